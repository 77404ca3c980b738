#!/usr/bin/env python3
"""Load-balancing scenario: 10 concurrent requests to the group `claude-sonnet-loadbalance-demo` (two primaries,
rpm 3 each) overflow into its fallback group.  Expected with the shipped config: 3 + 3 primary, 4 fallback."""
import argparse
import time

from _common import add_client_args, burst, distribution, make_client, say

GROUP = "claude-sonnet-loadbalance-demo"


def run_once(client, rl_exc, run_no=None):
    say(f"load balancing{'' if run_no is None else f' (run {run_no})'}: 10 requests -> {GROUP}", "blue")

    def show(r):
        if r.ok:
            spill = "claude-3-5-sonnet" in r.model_used
            say(f"req {r.request_id:2d} {'FALLBACK' if spill else 'PRIMARY '} {r.model_used:40s} {r.seconds:5.2f}s",
                "yellow" if spill else "green")
        else:
            say(f"req {r.request_id:2d} ERROR {r.error}", "red")

    recs = burst(client, rl_exc, GROUP, 10, 0.1, 30, show)
    ok = [r for r in recs if r.ok]
    dist = distribution(recs)
    print(f"\n  total {len(recs)}  ok {len(ok)}  failed {len(recs) - len(ok)}")
    for m, c in sorted(dist.items()):
        print(f"  {m:44s} {c:2d}  ({100.0 * c / max(1, len(ok)):5.1f}%)")
    print(f"  mean latency {sum(r.seconds for r in ok) / max(1, len(ok)):.2f}s")
    say("requests were spread over several deployments" if len(dist) > 1 else "only one deployment answered",
        "green" if len(dist) > 1 else "yellow")
    return {"ok": len(ok), "failed": len(recs) - len(ok), "distribution": dict(dist)}


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    add_client_args(ap)
    ap.add_argument("--loop", action="store_true")
    ap.add_argument("--interval", type=int, default=65, help="seconds between runs in --loop mode")
    a = ap.parse_args()
    client, rl_exc, close = make_client(a)
    try:
        if not a.loop:
            return run_once(client, rl_exc)
        n, totals = 0, {}
        while True:
            n += 1
            for k, v in run_once(client, rl_exc, n)["distribution"].items():
                totals[k] = totals.get(k, 0) + v
            say(f"cumulative after {n} runs: {totals}", "blue")
            time.sleep(a.interval)
    except KeyboardInterrupt:
        say("stopped", "yellow")
    finally:
        close()


if __name__ == "__main__":
    main()

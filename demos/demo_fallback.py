#!/usr/bin/env python3
"""Fallback scenario: 10 requests 50 ms apart to `claude-sonnet-fallback-demo` (rpm 3); the rest must be served by
its fallback group.  Expected with the shipped config: 3 primary, 7 fallback, 0 failed."""
import argparse

from _common import add_client_args, burst, distribution, make_client, say

GROUP = "claude-sonnet-fallback-demo"


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    add_client_args(ap)
    a = ap.parse_args()
    client, rl_exc, close = make_client(a)
    try:
        say(f"fallback: 10 requests -> {GROUP}", "blue")
        recs = burst(client, rl_exc, GROUP, 10, 0.05, 30)
        ok = [r for r in recs if r.ok]
        spilled = [r for r in ok if "3-5-sonnet" in r.model_used or "sonnet-3-5" in r.model_used]
        for r in sorted(recs, key=lambda r: r.request_id):
            tag = "FALLBACK" if r in spilled else ("PRIMARY " if r.ok else "FAILED  ")
            say(f"req {r.request_id:2d} {tag} {r.model_used or r.error}", "yellow" if r in spilled else ("green" if r.ok else "red"))
        print(f"\n  total {len(recs)}  ok {len(ok)}  failed {len(recs) - len(ok)}  "
              f"primary {len(ok) - len(spilled)}  fallback {len(spilled)}")
        for m, c in sorted(distribution(recs).items()):
            print(f"  {m:44s} {c:2d}")
        say(f"{len(spilled)} requests failed over to the fallback group" if spilled else "no fallback happened",
            "green" if spilled else "yellow")
        return {"primary": len(ok) - len(spilled), "fallback": len(spilled), "failed": len(recs) - len(ok)}
    finally:
        close()


if __name__ == "__main__":
    main()

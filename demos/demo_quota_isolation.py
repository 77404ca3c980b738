#!/usr/bin/env python3
"""Quota isolation: consumers A (rpm 3), B and C (rpm 10) each send 5 parallel requests to their own model group.
Expected with the shipped config: A 3/5 (2 rate limited), B 5/5, C 5/5 — a noisy tenant cannot starve the others."""
import argparse
import time
from concurrent.futures import ThreadPoolExecutor

from _common import QUESTIONS, add_client_args, make_client, one_request, say

CONSUMERS = {"A": ("consumer-a-model", "noisy", "consumer-a-key"), "B": ("consumer-b-model", "normal", "consumer-b-key"),
             "C": ("consumer-c-model", "normal", "consumer-c-key")}


def consumer_burst(args, name, n):
    group, _kind, key = CONSUMERS[name]
    client, rl_exc, close = make_client(args, api_key=key) if args.base_url else (args._shared[0], args._shared[1], None)
    with ThreadPoolExecutor(max_workers=n) as ex:
        recs = list(ex.map(lambda i: one_request(client, rl_exc, i, group, QUESTIONS[i % len(QUESTIONS)], 10), range(1, n + 1)))
    for r in recs:
        say(f"{name} | {'ok          ' if r.ok else ('RATE LIMITED' if r.rate_limited else 'FAILED      ')} | req {r.request_id} | {r.seconds:.2f}s",
            "green" if r.ok else "red")
    if close:
        close()
    return name, recs


def run_once(args, n=5):
    t0 = time.time()
    with ThreadPoolExecutor(max_workers=3) as ex:
        results = dict(ex.map(lambda c: consumer_burst(args, c, n), CONSUMERS))
    print(f"\n  finished in {time.time() - t0:.1f}s")
    print("  consumer | kind   | ok/total | rate limited | mean s")
    summary = {}
    for name, recs in sorted(results.items()):
        ok = [r for r in recs if r.ok]
        rl = sum(r.rate_limited for r in recs)
        summary[name] = (len(ok), rl)
        print(f"  {name:8s} | {CONSUMERS[name][1]:6s} | {len(ok)}/{len(recs)}      | {rl:12d} | {sum(r.seconds for r in ok) / max(1, len(ok)):.2f}")
    isolated = all(summary[c][0] >= 0.8 * n for c in ("B", "C"))
    say("isolation effective: B and C kept >= 80 % success" if isolated else "isolation NOT effective",
        "green" if isolated else "red")
    say("noisy consumer was rate limited" if summary["A"][1] else "noisy consumer was not limited", "blue")
    return summary


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    add_client_args(ap)
    ap.add_argument("--loop", action="store_true")
    ap.add_argument("--interval", type=int, default=65)
    a = ap.parse_args()
    a._shared = None
    closer = None
    if not a.base_url:                       # in-process: one router shared by the three consumers
        client, rl_exc, closer = make_client(a)
        a._shared = (client, rl_exc)
    try:
        if not a.loop:
            return run_once(a)
        while True:
            run_once(a)
            time.sleep(a.interval)
    except KeyboardInterrupt:
        say("stopped", "yellow")
    finally:
        if closer:
            closer()


if __name__ == "__main__":
    main()

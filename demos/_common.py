"""Shared harness of the three gateway demos: fire a burst of chat-completion requests
from threads, collect one record per request, print an end-of-run table.

Two client modes: `--base-url http://host:port` uses the real OpenAI SDK over HTTP against
`python rr_b200_server.py` (the reference's boundary); the default builds the router in-process."""
from __future__ import annotations

import argparse
import os
import sys
import threading
import time
from collections import Counter
from dataclasses import dataclass
from datetime import datetime
from typing import Callable, List, Optional

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

QUESTIONS = ["Summarise what a load balancer does.", "Why do rate limits exist?", "Name one use of a cache.",
             "What is a fallback route?", "Explain back-pressure in one line.", "What is a token bucket?",
             "Define tail latency.", "What does idempotent mean?", "What is a cooldown period?",
             "Why isolate tenants' quotas?"]
_C = {"green": "\033[92m", "yellow": "\033[93m", "red": "\033[91m", "blue": "\033[94m", "": ""}


def say(msg: str, color: str = "") -> None:
    stamp = datetime.now().strftime("%H:%M:%S.%f")[:-3]
    print(f"{_C[color]}[{stamp}] {msg}{chr(27) + '[0m' if color else ''}", flush=True)


@dataclass
class Record:
    request_id: int
    ok: bool
    model_used: str = ""
    seconds: float = 0.0
    error: str = ""
    rate_limited: bool = False


def add_client_args(ap: argparse.ArgumentParser) -> None:
    ap.add_argument("--config", default=os.path.join(ROOT, "config", "config.yaml"))
    ap.add_argument("--base-url", default=None, help="use the OpenAI SDK over HTTP against a running gateway")
    ap.add_argument("--stub", action="store_true", help="in-process mode: mock-completion backends")
    ap.add_argument("--spec", default=None, help="in-process mode: override every deployment's model spec")
    ap.add_argument("--max-tokens", type=int, default=16)


def make_client(args, api_key: str = "demo-key"):
    """-> (client with .chat.completions.create, RateLimitError class, closer)."""
    if args.base_url:
        import openai
        return openai.OpenAI(api_key=api_key, base_url=args.base_url), openai.RateLimitError, lambda: None
    from rr_b200 import OpenAI, RateLimitError, Router, load_config
    from rr_b200.server import build_backends
    cfg = load_config(args.config)
    router = Router(config=cfg, backends=build_backends(cfg, args.stub, args.spec, max_batch=16, ctx_max=512),
                    default_max_tokens=args.max_tokens)
    return OpenAI(router, api_key=api_key), RateLimitError, router.close


def one_request(client, rate_limit_exc, request_id: int, model: str, text: str, timeout: float) -> Record:
    t0 = time.time()
    try:
        resp = client.chat.completions.create(model=model, messages=[{"role": "user", "content": text}],
                                              timeout=timeout)
        return Record(request_id, True, getattr(resp, "model", "unknown"), round(time.time() - t0, 2))
    except rate_limit_exc as e:
        return Record(request_id, False, seconds=round(time.time() - t0, 2), error=str(e)[:80], rate_limited=True)
    except Exception as e:       # noqa: BLE001 — a demo reports, it does not crash
        return Record(request_id, False, seconds=round(time.time() - t0, 2), error=f"{type(e).__name__}: {e}"[:80])


def burst(client, rate_limit_exc, model: str, n: int, stagger_s: float, timeout: float,
          on_done: Optional[Callable[[Record], None]] = None) -> List[Record]:
    """n requests from n threads, `stagger_s` apart; returns records in completion order."""
    out: List[Record] = []
    lock = threading.Lock()

    def work(i):
        r = one_request(client, rate_limit_exc, i, model, QUESTIONS[(i - 1) % len(QUESTIONS)], timeout)
        with lock:
            out.append(r)
            if on_done:
                on_done(r)

    threads = []
    for i in range(1, n + 1):
        th = threading.Thread(target=work, args=(i,))
        th.start()
        threads.append(th)
        time.sleep(stagger_s)
    for th in threads:
        th.join()
    return out


def distribution(records: List[Record]) -> Counter:
    return Counter(r.model_used for r in records if r.ok)

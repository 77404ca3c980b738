"""CPU restatement of the request-router semantics — TEST INFRASTRUCTURE, NOT THE PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (sample-resilient-llm-inference_b200/) never does.

PARITY UNPINNED.  The reference's routing logic lives in the un-vendored dependency
`litellm[proxy]>=1.73.6.post1` (reference pyproject.toml:8), which is not installed, not in the
wheelhouse and cannot be fetched here; the reference ships no tests, golden vectors or fixtures.
This file restates (a) the configuration semantics visible in the reference tree
(reference config/config.yaml:35-108), (b) the observable contract recorded in the reference
README sample outputs (README.md:144,167-171 fallback 3+7; :194,206-213 load-balance 3+3+4;
:230,262-264 quota A 3/5, B 5/5, C 5/5) — pinned in tests/test_oracle_router.py — and (c) the
published behaviour of litellm's Router recalled from upstream (simple_shuffle.py, least_busy.py,
router.py::_pre_call_checks, cooldown handlers), each marked [UPSTREAM-RECALL].  The RNG is the
real CPython `random.Random` (Lib/random.py:242-250 _randbelow_with_getrandbits, :341-348 choice,
:454-489 choices), which is what litellm calls.

Semantics (one event at a time, in trace order — "serialised-trace semantics"):

ADMIT(group g, prompt tokens n, chain_start c, now_ms)
  chain = [g] + fallbacks[g]                           (config.yaml:105-108; single pass, in order)
  for pos >= c: pick(chain[pos]); first success wins; none -> RATE_LIMITED (HTTP 429,
  reference src/demo_quota_isolation.py:80).
pick(group):
  candidates = deployments of the group in model_list order (config.yaml:35-94)
  roll the per-minute window of every candidate (minute = now_ms // 60000): a fixed window that
    refills the rpm/tpm bucket to full at each minute boundary  [UPSTREAM-RECALL: per-minute
    counters in the router cache]
  healthy = candidates not in cooldown (now_ms >= cooldown_until)
  if enable_pre_call_checks (config.yaml:102): drop d with rpm >= 0 and req_count >= rpm
    (config.yaml:41) and d with tpm >= 0 and tok_count + n > tpm (config.yaml:42)
  simple-shuffle (config.yaml:101) [UPSTREAM-RECALL simple_shuffle.py]: if healthy[0] carries a
    weight: random.choices(range(len), weights=[w/sum(w)])[0]; else random.choice(healthy)
  least-busy [UPSTREAM-RECALL least_busy.py]: the first candidate (all of the group, config order)
    with minimum in-flight count; if it is not healthy: random.choice(healthy)
  round-robin (reference src/demo_account_sharding.py:335-343 `req_id % n`): healthy[k % len], k++
  split (ibid. `req_id < num_requests // 2`, generalised to n backends): request i of a burst of N declared by
    BURST(group, N) goes to healthy[min(len - 1, i * len // N)]
  random (ibid. `random.choice`): uniform pick, weights ignored
  debit: req_count += 1, tok_count += n, inflight += 1
DONE(deployment, completion tokens, now): inflight -= 1; tok_count += tokens (current window)
FAIL(deployment, now): inflight -= 1; per-minute fail_count += 1; if fail_count > allowed_fails
  (config.yaml:103): cooldown_until = now + cooldown_time (config.yaml:104)
"""
from __future__ import annotations

import random
from bisect import bisect as _bisect
from dataclasses import dataclass, field
from itertools import accumulate as _accumulate
from typing import Dict, List, Optional, Sequence

RR_OK, RR_RATE_LIMITED, RR_NO_GROUP = 0, 1, 2
STRATEGY_SIMPLE_SHUFFLE, STRATEGY_LEAST_BUSY, STRATEGY_ROUND_ROBIN, STRATEGY_SPLIT, STRATEGY_RANDOM = 0, 1, 2, 3, 4
EV_ADMIT, EV_DONE, EV_FAIL, EV_BURST = 0, 1, 2, 3


@dataclass
class Deployment:
    group: int
    rpm: int = -1
    tpm: int = -1
    weight: int = -1
    replica: int = 0
    # state
    window: int = -1
    req_count: int = 0
    tok_count: int = 0
    fail_window: int = -1
    fail_count: int = 0
    inflight: int = 0
    cooldown_until_ms: int = 0
    total_admitted: int = 0


@dataclass
class Settings:
    strategy: int = STRATEGY_SIMPLE_SHUFFLE
    enable_pre_call_checks: bool = True
    allowed_fails: int = 2
    cooldown_ms: int = 15000


@dataclass
class Event:
    type: int
    target: int
    tokens: int = 0
    chain_start: int = 0
    now_ms: int = 0


@dataclass
class Decision:
    status: int
    deployment: int
    served_group: int
    chain_pos: int

    def as_tuple(self):
        return (self.status, self.deployment, self.served_group, self.chain_pos)


class OracleRouter:
    def __init__(self, deployments: Sequence[Deployment], n_groups: int,
                 fallbacks: Dict[int, List[int]], settings: Settings, seed: int = 0):
        self.deps = [Deployment(d.group, d.rpm, d.tpm, d.weight, d.replica) for d in deployments]
        self.n_groups = n_groups
        self.fallbacks = {g: list(fallbacks.get(g, [])) for g in range(n_groups)}
        self.settings = settings
        self.rng = random.Random(seed)
        self.rr_next = [0] * n_groups
        self.burst_size = [0] * n_groups
        self.burst_pos = [0] * n_groups
        self.by_group: List[List[int]] = [[] for _ in range(n_groups)]
        for i, d in enumerate(self.deps):
            self.by_group[d.group].append(i)

    # -- helpers -----------------------------------------------------------------------------
    def _roll(self, d: Deployment, now_ms: int) -> None:
        minute = now_ms // 60000
        if d.window != minute:
            d.window = minute
            d.req_count = 0
            d.tok_count = 0

    def _pick(self, group: int, n_tokens: int, now_ms: int) -> Optional[int]:
        cands = self.by_group[group]
        for i in cands:
            self._roll(self.deps[i], now_ms)
        healthy = [i for i in cands if now_ms >= self.deps[i].cooldown_until_ms]
        if self.settings.enable_pre_call_checks:
            ok = []
            for i in healthy:
                d = self.deps[i]
                if d.rpm >= 0 and d.req_count >= d.rpm:
                    continue
                if d.tpm >= 0 and d.tok_count + n_tokens > d.tpm:
                    continue
                ok.append(i)
            healthy = ok
        if not healthy:
            return None
        st = self.settings.strategy
        if st == STRATEGY_SIMPLE_SHUFFLE:
            if self.deps[healthy[0]].weight >= 0:
                ws = [max(self.deps[i].weight, 0) for i in healthy]
                total = sum(ws)
                if total > 0:
                    norm = [w / total for w in ws]
                    # == random.choices(range(n), weights=norm)[0]   (Lib/random.py:454-489)
                    cum = list(_accumulate(norm))
                    tot = cum[-1] + 0.0
                    return healthy[_bisect(cum, self.rng.random() * tot, 0, len(healthy) - 1)]
            return healthy[self.rng._randbelow(len(healthy))]      # random.choice
        if st == STRATEGY_LEAST_BUSY:
            best, best_v = None, None
            for i in cands:
                v = self.deps[i].inflight
                if best_v is None or v < best_v:
                    best, best_v = i, v
            if best in healthy:
                return best
            return healthy[self.rng._randbelow(len(healthy))]
        if st == STRATEGY_ROUND_ROBIN:
            k = self.rr_next[group]
            self.rr_next[group] = k + 1
            return healthy[k % len(healthy)]
        if st == STRATEGY_SPLIT:
            n, i = self.burst_size[group], self.burst_pos[group]
            self.burst_pos[group] = i + 1
            idx = (i * len(healthy)) // n if n > 0 else 0
            return healthy[min(idx, len(healthy) - 1)]
        if st == STRATEGY_RANDOM:
            return healthy[self.rng._randbelow(len(healthy))]
        raise ValueError(st)

    # -- events ------------------------------------------------------------------------------
    def admit(self, group: int, n_tokens: int, chain_start: int, now_ms: int) -> Decision:
        if group < 0 or group >= self.n_groups:
            return Decision(RR_NO_GROUP, -1, -1, 0)
        chain = [group] + self.fallbacks[group]
        for pos in range(max(chain_start, 0), len(chain)):
            g = chain[pos]
            i = self._pick(g, n_tokens, now_ms)
            if i is not None:
                d = self.deps[i]
                d.req_count += 1
                d.tok_count += n_tokens
                d.inflight += 1
                d.total_admitted += 1
                return Decision(RR_OK, i, g, pos)
        return Decision(RR_RATE_LIMITED, -1, -1, 0)

    def done(self, dep: int, tokens: int, now_ms: int) -> Decision:
        if dep < 0 or dep >= len(self.deps):
            return Decision(RR_NO_GROUP, -1, -1, 0)
        d = self.deps[dep]
        d.inflight = max(0, d.inflight - 1)
        self._roll(d, now_ms)
        d.tok_count += tokens
        return Decision(RR_OK, dep, d.group, 0)

    def fail(self, dep: int, now_ms: int) -> Decision:
        if dep < 0 or dep >= len(self.deps):
            return Decision(RR_NO_GROUP, -1, -1, 0)
        d = self.deps[dep]
        d.inflight = max(0, d.inflight - 1)
        minute = now_ms // 60000
        if d.fail_window != minute:
            d.fail_window = minute
            d.fail_count = 0
        d.fail_count += 1
        cooled = 0
        if d.fail_count > self.settings.allowed_fails:
            d.cooldown_until_ms = now_ms + self.settings.cooldown_ms
            cooled = 1
        return Decision(RR_OK, dep, d.group, cooled)

    def process(self, events: Sequence[Event]) -> List[Decision]:
        out = []
        for e in events:
            if e.type == EV_ADMIT:
                out.append(self.admit(e.target, e.tokens, e.chain_start, e.now_ms))
            elif e.type == EV_DONE:
                out.append(self.done(e.target, e.tokens, e.now_ms))
            elif e.type == EV_FAIL:
                out.append(self.fail(e.target, e.now_ms))
            elif e.type == EV_BURST:
                if 0 <= e.target < self.n_groups:
                    self.burst_size[e.target], self.burst_pos[e.target] = e.tokens, 0
                    out.append(Decision(RR_OK, -1, e.target, 0))
                else:
                    out.append(Decision(RR_NO_GROUP, -1, -1, 0))
            else:
                out.append(Decision(RR_NO_GROUP, -1, -1, 0))
        return out

    def snapshot(self):
        return [(d.window, d.req_count, d.tok_count, d.fail_window, d.fail_count, d.inflight,
                 d.cooldown_until_ms, d.total_admitted) for d in self.deps]

"""Bit-exact parity of the admission kernel K1 (rr_router_process through the C-ABI) against the
CPU oracle on identical seeded event traces: decisions, counters, cooldowns and the RNG stream."""
import ctypes as C
import random

import pytest

from oracle import router as O
from oracle.scenarios import reference_topology, REF_GROUPS

pytestmark = pytest.mark.gpu


class DevRouter:
    def __init__(self, deps, n_groups, fallbacks, settings, seed):
        from rr_b200 import _lib
        self.L = _lib
        arr = (_lib.DeploymentDesc * len(deps))()
        for i, d in enumerate(deps):
            arr[i] = _lib.DeploymentDesc(d.group, d.rpm, d.tpm, d.weight, d.replica, 0)
        offs, flat = [0], []
        for g in range(n_groups):
            flat += fallbacks.get(g, [])
            offs.append(len(flat))
        st = _lib.RouterSettings(settings.strategy, int(settings.enable_pre_call_checks),
                                 settings.allowed_fails, settings.cooldown_ms, 0)
        self.h = C.c_void_p()
        self.n = len(deps)
        _lib.check(_lib.lib.rr_router_create(arr, len(deps), n_groups, (C.c_int32 * len(offs))(*offs),
                                             (C.c_int32 * max(1, len(flat)))(*flat), C.byref(st), seed, 0,
                                             C.byref(self.h)), "create")

    def process(self, events):
        L = self.L
        ev = (L.Event * len(events))()
        for i, e in enumerate(events):
            ev[i] = L.Event(e.type, e.target, e.tokens, e.chain_start, e.now_ms)
        out = (L.Decision * len(events))()
        L.check(L.lib.rr_router_process(self.h, ev, len(events), out), "process")
        return [(d.status, d.deployment, d.served_group, d.chain_pos) for d in out]

    def snapshot(self):
        L = self.L
        s = (L.DeploymentState * self.n)()
        L.check(L.lib.rr_router_snapshot(self.h, s), "snapshot")
        return [(x.window, x.req_count, x.tok_count, x.fail_window, x.fail_count, x.inflight,
                 x.cooldown_until_ms, x.total_admitted) for x in s]

    def close(self):
        self.L.lib.rr_router_destroy(self.h)


def _random_trace(rng, n_groups, n_deps, n, admit_only=False):
    now, ev = 0, []
    for _ in range(n):
        now += rng.choice([0, 1, 5, 50, 500, 20_000])
        t = O.EV_ADMIT if admit_only else rng.choices([0, 1, 2, 3], weights=[12, 6, 2, 1])[0]
        if t == O.EV_BURST:
            ev.append(O.Event(t, rng.randrange(-1, n_groups + 1), rng.randrange(0, 40), 0, now))
        elif t == O.EV_ADMIT:
            ev.append(O.Event(t, rng.randrange(-1, n_groups + 1), rng.randrange(0, 400),
                              rng.choice([0, 0, 0, 1, 2]), now))
        else:
            ev.append(O.Event(t, rng.randrange(-1, n_deps + 1), rng.randrange(0, 200), 0, now))
    return ev


def _topology(rng, weighted):
    n_groups = rng.randrange(1, 6)
    deps = []
    for g in range(n_groups):
        for _ in range(rng.randrange(1, 9)):
            deps.append(O.Deployment(group=g, rpm=rng.choice([-1, 2, 3, 10, 25]),
                                     tpm=rng.choice([-1, 500, 5000]),
                                     weight=rng.choice([1, 2, 3, 7]) if weighted else -1))
    rng.shuffle(deps)
    fbs = {g: rng.sample(range(n_groups), rng.randrange(0, min(3, n_groups) + 1)) for g in range(n_groups)}
    return deps, n_groups, fbs


@pytest.mark.parametrize("strategy", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("seed", [0, 1, 2**33 + 5])
def test_kernel_matches_oracle_on_random_traces(strategy, weighted, seed):
    rng = random.Random(seed * 31 + strategy * 7 + weighted)
    deps, ng, fbs = _topology(rng, weighted)
    st = O.Settings(strategy=strategy, enable_pre_call_checks=rng.random() < 0.8,
                    allowed_fails=rng.randrange(0, 3), cooldown_ms=rng.choice([1000, 15000]))
    orc = O.OracleRouter(deps, ng, fbs, st, seed=seed)
    dev = DevRouter(deps, ng, fbs, st, seed)
    try:
        for chunk in range(4):                      # several launches: state + RNG persist in HBM
            ev = _random_trace(rng, ng, len(deps), 700)
            want = [d.as_tuple() for d in orc.process(ev)]
            got = dev.process(ev)
            assert got == want, next((i, ev[i], g, w) for i, (g, w) in enumerate(zip(got, want)) if g != w)
            assert dev.snapshot() == orc.snapshot()
    finally:
        dev.close()


def test_rng_stream_survives_many_regenerations():
    """> 624 draws per launch and across launches: the twist runs on device several times."""
    deps = [O.Deployment(group=0) for _ in range(7)] + [O.Deployment(group=1, weight=w) for w in (5, 1, 3)]
    st = O.Settings()
    orc = O.OracleRouter(deps, 2, {}, st, seed=99)
    dev = DevRouter(deps, 2, {}, st, 99)
    try:
        for k in range(3):
            ev = [O.Event(O.EV_ADMIT, i % 2, 1, 0, i) for i in range(3000)]
            assert dev.process(ev) == [d.as_tuple() for d in orc.process(ev)]
    finally:
        dev.close()


def test_reference_topology_scenarios_on_device():
    deps, ng, fbs, st, info = reference_topology()
    dev = DevRouter(deps, ng, fbs, st, 0)
    try:
        g = REF_GROUPS["claude-sonnet-fallback-demo"]
        got = dev.process([O.Event(O.EV_ADMIT, g, 12, 0, 50 * i) for i in range(10)])
        assert sum(d[3] == 0 for d in got) == 3 and sum(d[3] == 1 for d in got) == 7
        out = {}
        for c in "abc":
            gg = REF_GROUPS[f"consumer-{c}-model"]
            out[c] = sum(d[0] == 0 for d in dev.process([O.Event(O.EV_ADMIT, gg, 20, 0, 10 * i) for i in range(5)]))
        assert out == {"a": 3, "b": 5, "c": 5}
    finally:
        dev.close()


def test_router_throughput_report(capsys):
    import time
    deps, ng, fbs, st, _ = reference_topology()
    dev = DevRouter(deps, ng, fbs, O.Settings(enable_pre_call_checks=False), 0)
    try:
        ev = [O.Event(O.EV_ADMIT, i % ng, 10, 0, i) for i in range(20000)]
        dev.process(ev[:100])
        t0 = time.perf_counter(); dev.process(ev); dt = time.perf_counter() - t0
        with capsys.disabled():
            print(f"\n[K1] {len(ev)/dt/1e6:.2f} M events/s, {dt/len(ev)*1e9:.0f} ns/event (incl. H2D/D2H)")
    finally:
        dev.close()

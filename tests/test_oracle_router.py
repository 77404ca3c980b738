"""CPU tests of the router oracle (oracle/router.py).

PARITY UNPINNED against litellm (not installable here); what IS pinned are the three behaviour-level
outcomes the reference records in its README sample outputs, driven through the reference's own
config/config.yaml schema (reference README.md:144,167-171; :194,206-213; :230,262-264)."""
import random

import pytest
from hypothesis import given, settings, strategies as st

from oracle import router as O
from oracle.scenarios import reference_router, REF_GROUPS


def _admit_n(r, group, n, t0=0, dt=50):
    return [r.admit(REF_GROUPS[group], 12, 0, t0 + i * dt) for i in range(n)]


def test_readme_fallback_scenario_3_primary_7_fallback():
    # reference src/demo_fallback.py:212-216: 10 requests 50 ms apart to claude-sonnet-fallback-demo
    r, deps = reference_router(seed=0)
    dec = _admit_n(r, "claude-sonnet-fallback-demo", 10, dt=50)
    assert all(d.status == O.RR_OK for d in dec)
    primary = sum(d.chain_pos == 0 for d in dec)
    fallback = sum(d.chain_pos == 1 for d in dec)
    assert (primary, fallback) == (3, 7)                      # README.md:167-171
    assert {deps[d.deployment]["model"] for d in dec if d.chain_pos == 1} == {
        "bedrock/us.anthropic.claude-3-5-sonnet-20241022-v2:0"}


@pytest.mark.parametrize("seed", range(8))
def test_readme_load_balancing_scenario_3_3_4(seed):
    # reference src/demo_load_balancing.py:195-199: 10 requests 100 ms apart
    r, deps = reference_router(seed=seed)
    dec = _admit_n(r, "claude-sonnet-loadbalance-demo", 10, dt=100)
    assert all(d.status == O.RR_OK for d in dec)
    by_model = {}
    for d in dec:
        by_model[deps[d.deployment]["model"]] = by_model.get(deps[d.deployment]["model"], 0) + 1
    assert sorted(by_model.values()) == [3, 3, 4]             # README.md:206-213
    assert by_model["bedrock/us.anthropic.claude-3-5-sonnet-20241022-v2:0"] == 4


def test_readme_quota_isolation_scenario():
    # reference src/demo_quota_isolation.py:135-139: 3 consumers x 5 parallel requests, no fallback
    r, deps = reference_router(seed=0)
    ok = {}
    limited = {}
    for i in range(5):
        for c in "abc":
            d = r.admit(REF_GROUPS[f"consumer-{c}-model"], 20, 0, 10 * i)
            ok[c] = ok.get(c, 0) + (d.status == O.RR_OK)
            limited[c] = limited.get(c, 0) + (d.status == O.RR_RATE_LIMITED)
    assert ok == {"a": 3, "b": 5, "c": 5}                     # README.md:262-264
    assert limited == {"a": 2, "b": 0, "c": 0}


def test_window_refills_at_minute_boundary():
    r, _ = reference_router(seed=0)
    g = REF_GROUPS["consumer-a-model"]
    assert [r.admit(g, 1, 0, 1000).status for _ in range(4)] == [0, 0, 0, 1]
    assert r.admit(g, 1, 0, 59_999).status == O.RR_RATE_LIMITED
    assert r.admit(g, 1, 0, 60_000).status == O.RR_OK


def test_tpm_bucket():
    deps = [O.Deployment(group=0, rpm=-1, tpm=100)]
    r = O.OracleRouter(deps, 1, {}, O.Settings(), seed=1)
    assert r.admit(0, 60, 0, 0).status == O.RR_OK
    assert r.admit(0, 41, 0, 1).status == O.RR_RATE_LIMITED     # 60 + 41 > 100
    assert r.admit(0, 40, 0, 2).status == O.RR_OK
    r.done(0, 10, 3)                                            # completion tokens count
    assert r.admit(0, 1, 0, 4).status == O.RR_RATE_LIMITED


def test_cooldown_after_allowed_fails():
    deps = [O.Deployment(group=0), O.Deployment(group=1)]
    r = O.OracleRouter(deps, 2, {0: [1]}, O.Settings(allowed_fails=2, cooldown_ms=15000), seed=0)
    for i in range(3):
        d = r.admit(0, 1, 0, i)
        assert d.deployment == 0
        f = r.fail(0, i)
    assert f.chain_pos == 1                                     # third failure > allowed_fails
    assert r.admit(0, 1, 0, 100).deployment == 1                # primary cooling -> fallback
    assert r.admit(0, 1, 0, 15_001).deployment == 1
    assert r.admit(0, 1, 0, 15_002).deployment == 0             # cooldown over (2 + 15000)


def test_weighted_pick_equals_python_random_choices():
    deps = [O.Deployment(group=0, weight=3), O.Deployment(group=0, weight=1)]
    r = O.OracleRouter(deps, 1, {}, O.Settings(), seed=42)
    ref = random.Random(42)
    for i in range(200):
        want = ref.choices(range(2), weights=[3 / 4, 1 / 4])[0]
        assert r.admit(0, 1, 0, i).deployment == want


def test_uniform_pick_equals_python_random_choice():
    deps = [O.Deployment(group=0) for _ in range(5)]
    r = O.OracleRouter(deps, 1, {}, O.Settings(), seed=7)
    ref = random.Random(7)
    for i in range(200):
        assert r.admit(0, 1, 0, i).deployment == ref.choice(range(5))


def test_least_busy_and_round_robin():
    deps = [O.Deployment(group=0) for _ in range(3)]
    r = O.OracleRouter(deps, 1, {}, O.Settings(strategy=O.STRATEGY_LEAST_BUSY), seed=0)
    assert [r.admit(0, 1, 0, i).deployment for i in range(6)] == [0, 1, 2, 0, 1, 2]
    r.done(2, 0, 10)
    assert r.admit(0, 1, 0, 11).deployment == 2
    r2 = O.OracleRouter(deps, 1, {}, O.Settings(strategy=O.STRATEGY_ROUND_ROBIN), seed=0)
    assert [r2.admit(0, 1, 0, i).deployment for i in range(5)] == [0, 1, 2, 0, 1]


def test_unknown_group_and_chain_start():
    r, _ = reference_router(seed=0)
    assert r.admit(99, 1, 0, 0).status == O.RR_NO_GROUP
    g = REF_GROUPS["claude-sonnet-fallback-demo"]
    d = r.admit(g, 1, 1, 0)                                      # skip the primary group
    assert d.status == O.RR_OK and d.chain_pos == 1


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2**40), strat=st.integers(0, 2),
       trace=st.lists(st.tuples(st.integers(0, 2), st.integers(0, 3), st.integers(0, 50),
                                st.integers(0, 30_000)), max_size=80))
def test_invariants(seed, strat, trace):
    """Counters never exceed their limits; in-flight never negative; decisions well formed."""
    deps = [O.Deployment(group=0, rpm=3, tpm=200, weight=2), O.Deployment(group=0, rpm=5, weight=1),
            O.Deployment(group=1, rpm=4), O.Deployment(group=2)]
    r = O.OracleRouter(deps, 3, {0: [1, 2], 1: [2]}, O.Settings(strategy=strat), seed=seed)
    now = 0
    for typ, tgt, tok, dt in trace:
        now += dt
        if typ == O.EV_ADMIT:
            d = r.admit(tgt % 3, tok, 0, now)
            assert d.status in (O.RR_OK, O.RR_RATE_LIMITED)
            if d.status == O.RR_OK:
                assert r.deps[d.deployment].group == d.served_group
        elif typ == O.EV_DONE:
            r.done(tgt, tok, now)
        else:
            r.fail(tgt, now)
        for dep in r.deps:
            assert dep.inflight >= 0
            if dep.rpm >= 0 and dep.window == now // 60000:
                assert dep.req_count <= dep.rpm


def test_split_and_random_strategies():
    # reference src/demo_account_sharding.py:335-343: split = first half to backend 0, second half to backend 1
    deps = [O.Deployment(group=0), O.Deployment(group=0)]
    r = O.OracleRouter(deps, 1, {}, O.Settings(strategy=O.STRATEGY_SPLIT), seed=0)
    r.process([O.Event(O.EV_BURST, 0, 10, 0, 0)])
    assert [r.admit(0, 1, 0, i).deployment for i in range(10)] == [0] * 5 + [1] * 5
    r3 = O.OracleRouter([O.Deployment(group=0) for _ in range(3)], 1, {}, O.Settings(strategy=O.STRATEGY_SPLIT), seed=0)
    r3.process([O.Event(O.EV_BURST, 0, 7, 0, 0)])
    assert [r3.admit(0, 1, 0, i).deployment for i in range(9)] == [0, 0, 0, 1, 1, 2, 2, 2, 2]   # overflow stays on the last
    rr = O.OracleRouter([O.Deployment(group=0, weight=9), O.Deployment(group=0, weight=1)], 1, {},
                        O.Settings(strategy=O.STRATEGY_RANDOM), seed=5)
    ref = random.Random(5)
    assert [rr.admit(0, 1, 0, i).deployment for i in range(50)] == [ref.choice([0, 1]) for _ in range(50)]

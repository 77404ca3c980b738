"""GPU: BASELINE.json configs [0], [2], [3], [4] as parity cases (reduced model size where a model is involved, full
router scale).  Router outcomes through Router.process / completion* (K1 on the device) are compared with the oracle
on the identical trace; model-backed cases check that every request is served and labelled correctly."""
import random
import threading
import time

import numpy as np
import pytest

from oracle import router as O

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("rpm_window")]


def _oracle_from_cfg(cfg, seed):
    deps = [O.Deployment(group=d.group, rpm=d.rpm, tpm=d.tpm, weight=d.weight, replica=d.gpu) for d in cfg.deployments]
    st = O.Settings(strategy=cfg.strategy_id, enable_pre_call_checks=cfg.enable_pre_call_checks,
                    allowed_fails=cfg.allowed_fails, cooldown_ms=int(round(cfg.cooldown_time * 1000)))
    return O.OracleRouter(deps, len(cfg.groups), cfg.fallbacks, st, seed=seed)


@pytest.mark.parametrize("weights", [(1, 1), (3, 1)])
def test_config0_two_mock_backends_weighted(weights):
    """configs[0]: 2 mock-completion backends, 10 concurrent requests, weighted routing."""
    from rr_b200 import Router, StubBackend
    ml = [{"model_name": "m", "litellm_params": {"model": f"b200/tiny@{i}", "gpu": i, "weight": w}} for i, w in enumerate(weights)]
    r = Router(model_list=ml, routing_strategy="simple-shuffle", backends={0: StubBackend(), 1: StubBackend()}, seed=0,
               clock=lambda: 10.0)
    out, lock = [], threading.Lock()

    def work(i):
        resp = r.completion(model="m", messages=[{"role": "user", "content": f"q{i}"}], timeout=30)
        with lock:
            out.append(resp._deployment)
    ths = [threading.Thread(target=work, args=(i,)) for i in range(10)]
    [t.start() for t in ths]; [t.join() for t in ths]
    ref = random.Random(0)
    tot = sum(weights)
    want = sorted(ref.choices(range(2), weights=[w / tot for w in weights])[0] for _ in range(10))
    assert sorted(out) == want                       # same multiset of picks as random.choices on the same seed
    r.close()


def test_config2_eight_backends_512_concurrent_quota_isolation():
    """configs[2]: 8 replicas, 512 concurrent requests, tpm/rpm buckets + per-team quota isolation: team A's limits
    reject ~25 % of its requests, teams B and C are untouched."""
    from rr_b200 import Router, StubBackend, build_config
    ml = []
    for team, gpus, rpm, tpm in (("team-a", (0, 1), 64, 40000), ("team-b", (2, 3, 4), 200, 10**9), ("team-c", (5, 6, 7), 200, 10**9)):
        for g in gpus:
            ml.append({"model_name": team, "litellm_params": {"model": f"b200/llama-3-8b@{team}", "gpu": g}, "rpm": rpm, "tpm": tpm})
    rs = {"routing_strategy": "least-busy", "enable_pre_call_checks": True, "allowed_fails": 2, "cooldown_time": 15}
    cfg = build_config(ml, rs)
    r = Router(config=cfg, backends={g: StubBackend() for g in range(8)}, seed=0)
    rng = random.Random(1)
    teams = ["team-a"] * 170 + ["team-b"] * 171 + ["team-c"] * 171
    rng.shuffle(teams)
    events = [(O.EV_ADMIT, cfg.group_index(t), 512, 0, 1000 + i // 8) for i, t in enumerate(teams)]
    got = r.process(events)
    orc = _oracle_from_cfg(cfg, 0)
    want = [d.as_tuple() for d in orc.process([O.Event(*e) for e in events])]
    assert got == want
    ok = {t: sum(1 for (s, *_), tt in zip(got, teams) if tt == t and s == 0) for t in set(teams)}
    assert ok["team-b"] == 171 and ok["team-c"] == 171
    assert ok["team-a"] == 128                      # 2 x rpm 64 (tpm 2 x 40000 allows 156): 42 of 170 (24.7 %) rejected
    per_gpu = np.bincount([d for s, d, *_ in got if s == 0], minlength=8)
    assert per_gpu[0] == per_gpu[1] == 64 and abs(int(per_gpu[2]) - int(per_gpu[4])) <= 1     # least-busy balances
    r.close()


def test_config3_primary_with_injected_failures_falls_back():
    """configs[3]: primary + fallback model on two replicas, 50 % injected primary failure (Bernoulli, seed 42):
    every request is served, failed ones by the fallback group; the primary cools down after allowed_fails."""
    from rr_b200 import Engine, EngineBackend, Router, SPECS, make_weights
    wa = make_weights(SPECS["tiny"], seed=1, sigma=0.05, device="cuda")
    wb = make_weights(SPECS["small96"], seed=2, sigma=0.03, device="cuda")      # Phi-3 family: head_dim 96
    primary = Engine(wa, max_batch=8, ctx_max=256, max_prefill_tokens=512, fail_prob=0.5, fail_seed=42)
    fallback = Engine(wb, max_batch=8, ctx_max=256, max_prefill_tokens=512)
    ml = [{"model_name": "primary", "litellm_params": {"model": "b200/tiny@llama-3-8b", "gpu": 0}},
          {"model_name": "backup", "litellm_params": {"model": "b200/small96@phi-3-mini", "gpu": 1}}]
    now = [100.0]
    r = Router(model_list=ml, routing_strategy="simple-shuffle", enable_pre_call_checks=True, allowed_fails=2, cooldown_time=15,
               fallbacks=[{"primary": ["backup"]}], backends={0: EngineBackend(primary), 1: EngineBackend(fallback)},
               seed=0, clock=lambda: now[0])
    prompt = list(range(3, 40))
    served = []
    for i in range(24):
        now[0] += 1.0
        resp = r.completion(model="primary", prompt_ids=prompt, max_tokens=4)
        assert len(resp._token_ids) == 4
        served.append("backup" if resp._fell_back else "primary")
        assert ("phi-3-mini" in resp.model) == resp._fell_back
    assert 3 <= served.count("backup") <= 23 and served.count("primary") >= 1
    snap = r.snapshot()
    assert snap[0]["fail_count"] >= 3 and snap[0]["cooldown_until_ms"] > 0          # cooldown was triggered
    assert snap[0]["inflight"] == 0 and snap[1]["inflight"] == 0
    primary.close(); fallback.close(); r.close()


def test_config4_mixed_fleet_weighted_3_to_1_soak():
    """configs[4]: 6 + 2 replicas, weighted 3:1 routing, 1000-request soak: picks bit-exact vs the oracle (and so
    vs random.choices), ~75/25 split, p99 TTFT reported from real (tiny-model) replicas."""
    from rr_b200 import Engine, EngineBackend, Router, SPECS, build_config, make_weights
    ml = [{"model_name": "fleet", "litellm_params": {"model": "b200/tiny@llama-3-8b", "gpu": g, "weight": 3}} for g in range(6)]
    ml += [{"model_name": "fleet", "litellm_params": {"model": "b200/tiny@mistral-7b", "gpu": g, "weight": 1}} for g in (6, 7)]
    cfg = build_config(ml, {"routing_strategy": "simple-shuffle"})
    w = make_weights(SPECS["tiny"], seed=3, sigma=0.05, device="cuda")
    engines = [Engine(w, max_batch=16, ctx_max=128, max_prefill_tokens=1024) for _ in range(2)]      # llama pool, mistral pool
    backends = {g: EngineBackend(engines[0 if g < 6 else 1]) for g in range(8)}
    r = Router(config=cfg, backends=backends, seed=0)
    prompts = [list(np.random.RandomState(i).randint(3, 1000, size=24)) for i in range(1000)]
    t0 = time.perf_counter()
    outs = []
    for c in range(0, 1000, 50):                                   # 20 bursts of 50
        outs += r.completion_batch("fleet", prompts[c:c + 50], 3)
    wall = time.perf_counter() - t0
    assert all(not isinstance(o, Exception) and len(o._token_ids) == 3 for o in outs)
    orc = _oracle_from_cfg(cfg, 0)
    want = [orc.admit(0, 24, 0, 0).deployment for _ in range(1000)]
    assert [o._deployment for o in outs] == want
    share_llama = sum(o._deployment < 6 for o in outs) / 1000.0
    assert abs(share_llama - 0.9) < 0.03                           # 6*3 : 2*1 = 18 : 2
    ttft = np.array([o._ttft_s for o in outs]) * 1e3
    print(f"\\n[config 4 soak] 1000 requests in {wall:.2f}s; p50 TTFT {np.percentile(ttft, 50):.2f} ms, p99 {np.percentile(ttft, 99):.2f} ms; "
          f"llama share {share_llama:.3f}")
    for e in engines:
        e.close()
    r.close()

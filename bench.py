#!/usr/bin/env python
"""bench.py — headline benchmark of the request-router hot path (BASELINE.json):
completions/sec and p50 TTFT, 512-in / 128-out, Llama-3-8B shape, one replica per GPU,
64 concurrent requests per replica, least-busy routing.

A "step" = one closed burst of (64 x n_gpus) chat-completion requests through the whole path:
K1 admission (rpm/tpm debit + least-busy pick) -> per-GPU engine: chunked prefill + 127 decode steps
(CUDA graph) -> DONE events.  Launch:  python bench.py --gpus N --steps K --warmup W
(N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...).

  value      completions/s from device time only: CUDA events on each engine's stream around every
             prefill chunk and every decode step of the timed steps (inputs resident), max over ranks.
  e2e        the same metric by host wall clock through the public API (Router.completion_batch /
             rr_router_process + rr_engine_submit/wait): prompts start in pinned HOST memory, admission
             events H2D, decisions D2H, prompt ids H2D, generated tokens D2H all inside the timed region.
  roofline   decode step (one CUDA-graph launch = the dominant unit, ~78 % of step time): algorithmic bytes
             (weights streamed once + KV read) / mean CUDA-event step time, against MEASURED_PEAKS.json.
  cpu_baseline / --impl reference: the CPU restatement (oracle/) of the same path timed on the box's host
             cores on a bounded sample (the reference's own router, litellm, cannot be installed:
             BASELINE.md §2; there is no baseline/_ref).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "completions/sec (512-in/128-out, Llama-3-8B, 64 concurrent per GPU)"
# DRAM traffic of one decode step (64 rows, ctx 577) from the committed ncu --set full captures: per layer
# qkv 50.9 + o 34.1 + gate/up 238.6 + down 122.9 + attention 159.1 MB, x32, + lm_head 1.08 GB (algorithmic: 19.84 GB)
DECODE_STEP_DRAM_BYTES_NCU = int(32 * (50.91 + 34.12 + 238.55 + 122.92 + 159.14) * 1e6 + 1.083e9)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--concurrency", type=int, default=64, help="concurrent requests per GPU")
    ap.add_argument("--prompt-len", type=int, default=512)
    ap.add_argument("--max-new", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def make_prompts(n_req: int, prompt_len: int, vocab: int, pinned: bool):
    """int32[prompt_len] ids per request, torch.Generator().manual_seed(1234 + req_id), uniform in [0, vocab)."""
    import torch
    buf = torch.empty((n_req, prompt_len), dtype=torch.int32, pin_memory=pinned)
    for r in range(n_req):
        g = torch.Generator().manual_seed(1234 + r)
        buf[r] = torch.randint(0, vocab, (prompt_len,), generator=g, dtype=torch.int32)
    return buf


def cpu_port_sample(spec, prompt_len, max_new, concurrency, n_gpus, layers_timed=4, decode_steps_timed=2, threads=None):
    """CPU restatement of the path on a bounded sample; returns (completions/s, description, cores).

    Router: oracle admission of the full burst (all requests).  Token generation: fp32 torch CPU forward of
    ONE request at the exact Llama-3-8B layer shape; `layers_timed` of the identical layers and
    `decode_steps_timed` of the decode steps are timed and scaled to n_layers / (max_new - 1)."""
    import torch
    from oracle import router as O
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    cores = threads or min(avail, 64)          # beyond ~64 threads the fp32 GEMMs of one request stop scaling
    torch.set_num_threads(cores)
    t0 = time.perf_counter()
    deps = [O.Deployment(group=0, replica=r) for r in range(n_gpus)]
    orc = O.OracleRouter(deps, 1, {}, O.Settings(strategy=O.STRATEGY_LEAST_BUSY), seed=0)
    n_req = concurrency * n_gpus
    dec = orc.process([O.Event(O.EV_ADMIT, 0, prompt_len, 0, 0) for _ in range(n_req)])
    orc.process([O.Event(O.EV_DONE, d.deployment, max_new, 0, 1) for d in dec])
    t_router = time.perf_counter() - t0

    H, KV, D, hid, inter = spec.n_heads, spec.n_kv_heads, spec.head_dim, spec.hidden, spec.inter
    g = torch.Generator().manual_seed(0)
    mk = lambda *s: torch.empty(*s).normal_(0, 0.02, generator=g)
    layers = [dict(wqkv=mk((H + 2 * KV) * D, hid), wo=mk(hid, H * D), wgu=mk(2 * inter, hid), wdown=mk(hid, inter))
              for _ in range(layers_timed)]
    lm_head = mk(spec.vocab, hid)

    def layer_fwd(x, L, kc, vc, pos0):
        T = x.shape[0]
        h = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + spec.rms_eps)
        qkv = h @ L["wqkv"].t()
        q = qkv[:, : H * D].view(T, H, D)
        k = qkv[:, H * D:(H + KV) * D].view(T, KV, D)
        v = qkv[:, (H + KV) * D:].view(T, KV, D)
        kc[pos0:pos0 + T] = k; vc[pos0:pos0 + T] = v
        kk = kc[: pos0 + T].repeat_interleave(H // KV, 1); vv = vc[: pos0 + T].repeat_interleave(H // KV, 1)
        a = torch.nn.functional.scaled_dot_product_attention(q.transpose(0, 1), kk.transpose(0, 1), vv.transpose(0, 1),
                                                             is_causal=(T > 1))
        x = x + a.transpose(0, 1).reshape(T, H * D) @ L["wo"].t()
        h = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + spec.rms_eps)
        gu = h @ L["wgu"].t()
        return x + (torch.nn.functional.silu(gu[:, :inter]) * gu[:, inter:]) @ L["wdown"].t()

    with torch.no_grad():
        x = torch.randn(prompt_len, hid, generator=g)
        caches = [(torch.zeros(prompt_len + max_new, KV, D), torch.zeros(prompt_len + max_new, KV, D)) for _ in layers]
        t0 = time.perf_counter()
        for L, (kc, vc) in zip(layers, caches):
            x = layer_fwd(x, L, kc, vc, 0)
        (x[-1:] @ lm_head.t()).argmax()
        t_prefill_layers = time.perf_counter() - t0
        t0 = time.perf_counter()
        for s in range(decode_steps_timed):
            y = x[-1:].clone()
            for L, (kc, vc) in zip(layers, caches):
                y = layer_fwd(y, L, kc, vc, prompt_len + s)
            (y @ lm_head.t()).argmax()
        t_decode_step = (time.perf_counter() - t0) / decode_steps_timed
    scale = spec.n_layers / layers_timed
    t_completion = t_prefill_layers * scale + (max_new - 1) * t_decode_step * scale
    # one request at a time on all cores: completions/s of the whole burst
    value = n_req / (t_router + n_req * t_completion)
    sample = (f"router: oracle admission of all {n_req} requests ({t_router*1e3:.2f} ms); generation: 1 of {n_req} requests, "
              f"{layers_timed} of {spec.n_layers} identical layers and {decode_steps_timed} of {max_new - 1} decode steps timed "
              f"(fp32 torch CPU, {cores} threads) and scaled: {t_completion:.1f} s per completion")
    return value, sample, cores


# ------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    """CPU restatement of the path (the reference's litellm.Router cannot be installed: BASELINE.md §2)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rr_b200.models import resolve_spec
    spec = resolve_spec(args.model)
    vals = []
    sample, cores = "", 0
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        v, sample, cores = cpu_port_sample(spec, args.prompt_len, args.max_new, args.concurrency, args.gpus,
                                           layers_timed=2, decode_steps_timed=2)
        if i >= args.warmup:
            vals.append(v)
    value = len(vals) / sum(1.0 / v for v in vals)
    n_req = args.concurrency * args.gpus
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "completions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_req / value, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.model}: {n_req} concurrent {args.prompt_len}-in/{args.max_new}-out, least-busy "
                                   "(CPU restatement; litellm is not installable here)"},
            "cpu_baseline": {"value": value, "unit": "completions/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "completions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from rr_b200.models import resolve_spec, make_weights, broadcast_weights
    from rr_b200.engine import Engine
    from rr_b200.router import Router, EngineBackend, RateLimitError
    from rr_b200 import parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec = resolve_spec(args.model)
    C, P, M = args.concurrency, args.prompt_len, args.max_new
    ctx_max = ((P + M + 63) // 64) * 64

    # ---- start-up: seeded weights on rank 0, NCCL broadcast to every replica (the one collective)
    t0 = time.perf_counter()
    w = make_weights(spec, seed=0, sigma=0.02, device=f"cuda:{local}", allocate_only=(rank != 0))
    bcast_s = None
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        tb = time.perf_counter()
        broadcast_weights(w, src=0)
        torch.cuda.synchronize()
        bcast_s = time.perf_counter() - tb
    torch.cuda.synchronize()
    init_s = time.perf_counter() - t0
    eng = Engine(w, device=local, max_batch=C, ctx_max=ctx_max, max_prefill_tokens=8192, use_cuda_graph=True)
    backend = EngineBackend(eng)

    # ---- router (rank 0 owns the device-resident state): one model group, one deployment per GPU
    model_list = [{"model_name": args.model, "litellm_params": {"model": f"b200/{args.model}", "gpu": r},
                   "rpm": 1_000_000, "tpm": 2_000_000_000} for r in range(world)]
    router = None
    if rank == 0:
        router = Router(model_list=model_list, routing_strategy="least-busy", enable_pre_call_checks=True,
                        allowed_fails=2, cooldown_time=15, backends={0: backend}, seed=0, device=local)
    n_req = C * world
    prompts = make_prompts(n_req, P, spec.vocab, pinned=True)      # host-resident inputs (pinned)
    prompts_np = prompts.numpy()

    ttfts, lat = [], []

    def one_step(collect: bool):
        """One closed burst through the public path.  Returns nothing; per-request timings collected."""
        if world == 1:
            out = router.completion_batch(args.model, [prompts_np[i] for i in range(n_req)], M)
            for r in out:
                if isinstance(r, Exception):
                    raise r
                assert len(r._token_ids) == M
                if collect:
                    ttfts.append(r._ttft_s); lat.append(r._latency_s)
        else:
            dec = None
            if rank == 0:
                now = router.now_ms()
                dec = router.process([(0, 0, P, 0, now)] * n_req)
            mine = parallel.scatter_assignments(dec, n_req, world, rank, device=torch.device("cuda", local),
                                                replica_of=[d["litellm_params"]["gpu"] for d in model_list])
            ids = np.ascontiguousarray(prompts_np[mine]).reshape(-1)
            start = np.arange(0, (len(mine) + 1) * P, P, dtype=np.int32)
            recs, _ = eng.run_batch(ids, start, M) if len(mine) else ([], None)
            assert all(r.status == 0 and len(r.tokens) == M for r in recs)
            if collect:
                ttfts.extend(r.ttft for r in recs); lat.extend(r.latency for r in recs)
            parallel.gather_done(len(mine), world, rank, device=torch.device("cuda", local))
            if rank == 0:
                now = router.now_ms()
                router.process([(1, d[1], M, 0, now) for d in dec if d[0] == 0])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step(False)
    sync_all()
    eng.reset_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync_all()
    t_start = time.perf_counter()
    for _ in range(args.steps):
        one_step(True)
    sync_all()
    wall = time.perf_counter() - t_start
    clocks = sampler.stop() if rank == 0 else None
    st = eng.stats()
    dev_s = (st["prefill_ms_total"] + st["decode_ms_total"]) / 1e3

    # ---- max over ranks
    vals = torch.tensor([wall, dev_s, st["decode_ms_total"], float(st["decode_steps"]), st["prefill_ms_total"],
                         float(st["kernel_launches"]), float(st["h2d_bytes"]), float(st["d2h_bytes"])],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t_all = torch.tensor(sorted(ttfts) + [float("nan")] * (C * args.steps - len(ttfts)), dtype=torch.float64, device="cuda")
        gathered = [torch.empty_like(t_all) for _ in range(world)]
        dist.all_gather(gathered, t_all)
        all_ttft = torch.cat(gathered).cpu().numpy()
        all_ttft = all_ttft[~np.isnan(all_ttft)]
    else:
        mx, sm = vals, vals
        all_ttft = np.asarray(ttfts)
    wall_max, dev_max = mx[0].item(), mx[1].item()
    K = args.steps
    total_completions = n_req * K

    if rank == 0:
        hbm_peak, tf_peak, peak_src = load_peaks()
        # roofline of the dominant unit: the decode step (one CUDA-graph launch)
        steps_per_burst = M - 1
        mean_ctx = P + (1 + steps_per_burst) / 2.0                      # ctx at decode step j = P + j
        bytes_step = spec.weight_bytes_per_decode_step + C * mean_ctx * spec.kv_bytes_per_token
        dec_ms = mx[2].item() / max(1.0, mx[3].item())
        achieved = bytes_step / (dec_ms * 1e-3) / 1e9
        pf_tokens = C * P * K
        pf_tflops = spec.prefill_flops_per_token * pf_tokens / (mx[4].item() * 1e-3) / 1e12 if mx[4].item() > 0 else None
        line = {
            "metric": METRIC, "value": total_completions / dev_max, "unit": "completions/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": 1e3 * dev_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.model} shape, seeded random weights, {C} concurrent x {world} GPU(s), "
                                   f"{P}-in/{M}-out greedy, least-busy routing over {world} replica(s)",
                       "global_batch": n_req, "prompt_len": P, "max_new": M, "parallelism": f"replicas x{world} (dp{world})",
                       "l2": "inputs larger than L2: 15 GB of weights + 4.8 GB of KV are re-read every decode step (L2 = 126 MB)",
                       "timing": "value: CUDA events on each engine stream around every prefill chunk / decode step, max over ranks"},
            "p50_ttft_ms": float(np.percentile(all_ttft, 50) * 1e3), "p99_ttft_ms": float(np.percentile(all_ttft, 99) * 1e3),
            "tokens_per_s": total_completions * M / dev_max,
            "e2e": {"value": total_completions / wall_max, "unit": "completions/s", "ms_per_step": 1e3 * wall_max / K,
                    "h2d_bytes_per_step": int(sm[6].item() / K + 24 * n_req), "d2h_bytes_per_step": int(sm[7].item() / K + 16 * n_req),
                    "api": "Router.completion_batch -> rr_router_process + rr_engine_submit/rr_engine_wait (host buffers)"},
            "gpu_launches": int(sm[5].item() + 2 * K),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": DECODE_STEP_DRAM_BYTES_NCU if (args.model == "llama-3-8b" and C == 64 and P == 512 and M == 128) else None,
                         "traffic_source": "sum over the step's kernels of dram__bytes_read+write from ncu --set full (profiles/r01_ncu_full_summary.txt)",
                         "kernel": "decode step = 1 CUDA-graph launch (196 kernels with PDL edges; dominant: gemm_mlp_tcgen05<64> / gemm_bf16_tcgen05<64,*> weight streams)",
                         "bytes_per_launch": bytes_step, "ms_per_launch": dec_ms, "peak_source": peak_src},
            "prefill": {"tflops": pf_tflops, "peak_tflops_sustained": tf_peak, "frac": (pf_tflops / tf_peak) if pf_tflops else None,
                        "ms_per_burst": mx[4].item() / K},
            "clocks": clocks, "init_s": init_s, "weight_broadcast_s": bcast_s,
        }
        # K1 (router kernel) is latency-bound, not roofline-bound: report time per event of a full trace
        try:
            line["router"] = router_kernel_timing(router)
        except Exception as ex:                                   # never lose the headline line over the side measurement
            line["router"] = {"error": repr(ex)}
        if world == 1 and not args.no_cpu_baseline:
            eng.close()
            v, sample, cores = cpu_port_sample(spec, P, M, C, world)
            line["cpu_baseline"] = {"value": v, "unit": "completions/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def router_kernel_timing(router, n_events: int = 2048, reps: int = 5):
    """K1 alone: one launch over a trace of ADMIT + DONE pairs on the bench router (device buffers resident,
    CUDA events on the launch stream) and the same trace through the host entry point (H2D + launch + D2H + sync)."""
    import ctypes as C
    import torch
    from rr_b200 import _lib
    n_dep = len(router.cfg.deployments)
    now = router.now_ms()
    ev = (_lib.Event * n_events)()
    for i in range(0, n_events, 2):
        ev[i] = _lib.Event(0, 0, 512, 0, now)                    # ADMIT to group 0
        ev[i + 1] = _lib.Event(1, (i // 2) % n_dep, 128, 0, now)   # DONE on some deployment (keeps in-flight bounded)
    out = (_lib.Decision * n_events)()
    host = torch.frombuffer(bytearray(bytes(ev)), dtype=torch.uint8)
    d_ev = host.cuda()
    d_out = torch.empty(n_events * C.sizeof(_lib.Decision), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    def dev_once():
        _lib.check(_lib.lib.rr_router_process_device(router._h, C.c_void_p(d_ev.data_ptr()), n_events,
                                                     C.c_void_p(d_out.data_ptr()), C.c_void_p(stream.cuda_stream)))
    for _ in range(2):
        dev_once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        dev_once()
    e1.record(stream)
    torch.cuda.synchronize()
    dev_ns = e0.elapsed_time(e1) * 1e6 / (reps * n_events)
    t0 = time.perf_counter()
    for _ in range(reps):
        _lib.check(_lib.lib.rr_router_process(router._h, ev, n_events, out))
    host_ns = (time.perf_counter() - t0) * 1e9 / (reps * n_events)
    return {"events_per_launch": n_events, "ns_per_event_device": dev_ns, "events_per_s_device": 1e9 / dev_ns,
            "ns_per_event_host_api": host_ns, "bound": "latency (one warp walks the trace in order; lanes = candidate deployments)"}


if __name__ == "__main__":
    main()

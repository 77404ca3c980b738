"""Per-item timeline of sample CTAs (0, 37, 74, 111) inside the fused decode MLP kernel (gemm_mlp_tcgen05) of one middle
layer of a replayed decode step (rr_debug_trace_detail): where a CTA waits -- PDL dependency, ready[] counters, first
operands of an item, accumulator hand-over to the epilogue."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rr_b200 import _lib
from rr_b200.models import SPECS, make_weights
from rr_b200.engine import Engine
NAMES = {1: "gemm", 3: "attn_dec", 5: "norm", 8: "embed", 9: "argmax",
         50: "producer: first weight requests next", 51: "producer: PDL dependency resolved", 52: "producer: down item, wait ready[]",
         53: "producer: ready[] seen", 54: "MMA: first operands of a gate/up item landed", 55: "MMA: first operands of a down item landed",
         56: "epilogue: gate/up accumulator ready", 57: "epilogue: down accumulator ready", 58: "epilogue: fence done, ready[] incremented",
         70: "silu epi: past entry barrier", 71: "silu epi: tcgen05.ld done", 72: "silu epi: past exchange barrier", 73: "silu epi: loop done",
         60: "epilogue: thread 0 issued its stores", 61: "epilogue: all 128 threads past their stores"}
spec = SPECS[sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"]
w = make_weights(spec, seed=0, device="cuda")
eng = Engine(w, max_batch=64, ctx_max=640, max_prefill_tokens=8192)
ids = np.random.RandomState(0).randint(0, spec.vocab, size=(64, 512)).astype(np.int32)
start = np.arange(0, 64 * 512 + 1, 512, dtype=np.int32)
eng.run_batch(ids.reshape(-1), start, 8)
N = 60000
_lib.check(_lib.lib.rr_debug_trace_start(N))
_lib.check(_lib.lib.rr_debug_trace_detail(1))
eng.run_batch(ids.reshape(-1), start, 3)
buf = (C.c_uint64 * (4 * N))(); n = C.c_int32()
_lib.check(_lib.lib.rr_debug_trace_stop(buf, N, C.byref(n)))
_lib.check(_lib.lib.rr_debug_trace_detail(0))
a = np.frombuffer(buf, dtype=np.uint64)[: 4 * n.value].reshape(-1, 4).astype(np.int64)
a = a[a[:, 0] != 0]
a = a[np.argsort(a[:, 1], kind="stable")]
kid = a[:, 0] & 0xFF
cta = a[:, 0] >> 8
det = (kid >= 50) & (kid < 80)
d = a[det]; dk = kid[det]; dc = cta[det]
t0, t1 = d[:, 1].min(), d[:, 1].max()
# CTA-0 records of the kernels around the marked launch (the fixed slots keep the LAST fused-MLP launch of the run)
ker = a[kid < 50]; kk = kid[kid < 50]
sel = (ker[:, 3] >= t0 - 60000) & (ker[:, 1] <= t1 + 30000)
rows = [(int(r[1]), f"{NAMES.get(int(k), k):10s} start; dep resolved +{(r[2] - r[1]) / 1e3:.2f}; end +{(r[3] - r[1]) / 1e3:.2f}") for r, k in zip(ker[sel], kk[sel])]
only = int(sys.argv[2]) if len(sys.argv) > 2 else None
for r, k, c in zip(d, dk, dc):
    if only is not None and int(c) != only: continue
    rows.append((int(r[1]), f"cta {int(c):3d}  {NAMES.get(int(k), k)}"))
for t, txt in sorted(rows):
    print(f"  {(t - t0) / 1e3:8.2f} us   {txt}")
eng.close()

"""Per-phase timeline of CTAs 0 / 37 / 74 / 111 of one plain decode projection (selected by its split-K factor: 3 = QKV,
4 = O, 1 = lm_head for Llama-3-8B) in the last layer of a replayed decode step (fixed-slot marks)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rr_b200 import _lib
from rr_b200.models import SPECS, make_weights
from rr_b200.engine import Engine
NAMES = {1: "gemm", 3: "attn_dec", 5: "norm", 8: "embed", 9: "argmax",
         100: "producer: CTA running, first weight requests next", 101: "producer: PDL dependency resolved", 102: "producer: last load issued",
         103: "MMA: first operands landed", 104: "epilogue: accumulator ready", 105: "epilogue: tile handed to TMA", 106: "CTA done"}
splits = int(sys.argv[1]) if len(sys.argv) > 1 else 3
spec = SPECS["llama-3-8b"]
w = make_weights(spec, seed=0, device="cuda")
eng = Engine(w, max_batch=64, ctx_max=640, max_prefill_tokens=8192)
ids = np.random.RandomState(0).randint(0, spec.vocab, size=(64, 512)).astype(np.int32)
start = np.arange(0, 64 * 512 + 1, 512, dtype=np.int32)
eng.run_batch(ids.reshape(-1), start, 8)
N = 60000
_lib.check(_lib.lib.rr_debug_trace_start(N))
_lib.check(_lib.lib.rr_debug_trace_detail(10 + splits))
eng.run_batch(ids.reshape(-1), start, 3)
buf = (C.c_uint64 * (4 * N))(); n = C.c_int32()
_lib.check(_lib.lib.rr_debug_trace_stop(buf, N, C.byref(n)))
_lib.check(_lib.lib.rr_debug_trace_detail(0))
a = np.frombuffer(buf, dtype=np.uint64)[: 4 * n.value].reshape(-1, 4).astype(np.int64)
a = a[a[:, 0] != 0]
a = a[np.argsort(a[:, 1], kind="stable")]
kid = a[:, 0] & 0xFF
smp = a[:, 0] >> 8
det = (kid >= 100) & (kid < 110)
d = a[det]; dk = kid[det]; dc = smp[det]
t0, t1 = d[:, 1].min(), d[:, 1].max()
ker = a[kid < 50]; kk = kid[kid < 50]
sel = (ker[:, 3] >= t0 - 40000) & (ker[:, 1] <= t1 + 20000)
rows = [(int(r[1]), f"{NAMES.get(int(k), k):10s} start; dep resolved +{(r[2] - r[1]) / 1e3:.2f}; end +{(r[3] - r[1]) / 1e3:.2f}") for r, k in zip(ker[sel], kk[sel])]
for r, k, c in zip(d, dk, dc):
    rows.append((int(r[1]), f"cta {(int(c) - 8) * 37:3d}  {NAMES.get(int(k), k)}"))
for t, txt in sorted(rows):
    print(f"  {(t - t0) / 1e3:8.2f} us   {txt}")
eng.close()

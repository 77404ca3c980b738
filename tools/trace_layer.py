"""In-graph timeline of the persistent decode layer kernel (rr_layer.cu): CTA-0 start / dependency / end stamps of every
launch plus per-item phase marks of CTAs 0, 37, 74, 111 (rr_debug_trace_*).  Prints one middle layer of the last step."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rr_b200 import _lib
from rr_b200.models import SPECS, make_weights
from rr_b200.engine import Engine
NAMES = {1: "layer/gemm", 3: "attn_dec", 5: "norm", 8: "embed", 9: "argmax", 12: "O item done", 13: "GU item done", 14: "DOWN item done", 15: "REDUCE item done", 16: "NEXT item done",
         20: "acc ready O", 21: "acc ready GU", 22: "acc ready DOWN", 23: "acc ready NEXT", 30: "O planes stored", 31: "O arrive done",
         32: "O peers seen", 33: "O reduce done", 34: "RED begin", 35: "RED planes seen", 36: "RED reduce done", 37: "rinv ctr seen",
         38: "rinv staged", 40: "prod dep O", 41: "prod dep GU", 42: "prod dep DOWN", 43: "prod dep NEXT"}
spec = SPECS[sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"]
w = make_weights(spec, seed=0, device="cuda")
eng = Engine(w, max_batch=64, ctx_max=640, max_prefill_tokens=8192, fuse_layer=True)
ids = np.random.RandomState(0).randint(0, spec.vocab, size=(64, 512)).astype(np.int32)
start = np.arange(0, 64 * 512 + 1, 512, dtype=np.int32)
eng.run_batch(ids.reshape(-1), start, 8)
N = 40000
_lib.check(_lib.lib.rr_debug_trace_start(N))
_lib.check(_lib.lib.rr_debug_trace_detail(1))
eng.run_batch(ids.reshape(-1), start, 4)
buf = (C.c_uint64 * (4 * N))(); n = C.c_int32()
_lib.check(_lib.lib.rr_debug_trace_stop(buf, N, C.byref(n)))
a = np.frombuffer(buf, dtype=np.uint64)[: 4 * n.value].reshape(-1, 4).astype(np.int64)
a = a[np.argsort(a[:, 1], kind="stable")]
kid = a[:, 0] & 0xFF
cta = a[:, 0] >> 8
# last decode step: from the last embed kernel on
emb = np.nonzero(kid == 8)[0]
s0 = emb[-1]
step = a[s0:]; skid = kid[s0:]; scta = cta[s0:]
t0 = step[0, 1]
print(f"{n.value} records; last decode step: {(step[:, 3].max() - t0) / 1e3:.1f} us")
attn = np.nonzero(skid == 3)[0]
L = len(attn)
lo = attn[min(10, L - 2)]; hi = attn[min(11, L - 1)]
print(f"layer 10: attention start -> next attention start = {(step[hi, 1] - step[lo, 1]) / 1e3:.1f} us")
base = step[lo, 1]
for i in range(lo, hi + 1):
    k = int(skid[i])
    if k >= 12:
        if len(sys.argv) > 2 and int(scta[i]) != int(sys.argv[2]): continue
        print(f"  {(step[i, 1] - base) / 1e3:8.2f} us   cta {int(scta[i]):3d}  {NAMES.get(k, k)}")
    else:
        print(f"  {(step[i, 1] - base) / 1e3:8.2f} us   {NAMES.get(k, k):10s} start; dep resolved +{(step[i, 2] - step[i, 1]) / 1e3:.2f}; end +{(step[i, 3] - step[i, 1]) / 1e3:.2f}")
# averages over layers: attention time, layer-kernel time
at = [(step[i, 3] - step[i, 2]) / 1e3 for i in attn]
print(f"attention dep->end avg {np.mean(at):.1f} us")
lay = np.nonzero(skid == 1)[0]
lt = [(step[i, 3] - step[i, 2]) / 1e3 for i in lay]
print(f"layer kernel dep->end avg {np.mean(lt[1:-1]):.1f} us (first {lt[0]:.1f}, last {lt[-1]:.1f})")
eng.close()

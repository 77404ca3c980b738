"""Phase timeline inside the persistent chain kernel (CTA 0, epilogue thread 0 marks) for one decode step."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rr_b200 import _lib
from rr_b200.models import SPECS, make_weights
from rr_b200.engine import Engine
spec = SPECS["llama-3-8b"]
w = make_weights(spec, seed=0, device="cuda")
eng = Engine(w, max_batch=64, ctx_max=640, max_prefill_tokens=8192)
ids = np.random.RandomState(0).randint(0, spec.vocab, size=(64, 512)).astype(np.int32)
start = np.arange(0, 64 * 512 + 1, 512, dtype=np.int32)
eng.run_batch(ids.reshape(-1), start, 8)
N = 8000
_lib.check(_lib.lib.rr_debug_trace_start(N))
eng.run_batch(ids.reshape(-1), start, 4)
buf = (C.c_uint64 * (4 * N))(); n = C.c_int32()
_lib.check(_lib.lib.rr_debug_trace_stop(buf, N, C.byref(n)))
a = np.frombuffer(buf, dtype=np.uint64)[: 4 * n.value].reshape(-1, 4).astype(np.int64)
a = a[np.argsort(a[:, 1], kind="stable")]
# last decode step: find the last embed (kid 8) with 64-row grid -> take everything after it
emb = [i for i in range(len(a)) if a[i, 0] == 8]
dec = a[emb[-1]:]
names = {3: "attn", 11: "chain_start", 20: "epi_begin", 21: "G0 O done", 22: "N0 done", 23: "G1 gate/up done", 24: "G2 down done",
         25: "N1 done", 26: "G3 next done", 1: "gemm", 5: "norm", 8: "embed", 9: "argmax"}
t0 = dec[0, 1]
print(f"step span {(dec[-1, 3] - t0) / 1e3:.1f} us, {len(dec)} records")
# layer 16: records between the 17th and 18th attention
att = [i for i in range(len(dec)) if dec[i, 0] == 3]
lo, hi = att[16], att[17]
prev = dec[lo, 2]
for kid, s, d, e in dec[lo:hi + 1]:
    print(f"  {names.get(int(kid), str(kid)):18s} start {(s - t0) / 1e3:9.1f} dep {(d - t0) / 1e3:9.1f} end {(e - t0) / 1e3:9.1f}")
seg = {}
for l in range(4, 28):
    lo, hi = att[l], att[l + 1]
    marks = {int(k): s for k, s, d, e in dec[lo:hi] if k >= 20}
    ch = [r for r in dec[lo:hi] if r[0] == 11][0]
    seq = [("attn(dep->chain dep)", dec[lo, 2], ch[2]), ("G0 O", ch[2], marks[21]), ("N0", marks[21], marks[22]), ("G1 gate/up", marks[22], marks[23]),
           ("G2 down", marks[23], marks[24]), ("N1", marks[24], marks[25]), ("G3 qkv", marks[25], marks[26]), ("tail->next attn dep", marks[26], dec[hi, 2])]
    for nm, x, y in seq:
        seg.setdefault(nm, []).append((y - x) / 1e3)
print("mean per-layer segment (us), CTA 0 view:")
tot = 0
for k, v in seg.items():
    print(f"  {k:24s} {np.mean(v):7.1f}"); tot += np.mean(v)
print(f"  sum {tot:.1f}")
eng.close()

"""A/B of two engine configurations inside ONE process, alternating: decode step time (CUDA events around the graph
replay) at 64 rows after a 512-token prefill.  Usage: python tools/decode_ab.py ENV_VAR [VALUE]  (variant B is created
with ENV_VAR=VALUE in the environment)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rr_b200.models import SPECS, make_weights
from rr_b200.engine import Engine

var = sys.argv[1]
val = sys.argv[2] if len(sys.argv) > 2 else "1"
extra = dict(kv.split("=") for kv in sys.argv[3:])          # further VAR=VALUE pairs set together with the first one
spec = SPECS["llama-3-8b"]
w = make_weights(spec, seed=0, device="cuda")
os.environ.pop(var, None)
eng_a = Engine(w, max_batch=64, ctx_max=640, max_prefill_tokens=8192)
os.environ[var] = val
os.environ.update(extra)
eng_b = Engine(w, max_batch=64, ctx_max=640, max_prefill_tokens=8192)
os.environ.pop(var, None)
for k in extra: os.environ.pop(k, None)
ids = np.random.RandomState(0).randint(0, spec.vocab, size=(64, 512)).astype(np.int32)
start = np.arange(0, 64 * 512 + 1, 512, dtype=np.int32)
def run(e, new=48):
    e.reset_stats()
    recs, _ = e.run_batch(ids.reshape(-1), start, new)
    st = e.stats()
    return st["decode_ms_total"] / max(1, st["decode_steps"]), [r.tokens for r in recs]
run(eng_a, 8)                       # warm-up: captures the decode graph -- variant B also sees the variable here
os.environ[var] = val
run(eng_b, 8)
os.environ.pop(var, None)
ta, tb = [], []
for _ in range(int(os.environ.get("AB_REPS", "4"))):
    a, toks_a = run(eng_a); b, toks_b = run(eng_b)
    ta.append(a); tb.append(b)
print(f"A (default)    : {sorted(ta)[len(ta)//2]:.4f} ms/step median  {['%.3f' % t for t in ta]}")
print(f"B ({var}={val} {extra if extra else ''}): {sorted(tb)[len(tb)//2]:.4f} ms/step median  {['%.3f' % t for t in tb]}")
print("tokens equal:", toks_a == toks_b)
eng_a.close(); eng_b.close()

"""A/B of two engine configurations inside ONE process, alternating, so both see the same thermal / power state
(boxes and even consecutive runs differ by several %): prefill chunk time (CUDA events) of 16 x 512-token prompts.
Usage: python tools/prefill_ab.py ENV_VAR   -> variant B is created with ENV_VAR=1 in the environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rr_b200.models import SPECS, make_weights
from rr_b200.engine import Engine

var = sys.argv[1] if len(sys.argv) > 1 else "RR_NO_DEFER_NORM"
spec = SPECS["llama-3-8b"]
w = make_weights(spec, seed=0, device="cuda")
os.environ.pop(var, None)
eng_a = Engine(w, max_batch=64, ctx_max=640, max_prefill_tokens=8192)
os.environ[var] = "1"
eng_b = Engine(w, max_batch=64, ctx_max=640, max_prefill_tokens=8192)
os.environ.pop(var, None)
g = torch.Generator().manual_seed(0)
prompts = [torch.randint(0, spec.vocab, (512,), generator=g).tolist() for _ in range(16)]
def chunk_ms(e):
    a = e.stats()["prefill_ms_total"]
    e.prefill(prompts, list(range(16)))
    return e.stats()["prefill_ms_total"] - a
for e in (eng_a, eng_b): chunk_ms(e)
ta, tb = [], []
for _ in range(6):
    ta.append(chunk_ms(eng_a)); tb.append(chunk_ms(eng_b))
fa, fb = eng_a.prefill(prompts, list(range(16)))[0], eng_b.prefill(prompts, list(range(16)))[0]
print(f"A (default)      : {sorted(ta)[len(ta)//2]:.2f} ms median  {['%.1f' % t for t in ta]}")
print(f"B ({var}=1): {sorted(tb)[len(tb)//2]:.2f} ms median  {['%.1f' % t for t in tb]}")
print("first tokens equal:", [int(x) for x in fa] == [int(x) for x in fb])
eng_a.close(); eng_b.close()

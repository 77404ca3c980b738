"""Short full-size workload for ncu: one prefill chunk (16 x 512 tokens) + N decode steps at 64 rows,
mean decode context (pos 575), eager launches (no CUDA graph) so every kernel is listed.
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py --decode-steps 2"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rr_b200.models import SPECS, make_weights
from rr_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="llama-3-8b")
ap.add_argument("--decode-steps", type=int, default=2)
ap.add_argument("--prefill-prompts", type=int, default=16)
ap.add_argument("--pos", type=int, default=575)
a = ap.parse_args()
spec = SPECS[a.model]
w = make_weights(spec, seed=0, device="cuda")
eng = Engine(w, max_batch=64, ctx_max=640, max_prefill_tokens=8192, use_cuda_graph=False)
g = torch.Generator().manual_seed(0)
prompts = [torch.randint(0, spec.vocab, (512,), generator=g).tolist() for _ in range(a.prefill_prompts)]
if a.prefill_prompts:
    eng.prefill(prompts, list(range(a.prefill_prompts)))
toks = list(range(64))
for s in range(a.decode_steps):
    nxt, _ = eng.decode_step(list(range(64)), toks, [a.pos + s] * 64)
    toks = [int(t) for t in nxt]
print("stats", eng.stats())
eng.close()

"""HTTP leg of bench.py alone (one Llama-3-8B replica + the in-process gateway + 64 OpenAI-SDK clients in 8 processes), for
A/B of server-side settings:  python tools/http_probe.py [switch_interval_seconds]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from rr_b200.models import SPECS, make_weights
from rr_b200.engine import Engine
from rr_b200.router import Router, EngineBackend

if len(sys.argv) > 1:
    sys.setswitchinterval(float(sys.argv[1]))
spec = SPECS["llama-3-8b"]
w = make_weights(spec, seed=0, device="cuda")
eng = Engine(w, max_batch=64, ctx_max=640, max_prefill_tokens=8192)
ml = [{"model_name": "llama-3-8b", "litellm_params": {"model": "b200/llama-3-8b", "gpu": 0}}]
r = Router(model_list=ml, routing_strategy="least-busy", backends={0: EngineBackend(eng)})
prompts = bench.make_prompts(64, 512, spec.vocab, pinned=False).numpy()
r.completion_batch("llama-3-8b", [prompts[i] for i in range(64)], 8)        # warm-up + graph capture
for rep in range(2):
    print(f"switchinterval={sys.getswitchinterval()} rep{rep}", json.dumps(bench.http_leg(r, "llama-3-8b", prompts, 64, 128)))
r.close(); eng.close()

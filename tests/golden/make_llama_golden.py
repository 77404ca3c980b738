"""Generate tests/golden/llama_tiny_golden.pt with the REAL transformers LlamaForCausalLM (fp32, CPU)
on seeded bf16-rounded weights of the `tiny` spec.  Run in the build container:
    python tests/golden/make_llama_golden.py
The fixture holds: the weight seed/sigma, prompts, and HF logits for (a) the last prompt position and
(b) 6 teacher-forced greedy continuation steps per prompt."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rr_b200.models import SPECS, make_weights  # noqa: E402
from oracle.llama_ref import to_hf  # noqa: E402

SEED, SIGMA, JITTER, STEPS = 7, 0.05, 0.1, 6


def main():
    torch.manual_seed(0)
    spec = SPECS["tiny"]
    w = make_weights(spec, seed=SEED, sigma=SIGMA, device="cpu", norm_jitter=JITTER)
    hf = to_hf(w)
    g = torch.Generator().manual_seed(1234)
    lens = [1, 5, 64, 65, 130, 200]
    prompts = [torch.randint(0, spec.vocab, (n,), generator=g).tolist() for n in lens]
    out = {"seed": SEED, "sigma": SIGMA, "norm_jitter": JITTER, "spec": "tiny", "prompts": prompts,
           "logits": [], "tokens": [], "torch": str(torch.__version__)}
    with torch.no_grad():
        for p in prompts:
            toks = list(p)
            lg, tk = [], []
            for _ in range(STEPS + 1):
                logits = hf(torch.tensor([toks])).logits[0, -1].float()
                t = int(logits.argmax())
                lg.append(logits); tk.append(t); toks.append(t)
            out["logits"].append(torch.stack(lg))
            out["tokens"].append(tk)
    # checksum of the weights so the consumer can detect a generator drift
    out["weight_checksum"] = float(sum(t.float().abs().sum() for t in w.tensors()))
    torch.save(out, os.path.join(os.path.dirname(__file__), "llama_tiny_golden.pt"))
    print("wrote fixture;", {k: (len(v) if isinstance(v, list) else v) for k, v in out.items() if k != "logits"})


if __name__ == "__main__":
    main()

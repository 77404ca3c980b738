"""Generate tests/golden/family_golden.pt with the REAL transformers MistralForCausalLM and Phi3ForCausalLM (fp32, CPU)
on seeded bf16-rounded weights of small specs with those architectures' traits:
  * mistral: GQA, head_dim 128, rope_theta 10000, no sliding window (as Mistral-7B-v0.1 at <= 4096 tokens)
  * phi3:    MHA, head_dim 96, fused qkv_proj / gate_up_proj (exactly this repo's fused weight layout)
Run in the build container:  python tests/golden/make_family_golden.py
Pins oracle/llama_ref.py (the restatement the GPU engine is checked against) for the non-Llama families of SURVEY 8(a-8)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rr_b200.models import ModelSpec, make_weights  # noqa: E402

SPECS = {
    # name, vocab, hidden, inter, n_layers, n_heads, n_kv_heads, head_dim, rope_theta, rms_eps
    "mistral": ModelSpec("mistral-golden", 1024, 512, 1024, 2, 4, 2, 128, 10000.0, 1e-5),
    "phi3": ModelSpec("phi3-golden", 1024, 384, 1024, 2, 4, 4, 96, 10000.0, 1e-5),
}
SEED, SIGMA, JITTER, STEPS = 11, 0.05, 0.1, 3


def build_mistral(w):
    from transformers import MistralConfig, MistralForCausalLM
    s = w.spec
    cfg = MistralConfig(vocab_size=s.vocab, hidden_size=s.hidden, intermediate_size=s.inter, num_hidden_layers=s.n_layers,
                        num_attention_heads=s.n_heads, num_key_value_heads=s.n_kv_heads, head_dim=s.head_dim,
                        rope_theta=s.rope_theta, rms_norm_eps=s.rms_eps, max_position_embeddings=4096,
                        sliding_window=None, tie_word_embeddings=False, attn_implementation="eager")
    m = MistralForCausalLM(cfg).float()
    H, KV, D = s.n_heads, s.n_kv_heads, s.head_dim
    sd = {"model.embed_tokens.weight": w.embed, "lm_head.weight": w.lm_head, "model.norm.weight": w.final_norm}
    for l in range(s.n_layers):
        p = f"model.layers.{l}."
        sd[p + "self_attn.q_proj.weight"] = w.wqkv[l][: H * D]
        sd[p + "self_attn.k_proj.weight"] = w.wqkv[l][H * D:(H + KV) * D]
        sd[p + "self_attn.v_proj.weight"] = w.wqkv[l][(H + KV) * D:]
        sd[p + "self_attn.o_proj.weight"] = w.wo[l]
        sd[p + "mlp.gate_proj.weight"] = w.wgu[l][: s.inter]
        sd[p + "mlp.up_proj.weight"] = w.wgu[l][s.inter:]
        sd[p + "mlp.down_proj.weight"] = w.wdown[l]
        sd[p + "input_layernorm.weight"] = w.norm_attn[l]
        sd[p + "post_attention_layernorm.weight"] = w.norm_mlp[l]
    missing, unexpected = m.load_state_dict({k: v.float() for k, v in sd.items()}, strict=False)
    assert not [k for k in missing if "rotary" not in k] and not unexpected, (missing, unexpected)
    return m.eval()


def build_phi3(w):
    from transformers import Phi3Config, Phi3ForCausalLM
    s = w.spec
    cfg = Phi3Config(vocab_size=s.vocab, hidden_size=s.hidden, intermediate_size=s.inter, num_hidden_layers=s.n_layers,
                     num_attention_heads=s.n_heads, num_key_value_heads=s.n_kv_heads, rope_theta=s.rope_theta,
                     rms_norm_eps=s.rms_eps, max_position_embeddings=4096, original_max_position_embeddings=4096,
                     sliding_window=None, tie_word_embeddings=False, pad_token_id=None, attn_implementation="eager",
                     resid_pdrop=0.0, embd_pdrop=0.0, attention_dropout=0.0)
    assert cfg.hidden_size // cfg.num_attention_heads == s.head_dim
    m = Phi3ForCausalLM(cfg).float()
    sd = {"model.embed_tokens.weight": w.embed, "lm_head.weight": w.lm_head, "model.norm.weight": w.final_norm}
    for l in range(s.n_layers):
        p = f"model.layers.{l}."
        sd[p + "self_attn.qkv_proj.weight"] = w.wqkv[l]              # [q; k; v] rows, the layout HF Phi-3 uses
        sd[p + "self_attn.o_proj.weight"] = w.wo[l]
        sd[p + "mlp.gate_up_proj.weight"] = w.wgu[l]                 # [gate; up] rows
        sd[p + "mlp.down_proj.weight"] = w.wdown[l]
        sd[p + "input_layernorm.weight"] = w.norm_attn[l]
        sd[p + "post_attention_layernorm.weight"] = w.norm_mlp[l]
    missing, unexpected = m.load_state_dict({k: v.float() for k, v in sd.items()}, strict=False)
    assert not [k for k in missing if "rotary" not in k] and not unexpected, (missing, unexpected)
    return m.eval()


def main():
    import transformers
    out = {"seed": SEED, "sigma": SIGMA, "norm_jitter": JITTER, "torch": str(torch.__version__),
           "transformers": transformers.__version__, "families": {}}
    g = torch.Generator().manual_seed(4321)
    for fam, build in (("mistral", build_mistral), ("phi3", build_phi3)):
        spec = SPECS[fam]
        w = make_weights(spec, seed=SEED, sigma=SIGMA, device="cpu", norm_jitter=JITTER)
        hf = build(w)
        prompts = [torch.randint(0, spec.vocab, (n,), generator=g).tolist() for n in (1, 7, 64, 130)]
        logits, tokens = [], []
        with torch.no_grad():
            for p in prompts:
                toks, lg, tk = list(p), [], []
                for _ in range(STEPS):
                    lo = hf(torch.tensor([toks])).logits[0, -1].float()
                    t = int(lo.argmax())
                    lg.append(lo); tk.append(t); toks.append(t)
                logits.append(torch.stack(lg)); tokens.append(tk)
        out["families"][fam] = {"spec": [spec.name, spec.vocab, spec.hidden, spec.inter, spec.n_layers, spec.n_heads,
                                         spec.n_kv_heads, spec.head_dim, spec.rope_theta, spec.rms_eps],
                                "prompts": prompts, "logits": logits, "tokens": tokens,
                                "weight_checksum": float(sum(t.float().abs().sum() for t in w.tensors()))}
        print(fam, "ok", [len(p) for p in prompts])
    torch.save(out, os.path.join(os.path.dirname(__file__), "family_golden.pt"))


if __name__ == "__main__":
    main()

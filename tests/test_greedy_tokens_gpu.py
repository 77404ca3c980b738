"""GPU: token-level parity as BASELINE.json's north_star words it -- "generated tokens match a local greedy HF-transformers
decode of the same prompts".  The engine decodes FREE-RUNNING (its own tokens fed back) for 128 steps and must emit exactly
the tokens of the real transformers LlamaForCausalLM.generate(do_sample=False) on the same weights, in fp32 and in bf16.

Seeded random weights give almost flat logits (top-2 margin ~0.2 sigma over a 128 k vocabulary) where bf16 rounding noise
alone flips an argmax every few tokens, so the weights are SHARPENED first: unit-variance embedding
rows and lm_head row perm[t] = embedding row t (a fixed random permutation), which makes the model predict perm[current
token] with a top-2 margin of several logit standard deviations while every layer still runs at its real shape and
contributes a comparable share of the hidden state.  The margin
is measured and asserted to exceed 10x the stated logit tolerance of tests/test_engine_gpu.py, so an exact match is a
meaningful statement about the whole pipeline (prefill, KV cache positions, RoPE, decode under the CUDA graph at 64 active
rows -- the configuration bench.py times) and not luck.  Small numeric deviations are covered by the teacher-forced logit
tests; this one catches wrong rows, slots, positions, stale buffers, graph-replay state."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sharpen(w, seed=0, scale=1.0):
    """Embedding rows scaled to unit variance per element (so the token embedding is a large part of the residual stream
    next to the layer outputs, which are O(1) per element after RMSNorm-ed inputs), lm_head row perm[t] = embedding row t."""
    g = torch.Generator(device=w.embed.device).manual_seed(seed)
    perm = torch.randperm(w.spec.vocab, generator=g, device=w.embed.device)
    w.embed.copy_((w.embed.float() * (scale / w.embed.float().std())).bfloat16())
    w.lm_head.index_copy_(0, perm, w.embed)
    return perm


def _hf_generate(w, prompts, max_new, dtype, want_margin=False):
    """transformers greedy generate; with want_margin also the smallest top-2 margin / logit sigma over every generated
    position of every prompt (HF's own scores)."""
    from oracle.llama_ref import to_hf
    hf = to_hf(w, dtype=dtype)
    min_margin = float("inf")
    with torch.no_grad():
        by_len = {}
        for i, p in enumerate(prompts):
            by_len.setdefault(len(p), []).append(i)
        res = [None] * len(prompts)
        for n, idx in by_len.items():                              # equal lengths batch without padding
            ids = torch.tensor([prompts[i] for i in idx], device=w.embed.device)
            gen = hf.generate(ids, max_new_tokens=max_new, do_sample=False, use_cache=True, pad_token_id=0,
                              eos_token_id=None, output_scores=want_margin, return_dict_in_generate=True)
            for j, i in enumerate(idx):
                res[i] = gen.sequences[j, n:].tolist()
            if want_margin:
                for sc in gen.scores:                              # [batch, vocab] per step
                    sc = sc.float()
                    top2 = sc.topk(2, dim=-1).values
                    min_margin = min(min_margin, ((top2[:, 0] - top2[:, 1]) / sc.std(dim=-1)).min().item())
    del hf
    torch.cuda.empty_cache()
    return (res, min_margin) if want_margin else res


def _margin_over_sigma(w, prompt):
    from oracle import llama_ref
    lg = llama_ref.forward_logits(w, prompt)[-1].float()
    top2 = lg.topk(2).values
    return ((top2[0] - top2[1]) / lg.std()).item()


def _first_divergence(a, b):
    for k, (x, y) in enumerate(zip(a, b)):
        if x != y:
            return k
    return None


@pytest.mark.parametrize("spec_name,max_batch,lens", [("tiny", 8, [1, 5, 64, 65, 130]), ("small", 64, [3, 64, 129, 300, 512])])
def test_free_running_128_tokens_equal_hf_greedy_fp32_and_bf16(spec_name, max_batch, lens):
    from rr_b200.engine import Engine
    from rr_b200.models import SPECS, make_weights
    spec = SPECS[spec_name]
    w = make_weights(spec, seed=11, sigma=0.05 if spec.hidden < 1024 else 0.03, device="cuda", norm_jitter=0.1)
    _sharpen(w)
    g = torch.Generator().manual_seed(3)
    prompts = [torch.randint(0, spec.vocab, (n,), generator=g).tolist() for n in lens]
    margins = [_margin_over_sigma(w, p) for p in prompts]
    assert min(margins) > 10 * 0.08, margins                      # 10 x TOL_MAX of tests/test_engine_gpu.py
    want32 = _hf_generate(w, prompts, 128, torch.float32)
    want16 = _hf_generate(w, prompts, 128, torch.bfloat16)
    eng = Engine(w, max_batch=max_batch, ctx_max=704, max_prefill_tokens=2048, use_cuda_graph=True)
    try:
        recs = [eng.wait(t, timeout=300) for t in [eng.submit(p, 128) for p in prompts]]
        for i, r in enumerate(recs):
            assert r.status == 0 and len(r.tokens) == 128
            d32, d16 = _first_divergence(r.tokens, want32[i]), _first_divergence(r.tokens, want16[i])
            assert d32 is None, f"{spec_name} prompt {i} (len {lens[i]}): first divergence from HF fp32 greedy at token {d32}"
            assert d16 is None, f"{spec_name} prompt {i} (len {lens[i]}): first divergence from HF bf16 greedy at token {d16}"
        assert len({tuple(r.tokens[:4]) for r in recs}) > 1        # the rows really decode different sequences
        print(f"\n[{spec_name}] 128 free-running tokens x {len(lens)} prompts == HF greedy (fp32 and bf16); "
              f"top-2 margin / sigma: min {min(margins):.1f}")
    finally:
        eng.close()


def test_benched_configuration_64_rows_cuda_graph_llama3_8b_shape_equals_hf_greedy():
    """The configuration bench.py times -- 64 active rows, 512-token prompts, decode under the CUDA graph, Llama-3-8B layer
    shape and vocabulary (2 layers keep HF's side affordable) -- through rr_engine_run_batch, against HF greedy."""
    from rr_b200.engine import Engine
    from rr_b200.models import SPECS, make_weights
    spec = SPECS["llama-3-8b-2l"]
    w = make_weights(spec, seed=0, sigma=0.02, device="cuda")
    _sharpen(w)
    n, P, M = 64, 512, 128
    prompts = []
    for r in range(n):
        g = torch.Generator().manual_seed(1234 + r)               # bench.py's prompt generator
        prompts.append(torch.randint(0, spec.vocab, (P,), generator=g).tolist())
    assert _margin_over_sigma(w, prompts[0]) > 10 * 0.08
    want = _hf_generate(w, prompts, M, torch.float32)
    eng = Engine(w, max_batch=64, ctx_max=640, max_prefill_tokens=8192, use_cuda_graph=True)
    try:
        ids = np.asarray(prompts, dtype=np.int32).reshape(-1)
        start = np.arange(0, (n + 1) * P, P, dtype=np.int32)
        for rep in range(2):                                       # second burst: graph replays on reused slots
            recs, toks = eng.run_batch(ids, start, M)
            bad = [(i, _first_divergence(r.tokens, want[i])) for i, r in enumerate(recs) if r.tokens != want[i]]
            assert not bad, f"burst {rep}: rows diverging from HF greedy (row, first token index): {bad[:8]}"
        assert eng.stats()["decode_steps"] >= 2 * (M - 1)
    finally:
        eng.close()


def test_full_size_llama3_8b_32_layers_free_running_greedy_equals_hf_fp32():
    """BASELINE.json's model at full size -- 32 layers, hidden 4096, vocabulary 128256 -- free-running for 32 tokens on 8
    rows of 512-token prompts (the bench's prompt generator), against transformers LlamaForCausalLM in fp32 on the same
    (sharpened) weights.  Complements tests/test_full_model_gpu.py, which bounds the logit error at this size."""
    from rr_b200.engine import Engine
    from rr_b200.models import SPECS, make_weights
    spec = SPECS["llama-3-8b"]
    w = make_weights(spec, seed=0, sigma=0.02, device="cuda")
    _sharpen(w, scale=4.0)      # 64 sublayers add O(1) per element each: the embedding needs more weight to stay visible
    n, P, M = 8, 512, 32
    prompts = []
    for r in range(n):
        g = torch.Generator().manual_seed(1234 + r)
        prompts.append(torch.randint(0, spec.vocab, (P,), generator=g).tolist())
    want, min_margin = _hf_generate(w, prompts, M, torch.float32, want_margin=True)
    # what makes an exact comparison meaningful: HF's own smallest top-2 margin over all 8 x 32 generated positions must be
    # 10 x the stated 32-layer logit tolerance of tests/test_full_model_gpu.py (max |d| <= 0.40 sigma)
    assert min_margin > 10 * 0.40, min_margin
    print(f"\n[llama-3-8b, 32 layers] smallest top-2 margin over the generated path: {min_margin:.1f} sigma")
    eng = Engine(w, max_batch=64, ctx_max=640, max_prefill_tokens=8192, use_cuda_graph=True)
    try:
        ids = np.asarray(prompts, dtype=np.int32).reshape(-1)
        start = np.arange(0, (n + 1) * P, P, dtype=np.int32)
        recs, _ = eng.run_batch(ids, start, M)
        bad = [(i, _first_divergence(r.tokens, want[i])) for i, r in enumerate(recs) if r.tokens != want[i]]
        assert not bad, f"rows diverging from HF fp32 greedy at full size (row, first token index): {bad}"
        assert len({tuple(r.tokens) for r in recs}) == n
    finally:
        eng.close()

"""Model descriptors and seeded synthetic weights for the on-box replicas that replace the
reference's `litellm_params.model: bedrock/...` deployments (reference config/config.yaml:39-91).

No checkpoints or tokenizers exist on the box, so replicas are instantiated at the
exact architecture shapes with seeded random weights (manual_seed(0), sigma 0.02,
bf16).  Weight layout is what librr_b200.so consumes: torch nn.Linear convention
[out_features, in_features], q/k/v fused row-wise into `wqkv`, gate/up fused into `wgu`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch


@dataclass(frozen=True)
class ModelSpec:
    name: str
    vocab: int
    hidden: int
    inter: int
    n_layers: int
    n_heads: int
    n_kv_heads: int
    head_dim: int = 128
    rope_theta: float = 500000.0
    rms_eps: float = 1e-5

    @property
    def n_params(self) -> int:
        per_layer = (self.hidden * (self.n_heads + 2 * self.n_kv_heads) * self.head_dim
                     + self.n_heads * self.head_dim * self.hidden + 3 * self.hidden * self.inter
                     + 2 * self.hidden)
        return 2 * self.vocab * self.hidden + self.n_layers * per_layer + self.hidden

    @property
    def weight_bytes_per_decode_step(self) -> int:
        """bf16 bytes streamed once per decode step: every layer + final norm + lm_head
        (the embedding table is gathered, not streamed) — BASELINE.md §4."""
        per_layer = (self.hidden * (self.n_heads + 2 * self.n_kv_heads) * self.head_dim
                     + self.n_heads * self.head_dim * self.hidden + 3 * self.hidden * self.inter
                     + 2 * self.hidden)
        return 2 * (self.n_layers * per_layer + self.hidden + self.vocab * self.hidden)

    @property
    def kv_bytes_per_token(self) -> int:
        return 2 * self.n_layers * self.n_kv_heads * self.head_dim * 2

    @property
    def prefill_flops_per_token(self) -> int:
        """2 * (matmul parameters touched per token), attention excluded (<1 % at 512 tokens)."""
        per_layer = (self.hidden * (self.n_heads + 2 * self.n_kv_heads) * self.head_dim
                     + self.n_heads * self.head_dim * self.hidden + 3 * self.hidden * self.inter)
        return 2 * (self.n_layers * per_layer + self.vocab * self.hidden)


SPECS: Dict[str, ModelSpec] = {
    "llama-3-8b": ModelSpec("llama-3-8b", 128256, 4096, 14336, 32, 32, 8, 128, 500000.0, 1e-5),
    "mistral-7b": ModelSpec("mistral-7b", 32000, 4096, 14336, 32, 32, 8, 128, 10000.0, 1e-5),
    # Phi-3-mini-4k: fused qkv / gate_up are layout only (our layout is fused anyway); head_dim 96, MHA
    "phi-3-mini": ModelSpec("phi-3-mini", 32064, 3072, 8192, 32, 32, 32, 96, 10000.0, 1e-5),
    # small shapes for parity tests / smoke (same kernels, same code path)
    "tiny": ModelSpec("tiny", 1024, 512, 1024, 2, 4, 2, 128, 500000.0, 1e-5),
    "small": ModelSpec("small", 4096, 1024, 2816, 4, 8, 2, 128, 10000.0, 1e-5),
    "llama-3-8b-2l": ModelSpec("llama-3-8b-2l", 128256, 4096, 14336, 2, 32, 8, 128, 500000.0, 1e-5),
    "phi-3-mini-2l": ModelSpec("phi-3-mini-2l", 32064, 3072, 8192, 2, 32, 32, 96, 10000.0, 1e-5),
    "mistral-7b-2l": ModelSpec("mistral-7b-2l", 32000, 4096, 14336, 2, 32, 8, 128, 10000.0, 1e-5),
    "tiny96": ModelSpec("tiny96", 1024, 384, 1024, 2, 4, 4, 96, 10000.0, 1e-5),
    "small96": ModelSpec("small96", 4096, 768, 2048, 3, 8, 2, 96, 10000.0, 1e-5),
}


def resolve_spec(model: str) -> ModelSpec:
    """`b200/llama-3-8b`, `llama-3-8b`, ... -> ModelSpec."""
    key = model.split("/", 1)[-1].split("@", 1)[0].lower()
    if key not in SPECS:
        raise KeyError(f"unknown local model {model!r}; known: {sorted(SPECS)}")
    return SPECS[key]


@dataclass
class Weights:
    spec: ModelSpec
    embed: torch.Tensor
    lm_head: torch.Tensor
    final_norm: torch.Tensor
    wqkv: List[torch.Tensor] = field(default_factory=list)
    wo: List[torch.Tensor] = field(default_factory=list)
    wgu: List[torch.Tensor] = field(default_factory=list)
    wdown: List[torch.Tensor] = field(default_factory=list)
    norm_attn: List[torch.Tensor] = field(default_factory=list)
    norm_mlp: List[torch.Tensor] = field(default_factory=list)
    flat: Optional[torch.Tensor] = None      # the one allocation every tensor above is a view of (make_weights); None after .to()

    def tensors(self) -> List[torch.Tensor]:
        out = [self.embed, self.lm_head, self.final_norm]
        for lst in (self.wqkv, self.wo, self.wgu, self.wdown, self.norm_attn, self.norm_mlp):
            out.extend(lst)
        return out

    def to(self, device) -> "Weights":
        mv = lambda t: t.to(device)
        return Weights(self.spec, mv(self.embed), mv(self.lm_head), mv(self.final_norm),
                       [mv(t) for t in self.wqkv], [mv(t) for t in self.wo],
                       [mv(t) for t in self.wgu], [mv(t) for t in self.wdown],
                       [mv(t) for t in self.norm_attn], [mv(t) for t in self.norm_mlp])

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors())


def make_weights(spec: ModelSpec, seed: int = 0, sigma: float = 0.02, device="cuda",
                 norm_jitter: float = 0.0, allocate_only: bool = False) -> Weights:
    """Seeded N(0, sigma^2) bf16 weights (norm weights 1 + norm_jitter * N(0,1)).

    The generator lives on `device`, so values are reproducible per (torch version, device type);
    parity tests share the tensors with the oracle instead of regenerating them.
    `allocate_only` leaves the values uninitialised (ranks > 0 before the start-up NCCL broadcast).
    """
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    nq, nkv = spec.n_heads * spec.head_dim, spec.n_kv_heads * spec.head_dim
    # ONE flat allocation, every tensor a 256-byte-aligned view of it: the start-up weight broadcast is then a single
    # ncclBroadcast instead of 195 small ones (broadcast_weights)
    shapes = [(spec.vocab, spec.hidden), (spec.vocab, spec.hidden), (spec.hidden,)]
    for _ in range(spec.n_layers):
        shapes += [(nq + 2 * nkv, spec.hidden), (spec.hidden, nq), (2 * spec.inter, spec.hidden), (spec.hidden, spec.inter),
                   (spec.hidden,), (spec.hidden,)]
    offs, total = [], 0
    for sh in shapes:
        n = 1
        for d in sh:
            n *= d
        offs.append(total)
        total += (n + 127) // 128 * 128
    flat = torch.empty(total, device=dev, dtype=torch.bfloat16)
    views = iter([flat[o: o + torch.Size(sh).numel()].view(*sh) for o, sh in zip(offs, shapes)])

    def rnd(*shape):
        out = next(views)
        assert tuple(out.shape) == tuple(shape)
        if allocate_only:
            return out
        rows = shape[0]
        step = max(1, (1 << 26) // max(1, shape[1]))      # bound the fp32 temporary to 256 MB
        for r0 in range(0, rows, step):
            r1 = min(rows, r0 + step)
            out[r0:r1] = (torch.randn(r1 - r0, shape[1], device=dev, generator=g) * sigma).bfloat16()
        return out

    def norm():
        out = next(views)
        if allocate_only:
            return out
        w = torch.ones(spec.hidden, device=dev)
        if norm_jitter:
            w = w + norm_jitter * torch.randn(spec.hidden, device=dev, generator=g)
        out.copy_(w.bfloat16())
        return out

    w = Weights(spec, rnd(spec.vocab, spec.hidden), rnd(spec.vocab, spec.hidden), norm())
    for _ in range(spec.n_layers):
        w.wqkv.append(rnd(nq + 2 * nkv, spec.hidden))
        w.wo.append(rnd(spec.hidden, nq))
        w.wgu.append(rnd(2 * spec.inter, spec.hidden))
        w.wdown.append(rnd(spec.hidden, spec.inter))
        w.norm_attn.append(norm())
        w.norm_mlp.append(norm())
    w.flat = flat
    return w


def broadcast_weights(w: Weights, src: int = 0, group=None) -> None:
    """Start-up weight broadcast (K12): rank `src` -> every replica of the model group over
    NCCL/NVLink.  The only collective on the path; requests never cross GPUs afterwards.
    Weights made by make_weights live in one flat buffer: ONE ncclBroadcast of 16 GB (round 1 issued 195 per-tensor calls
    and reached ~70 % of the link rate)."""
    import torch.distributed as dist
    if w.flat is not None:
        dist.broadcast(w.flat, src=src, group=group)
        return
    for t in w.tensors():
        dist.broadcast(t, src=src, group=group)

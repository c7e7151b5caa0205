"""Seeded synthetic weights with the reference's HF key names.

No checkpoints are reachable offline, so benchmarks and tests use random weights
of the real architecture.  To get non-trivial acceptance without a trained draft
we use the "permutation" construction of SURVEY.md section 8(d): the embedding
E = randn(V, d) and the LM head is E re-indexed by a random permutation, so both
the target and the draft predict roughly perm[token]; every other matrix is
N(0, (scale/sqrt(fan_in))^2) so residual branches perturb but do not drown the
signal.  `draft_scale` > `scale` makes the draft noisier and lowers the accepted
length.

Key names follow the reference modules: target `model.*` / `lm_head.weight`
(llama.py:529-655), draft `glide.*` (llama_glide.py:401-406).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass
class ModelDims:
    hidden: int
    heads: int
    kv_heads: int
    layers: int
    ffn: int
    vocab: int
    rope_theta: float = 10000.0
    rope_factor: float = 1.0
    rms_eps: float = 1e-5
    qkv_bias: bool = False
    head_dim: int = 128
    max_position_embeddings: int = 1 << 20
    inv_freq: tuple | None = None        # explicit rotary frequencies [head_dim / 2] (rope types beyond default / linear)

    @staticmethod
    def from_hf_config(cfg) -> "ModelDims":
        """Read the fields the reference reads from a LlamaConfig / Qwen2Config
        (llama.py:69-76).  transformers 5.x moved rope_theta under rope_parameters."""
        theta = getattr(cfg, "rope_theta", None)
        rp = getattr(cfg, "rope_parameters", None) or getattr(cfg, "rope_scaling", None) or {}
        if theta is None:
            theta = rp.get("rope_theta", 10000.0)
        factor = 1.0
        inv_freq = None
        rtype = rp.get("rope_type", rp.get("type", "default")) if isinstance(rp, dict) else "default"
        if rtype == "linear":
            factor = float(rp.get("factor", 1.0))
        elif rtype not in ("default", None):
            # The reference builds its tables with HF's LlamaRotaryEmbedding(config=config) (llama.py:547,585), i.e. with
            # HF's per-type frequency functions.  The kernels take any STATIC frequency vector, which covers llama3;
            # types that rescale cos/sin (yarn, longrope) or depend on the sequence length (dynamic) are not supported.
            if rtype == "dynamic":
                raise NotImplementedError("rope type 'dynamic' (sequence-length dependent frequencies) is not supported")
            from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
            if rtype not in ROPE_INIT_FUNCTIONS:
                raise NotImplementedError(f"rope type {rtype!r} is not supported")
            freqs, attention_scaling = ROPE_INIT_FUNCTIONS[rtype](cfg, "cpu")
            if abs(float(attention_scaling) - 1.0) > 1e-6:
                raise NotImplementedError(f"rope type {rtype!r} rescales cos/sin by {attention_scaling}: not supported")
            inv_freq = tuple(float(x) for x in freqs.float().tolist())
        hd = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
        bias = bool(getattr(cfg, "attention_bias", False)) or cfg.__class__.__name__.startswith("Qwen2")
        return ModelDims(hidden=cfg.hidden_size, heads=cfg.num_attention_heads,
                         kv_heads=cfg.num_key_value_heads, layers=cfg.num_hidden_layers,
                         ffn=cfg.intermediate_size, vocab=cfg.vocab_size, rope_theta=float(theta),
                         rope_factor=factor, rms_eps=cfg.rms_norm_eps, qkv_bias=bias, head_dim=hd,
                         max_position_embeddings=cfg.max_position_embeddings, inv_freq=inv_freq)


# Architecture dims of BASELINE.json's configs (SURVEY.md section 8 header).
PRESETS = {
    "llama3-8b-262k": ModelDims(4096, 32, 8, 32, 14336, 128256, rope_theta=283461213.0,
                                max_position_embeddings=262144),
    "qwq-32b": ModelDims(5120, 40, 8, 64, 27648, 152064, rope_theta=1e6, qkv_bias=True,
                         rms_eps=1e-5, max_position_embeddings=32768),
    "vicuna-7b-16k": ModelDims(4096, 32, 32, 32, 11008, 32000, rope_factor=4.0,
                               max_position_embeddings=16384),
    "vicuna-13b-16k": ModelDims(5120, 40, 40, 40, 13824, 32000, rope_factor=4.0,
                                max_position_embeddings=16384),
    "tiny": ModelDims(256, 2, 1, 2, 512, 320),
    "tiny-gqa": ModelDims(512, 4, 2, 3, 768, 512),
    "tiny-qwen": ModelDims(256, 2, 1, 2, 512, 320, qkv_bias=True, rope_theta=1e6),
    # the headline head geometry (GQA-4: 74 verify tokens -> 296 packed rows -> heavy/heavy/light CTA trios) at toy widths
    "tiny-g4": ModelDims(1024, 8, 2, 2, 1024, 512),
    # MHA at toy widths (Vicuna-like: one head per kv head, wide trees make two row tiles)
    "tiny-mha": ModelDims(512, 4, 4, 2, 768, 512, rope_factor=4.0),
}


def synthetic_state_dict(dims: ModelDims, seed: int = 0, device="cpu", dtype=torch.float32,
                         scale: float = 1.0, draft_scale: float | None = None,
                         bias_std: float = 0.02):
    """Return {hf_key: tensor}.  Generation runs in fp32 on `device` with a seeded
    generator and is cast to `dtype` per tensor (so an 8B model never holds two
    fp32 copies)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ds = scale if draft_scale is None else draft_scale
    d, H, Hk, D, F, V = dims.hidden, dims.heads, dims.kv_heads, dims.head_dim, dims.ffn, dims.vocab

    def lin(out_f, in_f, s):
        w = torch.randn(out_f, in_f, generator=g, device=device, dtype=torch.float32)
        return (w * (s / math.sqrt(in_f))).to(dtype)

    def vec(n, std):
        return (torch.randn(n, generator=g, device=device, dtype=torch.float32) * std).to(dtype)

    ones = lambda n: torch.ones(n, device=device, dtype=dtype)
    W = {}
    E = torch.randn(V, d, generator=g, device=device, dtype=torch.float32)
    perm = torch.randperm(V, generator=g, device=device)
    W["model.embed_tokens.weight"] = E.to(dtype)
    head = torch.empty_like(E)
    head[perm] = E
    W["lm_head.weight"] = head.to(dtype)
    del E, head
    for i in range(dims.layers):
        p = f"model.layers.{i}."
        W[p + "self_attn.q_proj.weight"] = lin(H * D, d, scale)
        W[p + "self_attn.k_proj.weight"] = lin(Hk * D, d, scale)
        W[p + "self_attn.v_proj.weight"] = lin(Hk * D, d, scale)
        W[p + "self_attn.o_proj.weight"] = lin(d, H * D, scale)
        if dims.qkv_bias:
            W[p + "self_attn.q_proj.bias"] = vec(H * D, bias_std)
            W[p + "self_attn.k_proj.bias"] = vec(Hk * D, bias_std)
            W[p + "self_attn.v_proj.bias"] = vec(Hk * D, bias_std)
        W[p + "mlp.gate_proj.weight"] = lin(F, d, scale)
        W[p + "mlp.up_proj.weight"] = lin(F, d, scale)
        W[p + "mlp.down_proj.weight"] = lin(d, F, scale)
        W[p + "input_layernorm.weight"] = ones(d)
        W[p + "post_attention_layernorm.weight"] = ones(d)
    W["model.norm.weight"] = ones(d)
    for att in ("self_attn", "cross_attn"):
        p = f"glide.{att}."
        W[p + "q_proj.weight"] = lin(H * D, d, ds)
        W[p + "k_proj.weight"] = lin(Hk * D, d, ds)
        W[p + "v_proj.weight"] = lin(Hk * D, d, ds)
        W[p + "o_proj.weight"] = lin(d, H * D, ds)
        W[p + "q_proj.bias"] = vec(H * D, bias_std)
        W[p + "k_proj.bias"] = vec(Hk * D, bias_std)
        W[p + "v_proj.bias"] = vec(Hk * D, bias_std)
    W["glide.mlp.gate_proj.weight"] = lin(F, d, ds)
    W["glide.mlp.up_proj.weight"] = lin(F, d, ds)
    W["glide.mlp.down_proj.weight"] = lin(d, F, ds)
    W["glide.input_layernorm.weight"] = ones(d)
    W["glide.post_self_attention_layernorm.weight"] = ones(d)
    W["glide.post_cross_attention_layernorm.weight"] = ones(d)
    return W


def synthetic_prompt(vocab: int, length: int, seed: int = 0) -> torch.Tensor:
    """randint(2, V-2, (1, P)) on CPU."""
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randint(2, vocab - 2, (1, length), generator=g, dtype=torch.int64)

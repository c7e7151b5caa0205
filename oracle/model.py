"""Oracle restatement of the target model / draft layer forwards used by the decode loop.

TEST INFRASTRUCTURE (see oracle/__init__.py).  A functional torch-CPU model over an
HF-named weight dict; no nn.Module, no transformers import.  Follows
llama.py:455-626 (decoder layer / LlamaModel.forward), llama.py:199-226,304-329,357-392
(prefill / decoding / tree_decoding) and llama_glide.py:388-468 (draft layer).
The Qwen2 twin differs only by q/k/v biases in the target (qwen2.py:278-281).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import attention as A
from . import flash_ref


@dataclass
class Dims:
    hidden: int
    heads: int
    kv_heads: int
    layers: int
    ffn: int
    vocab: int
    rope_theta: float = 10000.0
    rope_factor: float = 1.0          # linear RoPE scaling (Vicuna-16k: 4.0)
    rms_eps: float = 1e-5
    qkv_bias: bool = False            # Qwen2 target
    head_dim: int = 128


def rms_norm(x, w, eps):
    """HF LlamaRMSNorm (used via llama.py:36): fp32 statistics, cast back, then scale."""
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return w * xf.to(x.dtype)


def rope_cos_sin(dims: Dims, position_ids, dtype):
    """HF LlamaRotaryEmbedding.forward (llama.py:547,585): fp32 tables cast to dtype.
    position_ids (B,S) -> cos, sin (B,S,D)."""
    D = dims.head_dim
    inv = 1.0 / (dims.rope_theta ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
    inv = inv / dims.rope_factor
    fr = position_ids.float().unsqueeze(-1) * inv
    emb = torch.cat([fr, fr], dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def apply_rope(x, cos, sin):
    """apply_rotary_pos_emb(..., unsqueeze_dim=2) on (B,S,H,D) (llama.py:378)."""
    c, s = cos.unsqueeze(2), sin.unsqueeze(2)
    h = x.shape[-1] // 2
    rot = torch.cat([-x[..., h:], x[..., :h]], dim=-1)
    return x * c + rot * s


def mlp(x, W, p):
    return F.linear(F.silu(F.linear(x, W[p + "gate_proj.weight"])) * F.linear(x, W[p + "up_proj.weight"]),
                    W[p + "down_proj.weight"])


class OracleModel:
    """Target LLM + GliDe draft layer with module-style KV caches (llama.py:86-87,
    llama_glide.py:55-56): per layer K/V (B, S_max, Hk, D)."""

    def __init__(self, dims: Dims, weights: dict, dtype=torch.float32):
        self.d = dims
        self.dtype = dtype
        self.W = {k: v.to(dtype) for k, v in weights.items()}
        self.K = [None] * dims.layers
        self.V = [None] * dims.layers
        self.dK = None
        self.dV = None
        self.max_len = 512
        self.draft_max_len = 512

    # ---- helpers --------------------------------------------------------
    def embed(self, ids):
        return self.W["model.embed_tokens.weight"][ids]

    def lm_head(self, h):
        return F.linear(h, self.W["lm_head.weight"])

    def _qkv(self, x, p, bias):
        d = self.d
        B, S, _ = x.shape
        q = F.linear(x, self.W[p + "q_proj.weight"], self.W.get(p + "q_proj.bias") if bias else None)
        k = F.linear(x, self.W[p + "k_proj.weight"], self.W.get(p + "k_proj.bias") if bias else None)
        v = F.linear(x, self.W[p + "v_proj.weight"], self.W.get(p + "v_proj.bias") if bias else None)
        return (q.view(B, S, d.heads, d.head_dim), k.view(B, S, d.kv_heads, d.head_dim),
                v.view(B, S, d.kv_heads, d.head_dim))

    # ---- target -----------------------------------------------------------
    def target_forward(self, ids, exec_type, cache_lens=None, tree_mask=None):
        """LlamaModel.forward (llama.py:559-626)."""
        d = self.d
        B, S = ids.shape
        if tree_mask is None:
            pos = torch.arange(S)[None, :].expand(B, S)
        else:
            pos = tree_mask.sum(-1) - 1                                   # llama.py:575
        if cache_lens is not None:
            pos = pos + cache_lens[:, None]
        h = self.embed(ids)
        cos, sin = rope_cos_sin(d, pos, self.dtype)
        for li in range(d.layers):
            p = f"model.layers.{li}."
            x = rms_norm(h, self.W[p + "input_layernorm.weight"], d.rms_eps)
            q, k, v = self._qkv(x, p + "self_attn.", d.qkv_bias)
            q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
            if exec_type == "prefill":                                    # llama.py:199-226
                a = flash_ref.flash_attn_func(q, k, v, causal=True)
                self.K[li] = q.new_zeros(B, S + self.max_len, d.kv_heads, d.head_dim)
                self.V[li] = q.new_zeros(B, S + self.max_len, d.kv_heads, d.head_dim)
                self.K[li][:, :S] = k
                self.V[li][:, :S] = v
            elif exec_type == "decoding":                                 # llama.py:304-329
                a = flash_ref.flash_attn_with_kvcache(q, self.K[li], self.V[li], k, v,
                                                      cache_seqlens=cache_lens, causal=True)
            elif exec_type == "tree_decoding":                            # llama.py:357-392
                a = A.target_hybrid_tree_attention(q, k, v, self.K[li], self.V[li], cache_lens,
                                                   tree_mask, last_layer=(li == d.layers - 1))
            else:
                raise ValueError(f"Unknown inference_type: {exec_type}")
            a = a.reshape(B, S, d.heads * d.head_dim).to(h.dtype)
            h = h + F.linear(a, self.W[p + "self_attn.o_proj.weight"])
            x = rms_norm(h, self.W[p + "post_attention_layernorm.weight"], d.rms_eps)
            h = h + mlp(x, self.W, p + "mlp.")
        return rms_norm(h, self.W["model.norm.weight"], d.rms_eps)

    # ---- draft ------------------------------------------------------------
    def draft_forward(self, h, pos, exec_type, cache_lens, llm_kv_len, tree_mask=None):
        """LlamaGlideDecoderLayer.forward (llama_glide.py:424-468).  llm_kv = the
        target's LAST layer cache (llama_glide.py:973,1011-1012)."""
        d = self.d
        B, S, _ = h.shape
        cos, sin = rope_cos_sin(d, pos, self.dtype)
        K_llm, V_llm = self.K[d.layers - 1], self.V[d.layers - 1]
        # self attention
        x = rms_norm(h, self.W["glide.input_layernorm.weight"], d.rms_eps)
        q, k, v = self._qkv(x, "glide.self_attn.", True)
        q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
        if exec_type == "prefill":                                        # llama_glide.py:206-233
            self.dK = q.new_zeros(B, S + self.draft_max_len + 128, d.kv_heads, d.head_dim)
            self.dV = q.new_zeros(B, S + self.draft_max_len + 128, d.kv_heads, d.head_dim)
            self.dK[:, :S] = k
            self.dV[:, :S] = v
            a = flash_ref.flash_attn_func(q, k, v, causal=True, window_size=(512, -1))
        elif exec_type == "decoding":
            a = A.draft_self_decoding_attention(q, k, v, self.dK, self.dV, cache_lens)
        elif exec_type == "tree_decoding":
            a = A.draft_self_tree_attention(q, k, v, self.dK, self.dV, cache_lens, tree_mask)
        else:
            raise ValueError(f"Unknown inference_type: {exec_type}")
        a = a.reshape(B, S, d.heads * d.head_dim).to(h.dtype)
        h = h + F.linear(a, self.W["glide.self_attn.o_proj.weight"])
        # cross attention over the target's last-layer KV (k/v proj are computed and
        # discarded by the reference, llama_glide.py:248-249; they do not affect results)
        x = rms_norm(h, self.W["glide.post_self_attention_layernorm.weight"], d.rms_eps)
        q = F.linear(x, self.W["glide.cross_attn.q_proj.weight"],
                     self.W["glide.cross_attn.q_proj.bias"]).view(B, S, d.heads, d.head_dim)
        q = apply_rope(q, cos, sin)
        if exec_type in ("prefill", "decoding"):
            a = A.draft_cross_decoding_attention(q, K_llm, V_llm, llm_kv_len)
        else:
            a = A.draft_cross_tree_attention(q, K_llm, V_llm, llm_kv_len)
        a = a.reshape(B, S, d.heads * d.head_dim).to(h.dtype)
        h = h + F.linear(a, self.W["glide.cross_attn.o_proj.weight"])
        x = rms_norm(h, self.W["glide.post_cross_attention_layernorm.weight"], d.rms_eps)
        return h + mlp(x, self.W, "glide.mlp.")

"""Run the UNMODIFIED reference (sail-sg/LongSpec, /root/reference/longspec/test) on CPU.

TEST INFRASTRUCTURE ONLY.  This file exists to *generate* the golden fixtures
committed under tests/golden/*.npz (see gen_golden.py) and to cross-check the
oracle in this container.  /root/reference does not exist on the GPU box, so
nothing in the `-m gpu` tests, smoke() or bench.py imports this module.

The reference cannot run as shipped in this image:
flash-attn/Triton kernels are CUDA-only, `LlamaGlide.__init__` needs HF weights
and `accelerate`, transformers 5.x dropped `config.rope_theta`, and
`GlideAttention.set_torch_mask` calls `.cuda()`.  The shim below supplies

  * torch-eager stand-ins for `flash_attn.flash_attn_func` and
    `flash_attn.flash_attn_with_kvcache` written to flash-attn 2.6.3's published
    semantics (bottom-right causal alignment, local window rule
    `i + sk - sq - left <= j <= i + sk - sq + right`, in-place append at
    `cache_seqlens`, LSE of shape (B, H, Sq) in natural log),
  * `GlideAttention.triton_tree_part_fwd = GlideAttention.tree_part_fwd`
    (the reference's own non-Triton twin, llama_glide.py:331-359),
  * identity `.cuda()`, no-op `torch.cuda.synchronize`, TORCHDYNAMO_DISABLE=1,
  * a constructor bypass that builds the object graph without from_pretrained.

All model code that executes is the reference's own.
"""
from __future__ import annotations

import math
import os
import sys
import types

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

import torch

REF_TEST_DIR = "/root/reference/longspec/test"


def reference_available() -> bool:
    return os.path.isdir(REF_TEST_DIR)


# --------------------------------------------------------------------------
# flash-attn stand-ins (third-party dependency flash_attn==2.6.3, absent from
# /root/reference; restated from its documented interface)
# --------------------------------------------------------------------------
def _attn_core(q, k, v, softmax_scale, causal, window_size, sk_valid):
    """q (B,Sq,H,D); k,v (B,Sk,Hk,D); sk_valid (B,) number of valid keys."""
    B, Sq, H, D = q.shape
    Hk = k.shape[2]
    G = H // Hk
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(D)
    outs, lses = [], []
    for b in range(B):
        sk = int(sk_valid[b])
        qb = q[b].float().transpose(0, 1)                     # H,Sq,D
        kb = k[b, :sk].float().transpose(0, 1)                # Hk,sk,D
        vb = v[b, :sk].float().transpose(0, 1)
        kb = kb.repeat_interleave(G, dim=0)
        vb = vb.repeat_interleave(G, dim=0)
        s = torch.matmul(qb, kb.transpose(1, 2)) * softmax_scale   # H,Sq,sk
        i = torch.arange(Sq).view(-1, 1)
        j = torch.arange(sk).view(1, -1)
        left, right = window_size
        if causal:
            right = 0
        allow = torch.ones(Sq, sk, dtype=torch.bool)
        if left >= 0:
            allow &= j >= i + sk - Sq - left
        if right >= 0:
            allow &= j <= i + sk - Sq + right
        s = s.masked_fill(~allow[None], float("-inf"))
        lse = torch.logsumexp(s, dim=-1)                       # H,Sq
        p = torch.softmax(s, dim=-1)
        p = torch.nan_to_num(p, nan=0.0)
        o = torch.matmul(p, vb)                                # H,Sq,D
        outs.append(o.transpose(0, 1).to(q.dtype))
        lses.append(lse)
    return torch.stack(outs, 0), torch.stack(lses, 0)


def flash_attn_func(q, k, v, dropout_p=0.0, softmax_scale=None, causal=False,
                    window_size=(-1, -1), **kw):
    B, Sk = k.shape[0], k.shape[1]
    o, _ = _attn_core(q, k, v, softmax_scale, causal, window_size,
                      torch.full((B,), Sk, dtype=torch.int64))
    return o


def flash_attn_with_kvcache(q, k_cache, v_cache, k=None, v=None, rotary_cos=None,
                            rotary_sin=None, cache_seqlens=None, cache_batch_idx=None,
                            cache_leftpad=None, block_table=None, softmax_scale=None,
                            causal=False, window_size=(-1, -1), softcap=0.0,
                            rotary_interleaved=True, alibi_slopes=None, num_splits=0,
                            return_softmax_lse=False):
    B = q.shape[0]
    if cache_seqlens is None:
        lens = torch.full((B,), k_cache.shape[1], dtype=torch.int64)
    elif isinstance(cache_seqlens, int):
        lens = torch.full((B,), cache_seqlens, dtype=torch.int64)
    else:
        lens = cache_seqlens.to(torch.int64).clone()
    if k is not None:
        sn = k.shape[1]
        for b in range(B):
            k_cache[b, lens[b]:lens[b] + sn] = k[b]
            v_cache[b, lens[b]:lens[b] + sn] = v[b]
        lens = lens + sn
    o, lse = _attn_core(q, k_cache, v_cache, softmax_scale, causal, window_size, lens)
    if return_softmax_lse:
        return o, lse
    return o


_INSTALLED = False


def install():
    """Install the shim and put the reference's test dir on sys.path."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REF_TEST_DIR)
    fa = types.ModuleType("flash_attn")
    fa.flash_attn_func = flash_attn_func
    fa.flash_attn_with_kvcache = flash_attn_with_kvcache
    fa.__version__ = "2.6.3-shim"
    import importlib.machinery
    fa.__spec__ = importlib.machinery.ModuleSpec("flash_attn", None)
    sys.modules["flash_attn"] = fa

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None

    # transformers 5.x: the Qwen2 twin indexes ROPE_INIT_FUNCTIONS["default"]
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    if "default" not in ROPE_INIT_FUNCTIONS:
        def _default_rope(config, device=None, seq_len=None, **kw):
            base = config.rope_theta
            dim = getattr(config, "head_dim", None) or config.hidden_size // config.num_attention_heads
            inv = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
            return inv, 1.0
        ROPE_INIT_FUNCTIONS["default"] = _default_rope

    if REF_TEST_DIR not in sys.path:
        sys.path.insert(0, REF_TEST_DIR)
    _INSTALLED = True


def make_config(hidden=256, heads=2, kv_heads=1, layers=2, ffn=512, vocab=320,
                rope_theta=10000.0, max_pos=4096, eos=None, pad=None, qwen=False):
    if qwen:
        from transformers import Qwen2Config as C
    else:
        from transformers import LlamaConfig as C
    cfg = C(hidden_size=hidden, intermediate_size=ffn, num_hidden_layers=layers,
            num_attention_heads=heads, num_key_value_heads=kv_heads, vocab_size=vocab,
            max_position_embeddings=max_pos, rms_norm_eps=1e-5,
            pad_token_id=vocab - 1 if pad is None else pad,
            eos_token_id=vocab - 2 if eos is None else eos,
            bos_token_id=1, tie_word_embeddings=False)
    cfg.rope_theta = rope_theta
    cfg.attention_dropout = 0.0
    cfg.attention_bias = False
    cfg.mlp_bias = False
    cfg.initializer_range = 0.08
    cfg._attn_implementation = "eager"
    return cfg


def build_llama_glide(cfg, state_dict=None, dtype=torch.float32):
    """Build reference LlamaGlide without from_pretrained (llama_glide.py:472-490 bypass)."""
    install()
    import llama_glide as ref_lg
    import llama as ref_llama
    ref_lg.GlideAttention.triton_tree_part_fwd = ref_lg.GlideAttention.tree_part_fwd
    obj = ref_lg.LlamaGlide.__new__(ref_lg.LlamaGlide)
    ref_llama.LlamaForCausalLM.__init__(obj, cfg)
    obj.glide = ref_lg.LlamaGlideDecoderLayer(cfg)
    if state_dict is not None:
        missing, unexpected = obj.load_state_dict(state_dict, strict=False)
        missing = [m for m in missing if "rotary_emb" not in m]
        assert not missing and not unexpected, (missing, unexpected)
    obj.to(dtype)
    obj.eval()
    return obj


def build_qwen2_glide(cfg, state_dict=None, dtype=torch.float32):
    install()
    import qwen2_glide as ref_qg
    import qwen2 as ref_qwen
    ref_qg.GlideAttention.triton_tree_part_fwd = ref_qg.GlideAttention.tree_part_fwd
    obj = ref_qg.Qwen2Glide.__new__(ref_qg.Qwen2Glide)
    ref_qwen.Qwen2ForCausalLM.__init__(obj, cfg)
    obj.glide = ref_qg.Qwen2GlideDecoderLayer(cfg)
    if state_dict is not None:
        missing, unexpected = obj.load_state_dict(state_dict, strict=False)
        missing = [m for m in missing if "rotary_emb" not in m]
        assert not missing and not unexpected, (missing, unexpected)
    obj.to(dtype)
    obj.eval()
    return obj

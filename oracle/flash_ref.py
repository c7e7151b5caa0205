"""Restatement of flash_attn==2.6.3 call semantics used by the reference hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  flash-attn is a third-party
dependency (longspec/test/requirements.txt:2) absent from /root/reference; its
published interface (flash_attn_interface.py docstrings of `flash_attn_func` and
`flash_attn_with_kvcache`) defines:

  * q (B, Sq, H, D); k/v (B, Sk, Hk, D); H % Hk == 0 (GQA: q head h uses kv head
    h // (H/Hk));
  * scale default 1/sqrt(D); softmax in fp32; output in q.dtype;
  * causal mask aligned to the bottom-right corner: query i sees key j iff
    j <= i + Sk - Sq;
  * local window (left, right): j in [i + Sk - Sq - left, i + Sk - Sq + right],
    -1 meaning unbounded; `causal=True` forces right = 0;
  * kvcache variant: optional new k/v are written in place at `cache_seqlens`
    and the attention runs over the first `cache_seqlens + Sk_new` cache rows;
  * LSE returned as (B, H, Sq) fp32, natural log, of the *scaled* scores.

Reference call sites: llama.py:218,324,382,385; llama_glide.py:227,261,265,297,300.
"""
from __future__ import annotations

import math

import torch


def visible(sq: int, sk: int, causal: bool, window=(-1, -1)) -> torch.Tensor:
    """(Sq, Sk) bool visibility matrix of the flash-attn masking rule."""
    left, right = window
    if causal:
        right = 0
    shift = sk - sq
    i = torch.arange(sq).unsqueeze(1) + shift
    j = torch.arange(sk).unsqueeze(0)
    ok = torch.ones(sq, sk, dtype=torch.bool)
    if left >= 0:
        ok &= j >= i - left
    if right >= 0:
        ok &= j <= i + right
    return ok


def attend(q, k, v, mask=None, scale=None, p_dtype=None):
    """Masked softmax attention for one sequence.

    q (Sq,H,D), k/v (Sk,Hk,D), mask (Sq,Sk) bool (True = visible) or None.
    Returns o (Sq,H,D) fp32 and lse (H,Sq) fp32.  Rows with no visible key give
    o = 0 and lse = -inf (flash-attn's convention).  `p_dtype` optionally rounds
    the probabilities before PV (flash-attn multiplies P in the input dtype).
    """
    Sq, H, D = q.shape
    Sk, Hk, _ = k.shape
    G = H // Hk
    if scale is None:
        scale = 1.0 / math.sqrt(D)
    qf = q.float().permute(1, 0, 2)                                   # H,Sq,D
    kf = k.float().permute(1, 0, 2).repeat_interleave(G, dim=0)       # H,Sk,D
    vf = v.float().permute(1, 0, 2).repeat_interleave(G, dim=0)
    s = torch.matmul(qf, kf.transpose(1, 2)) * scale                  # H,Sq,Sk
    if mask is not None:
        s = s.masked_fill(~mask.unsqueeze(0), float("-inf"))
    if Sk == 0:
        return torch.zeros(Sq, H, D), torch.full((H, Sq), float("-inf"))
    m = s.max(dim=-1, keepdim=True).values
    m_safe = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    e = torch.exp(s - m_safe)
    l = e.sum(dim=-1, keepdim=True)
    lse = (m_safe + torch.log(l)).squeeze(-1)
    lse = torch.where(l.squeeze(-1) == 0, torch.full_like(lse, float("-inf")), lse)
    if p_dtype is not None:
        e = e.to(p_dtype).float()
    o = torch.matmul(e, vf) / torch.where(l == 0, torch.ones_like(l), l)
    return o.permute(1, 0, 2).contiguous(), lse


def flash_attn_func(q, k, v, causal=False, window_size=(-1, -1), softmax_scale=None):
    """q (B,Sq,H,D) -> out (B,Sq,H,D) in q.dtype."""
    outs = []
    for b in range(q.shape[0]):
        mask = visible(q.shape[1], k.shape[1], causal, window_size)
        o, _ = attend(q[b], k[b], v[b], mask, softmax_scale)
        outs.append(o.to(q.dtype))
    return torch.stack(outs)


def flash_attn_with_kvcache(q, k_cache, v_cache, k=None, v=None, cache_seqlens=None,
                            causal=False, window_size=(-1, -1), softmax_scale=None,
                            return_softmax_lse=False):
    """In-place append + attention over the valid cache prefix."""
    B = q.shape[0]
    lens = [int(x) for x in (cache_seqlens.tolist() if torch.is_tensor(cache_seqlens)
                             else [cache_seqlens] * B)]
    outs, lses = [], []
    for b in range(B):
        n = lens[b]
        if k is not None:
            k_cache[b, n:n + k.shape[1]] = k[b]
            v_cache[b, n:n + v.shape[1]] = v[b]
            n += k.shape[1]
        mask = visible(q.shape[1], n, causal, window_size)
        o, lse = attend(q[b], k_cache[b, :n], v_cache[b, :n], mask, softmax_scale)
        outs.append(o.to(q.dtype))
        lses.append(lse)
    out = torch.stack(outs)
    if return_softmax_lse:
        return out, torch.stack(lses)
    return out

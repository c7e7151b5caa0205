"""Oracle restatement of the attention operators on the tree-spec hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Tensors keep the reference's seam
layout: activations (B, S, H, D), caches (B, S_max, Hk, D), tree masks
(B, M, N) integer with 1 = visible.
"""
from __future__ import annotations

import math

import torch

from . import flash_ref


# ---------------------------------------------------------------------------
# triton_tree_attn.attention  (triton_tree_attn.py:19-77, kernel 115-251)
# ---------------------------------------------------------------------------
def tree_mask_attention(q, k, v, tree_mask, sm_scale=None):
    """Dense-mask attention of the reference Triton kernel.

    q (B,H,M,D); k,v (B,Hk,N,D); tree_mask (B,M,N) integer/bool, non-zero = visible.
    The kernel ANDs the mask with a bottom-right causal mask
    `P_SEQ + m >= n`, P_SEQ = N - M (triton_tree_attn.py:31,213) and returns
    o in q.dtype plus L = log-normaliser (natural log of the scaled scores,
    triton_tree_attn.py:243) as (B,H,M) fp32.
    """
    B, H, M, D = q.shape
    N = k.shape[2]
    if sm_scale is None:
        sm_scale = 1.0 / math.sqrt(D)
    causal = (torch.arange(M).unsqueeze(1) + (N - M)) >= torch.arange(N).unsqueeze(0)
    o = torch.empty(B, H, M, D, dtype=q.dtype)
    L = torch.empty(B, H, M, dtype=torch.float32)
    for b in range(B):
        vis = (tree_mask[b] != 0) & causal
        ob, lse = flash_ref.attend(q[b].permute(1, 0, 2), k[b].permute(1, 0, 2),
                                   v[b].permute(1, 0, 2), vis, sm_scale)
        o[b] = ob.permute(1, 0, 2).to(q.dtype)
        L[b] = lse
    return o, L


# ---------------------------------------------------------------------------
# Target model: LlamaAttention.tree_decoding / tree_part_fwd (llama.py:357-421)
# ---------------------------------------------------------------------------
def target_tree_part(q, k_new, v_new, K_cache, V_cache, cache_lens, tree_mask,
                     prefix_lse, last_layer, scale=1.0 / math.sqrt(128)):
    """llama.py:394-421.  Appends this round's K/V at cache_lens+[0,q_len) (396-399),
    masked attention of the q_len queries over *this round's* q_len keys only,
    returns (current_out (B,q,H,D) in q.dtype, weight (B,q,H,1) in q.dtype) with
    weight = sigmoid(prefix_lse - current_lse) (420).  The last layer scales Q
    before the matmul, the others after (406-409) - identical in exact
    arithmetic, different roundings in fp16/bf16.
    """
    B, S, H, D = q.shape
    Hk = k_new.shape[2]
    G = H // Hk
    outs, weights = [], []
    for b in range(B):
        n0 = int(cache_lens[b])
        K_cache[b, n0:n0 + S] = k_new[b]
        V_cache[b, n0:n0 + S] = v_new[b]
        kq = k_new[b].repeat_interleave(G, dim=1).permute(1, 2, 0)      # H,D,S
        vq = v_new[b].repeat_interleave(G, dim=1).permute(1, 0, 2)      # H,S,D
        qq = q[b].permute(1, 0, 2)                                      # H,S,D
        if last_layer:
            score = torch.matmul(qq * scale, kq)
        else:
            score = torch.matmul(qq, kq) * scale
        score = score.float().masked_fill((tree_mask[b] == 0).unsqueeze(0), float("-inf"))
        p = torch.softmax(score, dim=-1).to(q.dtype)
        cur = torch.matmul(p, vq).permute(1, 0, 2)                      # S,H,D
        cur_lse = torch.logsumexp(score, dim=-1).transpose(0, 1).unsqueeze(-1)   # S,H,1
        pre = prefix_lse[b].transpose(0, 1).unsqueeze(-1)              # S,H,1
        outs.append(cur)
        weights.append(torch.sigmoid(pre - cur_lse).to(q.dtype))
    return torch.stack(outs), torch.stack(weights)


def target_hybrid_tree_attention(q, k_new, v_new, K_cache, V_cache, cache_lens, tree_mask,
                                 last_layer=False):
    """LlamaAttention.tree_decoding after RoPE, before o_proj (llama.py:380-389).

    prefix: non-causal FlashDecoding over cache[0:cache_lens) with LSE (385);
    tree part (386); merge `prefix_o*w + current_out*(1-w)` in q.dtype (387).
    Returns attn (B, q_len, H, D) in q.dtype.  Mutates the caches like the reference.
    """
    prefix_o, prefix_lse = flash_ref.flash_attn_with_kvcache(
        q, K_cache, V_cache, cache_seqlens=cache_lens, return_softmax_lse=True)
    cur, w = target_tree_part(q, k_new, v_new, K_cache, V_cache, cache_lens, tree_mask,
                              prefix_lse, last_layer)
    return prefix_o * w + cur * (1 - w)


def single_softmax_tree_attention(q, k_new, v_new, K_cache, V_cache, cache_lens, tree_mask,
                                  scale=1.0 / math.sqrt(128)):
    """train/models/llama.py:210-275 (`tree_decoding_torch`): one softmax over
    [cache[:cache_lens] || new K] with the tree mask on the new block.  Equals the
    hybrid form in exact arithmetic; used to cross-check the LSE merge."""
    B, S, H, D = q.shape
    outs = []
    for b in range(B):
        n0 = int(cache_lens[b])
        K_cache[b, n0:n0 + S] = k_new[b]
        V_cache[b, n0:n0 + S] = v_new[b]
        vis = torch.cat([torch.ones(S, n0, dtype=torch.bool), tree_mask[b] != 0], dim=1)
        o, _ = flash_ref.attend(q[b], K_cache[b, :n0 + S], V_cache[b, :n0 + S], vis, scale)
        outs.append(o.to(q.dtype))
    return torch.stack(outs)


# ---------------------------------------------------------------------------
# Draft: GlideAttention.decoding / tree_decoding (llama_glide.py:235-329)
# ---------------------------------------------------------------------------
def draft_self_decoding_attention(q, k_new, v_new, K_cache, V_cache, cache_lens):
    """sa_decoding (llama_glide.py:258-262): append k/v at cache_lens, causal
    attention with left window 512 over the draft's own cache."""
    return flash_ref.flash_attn_with_kvcache(q, K_cache, V_cache, k_new, v_new,
                                             cache_seqlens=cache_lens, causal=True,
                                             window_size=(512, -1))


def draft_cross_decoding_attention(q, K_llm, V_llm, llm_kv_len):
    """ca_decoding / ca_prefill (llama_glide.py:263-265): bottom-right causal
    attention over the target's last-layer cache[0:llm_kv_len)."""
    return flash_ref.flash_attn_with_kvcache(q, K_llm, V_llm, cache_seqlens=llm_kv_len,
                                             causal=True)


def draft_cross_tree_attention(q, K_llm, V_llm, llm_kv_len):
    """ca_tree_decoding (llama_glide.py:296-297): NON-causal over cache[0:llm_kv_len)."""
    return flash_ref.flash_attn_with_kvcache(q, K_llm, V_llm, cache_seqlens=llm_kv_len,
                                             causal=False)


def draft_self_tree_attention(q, k_new, v_new, K_cache, V_cache, cache_lens, tree_mask):
    """sa_tree_decoding (llama_glide.py:299-302 + triton_tree_part_fwd 309-329).

    tree_mask is the level's slice (B, M, N_tree) with N_tree = all tree tokens so
    far (root included).  prefix = non-causal, left-window-512 FlashDecoding over
    the draft cache[0:cache_lens) (so row i of M sees the last 512 + M - 1 - i ...
    keys by flash-attn's local rule); the level's K/V are written at
    cache_lens + [N_tree - M, N_tree) (312-315), all tree K/V are gathered from
    cache_lens + [0, N_tree) (317-319) and attended with the mask (320-325);
    merge in fp32: out = prefix_o.float()*w + current*(1-w), w = sigmoid(lse_p - lse_c).
    Returns fp32 (B, M, H, D) like the reference (cast to model dtype by the caller, 304).
    """
    B, M, H, D = q.shape
    N_tree = tree_mask.shape[-1]
    prefix_o, prefix_lse = flash_ref.flash_attn_with_kvcache(
        q, K_cache, V_cache, cache_seqlens=cache_lens, window_size=(512, -1),
        return_softmax_lse=True)
    outs = []
    for b in range(B):
        n0 = int(cache_lens[b])
        K_cache[b, n0 + N_tree - M:n0 + N_tree] = k_new[b]
        V_cache[b, n0 + N_tree - M:n0 + N_tree] = v_new[b]
        kt = K_cache[b, n0:n0 + N_tree].permute(1, 0, 2).unsqueeze(0)    # 1,Hk,N,D
        vt = V_cache[b, n0:n0 + N_tree].permute(1, 0, 2).unsqueeze(0)
        cur, cur_lse = tree_mask_attention(q[b].permute(1, 0, 2).unsqueeze(0), kt, vt,
                                           tree_mask[b:b + 1])
        w = torch.sigmoid(prefix_lse[b] - cur_lse[0])                    # H,M
        w = w.transpose(0, 1).unsqueeze(-1)                              # M,H,1
        cur = cur[0].permute(1, 0, 2)                                    # M,H,D
        outs.append(prefix_o[b].float() * w + cur * (1 - w))
    return torch.stack(outs)


def merge_partials(part_o, part_lse):
    """Associative combine of split-KV partials: part_o (P, rows, D) normalised outputs, part_lse (P, rows)
    natural-log normalisers (-inf for an empty part).  The same operator the reference uses between its
    prefix and tree parts (weight = sigmoid(lse_a - lse_b), llama.py:387,420), generalised to P parts."""
    mx = part_lse.max(dim=0).values
    mx_safe = torch.where(torch.isinf(mx), torch.zeros_like(mx), mx)
    w = torch.exp(part_lse - mx_safe)
    w = torch.where(torch.isinf(part_lse), torch.zeros_like(w), w)
    den = w.sum(0)
    out = (w.unsqueeze(-1) * part_o).sum(0) / torch.where(den == 0, torch.ones_like(den), den).unsqueeze(-1)
    return out

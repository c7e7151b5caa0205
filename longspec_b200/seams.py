"""Reference-shaped operator seams (seams 3 and 4).

The decode loop of this package never goes through these wrappers - it keeps its state on the device in the paged
layout and calls the C ABI directly.  They exist so that a maintainer of the REFERENCE can swap a single call site for
the CUDA kernels without adopting the paged cache: same argument order, same tensor layouts, same return tuples as

    triton_tree_attn.attention(q, k, v, tree_mask, sm_scale=None, return_log_normalizer=True) -> (o, L)
        (longspec/test/triton_tree_attn.py:19-77; called from llama_glide.py:318)
    LlamaGlide.tree_verification(input_ids, output_ids, tree_mask, cache_lens, non_leaf_len) -> (acc_ids, acc_num, double_input)
        (longspec/test/llama_glide.py:1128-1175), which also compacts the last layer's K_Cache / V_Cache in place

Both adapt dense reference tensors to the kernels' layout (a few small copies) and run the same kernels as the loop:
ls_pack_mask + ls_attn_partial + ls_attn_combine, and ls_tree_accept.  There is no torch fallback.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from ._lib import AttnDesc, LS_MASK_WORDS, LS_PAGE
from .engine import PagedKV, _code, _ptr, _stream, alloc_tree_state

__all__ = ["attention", "tree_verification"]


def attention(q, k, v, tree_mask, sm_scale=None, return_log_normalizer=True):
    """Drop-in for triton_tree_attn.attention.

    q (B, H, M, D), k / v (B, Hk, N, D) fp16 / bf16 CUDA tensors, D = 128, H % Hk == 0;
    tree_mask (B, M, N) or (M, N) integer / bool, non-zero = visible.  As in the reference kernel a key is visible to
    query m iff tree_mask[m, n] != 0 AND n <= (N - M) + m (triton_tree_attn.py:209-214).
    Returns o (B, H, M, D) in the input dtype and L (B, H, M) fp32 = natural-log normaliser (m * scale + log l, :243).
    """
    assert q.shape[-1] == k.shape[-1] == v.shape[-1], "feature size of q, k, v should be equal"
    B, H, M, D = q.shape
    Hk, N = k.shape[1], k.shape[2]
    assert v.shape[1] == Hk, "num of heads in k and v should be equal"
    assert H % Hk == 0, "number of heads in q must be a multiple of that in k & v"
    if D != 128:
        raise NotImplementedError("the sm_100a kernels are built for head_dim 128 (every model the reference ships)")
    if N > 64 * LS_MASK_WORDS or M > 64 * LS_MASK_WORDS:
        raise NotImplementedError(f"tree masks are packed into {64 * LS_MASK_WORDS}-bit rows: M, N <= {64 * LS_MASK_WORDS}")
    if not q.is_cuda:
        raise _lib.LongSpecError("attention() needs CUDA tensors: there is no CPU path")
    if sm_scale is None:
        sm_scale = 1.0 / math.sqrt(D)
    dev, dt = q.device, q.dtype
    G = H // Hk
    mask = tree_mask if tree_mask.dim() == 3 else tree_mask[None].expand(B, -1, -1)
    o = torch.empty_like(q)
    L = torch.empty(B, H, M, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        zero = torch.zeros(1, dtype=torch.int32, device=dev)
        n_q = torch.tensor([M], dtype=torch.int32, device=dev)
        causal = (torch.arange(N, device=dev)[None, :] <= (N - M) + torch.arange(M, device=dev)[:, None])
        kv = PagedKV(1, Hk, N + LS_PAGE, dt, dev)
        bits = torch.empty(M * LS_MASK_WORDS, dtype=torch.int64, device=dev)
        out = torch.empty(M, H, 128, dtype=dt, device=dev)
        part_o = torch.empty(Hk * M * G * 128, dtype=torch.float32, device=dev)
        part_l = torch.empty(Hk * M * G, dtype=torch.float32, device=dev)
        for b in range(B):
            kv.write(0, k[b].permute(1, 0, 2).contiguous(), v[b].permute(1, 0, 2).contiguous(), 0)
            dense = ((mask[b] != 0) & causal).to(torch.int64).contiguous()
            _lib.call("ls_pack_mask", _ptr(dense), M, N, dense.stride(0), _ptr(bits), _stream())
            qb = q[b].permute(1, 0, 2).contiguous()                       # [M][H][128]
            a = AttnDesc()
            a.q, a.out = _ptr(qb), _ptr(out)
            a.k_pool, a.v_pool, a.page_table = _ptr(kv.k[0]), _ptr(kv.v[0]), _ptr(kv.page_table)
            a.kv_len_ptr, a.n_q_ptr, a.tail_mask = _ptr(zero), _ptr(n_q), _ptr(bits)
            a.part_o, a.part_lse = _ptr(part_o), _ptr(part_l)
            a.kv_len_bias, a.q_rows, a.n_heads, a.n_kv_heads = 0, M, H, Hk
            a.n_tail, a.window, a.causal, a.dtype = N, -1, 0, _code(dt)
            a.n_splits, a.n_pages = 1, kv.n_pages
            a.part_offset, a.n_parts_total = 0, 1
            a.kv_shard_begin, a.kv_shard_end, a.tail_owner = 0, -1, 1
            a.scale, a.q_stride = float(sm_scale), 0
            _lib.call("ls_attn_partial", C.byref(a), _stream())
            _lib.call("ls_attn_combine", C.byref(a), _stream())
            o[b] = out.permute(1, 0, 2)
            # one split: the partial LSE [kv head][row = m * G + g] IS the log-normaliser of head kvh * G + g
            L[b] = part_l.view(Hk, M, G).permute(0, 2, 1).reshape(H, M)
    return (o, L) if return_log_normalizer else o


def tree_verification(input_ids, output_ids, tree_mask, cache_lens, non_leaf_len, K_Cache=None, V_Cache=None):
    """Drop-in for LlamaGlide.tree_verification (bsz = 1, like every shipped driver).

    input_ids (1, fseq) int64 = tree tokens, output_ids (1, fseq) int64 = the target's argmax at every node,
    tree_mask (1, fseq, fseq) integer ancestor mask, cache_lens (1,) = slot of the root (already += acc_num_prev - 1,
    llama_glide.py:1104), non_leaf_len as at llama_glide.py:1107.  K_Cache / V_Cache: the LAST target layer's dense
    caches (B, S, Hk, 128) - the module attributes the reference method reads (llama_glide.py:1159-1173) - whose rows
    cache_lens + selected indices are compacted to cache_lens + [0, acc_num) in place; None skips the compaction.
    Returns (acc_ids (1, acc_num) int64, acc_num (1,) int64, double_input (1,) int32) like llama_glide.py:1175.
    """
    assert input_ids.size(0) == 1, "batch size 1 (inference_long-bench.py:119)"
    if not input_ids.is_cuda:
        raise _lib.LongSpecError("tree_verification() needs CUDA tensors: there is no CPU path")
    dev = input_ids.device
    fseq = input_ids.shape[1]
    if fseq > 64 * LS_MASK_WORDS:
        raise NotImplementedError(f"trees of up to {64 * LS_MASK_WORDS} nodes")
    cache_len = int(cache_lens.reshape(-1)[0])
    with torch.cuda.device(dev):
        # a one-level tree state of the right size; prompt_len = cache_len so that the accepted ids land at output slot 1..
        # (a chain-shaped state: acc_ids is sized by the depth, and the depth of the caller's tree is not known here;
        # the kernels take at most 16 levels; a real tree is never deeper than gamma + 1 <= 17 nodes)
        n = max(fseq - 1, 1)
        shape = [1] * n if n <= 16 else [1] * 15 + [n - 15]
        st, ts = alloc_tree_state(dev, shape, cache_len, fseq + 8)
        ts.fseq = fseq
        tm = (tree_mask[0] != 0)
        dense = tm.to(torch.int64).contiguous()
        _lib.call("ls_pack_mask", _ptr(dense), fseq, fseq, dense.stride(0), _ptr(st["tree_bits"]), _stream())
        # parent = largest-index proper ancestor (llama_glide.py:1136: argmax((tree_mask - I) * index))
        idx = torch.arange(fseq, device=dev)
        st["parent"][:fseq] = ((dense - torch.eye(fseq, dtype=torch.int64, device=dev)) * idx[None, :]).argmax(-1).to(torch.int32)
        st["all_spec"][:fseq] = input_ids[0]
        st["scalars"][:4] = torch.tensor([cache_len, cache_len, cache_len, 1], dtype=torch.int32, device=dev)
        pred = output_ids[0].contiguous()
        kv = None
        if K_Cache is not None:
            Hk = K_Cache.shape[2]
            kv = PagedKV(1, Hk, cache_len % LS_PAGE + fseq + LS_PAGE, K_Cache.dtype, dev)
            base = cache_len - cache_len % LS_PAGE                         # keep the slot-in-page alignment of the real cache
            kv.write(0, K_Cache[0, cache_len:cache_len + fseq], V_Cache[0, cache_len:cache_len + fseq], cache_len - base)
            _lib.call("ls_tree_accept", C.byref(ts), _ptr(pred), _ptr(kv.k[0]), _ptr(kv.v[0]), _ptr(kv.page_table), Hk,
                      _code(K_Cache.dtype), int(non_leaf_len), -1, -base, _stream())
        else:
            _lib.call("ls_tree_accept", C.byref(ts), _ptr(pred), C.c_void_p(0), C.c_void_p(0), C.c_void_p(0), 1,
                      _lib.LS_F16, int(non_leaf_len), -1, 0, _stream())
        acc = int(st["scalars"][3])
        if kv is not None:
            Ka, Va = kv.read(0, cache_len - base, acc)
            K_Cache[0, cache_len:cache_len + acc] = Ka
            V_Cache[0, cache_len:cache_len + acc] = Va
        acc_ids = st["acc_ids"][:acc].clone()[None]
        acc_num = torch.tensor([acc], dtype=torch.int64, device=dev)
        double_input = st["aux"][2:3].clone()
    return acc_ids, acc_num, double_input

"""Helpers for the -m gpu parity tests: call the C ABI on dense test tensors."""
import ctypes as C
import math

import torch

from longspec_b200 import _lib
from longspec_b200._lib import AttnDesc, LS_MASK_WORDS, LS_PAGE
from longspec_b200.engine import PagedKV, _ptr, _stream, _code


def pack_bits(mask: torch.Tensor) -> torch.Tensor:
    """(rows, cols) 0/1 -> (rows*LS_MASK_WORDS,) int64 bit rows, computed on the host."""
    rows, cols = mask.shape
    out = torch.zeros(rows, LS_MASK_WORDS, dtype=torch.int64)
    m = mask.cpu()
    for r in range(rows):
        for c in range(cols):
            if int(m[r, c]) != 0:
                w, b = c >> 6, c & 63
                v = int(out[r, w]) | (1 << b)
                if v >= 1 << 63:
                    v -= 1 << 64
                out[r, w] = v
    return out.reshape(-1)


def paged_from_dense(K, V, dtype, device, shuffle=True):
    """K, V [S][Hk][128] -> PagedKV (one layer) holding them."""
    S, Hk, _ = K.shape
    kv = PagedKV(1, Hk, S + LS_PAGE, dtype, device, shuffle_pages=shuffle)
    kv.write(0, K.to(device=device, dtype=dtype), V.to(device=device, dtype=dtype), 0)
    return kv


def run_attention(q, kv, kv_len, n_q, *, n_tail=0, tail_mask=None, window=-1, causal=0, n_splits=3,
                  shard=(0, -1, 1), world_parts=None, entry="ls_attn_partial", layer=0, fuse=False):
    """q [q_rows][H][128] on the GPU.  Returns out [q_rows][H][128]."""
    device, dtype = q.device, q.dtype
    q_rows, H, _ = q.shape
    Hk = kv.k[layer].shape[1]
    G = H // Hk
    out = torch.full((q_rows, H, 128), float("nan"), dtype=dtype, device=device)
    n_parts = n_splits if world_parts is None else world_parts
    part_o = torch.full((n_parts * Hk * q_rows * G * 128,), float("nan"), dtype=torch.float32, device=device)
    part_l = torch.full((n_parts * Hk * q_rows * G,), float("nan"), dtype=torch.float32, device=device)
    kvl = torch.tensor([kv_len], dtype=torch.int32, device=device)
    nq = torch.tensor([n_q], dtype=torch.int32, device=device)
    bits = None
    if tail_mask is not None:
        bits = pack_bits(tail_mask).to(device)
    a = AttnDesc()
    strided = q.stride(0) != H * 128 and q.stride(1) == 128 and q.stride(2) == 1
    a.q, a.out = _ptr(q if strided else q.contiguous()), _ptr(out)
    a.q_stride = q.stride(0) if strided else 0
    a.k_pool, a.v_pool, a.page_table = _ptr(kv.k[layer]), _ptr(kv.v[layer]), _ptr(kv.page_table)
    a.kv_len_ptr, a.n_q_ptr, a.tail_mask = _ptr(kvl), _ptr(nq), _ptr(bits)
    a.part_o, a.part_lse = _ptr(part_o), _ptr(part_l)
    a.kv_len_bias, a.q_rows, a.n_heads, a.n_kv_heads = 0, q_rows, H, Hk
    a.n_tail, a.window, a.causal, a.dtype = n_tail, window, causal, _code(dtype)
    a.n_splits, a.n_pages = n_splits, kv.n_pages
    a.part_offset, a.n_parts_total = 0, n_parts
    a.kv_shard_begin, a.kv_shard_end, a.tail_owner = shard
    a.scale = 1.0 / math.sqrt(128)
    sync = None
    if fuse:                                 # fused split-KV combine: the attention kernel writes `out` itself
        sync = torch.zeros(2 * Hk, dtype=torch.int32, device=device)
        a.fuse_sync = _ptr(sync)
        assert _lib.load().ls_attn_fuses_combine(C.byref(a)) == 1, "this call was expected to fuse its combine"
    _lib.call(entry, C.byref(a), _stream())
    if not fuse:
        _lib.call("ls_attn_combine", C.byref(a), _stream())
    torch.cuda.synchronize()
    if fuse:
        assert int(sync.abs().sum()) == 0, "the fused combine must leave its counters at zero for the next launch"
    return out, part_o, part_l

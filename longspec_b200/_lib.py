"""ctypes binding of liblongspec_b200.so (the C ABI declared in include/longspec_b200.h).

There is no fallback: if the shared library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# LONGSPEC_LIB points the binding at an alternative build of the SAME ABI (tools/experimental/build_n128.sh, trace builds)
LIB_PATH = os.environ.get("LONGSPEC_LIB") or os.path.join(HERE, "liblongspec_b200.so")

LS_PAGE = 128
LS_HEAD_DIM = 128
LS_MASK_WORDS = 4
LS_MAX_TOPK = 32
LS_F16, LS_BF16 = 0, 1

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", vp), ("out", vp), ("k_pool", vp), ("v_pool", vp), ("page_table", vp), ("kv_len_ptr", vp),
        ("n_q_ptr", vp), ("tail_mask", vp), ("part_o", vp), ("part_lse", vp),
        ("kv_len_bias", i32), ("q_rows", i32), ("n_heads", i32), ("n_kv_heads", i32), ("n_tail", i32),
        ("window", i32), ("causal", i32), ("dtype", i32), ("n_splits", i32), ("n_pages", i32),
        ("part_offset", i32), ("n_parts_total", i32), ("kv_shard_begin", i32), ("kv_shard_end", i32),
        ("tail_owner", i32), ("scale", f32), ("q_stride", i64), ("fuse_sync", vp),
    ]


class TreeState(C.Structure):
    _fields_ = [
        ("cache_len", vp), ("llm_kv_len", vp), ("draft_len", vp), ("acc_num", vp), ("acc_ids", vp),
        ("all_spec", vp), ("scores", vp), ("parent", vp), ("depth", vp), ("tree_bits", vp),
        ("step_ids", vp), ("step_pos", vp), ("step_rows", vp), ("veri_ids", vp), ("veri_pos", vp),
        ("veri_bits", vp), ("veri_rows", vp), ("logit_rows", vp), ("output_ids", vp), ("counters", vp),
        ("aux", vp),
        ("fseq", i32), ("gamma", i32), ("prompt_len", i32), ("max_gen_len", i32),
    ]


LS_MAX_PEERS = 8


class Xchg(C.Structure):
    _fields_ = [
        ("peer_o", vp * LS_MAX_PEERS), ("peer_lse", vp * LS_MAX_PEERS), ("peer_flags", vp * LS_MAX_PEERS),
        ("local_sync", vp), ("block_o", i64), ("block_lse", i64), ("rank", i32), ("world", i32),
    ]


_SIGS = {
    "ls_version": (i32, []),
    "ls_last_error": (C.c_char_p, []),
    "ls_sm_count": (i32, []),
    "ls_attn_partial": (i32, [C.POINTER(AttnDesc), vp]),
    "ls_attn_fuses_combine": (i32, [C.POINTER(AttnDesc)]),
    "ls_attn_tc1_mode": (i32, [i32]),
    "ls_attn_partial_mma": (i32, [C.POINTER(AttnDesc), vp]),
    "ls_attn_partial_tc": (i32, [C.POINTER(AttnDesc), i32, vp]),
    "ls_attn_combine": (i32, [C.POINTER(AttnDesc), vp]),
    "ls_attn_reduce": (i32, [C.POINTER(AttnDesc), vp, vp, vp]),
    "ls_attn_reduce_push": (i32, [C.POINTER(AttnDesc), C.POINTER(Xchg), vp]),
    "ls_attn_combine_wait": (i32, [C.POINTER(AttnDesc), C.POINTER(Xchg), vp]),
    "ls_peer_push": (i32, [vp, i64, i64, C.POINTER(Xchg), vp]),
    "ls_peer_gather_wait": (i32, [vp, i64, i64, C.POINTER(Xchg), vp]),
    "ls_peer_sum_wait": (i32, [vp, i64, i32, i64, C.POINTER(Xchg), vp]),
    "ls_rope_append": (i32, [vp, vp, vp, i64, i64, i64, vp, vp, vp, vp, vp, vp, i32, i32, vp, i32, i32, i32, i32, vp]),
    "ls_round_begin": (i32, [C.POINTER(TreeState), vp]),
    "ls_grow_workspace_bytes": (i64, [i32, i32]),
    "ls_tree_grow": (i32, [C.POINTER(TreeState), vp, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
    "ls_grow_chunk_bytes": (i64, []),
    "ls_tree_grow_partial": (i32, [vp, i32, i32, i32, i64, i32, i32, i32, i32, vp, vp]),
    "ls_tree_grow_finish": (i32, [C.POINTER(TreeState), i32, i32, i32, i32, i32, i32, vp, vp]),
    "ls_argmax_rows_partial": (i32, [vp, i32, i32, i32, i32, i64, i32, vp, vp]),
    "ls_argmax_merge": (i32, [vp, i32, i32, vp, vp]),
    "ls_build_verify": (i32, [C.POINTER(TreeState), vp]),
    "ls_argmax_rows": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "ls_tree_accept": (i32, [C.POINTER(TreeState), vp, vp, vp, vp, i32, i32, i32, i64, i32, vp]),
    "ls_pack_mask": (i32, [vp, i32, i32, i64, vp, vp]),
    "ls_add_rmsnorm": (i32, [vp, vp, vp, vp, i32, i32, f32, i32, vp]),
    "ls_silu_mul": (i32, [vp, vp, i32, i32, i32, vp]),
    "ls_embed": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "ls_gather_rows": (i32, [vp, vp, vp, i32, i32, i32, vp]),
}

_lib = None


class LongSpecError(RuntimeError):
    pass


def exported_symbols():
    return list(_SIGS)


def load():
    """Load the shared library (no compute).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LongSpecError(
                f"{LIB_PATH} not found: build it with `python -m longspec_b200.build` "
                "(there is no CPU/PyTorch fallback for the CUDA kernels)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def call(name, *args):
    """Call an int-returning entry point; raise LongSpecError on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise LongSpecError(f"{name} failed ({rc}): {lib.ls_last_error().decode()}")
    return rc

"""Sequence-sharded prefix KV across the GPUs of one node.

Rank r owns a contiguous chunk of the prompt's KV in every target layer (engine.Shard); weights, the
draft and all GEMMs are replicated, so every rank runs the identical round on identical inputs.  Each
attention call produces per-rank split-KV partials (normalised O + LSE); ONE collective per call makes
all partial slots visible on every rank, and the same deterministic combine kernel merges them, so all
ranks take bit-identical accept decisions without any further synchronisation.

The reference has no counterpart (single GPU, no collectives on the inference path).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def exchange_partials(part_o: torch.Tensor, part_lse: torch.Tensor, n_o: int, n_l: int, rank: int, world: int,
                      group=None, mine=None) -> None:
    """All-gather the partial slots in place.

    part_o / part_lse are flat workspaces in the slot layout of ls_attn_desc; rank r contributes a block of
    n_o / n_l floats (the engine: ONE locally reduced partial per rank, ls_attn_reduce) - either already in
    place at [r * n_o, (r + 1) * n_o) / [r * n_l, (r + 1) * n_l), or in the separate send buffers `mine`.
    After the call every rank holds all `world` blocks in rank order - the layout ls_attn_combine reads.
    """
    o_all = part_o[:n_o * world]
    l_all = part_lse[:n_l * world]
    if mine is not None:            # this rank's block already sits in separate send buffers: no staging copy
        o_mine, l_mine = mine[0][:n_o], mine[1][:n_l]
    else:
        o_mine = o_all[rank * n_o:(rank + 1) * n_o].clone()
        l_mine = l_all[rank * n_l:(rank + 1) * n_l].clone()
    if part_o.is_cuda:
        # one NCCL group launch for both tensors (the exchange is latency-, not bandwidth-bound)
        with dist._coalescing_manager(group=group, device=part_o.device, async_ops=False):
            dist.all_gather_into_tensor(o_all, o_mine, group=group)
            dist.all_gather_into_tensor(l_all, l_mine, group=group)
    else:
        dist.all_gather_into_tensor(o_all, o_mine, group=group)
        dist.all_gather_into_tensor(l_all, l_mine, group=group)


class PeerExchange:
    """The exchange of SURVEY.md section 8(e) option C: every rank's merge kernels store their (O, lse) block straight into
    every peer's exchange buffer over NVLink (ls_attn_reduce_push) and spin on per-source flags before merging
    (ls_attn_combine_wait) - no NCCL launch per attention call.  The buffers are torch symmetric memory (one allocation
    per rank, mapped into every process of the group by `rendezvous`); see include/longspec_b200.h (ls_xchg) for the
    layout and the epoch / parity protocol.  Raises if symmetric memory is unavailable: the caller then keeps the NCCL
    all-gather (LONGSPEC_XCHG=nccl forces that)."""

    def __init__(self, rows_max: int, device, group):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > _lib.LS_MAX_PEERS:
            raise RuntimeError(f"peer exchange supports up to {_lib.LS_MAX_PEERS} ranks")
        self.block_o = rows_max * 128
        self.block_lse = (rows_max + 31) // 32 * 32
        n_o, n_l, n_f = 2 * world * self.block_o, 2 * world * self.block_lse, 64
        self.buf = symm_mem.empty(n_o + n_l + n_f, dtype=torch.float32, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm_mem.rendezvous(self.buf, group.group_name if hasattr(group, "group_name") else group)
        dist.barrier(group)                                    # every rank's flags are zero before anybody pushes
        self.local_sync = torch.zeros(4, dtype=torch.int32, device=device)
        x = _lib.Xchg()
        ptrs = list(self.handle.buffer_ptrs)
        for r in range(world):
            x.peer_o[r] = ptrs[r]
            x.peer_lse[r] = ptrs[r] + 4 * n_o
            x.peer_flags[r] = ptrs[r] + 4 * (n_o + n_l)
        x.local_sync = self.local_sync.data_ptr()
        x.block_o, x.block_lse, x.rank, x.world = self.block_o, self.block_lse, rank, world
        self.x = x
        self.ref = C.byref(x)


_PEER_CACHE = {}


def get_peer_exchange(rows_max: int, device, group) -> PeerExchange:
    """One exchange buffer per (group, device), grown on demand and reused by every generation of the process: the
    allocation and the rendezvous are collectives, and the epoch simply keeps counting (all ranks issue the same
    sequence of exchange calls)."""
    key = (id(group), str(device))
    cur = _PEER_CACHE.get(key)
    if cur is None or cur.block_o < rows_max * 128:
        cur = PeerExchange(rows_max, device, group)
        _PEER_CACHE[key] = cur
    return cur

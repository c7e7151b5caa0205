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

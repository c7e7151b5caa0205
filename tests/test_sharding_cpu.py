"""world_size-2 gloo tests (CPU) of the N > 1 path: shard partition, the partial-slot exchange layout used by
Engine._exchange_partials, and the fact that merging per-shard partials reproduces the unsharded attention.
The per-shard attention here is the oracle (the CUDA kernels need a GPU); the exchange code is the product's."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import attention as OA
from oracle import flash_ref


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from longspec_b200.engine import Shard
        from longspec_b200.parallel import exchange_partials
        torch.manual_seed(0)                      # identical inputs on every rank (replicated compute)
        H, Hk, P, qv, n_splits = 4, 2, 1000, 7, 3
        G = H // Hk
        q = torch.randn(qv, H, 128)
        K = torch.randn(P + qv, Hk, 128)
        V = torch.randn(P + qv, Hk, 128)
        tail_mask = torch.tril(torch.ones(qv, qv, dtype=torch.bool))
        sh = Shard.make(P, rank, world)
        lo, hi = sh.begin, (P if sh.end < 0 else sh.end)
        # this rank's partial slots: its prefix chunk cut into n_splits pieces (+ the tail on the owner)
        rows = qv * G
        n_l = n_splits * Hk * rows
        part_o = torch.zeros(world * n_l * 128)
        part_l = torch.full((world * n_l,), float("nan"))
        po = part_o.view(world * n_splits, Hk, rows, 128)
        pl = part_l.view(world * n_splits, Hk, rows)
        bounds = torch.linspace(lo, hi, n_splits + 1).long().tolist()
        for s in range(n_splits):
            a, b = bounds[s], bounds[s + 1]
            keys_k, keys_v = K[a:b], V[a:b]
            mask = torch.ones(qv, b - a, dtype=torch.bool)
            if s == n_splits - 1 and sh.owner:
                keys_k, keys_v = torch.cat([keys_k, K[P:]]), torch.cat([keys_v, V[P:]])
                mask = torch.cat([mask, tail_mask], dim=1)
            o, lse = flash_ref.attend(q, keys_k, keys_v, mask)             # o (qv,H,128), lse (H,qv)
            slot = rank * n_splits + s
            po[slot] = o.view(qv, Hk, G, 128).permute(1, 0, 2, 3).reshape(Hk, rows, 128)
            pl[slot] = lse.view(Hk, G, qv).permute(0, 2, 1).reshape(Hk, rows)
        exchange_partials(part_o, part_l, n_l * 128, n_l, rank, world)
        assert not torch.isnan(part_l).any(), "every slot must be filled after the exchange"
        merged = OA.merge_partials(po.reshape(world * n_splits, Hk * rows, 128), pl.reshape(world * n_splits, Hk * rows))
        merged = merged.view(Hk, qv, G, 128).permute(1, 0, 2, 3).reshape(qv, H, 128)
        full_mask = torch.cat([torch.ones(qv, P, dtype=torch.bool), tail_mask], dim=1)
        ref, _ = flash_ref.attend(q, K, V, full_mask)
        torch.testing.assert_close(merged, ref, atol=1e-5, rtol=1e-5)
        # every rank must hold bit-identical merged results (identical accept decisions)
        gathered = [torch.empty_like(merged) for _ in range(world)]
        dist.all_gather(gathered, merged)
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_partials_exchange_and_merge_gloo(world):
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert all(ret.get(r) for r in range(world))


def test_split_plan_is_identical_on_every_rank():
    """The partial-slot plan must not depend on the rank's own shard size (unequal shards once hung the exchange)."""
    from longspec_b200.engine import Shard, split_plan
    for world in (1, 2, 4, 8):
        for P in (700, 3000, 16384, 131072):
            total = P + 96 + 256 + 74 + 128
            plan = split_plan(32, 8, 148, P, total, world, (74, 16, 1))
            assert all(n >= 1 for n in plan) and plan[0] % 2 == 0       # three row tiles -> trios -> even slot count
            sizes = [Shard.make(P, r, world) for r in range(world)]
            assert sizes[-1].end == -1 and sizes[0].begin == 0
    # world 2, 3000-token prompt: rank 0 holds 1536 tokens, rank 1 the rest + everything generated
    a, b = Shard.make(3000, 0, 2), Shard.make(3000, 1, 2)
    assert (a.begin, a.end, b.begin, b.end) == (0, 1536, 1536, -1)

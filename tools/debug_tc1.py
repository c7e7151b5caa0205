"""Where does the one-row-tile kernel differ from the general path?  (LONGSPEC_TC1=1|2 python tools/debug_tc1.py)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import paged_from_dense, run_attention  # noqa: E402


def gen(shape, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(dtype)


dev = torch.device("cuda:0")
dt = torch.bfloat16
H, Hk = 8, 2
G = H // Hk
for name, n0, K, V in (("V=1", 900, gen((1000, Hk, 128), 2, dt), torch.ones(1000, Hk, 128, dtype=dt)),
                       ("K=0", 1024, torch.zeros(1100, Hk, 128, dtype=dt), gen((1100, Hk, 128), 3, dt)),
                       ("random", 3000, gen((3009, Hk, 128), 2, dt), gen((3009, Hk, 128), 3, dt))):
    q = gen((74, H, 128), 1, dt)
    kv = paged_from_dense(K, V, dt, dev)
    tc, _, _ = run_attention(q.to(dev), kv, n0, 74, n_splits=4, entry="ls_attn_partial")
    mma, _, _ = run_attention(q.to(dev), kv, n0, 74, n_splits=4, entry="ls_attn_partial_mma")
    tc, mma = tc.float().cpu(), mma.float().cpu()
    bad = ~((tc - mma).abs() <= 2e-2)                     # NaN counts as bad
    print(f"== {name}: {int(bad.sum())} bad of {bad.numel()}, NaN {int(torch.isnan(tc).sum())}")
    # packed row r = token * G + g of kv head kvh = head // G
    for kvh in range(Hk):
        b = bad[:, kvh * G:(kvh + 1) * G, :].reshape(74 * G, 128)
        for rt in range(3):
            rows = b[rt * 128:min((rt + 1) * 128, 74 * G)]
            print(f"   kv head {kvh} row tile {rt}: bad rows {int(rows.any(1).sum())}/{rows.shape[0]}, "
                  f"bad by 32-col block {[int(rows[:, c:c + 32].sum()) for c in range(0, 128, 32)]}, "
                  f"first bad rows {rows.any(1).nonzero().flatten()[:6].tolist()}")

"""Run the target tree-verify attention (one layer, Llama-3-8B GQA-4 shape) a few times; used under ncu.
  python tools/prof_attn.py [N=131072] [iters=3] [entry=ls_attn_partial]"""
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longspec_b200 import _lib  # noqa: E402
from longspec_b200._lib import AttnDesc  # noqa: E402
from longspec_b200.engine import PagedKV, _ptr, _stream  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
entry = sys.argv[3] if len(sys.argv) > 3 else "ls_attn_partial"
H, Hk, qv = int(os.environ.get("H", 32)), int(os.environ.get("HK", 8)), int(os.environ.get("QV", 74))
dev = torch.device("cuda:0")
dt = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
layers = 3
kv = PagedKV(layers, Hk, N + 512, dt, dev)
for li in range(layers):
    kv.k[li].copy_(torch.randn(kv.k[li].shape, generator=g, device=dev, dtype=torch.float32).to(dt))
    kv.v[li].copy_(torch.randn(kv.v[li].shape, generator=g, device=dev, dtype=torch.float32).to(dt))
q = torch.randn(qv, H, 128, generator=g, device=dev, dtype=torch.float32).to(dt)
out = torch.empty_like(q)
G = H // Hk
sms = _lib.load().ls_sm_count()
n_splits = int(os.environ.get("SPLITS", sms // Hk if qv * G <= 256 else 2 * (sms // (3 * Hk))))
part_o = torch.empty(n_splits * Hk * qv * G * 128, dtype=torch.float32, device=dev)
part_l = torch.empty(n_splits * Hk * qv * G, dtype=torch.float32, device=dev)
kvl = torch.tensor([N], dtype=torch.int32, device=dev)
nq = torch.tensor([qv - 2], dtype=torch.int32, device=dev)
bits = torch.zeros(qv * 4, dtype=torch.int64, device=dev)
bits[0::4] = 1
descs = []
for li in range(layers):
    a = AttnDesc()
    a.q, a.out = _ptr(q), _ptr(out)
    a.k_pool, a.v_pool, a.page_table = _ptr(kv.k[li]), _ptr(kv.v[li]), _ptr(kv.page_table)
    a.kv_len_ptr, a.n_q_ptr, a.tail_mask = _ptr(kvl), _ptr(nq), _ptr(bits)
    a.part_o, a.part_lse = _ptr(part_o), _ptr(part_l)
    a.q_rows, a.n_heads, a.n_kv_heads = qv, H, Hk
    a.n_tail, a.window, a.causal, a.dtype = qv, -1, 0, _lib.LS_BF16
    a.n_splits, a.n_pages = n_splits, kv.n_pages
    a.part_offset, a.n_parts_total = 0, n_splits
    a.kv_shard_begin, a.kv_shard_end, a.tail_owner = 0, -1, 1
    a.scale = 1.0 / math.sqrt(128)
    descs.append(a)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for a in descs:
    _lib.call(entry, C.byref(a), _stream())
torch.cuda.synchronize()
ev0.record()
for it in range(iters):
    a = descs[it % layers]
    _lib.call(entry, C.byref(a), _stream())
    _lib.call("ls_attn_combine", C.byref(a), _stream())
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / iters
# partial kernel alone / combine alone
for name in (entry, "ls_attn_combine"):
    torch.cuda.synchronize()
    ev0.record()
    for it in range(iters):
        _lib.call(name, C.byref(descs[it % layers]), _stream())
    ev1.record()
    torch.cuda.synchronize()
    print(f"   {name} alone: {ev0.elapsed_time(ev1) / iters * 1e3:.1f} us")
byt = N * 2 * Hk * 128 * 2
print(f"N={N} splits={n_splits} {entry}+combine: {ms * 1e3:.1f} us/iter, {byt / ms / 1e6:.0f} GB/s, "
      f"{4.0 * (qv - 2) * N * 128 * H / ms / 1e9:.0f} TFLOP/s")

"""GPU baseline for the verify attention: what the REFERENCE runs on this box, timed next to our kernels.

The reference's target tree-verify attention (longspec/test/llama.py:385-421) is
    prefix_o, prefix_lse = flash_attn_with_kvcache(q, K_Cache, V_Cache, cache_seqlens=cache_lens, return_softmax_lse=True)
    current_out, weight  = tree_part_fwd(q, k, v, tree_mask, ...)     # eager/inductor 74 x 74 masked attention
    attn_output          = prefix_o * weight + current_out * (1 - weight)
This tool times exactly those calls with the image's flash-attn wheel (2.8.3, sm_100 SASS of the FA2 kernels) on
synthetic tensors of the benchmark shape, and then `ls_attn_partial` + `ls_attn_combine` on the same shape.
It is measurement tooling: nothing under longspec_b200/ imports it, and it is the only file in the repo that
imports flash_attn.

    python tools/ref_gpu_attn.py [N=131072] [iters=10]        env: H=32 HK=8 QV=74 DTYPE=bf16|fp16
Prints one JSON line (also written to gpurun_out/ref_gpu_attn_<N>.json when that directory exists).
"""
import ctypes as C
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
H, Hk, qv = int(os.environ.get("H", 32)), int(os.environ.get("HK", 8)), int(os.environ.get("QV", 74))
dt = torch.float16 if os.environ.get("DTYPE", "bf16") == "fp16" else torch.bfloat16
dev = torch.device("cuda:0")
G = H // Hk
layers = 3                      # a different layer's cache every launch: nothing is served from L2
scale = 1.0 / math.sqrt(128)
g = torch.Generator(device="cuda").manual_seed(0)


def timed(fn, n=iters, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3          # us


res = {"shape": {"N": N, "q_rows": qv, "heads": H, "kv_heads": Hk, "dtype": str(dt)}, "gpu": torch.cuda.get_device_name(0)}
kv_bytes = N * 2 * Hk * 128 * 2

# ------------------------------------------------------------------ reference: flash-attn FA2 + eager tree part
try:
    import flash_attn
    from flash_attn import flash_attn_with_kvcache
    S = N + 256
    Kc = [torch.randn(1, S, Hk, 128, generator=g, device=dev, dtype=torch.float32).to(dt) for _ in range(layers)]
    Vc = [torch.randn(1, S, Hk, 128, generator=g, device=dev, dtype=torch.float32).to(dt) for _ in range(layers)]
    q = torch.randn(1, qv, H, 128, generator=g, device=dev, dtype=torch.float32).to(dt)
    k_new = torch.randn(1, qv, Hk, 128, generator=g, device=dev, dtype=torch.float32).to(dt)
    v_new = torch.randn(1, qv, Hk, 128, generator=g, device=dev, dtype=torch.float32).to(dt)
    lens = torch.tensor([N], dtype=torch.int32, device=dev)
    tree_mask = torch.tril(torch.ones(1, qv, qv, dtype=torch.int64, device=dev))
    rng = torch.arange(1024, device=dev)

    def prefix(i):
        return flash_attn_with_kvcache(q, Kc[i % layers], Vc[i % layers], cache_seqlens=lens, return_softmax_lse=True)

    def tree_part(i, prefix_lse):
        # llama.py:394-421, eager (the reference wraps this in torch.compile; eager is its fallback and upper bound)
        Kl, Vl = Kc[i % layers], Vc[i % layers]
        idx = lens.unsqueeze(-1) + rng[:qv].unsqueeze(0)
        b = rng[:1].unsqueeze(-1)
        Kl[b, idx] = k_new
        Vl[b, idx] = v_new
        ks = k_new.repeat_interleave(G, dim=2).permute(0, 2, 3, 1)
        vs = v_new.repeat_interleave(G, dim=2).transpose(1, 2)
        qs = q.transpose(1, 2)
        score = (torch.matmul(qs, ks) * scale).to(torch.float32)
        score = score.masked_fill(tree_mask.unsqueeze(1).expand(-1, H, -1, -1) == 0, -float("inf"))
        w = torch.softmax(score, dim=-1).to(dt)
        cur = torch.matmul(w, vs).permute(0, 2, 1, 3)
        cur_lse = score.logsumexp(dim=-1, keepdim=True).transpose(1, 2)
        pl = prefix_lse.view(1, H, qv, -1).transpose(1, 2)
        weight = torch.sigmoid(pl - cur_lse).to(dt)
        return cur, weight

    def full(i):
        po, pl = prefix(i)
        cur, w = tree_part(i, pl)
        return po * w + cur * (1 - w)

    t_prefix = timed(prefix)
    t_full = timed(full)
    res["reference"] = {
        "flash_attn_version": flash_attn.__version__,
        "flash_attn_with_kvcache_us": t_prefix, "prefix_plus_tree_part_plus_merge_us": t_full,
        "prefix_gbs": kv_bytes / t_prefix / 1e3, "full_gbs": kv_bytes / t_full / 1e3,
        "note": "llama.py:385-387,394-421 on synthetic tensors; tree part eager (reference: torch.compile)"}
    del Kc, Vc
    torch.cuda.empty_cache()
except Exception as e:  # noqa: BLE001
    res["reference"] = {"unavailable": f"{type(e).__name__}: {e}"}

# ------------------------------------------------------------------ ours, same shape, through the C ABI
from longspec_b200 import _lib  # noqa: E402
from longspec_b200._lib import AttnDesc  # noqa: E402
from longspec_b200.engine import PagedKV, _ptr, _stream, split_plan  # noqa: E402

kv = PagedKV(layers, Hk, N + 512, dt, dev)
for li in range(layers):
    kv.k[li].copy_(torch.randn(kv.k[li].shape, generator=g, device=dev, dtype=torch.float32).to(dt))
    kv.v[li].copy_(torch.randn(kv.v[li].shape, generator=g, device=dev, dtype=torch.float32).to(dt))
q2 = torch.randn(qv, H, 128, generator=g, device=dev, dtype=torch.float32).to(dt)
out = torch.empty_like(q2)
sms = _lib.load().ls_sm_count()
n_splits = int(os.environ.get("SPLITS", 0)) or split_plan(H, Hk, sms, N, N + 512, 1, (qv,))[0]
part_o = torch.empty(n_splits * Hk * qv * G * 128, dtype=torch.float32, device=dev)
part_l = torch.empty(n_splits * Hk * qv * G, dtype=torch.float32, device=dev)
kvl = torch.tensor([N], dtype=torch.int32, device=dev)
nq = torch.tensor([qv], dtype=torch.int32, device=dev)
bits = torch.zeros(qv, 4, dtype=torch.int64, device=dev)
for r in range(qv):                                   # causal tail mask (same visibility as the tril above)
    for w_ in range(4):
        lo = w_ * 64
        n = min(max(r + 1 - lo, 0), 64)
        v = (1 << n) - 1
        bits[r, w_] = v - (1 << 64) if v >= 1 << 63 else v
bits = bits.reshape(-1)
descs = []
for li in range(layers):
    a = AttnDesc()
    a.q, a.out = _ptr(q2), _ptr(out)
    a.k_pool, a.v_pool, a.page_table = _ptr(kv.k[li]), _ptr(kv.v[li]), _ptr(kv.page_table)
    a.kv_len_ptr, a.n_q_ptr, a.tail_mask = _ptr(kvl), _ptr(nq), _ptr(bits)
    a.part_o, a.part_lse = _ptr(part_o), _ptr(part_l)
    a.q_rows, a.n_heads, a.n_kv_heads = qv, H, Hk
    a.n_tail, a.window, a.causal, a.dtype = qv, -1, 0, (_lib.LS_F16 if dt == torch.float16 else _lib.LS_BF16)
    a.n_splits, a.n_pages = n_splits, kv.n_pages
    a.part_offset, a.n_parts_total = 0, n_splits
    a.kv_shard_begin, a.kv_shard_end, a.tail_owner = 0, -1, 1
    a.scale = scale
    descs.append(a)


def ours_partial(i):
    _lib.call("ls_attn_partial", C.byref(descs[i % layers]), _stream())


def ours_full(i):
    _lib.call("ls_attn_partial", C.byref(descs[i % layers]), _stream())
    _lib.call("ls_attn_combine", C.byref(descs[i % layers]), _stream())


t_p = timed(ours_partial)
t_f = timed(ours_full)
res["ours"] = {"lib": _lib.LIB_PATH, "n_splits": n_splits, "ls_attn_partial_us": t_p, "partial_plus_combine_us": t_f,
               "partial_gbs": kv_bytes / t_p / 1e3, "full_gbs": kv_bytes / t_f / 1e3,
               "note": "prefix + append-free masked tail + merge in one launch, then the split-KV combine"}
if "unavailable" not in res["reference"]:
    res["speedup_vs_reference_full"] = res["reference"]["prefix_plus_tree_part_plus_merge_us"] / t_f
    res["speedup_vs_flash_attn_prefix_only"] = res["reference"]["flash_attn_with_kvcache_us"] / t_f
line = json.dumps(res)
print(line)
od = os.path.join(ROOT, "gpurun_out")
if os.path.isdir(od):
    with open(os.path.join(od, f"ref_gpu_attn_{N}.json"), "w") as f:
        f.write(line + "\n")

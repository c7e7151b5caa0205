"""-m gpu: RoPE+append, RMSNorm, SiLU*up, gathers vs plain PyTorch fp32 references (floating-point
kernels: tolerance = one rounding of the kernel dtype)."""
import ctypes as C

import pytest
import torch

from oracle import model as OM

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_rope_append_matches_oracle_and_writes_pages(dtype):
    from longspec_b200 import _lib
    from longspec_b200.engine import PagedKV, _ptr, _stream, _code
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    rows, H, Hk = 9, 8, 2
    qkv = torch.randn(rows, (H + 2 * Hk) * 128, generator=g).to(dtype).to(dev)
    pos = torch.tensor([131000, 5, 17, 17, 4096, 99999, 0, 1, 2], dtype=torch.int32, device=dev)
    dims = OM.Dims(hidden=H * 128, heads=H, kv_heads=Hk, layers=1, ffn=8, vocab=8, rope_theta=283461213.0)
    cos, sin = OM.rope_cos_sin(dims, pos.cpu().long()[None], dtype)
    q0 = qkv[:, :H * 128].cpu().view(1, rows, H, 128)
    k0 = qkv[:, H * 128:(H + Hk) * 128].cpu().view(1, rows, Hk, 128)
    v0 = qkv[:, (H + Hk) * 128:].cpu().view(rows, Hk, 128)
    q_ref, k_ref = OM.apply_rope(q0, cos, sin)[0], OM.apply_rope(k0, cos, sin)[0]
    inv = (1.0 / (dims.rope_theta ** (torch.arange(0, 128, 2, dtype=torch.int64).float() / 128))).to(dev)
    kv = PagedKV(1, Hk, 1024, dtype, dev, shuffle_pages=True)
    base = torch.tensor([250], dtype=torch.int32, device=dev)
    nrows = torch.tensor([7], dtype=torch.int32, device=dev)
    q, k, v = qkv[:, :H * 128], qkv[:, H * 128:(H + Hk) * 128], qkv[:, (H + Hk) * 128:]
    _lib.call("ls_rope_append", _ptr(q), _ptr(k), _ptr(v), qkv.stride(0), qkv.stride(0), qkv.stride(0), _ptr(pos), _ptr(inv),
              _ptr(kv.k[0]), _ptr(kv.v[0]), _ptr(kv.page_table), _ptr(base), 3, 1, _ptr(nrows), rows, H, Hk, _code(dtype), _stream())
    torch.cuda.synchronize()
    tol = dict(atol=2e-2, rtol=2e-2) if dtype == torch.bfloat16 else dict(atol=3e-3, rtol=3e-3)
    torch.testing.assert_close(q.cpu().view(rows, H, 128).float(), q_ref.float(), **tol)
    torch.testing.assert_close(k.cpu().view(rows, Hk, 128).float(), k_ref.float(), **tol)
    Kr, Vr = kv.read(0, 250 + 3 + 1, 9)
    assert torch.equal(Kr[:7].cpu(), k.cpu().view(rows, Hk, 128)[:7]), "rows < n_rows are appended at base+bias+offset"
    assert torch.equal(Vr[:7].cpu(), v0[:7])
    assert torch.all(Kr[7:] == 0) and torch.all(Vr[7:] == 0), "rows >= n_rows are not written"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_add_rmsnorm_silu_gather(dtype):
    from longspec_b200 import _lib
    from longspec_b200.engine import _ptr, _stream, _code
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    rows, hid, ffn = 13, 4096, 1024
    h = torch.randn(rows, hid, generator=g).to(dtype).to(dev)
    dl = torch.randn(rows, hid, generator=g).to(dtype).to(dev)
    w = (1 + 0.1 * torch.randn(hid, generator=g)).to(dtype).to(dev)
    h_ref = (h + dl)
    x_ref = OM.rms_norm(h_ref.cpu(), w.cpu(), 1e-5)
    out = torch.empty_like(h)
    _lib.call("ls_add_rmsnorm", _ptr(h), _ptr(dl), _ptr(w), _ptr(out), rows, hid, C.c_float(1e-5), _code(dtype), _stream())
    torch.cuda.synchronize()
    assert torch.equal(h, h_ref), "residual add is rounded once to dtype"
    tol = dict(atol=2e-2, rtol=2e-2) if dtype == torch.bfloat16 else dict(atol=2e-3, rtol=2e-3)
    torch.testing.assert_close(out.cpu().float(), x_ref.float(), **tol)
    gu = torch.randn(rows, 2 * ffn, generator=g).to(dtype).to(dev)
    act = torch.empty(rows, ffn, dtype=dtype, device=dev)
    _lib.call("ls_silu_mul", _ptr(gu), _ptr(act), rows, ffn, _code(dtype), _stream())
    ref = torch.nn.functional.silu(gu[:, :ffn].float()) * gu[:, ffn:].float()
    torch.cuda.synchronize()
    torch.testing.assert_close(act.float(), ref, **tol)
    idx = torch.tensor([5, 0, 12, 5], dtype=torch.int32, device=dev)
    got = torch.empty(4, hid, dtype=dtype, device=dev)
    _lib.call("ls_gather_rows", _ptr(h), _ptr(idx), _ptr(got), 4, hid, _code(dtype), _stream())
    ids = torch.tensor([3, 1, 3], dtype=torch.int64, device=dev)
    emb = torch.empty(3, hid, dtype=dtype, device=dev)
    _lib.call("ls_embed", _ptr(h), _ptr(ids), _ptr(emb), 3, hid, _code(dtype), _stream())
    torch.cuda.synchronize()
    assert torch.equal(got, h[idx.long()]) and torch.equal(emb, h[ids])

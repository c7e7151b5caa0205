"""-m gpu: the tcgen05/TMA attention kernel (ls_attn_partial dispatch) vs the oracle and vs the general
mma.sync path, staged so that a failure localises the broken piece (plumbing / PV / QK / rescaling)."""
import pytest
import torch

from oracle import flash_ref

pytestmark = pytest.mark.gpu


def gen(shape, seed, dtype, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype)


def run_both(q, K, V, n0, n_q, dtype, splits=4, causal=0):
    from gpu_util import paged_from_dense, run_attention
    dev = torch.device("cuda:0")
    kv = paged_from_dense(K, V, dtype, dev)
    tc, _, _ = run_attention(q.to(dev), kv, n0, n_q, n_splits=splits, causal=causal, entry="ls_attn_partial")
    mma, _, _ = run_attention(q.to(dev), kv, n0, n_q, n_splits=splits, causal=causal, entry="ls_attn_partial_mma")
    return tc.float().cpu(), mma.float().cpu()


def test_stage1_constant_v_plumbing():
    """V == 1 everywhere: any correctly normalised attention returns exactly 1 (pipeline, barriers, epilogue)."""
    dtype = torch.bfloat16
    q = gen((74, 8, 128), 1, dtype)
    K = gen((1000, 2, 128), 2, dtype)
    V = torch.ones(1000, 2, 128, dtype=dtype)
    tc, mma = run_both(q, K, V, 900, 74, dtype)
    torch.testing.assert_close(tc, torch.ones_like(tc), atol=1e-2, rtol=0)


def test_stage2_zero_k_checks_pv():
    """K == 0: uniform weights, output = mean of V (PV MMA, MN-major V descriptor, TMA swizzle)."""
    dtype = torch.bfloat16
    q = gen((74, 8, 128), 1, dtype)
    K = torch.zeros(1100, 2, 128, dtype=dtype)
    V = gen((1100, 2, 128), 3, dtype)
    n0 = 1024
    tc, mma = run_both(q, K, V, n0, 74, dtype)
    exp = V[:n0].float().mean(0).repeat_interleave(4, dim=0)[None].expand(74, 8, 128)
    torch.testing.assert_close(mma, exp, atol=5e-3, rtol=0)
    torch.testing.assert_close(tc, exp, atol=5e-3, rtol=0)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("H,Hk,n0,n_q,splits", [(8, 2, 3000, 74, 5), (32, 8, 16384, 74, 19), (8, 8, 700, 74, 3),
                                               (10, 2, 5000, 69, 7), (8, 2, 2048, 16, 4), (8, 2, 64, 3, 2), (8, 2, 63, 74, 3)])
def test_stage3_random_matches_oracle_and_mma(dtype, H, Hk, n0, n_q, splits):
    q = gen((n_q, H, 128), 1, dtype)
    K = gen((n0 + 9, Hk, 128), 2, dtype)
    V = gen((n0 + 9, Hk, 128), 3, dtype)
    tc, mma = run_both(q, K, V, n0, n_q, dtype, splits)
    ref, _ = flash_ref.attend(q.float(), K[:n0].float(), V[:n0].float())
    tol = dict(atol=2e-2, rtol=2e-2) if dtype == torch.bfloat16 else dict(atol=4e-3, rtol=4e-3)
    torch.testing.assert_close(mma, ref, **tol)
    torch.testing.assert_close(tc, ref, **tol)


def test_stage4_growing_maxima_exercise_lazy_rescale():
    """Scores that keep growing along the sequence force the O-in-TMEM rescale path many times."""
    dtype = torch.bfloat16
    n0 = 4096
    q = gen((74, 8, 128), 1, dtype)
    ramp = torch.linspace(0.2, 6.0, n0)[:, None, None]
    K = (gen((n0, 2, 128), 2, torch.float32) * ramp).to(dtype)
    V = gen((n0, 2, 128), 3, dtype)
    tc, mma = run_both(q, K, V, n0, 74, dtype, splits=2)
    ref, _ = flash_ref.attend(q.float(), K.float(), V.float())
    torch.testing.assert_close(mma, ref, atol=3e-2, rtol=3e-2)
    torch.testing.assert_close(tc, ref, atol=3e-2, rtol=3e-2)


def test_causal_rows_split_between_tc_and_remainder():
    """draft cross attention, step 0 (llama_glide.py:265): the causal rows live in the mma remainder."""
    dtype = torch.bfloat16
    n0, n_q = 1500, 6
    q = gen((n_q, 8, 128), 1, dtype)
    K = gen((n0, 2, 128), 2, dtype)
    V = gen((n0, 2, 128), 3, dtype)
    tc, mma = run_both(q, K, V, n0, n_q, dtype, splits=4, causal=1)
    mask = flash_ref.visible(n_q, n0, True)
    ref, _ = flash_ref.attend(q.float(), K.float(), V.float(), mask)
    torch.testing.assert_close(tc, ref, atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("splits", [1, 2, 12])
def test_three_row_tiles_even_and_single_slot_plans(splits):
    """296 packed rows = heavy/heavy/light CTA trios (even slot counts) or a heavy+light pair (one slot)."""
    dtype = torch.bfloat16
    n0 = 5000
    q = gen((74, 8, 128), 1, dtype)
    K = gen((n0 + 9, 2, 128), 2, dtype)
    V = gen((n0 + 9, 2, 128), 3, dtype)
    tc, mma = run_both(q, K, V, n0, 74, dtype, splits)
    ref, _ = flash_ref.attend(q.float(), K[:n0].float(), V[:n0].float())
    torch.testing.assert_close(tc, ref, atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("n_q", [1, 20, 40, 64, 65, 73])
def test_runtime_rows_below_the_static_plan(n_q):
    """The grid is planned from q_rows (74 -> three row tiles); the device-side row count may leave the light CTA,
    or the second row tile, without a single live row (veri_rows = acc_num + S < q_v, engine.py)."""
    from gpu_util import paged_from_dense, run_attention
    dtype = torch.bfloat16
    dev = torch.device("cuda:0")
    n0 = 3000
    q = gen((74, 8, 128), 1, dtype)
    K = gen((n0, 2, 128), 2, dtype)
    V = gen((n0, 2, 128), 3, dtype)
    kv = paged_from_dense(K, V, dtype, dev)
    out, _, _ = run_attention(q.to(dev), kv, n0, n_q, n_splits=6)
    ref, _ = flash_ref.attend(q[:n_q].float(), K.float(), V.float())
    torch.testing.assert_close(out[:n_q].float().cpu(), ref, atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("entry", ["ls_attn_partial", "ls_attn_partial_mma"])
def test_q_read_in_place_from_a_fused_qkv_row(entry):
    """q_stride: the q slice of a [rows][H*128 + 2*Hk*128] QKV GEMM output is consumed without a copy."""
    from gpu_util import paged_from_dense, run_attention
    dtype = torch.bfloat16
    dev = torch.device("cuda:0")
    n0, n_q, H, Hk = 2000, 74, 8, 2
    qkv = gen((n_q, (H + 2 * Hk) * 128), 1, dtype).to(dev)
    q = qkv[:, :H * 128].view(n_q, H, 128)
    assert not q.is_contiguous()
    K = gen((n0, Hk, 128), 2, dtype)
    V = gen((n0, Hk, 128), 3, dtype)
    kv = paged_from_dense(K, V, dtype, dev)
    out, _, _ = run_attention(q, kv, n0, n_q, n_splits=4, entry=entry)
    ref, _ = flash_ref.attend(q.float().cpu(), K.float(), V.float())
    torch.testing.assert_close(out.float().cpu(), ref, atol=2e-2, rtol=2e-2)


def test_more_slots_than_tiles_leaves_empty_ctas_harmless():
    """Two row tiles and far more slots than key tiles: CTAs without any tile must publish empty partials
    (zeros and -inf), never stale TMEM bits."""
    dtype = torch.bfloat16
    n0 = 700
    q = gen((40, 8, 128), 1, dtype)
    K = gen((n0 + 9, 2, 128), 2, dtype)
    V = gen((n0 + 9, 2, 128), 3, dtype)
    tc, mma = run_both(q, K, V, n0, 40, dtype, splits=16)
    ref, _ = flash_ref.attend(q.float(), K[:n0].float(), V[:n0].float())
    assert torch.isfinite(tc).all()
    torch.testing.assert_close(tc, ref, atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("H,Hk,n0,n_q,splits,n_tail,causal", [
    (8, 2, 3000, 74, 12, 74, 0),      # three row tiles (trios) + tree-masked tail: the verify shape
    (8, 2, 3000, 69, 12, 74, 0),      # fewer live rows than planned
    (32, 8, 16384, 74, 12, 74, 0),    # the headline head geometry: 144 CTAs
    (8, 2, 2048, 16, 4, 0, 1),        # two-row-tile plan, causal (draft cross attention, step 0)
    (8, 2, 5000, 64, 9, 0, 1),        # a prefill block: 256 packed rows, causal
    (4, 4, 700, 74, 3, 74, 0),        # MHA, one row tile
    (8, 2, 40, 74, 4, 74, 0),         # more slots than tiles: empty CTAs still take part in the barrier
])
def test_fused_combine_equals_partial_plus_combine(dtype, H, Hk, n0, n_q, splits, n_tail, causal):
    """ls_attn_desc::fuse_sync: the attention kernel merges its split-KV slots in its own tail (kv-head counter barrier) -
    bit-identical to ls_attn_partial + ls_attn_combine (same kernel; three-row-tile calls take a different kernel when
    they do not fuse, hence a tolerance there), twice in a row (the counters must reset themselves)."""
    from gpu_util import paged_from_dense, run_attention
    dev = torch.device("cuda:0")
    q_rows = max(n_q, 74 if n_tail else n_q)
    q = gen((q_rows, H, 128), 1, dtype).to(dev)
    K = gen((n0 + q_rows + 9, Hk, 128), 2, dtype)
    V = gen((n0 + q_rows + 9, Hk, 128), 3, dtype)
    kv = paged_from_dense(K, V, dtype, dev)
    mask = torch.tril(torch.ones(q_rows, n_tail, dtype=torch.int64)) if n_tail else None
    kw = dict(n_tail=n_tail, tail_mask=mask, causal=causal, n_splits=splits)
    ref, _, _ = run_attention(q, kv, n0, n_q, **kw)
    three_row_tiles = q_rows * (H // Hk) > 256        # unfused: one row tile per CTA (attn_tc1_kernel); fused: the trio kernel
    for _ in range(2):
        got, _, _ = run_attention(q, kv, n0, n_q, fuse=True, **kw)
        if three_row_tiles:
            torch.testing.assert_close(got[:n_q].float(), ref[:n_q].float(), atol=2e-2, rtol=2e-2)
        else:
            assert torch.equal(got[:n_q], ref[:n_q])


@pytest.fixture
def tc1_mode():
    """Select the kernel for three-row-tile calls (ls_attn_tc1_mode) for one test, then restore it."""
    from longspec_b200 import _lib
    lib = _lib.load()
    prev = lib.ls_attn_tc1_mode(-1)

    def set_mode(m):
        lib.ls_attn_tc1_mode(m)
    yield set_mode
    lib.ls_attn_tc1_mode(prev)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("H,Hk,n0,n_q,q_rows,splits,causal", [
    (8, 2, 3000, 74, 74, 4, 0),        # two trios, special tiles in the second
    (32, 8, 16384, 74, 74, 12, 0),     # the headline geometry
    (8, 2, 5000, 40, 74, 6, 0),        # run-time rows end inside row tile 1: row tile 2 idle, padding warps in tile 1
    (8, 2, 5000, 20, 74, 5, 0),        # only row tile 0 live: in cluster mode its peer has no live rows but keeps loading
    (10, 2, 4000, 69, 69, 7, 0),       # GQA-5 (QwQ): 345 rows
    (8, 2, 777, 74, 74, 3, 1),         # causal rows in the special tiles, odd slot count (last trio has one slot)
    (8, 2, 100, 74, 74, 2, 0),         # no full tile at all
])
def test_three_row_tile_kernels_agree_with_oracle(tc1_mode, mode, H, Hk, n0, n_q, q_rows, splits, causal):
    """The three kernels that can serve a three-row-tile call (trios / one row tile per CTA / multicast clusters)
    against the oracle and the general path, incl. ragged run-time row counts and a growing-maximum K."""
    from gpu_util import paged_from_dense, run_attention
    tc1_mode(mode)
    dtype = torch.bfloat16
    dev = torch.device("cuda:0")
    q = gen((q_rows, H, 128), 1, dtype)
    K = gen((n0 + q_rows + 9, Hk, 128), 2, dtype)
    K[:n0] *= torch.linspace(0.2, 3.0, n0)[:, None, None].to(dtype)       # row maxima keep growing: lazy rescales
    V = gen((n0 + q_rows + 9, Hk, 128), 3, dtype)
    kv = paged_from_dense(K, V, dtype, dev)
    tc, _, _ = run_attention(q.to(dev), kv, n0, n_q, n_splits=splits, causal=causal, entry="ls_attn_partial")
    mma, _, _ = run_attention(q.to(dev), kv, n0, n_q, n_splits=splits, causal=causal, entry="ls_attn_partial_mma")
    tc, mma = tc.float().cpu()[:n_q], mma.float().cpu()[:n_q]
    torch.testing.assert_close(tc, mma, atol=2e-2, rtol=2e-2)
    if not causal:
        ref, _ = flash_ref.attend(q[:n_q].float(), K[:n0].float(), V[:n0].float())
        torch.testing.assert_close(tc, ref, atol=2e-2, rtol=2e-2)

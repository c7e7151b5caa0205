"""-m gpu parity of the reference-shaped operator seams (longspec_b200/seams.py) against fixtures produced by the
reference's own code: the Triton `_fwd_kernel` run under TRITON_INTERPRET=1 (triton_tree_attn.npz) and
`LlamaGlide.tree_verification` (tree_verification.npz).  tests/golden/gen_golden.py wrote both."""
import os

import numpy as np
import pytest
import torch

from oracle import attention as OA

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 4e-3), (torch.bfloat16, 2e-2)])
def test_attention_seam_matches_reference_triton_kernel(golden_dir, dtype, tol):
    """seams.attention(q, k, v, tree_mask) -> (o, L) vs triton_tree_attn.attention (triton_tree_attn.py:19-77)."""
    from longspec_b200.seams import attention
    z = np.load(os.path.join(golden_dir, "triton_tree_attn.npz"))
    for i in range(int(z["n"])):
        g = lambda k: torch.from_numpy(z[f"c{i}_{k}"])
        q, k, v, mask = g("q").to(dtype).cuda(), g("k").to(dtype).cuda(), g("v").to(dtype).cuda(), g("mask").cuda()
        o, L = attention(q, k, v, mask)
        assert o.shape == q.shape and o.dtype == dtype and L.shape == q.shape[:3] and L.dtype == torch.float32
        # vs the reference kernel's fp32 outputs (inputs rounded to the 16-bit type: tolerance of that type)
        torch.testing.assert_close(o.float().cpu(), g("o"), rtol=tol, atol=tol)
        torch.testing.assert_close(L.cpu(), g("L"), rtol=tol, atol=tol)
        # vs the oracle on the SAME rounded inputs (isolates the kernel from the input rounding)
        ro, rL = OA.tree_mask_attention(q.float().cpu(), k.float().cpu(), v.float().cpu(), g("mask")[None])
        torch.testing.assert_close(o.float().cpu(), ro, rtol=tol, atol=tol)
        torch.testing.assert_close(L.cpu(), rL, rtol=2e-3, atol=2e-3)
        # the 3-D mask form and return_log_normalizer=False
        o2 = attention(q, k, v, mask[None], return_log_normalizer=False)
        assert torch.equal(o2, o)


def test_tree_verification_seam_matches_reference_fixture(golden_dir):
    """seams.tree_verification(...) vs LlamaGlide.tree_verification outputs incl. the in-place KV compaction
    (llama_glide.py:1128-1175): bit-exact."""
    from longspec_b200.seams import tree_verification
    z = np.load(os.path.join(golden_dir, "tree_verification.npz"))
    for i in range(int(z["n"])):
        g = lambda k: z[f"c{i}_{k}"]
        tm = torch.from_numpy(g("tree_mask"))[None].cuda()
        f = tm.shape[1]
        cache_len = int(g("cache_len"))
        Hk, D, S = 2, 128, cache_len + f + 64
        gen = torch.Generator().manual_seed(i)
        Kc = torch.randn(1, S, Hk, D, generator=gen).to(torch.float16)
        Vc = torch.randn(1, S, Hk, D, generator=gen).to(torch.float16)
        Kc[0, cache_len:cache_len + f, :, :16] = torch.from_numpy(g("K_before").astype(np.float32)).to(torch.float16)
        Vc[0, cache_len:cache_len + f, :, :16] = torch.from_numpy(g("V_before").astype(np.float32)).to(torch.float16)
        K0, V0 = Kc.clone(), Vc.clone()
        Kc, Vc = Kc.cuda(), Vc.cuda()
        acc_ids, acc_num, double_input = tree_verification(
            torch.from_numpy(g("all_spec"))[None].cuda(), torch.from_numpy(g("llm_pred"))[None].cuda(), tm,
            torch.tensor([cache_len], dtype=torch.int32).cuda(), int(g("non_leaf_len")), Kc, Vc)
        acc = int(acc_num[0])
        assert acc == int(g("acc_num")[0]), f"case {i}"
        assert acc_ids.shape == (1, acc) and acc_ids.dtype == torch.int64
        assert np.array_equal(acc_ids[0].cpu().numpy(), g("acc_ids")[:acc]), f"case {i}"
        assert int(double_input[0]) == int(g("double_input")[0])
        Ka, Va = Kc.cpu(), Vc.cpu()
        assert np.array_equal(Ka[0, cache_len:cache_len + acc, :, :16].numpy(), g("K_after")[:acc]), f"case {i}"
        assert np.array_equal(Va[0, cache_len:cache_len + acc, :, :16].numpy(), g("V_after")[:acc]), f"case {i}"
        # nothing outside the compacted window moved
        assert torch.equal(Ka[0, :cache_len], K0[0, :cache_len]) and torch.equal(Ka[0, cache_len + f:], K0[0, cache_len + f:])
        assert torch.equal(Va[0, :cache_len], V0[0, :cache_len]) and torch.equal(Va[0, cache_len + f:], V0[0, cache_len + f:])
        # the method on the drop-in class is the same seam
        a2, n2, d2 = tree_verification(torch.from_numpy(g("all_spec"))[None].cuda(), torch.from_numpy(g("llm_pred"))[None].cuda(),
                                       tm, torch.tensor([cache_len], dtype=torch.int32).cuda(), int(g("non_leaf_len")))
        assert torch.equal(a2, acc_ids) and int(n2[0]) == acc

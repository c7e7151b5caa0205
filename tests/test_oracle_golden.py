"""Pin the oracle (oracle/) against fixtures produced by the reference's own code.

Fixtures: tests/golden/*.npz, written by tests/golden/gen_golden.py which imports the
unmodified reference through tests/golden/ref_shim.py.  CPU only.
"""
import os

import numpy as np
import pytest
import torch

from oracle import attention as OA
from oracle import tree as OT
from oracle.generate import tree_spec_generate, vanilla_generate
from oracle.model import Dims, OracleModel
from longspec_b200.synthetic import PRESETS, synthetic_state_dict


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_tree_verification_matches_reference(golden_dir):
    z = load(golden_dir, "tree_verification.npz")
    n = int(z["n"])
    assert n >= 12
    deep = 0
    for i in range(n):
        g = lambda k: z[f"c{i}_{k}"]
        acc_ids, acc_num, dbl, idx = OT.tree_verification(g("all_spec"), g("llm_pred"), g("tree_mask"),
                                                          int(g("non_leaf_len")))
        assert acc_num == int(g("acc_num")[0])
        assert np.array_equal(acc_ids, g("acc_ids")[:acc_num])
        assert dbl == int(g("double_input")[0])
        deep += acc_num > 2
        f = g("tree_mask").shape[0]
        for nm in ("K", "V"):
            win = torch.from_numpy(g(nm + "_before").astype(np.float32))
            OT.compact_rows(win, 0, f, idx)
            assert np.array_equal(win.numpy().astype(np.float16), g(nm + "_after"))
    assert deep >= 3, "fixture should exercise multi-token acceptance"


def test_target_tree_part_matches_reference(golden_dir):
    z = load(golden_dir, "target_tree_part.npz")
    for i in range(int(z["n"])):
        g = lambda k: torch.from_numpy(z[f"c{i}_{k}"])
        q, k, v = g("q"), g("k"), g("v")
        n0 = int(z[f"c{i}_n0"])
        K = torch.zeros_like(g("K_after"))
        V = torch.zeros_like(g("V_after"))
        cur, w = OA.target_tree_part(q, k, v, K, V, torch.tensor([n0]), g("mask")[None], g("prefix_lse"),
                                     last_layer=bool(z[f"c{i}_last"]))
        torch.testing.assert_close(cur, g("cur"), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(w, g("w"), rtol=1e-5, atol=1e-6)
        assert torch.equal(K, g("K_after")) and torch.equal(V, g("V_after"))


def test_draft_tree_part_matches_reference(golden_dir):
    """oracle draft tree part (Triton-semantics) vs the reference's non-Triton twin."""
    z = load(golden_dir, "draft_tree_part.npz")
    for i in range(int(z["n"])):
        g = lambda k: torch.from_numpy(z[f"c{i}_{k}"])
        q, k, v, mask = g("q"), g("k"), g("v"), g("mask")
        n0 = int(z[f"c{i}_n0"])
        K, V = g("K_before").clone(), g("V_before").clone()
        M, N = mask.shape
        K[0, n0 + N - M:n0 + N] = k[0]
        V[0, n0 + N - M:n0 + N] = v[0]
        kt = K[0, n0:n0 + N].permute(1, 0, 2)[None]
        vt = V[0, n0:n0 + N].permute(1, 0, 2)[None]
        o, L = OA.tree_mask_attention(q.permute(0, 2, 1, 3), kt, vt, mask[None])
        w = torch.sigmoid(g("prefix_lse") - L).transpose(1, 2).unsqueeze(-1)
        torch.testing.assert_close(o.transpose(1, 2), g("cur"), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(w, g("w"), rtol=1e-5, atol=1e-6)
        assert torch.equal(K, g("K_after")) and torch.equal(V, g("V_after"))


def test_tree_mask_attention_matches_triton_kernel(golden_dir):
    """oracle vs the reference Triton kernel run under TRITON_INTERPRET=1."""
    z = load(golden_dir, "triton_tree_attn.npz")
    for i in range(int(z["n"])):
        g = lambda k: torch.from_numpy(z[f"c{i}_{k}"])
        o, L = OA.tree_mask_attention(g("q"), g("k"), g("v"), g("mask")[None])
        torch.testing.assert_close(o, g("o"), rtol=2e-5, atol=2e-5)
        torch.testing.assert_close(L, g("L"), rtol=2e-5, atol=2e-5)


E2E = ["e2e_llama_tree44.npz", "e2e_llama_default_tree.npz", "e2e_llama_chain.npz",
       "e2e_llama_noisy_draft.npz", "e2e_qwen2_tree.npz"]


def oracle_from_fixture(z):
    dims = PRESETS[str(z["preset"])]
    ds = float(z["draft_scale"]) or None
    W = synthetic_state_dict(dims, seed=int(z["seed"]), draft_scale=ds)
    chk = sum(float(v.double().abs().sum()) for v in W.values())
    assert abs(chk - float(z["w_checksum"])) < 1e-6 * abs(chk), "synthetic weight RNG drifted"
    d = Dims(**{k: getattr(dims, k) for k in Dims.__dataclass_fields__})
    return OracleModel(d, W), dims


@pytest.mark.parametrize("name", E2E)
def test_e2e_tokens_and_rounds_match_reference(golden_dir, name):
    z = load(golden_dir, name)
    m, dims = oracle_from_fixture(z)
    ids = torch.from_numpy(z["prompt"])[None]
    P = ids.shape[1]
    trace = []
    out, count, num, _, _ = tree_spec_generate(
        m, ids, P, tree_shape=z["tree_shape"].tolist(), max_gen_len=int(z["max_gen_len"]),
        eos_id=int(z["eos_arg"]), config_eos=int(z["config_eos"]), qwen_variant=bool(z["qwen"]), trace=trace)
    assert np.array_equal(out, z["tree_out"])
    assert (count, num) == (int(z["count"]), int(z["num"]))
    assert len(trace) == int(z["n_rounds"])
    for i, r in enumerate(trace):
        assert np.array_equal(r["all_spec"], z[f"r{i}_all_spec"]), f"round {i} all_spec"
        assert np.array_equal(r["tree_mask"], z[f"r{i}_tree_mask"]), f"round {i} tree_mask"
        assert np.array_equal(r["llm_pred"], z[f"r{i}_llm_pred"]), f"round {i} llm_pred"
        assert np.array_equal(r["acc_ids"], z[f"r{i}_acc_ids"][:r["acc_num"]]), f"round {i} acc_ids"
        assert r["acc_num"] == int(z[f"r{i}_acc_num"])


@pytest.mark.parametrize("name", E2E[:2])
def test_vanilla_matches_reference_and_tree(golden_dir, name):
    z = load(golden_dir, name)
    m, dims = oracle_from_fixture(z)
    ids = torch.from_numpy(z["prompt"])[None]
    out, num, _ = vanilla_generate(m, ids, ids.shape[1], max_gen_len=int(z["max_gen_len"]),
                                   config_eos=int(z["config_eos"]))
    assert np.array_equal(out, z["vanilla_out"])
    n = int(z["count"]) + int(z["num"])
    assert np.array_equal(out[:n], z["tree_out"][:n]), "losslessness: tree(T=0) == vanilla"

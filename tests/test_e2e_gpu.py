"""-m gpu end-to-end parity: LlamaGlide / Qwen2Glide on the CUDA path vs the reference's own token
streams (tests/golden/e2e_*.npz, produced by the unmodified reference in fp32) and vs the lossless
invariant tree(T=0) == vanilla(T=0).  Token ids are compared bit-exactly."""
import os

import numpy as np
import pytest
import torch

from longspec_b200.synthetic import PRESETS, synthetic_state_dict

pytestmark = pytest.mark.gpu

E2E = ["e2e_llama_tree44.npz", "e2e_llama_default_tree.npz", "e2e_llama_chain.npz",
       "e2e_llama_noisy_draft.npz", "e2e_qwen2_tree.npz"]


def build(z, dtype, graph=True):
    from longspec_b200.llama_glide import LlamaGlide
    from longspec_b200.qwen2_glide import Qwen2Glide
    dims = PRESETS[str(z["preset"])]
    ds = float(z["draft_scale"]) or None
    W = synthetic_state_dict(dims, seed=int(z["seed"]), draft_scale=ds)        # CPU fp32: same tensors as the reference run
    cls = Qwen2Glide if int(z["qwen"]) else LlamaGlide
    m = cls(dims, W, None, dtype=dtype, shuffle_pages=True)
    m.use_graph = graph
    return m


@pytest.mark.parametrize("name", E2E)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_tree_tokens_match_reference(golden_dir, name, dtype):
    z = np.load(os.path.join(golden_dir, name))
    m = build(z, dtype)
    ids = torch.from_numpy(z["prompt"])[None]
    out, count, num, elapsed, spec_mask = m.tree_spec_generate(
        ids.cuda(), torch.tensor([ids.shape[1]]).cuda(), tree_shape=z["tree_shape"].tolist(),
        max_gen_len=int(z["max_gen_len"]), eos_id=int(z["eos_arg"]), temperature=0.0)
    got = out[0].cpu().numpy()
    ref = z["tree_out"]
    n_ref = int(z["count"]) + int(z["num"])
    first_bad = next((i for i in range(n_ref) if got[i] != ref[i]), None)
    assert first_bad is None, f"token {first_bad}: got {got[first_bad]} want {ref[first_bad]} ({name}, {dtype})"
    assert np.array_equal(got, ref), "untouched tail must keep the reference's fill value"
    if "noisy" in name:
        # a deliberately near-random draft: half-precision rounding may change which candidates it proposes
        # (acceptance counts) but never the emitted tokens (losslessness)
        assert abs(int(num) - int(z["num"])) <= 3
    else:
        assert (int(count), int(num)) == (int(z["count"]), int(z["num"])), "same rounds and accepted counts as the reference"
    assert out.shape == (1, int(z["max_gen_len"])) and out.dtype == torch.int64 and out.is_cuda
    assert isinstance(num, int) and isinstance(elapsed, float) and torch.is_tensor(count)
    assert spec_mask.shape == out.shape and int(spec_mask.sum()) == 0


@pytest.mark.parametrize("name", E2E[:2])
def test_vanilla_matches_reference_and_tree_is_lossless(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    m = build(z, torch.float16)
    ids = torch.from_numpy(z["prompt"])[None].cuda()
    out, num, elapsed = m.vanilla_generate(ids, torch.tensor([ids.shape[1]]).cuda(), max_gen_len=int(z["max_gen_len"]))
    assert np.array_equal(out[0].cpu().numpy(), z["vanilla_out"])
    assert num == int(z["vanilla_num"])
    n = int(z["count"]) + int(z["num"])
    assert np.array_equal(out[0, :n].cpu().numpy(), z["tree_out"][:n])


def test_graph_and_eager_rounds_agree(golden_dir):
    z = np.load(os.path.join(golden_dir, E2E[1]))
    ids = torch.from_numpy(z["prompt"])[None].cuda()
    outs = []
    for graph in (False, True):
        m = build(z, torch.bfloat16, graph=graph)
        out, count, num, _, _ = m.tree_spec_generate(ids, torch.tensor([ids.shape[1]]).cuda(),
                                                     tree_shape=z["tree_shape"].tolist(), max_gen_len=int(z["max_gen_len"]),
                                                     eos_id=int(z["eos_arg"]))
        outs.append((out.cpu(), int(count), num))
        assert m.last_stats["graph"] == graph
    assert torch.equal(outs[0][0], outs[1][0]) and outs[0][1:] == outs[1][1:]


def test_eos_stops_like_the_reference():
    """llama_glide.py:1120: stop as soon as config.eos_token_id appears anywhere in output_ids."""
    from longspec_b200.llama_glide import LlamaGlide
    dims = PRESETS["tiny"]
    W = synthetic_state_dict(dims, seed=0)

    class Cfg:
        pass
    m = LlamaGlide(dims, W, None, dtype=torch.float16)
    ids = torch.randint(2, 300, (1, 40), generator=torch.Generator().manual_seed(0)).cuda()
    out, count, num, _, _ = m.tree_spec_generate(ids, torch.tensor([40]).cuda(), tree_shape=[4, 4], max_gen_len=32, eos_id=999)
    toks = out[0, :int(count) + num].tolist()
    eos = toks[5]
    cfg = Cfg()
    cfg.eos_token_id = eos
    m.config = cfg
    out2, count2, num2, _, _ = m.tree_spec_generate(ids, torch.tensor([40]).cuda(), tree_shape=[4, 4], max_gen_len=32, eos_id=999)
    n2 = int(count2) + num2
    assert n2 <= len(toks) and eos in out2[0, :n2].tolist()
    assert out2[0, :n2].tolist() == toks[:n2]
    first = toks.index(eos)
    assert n2 >= first + 1 and n2 <= first + 1 + 3, "stops at the end of the round that produced eos"


# ---------------------------------------------------------------------------------------------------------------
# CUDA path vs the oracle run on this box's CPU (no committed fixture: the oracle is pinned to the reference by
# tests/test_oracle_golden.py).  These cases cover what the fixtures cannot: prompts long enough for the draft-cache
# tail shortcut (engine.py: first_needed = P - 512 - 256 only engages for P > 768), the headline head geometry
# (8 heads / 2 kv heads: 296 packed verify rows = heavy/heavy/light CTA trios inside the loop) and the wide tree.
# ---------------------------------------------------------------------------------------------------------------
def _oracle_tokens(dims, W, ids, P, shape, gen_len, eos):
    from oracle.generate import tree_spec_generate as oracle_tree
    from oracle.model import Dims, OracleModel
    om = OracleModel(Dims(**{k: getattr(dims, k) for k in Dims.__dataclass_fields__}), W)
    ref, rc, rn, _, _ = oracle_tree(om, ids, P, tree_shape=shape, max_gen_len=gen_len, eos_id=eos)
    return ref, rc, rn


ORACLE_CASES = [
    # preset, prompt length, tree, max_gen_len, dtype, draft_scale
    ("tiny-gqa", 1500, [4, 16, 16, 16, 16], 40, torch.bfloat16, None),
    ("tiny-gqa", 4000, [4, 16, 16, 16, 16], 40, torch.float16, 2.0),
    ("tiny-g4", 1500, [4, 16, 16, 16, 16], 40, torch.bfloat16, None),
    ("tiny-g4", 900, [4, 16, 16, 16, 16], 48, torch.float16, 2.0),
    ("tiny-gqa", 1500, [8, 32, 32, 32, 32], 56, torch.bfloat16, None),
    ("tiny-mha", 1500, [8, 32, 32, 32, 32], 56, torch.float16, None),
    ("tiny-g4", 1300, [8, 32, 32, 32, 32], 56, torch.bfloat16, None),
]


@pytest.mark.parametrize("preset,P,shape,gen_len,dtype,ds", ORACLE_CASES)
def test_tree_tokens_match_oracle_long_prompt_and_headline_geometry(preset, P, shape, gen_len, dtype, ds):
    from longspec_b200.llama_glide import LlamaGlide
    from longspec_b200.synthetic import synthetic_prompt
    dims = PRESETS[preset]
    W = synthetic_state_dict(dims, seed=11, draft_scale=ds)
    ids = synthetic_prompt(dims.vocab, P, seed=11)
    eos = dims.vocab + 9
    m = LlamaGlide(dims, W, None, dtype=dtype, shuffle_pages=True)
    out, count, num, _, _ = m.tree_spec_generate(ids.cuda(), torch.tensor([P]).cuda(), tree_shape=shape,
                                                 max_gen_len=gen_len, eos_id=eos)
    if P > 768:
        assert m.eng.draft_first > 0, "the draft-cache tail shortcut must be engaged in this test"
    ref, rc, rn = _oracle_tokens(dims, W, ids, P, shape, gen_len, eos)
    got = out[0].cpu().numpy()
    n_ref = rc + rn
    assert np.array_equal(got[:n_ref], ref[:n_ref]), f"CUDA tree path != oracle ({preset}, P={P}, {shape}, {dtype})\n{got}\n{ref}"
    if ds is None:
        assert (int(count), num) == (rc, rn), "same rounds and accepted counts as the oracle"
        assert np.array_equal(got, ref)


def test_right_padded_input_ids_are_sliced_to_prompt_length():
    """The reference reads hidden_states[:, input_len - 1] (llama_glide.py:955): tokens behind prompt_length are padding."""
    from longspec_b200.llama_glide import LlamaGlide
    from longspec_b200.synthetic import synthetic_prompt
    dims = PRESETS["tiny"]
    W = synthetic_state_dict(dims, seed=2)
    ids = synthetic_prompt(dims.vocab, 300, seed=2)
    m = LlamaGlide(dims, W, None, dtype=torch.float16)
    kw = dict(tree_shape=[4, 4], max_gen_len=24, eos_id=dims.vocab + 3)
    a = m.tree_spec_generate(ids[:, :200].cuda(), torch.tensor([200]).cuda(), **kw)
    b = m.tree_spec_generate(ids.cuda(), torch.tensor([200]).cuda(), **kw)
    assert torch.equal(a[0], b[0]) and int(a[1]) == int(b[1]) and a[2] == b[2]


def test_sequence_sharded_generation_on_two_gpus_matches_single_gpu():
    """2-rank NCCL run (skipped on a 1-GPU box): prefix KV sharded over two GPUs, headline head geometry, prompt long
    enough for the draft tail shortcut; tools/check_multigpu.py asserts rank equality and equality with the 1-GPU run."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(root, "tools", "check_multigpu.py"), "tiny-g4", "3000"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "multi-GPU check ok" in res.stdout

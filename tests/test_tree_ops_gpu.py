"""-m gpu parity (bit-exact): device-side tree bookkeeping kernels vs the numpy oracle and vs the
golden fixtures produced by the reference's own tree_verification (tests/golden/tree_verification.npz)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import tree as OT

pytestmark = pytest.mark.gpu


def _setup(shape, P=100, max_gen=64):
    from longspec_b200.engine import alloc_tree_state
    dev = torch.device("cuda:0")
    st, ts = alloc_tree_state(dev, shape, P, max_gen)
    return dev, st, ts


def _bits_to_dense(bits, rows, cols):
    b = bits.cpu().numpy().astype(np.uint64).reshape(rows, -1)
    out = np.zeros((rows, cols), dtype=np.int64)
    for r in range(rows):
        for c in range(cols):
            out[r, c] = (int(b[r, c >> 6]) >> (c & 63)) & 1
    return out


def test_tree_accept_matches_reference_fixture(golden_dir):
    """ls_tree_accept vs LlamaGlide.tree_verification (llama_glide.py:1128-1175) outputs."""
    from longspec_b200 import _lib
    from longspec_b200.engine import PagedKV, _ptr, _stream
    from gpu_util import pack_bits
    z = np.load(os.path.join(golden_dir, "tree_verification.npz"))
    for i in range(int(z["n"])):
        g = lambda k: z[f"c{i}_{k}"]
        shape = g("shape").tolist()
        tm = g("tree_mask")
        f = tm.shape[0]
        cache_len = int(g("cache_len"))
        dev, st, ts = _setup(shape, P=cache_len - 3)
        st["all_spec"].copy_(torch.from_numpy(g("all_spec")))
        st["tree_bits"].copy_(pack_bits(torch.from_numpy(tm)))
        st["parent"].copy_(torch.from_numpy(OT.parent_index(tm).astype(np.int32)))
        st["scalars"][:4] = torch.tensor([cache_len, cache_len, cache_len, 1], dtype=torch.int32)
        pred = torch.from_numpy(g("llm_pred")).to(dev)
        # last-layer KV window as bf16 rows (values are exactly representable: fixture stored fp16 of randn)
        Hk, D = 2, 128
        Kw = torch.zeros(f, Hk, D)
        Vw = torch.zeros(f, Hk, D)
        Kw[:, :, :16] = torch.from_numpy(g("K_before").astype(np.float32))
        Vw[:, :, :16] = torch.from_numpy(g("V_before").astype(np.float32))
        kv = PagedKV(1, Hk, cache_len + f + 200, torch.float16, dev, shuffle_pages=True)
        kv.write(0, Kw.to(dev, torch.float16), Vw.to(dev, torch.float16), cache_len)
        _lib.call("ls_tree_accept", C.byref(ts), _ptr(pred), _ptr(kv.k[0]), _ptr(kv.v[0]), _ptr(kv.page_table), Hk,
                  _lib.LS_F16, int(g("non_leaf_len")), -1, 0, _stream())
        torch.cuda.synchronize()
        acc = int(st["scalars"][3])
        assert acc == int(g("acc_num")[0]), f"case {i}"
        assert np.array_equal(st["acc_ids"][:acc].cpu().numpy(), g("acc_ids")[:acc]), f"case {i}"
        assert int(st["aux"][2]) == int(g("double_input")[0])
        Ka, Va = kv.read(0, cache_len, f)
        assert np.array_equal(Ka[:acc, :, :16].float().cpu().numpy().astype(np.float16), g("K_after")[:acc]), f"case {i}"
        assert np.array_equal(Va[:acc, :, :16].float().cpu().numpy().astype(np.float16), g("V_after")[:acc]), f"case {i}"
        assert int(st["scalars"][0]) == cache_len + 1 and int(st["scalars"][1]) == cache_len + acc
        base = cache_len + 1 - (cache_len - 3)
        assert np.array_equal(st["output_ids"][base:base + acc].cpu().numpy(), g("acc_ids")[:acc])


@pytest.mark.parametrize("shape", [[4, 16, 16, 16, 16], [4, 4], [1, 1, 1, 1, 1], [8, 32, 32, 32, 32], [3, 5, 2]])
@pytest.mark.parametrize("logit_kind", ["fp32", "bf16"])
def test_tree_growth_and_verify_inputs_match_oracle(shape, logit_kind):
    """ls_round_begin + ls_tree_grow (all levels) + ls_build_verify vs oracle.tree on random logits.
    fp32 logits are tie-free; bf16 logits contain many exact ties and exercise the tie rule
    (score desc, flat index asc) that the oracle and the kernel share."""
    from longspec_b200 import _lib
    from longspec_b200.engine import _ptr, _stream
    V = 5003
    dev, st, ts = _setup(shape)
    offs = OT.level_offsets(shape)
    gamma, fseq = len(shape), offs[-1]
    gen = torch.Generator().manual_seed(7)
    acc_ids = np.array([11, 22, 33], dtype=np.int64)
    st["acc_ids"][:3] = torch.from_numpy(acc_ids)
    cache_len, draft_len = 500, 498
    st["scalars"][:4] = torch.tensor([cache_len, 503, draft_len, 3], dtype=torch.int32)
    ws = torch.empty(int(_lib.load().ls_grow_workspace_bytes(32, V)), dtype=torch.uint8, device=dev)
    _lib.call("ls_round_begin", C.byref(ts), _stream())
    tree_mask = OT.fresh_tree_mask(fseq)
    all_spec = np.zeros(fseq, dtype=np.int64)
    all_spec[0] = acc_ids[-1]
    scores = np.zeros(fseq, dtype=np.float32)
    torch.cuda.synchronize()
    assert st["step_ids"][:3].tolist() == acc_ids.tolist() and st["step_pos"][:3].tolist() == [498, 499, 500]
    assert int(st["aux"][0]) == draft_len + 3 and int(st["aux"][1]) == 2
    fp32 = logit_kind == "fp32"
    for lv in range(gamma):
        M = 1 if lv == 0 else shape[lv - 1]
        logits = torch.randn(M, V, generator=gen) * 3
        if not fp32:
            logits = logits.to(torch.bfloat16)
        dl = logits.to(dev)
        lo, hi = (0, 1) if lv == 0 else (offs[lv - 1], offs[lv])
        _lib.call("ls_tree_grow", C.byref(ts), _ptr(dl), int(fp32), _lib.LS_BF16, V, lv, lo, hi, shape[lv], _ptr(ws), _stream())
        logp = logits.float().log_softmax(-1).numpy()
        if lv == 0:
            OT.grow_root_level(tree_mask, all_spec, scores, logp[0], shape[0])
        else:
            OT.grow_level(tree_mask, all_spec, scores, logp, lv, offs)
        torch.cuda.synchronize()
        nb, ne = offs[lv], offs[lv + 1]
        assert np.array_equal(st["all_spec"][:ne].cpu().numpy(), all_spec[:ne]), f"level {lv} tokens"
        np.testing.assert_allclose(st["scores"][:ne].cpu().numpy(), scores[:ne], rtol=0, atol=2e-5)
        dense = _bits_to_dense(st["tree_bits"], fseq, fseq)
        assert np.array_equal(dense[:ne], tree_mask[:ne]), f"level {lv} mask"
        assert np.array_equal(st["parent"][:ne].cpu().numpy(), OT.parent_index(tree_mask)[:ne]) or lv == 0
        # next draft step inputs: this level's tokens at depth-based positions (llama_glide.py:1031-1033)
        n_new = ne - nb
        assert st["step_ids"][:n_new].tolist() == all_spec[nb:ne].tolist()
        exp_pos = OT.depth_positions(tree_mask[nb:ne, :ne], draft_len + 2)      # draft_len += acc_num - 1 (1018)
        assert st["step_pos"][:n_new].tolist() == exp_pos.tolist()
        assert int(st["scalars"][4]) == n_new
    assert int(st["scalars"][2]) == draft_len + 2
    _lib.call("ls_build_verify", C.byref(ts), _stream())
    torch.cuda.synchronize()
    veri, vmask = OT.build_verify_inputs(acc_ids, all_spec, tree_mask, gamma)
    qv = veri.shape[0]
    n_valid = 3 + fseq - 1
    assert np.array_equal(st["veri_ids"].cpu().numpy()[:n_valid], veri[:n_valid])
    assert np.array_equal(_bits_to_dense(st["veri_bits"], qv, qv)[:n_valid], vmask[:n_valid])
    assert np.array_equal(st["veri_pos"].cpu().numpy()[:n_valid], OT.depth_positions(vmask, cache_len)[:n_valid])
    assert int(st["scalars"][5]) == n_valid and int(st["scalars"][2]) == draft_len + 3
    assert st["logit_rows"].tolist() == list(range(2, 2 + fseq))


def test_argmax_rows_first_index_on_ties():
    from longspec_b200 import _lib
    from longspec_b200.engine import _ptr, _stream
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    for V in (128256, 32000, 5003):
        x = (torch.randn(7, V, generator=g) * 2).to(torch.bfloat16)
        x[3, 100] = x[3].max()          # duplicate maximum: first index wins
        x[3, 4000] = x[3].max()
        out = torch.zeros(7, dtype=torch.int64, device=dev)
        _lib.call("ls_argmax_rows", _ptr(x.to(dev)), 0, _lib.LS_BF16, 7, V, _ptr(out), _stream())
        torch.cuda.synchronize()
        exp = [int(np.argmax(r)) for r in x.float().numpy()]
        assert out.tolist() == exp


def test_pack_mask_matches_host_packing():
    from longspec_b200 import _lib
    from longspec_b200.engine import _ptr, _stream
    from gpu_util import pack_bits
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    m = (torch.rand(74, 150, generator=g) > 0.5).long()
    bits = torch.zeros(74 * 4, dtype=torch.int64, device=dev)
    md = m.to(dev)
    _lib.call("ls_pack_mask", _ptr(md), 74, 150, md.stride(0), _ptr(bits), _stream())
    torch.cuda.synchronize()
    assert torch.equal(bits.cpu(), pack_bits(m))


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("logit_kind", ["fp32", "bf16"])
def test_vocab_sharded_tree_growth_equals_unsharded(world, logit_kind):
    """ls_tree_grow_partial per vocabulary slice (as each rank of a vocab-sharded lm_head runs it) + ls_tree_grow_finish on
    the concatenated chunk blocks == ls_tree_grow on the whole vocabulary, bit for bit, incl. the tie rule (bf16 logits
    are full of exact ties) and slices that are short or empty (V = 5003 = 3 chunks of 2048 over up to 8 ranks)."""
    from longspec_b200 import _lib
    from longspec_b200.engine import _ptr, _stream
    shape, V = [4, 16, 16, 16, 16], 5003
    offs = OT.level_offsets(shape)
    fp32 = logit_kind == "fp32"
    lib = _lib.load()
    cb = int(lib.ls_grow_chunk_bytes())
    n_chunks = (V + 2047) // 2048
    cpr = (n_chunks + world - 1) // world
    results = []
    for sharded in (False, True):
        dev, st, ts = _setup(shape)
        st["acc_ids"][:3] = torch.tensor([11, 22, 33])
        st["scalars"][:4] = torch.tensor([500, 503, 498, 3], dtype=torch.int32)
        ws = torch.empty(int(lib.ls_grow_workspace_bytes(32, V)), dtype=torch.uint8, device=dev)
        _lib.call("ls_round_begin", C.byref(ts), _stream())
        gen = torch.Generator().manual_seed(7)
        for lv in range(len(shape)):
            M = 1 if lv == 0 else shape[lv - 1]
            logits = torch.randn(M, V, generator=gen) * 3
            if not fp32:
                logits = logits.to(torch.bfloat16)
            lo, hi = (0, 1) if lv == 0 else (offs[lv - 1], offs[lv])
            if not sharded:
                dl = logits.to(dev)
                _lib.call("ls_tree_grow", C.byref(ts), _ptr(dl), int(fp32), _lib.LS_BF16, V, lv, lo, hi, shape[lv], _ptr(ws), _stream())
            else:
                ws.fill_(0x7f)                                     # stale garbage must not leak into the merge
                keep = []
                for r in range(world):
                    c0 = r * cpr
                    v0, v1 = min(c0 * 2048, V), min((c0 + cpr) * 2048, V)
                    loc = logits[:, v0:v1].contiguous().to(dev) if v1 > v0 else torch.zeros(8, dtype=logits.dtype, device=dev)
                    keep.append(loc)
                    _lib.call("ls_tree_grow_partial", _ptr(loc), int(fp32), _lib.LS_BF16, M, v1 - v0, V, c0, cpr, shape[lv],
                              _ptr(ws[r * cpr * cb:]), _stream())
                _lib.call("ls_tree_grow_finish", C.byref(ts), V, cpr * world, lv, lo, hi, shape[lv], _ptr(ws), _stream())
            torch.cuda.synchronize()
        results.append({k: st[k].clone().cpu() for k in ("all_spec", "scores", "parent", "depth", "tree_bits", "step_ids", "step_pos")})
    for k in results[0]:
        assert torch.equal(results[0][k], results[1][k]), f"{k} differs between the sharded and the unsharded growth"


@pytest.mark.parametrize("world", [2, 5])
def test_vocab_sharded_argmax_equals_unsharded(world):
    from longspec_b200 import _lib
    from longspec_b200.engine import _ptr, _stream
    dev = torch.device("cuda:0")
    rows, V = 69, 5003
    logits = (torch.randn(rows, V, generator=torch.Generator().manual_seed(3)) * 2).to(torch.bfloat16)   # ties galore
    logits[5, 4000] = logits[5, 17] = 30.0                          # an exact tie across slices: the lower index wins
    dl = logits.to(dev)
    ref = torch.zeros(rows, dtype=torch.int64, device=dev)
    _lib.call("ls_argmax_rows", _ptr(dl), 0, _lib.LS_BF16, rows, V, _ptr(ref), _stream())
    n_chunks = (V + 2047) // 2048
    cpr = (n_chunks + world - 1) // world
    pairs = torch.zeros(world * rows * 2, dtype=torch.int64, device=dev)
    keep = []
    for r in range(world):
        v0, v1 = min(r * cpr * 2048, V), min((r + 1) * cpr * 2048, V)
        loc = logits[:, v0:v1].contiguous().to(dev) if v1 > v0 else torch.zeros(8, dtype=torch.bfloat16, device=dev)
        keep.append(loc)
        _lib.call("ls_argmax_rows_partial", _ptr(loc), 0, _lib.LS_BF16, rows, v1 - v0, v1 - v0, v0, _ptr(pairs[r * rows * 2:]), _stream())
    out = torch.zeros(rows, dtype=torch.int64, device=dev)
    _lib.call("ls_argmax_merge", _ptr(pairs), world, rows, _ptr(out), _stream())
    torch.cuda.synchronize()
    assert torch.equal(out, ref) and int(out[5]) == 17
    assert torch.equal(out.cpu(), logits.float().argmax(-1)) or True   # torch's tie order is unspecified; ours is first index

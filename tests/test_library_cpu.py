"""CPU-only checks of the C-ABI library and host logic (no compute calls without a GPU)."""
import ctypes
import inspect
import os
import re

import pytest
import torch


def test_library_loads_and_exports_every_declared_symbol():
    from longspec_b200 import _lib
    from longspec_b200.build import build
    build()
    lib = _lib.load()
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "longspec_b200.h")).read()
    declared = set(re.findall(r"\b(ls_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ls_attn_desc", "ls_tree_state"}
    assert len(declared) >= 18
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.exported_symbols(), f"{name} has no ctypes signature"
    assert lib.ls_version() >= 100


def test_struct_layouts_match_the_header():
    """ctypes mirrors must have the header's field order/sizes (a mismatch corrupts every call)."""
    from longspec_b200._lib import AttnDesc, TreeState
    assert ctypes.sizeof(AttnDesc) == 10 * 8 + 16 * 4 + 8 + 8
    assert ctypes.sizeof(TreeState) == 21 * 8 + 4 * 4
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "longspec_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    body = hdr[hdr.index("typedef struct ls_attn_desc {"):hdr.index("} ls_attn_desc;")]
    fields = re.findall(r"[\*\s]([a-z_0-9]+);", body)
    assert fields == [f[0] for f in AttnDesc._fields_]
    body = hdr[hdr.index("typedef struct ls_tree_state {"):hdr.index("} ls_tree_state;")]
    fields = re.findall(r"[\*\s]([a-z_0-9]+);", body)
    assert fields == [f[0] for f in TreeState._fields_]


def test_invalid_arguments_are_rejected_without_a_gpu():
    from longspec_b200 import _lib
    from longspec_b200._lib import AttnDesc
    lib = _lib.load()
    a = AttnDesc()
    assert lib.ls_attn_partial(ctypes.byref(a), None) == -1
    assert b"null" in lib.ls_last_error()
    with pytest.raises(_lib.LongSpecError):
        _lib.call("ls_argmax_rows", None, 0, 1, 0, 0, None, None)


def test_product_path_fails_loudly_without_cuda():
    from longspec_b200.llama_glide import LlamaGlide
    from longspec_b200.synthetic import PRESETS
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(Exception) as e:
        LlamaGlide(PRESETS["tiny"], "synthetic")
    assert "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    root = os.path.join(os.path.dirname(__file__), "..", "longspec_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"


def test_surface_matches_the_reference_signatures():
    """Same method names / parameters as llama_glide.py:552,915 (checked against the reference source
    when it is mounted; parameter lists are also pinned literally so the test is meaningful without it)."""
    from longspec_b200.llama_glide import LlamaGlide
    from longspec_b200.qwen2_glide import Qwen2Glide
    sig = inspect.signature(LlamaGlide.tree_spec_generate)
    assert list(sig.parameters)[:7] == ["self", "input_ids", "prompt_length", "tree_shape", "max_gen_len", "eos_id", "temperature"]
    assert sig.parameters["max_gen_len"].default == 64 and sig.parameters["eos_id"].default == 151645
    assert sig.parameters["temperature"].default == 0.0 and sig.parameters["tree_shape"].default is None
    sig = inspect.signature(LlamaGlide.vanilla_generate)
    assert list(sig.parameters) == ["self", "input_ids", "prompt_length", "max_gen_len", "eos_id"]
    assert list(inspect.signature(LlamaGlide.__init__).parameters)[:4] == ["self", "config", "target_model_path", "glide_path"]
    for name in ("vanilla_torch_generate", "spec_generate", "magicdec_generate"):
        assert hasattr(LlamaGlide, name) and hasattr(Qwen2Glide, name)
    ref = "/root/reference/longspec/test/llama_glide.py"
    if os.path.exists(ref):
        src = open(ref).read()
        assert "def tree_spec_generate(self, input_ids, prompt_length, tree_shape=None, max_gen_len=64, eos_id=151645, temperature=0.0)" in src
        assert "def vanilla_generate(self, input_ids, prompt_length, max_gen_len=64, eos_id=151645)" in src


def test_shard_partition_covers_the_prompt():
    from longspec_b200.engine import Shard
    for P in (1000, 16384, 131072, 131073):
        for W in (1, 2, 4, 8):
            sh = [Shard.make(P, r, W) for r in range(W)]
            assert sh[0].begin == 0 and sh[-1].end == -1 and sh[-1].owner
            for a, b in zip(sh[:-1], sh[1:]):
                assert a.end == b.begin and a.begin % 128 == 0 and not a.owner
            assert sh[-1].begin <= P


def test_checkpoint_directories_load_with_the_reference_key_names(tmp_path):
    """HF-style safetensors directories (target: `model.*` / `lm_head.*`; draft: LlamaGlideDecoderLayer keys without a
    prefix, llama_glide.py:480) are read shard by shard; the draft's keys get the `glide.` prefix Weights expects."""
    from safetensors.torch import save_file
    from longspec_b200.llama_glide import _load_dir
    from longspec_b200.synthetic import PRESETS, synthetic_state_dict
    dims = PRESETS["tiny-gqa"]
    sd = synthetic_state_dict(dims, seed=5)
    target = {k: v.contiguous() for k, v in sd.items() if not k.startswith("glide.")}
    draft = {k[len("glide."):]: v.contiguous() for k, v in sd.items() if k.startswith("glide.")}
    tdir, gdir = tmp_path / "target", tmp_path / "glide"
    tdir.mkdir()
    gdir.mkdir()
    keys = sorted(target)
    half = len(keys) // 2
    save_file({k: target[k] for k in keys[:half]}, str(tdir / "model-00001-of-00002.safetensors"))
    save_file({k: target[k] for k in keys[half:]}, str(tdir / "model-00002-of-00002.safetensors"))
    save_file(draft, str(gdir / "model.safetensors"))
    got = _load_dir(str(tdir))
    got.update(_load_dir(str(gdir), prefix="glide."))
    assert sorted(got) == sorted(sd)
    for k in sd:
        assert torch.equal(got[k], sd[k]), k
    with pytest.raises(FileNotFoundError):
        _load_dir(str(tmp_path / "missing"))


def _tc_roles(n_slots, nrs, all_tiles):
    """Host-side mirror of the role / tile assignment of attn_tc_kernel (attn_tc.cu, "role of this CTA"): returns one
    dict per CTA of a kv head.  Kept next to the kernel's integer logic as an executable statement of its contract."""
    nrl_max = 2
    grid_x = n_slots + (n_slots + 1) // 2 if nrs > nrl_max else n_slots
    per = (all_tiles + n_slots - 1) // n_slots
    ctas = []
    for bx in range(grid_x):
        slot, rt0, share0, nshare, first, empty_slot = bx, 0, bx, 1, 0, -1
        if nrs > nrl_max:
            light = bx >= n_slots
            g = bx - n_slots if light else bx >> 1
            share0 = 2 * g
            nshare = min(2, n_slots - share0)
            if light:
                slot, rt0, empty_slot = share0, nrl_max, (share0 + 1 if nshare == 2 else -1)
            else:
                first = bx & 1
        stride = nshare if (nrs > nrl_max and rt0 == 0) else 1
        lo = min(share0 * per, all_tiles) + first
        hi = min((share0 + nshare) * per, all_tiles)
        tiles = list(range(lo, hi, stride)) if hi > lo else []
        rows = list(range(rt0, min(nrs, rt0 + nrl_max)))
        ctas.append(dict(slot=slot, rows=rows, tiles=tiles, empty_slot=empty_slot))
    return ctas


@pytest.mark.parametrize("nrs", [1, 2, 3])
def test_tc_kernel_role_assignment_covers_every_tile_and_slot_once(nrs):
    """Every (key tile, row tile) is processed by exactly one CTA, and every (partial slot, row tile) has exactly one
    writer (a real one or the light CTA's explicit empty partial) - the contract ls_attn_combine relies on."""
    for n_slots in range(1, 20):
        for all_tiles in (0, 1, 2, 3, 7, 12, 13, 64, 1031):
            ctas = _tc_roles(n_slots, nrs, all_tiles)
            seen, writers = {}, {}
            for c in ctas:
                for r in c["rows"]:
                    writers[(c["slot"], r)] = writers.get((c["slot"], r), 0) + 1
                    if c["empty_slot"] >= 0:
                        writers[(c["empty_slot"], r)] = writers.get((c["empty_slot"], r), 0) + 1
                    for t in c["tiles"]:
                        seen[(t, r)] = seen.get((t, r), 0) + 1
            assert all(v == 1 for v in seen.values())
            assert set(seen) == {(t, r) for t in range(all_tiles) for r in range(nrs)}
            assert writers == {(s, r): 1 for s in range(n_slots) for r in range(nrs)}, (n_slots, nrs, all_tiles)
            if nrs == 3:     # trios are balanced: a light CTA runs as many (tile, row tile) steps as each of its heavy CTAs (+-1 tile)
                for c in ctas:
                    if c["rows"] == [2] and c["empty_slot"] >= 0:
                        heavy = [h for h in ctas if h["rows"] == [0, 1] and h["slot"] in (c["slot"], c["empty_slot"])]
                        assert all(abs(2 * len(h["tiles"]) - len(c["tiles"])) <= 1 for h in heavy)


def _tc1_roles(n_slots, all_tiles):
    """Host-side mirror of attn_tc1_kernel's role assignment (one row tile per CTA, three row tiles per call): trio g = the
    key tiles of slots 2g, 2g+1; CTAs [0, 2T) = row tile bx & 1 of trio bx >> 1, CTAs [2T, 3T) = row tile 2."""
    n_trios = (n_slots + 1) >> 1
    per = (all_tiles + n_slots - 1) // n_slots
    ctas = []
    for bx in range(3 * n_trios):
        light = bx >= 2 * n_trios
        g = bx - 2 * n_trios if light else bx >> 1
        rt = 2 if light else bx & 1
        share0 = 2 * g
        nshare = min(2, n_slots - share0)
        lo, hi = min(share0 * per, all_tiles), min((share0 + nshare) * per, all_tiles)
        ctas.append(dict(slot=share0, empty_slot=share0 + 1 if nshare == 2 else -1, row=rt, tiles=list(range(lo, hi)),
                         cluster=bx >> 1, heavy=not light))
    return ctas


def test_tc1_kernel_role_assignment_covers_every_tile_and_slot_once():
    """attn_tc1_kernel: every (key tile, row tile) has one CTA, every (slot, row tile) one writer; the two CTAs of a
    multicast cluster (LONGSPEC_TC1=2: blockIdx.x pairs) are the row tiles 0 / 1 of ONE trio and walk the same tiles."""
    for n_slots in range(1, 20):
        for all_tiles in (0, 1, 2, 3, 7, 12, 13, 64, 1031):
            ctas = _tc1_roles(n_slots, all_tiles)
            seen, writers = {}, {}
            for c in ctas:
                writers[(c["slot"], c["row"])] = writers.get((c["slot"], c["row"]), 0) + 1
                if c["empty_slot"] >= 0:
                    writers[(c["empty_slot"], c["row"])] = writers.get((c["empty_slot"], c["row"]), 0) + 1
                for t in c["tiles"]:
                    seen[(t, c["row"])] = seen.get((t, c["row"]), 0) + 1
            assert all(v == 1 for v in seen.values())
            assert set(seen) == {(t, r) for t in range(all_tiles) for r in range(3)}
            assert writers == {(s, r): 1 for s in range(n_slots) for r in range(3)}, (n_slots, all_tiles)
            heavy = [c for c in ctas if c["heavy"]]
            for a, b in zip(heavy[0::2], heavy[1::2]):
                assert a["cluster"] == b["cluster"] and (a["row"], b["row"]) == (0, 1) and a["tiles"] == b["tiles"]
            assert len({len(c["tiles"]) for c in ctas if c["slot"] == 0}) == 1      # a trio's three CTAs run the same number of pairs

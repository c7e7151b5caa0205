"""CPU tests of the host-side logic added in round 2: vocabulary slicing of the sharded lm_head, prefill split plan,
bench workload records, seam wrappers' refusal to run without CUDA, checkpoint resolution, rope frequency handling."""
import json
import os

import pytest
import torch

from longspec_b200.engine import Shard, split_plan
from longspec_b200.synthetic import PRESETS, ModelDims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("vocab", [320, 5003, 32000, 128256, 152064])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_vocab_slices_tile_the_vocabulary_exactly_once(vocab, world):
    """Engine.begin's slicing rule: chunk-aligned (2048 ids), equal chunk counts per rank, last slices short or empty."""
    n_chunks = (vocab + 2047) // 2048
    cpr = (n_chunks + world - 1) // world
    covered = []
    for r in range(world):
        c0 = r * cpr
        v0, v1 = min(c0 * 2048, vocab), min((c0 + cpr) * 2048, vocab)
        assert v0 % 2048 == 0 or v0 == vocab
        covered += list(range(v0, v1, 2048))
        assert v1 - v0 <= cpr * 2048
    assert covered == list(range(0, vocab, 2048)), "every chunk belongs to exactly one rank"
    assert cpr * world >= n_chunks and cpr * world - n_chunks < world + cpr


def test_prefill_split_plan_never_exceeds_the_workspace():
    """Engine._prefill_splits(end) <= the slot count the workspaces were sized for (split_plan for one query row)."""
    class E:                      # the two attributes _prefill_splits reads
        pass
    from longspec_b200.engine import Engine
    for preset, P, world in [("llama3-8b-262k", 131072, 1), ("llama3-8b-262k", 131072, 8), ("vicuna-13b-16k", 16384, 1),
                             ("tiny-g4", 3000, 2), ("tiny", 300, 1)]:
        d = PRESETS[preset]
        for rank in range(world):
            e = E()
            e.dims, e.sms, e.shard = d, 148, Shard.make(P, rank, world)
            total = P + 64 + 256 + 74 + 128
            cap = max(split_plan(d.heads, d.kv_heads, 148, P, total, world, (74, 16, 1)))
            for end in (1, 64, 129, P // 3, P - 1, P):
                n = Engine._prefill_splits(e, end)
                assert 1 <= n <= max(cap, 148 // d.kv_heads), (preset, rank, end, n, cap)


def test_bench_workload_records_are_consistent():
    """tests/golden/bench_workload.json (made on a B200 by tools/make_bench_golden.py): every record's tree stream was
    checked against its greedy stream when it was made; here: shapes, and that bench.py finds the default workload."""
    import bench
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_workload.json")))
    assert bench.workload_key() in g, "the default bench workload must have a committed greedy record"
    for key, rec in g.items():
        assert len(rec["greedy"]) >= 128 and all(isinstance(t, int) for t in rec["greedy"])
        tpr = rec["tokens_per_round"]
        gamma = len(key.split("tree=")[1].split("|")[0].split())
        assert len(tpr) >= 25 and all(1 <= t <= gamma + 1 for t in tpr)
        assert rec["tree_tokens_checked"] >= 100
    tau, src = bench.golden_tau(5, 20, 3.25)
    assert 1.0 <= tau <= 6.0 and "seeded workload" in src
    tau2, src2 = bench.golden_tau(5, 10 ** 6, 3.25)
    assert tau2 == 3.25 and "published" in src2


def test_seam_wrappers_refuse_cpu_tensors():
    from longspec_b200 import _lib
    from longspec_b200.seams import attention, tree_verification
    q = torch.zeros(1, 4, 3, 128, dtype=torch.float16)
    k = torch.zeros(1, 2, 5, 128, dtype=torch.float16)
    with pytest.raises(_lib.LongSpecError):
        attention(q, k, k, torch.ones(3, 5, dtype=torch.int64))
    with pytest.raises(NotImplementedError):
        attention(torch.zeros(1, 4, 3, 64), torch.zeros(1, 2, 5, 64), torch.zeros(1, 2, 5, 64), torch.ones(3, 5))
    with pytest.raises(_lib.LongSpecError):
        tree_verification(torch.zeros(1, 9, dtype=torch.int64), torch.zeros(1, 9, dtype=torch.int64),
                          torch.ones(1, 9, 9, dtype=torch.int64), torch.tensor([100], dtype=torch.int32), 5)


def test_hub_ids_resolve_through_the_local_cache_only(tmp_path, monkeypatch):
    from longspec_b200.llama_glide import _load_dir, _resolve_checkpoint_dir
    monkeypatch.setenv("HF_HOME", str(tmp_path / "hf"))
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    with pytest.raises(FileNotFoundError) as e:
        _resolve_checkpoint_dir("nobody/not-a-model")
    assert "local Hugging Face cache" in str(e.value)
    d = tmp_path / "ckpt"
    d.mkdir()
    assert _resolve_checkpoint_dir(str(d)) == str(d)
    with pytest.raises(FileNotFoundError):
        _load_dir(str(d))                                   # a directory without weights
    torch.save({"model.norm.weight": torch.ones(4)}, d / "pytorch_model.bin")
    assert list(_load_dir(str(d), prefix="x.")) == ["x.model.norm.weight"]


def test_rope_types():
    from transformers import LlamaConfig
    base = dict(hidden_size=4096, num_attention_heads=32, num_key_value_heads=8, num_hidden_layers=2, intermediate_size=14336,
                vocab_size=128256, max_position_embeddings=131072, rope_theta=500000.0)
    d = ModelDims.from_hf_config(LlamaConfig(**base))
    assert d.inv_freq is None and d.rope_factor == 1.0
    d = ModelDims.from_hf_config(LlamaConfig(**base, rope_scaling={"rope_type": "linear", "factor": 4.0}))
    assert d.inv_freq is None and d.rope_factor == 4.0
    d = ModelDims.from_hf_config(LlamaConfig(**base, rope_scaling={"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0,
                                                                   "high_freq_factor": 4.0, "original_max_position_embeddings": 8192}))
    assert d.inv_freq is not None and len(d.inv_freq) == 64
    plain = [1.0 / (500000.0 ** (i / 64)) for i in range(64)]
    assert abs(d.inv_freq[0] - plain[0]) < 1e-9 and abs(d.inv_freq[-1] - plain[-1] / 8.0) < 1e-12    # high freq kept, low freq / factor
    with pytest.raises(NotImplementedError):
        ModelDims.from_hf_config(LlamaConfig(**base, rope_scaling={"rope_type": "dynamic", "factor": 2.0}))

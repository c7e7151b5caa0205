"""Record the facts about bench.py's SEEDED synthetic workload that both bench arms and the N > 1 runs check against
(tests/golden/bench_workload.json):

  greedy            the target's greedy continuation of the prompt (LlamaGlide.vanilla_generate on ONE GPU) - what any
                    lossless tree-spec run, single- or multi-GPU, must emit token for token
  tokens_per_round  tokens emitted by every round of the tree-spec generation (accepted length = a property of the seeded
                    workload at T = 0), so that `bench.py --impl reference` uses the same accepted length as the GPU arm

    python tools/make_bench_golden.py            (on a B200; env as bench.py: LONGSPEC_BENCH_MODEL/_PROMPT/_TREE/_DRAFT_SCALE)
writes gpurun_out/bench_workload_<key>.json; `python tools/make_bench_golden.py merge` (anywhere) folds every such
file into tests/golden/bench_workload.json.
"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden", "bench_workload.json")

if len(sys.argv) > 1 and sys.argv[1] == "merge":
    merged = json.load(open(GOLDEN)) if os.path.exists(GOLDEN) else {}
    for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "bench_workload_*.json"))):
        merged.update(json.load(open(f)))
    with open(GOLDEN, "w") as fh:
        json.dump(merged, fh, indent=1, sort_keys=True)
    print(f"{GOLDEN}: {len(merged)} workloads: {sorted(merged)}")
    sys.exit(0)

import torch  # noqa: E402

import bench  # noqa: E402
from longspec_b200.llama_glide import LlamaGlide  # noqa: E402
from longspec_b200.synthetic import PRESETS, synthetic_prompt  # noqa: E402

N_GREEDY = int(os.environ.get("GOLDEN_TOKENS", 192))
N_ROUNDS = int(os.environ.get("GOLDEN_ROUNDS", 48))
dims = PRESETS[bench.MODEL]
spec = "synthetic:seed=0" + (f",draft_scale={bench.DRAFT_SCALE}" if bench.DRAFT_SCALE else "")
m = LlamaGlide(dims, spec, None, dtype=torch.bfloat16, device="cuda:0")
ids = synthetic_prompt(dims.vocab, bench.PROMPT, seed=0).pin_memory()
plen = torch.tensor([bench.PROMPT])
eos = dims.vocab + 5
out, num, el = m.vanilla_generate(ids, plen, max_gen_len=N_GREEDY, eos_id=eos)
greedy = out[0].cpu().tolist()
print(f"vanilla: {num} steps in {el:.2f} s ({el / max(num, 1) * 1e3:.2f} ms/token)", file=sys.stderr)
gamma = len(bench.TREE)
max_gen = (2 * N_ROUNDS + 6) * (gamma + 1) + 64
out, count, num, el, _ = m.tree_spec_generate(ids, plen, tree_shape=bench.TREE, max_gen_len=max_gen, eos_id=eos,
                                              max_rounds=N_ROUNDS)
tree = out[0].cpu().tolist()
n_tok = int(count) + num
written = [1] + [t for _, t in m.last_stats["round_log"]]        # tokens in output_ids after the prefill / after each round
per_round = [written[i + 1] - written[i] for i in range(len(written) - 1)]
n_cmp = min(n_tok, len(greedy))
assert tree[:n_cmp] == greedy[:n_cmp], "tree-spec output != greedy continuation: the workload is not usable as a golden"
rec = {bench.workload_key(): {
    "greedy": greedy, "tokens_per_round": per_round,
    "tau_all_rounds": sum(per_round) / len(per_round),
    "made_by": "tools/make_bench_golden.py on 1 x B200 (vanilla_generate / tree_spec_generate of this repo; the tree stream was "
               "checked against the greedy stream when the record was made)",
    "tree_tokens_checked": n_cmp, "ms_per_round": el / num * 1e3}}
od = os.path.join(ROOT, "gpurun_out")
os.makedirs(od, exist_ok=True)
name = "".join(c if c.isalnum() else "_" for c in bench.workload_key())
with open(os.path.join(od, f"bench_workload_{name}.json"), "w") as fh:
    json.dump(rec, fh)
print(json.dumps({"key": bench.workload_key(), "tokens_per_round": per_round, "tau": rec[bench.workload_key()]["tau_all_rounds"],
                  "tree_tokens_checked": n_cmp, "ms_per_round": el / num * 1e3}))

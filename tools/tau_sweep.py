"""Accepted length tau of the synthetic "permutation" weights vs the draft's noise scale (picks bench defaults)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longspec_b200.llama_glide import LlamaGlide  # noqa: E402
from longspec_b200.synthetic import PRESETS, synthetic_prompt  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "llama3-8b-262k"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dims = PRESETS[model]
ids = synthetic_prompt(dims.vocab, P, seed=0)
for ds in [float(x) for x in (sys.argv[3:] or ["1.0", "1.5", "2.0", "2.5", "3.0", "4.0"])]:
    m = LlamaGlide(dims, f"synthetic:seed=0,draft_scale={ds}", None, dtype=torch.bfloat16)
    out, count, num, el, _ = m.tree_spec_generate(ids.cuda(), torch.tensor([P]).cuda(), max_gen_len=400, eos_id=dims.vocab + 5)
    vout, vnum, vel = m.vanilla_generate(ids.cuda(), torch.tensor([P]).cuda(), max_gen_len=48)
    n = int(count) + num
    same = bool((out[0, :48] == vout[0, :48]).all())
    print(f"draft_scale={ds}: tau={(n) / num:.2f} rounds={num} tokens={n} {el / num * 1e3:.2f} ms/round  lossless(48)={same}  vanilla {vel / max(vnum, 1) * 1e3:.2f} ms/token", flush=True)
    del m
    torch.cuda.empty_cache()

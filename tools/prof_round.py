"""Per-kernel time breakdown of one decode round (torch.profiler / CUPTI), eager launches.
   python tools/prof_round.py [prompt=16384] [model=llama3-8b-262k]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longspec_b200.llama_glide import LlamaGlide  # noqa: E402
from longspec_b200.synthetic import PRESETS, synthetic_prompt  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
model = sys.argv[2] if len(sys.argv) > 2 else "llama3-8b-262k"
dims = PRESETS[model]
m = LlamaGlide(dims, "synthetic:seed=0", None, dtype=torch.bfloat16)
m.use_graph = False
ids = synthetic_prompt(dims.vocab, P, seed=0)
out = m.tree_spec_generate(ids.cuda(), torch.tensor([P]).cuda(), max_gen_len=256, eos_id=dims.vocab + 5, max_rounds=3)
eng = m.eng
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    eng.round(-1)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in ev:
    k = e.name[:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
print(f"kernels: {sum(v[0] for v in agg.values())}, total device time {tot / 1e3:.2f} ms")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:28]:
    print(f"{v[1] / 1e3:8.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  avg {v[1] / v[0]:7.1f} us  {k}")
t0 = min(e.time_range.start for e in ev)
t1 = max(e.time_range.end for e in ev)
print(f"span {(t1 - t0) / 1e3:.2f} ms")

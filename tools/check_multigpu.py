"""torchrun --nproc-per-node N tools/check_multigpu.py : sequence-sharded generation must emit exactly the
tokens of the single-GPU generation (run on rank 0 afterwards).  Uses a small synthetic model."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longspec_b200.llama_glide import LlamaGlide  # noqa: E402
from longspec_b200.synthetic import PRESETS, synthetic_prompt, synthetic_state_dict  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
preset = sys.argv[1] if len(sys.argv) > 1 else "tiny-gqa"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
dims = PRESETS[preset]
W = synthetic_state_dict(dims, seed=7)
ids = synthetic_prompt(dims.vocab, P, seed=7)
kw = dict(tree_shape=[4, 16, 16, 16, 16], max_gen_len=96, eos_id=dims.vocab + 5)
m = LlamaGlide(dims, W, None, dtype=torch.bfloat16, group=dist.group.WORLD)
out, count, num, el, _ = m.tree_spec_generate(ids.cuda(), torch.tensor([P]).cuda(), **kw)
sh = m.eng.shard
print(f"[rank {rank}] shard [{sh.begin},{sh.end}) owner={sh.owner} tokens={int(count) + num} rounds={num} {el * 1e3:.1f} ms", flush=True)
gathered = [torch.empty_like(out) for _ in range(world)]
dist.all_gather(gathered, out)
assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks disagree on the generated tokens"
dist.barrier()
if rank == 0:
    m1 = LlamaGlide(dims, W, None, dtype=torch.bfloat16)
    ref, c1, n1, _, _ = m1.tree_spec_generate(ids.cuda(), torch.tensor([P]).cuda(), **kw)
    assert torch.equal(ref, out), f"sharded != single GPU\n{ref}\n{out}"
    assert (int(c1), n1) == (int(count), num)
    print(f"multi-GPU check ok: world={world}, {int(count) + num} tokens identical to the single-GPU run", flush=True)
# teardown in dependency order: the captured graphs (they hold NCCL kernels) before the communicator
m.last_graph = None
m.eng = None
del m
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()

"""Drop-in for the reference's `llama_glide` module (longspec/test/llama_glide.py).

`LlamaGlide(config, target_model_path, glide_path=None)` exposes the generation methods the
unchanged drivers call (inference_long-bench.py:112,132-250): `tree_spec_generate`,
`vanilla_generate`, `vanilla_torch_generate`, `spec_generate`, `magicdec_generate`, with the
reference's argument names and return tuples.  The decode loop runs on hand-written sm_100a
kernels through the C ABI (include/longspec_b200.h); there is no Triton, no flash-attn and no
CPU fallback on this path.

Weights: `target_model_path` / `glide_path` are local HF checkpoint directories (safetensors),
a state dict, or the string "synthetic[:seed=N,scale=X,draft_scale=Y]" for seeded random weights
of the config's architecture (no checkpoints are reachable offline).
Environment knobs (the drivers cannot pass new arguments): LONGSPEC_DTYPE=float16|bfloat16
(default float16 like llama_glide.py:474), LONGSPEC_CUDA_GRAPH=0|1 (default 1).
"""
from __future__ import annotations

import glob
import os
import time

import torch
import torch.nn.functional as F

from . import _lib
from .engine import Engine, Shard, Weights
from .synthetic import ModelDims, synthetic_state_dict

_DTYPES = {"float16": torch.float16, "fp16": torch.float16, "bfloat16": torch.bfloat16, "bf16": torch.bfloat16}


def _load_dir(path, prefix=""):
    from safetensors.torch import load_file
    files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {path!r} (offline: hub ids cannot be resolved)")
    sd = {}
    for f in files:
        for k, v in load_file(f).items():
            sd[prefix + k] = v
    return sd


def _parse_synth(spec):
    kw = {"seed": 0, "scale": 1.0, "draft_scale": None}
    if ":" in spec:
        for item in spec.split(":", 1)[1].split(","):
            k, v = item.split("=")
            kw[k] = int(v) if k == "seed" else float(v)
    return kw


class LlamaGlide:
    """Same constructor and generation surface as llama_glide.py:471-1245 (T = 0 tree path)."""

    variant = "llama"        # Qwen2Glide overrides: output fill / eos test differ (qwen2_glide.py:766,949)

    def __init__(self, config, target_model_path, glide_path=None, *, dtype=None, device=None,
                 shard: Shard | None = None, group=None, shuffle_pages=False):
        _lib.load()          # fail loudly if the CUDA library is missing
        if not torch.cuda.is_available():
            raise _lib.LongSpecError("LlamaGlide needs a CUDA device (B200); there is no CPU path")
        self.config = config
        self.dims = config if isinstance(config, ModelDims) else ModelDims.from_hf_config(config)
        if self.dims.head_dim != 128:
            raise ValueError("head_dim must be 128 (the reference hard-codes 1/sqrt(128), llama.py:95)")
        self.dtype = dtype or _DTYPES[os.environ.get("LONGSPEC_DTYPE", "float16")]
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.use_graph = os.environ.get("LONGSPEC_CUDA_GRAPH", "1") != "0"
        sd = self._resolve_weights(target_model_path, glide_path)
        self.weights = Weights(self.dims, sd, self.dtype, self.device)
        del sd
        self.shard_spec = shard
        self.group = group
        self.shuffle_pages = shuffle_pages
        self.eng = None
        self.last_stats = {}

    # -------------------------------------------------------------- weights
    def _resolve_weights(self, target, glide):
        if isinstance(target, dict):
            sd = dict(target)
        elif isinstance(target, str) and target.startswith("synthetic"):
            kw = _parse_synth(target)
            sd = synthetic_state_dict(self.dims, seed=kw["seed"], device=str(self.device), dtype=self.dtype,
                                      scale=kw["scale"], draft_scale=kw["draft_scale"])
        else:
            sd = _load_dir(target)
        if glide is None or (isinstance(glide, str) and glide.startswith("synthetic")):
            if not any(k.startswith("glide.") for k in sd):
                # llama_glide.py:477-478: a freshly initialised draft layer
                extra = synthetic_state_dict(self.dims, seed=1, device=str(self.device), dtype=self.dtype)
                sd.update({k: v for k, v in extra.items() if k.startswith("glide.")})
        elif isinstance(glide, dict):
            sd.update({(k if k.startswith("glide.") else "glide." + k): v for k, v in glide.items()})
        else:
            sd.update(_load_dir(glide, prefix="glide."))
        return sd

    @property
    def eos_token_id(self):
        return getattr(self.config, "eos_token_id", None)

    # -------------------------------------------------------------- helpers
    def _new_engine(self, prompt_len, max_gen_len, tree_shape):
        shard = self.shard_spec
        if shard is None and self.group is not None:
            import torch.distributed as dist
            shard = Shard.make(prompt_len, dist.get_rank(self.group), dist.get_world_size(self.group))
        self.eng = None
        torch.cuda.empty_cache()
        eng = Engine(self.weights, self.dtype, self.device, shard, self.group, self.shuffle_pages)
        eng.begin(prompt_len, max_gen_len, tree_shape)
        self.eng = eng
        return eng

    def _first_token(self, eng, input_ids):
        last = eng.prefill(input_ids)
        return int(F.linear(last, self.weights.lm_head).argmax(-1))

    # -------------------------------------------------------------- vanilla (llama_glide.py:552-585)
    def vanilla_generate(self, input_ids, prompt_length, max_gen_len=64, eos_id=151645):
        assert input_ids is not None, "please give the input"
        assert input_ids.size(0) == 1, "batch size 1 (inference_long-bench.py:119)"
        P = int(prompt_length.reshape(-1)[0]) if torch.is_tensor(prompt_length) else int(prompt_length)
        eng = self._new_engine(P, max_gen_len, [1])
        out = torch.zeros(max_gen_len, dtype=torch.int64, device=self.device)
        out[0] = self._first_token(eng, input_ids)
        eng.set_lengths(P, P, P, [int(out[0])])
        stop_id = self.eos_token_id
        num = 0
        pos = eng.p_cache_len
        tok_in = out[0:1].clone()
        tok_out = torch.zeros(1, dtype=torch.int64, device=self.device)
        idx = torch.ones(1, dtype=torch.int64, device=self.device)

        def step():
            # one token: append at cache_len, causal attention, argmax (llama_glide.py:572-577), all on the device
            eng.decode_one(tok_in, pos, tok_out)
            eng.p_cache_len.add_(1)
            out.scatter_(0, idx, tok_out)
            idx.add_(1)
            tok_in.copy_(tok_out)

        graph = None
        torch.cuda.synchronize()
        t0 = time.time()
        for i in range(1, max_gen_len):
            if graph is not None:
                graph.replay()
            else:
                step()
                if self.use_graph and i == 1:
                    graph = torch.cuda.CUDAGraph()
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        with torch.cuda.graph(graph, stream=side):
                            step()
                    torch.cuda.current_stream().wait_stream(side)
            num += 1
            if stop_id is not None and i % 16 == 0 and bool((out[:i + 1] == stop_id).any()):
                break           # the reference tests every step (578); the tail is trimmed below
        torch.cuda.synchronize()
        elapsed = time.time() - t0
        if stop_id is not None:
            hit = (out == stop_id).nonzero()
            if hit.numel():
                cut = int(hit[0]) + 1
                num = min(num, cut - 1)
                out[cut:] = 0
        return out[None], num, elapsed

    def vanilla_torch_generate(self, input_ids, prompt_length, max_gen_len=64, eos_id=151645):
        """The reference's pure-torch attention baseline (llama_glide.py:587-619) produces the same
        tokens as vanilla_generate; here both run the same kernels."""
        return self.vanilla_generate(input_ids, prompt_length, max_gen_len, eos_id)

    # -------------------------------------------------------------- tree (llama_glide.py:915-1126)
    def tree_spec_generate(self, input_ids, prompt_length, tree_shape=None, max_gen_len=64, eos_id=151645,
                           temperature=0.0, max_rounds=None):
        assert input_ids is not None, "please give the input"
        assert input_ids.size(0) == 1, "batch size 1 (inference_long-bench.py:119)"
        if temperature > 0:
            raise NotImplementedError("stochastic tree verification (llama_glide.py:1177-1245) is out of scope: "
                                      "parity is defined at temperature 0")
        shape = [4, 16, 16, 16, 16] if tree_shape is None else list(tree_shape)
        P = int(prompt_length.reshape(-1)[0]) if torch.is_tensor(prompt_length) else int(prompt_length)
        eng = self._new_engine(P, max_gen_len, shape)
        st = eng.st
        fill = eos_id if self.variant == "llama" else 0                     # llama_glide.py:937 / qwen2_glide.py:766
        stop_id = self.eos_token_id if self.variant == "llama" else eos_id    # 1120 / qwen2_glide.py:949
        stop_id = -1 if stop_id is None else int(stop_id)
        first = self._first_token(eng, input_ids)
        st["output_ids"].fill_(fill)
        st["output_ids"][0] = first
        eng.set_lengths(P, P, P, [first])
        st["counters"].copy_(torch.tensor([0, 1, 1, 0], dtype=torch.int32))
        # the reference tests the WHOLE output buffer (pre-fill value and first token included)
        sticky_stop = stop_id >= 0 and (first == stop_id or (fill == stop_id and max_gen_len > 1))
        # Pipelined host loop: round r+1 is enqueued BEFORE round r's 16-byte result is read, so the GPU never waits
        # for the host (graph launch, D2H, Python).  A round enqueued after the stop condition is harmless:
        # ls_tree_accept is a no-op once the device-side stop flag is set.
        hosts = [torch.zeros(4, dtype=torch.int32).pin_memory() for _ in range(2)]
        events = [torch.cuda.Event(), torch.cuda.Event()]
        graph = None
        rounds = 0
        round_log = []                       # (host time, tokens written) after every round's stop test
        launches0 = eng.launches
        limit = max_gen_len - 1 if max_rounds is None else min(max_gen_len - 1, max_rounds)

        queued = 0                           # rounds enqueued so far

        def enqueue():
            nonlocal queued
            if graph is not None:
                graph.replay()
            else:
                eng.round(stop_id)
            hosts[queued % 2].copy_(st["counters"], non_blocking=True)
            events[queued % 2].record()
            queued += 1

        torch.cuda.synchronize()
        start_time = time.time()
        host = hosts[0]
        if limit >= 1:
            enqueue()                        # round 0 runs eagerly
        while rounds < queued:
            if rounds == 0:
                events[0].synchronize()
                self.launches_per_round = eng.launches - launches0
                if self.use_graph and not sticky_stop:
                    graph = self._capture_round(eng, stop_id)
            # keep one more round in flight before blocking on this round's result
            if queued == rounds + 1 and queued < limit and not sticky_stop:
                enqueue()
            events[rounds % 2].synchronize()
            host = hosts[rounds % 2]
            stop_now = bool(host[3] != 0) or sticky_stop
            rounds += 1
            round_log.append((time.time(), int(host[2])))
            if stop_now or rounds >= limit:
                break
        host = host.clone()
        elapsed_time = time.time() - start_time
        count = torch.tensor(int(host[0]), device=self.device)
        num = int(host[1])
        self.last_stats = dict(rounds=rounds, launches_per_round=getattr(self, "launches_per_round", 0),
                               graph=graph is not None, round_log=round_log, start_time=start_time)
        self.last_graph = graph
        spec_mask = torch.zeros(1, max_gen_len, dtype=torch.int64, device=self.device)
        return st["output_ids"].clone()[None], count, num, elapsed_time, spec_mask

    def _capture_round(self, eng, stop_id):
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                eng.round(stop_id)
        torch.cuda.current_stream().wait_stream(s)
        return g

    # -------------------------------------------------------------- other driver methods
    def spec_generate(self, *a, **k):
        raise NotImplementedError("--method seq (llama_glide.py:621-774) is outside the tree-spec hot path")

    def magicdec_generate(self, *a, **k):
        raise NotImplementedError("--method magicdec (llama_glide.py:776-913) is a competitor baseline, out of scope")

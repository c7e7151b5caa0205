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


def _resolve_checkpoint_dir(path):
    """A local directory, or a hub id resolved through the LOCAL HF cache (the reference drivers pass hub ids,
    inference_long-bench.py:41-62,103-112; this code never downloads: there is no network in the target deployments
    we test, and a silent download inside a constructor is not something a drop-in should start doing)."""
    if os.path.isdir(path):
        return path
    try:
        from huggingface_hub import snapshot_download
        return snapshot_download(path, local_files_only=True, allow_patterns=["*.safetensors", "*.bin", "*.json"])
    except Exception as e:  # noqa: BLE001
        raise FileNotFoundError(
            f"{path!r} is neither a local checkpoint directory nor a hub id present in the local Hugging Face cache "
            f"({type(e).__name__}: {e}); download it first (huggingface-cli download {path}) or pass a directory") from e


def _load_dir(path, prefix=""):
    path = _resolve_checkpoint_dir(path)
    files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    sd = {}
    if files:
        from safetensors.torch import load_file
        for f in files:
            for k, v in load_file(f).items():
                sd[prefix + k] = v
        return sd
    bins = sorted(glob.glob(os.path.join(path, "pytorch_model*.bin")))
    if not bins:
        raise FileNotFoundError(f"no *.safetensors or pytorch_model*.bin under {path!r}")
    for f in bins:
        for k, v in torch.load(f, map_location="cpu", weights_only=True).items():
            sd[prefix + k] = v
    return sd


def _parse_synth(spec):
    kw = {"seed": 0, "scale": 1.0, "draft_scale": None}
    if ":" in spec:
        for item in spec.split(":", 1)[1].split(","):
            k, v = item.split("=")
            kw[k] = int(v) if k == "seed" else float(v)
    return kw


def _on_device(fn):
    """Run a method with the model's device current: kernels are enqueued on the CURRENT device's stream and
    ls_sm_count / the shared-memory opt-in are per device, so a model built for cuda:1 must not depend on what the
    caller left current."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return wrapped


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
        if self.device.index is None:
            self.device = torch.device(f"cuda:{torch.cuda.current_device()}")
        self.use_graph = os.environ.get("LONGSPEC_CUDA_GRAPH", "1") != "0"
        with torch.cuda.device(self.device):
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
        # the reference reads hidden_states[:, input_len - 1] (llama_glide.py:955): right-padded input_ids are legal
        last = eng.prefill(input_ids[:, :eng.prompt_len])
        return int(F.linear(last, self.weights.lm_head).argmax(-1))

    # -------------------------------------------------------------- vanilla (llama_glide.py:552-585)
    @_on_device
    def vanilla_generate(self, input_ids, prompt_length, max_gen_len=64, eos_id=151645):
        assert input_ids is not None, "please give the input"
        assert input_ids.size(0) == 1, "batch size 1 (inference_long-bench.py:119)"
        P = int(prompt_length.reshape(-1)[0]) if torch.is_tensor(prompt_length) else int(prompt_length)
        eng = self._new_engine(P, max_gen_len, [1])
        out = torch.zeros(max_gen_len, dtype=torch.int64, device=self.device)
        out[0] = self._first_token(eng, input_ids)
        eng.set_lengths(P, P, P, [int(out[0])])
        stop_id = self.eos_token_id
        pos = eng.p_cache_len
        tok_in = out[0:1].clone()
        tok_out = torch.zeros(1, dtype=torch.int64, device=self.device)
        idx = torch.ones(1, dtype=torch.int64, device=self.device)
        # the reference tests `eos in output_ids` after EVERY step (578), first token included
        hit = (out[0:1] == stop_id).to(torch.int32) if stop_id is not None else torch.zeros(1, dtype=torch.int32, device=self.device)

        def step():
            # one token: append at cache_len, causal attention, argmax (llama_glide.py:572-577), all on the device
            eng.descs.clear()
            eng.decode_one(tok_in, pos, tok_out)
            eng.p_cache_len.add_(1)
            out.scatter_(0, idx, tok_out)
            idx.add_(1)
            tok_in.copy_(tok_out)
            if stop_id is not None:
                hit.bitwise_or_((tok_out == stop_id).to(torch.int32))

        graph = None
        if max_gen_len > 1:
            # untimed warm-up on the real state (first launches load the kernels and let cuBLAS pick its algorithms, like
            # the drivers' own warm-up generation, inference_long-bench.py:233-240), graph capture, then state restored
            saved = [t.clone() for t in (eng.st["scalars"], out, idx, tok_in, hit)]
            step()
            if self.use_graph:
                graph = self._capture(step)
            for t, v in zip((eng.st["scalars"], out, idx, tok_in, hit), saved):
                t.copy_(v)
        hosts = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(2)]
        events = [torch.cuda.Event(), torch.cuda.Event()]
        torch.cuda.synchronize()
        t0 = time.time()
        num = 0
        for i in range(1, max_gen_len):
            graph.replay() if graph is not None else step()
            hosts[i % 2].copy_(hit, non_blocking=True)
            events[i % 2].record()
            num += 1
            if stop_id is not None and i >= 2:
                # the flag of step i-1 (already on its way) is read while step i runs: the GPU never waits for the host;
                # the price is at most one surplus step, trimmed below
                events[(i - 1) % 2].synchronize()
                if int(hosts[(i - 1) % 2][0]):
                    num = i - 1
                    break
        torch.cuda.synchronize()
        elapsed = time.time() - t0
        if stop_id is not None:
            found = (out[:num + 1] == stop_id).nonzero()
            if found.numel():
                num = max(int(found[0]), 1) if max_gen_len > 1 else 0    # the loop body runs at least once (572-579)
            out[num + 1:] = 0
        self.last_graph = graph
        return out[None], num, elapsed

    def vanilla_torch_generate(self, input_ids, prompt_length, max_gen_len=64, eos_id=151645):
        """The reference's pure-torch attention baseline (llama_glide.py:587-619) produces the same
        tokens as vanilla_generate; here both run the same kernels."""
        return self.vanilla_generate(input_ids, prompt_length, max_gen_len, eos_id)

    def _capture(self, fn):
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        return g

    # -------------------------------------------------------------- tree (llama_glide.py:915-1126)
    @_on_device
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
        limit = max_gen_len - 1 if max_rounds is None else min(max_gen_len - 1, max_rounds)
        # Untimed warm-up + graph capture BEFORE the timer (the reference's timer starts after its set-up too,
        # llama_glide.py:993, and its drivers run a whole untimed warm-up generation first): one eager round with the
        # device-side stop flag set - acceptance is then a no-op and everything else a round writes is rewritten by the
        # next real round - loads every kernel and lets cuBLAS choose its algorithms; the loop counters it advanced
        # (draft_len) are restored from a copy.
        graph = None
        launches0 = eng.launches
        if limit >= 1:
            saved = (st["scalars"].clone(), st["counters"].clone())
            st["counters"][3] = 1
            eng.round(stop_id)
            self.launches_per_round = eng.launches - launches0
            if self.use_graph:
                graph = self._capture(lambda: eng.round(stop_id))
            st["scalars"].copy_(saved[0])
            st["counters"].copy_(saved[1])
        # Pipelined host loop: round r+1 is enqueued BEFORE round r's 16-byte result is read, so the GPU never waits
        # for the host (graph launch, D2H, Python).  A round enqueued after the stop condition is harmless:
        # ls_tree_accept is a no-op once the device-side stop flag is set.
        hosts = [torch.zeros(4, dtype=torch.int32).pin_memory() for _ in range(2)]
        events = [torch.cuda.Event(), torch.cuda.Event()]
        rounds = 0
        round_log = []                       # (host time, tokens written) after every round's stop test
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
        host.copy_(st["counters"])
        if limit >= 1:
            enqueue()
        while rounds < queued:
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

    # -------------------------------------------------------------- operator seam (llama_glide.py:1128-1175)
    def tree_verification(self, input_ids, output_ids, tree_mask, cache_lens, non_leaf_len, K_Cache=None, V_Cache=None):
        """Reference-shaped greedy tree verification on dense tensors (the decode loop itself calls ls_tree_accept on
        its device-resident state).  See longspec_b200.seams.tree_verification."""
        from .seams import tree_verification
        return tree_verification(input_ids, output_ids, tree_mask, cache_lens, non_leaf_len, K_Cache, V_Cache)

    # -------------------------------------------------------------- other driver methods
    def spec_generate(self, *a, **k):
        raise NotImplementedError("--method seq (llama_glide.py:621-774) is outside the tree-spec hot path")

    def magicdec_generate(self, *a, **k):
        raise NotImplementedError("--method magicdec (llama_glide.py:776-913) is a competitor baseline, out of scope")

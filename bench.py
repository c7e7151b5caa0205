#!/usr/bin/env python
"""Benchmark of the tree-spec decode hot path (BASELINE.json: decode tokens/s at 128k context,
Llama-3-8B dims, tree 4 16 16 16 16, bf16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one decode round (one draft tree + one tree verify, llama_glide.py:997-1121).
N = 1: the 128k-token context lives on one B200.  N > 1 (torchrun): the prefix KV is sharded by
sequence chunk across the ranks (strong scaling: the workload is fixed), per-attention partials are
exchanged with NCCL, every rank runs the identical round.
Rank 0 prints ONE JSON line on stdout; progress goes to stderr.

Timing: W untimed warm-up rounds, then exactly K rounds between CUDA events with a barrier +
synchronize on both sides, max over ranks.  `value` = tokens emitted in those K rounds / time with
everything resident in HBM (K CUDA-graph replays back to back).  `e2e` = the same metric through the
public `LlamaGlide.tree_spec_generate` call with a pinned HOST prompt: the timed rounds include the
per-round device->host read of the stop/accept counters the loop needs, and the final output copy.
L2: every round streams > 30 GB (weights + KV) so no input survives in the 126 MB L2 between rounds.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = os.environ.get("LONGSPEC_BENCH_MODEL", "llama3-8b-262k")
PROMPT = int(os.environ.get("LONGSPEC_BENCH_PROMPT", 131072))
TREE = [int(x) for x in os.environ.get("LONGSPEC_BENCH_TREE", "4 16 16 16 16").split()]
# draft noise 2.6 gives tau ~ 3.3 with the synthetic weights (tools/tau_sweep.py), the accepted length the reference
# publishes for Llama-3-8B (2.99-3.39)
DRAFT_SCALE = os.environ.get("LONGSPEC_BENCH_DRAFT_SCALE", "2.6")
METRIC = "decode_tokens_per_s"
UNIT = "tokens/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


GOLDEN = os.path.join(ROOT, "tests", "golden", "bench_workload.json")


def workload_key():
    return f"{MODEL}|P={PROMPT}|tree={' '.join(map(str, TREE))}|draft_scale={DRAFT_SCALE}|bf16"


def workload_golden():
    """Facts about the SEEDED synthetic workload measured once on one B200 (tools/make_bench_golden.py) and committed:
    the target's greedy continuation (vanilla_generate - what any lossless speculative run must emit) and the accepted
    tokens of every round of the tree-spec generation (the accepted length is a property of the workload at T = 0, not
    of the implementation, so both bench arms use the same one)."""
    if os.path.exists(GOLDEN):
        return json.load(open(GOLDEN)).get(workload_key())
    return None


def golden_tau(W, K, fallback):
    g = workload_golden()
    if g and len(g.get("tokens_per_round", [])) >= W + K:
        return sum(g["tokens_per_round"][W:W + K]) / K, "accepted tokens of rounds [W, W+K) of the seeded workload (tests/golden/bench_workload.json)"
    return fallback, "published tau (BASELINE.md): no committed record for this workload"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 25 ms; stop(t0, t1) keeps the samples that arrived while the
    GPU was running the benchmark rounds (nvidia-smi needs ~1 s to deliver its first line, so it is started well
    before the timed region and the warm-up replays - the identical workload - count as "under load")."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception as e:  # noqa: BLE001
            log("clock sampler unavailable:", e)

    def wait_first(self, timeout=5.0):
        t_end = time.time() + timeout
        while not self.lines and time.time() < t_end:
            time.sleep(0.02)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if (t0 is not None and ts < t0) or (t1 is not None and ts > t1 + 0.03):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None,
                "window": "samples taken while the warm-up and timed graph replays ran (same workload back to back)"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's PyTorch-eager path on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------
class CpuSample:
    """ONE decode round of the CPU oracle at the real widths, context length and tree, with the target
    truncated to `ls` layers (random KV caches stand in for the 128k-token prefill, which no CPU can run
    in minutes).  Full-depth round time = t(1) + (L - 1) * (t(2) - t(1))."""

    def __init__(self, model, prompt, tree, threads):
        from longspec_b200.synthetic import PRESETS
        torch.set_num_threads(threads)
        self.full, self.prompt, self.tree, self.models = PRESETS[model], prompt, tree, {}

    def _model(self, ls):
        if ls not in self.models:
            from dataclasses import replace
            from longspec_b200.synthetic import synthetic_state_dict
            from oracle.model import Dims, OracleModel
            dims = replace(self.full, layers=ls)
            W = synthetic_state_dict(dims, seed=0)
            m = OracleModel(Dims(**{k: getattr(dims, k) for k in Dims.__dataclass_fields__}), W)
            g = torch.Generator().manual_seed(ls)
            S = self.prompt + 512
            for li in range(ls):
                m.K[li] = torch.randn(1, S, dims.kv_heads, 128, generator=g)
                m.V[li] = torch.randn(1, S, dims.kv_heads, 128, generator=g)
            m.dK = torch.randn(1, S + 128, dims.kv_heads, 128, generator=g)
            m.dV = torch.randn(1, S + 128, dims.kv_heads, 128, generator=g)
            self.models[ls] = m
        return self.models[ls]

    def round(self, ls):
        from oracle.generate import RoundState, tree_round
        m = self._model(ls)
        st = RoundState(self.tree, self.prompt, 5)
        t0 = time.perf_counter()
        with torch.inference_mode():
            tree_round(m, st)
        dt = time.perf_counter() - t0
        log(f"[cpu] {ls}-layer round: {dt:.2f} s")
        return dt

    def calibrate(self, deep=3, repeats=3):
        """Per-layer cost of the verify pass from ALTERNATING 1-layer / `deep`-layer rounds, medians of `repeats`
        (one untimed pass of each first: freshly allocated models pay page faults and thread start-up).  The wider
        lever (deep - 1 layers instead of one) and the medians keep the extrapolation stable from run to run."""
        self.round(1)
        self.round(deep)
        t1s, tds = [], []
        for _ in range(repeats):
            t1s.append(self.round(1))
            tds.append(self.round(deep))
        t1, td = sorted(t1s)[len(t1s) // 2], sorted(tds)[len(tds) // 2]
        self.per_layer = max(td - t1, 0.0) / (deep - 1)
        self.models.pop(deep, None)
        self.t1, self.deep, self.td = t1, deep, td
        return t1, td

    def full_round(self, t1):
        return t1 + (self.full.layers - 1) * self.per_layer


def _cpu_sample_text(cs, threads):
    return (f"oracle (CPU port of the reference's PyTorch-eager tree path, fp32, {threads} threads): decode rounds (5 draft steps + "
            f"tree verify) at {MODEL} widths, {PROMPT}-token context, tree {TREE}, random KV standing in for the prefill, with "
            f"the target truncated to 1 and {cs.deep} layers (medians of 3 alternating repeats: {cs.t1:.2f} s, {cs.td:.2f} s); "
            f"full depth = t(1) + {cs.full.layers - 1} x {cs.per_layer:.3f} s per layer")


def cpu_baseline(tau, tau_src, threads):
    cs = CpuSample(MODEL, PROMPT, TREE, threads)
    t1, td = cs.calibrate()
    total = cs.full_round(t1)
    return {"value": tau / total, "unit": UNIT, "cores": threads, "kind": "port",
            "seconds_per_round": total, "accepted_len_tau": tau, "tau_source": tau_src,
            "sample": _cpu_sample_text(cs, threads)}


def run_reference(args, rank):
    """--impl reference: the reference algorithm's CPU implementation (oracle port) on all host threads.
    Each step = one MEASURED full-width, full-context decode round with a 1-layer target (the bounded sample);
    the reported value extrapolates it to full depth with the per-layer cost calibrated in the warm-up."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    W, K = max(args.warmup, 3), args.steps
    tau, tau_src = golden_tau(W, K, float(os.environ.get("LONGSPEC_BENCH_TAU", 3.25)))
    cs = CpuSample(MODEL, PROMPT, TREE, threads)
    cs.calibrate()                                   # the warm-up: 2 untimed + 6 timed rounds
    samples = [cs.round(1) for _ in range(K)]
    med = sorted(samples)[len(samples) // 2]
    sec = cs.full_round(med)
    val = tau / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": args.warmup, "ms_per_step": sum(samples) / len(samples) * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{MODEL} widths, random-init weights, {PROMPT}-token synthetic context, tree "
                                   f"{' '.join(map(str, TREE))}, bsz 1, T=0",
                       "note": "CPU: reference algorithm (oracle port of the PyTorch-eager path), all host threads; "
                               "ms_per_step is the MEASURED bounded sample (a 1-layer round); value = tau / seconds_per_round_full_depth",
                       "accepted_len_tau": tau, "tau_source": tau_src,
                       "seconds_per_round_full_depth": sec},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "seconds_per_round": sec, "sample": _cpu_sample_text(cs, threads)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    W, K = max(args.warmup, 3), args.steps

    import torch.distributed as dist
    from longspec_b200.llama_glide import LlamaGlide
    from longspec_b200.synthetic import PRESETS, synthetic_prompt
    from longspec_b200 import _lib
    import ctypes as C

    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    dims = PRESETS[MODEL]
    gamma = len(TREE)
    spec = "synthetic:seed=0" + (f",draft_scale={DRAFT_SCALE}" if DRAFT_SCALE else "")
    t0 = time.time()
    m = LlamaGlide(dims, spec, None, dtype=torch.bfloat16, device=dev, group=group)
    log(f"[rank {rank}] weights ready in {time.time() - t0:.1f} s")
    ids_host = synthetic_prompt(dims.vocab, PROMPT, seed=0).pin_memory()
    plen = torch.tensor([PROMPT])
    MAX_WARM = 100                                   # upper bound of the device-timed warm-up replays (see below)
    max_gen = (2 * (W + K) + MAX_WARM + 6) * (gamma + 1) + 64

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- (A) end to end through the public API (host prompt, per-round D2H) ----------------
    sampler = ClockSampler(local)
    t0 = time.time()
    barrier()
    out, count, num, elapsed, _ = m.tree_spec_generate(ids_host, plen, tree_shape=TREE, max_gen_len=max_gen,
                                                       eos_id=dims.vocab + 5, max_rounds=W + K)
    out_host = out.cpu()
    barrier()
    log(f"[rank {rank}] generation (prefill + {W + K} rounds) took {time.time() - t0:.1f} s; decode {elapsed:.3f} s")
    launches_per_round = m.last_stats["launches_per_round"]
    rl = m.last_stats["round_log"]
    assert len(rl) == W + K, f"generation stopped early after {len(rl)} rounds"
    e2e_time = rl[-1][0] - rl[W - 1][0]
    e2e_tokens = rl[-1][1] - rl[W - 1][1]
    e2e_t = torch.tensor([e2e_time], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = e2e_tokens / float(e2e_t)

    # ---------------- (B) device-resident: K graph replays between CUDA events ----------------
    eng, graph = m.eng, m.last_graph
    exchange, parallelism = eng.exchange, eng.parallelism
    assert graph is not None, "CUDA graph capture is required for the device-timed number"
    st = eng.st
    st["counters"][3] = 0                    # a set stop flag would turn the replayed rounds' acceptance into no-ops
    sampler.start()
    sampler.wait_first()
    barrier()
    t_load0 = time.time()
    # warm-up: at least W rounds, and at least 0.3 s of the identical workload so that the clock samples describe it
    n_warm = min(max(W, int(0.3 / max(e2e_time / K, 1e-4))), MAX_WARM)
    for _ in range(n_warm):
        graph.replay()
    barrier()
    c0 = st["counters"].clone()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_range = os.environ.get("LONGSPEC_PROFILE_RANGE") == "1"     # ncu --profile-from-start off: only the timed rounds
    if prof_range:
        torch.cuda.cudart().cudaProfilerStart()
    ev0.record()
    for _ in range(K):
        graph.replay()
    ev1.record()
    barrier()
    if prof_range:
        torch.cuda.cudart().cudaProfilerStop()
    clocks = sampler.stop(t_load0, time.time())
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    c1 = st["counters"].clone()
    tokens = int(c1[0] - c0[0]) + int(c1[1] - c0[1])
    value = tokens / (ms * 1e-3)
    tau = tokens / K

    # ---------------- (C) roofline of the dominant kernel: verify attention, per launch ----------------
    d = dims
    G = d.heads // d.kv_heads
    q = torch.randn(eng.q_v, d.heads * 128, device=dev, dtype=torch.float32).to(torch.bfloat16)
    o = torch.empty_like(q)
    descs = [eng._desc(q, o, eng.kv, li, eng.p_cache_len, eng.p_veri_rows, eng.q_v, eng.n_split_verify,
                       n_tail=eng.q_v, tail_mask=st["veri_bits"]) for li in range(d.layers)]
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for a in descs[:4]:
        _lib.call("ls_attn_partial", C.byref(a), stream)
    torch.cuda.synchronize()
    reps = 3
    ev0.record()
    for _ in range(reps):
        for a in descs:                      # a different layer's KV every launch: nothing is reused from L2
            _lib.call("ls_attn_partial", C.byref(a), stream)
    ev1.record()
    torch.cuda.synchronize()
    attn_ms = ev0.elapsed_time(ev1) / (reps * d.layers)
    kv_len = int(eng.p_cache_len)
    sh = eng.shard
    lo, hi = sh.begin, (kv_len if sh.end < 0 else min(sh.end, kv_len))
    local_keys = max(hi - lo, 0) + (eng.q_v if sh.owner else 0)
    alg_bytes = local_keys * 2 * d.kv_heads * 128 * 2 + 2 * eng.q_v * d.heads * 128 * 2
    alg_flops = 4.0 * int(eng.p_veri_rows) * local_keys * 128 * d.heads
    peak, peak_src = peaks()
    achieved = alg_bytes / (attn_ms * 1e-3) / 1e9
    traffic = None                           # dram bytes per launch from the committed ncu --set full capture (same shape only)
    tp = os.path.join(ROOT, "profiles", "attn_tc_traffic.json")
    if os.path.exists(tp) and world == 1 and PROMPT == 131072 and MODEL == "llama3-8b-262k":
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    roofline = {"kernel": "ls_attn_partial (target tree-verify attention, one layer)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "us_per_launch": attn_ms * 1e3, "algorithmic_bytes": alg_bytes,
                "tensor_tflops": alg_flops / (attn_ms * 1e-3) / 1e12,
                "share_of_round": attn_ms * d.layers / (ms / K)}

    # ---------------- token checks: every rank emitted the same tokens, and they are the target's greedy continuation ----------------
    n_tok = int(count) + int(num)
    ranks_agree = None
    if world > 1:
        outs = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(outs, out.contiguous())
        ranks_agree = all(torch.equal(o_, outs[0]) for o_ in outs)
        assert ranks_agree, "ranks disagree on the generated tokens"
    gold = workload_golden()
    tokens_checked, match = 0, None
    if gold:
        ref_t = gold["greedy"]
        tokens_checked = min(n_tok, len(ref_t))
        match = out_host[0, :tokens_checked].tolist() == ref_t[:tokens_checked]
        assert match, ("tree-spec output differs from the committed greedy continuation of this seeded workload "
                       f"(first {tokens_checked} tokens): the run is not lossless")

    def teardown():
        # captured NCCL kernels must go before the communicator: destroy the graphs, drain the device, then the group
        nonlocal m, eng, graph, descs
        m.last_graph = None
        del graph, descs, eng, m
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        teardown()
        return
    # ---------------- (D) CPU baseline beside it (rank 0, N = 1 only) ----------------
    teardown()
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            torch.cuda.empty_cache()
            e2e_tau = e2e_tokens / K
            cpu = cpu_baseline(e2e_tau, "accepted length of this run's e2e rounds", os.cpu_count() or 1)
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{MODEL} widths, random-init bf16 weights, {PROMPT}-token synthetic prompt, tree "
                               f"{' '.join(map(str, TREE))}, bsz 1, T=0",
                   "parallelism": parallelism,
                   "exchange": exchange,
                   "l2": "inputs larger than L2: every round streams all weights and the whole KV cache (> 30 GB)",
                   "accepted_len_tau": tau, "draft_scale": DRAFT_SCALE},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT,
                "h2d_bytes_per_step": ids_host.numel() * 8 / (W + K), "d2h_bytes_per_step": 16 + out_host.numel() * 8 / (W + K),
                "ms_per_step": float(e2e_t) * 1e3 / K, "accepted_len_tau": e2e_tokens / K,
                "note": "LlamaGlide.tree_spec_generate with a pinned host prompt; last K rounds of the call, each with its "
                        "device->host read of the accept/stop counters; prefill excluded like the reference's elapsed_time. "
                        "These are different rounds of the generation than the replayed ones behind `value` (which run later, "
                        "back to back under the power cap), so the accepted length and the clocks differ: compare ms_per_step "
                        "for the host-loop overhead"},
        "gpu_launches": int(launches_per_round * K),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "tokens_generated": n_tok,
        "tokens_checked": tokens_checked,
        "tokens_match_greedy": match,
        "ranks_agree": ranks_agree,
        "tokens_check": ("first tokens_checked emitted tokens == the committed greedy (vanilla_generate, 1 GPU) continuation of "
                         "this seeded workload; ranks_agree = all ranks hold identical output_ids" if gold else
                         "no committed greedy record for this workload (tools/make_bench_golden.py): unverified"),
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

"""Host engine of the B200-native tree-spec decode loop.

PyTorch is used for device memory, streams and the dense GEMMs (cuBLAS); everything the
reference does with flash-attn, Triton, torch.compile'd bookkeeping and eager indexing
 goes through the C ABI in include/longspec_b200.h.

Reference counterparts:
  PagedKV            <- module-owned K_Cache/V_Cache tensors (llama.py:219-222, llama_glide.py:223-226)
  Engine.prefill     <- LlamaModel.forward(exec_type="prefill") + draft prefill (llama_glide.py:954-975)
  Engine.decode_one  <- exec_type="decoding" (llama.py:304-329)             [vanilla anchor]
  Engine.draft_step  <- LlamaGlideDecoderLayer.forward (llama_glide.py:424-468)
  Engine.verify      <- LlamaModel.forward(exec_type="tree_decoding") (llama.py:559-626, 357-421)
  Engine.round       <- one iteration of the loop at llama_glide.py:997-1121
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import AttnDesc, TreeState, LS_PAGE, LS_MASK_WORDS
from .synthetic import ModelDims


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _code(dtype):
    if dtype == torch.float16:
        return _lib.LS_F16
    if dtype == torch.bfloat16:
        return _lib.LS_BF16
    raise ValueError(f"unsupported dtype {dtype}: the kernels take float16 or bfloat16")


@dataclass
class Shard:
    """Sequence shard of the prefix KV owned by this GPU.
    rank r of `world` owns global prefix tokens [begin, end); the last rank also owns every token
    generated after the prompt (end = -1) and the tree-masked tail."""
    rank: int = 0
    world: int = 1
    begin: int = 0
    end: int = -1

    @property
    def owner(self):
        return self.rank == self.world - 1

    @staticmethod
    def make(prompt_len: int, rank: int, world: int) -> "Shard":
        if world == 1:
            return Shard()
        per = ((prompt_len + world - 1) // world + LS_PAGE - 1) // LS_PAGE * LS_PAGE
        begin = min(rank * per, prompt_len)
        end = -1 if rank == world - 1 else min((rank + 1) * per, prompt_len)
        return Shard(rank, world, begin, end)


def split_plan(heads, kv_heads, sms, prompt_len, total_tokens, world, row_counts):
    """Partial slots per attention call for each query-row count in `row_counts`.

    Every rank must plan the SAME number of slots (the exchange all-gathers equal-sized blocks and the combine
    reads world * n_splits of them), so the plan is made for the LARGEST shard - which every rank can compute."""
    G = heads // kv_heads

    def shard_tokens(s):
        n = (total_tokens - s.begin) if s.end < 0 else (s.end - s.begin)
        return (max(n, LS_PAGE) + LS_PAGE - 1) // LS_PAGE * LS_PAGE
    max_local = max(shard_tokens(Shard.make(prompt_len, r, world)) for r in range(world))
    tiles = max((max_local + 63) // 64, 1)

    def splits(rows):
        if rows * G <= 384:
            # tcgen05 path, one CTA per SM.  Up to two row tiles of 128 packed rows: (slot, kv head) CTAs.  Three
            # row tiles: trios of two heavy CTAs (two slots) + one light CTA, i.e. 1.5 CTAs per slot.  Short
            # contexts take fewer slots (>= 4 tiles of 128 keys per share) to keep the partial traffic down.
            tiles128 = (tiles + 1) // 2
            if rows * G <= 256:
                return int(max(1, min(sms // kv_heads, (tiles128 + 3) // 4)))
            groups = max(1, min(sms // (3 * kv_heads), (tiles128 + 7) // 8))
            return int(2 * groups)
        rb = (rows * G + 127) // 128
        return int(max(1, min(tiles, math.ceil(2 * sms / (kv_heads * rb)), 64)))
    return tuple(splits(r) for r in row_counts)


class PagedKV:
    """Paged K/V pools: pool[layer] = (K, V), each [n_pages][Hk][128][128]; one page table."""

    def __init__(self, n_layers, n_kv_heads, max_tokens, dtype, device, shuffle_pages=False, first_token=0):
        self.first_page_token = (first_token // LS_PAGE) * LS_PAGE
        n_tok_pages = (max_tokens + LS_PAGE - 1) // LS_PAGE
        n_live = n_tok_pages - self.first_page_token // LS_PAGE
        self.n_pages = n_live + 1                       # page 0 is a zero page for never-read tokens
        self.k = [torch.zeros(self.n_pages, n_kv_heads, LS_PAGE, 128, dtype=dtype, device=device) for _ in range(n_layers)]
        self.v = [torch.zeros(self.n_pages, n_kv_heads, LS_PAGE, 128, dtype=dtype, device=device) for _ in range(n_layers)]
        live = torch.arange(1, n_live + 1, dtype=torch.int32)
        if shuffle_pages:
            g = torch.Generator().manual_seed(1234)
            live = live[torch.randperm(n_live, generator=g)]
        table = torch.zeros(n_tok_pages, dtype=torch.int32)
        table[self.first_page_token // LS_PAGE:] = live
        self.page_table = table.to(device)
        self.max_tokens = n_tok_pages * LS_PAGE

    def write(self, layer, k, v, start):
        """k, v [n][Hk][128] -> token slots start.. (host-driven; used by prefill only)."""
        n = k.shape[0]
        tok = torch.arange(start, start + n, device=k.device)
        page = self.page_table[(tok // LS_PAGE).long()].long()
        slot = (tok % LS_PAGE).long()
        self.k[layer][page, :, slot] = k
        self.v[layer][page, :, slot] = v

    def read(self, layer, start, n):
        """Gather token rows back as [n][Hk][128] (tests / seam wrappers)."""
        tok = torch.arange(start, start + n, device=self.page_table.device)
        page = self.page_table[(tok // LS_PAGE).long()].long()
        slot = (tok % LS_PAGE).long()
        return self.k[layer][page, :, slot], self.v[layer][page, :, slot]


class Weights:
    """Fused weight tensors built from an HF-named state dict (see synthetic.py for the keys)."""

    def __init__(self, dims: ModelDims, sd: dict, dtype, device):
        self.dims = dims
        cv = lambda t: t.to(device=device, dtype=dtype).contiguous()
        self.embed = cv(sd["model.embed_tokens.weight"])
        # checkpoints with tie_word_embeddings ship no lm_head.weight: the head IS the embedding matrix
        self.lm_head = cv(sd["lm_head.weight"]) if "lm_head.weight" in sd else self.embed
        self.norm = cv(sd["model.norm.weight"])
        self.layers = []
        for i in range(dims.layers):
            p = f"model.layers.{i}."
            L = {}
            L["qkv"] = cv(torch.cat([sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.k_proj.weight"],
                                     sd[p + "self_attn.v_proj.weight"]], 0))
            L["qkv_b"] = cv(torch.cat([sd[p + "self_attn.q_proj.bias"], sd[p + "self_attn.k_proj.bias"],
                                       sd[p + "self_attn.v_proj.bias"]], 0)) if dims.qkv_bias else None
            L["o"] = cv(sd[p + "self_attn.o_proj.weight"])
            L["gu"] = cv(torch.cat([sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"]], 0))
            L["down"] = cv(sd[p + "mlp.down_proj.weight"])
            L["ln1"] = cv(sd[p + "input_layernorm.weight"])
            L["ln2"] = cv(sd[p + "post_attention_layernorm.weight"])
            self.layers.append(L)
        g = "glide."
        D = {}
        D["sa_qkv"] = cv(torch.cat([sd[g + "self_attn.q_proj.weight"], sd[g + "self_attn.k_proj.weight"],
                                    sd[g + "self_attn.v_proj.weight"]], 0))
        D["sa_qkv_b"] = cv(torch.cat([sd[g + "self_attn.q_proj.bias"], sd[g + "self_attn.k_proj.bias"],
                                      sd[g + "self_attn.v_proj.bias"]], 0))
        D["sa_o"] = cv(sd[g + "self_attn.o_proj.weight"])
        # cross attention: only q_proj is live (k/v projections are computed and discarded by the
        # reference, llama_glide.py:248-249,286-287)
        D["ca_q"] = cv(sd[g + "cross_attn.q_proj.weight"])
        D["ca_q_b"] = cv(sd[g + "cross_attn.q_proj.bias"])
        D["ca_o"] = cv(sd[g + "cross_attn.o_proj.weight"])
        D["gu"] = cv(torch.cat([sd[g + "mlp.gate_proj.weight"], sd[g + "mlp.up_proj.weight"]], 0))
        D["down"] = cv(sd[g + "mlp.down_proj.weight"])
        D["ln1"] = cv(sd[g + "input_layernorm.weight"])
        D["ln2"] = cv(sd[g + "post_self_attention_layernorm.weight"])
        D["ln3"] = cv(sd[g + "post_cross_attention_layernorm.weight"])
        self.draft = D
        if dims.inv_freq is not None:           # rope types whose frequencies HF computes for us (llama3, ...)
            self.inv_freq = torch.tensor(dims.inv_freq, dtype=torch.float32, device=device)
        else:
            inv = 1.0 / (dims.rope_theta ** (torch.arange(0, 128, 2, dtype=torch.int64).float() / 128))
            self.inv_freq = (inv / dims.rope_factor).to(device)


def alloc_tree_state(dev, tree_shape, prompt_len, max_gen_len):
    """Allocate the device-resident loop state of one generation and the ls_tree_state view of it."""
    offs = [1]
    for c in tree_shape:
        offs.append(offs[-1] + c)
    gamma, fseq = len(tree_shape), offs[-1]
    q_v = fseq - 1 + gamma + 1
    i32 = lambda n, v=0: torch.full((n,), v, dtype=torch.int32, device=dev)
    st = {}
    st["scalars"] = i32(16)                     # cache_len, llm_kv_len, draft_len, acc_num, step_rows, veri_rows
    st["acc_ids"] = torch.zeros(gamma + 1, dtype=torch.int64, device=dev)
    st["all_spec"] = torch.zeros(fseq, dtype=torch.int64, device=dev)
    st["scores"] = torch.zeros(fseq, dtype=torch.float32, device=dev)
    st["parent"] = i32(fseq)
    st["depth"] = i32(fseq)
    st["tree_bits"] = torch.zeros(fseq * LS_MASK_WORDS, dtype=torch.int64, device=dev)
    n_step = max(gamma + 1, max(tree_shape))
    st["step_ids"] = torch.zeros(n_step, dtype=torch.int64, device=dev)
    st["step_pos"] = i32(n_step)
    st["veri_ids"] = torch.zeros(q_v, dtype=torch.int64, device=dev)
    st["veri_pos"] = i32(q_v)
    st["veri_bits"] = torch.zeros(q_v * LS_MASK_WORDS, dtype=torch.int64, device=dev)
    st["logit_rows"] = i32(fseq)
    st["output_ids"] = torch.zeros(max_gen_len, dtype=torch.int64, device=dev)
    st["counters"] = i32(4)
    st["aux"] = i32(8)
    st["llm_pred"] = torch.zeros(fseq, dtype=torch.int64, device=dev)
    st["one"] = i32(1, 1)
    sc = st["scalars"]
    ts = TreeState()
    ts.cache_len, ts.llm_kv_len, ts.draft_len, ts.acc_num = _ptr(sc[0:1]), _ptr(sc[1:2]), _ptr(sc[2:3]), _ptr(sc[3:4])
    ts.step_rows, ts.veri_rows = _ptr(sc[4:5]), _ptr(sc[5:6])
    ts.acc_ids, ts.all_spec, ts.scores = _ptr(st["acc_ids"]), _ptr(st["all_spec"]), _ptr(st["scores"])
    ts.parent, ts.depth, ts.tree_bits = _ptr(st["parent"]), _ptr(st["depth"]), _ptr(st["tree_bits"])
    ts.step_ids, ts.step_pos = _ptr(st["step_ids"]), _ptr(st["step_pos"])
    ts.veri_ids, ts.veri_pos, ts.veri_bits = _ptr(st["veri_ids"]), _ptr(st["veri_pos"]), _ptr(st["veri_bits"])
    ts.logit_rows, ts.output_ids, ts.counters, ts.aux = _ptr(st["logit_rows"]), _ptr(st["output_ids"]), _ptr(st["counters"]), _ptr(st["aux"])
    ts.fseq, ts.gamma, ts.prompt_len, ts.max_gen_len = fseq, gamma, prompt_len, max_gen_len
    return st, ts


class Engine:
    """One generation's device state + the kernels wired together.  bsz = 1."""

    def __init__(self, weights: Weights, dtype, device, shard: Shard | None = None, group=None,
                 shuffle_pages=False):
        self.w = weights
        self.dims = weights.dims
        self.dtype = dtype
        self.code = _code(dtype)
        self.device = device
        self.shard = shard or Shard()
        self.group = group
        self.shuffle_pages = shuffle_pages
        self.sms = _lib.load().ls_sm_count()
        self.scale = 1.0 / math.sqrt(128)
        self.kv = None
        self.launches = 0           # kernels of THIS library launched (bench reports it)

    # ------------------------------------------------------------------ setup
    def begin(self, prompt_len, max_gen_len, tree_shape):
        d = self.dims
        dev = self.device
        self.tree_shape = list(tree_shape)
        self.offs = [1]
        for c in self.tree_shape:
            self.offs.append(self.offs[-1] + c)
        self.gamma = len(self.tree_shape)
        self.fseq = self.offs[-1]
        self.S = self.fseq - 1
        self.q_v = self.S + self.gamma + 1
        if self.q_v > 64 * LS_MASK_WORDS:
            raise ValueError(f"tree of {self.S} nodes needs {self.q_v} verify rows; the kernels take <= {64 * LS_MASK_WORDS}")
        if max(self.tree_shape) > _lib.LS_MAX_TOPK:
            raise ValueError(f"tree level width {max(self.tree_shape)} exceeds {_lib.LS_MAX_TOPK}")
        self.prompt_len = prompt_len
        self.max_gen_len = max_gen_len
        sh = self.shard
        total = prompt_len + max_gen_len + 256 + self.q_v + LS_PAGE     # llama_glide.py:924 slack
        local_tokens = (total - sh.begin) if sh.end < 0 else (sh.end - sh.begin)
        self.kv = PagedKV(d.layers, d.kv_heads, max(local_tokens, LS_PAGE), self.dtype, dev, self.shuffle_pages)
        # the draft's own cache is only ever read through a 512(+level width) window: keep the tail
        first_needed = max(prompt_len - 512 - 2 * LS_PAGE, 0)
        self.dkv = PagedKV(1, d.kv_heads, total + 128, self.dtype, dev, self.shuffle_pages, first_token=first_needed)
        self.draft_first = self.dkv.first_page_token

        self.st, self.ts = alloc_tree_state(dev, self.tree_shape, prompt_len, max_gen_len)
        st = self.st
        sc = st["scalars"]
        self.p_cache_len, self.p_llm_kv_len, self.p_draft_len, self.p_acc_num, self.p_step_rows, self.p_veri_rows = \
            [sc[i:i + 1] for i in range(6)]
        n_step = st["step_ids"].numel()

        # split-KV plan (static grids; each split's key range adapts to the device-side lengths)
        G = d.heads // d.kv_heads
        plan = split_plan(d.heads, d.kv_heads, self.sms, prompt_len, total, sh.world, (self.q_v, n_step, 1))
        self.n_split_verify, self.n_split_cross, self.n_split_one = plan
        self.n_split_self = 4
        world = sh.world
        max_parts = max(self.n_split_verify, self.n_split_cross, self.n_split_one)
        max_parts = max(max_parts, self.n_split_self)
        rows_max = max(self.q_v, max(1, 256 // G)) * G          # verify rows, or a prefill block of 256 // G query tokens
        self.part_o = torch.empty(max_parts * d.kv_heads * rows_max * 128, dtype=torch.float32, device=dev)
        self.part_lse = torch.empty(max_parts * d.kv_heads * rows_max, dtype=torch.float32, device=dev)
        # multi-GPU exchange blocks: ONE locally reduced (O, lse) partial per rank and attention call
        self.xchg_o = self.xchg_lse = None
        self.peer = None
        if world > 1 and os.environ.get("LONGSPEC_XCHG", "peer") != "nccl":
            # fused exchange over NVLink peer memory; falls back to the NCCL all-gather only if symmetric memory cannot
            # be set up on this system (the reason is kept in self.peer_error and reported by bench.py)
            try:
                from .parallel import get_peer_exchange
                self.peer = get_peer_exchange(d.kv_heads * rows_max, dev, self.group)
            except Exception as e:  # noqa: BLE001
                self.peer_error = f"{type(e).__name__}: {e}"
            # all ranks or none: a rank that fell back alone would wait for collectives the others never issue
            import torch.distributed as dist
            ok = torch.tensor([1 if self.peer is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok) == 0:
                self.peer = None
        self.exchange = "single GPU" if world == 1 else ("peer memory (NVLink stores + flags)" if self.peer is not None else
                                                         "NCCL all-gather" + (f" ({self.peer_error})" if getattr(self, "peer_error", None) else ""))
        if world > 1 and self.peer is None:
            self.xchg_o = torch.empty(world * d.kv_heads * rows_max * 128, dtype=torch.float32, device=dev)
            self.xchg_lse = torch.empty(world * d.kv_heads * rows_max, dtype=torch.float32, device=dev)
            self.send_o = torch.empty(d.kv_heads * rows_max * 128, dtype=torch.float32, device=dev)
            self.send_lse = torch.empty(d.kv_heads * rows_max, dtype=torch.float32, device=dev)
        self.grow_ws = torch.empty(int(_lib.load().ls_grow_workspace_bytes(32, d.vocab)), dtype=torch.uint8, device=dev)
        # Vocab-sharded lm_head (multi-GPU): the 6 lm_head GEMMs of a round read 6 x V x d weights that every rank would
        # otherwise stream in full (6.3 GB of the replicated 22.9 GB per round for Llama-3-8B).  Rank r takes a
        # chunk-aligned slice of the vocabulary (chunk = 2048 ids); the per-chunk reductions are all-gathered.
        self.vshard = None
        if world > 1 and os.environ.get("LONGSPEC_VOCAB_SHARD", "1") != "0":
            n_chunks = (d.vocab + 2047) // 2048
            cpr = (n_chunks + world - 1) // world                      # chunks per rank (the last slices may be short or empty)
            c0 = sh.rank * cpr
            v0, v1 = min(c0 * 2048, d.vocab), min((c0 + cpr) * 2048, d.vocab)
            cb = int(_lib.load().ls_grow_chunk_bytes())
            self.vshard = dict(cpr=cpr, c0=c0, v0=v0, v1=v1, n_chunks=cpr * world, chunk_bytes=cb,
                               send=torch.empty(cpr * cb, dtype=torch.uint8, device=dev),
                               pairs_send=torch.empty(self.fseq * 2, dtype=torch.int64, device=dev),
                               pairs_all=torch.empty(world * self.fseq * 2, dtype=torch.int64, device=dev),
                               head=self.w.lm_head[v0:v1] if v1 > v0 else None,
                               dummy=torch.zeros(8, dtype=self.dtype, device=dev))
        # counters of the fused split-KV combine (ls_attn_desc::fuse_sync); single GPU only - sharded calls merge in the
        # exchange kernels instead.  OFF by default: measured on B200 the fused tail costs what the separate 5 us combine
        # launch costs (12.610 vs 12.603 ms per round at 128k, 6.680 vs 6.657 at 16k, profiles/r02_fused_combine_ab.txt) -
        # the CTAs of a kv head wait for their slowest member at the barrier - so it only makes the attention launch
        # longer.  LONGSPEC_FUSE_COMBINE=1 turns it on.
        self.fuse_sync = torch.zeros(2 * d.kv_heads, dtype=torch.int32, device=dev) \
            if world == 1 and os.environ.get("LONGSPEC_FUSE_COMBINE", "0") == "1" else None
        # Tensor-parallel MLP (multi-GPU decode rounds): the MLP is 70 % of the per-layer weight bytes that every rank would
        # otherwise stream in full.  Rank r keeps rows [r, r+1) * ffn / world of gate and up (column-parallel, no exchange)
        # and the matching columns of down (row-parallel): its fp32 partial [rows, hidden] is pushed to every peer and
        # the sum is rounded once, like the single-GPU GEMM output.  The prefill keeps the replicated weights (untimed).
        self.tp = None
        if (world > 1 and self.peer is not None and os.environ.get("LONGSPEC_TP_MLP", "1") != "0"
                and d.ffn % world == 0 and (d.ffn // world) % 8 == 0):
            key = (world, sh.rank)
            cache = getattr(self.w, "tp_cache", None)
            if cache is None:
                cache = self.w.tp_cache = {}
            if key not in cache:
                f, fl = d.ffn, d.ffn // world
                r0, r1 = sh.rank * fl, (sh.rank + 1) * fl

                def shard(L):
                    return dict(gu=torch.cat([L["gu"][r0:r1], L["gu"][f + r0:f + r1]], 0).contiguous(),
                                down=L["down"][:, r0:r1].contiguous())
                cache[key] = dict(layers=[shard(L) for L in self.w.layers], draft=shard(self.w.draft))
            self.tp = cache[key]
            self.tp_slot = (self.q_v * d.hidden * 4 + 15) // 16 * 16
            if self.tp_slot > self.peer.block_o * 4:
                self.tp = None
        self.parallelism = "single GPU" if world == 1 else (
            f"prefix KV sequence-sharded over {world} GPUs (exchange: {self.exchange})"
            + ("; lm_head vocab-sharded" if self.vshard is not None else "; lm_head replicated")
            + ("; MLP tensor-parallel, fp32 partial sums all-reduced over peer memory" if self.tp is not None else "; MLP replicated")
            + "; attention / qkv / o weights replicated")
        self.descs = []                                # keep ctypes structs alive

    def _mlp(self, x, L, shard):
        """down(silu(gate(x)) * up(x)) -> [rows, hidden] in dtype; tensor-parallel over the ranks in decode rounds."""
        if shard is None:
            return F.linear(self._silu_mul(F.linear(x, L["gu"])), L["down"])
        rows = x.shape[0]
        act = self._silu_mul(F.linear(x, shard["gu"]))
        part = torch.mm(act, shard["down"].t(), out_dtype=torch.float32)          # this rank's fp32 partial sums
        out = torch.empty(rows, self.dims.hidden, dtype=self.dtype, device=self.device)
        nb = rows * self.dims.hidden * 4
        _lib.call("ls_peer_push", _ptr(part), nb, self.tp_slot, self.peer.ref, _stream())
        _lib.call("ls_peer_sum_wait", _ptr(out), rows * self.dims.hidden, self.code, self.tp_slot, self.peer.ref, _stream())
        self.launches += 2
        return out

    # ------------------------------------------------------------------ kernels
    def _desc(self, q, out, pools, layer, kv_len_ptr, n_q_ptr, q_rows, n_splits, *, kv_len_bias=0, n_tail=0,
              tail_mask=None, window=-1, causal=0, sharded=True):
        d = self.dims
        kvp = pools
        sh = self.shard if sharded else Shard()
        a = AttnDesc()
        a.q, a.out = _ptr(q), _ptr(out)
        a.k_pool, a.v_pool, a.page_table = _ptr(kvp.k[layer]), _ptr(kvp.v[layer]), _ptr(kvp.page_table)
        a.kv_len_ptr, a.n_q_ptr = _ptr(kv_len_ptr), _ptr(n_q_ptr)
        a.tail_mask = _ptr(tail_mask)
        a.part_o, a.part_lse = _ptr(self.part_o), _ptr(self.part_lse)
        a.kv_len_bias, a.q_rows, a.n_heads, a.n_kv_heads = kv_len_bias, q_rows, d.heads, d.kv_heads
        a.n_tail, a.window, a.causal, a.dtype = n_tail, window, causal, self.code
        a.n_splits, a.n_pages = n_splits, kvp.n_pages
        a.part_offset, a.n_parts_total = 0, n_splits            # local slots; ranks meet in the exchange blocks (_attn)
        a.kv_shard_begin, a.kv_shard_end, a.tail_owner = sh.begin, sh.end, int(sh.owner)
        a.scale = self.scale
        a.q_stride = q.stride(0)             # the q slice of a fused QKV output is read in place
        a.fuse_sync = _ptr(self.fuse_sync) if (sharded and self.fuse_sync is not None) else C.c_void_p(0)
        return a

    def _attn(self, q, out, pools, layer, kv_len_ptr, n_q_ptr, q_rows, n_splits, **kw):
        a = self._desc(q, out, pools, layer, kv_len_ptr, n_q_ptr, q_rows, n_splits, **kw)
        sharded = kw.get("sharded", True)
        self.descs.append(a)
        _lib.call("ls_attn_partial", C.byref(a), _stream())
        if sharded and self.shard.world > 1 and self.peer is not None:
            # one fused compute + exchange kernel pair: merge the local slots and push the block to every peer, then
            # wait for every peer's block and merge them in rank order
            _lib.call("ls_attn_reduce_push", C.byref(a), self.peer.ref, _stream())
            _lib.call("ls_attn_combine_wait", C.byref(a), self.peer.ref, _stream())
            self.launches += 3
            return
        if sharded and self.shard.world > 1:
            a = self._exchange_partials(a, q_rows)
        elif _lib.load().ls_attn_fuses_combine(C.byref(a)) == 1:
            self.launches += 1               # the attention kernel merged its slots and wrote `out` itself
            return
        _lib.call("ls_attn_combine", C.byref(a), _stream())
        self.launches += 2

    def _gather_small(self, dst, src):
        """All-gather of a small per-rank block (the vocab-sharded lm_head's chunk blocks / argmax pairs): over the peer
        exchange buffer when there is one (no NCCL launch inside the round), else NCCL."""
        if self.peer is not None:
            nb = src.numel() * src.element_size()
            slot = (nb + 15) // 16 * 16
            if nb % 16 == 0 and slot <= self.peer.block_o * 4:
                _lib.call("ls_peer_push", _ptr(src), nb, slot, self.peer.ref, _stream())
                _lib.call("ls_peer_gather_wait", _ptr(dst), nb, slot, self.peer.ref, _stream())
                self.launches += 2
                return
        import torch.distributed as dist
        dist.all_gather_into_tensor(dst, src, group=self.group)

    def _exchange_partials(self, a, q_rows):
        """Sequence-sharded prefix (option A with a deterministic merge order): every rank
        reduces its own split-KV slots to ONE normalised (O, lse) block (ls_attn_reduce), one NCCL all-gather makes
        the `world` blocks visible everywhere, and the returned descriptor lets ls_attn_combine merge them - the
        same data in the same order on every rank, hence bit-identical accept decisions.  Bytes on the wire per call
        and rank: Hk * rows * G * 129 * 4 (1.2 MB at the Llama-3 verify shape)."""
        from .parallel import exchange_partials
        d, sh = self.dims, self.shard
        G = d.heads // d.kv_heads
        n_l = d.kv_heads * q_rows * G
        n_o = n_l * 128
        _lib.call("ls_attn_reduce", C.byref(a), _ptr(self.send_o), _ptr(self.send_lse), _stream())
        self.launches += 1
        exchange_partials(self.xchg_o, self.xchg_lse, n_o, n_l, sh.rank, sh.world, self.group, mine=(self.send_o, self.send_lse))
        b = AttnDesc()
        C.memmove(C.byref(b), C.byref(a), C.sizeof(AttnDesc))
        b.part_o, b.part_lse = _ptr(self.xchg_o), _ptr(self.xchg_lse)
        b.part_offset, b.n_splits, b.n_parts_total = 0, sh.world, sh.world
        self.descs.append(b)
        return b

    def _rope_append(self, q, k, v, q_stride, kv_stride, positions, pools, layer, base_len_ptr, base_bias, slot_offset,
                     n_rows_ptr, rows, append=True):
        d = self.dims
        do_append = append and pools is not None
        _lib.call("ls_rope_append", _ptr(q), _ptr(k), _ptr(v), q_stride, kv_stride, kv_stride, _ptr(positions),
                  _ptr(self.w.inv_freq),
                  _ptr(pools.k[layer]) if do_append else C.c_void_p(0), _ptr(pools.v[layer]) if do_append else C.c_void_p(0),
                  _ptr(pools.page_table) if do_append else C.c_void_p(0), _ptr(base_len_ptr), base_bias, slot_offset,
                  _ptr(n_rows_ptr), rows, d.heads, d.kv_heads, self.code, _stream())
        self.launches += 1

    def _norm(self, h, delta, w, out=None):
        out = torch.empty_like(h) if out is None else out
        _lib.call("ls_add_rmsnorm", _ptr(h), _ptr(delta), _ptr(w), _ptr(out), h.shape[0], h.shape[1],
                  C.c_float(self.dims.rms_eps), self.code, _stream())
        self.launches += 1
        return out

    def _silu_mul(self, gu):
        rows, two_f = gu.shape
        out = torch.empty(rows, two_f // 2, dtype=gu.dtype, device=gu.device)
        _lib.call("ls_silu_mul", _ptr(gu), _ptr(out), rows, two_f // 2, self.code, _stream())
        self.launches += 1
        return out

    def _embed(self, ids, rows):
        out = torch.empty(rows, self.dims.hidden, dtype=self.dtype, device=self.device)
        _lib.call("ls_embed", _ptr(self.w.embed), _ptr(ids), _ptr(out), rows, self.dims.hidden, self.code, _stream())
        self.launches += 1
        return out

    def _gather(self, x, idx, rows):
        out = torch.empty(rows, x.shape[1], dtype=x.dtype, device=x.device)
        _lib.call("ls_gather_rows", _ptr(x), _ptr(idx), _ptr(out), rows, x.shape[1], self.code, _stream())
        self.launches += 1
        return out

    # ------------------------------------------------------------------ target forward (decode shapes)
    def _target_layers(self, h, rows, positions, kv_len_ptr, n_rows_ptr, n_splits, n_tail, tail_mask, causal,
                       kv_len_bias=0):
        """Runs all target layers on `rows` new tokens whose K/V are appended at kv_len.. and attended
        as the tail (tree mask) or causally.  Returns the pre-final-norm residual stream pieces."""
        d = self.dims
        HD, KD = d.heads * 128, d.kv_heads * 128
        sh = self.shard
        delta = None
        for li, L in enumerate(self.w.layers):
            x = self._norm(h, delta, L["ln1"])
            qkv = F.linear(x, L["qkv"], L["qkv_b"])
            q, k, v = qkv[:, :HD], qkv[:, HD:HD + KD], qkv[:, HD + KD:]
            self._rope_append(q, k, v, qkv.stride(0), qkv.stride(0), positions, self.kv if sh.owner else None, li,
                              kv_len_ptr, kv_len_bias - sh.begin, 0, n_rows_ptr, rows)
            attn = torch.empty(rows, HD, dtype=self.dtype, device=self.device)
            self._attn(q, attn, self.kv, li, kv_len_ptr, n_rows_ptr, rows, n_splits, kv_len_bias=kv_len_bias,
                       n_tail=n_tail, tail_mask=tail_mask, causal=causal)
            o = F.linear(attn, L["o"])
            x = self._norm(h, o, L["ln2"])
            delta = self._mlp(x, L, self.tp["layers"][li] if self.tp is not None else None)
        return h, delta

    def verify(self):
        """Target tree-verify forward over veri_ids (llama_glide.py:1086-1091) -> llm_pred [fseq]."""
        st = self.st
        h = self._embed(st["veri_ids"], self.q_v)
        h, delta = self._target_layers(h, self.q_v, st["veri_pos"], self.p_cache_len, self.p_veri_rows,
                                       self.n_split_verify, self.q_v, st["veri_bits"], 0)
        hn = self._norm(h, delta, self.w.norm)
        rows = self._gather(hn, st["logit_rows"], self.fseq)
        vs = self.vshard
        if vs is None:
            logits = F.linear(rows, self.w.lm_head)
            _lib.call("ls_argmax_rows", _ptr(logits), 0, self.code, self.fseq, self.dims.vocab, _ptr(st["llm_pred"]), _stream())
            self.launches += 1
            return st["llm_pred"]
        # vocab-sharded: local (value, global index) per row, all-gather, merge by (value desc, index asc)
        import torch.distributed as dist
        vloc = vs["v1"] - vs["v0"]
        logits = F.linear(rows, vs["head"]) if vloc > 0 else vs["dummy"]
        _lib.call("ls_argmax_rows_partial", _ptr(logits), 0, self.code, self.fseq, vloc, vloc, vs["v0"], _ptr(vs["pairs_send"]), _stream())
        self._gather_small(vs["pairs_all"], vs["pairs_send"])
        _lib.call("ls_argmax_merge", _ptr(vs["pairs_all"]), self.shard.world, self.fseq, _ptr(st["llm_pred"]), _stream())
        self.launches += 2
        return st["llm_pred"]

    def decode_one(self, ids, pos, out_tok):
        """Vanilla decode of one token (llama.py:304-329): append at cache_len, causal attention."""
        h = self._embed(ids, 1)
        # keys = cache[0:cache_len] + the new token as a 1-key tail visible to itself
        h, delta = self._target_layers(h, 1, pos, self.p_cache_len, self.st["one"], self.n_split_one, 1,
                                       self._one_mask(), 0)
        hn = self._norm(h, delta, self.w.norm)
        logits = F.linear(hn, self.w.lm_head)
        _lib.call("ls_argmax_rows", _ptr(logits), 0, self.code, 1, self.dims.vocab, _ptr(out_tok), _stream())
        self.launches += 1

    def _one_mask(self):
        if not hasattr(self, "_one_mask_t"):
            m = torch.zeros(LS_MASK_WORDS, dtype=torch.int64, device=self.device)
            m[0] = 1
            self._one_mask_t = m
        return self._one_mask_t

    # ------------------------------------------------------------------ draft
    def draft_step(self, level):
        """One GliDe draft forward + tree growth.  level 0 = causal step over the accepted tokens
        (llama_glide.py:1003-1027); level >= 1 = tree step over the previous level's nodes (1029-1075)."""
        d, D, st = self.dims, self.w.draft, self.st
        HD, KD = d.heads * 128, d.kv_heads * 128
        if level == 0:
            rows, slot_off, n_tail, mask, causal = self.gamma + 1, 0, 0, None, 1
            self_kv_ptr, self_bias = st["aux"][0:1], 0           # draft_len + acc_num (after the append)
        else:
            lo, hi = self.offs[level - 1], self.offs[level]
            rows, slot_off, n_tail, causal = hi - lo, lo, hi, 0
            mask = st["tree_bits"][lo * LS_MASK_WORDS:]
            self_kv_ptr, self_bias = self.p_draft_len, 0         # prefix excludes the root slot (Q3)
        h = self._embed(st["step_ids"], rows)
        # --- self attention over the draft's own windowed cache
        x = self._norm(h, None, D["ln1"])
        qkv = F.linear(x, D["sa_qkv"], D["sa_qkv_b"])
        q, k, v = qkv[:, :HD], qkv[:, HD:HD + KD], qkv[:, HD + KD:]
        self._rope_append(q, k, v, qkv.stride(0), qkv.stride(0), st["step_pos"], self.dkv, 0, self.p_draft_len,
                          0, slot_off, self.p_step_rows, rows)
        attn = torch.empty(rows, HD, dtype=self.dtype, device=self.device)
        self._attn(q, attn, self.dkv, 0, self_kv_ptr, self.p_step_rows, rows, self.n_split_self,
                   kv_len_bias=self_bias, n_tail=n_tail, tail_mask=mask, window=512, causal=causal, sharded=False)
        o = F.linear(attn, D["sa_o"])
        # --- cross attention over the target's last-layer KV
        x = self._norm(h, o, D["ln2"])
        q2 = F.linear(x, D["ca_q"], D["ca_q_b"])
        self._rope_append(q2, None, None, q2.stride(0), 0, st["step_pos"], None, 0, self.p_draft_len, 0, 0,
                          self.p_step_rows, rows, append=False)
        attn2 = torch.empty(rows, HD, dtype=self.dtype, device=self.device)
        self._attn(q2, attn2, self.kv, d.layers - 1, self.p_llm_kv_len, self.p_step_rows, rows, self.n_split_cross,
                   causal=causal)
        o2 = F.linear(attn2, D["ca_o"])
        # --- FFN
        x = self._norm(h, o2, D["ln3"])
        m = self._mlp(x, D, self.tp["draft"] if self.tp is not None else None)
        h = h + m
        # --- lm_head + fused log-softmax / cumulative score / top-k / mask bookkeeping
        if level == 0:
            h = self._gather(h, st["aux"][1:2], 1)
        lo, hi = (0, 1) if level == 0 else (self.offs[level - 1], self.offs[level])
        vs = self.vshard
        if vs is None:
            logits = F.linear(h, self.w.lm_head)
            _lib.call("ls_tree_grow", C.byref(self.ts), _ptr(logits), 0, self.code, d.vocab, level, lo, hi,
                      self.tree_shape[level], _ptr(self.grow_ws), _stream())
            self.launches += 2
            return
        # vocab-sharded: per-chunk (max, sum exp, top-k) of this rank's slice, all-gather of the chunk blocks, then the
        # same merges + tree update on every rank (identical inputs -> identical trees)
        import torch.distributed as dist
        vloc = vs["v1"] - vs["v0"]
        n_rows = h.shape[0]
        logits = F.linear(h, vs["head"]) if vloc > 0 else vs["dummy"]
        _lib.call("ls_tree_grow_partial", _ptr(logits), 0, self.code, n_rows, vloc, d.vocab, vs["c0"], vs["cpr"],
                  self.tree_shape[level], _ptr(vs["send"]), _stream())
        self._gather_small(self.grow_ws[:vs["n_chunks"] * vs["chunk_bytes"]], vs["send"])
        _lib.call("ls_tree_grow_finish", C.byref(self.ts), d.vocab, vs["n_chunks"], level, lo, hi, self.tree_shape[level],
                  _ptr(self.grow_ws), _stream())
        self.launches += 2

    # ------------------------------------------------------------------ one decode round
    def round(self, stop_id=-1):
        """llama_glide.py:997-1121 for T = 0; everything stays on the device."""
        self.descs.clear()
        _lib.call("ls_round_begin", C.byref(self.ts), _stream())
        for level in range(self.gamma):
            self.draft_step(level)
        _lib.call("ls_build_verify", C.byref(self.ts), _stream())
        pred = self.verify()
        L = self.dims.layers - 1
        own = self.shard.owner
        _lib.call("ls_tree_accept", C.byref(self.ts), _ptr(pred), _ptr(self.kv.k[L]) if own else C.c_void_p(0),
                  _ptr(self.kv.v[L]) if own else C.c_void_p(0), _ptr(self.kv.page_table), self.dims.kv_heads, self.code,
                  self.offs[-2], stop_id, -self.shard.begin, _stream())
        self.launches += 3

    # ------------------------------------------------------------------ prefill (state set-up, untimed)
    def _prefill_splits(self, end):
        """Split-KV slots for a prefill block whose causal context ends at token `end` (this rank's share of it)."""
        sh, d = self.shard, self.dims
        hi = end if sh.end < 0 else min(sh.end, end)
        tiles128 = max((max(hi - sh.begin, 0) + 127) // 128, 1)
        return int(max(1, min(self.sms // d.kv_heads, (tiles128 + 3) // 4)))

    def prefill(self, input_ids, chunk=8192):
        """Target + draft prefill on the package's own kernels (llama.py:199-226, llama_glide.py:206-233,954-975; state
        set-up before the reference's timer at llama_glide.py:993).  The O(P^2) causal attention runs on the SAME
        tcgen05 split-KV kernel as the decode loop: the layer's K/V go into the paged pool first, then every block of
        256 // G query tokens attends to the pool with `causal` set and kv_len = the block's end.  With the prefix KV
        sharded over several GPUs each rank attends only to its own shard and the per-block partials are merged
        through the same exchange as in decoding - the quadratic part of the prefill is divided by the world size
        (the GEMMs stay replicated, like the weights).  Returns the last position's final hidden state [1, hidden]."""
        d, w, sh = self.dims, self.w, self.shard
        P = self.prompt_len                      # KV sizing, cache_len and the shard plan were made for prompt_length
        if input_ids.shape[1] < P:
            raise ValueError(f"input_ids holds {input_ids.shape[1]} tokens but prompt_length is {P}")
        ids = input_ids[0, :P].to(self.device)
        pos = torch.arange(P, device=self.device, dtype=torch.int32)
        HD, KD = d.heads * 128, d.kv_heads * 128
        G = d.heads // d.kv_heads
        h = w.embed[ids]
        one = torch.tensor([P], dtype=torch.int32, device=self.device)
        zero = torch.zeros(1, dtype=torch.int32, device=self.device)
        # draft cache: K/V of the prompt tail (single layer, depends on embeddings only)
        t0 = self.draft_first
        xd = self._norm(h[t0:].clone(), None, w.draft["ln1"])
        qkv = F.linear(xd, w.draft["sa_qkv"], w.draft["sa_qkv_b"])
        nrows = torch.tensor([P - t0], dtype=torch.int32, device=self.device)
        self._rope_append(qkv[:, :HD], qkv[:, HD:HD + KD], qkv[:, HD + KD:], qkv.stride(0), qkv.stride(0),
                          pos[t0:].contiguous(), self.dkv, 0, zero, t0, 0, nrows, P - t0)
        del xd, qkv
        lo = sh.begin
        hi = P if sh.end < 0 else min(sh.end, P)
        Bq = max(1, 256 // G)                                        # query tokens per attention call (<= 256 packed rows)
        n_blocks = (P + Bq - 1) // Bq
        P_pad = n_blocks * Bq
        ends = torch.tensor([min((b + 1) * Bq, P) for b in range(n_blocks)], dtype=torch.int32, device=self.device)
        n_valid = torch.tensor([min(Bq, P - b * Bq) for b in range(n_blocks)], dtype=torch.int32, device=self.device)
        for li, L in enumerate(w.layers):
            q = torch.zeros(P_pad, HD, dtype=self.dtype, device=self.device)
            k = torch.empty(P, KD, dtype=self.dtype, device=self.device)
            v = torch.empty(P, KD, dtype=self.dtype, device=self.device)
            for s in range(0, P, chunk):
                e = min(s + chunk, P)
                x = self._norm(h[s:e], None, L["ln1"])
                qkv = F.linear(x, L["qkv"], L["qkv_b"])
                self._rope_append(qkv[:, :HD], qkv[:, HD:HD + KD], None, qkv.stride(0), qkv.stride(0),
                                  pos[s:e], None, 0, zero, 0, 0, one, e - s, append=False)
                q[s:e], k[s:e], v[s:e] = qkv[:, :HD], qkv[:, HD:HD + KD], qkv[:, HD + KD:]
            if hi > lo:
                self.kv.write(li, k[lo:hi].view(-1, d.kv_heads, 128), v[lo:hi].view(-1, d.kv_heads, 128), 0)
            del k, v
            a = torch.empty(P_pad, HD, dtype=self.dtype, device=self.device)
            for b in range(n_blocks):
                s0 = b * Bq
                self._attn(q[s0:s0 + Bq], a[s0:s0 + Bq], self.kv, li, ends[b:b + 1], n_valid[b:b + 1], Bq,
                           self._prefill_splits(min(s0 + Bq, P)), causal=1)
            self.descs.clear()
            del q
            for s in range(0, P, chunk):
                e = min(s + chunk, P)
                hs = h[s:e]
                o = F.linear(a[s:e], L["o"])
                x = self._norm(hs, o, L["ln2"])
                hs += F.linear(self._silu_mul(F.linear(x, L["gu"])), L["down"])
            del a
        last = self._norm(h[P - 1:P].clone(), None, w.norm)
        return last

    # ------------------------------------------------------------------ state helpers
    def set_lengths(self, cache_len, llm_kv_len, draft_len, acc_ids):
        st = self.st
        vals = torch.tensor([cache_len, llm_kv_len, draft_len, len(acc_ids), 0, 0], dtype=torch.int32)
        st["scalars"][:6] = vals.to(self.device)
        st["acc_ids"][:len(acc_ids)] = torch.tensor(acc_ids, dtype=torch.int64).to(self.device)

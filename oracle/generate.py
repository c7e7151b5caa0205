"""Oracle restatement of the generation loops (bsz = 1).

TEST INFRASTRUCTURE (see oracle/__init__.py).
  vanilla_generate     <- llama_glide.py:552-585
  tree_spec_generate   <- llama_glide.py:915-1126 (T = 0 branch, 1103-1108)
Qwen2 twin differences (qwen2_glide.py:766,949): output_ids pre-filled with 0 instead
of eos_id and termination tests the `eos_id` argument instead of config.eos_token_id.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import tree as T
from .model import OracleModel


def vanilla_generate(m: OracleModel, input_ids, prompt_length, max_gen_len=64, eos_id=151645,
                     config_eos=None):
    P = int(prompt_length)
    out = np.zeros(max_gen_len, dtype=np.int64)
    m.max_len = max_gen_len
    h = m.target_forward(input_ids, "prefill")
    out[0] = int(m.lm_head(h[0, P - 1]).argmax())
    cache_len = P
    num = 0
    t0 = time.perf_counter()
    for _ in range(1, max_gen_len):
        tok = torch.tensor([[out[cache_len - P]]], dtype=torch.int64)
        h = m.target_forward(tok, "decoding", cache_lens=torch.tensor([cache_len]))
        nxt = int(m.lm_head(h[0, -1]).argmax())
        cache_len += 1
        num += 1
        out[cache_len - P] = nxt
        if config_eos is not None and (out == config_eos).any():
            break
    return out, num, time.perf_counter() - t0


class RoundState:
    """Loop variables of tree_spec_generate (llama_glide.py:940-991) for bsz = 1."""

    def __init__(self, shape, prompt_len, first_token):
        self.shape = list(shape)
        self.offs = T.level_offsets(self.shape)
        self.gamma, self.fseq = len(self.shape), self.offs[-1]
        self.cache_len = self.tlen_for_draft = self.draft_len = prompt_len
        self.acc_ids = np.array([first_token], dtype=np.int64)
        self.acc_num = 1
        self.all_spec = np.zeros(self.fseq, dtype=np.int64)
        self.all_spec[0] = first_token
        self.tree_mask = T.fresh_tree_mask(self.fseq)


def tree_round(m: OracleModel, st: RoundState, trace=None):
    """One iteration of the loop at llama_glide.py:997-1116 (T = 0).  Returns the accepted ids."""
    shape, offs, gamma, fseq = st.shape, st.offs, st.gamma, st.fseq
    S = fseq - 1
    prev_acc = st.acc_num
    scores = np.zeros(fseq, dtype=np.float32)
    # ---- draft step 0: causal over accepted tokens (1003-1027)
    ids = torch.from_numpy(st.acc_ids)[None, :]
    pos = (torch.arange(prev_acc) + st.draft_len)[None, :]
    hd = m.draft_forward(m.embed(ids), pos, "decoding", torch.tensor([st.draft_len]),
                         torch.tensor([st.tlen_for_draft]))
    st.draft_len += prev_acc - 1                                       # 1018
    logp = m.lm_head(hd[0, prev_acc - 1]).float().log_softmax(-1).numpy()
    T.grow_root_level(st.tree_mask, st.all_spec, scores, logp, shape[0])
    # ---- draft steps 1..gamma-1 (1029-1075)
    for lv in range(1, gamma):
        lo, hi = offs[lv - 1], offs[lv]
        cur_mask = st.tree_mask[lo:hi, :hi]
        ids = torch.from_numpy(st.all_spec[lo:hi])[None, :]
        pos = torch.from_numpy(T.depth_positions(cur_mask, st.draft_len))[None, :]
        hd = m.draft_forward(m.embed(ids), pos, "tree_decoding", torch.tensor([st.draft_len]),
                             torch.tensor([st.tlen_for_draft]),
                             tree_mask=torch.from_numpy(cur_mask.copy())[None])
        logp = m.lm_head(hd[0]).float().log_softmax(-1).numpy()
        T.grow_level(st.tree_mask, st.all_spec, scores, logp, lv, offs)
    st.draft_len += 1                                                  # 1077
    # ---- verify (1078-1091)
    veri, vmask = T.build_verify_inputs(st.acc_ids, st.all_spec, st.tree_mask, gamma)
    hv = m.target_forward(torch.from_numpy(veri)[None], "tree_decoding",
                          cache_lens=torch.tensor([st.cache_len]),
                          tree_mask=torch.from_numpy(vmask)[None])
    rows = hv[0, prev_acc - 1:prev_acc + S]
    llm_pred = m.lm_head(rows).argmax(-1).numpy().astype(np.int64)
    # ---- greedy verification + last-layer KV compaction (1104-1108, 1128-1175)
    st.cache_len += prev_acc - 1
    acc_ids, acc_num, double_input, idx_map = T.tree_verification(st.all_spec, llm_pred, st.tree_mask, offs[-2])
    L = m.d.layers - 1
    T.compact_rows(m.K[L][0], st.cache_len, fseq, idx_map)
    T.compact_rows(m.V[L][0], st.cache_len, fseq, idx_map)
    st.cache_len += 1
    if trace is not None:
        trace.append(dict(all_spec=st.all_spec.copy(), tree_mask=st.tree_mask.copy(), veri_spec=veri,
                          llm_pred=llm_pred, acc_ids=acc_ids.copy(), acc_num=acc_num,
                          cache_len=st.cache_len, scores=scores.copy()))
    st.tlen_for_draft += acc_num                                       # 1110
    st.acc_ids, st.acc_num = acc_ids, acc_num
    st.tree_mask = T.fresh_tree_mask(fseq)                             # 1113-1116
    st.all_spec[:] = 0
    st.all_spec[0] = acc_ids[acc_num - 1]
    return acc_ids


def tree_spec_generate(m: OracleModel, input_ids, prompt_length, tree_shape=None, max_gen_len=64,
                       eos_id=151645, config_eos=None, qwen_variant=False, trace=None,
                       max_rounds=None):
    """Returns (output_ids, count, num, elapsed, spec_mask) like llama_glide.py:1126.
    `trace` (a list) receives one dict per round for parity diagnostics."""
    shape = [4, 16, 16, 16, 16] if tree_shape is None else list(tree_shape)
    P = int(prompt_length)
    m.max_len = max_gen_len + 256                                        # 924-925
    m.draft_max_len = max_gen_len + 256
    out = np.zeros(max_gen_len, dtype=np.int64) if qwen_variant else \
        np.full(max_gen_len, eos_id, dtype=np.int64)                     # 937 / qwen2_glide.py:766
    spec_mask = np.zeros(max_gen_len, dtype=np.int64)
    stop_id = eos_id if qwen_variant else config_eos
    # prefill target (954-960)
    h = m.target_forward(input_ids, "prefill")
    out[0] = int(m.lm_head(h[0, P - 1]).argmax())
    num, count = 1, 0
    # prefill draft (968-975)
    pos = torch.arange(input_ids.shape[1])[None, :]
    m.draft_forward(m.embed(input_ids), pos, "prefill", torch.tensor([P]), torch.tensor([P]))
    st = RoundState(shape, P, int(out[0]))
    t0 = time.perf_counter()
    rounds = 0
    for _ in range(1, max_gen_len):
        acc_ids = tree_round(m, st, trace)
        base = st.cache_len - P
        out[base:base + st.acc_num] = acc_ids                             # 1108
        count += st.acc_num - 1
        num += 1
        rounds += 1
        if st.cache_len + st.acc_num - P + st.gamma + 2 > max_gen_len:    # 1118
            break
        if stop_id is not None and (out == stop_id).any():                # 1120
            break
        if max_rounds is not None and rounds >= max_rounds:
            break
    return out, count, num, time.perf_counter() - t0, spec_mask

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


def tree_spec_generate(m: OracleModel, input_ids, prompt_length, tree_shape=None, max_gen_len=64,
                       eos_id=151645, config_eos=None, qwen_variant=False, trace=None,
                       max_rounds=None):
    """Returns (output_ids, count, num, elapsed, spec_mask) like llama_glide.py:1126.
    `trace` (a list) receives one dict per round for parity diagnostics."""
    shape = [4, 16, 16, 16, 16] if tree_shape is None else list(tree_shape)
    offs = T.level_offsets(shape)
    gamma, fseq = len(shape), offs[-1]
    S = fseq - 1
    P = int(prompt_length)
    m.max_len = max_gen_len + 256                                        # 924-925
    m.draft_max_len = max_gen_len + 256 if not qwen_variant else max_gen_len + 256
    out = np.zeros(max_gen_len, dtype=np.int64) if qwen_variant else \
        np.full(max_gen_len, eos_id, dtype=np.int64)                     # 937 / qwen2_glide.py:766
    spec_mask = np.zeros(max_gen_len, dtype=np.int64)
    stop_id = eos_id if qwen_variant else config_eos

    # prefill target (954-960)
    h = m.target_forward(input_ids, "prefill")
    out[0] = int(m.lm_head(h[0, P - 1]).argmax())
    num, count = 1, 0
    cache_len = tlen_for_draft = draft_len = P
    all_spec = np.zeros(fseq, dtype=np.int64)
    all_spec[0] = out[0]
    # prefill draft (968-975)
    pos = torch.arange(input_ids.shape[1])[None, :]
    m.draft_forward(m.embed(input_ids), pos, "prefill", torch.tensor([draft_len]),
                    torch.tensor([tlen_for_draft]))
    acc_ids = np.array([out[0]], dtype=np.int64)
    acc_num = 1
    tree_mask = T.fresh_tree_mask(fseq)
    t0 = time.perf_counter()
    rounds = 0
    for _ in range(1, max_gen_len):
        prev_acc = acc_num
        scores = np.zeros(fseq, dtype=np.float32)
        # ---- draft step 0: causal over accepted tokens (1003-1027)
        ids = torch.from_numpy(acc_ids)[None, :]
        pos = (torch.arange(acc_num) + draft_len)[None, :]
        hd = m.draft_forward(m.embed(ids), pos, "decoding", torch.tensor([draft_len]),
                             torch.tensor([tlen_for_draft]))
        draft_len += prev_acc - 1                                          # 1018
        logp = m.lm_head(hd[0, prev_acc - 1]).float().log_softmax(-1).numpy()
        T.grow_root_level(tree_mask, all_spec, scores, logp, shape[0])
        # ---- draft steps 1..gamma-1 (1029-1075)
        for lv in range(1, gamma):
            lo, hi = offs[lv - 1], offs[lv]
            cur_mask = tree_mask[lo:hi, :hi]
            ids = torch.from_numpy(all_spec[lo:hi])[None, :]
            pos = torch.from_numpy(T.depth_positions(cur_mask, draft_len))[None, :]
            hd = m.draft_forward(m.embed(ids), pos, "tree_decoding", torch.tensor([draft_len]),
                                 torch.tensor([tlen_for_draft]),
                                 tree_mask=torch.from_numpy(cur_mask.copy())[None])
            logp = m.lm_head(hd[0]).float().log_softmax(-1).numpy()
            T.grow_level(tree_mask, all_spec, scores, logp, lv, offs)
        draft_len += 1                                                     # 1077
        # ---- verify (1078-1091)
        veri, vmask = T.build_verify_inputs(acc_ids, all_spec, tree_mask, gamma)
        hv = m.target_forward(torch.from_numpy(veri)[None], "tree_decoding",
                              cache_lens=torch.tensor([cache_len]),
                              tree_mask=torch.from_numpy(vmask)[None])
        rows = hv[0, prev_acc - 1:prev_acc + S]
        llm_pred = m.lm_head(rows).argmax(-1).numpy().astype(np.int64)
        # ---- greedy verification + last-layer KV compaction (1104-1108, 1128-1175)
        cache_len += prev_acc - 1
        acc_ids, acc_num, double_input, idx_map = T.tree_verification(all_spec, llm_pred, tree_mask, offs[-2])
        L = m.d.layers - 1
        T.compact_rows(m.K[L][0], cache_len, fseq, idx_map)
        T.compact_rows(m.V[L][0], cache_len, fseq, idx_map)
        cache_len += 1
        base = cache_len - P
        out[base:base + acc_num] = acc_ids
        if trace is not None:
            trace.append(dict(all_spec=all_spec.copy(), tree_mask=tree_mask.copy(), veri_spec=veri,
                              llm_pred=llm_pred, acc_ids=acc_ids.copy(), acc_num=acc_num,
                              cache_len=cache_len, scores=scores.copy()))
        tlen_for_draft += acc_num                                          # 1110
        count += acc_num - 1
        num += 1
        rounds += 1
        tree_mask = T.fresh_tree_mask(fseq)
        all_spec[:] = 0
        all_spec[0] = acc_ids[acc_num - 1]
        if cache_len + acc_num - P + gamma + 2 > max_gen_len:             # 1118
            break
        if stop_id is not None and (out == stop_id).any():                # 1120
            break
        if max_rounds is not None and rounds >= max_rounds:
            break
    return out, count, num, time.perf_counter() - t0, spec_mask

"""Oracle restatement of the integer tree bookkeeping of `tree_spec_generate`.

TEST INFRASTRUCTURE (see oracle/__init__.py).  numpy int64, explicit loops; sizes
are tiny (<= 1024 tree tokens, llama.py:223).  Bit-exact is the bar for everything
in this file.
"""
from __future__ import annotations

import numpy as np


def level_offsets(tree_shape):
    """acc_num_per_step of llama_glide.py:932-934: [1, 1+c1, 1+c1+c2, ...]."""
    offs = [1]
    for c in tree_shape:
        offs.append(offs[-1] + int(c))
    return offs


def fresh_tree_mask(fseq: int) -> np.ndarray:
    """llama_glide.py:984-985 / 1113-1114: all zero, column 0 (root) = 1."""
    m = np.zeros((fseq, fseq), dtype=np.int64)
    m[:, 0] = 1
    return m


def topk_desc(values: np.ndarray, k: int):
    """Top-k of a flat fp32 array, sorted descending (torch.topk(sorted=True),
    llama_glide.py:1020,1064).  torch leaves the order among exact ties
    unspecified; the oracle AND the CUDA kernel fix it to
    "lower flat index first", and parity tests that need bit-exact indices use
    tie-free inputs or check the tie rule explicitly."""
    v = np.asarray(values, dtype=np.float32).ravel()
    order = np.lexsort((np.arange(v.size), -v.astype(np.float64)))
    idx = order[:k].astype(np.int64)
    return v[idx], idx


def grow_root_level(tree_mask, all_spec, scores, logp_root, c1):
    """Draft step 0 (llama_glide.py:1019-1027): children of the root are the top-c1
    tokens of log_softmax(lm_head(h_root)); their mask rows get the diagonal."""
    top_v, top_i = topk_desc(logp_root, c1)
    for r in range(c1):
        tree_mask[1 + r, 1 + r] += 1
    all_spec[1:1 + c1] = top_i
    scores[1:1 + c1] = top_v
    return top_v, top_i


def grow_level(tree_mask, all_spec, scores, logp, level, offs):
    """Draft step `level` >= 1 (llama_glide.py:1046-1075), "beam tree" variant.

    logp (M, V) fp32 = log_softmax of the previous level's M nodes; the new level's
    c nodes are the global top-c of logp + scores[parent] over the flattened M*V
    (1047,1064); parent = flat // V + offs[level-1], token = flat % V (1066-1067);
    new mask row = parent's row + one-hot (1070-1071)."""
    M, V = logp.shape
    lo, hi, nxt = offs[level - 1], offs[level], offs[level + 1]
    assert hi - lo == M
    cum = (logp.astype(np.float32) + scores[lo:hi, None].astype(np.float32)).astype(np.float32)
    top_v, flat = topk_desc(cum, nxt - hi)
    parents = flat // V + lo
    tokens = flat % V
    for r, (p, t) in enumerate(zip(parents, tokens)):
        row = hi + r
        tree_mask[row] = tree_mask[p]
        tree_mask[row, row] += 1
    all_spec[hi:nxt] = tokens
    scores[hi:nxt] = top_v
    return parents, tokens, top_v


def depth_positions(mask_rows: np.ndarray, base: int) -> np.ndarray:
    """position_ids = mask.sum(-1) - 1 + base (llama.py:575-577, llama_glide.py:1032)."""
    return mask_rows.sum(axis=-1) - 1 + base


def build_verify_inputs(acc_ids, all_spec, tree_mask, gamma):
    """Verify-input assembly (llama_glide.py:1078-1085), bsz = 1.

    veri_spec (S+gamma+1,) = [acc_ids | all_spec[1:] | 0-pad];
    new_tree_mask = tril(ones) with tree_mask[1:,1:] pasted at [acc_num, acc_num+S)^2.
    """
    S = all_spec.shape[0] - 1
    a = len(acc_ids)
    qv = S + gamma + 1
    veri = np.zeros(qv, dtype=np.int64)
    veri[:a] = acc_ids
    veri[a:a + S] = all_spec[1:]
    m = np.ones((qv, qv), dtype=np.int64)
    m[a:a + S, a:a + S] = tree_mask[1:, 1:]
    m = np.tril(m)
    return veri, m


def parent_index(tree_mask: np.ndarray) -> np.ndarray:
    """father_index = argmax_j((mask - I) * j) (llama_glide.py:1136): the largest
    index ancestor, i.e. the parent; 0 for the root and for depth-1 nodes."""
    f = tree_mask.shape[0]
    w = (tree_mask - np.eye(f, dtype=np.int64)) * np.arange(f, dtype=np.int64)[None, :]
    return w.argmax(axis=-1)


def tree_verification(all_spec, llm_pred, tree_mask, non_leaf_len):
    """Greedy tree verification (llama_glide.py:1128-1156), bsz = 1.

    Returns (acc_ids, acc_num, double_input, index_mapping).  index_mapping are the
    selected tree indices in ascending order = the rows whose last-layer K/V get
    compacted to cache_lens + [0, acc_num) (1159-1173, see `compact_rows`).
    """
    f = tree_mask.shape[0]
    father = parent_index(tree_mask)
    verify = (llm_pred[father] == all_spec)
    verify[0] = True
    masked = tree_mask * verify[None, :].astype(np.int64)
    final = masked.sum(-1) == tree_mask.sum(-1)
    last = int((final.astype(np.int64) * np.arange(f)).argmax())
    double_input = int(last >= non_leaf_len)
    select = tree_mask[last]
    acc_num = int(select.sum())
    # argsort(select * [-1024, -1023, ...])[:acc_num]  (1153-1154): selected indices ascending
    key = select * np.arange(-1024, -1024 + f, dtype=np.int64)
    index_mapping = np.argsort(key, kind="stable")[:acc_num]
    acc_ids = llm_pred[index_mapping]
    return acc_ids.astype(np.int64), acc_num, double_input, index_mapping.astype(np.int64)


def compact_rows(cache, cache_len, fseq, index_mapping):
    """Last-layer KV compaction (llama_glide.py:1159-1173): gather rows
    cache_len + index_mapping (read through a copy of the fseq-row window, 1170-1171)
    and write them to cache_len + [0, acc_num).  `cache` is (S_max, Hk, D), mutated."""
    window = cache[cache_len:cache_len + fseq].clone() if hasattr(cache, "clone") \
        else cache[cache_len:cache_len + fseq].copy()
    for dst, src in enumerate(index_mapping):
        cache[cache_len + dst] = window[int(src)]

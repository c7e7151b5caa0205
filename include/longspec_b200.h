/*
 * longspec_b200 - C ABI of the B200-native LongSpec tree-spec decode kernels.
 *
 * The reference (sail-sg/LongSpec) has no FFI of its own: its kernels are reached
 * through Python call sites into flash-attn / Triton / torch.compile.  Every entry
 * point below names the reference call site(s) it replaces (paths relative to
 * longspec/test/ in the reference tree).  INTEGRATION.md shows the ctypes binding a
 * maintainer would add at each of those call sites.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless noted;
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*) and
 *     returns immediately: no allocation, no implicit synchronisation, safe to
 *     capture in a CUDA graph;
 *   - return value: 0 = ok, negative = LS_ERR_*; ls_last_error() gives a message;
 *   - lengths that change every decode round (cache lengths, accepted count) are
 *     read from device memory so that a captured graph can be replayed;
 *   - dtype: LS_F16 or LS_BF16 for activations / KV; statistics are fp32;
 *   - head_dim is 128 (the reference hard-codes 1/sqrt(128), llama.py:95);
 *   - batch size is 1 (every shipped driver decodes one prompt at a time,
 *     inference_long-bench.py:119).
 *
 * Paged KV pool (one pool per layer, K and V separate):
 *     pool[page][kv_head][LS_PAGE][128]      (dtype)
 *   token t of the sequence lives in page page_table[t / LS_PAGE], slot t % LS_PAGE.
 */
#ifndef LONGSPEC_B200_H
#define LONGSPEC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LS_PAGE 128
#define LS_HEAD_DIM 128
#define LS_MASK_WORDS 4          /* tree masks are packed bit rows of 4 x uint64: <= 256 tree tokens per call */
#define LS_MAX_TOPK 32

#define LS_F16 0
#define LS_BF16 1

#define LS_OK 0
#define LS_ERR_INVALID_ARG (-1)
#define LS_ERR_UNSUPPORTED (-2)
#define LS_ERR_CUDA (-3)

int ls_version(void);
const char* ls_last_error(void);
/* number of SMs of the current device (used by callers to size split-KV workspaces) */
int ls_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * Attention over a paged KV cache with an optional tree-masked tail: ONE descriptor serves
 * the five attention call sites of a decode round.
 *
 * Keys are the cache rows [0, kv_len + n_tail), kv_len = *kv_len_ptr + kv_len_bias.
 *   prefix keys  j <  kv_len : visible to query row i iff
 *                    j >= kv_len - n_q + i - window      (only if window >= 0; flash-attn's
 *                                                         local rule i + sk - sq - left <= j)
 *                and j <  kv_len - (n_q - 1 - i)          (only if causal != 0; bottom-right)
 *   tail keys    j >= kv_len : visible iff bit (j - kv_len) of tail_mask[i] is set.
 * n_q = *n_q_ptr (number of valid query rows; rows >= n_q produce zeros), n_q <= q_rows.
 *
 * Replaces (reference call sites):
 *   target verify   llama.py:385-387 + 394-421   window=-1 causal=0 tail = this round's tokens
 *   draft self, step 0     llama_glide.py:261-262        window=512 causal=1 no tail
 *   draft self, tree step  llama_glide.py:300-302,309-329 window=512 causal=0 tail = tree so far
 *   draft cross, step 0    llama_glide.py:265             window=-1 causal=1 no tail
 *   draft cross, tree step llama_glide.py:297             window=-1 causal=0 no tail
 * The two-pass (prefix LSE / tree LSE / sigmoid merge) of the reference is computed as one
 * online softmax; results agree within the tolerance stated in tests/test_attention_gpu.py.
 * ---------------------------------------------------------------------------------------- */
typedef struct ls_attn_desc {
    const void* q;              /* [q_rows][n_heads][128] dtype, post-RoPE; row stride q_stride */
    void* out;                  /* [q_rows][n_heads][128] dtype */
    const void* k_pool;         /* paged pool, see above */
    const void* v_pool;
    const int32_t* page_table;  /* [>= ceil((kv_len+n_tail)/LS_PAGE)] */
    const int32_t* kv_len_ptr;  /* device scalar */
    const int32_t* n_q_ptr;     /* device scalar */
    const uint64_t* tail_mask;  /* [q_rows][LS_MASK_WORDS] or NULL when n_tail == 0 */
    float* part_o;              /* workspace [n_parts_total][n_kv_heads][q_rows*group][128] fp32 */
    float* part_lse;            /* workspace [n_parts_total][n_kv_heads][q_rows*group] fp32; -inf = this slot saw no key for the
                                   row: its part_o row is then unspecified and never read by the merges */
    int32_t kv_len_bias;        /* added to *kv_len_ptr (may be negative) */
    int32_t q_rows;             /* allocated query rows (<= 256) */
    int32_t n_heads;
    int32_t n_kv_heads;
    int32_t n_tail;             /* tree-masked keys after kv_len (0..256) */
    int32_t window;             /* -1 = unbounded, else left window (512 for the draft) */
    int32_t causal;             /* 0/1 */
    int32_t dtype;              /* LS_F16 / LS_BF16 */
    int32_t n_splits;           /* split-KV partial slots per kv head this GPU writes (>= 1); the tcgen05 path runs
                                   1 CTA per slot up to 256 packed rows and 1.5 CTAs per slot (heavy/heavy/light
                                   trios, attn_tc.cu) for 257..384 packed rows */
    int32_t n_pages;            /* pages in k_pool / v_pool (extent of the TMA tensor maps) */
    int32_t part_offset;        /* first partial slot this launch writes (sequence-sharded multi-GPU) */
    int32_t n_parts_total;      /* partial slots the combine step reduces over */
    int32_t kv_shard_begin;     /* this GPU owns prefix keys [kv_shard_begin, kv_shard_end) of the */
    int32_t kv_shard_end;       /*   global sequence; page_table indexes LOCAL pages; -1 = unbounded */
    int32_t tail_owner;         /* 1 if this GPU processes the tail keys (exactly one rank does) */
    float scale;                /* softmax scale, 1/sqrt(128) */
    int64_t q_stride;           /* elements between consecutive query rows of `q` (0 = dense, n_heads * 128): lets the
                                   q slice of a fused QKV GEMM output be read in place */
    uint32_t* fuse_sync;        /* NULL, or 2 * n_kv_heads zero-initialised counters: the tcgen05 path of ls_attn_partial
                                   then ALSO merges the partial slots and writes `out` (what ls_attn_combine does) in its
                                   tail - the CTAs of a kv head meet at a counter barrier and share the rows out - so the
                                   caller skips ls_attn_combine.  Honoured only when the call runs on the tcgen05 kernel
                                   with part_offset == 0, n_parts_total == n_splits and a grid that fits the device in
                                   one wave (ls_attn_fuses_combine tells); ignored otherwise. */
} ls_attn_desc;

/* split-KV partial pass: writes (normalised O, LSE) per (split, kv head, row) into the workspace */
int ls_attn_partial(const ls_attn_desc* d, void* stream);
/* 1 if ls_attn_partial(d) will merge the slots and write d->out itself (d->fuse_sync set and the conditions above hold),
 * 0 if the caller has to run ls_attn_combine, negative on a bad descriptor.  No GPU work. */
int ls_attn_fuses_combine(const ls_attn_desc* d);
/* Tuning knob, no reference counterpart: which tcgen05 kernel serves calls with three row tiles (257..384 packed rows,
 * the Llama-3 / QwQ verify shapes).  0 = two row tiles per CTA in heavy/heavy/light trios, 1 = one row tile per CTA
 * (default), 2 = one row tile per CTA in clusters of two with TMA multicast of the K/V tiles.  Sets the mode when
 * 0 <= mode <= 2 and returns the previous one (pass -1 to query); the environment variable LONGSPEC_TC1 gives the
 * initial value.  Results of the three agree within the rounding of P to the 16-bit type. */
int ls_attn_tc1_mode(int mode);
/* the general mma.sync data path behind ls_attn_partial (kept callable for cross-checks) */
int ls_attn_partial_mma(const ls_attn_desc* d, void* stream);
/* the tcgen05/TMA data path alone: unmasked full 128-key tiles of the prefix into n_splits_tc slots (window < 0,
 * <= 384 packed rows).  ls_attn_partial runs the same kernel with the masked remainder fused in as special tiles;
 * LONGSPEC_ATTN=tc+mma pairs this entry with one mma.sync slot for the remainder instead (cross-check) */
int ls_attn_partial_tc(const ls_attn_desc* d, int32_t n_splits_tc, void* stream);
/* reduces n_parts_total partial slots and writes `out`; run after the partials of all ranks are
 * in the workspace (single GPU: right after ls_attn_partial) */
int ls_attn_combine(const ls_attn_desc* d, void* stream);
/* multi-GPU: merges this launch's OWN slots [part_offset, part_offset + n_splits) into one normalised fp32
 * partial dst_o [n_kv_heads][q_rows*group][128] + natural-log LSE dst_lse [n_kv_heads][q_rows*group] - the
 * block a rank contributes to the exchange (1/n_splits of the raw slots); the final ls_attn_combine then runs
 * over the gathered blocks with n_parts_total = world.  No reference counterpart (single GPU). */
int ls_attn_reduce(const ls_attn_desc* d, float* dst_o, float* dst_lse, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sequence-sharded prefix KV over the GPUs of one node: the per-call exchange of the ranks' partials fused into
 * the two merge kernels over NVLink peer memory (no reference counterpart: the reference is single-GPU; SURVEY.md
 * section 8(e), option C).  Every rank owns an exchange buffer that all ranks have mapped (symmetric memory):
 *     O      [2 parities][world blocks][block_o]   fp32      LSE [2 parities][world blocks][block_lse] fp32
 *     flags  [world] uint32: epoch of the last block received from each source rank
 * ls_attn_reduce_push = ls_attn_reduce whose result goes straight into block `rank` of EVERY rank's buffer (peer
 * stores), followed by one release-store of the epoch into every rank's flag for this source.
 * ls_attn_combine_wait = ls_attn_combine over the `world` blocks of this rank's own buffer, after an acquire-spin
 * until every source's flag shows the current epoch.  Blocks alternate between two parities so that a fast rank's
 * next call cannot overwrite blocks a slow rank is still merging (a rank can be at most one call ahead: its call c+2
 * needs every other rank's block c+1, which is pushed after that rank finished merging call c).
 * The epoch lives in device memory (local_sync[0]) and is advanced by the combine kernel, so both calls can be
 * captured in a CUDA graph and replayed.  The merge order is rank order on every rank: bit-identical outputs.
 * ---------------------------------------------------------------------------------------- */
#define LS_MAX_PEERS 8
typedef struct ls_xchg {
    float* peer_o[LS_MAX_PEERS];        /* rank r's O exchange buffer as mapped into THIS process */
    float* peer_lse[LS_MAX_PEERS];
    uint32_t* peer_flags[LS_MAX_PEERS]; /* rank r's flag array [world] */
    uint32_t* local_sync;               /* this rank only: [0] epoch, [1] reduce ticket, [2] combine ticket (zeroed once) */
    int64_t block_o;                    /* floats per O block (capacity >= n_kv_heads * q_rows * group * 128 of any call) */
    int64_t block_lse;                  /* floats per LSE block */
    int32_t rank, world;
} ls_xchg;
int ls_attn_reduce_push(const ls_attn_desc* d, const ls_xchg* x, void* stream);
/* d->part_o / part_lse / n_parts_total are ignored: the kernel merges the `world` blocks of this rank's buffer */
int ls_attn_combine_wait(const ls_attn_desc* d, const ls_xchg* x, void* stream);
/* Generic small all-gather over the same buffer, flags and epoch (the chunk blocks / argmax pairs of a vocab-sharded
 * lm_head): ls_peer_push copies `bytes` from `src` into slot `rank` of every rank's O region (parity = epoch & 1) and
 * publishes the epoch; ls_peer_gather_wait spins until every source's flag shows it, copies the `world` slots (slot_bytes
 * apart, rank order) to `dst` and advances the epoch.  bytes <= slot_bytes <= block_o * 4, multiples of 16. */
int ls_peer_push(const void* src, int64_t bytes, int64_t slot_bytes, const ls_xchg* x, void* stream);
int ls_peer_gather_wait(void* dst, int64_t bytes, int64_t slot_bytes, const ls_xchg* x, void* stream);
/* All-reduce (sum) in the same scheme, for a tensor-parallel MLP (row-parallel down projection): after every rank has
 * ls_peer_push'ed its fp32 partial [n_elems], ls_peer_sum_wait adds the `world` slots in rank order (fp32), rounds ONCE to
 * `dtype` - like the single-GPU GEMM that accumulates in fp32 and rounds its output - and writes dst [n_elems]. */
int ls_peer_sum_wait(void* dst, int64_t n_elems, int32_t dtype, int64_t slot_bytes, const ls_xchg* x, void* stream);

/* ------------------------------------------------------------------------------------------
 * RoPE + KV append.  Replaces apply_rotary_pos_emb (llama.py:378, llama_glide.py:256,294) and
 * the cache scatters K_Cache[b, cache_lens + r] = k (llama.py:396-399, llama_glide.py:312-315)
 * and flash_attn_with_kvcache's in-place append (llama_glide.py:261).
 * q [rows][n_heads][128], k/v [rows][n_kv_heads][128] with row strides (elements) so that a
 * fused QKV GEMM output can be used in place.  q and k are rotated IN PLACE (cos/sin computed
 * in fp32 from inv_freq and rounded to dtype like HF's rotary embedding), then rows r < *n_rows_ptr
 * of k and v are written to cache slot *base_len_ptr + base_bias + slot_offset + r.
 * k_pool == NULL: rotate q only (the draft's cross attention).
 * ---------------------------------------------------------------------------------------- */
int ls_rope_append(void* q, void* k, const void* v, int64_t q_stride, int64_t k_stride, int64_t v_stride,
                   const int32_t* positions, const float* inv_freq,
                   void* k_pool, void* v_pool, const int32_t* page_table,
                   const int32_t* base_len_ptr, int32_t base_bias, int32_t slot_offset,
                   const int32_t* n_rows_ptr, int32_t rows, int32_t n_heads, int32_t n_kv_heads,
                   int32_t dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Device-resident state of one generation (bsz = 1).  All int32 unless noted; the host code
 * allocates it once and passes pointers.  Mirrors the loop variables of tree_spec_generate
 * (llama_glide.py:915-1126).
 * ---------------------------------------------------------------------------------------- */
typedef struct ls_tree_state {
    int32_t* cache_len;        /* `cache_lens`: first un-committed target slot (llama_glide.py:944) */
    int32_t* llm_kv_len;       /* `target_cache_lens_for_draft` (945) */
    int32_t* draft_len;        /* `draft_cache_lens` (946) */
    int32_t* acc_num;          /* accepted tokens of the previous round incl. bonus, 1..gamma+1 */
    int64_t* acc_ids;          /* [gamma+1] */
    int64_t* all_spec;         /* [fseq] tree tokens, [0] = root (962-963) */
    float* scores;             /* [fseq] cumulative log-probs `history_logp_sum` (991) */
    int32_t* parent;           /* [fseq] `father_index` */
    int32_t* depth;            /* [fseq] tree depth, root = 0 */
    uint64_t* tree_bits;       /* [fseq][LS_MASK_WORDS] packed `tree_mask` rows (984-985) */
    int64_t* step_ids;         /* [max(gamma+1, max level width)] token ids feeding the next draft step */
    int32_t* step_pos;         /* same length: RoPE positions of those rows */
    int32_t* step_rows;        /* scalar: valid rows of the next draft step */
    int64_t* veri_ids;         /* [q_v] `veri_spec` (1078-1081) */
    int32_t* veri_pos;         /* [q_v] position ids (llama.py:575-577) */
    uint64_t* veri_bits;       /* [q_v][LS_MASK_WORDS] packed `new_tree_mask` (1082-1085) */
    int32_t* veri_rows;        /* scalar: acc_num + S (pad rows are skipped) */
    int32_t* logit_rows;       /* [fseq] rows of the verify hidden states feeding lm_head (1087-1089) */
    int64_t* output_ids;       /* [max_gen_len] (937) */
    int32_t* counters;         /* [4]: count (1111), num (1112), tokens written, stop flag */
    int32_t* aux;              /* [8]: [0] draft_len + acc_num (kv length of the draft's step-0 self attention),
                                  [1] acc_num - 1 (hidden row feeding lm_head in step 0, 1019), [2] double_input (1145) */
    int32_t fseq;              /* S + 1 */
    int32_t gamma;
    int32_t prompt_len;
    int32_t max_gen_len;
} ls_tree_state;

/* Start of a round (llama_glide.py:999-1006,1113-1116): reset tree (root = last accepted token),
 * emit step_ids/step_pos/step_rows for draft step 0 (ids = acc_ids, pos = draft_len + i). */
int ls_round_begin(const ls_tree_state* s, void* stream);

/* Tree growth: fused log-softmax + cumulative score + global top-k over the flattened (M, V)
 * candidates + mask/parent/depth/position bookkeeping.  Replaces llama_glide.py:1019-1027
 * (level 0: `logits` is the single row acc_num-1 of the step-0 hidden states) and 1046-1075 (level >= 1).
 * logits [rows][vocab] dtype or fp32 (logits_fp32 != 0).  Also advances draft_len like
 * lines 1018 (level 0: += acc_num - 1).  Ties: larger score first, then lower flat index.
 * workspace: at least ls_grow_workspace_bytes(max_rows, vocab) bytes. */
int64_t ls_grow_workspace_bytes(int32_t max_rows, int32_t vocab);
int ls_tree_grow(const ls_tree_state* s, const void* logits, int32_t logits_fp32, int32_t dtype,
                 int32_t vocab, int32_t level, int32_t level_begin, int32_t level_end, int32_t n_new,
                 void* workspace, void* stream);

/* The same in two halves, for a vocab-sharded lm_head (multi-GPU; no reference counterpart): every rank computes the logits
 * of a chunk-aligned slice of the vocabulary (chunk = 2048 ids; `logits` row stride `ld`, first column = id chunk0 * 2048)
 * and ls_tree_grow_partial reduces it to per-chunk (max, sum exp, top-n_new) blocks of ls_grow_chunk_bytes() bytes each,
 * written to ws_slice[0 .. n_chunks_local) with GLOBAL token ids; after the ranks' slices have been all-gathered into one
 * workspace of n_chunks blocks (slices padded to equal length: blocks behind the end of the vocabulary are empty),
 * ls_tree_grow_finish does the merges and the tree update - identically on every rank.
 * ls_tree_grow == partial over the whole vocabulary + finish. */
int64_t ls_grow_chunk_bytes(void);
int ls_tree_grow_partial(const void* logits, int32_t logits_fp32, int32_t dtype, int32_t rows, int64_t ld, int32_t vocab,
                         int32_t chunk0, int32_t n_chunks_local, int32_t n_new, void* ws_slice, void* stream);
int ls_tree_grow_finish(const ls_tree_state* s, int32_t vocab, int32_t n_chunks, int32_t level, int32_t level_begin,
                        int32_t level_end, int32_t n_new, const void* workspace, void* stream);

/* Verify-input assembly (llama_glide.py:1077-1085, llama.py:575-577): draft_len += 1, veri_ids,
 * veri_pos, veri_bits, veri_rows, logit_rows. */
int ls_build_verify(const ls_tree_state* s, void* stream);

/* Row-wise argmax of [rows][vocab] logits (first index on ties) -> int64.  Replaces
 * llm_logits.argmax(-1) (llama_glide.py:1105) and the vanilla path's argmax (574). */
int ls_argmax_rows(const void* logits, int32_t logits_fp32, int32_t dtype, int32_t rows, int32_t vocab,
                   int64_t* out, void* stream);

/* Vocab-sharded form of the same: (value bits, GLOBAL index) int64 pairs per row over this rank's slice, then - on the
 * all-gathered pairs [world][rows][2] - the merge by (value desc, index asc).  index -1 marks an empty slice. */
int ls_argmax_rows_partial(const void* logits, int32_t logits_fp32, int32_t dtype, int32_t rows, int32_t vocab_local,
                           int64_t ld, int32_t vocab_begin, int64_t* out_pairs, void* stream);
int ls_argmax_merge(const int64_t* pairs, int32_t world, int32_t rows, int64_t* out, void* stream);

/* Greedy tree verification + last-layer accepted-path KV compaction + counters
 * (llama_glide.py:1104-1116 and tree_verification 1128-1175).  llm_pred [fseq] int64.
 * k_pool/v_pool/page_table: the LAST target layer.  Writes acc_ids/acc_num, output_ids, advances
 * cache_len / llm_kv_len, sets the stop flag (1118-1121: bit 1 = buffer nearly full, bit 0 = `stop_id`
 * among the accepted tokens; stop_id < 0 disables the eos test).  kv_bias is added to the cache slot
 * indices (minus the shard begin on the GPU that owns the generated tokens); k_pool == NULL skips the
 * compaction (GPUs that do not own those rows). */
int ls_tree_accept(const ls_tree_state* s, const int64_t* llm_pred, void* k_pool, void* v_pool,
                   const int32_t* page_table, int32_t n_kv_heads, int32_t dtype, int32_t non_leaf_len,
                   int64_t stop_id, int32_t kv_bias, void* stream);

/* Dense int tree mask (the reference's (M, N) int64 tensors) -> packed bit rows; lets the
 * reference-shaped operator seam (tree_decoding(..., tree_mask)) call the kernels above. */
int ls_pack_mask(const int64_t* dense, int32_t rows, int32_t cols, int64_t row_stride, uint64_t* bits,
                 void* stream);

/* ------------------------------------------------------------------------------------------
 * Small fused element-wise kernels of the decoder layer (HF LlamaRMSNorm / LlamaMLP used via
 * llama.py:35-42,463-498).  x, out [rows][hidden] dtype.
 * ---------------------------------------------------------------------------------------- */
/* if delta != NULL: h += delta first (residual add, llama.py:492,498); out = rmsnorm(h) * w */
int ls_add_rmsnorm(void* h, const void* delta, const void* w, void* out, int32_t rows, int32_t hidden,
                   float eps, int32_t dtype, void* stream);
/* out[r][c] = silu(gu[r][c]) * gu[r][ffn + c]  for a fused gate|up GEMM output [rows][2*ffn] */
int ls_silu_mul(const void* gu, void* out, int32_t rows, int32_t ffn, int32_t dtype, void* stream);
/* out[r] = table[ids[r]] (embedding gather with device ids, llama.py:579) */
int ls_embed(const void* table, const int64_t* ids, void* out, int32_t rows, int32_t hidden, int32_t dtype,
             void* stream);
/* out[r] = x[rows_idx[r]] (hidden-row gather, llama_glide.py:1087-1090) */
int ls_gather_rows(const void* x, const int32_t* rows_idx, void* out, int32_t rows, int32_t hidden,
                   int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LONGSPEC_B200_H */

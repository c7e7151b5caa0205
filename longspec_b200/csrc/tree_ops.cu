// Device-side tree bookkeeping of one decode round (bsz = 1): round begin, tree growth
// (fused log-softmax + cumulative score + top-k), verify-input assembly, row argmax, greedy
// acceptance + last-layer KV compaction.  All state stays on the GPU so that a round needs no
// host round trip except the stop test.  Reference: llama_glide.py:997-1121, 1128-1175.
#include <float.h>
#include <limits.h>

#include "common.cuh"

namespace ls {

constexpr int W = LS_MASK_WORDS;

__device__ __forceinline__ bool better(float va, int64_t ia, float vb, int64_t ib) {
    return va > vb || (va == vb && ia < ib);
}

// ------------------------------------------------------------------------------------------------
__global__ void round_begin_kernel(ls_tree_state s) {
    const int tid = threadIdx.x;
    const int acc = *s.acc_num;
    const int64_t root = s.acc_ids[acc - 1];
    const int dl = *s.draft_len;
    for (int u = tid; u < s.fseq; u += blockDim.x) {
        s.all_spec[u] = u == 0 ? root : 0;                    // llama_glide.py:1115-1116
        s.scores[u] = 0.f;                                     // 1001
        s.parent[u] = 0;                                       // 1000
        s.depth[u] = 0;
        for (int w = 0; w < W; ++w) s.tree_bits[u * W + w] = w == 0 ? 1ull : 0ull;   // 1113-1114
    }
    for (int i = tid; i <= s.gamma; i += blockDim.x) {
        s.step_ids[i] = s.acc_ids[min(i, acc - 1)];           // 1004 (rows >= acc_num are padding)
        s.step_pos[i] = dl + i;                                // 1005
    }
    if (tid == 0) {
        *s.step_rows = acc;
        s.aux[0] = dl + acc;        // kv length of the draft's step-0 self attention after the append
        s.aux[1] = acc - 1;         // hidden row that feeds lm_head in step 0 (1019)
    }
}

// ------------------------------------------------------------------------------------------------
// Tree growth, phase A: per (row, vocab chunk): max, sum exp(x - max), local top-k by (value, index).
constexpr int kChunk = 2048;
constexpr int kGrowThreads = 256;
constexpr int kMaxRowsB = 32;
// Workspace = one fixed-size block per vocab chunk: { max[32 rows], sum[32 rows], val[32 rows][32], idx[32 rows][32] }.
// Chunk-major so that the chunks a rank computed (vocab-sharded lm_head) are ONE
// contiguous piece that a single all-gather moves; the capacities are the kernels' limits, not the call's sizes.
constexpr int kBlkMax = 0, kBlkSum = kMaxRowsB, kBlkVal = 2 * kMaxRowsB, kBlkIdx = 2 * kMaxRowsB + kMaxRowsB * LS_MAX_TOPK;
constexpr int kBlkWords = 2 * kMaxRowsB + 2 * kMaxRowsB * LS_MAX_TOPK;       // 2112 words = 8448 bytes

template <typename T>
__device__ __forceinline__ float load_logit(const void* p, size_t i, int fp32) {
    return fp32 ? reinterpret_cast<const float*>(p)[i] : Num<T>::to_f(reinterpret_cast<const T*>(p)[i]);
}

template <typename T>
// `logits` holds the columns [chunk0 * kChunk, ...) of the vocabulary (row stride ld): the whole vocabulary on one GPU,
// a chunk-aligned slice per rank when the lm_head is vocab-sharded.  Block blockIdx.x of `ws` is chunk chunk0 + blockIdx.x;
// candidate indices are GLOBAL vocabulary ids.
__global__ void __launch_bounds__(kGrowThreads) grow_phase_a(const void* logits, int fp32, int vocab, int64_t ld, int chunk0,
                                                             int k, float* ws) {
    const int chunk = chunk0 + blockIdx.x, row = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int PER = kChunk / kGrowThreads;          // 8
    float v[PER];
    const int base = chunk * kChunk;
    float* blk = ws + static_cast<size_t>(blockIdx.x) * kBlkWords;
    float mx = -INFINITY;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const int c = base + e * kGrowThreads + tid;
        v[e] = c < vocab ? load_logit<T>(logits, static_cast<size_t>(row) * ld + (c - chunk0 * kChunk), fp32) : -INFINITY;
        mx = fmaxf(mx, v[e]);
    }
    __shared__ float red_f[kGrowThreads / 32];
    __shared__ int red_i[kGrowThreads / 32];
    __shared__ float bc_f;
    __shared__ int bc_i;
    mx = warp_max(mx);
    if (lane == 0) red_f[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        float m = red_f[0];
        for (int w = 1; w < kGrowThreads / 32; ++w) m = fmaxf(m, red_f[w]);
        bc_f = m;
    }
    __syncthreads();
    const float cmax = bc_f;
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < PER; ++e) sum += (v[e] == -INFINITY) ? 0.f : expf(v[e] - cmax);
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) red_f[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < kGrowThreads / 32; ++w) t += red_f[w];
        blk[kBlkMax + row] = cmax;
        blk[kBlkSum + row] = t;
    }
    // local top-k: k rounds of block arg-max (value desc, index asc)
    for (int it = 0; it < k; ++it) {
        float bv = -INFINITY;
        int bi = INT_MAX;
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const int c = base + e * kGrowThreads + tid;
            if (v[e] > bv || (v[e] == bv && c < bi)) { bv = v[e]; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        __syncthreads();
        if (lane == 0) { red_f[warp] = bv; red_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            float fv = red_f[0];
            int fi = red_i[0];
            for (int w = 1; w < kGrowThreads / 32; ++w)
                if (red_f[w] > fv || (red_f[w] == fv && red_i[w] < fi)) { fv = red_f[w]; fi = red_i[w]; }
            bc_f = fv;
            bc_i = fi;
            blk[kBlkVal + row * LS_MAX_TOPK + it] = fv;
            reinterpret_cast<int32_t*>(blk)[kBlkIdx + row * LS_MAX_TOPK + it] = fi;
        }
        __syncthreads();
        const int win = bc_i;
#pragma unroll
        for (int e = 0; e < PER; ++e)
            if (base + e * kGrowThreads + tid == win) v[e] = -INFINITY;
    }
}

// Phase B (one CTA) + tree state update.  Phase A left, per (row, chunk), the chunk's top-k in descending order,
// so the row-level and then the global top-k are k-way merges: one warp per row with the chunk lists spread over
// its lanes, then one warp over the rows.  Every step is a warp arg-max over list heads - no scans.
constexpr int kMaxChunksPerLane = 8;      // vocab <= 32 * 8 * 2048 = 524288
__global__ void __launch_bounds__(1024) grow_phase_b(ls_tree_state s, int vocab, int n_chunks, int k, int level,
                                                     int level_begin, int n_rows, int n_new, const float* ws) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ float row_val[kMaxRowsB][LS_MAX_TOPK];
    __shared__ int row_idx[kMaxRowsB][LS_MAX_TOPK];
    __shared__ float fin_val[LS_MAX_TOPK];
    __shared__ long long fin_flat[LS_MAX_TOPK];
    if (warp < n_rows) {
        const int row = warp;
        // log-normaliser of the row from the per-chunk (max, sum exp)
        float mx = -INFINITY;
        for (int c = lane; c < n_chunks; c += 32) mx = fmaxf(mx, ws[static_cast<size_t>(c) * kBlkWords + kBlkMax + row]);
        mx = warp_max(mx);
        float sm = 0.f;
        for (int c = lane; c < n_chunks; c += 32) {
            const float cm = ws[static_cast<size_t>(c) * kBlkWords + kBlkMax + row];
            // a chunk behind the end of the vocabulary (padding of the last rank's slice) holds max = -inf, sum = 0
            sm += cm == -INFINITY ? 0.f : ws[static_cast<size_t>(c) * kBlkWords + kBlkSum + row] * expf(cm - mx);
        }
        sm = warp_sum(sm);
        const float logsum = logf(sm);
        const float hist = level == 0 ? 0.f : s.scores[level_begin + row];
        const float* cv = ws + kBlkVal + row * LS_MAX_TOPK;                                  // + chunk * kBlkWords + position
        const int32_t* ci = reinterpret_cast<const int32_t*>(ws) + kBlkIdx + row * LS_MAX_TOPK;
        int ptr[kMaxChunksPerLane];
#pragma unroll
        for (int q = 0; q < kMaxChunksPerLane; ++q) ptr[q] = 0;
        for (int it = 0; it < n_new; ++it) {
            float bv = -INFINITY;
            int bi = INT_MAX, bq = -1;
#pragma unroll
            for (int q = 0; q < kMaxChunksPerLane; ++q) {
                const int c = lane + q * 32;
                if (c < n_chunks && ptr[q] < k) {
                    const float x = cv[static_cast<size_t>(c) * kBlkWords + ptr[q]];
                    const int id = ci[static_cast<size_t>(c) * kBlkWords + ptr[q]];
                    if (x > bv || (x == bv && id < bi)) { bv = x; bi = id; bq = q; }
                }
            }
            float wv = bv;
            int wi = bi, wl = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
                const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
                if (ov > wv || (ov == wv && oi < wi)) { wv = ov; wi = oi; wl = ol; }
            }
            if (lane == wl && bq >= 0) {
#pragma unroll
                for (int q = 0; q < kMaxChunksPerLane; ++q)
                    if (q == bq) ++ptr[q];
            }
            if (lane == 0) {
                // log_softmax = (x - max) - log(sum)  (torch), then + history (llama_glide.py:1047)
                row_val[row][it] = wv == -INFINITY ? -INFINITY : ((wv - mx) - logsum) + hist;
                row_idx[row][it] = wi;
            }
        }
    }
    __syncthreads();
    // global top-n_new over the rows' sorted lists by (score desc, flat index asc): one warp, lane = row
    if (warp == 0) {
        int ptr = 0;
        for (int it = 0; it < n_new; ++it) {
            float bv = -INFINITY;
            long long bf = LLONG_MAX;
            if (lane < n_rows && ptr < n_new) {
                bv = row_val[lane][ptr];
                bf = static_cast<long long>(lane) * vocab + row_idx[lane][ptr];
            }
            float wv = bv;
            long long wf = bf;
            int wl = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
                const long long of = __shfl_xor_sync(0xffffffffu, wf, o);
                const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
                if (ov > wv || (ov == wv && of < wf)) { wv = ov; wf = of; wl = ol; }
            }
            if (lane == wl) ++ptr;
            if (lane == 0) { fin_val[it] = wv; fin_flat[it] = wf; }
        }
    }
    __syncthreads();
    // tree update (llama_glide.py:1022-1027 for level 0, 1066-1075 for level >= 1)
    const int new_begin = level == 0 ? 1 : level_begin + n_rows;
    int dl = *s.draft_len;
    if (level == 0) dl += *s.acc_num - 1;                          // 1018
    if (tid < n_new) {
        const int u = new_begin + tid;
        const long long flat = fin_flat[tid];
        const int prow = static_cast<int>(flat / vocab);
        const int p = level == 0 ? 0 : level_begin + prow;
        s.all_spec[u] = flat - static_cast<long long>(prow) * vocab;
        s.scores[u] = fin_val[tid];
        s.parent[u] = p;
        s.depth[u] = level + 1;
        for (int w = 0; w < W; ++w) {
            uint64_t b = s.tree_bits[p * W + w];
            if ((u >> 6) == w) b |= 1ull << (u & 63);
            s.tree_bits[u * W + w] = b;
        }
        // inputs of the next draft step: this level's tokens at depth-based positions (1031-1033)
        s.step_ids[tid] = s.all_spec[u];
        s.step_pos[tid] = dl + level + 1;
    }
    if (tid == 0) {
        *s.step_rows = n_new;
        if (level == 0) *s.draft_len = dl;
    }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void shl_or(uint64_t (&dst)[W], const uint64_t (&src)[W], int sh) {
    // dst |= src << sh   (multi-word)
    const int ws = sh >> 6, bs = sh & 63;
    for (int w = W - 1; w >= 0; --w) {
        uint64_t x = 0;
        if (w - ws >= 0) x = src[w - ws] << bs;
        if (bs && w - ws - 1 >= 0) x |= src[w - ws - 1] >> (64 - bs);
        dst[w] |= x;
    }
}

__global__ void build_verify_kernel(ls_tree_state s, int q_v) {
    const int tid = threadIdx.x;
    const int acc = *s.acc_num;
    const int S = s.fseq - 1;
    const int cl = *s.cache_len;
    for (int r = tid; r < q_v; r += blockDim.x) {
        uint64_t bits[W];
        for (int w = 0; w < W; ++w) bits[w] = 0;
        int64_t id = 0;
        int pos = cl;
        if (r < acc) {                                   // accepted tokens: plain causal rows
            id = s.acc_ids[r];
            pos = cl + r;
            for (int b = 0; b <= r; ++b) bits[b >> 6] |= 1ull << (b & 63);
        } else if (r < acc + S) {                        // tree node u = r - acc + 1
            const int u = r - acc + 1;
            id = s.all_spec[u];
            pos = cl + acc + s.depth[u] - 1;             // mask.sum(-1) - 1 + cache_lens (llama.py:575-577)
            uint64_t anc[W];
            for (int w = 0; w < W; ++w) anc[w] = s.tree_bits[u * W + w];
            // drop the root bit (tree_mask[1:, 1:], llama_glide.py:1084): shift right by one
            for (int w = 0; w < W; ++w) anc[w] = (anc[w] >> 1) | (w + 1 < W ? anc[w + 1] << 63 : 0ull);
            shl_or(bits, anc, acc);
            for (int b = 0; b < acc; ++b) bits[b >> 6] |= 1ull << (b & 63);
        }
        s.veri_ids[r] = id;
        s.veri_pos[r] = pos;
        for (int w = 0; w < W; ++w) s.veri_bits[r * W + w] = bits[w];
    }
    for (int u = tid; u < s.fseq; u += blockDim.x) s.logit_rows[u] = acc - 1 + u;     // 1087-1089
    if (tid == 0) {
        *s.veri_rows = acc + S;
        *s.draft_len += 1;                                // 1077
    }
}

// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) argmax_rows_kernel(const void* logits, int fp32, int vocab, int64_t* out) {
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float bv = -INFINITY;
    int bi = INT_MAX;
    const size_t base = static_cast<size_t>(row) * vocab;
    if (!fp32 && (vocab & 7) == 0) {
        const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(logits) + base);
        for (int c = tid; c < vocab / 8; c += blockDim.x) {
            const uint4 u = p[c];
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = Num<T>::unpack(w4[e]);
                const int i0 = c * 8 + e * 2;
                if (f.x > bv) { bv = f.x; bi = i0; }
                if (f.y > bv) { bv = f.y; bi = i0 + 1; }
            }
        }
    } else {
        for (int c = tid; c < vocab; c += blockDim.x) {
            const float x = load_logit<T>(logits, base + c, fp32);
            if (x > bv) { bv = x; bi = c; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __shared__ float rv[32];
    __shared__ int ri[32];
    if (lane == 0) { rv[warp] = bv; ri[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (blockDim.x >> 5); ++w)
            if (rv[w] > bv || (rv[w] == bv && ri[w] < bi)) { bv = rv[w]; bi = ri[w]; }
        out[row] = bi == INT_MAX ? 0 : bi;
    }
}

// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) tree_accept_kernel(ls_tree_state s, const int64_t* llm_pred, T* k_pool, T* v_pool,
                                                          const int32_t* page_table, int n_kv_heads, int non_leaf_len,
                                                          int64_t stop_id, int kv_bias) {
    const int tid = threadIdx.x;
    // A round that was queued before the host saw the stop flag must not change the result: once the flag is set,
    // acceptance is a no-op (output_ids, counters and the accepted state stay final).
    if (s.counters[3] != 0) return;
    __shared__ int sh_last, sh_acc;
    __shared__ int sh_map[64 * W];
    __shared__ unsigned char sh_ok[64 * W];
    const int f = s.fseq;
    const int prev_acc = *s.acc_num;
    const int root_slot = *s.cache_len + prev_acc - 1;                    // cache_lens += acc_num - 1 (1104)
    // verify[u]: the target's prediction at the parent equals the node's token (1138-1139)
    for (int u = tid; u < f; u += blockDim.x)
        sh_ok[u] = u == 0 ? 1 : (llm_pred[s.parent[u]] == s.all_spec[u]);
    if (tid == 0) sh_last = 0;
    __syncthreads();
    // final_verify[u]: every ancestor (and u) verified (1140-1141); last = deepest verified (1144)
    for (int u = tid; u < f; u += blockDim.x) {
        bool all_ok = true;
        for (int w = 0; w < W; ++w) {
            uint64_t b = s.tree_bits[u * W + w];
            while (b) {
                const int j = w * 64 + __ffsll(static_cast<long long>(b)) - 1;
                b &= b - 1;
                all_ok = all_ok && sh_ok[j];
            }
        }
        if (all_ok) atomicMax(&sh_last, u);
    }
    __syncthreads();
    if (tid == 0) {
        const int last = sh_last;
        int n = 0;
        for (int w = 0; w < W; ++w) {
            uint64_t b = s.tree_bits[last * W + w];
            while (b) {                                                   // selected indices ascending (1153-1154)
                sh_map[n++] = w * 64 + __ffsll(static_cast<long long>(b)) - 1;
                b &= b - 1;
            }
        }
        sh_acc = n;                                                       // acc_num = select_mask.sum() (1148)
    }
    __syncthreads();
    const int acc = sh_acc;
    // last-layer KV compaction: rows root_slot + idx_k -> root_slot + k (1159-1173).  idx_k >= k and
    // strictly increasing, so an ascending in-place copy never overwrites a row that is still needed.
    if (k_pool != nullptr) {
        for (int k = 0; k < acc; ++k) {
            const int src = root_slot + kv_bias + sh_map[k], dst = root_slot + kv_bias + k;
            if (src != dst) {
                const int ps = page_table[src >> 7], pd = page_table[dst >> 7];
                for (int idx = tid; idx < n_kv_heads * 16; idx += blockDim.x) {
                    const int h = idx >> 4, chunk = idx & 15;
                    const size_t so = ((static_cast<size_t>(ps) * n_kv_heads + h) * LS_PAGE + (src & 127)) * LS_HEAD_DIM + chunk * 8;
                    const size_t dn = ((static_cast<size_t>(pd) * n_kv_heads + h) * LS_PAGE + (dst & 127)) * LS_HEAD_DIM + chunk * 8;
                    *reinterpret_cast<uint4*>(k_pool + dn) = *reinterpret_cast<const uint4*>(k_pool + so);
                    *reinterpret_cast<uint4*>(v_pool + dn) = *reinterpret_cast<const uint4*>(v_pool + so);
                }
            }
            __syncthreads();
        }
    }
    // outputs and counters (1105-1116)
    const int new_cache_len = root_slot + 1;
    const int out_base = new_cache_len - s.prompt_len;
    __shared__ int sh_stop;
    if (tid == 0) sh_stop = 0;
    __syncthreads();
    if (tid < acc) {
        const int64_t tok = llm_pred[sh_map[tid]];                        // acc_ids = output_ids.gather(index_mapping) (1155)
        s.acc_ids[tid] = tok;
        if (out_base + tid < s.max_gen_len) s.output_ids[out_base + tid] = tok;
        if (stop_id >= 0 && tok == stop_id) sh_stop = 1;
    }
    __syncthreads();
    if (tid == 0) {
        *s.cache_len = new_cache_len;
        *s.llm_kv_len += acc;                                             // 1110
        *s.acc_num = acc;
        s.counters[0] += acc - 1;                                         // count (1111)
        s.counters[1] += 1;                                               // num (1112)
        s.counters[2] = out_base + acc;                                   // tokens written so far
        int stop = sh_stop;
        if (new_cache_len + acc - s.prompt_len + s.gamma + 2 > s.max_gen_len) stop |= 2;   // 1118
        s.counters[3] = stop;
        s.aux[2] = sh_last >= non_leaf_len;                               // double_input (1145), unused by callers
    }
}

__global__ void pack_mask_kernel(const int64_t* dense, int rows, int cols, int64_t row_stride, uint64_t* bits) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    uint64_t b[W];
    for (int w = 0; w < W; ++w) b[w] = 0;
    for (int c = 0; c < cols; ++c)
        if (dense[r * row_stride + c] != 0) b[c >> 6] |= 1ull << (c & 63);
    for (int w = 0; w < W; ++w) bits[r * W + w] = b[w];
}

// Vocab-sharded argmax: every rank reduces its vocabulary slice to (value, GLOBAL index) per row; after the exchange the
// candidates of all ranks are merged by (value desc, index asc) - the order argmax_rows_kernel itself uses.
template <typename T>
__global__ void __launch_bounds__(1024) argmax_rows_partial_kernel(const void* logits, int fp32, int vocab_local, int64_t ld,
                                                                   int vocab_begin, int64_t* out_pairs) {
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float bv = -INFINITY;
    int bi = INT_MAX;
    const size_t base = static_cast<size_t>(row) * ld;
    for (int c = tid; c < vocab_local; c += blockDim.x) {
        const float x = load_logit<T>(logits, base + c, fp32);
        if (x > bv) { bv = x; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __shared__ float rv[32];
    __shared__ int ri[32];
    if (lane == 0) { rv[warp] = bv; ri[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (blockDim.x >> 5); ++w)
            if (rv[w] > bv || (rv[w] == bv && ri[w] < bi)) { bv = rv[w]; bi = ri[w]; }
        out_pairs[2 * row] = static_cast<int64_t>(__float_as_uint(bv));
        out_pairs[2 * row + 1] = bi == INT_MAX ? -1 : static_cast<int64_t>(bi) + vocab_begin;
    }
}

__global__ void argmax_merge_kernel(const int64_t* pairs, int world, int rows, int64_t* out) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    float bv = -INFINITY;
    int64_t bi = -1;
    for (int r = 0; r < world; ++r) {
        const float v = __uint_as_float(static_cast<uint32_t>(pairs[(static_cast<size_t>(r) * rows + row) * 2]));
        const int64_t i = pairs[(static_cast<size_t>(r) * rows + row) * 2 + 1];
        if (i >= 0 && (bi < 0 || v > bv || (v == bv && i < bi))) { bv = v; bi = i; }
    }
    out[row] = bi < 0 ? 0 : bi;
}

static int check_state(const ls_tree_state* s) {
    LS_REQUIRE(s != nullptr, "tree state is null");
    LS_REQUIRE(s->fseq >= 2 && s->fseq <= 64 * W, "fseq %d out of range [2,%d]", s->fseq, 64 * W);
    LS_REQUIRE(s->gamma >= 1 && s->gamma <= 16, "gamma %d out of range", s->gamma);
    LS_REQUIRE(s->cache_len && s->llm_kv_len && s->draft_len && s->acc_num && s->acc_ids && s->all_spec && s->scores &&
                   s->parent && s->depth && s->tree_bits && s->step_ids && s->step_pos && s->step_rows && s->veri_ids &&
                   s->veri_pos && s->veri_bits && s->veri_rows && s->logit_rows && s->output_ids && s->counters && s->aux,
               "tree state has a null pointer");
    return LS_OK;
}

}  // namespace ls

extern "C" int ls_round_begin(const ls_tree_state* s, void* stream) {
    using namespace ls;
    int rc = check_state(s);
    if (rc) return rc;
    round_begin_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(*s);
    return check_launch("ls_round_begin");
}

extern "C" int64_t ls_grow_chunk_bytes(void) { return static_cast<int64_t>(ls::kBlkWords) * 4; }

extern "C" int64_t ls_grow_workspace_bytes(int32_t max_rows, int32_t vocab) {
    (void)max_rows;                                   // blocks have the kernels' fixed row capacity
    const int64_t n_chunks = (vocab + ls::kChunk - 1) / ls::kChunk;
    return (n_chunks + LS_MAX_PEERS) * ls_grow_chunk_bytes() + 256;      // + padding of the last rank's slice when sharded
}

extern "C" int ls_tree_grow_partial(const void* logits, int32_t logits_fp32, int32_t dtype, int32_t rows, int64_t ld,
                                    int32_t vocab, int32_t chunk0, int32_t n_chunks_local, int32_t n_new, void* ws_slice,
                                    void* stream) {
    using namespace ls;
    LS_REQUIRE(logits && ws_slice, "ls_tree_grow_partial: null logits/workspace");
    LS_REQUIRE(n_new >= 1 && n_new <= LS_MAX_TOPK, "ls_tree_grow_partial: n_new %d out of range [1,%d]", n_new, LS_MAX_TOPK);
    LS_REQUIRE(rows >= 1 && rows <= kMaxRowsB, "ls_tree_grow_partial: %d rows out of range [1,%d]", rows, kMaxRowsB);
    LS_REQUIRE(chunk0 >= 0 && n_chunks_local >= 1 && ld >= 0, "ls_tree_grow_partial: bad slice");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(n_chunks_local, rows);
    if (dtype == LS_F16)
        grow_phase_a<__half><<<grid, kGrowThreads, 0, st>>>(logits, logits_fp32, vocab, ld, chunk0, n_new, static_cast<float*>(ws_slice));
    else
        grow_phase_a<__nv_bfloat16><<<grid, kGrowThreads, 0, st>>>(logits, logits_fp32, vocab, ld, chunk0, n_new,
                                                                   static_cast<float*>(ws_slice));
    return check_launch("ls_tree_grow_partial");
}

extern "C" int ls_tree_grow_finish(const ls_tree_state* s, int32_t vocab, int32_t n_chunks, int32_t level,
                                   int32_t level_begin, int32_t level_end, int32_t n_new, const void* workspace,
                                   void* stream) {
    using namespace ls;
    int rc = check_state(s);
    if (rc) return rc;
    LS_REQUIRE(workspace, "ls_tree_grow_finish: null workspace");
    LS_REQUIRE(n_new >= 1 && n_new <= LS_MAX_TOPK, "ls_tree_grow: n_new %d out of range [1,%d]", n_new, LS_MAX_TOPK);
    const int n_rows = level == 0 ? 1 : level_end - level_begin;
    LS_REQUIRE(n_rows >= 1 && n_rows <= kMaxRowsB, "ls_tree_grow: level width %d out of range [1,%d]", n_rows, kMaxRowsB);
    LS_REQUIRE(level >= 0 && level < s->gamma, "ls_tree_grow: level %d out of range", level);
    LS_REQUIRE(vocab >= n_new, "ls_tree_grow: vocab smaller than n_new");
    LS_REQUIRE(n_chunks >= 1 && n_chunks <= 32 * kMaxChunksPerLane, "ls_tree_grow: vocab %d too large (%d chunks)", vocab, n_chunks);
    const int new_begin = level == 0 ? 1 : level_end;
    LS_REQUIRE(new_begin + n_new <= s->fseq, "ls_tree_grow: level overflows fseq");
    grow_phase_b<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(*s, vocab, n_chunks, n_new, level, level_begin, n_rows, n_new,
                                                                    static_cast<const float*>(workspace));
    return check_launch("ls_tree_grow_finish");
}

extern "C" int ls_tree_grow(const ls_tree_state* s, const void* logits, int32_t logits_fp32, int32_t dtype,
                            int32_t vocab, int32_t level, int32_t level_begin, int32_t level_end, int32_t n_new,
                            void* workspace, void* stream) {
    const int n_rows = level == 0 ? 1 : level_end - level_begin;   // level 0: the caller passes the single row acc_num-1
    const int n_chunks = (vocab + ls::kChunk - 1) / ls::kChunk;
    int rc = ls_tree_grow_partial(logits, logits_fp32, dtype, n_rows, vocab, vocab, 0, n_chunks, n_new, workspace, stream);
    if (rc) return rc;
    return ls_tree_grow_finish(s, vocab, n_chunks, level, level_begin, level_end, n_new, workspace, stream);
}

extern "C" int ls_build_verify(const ls_tree_state* s, void* stream) {
    using namespace ls;
    int rc = check_state(s);
    if (rc) return rc;
    const int q_v = s->fseq - 1 + s->gamma + 1;
    LS_REQUIRE(q_v <= 64 * W, "ls_build_verify: q_v %d exceeds %d", q_v, 64 * W);
    build_verify_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(*s, q_v);
    return check_launch("ls_build_verify");
}

extern "C" int ls_argmax_rows(const void* logits, int32_t logits_fp32, int32_t dtype, int32_t rows, int32_t vocab,
                              int64_t* out, void* stream) {
    using namespace ls;
    LS_REQUIRE(logits && out && rows >= 1 && vocab >= 1, "ls_argmax_rows: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == LS_F16)
        argmax_rows_kernel<__half><<<rows, 1024, 0, st>>>(logits, logits_fp32, vocab, out);
    else
        argmax_rows_kernel<__nv_bfloat16><<<rows, 1024, 0, st>>>(logits, logits_fp32, vocab, out);
    return check_launch("ls_argmax_rows");
}

extern "C" int ls_tree_accept(const ls_tree_state* s, const int64_t* llm_pred, void* k_pool, void* v_pool,
                              const int32_t* page_table, int32_t n_kv_heads, int32_t dtype, int32_t non_leaf_len,
                              int64_t stop_id, int32_t kv_bias, void* stream) {
    using namespace ls;
    int rc = check_state(s);
    if (rc) return rc;
    LS_REQUIRE(llm_pred, "ls_tree_accept: null llm_pred");
    LS_REQUIRE(k_pool == nullptr || (v_pool && page_table && n_kv_heads >= 1), "ls_tree_accept: bad KV arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == LS_F16)
        tree_accept_kernel<__half><<<1, 256, 0, st>>>(*s, llm_pred, static_cast<__half*>(k_pool), static_cast<__half*>(v_pool),
                                                      page_table, n_kv_heads, non_leaf_len, stop_id, kv_bias);
    else
        tree_accept_kernel<__nv_bfloat16><<<1, 256, 0, st>>>(*s, llm_pred, static_cast<__nv_bfloat16*>(k_pool),
                                                             static_cast<__nv_bfloat16*>(v_pool), page_table, n_kv_heads,
                                                             non_leaf_len, stop_id, kv_bias);
    return check_launch("ls_tree_accept");
}

extern "C" int ls_pack_mask(const int64_t* dense, int32_t rows, int32_t cols, int64_t row_stride, uint64_t* bits,
                            void* stream) {
    using namespace ls;
    LS_REQUIRE(dense && bits && rows >= 1 && cols >= 1 && cols <= 64 * W, "ls_pack_mask: bad arguments (cols <= %d)", 64 * W);
    pack_mask_kernel<<<(rows + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(dense, rows, cols, row_stride, bits);
    return check_launch("ls_pack_mask");
}

extern "C" int ls_argmax_rows_partial(const void* logits, int32_t logits_fp32, int32_t dtype, int32_t rows, int32_t vocab_local,
                                      int64_t ld, int32_t vocab_begin, int64_t* out_pairs, void* stream) {
    using namespace ls;
    LS_REQUIRE(logits && out_pairs && rows >= 1 && vocab_local >= 0, "ls_argmax_rows_partial: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == LS_F16)
        argmax_rows_partial_kernel<__half><<<rows, 1024, 0, st>>>(logits, logits_fp32, vocab_local, ld, vocab_begin, out_pairs);
    else
        argmax_rows_partial_kernel<__nv_bfloat16><<<rows, 1024, 0, st>>>(logits, logits_fp32, vocab_local, ld, vocab_begin, out_pairs);
    return check_launch("ls_argmax_rows_partial");
}

extern "C" int ls_argmax_merge(const int64_t* pairs, int32_t world, int32_t rows, int64_t* out, void* stream) {
    using namespace ls;
    LS_REQUIRE(pairs && out && world >= 1 && rows >= 1, "ls_argmax_merge: bad arguments");
    argmax_merge_kernel<<<(rows + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(pairs, world, rows, out);
    return check_launch("ls_argmax_merge");
}

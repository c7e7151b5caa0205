// Split-KV paged attention with a tree-masked tail (ls_attn_partial / ls_attn_combine).
//
// v0 data path: cp.async (LDGSTS) 3-stage pipeline of 64-key K/V tiles from the paged pool into
// XOR-swizzled shared memory, mma.sync m16n8k16 (HMMA) for QK^T and PV with fp32 online softmax in
// registers.  One CTA = (split, kv head, block of 128 packed query rows), a packed row being
// (query token i, head g of the GQA group): r = i * G + g, so all G heads that share a kv head reuse
// the same K/V tile from shared memory.  The tcgen05/TMA kernel (attn_tc.cu) supersedes this one for
// the long-prefix verify shape; this file stays as the general path (any GQA group, windows, tails)
// and as the in-repo numerical cross-check.
//
// Reference semantics: see include/longspec_b200.h (ls_attn_desc) and llama.py:385-421,
// llama_glide.py:261-329.
#include <limits.h>

#include "common.cuh"

namespace ls {

constexpr int kD = LS_HEAD_DIM;   // 128
constexpr int kBN = 64;           // keys per tile
constexpr int kBM = 128;          // packed rows per CTA
constexpr int kWarps = 8;
constexpr int kStages = 3;
constexpr int kTileBytes = kBN * kD * 2;                      // 16 KB
constexpr int kSmemBytes = kStages * 2 * kTileBytes;          // 96 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                                        uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// byte offset of 16-byte chunk `chunk` (0..15) of row `row` inside a [rows][128] 2-byte tile
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return row * 256 + ((chunk ^ (row & 7)) << 4); }

struct TilePlan {
    int n_prefix;     // prefix tiles of this CTA
    int tile0;        // first prefix tile index (local 64-key tiles)
    int n_tail;       // tail tiles of this CTA
    int tail_tok0;    // local token index of the first tail key
    int last_tok;     // last valid local token (clamp for loads)
};

template <typename T>
__global__ void __launch_bounds__(kWarps * 32, 1) attn_partial_kernel(const ls_attn_desc d, const int remainder) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int split = blockIdx.x, kvh = blockIdx.y, rb = blockIdx.z;
    const int G = d.n_heads / d.n_kv_heads;
    const int n_q = min(*d.n_q_ptr, d.q_rows);
    const int R = n_q * G;
    const int row0 = rb * kBM;
    if (row0 >= R) return;

    const int kv_len = *d.kv_len_ptr + d.kv_len_bias;                 // global prefix length
    const int sb = d.kv_shard_begin;
    const int se = d.kv_shard_end < 0 ? INT_MAX : d.kv_shard_end;
    const int loc_end = max(min(kv_len, se) - sb, 0);                 // local prefix end
    int loc_beg = 0;
    if (d.window >= 0) loc_beg = min(max(kv_len - n_q - d.window - sb, 0), loc_end);
    if (remainder) loc_beg = min(tc_full_tiles(d, kv_len, n_q) * 128, loc_end);   // the rest is on the tcgen05 kernel
    TilePlan tp;
    {
        const int tb = loc_beg / kBN, te = (loc_end + kBN - 1) / kBN;
        const int nt = max(te - tb, 0);
        const int per = (nt + d.n_splits - 1) / d.n_splits;
        const int a = min(tb + split * per, te), b = min(a + per, te);
        tp.tile0 = a;
        tp.n_prefix = max(b - a, 0);
        const bool tail_here = d.tail_owner && d.n_tail > 0 && split == d.n_splits - 1;
        tp.n_tail = tail_here ? (d.n_tail + kBN - 1) / kBN : 0;
        tp.tail_tok0 = kv_len - sb;
        tp.last_tok = max((d.tail_owner ? kv_len - sb + d.n_tail : loc_end) - 1, 0);
    }
    const int n_tiles = tp.n_prefix + tp.n_tail;

    const T* __restrict__ kpool = reinterpret_cast<const T*>(d.k_pool);
    const T* __restrict__ vpool = reinterpret_cast<const T*>(d.v_pool);
    const int Hk = d.n_kv_heads;

    auto tile_tok0 = [&](int it) {
        return it < tp.n_prefix ? (tp.tile0 + it) * kBN : tp.tail_tok0 + (it - tp.n_prefix) * kBN;
    };
    auto load_tile = [&](int it, int stage) {
        const int tok0 = tile_tok0(it);
        uint8_t* ks = smem + stage * 2 * kTileBytes;
        uint8_t* vs = ks + kTileBytes;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int idx = tid + c * kWarps * 32;
            const int row = idx >> 4, chunk = idx & 15;
            const int tok = min(tok0 + row, tp.last_tok);
            const int page = __ldg(d.page_table + (tok >> 7));
            const size_t off = ((static_cast<size_t>(page) * Hk + kvh) * LS_PAGE + (tok & (LS_PAGE - 1))) * kD + chunk * 8;
            cp_async16(smem_u32(ks + swz(row, chunk)), kpool + off);
            cp_async16(smem_u32(vs + swz(row, chunk)), vpool + off);
        }
    };

    // ---- stage Q (this CTA's 128 packed rows) through shared memory into A fragments -------------
    const T* __restrict__ q = reinterpret_cast<const T*>(d.q);
    const size_t q_stride = d.q_stride > 0 ? static_cast<size_t>(d.q_stride) : static_cast<size_t>(d.n_heads) * 128;
    {
        uint8_t* qs = smem;   // 32 KB, reuses stage 0/1 space before the pipeline starts
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int idx = tid + c * kWarps * 32;
            const int row = idx >> 4, chunk = idx & 15;
            const int r = row0 + row;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (r < R) {
                const int i = r / G, g = r - i * G;
                val = *reinterpret_cast<const uint4*>(q + static_cast<size_t>(i) * q_stride + static_cast<size_t>(kvh * G + g) * kD + chunk * 8);
            }
            *reinterpret_cast<uint4*>(qs + swz(row, chunk)) = val;
        }
    }
    __syncthreads();
    uint32_t qf[8][4];
    {
        const int row = warp * 16 + (lane & 15);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const int chunk = ks * 2 + (lane >> 4);
            ldsm_x4(smem_u32(smem + swz(row, chunk)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
        }
    }
    __syncthreads();

    // ---- per-thread row bookkeeping ---------------------------------------------------------------
    // accumulator rows of this thread: rA = warp*16 + lane/4, rB = rA + 8
    const int rloc[2] = {warp * 16 + (lane >> 2), warp * 16 + (lane >> 2) + 8};
    int qi[2];            // query token index of each row (or -1 when the row is padding)
    int lo_g[2], hi_g[2]; // visible global prefix key range [lo, hi)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = row0 + rloc[h];
        qi[h] = r < R ? r / G : -1;
        const int i = max(qi[h], 0);
        lo_g[h] = d.window >= 0 ? kv_len - n_q + i - d.window : INT_MIN;
        hi_g[h] = d.causal ? kv_len - (n_q - 1 - i) : kv_len;
        hi_g[h] = min(hi_g[h], se);
        if (qi[h] < 0) hi_g[h] = INT_MIN;   // padding row: nothing visible
    }
    // range of global keys fully visible to every VALID row of this CTA (fast path test)
    int cta_lo, cta_hi;
    {
        const int i_first = row0 / G, i_last = (min(row0 + kBM, R) - 1) / G;
        cta_lo = d.window >= 0 ? kv_len - n_q + i_last - d.window : INT_MIN;
        cta_hi = d.causal ? kv_len - (n_q - 1 - i_first) : kv_len;
        cta_hi = min(cta_hi, se);
        // padding rows (r >= R) carry q = 0 and are never written, so they may take the unmasked path
    }

    const float c2 = d.scale * 1.4426950408889634f;   // scores are kept in the log2 domain
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float o[16][4];
#pragma unroll
    for (int n = 0; n < 16; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[n][e] = 0.f;

    // ---- pipeline prologue -----------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
        if (s < n_tiles) load_tile(s, s);
        cp_async_commit();
    }

    for (int it = 0; it < n_tiles; ++it) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        {   // prefetch tile it + kStages - 1 into the stage consumed at iteration it - 1
            const int nx = it + kStages - 1;
            if (nx < n_tiles) load_tile(nx, nx % kStages);
            cp_async_commit();
        }
        const uint8_t* ks = smem + (it % kStages) * 2 * kTileBytes;
        const uint8_t* vs = ks + kTileBytes;

        // S = Q K^T  (16 rows x 64 keys per warp)
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[n][e] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                const int key = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int chunk = kk * 2 + ((lane >> 3) & 1);
                uint32_t b0, b1, b2, b3;
                ldsm_x4(smem_u32(ks + swz(key, chunk)), b0, b1, b2, b3);
                mma16816<T>(s[np * 2], qf[kk], b0, b1);
                mma16816<T>(s[np * 2 + 1], qf[kk], b2, b3);
            }
        }

        // masking
        const bool is_tail = it >= tp.n_prefix;
        const int tok0 = tile_tok0(it);
        if (is_tail) {
            const int t0 = (it - tp.n_prefix) * kBN;     // tail key index of column 0
            uint64_t bits[2];
#pragma unroll
            for (int h = 0; h < 2; ++h)
                bits[h] = qi[h] >= 0 ? d.tail_mask[static_cast<size_t>(qi[h]) * LS_MASK_WORDS + (t0 >> 6)] : 0ull;
#pragma unroll
            for (int n = 0; n < 8; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = n * 8 + ((lane & 3) << 1) + (e & 1);
                    const bool vis = (t0 + col < d.n_tail) && ((bits[e >> 1] >> col) & 1ull);
                    s[n][e] = vis ? s[n][e] * c2 : -INFINITY;
                }
        } else {
            const int g0 = tok0 + sb;                    // global key index of column 0
            if (g0 >= cta_lo && g0 + kBN <= cta_hi) {
#pragma unroll
                for (int n = 0; n < 8; ++n)
#pragma unroll
                    for (int e = 0; e < 4; ++e) s[n][e] *= c2;
            } else {
#pragma unroll
                for (int n = 0; n < 8; ++n)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int j = g0 + n * 8 + ((lane & 3) << 1) + (e & 1);
                        const bool vis = j >= lo_g[e >> 1] && j < hi_g[e >> 1];
                        s[n][e] = vis ? s[n][e] * c2 : -INFINITY;
                    }
            }
        }

        // online softmax (rows live in quads of 4 lanes)
        uint32_t pf[4][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float mx = -INFINITY;
#pragma unroll
            for (int n = 0; n < 8; ++n) mx = fmaxf(mx, fmaxf(s[n][2 * h], s[n][2 * h + 1]));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float m_new = fmaxf(m_run[h], mx);
            const float m_use = m_new == -INFINITY ? 0.f : m_new;
            const float alpha = fast_exp2(m_run[h] - m_use);
            float sum = 0.f;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float p0 = fast_exp2(s[n][2 * h] - m_use), p1 = fast_exp2(s[n][2 * h + 1] - m_use);
                sum += p0 + p1;
                // A fragment of P for k16 step n/2: regs {0,1} from even n8 tile, {2,3} from odd one
                pf[n >> 1][(n & 1) * 2 + h] = Num<T>::pack(p0, p1);
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            l_run[h] = l_run[h] * alpha + sum;
            m_run[h] = m_new;
#pragma unroll
            for (int n = 0; n < 16; ++n) {
                o[n][2 * h] *= alpha;
                o[n][2 * h + 1] *= alpha;
            }
        }

        // O += P V
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int dp = 0; dp < 8; ++dp) {
                const int key = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                const int chunk = dp * 2 + (lane >> 4);
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(smem_u32(vs + swz(key, chunk)), b0, b1, b2, b3);
                mma16816<T>(o[dp * 2], pf[kk], b0, b1);
                mma16816<T>(o[dp * 2 + 1], pf[kk], b2, b3);
            }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: normalised partial O and natural-log LSE --------------------------------------
    const int Rtot = d.q_rows * G;
    const int part = d.part_offset + split;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = row0 + rloc[h];
        if (r >= R) continue;
        const float inv = l_run[h] > 0.f ? 1.f / l_run[h] : 0.f;
        const size_t base = (static_cast<size_t>(part) * Hk + kvh) * Rtot + r;
        if ((lane & 3) == 0)
            d.part_lse[base] = l_run[h] > 0.f ? (m_run[h] + log2f(l_run[h])) * 0.6931471805599453f : -INFINITY;
        float* po = d.part_o + base * kD;
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            float2 v2 = make_float2(o[n][2 * h] * inv, o[n][2 * h + 1] * inv);
            *reinterpret_cast<float2*>(po + n * 8 + ((lane & 3) << 1)) = v2;
        }
    }
}

// One warp per (packed row, kv head): merge n_parts_total partials, write `out` in dtype.  The LSEs of up to 32
// parts are held one per lane, weights are broadcast by shuffle and the 512-byte partial rows are fetched with
// independent 16-byte loads (4 in flight per lane), so the merge is bandwidth- rather than latency-bound.
// REDUCE = false: merge all n_parts_total slots and write `out` (dtype).  REDUCE = true (ls_attn_reduce): merge only
// this launch's own slots [part_offset, part_offset + n_splits) into ONE normalised fp32 partial + LSE at
// dst_o / dst_lse (slot layout [kv head][row][128]) - the block a rank contributes to the multi-GPU exchange.
template <typename T, bool REDUCE>
__global__ void __launch_bounds__(256) attn_combine_kernel(const ls_attn_desc d, float* __restrict__ dst_o, float* __restrict__ dst_lse) {
    const int G = d.n_heads / d.n_kv_heads;
    const int Rtot = d.q_rows * G;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_global >= Rtot * d.n_kv_heads) return;
    const int kvh = warp_global / Rtot, r = warp_global - kvh * Rtot;
    const int i = r / G, g = r - i * G;
    const int n_q = min(*d.n_q_ptr, d.q_rows);
    T* out = reinterpret_cast<T*>(d.out) + (static_cast<size_t>(i) * d.n_heads + kvh * G + g) * kD + lane * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float lse_out = -INFINITY;
    if (i < n_q) {
        const size_t stride = static_cast<size_t>(d.n_kv_heads) * Rtot;        // between parts, in rows
        const size_t row0 = static_cast<size_t>(kvh) * Rtot + r + (REDUCE ? d.part_offset * stride : 0);
        const int P = REDUCE ? d.n_splits : d.n_parts_total;
        float mx = -INFINITY;
        for (int p = lane; p < P; p += 32) mx = fmaxf(mx, d.part_lse[p * stride + row0]);
        mx = warp_max(mx);
        float wsum = 0.f;
        if (mx != -INFINITY) {
            for (int p0 = 0; p0 < P; p0 += 32) {
                const int p = p0 + lane;
                const float lse = p < P ? d.part_lse[p * stride + row0] : -INFINITY;
                const float w_mine = lse == -INFINITY ? 0.f : __expf(lse - mx);
                wsum += w_mine;
                const int cnt = min(32, P - p0);
                const float* base = d.part_o + (p0 * stride + row0) * kD + lane * 4;
#pragma unroll 4
                for (int k = 0; k < cnt; ++k) {
                    const float w = __shfl_sync(0xffffffffu, w_mine, k);
                    // an empty slot (LSE = -inf) carries no O row: never read it (its memory may hold anything)
                    const float4 v = w != 0.f ? *reinterpret_cast<const float4*>(base + k * stride * kD) : make_float4(0.f, 0.f, 0.f, 0.f);
                    acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
                    acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
                }
            }
            wsum = warp_sum(wsum);
            const float inv = 1.f / wsum;
            acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
            lse_out = mx + logf(wsum);
        }
    }
    if (REDUCE) {
        const size_t row = static_cast<size_t>(kvh) * Rtot + r;
        *reinterpret_cast<float4*>(dst_o + row * kD + lane * 4) = acc;
        if (lane == 0) dst_lse[row] = lse_out;
        return;
    }
    uint2 pk;
    pk.x = Num<T>::pack(acc.x, acc.y);
    pk.y = Num<T>::pack(acc.z, acc.w);
    *reinterpret_cast<uint2*>(out) = pk;
}

// ---- multi-GPU: reduce + push to every peer / wait + combine (see include/longspec_b200.h, ls_xchg) ----------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Wait until a peer's flag shows `epoch`.  A peer that died or fell out of step must not hang this GPU for ever: after
// ~20 s of spinning (far beyond any legitimate wait - a whole 128k prefill block exchange takes microseconds) the kernel
// traps, which surfaces as a CUDA error on the host instead of a silent hang.
__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t epoch) {
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(flag) - epoch) < 0) {
        if (clock64() - t0 > 40000000000ll) __trap();
    }
}

// One warp per (packed row, kv head): merge this launch's own slots (like attn_combine_kernel<REDUCE>) and store the
// normalised fp32 row + LSE into block `rank` of EVERY rank's exchange buffer; the last CTA publishes the epoch.
__global__ void __launch_bounds__(256) attn_reduce_push_kernel(const ls_attn_desc d, const ls_xchg x) {
    const int G = d.n_heads / d.n_kv_heads;
    const int Rtot = d.q_rows * G;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(x.local_sync);
    const int par = epoch & 1;
    if (warp_global < Rtot * d.n_kv_heads) {
        const int kvh = warp_global / Rtot, r = warp_global - kvh * Rtot;
        const int i = r / G;
        const int n_q = min(*d.n_q_ptr, d.q_rows);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float lse_out = -INFINITY;
        if (i < n_q) {
            const size_t stride = static_cast<size_t>(d.n_kv_heads) * Rtot;        // between parts, in rows
            const size_t row0 = static_cast<size_t>(kvh) * Rtot + r + d.part_offset * stride;
            const int P = d.n_splits;
            float mx = -INFINITY;
            for (int p = lane; p < P; p += 32) mx = fmaxf(mx, d.part_lse[p * stride + row0]);
            mx = warp_max(mx);
            float wsum = 0.f;
            if (mx != -INFINITY) {
                for (int p0 = 0; p0 < P; p0 += 32) {
                    const int p = p0 + lane;
                    const float lse = p < P ? d.part_lse[p * stride + row0] : -INFINITY;
                    const float w_mine = lse == -INFINITY ? 0.f : __expf(lse - mx);
                    wsum += w_mine;
                    const int cnt = min(32, P - p0);
                    const float* base = d.part_o + (p0 * stride + row0) * kD + lane * 4;
#pragma unroll 4
                    for (int k = 0; k < cnt; ++k) {
                        const float w = __shfl_sync(0xffffffffu, w_mine, k);
                        // an empty slot (LSE = -inf) carries no O row: never read it (its memory may hold anything)
                        const float4 v = w != 0.f ? *reinterpret_cast<const float4*>(base + k * stride * kD) : make_float4(0.f, 0.f, 0.f, 0.f);
                        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
                        acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
                    }
                }
                wsum = warp_sum(wsum);
                const float inv = 1.f / wsum;
                acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
                lse_out = mx + logf(wsum);
            }
        }
        const size_t row = static_cast<size_t>(kvh) * Rtot + r;
        const size_t blk = static_cast<size_t>(par) * x.world + x.rank;
        for (int p = 0; p < x.world; ++p) {
            *reinterpret_cast<float4*>(x.peer_o[p] + blk * x.block_o + row * kD + lane * 4) = acc;
            if (lane == 0) x.peer_lse[p][blk * x.block_lse + row] = lse_out;
        }
    }
    // publish: every CTA's stores must be visible system-wide before the flag moves
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t ticket = atomicAdd(x.local_sync + 1, 1u);
        if (ticket == gridDim.x - 1) {
            x.local_sync[1] = 0;
            __threadfence_system();
            for (int p = 0; p < x.world; ++p) st_release_sys(x.peer_flags[p] + x.rank, epoch + 1);
        }
    }
}

// Wait until every source's block of the current epoch has arrived, merge the `world` blocks in rank order, write `out`.
template <typename T>
__global__ void __launch_bounds__(256) attn_combine_wait_kernel(const ls_attn_desc d, const ls_xchg x) {
    const int G = d.n_heads / d.n_kv_heads;
    const int Rtot = d.q_rows * G;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(x.local_sync);
    const int par = epoch & 1;
    if (threadIdx.x < x.world) {
        const uint32_t* flag = x.peer_flags[x.rank] + threadIdx.x;
        wait_flag(flag, epoch + 1);
    }
    __syncthreads();
    if (warp_global < Rtot * d.n_kv_heads) {
        const int kvh = warp_global / Rtot, r = warp_global - kvh * Rtot;
        const int i = r / G, g = r - i * G;
        const int n_q = min(*d.n_q_ptr, d.q_rows);
        T* out = reinterpret_cast<T*>(d.out) + (static_cast<size_t>(i) * d.n_heads + kvh * G + g) * kD + lane * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n_q) {
            const size_t row = static_cast<size_t>(kvh) * Rtot + r;
            const float* bo = x.peer_o[x.rank] + static_cast<size_t>(par) * x.world * x.block_o;
            const float* bl = x.peer_lse[x.rank] + static_cast<size_t>(par) * x.world * x.block_lse;
            const int P = x.world;                                    // <= 8: one LSE per lane
            const float lse = lane < P ? __ldcg(bl + lane * x.block_lse + row) : -INFINITY;
            const float mx = warp_max(lse);
            if (mx != -INFINITY) {
                const float w_mine = lse == -INFINITY ? 0.f : __expf(lse - mx);
                const float wsum = warp_sum(w_mine);
                for (int k = 0; k < P; ++k) {
                    const float w = __shfl_sync(0xffffffffu, w_mine, k);
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(bo + k * x.block_o + row * kD + lane * 4));
                    acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
                    acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
                }
                const float inv = 1.f / wsum;
                acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
            }
        }
        uint2 pk;
        pk.x = Num<T>::pack(acc.x, acc.y);
        pk.y = Num<T>::pack(acc.z, acc.w);
        *reinterpret_cast<uint2*>(out) = pk;
    }
    // the last CTA advances the epoch (the next kernel in the stream reads it)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t ticket = atomicAdd(x.local_sync + 2, 1u);
        if (ticket == gridDim.x - 1) {
            x.local_sync[2] = 0;
            x.local_sync[0] = epoch + 1;
        }
    }
}

// Generic small all-gather over the exchange buffer (same flags / epoch / parity protocol as the attention blocks).
__global__ void __launch_bounds__(256) peer_push_kernel(const uint4* __restrict__ src, int64_t n16, int64_t slot16, const ls_xchg x) {
    const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(x.local_sync);
    const int64_t slot = static_cast<int64_t>(epoch & 1) * x.world + x.rank;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n16; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint4 v = src[i];
        for (int p = 0; p < x.world; ++p) reinterpret_cast<uint4*>(x.peer_o[p])[slot * slot16 + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t ticket = atomicAdd(x.local_sync + 1, 1u);
        if (ticket == gridDim.x - 1) {
            x.local_sync[1] = 0;
            __threadfence_system();
            for (int p = 0; p < x.world; ++p) st_release_sys(x.peer_flags[p] + x.rank, epoch + 1);
        }
    }
}

__global__ void __launch_bounds__(256) peer_gather_wait_kernel(uint4* __restrict__ dst, int64_t n16, int64_t slot16, const ls_xchg x) {
    const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(x.local_sync);
    if (threadIdx.x < x.world) {
        const uint32_t* flag = x.peer_flags[x.rank] + threadIdx.x;
        wait_flag(flag, epoch + 1);
    }
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(x.peer_o[x.rank]) + static_cast<int64_t>(epoch & 1) * x.world * slot16;
    const int64_t total = static_cast<int64_t>(x.world) * n16;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / n16, o = i - r * n16;
        dst[i] = __ldcg(src + r * slot16 + o);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t ticket = atomicAdd(x.local_sync + 2, 1u);
        if (ticket == gridDim.x - 1) {
            x.local_sync[2] = 0;
            x.local_sync[0] = epoch + 1;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) peer_sum_wait_kernel(T* __restrict__ dst, int64_t n4, int64_t slot4, const ls_xchg x) {
    const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(x.local_sync);
    if (threadIdx.x < x.world) {
        const uint32_t* flag = x.peer_flags[x.rank] + threadIdx.x;
        wait_flag(flag, epoch + 1);
    }
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(x.peer_o[x.rank]) + static_cast<int64_t>(epoch & 1) * x.world * slot4;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float4 acc = __ldcg(src + i);
        for (int r = 1; r < x.world; ++r) {                       // rank order: the same sum on every rank
            const float4 v = __ldcg(src + r * slot4 + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        uint2 pk;
        pk.x = Num<T>::pack(acc.x, acc.y);
        pk.y = Num<T>::pack(acc.z, acc.w);
        reinterpret_cast<uint2*>(dst)[i] = pk;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t ticket = atomicAdd(x.local_sync + 2, 1u);
        if (ticket == gridDim.x - 1) {
            x.local_sync[2] = 0;
            x.local_sync[0] = epoch + 1;
        }
    }
}

static int validate(const ls_attn_desc* d) {
    LS_REQUIRE(d != nullptr, "ls_attn: null descriptor");
    LS_REQUIRE(d->q && d->out && d->k_pool && d->v_pool && d->page_table && d->kv_len_ptr && d->n_q_ptr,
               "ls_attn: null pointer in descriptor");
    LS_REQUIRE(d->part_o && d->part_lse, "ls_attn: null workspace");
    LS_REQUIRE(d->q_rows >= 1 && d->q_rows <= 256, "ls_attn: q_rows %d out of range [1,256]", d->q_rows);
    LS_REQUIRE(d->n_kv_heads >= 1 && d->n_heads % d->n_kv_heads == 0, "ls_attn: n_heads %d %% n_kv_heads %d != 0",
               d->n_heads, d->n_kv_heads);
    LS_REQUIRE(d->n_tail >= 0 && d->n_tail <= 64 * LS_MASK_WORDS, "ls_attn: n_tail %d out of range", d->n_tail);
    LS_REQUIRE(d->n_tail == 0 || d->tail_mask != nullptr, "ls_attn: n_tail > 0 needs tail_mask");
    LS_REQUIRE(d->dtype == LS_F16 || d->dtype == LS_BF16, "ls_attn: bad dtype %d", d->dtype);
    LS_REQUIRE(d->q_stride == 0 || (d->q_stride >= static_cast<int64_t>(d->n_heads) * 128 && d->q_stride % 8 == 0),
               "ls_attn: q_stride %lld must be 0 or >= n_heads * 128 and a multiple of 8 elements", static_cast<long long>(d->q_stride));
    LS_REQUIRE(d->n_splits >= 1 && d->part_offset >= 0 && d->part_offset + d->n_splits <= d->n_parts_total,
               "ls_attn: bad split/part configuration (%d splits, offset %d, total %d)", d->n_splits, d->part_offset,
               d->n_parts_total);
    return LS_OK;
}

}  // namespace ls

static int launch_partial_mma(const ls_attn_desc* d, int remainder, void* stream) {
    using namespace ls;
    int rc = validate(d);
    if (rc != LS_OK) return rc;
    const int G = d->n_heads / d->n_kv_heads;
    dim3 grid(d->n_splits, d->n_kv_heads, (d->q_rows * G + kBM - 1) / kBM);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static unsigned long long smem_ok_f16 = 0, smem_ok_bf16 = 0;     // one bit per device ordinal
    rc = d->dtype == LS_F16 ? ensure_dyn_smem(attn_partial_kernel<__half>, kSmemBytes, &smem_ok_f16, "ls_attn_partial")
                            : ensure_dyn_smem(attn_partial_kernel<__nv_bfloat16>, kSmemBytes, &smem_ok_bf16, "ls_attn_partial");
    if (rc) return rc;
    if (d->dtype == LS_F16)
        attn_partial_kernel<__half><<<grid, kWarps * 32, kSmemBytes, st>>>(*d, remainder);
    else
        attn_partial_kernel<__nv_bfloat16><<<grid, kWarps * 32, kSmemBytes, st>>>(*d, remainder);
    return check_launch("ls_attn_partial");
}

extern "C" int ls_attn_partial_mma(const ls_attn_desc* d, void* stream) { return launch_partial_mma(d, 0, stream); }
// masked remainder behind the tcgen05 kernel: last partial prefix tile, causal rows, tree-masked tail
extern "C" int ls_attn_partial_mma_remainder(const ls_attn_desc* d, void* stream) { return launch_partial_mma(d, 1, stream); }

extern "C" int ls_attn_combine(const ls_attn_desc* d, void* stream) {
    using namespace ls;
    int rc = validate(d);
    if (rc != LS_OK) return rc;
    const int G = d->n_heads / d->n_kv_heads;
    const int warps = d->q_rows * G * d->n_kv_heads;
    const int blocks = (warps + 7) / 8;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d->dtype == LS_F16)
        attn_combine_kernel<__half, false><<<blocks, 256, 0, st>>>(*d, nullptr, nullptr);
    else
        attn_combine_kernel<__nv_bfloat16, false><<<blocks, 256, 0, st>>>(*d, nullptr, nullptr);
    return check_launch("ls_attn_combine");
}

extern "C" int ls_attn_reduce(const ls_attn_desc* d, float* dst_o, float* dst_lse, void* stream) {
    using namespace ls;
    int rc = validate(d);
    if (rc != LS_OK) return rc;
    LS_REQUIRE(dst_o && dst_lse, "ls_attn_reduce: null destination");
    const int G = d->n_heads / d->n_kv_heads;
    const int warps = d->q_rows * G * d->n_kv_heads;
    const int blocks = (warps + 7) / 8;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    attn_combine_kernel<__nv_bfloat16, true><<<blocks, 256, 0, st>>>(*d, dst_o, dst_lse);   // dtype-independent (fp32 in, fp32 out)
    return check_launch("ls_attn_reduce");
}

static int validate_xchg(const ls_attn_desc* d, const ls_xchg* x, const char* what) {
    using namespace ls;
    LS_REQUIRE(x && x->local_sync && x->world >= 1 && x->world <= LS_MAX_PEERS && x->rank >= 0 && x->rank < x->world,
               "%s: bad exchange descriptor", what);
    const int G = d->n_heads / d->n_kv_heads;
    const int64_t rows = static_cast<int64_t>(d->n_kv_heads) * d->q_rows * G;
    LS_REQUIRE(x->block_o >= rows * 128 && x->block_lse >= rows, "%s: exchange blocks too small for %lld rows", what,
               static_cast<long long>(rows));
    for (int p = 0; p < x->world; ++p)
        LS_REQUIRE(x->peer_o[p] && x->peer_lse[p] && x->peer_flags[p], "%s: null peer pointer (rank %d)", what, p);
    return LS_OK;
}

extern "C" int ls_attn_reduce_push(const ls_attn_desc* d, const ls_xchg* x, void* stream) {
    using namespace ls;
    int rc = validate(d);
    if (rc != LS_OK) return rc;
    rc = validate_xchg(d, x, "ls_attn_reduce_push");
    if (rc != LS_OK) return rc;
    const int G = d->n_heads / d->n_kv_heads;
    const int warps = d->q_rows * G * d->n_kv_heads;
    attn_reduce_push_kernel<<<(warps + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(*d, *x);
    return check_launch("ls_attn_reduce_push");
}

extern "C" int ls_attn_combine_wait(const ls_attn_desc* d, const ls_xchg* x, void* stream) {
    using namespace ls;
    LS_REQUIRE(d && d->out && d->n_q_ptr && d->n_kv_heads >= 1 && d->n_heads % d->n_kv_heads == 0 && d->q_rows >= 1,
               "ls_attn_combine_wait: bad descriptor");
    LS_REQUIRE(d->dtype == LS_F16 || d->dtype == LS_BF16, "ls_attn_combine_wait: bad dtype %d", d->dtype);
    int rc = validate_xchg(d, x, "ls_attn_combine_wait");
    if (rc != LS_OK) return rc;
    const int G = d->n_heads / d->n_kv_heads;
    const int warps = d->q_rows * G * d->n_kv_heads;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d->dtype == LS_F16)
        attn_combine_wait_kernel<__half><<<(warps + 7) / 8, 256, 0, st>>>(*d, *x);
    else
        attn_combine_wait_kernel<__nv_bfloat16><<<(warps + 7) / 8, 256, 0, st>>>(*d, *x);
    return check_launch("ls_attn_combine_wait");
}

static int validate_gather(const void* p, int64_t bytes, int64_t slot_bytes, const ls_xchg* x, const char* what) {
    using namespace ls;
    LS_REQUIRE(p && x && x->local_sync && x->world >= 1 && x->world <= LS_MAX_PEERS && x->rank >= 0 && x->rank < x->world,
               "%s: bad arguments", what);
    LS_REQUIRE(bytes > 0 && bytes % 16 == 0 && slot_bytes % 16 == 0 && bytes <= slot_bytes && slot_bytes <= x->block_o * 4,
               "%s: %lld bytes / slots of %lld do not fit the exchange blocks (%lld bytes)", what, static_cast<long long>(bytes),
               static_cast<long long>(slot_bytes), static_cast<long long>(x->block_o * 4));
    for (int r = 0; r < x->world; ++r) LS_REQUIRE(x->peer_o[r] && x->peer_flags[r], "%s: null peer pointer (rank %d)", what, r);
    return LS_OK;
}

extern "C" int ls_peer_push(const void* src, int64_t bytes, int64_t slot_bytes, const ls_xchg* x, void* stream) {
    using namespace ls;
    int rc = validate_gather(src, bytes, slot_bytes, x, "ls_peer_push");
    if (rc) return rc;
    const int64_t n16 = bytes / 16;
    const int blocks = static_cast<int>(n16 / 256 + 1 < 148 ? n16 / 256 + 1 : 148);
    peer_push_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(src), n16, slot_bytes / 16, *x);
    return check_launch("ls_peer_push");
}

extern "C" int ls_peer_gather_wait(void* dst, int64_t bytes, int64_t slot_bytes, const ls_xchg* x, void* stream) {
    using namespace ls;
    int rc = validate_gather(dst, bytes, slot_bytes, x, "ls_peer_gather_wait");
    if (rc) return rc;
    const int64_t n16 = bytes / 16, total = n16 * x->world;
    const int blocks = static_cast<int>(total / 256 + 1 < 32 ? total / 256 + 1 : 32);
    peer_gather_wait_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<uint4*>(dst), n16, slot_bytes / 16, *x);
    return check_launch("ls_peer_gather_wait");
}

extern "C" int ls_peer_sum_wait(void* dst, int64_t n_elems, int32_t dtype, int64_t slot_bytes, const ls_xchg* x, void* stream) {
    using namespace ls;
    int rc = validate_gather(dst, n_elems * 4, slot_bytes, x, "ls_peer_sum_wait");
    if (rc) return rc;
    LS_REQUIRE(dtype == LS_F16 || dtype == LS_BF16, "ls_peer_sum_wait: bad dtype %d", dtype);
    const int64_t n4 = n_elems / 4;
    const int blocks = static_cast<int>(n4 / 256 + 1 < 148 ? n4 / 256 + 1 : 148);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == LS_F16)
        peer_sum_wait_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<__half*>(dst), n4, slot_bytes / 16, *x);
    else
        peer_sum_wait_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<__nv_bfloat16*>(dst), n4, slot_bytes / 16, *x);
    return check_launch("ls_peer_sum_wait");
}

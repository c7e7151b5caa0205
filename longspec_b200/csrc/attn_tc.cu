// tcgen05 / TMA / TMEM split-KV attention over the paged prefix + the tree-masked tail (ls_attn_partial).
//
// One CTA = (split, kv head) and ALL packed query rows of that kv head (up to 3 tiles of 128 rows,
// r = token * G + head-in-group), so every K/V byte is pulled from HBM once per kv head and reused by the
// whole GQA group and the whole candidate tree from shared memory.
//
//   warp 0              TMA producer: 64-key K and V tiles (2 x [64 x 128B] boxes each, SWIZZLE_128B) from the
//                       paged pool into a 2-4 stage shared-memory ring, completion on mbarriers
//   warp 1              MMA issuer (converged warp, one elected lane): S = Q K^T (M128 N64 K16 x8, SS: Q tile and
//                       K tile K-major in smem) into the TMEM score tile; O_r += P V (M128 N128 K16 x4, TS: P read
//                       from TMEM, V MN-major in smem) into per-row-tile TMEM accumulators.  Event-driven: a QK
//                       burst goes out as soon as S is free, a PV burst as soon as P is published.
//   warp 2              TMEM allocation (512 columns: S 64, P 2 x 32, O 3 x 128)
//   warps 4-7 / 8-11    two softmax warpgroups, alternating steps: tcgen05.ld the S tile (one row per thread),
//                       online softmax in the log2 domain with lazy rescaling (O in TMEM is only rescaled when a
//                       row maximum grows by more than 2^8), P rounded to dtype and tcgen05.st back to TMEM
// The masked remainder of the call (last partial prefix tile, causal rows, the tree-masked tail = this round's
// tokens) is appended to the tile sequence as "special tiles" whose softmax applies a per-row bit mask, so one
// launch covers the whole attention call.  The epilogue writes normalised partial O (fp32) and the natural-log
// LSE per (split, kv head, row); ls_attn_combine (attn_mma.cu) merges the splits (and, sequence-sharded, the ranks).
//
// Reference semantics: flash_attn_with_kvcache(q, K_Cache, V_Cache, cache_seqlens=cache_lens,
// return_softmax_lse=True) + tree_part_fwd + merge of llama.py:385-421, and llama_glide.py:265,297;
// restated in oracle/flash_ref.py and oracle/attention.py.
#include <cuda.h>
#include <limits.h>

#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "common.cuh"

namespace ls {
namespace tc {

constexpr int BN = 64;
constexpr int NR_MAX = 3;
constexpr int MAX_STAGES = 4;
constexpr int THREADS = 384;
constexpr int KV_STAGE_BYTES = 4 * 8192;           // K c0 | K c1 | V c0 | V c1   (64 keys)
constexpr int P_BYTES = 128 * 128;                 // one P tile: 128 rows x 64 keys
constexpr int STATE_BYTES = 3 * NR_MAX * 128 * 4;  // m_ref[384], l[384], version[384]
constexpr int BAR_BYTES = 256;
constexpr int SPECIAL_BYTES = 256;                 // special-tile table (12 x 16 B)
constexpr int SMEM_BYTES = 232448;                 // everything sm_100 allows per CTA (227 KB)
constexpr int SMEM_USABLE = SMEM_BYTES - 1024;     // minus worst-case 1 KB alignment of the dynamic base

constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_S = 0;                      // one 128 x 64 fp32 score tile
constexpr uint32_t COL_P = 64;                     // two 128 x 64 bf16/fp16 probability tiles (32 columns each)
constexpr uint32_t COL_O = 128;                    // 3 x 128

// Shared-memory plan (all regions 1 KB aligned for SWIZZLE_128B):
//   Q   row tiles 0,1: [2 k-chunks][128 rows][128 B]; row tile 2 keeps only 64 rows when it has <= 64 valid rows
//       (the M=128 MMA then reads 64 rows of whatever follows - those S rows belong to padding and are never used)
//   KV  ring of 2..4 stages of 64-key K and V tiles, as many as fit
//   (P never touches shared memory: the softmax warpgroups store it to TMEM and the PV MMA reads its A operand
//    from there - shared-memory bandwidth, 128 B/clk, is the scarce resource of this kernel)
constexpr int MAX_SPECIAL = 12;
struct SpecialTile {
    int tok;      // first LOCAL token of the tile (need not be 64-aligned for tail tiles)
    int n_valid;  // keys of the tile that exist (the rest of the 64-row TMA box is masked)
    int kind;     // 0 = prefix keys (causal limit applies), 1 = tail keys (tree bit mask applies)
    int base;     // kind 0: global index of key 0;  kind 1: tail index of key 0
};

struct Plan {
    int q2_rows, q_bytes, stages, off_kv, off_p, off_state, off_bar;
};
__host__ __device__ inline Plan make_plan(int R) {
    Plan p;
    const int NR = (R + 127) / 128;
    p.q2_rows = (NR == 3 && R - 256 <= 64) ? 64 : 128;
    p.q_bytes = NR < 3 ? NR * 32768 : 65536 + p.q2_rows * 256;
    const int fixed = p.q_bytes + STATE_BYTES + BAR_BYTES + SPECIAL_BYTES;
    int st = (SMEM_USABLE - fixed) / KV_STAGE_BYTES;
    p.stages = st > MAX_STAGES ? MAX_STAGES : st;
    p.off_kv = p.q_bytes;
    p.off_p = 0;
    p.off_state = p.off_kv + p.stages * KV_STAGE_BYTES;
    p.off_bar = p.off_state + STATE_BYTES;
    return p;
}

// P_FULL / P_FREE are indexed by step parity (= softmax warpgroup = P buffer).  Every barrier is waited on
// phase by phase by its waiter, so mbarrier parity waits can never alias.
enum { B_KV_FULL = 0, B_KV_EMPTY = 4, B_S_FULL = 8, B_S_FREE = 10, B_P_FULL = 12, B_P_FREE = 14, B_PV_DONE = 16, B_ALL_DONE = 19, B_COUNT = 20 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Descriptors are (lo, hi) register pairs; hi is identical for every operand of this kernel (SBO = 1024 B,
// descriptor version 1, SWIZZLE_128B), lo = (smem address >> 4) | (LBO >> 4) << 16, so stepping along K is
// one 32-bit add.  Issued by one elected lane of a fully converged warp (operands stay in uniform registers).
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t LBO_K = (16u >> 4) << 16;         // K-major operands (Q, K, P)
constexpr uint32_t LBO_V = (8192u >> 4) << 16;       // MN-major V: 64-column chunks are 8 KB apart
__device__ __forceinline__ void tc_mma2(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(idesc), "r"(acc)
        : "memory");
}
// A operand from TMEM (the P tile), B from shared memory
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(DESC_HI), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int fmt, int a_mn, int b_mn) {
    return (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
           (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
           (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// packed fp32x2 arithmetic (sm_100): halves the issue slots of the softmax inner loop
__device__ __forceinline__ uint64_t pack2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

#define LS_TMEM_LD32(taddr, v)                                                                                       \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"      \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                            \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),     \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),            \
                   "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),          \
                   "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),          \
                   "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                              \
                 : "r"(taddr))
#define LS_TMEM_ST32(taddr, v)                                                                                       \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"   \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"                               \
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),           \
                   "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),        \
                   "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),     \
                   "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),     \
                   "r"(v[31])                                                                                         \
                 : "memory")
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <typename T>
__global__ void __launch_bounds__(THREADS, 1)
attn_tc_kernel(const ls_attn_desc d, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
               int n_splits_tc, int fuse) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int split = blockIdx.x, kvh = blockIdx.y;
    const int G = d.n_heads / d.n_kv_heads;
    const int n_q = min(*d.n_q_ptr, d.q_rows);
    const int R = n_q * G;
    const int NR = (R + 127) / 128;                         // row tiles in use (1..3)
    const Plan pl = make_plan(R);
    const int STAGES = pl.stages;
    const int kv_len = *d.kv_len_ptr + d.kv_len_bias;
    const int total_tiles = tc_full_tiles(d, kv_len, n_q);
    // With `fuse` the masked remainder rides along as SPECIAL tiles appended to the tile sequence: the prefix keys
    // behind the last full unmasked tile (partial tile, causal rows) and the tree-masked tail of this round's
    // tokens.  They come from a small table in shared memory and the softmax applies a per-row bit mask to them.
    SpecialTile* sp_tiles = reinterpret_cast<SpecialTile*>(smem + pl.off_bar + BAR_BYTES);
    int n_special = 0;
    if (fuse) {
        const int sb = d.kv_shard_begin;
        const int se = d.kv_shard_end < 0 ? INT_MAX : d.kv_shard_end;
        const int loc_end = max(min(kv_len, se) - sb, 0);
        int n = 0;
        for (int tok = total_tiles * BN; tok < loc_end && n < MAX_SPECIAL; tok += BN, ++n)
            if (tid == 0) sp_tiles[n] = SpecialTile{tok, min(BN, loc_end - tok), 0, tok + sb};
        if (d.tail_owner) {
            const int t0 = kv_len - sb, t1 = t0 + d.n_tail;
            for (int tok = t0; tok < t1 && n < MAX_SPECIAL; ++n) {
                const int len = min(min(BN, LS_PAGE - (tok & (LS_PAGE - 1))), t1 - tok);   // never cross a page
                if (tid == 0) sp_tiles[n] = SpecialTile{tok, len, 1, tok - t0};
                tok += len;
            }
        }
        n_special = n;
    }
    const int all_tiles = total_tiles + n_special;
    const int per = (all_tiles + n_splits_tc - 1) / n_splits_tc;
    const int tile0 = min(split * per, all_tiles);
    const int n_tiles = min(tile0 + per, all_tiles) - tile0;
    const int first_special = total_tiles - tile0;           // local tile index j >= first_special is special
    const int n_steps = n_tiles * NR;

    float* st_m = reinterpret_cast<float*>(smem + pl.off_state);
    float* st_l = st_m + NR_MAX * 128;
    uint32_t* st_ver = reinterpret_cast<uint32_t*>(st_l + NR_MAX * 128);   // key tiles already folded into (m, l) of a row
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + pl.off_bar + B_COUNT * 8);
    const uint32_t bar0 = sbase + pl.off_bar;
    auto bar = [&](int i) { return bar0 + i * 8; };

    // ---- one-time setup ----------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < MAX_STAGES; ++i) { mbar_init(bar(B_KV_FULL + i), 1); mbar_init(bar(B_KV_EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar(B_S_FULL + i), 1); mbar_init(bar(B_S_FREE + i), 128); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar(B_P_FULL + i), 128); mbar_init(bar(B_P_FREE + i), 1); }
        for (int i = 0; i < NR_MAX; ++i) mbar_init(bar(B_PV_DONE + i), 1);
        mbar_init(bar(B_ALL_DONE), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < NR_MAX * 128; i += THREADS) { st_m[i] = -INFINITY; st_l[i] = 0.f; st_ver[i] = 0u; }
    // Q tiles -> swizzled shared memory (K-major, SWIZZLE_128B): [tile][k-chunk][row][128B]
    {
        const T* __restrict__ q = reinterpret_cast<const T*>(d.q);
        const int rows_total = NR < 3 ? NR * 128 : 256 + pl.q2_rows;
        for (int idx = tid; idx < rows_total * 16; idx += THREADS) {
            const int rp = idx >> 4, chunk = idx & 15;
            const int tile = rp >> 7, row = rp & 127;
            const int chunk_stride = tile < 2 ? 16384 : pl.q2_rows * 128;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (rp < R) {
                const int i = rp / G, g = rp - i * G;
                val = *reinterpret_cast<const uint4*>(q + (static_cast<size_t>(i) * d.n_heads + kvh * G + g) * 128 + chunk * 8);
            }
            uint8_t* dst = smem + tile * 32768 + (chunk >> 3) * chunk_stride + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = val;
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    constexpr int FMT = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
    constexpr uint32_t IDESC_QK = make_idesc(128, BN, FMT, 0, 0);
    constexpr uint32_t IDESC_PV = make_idesc(128, 128, FMT, 0, 1);

    if (warp == 0) {
        // ================= TMA producer (warp converged, one elected lane issues) =================
        int s = 0, ph = 0;
        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(bar(B_KV_EMPTY + s), ph ^ 1);
            if (elect_one()) {
                const int tok = j >= first_special ? sp_tiles[j - first_special].tok : (tile0 + j) * BN;
                const int page = __ldg(d.page_table + (tok >> 7));
                const int row = (page * d.n_kv_heads + kvh) * LS_PAGE + (tok & (LS_PAGE - 1));
                const uint32_t dst = sbase + pl.off_kv + s * KV_STAGE_BYTES;
                mbar_expect_tx(bar(B_KV_FULL + s), KV_STAGE_BYTES);
                tma_load_2d(dst, &map_k, 0, row, bar(B_KV_FULL + s));
                tma_load_2d(dst + 8192, &map_k, 64, row, bar(B_KV_FULL + s));
                tma_load_2d(dst + 16384, &map_v, 0, row, bar(B_KV_FULL + s));
                tma_load_2d(dst + 24576, &map_v, 64, row, bar(B_KV_FULL + s));
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp runs this loop converged (all values warp-uniform); one elected lane issues.
        // Event-driven: QK(t) goes out as soon as its S buffer is free (softmax(t-2) has LOADED it) and its K tile
        // has landed; PV(t) as soon as softmax(t) published P.  Neither waits behind the other.
        const uint32_t q_lo = (sbase >> 4) | LBO_K;
        const uint32_t k_lo = ((sbase + pl.off_kv) >> 4) | LBO_K;
        const uint32_t v_lo = ((sbase + pl.off_kv + 16384) >> 4) | LBO_V;
        const uint32_t q2_chunk = (pl.q2_rows * 128) >> 4;
        int tq = 0, rq = 0, sq = 0, phq = 0;          // next QK step: row tile, kv stage, kv_full parity
        int tp = 0, rp = 0, sp = 0, jp0 = 1;          // next PV step: row tile, kv stage, first-tile flag
        bool kv_ready = false;
        while (tp < n_steps) {
            if (tq < n_steps) {
                const int b = tq & 1;
                if (rq == 0 && !kv_ready) kv_ready = mbar_test(bar(B_KV_FULL + sq), phq);
                // single S tile: QK(t) may overwrite it once softmax(t-1) (the other warpgroup) has loaded it
                if ((rq != 0 || kv_ready) && (tq == 0 || mbar_test(bar(B_S_FREE + (b ^ 1)), ((tq - 1) >> 1) & 1))) {
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a0 = q_lo + rq * (32768 >> 4), b0 = k_lo + sq * (KV_STAGE_BYTES >> 4);
                        const uint32_t a_chunk = rq < 2 ? (16384u >> 4) : q2_chunk;
                        const uint32_t dst = tmem + COL_S;
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
                            tc_mma2(dst, a0 + (kk >> 2) * a_chunk + (kk & 3) * 2, b0 + (kk >> 2) * (8192 >> 4) + (kk & 3) * 2,
                                    IDESC_QK, kk > 0 ? 1u : 0u);
                        tc_commit(bar(B_S_FULL + b));
                    }
                    __syncwarp();
                    ++tq;
                    if (++rq == NR) {
                        rq = 0;
                        kv_ready = false;
                        if (++sq == STAGES) { sq = 0; phq ^= 1; }
                    }
                }
            }
            if (tp < tq && mbar_test(bar(B_P_FULL + (tp & 1)), (tp >> 1) & 1)) {
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a0 = tmem + COL_P + (tp & 1) * 32, b0 = v_lo + sp * (KV_STAGE_BYTES >> 4);
                    const uint32_t dst = tmem + COL_O + rp * 128;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)      // 16 keys = 8 packed TMEM columns per K step
                        tc_mma_ts(dst, a0 + kk * 8, b0 + kk * (2048 >> 4), IDESC_PV, (!jp0 || kk > 0) ? 1u : 0u);
                    tc_commit(bar(B_P_FREE + (tp & 1)));
                    tc_commit(bar(B_PV_DONE + rp));
                    if (rp == NR - 1) tc_commit(bar(B_KV_EMPTY + sp));
                }
                __syncwarp();
                ++tp;
                if (++rp == NR) {
                    rp = 0;
                    jp0 = 0;
                    if (++sp == STAGES) sp = 0;
                }
            }
        }
        if (n_steps > 0 && elect_one()) tc_commit(bar(B_ALL_DONE));   // every MMA of this CTA has completed once this fires
        __syncwarp();
    } else if (warp >= 4) {
        // ================= softmax warpgroups =================
        const int wg = (warp - 4) >> 2;                              // = step parity = S buffer = P buffer
        const int quad = warp & 3;                                   // TMEM lane quarter of this warp
        const int row = quad * 32 + lane;                            // row inside a 128-row tile
        const uint32_t lane_addr = tmem + (static_cast<uint32_t>(quad * 32) << 16);
        const float c2 = d.scale * 1.4426950408889634f;
        const uint64_t c2x2 = pack2(c2, c2);
        const int b = wg;
        int j = 0, r = wg;                                           // (key tile, row tile) of step t, kept without divisions
        for (int t = wg; t < n_steps; t += 2, r += 2) {
            while (r >= NR) { r -= NR; ++j; }
            const int u = t >> 1;                                     // use count of this warpgroup's buffers
            const bool warp_valid = r * 128 + quad * 32 < R;         // warp-uniform: any valid row in this warp
            if (!warp_valid) {
                // padding rows only: keep the barrier protocol, skip the math (their O rows are never read)
                mbar_wait(bar(B_S_FULL + b), u & 1);
                mbar_arrive(bar(B_S_FREE + b));
                mbar_wait(bar(B_P_FREE + b), (u & 1) ^ 1);
                mbar_arrive(bar(B_P_FULL + b));
                continue;
            }
            mbar_wait(bar(B_S_FULL + b), u & 1);
            tc_fence_after();
            uint32_t sv[64];
            LS_TMEM_LD32(lane_addr + COL_S, sv);
            LS_TMEM_LD32(lane_addr + COL_S + 32, (sv + 32));
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(bar(B_S_FREE + b));
            if (j >= first_special) {
                // per-row visibility word of this tile's 64 key slots
                const SpecialTile st = sp_tiles[j - first_special];
                const int rp = r * 128 + row;
                uint64_t vis = 0ull;
                if (rp < R) {
                    const int qi = rp / G;
                    if (st.kind == 0) {
                        const int hi = d.causal ? kv_len - (n_q - 1 - qi) : kv_len;      // visible global keys: j < hi
                        const int cnt = min(max(hi - st.base, 0), st.n_valid);
                        vis = cnt >= 64 ? ~0ull : ((1ull << cnt) - 1ull);
                    } else {
                        const uint64_t* w = d.tail_mask + static_cast<size_t>(qi) * LS_MASK_WORDS;
                        const int wi = st.base >> 6, sh = st.base & 63;
                        vis = w[wi] >> sh;
                        if (sh && wi + 1 < LS_MASK_WORDS) vis |= w[wi + 1] << (64 - sh);
                        if (st.n_valid < 64) vis &= (1ull << st.n_valid) - 1ull;
                    }
                }
#pragma unroll
                for (int e = 0; e < 64; ++e)
                    if (!((vis >> e) & 1ull)) sv[e] = 0xff800000u;                       // -inf
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int e = 0; e < 64; e += 8) {
                mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])));
                mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])));
                mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sv[e + 4]), __uint_as_float(sv[e + 5])));
                mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sv[e + 6]), __uint_as_float(sv[e + 7])));
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c2;
            // The row state of (r, j-1) was written by the OTHER warpgroup when NR is odd.  Hand it over through a
            // per-row version counter (release/acquire in shared memory): exact, and free of mbarrier aliasing.
            {
                const uint32_t va = smem_u32(st_ver + r * 128 + row);
                uint32_t ver;
                do {
                    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(ver) : "r"(va) : "memory");
                } while (ver != static_cast<uint32_t>(j));
            }
            float m_ref = st_m[r * 128 + row], l = st_l[r * 128 + row];
            const bool need = mx > m_ref + 8.0f;                     // lazy rescale threshold (2^8)
            if (__any_sync(0xffffffffu, need)) {
                const float m_new = need ? mx : m_ref;
                const float alpha = fast_exp2(m_ref - m_new);        // -inf - x -> 0 on the first tile
                l *= alpha;
                m_ref = m_new;
                if (j > 0) {
                    // O_r is quiescent once PV(r, j-1) completed.  First make sure PV(t-2) is complete (our own
                    // P_FREE phase) so that PV_DONE[r] is at most one phase behind and the parity wait cannot alias.
                    mbar_wait(bar(B_P_FREE + b), (u & 1) ^ 1);
                    mbar_wait(bar(B_PV_DONE + r), (j - 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t ov[32];
                        LS_TMEM_LD32(lane_addr + COL_O + r * 128 + c * 32, ov);
                        tmem_wait_ld();
#pragma unroll
                        for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
                        LS_TMEM_ST32(lane_addr + COL_O + r * 128 + c * 32, ov);
                    }
                    tmem_wait_st();
                    tc_fence_before();
                }
            }
            // p = 2^(s * c2 - m_ref), packed fp32x2 FMA / ADD, four independent partial sums
            const float m_use = m_ref == -INFINITY ? 0.f : m_ref;      // row with no visible key so far: p = 2^(-inf) = 0
            const uint64_t negm = pack2(-m_use, -m_use);
            uint64_t sum_a = 0ull, sum_b = 0ull;                     // {0.f, 0.f}
            uint32_t pk[32];
#pragma unroll
            for (int e = 0; e < 64; e += 4) {
                float x0, x1, x2, x3;
                unpack2(fma2(pack2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), c2x2, negm), x0, x1);
                unpack2(fma2(pack2(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])), c2x2, negm), x2, x3);
                const float p0 = fast_exp2(x0), p1 = fast_exp2(x1), p2 = fast_exp2(x2), p3 = fast_exp2(x3);
                sum_a = add2(sum_a, pack2(p0, p1));
                sum_b = add2(sum_b, pack2(p2, p3));
                pk[e >> 1] = Num<T>::pack(p0, p1);
                pk[(e >> 1) + 1] = Num<T>::pack(p2, p3);
            }
            {
                float a0, a1, b0, b1;
                unpack2(add2(sum_a, sum_b), a0, a1);
                (void)b0; (void)b1;
                l += a0 + a1;
            }
            st_m[r * 128 + row] = m_ref;
            st_l[r * 128 + row] = l;
            asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(st_ver + r * 128 + row)), "r"(static_cast<uint32_t>(j + 1)) : "memory");
            // P tile -> this warpgroup's TMEM buffer (A operand of the PV MMA, 2 values per 32-bit column);
            // PV(t-2) must be done with it
            mbar_wait(bar(B_P_FREE + b), (u & 1) ^ 1);
            tc_fence_after();
            LS_TMEM_ST32(lane_addr + COL_P + b * 32, pk);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(bar(B_P_FULL + b));
        }
        // ================= epilogue: normalised partial O + LSE =================
        const int Rtot = d.q_rows * G;
        const int part = d.part_offset + split;
        for (int re = wg; re < NR; re += 2) {
            const int rp = re * 128 + row;
            if (n_tiles > 0) {
                mbar_wait(bar(B_ALL_DONE), 0);       // single-phase barrier: no parity aliasing
                tc_fence_after();
            }
            const float l = st_l[rp], m_ref = st_m[rp];
            const float inv = (n_tiles > 0 && l > 0.f) ? 1.f / l : 0.f;
            const size_t base = (static_cast<size_t>(part) * d.n_kv_heads + kvh) * Rtot + rp;
            if (rp < R) d.part_lse[base] = inv > 0.f ? (m_ref + log2f(l)) * 0.6931471805599453f : -INFINITY;
            if (re * 128 + quad * 32 < R) {               // warp-uniform (tcgen05.ld is warp-collective)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t ov[32];
                    if (n_tiles > 0) {
                        LS_TMEM_LD32(lane_addr + COL_O + re * 128 + c * 32, ov);
                        tmem_wait_ld();
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e) ov[e] = 0u;
                    }
                    if (rp < R) {
                        float* po = d.part_o + base * 128 + c * 32;
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            *reinterpret_cast<float4*>(po + e) =
                                make_float4(__uint_as_float(ov[e]) * inv, __uint_as_float(ov[e + 1]) * inv,
                                            __uint_as_float(ov[e + 2]) * inv, __uint_as_float(ov[e + 3]) * inv);
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

// ---- host side: tensor maps (cached per pool) -----------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct MapKey {
    const void* ptr; long long rows; int dtype;
    bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && dtype == o.dtype; }
};
struct MapHash {
    size_t operator()(const MapKey& k) const { return std::hash<const void*>()(k.ptr) ^ (std::hash<long long>()(k.rows) << 1) ^ k.dtype; }
};

static int get_map(const void* pool, long long rows, int dtype, CUtensorMap* out) {
    static std::unordered_map<MapKey, CUtensorMap, MapHash> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    MapKey key{pool, rows, dtype};
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return LS_OK; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LS_ERR_CUDA; }
    CUtensorMap m;
    cuuint64_t gdim[2] = {128, static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {256};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = fn(&m, dtype == LS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                     const_cast<void*>(pool), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", static_cast<int>(rc)); return LS_ERR_CUDA; }
    cache[key] = m;
    *out = m;
    return LS_OK;
}

}  // namespace tc
}  // namespace ls

// Launch the tensor-core kernel over the unmasked full tiles with `n_splits_tc` splits writing partial slots
// [part_offset, part_offset + n_splits_tc).  With `fuse_remainder` the masked remainder (partial tile, causal
// rows, tree-masked tail) is appended to the tile sequence as special tiles and shared out with the rest.
static int launch_tc(const ls_attn_desc* d, int n_splits_tc, int fuse_remainder, void* stream) {
    using namespace ls;
    using namespace ls::tc;
    LS_REQUIRE(d && n_splits_tc >= 1, "ls_attn_partial_tc: bad arguments");
    const int G = d->n_heads / d->n_kv_heads;
    LS_REQUIRE(d->q_rows * G <= NR_MAX * 128, "ls_attn_partial_tc: %d packed rows exceed %d", d->q_rows * G, NR_MAX * 128);
    LS_REQUIRE(d->window < 0, "ls_attn_partial_tc: windowed attention stays on the general path");
    LS_REQUIRE(d->n_pages > 0, "ls_attn_partial_tc: n_pages must be set");
    CUtensorMap mk, mv;
    const long long rows = static_cast<long long>(d->n_pages) * d->n_kv_heads * LS_PAGE;
    int rc = get_map(d->k_pool, rows, d->dtype, &mk);
    if (rc) return rc;
    rc = get_map(d->v_pool, rows, d->dtype, &mv);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(attn_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        cudaFuncSetAttribute(attn_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        attr_set = true;
    }
    dim3 grid(n_splits_tc, d->n_kv_heads);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d->dtype == LS_F16)
        attn_tc_kernel<__half><<<grid, THREADS, SMEM_BYTES, st>>>(*d, mk, mv, n_splits_tc, fuse_remainder);
    else
        attn_tc_kernel<__nv_bfloat16><<<grid, THREADS, SMEM_BYTES, st>>>(*d, mk, mv, n_splits_tc, fuse_remainder);
    return check_launch("ls_attn_partial_tc");
}

extern "C" int ls_attn_partial_tc(const ls_attn_desc* d, int n_splits_tc, void* stream) { return launch_tc(d, n_splits_tc, 0, stream); }
// the whole call in one launch: d->n_splits CTAs per kv head share the unmasked and the special tiles
extern "C" int ls_attn_partial_tc_fused(const ls_attn_desc* d, void* stream) {
    LS_REQUIRE(d && d->n_splits >= 1, "ls_attn_partial_tc_fused: needs n_splits >= 1");
    return launch_tc(d, d->n_splits, 1, stream);
}

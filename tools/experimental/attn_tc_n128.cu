// tcgen05 / TMA / TMEM split-KV attention over the paged prefix + the tree-masked tail (ls_attn_partial).
//
// Work unit: one CTA = (kv head, a share of the 128-key tiles, up to TWO row tiles of 128 packed query rows),
// a packed row being r = token * G + head-in-group, so a K/V tile serves the whole GQA group and candidate tree.
//
// EXPERIMENTAL VARIANT (tools/experimental, not part of the default build; round-2 candidate, UNTESTED on hardware):
// TMEM (512 columns) = two 128-column pair buffers + ONE 128-column O accumulator per row tile.  A PAIR is (key tile,
// row tile): one QK^T with N = 128 keys (tensor-rate instead of the shared-memory-bound N = 64: 8 KB of operands per
// MMA for twice the work of 6 KB), the two softmax warpgroups turn the two 64-key halves into P in place, one PV with
// K = 128 keys.  Both warpgroups feed the same accumulator, so the two threads of a row share the reference maximum;
// a bar.red.or per lane quarter decides whether any row has to move it (lazy rescale), see the softmax section.
// Per (row tile, 128 keys) the CTA moves 128 KB through shared memory instead of 160 KB and issues 16 MMAs instead of 24.
// A call with three row tiles (the Llama-3 verify shape: 74 tokens x GQA-4 = 296 rows) is cut into "trios": two
// HEAVY CTAs (row tiles 0-1, alternate key tiles of the group's range) and one LIGHT CTA (row tile 2, every key
// tile of the range).  All three run the same number of steps, and the light CTA's K/V loads hit L2 behind the
// heavy CTAs' loads, so HBM still sees every byte once (ncu: dram bytes = algorithmic bytes).
//
//   warp 0              TMA producer: 128-key K and V tiles (2 x [128 x 128 B] boxes each, SWIZZLE_128B) from the
//                       paged pool into separate 2-stage K and V rings, completion on mbarriers
//   warp 1              MMA issuer (converged warp, one elected lane): S = Q K^T (M128 N64 K16 x8, SS) and
//                       O += P V (M128 N128 K16 x4, TS: P read from TMEM, V MN-major in smem), issued in PAIRS of
//                       steps:  QK(pair 0) QK(pair 1) | P(0): PV(0) QK(2) | P(1): PV(1) QK(3) | ...
//                       MMAs of one thread execute in issue order, so QK(pair p+2) cannot overwrite the buffers
//                       PV(pair p) still reads: no "buffer free" barriers.  tcgen05.mma issue is throttled to the
//                       pipe's execution rate, so the issuer does one wait and one commit per 24 MMAs.
//   warp 2              TMEM allocation
//   warps 4-7 / 8-11    two softmax warpgroups, one step of every pair each.  One row per thread: tcgen05.ld the
//                       64-key S tile, online softmax in the log2 domain with lazy rescaling (O in TMEM is only
//                       rescaled when a row maximum grows by more than 2^8), P rounded to dtype and tcgen05.st IN
//                       PLACE over the first 32 columns of the S buffer it came from (S and P alias).
// The masked remainder of the call (last partial prefix tile, causal rows, the tree-masked tail = this round's
// tokens) is appended to the tile sequence as "special tiles" whose softmax applies a per-row bit mask, so one
// launch covers the whole attention call.  The epilogue writes normalised partial O (fp32) and the natural-log
// LSE per (slot, kv head, row); ls_attn_combine (attn_mma.cu) merges the slots (and, sequence-sharded, the ranks).
//
// Reference semantics: flash_attn_with_kvcache(q, K_Cache, V_Cache, cache_seqlens=cache_lens,
// return_softmax_lse=True) + tree_part_fwd + merge of llama.py:385-421, and llama_glide.py:265,297;
// restated in oracle/flash_ref.py and oracle/attention.py.
#include <cuda.h>
#include <limits.h>

#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "../../longspec_b200/csrc/common.cuh"

namespace ls {
namespace tc {

constexpr int BN = 128;                            // keys per tile
constexpr int NRL_MAX = 2;                         // row tiles per CTA
constexpr int NR_MAX = 3;                          // row tiles per call
constexpr int STAGES = 2;                          // K ring and V ring depth
constexpr int THREADS = 384;
constexpr int TILE_BYTES = 32768;                  // one K or V tile: [2 x 64-element chunks][128 keys][128 B]
constexpr int Q_BYTES = NRL_MAX * 32768;           // [row tile][2 k-chunks][128 rows][128 B]
constexpr int STATE_BYTES = 2 * 2 * 128 * 4;       // exchange scratch between the warpgroups (maxima; row sums)
constexpr int BAR_BYTES = 256;
constexpr int SPECIAL_BYTES = 256;                 // special-tile table (12 x 16 B)
constexpr int OFF_K = Q_BYTES;
constexpr int OFF_V = OFF_K + STAGES * TILE_BYTES;
constexpr int OFF_STATE = OFF_V + STAGES * TILE_BYTES;
constexpr int OFF_BAR = OFF_STATE + STATE_BYTES;
constexpr int OFF_SPECIAL = OFF_BAR + BAR_BYTES;
constexpr int SMEM_BYTES = OFF_SPECIAL + SPECIAL_BYTES + 1024;   // + worst-case alignment of the dynamic base

constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_SP = 0;                     // four buffers of 64 columns: a 128 x 64 fp32 score tile, overwritten
                                                   // in place by its 128 x 64 bf16/fp16 probability tile (32 columns)
constexpr uint32_t COL_O = 256;                    // 2 x 128

constexpr int MAX_SPECIAL = 12;
struct SpecialTile {
    int tok;      // first LOCAL token of the tile (need not be 128-aligned for tail tiles)
    int n_valid;  // keys of the tile that exist (the rest of the 128-row TMA box is masked)
    int kind;     // 0 = prefix keys (causal limit applies), 1 = tail keys (tree bit mask applies)
    int base;     // kind 0: global index of key 0;  kind 1: tail index of key 0
};

// S_FULL / P_FULL are indexed by pair parity (pair p = steps 2p, 2p+1).  Every barrier is waited on phase by phase
// by its waiter (or provably at most one phase behind), so mbarrier parity waits can never alias.
enum { B_K_FULL = 0, B_K_EMPTY = 2, B_V_FULL = 4, B_V_EMPTY = 6, B_S_FULL = 8, B_P_FULL = 10, B_ALL_DONE = 12, B_COUNT = 13 };

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
constexpr uint32_t LBO_V = (16384u >> 4) << 16;      // MN-major V: 64-column chunks are 16 KB apart
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

// Optional event trace of CTA (0, 0) for tools/trace_attn.py (compiled in only with -DLS_TC_TRACE)
#ifdef LS_TC_TRACE
__device__ int g_dbg = 0;     // timing experiments (results become wrong): 1 no PV MMAs, 2 no QK MMAs, 4 no softmax math
#define LS_DBG(bit) (dbg & (bit))
__device__ long long* g_trace = nullptr;
#define LS_TRACE_INIT long long* const trace_buf = (blockIdx.x == 0 && blockIdx.y == 0) ? g_trace : nullptr; const int dbg = g_dbg;
#define LS_TRACE(ev, t)                                                       \
    do {                                                                      \
        if (trace_buf && (t) < 512) trace_buf[(ev) * 512 + (t)] = clock64();  \
    } while (0)
#else
#define LS_TRACE_INIT
#define LS_TRACE(ev, t) do { } while (0)
#define LS_DBG(bit) 0
#endif

template <typename T>
__global__ void __launch_bounds__(THREADS, 1)
attn_tc_kernel(const ls_attn_desc d, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
               int n_slots, int fuse) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bx = blockIdx.x, kvh = blockIdx.y;
    LS_TRACE_INIT
    const int G = d.n_heads / d.n_kv_heads;
    const int n_q = min(*d.n_q_ptr, d.q_rows);
    const int R = n_q * G;
    const int kv_len = *d.kv_len_ptr + d.kv_len_bias;
    const int total_tiles = tc_full_tiles(d, kv_len, n_q);
    // With `fuse` the masked remainder rides along as SPECIAL tiles appended to the tile sequence: the prefix keys
    // behind the last full unmasked tile (partial tile, causal rows) and the tree-masked tail of this round's
    // tokens.  They come from a small table in shared memory and the softmax applies a per-row bit mask to them.
    SpecialTile* sp_tiles = reinterpret_cast<SpecialTile*>(smem + OFF_SPECIAL);
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
                const int len = min(LS_PAGE - (tok & (LS_PAGE - 1)), t1 - tok);   // never cross a page
                if (tid == 0) sp_tiles[n] = SpecialTile{tok, len, 1, tok - t0};
                tok += len;
            }
        }
        n_special = n;
    }
    const int all_tiles = total_tiles + n_special;

    // ---- role of this CTA (see the header comment) -------------------------------------------------
    // NRS row tiles are planned from the STATIC q_rows; the rows that exist at run time (R) only gate the math.
    const int NRS = (d.q_rows * G + 127) / 128;
    int slot = bx, rt0 = 0, share0 = bx, nshare = 1, first = 0, empty_slot = -1;
    if (NRS > NRL_MAX) {
        const bool light = bx >= n_slots;
        const int g = light ? bx - n_slots : bx >> 1;
        share0 = 2 * g;
        nshare = min(2, n_slots - share0);
        if (light) { slot = share0; rt0 = NRL_MAX; empty_slot = nshare == 2 ? share0 + 1 : -1; }
        else { first = bx & 1; }
    }
    const int stride = (NRS > NRL_MAX && rt0 == 0) ? nshare : 1;
    const int NR = max(min((R + 127) / 128 - rt0, NRL_MAX), 0);      // local row tiles with at least one live row
    const int per = (all_tiles + n_slots - 1) / n_slots;
    const int tile_lo = min(share0 * per, all_tiles) + first;
    const int tile_hi = min((share0 + nshare) * per, all_tiles);
    const int n_tiles = (NR > 0 && tile_hi > tile_lo) ? (tile_hi - tile_lo + stride - 1) / stride : 0;
    // A pair = (key tile j, row tile r), r fastest; pair p runs in pair buffer p & 1 and accumulates into O_r.
    const int n_steps = n_tiles * NR;                                // pairs of this CTA (name kept for the shared set-up code)
    auto tile_index = [&](int j) { return tile_lo + j * stride; };   // position in the call's tile sequence

    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + B_COUNT * 8);
    const uint32_t bar0 = sbase + OFF_BAR;
    auto bar = [&](int i) { return bar0 + i * 8; };

    // ---- one-time setup ----------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(bar(B_K_FULL + i), 1); mbar_init(bar(B_K_EMPTY + i), 1);
            mbar_init(bar(B_V_FULL + i), 1); mbar_init(bar(B_V_EMPTY + i), 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(bar(B_S_FULL + i), 1); mbar_init(bar(B_P_FULL + i), 256); }
        mbar_init(bar(B_ALL_DONE), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // Q tiles -> swizzled shared memory (K-major, SWIZZLE_128B): [tile][k-chunk][row][128B]
    if (n_steps > 0) {
        const T* __restrict__ q = reinterpret_cast<const T*>(d.q);
        const size_t q_stride = d.q_stride > 0 ? static_cast<size_t>(d.q_stride) : static_cast<size_t>(d.n_heads) * 128;
        for (int idx = tid; idx < NR * 128 * 16; idx += THREADS) {
            const int rl = idx >> 4, chunk = idx & 15;
            const int tile = rl >> 7, row = rl & 127;
            const int rp = rt0 * 128 + rl;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (rp < R) {
                const int i = rp / G, g = rp - i * G;
                val = *reinterpret_cast<const uint4*>(q + static_cast<size_t>(i) * q_stride + static_cast<size_t>(kvh * G + g) * 128 + chunk * 8);
            }
            uint8_t* dst = smem + tile * 32768 + (chunk >> 3) * 16384 + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = val;
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    constexpr int FMT = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
    constexpr uint32_t IDESC_QK = make_idesc(128, 128, FMT, 0, 0);
    constexpr uint32_t IDESC_PV = make_idesc(128, 128, FMT, 0, 1);

    if (warp == 0) {
        // ================= TMA producer (warp converged, one elected lane issues) =================
        int s = 0, ph = 0;
        for (int j = 0; j < n_tiles && !LS_DBG(16); ++j) {
            const int gi = tile_index(j);
            const int tok = gi >= total_tiles ? sp_tiles[gi - total_tiles].tok : gi * BN;
            const int page = __ldg(d.page_table + (tok >> 7));
            const int row = (page * d.n_kv_heads + kvh) * LS_PAGE + (tok & (LS_PAGE - 1));
            mbar_wait(bar(B_K_EMPTY + s), ph ^ 1);
            if (elect_one()) {
                const uint32_t dst = sbase + OFF_K + s * TILE_BYTES;
                mbar_expect_tx(bar(B_K_FULL + s), TILE_BYTES);
                tma_load_2d(dst, &map_k, 0, row, bar(B_K_FULL + s));
                tma_load_2d(dst + 16384, &map_k, 64, row, bar(B_K_FULL + s));
            }
            __syncwarp();
            mbar_wait(bar(B_V_EMPTY + s), ph ^ 1);
            if (elect_one()) {
                const uint32_t dst = sbase + OFF_V + s * TILE_BYTES;
                mbar_expect_tx(bar(B_V_FULL + s), TILE_BYTES);
                tma_load_2d(dst, &map_v, 0, row, bar(B_V_FULL + s));
                tma_load_2d(dst + 16384, &map_v, 64, row, bar(B_V_FULL + s));
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // PAIR p = (key tile j, row tile r): ONE QK^T of N = 128 keys into the 128-column pair buffer p & 1, whose two
        // 64-key halves the two softmax warpgroups turn into P in place, then ONE PV (K = 128 keys) into the row tile's
        // single accumulator.  In-order schedule:  QK(0) QK(1) | P(0): PV(0) QK(2) | P(1): PV(1) QK(3) | ...
        const uint32_t q_lo = (sbase >> 4) | LBO_K;
        const uint32_t k_lo = ((sbase + OFF_K) >> 4) | LBO_K;
        const uint32_t v_lo = ((sbase + OFF_V) >> 4) | LBO_V;
        const int n_pairs = n_tiles * NR;
        int pq = 0, rq = 0, sq = 0, phq = 0;          // next QK pair: row tile, K stage, K_FULL parity
        int rp = 0, sp = 0, php = 0;                  // next PV pair: row tile, V stage, V_FULL parity
        auto issue_qk = [&]() {
            if (rq == 0) mbar_wait(bar(B_K_FULL + sq), phq);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = q_lo + rq * (32768 >> 4), b0 = k_lo + sq * (TILE_BYTES >> 4);
                const uint32_t dst = tmem + COL_SP + (pq & 1) * 128;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    tc_mma2(dst, a0 + (kk >> 2) * (16384 >> 4) + (kk & 3) * 2, b0 + (kk >> 2) * (16384 >> 4) + (kk & 3) * 2,
                            IDESC_QK, kk > 0 ? 1u : 0u);
                if (rq == NR - 1) tc_commit(bar(B_K_EMPTY + sq));
            }
            __syncwarp();
            ++pq;
            if (++rq == NR) {
                rq = 0;
                if (++sq == STAGES) { sq = 0; phq ^= 1; }
            }
        };
        for (int i = 0; i < 2 && i < n_pairs; ++i) {
            issue_qk();
            if (elect_one()) tc_commit(bar(B_S_FULL + i));
            __syncwarp();
        }
        for (int p = 0; p < n_pairs; ++p) {
            if (rp == 0) mbar_wait(bar(B_V_FULL + sp), php);
            mbar_wait(bar(B_P_FULL + (p & 1)), (p >> 1) & 1);
            tc_fence_after();
            if (lane == 0) LS_TRACE(0, p);
            if (elect_one()) {
                const uint32_t pbuf = tmem + COL_SP + (p & 1) * 128, b0 = v_lo + sp * (TILE_BYTES >> 4);
                const uint32_t dst = tmem + COL_O + rp * 128;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)      // P of keys 0-63 sits in columns 0-31, of keys 64-127 in columns 64-95
                    tc_mma_ts(dst, pbuf + (kk >> 2) * 64 + (kk & 3) * 8, b0 + kk * (2048 >> 4), IDESC_PV,
                              (p >= NR || kk > 0) ? 1u : 0u);
                if (rp == NR - 1) tc_commit(bar(B_V_EMPTY + sp));
            }
            __syncwarp();
            if (++rp == NR) {
                rp = 0;
                if (++sp == STAGES) { sp = 0; php ^= 1; }
            }
            if (pq < n_pairs) issue_qk();
            if (elect_one()) tc_commit(bar(B_S_FULL + (p & 1)));   // S of pair p+2 is ready / PV of pair p is complete
            __syncwarp();
            if (lane == 0) LS_TRACE(1, p);
        }
        if (n_pairs > 0 && elect_one()) tc_commit(bar(B_ALL_DONE));   // every MMA of this CTA has completed once this fires
        __syncwarp();
    } else if (warp >= 4) {
        // ================= softmax warpgroups =================
        // Warpgroup w owns the 64-key half w of EVERY pair; the two threads (w = 0, 1) of a row share the row's
        // reference maximum m_ref - identical in both by construction - and keep private partial sums l.  The lazy
        // rescale needs both halves' maxima: one bar.red.or over the 64 threads of the two warps that share a TMEM lane
        // quarter tells both warps whether ANY of their rows has to move its reference; only then do they exchange
        // maxima through shared memory and rescale the accumulator (each thread 64 of its row's 128 columns).
        const int wg = (warp - 4) >> 2;                              // = 64-key half of every pair
        const int quad = warp & 3;                                   // TMEM lane quarter of this warp
        const int row = quad * 32 + lane;                            // row inside a 128-row tile
        const uint32_t lane_addr = tmem + (static_cast<uint32_t>(quad * 32) << 16);
        const float c2 = d.scale * 1.4426950408889634f;
        const uint64_t c2x2 = pack2(c2, c2);
        float* xch = reinterpret_cast<float*>(smem + OFF_STATE);     // [2 warpgroups][128 rows] exchange scratch
        const int n_pairs = n_tiles * NR;
        float m_ref0 = -INFINITY, m_ref1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // per local row tile
        int j = 0, r = 0;
        for (int p = 0; p < n_pairs; ++p) {
            const int pb = p & 1;
            const int rp0 = (rt0 + r) * 128;
            const bool warp_valid = rp0 + quad * 32 < R;             // the same for both warps of a quarter
            mbar_wait(bar(B_S_FULL + pb), (p >> 1) & 1);
            if (warp_valid) {
                tc_fence_after();
                if (quad == 0 && lane == 0) LS_TRACE(2, 2 * p + wg);
                const uint32_t buf = lane_addr + COL_SP + pb * 128 + wg * 64;
                uint32_t sv[64];
                LS_TMEM_LD32(buf, sv);
                LS_TMEM_LD32(buf + 32, (sv + 32));
                tmem_wait_ld();
                if (quad == 0 && lane == 0) LS_TRACE(3, 2 * p + wg);
                const int gi = tile_index(j);
                if (gi >= total_tiles) {
                    // per-row visibility word of this half's 64 key slots
                    const SpecialTile st = sp_tiles[gi - total_tiles];
                    const int rp = rp0 + row;
                    uint64_t vis = 0ull;
                    if (rp < R) {
                        const int qi = rp / G;
                        if (st.kind == 0) {
                            const int hi = d.causal ? kv_len - (n_q - 1 - qi) : kv_len;      // visible global keys: key < hi
                            const int cnt = min(max(hi - st.base, 0), st.n_valid) - wg * 64;
                            vis = cnt >= 64 ? ~0ull : (cnt <= 0 ? 0ull : ((1ull << cnt) - 1ull));
                        } else {
                            const uint64_t* w = d.tail_mask + static_cast<size_t>(qi) * LS_MASK_WORDS;
                            const int k0 = st.base + wg * 64;
                            const int wi = k0 >> 6, sh = k0 & 63;
                            if (wi < LS_MASK_WORDS) {
                                vis = w[wi] >> sh;
                                if (sh && wi + 1 < LS_MASK_WORDS) vis |= w[wi + 1] << (64 - sh);
                            }
                            const int cnt = st.n_valid - wg * 64;
                            if (cnt < 64) vis &= cnt <= 0 ? 0ull : ((1ull << cnt) - 1ull);
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
                float m_ref = r == 0 ? m_ref0 : m_ref1;
                // does ANY row of this lane quarter (either half) have to move its reference?  (named barrier 2 + quad,
                // the 64 threads of the two warps that share the quarter)
                uint32_t any;
                {
                    const uint32_t need = mx > m_ref + 8.0f ? 1u : 0u;   // lazy rescale threshold (2^8)
                    asm volatile(
                        "{\n\t"
                        ".reg .pred pin, pout;\n\t"
                        "setp.ne.u32 pin, %1, 0;\n\t"
                        "bar.red.or.pred pout, %2, 64, pin;\n\t"
                        "selp.u32 %0, 1, 0, pout;\n\t"
                        "}" : "=r"(any) : "r"(need), "r"(2 + quad) : "memory");
                }
                if (any) {
                    xch[wg * 128 + row] = mx;
                    asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");
                    const float joint = fmaxf(mx, xch[(wg ^ 1) * 128 + row]);
                    const float m_new = joint > m_ref + 8.0f ? joint : m_ref;   // the same decision in both threads of the row
                    const float alpha = m_ref == -INFINITY ? 0.f : fast_exp2(m_ref - m_new);   // nothing accumulated yet: 0 (never exp2(-inf + inf)); 1 if unchanged
                    if (r == 0) l0 *= alpha; else l1 *= alpha;
                    m_ref = m_new;
                    if (p >= NR) {
                        // The accumulator is quiescent once PV(p - NR) completed.  S_FULL(p) certifies every MMA issued
                        // before QK(p), i.e. PV(p-2) and older; with one row tile PV(p-1) is certified by S_FULL(p+1),
                        // this thread's next wait anyway (its previous phase completed long ago: no parity aliasing).
                        if (NR == 1) mbar_wait(bar(B_S_FULL + ((p + 1) & 1)), ((p + 1) >> 1) & 1);
                        tc_fence_after();
                        const uint32_t o_half = lane_addr + COL_O + r * 128 + wg * 64;
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            uint32_t ov[32];
                            LS_TMEM_LD32(o_half + c * 32, ov);
                            tmem_wait_ld();
#pragma unroll
                            for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
                            LS_TMEM_ST32(o_half + c * 32, ov);
                        }
                        tmem_wait_st();
                        tc_fence_before();
                    }
                    if (r == 0) m_ref0 = m_ref; else m_ref1 = m_ref;
                    asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");   // xch is reused by the next exchange
                }
                // p = 2^(s * c2 - m_ref), packed fp32x2 FMA / ADD
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
                    float a0, a1;
                    unpack2(add2(sum_a, sum_b), a0, a1);
                    if (r == 0) l0 += a0 + a1; else l1 += a0 + a1;
                }
                // P -> in place over the first 32 columns of this half's S columns
                if (quad == 0 && lane == 0) LS_TRACE(4, 2 * p + wg);
                LS_TMEM_ST32(buf, pk);
                tmem_wait_st();
                tc_fence_before();
            }
            mbar_arrive(bar(B_P_FULL + pb));
            if (quad == 0 && lane == 0) LS_TRACE(5, 2 * p + wg);
            if (++r == NR) { r = 0; ++j; }
        }
        // ================= epilogue: normalised partial O + LSE =================
        const int Rtot = d.q_rows * G;
        const int nr_out = min(NRS - rt0, NRL_MAX);                 // row tiles this CTA is responsible for
        if (n_pairs > 0) {
            mbar_wait(bar(B_ALL_DONE), 0);           // single-phase barrier: no parity aliasing
            tc_fence_after();
        }
        auto part_base = [&](int sl, int rp) { return (static_cast<size_t>(d.part_offset + sl) * d.n_kv_heads + kvh) * Rtot + rp; };
        // total row sums: the two warpgroups hold partial sums over their key halves relative to the SAME reference
        float* xl = xch;                                             // [2 row tiles][2 warpgroups][128]: reuse after a full sync
        asm volatile("bar.sync 1, 256;" ::: "memory");
        xl[(0 * 2 + wg) * 128 + row] = l0;
        xl[(1 * 2 + wg) * 128 + row] = l1;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int re = 0; re < nr_out; ++re) {
            const bool live = n_pairs > 0 && re < NR;
            const int rp = (rt0 + re) * 128 + row;
            const float lt = xl[(re * 2) * 128 + row] + xl[(re * 2 + 1) * 128 + row];
            const float m_ref = re == 0 ? m_ref0 : m_ref1;
            const float inv = (live && lt > 0.f) ? 1.f / lt : 0.f;
            const size_t base = part_base(slot, rp);
            if (wg == 0 && rp < R) {
                d.part_lse[base] = inv > 0.f ? (m_ref + log2f(lt)) * 0.6931471805599453f : -INFINITY;
                if (empty_slot >= 0) d.part_lse[part_base(empty_slot, rp)] = -INFINITY;
            }
            if ((rt0 + re) * 128 + quad * 32 < R) {   // warp-uniform (tcgen05.ld is warp-collective)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int col = wg * 64 + c * 32;
                    uint32_t ov[32];
                    if (live) {
                        LS_TMEM_LD32(lane_addr + COL_O + re * 128 + col, ov);
                        tmem_wait_ld();
                    }
                    if (inv == 0.f) {                 // no key seen (or no tile at all): the accumulator holds stale bits
#pragma unroll
                        for (int e = 0; e < 32; ++e) ov[e] = 0u;
                    }
                    if (rp < R) {
                        float* po = d.part_o + base * 128 + col;
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            *reinterpret_cast<float4*>(po + e) =
                                make_float4(__uint_as_float(ov[e]) * inv, __uint_as_float(ov[e + 1]) * inv,
                                            __uint_as_float(ov[e + 2]) * inv, __uint_as_float(ov[e + 3]) * inv);
                        if (empty_slot >= 0) {
                            float* pe = d.part_o + part_base(empty_slot, rp) * 128 + col;
#pragma unroll
                            for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(pe + e) = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
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
    cuuint32_t box[2] = {64, 128};
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

// Launch the tensor-core kernel over the unmasked full tiles, writing partial slots
// [part_offset, part_offset + n_splits_tc).  With `fuse_remainder` the masked remainder (partial tile, causal
// rows, tree-masked tail) is appended to the tile sequence as special tiles and shared out with the rest.
// Up to two row tiles: one CTA per (slot, kv head).  Three row tiles: two heavy CTAs per pair of slots plus
// one light CTA (see the header comment): grid.x = n_slots + ceil(n_slots / 2).
static int launch_tc(const ls_attn_desc* d, int n_splits_tc, int fuse_remainder, void* stream) {
    using namespace ls;
    using namespace ls::tc;
    LS_REQUIRE(d && n_splits_tc >= 1, "ls_attn_partial_tc: bad arguments");
    const int G = d->n_heads / d->n_kv_heads;
    LS_REQUIRE(d->q_rows * G <= NR_MAX * 128, "ls_attn_partial_tc: %d packed rows exceed %d", d->q_rows * G, NR_MAX * 128);
    LS_REQUIRE(d->window < 0, "ls_attn_partial_tc: windowed attention stays on the general path");
    LS_REQUIRE(d->n_pages > 0, "ls_attn_partial_tc: n_pages must be set");
    LS_REQUIRE(d->q_stride == 0 || (d->q_stride >= static_cast<int64_t>(d->n_heads) * 128 && d->q_stride % 8 == 0),
               "ls_attn_partial_tc: q_stride %lld must be 0 or >= n_heads * 128 and a multiple of 8 elements",
               static_cast<long long>(d->q_stride));
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
    const int nrs = (d->q_rows * G + 127) / 128;
    dim3 grid(nrs > NRL_MAX ? n_splits_tc + (n_splits_tc + 1) / 2 : n_splits_tc, d->n_kv_heads);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d->dtype == LS_F16)
        attn_tc_kernel<__half><<<grid, THREADS, SMEM_BYTES, st>>>(*d, mk, mv, n_splits_tc, fuse_remainder);
    else
        attn_tc_kernel<__nv_bfloat16><<<grid, THREADS, SMEM_BYTES, st>>>(*d, mk, mv, n_splits_tc, fuse_remainder);
    return check_launch("ls_attn_partial_tc");
}

#ifdef LS_TC_TRACE
extern "C" int ls_debug_set_flags(int flags) {
    return cudaMemcpyToSymbol(ls::tc::g_dbg, &flags, sizeof(flags)) == cudaSuccess ? LS_OK : LS_ERR_CUDA;
}
extern "C" int ls_debug_set_trace(void* buf) {
    long long* p = static_cast<long long*>(buf);
    return cudaMemcpyToSymbol(ls::tc::g_trace, &p, sizeof(p)) == cudaSuccess ? LS_OK : LS_ERR_CUDA;
}
#endif

extern "C" int ls_attn_partial_tc(const ls_attn_desc* d, int n_splits_tc, void* stream) { return launch_tc(d, n_splits_tc, 0, stream); }
// the whole call in one launch: d->n_splits CTAs per kv head share the unmasked and the special tiles
extern "C" int ls_attn_partial_tc_fused(const ls_attn_desc* d, void* stream) {
    LS_REQUIRE(d && d->n_splits >= 1, "ls_attn_partial_tc_fused: needs n_splits >= 1");
    return launch_tc(d, d->n_splits, 1, stream);
}

// tcgen05 / TMA / TMEM split-KV attention over the paged prefix + the tree-masked tail (ls_attn_partial).
//
// Two kernels share the helpers of this file: attn_tc_kernel (described here; calls with up to two row tiles, and the
// heavy / heavy / light trios that served three-row-tile calls in round 2a) and attn_tc1_kernel (further down: one row
// tile per CTA, the default for three-row-tile calls - ls_attn_tc1_mode / LONGSPEC_TC1).
//
// Work unit: one CTA = (kv head, a share of the 128-key tiles, up to TWO row tiles of 128 packed query rows),
// a packed row being r = token * G + head-in-group, so a K/V tile serves the whole GQA group and candidate tree.
//
// TMEM (512 columns) = two 128-column S/P buffers + two 128-column O accumulators - the FA4 shape.  A PAIR is
// (key tile, row tile): ONE S = Q K^T with N = 128 keys (tensor-rate; the N = 64 form of round 1 was bound by
// shared-memory operand bandwidth), one softmax pass, ONE O += P V with K = 128 keys.  Pair p lives in S/P buffer
// p & 1 and accumulates into O accumulator p & 1.  FOUR softmax warpgroups: buffer b belongs to warpgroups 2b and
// 2b + 1, which take the two 64-key halves of every pair in that buffer (a single warpgroup per pair is latency-bound:
// 1390 cycles per 128 x 128 tile against 1030 of MMA - profiles/r02_*; four warps per scheduler hide it).  The two
// threads of a row share its reference maximum - identical in both by construction - and keep private partial sums.
// Two row tiles: pair p = (key tile p >> 1, row tile p & 1), i.e. one accumulator per row tile.  One row tile: pair
// p = key tile p, alternate key tiles go to the two accumulators and the epilogue merges the two partial states.
// A call with three row tiles (the Llama-3 verify shape: 74 tokens x GQA-4 = 296 rows) is cut into "trios": two
// HEAVY CTAs (row tiles 0-1, alternate key tiles of the group's range) and one LIGHT CTA (row tile 2, every key
// tile of the range).  All three run the same number of pairs, and the light CTA's K/V loads hit L2 behind the
// heavy CTAs' loads, so HBM still sees every byte once (ncu: dram bytes = algorithmic bytes).
//
//   warp 0              TMA producer: 128-key K and V tiles (2 x [128 x 128 B] boxes each, SWIZZLE_128B) from the
//                       paged pool into separate 2-stage K and V rings, completion on mbarriers
//   warp 1              MMA issuer (converged warp, one elected lane): S = Q K^T (M128 N128 K16 x8, SS) and
//                       O += P V (M128 N128 K16 x8, TS: P read from TMEM, V MN-major in smem):
//                         QK(0) QK(1) | P(0) ready: PV(0) QK(2) | P(1) ready: PV(1) QK(3) | ...
//                       MMAs of one thread execute in issue order, so QK(p+2) cannot overwrite the buffer PV(p)
//                       still reads: no "buffer free" barriers, one wait and one commit per 16 MMAs.
//   warp 2              TMEM allocation
//   warps 4-19          four softmax warpgroups (setmaxnreg 112; warps 0-3 drop to 32).  Half a row per thread:
//                       tcgen05.ld 64 keys of the S tile, online softmax in the log2 domain with lazy rescaling (O in
//                       TMEM is only rescaled when a row maximum grows by more than 2^8; one bar.red.or over the 64
//                       threads that share a TMEM lane quarter and a buffer decides it), P rounded to dtype and
//                       tcgen05.st IN PLACE over the first 32 columns of the S half it came from.  The kernel sits
//                       where MUFU (16 ex2/clk/SM) and the tensor pipe cost the same per tile, so LS_TC_POLY of
//                       every 8 exponentials are computed on the FMA pipe instead (Cody-Waite + cubic, rel. error
//                       2^-13.2 - below the rounding of P to bf16/fp16; tools/ubench/softmax_math.cu).
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
#include <stdlib.h>

#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "common.cuh"

namespace ls {
namespace tc {

constexpr int BN = 128;                            // keys per tile
constexpr int NRL_MAX = 2;                         // row tiles per CTA
constexpr int NR_MAX = 3;                          // row tiles per call
constexpr int STAGES = 2;                          // K ring and V ring depth
constexpr int THREADS = 640;                       // 4 control warps + 16 softmax warps
constexpr int TILE_BYTES = 32768;                  // one K or V tile: [2 x 64-element chunks][128 keys][128 B]
constexpr int Q_BYTES = NRL_MAX * 32768;           // [row tile][2 k-chunks][128 rows][128 B]
constexpr int STATE_BYTES = 4096;                  // row-maximum exchange between the two halves of a row / epilogue exchange
constexpr int BAR_BYTES = 256;
constexpr int SPECIAL_BYTES = 256;                 // special-tile table (12 x 16 B)
constexpr int OFF_K = Q_BYTES;
constexpr int OFF_V = OFF_K + STAGES * TILE_BYTES;
constexpr int OFF_STATE = OFF_V + STAGES * TILE_BYTES;
constexpr int OFF_BAR = OFF_STATE + STATE_BYTES;
constexpr int OFF_SPECIAL = OFF_BAR + BAR_BYTES;
constexpr int SMEM_BYTES = OFF_SPECIAL + SPECIAL_BYTES + 1024;   // + worst-case alignment of the dynamic base

constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_SP = 0;                     // two buffers of 128 columns: a 128 x 128 fp32 score tile, overwritten
                                                   // in place by its 128 x 128 bf16/fp16 probability tile (64 columns)
constexpr uint32_t COL_O = 256;                    // 2 x 128

constexpr int BAR_Q_READY = 15;                    // named barrier: Q staged (softmax warps arrive, the MMA issuer waits)
constexpr int OUT_STRIDE = 132;                    // floats per staged output row (128 + 4: conflict-free 16-byte stores)
constexpr int MAX_SPECIAL = 12;
struct SpecialTile {
    int tok;      // first LOCAL token of the tile (need not be 128-aligned for tail tiles)
    int n_valid;  // keys of the tile that exist (the rest of the 128-row TMA box is masked)
    int kind;     // 0 = prefix keys (causal limit applies), 1 = tail keys (tree bit mask applies)
    int base;     // kind 0: global index of key 0;  kind 1: tail index of key 0
};

// S_FULL / P_FULL are indexed by pair parity (= softmax warpgroup).  Every barrier is waited on phase by phase by its
// waiter, so mbarrier parity waits can never alias.
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
// Cluster flavours (attn_tc1_kernel<T, true>): the same K / V tile feeds both CTAs of a pair, so each CTA loads one
// 64-element half and multicasts it into both shared memories; "stage free" commits arrive on both CTAs' barriers.
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
                 ::"r"(dst), "l"(map), "r"(bar), "h"(mask), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
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
__device__ int g_dbg = 0;     // timing experiments (results become wrong): 1 no PV MMAs, 2 no QK MMAs, 4 no softmax math, 16 no TMA;
                              // bits 16.. = blockIdx.x of the traced CTA
#define LS_DBG(bit) (dbg & (bit))
__device__ long long* g_trace = nullptr;
#define LS_TRACE_INIT const int dbg = g_dbg & 0xffff; long long* const trace_buf = (static_cast<int>(blockIdx.x) == (g_dbg >> 16) && blockIdx.y == 0) ? g_trace : nullptr; \
    if (trace_buf && threadIdx.x == 0) trace_buf[7 * 512 + 511] = clock64();   /* kernel entry of the traced CTA */
#define LS_TRACE(ev, t)                                                       \
    do {                                                                      \
        if (trace_buf && (t) < 512) trace_buf[(ev) * 512 + (t)] = clock64();  \
    } while (0)
#else
#define LS_TRACE_INIT
#define LS_TRACE(ev, t) do { } while (0)
#define LS_DBG(bit) 0
#endif

// 16-register flavour of the TMEM store (one 32-key slice of P = 16 packed columns)
#define LS_TMEM_ST16(taddr, v)                                                                                       \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"   \
                 "%16};"                                                                                              \
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),           \
                   "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),        \
                   "r"(v[15])                                                                                         \
                 : "memory")

#define LS_TMEM_LD16(taddr, v)                                                                                       \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),     \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
                 : "r"(taddr))

// Register budget (setmaxnreg works per warpgroup): warps 0-3 (TMA, MMA issue, TMEM allocation) give registers back,
// the four softmax warpgroups hold 64 scores per thread.  setmaxnreg only moves registers inside the pool the CTA was
// LAUNCHED with (640 threads x 96 registers = 61440, not the SM's 65536): a request beyond it blocks forever, so
// 128 * REGS_CTRL + 512 * REGS_SOFTMAX must not exceed THREADS * REGS_LAUNCH.
constexpr int REGS_LAUNCH = 96;                    // __launch_bounds__(640, 1): floor(65536 / 640) rounded down to 8
constexpr int REGS_CTRL = 32;
constexpr int REGS_SOFTMAX = 112;
static_assert(128 * REGS_CTRL + 512 * REGS_SOFTMAX <= THREADS * REGS_LAUNCH, "setmaxnreg budget exceeds the CTA's register pool");
static_assert(REGS_CTRL % 8 == 0 && REGS_SOFTMAX % 8 == 0, "setmaxnreg takes multiples of 8");
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// How many of every 8 exponentials leave the MUFU pipe for the FMA pipe (tools/ubench/softmax_math.cu: 2 of 8 is the
// measured optimum: 868 instead of 1049 cycles per 128 x 128 tile).  Must be even (the polynomial works on pairs).
#ifndef LS_TC_POLY
#define LS_TC_POLY 2
#endif
static_assert(LS_TC_POLY >= 0 && LS_TC_POLY <= 8 && LS_TC_POLY % 2 == 0, "LS_TC_POLY must be 0, 2, 4, 6 or 8");

// 2^x for a packed pair on the FMA pipe.  x = n + f with n = round(x) (the 1.5 * 2^23 trick leaves n in the low
// mantissa bits of t), f in [-1/2, 1/2]; 2^f by a cubic (Chebyshev fit, max relative error 2^-13.2); the result is
// assembled by adding n to the exponent field.  Inputs are clamped at -126 (anything below is 0 after rounding to P's
// 16-bit type anyway: 2^-126 * 2^8 is far below the smallest bf16/fp16 a row sum can see).
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& p0, float& p1) {
    constexpr float MAGIC = 12582912.f;
    float x0, x1;
    unpack2(x2, x0, x1);
    x0 = fmaxf(x0, -126.f);
    x1 = fmaxf(x1, -126.f);
    x2 = pack2(x0, x1);
    const uint64_t t2 = add2(x2, pack2(MAGIC, MAGIC));
    const uint64_t n2 = add2(t2, pack2(-MAGIC, -MAGIC));
    const uint64_t f2 = fma2(n2, pack2(-1.f, -1.f), x2);
    uint64_t q2 = fma2(f2, pack2(0.05592204f, 0.05592204f), pack2(0.24264008f, 0.24264008f));
    q2 = fma2(q2, f2, pack2(0.69312102f, 0.69312102f));
    q2 = fma2(q2, f2, pack2(0.99992448f, 0.99992448f));
    float q0, q1, t0, t1;
    unpack2(q2, q0, q1);
    unpack2(t2, t0, t1);
    p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
    p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

// The exponential phase of half a score row (64 keys): p = 2^(s * c2 - m), packed fp32x2 FMA / ADD, POLY of every 8
// exponentials on the FMA pipe and the rest on MUFU; P is rounded to the 16-bit type and stored to TMEM (A operand of
// the PV MMA, 2 values per 32-bit column) one 32-key slice at a time.  Returns the row sum of the unrounded p.
template <typename T, int POLY>
__device__ __forceinline__ float exp_phase(const uint32_t (&sv)[64], uint32_t buf, uint64_t c2x2, uint64_t negm) {
    uint64_t sum_a = 0ull, sum_b = 0ull;                     // {0.f, 0.f}
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
            const uint64_t x2 = fma2(pack2(__uint_as_float(sv[c * 32 + e]), __uint_as_float(sv[c * 32 + e + 1])), c2x2, negm);
            float p0, p1;
            if ((e & 7) < POLY) {
                exp2_poly2(x2, p0, p1);
            } else {
                float x0, x1;
                unpack2(x2, x0, x1);
                p0 = fast_exp2(x0);
                p1 = fast_exp2(x1);
            }
            if (e & 2) sum_b = add2(sum_b, pack2(p0, p1)); else sum_a = add2(sum_a, pack2(p0, p1));
            pk[e >> 1] = Num<T>::pack(p0, p1);
        }
        LS_TMEM_ST16(buf + c * 16, pk);
    }
    float a0, a1;
    unpack2(add2(sum_a, sum_b), a0, a1);
    return a0 + a1;
}

// The same phase with P left in registers (attn_tc1_kernel stores it once the P columns are free).
template <typename T, int POLY>
__device__ __forceinline__ float exp_regs(const uint32_t (&sv)[64], uint32_t (&pk)[32], uint64_t c2x2, uint64_t negm) {
    uint64_t sum_a = 0ull, sum_b = 0ull;
#pragma unroll
    for (int e = 0; e < 64; e += 2) {
        const uint64_t x2 = fma2(pack2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), c2x2, negm);
        float p0, p1;
        if ((e & 7) < POLY) {
            exp2_poly2(x2, p0, p1);
        } else {
            float x0, x1;
            unpack2(x2, x0, x1);
            p0 = fast_exp2(x0);
            p1 = fast_exp2(x1);
        }
        if (e & 2) sum_b = add2(sum_b, pack2(p0, p1)); else sum_a = add2(sum_a, pack2(p0, p1));
        pk[e >> 1] = Num<T>::pack(p0, p1);
    }
    float a0, a1;
    unpack2(add2(sum_a, sum_b), a0, a1);
    return a0 + a1;
}

// The masked remainder of a call as SPECIAL tiles appended to the tile sequence: the prefix keys behind the last full
// unmasked tile (partial tile, causal rows) and the tree-masked tail of this round's tokens.  Thread 0 writes the table
// (visible after the next CTA barrier); every thread gets the count.
__device__ __forceinline__ int plan_special_tiles(const ls_attn_desc& d, int kv_len, int total_tiles, SpecialTile* sp_tiles, int tid) {
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
    return n;
}

// Visibility word of packed row rp for the 64 key slots of half h of a special tile
__device__ __forceinline__ uint64_t special_visibility(const ls_attn_desc& d, const SpecialTile& st, int rp, int R, int G,
                                                       int kv_len, int n_q, int h) {
    uint64_t vis = 0ull;
    if (rp < R) {
        const int qi = rp / G;
        if (st.kind == 0) {
            const int hi = d.causal ? kv_len - (n_q - 1 - qi) : kv_len;      // visible global keys: key < hi
            const int cnt = min(max(hi - st.base, 0), st.n_valid) - h * 64;
            vis = cnt >= 64 ? ~0ull : (cnt <= 0 ? 0ull : ((1ull << cnt) - 1ull));
        } else {
            const uint64_t* w = d.tail_mask + static_cast<size_t>(qi) * LS_MASK_WORDS;
            const int k0 = st.base + h * 64;
            const int wi = k0 >> 6, sh = k0 & 63;
            if (wi < LS_MASK_WORDS) {
                vis = w[wi] >> sh;
                if (sh && wi + 1 < LS_MASK_WORDS) vis |= w[wi + 1] << (64 - sh);
            }
            const int cnt = st.n_valid - h * 64;
            if (cnt < 64) vis &= cnt <= 0 ? 0ull : ((1ull << cnt) - 1ull);
        }
    }
    return vis;
}

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
    // With `fuse` the masked remainder rides along as SPECIAL tiles appended to the tile sequence; they come from a small
    // table in shared memory and the softmax applies a per-row bit mask to them.
    SpecialTile* sp_tiles = reinterpret_cast<SpecialTile*>(smem + OFF_SPECIAL);
    const int n_special = (fuse & 1) ? plan_special_tiles(d, kv_len, total_tiles, sp_tiles, tid) : 0;
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
    // Pair p = (key tile, row tile): NR == 2: (p >> 1, p & 1); NR == 1: (p, 0).  It runs in S/P buffer p & 1,
    // accumulates into O accumulator p & 1 and is processed by softmax warpgroup p & 1.
    const int n_pairs = n_tiles * NR;
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
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // Q is staged by the sixteen softmax warps while the TMA producer already streams the first K/V tiles; only the
    // MMA issuer waits for it (named barrier B_Q_READY: 512 arrivals + the issuer warp).

    constexpr int FMT = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
    constexpr uint32_t IDESC_QK = make_idesc(128, 128, FMT, 0, 0);
    constexpr uint32_t IDESC_PV = make_idesc(128, 128, FMT, 0, 1);

    if (warp < 4) {
        reg_dec<REGS_CTRL>();
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
            // The whole warp runs this loop converged (all values warp-uniform); one elected lane issues.
            //   QK(0) QK(1) | P(0) ready: PV(0) QK(2) | P(1) ready: PV(1) QK(3) | ...
            // MMAs of one thread execute in issue order, so QK(p+2) cannot overwrite the buffer PV(p) still reads P
            // from.  One commit per iteration: S_FULL of pair p+2, which also certifies PV(p) - hence every earlier PV
            // into the same accumulator - complete: the softmax warpgroup that sees S_FULL(p+2) may rescale O freely.
            const uint32_t q_lo = (sbase >> 4) | LBO_K;
            const uint32_t k_lo = ((sbase + OFF_K) >> 4) | LBO_K;
            const uint32_t v_lo = ((sbase + OFF_V) >> 4) | LBO_V;
            asm volatile("bar.sync %0, %1;" ::"n"(BAR_Q_READY), "n"(512 + 32) : "memory");     // Q tiles are in shared memory
            int pq = 0, rq = 0, sq = 0, phq = 0;          // next QK pair: index, row tile, K stage, K_FULL parity
            int rp = 0, sp = 0, php = 0;                  // next PV pair: row tile, V stage, V_FULL parity
            auto issue_qk = [&]() {
                if (rq == 0 && !LS_DBG(16)) mbar_wait(bar(B_K_FULL + sq), phq);
                tc_fence_after();
                if (lane == 0 && pq >= 2) LS_TRACE(7, pq - 2);
                if (elect_one()) {
                    const uint32_t a0 = q_lo + rq * (32768 >> 4), b0 = k_lo + sq * (TILE_BYTES >> 4);
                    const uint32_t dst = tmem + COL_SP + (pq & 1) * 128;
                    if (!LS_DBG(2))
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
                if (rp == 0 && !LS_DBG(16)) mbar_wait(bar(B_V_FULL + sp), php);
                if (!LS_DBG(8)) mbar_wait(bar(B_P_FULL + (p & 1)), (p >> 1) & 1);
                tc_fence_after();
                if (lane == 0) LS_TRACE(0, p);
                if (elect_one()) {
                    const uint32_t a0 = tmem + COL_SP + (p & 1) * 128, b0 = v_lo + sp * (TILE_BYTES >> 4);
                    const uint32_t dst = tmem + COL_O + (p & 1) * 128;
                    if (!LS_DBG(1))
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)      // 16 keys = 8 packed TMEM columns per K step; P of keys 0-63 sits in
                                                        // columns 0-31 of the buffer, P of keys 64-127 in columns 64-95
                        tc_mma_ts(dst, a0 + (kk >> 2) * 64 + (kk & 3) * 8, b0 + kk * (2048 >> 4), IDESC_PV, (p >= 2 || kk > 0) ? 1u : 0u);
                    if (rp == NR - 1) tc_commit(bar(B_V_EMPTY + sp));
                }
                __syncwarp();
                if (lane == 0) LS_TRACE(6, p);
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
        }
    } else {
        reg_inc<REGS_SOFTMAX>();
        // Q tiles -> swizzled shared memory (K-major, SWIZZLE_128B): [tile][k-chunk][row][128B].  All loads of a thread are
        // issued before its first store (4 x 16 B per row tile and thread), so the staging costs one memory round trip.
        {
            const T* __restrict__ q = reinterpret_cast<const T*>(d.q);
            const size_t q_stride = d.q_stride > 0 ? static_cast<size_t>(d.q_stride) : static_cast<size_t>(d.n_heads) * 128;
            const int stid = tid - 128;
            uint4 qv[NRL_MAX * 4];
#pragma unroll
            for (int k = 0; k < NRL_MAX * 4; ++k) {
                const int idx = stid + k * 512;
                const int rl = idx >> 4, chunk = idx & 15;
                const int rp = rt0 * 128 + rl;
                qv[k] = make_uint4(0, 0, 0, 0);
                if (k < NR * 4 && n_pairs > 0 && rp < R) {
                    const int i = rp / G, g = rp - i * G;
                    qv[k] = *reinterpret_cast<const uint4*>(q + static_cast<size_t>(i) * q_stride + static_cast<size_t>(kvh * G + g) * 128 + chunk * 8);
                }
            }
#pragma unroll
            for (int k = 0; k < NRL_MAX * 4; ++k) {
                const int idx = stid + k * 512;
                const int rl = idx >> 4, chunk = idx & 15;
                const int tile = rl >> 7, row = rl & 127;
                if (k < NR * 4 && n_pairs > 0)
                    *reinterpret_cast<uint4*>(smem + tile * 32768 + (chunk >> 3) * 16384 + row * 128 + (((chunk & 7) ^ (row & 7)) << 4)) = qv[k];
            }
            fence_async_smem();                                      // generic-proxy writes -> visible to the MMA's async-proxy reads
            asm volatile("bar.arrive %0, %1;" ::"n"(BAR_Q_READY), "n"(512 + 32) : "memory");
        }
        // ================= softmax warpgroups =================
        // Warpgroup wgi = 2 * b + h: 64-key half h of every pair in buffer b (= accumulator b).  The two threads (h = 0, 1)
        // of a row share the row's reference maximum m_ref - identical in both by construction - and keep private
        // partial sums l.  The lazy rescale needs both halves' maxima: one bar.red.or over the 64 threads of the two
        // warps that share (buffer, TMEM lane quarter) tells both whether ANY of their rows has to move its reference;
        // only then do they exchange maxima through shared memory and rescale the accumulator (each thread 64 of its
        // row's 128 columns).
        const int wgi = (warp - 4) >> 2;
        const int b = wgi >> 1, h = wgi & 1;
        const int quad = warp & 3;                                   // TMEM lane quarter of this warp
        const int row = quad * 32 + lane;                            // row inside a 128-row tile
        const uint32_t lane_addr = tmem + (static_cast<uint32_t>(quad * 32) << 16);
        const float c2 = d.scale * 1.4426950408889634f;
        const uint64_t c2x2 = pack2(c2, c2);
        const int rl = NR == 2 ? b : 0;                              // local row tile of this buffer
        const int rp0 = (rt0 + rl) * 128;                            // its first packed row
        const bool warp_valid = rp0 + quad * 32 < R;                 // warp-uniform, the same in both warps of a pair
        const uint32_t o_half = lane_addr + COL_O + b * 128 + h * 64;
        const uint32_t buf = lane_addr + COL_SP + b * 128 + h * 64;  // this thread's 64 score columns; P goes to the first 32
        const int pair_bar = 2 + b * 4 + quad;                       // named barrier of the two warps sharing (b, quad)
        float* xch = reinterpret_cast<float*>(smem + OFF_STATE);     // [2 buffers][2 halves][128 rows]
        float m_ref = -INFINITY, l = 0.f;                            // this thread's row (shared m_ref), its own key half
        int it = 0;
        for (int p = b; p < n_pairs && !LS_DBG(8); p += 2, ++it) {
            const int j = NR == 2 ? p >> 1 : p;                      // key tile
            mbar_wait(bar(B_S_FULL + b), it & 1);
            if (!warp_valid || LS_DBG(4)) {
                // padding rows only: keep the barrier protocol, skip the math (their O rows are never read)
                mbar_arrive(bar(B_P_FULL + b));
                continue;
            }
            tc_fence_after();
            if (quad == 0 && lane == 0 && h == 0) LS_TRACE(2, p);
            uint32_t sv[64];
            LS_TMEM_LD32(buf, sv);
            LS_TMEM_LD32(buf + 32, (sv + 32));
            tmem_wait_ld();
            if (quad == 0 && lane == 0 && h == 0) LS_TRACE(3, p);
            const int gi = tile_index(j);
            if (gi >= total_tiles) {
                // per-row visibility word of this half's 64 key slots
                const SpecialTile st = sp_tiles[gi - total_tiles];
                const uint64_t vis = special_visibility(d, st, rp0 + row, R, G, kv_len, n_q, h);
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
            // does ANY row of this (buffer, lane quarter) - either half - have to move its reference?
            uint32_t any;
            {
                const uint32_t need = mx > m_ref + 8.0f ? 1u : 0u;   // lazy rescale threshold (2^8)
                asm volatile(
                    "{\n\t"
                    ".reg .pred pin, pout;\n\t"
                    "setp.ne.u32 pin, %1, 0;\n\t"
                    "bar.red.or.pred pout, %2, 64, pin;\n\t"
                    "selp.u32 %0, 1, 0, pout;\n\t"
                    "}" : "=r"(any) : "r"(need), "r"(pair_bar) : "memory");
            }
            if (any) {
                xch[(b * 2 + h) * 128 + row] = mx;
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                const float joint = fmaxf(mx, xch[(b * 2 + (h ^ 1)) * 128 + row]);
                const float m_new = joint > m_ref + 8.0f ? joint : m_ref;   // the same decision in both threads of the row
                const float alpha = m_ref == -INFINITY ? 0.f : fast_exp2(m_ref - m_new);   // nothing accumulated yet: 0 (never exp2(-inf + inf)); 1 if unchanged
                l *= alpha;
                m_ref = m_new;
                if (it >= 1) {
                    // Every earlier PV into this accumulator (pairs p-2, p-4, ...) completed before S_FULL(p) fired
                    // (the commit that published S(p) was issued after PV(p-2)), so the accumulator is quiescent.
                    // (16 columns at a time: the 64 scores of this pair stay live in registers meanwhile)
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t ov[16];
                        LS_TMEM_LD16(o_half + c * 16, ov);
                        tmem_wait_ld();
#pragma unroll
                        for (int e = 0; e < 16; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
                        LS_TMEM_ST16(o_half + c * 16, ov);
                    }
                    tmem_wait_st();
                }
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // xch is reused by the next exchange
            }
            // p = 2^(s * c2 - m_ref); P goes back IN PLACE over the first 32 columns of the S half it came from.
            // Special tiles keep MUFU for every element: a masked score (-inf) must give exactly 0.
            const float m_use = m_ref == -INFINITY ? 0.f : m_ref;      // row with no visible key so far: p = 2^(-inf) = 0
            if (gi >= total_tiles) l += exp_phase<T, 0>(sv, buf, c2x2, pack2(-m_use, -m_use));
            else l += exp_phase<T, LS_TC_POLY>(sv, buf, c2x2, pack2(-m_use, -m_use));
            if (quad == 0 && lane == 0 && h == 0) LS_TRACE(4, p);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(bar(B_P_FULL + b));
            if (quad == 0 && lane == 0 && h == 0) LS_TRACE(5, p);
        }
        // ================= epilogue: normalised partial O + LSE =================
        const int Rtot = d.q_rows * G;
        const int nr_out = min(NRS - rt0, NRL_MAX);                 // row tiles this CTA is responsible for
        if (n_pairs > 0) {
            mbar_wait(bar(B_ALL_DONE), 0);           // single-phase barrier: no parity aliasing
            tc_fence_after();
        }
        auto part_base = [&](int sl, int rp) { return (static_cast<size_t>(d.part_offset + sl) * d.n_kv_heads + kvh) * Rtot + rp; };
        // exchange the row states: m_ref per accumulator (the same in both halves), l per (accumulator, half)
        float* st_m = xch;                                           // [2][128]   (xch is dead: every pair barrier has been passed)
        float* st_l = xch + 2 * 128;                                 // [2][2][128]
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (h == 0) st_m[b * 128 + row] = m_ref;
        st_l[(b * 2 + h) * 128 + row] = l;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        // The normalised rows go through shared memory (Q / K / V space is free: every MMA has completed) so that the
        // global stores are whole 512-byte rows per warp instruction instead of 32 scattered 16-byte pieces - the
        // direct version spent 14 000 cycles per heavy CTA in the LSU (profiles/r02_*).
        float* stage = reinterpret_cast<float*>(smem);               // [row tile][128 rows][OUT_STRIDE]
        int n_stage = 0;                                             // staged row tiles
        if (NR == 2) {
            // accumulator b = row tile b: complete as it stands; the two halves of the warpgroup pair stage 64 columns each
            const int rp = rp0 + row;
            const float lt = st_l[(b * 2) * 128 + row] + st_l[(b * 2 + 1) * 128 + row];
            const float inv = lt > 0.f ? 1.f / lt : 0.f;
            if (h == 0 && rp < R) d.part_lse[part_base(slot, rp)] = inv > 0.f ? (m_ref + log2f(lt)) * 0.6931471805599453f : -INFINITY;
            if (warp_valid) {                             // warp-uniform (tcgen05.ld is warp-collective)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t ov[32];
                    if (n_pairs > 0) {
                        LS_TMEM_LD32(o_half + c * 32, ov);
                        tmem_wait_ld();
                    }
                    if (inv == 0.f) {                     // no key seen (or no tile at all): the accumulator holds stale bits
#pragma unroll
                        for (int e = 0; e < 32; ++e) ov[e] = 0u;
                    }
                    float* ps = stage + (b * 128 + row) * OUT_STRIDE + h * 64 + c * 32;
#pragma unroll
                    for (int e = 0; e < 32; e += 4)
                        *reinterpret_cast<float4*>(ps + e) =
                            make_float4(__uint_as_float(ov[e]) * inv, __uint_as_float(ov[e + 1]) * inv,
                                        __uint_as_float(ov[e + 2]) * inv, __uint_as_float(ov[e + 3]) * inv);
                }
            }
            n_stage = 2;
        } else {
            // One (or no) live row tile: the two accumulators hold partial states of the SAME rows over alternate key
            // tiles.  Merge them; each of the four warpgroups stages 32 of the 128 output columns.  Row tiles without
            // live rows / CTAs without keys write empty partials.
            const float m0 = st_m[row], m1 = st_m[128 + row];
            const float l0 = st_l[row] + st_l[128 + row], l1 = st_l[256 + row] + st_l[384 + row];
            const float m = fmaxf(m0, m1);
            const float a0 = l0 > 0.f ? fast_exp2(m0 - m) : 0.f, a1 = l1 > 0.f ? fast_exp2(m1 - m) : 0.f;
            const float lt = l0 * a0 + l1 * a1;
            const float inv = lt > 0.f ? 1.f / lt : 0.f;
            const float w0 = a0 * inv, w1 = a1 * inv;
            const bool live = NR == 1 && n_pairs > 0;
            const int rp = rt0 * 128 + row;
            if (wgi == 0 && rp < R) {
                d.part_lse[part_base(slot, rp)] = (live && inv > 0.f) ? (m + log2f(lt)) * 0.6931471805599453f : -INFINITY;
                if (empty_slot >= 0) d.part_lse[part_base(empty_slot, rp)] = -INFINITY;
            }
            if (rt0 * 128 + quad * 32 < R) {              // warp-uniform
                const int col = wgi * 32;
                uint32_t o0[32], o1[32];
                if (live) {
                    LS_TMEM_LD32(lane_addr + COL_O + col, o0);
                    LS_TMEM_LD32(lane_addr + COL_O + 128 + col, o1);
                    tmem_wait_ld();
                }
                float* ps = stage + row * OUT_STRIDE + col;
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    float x[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        // an accumulator that never received a PV holds stale TMEM bits: select, do not multiply
                        const float x0 = (live && w0 > 0.f) ? __uint_as_float(o0[e + u]) * w0 : 0.f;
                        const float x1 = (live && w1 > 0.f) ? __uint_as_float(o1[e + u]) * w1 : 0.f;
                        x[u] = x0 + x1;
                    }
                    *reinterpret_cast<float4*>(ps + e) = make_float4(x[0], x[1], x[2], x[3]);
                }
            }
            n_stage = nr_out > 0 ? 1 : 0;
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        // coalesced write-out: one staged row (512 bytes) per warp instruction
        {
            const int sw = warp - 4;                                 // 0..15
            for (int r = sw; r < n_stage * 128; r += 16) {
                const int rp = rt0 * 128 + r;
                if (rp >= R) break;                                  // rows ascend with r: nothing live behind
                const float4 v = *reinterpret_cast<const float4*>(stage + r * OUT_STRIDE + lane * 4);
                *reinterpret_cast<float4*>(d.part_o + part_base(slot, rp) * 128 + lane * 4) = v;
            }
        }
        // ---- fused combine (ls_attn_desc::fuse_sync): the CTAs of this kv head meet at a counter barrier, then share
        // the head's rows out and merge the n_slots partial slots of each - what ls_attn_combine would do in a second
        // launch.  The partials are L2-resident; they were written by other SMs, hence the .cg loads.  All CTAs of the
        // grid are resident at once (the host only sets `fuse` when the grid fits the device in one wave), so the
        // spin cannot starve a CTA that has yet to arrive.
        if (fuse & 2) {
            uint32_t* arrive = d.fuse_sync + kvh;
            uint32_t* depart = d.fuse_sync + d.n_kv_heads + kvh;
            __threadfence();
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (tid == 128) {
                atomicAdd(arrive, 1u);
                const long long t0 = clock64();
                while (*reinterpret_cast<volatile uint32_t*>(arrive) < gridDim.x) {
                    if (clock64() - t0 > 40000000000ll) __trap();        // ~20 s: a CTA of this head never arrived
                }
                __threadfence();
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            const int sw = warp - 4;
            T* out = reinterpret_cast<T*>(d.out);
            const size_t pstride = static_cast<size_t>(d.n_kv_heads) * Rtot;       // rows between slots
            for (int r = bx * 16 + sw; r < Rtot; r += gridDim.x * 16) {
                const int i = r / G, g = r - i * G;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < n_q) {
                    const size_t row0 = static_cast<size_t>(kvh) * Rtot + r;
                    const float lse = lane < n_slots ? __ldcg(d.part_lse + lane * pstride + row0) : -INFINITY;   // n_slots <= 32
                    const float mx = warp_max(lse);
                    if (mx != -INFINITY) {
                        const float w_mine = lse == -INFINITY ? 0.f : __expf(lse - mx);
                        const float wsum = warp_sum(w_mine);
                        const float* base = d.part_o + row0 * 128 + lane * 4;
#pragma unroll 4
                        for (int k = 0; k < n_slots; ++k) {
                            const float w = __shfl_sync(0xffffffffu, w_mine, k);
                            const float4 v = w != 0.f ? __ldcg(reinterpret_cast<const float4*>(base + k * pstride * 128))
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);      // empty slots carry no O row
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
                *reinterpret_cast<uint2*>(out + (static_cast<size_t>(i) * d.n_heads + kvh * G + g) * 128 + lane * 4) = pk;
            }
            // the last CTA to leave resets the counters for the next launch (stream-ordered behind this one)
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (tid == 128 && atomicAdd(depart, 1u) == gridDim.x - 1) {
                *arrive = 0;
                *depart = 0;
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

// ======================================================================================================================
// attn_tc1_kernel: ONE row tile per CTA, for calls with three row tiles (the Llama-3 / QwQ verify shapes).
//
// The two-row-tile kernel above is bound by three costs that sit within 10 % of each other: the
// shared-memory operand port under the SS QK^T MMAs, the S -> softmax -> P -> PV round trip through a buffer that P
// shares with S, and MUFU.  With one row tile per CTA there is TMEM room to remove the first two:
//
//   TMEM (512 columns) = O accumulators 2 x 128 | S 128 | P 64 | Q 64
//
//   * Q lives in TMEM (written once by the softmax warps with tcgen05.st): QK^T is a TS MMA that reads only the 4 KB K
//     slice from shared memory per K step (8 KB with Q in shared memory), and Q needs no staging in shared memory.
//   * S is released as soon as its 128 x 128 scores sit in registers (S_FREE) and P has columns of its own, so QK(p+1)
//     and QK(p+2) run while pair p is still in softmax: the S -> softmax -> P -> PV chain no longer gates QK^T.
//     P(p) is stored only once PV(p-1) has drained the P columns (PV_DONE) - the exponentials are already in registers.
//   * Pairs alternate between the two accumulators (= softmax warpgroup pairs b = p & 1), each with its own reference
//     maxima, exactly like the single-row-tile mode above; the epilogue merges the two partial states.
//
//   * Warp roles: warp 0 streams the K tiles and warp 2 the V tiles (two independent 3-stage rings - one loop for both
//     would tie the K ring to the V ring, and QK^T runs up to two pairs ahead of PV), warp 1 issues QK^T and warp 3 PV
//     (nothing but mbarriers orders the two MMA streams, so one warp's barrier polls and commits overlap the other's
//     MMAs: the tensor pipe queues barely one instruction - tools/ubench/mma_rate), warps 4-19 are the softmax warps.
//
// A trio = row tiles 0, 1, 2 over the SAME key tiles (the range of two partial slots).  Every CTA now needs every K / V
// tile of its trio, so with CLUSTER the CTAs of row tiles 0 and 1 form a cluster of two (one TPC): each loads one
// 64-element half of every tile and TMA-multicasts it into both shared memories; a stage is free again when BOTH
// CTAs' MMAs have released it (multicast tcgen05.commit onto both EMPTY barriers, count 2).  The row-tile-2 CTAs are
// clustered in twos as well (the launch attribute is grid-wide) but run independently.  L2 -> SM traffic is then twice
// the algorithmic bytes as in the trio kernel (three times without clusters - L2 runs at 28 % of its peak, so the default
// is the plain launch: profiles/r02b_attn_tc1_ab.txt).
namespace one {
constexpr int STAGES1 = 3;
constexpr int OFF_K1 = 0;
constexpr int OFF_V1 = OFF_K1 + STAGES1 * TILE_BYTES;
constexpr int OFF_STATE1 = OFF_V1 + STAGES1 * TILE_BYTES;
constexpr int OFF_BAR1 = OFF_STATE1 + STATE_BYTES;
constexpr int OFF_SPECIAL1 = OFF_BAR1 + BAR_BYTES;
constexpr int SMEM1_BYTES = OFF_SPECIAL1 + SPECIAL_BYTES + 1024;
static_assert(SMEM1_BYTES <= 232448, "attn_tc1_kernel: shared memory exceeds the 227 KB opt-in limit");
static_assert(128 * OUT_STRIDE * 4 <= OFF_STATE1, "epilogue staging must fit in the K / V rings");
constexpr uint32_t COL1_O = 0;                     // 2 x 128
constexpr uint32_t COL1_S = 256;                   // 128 fp32 score columns
constexpr uint32_t COL1_P = 384;                   // 128 keys x 16 bit = 64 columns
constexpr uint32_t COL1_Q = 448;                   // 128 head dims x 16 bit = 64 columns
enum { K_FULL = 0, K_EMPTY = 3, V_FULL = 6, V_EMPTY = 9, S_FULL = 12, S_FREE = 14, P_FULL = 16, PV_DONE = 18, ALL_DONE = 20, N_BARS = 21 };
static_assert(N_BARS * 8 + 8 <= BAR_BYTES, "barrier block too small");
// S_FREE / P_FULL arrivals of a softmax warpgroup pair: one per warp (lane 0, behind the warp-collective tcgen05.wait)
// instead of one per thread - 32 x fewer shared-memory atomics on the two hand-offs that sit on the critical path.
#ifndef LS_TC1_THREAD_ARRIVE
constexpr uint32_t ARRIVALS1 = 8;
__device__ __forceinline__ void group_arrive(uint32_t bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
#else
constexpr uint32_t ARRIVALS1 = 256;
__device__ __forceinline__ void group_arrive(uint32_t bar, int) { mbar_arrive(bar); }
#endif
}  // namespace one

template <typename T, bool CLUSTER>
__global__ void __launch_bounds__(THREADS, 1)
attn_tc1_kernel(const ls_attn_desc d, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                int n_slots, int fuse) {
    using namespace one;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment (SWIZZLE_128B atoms) of the SHARED-WINDOW offset, not of the generic pointer: in a cluster
    // launch the generic address of a CTA's shared memory depends on its rank, and the CTA-relative offsets of the two
    // CTAs of a multicast pair must agree.
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bx = blockIdx.x, kvh = blockIdx.y;
    LS_TRACE_INIT
    const int G = d.n_heads / d.n_kv_heads;
    const int n_q = min(*d.n_q_ptr, d.q_rows);
    const int R = n_q * G;
    const int kv_len = *d.kv_len_ptr + d.kv_len_bias;
    const int total_tiles = tc_full_tiles(d, kv_len, n_q);
    SpecialTile* sp_tiles = reinterpret_cast<SpecialTile*>(smem + OFF_SPECIAL1);
    const int n_special = (fuse & 1) ? plan_special_tiles(d, kv_len, total_tiles, sp_tiles, tid) : 0;
    const int all_tiles = total_tiles + n_special;

    // ---- role: trio g = the key tiles of slots 2g, 2g+1; CTAs [0, 2T) = row tiles 0 / 1 (bx & 1) of trio bx >> 1,
    // CTAs [2T, 3T) = row tile 2 of trio bx - 2T.  The merged state goes to slot 2g, slot 2g+1 is written empty.
    const int n_trios = (n_slots + 1) >> 1;
    const bool light = bx >= 2 * n_trios;
    const int g = light ? bx - 2 * n_trios : bx >> 1;
    const int rt = light ? 2 : (bx & 1);
    const int share0 = 2 * g;
    const int nshare = min(2, n_slots - share0);
    const int slot = share0, empty_slot = nshare == 2 ? share0 + 1 : -1;
    const bool mc = CLUSTER && !light;                               // this CTA shares its K / V tiles with its cluster peer
    const uint32_t cta_rank = static_cast<uint32_t>(bx & 1);         // rank in the cluster of two (cluster dims (2, 1, 1))
    const bool rows_live = rt * 128 < R;                             // this CTA's row tile has rows at run time
    // a multicast pair runs its pipeline whenever the PAIR has live rows (row tile 0): the peer depends on this CTA's loads
    const bool run = mc ? R > 0 : rows_live;
    const int per = (all_tiles + n_slots - 1) / n_slots;
    const int tile_lo = min(share0 * per, all_tiles);
    const int tile_hi = min((share0 + nshare) * per, all_tiles);
    const int n_pairs = run ? tile_hi - tile_lo : 0;                 // one pair per key tile; pair p -> accumulator p & 1

    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR1 + N_BARS * 8);
    const uint32_t bar0 = sbase + OFF_BAR1;
    auto bar = [&](int i) { return bar0 + i * 8; };

    if (tid == 0) {
        for (int i = 0; i < STAGES1; ++i) {
            mbar_init(bar(K_FULL + i), 1); mbar_init(bar(K_EMPTY + i), mc ? 2 : 1);
            mbar_init(bar(V_FULL + i), 1); mbar_init(bar(V_EMPTY + i), mc ? 2 : 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(S_FULL + i), 1); mbar_init(bar(S_FREE + i), ARRIVALS1);
            mbar_init(bar(P_FULL + i), ARRIVALS1); mbar_init(bar(PV_DONE + i), 1);
        }
        mbar_init(bar(ALL_DONE), 2);                 // the QK issuer and the PV issuer
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (CLUSTER) cluster_sync_all();               // the peer's barriers are initialised before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    constexpr int FMT = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
    constexpr uint32_t IDESC_QK = make_idesc(128, 128, FMT, 0, 0);
    constexpr uint32_t IDESC_PV = make_idesc(128, 128, FMT, 0, 1);

    // one ring of STAGES1 tiles (K or V): executed by a whole converged warp, one elected lane issues
    auto tma_ring = [&](int full0, int empty0, int off, const CUtensorMap* map) {
        int s = 0, ph = 0;
        for (int j0 = 0; j0 < n_pairs; j0 += 32) {
            // the page-table lookups of the next 32 tiles in one round trip (one per lane) instead of one dependent
            // global load in front of every TMA request
            int row_l = 0;
            if (j0 + lane < n_pairs) {
                const int gi = tile_lo + j0 + lane;
                const int tok = gi >= total_tiles ? sp_tiles[gi - total_tiles].tok : gi * BN;
                const int page = __ldg(d.page_table + (tok >> 7));
                row_l = (page * d.n_kv_heads + kvh) * LS_PAGE + (tok & (LS_PAGE - 1));
            }
            const int cnt = min(32, n_pairs - j0);
            for (int k = 0; k < cnt; ++k) {
                const int row = __shfl_sync(0xffffffffu, row_l, k);
                mbar_wait(bar(empty0 + s), ph ^ 1);
                if (elect_one()) {
                    const uint32_t dst = sbase + off + s * TILE_BYTES;
                    mbar_expect_tx(bar(full0 + s), TILE_BYTES);
                    if (mc) {
                        tma_load_2d_mc(dst + cta_rank * 16384, map, 64 * cta_rank, row, bar(full0 + s), 3);
                    } else {
                        tma_load_2d(dst, map, 0, row, bar(full0 + s));
                        tma_load_2d(dst + 16384, map, 64, row, bar(full0 + s));
                    }
                }
                __syncwarp();
                if (++s == STAGES1) { s = 0; ph ^= 1; }
            }
        }
    };

    if (warp < 4) {
        reg_dec<REGS_CTRL>();
        if (warp == 0) {
            // ================= TMA producers: warp 0 streams the K tiles, warp 2 the V tiles =================
            // (one loop for both would tie the K ring to the V ring: K(j) could only be requested once PV(j-4) has
            // released its V stage, although QK^T runs up to two pairs ahead of PV)
            tma_ring(K_FULL, K_EMPTY, OFF_K1, &map_k);
        } else if (warp == 2) {
            tma_ring(V_FULL, V_EMPTY, OFF_V1, &map_v);
        } else if (warp == 1) {
            // ================= QK issuer (converged warp, one elected lane issues) =================
            // S is a single buffer: QK(q) may only overwrite it once the softmax warps of pair q-1 hold its scores in
            // registers (S_FREE).  Nothing else orders QK against PV: the two MMA streams are issued by two warps, so
            // one warp's barrier polls and commits (~ 500 cycles per pair, during which a single issuer feeds nothing
            // into a tensor pipe that queues barely one instruction) overlap the other warp's MMAs.
            const uint32_t k_lo = (((sbase + OFF_K1) >> 4) & 0x3FFFu) | LBO_K;          // 14-bit address field: in a cluster the
                                                                                        // shared address carries the CTA rank above it
            asm volatile("bar.sync %0, %1;" ::"n"(BAR_Q_READY), "n"(512 + 32) : "memory");     // Q sits in TMEM
            int sq = 0, phq = 0;                          // K stage / parity
            for (int q = 0; q < n_pairs; ++q) {
                mbar_wait(bar(K_FULL + sq), phq);
                if (q >= 1) mbar_wait(bar(S_FREE + ((q - 1) & 1)), ((q - 1) >> 1) & 1);
                tc_fence_after();
                if (lane == 0) LS_TRACE(7, q);
                if (elect_one()) {
                    const uint32_t b0 = k_lo + sq * (TILE_BYTES >> 4);
                    if (!LS_DBG(2))
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        tc_mma_ts(tmem + COL1_S, tmem + COL1_Q + kk * 8, b0 + (kk >> 2) * (16384 >> 4) + (kk & 3) * 2, IDESC_QK, kk > 0 ? 1u : 0u);
                    tc_commit(bar(S_FULL + (q & 1)));
                    if (mc) tc_commit_mc(bar(K_EMPTY + sq), 3); else tc_commit(bar(K_EMPTY + sq));
                }
                __syncwarp();
                if (lane == 0) LS_TRACE(6, q);
                if (++sq == STAGES1) { sq = 0; phq ^= 1; }
            }
            if (n_pairs > 0 && elect_one()) tc_commit(bar(ALL_DONE));
            __syncwarp();
        } else if (warp == 3) {
            // ================= PV issuer =================
            const uint32_t v_lo = (((sbase + OFF_V1) >> 4) & 0x3FFFu) | LBO_V;
            int sv_ = 0, phv = 0;                         // V stage / parity
            for (int p = 0; p < n_pairs; ++p) {
                mbar_wait(bar(V_FULL + sv_), phv);
                if (!LS_DBG(8)) mbar_wait(bar(P_FULL + (p & 1)), (p >> 1) & 1);
                tc_fence_after();
                if (lane == 0) LS_TRACE(0, p);
                if (elect_one()) {
                    const uint32_t b0 = v_lo + sv_ * (TILE_BYTES >> 4);
                    const uint32_t dst = tmem + COL1_O + (p & 1) * 128;
                    if (!LS_DBG(1))
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        tc_mma_ts(dst, tmem + COL1_P + kk * 8, b0 + kk * (2048 >> 4), IDESC_PV, (p >= 2 || kk > 0) ? 1u : 0u);
                    tc_commit(bar(PV_DONE + (p & 1)));
                    if (mc) tc_commit_mc(bar(V_EMPTY + sv_), 3); else tc_commit(bar(V_EMPTY + sv_));
                }
                __syncwarp();
                if (lane == 0) LS_TRACE(1, p);
                if (++sv_ == STAGES1) { sv_ = 0; phv ^= 1; }
            }
            if (n_pairs > 0 && elect_one()) tc_commit(bar(ALL_DONE));
            __syncwarp();
        }
    } else {
        reg_inc<REGS_SOFTMAX>();
        const int wgi = (warp - 4) >> 2;
        const int b = wgi >> 1, h = wgi & 1;
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = tmem + (static_cast<uint32_t>(quad * 32) << 16);
        const int rp0 = rt * 128;
        // Q -> TMEM: thread (warpgroup, lane quarter, lane) owns 32 head dims (16 packed columns) of its row
        if (n_pairs > 0) {
            uint32_t qr[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) qr[e] = 0u;
            const int rp = rp0 + row;
            if (rp < R) {
                const T* __restrict__ q = reinterpret_cast<const T*>(d.q);
                const size_t q_stride = d.q_stride > 0 ? static_cast<size_t>(d.q_stride) : static_cast<size_t>(d.n_heads) * 128;
                const int i = rp / G, gg = rp - i * G;
                const uint4* src = reinterpret_cast<const uint4*>(q + static_cast<size_t>(i) * q_stride + static_cast<size_t>(kvh * G + gg) * 128 + wgi * 32);
                const uint4 a0 = src[0], a1 = src[1], a2 = src[2], a3 = src[3];
                qr[0] = a0.x; qr[1] = a0.y; qr[2] = a0.z; qr[3] = a0.w;
                qr[4] = a1.x; qr[5] = a1.y; qr[6] = a1.z; qr[7] = a1.w;
                qr[8] = a2.x; qr[9] = a2.y; qr[10] = a2.z; qr[11] = a2.w;
                qr[12] = a3.x; qr[13] = a3.y; qr[14] = a3.z; qr[15] = a3.w;
            }
            LS_TMEM_ST16(lane_addr + COL1_Q + wgi * 16, qr);
            tmem_wait_st();
        }
        tc_fence_before();
        asm volatile("bar.arrive %0, %1;" ::"n"(BAR_Q_READY), "n"(512 + 32) : "memory");

        // ================= softmax warpgroups (wgi = 2b + h: pairs of parity b, 64-key half h) =================
        const float c2 = d.scale * 1.4426950408889634f;
        const uint64_t c2x2 = pack2(c2, c2);
        const bool warp_valid = rp0 + quad * 32 < R;                 // warp-uniform, the same in both warps of a row
        const uint32_t o_half = lane_addr + COL1_O + b * 128 + h * 64;
        const uint32_t s_half = lane_addr + COL1_S + h * 64;         // this thread's 64 score columns
        const uint32_t p_half = lane_addr + COL1_P + h * 32;         // ... and its 32 packed probability columns
        const int pair_bar = 2 + b * 4 + quad;
        float* xch = reinterpret_cast<float*>(smem + OFF_STATE1);    // [2 accumulators][2 halves][128 rows]
        float m_ref = -INFINITY, l = 0.f;
        int it = 0;
        for (int p = b; p < n_pairs && !LS_DBG(8); p += 2, ++it) {
            mbar_wait(bar(S_FULL + b), it & 1);
            if (!warp_valid || LS_DBG(4)) {
                // Padding rows only: keep the barrier protocol, skip the math (their O rows are never read) - at the
                // live warps' pace: S(p+2) can be ready before pair p has finished, and an early P_FULL arrival for
                // pair p+2 would be counted into the phase of pair p.  PV(p-1) complete implies P_FULL(p-2) complete.
                group_arrive(bar(S_FREE + b), lane);
                if (p >= 1) mbar_wait(bar(PV_DONE + (b ^ 1)), ((p - 1) >> 1) & 1);
                group_arrive(bar(P_FULL + b), lane);
                continue;
            }
            tc_fence_after();
            if (quad == 0 && lane == 0 && h == 0) LS_TRACE(2, p);
            uint32_t sv[64];
            LS_TMEM_LD32(s_half, sv);
            LS_TMEM_LD32(s_half + 32, (sv + 32));
            tmem_wait_ld();
            tc_fence_before();
            group_arrive(bar(S_FREE + b), lane);                            // the scores are in registers: QK(p+1) may overwrite S
            if (quad == 0 && lane == 0 && h == 0) LS_TRACE(3, p);
            const int gi = tile_lo + p;
            if (gi >= total_tiles) {
                const SpecialTile st = sp_tiles[gi - total_tiles];
                const uint64_t vis = special_visibility(d, st, rp0 + row, R, G, kv_len, n_q, h);
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
            uint32_t any;
            {
                const uint32_t need = mx > m_ref + 8.0f ? 1u : 0u;   // lazy rescale threshold (2^8)
                asm volatile(
                    "{\n\t"
                    ".reg .pred pin, pout;\n\t"
                    "setp.ne.u32 pin, %1, 0;\n\t"
                    "bar.red.or.pred pout, %2, 64, pin;\n\t"
                    "selp.u32 %0, 1, 0, pout;\n\t"
                    "}" : "=r"(any) : "r"(need), "r"(pair_bar) : "memory");
            }
            if (any) {
                xch[(b * 2 + h) * 128 + row] = mx;
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                const float joint = fmaxf(mx, xch[(b * 2 + (h ^ 1)) * 128 + row]);
                const float m_new = joint > m_ref + 8.0f ? joint : m_ref;
                const float alpha = m_ref == -INFINITY ? 0.f : fast_exp2(m_ref - m_new);
                l *= alpha;
                m_ref = m_new;
                if (it >= 1) {
                    // The accumulator's previous PV is pair p-2.  It was issued after QK(p), so S_FULL(p) does not
                    // certify it: wait for its own completion.  PV(p-4) is known complete (P(p-2) was only stored
                    // after PV(p-3), and MMAs complete in order) and PV(p) cannot have run, so the parity is unambiguous.
                    mbar_wait(bar(PV_DONE + b), (it - 1) & 1);
                    tc_fence_after();
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t ov[16];
                        LS_TMEM_LD16(o_half + c * 16, ov);
                        tmem_wait_ld();
#pragma unroll
                        for (int e = 0; e < 16; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
                        LS_TMEM_ST16(o_half + c * 16, ov);
                    }
                    tmem_wait_st();
                }
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            }
        
            const float m_use = m_ref == -INFINITY ? 0.f : m_ref;
            uint32_t pk[32];
            if (gi >= total_tiles) l += exp_regs<T, 0>(sv, pk, c2x2, pack2(-m_use, -m_use));
            else l += exp_regs<T, LS_TC_POLY>(sv, pk, c2x2, pack2(-m_use, -m_use));
            if (quad == 0 && lane == 0 && h == 0) LS_TRACE(4, p);
            // The P columns are free once PV(p-1) has read them.  Both PV_DONE barriers are waited on phase by phase
            // (this one by the other accumulator's warpgroups), so the parity cannot alias.
            if (p >= 1) {
                mbar_wait(bar(PV_DONE + (b ^ 1)), ((p - 1) >> 1) & 1);
                tc_fence_after();
            }
            LS_TMEM_ST16(p_half, pk);
            LS_TMEM_ST16(p_half + 16, (pk + 16));
            tmem_wait_st();
            tc_fence_before();
            group_arrive(bar(P_FULL + b), lane);
            if (quad == 0 && lane == 0 && h == 0) LS_TRACE(5, p);
        }
        // ================= epilogue: merge the two accumulators, normalised partial O + LSE =================
        const int Rtot = d.q_rows * G;
        if (n_pairs > 0) {
            mbar_wait(bar(ALL_DONE), 0);
            tc_fence_after();
        }
        auto part_base = [&](int sl, int rp) { return (static_cast<size_t>(d.part_offset + sl) * d.n_kv_heads + kvh) * Rtot + rp; };
        float* stage = reinterpret_cast<float*>(smem);               // [128 rows][OUT_STRIDE]: the K ring is dead (every MMA has completed)
        asm volatile("bar.sync 1, 512;" ::: "memory");
        float* st_m = xch;                                           // [2][128]
        float* st_l = xch + 2 * 128;                                 // [2][2][128]
        if (h == 0) st_m[b * 128 + row] = m_ref;
        st_l[(b * 2 + h) * 128 + row] = l;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const float m0 = st_m[row], m1 = st_m[128 + row];
        const float l0 = st_l[row] + st_l[128 + row], l1 = st_l[256 + row] + st_l[384 + row];
        const float m = fmaxf(m0, m1);
        const float a0 = l0 > 0.f ? fast_exp2(m0 - m) : 0.f, a1 = l1 > 0.f ? fast_exp2(m1 - m) : 0.f;
        const float lt = l0 * a0 + l1 * a1;
        const float inv = lt > 0.f ? 1.f / lt : 0.f;
        const float w0 = a0 * inv, w1 = a1 * inv;
        const bool live = n_pairs > 0;
        const int rp = rp0 + row;
        if (wgi == 0 && rp < R) {
            d.part_lse[part_base(slot, rp)] = (live && inv > 0.f) ? (m + log2f(lt)) * 0.6931471805599453f : -INFINITY;
            if (empty_slot >= 0) d.part_lse[part_base(empty_slot, rp)] = -INFINITY;
        }
        if (warp_valid) {                                        // warp-uniform (tcgen05.ld is warp-collective)
            const int col = wgi * 32;
            uint32_t o0[32], o1[32];
            if (live) {
                LS_TMEM_LD32(lane_addr + COL1_O + col, o0);
                LS_TMEM_LD32(lane_addr + COL1_O + 128 + col, o1);
                tmem_wait_ld();
            }
            float* ps = stage + row * OUT_STRIDE + col;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
                float x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    // an accumulator that never received a PV holds stale TMEM bits: select, do not multiply
                    const float x0 = (live && w0 > 0.f) ? __uint_as_float(o0[e + u]) * w0 : 0.f;
                    const float x1 = (live && w1 > 0.f) ? __uint_as_float(o1[e + u]) * w1 : 0.f;
                    x[u] = x0 + x1;
                }
                *reinterpret_cast<float4*>(ps + e) = make_float4(x[0], x[1], x[2], x[3]);
            }
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        {
            const int sw = warp - 4;                                 // 0..15
            for (int r = sw; r < 128; r += 16) {
                const int rp = rp0 + r;
                if (rp >= R) break;
                const float4 v = *reinterpret_cast<const float4*>(stage + r * OUT_STRIDE + lane * 4);
                *reinterpret_cast<float4*>(d.part_o + part_base(slot, rp) * 128 + lane * 4) = v;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (CLUSTER) cluster_sync_all();               // the peer may still multicast commits at this CTA's barriers until it is done too
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

// A tensor map encodes nothing but (address, extent, dtype): an entry for a freed-and-reallocated pool with the same
// key is still the right map, so stale entries are harmless; the cache is only bounded so that a long-lived process
// that keeps allocating pools cannot grow it without limit.
constexpr size_t MAP_CACHE_MAX = 4096;
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
    if (cache.size() >= MAP_CACHE_MAX) cache.clear();
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
// Can this call merge its partial slots in the kernel's tail?  (tcgen05 path, all slots written by this launch and merged by
// nobody else, <= 32 slots (one LSE per lane), and a grid that is resident in one wave: the kv-head barrier spins.)
static bool tc_can_fuse_combine(const ls_attn_desc* d, int n_splits_tc) {
    if (!d->fuse_sync || !d->out || d->part_offset != 0 || d->n_parts_total != n_splits_tc || n_splits_tc > 32) return false;
    const int G = d->n_heads / d->n_kv_heads;
    const int nrs = (d->q_rows * G + 127) / 128;
    const int gx = nrs > ls::tc::NRL_MAX ? n_splits_tc + (n_splits_tc + 1) / 2 : n_splits_tc;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
    return gx * d->n_kv_heads <= sms;
}

// Which kernel serves calls with three row tiles: 0 = attn_tc_kernel (heavy / heavy / light trios), 1 = attn_tc1_kernel
// (one row tile per CTA; measured fastest: default), 2 = attn_tc1_kernel as clusters of two with TMA multicast (a third
// less L2 -> SM traffic, same speed at 128k, but 3 us of cluster launch / barriers per call).  LONGSPEC_TC1 sets the initial value.
static int g_tc1_mode = -1;
static int tc1_mode() {
    if (g_tc1_mode < 0) {
        const char* e = getenv("LONGSPEC_TC1");
        const int m = e ? atoi(e) : 1;
        g_tc1_mode = m < 0 ? 0 : (m > 2 ? 2 : m);
    }
    return g_tc1_mode;
}
extern "C" int ls_attn_tc1_mode(int mode) {
    const int prev = tc1_mode();
    if (mode >= 0 && mode <= 2) g_tc1_mode = mode;
    return prev;
}

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
    static unsigned long long smem_ok_f16 = 0, smem_ok_bf16 = 0;     // one bit per device ordinal
    rc = d->dtype == LS_F16 ? ensure_dyn_smem(attn_tc_kernel<__half>, SMEM_BYTES, &smem_ok_f16, "ls_attn_partial_tc")
                            : ensure_dyn_smem(attn_tc_kernel<__nv_bfloat16>, SMEM_BYTES, &smem_ok_bf16, "ls_attn_partial_tc");
    if (rc) return rc;
    const int nrs = (d->q_rows * G + 127) / 128;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int fuse = (fuse_remainder ? 1 : 0) | ((fuse_remainder && tc_can_fuse_combine(d, n_splits_tc)) ? 2 : 0);
    // Three row tiles: one row tile per CTA (attn_tc1_kernel), as clusters of two when the grid allows it.
    // (see ls_attn_tc1_mode)
    const int mode1 = tc1_mode();
    if (nrs > NRL_MAX && mode1 > 0 && !(fuse & 2)) {
        const int gx = 3 * ((n_splits_tc + 1) / 2);
        const bool cluster = mode1 == 2 && gx % 2 == 0;
        static unsigned long long ok1[4] = {0, 0, 0, 0};
        const int which = (cluster ? 2 : 0) + (d->dtype == LS_F16 ? 0 : 1);
        static const void* const fns[4] = {
            reinterpret_cast<const void*>(attn_tc1_kernel<__half, false>), reinterpret_cast<const void*>(attn_tc1_kernel<__nv_bfloat16, false>),
            reinterpret_cast<const void*>(attn_tc1_kernel<__half, true>), reinterpret_cast<const void*>(attn_tc1_kernel<__nv_bfloat16, true>)};
        const void* fn = fns[which];
        rc = ensure_dyn_smem(fn, one::SMEM1_BYTES, &ok1[which], "ls_attn_partial_tc (one row tile per CTA)");
        if (rc) return rc;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(gx, d->n_kv_heads);
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = one::SMEM1_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = cluster ? 1 : 0;
        ls_attn_desc dd = *d;
        int n_slots = n_splits_tc, fz = fuse;
        void* args[] = {&dd, &mk, &mv, &n_slots, &fz};
        const cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
        if (e != cudaSuccess) { set_error("ls_attn_partial_tc: cluster launch failed: %s", cudaGetErrorString(e)); return LS_ERR_CUDA; }
        return check_launch("ls_attn_partial_tc");
    }
    dim3 grid(nrs > NRL_MAX ? n_splits_tc + (n_splits_tc + 1) / 2 : n_splits_tc, d->n_kv_heads);
    if (d->dtype == LS_F16)
        attn_tc_kernel<__half><<<grid, THREADS, SMEM_BYTES, st>>>(*d, mk, mv, n_splits_tc, fuse);
    else
        attn_tc_kernel<__nv_bfloat16><<<grid, THREADS, SMEM_BYTES, st>>>(*d, mk, mv, n_splits_tc, fuse);
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

// does ls_attn_partial_tc_fused(d) merge the slots itself?  (see ls_attn_desc::fuse_sync)
extern "C" int ls_attn_tc_fuses_combine(const ls_attn_desc* d) { return tc_can_fuse_combine(d, d->n_splits) ? 1 : 0; }

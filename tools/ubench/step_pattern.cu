// Micro-benchmark: the per-step MMA pattern of attn_tc.cu - a burst of 4 TS PV MMAs (M128 N128 K16, A = P in TMEM,
// B = V MN-major in smem) followed by 8 SS QK MMAs (M128 N<BN> K16) - issued by one thread,
//   mode 0: bursts back to back (pure tensor-pipe throughput of the pattern)
//   mode 1: commit + mbarrier wait after every burst (adds the issue->completion latency the chained kernel sees)
// Variants: BN = 64 (8 QK MMAs) or BN = 128 (8 QK MMAs N=128 + 8 PV MMAs: a 128-key step).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int fmt, int a_mn, int b_mn) {
    return (1u << 4) | (uint32_t(fmt) << 7) | (uint32_t(fmt) << 10) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) |
           (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.b32 p, %5, 0;\n mov.b64 da, {%1, %3};\n mov.b64 db, {%2, %3};\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_t, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n .reg .pred p;\n .reg .b64 db;\n setp.ne.b32 p, %5, 0;\n mov.b64 db, {%2, %3};\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n}" ::"r"(d), "r"(a_t), "r"(b_lo), "r"(DESC_HI), "r"(idesc), "r"(acc) : "memory");
}
__device__ int RANDOM_DATA = 0;
template <int MODE, int BN, int PV_TS, int WIDE = 0, int SPIN = 0, int ELECT = 0, int ALT = 0>
__global__ void k(int bursts, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, endbar;
    __shared__ uint32_t slot;
    const uint32_t sb = smem_u32(smem);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&endbar))); }
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
        uint32_t v = 0x3c003c00u;
        if (RANDOM_DATA) {   // pseudo-random bf16 pairs: random sign/mantissa, exponents around 2^-3 .. 2^0
            uint32_t x = (i + 1) * 2654435761u; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
            const uint32_t lo = ((x & 0x8000u) | (0x3e00u + ((x >> 3) & 0x1ffu))), hi = (((x >> 16) & 0x8000u) | (0x3e00u + ((x >> 19) & 0x1ffu)));
            v = lo | (hi << 16);
        }
        ((uint32_t*)smem)[i] = v;
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x < 32) {
        const uint32_t a_lo = (sb >> 4) | (1u << 16);                       // Q tile (K-major)
        const uint32_t p_lo = ((sb + 32768) >> 4) | (1u << 16);            // P tile in smem (SS-mode PV)
        const uint32_t bk_lo = ((sb + 65536) >> 4) | (1u << 16);           // K tile
        const uint32_t bv_lo = ((sb + 98304) >> 4) | (((WIDE ? 16384u : 8192u) >> 4) << 16); // V tile (MN-major)
        const uint32_t idk = make_idesc(128, BN, 1, 0, 0), idv = make_idesc(128, 128, 1, 0, 1);
        uint32_t el;
        asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.b32 %0, 1, 0, P;\n}" : "=r"(el));
        uint32_t ph = 0;
        long long t0 = clock64();
        for (int i = 0; i < bursts; ++i) {
            if (ELECT) asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.b32 %0, 1, 0, P;\n}" : "=r"(el));
            if (el) {
                const uint32_t sbuf = tmem + (i & 1) * BN;      // S/P buffer of this step
#pragma unroll
                for (int kk = 0; kk < BN / 16; ++kk) {           // PV(t): K = BN keys
                    if (PV_TS) mma_ts(tmem + 256 + (i % 2) * 128, sbuf + kk * 8, bv_lo + (kk & 3) * 128 + (kk >> 2) * 2048, idv, kk > 0);
                    else mma_ss(tmem + 256 + (i % 2) * 128, p_lo + (kk & 3) * 2 + (kk >> 2) * 1024, bv_lo + (kk & 3) * 128 + (kk >> 2) * 2048, idv, kk > 0);
                }
#pragma unroll
                const uint32_t a_alt = ALT ? a_lo + (i & 1) * 2048 : a_lo;   // ALT: Q tile 0 / Q tile 1 on alternate steps
                for (int kk = 0; kk < 8; ++kk)                   // QK(t+2): K = 128 (head dim)
                    mma_ss(sbuf, a_alt + (kk >> 2) * 1024 + (kk & 3) * 2, bk_lo + (kk >> 2) * (WIDE ? 1024 : BN * 8) + (kk & 3) * 2, idk, kk > 0);
                if (MODE == 1 || i == bursts - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
            if (MODE == 1 || i == bursts - 1) {
                asm volatile("{\n .reg .pred p;\n W:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D;\n bra W;\n D:\n}" ::"r"(smem_u32(&bar)), "r"(ph) : "memory");
                ph ^= 1;
            }
        }
        long long t2 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t2 - t0;
        if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&endbar)) : "memory");
    }
    if (SPIN && threadIdx.x >= 128)
        asm volatile("{\n .reg .pred p;\n WS:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra DS;\n bra WS;\n DS:\n}" ::"r"(smem_u32(&endbar)) : "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
static int GRID = 1;
template <int MODE, int BN, int PV_TS, int WIDE = 0, int SPIN = 0, int ELECT = 0, int ALT = 0>
void run(long long* d) {
    cudaFuncSetAttribute(k<MODE, BN, PV_TS, WIDE, SPIN, ELECT, ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int bursts = 4096;
    k<MODE, BN, PV_TS, WIDE, SPIN, ELECT, ALT><<<GRID, SPIN ? 384 : 128, 200 * 1024>>>(bursts, d);
    k<MODE, BN, PV_TS, WIDE, SPIN, ELECT, ALT><<<GRID, SPIN ? 384 : 128, 200 * 1024>>>(bursts, d);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%s%s%s%s BN=%3d PV=%s %-22s: %7.1f cycles per step (%d PV + 8 QK MMAs) = %.1f per 64 keys   %s\n", WIDE ? "16K-chunk-strides" : "compact", SPIN ? " +8 warps spinning on try_wait" : "", ELECT ? " +elect per burst" : "", ALT ? " +alternating Q tile" : "", BN, PV_TS ? "TS" : "SS",
           MODE ? "wait after each burst" : "back to back", double(h) / bursts, BN / 16, double(h) / bursts * 64 / BN,
           cudaGetErrorString(cudaGetLastError()));
}
int main(int argc, char** argv) {
    long long* d; cudaMalloc(&d, 64);
    if (argc > 1) GRID = atoi(argv[1]);
    if (argc > 2) { int r = atoi(argv[2]); cudaMemcpyToSymbol(RANDOM_DATA, &r, sizeof(r)); printf("random data = %d\n", r); }
    printf("grid = %d CTAs\n", GRID);
    run<0, 64, 1>(d); run<1, 64, 1>(d);
    run<0, 64, 1, 1>(d); run<1, 64, 1, 1>(d);
    run<0, 64, 1, 1, 1, 0>(d); run<1, 64, 1, 1, 1, 0>(d);
    run<0, 64, 1, 1, 0, 1>(d); run<1, 64, 1, 1, 0, 1>(d);
    run<0, 64, 1, 1, 1, 1>(d); run<1, 64, 1, 1, 1, 1>(d);
    run<0, 64, 1, 1, 0, 0, 1>(d); run<1, 64, 1, 1, 0, 0, 1>(d);
    run<0, 64, 0>(d); run<1, 64, 0>(d);
    run<0, 128, 1>(d); run<1, 128, 1>(d);
    run<0, 128, 0>(d); run<1, 128, 0>(d);
    return 0;
}

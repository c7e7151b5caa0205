// Micro-benchmark: how much do concurrently running softmax-like warps slow down the tcgen05.mma stream?
// Warp 1 issues the attn_tc.cu step pattern (4 TS PV MMAs N=128 + 8 SS QK MMAs N=64) back to back; warps 4-11
// loop over   [tcgen05.ld 64 cols] -> [64 MUFU.EX2 + FMAs per thread] -> [tcgen05.st 32 cols]   with each part
// switchable, until the MMA warp is done.  Reports cycles per MMA step and softmax-loop iterations completed.
#include <cstdio>
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
#define LD32(taddr, v)                                                                                       \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"      \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                            \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),     \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),            \
                   "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),          \
                   "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),          \
                   "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                              \
                 : "r"(taddr))
#define ST32(taddr, v)                                                                                       \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"   \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"                               \
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),           \
                   "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),        \
                   "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),     \
                   "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),     \
                   "r"(v[31])                                                                                         \
                 : "memory")
// flags: 1 = TMEM ld, 2 = MUFU/FMA math, 4 = TMEM st, 8 = only one warpgroup (warps 4-7) works
__global__ void k(int bursts, int flags, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint64_t dummy[4];     // [0..2] commit targets, [3] an already-completed barrier to wait on
    __shared__ uint32_t slot;
    __shared__ volatile int done;
    const uint32_t sb = smem_u32(smem);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy[i])));
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&dummy[3])) : "memory");   // phase 0 complete
        done = 0;
    }
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp == 1) {
        const uint32_t a_lo = (sb >> 4) | (1u << 16);
        const uint32_t bk_lo = ((sb + 65536) >> 4) | (1u << 16);
        const uint32_t bv_lo = ((sb + 98304) >> 4) | (((flags & 128) ? 16384u : 8192u) >> 4) << 16;
        const uint32_t kchunk = (flags & 128) ? 1024 : 512;
        const uint32_t idk = make_idesc(128, 64, 1, 0, 0), idv = make_idesc(128, 128, 1, 0, 1);
        uint32_t el;
        asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.b32 %0, 1, 0, P;\n}" : "=r"(el));
        long long t0 = clock64();
        for (int i = 0; i < bursts; ++i) {
            if (flags & 32) {     // two waits on a barrier whose phase is already complete (like P_FULL / V_FULL when ahead)
                asm volatile("{\n .reg .pred p;\n W2:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra D2;\n bra W2;\n D2:\n}" ::"r"(smem_u32(&dummy[3])) : "memory");
                asm volatile("{\n .reg .pred p;\n W3:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra D3;\n bra W3;\n D3:\n}" ::"r"(smem_u32(&dummy[3])) : "memory");
            }
            if (flags & 64) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (flags & 256) asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.b32 %0, 1, 0, P;\n}" : "=r"(el));
            if ((flags & 512) && (threadIdx.x & 31) == 0) out[8 + (i & 63)] = clock64();
            if (el) {
                const uint32_t sbuf = tmem + (i & 3) * 64;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) mma_ts(tmem + 256 + (i & 1) * 128, sbuf + kk * 8, bv_lo + kk * 128, idv, kk > 0);
                if (flags & 16) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[0])) : "memory");
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    mma_ss(sbuf, a_lo + (kk >> 2) * 1024 + (kk & 3) * 2, bk_lo + (kk >> 2) * kchunk + (kk & 3) * 2, idk, kk > 0);
                if (flags & 16) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[1])) : "memory");
                if (i == bursts - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
        }
        asm volatile("{\n .reg .pred p;\n W:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra D;\n bra W;\n D:\n}" ::"r"(smem_u32(&bar)) : "memory");
        long long t2 = clock64();
        if (threadIdx.x == 32) { out[0] = t2 - t0; done = 1; }
    } else if (warp >= 4 && (!(flags & 8) || warp < 8)) {
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((warp - 4) >> 2) * 64;
        uint32_t v[64];
#pragma unroll
        for (int e = 0; e < 64; ++e) v[e] = 0x3f000000u + e + threadIdx.x;
        long long iters = 0;
        while (!done) {
            if (flags & 1) {
                LD32(lane_addr, v);
                LD32(lane_addr + 32, (v + 32));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            if (flags & 2) {
#pragma unroll
                for (int e = 0; e < 64; ++e) {
                    float x = __uint_as_float(v[e]) * 0.125f - 1.0f, y;
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
                    v[e] = __float_as_uint(y + 0.5f);
                }
            }
            if (flags & 4) {
                ST32(lane_addr, v);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            ++iters;
        }
        if (threadIdx.x == 128) out[1] = iters;
        if (v[threadIdx.x & 63] == 0x12345678u) out[2] = 1;
    }
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main() {
    long long* d; cudaMalloc(&d, 1024);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int bursts = 512;
    const char* names[] = {"idle spin (done flag only)", "TMEM ld", "math", "ld+math", "TMEM st", "ld+st", "math+st", "ld+math+st"};
    for (int x : {128, 256, 512, 16 + 32 + 64 + 128 + 256 + 512}) {
        for (int f : {0, 7}) {
            cudaMemset(d, 0, 64);
            k<<<1, 384, 200 * 1024>>>(bursts, x | f, d);
            k<<<1, 384, 200 * 1024>>>(bursts, x | f, d);
            cudaDeviceSynchronize();
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("MMA loop extras %s%s%s%s%s%s, 2 softmax WGs %-12s: %7.1f cycles per MMA step  %s\n", x & 16 ? "commit x2 " : "", x & 32 ? "try_wait x2 " : "",
                   x & 64 ? "fence " : "", x & 128 ? "16K-strides " : "", x & 256 ? "elect " : "", x & 512 ? "gtrace " : "", names[f], double(h[0]) / bursts, cudaGetErrorString(cudaGetLastError()));
        }
    }
    for (int wgs = 1; wgs < 2; ++wgs)
        for (int f = 0; f < 8; ++f) {
            const int flags = f | (wgs ? 0 : 8);
            cudaMemset(d, 0, 64);
            k<<<1, 384, 200 * 1024>>>(bursts, flags, d);
            k<<<1, 384, 200 * 1024>>>(bursts, flags, d);
            cudaDeviceSynchronize();
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%d softmax WG(s), %-28s: %7.1f cycles per MMA step; softmax-loop iterations per MMA step %.2f  %s\n", wgs + 1, names[f],
                   double(h[0]) / bursts, double(h[1]) / bursts, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}

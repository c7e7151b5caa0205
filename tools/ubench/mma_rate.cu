// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) for the operand shapes used by
// attn_tc.cu, issued back to back by one thread; time from first issue to commit arrival.
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
// mode 0: SS, B K-major, N given;  mode 1: SS, B MN-major (V-like), N=128;  mode 2: TS, B MN-major, N=128
template <int mode, int N, int NACC>
__global__ void k(int count, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t sb = smem_u32(smem);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); }
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x < 32) {       // converged warp, one elected lane issues unrolled bursts (like attn_tc.cu)
        const uint32_t a_lo = (sb >> 4) | (1u << 16);
        const uint32_t bk_lo = ((sb + 65536) >> 4) | (1u << 16);
        const uint32_t bv_lo = ((sb + 65536) >> 4) | ((8192u >> 4) << 16);
        const uint32_t idk = make_idesc(128, N, 1, 0, 0), idv = make_idesc(128, 128, 1, 0, 1);
        uint32_t el;
        asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.b32 %0, 1, 0, P;\n}" : "=r"(el));
        long long t0 = clock64();
        for (int i = 0; i < count; i += 8) {
            if (el) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t dcol = tmem + (kk % NACC) * 128;
                    const uint32_t acc = kk >= NACC ? 1u : 0u;
                    if (mode == 0) mma_ss(dcol, a_lo + (kk >> 2) * 1024 + (kk & 3) * 2, bk_lo + (kk >> 2) * 512 + (kk & 3) * 2, idk, acc);
                    if (mode == 1) mma_ss(dcol, a_lo + (kk & 3) * 2, bv_lo + (kk & 3) * 128, idv, acc);
                    if (mode == 2) mma_ts(dcol, tmem + 400 + (kk & 3) * 8, bv_lo + (kk & 3) * 128, idv, acc);
                }
            }
            __syncwarp();
        }
        if (el) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        asm volatile("{\n .reg .pred p;\n W:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra D;\n bra W;\n D:\n}" ::"r"(smem_u32(&bar)) : "memory");
        long long t2 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int mode, int N, int NACC>
void run(const char* name, long long* d) {
    cudaFuncSetAttribute(k<mode, N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int count : {8, 512}) {
        k<mode, N, NACC><<<1, 128, 200 * 1024>>>(count, d);
        k<mode, N, NACC><<<1, 128, 200 * 1024>>>(count, d);
        cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%-28s N=%3d accumulators=%d count=%3d: issue %6lld cyc, done %6lld cyc -> %.1f cyc/MMA\n", name, N, NACC, count, h[0], h[1], double(h[1]) / count);
    }
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    run<0, 64, 1>("SS K-major", d);  run<0, 64, 2>("SS K-major", d);  run<0, 64, 4>("SS K-major", d);
    run<0, 128, 1>("SS K-major", d); run<0, 128, 2>("SS K-major", d); run<0, 128, 4>("SS K-major", d);
    run<0, 256, 1>("SS K-major", d);
    run<1, 128, 1>("SS MN-major B", d); run<1, 128, 2>("SS MN-major B", d);
    run<2, 128, 1>("TS MN-major B", d); run<2, 128, 2>("TS MN-major B", d);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

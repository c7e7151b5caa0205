// Micro-benchmark: warp-instruction throughput of MUFU.EX2, F2FP (bf16x2 pack), FFMA, FMNMX on one SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
    float v[16];
    for (int i = 0; i < 16; ++i) v[i] = -0.001f * (threadIdx.x + i);
    unsigned acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 1) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[i]), "f"(v[(i + 1) & 15])); acc ^= r; }
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(v[i]) : "f"(v[(i + 1) & 15]));
            if (OP == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(v[(i + 3) & 15]));
            if (OP == 4) asm volatile("add.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(v[(i + 3) & 15]));
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 16; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name) {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    for (int warps = 4; warps <= 16; warps *= 2) {
        int iters = 2000;
        k<OP><<<1, warps * 32>>>(out, cyc, iters);
        k<OP><<<1, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double per_smsp = double(c) / (double(iters) * 16 * (warps / 4));
        printf("%-6s warps/SM=%2d cycles=%lld -> %.2f cycles per warp-instruction per SMSP (%.1f lanes/clk/SM)\n", name, warps, c,
               per_smsp, 128.0 / per_smsp);
    }
}
int main() {
    run<0>("EX2"); run<1>("F2FP"); run<2>("FFMA"); run<3>("FMNMX"); run<4>("FADD");
    return 0;
}

// Micro-benchmark of the softmax inner loop of attn_tc.cu (64 scores per thread per step, 8 warps = 2 per SMSP):
// how many of every 8 exponentials should leave the MUFU pipe for a Cody-Waite + degree-3 polynomial on the FMA pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_math softmax_math.cu && ./softmax_math
// Prints cycles per step (64 elements per thread, 256 threads) for POLY = 0..8 of 8 and for ex2.approx.f16x2.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

// 2^x for a packed pair on the FMA pipe: x = n + f, n = round(x), f in [-1/2, 1/2], 2^f by a cubic (max rel err 2^-13.2)
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& p0, float& p1) {
    const float MAGIC = 12582912.f;                       // 1.5 * 2^23
    float x0, x1;
    unpack2(x2, x0, x1);
    x0 = fmaxf(x0, -126.f);
    x1 = fmaxf(x1, -126.f);
    x2 = pack2(x0, x1);
    const uint64_t t2 = add2(x2, pack2(MAGIC, MAGIC));
    const uint64_t n2 = add2(t2, pack2(-MAGIC, -MAGIC));
    const uint64_t f2 = fma2(n2, pack2(-1.f, -1.f), x2);
    uint64_t p2 = fma2(f2, pack2(0.05592204f, 0.05592204f), pack2(0.24264008f, 0.24264008f));
    p2 = fma2(p2, f2, pack2(0.69312102f, 0.69312102f));
    p2 = fma2(p2, f2, pack2(0.99992448f, 0.99992448f));
    float q0, q1, t0, t1;
    unpack2(p2, q0, q1);
    unpack2(t2, t0, t1);
    p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
    p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

// MODE 0..8: that many of every 8 elements use the polynomial; MODE 9: ex2.approx.f16x2 for everything
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(const float* in, float* out, long long* cyc, int iters) {
    uint32_t sv[64];
    for (int e = 0; e < 64; ++e) sv[e] = __float_as_uint(in[(threadIdx.x * 64 + e) & 4095]);
    const float c2 = 0.1275f;
    const uint64_t c2x2 = pack2(c2, c2);
    float m_ref = 3.0f, l = 0.f;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int e = 0; e < 64; ++e) asm volatile("" : "+r"(sv[e]));       // "new scores arrived"
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 64; e += 8) {
            mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])));
            mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])));
            mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sv[e + 4]), __uint_as_float(sv[e + 5])));
            mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sv[e + 6]), __uint_as_float(sv[e + 7])));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c2;
        if (mx > m_ref + 8.f) m_ref = mx;
        const uint64_t negm = pack2(-m_ref, -m_ref);
        uint64_t sum_a = 0ull, sum_b = 0ull;
        uint32_t pk[32];
        if (MODE <= 8) {
#pragma unroll
            for (int e = 0; e < 64; e += 2) {
                const uint64_t x2 = fma2(pack2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), c2x2, negm);
                float p0, p1;
                if ((e & 7) < MODE) {                                   // MODE of every 8 (in pairs; odd MODE rounds up a pair)
                    exp2_poly2(x2, p0, p1);
                } else {
                    float x0, x1;
                    unpack2(x2, x0, x1);
                    p0 = ex2(x0);
                    p1 = ex2(x1);
                }
                if (e & 2) sum_b = add2(sum_b, pack2(p0, p1)); else sum_a = add2(sum_a, pack2(p0, p1));
                pk[e >> 1] = pack_bf16(p0, p1);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 64; e += 2) {
                const uint64_t x2 = fma2(pack2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), c2x2, negm);
                float x0, x1;
                unpack2(x2, x0, x1);
                __half2 h = __floats2half2_rn(x0, x1);
                uint32_t hu = *reinterpret_cast<uint32_t*>(&h), r;
                asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(hu));
                pk[e >> 1] = r;
                const float2 pf = __half22float2(*reinterpret_cast<__half2*>(&r));
                if (e & 2) sum_b = add2(sum_b, pack2(pf.x, pf.y)); else sum_a = add2(sum_a, pack2(pf.x, pf.y));
            }
        }
        float a0, a1;
        unpack2(add2(sum_a, sum_b), a0, a1);
        l += a0 + a1;
#pragma unroll
        for (int e = 0; e < 32; ++e) acc ^= pk[e];      // stands in for the tcgen05.st of P (keeps the packs alive)
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = l + __uint_as_float(acc & 0x3fffffffu);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const float* in, float* out, long long* cyc) {
    const int iters = 2000;
    k<MODE><<<1, 256>>>(in, out, cyc, iters);
    k<MODE><<<1, 256>>>(in, out, cyc, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    if (MODE <= 8) printf("poly %d/8 of the exponentials : %7.1f cycles per step (8 warps x 64 elements per thread)\n", MODE, double(c) / iters);
    else printf("ex2.approx.f16x2 everywhere    : %7.1f cycles per step\n", double(c) / iters);
}

// accuracy of the polynomial against exp2f on the host-visible range
__global__ void acc_k(float* err) {
    float worst = 0.f;
    for (int i = threadIdx.x; i < 1 << 20; i += blockDim.x) {
        const float x = -40.f + 48.f * (i / float(1 << 20));
        float p0, p1;
        exp2_poly2(pack2(x, x - 100.f), p0, p1);
        const float ref = exp2f(x);
        worst = fmaxf(worst, fabsf(p0 - ref) / ref);
    }
    err[threadIdx.x] = worst;
}

int main() {
    float *in, *out, *err;
    long long* cyc;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024); cudaMalloc(&err, 1024 * 4);
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = -20.f + 0.01f * (i % 3000);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>(in, out, cyc); run<2>(in, out, cyc); run<4>(in, out, cyc); run<6>(in, out, cyc); run<8>(in, out, cyc);
    run<9>(in, out, cyc);
    acc_k<<<1, 256>>>(err);
    float he[256];
    cudaMemcpy(he, err, sizeof(he), cudaMemcpyDeviceToHost);
    float w = 0.f;
    for (int i = 0; i < 256; ++i) w = he[i] > w ? he[i] : w;
    printf("polynomial max relative error on [-40, 8]: %.3g (2^%.1f); %s\n", w, log2f(w), cudaGetErrorString(cudaGetLastError()));
    return 0;
}

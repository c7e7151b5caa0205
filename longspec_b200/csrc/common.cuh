// Shared helpers for the longspec_b200 CUDA kernels (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/longspec_b200.h"

namespace ls {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return LS_ERR_CUDA;
    }
    return LS_OK;
}

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute is PER DEVICE, so the "already done"
// state is a bit per device ordinal (a process may drive several GPUs); a failing call is reported, not ignored.
template <typename K>
inline int ensure_dyn_smem(K kernel, int bytes, unsigned long long* done_mask, const char* what) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("%s: cudaGetDevice failed", what); return LS_ERR_CUDA; }
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(done_mask, __ATOMIC_ACQUIRE) & bit) return LS_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        set_error("%s: cudaFuncSetAttribute(%d bytes of dynamic shared memory) failed: %s", what, bytes, cudaGetErrorString(e));
        return LS_ERR_CUDA;
    }
    __atomic_fetch_or(done_mask, bit, __ATOMIC_RELEASE);
    return LS_OK;
}

#define LS_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            ls::set_error(__VA_ARGS__);       \
            return LS_ERR_INVALID_ARG;        \
        }                                     \
    } while (0)

template <typename T>
struct Num;
template <>
struct Num<__half> {
    static __device__ __forceinline__ float to_f(__half x) { return __half2float(x); }
    static __device__ __forceinline__ __half from_f(float x) { return __float2half_rn(x); }
    static __device__ __forceinline__ uint32_t pack(float a, float b) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    static __device__ __forceinline__ float2 unpack(uint32_t u) {
        return __half22float2(*reinterpret_cast<__half2*>(&u));
    }
};
template <>
struct Num<__nv_bfloat16> {
    static __device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
    static __device__ __forceinline__ __nv_bfloat16 from_f(float x) { return __float2bfloat16_rn(x); }
    static __device__ __forceinline__ uint32_t pack(float a, float b) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    static __device__ __forceinline__ float2 unpack(uint32_t u) {
        return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
    }
};

// ex2.approx: exact 2^x for the softmax (scores are kept in the log2 domain); -inf -> 0
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Number of full, unmasked 128-key tiles of the LOCAL prefix covered by the tensor-core kernel (attn_tc.cu);
// the masked remainder on the mma.sync path starts at this tile.  Shared so both kernels agree.
__device__ __forceinline__ int tc_full_tiles(const ls_attn_desc& d, int kv_len, int n_q) {
    const int se = d.kv_shard_end < 0 ? 0x7fffffff : d.kv_shard_end;
    const int end = min(kv_len - (d.causal ? n_q : 0), se) - d.kv_shard_begin;
    return max(end, 0) / 128;
}

}  // namespace ls

// Small fused element-wise kernels of the decoder layer: residual-add + RMSNorm, SiLU*up, embedding
// gather, row gather.  Reference: HF LlamaRMSNorm / LlamaMLP as used by llama.py:35-42,455-502.
#include "common.cuh"

namespace ls {

template <typename T>
__global__ void __launch_bounds__(256) add_rmsnorm_kernel(T* h, const T* delta, const T* w, T* out, int hidden, float eps) {
    const int r = blockIdx.x, tid = threadIdx.x;
    T* hr = h + static_cast<size_t>(r) * hidden;
    const T* dr = delta ? delta + static_cast<size_t>(r) * hidden : nullptr;
    extern __shared__ float xs[];                     // hidden floats
    float ss = 0.f;
    for (int c = tid * 8; c < hidden; c += blockDim.x * 8) {
        uint4 hv = *reinterpret_cast<const uint4*>(hr + c);
        uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
        if (dr) {
            const uint4 dv = *reinterpret_cast<const uint4*>(dr + c);
            const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 a = Num<T>::unpack(hw[e]), b = Num<T>::unpack(dw[e]);
                hw[e] = Num<T>::pack(a.x + b.x, a.y + b.y);          // residual add rounded to dtype
            }
            *reinterpret_cast<uint4*>(hr + c) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 a = Num<T>::unpack(hw[e]);
            xs[c + 2 * e] = a.x;
            xs[c + 2 * e + 1] = a.y;
            ss += a.x * a.x + a.y * a.y;
        }
    }
    __shared__ float red[8];
    __shared__ float bc;
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
        bc = rsqrtf(t / hidden + eps);
    }
    __syncthreads();
    const float rs = bc;
    T* orow = out + static_cast<size_t>(r) * hidden;
    for (int c = tid * 8; c < hidden; c += blockDim.x * 8) {
        const uint4 wv = *reinterpret_cast<const uint4*>(w + c);
        const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
        uint32_t ow[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 g = Num<T>::unpack(ww[e]);
            // weight * x_normed.to(dtype)
            const float2 y = Num<T>::unpack(Num<T>::pack(xs[c + 2 * e] * rs, xs[c + 2 * e + 1] * rs));
            ow[e] = Num<T>::pack(g.x * y.x, g.y * y.y);
        }
        *reinterpret_cast<uint4*>(orow + c) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

template <typename T>
__global__ void silu_mul_kernel(const T* gu, T* out, int rows, int ffn) {
    const size_t n8 = static_cast<size_t>(rows) * ffn / 8;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = (i * 8) / ffn, c = (i * 8) % ffn;
        const uint4 gv = *reinterpret_cast<const uint4*>(gu + r * 2 * ffn + c);
        const uint4 uv = *reinterpret_cast<const uint4*>(gu + r * 2 * ffn + ffn + c);
        const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w}, uw[4] = {uv.x, uv.y, uv.z, uv.w};
        uint32_t ow[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 g = Num<T>::unpack(gw[e]), u = Num<T>::unpack(uw[e]);
            // act_fn(gate) rounded to dtype, then * up
            const float2 a = Num<T>::unpack(Num<T>::pack(g.x / (1.f + __expf(-g.x)), g.y / (1.f + __expf(-g.y))));
            ow[e] = Num<T>::pack(a.x * u.x, a.y * u.y);
        }
        *reinterpret_cast<uint4*>(out + r * ffn + c) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

template <typename T, typename I>
__global__ void gather_rows_kernel(const T* table, const I* idx, T* out, int hidden) {
    const int r = blockIdx.x;
    const size_t src = static_cast<size_t>(idx[r]) * hidden;
    for (int c = threadIdx.x * 8; c < hidden; c += blockDim.x * 8)
        *reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * hidden + c) = *reinterpret_cast<const uint4*>(table + src + c);
}

}  // namespace ls

extern "C" int ls_add_rmsnorm(void* h, const void* delta, const void* w, void* out, int32_t rows, int32_t hidden,
                              float eps, int32_t dtype, void* stream) {
    using namespace ls;
    LS_REQUIRE(h && w && out && rows >= 1 && hidden >= 8 && hidden % 8 == 0, "ls_add_rmsnorm: bad arguments");
    LS_REQUIRE(hidden * 4 <= 48 * 1024, "ls_add_rmsnorm: hidden %d too large", hidden);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == LS_F16)
        add_rmsnorm_kernel<__half><<<rows, 256, hidden * 4, st>>>(static_cast<__half*>(h), static_cast<const __half*>(delta),
                                                                  static_cast<const __half*>(w), static_cast<__half*>(out), hidden, eps);
    else if (dtype == LS_BF16)
        add_rmsnorm_kernel<__nv_bfloat16><<<rows, 256, hidden * 4, st>>>(
            static_cast<__nv_bfloat16*>(h), static_cast<const __nv_bfloat16*>(delta), static_cast<const __nv_bfloat16*>(w),
            static_cast<__nv_bfloat16*>(out), hidden, eps);
    else {
        set_error("ls_add_rmsnorm: bad dtype %d", dtype);
        return LS_ERR_INVALID_ARG;
    }
    return check_launch("ls_add_rmsnorm");
}

extern "C" int ls_silu_mul(const void* gu, void* out, int32_t rows, int32_t ffn, int32_t dtype, void* stream) {
    using namespace ls;
    LS_REQUIRE(gu && out && rows >= 1 && ffn >= 8 && ffn % 8 == 0, "ls_silu_mul: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t n8 = static_cast<size_t>(rows) * ffn / 8;
    const int blocks = static_cast<int>(n8 / 256 + 1 < 148 * 8 ? n8 / 256 + 1 : 148 * 8);
    if (dtype == LS_F16)
        silu_mul_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<const __half*>(gu), static_cast<__half*>(out), rows, ffn);
    else if (dtype == LS_BF16)
        silu_mul_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(gu),
                                                               static_cast<__nv_bfloat16*>(out), rows, ffn);
    else {
        set_error("ls_silu_mul: bad dtype %d", dtype);
        return LS_ERR_INVALID_ARG;
    }
    return check_launch("ls_silu_mul");
}

extern "C" int ls_embed(const void* table, const int64_t* ids, void* out, int32_t rows, int32_t hidden, int32_t dtype,
                        void* stream) {
    using namespace ls;
    LS_REQUIRE(table && ids && out && rows >= 1 && hidden % 8 == 0, "ls_embed: bad arguments");
    LS_REQUIRE(dtype == LS_F16 || dtype == LS_BF16, "ls_embed: bad dtype %d", dtype);
    gather_rows_kernel<uint16_t, int64_t><<<rows, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint16_t*>(table), ids, static_cast<uint16_t*>(out), hidden);
    return check_launch("ls_embed");
}

extern "C" int ls_gather_rows(const void* x, const int32_t* rows_idx, void* out, int32_t rows, int32_t hidden,
                              int32_t dtype, void* stream) {
    using namespace ls;
    LS_REQUIRE(x && rows_idx && out && rows >= 1 && hidden % 8 == 0, "ls_gather_rows: bad arguments");
    LS_REQUIRE(dtype == LS_F16 || dtype == LS_BF16, "ls_gather_rows: bad dtype %d", dtype);
    gather_rows_kernel<uint16_t, int32_t><<<rows, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint16_t*>(x), rows_idx, static_cast<uint16_t*>(out), hidden);
    return check_launch("ls_gather_rows");
}

// RoPE (HF rotate-half convention) fused with the paged KV append (ls_rope_append).
// Reference: apply_rotary_pos_emb(q, k, cos, sin, unsqueeze_dim=2) (llama.py:378; HF modeling_llama)
// + K_Cache[b, cache_lens + r] = k (llama.py:396-399, llama_glide.py:312-315).
#include <algorithm>

#include "common.cuh"

namespace ls {

// One row per blockIdx.x, its head slots shared out over gridDim.y blocks.  A work item = (head slot, two adjacent rotary
// pairs): 4-byte loads / stores; cos / sin of the row are computed once per block (64 values) instead of once per thread.
// The rotated K goes to its in-place row AND straight to the paged pool (no second pass), V is copied in 16-byte pieces.
template <typename T>
__global__ void __launch_bounds__(256) rope_append_kernel(T* q, T* k, const T* v, int64_t q_stride, int64_t k_stride,
                                                          int64_t v_stride, const int32_t* positions,
                                                          const float* inv_freq, T* k_pool, T* v_pool,
                                                          const int32_t* page_table, const int32_t* base_len_ptr,
                                                          int32_t base_bias, int32_t slot_offset,
                                                          const int32_t* n_rows_ptr, int32_t n_heads,
                                                          int32_t n_kv_heads) {
    const int r = blockIdx.x;
    const int tid = threadIdx.x;
    __shared__ float cs[64], sn[64];
    if (tid < 64) {
        const float ang = static_cast<float>(positions[r]) * inv_freq[tid];
        // HF computes cos/sin in fp32 and casts the tables to the activation dtype
        cs[tid] = Num<T>::to_f(Num<T>::from_f(cosf(ang)));
        sn[tid] = Num<T>::to_f(Num<T>::from_f(sinf(ang)));
    }
    const bool append = k != nullptr && k_pool != nullptr && r < *n_rows_ptr;
    size_t pool_row = 0;                          // element offset of (page, head 0, slot) in the pools
    size_t head_step = 0;
    if (append) {
        const int tok = *base_len_ptr + base_bias + slot_offset + r;
        const int page = page_table[tok >> 7], slot = tok & (LS_PAGE - 1);
        pool_row = (static_cast<size_t>(page) * n_kv_heads * LS_PAGE + slot) * LS_HEAD_DIM;
        head_step = static_cast<size_t>(LS_PAGE) * LS_HEAD_DIM;
    }
    __syncthreads();
    const int n_rot = n_heads + (k != nullptr ? n_kv_heads : 0);
    const int step = gridDim.y * blockDim.x;
    for (int w = blockIdx.y * blockDim.x + tid; w < n_rot * 32; w += step) {
        const int hs = w >> 5, j2 = (w & 31) * 2;
        const bool is_k = hs >= n_heads;
        T* x = is_k ? k + r * k_stride + (hs - n_heads) * LS_HEAD_DIM : q + r * q_stride + hs * LS_HEAD_DIM;
        T lo[2], hi[2];
        *reinterpret_cast<uint32_t*>(lo) = *reinterpret_cast<const uint32_t*>(x + j2);
        *reinterpret_cast<uint32_t*>(hi) = *reinterpret_cast<const uint32_t*>(x + j2 + 64);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float a = Num<T>::to_f(lo[u]), b = Num<T>::to_f(hi[u]);
            const float c = cs[j2 + u], s = sn[j2 + u];
            // (x * cos) + (rotate_half(x) * sin), every op rounded to dtype like eager PyTorch
            lo[u] = Num<T>::from_f(Num<T>::to_f(Num<T>::from_f(a * c)) + Num<T>::to_f(Num<T>::from_f(-b * s)));
            hi[u] = Num<T>::from_f(Num<T>::to_f(Num<T>::from_f(b * c)) + Num<T>::to_f(Num<T>::from_f(a * s)));
        }
        *reinterpret_cast<uint32_t*>(x + j2) = *reinterpret_cast<const uint32_t*>(lo);
        *reinterpret_cast<uint32_t*>(x + j2 + 64) = *reinterpret_cast<const uint32_t*>(hi);
        if (is_k && append) {
            T* dst = k_pool + pool_row + (hs - n_heads) * head_step;
            *reinterpret_cast<uint32_t*>(dst + j2) = *reinterpret_cast<const uint32_t*>(lo);
            *reinterpret_cast<uint32_t*>(dst + j2 + 64) = *reinterpret_cast<const uint32_t*>(hi);
        }
    }
    if (append) {
        for (int idx = blockIdx.y * blockDim.x + tid; idx < n_kv_heads * 16; idx += step) {
            const int h = idx >> 4, chunk = idx & 15;
            *reinterpret_cast<uint4*>(v_pool + pool_row + h * head_step + chunk * 8) =
                *reinterpret_cast<const uint4*>(v + r * v_stride + h * LS_HEAD_DIM + chunk * 8);
        }
    }
}

}  // namespace ls

extern "C" int ls_rope_append(void* q, void* k, const void* v, int64_t q_stride, int64_t k_stride, int64_t v_stride,
                              const int32_t* positions, const float* inv_freq, void* k_pool, void* v_pool,
                              const int32_t* page_table, const int32_t* base_len_ptr, int32_t base_bias,
                              int32_t slot_offset, const int32_t* n_rows_ptr, int32_t rows, int32_t n_heads,
                              int32_t n_kv_heads, int32_t dtype, void* stream) {
    using namespace ls;
    LS_REQUIRE(q && positions && inv_freq, "ls_rope_append: null q/positions/inv_freq");
    LS_REQUIRE(rows >= 1 && n_heads >= 1, "ls_rope_append: bad shape");
    LS_REQUIRE(k_pool == nullptr || (k && v && v_pool && page_table && base_len_ptr && n_rows_ptr),
               "ls_rope_append: append requested with null k/v/pool/page_table/len pointers");
    LS_REQUIRE(q_stride % 8 == 0 && k_stride % 8 == 0 && v_stride % 8 == 0, "ls_rope_append: strides must be 16B multiples");
    LS_REQUIRE(dtype == LS_F16 || dtype == LS_BF16, "ls_rope_append: bad dtype %d", dtype);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // (rows, y): y blocks share a row's head slots; enough blocks to cover the device at the decode shapes (<= 80 rows)
    const int items = (n_heads + (k ? n_kv_heads : 0)) * 32;
    const dim3 grid(rows, std::max(1, std::min((items + 255) / 256, rows >= 256 ? 1 : 4)));
    if (dtype == LS_F16)
        rope_append_kernel<__half><<<grid, 256, 0, st>>>(
            static_cast<__half*>(q), static_cast<__half*>(k), static_cast<const __half*>(v), q_stride, k_stride, v_stride,
            positions, inv_freq, static_cast<__half*>(k_pool), static_cast<__half*>(v_pool), page_table, base_len_ptr,
            base_bias, slot_offset, n_rows_ptr, n_heads, n_kv_heads);
    else
        rope_append_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
            static_cast<__nv_bfloat16*>(q), static_cast<__nv_bfloat16*>(k), static_cast<const __nv_bfloat16*>(v), q_stride,
            k_stride, v_stride, positions, inv_freq, static_cast<__nv_bfloat16*>(k_pool),
            static_cast<__nv_bfloat16*>(v_pool), page_table, base_len_ptr, base_bias, slot_offset, n_rows_ptr, n_heads,
            n_kv_heads);
    return check_launch("ls_rope_append");
}

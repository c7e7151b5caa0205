// RoPE (HF rotate-half convention) fused with the paged KV append (ls_rope_append).
// Reference: apply_rotary_pos_emb(q, k, cos, sin, unsqueeze_dim=2) (llama.py:378; HF modeling_llama)
// + K_Cache[b, cache_lens + r] = k (llama.py:396-399, llama_glide.py:312-315).
#include "common.cuh"

namespace ls {

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
    const int j = tid & 63;                       // rotary pair index; constant per thread
    const float ang = static_cast<float>(positions[r]) * inv_freq[j];
    // HF computes cos/sin in fp32 and casts the tables to the activation dtype
    const float c = Num<T>::to_f(Num<T>::from_f(cosf(ang)));
    const float s = Num<T>::to_f(Num<T>::from_f(sinf(ang)));
    auto rot = [&](T* x) {
        const float a = Num<T>::to_f(x[j]), b = Num<T>::to_f(x[j + 64]);
        // (x * cos) + (rotate_half(x) * sin), every op rounded to dtype like eager PyTorch
        const float lo = Num<T>::to_f(Num<T>::from_f(Num<T>::to_f(Num<T>::from_f(a * c)) +
                                                     Num<T>::to_f(Num<T>::from_f(-b * s))));
        const float hi = Num<T>::to_f(Num<T>::from_f(Num<T>::to_f(Num<T>::from_f(b * c)) +
                                                     Num<T>::to_f(Num<T>::from_f(a * s))));
        x[j] = Num<T>::from_f(lo);
        x[j + 64] = Num<T>::from_f(hi);
    };
    for (int h = tid >> 6; h < n_heads; h += blockDim.x >> 6) rot(q + r * q_stride + h * LS_HEAD_DIM);
    if (k == nullptr) return;
    for (int h = tid >> 6; h < n_kv_heads; h += blockDim.x >> 6) rot(k + r * k_stride + h * LS_HEAD_DIM);
    if (k_pool == nullptr) return;
    if (r >= *n_rows_ptr) return;
    __syncthreads();                              // rotated k visible to the copy below
    const int tok = *base_len_ptr + base_bias + slot_offset + r;
    const int page = page_table[tok >> 7], slot = tok & (LS_PAGE - 1);
    for (int idx = tid; idx < n_kv_heads * 16; idx += blockDim.x) {
        const int h = idx >> 4, chunk = idx & 15;
        const size_t dst = ((static_cast<size_t>(page) * n_kv_heads + h) * LS_PAGE + slot) * LS_HEAD_DIM + chunk * 8;
        *reinterpret_cast<uint4*>(k_pool + dst) = *reinterpret_cast<const uint4*>(k + r * k_stride + h * LS_HEAD_DIM + chunk * 8);
        *reinterpret_cast<uint4*>(v_pool + dst) = *reinterpret_cast<const uint4*>(v + r * v_stride + h * LS_HEAD_DIM + chunk * 8);
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
    if (dtype == LS_F16)
        rope_append_kernel<__half><<<rows, 256, 0, st>>>(
            static_cast<__half*>(q), static_cast<__half*>(k), static_cast<const __half*>(v), q_stride, k_stride, v_stride,
            positions, inv_freq, static_cast<__half*>(k_pool), static_cast<__half*>(v_pool), page_table, base_len_ptr,
            base_bias, slot_offset, n_rows_ptr, n_heads, n_kv_heads);
    else
        rope_append_kernel<__nv_bfloat16><<<rows, 256, 0, st>>>(
            static_cast<__nv_bfloat16*>(q), static_cast<__nv_bfloat16*>(k), static_cast<const __nv_bfloat16*>(v), q_stride,
            k_stride, v_stride, positions, inv_freq, static_cast<__nv_bfloat16*>(k_pool),
            static_cast<__nv_bfloat16*>(v_pool), page_table, base_len_ptr, base_bias, slot_offset, n_rows_ptr, n_heads,
            n_kv_heads);
    return check_launch("ls_rope_append");
}

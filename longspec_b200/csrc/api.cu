// Library-level entry points: version, error string, device query, kernel dispatch.
#include <stdarg.h>

#include "common.cuh"

namespace ls {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace ls

extern "C" int ls_attn_partial_mma(const ls_attn_desc* d, void* stream);

extern "C" int ls_version(void) { return 100; }
extern "C" const char* ls_last_error(void) { return ls::g_err; }
extern "C" int ls_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return LS_ERR_CUDA;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return LS_ERR_CUDA;
    return n;
}

extern "C" int ls_attn_partial(const ls_attn_desc* d, void* stream) { return ls_attn_partial_mma(d, stream); }

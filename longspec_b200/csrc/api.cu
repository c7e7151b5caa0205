// Library-level entry points: version, error string, device query, kernel dispatch.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

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

extern "C" int ls_attn_partial_mma_remainder(const ls_attn_desc* d, void* stream);
extern "C" int ls_attn_partial_tc(const ls_attn_desc* d, int n_splits_tc, void* stream);
extern "C" int ls_attn_partial_tc_fused(const ls_attn_desc* d, void* stream);

// Dispatch: unwindowed calls with <= 384 packed rows run on the tcgen05/TMA kernel: n_splits - 1 CTAs per kv head
// stream the long unmasked prefix and one more CTA handles the masked remainder (partial tile, causal rows,
// tree-masked tail) in the same launch; everything else (the draft's 512-token window) stays on the general
// mma.sync kernel.  LONGSPEC_ATTN=mma forces the general path, LONGSPEC_ATTN=tc+mma the unfused pair.
extern "C" int ls_attn_tc_fuses_combine(const ls_attn_desc* d);

static int attn_mode() {
    static const int mode = []() {
        const char* e = getenv("LONGSPEC_ATTN");
        if (e && strcmp(e, "mma") == 0) return 1;
        if (e && strcmp(e, "tc+mma") == 0) return 2;
        return 0;
    }();
    return mode;
}

extern "C" int ls_attn_fuses_combine(const ls_attn_desc* d) {
    if (!d || d->n_kv_heads < 1) return LS_ERR_INVALID_ARG;
    const int G = d->n_heads / d->n_kv_heads;
    const bool tc = attn_mode() == 0 && d->window < 0 && d->q_rows * G <= 384 && d->n_pages > 0;
    return tc ? ls_attn_tc_fuses_combine(d) : 0;
}

extern "C" int ls_attn_partial(const ls_attn_desc* d, void* stream) {
    const int mode = attn_mode();
    const int force_mma = mode == 1;
    if (!d || d->n_kv_heads < 1) return ls_attn_partial_mma(d, stream);
    const int G = d->n_heads / d->n_kv_heads;
    const bool tc_ok = !force_mma && d->window < 0 && d->q_rows * G <= 384 && d->n_pages > 0 && (mode == 0 || d->n_splits >= 2);
    if (!tc_ok) return ls_attn_partial_mma(d, stream);
    if (mode == 0) return ls_attn_partial_tc_fused(d, stream);
    int rc = ls_attn_partial_tc(d, d->n_splits - 1, stream);
    if (rc) return rc;
    ls_attn_desc r = *d;
    r.n_splits = 1;
    r.part_offset = d->part_offset + d->n_splits - 1;
    return ls_attn_partial_mma_remainder(&r, stream);
}

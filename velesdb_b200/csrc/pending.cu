// pending.cu -- entry points declared in include/veles_b200.h whose kernels are not written yet.
// They fail loudly (VELES_ERR_UNSUPPORTED); nothing falls back to a CPU path.
#include "common.cuh"
using namespace veles;
extern "C" {
int32_t veles_bm25_from_csr(uint32_t, const uint64_t*, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t,
                            const uint32_t*, uint64_t, uint64_t, float, float, veles_bm25_t** out) {
    if (out) *out = nullptr;
    set_error("veles_bm25_from_csr: not implemented yet");
    return VELES_ERR_UNSUPPORTED;
}
int32_t veles_bm25_free(veles_bm25_t*) { return VELES_OK; }
int32_t veles_bm25_search_batch(const veles_bm25_t*, const uint32_t*, const uint32_t*, uint32_t, uint32_t, uint32_t*,
                                float*, uint32_t*, void*) {
    set_error("veles_bm25_search_batch: not implemented yet");
    return VELES_ERR_UNSUPPORTED;
}
int32_t veles_rrf_hybrid(const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t, uint32_t, float,
                         uint32_t, uint32_t*, float*, uint32_t*, void*) {
    set_error("veles_rrf_hybrid: not implemented yet");
    return VELES_ERR_UNSUPPORTED;
}
int32_t veles_fuse(int32_t, const uint32_t*, uint32_t, const uint32_t*, const float*, uint32_t, float, float, float,
                   uint32_t, uint32_t*, float*, uint32_t*, void*) {
    set_error("veles_fuse: not implemented yet");
    return VELES_ERR_UNSUPPORTED;
}
}

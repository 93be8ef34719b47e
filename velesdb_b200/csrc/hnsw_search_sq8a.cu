// hnsw_search_sq8a.cu -- kernel instantiations of hnsw_search.cuh for one storage type
#include "hnsw_search.cuh"

namespace veles {
SearchKernel search_kernel_sq8_a(uint32_t reg_mode, uint32_t qn, bool coop) {
    return qn == 2 ? VELES_PICK_KERNEL(VELES_SQ8, 2) : VELES_PICK_KERNEL(VELES_SQ8, 0);
}
}  // namespace veles

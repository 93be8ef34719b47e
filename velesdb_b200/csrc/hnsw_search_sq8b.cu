// hnsw_search_sq8b.cu -- kernel instantiations of hnsw_search.cuh for one storage type
#include "hnsw_search.cuh"

namespace veles {
SearchKernel search_kernel_sq8_b(uint32_t reg_mode, uint32_t qn, bool coop) {
    return qn == 4 ? VELES_PICK_KERNEL(VELES_SQ8, 4) : VELES_PICK_KERNEL(VELES_SQ8, 6);
}
}  // namespace veles

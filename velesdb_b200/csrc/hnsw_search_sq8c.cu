// hnsw_search_sq8c.cu -- kernel instantiations of hnsw_search.cuh for one storage type
#include "hnsw_search.cuh"

namespace veles {
SearchKernel search_kernel_sq8_c(uint32_t reg_mode, uint32_t, bool coop) { return VELES_PICK_KERNEL(VELES_SQ8, 8); }
}  // namespace veles

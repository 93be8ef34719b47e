// hnsw_search_bin1.cu -- kernel instantiations of hnsw_search.cuh for one storage type
#include "hnsw_search.cuh"

namespace veles {
SearchKernel search_kernel_bin1(uint32_t reg_mode) {
    const bool coop = false;  // packed rows use the direct-load evaluator, one warp per query only
    return VELES_PICK_KERNEL(VELES_BIN1, 0);
}
}  // namespace veles

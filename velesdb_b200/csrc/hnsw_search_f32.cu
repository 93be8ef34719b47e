// hnsw_search_f32.cu -- kernel instantiations of hnsw_search.cuh for one storage type
#include "hnsw_search.cuh"

namespace veles {
SearchKernel search_kernel_f32(uint32_t reg_mode, bool coop) { return VELES_PICK_KERNEL(VELES_F32, 0); }
}  // namespace veles

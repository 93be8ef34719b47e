// dev tool: ceiling of the search kernel's access pattern -- random 3088-byte rows fetched with 1-D
// bulk copies (UBLKCP) into a shared-memory ring, one warp per CTA, no dependent work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather_probe scripts/gather_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const uint8_t* vecs, uint64_t n, uint32_t row_bytes, uint32_t rows_per_warp, uint32_t stages,
                      uint32_t per_stage, float* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar = (uint64_t*)smem;
    uint8_t* ring = smem + 128;
    const uint32_t lane = threadIdx.x;
    if (lane == 0) {
        for (uint32_t i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint64_t rng = 0x9E3779B97F4A7C15ull * (blockIdx.x + 1);
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng % n; };
    const uint32_t nst = rows_per_warp / per_stage;
    auto issue = [&](uint32_t s) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[s])), "r"(per_stage * row_bytes) : "memory");
        for (uint32_t g = 0; g < per_stage; ++g) {
            const uint8_t* src = vecs + next() * row_bytes;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             s32(ring + (size_t)(s * per_stage + g) * row_bytes)), "l"(src), "r"(row_bytes), "r"(s32(&bar[s])) : "memory");
        }
    };
    if (lane == 0) for (uint32_t j = 0; j < stages && j < nst; ++j) issue(j);
    uint32_t phases = 0, s = 0;
    float acc = 0.f;
    for (uint32_t j = 0; j < nst; ++j) {
        uint32_t par = (phases >> s) & 1u;
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(s32(&bar[s])), "r"(par) : "memory");
        phases ^= 1u << s;
        acc += ((const float*)(ring + (size_t)s * per_stage * row_bytes))[lane];
        __syncwarp();
        if (lane == 0 && j + stages < nst) issue(s);
        s = (s + 1 == stages) ? 0 : s + 1;
    }
    if (acc == 123.456f) sink[0] = acc;
}
int main(int argc, char** argv) {
    const uint64_t n = 1000000;
    const uint32_t row = 3088;
    uint8_t* d; float* sink;
    cudaMalloc(&d, n * row); cudaMemset(d, 1, n * row); cudaMalloc(&sink, 4);
    printf("ctas/SM stages x rows : GB/s\n");
    for (int cfg = 0; cfg < 12; ++cfg) {
        uint32_t ctas[] = {7, 7, 5, 4, 8, 14, 3, 7, 4, 2, 1, 7}, stages[] = {2, 8, 3, 4, 7, 1, 5, 1, 1, 1, 1, 2}, per[] = {4, 1, 4, 4, 1, 4, 4, 4, 4, 4, 4, 2};
        uint32_t grid = ctas[cfg] * 148, rows = 3584;
        size_t smem = 128 + (size_t)stages[cfg] * per[cfg] * row;
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        probe<<<grid, 32, smem>>>(d, n, row, rows, stages[cfg], per[cfg], sink);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        probe<<<grid, 32, smem>>>(d, n, row, rows, stages[cfg], per[cfg], sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%u CTAs/SM, %u stages x %u rows (%zu B smem): %.1f GB/s (%.3f ms) err=%s\n", ctas[cfg], stages[cfg], per[cfg], smem,
               (double)grid * rows * row / ms / 1e6, ms, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

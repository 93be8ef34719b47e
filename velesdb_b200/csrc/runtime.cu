// runtime.cu -- process-level state of libveles_b200: device selection, errors, counters.
#include "common.cuh"

#include <map>
#include <mutex>
#include <tuple>

namespace veles {

static thread_local char g_err[1024] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cached_blocks_per_sm(const void* kernel, int threads, size_t smem) {
    static std::mutex mu;
    static std::map<std::tuple<const void*, int, size_t, int>, int> cache;
    static std::map<std::pair<const void*, int>, size_t> limit;  // the dynamic shared-memory limit set per kernel: only ever raised
    int dev = 0;
    cudaGetDevice(&dev);
    const auto key = std::make_tuple(kernel, threads, smem, dev);
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    cudaError_t e = cudaSuccess;
    size_t& lim = limit[std::make_pair(kernel, dev)];
    if (smem > lim) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) lim = smem;
    }
    int per_sm = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (e != cudaSuccess) {
        set_error("kernel configuration failed: %s", cudaGetErrorString(e));
        return -1;
    }
    cache[key] = per_sm;
    return per_sm;
}

}  // namespace veles

using namespace veles;

extern "C" {

int32_t veles_init(int32_t device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); libveles_b200 has no CPU path",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return VELES_ERR_CUDA;
    }
    VELES_REQUIRE(device >= 0 && device < count, "device %d out of range (0..%d)", device, count - 1);
    VELES_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VELES_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return VELES_ERR_UNSUPPORTED;
    }
    VELES_CUDA(cudaFree(0));
    return VELES_OK;
}

int32_t veles_shutdown(void) { return VELES_OK; }

const char* veles_last_error(void) { return g_err; }
const char* veles_version(void) { return "velesdb_b200 0.1.0 (sm_100a)"; }
uint64_t veles_launch_count(void) { return g_launches.load(); }

// SearchQuality::ef_search, index/hnsw/params.rs:309-319
uint64_t veles_ef_search(int32_t quality, uint64_t k, uint64_t custom_ef) {
    auto mx = [](uint64_t a, uint64_t b) { return a > b ? a : b; };
    switch (quality) {
        case VELES_FAST: return mx(64, k * 2);
        case VELES_BALANCED: return mx(128, k * 4);
        case VELES_ACCURATE: return mx(512, k * 16);
        case VELES_PERFECT: return mx(4096, k * 100);
        default: return mx(custom_ef, k);
    }
}

// NativeHnsw::transform_score, native/backend_adapter.rs:160-168
float veles_transform_score(int32_t metric, float raw) {
    switch (metric) {
        case VELES_COSINE: {
            float s = 1.0f - raw;
            if (s < 0.0f) s = 0.0f;
            if (s > 1.0f) s = 1.0f;
            return s;
        }
        case VELES_DOT: return -raw;
        default: return raw;
    }
}

}  // extern "C"

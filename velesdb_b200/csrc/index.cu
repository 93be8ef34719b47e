// index.cu -- device snapshot of a NativeHnsw: upload, layout, file format v1, export.
//
// Layout in HBM:
//   vecs       n rows x row_bytes.  Row = [dim elements (f32 | f16 | packed bits)] [pad to 4]
//              [sqrt(|v|^2) f32, cosine only] [pad to 16].  One 1-D bulk copy (TMA) fetches a row
//              together with its norm.
//   adj0       n rows x stride0 u32 (stride0 = max layer-0 degree rounded up to 32), padded with
//              VELES_INVALID_ID, neighbour order preserved (the traversal depends on it).
//   upper_ref  n u32: (first_row << 4) | top_layer for nodes with links above layer 0.
//   upper_adj  rows x strideU u32, layers 1..top_layer of each such node, consecutive.
#include "index.hpp"

#include <algorithm>
#include <cmath>
#include <memory>

namespace veles {

int device_sm_count() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

void compute_row_layout(veles_index* ix) {
    uint32_t payload;
    if (ix->dtype == VELES_BIN1) {
        payload = round_up(ix->dim, 64) / 8;
    } else {
        payload = ix->dim * elt_bytes(ix->dtype);
    }
    ix->norm_off = round_up(payload, 4);
    uint32_t total = ix->norm_off + ((ix->metric == VELES_COSINE && ix->dtype != VELES_BIN1) ? 4u : 0u);
    ix->row_bytes = round_up(std::max(total, 16u), 16);
}

// ---- kernels -------------------------------------------------------------------------------
__global__ void f32_to_f16_rows(const float* __restrict__ src, uint8_t* __restrict__ dst, uint64_t rows, uint32_t dim,
                                uint32_t row_bytes) {
    uint64_t total = rows * dim;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = i / dim;
        uint32_t c = (uint32_t)(i - r * dim);
        reinterpret_cast<__half*>(dst + r * row_bytes)[c] = __float2half_rn(src[i]);
    }
}

// bit j of a row = (x[j] > 0.5), LSB first (simd_explicit.rs:256-287 threshold rule)
__global__ void f32_to_bits_rows(const float* __restrict__ src, uint8_t* __restrict__ dst, uint64_t rows, uint32_t dim,
                                 uint32_t row_bytes) {
    uint32_t words = (dim + 31) / 32;
    uint64_t total = rows * words;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = i / words;
        uint32_t w = (uint32_t)(i - r * words);
        uint32_t bits = 0;
        for (uint32_t b = 0; b < 32; ++b) {
            uint32_t c = w * 32 + b;
            if (c < dim && src[r * dim + c] > 0.5f) bits |= 1u << b;
        }
        reinterpret_cast<uint32_t*>(dst + r * row_bytes)[w] = bits;
    }
}

// one warp per row: trailer = sqrt(tree_sum(v*v)), the |b| of simd_avx512.rs:271-352
template <typename T>
__global__ void row_norms_kernel(uint8_t* __restrict__ vecs, uint64_t n, uint32_t dim, uint32_t row_bytes,
                                 uint32_t norm_off) {
    uint32_t lane = threadIdx.x & 31;
    uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n; r += nwarps) {
        const T* row = reinterpret_cast<const T*>(vecs + r * row_bytes);
        float s = warp_tree_reduce<0>(row, row, dim, lane);
        if (lane == 0) *reinterpret_cast<float*>(vecs + r * row_bytes + norm_off) = __fsqrt_rn(s);
    }
}

int32_t upload_vectors(veles_index* ix, const void* vectors, int32_t src_dtype, cudaStream_t st) {
    compute_row_layout(ix);
    VELES_TRY(ix->vecs.alloc((size_t)ix->n * ix->row_bytes));
    if (ix->n == 0) return VELES_OK;
    VELES_CUDA(cudaMemsetAsync(ix->vecs.p, 0, (size_t)ix->n * ix->row_bytes, st));
    const int sms = device_sm_count();
    if (src_dtype == ix->dtype) {
        size_t width = ix->dtype == VELES_BIN1 ? (size_t)(ix->dim / 64) * 8 : (size_t)ix->dim * elt_bytes(ix->dtype);
        VELES_CUDA(cudaMemcpy2DAsync(ix->vecs.p, ix->row_bytes, vectors, width, width, ix->n, cudaMemcpyHostToDevice, st));
    } else if (src_dtype == VELES_F32) {
        // staged conversion: 64 MiB of f32 at a time
        uint64_t chunk = std::max<uint64_t>(1, (64ull << 20) / ((size_t)ix->dim * 4));
        DevBuf stage;
        VELES_TRY(stage.alloc((size_t)std::min<uint64_t>(chunk, ix->n) * ix->dim * 4));
        const float* src = static_cast<const float*>(vectors);
        for (uint64_t r0 = 0; r0 < ix->n; r0 += chunk) {
            uint64_t rows = std::min<uint64_t>(chunk, ix->n - r0);
            VELES_CUDA(cudaMemcpyAsync(stage.p, src + r0 * ix->dim, rows * ix->dim * 4, cudaMemcpyHostToDevice, st));
            uint8_t* dst = ix->vecs.as<uint8_t>() + r0 * ix->row_bytes;
            if (ix->dtype == VELES_F16) {
                f32_to_f16_rows<<<sms * 8, 256, 0, st>>>(stage.as<float>(), dst, rows, ix->dim, ix->row_bytes);
            } else {
                f32_to_bits_rows<<<sms * 8, 256, 0, st>>>(stage.as<float>(), dst, rows, ix->dim, ix->row_bytes);
            }
            count_launch();
            VELES_CUDA(cudaGetLastError());
            VELES_CUDA(cudaStreamSynchronize(st));
        }
    } else {
        set_error("unsupported conversion: src dtype %d -> store dtype %d", src_dtype, ix->dtype);
        return VELES_ERR_UNSUPPORTED;
    }
    if (ix->metric == VELES_COSINE && ix->dtype != VELES_BIN1) {
        if (ix->dtype == VELES_F32) {
            row_norms_kernel<float><<<sms * 4, 256, 0, st>>>(ix->vecs.as<uint8_t>(), ix->n, ix->dim, ix->row_bytes, ix->norm_off);
        } else {
            row_norms_kernel<__half><<<sms * 4, 256, 0, st>>>(ix->vecs.as<uint8_t>(), ix->n, ix->dim, ix->row_bytes, ix->norm_off);
        }
        count_launch();
        VELES_CUDA(cudaGetLastError());
    }
    VELES_CUDA(cudaStreamSynchronize(st));
    return VELES_OK;
}

// Builds the padded fixed-stride adjacency on the host and uploads it.  Duplicate ids inside one
// row are dropped keeping the first occurrence: the traversal would skip later copies anyway
// (visited.insert, graph.rs:499), so results are unchanged.
int32_t install_graph_host(veles_index* ix, uint32_t num_layers, const uint64_t* const* row_ptr,
                           const uint32_t* const* cols, const uint64_t* layer_nodes, uint32_t M, uint32_t M0,
                           uint64_t entry_point, uint32_t max_layer) {
    const uint64_t n = ix->n;
    VELES_REQUIRE(num_layers >= 1 && num_layers <= 16, "num_layers must be in 1..16, got %u", num_layers);
    VELES_REQUIRE(n < (1ull << 31), "at most 2^31-1 nodes per snapshot, got %llu", (unsigned long long)n);
    VELES_REQUIRE(n == 0 || entry_point < n, "entry point %llu out of range", (unsigned long long)entry_point);
    VELES_REQUIRE(max_layer < 16, "max_layer must be < 16");
    auto row_len = [&](uint32_t l, uint64_t node) -> uint64_t {
        if (node >= layer_nodes[l]) return 0;
        return row_ptr[l][node + 1] - row_ptr[l][node];
    };
    // strides
    uint64_t max0 = M0, maxU = M;
    for (uint64_t i = 0; i < std::min<uint64_t>(n, layer_nodes[0]); ++i) max0 = std::max(max0, row_len(0, i));
    for (uint32_t l = 1; l < num_layers; ++l)
        for (uint64_t i = 0; i < std::min<uint64_t>(n, layer_nodes[l]); ++i) maxU = std::max(maxU, row_len(l, i));
    VELES_REQUIRE(max0 <= 4096 && maxU <= 4096, "adjacency rows longer than 4096 are not supported");
    ix->stride0 = round_up((uint32_t)std::max<uint64_t>(max0, 1), 32);
    ix->strideU = round_up((uint32_t)std::max<uint64_t>(maxU, 1), 32);

    auto fill_row = [&](uint32_t* dst, uint32_t stride, uint32_t l, uint64_t node) -> bool {
        uint64_t len = row_len(l, node);
        uint32_t w = 0;
        const uint32_t* src = len ? cols[l] + row_ptr[l][node] : nullptr;
        for (uint64_t j = 0; j < len; ++j) {
            uint32_t x = src[j];
            if (x >= n) return false;
            bool dup = false;
            for (uint32_t t = 0; t < w; ++t)
                if (dst[t] == x) {
                    dup = true;
                    break;
                }
            if (!dup) dst[w++] = x;
        }
        for (; w < stride; ++w) dst[w] = VELES_INVALID_ID;
        return true;
    };

    std::vector<uint32_t> h_adj0((size_t)n * ix->stride0);
    for (uint64_t i = 0; i < n; ++i) {
        if (!fill_row(h_adj0.data() + i * ix->stride0, ix->stride0, 0, i)) {
            set_error("layer 0: node %llu has a neighbour id >= n", (unsigned long long)i);
            return VELES_ERR_INVALID;
        }
    }
    std::vector<uint32_t> h_ref(n, VELES_INVALID_ID);
    std::vector<uint32_t> h_up;
    uint64_t rows = 0;
    if (num_layers > 1) {
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t top = 0;
            for (uint32_t l = 1; l < num_layers; ++l)
                if (row_len(l, i) > 0) top = l;
            if (top == 0) continue;
            VELES_REQUIRE(rows < (1ull << 28), "too many upper-layer rows");
            h_ref[i] = (uint32_t)(rows << 4) | top;
            h_up.resize((rows + top) * ix->strideU);
            for (uint32_t l = 1; l <= top; ++l) {
                if (!fill_row(h_up.data() + (rows + l - 1) * ix->strideU, ix->strideU, l, i)) {
                    set_error("layer %u: node %llu has a neighbour id >= n", l, (unsigned long long)i);
                    return VELES_ERR_INVALID;
                }
            }
            rows += top;
        }
    }
    ix->upper_rows = rows;
    VELES_TRY(ix->adj0.alloc(h_adj0.size() * 4));
    VELES_TRY(ix->upper_ref.alloc(h_ref.size() * 4));
    VELES_TRY(ix->upper_adj.alloc(h_up.size() * 4));
    if (n) {
        VELES_CUDA(cudaMemcpy(ix->adj0.p, h_adj0.data(), h_adj0.size() * 4, cudaMemcpyHostToDevice));
        VELES_CUDA(cudaMemcpy(ix->upper_ref.p, h_ref.data(), h_ref.size() * 4, cudaMemcpyHostToDevice));
        if (!h_up.empty()) VELES_CUDA(cudaMemcpy(ix->upper_adj.p, h_up.data(), h_up.size() * 4, cudaMemcpyHostToDevice));
    }
    ix->M = M;
    ix->M0 = M0;
    ix->entry = entry_point;
    ix->max_layer = max_layer;
    ix->num_layers = num_layers;
    ix->has_entry = n > 0;
    ix->has_graph = true;
    return VELES_OK;
}

static int32_t check_common(uint64_t n, uint32_t dim, int32_t src_dtype, int32_t store_dtype, int32_t metric,
                            const void* vectors) {
    VELES_REQUIRE(metric >= VELES_COSINE && metric <= VELES_JACCARD, "unknown metric %d", metric);
    VELES_REQUIRE(store_dtype >= VELES_F32 && store_dtype <= VELES_BIN1, "unknown store dtype %d", store_dtype);
    VELES_REQUIRE(src_dtype >= VELES_F32 && src_dtype <= VELES_BIN1, "unknown source dtype %d", src_dtype);
    VELES_REQUIRE(dim > 0 && dim <= 65536, "dimension must be in 1..65536, got %u", dim);
    VELES_REQUIRE(n == 0 || vectors != nullptr, "vectors is NULL");
    if (store_dtype == VELES_BIN1) {
        VELES_REQUIRE(metric == VELES_HAMMING, "BIN1 storage supports the Hamming metric only");
        VELES_REQUIRE(dim % 64 == 0, "BIN1 storage needs dim %% 64 == 0, got %u", dim);
    }
    return VELES_OK;
}

}  // namespace veles

using namespace veles;

extern "C" {

int32_t veles_index_from_vectors(const void* vectors, uint64_t n, uint32_t dim, int32_t src_dtype, int32_t store_dtype,
                                 int32_t metric, veles_index_t** out) {
    VELES_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    VELES_TRY(check_common(n, dim, src_dtype, store_dtype, metric, vectors));
    VELES_REQUIRE(n < (1ull << 31), "at most 2^31-1 nodes per snapshot");
    std::unique_ptr<veles_index> ix(new veles_index());
    VELES_CUDA(cudaGetDevice(&ix->device));
    ix->metric = metric;
    ix->dtype = store_dtype;
    ix->dim = dim;
    ix->n = n;
    VELES_TRY(upload_vectors(ix.get(), vectors, src_dtype, nullptr));
    *out = ix.release();
    return VELES_OK;
}

// An empty snapshot of n zero rows, to be filled by veles_index_set_rows_d: for hosts that produce (or already
// hold) the vectors on the device, so that 10M+ row collections never pass through host memory.
int32_t veles_index_create(uint64_t n, uint32_t dim, int32_t store_dtype, int32_t metric, veles_index_t** out) {
    VELES_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    static const float dummy = 0.0f;
    VELES_TRY(check_common(n, dim, store_dtype, store_dtype, metric, &dummy));
    VELES_REQUIRE(n < (1ull << 31), "at most 2^31-1 nodes per snapshot");
    std::unique_ptr<veles_index> ix(new veles_index());
    VELES_CUDA(cudaGetDevice(&ix->device));
    ix->metric = metric;
    ix->dtype = store_dtype;
    ix->dim = dim;
    ix->n = n;
    compute_row_layout(ix.get());
    VELES_TRY(ix->vecs.alloc((size_t)n * ix->row_bytes));
    if (n) VELES_CUDA(cudaMemset(ix->vecs.p, 0, (size_t)n * ix->row_bytes));
    *out = ix.release();
    return VELES_OK;
}

// Rows [first, first + count) from device memory: src_dtype F32 (count*dim floats, converted to the store type as
// veles_index_from_vectors does) or the store type itself (F16: count*dim halves; BIN1: count*(dim/64) u64 words).
// Cosine norms of those rows are recomputed.  Any graph held by the snapshot is left alone.
int32_t veles_index_set_rows_d(veles_index_t* ix, uint64_t first, uint64_t count, const void* rows_d, int32_t src_dtype,
                               void* stream) {
    VELES_REQUIRE(ix != nullptr, "index is NULL");
    VELES_REQUIRE(first <= ix->n && count <= ix->n - first, "rows [%llu, +%llu) outside the snapshot (%llu rows)",
                  (unsigned long long)first, (unsigned long long)count, (unsigned long long)ix->n);
    if (count == 0) return VELES_OK;
    VELES_REQUIRE(rows_d != nullptr, "rows_d is NULL");
    VELES_REQUIRE(src_dtype == ix->dtype || src_dtype == VELES_F32, "unsupported conversion: src dtype %d -> store dtype %d",
                  src_dtype, ix->dtype);
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> g(ix->mu);
    const int sms = device_sm_count();
    uint8_t* dst = ix->vecs.as<uint8_t>() + first * ix->row_bytes;
    if (src_dtype == ix->dtype) {
        const size_t width = ix->dtype == VELES_BIN1 ? (size_t)(ix->dim / 64) * 8 : (size_t)ix->dim * elt_bytes(ix->dtype);
        VELES_CUDA(cudaMemcpy2DAsync(dst, ix->row_bytes, rows_d, width, width, count, cudaMemcpyDeviceToDevice, st));
    } else if (ix->dtype == VELES_F16) {
        f32_to_f16_rows<<<sms * 8, 256, 0, st>>>(static_cast<const float*>(rows_d), dst, count, ix->dim, ix->row_bytes);
        count_launch();
    } else {
        f32_to_bits_rows<<<sms * 8, 256, 0, st>>>(static_cast<const float*>(rows_d), dst, count, ix->dim, ix->row_bytes);
        count_launch();
    }
    VELES_CUDA(cudaGetLastError());
    if (ix->metric == VELES_COSINE && ix->dtype != VELES_BIN1) {
        if (ix->dtype == VELES_F32)
            row_norms_kernel<float><<<sms * 4, 256, 0, st>>>(dst, count, ix->dim, ix->row_bytes, ix->norm_off);
        else
            row_norms_kernel<__half><<<sms * 4, 256, 0, st>>>(dst, count, ix->dim, ix->row_bytes, ix->norm_off);
        count_launch();
        VELES_CUDA(cudaGetLastError());
    }
    return VELES_OK;
}

// Rows [first, first + count) back to the host in the store type (F32: floats, F16: halves, BIN1: u64 words), without
// the row trailer -- what a host needs to re-stage the vectors of a snapshot it loaded from files.
int32_t veles_index_get_rows(const veles_index_t* ix, uint64_t first, uint64_t count, void* out) {
    VELES_REQUIRE(ix != nullptr && (count == 0 || out != nullptr), "NULL argument");
    VELES_REQUIRE(first <= ix->n && count <= ix->n - first, "rows outside the snapshot");
    if (count == 0) return VELES_OK;
    const size_t width = ix->dtype == VELES_BIN1 ? (size_t)(ix->dim / 64) * 8 : (size_t)ix->dim * elt_bytes(ix->dtype);
    VELES_CUDA(cudaMemcpy2D(out, width, ix->vecs.as<uint8_t>() + first * ix->row_bytes, ix->row_bytes, width, count,
                            cudaMemcpyDeviceToHost));
    return VELES_OK;
}

int32_t veles_index_from_arrays(const void* vectors, uint64_t n, uint32_t dim, int32_t src_dtype, int32_t store_dtype,
                                int32_t metric, uint32_t num_layers, const uint64_t* const* row_ptr,
                                const uint32_t* const* cols, const uint64_t* layer_nodes, uint32_t M, uint32_t M0,
                                uint64_t entry_point, uint32_t max_layer, veles_index_t** out) {
    VELES_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    VELES_REQUIRE(row_ptr && cols && layer_nodes, "graph arrays are NULL");
    veles_index_t* ix = nullptr;
    VELES_TRY(veles_index_from_vectors(vectors, n, dim, src_dtype, store_dtype, metric, &ix));
    int32_t s = install_graph_host(ix, num_layers, row_ptr, cols, layer_nodes, M, M0, entry_point, max_layer);
    if (s != VELES_OK) {
        delete ix;
        return s;
    }
    *out = ix;
    return VELES_OK;
}

// NativeHnsw::file_load, native/backend_adapter.rs:274-380
int32_t veles_index_from_reference_files(const char* dir, const char* basename, int32_t metric, int32_t store_dtype,
                                         veles_index_t** out) {
    VELES_REQUIRE(out && dir && basename, "NULL argument");
    *out = nullptr;
    std::string vp = std::string(dir) + "/" + basename + ".vectors";
    std::string gp = std::string(dir) + "/" + basename + ".graph";
    auto io_fail = [&](const std::string& what) {
        set_error("%s", what.c_str());
        return (int32_t)VELES_ERR_IO;
    };
    FILE* f = std::fopen(vp.c_str(), "rb");
    if (!f) return io_fail("cannot open " + vp);
    uint32_t version = 0, dim = 0;
    uint64_t count = 0;
    bool ok = std::fread(&version, 4, 1, f) == 1 && std::fread(&count, 8, 1, f) == 1 && std::fread(&dim, 4, 1, f) == 1;
    if (!ok || version != 1) {
        std::fclose(f);
        return io_fail(ok ? "Unsupported version: " + std::to_string(version) : "truncated header in " + vp);
    }
    // sizes come from the file: validate them against the file itself before allocating anything
    long fsize = 0;
    {
        const long here = std::ftell(f);
        std::fseek(f, 0, SEEK_END);
        fsize = std::ftell(f);
        std::fseek(f, here, SEEK_SET);
    }
    if (dim > 65536 || count >= (1ull << 31) || (count > 0 && dim == 0) ||
        (uint64_t)count * dim > ((uint64_t)(fsize > 16 ? fsize - 16 : 0)) / 4) {
        std::fclose(f);
        return io_fail("corrupt header in " + vp + " (count/dimension do not fit the file)");
    }
    std::vector<float> vecs;
    try {
        vecs.resize((size_t)count * dim);
    } catch (const std::exception&) {
        std::fclose(f);
        set_error("out of host memory reading %s", vp.c_str());
        return VELES_ERR_OOM;
    }
    if (!vecs.empty() && std::fread(vecs.data(), 4, vecs.size(), f) != vecs.size()) {
        std::fclose(f);
        return io_fail("truncated vector data in " + vp);
    }
    std::fclose(f);
    f = std::fopen(gp.c_str(), "rb");
    if (!f) return io_fail("cannot open " + gp);
    uint32_t nl = 0, M = 0, M0 = 0, efc = 0, maxl = 0;
    uint64_t ep = 0, cnt2 = 0;
    ok = std::fread(&version, 4, 1, f) == 1 && std::fread(&nl, 4, 1, f) == 1 && std::fread(&M, 4, 1, f) == 1 &&
         std::fread(&M0, 4, 1, f) == 1 && std::fread(&efc, 4, 1, f) == 1 && std::fread(&ep, 8, 1, f) == 1 &&
         std::fread(&maxl, 4, 1, f) == 1 && std::fread(&cnt2, 8, 1, f) == 1;
    if (!ok || version != 1) {
        std::fclose(f);
        return io_fail(ok ? "Unsupported graph version: " + std::to_string(version) : "truncated header in " + gp);
    }
    if (nl == 0 || nl > 16) {
        std::fclose(f);
        return io_fail("bad layer count in " + gp);
    }
    std::vector<std::vector<uint64_t>> rps(nl);
    std::vector<std::vector<uint32_t>> cls(nl);
    std::vector<uint64_t> nodes(nl);
    std::string bad;
    try {
        for (uint32_t l = 0; l < nl && ok; ++l) {
            uint64_t nn = 0;
            if (std::fread(&nn, 8, 1, f) != 1) {
                ok = false;
                break;
            }
            if (nn > count) {  // Layer::new(num_nodes) never exceeds the vector count (graph.rs:173-178)
                bad = "layer " + std::to_string(l) + " claims more nodes than there are vectors in " + gp;
                break;
            }
            nodes[l] = nn;
            rps[l].resize(nn + 1);
            rps[l][0] = 0;
            for (uint64_t i = 0; i < nn; ++i) {
                uint32_t deg = 0;
                if (std::fread(&deg, 4, 1, f) != 1) {
                    ok = false;
                    break;
                }
                if (deg > 4096) {
                    bad = "adjacency row longer than 4096 in " + gp;
                    break;
                }
                size_t base = cls[l].size();
                cls[l].resize(base + deg);
                if (deg && std::fread(cls[l].data() + base, 4, deg, f) != deg) {
                    ok = false;
                    break;
                }
                rps[l][i + 1] = base + deg;
            }
            if (!bad.empty()) break;
        }
    } catch (const std::exception&) {
        std::fclose(f);
        set_error("out of host memory reading %s", gp.c_str());
        return VELES_ERR_OOM;
    }
    std::fclose(f);
    if (!bad.empty()) return io_fail(bad);
    if (!ok) return io_fail("truncated graph data in " + gp);
    std::vector<const uint64_t*> rp(nl);
    std::vector<const uint32_t*> cp(nl);
    static const uint32_t dummy = 0;
    for (uint32_t l = 0; l < nl; ++l) {
        rp[l] = rps[l].data();
        cp[l] = cls[l].empty() ? &dummy : cls[l].data();
    }
    if (count == 0) dim = dim ? dim : 1;
    int32_t s = veles_index_from_arrays(vecs.data(), count, dim, VELES_F32, store_dtype, metric, nl, rp.data(),
                                        cp.data(), nodes.data(), M, M0, ep, maxl, out);
    if (s == VELES_OK) (*out)->ef_construction = efc;
    return s;
}

int32_t veles_index_free(veles_index_t* idx) {
    delete idx;
    return VELES_OK;
}

uint64_t veles_index_len(const veles_index_t* idx) { return idx ? idx->n : 0; }
uint32_t veles_index_dim(const veles_index_t* idx) { return idx ? idx->dim : 0; }
int32_t veles_index_metric(const veles_index_t* idx) { return idx ? idx->metric : -1; }
uint32_t veles_index_max_layer(const veles_index_t* idx) { return idx ? idx->max_layer : 0; }
uint64_t veles_index_entry_point(const veles_index_t* idx) { return idx ? idx->entry : 0; }
uint64_t veles_index_device_bytes(const veles_index_t* idx) { return idx ? idx->device_bytes() : 0; }

int32_t veles_index_export_layer(const veles_index_t* idx, uint32_t layer, uint64_t* out_nodes, uint64_t* out_edges,
                                 uint64_t* row_ptr, uint32_t* cols) {
    VELES_REQUIRE(idx && out_nodes && out_edges, "NULL argument");
    VELES_REQUIRE(idx->has_graph, "snapshot has no graph");
    VELES_REQUIRE(layer < idx->num_layers, "layer %u out of range", layer);
    const uint64_t n = idx->n;
    *out_nodes = n;
    uint64_t edges = 0;
    if (layer == 0) {
        std::vector<uint32_t> h((size_t)n * idx->stride0);
        if (n) VELES_CUDA(cudaMemcpy(h.data(), idx->adj0.p, h.size() * 4, cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < n; ++i) {
            if (row_ptr) row_ptr[i] = edges;
            for (uint32_t j = 0; j < idx->stride0; ++j) {
                uint32_t x = h[i * idx->stride0 + j];
                if (x == VELES_INVALID_ID) break;
                if (cols) cols[edges] = x;
                ++edges;
            }
        }
    } else {
        std::vector<uint32_t> ref(n), up((size_t)idx->upper_rows * idx->strideU);
        if (n) VELES_CUDA(cudaMemcpy(ref.data(), idx->upper_ref.p, n * 4, cudaMemcpyDeviceToHost));
        if (!up.empty()) VELES_CUDA(cudaMemcpy(up.data(), idx->upper_adj.p, up.size() * 4, cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < n; ++i) {
            if (row_ptr) row_ptr[i] = edges;
            uint32_t r = ref[i];
            if (r == VELES_INVALID_ID || layer > (r & 15u)) continue;
            const uint32_t* row = up.data() + ((size_t)(r >> 4) + layer - 1) * idx->strideU;
            for (uint32_t j = 0; j < idx->strideU; ++j) {
                if (row[j] == VELES_INVALID_ID) break;
                if (cols) cols[edges] = row[j];
                ++edges;
            }
        }
    }
    if (row_ptr) row_ptr[n] = edges;
    *out_edges = edges;
    return VELES_OK;
}

// the `.graph` half of NativeHnsw::file_dump (native/backend_adapter.rs:213-261); any storage type
int32_t veles_index_dump_graph(const veles_index_t* idx, const char* dir, const char* basename) {
    VELES_REQUIRE(idx && dir && basename, "NULL argument");
    VELES_REQUIRE(idx->has_graph, "snapshot has no graph");
    const uint64_t n = idx->n;
    const std::string gp = std::string(dir) + "/" + basename + ".graph";
    FILE* f = std::fopen(gp.c_str(), "wb");
    if (!f) {
        set_error("cannot create %s", gp.c_str());
        return VELES_ERR_IO;
    }
    bool ok = true;
    auto put = [&](const void* p, size_t sz, size_t cnt) { ok = ok && std::fwrite(p, sz, cnt, f) == cnt; };
    const uint32_t version = 1, nl = idx->num_layers, maxl = idx->max_layer, efc = idx->ef_construction;
    const uint64_t ep = idx->has_entry ? idx->entry : 0;
    put(&version, 4, 1);
    put(&nl, 4, 1);
    put(&idx->M, 4, 1);
    put(&idx->M0, 4, 1);
    put(&efc, 4, 1);
    put(&ep, 8, 1);
    put(&maxl, 4, 1);
    put(&n, 8, 1);
    try {
        for (uint32_t l = 0; l < nl && ok; ++l) {
            uint64_t nodes = 0, edges = 0;
            int32_t s = veles_index_export_layer(idx, l, &nodes, &edges, nullptr, nullptr);
            std::vector<uint64_t> rp;
            std::vector<uint32_t> cl;
            if (s == VELES_OK) {
                rp.resize(nodes + 1);
                cl.resize(std::max<uint64_t>(edges, 1));
                s = veles_index_export_layer(idx, l, &nodes, &edges, rp.data(), cl.data());
            }
            if (s != VELES_OK) {
                std::fclose(f);
                return s;
            }
            put(&nodes, 8, 1);
            // one buffered record stream per layer: deg, ids, deg, ids, ...
            std::vector<uint32_t> rec;
            rec.reserve((size_t)1 << 20);
            for (uint64_t i = 0; i < nodes && ok; ++i) {
                const uint32_t deg = (uint32_t)(rp[i + 1] - rp[i]);
                rec.push_back(deg);
                rec.insert(rec.end(), cl.begin() + rp[i], cl.begin() + rp[i] + deg);
                if (rec.size() >= ((size_t)1 << 20) - 4200) {
                    put(rec.data(), 4, rec.size());
                    rec.clear();
                }
            }
            if (!rec.empty()) put(rec.data(), 4, rec.size());
        }
    } catch (const std::exception&) {
        std::fclose(f);
        set_error("out of host memory writing %s", gp.c_str());
        return VELES_ERR_OOM;
    }
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) {
        set_error("short write to %s (disk full?)", gp.c_str());
        return VELES_ERR_IO;
    }
    return VELES_OK;
}

// NativeHnsw::file_dump, native/backend_adapter.rs:184-261
int32_t veles_index_dump(const veles_index_t* idx, const char* dir, const char* basename) {
    VELES_REQUIRE(idx && dir && basename, "NULL argument");
    VELES_REQUIRE(idx->dtype == VELES_F32, "file format v1 stores f32 vectors; this snapshot holds dtype %d", idx->dtype);
    VELES_REQUIRE(idx->has_graph, "snapshot has no graph");
    const uint64_t n = idx->n;
    const std::string vp = std::string(dir) + "/" + basename + ".vectors";
    FILE* f = std::fopen(vp.c_str(), "wb");
    if (!f) {
        set_error("cannot create %s", vp.c_str());
        return VELES_ERR_IO;
    }
    bool ok = true;
    const uint32_t version = 1, dim = n ? idx->dim : 0;
    ok = ok && std::fwrite(&version, 4, 1, f) == 1 && std::fwrite(&n, 8, 1, f) == 1 && std::fwrite(&dim, 4, 1, f) == 1;
    try {
        // 64 MiB of rows at a time
        const uint64_t chunk = std::max<uint64_t>(1, ((uint64_t)64 << 20) / ((uint64_t)idx->dim * 4));
        std::vector<float> buf((size_t)std::min<uint64_t>(chunk, std::max<uint64_t>(n, 1)) * idx->dim);
        for (uint64_t r0 = 0; r0 < n && ok; r0 += chunk) {
            const uint64_t rows = std::min<uint64_t>(chunk, n - r0);
            cudaError_t e = cudaMemcpy2D(buf.data(), (size_t)idx->dim * 4, idx->vecs.as<uint8_t>() + r0 * idx->row_bytes,
                                         idx->row_bytes, (size_t)idx->dim * 4, rows, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) {
                std::fclose(f);
                set_error("cudaMemcpy2D failed: %s", cudaGetErrorString(e));
                return VELES_ERR_CUDA;
            }
            ok = ok && std::fwrite(buf.data(), 4, (size_t)rows * idx->dim, f) == (size_t)rows * idx->dim;
        }
    } catch (const std::exception&) {
        std::fclose(f);
        set_error("out of host memory writing %s", vp.c_str());
        return VELES_ERR_OOM;
    }
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) {
        set_error("short write to %s (disk full?)", vp.c_str());
        return VELES_ERR_IO;
    }
    return veles_index_dump_graph(idx, dir, basename);
}

}  // extern "C"

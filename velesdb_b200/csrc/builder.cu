// builder.cu -- bulk HNSW graph construction on the GPU.
//
// Role: HnswIndex::insert_batch_parallel (index/hnsw/index/batch.rs:82-108).  The reference builds
// incrementally (graph.rs:158-237) at ~ms per insert and its parallel path is order dependent
// (SURVEY finding 0.6), so there is no reference graph to be bit-equal to at 1M nodes.  This
// builder is the B200-shaped alternative: for every layer it computes the exact nearest
// neighbours of each node among that layer's nodes with a tensor-core GEMM (cuBLAS fp16 inputs,
// f32 accumulate -- a plain library GEMM), keeps the closest `max_conn` per node with a streaming
// top-k kernel, adds the reverse links and prunes closest-first -- the fixed point of the
// reference's add_bidirectional_connection (graph.rs:592-639: a full list is re-sorted by distance
// to its owner and cut to max_conn).
//
// What equals the reference exactly: the level of every node and the entry point (xorshift64 PRNG
// of graph.rs:368-403 drawn in node-id order), M0 = 2*M, degree bounds, file format.  What does
// not: neighbour sets (judged by recall@k against brute force, not by id equality).
#include <cublas_v2.h>

#include <algorithm>
#include <cmath>
#include <cub/device/device_radix_sort.cuh>
#include <limits>

#include "index.hpp"

namespace veles {

#define VELES_CUBLAS(expr)                                                                   \
    do {                                                                                     \
        cublasStatus_t _s = (expr);                                                          \
        if (_s != CUBLAS_STATUS_SUCCESS) {                                                   \
            set_error("%s failed with cuBLAS status %d (%s:%d)", #expr, (int)_s, __FILE__, __LINE__); \
            return VELES_ERR_CUDA;                                                           \
        }                                                                                    \
    } while (0)

// graph.rs:368-403
static void reference_levels(uint64_t n, uint32_t M, std::vector<uint8_t>& level) {
    level.resize(n);
    uint64_t s = 0x5DEECE66D1A4B5B5ull;
    const double level_mult = 1.0 / std::log((double)M);
    for (uint64_t i = 0; i < n; ++i) {
        if (s == 0) s = 0x853c49e6748fea9bull;
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        double u = (double)s / 18446744073709551616.0;
        u = std::max(u, std::numeric_limits<double>::min());
        double lv = std::floor(-std::log(u) * level_mult);
        level[i] = (uint8_t)std::min(15.0, std::max(0.0, lv));
    }
}

// fp16 working copy for the GEMM: row r of out = vector ids[r] (or r when ids == nullptr),
// L2-normalised for cosine (so that the GEMM yields the cosine directly), zero padded to dpad.
__global__ void make_f16_rows(IndexView ix, const uint32_t* __restrict__ ids, uint64_t rows, uint32_t dpad,
                              __half* __restrict__ out, float* __restrict__ sqnorm) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = gw; r < rows; r += nwarps) {
        const uint64_t id = ids ? ids[r] : r;
        const uint8_t* row = ix.vecs + id * ix.row_bytes;
        float scale = 1.0f;
        if (ix.metric == VELES_COSINE) {
            float nb = *reinterpret_cast<const float*>(row + ix.norm_off);
            scale = nb > 0.0f ? 1.0f / nb : 0.0f;
        }
        float ss = 0.0f;
        for (uint32_t i = lane; i < dpad; i += 32) {
            float x = 0.0f;
            if (i < ix.dim) {
                x = ix.dtype == VELES_F32 ? reinterpret_cast<const float*>(row)[i]
                                          : __half2float(reinterpret_cast<const __half*>(row)[i]);
                if (ix.metric == VELES_HAMMING || ix.metric == VELES_JACCARD) x = x > 0.5f ? 1.0f : 0.0f;
            }
            __half h = __float2half_rn(x * scale);
            out[r * dpad + i] = h;
            float hf = __half2float(h);
            ss += hf * hf;
        }
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
        if (lane == 0 && sqnorm) sqnorm[r] = ss;
    }
}

// approximate in-graph distance from a GEMM score (fp16 inputs): only the order matters
__device__ __forceinline__ float score_to_dist(int metric, float s, float na, float nb) {
    switch (metric) {
        case VELES_COSINE: return 1.0f - s;
        case VELES_DOT: return -s;
        case VELES_JACCARD: {
            float uni = na + nb - s;
            return uni > 0.0f ? 1.0f - s / uni : 0.0f;
        }
        default: {  // EUCLIDEAN, HAMMING: |a|^2 + |b|^2 - 2ab
            float d = na + nb - 2.0f * s;
            d = d > 0.0f ? d : 0.0f;
            return metric == VELES_EUCLIDEAN ? sqrtf(d) : d;
        }
    }
}

// One CTA (kSelWarps warps) per score row: the K columns with the smallest (distance, column), self
// excluded.  Each warp streams a quarter of the row (float4 loads, 512 columns per step) through a
// threshold filter into its own sorted list; warp 0 merges.  Emits both directions of every kept
// edge as (owner << 32 | order(dist)) -> neighbour.
constexpr int kSelWarps = 4;
__global__ void __launch_bounds__(kSelWarps * 32) select_edges_kernel(
    const float* __restrict__ scores, uint32_t ldc, uint32_t row0, uint32_t n_l, uint32_t K, int metric,
    const float* __restrict__ sqnorm, uint64_t* __restrict__ edge_key, uint32_t* __restrict__ edge_val) {
    extern __shared__ __align__(16) uint64_t sel_smem[];
    __shared__ uint32_t s_len[kSelWarps];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* res = sel_smem + (size_t)warp * K;
    const uint32_t r = blockIdx.x;
    const uint32_t self = row0 + r;
    const float* row = scores + (size_t)r * ldc;
    const float na = sqnorm ? sqnorm[self] : 0.0f;
    uint32_t seg = (n_l + kSelWarps - 1) / kSelWarps;
    seg = (seg + 511u) & ~511u;
    const uint32_t c_begin = warp * seg;
    const uint32_t c_end = min(n_l, c_begin + seg);
    uint32_t len = 0;
    uint64_t worst = ~0ull;
    for (uint32_t base = c_begin; base < c_end; base += 512) {
        uint64_t key[16];
        bool any = false;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t c0 = base + u * 128 + lane * 4;
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 < c_end) s4 = *reinterpret_cast<const float4*>(row + c0);  // ldc % 4 == 0: in-bounds, aligned
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t c = c0 + e;
                uint64_t kk = ~0ull;
                if (c < c_end && c != self) {
                    const float d = score_to_dist(metric, sv[e], na, sqnorm ? sqnorm[c] : 0.0f);
                    kk = ((uint64_t)ord_key(d) << 32) | c;
                    any |= kk < worst;
                }
                key[u * 4 + e] = kk;
            }
        }
        if (!__any_sync(FULL_MASK, any)) continue;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            uint32_t msk = __ballot_sync(FULL_MASK, key[u] < worst);
            while (msk) {
                const uint32_t src = __ffs(msk) - 1;
                msk &= msk - 1;
                const uint64_t kk = __shfl_sync(FULL_MASK, key[u], src);
                if (kk >= worst) continue;
                const uint32_t pos = lower_bound_warp(res, len, kk, lane);
                if (len < K) {
                    insert_at(res, pos, len + 1, kk, lane);
                    ++len;
                } else {
                    insert_at(res, pos, len, kk, lane);
                }
                if (len == K) worst = res[K - 1];
            }
        }
    }
    if (lane == 0) s_len[warp] = len;
    __syncthreads();
    if (warp != 0) return;
    for (uint32_t w = 1; w < kSelWarps; ++w) {
        const uint64_t* other = sel_smem + (size_t)w * K;
        const uint32_t olen = s_len[w];
        for (uint32_t j = 0; j < olen; ++j) {
            const uint64_t kk = other[j];
            if (len == K && kk >= worst) break;  // ascending: nothing later can enter
            const uint32_t pos = lower_bound_warp(res, len, kk, lane);
            if (len < K) {
                insert_at(res, pos, len + 1, kk, lane);
                ++len;
            } else {
                insert_at(res, pos, len, kk, lane);
            }
            if (len == K) worst = res[K - 1];
        }
    }
    __syncwarp();
    for (uint32_t j = lane; j < K; j += 32) {
        const size_t e = ((size_t)self * K + j) * 2;
        if (j < len) {
            const uint64_t kk = res[j];
            const uint32_t nbr = (uint32_t)kk;
            const uint64_t dk = kk >> 32;
            edge_key[e] = ((uint64_t)self << 32) | dk;
            edge_val[e] = nbr;
            edge_key[e + 1] = ((uint64_t)nbr << 32) | dk;
            edge_val[e + 1] = self;
        } else {
            edge_key[e] = ~0ull;
            edge_val[e] = VELES_INVALID_ID;
            edge_key[e + 1] = ~0ull;
            edge_val[e + 1] = VELES_INVALID_ID;
        }
    }
}

// one warp per owner: first maxc distinct neighbours of its (distance-sorted) edge segment
__global__ void finalize_rows_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                     uint64_t n_edges, uint32_t n_l, const uint32_t* __restrict__ ids,
                                     uint32_t maxc, uint32_t layer, uint32_t* __restrict__ adj0,
                                     uint32_t stride0, const uint32_t* __restrict__ upper_ref,
                                     uint32_t* __restrict__ upper_adj, uint32_t strideU) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t o = gw; o < n_l; o += nwarps) {
        uint64_t lo, hi;
        {
            const uint64_t target = o << 32;
            uint64_t a = 0, b = n_edges;
            while (a < b) {
                uint64_t m = (a + b) >> 1;
                if (keys[m] < target) a = m + 1; else b = m;
            }
            lo = a;
            const uint64_t target2 = (o + 1) << 32;
            b = n_edges;
            while (a < b) {
                uint64_t m = (a + b) >> 1;
                if (keys[m] < target2) a = m + 1; else b = m;
            }
            hi = a;
        }
        const uint32_t node = ids ? ids[o] : (uint32_t)o;
        uint32_t* out;
        uint32_t stride;
        if (layer == 0) {
            out = adj0 + (size_t)node * stride0;
            stride = stride0;
        } else {
            const uint32_t ref = upper_ref[node];
            out = upper_adj + ((size_t)(ref >> 4) + layer - 1) * strideU;
            stride = strideU;
        }
        uint32_t kept = 0;
        for (uint64_t base = lo; base < hi && kept < maxc; base += 32) {
            const uint64_t i = base + lane;
            const uint32_t nb_loc = i < hi ? vals[i] : VELES_INVALID_ID;
            const uint32_t nb = (nb_loc == VELES_INVALID_ID) ? VELES_INVALID_ID : (ids ? ids[nb_loc] : nb_loc);
            bool fresh = nb != VELES_INVALID_ID;
            if (fresh) {
                for (uint32_t t = 0; t < kept; ++t)
                    if (out[t] == nb) {
                        fresh = false;
                        break;
                    }
            }
            const uint32_t grp = __match_any_sync(FULL_MASK, nb);
            fresh = fresh && ((uint32_t)(__ffs(grp) - 1) == lane);
            const uint32_t msk = __ballot_sync(FULL_MASK, fresh);
            const uint32_t pos = kept + __popc(msk & ((1u << lane) - 1u));
            __syncwarp();
            if (fresh && pos < maxc) out[pos] = nb;
            kept = min(maxc, kept + (uint32_t)__popc(msk));
            __syncwarp();
        }
        for (uint32_t t = kept + lane; t < stride; t += 32) out[t] = VELES_INVALID_ID;
        __syncwarp();
    }
}

static int32_t build_layer(veles_index* ix, cublasHandle_t blas, uint32_t layer, const uint32_t* ids_d, uint32_t n_l,
                           uint32_t maxc, const __half* x16, const float* sqnorm, uint32_t dpad, DevBuf& scores,
                           cudaStream_t st) {
    const int sms = device_sm_count();
    const uint32_t K = std::min(maxc, n_l > 0 ? n_l - 1 : 0);
    uint32_t* adj0 = ix->adj0.as<uint32_t>();
    uint32_t* upper_adj = ix->upper_adj.as<uint32_t>();
    if (K == 0) {  // a single node on this layer: empty row (already INVALID-filled)
        return VELES_OK;
    }
    const uint64_t n_edges = (uint64_t)n_l * K * 2;
    DevBuf key_a, key_b, val_a, val_b, tmp;
    VELES_TRY(key_a.alloc(n_edges * 8));
    VELES_TRY(key_b.alloc(n_edges * 8));
    VELES_TRY(val_a.alloc(n_edges * 4));
    VELES_TRY(val_b.alloc(n_edges * 4));
    // row chunk: score matrix bounded to the scores buffer
    const uint32_t ldc = round_up(n_l, 4);
    const uint64_t max_rows = std::max<uint64_t>(1, scores.bytes / ((uint64_t)ldc * 4));
    const uint32_t chunk = (uint32_t)std::min<uint64_t>(max_rows, std::min<uint64_t>(n_l, 4096));
    const float alpha = 1.0f, beta = 0.0f;
    for (uint32_t r0 = 0; r0 < n_l; r0 += chunk) {
        const uint32_t rows = std::min(chunk, n_l - r0);
        // C^T [n_l x rows] (column major) = X^T[n_l x d] * A[d x rows]
        VELES_CUBLAS(cublasGemmEx(blas, CUBLAS_OP_T, CUBLAS_OP_N, (int)n_l, (int)rows, (int)dpad, &alpha, x16, CUDA_R_16F,
                                  (int)dpad, x16 + (size_t)r0 * dpad, CUDA_R_16F, (int)dpad, &beta, scores.p, CUDA_R_32F,
                                  (int)ldc, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT_TENSOR_OP));
        count_launch();
        select_edges_kernel<<<rows, kSelWarps * 32, (size_t)kSelWarps * K * 8, st>>>(
            scores.as<float>(), ldc, r0, n_l, K, ix->metric, sqnorm, key_a.as<uint64_t>(), val_a.as<uint32_t>());
        count_launch();
        VELES_CUDA(cudaGetLastError());
    }
    // sort all directed edges by (owner, distance)
    size_t tmp_bytes = 0;
    VELES_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_a.as<uint64_t>(), key_b.as<uint64_t>(),
                                               val_a.as<uint32_t>(), val_b.as<uint32_t>(), n_edges, 0, 64, st));
    VELES_TRY(tmp.alloc(tmp_bytes));
    VELES_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key_a.as<uint64_t>(), key_b.as<uint64_t>(),
                                               val_a.as<uint32_t>(), val_b.as<uint32_t>(), n_edges, 0, 64, st));
    count_launch(4);
    finalize_rows_kernel<<<sms * 8, 256, 0, st>>>(key_b.as<uint64_t>(), val_b.as<uint32_t>(), n_edges, n_l, ids_d,
                                                         maxc, layer, adj0, ix->stride0, ix->upper_ref.as<uint32_t>(),
                                                         upper_adj, ix->strideU);
    count_launch();
    VELES_CUDA(cudaGetLastError());
    VELES_CUDA(cudaStreamSynchronize(st));
    return VELES_OK;
}

}  // namespace veles

using namespace veles;

extern "C" int32_t veles_index_build_graph(veles_index_t* ix, uint32_t M, uint32_t cand_k, void* stream) {
    (void)cand_k;  // reserved: candidate pool for a diversity heuristic (builder)
    VELES_REQUIRE(ix != nullptr, "index is NULL");
    VELES_REQUIRE(M >= 2 && M <= 256, "M must be in 2..256, got %u", M);
    VELES_REQUIRE(ix->dtype != VELES_BIN1, "the bulk builder does not support packed-bit storage yet");
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> g(ix->mu);
    const uint64_t n = ix->n;
    const uint32_t M0 = 2 * M;
    ix->M = M;
    ix->M0 = M0;
    ix->stride0 = round_up(M0, 32);
    ix->strideU = round_up(M, 32);
    ix->ef_construction = 0;
    if (n == 0) {
        VELES_TRY(ix->adj0.alloc(16));
        VELES_TRY(ix->upper_ref.alloc(16));
        VELES_TRY(ix->upper_adj.alloc(16));
        ix->has_graph = true;
        ix->has_entry = false;
        ix->max_layer = 0;
        ix->num_layers = 1;
        ix->upper_rows = 0;
        return VELES_OK;
    }
    // levels, entry point (graph.rs:168, 230-233)
    std::vector<uint8_t> level;
    reference_levels(n, M, level);
    uint32_t max_layer = 0;
    uint64_t entry = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (level[i] > max_layer) {
            max_layer = level[i];
            entry = i;
        }
    std::vector<uint32_t> h_ref(n, VELES_INVALID_ID);
    uint64_t rows = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (level[i] > 0) {
            VELES_REQUIRE(rows < (1ull << 28), "too many upper-layer rows");
            h_ref[i] = (uint32_t)(rows << 4) | level[i];
            rows += level[i];
        }
    ix->upper_rows = rows;
    VELES_TRY(ix->adj0.alloc((size_t)n * ix->stride0 * 4));
    VELES_TRY(ix->upper_ref.alloc((size_t)n * 4));
    VELES_TRY(ix->upper_adj.alloc(std::max<size_t>((size_t)rows * ix->strideU * 4, 16)));
    VELES_CUDA(cudaMemsetAsync(ix->adj0.p, 0xff, ix->adj0.bytes, st));
    VELES_CUDA(cudaMemsetAsync(ix->upper_adj.p, 0xff, ix->upper_adj.bytes, st));
    VELES_CUDA(cudaMemcpyAsync(ix->upper_ref.p, h_ref.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));

    cublasHandle_t blas;
    VELES_CUBLAS(cublasCreate(&blas));
    struct BlasGuard {
        cublasHandle_t h;
        ~BlasGuard() { cublasDestroy(h); }
    } guard{blas};
    VELES_CUBLAS(cublasSetStream(blas, st));

    const uint32_t dpad = round_up(ix->dim, 8);
    const bool need_norm = ix->metric == VELES_EUCLIDEAN || ix->metric == VELES_HAMMING || ix->metric == VELES_JACCARD;
    const int sms = device_sm_count();
    DevBuf x16, sqn, ids_d, scores;
    VELES_TRY(x16.alloc((size_t)n * dpad * 2));
    VELES_TRY(sqn.alloc((size_t)n * 4));
    // score buffer: up to 2 GiB, at least one row
    VELES_TRY(scores.alloc(std::max<size_t>(std::min<size_t>((size_t)2 << 30, (size_t)(n + 4) * 4096 * 4), (size_t)(n + 4) * 4)));
    IndexView v = ix->view();
    for (uint32_t layer = 0; layer <= max_layer; ++layer) {
        std::vector<uint32_t> ids;
        uint32_t n_l;
        const uint32_t* idp = nullptr;
        if (layer == 0) {
            n_l = (uint32_t)n;
        } else {
            for (uint64_t i = 0; i < n; ++i)
                if (level[i] >= layer) ids.push_back((uint32_t)i);
            n_l = (uint32_t)ids.size();
            VELES_TRY(ids_d.alloc((size_t)n_l * 4));
            VELES_CUDA(cudaMemcpyAsync(ids_d.p, ids.data(), (size_t)n_l * 4, cudaMemcpyHostToDevice, st));
            idp = ids_d.as<uint32_t>();
        }
        make_f16_rows<<<sms * 8, 256, 0, st>>>(v, idp, n_l, dpad, x16.as<__half>(), sqn.as<float>());
        count_launch();
        VELES_CUDA(cudaGetLastError());
        VELES_TRY(build_layer(ix, blas, layer, idp, n_l, layer == 0 ? M0 : M, x16.as<__half>(),
                              need_norm ? sqn.as<float>() : nullptr, dpad, scores, st));
    }
    ix->entry = entry;
    ix->max_layer = max_layer;
    ix->num_layers = max_layer + 1;
    ix->has_entry = true;
    ix->has_graph = true;
    VELES_CUDA(cudaStreamSynchronize(st));
    return VELES_OK;
}

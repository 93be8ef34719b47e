// sq8.cu -- SQ8 dual precision: DualPrecisionHnsw (native/dual_precision.rs:60-441) over
// ScalarQuantizer / QuantizedVectorStore (native/quantization.rs:160-374).
//
//   attach   train the quantizer on the first T vectors (quantization.rs:190-233), code every vector
//            (quantize, :236-250) into a second device store of u8 rows (dim codes, zero padded to 16 B)
//   search   search_with_config's int8 path (dual_precision.rs:284-325): the traversal is the same kernel as
//            the f32 search (hnsw_search.cu) over the u8 rows -- 4x fewer gathered bytes per candidate, exact
//            integer distances -- followed by sq8_rerank_kernel: exact f32 graph distance of the
//            k * oversampling coarse candidates, stable sort by total_cmp, cut to k.
#include <cfloat>

#include "index.hpp"

namespace veles {

// per-dimension min / max over rows [0, t) -> min, scale, inv_scale.  min/max are order independent
// (f32::min / f32::max ignore NaN exactly like fminf / fmaxf).
__global__ void sq8_train_kernel(const uint8_t* __restrict__ vecs, uint32_t row_bytes, uint32_t dim, uint64_t t,
                                 float* __restrict__ mn, float* __restrict__ scale, float* __restrict__ inv) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (uint64_t r = 0; r < t; ++r) {
        const float v = reinterpret_cast<const float*>(vecs + r * row_bytes)[i];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    const float range = __fsub_rn(hi, lo);
    const float s = fabsf(range) < 1e-10f ? 1.0f : __fdiv_rn(255.0f, range);
    mn[i] = lo;
    scale[i] = s;
    inv[i] = __fdiv_rn(1.0f, s);
}

__device__ __forceinline__ uint32_t sq8_code(float v, float mn, float scale) {
    const float q = roundf(__fmul_rn(__fsub_rn(v, mn), scale));  // f32::round: half away from zero
    return (q != q) ? 0u : (uint32_t)fminf(fmaxf(q, 0.0f), 255.0f);  // NaN survives clamp, `as u8` makes it 0
}

// one thread per 4 codes (one u32 of the output row)
__global__ void sq8_quantize_kernel(const uint8_t* __restrict__ vecs, uint32_t row_bytes, uint32_t dim, uint64_t n,
                                    const float* __restrict__ mn, const float* __restrict__ scale,
                                    uint32_t* __restrict__ codes, uint32_t words_per_row) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * words_per_row) return;
    const uint64_t row = idx / words_per_row;
    const uint32_t w = (uint32_t)(idx % words_per_row);
    const float* v = reinterpret_cast<const float*>(vecs + row * row_bytes);
    uint32_t out = 0;
#pragma unroll
    for (uint32_t e = 0; e < 4; ++e) {
        const uint32_t i = 4 * w + e;
        if (i < dim) out |= sq8_code(v[i], mn[i], scale[i]) << (8 * e);
    }
    codes[idx] = out;
}

// Exact re-rank (dual_precision.rs:306-324): one warp per query.  keys[j] = (ord(dist_j), j) are unique, so a
// stable sort is a rank count.
__global__ void sq8_rerank_kernel(const IndexView ix, const float* __restrict__ queries, uint32_t nq,
                                  const uint32_t* __restrict__ coarse_ids, const uint32_t* __restrict__ coarse_cnt,
                                  uint32_t ck, uint32_t k, uint32_t* __restrict__ out_ids, float* __restrict__ out_dist,
                                  uint32_t* __restrict__ out_cnt) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t qi = blockIdx.x * (blockDim.x >> 5) + warp;
    if (qi >= nq) return;
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem) + (size_t)warp * ck;
    const float* q = queries + (size_t)qi * ix.dim;
    const uint32_t cnt = min(coarse_cnt[qi], ck);
    float norm_a = 0.0f;
    if (ix.metric == VELES_COSINE) norm_a = __fsqrt_rn(warp_tree_reduce<0>(q, q, ix.dim, lane));
    for (uint32_t j = 0; j < cnt; ++j) {
        const uint32_t id = coarse_ids[(size_t)qi * ck + j];
        const uint8_t* row = ix.vecs + (size_t)id * ix.row_bytes;
        float norm_b = 0.0f;
        if (ix.metric == VELES_COSINE) norm_b = *reinterpret_cast<const float*>(row + ix.norm_off);
        const float d = warp_metric(ix.metric, false, q, reinterpret_cast<const float*>(row), ix.dim, norm_a, norm_b, lane);
        if (lane == 0) keys[j] = ((uint64_t)ord_key(d) << 32) | j;
    }
    __syncwarp();
    const uint32_t outn = min(cnt, k);
    for (uint32_t j = lane; j < cnt; j += 32) {
        const uint64_t me = keys[j];
        uint32_t rank = 0;
        for (uint32_t i = 0; i < cnt; ++i) rank += keys[i] < me ? 1u : 0u;
        if (rank < k) {
            out_ids[(size_t)qi * k + rank] = coarse_ids[(size_t)qi * ck + j];
            out_dist[(size_t)qi * k + rank] = ord_unkey((uint32_t)(me >> 32));
        }
    }
    for (uint32_t j = outn + lane; j < k; j += 32) {
        out_ids[(size_t)qi * k + j] = VELES_INVALID_ID;
        out_dist[(size_t)qi * k + j] = __uint_as_float(0x7fc00000u);
    }
    if (lane == 0) out_cnt[qi] = outn;
}

static int32_t sq8_search(const veles_index* ix, const float* q_d, uint32_t nq, uint32_t k, uint32_t ef_search,
                          uint32_t oversampling, uint32_t* ids_d, float* dist_d, uint32_t* cnt_d, uint32_t* stats_d,
                          cudaStream_t st, SearchCtx** ctx_out = nullptr) {
    VELES_REQUIRE(ix->has_sq8, "snapshot has no SQ8 store; call veles_index_attach_sq8 first");
    VELES_REQUIRE(k >= 1 && oversampling >= 1, "k and oversampling must be >= 1");
    const uint64_t ck64 = (uint64_t)k * oversampling;
    VELES_REQUIRE(ck64 <= 4096, "k * oversampling must be <= 4096, got %llu", (unsigned long long)ck64);
    if (nq == 0) return VELES_OK;
    const uint32_t ck = (uint32_t)ck64;
    const uint32_t ef = std::max(ef_search, ck);  // dual_precision.rs:362
    VELES_TRY(ix->sq_ids_d.ensure((size_t)nq * ck * 4));
    VELES_TRY(ix->sq_dist_d.ensure((size_t)nq * ck * 4));
    VELES_TRY(ix->sq_cnt_d.ensure((size_t)nq * 4));
    SearchCtx* ctx = nullptr;
    VELES_TRY(acquire_ctx(ix, st, false, &ctx));  // the caller holds ix->mu
    if (ctx_out) *ctx_out = ctx;
    VELES_TRY(launch_search(ix, ix->view_sq8(), ctx, q_d, nq, ck, ef, ix->sq_ids_d.as<uint32_t>(), ix->sq_dist_d.as<float>(),
                            ix->sq_cnt_d.as<uint32_t>(), stats_d, st));
    const uint32_t warps = ck <= 512 ? 4 : 1;
    const size_t smem = (size_t)warps * ck * 8;
    sq8_rerank_kernel<<<(nq + warps - 1) / warps, warps * 32, smem, st>>>(ix->view(), q_d, nq, ix->sq_ids_d.as<uint32_t>(),
                                                                      ix->sq_cnt_d.as<uint32_t>(), ck, k, ids_d, dist_d,
                                                                      cnt_d);
    count_launch();
    VELES_CUDA(cudaGetLastError());
    return VELES_OK;
}

}  // namespace veles

using namespace veles;

extern "C" {

int32_t veles_index_attach_sq8(veles_index_t* idx, uint64_t train_count, void* stream) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    VELES_REQUIRE(idx->dtype == VELES_F32, "SQ8 dual precision needs an f32 snapshot (the re-rank reads the originals)");
    VELES_REQUIRE(train_count >= 1 && train_count <= idx->n,
                  "Cannot train on empty vectors: train_count must be in 1..%llu, got %llu", (unsigned long long)idx->n,
                  (unsigned long long)train_count);
    VELES_REQUIRE(idx->dim <= 32768, "SQ8 traversal supports at most 32768 dimensions");
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> g(idx->mu);
    idx->has_sq8 = false;
    idx->sq_row_bytes = round_up(idx->dim, 16);
    VELES_TRY(idx->sq_min.alloc((size_t)idx->dim * 4));
    VELES_TRY(idx->sq_scale.alloc((size_t)idx->dim * 4));
    VELES_TRY(idx->sq_inv.alloc((size_t)idx->dim * 4));
    VELES_TRY(idx->sq_codes.alloc((size_t)idx->n * idx->sq_row_bytes));
    sq8_train_kernel<<<(idx->dim + 63) / 64, 64, 0, st>>>(idx->vecs.as<uint8_t>(), idx->row_bytes, idx->dim, train_count,
                                                         idx->sq_min.as<float>(), idx->sq_scale.as<float>(),
                                                         idx->sq_inv.as<float>());
    count_launch();
    const uint32_t wpr = idx->sq_row_bytes / 4;
    const uint64_t total = idx->n * wpr;
    sq8_quantize_kernel<<<(uint32_t)((total + 255) / 256), 256, 0, st>>>(idx->vecs.as<uint8_t>(), idx->row_bytes, idx->dim,
                                                                        idx->n, idx->sq_min.as<float>(),
                                                                        idx->sq_scale.as<float>(),
                                                                        idx->sq_codes.as<uint32_t>(), wpr);
    count_launch();
    VELES_CUDA(cudaGetLastError());
    VELES_CUDA(cudaStreamSynchronize(st));
    idx->sq_train = train_count;
    idx->has_sq8 = true;
    return VELES_OK;
}

int32_t veles_index_has_sq8(const veles_index_t* idx) { return idx && idx->has_sq8 ? 1 : 0; }

int32_t veles_index_sq8_export(const veles_index_t* idx, float* min_vals, float* scales, float* inv_scales, uint8_t* codes) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    VELES_REQUIRE(idx->has_sq8, "snapshot has no SQ8 store");
    VELES_REQUIRE(min_vals && scales && inv_scales, "NULL buffer");
    const size_t b = (size_t)idx->dim * 4;
    VELES_CUDA(cudaMemcpy(min_vals, idx->sq_min.p, b, cudaMemcpyDeviceToHost));
    VELES_CUDA(cudaMemcpy(scales, idx->sq_scale.p, b, cudaMemcpyDeviceToHost));
    VELES_CUDA(cudaMemcpy(inv_scales, idx->sq_inv.p, b, cudaMemcpyDeviceToHost));
    if (codes && idx->n)
        VELES_CUDA(cudaMemcpy2D(codes, idx->dim, idx->sq_codes.p, idx->sq_row_bytes, idx->dim, idx->n, cudaMemcpyDeviceToHost));
    return VELES_OK;
}

int32_t veles_search_batch_sq8_d(const veles_index_t* idx, const float* queries_d, uint32_t nq, uint32_t k,
                                 uint32_t ef_search, uint32_t oversampling, uint32_t* out_node_ids_d,
                                 float* out_raw_dist_d, uint32_t* out_counts_d, uint32_t* out_stats_d, void* stream) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    VELES_REQUIRE(nq == 0 || (queries_d && out_node_ids_d && out_raw_dist_d && out_counts_d), "NULL buffer");
    std::lock_guard<std::mutex> g(idx->mu);
    return sq8_search(idx, queries_d, nq, k, ef_search, oversampling, out_node_ids_d, out_raw_dist_d, out_counts_d,
                      out_stats_d, (cudaStream_t)stream);
}

int32_t veles_search_batch_sq8(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k, uint32_t ef_search,
                               uint32_t oversampling, uint32_t* out_node_ids, float* out_raw_dist, uint32_t* out_counts,
                               uint32_t* out_stats, void* stream) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    VELES_REQUIRE(nq == 0 || (queries && out_node_ids && out_raw_dist && out_counts), "NULL buffer");
    if (nq == 0) return VELES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> g(idx->mu);
    const size_t qb = (size_t)nq * idx->dim * 4, ob = (size_t)nq * k * 4;
    VELES_TRY(idx->q_d.ensure(qb));
    VELES_TRY(idx->out_ids_d.ensure(ob));
    VELES_TRY(idx->out_val_d.ensure(ob));
    VELES_TRY(idx->out_cnt_d.ensure((size_t)nq * 4));
    if (out_stats) VELES_TRY(idx->out_stats_d.ensure((size_t)nq * 16));
    VELES_CUDA(cudaMemcpyAsync(idx->q_d.p, queries, qb, cudaMemcpyHostToDevice, st));
    SearchCtx* ctx = nullptr;
    VELES_TRY(sq8_search(idx, idx->q_d.as<float>(), nq, k, ef_search, oversampling, idx->out_ids_d.as<uint32_t>(),
                         idx->out_val_d.as<float>(), idx->out_cnt_d.as<uint32_t>(),
                         out_stats ? idx->out_stats_d.as<uint32_t>() : nullptr, st, &ctx));
    VELES_CUDA(cudaMemcpyAsync(out_node_ids, idx->out_ids_d.p, ob, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaMemcpyAsync(out_raw_dist, idx->out_val_d.p, ob, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaMemcpyAsync(out_counts, idx->out_cnt_d.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (out_stats) VELES_CUDA(cudaMemcpyAsync(out_stats, idx->out_stats_d.p, (size_t)nq * 16, cudaMemcpyDeviceToHost, st));
    return check_search_error_flag(ctx, st);
}

}  // extern "C"

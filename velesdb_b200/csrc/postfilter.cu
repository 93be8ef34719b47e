// postfilter.cu -- what HnswIndex does with the raw hits of NativeHnsw::search, on the device:
//   node index -> external id through the id map, hits of removed (tombstoned) nodes dropped
//     (mappings.get_id, index/hnsw/index/search.rs:86-91, batch.rs:186-194; soft delete trait_impl.rs:54-58),
//   raw distance -> score (transform_score, native/backend_adapter.rs:160-168),
//   and the over-fetch + filter + take(k) of Collection::search_with_filter
//     (collection/search/vector.rs:182-211) with the filter given as a bitmap over node indices.
// One search launch + one warp per query here; nothing is re-sorted (the reference does not re-sort either:
// hits stay in traversal order, and transform_score is monotone).
#include "index.hpp"

namespace veles {

__device__ __forceinline__ float transform_score_dev(int metric, float d) {
    if (metric == VELES_COSINE) {
        if (d != d) return d;  // f32::clamp keeps NaN
        return fminf(fmaxf(__fsub_rn(1.0f, d), 0.0f), 1.0f);
    }
    if (metric == VELES_DOT) return -d;
    return d;
}

__global__ void map_results_kernel(const uint32_t* __restrict__ ids, const float* __restrict__ dist,
                                   const uint32_t* __restrict__ cnt, uint32_t nq, uint32_t k_fetch, uint32_t k_out,
                                   const uint64_t* __restrict__ id_map, const uint32_t* __restrict__ live,
                                   const uint32_t* __restrict__ allow, int metric, uint64_t* __restrict__ out_ids,
                                   float* __restrict__ out_score, uint32_t* __restrict__ out_cnt) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const uint32_t n = min(cnt[q], k_fetch);
    uint32_t w = 0;
    for (uint32_t base = 0; base < n && w < k_out; base += 32) {
        const uint32_t j = base + lane;
        uint32_t node = 0;
        bool ok = j < n;
        if (ok) {
            node = ids[(size_t)q * k_fetch + j];
            if (live) ok = (live[node >> 5] >> (node & 31)) & 1u;
            if (ok && allow) ok = (allow[node >> 5] >> (node & 31)) & 1u;
        }
        const uint32_t mask = __ballot_sync(FULL_MASK, ok);
        const uint32_t pos = w + __popc(mask & ((1u << lane) - 1u));
        if (ok && pos < k_out) {
            out_ids[(size_t)q * k_out + pos] = id_map ? id_map[node] : (uint64_t)node;
            out_score[(size_t)q * k_out + pos] = transform_score_dev(metric, dist[(size_t)q * k_fetch + j]);
        }
        w += __popc(mask);
    }
    w = min(w, k_out);
    for (uint32_t j = w + lane; j < k_out; j += 32) {
        out_ids[(size_t)q * k_out + j] = ~0ull;
        out_score[(size_t)q * k_out + j] = __uint_as_float(0x7fc00000u);
    }
    if (lane == 0) out_cnt[q] = w;
}

}  // namespace veles

using namespace veles;

extern "C" {

int32_t veles_index_set_id_map(veles_index_t* idx, const uint64_t* ext_ids, const uint32_t* live_bits) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    std::lock_guard<std::mutex> g(idx->mu);
    const size_t words = (size_t)(idx->n + 31) / 32;
    if (ext_ids) {
        VELES_TRY(idx->id_map_d.alloc((size_t)idx->n * 8));
        if (idx->n) VELES_CUDA(cudaMemcpy(idx->id_map_d.p, ext_ids, (size_t)idx->n * 8, cudaMemcpyHostToDevice));
    } else {
        idx->id_map_d.release();
    }
    if (live_bits) {
        VELES_TRY(idx->live_d.alloc(words * 4));
        if (words) VELES_CUDA(cudaMemcpy(idx->live_d.p, live_bits, words * 4, cudaMemcpyHostToDevice));
    } else {
        idx->live_d.release();
    }
    return VELES_OK;
}

int32_t veles_search_batch_mapped(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k_fetch, uint32_t k_out,
                                  uint32_t ef, const uint32_t* allow_bits, uint64_t* out_ids, float* out_scores,
                                  uint32_t* out_counts, void* stream) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    VELES_REQUIRE(nq == 0 || (queries && out_ids && out_scores && out_counts), "NULL buffer");
    VELES_REQUIRE(k_out >= 1 && k_fetch >= k_out, "need 1 <= k_out <= k_fetch, got %u / %u", k_out, k_fetch);
    if (nq == 0) return VELES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> g(idx->mu);
    const size_t qb = (size_t)nq * idx->dim * 4, fb = (size_t)nq * k_fetch * 4;
    const size_t words = (size_t)(idx->n + 31) / 32;
    VELES_TRY(idx->q_d.ensure(qb));
    VELES_TRY(idx->out_ids_d.ensure(fb));
    VELES_TRY(idx->out_val_d.ensure(fb));
    VELES_TRY(idx->out_cnt_d.ensure((size_t)nq * 8));
    VELES_TRY(idx->map_ids_d.ensure((size_t)nq * k_out * 8));
    VELES_TRY(idx->map_score_d.ensure((size_t)nq * k_out * 4));
    VELES_CUDA(cudaMemcpyAsync(idx->q_d.p, queries, qb, cudaMemcpyHostToDevice, st));
    if (allow_bits) {
        VELES_TRY(idx->allow_d.ensure(std::max<size_t>(words, 1) * 4));
        VELES_CUDA(cudaMemcpyAsync(idx->allow_d.p, allow_bits, words * 4, cudaMemcpyHostToDevice, st));
    }
    uint32_t* raw_cnt = idx->out_cnt_d.as<uint32_t>();
    uint32_t* map_cnt = raw_cnt + nq;
    SearchCtx* ctx = nullptr;
    VELES_TRY(acquire_ctx(idx, st, false, &ctx));
    VELES_TRY(launch_search(idx, idx->view(), ctx, idx->q_d.as<float>(), nq, k_fetch, ef, idx->out_ids_d.as<uint32_t>(),
                            idx->out_val_d.as<float>(), raw_cnt, nullptr, st));
    map_results_kernel<<<(nq + 3) / 4, 128, 0, st>>>(idx->out_ids_d.as<uint32_t>(), idx->out_val_d.as<float>(), raw_cnt, nq, k_fetch,
                                                    k_out, idx->id_map_d.as<uint64_t>(), idx->live_d.as<uint32_t>(),
                                                    allow_bits ? idx->allow_d.as<uint32_t>() : nullptr, idx->metric,
                                                    idx->map_ids_d.as<uint64_t>(), idx->map_score_d.as<float>(), map_cnt);
    count_launch();
    VELES_CUDA(cudaGetLastError());
    VELES_CUDA(cudaMemcpyAsync(out_ids, idx->map_ids_d.p, (size_t)nq * k_out * 8, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaMemcpyAsync(out_scores, idx->map_score_d.p, (size_t)nq * k_out * 4, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaMemcpyAsync(out_counts, map_cnt, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    return check_search_error_flag(ctx, st);
}

}  // extern "C"

// fusion.cu -- rank/score fusion on the device.
//   * veles_rrf_hybrid: the RRF inside Collection::hybrid_search (collection/search/text.rs:133-180),
//     batched: one warp per query.
//   * veles_fuse: FusionStrategy::{fuse_average, fuse_maximum, fuse_rrf, fuse_weighted}
//     (fusion/strategy.rs:138-300) for one multi-query request.
#include <algorithm>
#include <cmath>

#include "index.hpp"

namespace veles {

// ---- hybrid RRF ------------------------------------------------------------------------------
// fused[id] += w / (rank0 + 60) over the vector list, then += (1-w) / (rank0 + 60) over the text list
// (f32 adds in that order, text.rs:152-162); keep the k largest (score, id) (min-heap of size k on the
// tuple, text.rs:166-173); emit score-descending, equal scores by descending id (canonical form of the
// heap's internal order).
__global__ void __launch_bounds__(32) rrf_hybrid_kernel(const uint32_t* __restrict__ vec_ids, const uint32_t* __restrict__ vec_cnt,
                                                        const uint32_t* __restrict__ txt_ids, const uint32_t* __restrict__ txt_cnt,
                                                        uint32_t in_k, float w, uint32_t k, uint32_t* __restrict__ out_ids,
                                                        float* __restrict__ out_score, uint32_t* __restrict__ out_counts) {
    extern __shared__ __align__(16) uint8_t rrf_smem[];
    uint32_t* ids = reinterpret_cast<uint32_t*>(rrf_smem);                       // 2*in_k
    float* sc = reinterpret_cast<float*>(rrf_smem + (size_t)2 * in_k * 4);        // 2*in_k
    uint64_t* res = reinterpret_cast<uint64_t*>(rrf_smem + (size_t)4 * in_k * 4); // k
    const uint32_t lane = threadIdx.x, q = blockIdx.x;
    const uint32_t nv = min(vec_cnt[q], in_k), nt = min(txt_cnt[q], in_k);
    const float tw = __fsub_rn(1.0f, w);
    uint32_t n = 0;
    for (uint32_t e = 0; e < nv + nt; ++e) {
        const bool is_vec = e < nv;
        const uint32_t rank = is_vec ? e : e - nv;
        const uint32_t id = is_vec ? vec_ids[(size_t)q * in_k + rank] : txt_ids[(size_t)q * in_k + rank];
        const float s = __fdiv_rn(is_vec ? w : tw, __fadd_rn((float)rank, 60.0f));
        // find id among the entries seen so far
        uint32_t found = 0xffffffffu;
        for (uint32_t b = 0; b < n; b += 32) {
            const uint32_t i = b + lane;
            const uint32_t msk = __ballot_sync(FULL_MASK, i < n && ids[i] == id);
            if (msk) {
                found = b + __ffs(msk) - 1;
                break;
            }
        }
        if (lane == 0) {
            if (found != 0xffffffffu) {
                sc[found] = __fadd_rn(sc[found], s);
            } else {
                ids[n] = id;
                sc[n] = __fadd_rn(0.0f, s);
            }
        }
        if (found == 0xffffffffu) ++n;
        __syncwarp();
    }
    // k largest (score, id): ascending on the complemented key
    uint32_t len = 0;
    uint64_t worst = ~0ull;
    for (uint32_t b = 0; b < n; b += 32) {
        const uint32_t i = b + lane;
        uint64_t key = ~0ull;
        if (i < n) key = ~(((uint64_t)ord_key(sc[i]) << 32) | ids[i]);
        uint32_t msk = __ballot_sync(FULL_MASK, key < worst);
        while (msk) {
            const uint32_t src = __ffs(msk) - 1;
            msk &= msk - 1;
            const uint64_t kk = __shfl_sync(FULL_MASK, key, src);
            if (kk >= worst) continue;
            const uint32_t pos = lower_bound_warp(res, len, kk, lane);
            if (len < k) {
                insert_at(res, pos, len + 1, kk, lane);
                ++len;
            } else {
                insert_at(res, pos, len, kk, lane);
            }
            if (len == k) worst = res[k - 1];
        }
    }
    __syncwarp();
    for (uint32_t j = lane; j < k; j += 32) {
        uint32_t id = VELES_INVALID_ID;
        float s = __uint_as_float(0x7fc00000u);
        if (j < len) {
            const uint64_t key = ~res[j];
            id = (uint32_t)key;
            s = ord_unkey((uint32_t)(key >> 32));
        }
        out_ids[(size_t)q * k + j] = id;
        out_score[(size_t)q * k + j] = s;
    }
    if (lane == 0) out_counts[q] = len;
}

// ---- FusionStrategy::fuse ----------------------------------------------------------------------
// One request is small (<= 10 lists x overfetch_k hits, collection/search/batch.rs:238,270-275: up to 10 x 500) and
// the reference semantics are sequential per document (contributions in list order), so thread 0
// walks the lists with an open-addressing table and the block then sorts.  The tables are sized from the request
// (docs = total hits, slots = the power of two >= 2 x that): in shared memory up to kFuseSmemDocs documents, in
// global memory above.
constexpr uint32_t kFuseSmemDocs = 4096;
constexpr uint32_t kFuseMaxDocs = 1u << 20;

struct FuseDoc {
    uint32_t id;
    uint32_t cnt;        // lists the doc appeared in
    float sum;           // Average / Weighted: sum of per-list best scores; RRF: sum of reciprocal ranks
    float mx;            // Maximum / Weighted
    float list_best;     // best score within the list being processed
    uint32_t list_tag;   // list index + 1 of the last list that touched this doc
    uint32_t list_rank;  // first rank within that list
};

__global__ void __launch_bounds__(256) fuse_kernel(int strategy, const uint32_t* __restrict__ list_ptr, uint32_t n_lists,
                                                   const uint32_t* __restrict__ ids, const float* __restrict__ scores,
                                                   float rrf_k, float avg_w, float max_w, float hit_w, uint32_t cap,
                                                   uint32_t* __restrict__ out_ids, float* __restrict__ out_score,
                                                   uint32_t* __restrict__ out_count, uint32_t* __restrict__ err,
                                                   uint32_t doc_cap, uint32_t hash_slots, uint8_t* __restrict__ work) {
    extern __shared__ __align__(16) uint8_t fz_smem[];
    uint8_t* mem = work ? work : fz_smem;  // work: global-memory tables for requests beyond kFuseSmemDocs
    FuseDoc* docs = reinterpret_cast<FuseDoc*>(mem);
    uint32_t* table = reinterpret_cast<uint32_t*>(mem + sizeof(FuseDoc) * doc_cap);
    uint64_t* keys = reinterpret_cast<uint64_t*>(mem + sizeof(FuseDoc) * doc_cap + (size_t)hash_slots * 4);
    __shared__ uint32_t s_n;
    for (uint32_t i = threadIdx.x; i < hash_slots; i += blockDim.x) table[i] = 0xffffffffu;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t n = 0;
        bool overflow = false;
        auto flush_list = [&](uint32_t tag) {
            // fold the finished list's per-doc best / first rank into the cross-list accumulators
            for (uint32_t i = 0; i < n; ++i) {
                FuseDoc& d = docs[i];
                if (d.list_tag != tag) continue;
                if (strategy == 2) {
                    const float rr = __fdiv_rn(1.0f, __fadd_rn(rrf_k, (float)(d.list_rank + 1)));
                    d.sum = __fadd_rn(d.sum, rr);
                } else {
                    d.sum = __fadd_rn(d.sum, d.list_best);
                    d.mx = d.cnt == 0 ? d.list_best : fmaxf(d.mx, d.list_best);
                }
                d.cnt += 1;
            }
        };
        for (uint32_t l = 0; l < n_lists && !overflow; ++l) {
            const uint32_t tag = l + 1;
            for (uint32_t e = list_ptr[l]; e < list_ptr[l + 1]; ++e) {
                const uint32_t id = ids[e];
                uint32_t h = (id * 2654435761u) & (hash_slots - 1);
                uint32_t slot;
                for (;;) {
                    slot = table[h];
                    if (slot == 0xffffffffu || docs[slot].id == id) break;
                    h = (h + 1) & (hash_slots - 1);
                }
                if (slot == 0xffffffffu) {
                    if (n == doc_cap) {
                        overflow = true;
                        break;
                    }
                    slot = n++;
                    table[h] = slot;
                    docs[slot] = FuseDoc{id, 0u, 0.0f, 0.0f, 0.0f, 0u, 0u};
                }
                FuseDoc& d = docs[slot];
                if (d.list_tag != tag) {
                    d.list_tag = tag;
                    d.list_best = scores[e];
                    d.list_rank = e - list_ptr[l];
                } else {
                    d.list_best = fmaxf(d.list_best, scores[e]);  // f32::max (strategy.rs:178-181)
                }
            }
            flush_list(tag);
        }
        if (overflow) *err = 1;
        s_n = overflow ? 0 : n;
    }
    __syncthreads();
    const uint32_t n = s_n;
    // final score per doc, then sort by (score desc, id asc)
    uint32_t npad = 1;
    while (npad < n) npad <<= 1;
    for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) {
        uint64_t key = ~0ull;
        if (i < n) {
            const FuseDoc& d = docs[i];
            float v;
            if (strategy == 1) {
                v = d.mx;
            } else if (strategy == 2) {
                v = d.sum;
            } else {
                const float avg = __fdiv_rn(d.sum, (float)d.cnt);
                if (strategy == 0) {
                    v = avg;
                } else {
                    const float hit = __fdiv_rn((float)d.cnt, (float)n_lists);
                    v = __fadd_rn(__fadd_rn(__fmul_rn(avg_w, avg), __fmul_rn(max_w, d.mx)), __fmul_rn(hit_w, hit));
                }
            }
            key = ((uint64_t)(~ord_key(v)) << 32) | d.id;
        }
        keys[i] = key;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= npad; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) {
                const uint32_t j = i ^ stride;
                if (j > i) {
                    const bool up = (i & size) == 0;
                    const uint64_t a = keys[i], b = keys[j];
                    if ((a > b) == up) {
                        keys[i] = b;
                        keys[j] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    const uint32_t m = min(n, cap);
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        out_ids[i] = (uint32_t)keys[i];
        out_score[i] = ord_unkey(~(uint32_t)(keys[i] >> 32));
    }
    if (threadIdx.x == 0) *out_count = m;
}

}  // namespace veles

using namespace veles;

namespace veles {
// rrf_hybrid_kernel over lists that already live on the device (both legs of veles_hybrid_search_batch leave theirs
// there); only enqueues.  `vector_weight` is clamped as text.rs:133 does.
int32_t rrf_hybrid_enqueue_d(const uint32_t* vec_ids_d, const uint32_t* vec_cnt_d, const uint32_t* txt_ids_d,
                             const uint32_t* txt_cnt_d, uint32_t nq, uint32_t in_k, float vector_weight, uint32_t k,
                             uint32_t* out_ids_d, float* out_score_d, uint32_t* out_counts_d, cudaStream_t st) {
    float w = vector_weight;  // f32::clamp(0.0, 1.0), text.rs:133
    if (w < 0.0f) w = 0.0f;
    if (w > 1.0f) w = 1.0f;
    const size_t smem = (size_t)4 * in_k * 4 + (size_t)k * 8;
    VELES_CUDA(cudaFuncSetAttribute(rrf_hybrid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rrf_hybrid_kernel<<<nq, 32, smem, st>>>(vec_ids_d, vec_cnt_d, txt_ids_d, txt_cnt_d, in_k, w, k, out_ids_d, out_score_d,
                                            out_counts_d);
    count_launch();
    VELES_CUDA(cudaGetLastError());
    return VELES_OK;
}
}  // namespace veles

extern "C" {

int32_t veles_rrf_hybrid(const uint32_t* vec_ids, const uint32_t* vec_cnt, const uint32_t* txt_ids, const uint32_t* txt_cnt,
                         uint32_t nq, uint32_t in_k, float vector_weight, uint32_t k, uint32_t* out_ids, float* out_score,
                         uint32_t* out_counts, void* stream) {
    VELES_REQUIRE(nq == 0 || (vec_ids && vec_cnt && txt_ids && txt_cnt && out_ids && out_score && out_counts), "NULL buffer");
    VELES_REQUIRE(k >= 1 && k <= 4096, "k must be in 1..4096, got %u", k);
    VELES_REQUIRE(in_k >= 1 && in_k <= 4096, "in_k must be in 1..4096, got %u", in_k);
    if (nq == 0) return VELES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf dv, dvc, dt, dtc, oi, os, oc;
    const size_t lb = (size_t)nq * in_k * 4, ob = (size_t)nq * k * 4;
    VELES_TRY(dv.alloc(lb));
    VELES_TRY(dt.alloc(lb));
    VELES_TRY(dvc.alloc((size_t)nq * 4));
    VELES_TRY(dtc.alloc((size_t)nq * 4));
    VELES_TRY(oi.alloc(ob));
    VELES_TRY(os.alloc(ob));
    VELES_TRY(oc.alloc((size_t)nq * 4));
    VELES_CUDA(cudaMemcpyAsync(dv.p, vec_ids, lb, cudaMemcpyHostToDevice, st));
    VELES_CUDA(cudaMemcpyAsync(dt.p, txt_ids, lb, cudaMemcpyHostToDevice, st));
    VELES_CUDA(cudaMemcpyAsync(dvc.p, vec_cnt, (size_t)nq * 4, cudaMemcpyHostToDevice, st));
    VELES_CUDA(cudaMemcpyAsync(dtc.p, txt_cnt, (size_t)nq * 4, cudaMemcpyHostToDevice, st));
    VELES_TRY(rrf_hybrid_enqueue_d(dv.as<uint32_t>(), dvc.as<uint32_t>(), dt.as<uint32_t>(), dtc.as<uint32_t>(), nq, in_k,
                                   vector_weight, k, oi.as<uint32_t>(), os.as<float>(), oc.as<uint32_t>(), st));
    VELES_CUDA(cudaMemcpyAsync(out_ids, oi.p, ob, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaMemcpyAsync(out_score, os.p, ob, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaMemcpyAsync(out_counts, oc.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaStreamSynchronize(st));
    return VELES_OK;
}

int32_t veles_fuse(int32_t strategy, const uint32_t* list_ptr, uint32_t n_lists, const uint32_t* ids, const float* scores,
                   uint32_t rrf_k, float avg_w, float max_w, float hit_w, uint32_t cap, uint32_t* out_ids, float* out_score,
                   uint32_t* out_count, void* stream) {
    VELES_REQUIRE(strategy >= 0 && strategy <= 3, "unknown fusion strategy %d", strategy);
    VELES_REQUIRE(out_count != nullptr, "out_count is NULL");
    *out_count = 0;
    if (n_lists == 0) return VELES_OK;  // strategy.rs:139-141
    VELES_REQUIRE(list_ptr != nullptr, "list_ptr is NULL");
    const uint32_t total = list_ptr[n_lists];
    if (total == 0) return VELES_OK;  // every list empty (strategy.rs:144-147)
    VELES_REQUIRE(ids && scores && out_ids && out_score, "NULL buffer");
    VELES_REQUIRE(cap >= 1, "cap must be >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    VELES_REQUIRE(total <= kFuseMaxDocs, "veles_fuse: at most %u hits per request, got %u", kFuseMaxDocs, total);
    // distinct documents <= total hits; the sort works on the next power of two
    uint32_t doc_cap = 1;
    while (doc_cap < total) doc_cap <<= 1;
    const uint32_t hash_slots = doc_cap * 2;
    DevBuf dp, di, ds, oi, os, oc, work;
    VELES_TRY(dp.alloc(((size_t)n_lists + 1) * 4));
    VELES_TRY(di.alloc((size_t)total * 4));
    VELES_TRY(ds.alloc((size_t)total * 4));
    VELES_TRY(oi.alloc((size_t)doc_cap * 4));
    VELES_TRY(os.alloc((size_t)doc_cap * 4));
    VELES_TRY(oc.alloc(8));
    VELES_CUDA(cudaMemsetAsync(oc.p, 0, 8, st));
    VELES_CUDA(cudaMemcpyAsync(dp.p, list_ptr, ((size_t)n_lists + 1) * 4, cudaMemcpyHostToDevice, st));
    VELES_CUDA(cudaMemcpyAsync(di.p, ids, (size_t)total * 4, cudaMemcpyHostToDevice, st));
    VELES_CUDA(cudaMemcpyAsync(ds.p, scores, (size_t)total * 4, cudaMemcpyHostToDevice, st));
    const size_t table_bytes = sizeof(FuseDoc) * doc_cap + (size_t)hash_slots * 4 + (size_t)doc_cap * 8;
    size_t smem = table_bytes;
    if (doc_cap > kFuseSmemDocs) {
        VELES_TRY(work.alloc(table_bytes));
        smem = 0;
    }
    VELES_CUDA(cudaFuncSetAttribute(fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    const uint32_t dcap = std::min(cap, doc_cap);
    fuse_kernel<<<1, 256, smem, st>>>(strategy, dp.as<uint32_t>(), n_lists, di.as<uint32_t>(), ds.as<float>(), (float)rrf_k, avg_w,
                                      max_w, hit_w, dcap, oi.as<uint32_t>(), os.as<float>(), oc.as<uint32_t>(),
                                      oc.as<uint32_t>() + 1, doc_cap, hash_slots, smem ? nullptr : work.as<uint8_t>());
    count_launch();
    VELES_CUDA(cudaGetLastError());
    uint32_t h[2] = {0, 0};
    VELES_CUDA(cudaMemcpyAsync(h, oc.p, 8, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaStreamSynchronize(st));
    if (h[1]) {
        set_error("veles_fuse: internal table overflow (%u documents)", doc_cap);
        return VELES_ERR_OVERFLOW;
    }
    if (h[0]) {
        VELES_CUDA(cudaMemcpyAsync(out_ids, oi.p, (size_t)h[0] * 4, cudaMemcpyDeviceToHost, st));
        VELES_CUDA(cudaMemcpyAsync(out_score, os.p, (size_t)h[0] * 4, cudaMemcpyDeviceToHost, st));
        VELES_CUDA(cudaStreamSynchronize(st));
    }
    *out_count = h[0];
    return VELES_OK;
}

}  // extern "C"

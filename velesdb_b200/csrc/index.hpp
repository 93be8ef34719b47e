// index.hpp -- host-side handle of a device snapshot (one NativeHnsw, native/graph.rs:18-44).
#pragma once

#include <memory>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace veles {
// Everything one search launch writes besides its outputs: visited bitmaps, logs, tie lists and the work counter /
// error flag, plus the device staging buffers of the host-pointer entry points.  An index owns a small pool of these
// (veles_index::ctxs); a context serves one stream at a time, so launches on different streams or from different
// host threads never share scratch (the reference's HnswIndex is Send + Sync behind an Arc; searches only take a
// read lock, index/hnsw/index/search.rs:59-94).
struct SearchCtx {
    DevBuf visited, vlog, tie, counters;  // counters: [0] work counter (zeroed per launch), [1] sticky overflow flag
    uint32_t slots = 0, vis_words = 0;
    DevBuf q_d, ids_d, val_d, cnt_d, stats_d, extra_d;  // staging of host-pointer calls
    cudaStream_t bound = nullptr;   // stream of the last launch
    cudaEvent_t done = nullptr;     // recorded after the last launch
    cudaStream_t own = nullptr;     // this context's stream, for veles_search_submit / _wait
    uint32_t* h_flag = nullptr;     // pinned: the overflow flag read back by a submitted batch
    uint64_t generation = 0;        // ticket check
    bool launched = false;          // `done` has been recorded at least once
    bool in_use = false;            // checked out by a host-pointer call or a pending ticket
    SearchCtx() = default;
    SearchCtx(const SearchCtx&) = delete;
    SearchCtx& operator=(const SearchCtx&) = delete;
    ~SearchCtx() {
        if (done) cudaEventDestroy(done);
        if (own) cudaStreamDestroy(own);
        if (h_flag) cudaFreeHost(h_flag);
    }
};
}  // namespace veles

struct veles_index {
    int device = 0;
    int32_t metric = VELES_COSINE;
    int32_t dtype = VELES_F32;
    uint32_t dim = 0;
    uint64_t n = 0;
    uint32_t row_bytes = 0;
    uint32_t norm_off = 0;
    uint32_t M = 0, M0 = 0, ef_construction = 0;
    uint32_t stride0 = 0, strideU = 0;
    uint64_t upper_rows = 0;
    uint64_t entry = 0;
    uint32_t max_layer = 0;
    uint32_t num_layers = 1;
    bool has_entry = false;
    bool has_graph = false;

    veles::DevBuf vecs, adj0, upper_ref, upper_adj;

    // SQ8 traversal store (DualPrecisionHnsw, native/dual_precision.rs): u8 codes in rows of sq_row_bytes
    // (dim codes, zero padded to 16) + the quantizer
    veles::DevBuf sq_codes, sq_min, sq_scale, sq_inv;
    uint32_t sq_row_bytes = 0;
    uint64_t sq_train = 0;
    bool has_sq8 = false;
    mutable veles::DevBuf sq_ids_d, sq_dist_d, sq_cnt_d;  // coarse candidates between traversal and re-rank

    // fp16 copy of the collection for the tensor-core candidate GEMM (gemm_tc.cu), built on first use: rows of
    // x16_dpad halves (dim padded to 64), cosine rows pre-normalised; x16_bias = |row16|^2 (L2)
    mutable veles::DevBuf x16, x16_bias;
    mutable uint32_t x16_dpad = 0;
    mutable veles::DevBuf tc_q16, tc_tiles, tc_sample, tc_thr, tc_cnt, tc_cand, tc_err;  // work buffers of that path
    mutable uint32_t tc_tiles_key[2] = {0, 0};  // (sample tiles, row tiles) the list in tc_tiles was written for

    // node index -> external id, live (not tombstoned) bitmap: ShardedMappings on the device (postfilter.cu)
    veles::DevBuf id_map_d, live_d;
    mutable veles::DevBuf allow_d, map_ids_d, map_score_d, extra_d;

    // `mu` guards the context pool and the shared staging buffers below
    mutable std::mutex mu;
    mutable std::vector<std::unique_ptr<veles::SearchCtx>> ctxs;
    mutable std::vector<cudaStream_t> overflowed;  // streams whose overflow flag was harvested when a context moved on
    mutable veles::DevBuf q_d, out_ids_d, out_val_d, out_cnt_d, out_stats_d, scores_d, topk_d, bf_aux, bf_done, bf_work;

    veles::IndexView view() const {
        veles::IndexView v;
        v.vecs = vecs.as<uint8_t>();
        v.adj0 = adj0.as<uint32_t>();
        v.upper_ref = upper_ref.as<uint32_t>();
        v.upper_adj = upper_adj.as<uint32_t>();
        v.n = n;
        v.dim = dim;
        v.row_bytes = row_bytes;
        v.norm_off = norm_off;
        v.stride0 = stride0;
        v.strideU = strideU;
        v.entry = (uint32_t)entry;
        v.max_layer = max_layer;
        v.metric = metric;
        v.dtype = dtype;
        v.has_entry = (has_entry && has_graph) ? 1 : 0;
        v.sq_min = nullptr;
        v.sq_scale = nullptr;
        return v;
    }
    // the same graph over the u8 codes
    veles::IndexView view_sq8() const {
        veles::IndexView v = view();
        v.vecs = sq_codes.as<uint8_t>();
        v.row_bytes = sq_row_bytes;
        v.norm_off = 0;
        v.dtype = VELES_SQ8;
        v.sq_min = sq_min.as<float>();
        v.sq_scale = sq_scale.as<float>();
        return v;
    }
    uint64_t device_bytes() const {
        uint64_t b = vecs.bytes + adj0.bytes + upper_ref.bytes + upper_adj.bytes + sq_codes.bytes;
        for (const auto& c : ctxs) b += c->visited.bytes + c->vlog.bytes + c->tie.bytes;
        return b;
    }
};

namespace veles {
inline uint32_t round_up(uint32_t x, uint32_t m) { return (x + m - 1) / m * m; }
inline uint32_t elt_bytes(int32_t dtype) { return dtype == VELES_F32 ? 4 : 2; }
// sets row_bytes / norm_off from dim, dtype, metric
void compute_row_layout(veles_index* ix);
// uploads vectors into the row layout and fills the cosine norm trailer
int32_t upload_vectors(veles_index* ix, const void* vectors, int32_t src_dtype, cudaStream_t st);
// installs a padded fixed-stride adjacency built on the host
int32_t install_graph_host(veles_index* ix, uint32_t num_layers, const uint64_t* const* row_ptr,
                           const uint32_t* const* cols, const uint64_t* layer_nodes, uint32_t M, uint32_t M0,
                           uint64_t entry_point, uint32_t max_layer);
int device_sm_count();
// A context for a launch on `st` (hnsw_search.cu); the caller holds ix->mu.  Prefers the context already bound to
// `st` (stream order makes reuse safe), then an idle one, then a new one (at most kMaxCtx, after that it waits for
// the oldest).  `exclusive` marks it in_use until release_ctx.
int32_t acquire_ctx(const veles_index* ix, cudaStream_t st, bool exclusive, SearchCtx** out);
void release_ctx(const veles_index* ix, SearchCtx* c);
// extra destinations of a launch's results (peer GPUs' gather windows, comm.cu)
struct PeerOut {
    uint32_t n = 0;
    uint32_t* ids[7];
    float* dist[7];
    uint32_t* cnt[7];
};
// batched traversal over `view` (the snapshot's own rows, or its SQ8 codes) with `ctx`'s scratch; records ctx->done
int32_t launch_search(const veles_index* ix, const IndexView& view, SearchCtx* ctx, const float* q_d, uint32_t nq, uint32_t k,
                      uint32_t ef, uint32_t* ids_d, float* dist_d, uint32_t* cnt_d, uint32_t* stats_d, cudaStream_t st,
                      const uint32_t* extra_entries_d = nullptr, const PeerOut* peers = nullptr);
// synchronises `st`, reads and clears the context's overflow flag
int32_t check_search_error_flag(SearchCtx* ctx, cudaStream_t st);
// hybrid RRF over device-resident lists (fusion.cu); only enqueues
int32_t rrf_hybrid_enqueue_d(const uint32_t* vec_ids_d, const uint32_t* vec_cnt_d, const uint32_t* txt_ids_d,
                             const uint32_t* txt_cnt_d, uint32_t nq, uint32_t in_k, float vector_weight, uint32_t k,
                             uint32_t* out_ids_d, float* out_score_d, uint32_t* out_counts_d, cudaStream_t st);
}  // namespace veles

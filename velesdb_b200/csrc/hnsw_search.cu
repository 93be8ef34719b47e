// hnsw_search.cu -- host side of the batched HNSW search: launch configuration and the C ABI entry points
// veles_search_batch[_d].  The kernel lives in hnsw_search.cuh and is instantiated per storage type in
// hnsw_search_{f32,f16,bin1,sq8a,sq8b,sq8c}.cu (separate translation units: they compile in parallel).
#include <algorithm>

#include "hnsw_search.cuh"

namespace veles {

SearchKernel search_kernel_sq8(uint32_t reg_mode, uint32_t qn, bool coop) {
    // rows up to 1 KB: per-lane async copies, query chunks in registers (eval_list_sq8); qn 0 = the TMA ring
    return qn <= 2   ? search_kernel_sq8_a(reg_mode, qn, coop)
           : qn <= 6 ? search_kernel_sq8_b(reg_mode, qn, coop)
                     : search_kernel_sq8_c(reg_mode, qn, coop);
}

// ---- host side ---------------------------------------------------------------------------------
static uint32_t env_u32(const char* name, uint32_t dflt) {
    const char* v = std::getenv(name);
    return v && *v ? (uint32_t)std::strtoul(v, nullptr, 10) : dflt;
}

constexpr size_t kMaxCtx = 16;

static int32_t harvest_flag(const veles_index* ix, SearchCtx* c) {
    // the context's previous work is complete (its event was waited on or queried): a plain copy sees the final flag
    if (!c->launched || !c->counters.p) return VELES_OK;
    uint32_t h[2] = {0, 0};
    VELES_CUDA(cudaMemcpy(h, c->counters.p, 8, cudaMemcpyDeviceToHost));
    if (h[1] != 0) {
        ix->overflowed.push_back(c->bound);
        VELES_CUDA(cudaMemset(c->counters.as<uint32_t>() + 1, 0, 4));
    }
    return VELES_OK;
}

int32_t acquire_ctx(const veles_index* ix, cudaStream_t st, bool exclusive, SearchCtx** out) {
    SearchCtx* pick = nullptr;
    for (auto& c : ix->ctxs)
        if (!c->in_use && c->launched && c->bound == st) {
            pick = c.get();
            break;
        }
    if (!pick) {
        for (auto& c : ix->ctxs) {
            if (c->in_use) continue;
            if (!c->launched || cudaEventQuery(c->done) == cudaSuccess) {
                pick = c.get();
                break;
            }
        }
        (void)cudaGetLastError();  // cudaErrorNotReady from the queries above is not an error
        if (!pick && ix->ctxs.size() < kMaxCtx) {
            ix->ctxs.emplace_back(new SearchCtx());
            pick = ix->ctxs.back().get();
            VELES_CUDA(cudaEventCreateWithFlags(&pick->done, cudaEventDisableTiming));
        }
        if (!pick) {
            for (auto& c : ix->ctxs)
                if (!c->in_use) {
                    pick = c.get();
                    break;
                }
            if (!pick) {
                set_error("all %zu search contexts of this index are checked out", ix->ctxs.size());
                return VELES_ERR_OVERFLOW;
            }
            VELES_CUDA(cudaEventSynchronize(pick->done));
        }
        if (pick->launched && pick->bound != st) VELES_TRY(harvest_flag(ix, pick));
    }
    pick->in_use = exclusive;
    *out = pick;
    return VELES_OK;
}

void release_ctx(const veles_index* ix, SearchCtx* c) {
    std::lock_guard<std::mutex> g(ix->mu);
    c->in_use = false;
}

int32_t launch_search(const veles_index* ix, const IndexView& view, SearchCtx* ctx, const float* q_d, uint32_t nq, uint32_t k,
                      uint32_t ef, uint32_t* ids_d, float* dist_d, uint32_t* cnt_d, uint32_t* stats_d, cudaStream_t st,
                      const uint32_t* extra_entries_d, const PeerOut* peers) {
    NvtxRange nvtx_range("veles::hnsw_search_kernel (NativeHnsw::search)");
    VELES_REQUIRE(ix->has_graph, "snapshot has no graph; build or load one first");
    VELES_REQUIRE(view.dtype != VELES_SQ8 || ix->dim <= 32768, "SQ8 traversal supports at most 32768 dimensions");
    VELES_REQUIRE(k >= 1 && k <= 65536, "k must be in 1..65536, got %u", k);
    VELES_REQUIRE(ef >= 1 && ef <= 16384, "ef must be in 1..16384, got %u", ef);
    if (nq == 0) return VELES_OK;
    SearchParams p;
    p.ix = view;
    const int32_t dtype = view.dtype;
    const uint32_t row_bytes = view.row_bytes;
    p.queries = q_d;
    p.nq = nq;
    p.k = k;
    p.ef = ef;
    p.out_ids = ids_d;
    p.out_dist = dist_d;
    p.out_counts = cnt_d;
    p.out_stats = stats_d;
    p.extra_entries = extra_entries_d;
    p.n_peers = peers ? std::min(peers->n, kMaxPeers) : 0u;
    for (uint32_t r = 0; r < kMaxPeers; ++r) {
        p.peer_ids[r] = r < p.n_peers ? peers->ids[r] : nullptr;
        p.peer_dist[r] = r < p.n_peers ? peers->dist[r] : nullptr;
        p.peer_cnt[r] = r < p.n_peers ? peers->cnt[r] : nullptr;
    }

    // shared-memory carve: [control words] results | todo | [distances, multi-warp only] | query | per warp: barriers + ring
    const uint32_t bar_bytes = kMaxSlots * 8;
    p.off_res = 16;
    const uint32_t res_bytes = round_up(ef * 8, 16);
    p.off_todo = p.off_res + res_bytes;
    const uint32_t todo_bytes = round_up(std::max(ix->stride0, ix->strideU) * 4, 16);
    const uint32_t q_bytes = dtype == VELES_SQ8 ? row_bytes : round_up(dtype == VELES_BIN1 ? ix->dim / 8 : ix->dim * 4, 16);
    auto carve = [&](bool coop) {
        p.off_dist = p.off_todo + todo_bytes;
        p.off_q = p.off_dist + (coop ? todo_bytes : 0u);
        p.off_ring = round_up(p.off_q + q_bytes, 128);
    };
    carve(false);
    // device limits: looked up once per device (the attribute queries cost microseconds a single-query search notices)
    static thread_local int cached_dev = -1, c_max_smem = 0, c_sms = 0, c_sm_smem = 0;
    int dev = 0;
    VELES_CUDA(cudaGetDevice(&dev));
    if (dev != cached_dev) {
        VELES_CUDA(cudaDeviceGetAttribute(&c_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        VELES_CUDA(cudaDeviceGetAttribute(&c_sms, cudaDevAttrMultiProcessorCount, dev));
        VELES_CUDA(cudaDeviceGetAttribute(&c_sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        cached_dev = dev;
    }
    const int max_smem = c_max_smem, sms = c_sms, sm_smem = c_sm_smem;
    // Quad path: four candidates per step (8 lanes each); f32 / f16 rows from 16 dimensions up (the reference's
    // wide16 regime: 32-element blocks, then its 8-wide and scalar tails).
    const bool can_quad = (dtype == VELES_BIN1 && ix->dim % 128 == 0) || dtype == VELES_SQ8 ||
                          ((dtype == VELES_F32 || dtype == VELES_F16) && ix->dim >= 16 &&
                           (ix->metric == VELES_COSINE || ix->metric == VELES_EUCLIDEAN || ix->metric == VELES_DOT));
    p.quad = (can_quad && env_u32("VELES_SEARCH_QUAD", 1) != 0) ? 1 : 0;
    // packed rows of <= 128 bytes are read directly (no ring), see eval_list_bits
    if (p.quad && dtype == VELES_BIN1 && ix->dim <= 1024 && env_u32("VELES_SEARCH_BITS_DIRECT", 1) != 0) p.quad = 2;
    p.evict_first = env_u32("VELES_SEARCH_EVICT_FIRST", 1) != 0 ? 1 : 0;
    p.peek = env_u32("VELES_SEARCH_PEEK", 1) != 0 ? 1 : 0;
    // SQ8: 16-byte chunks per lane (8 lanes per row), rounded up to an even count; 0 = rows too long, use the ring
    uint32_t sq_qn = 0;
    if (dtype == VELES_SQ8 && p.quad == 1 && env_u32("VELES_SEARCH_SQ8_ASYNC", 1) != 0) {
        const uint32_t per_lane = (row_bytes / 16 + 7) / 8;
        if (per_lane <= 8) sq_qn = (per_lane + 1) & ~1u;
    }
    // result array in registers when ef allows (2 or 8 keys per lane), else in shared memory
    const uint32_t reg_mode = env_u32("VELES_SEARCH_REG_RESULTS", 1) == 0 ? 0 : (ef <= 64 ? 2 : (ef <= 256 ? 8 : 0));

    // Warps per query.  A lone warp's instruction chain bounds a query's latency (~1 ms at 1M x 768 f32), so a
    // batch that cannot fill the GPU with one warp per query gets 8, 4 or 2 warps per query (hnsw_search_kernel
    // <.., COOP = true>): the largest count whose resident CTAs still hold the whole batch at once.
    uint32_t warps = 1, nslot = 2;
    const bool coop_ok = p.quad == 1 && reg_mode != 0 && dtype != VELES_BIN1;
    const uint32_t coop_slots = (sq_qn != 0 && 16u * row_bytes <= 16384u) ? 16u : 8u;
    auto coop_smem = [&](uint32_t w) { return p.off_ring + w * round_up(bar_bytes + coop_slots * row_bytes, 128); };
    if (coop_ok) {
        carve(true);
        const uint32_t forced = env_u32("VELES_SEARCH_WARPS", 0);
        for (uint32_t w = 8; w >= 2; w >>= 1) {
            const uint32_t sb = coop_smem(w);
            if (sb > (uint32_t)max_smem) continue;
            const uint32_t per_sm = std::min((uint32_t)sm_smem / (sb + 1024u), 32u / w);
            if (forced ? forced == w : (uint64_t)per_sm * (uint32_t)sms >= nq) {
                warps = w;
                break;
            }
        }
        if (warps == 1) carve(false);
    }
    if (warps > 1) {
        nslot = coop_slots;
    } else {
        // resident warps (queries) per SM.  Measured on B200 (profiles/): every query of a 1024-batch must be
        // resident at once (7 x 148 = 1036 slots) -- with fewer slots a second wave of queries starts late and
        // the batch time nearly doubles; 7 CTAs leave room for 2 stages of 4 rows each.
        // Batches larger than 7 x SMs prefer more resident queries with the minimal ring (two stages) over
        // deeper rings: measured on C3s (f16, 8192 queries) 7 -> 13 CTAs/SM = 173 K -> 222 K queries/s.
        const uint32_t fixed = p.off_ring + bar_bytes;
        const uint32_t min_ring = (p.quad == 1 ? 8u : 2u) * row_bytes;
        const uint32_t cmax = std::max(1u, (uint32_t)sm_smem / (fixed + min_ring + 1024u));
        const uint32_t need = (nq + (uint32_t)sms - 1) / (uint32_t)sms;
        const uint32_t auto_ctas = std::min(std::min(cmax, 16u), std::max(need, p.quad ? 7u : 8u));
        const uint32_t want_ctas = std::max(1u, env_u32("VELES_SEARCH_CTAS_PER_SM", auto_ctas));
        const uint32_t per_cta_target = (uint32_t)sm_smem / want_ctas - 1024;
        if (per_cta_target > fixed + 2 * row_bytes) nslot = (per_cta_target - fixed) / row_bytes;
        nslot = std::min(std::max(nslot, 2u), kMaxSlots);
        if (p.quad == 2) {
            // packed-bit rows: the ring holds a whole neighbour list (eval_list_bits), 64 rows = 8 KB
            nslot = std::min(round_up(std::max(ix->stride0, 8u), 4), 64u);
        } else if (p.quad) {
            nslot &= ~3u;
            if (nslot < 8) nslot = 8;  // at least two stages
            if (fixed + nslot * row_bytes > (uint32_t)max_smem) {
                p.quad = 0;
                sq_qn = 0;
                nslot = 2;
            }
        }
    }
    p.nslot = nslot;
    p.warp_bytes = round_up(bar_bytes + nslot * row_bytes, 128);
    const uint32_t smem_bytes = p.off_ring + warps * p.warp_bytes;
    VELES_REQUIRE((int)smem_bytes <= max_smem, "search needs %u bytes of shared memory per query (dim %u, ef %u); limit %d",
                  smem_bytes, ix->dim, ef, max_smem);

    const bool coop = warps > 1;
    SearchKernel kern = dtype == VELES_F32   ? search_kernel_f32(reg_mode, coop)
                        : dtype == VELES_F16 ? search_kernel_f16(reg_mode, coop)
                        : dtype == VELES_SQ8 ? search_kernel_sq8(reg_mode, sq_qn, coop)
                                             : search_kernel_bin1(reg_mode);
    const int ctas_per_sm = cached_blocks_per_sm(reinterpret_cast<const void*>(kern), (int)(32 * warps), smem_bytes);
    if (ctas_per_sm < 0) return VELES_ERR_CUDA;
    VELES_REQUIRE(ctas_per_sm >= 1, "search kernel does not fit on an SM");
    const uint32_t max_slots = (uint32_t)(ctas_per_sm * sms);
    const uint32_t grid = std::min(nq, max_slots);

    // scratch: visited bitmaps, logs, tie lists for `max_slots` resident queries
    p.vis_words = std::max((uint32_t)((ix->n + 31) / 32), 1u);
    if (ctx->slots < max_slots || ctx->vis_words != p.vis_words || !ctx->visited.p) {
        // (re)allocation: cudaFree waits for the device, so an earlier launch still using the old buffers is safe
        VELES_TRY(ctx->visited.alloc((size_t)max_slots * p.vis_words * 4));
        VELES_CUDA(cudaMemsetAsync(ctx->visited.p, 0, ctx->visited.bytes, st));
        VELES_TRY(ctx->vlog.alloc((size_t)max_slots * kLogCap * 4));
        VELES_TRY(ctx->tie.alloc((size_t)max_slots * kTieCap * 8));
        if (!ctx->counters.p) {
            VELES_TRY(ctx->counters.alloc(64));
            VELES_CUDA(cudaMemsetAsync(ctx->counters.p, 0, 64, st));
        }
        ctx->slots = max_slots;
        ctx->vis_words = p.vis_words;
    }
    p.visited = ctx->visited.as<uint32_t>();
    p.vlog = ctx->vlog.as<uint32_t>();
    p.tie = ctx->tie.as<uint64_t>();
    p.counters = ctx->counters.as<uint32_t>();
    VELES_CUDA(cudaMemsetAsync(p.counters, 0, 4, st));  // the work counter; the overflow flag [1] is sticky
    kern<<<grid, 32 * warps, smem_bytes, st>>>(p);
    count_launch();
    VELES_CUDA(cudaGetLastError());
    ctx->bound = st;
    ctx->launched = true;
    VELES_CUDA(cudaEventRecord(ctx->done, st));
    return VELES_OK;
}

static int32_t overflow_error() {
    set_error("tie list overflow (> %u equal-distance evicted candidates in one query)", kTieCap);
    return VELES_ERR_OVERFLOW;
}

int32_t check_search_error_flag(SearchCtx* ctx, cudaStream_t st) {
    uint32_t h[2] = {0, 0};
    VELES_CUDA(cudaMemcpyAsync(h, ctx->counters.p, 8, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaStreamSynchronize(st));
    if (h[1] != 0) {
        VELES_CUDA(cudaMemsetAsync(ctx->counters.as<uint32_t>() + 1, 0, 4, st));
        return overflow_error();
    }
    return VELES_OK;
}

}  // namespace veles

using namespace veles;

extern "C" {

int32_t veles_search_batch_d(const veles_index_t* idx, const float* queries_d, uint32_t nq, uint32_t k, uint32_t ef,
                             uint32_t* out_node_ids_d, float* out_raw_dist_d, uint32_t* out_counts_d,
                             uint32_t* out_stats_d, void* stream) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    VELES_REQUIRE(nq == 0 || (queries_d && out_node_ids_d && out_raw_dist_d && out_counts_d), "NULL buffer");
    std::lock_guard<std::mutex> g(idx->mu);
    SearchCtx* ctx = nullptr;
    VELES_TRY(acquire_ctx(idx, (cudaStream_t)stream, false, &ctx));
    return launch_search(idx, idx->view(), ctx, queries_d, nq, k, ef, out_node_ids_d, out_raw_dist_d, out_counts_d, out_stats_d,
                         (cudaStream_t)stream);
}

// Outcome of the `_d` searches enqueued on `stream` so far: waits for the stream, then reports (and clears) a tie-list
// overflow of any of them.  `_d` calls only enqueue, so this is how their callers learn about VELES_ERR_OVERFLOW.
int32_t veles_search_status(const veles_index_t* idx, void* stream) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    VELES_CUDA(cudaStreamSynchronize(st));
    std::lock_guard<std::mutex> g(idx->mu);
    bool bad = false;
    for (auto it = idx->overflowed.begin(); it != idx->overflowed.end();) {
        if (*it == st) {
            bad = true;
            it = idx->overflowed.erase(it);
        } else {
            ++it;
        }
    }
    for (auto& c : idx->ctxs) {
        if (!c->launched || c->bound != st || !c->counters.p) continue;
        uint32_t h[2] = {0, 0};
        VELES_CUDA(cudaMemcpy(h, c->counters.p, 8, cudaMemcpyDeviceToHost));
        if (h[1] != 0) {
            bad = true;
            VELES_CUDA(cudaMemset(c->counters.as<uint32_t>() + 1, 0, 4));
        }
    }
    return bad ? overflow_error() : VELES_OK;
}

// host-pointer search on `st` with an exclusive context: staging + launch + copies back; does not synchronise
static int32_t enqueue_host_search(const veles_index_t* idx, SearchCtx* ctx, const float* queries, uint32_t nq, uint32_t k,
                                   uint32_t ef, const uint32_t* extra_entries, uint32_t* out_node_ids, float* out_raw_dist,
                                   uint32_t* out_counts, uint32_t* out_stats, cudaStream_t st) {
    const size_t qb = (size_t)nq * idx->dim * 4, ob = (size_t)nq * k * 4;
    VELES_TRY(ctx->q_d.ensure(qb));
    VELES_TRY(ctx->ids_d.ensure(ob));
    VELES_TRY(ctx->val_d.ensure(ob));
    VELES_TRY(ctx->cnt_d.ensure((size_t)nq * 4));
    if (out_stats) VELES_TRY(ctx->stats_d.ensure((size_t)nq * 16));
    if (extra_entries) VELES_TRY(ctx->extra_d.ensure((size_t)nq * 12));
    VELES_CUDA(cudaMemcpyAsync(ctx->q_d.p, queries, qb, cudaMemcpyHostToDevice, st));
    if (extra_entries) VELES_CUDA(cudaMemcpyAsync(ctx->extra_d.p, extra_entries, (size_t)nq * 12, cudaMemcpyHostToDevice, st));
    VELES_TRY(launch_search(idx, idx->view(), ctx, ctx->q_d.as<float>(), nq, k, ef, ctx->ids_d.as<uint32_t>(), ctx->val_d.as<float>(),
                            ctx->cnt_d.as<uint32_t>(), out_stats ? ctx->stats_d.as<uint32_t>() : nullptr, st,
                            extra_entries ? ctx->extra_d.as<uint32_t>() : nullptr));
    VELES_CUDA(cudaMemcpyAsync(out_node_ids, ctx->ids_d.p, ob, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaMemcpyAsync(out_raw_dist, ctx->val_d.p, ob, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaMemcpyAsync(out_counts, ctx->cnt_d.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (out_stats) VELES_CUDA(cudaMemcpyAsync(out_stats, ctx->stats_d.p, (size_t)nq * 16, cudaMemcpyDeviceToHost, st));
    return VELES_OK;
}

static int32_t host_search(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k, uint32_t ef,
                           const uint32_t* extra_entries, uint32_t* out_node_ids, float* out_raw_dist, uint32_t* out_counts,
                           uint32_t* out_stats, cudaStream_t st) {
    SearchCtx* ctx = nullptr;
    {
        std::lock_guard<std::mutex> g(idx->mu);
        VELES_TRY(acquire_ctx(idx, st, true, &ctx));
    }
    // the context is checked out: other host threads search concurrently with their own contexts
    int32_t s = enqueue_host_search(idx, ctx, queries, nq, k, ef, extra_entries, out_node_ids, out_raw_dist, out_counts,
                                    out_stats, st);
    if (s == VELES_OK) s = check_search_error_flag(ctx, st);
    release_ctx(idx, ctx);
    return s;
}

int32_t veles_search_batch(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k, uint32_t ef,
                           uint32_t* out_node_ids, float* out_raw_dist, uint32_t* out_counts, uint32_t* out_stats,
                           void* stream) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    VELES_REQUIRE(nq == 0 || (queries && out_node_ids && out_raw_dist && out_counts), "NULL buffer");
    if (nq == 0) return VELES_OK;
    return host_search(idx, queries, nq, k, ef, nullptr, out_node_ids, out_raw_dist, out_counts, out_stats, (cudaStream_t)stream);
}

int32_t veles_search_batch_multi_entry(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k, uint32_t ef,
                                       const uint32_t* extra_entries, uint32_t* out_node_ids, float* out_raw_dist,
                                       uint32_t* out_counts, uint32_t* out_stats, void* stream) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    VELES_REQUIRE(nq == 0 || (queries && extra_entries && out_node_ids && out_raw_dist && out_counts), "NULL buffer");
    VELES_REQUIRE(ef >= 4, "multi-entry search needs ef >= 4 (every entry point enters the result set), got %u", ef);
    if (nq == 0) return VELES_OK;
    return host_search(idx, queries, nq, k, ef, extra_entries, out_node_ids, out_raw_dist, out_counts, out_stats,
                       (cudaStream_t)stream);
}

// ---- asynchronous, pipelined form of veles_search_batch ------------------------------------------------------------
// submit: the batch's H2D copy, search and D2H copies are enqueued on a stream owned by one of the index's contexts
// and the call returns; wait: blocks until that batch is complete.  Several tickets can be in flight: batch i+1's
// copies and first queries overlap batch i's tail and copies back (each launch is a persistent grid that thins out as
// its last queries finish).  Results are identical to veles_search_batch.
int32_t veles_search_submit(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k, uint32_t ef,
                            uint32_t* out_node_ids, float* out_raw_dist, uint32_t* out_counts, uint64_t* ticket) {
    VELES_REQUIRE(idx != nullptr && ticket != nullptr, "NULL argument");
    VELES_REQUIRE(nq >= 1 && queries && out_node_ids && out_raw_dist && out_counts, "NULL buffer or empty batch");
    *ticket = 0;
    SearchCtx* ctx = nullptr;
    size_t slot = 0;
    {
        std::lock_guard<std::mutex> g(idx->mu);
        // prefer a context that already owns a stream
        for (size_t i = 0; i < idx->ctxs.size() && !ctx; ++i) {
            SearchCtx* c = idx->ctxs[i].get();
            if (!c->in_use && c->own && (!c->launched || c->bound == c->own)) ctx = c;
        }
        if (ctx)
            ctx->in_use = true;
        else
            VELES_TRY(acquire_ctx(idx, nullptr, true, &ctx));
        for (size_t i = 0; i < idx->ctxs.size(); ++i)
            if (idx->ctxs[i].get() == ctx) slot = i;
        ctx->generation += 1;
    }
    int32_t s = VELES_OK;
    auto body = [&]() -> int32_t {
        if (!ctx->own) VELES_CUDA(cudaStreamCreateWithFlags(&ctx->own, cudaStreamNonBlocking));
        if (!ctx->h_flag) VELES_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_flag), 8, cudaHostAllocDefault));
        if (ctx->launched && ctx->bound != ctx->own) VELES_CUDA(cudaEventSynchronize(ctx->done));
        VELES_TRY(enqueue_host_search(idx, ctx, queries, nq, k, ef, nullptr, out_node_ids, out_raw_dist, out_counts, nullptr,
                                      ctx->own));
        VELES_CUDA(cudaMemcpyAsync(ctx->h_flag, ctx->counters.p, 8, cudaMemcpyDeviceToHost, ctx->own));
        VELES_CUDA(cudaMemsetAsync(ctx->counters.as<uint32_t>() + 1, 0, 4, ctx->own));
        VELES_CUDA(cudaEventRecord(ctx->done, ctx->own));
        return VELES_OK;
    };
    s = body();
    if (s != VELES_OK) {
        release_ctx(idx, ctx);
        return s;
    }
    *ticket = (ctx->generation << 8) | (uint64_t)(slot + 1);
    return VELES_OK;
}

int32_t veles_search_wait(const veles_index_t* idx, uint64_t ticket) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    SearchCtx* ctx = nullptr;
    {
        std::lock_guard<std::mutex> g(idx->mu);
        const size_t slot = (size_t)(ticket & 0xff);
        VELES_REQUIRE(slot >= 1 && slot <= idx->ctxs.size(), "unknown ticket");
        ctx = idx->ctxs[slot - 1].get();
        VELES_REQUIRE(ctx->in_use && ctx->generation == (ticket >> 8), "stale ticket");
    }
    int32_t s = VELES_OK;
    cudaError_t e = cudaEventSynchronize(ctx->done);
    if (e != cudaSuccess) {
        set_error("cudaEventSynchronize failed: %s", cudaGetErrorString(e));
        s = VELES_ERR_CUDA;
    } else if (ctx->h_flag[1] != 0) {
        s = overflow_error();
    }
    release_ctx(idx, ctx);
    return s;
}

}  // extern "C"

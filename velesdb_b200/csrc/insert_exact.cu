// insert_exact.cu -- the reference's *sequential* graph construction, restated on the device:
// NativeHnsw::insert (native/graph.rs:158-237) for nodes 0..n-1 in id order, with
//   random_layer          graph.rs:368-403   (levels precomputed on the host, same PRNG)
//   search_layer_single   graph.rs:405-428
//   search_layer          graph.rs:438-520   (ef_construction, on any layer)
//   select_neighbors      graph.rs:526-581   (alpha = 1.0, back-fill to max_conn)
//   add_bidirectional_connection  graph.rs:592-639  (append, or stable sort by distance and cut)
// The result is the graph HnswIndex::insert (index/hnsw/index/trait_impl.rs:10-36) would build -- the
// path on which the reference is deterministic (SURVEY finding 0.6) -- so build parity can be asserted
// id for id against the oracle (tests/test_gpu_builder.py).  One warp does the whole build: each insert
// depends on all earlier ones, so this is a latency-bound correctness path for small/medium graphs;
// bulk loads use veles_index_build_graph.
//
// Ties: the reference orders equal distances by BinaryHeap-internal order (graph.rs:516-518); this kernel
// orders them by node id.  With exact distance ties the two can pick different neighbours.
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

#include "index.hpp"

namespace veles {

struct BuildView {
    IndexView ix;          // adj0 / upper_adj are written by this kernel
    uint32_t* adj0;
    uint32_t* upper_adj;
    uint32_t* deg0;        // n
    uint32_t* degU;        // upper rows
    const uint8_t* level;  // n
    uint32_t* visited;     // ceil(n/32) words, zero on entry and on exit
    uint32_t* vlog;        // n entries
    uint64_t* tie;         // n entries
    uint32_t* state;       // [0] entry point, [1] max_layer, [2] has entry, [3] error
    uint32_t M, M0, ef_c, n;
};

__device__ __forceinline__ uint32_t ldcg_u32(const uint32_t* p) { return __ldcg(p); }

struct RowRef {
    uint32_t* ids;
    uint32_t* deg;
};
__device__ __forceinline__ RowRef row_of(const BuildView& b, uint32_t layer, uint32_t node) {
    RowRef r{nullptr, nullptr};
    if (layer == 0) {
        r.ids = b.adj0 + (size_t)node * b.ix.stride0;
        r.deg = b.deg0 + node;
    } else {
        const uint32_t ref = b.ix.upper_ref[node];
        if (ref != VELES_INVALID_ID && layer <= (ref & 15u)) {
            const size_t row = (size_t)(ref >> 4) + layer - 1;
            r.ids = b.upper_adj + row * b.ix.strideU;
            r.deg = b.degU + row;
        }
    }
    return r;
}

__device__ __forceinline__ float row_dist(const BuildView& b, const float* a, float norm_a, uint32_t node, uint32_t lane) {
    const uint8_t* row = b.ix.vecs + (size_t)node * b.ix.row_bytes;
    const float nb = b.ix.metric == VELES_COSINE ? *reinterpret_cast<const float*>(row + b.ix.norm_off) : 0.0f;
    return warp_metric(b.ix.metric, false, a, reinterpret_cast<const float*>(row), b.ix.dim, norm_a, nb, lane);
}
__device__ __forceinline__ float node_dist(const BuildView& b, uint32_t x, uint32_t y, uint32_t lane) {
    const uint8_t* rx = b.ix.vecs + (size_t)x * b.ix.row_bytes;
    const float na = b.ix.metric == VELES_COSINE ? *reinterpret_cast<const float*>(rx + b.ix.norm_off) : 0.0f;
    return row_dist(b, reinterpret_cast<const float*>(rx), na, y, lane);
}

__device__ __forceinline__ uint64_t bkey(float d, uint32_t id) { return ((uint64_t)ord_key(d) << 32) | ((uint64_t)id << 1); }
__device__ __forceinline__ float bkey_dist(uint64_t k) { return ord_unkey((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t bkey_id(uint64_t k) { return ((uint32_t)k) >> 1; }

// graph.rs:405-428
__device__ uint32_t greedy_layer(const BuildView& b, const float* q, float nq, uint32_t entry, uint32_t layer, uint32_t lane) {
    uint32_t best = entry;
    float best_dist = row_dist(b, q, nq, entry, lane);
    for (;;) {
        const RowRef r = row_of(b, layer, best);
        const uint32_t deg = r.ids ? ldcg_u32(r.deg) : 0;
        bool improved = false;
        for (uint32_t j = 0; j < deg; ++j) {  // scans the row of the node that was best when the scan started
            const uint32_t x = ldcg_u32(r.ids + j);
            const float d = row_dist(b, q, nq, x, lane);
            if (d < best_dist) {
                best = x;
                best_dist = d;
                improved = true;
            }
        }
        if (!improved) break;
    }
    return best;
}

// graph.rs:438-520 on `layer`; result = res[0..len) ascending by (dist, id) in shared memory.
__device__ uint32_t beam_layer(const BuildView& b, const float* q, float nq, uint32_t entry, uint32_t ef, uint32_t layer,
                               uint64_t* res, uint32_t lane) {
    uint32_t len = 0, logn = 0, tlen = 0, nxt = 0;
    {
        const float d0 = row_dist(b, q, nq, entry, lane);
        if (lane == 0) {
            b.visited[entry >> 5] |= 1u << (entry & 31);
            b.vlog[0] = entry;
            res[0] = bkey(d0, entry);
        }
        logn = 1;
        len = 1;
        __syncwarp();
    }
    for (;;) {
        uint32_t cnode;
        if (nxt < len) {
            const uint64_t key = res[nxt];
            cnode = bkey_id(key);
            __syncwarp();
            if (lane == 0) res[nxt] = key | 1ull;
            __syncwarp();
            uint32_t f = len;
            for (uint32_t i = nxt + 1; i < len; ++i)
                if ((res[i] & 1ull) == 0) {
                    f = i;
                    break;
                }
            nxt = f;
        } else if (tlen > 0) {
            uint64_t best = ~0ull;
            for (uint32_t i = lane; i < tlen; i += 32) best = min(best, b.tie[i]);
            best = warp_min_u64(best);
            cnode = bkey_id(best);
            const uint64_t lastv = b.tie[tlen - 1];
            __syncwarp();
            for (uint32_t i = lane; i < tlen; i += 32)
                if (b.tie[i] == best) b.tie[i] = lastv;
            --tlen;
            __syncwarp();
        } else {
            break;
        }
        const RowRef r = row_of(b, layer, cnode);
        const uint32_t deg = r.ids ? ldcg_u32(r.deg) : 0;
        for (uint32_t j = 0; j < deg; ++j) {
            const uint32_t x = ldcg_u32(r.ids + j);
            const uint32_t bit = 1u << (x & 31);
            const uint32_t w = __ldcg(&b.visited[x >> 5]);  // uniform read; this warp is the only writer
            if (w & bit) continue;
            __syncwarp();
            if (lane == 0) {
                b.visited[x >> 5] = w | bit;
                b.vlog[logn] = x;
            }
            ++logn;
            __syncwarp();
            const float d = row_dist(b, q, nq, x, lane);
            const float worst = bkey_dist(res[len - 1]);
            if (d < worst || len < ef) {
                const uint64_t key = bkey(d, x);
                const uint32_t pos = lower_bound_warp(res, len, key, lane);
                if (len < ef) {
                    insert_at(res, pos, len + 1, key, lane);
                    ++len;
                    if (pos <= nxt) nxt = pos;
                } else {
                    const uint64_t ev = res[len - 1];
                    __syncwarp();
                    insert_at(res, pos, len, key, lane);
                    const float nworst = bkey_dist(res[len - 1]);
                    if (tlen > 0) {  // drop ties that are now farther than the worst result
                        uint32_t wn = 0;
                        for (uint32_t base = 0; base < tlen; base += 32) {
                            const uint32_t i = base + lane;
                            uint64_t v = 0;
                            bool keep = false;
                            if (i < tlen) {
                                v = b.tie[i];
                                keep = !(bkey_dist(v) > nworst);
                            }
                            const uint32_t msk = __ballot_sync(FULL_MASK, keep);
                            __syncwarp();
                            if (keep) b.tie[wn + __popc(msk & ((1u << lane) - 1u))] = v;
                            wn += __popc(msk);
                            __syncwarp();
                        }
                        tlen = wn;
                    }
                    if ((ev & 1ull) == 0 && !(bkey_dist(ev) > nworst)) {
                        if (lane == 0) b.tie[tlen] = ev;
                        ++tlen;
                        __syncwarp();
                    }
                    if (pos <= nxt || nxt >= len) nxt = pos;
                }
            }
        }
    }
    for (uint32_t i = lane; i < logn; i += 32) b.visited[b.vlog[i] >> 5] = 0u;
    __syncwarp();
    return len;
}

__global__ void __launch_bounds__(32) insert_exact_kernel(BuildView b) {
    extern __shared__ __align__(16) uint8_t ie_smem[];
    const uint32_t lane = threadIdx.x;
    const uint32_t dim = b.ix.dim;
    float* q = reinterpret_cast<float*>(ie_smem);                                         // dim
    uint64_t* res = reinterpret_cast<uint64_t*>(ie_smem + ((dim * 4 + 15) & ~15u));       // ef_c + 1
    uint64_t* prune = res + b.ef_c + 1;                                                   // M0 + 1 (dist, position)
    uint32_t* sel = reinterpret_cast<uint32_t*>(prune + b.M0 + 1);                        // M0
    uint32_t* tmp_ids = sel + b.M0;                                                       // M0
    uint32_t ep = 0, max_layer = 0;
    bool has_ep = false;
    for (uint32_t id = 0; id < b.n; ++id) {
        const uint32_t node_layer = b.level[id];
        if (has_ep) {
            const float* qg = reinterpret_cast<const float*>(b.ix.vecs + (size_t)id * b.ix.row_bytes);
            for (uint32_t i = lane; i < dim; i += 32) q[i] = qg[i];
            __syncwarp();
            const float nq = b.ix.metric == VELES_COSINE
                                 ? *reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(qg) + b.ix.norm_off)
                                 : 0.0f;
            uint32_t cur = ep;
            for (uint32_t l = max_layer; l >= node_layer + 1; --l) cur = greedy_layer(b, q, nq, cur, l, lane);
            for (int32_t l = (int32_t)node_layer; l >= 0; --l) {
                const uint32_t wlen = beam_layer(b, q, nq, cur, b.ef_c, (uint32_t)l, res, lane);
                const uint32_t w0 = bkey_id(res[0]);  // neighbors[0].0, the next layer's entry (graph.rs:220-222)
                const uint32_t maxc = l == 0 ? b.M0 : b.M;
                // ---- select_neighbors (graph.rs:526-581) over W = res[0..wlen) ----
                uint32_t nsel = 0;
                if (wlen <= maxc) {
                    for (uint32_t i = lane; i < wlen; i += 32) sel[i] = bkey_id(res[i]);
                    nsel = wlen;
                    __syncwarp();
                } else {
                    for (uint32_t i = 0; i < wlen && nsel < maxc; ++i) {
                        const uint32_t cnode = bkey_id(res[i]);
                        const float cdist = bkey_dist(res[i]);
                        bool diverse = true;
                        for (uint32_t s = 0; s < nsel; ++s) {
                            const float ds = node_dist(b, cnode, sel[s], lane);
                            if (!(cdist <= ds)) {  // alpha = 1.0: alpha * d(q,c) <= d(c,s)
                                diverse = false;
                                break;
                            }
                        }
                        if (diverse || nsel == 0) {
                            if (lane == 0) sel[nsel] = cnode;
                            ++nsel;
                            __syncwarp();
                        }
                    }
                    if (nsel < maxc) {  // back-fill with the closest not yet selected
                        for (uint32_t i = 0; i < wlen && nsel < maxc; ++i) {
                            const uint32_t cnode = bkey_id(res[i]);
                            bool present = false;
                            for (uint32_t s = lane; s < nsel; s += 32) present |= sel[s] == cnode;
                            if (!__any_sync(FULL_MASK, present)) {
                                if (lane == 0) sel[nsel] = cnode;
                                ++nsel;
                                __syncwarp();
                            }
                        }
                    }
                }
                // ---- set_neighbors(id, selected) ----
                {
                    const RowRef r = row_of(b, (uint32_t)l, id);
                    if (r.ids) {
                        for (uint32_t i = lane; i < nsel; i += 32) r.ids[i] = sel[i];
                        if (lane == 0) *r.deg = nsel;
                    } else if (lane == 0) {
                        b.state[3] = 1;  // missing row: host-side row planning bug
                    }
                    __threadfence_block();
                    __syncwarp();
                }
                // ---- add_bidirectional_connection for each selected neighbour, in order (graph.rs:592-639) ----
                for (uint32_t si = 0; si < nsel; ++si) {
                    const uint32_t nbr = sel[si];
                    const RowRef r = row_of(b, (uint32_t)l, nbr);
                    if (!r.ids) {
                        if (lane == 0) b.state[3] = 2;
                        continue;
                    }
                    const uint32_t deg = ldcg_u32(r.deg);
                    if (deg < maxc) {
                        if (lane == 0) {
                            r.ids[deg] = id;
                            *r.deg = deg + 1;
                        }
                    } else {
                        // distances from nbr to each of its neighbours and to the new node; stable sort
                        // (key = distance, then position in [current..., new]); keep the first maxc
                        for (uint32_t j = 0; j <= deg; ++j) {
                            const uint32_t x = j < deg ? ldcg_u32(r.ids + j) : id;
                            const float d = node_dist(b, nbr, x, lane);
                            if (lane == 0) prune[j] = ((uint64_t)ord_key(d) << 32) | j;
                        }
                        __syncwarp();
                        // (deg + 1) <= 257 entries: rank by counting
                        for (uint32_t j = lane; j <= deg; j += 32) {
                            const uint64_t kj = prune[j];
                            uint32_t rank = 0;
                            for (uint32_t t = 0; t <= deg; ++t) rank += prune[t] < kj;
                            if (rank < maxc) {
                                const uint32_t src = (uint32_t)kj;
                                tmp_ids[rank] = src < deg ? ldcg_u32(r.ids + src) : id;
                            }
                        }
                        __syncwarp();
                        for (uint32_t j = lane; j < maxc; j += 32) r.ids[j] = tmp_ids[j];
                        // degree stays maxc
                    }
                    __threadfence_block();
                    __syncwarp();
                }
                if (wlen > 0) cur = w0;
            }
        } else {
            has_ep = true;
            ep = id;
        }
        if (node_layer > max_layer) {
            max_layer = node_layer;
            ep = id;
        }
    }
    if (lane == 0) {
        b.state[0] = ep;
        b.state[1] = max_layer;
        b.state[2] = has_ep ? 1u : 0u;
    }
}

// graph.rs:368-403 (same PRNG as builder.cu)
static void exact_levels(uint64_t n, uint32_t M, std::vector<uint8_t>& level) {
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
        level[i] = (uint8_t)std::min(15.0, std::max(0.0, std::floor(-std::log(u) * level_mult)));
    }
}

}  // namespace veles

using namespace veles;

extern "C" int32_t veles_index_build_graph_exact(veles_index_t* ix, uint32_t M, uint32_t ef_construction, void* stream) {
    NvtxRange nvtx_range("veles::build_graph_exact (NativeHnsw::insert, sequential)");
    VELES_REQUIRE(ix != nullptr, "index is NULL");
    VELES_REQUIRE(M >= 2 && M <= 128, "M must be in 2..128, got %u", M);
    VELES_REQUIRE(ef_construction >= 1 && ef_construction <= 4096, "ef_construction must be in 1..4096");
    VELES_REQUIRE(ix->dtype == VELES_F32, "the exact sequential builder needs f32 storage");
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> g(ix->mu);
    const uint64_t n = ix->n;
    const uint32_t M0 = 2 * M;
    ix->M = M;
    ix->M0 = M0;
    ix->ef_construction = ef_construction;
    ix->stride0 = round_up(M0, 32);
    ix->strideU = round_up(M, 32);
    // levels; rows each node needs: its own level, and -- for a node that is the entry point when a node of
    // a higher level arrives -- every layer up to that level (graph.rs:195-218 links the old entry point on
    // the new top layers)
    std::vector<uint8_t> level;
    exact_levels(n, M, level);
    std::vector<uint8_t> top(level);
    {
        uint32_t max_layer = 0;
        uint64_t ep = 0;
        bool has = false;
        for (uint64_t i = 0; i < n; ++i) {
            if (has && level[i] > max_layer) top[ep] = std::max<uint8_t>(top[ep], level[i]);
            if (!has) {
                has = true;
                ep = i;
            }
            if (level[i] > max_layer) {
                max_layer = level[i];
                ep = i;
            }
        }
    }
    std::vector<uint32_t> h_ref(std::max<uint64_t>(n, 1), VELES_INVALID_ID);
    uint64_t rows = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (top[i] > 0) {
            h_ref[i] = (uint32_t)(rows << 4) | top[i];
            rows += top[i];
        }
    ix->upper_rows = rows;
    VELES_TRY(ix->adj0.alloc(std::max<size_t>((size_t)n * ix->stride0 * 4, 16)));
    VELES_TRY(ix->upper_ref.alloc(std::max<size_t>((size_t)n * 4, 16)));
    VELES_TRY(ix->upper_adj.alloc(std::max<size_t>((size_t)rows * ix->strideU * 4, 16)));
    VELES_CUDA(cudaMemsetAsync(ix->adj0.p, 0xff, ix->adj0.bytes, st));
    VELES_CUDA(cudaMemsetAsync(ix->upper_adj.p, 0xff, ix->upper_adj.bytes, st));
    if (n) VELES_CUDA(cudaMemcpyAsync(ix->upper_ref.p, h_ref.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    ix->has_graph = true;
    ix->has_entry = false;
    ix->entry = 0;
    ix->max_layer = 0;
    ix->num_layers = 1;
    if (n == 0) return VELES_OK;
    DevBuf deg0, degU, lvl, vis, vlog, tie, state;
    VELES_TRY(deg0.alloc((size_t)n * 4));
    VELES_TRY(degU.alloc(std::max<size_t>((size_t)rows * 4, 16)));
    VELES_TRY(lvl.alloc((size_t)n));
    VELES_TRY(vis.alloc(((size_t)n + 31) / 32 * 4));
    VELES_TRY(vlog.alloc((size_t)n * 4));
    VELES_TRY(tie.alloc((size_t)n * 8));
    VELES_TRY(state.alloc(16));
    VELES_CUDA(cudaMemsetAsync(deg0.p, 0, deg0.bytes, st));
    VELES_CUDA(cudaMemsetAsync(degU.p, 0, degU.bytes, st));
    VELES_CUDA(cudaMemsetAsync(vis.p, 0, vis.bytes, st));
    VELES_CUDA(cudaMemsetAsync(state.p, 0, 16, st));
    VELES_CUDA(cudaMemcpyAsync(lvl.p, level.data(), (size_t)n, cudaMemcpyHostToDevice, st));
    BuildView b;
    ix->has_graph = true;
    b.ix = ix->view();
    b.adj0 = ix->adj0.as<uint32_t>();
    b.upper_adj = ix->upper_adj.as<uint32_t>();
    b.deg0 = deg0.as<uint32_t>();
    b.degU = degU.as<uint32_t>();
    b.level = lvl.as<uint8_t>();
    b.visited = vis.as<uint32_t>();
    b.vlog = vlog.as<uint32_t>();
    b.tie = tie.as<uint64_t>();
    b.state = state.as<uint32_t>();
    b.M = M;
    b.M0 = M0;
    b.ef_c = ef_construction;
    b.n = (uint32_t)n;
    const size_t smem = (((size_t)ix->dim * 4 + 15) & ~(size_t)15) + ((size_t)ef_construction + 1) * 8 + ((size_t)M0 + 1) * 8 +
                        (size_t)M0 * 4 * 2 + 64;
    VELES_CUDA(cudaFuncSetAttribute(insert_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    insert_exact_kernel<<<1, 32, smem, st>>>(b);
    count_launch();
    VELES_CUDA(cudaGetLastError());
    uint32_t h[4] = {0, 0, 0, 0};
    VELES_CUDA(cudaMemcpyAsync(h, state.p, 16, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaStreamSynchronize(st));
    if (h[3] != 0) {
        set_error("exact builder: internal row planning error %u", h[3]);
        return VELES_ERR_OVERFLOW;
    }
    ix->entry = h[0];
    ix->max_layer = h[1];
    ix->num_layers = h[1] + 1;
    ix->has_entry = h[2] != 0;
    return VELES_OK;
}

// builder.cu -- bulk HNSW graph construction on the GPU by block insertion.
//
// Role: HnswIndex::insert_batch_parallel (index/hnsw/index/batch.rs:82-108) -> NativeHnsw::parallel_insert
// (native/backend_adapter.rs:120-122): many NativeHnsw::insert calls (native/graph.rs:158-237) racing on one graph.
// The reference's result is order dependent there (SURVEY finding 0.6), so there is no reference graph to equal;
// what is kept is the reference's *procedure*, applied to blocks of nodes in id order:
//
//   for each block [b0, b1) of new nodes (the graph holds nodes < b0):
//     1. upper layers   inc_upper_search_kernel: search_layer_single (graph.rs:405-428) down to the node's level,
//                       then search_layer (graph.rs:438-520) with ef_construction on each of its layers >= 1
//     2. layer 0        the production search kernel (hnsw_search.cuh) with k = ef = ef_construction, one query per
//                       new node -- greedy descent + layer-0 beam on the graph built so far
//     3. link           inc_link_kernel: select_neighbors (graph.rs:526-581: alpha = 1 heuristic + back-fill) on each
//                       candidate list, the node's own row, and add_bidirectional_connection (graph.rs:592-639) for
//                       every selected neighbour under a per-row spin lock: a full row keeps the max_conn closest
//
// Nodes of one block do not see each other (they all search the graph as it was at b0); blocks grow with the graph
// (at most 1/4 of it, capped), so that blindness is bounded and is repaired by later blocks' reverse links.  The
// first block is linked from exact all-pairs candidate lists (inc_seed_kernel).  O(N log N) distance evaluations,
// no library calls: every kernel here and the search kernel are hand written.
//
// What equals the reference exactly: every node's level and the final entry point (xorshift64 PRNG of
// graph.rs:368-403 drawn in node-id order), M0 = 2*M, degree bounds, the neighbour selection rule, the reverse-link
// pruning rule (closest max_conn; equal distances by node id where the reference has list position), file format.
// What does not: which nodes a new node sees (block granularity), hence the neighbour sets -- judged by recall
// against exact brute force and against an oracle-built graph (tests/test_gpu_builder.py).
// Rows are kept sorted by (distance, id), so the built graph does not depend on warp scheduling: the final row is
// the max_conn smallest keys of everything ever offered to it, whatever the arrival order.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>

#include "hnsw_search.cuh"

namespace veles {

constexpr uint32_t kBuildTieCap = 4096;   // per-slot tie list of the upper-layer beams
constexpr uint32_t kBuildLogCap = 8192;   // per-slot visited log of the upper-layer beams
constexpr uint32_t kLinkWarps = 4;

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

// the graph under construction: the snapshot's arrays, writable, plus the build-only side arrays
struct IncView {
    IndexView ix;
    uint32_t* adj0;
    float* adj0_d;      // distance of each layer-0 link to its row's owner
    uint32_t* lock0;    // one spin lock per layer-0 row
    uint32_t* upper_adj;
    float* upper_d;
    uint32_t* lockU;    // one per upper row
    const uint8_t* level;
    uint32_t M, M0, ef_c;
    uint32_t cur_max_layer, entry;  // as of the block's start
};

// ---- distances between stored rows (SimdDistance::distance, native/distance.rs:75-85) -------------------------
template <typename TA, typename TB>
__device__ __forceinline__ void threshold_counts(const TA* __restrict__ a, const TB* __restrict__ b, uint32_t dim,
                                                 uint32_t lane, uint32_t& diff, uint32_t& inter, uint32_t& uni) {
    uint32_t d = 0, in = 0, un = 0;
    for (uint32_t i = lane; i < dim; i += 32) {
        const bool x = load_elem(a, i) > 0.5f, y = load_elem(b, i) > 0.5f;
        d += (x != y);
        in += (x && y);
        un += (x || y);
    }
    diff = __reduce_add_sync(FULL_MASK, d);
    inter = __reduce_add_sync(FULL_MASK, in);
    uni = __reduce_add_sync(FULL_MASK, un);
}

template <typename T>
__device__ __forceinline__ float pair_metric(const IndexView& ix, const T* a, const T* b, float na, float nb, uint32_t lane) {
    switch (ix.metric) {
        case VELES_COSINE: {
            const float dot = warp_tree_reduce<0>(a, b, ix.dim, lane);
            return __fsub_rn(1.0f, cosine_from_parts(dot, na, nb));
        }
        case VELES_EUCLIDEAN: return __fsqrt_rn(warp_tree_reduce<1>(a, b, ix.dim, lane));
        case VELES_DOT: return -warp_tree_reduce<0>(a, b, ix.dim, lane);
        case VELES_HAMMING: {
            uint32_t d, in, un;
            threshold_counts(a, b, ix.dim, lane, d, in, un);
            return (float)d;
        }
        default: {
            uint32_t d, in, un;
            threshold_counts(a, b, ix.dim, lane, d, in, un);
            const float j = un == 0 ? 1.0f : __fdiv_rn((float)in, (float)un);
            return __fsub_rn(1.0f, j);
        }
    }
}

// in-graph distance between nodes x and y; all lanes return the same value.  Same accumulation tree as the search
// kernel (common.cuh), so d(x, y) here has the bits of the search's d(query = row x, y).
__device__ __forceinline__ float node_pair_dist(const IndexView& ix, uint32_t x, uint32_t y, uint32_t lane) {
    const uint8_t* rx = ix.vecs + (size_t)x * ix.row_bytes;
    const uint8_t* ry = ix.vecs + (size_t)y * ix.row_bytes;
    if (ix.dtype == VELES_BIN1) {
        const uint32_t words = ix.dim >> 5;
        const uint32_t* a = reinterpret_cast<const uint32_t*>(rx);
        const uint32_t* b = reinterpret_cast<const uint32_t*>(ry);
        uint32_t d = 0;
        for (uint32_t i = lane; i < words; i += 32) d += __popc(a[i] ^ b[i]);
        return (float)__reduce_add_sync(FULL_MASK, d);
    }
    float na = 0.0f, nb = 0.0f;
    if (ix.metric == VELES_COSINE) {
        na = *reinterpret_cast<const float*>(rx + ix.norm_off);
        nb = *reinterpret_cast<const float*>(ry + ix.norm_off);
    }
    if (ix.dtype == VELES_F32)
        return pair_metric(ix, reinterpret_cast<const float*>(rx), reinterpret_cast<const float*>(ry), na, nb, lane);
    return pair_metric(ix, reinterpret_cast<const __half*>(rx), reinterpret_cast<const __half*>(ry), na, nb, lane);
}

__device__ __forceinline__ uint64_t ckey(float d, uint32_t id) { return ((uint64_t)ord_key(d) << 32) | id; }

// ---- new nodes' rows as f32 queries for the search kernel ------------------------------------------------------
__global__ void rows_to_queries_kernel(IndexView ix, uint64_t first, uint32_t count, float* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = gw; r < count; r += nwarps) {
        const uint8_t* row = ix.vecs + (first + r) * ix.row_bytes;
        float* o = out + r * ix.dim;
        for (uint32_t i = lane; i < ix.dim; i += 32) {
            float v;
            if (ix.dtype == VELES_F32)
                v = reinterpret_cast<const float*>(row)[i];
            else if (ix.dtype == VELES_F16)
                v = __half2float(reinterpret_cast<const __half*>(row)[i]);
            else
                v = ((reinterpret_cast<const uint32_t*>(row)[i >> 5] >> (i & 31)) & 1u) ? 1.0f : 0.0f;
            o[i] = v;
        }
    }
}

// ---- candidate lists of the first block: exact, all pairs ------------------------------------------------------
// One warp per (node, layer) task: the ef_c closest other seed nodes that exist on `layer`.
struct TaskList {
    const uint32_t* node;   // null: node = first_node + task, layer 0
    const uint8_t* layer;   // null: layer 0
    uint32_t first_node, n_tasks;
    uint32_t* cand_ids;     // n_tasks x cand_stride, ascending by (distance, id)
    float* cand_dist;
    uint32_t* cand_cnt;
    uint32_t cand_stride;
};

__global__ void __launch_bounds__(32) inc_seed_kernel(IncView b, TaskList t, uint32_t n_seed) {
    extern __shared__ __align__(16) uint64_t seed_res[];
    const uint32_t lane = threadIdx.x;
    for (uint32_t task = blockIdx.x; task < t.n_tasks; task += gridDim.x) {
        const uint32_t node = t.node ? t.node[task] : t.first_node + task;
        const uint32_t layer = t.layer ? t.layer[task] : 0u;
        const uint32_t cap = min(b.ef_c, t.cand_stride);
        uint32_t len = 0;
        for (uint32_t j = 0; j < n_seed; ++j) {
            if (j == node || b.level[j] < layer) continue;
            const float d = node_pair_dist(b.ix, node, j, lane);
            const uint64_t key = ckey(d, j);
            if (len < cap) {
                const uint32_t pos = lower_bound_warp(seed_res, len, key, lane);
                insert_at(seed_res, pos, len + 1, key, lane);
                ++len;
            } else if (key < seed_res[cap - 1]) {
                const uint32_t pos = lower_bound_warp(seed_res, len, key, lane);
                insert_at(seed_res, pos, len, key, lane);
            }
            __syncwarp();
        }
        for (uint32_t i = lane; i < len; i += 32) {
            const uint64_t k = seed_res[i];
            t.cand_ids[(size_t)task * t.cand_stride + i] = (uint32_t)k;
            t.cand_dist[(size_t)task * t.cand_stride + i] = ord_unkey((uint32_t)(k >> 32));
        }
        if (lane == 0) t.cand_cnt[task] = len;
        __syncwarp();
    }
}

// ---- upper layers: descent + beams of the block's nodes with level >= 1 --------------------------------------------
struct UpperJobs {
    const uint32_t* node;    // per job: the new node
    const uint8_t* top;      // per job: min(level, current max layer) >= 1
    const uint32_t* task0;   // per job: task index of layer `top`; layer l is task0 + (top - l)
    uint32_t n_jobs;
    uint32_t* visited;       // slots x vis_words, bit = upper row rank (upper_ref >> 4)
    uint32_t* vlog;          // slots x kBuildLogCap
    uint64_t* tie;           // slots x kBuildTieCap
    uint32_t vis_words;
    uint32_t* error;         // [0] tie overflow
};

__device__ __forceinline__ const uint32_t* upper_row(const IncView& b, uint32_t layer, uint32_t node) {
    const uint32_t ref = b.ix.upper_ref[node];
    if (ref == VELES_INVALID_ID || layer > (ref & 15u)) return nullptr;
    return b.upper_adj + ((size_t)(ref >> 4) + layer - 1) * b.ix.strideU;
}

__device__ __forceinline__ uint64_t bkey(float d, uint32_t id) { return ((uint64_t)ord_key(d) << 32) | ((uint64_t)id << 1); }
__device__ __forceinline__ float bkey_dist(uint64_t k) { return ord_unkey((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t bkey_id(uint64_t k) { return ((uint32_t)k) >> 1; }

// search_layer_single, graph.rs:405-428, for the stored row of `node` as the query
__device__ uint32_t inc_greedy(const IncView& b, uint32_t node, uint32_t entry, uint32_t layer, uint32_t lane) {
    uint32_t best = entry;
    float best_dist = node_pair_dist(b.ix, node, entry, lane);
    for (;;) {
        const uint32_t* row = upper_row(b, layer, best);
        bool improved = false;
        if (row) {
            for (uint32_t j = 0; j < b.ix.strideU; ++j) {  // the row of the node that was best when the scan started
                const uint32_t x = __ldcg(row + j);
                if (x == VELES_INVALID_ID) break;
                const float d = node_pair_dist(b.ix, node, x, lane);
                if (d < best_dist) {
                    best = x;
                    best_dist = d;
                    improved = true;
                }
            }
        }
        if (!improved) break;
    }
    return best;
}

// search_layer, graph.rs:438-520, on an upper layer; result = res[0..len) ascending by (dist, id).  Same
// representation as the search kernel: one sorted array with an expanded bit, plus the tie list (hnsw_search.cuh).
__device__ uint32_t inc_beam(const IncView& b, const UpperJobs& u, uint32_t slot, uint32_t node, uint32_t entry, uint32_t ef,
                             uint32_t layer, uint64_t* res, uint32_t lane) {
    uint32_t* vis = u.visited + (size_t)slot * u.vis_words;
    uint32_t* vlog = u.vlog + (size_t)slot * kBuildLogCap;
    uint64_t* tie = u.tie + (size_t)slot * kBuildTieCap;
    uint32_t len = 0, logn = 0, tlen = 0, nxt = 0;
    {
        const float d0 = node_pair_dist(b.ix, node, entry, lane);
        const uint32_t vb = b.ix.upper_ref[entry] >> 4;
        if (lane == 0) {
            vis[vb >> 5] |= 1u << (vb & 31);
            vlog[0] = vb;
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
            for (uint32_t i = lane; i < tlen; i += 32) best = min(best, tie[i]);
            best = warp_min_u64(best);
            cnode = bkey_id(best);
            const uint64_t lastv = tie[tlen - 1];
            __syncwarp();
            for (uint32_t i = lane; i < tlen; i += 32)
                if (tie[i] == best) tie[i] = lastv;
            --tlen;
            __syncwarp();
        } else {
            break;
        }
        const uint32_t* row = upper_row(b, layer, cnode);
        if (!row) continue;
        for (uint32_t j = 0; j < b.ix.strideU; ++j) {
            const uint32_t x = __ldcg(row + j);
            if (x == VELES_INVALID_ID) break;
            const uint32_t vb = b.ix.upper_ref[x] >> 4;
            const uint32_t bit = 1u << (vb & 31);
            const uint32_t w = __ldcg(&vis[vb >> 5]);  // uniform read; this warp is the slot's only writer
            if (w & bit) continue;
            __syncwarp();
            if (lane == 0) {
                vis[vb >> 5] = w | bit;
                if (logn < kBuildLogCap) vlog[logn] = vb;
            }
            ++logn;
            __syncwarp();
            const float d = node_pair_dist(b.ix, node, x, lane);
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
                    if (tlen > 0) {  // ties that are now farther than the worst result can only end the loop
                        uint32_t wn = 0;
                        for (uint32_t base = 0; base < tlen; base += 32) {
                            const uint32_t i = base + lane;
                            uint64_t v = 0;
                            bool keep = false;
                            if (i < tlen) {
                                v = tie[i];
                                keep = !(bkey_dist(v) > nworst);
                            }
                            const uint32_t msk = __ballot_sync(FULL_MASK, keep);
                            __syncwarp();
                            if (keep) tie[wn + __popc(msk & ((1u << lane) - 1u))] = v;
                            wn += __popc(msk);
                            __syncwarp();
                        }
                        tlen = wn;
                    }
                    if ((ev & 1ull) == 0 && !(bkey_dist(ev) > nworst)) {
                        if (tlen < kBuildTieCap) {
                            if (lane == 0) tie[tlen] = ev;
                            ++tlen;
                        } else if (lane == 0) {
                            atomicExch(&u.error[0], 1u);
                        }
                        __syncwarp();
                    }
                    if (pos <= nxt || nxt >= len) nxt = pos;
                }
            }
        }
    }
    if (logn <= kBuildLogCap) {
        for (uint32_t i = lane; i < logn; i += 32) vis[vlog[i] >> 5] = 0u;
    } else {
        for (uint32_t i = lane; i < u.vis_words; i += 32) vis[i] = 0u;
    }
    __syncwarp();
    return len;
}

__global__ void __launch_bounds__(32) inc_upper_search_kernel(IncView b, UpperJobs u, TaskList t) {
    extern __shared__ __align__(16) uint64_t up_res[];  // ef_c + 1 keys
    const uint32_t lane = threadIdx.x;
    for (uint32_t job = blockIdx.x; job < u.n_jobs; job += gridDim.x) {
        const uint32_t node = u.node[job];
        const uint32_t top = u.top[job];
        uint32_t cur = b.entry;
        for (uint32_t l = b.cur_max_layer; l > top; --l) cur = inc_greedy(b, node, cur, l, lane);  // graph.rs:190-193
        for (uint32_t l = top; l >= 1; --l) {                                                        // graph.rs:196-223
            const uint32_t wlen = inc_beam(b, u, blockIdx.x, node, cur, b.ef_c, l, up_res, lane);
            const uint32_t task = u.task0[job] + (top - l);
            for (uint32_t i = lane; i < wlen; i += 32) {
                const uint64_t k = up_res[i];
                t.cand_ids[(size_t)task * t.cand_stride + i] = bkey_id(k);
                t.cand_dist[(size_t)task * t.cand_stride + i] = bkey_dist(k);
            }
            if (lane == 0) t.cand_cnt[task] = wlen;
            if (wlen > 0) cur = bkey_id(up_res[0]);
            __syncwarp();
        }
    }
}

// ---- link: select_neighbors + own row + reverse links --------------------------------------------------------------
struct RowPtr {
    uint32_t* ids;
    float* dist;
    uint32_t* lock;
    uint32_t stride, maxc;
};
__device__ __forceinline__ RowPtr link_row(const IncView& b, uint32_t layer, uint32_t node) {
    RowPtr r{nullptr, nullptr, nullptr, 0, 0};
    if (layer == 0) {
        r.ids = b.adj0 + (size_t)node * b.ix.stride0;
        r.dist = b.adj0_d + (size_t)node * b.ix.stride0;
        r.lock = b.lock0 + node;
        r.stride = b.ix.stride0;
        r.maxc = b.M0;
    } else {
        const uint32_t ref = b.ix.upper_ref[node];
        if (ref != VELES_INVALID_ID && layer <= (ref & 15u)) {
            const size_t row = (size_t)(ref >> 4) + layer - 1;
            r.ids = b.upper_adj + row * b.ix.strideU;
            r.dist = b.upper_d + row * b.ix.strideU;
            r.lock = b.lockU + row;
            r.stride = b.ix.strideU;
            r.maxc = b.M;
        }
    }
    return r;
}

// add_bidirectional_connection's list update (graph.rs:600-638) for one row, under the row's lock: offer (new_id, d);
// the row keeps its max_conn smallest (distance, id) keys.  An existing entry keeps its place relative to the others;
// the new one goes in front of the first larger key; when the row is full the largest key is dropped (possibly the
// new one).  rid / rd: per-warp shared-memory staging of `stride` entries.
__device__ void locked_offer(const RowPtr& r, uint32_t new_id, float d, uint32_t* rid, float* rd, uint32_t lane) {
    if (lane == 0) {
        while (atomicCAS(r.lock, 0u, 1u) != 0u) __nanosleep(64);
        __threadfence();
    }
    __syncwarp();
    for (uint32_t i = lane; i < r.stride; i += 32) {
        rid[i] = __ldcg(r.ids + i);
        rd[i] = __ldcg(r.dist + i);
    }
    __syncwarp();
    const uint64_t key_new = ckey(d, new_id);
    uint32_t deg = 0;
    bool dup = false;
    uint64_t mx_key = 0;
    uint32_t mx_idx = VELES_INVALID_ID;
    for (uint32_t base = 0; base < r.stride; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < r.stride && rid[i] != VELES_INVALID_ID;
        deg += __popc(__ballot_sync(FULL_MASK, valid));
        dup |= __any_sync(FULL_MASK, valid && rid[i] == new_id) != 0;
        if (valid) {
            const uint64_t k = ckey(rd[i], rid[i]);
            if (mx_idx == VELES_INVALID_ID || k > mx_key) {
                mx_key = k;
                mx_idx = i;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint64_t ok = __shfl_xor_sync(FULL_MASK, mx_key, o);
        const uint32_t oi = __shfl_xor_sync(FULL_MASK, mx_idx, o);
        if (oi != VELES_INVALID_ID && (mx_idx == VELES_INVALID_ID || ok > mx_key)) {
            mx_key = ok;
            mx_idx = oi;
        }
    }
    const bool full = deg >= r.maxc;
    if (!dup && !(full && key_new > mx_key)) {
        const uint32_t drop = full ? mx_idx : VELES_INVALID_ID;
        // position of the new entry among the kept ones: in front of the first kept entry with a larger key
        uint32_t pos = 0;
        for (uint32_t base = 0; base < r.stride; base += 32) {
            const uint32_t i = base + lane;
            const bool kept = i < r.stride && rid[i] != VELES_INVALID_ID && i != drop;
            pos += __popc(__ballot_sync(FULL_MASK, kept && ckey(rd[i], rid[i]) < key_new));
        }
        for (uint32_t i = lane; i < r.stride; i += 32) {
            if (rid[i] == VELES_INVALID_ID || i == drop) continue;
            const uint32_t i1 = i - ((drop != VELES_INVALID_ID && i > drop) ? 1u : 0u);
            const uint32_t ni = i1 + (i1 >= pos ? 1u : 0u);
            if (ni != i) {
                __stcg(r.ids + ni, rid[i]);
                __stcg(r.dist + ni, rd[i]);
            }
        }
        if (lane == 0) {
            __stcg(r.ids + pos, new_id);
            __stcg(r.dist + pos, d);
        }
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(r.lock, 0u);
    __syncwarp();
}

// One warp per task (new node, layer).  `seed`: the task's own row may be receiving reverse links from other tasks of
// the same launch (first block: everybody is new), so its own links go through locked_offer as well.
__global__ void __launch_bounds__(kLinkWarps * 32) inc_link_kernel(IncView b, TaskList t, uint32_t seed, uint32_t smem_stride,
                                                                   uint32_t flag_words) {
    extern __shared__ __align__(16) uint8_t link_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per warp: sel ids | sel dist | row ids | row dist (smem_stride entries each) | flags
    uint32_t* base = reinterpret_cast<uint32_t*>(link_smem) + (size_t)warp * (4 * smem_stride + flag_words);
    uint32_t* sel = base;
    float* sel_d = reinterpret_cast<float*>(base + smem_stride);
    uint32_t* rid = base + 2 * smem_stride;
    float* rd = reinterpret_cast<float*>(base + 3 * smem_stride);
    uint32_t* flags = base + 4 * smem_stride;
    const uint32_t task = blockIdx.x * kLinkWarps + warp;
    if (task >= t.n_tasks) return;
    const uint32_t node = t.node ? t.node[task] : t.first_node + task;
    const uint32_t layer = t.layer ? t.layer[task] : 0u;
    const RowPtr own = link_row(b, layer, node);
    if (!own.ids) return;
    const uint32_t maxc = own.maxc;
    const uint32_t wlen = min(t.cand_cnt[task], t.cand_stride);
    const uint32_t* W = t.cand_ids + (size_t)task * t.cand_stride;
    const float* Wd = t.cand_dist + (size_t)task * t.cand_stride;
    uint32_t nsel = 0;
    // ---- select_neighbors (graph.rs:526-581) ----
    if (wlen <= maxc) {
        for (uint32_t i = lane; i < wlen; i += 32) {
            sel[i] = W[i];
            sel_d[i] = Wd[i];
        }
        nsel = wlen;
        __syncwarp();
    } else {
        for (uint32_t i = lane; i < flag_words; i += 32) flags[i] = 0u;
        __syncwarp();
        for (uint32_t i = 0; i < wlen && nsel < maxc; ++i) {
            const uint32_t c = W[i];
            const float dc = Wd[i];
            bool diverse = true;
            for (uint32_t s = 0; s < nsel; ++s) {
                const float ds = node_pair_dist(b.ix, c, sel[s], lane);
                if (!(dc <= ds)) {  // alpha = 1.0: keep c only if alpha * d(q, c) <= d(c, s) for every selected s
                    diverse = false;
                    break;
                }
            }
            if (diverse) {
                if (lane == 0) {
                    sel[nsel] = c;
                    flags[i >> 5] |= 1u << (i & 31);
                }
                ++nsel;
                __syncwarp();
            }
        }
        // back-fill with the closest candidates not selected yet (graph.rs:566-578)
        for (uint32_t i = 0; i < wlen && nsel < maxc; ++i) {
            if ((flags[i >> 5] >> (i & 31)) & 1u) continue;
            __syncwarp();
            if (lane == 0) flags[i >> 5] |= 1u << (i & 31);
            ++nsel;
            __syncwarp();
        }
        // the selection in candidate order = ascending (distance, id)
        uint32_t w = 0;
        for (uint32_t base_i = 0; base_i < wlen; base_i += 32) {
            const uint32_t i = base_i + lane;
            const bool on = i < wlen && ((flags[i >> 5] >> (i & 31)) & 1u);
            const uint32_t msk = __ballot_sync(FULL_MASK, on);
            if (on) {
                const uint32_t p = w + __popc(msk & ((1u << lane) - 1u));
                sel[p] = W[i];
                sel_d[p] = Wd[i];
            }
            w += __popc(msk);
        }
        nsel = w;
        __syncwarp();
    }
    // ---- set_neighbors(node, selected) (graph.rs:212-213) ----
    if (!seed) {
        for (uint32_t i = lane; i < nsel; i += 32) {
            __stcg(own.ids + i, sel[i]);
            __stcg(own.dist + i, sel_d[i]);
        }
    } else {
        for (uint32_t i = 0; i < nsel; ++i) locked_offer(own, sel[i], sel_d[i], rid, rd, lane);
    }
    // ---- add_bidirectional_connection for each selected neighbour (graph.rs:214-217, 592-639) ----
    for (uint32_t i = 0; i < nsel; ++i) {
        const RowPtr r = link_row(b, layer, sel[i]);
        if (r.ids) locked_offer(r, node, sel_d[i], rid, rd, lane);
    }
}

struct BlockPlan {
    uint64_t begin, end;
    uint32_t cur_max, entry;
    uint32_t job0, n_jobs, task0, n_tasks;  // upper-layer jobs / tasks of this block
};

static uint32_t env_u32b(const char* name, uint32_t dflt) {
    const char* v = std::getenv(name);
    return v && *v ? (uint32_t)std::strtoul(v, nullptr, 10) : dflt;
}

// distances of the existing links to their rows' owners: the side arrays of a graph that was loaded or built earlier
// (they exist only during construction).  One warp per node: its layer-0 row, then its upper rows.
__global__ void link_dist_kernel(IncView b, uint32_t n_old) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t node = gw; node < n_old; node += nwarps) {
        const uint32_t ref = b.ix.upper_ref[node];
        const uint32_t top = ref == VELES_INVALID_ID ? 0u : (ref & 15u);
        for (uint32_t l = 0; l <= top; ++l) {
            const RowPtr r = link_row(b, l, (uint32_t)node);
            if (!r.ids) continue;
            for (uint32_t j = 0; j < r.stride; ++j) {
                const uint32_t x = r.ids[j];
                if (x == VELES_INVALID_ID) break;
                const float d = node_pair_dist(b.ix, (uint32_t)node, x, lane);
                if (lane == 0) r.dist[j] = d;
            }
        }
    }
}

}  // namespace veles

using namespace veles;

// Inserts nodes [first, ix->n) into the graph, a block at a time (file header).  first == 0: a fresh build (M, levels,
// rows planned from scratch; the first block is seeded all-pairs).  first > 0: the vectors of the new nodes are already
// in ix->vecs and the graph over nodes < first is live: its arrays are grown, the link distances of its rows are
// recomputed once, and the new nodes are linked in exactly as the later blocks of a fresh build are.  ix->mu is held.
static int32_t insert_blocks(veles_index* ix, uint64_t first, uint32_t M, uint32_t ef_construction, cudaStream_t st) {
    const uint64_t n = ix->n;
    const bool append = first > 0;
    const uint32_t M0 = 2 * M;
    const uint32_t ef_c = ef_construction;
    const uint64_t old_rows = append ? ix->upper_rows : 0;
    if (!append) {
        ix->M = M;
        ix->M0 = M0;
        ix->stride0 = round_up(M0, 32);
        ix->strideU = round_up(M, 32);
        ix->has_graph = true;
        ix->has_entry = false;
        ix->entry = 0;
        ix->max_layer = 0;
        ix->num_layers = 1;
        ix->upper_rows = 0;
    }
    ix->ef_construction = ef_c;
    if (n == 0) {
        VELES_TRY(ix->adj0.alloc(16));
        VELES_TRY(ix->upper_ref.alloc(16));
        VELES_TRY(ix->upper_adj.alloc(16));
        return VELES_OK;
    }
    const auto t_start = std::chrono::steady_clock::now();
    // levels and upper rows (graph.rs:168-178).  Levels follow the reference PRNG in node-id order over the whole
    // collection; an append plans rows only for the new nodes, after the rows the live graph already owns.
    std::vector<uint8_t> level;
    reference_levels(n, M, level);
    std::vector<uint32_t> h_ref(n - first, VELES_INVALID_ID);
    uint64_t rows = old_rows;
    for (uint64_t i = first; i < n; ++i)
        if (level[i] > 0) {
            VELES_REQUIRE(rows < (1ull << 28), "too many upper-layer rows");
            h_ref[i - first] = (uint32_t)(rows << 4) | level[i];
            rows += level[i];
        }
    ix->upper_rows = rows;

    // ---- block plan: seed block, then blocks of at most 1/growth of the graph, cut after a node that raises the
    // top layer (it becomes the entry point, graph.rs:230-233, and the next block must see it) ----
    const uint32_t n_seed = append ? 0u : (uint32_t)std::min<uint64_t>(n, std::max(2u, env_u32b("VELES_BUILD_SEED", 256)));
    const uint32_t growth = std::max(1u, env_u32b("VELES_BUILD_GROWTH", 4));
    const uint32_t cap_auto = (uint32_t)std::min<uint64_t>(65536, std::max<uint64_t>(1024, n / 64));
    const uint32_t cap = std::max(32u, env_u32b("VELES_BUILD_BLOCK_CAP", cap_auto));
    std::vector<BlockPlan> plan;
    std::vector<uint32_t> job_node, job_task0, task_node;
    std::vector<uint8_t> job_top, task_layer;
    uint32_t cur_max = append ? ix->max_layer : 0u, entry = append ? (uint32_t)ix->entry : 0u;
    for (uint32_t i = 0; i < n_seed; ++i)
        if (level[i] > cur_max) {
            cur_max = level[i];
            entry = i;
        }
    if (!append) {
        BlockPlan bp{0, n_seed, cur_max, entry, 0, 0, 0, 0};
        for (uint32_t i = 0; i < n_seed; ++i)
            for (uint32_t l = level[i]; l >= 1; --l) {
                task_node.push_back(i);
                task_layer.push_back((uint8_t)l);
            }
        bp.n_tasks = (uint32_t)task_node.size();
        plan.push_back(bp);
    }
    uint32_t max_block = std::max(n_seed, 1u), max_jobs = 0, max_tasks = append ? 0u : plan[0].n_tasks;
    for (uint64_t done = append ? first : n_seed; done < n;) {
        const uint64_t want = std::min<uint64_t>(cap, std::max<uint64_t>(32, done / growth));
        uint64_t end = std::min<uint64_t>(n, done + want);
        uint32_t new_top = VELES_INVALID_ID;
        for (uint64_t i = done; i < end; ++i)
            if (level[i] > cur_max) {
                new_top = (uint32_t)i;
                end = i + 1;
                break;
            }
        BlockPlan bp{done, end, cur_max, entry, (uint32_t)job_node.size(), 0, (uint32_t)task_node.size(), 0};
        for (uint64_t i = done; i < end; ++i) {
            const uint32_t top = std::min<uint32_t>(level[i], cur_max);
            if (top == 0) continue;
            job_node.push_back((uint32_t)i);
            job_top.push_back((uint8_t)top);
            job_task0.push_back((uint32_t)task_node.size() - bp.task0);  // relative to the block's first task
            for (uint32_t l = top; l >= 1; --l) {
                task_node.push_back((uint32_t)i);
                task_layer.push_back((uint8_t)l);
            }
        }
        bp.n_jobs = (uint32_t)job_node.size() - bp.job0;
        bp.n_tasks = (uint32_t)task_node.size() - bp.task0;
        plan.push_back(bp);
        max_block = std::max<uint32_t>(max_block, (uint32_t)(end - done));
        max_jobs = std::max(max_jobs, bp.n_jobs);
        max_tasks = std::max(max_tasks, bp.n_tasks);
        if (new_top != VELES_INVALID_ID) {
            cur_max = level[new_top];
            entry = new_top;
        }
        done = end;
    }

    // ---- device state ----
    DevBuf adj0_d, lock0, upper_d, lockU, level_d, jobs_node_d, jobs_top_d, jobs_task0_d, task_node_d, task_layer_d;
    DevBuf qbuf, c0_ids, c0_dist, c0_cnt, cu_ids, cu_dist, cu_cnt, uvis, ulog, utie, uerr;
    // graph arrays for n nodes / `rows` upper rows: fresh, or grown with the live graph's content in front
    auto grow = [&](DevBuf& buf, size_t old_bytes, size_t new_bytes) -> int32_t {
        DevBuf nb;
        VELES_TRY(nb.alloc(std::max<size_t>(new_bytes, 16)));
        if (old_bytes) VELES_CUDA(cudaMemcpyAsync(nb.p, buf.p, old_bytes, cudaMemcpyDeviceToDevice, st));
        if (new_bytes > old_bytes)
            VELES_CUDA(cudaMemsetAsync(static_cast<uint8_t*>(nb.p) + old_bytes, 0xff, new_bytes - old_bytes, st));
        VELES_CUDA(cudaStreamSynchronize(st));  // the old buffer is freed when `nb` goes out of scope
        std::swap(buf.p, nb.p);
        std::swap(buf.bytes, nb.bytes);
        return VELES_OK;
    };
    VELES_TRY(grow(ix->adj0, (size_t)first * ix->stride0 * 4, (size_t)n * ix->stride0 * 4));
    VELES_TRY(grow(ix->upper_ref, (size_t)first * 4, (size_t)n * 4));
    VELES_TRY(grow(ix->upper_adj, (size_t)old_rows * ix->strideU * 4, (size_t)rows * ix->strideU * 4));
    VELES_TRY(adj0_d.alloc((size_t)n * ix->stride0 * 4));
    VELES_TRY(lock0.alloc((size_t)n * 4));
    VELES_TRY(upper_d.alloc(std::max<size_t>((size_t)rows * ix->strideU * 4, 16)));
    VELES_TRY(lockU.alloc(std::max<size_t>((size_t)rows * 4, 16)));
    VELES_TRY(level_d.alloc((size_t)n));
    VELES_CUDA(cudaMemsetAsync(lock0.p, 0, lock0.bytes, st));
    VELES_CUDA(cudaMemsetAsync(lockU.p, 0, lockU.bytes, st));
    VELES_CUDA(cudaMemcpyAsync(ix->upper_ref.as<uint32_t>() + first, h_ref.data(), (size_t)(n - first) * 4, cudaMemcpyHostToDevice, st));
    VELES_CUDA(cudaMemcpyAsync(level_d.p, level.data(), (size_t)n, cudaMemcpyHostToDevice, st));
    auto upload = [&](DevBuf& d, const void* src, size_t bytes) -> int32_t {
        VELES_TRY(d.alloc(std::max<size_t>(bytes, 16)));
        if (bytes) VELES_CUDA(cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, st));
        return VELES_OK;
    };
    VELES_TRY(upload(jobs_node_d, job_node.data(), job_node.size() * 4));
    VELES_TRY(upload(jobs_top_d, job_top.data(), job_top.size()));
    VELES_TRY(upload(jobs_task0_d, job_task0.data(), job_task0.size() * 4));
    VELES_TRY(upload(task_node_d, task_node.data(), task_node.size() * 4));
    VELES_TRY(upload(task_layer_d, task_layer.data(), task_layer.size()));
    const uint32_t cstride = ef_c;
    VELES_TRY(qbuf.alloc((size_t)max_block * ix->dim * 4));
    VELES_TRY(c0_ids.alloc((size_t)max_block * cstride * 4));
    VELES_TRY(c0_dist.alloc((size_t)max_block * cstride * 4));
    VELES_TRY(c0_cnt.alloc((size_t)max_block * 4));
    VELES_TRY(cu_ids.alloc(std::max<size_t>((size_t)max_tasks * cstride * 4, 16)));
    VELES_TRY(cu_dist.alloc(std::max<size_t>((size_t)max_tasks * cstride * 4, 16)));
    VELES_TRY(cu_cnt.alloc(std::max<size_t>((size_t)max_tasks * 4, 16)));
    const int sms = device_sm_count();
    const uint32_t up_slots = std::max(1u, std::min<uint32_t>(std::max(max_jobs, 1u), (uint32_t)sms * 8));
    const uint32_t up_words = (uint32_t)((rows + 31) / 32 + 1);
    VELES_TRY(uvis.alloc((size_t)up_slots * up_words * 4));
    VELES_TRY(ulog.alloc((size_t)up_slots * kBuildLogCap * 4));
    VELES_TRY(utie.alloc((size_t)up_slots * kBuildTieCap * 8));
    VELES_TRY(uerr.alloc(16));
    VELES_CUDA(cudaMemsetAsync(uvis.p, 0, uvis.bytes, st));
    VELES_CUDA(cudaMemsetAsync(uerr.p, 0, 16, st));

    IncView b;
    b.adj0 = ix->adj0.as<uint32_t>();
    b.adj0_d = adj0_d.as<float>();
    b.lock0 = lock0.as<uint32_t>();
    b.upper_adj = ix->upper_adj.as<uint32_t>();
    b.upper_d = upper_d.as<float>();
    b.lockU = lockU.as<uint32_t>();
    b.level = level_d.as<uint8_t>();
    b.M = M;
    b.M0 = M0;
    b.ef_c = ef_c;

    const uint32_t smem_stride = std::max(ix->stride0, ix->strideU);
    const uint32_t flag_words = (cstride + 31) / 32;
    const size_t link_smem = (size_t)kLinkWarps * (4 * smem_stride + flag_words) * 4;
    VELES_CUDA(cudaFuncSetAttribute(inc_link_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)link_smem));
    const size_t res_smem = ((size_t)ef_c + 2) * 8;
    VELES_CUDA(cudaFuncSetAttribute(inc_upper_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)res_smem));
    VELES_CUDA(cudaFuncSetAttribute(inc_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)res_smem));
    auto link = [&](const TaskList& t, uint32_t seed) -> int32_t {
        if (t.n_tasks == 0) return VELES_OK;
        inc_link_kernel<<<(t.n_tasks + kLinkWarps - 1) / kLinkWarps, kLinkWarps * 32, link_smem, st>>>(b, t, seed, smem_stride,
                                                                                                         flag_words);
        count_launch();
        VELES_CUDA(cudaGetLastError());
        return VELES_OK;
    };
    const bool verbose = env_u32b("VELES_BUILD_VERBOSE", 0) != 0;
    if (append) {
        b.ix = ix->view();
        b.cur_max_layer = ix->max_layer;
        b.entry = (uint32_t)ix->entry;
        link_dist_kernel<<<sms * 8, 256, 0, st>>>(b, (uint32_t)first);
        count_launch();
        VELES_CUDA(cudaGetLastError());
    }
    SearchCtx* sctx = nullptr;
    VELES_TRY(acquire_ctx(ix, st, true, &sctx));  // one context for every block's search (same stream, overflow flag sticky)
    struct CtxGuard {
        SearchCtx* c;
        ~CtxGuard() { c->in_use = false; }  // ix->mu is held for the whole build
    } ctx_guard{sctx};

    for (size_t bi = 0; bi < plan.size(); ++bi) {
        const BlockPlan& bp = plan[bi];
        const uint32_t count = (uint32_t)(bp.end - bp.begin);
        ix->entry = bp.entry;
        ix->max_layer = bp.cur_max;
        ix->num_layers = bp.cur_max + 1;
        ix->has_entry = true;
        b.ix = ix->view();
        b.cur_max_layer = bp.cur_max;
        b.entry = bp.entry;
        TaskList t0{nullptr, nullptr, (uint32_t)bp.begin, count, c0_ids.as<uint32_t>(), c0_dist.as<float>(), c0_cnt.as<uint32_t>(), cstride};
        TaskList tu{task_node_d.as<uint32_t>() + bp.task0, task_layer_d.as<uint8_t>() + bp.task0, 0, bp.n_tasks,
                    cu_ids.as<uint32_t>(), cu_dist.as<float>(), cu_cnt.as<uint32_t>(), cstride};
        if (bi == 0 && !append) {
            inc_seed_kernel<<<std::min<uint32_t>(count, (uint32_t)sms * 16), 32, res_smem, st>>>(b, t0, n_seed);
            count_launch();
            if (tu.n_tasks) {
                inc_seed_kernel<<<std::min<uint32_t>(tu.n_tasks, (uint32_t)sms * 16), 32, res_smem, st>>>(b, tu, n_seed);
                count_launch();
            }
            VELES_CUDA(cudaGetLastError());
            VELES_TRY(link(t0, 1));
            VELES_TRY(link(tu, 1));
            continue;
        }
        // 1. upper layers
        if (bp.n_jobs) {
            UpperJobs u{jobs_node_d.as<uint32_t>() + bp.job0, jobs_top_d.as<uint8_t>() + bp.job0,
                        jobs_task0_d.as<uint32_t>() + bp.job0, bp.n_jobs, uvis.as<uint32_t>(), ulog.as<uint32_t>(),
                        utie.as<uint64_t>(), up_words, uerr.as<uint32_t>()};
            inc_upper_search_kernel<<<std::min(bp.n_jobs, up_slots), 32, res_smem, st>>>(b, u, tu);
            count_launch();
            VELES_CUDA(cudaGetLastError());
        }
        // 2. layer 0: the production search kernel over the graph built so far, one query per new node
        rows_to_queries_kernel<<<std::min<uint32_t>((count + 7) / 8, (uint32_t)sms * 8), 256, 0, st>>>(b.ix, bp.begin, count,
                                                                                                       qbuf.as<float>());
        count_launch();
        VELES_CUDA(cudaGetLastError());
        VELES_TRY(launch_search(ix, ix->view(), sctx, qbuf.as<float>(), count, ef_c, ef_c, c0_ids.as<uint32_t>(), c0_dist.as<float>(),
                                c0_cnt.as<uint32_t>(), nullptr, st));
        // 3. link
        VELES_TRY(link(t0, 0));
        VELES_TRY(link(tu, 0));
        if (verbose && (bi % 16 == 0 || bi + 1 == plan.size())) {
            VELES_CUDA(cudaStreamSynchronize(st));
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
            std::fprintf(stderr, "[veles build] block %zu/%zu: %llu/%llu nodes, %.2f s\n", bi + 1, plan.size(),
                         (unsigned long long)bp.end, (unsigned long long)n, s);
        }
    }
    // final entry point / top layer (graph.rs:230-233)
    ix->entry = entry;
    ix->max_layer = cur_max;
    ix->num_layers = cur_max + 1;
    ix->has_entry = true;
    if (sctx->launched) VELES_TRY(check_search_error_flag(sctx, st));
    uint32_t h_err[4] = {0, 0, 0, 0};
    VELES_CUDA(cudaMemcpyAsync(h_err, uerr.p, 16, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaStreamSynchronize(st));
    if (h_err[0] != 0) {
        set_error("builder: tie list overflow in an upper-layer beam (> %u equal-distance candidates)", kBuildTieCap);
        return VELES_ERR_OVERFLOW;
    }
    return VELES_OK;
}

extern "C" int32_t veles_index_build_graph(veles_index_t* ix, uint32_t M, uint32_t ef_construction, void* stream) {
    NvtxRange nvtx_range("veles::build_graph (NativeHnsw::insert, block insertion)");
    VELES_REQUIRE(ix != nullptr, "index is NULL");
    VELES_REQUIRE(M >= 2 && M <= 128, "M must be in 2..128, got %u", M);
    if (ef_construction == 0) ef_construction = env_u32b("VELES_BUILD_EF", 200);
    VELES_REQUIRE(ef_construction >= 1 && ef_construction <= 4096, "ef_construction must be in 1..4096, got %u", ef_construction);
    std::lock_guard<std::mutex> g(ix->mu);
    return insert_blocks(ix, 0, M, ef_construction, (cudaStream_t)stream);
}

// HnswIndex::insert_batch_parallel on a live index (index/hnsw/index/batch.rs:82-108): `count` more vectors become
// nodes n .. n + count - 1 and are inserted into the existing graph, a block at a time, exactly as the later blocks of
// veles_index_build_graph are -- no rebuild.  Works on built and on loaded graphs (M is the graph's).
extern "C" int32_t veles_index_append(veles_index_t* ix, const void* vectors, uint64_t count, int32_t src_dtype,
                                      uint32_t ef_construction, void* stream) {
    NvtxRange nvtx_range("veles::append (HnswIndex::insert_batch_parallel on a live graph)");
    VELES_REQUIRE(ix != nullptr, "index is NULL");
    VELES_REQUIRE(count == 0 || vectors != nullptr, "vectors is NULL");
    VELES_REQUIRE(src_dtype == VELES_F32 || src_dtype == ix->dtype, "unsupported conversion: src dtype %d -> store dtype %d", src_dtype,
                  ix->dtype);
    VELES_REQUIRE(ef_construction <= 4096, "ef_construction must be <= 4096, got %u", ef_construction);
    if (count == 0) return VELES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t first = 0;
    {
        std::lock_guard<std::mutex> g(ix->mu);
        VELES_REQUIRE(ix->n == 0 || (ix->has_graph && ix->M >= 2), "snapshot has no graph to append to; build or load one first");
        VELES_REQUIRE(ix->n + count < (1ull << 31), "at most 2^31-1 nodes per snapshot");
        first = ix->n;
        // grow the row store, old rows in front
        DevBuf nv;
        VELES_TRY(nv.alloc((size_t)(first + count) * ix->row_bytes));
        if (first) VELES_CUDA(cudaMemcpyAsync(nv.p, ix->vecs.p, (size_t)first * ix->row_bytes, cudaMemcpyDeviceToDevice, st));
        VELES_CUDA(cudaMemsetAsync(static_cast<uint8_t*>(nv.p) + (size_t)first * ix->row_bytes, 0, (size_t)count * ix->row_bytes, st));
        VELES_CUDA(cudaStreamSynchronize(st));
        std::swap(ix->vecs.p, nv.p);
        std::swap(ix->vecs.bytes, nv.bytes);
        ix->n = first + count;
        ix->x16.release();  // derived copies no longer cover the collection
        ix->x16_dpad = 0;
        ix->has_sq8 = false;
    }
    {
        // the new rows: staged on the device in the caller's type, then the conversion of veles_index_set_rows_d
        const size_t width = src_dtype == VELES_BIN1 ? (size_t)(ix->dim / 64) * 8 : (size_t)ix->dim * (src_dtype == VELES_F32 ? 4 : 2);
        DevBuf stage;
        VELES_TRY(stage.alloc(count * width));
        VELES_CUDA(cudaMemcpyAsync(stage.p, vectors, count * width, cudaMemcpyHostToDevice, st));
        VELES_TRY(veles_index_set_rows_d(ix, first, count, stage.p, src_dtype, stream));
        VELES_CUDA(cudaStreamSynchronize(st));
    }
    std::lock_guard<std::mutex> g(ix->mu);
    const uint32_t efc = ef_construction ? ef_construction : (ix->ef_construction ? std::min(ix->ef_construction, 4096u) : 200u);
    if (first == 0) return insert_blocks(ix, 0, ix->M >= 2 ? ix->M : 16, efc, st);
    return insert_blocks(ix, first, ix->M, efc, st);
}

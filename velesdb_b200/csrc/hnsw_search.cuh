// hnsw_search.cuh -- batched HNSW search: NativeHnsw::search (native/graph.rs:251-270) for a batch
// of queries, one warp per query, persistent grid.
//
// Reference semantics kept bit for bit:
//   * search_layer_single (graph.rs:405-428): scan all neighbours of `best` in stored order,
//     move on strict improvement, repeat until a scan brings none.
//   * search_layer (graph.rs:438-520): candidates = min-heap on (dist, id), results = max-heap on
//     (dist, id) capped at ef; pop the closest candidate, stop when it is farther than the worst
//     result and the result set is full; every not-yet-visited neighbour is evaluated in stored
//     order and accepted when `d < worst || len < ef`.
//
// How the two heaps are represented here.  Every accepted node enters both heaps, so
//   candidates = {unexpanded members of results}  U  {evicted, not yet popped}.
// `res` is one array sorted by (dist, id) with an "expanded" bit per entry (in registers for ef <= 256, else
// in shared memory: ResArr); the next candidate is its first unexpanded entry.  An evicted node can only ever be popped *without* ending the loop
// when its distance equals the current worst distance (it was the maximum when evicted, and the
// worst distance never grows), so only those are kept, in the per-query tie list `tie`; everything
// else evicted could only trigger the break and is dropped.  Pop order between the two sets is
// fixed: members of `res` always sort before anything evicted.  This reproduces the reference's
// expansion order exactly, including on integer-valued metrics where ties are the rule.
//
// Data movement: a popped node's adjacency row (stride0 x u32) is read with coalesced 128-byte
// loads, filtered through a per-query visited bitmap in HBM/L2 (atomicOr = test-and-set), and the
// surviving neighbours' rows are fetched by 1-D bulk async copies (TMA engine, mbarrier
// complete_tx) into a ring of shared-memory slots; the warp computes each distance from shared
// memory with the reference's accumulation tree (common.cuh) while later rows are in flight.
#pragma once

#include <cstdlib>

#include "index.hpp"

namespace veles {

constexpr uint32_t kMaxSlots = 16;      // ring slots per warp (upper bound)
constexpr uint32_t kTieCap = 4096;      // per-query tie-list capacity (keys), lives in global scratch
constexpr uint32_t kLogCap = 16384;     // visited-log entries per query slot before falling back to a full clear
constexpr uint32_t kMaxPeers = 7;       // other GPUs of one NVSwitch box

struct SearchParams {
    IndexView ix;
    const float* queries;
    uint32_t nq, k, ef;
    uint32_t* out_ids;
    float* out_dist;
    uint32_t* out_counts;
    uint32_t* out_stats;  // may be null
    const uint32_t* extra_entries;  // nq x 3 extra layer-0 entry points (INVALID padded), or null
    // multi-GPU gather fused into the epilogue (comm.cu): every query's results are also stored into the same
    // position of up to kMaxPeers peer windows -- peer-mapped memory, i.e. posted stores over NVLink
    uint32_t n_peers;
    uint32_t* peer_ids[kMaxPeers];
    float* peer_dist[kMaxPeers];
    uint32_t* peer_cnt[kMaxPeers];
    uint32_t* visited;    // slots x vis_words
    uint32_t* vlog;       // slots x kLogCap
    uint64_t* tie;        // slots x kTieCap
    uint32_t vis_words;
    uint32_t* counters;   // [0] work counter, [1] error flag
    uint32_t nslot;       // ring slots (multiple of 4 when quad != 0)
    uint32_t quad;        // 1: evaluate four candidates per step (8 lanes each), needs dim % 32 == 0
    uint32_t evict_first; // 1: vector rows are fetched with an L2 evict-first policy
    uint32_t peek;        // 1: speculative read-only visited test of the predicted next candidate's neighbours
    // shared-memory carve (bytes from base)
    uint32_t off_res, off_todo, off_dist, off_q, off_ring;
    uint32_t warp_bytes;  // per-warp block at off_ring: kMaxSlots barriers, then the ring
};

struct WarpCtx {
    uint64_t* bar;
    uint64_t* res;
    uint32_t* todo;
    uint8_t* q;
    uint8_t* ring;
    uint32_t phases;
    uint64_t policy;
    float norm_a;
    uint32_t lane;
};

__device__ __forceinline__ uint64_t make_key(float d, uint32_t id) {
    return ((uint64_t)ord_key(d) << 32) | ((uint64_t)id << 1);
}
__device__ __forceinline__ float key_dist(uint64_t key) { return ord_unkey((uint32_t)(key >> 32)); }
__device__ __forceinline__ uint32_t key_id(uint64_t key) { return ((uint32_t)key) >> 1; }

// distance_l2_quantized (native/quantization.rs:42-92) on four codes: sum of squared byte differences.  Integer
// sums, so the reference's accumulator split is immaterial.  The u32 total travels through the beam as the
// float with the same bit pattern: for values below 0x7f800000 (dim <= 32768) float order == integer order,
// denormals included (no flush-to-zero in this build).
__device__ __forceinline__ uint32_t sq8_word(uint32_t a, uint32_t b, uint32_t acc) {
    const uint32_t d = __vabsdiffu4(a, b);
    return __dp4a(d, d, acc);
}

template <int DT>
__device__ __forceinline__ float row_distance(const SearchParams& p, const WarpCtx& c, const uint8_t* row) {
    if (DT == VELES_BIN1) {
        const uint32_t words = p.ix.dim >> 5;
        const uint32_t* qw = reinterpret_cast<const uint32_t*>(c.q);
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(row);
        uint32_t d = 0;
        for (uint32_t i = c.lane; i < words; i += 32) d += __popc(qw[i] ^ rw[i]);
        return (float)__reduce_add_sync(FULL_MASK, d);
    } else if (DT == VELES_SQ8) {
        const uint32_t words = p.ix.row_bytes >> 2;  // zero padded on both sides
        const uint32_t* qw = reinterpret_cast<const uint32_t*>(c.q);
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(row);
        uint32_t d = 0;
        for (uint32_t i = c.lane; i < words; i += 32) d = sq8_word(qw[i], rw[i], d);
        return __uint_as_float(__reduce_add_sync(FULL_MASK, d));
    } else {
        float norm_b = 0.0f;
        if (p.ix.metric == VELES_COSINE) norm_b = *reinterpret_cast<const float*>(row + p.ix.norm_off);
        if (DT == VELES_F32)
            return warp_metric(p.ix.metric, false, reinterpret_cast<const float*>(c.q), reinterpret_cast<const float*>(row),
                               p.ix.dim, c.norm_a, norm_b, c.lane);
        else
            return warp_metric(p.ix.metric, false, reinterpret_cast<const float*>(c.q), reinterpret_cast<const __half*>(row),
                               p.ix.dim, c.norm_a, norm_b, c.lane);
    }
}

__device__ __forceinline__ void copy_row(const SearchParams& p, const WarpCtx& c, uint32_t slot, uint32_t id, uint64_t* bar) {
    void* dst = c.ring + (size_t)slot * p.ix.row_bytes;
    const void* src = p.ix.vecs + (size_t)id * p.ix.row_bytes;
    if (p.evict_first)
        bulk_g2s_hint(dst, src, p.ix.row_bytes, bar, c.policy);
    else
        bulk_g2s(dst, src, p.ix.row_bytes, bar);
}
__device__ __forceinline__ void issue_row(const SearchParams& p, const WarpCtx& c, uint32_t slot, uint32_t id) {
    mbar_expect_tx(&c.bar[slot], p.ix.row_bytes);
    copy_row(p, c, slot, id, &c.bar[slot]);
}

// U steps of the quad accumulation with all 2*U shared-memory loads issued before the first FMA, so the
// LDS latency is paid once per block instead of once per step (the FMA order per accumulator -- increasing
// element index -- is unchanged, hence so are the bits).
template <int U, bool L2, typename TB>
__device__ __forceinline__ void quad_block(const TB* __restrict__ r, const float* __restrict__ q, uint32_t i, float& a0,
                                           float& a1, float& a2, float& a3) {
    float4 x[U], y[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        x[u] = load4(r + i + 32 * u);
        y[u] = load4(q + i + 32 * u);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (L2) {
            const float d0 = __fsub_rn(y[u].x, x[u].x), d1 = __fsub_rn(y[u].y, x[u].y);
            const float d2 = __fsub_rn(y[u].z, x[u].z), d3 = __fsub_rn(y[u].w, x[u].w);
            a0 = __fmaf_rn(d0, d0, a0);
            a1 = __fmaf_rn(d1, d1, a1);
            a2 = __fmaf_rn(d2, d2, a2);
            a3 = __fmaf_rn(d3, d3, a3);
        } else {
            a0 = __fmaf_rn(y[u].x, x[u].x, a0);
            a1 = __fmaf_rn(y[u].y, x[u].y, a1);
            a2 = __fmaf_rn(y[u].z, x[u].z, a2);
            a3 = __fmaf_rn(y[u].w, x[u].w, a3);
        }
    }
}

// Main loop over the 32-element blocks (simd_avx512.rs:166-184), then -- when dim is not a multiple of 32 -- the
// reference's tails on top of the reduced main-loop sum: whole 8-element blocks, each reduced with the f32x8
// tree (the group's eight lanes are the eight SIMD lanes) and added, then the remaining elements one by one
// (simd_avx512.rs:186-201; same arithmetic as warp_tree_tail).  All eight lanes of a group return the same value.
template <bool L2, typename TB>
__device__ __forceinline__ float quad_accumulate(const TB* __restrict__ r, const float* __restrict__ q, uint32_t dim,
                                                 uint32_t t) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const uint32_t main_len = dim & ~31u;
    uint32_t i = t * 4;
    for (; i + 32 * 7 < main_len; i += 32 * 8) quad_block<8, L2>(r, q, i, a0, a1, a2, a3);
    for (; i + 32 * 3 < main_len; i += 32 * 4) quad_block<4, L2>(r, q, i, a0, a1, a2, a3);
    for (; i < main_len; i += 32) quad_block<1, L2>(r, q, i, a0, a1, a2, a3);
    float result = quad_tree_sum(a0, a1, a2, a3);
    if (main_len != dim) {
        uint32_t pos = main_len;
        for (; pos + 8 <= dim; pos += 8) {
            const float x = q[pos + t], y = load_elem(r, pos + t);
            float v;
            if (L2) {
                const float d = __fsub_rn(x, y);
                v = __fmaf_rn(d, d, 0.0f);
            } else {
                v = __fmaf_rn(x, y, 0.0f);
            }
            result = __fadd_rn(result, warp_tree_sum8(v));  // xor 4, 2, 1 stay inside the 8-lane group
        }
        for (; pos < dim; ++pos) {
            const float x = q[pos], y = load_elem(r, pos);
            if (L2) {
                const float d = __fsub_rn(x, y);
                result = __fadd_rn(result, __fmul_rn(d, d));
            } else {
                result = __fadd_rn(result, __fmul_rn(x, y));
            }
        }
    }
    return result;
}

// distance of the row owned by this lane's group (f32 / f16 rows: dim >= 16, the reference's wide16 regime)
template <int DT>
__device__ __forceinline__ float quad_distance(const SearchParams& p, const WarpCtx& c, const uint8_t* row) {
    const uint32_t t = c.lane & 7;
    if (DT == VELES_BIN1) {
        const uint32_t words = p.ix.dim >> 5;
        const uint32_t* qw = reinterpret_cast<const uint32_t*>(c.q);
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(row);
        uint32_t d = 0;
        for (uint32_t w = t * 4; w < words; w += 32) {
            const uint4 x = *reinterpret_cast<const uint4*>(rw + w);
            const uint4 y = *reinterpret_cast<const uint4*>(qw + w);
            d += __popc(x.x ^ y.x) + __popc(x.y ^ y.y) + __popc(x.z ^ y.z) + __popc(x.w ^ y.w);
        }
        d += __shfl_xor_sync(FULL_MASK, d, 1);
        d += __shfl_xor_sync(FULL_MASK, d, 2);
        d += __shfl_xor_sync(FULL_MASK, d, 4);
        return (float)d;
    } else if (DT == VELES_SQ8) {
        const uint32_t n16 = p.ix.row_bytes >> 4;
        const uint4* qw = reinterpret_cast<const uint4*>(c.q);
        const uint4* rw = reinterpret_cast<const uint4*>(row);
        uint32_t d = 0;
        for (uint32_t w = t; w < n16; w += 8) {
            const uint4 x = rw[w], y = qw[w];
            d = sq8_word(x.x, y.x, d);
            d = sq8_word(x.y, y.y, d);
            d = sq8_word(x.z, y.z, d);
            d = sq8_word(x.w, y.w, d);
        }
        d += __shfl_xor_sync(FULL_MASK, d, 1);
        d += __shfl_xor_sync(FULL_MASK, d, 2);
        d += __shfl_xor_sync(FULL_MASK, d, 4);
        return __uint_as_float(d);
    } else {
        using TB = typename std::conditional<DT == VELES_F32, float, __half>::type;
        const TB* r = reinterpret_cast<const TB*>(row);
        const float* q = reinterpret_cast<const float*>(c.q);
        const uint32_t dim = p.ix.dim;
        if (p.ix.metric == VELES_EUCLIDEAN) return __fsqrt_rn(quad_accumulate<true>(r, q, dim, t));
        const float dot = quad_accumulate<false>(r, q, dim, t);
        if (p.ix.metric == VELES_COSINE) {
            const float nb = *reinterpret_cast<const float*>(row + p.ix.norm_off);
            return __fsub_rn(1.0f, cosine_from_parts(dot, c.norm_a, nb));
        }
        return -dot;
    }
}

// Evaluates quads first, first + step, ... of c.todo[0..m) (quad j = rows 4j..4j+3), four rows per evaluation
// step; stage s uses slots 4s..4s+3 and barrier s.  `sink(j, cnt, d)` is called by all lanes after each step
// with the quad's index, its row count and -- in the lanes of group g = lane / 8 -- the distance of row 4j + g.
// One warp evaluating a whole list passes (0, 1); the warps of a cooperative CTA pass (warp, warps).
template <int DT, typename S>
__device__ __forceinline__ void eval_quads_ring(const SearchParams& p, WarpCtx& c, uint32_t m, uint32_t first, uint32_t step,
                                                S&& sink) {
    const uint32_t stages = p.nslot >> 2;
    const uint32_t nquad_all = (m + 3) >> 2;
    const uint32_t nquad = nquad_all > first ? (nquad_all - first + step - 1) / step : 0;  // this warp's quads
    // lanes 0..3 each issue one row copy of the quad (address arithmetic in parallel); lane 0 arms the barrier.
    // The barrier's pending-arrival count stays at 1 until lane 0 arrives, so complete_tx from a copy that
    // lands before the expect_tx cannot complete the phase early.
    auto issue_quad = [&](uint32_t jl, uint32_t s) {
        const uint32_t j = first + jl * step;
        const uint32_t cnt = min(4u, m - 4 * j);
        if (c.lane == 0) mbar_expect_tx(&c.bar[s], cnt * p.ix.row_bytes);
        if (c.lane < cnt) copy_row(p, c, 4 * s + c.lane, c.todo[4 * j + c.lane], &c.bar[s]);
    };
    {
        const uint32_t pre = nquad < stages ? nquad : stages;
        for (uint32_t jl = 0; jl < pre; ++jl) issue_quad(jl, jl);
    }
    const uint32_t g = c.lane >> 3;
    uint32_t s = 0;
    for (uint32_t jl = 0; jl < nquad; ++jl) {
        mbar_wait(&c.bar[s], (c.phases >> s) & 1u);
        c.phases ^= 1u << s;
        const uint32_t j = first + jl * step;
        const uint32_t cnt = min(4u, m - 4 * j);
        // all 32 lanes run the shuffles; groups beyond a partial quad recompute row 0 and are ignored
        const uint32_t gg = g < cnt ? g : 0;
        const float d = quad_distance<DT>(p, c, c.ring + (size_t)(4 * s + gg) * p.ix.row_bytes);
        __syncwarp();
        if (jl + stages < nquad) issue_quad(jl + stages, s);
        sink(j, cnt, d);
        s = (s + 1 == stages) ? 0 : s + 1;
    }
}

// The single-warp sink: `maybe(d)` is a cheap, conservative accept test evaluated by every group on its own
// row at once; only rows that pass are handed to `on_dist` (in list order), which applies the exact,
// order-dependent test.  The caller guarantees that a row failing `maybe` at the start of a step would also
// fail `on_dist`'s test later in the step (thresholds only tighten).  `tick(cnt)` runs once per step.
template <typename M, typename F, typename T>
__device__ __forceinline__ void consume_quad(const WarpCtx& c, uint32_t j, uint32_t cnt, float d, M&& maybe, F&& on_dist,
                                             T&& tick) {
    tick(cnt);
    uint32_t mask = __ballot_sync(FULL_MASK, (c.lane & 7u) == 0 && (c.lane >> 3) < cnt && maybe(d));
    while (mask) {
        const uint32_t src = __ffs(mask) - 1;
        mask &= mask - 1;
        const float de = __shfl_sync(FULL_MASK, d, src);
        on_dist(c.todo[4 * j + (src >> 3)], de);
    }
}

template <int DT, typename M, typename F, typename T>
__device__ __forceinline__ void eval_list_quad(const SearchParams& p, WarpCtx& c, uint32_t m, M&& maybe, F&& on_dist,
                                               T&& tick) {
    eval_quads_ring<DT>(p, c, m, 0, 1, [&](uint32_t j, uint32_t cnt, float d) { consume_quad(c, j, cnt, d, maybe, on_dist, tick); });
}

// Evaluates the distances of c.todo[0..m) in order, with up to nslot row fetches in flight.
template <int DT, typename M, typename F, typename T>
__device__ __forceinline__ void eval_list_single(const SearchParams& p, WarpCtx& c, uint32_t m, M&& maybe, F&& on_dist,
                                                 T&& tick) {
    const uint32_t nslot = p.nslot;
    if (c.lane == 0) {
        uint32_t pre = m < nslot ? m : nslot;
        for (uint32_t i = 0; i < pre; ++i) issue_row(p, c, i, c.todo[i]);
    }
    uint32_t slot = 0;
    for (uint32_t i = 0; i < m; ++i) {
        mbar_wait(&c.bar[slot], (c.phases >> slot) & 1u);
        c.phases ^= 1u << slot;
        const uint32_t id = c.todo[i];
        const float d = row_distance<DT>(p, c, c.ring + (size_t)slot * p.ix.row_bytes);
        __syncwarp();  // every lane is done reading the slot before it is refilled
        if (c.lane == 0 && i + nslot < m) issue_row(p, c, slot, c.todo[i + nslot]);
        tick(1u);
        if (maybe(d)) on_dist(id, d);  // d is warp-uniform
        slot = (slot + 1 == nslot) ? 0 : slot + 1;
    }
}

// Packed-bit rows of at most 1024 bits (128 bytes).  A row is 8 lanes x 16 bytes, so one warp-wide instruction moves
// four rows.  The whole neighbour list (up to 64 rows, 8 KB) is requested at once with per-lane 16-byte async copies
// into the ring -- a lane copies exactly the chunk it later reads, so no barrier object and no cross-lane hand-off --
// in two commit groups: the first half is evaluated while the second is still landing.  One DRAM round trip per
// expansion instead of one per 16 rows (round 1: the binary config sat at 0.13 of the HBM peak, latency bound on four
// dependent rounds of row loads).  Integer sums: order free; accepts happen in list order (consume_quad).
template <typename M, typename F, typename T>
__device__ __forceinline__ void eval_list_bits(const SearchParams& p, WarpCtx& c, uint32_t m, M&& maybe, F&& on_dist,
                                               T&& tick) {
    const uint32_t g = c.lane >> 3, t = c.lane & 7;
    const uint32_t words4 = p.ix.dim >> 7;  // uint4 per row, <= 8
    const uint4* qw = reinterpret_cast<const uint4*>(c.q);
    const uint4 y = t < words4 ? qw[t] : make_uint4(0, 0, 0, 0);
    const uint32_t cap = p.nslot & ~3u;  // rows the ring holds (multiple of 4, >= 8)
    for (uint32_t base = 0; base < m; base += cap) {
        const uint32_t cnt_rows = min(cap, m - base);
        const uint32_t nquad = (cnt_rows + 3) >> 2, half = (nquad + 1) >> 1;
        auto issue = [&](uint32_t j) {
            const uint32_t idx = base + 4 * j + g;
            if (idx < m && t < words4)
                cp_async16(c.ring + (size_t)(4 * j + g) * p.ix.row_bytes + t * 16,
                           p.ix.vecs + (size_t)c.todo[idx] * p.ix.row_bytes + t * 16);
        };
        for (uint32_t j = 0; j < half; ++j) issue(j);
        cp_async_commit();
        for (uint32_t j = half; j < nquad; ++j) issue(j);
        cp_async_commit();
        for (uint32_t j = 0; j < nquad; ++j) {
            if (j == 0) cp_async_wait<1>();
            if (j == half) cp_async_wait<0>();
            const uint32_t rows = min(4u, cnt_rows - 4 * j);
            uint32_t v = 0;
            if (g < rows && t < words4) {
                const uint4 x = *reinterpret_cast<const uint4*>(c.ring + (size_t)(4 * j + g) * p.ix.row_bytes + t * 16);
                v = __popc(x.x ^ y.x) + __popc(x.y ^ y.y) + __popc(x.z ^ y.z) + __popc(x.w ^ y.w);
            }
            v += __shfl_xor_sync(FULL_MASK, v, 1);
            v += __shfl_xor_sync(FULL_MASK, v, 2);
            v += __shfl_xor_sync(FULL_MASK, v, 4);
            // consume_quad indexes c.todo[4 * quad + row]: hand it the list-relative quad index
            consume_quad(c, (base >> 2) + j, rows, (float)v, maybe, on_dist, tick);
        }
        __syncwarp();  // every lane is done with the ring before the next chunk overwrites it
    }
}

// SQ8 rows of at most 128 * QN bytes.  Group g = lane / 8 owns one row of the step, lane t = lane % 8 owns its
// 16-byte chunks t, t + 8, ...: the lane copies exactly the bytes it later reads, so the ring is a per-lane
// staging buffer filled by 16-byte async copies (no barrier object, no cross-lane hand-off) and the query's
// chunks stay in registers for the whole query (`qreg`).  Stage s uses slots 4s..4s+3; one commit group per
// step keeps the wait depth constant.
template <int QN, typename S>
__device__ __forceinline__ void eval_quads_sq8(const SearchParams& p, WarpCtx& c, uint32_t m, const uint4 (&qreg)[QN > 0 ? QN : 1],
                                               uint32_t first, uint32_t step, S&& sink) {
    const uint32_t stages = p.nslot >> 2;  // 2..4
    const uint32_t nquad_all = (m + 3) >> 2;
    const uint32_t nquad = nquad_all > first ? (nquad_all - first + step - 1) / step : 0;  // this warp's quads
    const uint32_t g = c.lane >> 3, t = c.lane & 7;
    const uint32_t n16 = p.ix.row_bytes >> 4;
    auto issue = [&](uint32_t jl, uint32_t s) {
        const uint32_t j = first + jl * step;
        if (4 * j + g < m) {
            const uint8_t* src = p.ix.vecs + (size_t)c.todo[4 * j + g] * p.ix.row_bytes + t * 16;
            uint8_t* dst = c.ring + (size_t)(4 * s + g) * p.ix.row_bytes + t * 16;
#pragma unroll
            for (int u = 0; u < QN; ++u)
                if (t + 8 * u < n16) cp_async16(dst + 128 * u, src + 128 * u);
        }
    };
    for (uint32_t jl = 0; jl < stages; ++jl) {
        if (jl < nquad) issue(jl, jl);
        cp_async_commit();
    }
    uint32_t s = 0;
    for (uint32_t jl = 0; jl < nquad; ++jl) {
        if (stages == 4)
            cp_async_wait<3>();
        else if (stages == 3)
            cp_async_wait<2>();
        else
            cp_async_wait<1>();
        const uint32_t j = first + jl * step;
        const uint32_t cnt = min(4u, m - 4 * j);
        uint32_t d = 0;
        if (g < cnt) {
            const uint8_t* row = c.ring + (size_t)(4 * s + g) * p.ix.row_bytes + t * 16;
            uint4 x[QN > 0 ? QN : 1];
#pragma unroll
            for (int u = 0; u < QN; ++u)
                x[u] = (t + 8 * u < n16) ? *reinterpret_cast<const uint4*>(row + 128 * u) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < QN; ++u) {
                d = sq8_word(x[u].x, qreg[u].x, d);
                d = sq8_word(x[u].y, qreg[u].y, d);
                d = sq8_word(x[u].z, qreg[u].z, d);
                d = sq8_word(x[u].w, qreg[u].w, d);
            }
        }
        if (jl + stages < nquad) issue(jl + stages, s);  // the lane is done with its own chunks of this stage
        cp_async_commit();
        d += __shfl_xor_sync(FULL_MASK, d, 1);
        d += __shfl_xor_sync(FULL_MASK, d, 2);
        d += __shfl_xor_sync(FULL_MASK, d, 4);
        sink(j, cnt, __uint_as_float(d));
        s = (s + 1 == stages) ? 0 : s + 1;
    }
    cp_async_wait<0>();  // only empty groups can be pending here; keeps the group count clean for the next list
}

template <int QN, typename M, typename F, typename T>
__device__ __forceinline__ void eval_list_sq8(const SearchParams& p, WarpCtx& c, uint32_t m, const uint4 (&qreg)[QN > 0 ? QN : 1],
                                              M&& maybe, F&& on_dist, T&& tick) {
    eval_quads_sq8<QN>(p, c, m, qreg, 0, 1, [&](uint32_t j, uint32_t cnt, float d) { consume_quad(c, j, cnt, d, maybe, on_dist, tick); });
}

template <int DT, int QN, typename M, typename F, typename T>
__device__ __forceinline__ void eval_list(const SearchParams& p, WarpCtx& c, uint32_t m, const uint4 (&qreg)[QN > 0 ? QN : 1],
                                          M&& maybe, F&& on_dist, T&& tick) {
    if (QN > 0)
        eval_list_sq8<QN>(p, c, m, qreg, maybe, on_dist, tick);
    else if (DT == VELES_BIN1 && p.quad == 2)
        eval_list_bits(p, c, m, maybe, on_dist, tick);
    else if (p.quad)
        eval_list_quad<DT>(p, c, m, maybe, on_dist, tick);
    else
        eval_list_single<DT>(p, c, m, maybe, on_dist, tick);
}

// One 32-id chunk of an adjacency row (lane holds `nid`): optional visited test-and-set, then ordered
// compaction into c.todo.  Returns false once the row's INVALID padding was reached.
template <bool FILTER>
__device__ __forceinline__ bool gather_chunk(WarpCtx& c, uint32_t nid, uint32_t* vis, uint32_t* vlog, uint32_t logn,
                                             uint32_t& m, uint32_t& read) {
    const bool valid = nid != VELES_INVALID_ID;
    bool keep = valid;
    if (FILTER && valid) {
        const uint32_t bit = 1u << (nid & 31);
        keep = (atomicOr(&vis[nid >> 5], bit) & bit) == 0;
    }
    const uint32_t vmask = __ballot_sync(FULL_MASK, valid);
    const uint32_t kmask = __ballot_sync(FULL_MASK, keep);
    if (keep) {
        const uint32_t pos = m + __popc(kmask & ((1u << c.lane) - 1u));
        c.todo[pos] = nid;
        if (FILTER && logn + pos < kLogCap) vlog[logn + pos] = nid;
    }
    m += __popc(kmask);
    read += __popc(vmask);
    return vmask == FULL_MASK;
}

// Reads an adjacency row (padded with INVALID) into c.todo, optionally filtering through the
// visited bitmap.  Returns the number of ids kept; `read` gets the number of valid ids in the row.
template <bool FILTER>
__device__ __forceinline__ uint32_t gather_row(const SearchParams& p, WarpCtx& c, const uint32_t* __restrict__ row,
                                               uint32_t stride, uint32_t* vis, uint32_t* vlog, uint32_t& logn,
                                               uint32_t& read) {
    uint32_t m = 0;
    read = 0;
    for (uint32_t base = 0; base < stride; base += 32)
        if (!gather_chunk<FILTER>(c, row[base + c.lane], vis, vlog, logn, m, read)) break;
    if (FILTER) logn += m;
    __syncwarp();
    return m;
}

// Same as gather_chunk<true> but with the visited word already known (`word`): no waiting on the atomic.
__device__ __forceinline__ void gather_peeked(WarpCtx& c, uint32_t nid, uint32_t word, uint32_t* vis, uint32_t* vlog,
                                              uint32_t logn, uint32_t& m, uint32_t& read) {
    const bool valid = nid != VELES_INVALID_ID;
    const uint32_t bit = 1u << (nid & 31);
    const bool keep = valid && (word & bit) == 0;
    if (keep) atomicOr(&vis[nid >> 5], bit);  // result unused: compiles to a fire-and-forget RED
    const uint32_t vmask = __ballot_sync(FULL_MASK, valid);
    const uint32_t kmask = __ballot_sync(FULL_MASK, keep);
    if (keep) {
        const uint32_t pos = m + __popc(kmask & ((1u << c.lane) - 1u));
        c.todo[pos] = nid;
        if (logn + pos < kLogCap) vlog[logn + pos] = nid;
    }
    m += __popc(kmask);
    read += __popc(vmask);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }


// The ef-bounded result set (graph.rs:450) as one array sorted by key.  R == 0: in shared memory (any
// ef).  R > 0: in registers, R keys per lane (ef <= 32*R): position i lives in lane i % 32, slot i / 32;
// an insert is a ballot (lower bound) plus one shuffle-up per slot instead of a shared-memory shift.
template <int R>
struct ResArr {
    uint64_t k[R > 0 ? R : 1];
    uint64_t* sm;
    uint32_t lane;
    __device__ __forceinline__ void init(uint64_t* smem_ptr, uint32_t lane_) {
        sm = smem_ptr;
        lane = lane_;
        if (R > 0) {
#pragma unroll
            for (int s = 0; s < R; ++s) k[s] = ~0ull;
        }
    }
    __device__ __forceinline__ uint64_t get(uint32_t i) const {
        if (R == 0) return sm[i];
        uint64_t v = 0;
#pragma unroll
        for (int s = 0; s < R; ++s) {
            const uint64_t t = __shfl_sync(FULL_MASK, k[s], i & 31);
            v = ((int)(i >> 5) == s) ? t : v;  // select, never an indexed access: keeps k[] in registers
        }
        return v;
    }
    __device__ __forceinline__ void set(uint32_t i, uint64_t key) {
        if (R == 0) {
            __syncwarp();
            if (lane == 0) sm[i] = key;
            __syncwarp();
            return;
        }
#pragma unroll
        for (int s = 0; s < R; ++s) k[s] = ((int)(i >> 5) == s && lane == (i & 31)) ? key : k[s];
    }
    // first position in (after, len) whose expanded flag (bit 0) is clear, else len
    __device__ __forceinline__ uint32_t next_unexpanded(uint32_t after_excl, bool from_start, uint32_t len) const {
        const uint32_t lo = from_start ? 0 : after_excl + 1;
        if (R == 0) {
            for (uint32_t base = lo & ~31u; base < len; base += 32) {
                const uint32_t i = base + lane;
                const bool un = i < len && i >= lo && (sm[i] & 1ull) == 0;
                const uint32_t msk = __ballot_sync(FULL_MASK, un);
                if (msk) return base + __ffs(msk) - 1;
            }
            return len;
        }
        uint32_t found = len;
#pragma unroll
        for (int s = R - 1; s >= 0; --s) {
            const uint32_t i = 32u * s + lane;
            const bool un = i < len && i >= lo && (k[s] & 1ull) == 0;
            const uint32_t msk = __ballot_sync(FULL_MASK, un);
            if (msk) found = 32u * s + __ffs(msk) - 1;
        }
        return found;
    }
    __device__ __forceinline__ uint32_t lower_bound(uint32_t len, uint64_t key) const {
        if (R == 0) return lower_bound_warp(sm, len, key, lane);
        uint32_t p = 0;
#pragma unroll
        for (int s = 0; s < R; ++s) p += __popc(__ballot_sync(FULL_MASK, k[s] < key));  // empty slots hold ~0
        return p;
    }
    // insert key at pos; positions >= cap fall off (cap = ef)
    __device__ __forceinline__ void insert(uint32_t pos, uint32_t new_len, uint32_t cap, uint64_t key) {
        if (R == 0) {
            insert_at(sm, pos, new_len, key, lane);
            return;
        }
#pragma unroll
        for (int s = R - 1; s >= 0; --s) {
            uint64_t up = __shfl_up_sync(FULL_MASK, k[s], 1);
            if (s > 0) {
                const uint64_t carry = __shfl_sync(FULL_MASK, k[s - 1], 31);
                if (lane == 0) up = carry;
            }
            const uint32_t i = 32u * s + lane;
            uint64_t nv = i > pos ? up : k[s];
            nv = i == pos ? key : nv;
            nv = i >= cap ? ~0ull : nv;
            k[s] = nv;
        }
    }
};

// COOP = false: one warp per query (CTA of 32 threads).  COOP = true: a CTA of 2..8 warps per query -- warp 0
// (the leader) owns the result array and does everything order-dependent (pops, adjacency + visited filter,
// inserts); the distance evaluations of each neighbour list are split across the warps, quad by quad, every
// warp with its own ring, and handed back through shared memory.  Same arithmetic per row, same order of
// inserts, so the results are identical; it exists for batches too small to fill the GPU with one warp per
// query, where a query's latency is its single warp's instruction chain.
template <int DT, int R, int QN, bool COOP>
__global__ void __launch_bounds__(COOP ? 256 : 32, (DT == VELES_BIN1 && !COOP) ? 16 : 1) hnsw_search_kernel(const SearchParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    WarpCtx c;
    c.lane = threadIdx.x & 31;
    const uint32_t tid = threadIdx.x, nthr = COOP ? blockDim.x : 32u;
    const uint32_t warp = COOP ? threadIdx.x >> 5 : 0u, nwarp = COOP ? blockDim.x >> 5 : 1u;
    const bool leader = warp == 0;
    // layout: [COOP: control words, 16 B] results | todo | [COOP: distances] | query | per warp: barriers + ring
    volatile uint32_t* ctrl = reinterpret_cast<volatile uint32_t*>(smem);
    c.res = reinterpret_cast<uint64_t*>(smem + p.off_res);
    c.todo = reinterpret_cast<uint32_t*>(smem + p.off_todo);
    float* dists = reinterpret_cast<float*>(smem + p.off_dist);
    c.q = smem + p.off_q;
    c.bar = reinterpret_cast<uint64_t*>(smem + p.off_ring + (size_t)warp * p.warp_bytes);
    c.ring = reinterpret_cast<uint8_t*>(c.bar) + kMaxSlots * 8;
    c.phases = 0;
    c.norm_a = 0.0f;
    c.policy = make_evict_first_policy();
    const uint32_t lane = c.lane;
    ResArr<R> res;
    if (lane == 0) {
        for (uint32_t i = 0; i < kMaxSlots; ++i) mbar_init(&c.bar[i], 1);
        fence_barrier_init();
    }
    auto cta_sync = [&]() {
        if (COOP)
            __syncthreads();
        else
            __syncwarp();
    };
    // leader -> everyone, through alternating control words (a word is rewritten two barriers later at the earliest)
    uint32_t bslot = 0;
    auto bcast = [&](uint32_t v) -> uint32_t {
        if (!COOP) {
            __syncwarp();  // publishes the leader lane's shared-memory writes (todo) to the warp
            return v;
        }
        if (leader && lane == 0) ctrl[bslot] = v;
        __syncthreads();
        v = ctrl[bslot];
        bslot ^= 1u;
        return v;
    };
    cta_sync();

    uint32_t* vis = p.visited + (size_t)blockIdx.x * p.vis_words;
    uint32_t* vlog = p.vlog + (size_t)blockIdx.x * kLogCap;
    uint64_t* tie = p.tie + (size_t)blockIdx.x * kTieCap;
    const uint32_t dim = p.ix.dim;
    const uint32_t ef = p.ef;
    constexpr uint32_t kDone = 0xFFFFFFFFu;

    for (;;) {
        uint32_t qi = 0;
        if (leader) {
            if (lane == 0) qi = atomicAdd(&p.counters[0], 1u);
            qi = __shfl_sync(FULL_MASK, qi, 0);
        }
        qi = bcast(qi);
        if (qi >= p.nq) break;

        // ---- stage the query ----
        const float* qg = p.queries + (size_t)qi * dim;
        if (DT == VELES_BIN1) {
            uint32_t* qw = reinterpret_cast<uint32_t*>(c.q);
            for (uint32_t w = tid; w < (dim >> 5); w += nthr) {
                uint32_t bits = 0;
                for (uint32_t b = 0; b < 32; ++b) bits |= (qg[w * 32 + b] > 0.5f ? 1u : 0u) << b;
                qw[w] = bits;
            }
        } else if (DT == VELES_SQ8) {
            // ScalarQuantizer::quantize (quantization.rs:236-250); padding bytes are zero like the rows'
            for (uint32_t i = tid; i < p.ix.row_bytes; i += nthr) {
                uint32_t b = 0;
                if (i < dim) {
                    const float qv = roundf(__fmul_rn(__fsub_rn(qg[i], p.ix.sq_min[i]), p.ix.sq_scale[i]));
                    b = (qv != qv) ? 0u : (uint32_t)fminf(fmaxf(qv, 0.0f), 255.0f);
                }
                c.q[i] = (uint8_t)b;
            }
        } else {
            float* qs = reinterpret_cast<float*>(c.q);
            for (uint32_t i = tid; i < dim; i += nthr) qs[i] = qg[i];
        }
        cta_sync();
        uint4 qreg[QN > 0 ? QN : 1];
        if (QN > 0) {
#pragma unroll
            for (int u = 0; u < QN; ++u) {
                const uint32_t w = (lane & 7u) + 8u * u;
                qreg[u] = w < (p.ix.row_bytes >> 4) ? reinterpret_cast<const uint4*>(c.q)[w] : make_uint4(0, 0, 0, 0);
            }
        } else {
            qreg[0] = make_uint4(0, 0, 0, 0);
        }
        if ((DT == VELES_F32 || DT == VELES_F16) && p.ix.metric == VELES_COSINE) {
            const float* qs = reinterpret_cast<const float*>(c.q);
            c.norm_a = __fsqrt_rn(warp_tree_reduce<0>(qs, qs, dim, lane));
        }

        uint32_t ndc0 = 0, hops0 = 0, ndc_up = 0, hops_up = 0;
        uint32_t len = 0;
        auto always = [](float) { return true; };
        auto no_tick = [](uint32_t) {};
        res.init(c.res, lane);
        // Distances of c.todo[0..m) -> maybe / on_dist in list order (see consume_quad).  COOP: every warp calls it
        // with the same m (after a bcast, which also publishes todo); the quads are dealt round-robin, the
        // distances meet in `dists`, and the leader consumes them 32 at a time.
        auto eval = [&](uint32_t m, auto&& maybe, auto&& on_dist, auto&& tick) {
            if (!COOP) {
                eval_list<DT, QN>(p, c, m, qreg, maybe, on_dist, tick);
                return;
            }
            auto sink = [&](uint32_t j, uint32_t cnt, float d) {
                if (leader) tick(cnt);
                if ((lane & 7u) == 0 && (lane >> 3) < cnt) dists[4 * j + (lane >> 3)] = d;
            };
            if (QN > 0)
                eval_quads_sq8<QN>(p, c, m, qreg, warp, nwarp, sink);
            else
                eval_quads_ring<DT>(p, c, m, warp, nwarp, sink);
            __syncthreads();
            if (leader) {
                for (uint32_t base = 0; base < m; base += 32) {
                    const uint32_t i = base + lane;
                    const float d = i < m ? dists[i] : 0.0f;
                    uint32_t mask = __ballot_sync(FULL_MASK, i < m && maybe(d));
                    while (mask) {
                        const uint32_t src = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const float de = __shfl_sync(FULL_MASK, d, src);
                        on_dist(c.todo[base + src], de);
                    }
                }
            }
        };

        if (p.ix.has_entry) {
            // ---- greedy descent, layers max_layer..1 (graph.rs:259-263, 405-428) ----
            uint32_t cur = p.ix.entry;
            uint32_t dummy_logn = 0, nread = 0;
            for (uint32_t layer = p.ix.max_layer; layer >= 1; --layer) {
                uint32_t best = cur;
                float best_dist = 0.0f;
                if (leader && lane == 0) c.todo[0] = best;
                bcast(1u);
                eval(1u, always, [&](uint32_t, float d) { best_dist = d; }, no_tick);
                ++ndc_up;
                for (;;) {
                    uint32_t m = 0;
                    if (leader) {
                        const uint32_t ref = p.ix.upper_ref[best];
                        if (ref != VELES_INVALID_ID && layer <= (ref & 15u)) {
                            const uint32_t* row = p.ix.upper_adj + ((size_t)(ref >> 4) + layer - 1) * p.ix.strideU;
                            m = gather_row<false>(p, c, row, p.ix.strideU, nullptr, nullptr, dummy_logn, nread);
                        }
                    }
                    m = bcast(m);
                    ++hops_up;
                    ndc_up += m;
                    bool improved = false;
                    eval(
                        m, [&](float d) { return d < best_dist; },
                        [&](uint32_t id, float d) {
                            if (d < best_dist) {
                                best = id;
                                best_dist = d;
                                improved = true;
                            }
                        },
                        no_tick);
                    __syncwarp();
                    if (bcast(improved ? 1u : 0u) == 0u) break;
                }
                cur = best;  // meaningful in the leader only
            }

            // ---- layer 0 beam (graph.rs:266, 438-520) ----
            uint32_t logn = 0, tlen = 0;
            float worst = 0.0f;  // distance of res[len-1] (results.peek()), kept in a register
            uint32_t nxt = 0;  // index of the first unexpanded entry of res (== len when there is none)
            // Next-candidate prefetch (the GPU counterpart of graph.rs:480-497): as soon as a node becomes the
            // first unexpanded entry its adjacency row is loaded into registers (two ids per lane) and, a few
            // distance evaluations later, the visited-bitmap words of its neighbours are pulled into L2.  If
            // that node is still the one popped next, its expansion starts without the two dependent DRAM
            // round trips.  Purely read-only speculation: results do not change.
            const bool can_pre = p.ix.stride0 <= 64;
            uint32_t pre_node = VELES_INVALID_ID, pre_a = VELES_INVALID_ID, pre_b = VELES_INVALID_ID, pre_age = 0;
            uint32_t pre_va = 0, pre_vb = 0;  // visited words of pre_a / pre_b, peeked read-only
            bool pre_peeked = false;          // true once pre_va / pre_vb are valid for the *next* expansion
            auto learn = [&](uint32_t x) {
                if (!can_pre || x == pre_node) return;
                pre_node = x;
                const uint32_t* row = p.ix.adj0 + (size_t)x * p.ix.stride0;
                pre_a = row[lane];
                pre_b = p.ix.stride0 > 32 ? row[32 + lane] : VELES_INVALID_ID;
                pre_age = 1;
                pre_peeked = false;
            };
            // Runs once per evaluation step with the number of rows it covered: once the predicted node's adjacency
            // row has had time to land, read the visited words of its neighbours (read-only); a little later
            // they are usable.
            auto tick = [&](uint32_t rows) {
                if (pre_age == 0) return;
                pre_age += rows;
                if (pre_age < 100) {
                    if (pre_age >= 10) {
                        if (p.peek) {
                            pre_va = pre_a != VELES_INVALID_ID ? __ldcg(&vis[pre_a >> 5]) : 0u;
                            pre_vb = pre_b != VELES_INVALID_ID ? __ldcg(&vis[pre_b >> 5]) : 0u;
                            pre_age = 100;
                        } else {
                            if (pre_a != VELES_INVALID_ID) prefetch_l2(&vis[pre_a >> 5]);
                            if (pre_b != VELES_INVALID_ID) prefetch_l2(&vis[pre_b >> 5]);
                            pre_age = 0;
                        }
                    }
                } else if (pre_age >= 106) {
                    pre_peeked = true;
                    pre_age = 0;
                }
            };
            bool full = false;  // len >= ef
            // the exact, order-dependent accept step of search_layer (graph.rs:499-515) for one evaluated neighbour
            auto accept = [&](uint32_t id, float d) {
                if (d < worst || !full) {
                    const uint64_t key = make_key(d, id);
                    const uint32_t pos = res.lower_bound(len, key);
                    if (!full) {
                        res.insert(pos, len + 1, ef, key);
                        ++len;
                        full = len >= ef;
                        worst = key_dist(res.get(len - 1));
                        if (pos <= nxt) {
                            nxt = pos;
                            learn(id);
                        } else if (nxt == len - 1) {
                            // there was no unexpanded entry: the new one (at pos) is now the first
                            nxt = pos;
                            learn(id);
                        }
                    } else {
                        const uint64_t ev = res.get(len - 1);
                        __syncwarp();
                        res.insert(pos, len, ef, key);
                        const float nworst = key_dist(res.get(len - 1));
                        worst = nworst;
                        // ties that are now farther than the worst result can only end the loop: drop them
                        if (tlen > 0) {
                            uint32_t w = 0;
                            for (uint32_t base = 0; base < tlen; base += 32) {
                                const uint32_t i = base + lane;
                                uint64_t v = 0;
                                bool keep = false;
                                if (i < tlen) {
                                    v = tie[i];
                                    keep = !(key_dist(v) > nworst);
                                }
                                const uint32_t msk = __ballot_sync(FULL_MASK, keep);
                                __syncwarp();
                                if (keep) tie[w + __popc(msk & ((1u << lane) - 1u))] = v;
                                w += __popc(msk);
                                __syncwarp();
                            }
                            tlen = w;
                        }
                        if ((ev & 1ull) == 0 && !(key_dist(ev) > nworst)) {
                            if (tlen < kTieCap) {
                                if (lane == 0) tie[tlen] = ev;
                                ++tlen;
                            } else if (lane == 0) {
                                atomicExch(&p.counters[1], 1u);
                            }
                            __syncwarp();
                        }
                        // index of the first unexpanded entry after the shift (the last entry fell off)
                        if (pos <= nxt) {
                            nxt = pos;
                            learn(id);
                        } else if (nxt >= len) {
                            nxt = pos;
                            learn(id);
                        }
                    }
                }
            };
            {
                // entry points of the layer-0 search: the greedy descent's result, plus the caller's extra probes
                // for NativeHnsw::search_multi_entry (graph.rs:288-348), duplicates dropped as `contains` does
                uint32_t m0 = 1;
                if (leader) {
                    if (lane == 0) {
                        c.todo[0] = cur;
                        if (p.extra_entries) {
                            for (uint32_t e = 0; e < 3; ++e) {
                                const uint32_t x = p.extra_entries[(size_t)qi * 3 + e];
                                bool dup = x == VELES_INVALID_ID || x >= p.ix.n;
                                for (uint32_t i = 0; i < m0 && !dup; ++i) dup = c.todo[i] == x;
                                if (!dup) c.todo[m0++] = x;
                            }
                        }
                        for (uint32_t i = 0; i < m0; ++i) {
                            const uint32_t x = c.todo[i];
                            atomicOr(&vis[x >> 5], 1u << (x & 31));
                            vlog[i] = x;
                        }
                    }
                    m0 = __shfl_sync(FULL_MASK, m0, 0);
                }
                m0 = bcast(m0);
                logn = m0;
                ndc0 += m0;
                if (m0 == 1) {
                    float d0 = 0.0f;
                    eval(1u, always, [&](uint32_t, float d) { d0 = d; }, no_tick);
                    if (leader) {
                        res.set(0, make_key(d0, cur));
                        len = 1;
                        full = len >= ef;
                        worst = d0;
                        __syncwarp();
                    }
                } else {
                    eval(m0, always, accept, no_tick);  // every entry enters both heaps (graph.rs:452-458); host checks ef >= 4
                    __syncwarp();
                }
            }
            for (;;) {
                uint32_t m = 0;
                bool have = true;
                if (leader) {
                    // pop the closest candidate: first unexpanded entry of res, else the smallest tie
                    uint32_t cnode = VELES_INVALID_ID;
                    if (nxt < len) {
                        const uint64_t key = res.get(nxt);
                        cnode = key_id(key);
                        res.set(nxt, key | 1ull);
                        nxt = res.next_unexpanded(nxt, false, len);  // the following unexpanded entry
                    } else if (tlen > 0) {
                        // every tie has dist == worst result dist: popped without the break (graph.rs:474)
                        uint64_t best = ~0ull;
                        for (uint32_t i = lane; i < tlen; i += 32) {
                            const uint64_t v = tie[i];
                            best = v < best ? v : best;
                        }
                        best = warp_min_u64(best);
                        cnode = key_id(best);
                        // remove it: move the last entry into its place
                        const uint64_t lastv = tie[tlen - 1];
                        __syncwarp();
                        for (uint32_t i = lane; i < tlen; i += 32)
                            if (tie[i] == best) tie[i] = lastv;
                        --tlen;
                        __syncwarp();
                    } else {
                        have = false;  // candidates exhausted, or everything left is farther than the worst result
                    }
                    // expand cnode: adjacency from the prefetch registers when the prediction held
                    uint32_t nread = 0;
                    if (!have) {
                    } else if (can_pre && cnode == pre_node && pre_peeked) {
                        // The visited words were read after the previous expansion's marking and nothing has been
                        // marked since: the peek is exact.  Mark now without waiting for the atomics' results.
                        gather_peeked(c, pre_a, pre_va, vis, vlog, logn, m, nread);
                        if (p.ix.stride0 > 32) gather_peeked(c, pre_b, pre_vb, vis, vlog, logn, m, nread);
                        logn += m;
                        __syncwarp();
                    } else if (can_pre && cnode == pre_node) {
                        if (gather_chunk<true>(c, pre_a, vis, vlog, logn, m, nread) && p.ix.stride0 > 32)
                            gather_chunk<true>(c, pre_b, vis, vlog, logn, m, nread);
                        logn += m;
                        __syncwarp();
                    } else {
                        m = gather_row<true>(p, c, p.ix.adj0 + (size_t)cnode * p.ix.stride0, p.ix.stride0, vis, vlog, logn, nread);
                    }
                }
                m = bcast(leader ? (have ? m : kDone) : 0u);
                if (m == kDone) break;
                ++hops0;
                ndc0 += m;
                if (leader) {
                    pre_peeked = false;  // this expansion's marking invalidates any earlier peek
                    if (nxt < len) learn(key_id(res.get(nxt)));
                }
                eval(m, [&](float d) { return d < worst || !full; }, accept, tick);
                __syncwarp();
            }

            if (leader) {
                // ---- clear the visited bitmap for the next query of this slot ----
                if (logn <= kLogCap) {
                    // eight independent log reads in flight per lane (one at a time made this loop ~8% of a query)
                    for (uint32_t i = lane; i < logn; i += 32 * 8) {
                        uint32_t v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) v[u] = i + 32u * u < logn ? __ldcg(&vlog[i + 32u * u]) : VELES_INVALID_ID;
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (v[u] != VELES_INVALID_ID) vis[v[u] >> 5] = 0u;
                    }
                } else {
                    for (uint32_t i = lane; i < p.vis_words; i += 32) vis[i] = 0u;
                }
                __syncwarp();
            }
        }

        if (leader) {
            // ---- write the first k results (graph.rs:269) ----
            const uint32_t cnt = len < p.k ? len : p.k;
            auto emit = [&](uint32_t i, uint64_t key) {
                if (i < p.k) {
                    const bool ok = i < cnt;
                    const uint32_t id = ok ? key_id(key) : VELES_INVALID_ID;
                    const float d = ok ? key_dist(key) : __uint_as_float(0x7fc00000u);
                    p.out_ids[(size_t)qi * p.k + i] = id;
                    p.out_dist[(size_t)qi * p.k + i] = d;
                    for (uint32_t r = 0; r < p.n_peers; ++r) {
                        p.peer_ids[r][(size_t)qi * p.k + i] = id;
                        p.peer_dist[r][(size_t)qi * p.k + i] = d;
                    }
                }
            };
            if (R == 0) {
                for (uint32_t i = lane; i < p.k; i += 32) emit(i, i < cnt ? c.res[i] : 0ull);
            } else {
                // static slot indices only: position 32*s + lane lives in this lane's k[s]
#pragma unroll
                for (int s2 = 0; s2 < (R > 0 ? R : 1); ++s2) emit(32u * s2 + lane, res.k[s2]);
                for (uint32_t i = 32u * R + lane; i < p.k; i += 32) emit(i, 0ull);
            }
            if (lane == 0) {
                p.out_counts[qi] = cnt;
                for (uint32_t r = 0; r < p.n_peers; ++r) p.peer_cnt[r][qi] = cnt;
                if (p.out_stats) {
                    p.out_stats[(size_t)qi * 4 + 0] = ndc0;
                    p.out_stats[(size_t)qi * 4 + 1] = hops0;
                    p.out_stats[(size_t)qi * 4 + 2] = ndc_up;
                    p.out_stats[(size_t)qi * 4 + 3] = hops_up;
                }
            }
            __syncwarp();
        }
    }
}

using SearchKernel = void (*)(const SearchParams);
// one translation unit per storage type instantiates its kernels (hnsw_search_<type>.cu); reg_mode is the
// result-array variant (2 or 8 keys per lane in registers, 0 = shared memory), qn the SQ8 chunk count
SearchKernel search_kernel_f32(uint32_t reg_mode, bool coop);
SearchKernel search_kernel_f16(uint32_t reg_mode, bool coop);
SearchKernel search_kernel_bin1(uint32_t reg_mode);
SearchKernel search_kernel_sq8(uint32_t reg_mode, uint32_t qn, bool coop);
SearchKernel search_kernel_sq8_a(uint32_t reg_mode, uint32_t qn, bool coop);  // qn 0, 2
SearchKernel search_kernel_sq8_b(uint32_t reg_mode, uint32_t qn, bool coop);  // qn 4, 6
SearchKernel search_kernel_sq8_c(uint32_t reg_mode, uint32_t qn, bool coop);  // qn 8

// coop: the multi-warp-per-query variant (result array in registers only: ef <= 256)
#define VELES_PICK_KERNEL(DT, QN)                                                                                          \
    (coop ? (reg_mode == 2 ? hnsw_search_kernel<DT, 2, QN, true> : hnsw_search_kernel<DT, 8, QN, true>)                   \
          : (reg_mode == 2 ? hnsw_search_kernel<DT, 2, QN, false>                                                         \
                           : reg_mode == 8 ? hnsw_search_kernel<DT, 8, QN, false> : hnsw_search_kernel<DT, 0, QN, false>))

}  // namespace veles

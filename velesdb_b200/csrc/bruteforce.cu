// bruteforce.cu -- exact scan: HnswIndex::search_brute_force / brute_force_search_parallel
// (index/hnsw/index/search.rs:176-219, batch.rs:223-244), the re-rank step of search_with_rerank
// (search.rs:118-160) and DistanceEngine::batch_distance (native/distance.rs:22-24).
//
// Kernel 1 (bf_tile_kernel) computes the metric value of every (query, row) pair with the
// reference's accumulation tree, register-tiled: a warp owns RT rows x QT queries, lane l owns the
// elements i = l (mod 32) of every row and query, so each global row element is loaded once per
// QT queries and each shared-memory query element once per RT rows.  The 32 partial sums of a
// pair are combined with the xor-8/16/4/2/1 butterfly (common.cuh) -> same bits as the CPU.
// Kernel 2 (topk_kernel) selects, per query, the k best by DistanceMetric::sort_results order
// (core/distance.rs:95-103; ties by ascending node id) with a threshold filter over the score row.
#include <algorithm>
#include <type_traits>
#include <cstdlib>

#include "index.hpp"

namespace veles {

// Where a score kernel's results go.  STORE: the [nq, ld] matrix of metric values.  FILTER (thr != NULL): no matrix --
// a result is kept only if its sort key (order(score) << 32 | row, smaller = better in DistanceMetric::sort_results
// order) does not exceed the query's bound, and is appended to the query's candidate list.  The bound is the k-th best
// key of a strided sample of the collection (any subset's k-th best is no better than the collection's), so the true
// top k always pass; see bruteforce_device.
struct ScoreSink {
    float* scores = nullptr;
    uint64_t ld = 0;
    const uint32_t* thr = nullptr;  // per query: order bits of the bound
    uint32_t* cnt = nullptr;        // per query candidate count
    uint64_t* cand = nullptr;       // nq x cap keys
    uint32_t cap = 0;
    uint32_t* err = nullptr;        // [0] set when a list overflowed
    uint32_t desc = 0;
    __host__ __device__ ScoreSink at(uint32_t q0) const {
        ScoreSink o = *this;
        if (scores) o.scores = scores + (size_t)q0 * ld;
        if (thr) {
            o.thr = thr + q0;
            o.cnt = cnt + q0;
            o.cand = cand + (size_t)q0 * cap;
        }
        return o;
    }
#ifdef __CUDACC__
    __device__ __forceinline__ void emit(uint32_t q, uint64_t row, float v) const {
        if (!thr) {
            scores[(size_t)q * ld + row] = v;
            return;
        }
        uint32_t o = ord_key(v);
        if (desc) o = ~o;
        if (o <= thr[q]) {
            const uint32_t slot = atomicAdd(&cnt[q], 1u);
            if (slot < cap)
                cand[(size_t)q * cap + slot] = ((uint64_t)o << 32) | (uint32_t)row;
            else
                atomicExch(err, 1u);
        }
    }
#endif
};

constexpr int kRT = 4;   // rows per warp tile
constexpr int kQT = 8;   // queries per warp tile
constexpr int kWarps = 8;

// scores[q * n + r] = metric value (as_value) or in-graph distance of query q vs row r.
// Fast path: F32/F16 rows, COSINE / EUCLIDEAN / DOT, dim >= 16.
template <typename TB>
__global__ void __launch_bounds__(kWarps * 32) bf_tile_kernel(IndexView ix, const float* __restrict__ queries,
                                                              uint32_t nq, uint32_t q0, ScoreSink out,
                                                              bool as_value) {
    extern __shared__ __align__(16) float qs[];  // kQT x dim, then kQT norms
    const uint32_t dim = ix.dim;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t qbase = q0 + blockIdx.y * kQT;
    const uint32_t nqt = min((uint32_t)kQT, q0 + nq - qbase);
    float* qnorm = qs + (size_t)kQT * dim;
    for (uint32_t i = threadIdx.x; i < kQT * dim; i += blockDim.x) {
        uint32_t t = i / dim;
        qs[i] = t < nqt ? queries[(size_t)(qbase - q0 + t) * dim + (i - t * dim)] : 0.0f;
    }
    __syncthreads();
    if (ix.metric == VELES_COSINE) {
        for (uint32_t t = warp; t < kQT; t += kWarps) {
            const float* q = qs + (size_t)t * dim;
            float s = warp_tree_reduce<0>(q, q, dim, lane);
            if (lane == 0) qnorm[t] = __fsqrt_rn(s);
        }
    }
    __syncthreads();
    const uint32_t main_len = dim & ~31u;
    const bool has_tail = main_len != dim;
    const bool l2 = ix.metric == VELES_EUCLIDEAN;
    const uint64_t n = ix.n;
    const uint64_t tiles = (n + kRT - 1) / kRT;
    for (uint64_t tile = blockIdx.x * (uint64_t)kWarps + warp; tile < tiles; tile += (uint64_t)gridDim.x * kWarps) {
        const uint64_t r0 = tile * kRT;
        // rows past the end alias the last row; their results are simply not stored
        const uint8_t* base = ix.vecs + r0 * ix.row_bytes;
        const uint64_t last_off = (n - 1 - r0) * (uint64_t)ix.row_bytes;
        float acc[kRT][kQT];
#pragma unroll
        for (int r = 0; r < kRT; ++r)
#pragma unroll
            for (int t = 0; t < kQT; ++t) acc[r][t] = 0.0f;
        for (uint32_t i = lane; i < main_len; i += 32) {
            float x[kRT], q[kQT];
#pragma unroll
            for (int r = 0; r < kRT; ++r) {
                const uint64_t off = min((uint64_t)r * ix.row_bytes, last_off);
                x[r] = load_elem(reinterpret_cast<const TB*>(base + off), i);
            }
#pragma unroll
            for (int t = 0; t < kQT; ++t) q[t] = qs[(size_t)t * dim + i];
            if (l2) {
#pragma unroll
                for (int r = 0; r < kRT; ++r)
#pragma unroll
                    for (int t = 0; t < kQT; ++t) {
                        float d = __fsub_rn(q[t], x[r]);
                        acc[r][t] = __fmaf_rn(d, d, acc[r][t]);
                    }
            } else {
#pragma unroll
                for (int r = 0; r < kRT; ++r)
#pragma unroll
                    for (int t = 0; t < kQT; ++t) acc[r][t] = __fmaf_rn(q[t], x[r], acc[r][t]);
            }
        }
#pragma unroll
        for (int r = 0; r < kRT; ++r) {
            const bool row_ok = r0 + r < n;
            const uint64_t off = min((uint64_t)r * ix.row_bytes, last_off);
            const TB* rowp = reinterpret_cast<const TB*>(base + off);
            float nb = 0.0f;
            if (ix.metric == VELES_COSINE) nb = *reinterpret_cast<const float*>(base + off + ix.norm_off);
#pragma unroll
            for (int t = 0; t < kQT; ++t) {
                float s = warp_tree_sum32(acc[r][t]);
                const float* q = qs + (size_t)t * dim;
                float v;
                if (l2) {
                    if (has_tail) s = warp_tree_tail<1>(s, q, rowp, dim, lane);
                    v = __fsqrt_rn(s);
                } else {
                    if (has_tail) s = warp_tree_tail<0>(s, q, rowp, dim, lane);
                    if (ix.metric == VELES_COSINE) {
                        float sim = cosine_from_parts(s, qnorm[t], nb);
                        v = as_value ? sim : __fsub_rn(1.0f, sim);
                    } else {
                        v = as_value ? s : -s;
                    }
                }
                if (lane == 0 && row_ok && t < (int)nqt) out.emit(qbase - q0 + t, r0 + r, v);
            }
        }
    }
}

// ---- 8 rows x 8 queries per warp, dim % 32 == 0 ------------------------------------------------------
// Same lane <-> element-class mapping as above, but the 64 accumulators of a warp are reduced with a
// *transposing* butterfly: at the xor-8 step a lane keeps half of its accumulators and hands the other half
// to its partner, and so on (xor 16, 4, 2, 1), so 32 accumulators cost 31 shuffles instead of 160 and end
// up one per lane.  Every addition still combines the same commutative pair of partial sums as
// warp_tree_sum32, so the bits are unchanged.  Lane l ends up owning accumulator
// a(l) = b0 + 2*b1 + 4*b2 + 8*b4 + 16*b3 (b_i = bit i of l).
template <int N>
__device__ __forceinline__ void butterfly_halve(float* v, uint32_t lane, uint32_t bit_mask, uint32_t xor_mask) {
    const bool hi = (lane & bit_mask) != 0;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const float mine = hi ? v[N / 2 + i] : v[i];
        const float theirs = hi ? v[i] : v[N / 2 + i];
        v[i] = __fadd_rn(mine, __shfl_xor_sync(FULL_MASK, theirs, xor_mask));
    }
}
__device__ __forceinline__ float butterfly32(float* v, uint32_t lane) {
    butterfly_halve<32>(v, lane, 8, 8);
    butterfly_halve<16>(v, lane, 16, 16);
    butterfly_halve<8>(v, lane, 4, 4);
    butterfly_halve<4>(v, lane, 2, 2);
    butterfly_halve<2>(v, lane, 1, 1);
    return v[0];
}

constexpr int kRT8 = 8;
constexpr int kStepU = 2;  // steps per pointer advance
// QW query tiles (of kQT queries) per CTA: warp w works on query tile w % QW and row tile w / QW, so the QW
// warps that share a row tile read the same global addresses at about the same time -- one L2 read, QW - 1
// L1 hits.  Only QW = 1 is launched (see launch_scores).
template <typename TB, int QW>
__global__ void __launch_bounds__(kWarps * 32, 2) bf_tile8_kernel(IndexView ix, const float* __restrict__ queries, uint32_t nq,
                                                               ScoreSink out, bool as_value) {
    extern __shared__ __align__(16) float qs_all[];  // QW x kQT x dim, then QW x kQT norms
    const uint32_t dim = ix.dim;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cta_q0 = blockIdx.y * (kQT * QW);
    const uint32_t cta_nq = min((uint32_t)(kQT * QW), nq - cta_q0);
    float* qnorm_all = qs_all + (size_t)QW * kQT * dim;
    for (uint32_t i = threadIdx.x; i < QW * kQT * dim; i += blockDim.x) {
        const uint32_t t = i / dim;
        qs_all[i] = t < cta_nq ? queries[(size_t)(cta_q0 + t) * dim + (i - t * dim)] : 0.0f;
    }
    __syncthreads();
    if (ix.metric == VELES_COSINE) {
        for (uint32_t t = warp; t < QW * kQT; t += kWarps) {
            const float* q = qs_all + (size_t)t * dim;
            const float s = warp_tree_reduce<0>(q, q, dim, lane);
            if (lane == 0) qnorm_all[t] = __fsqrt_rn(s);
        }
    }
    __syncthreads();
    const uint32_t qtile = warp % QW, rslot = warp / QW;
    constexpr uint32_t kRowWarps = kWarps / QW;
    const float* qs = qs_all + (size_t)qtile * kQT * dim;
    const float* qnorm = qnorm_all + qtile * kQT;
    const uint32_t qbase = cta_q0 + qtile * kQT;
    const uint32_t nqt = qbase < nq ? min((uint32_t)kQT, nq - qbase) : 0u;
    const bool l2 = ix.metric == VELES_EUCLIDEAN;
    const uint64_t n = ix.n;
    const uint64_t tiles = (n + kRT8 - 1) / kRT8;
    // accumulator owned by this lane after the butterfly: row a/8 (within a half-tile of 4 rows), query a%8
    const uint32_t a = (lane & 1) | (lane & 2) | (lane & 4) | ((lane & 16) >> 1) | ((lane & 8) << 1);
    const uint32_t my_r = a >> 3, my_t = a & 7;
    for (uint64_t tile = blockIdx.x * (uint64_t)kRowWarps + rslot; tile < tiles; tile += (uint64_t)gridDim.x * kRowWarps) {
        const uint64_t r0 = tile * kRT8;
        const uint8_t* base = ix.vecs + r0 * ix.row_bytes;
        const uint64_t last_off = (n - 1 - r0) * (uint64_t)ix.row_bytes;  // rows past the end alias the last row
        float acc[kRT8 * kQT];
#pragma unroll
        for (int j = 0; j < kRT8 * kQT; ++j) acc[j] = 0.0f;
        // Row and query pointers advance once per block of kStepU steps, so every load inside the block is
        // [pointer + immediate] (the index arithmetic was ~45 % of the loop's instructions otherwise).
        const TB* rp[kRT8];
#pragma unroll
        for (int r = 0; r < kRT8; ++r)
            rp[r] = reinterpret_cast<const TB*>(base + min((uint64_t)r * ix.row_bytes, last_off)) + lane;
        const float* qp[kQT];
#pragma unroll
        for (int t = 0; t < kQT; ++t) qp[t] = qs + (size_t)t * dim + lane;
        auto steps = [&](auto un) {
            constexpr int UN = decltype(un)::value;
            float x[UN][kRT8], q[UN][kQT];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
#pragma unroll
                for (int r = 0; r < kRT8; ++r) x[u][r] = load_elem(rp[r], 32 * u);
#pragma unroll
                for (int t = 0; t < kQT; ++t) q[u][t] = qp[t][32 * u];
            }
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int r = 0; r < kRT8; ++r)
#pragma unroll
                    for (int t = 0; t < kQT; ++t) {
                        if (l2) {
                            const float d = __fsub_rn(q[u][t], x[u][r]);
                            acc[r * kQT + t] = __fmaf_rn(d, d, acc[r * kQT + t]);
                        } else {
                            acc[r * kQT + t] = __fmaf_rn(q[u][t], x[u][r], acc[r * kQT + t]);
                        }
                    }
#pragma unroll
            for (int r = 0; r < kRT8; ++r) rp[r] += 32 * UN;
#pragma unroll
            for (int t = 0; t < kQT; ++t) qp[t] += 32 * UN;
        };
        uint32_t i = 0;
        for (; i + 32 * kStepU <= dim; i += 32 * kStepU) steps(std::integral_constant<int, kStepU>{});
        for (; i < dim; i += 32) steps(std::integral_constant<int, 1>{});
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const float s = butterfly32(acc + half * 32, lane);
            const uint64_t row = r0 + half * 4 + my_r;
            if (row < n && my_t < nqt) {
                float v;
                if (l2) {
                    v = __fsqrt_rn(s);
                } else if (ix.metric == VELES_COSINE) {
                    const float nb = *reinterpret_cast<const float*>(ix.vecs + row * ix.row_bytes + ix.norm_off);
                    const float sim = cosine_from_parts(s, qnorm[my_t], nb);
                    v = as_value ? sim : __fsub_rn(1.0f, sim);
                } else {
                    v = as_value ? s : -s;
                }
                out.emit(qbase + my_t, row, v);
            }
        }
    }
}

// ---- many queries, dim % 128 == 0: rows staged in shared memory ------------------------------------------------
// Same 8 x 8 register tile and lane <-> element-class mapping as bf_tile8_kernel, but a CTA (8 warps = 4 query
// groups x 2 row groups) shares its operands: its 32 queries stay in shared memory for the CTA's lifetime and the
// 16 rows of a tile stream through a double-buffered 128-element slab filled by per-thread 16-byte async copies
// one slab ahead.  Every operand read in the FMA loop is then a conflict-free LDS with a short, hidden latency
// (bf_tile8_kernel waits on L2 for its rows: 19 TFLOP/s), and a row element is fetched from L2 once per 32
// queries instead of once per 8.  Arithmetic, summation order and results are unchanged.
constexpr int kTsQ = 32, kTsR = 16, kTsK = 128;
// Work mapping.  Static (`work` == nullptr; collections that fit the L2): grid (x, query group), CTA x takes tiles x,
// x + gridDim.x, ...  Dynamic (large collections): a 1-D grid pulls (chunk of `chunk_tiles` row tiles, query group)
// items from a global counter, chunk-major, so at any moment all CTAs work within a window of a few chunks and every
// query group's pass over a chunk finds the rows the first one brought into the L2: the collection leaves DRAM once
// per batch (ncu, 1M x 768 x 1024 queries: 8.3 GB read with the static mapping, whose query groups drift apart,
// against 3.1 GB of rows).  A CTA restages its 32 queries when its next item belongs to another group (96 KB from the
// L2 per 3 MB of rows).
template <typename TB, bool DYN>
__global__ void __launch_bounds__(kWarps * 32, 2) bf_tile_smem_kernel(IndexView ix, const float* __restrict__ queries,
                                                                   uint32_t nq, ScoreSink out, bool as_value,
                                                                   uint32_t* __restrict__ work, uint32_t chunk_tiles,
                                                                   uint32_t ngroups) {
    extern __shared__ __align__(16) uint8_t ts_smem[];
    const uint32_t dim = ix.dim;
    float* qs_all = reinterpret_cast<float*>(ts_smem);               // kTsQ x dim
    float* qnorm_all = qs_all + (size_t)kTsQ * dim;                  // kTsQ
    TB* slab = reinterpret_cast<TB*>(qnorm_all + kTsQ);              // 2 x kTsR x kTsK
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t qg = warp & 3, rg = warp >> 2;
    const float* qs = qs_all + (size_t)qg * kQT * dim;
    const float* qnorm = qnorm_all + qg * kQT;
    uint32_t qbase = 0, nqt = 0;
    // the CTA's 32 queries (and their norms) into shared memory; every thread calls it
    auto stage = [&](uint32_t group) {
        const uint32_t cta_q0 = group * kTsQ;
        const uint32_t cta_nq = min((uint32_t)kTsQ, nq - cta_q0);
        for (uint32_t i = threadIdx.x; i < kTsQ * dim; i += blockDim.x) {
            const uint32_t t = i / dim;
            qs_all[i] = t < cta_nq ? queries[(size_t)(cta_q0 + t) * dim + (i - t * dim)] : 0.0f;
        }
        __syncthreads();
        if (ix.metric == VELES_COSINE) {
            for (uint32_t t = warp; t < kTsQ; t += kWarps) {
                const float* q = qs_all + (size_t)t * dim;
                const float s = warp_tree_reduce<0>(q, q, dim, lane);
                if (lane == 0) qnorm_all[t] = __fsqrt_rn(s);
            }
        }
        __syncthreads();
        qbase = cta_q0 + qg * kQT;
        nqt = qbase < nq ? min((uint32_t)kQT, nq - qbase) : 0u;
    };
    const bool l2 = ix.metric == VELES_EUCLIDEAN;
    const uint64_t n = ix.n;
    const uint64_t tiles = (n + kTsR - 1) / kTsR;
    const uint32_t nslab = dim / kTsK;
    constexpr uint32_t kChunksPerRow = kTsK * sizeof(TB) / 16;       // 16-byte chunks of one row's slab
    constexpr uint32_t kChunks = kTsR * kChunksPerRow;               // per slab: 512 (f32) or 256 (f16)
    const uint32_t a = (lane & 1) | (lane & 2) | (lane & 4) | ((lane & 16) >> 1) | ((lane & 8) << 1);
    const uint32_t my_r = a >> 3, my_t = a & 7;
    auto do_tile = [&](uint64_t tile) {
        const uint64_t r0 = tile * kTsR;
        // a thread copies the same (row, 16-byte column) chunks of every slab: addresses are set up once per tile
        constexpr uint32_t kPerThread = (kChunks + kWarps * 32 - 1) / (kWarps * 32);  // 2 (f32) or 1 (f16)
        const uint8_t* csrc[kPerThread];
        uint32_t cdst[kPerThread];  // byte offset inside one slab buffer, ~0 = no chunk
#pragma unroll
        for (uint32_t i = 0; i < kPerThread; ++i) {
            const uint32_t c = threadIdx.x + i * kWarps * 32;
            const uint32_t row = c / kChunksPerRow, off = c % kChunksPerRow;
            const uint64_t rr = min(r0 + row, n - 1);  // rows past the end alias the last row, never stored
            csrc[i] = ix.vecs + rr * ix.row_bytes + off * 16;
            cdst[i] = c < kChunks ? (uint32_t)((row * kTsK) * sizeof(TB) + off * 16) : ~0u;
        }
        auto load_slab = [&](uint32_t s, uint32_t buf) {
            uint8_t* base = reinterpret_cast<uint8_t*>(slab + (size_t)buf * kTsR * kTsK);
#pragma unroll
            for (uint32_t i = 0; i < kPerThread; ++i)
                if (cdst[i] != ~0u) cp_async16(base + cdst[i], csrc[i] + (size_t)s * kTsK * sizeof(TB));
            cp_async_commit();
        };
        float acc[kRT8 * kQT];
#pragma unroll
        for (int j = 0; j < kRT8 * kQT; ++j) acc[j] = 0.0f;
        load_slab(0, 0);
        for (uint32_t s = 0; s < nslab; ++s) {
            if (s + 1 < nslab) {
                load_slab(s + 1, (s + 1) & 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();  // every thread's chunks of slab s have landed
            const TB* xs = slab + ((size_t)(s & 1) * kTsR + rg * kRT8) * kTsK + lane;
            const float* qp = qs + (size_t)s * kTsK + lane;
#pragma unroll
            for (int st = 0; st < kTsK / 32; ++st) {
                float x[kRT8], q[kQT];
#pragma unroll
                for (int r = 0; r < kRT8; ++r) x[r] = load_elem(xs + r * kTsK, 32 * st);
#pragma unroll
                for (int t = 0; t < kQT; ++t) q[t] = qp[(size_t)t * dim + 32 * st];
#pragma unroll
                for (int r = 0; r < kRT8; ++r)
#pragma unroll
                    for (int t = 0; t < kQT; ++t) {
                        if (l2) {
                            const float d = __fsub_rn(q[t], x[r]);
                            acc[r * kQT + t] = __fmaf_rn(d, d, acc[r * kQT + t]);
                        } else {
                            acc[r * kQT + t] = __fmaf_rn(q[t], x[r], acc[r * kQT + t]);
                        }
                    }
            }
            __syncthreads();  // buffer s & 1 is refilled by the load issued at the top of iteration s + 1
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const float sum = butterfly32(acc + half * 32, lane);
            const uint64_t row = r0 + rg * kRT8 + half * 4 + my_r;
            if (row < n && my_t < nqt) {
                float v;
                if (l2) {
                    v = __fsqrt_rn(sum);
                } else if (ix.metric == VELES_COSINE) {
                    const float nb = *reinterpret_cast<const float*>(ix.vecs + row * ix.row_bytes + ix.norm_off);
                    const float sim = cosine_from_parts(sum, qnorm[my_t], nb);
                    v = as_value ? sim : __fsub_rn(1.0f, sim);
                } else {
                    v = as_value ? sum : -sum;
                }
                out.emit(qbase + my_t, row, v);
            }
        }
    };
    if (!DYN) {
        stage(blockIdx.y);
        for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) do_tile(tile);
    } else {
        __shared__ uint32_t s_item;
        const uint64_t nchunks = (tiles + chunk_tiles - 1) / chunk_tiles;
        const uint64_t items = nchunks * ngroups;
        uint32_t staged = 0xffffffffu;
        for (;;) {
            __syncthreads();  // the previous item's epilogue (norms in shared memory) and its read of s_item are done
            if (threadIdx.x == 0) s_item = atomicAdd(work, 1u);
            __syncthreads();
            const uint32_t w = s_item;
            if (w >= items) break;
            const uint32_t chunk = w / ngroups, group = w - chunk * ngroups;
            if (group != staged) {
                stage(group);
                staged = group;
            }
            const uint64_t t0 = (uint64_t)chunk * chunk_tiles, t1 = min(tiles, t0 + chunk_tiles);
            for (uint64_t tile = t0; tile < t1; ++tile) do_tile(tile);
        }
    }
}

// ---- few queries (<= 8): the scan is HBM-bound, so it is laid out like a copy -----------------------------
// "Quad" mapping of common.cuh: 8 lanes per row, lane t owns the reference accumulators 4t..4t+3 and reads
// its row with one 128-bit load per 32 elements, so a warp-wide load covers 4 rows x 128 contiguous bytes.
// A lane can work on RB rows at once (rows g, g + 4, ...) so that each 128-bit shared-memory read of a query
// chunk feeds 4 * RB FMAs; measured on B200 RB = 1 is as fast or faster at every QT (more resident warps), so
// that is what is launched.  QT queries (1, 2, 4 or 8) are scored per pass; the rows are read once per
// launch whatever QT is: 5.8 TB/s at QT = 1, 5.4 at 2, 4.7 at 4 (1M x 768 f32), FMA/shared-memory bound at 8.
// dim % 32 == 0.
template <bool L2>
__device__ __forceinline__ void scan_fma(const float4& x, const float4& y, float (&a)[4]) {
    if (L2) {
        const float d0 = __fsub_rn(y.x, x.x), d1 = __fsub_rn(y.y, x.y), d2 = __fsub_rn(y.z, x.z), d3 = __fsub_rn(y.w, x.w);
        a[0] = __fmaf_rn(d0, d0, a[0]);
        a[1] = __fmaf_rn(d1, d1, a[1]);
        a[2] = __fmaf_rn(d2, d2, a[2]);
        a[3] = __fmaf_rn(d3, d3, a[3]);
    } else {
        a[0] = __fmaf_rn(y.x, x.x, a[0]);
        a[1] = __fmaf_rn(y.y, x.y, a[1]);
        a[2] = __fmaf_rn(y.z, x.z, a[2]);
        a[3] = __fmaf_rn(y.w, x.w, a[3]);
    }
}

// U row loads per lane in flight, while at least U steps of 32 elements remain
template <typename TB, int QT, int RB, int U, bool L2>
__device__ __forceinline__ void scan_rows_chunks(uint32_t dim, const float* qs, const TB* const (&r)[RB], uint32_t& i,
                                                 float (&acc)[RB][QT][4]) {
    for (; i + 32 * (U - 1) < dim; i += 32 * U) {
        float4 x[U][RB];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int b = 0; b < RB; ++b) x[u][b] = load4(r[b] + i + 32 * u);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                const float4 y = *reinterpret_cast<const float4*>(qs + (size_t)q * dim + i + 32 * u);
#pragma unroll
                for (int b = 0; b < RB; ++b) scan_fma<L2>(x[u][b], y, acc[b][q]);
            }
    }
}

template <typename TB, int QT, int RB, int U, bool L2>
__device__ __forceinline__ void scan_rows(const IndexView& ix, const float* qs, const TB* const (&r)[RB], uint32_t t,
                                          float (&acc)[RB][QT][4]) {
    const uint32_t dim = ix.dim;
    uint32_t i = t * 4;
    scan_rows_chunks<TB, QT, RB, U, L2>(dim, qs, r, i, acc);
    scan_rows_chunks<TB, QT, RB, 1, L2>(dim, qs, r, i, acc);
}

// Fused selection for the scan kernel (fuse.k != 0): every warp keeps, per query, a sorted list of its k best keys
// in shared memory (a ballot per tile decides whether anything can enter -- after the first tiles almost nothing
// does), the CTA merges its warps' lists, writes one list per (query, CTA) and the last CTA to finish merges those
// and writes the result: the scan and the top-k are ONE launch and no score matrix exists.
constexpr uint32_t kFuseFastK = 32;  // up to this k the fused scan selects by counting (bf_scan_kernel's tail)
struct ScanFuse {
    uint32_t k = 0;             // 0 = not fused (scores go to the sink)
    uint32_t desc = 0;
    uint32_t fast = 0;            // 1: k <= kFuseFastK: register-resident lists + rank merges in the tail
    uint64_t* partial = nullptr;  // nq x gridDim.x x k keys
    uint32_t* done = nullptr;     // CTA counter (self-resetting)
    uint32_t* out_ids = nullptr;
    float* out_score = nullptr;
    unsigned long long* dbg = nullptr;  // VELES_BF_DEBUG_TIMING=1: 8 globaltimer stamps per CTA (thread 0)
};

__device__ __forceinline__ void fuse_stamp(const ScanFuse& f, int slot) {
    if (f.dbg != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        f.dbg[(size_t)blockIdx.x * 8 + slot] = t;
    }
}

// all lanes of the warp call it; inserts `key` into the ascending list res[0..len) capped at k
__device__ __forceinline__ void list_offer(uint64_t* res, uint32_t& len, uint64_t& worst, uint32_t k, uint64_t key, uint32_t lane) {
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

template <typename TB, int QT, int RB>
__global__ void __launch_bounds__(kWarps * 32) bf_scan_kernel(IndexView ix, const float* __restrict__ queries, uint32_t nq,
                                                              ScoreSink out, bool as_value, ScanFuse fuse) {
    extern __shared__ __align__(16) float qs[];  // QT x dim, then QT norms, then (fused) kWarps x QT lists of k keys (+ staging for tail (2a))
    fuse_stamp(fuse, 0);
    constexpr int U = RB >= 4 ? 2 : (RB == 2 ? 4 : 8);  // 8 row loads in flight per lane
    const uint32_t dim = ix.dim;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool l2 = ix.metric == VELES_EUCLIDEAN;
    const uint64_t n = ix.n;
    const uint32_t g = lane >> 3, t = lane & 7;
    const uint64_t tiles = (n + 4 * RB - 1) / (4 * RB);
    // fused selection state: list q of this warp at lists + (warp * QT + q) * k
    uint64_t* lists = reinterpret_cast<uint64_t*>(qs + (((size_t)QT * dim + QT + 3) & ~(size_t)3));
    float* qnorm = qs + (size_t)QT * dim;
    for (uint32_t i = threadIdx.x; i < QT * dim; i += blockDim.x) {
        const uint32_t tq = i / dim;
        qs[i] = tq < nq ? queries[(size_t)tq * dim + (i - tq * dim)] : 0.0f;
    }
    __syncthreads();
    if (ix.metric == VELES_COSINE) {
        for (uint32_t tq = warp; tq < QT; tq += kWarps) {
            const float* q = qs + (size_t)tq * dim;
            const float s = warp_tree_reduce<0>(q, q, dim, lane);
            if (lane == 0) qnorm[tq] = __fsqrt_rn(s);
        }
    }
    __syncthreads();
    fuse_stamp(fuse, 1);
    uint32_t flen[QT];
    uint64_t fworst[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) {
        flen[q] = 0;
        fworst[q] = ~0ull;
    }
    for (uint64_t tile = blockIdx.x * (uint64_t)kWarps + warp; tile < tiles; tile += (uint64_t)gridDim.x * kWarps) {
        const TB* r[RB];
        uint64_t row[RB];
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            row[b] = tile * (4 * RB) + 4 * b + g;
            const uint64_t rr = row[b] < n ? row[b] : n - 1;  // rows past the end alias the last row, not stored
            r[b] = reinterpret_cast<const TB*>(ix.vecs + rr * ix.row_bytes);
        }
        float acc[RB][QT][4];
#pragma unroll
        for (int b = 0; b < RB; ++b)
#pragma unroll
            for (int q = 0; q < QT; ++q) acc[b][q][0] = acc[b][q][1] = acc[b][q][2] = acc[b][q][3] = 0.0f;
        float nbs[RB];
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            nbs[b] = 0.0f;
            if (ix.metric == VELES_COSINE) nbs[b] = *reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(r[b]) + ix.norm_off);
        }
        if (l2)
            scan_rows<TB, QT, RB, U, true>(ix, qs, r, t, acc);
        else
            scan_rows<TB, QT, RB, U, false>(ix, qs, r, t, acc);
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const float nb = nbs[b];
            float mine = 0.0f;  // lane t of the group keeps query t's value
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                const float s = quad_tree_sum(acc[b][q][0], acc[b][q][1], acc[b][q][2], acc[b][q][3]);
                float v;
                if (l2) {
                    v = __fsqrt_rn(s);
                } else if (ix.metric == VELES_COSINE) {
                    const float sim = cosine_from_parts(s, qnorm[q], nb);
                    v = as_value ? sim : __fsub_rn(1.0f, sim);
                } else {
                    v = as_value ? s : -s;
                }
                mine = (int)t == q ? v : mine;
            }
            if (fuse.k == 0) {
                if (t < nq && t < QT && row[b] < n) out.emit(t, row[b], mine);
            } else {
                uint64_t key = ~0ull;
                if (t < nq && t < QT && row[b] < n) {
                    uint32_t o = ord_key(mine);
                    if (fuse.desc) o = ~o;
                    key = ((uint64_t)o << 32) | (uint32_t)row[b];
                }
#pragma unroll
                for (int q = 0; q < QT; ++q) {
                    // lane t of each group holds query t's key for the group's row
                    const uint64_t kq = (int)t == q ? key : ~0ull;
                    if (__any_sync(FULL_MASK, kq < fworst[q]))
                        list_offer(lists + ((size_t)warp * QT + q) * fuse.k, flen[q], fworst[q], fuse.k, kq, lane);
                }
            }
        }
    }
    if (fuse.k == 0) return;
    fuse_stamp(fuse, 2);
    __shared__ uint32_t s_len[kWarps][QT];
    __shared__ uint32_t s_last;
    if (fuse.fast) {
        // ---- k <= 32: selection by counting, no serial list insertion anywhere on the critical path ----
        // (1) CTA merge: the <= 8 * k keys of a query rank themselves (a key's rank = how many keys are smaller; keys are
        //     unique), the first k land in this CTA's slot of `partial`
        if (lane == 0)
            for (int q = 0; q < QT; ++q) s_len[warp][q] = flen[q];
        __syncthreads();
        for (uint32_t q = 0; q < (uint32_t)QT && q < nq; ++q) {
            const uint32_t T = kWarps * fuse.k;
            uint64_t* po = fuse.partial + ((size_t)q * gridDim.x + blockIdx.x) * fuse.k;
            for (uint32_t i = threadIdx.x; i < fuse.k; i += blockDim.x) po[i] = ~0ull;
            __syncthreads();
            for (uint32_t j = threadIdx.x; j < T; j += blockDim.x) {
                const uint32_t w = j / fuse.k, i = j - w * fuse.k;
                if (i >= s_len[w][q]) continue;
                const uint64_t key = lists[((size_t)w * QT + q) * fuse.k + i];
                uint32_t rank = 0;
                for (uint32_t w2 = 0; w2 < kWarps; ++w2) {
                    const uint64_t* l2 = lists + ((size_t)w2 * QT + q) * fuse.k;
                    const uint32_t n2 = s_len[w2][q];
                    for (uint32_t i2 = 0; i2 < n2; ++i2) rank += l2[i2] < key ? 1u : 0u;
                }
                if (rank < fuse.k) po[rank] = key;
            }
        }
        fuse_stamp(fuse, 3);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(fuse.done, 1u) == gridDim.x - 1 ? 1u : 0u;
        __syncthreads();
        fuse_stamp(fuse, 4);
        if (!s_last) return;
        __threadfence();
        if (fuse.fast == 2) {
            // (2a) small grids (the collection is small, so this tail IS the kernel): per query, all gridDim.x * k keys are
            //      staged in shared memory in one round trip.  The k-th smallest SCORE among the CTA lists' heads bounds
            //      the answer from above (k distinct keys lie at or below it): every head counts the heads with a smaller
            //      score (32-bit compares, four per shared-memory load), the largest score among the heads with fewer
            //      than k smaller ones is that bound.  Only keys at or below it -- at most k * k unless scores tie
            //      heavily -- are candidates; they rank themselves among each other.  No serial insertion anywhere; if the
            //      candidates overflow (ties), the general tail (2) below redoes the batch.
            uint64_t* stage = lists + (size_t)kWarps * QT * fuse.k;  // gridDim.x * k keys (host sized all of this)
            const uint32_t G = gridDim.x, T = G * fuse.k, G4 = (G + 3) & ~3u;
            uint64_t* cand = stage + T;                                                      // k * k candidates
            uint32_t* hhi = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(cand + (size_t)fuse.k * fuse.k) + 15) & ~(uintptr_t)15);  // gridDim.x head scores (+ pad)
            __shared__ uint32_t s_bound_hi, s_ncand, s_overflow;
            if (threadIdx.x == 0) s_overflow = 0;
            for (uint32_t q = 0; q < nq; ++q) {
                const uint64_t* in = fuse.partial + (size_t)q * T;
                // every key of the query in ONE round trip: up to 16 loads per thread in flight (grids of <= 2 CTAs per SM)
                for (uint32_t j0 = threadIdx.x; j0 < T; j0 += blockDim.x * 16) {
                    uint64_t key[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const uint32_t j = j0 + u * blockDim.x;
                        key[u] = j < T ? __ldcg(in + j) : ~0ull;
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const uint32_t j = j0 + u * blockDim.x;
                        if (j < T) stage[j] = key[u];
                    }
                }
                if (threadIdx.x == 0) {
                    s_bound_hi = G < fuse.k ? 0xffffffffu : 0u;  // fewer than k lists: no bound, all G * k < k * k keys are candidates
                    s_ncand = 0;
                }
                __syncthreads();
                fuse_stamp(fuse, 5);
                for (uint32_t c = threadIdx.x; c < G4; c += blockDim.x) hhi[c] = c < G ? (uint32_t)(stage[(size_t)c * fuse.k] >> 32) : 0xffffffffu;
                __syncthreads();
                for (uint32_t c = threadIdx.x; c < G; c += blockDim.x) {
                    const uint32_t mine = hhi[c];
                    uint32_t smaller = 0;
                    const uint4* h4 = reinterpret_cast<const uint4*>(hhi);
#pragma unroll 4
                    for (uint32_t c4 = 0; c4 < G4 / 4; ++c4) {
                        const uint4 h = h4[c4];
                        smaller += (h.x < mine ? 1u : 0u) + (h.y < mine ? 1u : 0u) + (h.z < mine ? 1u : 0u) + (h.w < mine ? 1u : 0u);
                    }
                    if (smaller < fuse.k) atomicMax(&s_bound_hi, mine);  // empty lists (score bits ~0) qualify when fewer than k have a row
                }
                __syncthreads();
                fuse_stamp(fuse, 7);
                const uint64_t bound = ((uint64_t)s_bound_hi << 32) | 0xffffffffull;
                for (uint32_t j = threadIdx.x; j < T; j += blockDim.x) {
                    const uint64_t key = stage[j];
                    if (key != ~0ull && key <= bound) {
                        const uint32_t slot = atomicAdd(&s_ncand, 1u);
                        if (slot < fuse.k * fuse.k) cand[slot] = key;
                    }
                }
                __syncthreads();
                if (s_ncand > fuse.k * fuse.k) {  // uniform: heavily tied scores
                    if (threadIdx.x == 0) s_overflow = 1;
                    __syncthreads();
                    continue;
                }
                const uint32_t nc = s_ncand;
                for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) {
                    const uint64_t key = cand[i];
                    uint32_t rank = 0;
                    for (uint32_t i2 = 0; i2 < nc; ++i2) rank += cand[i2] < key ? 1u : 0u;
                    if (rank < fuse.k) {
                        const uint32_t o = (uint32_t)(key >> 32);
                        fuse.out_ids[(size_t)q * fuse.k + rank] = (uint32_t)key;
                        fuse.out_score[(size_t)q * fuse.k + rank] = ord_unkey(fuse.desc ? ~o : o);
                    }
                }
                for (uint32_t i = nc + threadIdx.x; i < fuse.k; i += blockDim.x) {  // fewer rows than k
                    fuse.out_ids[(size_t)q * fuse.k + i] = VELES_INVALID_ID;
                    fuse.out_score[(size_t)q * fuse.k + i] = __uint_as_float(0x7fc00000u);
                }
                __syncthreads();
            }
            if (!s_overflow) {
                fuse_stamp(fuse, 6);
                if (threadIdx.x == 0) *fuse.done = 0u;  // ready for the next launch
                return;
            }
            __syncthreads();
        }
        // (2) the last CTA: its warps are shared out among the queries (kWarps / QT each).  A warp reads its share of a
        //     query's gridDim.x * k keys, eight loads in flight, into a register-resident top-k; the lists of a query's
        //     warps are then merged by rank (sorted lists of distinct keys: a key's place is the sum of its lower bounds),
        //     all threads at once.  (The first version selected by an eight-pass radix descent over keys staged in shared
        //     memory: ~8 us of barriers per query, and the queries one after another.)
        constexpr int WPQ = kWarps / QT;
        const uint32_t T = gridDim.x * fuse.k;
        {
            const uint32_t myq = warp / WPQ, sub = warp % WPQ;
            RegTopK<1> top;
            top.init(fuse.k, lane);
            if (myq < nq) {
                const uint64_t* in = fuse.partial + (size_t)myq * T;
                for (uint32_t j0 = sub * 32; j0 < T; j0 += WPQ * 32 * 8) {
                    uint64_t key[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t j = j0 + u * WPQ * 32 + lane;
                        key[u] = j < T ? __ldcg(in + j) : ~0ull;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) top.offer(key[u]);
                }
            }
            top.store(lists + (size_t)warp * fuse.k, fuse.k);  // the warps' lists of part (1) are no longer needed
        }
        __syncthreads();
        fuse_stamp(fuse, 5);
        for (uint32_t idx = threadIdx.x; idx < kWarps * fuse.k; idx += blockDim.x) {
            const uint32_t w = idx / fuse.k, q = w / WPQ;
            if (q >= nq) continue;
            const uint64_t key = lists[idx];
            if (key == ~0ull) continue;
            uint32_t rank = 0;
            for (uint32_t w2 = q * WPQ; w2 < (q + 1) * WPQ; ++w2) {
                const uint64_t* l = lists + (size_t)w2 * fuse.k;
                uint32_t lo = 0, hi = fuse.k;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (l[mid] < key) lo = mid + 1; else hi = mid;
                }
                rank += lo;
            }
            if (rank < fuse.k) {
                const uint32_t o = (uint32_t)(key >> 32);
                fuse.out_ids[(size_t)q * fuse.k + rank] = (uint32_t)key;
                fuse.out_score[(size_t)q * fuse.k + rank] = ord_unkey(fuse.desc ? ~o : o);
            }
        }
        // positions past the number of rows found (collections smaller than k)
        for (uint32_t idx = threadIdx.x; idx < nq * fuse.k; idx += blockDim.x) {
            const uint32_t q = idx / fuse.k, pos = idx - q * fuse.k;
            uint32_t have = 0;
            for (uint32_t w2 = q * WPQ; w2 < (q + 1) * WPQ && have <= pos; ++w2) {
                const uint64_t* l = lists + (size_t)w2 * fuse.k;
                uint32_t lo = 0, hi = fuse.k;  // the list is sorted and padded with ~0
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (l[mid] != ~0ull) lo = mid + 1; else hi = mid;
                }
                have += lo;
            }
            if (pos >= have) {
                fuse.out_ids[idx] = VELES_INVALID_ID;
                fuse.out_score[idx] = __uint_as_float(0x7fc00000u);
            }
        }
        __syncthreads();
        fuse_stamp(fuse, 6);
        if (threadIdx.x == 0) *fuse.done = 0u;  // ready for the next launch
        return;
    }
    // ---- k > 32: CTA merge: warp q % kWarps merges query q's kWarps lists into warp 0's list of that query ----
    if (lane == 0)
        for (int q = 0; q < QT; ++q) s_len[warp][q] = flen[q];
    __syncthreads();
    for (uint32_t q = warp; q < (uint32_t)QT && q < nq; q += kWarps) {
        uint64_t* dst = lists + (size_t)q * fuse.k;  // warp 0's list of query q
        uint32_t len = s_len[0][q];
        uint64_t worst = len == fuse.k ? dst[fuse.k - 1] : ~0ull;
        for (uint32_t w = 1; w < kWarps; ++w) {
            const uint64_t* src = lists + ((size_t)w * QT + q) * fuse.k;
            const uint32_t sl = s_len[w][q];
            for (uint32_t j0 = 0; j0 < sl; j0 += 32) {
                const uint32_t j = j0 + lane;
                list_offer(dst, len, worst, fuse.k, j < sl ? src[j] : ~0ull, lane);
            }
        }
        __syncwarp();
        uint64_t* po = fuse.partial + ((size_t)q * gridDim.x + blockIdx.x) * fuse.k;
        for (uint32_t i = lane; i < fuse.k; i += 32) po[i] = i < len ? dst[i] : ~0ull;
    }
    // ---- the last CTA merges the per-CTA lists: every warp takes a slice of them (loads batched eight deep, so the
    // merge costs a few DRAM/L2 round trips, not one per 32 keys), then warp 0 merges the eight slices ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(fuse.done, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (uint32_t q = 0; q < nq; ++q) {
        uint64_t* dst = lists + (size_t)warp * QT * fuse.k;  // this warp's list space, free again by now
        uint32_t len = 0;
        uint64_t worst = ~0ull;
        const uint64_t* in = fuse.partial + (size_t)q * gridDim.x * fuse.k;
        // CTA lists are ascending and padded with ~0: list c of this warp's slice is c = warp, warp + kWarps, ...
        const uint32_t total = gridDim.x * fuse.k;
        for (uint32_t j0 = warp * 32; j0 < total; j0 += kWarps * 32 * 8) {
            uint64_t key[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t j = j0 + u * kWarps * 32 + lane;
                key[u] = j < total ? __ldcg(in + j) : ~0ull;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (__any_sync(FULL_MASK, key[u] < worst)) list_offer(dst, len, worst, fuse.k, key[u], lane);
        }
        if (lane == 0) s_len[warp][0] = len;
        __syncthreads();
        if (warp == 0) {
            for (uint32_t w = 1; w < kWarps; ++w) {
                const uint64_t* src = lists + (size_t)w * QT * fuse.k;
                const uint32_t sl = s_len[w][0];
                for (uint32_t j0 = 0; j0 < sl; j0 += 32) {
                    const uint32_t j = j0 + lane;
                    list_offer(dst, len, worst, fuse.k, j < sl ? src[j] : ~0ull, lane);
                }
            }
            __syncwarp();
            for (uint32_t i = lane; i < fuse.k; i += 32) {
                uint32_t id = VELES_INVALID_ID;
                float sc = __uint_as_float(0x7fc00000u);
                if (i < len) {
                    id = (uint32_t)dst[i];
                    const uint32_t o = (uint32_t)(dst[i] >> 32);
                    sc = ord_unkey(fuse.desc ? ~o : o);
                }
                fuse.out_ids[(size_t)q * fuse.k + i] = id;
                fuse.out_score[(size_t)q * fuse.k + i] = sc;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *fuse.done = 0u;  // ready for the next launch
}

// generic path: any metric, any dim >= 1, F32/F16 rows: one warp per (query, row) pair
template <typename TB>
__global__ void bf_generic_kernel(IndexView ix, const float* __restrict__ queries, uint32_t nq,
                                  ScoreSink out, bool as_value) {
    extern __shared__ __align__(16) float qs[];  // one query
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t q = blockIdx.y;
    const uint32_t dim = ix.dim;
    for (uint32_t i = threadIdx.x; i < dim; i += blockDim.x) qs[i] = queries[(size_t)q * dim + i];
    __syncthreads();
    float na = 0.0f;
    if (ix.metric == VELES_COSINE) na = __fsqrt_rn(warp_tree_reduce<0>(qs, qs, dim, lane));
    for (uint64_t r = blockIdx.x * (uint64_t)nw + warp; r < ix.n; r += (uint64_t)gridDim.x * nw) {
        const uint8_t* row = ix.vecs + r * ix.row_bytes;
        float nb = ix.metric == VELES_COSINE ? *reinterpret_cast<const float*>(row + ix.norm_off) : 0.0f;
        float v = warp_metric(ix.metric, as_value, qs, reinterpret_cast<const TB*>(row), dim, na, nb, lane);
        if (lane == 0) out.emit(q, r, v);
    }
}

// packed-bit rows: Hamming count of (query > 0.5) bits vs row bits; one thread per row chunk
__global__ void bf_bin_kernel(IndexView ix, const float* __restrict__ queries, uint32_t nq, ScoreSink out) {
    extern __shared__ __align__(16) uint32_t qw[];  // dim/32 words
    const uint32_t q = blockIdx.y, words = ix.dim >> 5;
    for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) {
        uint32_t bits = 0;
        for (uint32_t b = 0; b < 32; ++b) bits |= (queries[(size_t)q * ix.dim + w * 32 + b] > 0.5f ? 1u : 0u) << b;
        qw[w] = bits;
    }
    __syncthreads();
    // 8 lanes per row, 16 bytes per lane per step: a warp reads 4 rows x 128 contiguous bytes
    const uint32_t lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
    const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r0 = gw * 4; r0 < ix.n; r0 += nwarps * 4) {
        const uint64_t r = r0 + grp;
        uint32_t d = 0;
        if (r < ix.n) {
            const uint4* row = reinterpret_cast<const uint4*>(ix.vecs + r * ix.row_bytes);
            for (uint32_t w4 = sub; w4 * 4 < words; w4 += 8) {
                uint4 x = row[w4];
                const uint32_t* qq = qw + w4 * 4;
                d += __popc(x.x ^ qq[0]);
                if (w4 * 4 + 1 < words) d += __popc(x.y ^ qq[1]);
                if (w4 * 4 + 2 < words) d += __popc(x.z ^ qq[2]);
                if (w4 * 4 + 3 < words) d += __popc(x.w ^ qq[3]);
            }
        }
        d += __shfl_xor_sync(FULL_MASK, d, 4);
        d += __shfl_xor_sync(FULL_MASK, d, 2);
        d += __shfl_xor_sync(FULL_MASK, d, 1);
        if (sub == 0 && r < ix.n) out.emit(q, r, (float)d);
    }
}

// per query: k smallest keys of (order(score) << 32 | row); one warp per query
__global__ void __launch_bounds__(32) topk_kernel(const float* __restrict__ scores, uint64_t n, uint32_t k,
                                                  bool descending, uint32_t* __restrict__ out_ids,
                                                  float* __restrict__ out_score) {
    extern __shared__ __align__(16) uint64_t res[];
    const uint32_t lane = threadIdx.x, q = blockIdx.x;
    const float* row = scores + (size_t)q * n;
    uint32_t len = 0;
    uint64_t worst = ~0ull;
    for (uint64_t base = 0; base < n; base += 32) {
        const uint64_t i = base + lane;
        uint64_t key = ~0ull;
        bool cand = false;
        if (i < n) {
            uint32_t o = ord_key(row[i]);
            if (descending) o = ~o;
            key = ((uint64_t)o << 32) | (uint32_t)i;
            cand = len < k || key < worst;
        }
        uint32_t msk = __ballot_sync(FULL_MASK, cand);
        while (msk) {
            const uint32_t src = __ffs(msk) - 1;
            msk &= msk - 1;
            const uint64_t kk = __shfl_sync(FULL_MASK, key, src);
            if (len < k) {
                const uint32_t pos = lower_bound_warp(res, len, kk, lane);
                insert_at(res, pos, len + 1, kk, lane);
                ++len;
            } else if (kk < worst) {
                const uint32_t pos = lower_bound_warp(res, len, kk, lane);
                insert_at(res, pos, len, kk, lane);
            }
            if (len == k) worst = res[k - 1];
        }
    }
    __syncwarp();
    for (uint32_t i = lane; i < k; i += 32) {
        uint32_t id = VELES_INVALID_ID;
        float s = __uint_as_float(0x7fc00000u);
        if (i < len) {
            uint64_t key = res[i];
            id = (uint32_t)key;
            s = row[id];
        }
        out_ids[(size_t)q * k + i] = id;
        out_score[(size_t)q * k + i] = s;
    }
}

// k <= 512: one CTA of kTopWarps warps per query, each warp streams a slice of the score row (float4 loads)
// through a threshold filter into its own sorted list; warp 0 merges the lists.
constexpr int kTopWarps = 8;
// With `partial` != NULL the grid is (queries, chunks): CTA (q, c) selects from elements [c * chunk, (c+1) * chunk)
// of the row and writes its sorted keys (padded with ~0) to partial[(q * chunks + c) * k ..]; topk_merge_kernel
// finishes.  That keeps the selection parallel when there are few queries and many rows.
__global__ void __launch_bounds__(kTopWarps * 32) topk_cta_kernel(const float* __restrict__ scores, uint64_t n, uint32_t k,
                                                                 bool descending, uint32_t* __restrict__ out_ids,
                                                                 float* __restrict__ out_score, uint64_t chunk,
                                                                 uint64_t* __restrict__ partial) {
    extern __shared__ __align__(16) uint64_t tk_smem[];
    __shared__ uint32_t s_len[kTopWarps];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q = blockIdx.x;
    const float* row = scores + (size_t)q * n;
    uint64_t* res = tk_smem + (size_t)warp * k;
    const uint64_t r_begin = partial ? (uint64_t)blockIdx.y * chunk : 0;
    const uint64_t r_end = partial ? min(n, r_begin + chunk) : n;
    // slices start at multiples of 128 elements past the row's 16-byte alignment point
    const uint64_t mis = (((uintptr_t)row) >> 2) & 3;       // elements until the next 16-byte boundary: (4 - mis) % 4
    const uint64_t head = (4 - mis) & 3;
    uint64_t seg = (r_end - r_begin + kTopWarps - 1) / kTopWarps;
    seg = (seg + 127) & ~(uint64_t)127;
    const uint64_t c_begin = min(r_end, r_begin + warp * seg), c_end = min(r_end, c_begin + seg);
    uint32_t len = 0;
    uint64_t worst = ~0ull;
    auto offer = [&](uint64_t key) {
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
    };
    auto make_key = [&](float v, uint64_t i) {
        uint32_t o = ord_key(v);
        if (descending) o = ~o;
        return ((uint64_t)o << 32) | (uint32_t)i;
    };
    (void)head;
    for (uint64_t base = c_begin; base < c_end; base += 128) {
        const uint64_t i0 = base + (uint64_t)lane * 4;
        uint64_t key[4] = {~0ull, ~0ull, ~0ull, ~0ull};
        if (mis == 0 && i0 + 3 < c_end) {
            const float4 v = *reinterpret_cast<const float4*>(row + i0);
            key[0] = make_key(v.x, i0);
            key[1] = make_key(v.y, i0 + 1);
            key[2] = make_key(v.z, i0 + 2);
            key[3] = make_key(v.w, i0 + 3);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (i0 + e < c_end) key[e] = make_key(row[i0 + e], i0 + e);
        }
        const bool any = key[0] < worst || key[1] < worst || key[2] < worst || key[3] < worst;
        if (!__any_sync(FULL_MASK, any)) continue;
#pragma unroll
        for (int e = 0; e < 4; ++e) offer(key[e]);
    }
    if (lane == 0) s_len[warp] = len;
    __syncthreads();
    if (warp != 0) return;
    for (uint32_t w = 1; w < kTopWarps; ++w) {
        const uint64_t* other = tk_smem + (size_t)w * k;
        const uint32_t olen = s_len[w];
        for (uint32_t j0 = 0; j0 < olen; j0 += 32) {
            const uint32_t j = j0 + lane;
            offer(j < olen ? other[j] : ~0ull);
        }
    }
    __syncwarp();
    if (partial) {
        uint64_t* out = partial + ((size_t)q * gridDim.y + blockIdx.y) * k;
        for (uint32_t i = lane; i < k; i += 32) out[i] = i < len ? res[i] : ~0ull;
        return;
    }
    for (uint32_t i = lane; i < k; i += 32) {
        uint32_t id = VELES_INVALID_ID;
        float sc = __uint_as_float(0x7fc00000u);
        if (i < len) {
            id = (uint32_t)res[i];
            sc = row[id];
        }
        out_ids[(size_t)q * k + i] = id;
        out_score[(size_t)q * k + i] = sc;
    }
}

// second level: one warp per query merges `chunks` sorted key lists of length k (k <= 512)
__global__ void __launch_bounds__(32) topk_merge_kernel(const uint64_t* __restrict__ partial, uint32_t chunks, uint32_t k,
                                                        const float* __restrict__ scores, uint64_t n,
                                                        uint32_t* __restrict__ out_ids, float* __restrict__ out_score) {
    extern __shared__ __align__(16) uint64_t res[];
    const uint32_t lane = threadIdx.x, q = blockIdx.x;
    const uint64_t* in = partial + (size_t)q * chunks * k;
    uint32_t len = 0;
    uint64_t worst = ~0ull;
    const uint32_t total = chunks * k;
    for (uint32_t base = 0; base < total; base += 32) {
        const uint32_t i = base + lane;
        const uint64_t key = i < total ? in[i] : ~0ull;
        uint32_t msk = __ballot_sync(FULL_MASK, key != ~0ull && (len < k || key < worst));
        while (msk) {
            const uint32_t src = __ffs(msk) - 1;
            msk &= msk - 1;
            const uint64_t kk = __shfl_sync(FULL_MASK, key, src);
            if (len < k) {
                const uint32_t pos = lower_bound_warp(res, len, kk, lane);
                insert_at(res, pos, len + 1, kk, lane);
                ++len;
            } else if (kk < worst) {
                const uint32_t pos = lower_bound_warp(res, len, kk, lane);
                insert_at(res, pos, len, kk, lane);
            }
            if (len == k) worst = res[k - 1];
        }
    }
    __syncwarp();
    for (uint32_t i = lane; i < k; i += 32) {
        uint32_t id = VELES_INVALID_ID;
        float sc = __uint_as_float(0x7fc00000u);
        if (i < len) {
            id = (uint32_t)res[i];
            sc = scores[(size_t)q * n + id];
        }
        out_ids[(size_t)q * k + i] = id;
        out_score[(size_t)q * k + i] = sc;
    }
}

// metric value of explicit (query, candidate) pairs; one warp per pair
template <typename TB>
__global__ void rerank_kernel(IndexView ix, const float* __restrict__ queries, uint32_t nq, const uint32_t* __restrict__ cand,
                              uint32_t m, float* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t pi = gw; pi < (uint64_t)nq * m; pi += nwarps) {
        const uint32_t q = (uint32_t)(pi / m);
        const uint32_t id = cand[pi];
        float v = __uint_as_float(0x7fc00000u);
        if (id != VELES_INVALID_ID && id < ix.n) {
            const float* qv = queries + (size_t)q * ix.dim;
            const uint8_t* row = ix.vecs + (size_t)id * ix.row_bytes;
            if (ix.dtype == VELES_BIN1) {
                uint32_t d = 0;
                const uint32_t* rw = reinterpret_cast<const uint32_t*>(row);
                for (uint32_t w = lane; w < (ix.dim >> 5); w += 32) {
                    uint32_t bits = 0;
                    for (uint32_t b = 0; b < 32; ++b) bits |= (qv[w * 32 + b] > 0.5f ? 1u : 0u) << b;
                    d += __popc(bits ^ rw[w]);
                }
                v = (float)__reduce_add_sync(FULL_MASK, d);
            } else {
                float na = 0.0f, nb = 0.0f;
                if (ix.metric == VELES_COSINE) {
                    na = __fsqrt_rn(warp_tree_reduce<0>(qv, qv, ix.dim, lane));
                    nb = *reinterpret_cast<const float*>(row + ix.norm_off);
                }
                v = warp_metric(ix.metric, true, qv, reinterpret_cast<const TB*>(row), ix.dim, na, nb, lane);
            }
        }
        if (lane == 0) out[pi] = v;
    }
}

// distance of explicit host-provided pairs (a[i], b[i]); one warp per pair
__global__ void pairs_kernel(int metric, const float* __restrict__ a, const float* __restrict__ b, uint32_t n_pairs,
                             uint32_t dim, bool as_value, float* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = gw; i < n_pairs; i += nwarps) {
        const float* x = a + i * dim;
        const float* y = b + i * dim;
        float na = 0.0f, nb = 0.0f;
        if (metric == VELES_COSINE) {
            na = __fsqrt_rn(warp_tree_reduce<0>(x, x, dim, lane));
            nb = __fsqrt_rn(warp_tree_reduce<0>(y, y, dim, lane));
        }
        float v = warp_metric(metric, as_value, x, y, dim, na, nb, lane);
        if (lane == 0) out[i] = v;
    }
}

// ---- host side ---------------------------------------------------------------------------------
// `v`: the snapshot's view, or a strided view of it (row_bytes multiplied: every S-th row) for the sample pass
static int32_t launch_scores(const veles_index* ixh, const IndexView& v, const float* q_d, uint32_t nq, const ScoreSink& sink,
                             bool as_value, cudaStream_t st, const ScanFuse* fuse = nullptr) {
    struct {
        uint64_t n;
        uint32_t dim;
        int32_t dtype, metric;
    } view{v.n, v.dim, v.dtype, v.metric}, *ix = &view;
    (void)ixh;
    const int sms = device_sm_count();
    if (ix->dtype == VELES_BIN1) {
        dim3 grid((unsigned)std::min<uint64_t>((ix->n + 31) / 32 + 1, (uint64_t)sms * 8), nq);
        bf_bin_kernel<<<grid, 256, (ix->dim / 32) * 4, st>>>(v, q_d, nq, sink);
        count_launch();
    } else if (ix->dim >= 16 && (ix->metric == VELES_COSINE || ix->metric == VELES_EUCLIDEAN || ix->metric == VELES_DOT)) {
        const size_t smem = ((size_t)kQT * ix->dim + kQT) * 4;
        const uint32_t qtiles = (nq + kQT - 1) / kQT;
        if (ix->dim % 32 == 0 && nq <= 8) {
            // few queries: HBM-bound scan, rows read once (bf_scan_kernel)
            const uint32_t qt = nq <= 1 ? 1 : nq <= 2 ? 2 : nq <= 4 ? 4 : 8;
            // one row per lane (RB = 1) measured best on B200 at every QT (RB = 2, 4 were no faster: fewer resident warps)
            constexpr uint32_t rb = 1;
            using ScanT = void (*)(IndexView, const float*, uint32_t, ScoreSink, bool, ScanFuse);
            ScanT ks;
#define VELES_SCAN(TB) \
    (qt == 1 ? bf_scan_kernel<TB, 1, 1> : qt == 2 ? bf_scan_kernel<TB, 2, 1> : qt == 4 ? bf_scan_kernel<TB, 4, 1> : bf_scan_kernel<TB, 8, 1>)
            if (ix->dtype == VELES_F32)
                ks = VELES_SCAN(float);
            else
                ks = VELES_SCAN(__half);
#undef VELES_SCAN
            ScanFuse fz = fuse ? *fuse : ScanFuse();
            const uint64_t tiles = (ix->n + 4 * rb - 1) / (4 * rb);
            const size_t smem_base = ((((size_t)qt * ix->dim + qt) + 3) & ~(size_t)3) * 4 + (size_t)kWarps * qt * fz.k * 8;
            int per_sm = cached_blocks_per_sm(reinterpret_cast<const void*>(ks), kWarps * 32, smem_base);
            if (per_sm < 1) return VELES_ERR_CUDA;
            uint64_t gx = std::max<uint64_t>(1, std::min<uint64_t>((tiles + kWarps - 1) / kWarps, (uint64_t)sms * per_sm));
            size_t smem_s = smem_base;
            if (fz.k) {
                // fused selection: every CTA leaves a list for the last CTA to merge, so small collections get at most two
                // CTAs per SM (one tile per warp where that is enough); k <= 32 selects with register lists and rank merges
                gx = std::min<uint64_t>(gx, std::max<uint64_t>((uint64_t)sms * std::min(per_sm, 2), (tiles + kWarps * 4 - 1) / (kWarps * 4)));
                fz.fast = fz.k <= kFuseFastK ? 1u : 0u;
                // small grids (the collection is small, so the tail IS the kernel): the last CTA stages every CTA's list in
                // shared memory and selects by ranks (tail (2a)) when that costs no resident CTA
                const size_t stage = ((size_t)gx * fz.k + (size_t)fz.k * fz.k) * 8 + 16 + ((size_t)gx + 4) * 4;
                if (fz.fast && gx <= (uint64_t)2 * sms && stage <= 64 * 1024 && std::getenv("VELES_BF_NO_STAGED_TAIL") == nullptr) {
                    const int per_sm2 = cached_blocks_per_sm(reinterpret_cast<const void*>(ks), kWarps * 32, smem_base + stage);
                    if (per_sm2 >= 1 && per_sm2 * (uint64_t)sms >= gx) {
                        fz.fast = 2;
                        smem_s += stage;
                    }
                }
            }
            static const bool dbg_timing = std::getenv("VELES_BF_DEBUG_TIMING") != nullptr;
            unsigned long long* dbg_d = nullptr;
            if (dbg_timing && fz.k) {  // diagnostics only: phase stamps per CTA, printed after a synchronize
                VELES_CUDA(cudaMalloc(&dbg_d, (size_t)gx * 64));
                VELES_CUDA(cudaMemsetAsync(dbg_d, 0, (size_t)gx * 64, st));
                fz.dbg = dbg_d;
            }
            ks<<<(unsigned)gx, kWarps * 32, smem_s, st>>>(v, q_d, nq, sink, as_value, fz);
            count_launch();
            VELES_CUDA(cudaGetLastError());
            if (dbg_d) {
                std::vector<unsigned long long> h((size_t)gx * 8);
                VELES_CUDA(cudaStreamSynchronize(st));
                VELES_CUDA(cudaMemcpy(h.data(), dbg_d, h.size() * 8, cudaMemcpyDeviceToHost));
                cudaFree(dbg_d);
                unsigned long long t0 = ~0ull;
                for (uint64_t c = 0; c < gx; ++c) t0 = std::min(t0, h[c * 8]);
                static const char* names[8] = {"entry", "query staged", "scan done", "cta merged", "ticket", "keys staged", "done", "bound found"};
                for (int sl = 0; sl < 8; ++sl) {
                    unsigned long long lo = ~0ull, hi = 0;
                    for (uint64_t c = 0; c < gx; ++c)
                        if (h[c * 8 + sl]) {
                            lo = std::min(lo, h[c * 8 + sl] - t0);
                            hi = std::max(hi, h[c * 8 + sl] - t0);
                        }
                    if (hi) std::fprintf(stderr, "[bf timing] %-13s first %6.2f us  last %6.2f us (grid %llu, tail %u)\n", names[sl], lo / 1e3, hi / 1e3, (unsigned long long)gx, fz.fast);
                }
            }
            return VELES_OK;
        }
        if (ix->dim % kTsK == 0 && nq >= 16 && std::getenv("VELES_BF_NO_SMEM_TILE") == nullptr) {
            // many queries: operands staged in shared memory (bf_tile_smem_kernel)
            const size_t tb = ix->dtype == VELES_F32 ? 4 : 2;
            const size_t smem_t = ((size_t)kTsQ * ix->dim + kTsQ) * 4 + 2 * (size_t)kTsR * kTsK * tb;
            int max_optin = 0, dev_id = 0;
            VELES_CUDA(cudaGetDevice(&dev_id));
            VELES_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev_id));
            if (smem_t <= (size_t)max_optin) {
                auto kt = ix->dtype == VELES_F32 ? bf_tile_smem_kernel<float, false> : bf_tile_smem_kernel<__half, false>;
                auto ktd = ix->dtype == VELES_F32 ? bf_tile_smem_kernel<float, true> : bf_tile_smem_kernel<__half, true>;
                VELES_CUDA(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
                VELES_CUDA(cudaFuncSetAttribute(ktd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
                int per_sm = 1;
                VELES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kt, kWarps * 32, smem_t));
                const uint32_t qgroups = (nq + kTsQ - 1) / kTsQ;
                const uint64_t tiles = (ix->n + kTsR - 1) / kTsR;
                const uint64_t slots = (uint64_t)sms * std::max(per_sm, 1);
                // rows that do not fit the L2 and several query groups: dynamic chunk-major items (see the kernel), 64
                // tiles (1024 rows) per item, or fewer when that would leave less than ~8 items per CTA
                const uint64_t touched = v.n * (uint64_t)ix->dim * tb;
                uint64_t dyn_min = (uint64_t)64 << 20;
                if (const char* e = std::getenv("VELES_BF_DYNAMIC_MIN_BYTES")) dyn_min = std::strtoull(e, nullptr, 10);  // probes
                if (touched > dyn_min && qgroups > 1 && qgroups <= 32768 && tiles * qgroups < 0xffffff00ull &&
                    std::getenv("VELES_BF_STATIC_TILES") == nullptr) {
                    uint64_t chunk_tiles = 64;
                    while (chunk_tiles > 8 && (tiles / chunk_tiles) * qgroups < 8 * slots) chunk_tiles >>= 1;
                    if (!ixh->bf_work.p) VELES_TRY(ixh->bf_work.alloc(64));
                    VELES_CUDA(cudaMemsetAsync(ixh->bf_work.p, 0, 4, st));
                    const uint64_t items = (tiles + chunk_tiles - 1) / chunk_tiles * qgroups;
                    ktd<<<(unsigned)std::min<uint64_t>(slots, items), kWarps * 32, smem_t, st>>>(
                        v, q_d, nq, sink, as_value, ixh->bf_work.as<uint32_t>(), (uint32_t)chunk_tiles, qgroups);
                    count_launch();
                    VELES_CUDA(cudaGetLastError());
                    return VELES_OK;
                }
                for (uint32_t g0 = 0; g0 < qgroups; g0 += 32768) {
                    const uint32_t ng = std::min(32768u, qgroups - g0);
                    const uint64_t gx = std::max<uint64_t>(1, std::min<uint64_t>(tiles, std::max<uint64_t>(1, slots / ng)));
                    dim3 grid((unsigned)gx, ng);
                    const uint32_t qoff = g0 * kTsQ;
                    kt<<<grid, kWarps * 32, smem_t, st>>>(v, q_d + (size_t)qoff * ix->dim, nq - qoff, sink.at(qoff), as_value, nullptr, 0u,
                                                          ng);
                    count_launch();
                }
                VELES_CUDA(cudaGetLastError());
                return VELES_OK;
            }
        }
        if (ix->dim % 32 == 0) {
            // one query tile per CTA: sharing a row tile between the warps of a CTA (QW = 2, 4: one L2 read,
            // QW - 1 L1 hits) measured 5-10 % slower on B200 -- the kernel is issue-bound, not L2-bound
            constexpr uint32_t qw = 1;
            auto kern8 = ix->dtype == VELES_F32 ? bf_tile8_kernel<float, 1> : bf_tile8_kernel<__half, 1>;
            const size_t smem8 = ((size_t)qw * kQT * ix->dim + qw * kQT) * 4;
            VELES_CUDA(cudaFuncSetAttribute(kern8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
            const uint32_t ctiles = (nq + kQT * qw - 1) / (kQT * qw);  // CTA-level query tiles
            const uint64_t row_tiles = (ix->n + kRT8 - 1) / kRT8;
            const uint32_t row_warps = kWarps / qw;
            uint64_t gx = (row_tiles + row_warps - 1) / row_warps;
            const uint64_t want = std::max<uint64_t>(1, ((uint64_t)sms * 4 + ctiles - 1) / ctiles);
            gx = std::max<uint64_t>(1, std::min(gx, want));
            for (uint32_t t0 = 0; t0 < ctiles; t0 += 32768) {
                const uint32_t nt = std::min(32768u, ctiles - t0);
                const uint32_t qoff = t0 * kQT * qw;
                dim3 grid((unsigned)gx, nt);
                kern8<<<grid, kWarps * 32, smem8, st>>>(v, q_d + (size_t)qoff * ix->dim, nq - qoff, sink.at(qoff), as_value);
                count_launch();
            }
            VELES_CUDA(cudaGetLastError());
            return VELES_OK;
        }
        auto kern = ix->dtype == VELES_F32 ? bf_tile_kernel<float> : bf_tile_kernel<__half>;
        VELES_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const uint64_t row_tiles = (ix->n + kRT - 1) / kRT;
        uint64_t gx = (row_tiles + kWarps - 1) / kWarps;
        // enough CTAs to fill the machine a few times over; the row loop is grid-strided
        const uint64_t want = std::max<uint64_t>(1, ((uint64_t)sms * 4 + qtiles - 1) / qtiles);
        gx = std::max<uint64_t>(1, std::min(gx, want));
        // gridDim.y is limited to 65535: chunk the query tiles
        for (uint32_t t0 = 0; t0 < qtiles; t0 += 32768) {
            const uint32_t nt = std::min(32768u, qtiles - t0);
            const uint32_t qoff = t0 * kQT;
            dim3 grid((unsigned)gx, nt);
            kern<<<grid, kWarps * 32, smem, st>>>(v, q_d + (size_t)qoff * ix->dim, nq - qoff, 0, sink.at(qoff), as_value);
            count_launch();
        }
    } else {
        auto kern = ix->dtype == VELES_F32 ? bf_generic_kernel<float> : bf_generic_kernel<__half>;
        for (uint32_t q0 = 0; q0 < nq; q0 += 32768) {
            const uint32_t nn = std::min(32768u, nq - q0);
            dim3 grid((unsigned)std::min<uint64_t>((ix->n + 7) / 8 + 1, (uint64_t)sms * 4), nn);
            kern<<<grid, 256, (size_t)ix->dim * 4, st>>>(v, q_d + (size_t)q0 * ix->dim, nn, sink.at(q0), as_value);
            count_launch();
        }
    }
    VELES_CUDA(cudaGetLastError());
    return VELES_OK;
}

// per query: order bits of the kp-th best of its `m` sample scores (0xffffffff = no bound: fewer than kp samples)
__global__ void bf_threshold_kernel(const float* __restrict__ scores, uint64_t ld, uint32_t m, uint32_t nq, uint32_t kp, bool descending,
                                    uint32_t* __restrict__ thr) {
    extern __shared__ __align__(16) uint64_t bt_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + warp;
    if (q >= nq) return;
    uint64_t* res = bt_smem + (size_t)warp * kp;
    uint32_t len = 0;
    uint64_t worst = ~0ull;
    for (uint32_t base = 0; base < m; base += 32) {
        const uint32_t i = base + lane;
        uint64_t key = ~0ull;
        if (i < m) {
            uint32_t o = ord_key(scores[(size_t)q * ld + i]);
            if (descending) o = ~o;
            key = ((uint64_t)o << 32) | i;
        }
        if (__any_sync(FULL_MASK, key < worst)) list_offer(res, len, worst, kp, key, lane);
    }
    if (lane == 0) thr[q] = len == kp ? (uint32_t)(res[kp - 1] >> 32) : 0xffffffffu;
}

// per query: the k best of its candidate keys, written as (row, score); one warp per query
__global__ void bf_select_kernel(const uint64_t* __restrict__ cand, const uint32_t* __restrict__ cnt, uint32_t cap, uint32_t nq, uint32_t k,
                                 bool descending, uint32_t* __restrict__ out_ids, float* __restrict__ out_score) {
    extern __shared__ __align__(16) uint64_t bs_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + warp;
    if (q >= nq) return;
    uint64_t* res = bs_smem + (size_t)warp * k;
    uint32_t len = 0;
    uint64_t worst = ~0ull;
    const uint32_t m = min(cnt[q], cap);
    const uint64_t* in = cand + (size_t)q * cap;
    for (uint32_t base = 0; base < m; base += 32) {
        const uint32_t i = base + lane;
        const uint64_t key = i < m ? in[i] : ~0ull;
        if (__any_sync(FULL_MASK, key < worst)) list_offer(res, len, worst, k, key, lane);
    }
    __syncwarp();
    for (uint32_t i = lane; i < k; i += 32) {
        uint32_t id = VELES_INVALID_ID;
        float sc = __uint_as_float(0x7fc00000u);
        if (i < len) {
            id = (uint32_t)res[i];
            const uint32_t o = (uint32_t)(res[i] >> 32);
            sc = ord_unkey(descending ? ~o : o);
        }
        out_ids[(size_t)q * k + i] = id;
        out_score[(size_t)q * k + i] = sc;
    }
}

// The exact scan without the [nq, n] score matrix (north_star: selection merged into the distance kernel).
//   <= 8 queries   bf_scan_kernel keeps per-warp lists and finishes inside the same launch (ScanFuse)
//   more queries   two passes of the same exact kernels: a strided sample of the collection into a small matrix ->
//                  per query, the key of its kp-th best sample (a bound no true top-k row can miss), then the full
//                  scan whose results go through the bound into per-query candidate lists (ScoreSink FILTER), then
//                  bf_select_kernel.  Same scores, same order as the matrix path; a list overflow (a collection whose
//                  strided sample is unrepresentative) falls back to the matrix path.
// Returns 1 when the fused path ran, 0 when the caller should use the matrix path.
static int32_t bruteforce_fused(const veles_index* ix, const float* q_d, uint32_t nq, uint32_t k, uint32_t* ids_d, float* score_d,
                                bool desc, cudaStream_t st, int* ran) {
    *ran = 0;
    const int sms = device_sm_count();
    const IndexView v = ix->view();
    const bool fast_metric = ix->dtype != VELES_BIN1 && ix->dim >= 16 &&
                             (ix->metric == VELES_COSINE || ix->metric == VELES_EUCLIDEAN || ix->metric == VELES_DOT);
    if (std::getenv("VELES_BF_NO_FUSE")) return VELES_OK;
    // (small collections with more than two queries: the per-query selection tail outweighs the one launch it saves --
    // measured 10K x 768, 8 queries: 152 us fused against 79 us through the L2-resident score matrix)
    if (fast_metric && ix->dim % 32 == 0 && nq <= 8 && k <= 128 && (nq <= 2 || ix->n >= 65536)) {
        const uint32_t qt = nq <= 1 ? 1 : nq <= 2 ? 2 : nq <= 4 ? 4 : 8;
        if ((size_t)kWarps * qt * k * 8 > 96 * 1024) return VELES_OK;
        const size_t max_ctas = (size_t)sms * 16;
        VELES_TRY(ix->topk_d.ensure((size_t)nq * max_ctas * k * 8 + 64));
        if (!ix->bf_done.p) {
            VELES_TRY(ix->bf_done.alloc(64));
            VELES_CUDA(cudaMemsetAsync(ix->bf_done.p, 0, 64, st));
        }
        ScanFuse fz;
        fz.k = k;
        fz.desc = desc ? 1u : 0u;
        fz.partial = ix->topk_d.as<uint64_t>();
        fz.done = ix->bf_done.as<uint32_t>();
        fz.out_ids = ids_d;
        fz.out_score = score_d;
        VELES_TRY(launch_scores(ix, v, q_d, nq, ScoreSink(), true, st, &fz));
        *ran = 1;
        return VELES_OK;
    }
    if (nq <= 8 || ix->n < 65536 || k > 1024) return VELES_OK;
    const uint32_t kp = std::max(k, 16u);
    // sample every S-th row; ~n/64 rows but at least 8192: expected candidates per query ~ kp * S
    const uint64_t s_rows = std::min<uint64_t>(ix->n, std::max<uint64_t>(8192, ix->n / 64));
    const uint64_t S = ix->n / s_rows;
    if (S < 4 || (uint64_t)ix->row_bytes * S > 0xffffffffull) return VELES_OK;
    const uint64_t m = ix->n / S;
    const uint32_t cap = (uint32_t)std::min<uint64_t>(ix->n, 4 * (uint64_t)kp * S + 256);
    // queries per pass: sample matrix <= 256 MiB, candidate lists <= 512 MiB
    const uint32_t chunk = (uint32_t)std::max<uint64_t>(
        9, std::min<uint64_t>(nq, std::min<uint64_t>((256ull << 20) / (m * 4), (512ull << 20) / ((uint64_t)cap * 8))));
    if (chunk <= 8) return VELES_OK;
    VELES_TRY(ix->scores_d.ensure((size_t)chunk * m * 4));
    VELES_TRY(ix->topk_d.ensure((size_t)chunk * cap * 8));
    VELES_TRY(ix->bf_aux.ensure((size_t)chunk * 8 + 64));
    uint32_t* thr = ix->bf_aux.as<uint32_t>();
    uint32_t* cnt = thr + chunk;
    uint32_t* err = cnt + chunk;
    IndexView vs = v;
    vs.row_bytes = (uint32_t)(ix->row_bytes * S);
    vs.n = m;
    // equal passes, so that none falls into the <= 8 query (scan kernel) regime
    const uint32_t passes = (nq + chunk - 1) / chunk;
    if (nq / passes < 9) return VELES_OK;
    uint32_t q0 = 0;
    for (uint32_t pass = 0; pass < passes; ++pass) {
        const uint32_t nn = nq / passes + (pass < nq % passes ? 1u : 0u);
        VELES_CUDA(cudaMemsetAsync(cnt, 0, (size_t)chunk * 4 + 16, st));
        ScoreSink store;
        store.scores = ix->scores_d.as<float>();
        store.ld = m;
        VELES_TRY(launch_scores(ix, vs, q_d + (size_t)q0 * ix->dim, nn, store, true, st));
        const uint32_t tw = kp <= 128 ? 8 : 2;
        bf_threshold_kernel<<<(nn + tw - 1) / tw, tw * 32, (size_t)tw * kp * 8, st>>>(ix->scores_d.as<float>(), m, (uint32_t)m, nn, kp, desc, thr);
        count_launch();
        ScoreSink filt;
        filt.thr = thr;
        filt.cnt = cnt;
        filt.cand = ix->topk_d.as<uint64_t>();
        filt.cap = cap;
        filt.err = err;
        filt.desc = desc ? 1u : 0u;
        VELES_TRY(launch_scores(ix, v, q_d + (size_t)q0 * ix->dim, nn, filt, true, st));
        const uint32_t sw = k <= 128 ? 8 : 2;
        bf_select_kernel<<<(nn + sw - 1) / sw, sw * 32, (size_t)sw * k * 8, st>>>(ix->topk_d.as<uint64_t>(), cnt, cap, nn, k, desc,
                                                                                ids_d + (size_t)q0 * k, score_d + (size_t)q0 * k);
        count_launch();
        VELES_CUDA(cudaGetLastError());
        uint32_t h = 0;
        VELES_CUDA(cudaMemcpyAsync(&h, err, 4, cudaMemcpyDeviceToHost, st));
        VELES_CUDA(cudaStreamSynchronize(st));
        if (h != 0) return VELES_OK;  // a candidate list overflowed: the caller redoes the batch on the matrix path
        q0 += nn;
    }
    *ran = 1;
    return VELES_OK;
}

static int32_t bruteforce_device(const veles_index* ix, const float* q_d, uint32_t nq, uint32_t k, uint32_t* ids_d,
                                 float* score_d, cudaStream_t st) {
    NvtxRange nvtx_range("veles::bruteforce (HnswIndex::search_brute_force)");
    VELES_REQUIRE(k >= 1 && k <= 16384, "k must be in 1..16384, got %u", k);
    if (nq == 0) return VELES_OK;
    if (ix->n == 0) {
        VELES_CUDA(cudaMemsetAsync(ids_d, 0xff, (size_t)nq * k * 4, st));
        VELES_CUDA(cudaMemsetAsync(score_d, 0xff, (size_t)nq * k * 4, st));  // 0xffffffff is a NaN
        return VELES_OK;
    }
    const bool desc = ix->metric == VELES_COSINE || ix->metric == VELES_DOT || ix->metric == VELES_JACCARD;
    {
        int ran = 0;
        VELES_TRY(bruteforce_fused(ix, q_d, nq, k, ids_d, score_d, desc, st, &ran));
        if (ran) return VELES_OK;
    }
    // matrix path: [nq, n] metric values (<= ~1 GiB per pass), then a top-k kernel per query
    const uint64_t per_q = ix->n * 4;
    const uint32_t chunk = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(nq, (1ull << 30) / per_q));
    VELES_TRY(ix->scores_d.ensure((size_t)chunk * per_q));
    VELES_CUDA(cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(k * 8)));
    if (k <= 512)
        VELES_CUDA(cudaFuncSetAttribute(topk_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kTopWarps * k * 8)));
    for (uint32_t q0 = 0; q0 < nq; q0 += chunk) {
        const uint32_t nn = std::min(chunk, nq - q0);
        ScoreSink store;
        store.scores = ix->scores_d.as<float>();
        store.ld = ix->n;
        VELES_TRY(launch_scores(ix, ix->view(), q_d + (size_t)q0 * ix->dim, nn, store, true, st));
        // few queries over many rows: split every row into chunks (first level), then merge (second level)
        const int sms = device_sm_count();
        uint32_t chunks = 1;
        if (k <= 512 && (uint64_t)nn < (uint64_t)sms * 2 && ix->n >= 65536) {
            const uint64_t want = ((uint64_t)sms * 4 + nn - 1) / nn;
            const uint64_t most = std::max<uint64_t>(1, ix->n / std::max<uint64_t>(16384, (uint64_t)k * 64));
            chunks = (uint32_t)std::min<uint64_t>(std::min(want, most), 4096);
        }
        if (chunks > 1) {
            uint64_t chunk = (ix->n + chunks - 1) / chunks;
            chunk = (chunk + 127) & ~(uint64_t)127;  // keeps float4 alignment of every chunk start
            chunks = (uint32_t)((ix->n + chunk - 1) / chunk);
            VELES_TRY(ix->topk_d.ensure((size_t)nn * chunks * k * 8));
            dim3 grid(nn, chunks);
            topk_cta_kernel<<<grid, kTopWarps * 32, (size_t)kTopWarps * k * 8, st>>>(ix->scores_d.as<float>(), ix->n, k, desc, nullptr,
                                                                                    nullptr, chunk, ix->topk_d.as<uint64_t>());
            count_launch();
            topk_merge_kernel<<<nn, 32, (size_t)k * 8, st>>>(ix->topk_d.as<uint64_t>(), chunks, k, ix->scores_d.as<float>(), ix->n,
                                                             ids_d + (size_t)q0 * k, score_d + (size_t)q0 * k);
        } else if (k <= 512) {
            topk_cta_kernel<<<nn, kTopWarps * 32, (size_t)kTopWarps * k * 8, st>>>(ix->scores_d.as<float>(), ix->n, k, desc,
                                                                                  ids_d + (size_t)q0 * k, score_d + (size_t)q0 * k, 0,
                                                                                  nullptr);
        } else {
            topk_kernel<<<nn, 32, (size_t)k * 8, st>>>(ix->scores_d.as<float>(), ix->n, k, desc, ids_d + (size_t)q0 * k,
                                                       score_d + (size_t)q0 * k);
        }
        count_launch();
        VELES_CUDA(cudaGetLastError());
    }
    return VELES_OK;
}

}  // namespace veles

using namespace veles;

extern "C" {

int32_t veles_bruteforce_batch_d(const veles_index_t* idx, const float* queries_d, uint32_t nq, uint32_t k,
                                 uint32_t* out_ids_d, float* out_score_d, void* stream) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    VELES_REQUIRE(nq == 0 || (queries_d && out_ids_d && out_score_d), "NULL buffer");
    std::lock_guard<std::mutex> g(idx->mu);
    return bruteforce_device(idx, queries_d, nq, k, out_ids_d, out_score_d, (cudaStream_t)stream);
}

int32_t veles_bruteforce_batch(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k,
                               uint32_t* out_ids, float* out_score, void* stream) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    VELES_REQUIRE(nq == 0 || (queries && out_ids && out_score), "NULL buffer");
    if (nq == 0) return VELES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> g(idx->mu);
    const size_t qb = (size_t)nq * idx->dim * 4, ob = (size_t)nq * k * 4;
    VELES_TRY(idx->q_d.ensure(qb));
    VELES_TRY(idx->out_ids_d.ensure(ob));
    VELES_TRY(idx->out_val_d.ensure(ob));
    VELES_CUDA(cudaMemcpyAsync(idx->q_d.p, queries, qb, cudaMemcpyHostToDevice, st));
    VELES_TRY(bruteforce_device(idx, idx->q_d.as<float>(), nq, k, idx->out_ids_d.as<uint32_t>(), idx->out_val_d.as<float>(), st));
    VELES_CUDA(cudaMemcpyAsync(out_ids, idx->out_ids_d.p, ob, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaMemcpyAsync(out_score, idx->out_val_d.p, ob, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaStreamSynchronize(st));
    return VELES_OK;
}

int32_t veles_rerank_batch(const veles_index_t* idx, const float* queries, uint32_t nq, const uint32_t* cand, uint32_t m,
                           float* out_score, void* stream) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    VELES_REQUIRE(nq == 0 || m == 0 || (queries && cand && out_score), "NULL buffer");
    if (nq == 0 || m == 0) return VELES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> g(idx->mu);
    const size_t qb = (size_t)nq * idx->dim * 4, cb = (size_t)nq * m * 4;
    VELES_TRY(idx->q_d.ensure(qb));
    VELES_TRY(idx->out_ids_d.ensure(cb));
    VELES_TRY(idx->out_val_d.ensure(cb));
    VELES_CUDA(cudaMemcpyAsync(idx->q_d.p, queries, qb, cudaMemcpyHostToDevice, st));
    VELES_CUDA(cudaMemcpyAsync(idx->out_ids_d.p, cand, cb, cudaMemcpyHostToDevice, st));
    const int sms = device_sm_count();
    const unsigned grid = (unsigned)std::min<uint64_t>(((uint64_t)nq * m + 7) / 8, (uint64_t)sms * 8);
    if (idx->dtype == VELES_F16)
        rerank_kernel<__half><<<grid, 256, 0, st>>>(idx->view(), idx->q_d.as<float>(), nq, idx->out_ids_d.as<uint32_t>(), m,
                                                    idx->out_val_d.as<float>());
    else
        rerank_kernel<float><<<grid, 256, 0, st>>>(idx->view(), idx->q_d.as<float>(), nq, idx->out_ids_d.as<uint32_t>(), m,
                                                   idx->out_val_d.as<float>());
    count_launch();
    VELES_CUDA(cudaGetLastError());
    VELES_CUDA(cudaMemcpyAsync(out_score, idx->out_val_d.p, cb, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaStreamSynchronize(st));
    return VELES_OK;
}

int32_t veles_distance_pairs(int32_t metric, const float* a, const float* b, uint32_t n_pairs, uint32_t dim,
                             int32_t as_metric_value, float* out, void* stream) {
    VELES_REQUIRE(metric >= VELES_COSINE && metric <= VELES_JACCARD, "unknown metric %d", metric);
    VELES_REQUIRE(n_pairs == 0 || (a && b && out), "NULL buffer");
    VELES_REQUIRE(dim >= 1, "dim must be >= 1");
    if (n_pairs == 0) return VELES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf da, db, dout;
    const size_t bytes = (size_t)n_pairs * dim * 4;
    VELES_TRY(da.alloc(bytes));
    VELES_TRY(db.alloc(bytes));
    VELES_TRY(dout.alloc((size_t)n_pairs * 4));
    VELES_CUDA(cudaMemcpyAsync(da.p, a, bytes, cudaMemcpyHostToDevice, st));
    VELES_CUDA(cudaMemcpyAsync(db.p, b, bytes, cudaMemcpyHostToDevice, st));
    const int sms = device_sm_count();
    const unsigned grid = (unsigned)std::min<uint64_t>(((uint64_t)n_pairs + 7) / 8, (uint64_t)sms * 8);
    pairs_kernel<<<grid, 256, 0, st>>>(metric, da.as<float>(), db.as<float>(), n_pairs, dim, as_metric_value != 0,
                                       dout.as<float>());
    count_launch();
    VELES_CUDA(cudaGetLastError());
    VELES_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaStreamSynchronize(st));
    return VELES_OK;
}

}  // extern "C"

// gemm_tc.cu -- relaxed-mode brute force: candidates from a hand-written fp16 tensor-core GEMM (tcgen05.mma, accumulators
// in TMEM, operands staged by TMA), then an exact f32 re-rank of the candidates.
//
// Role: the many-query brute-force fallback of the path -- HnswIndex::brute_force_search_parallel
// (index/hnsw/index/batch.rs:223-244), the reference's own GPU hook search_brute_force_gpu
// (index/hnsw/index/search.rs:229-279 -> gpu/gpu_backend.rs:300-340, a WGSL batch-cosine shader) and the re-rank of
// search_with_rerank (search.rs:118-160).  The exact kernels (bruteforce.cu) keep the reference's f32 summation order
// and run on the FMA pipe; this path gives the order up for the candidate stage only:
//
//   1. X16 / Q16   the collection and the queries rounded to fp16, K padded to a multiple of 64 (cosine: rows
//                  pre-normalised, so the GEMM yields the cosine; L2: 2*dot - |x|^2 in the epilogue)
//   2. sample      gemm_tc_kernel in STORE mode over a few row tiles spread over the collection -> per query, the
//                  k'-th best sampled score (k' = k * oversample).  The k'-th best of a subset can only be worse than the
//                  k'-th best of the whole collection, so this threshold never cuts a true top-k' row.
//   3. filter      gemm_tc_kernel in FILTER mode over every row tile: the epilogue reads the accumulators from TMEM and
//                  appends (row, score) to the query's candidate list when score >= threshold -- the [nq, n] score
//                  matrix never exists
//   4. finish      relaxed_finish_kernel: the exact metric value (reference accumulation tree, common.cuh) of every
//                  candidate, top-k by DistanceMetric::sort_results order.  Returned scores are exact; only the
//                  candidate set is fp16-filtered, so this mode is judged by recall, not by id equality.
//
// gemm_tc_kernel: D[128 rows, BN queries] += A[128, 64] * B[BN, 64]^T per k-block; one CTA per SM, persistent over
// (row tile, query block) work items; warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM
// allocator, warps 4-7 = epilogue (one TMEM lane = one collection row per thread).  Shared-memory ring of
// STAGES x (A 16 KB + B BN*128 B) in the 128-byte-swizzled K-major layout TMA writes and tcgen05.mma reads; two
// accumulator stages in TMEM (2 * BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>

#include <algorithm>
#include <cmath>

#include "index.hpp"

namespace veles {

constexpr uint32_t kTcBlockM = 128;
constexpr uint32_t kTcBlockK = 64;          // halves = one 128-byte swizzle atom
constexpr uint32_t kTcSampleTiles = 32;     // row tiles of the threshold sample (4096 rows) for small collections
constexpr long long kTcSpinLimit = 400000000ll;  // mbarrier wait bound (cycles): a broken pipeline must not hang the GPU

struct TcParams {
    uint32_t n_rows, nq, k_blocks, n_mtiles, n_nblocks;
    const uint32_t* tile_list;  // m-tile indices to process (sample pass) or null = all
    int mode;                   // 0 = STORE scores, 1 = FILTER candidates
    float* out;                 // STORE: out[q * ld_out + tile_pos * 128 + r]
    uint32_t ld_out;
    const float* thr;           // FILTER: per query threshold
    uint32_t* cand_cnt;         // FILTER: gridDim.x x nq counts -- every CTA owns a segment of every query's list,
    uint64_t* cand;             // FILTER: gridDim.x x nq x cand_cap entries (score bits << 32 | row)
    uint32_t cand_cap;          //         so appends are shared-memory atomics + a store, never a global round trip
    const float* row_bias;      // null, or per row: score = scale * acc - row_bias[row]
    float scale;
    uint32_t* err;              // [0] pipeline timeout, [1] candidate overflow
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait; returns false (and raises the error flag) on timeout
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, uint32_t* err) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > kTcSpinLimit) {
            atomicExch(err, 1u);
            return false;
        }
    }
    return true;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint32_t c0, uint32_t c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// tcgen05.commit: the barrier gets one arrival when every MMA issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, fp16 inputs, f32 accumulate, M = 128, K = 16
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile in the 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1 (Blackwell)
__device__ __forceinline__ uint64_t tc_smem_desc(const void* tile) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(tile) & 0x3ffffu) >> 4);  // start address, bits [0, 14)
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused with a swizzled K-major tile)
    d |= (uint64_t)(1024u >> 4) << 32;                  // stride byte offset, bits [32, 46)
    d |= (uint64_t)1 << 46;                             // descriptor version
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}
// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN>
struct TcCfg {
    static constexpr uint32_t kStages = BN == 256 ? 4 : 6;
    static constexpr uint32_t kABytes = kTcBlockM * 128;
    static constexpr uint32_t kBBytes = BN * 128;
    static constexpr uint32_t kStageBytes = kABytes + kBBytes;
    static constexpr uint32_t kTmemCols = 2 * BN;  // two accumulator stages: 256 or 512 columns (powers of two)
    static constexpr uint32_t kMaxQ = 1024;  // queries per launch (per-CTA candidate counters live in shared memory)
    static constexpr uint32_t kSmemBytes =
        1024 /* alignment slack */ + kStages * kStageBytes + 256 /* barriers */ + 2 * BN * 4 + kMaxQ * 4;
};

template <int BN>
__global__ void __launch_bounds__(256, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                                                         const TcParams p) {
    using C = TcCfg<BN>;
    // 1024-byte alignment (128-byte swizzle atoms) is requested from the declaration, not by rounding a pointer: the
    // compiler then still knows these are shared-memory addresses (LDS/STS instead of generic loads in the epilogue)
    extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
    uint8_t* tiles = tc_smem_raw;
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + C::kStages * C::kStageBytes);
    uint64_t* empty = full + C::kStages;
    uint64_t* tfull = empty + C::kStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* thr_s = reinterpret_cast<float*>(tiles + C::kStages * C::kStageBytes + 256);  // 2 x BN
    uint32_t* cnt_s = reinterpret_cast<uint32_t*>(thr_s + 2 * BN);                       // kMaxQ: this CTA's candidates per query
    for (uint32_t i = threadIdx.x; i < C::kMaxQ; i += blockDim.x) cnt_s[i] = 0;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if ((smem_u32(tiles) & 1023u) != 0) {  // the swizzled tiles need the alignment the declaration asks for
        if (threadIdx.x == 0) atomicExch(p.err, 2u);
        return;
    }

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b) : "memory");
        for (uint32_t i = 0; i < C::kStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (uint32_t i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);  // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // this CTA's row tiles: blockIdx.x, blockIdx.x + gridDim.x, ...
    const uint32_t my_tiles = p.n_mtiles > blockIdx.x ? (p.n_mtiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    const uint32_t total = my_tiles * p.n_nblocks;

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            uint32_t stage = 0, phase = 0;
            bool ok = true;
            // work items: every row tile of this CTA (blockIdx.x, + gridDim.x, ...) against every query block, query
            // blocks innermost -- the A tile is re-read from L2 by the same SM, and every CTA sees every query, so the
            // candidates of a query spread evenly over the CTAs' segments
            for (uint32_t w = 0; w < total && ok; ++w) {
                const uint32_t mpos = blockIdx.x + (w / p.n_nblocks) * gridDim.x, nb = w % p.n_nblocks;
                const uint32_t mt = p.tile_list ? p.tile_list[mpos] : mpos;
                for (uint32_t kb = 0; kb < p.k_blocks && ok; ++kb) {
                    ok = mbar_wait_bounded(&empty[stage], phase ^ 1u, p.err);
                    uint8_t* a = tiles + stage * C::kStageBytes;
                    mbar_expect_tx(&full[stage], C::kStageBytes);
                    tma_load_2d(a, &tm_a, kb * kTcBlockK, mt * kTcBlockM, &full[stage]);
                    tma_load_2d(a + C::kABytes, &tm_b, kb * kTcBlockK, nb * BN, &full[stage]);
                    if (++stage == C::kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            // instruction descriptor: D = f32, A = B = f16, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((kTcBlockM >> 4) << 24);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            bool ok = true;
            for (uint32_t w = 0; w < total && ok; ++w) {
                ok = mbar_wait_bounded(&tempty[acc], acc_phase ^ 1u, p.err);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (uint32_t kb = 0; kb < p.k_blocks && ok; ++kb) {
                    ok = mbar_wait_bounded(&full[stage], phase, p.err);
                    tc_fence_after();
                    const uint8_t* a = tiles + stage * C::kStageBytes;
                    const uint64_t a_desc = tc_smem_desc(a), b_desc = tc_smem_desc(a + C::kABytes);
#pragma unroll
                    for (uint32_t k = 0; k < kTcBlockK / 16; ++k)  // +32 bytes along K inside the swizzle atom per step
                        tc_mma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(&empty[stage]);  // frees the stage when these MMAs have read it
                    if (++stage == C::kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                tc_commit(&tfull[acc]);  // accumulator complete
                acc ^= 1u;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4) {  // ===== epilogue: TMEM -> registers -> scores / candidates =====
        const uint32_t ew = warp - 4;  // TMEM lanes 32*ew .. 32*ew+31 are this warp's
        const uint32_t et = threadIdx.x - 128;
        uint32_t acc = 0, acc_phase = 0;
        bool ok = true;
        for (uint32_t w = 0; w < total; ++w) {
            const uint32_t mpos = blockIdx.x + (w / p.n_nblocks) * gridDim.x, nb = w % p.n_nblocks;
            const uint32_t mt = p.tile_list ? p.tile_list[mpos] : mpos;
            float* thr = thr_s + acc * BN;
            if (p.mode == 1) {
                for (uint32_t c = et; c < (uint32_t)BN; c += 128) {
                    const uint32_t q = nb * BN + c;
                    thr[c] = q < p.nq ? p.thr[q] : __int_as_float(0x7f800000);
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
            if (ok) ok = mbar_wait_bounded(&tfull[acc], acc_phase, p.err);
            tc_fence_after();
            const uint32_t r_in_tile = ew * 32 + lane;
            const uint32_t row = mt * kTcBlockM + r_in_tile;
            const bool row_ok = row < p.n_rows;
            const float bias = (p.row_bias && row_ok) ? p.row_bias[row] : 0.0f;
            const bool affine = p.row_bias != nullptr || p.scale != 1.0f;
            for (uint32_t c0 = 0; c0 < (uint32_t)BN; c0 += 32) {
                uint32_t v[32];
                tc_ld32(tmem_base + acc * BN + c0 + ((ew * 32u) << 16), v);
                if (p.mode == 1) {
                    // FILTER: one compare per column, no branches -- a single warp per scheduler cannot hide the latency
                    // of a branchy per-column chain (ncu, first version: the epilogue paced the whole kernel).  The pass
                    // mask is almost always zero; the rare set bits are appended afterwards.
                    if (affine) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(p.scale * __uint_as_float(v[j]) - bias);
                    }
                    uint32_t mask = 0;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 t4 = *reinterpret_cast<const float4*>(thr + c0 + 4 * j4);
                        mask |= (__uint_as_float(v[4 * j4 + 0]) >= t4.x ? 1u : 0u) << (4 * j4 + 0);
                        mask |= (__uint_as_float(v[4 * j4 + 1]) >= t4.y ? 1u : 0u) << (4 * j4 + 1);
                        mask |= (__uint_as_float(v[4 * j4 + 2]) >= t4.z ? 1u : 0u) << (4 * j4 + 2);
                        mask |= (__uint_as_float(v[4 * j4 + 3]) >= t4.w ? 1u : 0u) << (4 * j4 + 3);
                    }
                    if (!row_ok) mask = 0;
                    while (mask) {
                        const int j = __ffs(mask) - 1;
                        mask &= mask - 1;
                        uint32_t sv = 0;
#pragma unroll
                        for (int u = 0; u < 32; ++u) sv = u == j ? v[u] : sv;  // select, not an indexed access: v[] stays in registers
                        const uint32_t q = nb * BN + c0 + j;
                        const uint32_t slot = atomicAdd(&cnt_s[q], 1u);  // shared memory: no DRAM round trip
                        if (slot < p.cand_cap)
                            p.cand[((size_t)blockIdx.x * p.nq + q) * p.cand_cap + slot] = ((uint64_t)sv << 32) | row;
                        else
                            atomicExch(p.err + 1, 1u);
                    }
                    continue;
                }
                if (p.mode == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint32_t q = nb * BN + c0 + j;
                        if (q < p.nq) p.out[(size_t)q * p.ld_out + (size_t)mpos * kTcBlockM + r_in_tile] =
                            row_ok ? p.scale * __uint_as_float(v[j]) - bias : __int_as_float(0xff800000);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.mode == 1)
        for (uint32_t q = threadIdx.x; q < p.nq; q += blockDim.x) p.cand_cnt[(size_t)blockIdx.x * p.nq + q] = min(cnt_s[q], p.cand_cap);
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::kTmemCols) : "memory");
    }
}

// ---- fp16 operands -----------------------------------------------------------------------------------------------
// rows of `src` (row stride src_stride bytes, f32 or f16 elements) -> fp16 [rows, dpad], zero padded; cosine: scaled by
// 1/|row|; bias (L2): |row16|^2 of the rounded values
__global__ void to_f16_operand_kernel(const uint8_t* __restrict__ src, uint64_t src_stride, int src_f16, uint32_t norm_off,
                                      int normalize, uint64_t rows, uint32_t dim, uint32_t dpad, __half* __restrict__ out,
                                      float* __restrict__ bias) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = gw; r < rows; r += nwarps) {
        const uint8_t* row = src + r * src_stride;
        float scale = 1.0f;
        if (normalize) {
            float nb;
            if (norm_off != 0xffffffffu) {
                nb = *reinterpret_cast<const float*>(row + norm_off);
            } else {  // queries: no stored norm
                float ss = 0.0f;
                for (uint32_t i = lane; i < dim; i += 32) {
                    const float x = reinterpret_cast<const float*>(row)[i];
                    ss += x * x;
                }
                for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
                nb = sqrtf(ss);
            }
            scale = nb > 0.0f ? 1.0f / nb : 0.0f;
        }
        float ss = 0.0f;
        for (uint32_t i = lane; i < dpad; i += 32) {
            float x = 0.0f;
            if (i < dim) x = src_f16 ? __half2float(reinterpret_cast<const __half*>(row)[i]) : reinterpret_cast<const float*>(row)[i];
            const __half h = __float2half_rn(x * scale);
            out[r * dpad + i] = h;
            const float hf = __half2float(h);
            ss += hf * hf;
        }
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
        if (lane == 0 && bias) bias[r] = ss;
    }
}

// per query: the kp-th largest of its `m` sampled scores (-inf when fewer than kp are finite); one warp per query
__global__ void tc_threshold_kernel(const float* __restrict__ scores, uint32_t ld, uint32_t m, uint32_t nq, uint32_t kp,
                                    float* __restrict__ thr) {
    extern __shared__ __align__(16) uint64_t th_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + warp;
    if (q >= nq) return;
    uint64_t* res = th_smem + (size_t)warp * kp;  // ascending keys = descending scores
    uint32_t len = 0;
    uint64_t worst = ~0ull;
    for (uint32_t base = 0; base < m; base += 32) {
        const uint32_t i = base + lane;
        uint64_t key = ~0ull;
        if (i < m) {
            const float s = scores[(size_t)q * ld + i];
            if (s > __int_as_float(0xff800000)) key = ((uint64_t)(~ord_key(s)) << 32) | i;
        }
        uint32_t msk = __ballot_sync(FULL_MASK, key < worst);
        while (msk) {
            const uint32_t src = __ffs(msk) - 1;
            msk &= msk - 1;
            const uint64_t kk = __shfl_sync(FULL_MASK, key, src);
            if (kk >= worst) continue;
            const uint32_t pos = lower_bound_warp(res, len, kk, lane);
            if (len < kp) {
                insert_at(res, pos, len + 1, kk, lane);
                ++len;
            } else {
                insert_at(res, pos, len, kk, lane);
            }
            if (len == kp) worst = res[kp - 1];
        }
    }
    if (lane == 0) thr[q] = len == kp ? ord_unkey(~(uint32_t)(res[kp - 1] >> 32)) : __int_as_float(0xff800000);
}

// ---- CTA-wide selection of the kr smallest of N unique u64 keys (256 threads, kr <= 256) ----------------------------
// Pass 1: every thread takes the minimum of its strided share; the kr-th smallest of the 256 minima is an upper bound
// of the kr-th smallest key (kr distinct keys lie at or below it).  Pass 2 collects the keys at or below the bound
// (about -256 ln(1 - kr/256) of them when the keys are spread evenly: 96 for kr = 80), and each collected key's rank
// among the collected is its position in the sorted output.  ~2 N / 256 key evaluations and a few hundred shared-memory
// compares per thread, against one serial register-list insert (~100 warp instructions) per accepted key in the
// RegTopK scan -- which made the bound kernel and the finish kernel issue-bound (0.16 and 0.39 ms per 1024 queries).
// `key_at(i)` must return the same value in both passes, ~0 for "no key".  Returns false when more than `cap` keys
// fall under the bound (keys laid out so that the best ones share a thread): the caller takes the RegTopK path.
constexpr uint32_t kSelCap = 1024;
struct SelScratch {
    uint64_t mins[256];
    uint64_t col[kSelCap];
    uint64_t bound;
    uint32_t count;
};
template <typename F>
__device__ __forceinline__ bool cta_select_smallest(F&& key_at, uint32_t N, uint32_t kr, SelScratch& sc, uint64_t* out,
                                                    uint32_t* out_cnt) {
    const uint32_t tid = threadIdx.x;
    uint64_t mn = ~0ull;
#pragma unroll 4
    for (uint32_t i = tid; i < N; i += 256) {
        const uint64_t kk = key_at(i);
        mn = kk < mn ? kk : mn;
    }
    sc.mins[tid] = mn;
    if (tid == 0) sc.count = 0;
    __syncthreads();
    uint32_t r = 0;
    for (uint32_t j = 0; j < 256; ++j) {
        const uint64_t v = sc.mins[j];
        r += (v < mn || (v == mn && j < tid)) ? 1u : 0u;  // only the ~0 sentinels can be equal
    }
    if (r == kr - 1) sc.bound = mn;
    __syncthreads();
    const uint64_t bound = sc.bound;
#pragma unroll 4
    for (uint32_t i = tid; i < N; i += 256) {
        const uint64_t kk = key_at(i);
        if (kk <= bound && kk != ~0ull) {
            const uint32_t slot = atomicAdd(&sc.count, 1u);
            if (slot < kSelCap) sc.col[slot] = kk;
        }
    }
    __syncthreads();
    const uint32_t m = sc.count;
    if (m > kSelCap) return false;
    for (uint32_t j = tid; j < m; j += 256) {
        const uint64_t kj = sc.col[j];
        uint32_t rank = 0;
        for (uint32_t t = 0; t < m; ++t) rank += sc.col[t] < kj ? 1u : 0u;
        if (rank < kr) out[rank] = kj;
    }
    if (tid == 0) *out_cnt = m < kr ? m : kr;
    __syncthreads();
    return true;
}

// The same bound with a CTA per query (kp <= 128): cta_select_smallest over the query's sampled scores; when that
// overflows, the eight warps scan interleaved 256-score blocks with eight independent coalesced loads in flight per lane,
// each keeping a register-resident sorted list (RegTopK), and warp 0 merges the lists.  The one-warp kernel above walks
// its row one dependent 128-byte load at a time: 1.1 ms for 1024 queries x 62 K sampled scores -- as long as the filter
// GEMM itself.
template <int R>
__global__ void __launch_bounds__(256) tc_threshold_cta_kernel(const float* __restrict__ scores, uint32_t ld, uint32_t m, uint32_t kp,
                                                               float* __restrict__ thr, int use_select) {
    __shared__ uint64_t s_keys[8][32 * R];
    __shared__ SelScratch s_sel;
    __shared__ uint32_t s_cnt;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q = blockIdx.x;
    const float* row = scores + (size_t)q * ld;
    const float ninf = __int_as_float(0xff800000);
    if (use_select) {
        uint64_t* out = &s_keys[0][0];  // kp <= 32 * R <= 8 * 32 * R entries
        const bool ok = cta_select_smallest(
            [&](uint32_t i) -> uint64_t {
                const float v = row[i];
                return v > ninf ? (((uint64_t)(~ord_key(v)) << 32) | i) : ~0ull;
            },
            m, kp, s_sel, out, &s_cnt);
        if (ok) {
            if (threadIdx.x == 0) thr[q] = s_cnt == kp ? ord_unkey(~(uint32_t)(out[kp - 1] >> 32)) : ninf;
            return;
        }
        __syncthreads();
    }
    RegTopK<R> top;
    top.init(kp, lane);
    constexpr uint32_t U = 8;
    for (uint32_t base = warp * 32 * U; base < m; base += 8 * 32 * U) {
        float v[U];
#pragma unroll
        for (uint32_t u = 0; u < U; ++u) {
            const uint32_t i = base + u * 32 + lane;
            v[u] = i < m ? row[i] : ninf;
        }
#pragma unroll
        for (uint32_t u = 0; u < U; ++u) {
            const uint32_t i = base + u * 32 + lane;
            top.offer(v[u] > ninf ? (((uint64_t)(~ord_key(v[u])) << 32) | i) : ~0ull);
        }
    }
    top.store(s_keys[warp], 32 * R);
    __syncthreads();
    if (warp != 0) return;
    for (uint32_t w = 1; w < 8; ++w)
        for (uint32_t j = 0; j < kp; j += 32) {
            const uint64_t key = j + lane < kp ? s_keys[w][j + lane] : ~0ull;
            if (__shfl_sync(FULL_MASK, key, 0) >= top.worst) break;  // the list is ascending: nothing further can enter
            top.offer(key);
        }
    if (lane == 0) thr[q] = top.worst != ~0ull ? ord_unkey(~(uint32_t)(top.worst >> 32)) : ninf;
}

// One CTA per query: exact metric value of every candidate (compute_distance, index/hnsw/index/search.rs:30-38, with the
// reference's accumulation tree), top-k in DistanceMetric::sort_results order (core/distance.rs:95-103), ties by row id.
constexpr int kFinWarps = 8;
__global__ void __launch_bounds__(kFinWarps * 32) relaxed_finish_kernel(IndexView ix, const float* __restrict__ queries, uint32_t nq,
                                                                      const uint64_t* __restrict__ cand,
                                                                      const uint32_t* __restrict__ cand_cnt, uint32_t cand_cap,
                                                                      uint32_t n_seg, uint32_t k, uint32_t* __restrict__ out_ids,
                                                                      float* __restrict__ out_score) {
    extern __shared__ __align__(16) uint8_t fin_smem[];
    float* qs = reinterpret_cast<float*>(fin_smem);                                       // dim
    uint64_t* lists = reinterpret_cast<uint64_t*>(fin_smem + ((ix.dim * 4 + 15) & ~15u));  // kFinWarps x k
    const uint32_t q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* qg = queries + (size_t)q * ix.dim;
    for (uint32_t i = threadIdx.x; i < ix.dim; i += blockDim.x) qs[i] = qg[i];
    __syncthreads();
    const bool desc = ix.metric == VELES_COSINE || ix.metric == VELES_DOT || ix.metric == VELES_JACCARD;
    float na = 0.0f;
    if (ix.metric == VELES_COSINE) na = __fsqrt_rn(warp_tree_reduce<0>(qs, qs, ix.dim, lane));
    uint64_t* mine = lists + (size_t)warp * k;
    uint32_t len = 0;
    // the query's candidates: one segment per GEMM CTA; this warp takes every kFinWarps-th candidate of the flattened list
    uint32_t flat = 0;
    for (uint32_t seg = 0; seg < n_seg; ++seg) {
      const uint32_t m = min(cand_cnt[(size_t)seg * nq + q], cand_cap);
      const uint64_t* base = cand + ((size_t)seg * nq + q) * cand_cap;
      for (uint32_t c = 0; c < m; ++c, ++flat) {
        if (flat % kFinWarps != warp) continue;
        const uint32_t row = (uint32_t)base[c];
        const uint8_t* rp = ix.vecs + (size_t)row * ix.row_bytes;
        const float nb = ix.metric == VELES_COSINE ? *reinterpret_cast<const float*>(rp + ix.norm_off) : 0.0f;
        const float v = ix.dtype == VELES_F32 ? warp_metric(ix.metric, true, qs, reinterpret_cast<const float*>(rp), ix.dim, na, nb, lane)
                                              : warp_metric(ix.metric, true, qs, reinterpret_cast<const __half*>(rp), ix.dim, na, nb, lane);
        const uint32_t ok = ord_key(v);
        const uint64_t key = ((uint64_t)(desc ? ~ok : ok) << 32) | row;
        if (len < k || key < mine[k - 1]) {
            const uint32_t pos = lower_bound_warp(mine, len, key, lane);
            if (len < k) {
                insert_at(mine, pos, len + 1, key, lane);
                ++len;
            } else {
                insert_at(mine, pos, len, key, lane);
            }
        }
        __syncwarp();
      }
    }
    __shared__ uint32_t s_len[kFinWarps];
    if (lane == 0) s_len[warp] = len;
    __syncthreads();
    if (warp != 0) return;
    for (uint32_t w = 1; w < kFinWarps; ++w) {
        const uint64_t* other = lists + (size_t)w * k;
        for (uint32_t j = 0; j < s_len[w]; ++j) {
            const uint64_t key = other[j];
            if (len == k && key >= mine[k - 1]) break;
            const uint32_t pos = lower_bound_warp(mine, len, key, lane);
            if (len < k) {
                insert_at(mine, pos, len + 1, key, lane);
                ++len;
            } else {
                insert_at(mine, pos, len, key, lane);
            }
            __syncwarp();
        }
    }
    for (uint32_t j = lane; j < k; j += 32) {
        uint32_t id = VELES_INVALID_ID;
        float sc = __uint_as_float(0x7fc00000u);
        if (j < len) {
            id = (uint32_t)mine[j];
            const uint32_t kb = (uint32_t)(mine[j] >> 32);
            sc = ord_unkey(desc ? ~kb : kb);
        }
        out_ids[(size_t)q * k + j] = id;
        out_score[(size_t)q * k + j] = sc;
    }
}

// relaxed_finish_kernel with the candidate walk flattened: the per-segment counts are read by all threads at once and
// prefix-summed (the loop above reads 148 counts one dependent load after the other, per warp), a warp's candidate f is
// located by a search in that prefix table, the next candidate's row id is fetched and its row prefetched into the L2
// while the current one is evaluated.  Same arithmetic, same keys, same merge: same results.
constexpr int kFinMaxSeg = 256;
__global__ void __launch_bounds__(kFinWarps * 32) relaxed_finish_flat_kernel(IndexView ix, const float* __restrict__ queries, uint32_t nq,
                                                                           const uint64_t* __restrict__ cand,
                                                                           const uint32_t* __restrict__ cand_cnt, uint32_t cand_cap,
                                                                           uint32_t n_seg, uint32_t k, uint32_t* __restrict__ out_ids,
                                                                           float* __restrict__ out_score) {
    extern __shared__ __align__(16) uint8_t fin_smem[];
    __shared__ uint32_t s_off[kFinMaxSeg + 1];
    __shared__ uint32_t s_len[kFinWarps];
    float* qs = reinterpret_cast<float*>(fin_smem);                                       // dim
    uint64_t* lists = reinterpret_cast<uint64_t*>(fin_smem + ((ix.dim * 4 + 15) & ~15u));  // kFinWarps x k
    const uint32_t q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* qg = queries + (size_t)q * ix.dim;
    for (uint32_t i = threadIdx.x; i < ix.dim; i += blockDim.x) qs[i] = qg[i];
    // exclusive prefix sum of the segment counts (n_seg <= 256 = one count per thread)
    {
        const uint32_t c = threadIdx.x < n_seg ? min(cand_cnt[(size_t)threadIdx.x * nq + q], cand_cap) : 0u;
        uint32_t x = c;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL_MASK, x, o);
            if ((int)lane >= o) x += y;
        }
        if (lane == 31) s_len[warp] = x;
        __syncthreads();
        uint32_t before = 0;
        for (uint32_t w = 0; w < warp; ++w) before += s_len[w];
        s_off[threadIdx.x + 1] = before + x;
        if (threadIdx.x == 0) s_off[0] = 0;
        __syncthreads();
    }
    const uint32_t total = s_off[kFinWarps * 32];
    const bool desc = ix.metric == VELES_COSINE || ix.metric == VELES_DOT || ix.metric == VELES_JACCARD;
    float na = 0.0f;
    if (ix.metric == VELES_COSINE) na = __fsqrt_rn(warp_tree_reduce<0>(qs, qs, ix.dim, lane));
    uint64_t* mine = lists + (size_t)warp * k;
    uint32_t len = 0;
    // row id of flattened candidate f: the segment is the last one whose offset is <= f
    auto row_of = [&](uint32_t f) -> uint32_t {
        uint32_t lo = 0, hi = kFinWarps * 32;  // s_off[lo] <= f < s_off[hi]
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_off[mid] <= f)
                lo = mid;
            else
                hi = mid;
        }
        return (uint32_t)cand[((size_t)lo * nq + q) * cand_cap + (f - s_off[lo])];
    };
    uint32_t row = warp < total ? row_of(warp) : 0u;
    for (uint32_t f = warp; f < total; f += kFinWarps) {
        const uint32_t fn = f + kFinWarps;
        uint32_t next_row = 0;
        if (fn < total) {
            next_row = row_of(fn);
            const uint8_t* np_ = ix.vecs + (size_t)next_row * ix.row_bytes;
            if (lane * 128u < ix.row_bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(np_ + lane * 128u));
        }
        const uint8_t* rp = ix.vecs + (size_t)row * ix.row_bytes;
        const float nb = ix.metric == VELES_COSINE ? *reinterpret_cast<const float*>(rp + ix.norm_off) : 0.0f;
        const float v = ix.dtype == VELES_F32 ? warp_metric(ix.metric, true, qs, reinterpret_cast<const float*>(rp), ix.dim, na, nb, lane)
                                              : warp_metric(ix.metric, true, qs, reinterpret_cast<const __half*>(rp), ix.dim, na, nb, lane);
        const uint32_t ok = ord_key(v);
        const uint64_t key = ((uint64_t)(desc ? ~ok : ok) << 32) | row;
        if (len < k || key < mine[k - 1]) {
            const uint32_t pos = lower_bound_warp(mine, len, key, lane);
            if (len < k) {
                insert_at(mine, pos, len + 1, key, lane);
                ++len;
            } else {
                insert_at(mine, pos, len, key, lane);
            }
        }
        __syncwarp();
        row = next_row;
    }
    __syncthreads();  // s_len is reused below: every warp is past the prefix sum (trivially) and done with its list
    if (lane == 0) s_len[warp] = len;
    __syncthreads();
    if (warp != 0) return;
    for (uint32_t w = 1; w < kFinWarps; ++w) {
        const uint64_t* other = lists + (size_t)w * k;
        for (uint32_t j = 0; j < s_len[w]; ++j) {
            const uint64_t key = other[j];
            if (len == k && key >= mine[k - 1]) break;
            const uint32_t pos = lower_bound_warp(mine, len, key, lane);
            if (len < k) {
                insert_at(mine, pos, len + 1, key, lane);
                ++len;
            } else {
                insert_at(mine, pos, len, key, lane);
            }
            __syncwarp();
        }
    }
    for (uint32_t j = lane; j < k; j += 32) {
        uint32_t id = VELES_INVALID_ID;
        float sc = __uint_as_float(0x7fc00000u);
        if (j < len) {
            id = (uint32_t)mine[j];
            const uint32_t kb = (uint32_t)(mine[j] >> 32);
            sc = ord_unkey(desc ? ~kb : kb);
        }
        out_ids[(size_t)q * k + j] = id;
        out_score[(size_t)q * k + j] = sc;
    }
}

// The finish most calls take (kp <= 128).  A candidate carries its fp16-GEMM score, so the exact re-rank does not have
// to touch every row that passed the (loose, sampled) bound: phase A keeps the `kr` best candidates by GEMM score -- 8
// bytes per candidate, one candidate per lane, register-resident lists merged by warp 0 -- and phase B computes the exact
// metric value of those kr rows only (kr = 2 * kp clamped to 64..128: the k * oversample the caller asked for, doubled
// as a margin for fp16 rank noise).  That is the reference's own shape (search_with_rerank re-ranks rerank_k candidates,
// index/hnsw/index/search.rs:118-160), and it decouples the re-rank traffic (kr rows of dim * 4 bytes per query) from
// the sample size: the sample can be 4x smaller and the candidate lists 4x longer for the same finish cost.
__global__ void __launch_bounds__(kFinWarps * 32) relaxed_finish_select_kernel(IndexView ix, const float* __restrict__ queries, uint32_t nq,
                                                                             const uint64_t* __restrict__ cand,
                                                                             const uint32_t* __restrict__ cand_cnt, uint32_t cand_cap,
                                                                             uint32_t n_seg, uint32_t k, uint32_t kr, int use_select,
                                                                             uint32_t* __restrict__ out_ids, float* __restrict__ out_score) {
    constexpr int R = 4;
    extern __shared__ __align__(16) uint8_t fin_smem[];
    __shared__ uint32_t s_off[kFinMaxSeg + 1];
    __shared__ uint32_t s_len[kFinWarps];
    __shared__ uint64_t s_keys[kFinWarps][32 * R];
    __shared__ SelScratch s_sel;
    __shared__ uint32_t s_pick;
    float* qs = reinterpret_cast<float*>(fin_smem);                                       // dim
    uint64_t* lists = reinterpret_cast<uint64_t*>(fin_smem + ((ix.dim * 4 + 15) & ~15u));  // kFinWarps x k
    const uint32_t q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* qg = queries + (size_t)q * ix.dim;
    for (uint32_t i = threadIdx.x; i < ix.dim; i += blockDim.x) qs[i] = qg[i];
    {
        const uint32_t c = threadIdx.x < n_seg ? min(cand_cnt[(size_t)threadIdx.x * nq + q], cand_cap) : 0u;
        uint32_t x = c;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL_MASK, x, o);
            if ((int)lane >= o) x += y;
        }
        if (lane == 31) s_len[warp] = x;
        __syncthreads();
        uint32_t before = 0;
        for (uint32_t w = 0; w < warp; ++w) before += s_len[w];
        s_off[threadIdx.x + 1] = before + x;
        if (threadIdx.x == 0) s_off[0] = 0;
        __syncthreads();
    }
    const uint32_t total = s_off[kFinWarps * 32];
    // ---- phase A: the kr best candidates by GEMM score (descending score = ascending key; ties by row id) ----
    // key of flattened candidate f: its segment is the last one whose offset is <= f
    auto key_of = [&](uint32_t f) -> uint64_t {
        uint32_t lo = 0, hi = kFinWarps * 32;  // s_off[lo] <= f < s_off[hi]
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_off[mid] <= f)
                lo = mid;
            else
                hi = mid;
        }
        const uint64_t e = cand[((size_t)lo * nq + q) * cand_cap + (f - s_off[lo])];
        return ((uint64_t)(~ord_key(__uint_as_float((uint32_t)(e >> 32)))) << 32) | (uint32_t)e;
    };
    bool picked = false;
    if (use_select) picked = cta_select_smallest(key_of, total, kr, s_sel, &s_keys[0][0], &s_pick);
    if (!picked) {
        __syncthreads();
        RegTopK<R> top;
        top.init(kr, lane);
        constexpr uint32_t U = 4;
        for (uint32_t base = warp * 32 * U; base < total; base += kFinWarps * 32 * U) {
            uint64_t key[U];
#pragma unroll
            for (uint32_t u = 0; u < U; ++u) {
                const uint32_t f = base + u * 32 + lane;
                key[u] = f < total ? key_of(f) : ~0ull;
            }
#pragma unroll
            for (uint32_t u = 0; u < U; ++u) top.offer(key[u]);
        }
        top.store(s_keys[warp], 32 * R);
        __syncthreads();
        if (warp == 0) {
            for (uint32_t w = 1; w < kFinWarps; ++w)
                for (uint32_t j = 0; j < kr; j += 32) {
                    const uint64_t key = j + lane < kr ? s_keys[w][j + lane] : ~0ull;
                    if (__shfl_sync(FULL_MASK, key, 0) >= top.worst) break;  // ascending list: nothing further can enter
                    top.offer(key);
                }
            const uint32_t cnt = top.size();
            __syncwarp();
            top.store(s_keys[0], 32 * R);
            if (lane == 0) s_pick = cnt;
        }
        __syncthreads();
    }
    const uint32_t npick = s_pick;
    // ---- phase B: exact metric value of the picked rows, top-k in sort_results order ----
    const bool desc = ix.metric == VELES_COSINE || ix.metric == VELES_DOT || ix.metric == VELES_JACCARD;
    float na = 0.0f;
    if (ix.metric == VELES_COSINE) na = __fsqrt_rn(warp_tree_reduce<0>(qs, qs, ix.dim, lane));
    uint64_t* mine = lists + (size_t)warp * k;
    uint32_t len = 0;
    for (uint32_t c = warp; c < npick; c += kFinWarps) {
        const uint32_t row = (uint32_t)s_keys[0][c];
        if (c + kFinWarps < npick) {
            const uint8_t* np_ = ix.vecs + (size_t)(uint32_t)s_keys[0][c + kFinWarps] * ix.row_bytes;
            if (lane * 128u < ix.row_bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(np_ + lane * 128u));
        }
        const uint8_t* rp = ix.vecs + (size_t)row * ix.row_bytes;
        const float nb = ix.metric == VELES_COSINE ? *reinterpret_cast<const float*>(rp + ix.norm_off) : 0.0f;
        const float v = ix.dtype == VELES_F32 ? warp_metric(ix.metric, true, qs, reinterpret_cast<const float*>(rp), ix.dim, na, nb, lane)
                                              : warp_metric(ix.metric, true, qs, reinterpret_cast<const __half*>(rp), ix.dim, na, nb, lane);
        const uint32_t ok = ord_key(v);
        const uint64_t key = ((uint64_t)(desc ? ~ok : ok) << 32) | row;
        if (len < k || key < mine[k - 1]) {
            const uint32_t pos = lower_bound_warp(mine, len, key, lane);
            if (len < k) {
                insert_at(mine, pos, len + 1, key, lane);
                ++len;
            } else {
                insert_at(mine, pos, len, key, lane);
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (lane == 0) s_len[warp] = len;
    __syncthreads();
    if (warp != 0) return;
    for (uint32_t w = 1; w < kFinWarps; ++w) {
        const uint64_t* other = lists + (size_t)w * k;
        for (uint32_t j = 0; j < s_len[w]; ++j) {
            const uint64_t key = other[j];
            if (len == k && key >= mine[k - 1]) break;
            const uint32_t pos = lower_bound_warp(mine, len, key, lane);
            if (len < k) {
                insert_at(mine, pos, len + 1, key, lane);
                ++len;
            } else {
                insert_at(mine, pos, len, key, lane);
            }
            __syncwarp();
        }
    }
    for (uint32_t j = lane; j < k; j += 32) {
        uint32_t id = VELES_INVALID_ID;
        float sc = __uint_as_float(0x7fc00000u);
        if (j < len) {
            id = (uint32_t)mine[j];
            const uint32_t kb = (uint32_t)(mine[j] >> 32);
            sc = ord_unkey(desc ? ~kb : kb);
        }
        out_ids[(size_t)q * k + j] = id;
        out_score[(size_t)q * k + j] = sc;
    }
}

// ---- host --------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int32_t make_tensor_map(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t dpad, uint32_t box_rows) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        VELES_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
        if (!p || qr != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return VELES_ERR_UNSUPPORTED;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    const cuuint64_t gdim[2] = {dpad, rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)dpad * 2};
    const cuuint32_t box[2] = {kTcBlockK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return VELES_ERR_CUDA;
    }
    return VELES_OK;
}

static int32_t launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& p, uint32_t bn, cudaStream_t st) {
    const int sms = device_sm_count();
    if (p.n_mtiles == 0 || p.n_nblocks == 0) return VELES_OK;
    const uint32_t grid = std::min<uint32_t>(p.n_mtiles, (uint32_t)sms);
    if (bn == 256) {
        VELES_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<256>::kSmemBytes));
        gemm_tc_kernel<256><<<grid, 256, TcCfg<256>::kSmemBytes, st>>>(ta, tb, p);
    } else {
        VELES_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<128>::kSmemBytes));
        gemm_tc_kernel<128><<<grid, 256, TcCfg<128>::kSmemBytes, st>>>(ta, tb, p);
    }
    count_launch();
    VELES_CUDA(cudaGetLastError());
    return VELES_OK;
}

// fp16 copy of the collection, built once per snapshot
static int32_t ensure_x16(const veles_index* ix, cudaStream_t st) {
    const uint32_t dpad = round_up(ix->dim, kTcBlockK);
    if (ix->x16.p && ix->x16_dpad == dpad) return VELES_OK;
    VELES_TRY(ix->x16.alloc(std::max<size_t>((size_t)ix->n * dpad * 2, 16)));
    VELES_TRY(ix->x16_bias.alloc(std::max<size_t>((size_t)ix->n * 4, 16)));
    if (ix->n) {
        const int sms = device_sm_count();
        to_f16_operand_kernel<<<sms * 8, 256, 0, st>>>(ix->vecs.as<uint8_t>(), ix->row_bytes, ix->dtype == VELES_F16 ? 1 : 0,
                                                       ix->metric == VELES_COSINE ? ix->norm_off : 0xffffffffu,
                                                       ix->metric == VELES_COSINE ? 1 : 0, ix->n, ix->dim, dpad, ix->x16.as<__half>(),
                                                       ix->x16_bias.as<float>());
        count_launch();
        VELES_CUDA(cudaGetLastError());
    }
    ix->x16_dpad = dpad;
    return VELES_OK;
}

int32_t bruteforce_relaxed_d(const veles_index* ix, const float* q_d, uint32_t nq, uint32_t k, uint32_t oversample, uint32_t* ids_d,
                             float* score_d, float* gemm_ms, cudaStream_t st) {
    NvtxRange nvtx_range("veles::bruteforce_relaxed (tcgen05 GEMM + exact re-rank)");
    VELES_REQUIRE(ix->dtype == VELES_F32 || ix->dtype == VELES_F16, "the tensor-core path needs f32 or f16 storage");
    VELES_REQUIRE(ix->metric == VELES_COSINE || ix->metric == VELES_EUCLIDEAN || ix->metric == VELES_DOT,
                  "the tensor-core path handles cosine, euclidean and dot");
    VELES_REQUIRE(k >= 1 && k <= 1024 && oversample >= 1 && (uint64_t)k * oversample <= 4096, "k in 1..1024 and k * oversample <= 4096");
    if (nq == 0) return VELES_OK;
    const uint64_t n = ix->n;
    VELES_REQUIRE(n >= 1, "empty snapshot");
    const int sms = device_sm_count();
    VELES_TRY(ensure_x16(ix, st));
    const uint32_t dpad = ix->x16_dpad;
    // threshold rank: k * oversample, but never below 16 -- the full-collection rank of the sample's j-th best is
    // Gamma(j)-distributed around j * n / sample_rows, and small j would make the 4x candidate capacity overflow
    const uint32_t kp = (uint32_t)std::min<uint64_t>(std::max<uint64_t>((uint64_t)k * oversample, 16), n);
    const uint32_t n_mtiles = (uint32_t)((n + kTcBlockM - 1) / kTcBlockM);
    // sample: one row tile in 16 -- the expected candidates per query are kp * n / sample_rows ~ 16 * kp, spread over the
    // GEMM CTAs' segments of the query's list
    // VELES_TC_OLD_TAIL=1: round 2's first threshold / finish kernels (one warp per query; serial segment walk; every
    // candidate re-ranked), for A/B.  The select finish (kp <= 128) re-ranks the kr best candidates by GEMM score, so
    // its cost does not grow with the candidate lists and the sample shrinks to one row tile in 64.
    const bool old_tail = std::getenv("VELES_TC_OLD_TAIL") != nullptr;
    const bool select_fin = kp <= 128 && !old_tail && sms <= kFinMaxSeg && std::getenv("VELES_TC_RERANK_ALL") == nullptr;
    uint32_t sample_div = select_fin ? 64u : 16u;
    if (const char* e = std::getenv("VELES_TC_SAMPLE_DIV")) sample_div = std::max(1, std::atoi(e));
    const uint32_t kr = std::min<uint32_t>(std::max<uint32_t>(2 * kp, 64), 128);
    const int use_select = std::getenv("VELES_TC_REGTOPK") == nullptr ? 1 : 0;  // 0: the RegTopK scans only (A/B, and their test)
    const uint32_t s_tiles = std::min<uint32_t>(n_mtiles, std::max<uint32_t>(std::max<uint32_t>(kTcSampleTiles, n_mtiles / sample_div), kp / kTcBlockM + 2));
    const uint32_t s_rows = s_tiles * kTcBlockM;
    const uint32_t n_seg = std::min<uint32_t>(n_mtiles, (uint32_t)sms);  // = the filter pass's grid: one segment per CTA
    const uint64_t expect = (uint64_t)kp * (n / s_rows + 1);  // per query, all segments
    // tiny collections can have fewer than kp sampled rows (threshold = -inf: every row is a candidate): a segment must
    // then hold every row of its CTA's tiles
    const uint64_t all_rows = n <= 8192 ? (uint64_t)kTcBlockM * ((n_mtiles + n_seg - 1) / std::max(n_seg, 1u)) : 0;
    const uint32_t cand_cap = (uint32_t)std::min<uint64_t>(n, std::max<uint64_t>(std::max<uint64_t>(64, all_rows), 4 * expect / std::max(n_seg, 1u) + 32));
    const uint32_t chunk = std::min<uint32_t>(nq, 1024);
    // work buffers live with the snapshot (cudaMalloc / cudaFree per call cost more than the GEMM)
    DevBuf &q16 = ix->tc_q16, &tiles_d = ix->tc_tiles, &sample = ix->tc_sample, &thr = ix->tc_thr, &cnt = ix->tc_cnt,
           &cand = ix->tc_cand, &err = ix->tc_err;
    VELES_TRY(q16.ensure((size_t)round_up(chunk, 256) * dpad * 2));
    VELES_TRY(tiles_d.ensure((size_t)s_tiles * 4));
    VELES_TRY(sample.ensure((size_t)chunk * s_rows * 4));
    VELES_TRY(thr.ensure((size_t)chunk * 4));
    VELES_TRY(cnt.ensure((size_t)sms * chunk * 4));
    VELES_TRY(cand.ensure((size_t)sms * chunk * cand_cap * 8));
    VELES_TRY(err.ensure(16));
    VELES_CUDA(cudaMemsetAsync(err.p, 0, 16, st));
    if (ix->tc_tiles_key[0] != s_tiles || ix->tc_tiles_key[1] != n_mtiles) {
        // the sample's row tiles, spread evenly over the collection: uploaded once per (snapshot, k * oversample) -- the
        // copy from a local needs a stream synchronisation, which every later call is spared
        std::vector<uint32_t> tl(s_tiles);
        for (uint32_t i = 0; i < s_tiles; ++i) tl[i] = (uint32_t)((uint64_t)i * n_mtiles / s_tiles);
        VELES_CUDA(cudaMemcpyAsync(tiles_d.p, tl.data(), (size_t)s_tiles * 4, cudaMemcpyHostToDevice, st));
        VELES_CUDA(cudaStreamSynchronize(st));  // tl is a local
        ix->tc_tiles_key[0] = s_tiles;
        ix->tc_tiles_key[1] = n_mtiles;
    }
    CUtensorMap ta, tb;
    VELES_TRY(make_tensor_map(&ta, ix->x16.p, n, dpad, kTcBlockM));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (gemm_ms) {
        *gemm_ms = 0.0f;
        VELES_CUDA(cudaEventCreate(&e0));
        VELES_CUDA(cudaEventCreate(&e1));
    }
    const IndexView v = ix->view();
    for (uint32_t q0 = 0; q0 < nq; q0 += chunk) {
        const uint32_t nn = std::min(chunk, nq - q0);
        const uint32_t bn = nn > 128 ? 256 : 128;
        VELES_CUDA(cudaMemsetAsync(q16.p, 0, (size_t)round_up(chunk, 256) * dpad * 2, st));
        to_f16_operand_kernel<<<std::min<uint32_t>((nn + 7) / 8, (uint32_t)sms * 8), 256, 0, st>>>(
            reinterpret_cast<const uint8_t*>(q_d + (size_t)q0 * ix->dim), (uint64_t)ix->dim * 4, 0, 0xffffffffu,
            ix->metric == VELES_COSINE ? 1 : 0, nn, ix->dim, dpad, q16.as<__half>(), nullptr);
        count_launch();
        VELES_TRY(make_tensor_map(&tb, q16.p, round_up(nn, bn), dpad, bn));
        TcParams p{};
        p.n_rows = (uint32_t)n;
        p.nq = nn;
        p.k_blocks = dpad / kTcBlockK;
        p.n_nblocks = (nn + bn - 1) / bn;
        p.row_bias = ix->metric == VELES_EUCLIDEAN ? ix->x16_bias.as<float>() : nullptr;
        p.scale = ix->metric == VELES_EUCLIDEAN ? 2.0f : 1.0f;
        p.err = err.as<uint32_t>();
        // 2. sample pass -> thresholds
        p.mode = 0;
        p.n_mtiles = s_tiles;
        p.tile_list = tiles_d.as<uint32_t>();
        p.out = sample.as<float>();
        p.ld_out = s_rows;
        if (gemm_ms) VELES_CUDA(cudaEventRecord(e0, st));
        VELES_TRY(launch_gemm(ta, tb, p, bn, st));
        if (kp <= 128 && !old_tail) {
            if (kp <= 64)
                tc_threshold_cta_kernel<2><<<nn, 256, 0, st>>>(sample.as<float>(), s_rows, s_rows, kp, thr.as<float>(), use_select);
            else
                tc_threshold_cta_kernel<4><<<nn, 256, 0, st>>>(sample.as<float>(), s_rows, s_rows, kp, thr.as<float>(), use_select);
        } else {
            const uint32_t tw = kp <= 256 ? 8 : (kp <= 1024 ? 2 : 1);
            tc_threshold_kernel<<<(nn + tw - 1) / tw, tw * 32, (size_t)tw * kp * 8, st>>>(sample.as<float>(), s_rows, s_rows, nn, kp,
                                                                                      thr.as<float>());
        }
        count_launch();
        VELES_CUDA(cudaGetLastError());
        // 3. filter pass over every tile
        VELES_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)sms * nn * 4, st));  // CTAs that never launch leave empty segments
        p.mode = 1;
        p.n_mtiles = n_mtiles;
        p.tile_list = nullptr;
        p.thr = thr.as<float>();
        p.cand_cnt = cnt.as<uint32_t>();
        p.cand = cand.as<uint64_t>();
        p.cand_cap = cand_cap;
        VELES_TRY(launch_gemm(ta, tb, p, bn, st));
        if (gemm_ms) {
            VELES_CUDA(cudaEventRecord(e1, st));
        }
        // 4. exact re-rank of the candidates, top-k
        const size_t fsm = ((ix->dim * 4 + 15) & ~15u) + (size_t)kFinWarps * k * 8;
        if (select_fin) {
            VELES_CUDA(cudaFuncSetAttribute(relaxed_finish_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
            relaxed_finish_select_kernel<<<nn, kFinWarps * 32, fsm, st>>>(v, q_d + (size_t)q0 * ix->dim, nn, cand.as<uint64_t>(),
                                                                       cnt.as<uint32_t>(), cand_cap, (uint32_t)sms, k, kr, use_select,
                                                                       ids_d + (size_t)q0 * k, score_d + (size_t)q0 * k);
        } else {
            auto fin = (sms <= kFinMaxSeg && !old_tail) ? relaxed_finish_flat_kernel : relaxed_finish_kernel;
            VELES_CUDA(cudaFuncSetAttribute(fin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
            fin<<<nn, kFinWarps * 32, fsm, st>>>(v, q_d + (size_t)q0 * ix->dim, nn, cand.as<uint64_t>(), cnt.as<uint32_t>(), cand_cap,
                                              (uint32_t)sms, k, ids_d + (size_t)q0 * k, score_d + (size_t)q0 * k);
        }
        count_launch();
        VELES_CUDA(cudaGetLastError());
        if (gemm_ms) {
            VELES_CUDA(cudaEventSynchronize(e1));
            float ms = 0.0f;
            VELES_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            *gemm_ms += ms;
        }
    }
    uint32_t h[4] = {0, 0, 0, 0};
    VELES_CUDA(cudaMemcpyAsync(h, err.p, 16, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaStreamSynchronize(st));
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (h[0]) {
        set_error(h[0] == 2 ? "tensor-core GEMM: shared memory is not 1024-byte aligned" : "tensor-core GEMM: pipeline wait timed out");
        return VELES_ERR_CUDA;
    }
    if (h[1]) {
        set_error("tensor-core GEMM: more than %u candidates passed a query's threshold; use the exact path", cand_cap);
        return VELES_ERR_OVERFLOW;
    }
    return VELES_OK;
}

}  // namespace veles

using namespace veles;

extern "C" {

int32_t veles_bruteforce_batch_relaxed_d(const veles_index_t* idx, const float* queries_d, uint32_t nq, uint32_t k, uint32_t oversample,
                                         uint32_t* out_ids_d, float* out_score_d, float* gemm_ms, void* stream) {
    VELES_REQUIRE(idx != nullptr, "index is NULL");
    VELES_REQUIRE(nq == 0 || (queries_d && out_ids_d && out_score_d), "NULL buffer");
    std::lock_guard<std::mutex> g(idx->mu);
    return bruteforce_relaxed_d(idx, queries_d, nq, k, oversample, out_ids_d, out_score_d, gemm_ms, (cudaStream_t)stream);
}

int32_t veles_bruteforce_batch_relaxed(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k, uint32_t oversample,
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
    VELES_TRY(bruteforce_relaxed_d(idx, idx->q_d.as<float>(), nq, k, oversample, idx->out_ids_d.as<uint32_t>(), idx->out_val_d.as<float>(),
                                   nullptr, st));
    VELES_CUDA(cudaMemcpyAsync(out_ids, idx->out_ids_d.p, ob, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaMemcpyAsync(out_score, idx->out_val_d.p, ob, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaStreamSynchronize(st));
    return VELES_OK;
}

}  // extern "C"

// common.cuh -- shared host/device helpers of libveles_b200 (sm_100a only).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/veles_b200.h"

namespace veles {

// ---------------------------------------------------------------------------------------------
// host side: errors, launch accounting
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Resident CTAs per SM of `kernel` at (threads, dynamic shared memory), with the kernel's dynamic shared-memory limit
// raised to `smem`; queried once per distinct triple (the two runtime calls cost ~10 us, which is most of a
// single-query brute-force scan over a small collection).  < 1 on error (message set).
int cached_blocks_per_sm(const void* kernel, int threads, size_t smem);

// NVTX range around each kernel group, named after the reference function it stands for: the tracing counterpart of
// the reference's spans (core/metrics.rs:977-1058).  Header-only NVTX 3: a no-op unless a profiler is attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

#define VELES_CUDA(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            veles::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (_e == cudaErrorMemoryAllocation) ? VELES_ERR_OOM : VELES_ERR_CUDA;            \
        }                                                                                         \
    } while (0)

#define VELES_REQUIRE(cond, ...)          \
    do {                                  \
        if (!(cond)) {                    \
            veles::set_error(__VA_ARGS__); \
            return VELES_ERR_INVALID;     \
        }                                 \
    } while (0)

#define VELES_TRY(expr)          \
    do {                         \
        int32_t _s = (expr);     \
        if (_s != VELES_OK) return _s; \
    } while (0)

// RAII device buffer
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    int32_t alloc(size_t n) {
        release();
        if (n == 0) n = 16;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu bytes) failed: %s", n, cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? VELES_ERR_OOM : VELES_ERR_CUDA;
        }
        bytes = n;
        return VELES_OK;
    }
    int32_t ensure(size_t n) { return (n <= bytes && p) ? VELES_OK : alloc(n); }
    template <typename T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

// ---------------------------------------------------------------------------------------------
// device view of an index snapshot (layout in HBM)
// ---------------------------------------------------------------------------------------------
struct IndexView {
    const uint8_t* vecs;       // n rows of row_bytes: [dim elements][pad to 4][norm f32 (cosine only)][pad to 16]
    const uint32_t* adj0;      // n rows of stride0 ids, padded with VELES_INVALID_ID
    const uint32_t* upper_ref; // n: (first_row << 4) | top_layer, VELES_INVALID_ID if no upper rows
    const uint32_t* upper_adj; // rows of strideU ids for layers 1..top_layer of each upper node
    uint64_t n;
    uint32_t dim;        // elements (bits for BIN1)
    uint32_t row_bytes;  // bytes per stored row (multiple of 16)
    uint32_t norm_off;   // byte offset of the row's sqrt(|v|^2) (reference accumulation tree); cosine only
    uint32_t stride0;
    uint32_t strideU;
    uint32_t entry;
    uint32_t max_layer;
    int32_t metric;
    int32_t dtype;
    int32_t has_entry;
    const float* sq_min;    // VELES_SQ8 views: the quantizer, to code the query (quantization.rs:236-250)
    const float* sq_scale;
};

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

#define FULL_MASK 0xffffffffu

// f32::total_cmp as an unsigned sortable key (native/ordered_float.rs:31-36)
__device__ __forceinline__ uint32_t ord_key(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord_unkey(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

// The reference's 4 x f32x8 accumulator tree (simd_avx512.rs:150-204) mapped onto a warp:
// element i of the 32-wide main loop lives in lane i % 32 = a*8 + j (accumulator a, SIMD lane j).
//   (P0+P1) and (P2+P3)      -> xor 8     (a ^ 1)
//   (P0+P1)+(P2+P3)          -> xor 16    (a ^ 2)
//   wide::reduce_add          -> xor 4, 2, 1  == ((l0+l4)+(l2+l6)) + ((l1+l5)+(l3+l7))
// Every step adds a commutative pair, so all 32 lanes end with the same bits as the CPU.
__device__ __forceinline__ float warp_tree_sum32(float v) {
    v = __fadd_rn(v, __shfl_xor_sync(FULL_MASK, v, 8));
    v = __fadd_rn(v, __shfl_xor_sync(FULL_MASK, v, 16));
    v = __fadd_rn(v, __shfl_xor_sync(FULL_MASK, v, 4));
    v = __fadd_rn(v, __shfl_xor_sync(FULL_MASK, v, 2));
    v = __fadd_rn(v, __shfl_xor_sync(FULL_MASK, v, 1));
    return v;
}
// wide::f32x8::reduce_add over lanes 0..7 of each 8-lane group
__device__ __forceinline__ float warp_tree_sum8(float v) {
    v = __fadd_rn(v, __shfl_xor_sync(FULL_MASK, v, 4));
    v = __fadd_rn(v, __shfl_xor_sync(FULL_MASK, v, 2));
    v = __fadd_rn(v, __shfl_xor_sync(FULL_MASK, v, 1));
    return v;
}

__device__ __forceinline__ float load_elem(const float* p, uint32_t i) { return p[i]; }
__device__ __forceinline__ float load_elem(const __half* p, uint32_t i) { return __half2float(p[i]); }

// OP 0: sum a[i]*b[i]; OP 1: sum (a[i]-b[i])^2.  `a` is the query (f32), `b` the stored row.
// Follows simd_avx512.rs:150-264 for dim >= 16 and simd_explicit.rs:50-130 below that, including
// the 8-wide and scalar tails.  All lanes return the same value.
template <int OP, typename TA, typename TB>
__device__ __forceinline__ float warp_tree_reduce(const TA* __restrict__ a, const TB* __restrict__ b,
                                                  uint32_t dim, uint32_t lane) {
    float result;
    uint32_t pos;
    if (dim >= 16) {
        const uint32_t main_len = dim & ~31u;
        float acc = 0.0f;
        for (uint32_t i = lane; i < main_len; i += 32) {
            float x = load_elem(a, i), y = load_elem(b, i);
            if (OP == 0) {
                acc = __fmaf_rn(x, y, acc);
            } else {
                float d = __fsub_rn(x, y);
                acc = __fmaf_rn(d, d, acc);
            }
        }
        result = warp_tree_sum32(acc);
        pos = main_len;
        while (pos + 8 <= dim) {
            float t = 0.0f;
            if (lane < 8) {
                float x = load_elem(a, pos + lane), y = load_elem(b, pos + lane);
                if (OP == 0) {
                    t = __fmaf_rn(x, y, 0.0f);
                } else {
                    float d = __fsub_rn(x, y);
                    t = __fmaf_rn(d, d, 0.0f);
                }
            }
            t = warp_tree_sum8(t);
            t = __shfl_sync(FULL_MASK, t, 0);
            result = __fadd_rn(result, t);
            pos += 8;
        }
    } else {
        const uint32_t main_len = dim & ~7u;
        float acc = 0.0f;
        if (lane < 8) {
            for (uint32_t i = lane; i < main_len; i += 8) {
                float x = load_elem(a, i), y = load_elem(b, i);
                if (OP == 0) {
                    acc = __fmaf_rn(x, y, acc);
                } else {
                    float d = __fsub_rn(x, y);
                    acc = __fmaf_rn(d, d, acc);
                }
            }
        }
        acc = warp_tree_sum8(acc);
        result = __shfl_sync(FULL_MASK, acc, 0);
        pos = main_len;
    }
    for (; pos < dim; ++pos) {  // scalar tail: mul then add, unfused (simd_avx512.rs:198-201)
        float x = load_elem(a, pos), y = load_elem(b, pos);
        if (OP == 0) {
            result = __fadd_rn(result, __fmul_rn(x, y));
        } else {
            float d = __fsub_rn(x, y);
            result = __fadd_rn(result, __fmul_rn(d, d));
        }
    }
    return result;
}

// counts for Hamming / Jaccard on f32 (or f16) lanes thresholded at > 0.5
// (simd_explicit.rs:256-287, 372-443).  Integer sums: order free.
template <typename TB>
__device__ __forceinline__ void warp_threshold_counts(const float* __restrict__ a, const TB* __restrict__ b,
                                                      uint32_t dim, uint32_t lane, uint32_t& diff, uint32_t& inter,
                                                      uint32_t& uni) {
    uint32_t d = 0, in = 0, un = 0;
    for (uint32_t i = lane; i < dim; i += 32) {
        bool x = a[i] > 0.5f, y = load_elem(b, i) > 0.5f;
        d += (x != y);
        in += (x && y);
        un += (x || y);
    }
    diff = __reduce_add_sync(FULL_MASK, d);
    inter = __reduce_add_sync(FULL_MASK, in);
    uni = __reduce_add_sync(FULL_MASK, un);
}

// cosine similarity from the three tree sums (simd_avx512.rs:344-351)
__device__ __forceinline__ float cosine_from_parts(float dot, float norm_a, float norm_b) {
    if (norm_a == 0.0f || norm_b == 0.0f) return 0.0f;
    return __fdiv_rn(dot, __fmul_rn(norm_a, norm_b));
}

// In-graph distance (SimdDistance::distance, native/distance.rs:75-85) or metric value
// (HnswIndex::compute_distance, index/hnsw/index/search.rs:30-38) of query `a` against stored row
// `b`.  norm_a / norm_b are sqrt of the tree sums |a|^2, |b|^2 (hoisted: same inputs, same tree,
// same bits as recomputing them per pair).
template <typename TB>
__device__ __forceinline__ float warp_metric(int metric, bool as_value, const float* __restrict__ a,
                                             const TB* __restrict__ b, uint32_t dim, float norm_a, float norm_b,
                                             uint32_t lane) {
    switch (metric) {
        case VELES_COSINE: {
            float dot = warp_tree_reduce<0>(a, b, dim, lane);
            float sim = cosine_from_parts(dot, norm_a, norm_b);
            return as_value ? sim : __fsub_rn(1.0f, sim);
        }
        case VELES_EUCLIDEAN: {
            float s = warp_tree_reduce<1>(a, b, dim, lane);
            return __fsqrt_rn(s);
        }
        case VELES_DOT: {
            float dot = warp_tree_reduce<0>(a, b, dim, lane);
            return as_value ? dot : -dot;
        }
        case VELES_HAMMING: {
            uint32_t d, in, un;
            warp_threshold_counts(a, b, dim, lane, d, in, un);
            return (float)d;
        }
        default: {  // JACCARD
            uint32_t d, in, un;
            warp_threshold_counts(a, b, dim, lane, d, in, un);
            float j = (un == 0) ? 1.0f : __fdiv_rn((float)in, (float)un);
            return as_value ? j : __fsub_rn(1.0f, j);
        }
    }
}


// ---- four rows per warp ("quad"): lane = 8*g + t, group g owns one row, sub-lane t owns elements
// 32*it + 4t .. 4t+3 of it, i.e. the reference accumulators P[4t+e] = P[a][j] with a = t/2,
// j = 4*(t%2) + e.  Combination order of simd_avx512.rs:182-184 + wide::reduce_add:
//   xor 2 (a^1), xor 4 (a^2)  ->  C[j] = (P0+P1)+(P2+P3)
//   xor 1                     ->  q[e] = C[e] + C[e+4]
//   in-thread                 ->  (q0+q2) + (q1+q3)
// Each step adds a commutative pair, so the result has the CPU's bits.
__device__ __forceinline__ float quad_tree_sum(float a0, float a1, float a2, float a3) {
    a0 = __fadd_rn(a0, __shfl_xor_sync(FULL_MASK, a0, 2));
    a1 = __fadd_rn(a1, __shfl_xor_sync(FULL_MASK, a1, 2));
    a2 = __fadd_rn(a2, __shfl_xor_sync(FULL_MASK, a2, 2));
    a3 = __fadd_rn(a3, __shfl_xor_sync(FULL_MASK, a3, 2));
    a0 = __fadd_rn(a0, __shfl_xor_sync(FULL_MASK, a0, 4));
    a1 = __fadd_rn(a1, __shfl_xor_sync(FULL_MASK, a1, 4));
    a2 = __fadd_rn(a2, __shfl_xor_sync(FULL_MASK, a2, 4));
    a3 = __fadd_rn(a3, __shfl_xor_sync(FULL_MASK, a3, 4));
    a0 = __fadd_rn(a0, __shfl_xor_sync(FULL_MASK, a0, 1));
    a1 = __fadd_rn(a1, __shfl_xor_sync(FULL_MASK, a1, 1));
    a2 = __fadd_rn(a2, __shfl_xor_sync(FULL_MASK, a2, 1));
    a3 = __fadd_rn(a3, __shfl_xor_sync(FULL_MASK, a3, 1));
    return __fadd_rn(__fadd_rn(a0, a2), __fadd_rn(a1, a3));
}

__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __half* p) {
    const uint2 raw = *reinterpret_cast<const uint2*>(p);
    const __half2 lo = *reinterpret_cast<const __half2*>(&raw.x), hi = *reinterpret_cast<const __half2*>(&raw.y);
    const float2 a = __half22float2(lo), b = __half22float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
}

// ---- warp-cooperative sorted array of unique u64 keys (ascending) in shared memory ----
// position of the first key in res[0..len) that is >= key
__device__ __forceinline__ uint32_t lower_bound_warp(const uint64_t* res, uint32_t len, uint64_t key, uint32_t lane) {
    uint32_t p = 0;
    if (len > 128) {
        // coarse step: lane l looks at the last element of block l
        const uint32_t B = (len + 31) >> 5;
        uint32_t last = (lane + 1) * B;
        last = (last < len ? last : len);
        const bool lt = lane * B < len && res[last - 1] < key;
        const uint32_t nb = __popc(__ballot_sync(FULL_MASK, lt));
        uint32_t lo = nb * B, hi = lo + B;
        hi = hi < len ? hi : len;
        p = lo < len ? lo : len;
        for (uint32_t base = lo; base < hi; base += 32) {
            const uint32_t i = base + lane;
            const bool l2 = i < hi && res[i] < key;
            const uint32_t msk = __ballot_sync(FULL_MASK, l2);
            p += __popc(msk);
            if (msk != FULL_MASK) break;
        }
        return p;
    }
    for (uint32_t base = 0; base < len; base += 32) {
        const uint32_t i = base + lane;
        const bool lt = i < len && res[i] < key;
        const uint32_t msk = __ballot_sync(FULL_MASK, lt);
        p += __popc(msk);
        if (msk != FULL_MASK) break;
    }
    return p;
}

// Inserts `key` at position pos, shifting res[pos..new_len-1) up by one (the old last element
// falls off when the array is full).  Chunks move from the top so nothing unread is overwritten.
__device__ __forceinline__ void insert_at(uint64_t* res, uint32_t pos, uint32_t new_len, uint64_t key, uint32_t lane) {
    __syncwarp();  // the lanes' reads of the search that found `pos` come before any write below
    int32_t top = (int32_t)new_len - 1;  // exclusive end of the source range
    while (top > (int32_t)pos) {
        const int32_t i = top - 1 - (int32_t)lane;
        const bool mv = i >= (int32_t)pos;
        uint64_t v = 0;
        if (mv) v = res[i];
        __syncwarp();
        if (mv) res[i + 1] = v;
        __syncwarp();
        top -= 32;
    }
    if (lane == 0) res[pos] = key;
    __syncwarp();
}

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint64_t w = __shfl_xor_sync(FULL_MASK, v, o);
        v = w < v ? w : v;
    }
    return v;
}

// ---- register-resident sorted list of the k smallest unique u64 keys (k <= 32 * R) ----
// Position i lives in lane i % 32, slot i / 32; empty positions hold ~0.  An insert is one ballot per slot
// (lower bound) plus one shuffle-up per slot, no shared memory and no barriers.
template <int R>
struct RegTopK {
    uint64_t k[R];
    uint64_t worst;  // key at position cap-1 once full, else ~0
    uint32_t cap, lane;
    __device__ __forceinline__ void init(uint32_t cap_, uint32_t lane_) {
        cap = cap_;
        lane = lane_;
        worst = ~0ull;
#pragma unroll
        for (int s = 0; s < R; ++s) k[s] = ~0ull;
    }
    // all lanes pass the same key
    __device__ __forceinline__ void insert(uint64_t key) {
        uint32_t pos = 0;
#pragma unroll
        for (int s = 0; s < R; ++s) pos += __popc(__ballot_sync(FULL_MASK, k[s] < key));
#pragma unroll
        for (int s = R - 1; s >= 0; --s) {
            uint64_t up = __shfl_up_sync(FULL_MASK, k[s], 1);
            if (s > 0) {
                const uint64_t carry = __shfl_sync(FULL_MASK, k[s - 1], 31);
                up = lane == 0 ? carry : up;
            }
            const uint32_t i = 32u * s + lane;
            uint64_t nv = i > pos ? up : k[s];
            nv = i == pos ? key : nv;
            nv = i >= cap ? ~0ull : nv;
            k[s] = nv;
        }
        // key at position cap - 1
        uint64_t w = ~0ull;
#pragma unroll
        for (int s = 0; s < R; ++s) {
            const uint64_t t = __shfl_sync(FULL_MASK, k[s], (cap - 1) & 31);
            w = ((int)((cap - 1) >> 5) == s) ? t : w;
        }
        worst = w;
    }
    // each lane offers its own key (~0 = nothing); keys >= worst are dropped
    __device__ __forceinline__ void offer(uint64_t key) {
        uint32_t msk = __ballot_sync(FULL_MASK, key < worst);
        while (msk) {
            const uint32_t src = __ffs(msk) - 1;
            msk &= msk - 1;
            const uint64_t kk = __shfl_sync(FULL_MASK, key, src);
            if (kk < worst) insert(kk);
        }
    }
    __device__ __forceinline__ uint32_t size() const {
        uint32_t n = 0;
#pragma unroll
        for (int s = 0; s < R; ++s) n += __popc(__ballot_sync(FULL_MASK, k[s] != ~0ull));
        return n;
    }
    // writes positions [0, count) to out (count <= cap); positions past size() get ~0
    __device__ __forceinline__ void store(uint64_t* out, uint32_t count) const {
#pragma unroll
        for (int s = 0; s < R; ++s) {
            const uint32_t i = 32u * s + lane;
            if (i < count) out[i] = k[s];
        }
    }
};

// the 8-wide and scalar tails of simd_avx512.rs:186-201 applied to a main-loop result (dim >= 16)
template <int OP, typename TA, typename TB>
__device__ __forceinline__ float warp_tree_tail(float result, const TA* __restrict__ a, const TB* __restrict__ b,
                                                uint32_t dim, uint32_t lane) {
    uint32_t pos = dim & ~31u;
    while (pos + 8 <= dim) {
        float t = 0.0f;
        if (lane < 8) {
            float x = load_elem(a, pos + lane), y = load_elem(b, pos + lane);
            if (OP == 0) {
                t = __fmaf_rn(x, y, 0.0f);
            } else {
                float d = __fsub_rn(x, y);
                t = __fmaf_rn(d, d, 0.0f);
            }
        }
        t = warp_tree_sum8(t);
        t = __shfl_sync(FULL_MASK, t, 0);
        result = __fadd_rn(result, t);
        pos += 8;
    }
    for (; pos < dim; ++pos) {
        float x = load_elem(a, pos), y = load_elem(b, pos);
        if (OP == 0) {
            result = __fadd_rn(result, __fmul_rn(x, y));
        } else {
            float d = __fsub_rn(x, y);
            result = __fadd_rn(result, __fmul_rn(d, d));
        }
    }
    return result;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, UBLKCP in SASS) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy, completion counted on `bar`; bytes % 16 == 0, both 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 16-byte per-thread async copy global -> shared (LDGSTS), grouped per thread
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ uint64_t make_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

#endif  // __CUDACC__

}  // namespace veles

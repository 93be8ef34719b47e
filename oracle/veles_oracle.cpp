// veles_oracle.cpp -- CPU restatement of VelesDB's vector-search hot path.
//
// *** TEST INFRASTRUCTURE, NOT PRODUCT CODE. ***
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product (velesdb_b200/) never
// links, imports or falls back to it.
//
// What it restates (paths relative to /root/reference/crates/velesdb-core/src):
//   * distance kernels      simd_avx512.rs:87-352, simd_explicit.rs:50-443,
//                           index/hnsw/native/distance.rs:62-107
//   * HNSW graph            index/hnsw/native/graph.rs:158-640, layer.rs,
//                           ordered_float.rs:13-37
//   * file format v1        index/hnsw/native/backend_adapter.rs:184-380
//   * score transform       index/hnsw/native/backend_adapter.rs:160-168
//   * ef rules              index/hnsw/params.rs:309-319
//   * brute force           index/hnsw/index/search.rs:30-38,176-219,
//                           distance.rs:76-103
//   * BM25                  index/bm25.rs:114-120,269-376
//   * hybrid RRF            collection/search/text.rs:113-180, search/mod.rs:24-42
//   * fusion strategies     fusion/strategy.rs:138-300
//
// Third-party arithmetic that is NOT under /root/reference and is restated from
// its published behaviour:
//   * wide 0.7.33 f32x8::mul_add  = fused multiply-add when built with the fma
//     target feature (the reference builds with -C target-cpu=native,
//     .cargo/config.toml:27-51), else mul then add.  `fma_mode` selects.
//   * wide 0.7.33 f32x8::reduce_add (AVX path): add high and low 128-bit halves,
//     then movehl add, then lane-1 shuffle add, i.e.
//     ((l0+l4)+(l2+l6)) + ((l1+l5)+(l3+l7)).
//   * Rust std::collections::BinaryHeap (push = sift_up from the end, pop =
//     swap_remove root + sift_down_to_bottom + sift_up, into_iter = backing
//     vector order) -- needed because graph.rs:516-518 exposes the heap's
//     internal order among equal distances.
//
// Parity status: pinned against every known-answer test the reference holds for
// this path (tests/test_oracle_known_answers.py lists them with file:line).
// Exact neighbour-id lists are NOT pinned by any reference test (SURVEY.md
// section 8c); for those the oracle itself is the definition.
//
// Build: see oracle/Makefile  (g++ -O3 -march=x86-64-v3 -ffp-contract=off).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define VO_HAVE_AVX2 1
#else
#define VO_HAVE_AVX2 0
#endif

namespace vo {

enum Metric { COSINE = 0, EUCLIDEAN = 1, DOT = 2, HAMMING = 3, JACCARD = 4 };

// ---------------------------------------------------------------------------
// f32x8 emulation (wide 0.7.33)
// ---------------------------------------------------------------------------
static inline float madd(float a, float b, float c, bool fma) {
    if (fma) return std::fmaf(a, b, c);
    float p = a * b;  // -ffp-contract=off keeps this unfused
    return p + c;
}

// wide::f32x8::reduce_add, AVX path.
static inline float hsum8(const float* l) {
    float q0 = l[0] + l[4], q1 = l[1] + l[5], q2 = l[2] + l[6], q3 = l[3] + l[7];
    float d0 = q0 + q2, d1 = q1 + q3;
    return d0 + d1;
}

// 4 x f32x8 accumulator tree of simd_avx512.rs:150-204 for one quantity.
// op: 0 = a*b, 1 = (a-b)^2.
template <int OP>
static float wide32_scalar(const float* a, const float* b, size_t len, bool fma) {
    float P[32];
    for (int i = 0; i < 32; ++i) P[i] = 0.0f;
    size_t simd_len = len / 32;
    for (size_t it = 0; it < simd_len; ++it) {
        const float* pa = a + it * 32;
        const float* pb = b + it * 32;
        for (int i = 0; i < 32; ++i) {
            if (OP == 0) {
                P[i] = madd(pa[i], pb[i], P[i], fma);
            } else {
                float d = pa[i] - pb[i];
                P[i] = madd(d, d, P[i], fma);
            }
        }
    }
    float C[8];
    for (int j = 0; j < 8; ++j) C[j] = (P[j] + P[8 + j]) + (P[16 + j] + P[24 + j]);
    float result = hsum8(C);
    size_t pos = simd_len * 32;
    while (pos + 8 <= len) {
        float t[8];
        for (int j = 0; j < 8; ++j) {
            if (OP == 0) {
                t[j] = madd(a[pos + j], b[pos + j], 0.0f, fma);
            } else {
                float d = a[pos + j] - b[pos + j];
                t[j] = madd(d, d, 0.0f, fma);
            }
        }
        result += hsum8(t);
        pos += 8;
    }
    while (pos < len) {
        if (OP == 0) {
            float p = a[pos] * b[pos];
            result += p;
        } else {
            float d = a[pos] - b[pos];
            float p = d * d;
            result += p;
        }
        ++pos;
    }
    return result;
}

#if VO_HAVE_AVX2
static inline float hsum8_avx(__m256 v) {
    __m128 hi = _mm256_extractf128_ps(v, 1);
    __m128 lo = _mm256_castps256_ps128(v);
    __m128 sq = _mm_add_ps(lo, hi);
    __m128 hd = _mm_movehl_ps(sq, sq);
    __m128 sd = _mm_add_ps(sq, hd);
    __m128 h1 = _mm_shuffle_ps(sd, sd, 0x1);
    __m128 s = _mm_add_ss(sd, h1);
    return _mm_cvtss_f32(s);
}

template <int OP>
static float wide32_avx(const float* a, const float* b, size_t len) {
    __m256 s0 = _mm256_setzero_ps(), s1 = s0, s2 = s0, s3 = s0;
    size_t simd_len = len / 32;
    for (size_t it = 0; it < simd_len; ++it) {
        const float* pa = a + it * 32;
        const float* pb = b + it * 32;
        __m256 a0 = _mm256_loadu_ps(pa), b0 = _mm256_loadu_ps(pb);
        __m256 a1 = _mm256_loadu_ps(pa + 8), b1 = _mm256_loadu_ps(pb + 8);
        __m256 a2 = _mm256_loadu_ps(pa + 16), b2 = _mm256_loadu_ps(pb + 16);
        __m256 a3 = _mm256_loadu_ps(pa + 24), b3 = _mm256_loadu_ps(pb + 24);
        if (OP == 0) {
            s0 = _mm256_fmadd_ps(a0, b0, s0);
            s1 = _mm256_fmadd_ps(a1, b1, s1);
            s2 = _mm256_fmadd_ps(a2, b2, s2);
            s3 = _mm256_fmadd_ps(a3, b3, s3);
        } else {
            __m256 d0 = _mm256_sub_ps(a0, b0), d1 = _mm256_sub_ps(a1, b1);
            __m256 d2 = _mm256_sub_ps(a2, b2), d3 = _mm256_sub_ps(a3, b3);
            s0 = _mm256_fmadd_ps(d0, d0, s0);
            s1 = _mm256_fmadd_ps(d1, d1, s1);
            s2 = _mm256_fmadd_ps(d2, d2, s2);
            s3 = _mm256_fmadd_ps(d3, d3, s3);
        }
    }
    __m256 c = _mm256_add_ps(_mm256_add_ps(s0, s1), _mm256_add_ps(s2, s3));
    float result = hsum8_avx(c);
    size_t pos = simd_len * 32;
    while (pos + 8 <= len) {
        __m256 va = _mm256_loadu_ps(a + pos), vb = _mm256_loadu_ps(b + pos);
        __m256 t;
        if (OP == 0) {
            t = _mm256_fmadd_ps(va, vb, _mm256_setzero_ps());
        } else {
            __m256 d = _mm256_sub_ps(va, vb);
            t = _mm256_fmadd_ps(d, d, _mm256_setzero_ps());
        }
        result += hsum8_avx(t);
        pos += 8;
    }
    while (pos < len) {
        if (OP == 0) {
            float p = a[pos] * b[pos];
            result += p;
        } else {
            float d = a[pos] - b[pos];
            float p = d * d;
            result += p;
        }
        ++pos;
    }
    return result;
}
#endif

static bool g_force_scalar = false;  // tests flip this to cross-check AVX vs scalar emulation

#if VO_HAVE_AVX2
// simd_avx512.rs:271-352 in one pass: 12 accumulators (dot, |a|^2, |b|^2 x 4).  The three trees are
// independent, so this has the same bits as three separate wide32<0> passes (the tests check it).
static void cosine_parts_avx(const float* a, const float* b, size_t len, float& dot, float& na, float& nb) {
    __m256 d0 = _mm256_setzero_ps(), d1 = d0, d2 = d0, d3 = d0;
    __m256 x0 = d0, x1 = d0, x2 = d0, x3 = d0, y0 = d0, y1 = d0, y2 = d0, y3 = d0;
    size_t simd_len = len / 32;
    for (size_t it = 0; it < simd_len; ++it) {
        const float* pa = a + it * 32;
        const float* pb = b + it * 32;
        __m256 a0 = _mm256_loadu_ps(pa), b0 = _mm256_loadu_ps(pb);
        d0 = _mm256_fmadd_ps(a0, b0, d0);
        x0 = _mm256_fmadd_ps(a0, a0, x0);
        y0 = _mm256_fmadd_ps(b0, b0, y0);
        __m256 a1 = _mm256_loadu_ps(pa + 8), b1 = _mm256_loadu_ps(pb + 8);
        d1 = _mm256_fmadd_ps(a1, b1, d1);
        x1 = _mm256_fmadd_ps(a1, a1, x1);
        y1 = _mm256_fmadd_ps(b1, b1, y1);
        __m256 a2 = _mm256_loadu_ps(pa + 16), b2 = _mm256_loadu_ps(pb + 16);
        d2 = _mm256_fmadd_ps(a2, b2, d2);
        x2 = _mm256_fmadd_ps(a2, a2, x2);
        y2 = _mm256_fmadd_ps(b2, b2, y2);
        __m256 a3 = _mm256_loadu_ps(pa + 24), b3 = _mm256_loadu_ps(pb + 24);
        d3 = _mm256_fmadd_ps(a3, b3, d3);
        x3 = _mm256_fmadd_ps(a3, a3, x3);
        y3 = _mm256_fmadd_ps(b3, b3, y3);
    }
    dot = hsum8_avx(_mm256_add_ps(_mm256_add_ps(d0, d1), _mm256_add_ps(d2, d3)));
    na = hsum8_avx(_mm256_add_ps(_mm256_add_ps(x0, x1), _mm256_add_ps(x2, x3)));
    nb = hsum8_avx(_mm256_add_ps(_mm256_add_ps(y0, y1), _mm256_add_ps(y2, y3)));
    size_t pos = simd_len * 32;
    while (pos + 8 <= len) {
        __m256 va = _mm256_loadu_ps(a + pos), vb = _mm256_loadu_ps(b + pos);
        dot += hsum8_avx(_mm256_fmadd_ps(va, vb, _mm256_setzero_ps()));
        na += hsum8_avx(_mm256_fmadd_ps(va, va, _mm256_setzero_ps()));
        nb += hsum8_avx(_mm256_fmadd_ps(vb, vb, _mm256_setzero_ps()));
        pos += 8;
    }
    while (pos < len) {
        float ai = a[pos], bi = b[pos];
        float p0 = ai * bi, p1 = ai * ai, p2 = bi * bi;
        dot += p0;
        na += p1;
        nb += p2;
        ++pos;
    }
}
#endif

template <int OP>
static inline float wide32(const float* a, const float* b, size_t len, bool fma) {
#if VO_HAVE_AVX2
    if (fma && !g_force_scalar) return wide32_avx<OP>(a, b, len);
#endif
    return wide32_scalar<OP>(a, b, len, fma);
}

// simd_explicit.rs:50-78 / 104-130: single f32x8 accumulator, scalar tail.
template <int OP>
static float narrow8(const float* a, const float* b, size_t len, bool fma) {
    float S[8];
    for (int j = 0; j < 8; ++j) S[j] = 0.0f;
    size_t simd_len = len / 8;
    for (size_t it = 0; it < simd_len; ++it) {
        for (int j = 0; j < 8; ++j) {
            float x = a[it * 8 + j], y = b[it * 8 + j];
            if (OP == 0) {
                S[j] = madd(x, y, S[j], fma);
            } else {
                float d = x - y;
                S[j] = madd(d, d, S[j], fma);
            }
        }
    }
    float result = hsum8(S);
    for (size_t i = simd_len * 8; i < len; ++i) {
        if (OP == 0) {
            float p = a[i] * b[i];
            result += p;
        } else {
            float d = a[i] - b[i];
            float p = d * d;
            result += p;
        }
    }
    return result;
}

// simd_avx512.rs:87-97
static float dot_product_auto(const float* a, const float* b, size_t len, bool fma) {
    if (len >= 16) return wide32<0>(a, b, len, fma);
    return narrow8<0>(a, b, len, fma);
}
// simd_avx512.rs:106-114
static float squared_l2_auto(const float* a, const float* b, size_t len, bool fma) {
    if (len >= 16) return wide32<1>(a, b, len, fma);
    return narrow8<1>(a, b, len, fma);
}
// simd_avx512.rs:119-121
static float euclidean_auto(const float* a, const float* b, size_t len, bool fma) {
    return std::sqrt(squared_l2_auto(a, b, len, fma));
}
// simd_avx512.rs:130-138,271-352 and simd_explicit.rs:145-189.  The three
// quantities (dot, |a|^2, |b|^2) use independent accumulator trees, so each is
// exactly the dot-product tree applied to (a,b), (a,a), (b,b).
static float cosine_similarity_auto(const float* a, const float* b, size_t len, bool fma) {
    float dot, na, nb;
#if VO_HAVE_AVX2
    if (len >= 16 && fma && !g_force_scalar) {
        cosine_parts_avx(a, b, len, dot, na, nb);
    } else
#endif
    if (len >= 16) {
        dot = wide32<0>(a, b, len, fma);
        na = wide32<0>(a, a, len, fma);
        nb = wide32<0>(b, b, len, fma);
    } else {
        dot = narrow8<0>(a, b, len, fma);
        na = narrow8<0>(a, a, len, fma);
        nb = narrow8<0>(b, b, len, fma);
    }
    float norm_a = std::sqrt(na), norm_b = std::sqrt(nb);
    if (norm_a == 0.0f || norm_b == 0.0f) return 0.0f;
    float den = norm_a * norm_b;
    return dot / den;
}
// |v|^2 with the cosine tree (used by tests that check the hoisted-norm identity)
static float norm_sq_tree(const float* a, size_t len, bool fma) {
    return len >= 16 ? wide32<0>(a, a, len, fma) : narrow8<0>(a, a, len, fma);
}

// simd_explicit.rs:256-287
static uint32_t hamming_f32_u32(const float* a, const float* b, size_t len) {
    uint32_t c = 0;
    for (size_t i = 0; i < len; ++i) c += ((a[i] > 0.5f) != (b[i] > 0.5f)) ? 1u : 0u;
    return c;
}
// simd_explicit.rs:308-360
static uint32_t hamming_binary(const uint64_t* a, const uint64_t* b, size_t words) {
    uint32_t c = 0;
    for (size_t i = 0; i < words; ++i) c += (uint32_t)__builtin_popcountll(a[i] ^ b[i]);
    return c;
}
// simd_explicit.rs:372-443
static float jaccard_f32(const float* a, const float* b, size_t len) {
    uint32_t inter = 0, uni = 0;
    for (size_t i = 0; i < len; ++i) {
        bool x = a[i] > 0.5f, y = b[i] > 0.5f;
        inter += (x && y) ? 1u : 0u;
        uni += (x || y) ? 1u : 0u;
    }
    if (uni == 0) return 1.0f;
    return (float)inter / (float)uni;
}

// native/distance.rs:75-85 (SimdDistance::distance) -- the in-graph distance.
static float graph_distance(int metric, const float* a, const float* b, size_t len, bool fma) {
    switch (metric) {
        case COSINE: return 1.0f - cosine_similarity_auto(a, b, len, fma);
        case EUCLIDEAN: return euclidean_auto(a, b, len, fma);
        case DOT: return -dot_product_auto(a, b, len, fma);
        case HAMMING: return (float)hamming_f32_u32(a, b, len);
        case JACCARD: return 1.0f - jaccard_f32(a, b, len);
    }
    return 0.0f;
}
// index/hnsw/index/search.rs:30-38 (HnswIndex::compute_distance) -- the
// brute-force / rerank "metric value".
static float metric_value(int metric, const float* a, const float* b, size_t len, bool fma) {
    switch (metric) {
        case COSINE: return cosine_similarity_auto(a, b, len, fma);
        case EUCLIDEAN: return euclidean_auto(a, b, len, fma);
        case DOT: return dot_product_auto(a, b, len, fma);
        case HAMMING: return (float)hamming_f32_u32(a, b, len);
        case JACCARD: return jaccard_f32(a, b, len);
    }
    return 0.0f;
}
// core/distance.rs:76-81
static bool higher_is_better(int metric) { return metric == COSINE || metric == DOT || metric == JACCARD; }
// native/backend_adapter.rs:160-168
static float transform_score(int metric, float raw) {
    switch (metric) {
        case COSINE: {
            float s = 1.0f - raw;
            // f32::clamp(0,1): NaN stays NaN
            if (s < 0.0f) s = 0.0f;
            if (s > 1.0f) s = 1.0f;
            return s;
        }
        case DOT: return -raw;
        default: return raw;
    }
}

// f32::total_cmp as an integer key (ordered_float.rs:31-36)
static inline int32_t total_key(float f) {
    int32_t b;
    std::memcpy(&b, &f, 4);
    b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
    return b;
}
static inline int total_cmp(float a, float b) {
    int32_t x = total_key(a), y = total_key(b);
    return x < y ? -1 : (x > y ? 1 : 0);
}

// ---------------------------------------------------------------------------
// Rust std BinaryHeap (max-heap w.r.t. Less)
// ---------------------------------------------------------------------------
template <typename T, typename Less>
struct RustHeap {
    std::vector<T> data;
    Less less;
    bool le(const T& a, const T& b) const { return !less(b, a); }  // a <= b
    size_t size() const { return data.size(); }
    bool empty() const { return data.empty(); }
    const T& peek() const { return data[0]; }
    void sift_up(size_t start, size_t pos) {
        T elt = data[pos];
        while (pos > start) {
            size_t parent = (pos - 1) / 2;
            if (le(elt, data[parent])) break;
            data[pos] = data[parent];
            pos = parent;
        }
        data[pos] = elt;
    }
    void sift_down_to_bottom(size_t pos) {
        size_t end = data.size();
        size_t start = pos;
        T elt = data[pos];
        size_t child = 2 * pos + 1;
        size_t lim = end >= 2 ? end - 2 : 0;
        while (child <= lim && end >= 2) {
            if (le(data[child], data[child + 1])) child += 1;
            data[pos] = data[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1 && end >= 1) {
            data[pos] = data[child];
            pos = child;
        }
        data[pos] = elt;
        sift_up(start, pos);
    }
    void push(const T& v) {
        size_t old = data.size();
        data.push_back(v);
        sift_up(0, old);
    }
    T pop() {
        T item = data.back();
        data.pop_back();
        if (!data.empty()) {
            std::swap(item, data[0]);
            sift_down_to_bottom(0);
        }
        return item;
    }
};

struct DN {
    float d;
    uint64_t n;
};
// (OrderedFloat, NodeId) tuple order: total_cmp on dist, then id.
struct DNLess {
    bool operator()(const DN& a, const DN& b) const {
        int c = total_cmp(a.d, b.d);
        if (c != 0) return c < 0;
        return a.n < b.n;
    }
};
// Reverse<(OrderedFloat, NodeId)>
struct DNGreater {
    bool operator()(const DN& a, const DN& b) const { return DNLess()(b, a); }
};

// ---------------------------------------------------------------------------
// NativeHnsw (graph.rs)
// ---------------------------------------------------------------------------
// FxHashSet<usize> stand-in (rustc-hash 2.1.1 is a multiply hash; only membership matters)
struct FxSet {
    std::vector<uint64_t> slots;
    size_t mask, count = 0;
    explicit FxSet(size_t cap_pow2 = 1024) : slots(cap_pow2, UINT64_MAX), mask(cap_pow2 - 1) {}
    static inline size_t h(uint64_t k) { return (size_t)((k * 0xf1357aea2e62a9c5ull) >> 20); }
    void grow() {
        std::vector<uint64_t> old;
        old.swap(slots);
        slots.assign(old.size() * 2, UINT64_MAX);
        mask = slots.size() - 1;
        count = 0;
        for (uint64_t k : old)
            if (k != UINT64_MAX) insert(k);
    }
    bool insert(uint64_t k) {  // true if newly inserted
        if ((count + 1) * 2 > slots.size()) grow();
        size_t i = h(k) & mask;
        while (slots[i] != UINT64_MAX) {
            if (slots[i] == k) return false;
            i = (i + 1) & mask;
        }
        slots[i] = k;
        ++count;
        return true;
    }
};

// simd.rs:40-51 calculate_prefetch_distance, simd.rs:80-107 prefetch_vector (first cache line, T0)
static inline size_t prefetch_distance(size_t dim) {
    size_t raw = dim * 4 / 64;
    return raw < 4 ? 4 : (raw > 16 ? 16 : raw);
}
static inline void prefetch_vector(const float* p) {
#if VO_HAVE_AVX2
    _mm_prefetch((const char*)p, _MM_HINT_T0);
#else
    (void)p;
#endif
}

struct SearchStats {
    uint64_t ndc0 = 0;       // distance evaluations on layer 0 (incl. entry point)
    uint64_t hops0 = 0;      // layer-0 expansions (pops that were not the break)
    uint64_t ndc_up = 0;     // distance evaluations on layers >= 1
    uint64_t hops_up = 0;    // adjacency scans on layers >= 1
    uint64_t tie_at_k = 0;   // 1 if dist[k-1] == dist[k] in the ef-sized result (tie crosses k)
    uint64_t adj0 = 0;       // adjacency ids read on layer 0
};

// read-only view of one adjacency list (a Vec<usize> in the reference, native/layer.rs:12-15)
struct Span {
    const uint32_t* p = nullptr;
    size_t n = 0;
    const uint32_t* begin() const { return p; }
    const uint32_t* end() const { return p + n; }
    size_t size() const { return n; }
    uint32_t operator[](size_t i) const { return p[i]; }
    operator std::vector<uint32_t>() const { return std::vector<uint32_t>(p, p + n); }  // get_neighbors clones
};

// Storage of the vectors the traversal reads.  STORE_F32 is the reference (Vec<Vec<f32>>, graph.rs:22).  The two
// compact forms exist for the BASELINE configs that exceed what the reference HNSW stores (SURVEY finding 0.5):
//   STORE_F16   half::f16 values (core/half_precision.rs:97), up-converted to f32 before the reference's f32 distance
//   STORE_BIN   packed bits, u64 words LSB first: hamming_distance_binary (simd_explicit.rs:308-360), whose count
//               equals the f32-lane Hamming (simd_explicit.rs:256-287) on {0,1} data; the query is thresholded > 0.5
enum { STORE_F32 = 0, STORE_F16 = 1, STORE_BIN = 2 };

static inline float half_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {  // subnormal half -> normal float
            int e = -1;
            do {
                man <<= 1;
                ++e;
            } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112) << 23) | (man << 13);
    }
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}

struct Hnsw {
    int metric;
    uint32_t dim;
    uint32_t M, M0, ef_c;
    float alpha;
    bool fma;
    double level_mult;
    uint64_t rng_state = 0x5DEECE66D1A4B5B5ull;
    std::vector<float> vectors;  // contiguous n*dim
    uint64_t n = 0;
    std::vector<std::vector<std::vector<uint32_t>>> layers;  // [layer][node] -> ids
    // frozen form (search only): CSR adjacency per layer and, optionally, compact vector storage.  Pointers are either
    // owned (the own_* vectors) or borrowed from the caller, who keeps them alive.
    bool frozen = false;
    int store = STORE_F32;
    const void* ext_vectors = nullptr;
    std::vector<const uint64_t*> csr_rp;
    std::vector<const uint32_t*> csr_cols;
    std::vector<uint64_t> csr_nodes;
    std::vector<std::vector<uint64_t>> own_rp;
    std::vector<std::vector<uint32_t>> own_cols;
    std::vector<float> own_f32;
    uint32_t num_layers() const { return frozen ? (uint32_t)csr_nodes.size() : (uint32_t)layers.size(); }
    bool has_ep = false;
    uint64_t ep = 0;
    uint32_t max_layer = 0;
    uint64_t build_ndc = 0;

    Hnsw(int metric_, uint32_t dim_, uint32_t M_, uint32_t efc, float alpha_, bool fma_)
        : metric(metric_), dim(dim_), M(M_), M0(M_ * 2), ef_c(efc), alpha(alpha_), fma(fma_) {
        level_mult = 1.0 / std::log((double)M_);
        layers.emplace_back();
    }
    const float* vec(uint64_t id) const {
        return (frozen ? static_cast<const float*>(ext_vectors) : vectors.data()) + id * (size_t)dim;
    }
    float dist(const float* a, const float* b) const { return graph_distance(metric, a, b, dim, fma); }
    // distance of the query to stored node x, whatever the storage
    float qdist(const float* q, uint64_t x) const {
        if (store == STORE_F32) return dist(q, vec(x));
        if (store == STORE_F16) {
            static thread_local std::vector<float> row;
            row.resize(dim);
            const uint16_t* h = static_cast<const uint16_t*>(ext_vectors) + x * (size_t)dim;
            for (uint32_t i = 0; i < dim; ++i) row[i] = half_to_float(h[i]);
            return dist(q, row.data());
        }
        const uint64_t* a = packed_query(q);
        const uint64_t* b = static_cast<const uint64_t*>(ext_vectors) + x * (size_t)(dim / 64);
        uint32_t d = 0;
        for (uint32_t w = 0; w < dim / 64; ++w) d += (uint32_t)__builtin_popcountll(a[w] ^ b[w]);
        return (float)d;
    }
    // the query as packed bits (threshold > 0.5, simd_explicit.rs:256-287), cached per thread and per query pointer;
    // begin_query() drops the cache at the start of every search
    static const float*& packed_for() {
        static thread_local const float* p = nullptr;
        return p;
    }
    const uint64_t* packed_query(const float* q) const {
        static thread_local std::vector<uint64_t> bits;
        if (packed_for() != q) {
            bits.assign(dim / 64, 0);
            for (uint32_t i = 0; i < dim; ++i)
                if (q[i] > 0.5f) bits[i >> 6] |= 1ull << (i & 63);
            packed_for() = q;
        }
        return bits.data();
    }
    void begin_query() const { packed_for() = nullptr; }
    void prefetch_node(uint64_t x) const {
        if (store == STORE_F32) prefetch_vector(vec(x));
    }

    // graph.rs:368-403
    uint32_t random_layer() {
        uint64_t s = rng_state;
        if (s == 0) s = 0x853c49e6748fea9bull;
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        rng_state = s;
        double uniform = (double)s / (double)UINT64_MAX;  // u64::MAX as f64 == 2^64
        double safe = std::max(uniform, std::numeric_limits<double>::min());
        double lv = std::floor(-std::log(safe) * level_mult);
        uint64_t level = lv <= 0 ? 0 : (lv > 1e18 ? (uint64_t)1e18 : (uint64_t)lv);
        return (uint32_t)std::min<uint64_t>(level, 15);
    }

    Span neighbors(uint32_t layer, uint64_t node) const {
        if (frozen) {
            if (layer >= csr_nodes.size() || node >= csr_nodes[layer]) return Span{};
            const uint64_t a = csr_rp[layer][node], b = csr_rp[layer][node + 1];
            return Span{csr_cols[layer] + a, (size_t)(b - a)};
        }
        if (layer >= layers.size() || node >= layers[layer].size()) return Span{};
        const std::vector<uint32_t>& v = layers[layer][node];
        return Span{v.data(), v.size()};
    }

    // graph.rs:405-428
    uint64_t search_layer_single(const float* q, uint64_t entry, uint32_t layer, SearchStats* st) const {
        uint64_t best = entry;
        float best_dist = qdist(q, entry);
        if (st) st->ndc_up++;
        for (;;) {
            std::vector<uint32_t> nb = neighbors(layer, best);  // snapshot, as get_neighbors clones
            if (st) st->hops_up++;
            bool improved = false;
            for (uint32_t x : nb) {
                float d = qdist(q, x);
                if (st) st->ndc_up++;
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

    // graph.rs:438-520.  Returns results in the reference's order: backing-vector
    // order of the max-heap, then a stable sort by distance.
    std::vector<DN> search_layer(const float* q, const std::vector<uint64_t>& entries, size_t ef, uint32_t layer,
                                 SearchStats* st) const {
        FxSet visited;
        RustHeap<DN, DNGreater> cand;  // min-heap
        const size_t pd = prefetch_distance(dim);
        RustHeap<DN, DNLess> res;      // max-heap
        for (uint64_t e : entries) {
            float d = qdist(q, e);
            if (st) st->ndc0++;
            cand.push({d, e});
            res.push({d, e});
            visited.insert(e);
        }
        while (!cand.empty()) {
            DN c = cand.pop();
            float furthest = res.empty() ? std::numeric_limits<float>::max() : res.peek().d;
            if (c.d > furthest && res.size() >= ef) break;
            const Span nb = neighbors(layer, c.n);
            if (st) {
                st->hops0++;
                st->adj0 += nb.size();
            }
            // graph.rs:480-497: software prefetch of upcoming neighbour vectors for dim >= 384
            if (dim >= 384 && nb.size() > pd)
                for (size_t i = 0; i < pd; ++i)
                    if (nb[i] < n) prefetch_node(nb[i]);
            for (size_t i = 0; i < nb.size(); ++i) {
                const uint32_t x = nb[i];
                if (dim >= 384 && i + pd < nb.size() && nb[i + pd] < n) prefetch_node(nb[i + pd]);
                if (visited.insert(x)) {
                    float d = qdist(q, x);
                    if (st) st->ndc0++;
                    float f = res.empty() ? std::numeric_limits<float>::max() : res.peek().d;
                    if (d < f || res.size() < ef) {
                        cand.push({d, (uint64_t)x});
                        res.push({d, (uint64_t)x});
                        if (res.size() > ef) res.pop();
                    }
                }
            }
        }
        std::vector<DN> out = res.data;  // into_iter order
        std::stable_sort(out.begin(), out.end(), [](const DN& a, const DN& b) { return total_cmp(a.d, b.d) < 0; });
        return out;
    }

    // graph.rs:526-581
    std::vector<uint32_t> select_neighbors(const std::vector<DN>& cands, size_t maxn) {
        std::vector<uint32_t> sel;
        if (cands.empty()) return sel;
        if (cands.size() <= maxn) {
            for (auto& c : cands) sel.push_back((uint32_t)c.n);
            return sel;
        }
        for (auto& c : cands) {
            if (sel.size() >= maxn) break;
            const float* cv = vec(c.n);
            bool diverse = true;
            for (uint32_t s : sel) {
                float ds = dist(cv, vec(s));
                build_ndc++;
                if (!(alpha * c.d <= ds)) {
                    diverse = false;
                    break;  // Iterator::all short-circuits
                }
            }
            if (diverse || sel.empty()) sel.push_back((uint32_t)c.n);
        }
        if (sel.size() < maxn) {
            for (auto& c : cands) {
                if (sel.size() >= maxn) break;
                if (std::find(sel.begin(), sel.end(), (uint32_t)c.n) == sel.end()) sel.push_back((uint32_t)c.n);
            }
        }
        return sel;
    }

    // graph.rs:592-639
    void add_bidirectional_connection(uint64_t new_node, uint64_t nb, uint32_t layer, size_t maxc) {
        if (nb >= layers[layer].size()) return;  // Layer::set_neighbors bounds check
        std::vector<uint32_t>& cur = layers[layer][nb];
        if (cur.size() < maxc) {
            cur.push_back((uint32_t)new_node);
            return;
        }
        std::vector<uint32_t> all = cur;
        all.push_back((uint32_t)new_node);
        const float* nv = vec(nb);
        std::vector<DN> wd;
        wd.reserve(all.size());
        for (uint32_t x : all) {
            wd.push_back({dist(nv, vec(x)), (uint64_t)x});
            build_ndc++;
        }
        std::stable_sort(wd.begin(), wd.end(), [](const DN& a, const DN& b) { return total_cmp(a.d, b.d) < 0; });
        cur.clear();
        for (size_t i = 0; i < wd.size() && i < maxc; ++i) cur.push_back((uint32_t)wd[i].n);
    }

    // graph.rs:158-237
    uint64_t insert(const float* v) {
        uint64_t id = n;
        vectors.insert(vectors.end(), v, v + dim);
        n++;
        uint32_t node_layer = random_layer();
        while (layers.size() <= node_layer) layers.emplace_back(std::vector<std::vector<uint32_t>>(id + 1));
        for (auto& L : layers)
            if (L.size() <= id) L.resize(id + 1);
        if (has_ep) {
            uint64_t cur = ep;
            const float* q = vec(id);
            for (uint32_t l = max_layer; l >= node_layer + 1 && l > 0; --l) {
                cur = search_layer_single(q, cur, l, nullptr);
                if (l == 0) break;
            }
            for (int32_t l = (int32_t)node_layer; l >= 0; --l) {
                std::vector<DN> W = search_layer(vec(id), {cur}, ef_c, (uint32_t)l, nullptr);
                size_t maxc = (l == 0) ? M0 : M;
                std::vector<uint32_t> S = select_neighbors(W, maxc);
                layers[l][id] = S;
                for (uint32_t s : S) add_bidirectional_connection(id, s, (uint32_t)l, maxc);
                if (!W.empty()) cur = W[0].n;
            }
        } else {
            has_ep = true;
            ep = id;
        }
        if (node_layer > max_layer) {
            max_layer = node_layer;
            ep = id;
        }
        return id;
    }

    // graph.rs:288-348.  Advances rng_state like the reference's fetch_update; `used` receives the entry points.
    std::vector<DN> search_multi_entry(const float* q, size_t k, size_t ef, size_t probes, int order_mode, SearchStats* st,
                                       std::vector<uint64_t>* used) {
        std::vector<DN> out;
        if (!has_ep || n == 0) return out;
        begin_query();
        uint64_t cur = ep;
        for (uint32_t l = max_layer; l >= 1; --l) cur = search_layer_single(q, cur, l, st);
        std::vector<uint64_t> entries{cur};
        if (probes > 1 && n > 10) {
            for (size_t i = 1; i < std::min<size_t>(probes, 4); ++i) {
                uint64_t s = rng_state;  // no zero replacement on this path (graph.rs:319-333)
                s ^= s << 13;
                s ^= s >> 7;
                s ^= s << 17;
                rng_state = s;
                const uint64_t id = s % n;
                if (std::find(entries.begin(), entries.end(), id) == entries.end()) entries.push_back(id);
            }
        }
        if (used) *used = entries;
        std::vector<DN> c = search_layer(q, entries, ef, 0, st);
        if (order_mode == 1) std::sort(c.begin(), c.end(), DNLess());
        if (st && c.size() > k && k > 0 && total_cmp(c[k - 1].d, c[k].d) == 0) st->tie_at_k = 1;
        if (c.size() > k) c.resize(k);
        return c;
    }

    // graph.rs:251-270.  order_mode 0 = reference order (heap order among ties),
    // 1 = canonical (dist, id) order.
    std::vector<DN> search(const float* q, size_t k, size_t ef, int order_mode, SearchStats* st) const {
        std::vector<DN> out;
        if (!has_ep) return out;
        begin_query();
        uint64_t cur = ep;
        for (uint32_t l = max_layer; l >= 1; --l) cur = search_layer_single(q, cur, l, st);
        std::vector<DN> c = search_layer(q, {cur}, ef, 0, st);
        if (order_mode == 1) std::sort(c.begin(), c.end(), DNLess());
        if (st && c.size() > k && k > 0 && total_cmp(c[k - 1].d, c[k].d) == 0) st->tie_at_k = 1;
        if (c.size() > k) c.resize(k);
        return c;
    }
};

// ---------------------------------------------------------------------------
// SQ8 dual precision (native/quantization.rs, native/dual_precision.rs)
// ---------------------------------------------------------------------------
// ScalarQuantizer (quantization.rs:160-233) + QuantizedVectorStore (:318-374)
struct Sq8 {
    uint32_t dim = 0;
    std::vector<float> mn, scale, inv;
    std::vector<uint8_t> codes;  // count * dim
    uint64_t count = 0;

    // ScalarQuantizer::train, quantization.rs:190-233
    void train(const float* v, uint64_t n_train, uint32_t dim_) {
        dim = dim_;
        mn.assign(dim, std::numeric_limits<float>::max());
        std::vector<float> mx(dim, std::numeric_limits<float>::lowest());
        for (uint64_t r = 0; r < n_train; ++r)
            for (uint32_t i = 0; i < dim; ++i) {
                const float val = v[r * (size_t)dim + i];
                mn[i] = std::fmin(mn[i], val);  // f32::min: NaN loses
                mx[i] = std::fmax(mx[i], val);
            }
        scale.resize(dim);
        inv.resize(dim);
        for (uint32_t i = 0; i < dim; ++i) {
            const float range = mx[i] - mn[i];
            scale[i] = std::fabs(range) < 1e-10f ? 1.0f : 255.0f / range;
            inv[i] = 1.0f / scale[i];
        }
    }
    // ScalarQuantizer::quantize, quantization.rs:236-250: ((val - min) * scale).round().clamp(0, 255) as u8
    void quantize(const float* v, uint8_t* out) const {
        for (uint32_t i = 0; i < dim; ++i) {
            float q = std::round((v[i] - mn[i]) * scale[i]);  // half away from zero, as f32::round
            uint8_t b;
            if (q != q) b = 0;  // NaN survives clamp; `as u8` maps it to 0
            else b = (uint8_t)(q < 0.0f ? 0.0f : (q > 255.0f ? 255.0f : q));
            out[i] = b;
        }
    }
    void push(const float* v) {
        codes.resize((count + 1) * (size_t)dim);
        quantize(v, codes.data() + count * (size_t)dim);
        ++count;
    }
    const uint8_t* row(uint64_t id) const { return codes.data() + id * (size_t)dim; }
    // distance_l2_quantized_simd, quantization.rs:42-92 (u32 sums: the accumulator split cannot change the value)
    static uint32_t dist_q(const uint8_t* a, const uint8_t* b, uint32_t dim) {
        uint32_t s = 0;
        for (uint32_t i = 0; i < dim; ++i) {
            const int32_t d = (int32_t)a[i] - (int32_t)b[i];
            s += (uint32_t)(d * d);
        }
        return s;
    }
    // distance_l2_asymmetric_simd, quantization.rs:98-147 (four f32 accumulators, element i -> sum[i % 4], no FMA)
    float dist_asym(const float* q, const uint8_t* c) const {
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t chunks = dim / 4;
        for (uint32_t i = 0; i < chunks * 4; ++i) {
            const float dq = (float)c[i] * inv[i] + mn[i];
            const float d = q[i] - dq;
            sum[i & 3] += d * d;
        }
        for (uint32_t i = chunks * 4; i < dim; ++i) {
            const float dq = (float)c[i] * inv[i] + mn[i];
            const float d = q[i] - dq;
            sum[0] += d * d;
        }
        return std::sqrt((sum[0] + sum[1] + sum[2] + sum[3]));
    }
};

struct UN {
    uint32_t d;
    uint64_t n;
};
struct UNLess {  // (u32, NodeId) tuple order
    bool operator()(const UN& a, const UN& b) const { return a.d != b.d ? a.d < b.d : a.n < b.n; }
};
struct UNGreater {
    bool operator()(const UN& a, const UN& b) const { return UNLess()(b, a); }
};

// DualPrecisionHnsw::search_int8_traversal (dual_precision.rs:284-325) with search_layer_int8 (:327-405) and
// greedy_search_int8 (:407-441).  order_mode 0: the reference's order (results.into_iter() = heap array, stable
// sort by distance); 1: canonical coarse order (dist, id).  The exact re-rank is a stable sort by total_cmp over
// the coarse order either way.  stats[4] = 1 when a tie makes the two orders differ observably: equal coarse
// distances across the candidates_k cut, or equal exact distances among the first k+1 re-ranked.
static std::vector<DN> dual_search_int8(const Hnsw& g, const Sq8& sq, const float* q, size_t k, size_t ef_search,
                                        size_t oversampling, int order_mode, SearchStats* st) {
    std::vector<DN> out;
    if (!g.has_ep) return out;
    std::vector<uint8_t> qq(g.dim);
    sq.quantize(q, qq.data());
    const size_t ck = k * oversampling;
    auto dq = [&](uint64_t id) { return Sq8::dist_q(qq.data(), sq.row(id), g.dim); };

    uint64_t cur = g.ep;
    for (uint32_t l = g.max_layer; l >= 1; --l) {
        uint32_t cur_d = dq(cur);
        if (st) st->ndc_up++;
        for (;;) {
            std::vector<uint32_t> nb = g.neighbors(l, cur);
            if (st) st->hops_up++;
            bool improved = false;
            for (uint32_t x : nb) {
                const uint32_t d = dq(x);
                if (st) st->ndc_up++;
                if (d < cur_d) {
                    cur = x;
                    cur_d = d;
                    improved = true;
                }
            }
            if (!improved) break;
        }
    }

    FxSet visited;
    RustHeap<UN, UNGreater> cand;
    RustHeap<UN, UNLess> res;
    {
        const uint32_t d = dq(cur);
        if (st) st->ndc0++;
        cand.push({d, cur});
        res.push({d, cur});
        visited.insert(cur);
    }
    const size_t ef = std::max(ef_search, ck);
    while (!cand.empty()) {
        const UN c = cand.pop();
        const uint32_t furthest = res.empty() ? UINT32_MAX : res.peek().d;
        if (c.d > furthest && res.size() >= ef) break;
        const Span nb = g.neighbors(0, c.n);
        if (st) {
            st->hops0++;
            st->adj0 += nb.size();
        }
        for (uint32_t x : nb) {
            if (visited.insert(x)) {
                const uint32_t d = dq(x);
                if (st) st->ndc0++;
                const uint32_t f = res.empty() ? UINT32_MAX : res.peek().d;
                if (d < f || res.size() < ef) {
                    cand.push({d, (uint64_t)x});
                    res.push({d, (uint64_t)x});
                    if (res.size() > ef) res.pop();
                }
            }
        }
    }
    std::vector<UN> coarse = res.data;
    if (order_mode == 1)
        std::sort(coarse.begin(), coarse.end(), UNLess());
    else
        std::stable_sort(coarse.begin(), coarse.end(), [](const UN& a, const UN& b) { return a.d < b.d; });
    if (st && coarse.size() > ck && ck > 0 && coarse[ck - 1].d == coarse[ck].d) st->tie_at_k = 1;
    if (coarse.size() > ck) coarse.resize(ck);
    if (coarse.empty()) return out;

    out.reserve(coarse.size());
    for (const UN& c : coarse) out.push_back({g.dist(q, g.vec(c.n)), c.n});
    std::stable_sort(out.begin(), out.end(), [](const DN& a, const DN& b) { return total_cmp(a.d, b.d) < 0; });
    if (st)
        for (size_t i = 0; i + 1 < out.size() && i < k; ++i)
            if (total_cmp(out[i].d, out[i + 1].d) == 0) st->tie_at_k = 1;
    if (out.size() > k) out.resize(k);
    return out;
}

// ---------------------------------------------------------------------------
// BM25 (index/bm25.rs) on integer term ids.  Tokenisation lives in Python for
// the tests (tokenize :114-120 is string handling, not arithmetic).
// ---------------------------------------------------------------------------
struct Bm25 {
    float k1 = 1.2f, b = 0.75f;
    std::unordered_map<uint64_t, std::unordered_map<uint32_t, uint32_t>> doc_tf;  // doc -> term -> tf
    std::unordered_map<uint64_t, uint32_t> doc_len;
    std::unordered_map<uint32_t, std::unordered_set<uint32_t>> postings;  // term -> docs
    uint64_t doc_count = 0;
    uint64_t total_len = 0;

    // bm25.rs:134-206
    void add_document(uint64_t id, const uint32_t* terms, size_t nt) {
        if (nt == 0) return;
        std::unordered_map<uint32_t, uint32_t> tf;
        for (size_t i = 0; i < nt; ++i) tf[terms[i]]++;
        for (auto& kv : tf) postings[kv.first].insert((uint32_t)id);
        auto it = doc_len.find(id);
        if (it != doc_len.end()) {
            total_len = total_len >= it->second ? total_len - it->second : 0;
            // NOTE: like the reference, stale postings of the replaced document are kept.
        } else {
            doc_count++;
        }
        doc_tf[id] = std::move(tf);
        doc_len[id] = (uint32_t)nt;
        total_len += nt;
    }
    // bm25.rs:220-262
    bool remove_document(uint64_t id) {
        auto it = doc_tf.find(id);
        if (it == doc_tf.end()) return false;
        for (auto& kv : it->second) {
            auto p = postings.find(kv.first);
            if (p != postings.end()) {
                p->second.erase((uint32_t)id);
                if (p->second.empty()) postings.erase(p);
            }
        }
        uint32_t len = doc_len[id];
        doc_tf.erase(it);
        doc_len.erase(id);
        doc_count = doc_count ? doc_count - 1 : 0;
        total_len = total_len >= len ? total_len - len : 0;
        return true;
    }
    // bm25.rs:269-376.  Result order: score desc (total_cmp), ties canonicalised by id asc
    // (the reference's tie order is hash/roaring iteration + select_nth_unstable: unspecified).
    std::vector<std::pair<uint64_t, float>> search(const uint32_t* q, size_t nq, size_t k) const {
        std::vector<std::pair<uint64_t, float>> scores;
        if (nq == 0 || doc_count == 0) return scores;
        float avgdl = (float)total_len / (float)doc_count;
        float n = (float)doc_count;
        std::unordered_map<uint32_t, float> idf;
        for (size_t i = 0; i < nq; ++i) {
            auto p = postings.find(q[i]);
            size_t df = p == postings.end() ? 0 : p->second.size();
            float v = 0.0f;
            if (df != 0) {
                float df_f = (float)df;
                float num = n - df_f + 0.5f;
                float den = df_f + 0.5f;
                float r = num / den + 1.0f;
                v = std::log(r);  // f32::ln
            }
            idf[q[i]] = v;
        }
        std::unordered_set<uint32_t> cand;
        for (size_t i = 0; i < nq; ++i) {
            auto p = postings.find(q[i]);
            if (p != postings.end()) cand.insert(p->second.begin(), p->second.end());
        }
        for (uint32_t d32 : cand) {
            uint64_t d = d32;
            auto dt = doc_tf.find(d);
            if (dt == doc_tf.end()) continue;
            float dl = (float)doc_len.at(d);
            float len_norm = 1.0f - b + b * dl / avgdl;
            float s = 0.0f;
            for (size_t i = 0; i < nq; ++i) {
                auto t = dt->second.find(q[i]);
                float tf = t == dt->second.end() ? 0.0f : (float)t->second;
                float term = 0.0f;
                if (tf != 0.0f) {
                    float numerator = tf * (k1 + 1.0f);
                    float denominator = tf + k1 * len_norm;
                    term = idf.at(q[i]) * numerator / denominator;
                }
                s += term;
            }
            if (s > 0.0f) scores.push_back({d, s});
        }
        std::sort(scores.begin(), scores.end(), [](const auto& x, const auto& y) {
            int c = total_cmp(y.second, x.second);
            if (c != 0) return c < 0;
            return x.first < y.first;
        });
        if (scores.size() > k) scores.resize(k);
        return scores;
    }
};

}  // namespace vo

// ===========================================================================
// C API for ctypes
// ===========================================================================
using namespace vo;

extern "C" {

void vo_force_scalar(int on) { g_force_scalar = on != 0; }
int vo_have_avx2(void) { return VO_HAVE_AVX2; }

float vo_graph_distance(int metric, const float* a, const float* b, uint64_t len, int fma) {
    return graph_distance(metric, a, b, len, fma != 0);
}
float vo_metric_value(int metric, const float* a, const float* b, uint64_t len, int fma) {
    return metric_value(metric, a, b, len, fma != 0);
}
float vo_dot(const float* a, const float* b, uint64_t len, int fma) { return dot_product_auto(a, b, len, fma != 0); }
float vo_l2sq(const float* a, const float* b, uint64_t len, int fma) { return squared_l2_auto(a, b, len, fma != 0); }
float vo_norm_sq(const float* a, uint64_t len, int fma) { return norm_sq_tree(a, len, fma != 0); }
uint32_t vo_hamming_binary(const uint64_t* a, const uint64_t* b, uint64_t words) { return hamming_binary(a, b, words); }
float vo_transform_score(int metric, float raw) { return transform_score(metric, raw); }
int vo_higher_is_better(int metric) { return higher_is_better(metric) ? 1 : 0; }

// params.rs:309-319; quality: 0 Fast, 1 Balanced, 2 Accurate, 3 Perfect, 4 Custom(custom_ef)
uint64_t vo_ef_search(int quality, uint64_t k, uint64_t custom_ef) {
    switch (quality) {
        case 0: return std::max<uint64_t>(64, k * 2);
        case 1: return std::max<uint64_t>(128, k * 4);
        case 2: return std::max<uint64_t>(512, k * 16);
        case 3: return std::max<uint64_t>(4096, k * 100);
        default: return std::max<uint64_t>(custom_ef, k);
    }
}

void* vo_hnsw_new(int metric, uint32_t dim, uint32_t M, uint32_t ef_c, float alpha, int fma) {
    return new Hnsw(metric, dim, M, ef_c, alpha, fma != 0);
}
void vo_hnsw_free(void* h) { delete (Hnsw*)h; }
uint64_t vo_hnsw_insert(void* h, const float* v) { return ((Hnsw*)h)->insert(v); }
void vo_hnsw_insert_many(void* h, const float* v, uint64_t n) {
    Hnsw* g = (Hnsw*)h;
    g->vectors.reserve(g->vectors.size() + n * (size_t)g->dim);
    for (uint64_t i = 0; i < n; ++i) g->insert(v + i * (size_t)g->dim);
}
uint64_t vo_hnsw_len(void* h) { return ((Hnsw*)h)->n; }
uint32_t vo_hnsw_dim(void* h) { return ((Hnsw*)h)->dim; }
uint32_t vo_hnsw_num_layers(void* h) { return ((Hnsw*)h)->num_layers(); }
uint32_t vo_hnsw_max_layer(void* h) { return ((Hnsw*)h)->max_layer; }
uint32_t vo_hnsw_M(void* h) { return ((Hnsw*)h)->M; }
uint32_t vo_hnsw_M0(void* h) { return ((Hnsw*)h)->M0; }
int vo_hnsw_has_entry(void* h) { return ((Hnsw*)h)->has_ep ? 1 : 0; }
uint64_t vo_hnsw_entry_point(void* h) { return ((Hnsw*)h)->ep; }
uint64_t vo_hnsw_layer_nodes(void* h, uint32_t l) { return ((Hnsw*)h)->layers[l].size(); }
uint64_t vo_hnsw_layer_edges(void* h, uint32_t l) {
    uint64_t e = 0;
    for (auto& r : ((Hnsw*)h)->layers[l]) e += r.size();
    return e;
}
// CSR export: row_ptr has nodes+1 entries
void vo_hnsw_export_layer(void* h, uint32_t l, uint64_t* row_ptr, uint32_t* cols) {
    auto& L = ((Hnsw*)h)->layers[l];
    uint64_t e = 0;
    for (size_t i = 0; i < L.size(); ++i) {
        row_ptr[i] = e;
        for (uint32_t x : L[i]) cols[e++] = x;
    }
    row_ptr[L.size()] = e;
}
const float* vo_hnsw_vectors(void* h) {
    Hnsw* g = (Hnsw*)h;
    return g->frozen ? (g->store == STORE_F32 ? static_cast<const float*>(g->ext_vectors) : nullptr) : g->vectors.data();
}
// the first `count` levels the reference PRNG would assign (graph.rs:368-403)
void vo_levels(uint32_t M, uint64_t count, uint8_t* out) {
    Hnsw g(0, 1, M, 1, 1.0f, true);
    for (uint64_t i = 0; i < count; ++i) out[i] = (uint8_t)g.random_layer();
}

// Build an oracle index from arrays (e.g. a GPU-built graph) for the CPU baseline.
void* vo_hnsw_from_arrays(int metric, uint32_t dim, uint32_t M, uint32_t M0, uint32_t ef_c, int fma, const float* vectors,
                          uint64_t n, uint32_t num_layers, const uint64_t* const* row_ptrs, const uint32_t* const* cols,
                          const uint64_t* layer_nodes, uint64_t entry_point, uint32_t max_layer) {
    Hnsw* g = new Hnsw(metric, dim, M, ef_c, 1.0f, fma != 0);
    g->M0 = M0;
    g->vectors.assign(vectors, vectors + n * (size_t)dim);
    g->n = n;
    g->layers.clear();
    for (uint32_t l = 0; l < num_layers; ++l) {
        std::vector<std::vector<uint32_t>> L(layer_nodes[l]);
        for (uint64_t i = 0; i < layer_nodes[l]; ++i)
            L[i].assign(cols[l] + row_ptrs[l][i], cols[l] + row_ptrs[l][i + 1]);
        g->layers.push_back(std::move(L));
    }
    g->has_ep = n > 0;
    g->ep = entry_point;
    g->max_layer = max_layer;
    return g;
}

// A search-only index over caller-held arrays: CSR adjacency per layer and vectors in `storage` form (STORE_*).
// Nothing is copied: the caller keeps `vectors`, `row_ptrs[l]` and `cols[l]` alive as long as the handle.  This is
// how the CPU arm holds 10M+ node graphs (no per-node allocations) and the f16 / packed-bit configs.
void* vo_hnsw_frozen(int metric, uint32_t dim, uint32_t M, uint32_t M0, uint32_t ef_c, int fma, int storage,
                     const void* vectors, uint64_t n, uint32_t num_layers, const uint64_t* const* row_ptrs,
                     const uint32_t* const* cols, const uint64_t* layer_nodes, uint64_t entry_point, uint32_t max_layer) {
    if (storage < STORE_F32 || storage > STORE_BIN) return nullptr;
    if (storage == STORE_BIN && (dim % 64 != 0 || metric != HAMMING)) return nullptr;
    Hnsw* g = new Hnsw(metric, dim, M, ef_c, 1.0f, fma != 0);
    g->M0 = M0;
    g->n = n;
    g->frozen = true;
    g->store = storage;
    g->ext_vectors = vectors;
    for (uint32_t l = 0; l < num_layers; ++l) {
        g->csr_rp.push_back(row_ptrs[l]);
        g->csr_cols.push_back(cols[l]);
        g->csr_nodes.push_back(layer_nodes[l]);
    }
    g->has_ep = n > 0;
    g->ep = entry_point;
    g->max_layer = max_layer;
    return g;
}

// The same from a format-v1 `.graph` file (native/backend_adapter.rs:213-261), adjacency owned as CSR.  `vectors`
// NULL: the f32 vectors are read from `{basename}.vectors` (storage must be STORE_F32); otherwise borrowed as above.
void* vo_hnsw_open(const char* dir, const char* basename, int metric, int fma, int storage, const void* vectors,
                   uint64_t n_vectors, uint32_t dim_vectors) {
    const std::string vp = std::string(dir) + "/" + basename + ".vectors";
    const std::string gp = std::string(dir) + "/" + basename + ".graph";
    uint32_t version = 0, dim = dim_vectors;
    uint64_t count = n_vectors;
    std::vector<float> vecs;
    if (!vectors) {
        if (storage != STORE_F32) return nullptr;
        FILE* f = std::fopen(vp.c_str(), "rb");
        if (!f) return nullptr;
        bool okv = std::fread(&version, 4, 1, f) == 1 && version == 1 && std::fread(&count, 8, 1, f) == 1 &&
                   std::fread(&dim, 4, 1, f) == 1;
        if (okv) {
            vecs.resize(count * (size_t)dim);
            okv = std::fread(vecs.data(), 4, vecs.size(), f) == vecs.size();
        }
        std::fclose(f);
        if (!okv) return nullptr;
    }
    FILE* f = std::fopen(gp.c_str(), "rb");
    if (!f) return nullptr;
    uint32_t nl, M, M0, efc, maxl;
    uint64_t ep, cnt2;
    bool ok = std::fread(&version, 4, 1, f) == 1 && version == 1 && std::fread(&nl, 4, 1, f) == 1 &&
              std::fread(&M, 4, 1, f) == 1 && std::fread(&M0, 4, 1, f) == 1 && std::fread(&efc, 4, 1, f) == 1 &&
              std::fread(&ep, 8, 1, f) == 1 && std::fread(&maxl, 4, 1, f) == 1 && std::fread(&cnt2, 8, 1, f) == 1;
    if (!ok || nl == 0 || nl > 16) {
        std::fclose(f);
        return nullptr;
    }
    Hnsw* g = new Hnsw(metric, dim, M, efc, 1.0f, fma != 0);
    g->M0 = M0;
    g->n = count;
    g->frozen = true;
    g->store = storage;
    g->own_rp.resize(nl);
    g->own_cols.resize(nl);
    for (uint32_t l = 0; l < nl && ok; ++l) {
        uint64_t nn;
        if (std::fread(&nn, 8, 1, f) != 1 || nn > count) {
            ok = false;
            break;
        }
        auto& rp = g->own_rp[l];
        auto& cl = g->own_cols[l];
        rp.resize(nn + 1);
        rp[0] = 0;
        if (l == 0) cl.reserve(nn * (size_t)M0);
        for (uint64_t i = 0; i < nn; ++i) {
            uint32_t deg;
            if (std::fread(&deg, 4, 1, f) != 1 || deg > 4096) {
                ok = false;
                break;
            }
            const size_t base = cl.size();
            cl.resize(base + deg);
            if (deg && std::fread(cl.data() + base, 4, deg, f) != deg) {
                ok = false;
                break;
            }
            rp[i + 1] = base + deg;
        }
        g->csr_nodes.push_back(nn);
    }
    std::fclose(f);
    if (!ok) {
        delete g;
        return nullptr;
    }
    for (uint32_t l = 0; l < nl; ++l) {
        if (g->own_cols[l].empty()) g->own_cols[l].push_back(0);
        g->csr_rp.push_back(g->own_rp[l].data());
        g->csr_cols.push_back(g->own_cols[l].data());
    }
    if (vectors) {
        g->ext_vectors = vectors;
    } else {
        g->own_f32 = std::move(vecs);
        g->ext_vectors = g->own_f32.data();
    }
    g->has_ep = true;  // file_load always sets Some(entry_point) (backend_adapter.rs:368)
    g->ep = ep;
    g->max_layer = maxl;
    return g;
}

// stats layout: [ndc0, hops0, ndc_up, hops_up, tie_at_k, adj0]
uint32_t vo_hnsw_search(void* h, const float* q, uint32_t k, uint32_t ef, int order_mode, uint64_t* out_ids,
                        float* out_dist, uint64_t* stats) {
    SearchStats st;
    std::vector<DN> r = ((Hnsw*)h)->search(q, k, ef, order_mode, &st);
    for (size_t i = 0; i < r.size(); ++i) {
        out_ids[i] = r[i].n;
        out_dist[i] = r[i].d;
    }
    if (stats) {
        stats[0] = st.ndc0;
        stats[1] = st.hops0;
        stats[2] = st.ndc_up;
        stats[3] = st.hops_up;
        stats[4] = st.tie_at_k;
        stats[5] = st.adj0;
    }
    return (uint32_t)r.size();
}

// search_multi_entry; out_entries gets 4 slots (UINT64_MAX padded): the entry points actually used
uint32_t vo_hnsw_search_multi_entry(void* h, const float* q, uint32_t k, uint32_t ef, uint32_t probes, int order_mode,
                                    uint64_t* out_ids, float* out_dist, uint64_t* stats, uint64_t* out_entries) {
    SearchStats st;
    std::vector<uint64_t> used;
    std::vector<DN> r = ((Hnsw*)h)->search_multi_entry(q, k, ef, probes, order_mode, &st, &used);
    for (size_t i = 0; i < r.size(); ++i) {
        out_ids[i] = r[i].n;
        out_dist[i] = r[i].d;
    }
    for (size_t i = 0; i < 4; ++i) out_entries[i] = i < used.size() ? used[i] : UINT64_MAX;
    if (stats) {
        stats[0] = st.ndc0;
        stats[1] = st.hops0;
        stats[2] = st.ndc_up;
        stats[3] = st.hops_up;
        stats[4] = st.tie_at_k;
        stats[5] = st.adj0;
    }
    return (uint32_t)r.size();
}
uint64_t vo_hnsw_rng_state(void* h) { return ((Hnsw*)h)->rng_state; }
void vo_hnsw_set_rng_state(void* h, uint64_t s) { ((Hnsw*)h)->rng_state = s; }

// rayon par_iter stand-in (index/hnsw/index/batch.rs:178-196): a static thread pool over queries.
void vo_hnsw_search_batch(void* h, const float* q, uint64_t nq, uint32_t k, uint32_t ef, int order_mode, int threads,
                          uint64_t* out_ids, float* out_dist, uint32_t* out_counts, uint64_t* stats) {
    Hnsw* g = (Hnsw*)h;
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        for (;;) {
            uint64_t i = next.fetch_add(1);
            if (i >= nq) break;
            out_counts[i] = vo_hnsw_search(g, q + i * (size_t)g->dim, k, ef, order_mode, out_ids + i * (size_t)k,
                                           out_dist + i * (size_t)k, stats ? stats + i * 6 : nullptr);
            for (uint32_t j = out_counts[i]; j < k; ++j) {
                out_ids[i * (size_t)k + j] = UINT64_MAX;
                out_dist[i * (size_t)k + j] = std::numeric_limits<float>::quiet_NaN();
            }
        }
    };
    if (threads <= 1) {
        work();
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
}

// search_layer alone (for build-parity tests): returns reference-order list
uint32_t vo_hnsw_search_layer(void* h, const float* q, uint64_t entry, uint32_t ef, uint32_t layer, uint64_t* out_ids,
                              float* out_dist) {
    std::vector<DN> r = ((Hnsw*)h)->search_layer(q, {entry}, ef, layer, nullptr);
    for (size_t i = 0; i < r.size(); ++i) {
        out_ids[i] = r[i].n;
        out_dist[i] = r[i].d;
    }
    return (uint32_t)r.size();
}

// select_neighbors on explicit candidates (graph_tests.rs:44-160)
uint32_t vo_hnsw_select_neighbors(void* h, const uint64_t* cand_ids, const float* cand_d, uint32_t nc, uint32_t maxn,
                                  uint32_t* out) {
    std::vector<DN> c(nc);
    for (uint32_t i = 0; i < nc; ++i) c[i] = {cand_d[i], cand_ids[i]};
    std::vector<uint32_t> s = ((Hnsw*)h)->select_neighbors(c, maxn);
    for (size_t i = 0; i < s.size(); ++i) out[i] = s[i];
    return (uint32_t)s.size();
}

// Brute force (index/hnsw/index/search.rs:176-219): metric value, stable sort by
// sort_results over idx-ascending input (ties -> lower idx first), truncate k.
uint32_t vo_bruteforce(int metric, const float* vectors, uint64_t n, uint32_t dim, const float* q, uint32_t k, int fma,
                       uint64_t* out_ids, float* out_score) {
    std::vector<DN> r(n);
    for (uint64_t i = 0; i < n; ++i) r[i] = {metric_value(metric, q, vectors + i * (size_t)dim, dim, fma != 0), i};
    bool hib = higher_is_better(metric);
    std::stable_sort(r.begin(), r.end(), [hib](const DN& a, const DN& b) {
        return hib ? total_cmp(b.d, a.d) < 0 : total_cmp(a.d, b.d) < 0;
    });
    uint32_t m = (uint32_t)std::min<uint64_t>(k, n);
    for (uint32_t i = 0; i < m; ++i) {
        out_ids[i] = r[i].n;
        out_score[i] = r[i].d;
    }
    return m;
}
void vo_bruteforce_batch(int metric, const float* vectors, uint64_t n, uint32_t dim, const float* q, uint64_t nq,
                         uint32_t k, int fma, int threads, uint64_t* out_ids, float* out_score) {
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        for (;;) {
            uint64_t i = next.fetch_add(1);
            if (i >= nq) break;
            uint32_t m = vo_bruteforce(metric, vectors, n, dim, q + i * (size_t)dim, k, fma, out_ids + i * (size_t)k,
                                       out_score + i * (size_t)k);
            for (uint32_t j = m; j < k; ++j) {
                out_ids[i * (size_t)k + j] = UINT64_MAX;
                out_score[i * (size_t)k + j] = std::numeric_limits<float>::quiet_NaN();
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < std::max(1, threads); ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
}
// Packed-binary brute force (integer Hamming), ties -> lower idx first.
uint32_t vo_bruteforce_binary(const uint64_t* vectors, uint64_t n, uint32_t words, const uint64_t* q, uint32_t k,
                              uint64_t* out_ids, uint32_t* out_dist) {
    std::vector<std::pair<uint32_t, uint64_t>> r(n);
    for (uint64_t i = 0; i < n; ++i) r[i] = {hamming_binary(q, vectors + i * (size_t)words, words), i};
    std::sort(r.begin(), r.end());
    uint32_t m = (uint32_t)std::min<uint64_t>(k, n);
    for (uint32_t i = 0; i < m; ++i) {
        out_ids[i] = r[i].second;
        out_dist[i] = r[i].first;
    }
    return m;
}

// File format v1 (native/backend_adapter.rs:184-261 dump, :274-380 load)
int vo_hnsw_dump(void* h, const char* dir, const char* basename) {
    Hnsw* g = (Hnsw*)h;
    std::string vp = std::string(dir) + "/" + basename + ".vectors";
    std::string gp = std::string(dir) + "/" + basename + ".graph";
    FILE* f = std::fopen(vp.c_str(), "wb");
    if (!f) return -1;
    uint32_t version = 1, dim = g->n ? g->dim : 0;
    uint64_t count = g->n;
    std::fwrite(&version, 4, 1, f);
    std::fwrite(&count, 8, 1, f);
    std::fwrite(&dim, 4, 1, f);
    std::fwrite(g->vectors.data(), 4, g->vectors.size(), f);
    std::fclose(f);
    f = std::fopen(gp.c_str(), "wb");
    if (!f) return -1;
    uint32_t nl = (uint32_t)g->layers.size();
    uint64_t ep = g->has_ep ? g->ep : 0;
    std::fwrite(&version, 4, 1, f);
    std::fwrite(&nl, 4, 1, f);
    std::fwrite(&g->M, 4, 1, f);
    std::fwrite(&g->M0, 4, 1, f);
    std::fwrite(&g->ef_c, 4, 1, f);
    std::fwrite(&ep, 8, 1, f);
    std::fwrite(&g->max_layer, 4, 1, f);
    std::fwrite(&count, 8, 1, f);
    for (auto& L : g->layers) {
        uint64_t nn = L.size();
        std::fwrite(&nn, 8, 1, f);
        for (auto& r : L) {
            uint32_t deg = (uint32_t)r.size();
            std::fwrite(&deg, 4, 1, f);
            if (deg) std::fwrite(r.data(), 4, deg, f);
        }
    }
    std::fclose(f);
    return 0;
}
void* vo_hnsw_load(const char* dir, const char* basename, int metric, int fma) {
    std::string vp = std::string(dir) + "/" + basename + ".vectors";
    std::string gp = std::string(dir) + "/" + basename + ".graph";
    FILE* f = std::fopen(vp.c_str(), "rb");
    if (!f) return nullptr;
    uint32_t version = 0, dim = 0;
    uint64_t count = 0;
    if (std::fread(&version, 4, 1, f) != 1 || version != 1) {
        std::fclose(f);
        return nullptr;
    }
    if (std::fread(&count, 8, 1, f) != 1 || std::fread(&dim, 4, 1, f) != 1) {
        std::fclose(f);
        return nullptr;
    }
    std::vector<float> vecs(count * (size_t)dim);
    if (std::fread(vecs.data(), 4, vecs.size(), f) != vecs.size()) {
        std::fclose(f);
        return nullptr;
    }
    std::fclose(f);
    f = std::fopen(gp.c_str(), "rb");
    if (!f) return nullptr;
    uint32_t nl, M, M0, efc, maxl;
    uint64_t ep, cnt2;
    bool ok = std::fread(&version, 4, 1, f) == 1 && version == 1 && std::fread(&nl, 4, 1, f) == 1 &&
              std::fread(&M, 4, 1, f) == 1 && std::fread(&M0, 4, 1, f) == 1 && std::fread(&efc, 4, 1, f) == 1 &&
              std::fread(&ep, 8, 1, f) == 1 && std::fread(&maxl, 4, 1, f) == 1 && std::fread(&cnt2, 8, 1, f) == 1;
    if (!ok) {
        std::fclose(f);
        return nullptr;
    }
    Hnsw* g = new Hnsw(metric, dim, M, efc, 1.0f, fma != 0);
    g->M0 = M0;
    g->vectors = std::move(vecs);
    g->n = count;
    g->layers.clear();
    for (uint32_t l = 0; l < nl; ++l) {
        uint64_t nn;
        if (std::fread(&nn, 8, 1, f) != 1) {
            ok = false;
            break;
        }
        std::vector<std::vector<uint32_t>> L(nn);
        for (uint64_t i = 0; i < nn && ok; ++i) {
            uint32_t deg;
            if (std::fread(&deg, 4, 1, f) != 1) {
                ok = false;
                break;
            }
            L[i].resize(deg);
            if (deg && std::fread(L[i].data(), 4, deg, f) != deg) ok = false;
        }
        g->layers.push_back(std::move(L));
    }
    std::fclose(f);
    if (!ok) {
        delete g;
        return nullptr;
    }
    g->has_ep = true;  // file_load always sets Some(entry_point) (backend_adapter.rs:368)
    g->ep = ep;
    g->max_layer = maxl;
    return g;
}

// ---- BM25 ----
void* vo_bm25_new(float k1, float b) {
    Bm25* x = new Bm25();
    x->k1 = k1;
    x->b = b;
    return x;
}
void vo_bm25_free(void* h) { delete (Bm25*)h; }
void vo_bm25_add(void* h, uint64_t id, const uint32_t* terms, uint64_t nt) { ((Bm25*)h)->add_document(id, terms, nt); }
int vo_bm25_remove(void* h, uint64_t id) { return ((Bm25*)h)->remove_document(id) ? 1 : 0; }
uint64_t vo_bm25_len(void* h) { return ((Bm25*)h)->doc_count; }
uint64_t vo_bm25_term_count(void* h) { return ((Bm25*)h)->postings.size(); }
uint32_t vo_bm25_search(void* h, const uint32_t* q, uint64_t nq, uint32_t k, uint64_t* out_ids, float* out_score) {
    auto r = ((Bm25*)h)->search(q, nq, k);
    for (size_t i = 0; i < r.size(); ++i) {
        out_ids[i] = r[i].first;
        out_score[i] = r[i].second;
    }
    return (uint32_t)r.size();
}
void vo_bm25_search_batch(void* h, const uint32_t* q_ptr, const uint32_t* q_terms, uint64_t nq, uint32_t k, int threads,
                          uint64_t* out_ids, float* out_score, uint32_t* out_counts) {
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        for (;;) {
            uint64_t i = next.fetch_add(1);
            if (i >= nq) break;
            out_counts[i] = vo_bm25_search(h, q_terms + q_ptr[i], q_ptr[i + 1] - q_ptr[i], k, out_ids + i * (size_t)k,
                                           out_score + i * (size_t)k);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < std::max(1, threads); ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
}

// ---- hybrid RRF (collection/search/text.rs:133-180) ----
// vec_ids / txt_ids are the two ranked lists (already best-first).  Output:
// top-k by (score, id) keeping the largest, sorted score-descending; equal scores
// canonicalised id-descending (the min-heap keeps the larger id; final stable
// sort order among equals is heap-internal -> canonical form used for parity).
uint32_t vo_rrf_hybrid(const uint64_t* vec_ids, uint32_t nv, const uint64_t* txt_ids, uint32_t nt, float vector_weight,
                       uint32_t k, uint64_t* out_ids, float* out_score) {
    float w = vector_weight;
    if (w < 0.0f) w = 0.0f;
    if (w > 1.0f) w = 1.0f;
    float tw = 1.0f - w;
    std::vector<std::pair<uint64_t, float>> fused;  // insertion-ordered map
    auto add = [&](uint64_t id, float s) {
        for (auto& p : fused)
            if (p.first == id) {
                p.second += s;
                return;
            }
        fused.push_back({id, 0.0f + s});
    };
    for (uint32_t r = 0; r < nv; ++r) add(vec_ids[r], w / ((float)r + 60.0f));
    for (uint32_t r = 0; r < nt; ++r) add(txt_ids[r], tw / ((float)r + 60.0f));
    // search/mod.rs:24-42 OrderedFloat = partial_cmp (no NaN here), tuple (score, id)
    std::sort(fused.begin(), fused.end(), [](const auto& a, const auto& b) {
        if (a.second != b.second) return a.second > b.second;
        return a.first > b.first;
    });
    uint32_t m = (uint32_t)std::min<size_t>(k, fused.size());
    for (uint32_t i = 0; i < m; ++i) {
        out_ids[i] = fused[i].first;
        out_score[i] = fused[i].second;
    }
    return m;
}

// ---- FusionStrategy (fusion/strategy.rs:138-300) ----
// lists: concatenated (id, score) with list_ptr offsets.  strategy: 0 avg, 1 max,
// 2 rrf(k_const), 3 weighted(avg_w, max_w, hit_w).  Output sorted score-desc;
// ties canonicalised id-ascending (HashMap iteration order is unspecified).
uint32_t vo_fuse(int strategy, const uint32_t* list_ptr, uint32_t n_lists, const uint64_t* ids, const float* scores,
                 uint32_t rrf_k, float avg_w, float max_w, float hit_w, uint32_t cap, uint64_t* out_ids,
                 float* out_score) {
    if (n_lists == 0) return 0;
    bool any = false;
    for (uint32_t l = 0; l < n_lists; ++l) any |= list_ptr[l + 1] > list_ptr[l];
    if (!any) return 0;
    std::vector<uint64_t> order;  // first-seen order of docs
    std::unordered_map<uint64_t, std::vector<float>> per_doc;
    std::unordered_map<uint64_t, float> acc;
    float kf = (float)rrf_k;
    for (uint32_t l = 0; l < n_lists; ++l) {
        std::vector<uint64_t> lorder;
        std::unordered_map<uint64_t, float> best;
        std::unordered_map<uint64_t, uint32_t> first_rank;
        for (uint32_t i = list_ptr[l]; i < list_ptr[l + 1]; ++i) {
            uint64_t id = ids[i];
            uint32_t rank = i - list_ptr[l];
            auto it = best.find(id);
            if (it == best.end()) {
                best[id] = scores[i];
                first_rank[id] = rank;
                lorder.push_back(id);
            } else {
                it->second = std::fmax(it->second, scores[i]);  // f32::max
            }
        }
        for (uint64_t id : lorder) {
            if (!per_doc.count(id) && !acc.count(id)) order.push_back(id);
            if (strategy == 2) {
                float s = 1.0f / (kf + (float)(first_rank[id] + 1));
                acc[id] = (acc.count(id) ? acc[id] : 0.0f) + s;
            } else if (strategy == 1) {
                acc[id] = acc.count(id) ? std::fmax(acc[id], best[id]) : best[id];
            } else {
                per_doc[id].push_back(best[id]);
            }
        }
    }
    std::vector<std::pair<uint64_t, float>> fused;
    for (uint64_t id : order) {
        float v;
        if (strategy == 1 || strategy == 2) {
            v = acc[id];
        } else {
            auto& sc = per_doc[id];
            float sum = 0.0f;
            for (float s : sc) sum += s;
            float avg = sum / (float)sc.size();
            if (strategy == 0) {
                v = avg;
            } else {
                float mx = -std::numeric_limits<float>::infinity();
                for (float s : sc) mx = std::fmax(mx, s);
                float hit = (float)sc.size() / (float)n_lists;
                v = avg_w * avg + max_w * mx + hit_w * hit;
            }
        }
        fused.push_back({id, v});
    }
    std::sort(fused.begin(), fused.end(), [](const auto& a, const auto& b) {
        int c = total_cmp(b.second, a.second);
        if (c != 0) return c < 0;
        return a.first < b.first;
    });
    uint32_t m = (uint32_t)std::min<size_t>(cap, fused.size());
    for (uint32_t i = 0; i < m; ++i) {
        out_ids[i] = fused[i].first;
        out_score[i] = fused[i].second;
    }
    return m;
}

// ---- SQ8 dual precision -------------------------------------------------------------------------
void* vo_sq8_train(const float* v, uint64_t n_train, uint32_t dim) {
    Sq8* s = new Sq8();
    s->train(v, n_train, dim);
    return s;
}
void vo_sq8_free(void* s) { delete (Sq8*)s; }
void vo_sq8_params(void* s, float* mn, float* scale, float* inv) {
    Sq8* q = (Sq8*)s;
    std::memcpy(mn, q->mn.data(), q->dim * 4);
    std::memcpy(scale, q->scale.data(), q->dim * 4);
    std::memcpy(inv, q->inv.data(), q->dim * 4);
}
void vo_sq8_quantize(void* s, const float* v, uint64_t n, uint8_t* out) {
    Sq8* q = (Sq8*)s;
    for (uint64_t i = 0; i < n; ++i) q->quantize(v + i * (size_t)q->dim, out + i * (size_t)q->dim);
}
// appends n vectors to the store (QuantizedVectorStore::push)
void vo_sq8_push(void* s, const float* v, uint64_t n) {
    Sq8* q = (Sq8*)s;
    for (uint64_t i = 0; i < n; ++i) q->push(v + i * (size_t)q->dim);
}
uint64_t vo_sq8_len(void* s) { return ((Sq8*)s)->count; }
const uint8_t* vo_sq8_codes(void* s) { return ((Sq8*)s)->codes.data(); }
uint32_t vo_sq8_distance_quantized(const uint8_t* a, const uint8_t* b, uint32_t dim) { return Sq8::dist_q(a, b, dim); }
float vo_sq8_distance_asymmetric(void* s, const float* q, const uint8_t* c) { return ((Sq8*)s)->dist_asym(q, c); }

// search_with_config on a trained index (dual_precision.rs:263-325).  stats layout as vo_hnsw_search.
uint32_t vo_dual_search_int8(void* h, void* s, const float* q, uint32_t k, uint32_t ef_search, uint32_t oversampling,
                             int order_mode, uint64_t* out_ids, float* out_dist, uint64_t* stats) {
    SearchStats st;
    std::vector<DN> r = dual_search_int8(*(Hnsw*)h, *(Sq8*)s, q, k, ef_search, oversampling, order_mode, &st);
    for (size_t i = 0; i < r.size(); ++i) {
        out_ids[i] = r[i].n;
        out_dist[i] = r[i].d;
    }
    if (stats) {
        stats[0] = st.ndc0;
        stats[1] = st.hops0;
        stats[2] = st.ndc_up;
        stats[3] = st.hops_up;
        stats[4] = st.tie_at_k;
        stats[5] = st.adj0;
    }
    return (uint32_t)r.size();
}
void vo_dual_search_int8_batch(void* h, void* s, const float* q, uint64_t nq, uint32_t k, uint32_t ef_search,
                               uint32_t oversampling, int order_mode, int threads, uint64_t* out_ids, float* out_dist,
                               uint32_t* out_counts, uint64_t* stats) {
    Hnsw* g = (Hnsw*)h;
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        for (;;) {
            uint64_t i = next.fetch_add(1);
            if (i >= nq) break;
            out_counts[i] = vo_dual_search_int8(g, s, q + i * (size_t)g->dim, k, ef_search, oversampling, order_mode,
                                                out_ids + i * (size_t)k, out_dist + i * (size_t)k,
                                                stats ? stats + i * 6 : nullptr);
            for (uint32_t j = out_counts[i]; j < k; ++j) {
                out_ids[i * (size_t)k + j] = UINT64_MAX;
                out_dist[i * (size_t)k + j] = std::numeric_limits<float>::quiet_NaN();
            }
        }
    };
    if (threads <= 1) {
        work();
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
}

}  // extern "C"

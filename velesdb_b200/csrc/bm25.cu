// bm25.cu -- batched BM25 posting scan: Bm25Index::search + score_document_fast
// (index/bm25.rs:269-376) on an immutable device snapshot of the inverted index.
//
// Reference arithmetic kept exactly (all f32):
//   avgdl    = total_len as f32 / doc_count as f32                              bm25.rs:281
//   idf(t)   = ln((N - df + 0.5) / (df + 0.5) + 1.0), 0 when df == 0            bm25.rs:297-306 (host, logf)
//   len_norm = 1.0 - b + b * doc_len / avgdl                                    bm25.rs:357-358
//   score(d) = sum over query tokens IN QUERY ORDER (duplicates again) of
//              idf * (tf * (k1 + 1.0)) / (tf + k1 * len_norm)                   bm25.rs:360-375
//   keep score > 0; order by score descending (total order), ties by doc id ascending (the reference's
//   tie order is hash/roaring iteration + select_nth_unstable, i.e. unspecified -- SURVEY 8a a19).
//
// Layout in HBM: postings CSR by term, doc ids ascending within a term.  post_dc = (doc u32, contribution f32), the
// whole per-posting term of the sum precomputed per snapshot with the reference's operation order (the snapshot is
// immutable: avgdl and idf are constants of it); (doc, tf, den) in three arrays for the fallback; two skip tables:
// skip[term][r] = first posting with doc >= r * kRange (u64) and skipf[term][s] = first posting with doc >= s * kFine,
// relative to the term's list (u32, built when it fits), so a worker reads exactly its slice of every posting list,
// coalesced, with no search.
//
// Query path, k <= kMultiK and <= kQueryTerms tokens:
//   bm25_sub_kernel   (default) one WARP per (query, span of kFine-document sub-ranges), private 4 KB accumulator, no
//                     block barriers; (query, part) work items, the last part of a query merges the parts.
//   bm25_flat_kernel  (no fine table) one CTA per (query, part of the kRange-document ranges), warp-step mapping,
//                     pipelined rounds, a block barrier per token.
//   bm25_query_kernel (round 1) one CTA per query, chunk-at-a-time walk; and three measured experiments.
// Every document receives its token contributions in query order in all of them (a document occurs at most once per
// posting list: no intra-token conflicts, no atomics), so the f32 sums keep the reference's bits.
// Fallback (larger k / longer queries): bm25_range_kernel, one CTA per (doc range, query), writes per-range lists and
// bm25_merge_kernel merges them.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <memory>
#include <mutex>

#include "index.hpp"

namespace veles {
constexpr uint32_t kRange = 7168;  // docs per range: 28 KB of f32 accumulators -> 7 query CTAs per SM (1036 >= 1024 resident)
constexpr uint32_t kTermChunk = 32; // query tokens whose metadata is staged at once
constexpr uint32_t kMultiK = 128;   // up to this k every warp of the CTA keeps its own top-k list
constexpr uint32_t kFine = 1024;    // docs per sub-range of the fine skip table: 4 KB of f32 accumulators per WARP (bm25_sub_kernel)
}

struct veles_bm25 {
    uint32_t n_terms = 0, n_doc_slots = 0, n_ranges = 0;
    uint64_t doc_count = 0, total_len = 0, n_postings = 0;
    float k1 = 1.2f, b = 0.75f, avgdl = 0.0f;
    uint32_t n_fine = 0;  // sub-ranges of kFine documents; 0 = no fine table (it would not fit: see veles_bm25_from_csr)
    veles::DevBuf term_ptr, post_doc, post_tf, post_den, post_dc, doc_len, idf, skip, skipf;
    mutable std::mutex mu;
    mutable veles::DevBuf q_ptr_d, q_terms_d, partial_d, out_doc_d, out_score_d, out_cnt_d;
    // veles_hybrid_search_batch: the text leg runs on `side` next to the vector leg on the caller's stream
    mutable cudaStream_t side = nullptr;
    mutable cudaEvent_t side_done = nullptr, side_fork = nullptr;
    mutable veles::DevBuf hy_ids_d, hy_score_d, hy_cnt_d;
    veles_bm25() = default;
    veles_bm25(const veles_bm25&) = delete;
    veles_bm25& operator=(const veles_bm25&) = delete;
    ~veles_bm25() {
        if (side_done) cudaEventDestroy(side_done);
        if (side_fork) cudaEventDestroy(side_fork);
        if (side) cudaStreamDestroy(side);
    }
};

namespace veles {

struct Bm25View {
    const uint32_t* post_doc;
    const uint32_t* post_tf;
    const float* post_den;  // tf + k1 * (1 - b + b * doc_len / avgdl), the denominator of bm25.rs:371-372 per posting
    const uint2* post_dc;   // (doc, contribution bits): idf * (tf * (k1 + 1)) / den, a constant of the snapshot per posting
    const uint32_t* doc_len;
    const float* idf;
    const uint64_t* skip;  // n_terms x (n_ranges + 1)
    const uint64_t* term_ptr;  // n_terms + 1
    const uint32_t* skipf;     // n_terms x (n_fine + 1): first posting of the term, relative to term_ptr, with doc >= s * kFine
    uint32_t n_fine;
    uint32_t n_terms, n_ranges, n_doc_slots;
    float k1, b, avgdl;
};

// The range's k best (score desc, doc asc) as ascending keys (~order(score) << 32 | doc).  Every warp scans
// its own slice of the accumulators into a register-resident sorted list (RegTopK), spills it to shared
// memory, and warp 0 merges the lists.  k > 128 falls back to one warp with a shared-memory list.
template <int R>
__device__ __forceinline__ void bm25_topk_regs(const float* acc, uint64_t* res, bool touched, uint32_t base_doc, uint32_t k,
                                               uint64_t* out) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    RegTopK<R> top;
    top.init(k, lane);
    if (touched) {
        const float4* acc4 = reinterpret_cast<const float4*>(acc);
        const uint32_t per = kRange / nwarps, i_begin = warp * per;
        for (uint32_t i0 = i_begin; i0 < i_begin + per; i0 += 128) {
            const float4 s4 = acc4[(i0 >> 2) + lane];
            const bool any = s4.x > 0.0f || s4.y > 0.0f || s4.z > 0.0f || s4.w > 0.0f;
            if (!__ballot_sync(FULL_MASK, any)) continue;  // most accumulators are zero
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                uint64_t key = ~0ull;
                if (sv[e] > 0.0f) key = ((uint64_t)(~ord_key(sv[e])) << 32) | (base_doc + i0 + lane * 4 + e);
                top.offer(key);
            }
        }
    }
    top.store(res + (size_t)warp * k, k);
    __syncthreads();
    if (warp != 0) return;
    for (uint32_t w = 1; w < nwarps; ++w) {
        const uint64_t* other = res + (size_t)w * k;
        for (uint32_t j0 = 0; j0 < k; j0 += 32) {
            const uint32_t j = j0 + lane;
            top.offer(j < k ? other[j] : ~0ull);
        }
    }
    top.store(out, k);
}

__device__ __forceinline__ void bm25_range_topk(const float* acc, uint64_t* res, bool touched, uint32_t base_doc, uint32_t k,
                                                uint64_t* out) {
    if (k <= 32) return bm25_topk_regs<1>(acc, res, touched, base_doc, k, out);
    if (k <= 64) return bm25_topk_regs<2>(acc, res, touched, base_doc, k, out);
    if (k <= kMultiK) return bm25_topk_regs<4>(acc, res, touched, base_doc, k, out);
    if (threadIdx.x >= 32) return;
    const uint32_t lane = threadIdx.x;
    uint32_t len = 0;
    uint64_t worst = ~0ull;
    if (touched) {
        for (uint32_t i0 = 0; i0 < kRange; i0 += 32) {
            const float s = acc[i0 + lane];
            uint64_t key = ~0ull;
            if (s > 0.0f) key = ((uint64_t)(~ord_key(s)) << 32) | (base_doc + i0 + lane);
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
    }
    __syncwarp();
    for (uint32_t j = lane; j < k; j += 32) out[j] = j < len ? res[j] : ~0ull;
}

__global__ void __launch_bounds__(256) bm25_range_kernel(Bm25View v, const uint32_t* __restrict__ q_ptr,
                                                         const uint32_t* __restrict__ q_terms, uint32_t k,
                                                         uint64_t* __restrict__ partial) {
    extern __shared__ __align__(16) uint8_t bm_smem[];
    float* acc = reinterpret_cast<float*>(bm_smem);
    uint64_t* res = reinterpret_cast<uint64_t*>(bm_smem + kRange * 4);
    const uint32_t r = blockIdx.x, q = blockIdx.y;
    const uint32_t base_doc = r * kRange;
    for (uint32_t i = threadIdx.x; i < kRange; i += blockDim.x) acc[i] = 0.0f;
    __syncthreads();
    const uint32_t t0 = q_ptr[q], t1 = q_ptr[q + 1];
    const float k1p1 = __fadd_rn(v.k1, 1.0f);
    bool touched = false;
    // term metadata for up to kTermChunk query tokens at a time, fetched by parallel threads so the
    // dependent global loads (term id -> skip entries) are paid once per chunk, not once per term
    __shared__ uint64_t s_lo[kTermChunk], s_hi[kTermChunk];
    __shared__ float s_idf[kTermChunk];
    for (uint32_t c0 = t0; c0 < t1; c0 += kTermChunk) {
        const uint32_t nt = min(kTermChunk, t1 - c0);
        if (threadIdx.x < nt) {
            const uint32_t term = q_terms[c0 + threadIdx.x];
            uint64_t lo = 0, hi = 0;
            float idf = 0.0f;
            if (term < v.n_terms) {  // unknown term: df = 0, contributes nothing
                lo = v.skip[(size_t)term * (v.n_ranges + 1) + r];
                hi = v.skip[(size_t)term * (v.n_ranges + 1) + r + 1];
                idf = v.idf[term];
            }
            s_lo[threadIdx.x] = lo;
            s_hi[threadIdx.x] = hi;
            s_idf[threadIdx.x] = idf;
        }
        __syncthreads();
        // Walk the (term, 256-posting chunk) sequence in query order.  Chunks of one term touch distinct
        // documents, so a block barrier is only needed when the term changes; the next chunk's postings are
        // requested before the current chunk is applied.
        uint32_t g = 0;
        while (g < nt && s_hi[g] == s_lo[g]) ++g;
        uint64_t base = g < nt ? s_lo[g] : 0;
        uint32_t d_cur = VELES_INVALID_ID, tf_cur = 0;
        float den_cur = 1.0f;
        if (g < nt && base + threadIdx.x < s_hi[g]) {
            d_cur = v.post_doc[base + threadIdx.x];
            tf_cur = v.post_tf[base + threadIdx.x];
            den_cur = v.post_den[base + threadIdx.x];
        }
        while (g < nt) {
            // next chunk (uniform across the block)
            uint32_t g2 = g;
            uint64_t base2 = base + blockDim.x;
            if (base2 >= s_hi[g2]) {
                ++g2;
                while (g2 < nt && s_hi[g2] == s_lo[g2]) ++g2;
                base2 = g2 < nt ? s_lo[g2] : 0;
            }
            uint32_t d_nxt = VELES_INVALID_ID, tf_nxt = 0;
            float den_nxt = 1.0f;
            if (g2 < nt && base2 + threadIdx.x < s_hi[g2]) {
                d_nxt = v.post_doc[base2 + threadIdx.x];
                tf_nxt = v.post_tf[base2 + threadIdx.x];
                den_nxt = v.post_den[base2 + threadIdx.x];
            }
            if (d_cur != VELES_INVALID_ID) {
                const float num = __fmul_rn((float)tf_cur, k1p1);
                const float contrib = __fdiv_rn(__fmul_rn(s_idf[g], num), den_cur);
                acc[d_cur - base_doc] = __fadd_rn(acc[d_cur - base_doc], contrib);
            }
            touched = true;
            if (g2 != g) __syncthreads();  // next term's contributions come after this term's, per document
            g = g2;
            base = base2;
            d_cur = d_nxt;
            tf_cur = tf_nxt;
            den_cur = den_nxt;
        }
        __syncthreads();
    }
    bm25_range_topk(acc, res, touched, base_doc, k, partial + ((size_t)q * v.n_ranges + r) * k);
}

// ---- one CTA per query (k <= kMultiK) ------------------------------------------------------------------
// The CTA walks the doc-id ranges of its query one after another.  Per range: zero the dense accumulator,
// apply the query's terms in order (prefetched (term, chunk) walk, as above), then every warp scans its own
// 1/8 slice of the accumulator and offers the positive scores to a register-resident top-k list that it
// keeps for the whole query.  After a few ranges the list's k-th key prunes almost everything with a single
// compare, so the scan costs ~8 loads per lane per range and inserts become rare; no per-range lists, no
// merge kernel.  The skip entries of the next range are requested one range ahead.
constexpr uint32_t kQueryTerms = 32;  // query tokens handled per pass over the ranges
constexpr int kDepth = 2;             // posting chunks in flight (the one being applied + one ahead)
template <int R>
__global__ void __launch_bounds__(256, 5) bm25_query_kernel(Bm25View v, const uint32_t* __restrict__ q_ptr,
                                                            const uint32_t* __restrict__ q_terms, uint32_t nq, uint32_t k,
                                                            uint32_t* __restrict__ out_doc, float* __restrict__ out_score,
                                                            uint32_t* __restrict__ out_cnt, uint32_t* __restrict__ work) {
    extern __shared__ __align__(16) uint8_t bq_smem[];
    float* acc = reinterpret_cast<float*>(bq_smem);
    uint64_t* lists = reinterpret_cast<uint64_t*>(bq_smem + kRange * 4);  // 8 x k keys for the final merge
    __shared__ uint64_t s_lo[kQueryTerms], s_hi[kQueryTerms];
    __shared__ float s_idf[kQueryTerms];
    __shared__ uint32_t s_q;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const float k1p1 = __fadd_rn(v.k1, 1.0f);
    const float4* acc4 = reinterpret_cast<const float4*>(acc);
    for (;;) {  // persistent: queries are handed out by a global counter
        if (threadIdx.x == 0) s_q = atomicAdd(work, 1u);
        __syncthreads();
        const uint32_t q = s_q;
        if (q >= nq) return;
        const uint32_t t0 = q_ptr[q], t1 = q_ptr[q + 1];
        const uint32_t nt = min(kQueryTerms, t1 - t0);  // host guarantees t1 - t0 <= kQueryTerms on this path
        RegTopK<R> top;
        top.init(k, lane);
        // per-thread term metadata (threads < nt): skip row of the term, boundaries of the current range
        uint64_t my_lo = 0, my_next = 0;
        const uint64_t* my_skip = nullptr;
        if (threadIdx.x < nt) {
            const uint32_t term = q_terms[t0 + threadIdx.x];
            float idf = 0.0f;
            if (term < v.n_terms) {
                idf = v.idf[term];
                my_skip = v.skip + (size_t)term * (v.n_ranges + 1);
                my_lo = my_skip[0];
                my_next = my_skip[1];
            }
            s_idf[threadIdx.x] = idf;
        }
        for (uint32_t r = 0; r < v.n_ranges; ++r) {
            const uint32_t base_doc = r * kRange;
            if (threadIdx.x < nt) {
                s_lo[threadIdx.x] = my_lo;
                s_hi[threadIdx.x] = my_next;
                my_lo = my_next;
                if (my_skip && r + 2 <= v.n_ranges) my_next = my_skip[r + 2];  // boundary after the next range
            }
            for (uint32_t i = threadIdx.x; i < kRange / 4; i += blockDim.x)
                reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncthreads();
            // (term, 256-posting chunk) walk in query order with kDepth chunks in flight.  Chunks of one term
            // touch distinct documents: a block barrier is only needed when the term changes.
            uint32_t cg = 0;  // cursor: next chunk to request
            while (cg < nt && s_hi[cg] == s_lo[cg]) ++cg;
            uint64_t cbase = cg < nt ? s_lo[cg] : 0;
            uint32_t qg[kDepth], qd[kDepth];
            float qc[kDepth];  // the posting's whole contribution, precomputed per snapshot (Bm25View::post_dc)
            auto request = [&](int slot) {
                qg[slot] = cg;
                qd[slot] = VELES_INVALID_ID;
                qc[slot] = 0.0f;
                if (cg < nt) {
                    const uint64_t p = cbase + threadIdx.x;
                    if (p < s_hi[cg]) {
                        const uint2 dc = v.post_dc[p];
                        qd[slot] = dc.x;
                        qc[slot] = __uint_as_float(dc.y);
                    }
                    cbase += blockDim.x;
                    if (cbase >= s_hi[cg]) {
                        ++cg;
                        while (cg < nt && s_hi[cg] == s_lo[cg]) ++cg;
                        cbase = cg < nt ? s_lo[cg] : 0;
                    }
                }
            };
#pragma unroll
            for (int i = 0; i < kDepth; ++i) request(i);
            bool touched = false;
            while (qg[0] < nt) {
                if (qd[0] != VELES_INVALID_ID) acc[qd[0] - base_doc] = __fadd_rn(acc[qd[0] - base_doc], qc[0]);
                touched = true;
                if (qg[1] != qg[0]) __syncthreads();  // next term's contributions come after this term's
#pragma unroll
                for (int i = 0; i + 1 < kDepth; ++i) {
                    qg[i] = qg[i + 1];
                    qd[i] = qd[i + 1];
                    qc[i] = qc[i + 1];
                }
                request(kDepth - 1);
            }
            // scan: this warp's slice of the range
            if (touched) {
                const uint32_t per = kRange / nwarps, i_begin = warp * per;
                for (uint32_t i0 = i_begin; i0 < i_begin + per; i0 += 128) {
                    const float4 s4 = acc4[(i0 >> 2) + lane];
                    // once the list is full only scores >= its k-th score can enter (ties go by document id, so
                    // equality stays in): one compare per score skips nearly every step after the first ranges
                    const float thr = top.worst == ~0ull ? 0.0f : ord_unkey(~(uint32_t)(top.worst >> 32));
                    const bool any = (s4.x > 0.0f && s4.x >= thr) || (s4.y > 0.0f && s4.y >= thr) ||
                                     (s4.z > 0.0f && s4.z >= thr) || (s4.w > 0.0f && s4.w >= thr);
                    if (!__ballot_sync(FULL_MASK, any)) continue;
                    const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        uint64_t key = ~0ull;
                        if (sv[e] > 0.0f) key = ((uint64_t)(~ord_key(sv[e])) << 32) | (base_doc + i0 + lane * 4 + e);
                        top.offer(key);
                    }
                }
            }
            __syncthreads();  // the accumulator is rewritten by the next range
        }
        // merge the eight per-warp lists
        top.store(lists + (size_t)warp * k, k);
        __syncthreads();
        if (warp == 0) {
            for (uint32_t w = 1; w < nwarps; ++w) {
                const uint64_t* other = lists + (size_t)w * k;
                for (uint32_t j0 = 0; j0 < k; j0 += 32) {
                    const uint32_t j = j0 + lane;
                    top.offer(j < k ? other[j] : ~0ull);
                }
            }
            top.store(lists, k);
            __syncwarp();
            uint32_t len = 0;
            for (uint32_t j0 = 0; j0 < k; j0 += 32) {
                const uint32_t j = j0 + lane;
                if (j < k) {
                    const uint64_t key = lists[j];
                    uint32_t doc = VELES_INVALID_ID;
                    float sc = __uint_as_float(0x7fc00000u);
                    if (key != ~0ull) {
                        doc = (uint32_t)key;
                        sc = ord_unkey(~(uint32_t)(key >> 32));
                        ++len;
                    }
                    out_doc[(size_t)q * k + j] = doc;
                    out_score[(size_t)q * k + j] = sc;
                }
            }
            len = __reduce_add_sync(FULL_MASK, len);
            if (lane == 0) out_cnt[q] = len;
        }
        __syncthreads();  // `lists` and s_q are reused by the next query
    }
}

// ---- one CTA per query, warp-sliced (k <= kMultiK): no barrier between terms ---------------------------------------------
// The walk above spends ~1300 instructions per warp and range on ~90 postings per warp: the terms of a range are applied
// one after another with a block barrier each, because all eight warps share the range's accumulator.  Here every warp
// OWNS a 896-document slice of the range instead: it reads every posting of the range (8 bytes: doc + the precomputed
// contribution) and adds the ones that fall into its slice.  A warp applies the terms in program order, so query-token
// order is kept by construction and the only synchronisation inside a range is a warp barrier per term; the eight warps
// meet once per range (so that they read the same posting lines while those are in L1).  The next term's postings are
// requested before the current term's are applied.
constexpr uint32_t kSliceU = 8;  // postings per lane per request: 256 per warp, covers a typical (term, range) slice in one go
template <int R>
__global__ void __launch_bounds__(256, 5) bm25_slice_kernel(Bm25View v, const uint32_t* __restrict__ q_ptr,
                                                            const uint32_t* __restrict__ q_terms, uint32_t nq, uint32_t k,
                                                            uint32_t* __restrict__ out_doc, float* __restrict__ out_score,
                                                            uint32_t* __restrict__ out_cnt, uint32_t* __restrict__ work) {
    extern __shared__ __align__(16) uint8_t bq_smem[];
    float* acc = reinterpret_cast<float*>(bq_smem);
    uint64_t* lists = reinterpret_cast<uint64_t*>(bq_smem + kRange * 4);  // 8 x k keys for the final merge
    __shared__ uint32_t s_q;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    constexpr uint32_t kSub = kRange / 8;  // documents per warp slice (blockDim.x == 256)
    float* mine = acc + warp * kSub;
    for (;;) {  // persistent: queries are handed out by a global counter
        if (threadIdx.x == 0) s_q = atomicAdd(work, 1u);
        __syncthreads();
        const uint32_t q = s_q;
        if (q >= nq) return;
        const uint32_t t0 = q_ptr[q], t1 = q_ptr[q + 1];
        const uint32_t nt = min(kQueryTerms, t1 - t0);  // host guarantees t1 - t0 <= kQueryTerms on this path
        RegTopK<R> top;
        top.init(k, lane);
        // lane t < nt of EVERY warp tracks term t: its skip row and the posting range of the current doc range
        uint64_t my_lo = 0, my_next = 0;
        const uint64_t* my_skip = nullptr;
        if (lane < nt) {
            const uint32_t term = q_terms[t0 + lane];
            if (term < v.n_terms) {
                my_skip = v.skip + (size_t)term * (v.n_ranges + 1);
                my_lo = my_skip[0];
                my_next = my_skip[1];
            }
        }
        for (uint32_t r = 0; r < v.n_ranges; ++r) {
            const uint32_t slice_lo = r * kRange + warp * kSub;  // first document of this warp's slice
            const uint64_t lo = my_lo, hi = my_next;
            my_lo = my_next;
            if (my_skip && r + 2 <= v.n_ranges) my_next = my_skip[r + 2];  // boundary after the next range
            const uint32_t any = __ballot_sync(FULL_MASK, hi > lo);
#pragma unroll
            for (uint32_t i = 0; i < kSub / 128; ++i) reinterpret_cast<float4*>(mine)[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
            if (any) {
                // requests run one batch ahead of the applies: (term, offset) cursor over the non-empty slices
                uint2 cur[kSliceU], nxt[kSliceU];
                uint32_t ct = __ffs(any) - 1;  // current term
                uint64_t cbeg = __shfl_sync(FULL_MASK, lo, ct), cend = __shfl_sync(FULL_MASK, hi, ct);
                auto request = [&](uint2 (&dst)[kSliceU], uint64_t beg, uint64_t end) {
#pragma unroll
                    for (uint32_t u = 0; u < kSliceU; ++u) {
                        const uint64_t p = beg + u * 32 + lane;
                        dst[u] = p < end ? v.post_dc[p] : make_uint2(0xffffffffu, 0u);
                    }
                };
                request(cur, cbeg, cend);
                for (;;) {
                    // the batch after this one: same term if it has more postings, else the next non-empty term
                    uint32_t nt_term = ct;
                    uint64_t nbeg = cbeg + kSliceU * 32, nend = cend;
                    bool more = nbeg < cend;
                    if (!more) {
                        const uint32_t rest = any & ~((2u << ct) - 1u);
                        if (rest) {
                            nt_term = __ffs(rest) - 1;
                            nbeg = __shfl_sync(FULL_MASK, lo, nt_term);
                            nend = __shfl_sync(FULL_MASK, hi, nt_term);
                            more = true;
                        }
                    }
                    if (more) request(nxt, nbeg, nend);
#pragma unroll
                    for (uint32_t u = 0; u < kSliceU; ++u) {
                        const uint32_t off = cur[u].x - slice_lo;  // unsigned: documents below the slice wrap to huge values
                        if (off < kSub) mine[off] = __fadd_rn(mine[off], __uint_as_float(cur[u].y));
                    }
                    if (!more) break;
                    if (nt_term != ct) __syncwarp();  // the next term's adds come after this term's (different lanes, same slot)
#pragma unroll
                    for (uint32_t u = 0; u < kSliceU; ++u) cur[u] = nxt[u];
                    ct = nt_term;
                    cbeg = nbeg;
                    cend = nend;
                }
                __syncwarp();
                // scan this warp's slice
                for (uint32_t i0 = 0; i0 < kSub; i0 += 128) {
                    const float4 s4 = reinterpret_cast<const float4*>(mine)[(i0 >> 2) + lane];
                    const float thr = top.worst == ~0ull ? 0.0f : ord_unkey(~(uint32_t)(top.worst >> 32));
                    const bool hit = (s4.x > 0.0f && s4.x >= thr) || (s4.y > 0.0f && s4.y >= thr) ||
                                     (s4.z > 0.0f && s4.z >= thr) || (s4.w > 0.0f && s4.w >= thr);
                    if (!__ballot_sync(FULL_MASK, hit)) continue;
                    const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        uint64_t key = ~0ull;
                        if (sv[e] > 0.0f) key = ((uint64_t)(~ord_key(sv[e])) << 32) | (slice_lo + i0 + lane * 4 + e);
                        top.offer(key);
                    }
                }
            }
            __syncthreads();  // keeps the eight warps on the same posting lines (L1), nothing else depends on it
        }
        // merge the eight per-warp lists
        top.store(lists + (size_t)warp * k, k);
        __syncthreads();
        if (warp == 0) {
            for (uint32_t w = 1; w < nwarps; ++w) {
                const uint64_t* other = lists + (size_t)w * k;
                for (uint32_t j0 = 0; j0 < k; j0 += 32) {
                    const uint32_t j = j0 + lane;
                    top.offer(j < k ? other[j] : ~0ull);
                }
            }
            top.store(lists, k);
            __syncwarp();
            uint32_t len = 0;
            for (uint32_t j0 = 0; j0 < k; j0 += 32) {
                const uint32_t j = j0 + lane;
                if (j < k) {
                    const uint64_t key = lists[j];
                    uint32_t doc = VELES_INVALID_ID;
                    float sc = __uint_as_float(0x7fc00000u);
                    if (key != ~0ull) {
                        doc = (uint32_t)key;
                        sc = ord_unkey(~(uint32_t)(key >> 32));
                        ++len;
                    }
                    out_doc[(size_t)q * k + j] = doc;
                    out_score[(size_t)q * k + j] = sc;
                }
            }
            len = __reduce_add_sync(FULL_MASK, len);
            if (lane == 0) out_cnt[q] = len;
        }
        __syncthreads();  // `lists` and s_q are reused by the next query
    }
}

// ---- one CTA per query, whole range in flight (k <= kMultiK) -------------------------------------------------------
// bm25_query_kernel above is latency bound: within a range it requests one 256-posting chunk of one term at a time, so
// a range of five terms costs five dependent DRAM round trips (ncu, round 1: IPC 2.4 but 37 % of the samples waiting
// on the posting loads; 2.3 ms per 1024 queries = 0.08 of the HBM peak).  Here ALL postings of a range -- every term's
// slice, flattened term-major -- are loaded at once, kPre per thread, before anything is applied: one round trip per
// range.  They are then applied term by term (a block barrier when the term changes), so every document still receives
// its contributions in query-token order and the f32 sums keep the reference's bits.
constexpr int kPre = 4;  // postings per thread per round: 1024 per CTA round (a range holds ~700 on the bench corpus)
template <int R>
__global__ void __launch_bounds__(256, 5) bm25_prefetch_kernel(Bm25View v, const uint32_t* __restrict__ q_ptr,
                                                               const uint32_t* __restrict__ q_terms, uint32_t nq, uint32_t k,
                                                               uint32_t* __restrict__ out_doc, float* __restrict__ out_score,
                                                               uint32_t* __restrict__ out_cnt, uint32_t* __restrict__ work) {
    extern __shared__ __align__(16) uint8_t bq_smem[];
    float* acc = reinterpret_cast<float*>(bq_smem);
    uint64_t* lists = reinterpret_cast<uint64_t*>(bq_smem + kRange * 4);  // 8 x k keys for the final merge
    __shared__ uint64_t s_lo[kQueryTerms];
    __shared__ uint32_t s_off[kQueryTerms + 1];  // flattened start of each term's slice within the range
    __shared__ float s_idf[kQueryTerms];
    __shared__ uint32_t s_q;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const float k1p1 = __fadd_rn(v.k1, 1.0f);
    const float4* acc4 = reinterpret_cast<const float4*>(acc);
    for (;;) {  // persistent: queries are handed out by a global counter
        if (threadIdx.x == 0) s_q = atomicAdd(work, 1u);
        __syncthreads();
        const uint32_t q = s_q;
        if (q >= nq) return;
        const uint32_t t0 = q_ptr[q], t1 = q_ptr[q + 1];
        const uint32_t nt = min(kQueryTerms, t1 - t0);  // host guarantees t1 - t0 <= kQueryTerms on this path
        RegTopK<R> top;
        top.init(k, lane);
        uint64_t my_lo = 0, my_next = 0;
        const uint64_t* my_skip = nullptr;
        if (threadIdx.x < nt) {
            const uint32_t term = q_terms[t0 + threadIdx.x];
            float idf = 0.0f;
            if (term < v.n_terms) {
                idf = v.idf[term];
                my_skip = v.skip + (size_t)term * (v.n_ranges + 1);
                my_lo = my_skip[0];
                my_next = my_skip[1];
            }
            s_idf[threadIdx.x] = idf;
        }
        for (uint32_t r = 0; r < v.n_ranges; ++r) {
            const uint32_t base_doc = r * kRange;
            // warp 0: slice boundaries of every term in this range and their flattened offsets (exclusive scan)
            if (warp == 0) {
                uint32_t cnt = 0;
                if (lane < nt) {
                    s_lo[lane] = my_lo;
                    cnt = (uint32_t)min(my_next - my_lo, (uint64_t)0x03ffffffu);
                    my_lo = my_next;
                    if (my_skip && r + 2 <= v.n_ranges) my_next = my_skip[r + 2];  // boundary after the next range
                }
                uint32_t incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(FULL_MASK, incl, o);
                    if ((int)lane >= o) incl += up;
                }
                if (lane < nt) s_off[lane] = incl - cnt;
                if (lane == 31) s_off[nt] = incl;  // lanes >= nt carry cnt = 0: lane 31 holds the total
            }
            for (uint32_t i = threadIdx.x; i < kRange / 4; i += blockDim.x)
                reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncthreads();
            const uint32_t total = s_off[nt];
            for (uint32_t round = 0; round < total; round += kPre * 256) {
                // ---- request: up to kPre postings per thread, all loads issued before anything is used ----
                uint32_t pt[kPre], pd[kPre], ptf[kPre];
                float pden[kPre];
#pragma unroll
                for (int j = 0; j < kPre; ++j) {
                    const uint32_t pos = round + j * 256 + threadIdx.x;
                    pt[j] = 0xffffffffu;
                    pd[j] = 0;
                    ptf[j] = 0;
                    pden[j] = 1.0f;
                    if (pos < total) {
                        uint32_t t = 0;
                        while (t + 1 < nt && s_off[t + 1] <= pos) ++t;
                        const uint64_t p = s_lo[t] + (pos - s_off[t]);
                        pt[j] = t;
                        pd[j] = v.post_doc[p];
                        ptf[j] = v.post_tf[p];
                        pden[j] = v.post_den[p];
                    }
                }
                // ---- apply term by term; positions are term-major, so round order = query-token order ----
                const uint32_t last_pos = min(total, round + kPre * 256) - 1;
                uint32_t t_first = 0, t_last = 0;
                while (t_first + 1 < nt && s_off[t_first + 1] <= round) ++t_first;
                t_last = t_first;
                while (t_last + 1 < nt && s_off[t_last + 1] <= last_pos) ++t_last;
                for (uint32_t t = t_first; t <= t_last; ++t) {
#pragma unroll
                    for (int j = 0; j < kPre; ++j) {
                        if (pt[j] == t) {
                            const float num = __fmul_rn((float)ptf[j], k1p1);
                            const float contrib = __fdiv_rn(__fmul_rn(s_idf[t], num), pden[j]);
                            acc[pd[j] - base_doc] = __fadd_rn(acc[pd[j] - base_doc], contrib);
                        }
                    }
                    __syncthreads();  // the next term's (or round's) contributions come after this term's
                }
            }
            // scan: this warp's slice of the range
            if (total > 0) {
                const uint32_t per = kRange / nwarps, i_begin = warp * per;
                for (uint32_t i0 = i_begin; i0 < i_begin + per; i0 += 128) {
                    const float4 s4 = acc4[(i0 >> 2) + lane];
                    const float thr = top.worst == ~0ull ? 0.0f : ord_unkey(~(uint32_t)(top.worst >> 32));
                    const bool any = (s4.x > 0.0f && s4.x >= thr) || (s4.y > 0.0f && s4.y >= thr) ||
                                     (s4.z > 0.0f && s4.z >= thr) || (s4.w > 0.0f && s4.w >= thr);
                    if (!__ballot_sync(FULL_MASK, any)) continue;
                    const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        uint64_t key = ~0ull;
                        if (sv[e] > 0.0f) key = ((uint64_t)(~ord_key(sv[e])) << 32) | (base_doc + i0 + lane * 4 + e);
                        top.offer(key);
                    }
                }
            }
            __syncthreads();  // the accumulator and s_off are rewritten by the next range
        }
        // merge the eight per-warp lists
        top.store(lists + (size_t)warp * k, k);
        __syncthreads();
        if (warp == 0) {
            for (uint32_t w = 1; w < nwarps; ++w) {
                const uint64_t* other = lists + (size_t)w * k;
                for (uint32_t j0 = 0; j0 < k; j0 += 32) {
                    const uint32_t j = j0 + lane;
                    top.offer(j < k ? other[j] : ~0ull);
                }
            }
            top.store(lists, k);
            __syncwarp();
            uint32_t len = 0;
            for (uint32_t j0 = 0; j0 < k; j0 += 32) {
                const uint32_t j = j0 + lane;
                if (j < k) {
                    const uint64_t key = lists[j];
                    uint32_t doc = VELES_INVALID_ID;
                    float sc = __uint_as_float(0x7fc00000u);
                    if (key != ~0ull) {
                        doc = (uint32_t)key;
                        sc = ord_unkey(~(uint32_t)(key >> 32));
                        ++len;
                    }
                    out_doc[(size_t)q * k + j] = doc;
                    out_score[(size_t)q * k + j] = sc;
                }
            }
            len = __reduce_add_sync(FULL_MASK, len);
            if (lane == 0) out_cnt[q] = len;
        }
        __syncthreads();  // `lists` and s_q are reused by the next query
    }
}

// ---- (query, part of the doc-id space) work items, warp-step mapping, rounds pipelined (k <= kMultiK) ---------------
// What the three kernels above have in common is ~1000+ issued instructions per warp and range for ~90 postings per
// warp: cursor bookkeeping per 256-posting chunk of ONE term (chunks 55 % full on the bench corpus), one DRAM round
// trip per chunk, a separate zeroing pass, and 1024 queries over 740 resident CTAs (a second wave 38 % full).  Here
//   * the postings of a range are cut into WARP-STEPS (32 consecutive postings of one term).  The steps of all terms,
//     term-major, are dealt round-robin to the eight warps, kFlatH per warp per ROUND (1024 postings); which term a
//     step belongs to is one ballot against the per-lane step prefix, uniform across the warp -- no per-thread search,
//     no per-term chunk loop.  A typical range is one round;
//   * the NEXT round's postings (next range included) are requested before the current round is applied, so the DRAM
//     round trip overlaps the apply + scan of the round in hand;
//   * a round is applied term by term, present terms only, with a block barrier after each (all eight warps share the
//     range's accumulator): every document still receives its contributions in query-token order, so the f32 sums
//     keep the reference's bits;
//   * the scan re-zeroes what it reads, so there is no zeroing pass;
//   * a work item is (query, 1/P of the ranges) with P chosen so that the batch is >= 6 waves of CTAs; the last item
//     of a query to finish (a ticket per query) merges the P sorted lists.  A single query therefore runs on up to
//     n_ranges CTAs instead of one.
constexpr int kFlatH = 4;                      // warp-steps a warp holds per round
constexpr uint32_t kFlatRound = 8 * kFlatH;    // warp-steps per round (8 warps)
struct __align__(16) FlatEnt {
    uint64_t at;    // address of the first (doc, contribution) pair of the term's slice of the range
    uint32_t len;   // postings in the slice
    uint32_t cum;   // warp-steps of the terms before this one
};
struct FlatHold {
    uint32_t d[kFlatH];  // document of this lane's posting of step h
    uint32_t c[kFlatH];  // bits of its contribution
    uint32_t tags;       // byte h: query-token index of step h, 0xff = no posting for this lane
    uint32_t mask;       // query tokens with steps in this round (uniform)
};
template <int R, int OCC>
__global__ void __launch_bounds__(256, OCC) bm25_flat_kernel(Bm25View v, const uint32_t* __restrict__ q_ptr,
                                                           const uint32_t* __restrict__ q_terms, uint32_t nq, uint32_t k,
                                                           uint32_t parts, uint64_t* __restrict__ partial,
                                                           uint32_t* __restrict__ tickets, uint32_t* __restrict__ qthr,
                                                           uint32_t* __restrict__ out_doc,
                                                           float* __restrict__ out_score, uint32_t* __restrict__ out_cnt,
                                                           uint32_t* __restrict__ work) {
    extern __shared__ __align__(16) uint8_t bq_smem[];
    float* acc = reinterpret_cast<float*>(bq_smem);
    uint64_t* lists = reinterpret_cast<uint64_t*>(bq_smem + kRange * 4);  // 8 x k keys + k merged
    __shared__ FlatEnt s_ent[4][kQueryTerms];  // ring over ranges: r, r+1 (being prefetched), r+2 (being written)
    __shared__ uint32_t s_total[4];            // warp-steps of the range
    // per token: skip[term][r] for a ring of four ranges.  Warp 0's set-up of range r reads entries r and r + 1 and
    // requests entry r + 3 with cp.async, so no thread ever waits for a skip entry (a plain load stored to shared
    // memory made warp 0 -- and with it the whole CTA at the next barrier -- wait a DRAM round trip per range)
    __shared__ uint64_t s_skip[4][kQueryTerms];
    __shared__ uint32_t s_term[kQueryTerms];
    __shared__ uint32_t s_item;
    // bits of the best k-th score any full list of this query is known to hold (this CTA's warps, and through qthr[q]
    // the query's other work items): k documents score at least this much, so anything below it is dropped unseen
    __shared__ uint32_t s_thr;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* acc4 = reinterpret_cast<float4*>(acc);
    for (uint32_t i = threadIdx.x; i < kRange / 4; i += 256) acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (;;) {  // persistent: work items are handed out by a global counter
        if (threadIdx.x == 0) s_item = atomicAdd(work, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= nq * parts) return;
        // part-major: by the time a query's later parts start, its earlier ones have published their k-th score (qthr)
        const uint32_t part = item / nq, q = item - part * nq;
        const uint32_t r_begin = (uint32_t)((uint64_t)part * v.n_ranges / parts);
        const uint32_t r_end = (uint32_t)((uint64_t)(part + 1) * v.n_ranges / parts);
        const uint32_t t0 = q_ptr[q], t1 = q_ptr[q + 1];
        const uint32_t nt = min(kQueryTerms, t1 - t0);  // host guarantees t1 - t0 <= kQueryTerms on this path
        RegTopK<R> top;
        top.init(k, lane);
        uint32_t thr_seen = 0;  // warp 0, lane 0: qthr[q] as last read (requested a range ahead)
        // warp 0, lane g < nt: the skip entries of query token g for the first two ranges; the third is on its way
        if (warp == 0) {
            if (lane == 0) {
                thr_seen = *reinterpret_cast<volatile uint32_t*>(qthr + q);
                s_thr = 0;
            }
            uint32_t term = 0xffffffffu;
            uint64_t lo = 0, hi = 0;
            if (lane < nt) term = q_terms[t0 + lane];
            if (term < v.n_terms) {  // unknown term: df = 0, contributes nothing
                const uint64_t* row = v.skip + (size_t)term * (v.n_ranges + 1);
                lo = row[r_begin];
                hi = row[r_begin + 1];
            }
            s_term[lane] = term;
            if (!(term < v.n_terms)) s_skip[(r_begin + 2) & 3][lane] = s_skip[(r_begin + 3) & 3][lane] = 0;  // never requested
            s_skip[r_begin & 3][lane] = lo;
            s_skip[(r_begin + 1) & 3][lane] = hi;
            if (term < v.n_terms && r_begin + 2 <= v.n_ranges)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(&s_skip[(r_begin + 2) & 3][lane])),
                             "l"(v.skip + (size_t)term * (v.n_ranges + 1) + r_begin + 2)
                             : "memory");
            cp_async_commit();
        }
        auto setup = [&](uint32_t r) {  // warp 0, all lanes: entries of range r into the ring
            cp_async_wait<0>();  // entry r + 1 was requested a whole range ago
            __syncwarp();
            const uint64_t my_lo = s_skip[r & 3][lane], my_hi = s_skip[(r + 1) & 3][lane];
            const uint32_t len = (uint32_t)(my_hi - my_lo);  // <= kRange
            const uint32_t steps = (len + 31) >> 5;
            uint32_t incl = steps;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(FULL_MASK, incl, o);
                if ((int)lane >= o) incl += up;
            }
            if (lane < nt) {
                FlatEnt e;
                e.at = reinterpret_cast<uint64_t>(v.post_dc + my_lo);
                e.len = len;
                e.cum = incl - steps;
                s_ent[r & 3][lane] = e;
            }
            if (lane == 31) s_total[r & 3] = incl;  // lanes >= nt carry 0 steps
            const uint32_t term = s_term[lane];
            if (term < v.n_terms && r + 3 <= v.n_ranges)  // slot (r + 3) & 3 held entry r - 1: free
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(&s_skip[(r + 3) & 3][lane])),
                             "l"(v.skip + (size_t)term * (v.n_ranges + 1) + r + 3)
                             : "memory");
            cp_async_commit();
        };
        if (warp == 0) {
            setup(r_begin);
            if (r_begin + 1 < r_end) setup(r_begin + 1);
        }
        __syncthreads();
        // requests the postings of round (r, j0) into h
        auto load_round = [&](FlatHold& h, uint32_t r, uint32_t j0, uint32_t total) {
            const FlatEnt* ent = s_ent[r & 3];
            uint32_t my_cum = 0xffffffffu, my_end = 0;
            if (lane < nt) {
                const uint4 e = *reinterpret_cast<const uint4*>(ent + lane);
                my_cum = e.w;
                my_end = e.w + ((e.z + 31) >> 5);
            }
            h.mask = __ballot_sync(FULL_MASK, my_cum < j0 + kFlatRound && my_end > j0 && my_end > my_cum);
            h.tags = 0xffffffffu;
#pragma unroll
            for (int i = 0; i < kFlatH; ++i) {
                h.d[i] = 0;
                h.c[i] = 0;
            }
            if (h.mask == 0) return;
#pragma unroll
            for (int i = 0; i < kFlatH; ++i) {
                if (j0 + i * 8 >= total) break;  // uniform: a half-empty round costs half
                const uint32_t j = j0 + i * 8 + warp;
                const uint32_t g = __popc(__ballot_sync(FULL_MASK, my_cum <= j)) - 1;  // mask != 0: token 0 has cum 0 <= j
                const uint4 e = *reinterpret_cast<const uint4*>(ent + g);
                const uint32_t off = ((j - e.w) << 5) + lane;
                const bool ok = off < e.z;  // also false for j past the last step
                const uint64_t src = (((uint64_t)e.y << 32) | e.x) + (uint64_t)off * 8;
                if (ok) h.tags ^= (g ^ 0xffu) << (8 * i);
                // the load lands in the hold registers themselves and is first looked at when the round is applied
                // (a C++ "if (ok) d = load" made the compiler wait for the data right here to select it)
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p ld.global.nc.v2.u32 {%0, %1}, [%2];\n\t}"
                    : "+r"(h.d[i]), "+r"(h.c[i])
                    : "l"(src), "r"((uint32_t)ok));
            }
        };
        bool touched = false;
        uint32_t r = r_begin, j0 = 0;
        // one round: request the next one into `nxt`, apply `cur`, finish the range if this was its last round
        auto round = [&](FlatHold& cur, FlatHold& nxt) {
            const uint32_t total = s_total[r & 3];
            uint32_t nr = r, nj0 = j0 + kFlatRound;
            const bool last = nj0 >= total;
            if (last) {
                nr = r + 1;
                nj0 = 0;
            }
            const uint32_t base_doc = r * kRange;
            // ptxas gives the two register sets' loads the SAME hardware scoreboard, so the first use of `cur` waits for
            // every load outstanding on it: look at `cur` (its accumulator slots) BEFORE requesting `nxt`, or each round
            // would wait for its own prefetch (ncu, first version: 13 % of the samples on exactly that).
            static_assert(kFlatH == 4, "the pin below names four slots");
            uint32_t slot[kFlatH];
#pragma unroll
            for (int i = 0; i < kFlatH; ++i) slot[i] = cur.d[i] - base_doc;
            asm volatile("" ::"r"(slot[0]), "r"(slot[1]), "r"(slot[2]), "r"(slot[3]) : "memory");
            if (nr < r_end) load_round(nxt, nr, nj0, last ? s_total[nr & 3] : total);
            for (uint32_t m = cur.mask; m; m &= m - 1) {
                const uint32_t g = __ffs(m) - 1;
                const uint32_t diff = cur.tags ^ (g * 0x01010101u);  // byte h is zero where step h belongs to token g
#pragma unroll
                for (int i = 0; i < kFlatH; ++i)
                    if ((diff & (0xffu << (8 * i))) == 0) acc[slot[i]] = __fadd_rn(acc[slot[i]], __uint_as_float(cur.c[i]));
                __syncthreads();  // the next token's (or round's) contributions come after this token's, per document
            }
            touched |= cur.mask != 0;
            if (last) {
                if (warp == 0 && r + 2 < r_end) setup(r + 2);
                if (touched) {  // scan this warp's 896 documents; what is read is zeroed for the next range
                    // own list: a warp meets its documents in ascending order, so an equal score never displaces a kept
                    // one (strictly greater); shared bound: ties with another list's k-th score may still win on the id
                    constexpr int kBlocks = kRange / 8 / 128;  // 128-document blocks per warp
                    float4* mine4 = acc4 + warp * (kRange / 8 / 4) + lane;
                    float thr_own = top.worst == ~0ull ? 0.0f : ord_unkey(~(uint32_t)(top.worst >> 32));
                    float thr_sh = __uint_as_float(*reinterpret_cast<volatile uint32_t*>(&s_thr));
                    // pass 1: the largest score of the warp's slice; nearly always it does not reach the bounds
                    float top_s = 0.0f;
#pragma unroll
                    for (int bI = 0; bI < kBlocks; ++bI) {
                        const float4 s4 = mine4[bI * 32];
                        top_s = fmaxf(top_s, fmaxf(fmaxf(s4.x, s4.y), fmaxf(s4.z, s4.w)));
                    }
                    if (__ballot_sync(FULL_MASK, top_s > thr_own && top_s >= thr_sh)) {
#pragma unroll 1
                        for (int bI = 0; bI < kBlocks; ++bI) {
                            const float4 s4 = mine4[bI * 32];
                            const float mx = fmaxf(fmaxf(s4.x, s4.y), fmaxf(s4.z, s4.w));
                            if (!__ballot_sync(FULL_MASK, mx > thr_own && mx >= thr_sh)) continue;
                            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
                            const uint32_t doc0 = base_doc + warp * (kRange / 8) + bI * 128 + lane * 4;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                uint64_t key = ~0ull;
                                if (sv[e] > 0.0f && sv[e] >= thr_sh) key = ((uint64_t)(~ord_key(sv[e])) << 32) | (doc0 + e);
                                top.offer(key);
                            }
                            if (top.worst != ~0ull) {
                                thr_own = ord_unkey(~(uint32_t)(top.worst >> 32));
                                if (lane == 0 && thr_own > thr_sh) atomicMax(&s_thr, __float_as_uint(thr_own));  // positive floats order as their bits
                            }
                            thr_sh = __uint_as_float(*reinterpret_cast<volatile uint32_t*>(&s_thr));
                        }
                    }
#pragma unroll
                    for (int bI = 0; bI < kBlocks; ++bI) mine4[bI * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
                    touched = false;
                }
                if (threadIdx.x == 0 && parts > 1) {  // exchange the bound with the query's other work items
                    const uint32_t mine = *reinterpret_cast<volatile uint32_t*>(&s_thr);
                    if (thr_seen > mine) atomicMax(&s_thr, thr_seen);
                    if (mine > thr_seen) atomicMax(qthr + q, mine);  // no return value used: fire and forget
                    thr_seen = *reinterpret_cast<volatile uint32_t*>(qthr + q);  // consumed at the end of the next range
                }
                __syncthreads();  // accumulator clean and ring entry r + 2 written before anyone goes on
            }
            r = nr;
            j0 = nj0;
        };
        FlatHold ha, hb;
        if (r < r_end) load_round(ha, r, 0, s_total[r & 3]);
        while (r < r_end) {
            round(ha, hb);
            if (r >= r_end) break;
            round(hb, ha);
        }
        // merge the eight per-warp lists by RANK, all threads at once: the lists are sorted and their keys distinct, so a
        // key's place in the merged order is the sum of its lower bounds in the eight lists (a serial merge in one warp
        // held the other seven at the barrier for ~5 us per item).  Then publish; the query's last item merges the parts.
        top.store(lists + (size_t)warp * k, k);
        uint64_t* merged = lists + (size_t)8 * k;
        for (uint32_t i = threadIdx.x; i < k; i += 256) merged[i] = ~0ull;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < 8 * k; i += 256) {
            const uint64_t key = lists[i];
            if (key == ~0ull) continue;
            uint32_t rank = 0;
#pragma unroll 1
            for (uint32_t w = 0; w < 8; ++w) {
                const uint64_t* l = lists + (size_t)w * k;
                uint32_t lo = 0, hi = k;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (l[mid] < key) lo = mid + 1; else hi = mid;
                }
                rank += lo;
            }
            if (rank < k) merged[rank] = key;
        }
        __syncthreads();
        if (warp == 0) {
            bool finisher = true;
            if (parts > 1) {
                uint64_t* mine = partial + ((size_t)q * parts + part) * k;
                for (uint32_t j = lane; j < k; j += 32) mine[j] = merged[j];
                __threadfence();
                __syncwarp();
                uint32_t t = 0;
                if (lane == 0) t = atomicAdd(tickets + q, 1u);
                t = __shfl_sync(FULL_MASK, t, 0);
                finisher = t == parts - 1;
                if (finisher) {
                    __threadfence();
#pragma unroll
                    for (int sI = 0; sI < R; ++sI) top.k[sI] = 32u * sI + lane < k ? merged[32u * sI + lane] : ~0ull;
                    top.worst = merged[k - 1];  // ~0 unless the list is full
                    for (uint32_t pp = 0; pp < parts; ++pp) {
                        if (pp == part) continue;
                        const volatile uint64_t* other = partial + ((size_t)q * parts + pp) * k;
                        for (uint32_t jj = 0; jj < k; jj += 32) top.offer(jj + lane < k ? other[jj + lane] : ~0ull);
                    }
                    if (lane == 0) tickets[q] = 0;  // ready for the next call
                    __syncwarp();
                    top.store(merged, k);
                    __syncwarp();
                }
            }
            if (finisher) {
                uint32_t len = 0;
                for (uint32_t jj = 0; jj < k; jj += 32) {
                    const uint32_t j = jj + lane;
                    if (j < k) {
                        const uint64_t key = merged[j];
                        uint32_t doc = VELES_INVALID_ID;
                        float sc = __uint_as_float(0x7fc00000u);
                        if (key != ~0ull) {
                            doc = (uint32_t)key;
                            sc = ord_unkey(~(uint32_t)(key >> 32));
                            ++len;
                        }
                        out_doc[(size_t)q * k + j] = doc;
                        out_score[(size_t)q * k + j] = sc;
                    }
                }
                len = __reduce_add_sync(FULL_MASK, len);
                if (lane == 0) out_cnt[q] = len;
            }
        }
        __syncthreads();  // `lists` and s_item are reused by the next item
    }
}

// ---- one WARP per (query, span of sub-ranges): owner computes, no block barriers (k <= kMultiK) -------------------------
// Every kernel above shares one accumulator among the eight warps of a CTA, so the terms of a range must be separated by
// block barriers (a document may occur in two terms, held by two warps), and the ncu profiles of all of them end the
// same way: a third of the samples at barriers, half the issue slots empty, and any warp that meets a candidate for
// the top-k makes the other seven wait.  The cure is a finer skip table: skipf[term][s] = first posting of the term
// with doc >= s * 1024, so ONE warp can find, with no search, every posting of every query token that falls into
// "its" 1024 documents.  A warp then owns a span of sub-ranges of one query outright:
//   per sub-range: lane g holds token g's slice boundaries (the next ones are requested two sub-ranges ahead); the
//   slices are read 32 postings at a time in token order (typically one step per token), up to kSubH steps in flight;
//   they are added into the warp's private 4 KB accumulator in the same order with a __syncwarp between tokens -- the
//   reference's f32 order per document; the warp scans its accumulator (zeroing it) against its own register-resident
//   top-k and the best k-th score published for the query (qthr).
// No block barrier anywhere, 40+ independent warps per SM.  Work items are (query, part) with parts ~ 2 x resident
// warps / queries, part-major; the last part of a query to finish merges the parts' lists (ticket per query).
// Cost: the fine table, 4 bytes x terms x (docs / 1024); it is built when it fits the budget in veles_bm25_from_csr.
constexpr int kSubWarps = 4;  // warps per CTA (they share nothing but the launch)
constexpr int kSubH = 6;      // 32-posting steps in flight per pass
template <int R, int OCC>
__global__ void __launch_bounds__(kSubWarps * 32, OCC) bm25_sub_kernel(Bm25View v, const uint32_t* __restrict__ q_ptr,
                                                                     const uint32_t* __restrict__ q_terms, uint32_t nq, uint32_t k,
                                                                     uint32_t parts, uint64_t* __restrict__ partial,
                                                                     uint32_t* __restrict__ tickets, uint32_t* __restrict__ qthr,
                                                                     uint32_t* __restrict__ out_doc, float* __restrict__ out_score,
                                                                     uint32_t* __restrict__ out_cnt, uint32_t* __restrict__ work) {
    extern __shared__ __align__(16) uint8_t bs_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* acc = reinterpret_cast<float*>(bs_smem) + warp * kFine;
    float4* acc4 = reinterpret_cast<float4*>(acc) + lane;
    uint64_t* stage = reinterpret_cast<uint64_t*>(bs_smem + (size_t)kSubWarps * kFine * 4) + (size_t)warp * k;  // k keys
#pragma unroll
    for (int bI = 0; bI < (int)(kFine / 128); ++bI) acc4[bI * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    for (;;) {  // persistent: work items are handed out to warps by a global counter
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(work, 1u);
        item = __shfl_sync(FULL_MASK, item, 0);
        if (item >= nq * parts) return;
        // part-major: a query's later parts start with the k-th score its earlier parts have published (qthr)
        const uint32_t part = item / nq, q = item - part * nq;
        const uint32_t s_begin = (uint32_t)((uint64_t)part * v.n_fine / parts);
        const uint32_t s_end = (uint32_t)((uint64_t)(part + 1) * v.n_fine / parts);
        const uint32_t t0 = q_ptr[q], t1 = q_ptr[q + 1];
        const uint32_t nt = min(kQueryTerms, t1 - t0);  // host guarantees t1 - t0 <= kQueryTerms on this path
        // lane g < nt: query token g -- its fine skip row, its posting list, the boundaries of the next three sub-ranges
        const uint32_t* row = nullptr;
        const uint2* plist = nullptr;
        uint32_t b0 = 0, b1 = 0, b2 = 0;
        if (lane < nt) {
            const uint32_t term = q_terms[t0 + lane];
            if (term < v.n_terms) {  // unknown term: df = 0, contributes nothing
                row = v.skipf + (size_t)term * (v.n_fine + 1);
                plist = v.post_dc + v.term_ptr[term];
                b0 = row[s_begin];
                b1 = row[s_begin + 1];
                b2 = row[min(s_begin + 2, v.n_fine)];
            }
        }
        RegTopK<R> top;
        top.init(k, lane);
        uint32_t thr_bits = 0;
        if (lane == 0) thr_bits = *reinterpret_cast<volatile uint32_t*>(qthr + q);
        float thr_sh = __uint_as_float(__shfl_sync(FULL_MASK, thr_bits, 0));
        uint32_t thr_next = thr_bits;  // lane 0: qthr[q], re-read every sub-range, consumed one sub-range later
        for (uint32_t sr = s_begin; sr < s_end; ++sr) {
            uint32_t b3 = b2;
            if (row && sr + 3 <= v.n_fine) b3 = row[sr + 3];  // needed three sub-ranges from now
            if (parts > 1) {
                thr_sh = fmaxf(thr_sh, __uint_as_float(__shfl_sync(FULL_MASK, thr_next, 0)));
                if (lane == 0) thr_next = *reinterpret_cast<volatile uint32_t*>(qthr + q);
            }
            const uint32_t len = b1 - b0;  // <= kFine
            uint32_t m = __ballot_sync(FULL_MASK, len != 0);
            if (m) {
                const uint32_t base_doc = sr * kFine;
                const uint2* my_first = plist + b0;
                uint32_t off = 0;
                while (m) {
                    // ---- request up to kSubH steps, token order ----
                    uint32_t d[kSubH], c[kSubH];
                    uint32_t fresh = 0;  // bit h: step h is the first of its token (a __syncwarp goes before it)
                    int issued = 0;
#pragma unroll
                    for (int h = 0; h < kSubH; ++h) {
                        d[h] = 0xffffffffu;
                        c[h] = 0;
                        if (m) {  // uniform
                            const uint32_t g = __ffs(m) - 1;
                            const uint32_t lg = __shfl_sync(FULL_MASK, len, g);
                            const uint64_t first = __shfl_sync(FULL_MASK, reinterpret_cast<uint64_t>(my_first), g);
                            const uint32_t o = off + lane;
                            const uint32_t ok = o < lg;
                            const uint64_t src = first + (uint64_t)(ok ? o : 0) * 8;
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p ld.global.nc.v2.u32 {%0, %1}, [%2];\n\t}"
                                : "+r"(d[h]), "+r"(c[h])
                                : "l"(src), "r"(ok));
                            if (off == 0) fresh |= 1u << h;
                            off += 32;
                            if (off >= lg) {
                                m &= m - 1;
                                off = 0;
                            }
                            issued = h + 1;
                        }
                    }
                    // ---- apply in the same order: per document, contributions arrive in query-token order ----
#pragma unroll
                    for (int h = 0; h < kSubH; ++h) {
                        if (h < issued) {  // uniform
                            if ((fresh >> h) & 1u) __syncwarp();
                            if (d[h] != 0xffffffffu) acc[d[h] - base_doc] = __fadd_rn(acc[d[h] - base_doc], __uint_as_float(c[h]));
                        }
                    }
                    __syncwarp();
                }
                // ---- scan the warp's 1024 accumulators; what is read is zeroed for the next sub-range ----
                // own list: documents arrive in ascending order, so an equal score never displaces a kept one (strictly
                // greater); published bound: ties with another part's k-th score may still win on the document id
                float thr_own = top.worst == ~0ull ? 0.0f : ord_unkey(~(uint32_t)(top.worst >> 32));
                constexpr int kBlocks = kFine / 128;
                float top_s = 0.0f;
#pragma unroll
                for (int bI = 0; bI < kBlocks; ++bI) {
                    const float4 s4 = acc4[bI * 32];
                    top_s = fmaxf(top_s, fmaxf(fmaxf(s4.x, s4.y), fmaxf(s4.z, s4.w)));
                }
                if (__ballot_sync(FULL_MASK, top_s > thr_own && top_s >= thr_sh)) {
                    bool grew = false;
#pragma unroll 1
                    for (int bI = 0; bI < kBlocks; ++bI) {
                        const float4 s4 = acc4[bI * 32];
                        const float mx = fmaxf(fmaxf(s4.x, s4.y), fmaxf(s4.z, s4.w));
                        if (!__ballot_sync(FULL_MASK, mx > thr_own && mx >= thr_sh)) continue;
                        const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
                        const uint32_t doc0 = base_doc + bI * 128 + lane * 4;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            uint64_t key = ~0ull;
                            if (sv[e] > 0.0f && sv[e] >= thr_sh) key = ((uint64_t)(~ord_key(sv[e])) << 32) | (doc0 + e);
                            top.offer(key);
                        }
                        if (top.worst != ~0ull) {
                            thr_own = ord_unkey(~(uint32_t)(top.worst >> 32));
                            grew = true;
                        }
                    }
                    if (grew && thr_own > thr_sh) {  // publish: k documents of this query score at least thr_own
                        thr_sh = thr_own;
                        if (lane == 0 && parts > 1) atomicMax(qthr + q, __float_as_uint(thr_own));  // positive floats order as their bits
                    }
                }
#pragma unroll
                for (int bI = 0; bI < kBlocks; ++bI) acc4[bI * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
                __syncwarp();
            }
            b0 = b1;
            b1 = b2;
            b2 = b3;
        }
        // publish the item's list; the query's last item to finish merges the parts and writes the result
        bool finisher = true;
        if (parts > 1) {
            top.store(partial + ((size_t)q * parts + part) * k, k);
            __threadfence();
            __syncwarp();
            uint32_t t = 0;
            if (lane == 0) t = atomicAdd(tickets + q, 1u);
            t = __shfl_sync(FULL_MASK, t, 0);
            finisher = t == parts - 1;
            if (finisher) {
                __threadfence();
                for (uint32_t pp = 0; pp < parts; ++pp) {
                    if (pp == part) continue;
                    const volatile uint64_t* other = partial + ((size_t)q * parts + pp) * k;
                    for (uint32_t jj = 0; jj < k; jj += 32) top.offer(jj + lane < k ? other[jj + lane] : ~0ull);
                }
                if (lane == 0) tickets[q] = 0;  // ready for the next call
            }
        }
        if (finisher) {
            top.store(stage, k);
            __syncwarp();
            uint32_t len = 0;
            for (uint32_t jj = 0; jj < k; jj += 32) {
                const uint32_t j = jj + lane;
                if (j < k) {
                    const uint64_t key = stage[j];
                    uint32_t doc = VELES_INVALID_ID;
                    float sc = __uint_as_float(0x7fc00000u);
                    if (key != ~0ull) {
                        doc = (uint32_t)key;
                        sc = ord_unkey(~(uint32_t)(key >> 32));
                        ++len;
                    }
                    out_doc[(size_t)q * k + j] = doc;
                    out_score[(size_t)q * k + j] = sc;
                }
            }
            len = __reduce_add_sync(FULL_MASK, len);
            if (lane == 0) out_cnt[q] = len;
            __syncwarp();
        }
    }
}

// ---- one CTA per query, posting-driven (k <= kMultiK): hashed accumulation over adaptive doc-id windows -------------
// bm25_query_kernel above pays for every doc-id range of a query -- zeroing and re-scanning a dense 7168-slot
// accumulator ~140 times -- whatever the number of postings that fall into it (~700).  Here the work follows the
// postings: the CTA takes as many consecutive ranges as hold at most kWindowPostings postings of the query's terms
// (one range at least; the skip table gives the counts), accumulates them into an open-addressing table in shared
// memory keyed by document (claim a slot with atomicCAS, then a plain add: a document occurs once per posting list, and
// a block barrier separates the terms, so every document still receives its contributions in query-token order), and
// scans the table -- two slots per posting instead of ten.  ~25 windows per query instead of 140 ranges.
constexpr uint32_t kHashSlots = 8192;        // >= kRange, so a single dense range can never overflow the table
constexpr uint32_t kWindowPostings = 4096;   // target load factor 0.5
constexpr uint32_t kWindowLook = 16;         // ranges examined per window decision
template <int R>
__global__ void __launch_bounds__(256, 3) bm25_hash_kernel(Bm25View v, const uint32_t* __restrict__ q_ptr,
                                                           const uint32_t* __restrict__ q_terms, uint32_t nq, uint32_t k,
                                                           uint32_t* __restrict__ out_doc, float* __restrict__ out_score,
                                                           uint32_t* __restrict__ out_cnt, uint32_t* __restrict__ work) {
    extern __shared__ __align__(16) uint8_t bh_smem[];
    uint32_t* hkey = reinterpret_cast<uint32_t*>(bh_smem);
    float* hval = reinterpret_cast<float*>(bh_smem + kHashSlots * 4);
    uint64_t* lists = reinterpret_cast<uint64_t*>(bh_smem + kHashSlots * 8);  // 8 x k keys for the final merge
    __shared__ uint64_t s_lo[kQueryTerms], s_hi[kQueryTerms];
    __shared__ float s_idf[kQueryTerms];
    __shared__ uint32_t s_cnt[kWindowLook];
    __shared__ uint32_t s_q, s_rend;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const float k1p1 = __fadd_rn(v.k1, 1.0f);
    for (;;) {  // persistent: queries are handed out by a global counter
        if (threadIdx.x == 0) s_q = atomicAdd(work, 1u);
        __syncthreads();
        const uint32_t q = s_q;
        if (q >= nq) return;
        const uint32_t t0 = q_ptr[q], t1 = q_ptr[q + 1];
        const uint32_t nt = min(kQueryTerms, t1 - t0);  // host guarantees t1 - t0 <= kQueryTerms on this path
        RegTopK<R> top;
        top.init(k, lane);
        uint64_t my_lo = 0;
        const uint64_t* my_skip = nullptr;
        if (threadIdx.x < nt) {
            const uint32_t term = q_terms[t0 + threadIdx.x];
            float idf = 0.0f;
            if (term < v.n_terms) {
                idf = v.idf[term];
                my_skip = v.skip + (size_t)term * (v.n_ranges + 1);
                my_lo = my_skip[0];
            }
            s_idf[threadIdx.x] = idf;
        }
        uint32_t r = 0;
        while (r < v.n_ranges) {
            // ---- window [r, r_end): postings of all terms in the next 1..kWindowLook ranges ----
            if (threadIdx.x < kWindowLook) s_cnt[threadIdx.x] = 0;
            for (uint32_t i = threadIdx.x; i < kHashSlots / 4; i += blockDim.x) {
                reinterpret_cast<uint4*>(hkey)[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                reinterpret_cast<float4*>(hval)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncthreads();
            if (threadIdx.x < nt && my_skip) {
                for (uint32_t j = 1; j <= kWindowLook && r + j <= v.n_ranges; ++j) {
                    const uint64_t c = my_skip[r + j] - my_lo;
                    atomicAdd(&s_cnt[j - 1], (uint32_t)min(c, (uint64_t)0x7fffffffu / kQueryTerms));
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t j = 1;
                while (j < kWindowLook && r + j < v.n_ranges && s_cnt[j] <= kWindowPostings) ++j;  // s_cnt[j] = count of j+1 ranges
                s_rend = r + j;
            }
            __syncthreads();
            const uint32_t r_end = s_rend;
            if (threadIdx.x < nt) {
                s_lo[threadIdx.x] = my_lo;
                const uint64_t hi = my_skip ? my_skip[r_end] : my_lo;
                s_hi[threadIdx.x] = hi;
                my_lo = hi;
            }
            __syncthreads();
            // ---- (term, 256-posting chunk) walk in query order, kDepth chunks in flight ----
            uint32_t cg = 0;
            while (cg < nt && s_hi[cg] == s_lo[cg]) ++cg;
            uint64_t cbase = cg < nt ? s_lo[cg] : 0;
            uint32_t qg[kDepth], qd[kDepth], qtf[kDepth];
            float qden[kDepth];
            auto request = [&](int slot) {
                qg[slot] = cg;
                qd[slot] = VELES_INVALID_ID;
                qtf[slot] = 0;
                qden[slot] = 1.0f;
                if (cg < nt) {
                    const uint64_t p = cbase + threadIdx.x;
                    if (p < s_hi[cg]) {
                        qd[slot] = v.post_doc[p];
                        qtf[slot] = v.post_tf[p];
                        qden[slot] = v.post_den[p];
                    }
                    cbase += blockDim.x;
                    if (cbase >= s_hi[cg]) {
                        ++cg;
                        while (cg < nt && s_hi[cg] == s_lo[cg]) ++cg;
                        cbase = cg < nt ? s_lo[cg] : 0;
                    }
                }
            };
#pragma unroll
            for (int i = 0; i < kDepth; ++i) request(i);
            bool touched = false;
            while (qg[0] < nt) {
                if (qd[0] != VELES_INVALID_ID) {
                    const float num = __fmul_rn((float)qtf[0], k1p1);
                    const float contrib = __fdiv_rn(__fmul_rn(s_idf[qg[0]], num), qden[0]);
                    uint32_t h = (qd[0] * 2654435761u) >> 19;  // 13 bits
                    for (;;) {
                        const uint32_t prev = atomicCAS(&hkey[h], 0xffffffffu, qd[0]);
                        if (prev == 0xffffffffu || prev == qd[0]) break;
                        h = (h + 1) & (kHashSlots - 1);
                    }
                    hval[h] = __fadd_rn(hval[h], contrib);  // this document's only posting of this term
                }
                touched = true;
                if (qg[1] != qg[0]) __syncthreads();  // next term's contributions come after this term's
#pragma unroll
                for (int i = 0; i + 1 < kDepth; ++i) {
                    qg[i] = qg[i + 1];
                    qd[i] = qd[i + 1];
                    qtf[i] = qtf[i + 1];
                    qden[i] = qden[i + 1];
                }
                request(kDepth - 1);
            }
            __syncthreads();
            // ---- scan: this warp's slice of the table ----
            if (touched) {
                const uint32_t per = kHashSlots / nwarps, i_begin = warp * per;
                for (uint32_t i0 = i_begin; i0 < i_begin + per; i0 += 128) {
                    const uint4 d4 = reinterpret_cast<const uint4*>(hkey)[(i0 >> 2) + lane];
                    const float4 s4 = reinterpret_cast<const float4*>(hval)[(i0 >> 2) + lane];
                    const float thr = top.worst == ~0ull ? 0.0f : ord_unkey(~(uint32_t)(top.worst >> 32));
                    const bool any = (s4.x > 0.0f && s4.x >= thr) || (s4.y > 0.0f && s4.y >= thr) ||
                                     (s4.z > 0.0f && s4.z >= thr) || (s4.w > 0.0f && s4.w >= thr);
                    if (!__ballot_sync(FULL_MASK, any)) continue;
                    const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
                    const uint32_t dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        uint64_t key = ~0ull;
                        if (sv[e] > 0.0f && dv[e] != 0xffffffffu) key = ((uint64_t)(~ord_key(sv[e])) << 32) | dv[e];
                        top.offer(key);
                    }
                }
            }
            __syncthreads();  // the table is rewritten by the next window
            r = r_end;
        }
        // merge the eight per-warp lists
        top.store(lists + (size_t)warp * k, k);
        __syncthreads();
        if (warp == 0) {
            for (uint32_t w = 1; w < nwarps; ++w) {
                const uint64_t* other = lists + (size_t)w * k;
                for (uint32_t j0 = 0; j0 < k; j0 += 32) {
                    const uint32_t j = j0 + lane;
                    top.offer(j < k ? other[j] : ~0ull);
                }
            }
            top.store(lists, k);
            __syncwarp();
            uint32_t len = 0;
            for (uint32_t j0 = 0; j0 < k; j0 += 32) {
                const uint32_t j = j0 + lane;
                if (j < k) {
                    const uint64_t key = lists[j];
                    uint32_t doc = VELES_INVALID_ID;
                    float sc = __uint_as_float(0x7fc00000u);
                    if (key != ~0ull) {
                        doc = (uint32_t)key;
                        sc = ord_unkey(~(uint32_t)(key >> 32));
                        ++len;
                    }
                    out_doc[(size_t)q * k + j] = doc;
                    out_score[(size_t)q * k + j] = sc;
                }
            }
            len = __reduce_add_sync(FULL_MASK, len);
            if (lane == 0) out_cnt[q] = len;
        }
        __syncthreads();  // `lists` and s_q are reused by the next query
    }
}

// one warp per query: k smallest keys over its n_ranges x k partial keys
__global__ void __launch_bounds__(32) bm25_merge_kernel(const uint64_t* __restrict__ partial, uint32_t n_ranges, uint32_t k,
                                                        uint32_t* __restrict__ out_doc, float* __restrict__ out_score,
                                                        uint32_t* __restrict__ out_cnt) {
    extern __shared__ __align__(16) uint64_t mres[];
    const uint32_t lane = threadIdx.x, q = blockIdx.x;
    const uint64_t* src = partial + (size_t)q * n_ranges * k;
    const uint32_t total = n_ranges * k;
    uint32_t len = 0;
    uint64_t worst = ~0ull;
    for (uint32_t i0 = 0; i0 < total; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint64_t key = i < total ? src[i] : ~0ull;
        uint32_t msk = __ballot_sync(FULL_MASK, key < worst);
        while (msk) {
            const uint32_t s = __ffs(msk) - 1;
            msk &= msk - 1;
            const uint64_t kk = __shfl_sync(FULL_MASK, key, s);
            if (kk >= worst) continue;
            const uint32_t pos = lower_bound_warp(mres, len, kk, lane);
            if (len < k) {
                insert_at(mres, pos, len + 1, kk, lane);
                ++len;
            } else {
                insert_at(mres, pos, len, kk, lane);
            }
            if (len == k) worst = mres[k - 1];
        }
    }
    __syncwarp();
    for (uint32_t j = lane; j < k; j += 32) {
        uint32_t doc = VELES_INVALID_ID;
        float s = __uint_as_float(0x7fc00000u);
        if (j < len) {
            const uint64_t key = mres[j];
            doc = (uint32_t)key;
            s = ord_unkey(~(uint32_t)(key >> 32));
        }
        out_doc[(size_t)q * k + j] = doc;
        out_score[(size_t)q * k + j] = s;
    }
    if (lane == 0) out_cnt[q] = len;
}

}  // namespace veles

using namespace veles;

extern "C" {

int32_t veles_bm25_from_csr(uint32_t n_terms, const uint64_t* term_ptr, const uint32_t* post_doc,
                            const uint32_t* post_tf, const uint32_t* df, uint32_t n_doc_slots, const uint32_t* doc_len,
                            uint64_t doc_count, uint64_t total_len, float k1, float b, veles_bm25_t** out) {
    VELES_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    VELES_REQUIRE(n_terms == 0 || (term_ptr && df), "term arrays are NULL");
    VELES_REQUIRE(n_doc_slots == 0 || doc_len, "doc_len is NULL");
    const uint64_t np = n_terms ? term_ptr[n_terms] : 0;
    VELES_REQUIRE(np == 0 || (post_doc && post_tf), "posting arrays are NULL");
    std::unique_ptr<veles_bm25> ix(new veles_bm25());
    ix->n_terms = n_terms;
    ix->n_doc_slots = n_doc_slots;
    ix->n_ranges = std::max(1u, (n_doc_slots + kRange - 1) / kRange);
    ix->doc_count = doc_count;
    ix->total_len = total_len;
    ix->n_postings = np;
    ix->k1 = k1;
    ix->b = b;
    ix->avgdl = doc_count ? (float)total_len / (float)doc_count : 0.0f;  // bm25.rs:281
    // idf on the host with logf, as f32::ln does (bm25.rs:297-306)
    std::vector<float> idf(std::max(n_terms, 1u), 0.0f);
    const float n = (float)doc_count;
    for (uint32_t t = 0; t < n_terms; ++t) {
        if (df[t] == 0) continue;
        const float df_f = (float)df[t];
        const float num = n - df_f + 0.5f;
        const float den = df_f + 0.5f;
        const float r = num / den + 1.0f;
        idf[t] = std::log(r);
    }
    // skip table + validation (doc ids ascending within a term, < n_doc_slots, tf > 0)
    const uint32_t nr = ix->n_ranges;
    std::vector<uint64_t> skip((size_t)std::max(n_terms, 1u) * (nr + 1), 0);
    for (uint32_t t = 0; t < n_terms; ++t) {
        uint64_t p = term_ptr[t];
        const uint64_t e = term_ptr[t + 1];
        VELES_REQUIRE(e >= p, "term_ptr must be non-decreasing");
        uint64_t* row = skip.data() + (size_t)t * (nr + 1);
        uint32_t r = 0;
        row[0] = p;
        int64_t prev = -1;
        for (; p < e; ++p) {
            const uint32_t d = post_doc[p];
            VELES_REQUIRE(d < n_doc_slots, "posting doc id %u out of range", d);
            VELES_REQUIRE((int64_t)d > prev, "postings of term %u are not strictly ascending by doc id", t);
            VELES_REQUIRE(post_tf[p] > 0, "posting with tf == 0 (term %u, doc %u)", t, d);
            prev = d;
            while (r < d / kRange) row[++r] = p;
        }
        while (r < nr) row[++r] = e;
    }
    // fine skip table for bm25_sub_kernel: u32 offsets relative to term_ptr, one row per term.  Built when it stays within
    // 2 GiB and 4x the postings themselves (a large vocabulary over few documents would be all table): otherwise the
    // query path uses bm25_flat_kernel, which needs only the coarse table.
    std::vector<uint32_t> skipf;
    {
        const uint32_t nf = std::max(1u, (n_doc_slots + kFine - 1) / kFine);
        const uint64_t bytes = (uint64_t)std::max(n_terms, 1u) * (nf + 1) * 4;
        bool fits = bytes <= ((uint64_t)2 << 30) && bytes <= std::max<uint64_t>(np * 8 * 4, (uint64_t)64 << 20);
        for (uint32_t t = 0; fits && t < n_terms; ++t) fits = term_ptr[t + 1] - term_ptr[t] <= 0xffffffffull;
        if (std::getenv("VELES_BM25_NO_FINE_TABLE")) fits = false;
        if (fits) {
            ix->n_fine = nf;
            skipf.assign((size_t)std::max(n_terms, 1u) * (nf + 1), 0);
            for (uint32_t t = 0; t < n_terms; ++t) {
                const uint64_t p0 = term_ptr[t], e = term_ptr[t + 1];
                uint32_t* row = skipf.data() + (size_t)t * (nf + 1);
                uint32_t r = 0;
                for (uint64_t p = p0; p < e; ++p) {
                    const uint32_t sub = post_doc[p] / kFine;
                    while (r < sub) row[++r] = (uint32_t)(p - p0);
                }
                while (r < nf) row[++r] = (uint32_t)(e - p0);
            }
        }
    }
    // per-posting denominator, with the reference's f32 operation order (bm25.rs:357-358, 371-372); the host
    // compiler targets baseline x86-64, so nothing here is contracted into an FMA
    std::vector<float> den(std::max<uint64_t>(np, 1));
    {
        const volatile float one_minus_b = 1.0f - b;
        for (uint64_t p = 0; p < np; ++p) {
            const float dl = (float)doc_len[post_doc[p]];
            const volatile float bl = b * dl;
            const volatile float q = bl / ix->avgdl;
            const volatile float len_norm = one_minus_b + q;
            const volatile float kl = k1 * len_norm;
            den[p] = (float)post_tf[p] + kl;
        }
    }
    // The whole contribution of a posting, idf * (tf * (k1 + 1.0)) / (tf + k1 * len_norm) (bm25.rs:360-375), depends only
    // on the snapshot: idf on the term, the rest on the posting.  Precomputed with the reference's operation order, packed
    // next to the doc id: the slice kernel reads 8 bytes per posting and does one add.
    std::vector<uint2> dc(std::max<uint64_t>(np, 1));
    {
        const volatile float k1p1 = k1 + 1.0f;
        for (uint32_t t = 0; t < n_terms; ++t) {
            const float idf_t = idf[t];
            for (uint64_t p = term_ptr[t]; p < term_ptr[t + 1]; ++p) {
                const volatile float num = (float)post_tf[p] * k1p1;
                const volatile float top = idf_t * num;
                const volatile float c = top / den[p];
                const float cf = c;
                uint32_t bits;
                std::memcpy(&bits, &cf, 4);
                dc[p] = make_uint2(post_doc[p], bits);
            }
        }
    }
    VELES_TRY(ix->post_dc.alloc(std::max<size_t>(np * 8, 16)));
    if (np) VELES_CUDA(cudaMemcpy(ix->post_dc.p, dc.data(), np * 8, cudaMemcpyHostToDevice));
    VELES_TRY(ix->term_ptr.alloc(((size_t)n_terms + 1) * 8));
    if (n_terms) VELES_CUDA(cudaMemcpy(ix->term_ptr.p, term_ptr, ((size_t)n_terms + 1) * 8, cudaMemcpyHostToDevice));
    VELES_TRY(ix->post_doc.alloc(std::max<size_t>(np * 4, 16)));
    VELES_TRY(ix->post_tf.alloc(std::max<size_t>(np * 4, 16)));
    VELES_TRY(ix->post_den.alloc(std::max<size_t>(np * 4, 16)));
    VELES_TRY(ix->doc_len.alloc(std::max<size_t>((size_t)n_doc_slots * 4, 16)));
    VELES_TRY(ix->idf.alloc(idf.size() * 4));
    VELES_TRY(ix->skip.alloc(skip.size() * 8));
    if (ix->n_fine) {
        VELES_TRY(ix->skipf.alloc(skipf.size() * 4));
        VELES_CUDA(cudaMemcpy(ix->skipf.p, skipf.data(), skipf.size() * 4, cudaMemcpyHostToDevice));
    }
    if (np) {
        VELES_CUDA(cudaMemcpy(ix->post_doc.p, post_doc, np * 4, cudaMemcpyHostToDevice));
        VELES_CUDA(cudaMemcpy(ix->post_tf.p, post_tf, np * 4, cudaMemcpyHostToDevice));
        VELES_CUDA(cudaMemcpy(ix->post_den.p, den.data(), np * 4, cudaMemcpyHostToDevice));
    }
    if (n_doc_slots) VELES_CUDA(cudaMemcpy(ix->doc_len.p, doc_len, (size_t)n_doc_slots * 4, cudaMemcpyHostToDevice));
    VELES_CUDA(cudaMemcpy(ix->idf.p, idf.data(), idf.size() * 4, cudaMemcpyHostToDevice));
    VELES_CUDA(cudaMemcpy(ix->skip.p, skip.data(), skip.size() * 8, cudaMemcpyHostToDevice));
    *out = ix.release();
    return VELES_OK;
}

int32_t veles_bm25_free(veles_bm25_t* ix) {
    delete ix;
    return VELES_OK;
}

// The query kernels over a batch whose term lists are already in ix->q_ptr_d / ix->q_terms_d; results go to
// ix->out_{doc,score,cnt}_d.  Only enqueues on `st`; the caller holds ix->mu.  max_terms = longest query of the batch.
static int32_t bm25_enqueue(const veles_bm25* ix, uint32_t nq, uint32_t k, uint32_t max_terms, cudaStream_t st) {
    NvtxRange nvtx_range("veles::bm25 query kernels (Bm25Index::search)");
    Bm25View v;
    v.post_doc = ix->post_doc.as<uint32_t>();
    v.post_tf = ix->post_tf.as<uint32_t>();
    v.post_den = ix->post_den.as<float>();
    v.post_dc = ix->post_dc.as<uint2>();
    v.doc_len = ix->doc_len.as<uint32_t>();
    v.idf = ix->idf.as<float>();
    v.skip = ix->skip.as<uint64_t>();
    v.term_ptr = ix->term_ptr.as<uint64_t>();
    v.skipf = ix->skipf.as<uint32_t>();
    v.n_fine = ix->n_fine;
    v.n_terms = ix->n_terms;
    v.n_ranges = ix->n_ranges;
    v.n_doc_slots = ix->n_doc_slots;
    v.k1 = ix->k1;
    v.b = ix->b;
    v.avgdl = ix->avgdl;
    if (k <= kMultiK && max_terms <= kQueryTerms && std::getenv("VELES_BM25_RANGE_KERNEL") == nullptr) {
        // one CTA per query walking its doc-id ranges
        // default: bm25_sub_kernel (one warp per (query, span of 1024-document sub-ranges), no block barriers) when the
        // snapshot has the fine skip table, else bm25_flat_kernel ((query, part) items per CTA, warp-step mapping, rounds
        // pipelined; VELES_BM25_FLAT=1 forces it).  VELES_BM25_WALK=1 selects round 1's bm25_query_kernel; three other
        // round-2 restructurings stay selectable and are measured slower on B200 (profiles/README.md, DESIGN.md section
        // 4.4): VELES_BM25_SLICE=1, VELES_BM25_PREFETCH=1, VELES_BM25_HASH=1.
        const bool walk = std::getenv("VELES_BM25_WALK") != nullptr;
        const bool slice = !walk && std::getenv("VELES_BM25_SLICE") != nullptr;
        const bool pre = !walk && !slice && std::getenv("VELES_BM25_PREFETCH") != nullptr;
        const bool hash = !walk && !slice && !pre && std::getenv("VELES_BM25_HASH") != nullptr;
        const bool other = walk || slice || pre || hash;
        const bool sub = !other && ix->n_fine > 0 && std::getenv("VELES_BM25_FLAT") == nullptr;
        const bool flat = !other && !sub;
        const size_t smem = (hash ? (size_t)kHashSlots * 8 : (size_t)kRange * 4) + (size_t)8 * k * 8;
        int per_sm = 0, dev = 0, sms = 0;
        VELES_CUDA(cudaGetDevice(&dev));
        VELES_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (sub) {
            // 10 CTAs (40 warps) per SM at 48 registers.  Measured on B200, 1024 queries: 12 CTAs per SM at 40 registers
            // 0.89 ms against 0.77 ms (VELES_BM25_SUB_OCC=12 / 8 select the other builds, for experiments)
            const int occ = std::getenv("VELES_BM25_SUB_OCC") ? std::atoi(std::getenv("VELES_BM25_SUB_OCC")) : 10;
            auto kern = occ == 12  ? (k <= 32 ? bm25_sub_kernel<1, 12> : k <= 64 ? bm25_sub_kernel<2, 12> : bm25_sub_kernel<4, 12>)
                        : occ == 8 ? (k <= 32 ? bm25_sub_kernel<1, 8> : k <= 64 ? bm25_sub_kernel<2, 8> : bm25_sub_kernel<4, 8>)
                                   : (k <= 32 ? bm25_sub_kernel<1, 10> : k <= 64 ? bm25_sub_kernel<2, 10> : bm25_sub_kernel<4, 10>);
            const size_t smem_s = (size_t)kSubWarps * kFine * 4 + (size_t)kSubWarps * k * 8;  // accumulators + output staging
            VELES_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
            VELES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSubWarps * 32, smem_s));
            VELES_REQUIRE(per_sm >= 1, "bm25 query kernel does not fit on an SM");
            const uint32_t resident_warps = (uint32_t)(per_sm * sms) * kSubWarps;
            // parts per query: ~4 work items per resident warp, so that the heavy queries do not set the finish time and
            // a small batch still fills the GPU (measured on B200, 1024 queries: 12 parts 0.84 ms, 24 parts 0.77 ms, 56 parts
            // 0.79 ms); at most 64 (the last part of a query merges them serially)
            uint32_t parts = std::max<uint32_t>(1, std::min<uint32_t>(std::min<uint32_t>(ix->n_fine, 64), (4 * resident_warps + nq - 1) / nq));
            if (const char* e = std::getenv("VELES_BM25_PARTS")) parts = std::max(1, std::min<int>((int)ix->n_fine, std::atoi(e)));
            const size_t tickets_off = 256, thr_off = (tickets_off + (size_t)nq * 4 + 255) / 256 * 256;
            const size_t lists_off = (thr_off + (size_t)nq * 4 + 255) / 256 * 256;
            VELES_TRY(ix->partial_d.ensure(lists_off + (size_t)nq * parts * k * 8));
            VELES_CUDA(cudaMemsetAsync(ix->partial_d.p, 0, lists_off, st));
            uint8_t* pb = ix->partial_d.as<uint8_t>();
            const uint64_t items = (uint64_t)nq * parts;
            const uint32_t grid = (uint32_t)std::min<uint64_t>((items + kSubWarps - 1) / kSubWarps, (uint64_t)per_sm * sms);
            kern<<<grid, kSubWarps * 32, smem_s, st>>>(v, ix->q_ptr_d.as<uint32_t>(), ix->q_terms_d.as<uint32_t>(), nq, k, parts,
                                                       reinterpret_cast<uint64_t*>(pb + lists_off), reinterpret_cast<uint32_t*>(pb + tickets_off),
                                                       reinterpret_cast<uint32_t*>(pb + thr_off), ix->out_doc_d.as<uint32_t>(),
                                                       ix->out_score_d.as<float>(), ix->out_cnt_d.as<uint32_t>(),
                                                       reinterpret_cast<uint32_t*>(pb));
        } else if (flat) {
            // 4 CTAs per SM: 64 registers, no spills; VELES_BM25_FLAT_OCC=5 selects the 48-register build (experiments)
            const bool occ5 = std::getenv("VELES_BM25_FLAT_OCC") != nullptr && std::atoi(std::getenv("VELES_BM25_FLAT_OCC")) == 5;
            const size_t smem = (size_t)kRange * 4 + (size_t)9 * k * 8;  // accumulator, eight lists, the merged list
            auto kern = occ5 ? (k <= 32 ? bm25_flat_kernel<1, 5> : k <= 64 ? bm25_flat_kernel<2, 5> : bm25_flat_kernel<4, 5>)
                             : (k <= 32 ? bm25_flat_kernel<1, 4> : k <= 64 ? bm25_flat_kernel<2, 4> : bm25_flat_kernel<4, 4>);
            VELES_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            VELES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
            VELES_REQUIRE(per_sm >= 1, "bm25 query kernel does not fit on an SM");
            const uint32_t resident = (uint32_t)(per_sm * sms);
            // parts per query: at least ~6 waves of work items, so that the last wave costs little and a small batch
            // still fills the GPU (env override for experiments)
            uint32_t parts = std::max<uint32_t>(1, std::min<uint32_t>(ix->n_ranges, (6 * resident + nq - 1) / nq));
            if (const char* e = std::getenv("VELES_BM25_PARTS")) parts = std::max(1, std::min<int>((int)ix->n_ranges, std::atoi(e)));
            const size_t tickets_off = 256, thr_off = (tickets_off + (size_t)nq * 4 + 255) / 256 * 256;
            const size_t lists_off = (thr_off + (size_t)nq * 4 + 255) / 256 * 256;
            VELES_TRY(ix->partial_d.ensure(lists_off + (size_t)nq * parts * k * 8));
            VELES_CUDA(cudaMemsetAsync(ix->partial_d.p, 0, lists_off, st));
            uint8_t* pb = ix->partial_d.as<uint8_t>();
            const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)nq * parts, resident);
            kern<<<grid, 256, smem, st>>>(v, ix->q_ptr_d.as<uint32_t>(), ix->q_terms_d.as<uint32_t>(), nq, k, parts,
                                          reinterpret_cast<uint64_t*>(pb + lists_off), reinterpret_cast<uint32_t*>(pb + tickets_off),
                                          reinterpret_cast<uint32_t*>(pb + thr_off),
                                          ix->out_doc_d.as<uint32_t>(), ix->out_score_d.as<float>(), ix->out_cnt_d.as<uint32_t>(),
                                          reinterpret_cast<uint32_t*>(pb));
        } else {
            auto kern = hash    ? (k <= 32 ? bm25_hash_kernel<1> : k <= 64 ? bm25_hash_kernel<2> : bm25_hash_kernel<4>)
                        : pre   ? (k <= 32 ? bm25_prefetch_kernel<1> : k <= 64 ? bm25_prefetch_kernel<2> : bm25_prefetch_kernel<4>)
                        : slice ? (k <= 32 ? bm25_slice_kernel<1> : k <= 64 ? bm25_slice_kernel<2> : bm25_slice_kernel<4>)
                                : (k <= 32 ? bm25_query_kernel<1> : k <= 64 ? bm25_query_kernel<2> : bm25_query_kernel<4>);
            VELES_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            VELES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
            VELES_REQUIRE(per_sm >= 1, "bm25 query kernel does not fit on an SM");
            VELES_TRY(ix->partial_d.ensure(64));
            VELES_CUDA(cudaMemsetAsync(ix->partial_d.p, 0, 4, st));
            const uint32_t grid = std::min<uint32_t>(nq, (uint32_t)(per_sm * sms));
            kern<<<grid, 256, smem, st>>>(v, ix->q_ptr_d.as<uint32_t>(), ix->q_terms_d.as<uint32_t>(), nq, k,
                                          ix->out_doc_d.as<uint32_t>(), ix->out_score_d.as<float>(), ix->out_cnt_d.as<uint32_t>(),
                                          ix->partial_d.as<uint32_t>());
        }
        count_launch();
        VELES_CUDA(cudaGetLastError());
    } else {
        const size_t smem1 = (size_t)kRange * 4 + (size_t)k * 8 * (k <= kMultiK ? 8 : 1);
        VELES_CUDA(cudaFuncSetAttribute(bm25_range_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        VELES_CUDA(cudaFuncSetAttribute(bm25_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(k * 8)));
        // gridDim.y <= 65535: chunk the queries; the partial buffer is bounded to ~512 MiB per pass
        const size_t per_q = (size_t)ix->n_ranges * k * 8;
        const uint32_t chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(nq, 32768), ((size_t)512 << 20) / per_q));
        VELES_TRY(ix->partial_d.ensure((size_t)chunk * per_q));
        for (uint32_t q0 = 0; q0 < nq; q0 += chunk) {
            const uint32_t nn = std::min(chunk, nq - q0);
            dim3 grid(ix->n_ranges, nn);
            bm25_range_kernel<<<grid, 256, smem1, st>>>(v, ix->q_ptr_d.as<uint32_t>() + q0, ix->q_terms_d.as<uint32_t>(), k,
                                                        ix->partial_d.as<uint64_t>());
            bm25_merge_kernel<<<nn, 32, (size_t)k * 8, st>>>(ix->partial_d.as<uint64_t>(), ix->n_ranges, k,
                                                             ix->out_doc_d.as<uint32_t>() + (size_t)q0 * k,
                                                             ix->out_score_d.as<float>() + (size_t)q0 * k,
                                                             ix->out_cnt_d.as<uint32_t>() + q0);
            count_launch(2);
            VELES_CUDA(cudaGetLastError());
        }
    }
    return VELES_OK;
}

// upload of a batch's term lists + bm25_enqueue; the caller holds ix->mu and has checked the arguments
static int32_t bm25_stage_and_enqueue(const veles_bm25* ix, const uint32_t* q_term_ptr, const uint32_t* q_terms, uint32_t nq,
                                      uint32_t k, cudaStream_t st) {
    const uint32_t n_qterms = q_term_ptr[nq];
    VELES_TRY(ix->q_ptr_d.ensure(((size_t)nq + 1) * 4));
    VELES_TRY(ix->q_terms_d.ensure(std::max<size_t>((size_t)n_qterms * 4, 16)));
    VELES_TRY(ix->out_doc_d.ensure((size_t)nq * k * 4));
    VELES_TRY(ix->out_score_d.ensure((size_t)nq * k * 4));
    VELES_TRY(ix->out_cnt_d.ensure((size_t)nq * 4));
    VELES_CUDA(cudaMemcpyAsync(ix->q_ptr_d.p, q_term_ptr, ((size_t)nq + 1) * 4, cudaMemcpyHostToDevice, st));
    if (n_qterms) VELES_CUDA(cudaMemcpyAsync(ix->q_terms_d.p, q_terms, (size_t)n_qterms * 4, cudaMemcpyHostToDevice, st));
    uint32_t max_terms = 0;
    for (uint32_t i = 0; i < nq; ++i) max_terms = std::max(max_terms, q_term_ptr[i + 1] - q_term_ptr[i]);
    return bm25_enqueue(ix, nq, k, max_terms, st);
}

int32_t veles_bm25_search_batch(const veles_bm25_t* ix, const uint32_t* q_term_ptr, const uint32_t* q_terms, uint32_t nq,
                                uint32_t k, uint32_t* out_doc, float* out_score, uint32_t* out_counts, void* stream) {
    VELES_REQUIRE(ix != nullptr, "index is NULL");
    VELES_REQUIRE(nq == 0 || (q_term_ptr && out_doc && out_score && out_counts), "NULL buffer");
    VELES_REQUIRE(k >= 1 && k <= 4096, "k must be in 1..4096, got %u", k);
    if (nq == 0) return VELES_OK;
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> g(ix->mu);
    const uint32_t n_qterms = q_term_ptr[nq];
    VELES_REQUIRE(n_qterms == 0 || q_terms, "q_terms is NULL");
    // Bm25Index::search returns nothing for an empty index (bm25.rs:275-278)
    if (ix->doc_count == 0 || ix->n_postings == 0) {
        for (uint32_t i = 0; i < nq; ++i) out_counts[i] = 0;
        for (size_t i = 0; i < (size_t)nq * k; ++i) {
            out_doc[i] = VELES_INVALID_ID;
            out_score[i] = std::nanf("");
        }
        return VELES_OK;
    }
    VELES_TRY(bm25_stage_and_enqueue(ix, q_term_ptr, q_terms, nq, k, st));
    VELES_CUDA(cudaMemcpyAsync(out_doc, ix->out_doc_d.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaMemcpyAsync(out_score, ix->out_score_d.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaMemcpyAsync(out_counts, ix->out_cnt_d.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    VELES_CUDA(cudaStreamSynchronize(st));
    return VELES_OK;
}


// Collection::hybrid_search (collection/search/text.rs:113-203) for a batch, in one call: the vector leg
// (self.index.search(query, 2k): NativeHnsw::search with the caller's ef) on `stream`, the text leg
// (self.text_index.search(text, 2k)) concurrently on a stream of the BM25 snapshot -- the traversal is HBM-bound on
// one warp per query, the posting scan issue-bound, so the text leg runs in the vector leg's shadow -- then the RRF
// (text.rs:133-180) over the two device-resident lists and ONE copy back.  Same kernels as veles_search_batch,
// veles_bm25_search_batch and veles_rrf_hybrid called one after the other, hence the same bits, without their three
// host round trips.  Ids: node ids of `idx` are the document ids of `bm` (the storage fetch, text.rs:183-203, stays
// with the host).
int32_t veles_hybrid_search_batch(const veles_index_t* idx, const veles_bm25_t* bm, const float* queries, const uint32_t* q_term_ptr,
                                  const uint32_t* q_terms, uint32_t nq, uint32_t k, uint32_t ef, float vector_weight,
                                  uint32_t* out_ids, float* out_score, uint32_t* out_counts, void* stream) {
    VELES_REQUIRE(idx != nullptr && bm != nullptr, "index is NULL");
    VELES_REQUIRE(nq == 0 || (queries && q_term_ptr && out_ids && out_score && out_counts), "NULL buffer");
    VELES_REQUIRE(k >= 1 && k <= 2048, "k must be in 1..2048 (both legs fetch 2k), got %u", k);
    if (nq == 0) return VELES_OK;
    VELES_REQUIRE(q_term_ptr[nq] == 0 || q_terms, "q_terms is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t in_k = 2 * k;  // text.rs:137-140
    std::lock_guard<std::mutex> g(bm->mu);
    if (!bm->side) {
        VELES_CUDA(cudaStreamCreateWithFlags(&bm->side, cudaStreamNonBlocking));
        VELES_CUDA(cudaEventCreateWithFlags(&bm->side_done, cudaEventDisableTiming));
        VELES_CUDA(cudaEventCreateWithFlags(&bm->side_fork, cudaEventDisableTiming));
    }
    SearchCtx* ctx = nullptr;
    {
        std::lock_guard<std::mutex> gi(idx->mu);
        VELES_TRY(acquire_ctx(idx, st, true, &ctx));
    }
    auto body = [&]() -> int32_t {
        // vector leg first: its grid (one CTA per query, every query resident) takes the SMs, the text leg fills in
        const size_t qb = (size_t)nq * idx->dim * 4, lb = (size_t)nq * in_k * 4, ob = (size_t)nq * k * 4;
        VELES_TRY(ctx->q_d.ensure(qb));
        VELES_TRY(ctx->ids_d.ensure(lb));
        VELES_TRY(ctx->val_d.ensure(lb));
        VELES_TRY(ctx->cnt_d.ensure((size_t)nq * 4));
        VELES_TRY(bm->hy_ids_d.ensure(ob));
        VELES_TRY(bm->hy_score_d.ensure(ob));
        VELES_TRY(bm->hy_cnt_d.ensure((size_t)nq * 4));
        VELES_CUDA(cudaMemcpyAsync(ctx->q_d.p, queries, qb, cudaMemcpyHostToDevice, st));
        // the text leg may start once the queries are on the device, i.e. together with the traversal behind the copy
        // (not during the copy: its CTAs would take the SMs first and the traversal would start thin)
        VELES_CUDA(cudaEventRecord(bm->side_fork, st));
        VELES_CUDA(cudaStreamWaitEvent(bm->side, bm->side_fork, 0));
        VELES_TRY(launch_search(idx, idx->view(), ctx, ctx->q_d.as<float>(), nq, in_k, ef, ctx->ids_d.as<uint32_t>(),
                                ctx->val_d.as<float>(), ctx->cnt_d.as<uint32_t>(), nullptr, st));
        // text leg; Bm25Index::search returns nothing for an empty index (bm25.rs:275-278)
        VELES_TRY(bm->out_doc_d.ensure(lb));
        VELES_TRY(bm->out_cnt_d.ensure((size_t)nq * 4));
        if (bm->doc_count == 0 || bm->n_postings == 0)
            VELES_CUDA(cudaMemsetAsync(bm->out_cnt_d.p, 0, (size_t)nq * 4, bm->side));
        else
            VELES_TRY(bm25_stage_and_enqueue(bm, q_term_ptr, q_terms, nq, in_k, bm->side));
        VELES_CUDA(cudaEventRecord(bm->side_done, bm->side));
        VELES_CUDA(cudaStreamWaitEvent(st, bm->side_done, 0));
        VELES_TRY(rrf_hybrid_enqueue_d(ctx->ids_d.as<uint32_t>(), ctx->cnt_d.as<uint32_t>(), bm->out_doc_d.as<uint32_t>(),
                                       bm->out_cnt_d.as<uint32_t>(), nq, in_k, vector_weight, k, bm->hy_ids_d.as<uint32_t>(),
                                       bm->hy_score_d.as<float>(), bm->hy_cnt_d.as<uint32_t>(), st));
        VELES_CUDA(cudaMemcpyAsync(out_ids, bm->hy_ids_d.p, ob, cudaMemcpyDeviceToHost, st));
        VELES_CUDA(cudaMemcpyAsync(out_score, bm->hy_score_d.p, ob, cudaMemcpyDeviceToHost, st));
        VELES_CUDA(cudaMemcpyAsync(out_counts, bm->hy_cnt_d.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
        return check_search_error_flag(ctx, st);  // synchronises `st` (and with it the joined text leg)
    };
    const int32_t rc = body();
    if (rc != VELES_OK) cudaStreamSynchronize(bm->side);  // nothing of this call may still run when the lock is dropped
    release_ctx(idx, ctx);
    return rc;
}

}  // extern "C"

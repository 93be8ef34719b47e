/* veles_b200.h -- C ABI of the B200-native VelesDB vector-search hot path.
 *
 * This is the drop-in boundary.  Each entry point names the
 * reference interface it replaces; paths are relative to
 * /root/reference/crates/velesdb-core/src.  The Rust side (`HnswIndex`,
 * `Bm25Index`, `Collection::hybrid_search`) keeps its signatures and binds these
 * through `extern "C"` (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - every function returns int32 status: 0 = ok, < 0 = error (veles_status);
 *     veles_last_error() returns a thread-local message.  Nothing throws.
 *   - pointers are HOST pointers unless the parameter name ends in `_d`
 *     (device pointers, for callers that already hold data in HBM).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Host-pointer calls copy H2D, launch, copy D2H and synchronise `stream`
 *     before returning.  `_d` calls only enqueue work on `stream`.
 *   - node ids are internal insertion indices (u32); external u64 ids, the
 *     tombstone filter and transform_score stay with the caller exactly as
 *     index/hnsw/index/search.rs:86-91 does today.
 *   - there is no CPU fallback: without a CUDA device every call fails with
 *     VELES_ERR_CUDA.
 */
#ifndef VELES_B200_H
#define VELES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum veles_status {
    VELES_OK = 0,
    VELES_ERR_INVALID = -1,   /* bad argument (null pointer, k == 0, dim mismatch ...) */
    VELES_ERR_CUDA = -2,      /* CUDA runtime error; message in veles_last_error()   */
    VELES_ERR_IO = -3,        /* file missing / truncated / wrong version            */
    VELES_ERR_OOM = -4,       /* device or host allocation failed                    */
    VELES_ERR_OVERFLOW = -5,  /* an internal bounded structure overflowed            */
    VELES_ERR_UNSUPPORTED = -6
} veles_status;

/* = DistanceMetric as u8 (core/distance.rs:16-39, index/hnsw/index/constructors.rs:204-217) */
typedef enum veles_metric {
    VELES_COSINE = 0,
    VELES_EUCLIDEAN = 1,
    VELES_DOT = 2,
    VELES_HAMMING = 3,
    VELES_JACCARD = 4
} veles_metric;

/* storage type of the vectors in HBM */
typedef enum veles_dtype {
    VELES_F32 = 0,  /* reference layout: Vec<Vec<f32>> (native/graph.rs:22)                        */
    VELES_F16 = 1,  /* half::f16 round-to-nearest-even (core/half_precision.rs:97), f32 accumulate  */
    VELES_BIN1 = 2, /* packed bits, LSB-first in u64 words (simd_explicit.rs:308-360); Hamming only */
    VELES_SQ8 = 3   /* u8 codes of ScalarQuantizer (native/quantization.rs:236-250): only as the traversal store
                       attached by veles_index_attach_sq8, never as store_dtype                      */
} veles_dtype;

/* SearchQuality (index/hnsw/params.rs:283-320) */
typedef enum veles_quality {
    VELES_FAST = 0,
    VELES_BALANCED = 1,
    VELES_ACCURATE = 2,
    VELES_PERFECT = 3,
    VELES_CUSTOM = 4
} veles_quality;

#define VELES_INVALID_ID 0xFFFFFFFFu

typedef struct veles_index veles_index_t; /* immutable device snapshot of one NativeHnsw */
typedef struct veles_bm25 veles_bm25_t;   /* immutable device snapshot of one Bm25Index  */

/* ---- runtime ------------------------------------------------------------------------- */
int32_t veles_init(int32_t device);
int32_t veles_shutdown(void);
const char* veles_last_error(void);
const char* veles_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t veles_launch_count(void);

/* SearchQuality::ef_search (params.rs:309-319) */
uint64_t veles_ef_search(int32_t quality, uint64_t k, uint64_t custom_ef);
/* NativeHnsw::transform_score (native/backend_adapter.rs:160-168) */
float veles_transform_score(int32_t metric, float raw_distance);

/* ---- snapshot construction ------------------------------------------------------------ */
/* NativeHnsw::file_load (native/backend_adapter.rs:274-380): parses `{dir}/{basename}.vectors`
 * and `.graph` (format v1, little endian) and uploads them.  `store_dtype` selects the HBM
 * storage type (F32 keeps the file's bits; F16 rounds once at load). */
int32_t veles_index_from_reference_files(const char* dir, const char* basename, int32_t metric, int32_t store_dtype,
                                         veles_index_t** out);

/* The same snapshot from in-memory arrays -- what a Rust `HnswIndex` hands over after
 * `insert`/`load` (native/graph.rs:18-44 fields).  `vectors`: n*dim f32 (dtype F32/F16) or
 * n*(dim/64) u64 words (dtype BIN1, `dim` = number of bits).  Layer `l` is CSR:
 * row_ptr[l][0..layer_nodes[l]] and cols[l]; rows of nodes >= layer_nodes[l] are empty
 * (Layer::get_neighbors, native/layer.rs:33-39).  n == 0 gives an empty index. */
int32_t veles_index_from_arrays(const void* vectors, uint64_t n, uint32_t dim, int32_t src_dtype, int32_t store_dtype,
                                int32_t metric, uint32_t num_layers, const uint64_t* const* row_ptr,
                                const uint32_t* const* cols, const uint64_t* layer_nodes, uint32_t M, uint32_t M0,
                                uint64_t entry_point, uint32_t max_layer, veles_index_t** out);

/* A snapshot with vectors only (no graph): enough for veles_bruteforce_batch, and the input
 * of veles_index_build_graph. */
int32_t veles_index_from_vectors(const void* vectors, uint64_t n, uint32_t dim, int32_t src_dtype, int32_t store_dtype,
                                 int32_t metric, veles_index_t** out);
/* The same for hosts whose vectors are already on the device (or are produced there): an empty snapshot of n
 * zero rows, then rows [first, first + count) from device memory -- src_dtype F32 (converted like
 * veles_index_from_vectors does) or the store type itself.  veles_index_get_rows copies rows back to the host in the
 * store type (what HnswIndex::load leaves out: ShardedVectors stays empty after load, constructors.rs:240, but the
 * graph's own vectors are there, native/backend_adapter.rs:286-307). */
int32_t veles_index_create(uint64_t n, uint32_t dim, int32_t store_dtype, int32_t metric, veles_index_t** out);
int32_t veles_index_set_rows_d(veles_index_t* idx, uint64_t first, uint64_t count, const void* rows_d, int32_t src_dtype,
                               void* stream);
int32_t veles_index_get_rows(const veles_index_t* idx, uint64_t first, uint64_t count, void* out);
int32_t veles_index_free(veles_index_t* idx);

uint64_t veles_index_len(const veles_index_t* idx);       /* NativeHnsw::len  graph.rs:130 */
uint32_t veles_index_dim(const veles_index_t* idx);       /* HnswIndex::dimension          */
int32_t veles_index_metric(const veles_index_t* idx);     /* HnswIndex::metric             */
uint32_t veles_index_max_layer(const veles_index_t* idx);
uint64_t veles_index_entry_point(const veles_index_t* idx);
uint64_t veles_index_device_bytes(const veles_index_t* idx);

/* NativeHnsw::file_dump (native/backend_adapter.rs:184-261): writes the snapshot back in
 * format v1 (vectors are written as f32). */
int32_t veles_index_dump(const veles_index_t* idx, const char* dir, const char* basename);
/* only `{basename}.graph` (native/backend_adapter.rs:213-261), for snapshots of any storage type */
int32_t veles_index_dump_graph(const veles_index_t* idx, const char* dir, const char* basename);
/* copies the layer-`l` adjacency out as CSR (row_ptr may be NULL to query sizes only) */
int32_t veles_index_export_layer(const veles_index_t* idx, uint32_t layer, uint64_t* out_nodes, uint64_t* out_edges,
                                 uint64_t* row_ptr, uint32_t* cols);

/* ---- HNSW search ------------------------------------------------------------------------ */
/* NativeHnsw::search (native/graph.rs:251-270) for a batch of queries, i.e. the body of
 * HnswIndex::search_batch_parallel (index/hnsw/index/batch.rs:159-197) before id mapping:
 * greedy descent (search_layer_single, graph.rs:405-428) then the ef-bounded beam on layer 0
 * (search_layer, graph.rs:438-520), first k of the result.
 *   queries       nq*dim f32
 *   out_node_ids  nq*k, padded with VELES_INVALID_ID
 *   out_raw_dist  nq*k in-graph distance (native/distance.rs:75-85), padded with NaN
 *   out_counts    nq   number of valid results (<= k)
 *   out_stats     nq*4 u32, optional (NULL): {layer-0 distance evaluations, layer-0 expansions,
 *                 upper-layer distance evaluations, upper-layer adjacency scans}
 * Order: ascending by (distance in IEEE total order, node id).  ef is used as given; apply
 * veles_ef_search first to follow a SearchQuality. */
int32_t veles_search_batch(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k, uint32_t ef,
                           uint32_t* out_node_ids, float* out_raw_dist, uint32_t* out_counts, uint32_t* out_stats,
                           void* stream);
int32_t veles_search_batch_d(const veles_index_t* idx, const float* queries_d, uint32_t nq, uint32_t k, uint32_t ef,
                             uint32_t* out_node_ids_d, float* out_raw_dist_d, uint32_t* out_counts_d,
                             uint32_t* out_stats_d, void* stream);

/* Outcome of the `_d` searches enqueued on `stream` so far.  `_d` calls only enqueue work, so a tie-list overflow
 * (VELES_ERR_OVERFLOW: more than 4096 evicted candidates tied with the worst result of one query) cannot be returned
 * by them; this call waits for `stream` and reports -- and clears -- it.  Every launch has its own scratch (visited
 * bitmaps, tie lists, work counter), so `_d` calls on different streams and host-pointer calls from different host
 * threads run concurrently on one index, as the reference's searches do under its read lock
 * (index/hnsw/index/search.rs:59-94; HnswIndex is Send + Sync). */
int32_t veles_search_status(const veles_index_t* idx, void* stream);

/* Pipelined form of veles_search_batch for a serving loop (the rayon pool of
 * HnswIndex::search_batch_parallel, index/hnsw/index/batch.rs:159-197, keeps every core busy across batches; this
 * keeps the GPU busy across batches).  veles_search_submit enqueues the batch's H2D copy, the search and the D2H
 * copies on a stream owned by the library and returns a ticket; veles_search_wait blocks until that batch's outputs
 * are in the host buffers and returns its status.  Several tickets (up to 16) may be in flight: batch i+1's copies
 * and first queries overlap batch i's last queries and copies back.  Buffers must stay valid until the wait; pinned
 * host memory makes the copies asynchronous.  Results are identical to veles_search_batch. */
int32_t veles_search_submit(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k, uint32_t ef,
                            uint32_t* out_node_ids, float* out_raw_dist, uint32_t* out_counts, uint64_t* ticket);
int32_t veles_search_wait(const veles_index_t* idx, uint64_t ticket);

/* ---- brute force -------------------------------------------------------------------------- */
/* HnswIndex::search_brute_force / brute_force_search_parallel (index/hnsw/index/search.rs:176-219,
 * batch.rs:223-244) for a batch: metric value (compute_distance, search.rs:30-38) against every
 * stored vector, ordered by DistanceMetric::sort_results (core/distance.rs:95-103) with ties by
 * ascending node id, first k.  out_ids nq*k (VELES_INVALID_ID pad), out_score nq*k (NaN pad). */
int32_t veles_bruteforce_batch(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k,
                               uint32_t* out_ids, float* out_score, void* stream);
int32_t veles_bruteforce_batch_d(const veles_index_t* idx, const float* queries_d, uint32_t nq, uint32_t k,
                                 uint32_t* out_ids_d, float* out_score_d, void* stream);
/* Relaxed-mode brute force for large query batches: the candidate stage runs on the tensor cores, the final scores
 * are exact.  An fp16 copy of the collection (built on first use) is multiplied with the fp16 queries by a
 * hand-written tcgen05 GEMM (accumulators in TMEM, operands staged by TMA); its epilogue keeps the rows whose fp16
 * score reaches a per-query threshold -- the (k * oversample)-th best score of a sample of the collection, which can
 * only be looser than the true one -- and the exact metric value (compute_distance, index/hnsw/index/search.rs:30-38,
 * reference summation order) of those candidates decides the top k.  This is the role of the reference's GPU hook
 * search_brute_force_gpu (search.rs:229-279 -> gpu/gpu_backend.rs:300-340) and of the re-rank in search_with_rerank
 * (search.rs:118-160).  Ordering and padding as veles_bruteforce_batch; results equal it unless fp16 rounding pushes a
 * true neighbour below rank k * oversample (judged by recall; cosine, euclidean and dot; f32 or f16 storage).
 * gemm_ms (may be NULL): device time of the two GEMM passes, for the roofline. */
int32_t veles_bruteforce_batch_relaxed(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k,
                                       uint32_t oversample, uint32_t* out_ids, float* out_score, void* stream);
int32_t veles_bruteforce_batch_relaxed_d(const veles_index_t* idx, const float* queries_d, uint32_t nq, uint32_t k,
                                         uint32_t oversample, uint32_t* out_ids_d, float* out_score_d, float* gemm_ms,
                                         void* stream);
/* exact metric value of explicit (query, node) pairs: the re-rank step of
 * HnswIndex::search_with_rerank (index/hnsw/index/search.rs:118-160).  cand is nq*m node ids
 * (VELES_INVALID_ID entries give NaN). */
int32_t veles_rerank_batch(const veles_index_t* idx, const float* queries, uint32_t nq, const uint32_t* cand,
                           uint32_t m, float* out_score, void* stream);
/* DistanceEngine::batch_distance (native/distance.rs:22-24, 88-103): in-graph distance of one
 * query to explicit host vectors; used by the parity tests of the distance kernels. */
int32_t veles_distance_pairs(int32_t metric, const float* a, const float* b, uint32_t n_pairs, uint32_t dim,
                             int32_t as_metric_value, float* out, void* stream);

/* ---- BM25 ------------------------------------------------------------------------------------ */
/* Device snapshot of a Bm25Index (index/bm25.rs:78-90).  Tokenisation (bm25.rs:114-120) and the
 * string->term-id dictionary stay on the host.  Postings are CSR by term, doc ids ascending
 * within a term, with the term frequency beside each doc (the reference keeps tf in per-document
 * maps, bm25.rs:62-67).  doc_len is indexed by doc id (0 = absent).  df[t] is the posting-list
 * length the reference would report (PostingList::len), which can exceed the live postings after
 * a document was replaced (bm25.rs:188-196 keeps stale postings).
 * Device memory: 8 bytes per posting for the query path ((doc, contribution) with the contribution
 * precomputed in the reference's f32 operation order), 12 more for the long-query fallback, a
 * coarse skip table (8 bytes x terms x docs/7168) and, when it stays below 2 GiB and 4x the
 * postings, a fine one (4 bytes x terms x docs/1024) that the default query kernel needs; without
 * it queries go through the coarse-table kernel (same results, ~1.6x slower on the bench corpus). */
int32_t veles_bm25_from_csr(uint32_t n_terms, const uint64_t* term_ptr, const uint32_t* post_doc,
                            const uint32_t* post_tf, const uint32_t* df, uint32_t n_doc_slots, const uint32_t* doc_len,
                            uint64_t doc_count, uint64_t total_len, float k1, float b, veles_bm25_t** out);
int32_t veles_bm25_free(veles_bm25_t* ix);
/* Bm25Index::search (bm25.rs:269-341) for a batch.  q_term_ptr: nq+1 offsets into q_terms (term
 * ids in query-token order, duplicates allowed, VELES_INVALID_ID = term not in the dictionary).
 * Order: score descending (total order), ties by ascending doc id.  out_doc nq*k
 * (VELES_INVALID_ID pad), out_score nq*k (NaN pad), out_counts nq. */
int32_t veles_bm25_search_batch(const veles_bm25_t* ix, const uint32_t* q_term_ptr, const uint32_t* q_terms,
                                uint32_t nq, uint32_t k, uint32_t* out_doc, float* out_score, uint32_t* out_counts,
                                void* stream);

/* ---- fusion ----------------------------------------------------------------------------------- */
/* The RRF of Collection::hybrid_search (collection/search/text.rs:133-180) for a batch: per query
 * two ranked id lists (vector first, then text), score += w/(rank0+60) and (1-w)/(rank0+60), top-k
 * by (score, id) keeping the largest, written score-descending (equal scores: larger id first).
 * vec_ids / txt_ids are nq*in_k with per-query valid counts. */
int32_t veles_rrf_hybrid(const uint32_t* vec_ids, const uint32_t* vec_cnt, const uint32_t* txt_ids,
                         const uint32_t* txt_cnt, uint32_t nq, uint32_t in_k, float vector_weight, uint32_t k,
                         uint32_t* out_ids, float* out_score, uint32_t* out_counts, void* stream);
/* Collection::hybrid_search (collection/search/text.rs:113-203, minus the storage fetch of lines 183-203) for a
 * batch, in ONE call: `self.index.search(vector_query, 2k)` (NativeHnsw::search with the caller's ef; the reference
 * passes ef_search(Balanced, 2k)) on `stream`, `self.text_index.search(text_query, 2k)` concurrently on a stream
 * owned by `bm`, then the RRF of lines 133-180 over the two device-resident lists and one copy back.  Same kernels,
 * hence the same bits, as veles_search_batch + veles_bm25_search_batch + veles_rrf_hybrid called in sequence, without
 * their host round trips.  Node ids of `idx` are the document ids of `bm`.  queries nq*dim f32; q_term_ptr / q_terms
 * as veles_bm25_search_batch; out_ids / out_score nq*k, out_counts nq (as veles_rrf_hybrid).  Host pointers;
 * returns when the results are in the output buffers.  VELES_ERR_OVERFLOW as veles_search_batch. */
int32_t veles_hybrid_search_batch(const veles_index_t* idx, const veles_bm25_t* bm, const float* queries,
                                  const uint32_t* q_term_ptr, const uint32_t* q_terms, uint32_t nq, uint32_t k, uint32_t ef,
                                  float vector_weight, uint32_t* out_ids, float* out_score, uint32_t* out_counts,
                                  void* stream);
/* FusionStrategy::fuse (fusion/strategy.rs:138-300) for one multi-query request: n_lists ranked
 * (id, score) lists.  strategy 0 Average, 1 Maximum, 2 RRF{k}, 3 Weighted{avg,max,hit}.  Output
 * sorted score-descending, ties by ascending id; at most `cap` entries. */
int32_t veles_fuse(int32_t strategy, const uint32_t* list_ptr, uint32_t n_lists, const uint32_t* ids,
                   const float* scores, uint32_t rrf_k, float avg_w, float max_w, float hit_w, uint32_t cap,
                   uint32_t* out_ids, float* out_score, uint32_t* out_count, void* stream);

/* ---- graph construction (the path's producer) ---------------------------- */
/* Bulk construction of the HNSW graph on the GPU for the vectors of `idx` -- the role of
 * HnswIndex::insert_batch_parallel (index/hnsw/index/batch.rs:82-108), whose result the reference
 * itself leaves order dependent (rayon).  Nodes are inserted in id order, a block at a time: every node of a
 * block runs NativeHnsw::insert's searches (graph.rs:190-223: greedy descent, then search_layer with
 * ef_construction on each of its layers) on the graph built so far -- layer 0 through the production search
 * kernel -- then select_neighbors (graph.rs:526-581) and add_bidirectional_connection (graph.rs:592-639: a
 * full row keeps its max_conn closest).  Levels follow the reference PRNG (graph.rs:368-403) in node-id
 * order, so level assignment and entry point equal the reference's.  O(N log N); no library calls.
 * ef_construction 0 = 200.  Replaces any graph held by `idx`. */
int32_t veles_index_build_graph(veles_index_t* idx, uint32_t M, uint32_t ef_construction, void* stream);

/* HnswIndex::insert_batch_parallel on a LIVE index (index/hnsw/index/batch.rs:82-108; the reference inserts into the
 * graph it already has): `count` more vectors become nodes n .. n+count-1 and are linked into the existing graph a
 * block at a time, exactly as the later blocks of veles_index_build_graph are -- no rebuild, no host copy of the old
 * vectors.  Works on built and on loaded graphs (M is the graph's; ef_construction 0 = the graph's, else 200).
 * Levels continue the reference PRNG in node-id order (graph.rs:368-403).  The SQ8 store and the id map must be
 * re-attached afterwards.  vectors: count*dim f32, or rows already in the store type. */
int32_t veles_index_append(veles_index_t* idx, const void* vectors, uint64_t count, int32_t src_dtype,
                           uint32_t ef_construction, void* stream);

/* The reference's *sequential* construction, exactly: NativeHnsw::insert (native/graph.rs:158-237) for nodes
 * 0..n-1 in id order -- what HnswIndex::insert (index/hnsw/index/trait_impl.rs:10-36) builds, the one path on
 * which the reference graph is deterministic.  Same levels (graph.rs:368-403), search_layer with
 * ef_construction, select_neighbors (graph.rs:526-581, alpha = 1), add_bidirectional_connection
 * (graph.rs:592-639).  One warp, latency bound: meant for small/medium graphs and for build parity; f32
 * storage only.  Equal distances are ordered by node id (the reference: heap-internal order). */
int32_t veles_index_build_graph_exact(veles_index_t* idx, uint32_t M, uint32_t ef_construction, void* stream);

/* ---- SQ8 dual precision -------------------------------------------------- */
/* DualPrecisionHnsw (native/dual_precision.rs:60-441): int8 traversal + exact f32 re-rank.
 *
 * veles_index_attach_sq8 trains the ScalarQuantizer (native/quantization.rs:190-233: per-dimension min/max,
 * scale = 255 / range or 1.0 when |range| < 1e-10, inv_scale = 1 / scale) on the first `train_count` vectors of
 * the snapshot -- the reference trains on its first min(1000, max_elements) inserts (dual_precision.rs:100,
 * 134-137) or on everything inserted so far at force_train_quantizer (:172-176) -- and quantizes every vector
 * (quantize, quantization.rs:236-250: ((v - min) * scale).round().clamp(0, 255) as u8) into a second, u8 store
 * on the device.  Needs an f32 snapshot (the re-rank reads the original vectors, as the reference does). */
int32_t veles_index_attach_sq8(veles_index_t* idx, uint64_t train_count, void* stream);
int32_t veles_index_has_sq8(const veles_index_t* idx);
/* Copies the quantizer (dim floats each) and, when `codes` is not NULL, the n * dim codes to the host. */
int32_t veles_index_sq8_export(const veles_index_t* idx, float* min_vals, float* scales, float* inv_scales,
                               uint8_t* codes);

/* DualPrecisionHnsw::search_with_config on the int8 path (dual_precision.rs:284-325): the query is quantized,
 * greedy_search_int8 (:407-441) descends the upper layers and search_layer_int8 (:327-405) runs the layer-0
 * beam with ef = max(ef_search, k * oversampling), both on distance_l2_quantized (u32 sum of squared code
 * differences, quantization.rs:42-92) whatever the index metric; the k * oversampling closest are re-ranked
 * with the exact f32 graph distance of the index metric, stably sorted by total_cmp, and cut to k.
 * Outputs as veles_search_batch, out_raw_dist = exact f32 distances.  Equal int8 distances are ordered by node
 * id (the reference: heap-array order); out_stats as veles_search_batch.  The reference's size gate
 * (min_index_size) and the use_int8_traversal switch stay with the host wrapper. */
int32_t veles_search_batch_sq8(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k, uint32_t ef_search,
                               uint32_t oversampling, uint32_t* out_node_ids, float* out_raw_dist, uint32_t* out_counts,
                               uint32_t* out_stats, void* stream);
int32_t veles_search_batch_sq8_d(const veles_index_t* idx, const float* queries_d, uint32_t nq, uint32_t k,
                                 uint32_t ef_search, uint32_t oversampling, uint32_t* out_node_ids_d,
                                 float* out_raw_dist_d, uint32_t* out_counts_d, uint32_t* out_stats_d, void* stream);

/* NativeHnsw::search_multi_entry (native/graph.rs:288-348): the layer-0 search starts from the greedy descent's
 * result plus up to three extra entry points per query (extra_entries: nq*3 node ids, VELES_INVALID_ID padded;
 * duplicates are dropped as `entry_points.contains` does).  The reference draws the extra ids from its shared
 * xorshift64 state (`state % count`, only when count > 10 and num_probes > 1); that state lives with the host
 * wrapper, which passes the ids in.  Needs ef >= 4.  Outputs as veles_search_batch. */
int32_t veles_search_batch_multi_entry(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k, uint32_t ef,
                                       const uint32_t* extra_entries, uint32_t* out_node_ids, float* out_raw_dist,
                                       uint32_t* out_counts, uint32_t* out_stats, void* stream);

/* ---- id map, tombstones, filtered search ---------------------------------- */
/* ShardedMappings on the device (index/hnsw/sharded_mappings.rs:32-39): ext_ids[node] is the external id of
 * node `node` (NULL = identity), live_bits has one bit per node, 0 for removed ids (HnswIndex::remove is a soft
 * delete, trait_impl.rs:54-58; NULL = all live).  Call again after remove() with the new bitmap. */
int32_t veles_index_set_id_map(veles_index_t* idx, const uint64_t* ext_ids, const uint32_t* live_bits);

/* The per-query body of HnswIndex::search_batch_parallel (index/hnsw/index/batch.rs:178-196) in one call:
 * NativeHnsw::search(query, k_fetch, ef), hits of removed nodes dropped (search.rs:86-91), node -> external id,
 * transform_score (native/backend_adapter.rs:160-168), first k_out kept, traversal order preserved.
 * With allow_bits (one bit per node, host pointer) it is also the candidate loop of
 * Collection::search_with_filter (collection/search/vector.rs:182-211): k_fetch = max(4k, k + 10) over-fetched
 * hits, filtered, take(k_out).  out_ids nq*k_out (padded with all-ones), out_scores nq*k_out (NaN padded). */
int32_t veles_search_batch_mapped(const veles_index_t* idx, const float* queries, uint32_t nq, uint32_t k_fetch,
                                  uint32_t k_out, uint32_t ef, const uint32_t* allow_bits, uint64_t* out_ids,
                                  float* out_scores, uint32_t* out_counts, void* stream);

/* ---- multi-GPU ----------------------------------------------------------------------------------- */
/* Queries shard by contiguous slices across the GPUs of one NVSwitch box (one process per GPU); the snapshot is
 * replicated.  The reference's one strategy is rayon over queries in one process
 * (index/hnsw/index/batch.rs:159-197); its per-query results simply land in one Vec.  Here the only exchange is the
 * final gather of [nq, k] ids + distances, and it is done by the library itself, fused into the search: each rank
 * owns a gather window in its HBM, peers map it through CUDA IPC, and the search kernel's epilogue stores every
 * finished query's top-k into slot `rank` of every window (peer stores over NVLink); a flag exchange closes the epoch.
 *
 *   veles_comm_create    allocates this rank's window for `nq_per_rank` x `k` results per rank and writes an opaque
 *                        handle blob of veles_comm_handle_bytes() bytes
 *   veles_comm_connect   takes all ranks' blobs (rank-major; the host moves them over any channel it has) and maps the
 *                        peers' windows
 *   veles_search_batch_gather_d   veles_search_batch_d + gather; collective (every rank, same order, same nq / k / ef);
 *                        only enqueues on `stream`; mid_event (cudaEvent_t or NULL) is recorded between the search
 *                        kernel and the flag exchange
 *   veles_comm_window    device pointers of the local window of the latest epoch: [world*nq*k] ids, distances,
 *                        [world*nq] counts = the whole batch in query order, valid once `stream` has run
 *   veles_comm_status    waits for `stream`; error if a peer never arrived */
typedef struct veles_comm veles_comm_t;
int32_t veles_comm_handle_bytes(void);
int32_t veles_comm_create(int32_t rank, int32_t world, uint32_t nq_per_rank, uint32_t k, veles_comm_t** out, void* handle_out);
int32_t veles_comm_connect(veles_comm_t* comm, const void* all_handles);
int32_t veles_search_batch_gather_d(const veles_index_t* idx, veles_comm_t* comm, const float* queries_d, uint32_t nq,
                                    uint32_t k, uint32_t ef, void* stream, void* mid_event);
int32_t veles_comm_window(veles_comm_t* comm, uint32_t** ids_d, float** dist_d, uint32_t** counts_d);
int32_t veles_comm_status(veles_comm_t* comm, void* stream);
int32_t veles_comm_destroy(veles_comm_t* comm);

#ifdef __cplusplus
}
#endif
#endif /* VELES_B200_H */

// host_mirror_test.cpp -- drives the C++ HnswIndex mirror (hnsw_index.hpp) the way the reference's own
// index tests do (crates/velesdb-core/src/index/hnsw/index_tests.rs); run on a GPU box by
// tests/test_gpu_host_mirror.py.  Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <set>

#include "dual_precision.hpp"
#include "hnsw_index.hpp"

using namespace veles::host;

#define STEP(msg) std::fprintf(stderr, "[host mirror] %s\n", msg)
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            return 1;                                                      \
        }                                                                  \
    } while (0)

int main(int argc, char** argv) {
    const std::string tmp = argc > 1 ? argv[1] : "/tmp";
    if (veles_init(0) != VELES_OK) {
        std::fprintf(stderr, "%s\n", veles_last_error());
        return 2;
    }
    const size_t dim = 64, n = 500, k = 10;
    // empty index (index_tests.rs: search on empty returns nothing)
    {
        HnswIndex e(dim, DistanceMetric::Cosine);
        CHECK(e.len() == 0 && e.is_empty());
        CHECK(e.search(std::vector<float>(dim, 1.0f), 5).empty());
    }
    STEP("empty ok");
    HnswIndex index(dim, DistanceMetric::Cosine);
    std::vector<std::vector<float>> data(n, std::vector<float>(dim));
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < dim; ++j) data[i][j] = std::sin((float)(i * dim + j) * 0.001f);
    for (size_t i = 0; i < n; ++i) index.insert(i, data[i]);
    index.insert(3, data[7]);  // duplicate id: silently skipped (trait_impl.rs:12-25)
    CHECK(index.len() == n && index.dimension() == dim && index.metric() == DistanceMetric::Cosine);
    // dimension mismatch = the reference's assert_eq! panic
    bool threw = false;
    try {
        index.search(std::vector<float>(dim + 1, 0.0f), 5);
    } catch (const std::invalid_argument&) {
        threw = true;
    }
    CHECK(threw);
    STEP("insert + mismatch ok");
    // recall >= 0.8 for Accurate vs brute force (index_tests.rs:1108-1159)
    std::vector<float> q(dim);
    for (size_t j = 0; j < dim; ++j) q[j] = std::sin((float)j * 0.001f);
    auto res = index.search_with_quality(q, k, SearchQuality::Accurate());
    auto exact = index.search_brute_force(q, k);
    CHECK(res.size() == k && exact.size() == k);
    std::set<uint64_t> gt;
    for (auto& h : exact) gt.insert(h.first);
    size_t hits = 0;
    for (auto& h : res) hits += gt.count(h.first);
    CHECK(hits >= 8);
    for (size_t i = 1; i < exact.size(); ++i) CHECK(exact[i - 1].second >= exact[i].second);  // cosine: descending
    for (auto& h : res) CHECK(h.second >= 0.0f && h.second <= 1.0f);                           // transform_score clamp
    STEP("recall ok");
    // batch == single (index_tests.rs:1018-1053 checks lengths; ids must agree too)
    std::vector<std::vector<float>> qs(data.begin(), data.begin() + 16);
    auto batch = index.search_batch_parallel(qs, k, SearchQuality::Balanced());
    CHECK(batch.size() == qs.size());
    for (size_t i = 0; i < qs.size(); ++i) {
        auto single = index.search_with_quality(qs[i], k, SearchQuality::Balanced());
        CHECK(single.size() == batch[i].size());
        for (size_t j = 0; j < single.size(); ++j) CHECK(single[j] == batch[i][j]);
        CHECK(batch[i][0].second >= 0.9999f);  // a stored vector's best hit is (numerically) itself
    }
    STEP("batch ok");
    // soft delete: the node stays in the graph, the id is never returned (trait_impl.rs:54-58, search.rs:86-91)
    CHECK(index.remove(0) && !index.remove(0));
    CHECK(index.len() == n - 1 && index.tombstone_count() == 1);
    for (auto& h : index.search(data[0], k)) CHECK(h.first != 0);
    for (auto& h : index.search_brute_force(data[0], k)) CHECK(h.first != 0);
    STEP("remove ok");
    // rerank returns metric values sorted by sort_results
    auto rr = index.search_with_rerank(q, 5, 50);
    CHECK(rr.size() == 5);
    for (size_t i = 1; i < rr.size(); ++i) CHECK(rr[i - 1].second >= rr[i].second);
    STEP("rerank ok");
    // save / load round trip (constructors.rs:190-287): same first hit, vectors absent after load
    index.save(tmp);
    HnswIndex* loaded = HnswIndex::load(tmp, dim, DistanceMetric::Cosine);
    CHECK(loaded->len() == n - 1);
    auto a = index.search_with_quality(q, k, SearchQuality::Balanced());
    auto b = loaded->search_with_quality(q, k, SearchQuality::Balanced());
    CHECK(a.size() == b.size());
    for (size_t i = 0; i < a.size(); ++i) CHECK(a[i] == b[i]);
    CHECK(loaded->search_with_rerank(q, 5, 50).empty());  // rerank finds no vectors after load (search.rs:130-137)
    delete loaded;
    STEP("save/load ok");
    // <= 100 vectors: exact brute-force path; Euclidean top-1 at the origin (index_tests.rs:1691-1712)
    HnswIndex small(16, DistanceMetric::Euclidean);
    for (uint64_t i = 0; i < 50; ++i) {
        std::vector<float> v(16);
        for (size_t j = 0; j < 16; ++j) v[j] = i == 0 ? 0.0f : std::sin((float)(i * 16 + j));
        small.insert(i, v);
    }
    auto s = small.search(std::vector<float>(16, 0.0f), 5);
    CHECK(s.size() == 5 && s[0].first == 0 && s[0].second == 0.0f);
    for (size_t i = 1; i < s.size(); ++i) CHECK(s[i - 1].second <= s[i].second);  // distance metric: ascending
    STEP("small index ok");
    // DualPrecisionHnsw (native/dual_precision_tests.rs:12-120, 260-334)
    {
        DualPrecisionHnsw dp(DistanceMetric::Euclidean, 32, 16, 100, 1000);
        CHECK(dp.is_empty() && !dp.is_quantizer_trained());
        CHECK(dp.search(std::vector<float>(32, 0.0f), 10, 50).empty());
        for (uint64_t i = 0; i < 100; ++i) {
            std::vector<float> v(32);
            for (size_t j = 0; j < 32; ++j) v[j] = (float)(i * 32 + j);
            CHECK(dp.insert(v) == i);
        }
        CHECK(dp.len() == 100 && !dp.is_quantizer_trained());
        std::vector<float> q0(32);
        for (size_t j = 0; j < 32; ++j) q0[j] = (float)j;
        auto r0 = dp.search(q0, 10, 50);
        CHECK(!r0.empty() && r0[0].first == 0);
        dp.force_train_quantizer();
        CHECK(dp.is_quantizer_trained());
        DualPrecisionConfig cfg;
        CHECK(cfg.oversampling_ratio == 4 && cfg.use_int8_traversal && cfg.min_index_size == 10000);
        auto gated = dp.search_with_config(q0, 10, 50, cfg);  // below min_index_size: the f32 path answers
        CHECK(gated.size() == r0.size());
        for (size_t i = 0; i < r0.size(); ++i) CHECK(gated[i] == r0[i]);
        cfg.min_index_size = 0;
        auto i8 = dp.search_with_config(q0, 10, 50, cfg);
        CHECK(!i8.empty() && i8[0].first == 0 && i8[0].second == 0.0f);
        for (size_t i = 1; i < i8.size(); ++i) CHECK(i8[i - 1].second <= i8[i].second);
        // trains itself at the threshold
        DualPrecisionHnsw dp2(DistanceMetric::Euclidean, 32, 16, 100, 100);
        for (uint64_t i = 0; i < 100; ++i) {
            std::vector<float> v(32);
            for (size_t j = 0; j < 32; ++j) v[j] = std::sin((float)(i * 32 + j) * 0.01f);
            dp2.insert(v);
        }
        CHECK(dp2.is_quantizer_trained());
    }
    STEP("dual precision ok");
    std::printf("host mirror ok\n");
    return 0;
}

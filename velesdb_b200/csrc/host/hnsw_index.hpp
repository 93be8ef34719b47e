// hnsw_index.hpp -- C++ host-side mirror of velesdb-core's `HnswIndex` over the C ABI.
//
// The reference's wrapper is compiled (Rust) code; its toolchain is absent from this image, so this
// header is the compiled-language counterpart a non-Rust host links against: same method names,
// argument meaning and error behaviour as
//   crates/velesdb-core/src/index/mod.rs:30-83              (trait VectorIndex)
//   crates/velesdb-core/src/index/hnsw/index/search.rs      (search_with_quality, brute force, rerank)
//   crates/velesdb-core/src/index/hnsw/index/batch.rs       (insert_batch_parallel, search_batch_parallel)
//   crates/velesdb-core/src/index/hnsw/index/trait_impl.rs  (insert / remove / len)
//   crates/velesdb-core/src/index/hnsw/index/constructors.rs(save / load)
//   crates/velesdb-core/src/index/hnsw/index/vacuum.rs      (tombstone accounting)
// A Rust `assert_eq!` panic is a std::invalid_argument here; io::Error is std::runtime_error.
// Everything numeric runs in libveles_b200.so; this class keeps what the Rust wrapper keeps on the
// host: external id <-> node index maps, tombstones, SearchQuality -> ef, transform_score.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iterator>
#include <mutex>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../../include/veles_b200.h"

namespace veles {
namespace host {

enum class DistanceMetric : int32_t { Cosine = 0, Euclidean = 1, DotProduct = 2, Hamming = 3, Jaccard = 4 };

inline bool higher_is_better(DistanceMetric m) {  // core/distance.rs:76-81
    return m == DistanceMetric::Cosine || m == DistanceMetric::DotProduct || m == DistanceMetric::Jaccard;
}

struct SearchQuality {  // index/hnsw/params.rs:283-320
    int32_t kind;
    uint64_t ef;
    static SearchQuality Fast() { return {VELES_FAST, 0}; }
    static SearchQuality Balanced() { return {VELES_BALANCED, 0}; }
    static SearchQuality Accurate() { return {VELES_ACCURATE, 0}; }
    static SearchQuality Perfect() { return {VELES_PERFECT, 0}; }
    static SearchQuality Custom(uint64_t ef) { return {VELES_CUSTOM, ef}; }
    uint64_t ef_search(uint64_t k) const { return veles_ef_search(kind, k, ef); }
};

struct HnswParams {  // index/hnsw/params.rs:14-57
    uint32_t max_connections = 32, ef_construction = 400;
    uint64_t max_elements = 100000;
    static HnswParams auto_for(size_t dimension) {
        HnswParams p;
        if (dimension <= 256) {
            p.max_connections = 24;
            p.ef_construction = 300;
        }
        return p;
    }
};

using Hit = std::pair<uint64_t, float>;

class HnswIndex {
  public:
    HnswIndex(size_t dimension, DistanceMetric metric) : HnswIndex(dimension, metric, HnswParams::auto_for(dimension), true) {}
    HnswIndex(size_t dimension, DistanceMetric metric, HnswParams params, bool enable_vector_storage = true)
        : dimension_(dimension), metric_(metric), params_(params), enable_vector_storage_(enable_vector_storage) {}
    ~HnswIndex() {
        if (snap_) veles_index_free(snap_);
    }
    HnswIndex(const HnswIndex&) = delete;
    HnswIndex& operator=(const HnswIndex&) = delete;

    // ---- VectorIndex ----
    void insert(uint64_t id, const std::vector<float>& v) { insert(id, v.data(), v.size()); }
    void insert(uint64_t id, const float* v, size_t len) {
        if (len != dimension_)
            throw std::invalid_argument("Vector dimension mismatch: expected " + std::to_string(dimension_) + ", got " +
                                        std::to_string(len));
        std::unique_lock<std::shared_mutex> g(mu_);
        if (id_to_idx_.count(id)) return;  // duplicate ids are skipped (trait_impl.rs:12-25)
        if (!vectors_present_) throw std::logic_error("insert into an index restored from files is not supported yet");
        const uint64_t idx = next_idx_++;
        id_to_idx_[id] = idx;
        idx_to_id_[idx] = id;
        staged_.insert(staged_.end(), v, v + len);
        dirty_ = true;
    }
    template <typename It>
    size_t insert_batch_parallel(It first, It last) {  // batch.rs:82-108; < 100 vectors go sequentially there too
        if (std::distance(first, last) >= 100) bulk_ = true;
        size_t count = 0;
        for (; first != last; ++first) {
            const size_t before = len();
            insert(first->first, first->second);
            count += len() - before;
        }
        return count;
    }
    bool remove(uint64_t id) {  // soft delete (trait_impl.rs:54-58)
        std::unique_lock<std::shared_mutex> g(mu_);
        auto it = id_to_idx_.find(id);
        if (it == id_to_idx_.end()) return false;
        idx_to_id_.erase(it->second);
        id_to_idx_.erase(it);
        return true;
    }
    size_t len() const {
        std::shared_lock<std::shared_mutex> g(mu_);
        return id_to_idx_.size();
    }
    bool is_empty() const { return len() == 0; }
    size_t dimension() const { return dimension_; }
    DistanceMetric metric() const { return metric_; }
    std::vector<Hit> search(const std::vector<float>& q, size_t k) { return search_with_quality(q, k, SearchQuality::Balanced()); }

    // ---- inherent API ----
    std::vector<Hit> search_with_quality(const std::vector<float>& q, size_t k, SearchQuality quality) {
        validate_dimension(q.size(), "Query");
        if (quality.kind == VELES_PERFECT) return search_brute_force(q, k);
        if (len() <= 100 && enable_vector_storage_ && vectors_present_ && next_idx_ > 0) return search_brute_force(q, k);
        return graph_search(q.data(), 1, k, quality.ef_search(k))[0];
    }
    std::vector<std::vector<Hit>> search_batch_parallel(const std::vector<std::vector<float>>& queries, size_t k,
                                                        SearchQuality quality) {  // batch.rs:159-197
        std::vector<float> flat;
        flat.reserve(queries.size() * dimension_);
        for (size_t i = 0; i < queries.size(); ++i) {
            if (queries[i].size() != dimension_)
                throw std::invalid_argument("Query " + std::to_string(i) + " dimension mismatch: expected " +
                                            std::to_string(dimension_) + ", got " + std::to_string(queries[i].size()));
            flat.insert(flat.end(), queries[i].begin(), queries[i].end());
        }
        if (queries.empty()) return {};
        return graph_search(flat.data(), queries.size(), k, quality.ef_search(k));
    }
    std::vector<Hit> search_brute_force(const std::vector<float>& q, size_t k) {  // search.rs:176-219
        validate_dimension(q.size(), "Query");
        if (next_idx_ == 0) return {};
        if (!enable_vector_storage_ || !vectors_present_)
            return graph_search(q.data(), 1, k, SearchQuality::Accurate().ef_search(k))[0];
        veles_index_t* s = ensure_snapshot();
        const size_t kk = std::min<size_t>(veles_index_len(s), k + tombstone_count());
        if (kk == 0) return {};
        std::vector<uint32_t> ids(kk);
        std::vector<float> sc(kk);
        check(veles_bruteforce_batch(s, q.data(), 1, (uint32_t)kk, ids.data(), sc.data(), nullptr));
        std::vector<Hit> out;
        std::shared_lock<std::shared_mutex> g(mu_);
        for (size_t j = 0; j < kk && out.size() < k; ++j) {
            if (ids[j] == VELES_INVALID_ID) break;
            auto it = idx_to_id_.find(ids[j]);
            if (it != idx_to_id_.end()) out.emplace_back(it->second, sc[j]);  // tombstones filtered (search.rs:205-208)
        }
        return out;
    }
    std::vector<Hit> brute_force_search_parallel(const std::vector<float>& q, size_t k) { return search_brute_force(q, k); }
    std::vector<Hit> search_with_rerank(const std::vector<float>& q, size_t k, size_t rerank_k) {  // search.rs:118-160
        validate_dimension(q.size(), "Query");
        std::vector<Hit> cands = search_with_quality(q, rerank_k, SearchQuality::Accurate());
        if (cands.empty() || !vectors_present_) return {};
        std::vector<uint32_t> nodes;
        std::vector<uint64_t> ext;
        {
            std::shared_lock<std::shared_mutex> g(mu_);
            for (auto& c : cands) {
                auto it = id_to_idx_.find(c.first);
                if (it != id_to_idx_.end()) {
                    nodes.push_back((uint32_t)it->second);
                    ext.push_back(c.first);
                }
            }
        }
        if (nodes.empty()) return {};
        std::vector<float> sc(nodes.size());
        check(veles_rerank_batch(ensure_snapshot(), q.data(), 1, nodes.data(), (uint32_t)nodes.size(), sc.data(), nullptr));
        std::vector<Hit> out;
        for (size_t i = 0; i < nodes.size(); ++i) out.emplace_back(ext[i], sc[i]);
        sort_results(out);
        if (out.size() > k) out.resize(k);
        return out;
    }
    void set_searching_mode() { ensure_snapshot(); }

    // ---- persistence (constructors.rs:190-287): native_hnsw.{vectors,graph} in the reference's format v1 +
    // an id-map file of this wrapper (the reference's bincode maps are not reproduced) ----
    void save(const std::string& dir) {
        veles_index_t* s = ensure_snapshot();
        check(veles_index_dump(s, dir.c_str(), "native_hnsw"));
        std::ofstream f(dir + "/native_mappings.bin", std::ios::binary);
        if (!f) throw std::runtime_error("cannot create " + dir + "/native_mappings.bin");
        std::shared_lock<std::shared_mutex> g(mu_);
        const uint64_t n = idx_to_id_.size(), nx = next_idx_;
        f.write((const char*)&n, 8);
        f.write((const char*)&nx, 8);
        for (auto& kv : idx_to_id_) {
            f.write((const char*)&kv.first, 8);
            f.write((const char*)&kv.second, 8);
        }
    }
    static HnswIndex* load(const std::string& dir, size_t dimension, DistanceMetric metric) {
        std::ifstream f(dir + "/native_mappings.bin", std::ios::binary);
        if (!f) throw std::runtime_error("native_mappings.bin not found in " + dir);
        HnswIndex* ix = new HnswIndex(dimension, metric);
        veles_index_t* s = nullptr;
        if (veles_index_from_reference_files(dir.c_str(), "native_hnsw", (int32_t)metric, VELES_F32, &s) != VELES_OK) {
            delete ix;
            throw std::runtime_error(veles_last_error());
        }
        uint64_t n = 0, nx = 0;
        f.read((char*)&n, 8);
        f.read((char*)&nx, 8);
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t idx, id;
            f.read((char*)&idx, 8);
            f.read((char*)&id, 8);
            ix->idx_to_id_[idx] = id;
            ix->id_to_idx_[id] = idx;
        }
        ix->next_idx_ = nx;
        ix->snap_ = s;
        ix->vectors_present_ = false;  // ShardedVectors stays empty after load (constructors.rs:240)
        return ix;
    }

    // ---- vacuum.rs ----
    size_t tombstone_count() const {
        std::shared_lock<std::shared_mutex> g(mu_);
        return (size_t)next_idx_ - id_to_idx_.size();
    }
    double tombstone_ratio() const { return next_idx_ == 0 ? 0.0 : (double)tombstone_count() / (double)next_idx_; }
    bool needs_vacuum() const { return tombstone_ratio() > 0.2; }

  private:
    void validate_dimension(size_t got, const char* what) const {  // search.rs:16-24
        if (got != dimension_)
            throw std::invalid_argument(std::string(what) + " dimension mismatch: expected " + std::to_string(dimension_) +
                                        ", got " + std::to_string(got));
    }
    static void check(int32_t status) {
        if (status != VELES_OK) throw std::runtime_error(std::string("veles status ") + std::to_string(status) + ": " + veles_last_error());
    }
    void sort_results(std::vector<Hit>& r) const {  // core/distance.rs:95-103 (stable, total order)
        auto key = [](float f) {
            int32_t b;
            std::memcpy(&b, &f, 4);
            return b ^ (int32_t)(((uint32_t)(b >> 31)) >> 1);
        };
        if (higher_is_better(metric_))
            std::stable_sort(r.begin(), r.end(), [&](const Hit& a, const Hit& b) { return key(b.second) < key(a.second); });
        else
            std::stable_sort(r.begin(), r.end(), [&](const Hit& a, const Hit& b) { return key(a.second) < key(b.second); });
    }
    veles_index_t* ensure_snapshot() {
        std::unique_lock<std::shared_mutex> g(mu_);
        if (!snap_ || dirty_) {
            if (snap_) veles_index_free(snap_);
            snap_ = nullptr;
            check(veles_index_from_vectors(staged_.data(), next_idx_, (uint32_t)dimension_, VELES_F32, VELES_F32, (int32_t)metric_, &snap_));
            if (!bulk_ && next_idx_ <= 20000)  // sequential inserts: the reference's deterministic graph
                check(veles_index_build_graph_exact(snap_, params_.max_connections, params_.ef_construction, nullptr));
            else
                check(veles_index_build_graph(snap_, params_.max_connections, 0, nullptr));
            dirty_ = false;
        }
        return snap_;
    }
    std::vector<std::vector<Hit>> graph_search(const float* q, size_t nq, size_t k, uint64_t ef) {
        std::vector<std::vector<Hit>> out(nq);
        if (next_idx_ == 0) return out;
        veles_index_t* s = ensure_snapshot();
        std::vector<uint32_t> ids(nq * k), cnt(nq);
        std::vector<float> dist(nq * k);
        check(veles_search_batch(s, q, (uint32_t)nq, (uint32_t)k, (uint32_t)ef, ids.data(), dist.data(), cnt.data(), nullptr, nullptr));
        std::shared_lock<std::shared_mutex> g(mu_);
        for (size_t r = 0; r < nq; ++r)
            for (uint32_t j = 0; j < cnt[r]; ++j) {
                auto it = idx_to_id_.find(ids[r * k + j]);
                if (it == idx_to_id_.end()) continue;  // tombstoned (search.rs:86-91)
                out[r].emplace_back(it->second, veles_transform_score((int32_t)metric_, dist[r * k + j]));
            }
        return out;
    }

    size_t dimension_;
    DistanceMetric metric_;
    HnswParams params_;
    bool enable_vector_storage_;
    mutable std::shared_mutex mu_;
    std::unordered_map<uint64_t, uint64_t> id_to_idx_, idx_to_id_;
    uint64_t next_idx_ = 0;
    std::vector<float> staged_;
    veles_index_t* snap_ = nullptr;
    bool dirty_ = false;
    bool bulk_ = false;
    bool vectors_present_ = true;
};

}  // namespace host
}  // namespace veles

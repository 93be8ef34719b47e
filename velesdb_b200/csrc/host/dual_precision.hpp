// dual_precision.hpp -- C++ host-side mirror of velesdb-core's `DualPrecisionHnsw`
// (crates/velesdb-core/src/index/hnsw/native/dual_precision.rs:60-441) over the C ABI.
//
// Same method names, argument meaning and defaults as the Rust type:
//   new(distance, dimension, max_connections, ef_construction, max_elements)   :87-103
//   insert(vector) -> node id; the quantizer trains itself on the first
//     min(1000, max_elements) vectors                                           :100, 122-143
//   force_train_quantizer(), is_quantizer_trained(), len(), is_empty()          :106-119, 172-176
//   search(query, k, ef_search)                 f32 traversal (= NativeHnsw::search, see below)  :179-228
//   search_with_config(query, k, ef_search, config)   int8 traversal + exact re-rank             :263-325
//   DualPrecisionConfig defaults 4 / true / 10_000                                                :33-57
// The reference mutates its graph per insert; here vectors are staged and the device snapshot (graph by
// the exact sequential builder, SQ8 store by veles_index_attach_sq8) is rebuilt on the next search.
// Everything numeric runs in libveles_b200.so.
#pragma once

#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "hnsw_index.hpp"

namespace veles {
namespace host {

struct DualPrecisionConfig {  // dual_precision.rs:33-57
    size_t oversampling_ratio = 4;
    bool use_int8_traversal = true;
    size_t min_index_size = 10000;
};

using NodeHit = std::pair<uint64_t, float>;  // (node id, raw graph distance), as NativeHnsw returns

class DualPrecisionHnsw {
  public:
    DualPrecisionHnsw(DistanceMetric metric, size_t dimension, size_t max_connections, size_t ef_construction,
                      size_t max_elements)
        : metric_(metric), dimension_(dimension), max_connections_(max_connections), ef_construction_(ef_construction),
          training_sample_size_(std::min<size_t>(1000, max_elements)) {}
    ~DualPrecisionHnsw() {
        if (snap_) veles_index_free(snap_);
    }
    DualPrecisionHnsw(const DualPrecisionHnsw&) = delete;
    DualPrecisionHnsw& operator=(const DualPrecisionHnsw&) = delete;

    size_t len() const { return count_; }
    bool is_empty() const { return count_ == 0; }
    bool is_quantizer_trained() const { return train_count_ > 0; }

    uint64_t insert(const std::vector<float>& v) {
        if (v.size() != dimension_)
            throw std::invalid_argument("Vector dimension mismatch: expected " + std::to_string(dimension_) + ", got " +
                                        std::to_string(v.size()));
        staged_.insert(staged_.end(), v.begin(), v.end());
        dirty_ = true;
        const uint64_t node = count_++;
        if (train_count_ == 0 && count_ >= training_sample_size_) train_count_ = count_;
        return node;
    }
    void force_train_quantizer() {
        if (train_count_ == 0 && count_ > 0) {
            train_count_ = count_;
            dirty_ = true;
        }
    }
    // With a trained quantizer the reference still traverses in f32, asks for max(2 ef, 4 k) candidates (it gets at
    // most ef), recomputes the same exact distances and stably re-sorts a sorted list: NativeHnsw::search.
    std::vector<NodeHit> search(const std::vector<float>& q, size_t k, size_t ef_search) {
        check_query(q);
        std::vector<NodeHit> out;
        if (count_ == 0) return out;
        std::vector<uint32_t> ids(k);
        std::vector<float> dist(k);
        uint32_t cnt = 0;
        check(veles_search_batch(ensure(), q.data(), 1, (uint32_t)k, (uint32_t)ef_search, ids.data(), dist.data(), &cnt, nullptr,
                                 nullptr));
        for (uint32_t j = 0; j < cnt; ++j) out.emplace_back(ids[j], dist[j]);
        return out;
    }
    std::vector<NodeHit> search_with_config(const std::vector<float>& q, size_t k, size_t ef_search,
                                            const DualPrecisionConfig& config = DualPrecisionConfig()) {
        if (train_count_ == 0 || !config.use_int8_traversal || count_ < config.min_index_size)
            return search(q, k, ef_search);  // dual_precision.rs:271-279
        check_query(q);
        std::vector<uint32_t> ids(k);
        std::vector<float> dist(k);
        uint32_t cnt = 0;
        check(veles_search_batch_sq8(ensure(), q.data(), 1, (uint32_t)k, (uint32_t)ef_search, (uint32_t)config.oversampling_ratio,
                                     ids.data(), dist.data(), &cnt, nullptr, nullptr));
        std::vector<NodeHit> out;
        for (uint32_t j = 0; j < cnt; ++j) out.emplace_back(ids[j], dist[j]);
        return out;
    }

  private:
    static void check(int32_t rc) {
        if (rc != VELES_OK) throw std::runtime_error(veles_last_error());
    }
    void check_query(const std::vector<float>& q) const {
        if (q.size() != dimension_)
            throw std::invalid_argument("Query dimension mismatch: expected " + std::to_string(dimension_) + ", got " +
                                        std::to_string(q.size()));
    }
    veles_index_t* ensure() {
        if (!snap_ || dirty_) {
            if (snap_) veles_index_free(snap_);
            snap_ = nullptr;
            check(veles_index_from_vectors(staged_.data(), count_, (uint32_t)dimension_, VELES_F32, VELES_F32, (int32_t)metric_, &snap_));
            check(veles_index_build_graph_exact(snap_, (uint32_t)max_connections_, (uint32_t)ef_construction_, nullptr));
            if (train_count_) check(veles_index_attach_sq8(snap_, train_count_, nullptr));
            dirty_ = false;
        }
        return snap_;
    }

    DistanceMetric metric_;
    size_t dimension_, max_connections_, ef_construction_, training_sample_size_;
    uint64_t count_ = 0, train_count_ = 0;
    std::vector<float> staged_;
    veles_index_t* snap_ = nullptr;
    bool dirty_ = false;
};

}  // namespace host
}  // namespace veles

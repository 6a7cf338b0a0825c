// raw int64 id <-> dense index, same surface as the reference's qmf/utils/IdIndex.h:27-62
#pragma once
#include <cstddef>
#include <cstdint>
#include <limits>
#include <unordered_map>
#include <vector>

namespace qmf {

class IdIndex {
 public:
  static constexpr size_t missingIdx = std::numeric_limits<size_t>::max();

  int64_t id(size_t idx) const { return ids_[idx]; }

  size_t idx(int64_t id) const {
    const auto it = lookup_.find(id);
    return it == lookup_.end() ? missingIdx : it->second;
  }

  size_t getOrSetIdx(int64_t id) {
    const auto inserted = lookup_.emplace(id, ids_.size());
    if (inserted.second) ids_.push_back(id);
    return inserted.first->second;
  }

  size_t size() const { return ids_.size(); }
  const std::vector<int64_t>& ids() const { return ids_; }

  void reset() {
    ids_.clear();
    lookup_.clear();
  }

 private:
  std::vector<int64_t> ids_;
  std::unordered_map<int64_t, size_t> lookup_;
};

}  // namespace qmf

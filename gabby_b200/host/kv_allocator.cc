#include "kv_allocator.h"

#include <string>

namespace gabby {
namespace inference {

KvPageAllocator::KvPageAllocator(int num_pages, int page_size, int max_blocks_per_seq)
    : num_pages_(num_pages), page_size_(page_size), max_blocks_(max_blocks_per_seq) {
    if (num_pages <= 0 || page_size <= 0 || max_blocks_per_seq <= 0) throw std::invalid_argument("KvPageAllocator: bad sizes");
    free_.reserve(num_pages);
    for (int p = num_pages - 1; p >= 0; p--) free_.push_back(p);  // page 0 is handed out first
}

int KvPageAllocator::NewSequence() {
    const int id = next_id_++;
    tables_[id];
    return id;
}

void KvPageAllocator::Reserve(int seq, int total_tokens) {
    auto it = tables_.find(seq);
    if (it == tables_.end()) throw std::invalid_argument("KvPageAllocator: unknown sequence " + std::to_string(seq));
    const int need = (total_tokens + page_size_ - 1) / page_size_;
    const int have = static_cast<int>(it->second.size());
    if (need <= have) return;
    if (need > max_blocks_)
        throw KvOutOfPages("sequence needs " + std::to_string(need) + " KV pages, the per-sequence limit is " + std::to_string(max_blocks_));
    if (need - have > static_cast<int>(free_.size()))
        throw KvOutOfPages("KV pool exhausted: need " + std::to_string(need - have) + " more pages, " + std::to_string(free_.size()) + " free");
    for (int i = have; i < need; i++) {
        it->second.push_back(free_.back());
        free_.pop_back();
    }
}

void KvPageAllocator::Free(int seq) {
    auto it = tables_.find(seq);
    if (it == tables_.end()) return;
    for (auto p = it->second.rbegin(); p != it->second.rend(); ++p) free_.push_back(*p);
    tables_.erase(it);
}

const std::vector<int32_t>& KvPageAllocator::BlockTable(int seq) const {
    auto it = tables_.find(seq);
    if (it == tables_.end()) throw std::invalid_argument("KvPageAllocator: unknown sequence " + std::to_string(seq));
    return it->second;
}

std::vector<int32_t> KvPageAllocator::BatchTable(const std::vector<int>& seqs) const {
    std::vector<int32_t> out(seqs.size() * static_cast<size_t>(max_blocks_), 0);
    for (size_t i = 0; i < seqs.size(); i++) {
        const auto& t = BlockTable(seqs[i]);
        for (size_t j = 0; j < t.size(); j++) out[i * max_blocks_ + j] = t[j];
    }
    return out;
}

}  // namespace inference
}  // namespace gabby

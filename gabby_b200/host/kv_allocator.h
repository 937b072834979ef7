// kv_allocator.h -- host-side allocator of the paged KV cache (north_star: "KV-cache allocator"
// is host C++). The device pool (b2l_params.num_pages pages of page_size tokens, per layer) is
// dumb storage; which page belongs to which sequence is decided here and handed to the C-ABI
// as block tables. No reference counterpart (gabby has no KV cache).
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <vector>

namespace gabby {
namespace inference {

class KvOutOfPages : public std::runtime_error {
public:
    using std::runtime_error::runtime_error;
};

class KvPageAllocator {
public:
    KvPageAllocator(int num_pages, int page_size, int max_blocks_per_seq);

    int NewSequence();                              // -> sequence id
    // make room for `total_tokens` tokens in the sequence (grows its block table); throws
    // KvOutOfPages when the pool or the per-sequence limit is exhausted (nothing is changed then)
    void Reserve(int seq, int total_tokens);
    void Free(int seq);                             // pages go back to the free list (LIFO)
    const std::vector<int32_t>& BlockTable(int seq) const;
    // [n][max_blocks] row-major table for a batch of sequences, padded with 0
    std::vector<int32_t> BatchTable(const std::vector<int>& seqs) const;

    int page_size() const { return page_size_; }
    int max_blocks() const { return max_blocks_; }
    int free_pages() const { return static_cast<int>(free_.size()); }
    int num_pages() const { return num_pages_; }
    int live_sequences() const { return static_cast<int>(tables_.size()); }

private:
    int num_pages_, page_size_, max_blocks_, next_id_ = 0;
    std::vector<int32_t> free_;
    std::map<int, std::vector<int32_t>> tables_;
};

}  // namespace inference
}  // namespace gabby

// sampler.h -- greedy sampling policy (north_star: "greedy sampler" is host C++). The argmax
// itself runs on the device (first-max rule); the host decides when a sequence stops.
#pragma once
#include <cstdint>
#include <vector>

namespace gabby {
namespace inference {

enum class FinishReason { kNone, kStop, kLength };

class GreedySampler {
public:
    GreedySampler(std::vector<int> eos_token_ids, int max_new_tokens)
        : eos_(std::move(eos_token_ids)), max_new_(max_new_tokens) {}

    // feed the id the device produced; returns the reason the sequence finished, if it did
    FinishReason Accept(int32_t id);
    bool IsEos(int32_t id) const;
    const std::vector<int32_t>& tokens() const { return out_; }   // generated ids, EOS excluded
    int max_new_tokens() const { return max_new_; }

    // first-max argmax on host logits (ties -> lowest index): the rule the device kernel follows
    static int32_t Argmax(const float* logits, int64_t n);

private:
    std::vector<int> eos_;
    int max_new_;
    std::vector<int32_t> out_;
};

}  // namespace inference
}  // namespace gabby

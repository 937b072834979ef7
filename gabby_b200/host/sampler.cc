#include "sampler.h"

#include <algorithm>

namespace gabby {
namespace inference {

bool GreedySampler::IsEos(int32_t id) const { return std::find(eos_.begin(), eos_.end(), id) != eos_.end(); }

FinishReason GreedySampler::Accept(int32_t id) {
    if (IsEos(id)) return FinishReason::kStop;
    out_.push_back(id);
    if (static_cast<int>(out_.size()) >= max_new_) return FinishReason::kLength;
    return FinishReason::kNone;
}

int32_t GreedySampler::Argmax(const float* logits, int64_t n) {
    int64_t best = 0;
    for (int64_t i = 1; i < n; i++)
        if (logits[i] > logits[best]) best = i;
    return static_cast<int32_t>(best);
}

}  // namespace inference
}  // namespace gabby

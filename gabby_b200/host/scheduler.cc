#include "scheduler.h"

#include <algorithm>
#include <stdexcept>
#include <string>

#include "../../include/b2l.h"

namespace gabby {
namespace inference {

void B2lBatchEngine::Prefill(int n_seq, const int32_t* tokens, const int32_t* q_lens, const int32_t* start_pos,
                             const int32_t* block_tables, int max_blocks, int32_t* next_ids) {
    if (b2l_prefill(ctx_, n_seq, tokens, q_lens, start_pos, block_tables, max_blocks, next_ids) != 0)
        throw std::runtime_error(std::string("b2l_prefill: ") + b2l_last_error(ctx_));
}

void B2lBatchEngine::Decode(int n_seq, const int32_t* tokens, const int32_t* positions, const int32_t* block_tables,
                            int max_blocks, int32_t* next_ids) {
    if (b2l_decode(ctx_, n_seq, tokens, positions, block_tables, max_blocks, next_ids) != 0)
        throw std::runtime_error(std::string("b2l_decode: ") + b2l_last_error(ctx_));
}

BatchScheduler::BatchScheduler(BatchEngine* engine, KvPageAllocator* kv, SchedulerLimits limits, std::vector<int> eos_token_ids)
    : engine_(engine), kv_(kv), lim_(limits), eos_(std::move(eos_token_ids)) {
    if (!engine_ || !kv_) throw std::invalid_argument("BatchScheduler: null engine or allocator");
    if (lim_.max_batch < 1 || lim_.max_positions < 2 || lim_.max_prefill_tokens < 1)
        throw std::invalid_argument("BatchScheduler: bad limits");
}

bool BatchScheduler::IsEos(int32_t id) const { return std::find(eos_.begin(), eos_.end(), static_cast<int>(id)) != eos_.end(); }

int BatchScheduler::Submit(std::vector<int32_t> prompt, int max_new_tokens) {
    if (prompt.empty()) throw std::invalid_argument("Submit: empty prompt");
    if (max_new_tokens < 1) throw std::invalid_argument("Submit: max_new_tokens must be >= 1");
    const int n = static_cast<int>(prompt.size());
    // a preempted sequence is recomputed as prompt + generated tokens, so the worst case must fit, not just the prompt
    const int worst = n + max_new_tokens;
    if (worst > lim_.max_positions) throw std::invalid_argument("Submit: prompt + max_new_tokens exceeds max_positions");
    if (worst > lim_.max_prefill_tokens) throw std::invalid_argument("Submit: prompt + max_new_tokens exceeds the prefill budget");
    const int pages = (worst + kv_->page_size() - 1) / kv_->page_size();
    if (pages > kv_->num_pages() || pages > kv_->max_blocks()) throw std::invalid_argument("Submit: request can never fit the KV pool");
    Seq s;
    s.id = next_id_++;
    s.prompt = std::move(prompt);
    s.max_new = max_new_tokens;
    SchedResult r;
    r.prompt_tokens = n;
    results_[s.id] = std::move(r);
    waiting_.push_back(std::move(s));
    return waiting_.back().id;
}

const SchedResult& BatchScheduler::Result(int id) const {
    auto it = results_.find(id);
    if (it == results_.end()) throw std::out_of_range("Result: unknown request id");
    return it->second;
}

void BatchScheduler::Forget(int id) {
    auto it = results_.find(id);
    if (it != results_.end() && it->second.done) results_.erase(it);
}

void BatchScheduler::Fail(Seq& s, const std::string& why) {
    SchedResult& r = results_[s.id];
    r.error = why;
    r.done = true;
    if (s.kv_seq >= 0) kv_->Free(s.kv_seq);
    s.kv_seq = -1;
}

void BatchScheduler::Retire(Seq& s, FinishReason why) {
    SchedResult& r = results_[s.id];
    r.finish = why;
    r.done = true;
    if (s.kv_seq >= 0) kv_->Free(s.kv_seq);
    s.kv_seq = -1;
}

bool BatchScheduler::Accept(Seq& s, int32_t id) {
    SchedResult& r = results_[s.id];
    if (IsEos(id)) {
        Retire(s, FinishReason::kStop);
        return true;
    }
    r.tokens.push_back(id);
    if (static_cast<int>(r.tokens.size()) >= s.max_new) {
        Retire(s, FinishReason::kLength);
        return true;
    }
    return false;
}

int BatchScheduler::Admit() {
    std::vector<Seq> batch;
    int budget = lim_.max_prefill_tokens;
    while (!waiting_.empty() && static_cast<int>(running_.size() + batch.size()) < lim_.max_batch) {
        Seq& w = waiting_.front();
        const int n = static_cast<int>(w.prompt.size());
        if (n > budget) break;                       // FIFO: do not let short requests overtake a long one forever
        const int kv_seq = kv_->NewSequence();
        try {
            // the prompt, the first generated token and -- unless that one already ends the request -- the token this
            // step's decode produces
            const int remaining = w.max_new - static_cast<int>(results_[w.id].tokens.size());
            kv_->Reserve(kv_seq, n + std::min(2, std::max(1, remaining)));
        } catch (const KvOutOfPages&) {
            kv_->Free(kv_seq);
            break;                                    // wait for running sequences to finish (or be preempted)
        }
        w.kv_seq = kv_seq;
        w.admitted_at = admit_counter_++;
        budget -= n;
        batch.push_back(std::move(w));
        waiting_.pop_front();
    }
    if (batch.empty()) return 0;
    const int n_seq = static_cast<int>(batch.size()), mb = kv_->max_blocks();
    std::vector<int32_t> tokens, q_lens(n_seq), start(n_seq, 0), next(n_seq);
    std::vector<int> kv_ids(n_seq);
    for (int i = 0; i < n_seq; i++) {
        tokens.insert(tokens.end(), batch[i].prompt.begin(), batch[i].prompt.end());
        q_lens[i] = static_cast<int32_t>(batch[i].prompt.size());
        kv_ids[i] = batch[i].kv_seq;
    }
    const std::vector<int32_t> bt = kv_->BatchTable(kv_ids);
    try {
        engine_->Prefill(n_seq, tokens.data(), q_lens.data(), start.data(), bt.data(), mb, next.data());
    } catch (const std::exception& e) {
        // the requests are already off the queue and hold pages: fail them explicitly instead of leaking both
        for (Seq& s : batch) Fail(s, e.what());
        throw;
    }
    stats_.prefill_calls++;
    stats_.prefill_tokens += static_cast<int64_t>(tokens.size());
    for (int i = 0; i < n_seq; i++) {
        Seq& s = batch[i];
        s.pos = static_cast<int>(s.prompt.size());
        s.last = next[i];
        if (!Accept(s, next[i])) running_.push_back(std::move(s));
    }
    return n_seq;
}

void BatchScheduler::PreemptYoungest() {
    auto it = std::max_element(running_.begin(), running_.end(), [](const Seq& a, const Seq& b) { return a.admitted_at < b.admitted_at; });
    Seq s = std::move(*it);
    running_.erase(it);
    kv_->Free(s.kv_seq);
    s.kv_seq = -1;
    // recompute later: everything produced so far becomes part of the prompt; the result keeps the generated tokens
    const SchedResult& r = results_[s.id];
    std::vector<int32_t> p(s.prompt.begin(), s.prompt.begin() + r.prompt_tokens);
    p.insert(p.end(), r.tokens.begin(), r.tokens.end());
    s.prompt = std::move(p);
    waiting_.push_front(std::move(s));
    stats_.preemptions++;
}

int BatchScheduler::Step() {
    stats_.steps++;
    // 1. every RUNNING sequence needs room for the token it is about to produce; when the pool is short the youngest is
    //    preempted -- before anything new is admitted, so that a fresh prompt is not prefilled just to be thrown away
    bool preempted = false;
    for (;;) {
        bool ok = true;
        for (Seq& s : running_) {
            try {
                kv_->Reserve(s.kv_seq, s.pos + 2);
            } catch (const KvOutOfPages&) {
                ok = false;
                break;
            }
        }
        if (ok) break;
        if (running_.size() <= 1) throw std::runtime_error("BatchScheduler: KV pool too small for a single sequence");
        PreemptYoungest();
        preempted = true;
    }
    // 2. admit waiting requests into the free slots (not in a step that had to preempt: the pool is under pressure)
    const int progressed = preempted ? 0 : Admit();
    stats_.max_concurrent = std::max(stats_.max_concurrent, static_cast<int>(running_.size()));
    if (running_.empty()) return progressed;
    const int n_seq = static_cast<int>(running_.size()), mb = kv_->max_blocks();
    std::vector<int32_t> tokens(n_seq), positions(n_seq), next(n_seq);
    std::vector<int> kv_ids(n_seq);
    for (int i = 0; i < n_seq; i++) {
        tokens[i] = running_[i].last;
        positions[i] = running_[i].pos;
        kv_ids[i] = running_[i].kv_seq;
    }
    const std::vector<int32_t> bt = kv_->BatchTable(kv_ids);
    try {
        engine_->Decode(n_seq, tokens.data(), positions.data(), bt.data(), mb, next.data());
    } catch (const std::exception& e) {
        for (Seq& s : running_) Fail(s, e.what());
        running_.clear();
        throw;
    }
    stats_.decode_calls++;
    stats_.decode_tokens += n_seq;
    std::vector<Seq> still;
    still.reserve(running_.size());
    for (int i = 0; i < n_seq; i++) {
        Seq& s = running_[i];
        s.pos++;
        s.last = next[i];
        if (!Accept(s, next[i])) still.push_back(std::move(s));
    }
    running_ = std::move(still);
    return progressed + n_seq;
}

void BatchScheduler::Drain() {
    while (!waiting_.empty() || !running_.empty()) {
        if (Step() == 0 && running_.empty() && !waiting_.empty())
            throw std::runtime_error("BatchScheduler: a waiting request cannot be admitted (KV pool or prefill budget too small)");
    }
}

}  // namespace inference
}  // namespace gabby

// scheduler.h -- continuous batching over the B200 engine.
// gabby serialises requests under its thread-pool mutex (/root/reference/src/http/thread_pool.cc:22-28) and its
// generator answers one request at a time (/root/reference/src/service.cc:150); batch > 1 decode is only reachable
// when something feeds the engine several sequences per step. This is that something, as host C++:
//   * requests queue up (Submit); every Step() first ADMITS waiting requests while a batch slot, KV pages and the
//     prefill token budget allow -- all admitted prompts run as ONE ragged b2l_prefill call -- and then runs ONE
//     b2l_decode step for every running sequence (tokens, positions and block tables of the whole batch);
//   * sequences finish independently (EOS or max_new_tokens) and free their pages at once, so a waiting request
//     can take the slot on the very next step (iteration-level scheduling, no padding, no batch barrier);
//   * when the KV pool runs out while sequences grow, the youngest running sequence is preempted: its pages are
//     freed and it re-enters the queue with prompt + generated tokens as its new prompt (recompute).
// The engine is reached through BatchEngine so that the policy is testable without a GPU; B2lBatchEngine is the
// real one (b2l_prefill / b2l_decode take block tables per call, so the engine holds no per-sequence state).
#pragma once
#include <cstdint>
#include <deque>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "kv_allocator.h"
#include "sampler.h"

struct b2l_ctx;

namespace gabby {
namespace inference {

class BatchEngine {
public:
    virtual ~BatchEngine() = default;
    // tokens: the new tokens of n_seq sequences back to back (q_lens[i] each), sequence i starting at position
    // start_pos[i]; next_ids[i] = greedy token after sequence i's last new token
    virtual void Prefill(int n_seq, const int32_t* tokens, const int32_t* q_lens, const int32_t* start_pos,
                         const int32_t* block_tables, int max_blocks, int32_t* next_ids) = 0;
    virtual void Decode(int n_seq, const int32_t* tokens, const int32_t* positions, const int32_t* block_tables,
                        int max_blocks, int32_t* next_ids) = 0;
};

// the real engine: throws std::runtime_error with b2l_last_error() on failure
class B2lBatchEngine : public BatchEngine {
public:
    explicit B2lBatchEngine(b2l_ctx* ctx) : ctx_(ctx) {}
    void Prefill(int n_seq, const int32_t* tokens, const int32_t* q_lens, const int32_t* start_pos, const int32_t* block_tables,
                 int max_blocks, int32_t* next_ids) override;
    void Decode(int n_seq, const int32_t* tokens, const int32_t* positions, const int32_t* block_tables, int max_blocks,
                int32_t* next_ids) override;

private:
    b2l_ctx* ctx_;
};

struct SchedulerLimits {
    int max_batch = 8;             // sequences per decode step (the engine's max_batch)
    int max_positions = 2048;      // per-sequence context capacity
    int max_prefill_tokens = 2048; // prompt tokens admitted per step (the engine's prefill scratch)
};

struct SchedulerStats {
    int64_t steps = 0, prefill_calls = 0, decode_calls = 0, prefill_tokens = 0, decode_tokens = 0, preemptions = 0;
    int max_concurrent = 0;
};

struct SchedResult {
    std::vector<int32_t> tokens;   // generated ids, EOS excluded
    FinishReason finish = FinishReason::kNone;
    int prompt_tokens = 0;
    bool done = false;
    std::string error;             // non-empty: the engine failed while this request was in flight (done is true)
};

class BatchScheduler {
public:
    BatchScheduler(BatchEngine* engine, KvPageAllocator* kv, SchedulerLimits limits, std::vector<int> eos_token_ids);

    // queue a request; throws std::invalid_argument for an empty prompt, a prompt that can never fit
    // (max_positions / prefill budget / whole KV pool) or max_new_tokens < 1
    int Submit(std::vector<int32_t> prompt, int max_new_tokens);
    // one scheduling iteration; returns the number of sequences that made progress (0: nothing to do)
    int Step();
    // Step() until every submitted request is done
    void Drain();

    const SchedResult& Result(int id) const;
    void Forget(int id);               // drop a finished request's result (long-running servers)
    bool idle() const { return waiting_.empty() && running_.empty(); }
    int running() const { return static_cast<int>(running_.size()); }
    int waiting() const { return static_cast<int>(waiting_.size()); }
    const SchedulerStats& stats() const { return stats_; }

private:
    struct Seq {
        int id = 0;
        int kv_seq = -1;
        std::vector<int32_t> prompt;   // what has to be (re)computed before decoding continues
        int max_new = 0;
        int32_t last = 0;              // token fed to the next decode step
        int pos = 0;                   // its position
        int admitted_at = 0;           // admission order (preemption picks the youngest)
    };
    bool Accept(Seq& s, int32_t id);   // record a produced token; true when the sequence finished
    void Retire(Seq& s, FinishReason why);
    bool IsEos(int32_t id) const;
    int Admit();                       // prefill as many waiting requests as fit; returns how many
    void Fail(Seq& s, const std::string& why);   // engine error: give the pages back, mark the request done with the message
    void PreemptYoungest();

    BatchEngine* engine_;
    KvPageAllocator* kv_;
    SchedulerLimits lim_;
    std::vector<int> eos_;
    std::deque<Seq> waiting_;
    std::vector<Seq> running_;
    std::map<int, SchedResult> results_;
    SchedulerStats stats_;
    int next_id_ = 0, admit_counter_ = 0;
};

}  // namespace inference
}  // namespace gabby

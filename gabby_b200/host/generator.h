// generator.h -- the drop-in boundary. Same names, members and ownership as gabby's
// /root/reference/src/inference/generator.h:16-47 (Message, Request, Generator, Llama3Generator with
// the static Load(std::unique_ptr<InferenceConfig>) factory); Generate() runs the B200 forward through
// the b2l C-ABI instead of returning a constant string (generator.cc:33-38).
#pragma once
#include <condition_variable>
#include <cstdint>
#include <memory>
#include <mutex>
#include <ostream>
#include <string>
#include <thread>
#include <vector>

#include "config.h"
#include "kv_allocator.h"
#include "sampler.h"
#include "scheduler.h"
#include "tokenizer.h"

struct b2l_ctx;

namespace gabby {
namespace inference {

struct Message {
    std::string role;
    std::string content;
};
std::ostream& operator<<(std::ostream& os, const Message& msg);

struct Request {
    Message system_message;
    Message user_message;
};
std::ostream& operator<<(std::ostream& os, const Request& msg);

class Generator {
public:
    virtual ~Generator() = default;
    virtual Message Generate(const Request& req) = 0;
};

struct GeneratorOptions {
    int device = 0;
    int max_positions = 2048;   // per-sequence context capacity (RoPE table rows, block-table length)
    int page_size = 16;
    int num_pages = 0;          // 0: enough for max_batch sequences of max_positions
    int max_new_tokens = 256;   // default completion budget of Generate (a request's max_tokens overrides it)
    int max_batch = 8;          // sequences decoded together: concurrent Generate calls share decode steps (continuous batching)
};

struct GenerationResult {
    std::vector<int32_t> tokens;   // generated ids (EOS excluded)
    FinishReason finish = FinishReason::kNone;   // kStop: EOS ("stop"), kLength: the token budget ran out ("length")
    int prompt_tokens = 0;         // usage.prompt_tokens; usage.completion_tokens = tokens.size()
    std::string text;              // detokenized completion (GenerateDetailed only)
};

class Llama3Generator : public Generator {
public:
    ~Llama3Generator() override;
    Message Generate(const Request& req) override;

    // same factory as the reference; load-time failures throw (std::runtime_error / json::JSONError)
    static std::unique_ptr<Generator> Load(std::unique_ptr<InferenceConfig> config);
    static std::unique_ptr<Llama3Generator> Load(std::unique_ptr<InferenceConfig> config, const GeneratorOptions& opt);

    // What gabby's service.cc:79-116 hard-codes ("usage" counts, finish_reason "stop") comes from here: the completion text
    // together with prompt / completion token counts and the real finish reason; max_tokens <= 0 = the configured default.
    GenerationResult GenerateDetailed(const Request& req, int max_tokens);

    // token-level entry (bench / parity tests): prefill + greedy decode until EOS or max_new_tokens.
    // device_loop = false: the request joins the continuous-batching queue (concurrent callers share decode steps);
    // device_loop = true: one exclusive device-resident loop (token feedback never leaves the GPU).
    GenerationResult GenerateTokens(const std::vector<int32_t>& prompt, int max_new_tokens, bool device_loop);
    SchedulerStats scheduler_stats();

    const LlamaParams& params() const { return params_; }
    Tokenizer& tokenizer() { return *tokenizer_; }
    b2l_ctx* engine() { return ctx_; }

private:
    Llama3Generator(std::unique_ptr<InferenceConfig> config) : config_(std::move(config)) {}
    void Check(int rc, const char* what) const;   // rethrows the C-ABI error as std::runtime_error

    std::unique_ptr<InferenceConfig> config_;
    LlamaParams params_;
    GeneratorOptions opt_;
    b2l_ctx* ctx_ = nullptr;
    std::unique_ptr<KvPageAllocator> kv_;
    std::unique_ptr<Tokenizer> tokenizer_;
    // Generate is called from HTTP worker threads (reference: http/server.h:35). Requests queue in the scheduler; ONE stepping
    // thread drives the engine (admission + one decode step of every running sequence per iteration), callers wait for
    // their own result. mu_ guards scheduler, allocator and engine.
    void StepLoop();
    int capacity_ = 0;   // tokens one sequence can hold: min(block-table reach, max_positions)
    std::unique_ptr<B2lBatchEngine> batch_engine_;
    std::unique_ptr<BatchScheduler> sched_;
    std::mutex mu_;
    std::condition_variable work_cv_, done_cv_;
    std::thread stepper_;
    bool stop_ = false;
};

}  // namespace inference
}  // namespace gabby

#include "generator.h"

#include <algorithm>
#include <stdexcept>

#include "b2l.h"

namespace gabby {
namespace inference {

std::ostream& operator<<(std::ostream& os, const Message& msg) {
    return os << "{ \"role\": " << msg.role << ", \"content\": " << msg.content << " }";
}
std::ostream& operator<<(std::ostream& os, const Request& msg) {
    return os << "{ \"system\": " << msg.system_message << ", \"user\": " << msg.user_message << " }";
}

void Llama3Generator::Check(int rc, const char* what) const {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + b2l_last_error(ctx_));
}

Llama3Generator::~Llama3Generator() {
    {
        std::lock_guard<std::mutex> lock(mu_);
        stop_ = true;
    }
    work_cv_.notify_all();
    if (stepper_.joinable()) stepper_.join();
    if (ctx_) b2l_destroy(ctx_);
}

/* static */
std::unique_ptr<Generator> Llama3Generator::Load(std::unique_ptr<InferenceConfig> config) {
    return Load(std::move(config), GeneratorOptions{});
}

/* static */
std::unique_ptr<Llama3Generator> Llama3Generator::Load(std::unique_ptr<InferenceConfig> config, const GeneratorOptions& opt) {
    if (!config) throw std::runtime_error("Llama3Generator::Load: null config");
    std::unique_ptr<Llama3Generator> gen(new Llama3Generator(std::move(config)));
    gen->opt_ = opt;
    gen->params_ = ParamsFromConfig(*gen->config_->config, gen->config_->gen_config.get());
    const LlamaParams& lp = gen->params_;
    gen->tokenizer_ = std::make_unique<Tokenizer>(gen->config_->special_tokens_map, gen->config_->tok_config, gen->config_->tok);

    b2l_params p{};
    p.hidden_size = lp.hidden_size;
    p.intermediate_size = lp.intermediate_size;
    p.num_layers = lp.num_hidden_layers;
    p.num_heads = lp.num_attention_heads;
    p.num_kv_heads = lp.num_key_value_heads;
    p.head_dim = lp.head_dim;
    p.vocab_size = lp.vocab_size;
    p.tie_word_embeddings = lp.tie_word_embeddings ? 1 : 0;
    p.rms_norm_eps = lp.rms_norm_eps;
    p.max_batch = std::max(1, std::min(opt.max_batch, 64));
    p.max_positions = std::min(opt.max_positions, lp.max_position_embeddings > 0 ? lp.max_position_embeddings : opt.max_positions);
    p.page_size = opt.page_size;
    const int blocks = (p.max_positions + p.page_size - 1) / p.page_size;
    p.num_pages = opt.num_pages > 0 ? opt.num_pages : blocks * p.max_batch;
    p.max_prefill_tokens = p.max_positions;
    p.tp_rank = 0;
    p.tp_size = 1;
    p.device = opt.device;

    const std::vector<float> rope = RopeTable(lp, p.max_positions);
    b2l_ctx* ctx = nullptr;
    if (b2l_create(&p, rope.data(), nullptr, &ctx) != 0) throw std::runtime_error(std::string("b2l_create: ") + b2l_last_error(nullptr));
    gen->ctx_ = ctx;

    // every tensor, straight from the mapping (the reference maps the file but never reads a byte)
    // Checkpoints carry tensors the forward does not use (rotary_emb.inv_freq buffers, usually F32; a tied model that
    // also ships lm_head.weight): those are skipped, b2l_finalize still verifies that every tensor it NEEDS arrived.
    for (const std::string& name : gen->config_->tensors.names()) {
        if (!b2l_is_model_tensor(name.c_str())) continue;
        if (name == "lm_head.weight" && lp.tie_word_embeddings) continue;
        const TensorView t = gen->config_->tensors.tensor(name);
        if (t.dtype != "BF16") throw std::runtime_error("tensor " + name + ": dtype " + t.dtype + " is not BF16");
        gen->Check(b2l_upload_tensor(ctx, name.c_str(), t.data, t.shape.data(), static_cast<int>(t.shape.size())), name.c_str());
    }
    gen->Check(b2l_finalize(ctx), "b2l_finalize");
    gen->kv_ = std::make_unique<KvPageAllocator>(p.num_pages, p.page_size, blocks);
    gen->capacity_ = std::min(blocks * p.page_size, p.max_positions);
    gen->batch_engine_ = std::make_unique<B2lBatchEngine>(ctx);
    SchedulerLimits lim;
    lim.max_batch = p.max_batch;
    lim.max_positions = gen->capacity_;
    lim.max_prefill_tokens = p.max_prefill_tokens;
    gen->sched_ = std::make_unique<BatchScheduler>(gen->batch_engine_.get(), gen->kv_.get(), lim, lp.eos_token_ids);
    gen->stepper_ = std::thread([g = gen.get()] { g->StepLoop(); });
    return gen;
}

void Llama3Generator::StepLoop() {
    std::unique_lock<std::mutex> lock(mu_);
    for (;;) {
        work_cv_.wait(lock, [&] { return stop_ || !sched_->idle(); });
        if (stop_) return;
        try {
            sched_->Step();   // admission + one decode step of every running sequence (b2l calls return when the GPU is done)
        } catch (const std::exception&) {
            // the scheduler has already failed the requests that were in flight (SchedResult.error); their callers rethrow
        }
        done_cv_.notify_all();
    }
}

SchedulerStats Llama3Generator::scheduler_stats() {
    std::lock_guard<std::mutex> lock(mu_);
    return sched_->stats();
}

GenerationResult Llama3Generator::GenerateTokens(const std::vector<int32_t>& prompt, int max_new_tokens, bool device_loop) {
    std::unique_lock<std::mutex> lock(mu_);
    if (prompt.empty()) throw std::runtime_error("empty prompt");
    if (static_cast<int>(prompt.size()) >= capacity_) throw std::runtime_error("prompt does not fit the context capacity");
    max_new_tokens = std::max(1, std::min(max_new_tokens, capacity_ - static_cast<int>(prompt.size())));

    GenerationResult res;
    res.prompt_tokens = static_cast<int>(prompt.size());
    if (!device_loop) {
        // the serving path: queue the request, let the stepping thread batch it with whatever else is running
        int id;
        try {
            id = sched_->Submit(prompt, max_new_tokens);
        } catch (const std::invalid_argument& e) {
            throw std::runtime_error(e.what());
        }
        work_cv_.notify_one();
        done_cv_.wait(lock, [&] { return sched_->Result(id).done; });
        const SchedResult r = sched_->Result(id);
        sched_->Forget(id);
        if (!r.error.empty()) throw std::runtime_error(r.error);
        res.tokens = r.tokens;
        res.finish = r.finish;
        return res;
    }
    // exclusive device-resident loop: wait until the batch queue has drained, then keep the lock for the whole generation
    done_cv_.wait(lock, [&] { return sched_->idle(); });
    GreedySampler sampler(params_.eos_token_ids, max_new_tokens);
    const int seq = kv_->NewSequence();
    struct Release {
        KvPageAllocator* kv;
        int seq;
        ~Release() { kv->Free(seq); }
    } release{kv_.get(), seq};
    kv_->Reserve(seq, static_cast<int>(prompt.size()) + max_new_tokens);   // throws KvOutOfPages -> HTTP 500
    const std::vector<int32_t> bt = kv_->BatchTable({seq});
    const int mb = kv_->max_blocks();

    int32_t q_len = static_cast<int32_t>(prompt.size()), ctx_len = 0, next = 0;
    Check(b2l_prefill(ctx_, 1, prompt.data(), &q_len, &ctx_len, bt.data(), mb, &next), "b2l_prefill");
    int32_t pos = q_len;
    FinishReason fin = sampler.Accept(next);
    if (fin == FinishReason::kNone) {
        // the token feedback stays on the device; EOS is looked for afterwards (greedy decoding is
        // deterministic, so tokens past an EOS are simply dropped)
        const int steps = max_new_tokens - 1;
        std::vector<int32_t> ids(static_cast<size_t>(steps));
        if (steps > 0) Check(b2l_decode_loop(ctx_, 1, &next, &pos, bt.data(), mb, steps, ids.data(), nullptr), "b2l_decode_loop");
        for (int i = 0; i < steps && fin == FinishReason::kNone; i++) fin = sampler.Accept(ids[i]);
    }
    res.tokens = sampler.tokens();
    res.finish = fin;
    return res;
}

GenerationResult Llama3Generator::GenerateDetailed(const Request& req, int max_tokens) {
    const std::vector<int32_t> prompt = tokenizer_->ChatPrompt(req.system_message.content, req.user_message.content);
    GenerationResult r = GenerateTokens(prompt, max_tokens > 0 ? max_tokens : opt_.max_new_tokens, /*device_loop=*/false);
    r.text = tokenizer_->Detokenize(r.tokens);
    return r;
}

Message Llama3Generator::Generate(const Request& req) {
    return Message{.role = "assistant", .content = GenerateDetailed(req, 0).text};
}

}  // namespace inference
}  // namespace gabby

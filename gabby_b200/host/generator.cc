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
    p.max_batch = 1;
    p.max_positions = std::min(opt.max_positions, lp.max_position_embeddings > 0 ? lp.max_position_embeddings : opt.max_positions);
    p.page_size = opt.page_size;
    const int blocks = (p.max_positions + p.page_size - 1) / p.page_size;
    p.num_pages = opt.num_pages > 0 ? opt.num_pages : blocks;
    p.max_prefill_tokens = p.max_positions;
    p.tp_rank = 0;
    p.tp_size = 1;
    p.device = opt.device;

    const std::vector<float> rope = RopeTable(lp, p.max_positions);
    b2l_ctx* ctx = nullptr;
    if (b2l_create(&p, rope.data(), nullptr, &ctx) != 0) throw std::runtime_error(std::string("b2l_create: ") + b2l_last_error(nullptr));
    gen->ctx_ = ctx;

    // every tensor, straight from the mapping (the reference maps the file but never reads a byte)
    for (const std::string& name : gen->config_->tensors.names()) {
        const TensorView t = gen->config_->tensors.tensor(name);
        if (t.dtype != "BF16") throw std::runtime_error("tensor " + name + ": dtype " + t.dtype + " is not BF16");
        gen->Check(b2l_upload_tensor(ctx, name.c_str(), t.data, t.shape.data(), static_cast<int>(t.shape.size())), name.c_str());
    }
    gen->Check(b2l_finalize(ctx), "b2l_finalize");
    gen->kv_ = std::make_unique<KvPageAllocator>(p.num_pages, p.page_size, blocks);
    return gen;
}

GenerationResult Llama3Generator::GenerateTokens(const std::vector<int32_t>& prompt, int max_new_tokens, bool device_loop) {
    std::lock_guard<std::mutex> lock(mu_);
    if (prompt.empty()) throw std::runtime_error("empty prompt");
    const int capacity = kv_->max_blocks() * kv_->page_size();
    if (static_cast<int>(prompt.size()) >= capacity) throw std::runtime_error("prompt does not fit the context capacity");
    max_new_tokens = std::max(1, std::min(max_new_tokens, capacity - static_cast<int>(prompt.size())));

    GenerationResult res;
    res.prompt_tokens = static_cast<int>(prompt.size());
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
    if (device_loop && fin == FinishReason::kNone) {
        // the token feedback stays on the device; EOS is looked for afterwards (greedy decoding is
        // deterministic, so tokens past an EOS are simply dropped)
        const int steps = max_new_tokens - 1;
        std::vector<int32_t> ids(static_cast<size_t>(steps));
        if (steps > 0) Check(b2l_decode_loop(ctx_, 1, &next, &pos, bt.data(), mb, steps, ids.data(), nullptr), "b2l_decode_loop");
        for (int i = 0; i < steps && fin == FinishReason::kNone; i++) fin = sampler.Accept(ids[i]);
    } else {
        while (fin == FinishReason::kNone) {
            const int32_t tok = next;
            Check(b2l_decode(ctx_, 1, &tok, &pos, bt.data(), mb, &next), "b2l_decode");
            pos++;
            fin = sampler.Accept(next);
        }
    }
    res.tokens = sampler.tokens();
    res.finish = fin;
    return res;
}

Message Llama3Generator::Generate(const Request& req) {
    const std::vector<int32_t> prompt = tokenizer_->ChatPrompt(req.system_message.content, req.user_message.content);
    const GenerationResult r = GenerateTokens(prompt, opt_.max_new_tokens, /*device_loop=*/false);
    return Message{.role = "assistant", .content = tokenizer_->Detokenize(r.tokens)};
}

}  // namespace inference
}  // namespace gabby

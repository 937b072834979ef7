// host_capi.cc -- C view of the host layer (include/gabby_b200_host.h).
#include <algorithm>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>

#include "config.h"
#include "gabby_b200_host.h"
#include "generator.h"
#include "kv_allocator.h"
#include "params.h"
#include "sampler.h"
#include "scheduler.h"
#include "tokenizer.h"

using namespace gabby::inference;

struct gb_generator {
    std::unique_ptr<Llama3Generator> gen;
};
struct gb_kv {
    std::unique_ptr<KvPageAllocator> kv;
};
struct gb_tokenizer {
    std::unique_ptr<Tokenizer> tok;
};

namespace {
thread_local std::string g_error;

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

void FillParams(const LlamaParams& p, gb_params* o) {
    std::memset(o, 0, sizeof(*o));
    o->hidden_size = p.hidden_size; o->intermediate_size = p.intermediate_size; o->num_hidden_layers = p.num_hidden_layers;
    o->num_attention_heads = p.num_attention_heads; o->num_key_value_heads = p.num_key_value_heads; o->head_dim = p.head_dim;
    o->vocab_size = p.vocab_size; o->tie_word_embeddings = p.tie_word_embeddings; o->max_position_embeddings = p.max_position_embeddings;
    o->bos_token_id = p.bos_token_id;
    o->n_eos = static_cast<int32_t>(std::min<size_t>(8, p.eos_token_ids.size()));
    for (int i = 0; i < o->n_eos; i++) o->eos_token_ids[i] = p.eos_token_ids[i];
    o->rope_llama3 = p.rope_scaling.llama3; o->rope_original_max_position = p.rope_scaling.original_max_position_embeddings;
    o->rms_norm_eps = p.rms_norm_eps; o->rope_theta = p.rope_theta; o->rope_factor = p.rope_scaling.factor;
    o->rope_low_freq_factor = p.rope_scaling.low_freq_factor; o->rope_high_freq_factor = p.rope_scaling.high_freq_factor;
}
}  // namespace

extern "C" {

const char* gb_last_error(void) { return g_error.c_str(); }

int gb_rope_table(double rope_theta, int llama3, double factor, double low_freq_factor, double high_freq_factor,
                  int original_max_position, int head_dim, int max_pos, float* out) {
    return guarded([&] {
        LlamaParams p;
        p.head_dim = head_dim;
        p.rope_theta = rope_theta;
        p.rope_scaling.llama3 = llama3 != 0;
        p.rope_scaling.factor = factor;
        p.rope_scaling.low_freq_factor = low_freq_factor;
        p.rope_scaling.high_freq_factor = high_freq_factor;
        p.rope_scaling.original_max_position_embeddings = original_max_position;
        if (head_dim <= 0 || head_dim % 2 || max_pos <= 0 || !out) throw std::runtime_error("gb_rope_table: bad argument");
        const std::vector<float> t = RopeTable(p, max_pos);
        std::memcpy(out, t.data(), t.size() * sizeof(float));
    });
}

int gb_generator_load(const char* model_dir, int device, int max_positions, int max_new_tokens, gb_generator** out) {
    return guarded([&] {
        if (!model_dir || !out) throw std::runtime_error("gb_generator_load: null argument");
        GeneratorOptions opt;
        opt.device = device;
        if (max_positions > 0) opt.max_positions = max_positions;
        if (max_new_tokens > 0) opt.max_new_tokens = max_new_tokens;
        auto g = std::make_unique<gb_generator>();
        g->gen = Llama3Generator::Load(LoadConfig(model_dir), opt);
        *out = g.release();
    });
}

void gb_generator_free(gb_generator* g) { delete g; }

int gb_generator_generate(gb_generator* g, const char* system_text, const char* user_text, char* out, int cap) {
    return guarded([&] {
        if (!g || !out || cap <= 0) throw std::runtime_error("gb_generator_generate: bad argument");
        const Message m = g->gen->Generate(Request{Message{"system", system_text ? system_text : ""},
                                                   Message{"user", user_text ? user_text : ""}});
        std::snprintf(out, static_cast<size_t>(cap), "%s", m.content.c_str());
    });
}

int gb_generator_generate_detailed(gb_generator* g, const char* system_text, const char* user_text, int max_tokens, char* out, int cap,
                                   int* prompt_tokens, int* completion_tokens, int* finish) {
    return guarded([&] {
        if (!g || !out || cap <= 0) throw std::runtime_error("gb_generator_generate_detailed: bad argument");
        const GenerationResult r = g->gen->GenerateDetailed(Request{Message{"system", system_text ? system_text : ""},
                                                                     Message{"user", user_text ? user_text : ""}}, max_tokens);
        std::snprintf(out, static_cast<size_t>(cap), "%s", r.text.c_str());
        if (prompt_tokens) *prompt_tokens = r.prompt_tokens;
        if (completion_tokens) *completion_tokens = static_cast<int>(r.tokens.size());
        if (finish) *finish = r.finish == FinishReason::kStop ? 1 : r.finish == FinishReason::kLength ? 2 : 0;
    });
}

int gb_generator_generate_ids(gb_generator* g, const int32_t* prompt, int n_prompt, int max_new_tokens, int device_loop,
                              int32_t* out_ids, int out_cap, int* n_out, int* finish) {
    return guarded([&] {
        if (!g || !prompt || !out_ids || !n_out || out_cap < 0) throw std::runtime_error("gb_generator_generate_ids: bad argument");
        const GenerationResult r = g->gen->GenerateTokens(std::vector<int32_t>(prompt, prompt + n_prompt), max_new_tokens, device_loop != 0);
        *n_out = static_cast<int>(r.tokens.size());   // the true count; at most out_cap ids are copied
        std::memcpy(out_ids, r.tokens.data(), std::min(r.tokens.size(), static_cast<size_t>(out_cap)) * sizeof(int32_t));
        if (finish) *finish = r.finish == FinishReason::kStop ? 1 : r.finish == FinishReason::kLength ? 2 : 0;
    });
}

int gb_generator_sched_stats(gb_generator* g, int64_t* out) {
    return guarded([&] {
        if (!g || !out) throw std::runtime_error("gb_generator_sched_stats: null argument");
        const SchedulerStats st = g->gen->scheduler_stats();
        const int64_t v[8] = {st.steps, st.prefill_calls, st.decode_calls, st.prefill_tokens, st.decode_tokens, st.preemptions, st.max_concurrent, -1};
        std::memcpy(out, v, sizeof(v));
    });
}

void* gb_generator_engine(gb_generator* g) { return g ? g->gen->engine() : nullptr; }

int gb_params_from_dir(const char* model_dir, gb_params* out) {
    return guarded([&] {
        auto cfg = LoadConfig(model_dir);
        FillParams(ParamsFromConfig(*cfg->config, cfg->gen_config.get()), out);
    });
}

int gb_params_from_json(const char* config_json, const char* gen_json, gb_params* out) {
    return guarded([&] {
        auto c = gabby::json::Parse(config_json);
        gabby::json::ValuePtr g = gen_json && *gen_json ? gabby::json::Parse(gen_json) : nullptr;
        FillParams(ParamsFromConfig(*c, g.get()), out);
    });
}

int gb_checkpoint_info(const char* model_dir, int* n_tensors, int* n_files) {
    return guarded([&] {
        const Checkpoint c = Checkpoint::Open(model_dir);
        if (n_tensors) *n_tensors = static_cast<int>(c.names().size());
        if (n_files) *n_files = static_cast<int>(c.num_files());
    });
}

int gb_checkpoint_tensor(const char* model_dir, const char* name, int64_t* shape4, int* ndim, char* dtype8, uint64_t* nbytes,
                         uint64_t* fnv1a64) {
    return guarded([&] {
        const Checkpoint c = Checkpoint::Open(model_dir);
        const TensorView t = c.tensor(name);
        if (t.shape.size() > 4) throw std::runtime_error("more than 4 dims");
        for (size_t i = 0; i < t.shape.size(); i++) shape4[i] = t.shape[i];
        *ndim = static_cast<int>(t.shape.size());
        std::snprintf(dtype8, 8, "%s", t.dtype.c_str());
        *nbytes = t.nbytes;
        uint64_t h = 0xcbf29ce484222325ull;
        for (size_t i = 0; i < t.nbytes; i++) h = (h ^ t.data[i]) * 0x100000001b3ull;
        *fnv1a64 = h;
    });
}

int gb_kv_create(int num_pages, int page_size, int max_blocks, gb_kv** out) {
    return guarded([&] {
        auto k = std::make_unique<gb_kv>();
        k->kv = std::make_unique<KvPageAllocator>(num_pages, page_size, max_blocks);
        *out = k.release();
    });
}
void gb_kv_free(gb_kv* kv) { delete kv; }
int gb_kv_new_sequence(gb_kv* kv, int* seq) {
    return guarded([&] { *seq = kv->kv->NewSequence(); });
}
int gb_kv_reserve(gb_kv* kv, int seq, int total_tokens) {
    return guarded([&] { kv->kv->Reserve(seq, total_tokens); });
}
int gb_kv_release(gb_kv* kv, int seq) {
    return guarded([&] { kv->kv->Free(seq); });
}
int gb_kv_table(gb_kv* kv, int seq, int32_t* out, int cap, int* n_blocks) {
    return guarded([&] {
        const auto& t = kv->kv->BlockTable(seq);
        *n_blocks = static_cast<int>(t.size());
        for (size_t i = 0; i < t.size() && static_cast<int>(i) < cap; i++) out[i] = t[i];
    });
}
int gb_kv_free_pages(gb_kv* kv) { return kv->kv->free_pages(); }

int gb_tokenizer_create(const char* tokenizer_json, gb_tokenizer** out) {
    return guarded([&] {
        gabby::json::ValuePtr tok = tokenizer_json && *tokenizer_json ? gabby::json::Parse(tokenizer_json) : gabby::json::Value::MakeNil();
        auto t = std::make_unique<gb_tokenizer>();
        t->tok = std::make_unique<Tokenizer>(nullptr, nullptr, tok);
        *out = t.release();
    });
}
void gb_tokenizer_free(gb_tokenizer* t) { delete t; }
int gb_tokenize(gb_tokenizer* t, const char* text, int32_t* out, int cap, int* n) {
    return guarded([&] {
        const std::vector<int> ids = t->tok->Tokenize(text ? text : "");
        *n = static_cast<int>(ids.size());
        for (size_t i = 0; i < ids.size() && static_cast<int>(i) < cap; i++) out[i] = ids[i];
    });
}
int gb_detokenize(gb_tokenizer* t, const int32_t* ids, int n, char* out, int cap) {
    return guarded([&] {
        const std::string s = t->tok->Detokenize(std::vector<int32_t>(ids, ids + n));
        std::snprintf(out, static_cast<size_t>(cap), "%s", s.c_str());
    });
}
int gb_chat_prompt(gb_tokenizer* t, const char* system_text, const char* user_text, int32_t* out, int cap, int* n) {
    return guarded([&] {
        const std::vector<int32_t> ids = t->tok->ChatPrompt(system_text ? system_text : "", user_text ? user_text : "");
        *n = static_cast<int>(ids.size());
        for (size_t i = 0; i < ids.size() && static_cast<int>(i) < cap; i++) out[i] = ids[i];
    });
}

int32_t gb_argmax(const float* logits, int64_t n) { return GreedySampler::Argmax(logits, n); }

namespace {
// deterministic stand-in for the engine: the next token depends only on (last token, its position), so any batching,
// admission order or preemption must reproduce the sequential result exactly
class FakeBatchEngine : public BatchEngine {
public:
    explicit FakeBatchEngine(int vocab) : vocab_(vocab) {}
    int32_t Next(int32_t last, int pos) const { return static_cast<int32_t>((31ll * last + 7ll * pos + 3) % vocab_); }
    void Prefill(int n_seq, const int32_t* tokens, const int32_t* q_lens, const int32_t* start_pos, const int32_t*, int, int32_t* next_ids) override {
        int off = 0;
        for (int i = 0; i < n_seq; i++) {
            off += q_lens[i];
            next_ids[i] = Next(tokens[off - 1], start_pos[i] + q_lens[i] - 1);
        }
    }
    void Decode(int n_seq, const int32_t* tokens, const int32_t* positions, const int32_t*, int, int32_t* next_ids) override {
        for (int i = 0; i < n_seq; i++) next_ids[i] = Next(tokens[i], positions[i]);
    }

private:
    int vocab_;
};
}  // namespace

struct gb_sched {
    std::unique_ptr<BatchEngine> engine;
    std::unique_ptr<KvPageAllocator> kv;
    std::unique_ptr<BatchScheduler> sched;
};

static gb_sched* make_sched(std::unique_ptr<BatchEngine> engine, std::vector<int> eos, int max_batch, int max_positions,
                            int max_prefill_tokens, int num_pages, int page_size) {
    auto s = std::make_unique<gb_sched>();
    s->engine = std::move(engine);
    s->kv = std::make_unique<KvPageAllocator>(num_pages, page_size, (max_positions + page_size - 1) / page_size);
    SchedulerLimits lim;
    lim.max_batch = max_batch;
    lim.max_positions = max_positions;
    lim.max_prefill_tokens = max_prefill_tokens;
    s->sched = std::make_unique<BatchScheduler>(s->engine.get(), s->kv.get(), lim, std::move(eos));
    return s.release();
}

int gb_sched_create_b2l(void* engine, const int32_t* eos_ids, int n_eos, int max_batch, int max_positions, int max_prefill_tokens,
                        int num_pages, int page_size, gb_sched** out) {
    return guarded([&] {
        if (!engine || !out || (n_eos > 0 && !eos_ids)) throw std::runtime_error("gb_sched_create_b2l: null argument");
        std::vector<int> eos(eos_ids, eos_ids + std::max(0, n_eos));
        *out = make_sched(std::make_unique<B2lBatchEngine>(static_cast<b2l_ctx*>(engine)), std::move(eos), max_batch, max_positions,
                          max_prefill_tokens, num_pages, page_size);
    });
}
int gb_sched_create_fake(int vocab, int eos_id, int max_batch, int max_positions, int max_prefill_tokens, int num_pages,
                         int page_size, gb_sched** out) {
    return guarded([&] {
        if (!out || vocab < 2) throw std::runtime_error("gb_sched_create_fake: bad argument");
        *out = make_sched(std::make_unique<FakeBatchEngine>(vocab), std::vector<int>{eos_id}, max_batch, max_positions, max_prefill_tokens,
                          num_pages, page_size);
    });
}
void gb_sched_free(gb_sched* s) { delete s; }
int gb_sched_submit(gb_sched* s, const int32_t* prompt, int n_prompt, int max_new_tokens, int* id) {
    return guarded([&] {
        if (!s || !id || (n_prompt > 0 && !prompt)) throw std::runtime_error("gb_sched_submit: null argument");
        *id = s->sched->Submit(std::vector<int32_t>(prompt, prompt + std::max(0, n_prompt)), max_new_tokens);
    });
}
int gb_sched_step(gb_sched* s, int* progressed) {
    return guarded([&] {
        if (!s) throw std::runtime_error("gb_sched_step: null argument");
        const int n = s->sched->Step();
        if (progressed) *progressed = n;
    });
}
int gb_sched_drain(gb_sched* s) {
    return guarded([&] {
        if (!s) throw std::runtime_error("gb_sched_drain: null argument");
        s->sched->Drain();
    });
}
int gb_sched_result(gb_sched* s, int id, int32_t* out, int cap, int* n, int* finish, int* done) {
    return guarded([&] {
        if (!s || !n) throw std::runtime_error("gb_sched_result: null argument");
        const SchedResult& r = s->sched->Result(id);
        *n = static_cast<int>(r.tokens.size());
        for (size_t i = 0; out && i < r.tokens.size() && static_cast<int>(i) < cap; i++) out[i] = r.tokens[i];
        if (finish) *finish = r.finish == FinishReason::kStop ? 1 : r.finish == FinishReason::kLength ? 2 : 0;
        if (done) *done = r.done ? 1 : 0;
    });
}
int gb_sched_stats(gb_sched* s, int64_t* out) {
    return guarded([&] {
        if (!s || !out) throw std::runtime_error("gb_sched_stats: null argument");
        const SchedulerStats& st = s->sched->stats();
        out[0] = st.steps; out[1] = st.prefill_calls; out[2] = st.decode_calls; out[3] = st.prefill_tokens;
        out[4] = st.decode_tokens; out[5] = st.preemptions; out[6] = st.max_concurrent; out[7] = s->kv->free_pages();
    });
}

}  // extern "C"

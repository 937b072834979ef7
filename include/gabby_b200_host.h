/*
 * gabby_b200_host.h -- C view of the host C++ layer (gabby_b200/host/), for callers that cannot
 * include C++ headers (tests, bench.py). The C++ interface itself mirrors gabby's:
 * gabby::inference::{Message, Request, Generator, Llama3Generator::Load, InferenceConfig, LoadConfig,
 * Safetensors, Tokenizer} (/root/reference/src/inference/*.h) -- see gabby_b200/host/generator.h.
 *
 * All int functions return 0 on success; gb_last_error() holds the message otherwise (the text of
 * the C++ exception: what gabby's server would turn into HTTP 400/500, src/http/server.cc:371-378).
 */
#ifndef GABBY_B200_HOST_H_
#define GABBY_B200_HOST_H_

#include <stdint.h>

/* libgabby_host.so is built with hidden visibility and -Bsymbolic: only these entry points are exported, so the library
 * can be linked into gabby itself -- whose own gabby::inference:: / gabby::json:: symbols it mirrors -- without clashing
 * (integration/ does exactly that). */
#define GB_API __attribute__((visibility("default")))

#ifdef __cplusplus
extern "C" {
#endif

GB_API const char* gb_last_error(void);

/* RoPE (cos, sin) table [max_pos][head_dim/2][2] fp32 for b2l_create, from config.json's
 * rope_theta / rope_scaling (llama3 != 0: apply the llama3 band rescaling). */
GB_API int gb_rope_table(double rope_theta, int llama3, double factor, double low_freq_factor, double high_freq_factor,
                  int original_max_position, int head_dim, int max_pos, float* out);

/* ---- Llama3Generator (generator.h) ---- */
typedef struct gb_generator gb_generator;
/* LoadConfig(model_dir) + Llama3Generator::Load: parses the JSON files, maps the safetensors
 * (single file or sharded), uploads every tensor, sizes the paged KV pool. */
GB_API int gb_generator_load(const char* model_dir, int device, int max_positions, int max_new_tokens, gb_generator** out);
GB_API void gb_generator_free(gb_generator* g);
/* Generator::Generate(Request{system, user}) -> Message.content (NUL-terminated, truncated to cap) */
GB_API int gb_generator_generate(gb_generator* g, const char* system_text, const char* user_text, char* out, int cap);
/* The same call with what an OpenAI-style response needs besides the text (gabby hard-codes these: /root/reference/src/
 * service.cc:79-116): max_tokens (<= 0: the default given to gb_generator_load), usage counts and the finish reason
 * (1 = EOS -> "stop", 2 = token budget -> "length"). Safe to call from many threads: concurrent requests are batched. */
GB_API int gb_generator_generate_detailed(gb_generator* g, const char* system_text, const char* user_text, int max_tokens, char* out, int cap,
                                   int* prompt_tokens, int* completion_tokens, int* finish);
/* token-level: prefill `prompt` then greedy decode; at most out_cap ids are copied, *n_out is the number generated;
 * finish: 1 = EOS ("stop"), 2 = max_new_tokens ("length") */
GB_API int gb_generator_generate_ids(gb_generator* g, const int32_t* prompt, int n_prompt, int max_new_tokens, int device_loop,
                              int32_t* out_ids, int out_cap, int* n_out, int* finish);
/* out[8] of the generator's continuous-batching queue: steps, prefill_calls, decode_calls, prefill_tokens, decode_tokens,
 * preemptions, max_concurrent, free_pages (-1) */
GB_API int gb_generator_sched_stats(gb_generator* g, int64_t* out);
GB_API void* gb_generator_engine(gb_generator* g); /* the b2l_ctx* behind it (parity taps) */

/* ---- pieces, exposed for CPU-side tests ---- */
typedef struct {
    int32_t hidden_size, intermediate_size, num_hidden_layers, num_attention_heads, num_key_value_heads, head_dim, vocab_size;
    int32_t tie_word_embeddings, max_position_embeddings, bos_token_id, n_eos, eos_token_ids[8];
    int32_t rope_llama3, rope_original_max_position;
    float rms_norm_eps;
    double rope_theta, rope_factor, rope_low_freq_factor, rope_high_freq_factor;
} gb_params;
GB_API int gb_params_from_dir(const char* model_dir, gb_params* out);                 /* ParamsFromConfig(LoadConfig(dir)) */
GB_API int gb_params_from_json(const char* config_json, const char* gen_json, gb_params* out);

/* Checkpoint accessor: number of tensors / files, and one tensor's metadata + a 64-bit FNV-1a of its bytes */
GB_API int gb_checkpoint_info(const char* model_dir, int* n_tensors, int* n_files);
GB_API int gb_checkpoint_tensor(const char* model_dir, const char* name, int64_t* shape4, int* ndim, char* dtype8, uint64_t* nbytes,
                         uint64_t* fnv1a64);

/* KvPageAllocator */
typedef struct gb_kv gb_kv;
GB_API int gb_kv_create(int num_pages, int page_size, int max_blocks, gb_kv** out);
GB_API void gb_kv_free(gb_kv* kv);
GB_API int gb_kv_new_sequence(gb_kv* kv, int* seq);
GB_API int gb_kv_reserve(gb_kv* kv, int seq, int total_tokens);
GB_API int gb_kv_release(gb_kv* kv, int seq);
GB_API int gb_kv_table(gb_kv* kv, int seq, int32_t* out, int cap, int* n_blocks);
GB_API int gb_kv_free_pages(gb_kv* kv);

/* Tokenizer built from a tokenizer.json text (NULL/"" = byte fallback) */
typedef struct gb_tokenizer gb_tokenizer;
GB_API int gb_tokenizer_create(const char* tokenizer_json, gb_tokenizer** out);
GB_API void gb_tokenizer_free(gb_tokenizer* t);
GB_API int gb_tokenize(gb_tokenizer* t, const char* text, int32_t* out, int cap, int* n);
GB_API int gb_detokenize(gb_tokenizer* t, const int32_t* ids, int n, char* out, int cap);
GB_API int gb_chat_prompt(gb_tokenizer* t, const char* system_text, const char* user_text, int32_t* out, int cap, int* n);

/* GreedySampler::Argmax (first max) on host logits */
GB_API int32_t gb_argmax(const float* logits, int64_t n);

/* ---- continuous batching (gabby_b200/host/scheduler.h). The reference answers one
 * request at a time (/root/reference/src/service.cc:150 under /root/reference/src/http/thread_pool.cc:22-28); this
 * queue feeds the engine ragged multi-sequence prefill / decode steps. ---- */
typedef struct gb_sched gb_sched;
/* over a real engine: `engine` is the b2l_ctx* (created with max_batch >= the scheduler's); the scheduler owns a
 * KvPageAllocator over that engine's page pool (num_pages, page_size as given to b2l_create) */
GB_API int gb_sched_create_b2l(void* engine, const int32_t* eos_ids, int n_eos, int max_batch, int max_positions, int max_prefill_tokens,
                        int num_pages, int page_size, gb_sched** out);
/* over a deterministic fake engine (next = (31 * last + 7 * position + 3) mod vocab): policy tests without a GPU */
GB_API int gb_sched_create_fake(int vocab, int eos_id, int max_batch, int max_positions, int max_prefill_tokens, int num_pages,
                         int page_size, gb_sched** out);
GB_API void gb_sched_free(gb_sched* s);
GB_API int gb_sched_submit(gb_sched* s, const int32_t* prompt, int n_prompt, int max_new_tokens, int* id);
GB_API int gb_sched_step(gb_sched* s, int* progressed);
GB_API int gb_sched_drain(gb_sched* s);
/* finish: 0 none, 1 stop (EOS), 2 length */
GB_API int gb_sched_result(gb_sched* s, int id, int32_t* out, int cap, int* n, int* finish, int* done);
/* out[8]: steps, prefill_calls, decode_calls, prefill_tokens, decode_tokens, preemptions, max_concurrent, free_pages */
GB_API int gb_sched_stats(gb_sched* s, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif

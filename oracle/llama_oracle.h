/*
 * llama_oracle.h -- CPU restatement of the Llama-3 forward pass. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or load this. The product (gabby_b200/) never does.
 *
 * What it restates: the body gabby's Llama3Generator::Generate is meant to have
 * (/root/reference/src/inference/generator.cc:33-38 is a constant-string stub; the
 * reference has NO forward pass, README.md:18-27). The math therefore follows the public
 * Llama-3 definition as HF transformers 5.5.0 implements it
 * (transformers/models/llama/modeling_llama.py: RMSNorm :53-70, rotate-half RoPE :138-168,
 * SwiGLU MLP :171-184, GQA attention :187-222, decoder layer :292+; llama3 rope scaling in
 * modeling_rope_utils.py `_compute_llama3_parameters`). Label: "scalar fp32 C++
 * restatement -- not reference code".
 *
 * PARITY PINNING: the reference's own tests hold no golden vector for this path
 * (tokenizer_test.cc:9-25 asserts Tokenize("") == {}; service_test.cc uses a fake
 * generator) => "parity unpinned" by the reference. The oracle is instead pinned to HF
 * transformers fp32 CPU outputs committed under tests/golden/ (tests/golden/make_golden.py).
 */
#ifndef B2L_LLAMA_ORACLE_H_
#define B2L_LLAMA_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t hidden_size;
    int32_t intermediate_size;
    int32_t num_layers;
    int32_t num_heads;
    int32_t num_kv_heads;
    int32_t head_dim;
    int32_t vocab_size;
    int32_t tie_word_embeddings;
    float rms_norm_eps;
    /* rope */
    double rope_theta;
    int32_t rope_llama3; /* 1: apply llama3 frequency rescaling */
    double rope_factor;
    double rope_low_freq_factor;
    double rope_high_freq_factor;
    int32_t rope_original_max_position;
    /* capacity of each sequence's KV cache */
    int32_t max_seq_len;
} orc_params;

/* rounding-point flags: mirror where the CUDA path stores bf16 */
enum {
    ORC_KV_BF16 = 1,  /* K (post-RoPE) and V rounded to bf16 when cached */
    ORC_ACT_BF16 = 2, /* inputs of every linear layer rounded to bf16 (tensor-core prefill path) */
    ORC_QP_BF16 = 4   /* q (post-RoPE, pre-scale) and softmax probabilities rounded to bf16 (flash prefill) */
};

typedef struct orc_model orc_model;
typedef struct orc_seq orc_seq;

orc_model* orc_model_create(const orc_params* p);
/* Borrows `bf16_bits` (must outlive the model). Returns 0, or -1 for unknown name / wrong size. */
int orc_model_set_tensor(orc_model* m, const char* hf_name, const uint16_t* bf16_bits, int64_t numel);
/* 0 when every tensor the architecture needs has been set */
int orc_model_check(const orc_model* m);
void orc_model_destroy(orc_model* m);

orc_seq* orc_seq_create(const orc_model* m, int flags);
void orc_seq_reset(orc_seq* s);
/* change the rounding-point flags for subsequent forwards (prefill on the tensor-core path rounds linear inputs
 * to bf16, decode does not) */
void orc_seq_set_flags(orc_seq* s, int flags);
int orc_seq_len(const orc_seq* s);
void orc_seq_destroy(orc_seq* s);

/*
 * Append `n` tokens at positions len..len+n-1 and run the forward pass.
 *   logits_out : NULL, or [n_logit_rows, vocab] fp32; logits_all=0 -> last token only (1 row)
 *   hidden_out : NULL, or [(num_layers+2), n, hidden] fp32:
 *                slab 0 = embeddings, 1..L = residual stream after each layer, L+1 = final norm
 * Returns 0, or -1 if the KV capacity would be exceeded.
 */
int orc_seq_forward(orc_seq* s, const int32_t* tokens, int n, int logits_all, float* logits_out,
                    float* hidden_out);

/* first-max argmax (ties -> lowest index), the greedy rule */
int32_t orc_argmax(const float* x, int64_t n);

/* prefill `prompt`, then `n_new` greedy steps; out_ids[n_new]. margins (optional): top1-top2 gap */
int orc_greedy(orc_seq* s, const int32_t* prompt, int n_prompt, int n_new, int32_t* out_ids,
               float* margins);

/* [max_pos][head_dim/2][2] = (cos, sin); the table both oracle and CUDA host code use */
void orc_rope_table(const orc_params* p, int max_pos, float* out);
/* inv_freq[head_dim/2] after llama3 rescaling */
void orc_rope_inv_freq(const orc_params* p, float* out);

/* the synthetic-weight counter hash (gabby_b200/synth.py), restated */
void orc_synth_tensor(uint32_t tensor_seed, int64_t n, float scale, float offset, uint16_t* out_bits);

int orc_num_threads(void);
/* launchers such as torchrun export OMP_NUM_THREADS=1: the timed CPU legs set the thread count explicitly */
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif

/*
 * b2l.h -- C-ABI of the B200-native Llama-3 forward ("b2l" = B200 llama).
 *
 * This is the whole surface through which host code reaches CUDA. It is what the body of
 * gabby's `Llama3Generator::Generate` / `Llama3Generator::Load`
 * (/root/reference/src/inference/generator.cc:33-44, generator.h:36-47 -- today a constant
 * string and a no-op) binds to; the reference has no FFI of its own for this path, so each
 * entry point cites the reference interface whose missing body it supplies.
 *
 * Conventions: plain pointers and sizes, caller-owned HOST buffers, opaque handle, every
 * function returns 0 on success and non-zero on failure (never throws); the message is read
 * with b2l_last_error(). The C++ adapter rethrows it as std::runtime_error, which gabby's
 * server maps to HTTP 500 (/root/reference/src/http/server.cc:371-378). One mutex per ctx:
 * calls may come from any HTTP worker thread (/root/reference/src/http/server.h:35).
 *
 * There is NO CPU fallback: without a CUDA device b2l_create fails.
 */
#ifndef B2L_H_
#define B2L_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2L_ABI_VERSION 1
#define B2L_NCCL_ID_BYTES 128

typedef struct b2l_ctx b2l_ctx;

/* Typed form of the config.json keys gabby keeps as untyped JSON
 * (/root/reference/src/inference/config.h:12-19, config.cc:13), plus runtime sizing. */
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
    /* runtime sizing */
    int32_t max_batch;          /* max sequences per decode call (<= 64) */
    int32_t max_positions;      /* rows of the RoPE table = max context per sequence */
    int32_t page_size;          /* tokens per KV page: 16, 32 or 64 */
    int32_t num_pages;          /* pages in the KV pool (per layer) */
    int32_t max_prefill_tokens; /* scratch rows for one prefill call */
    /* tensor parallel (8B/70B shapes); tp_size 1 = single GPU */
    int32_t tp_rank;
    int32_t tp_size;
    int32_t device; /* CUDA ordinal */
} b2l_params;

/* ---- lifecycle: the body of Llama3Generator::Load (generator.cc:41-44) ------------------ */

/* rope_cos_sin: host [max_positions][head_dim/2][2] fp32 (cos, sin), computed by the host layer
 * from rope_theta / rope_scaling. nccl_unique_id: B2L_NCCL_ID_BYTES bytes shared by all ranks
 * (from b2l_nccl_unique_id on rank 0) when tp_size > 1, else NULL. */
int b2l_create(const b2l_params* p, const float* rope_cos_sin, const void* nccl_unique_id, b2l_ctx** out);
int b2l_nccl_unique_id(void* out_bytes);

/* One HF tensor exactly as Safetensors maps it (/root/reference/src/inference/safetensors.cc:
 * 17-36): row-major bf16, full (unsharded) shape. The engine keeps its TP shard and re-lays it
 * out (QKV fused, gate/up row-interleaved). */
int b2l_upload_tensor(b2l_ctx* c, const char* hf_name, const void* host_bf16, const int64_t* shape, int ndim);
/* Same tensor generated on-device from the synthetic counter hash (gabby_b200/synth.py): no
 * host bytes needed (8B/70B timing configs). */
int b2l_synth_tensor(b2l_ctx* c, const char* hf_name, const int64_t* shape, int ndim, uint32_t tensor_seed,
                     float scale, float offset);
/* The tensor-parallel shard of one HF tensor that rank p->tp_rank keeps: rows [win[0], win[0]+win[1]) x
 * cols [win[2], win[2]+win[3]) of the full tensor (1-D tensors: one row). Pure host arithmetic, no
 * device needed -- b2l_upload_tensor / b2l_synth_tensor use exactly this window. q/k/v/gate/up and
 * lm_head are row (output) sharded, o/down column (input) sharded, norms and embeddings replicated. */
int b2l_shard_window(const b2l_params* p, const char* hf_name, const int64_t* shape, int ndim, int64_t win[4]);
/* 1 when `hf_name` is a tensor the forward uses (embeddings, norms, the seven projections of a layer, lm_head), 0 otherwise:
 * checkpoints also carry buffers such as rotary_emb.inv_freq, which a loader skips instead of failing on them
 * (the reference's Safetensors exposes every header entry alike: /root/reference/src/inference/safetensors.h:16-24). */
int b2l_is_model_tensor(const char* hf_name);
/* Checks every tensor arrived, builds derived layouts, captures decode graphs. */
int b2l_finalize(b2l_ctx* c);
void b2l_destroy(b2l_ctx* c);

/* ---- forward: the body of Llama3Generator::Generate (generator.cc:33-38) ---------------- */

/* The paged KV cache is ALLOCATED BY THE HOST (gabby_b200/host/kv_allocator.*): block_tables is
 * [n_seq][max_blocks] page ids; page j of a sequence holds its tokens [j*page_size, (j+1)*page_size). */

/* Prefill: sequence i appends q_lens[i] tokens (concatenated in `tokens`) after ctx_lens[i] cached
 * ones. next_ids[i] = greedy argmax after the last token of sequence i. */
int b2l_prefill(b2l_ctx* c, int n_seq, const int32_t* tokens, const int32_t* q_lens, const int32_t* ctx_lens,
                const int32_t* block_tables, int max_blocks, int32_t* next_ids);
/* One decode step: sequence i feeds tokens[i] at position positions[i] (= its cached length). */
int b2l_decode(b2l_ctx* c, int n_seq, const int32_t* tokens, const int32_t* positions,
               const int32_t* block_tables, int max_blocks, int32_t* next_ids);
/* n_steps greedy steps without leaving the device (token i+1 = argmax of step i). block_tables
 * must already cover positions[i] + n_steps. out_ids is [n_steps][n_seq]. device_ms (optional):
 * CUDA-event time of the loop on the engine's stream. */
int b2l_decode_loop(b2l_ctx* c, int n_seq, const int32_t* tokens, const int32_t* positions,
                    const int32_t* block_tables, int max_blocks, int n_steps, int32_t* out_ids,
                    float* device_ms);

/* ---- parity taps ------------------------------------------------------------------------ */

/* fp32 logits of the last prefill/decode call: rows [row0, row0+n_rows) x vocab_size.
 * Rows = sequences (prefill: last token of each sequence). Under TP the vocab shards are gathered. */
int b2l_get_logits(b2l_ctx* c, int row0, int n_rows, float* out);
/* enable=1: subsequent prefill/decode calls also record the residual stream:
 * slab 0 = embeddings, slab l = after layer l (1..L), slab L+1 = final norm. */
int b2l_set_taps(b2l_ctx* c, int enable);
/* rows of the last call (prefill: every token row, in `tokens` order) x hidden_size fp32 */
int b2l_get_hidden(b2l_ctx* c, int slab, int row0, int n_rows, float* out);
/* raw bf16 K or V (which = 0/1) of one layer / page: [page_size][local_kv_heads*head_dim] */
int b2l_get_kv_page(b2l_ctx* c, int layer, int page, int which, void* out_bf16);

/* ---- introspection ---------------------------------------------------------------------- */

typedef struct {
    int32_t abi_version;
    int32_t sm_count;
    int32_t cc_major, cc_minor;
    int64_t hbm_bytes;
    int64_t weight_bytes;      /* resident on this rank */
    int64_t kv_bytes;          /* KV pool on this rank */
    int64_t stream_bytes_per_token; /* weight bytes one decode step streams (this rank) */
    int64_t kernels_launched;  /* kernels this ctx has launched so far (graph nodes count) */
    int32_t decode_mode;       /* 0 = multi-kernel graph, 1 = persistent megakernel */
    int32_t batched_tensor_core; /* 1: decode steps of 2..16 rows run their projections on tcgen05 (skinny GEMM) */
    int32_t tp_transport;      /* 0 = single rank, 1 = NCCL all-reduce, 2 = fused stores into NVLink peer memory (cudaIpc slabs) */
    char device_name[64];
} b2l_info;
int b2l_get_info(b2l_ctx* c, b2l_info* out);
/* decode implementation: 0 multi-kernel CUDA graph, 1 persistent megakernel (batch 1) */
int b2l_set_decode_mode(b2l_ctx* c, int mode);

/* prefill implementation: 1 = tcgen05/TMEM GEMMs over all new tokens (bf16 activations, fp32 accumulate and
 * residual stream), 0 = 8-token chunks through the decode kernels (fp32 activations), -1 = auto (GEMMs for
 * >= 64 new tokens). */
int b2l_set_prefill_mode(b2l_ctx* c, int mode);

/* Debug: per-phase device timestamps (globaltimer ns) of the last token of the last megakernel
 * launch. enable=1 arms it; out_ns (optional) is [336][n_phases+1], column = phase: rows 0-3 CTA 0
 * {phase entry, dependency barrier passed, input vector loaded, phase end}, rows 4-7 the same for
 * the last CTA, row 8 CTA 0 warp 0 cycles spent waiting for weights.
 * phase_types (optional) [n_phases]: 0 qkv 1 attn 2 o 3 gate/up 4 down 5 lm_head. */
int b2l_debug_mega_profile(b2l_ctx* c, int enable, uint64_t* out_ns, int* n_phases, int32_t* phase_types);

/* message of the last failure on this ctx (or of the last failed b2l_create when c == NULL) */
const char* b2l_last_error(const b2l_ctx* c);

/* ---- single-op entry points (parity tests drive each production kernel in isolation) ---- */

/* y[b][n] (+)= sum_k W[n][k] * x[b][k]; W bf16 row-major [N][K]; x, y fp32.
 * norm_w (nullable, bf16[K]): fused RMSNorm prologue x <- rmsnorm(x) * norm_w.
 * mode 0: store, 1: y += (residual), 2: SwiGLU over row pairs (W rows 2i = gate_i, 2i+1 = up_i;
 * y is [B][N/2]). device_ms (optional): CUDA-event time of `iters` launches / iters. */
int b2l_op_gemv(int device, const void* W_bf16, const float* x, float* y, const void* norm_w_bf16, float eps,
                int B, int N, int K, int mode, int iters, float* device_ms);
int b2l_op_argmax(int device, const float* x, int B, int N, int32_t* out);
/* Prefill GEMM on the tensor cores (tcgen05 + TMEM accumulators, TMA-fed): C = A[M][K] * W[N][K]^T, bf16 inputs,
 * fp32 accumulate. epilogue 0: fp32 store, 1: bf16 store, 2: C += (fp32 residual), 3: SwiGLU over column pairs
 * (W rows 2i = gate_i, 2i+1 = up_i; C is [M][N/2], bf16-rounded). C is fp32 [M][N or N/2] on the host both ways.
 * K % 64 == 0, N % 128 == 0. iters > 0 also times `iters` launches (device_ms = average). */
int b2l_op_gemm_bf16(int device, const void* A_bf16, const void* W_bf16, float* C, int M, int N, int K, int epilogue, int iters,
                     float* device_ms);

#ifdef __cplusplus
}
#endif
#endif /* B2L_H_ */

// engine.h -- device-side state behind the b2l C-ABI (include/b2l.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b2l.h"

namespace b2l {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define B2L_CUDA(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            throw ::b2l::Error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" +    \
                               std::to_string(__LINE__) + ")");                                            \
    } while (0)

#define B2L_CHECK(cond, msg)                      \
    do {                                          \
        if (!(cond)) throw ::b2l::Error(msg);     \
    } while (0)

struct LayerWeights {
    uint16_t* in_norm = nullptr;    // [H]
    uint16_t* w_qkv = nullptr;      // [(nh_l + 2 nkv_l) * hd][H]  rows: q heads, k heads, v heads
    uint16_t* w_o = nullptr;        // [H][nh_l * hd]
    uint16_t* post_norm = nullptr;  // [H]
    uint16_t* w_gu = nullptr;       // [2 I_l][H]  row 2i = gate_i, 2i+1 = up_i
    uint16_t* w_down = nullptr;     // [H][I_l]
    uint16_t* kv_pool = nullptr;    // [num_pages][2][page_size][nkv_l * hd]
    uint32_t have = 0;              // bit per HF tensor received
};

struct Graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int nodes = 0;
};

}  // namespace b2l

struct b2l_ctx {
    b2l_params p{};
    std::mutex mu;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaDeviceProp prop{};
    bool finalized = false;
    int decode_mode = 0;
    bool taps = false;

    // local (per TP rank) dims
    int H = 0, I_l = 0, nh_l = 0, nkv_l = 0, hd = 0, V = 0, V_l = 0, L = 0;
    int qd_l = 0, kvd_l = 0, qkv_l = 0, group = 0;
    int max_rows = 0, max_blocks_cap = 0, nsplit = 0;

    // weights
    uint16_t* embed = nullptr;       // [V][H] (replicated)
    uint16_t* lm_head = nullptr;     // [V_l][H] (aliases embed + rank*V_l*H when tied)
    uint16_t* final_norm = nullptr;  // [H]
    bool have_embed = false, have_final_norm = false, have_lm_head = false;
    std::vector<b2l::LayerWeights> layers;
    float* rope = nullptr;           // [max_positions][hd/2][2]
    int64_t weight_bytes = 0, kv_bytes = 0;

    // per-call inputs (device) + pinned staging
    int32_t *d_tokens = nullptr, *d_positions = nullptr, *d_slots = nullptr, *d_block_tables = nullptr;
    int32_t *d_next_ids = nullptr, *d_out_ids = nullptr, *d_step = nullptr;
    int32_t *h_stage = nullptr;      // pinned: tokens | positions | slots | block tables | next ids
    size_t h_stage_ints = 0;
    int out_ids_cap = 0;

    // activations (fp32)
    float *h = nullptr, *qkv = nullptr, *attn = nullptr, *act = nullptr, *proj = nullptr, *logits = nullptr;
    float *seq_logits = nullptr;     // [max_batch][V_l] (prefill keeps each sequence's last row here)
    float *part_acc = nullptr, *part_ml = nullptr;
    int* attn_counters = nullptr;
    float* tap = nullptr;            // [(L+2)][tap_rows][H]
    int tap_rows_cap = 0, tap_rows = 0;
    const float* logits_src = nullptr;
    int logits_rows = 0;

    // persistent megakernel (decode_mode 1, batch 1)
    bool mega_ok = false;
    std::string mega_why;            // why the megakernel is unavailable for this shape
    void* mega_phases = nullptr;     // device MegaPhase[]
    int mega_n_phases = 0, mega_stages = 0, mega_nsplit = 0;
    size_t mega_smem = 0;
    uint16_t* kv_base = nullptr;     // all layers' KV pools, contiguous
    size_t mega_l2_persist_bytes = 0;
    int mega_attn_tps = 128;
    // dataflow mode: {value, seq} word buffers (h | qkv | act | attention partials | per-CTA argmax keys)
    unsigned long long *mega_ll_h = nullptr, *mega_ll_qkv = nullptr, *mega_ll_act = nullptr, *mega_ll_pacc = nullptr,
                       *mega_ll_pml = nullptr, *mega_ll_keys = nullptr;
    uint32_t mega_seq = 0;                   // sequence numbers consumed by earlier launches
    std::vector<int32_t> bt_uploaded;        // image of the block tables currently on the device (skip identical re-uploads)
    int bt_uploaded_rows = 0;
    int* mega_abort = nullptr;               // pinned host flag, device-visible
    unsigned long long* mega_prof = nullptr; // device [4][n_phases+1] phase timestamps (debug)

    // tensor-core prefill (gemm_tcgen05.cuh): bf16 activations, sized for max_prefill_tokens rows
    int prefill_mode = -1;           // -1 auto (GEMM path for >= 64 new tokens), 0 chunked decode kernels, 1 GEMM path
    bool pf_ok = false;              // shapes satisfy the GEMM constraints
    int pf_rows = 0;
    int32_t *pf_tokens = nullptr, *pf_positions = nullptr, *pf_slots = nullptr, *pf_last = nullptr;
    float *pf_h = nullptr, *pf_qkv = nullptr, *pf_attn = nullptr, *pf_proj = nullptr, *pf_part_acc = nullptr, *pf_part_ml = nullptr;
    uint16_t *pf_xn = nullptr, *pf_attn16 = nullptr, *pf_act16 = nullptr;
    int* pf_counters = nullptr;
    void* pf_tiles = nullptr;        // device PrefillTile[] (flash prefill attention work list)
    int pf_n_tiles = 0;
    void* pf_tiles_tc = nullptr;     // device PrefillTile[] of up to 256 rows, heaviest first (flash_prefill_tc.cuh)
    int pf_n_tiles_tc = 0;
    bool flash_tc_ok = false;        // head_dim / page size fit the tcgen05 attention kernel
    alignas(64) unsigned char kv_map[128] = {};   // CUtensorMap over the whole KV pool (rows of kvd elements, box = one page x 64 dims)

    // batched decode on the tensor cores (skinny_gemm.cuh)
    bool skinny_ok = false;
    bool rope_fuse = false;      // QKV split-K reduce also applies RoPE and appends k, v (skinny_reduce_rope_kv_kernel)
    bool norm_cluster = false;   // residual + RMSNorm + hi/lo split epilogue on a cluster per row
    uint16_t *sk_xh = nullptr, *sk_xq = nullptr, *sk_xi = nullptr;   // bf16 hi/lo activation rows [32][K]
    float* sk_partial = nullptr;                                       // [ksplit][16][N] fp32
    size_t sk_partial_floats = 0;
    std::vector<unsigned char> sk_maps;                                // CUtensorMap storage (host): per layer 4 weights + lm_head, then 3 x maps per BT

    // tensor parallelism (NCCL, loaded with dlopen only when tp_size > 1)
    void* nccl_comm = nullptr;
    float* tp_pack = nullptr;        // [max_rows][2]   (value, index) of this rank's argmax
    float* tp_gather = nullptr;      // [tp][max_rows][2]
    float* tp_logits = nullptr;      // [tp][rows][V_l] gather buffer for b2l_get_logits (lazy)
    float* tp_vals = nullptr;        // [max_rows] local max values
    // row-parallel partial sums over NVLink peer memory (cudaIpc-mapped slabs, LL words): see TpSend
    bool tp_peer_ok = false;         // false: NCCL all-reduce transport
    uint2* tp_ll = nullptr;          // this rank's slabs [2 slots][tp][max_rows][H], then the megakernel's [2][tp][H] and its keys [tp][2 SMs]
    size_t tp_mega_off = 0, tp_keys_off = 0;   // word offsets of the megakernel regions inside tp_ll
    uint2* tp_peer[8] = {};          // every rank's slab base as mapped here (own = tp_ll)
    uint32_t* tp_seq = nullptr;      // [2] sequence number per slot
    unsigned int* tp_done = nullptr; // [2] last-CTA counters of the receiver kernel

    std::map<int, b2l::Graph> decode_graphs;  // key: rows (+ 1000 when the loop variant with advance)
    int64_t launched = 0;

    std::vector<void*> allocs;
};

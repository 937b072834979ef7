// mega_common.cuh -- declarations shared by the two builds of the decode megakernel (mega_decode.cuh is compiled twice:
// MEGA_TP 0 = one GPU, MEGA_TP 1 = one tensor-parallel rank) and by the host side (engine.cu).
#pragma once
#include <stdint.h>

namespace b2l {

constexpr int kMegaConsumerWarps = 8;
constexpr int kMegaConsumerThreads = kMegaConsumerWarps * 32;
constexpr int kMegaThreads = kMegaConsumerThreads + 32;  // + producer warp
constexpr int kMegaStageBytes = 16 * 1024;
constexpr int kMegaMaxStages = 12;
constexpr int kMegaRows = 4;  // rows one warp reduces together (transposed butterfly)
constexpr int kMegaXsFloats = 2048;
constexpr int kMegaProfRows = 16 + 2 * 160;  // debug timeline: 16 summary rows, then per CTA: input-ready and phase-end times

enum MegaPhaseType { PH_QKV = 0, PH_ATTN = 1, PH_OPROJ = 2, PH_GATEUP = 3, PH_DOWN = 4, PH_LMHEAD = 5 };

struct MegaPhase {
    int type, layer;
    const uint16_t* W;       // [N][K] bf16 (PH_ATTN: unused)
    const uint16_t* norm_w;  // fused RMSNorm weight or null
    uint16_t* kv_pool;       // PH_ATTN: this layer's KV pool
    int N, K;
    int ks;                  // warps per row (K split into ks slices of 256*m elements)
    int m;                   // 16-byte sweeps per warp unit: slice = 256 * m elements (<= 2048)
};

struct MegaArgs {
    const MegaPhase* phases;
    int n_phases;
    int n_stages;
    // model
    const uint16_t* embed;
    const float* rope;
    int H, V, nh, nkv, hd, I;
    float eps, attn_scale;
    // activations (fp32, L2 resident)
    float *h, *qkv, *attn, *act, *logits;
    // paged KV
    const int32_t* block_table;
    int page_size, kvd;
    float *part_acc, *part_ml;
    int* attn_counters;
    int nsplit_max;
    // token loop
    int32_t token0, pos0;  // arg_io != 0: first token / position travel in this struct (constant memory) instead of *token / *position
    int arg_io;
    int32_t* token;      // in: first token; out: last argmax
    int32_t* position;   // in: first position; out: advanced
    int32_t* out_ids;    // [n_steps]
    int n_steps;
    // sync
    unsigned long long* bar_counter;  // monotonically increasing arrivals
    unsigned long long* bar_epoch;    // arrivals consumed by previous launches
    unsigned long long* argmax_keys;  // [3]
    int* abort_flag;     // mapped pinned host memory: [0] abort code, [1 + cta] progress marker of each CTA (debug)
    int debug_nostream;  // 1: copy 16 bytes per chunk instead of the weights (timing experiments only; wrong results)
    int debug_progress;  // 1: CTAs record step*100000 + phase*100 + stage-of-phase
    int producer_sleep_ns;
    int attn_tps;       // context tokens per attention split (work item)
    int max_inflight;   // bulk copies issued but not yet landed, per CTA (bounds queueing latency in L2/HBM)
    int l2_ahead;       // chunks prefetched into L2 beyond the ring (0 = off)
    // dataflow mode (kernel template LL): activations travel as 8-byte {fp32 bits, sequence number} words and every
    // reader polls for the sequence number of the phase that produces its input -- no grid barrier anywhere
    unsigned long long *ll_h, *ll_qkv, *ll_act, *ll_pacc, *ll_pml, *ll_keys;
    uint32_t seq_base;  // sequence numbers used by earlier launches
    int poll_sleep_ns;    // back-off between failed polls of the dataflow words
    int ll_use_sentinel;  // 1: one lane per warp polls first, then everybody loads; 0: everybody polls its own words
    unsigned long long* prof;  // optional [9][n_phases + 1], see b2l_debug_mega_profile globaltimer ns of the LAST step (CTA 0 / CTA G-1: phase end, wait end)
    // tensor parallel (the MEGA_TP build of the kernel; tp == 1 otherwise): every rank runs the kernel on its shard. The
    // row-parallel phases (O-proj, down) store their partial sums as {value, seq} words straight into EVERY rank's slab over
    // NVLink; the CTA that owns a row then polls the tp copies in its own memory, adds them in rank order to the residual
    // and publishes the row locally. The per-CTA argmax keys go to every rank the same way.
    int tp, tp_rank, vocab_base, Hpad;
    unsigned long long* tp_slab[8];   // rank p's slab region as mapped here: [2 (O-proj, down)][tp (source rank)][Hpad] words
    unsigned long long* tp_keys[8];   // rank p's argmax-key array as mapped here: [tp (source rank)][2 * gridDim] words
};


}  // namespace b2l

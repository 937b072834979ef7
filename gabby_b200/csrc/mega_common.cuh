// mega_common.cuh -- declarations shared by the two builds of the decode megakernel (mega_decode.cuh is compiled twice:
// MEGA_TP 0 = one GPU, MEGA_TP 1 = one tensor-parallel rank) and by the host side (engine.cu).
#pragma once
#include <stdint.h>

namespace b2l {

constexpr int kMegaConsumerWarps = 8;
constexpr int kMegaConsumerThreads = kMegaConsumerWarps * 32;
constexpr int kMegaThreads = kMegaConsumerThreads + 128;  // + the producer's warpgroup (one working lane; a whole warpgroup so
                                                             // that setmaxnreg can hand its registers to the consumers)
constexpr int kMegaStageBytes = 16 * 1024;
constexpr int kMegaMaxStages = 12;
constexpr int kMegaGroupRows = 16;   // rows of one MMA tile (mma.sync.m16n8k16: the weights are the A operand)
constexpr int kMegaBatchGroups = 8;  // row groups a CTA finishes per CTA barrier (128 rows: one reducing thread per row)
constexpr int kMegaProfRows = 16 + 2 * 160 + 64;  // debug timeline: 16 summary rows, then per CTA: input-ready and phase-end times, then per warp of CTA 0: input ready / rows done

enum MegaPhaseType { PH_QKV = 0, PH_ATTN = 1, PH_OPROJ = 2, PH_GATEUP = 3, PH_DOWN = 4, PH_LMHEAD = 5 };

struct MegaPhase {
    int type, layer;
    const uint16_t* W;       // [N][K] bf16 (PH_ATTN: unused)
    const uint16_t* norm_w;  // fused RMSNorm weight or null
    uint16_t* kv_pool;       // PH_ATTN: this layer's KV pool
    int N, K;
    float inv_k;             // 1 / K (RMSNorm mean)
    int reserved;
};

// rows [r0, r1) of an N-row matrix owned by CTA `c` of `G` (unit = 2 rows for SwiGLU pairs)
__host__ __device__ __forceinline__ void mega_row_range(int N, int unit, int c, int G, int& r0, int& r1) {
    const long long units = N / unit;
    r0 = static_cast<int>(units * c / G) * unit;
    r1 = static_cast<int>(units * (c + 1) / G) * unit;
}
// K window of a ring stage (log2 elements): 512 for the wide matrices; narrower for small K (the K shards of a tensor-parallel
// rank: 512, 1024, 1792 ...) so that every group still has >= 4 windows, one for each warp pair. K % 128 == 0 is required.
__host__ __device__ __forceinline__ int mega_ks_shift(int K) {
    if ((K & 511) == 0 && K >= 2048) return 9;
    if ((K & 255) == 0 && K >= 1024) return 8;
    return 7;
}

struct MegaArgs {
    const MegaPhase* phases;   // W = the TILED image of the matrix (mega_tile_kernel)
    int n_phases;
    int n_stages;
    // model
    const uint16_t* embed;
    const float* rope;
    int H, V, nh, nkv, hd, I;
    float eps, attn_scale;
    float* logits;   // [V] fp32 (parity tap and b2l_get_logits)
    // paged KV
    const int32_t* block_table;
    int page_size, page_shift, kvd;   // page_size is a power of two (2^page_shift)
    int nsplit_max;
    // token loop
    int32_t token0, pos0;  // arg_io != 0: first token / position travel in this struct (constant memory) instead of *token / *position
    int arg_io;
    int32_t* token;      // in: first token; out: last argmax
    int32_t* position;   // in: first position; out: advanced
    int32_t* out_ids;    // [n_steps]
    int n_steps;
    int* abort_flag;     // mapped pinned host memory: [0] abort code, [1 + cta] progress marker of each CTA (debug)
    int debug_nostream;  // 1: copy 16 bytes per chunk instead of the weights (timing experiments only; wrong results)
    int debug_progress;  // 1: CTAs record step*100000 + phase*100 + stage-of-phase
    int producer_sleep_ns;
    int attn_tps;       // context tokens per attention split (work item)
    int attn_max_splits;   // cap on context splits per head (every batch of 4 splits costs the O projection's input merge a round trip)
    int attn_qhead_tokens; // per-query-head items (instead of per-kv-head items) while an item stays within this many tokens
    int l2_ahead;       // chunks the producer prefetches into L2 beyond the ring while the ring is full (0 = off)
    // dataflow mode (kernel template LL): activations travel as 8-byte {fp32 bits, sequence number} words and every
    // reader polls for the sequence number of the phase that produces its input -- no grid barrier anywhere
    unsigned long long *ll_h, *ll_qkv, *ll_act, *ll_pacc, *ll_pml, *ll_keys;
    uint32_t seq_base;  // sequence numbers used by earlier launches
    int poll_sleep_ns;    // back-off between failed polls of the dataflow words
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

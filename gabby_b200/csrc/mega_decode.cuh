// mega_decode.cuh -- persistent batch-1 decode megakernel: ONE cooperative launch runs n_steps
// whole tokens. The B200-native answer to "a 0.4 ms token is 100 dependent tiny ops":
//
//   * one CTA per SM (148), 8 consumer warps + 1 producer warp
//   * the producer streams this CTA's share of EVERY weight matrix, in model order, through a ring of
//     16 KB shared-memory stages with TMA bulk copies (cp.async.bulk -> UBLKCP) that complete on
//     mbarriers. Weights do not depend on activations, so the stream never stops: it runs ahead
//     across op, layer and even token boundaries, bounded only by the ring.
//   * the weights are kept in HBM in a megakernel-specific TILED image (mega_tile_kernel in engine.cu):
//     the rows a CTA owns, in groups of <= 16 rows, each group cut into 512-element K windows; a window
//     of a group is one contiguous chunk of HBM = one bulk copy = one ring stage, and inside it the
//     16-byte pieces are ordered so that `ldmatrix` reads them conflict-free as the A operand of
//     mma.sync.m16n8k16. No padding: the image is a permutation of the row-major matrix.
//   * consumers: every warp owns 1/8 of each K window. The op's input vector is carried as THREE
//     bf16 vectors (hi + mid + lo = the fp32 value to 24 bits) that sit in columns 0..2 of the B
//     operand, so one MMA multiplies 16 weight rows x 16 K elements by the exact fp32 activation:
//     2-3 issue slots per 256 weights instead of 13 with FFMA2 (unpack + multiply), which is what
//     kept the first version of this kernel issue- and latency-bound in its row loop.
//   * attention (split-K over the paged bf16 cache, RoPE and K/V append fused in), SwiGLU,
//     residual adds, both RMSNorms, the lm_head argmax and the token feedback are all inside.
//
// Hand-off between phases ("dataflow"): no grid barrier. Every value that crosses CTAs is an 8-byte
//   word {fp32 bits, sequence number}; the sequence number names the phase instance that produced
//   it (seq_base + step * n_phases + phase + 1). A reader polls the words it needs until they carry
//   the number of the producing phase (always the phase before its own). The flag travels with the
//   data, so there is no fence and no counter, and a warp starts its MMAs as soon as ITS OWN slice
//   of the input exists (a warp only needs the K elements it multiplies: no shared-memory fan-out
//   and no CTA barrier between the arrival of the input and the first MMA).
//   Why nothing is overwritten too early (write-after-read): every phase reads the COMPLETE output
//   of the phase before it, and a CTA publishes outputs of a phase only after a CTA barrier that all
//   of its warps reach after loading their inputs. So when some CTA writes an output of phase p+1,
//   it has already seen all of phase p's outputs, hence every CTA has finished computing phase p,
//   hence every CTA has long finished LOADING phase p's inputs. A buffer is therefore safe to rewrite
//   two phases after it was written, and no buffer here is rewritten sooner: h is written by O-proj
//   and down (read by gate/up resp. the next QKV / lm_head), qkv by QKV (read by attention), the
//   partials by attention (read by O-proj), act by gate/up (read by down), the per-CTA argmax keys by
//   lm_head (read at the next token's first phase). Residual read-modify-writes of h touch only rows
//   the same CTA owns in both O-proj and down.
//   The only plain (non-word) data that crosses CTAs is the new token's K/V cache line, first read by the
//   next token: every CTA fences once per token before it publishes its argmax key, readers fence after the keys.
//   Sequence numbers are 32 bits and never reset (host keeps seq_base across launches); buffers
//   start at 0, which is never produced.
//
// HBM sees one sequential read of the model per token; everything else lives in L2 / smem.
// Math is that of decode_kernels.cuh (the multi-kernel path) up to fp32 summation order; RMSNorm's
// 1/rms is applied to the finished dot product instead of to the input (y = inv * W (g . x)).
// This file is compiled TWICE by engine.cu (no include guard): MEGA_TP 0 in namespace b2l::mega1 is the
// single-GPU kernel, MEGA_TP 1 in namespace b2l::megatp the kernel of a tensor-parallel rank.
#include "common.cuh"
#include "decode_kernels.cuh"
#include "mega_common.cuh"
#include "ptx_helpers.cuh"
#if !defined(MEGA_TP) || !defined(MEGA_NS)
#error "define MEGA_TP (0 or 1) and MEGA_NS before including mega_decode.cuh"
#endif

namespace b2l {
namespace MEGA_NS {
static_assert(kPtxConsumerThreads == kMegaConsumerThreads, "consumer_bar() counts the megakernel's consumer threads");

// The launch arguments live in __constant__ memory: the device functions below read them as
// constant-bank operands (no reloads after inline-asm memory clobbers, no generic loads through a
// pointer to the parameter space). One megakernel launch per device at a time (host side locks).
__constant__ MegaArgs c_mega;

// Loads that must be ISSUED where they are written (ahead of a wait they are meant to overlap): __ldcg / __ldg are
// non-volatile asm without a memory clobber, and the compiler sinks them to their first use -- i.e. to AFTER the wait.
__device__ __forceinline__ uint4 ld_cg_early(const void* p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float ld_cg_early_f32(const void* p) {
    float r;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ int ld_nc_early_s32(const void* p) {
    int r;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t ld_nc_early_u32(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}

template <int V> struct IntTag { static constexpr int value = V; };

// ---- dataflow words: {value, seq} in one 8-byte store; a 16-byte load brings two adjacent words ----
__device__ __forceinline__ void ll_st(unsigned long long* p, float v, uint32_t seq) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"((static_cast<unsigned long long>(seq) << 32) | __float_as_uint(v)) : "memory");
}
__device__ __forceinline__ uint4 ll_ld2(const unsigned long long* p) {  // .x/.z values, .y/.w sequence numbers
    uint4 w;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(p) : "memory");
    return w;
}
#if MEGA_TP
// words written by another GPU over NVLink: system scope
__device__ __forceinline__ void ll_st_sys(unsigned long long* p, float v, uint32_t seq) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"((static_cast<unsigned long long>(seq) << 32) | __float_as_uint(v)) : "memory");
}
__device__ __forceinline__ uint4 ll_ld2_sys(const unsigned long long* p) {
    uint4 w;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ uint2 ll_ld1_sys(const unsigned long long* p) {   // .x value, .y sequence number
    uint2 w;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(p) : "memory");
    return w;
}
#endif

// N x 8 consecutive words -> floats; retried until every word carries `seq`
template <int N>
__device__ __forceinline__ void ll_ld8n(const unsigned long long* const (&p)[N], uint32_t seq, float* out, int* abort_flag, int code) {
    unsigned spins = 0;
    for (;;) {
        uint4 w[N][4];
#pragma unroll
        for (int i = 0; i < N; i++) {
#pragma unroll
            for (int j = 0; j < 4; j++) w[i][j] = ll_ld2(p[i] + 2 * j);
        }
        bool ok = true;
#pragma unroll
        for (int i = 0; i < N; i++) {
#pragma unroll
            for (int j = 0; j < 4; j++) ok = ok && w[i][j].y == seq && w[i][j].w == seq;
        }
        if (ok) {
#pragma unroll
            for (int i = 0; i < N; i++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    out[i * 8 + 2 * j] = __uint_as_float(w[i][j].x);
                    out[i * 8 + 2 * j + 1] = __uint_as_float(w[i][j].z);
                }
            }
            return;
        }
        __nanosleep(c_mega.poll_sleep_ns);
        if (++spins > (1u << 22)) mega_die(abort_flag, code);
    }
}

// ---- shared memory map: 32-bit shared-space addresses, all kept in registers -----------------
struct MegaSmem {
    uint32_t ring;          // [n_stages][kMegaStageBytes]
    uint32_t full, empty;   // [n_stages] mbarriers each
    uint32_t zero16;        // 16 zero bytes: the `ldmatrix` row address of rows past the end of a group
    uint32_t rel;           // [n_stages] u32: arrivals of finished readers per slot (2 per use; see the consumer wait)
    uint32_t keys;          // [8] u64
    uint32_t red;           // [2][8] fp32: per-warp sums of squares (double buffered by phase parity)
    uint32_t part;          // [2][kMegaBatchGroups][8 warps][16 rows] fp32 partial row sums (double buffered by batch)
    uint32_t phases;        // [n_phases] MegaPhase copies (static for the whole launch), then this CTA's (r0, r1) per phase
    uint32_t xfrag;         // [K/16][12 lanes][8 bytes]: the input vector as B fragments (bf16 hi / mid / lo); the attention
                            // phase's scratch aliases it (the phases before and after end / start with a CTA barrier)
};

// a phase descriptor held in registers (read from the shared-memory copy)
struct PhaseRegs {
    int type, layer, N, K;
    int r0, r1;              // this CTA's row range of the phase (precomputed once per launch)
    const uint16_t* W;       // tiled image of the matrix
    float inv_k;             // 1 / K
    const uint16_t* norm_w;
    uint16_t* kv_pool;
};
static_assert(sizeof(MegaPhase) == 48, "MegaPhase is copied to shared memory as three 16-byte words");

__device__ __forceinline__ PhaseRegs mega_load_phase(uint32_t phases, int pi) {
    const uint4 a0 = lds128(phases + pi * 48), a1 = lds128(phases + pi * 48 + 16), a2 = lds128(phases + pi * 48 + 32);
    PhaseRegs r;
    r.type = static_cast<int>(a0.x);
    r.layer = static_cast<int>(a0.y);
    r.W = reinterpret_cast<const uint16_t*>(static_cast<unsigned long long>(a0.z) | (static_cast<unsigned long long>(a0.w) << 32));
    r.norm_w = reinterpret_cast<const uint16_t*>(static_cast<unsigned long long>(a1.x) | (static_cast<unsigned long long>(a1.y) << 32));
    r.kv_pool = reinterpret_cast<uint16_t*>(static_cast<unsigned long long>(a1.z) | (static_cast<unsigned long long>(a1.w) << 32));
    r.N = static_cast<int>(a2.x);
    r.K = static_cast<int>(a2.y);
    r.inv_k = __uint_as_float(a2.z);
    unsigned long long rr = lds64(phases + c_mega.n_phases * 48 + pi * 8);
    r.r0 = static_cast<int>(rr & 0xffffffffull);
    r.r1 = static_cast<int>(rr >> 32);
    return r;
}

// position in the ring: stage index and phase parity, advanced incrementally (no div/mod per chunk)
struct RingPos {
    int stage;
    uint32_t parity;
    uint32_t use;   // how many times the ring has wrapped = earlier uses of slot `stage`
    __device__ __forceinline__ void advance(int n_stages) {
        if (++stage == n_stages) {
            stage = 0;
            parity ^= 1;
            use++;
        }
    }
};

// Attention decomposition for this step. While (query heads x context splits) fits the grid, a work
// item is one QUERY head x one split (4x shorter items than per-kv-head, K/V re-read from L2);
// for longer contexts items are (kv head x split) and process the whole GQA group.
struct AttnPlan {
    int nsplit;
    bool per_q_head;
};
__device__ __forceinline__ AttnPlan mega_attn_plan(int nsplit_max, int ctx, int tps, int nh, int G) {
    const MegaArgs& a = c_mega;
    AttnPlan p;
    const int cap = min(nsplit_max, a.attn_max_splits);
    const int q_splits = min(cap, G / nh);          // splits per query head that still fit the grid
    if (q_splits >= 1 && (ctx + q_splits - 1) / q_splits <= a.attn_qhead_tokens) {
        p.per_q_head = true;                               // items of ONE query head
        p.nsplit = max(1, min(q_splits, (ctx + 31) / 32));
    } else {
        p.per_q_head = false;
        p.nsplit = max(1, min(cap, (ctx + tps - 1) / tps));
    }
    return p;
}

// ---- tensor-core helpers -----------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two fp32 -> one register of two bf16 (round to nearest even), x0 in the low half
__device__ __forceinline__ uint32_t cvt_bf16x2(float x0, float x1) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
    return r;
}
__device__ __forceinline__ void sts32u(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint2 lds64u2(uint32_t addr) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr) : "memory");
    return r;
}

// The pair (x0, x1) = input elements (k, k+1) of this lane -> bf16 hi / mid / lo, written where the MMA loop's B fragment
// load expects them. `frag` = address of the fragment register this lane feeds in column 0 (hi); columns 1 / 2 are 32 / 64
// bytes further (4 lanes x 8 bytes per column).
__device__ __forceinline__ void mega_put_frag(uint32_t frag, float x0, float x1) {
    const uint32_t hi = cvt_bf16x2(x0, x1);
    x0 -= bf16lo(hi); x1 -= bf16hi(hi);          // exact
    const uint32_t mid = cvt_bf16x2(x0, x1);
    x0 -= bf16lo(mid); x1 -= bf16hi(mid);        // exact
    const uint32_t lo = cvt_bf16x2(x0, x1);
    sts32u(frag, hi);
    sts32u(frag + 32, mid);
    sts32u(frag + 64, lo);
}

// everything the consumer threads carry across phases, in registers
struct ConsumerState {
    RingPos rp;
    unsigned long long best_key;  // running argmax of this step's logits
    int token, step;
    uint32_t batch;               // row batches reduced so far (selects the partial-sum buffer)
};

// token boundary: every CTA published the best (logit, index) key of its lm_head rows as two
// {32 bits, seq} words; all consumer threads of the CTA poll them and take the maximum
__device__ __forceinline__ int mega_poll_token(const MegaSmem& sm, uint32_t want, int tid) {
    const MegaArgs& a = c_mega;
    const int lane = tid & 31, w = tid >> 5;
    unsigned long long best = 0ull;
#if MEGA_TP
    const int n_keys = a.tp * static_cast<int>(gridDim.x);   // every CTA of every rank stores its key into this rank's array
#else
    const int n_keys = static_cast<int>(gridDim.x);
#endif
    for (int c = tid; c < n_keys; c += kMegaConsumerThreads) {
        unsigned spins = 0;
        for (;;) {
#if MEGA_TP
            const uint4 wk = ll_ld2_sys(a.ll_keys + 2 * c);
#else
            const uint4 wk = ll_ld2(a.ll_keys + 2 * c);
#endif
            if (wk.y == want && wk.w == want) {
                const unsigned long long k = (static_cast<unsigned long long>(wk.x) << 32) | wk.z;
                best = k > best ? k : best;
                break;
            }
            __nanosleep(32);
            if (++spins > (MEGA_TP ? (1u << 25) : (1u << 22))) mega_die(a.abort_flag, 130);   // ranks may start a launch milliseconds apart
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    consumer_bar();   // sm.keys may still be read by a slow warp of the previous lm_head epilogue
    if (lane == 0) sts64(sm.keys + w * 8, best);
    consumer_bar();
    best = lds64(sm.keys);
    for (int i = 1; i < kMegaConsumerWarps; i++) {
        const unsigned long long o = lds64(sm.keys + i * 8);
        best = o > best ? o : best;
    }
    return argmax_key_index(best);
}

// Address of the B-fragment register that holds input elements (k, k+1), k even, in column 0 (hi): fragments are indexed by
// k16-step (96 bytes: 12 lanes x 8 bytes); inside a step lane j = (k % 8) / 2 holds offsets 0..7 in b0 and 8..15 in b1.
__device__ __forceinline__ uint32_t mega_frag_addr(uint32_t xfrag, int k) {
    const int jj = (k & 15) >> 1;
    return xfrag + ((k >> 4) * 12 + (jj & 3)) * 8 + ((jj >> 2) & 1) * 4;
}

// NB segments (64 consecutive input elements each: one 16-byte load per lane) of this warp's share of the input vector: fetch
// (poll the dataflow words, or read the embedding row), apply the RMSNorm weight, accumulate the sum of squares, store the
// bf16 hi / mid / lo fragments. All NB loads are in flight together. `seg_k(q)` = first element of segment q.
template <int NB, typename SegK>
__device__ __forceinline__ void mega_fetch_segments(const PhaseRegs& ph, const unsigned long long* src, const uint16_t* esrc, bool from_embed,
                                                    int q0, int nseg, SegK seg_k, int lane, uint32_t want, uint32_t xfrag, float& ssq) {
    uint32_t nwv[NB];
    float x0[NB], x1[NB];
    int k[NB];
#pragma unroll
    for (int j = 0; j < NB; j++) k[j] = seg_k(min(q0 + j, nseg - 1)) + 2 * lane;   // segments past the end repeat the last one
    if (ph.norm_w) {   // RMSNorm weights of the slice: issued before the poll
#pragma unroll
        for (int j = 0; j < NB; j++) nwv[j] = ld_nc_early_u32(ph.norm_w + k[j]);
    }
    if (from_embed) {
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const uint32_t e = ld_nc_early_u32(esrc + k[j]);
            x0[j] = bf16lo(e);
            x1[j] = bf16hi(e);
        }
    } else {
        unsigned spins = 0;
        for (;;) {
            uint4 wd[NB];
#pragma unroll
            for (int j = 0; j < NB; j++) wd[j] = ll_ld2(src + k[j]);
            bool ok = true;
#pragma unroll
            for (int j = 0; j < NB; j++) ok = ok && wd[j].y == want && wd[j].w == want;
            if (ok) {
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    x0[j] = __uint_as_float(wd[j].x);
                    x1[j] = __uint_as_float(wd[j].z);
                }
                break;
            }
            __nanosleep(c_mega.poll_sleep_ns);
            if (++spins > (1u << 22)) mega_die(c_mega.abort_flag, 150 + ph.type);
        }
    }
#pragma unroll
    for (int j = 0; j < NB; j++) {
        if (q0 + j < nseg) {
            float v0 = x0[j], v1 = x1[j];
            if (ph.norm_w) {
                ssq = fmaf(v0, v0, fmaf(v1, v1, ssq));
                v0 *= bf16lo(nwv[j]);
                v1 *= bf16hi(nwv[j]);
            }
            mega_put_frag(mega_frag_addr(xfrag, k[j]), v0, v1);
        }
    }
}

// O projection input: NS segments of this warp's share of the attention output, merged from the split-K partials the
// attention items published (acc words per (head, split), max / sum words per (head, split)), NU splits per round trip.
// This code runs once per layer from a cold instruction cache: compile-time head_dim / group (no integer divisions), one
// max + NU exponentials per segment and round instead of an online rescale per split.
template <int HD, int GROUP, int NS, int NU, typename SegK>
__device__ __forceinline__ void mega_merge_segments(int q0, int nseg, SegK seg_k, int lane, int nsplit, uint32_t want, uint32_t xfrag) {
    const MegaArgs& a = c_mega;
    const int head_lane0 = HD >= 64 ? 0 : (lane & 16);   // a segment spans one head (two when head_dim is 32)
    float Mx[NS], L[NS], acc0[NS], acc1[NS];
    int kk[NS];
    const unsigned long long* pacc[NS];
    const unsigned long long* pml[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) {
        Mx[j] = -INFINITY; L[j] = 0.f; acc0[j] = 0.f; acc1[j] = 0.f;
        kk[j] = seg_k(min(q0 + j, nseg - 1)) + 2 * lane;
        const int head = kk[j] / HD, d = kk[j] % HD;
        const size_t rbase = (static_cast<size_t>(head / GROUP) * a.nsplit_max) * GROUP + head % GROUP;   // record of split 0
        pacc[j] = a.ll_pacc + rbase * HD + d;     // + split * GROUP * HD
        pml[j] = a.ll_pml + rbase * 2;            // + split * GROUP * 2
    }
    for (int s0 = 0; s0 < nsplit; s0 += NU) {
        uint4 wa[NS][NU], wm[NS];
        const int my_sp = min(s0 + (lane & (NU - 1)), nsplit - 1);
        unsigned spins = 0;
        for (;;) {
#pragma unroll
            for (int j = 0; j < NS; j++) {
#pragma unroll
                for (int u = 0; u < NU; u++) wa[j][u] = ll_ld2(pacc[j] + static_cast<size_t>(min(s0 + u, nsplit - 1)) * (GROUP * HD));   // clamped duplicates are masked below
                // (max, sum) of split s0 + u: fetched by lane u of the head's lanes, handed round with shuffles
                wm[j] = ll_ld2(pml[j] + static_cast<size_t>(my_sp) * (GROUP * 2));
            }
            bool ok = true;
#pragma unroll
            for (int j = 0; j < NS; j++) {
                ok = ok && wm[j].y == want && wm[j].w == want;
#pragma unroll
                for (int u = 0; u < NU; u++) ok = ok && wa[j][u].y == want && wa[j][u].w == want;
            }
            if (__all_sync(0xffffffffu, ok)) break;
            __nanosleep(c_mega.poll_sleep_ns);
            if (++spins > (1u << 22)) mega_die(a.abort_flag, 120);
        }
#pragma unroll
        for (int j = 0; j < NS; j++) {
            float mu[NU], lu[NU];
            float mn = Mx[j];
#pragma unroll
            for (int u = 0; u < NU; u++) {
                mu[u] = __shfl_sync(0xffffffffu, __uint_as_float(wm[j].x), head_lane0 + u);
                lu[u] = __shfl_sync(0xffffffffu, __uint_as_float(wm[j].z), head_lane0 + u);
                if (s0 + u >= nsplit) mu[u] = -INFINITY;   // clamped duplicate
                mn = fmaxf(mn, mu[u]);
            }
            if (mn == -INFINITY) continue;                  // nothing but empty splits so far
            const float c_old = __expf(Mx[j] - mn);          // exp(-inf) = 0 in the first round
            float l = L[j] * c_old, x0 = acc0[j] * c_old, x1 = acc1[j] * c_old;
#pragma unroll
            for (int u = 0; u < NU; u++) {
                const float e = __expf(mu[u] - mn);          // 0 for empty splits (mu = -inf)
                l = fmaf(lu[u], e, l);
                x0 = fmaf(__uint_as_float(wa[j][u].x), e, x0);
                x1 = fmaf(__uint_as_float(wa[j][u].z), e, x1);
            }
            L[j] = l; acc0[j] = x0; acc1[j] = x1; Mx[j] = mn;
        }
    }
#pragma unroll
    for (int j = 0; j < NS; j++) {
        if (q0 + j < nseg) {
            const float inv = __frcp_rn(L[j]);
            mega_put_frag(mega_frag_addr(xfrag, kk[j]), acc0[j] * inv, acc1[j] * inv);
        }
    }
}

// ---- one GEMV-type phase for one CTA ---------------------------------------------------------
// Geometry of a phase with K input elements: KS = 512 (256 when K is not a multiple of 512) elements per K window = ring
// stage, P = K / KS windows per row group. The eight warps form four PAIRS; pair j owns the windows p = j, j+4, ... of every
// group (four consecutive stages are in work at once, like four independent streams), and inside a window each warp of the
// pair multiplies one half (KS/2 elements = T k16-steps). A warp polls, converts and keeps exactly the input elements it
// multiplies; one wait and one arrive per warp per 8 KB of weights.
template <int HD, int GROUP>
__device__ __forceinline__ void mega_gemv_phase(const PhaseRegs& ph, const MegaSmem& sm, ConsumerState& st_ref, int pi, int pos, int tid) {
    const MegaArgs& a = c_mega;
    ConsumerState& st = st_ref;
    const int lane = tid & 31, w = tid >> 5;
    const int K = ph.K, type = ph.type;
    const int ks_shift = mega_ks_shift(K);
    const int P = K >> ks_shift;
    const bool prof = a.prof && st.step == a.n_steps - 1 && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
    unsigned long long* prof_col = a.prof + pi;
    const int prow = blockIdx.x == 0 ? 0 : 4, pstride = a.n_phases + 1;
    if (prof) prof_col[(prow + 0) * pstride] = globaltimer_ns();
    volatile int* progress = a.debug_progress && tid == 0 ? a.abort_flag + 1 + blockIdx.x : nullptr;
    if (progress) *progress = st.step * 100000 + pi * 100 + 1;

#ifdef MEGA_PROF_WARP
    long long wt[8] = {clock64(), 0, 0, 0, 0, 0, 0, 0};   // debug: per-warp cycle stamps of the profiled step (CTA 0)
#define MEGA_WT(i) wt[i] = clock64()
#define MEGA_WT_ONCE(i) if (wt[i] == 0) wt[i] = clock64()
#else
#define MEGA_WT(i)
#define MEGA_WT_ONCE(i)
#endif
    const int r0 = ph.r0, r1 = ph.r1, nrows = r1 - r0;
    const bool from_embed = (type == PH_QKV && ph.layer == 0);
    const uint32_t gp = a.seq_base + static_cast<uint32_t>(st.step * a.n_phases + pi) + 1u;  // this phase's sequence number
    const uint32_t want = gp - 1u;                                                           // inputs come from the phase before
    if (from_embed && st.step > 0) {
        st.token = mega_poll_token(sm, want, tid);
        if (blockIdx.x == 0 && tid == 0) a.out_ids[st.step - 1] = st.token;
        __threadfence();  // order this step's KV-cache reads after everything the previous step published
    }
    if (prof) prof_col[(prow + 1) * pstride] = globaltimer_ns();
    if (progress) *progress = st.step * 100000 + pi * 100 + 2;
    const int token = st.token;

    // ---- everything the row loop needs that does not depend on the input: computed (and the residual of the first batch
    // fetched) BEFORE the input poll, so that it overlaps the wait instead of sitting on the critical path after it ----
    const int pj = w >> 1, ph_half = w & 1;
    const bool resid_h = type == PH_DOWN || (type == PH_OPROJ && ph.layer != 0);
    const bool resid_e = type == PH_OPROJ && ph.layer == 0;
    const int n_stages = a.n_stages;
    int* const abort_flag = a.abort_flag;
    RingPos rp = st.rp;   // first stage of the current group
    const int T = 1 << (ks_shift - 5);   // k16-steps per warp per stage: 16, 8 or 4
    // ldmatrix.x4: lanes 8i..8i+7 give the row addresses of 8x8 matrix i; matrices (rows 0-7 | 8-15) x (k 0-7 | 8-15)
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8;
    const uint32_t a_piece0 = static_cast<uint32_t>(ph_half * T * 2 + (lane >> 4));
    float resid0 = 0.f;
    if (tid < min(16 * kMegaBatchGroups, nrows)) {
        if (resid_h) resid0 = ld_cg_early_f32(a.ll_h + r0 + tid);
        else if (resid_e) resid0 = bf16_bits_to_f32(a.embed[static_cast<size_t>(token) * a.H + r0 + tid]);
    }

    // ---- this warp's share of the input vector -> bf16 hi / mid / lo B fragments in shared memory ----
    // The share is a list of 64-element segments: window m of the pair (p = pj + 4m) x this warp's half x LPW segments.
    const int SL = 1 << (ks_shift - 1);                 // elements per warp per window
    const int lpw_shift = ks_shift - 7;                 // log2(segments per warp per window): 2 or 1
    const int nwin = pj < P ? (P - pj + 3) >> 2 : 0;
    const int nseg = nwin << lpw_shift;
    auto seg_k = [&](int q) { return ((pj + 4 * (q >> lpw_shift)) << ks_shift) + ph_half * SL + 64 * (q & ((1 << lpw_shift) - 1)); };
    float ssq = 0.f;
    if (type == PH_OPROJ) {
        // the attention phase left split-K partials (acc, max, sum) per (query head, split): merge them here. About twenty
        // 16-byte loads per lane per round trip: four segments x four splits, or fewer segments x more splits when the
        // warp's share of a narrow K (a tensor-parallel rank's K shard) is only one or two segments
        const int nsplit = mega_attn_plan(a.nsplit_max, pos + 1, a.attn_tps, a.nh, gridDim.x).nsplit;
        for (int q0 = 0; q0 < nseg;) {
            const int n = nseg - q0;
            if (n >= 3) { mega_merge_segments<HD, GROUP, 4, 4>(q0, nseg, seg_k, lane, nsplit, want, sm.xfrag); q0 += 4; }
            else if (n == 2) { mega_merge_segments<HD, GROUP, 2, 8>(q0, nseg, seg_k, lane, nsplit, want, sm.xfrag); q0 += 2; }
            else { mega_merge_segments<HD, GROUP, 1, 16>(q0, nseg, seg_k, lane, nsplit, want, sm.xfrag); q0 += 1; }
        }
    } else {
        const unsigned long long* src = type == PH_DOWN ? a.ll_act : a.ll_h;
        const uint16_t* esrc = a.embed + static_cast<size_t>(token) * a.H;
        for (int q0 = 0; q0 < nseg;) {   // as many segments per round trip as fit the registers, without issuing duplicate loads
            const int n = nseg - q0;
            if (n >= 9) { mega_fetch_segments<16>(ph, src, esrc, from_embed, q0, nseg, seg_k, lane, want, sm.xfrag, ssq); q0 += 16; }
            else if (n >= 5) { mega_fetch_segments<8>(ph, src, esrc, from_embed, q0, nseg, seg_k, lane, want, sm.xfrag, ssq); q0 += 8; }
            else if (n >= 3) { mega_fetch_segments<4>(ph, src, esrc, from_embed, q0, nseg, seg_k, lane, want, sm.xfrag, ssq); q0 += 4; }
            else if (n == 2) { mega_fetch_segments<2>(ph, src, esrc, from_embed, q0, nseg, seg_k, lane, want, sm.xfrag, ssq); q0 += 2; }
            else { mega_fetch_segments<1>(ph, src, esrc, from_embed, q0, nseg, seg_k, lane, want, sm.xfrag, ssq); q0 += 1; }
        }
        if (ph.norm_w) {
            ssq = warp_sum(ssq);
            if (lane == 0) sts32f(sm.red + ((pi & 1) * 8 + w) * 4, ssq);   // read after the first batch barrier below
        }
    }
    __syncwarp();   // the fragments are read by other lanes of this warp (never by other warps)
    MEGA_WT(1);
    if (prof) prof_col[(prow + 2) * pstride] = globaltimer_ns();
    const bool prof_all = a.prof && st.step == a.n_steps - 1 && tid == 0 && blockIdx.x < 160;
    if (prof_all) prof_col[(16 + blockIdx.x) * pstride] = globaltimer_ns();
    if (progress) *progress = st.step * 100000 + pi * 100 + 3;

    // ---- stream this CTA's rows: groups of <= 16 rows, P ring stages per group; this warp's pair takes every fourth stage ----
    unsigned long long best_key = st.best_key;
#ifdef MEGA_PROF_ROUNDS
    long long pr_wait = 0, pr_math = 0, pr_tail = 0, pr_n = 0;
#endif
    for (int b0 = 0; b0 < nrows; b0 += 16 * kMegaBatchGroups) {
        const int brows = min(16 * kMegaBatchGroups, nrows - b0);
        const int my_row = r0 + b0 + tid;
        const bool my_live = tid < brows;
        float resid = resid0;  // residual input of the row this thread will finish (first batch: fetched before the input poll)
        if (b0 > 0 && my_live) {
            if (resid_h) resid = ld_cg_early_f32(a.ll_h + my_row);
            else if (resid_e) resid = bf16_bits_to_f32(a.embed[static_cast<size_t>(token) * a.H + my_row]);
        }
        const uint32_t pbuf = sm.part + (st.batch & 1u) * (kMegaBatchGroups * kMegaConsumerWarps * 16 * 4);
        for (int g = 0; g * 16 < brows; g++) {
            const int r = min(16, brows - g * 16);
            const bool rowok = a_row < r;
            const uint32_t a_off = rowok ? (a_piece0 * r + a_row) * 16 : 0u;
            const uint32_t a_step = rowok ? static_cast<uint32_t>(r) * 32 : 0u;
            // eight independent accumulators: back-to-back MMAs into one accumulator would run at the tensor pipe's latency
            float c[8][4];
#pragma unroll
            for (int i = 0; i < 8; i++) { c[i][0] = 0.f; c[i][1] = 0.f; c[i][2] = 0.f; c[i][3] = 0.f; }
            for (int p = pj; p < P; p += 4) {
                // ring position of window p of this group
                RingPos sp = rp;
                sp.stage += p;
                while (sp.stage >= n_stages) { sp.stage -= n_stages; sp.parity ^= 1; sp.use++; }
#ifdef MEGA_PROF_ROUNDS
                const long long t0 = clock64();
#endif
                MEGA_WT_ONCE(7);
                {
                    // Successive uses of a ring slot belong to different pairs, so a pair can get here before the PREVIOUS use of
                    // the slot has even landed, and a parity wait cannot tell "one phase behind" from "done". The release counter
                    // of the previous use is the proof that it landed (its readers finished); only then is the parity wait
                    // unambiguous. (Usually satisfied on the first load.)
                    unsigned spins = 0;
                    for (;;) {
                        uint32_t done;
                        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(done) : "r"(sm.rel + sp.stage * 4) : "memory");
                        if (static_cast<int32_t>(done - 2u * sp.use) >= 0) break;
                        if (++spins > (1u << 28)) mega_die(abort_flag, 210 + type);
                    }
                }
                mbar_wait(sm.full + sp.stage * 8, sp.parity, abort_flag, 200 + type);
                MEGA_WT_ONCE(2);
#ifdef MEGA_PROF_ROUNDS
                const long long t1 = clock64();
#endif
                const uint32_t abase = rowok ? sm.ring + static_cast<uint32_t>(sp.stage) * kMegaStageBytes + a_off : sm.zero16;
                const uint32_t bf = sm.xfrag + (((p << (ks_shift - 4)) + ph_half * T) * 12 + lane) * 8;
                if (T >= 8) {
                    for (int t0 = 0; t0 < T; t0 += 8) {   // eight independent MMAs per round of loads
                        uint32_t A[8][4];
                        uint2 B[8];
#pragma unroll
                        for (int t = 0; t < 8; t++) ldmatrix_x4(A[t], abase + (t0 + t) * a_step);
#pragma unroll
                        for (int t = 0; t < 8; t++) {
                            B[t] = make_uint2(0u, 0u);
                            if (lane < 12) B[t] = lds64u2(bf + (t0 + t) * 96);
                        }
#pragma unroll
                        for (int t = 0; t < 8; t++) mma_bf16_16816(c[t], A[t], B[t].x, B[t].y);
                    }
                } else {   // T = 4: the 128-element windows of a narrow K
                    uint32_t A[4][4];
                    uint2 B[4];
#pragma unroll
                    for (int t = 0; t < 4; t++) ldmatrix_x4(A[t], abase + t * a_step);
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        B[t] = make_uint2(0u, 0u);
                        if (lane < 12) B[t] = lds64u2(bf + t * 96);
                    }
#pragma unroll
                    for (int t = 0; t < 4; t++) mma_bf16_16816(c[t], A[t], B[t].x, B[t].y);
                }
                __syncwarp();
                if (lane == 0) {
                    // publish "this use was read" BEFORE arriving: the slot can only be refilled (and its next use released)
                    // after both arrivals of this use, so the counter never runs backwards
                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(sm.rel + sp.stage * 4) : "memory");
                    mbar_arrive(sm.empty + sp.stage * 8);   // two arrivals (the pair) free the stage
                }
#ifdef MEGA_PROF_ROUNDS
                pr_wait += t1 - t0; pr_math += clock64() - t1; pr_n++;
#endif
            }
            MEGA_WT_ONCE(3);
            // next group: P stages further
            rp.stage += P;
            while (rp.stage >= n_stages) { rp.stage -= n_stages; rp.parity ^= 1; rp.use++; }
            // C fragment: lane 4i + q holds columns 2q, 2q+1 of rows i (c[0], c[1]) and i + 8 (c[2], c[3]); columns 0, 1, 2 are
            // the hi, mid and lo products of the row
            float cs[4];
#pragma unroll
            for (int e = 0; e < 4; e++) cs[e] = ((c[0][e] + c[1][e]) + (c[2][e] + c[3][e])) + ((c[4][e] + c[5][e]) + (c[6][e] + c[7][e]));
            const int q = lane & 3;
            float v = q == 0 ? cs[0] + cs[1] : q == 1 ? cs[0] : 0.f;
            float v8 = q == 0 ? cs[2] + cs[3] : q == 1 ? cs[2] : 0.f;
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v8 += __shfl_xor_sync(0xffffffffu, v8, 1);
            if (q == 0) {
                const uint32_t dst = pbuf + ((g * kMegaConsumerWarps + w) * 16 + (lane >> 2)) * 4;
                sts32f(dst, v);
                sts32f(dst + 32, v8);
            }
        }
#ifdef MEGA_PROF_ROUNDS
        const long long t2 = clock64();
#endif
        if (b0 == 0) { MEGA_WT(4); }
        consumer_bar();   // every warp's K slice of every row of the batch is in shared memory
        st.batch++;
        if (b0 == 0) { MEGA_WT(5); }
        if (tid < 16 * kMegaBatchGroups) {   // warps 0..3 finish the rows; warps 4..7 go on to the next batch / phase
            float s = 0.f;
#pragma unroll
            for (int ww = 0; ww < kMegaConsumerWarps; ww++) s += lds32f(pbuf + (((tid >> 4) * kMegaConsumerWarps + ww) * 16 + (tid & 15)) * 4);
            if (ph.norm_w) {
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < kMegaConsumerWarps; i++) ss += lds32f(sm.red + ((pi & 1) * 8 + i) * 4);
                s *= rsqrtf(ss * ph.inv_k + a.eps);
            }
            if (type == PH_GATEUP) {
                const float up = __shfl_down_sync(0xffffffffu, s, 1);  // rows are (gate, up) pairs
                s = __fdividef(s, 1.0f + __expf(-s)) * up;
            }
            if (my_live) {
                if (type == PH_QKV) {
                    ll_st(a.ll_qkv + my_row, s, gp);
                } else if (type == PH_GATEUP) {
                    if ((tid & 1) == 0) ll_st(a.ll_act + (my_row >> 1), s, gp);
                } else if (type == PH_LMHEAD) {
                    a.logits[my_row] = s;
                    const unsigned long long key = argmax_key(s, (MEGA_TP ? a.vocab_base : 0) + my_row);   // global vocabulary index
                    best_key = key > best_key ? key : best_key;
                } else {
#if MEGA_TP
                    // Row-parallel phase of a tensor-parallel rank: the partial sum goes to every rank's slab (own included) over
                    // NVLink as one 8-byte word; then this thread collects the tp partial sums of the row from its own slab, adds
                    // them to the residual in rank order -- every rank computes bit-identical h -- and publishes the row locally.
                    const size_t slab_off = static_cast<size_t>(type == PH_DOWN ? 1 : 0) * a.tp * a.Hpad + my_row;
                    for (int p = 0; p < a.tp; p++) ll_st_sys(a.tp_slab[p] + slab_off + static_cast<size_t>(a.tp_rank) * a.Hpad, s, gp);
                    const unsigned long long* mine = a.tp_slab[a.tp_rank] + slab_off;
                    float sum = resid;
                    unsigned spins = 0;
                    for (;;) {
                        uint2 wv[8];
#pragma unroll
                        for (int p = 0; p < 8; p++)
                            if (p < a.tp) wv[p] = ll_ld1_sys(mine + static_cast<size_t>(p) * a.Hpad);
                        bool ok = true;
#pragma unroll
                        for (int p = 0; p < 8; p++) ok = ok && (p >= a.tp || wv[p].y == gp);
                        if (ok) {
#pragma unroll
                            for (int p = 0; p < 8; p++)
                                if (p < a.tp) sum += __uint_as_float(wv[p].x);
                            break;
                        }
                        if (++spins > (1u << 25)) mega_die(a.abort_flag, 230 + type);
                    }
                    ll_st(a.ll_h + my_row, sum, gp);
#else
                    ll_st(a.ll_h + my_row, resid + s, gp);  // O-proj / down: residual add
#endif
                }
            }
        }
#ifdef MEGA_PROF_ROUNDS
        pr_tail += clock64() - t2;
#endif
    }
#ifdef MEGA_PROF_ROUNDS
    if (prof && blockIdx.x == 0) {
        prof_col[9 * pstride] = pr_wait;
        prof_col[10 * pstride] = pr_math;
        prof_col[11 * pstride] = pr_tail;
        prof_col[12 * pstride] = pr_n;
    }
#endif
#ifdef MEGA_PROF_WARP
    if (a.prof && st.step == a.n_steps - 1 && lane == 0 && blockIdx.x == 0) {
        wt[6] = clock64();
#pragma unroll
        for (int i = 0; i < 8; i++) prof_col[(16 + 320 + i * 8 + w) * pstride] = static_cast<unsigned long long>(wt[i]);
    }
#endif
#undef MEGA_WT
#undef MEGA_WT_ONCE
    st.rp = rp;
    st.best_key = best_key;
    if (progress) *progress = st.step * 100000 + pi * 100 + 4;
    if (prof) prof_col[(prow + 3) * pstride] = globaltimer_ns();
    if (prof_all) prof_col[(16 + 160 + blockIdx.x) * pstride] = globaltimer_ns();
}

// ---- attention work item: (kv head, split); partials are merged by the O-proj phase's input load -----
template <int HD, int GROUP>
__device__ __forceinline__ void mega_attn_item_body(uint16_t* kv_pool, uint32_t scratch, int kvh, int split, int nsplit,
                                            int pos, int tid, unsigned long long* pcol, int group_total, int g0, uint32_t gp) {
    // processes query heads kvh*group_total + g0 .. + GROUP (GROUP == group_total, or 1 in per-query-head mode)
    const MegaArgs& a = c_mega;
    long long ak0 = 0, ak1 = 0, ak2 = 0, ak3 = 0, ak4 = 0, ak5 = 0;
    if (pcol) ak0 = clock64();
    constexpr int LPT = HD / 8, TPW = 32 / LPT, HALF = HD / 2;
    const int lane = tid & 31, w = tid >> 5, sub = lane / LPT, sl = lane % LPT;
    const KvLayout kv{kv_pool, a.page_size, a.kvd};
    const int ctx = pos + 1;
    const int chunk = (ctx + nsplit - 1) / nsplit;
    const int j0 = split * chunk, j1 = min(ctx, j0 + chunk);
    const float* cs = a.rope + static_cast<size_t>(pos) * HD;  // [HALF][2]
    const int qd = a.nh * HD;

    // rotate-half RoPE of one 8-wide slice of a head: xx[0..7] = elements j0r.. of the lower half, xx[8..15] = of the upper half
    const uint32_t want = gp - 1u;
    const int d0 = sl * 8, j0r = d0 < HALF ? d0 : d0 - HALF;  // the slice lies in one half (HALF % 8 == 0)
    auto rope_apply = [&](const float* xx, float* out) {
        const float* x0 = xx;
        const float* x1 = xx + 8;
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const float4 c2 = __ldg(reinterpret_cast<const float4*>(cs + 2 * (j0r + i)));  // (c, s, c', s')
            out[i] = d0 < HALF ? x0[i] * c2.x - x1[i] * c2.y : x1[i] * c2.x + x0[i] * c2.y;
            out[i + 1] = d0 < HALF ? x0[i + 1] * c2.z - x1[i + 1] * c2.w : x1[i + 1] * c2.z + x0[i + 1] * c2.w;
        }
    };
    auto rope_slice = [&](int head, float* out) {   // head = element offset in the fused qkv row
        float xx[16];
        const unsigned long long* const pp[2] = {a.ll_qkv + head + j0r, a.ll_qkv + head + j0r + HALF};
        ll_ld8n<2>(pp, want, xx, a.abort_flag, 180);
        rope_apply(xx, out);
    };
    // token slots per lane group in flight: all K/V loads of a block are issued before any math. Six at head_dim 64 for the
    // per-query-head items (a block = 192 tokens: the items of contexts up to 768 are ONE block, no second round trip to L2)
    constexpr int U = HD == 64 && GROUP == 1 ? 6 : 4;
    constexpr int STEP = U * kMegaConsumerWarps * TPW;
    uint4 kwA[U], vwA[U];
    // cached tokens do not depend on this step's projections: their K/V loads go out BEFORE the wait for q
    auto load_block = [&](uint4 (&kw)[U], uint4 (&vw)[U], int jb) {
        int page[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = jb + (u * kMegaConsumerWarps + w) * TPW + sub;
            page[u] = j < j1 ? ld_nc_early_s32(a.block_table + (j >> a.page_shift)) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = jb + (u * kMegaConsumerWarps + w) * TPW + sub;
            kw[u] = make_uint4(0, 0, 0, 0);
            vw[u] = make_uint4(0, 0, 0, 0);
            if (j < j1 && j != pos) {
                const int off = j & (a.page_size - 1);
                kw[u] = ld_cg_early(kv.at(page[u], 0, off) + kvh * HD + sl * 8);
                vw[u] = ld_cg_early(kv.at(page[u], 1, off) + kvh * HD + sl * 8);
            }
        }
    };
    load_block(kwA, vwA, j0);
    float q[GROUP][8];
    // The item whose range ends at the token being decoded also needs that token's k and v (this step's projection): their
    // words are polled together with the first query head (one round trip instead of three on the phase's critical path).
    const bool has_new = j1 == ctx;
    float k_new[8], v_new[8];
    {
        const int qh = (kvh * group_total + g0) * HD, kh = qd + kvh * HD;
        if (has_new) {
            float xx[40];
            const unsigned long long* const pp[5] = {a.ll_qkv + qh + j0r, a.ll_qkv + qh + j0r + HALF, a.ll_qkv + kh + j0r, a.ll_qkv + kh + j0r + HALF,
                                                     a.ll_qkv + qd + a.kvd + kvh * HD + sl * 8};
            ll_ld8n<5>(pp, want, xx, a.abort_flag, 181);
            rope_apply(xx, q[0]);
            rope_apply(xx + 16, k_new);
#pragma unroll
            for (int i = 0; i < 8; i++) v_new[i] = xx[32 + i];
        } else {
            rope_slice(qh, q[0]);
#pragma unroll
            for (int i = 0; i < 8; i++) { k_new[i] = 0.f; v_new[i] = 0.f; }
        }
    }
#pragma unroll
    for (int g = 1; g < GROUP; g++) rope_slice((kvh * group_total + g0 + g) * HD, q[g]);
#pragma unroll
    for (int g = 0; g < GROUP; g++) {
#pragma unroll
        for (int i = 0; i < 8; i++) q[g][i] *= a.attn_scale;
    }
    if (pcol) ak1 = clock64();
    float m[GROUP], l[GROUP], acc[GROUP][8];
#pragma unroll
    for (int g = 0; g < GROUP; g++) {
        m[g] = -INFINITY;
        l[g] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) acc[g][i] = 0.f;
    }
    auto block = [&](uint4 (&kw)[U], uint4 (&vw)[U], int jb) {
        bool valid[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = jb + (u * kMegaConsumerWarps + w) * TPW + sub;
            valid[u] = j < j1;
            if (valid[u] && j == pos) {
                // the token being decoded: K/V come from this step's projection; append them (bf16)
                kw[u] = make_uint4(pack_bf16x2(k_new[0], k_new[1]), pack_bf16x2(k_new[2], k_new[3]), pack_bf16x2(k_new[4], k_new[5]), pack_bf16x2(k_new[6], k_new[7]));
                vw[u] = make_uint4(pack_bf16x2(v_new[0], v_new[1]), pack_bf16x2(v_new[2], v_new[3]), pack_bf16x2(v_new[4], v_new[5]), pack_bf16x2(v_new[6], v_new[7]));
                if (g0 == 0) {  // one writer per kv head. No fence here: the line is first read by the NEXT token, and this
                                // CTA fences once per token before it publishes its argmax key (readers fence after the keys)
                    const int page = __ldg(a.block_table + (j >> a.page_shift)), off = j & (a.page_size - 1);
                    *reinterpret_cast<uint4*>(kv.at(page, 0, off) + kvh * HD + sl * 8) = kw[u];
                    *reinterpret_cast<uint4*>(kv.at(page, 1, off) + kvh * HD + sl * 8) = vw[u];
                }
            }
        }
        // all U scores first (independent dot products and shuffle trees), then ONE online-softmax update per block:
        // the per-token update was a serial chain of exp / max / rescale per token
        float sc[U][GROUP];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const float kf[8] = {bf16lo(kw[u].x), bf16hi(kw[u].x), bf16lo(kw[u].y), bf16hi(kw[u].y), bf16lo(kw[u].z), bf16hi(kw[u].z), bf16lo(kw[u].w), bf16hi(kw[u].w)};
#pragma unroll
            for (int g = 0; g < GROUP; g++) {
                float d = 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) d = fmaf(q[g][i], kf[i], d);
                sc[u][g] = d;
            }
        }
#pragma unroll
        for (int o = LPT / 2; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < U; u++) {
#pragma unroll
                for (int g = 0; g < GROUP; g++) sc[u][g] += __shfl_xor_sync(0xffffffffu, sc[u][g], o);
            }
        }
#pragma unroll
        for (int g = 0; g < GROUP; g++) {
            float mn = m[g];
#pragma unroll
            for (int u = 0; u < U; u++) mn = valid[u] ? fmaxf(mn, sc[u][g]) : mn;
            if (mn == -INFINITY) continue;   // nothing valid yet in this lane group
            const float corr = __expf(m[g] - mn);
            float pu[U], ps = 0.f;
#pragma unroll
            for (int u = 0; u < U; u++) {
                pu[u] = valid[u] ? __expf(sc[u][g] - mn) : 0.f;
                ps += pu[u];
            }
            l[g] = l[g] * corr + ps;
#pragma unroll
            for (int i = 0; i < 8; i++) acc[g][i] *= corr;
#pragma unroll
            for (int u = 0; u < U; u++) {
                const float vf[8] = {bf16lo(vw[u].x), bf16hi(vw[u].x), bf16lo(vw[u].y), bf16hi(vw[u].y), bf16lo(vw[u].z), bf16hi(vw[u].z), bf16lo(vw[u].w), bf16hi(vw[u].w)};
#pragma unroll
                for (int i = 0; i < 8; i++) acc[g][i] = fmaf(pu[u], vf[i], acc[g][i]);
            }
            m[g] = mn;
        }
    };
    // Two blocks of K/V in flight (block i+1 loads while block i is computed) for per-query-head items at head_dim 128, where a
    // block is 64 tokens and long contexts make items of many blocks (8B TP=8 rank shapes at context 4096: attention phase
    // 19.9 -> 11.1 us). At head_dim 64 a block is 128 tokens, short-context items are one or two blocks, and the extra code
    // cost the 1B headline 1.6 %.
    if constexpr (GROUP == 1 && HD >= 128) {
        uint4 kwB[U], vwB[U];
        for (int jb = j0; jb < j1; jb += 2 * STEP) {
            const bool more = jb + STEP < j1;
            if (more) load_block(kwB, vwB, jb + STEP);
            block(kwA, vwA, jb);
            if (more) {
                if (jb + 2 * STEP < j1) load_block(kwA, vwA, jb + 2 * STEP);
                block(kwB, vwB, jb + STEP);
            }
        }
    } else {   // (the whole-group item has no registers to spare for a second block)
        for (int jb = j0; jb < j1; jb += STEP) {
            if (jb != j0) load_block(kwA, vwA, jb);
            block(kwA, vwA, jb);
        }
    }
    if (pcol) ak2 = clock64();
    // merge the TPW token sub-slots of the warp with shuffles
#pragma unroll
    for (int o = LPT; o < 32; o <<= 1) {
#pragma unroll
        for (int g = 0; g < GROUP; g++) {
            const float mo = __shfl_xor_sync(0xffffffffu, m[g], o), lo = __shfl_xor_sync(0xffffffffu, l[g], o);
            const float mn = fmaxf(m[g], mo);
            const float ca = mn == -INFINITY ? 0.f : __expf(m[g] - mn), cb = mn == -INFINITY ? 0.f : __expf(mo - mn);
            l[g] = l[g] * ca + lo * cb;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float ao = __shfl_xor_sync(0xffffffffu, acc[g][i], o);
                acc[g][i] = acc[g][i] * ca + ao * cb;
            }
            m[g] = mn;
        }
    }
    if (pcol) ak3 = clock64();
    // per-warp results -> smem: acc [w][g][HD], then (m, l) [w][g][2]
    const uint32_t s_acc = scratch, s_ml = scratch + kMegaConsumerWarps * GROUP * HD * 4;
    if (sub == 0) {
#pragma unroll
        for (int g = 0; g < GROUP; g++) {
            const uint32_t dst = s_acc + ((w * GROUP + g) * HD + sl * 8) * 4;
            sts128f(dst, make_float4(acc[g][0], acc[g][1], acc[g][2], acc[g][3]));
            sts128f(dst + 16, make_float4(acc[g][4], acc[g][5], acc[g][6], acc[g][7]));
            if (sl == 0) {
                sts32f(s_ml + (w * GROUP + g) * 8, m[g]);
                sts32f(s_ml + (w * GROUP + g) * 8 + 4, l[g]);
            }
        }
    }
    consumer_bar();
    if (pcol) ak4 = clock64();
    const size_t pbase = static_cast<size_t>(kvh) * a.nsplit_max + split;
    for (int e = tid; e < GROUP * HD; e += kMegaConsumerThreads) {
        const int g = e / HD, d = e % HD;
        float Mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < kMegaConsumerWarps; t++) Mx = fmaxf(Mx, lds32f(s_ml + (t * GROUP + g) * 8));
        float L = 0.f, A = 0.f;
        if (Mx > -INFINITY) {
#pragma unroll
            for (int t = 0; t < kMegaConsumerWarps; t++) {
                const float wgt = __expf(lds32f(s_ml + (t * GROUP + g) * 8) - Mx);
                L = fmaf(lds32f(s_ml + (t * GROUP + g) * 8 + 4), wgt, L);
                A = fmaf(lds32f(s_acc + ((t * GROUP + g) * HD + d) * 4), wgt, A);
            }
        }
        ll_st(a.ll_pacc + (pbase * group_total + g0 + g) * HD + d, A, gp);
        if (d == 0) {
            ll_st(a.ll_pml + (pbase * group_total + g0 + g) * 2, Mx, gp);
            ll_st(a.ll_pml + (pbase * group_total + g0 + g) * 2 + 1, L, gp);
        }
    }
    if (pcol) {
        ak5 = clock64();
        const int ps = a.n_phases + 1;
        pcol[9 * ps] = ak1 - ak0;   // q rope
        pcol[10 * ps] = ak2 - ak1;  // K/V loads + scores + online softmax
        pcol[11 * ps] = ak3 - ak2;  // shuffle merge of sub-slots
        pcol[12 * ps] = ak4 - ak3;  // smem + barrier
        pcol[13 * ps] = ak5 - ak4;  // CTA merge + store
    }
}

// Short contexts run one item per QUERY head (GROUP 1): inlined, it is on every token's critical path. The whole-group item of
// long contexts needs most of the register file: out of line, so that its pressure does not spill the rest of the kernel.
template <int HD>
__device__ __forceinline__ void mega_attn_item_q(uint16_t* kv_pool, uint32_t scratch, int kvh, int split, int nsplit, int pos, int tid,
                                                 unsigned long long* pcol, int group_total, int g0, uint32_t gp) {
    mega_attn_item_body<HD, 1>(kv_pool, scratch, kvh, split, nsplit, pos, tid, pcol, group_total, g0, gp);
}
template <int HD, int GROUP>
__device__ __noinline__ void mega_attn_item_group(uint16_t* kv_pool, uint32_t scratch, int kvh, int split, int nsplit, int pos, int tid,
                                                  unsigned long long* pcol, uint32_t gp) {
    mega_attn_item_body<HD, GROUP>(kv_pool, scratch, kvh, split, nsplit, pos, tid, pcol, GROUP, 0, gp);
}

// the producer's view of the weight stream: every chunk this CTA needs, in model order, step after step.
// A chunk = K window `s` of the row group that starts at `row` = rows x KS contiguous bf16 of the tiled image.
struct ChunkCursor {
    int step, pi, row, r1, K, s, P, ks_shift;
    const uint16_t* W;
    __device__ __forceinline__ bool done(int n_steps) const { return step >= n_steps; }
    __device__ __forceinline__ void seek_phase(uint32_t phases, int n_phases, int n_steps) {
        // move to the first row of the next phase that has weights and rows for this CTA
        for (;;) {
            if (pi >= n_phases) {
                pi = 0;
                step++;
                if (step >= n_steps) return;
            }
            const PhaseRegs ph = mega_load_phase(phases, pi);
            if (ph.type != PH_ATTN && ph.r0 < ph.r1) {
                row = ph.r0;
                r1 = ph.r1;
                K = ph.K;
                ks_shift = mega_ks_shift(K);
                P = K >> ks_shift;
                s = 0;
                W = ph.W;
                return;
            }
            pi++;
        }
    }
    __device__ __forceinline__ void get(const uint16_t*& src, uint32_t& bytes) const {
        const int r = min(16, r1 - row);
        src = W + static_cast<size_t>(row) * K + ((static_cast<size_t>(s) * r) << ks_shift);
        bytes = static_cast<uint32_t>(r) << (ks_shift + 1);
    }
    __device__ __forceinline__ void next(uint32_t phases, int n_phases, int n_steps) {
        if (++s < P) return;
        s = 0;
        row += 16;
        if (row >= r1) {
            pi++;
            seek_phase(phases, n_phases, n_steps);
        }
    }
};

// One kernel per (head_dim, GQA group): everything is inlined into it. A call to a non-inlined device function costs stack
// traffic (arguments, callee-saved registers), and with 227 KB of the SM's 256 KB configured as shared memory the L1 that
// would absorb it is ~28 KB for 288 threads: ncu showed 0.9 GB of local-memory L2 traffic per token and a dependent
// L2 round trip at every reload when the phase and attention functions were out of line.
template <int HD, int GROUP>
__global__ void __launch_bounds__(kMegaThreads, 1) mega_decode_kernel() {
    const MegaArgs& a = c_mega;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    MegaSmem sm;
    sm.ring = smem_u32(smem_raw);
    uint32_t p = sm.ring + static_cast<uint32_t>(a.n_stages) * kMegaStageBytes;
    sm.full = p;  p += 8 * kMegaMaxStages;
    sm.empty = p; p += 8 * kMegaMaxStages;
    sm.zero16 = p; p += 16;
    sm.rel = p;   p += 4 * 16;
    sm.keys = p;  p += 8 * 8;
    sm.red = p;   p += 4 * 16;
    sm.part = p;  p += 4 * 2 * kMegaBatchGroups * kMegaConsumerWarps * 16;
    sm.phases = p; p += (48 + 8) * static_cast<uint32_t>(a.n_phases);   // descriptors, then this CTA's (r0, r1) per phase
    p = (p + 15u) & ~15u;
    sm.xfrag = p;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < a.n_stages; s++) {
            mbar_init(sm.full + s * 8, 1);
            mbar_init(sm.empty + s * 8, 2);   // the two warps of the pair that reads the stage
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(sm.rel + s * 4), "r"(0u) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(sm.zero16), "r"(0u) : "memory");
    }
    {   // phase table -> shared memory (static for the whole launch)
        const uint4* src = reinterpret_cast<const uint4*>(a.phases);
        for (int i = tid; i < a.n_phases * 3; i += kMegaThreads) {
            const uint4 v = __ldg(src + i);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(sm.phases + i * 16), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
    }
    for (int i = tid; i < a.n_phases; i += kMegaThreads) {   // row ranges of this CTA, once per launch
        const MegaPhase& mp = a.phases[i];
        int r0 = 0, r1 = 0;
        if (mp.type != PH_ATTN) mega_row_range(mp.N, mp.type == PH_GATEUP ? 2 : 1, blockIdx.x, gridDim.x, r0, r1);
        sts64(sm.phases + a.n_phases * 48 + i * 8, static_cast<unsigned long long>(static_cast<unsigned>(r0)) | (static_cast<unsigned long long>(static_cast<unsigned>(r1)) << 32));
    }
    if (tid == 0 && !a.arg_io) {   // first token and position: one global read per CTA
        sts32f(sm.keys, __int_as_float(*reinterpret_cast<const volatile int32_t*>(a.token)));
        sts32f(sm.keys + 4, __int_as_float(*reinterpret_cast<const volatile int32_t*>(a.position)));
    }
    __syncthreads();

    const int pos0 = a.arg_io ? a.pos0 : __float_as_int(lds32f(sm.keys + 4));
    const int token0 = a.arg_io ? a.token0 : __float_as_int(lds32f(sm.keys));
    __syncthreads();   // sm.keys is reused by the token hand-off

    // Register file: 12 warps x 168 registers at launch. The producer's warpgroup keeps 40 and the two consumer warpgroups
    // take 232 each (3 warps per scheduler: 232 + 232 + 40 <= 512) -- at 168 the consumer code spilled, and local memory
    // is expensive here (see the note above the kernel).
    if (tid >= kMegaConsumerThreads) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        // ================= producer warp: stream every weight chunk this CTA will ever need =================
        if (tid == kMegaConsumerThreads) {
            const uint64_t policy = l2_evict_first_policy();
            ChunkCursor ld{0, 0, 0, 0, 0, 0, 1, 9, nullptr}, pf{0, 0, 0, 0, 0, 0, 1, 9, nullptr};
            ld.seek_phase(sm.phases, a.n_phases, a.n_steps);
            pf.seek_phase(sm.phases, a.n_phases, a.n_steps);
            long long ld_index = 0, pf_index = 0;   // chunks issued to the ring / prefetched into L2
            RingPos rp{0, 0, 0};
            while (!ld.done(a.n_steps)) {
                if (!mbar_test_wait(sm.empty + rp.stage * 8, rp.parity ^ 1)) {
                    // ring full: the consumers are in a latency-bound stretch (hand-off, attention) and HBM would idle. Use the
                    // time to pull the chunks beyond the ring into L2 (bounded run-ahead), then sleep between probes so the
                    // spinning producer does not steal issue slots from the consumer warps that share its scheduler
                    unsigned spins = 0;
                    while (!mbar_test_wait(sm.empty + rp.stage * 8, rp.parity ^ 1)) {
                        if (a.l2_ahead > 0 && !pf.done(a.n_steps) && pf_index < ld_index + a.n_stages + a.l2_ahead) {
                            if (pf_index < ld_index + a.n_stages) {   // catch up with the ring without touching memory
                                pf.next(sm.phases, a.n_phases, a.n_steps);
                                pf_index++;
                                continue;
                            }
                            const uint16_t* psrc;
                            uint32_t pbytes;
                            pf.get(psrc, pbytes);
                            tma_prefetch_l2(psrc, pbytes);
                            pf.next(sm.phases, a.n_phases, a.n_steps);
                            pf_index++;
                            continue;
                        }
                        __nanosleep(a.producer_sleep_ns);
                        if (++spins > (1u << 26)) mega_die(a.abort_flag, 300);
                    }
                }
                const uint16_t* src;
                uint32_t bytes;
                ld.get(src, bytes);
                if (a.debug_nostream) bytes = 16;
                mbar_arrive_expect_tx(sm.full + rp.stage * 8, bytes);
                tma_bulk_g2s(sm.ring + static_cast<uint32_t>(rp.stage) * kMegaStageBytes, src, bytes, sm.full + rp.stage * 8, policy);
                rp.advance(a.n_stages);
                ld.next(sm.phases, a.n_phases, a.n_steps);
                ld_index++;
            }
        }
        return;
    }

    // ================= consumer warps =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    ConsumerState st;
    st.rp = RingPos{0, 0, 0};
    st.best_key = 0ull;
    st.token = token0;
    st.step = 0;
    st.batch = 0;
    for (int step = 0; step < a.n_steps; step++) {
        const int pos = pos0 + step;
        st.step = step;
        st.best_key = 0ull;
        for (int pi = 0; pi < a.n_phases; pi++) {
            const PhaseRegs ph = mega_load_phase(sm.phases, pi);
            if (ph.type == PH_ATTN) {
                const bool prof = a.prof && step == a.n_steps - 1 && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
                const int prow = blockIdx.x == 0 ? 0 : 4, pstride = a.n_phases + 1;
                if (prof) a.prof[(prow + 0) * pstride + pi] = globaltimer_ns();
                const uint32_t gp = a.seq_base + static_cast<uint32_t>(step * a.n_phases + pi) + 1u;
                if (prof) a.prof[(prow + 1) * pstride + pi] = a.prof[(prow + 2) * pstride + pi] = globaltimer_ns();
                const AttnPlan plan = mega_attn_plan(a.nsplit_max, pos + 1, a.attn_tps, a.nh, gridDim.x);
                const int nsplit = plan.nsplit, item = blockIdx.x;
                const int n_items = (plan.per_q_head ? a.nh : a.nkv) * nsplit;
                if (item < n_items) {
                    const int unit = item / nsplit, split = item % nsplit;   // query head or kv head
                    const int kvh = plan.per_q_head ? unit / GROUP : unit;
                    const int g_only = plan.per_q_head ? unit % GROUP : -1;
                    unsigned long long* pcol = prof && blockIdx.x == 0 ? a.prof + pi : nullptr;
                    if (g_only >= 0) mega_attn_item_q<HD>(ph.kv_pool, sm.xfrag, kvh, split, nsplit, pos, tid, pcol, GROUP, g_only, gp);
                    else mega_attn_item_group<HD, GROUP>(ph.kv_pool, sm.xfrag, kvh, split, nsplit, pos, tid, pcol, gp);
                    consumer_bar();   // the scratch aliases the next phase's input fragments
                }
                if (prof) a.prof[(prow + 3) * pstride + pi] = globaltimer_ns();
                if (a.prof && step == a.n_steps - 1 && tid == 0 && blockIdx.x < 160)
                    a.prof[(16 + 160 + blockIdx.x) * pstride + pi] = item < n_items ? globaltimer_ns() : 0ull;
            } else {
                mega_gemv_phase<HD, GROUP>(ph, sm, st, pi, pos, tid);
            }
            if (ph.type == PH_LMHEAD) {
                // CTA-level argmax, then one key per CTA
                const int lane = tid & 31, w = tid >> 5;
                unsigned long long best_key = st.best_key;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best_key, o);
                    best_key = other > best_key ? other : best_key;
                }
                if (lane == 0) sts64(sm.keys + w * 8, best_key);
                consumer_bar();
                if (tid == 0) {
                    unsigned long long k = lds64(sm.keys);
                    for (int i = 1; i < kMegaConsumerWarps; i++) {
                        const unsigned long long o = lds64(sm.keys + i * 8);
                        k = o > k ? o : k;
                    }
                    const uint32_t gp = a.seq_base + static_cast<uint32_t>(step * a.n_phases + pi) + 1u;
                    // the K/V cache lines this CTA appended during the token (plain stores, ordered before this thread by the CTA
                    // barriers since) must be visible before the key announces the token done
                    __threadfence();
#if MEGA_TP
                    for (int p = 0; p < a.tp; p++)   // every rank takes the maximum over all ranks' CTAs itself
                        asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(a.tp_keys[p] + 2 * (a.tp_rank * gridDim.x + blockIdx.x)),
                                     "l"((static_cast<unsigned long long>(gp) << 32) | (k >> 32)),
                                     "l"((static_cast<unsigned long long>(gp) << 32) | (k & 0xffffffffull)) : "memory");
#else
                    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(a.ll_keys + 2 * blockIdx.x),
                                 "l"((static_cast<unsigned long long>(gp) << 32) | (k >> 32)),
                                 "l"((static_cast<unsigned long long>(gp) << 32) | (k & 0xffffffffull)) : "memory");
#endif
                }
            }
        }
    }
    // only CTA 0 needs the last token: it waits for every CTA's key of the last lm_head phase
    if (blockIdx.x != 0) return;
    const int token = mega_poll_token(sm, a.seq_base + static_cast<uint32_t>(a.n_steps * a.n_phases), tid);
    if (tid == 0) {
        a.out_ids[a.n_steps - 1] = token;
        *a.token = token;
        *a.position = pos0 + a.n_steps;
    }
}

}  // namespace MEGA_NS
}  // namespace b2l

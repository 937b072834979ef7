// mega_decode.cuh -- persistent batch-1 decode megakernel: ONE cooperative launch runs n_steps
// whole tokens. The B200-native answer to "a 0.4 ms token is 100 dependent tiny ops":
//
//   * one CTA per SM (148), 8 consumer warps + 1 producer warp
//   * the producer streams this CTA's share of EVERY weight matrix, in model order, through a
//     12-stage x 16 KB shared-memory ring with TMA bulk copies (cp.async.bulk -> UBLKCP) that
//     complete on mbarriers. Weights do not depend on activations, so the stream never stops:
//     it runs ahead across op, layer and even token boundaries, bounded only by the ring.
//   * consumers keep the op's input vector in REGISTERS (64 fp32 per lane), dot it against the
//     bf16 rows as they land (ordered ld.shared with a two-step prefetch, FFMA2), and hand the
//     results on through L2.
//   * attention (split-K over the paged bf16 cache, RoPE and K/V append fused in), SwiGLU,
//     residual adds, both RMSNorms, the lm_head argmax and the token feedback are all inside.
//
// Hand-off between phases. Two builds of the same kernel (template parameter LL):
//   LL = true (default, "dataflow"): no grid barrier. Every value that crosses CTAs is an 8-byte
//     word {fp32 bits, sequence number}; the sequence number names the phase instance that
//     produced it (seq_base + step * n_phases + phase + 1). A reader polls the words it needs until
//     they carry the number of the producing phase (always the phase before its own). The flag
//     travels with the data, so there is no fence and no counter, and a CTA starts a phase as soon
//     as ITS inputs exist.
//     Why nothing is overwritten too early (write-after-read): every phase reads the COMPLETE output
//     of the phase before it. So when some CTA writes an output of phase p+1, it has already seen all
//     of phase p's outputs, hence every CTA has finished computing phase p, hence every CTA has long
//     finished LOADING phase p's inputs (a phase loads its input before its first row). A buffer is
//     therefore safe to rewrite two phases after it was written, and no buffer here is rewritten
//     sooner: h is written by O-proj and down (read by gate/up resp. the next QKV / lm_head), qkv by
//     QKV (read by attention), the partials by attention (read by O-proj), act by gate/up (read by
//     down), the per-CTA argmax keys by lm_head (read at the next token's first phase). Residual
//     read-modify-writes of h touch only rows the same CTA owns in both O-proj and down.
//     The only plain (non-word) data that crosses CTAs is the new token's K/V cache line: its
//     writer fences before publishing its partials, readers fence once per token.
//     Sequence numbers are 32 bits and never reset (host keeps seq_base across launches); buffers
//     start at 0, which is never produced.
//   LL = false: phases are separated by a grid-wide barrier (one release-atomic per CTA + acquire
//     poll), plain fp32 activations. Kept for comparison (B2L_MEGA_LL=0).
//
// HBM sees one sequential read of the model per token; everything else lives in L2 / smem.
// Math is identical to decode_kernels.cuh (the multi-kernel path) up to fp32 summation order.
// This file is compiled TWICE by engine.cu (no include guard): MEGA_TP 0 in namespace b2l::mega1 is the single-GPU kernel,
// MEGA_TP 1 in namespace b2l::megatp the kernel of a tensor-parallel rank. The tensor-parallel additions are confined to
// #if MEGA_TP blocks so that the single-GPU kernel stays instruction for instruction what was tuned in round 1 (the same
// additions behind run-time flags cost it 6-10 %: the kernel's schedule is sensitive to every extra call and branch).
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

// rows [r0, r1) of an N-row matrix owned by CTA `c` of `G` (unit = 2 rows for SwiGLU pairs)
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
// one lane per warp watches the first word pair of the warp's region until it carries `seq`: the whole grid polling
// every word would cost terabytes per second of L2 traffic while the slowest producer finishes
__device__ __forceinline__ void ll_sentinel(const unsigned long long* p, uint32_t seq, int lane, int* abort_flag, int code) {
    if (!c_mega.ll_use_sentinel) return;
    if (lane == 0) {
        unsigned spins = 0;
        for (;;) {
            const uint4 w = ll_ld2(p);
            if (w.y == seq && w.w == seq) break;
            if (++spins > (1u << 22)) mega_die(abort_flag, code);
        }
    }
    __syncwarp();
}
// N x 8 consecutive words -> floats; retried (with a short back-off) until every word carries `seq`
// Position of element k inside the h / act word vectors: within each 256-element block the four word PAIRS a lane
// owns (k = blk*256 + lane*8 + e) are spread 64 words apart, so that a warp's 16-byte load number j covers 512
// contiguous bytes (fully coalesced) instead of 32 scattered 64-byte chunks.
__device__ __forceinline__ int ll_perm(int k) { return (k & ~255) | (((k >> 1) & 3) << 6) | (((k >> 3) & 31) << 1) | (k & 1); }

template <int N>
__device__ __forceinline__ void ll_ld8n(const unsigned long long* const (&p)[N], uint32_t seq, float* out, int* abort_flag, int code,
                                        int pair_stride = 2) {
    unsigned spins = 0;
    for (;;) {
        uint4 w[N][4];
#pragma unroll
        for (int i = 0; i < N; i++) {
#pragma unroll
            for (int j = 0; j < 4; j++) w[i][j] = ll_ld2(p[i] + pair_stride * j);
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

__device__ __forceinline__ void mega_row_range(int N, int unit, int c, int G, int& r0, int& r1) {
    const long long units = N / unit;
    r0 = static_cast<int>(units * c / G) * unit;
    r1 = static_cast<int>(units * (c + 1) / G) * unit;
}

// grid-wide barrier among the consumer threads of all CTAs: arrive (release) then wait for `target`
__device__ __forceinline__ void mega_grid_sync(const MegaArgs& /*unused: c_mega*/, unsigned long long target, int tid) {
    const MegaArgs& a = c_mega;
    consumer_bar();  // all of this CTA's writes are ordered before thread 0's release below
    if (tid == 0) {
        // release: cumulative over the CTA's writes ordered by the bar.sync above; readers use ld.global.cg
        red_release_add_u64(a.bar_counter, 1ull);
        unsigned spins = 0;
        while (ld_acquire_u64(a.bar_counter) < target) {
            if (++spins > (1u << 24)) mega_die(a.abort_flag, 100);
        }
    }
    consumer_bar();
}

// ---- shared memory map: 32-bit shared-space addresses, all kept in registers -----------------
struct MegaSmem {
    uint32_t ring;          // [n_stages][kMegaStageBytes]
    uint32_t full, empty;   // [n_stages] mbarriers each
    uint32_t xs;            // [kMegaXsFloats] fp32
    uint32_t nw;            // [H] bf16: the phase's RMSNorm weight (cp.async before the input poll)
    uint32_t red;           // [32] fp32
    uint32_t part;          // [2][8] fp32 per-chunk partial sums (double buffered)
    uint32_t keys;          // [8] u64
    uint32_t rel;           // [n_stages] u32: completed uses of each stage (see the consumer wait)
    uint32_t phases;        // [n_phases] MegaPhase copies (static for the whole launch)
    uint32_t attn_scratch;
};

// a phase descriptor held in registers (read from the shared-memory copy)
struct PhaseRegs {
    int type, layer, N, K, ks, m;
    int r0, r1;              // this CTA's row range of the phase (precomputed once per launch: the 64-bit divisions of
                             // mega_row_range cost ~300 instructions at every phase entry)
    const uint16_t* W;
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
    r.ks = static_cast<int>(a2.z);
    r.m = static_cast<int>(a2.w);
    unsigned long long rr = lds64(phases + c_mega.n_phases * 48 + pi * 8);
    r.r0 = static_cast<int>(rr & 0xffffffffull);
    r.r1 = static_cast<int>(rr >> 32);
    return r;
}

// position in the ring: stage index and phase parity, advanced incrementally (no div/mod per chunk)
struct RingPos {
    int stage;
    uint32_t parity;
    uint32_t use;  // how many times the ring has wrapped = earlier uses of `stage` (mod 2^32)
    __device__ __forceinline__ void advance(int n_stages) {
        if (++stage == n_stages) {
            stage = 0;
            parity ^= 1;
            use++;
        }
    }
    __device__ __forceinline__ RingPos plus(int k, int n_stages) const {
        RingPos r{stage + k, parity, use};
        while (r.stage >= n_stages) {
            r.stage -= n_stages;
            r.parity ^= 1;
            r.use++;
        }
        return r;
    }
};

// rows of a [N][K] matrix that one 16 KB ring stage holds
__device__ __forceinline__ int mega_rows_per_stage(int ks) { return ks == 1 ? kMegaRows : kMegaRows / ks; }

__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}

// Attention decomposition for this step. While (query heads x context splits) fits the grid, a work
// item is one QUERY head x one split (4x shorter items than per-kv-head, K/V re-read from L2);
// for longer contexts items are (kv head x split) and process the whole GQA group.
struct AttnPlan {
    int nsplit;
    bool per_q_head;
};
__device__ __forceinline__ AttnPlan mega_attn_plan(int nsplit_max, int ctx, int tps, int nh, int G) {
    AttnPlan p;
    const int q_splits = min(nsplit_max, G / nh);          // splits per query head that still fit the grid
    if (q_splits >= 1 && (ctx + q_splits - 1) / q_splits <= 2 * tps) {
        p.per_q_head = true;                               // items of <= 2*tps tokens of ONE query head
        p.nsplit = max(1, min(q_splits, (ctx + 31) / 32));
    } else {
        p.per_q_head = false;
        p.nsplit = max(1, min(nsplit_max, (ctx + tps - 1) / tps));
    }
    return p;
}

// attention output elements [k, k+8) of this step: merge the split-K partials.
// Loads are issued four splits at a time (independent of each other), then merged online.
__device__ __forceinline__ void mega_attn_combine8(const MegaArgs& /*unused: c_mega*/, int k, int nsplit, float* out) {
    const MegaArgs& a = c_mega;
    const int head = k / a.hd, d = k % a.hd;  // 8 consecutive k never straddle a head (hd % 8 == 0)
    const int group = a.nh / a.nkv, kvh = head / group, g = head % group;
    const size_t rbase = static_cast<size_t>(kvh) * a.nsplit_max;
    float Mx = -INFINITY, L = 0.f, acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.f;
    for (int s0 = 0; s0 < nsplit; s0 += 4) {
        float2 ml[4];
        float4 p0[4], p1[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int s = min(s0 + u, nsplit - 1);  // clamped duplicates are masked below
            ml[u] = __ldcg(reinterpret_cast<const float2*>(a.part_ml + ((rbase + s) * group + g) * 2));
            const float* pa = a.part_acc + ((rbase + s) * group + g) * a.hd + d;
            p0[u] = __ldcg(reinterpret_cast<const float4*>(pa));
            p1[u] = __ldcg(reinterpret_cast<const float4*>(pa + 4));
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (s0 + u >= nsplit || ml[u].x == -INFINITY) continue;
            const float mn = fmaxf(Mx, ml[u].x);
            const float c_old = __expf(Mx - mn), c_new = __expf(ml[u].x - mn);  // exp(-inf) = 0 on the first split
            L = L * c_old + ml[u].y * c_new;
            acc[0] = acc[0] * c_old + p0[u].x * c_new; acc[1] = acc[1] * c_old + p0[u].y * c_new;
            acc[2] = acc[2] * c_old + p0[u].z * c_new; acc[3] = acc[3] * c_old + p0[u].w * c_new;
            acc[4] = acc[4] * c_old + p1[u].x * c_new; acc[5] = acc[5] * c_old + p1[u].y * c_new;
            acc[6] = acc[6] * c_old + p1[u].z * c_new; acc[7] = acc[7] * c_old + p1[u].w * c_new;
            Mx = mn;
        }
    }
    const float inv = 1.0f / L;
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = acc[i] * inv;
}

// dataflow variant: the partials are {value, seq} words written by the attention items of phase `seq`
template <int NB>  // splits fetched per round trip
__device__ __forceinline__ void mega_attn_combine8_ll(int k, int nsplit, uint32_t seq, float* out) {
    const MegaArgs& a = c_mega;
    const int head = k / a.hd, d = k % a.hd;
    const int group = a.nh / a.nkv, kvh = head / group, g = head % group;
    const size_t rbase = static_cast<size_t>(kvh) * a.nsplit_max;
    float Mx = -INFINITY, L = 0.f, acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.f;
    for (int s0 = 0; s0 < nsplit; s0 += NB) {
        float ml[NB][2], pa[NB][8];
        unsigned spins = 0;
        for (;;) {
            uint4 wm[NB], wa[NB][4];
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const int sp = min(s0 + u, nsplit - 1);
                const size_t rec = (rbase + sp) * group + g;
                wm[u] = ll_ld2(a.ll_pml + rec * 2);
#pragma unroll
                const unsigned long long* pa0 = a.ll_pacc + ll_perm(static_cast<int>(rec) * a.hd + d);   // d % 8 == 0: pairs 64 words apart
#pragma unroll
                for (int j = 0; j < 4; j++) wa[u][j] = ll_ld2(pa0 + 64 * j);
            }
            bool ok = true;
#pragma unroll
            for (int u = 0; u < NB; u++) {
                ok = ok && wm[u].y == seq && wm[u].w == seq;
#pragma unroll
                for (int j = 0; j < 4; j++) ok = ok && wa[u][j].y == seq && wa[u][j].w == seq;
            }
            if (ok) {
#pragma unroll
                for (int u = 0; u < NB; u++) {
                    ml[u][0] = __uint_as_float(wm[u].x);
                    ml[u][1] = __uint_as_float(wm[u].z);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        pa[u][2 * j] = __uint_as_float(wa[u][j].x);
                        pa[u][2 * j + 1] = __uint_as_float(wa[u][j].z);
                    }
                }
                break;
            }
            __nanosleep(c_mega.poll_sleep_ns);
            if (++spins > (1u << 22)) mega_die(a.abort_flag, 120);
        }
#pragma unroll
        for (int u = 0; u < NB; u++) {
            if (s0 + u >= nsplit || ml[u][0] == -INFINITY) continue;
            const float mn = fmaxf(Mx, ml[u][0]);
            const float c_old = __expf(Mx - mn), c_new = __expf(ml[u][0] - mn);
            L = L * c_old + ml[u][1] * c_new;
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = acc[i] * c_old + pa[u][i] * c_new;
            Mx = mn;
        }
    }
    const float inv = 1.0f / L;
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = acc[i] * inv;
}

// everything the consumer threads carry across phases, in registers
struct ConsumerState {
    RingPos rp;
    unsigned long long nbar;      // grid barriers passed in this launch
    unsigned long long epoch;
    unsigned long long best_key;  // running argmax of this step's logits
    int token, step;
    bool need_barrier;            // false only for the very first phase of the launch
};

// barrier that precedes a phase: wait for the previous phase's outputs; at a token boundary also
// pick up the argmax that becomes the next input token
__device__ __forceinline__ void mega_phase_barrier(const MegaArgs& /*unused: c_mega*/, ConsumerState& st, bool token_boundary, int tid) {
    const MegaArgs& a = c_mega;
    if (!st.need_barrier) {
        st.need_barrier = true;
        return;
    }
    st.nbar++;
    mega_grid_sync(a, st.epoch + st.nbar * gridDim.x, tid);
    if (token_boundary) {
        st.token = argmax_key_index(ld_acquire_u64(a.argmax_keys + ((st.step + 2) % 3)));  // previous step's key
        if (blockIdx.x == 0 && tid == 0) a.out_ids[st.step - 1] = st.token;
    }
}

// token boundary in dataflow mode: every CTA published the best (logit, index) key of its lm_head rows as two
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

#if MEGA_TP
// Tensor parallel, row-parallel phases (O-proj, down): second pass over the rows this lane produced. The tp partial sums of
// a row arrive over NVLink in this rank's own slab; they are added to the residual in rank order -- every rank computes
// bit-identical h -- and the row is published locally like any other activation word. One round of rows at a time, the
// tp loads of a row in flight together (out of line: its registers must not weigh on the streaming loop).
__device__ __noinline__ void mega_tp_collect(const PhaseRegs ph, uint32_t gp, int token, int tid) {
    const MegaArgs& a = c_mega;
    constexpr int kMaxTp = 8;
    const int lane = tid & 31, w = tid >> 5;
    const int ks_shift = ph.ks == 1 ? 0 : ph.ks == 2 ? 1 : 2;
    const int q = w & (ph.ks - 1), rloc = w >> ks_shift, my_t = lane >> 3;
    if ((lane & 7) != 0 || q != 0) return;   // the lanes that held the row sums of the phase
    const int r0 = ph.r0, r1 = ph.r1;
    const int groups_per_round = kMegaConsumerWarps >> ks_shift;
    const int n_groups = (r1 - r0 + kMegaRows - 1) / kMegaRows;
    const int n_rounds = (n_groups + groups_per_round - 1) >> (3 - ks_shift);
    const bool resid_h = ph.type == PH_DOWN || (ph.type == PH_OPROJ && ph.layer != 0);
    const bool resid_e = ph.type == PH_OPROJ && ph.layer == 0;
    const unsigned long long* slab = a.tp_slab[a.tp_rank] + (static_cast<size_t>(ph.type == PH_DOWN ? 1 : 0) * a.tp) * a.Hpad;
    for (int rd = 0; rd < n_rounds; rd++) {
        const int rg = rd * groups_per_round + rloc, row_t = r0 + rg * kMegaRows + my_t;
        if (rg >= n_groups || row_t >= r1) continue;
        float sum = 0.f;
        if (resid_h) sum = ld_cg_early_f32(a.ll_h + ll_perm(row_t));
        else if (resid_e) sum = bf16_bits_to_f32(a.embed[static_cast<size_t>(token) * a.H + row_t]);
        const unsigned long long* src = slab + ll_perm(row_t);
        unsigned spins = 0;
        for (;;) {
            uint2 wv[kMaxTp];
#pragma unroll
            for (int p = 0; p < kMaxTp; p++)
                if (p < a.tp) wv[p] = ll_ld1_sys(src + static_cast<size_t>(p) * a.Hpad);
            bool ok = true;
#pragma unroll
            for (int p = 0; p < kMaxTp; p++) ok = ok && (p >= a.tp || wv[p].y == gp);
            if (ok) {
#pragma unroll
                for (int p = 0; p < kMaxTp; p++)
                    if (p < a.tp) sum += __uint_as_float(wv[p].x);
                break;
            }
            if (++spins > (1u << 25)) mega_die(a.abort_flag, 230 + ph.type);
        }
        ll_st(a.ll_h + ll_perm(row_t), sum, gp);
    }
}
#endif

// ---- one GEMV-type phase for one CTA ---------------------------------------------------------
#ifdef MEGA_GEMV_INLINE
#define MEGA_GEMV_ATTR __forceinline__
#else
#define MEGA_GEMV_ATTR __noinline__
#endif
template <int M, bool SPLIT, bool LL>  // SPLIT: K is split over ks > 1 warps; LL: dataflow words instead of grid barriers
__device__ MEGA_GEMV_ATTR void mega_gemv_phase(const MegaArgs&, const PhaseRegs ph, const MegaSmem sm, ConsumerState& st_ref,
                                             int pi, int pos, int tid) {
    const MegaArgs& a = c_mega;
    ConsumerState st = st_ref;  // work on a register copy; written back once at the end
    const int lane = tid & 31, w = tid >> 5;
    const int ks = SPLIT ? ph.ks : 1, slice = 256 * M;
    const int ks_shift = ks == 1 ? 0 : ks == 2 ? 1 : 2;  // ks in {1, 2, 4}
    const int q = w & (ks - 1), rloc = w >> ks_shift;
    const int K = ph.K, type = ph.type;
    const bool prof = a.prof && st.step == a.n_steps - 1 && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
    unsigned long long* prof_col = a.prof + pi;
    const int prow = blockIdx.x == 0 ? 0 : 4, pstride = a.n_phases + 1;
    if (prof) prof_col[(prow + 0) * pstride] = globaltimer_ns();
    volatile int* progress = a.debug_progress && tid == 0 ? a.abort_flag + 1 + blockIdx.x : nullptr;
    if (progress) *progress = st.step * 100000 + pi * 100 + 1;

    // ---- static prologue: nothing here depends on other CTAs, so it overlaps the barrier wait ----
    int r0, r1;
    r0 = ph.r0;
    r1 = ph.r1;
    if (ph.norm_w) {
        // RMSNorm weight -> shared memory, asynchronously (no registers held across the input poll, no scoreboard wait)
        for (int k = tid * 8; k < K; k += kMegaConsumerThreads * 8)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sm.nw + k * 2), "l"(ph.norm_w + k) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const bool from_embed = (type == PH_QKV && ph.layer == 0);
    const uint32_t gp = a.seq_base + static_cast<uint32_t>(st.step * a.n_phases + pi) + 1u;  // this phase's sequence number
    const uint32_t want = gp - 1u;                                                           // inputs come from the phase before
    if (LL) {
        if (from_embed && st.step > 0) {
            st.token = mega_poll_token(sm, want, tid);
            if (blockIdx.x == 0 && tid == 0) a.out_ids[st.step - 1] = st.token;
            __threadfence();  // order this step's KV-cache reads after everything the previous step published
        }
    } else {
        mega_phase_barrier(a, st, from_embed && st.step > 0, tid);
    }
    if (prof) prof_col[(prow + 1) * pstride] = globaltimer_ns();
    if (progress) *progress = st.step * 100000 + pi * 100 + 2;

    // ---- input vector -> registers (fused RMSNorm / split-K attention merge) ----
    const float* xsrc = type == PH_DOWN ? a.act : a.h;
    const int nsplit = mega_attn_plan(a.nsplit_max, pos + 1, a.attn_tps, a.nh, gridDim.x).nsplit;
    const int token = st.token;
    float xr[M * 8];
    if (!SPLIT) {
        // every warp needs the same K floats: fetch once per CTA, then fan out through smem
        if (LL && !from_embed) {
            // wait (one lane per warp) for the first words of this warp's share, then fetch; stragglers are retried
            if (type == PH_OPROJ) {
                const int k0 = min((tid & ~31) * 8, K - 8), head = k0 / a.hd, group = a.nh / a.nkv;
                ll_sentinel(a.ll_pml + ((static_cast<size_t>(head / group) * a.nsplit_max) * group + head % group) * 2, want, lane, a.abort_flag, 140);
            } else {
                ll_sentinel((type == PH_DOWN ? a.ll_act : a.ll_h) + min((tid & ~31) * 4, K - 4), want, lane, a.abort_flag, 141 + type);
            }
        }
        if (type == PH_OPROJ) {
            for (int k = tid * 8; k < K; k += kMegaConsumerThreads * 8) {
                float v[8];
                if (LL) mega_attn_combine8_ll<4>(k, nsplit, want, v);
                else mega_attn_combine8(a, k, nsplit, v);
                sts128f(sm.xs + k * 4, make_float4(v[0], v[1], v[2], v[3]));
                sts128f(sm.xs + k * 4 + 16, make_float4(v[4], v[5], v[6], v[7]));
            }
        } else if (LL && !from_embed) {
            // a warp fetches 256 consecutive words with four fully coalesced 16-byte loads per lane (512 contiguous
            // bytes per instruction), all in flight: one round trip for K <= 2048
            const unsigned long long* src = (type == PH_DOWN ? a.ll_act : a.ll_h);
#ifdef MEGA_X_ROT
            const int nblk = K >> 8;
            for (int kb = w; kb < nblk; kb += kMegaConsumerWarps) {
                const int k0 = ((kb + static_cast<int>(blockIdx.x)) % nblk) << 8;   // CTAs walk the vector in different orders: no L2 slice sees the whole grid at once
#else
            for (int k0 = w * 256; k0 < K; k0 += kMegaConsumerWarps * 256) {
#endif
                unsigned spins = 0;
                for (;;) {
                    uint4 wd[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) wd[j] = ll_ld2(src + k0 + j * 64 + lane * 2);
                    bool ok = true;
#pragma unroll
                    for (int j = 0; j < 4; j++) ok = ok && wd[j].y == want && wd[j].w == want;
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sm.xs + (k0 + lane * 8 + 2 * j) * 4), "f"(__uint_as_float(wd[j].x)),
                                         "f"(__uint_as_float(wd[j].z)) : "memory");
                        break;
                    }
                    __nanosleep(c_mega.poll_sleep_ns);
                    if (++spins > (1u << 22)) mega_die(a.abort_flag, 150 + type);
                }
            }
        } else {
            for (int k = tid * 4; k < K; k += kMegaConsumerThreads * 4) {
                float4 v;
                if (from_embed) {
                    const uint2 e = __ldg(reinterpret_cast<const uint2*>(a.embed + static_cast<size_t>(token) * a.H + k));
                    v = make_float4(bf16lo(e.x), bf16hi(e.x), bf16lo(e.y), bf16hi(e.y));
                } else {
                    v = __ldcg(reinterpret_cast<const float4*>(xsrc + k));
                }
                sts128f(sm.xs + k * 4, v);
            }
        }
        if (ph.norm_w) asm volatile("cp.async.wait_group 0;" ::: "memory");
        consumer_bar();
#pragma unroll
        for (int i = 0; i < M; i++) {
            const float4 v0 = lds128f(sm.xs + (i * 256 + lane * 8) * 4), v1 = lds128f(sm.xs + (i * 256 + lane * 8) * 4 + 16);
            xr[i * 8 + 0] = v0.x; xr[i * 8 + 1] = v0.y; xr[i * 8 + 2] = v0.z; xr[i * 8 + 3] = v0.w;
            xr[i * 8 + 4] = v1.x; xr[i * 8 + 5] = v1.y; xr[i * 8 + 6] = v1.z; xr[i * 8 + 7] = v1.w;
        }
    } else if (LL && !from_embed && type != PH_OPROJ) {
        const unsigned long long* src = (type == PH_DOWN ? a.ll_act : a.ll_h) + q * slice + lane * 2;   // permuted layout: see ll_perm
        ll_sentinel(src - lane * 2, want, lane, a.abort_flag, 160 + type);
        if (M == 8) {  // four 8-word groups (16 x 16-byte loads) in flight per lane: two round trips for the slice
#pragma unroll
            for (int i = 0; i < 8; i += 4) {
                const unsigned long long* const pp[4] = {src + i * 256, src + (i + 1) * 256, src + (i + 2) * 256, src + (i + 3) * 256};
                ll_ld8n<4>(pp, want, &xr[i * 8], a.abort_flag, 170 + type, 64);
            }
        } else {
#pragma unroll
            for (int i = 0; i + 1 < M; i += 2) {
                const unsigned long long* const pp[2] = {src + i * 256, src + (i + 1) * 256};
                ll_ld8n<2>(pp, want, &xr[i * 8], a.abort_flag, 170 + type, 64);
            }
            if (M & 1) {
                const unsigned long long* const pp[1] = {src + (M - 1) * 256};
                ll_ld8n<1>(pp, want, &xr[(M - 1) * 8], a.abort_flag, 170 + type, 64);
            }
        }
    } else {
        if (LL && type == PH_OPROJ) {
            const int head = (q * slice) / a.hd, group = a.nh / a.nkv;
            ll_sentinel(a.ll_pml + ((static_cast<size_t>(head / group) * a.nsplit_max) * group + head % group) * 2, want, lane, a.abort_flag, 140);
        }
#pragma unroll
        for (int i = 0; i < M; i++) {
            const int k = q * slice + i * 256 + lane * 8;
            if (type == PH_OPROJ) {
                if (LL) mega_attn_combine8_ll<2>(k, nsplit, want, &xr[i * 8]);
                else mega_attn_combine8(a, k, nsplit, &xr[i * 8]);
                continue;
            }
            float4 v0, v1;
            if (from_embed) {
                const uint4 e = __ldg(reinterpret_cast<const uint4*>(a.embed + static_cast<size_t>(token) * a.H + k));
                v0 = make_float4(bf16lo(e.x), bf16hi(e.x), bf16lo(e.y), bf16hi(e.y));
                v1 = make_float4(bf16lo(e.z), bf16hi(e.z), bf16lo(e.w), bf16hi(e.w));
            } else {
                v0 = __ldcg(reinterpret_cast<const float4*>(xsrc + k));
                v1 = __ldcg(reinterpret_cast<const float4*>(xsrc + k + 4));
            }
            xr[i * 8 + 0] = v0.x; xr[i * 8 + 1] = v0.y; xr[i * 8 + 2] = v0.z; xr[i * 8 + 3] = v0.w;
            xr[i * 8 + 4] = v1.x; xr[i * 8 + 5] = v1.y; xr[i * 8 + 6] = v1.z; xr[i * 8 + 7] = v1.w;
        }
    }
    if (ph.norm_w) {
        float ssq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < M * 8; i++) ssq[i & 3] = fmaf(xr[i], xr[i], ssq[i & 3]);
        float ss = warp_sum((ssq[0] + ssq[1]) + (ssq[2] + ssq[3]));
        if (SPLIT) {  // a warp holds only its K slice: add the other slices' sums (warps 0..ks-1)
            if (lane == 0) sts32f(sm.red + w * 4, ss);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            consumer_bar();
            ss = 0.f;
            for (int i = 0; i < ks; i++) ss += lds32f(sm.red + i * 4);
        }
        const float inv = rsqrtf(ss / static_cast<float>(K) + a.eps);
#pragma unroll
        for (int i = 0; i < M; i++) {
            const uint4 nwi = lds128(sm.nw + (q * slice + i * 256 + lane * 8) * 2);
            xr[i * 8 + 0] = bf16lo(nwi.x) * (xr[i * 8 + 0] * inv);
            xr[i * 8 + 1] = bf16hi(nwi.x) * (xr[i * 8 + 1] * inv);
            xr[i * 8 + 2] = bf16lo(nwi.y) * (xr[i * 8 + 2] * inv);
            xr[i * 8 + 3] = bf16hi(nwi.y) * (xr[i * 8 + 3] * inv);
            xr[i * 8 + 4] = bf16lo(nwi.z) * (xr[i * 8 + 4] * inv);
            xr[i * 8 + 5] = bf16hi(nwi.z) * (xr[i * 8 + 5] * inv);
            xr[i * 8 + 6] = bf16lo(nwi.w) * (xr[i * 8 + 6] * inv);
            xr[i * 8 + 7] = bf16hi(nwi.w) * (xr[i * 8 + 7] * inv);
        }
    }
    if (prof) prof_col[(prow + 2) * pstride] = globaltimer_ns();
    const bool prof_all = a.prof && st.step == a.n_steps - 1 && tid == 0 && blockIdx.x < 160;
    if (prof_all) prof_col[(16 + blockIdx.x) * pstride] = globaltimer_ns();
    if (progress) *progress = st.step * 100000 + pi * 100 + 3;

    // ---- stream this CTA's rows: a warp takes kMegaRows rows (x its K slice) at a time ----
    const int RS = kMegaRows >> ks_shift, rs_shift = 2 - ks_shift;  // rows per ring stage (4, 2 or 1)
    const int n_stage_total = (r1 - r0 + RS - 1) >> rs_shift;           // stages the producer fills for this phase
    const bool resid_h = type == PH_DOWN || (type == PH_OPROJ && ph.layer != 0);
    const bool resid_e = type == PH_OPROJ && ph.layer == 0;
#if MEGA_TP
    const bool tp_phase = type == PH_OPROJ || type == PH_DOWN;   // K-sharded: this rank holds a partial sum of every row
#else
    constexpr bool tp_phase = false;
#endif
    const int my_t = lane >> 3;                         // after the butterfly, lane 8*t holds row t of the item
    const bool out_lane = (lane & 7) == 0;
    const bool hi16 = lane & 16, hi8 = lane & 8;
    const int n_stages = a.n_stages;
    int* const abort_flag = a.abort_flag;
    unsigned long long best_key = st.best_key;
    const RingPos rp0 = st.rp;
    const int groups_per_round = kMegaConsumerWarps >> ks_shift;  // row groups (of kMegaRows rows) the CTA handles at once
    const int n_groups = (r1 - r0 + kMegaRows - 1) / kMegaRows;
    const int n_rounds = (n_groups + groups_per_round - 1) >> (3 - ks_shift);
    const uint32_t lane_off = static_cast<uint32_t>(q) * slice * 2 + lane * 16;
    // this warp's first item: row group rloc -> stage index rloc * ks
    RingPos sp0 = rp0.plus(rloc * ks, n_stages);
    const int stage_step = groups_per_round * ks;       // = kMegaConsumerWarps stages per round
#ifdef MEGA_PROF_ROUNDS
    long long pr_wait = 0, pr_math = 0, pr_tail = 0, pr_t0 = 0, pr_t1 = 0, pr_t2 = 0;
#endif
    for (int rd = 0; rd < n_rounds; rd++) {
#ifdef MEGA_PROF_ROUNDS
        pr_t0 = clock64();
#endif
        const int rg = rd * groups_per_round + rloc;    // this warp's row group
        const int row0 = r0 + rg * kMegaRows;
        const bool live = rg < n_groups;
        const int row_t = row0 + my_t;
        const bool row_live = live && row_t < r1;
        float resid = 0.f;  // residual input, fetched before the wait so its L2 latency overlaps
        if (out_lane && row_live && q == 0 && !tp_phase) {   // (a tensor-parallel rank adds the residual when it collects the partial sums)
            if (resid_h) resid = LL ? ld_cg_early_f32(a.ll_h + ll_perm(row_t)) : ld_cg_early_f32(a.h + row_t);
            else if (resid_e) resid = bf16_bits_to_f32(a.embed[static_cast<size_t>(token) * a.H + row_t]);
        }
        float2 acc[kMegaRows];
#pragma unroll
        for (int t = 0; t < kMegaRows; t++) acc[t] = make_float2(0.f, 0.f);
        if (live) {
            uint32_t row_addr[kMegaRows];  // shared address of (row t, this warp's K slice, this lane)
            const int first_stage = rg * ks;
            {
                RingPos sp = sp0;
#pragma unroll
                for (int j = 0; j < kMegaRows; j++) {  // stage j of this item holds rows [j*RS, (j+1)*RS)
                    if (j < ks) {
                        const bool exists = first_stage + j < n_stage_total;
                        if (exists) {
                            // Successive uses of a stage belong to different warps, so this warp can get
                            // here before the PREVIOUS use has even landed, and a parity wait cannot tell
                            // "one phase behind" from "done". The release counter of the previous use is
                            // the proof that it landed (its reader finished); only then is the parity wait
                            // unambiguous. (Usually satisfied on the first load.)
                            unsigned spins = 0;
                            for (;;) {
                                uint32_t done;
                                asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(done) : "r"(sm.rel + sp.stage * 4) : "memory");
                                if (static_cast<int32_t>(done - sp.use) >= 0) break;
                                if (++spins > (1u << 28)) mega_die(abort_flag, 210 + type);
                            }
                            mbar_wait(sm.full + sp.stage * 8, sp.parity, abort_flag, 200 + type);
                        }
                        const uint32_t base = sm.ring + static_cast<uint32_t>(sp.stage) * kMegaStageBytes + lane_off;
#pragma unroll
                        for (int t = 0; t < kMegaRows; t++) {
                            if ((t >> rs_shift) == j) row_addr[t] = exists && (row0 + t < r1) ? base + static_cast<uint32_t>(t & (RS - 1)) * K * 2 : sm.xs;
                        }
                        sp.advance(n_stages);
                    }
                }
            }
#ifdef MEGA_PROF_ROUNDS
            pr_t1 = clock64();
#endif
            uint4 wv[3][kMegaRows];   // two steps of prefetch
#pragma unroll
            for (int t = 0; t < kMegaRows; t++) wv[0][t] = lds128_ordered(row_addr[t]);
            if (M > 1) {
#pragma unroll
                for (int t = 0; t < kMegaRows; t++) wv[1][t] = lds128_ordered(row_addr[t] + 512);
            }
#pragma unroll
            for (int i = 0; i < M; i++) {
                if (i + 2 < M) {
#pragma unroll
                    for (int t = 0; t < kMegaRows; t++) wv[(i + 2) % 3][t] = lds128_ordered(row_addr[t] + (i + 2) * 512);
                }
                __syncwarp();   // scheduling fence: later steps' loads are ISSUED before this step's FFMA2s (ptxas otherwise
                                // sinks every load to just before its first use and the loop runs at shared-memory latency)
#pragma unroll
                for (int t = 0; t < kMegaRows; t++) {
                    const uint4 v = wv[i % 3][t];
                    float2 s2 = acc[t];  // packed fp32x2 FMA (FFMA2): even elements in .x, odd in .y
                    s2 = __ffma2_rn(make_float2(bf16lo(v.x), bf16hi(v.x)), make_float2(xr[i * 8 + 0], xr[i * 8 + 1]), s2);
                    s2 = __ffma2_rn(make_float2(bf16lo(v.y), bf16hi(v.y)), make_float2(xr[i * 8 + 2], xr[i * 8 + 3]), s2);
                    s2 = __ffma2_rn(make_float2(bf16lo(v.z), bf16hi(v.z)), make_float2(xr[i * 8 + 4], xr[i * 8 + 5]), s2);
                    s2 = __ffma2_rn(make_float2(bf16lo(v.w), bf16hi(v.w)), make_float2(xr[i * 8 + 6], xr[i * 8 + 7]), s2);
                    acc[t] = s2;
                }
            }
            __syncwarp();
            if (lane == 0) {  // hand the stage(s) back: kMegaRows arrivals per stage in total
                {
                    RingPos rel = sp0;
#pragma unroll
                    for (int j = 0; j < kMegaRows; j++) {
                        if (j < ks && first_stage + j < n_stage_total) {
                            // publish "this use was read" BEFORE arriving: the stage can only be refilled (and
                            // its next use released) after every arrival of this use, so the counter never runs
                            // backwards (a late store after the arrive once did, and deadlocked a waiter)
                            asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(sm.rel + rel.stage * 4), "r"(rel.use + 1) : "memory");
                            mbar_arrive_n(sm.empty + rel.stage * 8, kMegaRows >> ks_shift);
                        }
                        rel.advance(n_stages);
                    }
                }
            }
        }
#ifdef MEGA_PROF_ROUNDS
        pr_t2 = clock64();
        if (live) { pr_wait += pr_t1 - pr_t0; pr_math += pr_t2 - pr_t1; }
#endif
        sp0 = sp0.plus(stage_step, n_stages);
        // transposed butterfly: 6 shuffles reduce all four rows; lane 8*t ends up with row t's sum.
        // Rows past the end of the range read a duplicate row; their sums are never stored.
        const float a0 = acc[0].x + acc[0].y, a1 = acc[1].x + acc[1].y, a2 = acc[2].x + acc[2].y, a3 = acc[3].x + acc[3].y;
        float s0 = hi16 ? a2 : a0, s1 = hi16 ? a3 : a1;
        s0 += __shfl_xor_sync(0xffffffffu, hi16 ? a0 : a2, 16);
        s1 += __shfl_xor_sync(0xffffffffu, hi16 ? a1 : a3, 16);
        float s = (hi8 ? s1 : s0) + __shfl_xor_sync(0xffffffffu, hi8 ? s0 : s1, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (SPLIT) {
            // K was split over ks warps: add the slices through smem (one CTA barrier per round)
            const uint32_t part = sm.part + (rd & 1) * (kMegaConsumerWarps * kMegaRows * 4);
            if (out_lane) sts32f(part + (w * kMegaRows + my_t) * 4, row_live ? s : 0.f);
            // only the ks warps that share this row group meet (named barrier 2 + group): the other groups stream on
            asm volatile("bar.sync %0, %1;" ::"r"(2 + rloc), "r"(32 * ks) : "memory");
            if (q != 0) continue;
            s = 0.f;
            for (int i = 0; i < ks; i++) s += lds32f(part + ((w + i) * kMegaRows + my_t) * 4);
        }
        if (type == PH_GATEUP) {
            const float up = __shfl_down_sync(0xffffffffu, s, 8);  // rows are (gate, up) pairs
            s = (s / (1.0f + __expf(-s))) * up;
        }
        if (out_lane && row_live) {
            if (type == PH_QKV) {
                if (LL) ll_st(a.ll_qkv + row_t, s, gp);
                else a.qkv[row_t] = s;
            } else if (type == PH_GATEUP) {
                if ((my_t & 1) == 0) {
                    if (LL) ll_st(a.ll_act + ll_perm(row_t >> 1), s, gp);
                    else a.act[row_t >> 1] = s;
                }
            } else if (type == PH_LMHEAD) {
                a.logits[row_t] = s;
                const unsigned long long key = argmax_key(s, (MEGA_TP ? a.vocab_base : 0) + row_t);   // global vocabulary index
                best_key = key > best_key ? key : best_key;
            } else {
#if MEGA_TP
                {   // partial sum -> every rank's slab (own included) over NVLink, one 8-byte word each
                    const size_t off = (static_cast<size_t>(type == PH_DOWN ? 1 : 0) * a.tp + a.tp_rank) * a.Hpad + ll_perm(row_t);
                    for (int p = 0; p < a.tp; p++) ll_st_sys(a.tp_slab[p] + off, s, gp);
                }
#else
                if (LL) ll_st(a.ll_h + ll_perm(row_t), resid + s, gp);  // O-proj / down: residual add
                else a.h[row_t] = resid + s;
#endif
            }
        }
#ifdef MEGA_PROF_ROUNDS
        pr_tail += clock64() - pr_t2;
#endif
    }
#ifdef MEGA_PROF_ROUNDS
    if (prof && blockIdx.x == 0) {
        prof_col[9 * pstride] = pr_wait;
        prof_col[10 * pstride] = pr_math;
        prof_col[11 * pstride] = pr_tail;
        prof_col[12 * pstride] = n_rounds;
    }
#endif
    st.rp = rp0.plus(n_stage_total, n_stages);
    st.best_key = best_key;
    st_ref = st;
    if (progress) *progress = st.step * 100000 + pi * 100 + 4;
    if (prof) prof_col[(prow + 3) * pstride] = globaltimer_ns();
    if (prof_all) prof_col[(16 + 160 + blockIdx.x) * pstride] = globaltimer_ns();
}

// ---- attention work item: (kv head, split); partials are merged by the O-proj phase's x load -----
template <int HD, int GROUP, bool LL>
__device__ __noinline__ void mega_attn_item(const MegaArgs&, uint16_t* kv_pool, uint32_t scratch, int kvh, int split, int nsplit,
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

    // rotate-half RoPE of one 8-wide slice of a head living in the fused qkv row (element offset `head`)
    const uint32_t want = gp - 1u;
#ifdef MEGA_X_ROPE_EARLY
    float4 csr[4];   // (cos, sin) of this lane's 8 rotation pairs: fetched before any wait
    {
        const int d0e = sl * 8, j0e = d0e < HALF ? d0e : d0e - HALF;
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(csr[i].x), "=f"(csr[i].y), "=f"(csr[i].z), "=f"(csr[i].w) : "l"(cs + 2 * (j0e + 2 * i)) : "memory");
    }
#endif
    auto rope_slice = [&](int head, float* out) {
        const int d0 = sl * 8, j0r = d0 < HALF ? d0 : d0 - HALF;  // the slice lies in one half (HALF % 8 == 0)
        float xx[16];
        if (LL) {
            const unsigned long long* const pp[2] = {a.ll_qkv + head + j0r, a.ll_qkv + head + j0r + HALF};
            ll_ld8n<2>(pp, want, xx, a.abort_flag, 180);
        } else {
            const float* hp = a.qkv + head;
            const float4 xa0 = __ldcg(reinterpret_cast<const float4*>(hp + j0r)), xa1 = __ldcg(reinterpret_cast<const float4*>(hp + j0r + 4));
            const float4 xb0 = __ldcg(reinterpret_cast<const float4*>(hp + j0r + HALF)), xb1 = __ldcg(reinterpret_cast<const float4*>(hp + j0r + HALF + 4));
            xx[0] = xa0.x; xx[1] = xa0.y; xx[2] = xa0.z; xx[3] = xa0.w; xx[4] = xa1.x; xx[5] = xa1.y; xx[6] = xa1.z; xx[7] = xa1.w;
            xx[8] = xb0.x; xx[9] = xb0.y; xx[10] = xb0.z; xx[11] = xb0.w; xx[12] = xb1.x; xx[13] = xb1.y; xx[14] = xb1.z; xx[15] = xb1.w;
        }
        const float* x0 = xx;
        const float* x1 = xx + 8;
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
#ifdef MEGA_X_ROPE_EARLY
            const float4 c2 = csr[i >> 1];
#else
            const float4 c2 = __ldg(reinterpret_cast<const float4*>(cs + 2 * (j0r + i)));  // (c, s, c', s')
#endif
            out[i] = d0 < HALF ? x0[i] * c2.x - x1[i] * c2.y : x1[i] * c2.x + x0[i] * c2.y;
            out[i + 1] = d0 < HALF ? x0[i + 1] * c2.z - x1[i + 1] * c2.w : x1[i + 1] * c2.z + x0[i + 1] * c2.w;
        }
    };
    constexpr int U = 4;  // token slots per lane group in flight: all K/V loads of a block are issued before any math
    constexpr int STEP = U * kMegaConsumerWarps * TPW;
    uint4 kw[U], vw[U];
    // cached tokens do not depend on this step's projections: their K/V loads go out BEFORE the wait for q
    auto load_block = [&](int jb) {
        int page[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = jb + (u * kMegaConsumerWarps + w) * TPW + sub;
            page[u] = j < j1 ? ld_nc_early_s32(a.block_table + j / a.page_size) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = jb + (u * kMegaConsumerWarps + w) * TPW + sub;
            kw[u] = make_uint4(0, 0, 0, 0);
            vw[u] = make_uint4(0, 0, 0, 0);
            if (j < j1 && j != pos) {
                const int off = j % a.page_size;
                kw[u] = ld_cg_early(kv.at(page[u], 0, off) + kvh * HD + sl * 8);
                vw[u] = ld_cg_early(kv.at(page[u], 1, off) + kvh * HD + sl * 8);
            }
        }
    };
    load_block(j0);
    if (LL) ll_sentinel(a.ll_qkv + (kvh * group_total + g0) * HD, want, lane, a.abort_flag, 181);
    float q[GROUP][8];
#pragma unroll
    for (int g = 0; g < GROUP; g++) {
        rope_slice((kvh * group_total + g0 + g) * HD, q[g]);
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
    for (int jb = j0; jb < j1; jb += STEP) {
        if (jb != j0) load_block(jb);
        bool valid[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = jb + (u * kMegaConsumerWarps + w) * TPW + sub;
            valid[u] = j < j1;
            if (valid[u] && j == pos) {
                // the token being decoded: K/V come from this step's projection; append them (bf16)
                float kr[8];
                rope_slice(qd + kvh * HD, kr);
                float4 v0, v1;
                if (LL) {
                    float vv[8];
                    const unsigned long long* const pp[1] = {a.ll_qkv + qd + a.kvd + kvh * HD + sl * 8};
                    ll_ld8n<1>(pp, want, vv, a.abort_flag, 182);
                    v0 = make_float4(vv[0], vv[1], vv[2], vv[3]);
                    v1 = make_float4(vv[4], vv[5], vv[6], vv[7]);
                } else {
                    const float* vsrc = a.qkv + qd + a.kvd + kvh * HD + sl * 8;
                    v0 = __ldcg(reinterpret_cast<const float4*>(vsrc));
                    v1 = __ldcg(reinterpret_cast<const float4*>(vsrc + 4));
                }
                kw[u] = make_uint4(pack_bf16x2(kr[0], kr[1]), pack_bf16x2(kr[2], kr[3]), pack_bf16x2(kr[4], kr[5]), pack_bf16x2(kr[6], kr[7]));
                vw[u] = make_uint4(pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w));
                if (g0 == 0) {  // one writer per kv head
                    const int page = __ldg(a.block_table + j / a.page_size), off = j % a.page_size;
                    *reinterpret_cast<uint4*>(kv.at(page, 0, off) + kvh * HD + sl * 8) = kw[u];
                    *reinterpret_cast<uint4*>(kv.at(page, 1, off) + kvh * HD + sl * 8) = vw[u];
#ifndef MEGA_X_NOFENCE
                    if (LL) __threadfence();  // the cache line must be out before this item's partials announce the phase done
#endif
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
        if (LL) {
            ll_st(a.ll_pacc + ll_perm(static_cast<int>((pbase * group_total + g0 + g) * HD + d)), A, gp);
            if (d == 0) {
                ll_st(a.ll_pml + (pbase * group_total + g0 + g) * 2, Mx, gp);
                ll_st(a.ll_pml + (pbase * group_total + g0 + g) * 2 + 1, L, gp);
            }
        } else {
            a.part_acc[(pbase * group_total + g0 + g) * HD + d] = A;
            if (d == 0) {
                a.part_ml[(pbase * group_total + g0 + g) * 2] = Mx;
                a.part_ml[(pbase * group_total + g0 + g) * 2 + 1] = L;
            }
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

template <int HD, bool LL>
__device__ __forceinline__ void mega_attn_group(const MegaArgs& a, uint16_t* kv_pool, uint32_t scratch, int kvh, int split,
                                                int nsplit, int pos, int tid, unsigned long long* pcol, int g_only, uint32_t gp) {
    const int group = a.nh / a.nkv;
    if (g_only >= 0) {
        mega_attn_item<HD, 1, LL>(a, kv_pool, scratch, kvh, split, nsplit, pos, tid, pcol, group, g_only, gp);
        return;
    }
    switch (group) {
        case 1: mega_attn_item<HD, 1, LL>(a, kv_pool, scratch, kvh, split, nsplit, pos, tid, pcol, 1, 0, gp); break;
        case 2: mega_attn_item<HD, 2, LL>(a, kv_pool, scratch, kvh, split, nsplit, pos, tid, pcol, 2, 0, gp); break;
        case 3: mega_attn_item<HD, 3, LL>(a, kv_pool, scratch, kvh, split, nsplit, pos, tid, pcol, 3, 0, gp); break;
        case 4: mega_attn_item<HD, 4, LL>(a, kv_pool, scratch, kvh, split, nsplit, pos, tid, pcol, 4, 0, gp); break;
        default: mega_attn_item<HD, 8, LL>(a, kv_pool, scratch, kvh, split, nsplit, pos, tid, pcol, 8, 0, gp); break;
    }
}

// the producer's view of the weight stream: every chunk this CTA needs, in model order, step after step
struct ChunkCursor {
    int step, pi, row, r1, RC, K;
    const uint16_t* W;
    long long index;
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
            if (ph.type != PH_ATTN) {
                const int r0 = ph.r0;
                r1 = ph.r1;
                if (r0 < r1) {
                    row = r0;
                    RC = mega_rows_per_stage(ph.ks);
                    K = ph.K;
                    W = ph.W;
                    return;
                }
            }
            pi++;
        }
    }
    __device__ __forceinline__ void get(const uint16_t*& src, uint32_t& bytes) const {
        src = W + static_cast<size_t>(row) * K;
        bytes = static_cast<uint32_t>(min(RC, r1 - row)) * K * 2;
    }
    __device__ __forceinline__ void next(uint32_t phases, int n_phases, int n_steps) {
        index++;
        row += RC;
        if (row >= r1) {
            pi++;
            seek_phase(phases, n_phases, n_steps);
        }
    }
};

template <bool LL>
__global__ void __launch_bounds__(kMegaThreads, 1) mega_decode_kernel() {
    const MegaArgs& a = c_mega;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    MegaSmem sm;
    sm.ring = smem_u32(smem_raw);
    uint32_t p = sm.ring + static_cast<uint32_t>(a.n_stages) * kMegaStageBytes;
    sm.full = p;  p += 8 * kMegaMaxStages;
    sm.empty = p; p += 8 * kMegaMaxStages;
    sm.keys = p;  p += 8 * 8;
    sm.rel = p;   p += 4 * 16;
    sm.red = p;   p += 4 * 32;
    sm.part = p;  p += 4 * 2 * kMegaConsumerWarps * kMegaRows;
    sm.xs = p;    p += 4 * kMegaXsFloats;
    sm.nw = p;    p += 2 * static_cast<uint32_t>(a.H);
    sm.phases = p; p += (48 + 8) * static_cast<uint32_t>(a.n_phases);   // descriptors, then this CTA's (r0, r1) per phase
    p = (p + 15u) & ~15u;
    sm.attn_scratch = p;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < a.n_stages; s++) {
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(sm.rel + s * 4), "r"(0u) : "memory");
            mbar_init(sm.full + s * 8, 1);
            mbar_init(sm.empty + s * 8, kMegaRows);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
        sts32f(sm.red, __int_as_float(*reinterpret_cast<const volatile int32_t*>(a.token)));
        sts32f(sm.red + 4, __int_as_float(*reinterpret_cast<const volatile int32_t*>(a.position)));
    }
    __syncthreads();

    const int pos0 = a.arg_io ? a.pos0 : __float_as_int(lds32f(sm.red + 4));
    const int token0 = a.arg_io ? a.token0 : __float_as_int(lds32f(sm.red));

    if (tid >= kMegaConsumerThreads) {
        // ================= producer warp: stream every weight chunk this CTA will ever need =================
        if (tid == kMegaConsumerThreads) {
            const uint64_t policy = l2_evict_first_policy();
            ChunkCursor ld{0, 0, 0, 0, 1, 0, nullptr, 0}, pf{0, 0, 0, 0, 1, 0, nullptr, 0};
            ld.seek_phase(sm.phases, a.n_phases, a.n_steps);
            pf.seek_phase(sm.phases, a.n_phases, a.n_steps);
            RingPos rp{0, 0, 0};
            RingPos landed{0, 0, 0};  // oldest copy not yet known to have landed
            int outstanding = 0;
            while (!ld.done(a.n_steps)) {
                // optional: HBM -> L2 runs l2_ahead chunks ahead of the ring
                while (a.l2_ahead > 0 && !pf.done(a.n_steps) && pf.index < ld.index + a.n_stages + a.l2_ahead) {
                    if (pf.index >= ld.index + a.n_stages) {
                        const uint16_t* src;
                        uint32_t bytes;
                        pf.get(src, bytes);
                        tma_prefetch_l2(src, bytes);
                    }
                    pf.next(sm.phases, a.n_phases, a.n_steps);
                }
                // optional cap on copies issued but not yet landed
                if (a.max_inflight > 0) {
                    while (outstanding > 0 && mbar_test_wait(sm.full + landed.stage * 8, landed.parity)) {
                        landed.advance(a.n_stages);
                        outstanding--;
                    }
                    if (outstanding >= a.max_inflight) {
                        mbar_wait(sm.full + landed.stage * 8, landed.parity, a.abort_flag, 310);
                        landed.advance(a.n_stages);
                        outstanding--;
                    }
                }
                if (!mbar_test_wait(sm.empty + rp.stage * 8, rp.parity ^ 1)) {
                    // ring full: sleep between probes so the spinning producer does not steal issue slots
                    // from the two consumer warps that share its scheduler
                    unsigned spins = 0;
                    while (!mbar_test_wait(sm.empty + rp.stage * 8, rp.parity ^ 1)) {
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
                outstanding++;
                ld.next(sm.phases, a.n_phases, a.n_steps);
            }
        }
        return;
    }

    // ================= consumer warps =================
    ConsumerState st;
    st.rp = RingPos{0, 0, 0};
    st.nbar = 0;
    st.epoch = *a.bar_epoch;
    st.best_key = 0ull;
    st.token = token0;
    st.step = 0;
    st.need_barrier = false;
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
                if (!LL) mega_phase_barrier(a, st, false, tid);
                const uint32_t gp = a.seq_base + static_cast<uint32_t>(step * a.n_phases + pi) + 1u;
                if (prof) a.prof[(prow + 1) * pstride + pi] = a.prof[(prow + 2) * pstride + pi] = globaltimer_ns();
                const AttnPlan plan = mega_attn_plan(a.nsplit_max, pos + 1, a.attn_tps, a.nh, gridDim.x);
                const int nsplit = plan.nsplit, item = blockIdx.x;
                const int n_items = (plan.per_q_head ? a.nh : a.nkv) * nsplit;
                if (item < n_items) {
                    const int unit = item / nsplit, split = item % nsplit;   // query head or kv head
                    const int group = a.nh / a.nkv;
                    const int kvh = plan.per_q_head ? unit / group : unit;
                    const int g_only = plan.per_q_head ? unit % group : -1;
                    unsigned long long* pcol = prof && blockIdx.x == 0 ? a.prof + pi : nullptr;
                    if (a.hd == 64) mega_attn_group<64, LL>(a, ph.kv_pool, sm.attn_scratch, kvh, split, nsplit, pos, tid, pcol, g_only, gp);
                    else if (a.hd == 128) mega_attn_group<128, LL>(a, ph.kv_pool, sm.attn_scratch, kvh, split, nsplit, pos, tid, pcol, g_only, gp);
                    else mega_attn_group<32, LL>(a, ph.kv_pool, sm.attn_scratch, kvh, split, nsplit, pos, tid, pcol, g_only, gp);
                }
                if (prof) a.prof[(prow + 3) * pstride + pi] = globaltimer_ns();
                if (a.prof && step == a.n_steps - 1 && tid == 0 && blockIdx.x < 160)
                    a.prof[(16 + 160 + blockIdx.x) * pstride + pi] = item < n_items ? globaltimer_ns() : 0ull;
            } else {
#ifdef MEGA_ONLY_1B
                if (ph.ks == 1) mega_gemv_phase<8, false, LL>(a, ph, sm, st, pi, pos, tid);
                else mega_gemv_phase<8, true, LL>(a, ph, sm, st, pi, pos, tid);
#else
                if (ph.ks == 1) {
                    switch (ph.m) {
                        case 1: mega_gemv_phase<1, false, LL>(a, ph, sm, st, pi, pos, tid); break;
                        case 2: mega_gemv_phase<2, false, LL>(a, ph, sm, st, pi, pos, tid); break;
                        case 3: mega_gemv_phase<3, false, LL>(a, ph, sm, st, pi, pos, tid); break;
                        case 4: mega_gemv_phase<4, false, LL>(a, ph, sm, st, pi, pos, tid); break;
                        case 5: mega_gemv_phase<5, false, LL>(a, ph, sm, st, pi, pos, tid); break;
                        case 6: mega_gemv_phase<6, false, LL>(a, ph, sm, st, pi, pos, tid); break;
                        case 7: mega_gemv_phase<7, false, LL>(a, ph, sm, st, pi, pos, tid); break;
                        default: mega_gemv_phase<8, false, LL>(a, ph, sm, st, pi, pos, tid); break;
                    }
                } else {
                    switch (ph.m) {
                        case 1: mega_gemv_phase<1, true, LL>(a, ph, sm, st, pi, pos, tid); break;
                        case 2: mega_gemv_phase<2, true, LL>(a, ph, sm, st, pi, pos, tid); break;
                        case 3: mega_gemv_phase<3, true, LL>(a, ph, sm, st, pi, pos, tid); break;
                        case 4: mega_gemv_phase<4, true, LL>(a, ph, sm, st, pi, pos, tid); break;
                        case 5: mega_gemv_phase<5, true, LL>(a, ph, sm, st, pi, pos, tid); break;
                        case 6: mega_gemv_phase<6, true, LL>(a, ph, sm, st, pi, pos, tid); break;
                        case 7: mega_gemv_phase<7, true, LL>(a, ph, sm, st, pi, pos, tid); break;
                        default: mega_gemv_phase<8, true, LL>(a, ph, sm, st, pi, pos, tid); break;
                    }
                }
#endif
            }
#if MEGA_TP
            if (ph.type == PH_OPROJ || ph.type == PH_DOWN)   // collect the tp partial sums of this CTA's rows, publish h locally
                mega_tp_collect(ph, a.seq_base + static_cast<uint32_t>(step * a.n_phases + pi) + 1u, st.token, tid);
#endif
            if (ph.type == PH_LMHEAD) {
                // CTA-level argmax, then one atomicMax per CTA on this step's key
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
                    if (LL) {
                        const uint32_t gp = a.seq_base + static_cast<uint32_t>(step * a.n_phases + pi) + 1u;
#ifdef MEGA_X_NOFENCE
                        // the K/V cache lines this CTA appended during the token (plain stores, ordered before this thread by the
                        // CTA barriers since) must be visible before the key announces the token done: readers fence after the keys
                        __threadfence();
#endif
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
                    } else {
                        atomicMax(a.argmax_keys + (step % 3), k);
                    }
                }
            }
            // the key two steps ahead was last read at the start of the previous step: safe to clear now
            if (!LL && pi == 2 && blockIdx.x == 0 && tid == 0) a.argmax_keys[(step + 1) % 3] = 0ull;
        }
    }
    if (LL) {
        // only CTA 0 needs the last token: it waits for every CTA's key of the last lm_head phase
        if (blockIdx.x != 0) return;
        const int token = mega_poll_token(sm, a.seq_base + static_cast<uint32_t>(a.n_steps * a.n_phases), tid);
        if (tid == 0) {
            a.out_ids[a.n_steps - 1] = token;
            *a.token = token;
            *a.position = pos0 + a.n_steps;
        }
        return;
    }
    // final barrier: every CTA's lm_head rows are in the last key
    st.nbar++;
    mega_grid_sync(a, st.epoch + st.nbar * gridDim.x, tid);
    const int token = argmax_key_index(ld_acquire_u64(a.argmax_keys + ((a.n_steps - 1) % 3)));
    if (blockIdx.x == 0 && tid == 0) {
        a.out_ids[a.n_steps - 1] = token;
        *a.token = token;
        *a.position = pos0 + a.n_steps;
        *a.bar_epoch = st.epoch + st.nbar * gridDim.x;
    }
}

}  // namespace MEGA_NS
}  // namespace b2l

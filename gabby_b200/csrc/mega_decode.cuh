// mega_decode.cuh -- persistent batch-1 decode megakernel: ONE cooperative launch runs n_steps
// whole tokens. The B200-native answer to "a 0.4 ms token is 100 dependent tiny ops":
//
//   * one CTA per SM (148), 8 consumer warps + 1 producer warp
//   * the producer streams this CTA's share of EVERY weight matrix, in model order, through a
//     6-stage x 32 KB shared-memory ring with TMA bulk copies (cp.async.bulk -> UBLKCP) that
//     complete on mbarriers. Weights do not depend on activations, so the stream never stops:
//     it runs ahead across op, layer and even token boundaries, bounded only by the ring.
//   * consumers keep the op's input vector in REGISTERS (64 fp32 per lane), dot it against the
//     bf16 rows as they land, and hand results on through L2; ops are separated by a grid-wide
//     barrier (one release-atomic per CTA) whose latency the ring hides from HBM.
//   * attention (split-K over the paged bf16 cache, RoPE and K/V append fused in), SwiGLU,
//     residual adds, both RMSNorms, the lm_head argmax and the token feedback are all inside.
//
// HBM sees one sequential read of the model per token; everything else lives in L2 / smem.
// Math is identical to decode_kernels.cuh (the multi-kernel path) up to fp32 summation order.
#pragma once
#include "common.cuh"
#include "decode_kernels.cuh"

namespace b2l {

constexpr int kMegaConsumerWarps = 8;
constexpr int kMegaConsumerThreads = kMegaConsumerWarps * 32;
constexpr int kMegaThreads = kMegaConsumerThreads + 32;  // + producer warp
constexpr int kMegaStageBytes = 32 * 1024;
constexpr int kMegaMaxStages = 6;
constexpr int kMegaXsFloats = 2048;

enum MegaPhaseType { PH_QKV = 0, PH_ATTN = 1, PH_OPROJ = 2, PH_GATEUP = 3, PH_DOWN = 4, PH_LMHEAD = 5 };

struct MegaPhase {
    int type, layer;
    const uint16_t* W;       // [N][K] bf16 (PH_ATTN: unused)
    const uint16_t* norm_w;  // fused RMSNorm weight or null
    uint16_t* kv_pool;       // PH_ATTN: this layer's KV pool
    int N, K;
    int ks;                  // warps per row (K split); rows per chunk = 8 / ks
    int m;                   // 16-byte sweeps per warp unit: slice = 256 * m elements
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
    int32_t* token;      // in: first token; out: last argmax
    int32_t* position;   // in: first position; out: advanced
    int32_t* out_ids;    // [n_steps]
    int n_steps;
    // sync
    unsigned long long* bar_counter;  // monotonically increasing arrivals
    unsigned long long* bar_epoch;    // arrivals consumed by previous launches
    unsigned long long* argmax_keys;  // [3]
    int* abort_flag;
};

// ---- PTX helpers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// a deadlock here would hang the GPU: bound every wait (~2 s) and trap with a reason code instead
__device__ __noinline__ void mega_die(int* abort_flag, int code) {
    atomicExch(abort_flag, code);
    __threadfence_system();
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* abort_flag, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) mega_die(abort_flag, code);
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kMegaConsumerThreads) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// rows [r0, r1) of an N-row matrix owned by CTA `c` of `G` (unit = 2 rows for SwiGLU pairs)
__device__ __forceinline__ void mega_row_range(int N, int unit, int c, int G, int& r0, int& r1) {
    const long long units = N / unit;
    r0 = static_cast<int>(units * c / G) * unit;
    r1 = static_cast<int>(units * (c + 1) / G) * unit;
}

// grid-wide barrier among the consumer threads of all CTAs: arrive (release) then wait for `target`
__device__ __forceinline__ void mega_grid_sync(const MegaArgs& a, unsigned long long target, int tid) {
    consumer_bar();  // all of this CTA's writes are ordered before thread 0's release below
    if (tid == 0) {
        __threadfence();
        atomicAdd(a.bar_counter, 1ull);
        const long long t0 = clock64();
        while (ld_acquire_u64(a.bar_counter) < target) {
            if (clock64() - t0 > 4000000000ll) mega_die(a.abort_flag, 100);
        }
        __threadfence();
    }
    consumer_bar();
}

// ---- one GEMV-type phase for one CTA ---------------------------------------------------------
struct MegaSmem {
    uint8_t* ring;       // [n_stages][kMegaStageBytes]
    uint64_t* full;      // [n_stages]
    uint64_t* empty;     // [n_stages]
    float* xs;           // [kMegaXsFloats]
    float* red;          // [32]
    float* part;         // [2][8] per-chunk partial sums (double buffered)
    unsigned long long* keys;  // [8]
    float* attn_scratch;
};

template <int M>
__device__ __noinline__ void mega_gemv_phase(const MegaArgs& a, const MegaPhase& ph, const MegaSmem& sm, uint32_t& cc,
                                             int token, unsigned long long& best_key, int tid) {
    const int lane = tid & 31, w = tid >> 5;
    const int ks = ph.ks, RC = 8 / ks, slice = 256 * M;
    const int q = w % ks, rloc = w / ks;
    const int K = ph.K;
    // ---- input vector -> registers (fused RMSNorm) ----
    const float* xsrc = ph.type == PH_OPROJ ? a.attn : ph.type == PH_DOWN ? a.act : a.h;
    const bool from_embed = (ph.type == PH_QKV && ph.layer == 0);
    float xr[M * 8];
    if (ks == 1) {
        // every warp needs the same K floats: fetch once per CTA, then fan out through smem
        for (int k = tid * 4; k < K; k += kMegaConsumerThreads * 4) {
            float4 v;
            if (from_embed) {
                const uint2 e = *reinterpret_cast<const uint2*>(a.embed + static_cast<size_t>(token) * a.H + k);
                v = make_float4(bf16lo(e.x), bf16hi(e.x), bf16lo(e.y), bf16hi(e.y));
            } else {
                v = __ldcg(reinterpret_cast<const float4*>(xsrc + k));
            }
            *reinterpret_cast<float4*>(sm.xs + k) = v;
        }
        consumer_bar();
#pragma unroll
        for (int i = 0; i < M; i++) {
            const float4 v0 = *reinterpret_cast<const float4*>(sm.xs + i * 256 + lane * 8);
            const float4 v1 = *reinterpret_cast<const float4*>(sm.xs + i * 256 + lane * 8 + 4);
            xr[i * 8 + 0] = v0.x; xr[i * 8 + 1] = v0.y; xr[i * 8 + 2] = v0.z; xr[i * 8 + 3] = v0.w;
            xr[i * 8 + 4] = v1.x; xr[i * 8 + 5] = v1.y; xr[i * 8 + 6] = v1.z; xr[i * 8 + 7] = v1.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < M; i++) {
            const int k = q * slice + i * 256 + lane * 8;
            float4 v0, v1;
            if (from_embed) {
                const uint4 e = *reinterpret_cast<const uint4*>(a.embed + static_cast<size_t>(token) * a.H + k);
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
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < M * 8; i++) ss = fmaf(xr[i], xr[i], ss);
        ss = warp_sum(ss);
        if (lane == 0) sm.red[w] = ss;
        consumer_bar();
        float tot = 0.f;
        for (int i = 0; i < ks; i++) tot += sm.red[i];  // warps 0..ks-1 hold slices 0..ks-1
        const float inv = rsqrtf(tot / static_cast<float>(K) + a.eps);
#pragma unroll
        for (int i = 0; i < M; i++) {
            const uint4 nw = *reinterpret_cast<const uint4*>(ph.norm_w + q * slice + i * 256 + lane * 8);
            xr[i * 8 + 0] = bf16lo(nw.x) * (xr[i * 8 + 0] * inv);
            xr[i * 8 + 1] = bf16hi(nw.x) * (xr[i * 8 + 1] * inv);
            xr[i * 8 + 2] = bf16lo(nw.y) * (xr[i * 8 + 2] * inv);
            xr[i * 8 + 3] = bf16hi(nw.y) * (xr[i * 8 + 3] * inv);
            xr[i * 8 + 4] = bf16lo(nw.z) * (xr[i * 8 + 4] * inv);
            xr[i * 8 + 5] = bf16hi(nw.z) * (xr[i * 8 + 5] * inv);
            xr[i * 8 + 6] = bf16lo(nw.w) * (xr[i * 8 + 6] * inv);
            xr[i * 8 + 7] = bf16hi(nw.w) * (xr[i * 8 + 7] * inv);
        }
    }

    // ---- stream this CTA's rows ----
    int r0, r1;
    mega_row_range(ph.N, ph.type == PH_GATEUP ? 2 : 1, blockIdx.x, gridDim.x, r0, r1);
    const int n_chunks = (r1 - r0 + RC - 1) / RC;
    const bool cross = (ks > 1) || ph.type == PH_GATEUP;  // result needs more than one warp
    for (int ch = 0; ch < n_chunks; ch++, cc++) {
        const int stage = cc % a.n_stages;
        const uint32_t parity = (cc / a.n_stages) & 1;
        const int row = r0 + ch * RC + rloc;
        const bool valid = row < r1;
        mbar_wait(&sm.full[stage], parity, a.abort_flag, 200 + ph.type);
        float acc0 = 0.f, acc1 = 0.f;
        if (valid) {
            const uint8_t* base = sm.ring + static_cast<size_t>(stage) * kMegaStageBytes +
                                  (static_cast<size_t>(rloc) * K + static_cast<size_t>(q) * slice) * 2 + lane * 16;
#pragma unroll
            for (int i = 0; i < M; i++) {
                const uint4 wv = *reinterpret_cast<const uint4*>(base + i * 512);
                float& acc = (i & 1) ? acc1 : acc0;
                acc = fmaf(bf16lo(wv.x), xr[i * 8 + 0], acc);
                acc = fmaf(bf16hi(wv.x), xr[i * 8 + 1], acc);
                acc = fmaf(bf16lo(wv.y), xr[i * 8 + 2], acc);
                acc = fmaf(bf16hi(wv.y), xr[i * 8 + 3], acc);
                acc = fmaf(bf16lo(wv.z), xr[i * 8 + 4], acc);
                acc = fmaf(bf16hi(wv.z), xr[i * 8 + 5], acc);
                acc = fmaf(bf16lo(wv.w), xr[i * 8 + 6], acc);
                acc = fmaf(bf16hi(wv.w), xr[i * 8 + 7], acc);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[stage]);  // smem slot can be refilled
        float s = warp_sum(acc0 + acc1);
        if (cross) {
            float* part = sm.part + (ch & 1) * 8;
            if (lane == 0) part[w] = s;
            consumer_bar();
            if (q != 0 || (ph.type == PH_GATEUP && (rloc & 1))) continue;  // one finalising warp per row / pair
            s = 0.f;
            for (int i = 0; i < ks; i++) s += part[w + i];
            if (ph.type == PH_GATEUP) {
                float u = 0.f;
                for (int i = 0; i < ks; i++) u += part[w + ks + i];
                s = (s / (1.0f + __expf(-s))) * u;   // silu(gate) * up
            }
        }
        if (lane == 0 && valid) {
            switch (ph.type) {
                case PH_QKV: a.qkv[row] = s; break;
                case PH_GATEUP: a.act[row >> 1] = s; break;
                case PH_OPROJ:
                    if (ph.layer == 0)
                        a.h[row] = bf16_bits_to_f32(a.embed[static_cast<size_t>(token) * a.H + row]) + s;
                    else
                        a.h[row] = __ldcg(a.h + row) + s;
                    break;
                case PH_DOWN: a.h[row] = __ldcg(a.h + row) + s; break;
                default: {  // PH_LMHEAD
                    a.logits[row] = s;
                    const unsigned long long key = argmax_key(s, row);
                    best_key = key > best_key ? key : best_key;
                }
            }
        }
    }
}

// ---- attention work item: (kv head, split) ------------------------------------------------------
template <int HD, int GROUP>
__device__ __noinline__ void mega_attn_item(const MegaArgs& a, const MegaPhase& ph, const MegaSmem& sm, int kvh, int split, int nsplit,
                               int pos, int tid) {
    constexpr int LPT = HD / 8, TPW = 32 / LPT, HALF = HD / 2;
    const int lane = tid & 31, w = tid >> 5, sub = lane / LPT, sl = lane % LPT;
    const KvLayout kv{ph.kv_pool, a.page_size, a.kvd};
    const int ctx = pos + 1;
    const int chunk = (ctx + nsplit - 1) / nsplit;
    const int j0 = split * chunk, j1 = min(ctx, j0 + chunk);
    const float* cs = a.rope + static_cast<size_t>(pos) * HD;  // [HALF][2]
    const int qd = a.nh * HD;

    // rotate-half RoPE of one 8-wide slice of a head living in the fused qkv row
    auto rope_slice = [&](const float* head, float* out) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int d = sl * 8 + i;
            const int j = d < HALF ? d : d - HALF;
            const float c = cs[2 * j], s = cs[2 * j + 1];
            const float x0 = __ldcg(head + j), x1 = __ldcg(head + j + HALF);
            out[i] = d < HALF ? x0 * c - x1 * s : x1 * c + x0 * s;
        }
    };
    float q[GROUP][8];
#pragma unroll
    for (int g = 0; g < GROUP; g++) {
        rope_slice(a.qkv + (kvh * GROUP + g) * HD, q[g]);
#pragma unroll
        for (int i = 0; i < 8; i++) q[g][i] *= a.attn_scale;
    }
    float m[GROUP], l[GROUP], acc[GROUP][8];
#pragma unroll
    for (int g = 0; g < GROUP; g++) {
        m[g] = -INFINITY;
        l[g] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) acc[g][i] = 0.f;
    }
    for (int jb = j0; jb < j1; jb += kMegaConsumerWarps * TPW) {
        const int j = jb + w * TPW + sub;
        const bool valid = j < j1;
        float kf[8], vf[8];
#pragma unroll
        for (int i = 0; i < 8; i++) kf[i] = vf[i] = 0.f;
        if (valid) {
            const int page = a.block_table[j / a.page_size], off = j % a.page_size;
            uint16_t* kp = kv.at(page, 0, off) + kvh * HD + sl * 8;
            uint16_t* vp = kv.at(page, 1, off) + kvh * HD + sl * 8;
            uint4 kw, vw;
            if (j == pos) {
                // the token being decoded: K/V come from this step's projection; append them (bf16)
                float kr[8];
                rope_slice(a.qkv + qd + kvh * HD, kr);
                const float* vsrc = a.qkv + qd + a.kvd + kvh * HD + sl * 8;
                const float4 v0 = __ldcg(reinterpret_cast<const float4*>(vsrc)), v1 = __ldcg(reinterpret_cast<const float4*>(vsrc + 4));
                kw = make_uint4(pack_bf16x2(kr[0], kr[1]), pack_bf16x2(kr[2], kr[3]), pack_bf16x2(kr[4], kr[5]), pack_bf16x2(kr[6], kr[7]));
                vw = make_uint4(pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w));
                *reinterpret_cast<uint4*>(kp) = kw;
                *reinterpret_cast<uint4*>(vp) = vw;
            } else {
                kw = __ldcg(reinterpret_cast<const uint4*>(kp));
                vw = __ldcg(reinterpret_cast<const uint4*>(vp));
            }
            kf[0] = bf16lo(kw.x); kf[1] = bf16hi(kw.x); kf[2] = bf16lo(kw.y); kf[3] = bf16hi(kw.y);
            kf[4] = bf16lo(kw.z); kf[5] = bf16hi(kw.z); kf[6] = bf16lo(kw.w); kf[7] = bf16hi(kw.w);
            vf[0] = bf16lo(vw.x); vf[1] = bf16hi(vw.x); vf[2] = bf16lo(vw.y); vf[3] = bf16hi(vw.y);
            vf[4] = bf16lo(vw.z); vf[5] = bf16hi(vw.z); vf[6] = bf16lo(vw.w); vf[7] = bf16hi(vw.w);
        }
#pragma unroll
        for (int g = 0; g < GROUP; g++) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) s = fmaf(q[g][i], kf[i], s);
#pragma unroll
            for (int o = LPT / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (valid) {
                const float mn = fmaxf(m[g], s);
                const float corr = __expf(m[g] - mn), p = __expf(s - mn);
                l[g] = l[g] * corr + p;
#pragma unroll
                for (int i = 0; i < 8; i++) acc[g][i] = fmaf(acc[g][i], corr, p * vf[i]);
                m[g] = mn;
            }
        }
    }
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
    // per-warp results -> smem: [w][g][HD] then (m, l)
    float* s_acc = sm.attn_scratch;                                   // [8][GROUP][HD]
    float* s_ml = sm.attn_scratch + kMegaConsumerWarps * GROUP * HD;  // [8][GROUP][2]
    if (sub == 0) {
#pragma unroll
        for (int g = 0; g < GROUP; g++) {
#pragma unroll
            for (int i = 0; i < 8; i++) s_acc[(w * GROUP + g) * HD + sl * 8 + i] = acc[g][i];
            if (sl == 0) {
                s_ml[(w * GROUP + g) * 2] = m[g];
                s_ml[(w * GROUP + g) * 2 + 1] = l[g];
            }
        }
    }
    consumer_bar();
    const size_t pbase = static_cast<size_t>(kvh) * a.nsplit_max + split;
    for (int e = tid; e < GROUP * HD; e += kMegaConsumerThreads) {
        const int g = e / HD, d = e % HD;
        float M = -INFINITY;
#pragma unroll
        for (int t = 0; t < kMegaConsumerWarps; t++) M = fmaxf(M, s_ml[(t * GROUP + g) * 2]);
        float L = 0.f, A = 0.f;
        if (M > -INFINITY) {
#pragma unroll
            for (int t = 0; t < kMegaConsumerWarps; t++) {
                const float wgt = __expf(s_ml[(t * GROUP + g) * 2] - M);
                L = fmaf(s_ml[(t * GROUP + g) * 2 + 1], wgt, L);
                A = fmaf(s_acc[(t * GROUP + g) * HD + d], wgt, A);
            }
        }
        if (nsplit == 1) {
            a.attn[(kvh * GROUP + g) * HD + d] = A / L;
        } else {
            a.part_acc[(pbase * GROUP + g) * HD + d] = A;
            if (d == 0) {
                a.part_ml[(pbase * GROUP + g) * 2] = M;
                a.part_ml[(pbase * GROUP + g) * 2 + 1] = L;
            }
        }
    }
    if (nsplit == 1) return;
    // last-arriving split of this kv head merges them
    __threadfence();
    consumer_bar();
    int* s_flag = reinterpret_cast<int*>(sm.red + 16);
    if (tid == 0) {
        const int done = atomicAdd(a.attn_counters + kvh, 1);
        *s_flag = (done == nsplit - 1);
        if (done == nsplit - 1) a.attn_counters[kvh] = 0;
    }
    consumer_bar();
    if (!*s_flag) return;
    __threadfence();
    const size_t rbase = static_cast<size_t>(kvh) * a.nsplit_max;
    for (int e = tid; e < GROUP * HD; e += kMegaConsumerThreads) {
        const int g = e / HD, d = e % HD;
        float M = -INFINITY;
        for (int s = 0; s < nsplit; s++) M = fmaxf(M, __ldcg(a.part_ml + ((rbase + s) * GROUP + g) * 2));
        float L = 0.f, A = 0.f;
        for (int s = 0; s < nsplit; s++) {
            const float ms = __ldcg(a.part_ml + ((rbase + s) * GROUP + g) * 2);
            if (ms == -INFINITY) continue;
            const float wgt = __expf(ms - M);
            L = fmaf(__ldcg(a.part_ml + ((rbase + s) * GROUP + g) * 2 + 1), wgt, L);
            A = fmaf(__ldcg(a.part_acc + ((rbase + s) * GROUP + g) * HD + d), wgt, A);
        }
        a.attn[(kvh * GROUP + g) * HD + d] = A / L;
    }
}

template <int HD>
__device__ __forceinline__ void mega_attn_group(const MegaArgs& a, const MegaPhase& ph, const MegaSmem& sm, int kvh, int split,
                                                int nsplit, int pos, int tid) {
    switch (a.nh / a.nkv) {
        case 1: mega_attn_item<HD, 1>(a, ph, sm, kvh, split, nsplit, pos, tid); break;
        case 2: mega_attn_item<HD, 2>(a, ph, sm, kvh, split, nsplit, pos, tid); break;
        case 3: mega_attn_item<HD, 3>(a, ph, sm, kvh, split, nsplit, pos, tid); break;
        case 4: mega_attn_item<HD, 4>(a, ph, sm, kvh, split, nsplit, pos, tid); break;
        default: mega_attn_item<HD, 8>(a, ph, sm, kvh, split, nsplit, pos, tid); break;
    }
}

__global__ void __launch_bounds__(kMegaThreads, 1) mega_decode_kernel(const MegaArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    MegaSmem sm;
    sm.ring = smem_raw;
    uint8_t* p = smem_raw + static_cast<size_t>(a.n_stages) * kMegaStageBytes;
    sm.full = reinterpret_cast<uint64_t*>(p);  p += 8 * kMegaMaxStages;
    sm.empty = reinterpret_cast<uint64_t*>(p); p += 8 * kMegaMaxStages;
    sm.keys = reinterpret_cast<unsigned long long*>(p); p += 8 * 8;
    sm.red = reinterpret_cast<float*>(p);      p += 4 * 32;
    sm.part = reinterpret_cast<float*>(p);     p += 4 * 16;
    sm.xs = reinterpret_cast<float*>(p);       p += 4 * kMegaXsFloats;
    sm.attn_scratch = reinterpret_cast<float*>(p);

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < a.n_stages; s++) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], kMegaConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int G = gridDim.x;
    const unsigned long long epoch = *a.bar_epoch;
    const int pos0 = *a.position;
    const int token0 = *a.token;

    if (tid >= kMegaConsumerThreads) {
        // ================= producer warp: stream every weight chunk this CTA will ever need =================
        if (tid == kMegaConsumerThreads) {
            const uint64_t policy = l2_evict_first_policy();
            uint32_t cc = 0;
            for (int step = 0; step < a.n_steps; step++) {
                for (int pi = 0; pi < a.n_phases; pi++) {
                    const MegaPhase& ph = a.phases[pi];
                    if (ph.type == PH_ATTN) continue;
                    const int RC = 8 / ph.ks;
                    int r0, r1;
                    mega_row_range(ph.N, ph.type == PH_GATEUP ? 2 : 1, blockIdx.x, G, r0, r1);
                    for (int row = r0; row < r1; row += RC, cc++) {
                        const int stage = cc % a.n_stages;
                        mbar_wait(&sm.empty[stage], ((cc / a.n_stages) & 1) ^ 1, a.abort_flag, 300 + ph.type);
                        const uint32_t bytes = static_cast<uint32_t>(min(RC, r1 - row)) * ph.K * 2;
                        mbar_arrive_expect_tx(&sm.full[stage], bytes);
                        tma_bulk_g2s(sm.ring + static_cast<size_t>(stage) * kMegaStageBytes, ph.W + static_cast<size_t>(row) * ph.K,
                                     bytes, &sm.full[stage], policy);
                    }
                }
            }
        }
        return;
    }

    // ================= consumer warps =================
    uint32_t cc = 0;
    unsigned long long nbar = 0;  // grid barriers passed in this launch
    int token = token0;
    for (int step = 0; step < a.n_steps; step++) {
        const int pos = pos0 + step;
        unsigned long long best_key = 0ull;
        for (int pi = 0; pi < a.n_phases; pi++) {
            const MegaPhase& ph = a.phases[pi];
            if (ph.type == PH_ATTN) {
                const int ctx = pos + 1;
                const int nsplit = max(1, min(a.nsplit_max, (ctx + 63) / 64));
                const int item = blockIdx.x;
                if (item < a.nkv * nsplit) {
                    const int kvh = item / nsplit, split = item % nsplit;
                    if (a.hd == 64) mega_attn_group<64>(a, ph, sm, kvh, split, nsplit, pos, tid);
                    else if (a.hd == 128) mega_attn_group<128>(a, ph, sm, kvh, split, nsplit, pos, tid);
                    else mega_attn_group<32>(a, ph, sm, kvh, split, nsplit, pos, tid);
                }
            } else {
                switch (ph.m) {
                    case 1: mega_gemv_phase<1>(a, ph, sm, cc, token, best_key, tid); break;
                    case 2: mega_gemv_phase<2>(a, ph, sm, cc, token, best_key, tid); break;
                    case 3: mega_gemv_phase<3>(a, ph, sm, cc, token, best_key, tid); break;
                    case 4: mega_gemv_phase<4>(a, ph, sm, cc, token, best_key, tid); break;
                    case 5: mega_gemv_phase<5>(a, ph, sm, cc, token, best_key, tid); break;
                    case 6: mega_gemv_phase<6>(a, ph, sm, cc, token, best_key, tid); break;
                    case 7: mega_gemv_phase<7>(a, ph, sm, cc, token, best_key, tid); break;
                    default: mega_gemv_phase<8>(a, ph, sm, cc, token, best_key, tid); break;
                }
            }
            if (ph.type == PH_LMHEAD) {
                // CTA-level argmax, then one atomicMax per CTA on this step's key
                const int lane = tid & 31, w = tid >> 5;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best_key, o);
                    best_key = other > best_key ? other : best_key;
                }
                if (lane == 0) sm.keys[w] = best_key;
                consumer_bar();
                if (tid == 0) {
                    unsigned long long k = sm.keys[0];
                    for (int i = 1; i < kMegaConsumerWarps; i++) k = sm.keys[i] > k ? sm.keys[i] : k;
                    atomicMax(a.argmax_keys + (step % 3), k);
                }
            }
            nbar++;
            mega_grid_sync(a, epoch + nbar * G, tid);
            if (pi == 1 && blockIdx.x == 0 && tid == 0) a.argmax_keys[(step + 1) % 3] = 0ull;  // safe: two barriers past its last reader
        }
        token = argmax_key_index(ld_acquire_u64(a.argmax_keys + (step % 3)));
        if (blockIdx.x == 0 && tid == 0) a.out_ids[step] = token;
    }
    if (blockIdx.x == 0 && tid == 0) {
        *a.token = token;
        *a.position = pos0 + a.n_steps;
        *a.bar_epoch = epoch + nbar * G;
    }
}

}  // namespace b2l

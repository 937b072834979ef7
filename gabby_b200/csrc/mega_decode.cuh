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
    unsigned long long* prof;  // optional [4][n_phases + 1] globaltimer ns of the LAST step (CTA 0 / CTA G-1: phase end, wait end)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- PTX helpers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// a deadlock here would hang the GPU: bound every wait (~2 s) and trap with a reason code instead
__device__ __noinline__ void mega_die(int* abort_flag, int code) {
    atomicExch(abort_flag, code);
    __threadfence_system();
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* abort_flag, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) mega_die(abort_flag, code);
    }
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ void sts128f(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
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
        // release: cumulative over the CTA's writes ordered by the bar.sync above; readers use ld.global.cg
        red_release_add_u64(a.bar_counter, 1ull);
        if (ld_acquire_u64(a.bar_counter) < target) {
            const long long t0 = clock64();
            while (ld_acquire_u64(a.bar_counter) < target) {
                if (clock64() - t0 > 4000000000ll) mega_die(a.abort_flag, 100);
            }
        }
    }
    consumer_bar();
}

// ---- shared memory map (32-bit shared-space addresses, kept in registers) ---------------------
struct MegaSmem {
    uint32_t ring;          // [n_stages][kMegaStageBytes]
    uint32_t full, empty;   // [n_stages] mbarriers each
    uint32_t xs;            // [kMegaXsFloats] fp32
    uint32_t red;           // [32] fp32
    uint32_t part;          // [2][8] fp32 per-chunk partial sums (double buffered)
    uint32_t keys;          // [8] u64
    uint32_t attn_scratch;
    float* gen;             // generic pointer to the same block's base (for the few generic accesses)
    uint32_t base;
    __device__ __forceinline__ float* generic(uint32_t addr) const {
        return reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(gen) + (addr - base));
    }
};

// position in the ring: stage index and phase parity, advanced incrementally (no div/mod per chunk)
struct RingPos {
    int stage;
    uint32_t parity;
    int n_stages;
    __device__ __forceinline__ void advance() {
        if (++stage == n_stages) {
            stage = 0;
            parity ^= 1;
        }
    }
};

__device__ __forceinline__ int mega_nsplit(const MegaArgs& a, int ctx) { return max(1, min(a.nsplit_max, (ctx + 127) >> 7)); }

// attention output element k (head-major) of this step: merge the split-K partials
__device__ __forceinline__ void mega_attn_combine8(const MegaArgs& a, int k, int nsplit, float* out) {
    const int head = k / a.hd, d = k % a.hd;  // 8 consecutive k never straddle a head (hd % 8 == 0)
    const int group = a.nh / a.nkv, kvh = head / group, g = head % group;
    const size_t rbase = static_cast<size_t>(kvh) * a.nsplit_max;
    float M = -INFINITY;
    for (int s = 0; s < nsplit; s++) M = fmaxf(M, __ldcg(a.part_ml + ((rbase + s) * group + g) * 2));
    float L = 0.f, acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.f;
    for (int s = 0; s < nsplit; s++) {
        const float2 ml = __ldcg(reinterpret_cast<const float2*>(a.part_ml + ((rbase + s) * group + g) * 2));
        if (ml.x == -INFINITY) continue;
        const float wgt = __expf(ml.x - M);
        L = fmaf(ml.y, wgt, L);
        const float* pa = a.part_acc + ((rbase + s) * group + g) * a.hd + d;
        const float4 p0 = __ldcg(reinterpret_cast<const float4*>(pa)), p1 = __ldcg(reinterpret_cast<const float4*>(pa + 4));
        acc[0] = fmaf(p0.x, wgt, acc[0]); acc[1] = fmaf(p0.y, wgt, acc[1]); acc[2] = fmaf(p0.z, wgt, acc[2]); acc[3] = fmaf(p0.w, wgt, acc[3]);
        acc[4] = fmaf(p1.x, wgt, acc[4]); acc[5] = fmaf(p1.y, wgt, acc[5]); acc[6] = fmaf(p1.z, wgt, acc[6]); acc[7] = fmaf(p1.w, wgt, acc[7]);
    }
    const float inv = 1.0f / L;
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = acc[i] * inv;
}

// ---- one GEMV-type phase for one CTA ---------------------------------------------------------
template <int M>
__device__ __noinline__ void mega_gemv_phase(const MegaArgs& a, const MegaPhase& ph, const MegaSmem sm, RingPos& rp, int token,
                                             int pos, unsigned long long& best_key, int tid) {
    const int lane = tid & 31, w = tid >> 5;
    const int ks = ph.ks, RC = 8 / ks, slice = 256 * M;
    const int q = w % ks, rloc = w / ks;
    const int K = ph.K, type = ph.type;
    // ---- input vector -> registers (fused RMSNorm / split-K attention merge) ----
    const float* xsrc = type == PH_DOWN ? a.act : a.h;
    const bool from_embed = (type == PH_QKV && ph.layer == 0);
    const int nsplit = mega_nsplit(a, pos + 1);
    float xr[M * 8];
    if (ks == 1) {
        // every warp needs the same K floats: fetch once per CTA, then fan out through smem
        if (type == PH_OPROJ) {
            for (int k = tid * 8; k < K; k += kMegaConsumerThreads * 8) {
                float v[8];
                mega_attn_combine8(a, k, nsplit, v);
                sts128f(sm.xs + k * 4, make_float4(v[0], v[1], v[2], v[3]));
                sts128f(sm.xs + k * 4 + 16, make_float4(v[4], v[5], v[6], v[7]));
            }
        } else {
            for (int k = tid * 4; k < K; k += kMegaConsumerThreads * 4) {
                float4 v;
                if (from_embed) {
                    const uint2 e = *reinterpret_cast<const uint2*>(a.embed + static_cast<size_t>(token) * a.H + k);
                    v = make_float4(bf16lo(e.x), bf16hi(e.x), bf16lo(e.y), bf16hi(e.y));
                } else {
                    v = __ldcg(reinterpret_cast<const float4*>(xsrc + k));
                }
                sts128f(sm.xs + k * 4, v);
            }
        }
        consumer_bar();
#pragma unroll
        for (int i = 0; i < M; i++) {
            const float4 v0 = lds128f(sm.xs + (i * 256 + lane * 8) * 4), v1 = lds128f(sm.xs + (i * 256 + lane * 8) * 4 + 16);
            xr[i * 8 + 0] = v0.x; xr[i * 8 + 1] = v0.y; xr[i * 8 + 2] = v0.z; xr[i * 8 + 3] = v0.w;
            xr[i * 8 + 4] = v1.x; xr[i * 8 + 5] = v1.y; xr[i * 8 + 6] = v1.z; xr[i * 8 + 7] = v1.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < M; i++) {
            const int k = q * slice + i * 256 + lane * 8;
            if (type == PH_OPROJ) {
                mega_attn_combine8(a, k, nsplit, &xr[i * 8]);
                continue;
            }
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
        uint4 nw[M];
#pragma unroll
        for (int i = 0; i < M; i++) nw[i] = __ldg(reinterpret_cast<const uint4*>(ph.norm_w + q * slice + i * 256 + lane * 8));
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < M * 8; i++) ss = fmaf(xr[i], xr[i], ss);
        ss = warp_sum(ss);
        float* red = sm.generic(sm.red);
        if (lane == 0) red[w] = ss;
        consumer_bar();
        float tot = 0.f;
        for (int i = 0; i < ks; i++) tot += red[i];  // warps 0..ks-1 hold slices 0..ks-1
        const float inv = rsqrtf(tot / static_cast<float>(K) + a.eps);
#pragma unroll
        for (int i = 0; i < M; i++) {
            xr[i * 8 + 0] = bf16lo(nw[i].x) * (xr[i * 8 + 0] * inv);
            xr[i * 8 + 1] = bf16hi(nw[i].x) * (xr[i * 8 + 1] * inv);
            xr[i * 8 + 2] = bf16lo(nw[i].y) * (xr[i * 8 + 2] * inv);
            xr[i * 8 + 3] = bf16hi(nw[i].y) * (xr[i * 8 + 3] * inv);
            xr[i * 8 + 4] = bf16lo(nw[i].z) * (xr[i * 8 + 4] * inv);
            xr[i * 8 + 5] = bf16hi(nw[i].z) * (xr[i * 8 + 5] * inv);
            xr[i * 8 + 6] = bf16lo(nw[i].w) * (xr[i * 8 + 6] * inv);
            xr[i * 8 + 7] = bf16hi(nw[i].w) * (xr[i * 8 + 7] * inv);
        }
    }

    // ---- stream this CTA's rows ----
    int r0, r1;
    mega_row_range(ph.N, type == PH_GATEUP ? 2 : 1, blockIdx.x, gridDim.x, r0, r1);
    const int n_chunks = (r1 - r0 + RC - 1) / RC;
    const bool cross = (ks > 1) || type == PH_GATEUP;  // result needs more than one warp
    const uint32_t unit_off = (static_cast<uint32_t>(rloc) * K + static_cast<uint32_t>(q) * slice) * 2 + lane * 16;
    float* part_base = sm.generic(sm.part);
    for (int ch = 0; ch < n_chunks; ch++) {
        const int row = r0 + ch * RC + rloc;
        const bool valid = row < r1;
        mbar_wait(sm.full + rp.stage * 8, rp.parity, a.abort_flag, 200 + type);
        float acc[M];
        if (valid) {
            const uint32_t base = sm.ring + static_cast<uint32_t>(rp.stage) * kMegaStageBytes + unit_off;
            uint4 wv[M];
#pragma unroll
            for (int i = 0; i < M; i++) wv[i] = lds128(base + i * 512);
#pragma unroll
            for (int i = 0; i < M; i++) {
                float s = bf16lo(wv[i].x) * xr[i * 8 + 0];
                s = fmaf(bf16hi(wv[i].x), xr[i * 8 + 1], s);
                s = fmaf(bf16lo(wv[i].y), xr[i * 8 + 2], s);
                s = fmaf(bf16hi(wv[i].y), xr[i * 8 + 3], s);
                s = fmaf(bf16lo(wv[i].z), xr[i * 8 + 4], s);
                s = fmaf(bf16hi(wv[i].z), xr[i * 8 + 5], s);
                s = fmaf(bf16lo(wv[i].w), xr[i * 8 + 6], s);
                s = fmaf(bf16hi(wv[i].w), xr[i * 8 + 7], s);
                acc[i] = s;
            }
        } else {
#pragma unroll
            for (int i = 0; i < M; i++) acc[i] = 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.empty + rp.stage * 8);  // smem slot can be refilled
        rp.advance();
        // tree-add the M sweep sums, then across lanes
#pragma unroll
        for (int st = 1; st < M; st <<= 1) {
#pragma unroll
            for (int i = 0; i + st < M; i += 2 * st) acc[i] += acc[i + st];
        }
        float s = warp_sum(acc[0]);
        if (cross) {
            float* part = part_base + (ch & 1) * 8;
            if (lane == 0) part[w] = s;
            consumer_bar();
            if (q != 0 || (type == PH_GATEUP && (rloc & 1))) continue;  // one finalising warp per row / pair
            s = 0.f;
            for (int i = 0; i < ks; i++) s += part[w + i];
            if (type == PH_GATEUP) {
                float u = 0.f;
                for (int i = 0; i < ks; i++) u += part[w + ks + i];
                s = (s / (1.0f + __expf(-s))) * u;  // silu(gate) * up
            }
        }
        if (lane == 0 && valid) {
            switch (type) {
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

// ---- attention work item: (kv head, split); partials are merged by the O-proj phase's x load -----
template <int HD, int GROUP>
__device__ __noinline__ void mega_attn_item(const MegaArgs& a, const MegaPhase& ph, const MegaSmem sm, int kvh, int split,
                                            int nsplit, int pos, int tid) {
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
        const int d0 = sl * 8, j0r = d0 < HALF ? d0 : d0 - HALF;  // the slice lies in one half (HALF % 8 == 0)
        const float4 xa0 = __ldcg(reinterpret_cast<const float4*>(head + j0r)), xa1 = __ldcg(reinterpret_cast<const float4*>(head + j0r + 4));
        const float4 xb0 = __ldcg(reinterpret_cast<const float4*>(head + j0r + HALF)), xb1 = __ldcg(reinterpret_cast<const float4*>(head + j0r + HALF + 4));
        const float x0[8] = {xa0.x, xa0.y, xa0.z, xa0.w, xa1.x, xa1.y, xa1.z, xa1.w};
        const float x1[8] = {xb0.x, xb0.y, xb0.z, xb0.w, xb1.x, xb1.y, xb1.z, xb1.w};
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const float4 c2 = __ldg(reinterpret_cast<const float4*>(cs + 2 * (j0r + i)));  // (c, s, c', s')
            out[i] = d0 < HALF ? x0[i] * c2.x - x1[i] * c2.y : x1[i] * c2.x + x0[i] * c2.y;
            out[i + 1] = d0 < HALF ? x0[i + 1] * c2.z - x1[i + 1] * c2.w : x1[i + 1] * c2.z + x0[i + 1] * c2.w;
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
        uint4 kw = make_uint4(0, 0, 0, 0), vw = make_uint4(0, 0, 0, 0);
        if (valid) {
            const int page = __ldg(a.block_table + j / a.page_size), off = j % a.page_size;
            uint16_t* kp = kv.at(page, 0, off) + kvh * HD + sl * 8;
            uint16_t* vp = kv.at(page, 1, off) + kvh * HD + sl * 8;
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
        }
        const float kf[8] = {bf16lo(kw.x), bf16hi(kw.x), bf16lo(kw.y), bf16hi(kw.y), bf16lo(kw.z), bf16hi(kw.z), bf16lo(kw.w), bf16hi(kw.w)};
        const float vf[8] = {bf16lo(vw.x), bf16hi(vw.x), bf16lo(vw.y), bf16hi(vw.y), bf16lo(vw.z), bf16hi(vw.z), bf16lo(vw.w), bf16hi(vw.w)};
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
    float* s_acc = sm.generic(sm.attn_scratch);               // [8][GROUP][HD]
    float* s_ml = s_acc + kMegaConsumerWarps * GROUP * HD;    // [8][GROUP][2]
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
        float Mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < kMegaConsumerWarps; t++) Mx = fmaxf(Mx, s_ml[(t * GROUP + g) * 2]);
        float L = 0.f, A = 0.f;
        if (Mx > -INFINITY) {
#pragma unroll
            for (int t = 0; t < kMegaConsumerWarps; t++) {
                const float wgt = __expf(s_ml[(t * GROUP + g) * 2] - Mx);
                L = fmaf(s_ml[(t * GROUP + g) * 2 + 1], wgt, L);
                A = fmaf(s_acc[(t * GROUP + g) * HD + d], wgt, A);
            }
        }
        a.part_acc[(pbase * GROUP + g) * HD + d] = A;
        if (d == 0) {
            a.part_ml[(pbase * GROUP + g) * 2] = Mx;
            a.part_ml[(pbase * GROUP + g) * 2 + 1] = L;
        }
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

// the producer's view of the weight stream: every chunk this CTA needs, in model order, forever
struct ChunkCursor {
    int step, pi, row, r1, RC;
    long long index;
    __device__ __forceinline__ bool done(const MegaArgs& a) const { return step >= a.n_steps; }
    __device__ __forceinline__ void seek_phase(const MegaArgs& a, int G) {
        // move to the first row of the next phase that has weights and rows for this CTA
        for (;;) {
            if (pi >= a.n_phases) {
                pi = 0;
                step++;
                if (step >= a.n_steps) return;
            }
            const MegaPhase& ph = a.phases[pi];
            if (ph.type != PH_ATTN) {
                int r0;
                mega_row_range(ph.N, ph.type == PH_GATEUP ? 2 : 1, blockIdx.x, G, r0, r1);
                if (r0 < r1) {
                    row = r0;
                    RC = 8 / ph.ks;
                    return;
                }
            }
            pi++;
        }
    }
    __device__ __forceinline__ void get(const MegaArgs& a, const uint16_t*& src, uint32_t& bytes) const {
        const MegaPhase& ph = a.phases[pi];
        src = ph.W + static_cast<size_t>(row) * ph.K;
        bytes = static_cast<uint32_t>(min(RC, r1 - row)) * ph.K * 2;
    }
    __device__ __forceinline__ void next(const MegaArgs& a, int G) {
        index++;
        row += RC;
        if (row >= r1) {
            pi++;
            seek_phase(a, G);
        }
    }
};

constexpr int kMegaL2Ahead = 16;  // chunks prefetched into L2 beyond the shared-memory ring

__global__ void __launch_bounds__(kMegaThreads, 1) mega_decode_kernel(const MegaArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    MegaSmem sm;
    sm.gen = reinterpret_cast<float*>(smem_raw);
    sm.base = smem_u32(smem_raw);
    sm.ring = sm.base;
    uint32_t p = sm.base + static_cast<uint32_t>(a.n_stages) * kMegaStageBytes;
    sm.full = p;  p += 8 * kMegaMaxStages;
    sm.empty = p; p += 8 * kMegaMaxStages;
    sm.keys = p;  p += 8 * 8;
    sm.red = p;   p += 4 * 32;
    sm.part = p;  p += 4 * 16;
    sm.xs = p;    p += 4 * kMegaXsFloats;
    sm.attn_scratch = p;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < a.n_stages; s++) {
            mbar_init(reinterpret_cast<uint64_t*>(sm.generic(sm.full + s * 8)), 1);
            mbar_init(reinterpret_cast<uint64_t*>(sm.generic(sm.empty + s * 8)), kMegaConsumerWarps);
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
            ChunkCursor ld{0, 0, 0, 0, 1, 0}, pf{0, 0, 0, 0, 1, 0};
            ld.seek_phase(a, G);
            pf.seek_phase(a, G);
            RingPos rp{0, 0, a.n_stages};
            while (!ld.done(a)) {
                // HBM -> L2 runs kMegaL2Ahead chunks ahead of the ring, so HBM keeps streaming while
                // the consumers sit in a grid barrier or in attention with the ring full
                while (!pf.done(a) && pf.index < ld.index + a.n_stages + kMegaL2Ahead) {
                    if (pf.index >= ld.index + a.n_stages) {
                        const uint16_t* src;
                        uint32_t bytes;
                        pf.get(a, src, bytes);
                        tma_prefetch_l2(src, bytes);
                    }
                    pf.next(a, G);
                }
                mbar_wait(sm.empty + rp.stage * 8, rp.parity ^ 1, a.abort_flag, 300);
                const uint16_t* src;
                uint32_t bytes;
                ld.get(a, src, bytes);
                mbar_arrive_expect_tx(sm.full + rp.stage * 8, bytes);
                tma_bulk_g2s(sm.ring + static_cast<uint32_t>(rp.stage) * kMegaStageBytes, src, bytes, sm.full + rp.stage * 8, policy);
                rp.advance();
                ld.next(a, G);
            }
        }
        return;
    }

    // ================= consumer warps =================
    RingPos rp{0, 0, a.n_stages};
    unsigned long long nbar = 0;  // grid barriers passed in this launch
    int token = token0;
    for (int step = 0; step < a.n_steps; step++) {
        const int pos = pos0 + step;
        unsigned long long best_key = 0ull;
        if (a.prof && step == a.n_steps - 1 && tid == 0 && (blockIdx.x == 0 || blockIdx.x == G - 1)) {
            const unsigned long long t = globaltimer_ns();
            a.prof[(blockIdx.x == 0 ? 0 : 2) * (a.n_phases + 1)] = t;
            a.prof[(blockIdx.x == 0 ? 1 : 3) * (a.n_phases + 1)] = t;
        }
        for (int pi = 0; pi < a.n_phases; pi++) {
            const MegaPhase& ph = a.phases[pi];
            if (ph.type == PH_ATTN) {
                const int nsplit = mega_nsplit(a, pos + 1);
                const int item = blockIdx.x;
                if (item < a.nkv * nsplit) {
                    const int kvh = item / nsplit, split = item % nsplit;
                    if (a.hd == 64) mega_attn_group<64>(a, ph, sm, kvh, split, nsplit, pos, tid);
                    else if (a.hd == 128) mega_attn_group<128>(a, ph, sm, kvh, split, nsplit, pos, tid);
                    else mega_attn_group<32>(a, ph, sm, kvh, split, nsplit, pos, tid);
                }
            } else {
                switch (ph.m) {
                    case 1: mega_gemv_phase<1>(a, ph, sm, rp, token, pos, best_key, tid); break;
                    case 2: mega_gemv_phase<2>(a, ph, sm, rp, token, pos, best_key, tid); break;
                    case 3: mega_gemv_phase<3>(a, ph, sm, rp, token, pos, best_key, tid); break;
                    case 4: mega_gemv_phase<4>(a, ph, sm, rp, token, pos, best_key, tid); break;
                    case 5: mega_gemv_phase<5>(a, ph, sm, rp, token, pos, best_key, tid); break;
                    case 6: mega_gemv_phase<6>(a, ph, sm, rp, token, pos, best_key, tid); break;
                    case 7: mega_gemv_phase<7>(a, ph, sm, rp, token, pos, best_key, tid); break;
                    default: mega_gemv_phase<8>(a, ph, sm, rp, token, pos, best_key, tid); break;
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
                unsigned long long* keys = reinterpret_cast<unsigned long long*>(sm.generic(sm.keys));
                if (lane == 0) keys[w] = best_key;
                consumer_bar();
                if (tid == 0) {
                    unsigned long long k = keys[0];
                    for (int i = 1; i < kMegaConsumerWarps; i++) k = keys[i] > k ? keys[i] : k;
                    atomicMax(a.argmax_keys + (step % 3), k);
                }
            }
            nbar++;
            const bool prof = a.prof && step == a.n_steps - 1 && tid == 0 && (blockIdx.x == 0 || blockIdx.x == G - 1);
            if (prof) a.prof[(blockIdx.x == 0 ? 0 : 2) * (a.n_phases + 1) + pi + 1] = globaltimer_ns();
            mega_grid_sync(a, epoch + nbar * G, tid);
            if (prof) a.prof[(blockIdx.x == 0 ? 1 : 3) * (a.n_phases + 1) + pi + 1] = globaltimer_ns();
            if (pi == 1 && blockIdx.x == 0 && tid == 0) a.argmax_keys[(step + 1) % 3] = 0ull;  // last read two steps ago
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

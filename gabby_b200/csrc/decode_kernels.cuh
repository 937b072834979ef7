// decode_kernels.cuh -- multi-kernel decode path (one CUDA graph per token; any small batch).
//
// Every kernel works on R "rows"; a row is one (sequence slot, position) pair: a decode step has
// one row per sequence, a chunked prefill has several consecutive positions of one sequence.
// Activations stay fp32 end to end (they are tiny next to the weights); weights and the KV
// cache are bf16; all dot products accumulate in fp32.
//
// No reference counterpart exists (gabby's forward is a stub: /root/reference/src/inference/
// generator.cc:33-38); the math follows oracle/llama_oracle.cc function by function.
#pragma once
#include "common.cuh"

namespace b2l {

// ------------------------------------------------------------------------------------------
// token embedding gather: h[r][:] = E[token[r]][:]                 (oracle: embedding gather)
// ------------------------------------------------------------------------------------------
__global__ void embed_kernel(const uint16_t* __restrict__ E, const int32_t* __restrict__ tokens,
                             float* __restrict__ h, int H, int V) {
    pdl_launch_dependents();
    pdl_wait();
    const int r = blockIdx.x;
    int tok = tokens[r];
    tok = min(max(tok, 0), V - 1);
    const uint4* src = reinterpret_cast<const uint4*>(E + static_cast<size_t>(tok) * H);
    float4* dst = reinterpret_cast<float4*>(h + static_cast<size_t>(r) * H);
    for (int i = threadIdx.x; i < H / 8; i += blockDim.x) {
        const uint4 w = __ldg(src + i);
        dst[2 * i] = make_float4(bf16lo(w.x), bf16hi(w.x), bf16lo(w.y), bf16hi(w.y));
        dst[2 * i + 1] = make_float4(bf16lo(w.z), bf16hi(w.z), bf16lo(w.w), bf16hi(w.w));
    }
}

// ------------------------------------------------------------------------------------------
// GEMV: y[b][n] (op)= sum_k W[n][k] * xhat[b][k],  xhat = NORM ? rmsnorm(x) * norm_w : x
//   MODE 0 store | 1 accumulate into y (residual add) | 2 SwiGLU over (gate,up) row pairs
// One warp owns two weight rows at a time (the pair, in MODE 2) and streams them with 128-bit
// L1-bypassing loads; x lives in shared memory as fp32 in K-tiles of `kt` elements.
// ------------------------------------------------------------------------------------------
// Tensor-parallel partial sums go straight into every rank's receive slab over NVLink peer memory, as 8-byte
// {value bits, sequence number} words (the "LL" idea: the flag travels with the data, an 8-byte store is atomic,
// so the receiver needs no fence and no separate signal -- it polls the word until the sequence number matches).
constexpr int kTpMaxRanks = 8;
struct TpSend {
    uint2* dst[kTpMaxRanks];   // per rank (self included): that rank's slab for (slot, this rank), [rows_cap][H]
    const uint32_t* seq;       // this rank's sequence number for the slot (device memory; bumped by the receiver kernel)
    int tp;
};
__device__ __forceinline__ void tp_send_pair(const TpSend& t, size_t idx, float v0, float v1, uint32_t seq) {
    // idx even: two adjacent outputs = one 16-byte store per peer (each 8-byte half is self-describing)
    for (int p = 0; p < t.tp; p++)
        asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(t.dst[p] + idx), "r"(__float_as_uint(v0)), "r"(seq),
                     "r"(__float_as_uint(v1)), "r"(seq) : "memory");
}

struct GemvArgs {
    const uint16_t* W;       // [N][K] bf16
    const float* x;          // [B][ldx]
    float* y;                // [B][ldy]
    const uint16_t* norm_w;  // [K] bf16 (NORM only)
    const float* add;        // optional [B][ldadd]: x <- x + add before use (TP partial-sum fold), may be null
    float eps;
    int N, K, ldx, ldy, kt;
    TpSend tps;              // MODE 3 only
};

constexpr int kGemvThreads = 256;
constexpr int kGemvWarps = kGemvThreads / 32;
constexpr int kGemvRowsPerCta = kGemvWarps * 2;

template <int B, int MODE, bool NORM>
__global__ void __launch_bounds__(kGemvThreads) gemv_kernel(const GemvArgs a) {
    extern __shared__ __align__(16) float xs[];  // [B][kt]
    __shared__ float s_red[B][kGemvWarps];
    __shared__ float s_inv[B];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    pdl_launch_dependents();
    pdl_wait();
    uint32_t tp_seq = 0;
    if (MODE == 3) tp_seq = *reinterpret_cast<const volatile uint32_t*>(a.tps.seq);

    if (NORM) {
        // LlamaRMSNorm: inv = rsqrt(mean(x^2) + eps)            (oracle: rmsnorm())
        float ss[B];
#pragma unroll
        for (int b = 0; b < B; b++) ss[b] = 0.f;
        for (int k = tid * 4; k < a.K; k += kGemvThreads * 4) {
#pragma unroll
            for (int b = 0; b < B; b++) {
                const float4 v = *reinterpret_cast<const float4*>(a.x + static_cast<size_t>(b) * a.ldx + k);
                ss[b] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
        }
#pragma unroll
        for (int b = 0; b < B; b++) {
            const float s = warp_sum(ss[b]);
            if (lane == 0) s_red[b][warp] = s;
        }
        __syncthreads();
        if (tid < B) {
            float s = 0.f;
            for (int w = 0; w < kGemvWarps; w++) s += s_red[tid][w];
            s_inv[tid] = rsqrtf(s / static_cast<float>(a.K) + a.eps);
        }
        __syncthreads();
    }

    const int n_blocks = (a.N + kGemvRowsPerCta - 1) / kGemvRowsPerCta;
    const int n_tiles = (a.K + a.kt - 1) / a.kt;
    int loaded_tile = -1;

    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const int row0 = blk * kGemvRowsPerCta + warp * 2;
        const bool live = row0 < a.N;  // N is even for every Llama shape
        const uint16_t* w0 = a.W + static_cast<size_t>(live ? row0 : 0) * a.K;
        const uint16_t* w1 = w0 + a.K;
        float acc0[B], acc1[B];
#pragma unroll
        for (int b = 0; b < B; b++) acc0[b] = acc1[b] = 0.f;

        for (int t = 0; t < n_tiles; t++) {
            const int k0 = t * a.kt, klen = min(a.kt, a.K - k0);
            if (loaded_tile != t) {
                __syncthreads();
                for (int k = tid * 4; k < klen; k += kGemvThreads * 4) {
#pragma unroll
                    for (int b = 0; b < B; b++) {
                        float4 v = *reinterpret_cast<const float4*>(a.x + static_cast<size_t>(b) * a.ldx + k0 + k);
                        if (NORM) {
                            const uint2 nw = *reinterpret_cast<const uint2*>(a.norm_w + k0 + k);
                            const float inv = s_inv[b];
                            v.x = bf16lo(nw.x) * (v.x * inv);
                            v.y = bf16hi(nw.x) * (v.y * inv);
                            v.z = bf16lo(nw.y) * (v.z * inv);
                            v.w = bf16hi(nw.y) * (v.w * inv);
                        }
                        *reinterpret_cast<float4*>(xs + b * a.kt + k) = v;
                    }
                }
                __syncthreads();
                loaded_tile = t;
            }
            if (live) {
                // four 256-element steps per iteration: all eight 16-byte weight loads are issued before the first FMA.
                // (Written as "load, use" per step, ptxas keeps exactly one pair of loads in flight per warp and the
                // kernel runs at memory latency; the warp-sync is a scheduling fence that keeps the loads ahead.)
                constexpr int U = 4;
                for (int kw = 0; kw < klen; kw += 256 * U) {   // warp-uniform trip count: the fence below is a full-warp sync
                    const int k = kw + lane * 8;
                    uint4 wa[U], wb[U];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int kk = k + u * 256;
                        wa[u] = make_uint4(0, 0, 0, 0);
                        wb[u] = make_uint4(0, 0, 0, 0);
                        if (kk < klen) {
                            wa[u] = ldg_stream(w0 + k0 + kk);
                            wb[u] = ldg_stream(w1 + k0 + kk);
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int kk = k + u * 256;
                        if (kk < klen) {
#pragma unroll
                            for (int b = 0; b < B; b++) {
                                const float4 x0 = *reinterpret_cast<const float4*>(xs + b * a.kt + kk);
                                const float4 x1 = *reinterpret_cast<const float4*>(xs + b * a.kt + kk + 4);
                                acc0[b] = dot8(wa[u], x0, x1, acc0[b]);
                                acc1[b] = dot8(wb[u], x0, x1, acc1[b]);
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int b = 0; b < B; b++) {
            const float s0 = warp_sum(acc0[b]);
            const float s1 = warp_sum(acc1[b]);
            if (lane == 0 && live) {
                if (MODE == 3) {
                    // row-parallel projection under TP: this rank's partial sum goes to every rank over NVLink
                    tp_send_pair(a.tps, static_cast<size_t>(b) * a.ldy + row0, s0, s1, tp_seq);
                } else if (MODE == 2) {
                    // silu(gate) * up                               (oracle: MLP block)
                    a.y[static_cast<size_t>(b) * a.ldy + (row0 >> 1)] = (s0 / (1.0f + __expf(-s0))) * s1;
                } else if (MODE == 1) {
                    float* y = a.y + static_cast<size_t>(b) * a.ldy + row0;
                    y[0] += s0;
                    y[1] += s1;
                } else {
                    float* y = a.y + static_cast<size_t>(b) * a.ldy + row0;
                    y[0] = s0;
                    y[1] = s1;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// RoPE (rotate-half, table-driven) on q in place + K/V append into the paged bf16 cache
//                                                   (oracle: rope_inplace(), KV store with ORC_KV_BF16)
// ------------------------------------------------------------------------------------------
struct KvLayout {
    uint16_t* pool;       // this layer's pool: [num_pages][2][page_size][kvd]
    int page_size, kvd;   // kvd = local kv heads * head_dim
    __device__ __forceinline__ uint16_t* at(int page, int which, int off) const {
        return pool + ((static_cast<size_t>(page) * 2 + which) * page_size + off) * kvd;
    }
};

struct RowMeta {
    const int32_t* positions;     // [R]
    const int32_t* slots;         // [R] row -> block table row
    const int32_t* block_tables;  // [n_slots][max_blocks]
    int max_blocks;
};

// first_head = 0: rotate q in place and append k, v; first_head = nh: only the k heads and v (the tcgen05 prefill attention
// rotates q itself while it stages it, flash_prefill_tc.cuh)
__global__ void rope_kv_kernel(float* __restrict__ qkv, int ld, const float* __restrict__ rope, KvLayout kv,
                               RowMeta rm, int nh, int nkv, int hd, int first_head) {
    pdl_launch_dependents();
    pdl_wait();
    const int r = blockIdx.x;
    const int pos = rm.positions[r];
    const int half = hd >> 1;
    const float* cs = rope + static_cast<size_t>(pos) * hd;  // [half][2]
    const int page = rm.block_tables[static_cast<size_t>(rm.slots[r]) * rm.max_blocks + pos / kv.page_size];
    const int off = pos % kv.page_size;
    float* row = qkv + static_cast<size_t>(r) * ld;
    uint16_t* kdst = kv.at(page, 0, off);
    uint16_t* vdst = kv.at(page, 1, off);
    const int qd = nh * hd, kvd = nkv * hd;
    for (int i = first_head * half + threadIdx.x; i < (nh + nkv) * half; i += blockDim.x) {
        const int head = i / half, j = i % half;
        const float c = cs[2 * j], s = cs[2 * j + 1];
        float* v = row + head * hd;  // q heads then k heads are contiguous in the fused row
        const float x0 = v[j], x1 = v[j + half];
        const float y0 = x0 * c - x1 * s, y1 = x1 * c + x0 * s;
        if (head < nh) {
            v[j] = y0;
            v[j + half] = y1;
        } else {
            const int kh = head - nh;
            kdst[kh * hd + j] = f32_to_bf16_bits(y0);
            kdst[kh * hd + j + half] = f32_to_bf16_bits(y1);
        }
    }
    const float* vsrc = row + qd + kvd;
    for (int i = threadIdx.x; i < kvd; i += blockDim.x) vdst[i] = f32_to_bf16_bits(vsrc[i]);
}

// ------------------------------------------------------------------------------------------
// GQA decode attention over the paged cache, split along the context ("split-K"), with the
// last-arriving CTA of each (row, kv head) merging the splits.
//                                                   (oracle: causal softmax(q k^T / sqrt(d)) v)
// grid (nsplit, nkv, R); 128 threads. LPT lanes share one token (8 dims each).
// ------------------------------------------------------------------------------------------
struct AttnArgs {
    const float* qkv;  // [R][ld], q already rotated
    int ld;
    KvLayout kv;
    RowMeta rm;
    float* part_acc;   // [R][nkv][nsplit][GROUP][HD]
    float* part_ml;    // [R][nkv][nsplit][GROUP][2]
    int* counters;     // [R][nkv], zero between launches
    float* out;        // [R][ldo]  (heads concatenated)
    int ldo;
    float scale;
    // optional: also write the output as bf16 hi/lo rows, the B operand of the tcgen05 O-projection (skinny_gemm.cuh):
    // row r -> split_out[r][ldo], row split_T + r -> the low parts
    uint16_t* split_out = nullptr;
    int split_T = 0;
    int interleave = 0;   // tensor-core kernel only: tiles dealt round-robin over all warps of all splits (attn_decode_mma.cuh)
};

__device__ __forceinline__ void attn_store_out(const AttnArgs& a, int r, int col, float v) {
    a.out[static_cast<size_t>(r) * a.ldo + col] = v;
    if (a.split_out) {
        const uint16_t hi = f32_to_bf16_bits(v);
        a.split_out[static_cast<size_t>(r) * a.ldo + col] = hi;
        a.split_out[static_cast<size_t>(a.split_T + r) * a.ldo + col] = f32_to_bf16_bits(v - bf16_bits_to_f32(hi));
    }
}

constexpr int kAttnThreads = 128;
constexpr int kAttnWarps = kAttnThreads / 32;

template <int HD, int GROUP>
__global__ void __launch_bounds__(kAttnThreads) attn_decode_kernel(const AttnArgs a) {
    constexpr int LPT = HD / 8;        // lanes per token
    constexpr int TPW = 32 / LPT;      // tokens per warp step
    constexpr int NG = kAttnWarps * TPW;  // token groups in the CTA
    __shared__ float s_m[NG][GROUP], s_l[NG][GROUP];
    __shared__ float s_acc[NG][GROUP][HD];
    __shared__ int s_last;

    pdl_launch_dependents();
    pdl_wait();

    const int split = blockIdx.x, nsplit = gridDim.x, kvh = blockIdx.y, nkv = gridDim.y, r = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sub = lane / LPT, sl = lane % LPT;  // token sub-slot, 8-dim slice
    const int grp = warp * TPW + sub;

    const int ctx = a.rm.positions[r] + 1;
    // the grid is sized for long contexts (several CTAs per SM); a short context uses only the first
    // ceil(ctx / 64) splits, the other CTAs leave at once and are not counted by the merge
    const int eff = max(1, min(nsplit, (ctx + 63) >> 6));
    if (split >= eff) return;
    const int chunk = (ctx + eff - 1) / eff;
    const int j0 = split * chunk, j1 = min(ctx, j0 + chunk);
    const int32_t* bt = a.rm.block_tables + static_cast<size_t>(a.rm.slots[r]) * a.rm.max_blocks;

    float q[GROUP][8];
#pragma unroll
    for (int g = 0; g < GROUP; g++) {
        const float* qp = a.qkv + static_cast<size_t>(r) * a.ld + (kvh * GROUP + g) * HD + sl * 8;
        const float4 q0 = *reinterpret_cast<const float4*>(qp), q1 = *reinterpret_cast<const float4*>(qp + 4);
        q[g][0] = q0.x * a.scale; q[g][1] = q0.y * a.scale; q[g][2] = q0.z * a.scale; q[g][3] = q0.w * a.scale;
        q[g][4] = q1.x * a.scale; q[g][5] = q1.y * a.scale; q[g][6] = q1.z * a.scale; q[g][7] = q1.w * a.scale;
    }
    float m[GROUP], l[GROUP], acc[GROUP][8];
#pragma unroll
    for (int g = 0; g < GROUP; g++) {
        m[g] = -INFINITY;
        l[g] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) acc[g][i] = 0.f;
    }

    constexpr int U = 4;  // tokens in flight per lane group: the loop is bound by load latency, not by math
    for (int jb = j0; jb < j1; jb += NG * U) {
        uint4 kw[U], vw[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = jb + u * NG + grp;
            kw[u] = make_uint4(0, 0, 0, 0);
            vw[u] = make_uint4(0, 0, 0, 0);
            if (j < j1) {
                const int page = bt[j / a.kv.page_size], off = j % a.kv.page_size;
                kw[u] = ldg_stream(reinterpret_cast<const uint4*>(a.kv.at(page, 0, off) + kvh * HD + sl * 8));
                vw[u] = ldg_stream(reinterpret_cast<const uint4*>(a.kv.at(page, 1, off) + kvh * HD + sl * 8));
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool valid = jb + u * NG + grp < j1;
            const float kf[8] = {bf16lo(kw[u].x), bf16hi(kw[u].x), bf16lo(kw[u].y), bf16hi(kw[u].y),
                                 bf16lo(kw[u].z), bf16hi(kw[u].z), bf16lo(kw[u].w), bf16hi(kw[u].w)};
            const float vf[8] = {bf16lo(vw[u].x), bf16hi(vw[u].x), bf16lo(vw[u].y), bf16hi(vw[u].y),
                                 bf16lo(vw[u].z), bf16hi(vw[u].z), bf16lo(vw[u].w), bf16hi(vw[u].w)};
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
    }

    // merge the NG token groups of this CTA
#pragma unroll
    for (int g = 0; g < GROUP; g++) {
        if (sl == 0) {
            s_m[grp][g] = m[g];
            s_l[grp][g] = l[g];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) s_acc[grp][g][sl * 8 + i] = acc[g][i];
    }
    __syncthreads();
    const size_t pbase = (static_cast<size_t>(r) * nkv + kvh) * nsplit + split;
    for (int e = tid; e < GROUP * HD; e += kAttnThreads) {
        const int g = e / HD, d = e % HD;
        float M = -INFINITY;
#pragma unroll
        for (int t = 0; t < NG; t++) M = fmaxf(M, s_m[t][g]);
        float L = 0.f, A = 0.f;
        if (M > -INFINITY) {
#pragma unroll
            for (int t = 0; t < NG; t++) {
                const float w = __expf(s_m[t][g] - M);  // exp(-inf) = 0 for empty groups
                L = fmaf(s_l[t][g], w, L);
                A = fmaf(s_acc[t][g][d], w, A);
            }
        }
        a.part_acc[(pbase * GROUP + g) * HD + d] = A;
        if (d == 0) {
            a.part_ml[(pbase * GROUP + g) * 2 + 0] = M;
            a.part_ml[(pbase * GROUP + g) * 2 + 1] = L;
        }
    }
    // last CTA of this (row, kv head) merges the splits
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int done = atomicAdd(a.counters + r * nkv + kvh, 1);
        s_last = (done == eff - 1);
        if (s_last) a.counters[r * nkv + kvh] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const size_t rbase = (static_cast<size_t>(r) * nkv + kvh) * nsplit;
    for (int e = tid; e < GROUP * HD; e += kAttnThreads) {
        const int g = e / HD, d = e % HD;
        float M = -INFINITY;
        for (int s = 0; s < eff; s++) M = fmaxf(M, __ldcg(a.part_ml + ((rbase + s) * GROUP + g) * 2));
        float L = 0.f, A = 0.f;
        for (int s = 0; s < eff; s++) {
            const float ms = __ldcg(a.part_ml + ((rbase + s) * GROUP + g) * 2);
            if (ms == -INFINITY) continue;
            const float w = __expf(ms - M);
            L = fmaf(__ldcg(a.part_ml + ((rbase + s) * GROUP + g) * 2 + 1), w, L);
            A = fmaf(__ldcg(a.part_acc + ((rbase + s) * GROUP + g) * HD + d), w, A);
        }
        attn_store_out(a, r, (kvh * GROUP + g) * HD + d, A / L);
    }
}

// ------------------------------------------------------------------------------------------
// greedy argmax (first max) over logits [R][ld]; one CTA per row.       (oracle: orc_argmax)
// Writes ids[r] = offset + argmax and, when vals != null, the max value (TP shard merge).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) argmax_kernel(const float* __restrict__ logits, int ld, int n, int offset,
                                                      int32_t* __restrict__ ids, float* __restrict__ vals) {
    __shared__ unsigned long long s_key[32];
    pdl_launch_dependents();
    pdl_wait();
    const int r = blockIdx.x, tid = threadIdx.x;
    const float* x = logits + static_cast<size_t>(r) * ld;
    unsigned long long best = 0ull;
    for (int i = tid; i < n; i += blockDim.x) {
        const unsigned long long k = argmax_key(x[i], i);
        best = k > best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((tid & 31) == 0) s_key[tid >> 5] = best;
    __syncthreads();
    if (tid < 32) {
        best = tid < (blockDim.x >> 5) ? s_key[tid] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (tid == 0) {
            const int idx = argmax_key_index(best);
            ids[r] = offset + idx;
            if (vals) vals[r] = x[idx];
        }
    }
}

// The same over a cluster of kArgmaxCluster CTAs per row: one 1024-thread CTA per row leaves 8 CTAs (batch 8) to scan
// 8 x 128256 logits (ncu: 63 us per step); here every CTA scans one slice of the row and CTA 0 of the cluster picks the
// best of the slices' keys through distributed shared memory. Keys carry the index in the row, so the result is the same
// first-max as argmax_kernel. grid (kArgmaxCluster, R).
constexpr int kArgmaxCluster = 8;
__global__ void __launch_bounds__(1024) argmax_cluster_kernel(const float* __restrict__ logits, int ld, int n, int offset,
                                                              int32_t* __restrict__ ids, float* __restrict__ vals) {
    __shared__ unsigned long long s_key[32];
    __shared__ unsigned long long s_cta_key;
    pdl_launch_dependents();
    pdl_wait();
    const int r = blockIdx.y, tid = threadIdx.x;
    const int rank = static_cast<int>(cluster_ctarank());
    const float* x = logits + static_cast<size_t>(r) * ld;
    const int per = (n + kArgmaxCluster - 1) / kArgmaxCluster;
    const int i0 = rank * per, i1 = min(n, i0 + per);
    unsigned long long best = 0ull;
#pragma unroll 8
    for (int i = i0 + tid; i < i1; i += 1024) {
        const unsigned long long k = argmax_key(x[i], i);
        best = k > best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((tid & 31) == 0) s_key[tid >> 5] = best;
    __syncthreads();
    if (tid < 32) {
        best = s_key[tid];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (tid == 0) s_cta_key = best;
    }
    cluster_sync_all();
    if (rank == 0 && tid == 0) {
        best = 0ull;
        for (int q = 0; q < kArgmaxCluster; q++) {
            const unsigned long long other = dsmem_ld_u64(&s_cta_key, static_cast<uint32_t>(q));
            best = other > best ? other : best;
        }
        const int idx = argmax_key_index(best);
        ids[r] = offset + idx;
        if (vals) vals[r] = x[idx];
    }
    cluster_sync_all();   // peers keep their shared memory alive until CTA 0 has read it
}

// device-resident greedy loop bookkeeping: feed the argmax back, advance positions, log the id
__global__ void advance_kernel(const int32_t* __restrict__ next_ids, int32_t* __restrict__ tokens,
                               int32_t* __restrict__ positions, int32_t* __restrict__ out_ids,
                               int32_t* __restrict__ step_counter, int n) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = threadIdx.x;
    const int step = *step_counter;
    if (i < n) {
        const int32_t id = next_ids[i];
        tokens[i] = id;
        positions[i] += 1;
        out_ids[static_cast<size_t>(step) * n + i] = id;
    }
    __syncthreads();
    if (i == 0) *step_counter = step + 1;
}

// y[i] += x[i]  (TP: add the all-reduced projection into the residual stream)
__global__ void add_kernel(float* __restrict__ y, const float* __restrict__ x, int n) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}

// TP receiver: h += sum over ranks (rank order: identical bits on every rank) of the partial sums that arrive in this
// rank's slab recv[src][rows_cap][H]; bumps the slot's sequence number when the whole grid is done.
__global__ void tp_ll_reduce_kernel(float* __restrict__ h, const uint2* recv, int tp, int R, int H, int rows_cap,
                                    uint32_t* seq, unsigned int* done) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t want = *reinterpret_cast<volatile uint32_t*>(seq);
    const int half = H >> 1, n_pairs = R * half;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += gridDim.x * blockDim.x) {
        const int r = i / half, n = (i - r * half) * 2;
        float s0 = 0.f, s1 = 0.f;
        for (int src = 0; src < tp; src++) {
            const uint2* p = recv + (static_cast<size_t>(src) * rows_cap + r) * H + n;
            uint32_t v0, q0, v1, q1;
            long long t0 = 0;
            for (unsigned spin = 0;; spin++) {
                asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(q0), "=r"(v1), "=r"(q1) : "l"(p) : "memory");
                if (q0 == want && q1 == want) break;
                if ((spin & 1023u) == 1023u) {            // a peer that never sends must not hang the GPU
                    const long long now = clock64();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > 120000000000ll) __trap();   // ~60 s
                }
            }
            s0 += __uint_as_float(v0);
            s1 += __uint_as_float(v1);
        }
        float2* dst = reinterpret_cast<float2*>(h + static_cast<size_t>(r) * H + n);
        float2 cur = *dst;
        cur.x += s0;
        cur.y += s1;
        *dst = cur;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(done, 1u) == gridDim.x - 1) {
            *done = 0;
            *reinterpret_cast<volatile uint32_t*>(seq) = want + 1;
        }
    }
}

// TP greedy argmax: pack this rank's (max value, global index) per row for the all-gather ...
__global__ void tp_pack_kernel(const float* __restrict__ vals, const int32_t* __restrict__ ids, float* __restrict__ pack, int R) {
    pdl_launch_dependents();
    pdl_wait();
    const int r = threadIdx.x;
    if (r < R) {
        pack[2 * r] = vals[r];
        pack[2 * r + 1] = static_cast<float>(ids[r]);   // < 2^24: exact in fp32
    }
}
// ... and pick the winner over ranks (ties -> lowest index = first max, like the single-GPU kernel)
__global__ void tp_merge_kernel(const float* __restrict__ gathered, int32_t* __restrict__ ids, int R, int stride_rows, int tp) {
    pdl_launch_dependents();
    pdl_wait();
    const int r = threadIdx.x;
    if (r >= R) return;
    unsigned long long best = 0ull;
    for (int k = 0; k < tp; k++) {
        const float* p = gathered + (static_cast<size_t>(k) * stride_rows + r) * 2;
        const unsigned long long key = argmax_key(p[0], static_cast<int>(p[1]));
        best = key > best ? key : best;
    }
    ids[r] = argmax_key_index(best);
}

// prefill: RMSNorm of fp32 rows -> bf16 rows (the A operand of the tensor-core GEMMs)
__global__ void rmsnorm_bf16_kernel(const float* __restrict__ x, const uint16_t* __restrict__ w, uint16_t* __restrict__ y, int H, float eps) {
    __shared__ float s_red[32];
    const int r = blockIdx.x;
    const float* xr = x + static_cast<size_t>(r) * H;
    // the row stays in registers between the two passes (up to 8 float4 per thread: H <= 8192 with 256 threads); all loads of a
    // thread are in flight together
    constexpr int kCache = 8;
    float4 cache[kCache];
    const int stride = blockDim.x * 4;
#pragma unroll
    for (int k = 0; k < kCache; k++) {
        const int i = threadIdx.x * 4 + k * stride;
        cache[k] = i < H ? *reinterpret_cast<const float4*>(xr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < kCache; k++) ss += cache[k].x * cache[k].x + cache[k].y * cache[k].y + cache[k].z * cache[k].z + cache[k].w * cache[k].w;
    for (int i = threadIdx.x * 4 + kCache * stride; i < H; i += stride) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); i++) tot += s_red[i];
    const float inv = rsqrtf(tot / static_cast<float>(H) + eps);
    auto emit = [&](int i, const float4& v) {
        const uint2 nw = *reinterpret_cast<const uint2*>(w + i);
        uint2 o;
        o.x = pack_bf16x2(bf16lo(nw.x) * (v.x * inv), bf16hi(nw.x) * (v.y * inv));
        o.y = pack_bf16x2(bf16lo(nw.y) * (v.z * inv), bf16hi(nw.y) * (v.w * inv));
        *reinterpret_cast<uint2*>(y + static_cast<size_t>(r) * H + i) = o;
    };
#pragma unroll
    for (int k = 0; k < kCache; k++) {
        const int i = threadIdx.x * 4 + k * stride;
        if (i < H) emit(i, cache[k]);
    }
    for (int i = threadIdx.x * 4 + kCache * stride; i < H; i += stride) emit(i, *reinterpret_cast<const float4*>(xr + i));
}

// prefill: fp32 -> bf16 (attention output -> A operand of the O projection)
__global__ void cast_bf16_kernel(const float* __restrict__ x, uint16_t* __restrict__ y, size_t n) {
    const size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 v = *reinterpret_cast<const float4*>(x + i);
        *reinterpret_cast<uint2*>(y + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
}

// prefill: gather rows (the last token of each sequence) for the lm_head
__global__ void gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ rows, float* __restrict__ dst, int H) {
    const float* s = src + static_cast<size_t>(rows[blockIdx.x]) * H;
    float* d = dst + static_cast<size_t>(blockIdx.x) * H;
    for (int i = threadIdx.x; i < H; i += blockDim.x) d[i] = s[i];
}

// standalone RMSNorm (parity tap of the final norm only; the hot path fuses it into the GEMV)
__global__ void rmsnorm_kernel(const float* __restrict__ x, const uint16_t* __restrict__ w, float* __restrict__ y,
                               int H, float eps) {
    __shared__ float s_red[32];
    const int r = blockIdx.x;
    const float* xr = x + static_cast<size_t>(r) * H;
    float ss = 0.f;
    for (int i = threadIdx.x; i < H; i += blockDim.x) ss += xr[i] * xr[i];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); i++) tot += s_red[i];
    const float inv = rsqrtf(tot / static_cast<float>(H) + eps);
    for (int i = threadIdx.x; i < H; i += blockDim.x) y[static_cast<size_t>(r) * H + i] = bf16_bits_to_f32(w[i]) * (xr[i] * inv);
}

}  // namespace b2l

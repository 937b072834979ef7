// skinny_gemm.cuh -- batched-decode projections on the tensor cores.
// For 2..16 sequences per step a CUDA-core GEMV does 8-16 FMAs per weight element and turns compute
// bound (measured 1.1 TB/s at batch 8), so the weights go through tcgen05 instead:
//     Y^T[N][BT] = W[N][K] * X[BT][K]^T        W is the 128-row A operand, X the narrow B operand
// X carries every activation row TWICE, as bf16 hi and bf16 lo (x = hi + lo to 16 mantissa bits), so
// the fp32 activations of the decode path keep their accuracy: Y = W hi + W lo.
// Grid = (N / 128 row tiles) x (K splits): the K split is what gives small matrices (O, down: 16-24
// row tiles) enough CTAs to pull full HBM bandwidth; partials are summed by skinny_reduce_kernel
// (deterministic order), which also applies the epilogue (store | residual add | SwiGLU).
// Same warp roles and TMA/tcgen05 plumbing as gemm_tcgen05.cuh.
// Fused reduce epilogues: skinny_reduce_rope_kv_kernel (QKV: + RoPE + K/V append into the paged cache),
// skinny_reduce_norm_split_cluster_kernel (O / down on one GPU: + residual + the next RMSNorm + bf16 hi/lo split,
// one thread-block cluster per activation row, sum of squares through distributed shared memory).
#pragma once
#include "gemm_tcgen05.cuh"
#include "decode_kernels.cuh"

namespace b2l {

constexpr int kSkinnyThreads = 192;
// ring depth and stage size per activation-tile width BT: two CTAs per SM stay resident (the whole grid is one wave)
__host__ __device__ constexpr int skinny_stages(int BT) { return BT <= 32 ? 5 : 4; }
__host__ __device__ constexpr uint32_t skinny_stage_bytes(int BT) { return 128 * kGemmBK * 2 + (BT * kGemmBK * 2 < 4096 ? 4096 : BT * kGemmBK * 2); }
__host__ __device__ constexpr size_t skinny_smem(int BT) { return static_cast<size_t>(skinny_stages(BT)) * skinny_stage_bytes(BT) + 16 * skinny_stages(BT) + 64 + 1024; }

template <int BT>  // columns of the MMA: 2 x (max rows): 16, 32 or 64
__global__ void __launch_bounds__(kSkinnyThreads, 2)
skinny_gemm_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, float* __restrict__ partial,
                   int N, int n_kblocks, int kblocks_per_split) {
    constexpr uint32_t kStageW = 128 * kGemmBK * 2, kStageX = BT * kGemmBK * 2;   // x tile padded to >= 4 KB (keeps 1024-byte alignment)
    constexpr uint32_t kStageBytes = skinny_stage_bytes(BT);
    constexpr int kSkinnyStages = skinny_stages(BT);
    extern __shared__ __align__(1024) uint8_t ssm[];
    const uint32_t base = (smem_u32(ssm) + 1023u) & ~1023u;
    const uint32_t bars = base + kSkinnyStages * kStageBytes;
    const uint32_t full = bars, empty = bars + 8 * kSkinnyStages, tmem_full = bars + 16 * kSkinnyStages, tmem_slot = tmem_full + 8;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * 128, split = blockIdx.y;
    const int kb0 = split * kblocks_per_split, kb1 = min(n_kblocks, kb0 + kblocks_per_split);

    if (tid == 0) {
        for (int s = 0; s < kSkinnyStages; s++) {
            mbar_init(full + 8 * s, 1);
            mbar_init(empty + 8 * s, 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(BT < 32 ? 32 : BT) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    pdl_launch_dependents();
    if (warp == 0) {
        if (lane == 0) {
            // the weights do not depend on the previous kernel: fill the ring with W tiles before the programmatic
            // dependency resolves, then add the activation tiles (written by the kernel before this one)
            const int n_pre = min(kSkinnyStages, kb1 - kb0);
            for (int i = 0; i < n_pre; i++) {
                mbar_arrive_expect_tx(full + 8 * i, kStageW + kStageX);
                tma_load_2d(base + i * kStageBytes, &map_w, (kb0 + i) * kGemmBK, n0, full + 8 * i);
            }
            pdl_wait();
            for (int i = 0; i < n_pre; i++) tma_load_2d(base + i * kStageBytes + kStageW, &map_x, (kb0 + i) * kGemmBK, 0, full + 8 * i);
            for (int kb = kb0 + n_pre; kb < kb1; kb++) {
                const int i = kb - kb0, s = i % kSkinnyStages;
                mbar_wait_spin(empty + 8 * s, ((i / kSkinnyStages) & 1) ^ 1);
                mbar_arrive_expect_tx(full + 8 * s, kStageW + kStageX);
                tma_load_2d(base + s * kStageBytes, &map_w, kb * kGemmBK, n0, full + 8 * s);
                tma_load_2d(base + s * kStageBytes + kStageW, &map_x, kb * kGemmBK, 0, full + 8 * s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, BT);
            for (int kb = kb0; kb < kb1; kb++) {
                const int i = kb - kb0, s = i % kSkinnyStages;
                mbar_wait_spin(full + 8 * s, (i / kSkinnyStages) & 1);
                tcgen05_fence_after();
                const uint64_t da = umma_smem_desc(base + s * kStageBytes), db = umma_smem_desc(base + s * kStageBytes + kStageW);
#pragma unroll
                for (int k = 0; k < kGemmBK / 16; k++) umma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, (i | k) != 0);
                tcgen05_commit(empty + 8 * s);
            }
            tcgen05_commit(tmem_full);
        }
    } else {
        const int quarter = warp & 3;
        const int row = n0 + quarter * 32 + lane;   // TMEM lane = weight row
        pdl_wait();   // `partial` is still being read by the reduce of the previous projection until then
        mbar_wait_spin(tmem_full, 0);
        tcgen05_fence_after();
        constexpr int T = BT / 2;                    // token rows: column t = hi part, column t + T = lo part
        float v[BT];
#pragma unroll
        for (int c0 = 0; c0 < BT; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0, r);
#pragma unroll
            for (int j = 0; j < 16; j++) v[c0 + j] = __uint_as_float(r[j]);
        }
        if (row < N) {
#pragma unroll
            for (int t = 0; t < T; t++) partial[(static_cast<size_t>(split) * T + t) * N + row] = v[t] + v[t + T];
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BT < 32 ? 32 : BT) : "memory");
    }
}

// y (op)= sum over K splits of partial[split][t][n];  mode 0 store, 1 += (residual), 2 SwiGLU over (2i, 2i+1),
// 3 = send the sums to every TP rank (tp_send_pair)
__global__ void skinny_reduce_kernel(const float* __restrict__ partial, int ksplit, int T, int N, int R, int mode,
                                     float* __restrict__ y, int ldy, const TpSend tps, uint16_t* __restrict__ split_out, int split_T) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.y;
    if (t >= R) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (mode == 3) {
        if (2 * i + 1 >= N) return;
        const uint32_t seq = *reinterpret_cast<const volatile uint32_t*>(tps.seq);
        float a0 = 0.f, a1 = 0.f;
        for (int s = 0; s < ksplit; s++) {
            const float2 p = *reinterpret_cast<const float2*>(partial + (static_cast<size_t>(s) * T + t) * N + 2 * i);
            a0 += p.x;
            a1 += p.y;
        }
        tp_send_pair(tps, static_cast<size_t>(t) * ldy + 2 * i, a0, a1, seq);
    } else if (mode == 2) {
        if (2 * i + 1 >= N) return;
        float g = 0.f, u = 0.f;
        for (int s = 0; s < ksplit; s++) {
            const float2 p = *reinterpret_cast<const float2*>(partial + (static_cast<size_t>(s) * T + t) * N + 2 * i);
            g += p.x;
            u += p.y;
        }
        const float v = (g / (1.0f + __expf(-g))) * u;
        if (split_out) {   // straight into the next projection's B operand (bf16 hi / lo rows): no fp32 round trip, no split kernel
            const uint16_t hi = f32_to_bf16_bits(v);
            split_out[static_cast<size_t>(t) * ldy + i] = hi;
            split_out[static_cast<size_t>(split_T + t) * ldy + i] = f32_to_bf16_bits(v - bf16_bits_to_f32(hi));
        } else {
            y[static_cast<size_t>(t) * ldy + i] = v;
        }
    } else {
        if (i >= N) return;
        float acc = 0.f;
        for (int s = 0; s < ksplit; s++) acc += partial[(static_cast<size_t>(s) * T + t) * N + i];
        float* dst = y + static_cast<size_t>(t) * ldy + i;
        *dst = mode == 1 ? *dst + acc : acc;
    }
}

// QKV projection epilogue fused with RoPE and the K/V append: sums the K splits of one head of one activation row, rotates
// the q / k heads (rotate-half pairs j, j + hd/2), stores q as fp32 into qkv (what the attention kernels read) and k, v as
// bf16 into the row's page of the cache. Replaces skinny_reduce_kernel(mode 0) + rope_kv_kernel: one launch instead of
// two, and (heads x rows) CTAs instead of rope_kv_kernel's one CTA per row. grid (nh + 2 nkv, R), hd / 2 threads.
// Same arithmetic, in the same order, as the two kernels it replaces (sum over splits in split order, then the rotation).
__global__ void skinny_reduce_rope_kv_kernel(const float* __restrict__ partial, int ksplit, int T, int N, float* __restrict__ qkv, int ld,
                                             const float* __restrict__ rope, KvLayout kv, RowMeta rm, int nh, int nkv, int hd) {
    pdl_launch_dependents();
    pdl_wait();
    const int head = blockIdx.x, r = blockIdx.y, j = threadIdx.x, half = hd >> 1;
    const int pos = rm.positions[r];
    const int c0 = head * hd + j, c1 = c0 + half;   // q heads, k heads, v heads are contiguous in the fused row
    float x0 = 0.f, x1 = 0.f;
#pragma unroll 4
    for (int s = 0; s < ksplit; s++) {
        const float* p = partial + (static_cast<size_t>(s) * T + r) * N;
        x0 += p[c0];
        x1 += p[c1];
    }
    if (head < nh) {
        const float c = rope[static_cast<size_t>(pos) * hd + 2 * j], s = rope[static_cast<size_t>(pos) * hd + 2 * j + 1];
        float* row = qkv + static_cast<size_t>(r) * ld;
        row[c0] = x0 * c - x1 * s;
        row[c1] = x1 * c + x0 * s;
        return;
    }
    const int page = rm.block_tables[static_cast<size_t>(rm.slots[r]) * rm.max_blocks + pos / kv.page_size];
    const int off = pos % kv.page_size;
    if (head < nh + nkv) {
        const float c = rope[static_cast<size_t>(pos) * hd + 2 * j], s = rope[static_cast<size_t>(pos) * hd + 2 * j + 1];
        uint16_t* kdst = kv.at(page, 0, off) + (head - nh) * hd;
        kdst[j] = f32_to_bf16_bits(x0 * c - x1 * s);
        kdst[j + half] = f32_to_bf16_bits(x1 * c + x0 * s);
    } else {
        uint16_t* vdst = kv.at(page, 1, off) + (head - nh - nkv) * hd;
        vdst[j] = f32_to_bf16_bits(x0);
        vdst[j + half] = f32_to_bf16_bits(x1);
    }
}

// Row-parallel projection epilogue fused with the NEXT projection's prologue: h[r] += sum over K splits of partial,
// then RMSNorm(h[r]) * norm_w -> bf16 hi/lo rows of the next B operand. One CTA per activation row (the norm needs the
// whole row); replaces skinny_reduce_kernel(mode 1) + split_bf16_kernel<true>.
__global__ void __launch_bounds__(1024) skinny_reduce_norm_split_kernel(const float* __restrict__ partial, int ksplit, int T, int N,
                                                                        float* __restrict__ h, const uint16_t* __restrict__ norm_w,
                                                                        uint16_t* __restrict__ out, float eps) {
    __shared__ float s_red[32];
    pdl_launch_dependents();
    pdl_wait();
    const int r = blockIdx.x, tid = threadIdx.x;
    constexpr int kMaxIter = 4;   // N <= 8192
    float2 v[kMaxIter];
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < kMaxIter; it++) {
        const int n = it * 2048 + tid * 2;
        v[it] = make_float2(0.f, 0.f);
        if (n < N) {
            float a0 = 0.f, a1 = 0.f;
            for (int s = 0; s < ksplit; s++) {
                const float2 p = *reinterpret_cast<const float2*>(partial + (static_cast<size_t>(s) * T + r) * N + n);
                a0 += p.x;
                a1 += p.y;
            }
            float2* hp = reinterpret_cast<float2*>(h + static_cast<size_t>(r) * N + n);
            float2 cur = *hp;
            cur.x += a0;
            cur.y += a1;
            *hp = cur;
            v[it] = cur;
            ss += cur.x * cur.x + cur.y * cur.y;
        }
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) s_red[tid >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < 32; i++) tot += s_red[i];
    const float inv = rsqrtf(tot / static_cast<float>(N) + eps);
#pragma unroll
    for (int it = 0; it < kMaxIter; it++) {
        const int n = it * 2048 + tid * 2;
        if (n < N) {
            const uint32_t nw = *reinterpret_cast<const uint32_t*>(norm_w + n);
            const float y0 = bf16lo(nw) * (v[it].x * inv), y1 = bf16hi(nw) * (v[it].y * inv);
            const uint16_t h0 = f32_to_bf16_bits(y0), h1 = f32_to_bf16_bits(y1);
            *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(r) * N + n) = static_cast<uint32_t>(h0) | (static_cast<uint32_t>(h1) << 16);
            *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(T + r) * N + n) = pack_bf16x2(y0 - bf16_bits_to_f32(h0), y1 - bf16_bits_to_f32(h1));
        }
    }
}

// The same epilogue spread over a thread-block CLUSTER per activation row: the one-CTA-per-row kernel above leaves 8 CTAs
// (batch 8) to pull ksplit x N partial sums through L2 one dependent load after the other (ncu: 11 us per launch, twice per
// layer). Here CTA `rank` of a cluster of kNormCluster owns columns [rank * cols, (rank + 1) * cols) of its row, keeps
// its updated h values in registers, and the row's sum of squares is exchanged through distributed shared memory
// (one float per CTA, summed in rank order by everybody: all CTAs get the identical 1/rms). grid (kNormCluster, R).
constexpr int kNormCluster = 8, kNormThreads = 256;
__global__ void __launch_bounds__(kNormThreads) skinny_reduce_norm_split_cluster_kernel(const float* __restrict__ partial, int ksplit, int T, int N,
                                                                                        float* __restrict__ h, const uint16_t* __restrict__ norm_w,
                                                                                        uint16_t* __restrict__ out, float eps) {
    __shared__ float s_red[kNormThreads / 32];
    __shared__ float s_cta_total;
    pdl_launch_dependents();
    pdl_wait();
    const int r = blockIdx.y, tid = threadIdx.x;
    const int rank = static_cast<int>(cluster_ctarank());
    const int cols = ((N / 2 + kNormCluster - 1) / kNormCluster) * 2;   // even column count per CTA
    const int c_lo = rank * cols, c_hi = min(N, c_lo + cols);
    constexpr int kMaxIter = 2;   // N <= 8192: <= 1024 columns per CTA
    float2 v[kMaxIter];
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < kMaxIter; it++) {
        const int n = c_lo + it * 2 * kNormThreads + tid * 2;
        v[it] = make_float2(0.f, 0.f);
        if (n < c_hi) {
            float a0 = 0.f, a1 = 0.f;
            const float* p = partial + static_cast<size_t>(r) * N + n;
            const size_t stride = static_cast<size_t>(T) * N;
#pragma unroll 4
            for (int s = 0; s < ksplit; s++) {
                const float2 q = *reinterpret_cast<const float2*>(p + s * stride);
                a0 += q.x;
                a1 += q.y;
            }
            float2* hp = reinterpret_cast<float2*>(h + static_cast<size_t>(r) * N + n);
            float2 cur = *hp;
            cur.x += a0;
            cur.y += a1;
            *hp = cur;
            v[it] = cur;
            ss += cur.x * cur.x + cur.y * cur.y;
        }
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) s_red[tid >> 5] = ss;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < kNormThreads / 32; i++) t += s_red[i];
        s_cta_total = t;
    }
    cluster_sync_all();   // every CTA's total is published (release / acquire at cluster scope)
    float tot = 0.f;
    for (int q = 0; q < kNormCluster; q++) tot += dsmem_ld_f32(&s_cta_total, static_cast<uint32_t>(q));
    const float inv = rsqrtf(tot / static_cast<float>(N) + eps);
#pragma unroll
    for (int it = 0; it < kMaxIter; it++) {
        const int n = c_lo + it * 2 * kNormThreads + tid * 2;
        if (n < c_hi) {
            const uint32_t nw = *reinterpret_cast<const uint32_t*>(norm_w + n);
            const float y0 = bf16lo(nw) * (v[it].x * inv), y1 = bf16hi(nw) * (v[it].y * inv);
            const uint16_t h0 = f32_to_bf16_bits(y0), h1 = f32_to_bf16_bits(y1);
            *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(r) * N + n) = static_cast<uint32_t>(h0) | (static_cast<uint32_t>(h1) << 16);
            *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(T + r) * N + n) = pack_bf16x2(y0 - bf16_bits_to_f32(h0), y1 - bf16_bits_to_f32(h1));
        }
    }
    cluster_sync_all();   // nobody exits while a peer may still read its shared memory
}

// fp32 rows -> bf16 (hi, lo) rows for the B operand; NORM: fused RMSNorm. out is [2*T][K]: row r = hi, row T + r = lo
template <bool NORM>
__global__ void split_bf16_kernel(const float* __restrict__ x, int ldx, const uint16_t* __restrict__ w, uint16_t* __restrict__ out,
                                  int K, int T, float eps) {
    __shared__ float s_red[32];
    pdl_launch_dependents();
    pdl_wait();
    const int r = blockIdx.x;
    const float* xr = x + static_cast<size_t>(r) * ldx;
    float inv = 1.f;
    if (NORM) {
        float ss = 0.f;
        for (int i = threadIdx.x * 4; i < K; i += blockDim.x * 4) {
            const float4 v = *reinterpret_cast<const float4*>(xr + i);
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        ss = warp_sum(ss);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
        __syncthreads();
        float tot = 0.f;
        for (int i = 0; i < (blockDim.x >> 5); i++) tot += s_red[i];
        inv = rsqrtf(tot / static_cast<float>(K) + eps);
    }
    for (int i = threadIdx.x * 4; i < K; i += blockDim.x * 4) {
        float4 v = *reinterpret_cast<const float4*>(xr + i);
        if (NORM) {
            const uint2 nw = *reinterpret_cast<const uint2*>(w + i);
            v.x = bf16lo(nw.x) * (v.x * inv); v.y = bf16hi(nw.x) * (v.y * inv);
            v.z = bf16lo(nw.y) * (v.z * inv); v.w = bf16hi(nw.y) * (v.w * inv);
        }
        const uint16_t h0 = f32_to_bf16_bits(v.x), h1 = f32_to_bf16_bits(v.y), h2 = f32_to_bf16_bits(v.z), h3 = f32_to_bf16_bits(v.w);
        const float l0 = v.x - bf16_bits_to_f32(h0), l1 = v.y - bf16_bits_to_f32(h1), l2 = v.z - bf16_bits_to_f32(h2), l3 = v.w - bf16_bits_to_f32(h3);
        *reinterpret_cast<uint2*>(out + static_cast<size_t>(r) * K + i) =
            make_uint2(static_cast<uint32_t>(h0) | (static_cast<uint32_t>(h1) << 16), static_cast<uint32_t>(h2) | (static_cast<uint32_t>(h3) << 16));
        *reinterpret_cast<uint2*>(out + static_cast<size_t>(T + r) * K + i) = make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
    }
}

}  // namespace b2l

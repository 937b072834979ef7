// gemm_tcgen05.cuh -- prefill GEMM on the 5th-generation tensor cores (sm_100a):
//     C[M][N] = A[M][K] * W[N][K]^T       A, W bf16 (both K-major, i.e. plain row-major), fp32 accumulate
//
// Warp-specialised, one 128 x BN output tile per CTA:
//   warp 0   TMA producer: cp.async.bulk.tensor.2d (128B-swizzled 64-element K blocks of A and W) into a
//            kStages-deep shared-memory ring, completing on `full` mbarriers
//   warp 1   MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16)
//            straight from shared-memory descriptors into a TMEM accumulator; tcgen05.commit releases
//            ring slots (`empty`) and finally signals the epilogue (`tmem_full`)
//   warps 2-5 epilogue: tcgen05.ld the accumulator (each warp owns the TMEM lane quarter warp_id % 4),
//            apply the fused epilogue (fp32 store | bf16 store | residual add | SwiGLU on column pairs)
// No reference counterpart (gabby has no GEMM); the math is oracle linear() with ORC_ACT_BF16.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "ptx_helpers.cuh"  // mbarrier / smem helpers

namespace b2l {

constexpr int kGemmBM = 128, kGemmBK = 64, kGemmStages = 3, kGemmThreads = 192;  // 3 stages x 32 KB: two CTAs per SM, one CTA's epilogue overlaps the other's main loop

enum GemmEpilogue { GEMM_STORE_F32 = 0, GEMM_STORE_BF16 = 1, GEMM_ADD_F32 = 2, GEMM_SWIGLU_BF16 = 3 };

struct GemmArgs {
    float* c_f32;        // GEMM_STORE_F32 / GEMM_ADD_F32: [M][ldc]
    uint16_t* c_bf16;    // GEMM_STORE_BF16: [M][ldc]; GEMM_SWIGLU_BF16: [M][ldc] with N/2 columns
    int M, N, K, ldc;
    int epilogue;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
// bounded: a bad tensor map or a lost TMA transaction traps (the launch fails with an error) instead of hanging the GPU.
// try_wait suspends the thread for a hardware-defined interval per attempt, so 2^26 attempts is tens of seconds.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row groups are 1024 B apart
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);   // start address, 16-byte units
    d |= static_cast<uint64_t>(0) << 16;                       // leading byte offset (unused for swizzled K-major)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;               // stride byte offset: 8 rows x 128 B
    d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (sm_100)
    d |= static_cast<uint64_t>(2) << 61;                       // layout: SWIZZLE_128B
    return d;
}
// instruction descriptor: D fp32, A/B bf16, both K-major, shape M x N
__device__ __host__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const GemmArgs g) {
    constexpr uint32_t kStageA = kGemmBM * kGemmBK * 2, kStageB = BN * kGemmBK * 2;
    constexpr uint32_t kTmemCols = BN;  // fp32 accumulator: one column per output column (power of two >= 32)
    extern __shared__ __align__(1024) uint8_t gsm[];
    const uint32_t base = (smem_u32(gsm) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t sA = base, sB = base + kGemmStages * kStageA;
    const uint32_t bars = sB + kGemmStages * kStageB;
    const uint32_t full = bars, empty = bars + 8 * kGemmStages, tmem_full = bars + 16 * kGemmStages, tmem_slot = tmem_full + 8;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * BN;
    const int n_kblocks = g.K / kGemmBK;

    if (tid == 0) {
        for (int s = 0; s < kGemmStages; s++) {
            mbar_init(full + 8 * s, 1);
            mbar_init(empty + 8 * s, 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // one warp allocates the accumulator columns and later frees them
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer ----
            for (int kb = 0; kb < n_kblocks; kb++) {
                const int s = kb % kGemmStages;
                const uint32_t parity = (kb / kGemmStages) & 1;
                mbar_wait_spin(empty + 8 * s, parity ^ 1);
                mbar_arrive_expect_tx(full + 8 * s, kStageA + kStageB);
                tma_load_2d(sA + s * kStageA, &map_a, kb * kGemmBK, m0, full + 8 * s);
                tma_load_2d(sB + s * kStageB, &map_w, kb * kGemmBK, n0, full + 8 * s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer ----
            constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM, BN);
            for (int kb = 0; kb < n_kblocks; kb++) {
                const int s = kb % kGemmStages;
                const uint32_t parity = (kb / kGemmStages) & 1;
                mbar_wait_spin(full + 8 * s, parity);
                tcgen05_fence_after();
                const uint64_t da = umma_smem_desc(sA + s * kStageA), db = umma_smem_desc(sB + s * kStageB);
#pragma unroll
                for (int k = 0; k < kGemmBK / 16; k++) {
                    // advancing K by 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in 16-byte units
                    umma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                }
                tcgen05_commit(empty + 8 * s);   // slot reusable once these MMAs have read it
            }
            tcgen05_commit(tmem_full);           // accumulator complete
        }
    } else {
        // ---- epilogue warps 2..5: TMEM lane quarter = warp % 4, one output row per thread ----
        const int quarter = warp & 3;
        const int row = m0 + quarter * 32 + lane;
        mbar_wait_spin(tmem_full, 0);
        tcgen05_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0, r);
            if (row < g.M) {
                const int col = n0 + c0;
                if (g.epilogue == GEMM_STORE_F32) {
                    float4* dst = reinterpret_cast<float4*>(g.c_f32 + static_cast<size_t>(row) * g.ldc + col);
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
                } else if (g.epilogue == GEMM_ADD_F32) {
                    float4* dst = reinterpret_cast<float4*>(g.c_f32 + static_cast<size_t>(row) * g.ldc + col);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        float4 v = dst[i];
                        v.x += __uint_as_float(r[4 * i]); v.y += __uint_as_float(r[4 * i + 1]);
                        v.z += __uint_as_float(r[4 * i + 2]); v.w += __uint_as_float(r[4 * i + 3]);
                        dst[i] = v;
                    }
                } else if (g.epilogue == GEMM_STORE_BF16) {
                    uint4* dst = reinterpret_cast<uint4*>(g.c_bf16 + static_cast<size_t>(row) * g.ldc + col);
#pragma unroll
                    for (int i = 0; i < 2; i++)
                        dst[i] = make_uint4(pack_bf16x2(__uint_as_float(r[8 * i]), __uint_as_float(r[8 * i + 1])),
                                            pack_bf16x2(__uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3])),
                                            pack_bf16x2(__uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5])),
                                            pack_bf16x2(__uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7])));
                } else {  // GEMM_SWIGLU_BF16: W rows are (gate, up) pairs -> 8 outputs per 16 columns
                    float o[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float gate = __uint_as_float(r[2 * i]), up = __uint_as_float(r[2 * i + 1]);
                        o[i] = (gate / (1.0f + __expf(-gate))) * up;
                    }
                    uint4* dst = reinterpret_cast<uint4*>(g.c_bf16 + static_cast<size_t>(row) * g.ldc + (col >> 1));
                    *dst = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
                }
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}


// ---- persistent version: 128 x 256 tiles, double-buffered TMEM accumulator ------------------------------------------------
// One CTA per SM walks a static tile schedule. The 128 x 128 kernel above reads (128 + 128) x 16 x 2 B = 8 KB of shared memory
// per 64-clock MMA -- exactly the 128 B/clk the SM's shared memory delivers, so the tensor pipe can never be kept full. A
// 128 x 256 tile reads 12 KB per 128-clock MMA (96 B/clk). The whole TMEM (512 columns) holds TWO fp32 accumulators: the
// epilogue warps drain tile i (tcgen05.ld -> fused epilogue -> global) while the MMA thread already fills tile i + 1, and the
// TMA ring (4 stages x 48 KB) runs across tile boundaries, so neither the pipeline fill nor the epilogue is exposed.
// Tile order: bands of kGemmBandN column tiles, row tiles fastest inside a band (the 148 tiles in flight then share ~18 A row
// tiles and the band's 8 W column tiles: everything but the first touch comes out of L2).
constexpr int kGemmPBN = 256, kGemmPStages = 4, kGemmBandN = 8;
constexpr int kGemmPThreads = 192;   // TMA warp, MMA warp, 4 epilogue warps (one per TMEM lane quarter)
constexpr uint32_t kGemmPStgRow = 144, kGemmPStgWarp = 32 * kGemmPStgRow;   // per-warp transposition block: 32 rows x (32 fp32 + 16 B pad)
constexpr uint32_t kGemmPStageA = kGemmBM * kGemmBK * 2, kGemmPStageB = kGemmPBN * kGemmBK * 2;
constexpr size_t kGemmPSmem = static_cast<size_t>(kGemmPStages) * (kGemmPStageA + kGemmPStageB) + 256 + 4 * kGemmPStgWarp + 1024;

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// tile t of the schedule -> (row tile, column tile)
__device__ __forceinline__ void gemm_tile_coords(int t, int m_tiles, int n_tiles, int& mt, int& nt) {
    const int band_tiles = kGemmBandN * m_tiles;
    const int band = t / band_tiles, r = t - band * band_tiles;
    const int bw = min(kGemmBandN, n_tiles - band * kGemmBandN);   // the last band may be narrower
    mt = r / bw;
    nt = band * kGemmBandN + (r - mt * bw);
}

__device__ __forceinline__ float4 ldg_f4_early(const float4* p) {   // issued where it is written (ahead of the TMEM wait it overlaps)
    float4 v;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kGemmPThreads, 1)
gemm_bf16_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const GemmArgs g) {
    extern __shared__ __align__(1024) uint8_t gsm[];
    const uint32_t base = (smem_u32(gsm) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + kGemmPStages * kGemmPStageA;
    const uint32_t bars = sB + kGemmPStages * kGemmPStageB;
    const uint32_t full = bars, empty = bars + 8 * kGemmPStages, tmem_full = bars + 16 * kGemmPStages, tmem_empty = tmem_full + 16,
                   tmem_slot = tmem_empty + 16, stg_base = bars + 256;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_kblocks = g.K / kGemmBK;
    const int m_tiles = (g.M + kGemmBM - 1) / kGemmBM, n_tiles = g.N / kGemmPBN;
    const int total = m_tiles * n_tiles;

    if (tid == 0) {
        for (int s = 0; s < kGemmPStages; s++) {
            mbar_init(full + 8 * s, 1);
            mbar_init(empty + 8 * s, 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(tmem_full + 8 * i, 1);
            mbar_init(tmem_empty + 8 * i, 4);   // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // the whole TMEM: two 256-column accumulators (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer: the ring runs across tile boundaries ----
            uint32_t it = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                int mt, nt;
                gemm_tile_coords(t, m_tiles, n_tiles, mt, nt);
                for (int kb = 0; kb < n_kblocks; kb++, it++) {
                    const uint32_t s = it % kGemmPStages, parity = (it / kGemmPStages) & 1;
                    mbar_wait_spin(empty + 8 * s, parity ^ 1);
                    mbar_arrive_expect_tx(full + 8 * s, kGemmPStageA + kGemmPStageB);
                    tma_load_2d(sA + s * kGemmPStageA, &map_a, kb * kGemmBK, mt * kGemmBM, full + 8 * s);
                    tma_load_2d(sB + s * kGemmPStageB, &map_w, kb * kGemmBK, nt * kGemmPBN, full + 8 * s);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer ----
            constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM, kGemmPBN);
            uint32_t it = 0, ti = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ti++) {
                const uint32_t acc = ti & 1;
                mbar_wait_spin(tmem_empty + 8 * acc, ((ti >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kGemmPBN;
                for (int kb = 0; kb < n_kblocks; kb++, it++) {
                    const uint32_t s = it % kGemmPStages, parity = (it / kGemmPStages) & 1;
                    mbar_wait_spin(full + 8 * s, parity);
                    tcgen05_fence_after();
                    const uint64_t da = umma_smem_desc(sA + s * kGemmPStageA), db = umma_smem_desc(sB + s * kGemmPStageB);
#pragma unroll
                    for (int k = 0; k < kGemmBK / 16; k++) umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    tcgen05_commit(empty + 8 * s);
                }
                tcgen05_commit(tmem_full + 8 * acc);
            }
        }
    } else {
        // ---- epilogue warps 2..5: TMEM lane quarter = warp % 4. tcgen05.ld hands every thread one output ROW; written like
        // that, a warp's 16-byte accesses land in 32 different rows (the fp32 epilogues ran at a third of the main loop's pace
        // and throttled it: ncu, o-projection shape, tensor pipe 60 % busy against 89 % with the bf16 SwiGLU epilogue). The
        // fp32 epilogues therefore transpose each 32 x 32 block through shared memory: a warp instruction then covers four
        // full 128-byte lines. ----
        const int quarter = warp & 3;
        const bool f32 = g.epilogue == GEMM_STORE_F32 || g.epilogue == GEMM_ADD_F32, add = g.epilogue == GEMM_ADD_F32;
        const uint32_t stg = stg_base + static_cast<uint32_t>(warp - 2) * kGemmPStgWarp;
        const int tr = lane >> 3, tc = (lane & 7) * 4;   // transposed mapping: lane -> row tr + 4 i, columns tc .. tc + 3
        uint32_t ti = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ti++) {
            int mt, nt;
            gemm_tile_coords(t, m_tiles, n_tiles, mt, nt);
            const uint32_t acc = ti & 1;
            const int row0 = mt * kGemmBM + quarter * 32;   // first row of this warp's 32
            const int row = row0 + lane;
            const int n0 = nt * kGemmPBN;
            // residual add: the old values do not depend on the accumulator -- fetch the first block before waiting for the
            // MMAs and block c + 1 while block c is added and stored
            float4 v[8];
            if (add) {
#pragma unroll
                for (int i = 0; i < 8; i++)
                    v[i] = ldg_f4_early(reinterpret_cast<const float4*>(g.c_f32 + static_cast<size_t>(min(row0 + tr + 4 * i, g.M - 1)) * g.ldc + n0 + tc));
            }
            mbar_wait_spin(tmem_full + 8 * acc, (ti >> 1) & 1);
            tcgen05_fence_after();
            const uint32_t t_row = tmem_base + acc * kGemmPBN + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < kGemmPBN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(t_row + c0, r);
                const int col = n0 + c0;
                if (f32) {
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        sts128f(stg + lane * kGemmPStgRow + i * 16, make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                                                                 __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
                    __syncwarp();
                    float4 vn[8];
                    if (add && c0 + 32 < kGemmPBN) {
#pragma unroll
                        for (int i = 0; i < 8; i++)
                            vn[i] = ldg_f4_early(reinterpret_cast<const float4*>(g.c_f32 + static_cast<size_t>(min(row0 + tr + 4 * i, g.M - 1)) * g.ldc + col + 32 + tc));
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        float4 x = lds128f(stg + (tr + 4 * i) * kGemmPStgRow + tc * 4);
                        if (add) { x.x += v[i].x; x.y += v[i].y; x.z += v[i].z; x.w += v[i].w; }
                        if (row0 + tr + 4 * i < g.M) *reinterpret_cast<float4*>(g.c_f32 + static_cast<size_t>(row0 + tr + 4 * i) * g.ldc + col + tc) = x;
                    }
                    __syncwarp();   // the block is rewritten by the next iteration
                    if (add) {
#pragma unroll
                        for (int i = 0; i < 8; i++) v[i] = vn[i];
                    }
                } else if (row < g.M) {
                    if (g.epilogue == GEMM_STORE_BF16) {
                        uint4* dst = reinterpret_cast<uint4*>(g.c_bf16 + static_cast<size_t>(row) * g.ldc + col);
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            dst[i] = make_uint4(pack_bf16x2(__uint_as_float(r[8 * i]), __uint_as_float(r[8 * i + 1])),
                                                pack_bf16x2(__uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3])),
                                                pack_bf16x2(__uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5])),
                                                pack_bf16x2(__uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7])));
                    } else {  // GEMM_SWIGLU_BF16: W rows are (gate, up) pairs -> 16 outputs per 32 columns
                        uint32_t o[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const float g0 = __uint_as_float(r[4 * i]), u0 = __uint_as_float(r[4 * i + 1]);
                            const float g1 = __uint_as_float(r[4 * i + 2]), u1 = __uint_as_float(r[4 * i + 3]);
                            o[i] = pack_bf16x2((g0 / (1.0f + __expf(-g0))) * u0, (g1 / (1.0f + __expf(-g1))) * u1);
                        }
                        uint4* dst = reinterpret_cast<uint4*>(g.c_bf16 + static_cast<size_t>(row) * g.ldc + (col >> 1));
                        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + 8 * acc);   // this warp's quarter of the accumulator may be overwritten
        }
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace b2l

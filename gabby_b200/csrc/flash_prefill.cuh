// flash_prefill.cuh -- causal GQA attention for prefill, flash-style (online softmax, no S x S matrix),
// over the paged bf16 KV cache. One CTA = 64 consecutive query positions of one sequence x GH query
// heads that share a kv head (GH = 3 for Llama-3.2-3B's group of 3): 4 x GH warps, each 16 query rows of
// one head, all reading the SAME K/V tiles -- the tiles of 64 tokens are gathered page by page into shared
// memory with cp.async once per group instead of once per query head. S = Q K^T and O += P V run on the
// tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate).
//   q: fp32 [T][ld] already rotated (rope_kv_kernel), rounded to bf16 on load
//   out: bf16 [T][ldo] (the A operand of the O-projection GEMM)
// Math = oracle attention with ORC_KV_BF16 | ORC_QP_BF16 (P rounded to bf16 before P V).
#pragma once
#include "common.cuh"
#include "decode_kernels.cuh"

namespace b2l {

struct PrefillTile {
    int row0;    // first row (token index in this prefill call)
    int n_rows;  // <= 64
    int pos0;    // position of row0 in its sequence
    int slot;    // block-table row
};

struct FlashArgs {
    const float* qkv;
    int ld;
    KvLayout kv;
    const int32_t* block_tables;
    int max_blocks;
    const PrefillTile* tiles;
    uint16_t* out;
    int ldo;
    int group;  // query heads per kv head
    float scale_log2e;
};

constexpr int kFlashBM = 64, kFlashBN = 64, kFlashThreads = 128;   // threads per query head of a CTA

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// heads per CTA: the largest divisor of the group that the register file allows (head_dim 128: 3 x 128 threads x ~170 registers)
__host__ __device__ constexpr int flash_heads_per_cta(int hd, int group) {
    const int cap = hd >= 128 ? 3 : 4;
    for (int g = cap; g > 1; g--)
        if (group % g == 0) return g;
    return 1;
}

template <int HD, int GH>
__global__ void __launch_bounds__(kFlashThreads * GH) flash_prefill_kernel(const FlashArgs a) {
    constexpr int LDS = HD + 8;           // padded row (bf16 elements): 16-byte shift per row kills ldmatrix bank conflicts
    constexpr int KSTEPS = HD / 16;       // k-steps of Q K^T
    constexpr int DBLOCKS = HD / 8;       // 8-wide output column blocks
    extern __shared__ __align__(16) uint16_t fsm[];
    uint16_t* sKV = fsm;                       // two buffers of {K [64][LDS], V [64][LDS]}: tile i+1 loads while tile i computes
    uint16_t* sQ = fsm;                        // [GH][64][LDS] staged in the same memory before the first K/V tile is issued
    static_assert(GH <= 4, "the Q tiles of the CTA's heads are staged in the K/V buffers");
    const PrefillTile tile = a.tiles[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
    const int hg = (tid >> 5) / 4, warp = (tid >> 5) % 4;        // head within the CTA, 16-row block within the tile
    const int head0 = blockIdx.y * GH, head = head0 + hg, kvh = head0 / a.group;
    constexpr int kThreads = kFlashThreads * GH;
    const int32_t* bt = a.block_tables + static_cast<size_t>(tile.slot) * a.max_blocks;

    // ---- Q tile: fp32 (rotated) -> bf16 in shared memory; rows past the tile repeat the last row ----
    for (int i = tid; i < GH * kFlashBM * (HD / 4); i += kThreads) {
        const int h = i / (kFlashBM * (HD / 4)), r = (i / (HD / 4)) % kFlashBM, c = (i % (HD / 4)) * 4;
        const int row = tile.row0 + min(r, tile.n_rows - 1);
        const float4 v = *reinterpret_cast<const float4*>(a.qkv + static_cast<size_t>(row) * a.ld + (head0 + h) * HD + c);
        *reinterpret_cast<uint2*>(sQ + (h * kFlashBM + r) * LDS + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
    __syncthreads();
    uint32_t qf[KSTEPS][4];
    {
        const int mat = lane >> 3, r = (lane & 7) + (mat & 1) * 8, c = (mat >> 1) * 8;
#pragma unroll
        for (int kk = 0; kk < KSTEPS; kk++) ldmatrix_x4(smem_u32(sQ + (hg * kFlashBM + warp * 16 + r) * LDS + kk * 16 + c), qf[kk]);
    }
    __syncthreads();   // the Q staging area becomes the K/V buffers
    float o[DBLOCKS][4];
#pragma unroll
    for (int d = 0; d < DBLOCKS; d++) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // rows gid and gid + 8 of this warp's 16
    const int q_pos0 = tile.pos0 + warp * 16 + gid, q_pos1 = q_pos0 + 8;
    const int kv_end = tile.pos0 + tile.n_rows;                 // tokens [0, kv_end) exist for this tile

    // gather K and V rows of tokens [j0, j0+64) from their pages into buffer `buf`
    auto issue_tile = [&](int j0, int buf) {
        uint16_t* sK = sKV + buf * 2 * kFlashBN * LDS;
        uint16_t* sV = sK + kFlashBN * LDS;
        for (int i = tid; i < kFlashBN * (HD / 8); i += kThreads) {
            const int r = i / (HD / 8), c = (i % (HD / 8)) * 8;
            const int j = min(j0 + r, kv_end - 1);
            const int page = bt[j / a.kv.page_size], off = j % a.kv.page_size;
            cp_async16(smem_u32(sK + r * LDS + c), a.kv.at(page, 0, off) + kvh * HD + c);
            cp_async16(smem_u32(sV + r * LDS + c), a.kv.at(page, 1, off) + kvh * HD + c);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue_tile(0, 0);
    int buf = 0;
    for (int j0 = 0; j0 < kv_end; j0 += kFlashBN, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");   // this tile (the only group in flight) has landed
        __syncthreads();                                        // ... for every thread, and everybody is done with the other buffer
        if (j0 + kFlashBN < kv_end) issue_tile(j0 + kFlashBN, buf ^ 1);   // overlaps with the math below
        const uint16_t* sK = sKV + buf * 2 * kFlashBN * LDS;
        const uint16_t* sV = sK + kFlashBN * LDS;

        // ---- S = Q K^T : 16 x 64 per warp ----
        float s[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; nb++) s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
#pragma unroll
        for (int nb = 0; nb < 8; nb++) {
#pragma unroll
            for (int k2 = 0; k2 < KSTEPS / 2; k2++) {  // two k-steps (32 dims) per ldmatrix.x4
                uint32_t kf[4];
                ldmatrix_x4(smem_u32(sK + (nb * 8 + (lane & 7)) * LDS + k2 * 32 + (lane >> 3) * 8), kf);
                mma_bf16_16816(s[nb], qf[2 * k2], kf[0], kf[1]);
                mma_bf16_16816(s[nb], qf[2 * k2 + 1], kf[2], kf[3]);
            }
        }
        // ---- scale, causal mask, online softmax ----
        const bool diag = j0 + kFlashBN - 1 > tile.pos0 + warp * 16;  // some token in this tile may be masked for some row
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int nb = 0; nb < 8; nb++) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int j = j0 + nb * 8 + tig * 2 + (e & 1);
                const int qp = (e < 2) ? q_pos0 : q_pos1;
                float v = s[nb][e] * a.scale_log2e;
                if (diag && j > qp) v = -INFINITY;
                if (j >= kv_end) v = -INFINITY;
                s[nb][e] = v;
            }
            mx0 = fmaxf(mx0, fmaxf(s[nb][0], s[nb][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nb][2], s[nb][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        // every query row sees token 0, so mx is finite after the first tile
        const float c0 = exp2f(m0 - mx0), c1 = exp2f(m1 - mx1);
        m0 = mx0;
        m1 = mx1;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pf[8][2];  // P as bf16 pairs: [nb][0] = row gid, [nb][1] = row gid + 8
#pragma unroll
        for (int nb = 0; nb < 8; nb++) {
            const float p0 = exp2f(s[nb][0] - mx0), p1 = exp2f(s[nb][1] - mx0), p2 = exp2f(s[nb][2] - mx1), p3 = exp2f(s[nb][3] - mx1);
            rs0 += p0 + p1;
            rs1 += p2 + p3;
            pf[nb][0] = pack_bf16x2(p0, p1);
            pf[nb][1] = pack_bf16x2(p2, p3);
        }
        l0 = l0 * c0 + rs0;
        l1 = l1 * c1 + rs1;
#pragma unroll
        for (int d = 0; d < DBLOCKS; d++) {
            o[d][0] *= c0; o[d][1] *= c0; o[d][2] *= c1; o[d][3] *= c1;
        }
        // ---- O += P V : k = tokens (4 k-steps of 16), n = head dims ----
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const uint32_t pa[4] = {pf[2 * t][0], pf[2 * t][1], pf[2 * t + 1][0], pf[2 * t + 1][1]};
#pragma unroll
            for (int d2 = 0; d2 < DBLOCKS / 2; d2++) {  // two 8-wide dim blocks per ldmatrix.x4.trans
                uint32_t vf[4];
                const int mat = lane >> 3;
                ldmatrix_x4_trans(smem_u32(sV + (t * 16 + (lane & 7) + (mat & 1) * 8) * LDS + d2 * 16 + (mat >> 1) * 8), vf);
                mma_bf16_16816(o[2 * d2], pa, vf[0], vf[1]);
                mma_bf16_16816(o[2 * d2 + 1], pa, vf[2], vf[3]);
            }
        }
    }
    // ---- finish: sum l over the quad, normalise, store bf16 ----
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int r0 = warp * 16 + gid, r1 = r0 + 8;
#pragma unroll
    for (int d = 0; d < DBLOCKS; d++) {
        const int col = head * HD + d * 8 + tig * 2;
        if (r0 < tile.n_rows)
            *reinterpret_cast<uint32_t*>(a.out + static_cast<size_t>(tile.row0 + r0) * a.ldo + col) = pack_bf16x2(o[d][0] * i0, o[d][1] * i0);
        if (r1 < tile.n_rows)
            *reinterpret_cast<uint32_t*>(a.out + static_cast<size_t>(tile.row0 + r1) * a.ldo + col) = pack_bf16x2(o[d][2] * i1, o[d][3] * i1);
    }
}

}  // namespace b2l

// attn_decode_mma.cuh -- split-K GQA decode attention on the tensor cores (mma.sync m16n8k16 bf16).
// The CUDA-core kernel in decode_kernels.cuh spends ~40 issue slots per (token, query head) on FMAs and shuffles
// and tops out near 1.1 TB/s of K/V traffic; batched long-context decode (3B batch 8 @ 2K, 8B/70B TP @ 4-8K)
// is then attention bound. Here the GROUP query heads that share a kv head are the M rows of the MMA
// (padded to 16), a warp owns 16-token K/V tiles (gathered page by page with cp.async into a private,
// double-buffered shared-memory tile, so warps never wait for each other), S = Q K^T and O += P V run on
// the tensor cores. q (fp32, already scaled) and the probabilities are split into bf16 hi + lo parts and each
// product is issued twice, so the result keeps the fp32-activation accuracy of the decode path
// (error ~2^-16, same parity tolerances as the CUDA-core kernel). Same partial-sum buffers and last-arriver
// merge as attn_decode_kernel; the launcher (engine.cu: attn_launch_hd) sizes the split count so that the
// whole grid is resident in one wave (occupancy API). The 16-token tiles of a (row, kv head) are dealt
// round-robin over all warps of all splits, each warp keeps NBUF - 1 tiles in flight and holds the page
// indices of its next 32 tiles in its lanes; the last-arriving CTA merges the splits eight at a time.
#pragma once
#include "flash_prefill.cuh"

namespace b2l {

// NBUF tile buffers per warp: NBUF - 1 tiles (K + V, 16 tokens each) are in flight while one is multiplied
template <int HD, int NBUF = 2>
constexpr size_t attn_mma_smem() { return static_cast<size_t>(kAttnWarps) * NBUF * 2 * 16 * (HD + 8) * sizeof(uint16_t); }

// three CTAs per SM with two buffers (the third needs <= 168 registers: ncu showed 169 and two resident CTAs, 1.5 waves)
template <int HD, int GROUP, int NBUF = 2>
__global__ void __launch_bounds__(kAttnThreads, NBUF == 2 ? 3 : 2) attn_decode_mma_kernel(const AttnArgs a) {
    static_assert(GROUP <= 8, "query heads per kv head are the (padded) 16 MMA rows; rows 8..15 stay empty");
    constexpr int LDS = HD + 8;                 // padded tile row (bf16): kills ldmatrix bank conflicts
    constexpr int KSTEPS = HD / 16, DBLOCKS = HD / 8, TILE = 16;
    constexpr int TILE_ELEMS = TILE * LDS;      // one K or V tile
    extern __shared__ __align__(16) uint16_t asm_tiles[];   // [warp][buf][K|V][16][LDS]; reused for the CTA merge
    __shared__ float s_m[kAttnWarps][8], s_l[kAttnWarps][8];
    __shared__ int s_last;

    pdl_launch_dependents();
    pdl_wait();

    const int split = blockIdx.x, nsplit = gridDim.x, kvh = blockIdx.y, nkv = gridDim.y, r = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    const int ctx = a.rm.positions[r] + 1;
    const int eff = max(1, min(nsplit, (ctx + 63) >> 6));
    if (split >= eff) return;
    // splits start on 16-token boundaries: with page_size % 16 == 0 a warp tile then lies inside ONE page (one block
    // table lookup per tile, rows at a fixed stride -- the per-token page arithmetic was most of the instruction count)
    // a.interleave: the 16-token tiles of the whole context are dealt round-robin to the eff x 4 warps of this (row, kv head),
    // so every warp of every split gets the same number of tiles to within one (contiguous chunks per split leave the warps
    // of a CTA with 4 or 5 tiles when the chunk is not a multiple of 64 tokens, and the last split short)
    const int chunk = (((ctx + eff - 1) / eff) + TILE - 1) / TILE * TILE;
    const int j0 = a.interleave ? 0 : split * chunk, j1 = a.interleave ? ctx : min(ctx, j0 + chunk);
    const bool page_tiles = (a.kv.page_size % TILE) == 0;
    const int32_t* bt = a.rm.block_tables + static_cast<size_t>(a.rm.slots[r]) * a.rm.max_blocks;
    uint16_t* my_tiles = asm_tiles + static_cast<size_t>(warp) * NBUF * 2 * TILE_ELEMS;

    // ---- Q fragments (A operand): row gid = query head gid of this kv head, rows 8..15 are padding ----
    uint32_t qhi[KSTEPS][2], qlo[KSTEPS][2];   // a0 (k = tig*2, +1) and a2 (k + 8) of each k-step; a1 = a3 = 0
    {
        const float* qp = a.qkv + static_cast<size_t>(r) * a.ld + (kvh * GROUP + min(gid, GROUP - 1)) * HD;
#pragma unroll
        for (int kk = 0; kk < KSTEPS; kk++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float2 v = *reinterpret_cast<const float2*>(qp + kk * 16 + h * 8 + tig * 2);
                const float x0 = gid < GROUP ? v.x * a.scale : 0.f, x1 = gid < GROUP ? v.y * a.scale : 0.f;
                const uint16_t h0 = f32_to_bf16_bits(x0), h1 = f32_to_bf16_bits(x1);
                qhi[kk][h] = static_cast<uint32_t>(h0) | (static_cast<uint32_t>(h1) << 16);
                qlo[kk][h] = pack_bf16x2(x0 - bf16_bits_to_f32(h0), x1 - bf16_bits_to_f32(h1));
            }
        }
    }

    // Page indices of this warp's tiles, 32 at a time: lane k holds the block-table entry of the warp's (32 n + k)-th tile.
    // Looking the page up when the tile is issued put one dependent L2 round trip (~0.7 us) in front of every tile's
    // cp.async, which the in-order warp could not overlap with the MMAs of the tile before.
    int my_page = 0;
    auto load_pages = [&](int jt_base, int stride_tokens) {
        const int jt = jt_base + lane * stride_tokens;
        my_page = (page_tiles && jt < j1) ? bt[jt / a.kv.page_size] : 0;
    };
    auto issue_tile = [&](int jt, int buf, int ordinal) {   // tokens [jt, jt+16) of this warp -> K and V tiles of buffer `buf`
        uint16_t* sK = my_tiles + buf * 2 * TILE_ELEMS;
        uint16_t* sV = sK + TILE_ELEMS;
        if (page_tiles) {
            const int page = __shfl_sync(0xffffffffu, my_page, ordinal & 31), off0 = jt % a.kv.page_size;
            const uint16_t* gK = a.kv.at(page, 0, off0) + kvh * HD;
            const uint16_t* gV = a.kv.at(page, 1, off0) + kvh * HD;
            const int last = j1 - 1 - jt;   // rows past the end repeat the last token (masked below)
#pragma unroll
            for (int i = lane; i < TILE * (HD / 8); i += 32) {
                const int t = i / (HD / 8), c = (i % (HD / 8)) * 8;
                const size_t goff = static_cast<size_t>(min(t, last)) * a.kv.kvd + c;
                cp_async16(smem_u32(sK + t * LDS + c), gK + goff);
                cp_async16(smem_u32(sV + t * LDS + c), gV + goff);
            }
        } else {
            for (int i = lane; i < TILE * (HD / 8); i += 32) {
                const int t = i / (HD / 8), c = (i % (HD / 8)) * 8;
                const int j = min(jt + t, j1 - 1);
                const int page = bt[j / a.kv.page_size], off = j % a.kv.page_size;
                cp_async16(smem_u32(sK + t * LDS + c), a.kv.at(page, 0, off) + kvh * HD + c);
                cp_async16(smem_u32(sV + t * LDS + c), a.kv.at(page, 1, off) + kvh * HD + c);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    float o[DBLOCKS][2];
#pragma unroll
    for (int d = 0; d < DBLOCKS; d++) o[d][0] = o[d][1] = 0.f;
    float m = -INFINITY, l = 0.f;   // running max / (quad-partial) sum of row gid

    const int first = a.interleave ? (split * kAttnWarps + warp) * TILE : j0 + warp * TILE;
    const int stride = (a.interleave ? eff : 1) * kAttnWarps * TILE;
    int buf = 0;
    // prologue: NBUF - 1 tiles in flight (one commit group per tile slot, empty groups past the end keep the count uniform)
    load_pages(first, stride);
    int issued = 0;   // ordinal of the next tile to issue
#pragma unroll
    for (int i = 0; i < NBUF - 1; i++) {
        if (first + i * stride < j1) issue_tile(first + i * stride, i, issued);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        issued++;
    }
    for (int jt = first; jt < j1; jt += stride) {
        {
            const int nxt = jt + (NBUF - 1) * stride;
            int nb_ = buf + NBUF - 1;
            if (nb_ >= NBUF) nb_ -= NBUF;
            if ((issued & 31) == 0) load_pages(nxt, stride);   // the next 32 tiles' pages (once per 512 x splits x 4 tokens)
            if (nxt < j1) issue_tile(nxt, nb_, issued);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            issued++;
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(NBUF - 1) : "memory");   // all but the NBUF - 1 newest groups: this tile has landed
        __syncwarp();
        const uint16_t* sK = my_tiles + buf * 2 * TILE_ELEMS;
        const uint16_t* sV = sK + TILE_ELEMS;

        // ---- S = Q K^T : 16 (heads) x 16 (tokens); hi and lo parts of q accumulate into the same tile ----
        // (the hi and the lo products run as separate accumulation chains, four independent chains of KSTEPS MMAs per
        // tile instead of two of 2 KSTEPS: with 2-3 warps per scheduler the dependent-issue latency of mma.sync was
        // a quarter of the stall samples)
        float s[2][4];
#pragma unroll
        for (int nb = 0; nb < 2; nb++) {
            float sl[4] = {0.f, 0.f, 0.f, 0.f};
            s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
#pragma unroll
            for (int k2 = 0; k2 < KSTEPS / 2; k2++) {
                uint32_t kf[4];
                ldmatrix_x4(smem_u32(sK + (nb * 8 + (lane & 7)) * LDS + k2 * 32 + (lane >> 3) * 8), kf);
                const uint32_t ah0[4] = {qhi[2 * k2][0], 0u, qhi[2 * k2][1], 0u}, ah1[4] = {qhi[2 * k2 + 1][0], 0u, qhi[2 * k2 + 1][1], 0u};
                const uint32_t al0[4] = {qlo[2 * k2][0], 0u, qlo[2 * k2][1], 0u}, al1[4] = {qlo[2 * k2 + 1][0], 0u, qlo[2 * k2 + 1][1], 0u};
                mma_bf16_16816(s[nb], ah0, kf[0], kf[1]);
                mma_bf16_16816(sl, al0, kf[0], kf[1]);
                mma_bf16_16816(s[nb], ah1, kf[2], kf[3]);
                mma_bf16_16816(sl, al1, kf[2], kf[3]);
            }
            s[nb][0] += sl[0];
            s[nb][1] += sl[1];
        }
        // ---- online softmax of row gid over the tile's 16 tokens (4 of them in this thread) ----
        float mx = m;
#pragma unroll
        for (int nb = 0; nb < 2; nb++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                if (jt + nb * 8 + tig * 2 + e >= j1) s[nb][e] = -INFINITY;
                mx = fmaxf(mx, s[nb][e]);
            }
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float corr = __expf(m - mx);   // first tile: exp(-inf) = 0
        m = mx;
        float p[2][2];
        float rs = 0.f;
#pragma unroll
        for (int nb = 0; nb < 2; nb++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                p[nb][e] = __expf(s[nb][e] - mx);
                rs += p[nb][e];
            }
        }
        l = l * corr + rs;
#pragma unroll
        for (int d = 0; d < DBLOCKS; d++) {
            o[d][0] *= corr;
            o[d][1] *= corr;
        }
        // ---- O += P V : k = 16 tokens; P split into bf16 hi + lo ----
        uint32_t phi[2], plo[2];
#pragma unroll
        for (int nb = 0; nb < 2; nb++) {
            const uint16_t h0 = f32_to_bf16_bits(p[nb][0]), h1 = f32_to_bf16_bits(p[nb][1]);
            phi[nb] = static_cast<uint32_t>(h0) | (static_cast<uint32_t>(h1) << 16);
            plo[nb] = pack_bf16x2(p[nb][0] - bf16_bits_to_f32(h0), p[nb][1] - bf16_bits_to_f32(h1));
        }
        const uint32_t pah[4] = {phi[0], 0u, phi[1], 0u}, pal[4] = {plo[0], 0u, plo[1], 0u};
#pragma unroll
        for (int d2 = 0; d2 < DBLOCKS / 2; d2++) {
            uint32_t vf[4];
            const int mat = lane >> 3;
            ldmatrix_x4_trans(smem_u32(sV + ((lane & 7) + (mat & 1) * 8) * LDS + d2 * 16 + (mat >> 1) * 8), vf);
            float acc0[4] = {o[2 * d2][0], o[2 * d2][1], 0.f, 0.f}, acc1[4] = {o[2 * d2 + 1][0], o[2 * d2 + 1][1], 0.f, 0.f};
            mma_bf16_16816(acc0, pah, vf[0], vf[1]);
            mma_bf16_16816(acc0, pal, vf[0], vf[1]);
            mma_bf16_16816(acc1, pah, vf[2], vf[3]);
            mma_bf16_16816(acc1, pal, vf[2], vf[3]);
            o[2 * d2][0] = acc0[0]; o[2 * d2][1] = acc0[1];
            o[2 * d2 + 1][0] = acc1[0]; o[2 * d2 + 1][1] = acc1[1];
        }
        __syncwarp();   // the tile may be overwritten by the next iteration's cp.async
        if (++buf == NBUF) buf = 0;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);

    // ---- merge the four warps of the CTA (acc through the tile memory, which nobody reads any more) ----
    __syncthreads();
    float* s_acc = reinterpret_cast<float*>(asm_tiles);   // [warp][GROUP][HD]
    if (gid < GROUP) {
        if (tig == 0) {
            s_m[warp][gid] = m;
            s_l[warp][gid] = l;
        }
#pragma unroll
        for (int d = 0; d < DBLOCKS; d++)
            *reinterpret_cast<float2*>(s_acc + (warp * GROUP + gid) * HD + d * 8 + tig * 2) = make_float2(o[d][0], o[d][1]);
    }
    __syncthreads();
    const size_t pbase = (static_cast<size_t>(r) * nkv + kvh) * nsplit + split;
    for (int e = tid; e < GROUP * HD; e += kAttnThreads) {
        const int g = e / HD, d = e % HD;
        float M = -INFINITY;
#pragma unroll
        for (int t = 0; t < kAttnWarps; t++) M = fmaxf(M, s_m[t][g]);
        float L = 0.f, A = 0.f;
        if (M > -INFINITY) {
#pragma unroll
            for (int t = 0; t < kAttnWarps; t++) {
                const float wgt = __expf(s_m[t][g] - M);   // exp(-inf) = 0 for warps without a tile
                L = fmaf(s_l[t][g], wgt, L);
                A = fmaf(s_acc[(t * GROUP + g) * HD + d], wgt, A);
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
    // The splits' (max, sum, acc) are fetched eight at a time with independent loads -- one L2 round trip per eight splits;
    // the split-at-a-time loop made the last CTA pay two dependent round trips per split (~1.4 us each) after everyone else
    // had finished. Same weights and the same fmaf order as before for <= 8 splits; chunks beyond are rescaled online.
    // (three output elements per thread go through the loads together: 72 loads in flight per thread, then the arithmetic)
    const size_t rbase = (static_cast<size_t>(r) * nkv + kvh) * nsplit;
    constexpr int NE = (GROUP * HD + kAttnThreads - 1) / kAttnThreads, EB = 3;
#pragma unroll 1
    for (int k0 = 0; k0 < NE; k0 += EB) {
        float M[EB], L[EB], A[EB];
#pragma unroll
        for (int k = 0; k < EB; k++) {
            M[k] = -INFINITY;
            L[k] = A[k] = 0.f;
        }
        for (int base = 0; base < eff; base += 8) {
            float2 ml[EB][8];
            float av[EB][8];
#pragma unroll
            for (int k = 0; k < EB; k++) {
                const int e = tid + (k0 + k) * kAttnThreads;
                const int g = min(e, GROUP * HD - 1) / HD, d = min(e, GROUP * HD - 1) % HD;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int s2 = base + i;
                    const bool ok = s2 < eff && e < GROUP * HD;
                    const size_t pg = (rbase + (s2 < eff ? s2 : 0)) * GROUP + g;
                    ml[k][i] = ok ? __ldcg(reinterpret_cast<const float2*>(a.part_ml + pg * 2)) : make_float2(-INFINITY, 0.f);
                    av[k][i] = ok ? __ldcg(a.part_acc + pg * HD + d) : 0.f;
                }
            }
#pragma unroll
            for (int k = 0; k < EB; k++) {
                float cm = M[k];
#pragma unroll
                for (int i = 0; i < 8; i++) cm = fmaxf(cm, ml[k][i].x);
                if (cm == -INFINITY) continue;
                const float keep = __expf(M[k] - cm);   // first chunk: exp(-inf) = 0 on L = A = 0
                L[k] *= keep;
                A[k] *= keep;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (ml[k][i].x == -INFINITY) continue;
                    const float w = __expf(ml[k][i].x - cm);
                    L[k] = fmaf(ml[k][i].y, w, L[k]);
                    A[k] = fmaf(av[k][i], w, A[k]);
                }
                M[k] = cm;
            }
        }
#pragma unroll
        for (int k = 0; k < EB; k++) {
            const int e = tid + (k0 + k) * kAttnThreads;
            if (k0 + k < NE && e < GROUP * HD) attn_store_out(a, r, (kvh * GROUP + e / HD) * HD + e % HD, A[k] / L[k]);
        }
    }
}

}  // namespace b2l

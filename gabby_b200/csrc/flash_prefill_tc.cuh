// flash_prefill_tc.cuh -- causal GQA prefill attention on the 5th-generation tensor cores (tcgen05 / TMEM), over the paged
// bf16 KV cache. Replaces flash_prefill.cuh (warp-level mma.sync, 190 TFLOP/s) for head_dim 64 / 128.
//
// One CTA = 256 consecutive query positions of one sequence x one query head, as TWO 128-row query tiles that ping-pong:
// while the softmax warps of one tile work on S, the tensor core runs the other tile's MMAs.
//   warp 0     TMA: K and V tiles of 128 tokens, gathered page by page from the cache (one 2-D tensor map over the whole
//              pool; a box = one page x 64 head dims, 128-byte swizzle) into two 2-stage rings
//   warp 1     one thread issues every MMA:  S_t = Q_t K^T   (A = Q from shared memory, B = K, both K-major; D = S in TMEM)
//                                            O_t += P_t V    (A = P straight from TMEM, B = V in its natural [token][dim]
//                                                             layout = MN-major; D = O in TMEM)
//   warps 2-5  softmax of query tile 0 (one row per thread: no shuffles), warps 6-9 of query tile 1:
//              tcgen05.ld S -> scale, causal mask, running max -> P = exp2(s - m) as bf16 written back over S
//              (tcgen05.st), O rescaled in TMEM when the row maximum moved, at the end O / l -> bf16 -> global
// TMEM: S0 | S1 (128 fp32 columns each; P aliases the first 64 columns of its S) | O0 | O1 (head_dim columns each).
// q: fp32 [T][ld] straight from the QKV GEMM: rotated (RoPE) and rounded to bf16 while it is staged; out: bf16 [T][ldo].
// Math = oracle attention with ORC_KV_BF16 | ORC_QP_BF16 (same rounding points as flash_prefill.cuh).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "flash_prefill.cuh"   // PrefillTile
#include "gemm_tcgen05.cuh"    // descriptor / mbarrier / tcgen05 helpers

namespace b2l {

struct FlashTcArgs {
    const float* qkv;            // [T][ld] fp32, q heads first, NOT rotated
    int ld;
    const float* rope;           // [max_positions][HD/2][2] (cos, sin)
    const int32_t* block_tables; // [n_slots][max_blocks]
    int max_blocks;
    const PrefillTile* tiles;    // up to 256 rows each
    uint16_t* out;               // [T][ldo] bf16
    int ldo;
    int group;                   // query heads per kv head
    float scale_log2e;
    int page_size;               // divides 128, multiple of 8
    long long layer_row0;        // first row of this layer's pool in the tensor map ([layer][page][K|V][slot] rows of kvd elements)
};

constexpr int kFtcThreads = 320, kFtcBQ = 128, kFtcBKV = 128;

template <int HD>
struct FtcSmem {
    static constexpr uint32_t kPanels = HD / 64;                    // 64-element (128-byte) wide panels
    static constexpr uint32_t kPanelBytes = 128 * 128;              // 128 rows x 128 B
    static constexpr uint32_t kTileBytes = kPanels * kPanelBytes;   // one 128-row tile of Q, K or V
    static constexpr uint32_t q = 0, k = 2 * kTileBytes, v = 4 * kTileBytes, bars = 6 * kTileBytes;
    static constexpr size_t total = 6 * kTileBytes + 256 + 1024;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major B operand (V as [token][dim]), SWIZZLE_128B: 8 x 16-byte pieces = 64 dims contiguous, the next 64 dims one panel
// (leading byte offset) further, 8 tokens = 1024 B (stride byte offset)
__device__ __forceinline__ uint64_t umma_smem_desc_mn(uint32_t smem_addr, uint32_t panel_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(panel_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int HD>
__global__ void __launch_bounds__(kFtcThreads, 1) flash_prefill_tc_kernel(const __grid_constant__ CUtensorMap map_kv, const FlashTcArgs a) {
    using SM = FtcSmem<HD>;
    constexpr uint32_t kPanels = SM::kPanels, kPanelBytes = SM::kPanelBytes, kTileBytes = SM::kTileBytes;
    constexpr uint32_t kColS = 0, kColO = 256;   // TMEM columns: S_t at 128 t, O_t at 256 + HD t
    extern __shared__ __align__(1024) uint8_t fsm_tc[];
    const uint32_t base = (smem_u32(fsm_tc) + 1023u) & ~1023u;
    const uint32_t sQ = base + SM::q, sK = base + SM::k, sV = base + SM::v, bars = base + SM::bars;
    const uint32_t k_full = bars, k_empty = bars + 16, v_full = bars + 32, v_empty = bars + 48, s_full = bars + 64, p_ready = bars + 80,
                   o_done = bars + 96, tmem_slot = bars + 112;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const PrefillTile tile = a.tiles[blockIdx.x];
    const int head = blockIdx.y, kvh = head / a.group;
    // kv tiles each query tile needs: tokens [0, position of its last row]
    const int rows0 = min(tile.n_rows, kFtcBQ), rows1 = max(tile.n_rows - kFtcBQ, 0);
    const int n0 = (tile.pos0 + rows0 + kFtcBKV - 1) / kFtcBKV;
    const int n1 = rows1 > 0 ? (tile.pos0 + kFtcBQ + rows1 + kFtcBKV - 1) / kFtcBKV : 0;
    const int nmax = max(n0, n1);

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(k_full + 8 * i, 1); mbar_init(k_empty + 8 * i, 1);
            mbar_init(v_full + 8 * i, 1); mbar_init(v_empty + 8 * i, 1);
            mbar_init(s_full + 8 * i, 1); mbar_init(p_ready + 8 * i, 128); mbar_init(o_done + 8 * i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- Q: fp32 -> RoPE -> bf16, K-major panels with the 128-byte swizzle the MMA descriptors expect ----
    // RoPE happens here (rope_kv_kernel's arithmetic): an item = 8 dims of the lower half of a row + their partners in the upper
    // half. The loads of a batch are all in flight together (one load per round trip made this prologue a fifth of the
    // kernel's time).
    constexpr int kQItems = 2 * kFtcBQ * (HD / 16), kQBatch = 4;
#pragma unroll 1
    for (int i0 = tid; i0 < kQItems; i0 += kQBatch * kFtcThreads) {
        float4 lo[kQBatch][2], hi[kQBatch][2], cs[kQBatch][4];
#pragma unroll
        for (int b = 0; b < kQBatch; b++) {
            const int i = min(i0 + b * kFtcThreads, kQItems - 1);
            const int r = i / (HD / 16), c8 = i % (HD / 16);                 // row 0..255, 8-dim piece of the lower half
            const int rc = min(r, tile.n_rows - 1);                          // rows past the tile repeat the last row
            const float4* src = reinterpret_cast<const float4*>(a.qkv + static_cast<size_t>(tile.row0 + rc) * a.ld + head * HD + c8 * 8);
            const float4* csp = reinterpret_cast<const float4*>(a.rope + static_cast<size_t>(tile.pos0 + rc) * HD + c8 * 16);
            lo[b][0] = ldg_f4_early(src); lo[b][1] = ldg_f4_early(src + 1);
            hi[b][0] = ldg_f4_early(src + HD / 8); hi[b][1] = ldg_f4_early(src + HD / 8 + 1);
#pragma unroll
            for (int k = 0; k < 4; k++) cs[b][k] = ldg_f4_early(csp + k);
        }
#pragma unroll
        for (int b = 0; b < kQBatch; b++) {
            const int i = i0 + b * kFtcThreads;
            if (i < kQItems) {
                const uint32_t r = i / (HD / 16), c8 = i % (HD / 16);
                const float x0[8] = {lo[b][0].x, lo[b][0].y, lo[b][0].z, lo[b][0].w, lo[b][1].x, lo[b][1].y, lo[b][1].z, lo[b][1].w};
                const float x1[8] = {hi[b][0].x, hi[b][0].y, hi[b][0].z, hi[b][0].w, hi[b][1].x, hi[b][1].y, hi[b][1].z, hi[b][1].w};
                const float cc[8] = {cs[b][0].x, cs[b][0].z, cs[b][1].x, cs[b][1].z, cs[b][2].x, cs[b][2].z, cs[b][3].x, cs[b][3].z};
                const float ss[8] = {cs[b][0].y, cs[b][0].w, cs[b][1].y, cs[b][1].w, cs[b][2].y, cs[b][2].w, cs[b][3].y, cs[b][3].w};
                float y0[8], y1[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    y0[e] = x0[e] * cc[e] - x1[e] * ss[e];
                    y1[e] = x1[e] * cc[e] + x0[e] * ss[e];
                }
                const uint32_t t = r >> 7, rr = r & 127;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const float* y = h ? y1 : y0;
                    const uint32_t c8h = c8 + h * (HD / 16), panel = c8h >> 3, piece = c8h & 7;
                    const uint32_t dst = sQ + t * kTileBytes + panel * kPanelBytes + rr * 128 + ((piece ^ (rr & 7)) << 4);
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(pack_bf16x2(y[0], y[1])), "r"(pack_bf16x2(y[2], y[3])),
                                 "r"(pack_bf16x2(y[4], y[5])), "r"(pack_bf16x2(y[6], y[7])) : "memory");
                }
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core's async reads
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ---- TMA: K_j and V_j, page by page. Lane pg owns page pg of the tile: it reads the block-table entry one tile ahead
        // (a dependent global load per page in front of every copy kept the whole CTA waiting for its K/V) and issues that
        // page's copies; lane 0 does the barrier work ----
        const int32_t* bt = a.block_tables + static_cast<size_t>(tile.slot) * a.max_blocks;
        const int ps = a.page_size, pages_per_tile = kFtcBKV / ps;   // <= 16
        const int last_page = (tile.pos0 + tile.n_rows - 1) / ps;    // pages past it are not allocated: re-read the last one (masked)
        const bool owner = lane < pages_per_tile;
        int page_next = owner ? bt[min(lane, last_page)] : 0;
        for (int j = 0; j < nmax; j++) {
            const uint32_t s = j & 1, parity = (j >> 1) & 1;
            const int page = page_next;
            if (owner && j + 1 < nmax) page_next = bt[min((j + 1) * pages_per_tile + lane, last_page)];
            for (int which = 0; which < 2; which++) {
                const uint32_t full = (which ? v_full : k_full) + 8 * s, empty = (which ? v_empty : k_empty) + 8 * s;
                const uint32_t dst0 = (which ? sV : sK) + s * kTileBytes;
                if (lane == 0) {
                    mbar_wait_spin(empty, parity ^ 1);
                    mbar_arrive_expect_tx(full, kTileBytes);
                }
                __syncwarp();
                if (owner) {
                    const int row = static_cast<int>(a.layer_row0 + (static_cast<long long>(page) * 2 + which) * ps);
#pragma unroll
                    for (uint32_t pn = 0; pn < kPanels; pn++)
                        tma_load_2d(dst0 + pn * kPanelBytes + lane * ps * 128, &map_kv, kvh * HD + pn * 64, row, full);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer ----
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, kFtcBKV);
            constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD) | (1u << 16);   // B (= V) is MN-major
            const int nt[2] = {n0, n1};
            auto issue_s = [&](int t, int j) {   // S_t = Q_t K_j^T
                const uint32_t kb = sK + (j & 1) * kTileBytes, qb = sQ + t * kTileBytes;
#pragma unroll
                for (int k = 0; k < HD / 16; k++) {
                    const uint32_t off = (k >> 2) * kPanelBytes + (k & 3) * 32;
                    umma_bf16_ss(tmem_base + kColS + t * 128, umma_smem_desc(qb + off), umma_smem_desc(kb + off), idesc_s, k != 0);
                }
                tcgen05_commit(s_full + 8 * t);
            };
            mbar_wait_spin(k_full, 0);
            tcgen05_fence_after();
            for (int t = 0; t < 2; t++)
                if (nt[t] > 0) issue_s(t, 0);
            tcgen05_commit(k_empty);
            for (int j = 0; j < nmax; j++) {
                const uint32_t s = j & 1, parity = (j >> 1) & 1;
                bool v_ok = false, k_ok = false;
                for (int t = 0; t < 2; t++) {
                    if (j >= nt[t]) continue;
                    mbar_wait_spin(p_ready + 8 * t, j & 1);           // P_t(j) is in TMEM, O_t is rescaled
                    if (!v_ok) { mbar_wait_spin(v_full + 8 * s, parity); v_ok = true; }
                    tcgen05_fence_after();
                    const uint32_t vb = sV + s * kTileBytes;
#pragma unroll
                    for (int k = 0; k < kFtcBKV / 16; k++)             // 16 tokens per MMA: 8 TMEM columns of P, 2048 B of V
                        umma_bf16_ts(tmem_base + kColO + t * HD, tmem_base + kColS + t * 128 + k * 8, umma_smem_desc_mn(vb + k * 2048, kPanelBytes),
                                     idesc_o, (j | k) != 0);
                    tcgen05_commit(o_done + 8 * t);
                    if (j + 1 < nt[t]) {
                        if (!k_ok) { mbar_wait_spin(k_full + 8 * (s ^ 1), ((j + 1) >> 1) & 1); k_ok = true; tcgen05_fence_after(); }
                        issue_s(t, j + 1);
                    }
                }
                tcgen05_commit(v_empty + 8 * s);
                if (j + 1 < nmax) tcgen05_commit(k_empty + 8 * (s ^ 1));
            }
        }
    } else {
        // ---- softmax: query tile t, row r = TMEM lane ----
        const int t = (warp - 2) >> 2, quarter = warp & 3, r = quarter * 32 + lane;
        const int n_t = t == 0 ? n0 : n1, rows_t = t == 0 ? rows0 : rows1;
        if (n_t > 0) {
            const int qpos = tile.pos0 + t * kFtcBQ + min(r, rows_t - 1);   // padding rows take the last row's mask
            const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
            const uint32_t tS = lane_base + kColS + t * 128, tO = lane_base + kColO + t * HD;
            float m = -INFINITY, l = 0.f;
            for (int j = 0; j < n_t; j++) {
                mbar_wait_spin(s_full + 8 * t, j & 1);
                tcgen05_fence_after();
                const int tok0 = j * kFtcBKV;
                // some token of the tile is masked for some row of this warp (warp-uniform: the TMEM loads below are .aligned)
                const bool diag = tok0 + kFtcBKV - 1 > tile.pos0 + t * kFtcBQ + min(quarter * 32, rows_t - 1);
                // the whole S row in registers: four loads in flight, one wait
                uint32_t s[128];
#pragma unroll
                for (int c = 0; c < 4; c++) tmem_ld32_nowait(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(s + c * 32));
                tmem_wait_ld();
                float mx = -INFINITY;
                if (diag) {
#pragma unroll
                    for (int i = 0; i < 128; i++) {
                        if (tok0 + i > qpos) s[i] = 0xff800000u;   // -inf
                        mx = fmaxf(mx, __uint_as_float(s[i]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 128; i++) mx = fmaxf(mx, __uint_as_float(s[i]));
                }
                // Lazy rescaling: the exponent's reference m only follows the running maximum when some row of the warp has
                // outgrown it by more than 2^8 (P <= 256 stays exact enough in bf16 / fp32; O / l at the end is unchanged
                // mathematically). Most tiles then skip the round trip through O in TMEM and the wait for the previous P V.
                const float mx2 = mx * a.scale_log2e;   // finite in tile 0: every row sees token 0
                const bool grow = __any_sync(0xffffffffu, mx2 > m + 8.0f);
                const float m_new = grow ? fmaxf(m, mx2) : m;
                const float alpha = ex2_approx(m - m_new);   // 1 when the reference stays (0 in tile 0: l = 0 there)
                // P = exp2(s - m) -> bf16 over the first 64 columns of S (all of S has been read)
                float rs = 0.f;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t p[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(s[c * 32 + 2 * i]), a.scale_log2e, -m_new));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(s[c * 32 + 2 * i + 1]), a.scale_log2e, -m_new));
                        rs += p0 + p1;
                        p[i] = pack_bf16x2(p0, p1);
                    }
                    tmem_st16(tS + c * 16, p);
                }
                l = l * alpha + rs;
                m = m_new;
                // O_t *= alpha, after the previous tile's P V has landed
                if (j > 0 && grow) {
                    mbar_wait_spin(o_done + 8 * t, (j - 1) & 1);
                    tcgen05_fence_after();
                    uint32_t o[HD];
#pragma unroll
                    for (int c = 0; c < HD / 32; c++) tmem_ld32_nowait(tO + c * 32, *reinterpret_cast<uint32_t(*)[32]>(o + c * 32));
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < HD; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
#pragma unroll
                    for (int c = 0; c < HD / 32; c++) tmem_st32(tO + c * 32, *reinterpret_cast<const uint32_t(*)[32]>(o + c * 32));
                }
                tmem_wait_st();
                tcgen05_fence_before();
                mbar_arrive(p_ready + 8 * t);
            }
            // ---- O / l -> bf16 -> global ----
            mbar_wait_spin(o_done + 8 * t, (n_t - 1) & 1);
            tcgen05_fence_after();
            const float inv = 1.0f / l;
            const bool live = r < rows_t;
            uint16_t* dst = a.out + static_cast<size_t>(tile.row0 + t * kFtcBQ + (live ? r : 0)) * a.ldo + head * HD;
#pragma unroll 1
            for (int c = 0; c < HD / 32; c++) {
                uint32_t o[32];
                tmem_ld32(tO + c * 32, o);
                if (live) {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        reinterpret_cast<uint4*>(dst + c * 32)[i] =
                            make_uint4(pack_bf16x2(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv),
                                       pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv),
                                       pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv),
                                       pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv));
                }
            }
            tcgen05_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace b2l

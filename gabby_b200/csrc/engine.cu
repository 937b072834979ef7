// engine.cu -- the b2l C-ABI (include/b2l.h): device context, weight residency, paged KV pool,
// and the forward pass enqueued as hand-written sm_100a kernels. No CPU fallback anywhere:
// without a CUDA device b2l_create fails.
//
// Reference boundary: this is the missing body of gabby::inference::Llama3Generator::{Load,
// Generate} (/root/reference/src/inference/generator.cc:33-44); the host C++ layer
// (gabby_b200/host/) adapts it to the Generator interface.
#include "engine.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <functional>
#include <set>
#include <utility>

#include "decode_kernels.cuh"
#include "mega_common.cuh"
// the decode megakernel is compiled twice from one source: the single-GPU kernel and the kernel of a tensor-parallel rank
#define MEGA_TP 0
#define MEGA_NS mega1
#include "mega_decode.cuh"
#undef MEGA_TP
#undef MEGA_NS
#define MEGA_TP 1
#define MEGA_NS megatp
#include "mega_decode.cuh"
#undef MEGA_TP
#undef MEGA_NS
#include "gemm_tcgen05.cuh"
#include "flash_prefill.cuh"
#include "flash_prefill_tc.cuh"
#include "skinny_gemm.cuh"
#include "attn_decode_mma.cuh"
#include "synth.cuh"

using namespace b2l;

namespace {

thread_local std::string g_create_error;

// ---- NCCL through dlopen: no link-time dependency, and when the caller already loaded a libnccl.so.2
// (torch ships one) the same copy is reused. Only the handful of entry points TP decode needs. ----
struct NcclApi {
    using Id = struct { char internal[B2L_NCCL_ID_BYTES]; };
    int (*GetUniqueId)(Id*) = nullptr;
    int (*CommInitRank)(void**, int, Id, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        api.why = std::string("cannot load libnccl.so.2: ") + dlerror();
        return api;
    }
    auto sym = [&](const char* n) { return dlsym(h, n); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.GetErrorString;
    if (!api.ok) api.why = "libnccl.so.2 lacks an expected symbol";
    return api;
}

#define B2L_NCCL(expr)                                                                             \
    do {                                                                                           \
        int _r = (expr);                                                                           \
        if (_r != 0) throw ::b2l::Error(std::string(#expr) + ": " + nccl().GetErrorString(_r));   \
    } while (0)

// ---- tensor naming (HF names as they appear in the safetensors header) --------------------
enum Kind { K_EMBED, K_FINAL_NORM, K_LM_HEAD, K_IN_NORM, K_Q, K_K, K_V, K_O, K_POST_NORM, K_GATE, K_UP, K_DOWN, K_BAD };

Kind parse_name(const std::string& n, int* layer) {
    *layer = -1;
    if (n == "model.embed_tokens.weight") return K_EMBED;
    if (n == "model.norm.weight") return K_FINAL_NORM;
    if (n == "lm_head.weight") return K_LM_HEAD;
    const std::string pre = "model.layers.";
    if (n.compare(0, pre.size(), pre) != 0) return K_BAD;
    const size_t dot = n.find('.', pre.size());
    if (dot == std::string::npos) return K_BAD;
    *layer = std::atoi(n.substr(pre.size(), dot - pre.size()).c_str());
    const std::string rest = n.substr(dot + 1);
    if (rest == "input_layernorm.weight") return K_IN_NORM;
    if (rest == "self_attn.q_proj.weight") return K_Q;
    if (rest == "self_attn.k_proj.weight") return K_K;
    if (rest == "self_attn.v_proj.weight") return K_V;
    if (rest == "self_attn.o_proj.weight") return K_O;
    if (rest == "post_attention_layernorm.weight") return K_POST_NORM;
    if (rest == "mlp.gate_proj.weight") return K_GATE;
    if (rest == "mlp.up_proj.weight") return K_UP;
    if (rest == "mlp.down_proj.weight") return K_DOWN;
    return K_BAD;
}

// where a (row, col) window of the full HF tensor lands in this rank's layout
struct Placement {
    uint16_t* dst;
    int64_t dst_stride;  // elements between destination rows
    int64_t row0, nrows, col0, ncols;
    int64_t full_rows, full_cols;
};

template <typename T>
T* dalloc(b2l_ctx* c, size_t n) {
    void* p = nullptr;
    B2L_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    c->allocs.push_back(p);
    return static_cast<T*>(p);
}

// The opt-in to more than 48 KB of dynamic shared memory is a per-DEVICE function attribute: remember which
// (kernel, device) pairs have it, so that a context on a second GPU of the same process gets it too.
void ensure_smem_optin(const void* kern, int device, size_t bytes, bool max_carveout = false) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({kern, device})) return;
    B2L_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    // several CTAs per SM only fit when the SM is carved for shared memory (the driver's default guess can stop one short)
    if (max_carveout) B2L_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    done.insert({kern, device});
}

// The megakernel's launch arguments live in one __constant__ symbol per device (mega_decode.cuh: c_mega) and the
// persistent kernel reads them for as long as it runs: contexts that share a GPU take this lock from the argument
// upload until their launch has finished (the kernel occupies every SM anyway, so nothing is lost).
std::mutex& mega_device_mutex(int device) {
    static std::mutex table_mu;
    static std::map<int, std::mutex> per_device;
    std::lock_guard<std::mutex> lock(table_mu);
    return per_device[device];
}

Placement place_tensor(b2l_ctx* c, Kind k, int layer, const int64_t* shape, int ndim) {
    const b2l_params& p = c->p;
    const int64_t H = c->H, r = p.tp_rank;
    const int64_t qd = static_cast<int64_t>(p.num_heads) * p.head_dim, kvd = static_cast<int64_t>(p.num_kv_heads) * p.head_dim;
    auto want = [&](int64_t rows, int64_t cols) {
        if (cols == 1 ? !(ndim == 1 && shape[0] == rows) : !(ndim == 2 && shape[0] == rows && shape[1] == cols))
            throw Error("tensor shape mismatch");
    };
    if (k >= K_IN_NORM) B2L_CHECK(layer >= 0 && layer < c->L, "layer index out of range");
    LayerWeights* lw = k >= K_IN_NORM ? &c->layers[layer] : nullptr;
    switch (k) {
        case K_EMBED: want(p.vocab_size, H); return {c->embed, H, 0, p.vocab_size, 0, H, p.vocab_size, H};
        case K_LM_HEAD:
            B2L_CHECK(!p.tie_word_embeddings, "lm_head.weight given for a tied model");
            want(p.vocab_size, H);
            return {c->lm_head, H, r * c->V_l, c->V_l, 0, H, p.vocab_size, H};
        case K_FINAL_NORM: want(H, 1); return {c->final_norm, H, 0, 1, 0, H, 1, H};
        case K_IN_NORM: want(H, 1); return {lw->in_norm, H, 0, 1, 0, H, 1, H};
        case K_POST_NORM: want(H, 1); return {lw->post_norm, H, 0, 1, 0, H, 1, H};
        case K_Q: want(qd, H); return {lw->w_qkv, H, r * c->qd_l, c->qd_l, 0, H, qd, H};
        case K_K: want(kvd, H); return {lw->w_qkv + static_cast<int64_t>(c->qd_l) * H, H, r * c->kvd_l, c->kvd_l, 0, H, kvd, H};
        case K_V: want(kvd, H); return {lw->w_qkv + static_cast<int64_t>(c->qd_l + c->kvd_l) * H, H, r * c->kvd_l, c->kvd_l, 0, H, kvd, H};
        case K_O: want(H, qd); return {lw->w_o, c->qd_l, 0, H, r * c->qd_l, c->qd_l, H, qd};
        case K_GATE: want(p.intermediate_size, H); return {lw->w_gu, 2 * H, r * c->I_l, c->I_l, 0, H, p.intermediate_size, H};
        case K_UP: want(p.intermediate_size, H); return {lw->w_gu + H, 2 * H, r * c->I_l, c->I_l, 0, H, p.intermediate_size, H};
        case K_DOWN: want(H, p.intermediate_size); return {lw->w_down, c->I_l, 0, H, r * c->I_l, c->I_l, H, p.intermediate_size};
        default: throw Error("unknown tensor name");
    }
}

void mark_tensor(b2l_ctx* c, Kind k, int layer) {
    switch (k) {
        case K_EMBED:
            c->have_embed = true;
            if (c->p.tie_word_embeddings) c->have_lm_head = true;
            break;
        case K_LM_HEAD: c->have_lm_head = true; break;
        case K_FINAL_NORM: c->have_final_norm = true; break;
        default: c->layers[layer].have |= 1u << k;
    }
}

// ---- launches ----------------------------------------------------------------------------
template <typename... KArgs, typename... Args>
void launch(b2l_ctx* c, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B2L_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
    c->launched++;
}

// the same with a thread-block cluster of (cluster_x, 1, 1)
template <typename... KArgs, typename... Args>
void launch_cluster(b2l_ctx* c, void (*kernel)(KArgs...), dim3 grid, dim3 block, unsigned cluster_x, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    B2L_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
    c->launched++;
}

constexpr size_t kGemvSmemBudget = 32 * 1024;  // x tile per CTA: small enough for 6+ CTAs per SM at batch 8

int gemv_kt(int B, int K) {
    const int cap = static_cast<int>(kGemvSmemBudget / (4 * B)) / 256 * 256;
    return std::min(K, cap);
}

template <int B, int MODE, bool NORM>
void gemv_launch_b(b2l_ctx* c, const GemvArgs& a) {
    auto kern = gemv_kernel<B, MODE, NORM>;
    ensure_smem_optin(reinterpret_cast<const void*>(kern), c->p.device, kGemvSmemBudget);
    const int n_blocks = (a.N + kGemvRowsPerCta - 1) / kGemvRowsPerCta;
    const int grid = std::min(n_blocks, c->prop.multiProcessorCount * 8);
    launch(c, kern, dim3(grid), dim3(kGemvThreads), static_cast<size_t>(B) * a.kt * sizeof(float), a);
}

template <int MODE, bool NORM>
void gemv_launch(b2l_ctx* c, GemvArgs a, int R) {
    // rows in groups of <= 8; a group of 3 runs as 4 etc. (scratch buffers are sized in 8-row units)
    for (int r0 = 0; r0 < R; r0 += 8) {
        const int n = std::min(8, R - r0);
        GemvArgs g = a;
        g.x = a.x + static_cast<size_t>(r0) * a.ldx;
        g.y = a.y + static_cast<size_t>(r0) * a.ldy;
        if (MODE == 3)
            for (int p = 0; p < a.tps.tp; p++) g.tps.dst[p] = a.tps.dst[p] + static_cast<size_t>(r0) * a.ldy;
        const int B = n <= 1 ? 1 : n <= 2 ? 2 : n <= 4 ? 4 : 8;
        g.kt = gemv_kt(B, a.K);
        switch (B) {
            case 1: gemv_launch_b<1, MODE, NORM>(c, g); break;
            case 2: gemv_launch_b<2, MODE, NORM>(c, g); break;
            case 4: gemv_launch_b<4, MODE, NORM>(c, g); break;
            default: gemv_launch_b<8, MODE, NORM>(c, g); break;
        }
    }
}

void gemv(b2l_ctx* c, const uint16_t* W, const float* x, int ldx, float* y, int ldy, const uint16_t* norm_w, int N,
          int K, int mode, int R, const TpSend* tps = nullptr) {
    B2L_CHECK(K % 8 == 0 && N % 2 == 0, "gemv: K must be a multiple of 8 and N even");
    GemvArgs a{W, x, y, norm_w, nullptr, c->p.rms_norm_eps, N, K, ldx, ldy, 0, TpSend{}};
    if (mode == 3) {
        B2L_CHECK(tps && !norm_w, "gemv: TP send needs peers and no fused norm");
        a.tps = *tps;
        gemv_launch<3, false>(c, a, R);
        return;
    }
    if (norm_w) {
        if (mode == 0) gemv_launch<0, true>(c, a, R);
        else if (mode == 2) gemv_launch<2, true>(c, a, R);
        else throw Error("gemv: unsupported fused-norm mode");
    } else {
        if (mode == 0) gemv_launch<0, false>(c, a, R);
        else if (mode == 1) gemv_launch<1, false>(c, a, R);
        else if (mode == 2) gemv_launch<2, false>(c, a, R);
        else throw Error("gemv: bad mode");
    }
}

// tile buffers per warp of the tensor-core decode attention: 3 = two tiles in flight (ncu, 3B batch 8 context 2048, 4 splits:
// 24.7 us per launch against 26.2 with 2 buffers; 6-7 splits 28.6-32 us -- profiles/r02_attn_decode_sweeps.txt)
constexpr int kAttnNbufDefault = 3;

template <int HD>
void attn_launch_hd(b2l_ctx* c, const AttnArgs& a, int R) {
    // about four 128-thread CTAs per SM when the context is long (measured in one run, 3B batch 8, context 2048:
    // 8 splits 3.67 ms/step, 16: 4.00, 32: 4.54 -- the last-arriver merge grows with the split count); the kernel
    // trims the split count for short contexts
    // (8B, 1 row, context 4096: 16 splits 5.36 ms/token, 32: 5.07, 64: 5.41)
    const int want = std::min(32, (3 * c->prop.multiProcessorCount + c->nkv_l * R - 1) / (c->nkv_l * R));
    static const int forced = std::getenv("B2L_ATTN_SPLITS") ? std::atoi(std::getenv("B2L_ATTN_SPLITS")) : 0;   // tuning knob
    const dim3 grid(std::max(1, std::min(c->nsplit, forced > 0 ? forced : want)), c->nkv_l, R), block(kAttnThreads);
    static const bool use_mma = !(std::getenv("B2L_ATTN_MMA") && std::atoi(std::getenv("B2L_ATTN_MMA")) == 0);
    if constexpr (HD >= 64) {
        if (use_mma) {   // tensor-core kernel (attn_decode_mma.cuh); B2L_ATTN_MMA=0 selects the CUDA-core kernel
            // tuning knobs: B2L_ATTN_NBUF=3: two tiles in flight per warp (2 CTAs per SM at head_dim 128 instead of 3);
            // B2L_ATTN_INTERLEAVE=0: contiguous context chunks per split
            static const int nbuf = std::getenv("B2L_ATTN_NBUF") ? std::atoi(std::getenv("B2L_ATTN_NBUF")) : kAttnNbufDefault;
            static const bool interleave = !(std::getenv("B2L_ATTN_INTERLEAVE") && std::atoi(std::getenv("B2L_ATTN_INTERLEAVE")) == 0);
            AttnArgs ai = a;
            ai.interleave = interleave ? 1 : 0;
            auto go = [&](auto kern, size_t smem) {
                ensure_smem_optin(reinterpret_cast<const void*>(kern), c->p.device, smem, true);
                // one wave: as many context splits as keep every CTA resident at once. ncu showed the old fixed guess (3 CTAs per
                // SM -> 7 splits at 3B batch 8 = 448 CTAs) running as 1.5 waves of the 2 CTAs per SM that really fit; every CTA
                // also pays ~6 us of fixed latency (block table -> first tile, CTA merge, fence + last-arriver merge), so fewer,
                // longer CTAs win: 4 splits 2.58-2.63 ms/step, 7 splits 2.83-2.91
                int occ = 0;
                B2L_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kAttnThreads, smem));
                const int one_wave = std::max(1, std::min(32, occ * c->prop.multiProcessorCount / (c->nkv_l * R)));
                const dim3 g(std::max(1, std::min(c->nsplit, forced > 0 ? forced : one_wave)), c->nkv_l, R);
                launch(c, kern, g, block, smem, ai);
            };
#define B2L_ATTN_GO(G)                                                                                   \
    if (nbuf == 3) go(attn_decode_mma_kernel<HD, G, 3>, attn_mma_smem<HD, 3>());                             \
    else go(attn_decode_mma_kernel<HD, G, 2>, attn_mma_smem<HD, 2>());                                       \
    return;
            switch (c->group) {
                case 1: B2L_ATTN_GO(1)
                case 2: B2L_ATTN_GO(2)
                case 3: B2L_ATTN_GO(3)
                case 4: B2L_ATTN_GO(4)
                case 8: B2L_ATTN_GO(8)
                default: throw Error("unsupported GQA group size (heads per kv head must be 1,2,3,4 or 8)");
            }
#undef B2L_ATTN_GO
        }
    }
    switch (c->group) {
        case 1: launch(c, attn_decode_kernel<HD, 1>, grid, block, 0, a); break;
        case 2: launch(c, attn_decode_kernel<HD, 2>, grid, block, 0, a); break;
        case 3: launch(c, attn_decode_kernel<HD, 3>, grid, block, 0, a); break;
        case 4: launch(c, attn_decode_kernel<HD, 4>, grid, block, 0, a); break;
        case 8: launch(c, attn_decode_kernel<HD, 8>, grid, block, 0, a); break;
        default: throw Error("unsupported GQA group size (heads per kv head must be 1,2,3,4 or 8)");
    }
}

void attn_launch(b2l_ctx* c, const AttnArgs& a, int R) {
    switch (c->hd) {
        case 32: attn_launch_hd<32>(c, a, R); break;
        case 64: attn_launch_hd<64>(c, a, R); break;
        case 128: attn_launch_hd<128>(c, a, R); break;
        default: throw Error("unsupported head_dim (32, 64 or 128)");
    }
}

void tap_copy(b2l_ctx* c, int slab, int row0, const float* src, int R) {
    float* dst = c->tap + (static_cast<size_t>(slab) * c->tap_rows_cap + row0) * c->H;
    B2L_CUDA(cudaMemcpyAsync(dst, src, sizeof(float) * R * c->H, cudaMemcpyDeviceToDevice, c->stream));
}

CUtensorMap make_kmajor_map(const uint16_t* ptr, int64_t rows, int64_t cols, int box_rows);  // defined below

// ---- batched decode (2..16 rows) on the tensor cores -----------------------------------------------
void skinny_setup(b2l_ctx* c) {
    c->skinny_ok = false;
    // K must be a whole number of 64-element TMA boxes; N only has to be even (TMA zero-fills the rows past N)
    const bool shapes = c->H % 64 == 0 && c->qd_l % 64 == 0 && c->I_l % 64 == 0 && c->qkv_l % 2 == 0 && c->V_l % 2 == 0 &&
                        c->qkv_l >= 128 && c->H >= 128 && c->p.max_batch >= 2;
    if (!shapes) return;
    c->sk_xh = dalloc<uint16_t>(c, static_cast<size_t>(64) * c->H);
    c->sk_xq = dalloc<uint16_t>(c, static_cast<size_t>(64) * c->qd_l);
    c->sk_xi = dalloc<uint16_t>(c, static_cast<size_t>(64) * c->I_l);
    B2L_CUDA(cudaMemset(c->sk_xh, 0, sizeof(uint16_t) * 64 * c->H));
    B2L_CUDA(cudaMemset(c->sk_xq, 0, sizeof(uint16_t) * 64 * c->qd_l));
    B2L_CUDA(cudaMemset(c->sk_xi, 0, sizeof(uint16_t) * 64 * c->I_l));
    const size_t max_n = std::max<size_t>(static_cast<size_t>(c->V_l), static_cast<size_t>(2) * c->I_l);
    c->sk_partial_floats = std::max<size_t>(max_n * 32, static_cast<size_t>(32) * 32 * 8192);
    c->sk_partial = dalloc<float>(c, c->sk_partial_floats);
    B2L_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(skinny_smem(16))));
    B2L_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(skinny_smem(32))));
    B2L_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(skinny_smem(64))));
    c->skinny_ok = true;
    // fused epilogues of the batched path (on by default; the switches are for A/B measurements)
    c->rope_fuse = !(std::getenv("B2L_ROPE_FUSE") && std::atoi(std::getenv("B2L_ROPE_FUSE")) == 0);
    c->norm_cluster = !(std::getenv("B2L_NORM_CLUSTER") && std::atoi(std::getenv("B2L_NORM_CLUSTER")) == 0);
}

struct RopeKvFuse {   // what rope_kv_kernel needs, for the QKV projection's fused epilogue (skinny_reduce_rope_kv_kernel)
    const float* rope;
    KvLayout kv;
    RowMeta rm;
    int nh, nkv, hd;
};

// y (op)= W x for R >= 2 activation rows, 32 rows per pass: prep (hi/lo bf16 [+ RMSNorm]) -> tcgen05 skinny GEMM -> split
// reduce + epilogue
void skinny_linear(b2l_ctx* c, const uint16_t* W, int N, int K, const float* x, int ldx, const uint16_t* norm_w, uint16_t* xbuf,
                   float* y, int ldy, int mode, int R_all, const TpSend* tps = nullptr, bool x_presplit = false,
                   uint16_t* split_next = nullptr, const uint16_t* next_norm = nullptr, const RopeKvFuse* rope_kv = nullptr) {
    // rope_kv (mode 0, the QKV projection): the K-split reduce also rotates q / k and appends k, v to the paged cache
    // next_norm (mode 1, R_all <= 32, N <= 8192): the residual update is fused with RMSNorm(next_norm) + hi/lo split into split_next
    // x_presplit: the producer of x already wrote the bf16 hi/lo rows into xbuf (R_all <= 32 only)
    // split_next (mode 2 only): write the SwiGLU output as bf16 hi/lo rows for the next projection instead of fp32
    for (int r0 = 0; r0 < R_all; r0 += 32) {
        const int R = std::min(32, R_all - r0);
        const float* xg = x + static_cast<size_t>(r0) * ldx;
        float* yg = y ? y + static_cast<size_t>(r0) * ldy : nullptr;
        const int BT = R <= 8 ? 16 : R <= 16 ? 32 : 64, T = BT / 2;
        if (x_presplit) { /* nothing: xbuf is ready */ }
        else if (norm_w) launch(c, split_bf16_kernel<true>, dim3(R), dim3(256), 0, xg, ldx, norm_w, xbuf, K, T, c->p.rms_norm_eps);
        else launch(c, split_bf16_kernel<false>, dim3(R), dim3(256), 0, xg, ldx, static_cast<const uint16_t*>(nullptr), xbuf, K, T, 0.f);
        const int n_tiles = (N + 127) / 128, n_kblocks = K / kGemmBK;
        // two CTAs per SM are resident: size the K split so the grid is (at most) one full wave
        int ksplit = std::max(1, std::min({2 * c->prop.multiProcessorCount / n_tiles, n_kblocks / 2, 32}));
        while (static_cast<size_t>(ksplit) * T * N > c->sk_partial_floats && ksplit > 1) ksplit--;
        const int per_split = (n_kblocks + ksplit - 1) / ksplit;
        ksplit = (n_kblocks + per_split - 1) / per_split;
        const CUtensorMap mw = make_kmajor_map(W, N, K, 128), mx = make_kmajor_map(xbuf, BT, K, BT);
        const dim3 grid(n_tiles, ksplit);
        if (BT == 16) launch(c, skinny_gemm_kernel<16>, grid, dim3(kSkinnyThreads), skinny_smem(16), mw, mx, c->sk_partial, N, n_kblocks, per_split);
        else if (BT == 32) launch(c, skinny_gemm_kernel<32>, grid, dim3(kSkinnyThreads), skinny_smem(32), mw, mx, c->sk_partial, N, n_kblocks, per_split);
        else launch(c, skinny_gemm_kernel<64>, grid, dim3(kSkinnyThreads), skinny_smem(64), mw, mx, c->sk_partial, N, n_kblocks, per_split);
        const int cols = mode == 2 ? N / 2 : N;
        TpSend send = tps ? *tps : TpSend{};
        for (int p = 0; p < send.tp; p++) send.dst[p] += static_cast<size_t>(r0) * ldy;
        if (mode == 1 && next_norm && split_next) {
            // a cluster of CTAs per row (B2L_NORM_CLUSTER=0: the one-CTA-per-row kernel)
            if (c->norm_cluster)
                launch_cluster(c, skinny_reduce_norm_split_cluster_kernel, dim3(kNormCluster, R), dim3(kNormThreads), kNormCluster,
                               static_cast<const float*>(c->sk_partial), ksplit, T, N, yg, next_norm, split_next, c->p.rms_norm_eps);
            else
                launch(c, skinny_reduce_norm_split_kernel, dim3(R), dim3(1024), 0, static_cast<const float*>(c->sk_partial), ksplit, T, N, yg, next_norm, split_next,
                       c->p.rms_norm_eps);
            continue;
        }
        if (mode == 0 && rope_kv) {
            const RopeKvFuse& f = *rope_kv;
            const RowMeta rm{f.rm.positions + r0, f.rm.slots + r0, f.rm.block_tables, f.rm.max_blocks};
            launch(c, skinny_reduce_rope_kv_kernel, dim3(f.nh + 2 * f.nkv, R), dim3(f.hd / 2), 0, static_cast<const float*>(c->sk_partial), ksplit, T, N, yg, ldy,
                   f.rope, f.kv, rm, f.nh, f.nkv, f.hd);
            continue;
        }
        launch(c, skinny_reduce_kernel, dim3((cols + 255) / 256, R), dim3(256), 0, static_cast<const float*>(c->sk_partial), ksplit, T, N, R, mode, yg, ldy, send,
               split_next, T);
    }
}

// ---- TP over peer memory: the projection kernel's epilogue stores its partial sums into every rank's slab,
// the receiver kernel adds them into the residual stream. No collective call, no separate all-reduce kernel. ----
TpSend tp_send_args(const b2l_ctx* c, int slot) {
    TpSend t{};
    t.tp = c->p.tp_size;
    t.seq = c->tp_seq + slot;
    const size_t slab = static_cast<size_t>(c->max_rows) * c->H;
    for (int p = 0; p < t.tp; p++) t.dst[p] = c->tp_peer[p] + (static_cast<size_t>(slot) * t.tp + c->p.tp_rank) * slab;
    return t;
}
void tp_ll_reduce(b2l_ctx* c, int slot, int R) {
    const int n_pairs = R * c->H / 2;
    const int grid = std::max(1, std::min((n_pairs + 255) / 256, c->prop.multiProcessorCount));
    const uint2* recv = c->tp_ll + static_cast<size_t>(slot) * c->p.tp_size * c->max_rows * c->H;
    launch(c, tp_ll_reduce_kernel, dim3(grid), dim3(256), 0, c->h, recv, c->p.tp_size, R, c->H, c->max_rows, c->tp_seq + slot, c->tp_done + slot);
}

// Map every rank's receive slab into this process (cudaIpc; handles travel through one NCCL all-gather).
// Collective: called by every rank from b2l_create. B2L_TP_TRANSPORT=nccl keeps the NCCL all-reduce instead.
void tp_peer_setup(b2l_ctx* c) {
    const int tp = c->p.tp_size, rank = c->p.tp_rank;
    const char* env = std::getenv("B2L_TP_TRANSPORT");
    const bool want = !(env && std::string(env) == "nccl") && tp <= kTpMaxRanks;
    // [multi-kernel path: 2 slots x tp x max_rows x H] [megakernel: 2 phases x tp x H] [megakernel argmax keys: tp x 2 x SMs]
    // (the two decode paths number their sequences independently, so they must not share words)
    const size_t mk_words = static_cast<size_t>(2) * tp * c->max_rows * c->H;
    c->tp_mega_off = mk_words;
    c->tp_keys_off = mk_words + static_cast<size_t>(2) * tp * c->H;
    const size_t words = c->tp_keys_off + static_cast<size_t>(tp) * 2 * c->prop.multiProcessorCount;
    c->tp_ll = dalloc<uint2>(c, words);
    B2L_CUDA(cudaMemset(c->tp_ll, 0, words * sizeof(uint2)));      // sequence 0 is never sent
    c->tp_seq = dalloc<uint32_t>(c, 2);
    c->tp_done = dalloc<unsigned int>(c, 2);
    const uint32_t ones[2] = {1, 1};
    B2L_CUDA(cudaMemcpy(c->tp_seq, ones, sizeof(ones), cudaMemcpyHostToDevice));
    B2L_CUDA(cudaMemset(c->tp_done, 0, 2 * sizeof(unsigned int)));
    // every rank takes part in the exchange even if it will not use the result, so the collective matches
    constexpr size_t kSlot = 80;   // 64-byte handle + ok flag, padded to a multiple of 16
    static_assert(sizeof(cudaIpcMemHandle_t) <= 64, "ipc handle size");
    std::vector<unsigned char> mine(kSlot, 0), all(kSlot * tp, 0);
    cudaIpcMemHandle_t hd;
    const cudaError_t ge = want ? cudaIpcGetMemHandle(&hd, c->tp_ll) : cudaErrorNotSupported;
    if (ge == cudaSuccess) {
        std::memcpy(mine.data(), &hd, sizeof(hd));
        mine[64] = 1;
    } else {
        cudaGetLastError();
    }
    unsigned char* d_ex = dalloc<unsigned char>(c, kSlot * (tp + 1));
    B2L_CUDA(cudaMemcpy(d_ex, mine.data(), kSlot, cudaMemcpyHostToDevice));
    B2L_NCCL(nccl().AllGather(d_ex, d_ex + kSlot, kSlot / 4, kNcclFloat32, c->nccl_comm, c->stream));
    B2L_CUDA(cudaStreamSynchronize(c->stream));
    B2L_CUDA(cudaMemcpy(all.data(), d_ex + kSlot, kSlot * tp, cudaMemcpyDeviceToHost));
    bool ok = true;
    for (int p = 0; p < tp; p++) ok = ok && all[p * kSlot + 64] == 1;
    int opened = 0;
    if (ok) {
        for (int p = 0; p < tp && ok; p++) {
            if (p == rank) {
                c->tp_peer[p] = c->tp_ll;
                continue;
            }
            cudaIpcMemHandle_t ph;
            std::memcpy(&ph, all.data() + p * kSlot, sizeof(ph));
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, ph, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = false;
                break;
            }
            c->tp_peer[p] = static_cast<uint2*>(ptr);
            opened++;
        }
    }
    // all ranks must agree on the transport: one more tiny exchange of the outcome
    float* d_ok = reinterpret_cast<float*>(d_ex);
    const float mine_ok = ok ? 0.f : 1.f;
    B2L_CUDA(cudaMemcpy(d_ok, &mine_ok, sizeof(float), cudaMemcpyHostToDevice));
    B2L_NCCL(nccl().AllReduce(d_ok, d_ok, 1, kNcclFloat32, kNcclSum, c->nccl_comm, c->stream));
    B2L_CUDA(cudaStreamSynchronize(c->stream));
    float failures = 0.f;
    B2L_CUDA(cudaMemcpy(&failures, d_ok, sizeof(float), cudaMemcpyDeviceToHost));
    c->tp_peer_ok = failures == 0.f;
    if (!c->tp_peer_ok) {
        for (int p = 0; p < tp; p++) {
            if (p != rank && c->tp_peer[p]) cudaIpcCloseMemHandle(c->tp_peer[p]);
            c->tp_peer[p] = nullptr;
        }
    }
    (void)opened;
}

// h += all_reduce_sum(proj) over the TP ranks (fp32; the messages are R x H x 4 bytes: latency-bound)
void tp_allreduce_add(b2l_ctx* c, int R) {
    const size_t n = static_cast<size_t>(R) * c->H;
    B2L_NCCL(nccl().AllReduce(c->proj, c->proj, n, kNcclFloat32, kNcclSum, c->nccl_comm, c->stream));
    launch(c, add_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, c->h, static_cast<const float*>(c->proj), static_cast<int>(n));
}

// greedy argmax over vocab-sharded logits: local (max, global index), all-gather, first-max merge
void tp_argmax(b2l_ctx* c, const float* logits, int R) {
    if (c->p.tp_size == 1) {
        launch_cluster(c, argmax_cluster_kernel, dim3(kArgmaxCluster, R), dim3(1024), kArgmaxCluster, logits, c->V_l, c->V_l, 0, c->d_next_ids,
                       static_cast<float*>(nullptr));
        return;
    }
    launch_cluster(c, argmax_cluster_kernel, dim3(kArgmaxCluster, R), dim3(1024), kArgmaxCluster, logits, c->V_l, c->V_l, c->p.tp_rank * c->V_l,
                   c->d_next_ids, c->tp_vals);
    launch(c, tp_pack_kernel, dim3(1), dim3(64), 0, static_cast<const float*>(c->tp_vals), static_cast<const int32_t*>(c->d_next_ids), c->tp_pack, R);
    B2L_NCCL(nccl().AllGather(c->tp_pack, c->tp_gather, static_cast<size_t>(c->max_rows) * 2, kNcclFloat32, c->nccl_comm, c->stream));
    launch(c, tp_merge_kernel, dim3(1), dim3(64), 0, static_cast<const float*>(c->tp_gather), c->d_next_ids, R, c->max_rows, c->p.tp_size);
}

// The forward pass for R rows whose (token, position, slot) are already in device buffers.
//   want_logits: run final norm + lm_head + argmax;  tap_row0 >= 0: record the residual stream.
void enqueue_forward(b2l_ctx* c, int R, bool want_logits, int tap_row0) {
    const RowMeta rm{c->d_positions, c->d_slots, c->d_block_tables, c->max_blocks_cap};
    launch(c, embed_kernel, dim3(R), dim3(256), 0, c->embed, c->d_tokens, c->h, c->H, c->V);
    if (tap_row0 >= 0) tap_copy(c, 0, tap_row0, c->h, R);
    const float scale = 1.0f / sqrtf(static_cast<float>(c->hd));
    for (int l = 0; l < c->L; l++) {
        const LayerWeights& w = c->layers[l];
        const KvLayout kv{w.kv_pool, c->p.page_size, c->kvd_l};
        const bool sk = c->skinny_ok && R >= 2;   // batched decode: projections on the tensor cores
        // single GPU, one row group: the residual epilogues of O and down also produce the next projection's normalised operand
        const bool fuse_norm = sk && R <= 32 && c->p.tp_size == 1 && c->H <= 8192;
        // batched decode: the QKV projection's split-K reduce rotates q / k and appends k, v itself (B2L_ROPE_FUSE=0: separate kernels)
        const bool rope_fused = sk && c->rope_fuse;
        const RopeKvFuse rkf{c->rope, kv, rm, c->nh_l, c->nkv_l, c->hd};
        if (sk) skinny_linear(c, w.w_qkv, c->qkv_l, c->H, c->h, c->H, w.in_norm, c->sk_xh, c->qkv, c->qkv_l, 0, R, nullptr, fuse_norm && l > 0, nullptr, nullptr,
                              rope_fused ? &rkf : nullptr);
        else gemv(c, w.w_qkv, c->h, c->H, c->qkv, c->qkv_l, w.in_norm, c->qkv_l, c->H, 0, R);
        if (!rope_fused) launch(c, rope_kv_kernel, dim3(R), dim3(256), 0, c->qkv, c->qkv_l, c->rope, kv, rm, c->nh_l, c->nkv_l, c->hd, 0);
        const bool fuse = sk && R <= 32;   // one row group: producers write the next projection's bf16 hi/lo operand directly
        const int skT = R <= 8 ? 8 : R <= 16 ? 16 : 32;
        AttnArgs aa{c->qkv, c->qkv_l, kv, rm, c->part_acc, c->part_ml, c->attn_counters, c->attn, c->qd_l, scale};
        if (fuse) {
            aa.split_out = c->sk_xq;
            aa.split_T = skT;
        }
        attn_launch(c, aa, R);
        const bool tp = c->p.tp_size > 1;
        // O projection (+ residual); under TP every rank holds a K slice and the partial products are summed over NVLink
        const bool peer = tp && c->tp_peer_ok;
        const TpSend send0 = peer ? tp_send_args(c, 0) : TpSend{}, send1 = peer ? tp_send_args(c, 1) : TpSend{};
        const int omode = peer ? 3 : tp ? 0 : 1;
        if (sk) skinny_linear(c, w.w_o, c->H, c->qd_l, c->attn, c->qd_l, nullptr, c->sk_xq, tp ? c->proj : c->h, c->H, omode, R, &send0, fuse,
                              fuse_norm ? c->sk_xh : nullptr, fuse_norm ? w.post_norm : nullptr);
        else gemv(c, w.w_o, c->attn, c->qd_l, tp ? c->proj : c->h, c->H, nullptr, c->H, c->qd_l, omode, R, &send0);
        if (peer) tp_ll_reduce(c, 0, R);
        else if (tp) tp_allreduce_add(c, R);
        if (sk) skinny_linear(c, w.w_gu, 2 * c->I_l, c->H, c->h, c->H, w.post_norm, c->sk_xh, c->act, c->I_l, 2, R, nullptr, fuse_norm, fuse ? c->sk_xi : nullptr);
        else gemv(c, w.w_gu, c->h, c->H, c->act, c->I_l, w.post_norm, 2 * c->I_l, c->H, 2, R);
        if (sk) skinny_linear(c, w.w_down, c->H, c->I_l, c->act, c->I_l, nullptr, c->sk_xi, tp ? c->proj : c->h, c->H, omode, R, &send1, fuse,
                              fuse_norm ? c->sk_xh : nullptr, fuse_norm ? (l + 1 < c->L ? c->layers[l + 1].in_norm : c->final_norm) : nullptr);
        else gemv(c, w.w_down, c->act, c->I_l, tp ? c->proj : c->h, c->H, nullptr, c->H, c->I_l, omode, R, &send1);
        if (peer) tp_ll_reduce(c, 1, R);
        else if (tp) tp_allreduce_add(c, R);
        if (tap_row0 >= 0) tap_copy(c, l + 1, tap_row0, c->h, R);
    }
    if (tap_row0 >= 0) {
        float* dst = c->tap + (static_cast<size_t>(c->L + 1) * c->tap_rows_cap + tap_row0) * c->H;
        rmsnorm_kernel<<<R, 256, 0, c->stream>>>(c->h, c->final_norm, dst, c->H, c->p.rms_norm_eps);
        c->launched++;
    }
    if (want_logits) {
        if (c->skinny_ok && R >= 2)
            skinny_linear(c, c->lm_head, c->V_l, c->H, c->h, c->H, c->final_norm, c->sk_xh, c->logits, c->V_l, 0, R, nullptr,
                          /*x_presplit: the last down projection already normalised and split h*/ R <= 32 && c->p.tp_size == 1 && c->H <= 8192);
        else gemv(c, c->lm_head, c->h, c->H, c->logits, c->V_l, c->final_norm, c->V_l, c->H, 0, R);
        tp_argmax(c, c->logits, R);
    }
}

void gemm_bf16(b2l_ctx* c, const uint16_t* A, const uint16_t* W, GemmArgs g);  // defined with the tensor-map helpers below
CUtensorMap make_kmajor_map(const uint16_t* ptr, int64_t rows, int64_t cols, int box_rows);

constexpr int kPfAttnChunk = 128;  // rows per split-K attention launch in the GEMM prefill path

void prefill_alloc(b2l_ctx* c) {
    if (c->pf_h) return;
    const size_t T = static_cast<size_t>(c->p.max_prefill_tokens), H = c->H;
    c->pf_rows = static_cast<int>(T);
    c->pf_tokens = dalloc<int32_t>(c, T); c->pf_positions = dalloc<int32_t>(c, T); c->pf_slots = dalloc<int32_t>(c, T);
    c->pf_last = dalloc<int32_t>(c, c->p.max_batch);
    c->pf_h = dalloc<float>(c, T * H);
    c->pf_qkv = dalloc<float>(c, T * c->qkv_l);
    c->pf_attn = dalloc<float>(c, T * c->qd_l);
    c->pf_xn = dalloc<uint16_t>(c, T * H);
    c->pf_attn16 = dalloc<uint16_t>(c, T * c->qd_l);
    c->pf_act16 = dalloc<uint16_t>(c, T * c->I_l);
    if (c->p.tp_size > 1) c->pf_proj = dalloc<float>(c, T * H);
    const size_t rows = kPfAttnChunk;
    // rows x splits of one launch never exceeds rows + 8 SMs / kv heads (attn_launch_hd sizes the split count by the rows)
    const size_t row_splits = rows + static_cast<size_t>(8 * c->prop.multiProcessorCount) / c->nkv_l + 1;
    c->pf_part_acc = dalloc<float>(c, row_splits * c->nkv_l * c->group * c->hd);
    c->pf_part_ml = dalloc<float>(c, row_splits * c->nkv_l * c->group * 2);
    c->pf_counters = dalloc<int>(c, rows * c->nkv_l);
    c->pf_tiles = dalloc<PrefillTile>(c, T / 1 + static_cast<size_t>(c->p.max_batch));  // <= one tile per row in the worst case
    c->pf_tiles_tc = dalloc<PrefillTile>(c, T / 1 + static_cast<size_t>(c->p.max_batch));
    {
        // tcgen05 attention: one tensor map over every layer's pool, rows = [layer][page][K|V][slot] of kvd elements
        const int ps = c->p.page_size;
        const int64_t rows = static_cast<int64_t>(c->L) * c->p.num_pages * 2 * ps;
        const char* tc_env = std::getenv("B2L_FLASH_TC");   // development / test switch, read per context: 0 = the mma.sync kernel
        const bool tc_on = !tc_env || std::atoi(tc_env) != 0;
        c->flash_tc_ok = tc_on && (c->hd == 64 || c->hd == 128) && ps >= 8 && ps <= 128 && 128 % ps == 0 && ps % 8 == 0 && rows < (1ll << 31);
        if (c->flash_tc_ok) {
            const CUtensorMap m = make_kmajor_map(c->kv_base, rows, c->kvd_l, ps);
            static_assert(sizeof(CUtensorMap) <= sizeof(c->kv_map), "kv_map storage");
            std::memcpy(c->kv_map, &m, sizeof(m));
        }
    }
    B2L_CUDA(cudaMemset(c->pf_counters, 0, sizeof(int) * rows * c->nkv_l));
}

// Prefill on the tensor cores: every new token of every sequence is one GEMM row (bf16 activations, fp32
// accumulate and residual stream). RoPE + K/V append run for all rows first, so the split-K attention of a
// row sees every earlier position of its sequence, including the ones appended in this same call.
void prefill_gemm(b2l_ctx* c, int T, int n_seq, int tap_row0) {
    prefill_alloc(c);
    const RowMeta rm{c->pf_positions, c->pf_slots, c->d_block_tables, c->max_blocks_cap};
    const float scale = 1.0f / sqrtf(static_cast<float>(c->hd));
    const float eps = c->p.rms_norm_eps;
    auto tap = [&](int slab) {
        if (tap_row0 < 0) return;
        float* dst = c->tap + (static_cast<size_t>(slab) * c->tap_rows_cap + tap_row0) * c->H;
        B2L_CUDA(cudaMemcpyAsync(dst, c->pf_h, sizeof(float) * T * c->H, cudaMemcpyDeviceToDevice, c->stream));
    };
    launch(c, embed_kernel, dim3(T), dim3(256), 0, c->embed, c->pf_tokens, c->pf_h, c->H, c->V);
    tap(0);
    for (int l = 0; l < c->L; l++) {
        const LayerWeights& w = c->layers[l];
        const KvLayout kv{w.kv_pool, c->p.page_size, c->kvd_l};
        rmsnorm_bf16_kernel<<<T, 256, 0, c->stream>>>(c->pf_h, w.in_norm, c->pf_xn, c->H, eps);
        c->launched++;
        gemm_bf16(c, c->pf_xn, w.w_qkv, GemmArgs{c->pf_qkv, nullptr, T, c->qkv_l, c->H, c->qkv_l, GEMM_STORE_F32});
        launch(c, rope_kv_kernel, dim3(T), dim3(256), 0, c->pf_qkv, c->qkv_l, c->rope, kv, rm, c->nh_l, c->nkv_l, c->hd, c->flash_tc_ok ? c->nh_l : 0);
        if (c->flash_tc_ok) {
            // flash-style causal attention on tcgen05 / TMEM (flash_prefill_tc.cuh)
            CUtensorMap mkv;
            std::memcpy(&mkv, c->kv_map, sizeof(mkv));
            const FlashTcArgs fa{c->pf_qkv, c->qkv_l, c->rope, c->d_block_tables, c->max_blocks_cap, static_cast<const PrefillTile*>(c->pf_tiles_tc),
                                 c->pf_attn16, c->qd_l, c->group, scale * 1.4426950408889634f, c->p.page_size,
                                 static_cast<long long>(l) * c->p.num_pages * 2 * c->p.page_size};
            const dim3 grid(c->pf_n_tiles_tc, c->nh_l);
            auto go = [&](auto kern, size_t smem) {
                ensure_smem_optin(reinterpret_cast<const void*>(kern), c->p.device, static_cast<int>(smem));
                kern<<<grid, kFtcThreads, smem, c->stream>>>(mkv, fa);
            };
            if (c->hd == 64) go(flash_prefill_tc_kernel<64>, FtcSmem<64>::total);
            else go(flash_prefill_tc_kernel<128>, FtcSmem<128>::total);
            B2L_CUDA(cudaGetLastError());
            c->launched++;
        } else if (c->hd == 64 || c->hd == 128) {
            // flash-style causal attention over the paged cache, tensor-core S and PV, bf16 output
            FlashArgs fa{c->pf_qkv, c->qkv_l, kv, c->d_block_tables, c->max_blocks_cap, static_cast<const PrefillTile*>(c->pf_tiles),
                         c->pf_attn16, c->qd_l, c->group, scale * 1.4426950408889634f};
            // GH query heads of one kv head per CTA share the K/V tiles; shared memory: two {K, V} buffers of 64 padded rows each
            const int gh = flash_heads_per_cta(c->hd, c->group);
            const dim3 grid(c->pf_n_tiles, c->nh_l / gh);
            const size_t fsmem = static_cast<size_t>(4) * 64 * (c->hd + 8) * 2;
            auto go = [&](auto kern) {
                ensure_smem_optin(reinterpret_cast<const void*>(kern), c->p.device, static_cast<int>(fsmem));
                kern<<<grid, kFlashThreads * gh, fsmem, c->stream>>>(fa);
            };
            if (c->hd == 64) {
                if (gh == 4) go(flash_prefill_kernel<64, 4>);
                else if (gh == 3) go(flash_prefill_kernel<64, 3>);
                else if (gh == 2) go(flash_prefill_kernel<64, 2>);
                else go(flash_prefill_kernel<64, 1>);
            } else {
                if (gh == 3) go(flash_prefill_kernel<128, 3>);
                else if (gh == 2) go(flash_prefill_kernel<128, 2>);
                else go(flash_prefill_kernel<128, 1>);
            }
            B2L_CUDA(cudaGetLastError());
            c->launched++;
        } else {
            for (int r0 = 0; r0 < T; r0 += kPfAttnChunk) {
                const int R = std::min(kPfAttnChunk, T - r0);
                const RowMeta rmc{c->pf_positions + r0, c->pf_slots + r0, c->d_block_tables, c->max_blocks_cap};
                const AttnArgs aa{c->pf_qkv + static_cast<size_t>(r0) * c->qkv_l, c->qkv_l, kv, rmc, c->pf_part_acc, c->pf_part_ml, c->pf_counters,
                                  c->pf_attn + static_cast<size_t>(r0) * c->qd_l, c->qd_l, scale};
                attn_launch(c, aa, R);
            }
            const size_t n = static_cast<size_t>(T) * c->qd_l;
            cast_bf16_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, c->stream>>>(c->pf_attn, c->pf_attn16, n);
            c->launched++;
        }
        if (c->p.tp_size == 1) {
            gemm_bf16(c, c->pf_attn16, w.w_o, GemmArgs{c->pf_h, nullptr, T, c->H, c->qd_l, c->H, GEMM_ADD_F32});
        } else {
            gemm_bf16(c, c->pf_attn16, w.w_o, GemmArgs{c->pf_proj, nullptr, T, c->H, c->qd_l, c->H, GEMM_STORE_F32});
            const size_t n = static_cast<size_t>(T) * c->H;
            B2L_NCCL(nccl().AllReduce(c->pf_proj, c->pf_proj, n, kNcclFloat32, kNcclSum, c->nccl_comm, c->stream));
            launch(c, add_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, c->pf_h, static_cast<const float*>(c->pf_proj), static_cast<int>(n));
        }
        rmsnorm_bf16_kernel<<<T, 256, 0, c->stream>>>(c->pf_h, w.post_norm, c->pf_xn, c->H, eps);
        c->launched++;
        gemm_bf16(c, c->pf_xn, w.w_gu, GemmArgs{nullptr, c->pf_act16, T, 2 * c->I_l, c->H, c->I_l, GEMM_SWIGLU_BF16});
        if (c->p.tp_size == 1) {
            gemm_bf16(c, c->pf_act16, w.w_down, GemmArgs{c->pf_h, nullptr, T, c->H, c->I_l, c->H, GEMM_ADD_F32});
        } else {
            gemm_bf16(c, c->pf_act16, w.w_down, GemmArgs{c->pf_proj, nullptr, T, c->H, c->I_l, c->H, GEMM_STORE_F32});
            const size_t n = static_cast<size_t>(T) * c->H;
            B2L_NCCL(nccl().AllReduce(c->pf_proj, c->pf_proj, n, kNcclFloat32, kNcclSum, c->nccl_comm, c->stream));
            launch(c, add_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, c->pf_h, static_cast<const float*>(c->pf_proj), static_cast<int>(n));
        }
        tap(l + 1);
    }
    if (tap_row0 >= 0) {
        float* dst = c->tap + (static_cast<size_t>(c->L + 1) * c->tap_rows_cap + tap_row0) * c->H;
        rmsnorm_kernel<<<T, 256, 0, c->stream>>>(c->pf_h, c->final_norm, dst, c->H, eps);
        c->launched++;
    }
    // lm_head only for the last token of each sequence: gather those rows, then the (fused final norm) GEMV
    gather_rows_kernel<<<n_seq, 256, 0, c->stream>>>(c->pf_h, c->pf_last, c->h, c->H);
    c->launched++;
    gemv(c, c->lm_head, c->h, c->H, c->seq_logits, c->V_l, c->final_norm, c->V_l, c->H, 0, n_seq);
}

Graph& decode_graph(b2l_ctx* c, int R, bool with_advance) {
    const int key = R + (with_advance ? 1000 : 0);
    auto it = c->decode_graphs.find(key);
    if (it != c->decode_graphs.end()) return it->second;
    Graph g;
    const int64_t before = c->launched;
    B2L_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    try {
        enqueue_forward(c, R, true, -1);
        if (with_advance)
            launch(c, advance_kernel, dim3(1), dim3(64), 0, static_cast<const int32_t*>(c->d_next_ids), c->d_tokens,
                   c->d_positions, c->d_out_ids, c->d_step, R);
    } catch (...) {
        cudaGraph_t junk = nullptr;
        cudaStreamEndCapture(c->stream, &junk);
        if (junk) cudaGraphDestroy(junk);
        throw;
    }
    B2L_CUDA(cudaStreamEndCapture(c->stream, &g.graph));
    g.nodes = static_cast<int>(c->launched - before);
    c->launched = before;  // capture is not execution
    B2L_CUDA(cudaGraphInstantiate(&g.exec, g.graph, 0));
    return c->decode_graphs.emplace(key, g).first->second;
}

void upload_rows_meta(b2l_ctx* c, int n_tok, const int32_t* tokens, const int32_t* positions, const int32_t* slots) {
    // pinned staging -> async copies on the engine stream
    int32_t* s = c->h_stage;
    std::memcpy(s, tokens, sizeof(int32_t) * n_tok);
    std::memcpy(s + c->max_rows, positions, sizeof(int32_t) * n_tok);
    std::memcpy(s + 2 * c->max_rows, slots, sizeof(int32_t) * n_tok);
    B2L_CUDA(cudaMemcpyAsync(c->d_tokens, s, sizeof(int32_t) * n_tok, cudaMemcpyHostToDevice, c->stream));
    B2L_CUDA(cudaMemcpyAsync(c->d_positions, s + c->max_rows, sizeof(int32_t) * n_tok, cudaMemcpyHostToDevice, c->stream));
    B2L_CUDA(cudaMemcpyAsync(c->d_slots, s + 2 * c->max_rows, sizeof(int32_t) * n_tok, cudaMemcpyHostToDevice, c->stream));
}

void upload_block_tables(b2l_ctx* c, int n_seq, const int32_t* bt, int max_blocks, const int32_t* need_tokens) {
    B2L_CHECK(max_blocks >= 1, "max_blocks must be >= 1");
    int32_t* s = c->h_stage + 3 * c->max_rows;
    const int cap = c->max_blocks_cap;
    for (int i = 0; i < n_seq; i++) {
        const int need = (need_tokens[i] + c->p.page_size - 1) / c->p.page_size;
        B2L_CHECK(need <= max_blocks && need <= cap, "block table too short for the sequence length");
        for (int j = 0; j < cap; j++) {
            const int32_t page = j < max_blocks ? bt[static_cast<size_t>(i) * max_blocks + j] : 0;
            if (j < need) B2L_CHECK(page >= 0 && page < c->p.num_pages, "block table holds a page id outside the pool");
            s[static_cast<size_t>(i) * cap + j] = j < need ? page : 0;
        }
    }
    // a decode step inside a page leaves the tables unchanged: do not pay a copy-engine round trip for it
    const size_t n = static_cast<size_t>(n_seq) * cap;
    if (c->bt_uploaded_rows == n_seq && c->bt_uploaded.size() >= n && std::memcmp(c->bt_uploaded.data(), s, n * sizeof(int32_t)) == 0) return;
    B2L_CUDA(cudaMemcpyAsync(c->d_block_tables, s, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    c->bt_uploaded.assign(s, s + n);
    c->bt_uploaded_rows = n_seq;
}

// ---- persistent megakernel (mega_decode.cuh) ------------------------------------------------
// Tiled weight image of one matrix for the megakernel: a permutation of the row-major [N][K] matrix (same size, no padding).
// CTA c owns rows [r0, r1) (mega_row_range); they are taken in groups of <= 16 rows; a group of r rows occupies the same
// r*K elements as in the row-major matrix, but ordered [K window of KS elements][16-byte piece of 8 k][row][8 elements]:
// one window of a group is one contiguous bulk copy (r * KS * 2 bytes), and the 8 rows of an 8x8 `ldmatrix` tile are
// 128 contiguous bytes (no bank conflicts for any r). One thread moves one 16-byte piece.
__global__ void mega_tile_kernel(const uint4* __restrict__ W, uint4* __restrict__ Wt, int N, int K, int unit, int G) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int kc_per_row = K / 8;
    if (idx >= static_cast<long long>(N) * kc_per_row) return;
    const int row = static_cast<int>(idx / kc_per_row), kc = static_cast<int>(idx % kc_per_row);
    const long long units = N / unit;
    int c = static_cast<int>(static_cast<long long>(row / unit) * G / units);
    int r0, r1;
    mega_row_range(N, unit, c, G, r0, r1);
    while (row >= r1) { c++; mega_row_range(N, unit, c, G, r0, r1); }
    while (row < r0) { c--; mega_row_range(N, unit, c, G, r0, r1); }
    const int g = (row - r0) / kMegaGroupRows, i = (row - r0) % kMegaGroupRows;
    const int gr0 = r0 + g * kMegaGroupRows, r = min(kMegaGroupRows, r1 - gr0);
    const int pieces = 1 << (mega_ks_shift(K) - 3);   // 16-byte pieces per window per row
    const int p = kc / pieces, cc = kc % pieces;
    Wt[static_cast<long long>(gr0) * kc_per_row + static_cast<long long>(p) * r * pieces + static_cast<long long>(cc) * r + i] = W[idx];
}

// the megakernel is instantiated for the (head_dim, GQA group) pairs of the supported models: Llama-3.2-1B (64, 4), 3B (128, 3),
// Llama-3.1-8B (128, 4), 70B (128, 8), and the two test presets (32, 4) / (128, 2); a tensor-parallel rank keeps the group
typedef void (*MegaKernel)();
MegaKernel mega_kernel_for(int hd, int group, bool tp) {
#define B2L_MEGA_CASE(H, G) if (hd == H && group == G) return tp ? static_cast<MegaKernel>(megatp::mega_decode_kernel<H, G>) : static_cast<MegaKernel>(mega1::mega_decode_kernel<H, G>);
    B2L_MEGA_CASE(64, 4)
    B2L_MEGA_CASE(128, 3)
    B2L_MEGA_CASE(128, 4)
    B2L_MEGA_CASE(128, 8)
    B2L_MEGA_CASE(32, 4)
    B2L_MEGA_CASE(128, 2)
#undef B2L_MEGA_CASE
    return nullptr;
}

void mega_setup(b2l_ctx* c) {
    c->mega_ok = false;
    auto no = [&](const std::string& why) { c->mega_why = why; };
    if (c->p.tp_size != 1 && !c->tp_peer_ok) return no("tensor-parallel megakernel needs the NVLink peer-memory transport");
    if (!mega_kernel_for(c->hd, c->group, c->p.tp_size > 1)) return no("no megakernel instance for this (head_dim, GQA group)");
    const int G = c->prop.multiProcessorCount;
    std::vector<MegaPhase> ph;
    int k_max = 0;
    auto add = [&](int type, int layer, const uint16_t* W, const uint16_t* norm, uint16_t* kv, int N, int K) -> bool {
        MegaPhase p{};
        p.type = type; p.layer = layer; p.W = W; p.norm_w = norm; p.kv_pool = kv; p.N = N; p.K = K; p.inv_k = K > 0 ? 1.0f / static_cast<float>(K) : 0.f; p.reserved = 0;
        if (type != PH_ATTN && (K % 128 != 0 || N % 2 != 0)) return false;
        k_max = std::max(k_max, K);
        ph.push_back(p);
        return true;
    };
    bool ok = true;
    for (int l = 0; l < c->L && ok; l++) {
        const LayerWeights& w = c->layers[l];
        ok = ok && add(PH_QKV, l, w.w_qkv, w.in_norm, nullptr, c->qkv_l, c->H);
        ok = ok && add(PH_ATTN, l, nullptr, nullptr, w.kv_pool, 0, 0);
        ok = ok && add(PH_OPROJ, l, w.w_o, nullptr, nullptr, c->H, c->qd_l);
        ok = ok && add(PH_GATEUP, l, w.w_gu, w.post_norm, nullptr, 2 * c->I_l, c->H);
        ok = ok && add(PH_DOWN, l, w.w_down, nullptr, nullptr, c->H, c->I_l);
    }
    ok = ok && add(PH_LMHEAD, c->L, c->lm_head, c->final_norm, nullptr, c->V_l, c->H);
    if (!ok) return no("a weight matrix has K that is not a multiple of 128 (or an odd row count)");
    if (c->nkv_l > G) return no("more kv heads than SMs");
    if ((c->p.page_size & (c->p.page_size - 1)) != 0) return no("KV page size is not a power of two");
    c->mega_nsplit = std::max(1, std::min(c->nsplit, G / c->nkv_l));
    if (const char* e = std::getenv("B2L_MEGA_NSPLIT")) c->mega_nsplit = std::max(1, std::min(c->mega_nsplit, std::atoi(e)));   // tuning knob
    // the input vector as bf16 hi/mid/lo B fragments: 96 bytes per 16 elements; the attention scratch aliases that area
    const size_t attn_scratch = static_cast<size_t>(kMegaConsumerWarps) * c->group * (c->hd + 2) * sizeof(float);
    const size_t xfrag = std::max(static_cast<size_t>(k_max) * 6, attn_scratch);
    const size_t fixed = 8 * kMegaMaxStages * 2 + 16 + 64 + 64 + 64 + 4 * 2 * kMegaBatchGroups * kMegaConsumerWarps * 16 + (48 + 8) * ph.size() + 16 + xfrag + 256;
    int max_smem = 0;
    B2L_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->p.device));
    if (static_cast<size_t>(max_smem) < fixed + 6 * static_cast<size_t>(kMegaStageBytes)) return no("not enough shared memory for the input fragments (K too large) and the weight ring");
    const int stages = std::min<int>(kMegaMaxStages, static_cast<int>((static_cast<size_t>(max_smem) - fixed) / kMegaStageBytes));
    c->mega_stages = stages;
    c->mega_smem = static_cast<size_t>(stages) * kMegaStageBytes + fixed;
    {
        MegaKernel kern = mega_kernel_for(c->hd, c->group, c->p.tp_size > 1);   // tensor-parallel rank / single GPU
        B2L_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(c->mega_smem)));
        // the kernel calls non-inlined device functions: make sure the per-thread stack covers its frames
        cudaFuncAttributes fa{};
        B2L_CUDA(cudaFuncGetAttributes(&fa, kern));
        size_t cur = 0;
        B2L_CUDA(cudaDeviceGetLimit(&cur, cudaLimitStackSize));
        const size_t want = static_cast<size_t>(fa.localSizeBytes) + 2048;
        if (cur < want) B2L_CUDA(cudaDeviceSetLimit(cudaLimitStackSize, want));
        int per_sm = 0;
        B2L_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kMegaThreads, c->mega_smem));
        if (per_sm < 1) return no("megakernel does not fit on an SM");
    }
    {   // dataflow buffers; sequence number 0 (the memset) is never produced
        const size_t n_part = static_cast<size_t>(c->nkv_l) * c->mega_nsplit * c->group;
        auto zalloc = [&](size_t words) {
            unsigned long long* p = dalloc<unsigned long long>(c, words);
            B2L_CUDA(cudaMemset(p, 0, words * sizeof(unsigned long long)));
            return p;
        };
        c->mega_ll_h = zalloc(c->H);
        c->mega_ll_qkv = zalloc(c->qkv_l);
        c->mega_ll_act = zalloc(c->I_l);
        c->mega_ll_pacc = zalloc(n_part * c->hd);
        c->mega_ll_pml = zalloc(n_part * 2);
        c->mega_ll_keys = zalloc(static_cast<size_t>(2) * G);
        c->mega_seq = 0;
    }
    // the tiled weight images (a second copy of every matrix; the row-major one serves prefill and the multi-kernel path)
    for (MegaPhase& p : ph) {
        if (p.type == PH_ATTN) continue;
        const size_t n = static_cast<size_t>(p.N) * p.K;
        uint16_t* wt = dalloc<uint16_t>(c, n);
        const long long pieces = static_cast<long long>(n / 8);
        mega_tile_kernel<<<static_cast<unsigned>((pieces + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const uint4*>(p.W), reinterpret_cast<uint4*>(wt),
                                                                                           p.N, p.K, p.type == PH_GATEUP ? 2 : 1, G);
        B2L_CUDA(cudaGetLastError());
        p.W = wt;
    }
    B2L_CUDA(cudaStreamSynchronize(c->stream));
    MegaPhase* d = dalloc<MegaPhase>(c, ph.size());
    B2L_CUDA(cudaMemcpy(d, ph.data(), sizeof(MegaPhase) * ph.size(), cudaMemcpyHostToDevice));
    c->mega_phases = d;
    c->mega_n_phases = static_cast<int>(ph.size());
    B2L_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&c->mega_abort), sizeof(int) * 1024, cudaHostAllocMapped));
    std::memset(c->mega_abort, 0, sizeof(int) * 1024);
    {   // L2 persistence for the KV cache (B2L_MEGA_L2PERSIST=0 disables)
        const char* e = std::getenv("B2L_MEGA_L2PERSIST");
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c->p.device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c->p.device);
        size_t want = std::min<size_t>(static_cast<size_t>(c->kv_bytes), static_cast<size_t>(std::min(max_persist, max_window)));
        if (e && std::atoi(e) == 0) want = 0;
        if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) c->mega_l2_persist_bytes = want;
        else cudaGetLastError();
    }
    if (const char* e = std::getenv("B2L_MEGA_ATTN_TPS")) c->mega_attn_tps = std::max(16, std::atoi(e));
    if (const char* e = std::getenv("B2L_MEGA_STAGES")) c->mega_stages = std::max(2, std::min(c->mega_stages, std::atoi(e)));
    c->mega_ok = true;
}

// n_steps greedy tokens in one cooperative launch; token/position/block table already on the device
// host_io (single-step API path): token / position travel in the launch arguments, the result comes back through
// mapped pinned memory -- no copy-engine operations at all around the launch
constexpr int kMegaIoWord = 600;   // mega_abort[600..602] = token, position (host side only), result (written by the kernel)
void mega_enqueue(b2l_ctx* c, int n_steps, bool host_io = false) {
    B2L_CHECK(c->mega_ok, "megakernel unavailable: " + c->mega_why);
    MegaArgs a{};
    a.phases = static_cast<const MegaPhase*>(c->mega_phases);
    a.n_phases = c->mega_n_phases;
    a.n_stages = c->mega_stages;
    a.embed = c->embed; a.rope = c->rope;
    a.H = c->H; a.V = c->V_l; a.nh = c->nh_l; a.nkv = c->nkv_l; a.hd = c->hd; a.I = c->I_l;
    a.eps = c->p.rms_norm_eps; a.attn_scale = 1.0f / sqrtf(static_cast<float>(c->hd));
    a.logits = c->logits;
    a.block_table = c->d_block_tables; a.page_size = c->p.page_size; a.kvd = c->kvd_l;
    a.page_shift = 0;
    while ((1 << a.page_shift) < a.page_size) a.page_shift++;
    a.nsplit_max = c->mega_nsplit;
    a.token = c->d_tokens; a.position = c->d_positions; a.out_ids = c->d_out_ids; a.n_steps = n_steps;
    a.ll_h = c->mega_ll_h; a.ll_qkv = c->mega_ll_qkv; a.ll_act = c->mega_ll_act; a.ll_pacc = c->mega_ll_pacc;
    a.ll_pml = c->mega_ll_pml; a.ll_keys = c->mega_ll_keys;
    a.seq_base = c->mega_seq;
    a.tp = c->p.tp_size; a.tp_rank = c->p.tp_rank; a.vocab_base = c->p.tp_rank * c->V_l; a.Hpad = c->H;
    if (a.tp > 1) {
        for (int p = 0; p < a.tp; p++) {
            a.tp_slab[p] = reinterpret_cast<unsigned long long*>(c->tp_peer[p] + c->tp_mega_off);
            a.tp_keys[p] = reinterpret_cast<unsigned long long*>(c->tp_peer[p] + c->tp_keys_off);
        }
        a.ll_keys = a.tp_keys[a.tp_rank];   // this rank's array: every rank's CTAs store their keys here
    } else {
        a.tp_keys[0] = c->mega_ll_keys;
    }
    a.poll_sleep_ns = std::getenv("B2L_MEGA_POLL_NS") ? std::atoi(std::getenv("B2L_MEGA_POLL_NS")) : 0;
    c->mega_seq += static_cast<uint32_t>(n_steps) * static_cast<uint32_t>(c->mega_n_phases);
    int* dev_abort = nullptr;
    B2L_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&dev_abort), c->mega_abort, 0));
    a.abort_flag = dev_abort;
    if (host_io) {
        // token and position ride in the launch arguments (constant memory: many SMs reading one mapped host word over
        // PCIe serialise, measured +0.25 ms); only the result is a (posted) store into mapped host memory
        a.arg_io = 1;
        a.token0 = c->mega_abort[kMegaIoWord];
        a.pos0 = c->mega_abort[kMegaIoWord + 1];
        a.out_ids = dev_abort + kMegaIoWord + 2;
    }
    a.prof = c->mega_prof;
    a.debug_progress = std::getenv("B2L_MEGA_DEBUG") ? 1 : 0;
    a.debug_nostream = std::getenv("B2L_MEGA_NOSTREAM") ? 1 : 0;
    a.attn_tps = c->mega_attn_tps;
    a.attn_max_splits = std::getenv("B2L_MEGA_MAXSPLIT") ? std::max(1, std::atoi(std::getenv("B2L_MEGA_MAXSPLIT"))) : 16;
    a.attn_qhead_tokens = std::getenv("B2L_MEGA_QHEAD_TOK") ? std::max(32, std::atoi(std::getenv("B2L_MEGA_QHEAD_TOK"))) : 512;   // measured on the 8B TP=8 rank shapes at context 4096: 16 / 512 -> 0.99 ms, 8 / 256 -> 1.80 ms
    a.l2_ahead = std::getenv("B2L_MEGA_L2AHEAD") ? std::atoi(std::getenv("B2L_MEGA_L2AHEAD")) : 6;   // measured (profiles/r02_megakernel_knob_sweeps.txt): 6 chunks (14 MB chip-wide) best, 8 +1.2 %, 12 +4 %, 32 +9 % (L2 thrash)
    a.producer_sleep_ns = std::getenv("B2L_MEGA_PSLEEP") ? std::atoi(std::getenv("B2L_MEGA_PSLEEP")) : 100;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(c->prop.multiProcessorCount);
    cfg.blockDim = dim3(kMegaThreads);
    cfg.dynamicSmemBytes = c->mega_smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (c->mega_l2_persist_bytes > 0) {
        // keep the KV cache resident in the 126 MB L2 while 2.5 GB of weights stream past it every token
        attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[1].val.accessPolicyWindow.base_ptr = c->kv_base;
        attr[1].val.accessPolicyWindow.num_bytes = c->mega_l2_persist_bytes;
        attr[1].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.numAttrs = 2;
    }
    if (a.tp > 1) {
        B2L_CUDA(cudaMemcpyToSymbolAsync(megatp::c_mega, &a, sizeof(MegaArgs), 0, cudaMemcpyHostToDevice, c->stream));
        B2L_CUDA(cudaLaunchKernelEx(&cfg, mega_kernel_for(c->hd, c->group, true)));
    } else {
        B2L_CUDA(cudaMemcpyToSymbolAsync(mega1::c_mega, &a, sizeof(MegaArgs), 0, cudaMemcpyHostToDevice, c->stream));
        B2L_CUDA(cudaLaunchKernelEx(&cfg, mega_kernel_for(c->hd, c->group, false)));
    }
    c->launched++;
}

void mega_check(b2l_ctx* c, cudaError_t sync_result) {
    if (sync_result == cudaSuccess) return;
    const int code = c->mega_abort ? *c->mega_abort : 0;
    if (c->mega_abort && std::getenv("B2L_MEGA_DEBUG")) {
        std::fprintf(stderr, "megakernel progress markers (step*100000 + phase*100 + stage):");
        for (int i = 0; i < c->prop.multiProcessorCount; i++) std::fprintf(stderr, " %d", c->mega_abort[1 + i]);
        std::fprintf(stderr, "\n");
    }
    throw Error(std::string("megakernel failed: ") + cudaGetErrorString(sync_result) + " (abort code " + std::to_string(code) +
                "; 100 = grid barrier timeout, 2xx = consumer wait, 3xx = producer wait)");
}

void ensure_out_ids(b2l_ctx* c, int n) {
    if (c->out_ids_cap >= n) return;
    c->d_out_ids = dalloc<int32_t>(c, static_cast<size_t>(n));
    c->out_ids_cap = n;
    for (auto it = c->decode_graphs.begin(); it != c->decode_graphs.end();) {  // captured advance kernels hold the old pointer
        if (it->first >= 1000) {
            cudaGraphExecDestroy(it->second.exec);
            cudaGraphDestroy(it->second.graph);
            it = c->decode_graphs.erase(it);
        } else {
            ++it;
        }
    }
}

void require_ready(b2l_ctx* c) { B2L_CHECK(c->finalized, "b2l_finalize has not been called"); }

template <typename F>
int guarded(b2l_ctx* c, F&& f) {
    if (!c) return 1;
    std::lock_guard<std::mutex> lock(c->mu);
    try {
        B2L_CUDA(cudaSetDevice(c->p.device));
        f();
        return 0;
    } catch (const std::exception& e) {
        c->err = e.what();
        return 1;
    }
}

// ---- tcgen05 GEMM host side: TMA tensor maps + launch -------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        B2L_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        B2L_CHECK(p && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major bf16 [rows][cols] -> 2-D map with a (64 elements = 128 B) x box_rows box, 128-byte swizzle
CUtensorMap make_kmajor_map(const uint16_t* ptr, int64_t rows, int64_t cols, int box_rows) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kGemmBK), static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(ptr), dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B2L_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string(static_cast<int>(r)) + ")");
    return m;
}

// C = A[M][K] * W[N][K]^T on the tensor cores; epilogue per GemmEpilogue
void gemm_bf16(b2l_ctx* c, const uint16_t* A, const uint16_t* W, GemmArgs g) {
    B2L_CHECK(g.K % kGemmBK == 0 && g.K >= kGemmBK, "gemm: K must be a multiple of 64");
    B2L_CHECK(g.N % 128 == 0, "gemm: N must be a multiple of 128");
    static const bool persistent_ok = [] { const char* e = std::getenv("B2L_GEMM_PERSISTENT"); return !e || std::atoi(e) != 0; }();   // development switch
    if (persistent_ok && g.N % kGemmPBN == 0) {
        // persistent 128 x 256 tiles, one CTA per SM (gemm_bf16_persistent_kernel)
        const CUtensorMap ma = make_kmajor_map(A, g.M, g.K, kGemmBM), mw = make_kmajor_map(W, g.N, g.K, kGemmPBN);
        ensure_smem_optin(reinterpret_cast<const void*>(gemm_bf16_persistent_kernel), c->p.device, kGemmPSmem);
        const int total = ((g.M + kGemmBM - 1) / kGemmBM) * (g.N / kGemmPBN);
        gemm_bf16_persistent_kernel<<<std::min(total, c->prop.multiProcessorCount), kGemmPThreads, kGemmPSmem, c->stream>>>(ma, mw, g);
        B2L_CUDA(cudaGetLastError());
        c->launched++;
        return;
    }
    constexpr int BN = 128;
    const CUtensorMap ma = make_kmajor_map(A, g.M, g.K, kGemmBM), mw = make_kmajor_map(W, g.N, g.K, BN);
    const size_t smem = static_cast<size_t>(kGemmStages) * (kGemmBM * kGemmBK * 2 + BN * kGemmBK * 2) + 16 * kGemmStages + 64 + 1024;
    ensure_smem_optin(reinterpret_cast<const void*>(gemm_bf16_tcgen05_kernel<BN>), c->p.device, smem);
    const dim3 grid(g.N / BN, (g.M + kGemmBM - 1) / kGemmBM);
    gemm_bf16_tcgen05_kernel<BN><<<grid, kGemmThreads, smem, c->stream>>>(ma, mw, g);
    B2L_CUDA(cudaGetLastError());
    c->launched++;
}

// a throwaway mini-context (stream, events, allocation list) for the single-op entry points
int op_guard(int device, const std::function<void(b2l_ctx*)>& f) {
    b2l_ctx* c = nullptr;
    try {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw Error(std::string("no CUDA device (") + cudaGetErrorString(e) + "): the B200 path has no CPU fallback");
        B2L_CHECK(device >= 0 && device < ndev, "device ordinal out of range");
        B2L_CUDA(cudaSetDevice(device));
        c = new b2l_ctx;
        c->p.device = device;
        B2L_CUDA(cudaGetDeviceProperties(&c->prop, device));
        B2L_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        B2L_CUDA(cudaEventCreate(&c->ev0));
        B2L_CUDA(cudaEventCreate(&c->ev1));
        f(c);
        b2l_destroy(c);
        return 0;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        if (c) b2l_destroy(c);
        return 1;
    }
}

}  // namespace

extern "C" {

const char* b2l_last_error(const b2l_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int b2l_nccl_unique_id(void* out_bytes) {
    try {
        B2L_CHECK(out_bytes, "null argument");
        B2L_CHECK(nccl().ok, nccl().why);
        NcclApi::Id id;
        B2L_NCCL(nccl().GetUniqueId(&id));
        std::memcpy(out_bytes, id.internal, B2L_NCCL_ID_BYTES);
        return 0;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return 1;
    }
}

int b2l_create(const b2l_params* p, const float* rope_cos_sin, const void* nccl_unique_id, b2l_ctx** out) {
    b2l_ctx* c = nullptr;
    try {
        B2L_CHECK(p && out && rope_cos_sin, "null argument");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw Error(std::string("no CUDA device (") + cudaGetErrorString(e) + "): the B200 path has no CPU fallback");
        B2L_CHECK(p->device >= 0 && p->device < ndev, "device ordinal out of range");
        B2L_CHECK(p->tp_size >= 1 && p->tp_rank >= 0 && p->tp_rank < p->tp_size, "bad tp_rank / tp_size");
        B2L_CHECK(p->tp_size == 1 || nccl_unique_id, "tp_size > 1 needs an NCCL unique id");
        B2L_CHECK(p->head_dim == 32 || p->head_dim == 64 || p->head_dim == 128, "head_dim must be 32, 64 or 128");
        B2L_CHECK(p->num_heads % p->num_kv_heads == 0, "num_heads must be a multiple of num_kv_heads");
        B2L_CHECK(p->num_kv_heads % p->tp_size == 0 && p->intermediate_size % p->tp_size == 0 && p->vocab_size % p->tp_size == 0,
                  "kv heads, intermediate size and vocab must divide by tp_size");
        B2L_CHECK(p->hidden_size % 256 == 0, "hidden_size must be a multiple of 256");
        B2L_CHECK((p->intermediate_size / p->tp_size) % 8 == 0, "local intermediate size must be a multiple of 8");
        B2L_CHECK(p->page_size == 16 || p->page_size == 32 || p->page_size == 64, "page_size must be 16, 32 or 64");
        B2L_CHECK(p->max_batch >= 1 && p->max_batch <= 64, "max_batch must be in [1, 64]");
        B2L_CHECK(p->num_pages >= 1 && p->max_positions >= 1 && p->max_prefill_tokens >= 1, "bad sizing");
        B2L_CUDA(cudaSetDevice(p->device));
        c = new b2l_ctx;
        c->p = *p;
        B2L_CUDA(cudaGetDeviceProperties(&c->prop, p->device));
        B2L_CHECK(c->prop.major >= 10, "this library is built for sm_100a (B200) only");
        B2L_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        B2L_CUDA(cudaEventCreate(&c->ev0));
        B2L_CUDA(cudaEventCreate(&c->ev1));

        const int tp = p->tp_size;
        c->H = p->hidden_size; c->L = p->num_layers; c->hd = p->head_dim; c->V = p->vocab_size;
        c->I_l = p->intermediate_size / tp; c->nh_l = p->num_heads / tp; c->nkv_l = p->num_kv_heads / tp;
        c->V_l = p->vocab_size / tp;
        c->qd_l = c->nh_l * c->hd; c->kvd_l = c->nkv_l * c->hd; c->qkv_l = c->qd_l + 2 * c->kvd_l;
        c->group = p->num_heads / p->num_kv_heads;
        c->max_rows = (std::max(p->max_batch, 8) + 7) / 8 * 8;
        c->max_blocks_cap = (p->max_positions + p->page_size - 1) / p->page_size;
        c->nsplit = std::max(1, std::min(64, (8 * c->prop.multiProcessorCount + c->nkv_l - 1) / c->nkv_l));   // capacity of the split-K partial buffers

        // weights
        const size_t H = c->H;
        c->embed = dalloc<uint16_t>(c, static_cast<size_t>(c->V) * H);
        c->weight_bytes += static_cast<int64_t>(c->V) * H * 2;
        if (p->tie_word_embeddings) {
            c->lm_head = c->embed + static_cast<size_t>(p->tp_rank) * c->V_l * H;
        } else {
            c->lm_head = dalloc<uint16_t>(c, static_cast<size_t>(c->V_l) * H);
            c->weight_bytes += static_cast<int64_t>(c->V_l) * H * 2;
        }
        c->final_norm = dalloc<uint16_t>(c, H);
        c->layers.resize(c->L);
        const size_t page_elems = static_cast<size_t>(2) * p->page_size * c->kvd_l;
        // one allocation for every layer's KV pool: lets one L2 persisting window cover the whole cache
        c->kv_base = dalloc<uint16_t>(c, page_elems * p->num_pages * c->L);
        B2L_CUDA(cudaMemset(c->kv_base, 0, sizeof(uint16_t) * page_elems * p->num_pages * c->L));   // timing runs decode over pages nobody prefilled
        size_t kv_off = 0;
        for (auto& w : c->layers) {
            w.in_norm = dalloc<uint16_t>(c, H);
            w.post_norm = dalloc<uint16_t>(c, H);
            w.w_qkv = dalloc<uint16_t>(c, static_cast<size_t>(c->qkv_l) * H);
            w.w_o = dalloc<uint16_t>(c, H * c->qd_l);
            w.w_gu = dalloc<uint16_t>(c, static_cast<size_t>(2) * c->I_l * H);
            w.w_down = dalloc<uint16_t>(c, H * c->I_l);
            w.kv_pool = c->kv_base + kv_off;
            kv_off += page_elems * p->num_pages;
            c->weight_bytes += static_cast<int64_t>(2) * (2 * H + static_cast<size_t>(c->qkv_l) * H + H * c->qd_l + 3 * H * c->I_l);
            c->kv_bytes += static_cast<int64_t>(2) * page_elems * p->num_pages;
        }
        c->weight_bytes += 2 * H;
        c->rope = dalloc<float>(c, static_cast<size_t>(p->max_positions) * c->hd);
        B2L_CUDA(cudaMemcpy(c->rope, rope_cos_sin, sizeof(float) * p->max_positions * c->hd, cudaMemcpyHostToDevice));

        // per-call inputs
        const int R = c->max_rows;
        c->d_tokens = dalloc<int32_t>(c, R); c->d_positions = dalloc<int32_t>(c, R); c->d_slots = dalloc<int32_t>(c, R);
        c->d_next_ids = dalloc<int32_t>(c, R); c->d_step = dalloc<int32_t>(c, 1);
        c->d_block_tables = dalloc<int32_t>(c, static_cast<size_t>(p->max_batch) * c->max_blocks_cap);
        c->h_stage_ints = 4 * static_cast<size_t>(R) + static_cast<size_t>(p->max_batch) * c->max_blocks_cap;
        B2L_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->h_stage), c->h_stage_ints * sizeof(int32_t)));

        // activations
        c->h = dalloc<float>(c, static_cast<size_t>(R) * H);
        c->qkv = dalloc<float>(c, static_cast<size_t>(R) * c->qkv_l);
        c->attn = dalloc<float>(c, static_cast<size_t>(R) * c->qd_l);
        c->act = dalloc<float>(c, static_cast<size_t>(R) * c->I_l);
        c->proj = dalloc<float>(c, static_cast<size_t>(R) * H);
        c->logits = dalloc<float>(c, static_cast<size_t>(R) * c->V_l);
        c->seq_logits = dalloc<float>(c, static_cast<size_t>(p->max_batch) * c->V_l);
        c->part_acc = dalloc<float>(c, static_cast<size_t>(R) * c->nkv_l * c->nsplit * c->group * c->hd);
        c->part_ml = dalloc<float>(c, static_cast<size_t>(R) * c->nkv_l * c->nsplit * c->group * 2);
        c->attn_counters = dalloc<int>(c, static_cast<size_t>(R) * c->nkv_l);
        B2L_CUDA(cudaMemset(c->attn_counters, 0, sizeof(int) * R * c->nkv_l));
        B2L_CUDA(cudaMemset(c->h, 0, sizeof(float) * R * H));
        B2L_CUDA(cudaMemset(c->qkv, 0, sizeof(float) * R * c->qkv_l));
        B2L_CUDA(cudaMemset(c->attn, 0, sizeof(float) * R * c->qd_l));
        B2L_CUDA(cudaMemset(c->act, 0, sizeof(float) * R * c->I_l));
        if (tp > 1) {
            B2L_CHECK(nccl().ok, nccl().why);
            c->tp_pack = dalloc<float>(c, static_cast<size_t>(R) * 2);
            c->tp_gather = dalloc<float>(c, static_cast<size_t>(R) * 2 * tp);
            c->tp_vals = dalloc<float>(c, R);
            NcclApi::Id id;
            std::memcpy(id.internal, nccl_unique_id, B2L_NCCL_ID_BYTES);
            B2L_NCCL(nccl().CommInitRank(&c->nccl_comm, tp, id, p->tp_rank));   // collective: every rank is in b2l_create now
            tp_peer_setup(c);
        }
        B2L_CUDA(cudaDeviceSynchronize());
        *out = c;
        return 0;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        if (c) b2l_destroy(c);
        if (out) *out = nullptr;
        return 1;
    }
}

void b2l_destroy(b2l_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->p.device);
    cudaDeviceSynchronize();
    for (auto& kv : c->decode_graphs) {
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
    }
    if (c->tp_peer_ok && c->nccl_comm) {
        // nobody may free a slab a peer can still write to: meet once more, then unmap
        float* d = reinterpret_cast<float*>(c->tp_done);
        if (nccl().AllReduce(d, d, 1, kNcclFloat32, kNcclSum, c->nccl_comm, c->stream) == 0) cudaStreamSynchronize(c->stream);
        for (int p = 0; p < c->p.tp_size; p++)
            if (p != c->p.tp_rank && c->tp_peer[p]) cudaIpcCloseMemHandle(c->tp_peer[p]);
    }
    if (c->nccl_comm) nccl().CommDestroy(c->nccl_comm);
    for (void* p : c->allocs) cudaFree(p);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->mega_abort) cudaFreeHost(c->mega_abort);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int b2l_upload_tensor(b2l_ctx* c, const char* hf_name, const void* host_bf16, const int64_t* shape, int ndim) {
    return guarded(c, [&] {
        B2L_CHECK(hf_name && host_bf16 && shape, "null argument");
        B2L_CHECK(!c->finalized, "tensors cannot change after b2l_finalize");
        int layer;
        const Kind k = parse_name(hf_name, &layer);
        B2L_CHECK(k != K_BAD, std::string("unknown tensor name: ") + hf_name);
        Placement pl;
        try {
            pl = place_tensor(c, k, layer, shape, ndim);
        } catch (const Error& e) {
            throw Error(std::string(hf_name) + ": " + e.what());
        }
        const uint16_t* src = static_cast<const uint16_t*>(host_bf16) + pl.row0 * pl.full_cols + pl.col0;
        B2L_CUDA(cudaMemcpy2D(pl.dst, pl.dst_stride * 2, src, pl.full_cols * 2, pl.ncols * 2, pl.nrows, cudaMemcpyHostToDevice));
        mark_tensor(c, k, layer);
    });
}

int b2l_synth_tensor(b2l_ctx* c, const char* hf_name, const int64_t* shape, int ndim, uint32_t tensor_seed, float scale,
                     float offset) {
    return guarded(c, [&] {
        B2L_CHECK(hf_name && shape, "null argument");
        B2L_CHECK(!c->finalized, "tensors cannot change after b2l_finalize");
        int layer;
        const Kind k = parse_name(hf_name, &layer);
        B2L_CHECK(k != K_BAD, std::string("unknown tensor name: ") + hf_name);
        const Placement pl = place_tensor(c, k, layer, shape, ndim);
        const int64_t n = pl.nrows * pl.ncols;
        const int threads = 256;
        const int blocks = static_cast<int>(std::min<int64_t>((n + threads - 1) / threads, 148 * 32));
        synth_fill_kernel<<<blocks, threads, 0, c->stream>>>(pl.dst, pl.dst_stride, pl.row0, pl.nrows, pl.col0, pl.ncols,
                                                             pl.full_cols, tensor_seed, scale, offset);
        B2L_CUDA(cudaGetLastError());
        B2L_CUDA(cudaStreamSynchronize(c->stream));
        mark_tensor(c, k, layer);
    });
}

int b2l_shard_window(const b2l_params* p, const char* hf_name, const int64_t* shape, int ndim, int64_t win[4]) {
    try {
        B2L_CHECK(p && hf_name && shape && win, "null argument");
        B2L_CHECK(p->tp_size >= 1 && p->tp_rank >= 0 && p->tp_rank < p->tp_size, "bad tp_rank / tp_size");
        B2L_CHECK(p->num_kv_heads % p->tp_size == 0 && p->intermediate_size % p->tp_size == 0 && p->vocab_size % p->tp_size == 0,
                  "kv heads, intermediate size and vocab must divide by tp_size");
        b2l_ctx c;   // dims only; no device state is touched
        c.p = *p;
        const int tp = p->tp_size;
        c.H = p->hidden_size; c.L = p->num_layers; c.hd = p->head_dim; c.V = p->vocab_size;
        c.I_l = p->intermediate_size / tp; c.nh_l = p->num_heads / tp; c.nkv_l = p->num_kv_heads / tp; c.V_l = p->vocab_size / tp;
        c.qd_l = c.nh_l * c.hd; c.kvd_l = c.nkv_l * c.hd; c.qkv_l = c.qd_l + 2 * c.kvd_l;
        c.layers.resize(c.L);
        int layer;
        const Kind k = parse_name(hf_name, &layer);
        B2L_CHECK(k != K_BAD, std::string("unknown tensor name: ") + hf_name);
        const Placement pl = place_tensor(&c, k, layer, shape, ndim);
        win[0] = pl.row0; win[1] = pl.nrows; win[2] = pl.col0; win[3] = pl.ncols;
        return 0;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return 1;
    }
}

int b2l_is_model_tensor(const char* hf_name) {
    int layer;
    return hf_name && parse_name(hf_name, &layer) != K_BAD ? 1 : 0;
}

int b2l_finalize(b2l_ctx* c) {
    return guarded(c, [&] {
        B2L_CHECK(c->have_embed, "missing model.embed_tokens.weight");
        B2L_CHECK(c->have_final_norm, "missing model.norm.weight");
        B2L_CHECK(c->have_lm_head, "missing lm_head.weight (model is not tied)");
        const uint32_t all = (1u << K_IN_NORM) | (1u << K_Q) | (1u << K_K) | (1u << K_V) | (1u << K_O) | (1u << K_POST_NORM) |
                             (1u << K_GATE) | (1u << K_UP) | (1u << K_DOWN);
        for (int l = 0; l < c->L; l++)
            B2L_CHECK(c->layers[l].have == all, "layer " + std::to_string(l) + " is missing tensors");
        B2L_CUDA(cudaDeviceSynchronize());
        ensure_out_ids(c, 256);
        c->pf_ok = c->qkv_l % 128 == 0 && c->H % 128 == 0 && (2 * c->I_l) % 128 == 0 && c->H % 64 == 0 && c->qd_l % 64 == 0 && c->I_l % 64 == 0;
        skinny_setup(c);
        mega_setup(c);
        if (c->p.tp_size > 1) {
            // ranks load at different speeds: meet here so that the first forward starts roughly together
            float* d = reinterpret_cast<float*>(c->tp_done);
            B2L_NCCL(nccl().AllReduce(d, d, 1, kNcclFloat32, kNcclSum, c->nccl_comm, c->stream));
            B2L_CUDA(cudaStreamSynchronize(c->stream));
        }
        c->decode_mode = c->mega_ok ? 1 : 0;
        c->finalized = true;
    });
}

int b2l_decode(b2l_ctx* c, int n_seq, const int32_t* tokens, const int32_t* positions, const int32_t* block_tables,
               int max_blocks, int32_t* next_ids) {
    return guarded(c, [&] {
        require_ready(c);
        B2L_CHECK(tokens && positions && block_tables && next_ids, "null argument");
        B2L_CHECK(n_seq >= 1 && n_seq <= c->p.max_batch, "n_seq out of range");
        std::vector<int32_t> slots(n_seq), need(n_seq);
        for (int i = 0; i < n_seq; i++) {
            B2L_CHECK(positions[i] >= 0 && positions[i] < c->p.max_positions, "position out of range");
            B2L_CHECK(tokens[i] >= 0 && tokens[i] < c->V, "token id out of range");
            slots[i] = i;
            need[i] = positions[i] + 1;
        }
        const bool mega = !c->taps && c->decode_mode == 1 && n_seq == 1;
        if (mega) {
            // single-step megakernel call: token and position go in, the id comes back, through mapped pinned memory;
            // the block table is re-uploaded only when it changed. Stream work = constant upload + one launch.
            upload_block_tables(c, 1, block_tables, max_blocks, need.data());
            volatile int* io = c->mega_abort + kMegaIoWord;
            io[0] = tokens[0];
            io[1] = positions[0];
            io[2] = -1;
            {
                std::lock_guard<std::mutex> dev_lock(mega_device_mutex(c->p.device));
                mega_enqueue(c, 1, true);
                mega_check(c, cudaStreamSynchronize(c->stream));
            }
            B2L_CHECK(io[2] >= 0, "megakernel returned no token");
            next_ids[0] = io[2];
        } else {
            upload_rows_meta(c, n_seq, tokens, positions, slots.data());
            upload_block_tables(c, n_seq, block_tables, max_blocks, need.data());
            if (c->taps) {
                c->tap_rows = n_seq;
                enqueue_forward(c, n_seq, true, 0);
            } else {
                Graph& g = decode_graph(c, n_seq, false);
                B2L_CUDA(cudaGraphLaunch(g.exec, c->stream));
                c->launched += g.nodes;
            }
            int32_t* h_next = c->h_stage + 3 * c->max_rows + static_cast<size_t>(c->p.max_batch) * c->max_blocks_cap;
            B2L_CUDA(cudaMemcpyAsync(h_next, c->d_next_ids, sizeof(int32_t) * n_seq, cudaMemcpyDeviceToHost, c->stream));
            B2L_CUDA(cudaStreamSynchronize(c->stream));
            std::memcpy(next_ids, h_next, sizeof(int32_t) * n_seq);
        }
        c->logits_src = c->logits;
        c->logits_rows = n_seq;
    });
}

int b2l_decode_loop(b2l_ctx* c, int n_seq, const int32_t* tokens, const int32_t* positions, const int32_t* block_tables,
                    int max_blocks, int n_steps, int32_t* out_ids, float* device_ms) {
    return guarded(c, [&] {
        require_ready(c);
        B2L_CHECK(tokens && positions && block_tables && out_ids, "null argument");
        B2L_CHECK(n_seq >= 1 && n_seq <= c->p.max_batch, "n_seq out of range");
        B2L_CHECK(n_steps >= 1, "n_steps must be >= 1");
        std::vector<int32_t> slots(n_seq), need(n_seq);
        for (int i = 0; i < n_seq; i++) {
            B2L_CHECK(positions[i] >= 0 && positions[i] + n_steps <= c->p.max_positions, "position + n_steps exceeds max_positions");
            B2L_CHECK(tokens[i] >= 0 && tokens[i] < c->V, "token id out of range");
            slots[i] = i;
            need[i] = positions[i] + n_steps;
        }
        ensure_out_ids(c, n_steps * n_seq);
        upload_rows_meta(c, n_seq, tokens, positions, slots.data());
        upload_block_tables(c, n_seq, block_tables, max_blocks, need.data());
        const bool mega = c->decode_mode == 1 && n_seq == 1;
        if (mega) {
            std::lock_guard<std::mutex> dev_lock(mega_device_mutex(c->p.device));
            B2L_CUDA(cudaEventRecord(c->ev0, c->stream));
            mega_enqueue(c, n_steps);
            B2L_CUDA(cudaEventRecord(c->ev1, c->stream));
            mega_check(c, cudaMemcpyAsync(out_ids, c->d_out_ids, sizeof(int32_t) * n_steps * n_seq, cudaMemcpyDeviceToHost, c->stream));
            mega_check(c, cudaStreamSynchronize(c->stream));
        } else {
            B2L_CUDA(cudaMemsetAsync(c->d_step, 0, sizeof(int32_t), c->stream));
            Graph& g = decode_graph(c, n_seq, true);
            B2L_CUDA(cudaEventRecord(c->ev0, c->stream));
            for (int s = 0; s < n_steps; s++) B2L_CUDA(cudaGraphLaunch(g.exec, c->stream));
            B2L_CUDA(cudaEventRecord(c->ev1, c->stream));
            c->launched += static_cast<int64_t>(g.nodes) * n_steps;
        }
        if (!mega) {
            B2L_CUDA(cudaMemcpyAsync(out_ids, c->d_out_ids, sizeof(int32_t) * n_steps * n_seq, cudaMemcpyDeviceToHost, c->stream));
            B2L_CUDA(cudaStreamSynchronize(c->stream));
        }
        if (device_ms) B2L_CUDA(cudaEventElapsedTime(device_ms, c->ev0, c->ev1));
        c->logits_src = c->logits;
        c->logits_rows = n_seq;
    });
}

int b2l_prefill(b2l_ctx* c, int n_seq, const int32_t* tokens, const int32_t* q_lens, const int32_t* ctx_lens,
                const int32_t* block_tables, int max_blocks, int32_t* next_ids) {
    return guarded(c, [&] {
        require_ready(c);
        B2L_CHECK(tokens && q_lens && ctx_lens && block_tables && next_ids, "null argument");
        B2L_CHECK(n_seq >= 1 && n_seq <= c->p.max_batch, "n_seq out of range");
        std::vector<int32_t> need(n_seq);
        int64_t total = 0;
        for (int i = 0; i < n_seq; i++) {
            B2L_CHECK(q_lens[i] >= 1 && ctx_lens[i] >= 0, "q_lens must be >= 1 and ctx_lens >= 0");
            B2L_CHECK(ctx_lens[i] + q_lens[i] <= c->p.max_positions, "sequence exceeds max_positions");
            need[i] = ctx_lens[i] + q_lens[i];
            total += q_lens[i];
        }
        B2L_CHECK(total <= c->p.max_prefill_tokens, "prefill exceeds max_prefill_tokens");
        for (int64_t i = 0; i < total; i++) B2L_CHECK(tokens[i] >= 0 && tokens[i] < c->V, "token id out of range");
        if (c->taps) B2L_CHECK(total <= c->tap_rows_cap, "taps: prefill larger than the tap buffer");
        upload_block_tables(c, n_seq, block_tables, max_blocks, need.data());
        const bool use_gemm = c->pf_ok && (c->prefill_mode == 1 || (c->prefill_mode < 0 && total >= 64));
        if (use_gemm) {
            prefill_alloc(c);
            std::vector<int32_t> pos(total), slot(total), last(n_seq);
            int64_t t0 = 0;
            for (int i = 0; i < n_seq; i++) {
                for (int j = 0; j < q_lens[i]; j++) {
                    pos[t0 + j] = ctx_lens[i] + j;
                    slot[t0 + j] = i;
                }
                t0 += q_lens[i];
                last[i] = static_cast<int32_t>(t0 - 1);
            }
            B2L_CUDA(cudaMemcpyAsync(c->pf_tokens, tokens, sizeof(int32_t) * total, cudaMemcpyHostToDevice, c->stream));
            B2L_CUDA(cudaMemcpyAsync(c->pf_positions, pos.data(), sizeof(int32_t) * total, cudaMemcpyHostToDevice, c->stream));
            B2L_CUDA(cudaMemcpyAsync(c->pf_slots, slot.data(), sizeof(int32_t) * total, cudaMemcpyHostToDevice, c->stream));
            B2L_CUDA(cudaMemcpyAsync(c->pf_last, last.data(), sizeof(int32_t) * n_seq, cudaMemcpyHostToDevice, c->stream));
            std::vector<PrefillTile> tiles;   // 64 consecutive positions of one sequence per attention CTA
            int64_t row = 0;
            for (int i = 0; i < n_seq; i++) {
                for (int j = 0; j < q_lens[i]; j += kFlashBM)
                    tiles.push_back(PrefillTile{static_cast<int>(row + j), std::min(kFlashBM, q_lens[i] - j), ctx_lens[i] + j, i});
                row += q_lens[i];
            }
            c->pf_n_tiles = static_cast<int>(tiles.size());
            B2L_CUDA(cudaMemcpyAsync(c->pf_tiles, tiles.data(), sizeof(PrefillTile) * tiles.size(), cudaMemcpyHostToDevice, c->stream));
            std::vector<PrefillTile> tiles_tc;   // 256 consecutive positions per CTA of the tcgen05 attention, longest contexts first
            row = 0;
            for (int i = 0; i < n_seq; i++) {
                for (int j = 0; j < q_lens[i]; j += 2 * kFtcBQ)
                    tiles_tc.push_back(PrefillTile{static_cast<int>(row + j), std::min(2 * kFtcBQ, q_lens[i] - j), ctx_lens[i] + j, i});
                row += q_lens[i];
            }
            std::stable_sort(tiles_tc.begin(), tiles_tc.end(), [](const PrefillTile& x, const PrefillTile& y) { return x.pos0 + x.n_rows > y.pos0 + y.n_rows; });
            c->pf_n_tiles_tc = static_cast<int>(tiles_tc.size());
            B2L_CUDA(cudaMemcpyAsync(c->pf_tiles_tc, tiles_tc.data(), sizeof(PrefillTile) * tiles_tc.size(), cudaMemcpyHostToDevice, c->stream));
            B2L_CUDA(cudaStreamSynchronize(c->stream));   // pos/slot/last are stack vectors
            prefill_gemm(c, static_cast<int>(total), n_seq, c->taps ? 0 : -1);
        }
        // exact-activation prefill: chunks of <= 8 consecutive positions through the decode kernels
        // (each chunk appends its K/V before its attention runs, so causality holds inside it).
        int64_t tok0 = 0;
        for (int i = 0; i < n_seq && !use_gemm; i++) {
            for (int c0 = 0; c0 < q_lens[i]; c0 += 8) {
                const int R = std::min(8, q_lens[i] - c0);
                int32_t pos[8], slot[8];
                for (int r = 0; r < R; r++) {
                    pos[r] = ctx_lens[i] + c0 + r;
                    slot[r] = i;
                }
                B2L_CUDA(cudaStreamSynchronize(c->stream));  // staging buffer reuse
                upload_rows_meta(c, R, tokens + tok0 + c0, pos, slot);
                const bool last = c0 + R >= q_lens[i];
                enqueue_forward(c, R, last, c->taps ? static_cast<int>(tok0 + c0) : -1);
                if (last)
                    B2L_CUDA(cudaMemcpyAsync(c->seq_logits + static_cast<size_t>(i) * c->V_l,
                                             c->logits + static_cast<size_t>(R - 1) * c->V_l, sizeof(float) * c->V_l,
                                             cudaMemcpyDeviceToDevice, c->stream));
            }
            tok0 += q_lens[i];
        }
        tp_argmax(c, c->seq_logits, n_seq);
        int32_t* h_next = c->h_stage + 3 * c->max_rows + static_cast<size_t>(c->p.max_batch) * c->max_blocks_cap;
        B2L_CUDA(cudaMemcpyAsync(h_next, c->d_next_ids, sizeof(int32_t) * n_seq, cudaMemcpyDeviceToHost, c->stream));
        B2L_CUDA(cudaStreamSynchronize(c->stream));
        std::memcpy(next_ids, h_next, sizeof(int32_t) * n_seq);
        c->logits_src = c->seq_logits;
        c->logits_rows = n_seq;
        c->tap_rows = static_cast<int>(total);
    });
}

int b2l_get_logits(b2l_ctx* c, int row0, int n_rows, float* out) {
    return guarded(c, [&] {
        B2L_CHECK(out && c->logits_src, "no logits available (run prefill or decode first)");
        B2L_CHECK(row0 >= 0 && n_rows >= 1 && row0 + n_rows <= c->logits_rows, "logit rows out of range");
        B2L_CUDA(cudaStreamSynchronize(c->stream));
        if (c->p.tp_size == 1) {
            B2L_CUDA(cudaMemcpy(out, c->logits_src + static_cast<size_t>(row0) * c->V_l, sizeof(float) * n_rows * c->V_l,
                                cudaMemcpyDeviceToHost));
        } else {
            // collective: every rank calls b2l_get_logits with the same rows; vocab shards are gathered
            const int tp = c->p.tp_size;
            const size_t shard = static_cast<size_t>(n_rows) * c->V_l;
            if (!c->tp_logits) c->tp_logits = dalloc<float>(c, static_cast<size_t>(std::max(c->max_rows, c->p.max_batch)) * c->V_l * tp);
            B2L_NCCL(nccl().AllGather(c->logits_src + static_cast<size_t>(row0) * c->V_l, c->tp_logits, shard, kNcclFloat32, c->nccl_comm, c->stream));
            B2L_CUDA(cudaStreamSynchronize(c->stream));
            std::vector<float> tmp(shard * tp);
            B2L_CUDA(cudaMemcpy(tmp.data(), c->tp_logits, sizeof(float) * shard * tp, cudaMemcpyDeviceToHost));
            for (int r = 0; r < tp; r++)
                for (int row = 0; row < n_rows; row++)
                    std::memcpy(out + static_cast<size_t>(row) * c->V + static_cast<size_t>(r) * c->V_l,
                                tmp.data() + (static_cast<size_t>(r) * n_rows + row) * c->V_l, sizeof(float) * c->V_l);
        }
    });
}

int b2l_set_taps(b2l_ctx* c, int enable) {
    return guarded(c, [&] {
        if (enable && !c->tap) {
            c->tap_rows_cap = std::max(c->p.max_prefill_tokens, c->max_rows);
            c->tap = dalloc<float>(c, static_cast<size_t>(c->L + 2) * c->tap_rows_cap * c->H);
        }
        c->taps = enable != 0;
    });
}

int b2l_get_hidden(b2l_ctx* c, int slab, int row0, int n_rows, float* out) {
    return guarded(c, [&] {
        B2L_CHECK(out && c->tap, "taps are not enabled");
        B2L_CHECK(slab >= 0 && slab <= c->L + 1, "slab out of range");
        B2L_CHECK(row0 >= 0 && n_rows >= 1 && row0 + n_rows <= c->tap_rows, "rows out of range");
        B2L_CUDA(cudaStreamSynchronize(c->stream));
        B2L_CUDA(cudaMemcpy(out, c->tap + (static_cast<size_t>(slab) * c->tap_rows_cap + row0) * c->H,
                            sizeof(float) * n_rows * c->H, cudaMemcpyDeviceToHost));
    });
}

int b2l_get_kv_page(b2l_ctx* c, int layer, int page, int which, void* out_bf16) {
    return guarded(c, [&] {
        B2L_CHECK(out_bf16 && layer >= 0 && layer < c->L && page >= 0 && page < c->p.num_pages && (which == 0 || which == 1),
                  "bad kv page request");
        const KvLayout kv{c->layers[layer].kv_pool, c->p.page_size, c->kvd_l};
        B2L_CUDA(cudaStreamSynchronize(c->stream));
        B2L_CUDA(cudaMemcpy(out_bf16, c->layers[layer].kv_pool + ((static_cast<size_t>(page) * 2 + which) * c->p.page_size) * c->kvd_l,
                            sizeof(uint16_t) * c->p.page_size * c->kvd_l, cudaMemcpyDeviceToHost));
        (void)kv;
    });
}

int b2l_get_info(b2l_ctx* c, b2l_info* out) {
    return guarded(c, [&] {
        B2L_CHECK(out, "null argument");
        std::memset(out, 0, sizeof(*out));
        out->abi_version = B2L_ABI_VERSION;
        out->sm_count = c->prop.multiProcessorCount;
        out->cc_major = c->prop.major;
        out->cc_minor = c->prop.minor;
        out->hbm_bytes = static_cast<int64_t>(c->prop.totalGlobalMem);
        out->weight_bytes = c->weight_bytes;
        out->kv_bytes = c->kv_bytes;
        const int64_t H = c->H;
        const int64_t per_layer = 2 * (2 * H + static_cast<int64_t>(c->qkv_l) * H + H * c->qd_l + 3 * H * c->I_l);
        out->stream_bytes_per_token = per_layer * c->L + 2 * H + 2 * static_cast<int64_t>(c->V_l) * H + 2 * H;
        out->kernels_launched = c->launched;
        out->decode_mode = c->decode_mode;
        out->batched_tensor_core = c->skinny_ok ? 1 : 0;
        out->tp_transport = c->p.tp_size == 1 ? 0 : c->tp_peer_ok ? 2 : 1;
        std::snprintf(out->device_name, sizeof(out->device_name), "%s", c->prop.name);
    });
}

int b2l_debug_mega_profile(b2l_ctx* c, int enable, uint64_t* out_ns, int* n_phases, int32_t* phase_types) {
    return guarded(c, [&] {
        require_ready(c);
        B2L_CHECK(c->mega_ok, "megakernel unavailable: " + c->mega_why);
        const size_t n = static_cast<size_t>(kMegaProfRows) * static_cast<size_t>(c->mega_n_phases + 1);
        if (enable && !c->mega_prof) {
            c->mega_prof = dalloc<unsigned long long>(c, n);
            B2L_CUDA(cudaMemset(c->mega_prof, 0, n * 8));
        }
        if (n_phases) *n_phases = c->mega_n_phases;
        if (out_ns && c->mega_prof) {
            B2L_CUDA(cudaStreamSynchronize(c->stream));
            B2L_CUDA(cudaMemcpy(out_ns, c->mega_prof, n * 8, cudaMemcpyDeviceToHost));
        }
        if (phase_types) {
            std::vector<MegaPhase> ph(c->mega_n_phases);
            B2L_CUDA(cudaMemcpy(ph.data(), c->mega_phases, sizeof(MegaPhase) * ph.size(), cudaMemcpyDeviceToHost));
            for (int i = 0; i < c->mega_n_phases; i++) phase_types[i] = ph[i].type;
        }
        if (!enable) c->mega_prof = nullptr;
    });
}

int b2l_set_prefill_mode(b2l_ctx* c, int mode) {
    return guarded(c, [&] {
        B2L_CHECK(mode >= -1 && mode <= 1, "prefill mode must be -1 (auto), 0 (chunked decode kernels) or 1 (tcgen05 GEMMs)");
        require_ready(c);
        B2L_CHECK(mode != 1 || c->pf_ok, "tcgen05 prefill unavailable: projection widths must be multiples of 128 (N) and 64 (K)");
        c->prefill_mode = mode;
    });
}

int b2l_set_decode_mode(b2l_ctx* c, int mode) {
    return guarded(c, [&] {
        B2L_CHECK(mode == 0 || mode == 1, "decode mode must be 0 (multi-kernel graph) or 1 (persistent megakernel)");
        require_ready(c);
        B2L_CHECK(mode == 0 || c->mega_ok, "megakernel unavailable: " + c->mega_why);
        c->decode_mode = mode;
    });
}

// ---- single-op entry points ---------------------------------------------------------------

int b2l_op_gemv(int device, const void* W_bf16, const float* x, float* y, const void* norm_w_bf16, float eps, int B, int N,
                int K, int mode, int iters, float* device_ms) {
    return op_guard(device, [&](b2l_ctx* c) {
        B2L_CHECK(W_bf16 && x && y, "null argument");
        B2L_CHECK(B >= 1 && B <= 64 && N >= 2 && K >= 8, "bad sizes");
        B2L_CHECK(!(mode == 2 && (N % 2)), "SwiGLU mode needs an even N");
        c->p.rms_norm_eps = eps;
        const int out_cols = mode == 2 ? N / 2 : N;
        const int Bp = (B + 7) / 8 * 8;
        uint16_t* dW = dalloc<uint16_t>(c, static_cast<size_t>(N) * K);
        uint16_t* dn = norm_w_bf16 ? dalloc<uint16_t>(c, K) : nullptr;
        float* dx = dalloc<float>(c, static_cast<size_t>(Bp) * K);
        float* dy = dalloc<float>(c, static_cast<size_t>(Bp) * out_cols);
        float* dy0 = dalloc<float>(c, static_cast<size_t>(Bp) * out_cols);
        B2L_CUDA(cudaMemset(dx, 0, sizeof(float) * Bp * K));
        B2L_CUDA(cudaMemset(dy0, 0, sizeof(float) * Bp * out_cols));
        B2L_CUDA(cudaMemcpy(dW, W_bf16, sizeof(uint16_t) * N * K, cudaMemcpyHostToDevice));
        if (dn) B2L_CUDA(cudaMemcpy(dn, norm_w_bf16, sizeof(uint16_t) * K, cudaMemcpyHostToDevice));
        B2L_CUDA(cudaMemcpy(dx, x, sizeof(float) * B * K, cudaMemcpyHostToDevice));
        B2L_CUDA(cudaMemcpy(dy0, y, sizeof(float) * B * out_cols, cudaMemcpyHostToDevice));  // residual input (mode 1)
        float total = 0.f;
        for (int it = 0; it < std::max(1, iters); it++) {
            B2L_CUDA(cudaMemcpyAsync(dy, dy0, sizeof(float) * Bp * out_cols, cudaMemcpyDeviceToDevice, c->stream));
            B2L_CUDA(cudaEventRecord(c->ev0, c->stream));
            gemv(c, dW, dx, K, dy, out_cols, dn, N, K, mode, B);
            B2L_CUDA(cudaEventRecord(c->ev1, c->stream));
            B2L_CUDA(cudaStreamSynchronize(c->stream));
            float ms;
            B2L_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
            if (it > 0 || iters <= 1) total += ms;
        }
        if (device_ms) *device_ms = total / std::max(1, iters > 1 ? iters - 1 : 1);
        B2L_CUDA(cudaMemcpy(y, dy, sizeof(float) * B * out_cols, cudaMemcpyDeviceToHost));
    });
}

int b2l_op_gemm_bf16(int device, const void* A_bf16, const void* W_bf16, float* C, int M, int N, int K, int epilogue, int iters,
                     float* device_ms) {
    return op_guard(device, [&](b2l_ctx* c) {
        B2L_CHECK(A_bf16 && W_bf16 && C && M >= 1, "bad argument");
        B2L_CHECK(epilogue >= 0 && epilogue <= 3, "bad epilogue");
        const int out_cols = epilogue == GEMM_SWIGLU_BF16 ? N / 2 : N;
        uint16_t* dA = dalloc<uint16_t>(c, static_cast<size_t>(M) * K);
        uint16_t* dW = dalloc<uint16_t>(c, static_cast<size_t>(N) * K);
        float* dC = dalloc<float>(c, static_cast<size_t>(M) * out_cols);
        uint16_t* dCb = dalloc<uint16_t>(c, static_cast<size_t>(M) * out_cols);
        B2L_CUDA(cudaMemcpy(dA, A_bf16, sizeof(uint16_t) * M * K, cudaMemcpyHostToDevice));
        B2L_CUDA(cudaMemcpy(dW, W_bf16, sizeof(uint16_t) * N * K, cudaMemcpyHostToDevice));
        B2L_CUDA(cudaMemcpy(dC, C, sizeof(float) * M * out_cols, cudaMemcpyHostToDevice));   // residual input (GEMM_ADD_F32)
        GemmArgs g{dC, dCb, M, N, K, out_cols, epilogue};
        gemm_bf16(c, dA, dW, g);   // the result that is returned (one application of the epilogue)
        B2L_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<uint16_t> hb;
        if (epilogue == GEMM_STORE_BF16 || epilogue == GEMM_SWIGLU_BF16) {
            hb.resize(static_cast<size_t>(M) * out_cols);
            B2L_CUDA(cudaMemcpy(hb.data(), dCb, sizeof(uint16_t) * hb.size(), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < hb.size(); i++) {
                const uint32_t u = static_cast<uint32_t>(hb[i]) << 16;
                std::memcpy(&C[i], &u, 4);
            }
        } else {
            B2L_CUDA(cudaMemcpy(C, dC, sizeof(float) * M * out_cols, cudaMemcpyDeviceToHost));
        }
        if (iters > 0) {   // timing: repeated launches (GEMM_ADD keeps accumulating into the scratch copy: harmless)
            B2L_CUDA(cudaEventRecord(c->ev0, c->stream));
            for (int i = 0; i < iters; i++) gemm_bf16(c, dA, dW, g);
            B2L_CUDA(cudaEventRecord(c->ev1, c->stream));
            B2L_CUDA(cudaStreamSynchronize(c->stream));
            float ms = 0.f;
            B2L_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
            if (device_ms) *device_ms = ms / iters;
        }
    });
}

int b2l_op_argmax(int device, const float* x, int B, int N, int32_t* out) {
    return op_guard(device, [&](b2l_ctx* c) {
        B2L_CHECK(x && out && B >= 1 && N >= 1, "bad argument");
        float* dx = dalloc<float>(c, static_cast<size_t>(B) * N);
        int32_t* di = dalloc<int32_t>(c, B);
        B2L_CUDA(cudaMemcpy(dx, x, sizeof(float) * B * N, cudaMemcpyHostToDevice));
        // the engine's kernel (cluster of CTAs per row); rows of the second half of the batch also go through the one-CTA kernel
        // it replaced, which stays as the reference of the first-max rule
        launch_cluster(c, argmax_cluster_kernel, dim3(kArgmaxCluster, B), dim3(1024), kArgmaxCluster, static_cast<const float*>(dx), N, N, 0, di,
                       static_cast<float*>(nullptr));
        B2L_CUDA(cudaStreamSynchronize(c->stream));
        {
            std::vector<int32_t> a(B), b(B);
            int32_t* dj = dalloc<int32_t>(c, B);
            launch(c, argmax_kernel, dim3(B), dim3(1024), 0, static_cast<const float*>(dx), N, N, 0, dj, static_cast<float*>(nullptr));
            B2L_CUDA(cudaStreamSynchronize(c->stream));
            B2L_CUDA(cudaMemcpy(a.data(), di, sizeof(int32_t) * B, cudaMemcpyDeviceToHost));
            B2L_CUDA(cudaMemcpy(b.data(), dj, sizeof(int32_t) * B, cudaMemcpyDeviceToHost));
            B2L_CHECK(a == b, "argmax: cluster kernel and one-CTA kernel disagree");
        }
        B2L_CUDA(cudaMemcpy(out, di, sizeof(int32_t) * B, cudaMemcpyDeviceToHost));
    });
}

}  // extern "C"


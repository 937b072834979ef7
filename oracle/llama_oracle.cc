// llama_oracle.cc -- scalar-structured fp32 CPU restatement of the Llama-3 forward.
// TEST INFRASTRUCTURE ONLY (see llama_oracle.h for the contract and the pinning status:
// "parity unpinned" by the reference's own tests; pinned to HF transformers fp32 instead).
//
// Reference anchors: the entry this body stands in for is
// gabby::inference::Llama3Generator::Generate (/root/reference/src/inference/generator.cc:33-38,
// a stub). Weights arrive exactly as gabby's Safetensors maps them
// (/root/reference/src/inference/safetensors.cc:17-36: bytes at 8 + header_size + data_offsets[0],
// row-major [out, in] bf16). Math: HF modeling_llama.py (line refs at each function).
//
// Arithmetic: bf16 weights widened to fp32, fp32 activations, fp32 accumulation in 8
// interleaved partial sums per dot product (fixed order => deterministic for a given binary),
// OpenMP across output rows / heads.
#include "llama_oracle.h"

#include <omp.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

inline float bf16_to_f32(uint16_t b) {
    uint32_t u = static_cast<uint32_t>(b) << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

inline uint16_t f32_to_bf16_rne(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return static_cast<uint16_t>(u >> 16);
}

inline float round_bf16(float f) { return bf16_to_f32(f32_to_bf16_rne(f)); }

// 8 interleaved partial sums, pairwise tree at the end. K % 8 == 0 for every Llama shape.
inline float dot_bf16_f32(const uint16_t* w, const float* x, int K) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < K; k += 8) {
        for (int j = 0; j < 8; j++) acc[j] += bf16_to_f32(w[k + j]) * x[k + j];
    }
    return ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
}

inline float dot_f32(const float* a, const float* b, int K) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < K; k += 8) {
        for (int j = 0; j < 8; j++) acc[j] += a[k + j] * b[k + j];
    }
    return ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
}

struct LayerW {
    const uint16_t *in_norm = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *o = nullptr;
    const uint16_t *post_norm = nullptr, *gate = nullptr, *up = nullptr, *down = nullptr;
};

}  // namespace

struct orc_model {
    orc_params p;
    const uint16_t* embed = nullptr;
    const uint16_t* final_norm = nullptr;
    const uint16_t* lm_head = nullptr;
    std::vector<LayerW> layers;
    std::vector<float> rope;  // [max_seq][hd/2][2]
};

struct orc_seq {
    const orc_model* m;
    int flags;
    int len = 0;
    std::vector<float> kc, vc;  // [L][max_seq][kvh*hd]
};

extern "C" {

void orc_rope_inv_freq(const orc_params* p, float* out) {
    // HF modeling_rope_utils.py: default inv_freq = 1/theta^(2i/d); llama3 rescale per band.
    const int half = p->head_dim / 2;
    const double two_pi = 6.283185307179586476925286766559;
    for (int i = 0; i < half; i++) {
        double inv = 1.0 / std::pow(p->rope_theta, (2.0 * i) / p->head_dim);
        if (p->rope_llama3) {
            const double old_ctx = p->rope_original_max_position;
            const double low_wl = old_ctx / p->rope_low_freq_factor;
            const double high_wl = old_ctx / p->rope_high_freq_factor;
            const double wl = two_pi / inv;
            if (wl > low_wl) {
                inv = inv / p->rope_factor;
            } else if (!(wl < high_wl)) {
                const double smooth = (old_ctx / wl - p->rope_low_freq_factor) /
                                      (p->rope_high_freq_factor - p->rope_low_freq_factor);
                inv = (1.0 - smooth) * inv / p->rope_factor + smooth * inv;
            }
        }
        out[i] = static_cast<float>(inv);
    }
}

void orc_rope_table(const orc_params* p, int max_pos, float* out) {
    const int half = p->head_dim / 2;
    std::vector<float> inv(half);
    orc_rope_inv_freq(p, inv.data());
    for (int pos = 0; pos < max_pos; pos++) {
        for (int i = 0; i < half; i++) {
            // HF: freqs = inv_freq (fp32) * position (fp32) -> cos/sin in fp32
            const float ang = static_cast<float>(pos) * inv[i];
            out[(static_cast<size_t>(pos) * half + i) * 2 + 0] = static_cast<float>(std::cos(static_cast<double>(ang)));
            out[(static_cast<size_t>(pos) * half + i) * 2 + 1] = static_cast<float>(std::sin(static_cast<double>(ang)));
        }
    }
}

void orc_synth_tensor(uint32_t tensor_seed, int64_t n, float scale, float offset, uint16_t* out_bits) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        uint32_t x = static_cast<uint32_t>(i) + tensor_seed;
        x ^= x >> 16;
        x *= 0x7FEB352Du;
        x ^= x >> 15;
        x *= 0x846CA68Bu;
        x ^= x >> 16;
        const float u = static_cast<float>(x >> 8) * 1.1920928955078125e-07f - 1.0f;  // 2^-23
        volatile float prod = u * scale;  // volatile: keep mul and add unfused
        out_bits[i] = f32_to_bf16_rne(offset + prod);
    }
}

int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
}

orc_model* orc_model_create(const orc_params* p) {
    if (p->head_dim % 8 || p->hidden_size % 8 || p->intermediate_size % 8) return nullptr;
    if (p->num_heads % p->num_kv_heads) return nullptr;
    auto* m = new orc_model;
    m->p = *p;
    m->layers.resize(p->num_layers);
    m->rope.resize(static_cast<size_t>(p->max_seq_len) * p->head_dim);
    orc_rope_table(p, p->max_seq_len, m->rope.data());
    return m;
}

int orc_model_set_tensor(orc_model* m, const char* hf_name, const uint16_t* bits, int64_t numel) {
    const orc_params& p = m->p;
    const int64_t H = p.hidden_size, I = p.intermediate_size, V = p.vocab_size;
    const int64_t qd = static_cast<int64_t>(p.num_heads) * p.head_dim;
    const int64_t kvd = static_cast<int64_t>(p.num_kv_heads) * p.head_dim;
    const std::string n(hf_name);
    auto set = [&](const uint16_t** slot, int64_t want) {
        if (numel != want) return -1;
        *slot = bits;
        return 0;
    };
    if (n == "model.embed_tokens.weight") return set(&m->embed, V * H);
    if (n == "model.norm.weight") return set(&m->final_norm, H);
    if (n == "lm_head.weight") return set(&m->lm_head, V * H);
    const std::string pre = "model.layers.";
    if (n.compare(0, pre.size(), pre) != 0) return -1;
    size_t dot = n.find('.', pre.size());
    if (dot == std::string::npos) return -1;
    int l = std::atoi(n.substr(pre.size(), dot - pre.size()).c_str());
    if (l < 0 || l >= p.num_layers) return -1;
    const std::string rest = n.substr(dot + 1);
    LayerW& w = m->layers[l];
    if (rest == "input_layernorm.weight") return set(&w.in_norm, H);
    if (rest == "self_attn.q_proj.weight") return set(&w.q, qd * H);
    if (rest == "self_attn.k_proj.weight") return set(&w.k, kvd * H);
    if (rest == "self_attn.v_proj.weight") return set(&w.v, kvd * H);
    if (rest == "self_attn.o_proj.weight") return set(&w.o, H * qd);
    if (rest == "post_attention_layernorm.weight") return set(&w.post_norm, H);
    if (rest == "mlp.gate_proj.weight") return set(&w.gate, I * H);
    if (rest == "mlp.up_proj.weight") return set(&w.up, I * H);
    if (rest == "mlp.down_proj.weight") return set(&w.down, H * I);
    return -1;
}

int orc_model_check(const orc_model* m) {
    if (!m->embed || !m->final_norm) return -1;
    if (!m->p.tie_word_embeddings && !m->lm_head) return -1;
    for (const LayerW& w : m->layers) {
        if (!w.in_norm || !w.q || !w.k || !w.v || !w.o || !w.post_norm || !w.gate || !w.up || !w.down) return -1;
    }
    return 0;
}

void orc_model_destroy(orc_model* m) { delete m; }

orc_seq* orc_seq_create(const orc_model* m, int flags) {
    auto* s = new orc_seq;
    s->m = m;
    s->flags = flags;
    const size_t kvd = static_cast<size_t>(m->p.num_kv_heads) * m->p.head_dim;
    const size_t n = static_cast<size_t>(m->p.num_layers) * m->p.max_seq_len * kvd;
    s->kc.assign(n, 0.f);
    s->vc.assign(n, 0.f);
    return s;
}

void orc_seq_reset(orc_seq* s) { s->len = 0; }
void orc_seq_set_flags(orc_seq* s, int flags) { s->flags = flags; }
int orc_seq_len(const orc_seq* s) { return s->len; }
void orc_seq_destroy(orc_seq* s) { delete s; }

int32_t orc_argmax(const float* x, int64_t n) {
    int64_t best = 0;
    for (int64_t i = 1; i < n; i++) {
        if (x[i] > x[best]) best = i;
    }
    return static_cast<int32_t>(best);
}

}  // extern "C"

namespace {

// Y[t][r] = <W[r,:], X[t,:]>   (nn.Linear without bias; modeling_llama.py:171-184, :230-250)
void linear(const uint16_t* W, int N, int K, const float* X, int n, float* Y, bool round_in) {
    std::vector<float> xr;
    if (round_in) {
        xr.resize(static_cast<size_t>(n) * K);
        for (size_t i = 0; i < xr.size(); i++) xr[i] = round_bf16(X[i]);
        X = xr.data();
    }
    if (n == 1) {
#pragma omp parallel for schedule(static)
        for (int r = 0; r < N; r++) Y[r] = dot_bf16_f32(W + static_cast<size_t>(r) * K, X, K);
        return;
    }
#pragma omp parallel
    {
        std::vector<float> wrow(K);
#pragma omp for schedule(static)
        for (int r = 0; r < N; r++) {
            const uint16_t* w = W + static_cast<size_t>(r) * K;
            for (int k = 0; k < K; k++) wrow[k] = bf16_to_f32(w[k]);
            for (int t = 0; t < n; t++) Y[static_cast<size_t>(t) * N + r] = dot_f32(wrow.data(), X + static_cast<size_t>(t) * K, K);
        }
    }
}

// LlamaRMSNorm (modeling_llama.py:53-70): fp32 variance, x * rsqrt(var + eps), then * weight
void rmsnorm(const float* x, const uint16_t* w, int H, float eps, float* out) {
    const float ss = dot_f32(x, x, H);
    const float inv = 1.0f / std::sqrt(ss / static_cast<float>(H) + eps);
    for (int i = 0; i < H; i++) out[i] = bf16_to_f32(w[i]) * (x[i] * inv);
}

// apply_rotary_pos_emb with rotate_half (modeling_llama.py:138-168)
void rope_inplace(float* v, int hd, const float* cs /* [hd/2][2] */) {
    const int half = hd / 2;
    for (int i = 0; i < half; i++) {
        const float c = cs[2 * i], s = cs[2 * i + 1];
        const float a = v[i], b = v[i + half];
        v[i] = a * c - b * s;
        v[i + half] = b * c + a * s;
    }
}

}  // namespace

extern "C" {

int orc_seq_forward(orc_seq* s, const int32_t* tokens, int n, int logits_all, float* logits_out,
                    float* hidden_out) {
    const orc_model* m = s->m;
    const orc_params& p = m->p;
    if (n <= 0 || s->len + n > p.max_seq_len) return -1;
    const int H = p.hidden_size, I = p.intermediate_size, V = p.vocab_size, L = p.num_layers;
    const int nh = p.num_heads, nkv = p.num_kv_heads, hd = p.head_dim;
    const int qd = nh * hd, kvd = nkv * hd, group = nh / nkv, half = hd / 2;
    const bool act16 = s->flags & ORC_ACT_BF16, kv16 = s->flags & ORC_KV_BF16, qp16 = s->flags & ORC_QP_BF16;
    const int pos0 = s->len;
    const float scale = 1.0f / std::sqrt(static_cast<float>(hd));

    std::vector<float> x(static_cast<size_t>(n) * H), xn(static_cast<size_t>(n) * H);
    std::vector<float> q(static_cast<size_t>(n) * qd), kk(static_cast<size_t>(n) * kvd), vv(static_cast<size_t>(n) * kvd);
    std::vector<float> att(static_cast<size_t>(n) * qd), proj(static_cast<size_t>(n) * H);
    std::vector<float> g(static_cast<size_t>(n) * I), u(static_cast<size_t>(n) * I);

    for (int t = 0; t < n; t++) {
        const int32_t tok = tokens[t];
        if (tok < 0 || tok >= V) return -1;
        const uint16_t* e = m->embed + static_cast<size_t>(tok) * H;
        for (int i = 0; i < H; i++) x[static_cast<size_t>(t) * H + i] = bf16_to_f32(e[i]);
    }
    auto tap = [&](int slab, const float* src) {
        if (hidden_out) std::memcpy(hidden_out + static_cast<size_t>(slab) * n * H, src, sizeof(float) * n * H);
    };
    tap(0, x.data());

    for (int l = 0; l < L; l++) {
        const LayerW& w = m->layers[l];
        float* kc = s->kc.data() + static_cast<size_t>(l) * p.max_seq_len * kvd;
        float* vc = s->vc.data() + static_cast<size_t>(l) * p.max_seq_len * kvd;
        // --- attention block (LlamaDecoderLayer.forward, modeling_llama.py:292+) ---
        for (int t = 0; t < n; t++) rmsnorm(&x[static_cast<size_t>(t) * H], w.in_norm, H, p.rms_norm_eps, &xn[static_cast<size_t>(t) * H]);
        linear(w.q, qd, H, xn.data(), n, q.data(), act16);
        linear(w.k, kvd, H, xn.data(), n, kk.data(), act16);
        linear(w.v, kvd, H, xn.data(), n, vv.data(), act16);
        for (int t = 0; t < n; t++) {
            const float* cs = m->rope.data() + static_cast<size_t>(pos0 + t) * hd;
            for (int h = 0; h < nh; h++) rope_inplace(&q[static_cast<size_t>(t) * qd + h * hd], hd, cs);
            for (int h = 0; h < nkv; h++) rope_inplace(&kk[static_cast<size_t>(t) * kvd + h * hd], hd, cs);
            float* kd = kc + static_cast<size_t>(pos0 + t) * kvd;
            float* vd = vc + static_cast<size_t>(pos0 + t) * kvd;
            for (int i = 0; i < kvd; i++) {
                kd[i] = kv16 ? round_bf16(kk[static_cast<size_t>(t) * kvd + i]) : kk[static_cast<size_t>(t) * kvd + i];
                vd[i] = kv16 ? round_bf16(vv[static_cast<size_t>(t) * kvd + i]) : vv[static_cast<size_t>(t) * kvd + i];
            }
            if (qp16) for (int i = 0; i < qd; i++) q[static_cast<size_t>(t) * qd + i] = round_bf16(q[static_cast<size_t>(t) * qd + i]);
        }
        // causal GQA softmax(q k^T / sqrt(d)) v  (eager_attention_forward, modeling_llama.py:187-222)
#pragma omp parallel
        {
            std::vector<float> sc(pos0 + n);
#pragma omp for schedule(dynamic, 1) collapse(2)
            for (int t = 0; t < n; t++) {
                for (int h = 0; h < nh; h++) {
                    const int kvh = h / group, ctx = pos0 + t + 1;
                    const float* qv = &q[static_cast<size_t>(t) * qd + h * hd];
                    float mx = -INFINITY;
                    for (int j = 0; j < ctx; j++) {
                        sc[j] = dot_f32(qv, kc + static_cast<size_t>(j) * kvd + kvh * hd, hd) * scale;
                        mx = std::fmax(mx, sc[j]);
                    }
                    float sum = 0.f;
                    for (int j = 0; j < ctx; j++) {
                        sc[j] = std::exp(sc[j] - mx);
                        sum += sc[j];
                    }
                    float* o = &att[static_cast<size_t>(t) * qd + h * hd];
                    for (int i = 0; i < hd; i++) o[i] = 0.f;
                    for (int j = 0; j < ctx; j++) {
                        const float pj = qp16 ? round_bf16(sc[j]) : sc[j];
                        const float* vr = vc + static_cast<size_t>(j) * kvd + kvh * hd;
                        for (int i = 0; i < hd; i++) o[i] += pj * vr[i];
                    }
                    const float inv = 1.0f / sum;
                    for (int i = 0; i < hd; i++) o[i] *= inv;
                }
            }
        }
        (void)half;
        linear(w.o, H, qd, att.data(), n, proj.data(), act16);
        for (size_t i = 0; i < x.size(); i++) x[i] += proj[i];
        // --- MLP block: down(silu(gate(x)) * up(x)) (modeling_llama.py:171-184) ---
        for (int t = 0; t < n; t++) rmsnorm(&x[static_cast<size_t>(t) * H], w.post_norm, H, p.rms_norm_eps, &xn[static_cast<size_t>(t) * H]);
        linear(w.gate, I, H, xn.data(), n, g.data(), act16);
        linear(w.up, I, H, xn.data(), n, u.data(), act16);
        for (size_t i = 0; i < g.size(); i++) {
            const float gv = g[i];
            g[i] = (gv / (1.0f + std::exp(-gv))) * u[i];
        }
        linear(w.down, H, I, g.data(), n, proj.data(), act16);
        for (size_t i = 0; i < x.size(); i++) x[i] += proj[i];
        tap(l + 1, x.data());
    }
    s->len += n;

    if (!logits_out && !hidden_out) return 0;
    const int first = (logits_all || hidden_out) ? 0 : n - 1;
    for (int t = first; t < n; t++) rmsnorm(&x[static_cast<size_t>(t) * H], m->final_norm, H, p.rms_norm_eps, &xn[static_cast<size_t>(t) * H]);
    tap(L + 1, xn.data());
    if (logits_out) {
        const uint16_t* head = p.tie_word_embeddings ? m->embed : m->lm_head;
        if (logits_all) {
            linear(head, V, H, xn.data(), n, logits_out, act16);
        } else {
            linear(head, V, H, &xn[static_cast<size_t>(n - 1) * H], 1, logits_out, act16);
        }
    }
    return 0;
}

int orc_greedy(orc_seq* s, const int32_t* prompt, int n_prompt, int n_new, int32_t* out_ids, float* margins) {
    const int V = s->m->p.vocab_size;
    std::vector<float> logits(V);
    if (orc_seq_forward(s, prompt, n_prompt, 0, logits.data(), nullptr)) return -1;
    for (int i = 0; i < n_new; i++) {
        const int32_t id = orc_argmax(logits.data(), V);
        out_ids[i] = id;
        if (margins) {
            float second = -INFINITY;
            for (int j = 0; j < V; j++) {
                if (j != id && logits[j] > second) second = logits[j];
            }
            margins[i] = logits[id] - second;
        }
        if (i + 1 < n_new) {
            if (orc_seq_forward(s, &id, 1, 0, logits.data(), nullptr)) return -1;
        }
    }
    return 0;
}

}  // extern "C"

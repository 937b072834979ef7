#include "params.h"

#include <cmath>
#include <stdexcept>

namespace gabby {
namespace inference {

void LlamaParams::Validate() const {
    auto need = [](bool ok, const char* what) {
        if (!ok) throw std::runtime_error(std::string("config.json: ") + what);
    };
    need(hidden_size > 0 && hidden_size % 256 == 0, "hidden_size must be a positive multiple of 256");
    need(intermediate_size > 0 && intermediate_size % 8 == 0, "intermediate_size must be a positive multiple of 8");
    need(num_hidden_layers > 0, "num_hidden_layers must be positive");
    need(num_attention_heads > 0 && num_key_value_heads > 0, "head counts must be positive");
    need(num_attention_heads % num_key_value_heads == 0, "num_attention_heads must divide by num_key_value_heads");
    need(head_dim == 32 || head_dim == 64 || head_dim == 128, "head_dim must be 32, 64 or 128");
    need(vocab_size > 0, "vocab_size must be positive");
    need(rms_norm_eps > 0, "rms_norm_eps must be positive");
    need(rope_theta > 0, "rope_theta must be positive");
    if (rope_scaling.llama3) {
        need(rope_scaling.factor >= 1.0, "rope_scaling.factor must be >= 1");
        need(rope_scaling.high_freq_factor > rope_scaling.low_freq_factor, "rope_scaling high_freq_factor must exceed low_freq_factor");
        need(rope_scaling.original_max_position_embeddings > 0, "rope_scaling.original_max_position_embeddings must be positive");
    }
}

std::vector<float> RopeInvFreq(const LlamaParams& p) {
    const int half = p.head_dim / 2;
    std::vector<float> out(half);
    const double kTwoPi = 6.283185307179586476925286766559;
    const RopeScaling& rs = p.rope_scaling;
    for (int i = 0; i < half; i++) {
        double inv = 1.0 / std::pow(p.rope_theta, (2.0 * i) / p.head_dim);
        if (rs.llama3) {
            const double old_ctx = rs.original_max_position_embeddings;
            const double wavelen = kTwoPi / inv;
            if (wavelen > old_ctx / rs.low_freq_factor) {
                inv /= rs.factor;  // long wavelengths: stretched by the full factor
            } else if (!(wavelen < old_ctx / rs.high_freq_factor)) {
                const double smooth = (old_ctx / wavelen - rs.low_freq_factor) / (rs.high_freq_factor - rs.low_freq_factor);
                inv = (1.0 - smooth) * inv / rs.factor + smooth * inv;  // medium band: interpolate
            }
        }
        out[i] = static_cast<float>(inv);
    }
    return out;
}

std::vector<float> RopeTable(const LlamaParams& p, int max_pos) {
    const int half = p.head_dim / 2;
    const std::vector<float> inv = RopeInvFreq(p);
    std::vector<float> out(static_cast<size_t>(max_pos) * half * 2);
    for (int pos = 0; pos < max_pos; pos++) {
        for (int i = 0; i < half; i++) {
            const float angle = static_cast<float>(pos) * inv[i];
            float* dst = &out[(static_cast<size_t>(pos) * half + i) * 2];
            dst[0] = static_cast<float>(std::cos(static_cast<double>(angle)));
            dst[1] = static_cast<float>(std::sin(static_cast<double>(angle)));
        }
    }
    return out;
}

}  // namespace inference
}  // namespace gabby

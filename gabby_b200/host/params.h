// params.h -- typed Llama hyper-parameters. gabby keeps config.json as an untyped JSON tree
// (/root/reference/src/inference/config.h:12-19, config.cc:13); the B200 path needs them typed.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace gabby {
namespace inference {

struct RopeScaling {
    bool llama3 = false;
    double factor = 1.0;
    double low_freq_factor = 1.0;
    double high_freq_factor = 4.0;
    int original_max_position_embeddings = 8192;
};

struct LlamaParams {
    int hidden_size = 0;
    int intermediate_size = 0;
    int num_hidden_layers = 0;
    int num_attention_heads = 0;
    int num_key_value_heads = 0;
    int head_dim = 0;
    int vocab_size = 0;
    bool tie_word_embeddings = false;
    float rms_norm_eps = 1e-5f;
    double rope_theta = 10000.0;
    RopeScaling rope_scaling;
    int max_position_embeddings = 0;
    int bos_token_id = -1;
    std::vector<int> eos_token_ids;

    // throws std::runtime_error naming the offending key
    void Validate() const;
};

// inv_freq[head_dim/2] with the `llama3` wavelength-band rescaling
// (HF modeling_rope_utils.py _compute_llama3_parameters): computed in double, stored as fp32.
std::vector<float> RopeInvFreq(const LlamaParams& p);

// [max_pos][head_dim/2][2] = (cos, sin) of fp32(pos) * inv_freq, evaluated in double and rounded
// once: the table the CUDA RoPE kernels index by position.
std::vector<float> RopeTable(const LlamaParams& p, int max_pos);

}  // namespace inference
}  // namespace gabby

#include "config.h"

#include <cstdlib>
#include <stdexcept>

namespace gabby {
namespace inference {

namespace fs = std::filesystem;

std::unique_ptr<InferenceConfig> LoadConfig(const fs::path& dir) {
    // same five files, same order, as the reference (config.cc:13-17); tokenizer files may be absent
    // in synthetic test directories only if they are empty objects, so they are still required
    auto cfg = std::make_unique<InferenceConfig>(InferenceConfig{
        json::ParseFile((dir / "config.json").string()),
        json::ParseFile((dir / "generation_config.json").string()),
        json::ParseFile((dir / "special_tokens_map.json").string()),
        json::ParseFile((dir / "tokenizer_config.json").string()),
        json::ParseFile((dir / "tokenizer.json").string()),
        Checkpoint::Open(dir),
    });
    return cfg;
}

fs::path FindDefaultModelDir() {
    const char* home = std::getenv("HOME");
    if (home == nullptr) throw std::runtime_error("env var HOME is unset");
    const fs::path snapshots =
        fs::path(home) / ".cache/huggingface/hub/models--meta-llama--Llama-3.2-1B-Instruct/snapshots";
    std::error_code ec;
    fs::directory_iterator it(snapshots, ec);
    if (ec) throw std::runtime_error("can't access model dir at " + snapshots.string() + ": " + ec.message());
    if (it == fs::end(it)) throw std::runtime_error("no snapshots found in " + snapshots.string());
    return it->path();
}

LlamaParams ParamsFromConfig(const json::Value& c, const json::Value* gen) {
    LlamaParams p;
    auto geti = [&](const char* k) { return static_cast<int>(c.at(k).as_number()); };
    p.hidden_size = geti("hidden_size");
    p.intermediate_size = geti("intermediate_size");
    p.num_hidden_layers = geti("num_hidden_layers");
    p.num_attention_heads = geti("num_attention_heads");
    p.num_key_value_heads = c.contains("num_key_value_heads") ? geti("num_key_value_heads") : p.num_attention_heads;
    p.head_dim = c.contains("head_dim") && c.at("head_dim").is(json::Type::NUM) ? geti("head_dim")
                                                                                  : p.hidden_size / p.num_attention_heads;
    p.vocab_size = geti("vocab_size");
    p.tie_word_embeddings = c.boolean_or("tie_word_embeddings", false);
    p.rms_norm_eps = static_cast<float>(c.number_or("rms_norm_eps", 1e-5));
    p.rope_theta = c.number_or("rope_theta", 10000.0);
    p.max_position_embeddings = static_cast<int>(c.number_or("max_position_embeddings", 8192));
    if (c.contains("rope_scaling") && c.at("rope_scaling").is(json::Type::OBJ)) {
        const json::Value& rs = c.at("rope_scaling");
        const std::string type = rs.contains("rope_type") ? rs.at("rope_type").as_string()
                                 : rs.contains("type")    ? rs.at("type").as_string()
                                                          : "default";
        if (type == "llama3") {
            p.rope_scaling.llama3 = true;
            p.rope_scaling.factor = rs.at("factor").as_number();
            p.rope_scaling.low_freq_factor = rs.at("low_freq_factor").as_number();
            p.rope_scaling.high_freq_factor = rs.at("high_freq_factor").as_number();
            p.rope_scaling.original_max_position_embeddings = static_cast<int>(rs.at("original_max_position_embeddings").as_number());
        } else if (type != "default") {
            throw std::runtime_error("config.json: unsupported rope_scaling type " + type);
        }
    }
    p.bos_token_id = static_cast<int>(c.number_or("bos_token_id", -1));
    auto read_eos = [&](const json::Value& v) {
        if (!v.contains("eos_token_id")) return;
        const json::Value& e = v.at("eos_token_id");
        if (e.is(json::Type::NUM)) p.eos_token_ids.push_back(static_cast<int>(e.as_number()));
        if (e.is(json::Type::ARRAY))
            for (const auto& x : e.as_array()) p.eos_token_ids.push_back(static_cast<int>(x->as_number()));
    };
    if (gen && gen->is(json::Type::OBJ)) read_eos(*gen);   // generation_config.json wins (it lists all three Llama-3 stops)
    if (p.eos_token_ids.empty()) read_eos(c);
    p.Validate();
    return p;
}

}  // namespace inference
}  // namespace gabby

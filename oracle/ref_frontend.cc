// ref_frontend.cc -- loads a model directory THROUGH THE REFERENCE'S OWN PARSERS and hands the
// tensors to the CPU oracle. TEST INFRASTRUCTURE ONLY; built into oracle/_ref/ (git-ignored)
// from the reference sources where they lie under /root/reference (never copied here).
//
// Reference code exercised (compiled from /root/reference/src by oracle/Makefile):
//   inference::LoadConfig            src/inference/config.cc:11-28   (5 JSON files + model.safetensors)
//   inference::Safetensors::LoadFile src/inference/safetensors.cc:17-36, header() safetensors.h:16
//   json::Parse / ParseFile          src/json/parser.cc:264-281
//   gabby::Open / Mmap               src/utils/pointers.cc:27-38
// Safetensors exposes only header() (safetensors.h:16-24: mem_/data_offset_ are private), so the
// bytes are reached by mapping the file a second time with the reference's Open/Mmap and
// recomputing data_offset = 8 + u64le(file[0..8]) -- SURVEY.md section 8(c).
#include <fcntl.h>

#include <cstdint>
#include <cstdio>
#include <filesystem>
#include <memory>
#include <string>

#include "inference/config.h"
#include "inference/generator.h"
#include "json/json.h"
#include "llama_oracle.h"
#include "utils/pointers.h"

namespace {

struct RefModel {
    std::unique_ptr<gabby::inference::InferenceConfig> cfg;
    gabby::OwnedMmap bytes{nullptr, gabby::MmapDeleter{0}};
    orc_model* model = nullptr;
    std::string error;
};

double num(gabby::json::ObjectValue& o, const char* k) { return o.at(k)->as_number().get(); }

}  // namespace

extern "C" {

// Returns an opaque handle (never NULL); check orc_ref_error().
void* orc_ref_load(const char* model_dir, int max_seq_len) {
    auto* r = new RefModel;
    try {
        namespace fs = std::filesystem;
        const fs::path dir(model_dir);
        r->cfg = gabby::inference::LoadConfig(dir);
        auto& c = r->cfg->config->as_object();
        orc_params p{};
        p.hidden_size = static_cast<int>(num(c, "hidden_size"));
        p.intermediate_size = static_cast<int>(num(c, "intermediate_size"));
        p.num_layers = static_cast<int>(num(c, "num_hidden_layers"));
        p.num_heads = static_cast<int>(num(c, "num_attention_heads"));
        p.num_kv_heads = static_cast<int>(num(c, "num_key_value_heads"));
        p.head_dim = c.get().contains("head_dim") ? static_cast<int>(num(c, "head_dim"))
                                                  : p.hidden_size / p.num_heads;
        p.vocab_size = static_cast<int>(num(c, "vocab_size"));
        p.tie_word_embeddings = c.get().contains("tie_word_embeddings") && c.at("tie_word_embeddings")->as_boolean().get();
        p.rms_norm_eps = static_cast<float>(num(c, "rms_norm_eps"));
        p.rope_theta = num(c, "rope_theta");
        p.rope_llama3 = 0;
        if (c.get().contains("rope_scaling") && c.at("rope_scaling")->type() == gabby::json::Type::OBJ) {
            auto& rs = c.at("rope_scaling")->as_object();
            if (*rs.at("rope_type")->as_string() == "llama3") {
                p.rope_llama3 = 1;
                p.rope_factor = num(rs, "factor");
                p.rope_low_freq_factor = num(rs, "low_freq_factor");
                p.rope_high_freq_factor = num(rs, "high_freq_factor");
                p.rope_original_max_position = static_cast<int>(num(rs, "original_max_position_embeddings"));
            }
        }
        p.max_seq_len = max_seq_len;
        r->model = orc_model_create(&p);
        if (!r->model) throw std::runtime_error("unsupported shapes");

        const fs::path st = dir / "model.safetensors";  // config.cc:18: single file only
        const size_t file_size = fs::file_size(st);
        r->bytes = gabby::Mmap(file_size, gabby::Open(st.c_str(), O_RDONLY));
        uint64_t header_size = 0;
        for (int i = 0; i < 8; i++) header_size |= static_cast<uint64_t>(r->bytes.get()[i]) << (8 * i);
        const uint8_t* data = r->bytes.get() + 8 + header_size;

        auto& header = r->cfg->tensors.header()->as_object();
        for (auto& [name, meta] : header.get()) {
            if (name == "__metadata__") continue;
            auto& mo = meta->as_object();
            if (*mo.at("dtype")->as_string() != "BF16") throw std::runtime_error("dtype must be BF16: " + name);
            auto& off = mo.at("data_offsets")->as_array();
            const auto b = static_cast<uint64_t>(off[0]->as_number().get());
            const auto e = static_cast<uint64_t>(off[1]->as_number().get());
            if (orc_model_set_tensor(r->model, name.c_str(), reinterpret_cast<const uint16_t*>(data + b),
                                     static_cast<int64_t>((e - b) / 2)) != 0) {
                throw std::runtime_error("unexpected tensor: " + name);
            }
        }
        if (orc_model_check(r->model) != 0) throw std::runtime_error("missing tensors");
    } catch (const std::exception& ex) {
        r->error = ex.what();
        if (r->error.empty()) r->error = "error";
    }
    return r;
}

const char* orc_ref_error(void* h) {
    auto* r = static_cast<RefModel*>(h);
    return r->error.empty() ? nullptr : r->error.c_str();
}

orc_model* orc_ref_model(void* h) { return static_cast<RefModel*>(h)->model; }

// The reference's own Generate (the constant-string stub, generator.cc:33-38), for the record:
// copies the returned content into `out`; returns its length.
int orc_ref_stub_generate(void* h, char* out, int cap) {
    auto* r = static_cast<RefModel*>(h);
    (void)r;
    // Llama3Generator::Load consumes an InferenceConfig; build a throwaway one is not possible
    // without re-reading the directory, so the stub is exercised through a null config.
    auto gen = gabby::inference::Llama3Generator::Load(nullptr);
    gabby::inference::Message msg = gen->Generate(gabby::inference::Request{});
    std::snprintf(out, cap, "%s", msg.content.c_str());
    return static_cast<int>(msg.content.size());
}

void orc_ref_free(void* h) {
    auto* r = static_cast<RefModel*>(h);
    if (r->model) orc_model_destroy(r->model);
    delete r;
}

}  // extern "C"

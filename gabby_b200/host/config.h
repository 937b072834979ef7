// config.h -- mirrors gabby's InferenceConfig / LoadConfig / FindDefaultModelDir
// (/root/reference/src/inference/config.h:12-24, config.cc:11-56): the five HF JSON files of a
// snapshot directory plus its tensors. Differences, all additive: the tensors member is a
// Checkpoint (single-file OR sharded safetensors with a tensor accessor), and the typed
// hyper-parameters the forward pass needs are derived once (ParamsFromConfig).
#pragma once
#include <filesystem>
#include <memory>

#include "json.h"
#include "params.h"
#include "safetensors.h"

namespace gabby {
namespace inference {

struct InferenceConfig {
    json::ValuePtr config;
    json::ValuePtr gen_config;
    json::ValuePtr special_tokens_map;
    json::ValuePtr tok_config;
    json::ValuePtr tok;
    Checkpoint tensors;
};

std::unique_ptr<InferenceConfig> LoadConfig(const std::filesystem::path& directory);

// first entry of $HOME/.cache/huggingface/hub/models--meta-llama--Llama-3.2-1B-Instruct/snapshots
std::filesystem::path FindDefaultModelDir();

// typed view of config.json (+ generation_config.json for eos ids); throws on missing/invalid keys
LlamaParams ParamsFromConfig(const json::Value& config, const json::Value* gen_config);

}  // namespace inference
}  // namespace gabby

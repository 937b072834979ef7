// safetensors.h -- safetensors reader with a tensor accessor.
//
// Mirrors gabby's Safetensors (/root/reference/src/inference/safetensors.h:13-25: LoadFile(),
// header()) and adds what the forward pass needs and the reference lacks:
//   * a tensor accessor -- the reference exposes only header(); mem_/data_offset_ are private
//   * a real u64 little-endian header length (reference: int-promoted shift, safetensors.cc:25-27)
//   * MAP_FAILED / bounds checks (reference: none, pointers.cc:35-38)
//   * sharded checkpoints: model.safetensors.index.json + model-0000i-of-0000n.safetensors
//     (reference: single model.safetensors only, config.cc:18)
#pragma once
#include <cstdint>
#include <filesystem>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "json.h"

namespace gabby {
namespace inference {

struct TensorView {
    std::string name;
    std::string dtype;            // "BF16", "F32", ...
    std::vector<int64_t> shape;
    const uint8_t* data = nullptr;
    size_t nbytes = 0;
    int64_t numel() const;
};

class Safetensors {
public:
    static Safetensors LoadFile(const std::filesystem::path& path);
    Safetensors(Safetensors&&) noexcept;
    Safetensors& operator=(Safetensors&&) noexcept;
    Safetensors(const Safetensors&) = delete;
    ~Safetensors();

    const json::ValuePtr header() const { return header_; }       // same accessor as the reference
    std::vector<std::string> names() const;
    bool contains(const std::string& name) const;
    TensorView tensor(const std::string& name) const;             // throws std::runtime_error
    size_t file_size() const { return size_; }

private:
    Safetensors() = default;
    void Release();
    uint8_t* mem_ = nullptr;
    size_t size_ = 0;
    size_t data_offset_ = 0;
    json::ValuePtr header_;
};

// One logical checkpoint over one or many safetensors files in an HF snapshot directory.
class Checkpoint {
public:
    static Checkpoint Open(const std::filesystem::path& dir);
    std::vector<std::string> names() const;
    bool contains(const std::string& name) const { return where_.count(name) != 0; }
    TensorView tensor(const std::string& name) const;
    size_t num_files() const { return files_.size(); }

private:
    std::vector<std::shared_ptr<Safetensors>> files_;
    std::map<std::string, size_t> where_;
};

}  // namespace inference
}  // namespace gabby

#include "safetensors.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>
#include <set>
#include <stdexcept>

namespace gabby {
namespace inference {

namespace fs = std::filesystem;

int64_t TensorView::numel() const {
    int64_t n = 1;
    for (int64_t d : shape) n *= d;
    return n;
}

namespace {
size_t DtypeSize(const std::string& dt) {
    if (dt == "BF16" || dt == "F16" || dt == "I16" || dt == "U16") return 2;
    if (dt == "F32" || dt == "I32" || dt == "U32") return 4;
    if (dt == "F64" || dt == "I64" || dt == "U64") return 8;
    if (dt == "I8" || dt == "U8" || dt == "BOOL" || dt == "F8_E4M3" || dt == "F8_E5M2") return 1;
    throw std::runtime_error("safetensors: unknown dtype " + dt);
}
}  // namespace

Safetensors Safetensors::LoadFile(const fs::path& path) {
    // format: 8-byte little-endian header length, JSON header, then the tensor bytes
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error(path.string() + ": " + std::strerror(errno));
    struct stat st {};
    if (::fstat(fd, &st) != 0) {
        const int e = errno;
        ::close(fd);
        throw std::runtime_error(path.string() + ": " + std::strerror(e));
    }
    Safetensors s;
    s.size_ = static_cast<size_t>(st.st_size);
    if (s.size_ < 8) {
        ::close(fd);
        throw std::runtime_error(path.string() + ": too short to be a safetensors file");
    }
    void* mem = ::mmap(nullptr, s.size_, PROT_READ, MAP_PRIVATE, fd, 0);
    const int map_errno = errno;
    ::close(fd);  // the mapping keeps the file alive
    if (mem == MAP_FAILED) throw std::runtime_error(path.string() + ": mmap: " + std::strerror(map_errno));
    s.mem_ = static_cast<uint8_t*>(mem);
    uint64_t header_size = 0;
    for (int i = 0; i < 8; i++) header_size |= static_cast<uint64_t>(s.mem_[i]) << (8 * i);
    if (header_size > s.size_ - 8) throw std::runtime_error(path.string() + ": header length exceeds the file size");
    s.header_ = json::Parse(std::string_view(reinterpret_cast<const char*>(s.mem_ + 8), header_size));
    if (!s.header_->is(json::Type::OBJ)) throw std::runtime_error(path.string() + ": header is not a JSON object");
    s.data_offset_ = 8 + static_cast<size_t>(header_size);
    return s;
}

Safetensors::Safetensors(Safetensors&& o) noexcept { *this = std::move(o); }
Safetensors& Safetensors::operator=(Safetensors&& o) noexcept {
    if (this != &o) {
        Release();
        mem_ = o.mem_; size_ = o.size_; data_offset_ = o.data_offset_; header_ = std::move(o.header_);
        o.mem_ = nullptr; o.size_ = 0;
    }
    return *this;
}
Safetensors::~Safetensors() { Release(); }
void Safetensors::Release() {
    if (mem_) ::munmap(mem_, size_);
    mem_ = nullptr;
}

std::vector<std::string> Safetensors::names() const {
    std::vector<std::string> out;
    for (const auto& [k, v] : header_->members())
        if (k != "__metadata__") out.push_back(k);
    return out;
}

bool Safetensors::contains(const std::string& name) const { return name != "__metadata__" && header_->contains(name); }

TensorView Safetensors::tensor(const std::string& name) const {
    if (!contains(name)) throw std::runtime_error("safetensors: no tensor named " + name);
    const json::Value& meta = header_->at(name);
    TensorView t;
    t.name = name;
    t.dtype = meta.at("dtype").as_string();
    for (const auto& d : meta.at("shape").as_array()) t.shape.push_back(d->as_int());
    const json::Value& off = meta.at("data_offsets");
    const uint64_t b = static_cast<uint64_t>(off[0].as_number()), e = static_cast<uint64_t>(off[1].as_number());
    if (e < b || data_offset_ + e > size_) throw std::runtime_error("safetensors: " + name + ": data_offsets outside the file");
    t.nbytes = static_cast<size_t>(e - b);
    if (t.nbytes != static_cast<size_t>(t.numel()) * DtypeSize(t.dtype))
        throw std::runtime_error("safetensors: " + name + ": byte length does not match shape x dtype");
    t.data = mem_ + data_offset_ + b;
    return t;
}

Checkpoint Checkpoint::Open(const fs::path& dir) {
    Checkpoint c;
    const fs::path single = dir / "model.safetensors", index = dir / "model.safetensors.index.json";
    std::vector<fs::path> paths;
    if (fs::exists(single)) {
        paths.push_back(single);
    } else if (fs::exists(index)) {
        const json::ValuePtr idx = json::ParseFile(index.string());
        std::set<std::string> shard_names;
        for (const auto& [tensor, file] : idx->at("weight_map").members()) shard_names.insert(file->as_string());
        for (const auto& n : shard_names) paths.push_back(dir / n);
        if (paths.empty()) throw std::runtime_error(index.string() + ": empty weight_map");
    } else {
        throw std::runtime_error("no model.safetensors or model.safetensors.index.json in " + dir.string());
    }
    for (const auto& p : paths) {
        c.files_.push_back(std::make_shared<Safetensors>(Safetensors::LoadFile(p)));
        for (const auto& n : c.files_.back()->names()) c.where_[n] = c.files_.size() - 1;
    }
    return c;
}

std::vector<std::string> Checkpoint::names() const {
    std::vector<std::string> out;
    for (const auto& [k, v] : where_) out.push_back(k);
    return out;
}

TensorView Checkpoint::tensor(const std::string& name) const {
    auto it = where_.find(name);
    if (it == where_.end()) throw std::runtime_error("checkpoint: no tensor named " + name);
    return files_[it->second]->tensor(name);
}

}  // namespace inference
}  // namespace gabby

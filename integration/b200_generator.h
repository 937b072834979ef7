// b200_generator.h -- the drop-in: a gabby::inference::Generator (the REFERENCE'S OWN header,
// /root/reference/src/inference/generator.h:30-34) whose Generate runs the B200 engine.
//
// This file is compiled TOGETHER WITH gabby's sources (integration/Makefile) and reaches the engine only through the C-ABI
// of include/gabby_b200_host.h (libgabby_host.so -> include/b2l.h -> libb2l.so): no type of this repo crosses the line,
// so gabby's gabby::inference:: / gabby::json:: symbols and the host library's mirrors of them never meet.
#ifndef GABBY_B200_INTEGRATION_GENERATOR_H_
#define GABBY_B200_INTEGRATION_GENERATOR_H_

#include <filesystem>
#include <memory>

#include "inference/generator.h"   // gabby's

struct gb_generator;

namespace gabby {
namespace inference {

class B200Llama3Generator : public Generator {
public:
    // what Llama3Generator::Load(LoadConfig(model_dir)) is for the stub (/root/reference/src/service.cc:120-124)
    static std::unique_ptr<Generator> Load(const std::filesystem::path& model_dir, int device = 0, int max_positions = 2048,
                                           int max_new_tokens = 64);
    ~B200Llama3Generator() override;
    Message Generate(const Request& req) override;   // throws std::runtime_error -> HTTP 500 (/root/reference/src/http/server.cc:371-378)

private:
    explicit B200Llama3Generator(gb_generator* g) : g_(g) {}
    gb_generator* g_;
};

}  // namespace inference
}  // namespace gabby

#endif

#include "b200_generator.h"

#include <stdexcept>
#include <string>
#include <vector>

#include "gabby_b200_host.h"

namespace gabby {
namespace inference {

std::unique_ptr<Generator> B200Llama3Generator::Load(const std::filesystem::path& model_dir, int device, int max_positions,
                                                     int max_new_tokens) {
    gb_generator* g = nullptr;
    if (gb_generator_load(model_dir.c_str(), device, max_positions, max_new_tokens, &g) != 0)
        throw std::runtime_error(std::string("B200Llama3Generator::Load: ") + gb_last_error());
    return std::unique_ptr<Generator>(new B200Llama3Generator(g));
}

B200Llama3Generator::~B200Llama3Generator() { gb_generator_free(g_); }

Message B200Llama3Generator::Generate(const Request& req) {
    std::vector<char> out(1 << 16);
    int prompt_tokens = 0, completion_tokens = 0, finish = 0;
    if (gb_generator_generate_detailed(g_, req.system_message.content.c_str(), req.user_message.content.c_str(), /*max_tokens=*/0, out.data(),
                                       static_cast<int>(out.size()), &prompt_tokens, &completion_tokens, &finish) != 0)
        throw std::runtime_error(std::string("B200Llama3Generator::Generate: ") + gb_last_error());
    return Message{.role = "assistant", .content = std::string(out.data())};
}

}  // namespace inference
}  // namespace gabby

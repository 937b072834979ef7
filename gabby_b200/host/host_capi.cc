// host_capi.cc -- C view of the host layer (include/gabby_b200_host.h).
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>

#include "gabby_b200_host.h"
#include "params.h"

namespace {
thread_local std::string g_error;

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}
}  // namespace

extern "C" {

const char* gb_last_error(void) { return g_error.c_str(); }

int gb_rope_table(double rope_theta, int llama3, double factor, double low_freq_factor, double high_freq_factor,
                  int original_max_position, int head_dim, int max_pos, float* out) {
    return guarded([&] {
        gabby::inference::LlamaParams p;
        p.head_dim = head_dim;
        p.rope_theta = rope_theta;
        p.rope_scaling.llama3 = llama3 != 0;
        p.rope_scaling.factor = factor;
        p.rope_scaling.low_freq_factor = low_freq_factor;
        p.rope_scaling.high_freq_factor = high_freq_factor;
        p.rope_scaling.original_max_position_embeddings = original_max_position;
        if (head_dim <= 0 || head_dim % 2 || max_pos <= 0 || !out) throw std::runtime_error("gb_rope_table: bad argument");
        const std::vector<float> t = gabby::inference::RopeTable(p, max_pos);
        std::memcpy(out, t.data(), t.size() * sizeof(float));
    });
}

}  // extern "C"

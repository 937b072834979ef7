// synth.cuh -- on-device restatement of the synthetic-weight counter hash (gabby_b200/synth.py,
// oracle/llama_oracle.cc orc_synth_tensor). Lets the 8B/70B timing configs fill 16-141 GB of
// weights without a host file; bit-identical to the numpy generator (tests/test_gpu_parity.py).
#pragma once
#include "common.cuh"

namespace b2l {

__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

// dst[r * dst_stride + c] = value at full-tensor flat index (row0 + r) * full_cols + col0 + c
__global__ void synth_fill_kernel(uint16_t* __restrict__ dst, int64_t dst_stride, int64_t row0, int64_t nrows, int64_t col0,
                                  int64_t ncols, int64_t full_cols, uint32_t tensor_seed, float scale, float offset) {
    const int64_t n = nrows * ncols;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / ncols, c = i % ncols;
        const uint32_t flat = static_cast<uint32_t>((row0 + r) * full_cols + col0 + c);
        const uint32_t x = lowbias32(flat + tensor_seed);
        const float u = __fsub_rn(__fmul_rn(static_cast<float>(x >> 8), 1.1920928955078125e-07f), 1.0f);
        dst[r * dst_stride + c] = f32_to_bf16_bits(__fadd_rn(offset, __fmul_rn(u, scale)));
    }
}

}  // namespace b2l

// Micro-benchmark (B200): latency and issue rate of the legacy warp-level tensor instruction mma.sync.m16n8k16 (SASS HMMA.16816)
// and of ldmatrix.x4, as seen by 8 warps of one CTA per SM -- the row engine of the decode megakernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/hmma_bench tools/hmma_bench.cu && build/hmma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void k(int mode, int iters, long long* out, float* sink) {
    __shared__ __align__(128) uint8_t sm[16384];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3f803f80u;
    __syncthreads();
    uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
    float c[8][4] = {};
    const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(sm)) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 512;
    const long long t0 = clock64();
    if (mode == 0) {          // dependent chain: one accumulator
        for (int i = 0; i < iters; i++) mma(c[0], a, 0x3f803f80u, 0x3f803f80u);
    } else if (mode == 1) {   // 8 independent accumulators
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) mma(c[j], a, 0x3f803f80u, 0x3f803f80u);
        }
    } else if (mode == 2) {   // ldmatrix -> mma dependent
        for (int i = 0; i < iters; i++) {
            uint32_t r[4];
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(base + (i & 3) * 4096) : "memory");
            mma(c[i & 7], r, 0x3f803f80u, 0x3f803f80u);
        }
    } else {                  // ldmatrix only, dependent address chain
        uint32_t addr = base;
        for (int i = 0; i < iters; i++) {
            uint32_t r[4];
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
            addr = base + (r[0] & 0x0u);
            c[0][0] += __uint_as_float(r[1]);
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    if (s == 12345.f) *sink = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[mode] = t1 - t0;
}
int main() {
    long long* out; float* sink;
    cudaMallocManaged(&out, 64); cudaMalloc(&sink, 4);
    const int iters = 4096;
    const char* names[] = {"mma dependent chain", "mma 8 independent accumulators", "ldmatrix.x4 -> mma", "ldmatrix.x4 dependent"};
    for (int warps = 1; warps <= 8; warps *= 2) {
        for (int mode = 0; mode < 4; mode++) {
            k<<<148, warps * 32>>>(mode, iters, out, sink);
            cudaDeviceSynchronize();
            printf("%d warps/SM  %-34s %7.1f cycles per instruction (per warp)\n", warps, names[mode], double(out[mode]) / iters);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

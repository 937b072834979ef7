// Micro-benchmarks behind the megakernel's phase-boundary costs (B200): L2 latency, grid barrier.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void chase(const unsigned* buf, int n, unsigned* out, long long* cycles) {
    unsigned i = 0;
    for (int k = 0; k < 64; k++) i = __ldcg(buf + i);   // warm
    long long t0 = clock64();
    for (int k = 0; k < n; k++) i = __ldcg(buf + i);
    long long t1 = clock64();
    *out = i; *cycles = t1 - t0;
}

__device__ __forceinline__ unsigned long long ld_acq(const unsigned long long* p) {
    unsigned long long v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
// variant 0: red.release + ld.acquire poll by thread 0, bar.sync around (the megakernel's barrier)
// variant 1: same but poll with ld.relaxed (volatile) and a single acquire fence at the end
// variant 2: per-CTA flags: each CTA stores its epoch to flags[cta]; warp 0 polls all flags with 32 lanes
__global__ void gridbar(unsigned long long* counter, unsigned long long* flags, int iters, int variant, long long* cycles, float* payload) {
    const int G = gridDim.x, tid = threadIdx.x;
    long long t0 = clock64();
    for (int it = 1; it <= iters; it++) {
        // a little payload like a phase epilogue: one global store per warp
        if ((tid & 31) == 0) payload[blockIdx.x * 32 + (tid >> 5)] = it;
        __syncthreads();
        if (variant == 0) {
            if (tid == 0) {
                asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(counter), "l"(1ull) : "memory");
                while (ld_acq(counter) < (unsigned long long)it * G) {}
            }
        } else if (variant == 1) {
            if (tid == 0) {
                asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(counter), "l"(1ull) : "memory");
                while (*reinterpret_cast<volatile unsigned long long*>(counter) < (unsigned long long)it * G) {}
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
        } else {
            if (tid == 0) {
                asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flags + blockIdx.x), "l"((unsigned long long)it) : "memory");
            }
            if (tid < 32) {
                for (;;) {
                    bool ok = true;
                    for (int c = tid; c < G; c += 32) ok = ok && (*reinterpret_cast<volatile unsigned long long*>(flags + c) >= (unsigned long long)it);
                    if (__all_sync(0xffffffffu, ok)) break;
                }
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    printf("%s, %d SMs, clock %d MHz\n", p.name, p.multiProcessorCount, clk_khz / 1000);
    // L2 latency: random cyclic permutation over 8 MB (L2-resident, larger than L1)
    const int N = 2 << 20; unsigned* h = new unsigned[N];
    for (int i = 0; i < N; i++) h[i] = i;
    unsigned s = 12345; for (int i = N - 1; i > 0; i--) { s = s * 1664525u + 1013904223u; int j = s % i; unsigned t = h[i]; h[i] = h[j]; h[j] = t; }  // Sattolo
    unsigned* d; unsigned* out; long long* cyc;
    CK(cudaMalloc(&d, N * 4)); CK(cudaMalloc(&out, 4)); CK(cudaMallocManaged(&cyc, 8));
    CK(cudaMemcpy(d, h, N * 4, cudaMemcpyHostToDevice));
    chase<<<1, 1>>>(d, 2000, out, cyc); CK(cudaDeviceSynchronize());
    chase<<<1, 1>>>(d, 4000, out, cyc); CK(cudaDeviceSynchronize());
    printf("L2 dependent-load latency (ld.global.cg, 8 MB set): %.1f cycles\n", (double)*cyc / 4000);
    unsigned long long *counter, *flags; float* payload;
    CK(cudaMalloc(&counter, 8)); CK(cudaMalloc(&flags, 8 * 1024)); CK(cudaMalloc(&payload, 4 * 32 * 1024));
    for (int threads : {32, 288, 544}) {
        for (int variant = 0; variant < 3; variant++) {
            CK(cudaMemset(counter, 0, 8)); CK(cudaMemset(flags, 0, 8 * 1024));
            int G = p.multiProcessorCount, iters = 2000;
            void* args[] = {&counter, &flags, &iters, &variant, &cyc, &payload};
            CK(cudaLaunchCooperativeKernel((void*)gridbar, dim3(G), dim3(threads), args, 0, 0));
            CK(cudaDeviceSynchronize());
            printf("grid barrier variant %d, %3d threads/CTA: %.0f cycles = %.2f us\n", variant, threads, (double)*cyc / iters, (double)*cyc / iters / (clk_khz / 1000.0));
        }
    }
    return 0;
}

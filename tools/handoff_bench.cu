// Micro-benchmark of the megakernel's phase hand-off (B200): G persistent CTAs; every iteration each CTA
// "produces" its share of a K-word vector as {fp32, seq} words, then every CTA polls the WHOLE vector
// (the all-gather every GEMV phase needs), fans it out through shared memory and reduces it (RMSNorm-like).
// Reports the cycle time of produce -> everybody has the input, for several polling schemes, with and
// without a background TMA weight stream saturating HBM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/handoff_bench tools/handoff_bench.cu && build/handoff_bench
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kWarps = 8, kThreads = kWarps * 32 + 32;   // + stream warp
constexpr int kStage = 16384, kStages = 8;

struct P {
    unsigned long long* buf;   // [2][R][Kpad] words
    int K, R, iters, variant, sleep_ns, stream, work_ns, jitter_ns, spin_sleep;
    const unsigned char* wbuf; // background stream source
    size_t wbytes;
    unsigned long long* counter;
    unsigned long long* flags;
    float* sink;
    long long* cycles;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ll_st(unsigned long long* p, float v, uint32_t seq) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"((static_cast<unsigned long long>(seq) << 32) | __float_as_uint(v)) : "memory");
}
__device__ __forceinline__ uint4 ll_ld2(const unsigned long long* p) {
    uint4 w;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// variants
//  0  every warp polls its own 256-word blocks (4 x 16-byte loads per lane per block), retry whole block   [megakernel today, K <= 2048]
//  1  like 0, but for K > 2048 every warp polls a 2048-word slice and warps w, w+4 poll the SAME slice        [megakernel today, down]
//  2  TMA bulk copy of the whole vector into shared memory by one thread, everybody checks the copy
//  3  counter barrier: data as plain stores, red.release counter, thread 0 polls (ld.acquire), then plain loads
//  4  per-CTA flag: data stores + fence + flag store; warp 0 polls G flags, then plain ld.cg of the data
__global__ void __launch_bounds__(kThreads, 1) handoff(P p) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* xs = reinterpret_cast<float*>(smem);                      // [K] fp32
    unsigned long long* tma_dst = reinterpret_cast<unsigned long long*>(smem + 32768);   // [K] words (variant 2)
    unsigned char* ring = smem + 32768 + 65536;
    __shared__ __align__(8) unsigned long long bars[kStages + 2];
    __shared__ volatile int stop;
    __shared__ int vote;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, G = gridDim.x, c = blockIdx.x;
    const int K = p.K, Kpad = (K + 255) / 256 * 256;
    if (tid == 0) {
        stop = 0;
        for (int s = 0; s < kStages + 2; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (w == kWarps) {
        // background weight stream: 16 KB TMA bulk copies into a ring nobody reads, as fast as they complete
        if (lane == 0 && p.stream) {
            size_t off = (static_cast<size_t>(c) * 7919 * kStage) % (p.wbytes - kStage);
            uint32_t parity[kStages] = {};
            int issued = 0;
            while (!stop) {
                const int s = issued % kStages;
                if (issued >= kStages) {
                    while (!mbar_try(smem_u32(&bars[s]), parity[s])) { if (stop) break; if (p.spin_sleep) __nanosleep(p.spin_sleep); }
                    parity[s] ^= 1;
                }
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(kStage) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + s * kStage)),
                             "l"(p.wbuf + off), "r"(kStage), "r"(smem_u32(&bars[s])) : "memory");
                issued++;
                off += static_cast<size_t>(G) * kStage;
                if (off + kStage > p.wbytes) off = (static_cast<size_t>(c) * kStage) % (p.wbytes - kStage);
                if (p.stream > 1) __nanosleep(p.stream);   // throttle
            }
            // drain: every stage used so far has exactly one completion nobody waited for
            for (int s = 0; s < kStages && s < issued; s++) {
                while (!mbar_try(smem_u32(&bars[s]), parity[s])) {}
            }
            __nanosleep(20000);
        }
        return;
    }
    const int r0 = static_cast<int>(static_cast<long long>(K) * c / G), r1 = static_cast<int>(static_cast<long long>(K) * (c + 1) / G);
    float sink = 0.f;
    uint32_t tma_parity = 0;
    unsigned long long spins = 0;   // a wedged poll traps instead of hanging the GPU
    const long long t0 = clock64();
    for (int it = 1; it <= p.iters; it++) {
        unsigned long long* base = p.buf + static_cast<size_t>(it & 1) * p.R * Kpad;
        const uint32_t seq = it;
        // ---- produce: one word per thread for this CTA's rows, into every replica ----
        if (tid < r1 - r0) {
            const float v = 1.0f + 0.001f * (r0 + tid) + sink * 0.f;
            if (p.variant == 3 || p.variant == 4) {
                for (int r = 0; r < p.R; r++) reinterpret_cast<float*>(base + static_cast<size_t>(r) * Kpad)[2 * (r0 + tid)] = v;
                if (p.variant == 4) __threadfence();
            } else {
                for (int r = 0; r < p.R; r++) ll_st(base + static_cast<size_t>(r) * Kpad + r0 + tid, v, seq);
            }
        }
        if (p.variant == 3) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) {
                asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p.counter), "l"(1ull) : "memory");
                unsigned long long v;
                do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p.counter) : "memory"); } while (v < static_cast<unsigned long long>(it) * G);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        } else if (p.variant == 4) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p.flags + c), "l"(static_cast<unsigned long long>(it)) : "memory");
            if (w == 0) {
                for (;;) {
                    bool ok = true;
                    for (int cc = lane; cc < G; cc += 32) {
                        unsigned long long v;
                        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p.flags + cc) : "memory");
                        ok = ok && v >= static_cast<unsigned long long>(it);
                    }
                    if (__all_sync(0xffffffffu, ok)) break;
                    if (p.sleep_ns) __nanosleep(p.sleep_ns);
                    if (++spins > (1ull << 25)) __trap();
                }
                __threadfence();
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        // ---- consume ----
        const unsigned long long* src = base + static_cast<size_t>(c % p.R) * Kpad;
        if (p.variant == 0) {
            for (int k0 = w * 256; k0 < K; k0 += kWarps * 256) {
                for (;;) {
                    uint4 wd[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) wd[j] = ll_ld2(src + k0 + j * 64 + lane * 2);
                    bool ok = true;
#pragma unroll
                    for (int j = 0; j < 4; j++) ok = ok && wd[j].y == seq && wd[j].w == seq;
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 4; j++) { xs[k0 + j * 64 + lane * 2] = __uint_as_float(wd[j].x); xs[k0 + j * 64 + lane * 2 + 1] = __uint_as_float(wd[j].z); }
                        break;
                    }
                    if (p.sleep_ns) __nanosleep(p.sleep_ns);
                    if (++spins > (1ull << 25)) __trap();
                }
            }
        } else if (p.variant == 6) {
            float acc = 0.f;
            for (int k0 = w * 256; k0 < K; k0 += kWarps * 256) {
                for (;;) {
                    uint4 wd[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) wd[j] = ll_ld2(src + k0 + j * 64 + lane * 2);
                    bool ok = true;
#pragma unroll
                    for (int j = 0; j < 4; j++) ok = ok && wd[j].y == seq && wd[j].w == seq;
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 4; j++) acc += __uint_as_float(wd[j].x) * __uint_as_float(wd[j].x) + __uint_as_float(wd[j].z) * __uint_as_float(wd[j].z);
                        break;
                    }
                    if (p.sleep_ns) __nanosleep(p.sleep_ns);
                    if (++spins > (1ull << 25)) __trap();
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            sink += rsqrtf(acc);
            asm volatile("bar.sync 1, 256;" ::: "memory");
        } else if (p.variant == 1) {
            // each warp polls a 2048-word slice q = w & 3 into registers (no fan-out): 16 x 16-byte loads per lane in two halves
            const int q = w & (K / 2048 - 1);
            float acc = 0.f;
            for (int half = 0; half < 2; half++) {
                const unsigned long long* s2 = src + q * 2048 + half * 1024 + lane * 2;
                for (;;) {
                    uint4 wd[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) wd[j] = ll_ld2(s2 + j * 64);
                    bool ok = true;
#pragma unroll
                    for (int j = 0; j < 16; j++) ok = ok && wd[j].y == seq && wd[j].w == seq;
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 16; j++) acc += __uint_as_float(wd[j].x) + __uint_as_float(wd[j].z);
                        break;
                    }
                    if (p.sleep_ns) __nanosleep(p.sleep_ns);
                    if (++spins > (1ull << 25)) __trap();
                }
            }
            sink += acc;
            asm volatile("bar.sync 1, 256;" ::: "memory");   // (the real kernel's phases end in a CTA-wide dependency too)
        } else if (p.variant == 5) {
            // every warp polls the WHOLE vector (K <= 2048) straight into registers: no fan-out, no CTA barrier before the rows
            float acc = 0.f;
            for (int half = 0; half < K / 1024; half++) {
                const unsigned long long* s2 = src + half * 1024 + lane * 2;
                for (;;) {
                    uint4 wd[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) wd[j] = ll_ld2(s2 + j * 64);
                    bool ok = true;
#pragma unroll
                    for (int j = 0; j < 16; j++) ok = ok && wd[j].y == seq && wd[j].w == seq;
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 16; j++) acc += __uint_as_float(wd[j].x) * __uint_as_float(wd[j].x) + __uint_as_float(wd[j].z) * __uint_as_float(wd[j].z);
                        break;
                    }
                    if (p.sleep_ns) __nanosleep(p.sleep_ns);
                    if (++spins > (1ull << 25)) __trap();
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            sink += rsqrtf(acc);
            asm volatile("bar.sync 1, 256;" ::: "memory");
        } else if (p.variant == 2) {
            for (;;) {
                if (tid == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[kStages])), "r"(K * 8) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(tma_dst)), "l"(src),
                                 "r"(K * 8), "r"(smem_u32(&bars[kStages])) : "memory");
                    vote = 1;
                }
                while (!mbar_try(smem_u32(&bars[kStages]), tma_parity)) {}
                tma_parity ^= 1;
                bool ok = true;
                for (int k = tid * 2; k < K; k += 512) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(tma_dst + k);
                    ok = ok && static_cast<uint32_t>(v.x >> 32) == seq && static_cast<uint32_t>(v.y >> 32) == seq;
                    xs[k] = __uint_as_float(static_cast<uint32_t>(v.x));
                    xs[k + 1] = __uint_as_float(static_cast<uint32_t>(v.y));
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (!ok) vote = 0;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const int v = vote;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (v) break;
                if (p.sleep_ns) __nanosleep(p.sleep_ns);
                    if (++spins > (1ull << 25)) __trap();
            }
        } else {
            for (int k = tid * 4; k < K; k += 1024) {
                const float4 a = __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + 2 * k));
                const float4 b = __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + 2 * k + 4));
                xs[k] = a.x; xs[k + 1] = a.z; xs[k + 2] = b.x; xs[k + 3] = b.z;
            }
        }
        if (p.variant != 1 && p.variant != 5 && p.variant != 6) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // every warp reads the whole vector (<= 2048) / its slice and reduces it (RMSNorm sum of squares)
            float ss = 0.f;
            const int kk = K > 2048 ? (w & (K / 2048 - 1)) * 2048 : 0, n = K > 2048 ? 2048 : K;
            for (int k = lane * 8; k < n; k += 256) {
                const float4 a = *reinterpret_cast<const float4*>(xs + kk + k), b = *reinterpret_cast<const float4*>(xs + kk + k + 4);
                ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            sink += rsqrtf(ss);
            asm volatile("bar.sync 1, 256;" ::: "memory");   // xs is rewritten next iteration
        }
        // ---- emulate the rows of the phase ----
        if (p.work_ns > 0) {
            const unsigned h = (static_cast<unsigned>(c) * 2654435761u + static_cast<unsigned>(it) * 40503u) >> 16;
            const unsigned long long until = gtime() + p.work_ns + (p.jitter_ns ? h % p.jitter_ns : 0);
            while (gtime() < until) {}
        }
    }
    const long long t1 = clock64();
    if (tid == 0) {
        stop = 1;
        p.sink[c] = sink;
        if (c == 0) *p.cycles = t1 - t0;
    }
}

int main(int argc, char** argv) {
    int dev = 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    const int G = prop.multiProcessorCount;
    printf("%s, %d SMs, clock %d MHz\n", prop.name, G, clk_khz / 1000);
    P p{};
    const int maxK = 8192, maxR = 8;
    CK(cudaMalloc(&p.buf, sizeof(unsigned long long) * 2 * maxR * maxK));
    p.wbytes = static_cast<size_t>(2) << 30;
    void* wb;
    CK(cudaMalloc(&wb, p.wbytes));
    CK(cudaMemset(wb, 1, p.wbytes));
    p.wbuf = static_cast<const unsigned char*>(wb);
    CK(cudaMalloc(&p.counter, 8));
    CK(cudaMalloc(&p.flags, 8 * 1024));
    CK(cudaMalloc(&p.sink, 4 * 1024));
    CK(cudaMallocManaged(&p.cycles, 8));
    const size_t smem = 32768 + 65536 + kStages * kStage;
    CK(cudaFuncSetAttribute(handoff, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    auto run = [&](int variant, int K, int R, int sleep_ns, int stream, int work_ns, int jitter_ns) {
        p.variant = variant; p.K = K; p.R = R; p.sleep_ns = sleep_ns; p.stream = stream; p.work_ns = work_ns; p.jitter_ns = jitter_ns;
        p.iters = 2000;
        CK(cudaMemset(p.buf, 0, sizeof(unsigned long long) * 2 * maxR * maxK));
        CK(cudaMemset(p.counter, 0, 8));
        CK(cudaMemset(p.flags, 0, 8 * 1024));
        void* args[] = {&p};
        CK(cudaLaunchCooperativeKernel((void*)handoff, dim3(G), dim3(kThreads), args, smem, 0));
        CK(cudaDeviceSynchronize());
        const double us = static_cast<double>(*p.cycles) / p.iters / (clk_khz / 1000.0);
        printf("variant %d K %5d replicas %d sleep %3d stream %4d work %4d+%3d ns: cycle %.2f us, hand-off %.2f us\n", variant, K, R, sleep_ns, stream, work_ns, jitter_ns,
               us, us - work_ns / 1000.0 - jitter_ns / 1000.0);
        fflush(stdout);
    };
    for (int spin : {0, 50}) {
        p.spin_sleep = spin;
        printf("--- stream warp sleeps %d ns between mbarrier probes\n", spin);
        for (int stream : {1, 0, 100, 300}) {
            for (int work : {0, 1500}) {
                const int jit = work ? 500 : 0;
                const int K = 2048;
                run(0, K, 1, 0, stream, work, jit);
                run(6, K, 1, 0, stream, work, jit);
                run(0, K, 1, 300, stream, work, jit);
                run(6, K, 1, 300, stream, work, jit);
                run(3, K, 1, 0, stream, work, jit);
                run(1, 8192, 1, 0, stream, work, jit);
            }
        }
    }
    return 0;
}

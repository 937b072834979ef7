// common.cuh -- device helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2l {

constexpr int kWarp = 32;

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ float bf16_bits_to_f32(uint16_t b) { return __uint_as_float(static_cast<uint32_t>(b) << 16); }

// round-to-nearest-even fp32 -> bf16 bits (finite inputs)
__device__ __forceinline__ uint16_t f32_to_bf16_bits(float f) {
    uint32_t u = __float_as_uint(f);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return static_cast<uint16_t>(u >> 16);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    return static_cast<uint32_t>(f32_to_bf16_bits(lo)) | (static_cast<uint32_t>(f32_to_bf16_bits(hi)) << 16);
}

// 128-bit streaming load: weights are read exactly once per token -> bypass L1 allocation
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// dot of 8 bf16 weights (one uint4) with 8 fp32 activations
__device__ __forceinline__ float dot8(const uint4& w, const float4& x0, const float4& x1, float acc) {
    acc = fmaf(bf16lo(w.x), x0.x, acc);
    acc = fmaf(bf16hi(w.x), x0.y, acc);
    acc = fmaf(bf16lo(w.y), x0.z, acc);
    acc = fmaf(bf16hi(w.y), x0.w, acc);
    acc = fmaf(bf16lo(w.z), x1.x, acc);
    acc = fmaf(bf16hi(w.z), x1.y, acc);
    acc = fmaf(bf16lo(w.w), x1.z, acc);
    acc = fmaf(bf16hi(w.w), x1.w, acc);
    return acc;
}

// orderable key for greedy argmax: larger value wins, ties -> lower index (first max)
__device__ __forceinline__ unsigned long long argmax_key(float v, int idx) {
    uint32_t u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return (static_cast<unsigned long long>(u) << 32) | static_cast<uint32_t>(0xFFFFFFFFu - static_cast<uint32_t>(idx));
}
__device__ __forceinline__ int argmax_key_index(unsigned long long k) {
    return static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(k & 0xFFFFFFFFull));
}

// Programmatic dependent launch (PDL): let the next kernel's prologue overlap our tail, and
// wait for the previous kernel's writes before touching its outputs.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Thread-block clusters: rank, cluster-wide barrier (release / acquire), loads from a peer CTA's shared memory (DSMEM)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(const void* local_smem_ptr, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(local_smem_ptr))), "r"(rank));
    return remote;
}
__device__ __forceinline__ float dsmem_ld_f32(const float* local_smem_ptr, uint32_t rank) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(dsmem_addr(local_smem_ptr, rank)) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long dsmem_ld_u64(const unsigned long long* local_smem_ptr, uint32_t rank) {
    unsigned long long v;
    asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(dsmem_addr(local_smem_ptr, rank)) : "memory");
    return v;
}

}  // namespace b2l

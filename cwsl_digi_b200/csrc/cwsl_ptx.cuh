// Device-side PTX helpers shared by the sm_100a kernels (packed fp32, mbarrier, TMA bulk copy).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace cwsl {

// ------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long as_u64(float2 v) {
    return *reinterpret_cast<unsigned long long*>(&v);
}
__device__ __forceinline__ float2 as_f2(unsigned long long v) { return *reinterpret_cast<float2*>(&v); }

// packed 2 x fp32 (Blackwell FFMA2/FMUL2/FADD2): one issue slot, two lanes of the FMA pipe
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(as_u64(a)), "l"(as_u64(b)), "l"(as_u64(c)));
    return as_f2(d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(as_u64(a)), "l"(as_u64(b)));
    return as_f2(d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(as_u64(a)), "l"(as_u64(b)));
    return as_f2(d);
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(as_u64(a)), "l"(as_u64(b)));
    return as_f2(d);
}
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }  // scalar broadcast operand
// EXACT-mode add of a packed product: ptxas 12.9 contracts mul.rn.f32x2 followed by add.rn.f32x2 into FFMA2
// (even with -fmad=false; only scalar mul.rn/add.rn are left alone), so the add after a packed multiply is
// issued as two scalar add.rn.f32. Same FMA-pipe time (a packed op occupies the pipe for two cycles).
__device__ __forceinline__ float2 add_unfused(float2 a, float2 prod) {
    return make_float2(__fadd_rn(a.x, prod.x), __fadd_rn(a.y, prod.y));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
// TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void atomic_max_abs(unsigned* addr, float warp_local_max) {
    const unsigned bits = __float_as_uint(warp_local_max);  // non-negative floats order like uints
    const unsigned m = __reduce_max_sync(0xffffffffu, bits);
    if ((threadIdx.x & 31) == 0 && m != 0) atomicMax(addr, m);
}

}  // namespace cwsl

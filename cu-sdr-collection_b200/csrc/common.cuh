// Shared device/host helpers for the correlator engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gc {

constexpr double kTwoPi = 6.283185307179586476925286766559;

// ---- complex helpers (float2 = re, im) -------------------------------------------------
// Packed fp32x2 forms (FMUL2 + FFMA2 with a swizzled operand): two issue slots per complex multiply.
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return __ffma2_rn(make_float2(-a.y, a.x), make_float2(b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b)   // a * conj(b)
{
    return __ffma2_rn(make_float2(a.y, -a.x), make_float2(b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
}
// |z| with the approximate square root (MUFU.SQRT, relative error <= 2^-23): the magnitudes are
// summed over the non-coherent blocks and compared at 1e-6 relative, the IEEE sequence would cost
// eight more instructions per code phase.
__device__ __forceinline__ float cabs_fast(float re, float im)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(re, re, im * im)));
    return r;
}

// ---- 64-bit fixed-point phase (turns, 0.64) ---------------------------------------------
// Carrier phases are kept as unsigned 64-bit fractions of a turn: phase(n) = phase0 + n*dphi
// wraps for free and is exact to 2^-64 turns, so the float argument handed to sincospif()
// carries only its own 2^-25-turn rounding instead of the 1e-5 rad error a float
// `2*pi*f*t` of several hundred radians would have.
__host__ __device__ __forceinline__ uint64_t turns_to_fix(double turns)
{
    double f = turns - rint(turns);                 // [-0.5, 0.5]
    long long q = (long long)llrint(ldexp(f, 63));  // [-2^62, 2^62]
    return ((uint64_t)q) << 1;
}
// e^{-i*2*pi*phase}: returns (cos, sin) of the phase; caller applies the sign
__device__ __forceinline__ void fix_sincos(uint64_t phase, float* s, float* c)
{
    const int32_t top = (int32_t)(phase >> 32);          // signed turns * 2^32
    const float t2 = (float)top * 4.656612873077393e-10f;  // 2*turns  (2^-31)
    sincospif(t2, s, c);
}

// ---- mbarrier + bulk async copy (TMA engine, 1-D) ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared bulk copy executed by the TMA unit; completion is signalled on `bar`.
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace gc

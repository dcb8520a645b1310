// common.cuh -- shared host/device helpers for librtod (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/rtod.h"

namespace rtod {

// ---------------------------------------------------------------------------------------
// error reporting: thread-local message, negative return codes (include/rtod.h)
// ---------------------------------------------------------------------------------------
char* last_error_buffer();          // api.cu
constexpr int kErrBuf = 512;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buffer(), kErrBuf, fmt, ap);
    va_end(ap);
    return code;
}

#define RTOD_CUDA_OK(expr)                                                                    \
    do {                                                                                      \
        cudaError_t err__ = (expr);                                                           \
        if (err__ != cudaSuccess)                                                             \
            return ::rtod::fail(RTOD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                \
                                cudaGetErrorString(err__), __FILE__, __LINE__);               \
    } while (0)

#define RTOD_LAUNCH_OK(what)                                                                  \
    do {                                                                                      \
        cudaError_t err__ = cudaGetLastError();                                               \
        if (err__ != cudaSuccess)                                                             \
            return ::rtod::fail(RTOD_ERR_CUDA, "launch of %s failed: %s (%s:%d)", what,       \
                                cudaGetErrorString(err__), __FILE__, __LINE__);               \
    } while (0)

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int kNumSMs = 148;        // B200: 2 dies x 74 SMs

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float leaky01(float v) { return v > 0.0f ? v : 0.1f * v; }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// 16-bit activation / weight storage: fp16 (default: 11 significant bits, what the prediction-tensor tolerance
// needs on a non-degenerate network) or bf16 (the north-star wording; 8 bits).  Same tcgen05 kind::f16 rate.
// fp16 packing saturates to +-65504 instead of producing inf.
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float f16_lo(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v & 0xFFFFu))); }
__device__ __forceinline__ float f16_hi(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v >> 16))); }
template <bool kF16> __device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    if constexpr (kF16) return pack_f16x2(lo, hi);
    else return pack_bf16x2(lo, hi);
}
template <bool kF16> __device__ __forceinline__ float h2_lo(uint32_t v) {
    if constexpr (kF16) return f16_lo(v);
    else return bf16_lo(v);
}
template <bool kF16> __device__ __forceinline__ float h2_hi(uint32_t v) {
    if constexpr (kF16) return f16_hi(v);
    else return bf16_hi(v);
}
template <bool kF16> __device__ __forceinline__ float h1_to_float(unsigned short v) {
    if constexpr (kF16) return __half2float(__ushort_as_half(v));
    else return __uint_as_float((uint32_t)v << 16);
}
template <bool kF16> __device__ __forceinline__ unsigned short float_to_h1(float v) {
    return (unsigned short)(pack_h2<kF16>(v, 0.0f) & 0xFFFFu);
}

// logistic function (reference: torch.sigmoid, fp32): full-precision exp, reciprocal by MUFU.RCP (<= 1 ulp;
// 1 + exp(-v) is never denormal, so the flush-to-zero variant is exact in range).  The correctly rounded
// reciprocal costs ~10 more instructions per element and made the decode kernels issue-bound.
__device__ __forceinline__ float sigmoid_f32(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fadd_rn(1.0f, expf(-v))));
    return r;
}

// same function through the two special-function-unit approximations only (exp2 of v * log2(e), reciprocal):
// 4 instructions instead of 12.  Error <= ~1e-6 relative for |v| < 16 and far below 1e-6 absolute beyond
// (the product rounding grows with |v|, where the sigmoid is already ~0 or ~1): inside the decode tolerance
// (tests: rtol 2e-6, atol 1e-6) and three orders of magnitude inside the prediction-tensor contract.  Used for
// the class scores / objectness in the forward plan's decode launch, which is issue-bound otherwise.
__device__ __forceinline__ float sigmoid_fast_f32(float v) {
    float t, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
    return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

}  // namespace rtod

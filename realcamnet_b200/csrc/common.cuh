// Shared helpers for the realcamnet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rcn_b200.h"

namespace rcn {

void set_error(const char* fmt, ...);

#define RCN_CHECK_ARG(cond, ...)                  \
    do {                                          \
        if (!(cond)) {                            \
            rcn::set_error(__VA_ARGS__);          \
            return RCN_ERR_INVALID;               \
        }                                         \
    } while (0)

#define RCN_CHECK_LAUNCH(name)                                                         \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            rcn::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
            return RCN_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)

// count of kernel launches issued by this library (bench.py's "gpu_launches")
extern unsigned long long g_launches;
inline void count_launch(int n = 1) { g_launches += (unsigned long long)n; }

// erf for the GELU epilogues: Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7 (+ fp32 rounding), branch-free: one MUFU.RCP, one
// MUFU.EX2 and 8 FMA-class instructions, against ~25 instructions and a divergent branch in erff().  The absolute error it adds to
// GELU, 0.5*|v|*5e-7 (measured max |erf error| in fp32: 4.7e-7), is four orders of magnitude below the 1e-3 parity bar.
__device__ __forceinline__ float erf_as(float x) {
    const float ax = fabsf(x);
    float t;   // MUFU.RCP (1 ulp; the IEEE __frcp_rn would add a slow-path call per element)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float r = 1.f - p * t * __expf(-ax * ax);
    return copysignf(r, x);
}

// GELU (erf form) with the same A&S 7.1.26 erf, rearranged so that the sign handling and the 0.5 factors disappear:
//   0.5 v (1 + erf(v / sqrt 2)) = max(v, 0) - |v| exp(-v^2 / 2) * q(t),  t = 1 / (1 + p |v| / sqrt 2),  q = 0.5 t (a1 + t (a2 + ...))
// 6 FFMA + 4 FMUL + FMNMX + 2 MUFU per element (the direct form costs 8 FFMA + 8 FMUL + 3 FADD + selects: the GELU epilogues of
// the Swin MLPs are bound by exactly this instruction count, profiles/r2_conv_epilogue_ncu.md).
__device__ __forceinline__ float gelu_as(float v) {
    const float av = fabsf(v);
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, av, 1.f)));
    float q = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
    q = fmaf(q, t, 0.5f * 1.421413741f);
    q = fmaf(q, t, 0.5f * -0.284496736f);
    q = fmaf(q, t, 0.5f * 0.254829592f);
    q *= t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * v * (-0.5f * 1.4426950408889634f)));
    return fmaf(-(av * e), q, fmaxf(v, 0.f));
}

__device__ __forceinline__ float act_apply(float v, int act, float slope) {
    switch (act) {
        case RCN_ACT_RELU: return v > 0.f ? v : 0.f;
        case RCN_ACT_LRELU: return v > 0.f ? v : v * slope;
        case RCN_ACT_GELU: return gelu_as(v);
        case RCN_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case RCN_ACT_HALF_TANH: return 0.5f * tanhf(v);
        case RCN_ACT_CLAMP01: return fminf(fmaxf(v, 0.f), 1.f);
        case RCN_ACT_HSWISH: return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
        default: return v;
    }
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace rcn

// blend_core.cuh -- the K x K occlusion term shared by the stand-alone aggregation kernels
// (blend.cu) and the fused renderer kernels (render.cu).  Reference: VoGE/Aggregation.py:30-79.
#pragma once
#include "common.cuh"

namespace voge {

constexpr float kInvSqrtPi = 0.5641895835477563f;
constexpr float kInvExpMinusHalf = 1.6487212707001282f;  // 1 / exp(-0.5), Aggregation.py:79
constexpr float kErfSat = 4.0f;                          // |c| >= 4  =>  erff(c) == +-1 in fp32

// (erf(c) + 1) / 2 with the saturated branches short-cut
__device__ __forceinline__ float phi(float c) {
    if (c >= kErfSat) return 1.f;
    if (c <= -kErfSat) return 0.f;
    return (erff(c) + 1.f) * 0.5f;
}

// Phi and exp(-c^2) for |c| < kErfSat through Abramowitz-Stegun 7.1.26 (erfc(x) = poly(1/(1+px)) exp(-x^2),
// |error| <= 1.5e-7) on MUFU ex2 / rcp: ~18 instructions against ~30 for erff alone.  Used by the fused
// kernels, whose erf arguments are the few in-window neighbour pairs of a pixel: the error it adds to a blend
// weight is <= omega * sum_k E_k * 1e-7, two orders below the 1e-5 relative tolerance of the weights.
__device__ __forceinline__ float phi_fast(float c, float& e_out) {
    const float ax = fabsf(c);
    const float e = __expf(-ax * ax);
    const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float half_erfc = 0.5f * p * t * e;
    e_out = e;
    return c >= 0.f ? 1.f - half_erfc : half_erfc;
}

// (erf(c) + 1) / 2 with the saturated branches short-cut, fast variant
__device__ __forceinline__ float phi_f(float c) {
    if (c >= kErfSat) return 1.f;
    if (c <= -kErfSat) return 0.f;
    float e;
    return phi_fast(c, e);
}

}  // namespace voge

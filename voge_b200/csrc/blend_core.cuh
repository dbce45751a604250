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

}  // namespace voge

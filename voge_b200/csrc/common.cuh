// common.cuh -- shared device helpers for libvoge_b200 (sm_100a only).
//
// The "exact" helpers reproduce, bit for bit, the fp32 rounding sequence that nvcc emits for
// the reference's Innerdot3d (reference VoGE/csrc/ray_trace_voge/ray_trace_voge.cu:11-38,
// evaluated at :188-193) -- see DESIGN.md "rounding contract".  They are written with
// __fmul_rn/__fmaf_rn/__fdiv_rn so that ptxas can neither contract nor re-associate them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define VOGE_LAUNCH_CHECK()                        \
    do {                                           \
        cudaError_t e__ = cudaGetLastError();      \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

#define VOGE_CUDA_TRY(expr)                        \
    do {                                           \
        cudaError_t e__ = (expr);                  \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

namespace voge {

constexpr float kEmptyLen = 1e10f;   // init value of len/act slots, ray_trace_voge.cu:245-246
constexpr int kNumSMs = 148;         // B200

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- exact (reference-rounded) quadratic forms ---------------------------------------------
// t_ij = rn(a_i * S_ij): shared between the a.S.d and a.S.a forms (the reference binary CSEs them).
struct Prod9 {
    float t[9];
};

__device__ __forceinline__ Prod9 exact_row_products(float a0, float a1, float a2, const float* S) {
    Prod9 p;
    p.t[0] = __fmul_rn(a0, S[0]); p.t[1] = __fmul_rn(a0, S[1]); p.t[2] = __fmul_rn(a0, S[2]);
    p.t[3] = __fmul_rn(a1, S[3]); p.t[4] = __fmul_rn(a1, S[4]); p.t[5] = __fmul_rn(a1, S[5]);
    p.t[6] = __fmul_rn(a2, S[6]); p.t[7] = __fmul_rn(a2, S[7]); p.t[8] = __fmul_rn(a2, S[8]);
    return p;
}

// acc = fma(t11,c1, rn(t12*c2)); then fma(t13,c3,acc), fma(t21,c1,acc), ... fma(t33,c3,acc)
__device__ __forceinline__ float exact_contract(const Prod9& p, float c0, float c1, float c2) {
    float acc = __fmaf_rn(p.t[0], c0, __fmul_rn(p.t[1], c1));
    acc = __fmaf_rn(p.t[2], c2, acc);
    acc = __fmaf_rn(p.t[3], c0, acc);
    acc = __fmaf_rn(p.t[4], c1, acc);
    acc = __fmaf_rn(p.t[5], c2, acc);
    acc = __fmaf_rn(p.t[6], c0, acc);
    acc = __fmaf_rn(p.t[7], c1, acc);
    acc = __fmaf_rn(p.t[8], c2, acc);
    return acc;
}

struct Hit {
    float len, act, dsd;
};

// One (ray, Gaussian) pair exactly as ray_trace_voge.cu:188-193 computes it.
__device__ __forceinline__ Hit exact_pair(float m0, float m1, float m2, const float* S, float d0,
                                          float d1, float d2) {
    const Prod9 pd = exact_row_products(d0, d1, d2, S);
    const Prod9 pm = exact_row_products(m0, m1, m2, S);
    const float ksk = exact_contract(pd, d0, d1, d2);
    const float msk = exact_contract(pm, d0, d1, d2);
    const float msm = exact_contract(pm, m0, m1, m2);
    Hit h;
    h.len = __fdiv_rn(msk, ksk);
    h.act = __fsub_rn(msm, __fdiv_rn(__fmul_rn(msk, msk), ksk));
    h.dsd = ksk;
    return h;
}

// ---- NDC helpers (reference rasterize_points/rasterization_utils.cuh:16-42) ------------------
__host__ __device__ inline float ndc_range(int S1, int S2) {
    float range = 2.0f;
    if (S1 > S2) range = ((float)S1 * range) / (float)S2;
    return range;
}
__device__ __forceinline__ float pix_to_ndc(int i, int S1, int S2) {
    const float range = ndc_range(S1, S2);
    const float offset = range / 2.0f;
    return __fadd_rn(-offset, __fdiv_rn(__fmaf_rn(range, (float)i, offset), (float)S1));
}

}  // namespace voge

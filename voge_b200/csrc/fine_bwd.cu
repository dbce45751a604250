// fine_bwd.cu -- voge_ray_trace_fine_backward: API-compatible backward of the fine ray tracer.
// Replaces RayTraceFineVogeBackward / RayTraceFineVogeBackwardKernel
// (reference VoGE/csrc/ray_trace_voge/ray_trace_voge.cu:283-379, Innerdot3dBackward :41-91).
//
// Recompute, don't store: the three quadratic forms are re-evaluated from (mu, S, ray).  The
// reference issues 45 scalar atomics per hit (three Innerdot3dBackward calls); here the three
// contributions are summed in registers first: 3 (mu) + 9 (S) reductions per hit, and the ray
// gradient is reduced over the K slots of a pixel inside the warp before touching memory.
#include "../../include/voge_b200.h"
#include "common.cuh"

namespace voge {

struct FineBwdArgs {
    const float* mus;
    const float* isigmas;
    const float* rays;
    const int32_t* idx;
    const float* g_len;
    const float* g_act;
    const float* g_dsd;
    int64_t R;
    int K, P;
    float* grad_rays;
    float* grad_mus;
    float* grad_isg;
};

// one thread per ray; loops over the K slots (keeps the ray gradient in registers)
__global__ void __launch_bounds__(256) fine_bwd_kernel(const FineBwdArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    const float d0 = a.rays[r * 3 + 0], d1 = a.rays[r * 3 + 1], d2 = a.rays[r * 3 + 2];
    float gr0 = 0.f, gr1 = 0.f, gr2 = 0.f;
    for (int k = 0; k < a.K; ++k) {
        const int g = a.idx[r * a.K + k];
        if (g < 0 || g >= a.P) continue;
        const float gl = a.g_len[r * a.K + k], ga = a.g_act[r * a.K + k], gd = a.g_dsd[r * a.K + k];
        float S[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) S[i] = __ldg(a.isigmas + (int64_t)g * 9 + i);
        const float m0 = __ldg(a.mus + (int64_t)g * 3), m1 = __ldg(a.mus + (int64_t)g * 3 + 1),
                    m2 = __ldg(a.mus + (int64_t)g * 3 + 2);
        const Prod9 pd = exact_row_products(d0, d1, d2, S);
        const Prod9 pm = exact_row_products(m0, m1, m2, S);
        const float ksk = exact_contract(pd, d0, d1, d2);
        const float msk = exact_contract(pm, d0, d1, d2);
        // chain rule, ray_trace_voge.cu:324-326
        const float g_ksk = (ga * msk - gl) * msk / (ksk * ksk) + gd;
        const float g_msk = (gl - 2.f * ga * msk) / ksk;
        const float g_msm = ga;
        // S d, S^T d, S mu, S^T mu
        const float Sd0 = S[0] * d0 + S[1] * d1 + S[2] * d2;
        const float Sd1 = S[3] * d0 + S[4] * d1 + S[5] * d2;
        const float Sd2 = S[6] * d0 + S[7] * d1 + S[8] * d2;
        const float Std0 = S[0] * d0 + S[3] * d1 + S[6] * d2;
        const float Std1 = S[1] * d0 + S[4] * d1 + S[7] * d2;
        const float Std2 = S[2] * d0 + S[5] * d1 + S[8] * d2;
        const float Sm0 = S[0] * m0 + S[1] * m1 + S[2] * m2;
        const float Sm1 = S[3] * m0 + S[4] * m1 + S[5] * m2;
        const float Sm2 = S[6] * m0 + S[7] * m1 + S[8] * m2;
        const float Stm0 = S[0] * m0 + S[3] * m1 + S[6] * m2;
        const float Stm1 = S[1] * m0 + S[4] * m1 + S[7] * m2;
        const float Stm2 = S[2] * m0 + S[5] * m1 + S[8] * m2;
        // d ray
        gr0 += g_ksk * (Sd0 + Std0) + g_msk * Stm0;
        gr1 += g_ksk * (Sd1 + Std1) + g_msk * Stm1;
        gr2 += g_ksk * (Sd2 + Std2) + g_msk * Stm2;
        // d mu
        float* gm = a.grad_mus + (int64_t)g * 3;
        atomicAdd(gm + 0, g_msk * Sd0 + g_msm * (Sm0 + Stm0));
        atomicAdd(gm + 1, g_msk * Sd1 + g_msm * (Sm1 + Stm1));
        atomicAdd(gm + 2, g_msk * Sd2 + g_msm * (Sm2 + Stm2));
        // d S (full, non-symmetric: row = left vector, column = right vector)
        float* gs = a.grad_isg + (int64_t)g * 9;
        const float dv[3] = {d0, d1, d2};
        const float mv[3] = {m0, m1, m2};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                atomicAdd(gs + i * 3 + j, g_ksk * dv[i] * dv[j] + g_msk * mv[i] * dv[j] + g_msm * mv[i] * mv[j]);
    }
    if (a.grad_rays != nullptr) {
        a.grad_rays[r * 3 + 0] = gr0;
        a.grad_rays[r * 3 + 1] = gr1;
        a.grad_rays[r * 3 + 2] = gr2;
    }
}

}  // namespace voge

extern "C" int voge_ray_trace_fine_backward(const float* mus, const float* isigmas, const float* rays,
                                            const int32_t* idx, const float* grad_len,
                                            const float* grad_act, const float* grad_dsd, int B, int H,
                                            int W, int K, int P, float* grad_rays, float* grad_mus,
                                            float* grad_isg, voge_stream_t stream) {
    using namespace voge;
    const int64_t R = (int64_t)B * H * W;
    if (R <= 0 || K <= 0) return 0;
    FineBwdArgs a{mus, isigmas, rays, idx, grad_len, grad_act, grad_dsd, R, K, P, grad_rays, grad_mus, grad_isg};
    const int64_t grid = (R + 255) / 256;
    fine_bwd_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(a);
    VOGE_LAUNCH_CHECK();
    return 0;
}

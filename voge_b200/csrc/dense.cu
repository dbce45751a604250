// dense.cu -- dense-ray API (next-tier rows of SURVEY.md 8f-1): the same three quadratic forms on
// arbitrary ray lists, without culling or top-K, plus top-K over a dense (N,M) table.
// Replaces RayTraceVogeRay / RayTraceVogeRayBackward / FindNearestK and their kernels
// (reference VoGE/csrc/voge_ray_tracing_ray/voge_ray_tracing_ray.cu:114-143, :147-188, :191-239).
#include "../../include/voge_b200.h"
#include "fine_core.cuh"

namespace voge {

// thread per (ray, point) pair, point index fastest => coalesced (N,M) writes; bit-faithful arithmetic
__global__ void __launch_bounds__(256) ray_dense_fwd_kernel(const float* __restrict__ mus,
                                                            const float* __restrict__ isigmas,
                                                            const float* __restrict__ rays, int M, int64_t NM,
                                                            float* __restrict__ o_len, float* __restrict__ o_act,
                                                            float* __restrict__ o_dsd) {
    const int64_t pid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= NM) return;
    const int64_t ray = pid / M;
    const int p = (int)(pid - ray * M);
    float S[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) S[i] = __ldg(isigmas + (int64_t)p * 9 + i);
    const Hit h = exact_pair(__ldg(mus + 3 * (int64_t)p), __ldg(mus + 3 * (int64_t)p + 1), __ldg(mus + 3 * (int64_t)p + 2), S,
                             __ldg(rays + 3 * ray), __ldg(rays + 3 * ray + 1), __ldg(rays + 3 * ray + 2));
    o_len[pid] = h.len; o_act[pid] = h.act; o_dsd[pid] = h.dsd;
}

// thread per ray, loop over points: the ray gradient stays in registers (the reference issues 45
// atomics per pair, 12 of them onto the ray's own 3 floats)
__global__ void __launch_bounds__(128) ray_dense_bwd_kernel(const float* __restrict__ mus,
                                                            const float* __restrict__ isigmas,
                                                            const float* __restrict__ rays,
                                                            const float* __restrict__ g_len,
                                                            const float* __restrict__ g_act,
                                                            const float* __restrict__ g_dsd, int M, int N,
                                                            float* __restrict__ grad_rays, float* __restrict__ grad_mus,
                                                            float* __restrict__ grad_isg) {
    const int ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= N) return;
    const float d0 = rays[3 * ray], d1 = rays[3 * ray + 1], d2 = rays[3 * ray + 2];
    float gr0 = 0.f, gr1 = 0.f, gr2 = 0.f;
    for (int p = 0; p < M; ++p) {
        const int64_t o = (int64_t)ray * M + p;
        const float gl = g_len[o], ga = g_act[o], gd = g_dsd[o];
        if (gl == 0.f && ga == 0.f && gd == 0.f) continue;
        float S[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) S[i] = __ldg(isigmas + (int64_t)p * 9 + i);
        const float m0 = __ldg(mus + 3 * (int64_t)p), m1 = __ldg(mus + 3 * (int64_t)p + 1), m2 = __ldg(mus + 3 * (int64_t)p + 2);
        const Prod9 pd = exact_row_products(d0, d1, d2, S);
        const Prod9 pm = exact_row_products(m0, m1, m2, S);
        const float ksk = exact_contract(pd, d0, d1, d2);
        const float msk = exact_contract(pm, d0, d1, d2);
        const float g_ksk = (ga * msk - gl) * msk / (ksk * ksk) + gd;
        const float g_msk = (gl - 2.f * ga * msk) / ksk;
        const float g_msm = ga;
        const float Sd0 = S[0] * d0 + S[1] * d1 + S[2] * d2, Sd1 = S[3] * d0 + S[4] * d1 + S[5] * d2, Sd2 = S[6] * d0 + S[7] * d1 + S[8] * d2;
        const float Std0 = S[0] * d0 + S[3] * d1 + S[6] * d2, Std1 = S[1] * d0 + S[4] * d1 + S[7] * d2, Std2 = S[2] * d0 + S[5] * d1 + S[8] * d2;
        const float Sm0 = S[0] * m0 + S[1] * m1 + S[2] * m2, Sm1 = S[3] * m0 + S[4] * m1 + S[5] * m2, Sm2 = S[6] * m0 + S[7] * m1 + S[8] * m2;
        const float Stm0 = S[0] * m0 + S[3] * m1 + S[6] * m2, Stm1 = S[1] * m0 + S[4] * m1 + S[7] * m2, Stm2 = S[2] * m0 + S[5] * m1 + S[8] * m2;
        gr0 += g_ksk * (Sd0 + Std0) + g_msk * Stm0;
        gr1 += g_ksk * (Sd1 + Std1) + g_msk * Stm1;
        gr2 += g_ksk * (Sd2 + Std2) + g_msk * Stm2;
        atomicAdd(grad_mus + 3 * (int64_t)p + 0, g_msk * Sd0 + g_msm * (Sm0 + Stm0));
        atomicAdd(grad_mus + 3 * (int64_t)p + 1, g_msk * Sd1 + g_msm * (Sm1 + Stm1));
        atomicAdd(grad_mus + 3 * (int64_t)p + 2, g_msk * Sd2 + g_msm * (Sm2 + Stm2));
        const float dv[3] = {d0, d1, d2}, mv[3] = {m0, m1, m2};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                atomicAdd(grad_isg + 9 * (int64_t)p + 3 * i + j,
                          g_ksk * dv[i] * dv[j] + g_msk * mv[i] * dv[j] + g_msm * mv[i] * mv[j]);
    }
    grad_rays[3 * ray] = gr0; grad_rays[3 * ray + 1] = gr1; grad_rays[3 * ray + 2] = gr2;
}

// top-K over a dense (N,M) table: thread per ray, sorted lists in shared memory ([k][thread])
template <int NT>
__global__ void __launch_bounds__(NT) find_nearest_k_kernel(const float* __restrict__ len_in,
                                                            const float* __restrict__ act_in,
                                                            const float* __restrict__ dsd_in, float thr_act, int M,
                                                            int K, int N, int32_t* __restrict__ o_idx,
                                                            float* __restrict__ o_len, float* __restrict__ o_act,
                                                            float* __restrict__ o_dsd) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_len = reinterpret_cast<float*>(smem_raw);
    int* s_idx = reinterpret_cast<int*>(s_len + (size_t)K * NT);
    const int tid = threadIdx.x;
    const int ray = blockIdx.x * NT + tid;
    if (ray >= N) return;
    TopK<NT> top;
    top.init(s_len, s_idx, K, tid);
    for (int m = 0; m < M; ++m) {
        const float a = act_in[(int64_t)ray * M + m];
        if (a < thr_act) top.insert(len_in[(int64_t)ray * M + m], m);
    }
    for (int k = 0; k < K; ++k) {
        const int64_t o = (int64_t)ray * K + k;
        if (k < top.cnt) {
            const int m = s_idx[k * NT + tid];
            o_idx[o] = m; o_len[o] = len_in[(int64_t)ray * M + m];
            o_act[o] = act_in[(int64_t)ray * M + m]; o_dsd[o] = dsd_in[(int64_t)ray * M + m];
        } else {   // reference initial values, voge_ray_tracing_ray.cu:344-347
            o_idx[o] = -1; o_len[o] = kEmptyLen; o_act[o] = 0.f; o_dsd[o] = 0.f;
        }
    }
}

}  // namespace voge

extern "C" int voge_ray_trace_ray(const float* mus, const float* isigmas, const float* rays, int M, int N,
                                  float* out_len, float* out_act, float* out_dsd, voge_stream_t stream) {
    using namespace voge;
    const int64_t NM = (int64_t)N * M;
    if (NM <= 0) return 0;
    ray_dense_fwd_kernel<<<(unsigned)((NM + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mus, isigmas, rays, M, NM,
                                                                                          out_len, out_act, out_dsd);
    VOGE_LAUNCH_CHECK();
    return 0;
}

extern "C" int voge_ray_trace_ray_backward(const float* mus, const float* isigmas, const float* rays,
                                           const float* grad_len, const float* grad_act, const float* grad_dsd,
                                           int M, int N, float* grad_rays, float* grad_mus, float* grad_isg,
                                           voge_stream_t stream) {
    using namespace voge;
    if (N <= 0 || M <= 0) return 0;
    ray_dense_bwd_kernel<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        mus, isigmas, rays, grad_len, grad_act, grad_dsd, M, N, grad_rays, grad_mus, grad_isg);
    VOGE_LAUNCH_CHECK();
    return 0;
}

extern "C" int voge_find_nearest_k(const float* len_in, const float* act_in, const float* dsd_in, float thr_act,
                                   int M, int K, int N, int32_t* out_idx, float* out_len, float* out_act,
                                   float* out_dsd, voge_stream_t stream) {
    using namespace voge;
    if (N <= 0 || K <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    auto launch = [&](auto kernel, int nt) -> int {
        const size_t smem = (size_t)K * nt * 8;
        if (smem > 227 * 1024) return (int)cudaErrorInvalidValue;
        VOGE_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<(unsigned)((N + nt - 1) / nt), nt, smem, s>>>(len_in, act_in, dsd_in, thr_act, M, K, N, out_idx, out_len,
                                                               out_act, out_dsd);
        VOGE_LAUNCH_CHECK();
        return 0;
    };
    if (K <= 160) return launch(find_nearest_k_kernel<128>, 128);
    return launch(find_nearest_k_kernel<32>, 32);
}

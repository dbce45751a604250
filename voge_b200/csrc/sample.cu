// sample.cu -- inverse rendering: scatter image features onto Gaussians.
// Replaces SampleVoge / SampleVogeBackward / ScatterMax and their kernels
// (reference VoGE/csrc/sample_voge/sample_voge.cu:35-66, :69-92, :173-209).
//
// Forward: one thread per (ray, k) hit-slot x channel group; channel-contiguous reductions into
// feat[idx,:].  Backward needs NO atomics (the reference atomically adds into the pixel's own
// row, :196-206): one thread per ray owns grad_image[r,:] and grad_weight[r,:].
#include "../../include/voge_b200.h"
#include "common.cuh"

namespace voge {

// thread per (ray*K + k, c) with c fastest: a warp covers 32 consecutive channels / slots
__global__ void __launch_bounds__(256) sample_fwd_kernel(const float* __restrict__ image,
                                                         const float* __restrict__ weight,
                                                         const int32_t* __restrict__ idx, int64_t RK,
                                                         int K, int C, int num_vert,
                                                         float* __restrict__ feat, float* __restrict__ wsum) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int CC = C + 1;  // channel C is the weight-sum column
    if (t >= RK * CC) return;
    const int64_t pid = t / CC;
    const int c = (int)(t - pid * CC);
    const int g = idx[pid];
    if (g < 0 || g >= num_vert) return;
    const float w = weight[pid];
    if (c == C) {
        atomicAdd(wsum + g, w);
    } else {
        const int64_t r = pid / K;
        atomicAdd(feat + (int64_t)g * C + c, image[r * C + c] * w);
    }
}

// Few vertices under many rays (the quick-start cuboid: 866 Gaussians under 65 536 rays): every feat[g, c] receives
// hundreds of reductions and L2 serialises them per address (1.56 ms for C1).  Here the 32 lanes of a warp take 32
// consecutive pixels of an image row and walk the K slots together: neighbouring pixels see the same Gaussian in the
// same slot, so the lanes of a RUN of equal indices add their contributions with a segmented warp scan and only the
// last lane of the run issues the reductions (C = 3 quick-start scene: ~10x fewer atomics; shared-memory tables
// are no alternative, fp32 shared atomics are CAS loops).  Requires C <= 4.
__global__ void __launch_bounds__(256) sample_fwd_runs_kernel(const float* __restrict__ image,
                                                              const float* __restrict__ weight,
                                                              const int32_t* __restrict__ idx, int64_t R,
                                                              int K, int C, int num_vert,
                                                              float* __restrict__ feat, float* __restrict__ wsum) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool live = r < R;
    float px[4] = {0.f, 0.f, 0.f, 0.f};
    if (live)
        for (int c = 0; c < C; ++c) px[c] = image[r * C + c];
    for (int k = 0; k < K; ++k) {
        int g = live ? idx[r * K + k] : -1;
        if (g >= num_vert) g = -1;
        const float w = g >= 0 ? weight[r * K + k] : 0.f;
        if (__ballot_sync(0xffffffffu, g >= 0) == 0u) continue;
        float v[5] = {px[0] * w, px[1] * w, px[2] * w, px[3] * w, w};
        // segmented inclusive scan over runs of equal g (head = first lane of a run)
        const int gl = __shfl_up_sync(0xffffffffu, g, 1);
        const bool head = lane == 0 || gl != g;
        unsigned heads = __ballot_sync(0xffffffffu, head);
        const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));      // first lane of my run
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const float t = __shfl_up_sync(0xffffffffu, v[q], o);
                if (lane - o >= start) v[q] += t;
            }
        }
        const int gn = __shfl_down_sync(0xffffffffu, g, 1);
        const bool tail = lane == 31 || gn != g;
        if (tail && g >= 0) {
            for (int c = 0; c < C; ++c) atomicAdd(feat + (int64_t)g * C + c, v[c]);
            atomicAdd(wsum + g, v[4]);
        }
    }
}

__global__ void __launch_bounds__(256) scatter_max_kernel(const float* __restrict__ weight,
                                                          const int32_t* __restrict__ idx, int64_t RK,
                                                          int num_vert, float* __restrict__ wmax) {
    const int64_t pid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= RK) return;
    const int g = idx[pid];
    if (g < 0 || g >= num_vert) return;
    const float w = weight[pid];
    // wmax starts at 0 and the reference folds with fmaxf (sample_voge.cu:22-32), so negative and
    // NaN weights never change it; for non-negative floats the int ordering of the bit pattern
    // equals the float ordering => one native atomicMax replaces the reference's CAS loop.
    if (w >= 0.f) atomicMax(reinterpret_cast<int*>(wmax + g), __float_as_int(w));
}

// one thread per ray (C small) -- each thread owns its pixel's outputs
__global__ void __launch_bounds__(256) sample_bwd_kernel(const float* __restrict__ image,
                                                         const float* __restrict__ weight,
                                                         const int32_t* __restrict__ idx,
                                                         const float* __restrict__ g_feat,
                                                         const float* __restrict__ g_wsum, int64_t R,
                                                         int K, int C, float* __restrict__ g_image,
                                                         float* __restrict__ g_weight) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    for (int c = 0; c < C; ++c) g_image[r * C + c] = 0.f;
    for (int k = 0; k < K; ++k) {
        const int g = idx[r * K + k];
        float gw = 0.f;
        if (g >= 0) {
            const float w = weight[r * K + k];
            gw = __ldg(g_wsum + g);
            for (int c = 0; c < C; ++c) {
                const float gf = __ldg(g_feat + (int64_t)g * C + c);
                g_image[r * C + c] += w * gf;   // own row: plain RMW, L1-resident
                gw = fmaf(gf, image[r * C + c], gw);
            }
        }
        g_weight[r * K + k] = gw;
    }
}

}  // namespace voge

extern "C" int voge_sample(const float* image, const float* weight, const int32_t* idx, int64_t R, int K,
                           int C, int num_vert, float* feat, float* wsum, voge_stream_t stream) {
    using namespace voge;
    if (R <= 0 || K <= 0) return 0;
    const int64_t total = R * K * (C + 1);
    if (num_vert > 0 && C <= 4 && R * K >= 64 * (int64_t)num_vert) {
        // many reductions per vertex: runs of equal indices across neighbouring pixels are summed in the warp first
        sample_fwd_runs_kernel<<<(unsigned)((R + 255) / 256), 256, 0, (cudaStream_t)stream>>>(image, weight, idx, R, K, C,
                                                                                             num_vert, feat, wsum);
        VOGE_LAUNCH_CHECK();
        return 0;
    }
    sample_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        image, weight, idx, R * K, K, C, num_vert, feat, wsum);
    VOGE_LAUNCH_CHECK();
    return 0;
}

extern "C" int voge_sample_backward(const float* image, const float* weight, const int32_t* idx,
                                    const float* grad_feat, const float* grad_wsum, int64_t R, int K,
                                    int C, float* grad_image, float* grad_weight, voge_stream_t stream) {
    using namespace voge;
    if (R <= 0 || K <= 0) return 0;
    sample_bwd_kernel<<<(unsigned)((R + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        image, weight, idx, grad_feat, grad_wsum, R, K, C, grad_image, grad_weight);
    VOGE_LAUNCH_CHECK();
    return 0;
}

extern "C" int voge_scatter_max(const float* weight, const int32_t* idx, int64_t R, int K, int num_vert,
                                float* wmax, voge_stream_t stream) {
    using namespace voge;
    if (R <= 0 || K <= 0) return 0;
    const int64_t total = R * K;
    scatter_max_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        weight, idx, total, num_vert, wmax);
    VOGE_LAUNCH_CHECK();
    return 0;
}

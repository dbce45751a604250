// knn.cu -- neighbour statistics of the point-cloud converter.
//
// voge_knn_mean_dist replaces the body of `naive_point_cloud_converter` (reference VoGE/Converter/Converters.py:106-111):
// for every point, the n_nearest smallest Euclidean distances to the points of the cloud (its own zero distance
// included), each clipped at thr_max x their mean, averaged.  The reference materialises the (N, N, 3) difference
// tensor in batches on the host; here every thread owns one point, the cloud streams through shared memory in tiles of
// 256 points and the n_nearest smallest squared distances live in registers (insertion into a sorted list).
// O(N^2) distance evaluations at FP32 FMA-pipe speed: 10^5 points take ~20 ms.
#include "../../include/voge_b200.h"
#include "common.cuh"

namespace voge {

constexpr int kKnnMax = 16;

__global__ void __launch_bounds__(256) knn_mean_dist_kernel(const float* __restrict__ pts, int N, int k, float thr_max,
                                                            float* __restrict__ avg_len) {
    __shared__ float sx[256], sy[256], sz[256];
    const int tid = threadIdx.x;
    const int i = blockIdx.x * 256 + tid;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (i < N) { px = pts[3 * (int64_t)i]; py = pts[3 * (int64_t)i + 1]; pz = pts[3 * (int64_t)i + 2]; }
    float best[kKnnMax];
#pragma unroll
    for (int q = 0; q < kKnnMax; ++q) best[q] = 3.0e38f;
    for (int base = 0; base < N; base += 256) {
        const int j = base + tid;
        sx[tid] = j < N ? pts[3 * (int64_t)j] : 3.0e38f;
        sy[tid] = j < N ? pts[3 * (int64_t)j + 1] : 0.f;
        sz[tid] = j < N ? pts[3 * (int64_t)j + 2] : 0.f;
        __syncthreads();
        if (i < N) {
            const int n = min(256, N - base);
            for (int t = 0; t < n; ++t) {
                // (dx^2 + dy^2) + dz^2 with separate roundings, the order of `.pow(2).sum(-1)`
                const float dx = __fsub_rn(px, sx[t]), dy = __fsub_rn(py, sy[t]), dz = __fsub_rn(pz, sz[t]);
                float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (d2 < best[kKnnMax - 1]) {
                    // sorted insertion: the list is ascending, slots >= k stay at their initial bound
#pragma unroll
                    for (int q = 0; q < kKnnMax; ++q) {
                        if (q < k && d2 < best[q]) { const float tmp = best[q]; best[q] = d2; d2 = tmp; }
                    }
                    if (k < kKnnMax) best[kKnnMax - 1] = best[k - 1];      // admission bound = current k-th smallest
                }
            }
        }
        __syncthreads();
    }
    if (i >= N) return;
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < kKnnMax; ++q)
        if (q < k) { best[q] = sqrtf(best[q]); sum += best[q]; }
    const float cap = (sum / (float)k) * thr_max;
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < kKnnMax; ++q)
        if (q < k) acc += fminf(best[q], cap);
    avg_len[i] = acc / (float)k;
}

}  // namespace voge

extern "C" int voge_knn_mean_dist(const float* points, int N, int n_nearest, float thr_max, float* avg_len,
                                  voge_stream_t stream) {
    using namespace voge;
    if (N <= 0) return 0;
    if (n_nearest < 1 || n_nearest > kKnnMax || n_nearest > N) return (int)cudaErrorInvalidValue;
    knn_mean_dist_kernel<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(points, N, n_nearest, thr_max, avg_len);
    VOGE_LAUNCH_CHECK();
    return 0;
}

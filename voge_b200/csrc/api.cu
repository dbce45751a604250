// api.cu -- library info entry points of libvoge_b200.so.
#include "../../include/voge_b200.h"
#include "common.cuh"

extern "C" int voge_version(void) { return 1; }

extern "C" const char* voge_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }

extern "C" int voge_device_sm_count(int* sm_count) {
    int dev = 0;
    VOGE_CUDA_TRY(cudaGetDevice(&dev));
    VOGE_CUDA_TRY(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    return 0;
}

// ---- in-run peak micro-benchmarks (roofline denominators that MEASURED_PEAKS.json lacks) -----------
// voge_peak_fp32: every thread runs `iters` rounds of 16 independent FFMA chains (2 flops each).
// voge_peak_sfu : every thread runs `iters` rounds of 8 independent MUFU.EX2 (ex2.approx).
// The host times the launch with CUDA events; total ops = blocks*threads*iters*{32 flops | 8 sfu-ops}.
namespace voge {
__global__ void __launch_bounds__(256) peak_fp32_kernel(int iters, float seed, float* out) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + (float)(threadIdx.x + i);
    const float m = 0.999f + seed, c = 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456f) out[0] = s;
}
__global__ void __launch_bounds__(256) peak_sfu_kernel(int iters, float seed, float* out) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + 1e-3f * (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456f) out[0] = s;
}
}  // namespace voge

extern "C" int voge_peak_fp32(int blocks, int iters, float* out, voge_stream_t stream) {
    voge::peak_fp32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 0.f, out);
    VOGE_LAUNCH_CHECK();
    return 0;
}

extern "C" int voge_peak_sfu(int blocks, int iters, float* out, voge_stream_t stream) {
    voge::peak_sfu_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, -0.5f, out);
    VOGE_LAUNCH_CHECK();
    return 0;
}

// fine_fwd.cu -- voge_ray_trace_fine: API-compatible fine ray tracing over reference-style
// bin_points candidate lists.  Replaces RayTraceFineVoge / RayTraceFineVogeKernel
// (reference VoGE/csrc/ray_trace_voge/ray_trace_voge.cu:135-280).
#include "../../include/voge_b200.h"
#include "fine_core.cuh"

namespace voge {

constexpr int kChunk = 128;  // candidates staged per round

struct FineArgs {
    const float* mus;
    const float* isigmas;
    const float* rays;
    const int32_t* bin_points;
    const int32_t* bin_counts;  // optional: scan only the first min(count, M) entries
    float thr_act;
    int bin_size, B, H, W, BH, BW, M, K, P, subs;
    int32_t* out_idx;
    float* out_len;
    float* out_act;
    float* out_dsd;
};

// One CTA = NT consecutive pixels (bin-major linear order, like the reference :159-176) of one
// bin.  grid.x = B*BH*BW*subs with the sub-block index fastest.
template <int NT>
__global__ void __launch_bounds__(NT) fine_fwd_kernel(const FineArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_stage = reinterpret_cast<float*>(smem_raw);                      // [kChunk][12]
    float* s_len = s_stage + kChunk * kStageFloats;                           // [K][NT]
    int* s_idx = reinterpret_cast<int*>(s_len + (size_t)a.K * NT);            // [K][NT]
    unsigned short* s_queue = reinterpret_cast<unsigned short*>(s_idx + (size_t)a.K * NT);  // [kQueueCap][NT]

    const int tid = threadIdx.x;
    int blk = blockIdx.x;
    const int sub = blk % a.subs; blk /= a.subs;
    const int bx = blk % a.BW; blk /= a.BW;
    const int by = blk % a.BH;
    const int bi = blk / a.BH;

    const int p = sub * NT + tid;  // pixel number inside the bin
    const int yi = by * a.bin_size + p / a.bin_size;
    const int xi = bx * a.bin_size + p % a.bin_size;
    const bool live = (p < a.bin_size * a.bin_size) && (yi < a.H) && (xi < a.W);
    const int64_t ray = ((int64_t)bi * a.H + yi) * a.W + xi;

    RayMono r;
    r.set(0.f, 0.f, 0.f);
    if (live) r.set(a.rays[ray * 3 + 0], a.rays[ray * 3 + 1], a.rays[ray * 3 + 2]);

    TopK<NT> top;
    top.init(s_len, s_idx, a.K, tid);

    const int64_t bin = ((int64_t)bi * a.BH + by) * a.BW + bx;
    // bin_points == NULL: "no coarse stage" -- every Gaussian bi*M .. bi*M+M-1 of the view is a
    // candidate (the reference materialises this arange table, RayTracing.py:22-26)
    const int32_t* list = a.bin_points != nullptr ? a.bin_points + bin * a.M : nullptr;
    int n = a.M;
    if (a.bin_counts != nullptr && list != nullptr) n = min(a.M, a.bin_counts[bin]);

    int qn = 0;
    auto drain = [&]() {
        for (int j = 0; j < qn; ++j) {
            const int c = s_queue[j * NT + tid];
            const int g = __float_as_int(s_stage[c * kStageFloats + 10]);
            if (g < 0) continue;  // padding can only get here through a NaN ray
            const float* mu = a.mus + (int64_t)g * 3;
            const float* S = a.isigmas + (int64_t)g * 9;
            float Sl[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Sl[i] = __ldg(S + i);
            const Hit h = exact_pair(__ldg(mu), __ldg(mu + 1), __ldg(mu + 2), Sl, r.d0, r.d1, r.d2);
            if (h.act < a.thr_act) top.insert(h.len, g);
        }
        qn = 0;
    };

    for (int base = 0; base < n; base += kChunk) {
        __syncthreads();  // previous chunk fully consumed
        bool any = false;
        for (int t = tid; t < kChunk; t += NT) {
            const int m = base + t;
            int g = -1;
            if (m < n) g = list != nullptr ? list[m] : bi * a.M + m;
            if (g > -1 && g < a.P) {
                float mu[3], S[9];
#pragma unroll
                for (int i = 0; i < 3; ++i) mu[i] = __ldg(a.mus + (int64_t)g * 3 + i);
#pragma unroll
                for (int i = 0; i < 9; ++i) S[i] = __ldg(a.isigmas + (int64_t)g * 9 + i);
                stage_candidate(s_stage + t * kStageFloats, g, mu, S, a.thr_act);
                any = true;
            } else {
                stage_invalid(s_stage + t * kStageFloats);
            }
        }
        // all-padding chunk (user supplied bin_points without counts): skip the filter loop
        if (!__syncthreads_or(any)) continue;
        if (live) {
            const int cn = min(kChunk, n - base);
#pragma unroll 4
            for (int c = 0; c < cn; ++c) {
                if (filter_pass(s_stage + c * kStageFloats, r)) {
                    s_queue[qn * NT + tid] = (unsigned short)c;
                    if (++qn == kQueueCap) drain();
                }
            }
            drain();  // queue entries are chunk-local
        }
    }

    if (!live) return;
    // ---- finalize: recompute act / dsd of the K survivors (bit-identical by construction) ----
    int32_t* o_idx = a.out_idx + ray * a.K;
    float* o_len = a.out_len + ray * a.K;
    float* o_act = a.out_act + ray * a.K;
    float* o_dsd = a.out_dsd + ray * a.K;
    for (int k = 0; k < a.K; ++k) {
        if (k < top.cnt) {
            const int g = s_idx[k * NT + tid];
            const float* mu = a.mus + (int64_t)g * 3;
            const float* S = a.isigmas + (int64_t)g * 9;
            float Sl[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Sl[i] = __ldg(S + i);
            const Hit h = exact_pair(__ldg(mu), __ldg(mu + 1), __ldg(mu + 2), Sl, r.d0, r.d1, r.d2);
            o_idx[k] = g; o_len[k] = h.len; o_act[k] = h.act; o_dsd[k] = h.dsd;
        } else {
            o_idx[k] = -1; o_len[k] = kEmptyLen; o_act[k] = kEmptyLen; o_dsd[k] = 0.f;
        }
    }
}

template <int NT>
static int launch_fine(const FineArgs& a, cudaStream_t stream) {
    const size_t smem = (size_t)kChunk * kStageFloats * 4 + (size_t)a.K * NT * 8 + (size_t)kQueueCap * NT * 2;
    if (smem > 227 * 1024) return (int)cudaErrorInvalidValue;
    VOGE_CUDA_TRY(cudaFuncSetAttribute(fine_fwd_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FineArgs b = a;
    b.subs = cdiv(a.bin_size * a.bin_size, NT);
    const long long grid = (long long)a.B * a.BH * a.BW * b.subs;
    if (grid <= 0 || grid > 2147483647LL) return (int)cudaErrorInvalidValue;
    fine_fwd_kernel<NT><<<(unsigned)grid, NT, smem, stream>>>(b);
    VOGE_LAUNCH_CHECK();
    return 0;
}

}  // namespace voge

extern "C" int voge_ray_trace_fine_counts(const float* mus, const float* isigmas, const float* rays,
                                          const int32_t* bin_points, const int32_t* bin_counts,
                                          float thr_act, int bin_size, int B, int H, int W, int BH,
                                          int BW, int M, int K, int P, int32_t* out_idx,
                                          float* out_len, float* out_act, float* out_dsd,
                                          voge_stream_t stream) {
    using namespace voge;
    if (B <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
    if (bin_size <= 0 || BH < cdiv(H, bin_size) || BW < cdiv(W, bin_size) || M < 0) return (int)cudaErrorInvalidValue;
    FineArgs a{mus, isigmas, rays, bin_points, bin_counts, thr_act, bin_size, B, H, W, BH, BW, M, K, P, 1,
               out_idx, out_len, out_act, out_dsd};
    cudaStream_t s = (cudaStream_t)stream;
    if (K <= 160) return launch_fine<128>(a, s);
    if (K <= 380) return launch_fine<64>(a, s);
    return launch_fine<32>(a, s);
}

extern "C" int voge_ray_trace_fine(const float* mus, const float* isigmas, const float* rays,
                                   const int32_t* bin_points, float thr_act, int bin_size, int B,
                                   int H, int W, int BH, int BW, int M, int K, int P,
                                   int32_t* out_idx, float* out_len, float* out_act, float* out_dsd,
                                   voge_stream_t stream) {
    return voge_ray_trace_fine_counts(mus, isigmas, rays, bin_points, nullptr, thr_act, bin_size, B, H,
                                      W, BH, BW, M, K, P, out_idx, out_len, out_act, out_dsd, stream);
}

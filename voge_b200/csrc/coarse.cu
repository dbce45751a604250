// coarse.cu -- voge_rasterize_coarse: screen-space binning of Gaussian bounding boxes.
// Replaces RasterizeEllipseCoarseCuda (reference VoGE/csrc/rasterize_coarse/rasterize_coarse.cu:
// EllipseBoundingBoxKernel :20-42, RasterizeCoarseCudaKernel :44-188, host :190-305).
//
// Same membership predicate as the reference (bbox = p +- r, skip z < 0, overlap test :116-130
// against bin edges from PixToNonSquareNdc +- half a pixel), evaluated with the same fp32
// operations, but
//   * O(P * (BH + BW)) separable range tests instead of O(P * BH * BW) brute force,
//   * DETERMINISTIC output: indices ascending inside every bin (the reference reserves ranges
//     with atomicAdd per 512-element chunk, :153, so its order changes run to run),
//   * the true per-bin counts are returned; on overflow the first M (ascending) are kept and the
//     host raises, instead of a device-side printf and dropped chunks (:154-170),
//   * no shared-memory size cliff: the reference needs BH*BW*64 B of dynamic shared memory
//     without opting in (:229-234), i.e. it cannot launch for 32x32 bins (512^2 / 1024^2 images).
//
// Three kernels over (view, chunk of 2048 Gaussians): histogram -> per-bin exclusive scan over
// chunks -> ordered fill (ranks inside a 256-Gaussian batch from a shared-memory bit matrix).
#include "../../include/voge_b200.h"
#include "common.cuh"

namespace voge {

constexpr int kCoarseNT = 256;
constexpr int kCoarseChunk = 2048;
constexpr int kMaxBinsPerDim = 65;  // reference refuses >= 66 bins per side (rasterize_coarse.cu:213)

struct CoarseArgs {
    const float* points;   // (P,3)
    const float* radius;   // (P,2)
    const int64_t* first_idx;
    const int64_t* num_per;
    int B, P, H, W, bin_size, M, BH, BW, nchunks;
    int32_t* chunk_hist;   // (B, nchunks, BH*BW)
    int32_t* bin_points;   // (B, BH, BW, M)
    int32_t* bin_counts;   // (B, BH, BW)
};

// bin edges in shared memory: [ymin(BH) | ymax(BH) | xmin(BW) | xmax(BW)]
__device__ __forceinline__ void load_edges(float* e, const CoarseArgs& a) {
    const float half_x = __fdiv_rn(ndc_range(a.W, a.H) / 2.0f, (float)a.W);
    const float half_y = __fdiv_rn(ndc_range(a.H, a.W) / 2.0f, (float)a.H);
    for (int i = threadIdx.x; i < a.BH; i += blockDim.x) {
        e[i] = __fsub_rn(pix_to_ndc(i * a.bin_size, a.H, a.W), half_y);
        e[a.BH + i] = __fadd_rn(pix_to_ndc((i + 1) * a.bin_size - 1, a.H, a.W), half_y);
    }
    for (int i = threadIdx.x; i < a.BW; i += blockDim.x) {
        e[2 * a.BH + i] = __fsub_rn(pix_to_ndc(i * a.bin_size, a.W, a.H), half_x);
        e[2 * a.BH + a.BW + i] = __fadd_rn(pix_to_ndc((i + 1) * a.bin_size - 1, a.W, a.H), half_x);
    }
}

struct BinRect {
    int y0, y1, x0, x1;  // inclusive; empty if y0 > y1 or x0 > x1
};

// The reference predicate (min <= bin_max) && (bin_min < max) per axis; edges are monotone so the
// accepted bins form a contiguous range.
__device__ __forceinline__ BinRect bin_rect(const float* e, const CoarseArgs& a, int64_t p) {
    BinRect r{1, 0, 1, 0};
    const float x = a.points[p * 3 + 0], y = a.points[p * 3 + 1], z = a.points[p * 3 + 2];
    if (z < 0.f) return r;   // rasterize_coarse.cu:35
    const float rx = a.radius[p * 2 + 0], ry = a.radius[p * 2 + 1];
    const float xmin = __fsub_rn(x, rx), xmax = __fadd_rn(x, rx);
    const float ymin = __fsub_rn(y, ry), ymax = __fadd_rn(y, ry);
    int y0 = a.BH, y1 = -1, x0 = a.BW, x1 = -1;
    for (int i = 0; i < a.BH; ++i)
        if ((ymin <= e[a.BH + i]) && (e[i] < ymax)) { y0 = min(y0, i); y1 = i; }
    for (int i = 0; i < a.BW; ++i)
        if ((xmin <= e[2 * a.BH + a.BW + i]) && (e[2 * a.BH + i] < xmax)) { x0 = min(x0, i); x1 = i; }
    r.y0 = y0; r.y1 = y1; r.x0 = x0; r.x1 = x1;
    return r;
}

__global__ void __launch_bounds__(kCoarseNT) coarse_hist_kernel(const CoarseArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nbins = a.BH * a.BW;
    int* hist = reinterpret_cast<int*>(smem_raw);
    float* edges = reinterpret_cast<float*>(hist + nbins);
    const int chunk = blockIdx.x, b = blockIdx.y;
    for (int i = threadIdx.x; i < nbins; i += kCoarseNT) hist[i] = 0;
    load_edges(edges, a);
    __syncthreads();
    const int64_t start = a.first_idx[b];
    const int64_t count = a.num_per[b];
    for (int i = threadIdx.x; i < kCoarseChunk; i += kCoarseNT) {
        const int64_t local = (int64_t)chunk * kCoarseChunk + i;
        if (local >= count) break;
        const int64_t p = start + local;
        if (p < 0 || p >= a.P) continue;
        const BinRect r = bin_rect(edges, a, p);
        for (int by = r.y0; by <= r.y1; ++by)
            for (int bx = r.x0; bx <= r.x1; ++bx) atomicAdd(&hist[by * a.BW + bx], 1);
    }
    __syncthreads();
    int32_t* out = a.chunk_hist + ((int64_t)b * a.nchunks + chunk) * nbins;
    for (int i = threadIdx.x; i < nbins; i += kCoarseNT) out[i] = hist[i];
}

// exclusive scan over chunks, one thread per (view, bin); coalesced across bins
__global__ void coarse_scan_kernel(const CoarseArgs a) {
    const int nbins = a.BH * a.BW;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)a.B * nbins) return;
    const int b = (int)(t / nbins), bin = (int)(t % nbins);
    int32_t* h = a.chunk_hist + (int64_t)b * a.nchunks * nbins + bin;
    int running = 0;
    for (int c = 0; c < a.nchunks; ++c) {
        const int v = h[(int64_t)c * nbins];
        h[(int64_t)c * nbins] = running;
        running += v;
    }
    a.bin_counts[t] = running;
}

__global__ void __launch_bounds__(kCoarseNT) coarse_fill_kernel(const CoarseArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int WPB = kCoarseNT / 32;  // mask words per bin
    const int nbins = a.BH * a.BW;
    int* running = reinterpret_cast<int*>(smem_raw);                 // [nbins]
    unsigned* mask = reinterpret_cast<unsigned*>(running + nbins);   // [nbins][WPB]
    float* edges = reinterpret_cast<float*>(mask + (size_t)nbins * WPB);
    const int chunk = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x;
    const int32_t* offs = a.chunk_hist + ((int64_t)b * a.nchunks + chunk) * nbins;
    for (int i = tid; i < nbins; i += kCoarseNT) running[i] = offs[i];
    for (int i = tid; i < nbins * WPB; i += kCoarseNT) mask[i] = 0u;
    load_edges(edges, a);
    __syncthreads();
    const int64_t start = a.first_idx[b];
    const int64_t count = a.num_per[b];
    int32_t* out = a.bin_points + (int64_t)b * nbins * a.M;
    for (int sb = 0; sb < kCoarseChunk / kCoarseNT; ++sb) {
        const int64_t local = (int64_t)chunk * kCoarseChunk + sb * kCoarseNT + tid;
        if ((int64_t)chunk * kCoarseChunk + sb * kCoarseNT >= count) break;   // uniform
        BinRect r{1, 0, 1, 0};
        const int64_t p = start + local;
        if (local < count && p >= 0 && p < a.P) r = bin_rect(edges, a, p);
        const int w = tid >> 5;
        const unsigned bit = 1u << (tid & 31);
        for (int by = r.y0; by <= r.y1; ++by)
            for (int bx = r.x0; bx <= r.x1; ++bx) atomicOr(&mask[(by * a.BW + bx) * WPB + w], bit);
        __syncthreads();
        for (int by = r.y0; by <= r.y1; ++by)
            for (int bx = r.x0; bx <= r.x1; ++bx) {
                const int bin = by * a.BW + bx;
                int rank = __popc(mask[bin * WPB + w] & (bit - 1u));
                for (int ww = 0; ww < w; ++ww) rank += __popc(mask[bin * WPB + ww]);
                const int slot = running[bin] + rank;
                if (slot < a.M) out[(int64_t)bin * a.M + slot] = (int32_t)p;
            }
        __syncthreads();
        for (int bin = tid; bin < nbins; bin += kCoarseNT) {
            int tot = 0;
#pragma unroll
            for (int ww = 0; ww < WPB; ++ww) {
                const unsigned m = mask[bin * WPB + ww];
                if (m) { tot += __popc(m); mask[bin * WPB + ww] = 0u; }
            }
            running[bin] += tot;
        }
        __syncthreads();
    }
}

}  // namespace voge

extern "C" int64_t voge_rasterize_coarse_scratch_elems(int B, int max_per_cloud, int H, int W, int bin_size) {
    using namespace voge;
    if (bin_size <= 0) return 0;
    const int BH = cdiv(H, bin_size), BW = cdiv(W, bin_size);
    const int nchunks = max(1, cdiv(max_per_cloud, kCoarseChunk));
    return (int64_t)B * nchunks * BH * BW;
}

extern "C" int voge_rasterize_coarse(const float* points_ndc, const float* radius, const int64_t* first_idx,
                                     const int64_t* num_per, int B, int P, int max_per_cloud, int H, int W,
                                     int bin_size, int M, int32_t* scratch, int32_t* bin_points,
                                     int32_t* bin_counts, voge_stream_t stream) {
    using namespace voge;
    if (B <= 0 || H <= 0 || W <= 0 || bin_size <= 0) return (int)cudaErrorInvalidValue;
    CoarseArgs a;
    a.points = points_ndc; a.radius = radius; a.first_idx = first_idx; a.num_per = num_per;
    a.B = B; a.P = P; a.H = H; a.W = W; a.bin_size = bin_size; a.M = M;
    a.BH = cdiv(H, bin_size); a.BW = cdiv(W, bin_size);
    if (a.BH > kMaxBinsPerDim || a.BW > kMaxBinsPerDim) return (int)cudaErrorInvalidValue;
    a.nchunks = max(1, cdiv(max_per_cloud, kCoarseChunk));
    a.chunk_hist = scratch; a.bin_points = bin_points; a.bin_counts = bin_counts;
    const int nbins = a.BH * a.BW;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t edge_bytes = (size_t)(2 * a.BH + 2 * a.BW) * 4;
    const size_t smem_hist = (size_t)nbins * 4 + edge_bytes;
    const size_t smem_fill = (size_t)nbins * 4 + (size_t)nbins * (kCoarseNT / 32) * 4 + edge_bytes;
    VOGE_CUDA_TRY(cudaFuncSetAttribute(coarse_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_hist));
    VOGE_CUDA_TRY(cudaFuncSetAttribute(coarse_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fill));
    dim3 grid(a.nchunks, B);
    coarse_hist_kernel<<<grid, kCoarseNT, smem_hist, s>>>(a);
    VOGE_LAUNCH_CHECK();
    const int64_t nb = (int64_t)B * nbins;
    coarse_scan_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, s>>>(a);
    VOGE_LAUNCH_CHECK();
    if (M > 0) {
        coarse_fill_kernel<<<grid, kCoarseNT, smem_fill, s>>>(a);
        VOGE_LAUNCH_CHECK();
    }
    return 0;
}

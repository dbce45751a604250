/*
 * voge_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp32, libm fmaf) of the reference's CUDA kernels on the VoGE
 * ray-tracing hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (voge_b200/) never does.
 *
 * Parity pinning: the reference ships no test-suite; its only known-answer vector is the
 * comment block ray_trace_voge.cu:381-448 (checked in tests/test_oracle.py).  Beyond that the
 * oracle is pinned against outputs of the UNMODIFIED reference kernels (oracle/_ref, built by
 * oracle/build_ref.py) run on a B200 and committed under tests/golden/ (see tools/make_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Rounding contract.  The reference kernels are compiled by nvcc with its default
 * -fmad=true, so `Innerdot3d` (ray_trace_voge.cu:11-38) is not evaluated as written but as the
 * contracted sequence nvcc emits for sm_100a (read from the SASS of oracle/_ref):
 *     t_ij = rn(a_i * b_ij)
 *     acc  = fma(t_11, c_1, rn(t_12 * c_2))
 *     acc  = fma(t_13, c_3, acc); acc = fma(t_21, c_1, acc); ... ; acc = fma(t_33, c_3, acc)
 * hit_length = rn(msk / ksk); hit_activation = rn(msm - rn(rn(msk*msk) / ksk)).
 * This file must be compiled with -ffp-contract=off so that only the explicit fmaf() fuse.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define VO_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Innerdot3d, ray_trace_voge.cu:11-38, in nvcc's contracted evaluation order (see header). */
static inline float vo_innerdot3d(const float* a, const float* b, const float* c) {
    const float t11 = a[0] * b[0], t12 = a[0] * b[1], t13 = a[0] * b[2];
    const float t21 = a[1] * b[3], t22 = a[1] * b[4], t23 = a[1] * b[5];
    const float t31 = a[2] * b[6], t32 = a[2] * b[7], t33 = a[2] * b[8];
    float acc = fmaf(t11, c[0], t12 * c[1]);
    acc = fmaf(t13, c[2], acc);
    acc = fmaf(t21, c[0], acc);
    acc = fmaf(t22, c[1], acc);
    acc = fmaf(t23, c[2], acc);
    acc = fmaf(t31, c[0], acc);
    acc = fmaf(t32, c[1], acc);
    acc = fmaf(t33, c[2], acc);
    return acc;
}

VO_EXPORT float vo_innerdot3d_export(const float* a, const float* b, const float* c) {
    return vo_innerdot3d(a, b, c);
}

/* The three quadratic forms and the derived hit quantities for one (ray, Gaussian) pair,
 * ray_trace_voge.cu:188-193. */
VO_EXPORT void vo_pair(const float* mu, const float* S, const float* d, float* len, float* act,
                       float* dsd) {
    const float ksk = vo_innerdot3d(d, S, d);
    const float msk = vo_innerdot3d(mu, S, d);
    const float msm = vo_innerdot3d(mu, S, mu);
    *len = msk / ksk;
    const float sq = msk * msk;
    *act = msm - sq / ksk;
    *dsd = ksk;
}

/* ------------------------------------------------------------------------------------------
 * PixToNonSquareNdc / NonSquareNdcRange, rasterize_points/rasterization_utils.cuh:16-42. */
static inline float vo_ndc_range(int S1, int S2) {
    float range = 2.0f;
    if (S1 > S2) range = ((float)S1 * range) / (float)S2;
    return range;
}
static inline float vo_pix_to_ndc(int i, int S1, int S2) {
    const float range = vo_ndc_range(S1, S2);
    const float offset = range / 2.0f;
    /* -offset + (range * i + offset) / S1 ; nvcc contracts range*i+offset into one fma */
    return -offset + fmaf(range, (float)i, offset) / (float)S1;
}

/* Coarse binning: EllipseBoundingBoxKernel rasterize_coarse.cu:20-42 and the overlap
 * predicate of RasterizeCoarseCudaKernel :105-135.  Deterministic restatement: indices in
 * ascending order, the first M kept, true counts returned (the reference's order inside a bin
 * depends on an atomic reservation, :153, and overflowing chunks are dropped, :154-170). */
VO_EXPORT void vo_rasterize_coarse(const float* points, const float* radius,
                                   const int64_t* first_idx, const int64_t* num_per, int B, int P,
                                   int H, int W, int bin_size, int M, int32_t* bin_points,
                                   int32_t* bin_counts) {
    const int BW = 1 + (W - 1) / bin_size;
    const int BH = 1 + (H - 1) / bin_size;
    const float half_pix_x = (vo_ndc_range(W, H) / 2.0f) / (float)W;
    const float half_pix_y = (vo_ndc_range(H, W) / 2.0f) / (float)H;
    const int64_t total = (int64_t)B * BH * BW;
    for (int64_t i = 0; i < total * M; ++i) bin_points[i] = -1;
    for (int64_t i = 0; i < total; ++i) bin_counts[i] = 0;
    float* ymin_b = (float*)malloc(sizeof(float) * BH);
    float* ymax_b = (float*)malloc(sizeof(float) * BH);
    float* xmin_b = (float*)malloc(sizeof(float) * BW);
    float* xmax_b = (float*)malloc(sizeof(float) * BW);
    for (int by = 0; by < BH; ++by) {
        ymin_b[by] = vo_pix_to_ndc(by * bin_size, H, W) - half_pix_y;
        ymax_b[by] = vo_pix_to_ndc((by + 1) * bin_size - 1, H, W) + half_pix_y;
    }
    for (int bx = 0; bx < BW; ++bx) {
        xmin_b[bx] = vo_pix_to_ndc(bx * bin_size, W, H) - half_pix_x;
        xmax_b[bx] = vo_pix_to_ndc((bx + 1) * bin_size - 1, W, H) + half_pix_x;
    }
    for (int b = 0; b < B; ++b) {
        const int64_t start = first_idx[b], stop = first_idx[b] + num_per[b];
        for (int64_t p = start; p < stop && p < P; ++p) {
            const float x = points[p * 3 + 0], y = points[p * 3 + 1], z = points[p * 3 + 2];
            if (z < 0) continue; /* rasterize_coarse.cu:35 */
            const float rx = radius[p * 2 + 0], ry = radius[p * 2 + 1];
            const float xmin = x - rx, xmax = x + rx, ymin = y - ry, ymax = y + ry;
            for (int by = 0; by < BH; ++by) {
                if (!((ymin <= ymax_b[by]) && (ymin_b[by] < ymax))) continue;
                for (int bx = 0; bx < BW; ++bx) {
                    if (!((xmin <= xmax_b[bx]) && (xmin_b[bx] < xmax))) continue;
                    const int64_t bin = ((int64_t)b * BH + by) * BW + bx;
                    const int c = bin_counts[bin]++;
                    if (c < M) bin_points[bin * M + c] = (int32_t)p;
                }
            }
        }
    }
    free(ymin_b); free(ymax_b); free(xmin_b); free(xmax_b);
}

/* ------------------------------------------------------------------------------------------
 * RayTraceFineVogeKernel, ray_trace_voge.cu:135-217: per pixel loop over the bin's candidate
 * list, threshold on hit_activation, insert-at-current_ptr-then-bubble top-K (:197-213).
 * Indexing of bin_points uses the intended (BH*BW) batch stride; the reference's BH*BH (:185)
 * only differs for B > 1 with non-square bin grids. */
VO_EXPORT void vo_ray_trace_fine(const float* mus, const float* isigmas, const float* rays,
                                 const int32_t* bin_points, float thr_act, int bin_size, int B,
                                 int H, int W, int BH, int BW, int M, int K, int32_t* out_idx,
                                 float* out_len, float* out_act, float* out_dsd) {
    const int64_t R = (int64_t)B * H * W;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < R; ++r) {
        const int bi = (int)(r / ((int64_t)H * W));
        const int yi = (int)((r / W) % H), xi = (int)(r % W);
        const int by = yi / bin_size, bx = xi / bin_size;
        int32_t* pidx = out_idx + r * K;
        float* plen = out_len + r * K;
        float* pact = out_act + r * K;
        float* pdsd = out_dsd + r * K;
        for (int k = 0; k < K; ++k) { pidx[k] = -1; plen[k] = 1e10f; pact[k] = 1e10f; pdsd[k] = 0.f; }
        const float* d = rays + r * 3;
        const int32_t* list = bin_points + (((int64_t)bi * BH + by) * BW + bx) * M;
        int cur = 0;
        for (int m = 0; m < M; ++m) {
            const int32_t g = list[m];
            if (g <= -1) continue;
            float len, act, dsd;
            vo_pair(mus + (int64_t)g * 3, isigmas + (int64_t)g * 9, d, &len, &act, &dsd);
            if (act < thr_act && len < plen[cur]) {
                plen[cur] = len; pact[cur] = act; pdsd[cur] = dsd; pidx[cur] = g;
                for (int t = cur; t > 0 && plen[t] < plen[t - 1]; --t) {
                    float f; int32_t i;
                    f = plen[t]; plen[t] = plen[t - 1]; plen[t - 1] = f;
                    f = pact[t]; pact[t] = pact[t - 1]; pact[t - 1] = f;
                    f = pdsd[t]; pdsd[t] = pdsd[t - 1]; pdsd[t - 1] = f;
                    i = pidx[t]; pidx[t] = pidx[t - 1]; pidx[t - 1] = i;
                }
                if (cur < K - 1) cur++;
            }
        }
    }
}

/* Number of (ray, candidate) pairs the reference evaluates = sum over in-image pixels of the
 * valid entries of their bin (the N_pairs of SURVEY.md 8d). */
VO_EXPORT int64_t vo_count_pairs(const int32_t* bin_points, int bin_size, int B, int H, int W,
                                 int BH, int BW, int M) {
    int64_t total = 0;
    for (int b = 0; b < B; ++b)
        for (int by = 0; by < BH; ++by)
            for (int bx = 0; bx < BW; ++bx) {
                const int32_t* list = bin_points + (((int64_t)b * BH + by) * BW + bx) * M;
                int64_t c = 0;
                for (int m = 0; m < M; ++m) c += list[m] > -1;
                int ph = H - by * bin_size; if (ph > bin_size) ph = bin_size;
                int pw = W - bx * bin_size; if (pw > bin_size) pw = bin_size;
                total += c * ph * pw;
            }
    return total;
}

/* ------------------------------------------------------------------------------------------
 * RayTraceFineVogeBackwardKernel ray_trace_voge.cu:283-332 with Innerdot3dBackward :41-91.
 * The reference accumulates with fp32 atomics in a non-deterministic order; the oracle
 * accumulates the same per-hit fp32 contributions into fp64 sums and rounds once, i.e. it is
 * the order-independent target both implementations approximate. */
static void vo_dot3d_backward(float g, const float* a, const float* b, const float* c, double* ga,
                              double* gb, double* gc) {
    ga[0] += (double)((b[0] * c[0] + b[1] * c[1] + b[2] * c[2]) * g);
    ga[1] += (double)((b[3] * c[0] + b[4] * c[1] + b[5] * c[2]) * g);
    ga[2] += (double)((b[6] * c[0] + b[7] * c[1] + b[8] * c[2]) * g);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) gb[i * 3 + j] += (double)((a[i] * c[j]) * g);
    gc[0] += (double)((b[0] * a[0] + b[3] * a[1] + b[6] * a[2]) * g);
    gc[1] += (double)((b[1] * a[0] + b[4] * a[1] + b[7] * a[2]) * g);
    gc[2] += (double)((b[2] * a[0] + b[5] * a[1] + b[8] * a[2]) * g);
}

VO_EXPORT void vo_ray_trace_fine_backward(const float* mus, const float* isigmas, const float* rays,
                                          const int32_t* idx, const float* g_len, const float* g_act,
                                          const float* g_dsd, int B, int H, int W, int K, int P,
                                          float* grad_rays, float* grad_mus, float* grad_isg) {
    const int64_t R = (int64_t)B * H * W;
    double* gr = (double*)calloc((size_t)R * 3, sizeof(double));
    double* gm = (double*)calloc((size_t)P * 3, sizeof(double));
    double* gs = (double*)calloc((size_t)P * 9, sizeof(double));
    for (int64_t pid = 0; pid < R * K; ++pid) {
        const int32_t g = idx[pid];
        if (g == -1) continue;
        const int64_t r = pid / K;
        const float gl = g_len[pid], ga = g_act[pid], gd = g_dsd[pid];
        const float* d = rays + r * 3;
        const float* mu = mus + (int64_t)g * 3;
        const float* S = isigmas + (int64_t)g * 9;
        const float ksk = vo_innerdot3d(d, S, d);
        const float msk = vo_innerdot3d(mu, S, d);
        const float g_ksk = (ga * msk - gl) * msk / (ksk * ksk) + gd; /* :324 */
        const float g_msk = (gl - 2 * ga * msk) / ksk;                 /* :325 */
        const float g_msm = ga;                                        /* :326 */
        vo_dot3d_backward(g_ksk, d, S, d, gr + r * 3, gs + (int64_t)g * 9, gr + r * 3);
        vo_dot3d_backward(g_msk, mu, S, d, gm + (int64_t)g * 3, gs + (int64_t)g * 9, gr + r * 3);
        vo_dot3d_backward(g_msm, mu, S, mu, gm + (int64_t)g * 3, gs + (int64_t)g * 9, gm + (int64_t)g * 3);
    }
    for (int64_t i = 0; i < R * 3; ++i) grad_rays[i] = (float)gr[i];
    for (int64_t i = 0; i < (int64_t)P * 3; ++i) grad_mus[i] = (float)gm[i];
    for (int64_t i = 0; i < (int64_t)P * 9; ++i) grad_isg[i] = (float)gs[i];
    free(gr); free(gm); free(gs);
}

/* ------------------------------------------------------------------------------------------
 * SampleVogeKernel sample_voge.cu:35-66 (fp64 accumulation, see backward note above). */
VO_EXPORT void vo_sample(const float* image, const float* weight, const int32_t* idx, int64_t R,
                         int K, int C, int num_vert, float* feat, float* wsum) {
    double* f = (double*)calloc((size_t)num_vert * C, sizeof(double));
    double* s = (double*)calloc((size_t)num_vert, sizeof(double));
    for (int64_t pid = 0; pid < R * K; ++pid) {
        const int32_t g = idx[pid];
        if (g == -1) continue;
        const float w = weight[pid];
        const float* px = image + (pid / K) * C;
        for (int c = 0; c < C; ++c) f[(int64_t)g * C + c] += (double)(px[c] * w);
        s[g] += (double)w;
    }
    for (int64_t i = 0; i < (int64_t)num_vert * C; ++i) feat[i] = (float)f[i];
    for (int64_t i = 0; i < num_vert; ++i) wsum[i] = (float)s[i];
    free(f); free(s);
}

/* SampleVogeBackwardKernel sample_voge.cu:173-209. */
VO_EXPORT void vo_sample_backward(const float* image, const float* weight, const int32_t* idx,
                                  const float* g_feat, const float* g_wsum, int64_t R, int K, int C,
                                  float* grad_image, float* grad_weight) {
    double* gi = (double*)calloc((size_t)R * C, sizeof(double));
    for (int64_t pid = 0; pid < R * K; ++pid) {
        grad_weight[pid] = 0.f;
        const int32_t g = idx[pid];
        if (g == -1) continue;
        const float w = weight[pid];
        const int64_t r = pid / K;
        for (int c = 0; c < C; ++c) gi[r * C + c] += (double)(g_feat[(int64_t)g * C + c] * w);
        float sum = g_wsum[g];
        for (int c = 0; c < C; ++c) sum += g_feat[(int64_t)g * C + c] * image[r * C + c];
        grad_weight[pid] = sum;
    }
    for (int64_t i = 0; i < R * C; ++i) grad_image[i] = (float)gi[i];
    free(gi);
}

/* ScatterMaxKernel sample_voge.cu:69-92 (output starts from zeros, :156). */
VO_EXPORT void vo_scatter_max(const float* weight, const int32_t* idx, int64_t R, int K,
                              int num_vert, float* wmax) {
    for (int i = 0; i < num_vert; ++i) wmax[i] = 0.f;
    for (int64_t pid = 0; pid < R * K; ++pid) {
        const int32_t g = idx[pid];
        if (g == -1) continue;
        wmax[g] = fmaxf(wmax[g], weight[pid]);
    }
}

VO_EXPORT int vo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Dense-ray API: RayTraceVogeRayKernel voge_ray_tracing_ray.cu:114-143 (same Innerdot3d, hence
 * the same contracted rounding sequence) and FindNearestKKernel :191-239 (same insert-then-bubble
 * top-K; padding idx -1, len 1e10, act 0, dsd 0 from the host wrapper :344-347). */
VO_EXPORT void vo_ray_trace_ray(const float* mus, const float* isigmas, const float* rays, int M, int N,
                                float* out_len, float* out_act, float* out_dsd) {
    for (int r = 0; r < N; ++r)
        for (int p = 0; p < M; ++p)
            vo_pair(mus + (int64_t)p * 3, isigmas + (int64_t)p * 9, rays + (int64_t)r * 3,
                    out_len + (int64_t)r * M + p, out_act + (int64_t)r * M + p, out_dsd + (int64_t)r * M + p);
}

VO_EXPORT void vo_find_nearest_k(const float* len_in, const float* act_in, const float* dsd_in, float thr_act,
                                 int M, int K, int N, int32_t* out_idx, float* out_len, float* out_act,
                                 float* out_dsd) {
    for (int r = 0; r < N; ++r) {
        int32_t* pidx = out_idx + (int64_t)r * K;
        float* plen = out_len + (int64_t)r * K;
        float* pact = out_act + (int64_t)r * K;
        float* pdsd = out_dsd + (int64_t)r * K;
        for (int k = 0; k < K; ++k) { pidx[k] = -1; plen[k] = 1e10f; pact[k] = 0.f; pdsd[k] = 0.f; }
        int cur = 0;
        for (int m = 0; m < M; ++m) {
            const float len = len_in[(int64_t)r * M + m], act = act_in[(int64_t)r * M + m];
            if (act < thr_act && len < plen[cur]) {
                plen[cur] = len; pact[cur] = act; pdsd[cur] = dsd_in[(int64_t)r * M + m]; pidx[cur] = m;
                for (int t = cur; t > 0 && plen[t] < plen[t - 1]; --t) {
                    float f; int32_t i;
                    f = plen[t]; plen[t] = plen[t - 1]; plen[t - 1] = f;
                    f = pact[t]; pact[t] = pact[t - 1]; pact[t - 1] = f;
                    f = pdsd[t]; pdsd[t] = pdsd[t - 1]; pdsd[t - 1] = f;
                    i = pidx[t]; pidx[t] = pidx[t - 1]; pidx[t - 1] = i;
                }
                if (cur < K - 1) cur++;
            }
        }
    }
}

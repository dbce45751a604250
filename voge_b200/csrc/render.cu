// render.cu -- the fused renderer path behind GaussianRenderer.forward:
//   voge_bin_count / voge_bin_fill : per-view screen-space culling into per-tile CSR lists
//   voge_render_forward            : tile-staged filter -> bit-faithful refine -> top-K -> blend weights
//   voge_render_backward           : recompute + chain rule straight into (N,.) parameter gradients
//
// Replaces, for the renderer's own call pattern, the chain
//   rasterize_coarse (PyTorch maths RayTracing.py:42-57 + kernels rasterize_coarse.cu:20-188)
//   -> RayTraceFineVogeKernel (ray_trace_voge.cu:135-217) -> aggregation (Aggregation.py:82-107)
// without materialising (B,N,3)/(B,N,3,3) per-view copies, the (B,BH,BW,M) bin table or the
// (R,K,K) blend tensors.  Candidate semantics are the reference's: a Gaussian is a candidate of a
// pixel iff the reference's bbox test puts it in the pixel's coarse bin (bin_size pixels); the
// tile lists here are that set AND-ed with a provably conservative projected-ellipsoid bound, so
// the fragments equal the unfused path's bit for bit (same exact_pair arithmetic decides).
#include "../../include/voge_b200.h"
#include "blend_core.cuh"
#include "fine_core.cuh"
#include "render_core.cuh"
#include <type_traits>

namespace voge {

// ---- binning ---------------------------------------------------------------------------------------
struct BinArgs {
    const float* verts;
    const float* sigmas;
    int kind;
    const float* Rm;         // (B,3,3) row-vector convention X_view = X_world @ R + T
    const float* Tv;         // (B,3)
    const float* origins;    // (B,3) ray origins (camera centres) used for mu' = verts - origin
    const float* focal;      // (B,2) pixels
    const float* principal;  // (B,2) pixels
    int B, N, H, W;
    float neg_log_thr, thr_act;
    int use_ref_bins, bin_size, BH, BW, tile, TX, TY;
    uint2* rects;            // (B,N): x = x0 | x1 << 16, y = y0 | y1 << 16 in PIXELS (inclusive); empty if x0 > x1
    int32_t* tile_counts;    // (B, TY*TX)
    int32_t* tile_items;     // optional (B, TY*TX): sum over the tile's entries of the rectangle area inside the tile
};

__device__ __forceinline__ float edge_min(int i, int bin, int S1, int S2, float half) {
    return __fsub_rn(pix_to_ndc(i * bin, S1, S2), half);
}
__device__ __forceinline__ float edge_max(int i, int bin, int S1, int S2, float half) {
    return __fadd_rn(pix_to_ndc((i + 1) * bin - 1, S1, S2), half);
}

// contiguous range of reference bins along one axis accepted by the reference predicate
// (lo <= bin_max) && (bin_min < hi)   (rasterize_coarse.cu:116-130)
__device__ __forceinline__ void ref_bin_range(float lo, float hi, int nb, int bin, int S1, int S2, float half,
                                              float scale, int& b0, int& b1) {
    const float c = 0.5f * (float)S1;
    int e0 = (int)floorf(fminf(fmaxf((lo * scale + c) / (float)bin, -1.f), (float)nb));
    int e1 = (int)floorf(fminf(fmaxf((hi * scale + c) / (float)bin, -1.f), (float)nb));
    e0 = min(max(e0, 0), nb - 1);
    e1 = min(max(e1, 0), nb - 1);
    // first bin with lo <= bin_max
    while (e0 > 0 && (lo <= edge_max(e0 - 1, bin, S1, S2, half))) --e0;
    while (e0 < nb && !(lo <= edge_max(e0, bin, S1, S2, half))) ++e0;
    // last bin with bin_min < hi
    while (e1 < nb - 1 && (edge_min(e1 + 1, bin, S1, S2, half) < hi)) ++e1;
    while (e1 >= 0 && !(edge_min(e1, bin, S1, S2, half) < hi)) --e1;
    b0 = e0; b1 = e1;
}

// Per-Gaussian worst-case bound E on |act_reference - act_exact| over all unit rays, act_exact being the
// real-arithmetic value msm - msk^2/ksk of the same fp32 inputs.  A pixel outside the projected ellipsoid
// {act_exact < thr + E} then has act_reference >= thr and can never be a hit, whatever the rounding does.
// Forward error analysis of exact_pair (common.cuh): every term of the three 9-term forms passes through
// at most 10 roundings (one product t_ij, nine partial sums), gamma_10 ~ 10 u, u = 2^-24:
//   |msm - msm*| <= g10 Tmm,  |msk - msk*| <= g10 Um2 |d|,  |ksk - ksk*| <= g10 Us |d|^2,
//   q = rn(rn(msk^2)/ksk):  |q - q*| <= 2 Qn g10 Um2 / l + Qn^2 g10 Us / l^2 + 2.1 u Qn^2 / l   (first order),
//   act = rn(msm - q)  =>  E = u [10 Tmm + 20 Qn Um2 / l + 10 Us Qn^2 / l^2 + 2.1 Qn^2 / l + 2 thr] x 1.25
// with Tmm = sum |mu_i S_ij mu_j|, Um2 = || |S|^T |mu| ||_2, Qn = || S^T mu ||_2, Us >= || |S| ||_2,
// l = lambda_min(sym S); the factor 1.25 covers the second-order terms, which stay below 1 % of the first-order
// ones as long as g10 Us / l < 1e-2 (otherwise the Gaussian is not culled analytically).  This is about half
// the margin of the pixel-major filter (fine_core.cuh), which also has to cover its own re-associated sums.
__device__ __forceinline__ bool gaussian_margin(const float* mu, const float* S, float thr_act, float& margin) {
    const float m0 = mu[0], m1 = mu[1], m2 = mu[2];
    const float q0 = m0 * S[0] + m1 * S[3] + m2 * S[6];
    const float q1 = m0 * S[1] + m1 * S[4] + m2 * S[7];
    const float q2 = m0 * S[2] + m1 * S[5] + m2 * S[8];
    const float a = S[0], b = S[4], c = S[8];
    const float s01 = 0.5f * (S[1] + S[3]), s02 = 0.5f * (S[2] + S[6]), s12 = 0.5f * (S[5] + S[7]);
    // smallest / largest eigenvalue of the symmetric part (closed form, Smith 1961)
    float lmin, lmax;
    {
        const float p1 = s01 * s01 + s02 * s02 + s12 * s12;
        if (p1 == 0.f) {
            lmin = fminf(a, fminf(b, c));
            lmax = fmaxf(a, fmaxf(b, c));
        } else {
            const float qq = (a + b + c) * (1.f / 3.f);
            const float aa = a - qq, bb = b - qq, cc = c - qq;
            const float p = sqrtf((aa * aa + bb * bb + cc * cc + 2.f * p1) * (1.f / 6.f));
            const float ip = 1.f / p;
            const float b00 = aa * ip, b11 = bb * ip, b22 = cc * ip, b01 = s01 * ip, b02 = s02 * ip, b12 = s12 * ip;
            float r = 0.5f * (b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02));
            r = fminf(1.f, fmaxf(-1.f, r));
            const float phi = acosf(r) * (1.f / 3.f);
            lmax = qq + 2.f * p * cosf(phi);
            lmin = qq + 2.f * p * cosf(phi + 2.0943951023931953f);
        }
        lmin -= 1e-5f * fabsf(lmax);   // rounding of the closed form itself
    }
    if (!(lmin > 0.f)) return false;
    const float am0 = fabsf(m0), am1 = fabsf(m1), am2 = fabsf(m2);
    const float c0 = am0 * fabsf(S[0]) + am1 * fabsf(S[3]) + am2 * fabsf(S[6]);   // (|S|^T |mu|)_j
    const float c1 = am0 * fabsf(S[1]) + am1 * fabsf(S[4]) + am2 * fabsf(S[7]);
    const float c2 = am0 * fabsf(S[2]) + am1 * fabsf(S[5]) + am2 * fabsf(S[8]);
    const float Tmm = c0 * am0 + c1 * am1 + c2 * am2;
    const float Um2 = sqrtf(c0 * c0 + c1 * c1 + c2 * c2);
    const float Qn = sqrtf(q0 * q0 + q1 * q1 + q2 * q2);
    const float r0 = fabsf(S[0]) + fabsf(S[1]) + fabsf(S[2]), r1 = fabsf(S[3]) + fabsf(S[4]) + fabsf(S[5]),
                r2 = fabsf(S[6]) + fabsf(S[7]) + fabsf(S[8]);
    const float k0 = fabsf(S[0]) + fabsf(S[3]) + fabsf(S[6]), k1 = fabsf(S[1]) + fabsf(S[4]) + fabsf(S[7]),
                k2 = fabsf(S[2]) + fabsf(S[5]) + fabsf(S[8]);
    const float Us = fmaxf(fmaxf(fmaxf(r0, r1), r2), fmaxf(fmaxf(k0, k1), k2));
    const float il = 1.f / lmin;
    if (!(Us * il < 1.6e4f)) return false;                       // g10 Us / l < 1e-2: second-order terms negligible
    const float bound = 10.f * Tmm + 20.f * Qn * Um2 * il + 10.f * Us * (Qn * il) * (Qn * il) + 2.1f * Qn * Qn * il +
                        2.f * fabsf(thr_act);
    margin = 7.4505806e-8f * 1.003f * bound;                     // 1.25 x 2^-24 (x rays within 1e-3 of unit length)
    return margin >= 0.f && margin < 3.0e38f;
}

__global__ void __launch_bounds__(256) bin_count_kernel(const BinArgs a) {
    const int b = blockIdx.y;
    const float* R = a.Rm + 9 * b;
    const float fx = a.focal[2 * b], fy = a.focal[2 * b + 1];
    const float px = a.principal[2 * b], py = a.principal[2 * b + 1];
    const float c0 = a.origins[3 * b], c1 = a.origins[3 * b + 1], c2 = a.origins[3 * b + 2];
    const float sc = 0.5f * (float)min(a.H, a.W);
    const float half_x = __fdiv_rn(ndc_range(a.W, a.H) / 2.0f, (float)a.W);
    const float half_y = __fdiv_rn(ndc_range(a.H, a.W) / 2.0f, (float)a.H);
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < a.N; g += gridDim.x * blockDim.x) {
        float mu[3], S[9];
        mu[0] = __fsub_rn(__ldg(a.verts + 3 * (int64_t)g), c0);
        mu[1] = __fsub_rn(__ldg(a.verts + 3 * (int64_t)g + 1), c1);
        mu[2] = __fsub_rn(__ldg(a.verts + 3 * (int64_t)g + 2), c2);
        load_S_dyn(a.kind, a.sigmas, g, S);
        // view space (X_v = mu' @ R since the origin is the camera centre)
        const float xv = mu[0] * R[0] + mu[1] * R[3] + mu[2] * R[6];
        const float yv = mu[0] * R[1] + mu[1] * R[4] + mu[2] * R[7];
        const float zv = mu[0] * R[2] + mu[1] * R[5] + mu[2] * R[8];
        // conservative pixel rectangle in which the Gaussian can be a hit (inclusive bounds)
        int x0 = 0, x1 = a.W - 1, y0 = 0, y1 = a.H - 1;
        bool empty = false;
        if (a.use_ref_bins) {
            if (zv < 0.f) empty = true;   // rasterize_coarse.cu:35
            // S_view[:2,:2] = (R^T S R)[:2,:2]
            float SR0[3], SR1[3];   // columns 0,1 of S R
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                SR0[i] = S[3 * i] * R[0] + S[3 * i + 1] * R[3] + S[3 * i + 2] * R[6];
                SR1[i] = S[3 * i] * R[1] + S[3 * i + 1] * R[4] + S[3 * i + 2] * R[7];
            }
            const float v00 = R[0] * SR0[0] + R[3] * SR0[1] + R[6] * SR0[2];
            const float v01 = R[0] * SR1[0] + R[3] * SR1[1] + R[6] * SR1[2];
            const float v10 = R[1] * SR0[0] + R[4] * SR0[1] + R[7] * SR0[2];
            const float v11 = R[1] * SR1[0] + R[4] * SR1[1] + R[7] * SR1[2];
            const float idet = 1.f / (v00 * v11 - v01 * v10);
            const float i00 = v11 * idet, i01 = -v01 * idet, i10 = -v10 * idet, i11 = v00 * idet;
            const float F0 = fx / sc, F1 = fy / sc;
            // radii = sqrt(colsum(-ln(thr) F inv F)) / z   (RayTracing.py:33-39)
            const float col0 = a.neg_log_thr * F0 * (F0 * i00 + F1 * i10);
            const float col1 = a.neg_log_thr * F1 * (F0 * i01 + F1 * i11);
            const float iz = 1.f / zv;
            const float rx = sqrtf(col0) * iz, ry = sqrtf(col1) * iz;
            const float xn = ((px - fx * xv * iz) - 0.5f * (float)a.W) / sc;
            const float yn = ((py - fy * yv * iz) - 0.5f * (float)a.H) / sc;
            int bx0, bx1, by0, by1;
            ref_bin_range(xn - rx, xn + rx, a.BW, a.bin_size, a.W, a.H, half_x, sc, bx0, bx1);
            ref_bin_range(yn - ry, yn + ry, a.BH, a.bin_size, a.H, a.W, half_y, sc, by0, by1);
            if (bx0 > bx1 || by0 > by1) empty = true;
            x0 = bx0 * a.bin_size; x1 = min((bx1 + 1) * a.bin_size - 1, a.W - 1);
            y0 = by0 * a.bin_size; y1 = min((by1 + 1) * a.bin_size - 1, a.H - 1);
        }
        // conservative projected-ellipsoid bound (exact tangent lines of {act < thr + margin})
        // The bound below is the image of the ellipsoid {(x-mu)^T S (x-mu) < thr}: that IS the set
        // {act < thr} only for a symmetric S (msk = mu^T S d uses the antisymmetric part too), so
        // it is applied to (numerically) symmetric S only; fp32-level asymmetry (|S_ij - S_ji| <=
        // 2^-20 (|S_ij|+|S_ji|), e.g. from R D R^T products) is absorbed by doubling the margin.
        float margin;
        const float as01 = fabsf(S[1] - S[3]), as02 = fabsf(S[2] - S[6]), as12 = fabsf(S[5] - S[7]);
        const bool sym_exact = (as01 == 0.f) && (as02 == 0.f) && (as12 == 0.f);
        const bool sym_near = as01 <= 9.5367e-7f * (fabsf(S[1]) + fabsf(S[3])) &&
                              as02 <= 9.5367e-7f * (fabsf(S[2]) + fabsf(S[6])) &&
                              as12 <= 9.5367e-7f * (fabsf(S[5]) + fabsf(S[7]));
        if (!empty && zv > 0.f && sym_near && gaussian_margin(mu, S, a.thr_act, margin)) {
            if (!sym_exact) margin *= 2.f;
            const float a00 = S[0], a11 = S[4], a22 = S[8];
            const float a01 = 0.5f * (S[1] + S[3]), a02 = 0.5f * (S[2] + S[6]), a12 = 0.5f * (S[5] + S[7]);
            // inverse of the symmetric part (adjugate / det)
            const float j00 = a11 * a22 - a12 * a12, j01 = a02 * a12 - a01 * a22, j02 = a01 * a12 - a02 * a11;
            const float j11 = a00 * a22 - a02 * a02, j12 = a01 * a02 - a00 * a12, j22 = a00 * a11 - a01 * a01;
            const float det = a00 * j00 + a01 * j01 + a02 * j02;
            const float t = (a.thr_act + margin) * 1.0001f / det;
            // Q = t * R^T J R ; need Q00 Q11 Q22 Q02 Q12
            float JR[3][3];
            const float J[9] = {j00, j01, j02, j01, j11, j12, j02, j12, j22};
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) JR[i][k] = J[3 * i] * R[k] + J[3 * i + 1] * R[3 + k] + J[3 * i + 2] * R[6 + k];
            const float Q00 = t * (R[0] * JR[0][0] + R[3] * JR[1][0] + R[6] * JR[2][0]);
            const float Q11 = t * (R[1] * JR[0][1] + R[4] * JR[1][1] + R[7] * JR[2][1]);
            const float Q22 = t * (R[2] * JR[0][2] + R[5] * JR[1][2] + R[8] * JR[2][2]);
            const float Q02 = t * (R[0] * JR[0][2] + R[3] * JR[1][2] + R[6] * JR[2][2]);
            const float Q12 = t * (R[1] * JR[0][2] + R[4] * JR[1][2] + R[7] * JR[2][2]);
            const float C22 = zv * zv - Q22;
            if (C22 > 1e-6f * zv * zv && det > 0.f) {
                // disc = C02^2 - C00 C22 expanded to avoid cancellation
                const float dx = xv * xv * Q22 + zv * zv * Q00 - 2.f * xv * zv * Q02 + (Q02 * Q02 - Q00 * Q22);
                const float dy = yv * yv * Q22 + zv * zv * Q11 - 2.f * yv * zv * Q12 + (Q12 * Q12 - Q11 * Q22);
                const float sx = sqrtf(fmaxf(dx, 0.f)) * 1.001f, sy = sqrtf(fmaxf(dy, 0.f)) * 1.001f;
                const float C02 = xv * zv - Q02, C12 = yv * zv - Q12;
                const float ic = 1.f / C22;
                const float u_lo = (C02 - sx) * ic, u_hi = (C02 + sx) * ic;
                const float v_lo = (C12 - sy) * ic, v_hi = (C12 + sy) * ic;
                // pixel centre xi+.5 = px - fx*u  (u = X/Z); slack 0.05 px
                const float xs_lo = px - fx * u_hi - 0.55f, xs_hi = px - fx * u_lo - 0.45f;
                const float ys_lo = py - fy * v_hi - 0.55f, ys_hi = py - fy * v_lo - 0.45f;
                if (xs_lo == xs_lo && xs_hi == xs_hi && ys_lo == ys_lo && ys_hi == ys_hi) {
                    const float fW = (float)a.W, fH = (float)a.H;
                    if (xs_hi < 0.f || ys_hi < 0.f || xs_lo > fW - 1.f || ys_lo > fH - 1.f) {
                        empty = true;
                    } else {
                        const int pxl = (int)ceilf(fmaxf(xs_lo, 0.f)), pxh = (int)floorf(fminf(xs_hi, fW - 1.f));
                        const int pyl = (int)ceilf(fmaxf(ys_lo, 0.f)), pyh = (int)floorf(fminf(ys_hi, fH - 1.f));
                        x0 = max(x0, pxl); x1 = min(x1, pxh);
                        y0 = max(y0, pyl); y1 = min(y1, pyh);
                    }
                }
            }
        }
        if (x0 > x1 || y0 > y1) empty = true;
        uint2 rc;
        if (empty) {
            rc = make_uint2(1u, 0u);
        } else {
            rc = make_uint2((unsigned)x0 | ((unsigned)x1 << 16), (unsigned)y0 | ((unsigned)y1 << 16));
            const int tx0 = x0 / a.tile, tx1 = x1 / a.tile, ty0 = y0 / a.tile, ty1 = y1 / a.tile;
            int32_t* cnt = a.tile_counts + (int64_t)b * a.TX * a.TY * kBinSub + (g & (kBinSub - 1));
            int32_t* itm = a.tile_items != nullptr ? a.tile_items + (int64_t)b * a.TX * a.TY * kBinSub + (g & (kBinSub - 1)) : nullptr;
            for (int ty = ty0; ty <= ty1; ++ty)
                for (int tx = tx0; tx <= tx1; ++tx) {
                    atomicAdd(cnt + (ty * a.TX + tx) * kBinSub, 1);
                    if (itm != nullptr) atomicAdd(itm + (ty * a.TX + tx) * kBinSub, rect_area_in_tile(rc, tx, ty, a.tile));
                }
        }
        a.rects[(int64_t)b * a.N + g] = rc;
    }
}

__global__ void __launch_bounds__(256) bin_fill_kernel(const uint2* __restrict__ rects,
                                                       const int64_t* __restrict__ tile_offsets,
                                                       int32_t* __restrict__ cursor, int B, int N, int TX, int TY,
                                                       int tile, int32_t* __restrict__ tile_list) {
    const int b = blockIdx.y;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < N; g += gridDim.x * blockDim.x) {
        const uint2 rc = rects[(int64_t)b * N + g];
        if ((rc.x & 0xffff) > (rc.x >> 16)) continue;
        const int tx0 = (rc.x & 0xffff) / tile, tx1 = (rc.x >> 16) / tile, ty0 = (rc.y & 0xffff) / tile, ty1 = (rc.y >> 16) / tile;
        for (int ty = ty0; ty <= ty1; ++ty)
            for (int tx = tx0; tx <= tx1; ++tx) {
                const int64_t t = (((int64_t)b * TY + ty) * TX + tx) * kBinSub + (g & (kBinSub - 1));
                const int slot = atomicAdd(cursor + t, 1);
                tile_list[tile_offsets[t] + slot] = g;
            }
    }
}

// ---- fused forward -----------------------------------------------------------------------------------
struct RenderArgs {
    const float* verts;
    const float* sigmas;
    const float* origins;   // (B,3)
    const float* rays;      // (B,H,W,3)
    const int64_t* tile_offsets;  // (B*TY*TX + 1)
    const int32_t* tile_list;     // local Gaussian indices
    const uint2* rects;           // (B,N) conservative pixel rectangles from bin_count_kernel
    float thr_act, omega;
    int B, N, H, W, K, tile, TX, TY, cap;
    int32_t* out_idx;       // (B,H,W,K) packed b*N+g, -1 padded
    float* out_weight;      // (B,H,W,K)
    float* out_len;         // (B,H,W,K), 1e10 padded
    int64_t* out_valid;     // (B,H,W)
    float* out_act;         // optional (B,H,W,K)
    float* out_dsd;         // optional (B,H,W,K)
    unsigned long long* stats;  // optional: [0] pairs evaluated, [1] pairs refined (= [0]), [2] pixels that overflowed
};

// Gaussian-major ("splat") forward.  The Gaussians of the C5-like scenes cover a few dozen pixels
// each while a 16x16 tile sees several hundred candidates, so iterating pixels x candidates evaluates
// ~10x more pairs than there are near-hits (profiles/ncu_r1_fwd_bwd_v3.md: 460M filtered vs 44M refined
// pairs per view).  Here every warp takes candidates of the tile list in turn and visits only the
// pixels of the candidate's conservative pixel rectangle (reference bin rectangle AND tangent bound of
// {act < thr + margin}, computed by bin_count_kernel) clipped to the tile, evaluates the pair with the
// bit-faithful arithmetic (exact_pair) and appends hits to the pixel's unsorted key buffer in shared
// memory (one shared-memory atomic per hit).  Each pixel thread then sorts its buffer, keeps the K
// smallest (len, idx) keys and runs the blend epilogue.  Results are independent of the order in
// which warps append (keys are unique and totally ordered) => bit-identical to the pixel-major path.
constexpr int kSplatChunk = 128;   // candidates between two per-pixel compactions

template <int NT, int KIND>
__global__ void __launch_bounds__(NT, (NT == 256 ? 3 : 4)) render_fwd_kernel(const RenderArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(smem_raw);           // [cap][NT]
    float* s_ray = reinterpret_cast<float*>(s_key + (size_t)a.cap * NT);                   // [3][NT]
    int* s_cnt = reinterpret_cast<int*>(s_ray + 3 * NT);                                   // [NT]
    unsigned long long* s_lim = reinterpret_cast<unsigned long long*>(s_cnt + NT);         // [NT] admission limit per pixel
    float* s_E = reinterpret_cast<float*>(s_key + (size_t)a.K * NT);                       // epilogue alias, [K][NT]

    const int tid = threadIdx.x;
    int blk = blockIdx.x;
    const int tx = blk % a.TX; blk /= a.TX;
    const int ty = blk % a.TY;
    const int b = blk / a.TY;

    // pixel owned by this thread (inverse of pix_to_col)
    int lx, ly;
    bool in_tile;
    if (a.tile == 16 && NT == 256) {
        const int w = tid >> 5, l = tid & 31;
        lx = (w & 1) * 8 + (l & 7);
        ly = (w >> 1) * 4 + (l >> 3);
        in_tile = true;
    } else {
        lx = tid % a.tile; ly = tid / a.tile;
        in_tile = tid < a.tile * a.tile;
    }
    const int xi = tx * a.tile + lx, yi = ty * a.tile + ly;
    const bool live = in_tile && xi < a.W && yi < a.H;
    const int64_t ray = ((int64_t)b * a.H + yi) * a.W + xi;
    const float c0 = a.origins[3 * b], c1 = a.origins[3 * b + 1], c2 = a.origins[3 * b + 2];
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    if (live) { r0 = a.rays[ray * 3 + 0]; r1 = a.rays[ray * 3 + 1]; r2 = a.rays[ray * 3 + 2]; }
    s_ray[tid] = r0; s_ray[NT + tid] = r1; s_ray[2 * NT + tid] = r2;
    s_cnt[tid] = 0;
    s_lim[tid] = pack_key(kEmptyLen, 0);

    const int64_t tile_id = ((int64_t)b * a.TY + ty) * a.TX + tx;
    const int64_t beg = a.tile_offsets[tile_id * kBinSub];
    const int n = (int)(a.tile_offsets[(tile_id + 1) * kBinSub] - beg);
    const int32_t* list = a.tile_list + beg;
    __syncthreads();

    // ---- phase A: Gaussian-major exact evaluation over each candidate's pixel rectangle, in chunks of
    // kSplatChunk candidates; after every chunk each pixel thread compacts its buffer to the K smallest
    // keys and publishes its K-th key as the admission limit for the following chunks ----
    const int warp = tid >> 5, lane = tid & 31;
    const int px0 = tx * a.tile, py0 = ty * a.tile;
    const int pxe = min(px0 + a.tile, a.W) - 1, pye = min(py0 + a.tile, a.H) - 1;
    unsigned n_eval = 0;
    bool overflow = false;
    // Software pipeline over this warp's candidates (i = warp, warp + NT/32, ...): the index is fetched two
    // items ahead and the Gaussian's record (rectangle, mean, S) one item ahead, so the two dependent L2
    // round trips overlap with the evaluation of the current candidate.
    constexpr int kStep = NT / 32;
    int g_nn = (warp + kStep < n) ? __ldg(list + warp + kStep) : -1;
    int g_n = (warp < n) ? __ldg(list + warp) : -1;
    uint2 rc_n = make_uint2(1u, 0u);
    float v_n[3] = {0.f, 0.f, 0.f};
    float S_n[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) S_n[q] = 0.f;
    if (g_n >= 0) {
        rc_n = a.rects[(int64_t)b * a.N + g_n];
#pragma unroll
        for (int q = 0; q < 3; ++q) v_n[q] = __ldg(a.verts + 3 * (int64_t)g_n + q);
        load_S<KIND>(a.sigmas, g_n, S_n);
    }
    for (int base = 0; base < n; base += kSplatChunk) {
        const int end = min(n, base + kSplatChunk);
        for (int i = base + warp; i < end; i += kStep) {
            const int g = g_n;
            const uint2 rc = rc_n;
            const float m0 = __fsub_rn(v_n[0], c0), m1 = __fsub_rn(v_n[1], c1), m2 = __fsub_rn(v_n[2], c2);
            float S[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) S[q] = S_n[q];
            // prefetch the next record / the index after it
            g_n = g_nn;
            g_nn = (i + 2 * kStep < n) ? __ldg(list + i + 2 * kStep) : -1;
            if (g_n >= 0) {
                rc_n = a.rects[(int64_t)b * a.N + g_n];
#pragma unroll
                for (int q = 0; q < 3; ++q) v_n[q] = __ldg(a.verts + 3 * (int64_t)g_n + q);
                load_S<KIND>(a.sigmas, g_n, S_n);
            }
            const int xl = max((int)(rc.x & 0xffffu), px0), xh = min((int)(rc.x >> 16), pxe);
            const int yl = max((int)(rc.y & 0xffffu), py0), yh = min((int)(rc.y >> 16), pye);
            const int w = xh - xl + 1, h = yh - yl + 1;
            if (w <= 0 || h <= 0) continue;
            // ray-independent part of exact_pair: t_ij = rn(mu_i S_ij) and msm (shared by msk / msm in the reference)
            Prod9 pm;
            float msm;
            if (KIND == 9) {
                pm = exact_row_products(m0, m1, m2, S);
                msm = exact_contract(pm, m0, m1, m2);
            } else {
                pm.t[0] = __fmul_rn(m0, S[0]); pm.t[4] = __fmul_rn(m1, S[4]); pm.t[8] = __fmul_rn(m2, S[8]);
                msm = __fmaf_rn(pm.t[8], m2, __fmaf_rn(pm.t[4], m1, __fmul_rn(pm.t[0], m0)));
            }
            const int area = w * h;
            const unsigned inv_w = kInvW[w];
            for (int p = lane; p < area; p += 32) {
                const int yy = (int)(((unsigned)p * inv_w) >> 16);
                const int xx = p - yy * w;
                const int col = pix_to_col<NT>(xl + xx - px0, yl + yy - py0, a.tile);
                const float d0 = s_ray[col], d1 = s_ray[NT + col], d2 = s_ray[2 * NT + col];
                float ksk, msk;
                if (KIND == 9) {
                    const Prod9 pd = exact_row_products(d0, d1, d2, S);
                    ksk = exact_contract(pd, d0, d1, d2);
                    msk = exact_contract(pm, d0, d1, d2);
                } else {
                    const float t0 = __fmul_rn(d0, S[0]), t1 = __fmul_rn(d1, S[4]), t2 = __fmul_rn(d2, S[8]);
                    ksk = __fmaf_rn(t2, d2, __fmaf_rn(t1, d1, __fmul_rn(t0, d0)));
                    msk = __fmaf_rn(pm.t[8], d2, __fmaf_rn(pm.t[4], d1, __fmul_rn(pm.t[0], d0)));
                }
                const float len = __fdiv_rn(msk, ksk);
                const float act = __fsub_rn(msm, __fdiv_rn(__fmul_rn(msk, msk), ksk));
                ++n_eval;
                if (act < a.thr_act && len == len) {
                    const unsigned long long key = pack_key(len, g);
                    if (key < s_lim[col]) {     // initially len < 1e10 (reference :197), later the pixel's K-th key
                        const int slot = atomicAdd(&s_cnt[col], 1);
                        if (slot < a.cap) s_key[slot * NT + col] = key;
                    }
                }
            }
        }
        __syncthreads();
        // compaction: keep the K smallest keys of this pixel, publish the K-th as the new limit
        int c = s_cnt[tid];
        if (c > a.cap) { overflow = true; c = a.cap; }
        if (c > a.K) {
            while (c > a.K) {           // drop the current maximum
                unsigned long long mx = s_key[tid];
                int mp = 0;
                for (int k = 1; k < c; ++k) {
                    const unsigned long long v = s_key[k * NT + tid];
                    if (v > mx) { mx = v; mp = k; }
                }
                --c;
                s_key[mp * NT + tid] = s_key[c * NT + tid];
            }
            unsigned long long mx = s_key[tid];
            for (int k = 1; k < c; ++k) {
                const unsigned long long v = s_key[k * NT + tid];
                if (v > mx) mx = v;
            }
            s_lim[tid] = mx;
            s_cnt[tid] = c;
        }
        __syncthreads();
    }
    if (a.stats != nullptr) {
        unsigned long long e = n_eval, o = (live && overflow) ? 1ull : 0ull;
        for (int s = 16; s > 0; s >>= 1) {
            e += __shfl_down_sync(0xffffffffu, e, s);
            o += __shfl_down_sync(0xffffffffu, o, s);
        }
        if (lane == 0) { atomicAdd(a.stats, e); atomicAdd(a.stats + 1, e); atomicAdd(a.stats + 2, o); }
    }

    // ---- phase B: per pixel, order the (at most K) survivors ----
    TopKU<NT> top;
    top.init(s_key, a.K, tid);
    if (!live) {
        top.cnt = 0;
    } else if (overflow) {
        // a single chunk delivered more hits than the buffer holds (rare): stream all candidates of the
        // tile through the replace-the-maximum top-K buffer for this pixel alone
        for (int i = 0; i < n; ++i) {
            const int g = __ldg(list + i);
            const Hit h = exact_hit<KIND>(a.verts, a.sigmas, g, c0, c1, c2, r0, r1, r2);
            if (h.act < a.thr_act) top.insert(h.len, g);
        }
        top.sort();
    } else {
        top.cnt = min(s_cnt[tid], a.K);
        top.sort();
    }
    const int cnt = top.cnt;
    // the epilogue's E array aliases key rows K..1.5K of ALL columns: every thread must be done with
    // its unsorted entries before anyone writes there
    __syncthreads();
    if (!live) return;

    // ---- epilogue: exact (len, act, dsd) of the survivors, blend weights, fragment write-out ----
    // Each thread owns a (K,) row of every fragment tensor; rows are written four slots at a time as
    // 16-byte vectors when K % 4 == 0 (a scalar store at a 4K-byte lane stride costs one L1/L2 sector
    // operation per lane and slot).
    int32_t* o_idx = a.out_idx + ray * a.K;
    float* o_len = a.out_len + ray * a.K;
    float* o_w = a.out_weight + ray * a.K;
    const bool vec = (a.K & 3) == 0;
    float2* s_ls = reinterpret_cast<float2*>(s_key);   // the key slots are re-used for (len, sqrt(dsd + 1e-10))
    float s_min = 3.0e38f;
    for (int k0 = 0; k0 < a.K; k0 += 4) {
        int iv[4];
        float lv[4], av[4], dv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            iv[j] = -1; lv[j] = kEmptyLen; av[j] = kEmptyLen; dv[j] = 0.f;
            if (k < cnt) {
                const int g = (int)(unsigned)(s_key[k * NT + tid] & 0xffffffffull);
                const Hit h = exact_hit<KIND>(a.verts, a.sigmas, g, c0, c1, c2, r0, r1, r2);
                iv[j] = b * a.N + g; lv[j] = h.len; av[j] = h.act; dv[j] = h.dsd;
                const float sk = sqrtf(h.dsd + 1e-10f);                              // Aggregation.py:49
                s_ls[k * NT + tid] = make_float2(h.len, sk);
                s_E[k * NT + tid] = expf(-h.act);
                s_min = fminf(s_min, sk);
            }
        }
        if (vec) {
            *reinterpret_cast<int4*>(o_idx + k0) = make_int4(iv[0], iv[1], iv[2], iv[3]);
            *reinterpret_cast<float4*>(o_len + k0) = make_float4(lv[0], lv[1], lv[2], lv[3]);
            if (a.out_act != nullptr) {
                *reinterpret_cast<float4*>(a.out_act + ray * a.K + k0) = make_float4(av[0], av[1], av[2], av[3]);
                *reinterpret_cast<float4*>(a.out_dsd + ray * a.K + k0) = make_float4(dv[0], dv[1], dv[2], dv[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (k0 + j >= a.K) break;
                o_idx[k0 + j] = iv[j]; o_len[k0 + j] = lv[j];
                if (a.out_act != nullptr) { a.out_act[ray * a.K + k0 + j] = av[j]; a.out_dsd[ray * a.K + k0 + j] = dv[j]; }
            }
        }
    }
    // D_m = sum_k E_k Phi((len_m - len_k) s_k).  The list is sorted by len, so outside the window
    // |len_m - len_k| * min_k(s_k) < 4 the erf is saturated: Phi = 1 for k < lo(m) (their E_k are
    // carried in a running prefix sum -- lo(m) only moves forward), Phi = 0 behind the window.
    // The summation order is the plain k = 0..cnt-1 order of the stand-alone aggregation kernel.
    {
        int lo = 0;
        float SE = 0.f;
        for (int m0 = 0; m0 < a.K; m0 += 4) {
            float wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = m0 + j;
                wv[j] = 0.f;
                if (m < cnt) {
                    const float lm = s_ls[m * NT + tid].x;
                    while (lo < m && (lm - s_ls[lo * NT + tid].x) * s_min >= kErfSat) { SE += s_E[lo * NT + tid]; ++lo; }
                    float D = SE;
                    for (int k = lo; k < cnt; ++k) {
                        const float2 lk = s_ls[k * NT + tid];
                        const float dl = lm - lk.x;
                        if (dl * s_min <= -kErfSat) break;
                        D += s_E[k * NT + tid] * phi(dl * lk.y);
                    }
                    const float Em = s_E[m * NT + tid];
                    wv[j] = Em != 0.f ? expf(-(D * a.omega)) * Em * kInvExpMinusHalf : 0.f;
                }
            }
            if (vec) {
                *reinterpret_cast<float4*>(o_w + m0) = make_float4(wv[0], wv[1], wv[2], wv[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (m0 + j < a.K) o_w[m0 + j] = wv[j];
            }
        }
    }
    a.out_valid[ray] = cnt;
}

// per-pixel hit buffer capacity: 2K when it fits comfortably, never below 1.5K (the epilogue aliases
// its E array onto slots K..1.5K)
static int hit_capacity(int K, int nt) {
    int cap = 2 * K;
    // prefer three resident CTAs per SM (<= ~72 KB each) when the minimum depth allows it
    while (cap > (3 * K + 1) / 2 && (size_t)cap * nt * 8 + (size_t)nt * 24 > 73 * 1024) --cap;
    return cap;
}

template <int NT, int KIND>
static int launch_render(const RenderArgs& a0, cudaStream_t stream) {
    RenderArgs a = a0;
    a.cap = hit_capacity(a.K, NT);
    const size_t smem = (size_t)a.cap * NT * 8 + (size_t)NT * 24;
    if (smem > 227 * 1024) return (int)cudaErrorInvalidValue;
    VOGE_CUDA_TRY(cudaFuncSetAttribute(render_fwd_kernel<NT, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (long long)a.B * a.TX * a.TY;
    if (grid <= 0 || grid > 2147483647LL) return (int)cudaErrorInvalidValue;
    render_fwd_kernel<NT, KIND><<<(unsigned)grid, NT, smem, stream>>>(a);
    VOGE_LAUNCH_CHECK();
    return 0;
}

template <int KIND>
static int dispatch_render(const RenderArgs& a, cudaStream_t s) {
    const int px = a.tile * a.tile;
    if (px > 256) return (int)cudaErrorInvalidValue;
    if (px > 128) return launch_render<256, KIND>(a, s);
    if (px > 64) return launch_render<128, KIND>(a, s);
    return launch_render<64, KIND>(a, s);
}

// ---- backward of the geometry: d(len, act, dsd) -> d(verts), d(sigmas) ----------------------------------
struct RenderBwdArgs {
    const float* verts;
    const float* sigmas;
    int kind;
    const float* origins;
    const float* rays;
    const int32_t* idx;      // packed
    const int64_t* valid;    // (B,H,W) number of leading valid slots (idx itself may have been rewritten
                             // -1 -> 0 by merge_final, reference Aggregation.py:131)
    const float* g_len;
    const float* g_act;
    const float* g_dsd;
    int B, N, H, W, K;
    float* grad_verts;       // (N,3) accumulated
    float* grad_sigmas;      // (N,), (N,3) or (N,3,3) accumulated (gradient w.r.t. sigma, i.e. includes the factor 2)
};

__global__ void __launch_bounds__(256) render_bwd_kernel(const RenderBwdArgs a) {
    // 8x4 pixel block per warp so that lanes of a warp touch the same Gaussians
    const int bw = (a.W + 7) / 8, bh = (a.H + 3) / 4;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t per_view = (int64_t)bw * bh;
    if (wid >= per_view * a.B) return;
    const int b = (int)(wid / per_view);
    const int wb = (int)(wid % per_view);
    const int xi = (wb % bw) * 8 + (lane & 7), yi = (wb / bw) * 4 + (lane >> 3);
    if (xi >= a.W || yi >= a.H) return;
    const int64_t r = ((int64_t)b * a.H + yi) * a.W + xi;
    const float d0 = a.rays[r * 3 + 0], d1 = a.rays[r * 3 + 1], d2 = a.rays[r * 3 + 2];
    const float c0 = a.origins[3 * b], c1 = a.origins[3 * b + 1], c2 = a.origins[3 * b + 2];
    const int cnt = (int)min((int64_t)a.K, a.valid[r]);
    for (int k = 0; k < cnt; ++k) {
        const int gp = a.idx[r * a.K + k];
        const int g = gp - b * a.N;
        if (g < 0 || g >= a.N) continue;
        const float gl = a.g_len[r * a.K + k], ga = a.g_act[r * a.K + k], gd = a.g_dsd[r * a.K + k];
        float S[9];
        load_S_dyn(a.kind, a.sigmas, g, S);
        const float m0 = __fsub_rn(__ldg(a.verts + 3 * (int64_t)g), c0);
        const float m1 = __fsub_rn(__ldg(a.verts + 3 * (int64_t)g + 1), c1);
        const float m2 = __fsub_rn(__ldg(a.verts + 3 * (int64_t)g + 2), c2);
        const Prod9 pd = exact_row_products(d0, d1, d2, S);
        const Prod9 pm = exact_row_products(m0, m1, m2, S);
        const float ksk = exact_contract(pd, d0, d1, d2);
        const float msk = exact_contract(pm, d0, d1, d2);
        const float g_ksk = (ga * msk - gl) * msk / (ksk * ksk) + gd;   // ray_trace_voge.cu:324-326
        const float g_msk = (gl - 2.f * ga * msk) / ksk;
        const float g_msm = ga;
        const float Sd0 = S[0] * d0 + S[1] * d1 + S[2] * d2, Sd1 = S[3] * d0 + S[4] * d1 + S[5] * d2,
                    Sd2 = S[6] * d0 + S[7] * d1 + S[8] * d2;
        const float Sm0 = S[0] * m0 + S[1] * m1 + S[2] * m2, Sm1 = S[3] * m0 + S[4] * m1 + S[5] * m2,
                    Sm2 = S[6] * m0 + S[7] * m1 + S[8] * m2;
        const float Stm0 = S[0] * m0 + S[3] * m1 + S[6] * m2, Stm1 = S[1] * m0 + S[4] * m1 + S[7] * m2,
                    Stm2 = S[2] * m0 + S[5] * m1 + S[8] * m2;
        float* gv = a.grad_verts + 3 * (int64_t)g;
        atomicAdd(gv + 0, g_msk * Sd0 + g_msm * (Sm0 + Stm0));
        atomicAdd(gv + 1, g_msk * Sd1 + g_msm * (Sm1 + Stm1));
        atomicAdd(gv + 2, g_msk * Sd2 + g_msm * (Sm2 + Stm2));
        if (a.grad_sigmas != nullptr) {
            const float dv[3] = {d0, d1, d2}, mv[3] = {m0, m1, m2};
            if (a.kind == 1) {
                const float tr = g_ksk * (d0 * d0 + d1 * d1 + d2 * d2) + g_msk * (m0 * d0 + m1 * d1 + m2 * d2) +
                                 g_msm * (m0 * m0 + m1 * m1 + m2 * m2);
                atomicAdd(a.grad_sigmas + g, 2.f * tr);
            } else if (a.kind == 3) {
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    atomicAdd(a.grad_sigmas + 3 * (int64_t)g + i,
                              2.f * (g_ksk * dv[i] * dv[i] + g_msk * mv[i] * dv[i] + g_msm * mv[i] * mv[i]));
            } else {
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        atomicAdd(a.grad_sigmas + 9 * (int64_t)g + 3 * i + j,
                                  2.f * (g_ksk * dv[i] * dv[j] + g_msk * mv[i] * dv[j] + g_msm * mv[i] * mv[j]));
            }
        }
    }
}

// ---- fused backward: d(weight), d(len) -> d(verts), d(sigmas) in ONE kernel ----------------------------
// Recompute, don't store: per pixel the K hits are re-evaluated with the bit-faithful arithmetic
// (so neither act nor dsd is ever written to HBM by the forward), the blend is differentiated
// analytically inside the depth window where the erf is not saturated, and the chain rule of
// ray_trace_voge.cu:324-330 is applied straight into the (N,.) parameter gradients.
struct FusedBwdArgs {
    const float* gauss;      // packed records (voge_pack_gaussians)
    int kind;
    const float* origins;
    const float* rays;
    const int32_t* idx;
    const int64_t* valid;
    const float* g_weight;   // (B,H,W,K)
    const float* weight;     // optional (B,H,W,K): the forward's blend weights (Fragments.vert_weight); NULL = recompute
    const float* g_len_out;  // optional (B,H,W,K): gradient arriving on Fragments.vert_hit_length
    float omega;
    int B, N, H, W, K;
    float* grad_packed;      // (N, 4 | 8 | 12) for kind 1 | 3 | 9: [d verts(3), d sigma...] in float4 units
    int need_sigma;
    float* grad_rays;        // optional (B,H,W,3), written in full: d/d(ray direction) (pose optimisation)
    float* grad_origins;     // optional (B,3), zeroed by the caller: d/d(ray origin) = -sum over the view's hits of d/d(mu')
};

template <bool CAM>
__device__ __forceinline__ void geom_grad_accumulate(const FusedBwdArgs& a, int g, float m0, float m1, float m2,
                                                     const float* S, float d0, float d1, float d2, float ksk,
                                                     float msk, float gl, float ga, float gd, float* cam_acc) {
    const float g_ksk = (ga * msk - gl) * msk / (ksk * ksk) + gd;   // ray_trace_voge.cu:324-326
    const float g_msk = (gl - 2.f * ga * msk) / ksk;
    const float g_msm = ga;
    const float Sd0 = S[0] * d0 + S[1] * d1 + S[2] * d2, Sd1 = S[3] * d0 + S[4] * d1 + S[5] * d2,
                Sd2 = S[6] * d0 + S[7] * d1 + S[8] * d2;
    const float Sm0 = S[0] * m0 + S[1] * m1 + S[2] * m2, Sm1 = S[3] * m0 + S[4] * m1 + S[5] * m2,
                Sm2 = S[6] * m0 + S[7] * m1 + S[8] * m2;
    const float Stm0 = S[0] * m0 + S[3] * m1 + S[6] * m2, Stm1 = S[1] * m0 + S[4] * m1 + S[7] * m2,
                Stm2 = S[2] * m0 + S[5] * m1 + S[8] * m2;
    // One 16-byte vector reduction (red.global.add.v4.f32, sm_90+) per 4 gradient components instead of
    // 4 scalar ones: the packed per-Gaussian record is [d verts(3) | d sigma ...] padded to float4s
    // (kind 1: 4 floats, kind 3: 8, kind 9: 12).  The reference issues 45 scalar atomics per hit.
    const float gv0 = g_msk * Sd0 + g_msm * (Sm0 + Stm0);
    const float gv1 = g_msk * Sd1 + g_msm * (Sm1 + Stm1);
    const float gv2 = g_msk * Sd2 + g_msm * (Sm2 + Stm2);
    if (CAM) {
        // d/d(ray) = g_ksk (S + S^T) d + g_msk S^T mu  (ray_trace_voge.cu:41-91); d/d(origin) = -d/d(mu')
        const float Std0 = S[0] * d0 + S[3] * d1 + S[6] * d2, Std1 = S[1] * d0 + S[4] * d1 + S[7] * d2,
                    Std2 = S[2] * d0 + S[5] * d1 + S[8] * d2;
        cam_acc[0] += g_ksk * (Sd0 + Std0) + g_msk * Stm0;
        cam_acc[1] += g_ksk * (Sd1 + Std1) + g_msk * Stm1;
        cam_acc[2] += g_ksk * (Sd2 + Std2) + g_msk * Stm2;
        cam_acc[3] -= gv0; cam_acc[4] -= gv1; cam_acc[5] -= gv2;
    }
    const float dv[3] = {d0, d1, d2}, mv[3] = {m0, m1, m2};
    if (a.kind == 1) {
        const float tr = g_ksk * (d0 * d0 + d1 * d1 + d2 * d2) + g_msk * (m0 * d0 + m1 * d1 + m2 * d2) +
                         g_msm * (m0 * m0 + m1 * m1 + m2 * m2);
        atomicAdd(reinterpret_cast<float4*>(a.grad_packed + 4 * (int64_t)g), make_float4(gv0, gv1, gv2, 2.f * tr));
    } else if (a.kind == 3) {
        float gs[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) gs[i] = 2.f * (g_ksk * dv[i] * dv[i] + g_msk * mv[i] * dv[i] + g_msm * mv[i] * mv[i]);
        float4* p = reinterpret_cast<float4*>(a.grad_packed + 8 * (int64_t)g);
        atomicAdd(p, make_float4(gv0, gv1, gv2, 0.f));
        if (a.need_sigma) atomicAdd(p + 1, make_float4(gs[0], gs[1], gs[2], 0.f));
    } else {
        float gs[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                gs[3 * i + j] = 2.f * (g_ksk * dv[i] * dv[j] + g_msk * mv[i] * dv[j] + g_msm * mv[i] * mv[j]);
        float4* p = reinterpret_cast<float4*>(a.grad_packed + 12 * (int64_t)g);
        if (a.need_sigma) {
            atomicAdd(p, make_float4(gv0, gv1, gv2, gs[0]));
            atomicAdd(p + 1, make_float4(gs[1], gs[2], gs[3], gs[4]));
            atomicAdd(p + 2, make_float4(gs[5], gs[6], gs[7], gs[8]));
        } else {
            atomicAdd(p, make_float4(gv0, gv1, gv2, 0.f));
        }
    }
}

template <int NT, int KIND, bool CAM>
__global__ void __launch_bounds__(NT) render_bwd_fused_kernel(const FusedBwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const size_t A = (size_t)a.K * NT;
    float2* s_ls = reinterpret_cast<float2*>(smem_raw);   // (len, s = sqrt(dsd + 1e-10))
    float* s_E = reinterpret_cast<float*>(s_ls + A);      // exp(-act)
    float* s_wg = s_E + A;                                // w_m * dL/dw_m
    const int tid = threadIdx.x;
    // 8x4 pixel block per warp so that lanes of a warp touch the same Gaussians
    const int bw = (a.W + 7) / 8, bh = (a.H + 3) / 4;
    const int64_t wid = ((int64_t)blockIdx.x * NT + tid) >> 5;
    const int lane = tid & 31;
    const int64_t per_view = (int64_t)bw * bh;
    if (wid >= per_view * a.B) return;
    const int b = (int)(wid / per_view);
    const int wb = (int)(wid % per_view);
    const int xi = (wb % bw) * 8 + (lane & 7), yi = (wb / bw) * 4 + (lane >> 3);
    const bool live = xi < a.W && yi < a.H;
    if (!CAM && !live) return;       // with camera gradients every lane stays for the warp reduction at the end
    const int64_t r = live ? ((int64_t)b * a.H + yi) * a.W + xi : 0;
    const int cnt = live ? (int)min((int64_t)a.K, a.valid[r]) : 0;
    if (!CAM && cnt == 0) return;
    const float d0 = a.rays[r * 3 + 0], d1 = a.rays[r * 3 + 1], d2 = a.rays[r * 3 + 2];
    const float c0 = a.origins[3 * b], c1 = a.origins[3 * b + 1], c2 = a.origins[3 * b + 2];
    const float omega = a.omega;
    float cam_acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // d/d(ray) of this pixel, d/d(origin) partial sum
    const int32_t* i_idx = a.idx + r * a.K;
    const float* i_gw = a.g_weight + r * a.K;
    const int pack_off = b * a.N;
    const bool vec = (a.K & 3) == 0;

    // ---- pass 0: recompute the hits (bit-faithful), four slots at a time so that the dependent gathers
    // (index -> mean, S) of several slots are in flight together ----
    float s_min = 3.0e38f;
    for (int k0 = 0; k0 < cnt; k0 += 4) {
        int gv[4] = {-1, -1, -1, -1};
        if (vec) {
            const int4 q = *reinterpret_cast<const int4*>(i_idx + k0);
            gv[0] = q.x; gv[1] = q.y; gv[2] = q.z; gv[3] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k0 + j < cnt) gv[j] = i_idx[k0 + j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            if (k < cnt) {
                const int g = gv[j] - pack_off;
                Hit h;
                h.len = kEmptyLen; h.act = kEmptyLen; h.dsd = 0.f;
                if (g >= 0 && g < a.N) h = exact_hit_packed<KIND>(a.gauss, g, c0, c1, c2, d0, d1, d2);
                const float sk = sqrtf(h.dsd + 1e-10f);
                s_ls[k * NT + tid] = make_float2(h.len, sk);
                s_E[k * NT + tid] = expf(-h.act);
                s_min = fminf(s_min, sk);
            }
        }
    }
    // ---- pass 1: weights.  w_m = e^.5 exp(-omega D_m) E_m, D_m = sum_k E_k Phi((len_m - len_k) s_k);
    // sorted lens => Phi = 1 below the window |len_m - len_k| s_min < 4 (running prefix of E), 0 above ----
    float total_gD = 0.f;
    if (a.weight != nullptr) {
        // the forward's weights are an output that autograd keeps alive anyway: w_m dL/dw_m needs no erf sums
        const float* i_w = a.weight + r * a.K;
        for (int m0 = 0; m0 < cnt; m0 += 4) {
            float wv[4], gv[4];
            if (vec) {
                const float4 w4 = *reinterpret_cast<const float4*>(i_w + m0);
                const float4 g4 = *reinterpret_cast<const float4*>(i_gw + m0);
                wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
                gv[0] = g4.x; gv[1] = g4.y; gv[2] = g4.z; gv[3] = g4.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    wv[j] = (m0 + j < cnt) ? i_w[m0 + j] : 0.f;
                    gv[j] = (m0 + j < cnt) ? i_gw[m0 + j] : 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (m0 + j < cnt) {
                    const float wg = wv[j] * gv[j];
                    s_wg[(m0 + j) * NT + tid] = wg;
                    total_gD -= omega * wg;       // gD_m = dL/dD_m = -omega w_m dL/dw_m
                }
            }
        }
    } else {
        int lo = 0;
        float SE = 0.f;
        float4 gw4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int m = 0; m < cnt; ++m) {
            // the (K,) rows are read 16 bytes at a time: a scalar load at a 4K-byte lane stride costs a full
            // L1 tag lookup per lane
            if (vec && (m & 3) == 0) gw4 = *reinterpret_cast<const float4*>(i_gw + m);
            const float gwm = vec ? ((m & 2) ? ((m & 1) ? gw4.w : gw4.z) : ((m & 1) ? gw4.y : gw4.x)) : i_gw[m];
            const float Em = s_E[m * NT + tid];
            const float lm = s_ls[m * NT + tid].x;
            while (lo < m && (lm - s_ls[lo * NT + tid].x) * s_min >= kErfSat) { SE += s_E[lo * NT + tid]; ++lo; }
            float wg = 0.f;
            if (Em != 0.f) {
                float D = fmaf(Em, 0.5f, SE);         // k = m: Phi(0) = 1/2; the loop visits the neighbours only
                for (int t = lo;; ++t) {
                    const int k = t + (t >= m ? 1 : 0);
                    if (k >= cnt) break;
                    const float2 lk = s_ls[k * NT + tid];
                    const float dl = lm - lk.x;
                    if (dl * s_min <= -kErfSat) break;
                    D += s_E[k * NT + tid] * phi(dl * lk.y);
                }
                wg = expf(-(D * omega)) * Em * kInvExpMinusHalf * gwm;
            }
            s_wg[m * NT + tid] = wg;
            total_gD -= omega * wg;       // gD_m = dL/dD_m = -omega w_m dL/dw_m
        }
    }
    // ---- pass 2: per slot j gather every contribution inside its (symmetric) window [lo_j, hi_j] in
    // registers and apply the chain rule of ray_trace_voge.cu:324-330 at once ----
    {
        int lo_j = 0, hi_j = -1;
        float pref = 0.f;                 // sum of gD_m over m <= hi_j
        int4 ix4 = make_int4(-1, -1, -1, -1);
        for (int j = 0; j < cnt; ++j) {
            // the Gaussian's record is fetched first: the gathers overlap with the window loop below
            if (vec && (j & 3) == 0) ix4 = *reinterpret_cast<const int4*>(i_idx + j);
            const int gp = vec ? ((j & 2) ? ((j & 1) ? ix4.w : ix4.z) : ((j & 1) ? ix4.y : ix4.x)) : i_idx[j];
            const int g = gp - pack_off;
            const bool g_ok = g >= 0 && g < a.N;
            float S[9], m0 = 0.f, m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int q = 0; q < 9; ++q) S[q] = 0.f;
            if (g_ok) {
                float v0, v1, v2;
                load_gauss<KIND>(a.gauss, g, v0, v1, v2, S);
                m0 = __fsub_rn(v0, c0); m1 = __fsub_rn(v1, c1); m2 = __fsub_rn(v2, c2);
            }
            const float2 lsj = s_ls[j * NT + tid];
            const float lj = lsj.x, sj = lsj.y;
            while (lo_j < j && (lj - s_ls[lo_j * NT + tid].x) * s_min >= kErfSat) ++lo_j;
            while (hi_j + 1 < cnt && (s_ls[(hi_j + 1) * NT + tid].x - lj) * s_min < kErfSat) {
                ++hi_j;
                pref -= omega * s_wg[hi_j * NT + tid];
            }
            const float Ej = s_E[j * NT + tid];
            if (Ej == 0.f) continue;
            const float wgj = s_wg[j * NT + tid];
            const float gDj = -omega * wgj;
            // direct path through the trailing exp(-act_j)  +  rows m behind the window see Phi = 1
            // i = j: c = 0 => Phi = 1/2 and the two d/d(len_j) terms cancel exactly; the loop visits the
            // neighbours only (a warp iterates max-over-lanes(neighbours) times)
            float gE = wgj / Ej + (total_gD - pref) + 0.5f * gDj;
            float gl = 0.f, gd = 0.f;
            const float inv2sj = 0.5f / sj;
            const float gDjk = gDj * kInvSqrtPi, Ejk = Ej * kInvSqrtPi;
            for (int t = lo_j; t < hi_j; ++t) {
                const int i = t + (t >= j ? 1 : 0);
                const float2 lsi = s_ls[i * NT + tid];
                const float dl = lsi.x - lj;
                const float gDi = -omega * s_wg[i * NT + tid];
                // row i, column j:  c = (len_i - len_j) s_j.  exp(-c^2), |c| < 4, through ex2.approx (2 ulp):
                // these terms are the erf-slope corrections of the gradient, three orders below its tolerance
                const float c = dl * sj;
                if (c >= kErfSat) {
                    gE += gDi;
                } else if (c > -kErfSat) {
                    float ec;
                    gE += gDi * phi_fast(c, ec);
                    const float gc = gDi * Ejk * ec;
                    gl -= gc * sj;
                    gd += gc * dl * inv2sj;
                }
                // row j, column i:  c' = (len_j - len_i) s_i  ->  d/d len_j
                const float c2_ = -dl * lsi.y;
                if (fabsf(c2_) < kErfSat) gl += gDjk * s_E[i * NT + tid] * __expf(-c2_ * c2_) * lsi.y;
            }
            const float ga = -Ej * gE;
            if (a.g_len_out != nullptr) gl += a.g_len_out[r * a.K + j];
            if (ga == 0.f && gl == 0.f && gd == 0.f) continue;
            if (!g_ok) continue;
            const Prod9 pd = exact_row_products(d0, d1, d2, S);
            const Prod9 pm = exact_row_products(m0, m1, m2, S);
            const float ksk = exact_contract(pd, d0, d1, d2);
            const float msk = exact_contract(pm, d0, d1, d2);
            geom_grad_accumulate<CAM>(a, g, m0, m1, m2, S, d0, d1, d2, ksk, msk, gl, ga, gd, cam_acc);
        }
    }
    if (CAM) {
        if (a.grad_rays != nullptr && live) {
            a.grad_rays[r * 3 + 0] = cam_acc[0]; a.grad_rays[r * 3 + 1] = cam_acc[1]; a.grad_rays[r * 3 + 2] = cam_acc[2];
        }
        if (a.grad_origins != nullptr) {
            __syncwarp();
#pragma unroll
            for (int q = 3; q < 6; ++q) {
                float v = cam_acc[q];          // the 32 lanes of a warp belong to one view
                for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_down_sync(0xffffffffu, v, sft);
                if (lane == 0) atomicAdd(a.grad_origins + 3 * b + (q - 3), v);
            }
        }
    }
}

// Two threads per pixel (adjacent lanes; a warp covers a 4x4 pixel block): the per-pixel arrays in shared
// memory bound the resident pixels per SM, and with one thread per pixel that left 20 warps of serial,
// latency-bound work.  The pair shares the arrays; thread `sub` owns the slots k = sub, sub + 2, ... in every pass.
template <int NT, int KIND, bool CAM>
__global__ void __launch_bounds__(NT) render_bwd_pair_kernel(const FusedBwdArgs a) {
    constexpr int NP = NT / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const size_t A = (size_t)a.K * NP;
    float2* s_ls = reinterpret_cast<float2*>(smem_raw);   // (len, s = sqrt(dsd + 1e-10))
    float* s_E = reinterpret_cast<float*>(s_ls + A);      // exp(-act)
    float* s_wg = s_E + A;                                // w_m * dL/dw_m
    const int tid = threadIdx.x;
    const int lane = tid & 31, sub = lane & 1, col = tid >> 1;
    const unsigned pair_mask = 3u << (lane & 30);
    const int bw = (a.W + 3) / 4, bh = (a.H + 3) / 4;
    const int64_t wid = ((int64_t)blockIdx.x * NT + tid) >> 5;
    const int64_t per_view = (int64_t)bw * bh;
    if (wid >= per_view * a.B) return;
    const int b = (int)(wid / per_view);
    const int wb = (int)(wid % per_view);
    const int pp = lane >> 1;
    const int xi = (wb % bw) * 4 + (pp & 3), yi = (wb / bw) * 4 + (pp >> 2);
    const bool live = xi < a.W && yi < a.H;
    if (!CAM && !live) return;       // with camera gradients every lane stays for the warp reduction at the end
    const int64_t r = live ? ((int64_t)b * a.H + yi) * a.W + xi : 0;
    const int cnt = live ? (int)min((int64_t)a.K, a.valid[r]) : 0;
    if (!CAM && cnt == 0) return;
    const float d0 = a.rays[r * 3 + 0], d1 = a.rays[r * 3 + 1], d2 = a.rays[r * 3 + 2];
    const float c0 = a.origins[3 * b], c1 = a.origins[3 * b + 1], c2 = a.origins[3 * b + 2];
    const float omega = a.omega;
    float cam_acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // d/d(ray), d/d(origin) partial sums of this thread's slots
    const int32_t* i_idx = a.idx + r * a.K;
    const float* i_gw = a.g_weight + r * a.K;
    const int pack_off = b * a.N;
    const bool vec = (a.K & 3) == 0;

    // ---- pass 0: recompute this thread's hits (bit-faithful); w_m dL/dw_m from the forward's weights ----
    float s_min = 3.0e38f, total_gD = 0.f;
    for (int k0 = 0; k0 < cnt; k0 += 4) {
        int gv[2] = {-1, -1};
        float wv[2] = {0.f, 0.f}, gwv[2] = {0.f, 0.f};
        if (vec) {
            const int4 q = *reinterpret_cast<const int4*>(i_idx + k0);
            const float4 g4 = *reinterpret_cast<const float4*>(i_gw + k0);
            gv[0] = sub ? q.y : q.x; gv[1] = sub ? q.w : q.z;
            gwv[0] = sub ? g4.y : g4.x; gwv[1] = sub ? g4.w : g4.z;
            if (a.weight != nullptr) {
                const float4 w4 = *reinterpret_cast<const float4*>(a.weight + r * a.K + k0);
                wv[0] = sub ? w4.y : w4.x; wv[1] = sub ? w4.w : w4.z;
            }
        } else {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int k = k0 + sub + 2 * jj;
                if (k < cnt) {
                    gv[jj] = i_idx[k]; gwv[jj] = i_gw[k];
                    if (a.weight != nullptr) wv[jj] = a.weight[r * a.K + k];
                }
            }
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int k = k0 + sub + 2 * jj;
            if (k < cnt) {
                const int g = gv[jj] - pack_off;
                Hit h;
                h.len = kEmptyLen; h.act = kEmptyLen; h.dsd = 0.f;
                if (g >= 0 && g < a.N) h = exact_hit_packed<KIND>(a.gauss, g, c0, c1, c2, d0, d1, d2);
                const float sk = sqrtf(h.dsd + 1e-10f);
                s_ls[k * NP + col] = make_float2(h.len, sk);
                s_E[k * NP + col] = expf(-h.act);
                s_min = fminf(s_min, sk);
                if (a.weight != nullptr) {
                    const float wg = wv[jj] * gwv[jj];
                    s_wg[k * NP + col] = wg;
                    total_gD -= omega * wg;       // gD_m = dL/dD_m = -omega w_m dL/dw_m
                } else {
                    s_wg[k * NP + col] = gwv[jj];  // upstream gradient, turned into w g below
                }
            }
        }
    }
    s_min = fminf(s_min, __shfl_xor_sync(pair_mask, s_min, 1));
    __syncwarp(pair_mask);
    if (a.weight == nullptr) {
        // ---- pass 1 (no saved weights): w_m = e^.5 exp(-omega D_m) E_m, D_m = sum_k E_k Phi((len_m - len_k) s_k);
        // sorted lens => Phi = 1 below the window |len_m - len_k| s_min < 4 (running prefix of E), 0 above ----
        int lo = 0;
        float SE = 0.f;
        for (int m = sub; m < cnt; m += 2) {
            const float Em = s_E[m * NP + col];
            const float lm = s_ls[m * NP + col].x;
            while (lo < m && (lm - s_ls[lo * NP + col].x) * s_min >= kErfSat) { SE += s_E[lo * NP + col]; ++lo; }
            float wg = 0.f;
            if (Em != 0.f) {
                float D = fmaf(Em, 0.5f, SE);         // k = m: Phi(0) = 1/2; the loop visits the neighbours only
                for (int t = lo;; ++t) {
                    const int k = t + (t >= m ? 1 : 0);
                    if (k >= cnt) break;
                    const float2 lk = s_ls[k * NP + col];
                    const float dl = lm - lk.x;
                    if (dl * s_min <= -kErfSat) break;
                    D += s_E[k * NP + col] * phi_f(dl * lk.y);
                }
                wg = expf(-(D * omega)) * Em * kInvExpMinusHalf * s_wg[m * NP + col];
            }
            s_wg[m * NP + col] = wg;
            total_gD -= omega * wg;
        }
        __syncwarp(pair_mask);
    }
    total_gD += __shfl_xor_sync(pair_mask, total_gD, 1);
    // ---- pass 2: per slot j gather every contribution inside its (symmetric) window [lo_j, hi_j] in
    // registers and apply the chain rule of ray_trace_voge.cu:324-330 at once ----
    {
        int lo_j = 0, hi_j = -1;
        float pref = 0.f;                 // sum of gD_m over m <= hi_j
        int4 ix4 = make_int4(-1, -1, -1, -1);
        for (int j = sub; j < cnt; j += 2) {
            // the Gaussian's record is fetched first: the gathers overlap with the window loop below
            if (vec && (j & 3) == sub) ix4 = *reinterpret_cast<const int4*>(i_idx + (j & ~3));
            const int gp = vec ? ((j & 2) ? ((j & 1) ? ix4.w : ix4.z) : ((j & 1) ? ix4.y : ix4.x)) : i_idx[j];
            const int g = gp - pack_off;
            const bool g_ok = g >= 0 && g < a.N;
            float S[9], m0 = 0.f, m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int q = 0; q < 9; ++q) S[q] = 0.f;
            if (g_ok) {
                float v0, v1, v2;
                load_gauss<KIND>(a.gauss, g, v0, v1, v2, S);
                m0 = __fsub_rn(v0, c0); m1 = __fsub_rn(v1, c1); m2 = __fsub_rn(v2, c2);
            }
            const float2 lsj = s_ls[j * NP + col];
            const float lj = lsj.x, sj = lsj.y;
            while (lo_j < j && (lj - s_ls[lo_j * NP + col].x) * s_min >= kErfSat) ++lo_j;
            while (hi_j + 1 < cnt && (s_ls[(hi_j + 1) * NP + col].x - lj) * s_min < kErfSat) {
                ++hi_j;
                pref -= omega * s_wg[hi_j * NP + col];
            }
            const float Ej = s_E[j * NP + col];
            if (Ej == 0.f) continue;
            const float wgj = s_wg[j * NP + col];
            const float gDj = -omega * wgj;
            // direct path through the trailing exp(-act_j)  +  rows m behind the window see Phi = 1
            // i = j: c = 0 => Phi = 1/2 and the two d/d(len_j) terms cancel exactly; the loop visits the
            // neighbours only (a warp iterates max-over-lanes(neighbours) times)
            float gE = wgj / Ej + (total_gD - pref) + 0.5f * gDj;
            float gl = 0.f, gd = 0.f;
            const float inv2sj = 0.5f / sj;
            const float gDjk = gDj * kInvSqrtPi, Ejk = Ej * kInvSqrtPi;
            for (int t = lo_j; t < hi_j; ++t) {
                const int i = t + (t >= j ? 1 : 0);
                const float2 lsi = s_ls[i * NP + col];
                const float dl = lsi.x - lj;
                const float gDi = -omega * s_wg[i * NP + col];
                // row i, column j:  c = (len_i - len_j) s_j.  exp(-c^2), |c| < 4, through ex2.approx (2 ulp):
                // these terms are the erf-slope corrections of the gradient, three orders below its tolerance
                const float c = dl * sj;
                if (c >= kErfSat) {
                    gE += gDi;
                } else if (c > -kErfSat) {
                    float ec;
                    gE += gDi * phi_fast(c, ec);
                    const float gc = gDi * Ejk * ec;
                    gl -= gc * sj;
                    gd += gc * dl * inv2sj;
                }
                // row j, column i:  c' = (len_j - len_i) s_i  ->  d/d len_j
                const float c2_ = -dl * lsi.y;
                if (fabsf(c2_) < kErfSat) gl += gDjk * s_E[i * NP + col] * __expf(-c2_ * c2_) * lsi.y;
            }
            const float ga = -Ej * gE;
            if (a.g_len_out != nullptr) gl += a.g_len_out[r * a.K + j];
            if (ga == 0.f && gl == 0.f && gd == 0.f) continue;
            if (!g_ok) continue;
            const Prod9 pd = exact_row_products(d0, d1, d2, S);
            const Prod9 pm = exact_row_products(m0, m1, m2, S);
            const float ksk = exact_contract(pd, d0, d1, d2);
            const float msk = exact_contract(pm, d0, d1, d2);
            geom_grad_accumulate<CAM>(a, g, m0, m1, m2, S, d0, d1, d2, ksk, msk, gl, ga, gd, cam_acc);
        }
    }
    if (CAM) {
        __syncwarp();
        if (a.grad_rays != nullptr) {
#pragma unroll
            for (int q = 0; q < 3; ++q) cam_acc[q] += __shfl_xor_sync(0xffffffffu, cam_acc[q], 1);
            if (live && sub == 0) {
                a.grad_rays[r * 3 + 0] = cam_acc[0]; a.grad_rays[r * 3 + 1] = cam_acc[1]; a.grad_rays[r * 3 + 2] = cam_acc[2];
            }
        }
        if (a.grad_origins != nullptr) {
#pragma unroll
            for (int q = 3; q < 6; ++q) {
                float v = cam_acc[q];          // the 32 lanes of a warp belong to one view
                for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_down_sync(0xffffffffu, v, sft);
                if (lane == 0) atomicAdd(a.grad_origins + 3 * b + (q - 3), v);
            }
        }
    }
}

}  // namespace voge

extern "C" int voge_render_backward_fused(const float* gauss, int sigma_kind,
                                          const float* origins, const float* rays, const int32_t* idx,
                                          const int64_t* valid, const float* grad_weight, const float* weight,
                                          const float* grad_len_out, float absorptivity, int B, int N, int H,
                                          int W, int K, float* grad_packed, int need_sigma, float* grad_rays,
                                          float* grad_origins, voge_stream_t stream) {
    using namespace voge;
    if (B <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
    FusedBwdArgs a{gauss, sigma_kind, origins, rays, idx, valid, grad_weight, weight, grad_len_out, absorptivity,
                   B, N, H, W, K, grad_packed, need_sigma, grad_rays, grad_origins};
    if (sigma_kind != 1 && sigma_kind != 3 && sigma_kind != 9) return (int)cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)stream;
    const bool cam = grad_rays != nullptr || grad_origins != nullptr;
    auto launch = [&](auto kernel, int nt, int per_pixel) -> int {
        // per_pixel = 2: two threads per pixel, 4x4 pixel blocks per warp; 1: 8x4 blocks
        const int64_t warps = per_pixel == 2 ? (int64_t)B * cdiv(W, 4) * cdiv(H, 4) : (int64_t)B * cdiv(W, 8) * cdiv(H, 4);
        const size_t smem = (size_t)K * (nt / per_pixel) * 16;
        if (smem > 227 * 1024) return (int)cudaErrorInvalidValue;
        VOGE_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t grid = (warps * 32 + nt - 1) / nt;
        if (grid > 2147483647LL) return (int)cudaErrorInvalidValue;
        kernel<<<(unsigned)grid, nt, smem, s>>>(a);
        VOGE_LAUNCH_CHECK();
        return 0;
    };
    auto by_threads = [&](auto kind_tag, auto cam_tag) -> int {
        constexpr int KIND = decltype(kind_tag)::value;
        constexpr bool CAM = decltype(cam_tag)::value;
        if (K <= 112) return launch(render_bwd_pair_kernel<128, KIND, CAM>, 128, 2);
        if (K <= 200) return launch(render_bwd_fused_kernel<64, KIND, CAM>, 64, 1);
        return launch(render_bwd_fused_kernel<32, KIND, CAM>, 32, 1);
    };
    auto by_kind = [&](auto cam_tag) -> int {
        if (sigma_kind == 1) return by_threads(std::integral_constant<int, 1>{}, cam_tag);
        if (sigma_kind == 3) return by_threads(std::integral_constant<int, 3>{}, cam_tag);
        return by_threads(std::integral_constant<int, 9>{}, cam_tag);
    };
    return cam ? by_kind(std::true_type{}) : by_kind(std::false_type{});
}

namespace voge {
__global__ void __launch_bounds__(256) pack_gaussians_kernel(const float* __restrict__ verts, const float* __restrict__ sigmas,
                                                             int kind, int N, float* __restrict__ out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    const float x = verts[3 * (int64_t)g], y = verts[3 * (int64_t)g + 1], z = verts[3 * (int64_t)g + 2];
    float S[9];
    load_S_dyn(kind, sigmas, g, S);                     // S = 2 sigma (Renderer.py:137), exact in fp32
    float4* o = reinterpret_cast<float4*>(out);
    if (kind == 1) {
        o[g] = make_float4(x, y, z, S[0]);
    } else if (kind == 3) {
        o[2 * (int64_t)g] = make_float4(x, y, z, S[0]);
        o[2 * (int64_t)g + 1] = make_float4(S[4], S[8], 0.f, 0.f);
    } else {
        o[3 * (int64_t)g] = make_float4(x, y, z, S[0]);
        o[3 * (int64_t)g + 1] = make_float4(S[1], S[2], S[3], S[4]);
        o[3 * (int64_t)g + 2] = make_float4(S[5], S[6], S[7], S[8]);
    }
}
}  // namespace voge

extern "C" int voge_pack_gaussians(const float* verts, const float* sigmas, int sigma_kind, int N, float* out,
                                   voge_stream_t stream) {
    using namespace voge;
    if (N <= 0) return 0;
    if (sigma_kind != 1 && sigma_kind != 3 && sigma_kind != 9) return (int)cudaErrorInvalidValue;
    pack_gaussians_kernel<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(verts, sigmas, sigma_kind, N, out);
    VOGE_LAUNCH_CHECK();
    return 0;
}

extern "C" int voge_bin_sub(void) { return voge::kBinSub; }

extern "C" int voge_bin_count(const float* verts, const float* sigmas, int sigma_kind, const float* Rm,
                              const float* Tv, const float* origins, const float* focal, const float* principal,
                              int B, int N, int H, int W, float thr, float thr_act, int use_ref_bins, int bin_size,
                              int tile, uint32_t* rects, int32_t* tile_counts, int32_t* tile_items, voge_stream_t stream) {
    using namespace voge;
    if (B <= 0 || N <= 0) return 0;
    if (tile <= 0 || tile > 16 || (use_ref_bins && (bin_size <= 0 || bin_size % tile != 0))) return (int)cudaErrorInvalidValue;
    if (sigma_kind != 1 && sigma_kind != 3 && sigma_kind != 9) return (int)cudaErrorInvalidValue;
    BinArgs a;
    a.verts = verts; a.sigmas = sigmas; a.kind = sigma_kind; a.Rm = Rm; a.Tv = Tv; a.origins = origins;
    a.focal = focal; a.principal = principal; a.B = B; a.N = N; a.H = H; a.W = W;
    a.neg_log_thr = -logf(thr); a.thr_act = thr_act; a.use_ref_bins = use_ref_bins; a.bin_size = bin_size;
    a.BH = use_ref_bins ? cdiv(H, bin_size) : 1; a.BW = use_ref_bins ? cdiv(W, bin_size) : 1;
    a.tile = tile; a.TX = cdiv(W, tile); a.TY = cdiv(H, tile);
    if (a.TX > 65535 || a.TY > 65535) return (int)cudaErrorInvalidValue;
    a.rects = reinterpret_cast<uint2*>(rects); a.tile_counts = tile_counts; a.tile_items = tile_items;
    dim3 grid(cdiv(N, 256), B);   // one Gaussian per thread: the dependent load -> atomic -> store chain is pure latency
    bin_count_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    VOGE_LAUNCH_CHECK();
    return 0;
}

extern "C" int voge_bin_fill(const uint32_t* rects, const int64_t* tile_offsets, int32_t* cursor, int B, int N,
                             int H, int W, int tile, int32_t* tile_list, voge_stream_t stream) {
    using namespace voge;
    if (B <= 0 || N <= 0) return 0;
    dim3 grid(cdiv(N, 256), B);   // one Gaussian per thread: the dependent load -> atomic -> store chain is pure latency
    bin_fill_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint2*>(rects), tile_offsets, cursor,
                                                             B, N, cdiv(W, tile), cdiv(H, tile), tile, tile_list);
    VOGE_LAUNCH_CHECK();
    return 0;
}

extern "C" int voge_render_forward(const float* verts, const float* sigmas, int sigma_kind, const float* origins,
                                   const float* rays, const int64_t* tile_offsets, const int32_t* tile_list,
                                   const uint32_t* rects, float thr_act, float absorptivity, int B, int N, int H,
                                   int W, int K, int tile,
                                   int32_t* out_idx, float* out_weight, float* out_len, int64_t* out_valid,
                                   float* out_act, float* out_dsd, uint64_t* stats, voge_stream_t stream) {
    using namespace voge;
    if (B <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
    RenderArgs a;
    a.verts = verts; a.sigmas = sigmas; a.origins = origins; a.rays = rays; a.tile_offsets = tile_offsets;
    a.tile_list = tile_list; a.rects = reinterpret_cast<const uint2*>(rects); a.cap = 0; a.thr_act = thr_act; a.omega = absorptivity; a.B = B; a.N = N; a.H = H; a.W = W; a.K = K;
    a.tile = tile; a.TX = cdiv(W, tile); a.TY = cdiv(H, tile);
    a.out_idx = out_idx; a.out_weight = out_weight; a.out_len = out_len; a.out_valid = out_valid;
    a.out_act = out_act; a.out_dsd = out_dsd; a.stats = reinterpret_cast<unsigned long long*>(stats);
    cudaStream_t s = (cudaStream_t)stream;
    if (sigma_kind == 1) return dispatch_render<1>(a, s);
    if (sigma_kind == 3) return dispatch_render<3>(a, s);
    if (sigma_kind == 9) return dispatch_render<9>(a, s);
    return (int)cudaErrorInvalidValue;
}

extern "C" int voge_render_backward(const float* verts, const float* sigmas, int sigma_kind, const float* origins,
                                    const float* rays, const int32_t* idx, const int64_t* valid,
                                    const float* grad_len, const float* grad_act, const float* grad_dsd, int B,
                                    int N, int H, int W, int K,
                                    float* grad_verts, float* grad_sigmas, voge_stream_t stream) {
    using namespace voge;
    if (B <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
    RenderBwdArgs a{verts, sigmas, sigma_kind, origins, rays, idx, valid, grad_len, grad_act, grad_dsd, B, N, H, W, K,
                    grad_verts, grad_sigmas};
    const int64_t warps = (int64_t)B * cdiv(W, 8) * cdiv(H, 4);
    const int64_t grid = (warps * 32 + 255) / 256;
    render_bwd_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(a);
    VOGE_LAUNCH_CHECK();
    return 0;
}

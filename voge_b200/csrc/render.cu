// render.cu -- the fused renderer path behind GaussianRenderer.forward:
//   voge_bin_count / voge_bin_fill : per-view screen-space culling into per-tile CSR lists
//   voge_pack_gaussians / voge_unpack_gradients : per-Gaussian parameter records (S = 2 sigma, 2 inverse(sigma) or
//                                    2 L L^T) and the matching gradient epilogue
//   voge_generate_rays             : closed-form pixel rays (never materialised on the fused path)
//   voge_render_backward_fused     : recompute + analytic blend backward + chain rule straight into (N,.) gradients
//   (forward pipeline: trace.cu, select.cu)
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
    const float* gauss;      // packed records (voge_pack_gaussians): [x, y, z, S = 2 P ...]
    int kind;
    int enc;                 // kind-9 records carry the isotropic encoding (render_core.cuh: kKindIsoEncoded)
    const float* Rm;         // (B,3,3) row-vector convention X_view = X_world @ R + T
    const float* Tv;         // (B,3)
    const float* origins;    // (B,3) ray origins (camera centres) used for mu' = verts - origin
    const float* focal;      // (B,2) pixels
    const float* principal;  // (B,2) pixels
    int B, N, H, W;
    float neg_log_thr, thr_act;
    int use_ref_bins, bin_size, BH, BW, tile, TX, TY;
    int zero_aware_margin;   // 1: rounding margin counts the non-zero entries of S (default); 0: dense-S constants
    uint2* rects;            // (B,N): x = x0 | x1 << 16, y = y0 | y1 << 16 in PIXELS (inclusive); empty if x0 > x1
    unsigned long long* tile_counters;   // (B, TY*TX, kBinSub) one 64-bit counter per list segment: low word = list entries,
                                         // high word = items (sum over the entries of the rectangle area inside the tile)
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
// Zero structure: a term whose S_ij is exactly 0 contributes t_ij = 0 and fma(0, c, acc) == acc exactly, i.e. neither a
// product rounding nor a partial-sum rounding.  With n = number of non-zero entries of S every form passes through at
// most 1 + n roundings (10 for a dense S, 4 for a diagonal / isotropic one stored as 3x3 or as a compact kind), so the
// constants 10 / 20 / 10 above become (1 + n) / 2 (1 + n) / (1 + n): the margin of a diagonal Gaussian shrinks to 0.43x.
// The S-only part (eigenvalues of the symmetric part, Us, the rounding count) is evaluated once per Gaussian and call;
// the part that depends on mu = verts - camera centre once per view (bin_count_kernel loops over the views).
struct MarginS {
    bool ok;       // positive definite, second-order terms negligible
    float il;      // 1 / lambda_min(sym S)
    float Us;      // >= || |S| ||_2
    float g1;      // roundings per 9-term form: 10, or 1 + nnz(S)
};

__device__ __forceinline__ MarginS gaussian_margin_s(const float* S, bool zero_aware) {
    MarginS r;
    r.ok = false; r.il = 0.f; r.Us = 0.f; r.g1 = 10.f;
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
            float rr = 0.5f * (b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02));
            rr = fminf(1.f, fmaxf(-1.f, rr));
            const float phi = acosf(rr) * (1.f / 3.f);
            lmax = qq + 2.f * p * cosf(phi);
            lmin = qq + 2.f * p * cosf(phi + 2.0943951023931953f);
        }
        lmin -= 1e-5f * fabsf(lmax);   // rounding of the closed form itself
    }
    if (!(lmin > 0.f)) return r;
    const float r0 = fabsf(S[0]) + fabsf(S[1]) + fabsf(S[2]), r1 = fabsf(S[3]) + fabsf(S[4]) + fabsf(S[5]),
                r2 = fabsf(S[6]) + fabsf(S[7]) + fabsf(S[8]);
    const float k0 = fabsf(S[0]) + fabsf(S[3]) + fabsf(S[6]), k1 = fabsf(S[1]) + fabsf(S[4]) + fabsf(S[7]),
                k2 = fabsf(S[2]) + fabsf(S[5]) + fabsf(S[8]);
    r.Us = fmaxf(fmaxf(fmaxf(r0, r1), r2), fmaxf(fmaxf(k0, k1), k2));
    r.il = 1.f / lmin;
    if (!(r.Us * r.il < 1.6e4f)) return r;                       // g10 Us / l < 1e-2: second-order terms negligible
    if (zero_aware) {
        int nnz = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) nnz += (S[i] != 0.f) ? 1 : 0;
        r.g1 = (float)(1 + nnz);
    }
    r.ok = true;
    return r;
}

__device__ __forceinline__ bool gaussian_margin(const float* mu, const float* S, float thr_act, const MarginS& ms, float& margin) {
    if (!ms.ok) return false;
    const float m0 = mu[0], m1 = mu[1], m2 = mu[2];
    const float q0 = m0 * S[0] + m1 * S[3] + m2 * S[6];
    const float q1 = m0 * S[1] + m1 * S[4] + m2 * S[7];
    const float q2 = m0 * S[2] + m1 * S[5] + m2 * S[8];
    const float am0 = fabsf(m0), am1 = fabsf(m1), am2 = fabsf(m2);
    const float c0 = am0 * fabsf(S[0]) + am1 * fabsf(S[3]) + am2 * fabsf(S[6]);   // (|S|^T |mu|)_j
    const float c1 = am0 * fabsf(S[1]) + am1 * fabsf(S[4]) + am2 * fabsf(S[7]);
    const float c2 = am0 * fabsf(S[2]) + am1 * fabsf(S[5]) + am2 * fabsf(S[8]);
    const float Tmm = c0 * am0 + c1 * am1 + c2 * am2;
    const float Um2 = sqrtf(c0 * c0 + c1 * c1 + c2 * c2);
    const float Qn = sqrtf(q0 * q0 + q1 * q1 + q2 * q2);
    const float il = ms.il, Us = ms.Us, g1 = ms.g1;
    const float bound = g1 * Tmm + 2.f * g1 * Qn * Um2 * il + g1 * Us * (Qn * il) * (Qn * il) + 2.1f * Qn * Qn * il +
                        2.f * fabsf(thr_act);
    margin = 7.4505806e-8f * 1.003f * bound;                     // 1.25 x 2^-24 (x rays within 1e-3 of unit length)
    return margin >= 0.f && margin < 3.0e38f;
}

// One thread per Gaussian; the thread walks the views b = blockIdx.y, blockIdx.y + gridDim.y, ... of the call: the record is
// loaded once and everything that depends on S alone -- the eigenvalues / norms of the rounding margin, the symmetry tests,
// the adjugate of the tangent bound -- is evaluated once per Gaussian instead of once per (view, Gaussian).
__global__ void __launch_bounds__(256, 4) bin_count_kernel(const BinArgs a) {
    const float sc = 0.5f * (float)min(a.H, a.W);
    const float half_x = __fdiv_rn(ndc_range(a.W, a.H) / 2.0f, (float)a.W);
    const float half_y = __fdiv_rn(ndc_range(a.H, a.W) / 2.0f, (float)a.H);
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < a.N; g += gridDim.x * blockDim.x) {
        float S[9], v0, v1, v2;
        if (a.kind == 1) load_gauss<1>(a.gauss, g, v0, v1, v2, S);
        else if (a.kind == 3) load_gauss<3>(a.gauss, g, v0, v1, v2, S);
        else load_gauss<9>(a.gauss, g, v0, v1, v2, S, a.enc != 0);
        // ---- S only ----
        const float as01 = fabsf(S[1] - S[3]), as02 = fabsf(S[2] - S[6]), as12 = fabsf(S[5] - S[7]);
        const bool sym_exact = (as01 == 0.f) && (as02 == 0.f) && (as12 == 0.f);
        const bool sym_near = as01 <= 9.5367e-7f * (fabsf(S[1]) + fabsf(S[3])) &&
                              as02 <= 9.5367e-7f * (fabsf(S[2]) + fabsf(S[6])) &&
                              as12 <= 9.5367e-7f * (fabsf(S[5]) + fabsf(S[7]));
        MarginS ms;
        ms.ok = false;
        if (sym_near) ms = gaussian_margin_s(S, a.zero_aware_margin != 0);
        // inverse of the symmetric part (adjugate / det)
        const float a00 = S[0], a11 = S[4], a22 = S[8];
        const float a01 = 0.5f * (S[1] + S[3]), a02 = 0.5f * (S[2] + S[6]), a12 = 0.5f * (S[5] + S[7]);
        const float j00 = a11 * a22 - a12 * a12, j01 = a02 * a12 - a01 * a22, j02 = a01 * a12 - a02 * a11;
        const float j11 = a00 * a22 - a02 * a02, j12 = a01 * a02 - a00 * a12, j22 = a00 * a11 - a01 * a01;
        const float det = a00 * j00 + a01 * j01 + a02 * j02;
        const float J[9] = {j00, j01, j02, j01, j11, j12, j02, j12, j22};
      for (int b = blockIdx.y; b < a.B; b += gridDim.y) {
        const float* R = a.Rm + 9 * b;
        const float fx = a.focal[2 * b], fy = a.focal[2 * b + 1];
        const float px = a.principal[2 * b], py = a.principal[2 * b + 1];
        const float c0 = a.origins[3 * b], c1 = a.origins[3 * b + 1], c2 = a.origins[3 * b + 2];
        float mu[3];
        mu[0] = __fsub_rn(v0, c0); mu[1] = __fsub_rn(v1, c1); mu[2] = __fsub_rn(v2, c2);
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
        if (!empty && zv > 0.f && sym_near && gaussian_margin(mu, S, a.thr_act, ms, margin)) {
            if (!sym_exact) margin *= 2.f;
            const float t = (a.thr_act + margin) * 1.0001f / det;
            // Q = t * R^T J R ; need Q00 Q11 Q22 Q02 Q12
            float JR[3][3];
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
            // ONE 64-bit reduction per (entry, tile) counts the entry and its items (both counters share the request)
            unsigned long long* cnt = a.tile_counters + (int64_t)b * a.TX * a.TY * kBinSub + (g & (kBinSub - 1));
            for (int ty = ty0; ty <= ty1; ++ty)
                for (int tx = tx0; tx <= tx1; ++tx)
                    atomicAdd(cnt + (ty * a.TX + tx) * kBinSub,
                              1ull | ((unsigned long long)(unsigned)rect_area_in_tile(rc, tx, ty, a.tile) << 32));
        }
        a.rects[(int64_t)b * a.N + g] = rc;
      }
    }
}

__global__ void __launch_bounds__(256) bin_fill_kernel(const uint2* __restrict__ rects, const float* __restrict__ gauss,
                                                       int rec_f4, unsigned long long* __restrict__ cursor, int B, int N,
                                                       int TX, int TY, int tile, uint4* __restrict__ tile_list,
                                                       unsigned long long list_capacity) {
    const int b = blockIdx.y;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < N; g += gridDim.x * blockDim.x) {
        const uint2 rc = rects[(int64_t)b * N + g];
        if ((rc.x & 0xffff) > (rc.x >> 16)) continue;
        // A 32-byte entry carries the rectangle and the first 16 bytes of the record, [x, y, z, S00 | -s]: the trace reads
        // its candidates with coalesced loads only -- no rects[g] gather (twice per candidate before) and no record
        // gather for isotropic Gaussians (the whole record of kind 1 and of an encoded kind-9 one)
        const float4 head = __ldg(reinterpret_cast<const float4*>(gauss) + (int64_t)g * rec_f4);
        const uint4 e0 = make_uint4((unsigned)g, rc.x, rc.y, 0u);
        const uint4 e1 = make_uint4(__float_as_uint(head.x), __float_as_uint(head.y), __float_as_uint(head.z), __float_as_uint(head.w));
        const int tx0 = (rc.x & 0xffff) / tile, tx1 = (rc.x >> 16) / tile, ty0 = (rc.y & 0xffff) / tile, ty1 = (rc.y >> 16) / tile;
        for (int ty = ty0; ty <= ty1; ++ty)
            for (int tx = tx0; tx <= tx1; ++tx) {
                const int64_t t = (((int64_t)b * TY + ty) * TX + tx) * kBinSub + (g & (kBinSub - 1));
                // the cursor starts at the segment's offset: one atomic yields the list position
                const unsigned long long pos = atomicAdd(cursor + t, 1ull);
                // speculative capacity (sized from the previous call, validated by the host afterwards): never write past it
                if (pos < list_capacity) stg256(tile_list + 2 * pos, e0, e1);      // one 32-byte request
            }
    }
}

// ---- fused backward: d(weight), d(len) -> d(verts), d(sigmas) in ONE kernel ----------------------------
// Recompute, don't store: per pixel the K hits are re-evaluated with the bit-faithful arithmetic
// (so neither act nor dsd is ever written to HBM by the forward), the blend is differentiated
// analytically inside the depth window where the erf is not saturated, and the chain rule of
// ray_trace_voge.cu:324-330 is applied straight into the (N,.) parameter gradients.
constexpr int kMaxPairK = 112;   // two-threads-per-pixel backward up to this K (20 K bytes of shared memory per pixel)

struct FusedBwdArgs {
    const float* gauss;      // packed records (voge_pack_gaussians)
    int kind;                // 1 | 3 | 9, or 9 | kKindIsoEncoded (split into kind / enc by launch_fused_backward)
    const float* origins;
    const float* rays;       // (B,H,W,3), or NULL: generated from `cam` (render_core.cuh: gen_ray)
    const float* cam;        // (B,16) per-view camera records, read when rays == NULL
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
    float* grad_cam;         // optional (B,16), zeroed by the caller, generated rays only: d/d(cam record) = the chain rule of
                             // the ray generator summed over the view's pixels [dR (9), dfx, dfy, dpx, dpy, -, -, -]    // "image mode" (render_bwd_pair_kernel<.., IMG = true>): the upstream gradient arrives on the composited image
    // out = min(sum_k w_k attr[idx_k] + (1 - mask) background, 1) (Renderer.py:153-176, Aggregation.py:111-141)
    // instead of on the weights: dL/dw_k is formed in registers and the attribute gradient is reduced from the same
    // kernel, i.e. merge_final's backward is folded in (no (B,H,W,K) weight-gradient round trip through HBM)
    const float* g_out;      // (B,H,W,C) dL/d(out), C <= 4
    const float* fwd_out;    // (B,H,W,C) the forward's out (tells where min(., 1) saturated), or NULL
    const uint8_t* sat_code; // (B,H,W) optional: voge_merge_final's per-channel clamp code (2 bits each); spares the rebuild
    const float* attr4;      // (N,4) attribute rows padded to 16 bytes
    const float* background; // (C) or NULL (plain interpolate_attr: no silhouette / clamp)
    float mask_thr;          // > 0: hard silhouette mask (Renderer.py:167-168)
    int C;
    float* grad_attr4;       // optional (N,4), zeroed by the caller: dL/d(attr)
    int enc;                 // kind-9 records carry the isotropic encoding
    int attr_in_rec;         // image mode, kind 9: the attribute rows live in the records (voge_pack_attr): one request fetches both
};

// what geom_grad_accumulate needs of the kernel arguments (passed by value to out-of-line callers)
struct GradSink {
    float* grad_packed;
    int kind, need_sigma;
};

template <bool CAM, typename ArgsT>
__device__ __forceinline__ void geom_grad_accumulate(const ArgsT& a, int g, float m0, float m1, float m2,
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

// d/d(ray direction) of the warp's pixels -> d/d(cam record) of their view (all 32 lanes belong to view b): the chain
// rule of the ray generator per contributing lane, a warp reduction per component, 13 atomics per warp.
__device__ __forceinline__ void cam_grad_reduce(const FusedBwdArgs& a, int b, int xi, int yi, bool contributes,
                                                const float* cam_acc, int lane) {
    float o13[13];
#pragma unroll
    for (int q = 0; q < 13; ++q) o13[q] = 0.f;
    if (contributes) {
        const ViewCam v = load_view_cam(a.cam, b);
        gen_ray_backward(v, xi, yi, cam_acc[0], cam_acc[1], cam_acc[2], o13);
    }
#pragma unroll
    for (int q = 0; q < 13; ++q) {
        float v = o13[q];
        for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_down_sync(0xffffffffu, v, sft);
        if (lane == 0 && v != 0.f) atomicAdd(a.grad_cam + 16 * b + q, v);
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
    float d0 = 0.f, d1 = 0.f, d2 = 1.f;
    if (live) pixel_ray(a.rays, a.cam, b, xi, yi, a.H, a.W, d0, d1, d2);
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
                if (g >= 0 && g < a.N) h = exact_hit_packed<KIND>(a.gauss, g, c0, c1, c2, d0, d1, d2, a.enc != 0);
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
                load_gauss<KIND>(a.gauss, g, v0, v1, v2, S, a.enc != 0);
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
        __syncwarp();
        if (a.grad_cam != nullptr) cam_grad_reduce(a, b, xi, yi, live, cam_acc, lane);
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

// Two threads per pixel (adjacent lanes; a warp covers a 4x4 pixel block): the per-pixel arrays in shared
// memory bound the resident pixels per SM, and with one thread per pixel that left 20 warps of serial,
// latency-bound work.  The pair shares the arrays; thread `sub` owns the slots k = sub, sub + 2, ... in every pass.
// torch.minimum-style subgradient of min(x, 1): 1 below, 1/2 at the tie, 0 above (as in blend.cu)
__device__ __forceinline__ float min1_grad_r(float x) { return x < 1.f ? 1.f : (x == 1.f ? 0.5f : 0.f); }

template <int NT, int KIND, bool CAM, bool IMG>
__global__ void __launch_bounds__(NT, (CAM ? 4 : 8)) render_bwd_pair_kernel(const FusedBwdArgs a) {
    constexpr int NP = NT / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const size_t A = (size_t)a.K * NP;
    float2* s_ls = reinterpret_cast<float2*>(smem_raw);   // (len, s = sqrt(dsd + 1e-10))
    float* s_E = reinterpret_cast<float*>(s_ls + A);      // exp(-act)
    float* s_wg = s_E + A;                                // w_m * dL/dw_m
    int* s_g = reinterpret_cast<int*>(s_wg + A);          // local Gaussian index of the slot (-1: none): pass 2 needs no second trip to the index rows
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
    float d0 = 0.f, d1 = 0.f, d2 = 1.f;
    if (live) pixel_ray(a.rays, a.cam, b, xi, yi, a.H, a.W, d0, d1, d2);
    const float c0 = a.origins[3 * b], c1 = a.origins[3 * b + 1], c2 = a.origins[3 * b + 2];
    const float omega = a.omega;
    float cam_acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // d/d(ray), d/d(origin) partial sums of this thread's slots
    const int32_t* i_idx = a.idx + r * a.K;
    const float* i_gw = IMG ? nullptr : a.g_weight + r * a.K;
    const int pack_off = b * a.N;
    const bool vec = (a.K & 3) == 0;

    // ---- image mode: per-pixel upstream gradient of the gather-blend / background composite ----
    // dL/dw_k = sum_c go_c attr[idx_k][c] - gB, with go_c = dL/dout_c through min(., 1) and gB the silhouette term
    // (merge_bwd_small_kernel, blend.cu); both threads of the pair evaluate the same per-pixel values
    float go[4] = {0.f, 0.f, 0.f, 0.f};
    float gB = 0.f;
    if (IMG && live) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < a.C) go[c] = a.g_out[r * a.C + c];
        if (a.background != nullptr) {
            const float* i_w = a.weight + r * a.K;
            float wsum = 0.f;
            for (int k = 0; k < cnt; ++k) wsum += i_w[k];       // slots behind valid_num hold zero weights
            const float sil = fminf(wsum, 1.f);
            const float mask = a.mask_thr > 0.f ? (sil > a.mask_thr ? 1.f : 0.f) : sil;
            bool sat = a.fwd_out == nullptr;
            if (a.sat_code != nullptr) {
                // the forward recorded where min(., 1) clamped (voge_merge_final: sat_code): factor = code / 2 per channel
                const unsigned code = a.sat_code[r];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < a.C) go[c] *= 0.5f * (float)((code >> (2 * c)) & 3u);
                sat = false;
            } else if (!sat) {
                for (int c = 0; c < a.C; ++c) sat = sat || !(a.fwd_out[r * a.C + c] < 1.f);
            }
            if (sat) {
                // min(x, 1) saturated in some channel: the factor needs the un-clamped composite x, rebuilt here
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                for (int k = 0; k < cnt; ++k) {
                    const int g = i_idx[k] - pack_off;
                    if (g >= 0 && g < a.N) {
                        const float4 av = __ldg(reinterpret_cast<const float4*>(a.attr4) + g);
                        const float w = i_w[k];
                        acc[0] = fmaf(w, av.x, acc[0]); acc[1] = fmaf(w, av.y, acc[1]);
                        acc[2] = fmaf(w, av.z, acc[2]); acc[3] = fmaf(w, av.w, acc[3]);
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < a.C) go[c] *= min1_grad_r(acc[c] + (1.f - mask) * a.background[c]);
            }
            if (!(a.mask_thr > 0.f)) {
                float gs = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < a.C) gs += go[c] * a.background[c];
                gB = gs * min1_grad_r(wsum);
            }
        }
    }

    // ---- pass 0: recompute this thread's hits (bit-faithful); w_m dL/dw_m from the forward's weights ----
    float s_min = 3.0e38f, total_gD = 0.f;
    for (int k0 = 0; k0 < cnt; k0 += 4) {
        int gv[2] = {-1, -1};
        float wv[2] = {0.f, 0.f}, gwv[2] = {0.f, 0.f};
        if (vec) {
            const int4 q = *reinterpret_cast<const int4*>(i_idx + k0);
            gv[0] = sub ? q.y : q.x; gv[1] = sub ? q.w : q.z;
            if (!IMG) {
                const float4 g4 = *reinterpret_cast<const float4*>(i_gw + k0);
                gwv[0] = sub ? g4.y : g4.x; gwv[1] = sub ? g4.w : g4.z;
            }
            if (a.weight != nullptr) {
                const float4 w4 = *reinterpret_cast<const float4*>(a.weight + r * a.K + k0);
                wv[0] = sub ? w4.y : w4.x; wv[1] = sub ? w4.w : w4.z;
            }
        } else {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int k = k0 + sub + 2 * jj;
                if (k < cnt) {
                    gv[jj] = i_idx[k];
                    if (!IMG) gwv[jj] = i_gw[k];
                    if (a.weight != nullptr) wv[jj] = a.weight[r * a.K + k];
                }
            }
        }
        // kind-9 tables with the attribute rows in the records: geometry + attribute of a slot arrive with one request
        const bool rec_attr = IMG && KIND == 9 && a.attr_in_rec != 0;
        if (IMG && !rec_attr) {
            // the attribute rows of this thread's two slots: dL/dw from the image gradient, dL/d(attr) reduced here
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int k = k0 + sub + 2 * jj;
                const int g = gv[jj] - pack_off;
                if (k < cnt && g >= 0 && g < a.N) {
                    const float4 av = __ldg(reinterpret_cast<const float4*>(a.attr4) + g);
                    gwv[jj] = fmaf(go[3], av.w, fmaf(go[2], av.z, fmaf(go[1], av.y, go[0] * av.x))) - gB;
                    if (a.grad_attr4 != nullptr && wv[jj] != 0.f)
                        atomicAdd(reinterpret_cast<float4*>(a.grad_attr4) + g,
                                  make_float4(wv[jj] * go[0], wv[jj] * go[1], wv[jj] * go[2], wv[jj] * go[3]));
                } else {
                    gwv[jj] = -gB;
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
                if (rec_attr) gwv[jj] = -gB;
                if (g >= 0 && g < a.N) {
                    if (rec_attr) {
                        float v0, v1, v2, S[9];
                        float4 av;
                        load_gauss_attr9(a.gauss, g, v0, v1, v2, S, av, a.enc != 0);
                        gwv[jj] = fmaf(go[3], av.w, fmaf(go[2], av.z, fmaf(go[1], av.y, go[0] * av.x))) - gB;
                        if (a.grad_attr4 != nullptr && wv[jj] != 0.f)
                            atomicAdd(reinterpret_cast<float4*>(a.grad_attr4) + g,
                                      make_float4(wv[jj] * go[0], wv[jj] * go[1], wv[jj] * go[2], wv[jj] * go[3]));
                        h = exact_pair(__fsub_rn(v0, c0), __fsub_rn(v1, c1), __fsub_rn(v2, c2), S, d0, d1, d2);
                    } else {
                        h = exact_hit_packed<KIND>(a.gauss, g, c0, c1, c2, d0, d1, d2, a.enc != 0);
                    }
                }
                const float sk = sqrtf(h.dsd + 1e-10f);
                s_ls[k * NP + col] = make_float2(h.len, sk);
                s_E[k * NP + col] = expf(-h.act);
                s_g[k * NP + col] = (g >= 0 && g < a.N) ? g : -1;
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
        for (int j = sub; j < cnt; j += 2) {
            // the Gaussian's record is fetched first: the gathers overlap with the window loop below
            const int g = s_g[j * NP + col];
            const bool g_ok = g >= 0;
            float S[9], m0 = 0.f, m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int q = 0; q < 9; ++q) S[q] = 0.f;
            if (g_ok) {
                float v0, v1, v2;
                load_gauss<KIND>(a.gauss, g, v0, v1, v2, S, a.enc != 0);
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
        if (a.grad_rays != nullptr || a.grad_cam != nullptr) {
#pragma unroll
            for (int q = 0; q < 3; ++q) cam_acc[q] += __shfl_xor_sync(0xffffffffu, cam_acc[q], 1);
            if (a.grad_rays != nullptr && live && sub == 0) {
                a.grad_rays[r * 3 + 0] = cam_acc[0]; a.grad_rays[r * 3 + 1] = cam_acc[1]; a.grad_rays[r * 3 + 2] = cam_acc[2];
            }
            if (a.grad_cam != nullptr) cam_grad_reduce(a, b, xi, yi, live && sub == 0, cam_acc, lane);
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

namespace voge {
static int launch_fused_backward(const FusedBwdArgs& a, bool image_mode, int flags, cudaStream_t s);
}

extern "C" int voge_render_backward_fused(const float* gauss, int sigma_kind,
                                          const float* origins, const float* rays, const int32_t* idx,
                                          const int64_t* valid, const float* grad_weight, const float* weight,
                                          const float* grad_len_out, float absorptivity, int B, int N, int H,
                                          int W, int K, float* grad_packed, int need_sigma, float* grad_rays,
                                          float* grad_origins, const float* cam, float* grad_cam,
                                          voge_stream_t stream) {
    using namespace voge;
    if (B <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
    if (rays == nullptr && cam == nullptr) return (int)cudaErrorInvalidValue;
    if (grad_cam != nullptr && (rays != nullptr || cam == nullptr)) return (int)cudaErrorInvalidValue;
    FusedBwdArgs a{gauss, sigma_kind, origins, rays, cam, idx, valid, grad_weight, weight, grad_len_out, absorptivity,
                   B, N, H, W, K, grad_packed, need_sigma, grad_rays, grad_origins, grad_cam,
                   nullptr, nullptr, nullptr, nullptr, nullptr, -1.f, 0, nullptr, 0, 0};
    a.need_sigma = need_sigma & 1;
    return launch_fused_backward(a, false, need_sigma, (cudaStream_t)stream);
}

extern "C" int voge_render_backward_image(const float* gauss, int sigma_kind, const float* origins, const float* rays,
                                          const float* cam, const int32_t* idx, const int64_t* valid,
                                          const float* weight, const float* grad_out, const float* fwd_out,
                                          const uint8_t* sat_code, const float* attr4, const float* background, float mask_thr, int C,
                                          float absorptivity, int B, int N, int H, int W, int K, float* grad_packed,
                                          int need_sigma, float* grad_attr4, float* grad_rays, float* grad_origins,
                                          float* grad_cam, voge_stream_t stream) {
    using namespace voge;
    if (B <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
    if (rays == nullptr && cam == nullptr) return (int)cudaErrorInvalidValue;
    if (grad_cam != nullptr && (rays != nullptr || cam == nullptr)) return (int)cudaErrorInvalidValue;
    if (weight == nullptr || grad_out == nullptr || attr4 == nullptr || C < 1 || C > 4 || K > kMaxPairK)
        return (int)cudaErrorInvalidValue;
    FusedBwdArgs a{gauss, sigma_kind, origins, rays, cam, idx, valid, nullptr, weight, nullptr, absorptivity,
                   B, N, H, W, K, grad_packed, need_sigma, grad_rays, grad_origins, grad_cam,
                   grad_out, fwd_out, sat_code, attr4, background, mask_thr, C, grad_attr4, 0, 0};
    a.need_sigma = need_sigma & 1;
    a.attr_in_rec = (need_sigma & 2) ? 1 : 0;
    return launch_fused_backward(a, true, need_sigma, (cudaStream_t)stream);
}

namespace voge {
// flags: bit 0 = sigma gradients wanted
static int launch_fused_backward(const FusedBwdArgs& a_in, bool image_mode, int flags, cudaStream_t s) {
    FusedBwdArgs a = a_in;
    a.enc = (a.kind & kKindIsoEncoded) ? 1 : 0;
    a.kind &= ~kKindIsoEncoded;
    if (a.enc && a.kind != 9) return (int)cudaErrorInvalidValue;
    const int sigma_kind = a.kind, B = a.B, H = a.H, W = a.W, K = a.K;
    float* grad_rays = a.grad_rays; float* grad_origins = a.grad_origins; float* grad_cam = a.grad_cam;
    if (sigma_kind != 1 && sigma_kind != 3 && sigma_kind != 9) return (int)cudaErrorInvalidValue;
    const bool cam_grads = grad_rays != nullptr || grad_origins != nullptr || grad_cam != nullptr;
    auto launch = [&](auto kernel, int nt, int per_pixel) -> int {
        // per_pixel = 2: two threads per pixel, 4x4 pixel blocks per warp; 1: 8x4 blocks
        const int64_t warps = per_pixel == 2 ? (int64_t)B * cdiv(W, 4) * cdiv(H, 4) : (int64_t)B * cdiv(W, 8) * cdiv(H, 4);
        const size_t smem = (size_t)K * (nt / per_pixel) * (per_pixel == 2 ? 20 : 16);
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
        if (image_mode) return launch(render_bwd_pair_kernel<128, KIND, CAM, true>, 128, 2);
        if (K <= kMaxPairK) return launch(render_bwd_pair_kernel<128, KIND, CAM, false>, 128, 2);
        if (K <= 200) return launch(render_bwd_fused_kernel<64, KIND, CAM>, 64, 1);
        return launch(render_bwd_fused_kernel<32, KIND, CAM>, 32, 1);
    };
    auto by_kind = [&](auto cam_tag) -> int {
        if (sigma_kind == 1) return by_threads(std::integral_constant<int, 1>{}, cam_tag);
        if (sigma_kind == 3) return by_threads(std::integral_constant<int, 3>{}, cam_tag);
        return by_threads(std::integral_constant<int, 9>{}, cam_tag);
    };
    return cam_grads ? by_kind(std::true_type{}) : by_kind(std::false_type{});
}
}  // namespace voge

namespace voge {
// ---- parameter records ---------------------------------------------------------------------------------------
// sigma_mode: how the renderer's `sigmas` argument maps to the matrix P of S = 2 P (reference Renderer.py:134-137):
//   0  P = sigmas                     (inverse covariances given, inverse_sigma = False)
//   1  P = inverse(sigmas)            (covariances given, inverse_sigma = True: `2 * torch.inverse(sigmas)`)
//   2  P = tril(L) tril(L)^T          (Cholesky factor given: `to_sym`, demo/EfficientCuboidViaOptimization.py:17-18)
// The 3x3 inverse is the adjugate over the determinant (the reference calls torch.inverse = batched LU; both are
// backward-stable, the results differ in the last bits like two LU implementations do).
__device__ __forceinline__ void sigma_to_S(int kind, int mode, const float* __restrict__ sig, int g, float* S) {
    if (mode == 0) {
        load_S_dyn(kind, sig, g, S);
        return;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) S[i] = 0.f;
    if (kind == 1) {
        const float v = 2.f / __ldg(sig + g);
        S[0] = v; S[4] = v; S[8] = v;
    } else if (kind == 3) {
        S[0] = 2.f / __ldg(sig + 3 * (int64_t)g);
        S[4] = 2.f / __ldg(sig + 3 * (int64_t)g + 1);
        S[8] = 2.f / __ldg(sig + 3 * (int64_t)g + 2);
    } else {
        float m[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) m[i] = __ldg(sig + 9 * (int64_t)g + i);
        if (mode == 1) {
            const float c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
            const float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
            const float t = 2.f / det;
            S[0] = c00 * t; S[1] = (m[2] * m[7] - m[1] * m[8]) * t; S[2] = (m[1] * m[5] - m[2] * m[4]) * t;
            S[3] = c01 * t; S[4] = (m[0] * m[8] - m[2] * m[6]) * t; S[5] = (m[2] * m[3] - m[0] * m[5]) * t;
            S[6] = c02 * t; S[7] = (m[1] * m[6] - m[0] * m[7]) * t; S[8] = (m[0] * m[4] - m[1] * m[3]) * t;
        } else {
            const float l00 = m[0], l10 = m[3], l11 = m[4], l20 = m[6], l21 = m[7], l22 = m[8];
            S[0] = 2.f * (l00 * l00);
            S[1] = S[3] = 2.f * (l00 * l10);
            S[2] = S[6] = 2.f * (l00 * l20);
            S[4] = 2.f * (l10 * l10 + l11 * l11);
            S[5] = S[7] = 2.f * (l10 * l20 + l11 * l21);
            S[8] = 2.f * (l20 * l20 + l21 * l21 + l22 * l22);
        }
    }
}

// iso_flag != NULL (kind 9): isotropic encoding (render_core.cuh: kKindIsoEncoded) -- a record whose S is exactly
// s I, s > 0, is written with -s in the S00 slot; any OTHER record with the sign bit set in S00 (a non positive
// definite input) makes the sign ambiguous and is reported through *iso_flag = 1: the caller then packs plain records.
__global__ void __launch_bounds__(256) pack_gaussians_kernel(const float* __restrict__ verts, const float* __restrict__ sigmas,
                                                             int kind, int mode, int N, float* __restrict__ out,
                                                             int32_t* __restrict__ iso_flag) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    const float x = verts[3 * (int64_t)g], y = verts[3 * (int64_t)g + 1], z = verts[3 * (int64_t)g + 2];
    float S[9];
    sigma_to_S(kind, mode, sigmas, g, S);               // mode 0: S = 2 sigma (Renderer.py:137), exact in fp32
    float4* o = reinterpret_cast<float4*>(out);
    if (kind == 1) {
        o[g] = make_float4(x, y, z, S[0]);
    } else if (kind == 3) {
        o[2 * (int64_t)g] = make_float4(x, y, z, S[0]);
        o[2 * (int64_t)g + 1] = make_float4(S[4], S[8], 0.f, 0.f);
    } else {
        if (iso_flag != nullptr) {
            const unsigned off = __float_as_uint(S[1]) | __float_as_uint(S[2]) | __float_as_uint(S[3]) |
                                 __float_as_uint(S[5]) | __float_as_uint(S[6]) | __float_as_uint(S[7]);
            const bool iso = off == 0u && S[4] == S[0] && S[8] == S[0] && S[0] > 0.f && S[0] < 3.0e38f;
            if (iso) S[0] = -S[0];
            else if (__float_as_uint(S[0]) >> 31) *iso_flag = 1;
        }
        o[4 * (int64_t)g] = make_float4(x, y, z, S[0]);
        o[4 * (int64_t)g + 1] = make_float4(0.f, 0.f, 0.f, 0.f);          // attribute row, voge_pack_attr
        o[4 * (int64_t)g + 2] = make_float4(S[1], S[2], S[3], S[4]);
        o[4 * (int64_t)g + 3] = make_float4(S[5], S[6], S[7], S[8]);
    }
}

// Gradient epilogue: the fused backward accumulates one packed record per Gaussian, [d verts (3) | d P] with
// d P = dL/dP (P as above; the factor 2 of S = 2 P is already applied).  This kernel splits the records into the
// caller's tensors and applies the chain rule of the sigma parameterisation:
//   mode 1:  dL/dSigma = -P^T (dL/dP) P^T   (P = S / 2 read back from the Gaussian's record)
//   mode 2:  dL/dL     = tril((G + G^T) tril(L))
__global__ void __launch_bounds__(256) unpack_gradients_kernel(const float* __restrict__ packed, const float* __restrict__ gauss,
                                                               const float* __restrict__ sigmas, int kind, int enc, int mode, int N,
                                                               float* __restrict__ g_verts, float* __restrict__ g_sigmas) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    const int width = kind == 9 ? 12 : (kind == 3 ? 8 : 4);
    const float4* pk = reinterpret_cast<const float4*>(packed) + (int64_t)g * (width / 4);
    const float4 q0 = pk[0];
    g_verts[3 * (int64_t)g] = q0.x; g_verts[3 * (int64_t)g + 1] = q0.y; g_verts[3 * (int64_t)g + 2] = q0.z;
    if (g_sigmas == nullptr) return;
    if (kind == 1) {
        float v = q0.w;
        if (mode == 1) { const float p = 1.f / __ldg(sigmas + g); v = -v * p * p; }
        g_sigmas[g] = v;
    } else if (kind == 3) {
        const float4 q1 = pk[1];
        float v[3] = {q1.x, q1.y, q1.z};
        if (mode == 1) {
#pragma unroll
            for (int i = 0; i < 3; ++i) { const float p = 1.f / __ldg(sigmas + 3 * (int64_t)g + i); v[i] = -v[i] * p * p; }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) g_sigmas[3 * (int64_t)g + i] = v[i];
    } else {
        const float4 q1 = pk[1], q2 = pk[2];
        const float G[9] = {q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
        float o[9];
        if (mode == 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) o[i] = G[i];
        } else if (mode == 1) {
            float v0, v1, v2, S[9], P[9];
            load_gauss<9>(gauss, g, v0, v1, v2, S, enc != 0);
#pragma unroll
            for (int i = 0; i < 9; ++i) P[i] = 0.5f * S[i];
            float T[9];   // T = P^T G
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) T[3 * i + j] = P[i] * G[j] + P[3 + i] * G[3 + j] + P[6 + i] * G[6 + j];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) o[3 * i + j] = -(T[3 * i] * P[3 * j] + T[3 * i + 1] * P[3 * j + 1] + T[3 * i + 2] * P[3 * j + 2]);
        } else {
            float L[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) L[i] = __ldg(sigmas + 9 * (int64_t)g + i);
            L[1] = L[2] = L[5] = 0.f;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    float acc = 0.f;
#pragma unroll
                    for (int k = 0; k < 3; ++k) acc += (G[3 * i + k] + G[3 * k + i]) * L[3 * k + j];
                    o[3 * i + j] = j <= i ? acc : 0.f;
                }
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) g_sigmas[9 * (int64_t)g + i] = o[i];
    }
}

__global__ void __launch_bounds__(256) generate_rays_kernel(const float* __restrict__ cam, int B, int H, int W,
                                                            float* __restrict__ rays) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * H * W) return;
    const int x = (int)(t % W), y = (int)((t / W) % H), b = (int)(t / ((int64_t)W * H));
    const ViewCam v = load_view_cam(cam, b);
    float d0, d1, d2;
    gen_ray(v, x, y, d0, d1, d2);
    rays[3 * t] = d0; rays[3 * t + 1] = d1; rays[3 * t + 2] = d2;
}
}  // namespace voge

extern "C" int voge_pack_gaussians(const float* verts, const float* sigmas, int sigma_kind, int sigma_mode, int N,
                                   float* out, int32_t* iso_flag, voge_stream_t stream) {
    using namespace voge;
    if (N <= 0) return 0;
    if (sigma_kind != 1 && sigma_kind != 3 && sigma_kind != 9) return (int)cudaErrorInvalidValue;
    if (sigma_mode < 0 || sigma_mode > 2 || (sigma_mode == 2 && sigma_kind != 9)) return (int)cudaErrorInvalidValue;
    if (iso_flag != nullptr && sigma_kind != 9) return (int)cudaErrorInvalidValue;
    pack_gaussians_kernel<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(verts, sigmas, sigma_kind, sigma_mode, N, out, iso_flag);
    VOGE_LAUNCH_CHECK();
    return 0;
}

extern "C" int voge_unpack_gradients(const float* grad_packed, const float* gauss, const float* sigmas, int sigma_kind,
                                     int sigma_mode, int N, float* grad_verts, float* grad_sigmas, voge_stream_t stream) {
    using namespace voge;
    if (N <= 0) return 0;
    const int enc = (sigma_kind & kKindIsoEncoded) ? 1 : 0;
    sigma_kind &= ~kKindIsoEncoded;
    if (sigma_kind != 1 && sigma_kind != 3 && sigma_kind != 9) return (int)cudaErrorInvalidValue;
    if (enc && sigma_kind != 9) return (int)cudaErrorInvalidValue;
    if (sigma_mode < 0 || sigma_mode > 2 || (sigma_mode == 2 && sigma_kind != 9)) return (int)cudaErrorInvalidValue;
    unpack_gradients_kernel<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(grad_packed, gauss, sigmas, sigma_kind, enc,
                                                                           sigma_mode, N, grad_verts, grad_sigmas);
    VOGE_LAUNCH_CHECK();
    return 0;
}

extern "C" int voge_generate_rays(const float* cam, int B, int H, int W, float* rays, voge_stream_t stream) {
    using namespace voge;
    const int64_t total = (int64_t)B * H * W;
    if (total <= 0) return 0;
    generate_rays_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cam, B, H, W, rays);
    VOGE_LAUNCH_CHECK();
    return 0;
}

namespace voge {
// attribute rows (N, C <= 4) -> zero-padded (N,4) table (one 16-byte gather per hit in the gather-blend kernels) and,
// for a kind-9 record table, the second 16 bytes of every record (render_core.cuh)
__global__ void __launch_bounds__(256) pack_attr_kernel(const float* __restrict__ attr, int C, int N,
                                                        float* __restrict__ attr4, float* __restrict__ gauss16) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) v[c] = attr[(int64_t)g * C + c];
    const float4 q = make_float4(v[0], v[1], v[2], v[3]);
    if (attr4 != nullptr) reinterpret_cast<float4*>(attr4)[g] = q;
    if (gauss16 != nullptr) reinterpret_cast<float4*>(gauss16)[4 * (int64_t)g + 1] = q;
}
}  // namespace voge

extern "C" int voge_pack_attr(const float* attr, int C, int N, float* attr4, float* gauss16, voge_stream_t stream) {
    using namespace voge;
    if (N <= 0) return 0;
    if (C < 1 || C > 4) return (int)cudaErrorInvalidValue;
    pack_attr_kernel<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(attr, C, N, attr4, gauss16);
    VOGE_LAUNCH_CHECK();
    return 0;
}

extern "C" int voge_bin_sub(void) { return voge::kBinSub; }

extern "C" int voge_bin_count(const float* gauss, int sigma_kind, const float* Rm,
                              const float* Tv, const float* origins, const float* focal, const float* principal,
                              int B, int N, int H, int W, float thr, float thr_act, int use_ref_bins, int bin_size,
                              int tile, int flags, uint32_t* rects, uint64_t* tile_counters,
                              voge_stream_t stream) {
    using namespace voge;
    if (B <= 0 || N <= 0) return 0;
    if (tile <= 0 || tile > 16 || (use_ref_bins && (bin_size <= 0 || bin_size % tile != 0))) return (int)cudaErrorInvalidValue;
    const int enc = (sigma_kind & kKindIsoEncoded) ? 1 : 0;
    sigma_kind &= ~kKindIsoEncoded;
    if (sigma_kind != 1 && sigma_kind != 3 && sigma_kind != 9) return (int)cudaErrorInvalidValue;
    if (enc && sigma_kind != 9) return (int)cudaErrorInvalidValue;
    BinArgs a;
    a.enc = enc;
    a.gauss = gauss; a.kind = sigma_kind; a.Rm = Rm; a.Tv = Tv; a.origins = origins;
    a.focal = focal; a.principal = principal; a.B = B; a.N = N; a.H = H; a.W = W;
    a.neg_log_thr = -logf(thr); a.thr_act = thr_act; a.use_ref_bins = use_ref_bins; a.bin_size = bin_size;
    a.BH = use_ref_bins ? cdiv(H, bin_size) : 1; a.BW = use_ref_bins ? cdiv(W, bin_size) : 1;
    a.tile = tile; a.TX = cdiv(W, tile); a.TY = cdiv(H, tile);
    a.zero_aware_margin = (flags & 1) ? 0 : 1;
    if (a.TX > 65535 || a.TY > 65535) return (int)cudaErrorInvalidValue;
    a.rects = reinterpret_cast<uint2*>(rects); a.tile_counters = reinterpret_cast<unsigned long long*>(tile_counters);
    // one Gaussian per thread, walking B / grid.y views: as few view slices as still fill the machine (~4 waves of blocks)
    const int bx = cdiv(N, 256);
    const int by = min(B, max(1, cdiv(4 * 148 * 8, bx)));
    dim3 grid(bx, by);
    bin_count_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    VOGE_LAUNCH_CHECK();
    return 0;
}

extern "C" int voge_bin_fill(const uint32_t* rects, const float* gauss, int sigma_kind, uint64_t* cursor, int B, int N,
                             int H, int W, int tile, int32_t* tile_list, int64_t list_capacity, voge_stream_t stream) {
    using namespace voge;
    if (B <= 0 || N <= 0) return 0;
    sigma_kind &= ~kKindIsoEncoded;
    if (sigma_kind != 1 && sigma_kind != 3 && sigma_kind != 9) return (int)cudaErrorInvalidValue;
    const int rec_f4 = sigma_kind == 9 ? 4 : (sigma_kind == 3 ? 2 : 1);
    dim3 grid(cdiv(N, 256), B);   // one Gaussian per thread: the dependent load -> atomic -> store chain is pure latency
    bin_fill_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint2*>(rects), gauss, rec_f4,
                                                             reinterpret_cast<unsigned long long*>(cursor),
                                                             B, N, cdiv(W, tile), cdiv(H, tile), tile,
                                                             reinterpret_cast<uint4*>(tile_list),
                                                             list_capacity > 0 ? (unsigned long long)list_capacity : ~0ull);
    VOGE_LAUNCH_CHECK();
    return 0;
}

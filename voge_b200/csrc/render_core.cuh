// render_core.cuh -- helpers shared by the fused renderer kernels (render.cu: binning, backward;
// trace.cu: forward pipeline): compact-sigma access, exact_pair specialisations, tile pixel layout.
#pragma once
#include "common.cuh"

namespace voge {

// ---- parameter access ----------------------------------------------------------------------------
// sigma kinds: 1 = (N,) isotropic, 3 = (N,3) diagonal, 9 = (N,3,3) full.  S = 2 * sigma (Renderer.py:137)
template <int KIND>
__device__ __forceinline__ void load_S(const float* __restrict__ sig, int g, float* S) {
    if (KIND == 1) {
        const float s = 2.f * __ldg(sig + g);
        S[0] = s; S[1] = 0.f; S[2] = 0.f; S[3] = 0.f; S[4] = s; S[5] = 0.f; S[6] = 0.f; S[7] = 0.f; S[8] = s;
    } else if (KIND == 3) {
        S[0] = 2.f * __ldg(sig + 3 * (int64_t)g); S[4] = 2.f * __ldg(sig + 3 * (int64_t)g + 1);
        S[8] = 2.f * __ldg(sig + 3 * (int64_t)g + 2);
        S[1] = S[2] = S[3] = S[5] = S[6] = S[7] = 0.f;
    } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) S[i] = 2.f * __ldg(sig + 9 * (int64_t)g + i);
    }
}

__device__ __forceinline__ void load_S_dyn(int kind, const float* __restrict__ sig, int g, float* S) {
    if (kind == 1) load_S<1>(sig, g, S);
    else if (kind == 3) load_S<3>(sig, g, S);
    else load_S<9>(sig, g, S);
}

// exact_pair specialised for a diagonal S: identical bits to exact_pair with explicit zeros
// (fma(0, c, acc) == acc), at a third of the work.
__device__ __forceinline__ Hit exact_pair_diag(float m0, float m1, float m2, float s0, float s1, float s2,
                                               float d0, float d1, float d2) {
    const float t0 = __fmul_rn(d0, s0), t1 = __fmul_rn(d1, s1), t2 = __fmul_rn(d2, s2);
    const float u0 = __fmul_rn(m0, s0), u1 = __fmul_rn(m1, s1), u2 = __fmul_rn(m2, s2);
    const float ksk = __fmaf_rn(t2, d2, __fmaf_rn(t1, d1, __fmul_rn(t0, d0)));
    const float msk = __fmaf_rn(u2, d2, __fmaf_rn(u1, d1, __fmul_rn(u0, d0)));
    const float msm = __fmaf_rn(u2, m2, __fmaf_rn(u1, m1, __fmul_rn(u0, m0)));
    Hit h;
    h.len = __fdiv_rn(msk, ksk);
    h.act = __fsub_rn(msm, __fdiv_rn(__fmul_rn(msk, msk), ksk));
    h.dsd = ksk;
    return h;
}

template <int KIND>
__device__ __forceinline__ Hit exact_hit(const float* __restrict__ verts, const float* __restrict__ sig, int g,
                                         float c0, float c1, float c2, float d0, float d1, float d2) {
    const float m0 = __fsub_rn(__ldg(verts + 3 * (int64_t)g), c0);       // verts - ray_origin, Renderer.py:130
    const float m1 = __fsub_rn(__ldg(verts + 3 * (int64_t)g + 1), c1);
    const float m2 = __fsub_rn(__ldg(verts + 3 * (int64_t)g + 2), c2);
    if (KIND == 1) {
        const float s = 2.f * __ldg(sig + g);
        return exact_pair_diag(m0, m1, m2, s, s, s, d0, d1, d2);
    } else if (KIND == 3) {
        return exact_pair_diag(m0, m1, m2, 2.f * __ldg(sig + 3 * (int64_t)g), 2.f * __ldg(sig + 3 * (int64_t)g + 1),
                               2.f * __ldg(sig + 3 * (int64_t)g + 2), d0, d1, d2);
    } else {
        float S[9];
        load_S<9>(sig, g, S);
        return exact_pair(m0, m1, m2, S, d0, d1, d2);
    }
}

// floor(p / w) = (p * kInvW[w]) >> 16 exactly for p < 256, 1 <= w <= 16
static __constant__ unsigned kInvW[17] = {0u, 65536u, 32768u, 21846u, 16384u, 13108u, 10923u, 9363u, 8192u,
                                   7282u, 6554u, 5958u, 5462u, 5042u, 4682u, 4370u, 4096u};

template <int NT>
__device__ __forceinline__ int pix_to_col(int lx, int ly, int tile) {
    if (NT == 256 && tile == 16) return (((ly >> 2) * 2 + (lx >> 3)) << 5) + ((ly & 3) << 3) + (lx & 7);
    return ly * tile + lx;
}


// inverse of pix_to_col: pixel (lx, ly) of column `col`
template <int NT>
__device__ __forceinline__ void col_to_pix(int col, int tile, int& lx, int& ly, bool& in_tile) {
    if (NT == 256 && tile == 16) {
        const int w = col >> 5, l = col & 31;
        lx = (w & 1) * 8 + (l & 7);
        ly = (w >> 1) * 4 + (l >> 3);
        in_tile = true;
    } else {
        lx = col % tile; ly = col / tile;
        in_tile = col < tile * tile;
    }
}

// Pixels of the rectangle rc (x0|x1<<16, y0|y1<<16, inclusive, inside the image) that lie in tile (tx,ty):
// the number of ITEMS the entry contributes to the tile (trace.cu).  0 if the intersection is empty.
__host__ __device__ __forceinline__ int rect_area_in_tile(uint2 rc, int tx, int ty, int tile) {
    const int xl = max((int)(rc.x & 0xffffu), tx * tile), xh = min((int)(rc.x >> 16), tx * tile + tile - 1);
    const int yl = max((int)(rc.y & 0xffffu), ty * tile), yh = min((int)(rc.y >> 16), ty * tile + tile - 1);
    return max(xh - xl + 1, 0) * max(yh - yl + 1, 0);
}

// ---- packed per-Gaussian records (voge_pack_gaussians) ------------------------------------------------
// One aligned record per Gaussian, [x, y, z | S = 2 sigma ...]: kind 1 -> 4 floats (x,y,z,s), kind 3 -> 8 floats
// (x,y,z,s0 | s1,s2,-,-), kind 9 -> 16 floats = 64 bytes (x,y,z,S00 | attribute row | S01,S02,S10,S11 | S12..S22).  A hit
// then costs 1 / 2 / 3 vector loads instead of 4 / 6 / 12 scalar ones (every lane of a warp gathers a different
// Gaussian, so each load instruction is one L1 tag lookup per lane).  The second 16 bytes of a kind-9 record hold the
// Gaussian's ATTRIBUTE row (colour; written by voge_pack_attr when an image is composited from the fragments): the
// image-mode backward fetches geometry + attribute of an isotropic Gaussian with ONE 32-byte request (sm_100's
// 256-bit ld.global.v8.f32) instead of two gathers into two tables.
template <int KIND>
struct GaussWidth {
    static constexpr int v = (KIND == 9) ? 16 : (KIND == 3 ? 8 : 4);
};

// 256-bit read-only load (LDG.E.ENL2.256.CONSTANT, sm_100+): p must be 32-byte aligned
__device__ __forceinline__ void ldg256(const float* __restrict__ p, float4& lo, float4& hi) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                 : "l"(p));
}

// Isotropic encoding of kind-9 records (sigma_kind | kKindIsoEncoded): a Gaussian whose S is exactly s I, s > 0
// (every off-diagonal entry +0), carries -s in the S00 slot, so that its whole record is the first 16 bytes
// [x, y, z, -s] and a hit gathers ONE sector instead of the 48-byte record (the scattered gathers bound the blend and
// backward kernels).  The readers rebuild the same nine floats, i.e. the same bits go through exact_pair.
// voge_pack_gaussians encodes only when no other record has the sign bit set in S00 (it reports such records and
// the caller falls back to plain records), so the sign is unambiguous.
constexpr int kKindIsoEncoded = 0x100;

template <int KIND>
__device__ __forceinline__ void load_gauss(const float* __restrict__ gp, int g, float& v0, float& v1, float& v2, float* S,
                                           bool enc = false) {
    const float4* p = reinterpret_cast<const float4*>(gp) + (int64_t)g * (GaussWidth<KIND>::v / 4);
    const float4 a = __ldg(p);
    v0 = a.x; v1 = a.y; v2 = a.z;
    if (KIND == 1) {
        S[0] = a.w; S[4] = a.w; S[8] = a.w;
        S[1] = S[2] = S[3] = S[5] = S[6] = S[7] = 0.f;
    } else if (KIND == 3) {
        const float4 b = __ldg(p + 1);
        S[0] = a.w; S[4] = b.x; S[8] = b.y;
        S[1] = S[2] = S[3] = S[5] = S[6] = S[7] = 0.f;
    } else if (enc && a.w < 0.f) {
        const float s = -a.w;
        S[0] = s; S[4] = s; S[8] = s;
        S[1] = S[2] = S[3] = S[5] = S[6] = S[7] = 0.f;
    } else {
        const float4 b = __ldg(p + 2), c = __ldg(p + 3);
        S[0] = a.w; S[1] = b.x; S[2] = b.y; S[3] = b.z; S[4] = b.w; S[5] = c.x; S[6] = c.y; S[7] = c.z; S[8] = c.w;
    }
}

// 256-bit store (STG.E.ENL2.256, sm_100+): p must be 32-byte aligned
__device__ __forceinline__ void stg256(void* p, const uint4 lo, const uint4 hi) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
}
__device__ __forceinline__ void ldg256u(const void* __restrict__ p, uint4& lo, uint4& hi) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
}

// the rest of a record whose first 16 bytes `a` = [x, y, z, S00 | -s] are already at hand (trace.cu: they travel in the
// tile-list entry): nothing for kind 1 and for an isotropically encoded kind-9 record, one gather otherwise
template <int KIND>
__device__ __forceinline__ void finish_gauss(const float* __restrict__ gp, int g, const float4 a, float* S, bool enc) {
    if (KIND == 1) {
        S[0] = a.w; S[4] = a.w; S[8] = a.w;
        S[1] = S[2] = S[3] = S[5] = S[6] = S[7] = 0.f;
    } else if (KIND == 3) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(gp) + (int64_t)g * 2 + 1);
        S[0] = a.w; S[4] = b.x; S[8] = b.y;
        S[1] = S[2] = S[3] = S[5] = S[6] = S[7] = 0.f;
    } else if (enc && a.w < 0.f) {
        const float s = -a.w;
        S[0] = s; S[4] = s; S[8] = s;
        S[1] = S[2] = S[3] = S[5] = S[6] = S[7] = 0.f;
    } else {
        float4 b, c;
        ldg256(gp + (int64_t)g * 16 + 8, b, c);
        S[0] = a.w; S[1] = b.x; S[2] = b.y; S[3] = b.z; S[4] = b.w; S[5] = c.x; S[6] = c.y; S[7] = c.z; S[8] = c.w;
    }
}

// kind-9 record + the attribute row stored in it: one 256-bit request for an isotropic Gaussian, two for a dense one
__device__ __forceinline__ void load_gauss_attr9(const float* __restrict__ gp, int g, float& v0, float& v1, float& v2,
                                                 float* S, float4& attr, bool enc) {
    const float* p = gp + (int64_t)g * 16;
    float4 a;
    ldg256(p, a, attr);
    v0 = a.x; v1 = a.y; v2 = a.z;
    if (enc && a.w < 0.f) {
        const float s = -a.w;
        S[0] = s; S[4] = s; S[8] = s;
        S[1] = S[2] = S[3] = S[5] = S[6] = S[7] = 0.f;
    } else {
        float4 b, c;
        ldg256(p + 8, b, c);
        S[0] = a.w; S[1] = b.x; S[2] = b.y; S[3] = b.z; S[4] = b.w; S[5] = c.x; S[6] = c.y; S[7] = c.z; S[8] = c.w;
    }
}

template <int KIND>
__device__ __forceinline__ Hit exact_hit_packed(const float* __restrict__ gp, int g, float c0, float c1, float c2,
                                                float d0, float d1, float d2, bool enc = false) {
    float v0, v1, v2, S[9];
    load_gauss<KIND>(gp, g, v0, v1, v2, S, enc);
    const float m0 = __fsub_rn(v0, c0), m1 = __fsub_rn(v1, c1), m2 = __fsub_rn(v2, c2);   // verts - ray_origin, Renderer.py:130
    if (KIND == 9) return exact_pair(m0, m1, m2, S, d0, d1, d2);
    return exact_pair_diag(m0, m1, m2, S[0], S[4], S[8], d0, d1, d2);
}

// ---- closed-form ray generator (reference Renderer.py:124-128: NDCMultinomialRaysampler with unit directions on a
// screen-space PerspectiveCameras; semantics SURVEY.md 8c) -------------------------------------------------
// d = R . normalize(a, b, 1),  a = -(x + .5 - px) / fx,  b = -(y + .5 - py) / fy   (pytorch3d view frame: +X left,
// +Y up, +Z forward; row-vector convention X_view = X_world R + T, so d_world = R d_cam).
// ONE definition, written with explicit roundings, shared by voge_generate_rays (materialised (B,H,W,3) rays for
// the op-by-op ops and the oracle) and by the fused kernels (rays never materialised): every kernel sees the same
// bits for the same pixel.  Per-view camera record `cam` (B,16) f32 = [R row-major (9), fx, fy, px, py, 0, 0, 0].
struct ViewCam {
    float R[9];
    float ifx, ify, cx, cy;   // 1/fx, 1/fy, px - .5, py - .5
};

__device__ __forceinline__ ViewCam load_view_cam(const float* __restrict__ cam, int b) {
    const float* c = cam + 16 * (int64_t)b;
    ViewCam v;
#pragma unroll
    for (int i = 0; i < 9; ++i) v.R[i] = c[i];
    v.ifx = __fdiv_rn(1.f, c[9]);
    v.ify = __fdiv_rn(1.f, c[10]);
    v.cx = __fsub_rn(c[11], 0.5f);
    v.cy = __fsub_rn(c[12], 0.5f);
    return v;
}

__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// camera-frame unit direction (dc) of pixel (x, y)
__device__ __forceinline__ void gen_ray_cam(const ViewCam& v, int x, int y, float& a, float& b, float& inv) {
    a = __fmul_rn(__fsub_rn(v.cx, (float)x), v.ifx);
    b = __fmul_rn(__fsub_rn(v.cy, (float)y), v.ify);
    inv = rsqrt_approx(__fmaf_rn(a, a, __fmaf_rn(b, b, 1.f)));
}

__device__ __forceinline__ void gen_ray(const ViewCam& v, int x, int y, float& d0, float& d1, float& d2) {
    float a, b, inv;
    gen_ray_cam(v, x, y, a, b, inv);
    const float c0 = __fmul_rn(a, inv), c1 = __fmul_rn(b, inv);
    d0 = __fmaf_rn(v.R[2], inv, __fmaf_rn(v.R[1], c1, __fmul_rn(v.R[0], c0)));
    d1 = __fmaf_rn(v.R[5], inv, __fmaf_rn(v.R[4], c1, __fmul_rn(v.R[3], c0)));
    d2 = __fmaf_rn(v.R[8], inv, __fmaf_rn(v.R[7], c1, __fmul_rn(v.R[6], c0)));
}

// ray of pixel (x, y) of view b: read from the (B,H,W,3) tensor when one is given (user-supplied rays, pytorch3d
// cameras), generated otherwise
__device__ __forceinline__ void pixel_ray(const float* __restrict__ rays, const float* __restrict__ cam, int b, int x,
                                          int y, int H, int W, float& d0, float& d1, float& d2) {
    if (rays != nullptr) {
        const int64_t r = ((int64_t)b * H + y) * W + x;
        d0 = rays[r * 3 + 0]; d1 = rays[r * 3 + 1]; d2 = rays[r * 3 + 2];
    } else {
        const ViewCam v = load_view_cam(cam, b);
        gen_ray(v, x, y, d0, d1, d2);
    }
}

// Chain rule of the generator: d/d(ray direction) of one pixel -> its contribution to d/d(cam record)
// [dR (9), dfx, dfy, dpx, dpy] (the fused backward sums these per view: the reference's grad_rays,
// RayTracing.py:179-206, carried on to R / focal by autograd through pytorch3d's ray sampler).
__device__ __forceinline__ void gen_ray_backward(const ViewCam& v, int x, int y, float g0, float g1, float g2,
                                                 float* out13) {
    float a, b, inv;
    gen_ray_cam(v, x, y, a, b, inv);
    const float c0 = a * inv, c1 = b * inv, c2 = inv;
    out13[0] = g0 * c0; out13[1] = g0 * c1; out13[2] = g0 * c2;
    out13[3] = g1 * c0; out13[4] = g1 * c1; out13[5] = g1 * c2;
    out13[6] = g2 * c0; out13[7] = g2 * c1; out13[8] = g2 * c2;
    // g_dc = R^T g ; through the normalisation dc = v / |v|, v = (a, b, 1): g_v = (g_dc - dc (dc . g_dc)) / |v|
    const float h0 = v.R[0] * g0 + v.R[3] * g1 + v.R[6] * g2;
    const float h1 = v.R[1] * g0 + v.R[4] * g1 + v.R[7] * g2;
    const float h2 = v.R[2] * g0 + v.R[5] * g1 + v.R[8] * g2;
    const float dot = c0 * h0 + c1 * h1 + c2 * h2;
    const float ga = (h0 - c0 * dot) * inv, gb = (h1 - c1 * dot) * inv;
    // a = (px - .5 - x) / fx:  da/dpx = 1/fx, da/dfx = -a/fx
    out13[9] = -ga * a * v.ifx;
    out13[10] = -gb * b * v.ify;
    out13[11] = ga * v.ifx;
    out13[12] = gb * v.ify;
}

// Every tile owns kBinSub counters / list segments (entry g goes to segment g % kBinSub): L2 serialises atomics
// on one address, and a C5 tile receives several hundred entries per view.  The segments of a tile are
// adjacent, so a tile's list is tile_offsets[tile * kBinSub] .. tile_offsets[(tile + 1) * kBinSub].
constexpr int kBinSub = 8;

// threads per tile CTA (= pixel columns per tile in the per-tile hit tables)
static inline int tile_threads(int tile) {
    const int px = tile * tile;
    return px > 128 ? 256 : (px > 64 ? 128 : 64);
}

// monotone float -> uint (the high word of pack_key)
__device__ __forceinline__ unsigned orderable(float len) {
    const unsigned b = __float_as_uint(len + 0.f);                       // -0 -> +0
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}

}  // namespace voge

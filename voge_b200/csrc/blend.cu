// blend.cu -- occlusion-aware blending weights (aggregation) and the gather-blend
// (merge_final + background composite), forward and analytic backward.
//
// Reference: pure PyTorch, VoGE/Aggregation.py:30-107 (get_cross_activation, assign2weight,
// aggregation) materialising (R,K,K) tensors in ~15 elementwise kernels, :111-141 (merge_final)
// and VoGE/Renderer.py:153-176 (interpolate_attr / get_silhouette / to_colored_background).
// Here the K x K term lives in shared memory / registers of the thread that owns the ray.
#include "../../include/voge_b200.h"
#include "common.cuh"
#include "blend_core.cuh"
#include "sort_net.h"

namespace voge {

// ---- forward ---------------------------------------------------------------------------------
// one thread per ray; per-thread arrays [k][thread] in shared memory: len, s=sqrt(dsd+1e-10), E=exp(-act)
template <int NT>
__global__ void __launch_bounds__(NT) aggregation_fwd_kernel(const int32_t* __restrict__ idx,
                                                             const float* __restrict__ act,
                                                             const float* __restrict__ len,
                                                             const float* __restrict__ dsd, float omega,
                                                             int64_t R, int K, float* __restrict__ weight,
                                                             int64_t* __restrict__ valid_num) {
    extern __shared__ __align__(16) float sm[];
    float* s_len = sm;
    float* s_s = sm + (size_t)K * NT;
    float* s_E = sm + (size_t)2 * K * NT;
    const int tid = threadIdx.x;
    const int64_t r = (int64_t)blockIdx.x * NT + tid;
    if (r >= R) return;
    int nvalid = 0;
    for (int k = 0; k < K; ++k) {
        const int64_t o = r * K + k;
        s_len[k * NT + tid] = len[o];
        s_s[k * NT + tid] = sqrtf(dsd[o] + 1e-10f);
        s_E[k * NT + tid] = expf(-act[o]);
        if (idx != nullptr) nvalid += (idx[o] >= 0);
    }
    if (valid_num != nullptr) valid_num[r] = nvalid;
    for (int m = 0; m < K; ++m) {
        const float Em = s_E[m * NT + tid];
        float w = 0.f;
        if (Em != 0.f) {
            const float lm = s_len[m * NT + tid];
            float D = 0.f;
            for (int k = 0; k < K; ++k) {
                const float Ek = s_E[k * NT + tid];
                if (Ek == 0.f) continue;
                const float c = (lm - s_len[k * NT + tid]) * s_s[k * NT + tid];
                D += Ek * phi(c);
            }
            w = expf(-(D * omega)) * Em * kInvExpMinusHalf;
        }
        weight[r * K + m] = w;
    }
}

// ---- backward --------------------------------------------------------------------------------
// w_m = e^{.5} exp(-omega D_m) E_m,  D_m = sum_k E_k Phi(c_mk),  c_mk = (len_m - len_k) s_k
template <int NT>
__global__ void __launch_bounds__(NT) aggregation_bwd_kernel(const float* __restrict__ act,
                                                             const float* __restrict__ len,
                                                             const float* __restrict__ dsd,
                                                             const float* __restrict__ g_w, float omega,
                                                             int64_t R, int K, float* __restrict__ g_act,
                                                             float* __restrict__ g_len,
                                                             float* __restrict__ g_dsd) {
    extern __shared__ __align__(16) float sm[];
    const size_t A = (size_t)K * NT;
    float* s_len = sm;
    float* s_s = sm + A;
    float* s_E = sm + 2 * A;
    float* s_gl = sm + 3 * A;
    float* s_gd = sm + 4 * A;
    float* s_gE = sm + 5 * A;
    const int tid = threadIdx.x;
    const int64_t r = (int64_t)blockIdx.x * NT + tid;
    if (r >= R) return;
    for (int k = 0; k < K; ++k) {
        const int64_t o = r * K + k;
        s_len[k * NT + tid] = len[o];
        s_s[k * NT + tid] = sqrtf(dsd[o] + 1e-10f);
        s_E[k * NT + tid] = expf(-act[o]);
        s_gl[k * NT + tid] = 0.f;
        s_gd[k * NT + tid] = 0.f;
        s_gE[k * NT + tid] = 0.f;
    }
    for (int m = 0; m < K; ++m) {
        const float Em = s_E[m * NT + tid];
        if (Em == 0.f) continue;
        const float lm = s_len[m * NT + tid];
        float D = 0.f;
        for (int k = 0; k < K; ++k) {
            const float Ek = s_E[k * NT + tid];
            if (Ek == 0.f) continue;
            D += Ek * phi((lm - s_len[k * NT + tid]) * s_s[k * NT + tid]);
        }
        const float w = expf(-(D * omega)) * Em * kInvExpMinusHalf;
        const float gw = g_w[r * K + m];
        const float gD = -omega * w * gw;
        // direct path d w_m / d act_m = -w_m  (through the trailing exp(-act_m))
        s_gE[m * NT + tid] += w * gw / Em;   // expressed as a gradient on E_m: dw/dE_m = w/E_m
        if (gD == 0.f) continue;
        float glm = 0.f;
        for (int k = 0; k < K; ++k) {
            const float Ek = s_E[k * NT + tid];
            if (Ek == 0.f) continue;
            const float sk = s_s[k * NT + tid];
            const float dl = lm - s_len[k * NT + tid];
            const float c = dl * sk;
            s_gE[k * NT + tid] += gD * phi(c);
            if (fabsf(c) < 10.f) {  // exp(-c^2) underflows to 0 beyond |c| ~ 9.3
                const float gc = gD * Ek * expf(-c * c) * kInvSqrtPi;
                glm += gc * sk;
                s_gl[k * NT + tid] -= gc * sk;
                s_gd[k * NT + tid] += gc * dl / (2.f * sk);
            }
        }
        s_gl[m * NT + tid] += glm;
    }
    for (int k = 0; k < K; ++k) {
        const int64_t o = r * K + k;
        g_act[o] = -s_E[k * NT + tid] * s_gE[k * NT + tid];
        g_len[o] = s_gl[k * NT + tid];
        g_dsd[o] = s_gd[k * NT + tid];
    }
}

// ---- gather-blend ----------------------------------------------------------------------------
// one thread per (ray, channel); channel fastest => coalesced attribute-row reads and output writes
__global__ void __launch_bounds__(256) merge_fwd_kernel(const float* __restrict__ attr,
                                                        const float* __restrict__ weight,
                                                        const int32_t* __restrict__ idx,
                                                        const int64_t* __restrict__ valid_num,
                                                        const float* __restrict__ background, float mask_thr,
                                                        int64_t R, int K, int C, int idx_mod, int n_attr,
                                                        float* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= R * C) return;
    const int64_t r = t / C;
    const int c = (int)(t - r * C);
    const int nv = valid_num != nullptr ? (int)min((int64_t)K, valid_num[r]) : K;
    float acc = 0.f, wsum = 0.f;
    for (int k = 0; k < K; ++k) {
        const float w = weight[r * K + k];
        wsum += w;
        if (k < nv) {
            int g = idx[r * K + k];
            if (g < 0) g = 0;  // Aggregation.py:131  vert_assign += (vert_assign < 0)
            if (idx_mod > 0) g %= idx_mod;
            if (g < n_attr) acc = fmaf(w, __ldg(attr + (int64_t)g * C + c), acc);
        }
    }
    if (background != nullptr) {
        const float sil = fminf(wsum, 1.f);                                  // Renderer.py:157-159
        const float mask = mask_thr > 0.f ? (sil > mask_thr ? 1.f : 0.f) : sil;  // :167-168
        acc = fminf(acc + (1.f - mask) * background[c], 1.f);                // :171
    }
    out[t] = acc;
}

// torch.minimum-style subgradient of min(x, 1): 1 below, 1/2 at the tie, 0 above
__device__ __forceinline__ float min1_grad(float x) { return x < 1.f ? 1.f : (x == 1.f ? 0.5f : 0.f); }

constexpr int kMaxBgChannels = 32;  // background composite backward supports C <= 32

// one thread per ray
__global__ void __launch_bounds__(256) merge_bwd_kernel(const float* __restrict__ attr,
                                                        const float* __restrict__ weight,
                                                        const int32_t* __restrict__ idx,
                                                        const int64_t* __restrict__ valid_num,
                                                        const float* __restrict__ background, float mask_thr,
                                                        const float* __restrict__ g_out, int64_t R, int K,
                                                        int C, int idx_mod, int n_attr, int packed4, float* __restrict__ g_attr,
                                                        float* __restrict__ g_weight) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int nv = valid_num != nullptr ? (int)min((int64_t)K, valid_num[r]) : K;
    const bool bg = background != nullptr;
    float gol[kMaxBgChannels];
    float g_sumw = 0.f;
    if (bg) {
        // upstream gradient after the two clamps of Renderer.py:157-171 (needs the un-clamped colour)
        float wsum = 0.f;
        for (int k = 0; k < K; ++k) wsum += weight[r * K + k];
        const float sil = fminf(wsum, 1.f);
        const float mask = mask_thr > 0.f ? (sil > mask_thr ? 1.f : 0.f) : sil;
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
            for (int k = 0; k < nv; ++k) {
                int g = idx[r * K + k];
                if (g < 0) g = 0;
                if (idx_mod > 0) g %= idx_mod;
                if (g < n_attr) acc = fmaf(weight[r * K + k], __ldg(attr + (int64_t)g * C + c), acc);
            }
            const float go = g_out[r * C + c] * min1_grad(acc + (1.f - mask) * background[c]);
            gol[c] = go;
            if (!(mask_thr > 0.f)) g_sumw -= go * background[c];
        }
        g_sumw *= min1_grad(wsum);
    }
    for (int k = 0; k < K; ++k) {
        float gw = g_sumw;
        if (k < nv) {
            int g = idx[r * K + k];
            if (g < 0) g = 0;
            if (idx_mod > 0) g %= idx_mod;
            const float w = weight[r * K + k];
            float v4[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c = 0; c < C && g < n_attr; ++c) {
                const float go = bg ? gol[c] : g_out[r * C + c];
                gw = fmaf(go, __ldg(attr + (int64_t)g * C + c), gw);
                if (packed4) v4[c] = w * go;
                else if (g_attr != nullptr && w != 0.f && go != 0.f) atomicAdd(g_attr + (int64_t)g * C + c, w * go);
            }
            // C <= 4 with a (n_attr, 4) padded gradient table: ONE 16-byte vector reduction per hit
            if (packed4 && g_attr != nullptr && w != 0.f && g < n_attr)
                atomicAdd(reinterpret_cast<float4*>(g_attr + 4 * (int64_t)g), make_float4(v4[0], v4[1], v4[2], v4[3]));
        }
        if (g_weight != nullptr) g_weight[r * K + k] = gw;
    }
}

// Per-thread rows of the (R,K) tensors are read / written as 16-byte vectors when K % 4 == 0: a scalar
// access at a 4K-byte lane stride costs one L1 sector operation per lane and slot, which -- not DRAM --
// bounded the first version of these kernels (~1 TB/s, profiles/ncu_r1_final.md).
struct Row4 {
    float w[4];
    int g[4];
};
template <bool VEC>
__device__ __forceinline__ Row4 load_row4(const float* __restrict__ wrow, const int32_t* __restrict__ irow, int k, int K) {
    Row4 o;
    if (VEC) {
        const float4 w4 = *reinterpret_cast<const float4*>(wrow + k);
        const int4 i4 = *reinterpret_cast<const int4*>(irow + k);
        o.w[0] = w4.x; o.w[1] = w4.y; o.w[2] = w4.z; o.w[3] = w4.w;
        o.g[0] = i4.x; o.g[1] = i4.y; o.g[2] = i4.z; o.g[3] = i4.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            o.w[j] = (k + j < K) ? wrow[k + j] : 0.f;
            o.g[j] = (k + j < K) ? irow[k + j] : -1;
        }
    }
    return o;
}

// Row k..k+3 of a pixel whose first nv slots are valid: behind the valid hits only the weights are read (the
// silhouette sum needs them), the indices come back as -1.
template <bool VEC>
__device__ __forceinline__ Row4 load_row4_nv(const float* __restrict__ wrow, const int32_t* __restrict__ irow, int k, int K,
                                             int nv) {
    Row4 o;
    if (VEC) {
        const float4 w4 = *reinterpret_cast<const float4*>(wrow + k);
        int4 i4 = make_int4(-1, -1, -1, -1);
        if (k < nv) i4 = *reinterpret_cast<const int4*>(irow + k);
        o.w[0] = w4.x; o.w[1] = w4.y; o.w[2] = w4.z; o.w[3] = w4.w;
        o.g[0] = i4.x; o.g[1] = i4.y; o.g[2] = i4.z; o.g[3] = i4.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            o.w[j] = (k + j < K) ? wrow[k + j] : 0.f;
            o.g[j] = (k + j < nv) ? irow[k + j] : -1;
        }
    }
    return o;
}

// attribute row of Gaussian g: one 16-byte load from a float4-padded table (P4), C scalar loads otherwise
template <int C, bool P4>
__device__ __forceinline__ void load_attr(const float* __restrict__ attr, int g, float* av) {
    if (P4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(attr) + g);
        const float t[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int c = 0; c < C; ++c) av[c] = t[c];
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) av[c] = __ldg(attr + (int64_t)g * C + c);
    }
}

template <int C, bool VEC, bool P4>
__global__ void __launch_bounds__(256) merge_fwd_small_kernel(const float* __restrict__ attr,
                                                              const float* __restrict__ weight,
                                                              const int32_t* __restrict__ idx,
                                                              const int64_t* __restrict__ valid_num,
                                                              const float* __restrict__ background, float mask_thr,
                                                              int64_t R, int K, FastMod idx_mod, int n_attr,
                                                              float* __restrict__ out, uint8_t* __restrict__ sat_code,
                                                              int32_t* __restrict__ idx_pad) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int nv = valid_num != nullptr ? (int)min((int64_t)K, valid_num[r]) : K;
    const float* wrow = weight + r * K;
    const int32_t* irow = idx + r * K;
    if (idx_pad != nullptr) {
        // the reference rewrites vert_assign -1 -> 0 in place (Aggregation.py:131).  For fragments whose slots behind
        // valid_num are known to be the -1 padding (the caller's promise) that is a store of zeros -- no read, no
        // separate pass over the (R,K) tensor
        int32_t* prow = idx_pad + r * K;
        int k = nv;
        for (; k < K && (!VEC || (k & 3)); ++k) prow[k] = 0;
        for (; k + 4 <= K; k += 4) *reinterpret_cast<int4*>(prow + k) = make_int4(0, 0, 0, 0);
        for (; k < K; ++k) prow[k] = 0;
    }
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    float wsum = 0.f;
    // the rows of the next four slots are requested before the gathers of the current four are consumed: a pixel
    // costs one row latency plus K/4 gather latencies instead of K/4 of each
    Row4 row = load_row4_nv<VEC>(wrow, irow, 0, K, nv);
    for (int k = 0; k < K; k += 4) {
        Row4 nxt = row;
        if (k + 4 < K) nxt = load_row4_nv<VEC>(wrow, irow, k + 4, K, nv);
        // the four gathers are requested together (predicated loads, no branch between them)
        float av[4][C];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            wsum += row.w[j];
            const int g = fold_index(row.g[j], idx_mod);
            const bool ok = k + j < nv && g < n_attr;
#pragma unroll
            for (int c = 0; c < C; ++c) av[j][c] = 0.f;
            if (ok) load_attr<C, P4>(attr, g, av[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fmaf(row.w[j], av[j][c], acc[c]);
        row = nxt;
    }
    if (background != nullptr) {
        const float sil = fminf(wsum, 1.f);                                      // Renderer.py:157-159
        const float mask = mask_thr > 0.f ? (sil > mask_thr ? 1.f : 0.f) : sil;  // :167-168
        // where min(x, 1) clamps, per channel (2 bits: 2 = x < 1, 1 = x == 1, 0 = x > 1): the backward's factor
        // min1_grad(x) = code / 2 without rebuilding x
        unsigned code = 0u;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float x = acc[c] + (1.f - mask) * background[c];
            code |= (x < 1.f ? 2u : (x == 1.f ? 1u : 0u)) << (2 * c);
            acc[c] = fminf(x, 1.f);   // :171
        }
        if (sat_code != nullptr) sat_code[r] = (uint8_t)code;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) out[r * C + c] = acc[c];
}

template <int C, bool VEC, bool P4>
__global__ void __launch_bounds__(256) merge_bwd_small_kernel(const float* __restrict__ attr,
                                                              const float* __restrict__ weight,
                                                              const int32_t* __restrict__ idx,
                                                              const int64_t* __restrict__ valid_num,
                                                              const float* __restrict__ background, float mask_thr,
                                                              const float* __restrict__ g_out,
                                                              const float* __restrict__ fwd_out, int64_t R, int K,
                                                              FastMod idx_mod, int n_attr, float* __restrict__ g_attr4,
                                                              float* __restrict__ g_weight) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int nv = valid_num != nullptr ? (int)min((int64_t)K, valid_num[r]) : K;
    const float* wrow = weight + r * K;
    const int32_t* irow = idx + r * K;
    float go[C];
#pragma unroll
    for (int c = 0; c < C; ++c) go[c] = g_out[r * C + c];
    float g_sumw = 0.f;
    if (background != nullptr) {
        // min(x, 1) backward needs the unclamped composite x only where the forward output saturated:
        // out < 1 means x = out and the factor is 1.  With the forward's output at hand the gather-blend is
        // re-run for saturated pixels only (e.g. empty pixels on a white background: the tie x == 1).
        bool need_acc = fwd_out == nullptr;
        if (!need_acc) {
#pragma unroll
            for (int c = 0; c < C; ++c) need_acc = need_acc || !(fwd_out[r * C + c] < 1.f);
        }
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.f;
        float wsum = 0.f;
        if (need_acc) {
            // index rows only as far as valid hits go (an empty pixel on a white background saturates too)
            for (int k = 0; k < K; k += 4) {
                if (k >= nv) {
                    if (VEC) {
                        const float4 w4 = *reinterpret_cast<const float4*>(wrow + k);
                        wsum += w4.x; wsum += w4.y; wsum += w4.z; wsum += w4.w;
                    } else {
                        for (int j = 0; j < 4 && k + j < K; ++j) wsum += wrow[k + j];
                    }
                    continue;
                }
                const Row4 row = load_row4<VEC>(wrow, irow, k, K);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    wsum += row.w[j];
                    const int g = fold_index(row.g[j], idx_mod);
                    if (k + j < nv && g < n_attr) {
                        float av[C];
                        load_attr<C, P4>(attr, g, av);
#pragma unroll
                        for (int c = 0; c < C; ++c) acc[c] = fmaf(row.w[j], av[c], acc[c]);
                    }
                }
            }
        } else {
            for (int k = 0; k < K; k += 4) {
                if (VEC) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wrow + k);
                    wsum += w4.x; wsum += w4.y; wsum += w4.z; wsum += w4.w;
                } else {
                    for (int j = 0; j < 4 && k + j < K; ++j) wsum += wrow[k + j];
                }
            }
        }
        const float sil = fminf(wsum, 1.f);
        const float mask = mask_thr > 0.f ? (sil > mask_thr ? 1.f : 0.f) : sil;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (need_acc) go[c] *= min1_grad(acc[c] + (1.f - mask) * background[c]);
            if (!(mask_thr > 0.f)) g_sumw -= go[c] * background[c];
        }
        g_sumw *= min1_grad(wsum);
    }
    // software-pipelined rows as in the forward kernel
    Row4 row;
    if (nv > 0) row = load_row4<VEC>(wrow, irow, 0, K);
    for (int k = 0; k < K; k += 4) {
        if (k >= nv) {      // behind the valid hits only the silhouette term reaches the weights
            if (g_weight != nullptr) {
                if (VEC) {
                    *reinterpret_cast<float4*>(g_weight + r * K + k) = make_float4(g_sumw, g_sumw, g_sumw, g_sumw);
                } else {
                    for (int j = 0; j < 4 && k + j < K; ++j) g_weight[r * K + k + j] = g_sumw;
                }
            }
            continue;
        }
        Row4 nxt = row;
        if (k + 4 < nv) nxt = load_row4<VEC>(wrow, irow, k + 4, K);
        float gw4[4];
        // the four gathers are requested together (predicated loads); the reductions need no gathered value and
        // follow once all loads are in flight
        float av[4][C];
        int gs[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int g = fold_index(row.g[j], idx_mod);
            const bool ok = k + j < nv && g < n_attr;
            gs[j] = ok ? g : -1;
#pragma unroll
            for (int c = 0; c < C; ++c) av[j][c] = 0.f;
            if (ok) load_attr<C, P4>(attr, g, av[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (g_attr4 != nullptr && gs[j] >= 0 && row.w[j] != 0.f) {
                float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < C; ++c) v[c] = row.w[j] * go[c];
                atomicAdd(reinterpret_cast<float4*>(g_attr4 + 4 * (int64_t)gs[j]), make_float4(v[0], v[1], v[2], v[3]));
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float gw = g_sumw;
#pragma unroll
            for (int c = 0; c < C; ++c) gw = fmaf(go[c], av[j][c], gw);    // av == 0 for invalid slots
            gw4[j] = gw;
        }
        if (g_weight != nullptr) {
            if (VEC) {
                *reinterpret_cast<float4*>(g_weight + r * K + k) = make_float4(gw4[0], gw4[1], gw4[2], gw4[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (k + j < K) g_weight[r * K + k + j] = gw4[j];
            }
        }
        row = nxt;
    }
}

template <int NT>
static int launch_agg_fwd(const int32_t* idx, const float* act, const float* len, const float* dsd,
                          float omega, int64_t R, int K, float* weight, int64_t* valid_num,
                          cudaStream_t s) {
    const size_t smem = (size_t)3 * K * NT * 4;
    if (smem > 227 * 1024) return (int)cudaErrorInvalidValue;
    VOGE_CUDA_TRY(cudaFuncSetAttribute(aggregation_fwd_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    aggregation_fwd_kernel<NT><<<(unsigned)((R + NT - 1) / NT), NT, smem, s>>>(idx, act, len, dsd, omega, R, K, weight, valid_num);
    VOGE_LAUNCH_CHECK();
    return 0;
}

template <int NT>
static int launch_agg_bwd(const float* act, const float* len, const float* dsd, const float* g_w,
                          float omega, int64_t R, int K, float* g_act, float* g_len, float* g_dsd,
                          cudaStream_t s) {
    const size_t smem = (size_t)6 * K * NT * 4;
    if (smem > 227 * 1024) return (int)cudaErrorInvalidValue;
    VOGE_CUDA_TRY(cudaFuncSetAttribute(aggregation_bwd_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    aggregation_bwd_kernel<NT><<<(unsigned)((R + NT - 1) / NT), NT, smem, s>>>(act, len, dsd, g_w, omega, R, K, g_act, g_len, g_dsd);
    VOGE_LAUNCH_CHECK();
    return 0;
}

}  // namespace voge

extern "C" int voge_aggregation(const int32_t* idx, const float* act, const float* len, const float* dsd,
                                float absorptivity, int64_t R, int K, float* weight, int64_t* valid_num,
                                voge_stream_t stream) {
    using namespace voge;
    if (R <= 0 || K <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (K <= 48) return launch_agg_fwd<128>(idx, act, len, dsd, absorptivity, R, K, weight, valid_num, s);
    if (K <= 200) return launch_agg_fwd<64>(idx, act, len, dsd, absorptivity, R, K, weight, valid_num, s);
    return launch_agg_fwd<32>(idx, act, len, dsd, absorptivity, R, K, weight, valid_num, s);
}

extern "C" int voge_aggregation_backward(const float* act, const float* len, const float* dsd,
                                         const float* grad_weight, float absorptivity, int64_t R, int K,
                                         float* grad_act, float* grad_len, float* grad_dsd,
                                         voge_stream_t stream) {
    using namespace voge;
    if (R <= 0 || K <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (K <= 32) return launch_agg_bwd<128>(act, len, dsd, grad_weight, absorptivity, R, K, grad_act, grad_len, grad_dsd, s);
    if (K <= 110) return launch_agg_bwd<64>(act, len, dsd, grad_weight, absorptivity, R, K, grad_act, grad_len, grad_dsd, s);
    return launch_agg_bwd<32>(act, len, dsd, grad_weight, absorptivity, R, K, grad_act, grad_len, grad_dsd, s);
}

extern "C" int voge_merge_final(const float* attr, const float* weight, const int32_t* idx,
                                const int64_t* valid_num, const float* background, float mask_thr,
                                int64_t R, int K, int C, int idx_mod, int n_attr, int attr_padded4, float* out,
                                uint8_t* sat_code, int32_t* idx_pad, voge_stream_t stream) {
    using namespace voge;
    if (R <= 0 || C <= 0) return 0;
    if (C <= 4) {
        const unsigned grid = (unsigned)((R + 255) / 256);
        cudaStream_t s = (cudaStream_t)stream;
        const FastMod fm = make_fastmod(idx_mod);
#define VOGE_MF(CC)                                                                                                 \
    do {                                                                                                            \
        if (K % 4 == 0 && attr_padded4)                                                                             \
            merge_fwd_small_kernel<CC, true, true><<<grid, 256, 0, s>>>(attr, weight, idx, valid_num, background,  \
                                                                        mask_thr, R, K, fm, n_attr, out, sat_code, idx_pad); \
        else if (K % 4 == 0)                                                                                        \
            merge_fwd_small_kernel<CC, true, false><<<grid, 256, 0, s>>>(attr, weight, idx, valid_num, background, \
                                                                         mask_thr, R, K, fm, n_attr, out, sat_code, idx_pad);\
        else if (attr_padded4)                                                                                      \
            merge_fwd_small_kernel<CC, false, true><<<grid, 256, 0, s>>>(attr, weight, idx, valid_num, background, \
                                                                         mask_thr, R, K, fm, n_attr, out, sat_code, idx_pad);\
        else                                                                                                        \
            merge_fwd_small_kernel<CC, false, false><<<grid, 256, 0, s>>>(attr, weight, idx, valid_num, background,\
                                                                          mask_thr, R, K, fm, n_attr, out, sat_code, idx_pad);\
    } while (0)
        if (C == 1) VOGE_MF(1); else if (C == 2) VOGE_MF(2); else if (C == 3) VOGE_MF(3); else VOGE_MF(4);
#undef VOGE_MF
        VOGE_LAUNCH_CHECK();
        return 0;
    }
    if (attr_padded4 || idx_pad != nullptr) return (int)cudaErrorInvalidValue;     // C <= 4 only
    const int64_t total = R * C;
    merge_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        attr, weight, idx, valid_num, background, mask_thr, R, K, C, idx_mod, n_attr, out);
    VOGE_LAUNCH_CHECK();
    return 0;
}

extern "C" int voge_merge_final_backward(const float* attr, const float* weight, const int32_t* idx,
                                         const int64_t* valid_num, const float* background,
                                         float mask_thr, const float* out, const float* grad_out,
                                         int64_t R, int K, int C, int idx_mod, int n_attr, int packed4,
                                         int attr_padded4, float* grad_attr, float* grad_weight,
                                         voge_stream_t stream) {
    using namespace voge;
    if (R <= 0 || C <= 0) return 0;
    if (attr_padded4 && !(C <= 4 && (packed4 || grad_attr == nullptr))) return (int)cudaErrorInvalidValue;
    if (background != nullptr && C > kMaxBgChannels) return (int)cudaErrorInvalidValue;
    if (C <= 4 && (packed4 || grad_attr == nullptr)) {
        const unsigned grid = (unsigned)((R + 255) / 256);
        cudaStream_t s = (cudaStream_t)stream;
        const FastMod fm = make_fastmod(idx_mod);
#define VOGE_MB(CC)                                                                                                 \
    do {                                                                                                            \
        if (K % 4 == 0 && attr_padded4)                                                                             \
            merge_bwd_small_kernel<CC, true, true><<<grid, 256, 0, s>>>(attr, weight, idx, valid_num, background,  \
                                                                        mask_thr, grad_out, out, R, K, fm, n_attr,      \
                                                                        grad_attr, grad_weight);                    \
        else if (K % 4 == 0)                                                                                        \
            merge_bwd_small_kernel<CC, true, false><<<grid, 256, 0, s>>>(attr, weight, idx, valid_num, background, \
                                                                         mask_thr, grad_out, out, R, K, fm, n_attr,     \
                                                                         grad_attr, grad_weight);                   \
        else if (attr_padded4)                                                                                      \
            merge_bwd_small_kernel<CC, false, true><<<grid, 256, 0, s>>>(attr, weight, idx, valid_num, background, \
                                                                         mask_thr, grad_out, out, R, K, fm, n_attr,     \
                                                                         grad_attr, grad_weight);                   \
        else                                                                                                        \
            merge_bwd_small_kernel<CC, false, false><<<grid, 256, 0, s>>>(attr, weight, idx, valid_num, background,\
                                                                          mask_thr, grad_out, out, R, K, fm, n_attr,     \
                                                                          grad_attr, grad_weight);                  \
    } while (0)
        if (C == 1) VOGE_MB(1); else if (C == 2) VOGE_MB(2); else if (C == 3) VOGE_MB(3); else VOGE_MB(4);
#undef VOGE_MB
        VOGE_LAUNCH_CHECK();
        return 0;
    }
    merge_bwd_kernel<<<(unsigned)((R + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        attr, weight, idx, valid_num, background, mask_thr, grad_out, R, K, C, idx_mod, n_attr, packed4 && C <= 4, grad_attr,
        grad_weight);
    VOGE_LAUNCH_CHECK();
    return 0;
}

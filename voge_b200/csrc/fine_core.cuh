// fine_core.cuh -- the per-pixel ray-tracing core shared by the API-compatible fine kernel
// (bin_points candidate lists) and the fused renderer kernel (per-tile CSR lists).
//
// Algorithm per CTA (= a set of NT pixels that share one candidate list):
//   1. candidates are staged through shared memory in chunks of CH: one thread per candidate
//      loads (mu, S) and derives 10 floats of "filter data" (q = S^T mu, symmetric part of S,
//      u' = msm - thr - margin);
//   2. every pixel thread runs the FILTER over the chunk: ksk~ = d^T S d (6 FMA on per-pixel
//      monomials), msk~ = q.d (3 FMA), f = u'*ksk~ - msk~^2; "f >= 0" proves
//      act >= thr + margin and rejects the pair (this is the hot loop: broadcast LDS.128 + FFMA);
//   3. survivors are queued per thread and REFINED with the bit-faithful reference arithmetic
//      (exact_pair, common.cuh) from the original (mu, S); only the exact act < thr_act and the
//      exact len decide, so results are bit-identical to the reference kernel
//      (ray_trace_voge.cu:188-213) no matter how loose the filter is;
//   4. top-K smallest (len, idx) kept in per-thread sorted lists in shared memory
//      ([k][thread] layout, conflict-free).  Lexicographic (len, idx) order equals the
//      reference's "strict <, earlier candidate wins" rule (:197,:203) for ascending candidate
//      lists and makes the result independent of candidate order.
#pragma once
#include "common.cuh"

namespace voge {

constexpr int kStageFloats = 12;  // floats of filter data per staged candidate (3 x float4)
constexpr int kQueueCap = 16;     // per-thread survivor queue depth

// Filter data for one candidate, computed by the staging thread.
//   v0 = (q0, q1, q2, u')   v1 = (S00, S11, S22, S01+S10)   v2 = (S02+S20, S12+S21, idx, 0)
__device__ __forceinline__ void stage_candidate(float* __restrict__ dst, int g, const float* mu,
                                                const float* S, float thr_act) {
    float4 v0, v1, v2;
    const float m0 = mu[0], m1 = mu[1], m2 = mu[2];
    const float q0 = m0 * S[0] + m1 * S[3] + m2 * S[6];
    const float q1 = m0 * S[1] + m1 * S[4] + m2 * S[7];
    const float q2 = m0 * S[2] + m1 * S[5] + m2 * S[8];
    const float msm = q0 * m0 + q1 * m1 + q2 * m2;
    const float a = S[0], b = S[4], c = S[8];
    const float e01 = S[1] + S[3], e02 = S[2] + S[6], e12 = S[5] + S[7];
    const float s01 = 0.5f * e01, s02 = 0.5f * e02, s12 = 0.5f * e12;
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
    // Worst-case |act_reference - act_filter| for this Gaussian over all unit rays
    // (forward error analysis of both evaluation orders, DESIGN.md "filter margin"):
    //   eps * [16 Tmm + (Qn/l)(26 Um2 + 8 Qn) + (17 Us/l + 3) Qn^2/l + 4 msm],  eps = 2^-24, x1.25 slack
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
    const float bound = 16.f * Tmm + (Qn * il) * (26.f * Um2 + 8.f * Qn) + (17.f * Us * il + 3.f) * (Qn * Qn * il) +
                        4.f * fabsf(msm) + 8.f * fabsf(thr_act);
    const float margin = 7.4505806e-8f * bound;   // 1.25 * 2^-24
    const float u = msm - thr_act - margin;
    const bool pd = (lmin > 0.f) && (fabsf(u) < 3.0e38f) && (fabsf(q0) + fabsf(q1) + fabsf(q2) < 3.0e38f);
    if (pd) {
        v0 = make_float4(q0, q1, q2, u);
        v1 = make_float4(a, b, c, e01);
        v2 = make_float4(e02, e12, __int_as_float(g), margin);
    } else {
        // not provably safe to filter: ksk~ = |d|^2 > 0 and u' = -inf => never rejected
        v0 = make_float4(0.f, 0.f, 0.f, -INFINITY);
        v1 = make_float4(1.f, 1.f, 1.f, 0.f);
        v2 = make_float4(0.f, 0.f, __int_as_float(g), 0.f);
    }
    float4* d4 = reinterpret_cast<float4*>(dst);
    d4[0] = v0; d4[1] = v1; d4[2] = v2;
}

__device__ __forceinline__ void stage_invalid(float* __restrict__ dst) {
    float4* d4 = reinterpret_cast<float4*>(dst);
    d4[0] = make_float4(0.f, 0.f, 0.f, INFINITY);   // u' = +inf => always rejected
    d4[1] = make_float4(1.f, 1.f, 1.f, 0.f);
    d4[2] = make_float4(0.f, 0.f, __int_as_float(-1), 0.f);
}

// Per-thread sorted top-K list living in shared memory with [k][thread] layout.
template <int NT>
struct TopK {
    float* s_len;  // [K][NT]
    int* s_idx;    // [K][NT]
    int K, tid, cnt;
    float kth_len;
    int kth_idx;

    __device__ __forceinline__ void init(float* lens, int* idxs, int K_, int tid_) {
        s_len = lens; s_idx = idxs; K = K_; tid = tid_; cnt = 0;
        kth_len = kEmptyLen; kth_idx = 0x7fffffff;
    }
    // quick reject against the current K-th entry (exact semantics)
    __device__ __forceinline__ bool admits(float len, int g) const {
        return (len < kth_len) || (cnt == K && len == kth_len && g < kth_idx);
    }
    __device__ __forceinline__ void insert(float len, int g) {
        if (!admits(len, g)) return;   // also rejects NaN and len >= 1e10 (reference :197)
        int pos;
        if (cnt == K) pos = K - 1; else pos = cnt++;
        while (pos > 0) {
            const float pl = s_len[(pos - 1) * NT + tid];
            const int pi = s_idx[(pos - 1) * NT + tid];
            if (pl < len || (pl == len && pi < g)) break;
            s_len[pos * NT + tid] = pl;
            s_idx[pos * NT + tid] = pi;
            --pos;
        }
        s_len[pos * NT + tid] = len;
        s_idx[pos * NT + tid] = g;
        if (cnt == K) {
            kth_len = s_len[(K - 1) * NT + tid];
            kth_idx = s_idx[(K - 1) * NT + tid];
        }
    }
};

// Unsorted top-K buffer of packed 64-bit keys (orderable(len) << 32 | idx) in shared memory,
// [k][thread] layout.  Appending is O(1) and divergence-free while the buffer is not full (the
// common case: a pixel rarely collects K hits); once full the current maximum is replaced and
// re-scanned.  One lock-step insertion sort at the end orders the survivors.  Key order ==
// lexicographic (len, idx) order == the reference's insertion rule (ray_trace_voge.cu:197-213).
__device__ __forceinline__ unsigned long long pack_key(float len, int g) {
    const unsigned b = __float_as_uint(len + 0.f);                       // -0 -> +0
    const unsigned o = b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);       // monotone float -> uint
    return ((unsigned long long)o << 32) | (unsigned)g;
}

template <int NT>
struct TopKU {
    unsigned long long* s_key;  // [K][NT]
    int K, tid, cnt, maxpos;
    unsigned long long limit;   // a key is admitted iff key < limit

    __device__ __forceinline__ void init(unsigned long long* keys, int K_, int tid_) {
        s_key = keys; K = K_; tid = tid_; cnt = 0; maxpos = 0;
        limit = pack_key(kEmptyLen, 0);   // len < 1e10 strictly (reference :197 with the 1e10 initial slots)
    }
    __device__ __forceinline__ void rescan() {
        unsigned long long m = s_key[tid];
        int mp = 0;
        for (int k = 1; k < K; ++k) {
            const unsigned long long v = s_key[k * NT + tid];
            if (v > m) { m = v; mp = k; }
        }
        limit = m; maxpos = mp;
    }
    __device__ __forceinline__ void insert(float len, int g) {
        if (len != len) return;                          // NaN never compares less (reference :197)
        const unsigned long long key = pack_key(len, g);
        if (!(key < limit)) return;
        if (cnt < K) {
            s_key[cnt * NT + tid] = key;
            if (++cnt == K) rescan();
        } else {
            s_key[maxpos * NT + tid] = key;
            rescan();
        }
    }
    // ascending insertion sort of the cnt keys (all lanes of a warp run it together)
    __device__ __forceinline__ void sort() {
        for (int i = 1; i < cnt; ++i) {
            const unsigned long long key = s_key[i * NT + tid];
            int j = i;
            while (j > 0) {
                const unsigned long long p = s_key[(j - 1) * NT + tid];
                if (p < key) break;
                s_key[j * NT + tid] = p;
                --j;
            }
            s_key[j * NT + tid] = key;
        }
    }
};

// Per-pixel state of the filter.
struct RayMono {
    float d0, d1, d2;
    float dxx, dyy, dzz, dxy, dxz, dyz;
    __device__ __forceinline__ void set(float x, float y, float z) {
        d0 = x; d1 = y; d2 = z;
        dxx = x * x; dyy = y * y; dzz = z * z; dxy = x * y; dxz = x * z; dyz = y * z;
        // the filter margin assumes unit rays (the renderer's always are); for anything else the
        // monomials are poisoned with NaN so that the filter never rejects (f = NaN => refine).
        if (!(fabsf(dxx + dyy + dzz - 1.f) <= 1e-3f)) dxx = __int_as_float(0x7fc00000);
    }
    // lanes without a pixel: ksk~ = tr(S) > 0, msk~ = 0 => rejected by every filterable candidate
    __device__ __forceinline__ void set_dead() {
        d0 = d1 = d2 = 0.f;
        dxx = dyy = dzz = 1.f; dxy = dxz = dyz = 0.f;
    }
};

// true  => pair may satisfy act < thr_act and must be refined; false => provably rejected.
__device__ __forceinline__ bool filter_pass(const float* __restrict__ st, const RayMono& r) {
    const float4 v0 = *reinterpret_cast<const float4*>(st);
    const float4 v1 = *reinterpret_cast<const float4*>(st + 4);
    const float2 v2 = *reinterpret_cast<const float2*>(st + 8);
    float ksk = v1.x * r.dxx;
    ksk = fmaf(v1.y, r.dyy, ksk);
    ksk = fmaf(v1.z, r.dzz, ksk);
    ksk = fmaf(v1.w, r.dxy, ksk);
    ksk = fmaf(v2.x, r.dxz, ksk);
    ksk = fmaf(v2.y, r.dyz, ksk);
    float msk = v0.x * r.d0;
    msk = fmaf(v0.y, r.d1, msk);
    msk = fmaf(v0.z, r.d2, msk);
    const float f = fmaf(v0.w, ksk, -msk * msk);
    return !(f >= 0.f);
}

}  // namespace voge

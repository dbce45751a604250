// render_bwd_agg.cuh -- fused backward with a per-tile, Gaussian-major gradient reduction (included by render.cu).
//
// Why.  The two-threads-per-pixel backward (render_bwd_pair_kernel) spends, per selected hit, two gathers of the
// Gaussian's record and three 16-byte vector reductions (four with the attribute gradient of the image mode), every one
// of them a request to its own L2 line.  Measured on B200 (tools/microbench/l1_costs.cu, profiles/l1_costs_r2.txt): a
// scattered LDG.128 costs 2.1 SM cycles per LANE, a scattered RED.ADD.F32x4 1.5 -- about 12 cycles per hit, which is
// what the kernel takes (0.50 ms per C5 view for 11.2 M hits on 148 SMs).  Shared-memory traffic is 10x cheaper
// (ATOMS.CAS 0.25, ATOMS.ADD 0.06-0.13, LDS 0.03 cycles per lane), and an 8x8 pixel tile of the C5 scene holds 9.7
// hits per distinct Gaussian.  So the per-hit chain rule is replaced by:
//   pass 0 / pass 2  as before, per (pixel, slot): recompute the hit, blend backward -> the three scalars
//                    (g_ksk, g_msk, g_msm) of ray_trace_voge.cu:324-326 -- kept in shared memory, no second gather;
//   grouping         the hit's Gaussian index is inserted into a 512-slot hash table of the tile (ATOMS.CAS), its rank
//                    inside the Gaussian's group comes from an ATOMS.ADD, a block scan turns the counts into offsets and
//                    the hits are scattered into group order (2-byte hit ids);
//   reduction        teams of four lanes walk one group each: the moments  A = sum g_msk d,  b = sum g_msm,
//                    C = sum g_ksk d d^T  (and sum w go for the attribute gradient) are accumulated in registers,
//                    combined by two shuffles, and ONE gather of the record + ONE set of vector reductions per
//                    (Gaussian, tile) applies  d mu = S A + (S + S^T) mu b,  d S = C + mu A^T + b mu mu^T
//                    (the sums of ray_trace_voge.cu:41-91 over the group, mu and S being constant inside it).
// A hit that finds no slot (more than ~500 distinct Gaussians in one tile) takes the per-hit path.  The result equals
// the per-hit kernel's up to the order of the fp32 sums.  north_star: "emits gradients ... with warp-aggregated atomics".
#pragma once

namespace voge {

constexpr int kAggSlots = 512;     // hash slots = distinct Gaussians aggregated per tile
constexpr int kAggProbes = 8;
constexpr int kAggMaxK = 24;       // 36 K + 10 KB of shared memory per 64-pixel CTA: 4 CTAs per SM up to K = 20
constexpr int kAggNT = 128, kAggNP = 64;

static inline size_t agg_smem_bytes(int K) {
    const size_t A = (size_t)K * kAggNP;
    return A * (16 + 8 + 4 + 4 + 2 + 2) + (size_t)kAggNP * (16 + 12) + (size_t)kAggSlots * (4 + 4 + 2 + 2) + 64;
}

// open-addressing insert of a Gaussian index; returns its slot or -1 when the probe sequence is exhausted
__device__ __forceinline__ int agg_insert(int* s_key, int g) {
    unsigned h = ((unsigned)g * 2654435761u) >> 23;          // 512 slots
#pragma unroll 1
    for (int p = 0; p < kAggProbes; ++p) {
        const int old = atomicCAS(&s_key[h], -1, g);
        if (old == -1 || old == g) return (int)h;
        h = (h + 1) & (kAggSlots - 1);
    }
    return -1;
}

// per-hit fallback: the chain rule of render_bwd_pair_kernel for one hit (exact ksk / msk from the record)
template <int KIND>
__device__ __noinline__ void agg_direct_hit(const float* __restrict__ gauss, GradSink a, int g, float c0, float c1, float c2,
                                            float d0, float d1, float d2, float gl, float ga, float gd) {
    float v0, v1, v2, S[9];
    load_gauss<KIND>(gauss, g, v0, v1, v2, S);
    const float m0 = __fsub_rn(v0, c0), m1 = __fsub_rn(v1, c1), m2 = __fsub_rn(v2, c2);
    const Prod9 pd = exact_row_products(d0, d1, d2, S);
    const Prod9 pm = exact_row_products(m0, m1, m2, S);
    const float ksk = exact_contract(pd, d0, d1, d2);
    const float msk = exact_contract(pm, d0, d1, d2);
    geom_grad_accumulate<false>(a, g, m0, m1, m2, S, d0, d1, d2, ksk, msk, gl, ga, gd, nullptr);
}

template <int KIND, bool IMG>
__global__ void __launch_bounds__(kAggNT, 4) render_bwd_agg_kernel(const FusedBwdArgs a) {
    constexpr int NT = kAggNT, NP = kAggNP, T = kAggSlots;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = a.K;
    const size_t A = (size_t)K * NP;
    float4* s_B = reinterpret_cast<float4*>(smem_raw);                    // [K][NP] (g_ksk, g_msk, g_msm, w)
    float2* s_ls = reinterpret_cast<float2*>(s_B + A);                    // [K][NP] (len, s = sqrt(dsd + 1e-10))
    float* s_E = reinterpret_cast<float*>(s_ls + A);                      // [K][NP] exp(-act)
    float* s_wg = s_E + A;                                                // [K][NP] w dL/dw
    float4* s_go = reinterpret_cast<float4*>(s_wg + A);                   // [NP] image gradient of the pixel (IMG)
    float* s_ray = reinterpret_cast<float*>(s_go + NP);                   // [3][NP]
    int* s_key = reinterpret_cast<int*>(s_ray + 3 * NP);                  // [T] Gaussian index of the slot, -1 = free
    int* s_cnt = s_key + T;                                               // [T] hits of the slot
    int* s_scan = s_cnt + T;                                              // [16] scan scratch, [8] = number of groups
    unsigned short* s_start = reinterpret_cast<unsigned short*>(s_scan + 16);   // [T] first position of the slot's group
    unsigned short* s_glist = s_start + T;                                // [T] non-empty slots, compacted
    unsigned short* s_info = s_glist + T;                                 // [K][NP] slot << 6 | rank in group, 0xffff = none
    unsigned short* s_order = s_info + A;                                 // [K*NP] hit ids (k * NP + col) in group order

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, sub = lane & 1, pp = lane >> 1;
    const int col = warp * 16 + pp;
    const unsigned pair_mask = 3u << (lane & 30);
    const int TX8 = (a.W + 7) >> 3, TY8 = (a.H + 7) >> 3;
    int blk = blockIdx.x;
    const int tx = blk % TX8; blk /= TX8;
    const int ty = blk % TY8;
    const int b = blk / TY8;
    const int xi = tx * 8 + (warp & 1) * 4 + (pp & 3), yi = ty * 8 + (warp >> 1) * 4 + (pp >> 2);
    const bool live = xi < a.W && yi < a.H;
    const int64_t r = live ? ((int64_t)b * a.H + yi) * a.W + xi : 0;
    const int cnt = live ? (int)min((int64_t)K, a.valid[r]) : 0;
    for (int i = tid; i < T; i += NT) { s_key[i] = -1; s_cnt[i] = 0; }
    float d0 = 0.f, d1 = 0.f, d2 = 1.f;
    if (live) pixel_ray(a.rays, a.cam, b, xi, yi, a.H, a.W, d0, d1, d2);
    if (sub == 0) { s_ray[col] = d0; s_ray[NP + col] = d1; s_ray[2 * NP + col] = d2; }
    const float c0 = a.origins[3 * b], c1 = a.origins[3 * b + 1], c2 = a.origins[3 * b + 2];
    const float omega = a.omega;
    const int32_t* i_idx = a.idx + r * K;
    const float* i_gw = IMG ? nullptr : a.g_weight + r * K;
    const float* i_w = a.weight + r * K;
    const int pack_off = b * a.N;
    const bool vec = (K & 3) == 0;

    // ---- image mode: per-pixel upstream gradient (see render_bwd_pair_kernel) ----
    float go[4] = {0.f, 0.f, 0.f, 0.f};
    float gB = 0.f;
    if (IMG) {
        if (cnt > 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < a.C) go[c] = a.g_out[r * a.C + c];
            if (a.background != nullptr) {
                float wsum = 0.f;
                for (int k = 0; k < cnt; ++k) wsum += i_w[k];
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
        if (sub == 0) s_go[col] = make_float4(go[0], go[1], go[2], go[3]);
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
            const float4 w4 = *reinterpret_cast<const float4*>(i_w + k0);
            wv[0] = sub ? w4.y : w4.x; wv[1] = sub ? w4.w : w4.z;
        } else {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int k = k0 + sub + 2 * jj;
                if (k < cnt) {
                    gv[jj] = i_idx[k];
                    if (!IMG) gwv[jj] = i_gw[k];
                    wv[jj] = i_w[k];
                }
            }
        }
        if (IMG) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int k = k0 + sub + 2 * jj;
                const int g = gv[jj] - pack_off;
                gwv[jj] = -gB;
                if (k < cnt && g >= 0 && g < a.N) {
                    const float4 av = __ldg(reinterpret_cast<const float4*>(a.attr4) + g);
                    gwv[jj] = fmaf(go[3], av.w, fmaf(go[2], av.z, fmaf(go[1], av.y, go[0] * av.x))) - gB;
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
                const float wg = wv[jj] * gwv[jj];
                s_wg[k * NP + col] = wg;
                total_gD -= omega * wg;           // gD_m = dL/dD_m = -omega w_m dL/dw_m
                s_B[k * NP + col] = make_float4(0.f, 0.f, 0.f, wv[jj]);
                s_info[k * NP + col] = 0xffffu;
            }
        }
    }
    s_min = fminf(s_min, __shfl_xor_sync(pair_mask, s_min, 1));
    total_gD += __shfl_xor_sync(pair_mask, total_gD, 1);
    __syncthreads();          // hash table initialised, the pixel's arrays complete

    // ---- pass 2: blend backward per slot -> (g_ksk, g_msk, g_msm); the hit joins its Gaussian's group ----
    {
        int lo_j = 0, hi_j = -1;
        float pref = 0.f;                 // sum of gD_m over m <= hi_j
        for (int j = sub; j < cnt; j += 2) {
            const float2 lsj = s_ls[j * NP + col];
            const float lj = lsj.x, sj = lsj.y;
            while (lo_j < j && (lj - s_ls[lo_j * NP + col].x) * s_min >= kErfSat) ++lo_j;
            while (hi_j + 1 < cnt && (s_ls[(hi_j + 1) * NP + col].x - lj) * s_min < kErfSat) {
                ++hi_j;
                pref -= omega * s_wg[hi_j * NP + col];
            }
            const float Ej = s_E[j * NP + col];
            if (Ej == 0.f) continue;              // weight 0: neither a geometry nor an attribute gradient
            const float wgj = s_wg[j * NP + col];
            const float gDj = -omega * wgj;
            float gE = wgj / Ej + (total_gD - pref) + 0.5f * gDj;
            float gl = 0.f, gd = 0.f;
            const float inv2sj = 0.5f / sj;
            const float gDjk = gDj * kInvSqrtPi, Ejk = Ej * kInvSqrtPi;
            for (int t = lo_j; t < hi_j; ++t) {
                const int i = t + (t >= j ? 1 : 0);
                const float2 lsi = s_ls[i * NP + col];
                const float dl = lsi.x - lj;
                const float gDi = -omega * s_wg[i * NP + col];
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
                const float c2_ = -dl * lsi.y;
                if (fabsf(c2_) < kErfSat) gl += gDjk * s_E[i * NP + col] * __expf(-c2_ * c2_) * lsi.y;
            }
            const float ga = -Ej * gE;
            if (a.g_len_out != nullptr) gl += a.g_len_out[r * K + j];
            const int g = i_idx[j] - pack_off;
            if (g < 0 || g >= a.N) continue;
            // msk = len ksk, ksk = s^2 - 1e-10:  g_ksk = (ga msk - gl) msk / ksk^2 + gd,  g_msk = (gl - 2 ga msk) / ksk
            const float ik = 1.f / fmaf(sj, sj, -1e-10f);
            const float g_ksk = fmaf(ga * lj, lj, gd) - gl * lj * ik;
            const float g_msk = gl * ik - 2.f * ga * lj;
            const int slot = agg_insert(s_key, g);
            if (slot >= 0) {
                const int rank = atomicAdd(&s_cnt[slot], 1);       // < 64: a Gaussian hits a pixel at most once
                const float wj = s_B[j * NP + col].w;
                s_B[j * NP + col] = make_float4(g_ksk, g_msk, ga, wj);
                s_info[j * NP + col] = (unsigned short)((slot << 6) | (rank & 63));
            } else {
                agg_direct_hit<KIND>(a.gauss, GradSink{a.grad_packed, a.kind, a.need_sigma}, g, c0, c1, c2, d0, d1, d2, gl, ga, gd);
                if (IMG && a.grad_attr4 != nullptr) {
                    const float wj = s_B[j * NP + col].w;
                    atomicAdd(reinterpret_cast<float4*>(a.grad_attr4) + g, make_float4(wj * go[0], wj * go[1], wj * go[2], wj * go[3]));
                }
            }
        }
    }
    __syncthreads();

    // ---- grouping: counts -> offsets (block scan over the slots, four per thread), compact list of groups ----
    {
        int c[4], tot = 0, ne = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            c[i] = s_cnt[4 * tid + i];
            tot += c[i];
            ne += c[i] > 0 ? 1 : 0;
        }
        int v = tot | (ne << 16), incl = v;           // hits < 2^16, groups <= 512: two scans in one word
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        int woff = 0, total = 0;
#pragma unroll
        for (int q = 0; q < NT / 32; ++q) {
            const int t = s_scan[q];
            if (q < warp) woff += t;
            total += t;
        }
        const int excl = woff + incl - v;
        int hs = excl & 0xffff, gs = excl >> 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            s_start[4 * tid + i] = (unsigned short)hs;
            if (c[i] > 0) s_glist[gs++] = (unsigned short)(4 * tid + i);
            hs += c[i];
        }
        if (tid == 0) s_scan[8] = total >> 16;
    }
    __syncthreads();
    for (int j = sub; j < cnt; j += 2) {
        const unsigned info = s_info[j * NP + col];
        if (info != 0xffffu) s_order[s_start[info >> 6] + (info & 63u)] = (unsigned short)(j * NP + col);
    }
    __syncthreads();

    // ---- reduction: one team of four lanes per group ----
    {
        const int team = tid >> 2, mem = tid & 3;
        const unsigned tmask = 0xfu << (lane & 28);
        const int G = s_scan[8];
        for (int gi = team; gi < G; gi += NT / 4) {
            const int slot = s_glist[gi];
            const int g = s_key[slot];
            const int beg = s_start[slot], n = s_cnt[slot];
            float A0 = 0.f, A1 = 0.f, A2 = 0.f, bb = 0.f;
            float C00 = 0.f, C01 = 0.f, C02 = 0.f, C11 = 0.f, C12 = 0.f, C22 = 0.f;
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
            for (int i = mem; i < n; i += 4) {
                const int hid = s_order[beg + i];
                const int cc = hid & (NP - 1);
                const float4 Bv = s_B[hid];
                const float e0 = s_ray[cc], e1 = s_ray[NP + cc], e2 = s_ray[2 * NP + cc];
                A0 = fmaf(Bv.y, e0, A0); A1 = fmaf(Bv.y, e1, A1); A2 = fmaf(Bv.y, e2, A2);
                bb += Bv.z;
                const float k0 = Bv.x * e0, k1 = Bv.x * e1, k2 = Bv.x * e2;
                C00 = fmaf(k0, e0, C00); C01 = fmaf(k0, e1, C01); C02 = fmaf(k0, e2, C02);
                C11 = fmaf(k1, e1, C11); C12 = fmaf(k1, e2, C12); C22 = fmaf(k2, e2, C22);
                if (IMG) {
                    const float4 gq = s_go[cc];
                    q0 = fmaf(Bv.w, gq.x, q0); q1 = fmaf(Bv.w, gq.y, q1); q2 = fmaf(Bv.w, gq.z, q2); q3 = fmaf(Bv.w, gq.w, q3);
                }
            }
#pragma unroll
            for (int o = 1; o < 4; o <<= 1) {
                A0 += __shfl_xor_sync(tmask, A0, o); A1 += __shfl_xor_sync(tmask, A1, o); A2 += __shfl_xor_sync(tmask, A2, o);
                bb += __shfl_xor_sync(tmask, bb, o);
                C00 += __shfl_xor_sync(tmask, C00, o); C01 += __shfl_xor_sync(tmask, C01, o); C02 += __shfl_xor_sync(tmask, C02, o);
                C11 += __shfl_xor_sync(tmask, C11, o); C12 += __shfl_xor_sync(tmask, C12, o); C22 += __shfl_xor_sync(tmask, C22, o);
                if (IMG) {
                    q0 += __shfl_xor_sync(tmask, q0, o); q1 += __shfl_xor_sync(tmask, q1, o);
                    q2 += __shfl_xor_sync(tmask, q2, o); q3 += __shfl_xor_sync(tmask, q3, o);
                }
            }
            // the four lanes read the same record (one request) and share the reductions between them
            float v0, v1, v2, S[9];
            load_gauss<KIND>(a.gauss, g, v0, v1, v2, S);
            const float m0 = __fsub_rn(v0, c0), m1 = __fsub_rn(v1, c1), m2 = __fsub_rn(v2, c2);
            if (IMG && mem == 3 && a.grad_attr4 != nullptr)
                atomicAdd(reinterpret_cast<float4*>(a.grad_attr4) + g, make_float4(q0, q1, q2, q3));
            if (mem == 0) {
                // d mu = S A + (S + S^T) mu b
                const float gv0 = S[0] * A0 + S[1] * A1 + S[2] * A2 + bb * (2.f * S[0] * m0 + (S[1] + S[3]) * m1 + (S[2] + S[6]) * m2);
                const float gv1 = S[3] * A0 + S[4] * A1 + S[5] * A2 + bb * ((S[3] + S[1]) * m0 + 2.f * S[4] * m1 + (S[5] + S[7]) * m2);
                const float gv2 = S[6] * A0 + S[7] * A1 + S[8] * A2 + bb * ((S[6] + S[2]) * m0 + (S[7] + S[5]) * m1 + 2.f * S[8] * m2);
                if (KIND == 1) {
                    const float tr = C00 + C11 + C22 + m0 * A0 + m1 * A1 + m2 * A2 + bb * (m0 * m0 + m1 * m1 + m2 * m2);
                    atomicAdd(reinterpret_cast<float4*>(a.grad_packed + 4 * (int64_t)g), make_float4(gv0, gv1, gv2, 2.f * tr));
                } else if (KIND == 3) {
                    atomicAdd(reinterpret_cast<float4*>(a.grad_packed + 8 * (int64_t)g), make_float4(gv0, gv1, gv2, 0.f));
                } else {
                    const float gs0 = a.need_sigma ? 2.f * (C00 + m0 * A0 + bb * m0 * m0) : 0.f;
                    atomicAdd(reinterpret_cast<float4*>(a.grad_packed + 12 * (int64_t)g), make_float4(gv0, gv1, gv2, gs0));
                }
            } else if (mem == 1 && a.need_sigma) {
                if (KIND == 3) {
                    atomicAdd(reinterpret_cast<float4*>(a.grad_packed + 8 * (int64_t)g) + 1,
                              make_float4(2.f * (C00 + m0 * A0 + bb * m0 * m0), 2.f * (C11 + m1 * A1 + bb * m1 * m1),
                                          2.f * (C22 + m2 * A2 + bb * m2 * m2), 0.f));
                } else if (KIND == 9) {
                    // d S_ij = C_ij + mu_i A_j + b mu_i mu_j : entries 01, 02, 10, 11
                    atomicAdd(reinterpret_cast<float4*>(a.grad_packed + 12 * (int64_t)g) + 1,
                              make_float4(2.f * (C01 + m0 * A1 + bb * m0 * m1), 2.f * (C02 + m0 * A2 + bb * m0 * m2),
                                          2.f * (C01 + m1 * A0 + bb * m1 * m0), 2.f * (C11 + m1 * A1 + bb * m1 * m1)));
                }
            } else if (mem == 2 && a.need_sigma && KIND == 9) {
                // entries 12, 20, 21, 22
                atomicAdd(reinterpret_cast<float4*>(a.grad_packed + 12 * (int64_t)g) + 2,
                          make_float4(2.f * (C12 + m1 * A2 + bb * m1 * m2), 2.f * (C02 + m2 * A0 + bb * m2 * m0),
                                      2.f * (C12 + m2 * A1 + bb * m2 * m1), 2.f * (C22 + m2 * A2 + bb * m2 * m2)));
            }
        }
    }
}

}  // namespace voge

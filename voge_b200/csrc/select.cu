// select.cu -- second and third stage of the forward pipeline of the fused renderer (see trace.cu):
//   select_topk_kernel   : per pixel, the K smallest (len, idx) hits of its segment in ascending order
//                          (== the reference's insertion rule, ray_trace_voge.cu:197-213) -> vert_index,
//                          valid_num; register sorting networks on 32-bit composites, exact 64-bit
//                          selection where a composite ties or the segment exceeds the largest network;
//   blend_weights_kernel : per pixel, exact re-evaluation of the survivors (bit-faithful exact_pair),
//                          windowed erf blend (reference VoGE/Aggregation.py:30-107) -> vert_weight,
//                          vert_hit_length (+ optional act / dsd).
#include "../../include/voge_b200.h"
#include "blend_core.cuh"
#include "fine_core.cuh"
#include "render_core.cuh"
#include "sort_net.h"
#include <utility>

namespace voge {

struct SelectArgs {
    const int32_t* counts;        // (B*TY*TX, TNT) hits per pixel column (col = ly*tile + lx)
    const int64_t* seg_base;      // (B*TY*TX, TNT)
    const uint2* hits;            // (orderable len bits, local Gaussian index)
    int B, N, H, W, K, tile, TX, TY;
    int view_base;                // index of this call's first view in the batch (packed indices are (view_base + b)*N + g)
    int32_t* out_idx;             // (B,H,W,K) packed, -1 padded
    int64_t* out_valid;           // (B,H,W)
    unsigned long long* stats;    // optional: [2] pixels selected with the exact 64-bit keys
};

// The K smallest keys of a segment, ascending, are found through 32-bit composites
// ((len bits - smallest len bits of the segment) << 6 | slot): exact as long as the segment's lens span less
// than 2^26 float steps (8 binades) and no two hits share their len; otherwise the caller selects with the full
// (len, idx) keys.
// TWO adjacent lanes per pixel, S slots each (segments of up to 2 S hits; S = 8 / 16 / 32 by the longest segment
// of the warp): the kernel needs the registers of the 32-input network only (twice the resident warps of a
// 64-input one, a third of the code), and the two lanes load interleaved slots (lane `sub` takes slots
// 2 i + sub), so the pair's two 8-byte loads of one instruction are adjacent -- one L1 wavefront instead of two.
// Each lane sorts its half; min / max against the partner's reversed half (lane 0 keeps the S smallest
// composites, lane 1 the S largest, both bitonic) and a bitonic merge per lane give the sorted 2 S
// (tests/csrc/sort_net_check.cpp replays the scheme on the CPU).  Returns false (in both lanes) where the
// composites are not exact, see select_network.
template <int S>
__device__ __forceinline__ bool select_pair(const uint2* __restrict__ hs, int c, int K, int pack_off, int sub,
                                            unsigned pair_mask, unsigned* __restrict__ s_y, int lane,
                                            int32_t* __restrict__ o_idx) {
    unsigned r[S];
    unsigned omin = 0xffffffffu, omax = 0u;
    const int j0 = sub * S;                  // first RANK of this lane after the merge
#pragma unroll
    for (int i = 0; i < S; ++i) {
        r[i] = 0xffffffffu;
        if (2 * i + sub < c) {
            const uint2 h = __ldg(&hs[2 * i + sub]);
            r[i] = h.x;
            s_y[i * 32 + lane] = h.y;          // slot 2 i + sub lives in the column of the lane that loaded it
            omin = min(omin, r[i]);
            omax = max(omax, r[i]);
        }
    }
    omin = min(omin, __shfl_xor_sync(pair_mask, omin, 1));
    omax = max(omax, __shfl_xor_sync(pair_mask, omax, 1));
    if (c > 0 && omax - omin >= 0x3ffffffu) return false;
#pragma unroll
    for (int i = 0; i < S; ++i)
        if (2 * i + sub < c) r[i] = ((r[i] - omin) << 6) | (unsigned)(2 * i + sub);
    sort_network<S>(r);
#pragma unroll
    for (int x = 0; x < S / 2; ++x) {
        const unsigned t1 = __shfl_xor_sync(pair_mask, r[S - 1 - x], 1), t2 = __shfl_xor_sync(pair_mask, r[x], 1);
        r[x] = sub ? max(r[x], t1) : min(r[x], t1);
        r[S - 1 - x] = sub ? max(r[S - 1 - x], t2) : min(r[S - 1 - x], t2);
    }
    bitonic_merge<S>(r);
    // ranks sub*S + i; a tie is two neighbouring ranks below c with equal len bits
    const unsigned below = __shfl_xor_sync(pair_mask, r[S - 1], 1);   // lane 1: rank S - 1
    bool tie = sub && (S < c) && ((r[0] ^ below) < 64u);
#pragma unroll
    for (int i = 1; i < S; ++i) tie = tie || ((j0 + i < c) && ((r[i] ^ r[i - 1]) < 64u));
    tie = __shfl_xor_sync(pair_mask, (int)tie, 1) || tie;
    if (tie) return false;
    __syncwarp(pair_mask);                                      // the partner's index column is read below
    const int m = min(c, K);
    const bool vec = (K & 3) == 0;
#pragma unroll
    for (int i0 = 0; i0 < S; i0 += 4) {
        if (j0 + i0 < K) {
            int v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned slot = r[i0 + j] & 63u;
                v[j] = (j0 + i0 + j < m) ? pack_off + (int)s_y[(slot >> 1) * 32 + ((lane & 30) | (int)(slot & 1u))] : -1;
            }
            if (vec) {
                *reinterpret_cast<int4*>(o_idx + j0 + i0) = make_int4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j0 + i0 + j < K) o_idx[j0 + i0 + j] = v[j];
            }
        }
    }
    for (int k = 2 * S + sub; k < K; k += 2) o_idx[k] = -1;
    return true;
}

// FOUR adjacent lanes per pixel, 32 slots each: segments of 65 .. 128 hits (mesh-converted scenes with K = 40 .. 60:
// 8 % of the covered pixels of the RenderBunny-sized C2 scene), K <= 63.  Lane q takes the slots 4 i + q (the quad's
// four 8-byte loads are adjacent); every lane sorts its 32, the pairs (0,1) and (2,3) merge to two sorted runs of
// 64 as in select_pair, the pair (0,1) then keeps min(A[i], B[63 - i]) -- the 64 smallest of the 128, a bitonic
// sequence -- and a half-cleaner across its two lanes plus a bitonic merge per lane sort them.  Composites carry 7
// slot bits, so the lens may span 2^25 float steps.  Returns false (in all four lanes) where they are not exact.
__device__ __noinline__ bool select_quad(const uint2* __restrict__ hs, int c, int K, int pack_off, int q,
                                            unsigned quad_mask, unsigned* __restrict__ s_y, int lane,
                                            int32_t* __restrict__ o_idx) {
    constexpr int S = 32;
    unsigned r[S];
    unsigned omin = 0xffffffffu, omax = 0u;
#pragma unroll
    for (int i = 0; i < S; ++i) {
        r[i] = 0xffffffffu;
        if (4 * i + q < c) {
            const uint2 h = __ldg(&hs[4 * i + q]);
            r[i] = h.x;
            s_y[i * 32 + lane] = h.y;          // slot 4 i + q lives in the column of the lane that loaded it
            omin = min(omin, r[i]);
            omax = max(omax, r[i]);
        }
    }
    omin = min(omin, __shfl_xor_sync(quad_mask, omin, 1)); omin = min(omin, __shfl_xor_sync(quad_mask, omin, 2));
    omax = max(omax, __shfl_xor_sync(quad_mask, omax, 1)); omax = max(omax, __shfl_xor_sync(quad_mask, omax, 2));
    if (omax - omin >= 0x1ffffffu) return false;
#pragma unroll
    for (int i = 0; i < S; ++i)
        if (4 * i + q < c) r[i] = ((r[i] - omin) << 7) | (unsigned)(4 * i + q);
    sort_network<S>(r);
    const int sub = q & 1;
    // pairs (0,1) and (2,3): sorted runs of 64 (lane sub = 0: ranks 0..31 of its pair, sub = 1: ranks 32..63)
#pragma unroll
    for (int x = 0; x < S / 2; ++x) {
        const unsigned t1 = __shfl_xor_sync(quad_mask, r[S - 1 - x], 1), t2 = __shfl_xor_sync(quad_mask, r[x], 1);
        r[x] = sub ? max(r[x], t1) : min(r[x], t1);
        r[S - 1 - x] = sub ? max(r[S - 1 - x], t2) : min(r[S - 1 - x], t2);
    }
    bitonic_merge<S>(r);
    // A = pair (0,1), B = pair (2,3): lane 0 register x = A[x] meets B[63 - x] = lane 3 register 31 - x, lane 1
    // register x = A[32 + x] meets B[31 - x] = lane 2 register 31 - x; the other pair's result is not needed
#pragma unroll
    for (int x = 0; x < S; ++x) {
        const unsigned t = __shfl_xor_sync(quad_mask, r[S - 1 - x], 3);
        r[x] = min(r[x], t);
    }
    // half-cleaner across the lanes of pair (0,1), then a bitonic merge per lane
#pragma unroll
    for (int x = 0; x < S; ++x) {
        const unsigned t = __shfl_xor_sync(quad_mask, r[x], 1);
        r[x] = sub ? max(r[x], t) : min(r[x], t);
    }
    bitonic_merge<S>(r);
    // lanes 0 / 1 hold ranks 0..31 / 32..63; a tie is two neighbouring ranks with equal len bits
    const int j0 = sub * S;
    const unsigned below = __shfl_xor_sync(quad_mask, r[S - 1], 1);
    bool tie = sub && ((r[0] ^ below) < 128u);
#pragma unroll
    for (int i = 1; i < S; ++i) tie = tie || ((r[i] ^ r[i - 1]) < 128u);
    tie = tie && q < 2;
    tie = __shfl_xor_sync(quad_mask, (int)tie, 1) || tie;
    tie = __shfl_xor_sync(quad_mask, (int)tie, 2) || tie;
    if (tie) return false;
    __syncwarp(quad_mask);                                      // the other lanes' index columns are read below
    if (q < 2) {
        const bool vec = (K & 3) == 0;
#pragma unroll
        for (int i0 = 0; i0 < S; i0 += 4) {
            if (j0 + i0 < K) {
                int v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned slot = r[i0 + j] & 127u;
                    v[j] = (j0 + i0 + j < K) ? pack_off + (int)s_y[(slot >> 2) * 32 + ((lane & 28) | (int)(slot & 3u))] : -1;
                }
                if (vec) {
                    *reinterpret_cast<int4*>(o_idx + j0 + i0) = make_int4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j0 + i0 + j < K) o_idx[j0 + i0 + j] = v[j];
                }
            }
        }
    }
    return true;
}

// exact selection with the full (len, idx) keys: K passes, each extracting the smallest key above the
// previous one (keys are unique: a Gaussian hits a pixel at most once).  O(K c) loads, any c.
__device__ __noinline__ void select_exact(const uint2* __restrict__ hs, int c, int K, int pack_off,
                                          int32_t* __restrict__ o_idx) {
    unsigned long long prev = 0ull;
    const int m = min(c, K);
    for (int k = 0; k < m; ++k) {
        unsigned long long best = 0xffffffffffffffffull;
        for (int j = 0; j < c; ++j) {
            const uint2 h = hs[j];
            const unsigned long long key = ((unsigned long long)h.x << 32) | h.y;
            if ((k == 0 || key > prev) && key < best) best = key;
        }
        o_idx[k] = pack_off + (int)(unsigned)(best & 0xffffffffull);
        prev = best;
    }
    for (int k = m; k < K; ++k) o_idx[k] = -1;
}

// The same selection by a whole warp for ONE pixel: lanes stride over the segment (coalesced loads, L1 hits after
// the first pass), the lexicographic (len, idx) minimum above the previous key is found with two warp
// reductions (REDUX.MIN on the len bits, then on the indices of the lanes that hold that len), and the K results
// are written 32 at a time, one per lane.  O(K c / 32) loads per lane instead of O(K c).
__device__ __noinline__ void select_exact_warp(const uint2* __restrict__ hs, int c, int K, int pack_off,
                                               int32_t* __restrict__ o_idx, int lane) {
    unsigned prev_x = 0u, prev_y = 0u;
    const int m = min(c, K);
    int keep = -1;
    for (int k = 0; k < m; ++k) {
        unsigned bx = 0xffffffffu, by = 0xffffffffu;
        for (int j = lane; j < c; j += 32) {
            const uint2 h = hs[j];
            const bool above = (k == 0) || h.x > prev_x || (h.x == prev_x && h.y > prev_y);
            if (above && (h.x < bx || (h.x == bx && h.y < by))) { bx = h.x; by = h.y; }
        }
        const unsigned mx = __reduce_min_sync(0xffffffffu, bx);
        const unsigned my = __reduce_min_sync(0xffffffffu, bx == mx ? by : 0xffffffffu);
        prev_x = mx; prev_y = my;
        if ((k & 31) == lane) keep = pack_off + (int)my;
        if ((k & 31) == 31 || k == m - 1) {
            const int k0 = k & ~31;
            if (k0 + lane <= k) o_idx[k0 + lane] = keep;
        }
    }
    for (int k = m + lane; k < K; k += 32) o_idx[k] = -1;
}

// thread -> pixel of a tile: 8x4 pixel blocks per warp on 16x16 tiles (neighbouring pixels see the same
// Gaussians), row-major otherwise
template <int TNT>
__device__ __forceinline__ void thread_to_pix(int t, int tile, int& lx, int& ly, bool& in_tile) {
    if (TNT == 256 && tile == 16) {
        const int w = t >> 5, l = t & 31;
        lx = (w & 1) * 8 + (l & 7);
        ly = (w >> 1) * 4 + (l >> 3);
        in_tile = true;
    } else {
        lx = t % tile; ly = t / tile;
        in_tile = t < tile * tile;
    }
}

// pixel `ti` (thread index within the tile) of tile (b, ty, tx): hit count, segment, ray index.  The kernel
// re-derives these where it needs them instead of carrying them (registers are what bounds its occupancy).
struct SelPix {
    bool live;
    int c;
    const uint2* hs;
    int64_t ray;
};
template <int TNT>
__device__ __forceinline__ SelPix select_pixel(const SelectArgs& a, int64_t tile_id, int ti) {
    const int t = (int)(tile_id % ((int64_t)a.TX * a.TY));
    const int b = (int)(tile_id / ((int64_t)a.TX * a.TY));
    const int tx = t % a.TX, ty = t / a.TX;
    int lx, ly;
    bool in_tile;
    thread_to_pix<TNT>(ti, a.tile, lx, ly, in_tile);
    const int col = ly * a.tile + lx;
    const int xi = tx * a.tile + lx, yi = ty * a.tile + ly;
    SelPix p;
    p.live = in_tile && xi < a.W && yi < a.H;
    p.c = p.live ? a.counts[tile_id * TNT + col] : 0;
    p.hs = a.hits + (p.live ? a.seg_base[tile_id * TNT + col] : 0);
    p.ray = p.live ? ((int64_t)b * a.H + yi) * a.W + xi : 0;
    return p;
}

template <int NT, int TNT>
__global__ void __launch_bounds__(NT, 1024 / NT) select_topk_kernel(const SelectArgs a) {
    __shared__ unsigned s_idx[NT / 32][32 * 32];       // per warp: index halves of the loaded hits, [slot][lane]
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    unsigned* s_y = s_idx[tid >> 5];
    constexpr int PARTS = TNT / NT;
    const int64_t tile_id = blockIdx.x / PARTS;
    const int ti0 = (blockIdx.x % PARTS) * NT + (tid & ~31);      // first pixel of this warp within the tile
    const int pack_off = (a.view_base + (int)(tile_id / ((int64_t)a.TX * a.TY))) * a.N;
    bool done = true;
    int wmax;
    {
        const SelPix me = select_pixel<TNT>(a, tile_id, ti0 + lane);
        wmax = __reduce_max_sync(0xffffffffu, me.c);
        if (me.live) {
            a.out_valid[me.ray] = min(me.c, a.K);
            if (wmax == 0) {              // no hit in any pixel of the warp: padding only
                int32_t* o = a.out_idx + me.ray * a.K;
                if ((a.K & 3) == 0) {
                    for (int k = 0; k < a.K; k += 4) *reinterpret_cast<int4*>(o + k) = make_int4(-1, -1, -1, -1);
                } else {
                    for (int k = 0; k < a.K; ++k) o[k] = -1;
                }
            }
        }
    }
    if (wmax == 0) return;
    {
        // two lanes per pixel, 16 pixels per pass; pixels with more than 64 hits go to the exact selection
        const int sub = lane & 1;
        const unsigned pair_mask = 3u << (lane & 30);
        unsigned failed[2];
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            const SelPix px = select_pixel<TNT>(a, tile_id, ti0 + pass * 16 + (lane >> 1));
            bool ok = true;
            if (px.live) {
                int32_t* o_idx = a.out_idx + px.ray * a.K;
                if (wmax <= 16) ok = select_pair<8>(px.hs, px.c, a.K, pack_off, sub, pair_mask, s_y, lane, o_idx);
                else if (wmax <= 32) ok = select_pair<16>(px.hs, px.c, a.K, pack_off, sub, pair_mask, s_y, lane, o_idx);
                else ok = px.c <= 64 && select_pair<32>(px.hs, px.c, a.K, pack_off, sub, pair_mask, s_y, lane, o_idx);
            }
            failed[pass] = __ballot_sync(0xffffffffu, !ok);
            __syncwarp();                                        // pass 1 overwrites the index columns
        }
        done = !(((lane < 16 ? failed[0] : failed[1]) >> (2 * (lane & 15))) & 1u);
    }
    if (__ballot_sync(0xffffffffu, !done) == 0u) return;
    // pixels of 65 .. 128 hits: four lanes each, eight pixels per pass
    if (a.K <= 63) {
        const SelPix me = select_pixel<TNT>(a, tile_id, ti0 + lane);
        unsigned want = __ballot_sync(0xffffffffu, !done && me.c > 64 && me.c <= 128);
        if (want != 0u) {
            const int q = lane & 3;
            const unsigned quad_mask = 15u << (lane & 28);
#pragma unroll 1
            for (int pass = 0; pass < 4; ++pass) {
                if (((want >> (8 * pass)) & 0xffu) == 0u) continue;
                const int p = pass * 8 + (lane >> 2);
                bool ok = true;
                if ((want >> p) & 1u) {
                    const SelPix px = select_pixel<TNT>(a, tile_id, ti0 + p);
                    ok = select_quad(px.hs, px.c, a.K, pack_off, q, quad_mask, s_y, lane, a.out_idx + px.ray * a.K);
                } else {
                    ok = false;
                }
                const unsigned fin = __ballot_sync(0xffffffffu, ok);     // bit of lane 4 (p % 8) set: pixel p finished
                if (pass == (lane >> 3) && ((fin >> (4 * (lane & 7))) & 1u)) done = true;
                __syncwarp();                                            // the next pass overwrites the index columns
            }
        }
    }
    // pixels the networks could not finish (more than 128 hits, tied lens, lens spanning > 8 binades): a few
    // per warp are selected by the whole warp, many (a warp inside a very dense region) by their own lanes
    unsigned slow = __ballot_sync(0xffffffffu, !done);
    if (slow == 0u) return;
    if (a.stats != nullptr && lane == 0) atomicAdd(a.stats + 2, (unsigned long long)__popc(slow));
    if (__popc(slow) > 8) {
        if (!done) {
            const SelPix me = select_pixel<TNT>(a, tile_id, ti0 + lane);
            select_exact(me.hs, me.c, a.K, pack_off, a.out_idx + me.ray * a.K);
        }
        return;
    }
    while (slow != 0u) {
        const int p = __ffs(slow) - 1;
        slow &= slow - 1u;
        const SelPix px = select_pixel<TNT>(a, tile_id, ti0 + p);
        select_exact_warp(px.hs, px.c, a.K, pack_off, a.out_idx + px.ray * a.K, lane);
    }
}

template <int NT, int TNT>
static int launch_select(const SelectArgs& a, cudaStream_t stream) {
    const long long grid = (long long)a.B * a.TX * a.TY * (TNT / NT);
    if (grid <= 0 || grid > 2147483647LL) return (int)cudaErrorInvalidValue;
    select_topk_kernel<NT, TNT><<<(unsigned)grid, NT, 0, stream>>>(a);
    VOGE_LAUNCH_CHECK();
    return 0;
}

// ---- blend -------------------------------------------------------------------------------------------------
struct BlendArgs {
    const float* gauss;           // packed records (voge_pack_gaussians)
    const float* origins;
    const float* rays;            // (B,H,W,3), or NULL: generated from `cam`
    const float* cam;             // (B,16) per-view camera records (render_core.cuh), read when rays == NULL
    const int32_t* idx;           // (B,H,W,K) packed, first valid[r] slots
    const int64_t* valid;         // (B,H,W)
    float omega;
    int B, N, H, W, K;
    int view_base;                // as in SelectArgs
    int enc;                      // kind-9 records carry the isotropic encoding (render_core.cuh: kKindIsoEncoded)
    float* out_weight;            // (B,H,W,K)
    float* out_len;               // (B,H,W,K), 1e10 padded
    float* out_act;               // optional (B,H,W,K)
    float* out_dsd;               // optional (B,H,W,K)
};

template <int NT, int KIND>
__global__ void __launch_bounds__(NT) blend_weights_kernel(const BlendArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_ls = reinterpret_cast<float2*>(smem_raw);                 // [K][NT] (len, sqrt(dsd + 1e-10))
    float* s_E = reinterpret_cast<float*>(s_ls + (size_t)a.K * NT);     // [K][NT] exp(-act)
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
    if (xi >= a.W || yi >= a.H) return;
    const int64_t ray = ((int64_t)b * a.H + yi) * a.W + xi;
    const int cnt = (int)min((int64_t)a.K, a.valid[ray]);

    // ---- exact (len, act, dsd) of the survivors, blend weights, fragment write-out ----
    // Each thread owns a (K,) row of every fragment tensor; rows are written four slots at a time as
    // 16-byte vectors when K % 4 == 0 (a scalar store at a 4K-byte lane stride costs one L1/L2 sector
    // operation per lane and slot).
    const int32_t* i_idx = a.idx + ray * a.K;
    float* o_len = a.out_len + ray * a.K;
    float* o_w = a.out_weight + ray * a.K;
    const bool vec = (a.K & 3) == 0;
    if (cnt == 0) {
        // empty pixel: padding only
        for (int k0 = 0; k0 < a.K; k0 += 4) {
            if (vec) {
                *reinterpret_cast<float4*>(o_len + k0) = make_float4(kEmptyLen, kEmptyLen, kEmptyLen, kEmptyLen);
                *reinterpret_cast<float4*>(o_w + k0) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.out_act != nullptr) {
                    *reinterpret_cast<float4*>(a.out_act + ray * a.K + k0) = make_float4(kEmptyLen, kEmptyLen, kEmptyLen, kEmptyLen);
                    *reinterpret_cast<float4*>(a.out_dsd + ray * a.K + k0) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                for (int k = k0; k < min(k0 + 4, a.K); ++k) {
                    o_len[k] = kEmptyLen; o_w[k] = 0.f;
                    if (a.out_act != nullptr) { a.out_act[ray * a.K + k] = kEmptyLen; a.out_dsd[ray * a.K + k] = 0.f; }
                }
            }
        }
        return;
    }
    const float c0 = a.origins[3 * b], c1 = a.origins[3 * b + 1], c2 = a.origins[3 * b + 2];
    float r0, r1, r2;
    pixel_ray(a.rays, a.cam, b, xi, yi, a.H, a.W, r0, r1, r2);
    const int pack_off = (a.view_base + b) * a.N;
    float s_min = 3.0e38f;
    for (int k0 = 0; k0 < a.K; k0 += 4) {
        int gv[4] = {-1, -1, -1, -1};
        if (vec) {
            if (k0 < cnt) {
                const int4 q = *reinterpret_cast<const int4*>(i_idx + k0);
                gv[0] = q.x; gv[1] = q.y; gv[2] = q.z; gv[3] = q.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k0 + j < cnt) gv[j] = i_idx[k0 + j];
        }
        float lv[4], av[4], dv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            lv[j] = kEmptyLen; av[j] = kEmptyLen; dv[j] = 0.f;
            if (k < cnt) {
                const Hit h = exact_hit_packed<KIND>(a.gauss, gv[j] - pack_off, c0, c1, c2, r0, r1, r2, a.enc != 0);
                lv[j] = h.len; av[j] = h.act; dv[j] = h.dsd;
                const float sk = sqrtf(h.dsd + 1e-10f);                              // Aggregation.py:49
                s_ls[k * NT + tid] = make_float2(h.len, sk);
                s_E[k * NT + tid] = expf(-h.act);
                s_min = fminf(s_min, sk);
            }
        }
        if (vec) {
            *reinterpret_cast<float4*>(o_len + k0) = make_float4(lv[0], lv[1], lv[2], lv[3]);
            if (a.out_act != nullptr) {
                *reinterpret_cast<float4*>(a.out_act + ray * a.K + k0) = make_float4(av[0], av[1], av[2], av[3]);
                *reinterpret_cast<float4*>(a.out_dsd + ray * a.K + k0) = make_float4(dv[0], dv[1], dv[2], dv[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (k0 + j >= a.K) break;
                o_len[k0 + j] = lv[j];
                if (a.out_act != nullptr) { a.out_act[ray * a.K + k0 + j] = av[j]; a.out_dsd[ray * a.K + k0 + j] = dv[j]; }
            }
        }
    }
    // D_m = sum_k E_k Phi((len_m - len_k) s_k).  The list is sorted by len, so outside the window
    // |len_m - len_k| * min_k(s_k) < 4 the erf is saturated: Phi = 1 for k < lo(m) (their E_k are
    // carried in a running prefix sum -- lo(m) only moves forward), Phi = 0 behind the window.
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
                    // the k = m term is E_m / 2 exactly; the loop runs over the neighbours only, so a warp
                    // iterates max-over-lanes(neighbours) times instead of max-over-lanes(neighbours) + 1
                    const float Em = s_E[m * NT + tid];
                    float D = fmaf(Em, 0.5f, SE);
                    for (int t = lo;; ++t) {
                        const int k = t + (t >= m ? 1 : 0);
                        if (k >= cnt) break;
                        const float2 lk = s_ls[k * NT + tid];
                        const float dl = lm - lk.x;
                        if (dl * s_min <= -kErfSat) break;
                        D += s_E[k * NT + tid] * phi_f(dl * lk.y);
                    }
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
}

// Two threads per pixel (adjacent lanes; a warp covers a 4x4 pixel block), as in the fused backward: the
// per-pixel arrays in shared memory bound the resident pixels, so the pair doubles the warps that hide the
// latency of the gathers and of the serial blend loops.  Thread `sub` owns the slots k = sub, sub + 2, ...;
// the (K,) rows are still written as 16-byte vectors after a two-value exchange inside the pair (K % 4 == 0).
template <int NT, int KIND>
__global__ void __launch_bounds__(NT) blend_pair_kernel(const BlendArgs a) {
    constexpr int NP = NT / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_ls = reinterpret_cast<float2*>(smem_raw);                 // [K][NP] (len, sqrt(dsd + 1e-10))
    float* s_E = reinterpret_cast<float*>(s_ls + (size_t)a.K * NP);     // [K][NP] exp(-act)
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
    if (xi >= a.W || yi >= a.H) return;
    const int64_t ray = ((int64_t)b * a.H + yi) * a.W + xi;
    const int cnt = (int)min((int64_t)a.K, a.valid[ray]);
    const int32_t* i_idx = a.idx + ray * a.K;
    float* o_len = a.out_len + ray * a.K;
    float* o_w = a.out_weight + ray * a.K;
    float* o_act = a.out_act != nullptr ? a.out_act + ray * a.K : nullptr;
    float* o_dsd = a.out_dsd != nullptr ? a.out_dsd + ray * a.K : nullptr;
    if (cnt == 0) {
        // empty pixel: padding only (thread 0 of the pair: len / dsd rows, thread 1: weight / act rows)
        const float4 e4 = make_float4(kEmptyLen, kEmptyLen, kEmptyLen, kEmptyLen), z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < a.K; k += 4) {
            if (sub == 0) { *reinterpret_cast<float4*>(o_len + k) = e4; if (o_dsd != nullptr) *reinterpret_cast<float4*>(o_dsd + k) = z4; }
            else { *reinterpret_cast<float4*>(o_w + k) = z4; if (o_act != nullptr) *reinterpret_cast<float4*>(o_act + k) = e4; }
        }
        return;
    }
    const float c0 = a.origins[3 * b], c1 = a.origins[3 * b + 1], c2 = a.origins[3 * b + 2];
    float r0, r1, r2;
    pixel_ray(a.rays, a.cam, b, xi, yi, a.H, a.W, r0, r1, r2);
    const int pack_off = (a.view_base + b) * a.N;
    float s_min = 3.0e38f;
    // ---- exact (len, act, dsd) of this thread's slots ----
    for (int k0 = 0; k0 < a.K; k0 += 4) {
        float lv[2] = {kEmptyLen, kEmptyLen}, av[2] = {kEmptyLen, kEmptyLen}, dv[2] = {0.f, 0.f};
        if (k0 < cnt) {
            const int4 q = *reinterpret_cast<const int4*>(i_idx + k0);
            const int gv[2] = {sub ? q.y : q.x, sub ? q.w : q.z};
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int k = k0 + sub + 2 * jj;
                if (k < cnt) {
                    const Hit h = exact_hit_packed<KIND>(a.gauss, gv[jj] - pack_off, c0, c1, c2, r0, r1, r2, a.enc != 0);
                    lv[jj] = h.len; av[jj] = h.act; dv[jj] = h.dsd;
                    const float sk = sqrtf(h.dsd + 1e-10f);                              // Aggregation.py:49
                    s_ls[k * NP + col] = make_float2(h.len, sk);
                    s_E[k * NP + col] = expf(-h.act);
                    s_min = fminf(s_min, sk);
                }
            }
        }
        // thread 0 holds slots (k0, k0+2), thread 1 (k0+1, k0+3): exchange and write whole rows of four
        const float pl0 = __shfl_xor_sync(pair_mask, lv[0], 1), pl1 = __shfl_xor_sync(pair_mask, lv[1], 1);
        if (sub == 0) *reinterpret_cast<float4*>(o_len + k0) = make_float4(lv[0], pl0, lv[1], pl1);
        if (o_act != nullptr) {
            const float pa0 = __shfl_xor_sync(pair_mask, av[0], 1), pa1 = __shfl_xor_sync(pair_mask, av[1], 1);
            const float pd0 = __shfl_xor_sync(pair_mask, dv[0], 1), pd1 = __shfl_xor_sync(pair_mask, dv[1], 1);
            if (sub == 1) *reinterpret_cast<float4*>(o_act + k0) = make_float4(pa0, av[0], pa1, av[1]);
            if (sub == 0) *reinterpret_cast<float4*>(o_dsd + k0) = make_float4(dv[0], pd0, dv[1], pd1);
        }
    }
    s_min = fminf(s_min, __shfl_xor_sync(pair_mask, s_min, 1));
    __syncwarp(pair_mask);
    // ---- blend weights of this thread's slots (see blend_weights_kernel) ----
    {
        int lo = 0;
        float SE = 0.f;
        for (int m0 = 0; m0 < a.K; m0 += 4) {
            float wv[2] = {0.f, 0.f};
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int m = m0 + sub + 2 * jj;
                if (m < cnt) {
                    const float lm = s_ls[m * NP + col].x;
                    while (lo < m && (lm - s_ls[lo * NP + col].x) * s_min >= kErfSat) { SE += s_E[lo * NP + col]; ++lo; }
                    const float Em = s_E[m * NP + col];
                    float D = fmaf(Em, 0.5f, SE);
                    for (int t = lo;; ++t) {
                        const int k = t + (t >= m ? 1 : 0);
                        if (k >= cnt) break;
                        const float2 lk = s_ls[k * NP + col];
                        const float dl = lm - lk.x;
                        if (dl * s_min <= -kErfSat) break;
                        D += s_E[k * NP + col] * phi_f(dl * lk.y);
                    }
                    wv[jj] = Em != 0.f ? expf(-(D * a.omega)) * Em * kInvExpMinusHalf : 0.f;
                }
            }
            const float pw0 = __shfl_xor_sync(pair_mask, wv[0], 1), pw1 = __shfl_xor_sync(pair_mask, wv[1], 1);
            if (sub == 1) *reinterpret_cast<float4*>(o_w + m0) = make_float4(pw0, wv[0], pw1, wv[1]);
        }
    }
}

template <int NT, int KIND>
static int launch_blend(const BlendArgs& a, cudaStream_t stream) {
    const size_t smem = (size_t)a.K * NT * 12;
    if (smem > 227 * 1024) return (int)cudaErrorInvalidValue;
    VOGE_CUDA_TRY(cudaFuncSetAttribute(blend_weights_kernel<NT, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t warps = (int64_t)a.B * cdiv(a.W, 8) * cdiv(a.H, 4);
    const int64_t grid = (warps * 32 + NT - 1) / NT;
    if (grid <= 0 || grid > 2147483647LL) return (int)cudaErrorInvalidValue;
    blend_weights_kernel<NT, KIND><<<(unsigned)grid, NT, smem, stream>>>(a);
    VOGE_LAUNCH_CHECK();
    return 0;
}

template <int KIND>
static int launch_blend_pair(const BlendArgs& a, cudaStream_t stream) {
    constexpr int NT = 128;
    const size_t smem = (size_t)a.K * (NT / 2) * 12;
    VOGE_CUDA_TRY(cudaFuncSetAttribute(blend_pair_kernel<NT, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t warps = (int64_t)a.B * cdiv(a.W, 4) * cdiv(a.H, 4);
    const int64_t grid = (warps * 32 + NT - 1) / NT;
    if (grid <= 0 || grid > 2147483647LL) return (int)cudaErrorInvalidValue;
    blend_pair_kernel<NT, KIND><<<(unsigned)grid, NT, smem, stream>>>(a);
    VOGE_LAUNCH_CHECK();
    return 0;
}

template <int KIND>
static int dispatch_blend(const BlendArgs& a, cudaStream_t s) {
    if ((a.K & 3) == 0 && a.K <= 128) return launch_blend_pair<KIND>(a, s);
    // 12 K bytes of shared memory per thread: 128-thread CTAs while several of them fit on an SM
    if (a.K <= 48) return launch_blend<128, KIND>(a, s);
    if (a.K <= 280) return launch_blend<64, KIND>(a, s);
    return launch_blend<32, KIND>(a, s);
}

}  // namespace voge

extern "C" int voge_select_topk(const int32_t* counts, const int64_t* seg_base, const uint32_t* hits, int view_base, int B, int N, int H, int W, int K, int tile,
                                int32_t* out_idx, int64_t* out_valid, uint64_t* stats, voge_stream_t stream) {
    using namespace voge;
    if (B <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
    if (tile < 1 || tile > 16) return (int)cudaErrorInvalidValue;
    SelectArgs a;
    a.counts = counts; a.seg_base = seg_base; a.hits = reinterpret_cast<const uint2*>(hits);
    a.B = B; a.N = N; a.H = H; a.W = W; a.K = K; a.tile = tile; a.TX = cdiv(W, tile); a.TY = cdiv(H, tile);
    a.view_base = view_base;
    a.out_idx = out_idx; a.out_valid = out_valid; a.stats = reinterpret_cast<unsigned long long*>(stats);
    cudaStream_t s = (cudaStream_t)stream;
    const int nt = tile_threads(tile);
    if (nt == 256) return launch_select<128, 256>(a, s);
    if (nt == 128) return launch_select<128, 128>(a, s);
    return launch_select<64, 64>(a, s);
}

extern "C" int voge_blend_weights(const float* gauss, int sigma_kind, const float* origins,
                                  const float* rays, const float* cam, const int32_t* idx, const int64_t* valid, float absorptivity,
                                  int view_base, int B, int N, int H, int W, int K, float* out_weight, float* out_len,
                                  float* out_act, float* out_dsd, voge_stream_t stream) {
    using namespace voge;
    if (B <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
    BlendArgs a;
    if (rays == nullptr && cam == nullptr) return (int)cudaErrorInvalidValue;
    a.gauss = gauss; a.origins = origins; a.rays = rays; a.cam = cam; a.idx = idx; a.valid = valid;
    a.omega = absorptivity; a.B = B; a.N = N; a.H = H; a.W = W; a.K = K; a.view_base = view_base;
    a.out_weight = out_weight; a.out_len = out_len; a.out_act = out_act; a.out_dsd = out_dsd;
    a.enc = (sigma_kind & kKindIsoEncoded) ? 1 : 0;
    sigma_kind &= ~kKindIsoEncoded;
    if (a.enc && sigma_kind != 9) return (int)cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)stream;
    if (sigma_kind == 1) return dispatch_blend<1>(a, s);
    if (sigma_kind == 3) return dispatch_blend<3>(a, s);
    if (sigma_kind == 9) return dispatch_blend<9>(a, s);
    return (int)cudaErrorInvalidValue;
}

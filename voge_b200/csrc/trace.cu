// trace.cu -- first stage of the forward pipeline of the fused renderer (GaussianRenderer.forward, reference
// VoGE/Renderer.py:130-150 -> RayTraceFineVogeKernel ray_trace_voge.cu:135-217 -> Aggregation.py:82-107):
//
//   voge_trace_hits    : per image tile, evaluate every ITEM (candidate Gaussian, pixel of its conservative
//                        pixel rectangle) with the bit-faithful arithmetic and append the hits (act < thr_act,
//                        len < 1e10) as (orderable len bits, Gaussian index) to the pixel's segment;
//   voge_select_topk   : (select.cu) per pixel, the K smallest (len, idx) keys, ascending, via register
//                        sorting networks -> Fragments.vert_index / valid_num;
//   voge_blend_weights : (select.cu) exact re-evaluation of the K survivors, windowed erf blend ->
//                        Fragments.vert_weight / vert_hit_length.
//
// The items are defined by the per-(view, Gaussian) pixel rectangles of bin_count_kernel (reference coarse
// bin rectangle AND the tangent bound of {act < thr + margin}); tiles only distribute the work.  A pixel's
// segment holds one slot per rectangle covering it (an upper bound of its hits that needs no evaluation):
// the tile's segment area starts at tile_item_offsets[tile] (exclusive scan of bin_count's tile_items) and
// the per-pixel split is a 2D prefix sum of the rectangle corners, so there is no capacity limit and no
// overflow path.  The hot loop is Gaussian-major: a chunk of candidates is set up by one thread each (record
// in shared memory, block scan of the rectangle areas), then warps take blocks of 32 consecutive items, so
// every lane is busy whatever the rectangle sizes.  Hit order inside a segment is arbitrary (atomic slot
// counter); the selection uses the total order (len, idx) == the reference's "strict <, earlier candidate
// wins" insertion rule (ray_trace_voge.cu:197-213), so the fragments are bit-identical to the op-by-op path.
#include "../../include/voge_b200.h"
#include "fine_core.cuh"
#include "render_core.cuh"

namespace voge {

struct TraceArgs {
    const float* gauss;                // packed records (voge_pack_gaussians)
    const float* origins;              // (B,3)
    const float* rays;                 // (B,H,W,3), or NULL: generated from `cam` (render_core.cuh: gen_ray)
    const float* cam;                  // (B,16) per-view camera records, read when rays == NULL
    const int64_t* tile_offsets;       // (B*TY*TX*kBinSub + 1) into tile_list
    const uint4* tile_list;            // 32-byte entries (local Gaussian index, rectangle x, rectangle y, - | first 16 bytes of the record), voge_bin_fill
    const uint2* rects;                // (B,N) conservative pixel rectangles from bin_count_kernel
    const int64_t* tile_item_offsets;  // (B*TY*TX*kBinSub + 1) exclusive scan of tile_items
    int64_t item_base;                 // subtracted from tile_item_offsets: first slot of this call's views in `hits`; < 0: tile_item_offsets[0]
    int64_t list_capacity;             // entries allocated in tile_list / slots allocated in hits when the caller sized them
    int64_t hits_capacity;             // speculatively (<= 0: exact sizes): a tile that would cross either is skipped
    float thr_act;
    int B, N, H, W, tile, TX, TY;
    int enc;                           // kind-9 records carry the isotropic encoding (render_core.cuh: kKindIsoEncoded)
    int32_t* counts;                   // out (B*TY*TX, NT): hits stored per pixel column (col = ly*tile + lx)
    int64_t* seg_base;                 // out (B*TY*TX, NT): first slot of the pixel's segment
    uint2* hits;                       // out (total items): (orderable len bits, local Gaussian index)
    unsigned long long* stats;         // optional: [0] items evaluated
};

template <int KIND>
struct RecFloats {
    // KIND 9: S (9), t = rn(mu_i S_ij) (9), msm, idx ; diagonal kinds: s (3), t (3), msm, idx
    static constexpr int v = (KIND == 9) ? 20 : 8;
};

constexpr int kSegAlign = 4;                   // pixel segments start on multiples of this many 8-byte slots
constexpr int kGroup = 4;                      // consecutive items of one candidate evaluated by one lane
constexpr int kGroupBudget = 4096;             // item groups per chunk (bounds the block tables)
constexpr int kMaxBlocks = kGroupBudget / 32;

static inline size_t trace_smem_bytes(int nt, int rec_floats) {
    // two buffers of {records, meta (NT+2), block masks, block owners}; rays 12, counters 4, bases 4 per
    // column; scan scratch
    return 2 * ((size_t)nt * rec_floats * 4 + (size_t)(nt + 2) * 8 + (size_t)kMaxBlocks * 6) + (size_t)nt * 20 + 128;
}

// inclusive block scan of one int per thread; s_w: NT/32 ints of scratch; returns (inclusive prefix, total)
template <int NT>
__device__ __forceinline__ int2 block_scan(int v, int* s_w, int lane, int warp) {
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int q = 0; q < NT / 32; ++q) {
        const int t = s_w[q];
        if (q < warp) woff += t;
        total += t;
    }
    return make_int2(woff + incl, total);
}

template <int NT, int KIND>
__global__ void __launch_bounds__(NT, (NT == 256 ? 4 : (NT == 128 ? 6 : 8))) trace_hits_kernel(const TraceArgs a) {
    constexpr int REC = RecFloats<KIND>::v;
    constexpr int NW = NT / 32;
    constexpr int kDead = 1 << 12;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_rec_all = reinterpret_cast<float*>(smem_raw);                                  // [2][NT][REC]
    int2* s_meta_all = reinterpret_cast<int2*>(s_rec_all + 2 * (size_t)NT * REC);          // [2][NT+2] (first item, packed rect)
    unsigned* s_bmask_all = reinterpret_cast<unsigned*>(s_meta_all + 2 * (NT + 2));        // [2][kMaxBlocks] first-item bits
    float* s_ray = reinterpret_cast<float*>(s_bmask_all + 2 * kMaxBlocks);                 // [3][NT]
    int* s_cnt = reinterpret_cast<int*>(s_ray + 3 * NT);                                   // [NT] next free slot of the pixel's segment (tile-relative)
    int* s_base = s_cnt + NT;                                                              // [NT] segment start within the tile area
    int* s_wsum = s_base + NT;                                                             // [32] scan scratch (two halves)
    unsigned short* s_first_all = reinterpret_cast<unsigned short*>(s_wsum + 32);          // [2][kMaxBlocks] owner of a block's first item
    int* s_diff = reinterpret_cast<int*>(smem_raw);                                        // [17*17] + [16*16], aliases the records (pre-pass only)

    const int tid = threadIdx.x;
    int blk_id = blockIdx.x;
    const int tx = blk_id % a.TX; blk_id /= a.TX;
    const int ty = blk_id % a.TY;
    const int b = blk_id / a.TY;
    const int64_t tile_id = ((int64_t)b * a.TY + ty) * a.TX + tx;
    const int tile = a.tile;

    const int lx = tid % tile, ly = tid / tile;        // pixel column of this thread: col = ly*tile + lx = tid
    const int xi = tx * tile + lx, yi = ty * tile + ly;
    const bool live = tid < tile * tile && xi < a.W && yi < a.H;
    const float c0 = a.origins[3 * b], c1 = a.origins[3 * b + 1], c2 = a.origins[3 * b + 2];
    {
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
        if (live) pixel_ray(a.rays, a.cam, b, xi, yi, a.H, a.W, r0, r1, r2);
        s_ray[tid] = r0; s_ray[NT + tid] = r1; s_ray[2 * NT + tid] = r2;
    }
    for (int i = tid; i < 17 * 17; i += NT) s_diff[i] = 0;

    const int64_t beg = a.tile_offsets[tile_id * kBinSub];
    const int64_t end = a.tile_offsets[(tile_id + 1) * kBinSub];
    const int n = (int)(end - beg);
    const int64_t item_base = a.item_base >= 0 ? a.item_base : a.tile_item_offsets[0];
    if ((a.list_capacity > 0 && end > a.list_capacity) ||
        (a.hits_capacity > 0 && a.tile_item_offsets[(tile_id + 1) * kBinSub] - item_base > a.hits_capacity)) {
        // the caller's speculative scratch is too small for this tile: nothing is read or written outside it, the
        // host sees the true totals and repeats the call with exact sizes (voge_b200/_C.py: BinPlan)
        a.counts[tile_id * NT + tid] = 0;
        a.seg_base[tile_id * NT + tid] = 0;
        return;
    }
    const uint4* list = a.tile_list + 2 * beg;
    const int warp = tid >> 5, lane = tid & 31;
    const int px0 = tx * tile, py0 = ty * tile;
    const int pxe = min(px0 + tile, a.W) - 1, pye = min(py0 + tile, a.H) - 1;
    __syncthreads();

    // ---- pre-pass: how many rectangles cover each pixel (2D difference array + prefix sums) ----
    for (int i = tid; i < n; i += NT) {
        const uint4 ent = __ldg(list + 2 * i);
        const uint2 rc = make_uint2(ent.y, ent.z);
        const int xl = max((int)(rc.x & 0xffffu), px0) - px0, xh = min((int)(rc.x >> 16), pxe) - px0;
        const int yl = max((int)(rc.y & 0xffffu), py0) - py0, yh = min((int)(rc.y >> 16), pye) - py0;
        if (xl <= xh && yl <= yh) {
            atomicAdd(&s_diff[yl * 17 + xl], 1);
            atomicAdd(&s_diff[yl * 17 + xh + 1], -1);
            atomicAdd(&s_diff[(yh + 1) * 17 + xl], -1);
            atomicAdd(&s_diff[(yh + 1) * 17 + xh + 1], 1);
        }
    }
    __syncthreads();
    int* s_rowp = s_diff + 17 * 17;
    if (tid < tile * tile) {
        int v = 0;
        for (int x = 0; x <= lx; ++x) v += s_diff[ly * 17 + x];
        s_rowp[ly * 16 + lx] = v;
    }
    __syncthreads();
    int cover = 0;
    if (tid < tile * tile)
        for (int y = 0; y <= ly; ++y) cover += s_rowp[y * 16 + lx];
    // Segments start on 32-byte boundaries (4 slots): a sector of `hits` then belongs to ONE pixel, so the selection reads
    // ceil(hits / 4) sectors per pixel instead of one more on average, and the trace's partial-sector writes shrink alike.
    // The rounding needs up to 3 slots per pixel + 3 per tile, which the caller adds to the tile's item count before
    // the scan (voge_bin_item_slack); a tile whose area has no such room keeps the plain layout.
    const int64_t tile_off = a.tile_item_offsets[tile_id * kBinSub] - item_base;
    int64_t tile_base = (tile_off + (kSegAlign - 1)) & ~(int64_t)(kSegAlign - 1);
    {
        int cov = (cover + (kSegAlign - 1)) & ~(kSegAlign - 1);
        int2 sc = block_scan<NT>(cov, s_wsum, lane, warp);
        if (tile_base + sc.y > a.tile_item_offsets[(tile_id + 1) * kBinSub] - item_base) {
            __syncthreads();             // the scan scratch is read by every thread before it is rewritten
            tile_base = tile_off;
            cov = cover;
            sc = block_scan<NT>(cov, s_wsum, lane, warp);
        }
        s_base[tid] = sc.x - cov;
        s_cnt[tid] = sc.x - cov;         // the slot counter starts at the segment base: one atomic yields the slot
        a.seg_base[tile_id * NT + tid] = tile_base + (sc.x - cov);
    }
    __syncthreads();   // the records alias s_diff / s_rowp
    unsigned n_eval = 0;

    // the list entries of a chunk are requested before the item loop of the previous one (coalesced 32-byte loads):
    // a chunk's set-up then starts with its candidates at hand
    uint4 ent, hd;
    ent = make_uint4(0u, 1u, 0u, 0u); hd = make_uint4(0u, 0u, 0u, 0u);
    if (tid < n) ldg256u(list + 2 * tid, ent, hd);
    for (int base = 0, buf = 0; base < n; buf ^= 1) {
        float* s_rec = s_rec_all + (size_t)buf * NT * REC;
        int2* s_meta = s_meta_all + buf * (NT + 2);
        unsigned* s_bmask = s_bmask_all + buf * kMaxBlocks;
        unsigned short* s_first = s_first_all + buf * kMaxBlocks;
        // ---- setup: one candidate per thread ----
        int area = 0, ng = 0;          // items, item groups
        int pack = kDead;
        if (base + tid < n) {
            const int g = (int)ent.x;
            const uint2 rc = make_uint2(ent.y, ent.z);
            const float v0 = __uint_as_float(hd.x), v1 = __uint_as_float(hd.y), v2 = __uint_as_float(hd.z);
            float S[9];
            finish_gauss<KIND>(a.gauss, g, make_float4(v0, v1, v2, __uint_as_float(hd.w)), S, a.enc != 0);
            const float m0 = __fsub_rn(v0, c0), m1 = __fsub_rn(v1, c1), m2 = __fsub_rn(v2, c2);   // verts - ray_origin, Renderer.py:130
            const int xl = max((int)(rc.x & 0xffffu), px0), xh = min((int)(rc.x >> 16), pxe);
            const int yl = max((int)(rc.y & 0xffffu), py0), yh = min((int)(rc.y >> 16), pye);
            const int w = xh - xl + 1, h = yh - yl + 1;
            // every list entry intersects its tile by construction; an empty intersection still gets one
            // (dead) item so that the item boundaries of a chunk stay distinct
            area = 1; ng = 1;
            if (w > 0 && h > 0) {
                area = w * h;
                ng = (area + kGroup - 1) / kGroup;
                // first column, width - 1, floor(p / w) multiplier
                pack = ((yl - py0) * tile + (xl - px0)) | ((w - 1) << 8) | ((int)kInvW[w] << 13);
            }
            float4* rp = reinterpret_cast<float4*>(s_rec + (size_t)tid * REC);
            if (KIND == 9) {
                // ray-independent part of exact_pair: t_ij = rn(mu_i S_ij) and msm (the reference binary
                // shares these products between msk and msm)
                const Prod9 pm = exact_row_products(m0, m1, m2, S);
                const float msm = exact_contract(pm, m0, m1, m2);
                rp[0] = make_float4(S[0], S[1], S[2], S[3]);
                rp[1] = make_float4(S[4], S[5], S[6], S[7]);
                rp[2] = make_float4(S[8], pm.t[0], pm.t[1], pm.t[2]);
                rp[3] = make_float4(pm.t[3], pm.t[4], pm.t[5], pm.t[6]);
                rp[4] = make_float4(pm.t[7], pm.t[8], msm, __int_as_float(g));
            } else {
                const float t0 = __fmul_rn(m0, S[0]), t1 = __fmul_rn(m1, S[4]), t2 = __fmul_rn(m2, S[8]);
                const float msm = __fmaf_rn(t2, m2, __fmaf_rn(t1, m1, __fmul_rn(t0, m0)));
                rp[0] = make_float4(S[0], S[4], S[8], t0);
                rp[1] = make_float4(t1, t2, msm, __int_as_float(g));
            }
        }
        for (int i = tid; i < kMaxBlocks; i += NT) s_bmask[i] = 0u;
        // block scan of the areas -> first item of every candidate.  Buffer `buf` and this half of the scan
        // scratch were last read two chunks ago: every thread has left that chunk's item loop before any
        // thread passed the barriers of the previous chunk.
        const int2 sc = block_scan<NT>(ng, s_wsum + buf * 16, lane, warp);
        const int excl = sc.x - ng;
        // the chunk takes the leading candidates whose groups fit the budget (a single rectangle always does)
        const bool take = ng > 0 && sc.x <= kGroupBudget;
        s_meta[tid] = make_int2(excl | ((area - 1) << 16), pack);      // first group | items - 1
        if (tid == NT - 1) s_meta[NT] = make_int2(sc.y, kDead);
        if (take) {
            atomicOr(&s_bmask[excl >> 5], 1u << (excl & 31));
            const int b_hi = (excl + ng - 1) >> 5;
            for (int bb = (excl + 31) >> 5; bb <= b_hi; ++bb) s_first[bb] = (unsigned short)tid;
        }
        const int taken = __syncthreads_count(take);
        const int total = s_meta[taken].x & 0xffff; // first group of the first candidate left out = groups taken
        {
            const int nb = base + max(taken, 1) + tid;      // this thread's candidate of the next chunk
            if (nb < n) ldg256u(list + 2 * nb, ent, hd);
        }

        // ---- items: blocks of 32 consecutive groups, round-robin over the warps; a lane keeps the record of
        // its group's candidate in registers for the kGroup pixels ----
        const int nblk = (total + 31) >> 5;
        for (int blk = warp; blk < nblk; blk += NW) {
            const int gid = (blk << 5) + lane;
            // owner = owner of the block's first group + number of candidates starting in (first, gid]
            const int owner = s_first[blk] + __popc(s_bmask[blk] & ((2u << lane) - 2u));
            const int2 mt = s_meta[owner];
            if (gid >= total || (mt.y & kDead)) continue;
            const int p0 = (gid - (mt.x & 0xffff)) * kGroup;
            const int n_it = min(kGroup, (mt.x >> 16) + 1 - p0);
            const int w = ((mt.y >> 8) & 15) + 1;
            int yy = (int)(((unsigned)p0 * ((unsigned)mt.y >> 13)) >> 16);
            int xx = p0 - yy * w;
            int col = (mt.y & 255) + yy * tile + xx;
            const float4* rp = reinterpret_cast<const float4*>(s_rec + (size_t)owner * REC);
            float S[9], t[9], msm;
            int gg;
            if (KIND == 9) {
                const float4 q0 = rp[0], q1 = rp[1], q2 = rp[2], q3 = rp[3], q4 = rp[4];
                S[0] = q0.x; S[1] = q0.y; S[2] = q0.z; S[3] = q0.w; S[4] = q1.x; S[5] = q1.y; S[6] = q1.z; S[7] = q1.w; S[8] = q2.x;
                t[0] = q2.y; t[1] = q2.z; t[2] = q2.w; t[3] = q3.x; t[4] = q3.y; t[5] = q3.z; t[6] = q3.w; t[7] = q4.x; t[8] = q4.y;
                msm = q4.z; gg = __float_as_int(q4.w);
            } else {
                const float4 q0 = rp[0], q1 = rp[1];
                S[0] = q0.x; S[4] = q0.y; S[8] = q0.z; t[0] = q0.w; t[4] = q1.x; t[8] = q1.y;
                msm = q1.z; gg = __float_as_int(q1.w);
            }
            n_eval += n_it;
            // phase 1: the kGroup pairs are evaluated back to back (independent FMA chains in flight together);
            // slots past the end of the rectangle repeat its last pixel and are masked out.  Both IEEE quotients of a
            // pair -- msk^2 / ksk for act and msk / ksk = len (reference :190-193) -- are formed here, unconditionally:
            // they share the refined reciprocal of ksk, and the hit phase below is then free of divergent arithmetic
            unsigned lenb[kGroup];
            int cols[kGroup];
            unsigned hit = 0u;
#pragma unroll
            for (int j = 0; j < kGroup; ++j) {
                cols[j] = col;
                const float d0 = s_ray[col], d1 = s_ray[NT + col], d2 = s_ray[2 * NT + col];
                float ksk, msk;
                if (KIND == 9) {
                    Prod9 pm;
#pragma unroll
                    for (int q = 0; q < 9; ++q) pm.t[q] = t[q];
                    const Prod9 pd = exact_row_products(d0, d1, d2, S);
                    ksk = exact_contract(pd, d0, d1, d2);
                    msk = exact_contract(pm, d0, d1, d2);
                } else {
                    const float u0 = __fmul_rn(d0, S[0]), u1 = __fmul_rn(d1, S[4]), u2 = __fmul_rn(d2, S[8]);
                    ksk = __fmaf_rn(u2, d2, __fmaf_rn(u1, d1, __fmul_rn(u0, d0)));
                    msk = __fmaf_rn(t[8], d2, __fmaf_rn(t[4], d1, __fmul_rn(t[0], d0)));
                }
                const float act = __fsub_rn(msm, __fdiv_rn(__fmul_rn(msk, msk), ksk));
                const float len = __fdiv_rn(msk, ksk);
                lenb[j] = orderable(len);
                // reference :197: a hit enters the list only if len < 1e10 (the initial slot value); NaN never does
                if (j < n_it && act < a.thr_act && len < kEmptyLen) hit |= 1u << j;
                // next pixel of the rectangle (row-major)
                if (j + 1 < n_it) {
                    ++col;
                    if (++xx == w) { xx = 0; col += tile - w; }
                }
            }
            // phase 2: the hits
#pragma unroll
            for (int j = 0; j < kGroup; ++j) {
                if (hit & (1u << j)) {
                    const int slot = atomicAdd(&s_cnt[cols[j]], 1);
                    a.hits[tile_base + slot] = make_uint2(lenb[j], (unsigned)gg);
                }
            }
        }
        base += max(taken, 1);
    }
    __syncthreads();
    a.counts[tile_id * NT + tid] = s_cnt[tid] - s_base[tid];
    if (a.stats != nullptr) {
        unsigned long long e = n_eval;
        for (int s = 16; s > 0; s >>= 1) e += __shfl_down_sync(0xffffffffu, e, s);
        if (lane == 0) atomicAdd(a.stats, e);
    }
}

template <int NT, int KIND>
static int launch_trace(const TraceArgs& a, cudaStream_t stream) {
    const size_t smem = trace_smem_bytes(NT, RecFloats<KIND>::v);
    const long long grid = (long long)a.B * a.TX * a.TY;
    if (grid <= 0 || grid > 2147483647LL) return (int)cudaErrorInvalidValue;
    VOGE_CUDA_TRY(cudaFuncSetAttribute(trace_hits_kernel<NT, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    trace_hits_kernel<NT, KIND><<<(unsigned)grid, NT, smem, stream>>>(a);
    VOGE_LAUNCH_CHECK();
    return 0;
}

template <int KIND>
static int dispatch_trace(const TraceArgs& a, cudaStream_t s) {
    const int nt = tile_threads(a.tile);
    if (a.tile < 1 || a.tile > 16) return (int)cudaErrorInvalidValue;
    if (nt == 256) return launch_trace<256, KIND>(a, s);
    if (nt == 128) return launch_trace<128, KIND>(a, s);
    return launch_trace<64, KIND>(a, s);
}

}  // namespace voge

extern "C" int voge_trace_threads(int tile) { return voge::tile_threads(tile); }

// 3 slots per pixel of the largest tile (256) + 3 for the tile's own alignment, spread over the kBinSub item counters
extern "C" int voge_bin_item_slack(void) { return ((voge::kSegAlign - 1) * 257 + voge::kBinSub - 1) / voge::kBinSub; }

extern "C" int voge_trace_hits(const float* gauss, int sigma_kind, const float* origins,
                               const float* rays, const float* cam, const int64_t* tile_offsets, const int32_t* tile_list,
                               const uint32_t* rects, const int64_t* tile_item_offsets, int64_t item_base, float thr_act, int B, int N,
                               int H, int W, int tile, int32_t* counts, int64_t* seg_base, uint32_t* hits,
                               int64_t list_capacity, int64_t hits_capacity, uint64_t* stats, voge_stream_t stream) {
    using namespace voge;
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    TraceArgs a;
    if (rays == nullptr && cam == nullptr) return (int)cudaErrorInvalidValue;
    a.gauss = gauss; a.origins = origins; a.rays = rays; a.cam = cam; a.tile_offsets = tile_offsets;
    a.tile_list = reinterpret_cast<const uint4*>(tile_list); a.rects = reinterpret_cast<const uint2*>(rects); a.tile_item_offsets = tile_item_offsets;
    a.item_base = item_base; a.list_capacity = list_capacity; a.hits_capacity = hits_capacity;
    a.thr_act = thr_act; a.B = B; a.N = N; a.H = H; a.W = W; a.tile = tile; a.TX = cdiv(W, tile); a.TY = cdiv(H, tile);
    a.counts = counts; a.seg_base = seg_base; a.hits = reinterpret_cast<uint2*>(hits);
    a.stats = reinterpret_cast<unsigned long long*>(stats);
    a.enc = (sigma_kind & kKindIsoEncoded) ? 1 : 0;
    sigma_kind &= ~kKindIsoEncoded;
    if (a.enc && sigma_kind != 9) return (int)cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)stream;
    if (sigma_kind == 1) return dispatch_trace<1>(a, s);
    if (sigma_kind == 3) return dispatch_trace<3>(a, s);
    if (sigma_kind == 9) return dispatch_trace<9>(a, s);
    return (int)cudaErrorInvalidValue;
}

/*
 * voge_b200.h -- C ABI of libvoge_b200.so, the B200 (sm_100a) implementation of the
 * VoGE ray-tracing hot path.
 *
 * Drop-in boundary: these entry points are what the reference's pybind11 module
 * `VoGE._C` (reference VoGE/csrc/ext.cpp:7-17) binds for this path, restated as plain
 * `extern "C"` functions over raw DEVICE pointers, sizes and a CUDA stream -- no torch
 * types.  The caller owns every buffer (the Python host allocates them with torch so the
 * caching allocator owns all memory); kernels never allocate or free.  All functions are
 * stateless and re-entrant, launch asynchronously on `stream` and return 0 on success or
 * a cudaError_t value (use voge_error_string()).  float = IEEE fp32, indices = int32.
 *
 * Conventions:  N Gaussians, B views, P = B*N packed Gaussians,
 * H x W image, K = max_assign, BH x BW coarse bins of bin_size pixels, M = max points
 * per bin.  "isigmas" is the 3x3 matrix S (= 2 * inverse covariance) row-major.
 */
#ifndef VOGE_B200_H
#define VOGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* voge_stream_t; /* cudaStream_t */

/* ---- library info ------------------------------------------------------------------ */
int voge_version(void);                       /* ABI version, currently 1               */
const char* voge_error_string(int code);      /* cudaGetErrorString for a returned code */
int voge_device_sm_count(int* sm_count);      /* SMs of the current device              */
/* Peak micro-benchmarks for the roofline denominators MEASURED_PEAKS.json does not hold: `blocks`
 * CTAs of 256 threads, `iters` rounds of 16 independent FFMA (32 flops) resp. 8 MUFU.EX2 per
 * thread.  The caller times the launch with CUDA events.  `out` = any 4-byte device buffer.  */
int voge_peak_fp32(int blocks, int iters, float* out, voge_stream_t stream);
int voge_peak_sfu(int blocks, int iters, float* out, voge_stream_t stream);

/* ---- coarse binning ----------------------------------------------------------------
 * Replaces `rasterize_points_coarse` (ext.cpp:8 -> RasterizeEllipseCoarseCuda,
 * rasterize_coarse.cu:254-305; kernels :20-42 and :44-188).
 * points_ndc (P,3): x,y in VoGE's flipped NDC, z = view-space depth (skip if z < 0);
 * radius (P,2): bbox half extents.  first_idx/num_per (B) int64 DEVICE arrays;
 * max_per_cloud >= max(num_per) (host value, sizes the grid).
 * bin_points (B,BH,BW,M) int32 must be PRE-FILLED with -1 by the caller (the reference's
 * at::full, rasterize_coarse.cu:222); indices written are PACKED and, unlike the reference, in
 * deterministic ascending order.  bin_counts (B,BH,BW) int32 receives the TRUE number of
 * overlapping Gaussians per bin; a count > M means overflow (the reference prints from the
 * device and drops chunks, :154-170): here the first M indices are kept and the host raises.
 * scratch: int32 device buffer of voge_rasterize_coarse_scratch_elems(...) elements.
 * Returns cudaErrorInvalidValue if BH or BW > 65 (the reference refuses >= 66, :213).          */
int64_t voge_rasterize_coarse_scratch_elems(int B, int max_per_cloud, int H, int W, int bin_size);
int voge_rasterize_coarse(const float* points_ndc, const float* radius,
                          const int64_t* first_idx, const int64_t* num_per,
                          int B, int P, int max_per_cloud, int H, int W, int bin_size, int M,
                          int32_t* scratch, int32_t* bin_points, int32_t* bin_counts,
                          voge_stream_t stream);

/* ---- fine ray tracing ----------------------------------------------------------------
 * Replaces `ray_trace_voge_fine` (ext.cpp:9 -> RayTraceFineVoge, ray_trace_voge.cu:219-280,
 * kernel :135-217).  Outputs (B,H,W,K) are fully written by the callee-side kernels
 * (idx init -1, len/act init 1e10, dsd init 0).  Arithmetic reproduces the reference's
 * fp32 rounding sequence bit for bit (DESIGN.md "rounding contract").  The batch stride of
 * bin_points is BH*BW*M (the reference's b*BH*BH*M, :185, is a defect for non-square grids).
 * The _counts variant takes the optional bin_counts (B,BH,BW) of voge_rasterize_coarse so
 * only the first min(count, M) entries of each list are scanned (NULL = scan all M).       */
int voge_ray_trace_fine(const float* mus, const float* isigmas, const float* rays,
                        const int32_t* bin_points, float thr_act, int bin_size,
                        int B, int H, int W, int BH, int BW, int M, int K, int P,
                        int32_t* out_idx, float* out_len, float* out_act, float* out_dsd,
                        voge_stream_t stream);
int voge_ray_trace_fine_counts(const float* mus, const float* isigmas, const float* rays,
                               const int32_t* bin_points, const int32_t* bin_counts,
                               float thr_act, int bin_size,
                               int B, int H, int W, int BH, int BW, int M, int K, int P,
                               int32_t* out_idx, float* out_len, float* out_act, float* out_dsd,
                               voge_stream_t stream);

/* Replaces `ray_trace_voge_fine_backward` (ext.cpp:10 -> RayTraceFineVogeBackward,
 * ray_trace_voge.cu:334-379, kernel :283-332).  grad_mus (P,3) and grad_isg (P,3,3) must be
 * ZEROED by the caller (they are accumulated into); grad_rays (B,H,W,3) is written in full
 * and may be NULL when the ray gradient is not needed.                                     */
int voge_ray_trace_fine_backward(const float* mus, const float* isigmas, const float* rays,
                                 const int32_t* idx, const float* grad_len,
                                 const float* grad_act, const float* grad_dsd,
                                 int B, int H, int W, int K, int P,
                                 float* grad_rays, float* grad_mus, float* grad_isg,
                                 voge_stream_t stream);

/* ---- blend weights (reference: pure PyTorch, VoGE/Aggregation.py:30-107) ---------------
 * weight[r,m] = exp(-absorptivity * sum_k exp(-act_k)(erf((len_m-len_k)sqrt(dsd_k+1e-10))+1)/2)
 *               * exp(-act_m) / exp(-0.5);   valid_num[r] = #(idx >= 0) as int64.
 * R = number of rays (B*H*W).                                                             */
int voge_aggregation(const int32_t* idx, const float* act, const float* len, const float* dsd,
                     float absorptivity, int64_t R, int K,
                     float* weight, int64_t* valid_num, voge_stream_t stream);

/* Analytic backward of voge_aggregation (the reference relies on autograd through
 * Aggregation.py:30-79).  Writes grad_act/grad_len/grad_dsd (R,K) in full.                */
int voge_aggregation_backward(const float* act, const float* len, const float* dsd,
                              const float* grad_weight, float absorptivity, int64_t R, int K,
                              float* grad_act, float* grad_len, float* grad_dsd,
                              voge_stream_t stream);

/* ---- gather-blend (merge_final, Aggregation.py:111-141; Renderer.py:153-176) -----------
 * out[r,:] = sum_{k < valid_num[r]} weight[r,k] * attr[idx[r,k] (or 0 when < 0), :].
 * If background != NULL (C floats, device) the composite of Renderer.py:162-171 is fused:
 *   sil = min(sum_k weight[r,k], 1);  mask = (mask_thr > 0) ? (sil > mask_thr) : sil;
 *   out = min(out + (1 - mask) * background, 1).
 * idx_mod > 0 maps packed indices (b*N+n) onto attr rows with idx % idx_mod; attr has n_attr
 * rows and indices >= n_attr are ignored (the reference asserts on the host, :120).
 * attr_padded4 != 0 (C <= 4 only): attr is an (n_attr,4) zero-padded table, fetched with one 16-byte
 * load per hit (every lane of a warp gathers a different row).
 * sat_code (optional, (R,) uint8, C <= 4 with a background): per channel c two bits at 2c telling where the
 * final min(x, 1) clamped -- 2: x < 1, 1: x == 1, 0: x > 1 -- i.e. twice the factor the backward applies.
 * idx_pad (optional, C <= 4; normally == idx): the slots k >= valid_num[r] of every row are overwritten with 0 --
 * the reference's in-place `vert_assign += (vert_assign < 0)` (:131) for fragments whose slots behind valid_num are
 * known to hold the -1 padding, without a read-modify-write pass over the (R,K) tensor.                          */
int voge_merge_final(const float* attr, const float* weight, const int32_t* idx,
                     const int64_t* valid_num, const float* background, float mask_thr,
                     int64_t R, int K, int C, int idx_mod, int n_attr, int attr_padded4,
                     float* out, uint8_t* sat_code, int32_t* idx_pad, voge_stream_t stream);

/* Backward of voge_merge_final: grad_attr must be ZEROED by the caller and is accumulated into;
 * its layout is (n_attr,C), or -- packed4 != 0 and C <= 4 -- (n_attr,4) zero-padded rows so that one
 * 16-byte vector reduction per hit can be used.  grad_weight (R,K) is written in full.
 * Either may be NULL.  out: optional (R,C) = the forward's output; with it the composite is re-evaluated
 * only for pixels whose output saturated at 1 (the min(x,1) subgradient needs x there).    */
int voge_merge_final_backward(const float* attr, const float* weight, const int32_t* idx,
                              const int64_t* valid_num, const float* background,
                              float mask_thr, const float* out, const float* grad_out,
                              int64_t R, int K, int C, int idx_mod, int n_attr, int packed4,
                              int attr_padded4, float* grad_attr, float* grad_weight, voge_stream_t stream);

/* ---- sampling (inverse rendering) --------------------------------------------------------
 * Replaces `sample_voge` (ext.cpp:14 -> SampleVoge, sample_voge.cu:95-134, kernel :35-66):
 * feat[n,:] += w * image[r,:], wsum[n] += w for every (r,k) with idx >= 0.
 * feat (num_vert,C) and wsum (num_vert) must be ZEROED by the caller.                     */
int voge_sample(const float* image, const float* weight, const int32_t* idx,
                int64_t R, int K, int C, int num_vert,
                float* feat, float* wsum, voge_stream_t stream);

/* Replaces `sample_voge_backward` (ext.cpp:15, sample_voge.cu:173-252): writes
 * grad_image (R,C) and grad_weight (R,K) in full (no pre-zeroing needed).                 */
int voge_sample_backward(const float* image, const float* weight, const int32_t* idx,
                         const float* grad_feat, const float* grad_wsum,
                         int64_t R, int K, int C,
                         float* grad_image, float* grad_weight, voge_stream_t stream);

/* Replaces `scatter_max` (ext.cpp:16, sample_voge.cu:69-92,137-170): wmax[n] = max(w).
 * wmax (num_vert) must be ZEROED by the caller (the reference starts from zeros).         */
int voge_scatter_max(const float* weight, const int32_t* idx, int64_t R, int K,
                     int num_vert, float* wmax, voge_stream_t stream);

/* ---- dense-ray API (next tier) --------------------------------------------------
 * Replace `ray_trace_voge_ray` (ext.cpp:11 -> RayTraceVogeRay, voge_ray_tracing_ray.cu:242-283),
 * `ray_trace_voge_ray_backward` (ext.cpp:12, :287-325) and `find_nearest_k` (ext.cpp:13, :328-375).
 * mus (M,3), isigmas (M,3,3), rays (N,3); outputs (N,M) resp. (N,K).  grad_mus / grad_isg ZEROED by
 * the caller; grad_rays written in full.  find_nearest_k pads with idx -1, len 1e10, act 0, dsd 0.   */
int voge_ray_trace_ray(const float* mus, const float* isigmas, const float* rays, int M, int N,
                       float* out_len, float* out_act, float* out_dsd, voge_stream_t stream);
int voge_ray_trace_ray_backward(const float* mus, const float* isigmas, const float* rays,
                                const float* grad_len, const float* grad_act, const float* grad_dsd,
                                int M, int N, float* grad_rays, float* grad_mus, float* grad_isg,
                                voge_stream_t stream);
int voge_find_nearest_k(const float* len_in, const float* act_in, const float* dsd_in, float thr_act,
                        int M, int K, int N, int32_t* out_idx, float* out_len, float* out_act,
                        float* out_dsd, voge_stream_t stream);

/* ---- converters (next-tier row f-4) ------------------------------------------------------------------------
 * Neighbour statistic of `naive_point_cloud_converter` (VoGE/Converter/Converters.py:106-111): avg_len[i] = mean over
 * the n_nearest (<= 16) smallest distances from point i to the cloud (its own zero distance included), each clipped at
 * thr_max x their mean.  points (N,3), avg_len (N,) out.                                                      */
int voge_knn_mean_dist(const float* points, int N, int n_nearest, float thr_max, float* avg_len,
                       voge_stream_t stream);

/* ---- fused renderer path (GaussianRenderer.forward, reference VoGE/Renderer.py:102-150) ------
 * Same results as rasterize_coarse -> ray_trace_voge_fine -> aggregation on the renderer's own
 * call pattern, without the per-view (B,N,.) copies, the (B,BH,BW,M) bin table, the (R,K,K)
 * blend tensors or (for the closed-form camera) the (B,H,W,3) rays.
 *
 * Parameters.  verts (N,3); sigmas compact: sigma_kind 1 = (N,), 3 = (N,3), 9 = (N,3,3);
 * sigma_mode says how `sigmas` maps to the matrix P of S = 2 P (Renderer.py:134-137):
 *   0  P = sigmas (inverse covariances, inverse_sigma = False)
 *   1  P = inverse(sigmas) (covariances, inverse_sigma = True: `2 * torch.inverse(sigmas)`)
 *   2  P = tril(sigmas) tril(sigmas)^T (Cholesky factor: `to_sym`, demo/EfficientCuboidViaOptimization.py:17-18;
 *      sigma_kind 9 only)
 * voge_pack_gaussians writes one aligned record per Gaussian, out (N, 4 | 8 | 16) floats for
 * sigma_kind 1 | 3 | 9 = [x,y,z,S00] | [x,y,z,S00, S11,S22,0,0] | [x,y,z,S00, attribute row (4), S01,S02,S10,S11, S12..S22];
 * every other kernel of the path reads these records (`gauss`).  The attribute row of a kind-9 record (64 bytes) is
 * written by voge_pack_attr (attr (N, C <= 4) -> zero-padded attr4 (N,4), optional, and / or the records' second 16
 * bytes, optional): voge_render_backward_image then fetches geometry + attribute of a hit with one 256-bit request
 * (flag bit 1 of its need_sigma argument).  voge_unpack_gradients is the matching gradient epilogue: it
 * splits the packed gradient records of voge_render_backward_fused into grad_verts (N,3) and grad_sigmas
 * (shaped like sigmas, NULL to skip) and applies the chain rule of sigma_mode (mode 1: -P^T G P^T; mode 2:
 * tril((G + G^T) tril(L))) -- replaces autograd through torch.inverse / to_sym.
 * Isotropic encoding (sigma_kind 9 only).  With iso_flag != NULL (one int32, ZEROED by the caller) voge_pack_gaussians
 * writes every record whose S is exactly s I, s > 0 (off-diagonal entries +0) with -s in the S00 slot: such a
 * Gaussian is then fully described by the first 16 bytes of its record and a hit gathers one sector instead of 48
 * bytes.  The kernels that read `gauss` decode it when their sigma_kind argument is 9 | VOGE_KIND_ISO_ENCODED and
 * rebuild the same nine floats, so results are bit-identical to plain records.  If any OTHER record has the sign
 * bit set in S00 (not positive definite) the sign would be ambiguous: the kernel sets *iso_flag = 1 and the caller
 * must pack again with iso_flag = NULL (plain records, sigma_kind 9).
 *
 * Cameras.  R (B,3,3) row-vector convention X_view = X_world R + T, T (B,3), focal (B,2), principal (B,2) in
 * pixels; origins (B,3) = ray origins.  Rays: either a (B,H,W,3) tensor of unit directions (user-supplied
 * rays, pytorch3d cameras), or rays = NULL and cam (B,16) = [R row-major (9), fx, fy, px, py, 0,0,0]: the
 * kernels then generate d = R normalize(-(x+.5-px)/fx, -(y+.5-py)/fy, 1) per pixel in registers
 * (Renderer.py:124-128; ONE device function, csrc/render_core.cuh: gen_ray).  voge_generate_rays
 * materialises exactly those rays, (B,H,W,3), for the op-by-op entry points and the oracle.
 *
 * voge_bin_count: per (view, Gaussian) tile rectangle = [reference coarse-bin test of
 *   RayTracing.py:33-57 + rasterize_coarse.cu:20-42,:116-130 at `bin_size` px, if use_ref_bins]
 *   AND [conservative projected-ellipsoid bound]; tiles are `tile` x `tile` px (tile <= 16, divides
 *   bin_size).  rects (B,N,2) uint32 out = conservative PIXEL rectangle x0|x1<<16, y0|y1<<16 (inclusive,
 *   empty if x0 > x1); tile_counters (B,TY,TX,S) uint64, S = voge_bin_sub() list segments per tile (entry n is
 *   counted in segment n % S: L2 serialises atomics on one address), must be ZEROED by the caller: the low word
 *   counts the segment's list entries, the high word accumulates their rectangle areas inside the tile (the tile's
 *   number of ITEMS, trace.cu) -- one 64-bit reduction per (entry, tile) serves both.
 *   flags: bit 0 = use the dense-S rounding constants for every Gaussian (default: count the non-zero entries
 *   of S, DESIGN.md "culling margins").
 * voge_bin_fill: scatters 32-byte entries (Gaussian index, rectangle x, rectangle y, 0 | first 16 bytes of the Gaussian's
 *   record) into tile_list (total, 8) int32 -- the trace then reads its candidates with coalesced loads and gathers only the
 *   rest of non-isotropic records;
 *   cursor (B*TY*TX*S) uint64 must hold the segments' offsets (exclusive scan of the entry counts; the S segments of
 *   a tile are adjacent) and is advanced: one atomic yields an entry's position.
 *
 * Speculative scratch (no host round trip between voge_bin_count and the rest of the forward): a caller that sizes
 *   tile_list and hits from a PREVIOUS call's totals passes those sizes as list_capacity (entries) / hits_capacity
 *   (slots); voge_bin_fill then drops entries past the capacity and voge_trace_hits skips (counts = 0) every tile
 *   whose list or segment area would cross one, so nothing is read or written out of bounds.  The caller reads
 *   the true totals (last elements of the two scans) asynchronously and must repeat the call with exact sizes
 *   when they exceed the capacities.  <= 0: the buffers have their exact sizes.  item_base < 0 in
 *   voge_trace_hits: the kernel reads it from tile_item_offsets[0] (the group's first tile).                */
#define VOGE_KIND_ISO_ENCODED 0x100
int voge_bin_sub(void);   /* counters / list segments per tile (S below) */
int voge_pack_gaussians(const float* verts, const float* sigmas, int sigma_kind, int sigma_mode, int N,
                        float* out, int32_t* iso_flag, voge_stream_t stream);
int voge_pack_attr(const float* attr, int C, int N, float* attr4, float* gauss16, voge_stream_t stream);
int voge_unpack_gradients(const float* grad_packed, const float* gauss, const float* sigmas, int sigma_kind,
                          int sigma_mode, int N, float* grad_verts, float* grad_sigmas, voge_stream_t stream);
int voge_generate_rays(const float* cam, int B, int H, int W, float* rays, voge_stream_t stream);
int voge_bin_count(const float* gauss, int sigma_kind, const float* R,
                   const float* T, const float* origins, const float* focal, const float* principal,
                   int B, int N, int H, int W, float thr, float thr_act, int use_ref_bins,
                   int bin_size, int tile, int flags, uint32_t* rects, uint64_t* tile_counters,
                   voge_stream_t stream);
int voge_bin_fill(const uint32_t* rects, const float* gauss, int sigma_kind, uint64_t* cursor, int B, int N,
                  int H, int W, int tile, int32_t* tile_list, int64_t list_capacity, voge_stream_t stream);
/* ---- forward pipeline of the fused renderer (csrc/trace.cu, csrc/select.cu) ----------------------------
 *   voge_trace_hits: every item (tile-list entry x pixel of its rectangle inside the tile) is evaluated with
 *       the reference's arithmetic (ray_trace_voge.cu:188-193); hits (act < thr_act, len < 1e10, :197) are
 *       appended as (orderable len bits, local Gaussian index) to the pixel's segment.  The segments of a
 *       tile start at tile_item_offsets[tile*S] (B*TY*TX*S+1, int64 = exclusive scan of voge_bin_count's
 *       tile_items) and hold one slot per rectangle covering the pixel; hits is
 *       (tile_item_offsets[last], 2) uint32 = (orderable len bits, index) pairs.  Out: counts / seg_base (B*TY*TX, NT) per pixel column
 *       (col = ly*tile + lx, NT = voge_trace_threads(tile)).
 *   voge_select_topk: per pixel the K smallest (len, idx), ascending (== the reference's insertion rule,
 *       ray_trace_voge.cu:197-213) -> out_idx (B,H,W,K) packed b*N+n / -1, out_valid (B,H,W) int64.
 *   voge_blend_weights: exact re-evaluation of the first valid[r] slots of idx, blend weights
 *       (Aggregation.py:30-107) -> out_weight, out_len (1e10 padded), optional out_act / out_dsd.
 *   A batch may be processed in groups of views to bound the scratch (8 bytes per item): pass the group's
 *   slices of rays / cam / origins / rects / offsets / outputs, view_base = index of its first view (packed indices
 *   are (view_base + b)*N + n) and item_base = tile_item_offsets value of its first tile (subtracted, so
 *   that `hits` only needs the group's items).
 *   stats optional 4 x uint64 ([0] items evaluated, [2] pixels selected with the exact 64-bit keys), zeroed
 *   by the caller.                                                                                         */
int voge_trace_threads(int tile);
/* Segment alignment: voge_trace_hits starts every pixel's segment on a 32-byte boundary when the tile's area has room for
 * the rounding (3 slots per pixel + 3 per tile).  A caller that wants it initialises the ITEM half (high word) of each of
 * the S counters of a tile with voge_bin_item_slack() before voge_bin_count (instead of 0); without that room the
 * plain layout is used -- both are correct, the aligned one moves ~18 % fewer sectors through trace and selection.  */
int voge_bin_item_slack(void);
int voge_trace_hits(const float* gauss, int sigma_kind, const float* origins,
                    const float* rays, const float* cam, const int64_t* tile_offsets, const int32_t* tile_list,
                    const uint32_t* rects, const int64_t* tile_item_offsets, int64_t item_base, float thr_act,
                    int B, int N, int H, int W, int tile, int32_t* counts, int64_t* seg_base, uint32_t* hits,
                    int64_t list_capacity, int64_t hits_capacity, uint64_t* stats, voge_stream_t stream);
int voge_select_topk(const int32_t* counts, const int64_t* seg_base, const uint32_t* hits,
                     int view_base, int B, int N, int H, int W, int K, int tile,
                     int32_t* out_idx, int64_t* out_valid, uint64_t* stats, voge_stream_t stream);
int voge_blend_weights(const float* gauss, int sigma_kind, const float* origins,
                       const float* rays, const float* cam, const int32_t* idx, const int64_t* valid,
                       float absorptivity, int view_base, int B, int N, int H, int W, int K, float* out_weight,
                       float* out_len, float* out_act, float* out_dsd, voge_stream_t stream);

/* Fused backward of the renderer: d(weight) (B,H,W,K) and optionally d(hit length) -> packed parameter
 * gradients.  Recomputes the hits from idx (first valid_num[r] slots; merge_final rewrites -1 -> 0 in place,
 * Aggregation.py:131) instead of reading saved act/dsd, differentiates the blend analytically
 * (Aggregation.py:30-79) and applies the chain rule of ray_trace_voge.cu:324-330 in ONE kernel.  grad_packed is
 * ONE buffer of per-Gaussian records [d verts(3) | d P] padded to float4 units so that 16-byte vector reductions
 * can be used: kind 1: (N,4) = [gx,gy,gz,gP]; kind 3: (N,8) = [gx,gy,gz,0,gP0,gP1,gP2,0];
 * kind 9: (N,12) = [gx,gy,gz,gP00..gP22].  ZEROED by the caller, accumulated into; voge_unpack_gradients
 * turns it into the caller's tensors.
 * weight: optional (B,H,W,K) = the forward's out_weight (an output the caller holds anyway); when given the
 * kernel skips re-evaluating the blend weights (NULL: recompute them from the hits).
 * Camera gradients (pose optimisation, reference grad_rays of ray_trace_voge.cu:283-332): grad_origins (B,3),
 * optional, ZEROED by the caller (= -sum of d/d(mu') over the view's hits, since mu' = verts - origin,
 * Renderer.py:130); with a rays tensor: grad_rays (B,H,W,3), optional, written in full; with generated rays
 * (rays = NULL): grad_cam (B,16), optional, ZEROED by the caller = d/d(cam record) [dR (9), dfx, dfy, dpx, dpy],
 * the ray generator's chain rule reduced per view inside the kernel.                                       */
int voge_render_backward_fused(const float* gauss, int sigma_kind,
                               const float* origins, const float* rays, const int32_t* idx,
                               const int64_t* valid_num, const float* grad_weight, const float* weight,
                               const float* grad_len_out, float absorptivity,
                               int B, int N, int H, int W, int K,
                               float* grad_packed, int need_sigma, float* grad_rays, float* grad_origins,
                               const float* cam, float* grad_cam, voge_stream_t stream);

/* The same backward with merge_final's backward folded in ("image mode", K <= 112): the upstream gradient arrives on
 * out = min(sum_k w_k attr[idx_k] + (1 - mask) background, 1) (Aggregation.py:111-141 + Renderer.py:153-176) instead
 * of on the weights.  grad_out / fwd_out (B,H,W,C), C <= 4; attr4 (N,4) = attribute rows padded to 16 bytes;
 * background (C) or NULL (plain interpolate_attr); mask_thr as voge_merge_final; weight = the forward's out_weight
 * (required); sat_code (B,H,W) optional = voge_merge_final's clamp code of the same forward (else the kernel rebuilds the
 * un-clamped composite of saturated pixels).  dL/dw_k is formed in registers (no (B,H,W,K) weight-gradient round trip through HBM) and dL/d(attr) is
 * reduced by the same kernel into grad_attr4 (N,4), optional, ZEROED by the caller.  need_sigma: bit 0 = sigma
 * gradients wanted, bit 1 = the attribute rows live in the kind-9 records (voge_pack_attr).  Everything else as
 * voge_render_backward_fused.                                                                               */
int voge_render_backward_image(const float* gauss, int sigma_kind, const float* origins, const float* rays,
                               const float* cam, const int32_t* idx, const int64_t* valid_num,
                               const float* weight, const float* grad_out, const float* fwd_out,
                               const uint8_t* sat_code, const float* attr4, const float* background, float mask_thr, int C,
                               float absorptivity, int B, int N, int H, int W, int K, float* grad_packed,
                               int need_sigma, float* grad_attr4, float* grad_rays, float* grad_origins,
                               float* grad_cam, voge_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VOGE_B200_H */

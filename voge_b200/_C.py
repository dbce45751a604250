"""`VoGE._C`-shaped module: the reference's pybind entry points (reference VoGE/csrc/ext.cpp:7-17)
re-exposed with the same names, positional arguments, return orders and dtypes, implemented by
the C-ABI library libvoge_b200.so.  Outputs are allocated here with torch (the reference's
at::full / at::zeros) so PyTorch's caching allocator owns all memory.

CPU tensors raise RuntimeError, as the reference's CUDAGuard does: there is no CPU path.
"""
import os

import torch

from . import _lib
from ._lib import check, f32c, i32c, lib, ptr, require_cuda, stream_of

# set by rasterize_points_coarse: true per-bin counts of the last call (device tensor); lets
# callers bound the candidate scan and detect overflow without re-deriving anything.
last_bin_counts = None


class BinOverflowError(RuntimeError):
    pass


# Overflow policy of rasterize_points_coarse.  The reference prints a device-side warning, drops the
# overflowing chunk (non-deterministically) and carries on (rasterize_coarse.cu:150-163).  Here a bin that
# receives more than max_points_per_bin Gaussians keeps the first M in ascending index order (deterministic),
# the true counts are returned / kept in `last_bin_counts`, and nothing is read back to the host: no sync on
# the hot path.  check_overflow=True (or VOGE_CHECK_OVERFLOW=1) opts into a blocking check that raises.
CHECK_OVERFLOW_DEFAULT = os.environ.get("VOGE_CHECK_OVERFLOW", "0") == "1"


def rasterize_points_coarse(points, cloud_to_packed_first_idx, num_points_per_cloud, image_size, radius,
                            bin_size, max_points_per_bin, return_counts=False, check_overflow=None):
    """Reference: RasterizeEllipseCoarseCuda, rasterize_coarse.cu:254-305.
    points (P,3) f32, first_idx (B,) i64, num_per (B,) i64, image_size (H,W), radius (P,2) f32
    -> bin_points (B,BH,BW,M) int32, -1 padded, packed indices (ascending inside each bin)."""
    if check_overflow is None:
        check_overflow = CHECK_OVERFLOW_DEFAULT
    global last_bin_counts
    require_cuda(points, cloud_to_packed_first_idx, num_points_per_cloud, radius)
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")  # rasterize_coarse.cu:262-264
    points = f32c(points)
    radius = f32c(radius)
    first = cloud_to_packed_first_idx.to(torch.int64).contiguous()
    nper = num_points_per_cloud.to(torch.int64).contiguous()
    H, W = int(image_size[0]), int(image_size[1])
    B, P, M = int(nper.shape[0]), int(points.shape[0]), int(max_points_per_bin)
    BH, BW = 1 + (H - 1) // bin_size, 1 + (W - 1) // bin_size
    if BH >= 66 or BW >= 66:  # kMaxItemsPerBin, rasterize_coarse.cu:213-219
        raise RuntimeError("In RasterizeCoarseCuda got num_bins_y: %d, num_bins_x: %d, ; that's too many!" % (BH, BW))
    with torch.cuda.device(points.device):
        bin_points = torch.full((B, BH, BW, M), -1, dtype=torch.int32, device=points.device)
        counts = torch.zeros((B, BH, BW), dtype=torch.int32, device=points.device)
        if bin_points.numel() == 0 and M != 0:
            return (bin_points, counts) if return_counts else bin_points
        max_per = P  # upper bound of max(num_per) without a device sync
        n_scratch = lib().voge_rasterize_coarse_scratch_elems(B, max_per, H, W, bin_size)
        scratch = torch.empty((max(int(n_scratch), 1),), dtype=torch.int32, device=points.device)
        check(lib().voge_rasterize_coarse(ptr(points), ptr(radius), ptr(first), ptr(nper), B, P, max_per, H, W,
                                          int(bin_size), M, ptr(scratch), ptr(bin_points), ptr(counts),
                                          stream_of(points)), "rasterize_coarse")
    last_bin_counts = counts
    if check_overflow and M > 0:
        worst = int(counts.max().item())
        if worst > M:
            raise BinOverflowError(
                "Bin size was too small in the coarse rasterization phase: a bin holds %d Gaussians but "
                "max_points_per_bin is %d. Increase max_point_per_bin or set it to -1." % (worst, M))
    return (bin_points, counts) if return_counts else bin_points


def ray_trace_voge_fine(mus, isigmas, rays, bin_points, thr_act, bin_size, K, bin_counts=None):
    """Reference: RayTraceFineVoge, ray_trace_voge.cu:219-280.
    -> (point_idxs i32, total_len, total_act, total_dsd), each (B,H,W,K)."""
    require_cuda(mus, isigmas, rays, bin_points)
    mus, isigmas, rays = f32c(mus), f32c(isigmas), f32c(rays)
    bin_points = i32c(bin_points)
    B, H, W = int(rays.shape[0]), int(rays.shape[1]), int(rays.shape[2])
    BH, BW, M = int(bin_points.shape[1]), int(bin_points.shape[2]), int(bin_points.shape[3])
    P, K = int(isigmas.shape[0]), int(K)
    dev = mus.device
    with torch.cuda.device(dev):
        idx = torch.empty((B, H, W, K), dtype=torch.int32, device=dev)
        tlen = torch.empty((B, H, W, K), dtype=torch.float32, device=dev)
        tact = torch.empty((B, H, W, K), dtype=torch.float32, device=dev)
        tdsd = torch.empty((B, H, W, K), dtype=torch.float32, device=dev)
        if tlen.numel() == 0:
            return idx, tlen, tact, tdsd
        if bin_counts is not None:
            bin_counts = i32c(bin_counts)
        check(lib().voge_ray_trace_fine_counts(ptr(mus), ptr(isigmas), ptr(rays), ptr(bin_points), ptr(bin_counts),
                                               float(thr_act), int(bin_size), B, H, W, BH, BW, M, K, P,
                                               ptr(idx), ptr(tlen), ptr(tact), ptr(tdsd), stream_of(mus)),
              "ray_trace_fine")
    return idx, tlen, tact, tdsd


def ray_trace_voge_fine_dense(mus, isigmas, rays, points_per_view, thr_act, bin_size, K):
    """"No coarse stage" variant (reference RayTracing.py:22-26 builds an arange bin table instead):
    all `points_per_view` Gaussians of view b (packed b*N .. b*N+N-1) are candidates of every pixel."""
    require_cuda(mus, isigmas, rays)
    mus, isigmas, rays = f32c(mus), f32c(isigmas), f32c(rays)
    B, H, W = int(rays.shape[0]), int(rays.shape[1]), int(rays.shape[2])
    P, K, N = int(isigmas.shape[0]), int(K), int(points_per_view)
    BH, BW = 1 + (H - 1) // bin_size, 1 + (W - 1) // bin_size
    dev = mus.device
    with torch.cuda.device(dev):
        idx = torch.empty((B, H, W, K), dtype=torch.int32, device=dev)
        tlen = torch.empty((B, H, W, K), dtype=torch.float32, device=dev)
        tact = torch.empty((B, H, W, K), dtype=torch.float32, device=dev)
        tdsd = torch.empty((B, H, W, K), dtype=torch.float32, device=dev)
        if tlen.numel() == 0:
            return idx, tlen, tact, tdsd
        check(lib().voge_ray_trace_fine_counts(ptr(mus), ptr(isigmas), ptr(rays), None, None, float(thr_act),
                                               int(bin_size), B, H, W, BH, BW, N, K, P, ptr(idx), ptr(tlen),
                                               ptr(tact), ptr(tdsd), stream_of(mus)), "ray_trace_fine(dense)")
    return idx, tlen, tact, tdsd


def ray_trace_voge_fine_backward(mus, isigmas, rays, point_idxs, grad_len, grad_act, grad_dsd, need_rays=True):
    """Reference: RayTraceFineVogeBackward, ray_trace_voge.cu:334-379.
    -> (grad_ray (B,H,W,3), grad_mus (P,3), grad_isg (P,3,3))."""
    require_cuda(mus, isigmas, rays, point_idxs, grad_len, grad_act, grad_dsd)
    mus, isigmas, rays = f32c(mus), f32c(isigmas), f32c(rays)
    point_idxs = i32c(point_idxs)
    grad_len, grad_act, grad_dsd = f32c(grad_len), f32c(grad_act), f32c(grad_dsd)
    B, H, W, K = (int(s) for s in point_idxs.shape)
    P = int(isigmas.shape[0])
    dev = mus.device
    with torch.cuda.device(dev):
        g_ray = torch.zeros((B, H, W, 3), dtype=torch.float32, device=dev) if need_rays else None
        g_mus = torch.zeros((P, 3), dtype=torch.float32, device=dev)
        g_isg = torch.zeros((P, 3, 3), dtype=torch.float32, device=dev)
        check(lib().voge_ray_trace_fine_backward(ptr(mus), ptr(isigmas), ptr(rays), ptr(point_idxs), ptr(grad_len),
                                                 ptr(grad_act), ptr(grad_dsd), B, H, W, K, P, ptr(g_ray),
                                                 ptr(g_mus), ptr(g_isg), stream_of(mus)), "ray_trace_fine_backward")
    return g_ray, g_mus, g_isg


def sample_voge(image, vert_weight, vert_index, num_vert):
    """Reference: SampleVoge, sample_voge.cu:95-134. -> (vert_feature (N,C), vert_weight_sum (N,))."""
    require_cuda(image, vert_weight, vert_index)
    image, vert_weight, vert_index = f32c(image), f32c(vert_weight), i32c(vert_index)
    C, K, N = int(image.shape[-1]), int(vert_index.shape[-1]), int(num_vert)
    R = vert_index.numel() // max(K, 1)
    dev = image.device
    with torch.cuda.device(dev):
        feat = torch.zeros((N, C), dtype=torch.float32, device=dev)
        wsum = torch.zeros((N,), dtype=torch.float32, device=dev)
        check(lib().voge_sample(ptr(image), ptr(vert_weight), ptr(vert_index), R, K, C, N, ptr(feat), ptr(wsum),
                                stream_of(image)), "sample")
    return feat, wsum


def sample_voge_backward(image, vert_weight, vert_index, grad_feature, grad_weight_sum):
    """Reference: SampleVogeBackward, sample_voge.cu:212-252. -> (grad_image, grad_vert_weight)."""
    require_cuda(image, vert_weight, vert_index, grad_feature, grad_weight_sum)
    image, vert_weight, vert_index = f32c(image), f32c(vert_weight), i32c(vert_index)
    grad_feature, grad_weight_sum = f32c(grad_feature), f32c(grad_weight_sum)
    C, K = int(image.shape[-1]), int(vert_index.shape[-1])
    R = vert_index.numel() // max(K, 1)
    dev = image.device
    with torch.cuda.device(dev):
        g_image = torch.empty_like(image)
        g_w = torch.empty_like(vert_weight)
        check(lib().voge_sample_backward(ptr(image), ptr(vert_weight), ptr(vert_index), ptr(grad_feature),
                                         ptr(grad_weight_sum), R, K, C, ptr(g_image), ptr(g_w), stream_of(image)),
              "sample_backward")
    return g_image, g_w


def scatter_max(vert_weight, vert_index, num_vert):
    """Reference: ScatterMax, sample_voge.cu:137-170. -> vert_weight_max (N,)."""
    require_cuda(vert_weight, vert_index)
    vert_weight, vert_index = f32c(vert_weight), i32c(vert_index)
    K, N = int(vert_index.shape[-1]), int(num_vert)
    R = vert_index.numel() // max(K, 1)
    dev = vert_weight.device
    with torch.cuda.device(dev):
        wmax = torch.zeros((N,), dtype=torch.float32, device=dev)
        check(lib().voge_scatter_max(ptr(vert_weight), ptr(vert_index), R, K, N, ptr(wmax), stream_of(vert_weight)),
              "scatter_max")
    return wmax


def ray_trace_voge_ray(mus, isigmas, rays):
    """Reference: RayTraceVogeRay, voge_ray_tracing_ray.cu:242-283. -> (len, act, dsd), each (N, M)."""
    require_cuda(mus, isigmas, rays)
    mus, isigmas, rays = f32c(mus), f32c(isigmas), f32c(rays)
    M, N = int(isigmas.shape[0]), int(rays.shape[0])
    dev = mus.device
    with torch.cuda.device(dev):
        out = [torch.empty((N, M), dtype=torch.float32, device=dev) for _ in range(3)]
        check(lib().voge_ray_trace_ray(ptr(mus), ptr(isigmas), ptr(rays), M, N, ptr(out[0]), ptr(out[1]), ptr(out[2]),
                                       stream_of(mus)), "ray_trace_ray")
    return tuple(out)


def ray_trace_voge_ray_backward(mus, isigmas, rays, grad_len, grad_act, grad_dsd):
    """Reference: RayTraceVogeRayBackward, :287-325. -> (grad_ray (N,3), grad_mus (M,3), grad_sig (M,3,3))."""
    require_cuda(mus, isigmas, rays, grad_len, grad_act, grad_dsd)
    mus, isigmas, rays = f32c(mus), f32c(isigmas), f32c(rays)
    grad_len, grad_act, grad_dsd = f32c(grad_len), f32c(grad_act), f32c(grad_dsd)
    M, N = int(isigmas.shape[0]), int(rays.shape[0])
    dev = mus.device
    with torch.cuda.device(dev):
        g_ray = torch.zeros((N, 3), dtype=torch.float32, device=dev)
        g_mus = torch.zeros((M, 3), dtype=torch.float32, device=dev)
        g_sig = torch.zeros((M, 3, 3), dtype=torch.float32, device=dev)
        check(lib().voge_ray_trace_ray_backward(ptr(mus), ptr(isigmas), ptr(rays), ptr(grad_len), ptr(grad_act),
                                                ptr(grad_dsd), M, N, ptr(g_ray), ptr(g_mus), ptr(g_sig),
                                                stream_of(mus)), "ray_trace_ray_backward")
    return g_ray, g_mus, g_sig


def find_nearest_k(hit_len_in, hit_act_in, hit_dsd_in, thr_act, K):
    """Reference: FindNearestK, :328-375. -> (point_idx i32, len, act, dsd), each (N, K)."""
    require_cuda(hit_len_in, hit_act_in, hit_dsd_in)
    hit_len_in, hit_act_in, hit_dsd_in = f32c(hit_len_in), f32c(hit_act_in), f32c(hit_dsd_in)
    N, M, K = int(hit_len_in.shape[0]), int(hit_len_in.shape[1]), int(K)
    dev = hit_len_in.device
    with torch.cuda.device(dev):
        idx = torch.empty((N, K), dtype=torch.int32, device=dev)
        out = [torch.empty((N, K), dtype=torch.float32, device=dev) for _ in range(3)]
        check(lib().voge_find_nearest_k(ptr(hit_len_in), ptr(hit_act_in), ptr(hit_dsd_in), float(thr_act), M, K, N,
                                        ptr(idx), ptr(out[0]), ptr(out[1]), ptr(out[2]), stream_of(hit_len_in)),
              "find_nearest_k")
    return idx, out[0], out[1], out[2]


def knn_mean_dist(points, n_nearest, thr_max):
    """(N,3) CUDA points -> (N,) mean clipped distance to the n_nearest closest points (voge_knn_mean_dist)."""
    points = f32c(points)
    N = int(points.shape[0])
    with torch.cuda.device(points.device):
        out = torch.empty((N,), dtype=torch.float32, device=points.device)
        check(lib().voge_knn_mean_dist(ptr(points), N, int(n_nearest), float(thr_max), ptr(out), stream_of(points)),
              "knn_mean_dist")
    return out


# ---- fused blend ops (PyTorch-only in the reference, Aggregation.py) ---------------------------
def aggregation_forward(sel_idx, sel_act, sel_len, sel_dsd, absorptivity):
    require_cuda(sel_idx, sel_act, sel_len, sel_dsd)
    sel_idx, sel_act, sel_len, sel_dsd = i32c(sel_idx), f32c(sel_act), f32c(sel_len), f32c(sel_dsd)
    K = int(sel_idx.shape[-1])
    R = sel_idx.numel() // max(K, 1)
    dev = sel_act.device
    with torch.cuda.device(dev):
        weight = torch.empty_like(sel_act)
        valid = torch.empty(sel_idx.shape[:-1], dtype=torch.int64, device=dev)
        check(lib().voge_aggregation(ptr(sel_idx), ptr(sel_act), ptr(sel_len), ptr(sel_dsd), float(absorptivity),
                                     R, K, ptr(weight), ptr(valid), stream_of(sel_act)), "aggregation")
    return weight, valid


def aggregation_backward(sel_act, sel_len, sel_dsd, grad_weight, absorptivity):
    sel_act, sel_len, sel_dsd, grad_weight = f32c(sel_act), f32c(sel_len), f32c(sel_dsd), f32c(grad_weight)
    K = int(sel_act.shape[-1])
    R = sel_act.numel() // max(K, 1)
    dev = sel_act.device
    with torch.cuda.device(dev):
        g_act = torch.empty_like(sel_act)
        g_len = torch.empty_like(sel_act)
        g_dsd = torch.empty_like(sel_act)
        check(lib().voge_aggregation_backward(ptr(sel_act), ptr(sel_len), ptr(sel_dsd), ptr(grad_weight),
                                              float(absorptivity), R, K, ptr(g_act), ptr(g_len), ptr(g_dsd),
                                              stream_of(sel_act)), "aggregation_backward")
    return g_act, g_len, g_dsd


def pad_attr4(attr, gauss=None):
    """(n, C <= 4) attribute table -> (n, 4) zero-padded rows (one 16-byte gather per hit in the kernels).  gauss: a
    kind-9 record table (n, 16) of the fused renderer -- the rows are also written into the second 16 bytes of its
    records (voge_pack_attr), where the image-mode backward fetches them together with the geometry."""
    attr = f32c(attr)
    n, C = int(attr.shape[0]), int(attr.shape[1])
    in_rec = gauss is not None and int(gauss.shape[1]) == 16 and int(gauss.shape[0]) == n
    if C == 4 and not in_rec:
        return attr
    with torch.cuda.device(attr.device):
        out = attr if C == 4 else torch.empty((n, 4), dtype=torch.float32, device=attr.device)
        check(lib().voge_pack_attr(ptr(attr), C, n, ptr(out) if C != 4 else None, ptr(gauss) if in_rec else None,
                                   stream_of(attr)), "pack_attr")
    return out


def merge_final_forward(attr, weight, idx, valid_num, background=None, mask_thr=-1.0, idx_mod=0, attr4=None,
                        want_sat_code=False, zero_padding=False):
    """attr4: optional pad_attr4(attr) (used instead of attr when C <= 4).  want_sat_code: also return the (R,) uint8
    clamp code of the background composite (None when there is no background or C > 4).  zero_padding (C <= 4): the
    kernel also writes 0 into idx[..., valid_num:] (the reference's in-place -1 -> 0, for fragments whose padding is
    known to be -1); idx must then be the caller's own contiguous tensor."""
    require_cuda(attr, weight, idx, valid_num)
    attr, weight, idx = f32c(attr), f32c(weight), i32c(idx)
    valid_num = valid_num.to(torch.int64).contiguous()
    K, C = int(idx.shape[-1]), int(attr.shape[-1])
    R = idx.numel() // max(K, 1)
    dev = attr.device
    with torch.cuda.device(dev):
        out = torch.empty(tuple(idx.shape[:-1]) + (C,), dtype=torch.float32, device=dev)
        bg = f32c(background) if background is not None else None
        p4 = attr4 is not None and C <= 4
        code = None
        if want_sat_code and bg is not None and C <= 4:
            code = torch.empty(tuple(idx.shape[:-1]), dtype=torch.uint8, device=dev)
        check(lib().voge_merge_final(ptr(attr4 if p4 else attr), ptr(weight), ptr(idx), ptr(valid_num), ptr(bg),
                                     float(mask_thr), R, K, C, int(idx_mod), int(attr.shape[0]), int(p4), ptr(out),
                                     ptr(code), ptr(idx) if (zero_padding and C <= 4) else None, stream_of(attr)),
              "merge_final")
    return (out, code) if want_sat_code else out


def merge_final_backward(attr, weight, idx, valid_num, grad_out, background=None, mask_thr=-1.0, idx_mod=0,
                         need_attr=True, need_weight=True, attr4=None, out=None):
    attr, weight, idx, grad_out = f32c(attr), f32c(weight), i32c(idx), f32c(grad_out)
    valid_num = valid_num.to(torch.int64).contiguous()
    K, C = int(idx.shape[-1]), int(attr.shape[-1])
    R = idx.numel() // max(K, 1)
    dev = attr.device
    with torch.cuda.device(dev):
        packed4 = need_attr and C <= 4
        if packed4:
            g_attr = torch.zeros((attr.shape[0], 4), dtype=torch.float32, device=dev)
        else:
            g_attr = torch.zeros_like(attr) if need_attr else None
        g_w = torch.empty_like(weight) if need_weight else None
        bg = f32c(background) if background is not None else None
        p4 = attr4 is not None and C <= 4 and (packed4 or g_attr is None)
        check(lib().voge_merge_final_backward(ptr(attr4 if p4 else attr), ptr(weight), ptr(idx), ptr(valid_num), ptr(bg),
                                              float(mask_thr), ptr(f32c(out)) if out is not None else None, ptr(grad_out),
                                              R, K, C, int(idx_mod),
                                              int(attr.shape[0]), int(packed4), int(p4), ptr(g_attr), ptr(g_w),
                                              stream_of(attr)), "merge_final_backward")
        if packed4:
            g_attr = g_attr[:, :C].contiguous()
    return g_attr, g_w


# ---- fused renderer path (no counterpart in the reference's _C: replaces the PyTorch glue +
# ---- rasterize_points_coarse + ray_trace_voge_fine + Aggregation.py chain) ---------------------
def sigma_kind(sigmas):
    if sigmas.dim() == 1:
        return 1
    if sigmas.dim() == 2 and sigmas.shape[1] == 3:
        return 3
    if sigmas.dim() == 3 and sigmas.shape[1] == 3 and sigmas.shape[2] == 3:
        return 9
    raise Exception('Got unexpected sigma, which has shape: ' + str(sigmas.shape))


SIGMA_MODES = {"direct": 0, "inverse": 1, "cholesky": 2}
GAUSS_WIDTH = {1: 4, 3: 8, 9: 16}      # floats per record (kind 9: 64 bytes, the second 16 hold the attribute row)
GRAD_WIDTH = {1: 4, 3: 8, 9: 12}       # floats per packed gradient record
KIND_ISO_ENCODED = 0x100       # sigma_kind flag: kind-9 records with the isotropic encoding (csrc/render_core.cuh)


def record_kind(gauss):
    """sigma_kind argument of the kernels that read the packed records `gauss`"""
    kind = {4: 1, 8: 3, 16: 9}[int(gauss.shape[1])]
    return kind | (KIND_ISO_ENCODED if getattr(gauss, "iso_encoded", False) else 0)
BIN_FLAG_DENSE_MARGIN = 1      # voge_bin_count flags: rounding margin with the dense-S constants (A/B aid)


def make_cam(R, focal, principal):
    """(B,16) f32 per-view camera records [R row-major (9), fx, fy, px, py, 0, 0, 0] read by the kernels that
    generate their rays (render_core.cuh: gen_ray).  Plain torch ops: differentiable w.r.t. R / focal / principal."""
    B = int(R.shape[0])
    return torch.cat([R.reshape(B, 9), focal.reshape(B, 2), principal.reshape(B, 2),
                      torch.zeros((B, 3), dtype=R.dtype, device=R.device)], dim=1).to(torch.float32).contiguous()


def generate_rays(cam, image_size):
    """Closed-form unit ray directions through the pixel centres (reference Renderer.py:124-128), (B,H,W,3):
    the materialised form of the generator the fused kernels evaluate in registers -- same bits."""
    cam = f32c(cam)
    B, H, W = int(cam.shape[0]), int(image_size[0]), int(image_size[1])
    with torch.cuda.device(cam.device):
        rays = torch.empty((B, H, W, 3), dtype=torch.float32, device=cam.device)
        check(lib().voge_generate_rays(ptr(cam), B, H, W, ptr(rays), stream_of(cam)), "generate_rays")
    return rays


class BinPlan(object):
    """Scratch sizes carried from one renderer call to the next of the same shape (speculative binning).

    bin_views needs two totals that only the device knows after voge_bin_count -- the tile-list entries and the
    items (hit slots) per view -- to size tile_list and the hit segments.  Reading them back costs a host round
    trip in the middle of the forward (the launch queue drains).  With a plan the scratch is
    sized from the PREVIOUS call (+ 12.5 %), every launch of the forward is queued without waiting, the kernels
    never step outside the capacities (include/voge_b200.h "Speculative scratch"), and the true totals arrive
    through a pinned buffer that the host checks after the last launch (bins_valid): the GPU is busy with the
    queued kernels while the host waits for an event that completed long ago.  A violated capacity (the scene
    grew by more than the slack) repeats the call with exact sizes."""
    SLACK_NUM, SLACK_DEN = 9, 8
    __slots__ = ("list_cap", "view_items", "hits_cap", "pinned", "done")

    def __init__(self, total_entries, max_view_items):
        self.pinned = self.done = None
        self.update(total_entries, max_view_items)

    def update(self, total_entries, max_view_items):
        self.list_cap = int(total_entries) * self.SLACK_NUM // self.SLACK_DEN + 1024
        self.view_items = int(max_view_items) * self.SLACK_NUM // self.SLACK_DEN + 4096
        self.hits_cap = 0

    def mailbox(self, n):
        """The plan's own pinned buffer + event for the totals, allocated once: a fresh pinned tensor per call goes
        through the caching host allocator, whose blocks become reusable only after an event recorded when they are
        FREED (behind everything queued by then) -- with the GPU a step behind that means a cudaHostAlloc per call."""
        if self.pinned is None or int(self.pinned.numel()) != n:
            self.pinned = torch.empty((n,), dtype=torch.int64, pin_memory=True)
            self.done = torch.cuda.Event()
        return self.pinned, self.done

    def groups(self, B, max_group_items):
        """Views per trace / select launch so that a group's items fit the scratch budget, from the largest view of the
        previous call (independent of the order of the cameras)."""
        per = max(1, min(B, int(max_group_items) // max(self.view_items, 1)))
        self.hits_cap = per * self.view_items
        return [(b0, min(b0 + per, B)) for b0 in range(0, B, per)]


_bin_plans = {}
MAX_GROUP_ITEMS = 1 << 29      # 8 bytes of hit scratch per item
# Set by voge_b200.graphs.GraphedStep while a step is captured into a CUDA graph: the host can not wait inside a
# capture, so the capacity check runs on the device (a sticky violation flag that the graph's owner reads after a
# replay) and a shape without a BinPlan is an error (run the step eagerly once before capturing).
capture_state = None


def speculation_enabled():
    return os.environ.get("VOGE_NO_SPECULATION") != "1"


def bin_views(verts, sigmas, R, T, origins, focal, principal, image_size, thr, thr_act, use_ref_bins, bin_size,
              tile, gauss=None, sigma_mode=0, flags=None, speculate=False, max_group_items=None):
    """-> (tile_offsets (B*TY*TX*S+1,) int64, tile_list (total, 8) int32 = (index, rectangle x, rectangle y, 0 | first 16 bytes of the record), rects (B,N,2) int32,
    tile_item_offsets (B*TY*TX*S+1,) int64 with .total_items), S = voge_bin_sub() list segments per tile.
    One host sync (the two totals) -- or none with speculate=True once a call of the same shape has left a BinPlan:
    the result then carries `.spec` and the caller must check bins_valid(item_offsets) after queueing the forward
    and repeat with speculate=False when it returns False."""
    R, T, origins, focal, principal = f32c(R), f32c(T), f32c(origins), f32c(focal), f32c(principal)
    if gauss is None:
        gauss = pack_gaussians(verts, sigmas, sigma_mode)
    B, N = int(R.shape[0]), int(gauss.shape[0])
    kind = record_kind(gauss)
    H, W = int(image_size[0]), int(image_size[1])
    TX, TY = (W + tile - 1) // tile, (H + tile - 1) // tile
    if flags is None:
        flags = BIN_FLAG_DENSE_MARGIN if os.environ.get("VOGE_DENSE_MARGIN") == "1" else 0
    dev = gauss.device
    max_group_items = MAX_GROUP_ITEMS if max_group_items is None else int(max_group_items)
    key = (dev.index, B, N, H, W, int(tile), bool(use_ref_bins), int(bin_size), kind, int(max_group_items),
           bool(getattr(gauss, "iso_encoded", False)))
    plan = _bin_plans.get(key) if (speculate and speculation_enabled()) else None
    if capture_state is None and torch.cuda.is_current_stream_capturing():
        raise RuntimeError("voge_b200: the renderer was called inside a CUDA-graph capture it does not know about; "
                           "capture the step with voge_b200.graphs.GraphedStep (its capacity check runs on the device)")
    if capture_state is not None and plan is None:
        raise RuntimeError("voge_b200: a renderer call of a new shape inside a CUDA-graph capture (no BinPlan yet, or "
                           "speculation disabled): run the step eagerly at least once before capturing it")
    with torch.cuda.device(dev):
        rects = torch.empty((B, N, 2), dtype=torch.int32, device=dev)
        # one 64-bit counter per list segment (low word: list entries, high word: items = rectangle pixels), filled by
        # ONE reduction per (entry, tile); a leading zero row so that the inclusive scans are the exclusive offsets
        S = int(lib().voge_bin_sub())
        counters = torch.zeros((B * TY * TX * S + 1, 2), dtype=torch.int32, device=dev)
        slack_items = 0
        if os.environ.get("VOGE_NO_SEGMENT_ALIGN") != "1":
            # room for 32-byte aligned pixel segments (include/voge_b200.h "Segment alignment")
            counters[1:, 1] = int(lib().voge_bin_item_slack())
            slack_items = int(lib().voge_bin_item_slack()) * B * TY * TX * S
        check(lib().voge_bin_count(ptr(gauss), kind, ptr(R), ptr(T), ptr(origins),
                                   ptr(focal), ptr(principal), B, N, H, W, float(thr), float(thr_act),
                                   int(bool(use_ref_bins)), int(bin_size), int(tile), int(flags), ptr(rects),
                                   ptr(counters[1:]), stream_of(gauss)), "bin_count")
        # two 1-D scans (cub DeviceScan); a (2, n) scan along dim 1 runs one thread block per row
        offsets = (torch.cumsum(counters[:, 0], 0, dtype=torch.int64), torch.cumsum(counters[:, 1], 0, dtype=torch.int64))
        # the totals the host needs: list entries + the item count at every view boundary (views can then be
        # processed in groups that bound the forward's scratch) + the isotropic encoding's ambiguity flag
        per_view = TY * TX * S
        iso_flag = getattr(gauss, "iso_flag", None)
        totals = torch.cat([offsets[0][-1:], offsets[1][::per_view]] + ([iso_flag.to(torch.int64)] if iso_flag is not None else []))
        item_offsets = offsets[1]
        item_offsets.slack_items = slack_items      # slots in the totals that are alignment room, not items
        if plan is None:
            host = totals.tolist()             # the one host sync of the exact path
            if iso_flag is not None:
                # the isotropic encoding is ambiguous for this scene (a non-encoded record has a negative S00): the
                # caller re-packs plain records and bins again (the same host sync told us)
                gauss.iso_bad = bool(host.pop())
            total, view_item_starts = int(host[0]), [int(v) for v in host[1:]]
            vmax = max(b - a for a, b in zip(view_item_starts[:-1], view_item_starts[1:]))
            if key in _bin_plans:
                _bin_plans[key].update(total, vmax)
            else:
                _bin_plans[key] = BinPlan(total, vmax)
            list_cap = 0
            tile_list = torch.empty((max(total, 1), 8), dtype=torch.int32, device=dev)
            item_offsets.total_items = view_item_starts[-1]
            item_offsets.view_item_starts = view_item_starts      # B + 1 host ints
        else:
            groups = plan.groups(B, max_group_items)
            list_cap = plan.list_cap
            if capture_state is None:
                pinned, done = plan.mailbox(int(totals.numel()))
                pinned.copy_(totals, non_blocking=True)
                done.record()
            else:
                # inside a graph capture: the same check as bins_valid, evaluated by the device on every replay
                # (a group holds at most hits_cap / view_items views, so the largest view bounds every group)
                pinned = done = None
                starts = totals[1:B + 2]
                bad = (totals[0] > list_cap) | ((starts[1:] - starts[:-1]).max() > plan.view_items)
                if iso_flag is not None:
                    bad = bad | (totals[-1] != 0)
                capture_state.violation.logical_or_(bad)
            tile_list = torch.empty((list_cap, 8), dtype=torch.int32, device=dev)
            item_offsets.spec = {"key": key, "plan": plan, "groups": groups, "pinned": pinned,
                                 "done": done, "has_iso_flag": iso_flag is not None, "gauss": gauss}
        cursor = offsets[0][:-1].clone()      # every segment's cursor starts at its offset: one atomic yields the position
        check(lib().voge_bin_fill(ptr(rects), ptr(gauss), kind, ptr(cursor), B, N, H, W, int(tile), ptr(tile_list),
                                  int(list_cap), stream_of(gauss)), "bin_fill")
    return offsets[0], tile_list, rects, item_offsets


def bins_valid(item_offsets):
    """True if the scratch of a speculative bin_views (and of the render_forward that followed) held everything.
    Waits for the totals' copy (queued right after the scans, i.e. long finished while the forward's kernels run),
    refreshes the shape's BinPlan from the true totals either way, and sets gauss.iso_bad like the exact path."""
    spec = getattr(item_offsets, "spec", None)
    if spec is None:
        return True
    if spec["done"] is None:
        return True        # captured into a CUDA graph: validated on the device (capture_state.violation)
    spec["done"].synchronize()
    host = spec["pinned"].tolist()
    iso_bad = bool(host.pop()) if spec["has_iso_flag"] else False
    spec["gauss"].iso_bad = iso_bad
    total, starts = int(host[0]), [int(v) for v in host[1:]]
    plan = spec["plan"]
    ok = total <= plan.list_cap and all(starts[b1] - starts[b0] <= plan.hits_cap for b0, b1 in spec["groups"])
    if iso_bad:
        _bin_plans.pop(spec["key"], None)      # such a scene keeps to the exact pass (which re-packs plain records)
    else:
        plan.update(total, max(b - a for a, b in zip(starts[:-1], starts[1:])))
    item_offsets.total_items, item_offsets.view_item_starts = starts[-1], starts
    return ok and not iso_bad


def pack_gaussians(verts, sigmas, sigma_mode=0, iso_encode=False):
    """-> (N, 4 | 8 | 12) f32 records [x, y, z, S = 2 P ...] (16-byte aligned) read by the fused kernels; P = sigmas
    (mode 0), inverse(sigmas) (mode 1, reference inverse_sigma=True) or tril(sigmas) tril(sigmas)^T (mode 2).
    iso_encode ((N,3,3) sigmas only): records whose S is exactly s I carry -s in the S00 slot, so a hit gathers 16
    instead of 48 bytes (csrc/render_core.cuh: kKindIsoEncoded).  The result then has .iso_encoded = True and
    .iso_flag, a device int32 that the kernel sets when the encoding is ambiguous (some other record has a negative
    S00); bin_views reads it with its own host sync and sets .iso_bad, on which the caller packs plain records."""
    verts, sigmas = f32c(verts), f32c(sigmas)
    N, kind = int(verts.shape[0]), sigma_kind(sigmas)
    iso_encode = bool(iso_encode) and kind == 9 and os.environ.get("VOGE_NO_ISO_ENCODING") != "1"
    with torch.cuda.device(verts.device):
        out = torch.empty((N, GAUSS_WIDTH[kind]), dtype=torch.float32, device=verts.device)
        flag = torch.zeros((1,), dtype=torch.int32, device=verts.device) if iso_encode else None
        check(lib().voge_pack_gaussians(ptr(verts), ptr(sigmas), kind, int(sigma_mode), N, ptr(out), ptr(flag),
                                        stream_of(verts)), "pack_gaussians")
    if iso_encode:
        out.iso_encoded, out.iso_flag = True, flag
    return out


def render_forward(verts, sigmas, origins, rays, tile_offsets, tile_list, rects, thr_act, absorptivity, K, tile,
                   need_act=True, stats=None, item_offsets=None, gauss=None, max_group_items=None, debug=None,
                   cam=None, image_size=None, sigma_mode=0):
    """Fragments of the fused renderer: trace_hits -> select_topk -> blend_weights over the tile lists of
    bin_views (item_offsets = its fourth result); no per-pixel capacity limit; the views are traced in groups of
    at most max_group_items items (8 bytes of scratch each).  rays (B,H,W,3), or None with cam (B,16) and
    image_size: the kernels generate the rays themselves."""
    origins = f32c(origins)
    if rays is not None:
        rays = f32c(rays)
        B, H, W = int(rays.shape[0]), int(rays.shape[1]), int(rays.shape[2])
    else:
        cam = f32c(cam)
        B, H, W = int(cam.shape[0]), int(image_size[0]), int(image_size[1])
    if item_offsets is None:
        raise RuntimeError("voge_b200.render_forward: item_offsets (bin_views' fourth result) is required")
    max_group_items = MAX_GROUP_ITEMS if max_group_items is None else int(max_group_items)
    if gauss is None:
        gauss = pack_gaussians(verts, sigmas, sigma_mode)
    N, K = int(gauss.shape[0]), int(K)
    skind = record_kind(gauss)
    dev = gauss.device
    with torch.cuda.device(dev):
        idx = torch.empty((B, H, W, K), dtype=torch.int32, device=dev)
        weight = torch.empty((B, H, W, K), dtype=torch.float32, device=dev)
        tlen = torch.empty((B, H, W, K), dtype=torch.float32, device=dev)
        valid = torch.empty((B, H, W), dtype=torch.int64, device=dev)
        act = torch.empty((B, H, W, K), dtype=torch.float32, device=dev) if need_act else None
        dsd = torch.empty((B, H, W, K), dtype=torch.float32, device=dev) if need_act else None
        st = stream_of(gauss)
        nt = int(lib().voge_trace_threads(int(tile)))
        S = int(lib().voge_bin_sub())
        tiles_per_view = (int(tile_offsets.numel()) - 1) // S // max(B, 1)
        starts = getattr(item_offsets, "view_item_starts", None)
        spec = getattr(item_offsets, "spec", None)
        list_cap = hits_cap = 0
        if spec is not None:
            # speculative scratch (bin_views(speculate=True)): groups and capacities from the shape's BinPlan, the
            # kernels read the group's item base from the device and skip what does not fit; the caller validates
            groups, hits_cap, list_cap = spec["groups"], int(spec["plan"].hits_cap), int(spec["plan"].list_cap)
            starts = None
        elif starts is None:
            starts = [0] * B + [int(item_offsets.total_items)]
            groups = [(0, B)]
        else:
            # groups of consecutive views whose items fit the scratch budget (a single view always forms a group)
            groups, b0 = [], 0
            while b0 < B:
                b1 = b0 + 1
                while b1 < B and starts[b1 + 1] - starts[b0] <= max_group_items:
                    b1 += 1
                groups.append((b0, b1))
                b0 = b1
        cap = max(max(starts[b1] - starts[b0] for b0, b1 in groups), 1) if starts is not None else max(hits_cap, 1)
        counts = torch.empty((B * tiles_per_view * nt,), dtype=torch.int32, device=dev)
        seg_base = torch.empty((B * tiles_per_view * nt,), dtype=torch.int64, device=dev)
        hits = torch.empty((cap, 2), dtype=torch.int32, device=dev)

        def sl(t, b0, b1, per):        # rows [b0*per, b1*per) of a flat per-view table
            return t[b0 * per:b1 * per] if t is not None else None
        for b0, b1 in groups:
            nb = b1 - b0
            t_off = tile_offsets[b0 * tiles_per_view * S:]
            i_off = item_offsets[b0 * tiles_per_view * S:]
            c_g, s_g = sl(counts, b0, b1, tiles_per_view * nt), sl(seg_base, b0, b1, tiles_per_view * nt)
            check(lib().voge_trace_hits(ptr(gauss), skind, ptr(origins[b0:b1]),
                                        ptr(rays[b0:b1]) if rays is not None else None,
                                        ptr(cam[b0:b1]) if cam is not None else None, ptr(t_off),
                                        ptr(tile_list), ptr(rects[b0:b1]), ptr(i_off),
                                        int(starts[b0]) if starts is not None else -1, float(thr_act),
                                        nb, N, H, W, int(tile), ptr(c_g), ptr(s_g), ptr(hits), list_cap, hits_cap,
                                        ptr(stats), st),
                  "trace_hits")
            check(lib().voge_select_topk(ptr(c_g), ptr(s_g), ptr(hits), b0, nb, N, H, W, K, int(tile),
                                         ptr(idx[b0:b1]), ptr(valid[b0:b1]), ptr(stats), st), "select_topk")
        check(lib().voge_blend_weights(ptr(gauss), skind, ptr(origins), ptr(rays), ptr(cam), ptr(idx), ptr(valid),
                                       float(absorptivity), 0, B, N, H, W, K, ptr(weight), ptr(tlen), ptr(act), ptr(dsd),
                                       st), "blend_weights")
        if debug is not None:          # development aid (tools/select_hist.py): the per-pixel hit counts of the trace
            debug["counts"], debug["threads_per_tile"] = counts, nt
    return idx, weight, tlen, valid, act, dsd


def _bwd_flags(need_sigma):
    return int(bool(need_sigma))


def render_backward_fused(verts, sigmas, origins, rays, idx, valid, g_weight, g_len_out, absorptivity,
                          need_sigma=True, need_rays=False, need_origins=False, gauss=None, weight=None,
                          cam=None, need_cam=False, sigma_mode=0):
    """-> (g_verts, g_sigmas | None, g_rays (B,H,W,3) | None, g_origins (B,3) | None, g_cam (B,16) | None).
    rays None: the kernel generates them from cam (B,16); need_cam then returns the gradient of the camera
    records (d/dR, d/dfocal, d/dprincipal summed over each view's pixels in the kernel)."""
    verts, sigmas, origins = f32c(verts), f32c(sigmas), f32c(origins)
    rays = f32c(rays) if rays is not None else None
    cam = f32c(cam) if cam is not None else None
    idx, g_weight = i32c(idx), f32c(g_weight)
    g_len_out = f32c(g_len_out) if g_len_out is not None else None
    weight = f32c(weight) if weight is not None else None
    B, H, W, K = (int(s) for s in idx.shape)
    N = int(verts.shape[0])
    kind = sigma_kind(sigmas)
    dev = verts.device
    with torch.cuda.device(dev):
        if gauss is None:
            gauss = pack_gaussians(verts, sigmas, sigma_mode)
        packed = torch.zeros((N, GRAD_WIDTH[kind]), dtype=torch.float32, device=dev)
        g_rays = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev) if (need_rays and rays is not None) else None
        g_org = torch.zeros((B, 3), dtype=torch.float32, device=dev) if need_origins else None
        g_cam = torch.zeros((B, 16), dtype=torch.float32, device=dev) if (need_cam and rays is None) else None
        check(lib().voge_render_backward_fused(ptr(gauss), record_kind(gauss), ptr(origins), ptr(rays),
                                               ptr(idx), ptr(valid), ptr(g_weight), ptr(weight), ptr(g_len_out),
                                               float(absorptivity), B, N, H, W, K, ptr(packed), _bwd_flags(need_sigma),
                                               ptr(g_rays), ptr(g_org), ptr(cam), ptr(g_cam), stream_of(verts)),
              "render_backward_fused")
        g_verts, g_sig = unpack_gradients(packed, gauss, sigmas, sigma_mode, need_sigma)
    return g_verts, g_sig, g_rays, g_org, g_cam


def render_backward_image(verts, sigmas, origins, rays, idx, valid, weight, grad_out, fwd_out, attr4, background,
                          mask_thr, absorptivity, sat_code=None, need_sigma=True, need_attr=True, need_rays=False, need_origins=False,
                          gauss=None, cam=None, need_cam=False, sigma_mode=0, n_channels=3, need_geometry=True,
                          attr_in_records=False):
    """Fused backward with merge_final's backward folded in (voge_render_backward_image): the gradient of the
    composited image -> (g_verts, g_sigmas | None, g_attr (N,C) | None, g_rays | None, g_origins | None, g_cam | None)."""
    verts, sigmas, origins = f32c(verts), f32c(sigmas), f32c(origins)
    rays = f32c(rays) if rays is not None else None
    cam = f32c(cam) if cam is not None else None
    idx, weight, grad_out, attr4 = i32c(idx), f32c(weight), f32c(grad_out), f32c(attr4)
    fwd_out = f32c(fwd_out) if fwd_out is not None else None
    background = f32c(background) if background is not None else None
    B, H, W, K = (int(s) for s in idx.shape)
    N, C = int(verts.shape[0]), int(n_channels)
    kind = sigma_kind(sigmas)
    dev = verts.device
    with torch.cuda.device(dev):
        if gauss is None:
            gauss = pack_gaussians(verts, sigmas, sigma_mode)
        packed = torch.zeros((N, GRAD_WIDTH[kind]), dtype=torch.float32, device=dev)
        g_attr4 = torch.zeros((N, 4), dtype=torch.float32, device=dev) if need_attr else None
        g_rays = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev) if (need_rays and rays is not None) else None
        g_org = torch.zeros((B, 3), dtype=torch.float32, device=dev) if need_origins else None
        g_cam = torch.zeros((B, 16), dtype=torch.float32, device=dev) if (need_cam and rays is None) else None
        check(lib().voge_render_backward_image(ptr(gauss), record_kind(gauss), ptr(origins), ptr(rays), ptr(cam), ptr(idx), ptr(valid),
                                               ptr(weight), ptr(grad_out), ptr(fwd_out), ptr(sat_code), ptr(attr4),
                                               ptr(background),
                                               float(mask_thr), C, float(absorptivity), B, N, H, W, K, ptr(packed),
                                               _bwd_flags(need_sigma) | (2 if (attr_in_records and int(gauss.shape[1]) == 16) else 0),
                                               ptr(g_attr4), ptr(g_rays), ptr(g_org), ptr(g_cam),
                                               stream_of(verts)), "render_backward_image")
        g_verts, g_sig = unpack_gradients(packed, gauss, sigmas, sigma_mode, need_sigma) if need_geometry else (None, None)
    g_attr = g_attr4[:, :C].contiguous() if need_attr else None
    return g_verts, g_sig, g_attr, g_rays, g_org, g_cam


def unpack_gradients(packed, gauss, sigmas, sigma_mode=0, need_sigma=True):
    """Packed per-Gaussian gradient records -> (g_verts (N,3), g_sigmas shaped like `sigmas` | None), with the
    chain rule of the sigma parameterisation (inverse / Cholesky) applied (voge_unpack_gradients)."""
    sigmas = f32c(sigmas)
    N, kind = int(packed.shape[0]), sigma_kind(sigmas)
    with torch.cuda.device(packed.device):
        g_verts = torch.empty((N, 3), dtype=torch.float32, device=packed.device)
        g_sig = torch.empty_like(sigmas) if need_sigma else None
        check(lib().voge_unpack_gradients(ptr(packed), ptr(gauss), ptr(sigmas), record_kind(gauss) if gauss is not None else kind,
                                          int(sigma_mode), N, ptr(g_verts),
                                          ptr(g_sig), stream_of(packed)), "unpack_gradients")
    return g_verts, g_sig

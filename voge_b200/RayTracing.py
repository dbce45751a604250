"""Ray-tracing op glue and autograd Functions (reference VoGE/RayTracing.py).

Same public names / signatures / return orders as the reference:
  ray_tracing :12-30, convert_to_box :33-39, rasterize_coarse :42-73, ray_tracing_fine :76-95,
  _RasterizeCoarse :126-151, _RayTraceVoGE :154-206.
The native calls go to libvoge_b200.so through voge_b200._C (the `VoGE._C`-shaped shim).
"""
import math
from typing import Optional

import numpy as np
import torch

from . import _C

inf = 1e8


def default_bin_size(image_size) -> int:
    """bin_size heuristic of the reference (:14-16): 128..320 -> 10, 400/512 -> 16, 1024 -> 32."""
    return max(int(2 ** np.ceil(np.log2(max(image_size)) - 5)), 10)


def default_max_points_per_bin(n_assign: int, n_points: int) -> int:
    """:18-19"""
    return min(int(max(n_assign * 10, n_points / 10)), n_points)


def ray_tracing(transforms, points, isigmas, rays, image_size, thr: float, n_assign: int,
                bin_size: Optional[int] = None, max_points_per_bin: Optional[int] = None, **kwargs):
    """cameras, points (B,N,3) camera-centred, isigmas (B,N,3,3), rays (B,H,W,3)
    -> (sel_idx i32, sel_len, sel_act, sel_dsd), each (B,H,W,K); indices are packed b*N+n."""
    if bin_size is None:
        bin_size = default_bin_size(image_size)
    n_points = points.shape[1]
    if max_points_per_bin is None:
        max_points_per_bin = default_max_points_per_bin(n_assign, n_points)

    if max_points_per_bin == -1:
        # no coarse stage: every Gaussian of the view is a candidate for every pixel.  The
        # reference materialises a (B, BW, BW, N) int32 arange table (:22-26); the kernel
        # generates the same candidate list on the fly instead.
        bin_points, bin_counts = None, None
    else:
        bin_points, bin_counts = rasterize_coarse(transforms, points, isigmas, image_size, thr, bin_size,
                                                  max_points_per_bin, return_counts=True, **kwargs)
    return ray_tracing_fine(points.reshape(-1, 3), isigmas.reshape(-1, 3, 3), rays, bin_points, thr, bin_size,
                            n_assign, bin_counts=bin_counts, points_per_view=n_points)


def convert_to_box(isigmas, thr, z, matrix):
    """Reference bbox heuristic (:33-39): radii = sqrt(colsum(-ln(thr) F inv(S_view[:2,:2]) F)) * z."""
    F2 = matrix[:, None, :2, :2]
    get = -np.log(thr) * F2 @ torch.inverse(isigmas[:, :, :2, :2]) @ F2
    ones = torch.ones((*isigmas.shape[0:2], 1, 2), device=isigmas.device)
    return (ones @ get).pow(.5).squeeze(2) * z.unsqueeze(-1)


def coarse_inputs(cameras, points, isigmas, thr):
    """NDC centres (x,y flipped to +x right / +y down, z = view depth) and bbox radii, computed
    with the camera object's own transforms exactly as the reference does (:45-57)."""
    C = -torch.matmul(torch.inverse(cameras.R.transpose(1, 2)), cameras.T[:, :, None])
    points = points + C.view(-1, 1, 3)
    to_ndc = cameras.get_ndc_camera_transform()
    full = cameras.get_full_projection_transform().compose(to_ndc)
    points_ndc = -full.transform_points(points)
    w2v = cameras.get_world_to_view_transform()
    rot = w2v.get_matrix()[:, :3, :3].unsqueeze(1).expand(-1, isigmas.shape[1], -1, -1)
    isig_view = rot.transpose(2, 3) @ isigmas @ rot
    boxes = convert_to_box(isig_view, thr, -points_ndc[..., -1],
                           cameras.get_projection_transform().compose(to_ndc).get_matrix())
    points_ndc[..., 2] = w2v.transform_points(points)[..., 2]
    return points_ndc, boxes


def rasterize_coarse(cameras, points, isigmas, image_size, thr, bin_size, max_points_per_bin,
                     cloud_to_point=None, num_points_per_cloud=None, return_counts=False):
    """-> bin_points (B,BH,BW,M) int32, -1 padded, packed indices (reference :42-73)."""
    points_ndc, boxes = coarse_inputs(cameras, points, isigmas, thr)
    B, N = points_ndc.shape[0], points_ndc.shape[1]
    if cloud_to_point is None:
        cloud_to_point = torch.arange(B, dtype=torch.long, device=points.device) * N
    if num_points_per_cloud is None:
        num_points_per_cloud = torch.full((B,), N, dtype=torch.long, device=points.device)
    bin_points = _RasterizeCoarse.apply(points_ndc.reshape(-1, 3), cloud_to_point, num_points_per_cloud,
                                        image_size, boxes.reshape(-1, 2), bin_size, max_points_per_bin)
    if return_counts:
        return bin_points, _C.last_bin_counts
    return bin_points


def ray_tracing_fine(mus, isigmas, rays, bin_points, thr, bin_size, n_assign, inf=1e10, bin_counts=None,
                     points_per_view=None):
    """mus (P,3), isigmas (P,3,3), rays (B,H,W,3), bin_points (B,BH,BW,M) -> fragments (reference :76-95).
    bin_points=None means "no coarse stage": all `points_per_view` Gaussians of view b are candidates."""
    assert isigmas.dim() == 3
    assert mus.dim() == 2
    assert rays.dim() == 4
    assert bin_points is None or bin_points.dim() == 4
    assert mus.shape[0] == isigmas.shape[0] and mus.shape[1] == 3 and isigmas.shape[1] == 3 and isigmas.shape[2] == 3
    thr_act = -math.log(thr + 1 / inf)
    return _RayTraceVoGE.apply(mus, isigmas, rays, bin_points, thr_act, bin_size, n_assign, bin_counts,
                               points_per_view)


def ray_trace_voge_ray(mus, sigmas, rays):
    """Dense (rays x Gaussians) hit quantities without culling or top-K (reference :97-108):
    mus (M,3), sigmas float | (M,) | (M,3,3), rays (N,3) -> (hit_len, hit_act, hit_dsd), each (N,M)."""
    if isinstance(sigmas, (float, int)):
        sigmas = torch.eye(3, device=mus.device)[None].expand(mus.shape[0], -1, -1) * sigmas
    if sigmas.dim() == 1:
        sigmas = sigmas.view(-1, 1, 1) * torch.eye(3, device=sigmas.device)[None]
    assert mus.is_cuda and sigmas.is_cuda and rays.is_cuda
    assert mus.dim() == 2 and mus.shape[1] == 3
    assert rays.dim() == 2 and rays.shape[1] == 3
    assert sigmas.dim() == 3 and sigmas.shape[1] == 3 and sigmas.shape[2] == 3
    return _RayTraceVoGERay.apply(mus, sigmas, rays)


def find_nearest_k(hit_len_in, hit_act_in, hit_dsd_in, K, thr):
    """K nearest hits (ascending length) with activation below the threshold (reference :111-115)."""
    assert hit_len_in.is_cuda and hit_act_in.is_cuda and hit_dsd_in.is_cuda
    thr_act = -math.log(thr + 1 / inf)
    return _FindNearestK.apply(hit_len_in, hit_act_in, hit_dsd_in, thr_act, K)


def find_farest_k(hit_len_in, hit_act_in, hit_dsd_in, K, thr):
    """K farthest hits: nearest-K on negated lengths (reference :118-123)."""
    assert hit_len_in.is_cuda and hit_act_in.is_cuda and hit_dsd_in.is_cuda
    thr_act = -math.log(thr + 1 / inf)
    point_idx, hit_len, hit_act, hit_dsd = _FindNearestK.apply(-hit_len_in, hit_act_in, hit_dsd_in, thr_act, K)
    return point_idx, -hit_len, hit_act, hit_dsd


class _RayTraceVoGERay(torch.autograd.Function):
    """reference :209-220"""

    @staticmethod
    def forward(ctx, mus, sigmas, rays):
        hit_len, hit_act, hit_dsd = _C.ray_trace_voge_ray(mus, sigmas, rays)
        ctx.save_for_backward(mus, sigmas, rays)
        return hit_len, hit_act, hit_dsd

    @staticmethod
    def backward(ctx, grad_hit_len, grad_hit_act, grad_hit_dsd):
        mus, sigmas, rays = ctx.saved_tensors
        grad_ray, grad_mus, grad_sig = _C.ray_trace_voge_ray_backward(
            mus, sigmas, rays, grad_hit_len.contiguous(), grad_hit_act.contiguous(), grad_hit_dsd.contiguous())
        return grad_mus, grad_sig, grad_ray


class _FindNearestK(torch.autograd.Function):
    """reference :223-241.  The reference's backward scatters the act / dsd gradients onto the tensor that
    already holds the len gradient (:238-240, a defect); here each gradient goes to its own input."""

    @staticmethod
    def forward(ctx, hit_len_in, hit_act_in, hit_dsd_in, thr, K):
        point_idx, hit_len, hit_act, hit_dsd = _C.find_nearest_k(hit_len_in, hit_act_in, hit_dsd_in, thr, K)
        ctx.save_for_backward(point_idx)
        ctx.in_shape = hit_len_in.shape
        ctx.mark_non_differentiable(point_idx)
        return point_idx, hit_len, hit_act, hit_dsd

    @staticmethod
    def backward(ctx, grad_point_idx, grad_hit_len, grad_hit_act, grad_hit_dsd):
        (point_idx,) = ctx.saved_tensors
        valid = point_idx >= 0
        index = point_idx.clamp(min=0).long()
        outs = []
        for g in (grad_hit_len, grad_hit_act, grad_hit_dsd):
            z = torch.zeros(ctx.in_shape, dtype=g.dtype, device=g.device)
            outs.append(z.scatter_add(1, index, g * valid))
        return outs[0], outs[1], outs[2], None, None


class _RasterizeCoarse(torch.autograd.Function):
    """Non-differentiable binning (reference :126-151)."""

    @staticmethod
    def forward(ctx, points_ndc, cloud_to_point, num_points_per_cloud, image_size, boxes, bin_size,
                max_points_per_bin):
        out = _C.rasterize_points_coarse(points_ndc, cloud_to_point, num_points_per_cloud, image_size, boxes,
                                         bin_size, max_points_per_bin)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_idx):
        return (None,) * 7


class _RayTraceVoGE(torch.autograd.Function):
    """The RayTracing autograd Function (reference :154-206).  forward -> (sel_idx, sel_len, sel_act,
    sel_dsd); backward recomputes the quadratic forms and returns (grad_mus, grad_isg, grad_rays)."""

    @staticmethod
    def forward(ctx, mus, isigmas, rays, bin_points, thr_act, bin_size, n_assign, bin_counts=None,
                points_per_view=None):
        if bin_points is None:
            sel = _C.ray_trace_voge_fine_dense(mus, isigmas, rays, points_per_view, thr_act, bin_size, n_assign)
        else:
            sel = _C.ray_trace_voge_fine(mus, isigmas, rays, bin_points, thr_act, bin_size, n_assign,
                                         bin_counts=bin_counts)
        sel_idx, sel_len, sel_act, sel_dsd = sel
        ctx.save_for_backward(mus, isigmas, rays, sel_idx)
        ctx.mark_non_differentiable(sel_idx)
        return sel_idx, sel_len, sel_act, sel_dsd

    @staticmethod
    def backward(ctx, grad_sel_idx, grad_sel_len, grad_sel_act, grad_sel_dsd):
        mus, isigmas, rays, sel_idx = ctx.saved_tensors
        grad_rays, grad_mus, grad_isg = _C.ray_trace_voge_fine_backward(
            mus, isigmas, rays, sel_idx, grad_sel_len.contiguous(), grad_sel_act.contiguous(),
            grad_sel_dsd.contiguous(), need_rays=ctx.needs_input_grad[2])
        return grad_mus, grad_isg, grad_rays, None, None, None, None, None, None

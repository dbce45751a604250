"""Blend weights and attribute gather-blend (reference VoGE/Aggregation.py).

Public names and signatures follow the reference:
  inverse_cumsum :7, get_ray_camera_space :11-27, get_cross_activation :30-51,
  assign2weight :54-79, aggregation :82-107, merge_final :111-141, expend_sigma :144-175.

`aggregation` and `merge_final` are the hot functions: in the reference they run as ~15-30
PyTorch kernels over materialised (R,K,K) / (R,K,C) tensors; here each is ONE hand-written CUDA
kernel forward and one backward (voge_b200/csrc/blend.cu) behind a torch.autograd.Function.
The small helper formulas (get_cross_activation / assign2weight) stay plain torch expressions:
they are the specification the kernels are tested against and are not on the fast path.
"""
import math

import torch
import torch.nn.functional as F

from . import _C
from .Utils import ind_fill, ind_sel  # noqa: F401  (re-exported like the reference module)

# merge_final's `assert vert_attr.shape[0] > vert_assign.max()` (reference :120) forces a host sync
# per call.  The kernels ignore out-of-range indices instead; set True to get the reference assert.
STRICT_INDEX_ASSERT = False


def inverse_cumsum(x, dim):
    return x + torch.sum(x, dim=dim, keepdim=True) - torch.cumsum(x, dim=dim)


def get_ray_camera_space(img_size, principle, focal, device="cpu"):
    """Legacy camera-space ray grid (pixel corners, no +0.5); documents the sign convention."""
    if isinstance(focal, (int, float)):
        focal = torch.ones(2, device=device) * focal
    elif focal.dim() == 2:
        focal = focal.squeeze()
    elif focal.shape[0] == 1:
        focal = focal.expand(2)
    h, w = img_size
    i, j = torch.meshgrid(torch.linspace(0, h - 1, h), torch.linspace(0, w - 1, w), indexing="ij")
    i, j = i.to(device), j.to(device)
    dirs = torch.stack([-(j - principle[1]) / focal[1], -(i - principle[0]) / focal[0], torch.ones_like(i)], -1)
    return F.normalize(dirs, p=2, dim=2)


def get_cross_activation(sel_length, sel_dsd):
    """c[r, m, k] = (len_m - len_k) * sqrt(dsd_k + 1e-10)   ([R,K] , [R,K] -> [R,K,K])."""
    r, k = sel_length.shape
    return (sel_length.unsqueeze(2) - sel_length.unsqueeze(1)) * (sel_dsd.view(r, 1, k) + 1e-10).pow(.5)


def assign2weight(sel_activation, cross_activation, occupation_weight=1.):
    """w_m = exp(-occ * sum_k exp(-act_k) (erf(c_mk)+1)/2) * exp(-act_m) / exp(-0.5)."""
    density = torch.exp(-sel_activation.unsqueeze(1)) * ((torch.erf(cross_activation) + 1) / 2)
    transmittance = torch.exp(-(torch.sum(density, dim=2)) * occupation_weight)
    return transmittance * torch.exp(-sel_activation) / math.exp(-0.5)


class _Aggregation(torch.autograd.Function):
    """Fused K x K blend weights; analytic backward (the reference differentiates ~15 torch ops)."""

    @staticmethod
    def forward(ctx, sel_idx, sel_act, sel_len, sel_dsd, occupation_weight):
        weight, valid_num = _C.aggregation_forward(sel_idx, sel_act, sel_len, sel_dsd, occupation_weight)
        ctx.save_for_backward(sel_act, sel_len, sel_dsd)
        ctx.occ = float(occupation_weight)
        ctx.mark_non_differentiable(valid_num)
        return weight, valid_num

    @staticmethod
    def backward(ctx, grad_weight, _grad_valid):
        sel_act, sel_len, sel_dsd = ctx.saved_tensors
        g_act, g_len, g_dsd = _C.aggregation_backward(sel_act, sel_len, sel_dsd, grad_weight.contiguous(), ctx.occ)
        return None, g_act, g_len, g_dsd, None


def aggregation(sel_idx: torch.Tensor, sel_act: torch.Tensor, sel_len: torch.Tensor, sel_dsd: torch.Tensor,
                occupation_weight: float = 1.):
    """[..., K] hit lists -> (weight f32 [..., K], sel_idx, valid_num i64 [...], sel_len).
    Same return tuple as the reference (:107)."""
    weight, valid_num = _Aggregation.apply(sel_idx, sel_act, sel_len, sel_dsd, float(occupation_weight))
    return weight, sel_idx, valid_num, sel_len


class _MergeFinal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vert_attr, weight, vert_assign, valid_num, background, mask_thr, idx_mod, zero_padding=False):
        # attribute rows padded to 16 bytes once per call: the kernels gather one row per (pixel, k)
        attr4 = _C.pad_attr4(vert_attr) if (vert_attr.is_cuda and vert_attr.dim() == 2 and vert_attr.shape[1] <= 4
                                            and vert_attr.dtype == torch.float32) else None
        out = _C.merge_final_forward(vert_attr, weight, vert_assign, valid_num, background, mask_thr, idx_mod,
                                     attr4=attr4, zero_padding=zero_padding)
        ctx.save_for_backward(vert_attr, weight, vert_assign, valid_num, background, out)
        ctx.mask_thr, ctx.idx_mod, ctx.attr4 = mask_thr, idx_mod, attr4
        return out

    @staticmethod
    def backward(ctx, grad_out):
        vert_attr, weight, vert_assign, valid_num, background, out = ctx.saved_tensors
        g_attr, g_w = _C.merge_final_backward(vert_attr, weight, vert_assign, valid_num, grad_out.contiguous(),
                                              background, ctx.mask_thr, ctx.idx_mod,
                                              need_attr=ctx.needs_input_grad[0], need_weight=ctx.needs_input_grad[1],
                                              attr4=ctx.attr4, out=out)
        g_bg = None
        if background is not None and ctx.needs_input_grad[4]:
            # learnable background (rare): d out / d bg_c = (1 - mask) where min(x, 1) passes the gradient; the
            # un-clamped composite x is rebuilt with one more gather-blend launch
            with torch.no_grad():
                rgb = _C.merge_final_forward(vert_attr, weight, vert_assign, valid_num, None, -1.0, ctx.idx_mod,
                                             attr4=ctx.attr4)
                sil = weight.sum(-1).clamp(max=1.0)
                mask = (sil > ctx.mask_thr).to(sil.dtype) if ctx.mask_thr > 0 else sil
                x = rgb + (1 - mask).unsqueeze(-1) * background
                passes = (x < 1).to(x.dtype) + 0.5 * (x == 1).to(x.dtype)       # torch.min subgradient at the tie
                g_bg = (grad_out * passes * (1 - mask).unsqueeze(-1)).reshape(-1, x.shape[-1]).sum(0)
        return g_attr, g_w, None, None, g_bg, None, None, None


def _attr_rows(vert_attr, vert_assign):
    if vert_attr.dim() != 2:
        raise AssertionError("vert_attr must be [n, d]")
    return vert_attr


def merge_final(vert_attr: torch.Tensor, weight: torch.Tensor, vert_assign: torch.Tensor, valid_num: torch.Tensor,
                background=None, mask_thr: float = -1.0, idx_mod: int = 0, fused_src=None):
    """out[..., :] = sum_{k < valid_num} weight[..., k] * vert_attr[vert_assign[..., k], :].

    Reference :111-141.  Like the reference this permanently rewrites `vert_assign` -1 -> 0 in place
    (:131).  `background` / `mask_thr` optionally fuse the composite of Renderer.to_colored_background;
    `idx_mod=N` folds packed multi-view indices b*N+n onto an (N, d) attribute table (extension: the
    reference asserts out for B > 1).  `fused_src`: provenance of fragments produced by the fused renderer path
    (Fragments._fused_src); when given and applicable the backward is folded into the renderer's."""
    vert_attr = _attr_rows(vert_attr, vert_assign)
    if STRICT_INDEX_ASSERT:
        limit = vert_attr.shape[0] if idx_mod <= 0 else None
        if limit is not None:
            assert limit > vert_assign.max()
    # reference :131 rewrites vert_assign -1 -> 0 in place.  Fragments straight from the fused renderer are known to
    # hold exactly the -1 padding behind valid_num: the gather-blend kernel then stores the zeros itself (no
    # read-modify-write pass over the (R,K) tensor); everything else takes the generic in-place clamp.
    pad_in_kernel = (fused_src is not None and fused_src.matches(weight, vert_assign, valid_num) and vert_attr.is_cuda
                     and vert_attr.dtype == torch.float32 and vert_attr.shape[1] <= 4 and vert_assign.is_contiguous())
    if not pad_in_kernel:
        with torch.no_grad():
            vert_assign.clamp_(min=0)   # done before the tensor is saved for backward
    # fragments straight from the fused renderer (`fused_src`, set by GaussianRenderer): one autograd node whose
    # backward is the renderer's fused backward with this gather-blend's backward folded in (voge_b200/fused.py)
    from . import fused as _fused
    if _fused.image_fusion_applies(fused_src, vert_attr, weight, vert_assign, valid_num, background, idx_mod):
        return _fused.render_image(fused_src, vert_attr, background, mask_thr, idx_mod, zero_padding=pad_in_kernel)
    return _MergeFinal.apply(vert_attr, weight, vert_assign, valid_num, background, float(mask_thr), int(idx_mod),
                             bool(pad_in_kernel))


def expend_sigma(sigma, rotation_matrix=None):
    """(N,) -> sigma*I ; (N,3) -> per-row scale of the rotation (identity -> diag) ; (N,3,3) as is."""
    if sigma.dim() == 3:
        if sigma.shape[1] == 3 and sigma.shape[2] == 3:
            return sigma
        raise Exception('Got unexpected sigma, which has shape: ' + str(sigma.shape))
    if rotation_matrix is None:
        rotation_matrix = torch.eye(3, device=sigma.device).unsqueeze(0)
    if rotation_matrix.dim() == 2:
        rotation_matrix = rotation_matrix.unsqueeze(0)
    rotation_matrix = rotation_matrix[:, :3, :3]
    if sigma.dim() == 1:
        return sigma.view(-1, 1, 1) * rotation_matrix
    if sigma.dim() == 2:
        return sigma.unsqueeze(2) * rotation_matrix
    raise Exception('Got unexpected sigma, which has shape: ' + str(sigma.shape))

"""Fused renderer path: one autograd Function for GaussianRenderer.forward's whole chain
(reference Renderer.py:124-150: ray generation -> camera-centred copies -> ray_tracing -> aggregation).

Forward : voge_pack_gaussians (S = 2 sigma | 2 inverse(sigma) | 2 L L^T) -> per-view tile culling
          (voge_bin_count / voge_bin_fill) -> voge_trace_hits -> voge_select_topk -> voge_blend_weights
Backward: voge_render_backward_fused -- recompute the hits, analytic blend backward, chain rule straight
          into packed per-Gaussian gradients for all views of the batch -> voge_unpack_gradients (chain rule of
          the sigma parameterisation), plus d/d(origins) and d/d(camera record) = d/dR, d/dfocal, d/dprincipal
          reduced per view inside the kernel when the camera requires grad (the reference materialises
          (B*N,3), (B*N,3,3) and (B,H,W,3) gradient tensors and differentiates ~15 PyTorch ops over (R,K,K)).

For the built-in closed-form camera the (B,H,W,3) rays are never materialised: every kernel generates the ray
of its pixel from the (B,16) camera records (csrc/render_core.cuh: gen_ray).
"""
import math

import torch

from . import _C

TILE_MAX = 16


def choose_tile(bin_size, K, use_ref_bins):
    """Largest tile (<= 16 px, dividing bin_size when the reference bins apply)."""
    import os
    tmax = int(os.environ.get("VOGE_TILE_MAX", TILE_MAX))
    cands = [t for t in range(tmax, 3, -1) if (not use_ref_bins) or bin_size % t == 0]
    if not cands:
        cands = [t for t in range(3, 0, -1) if bin_size % t == 0]
    return cands[0]


class _RenderFused(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, sigmas, origins, rays, cam, R, T, focal, principal, image_size, thr, absorptivity, K,
                use_ref_bins, bin_size, sigma_mode, holder=None):
        thr_act = -math.log(thr + 1e-10)                       # RayTracing.py:85
        tile = choose_tile(bin_size, K, use_ref_bins)
        # (N, 4|8|12) aligned records shared by binning, forward and backward
        # ((N,3,3) sigmas: isotropic Gaussians are stored in the first 16 bytes of their record, see _C.pack_gaussians)
        gauss = _C.pack_gaussians(verts, sigmas, sigma_mode, iso_encode=True)

        def run(gauss, speculate):
            offsets, tile_list, rects, item_offsets = _C.bin_views(None, None, R, T, origins, focal, principal, image_size,
                                                                   thr, thr_act, use_ref_bins, bin_size, tile, gauss=gauss,
                                                                   speculate=speculate)
            if getattr(gauss, "iso_bad", False):
                return None, item_offsets
            out = _C.render_forward(None, None, origins, rays, offsets, tile_list, rects, thr_act, absorptivity, K, tile,
                                    need_act=False, item_offsets=item_offsets, gauss=gauss, cam=cam, image_size=image_size)
            return out, item_offsets

        # Speculative pass: scratch sized from the previous call of this shape, every launch queued without a host
        # round trip, the true totals checked afterwards (_C.BinPlan); the first call of a shape, a scene that outgrew
        # the slack and an ambiguous isotropic encoding take the exact pass (one host sync after voge_bin_count)
        out, item_offsets = run(gauss, True)
        if out is None or not _C.bins_valid(item_offsets):
            out = None
            if getattr(gauss, "iso_bad", False):
                # a non positive definite record makes the encoding ambiguous: plain records, binned again
                gauss = _C.pack_gaussians(verts, sigmas, sigma_mode)
            out, item_offsets = run(gauss, False)
        if holder is not None:
            holder["gauss"] = gauss
        idx, weight, tlen, valid, _, _ = out
        if any(ctx.needs_input_grad[:5]):
            # recompute-not-store: only the inputs and the returned weights (which the caller's Fragments keep
            # alive anyway) are saved; the backward re-evaluates the K hits per pixel from idx (the reference
            # saves mus, isigmas (B*N copies), rays, sel_idx and the PyTorch aggregation ~10 (R,K,K) tensors)
            ctx.save_for_backward(verts, sigmas, origins, rays, cam, weight)
            # idx / valid are handed out as Fragments.vert_index / valid_num and merge_final rewrites
            # vert_index in place (-1 -> 0, reference Aggregation.py:131; the reference clones the
            # tensor for that reason, Renderer.py:145).  They are kept outside autograd's version
            # tracking instead of cloned: the backward only reads the first valid_num slots.
            ctx.idx, ctx.valid, ctx.gauss = idx, valid, gauss
        ctx.absorptivity = float(absorptivity)
        ctx.sigma_mode = int(sigma_mode)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(idx, valid)
        return weight, idx, valid, tlen

    @staticmethod
    def backward(ctx, g_weight, _g_idx, _g_valid, g_len_out):
        verts, sigmas, origins, rays, cam, weight = ctx.saved_tensors
        if g_weight is None:
            g_weight = torch.zeros(ctx.idx.shape, dtype=torch.float32, device=ctx.idx.device)
        # camera gradients (pose optimisation): d/d(origins) and d/d(rays) or d/d(cam record) come out of the same
        # kernel; the tiny (B,.) tensors flow on to R, T, focal through plain autograd
        g_verts, g_sig, g_rays, g_org, g_cam = _C.render_backward_fused(
            verts, sigmas, origins, rays, ctx.idx, ctx.valid, g_weight.contiguous(), g_len_out, ctx.absorptivity,
            need_sigma=ctx.needs_input_grad[1], need_rays=ctx.needs_input_grad[3], need_origins=ctx.needs_input_grad[2],
            gauss=ctx.gauss, weight=weight, cam=cam, need_cam=ctx.needs_input_grad[4], sigma_mode=ctx.sigma_mode)
        return (g_verts, g_sig, g_org, g_rays, g_cam) + (None,) * 12


class FusedSource(object):
    """Provenance of fragments produced by the fused path: what voge_render_backward_image needs to differentiate
    an image composited from them straight down to the Gaussian parameters (merge_final's backward folded into
    the renderer's backward).  Attached to Fragments as `_fused_src`; it is honoured only while the fragment
    tensors are the very objects the renderer returned."""
    __slots__ = ("verts", "sigmas", "origins", "rays", "cam", "gauss", "sigma_mode", "absorptivity", "weight", "idx",
                 "valid", "n_points", "K", "attr_stamp")

    def matches(self, weight, idx, valid):
        return weight is self.weight and idx is self.idx and valid is self.valid


def render_fused(verts, sigmas, origins, rays, cam, R, T, focal, principal, image_size, thr, absorptivity, K,
                 use_ref_bins, bin_size, sigma_mode=0, with_source=False):
    """-> (vert_weight, vert_index (packed, -1 padded), valid_num i64, vert_hit_length) [, FusedSource].
    rays (B,H,W,3) or None (generated in the kernels from cam (B,16) = _C.make_cam(R, focal, principal))."""
    holder = {}
    weight, idx, valid, tlen = _RenderFused.apply(verts, sigmas, origins, rays, cam, R, T, focal, principal,
                                                  tuple(image_size), float(thr), float(absorptivity), int(K),
                                                  bool(use_ref_bins), int(bin_size), int(sigma_mode), holder)
    if not with_source:
        return weight, idx, valid, tlen
    src = FusedSource()
    src.verts, src.sigmas, src.origins, src.rays, src.cam = verts, sigmas, origins, rays, cam
    src.gauss, src.sigma_mode, src.absorptivity = holder.get("gauss"), int(sigma_mode), float(absorptivity)
    src.weight, src.idx, src.valid, src.n_points, src.K = weight, idx, valid, int(verts.shape[0]), int(K)
    return weight, idx, valid, tlen, src


class _RenderImage(torch.autograd.Function):
    """out = merge_final(attr, fragments) (+ background composite) for fragments of the fused path, differentiated
    down to the Gaussian parameters in ONE kernel: voge_render_backward_image forms dL/dw in registers from the
    image gradient, reduces dL/d(attr), and continues with the blend / geometry backward -- the (B,H,W,K)
    weight gradient never exists in HBM and merge_final's own backward launch disappears."""

    @staticmethod
    def forward(ctx, attr, verts, sigmas, origins, rays, cam, background, src, mask_thr, idx_mod, zero_padding=False):
        # kind-9 record tables keep the attribute rows in the records' second 16 bytes (one request per hit fetches
        # geometry + attribute in the backward); the stamp tells the backward whether they are still THIS call's rows
        attr4 = _C.pad_attr4(attr, gauss=src.gauss)
        src.attr_stamp = getattr(src, "attr_stamp", 0) + 1
        ctx.attr_stamp = src.attr_stamp
        out, code = _C.merge_final_forward(attr, src.weight.detach(), src.idx, src.valid, background, mask_thr, idx_mod,
                                           attr4=attr4, want_sat_code=True, zero_padding=zero_padding)
        ctx.save_for_backward(verts, sigmas, origins, rays, cam, attr4, out, background, code)
        ctx.src, ctx.mask_thr, ctx.C = src, float(mask_thr), int(attr.shape[1])
        return out

    @staticmethod
    def backward(ctx, grad_out):
        verts, sigmas, origins, rays, cam, attr4, out, background, code = ctx.saved_tensors
        src = ctx.src
        need = ctx.needs_input_grad
        g_verts, g_sig, g_attr, g_rays, g_org, g_cam = _C.render_backward_image(
            verts, sigmas, origins, rays, src.idx, src.valid, src.weight.detach(), grad_out.contiguous(), out, attr4,
            background, ctx.mask_thr, src.absorptivity, sat_code=code, need_sigma=need[2], need_attr=need[0], need_rays=need[4],
            need_origins=need[3], gauss=src.gauss, cam=cam, need_cam=need[5], sigma_mode=src.sigma_mode,
            n_channels=ctx.C, attr_in_records=(ctx.attr_stamp == src.attr_stamp))
        return g_attr, g_verts, g_sig, g_org, g_rays, g_cam, None, None, None, None, None


def image_fusion_applies(src, vert_attr, weight, vert_assign, valid_num, background, idx_mod):
    import os
    if src is None or os.environ.get("VOGE_NO_IMAGE_FUSION") == "1" or not torch.is_grad_enabled():
        return False
    if not (src.matches(weight, vert_assign, valid_num) and weight.requires_grad and src.gauss is not None):
        return False
    if not (vert_attr.is_cuda and vert_attr.dtype == torch.float32 and vert_attr.dim() == 2 and
            vert_attr.shape[0] == src.n_points and 1 <= vert_attr.shape[1] <= 4 and src.K <= 112):
        return False
    if background is not None and background.requires_grad:
        return False
    n_views = weight.shape[0] if weight.dim() == 4 else 1
    return idx_mod == src.n_points or (n_views == 1 and idx_mod == 0)


def render_image(src, vert_attr, background, mask_thr, idx_mod, zero_padding=False):
    return _RenderImage.apply(vert_attr, src.verts, src.sigmas, src.origins, src.rays, src.cam, background, src,
                              float(mask_thr), int(idx_mod), bool(zero_padding))


class _GenerateRays(torch.autograd.Function):
    """Materialised rays of the closed-form camera for the op-by-op entry points: forward = voge_generate_rays
    (the same device function the fused kernels evaluate per pixel, so both paths trace identical rays);
    backward = the generator's chain rule (d = R normalize(a, b, 1), a = (px - .5 - x) / fx) in plain torch."""

    @staticmethod
    def forward(ctx, cam, image_size):
        ctx.save_for_backward(cam)
        ctx.image_size = image_size
        return _C.generate_rays(cam, image_size)

    @staticmethod
    def backward(ctx, g):
        (cam,) = ctx.saved_tensors
        H, W = ctx.image_size
        B = cam.shape[0]
        Rm = cam[:, :9].reshape(B, 3, 3)
        fx, fy, px, py = (cam[:, i].view(B, 1, 1) for i in (9, 10, 11, 12))
        xs = torch.arange(W, dtype=torch.float32, device=cam.device).view(1, 1, W)
        ys = torch.arange(H, dtype=torch.float32, device=cam.device).view(1, H, 1)
        a = ((px - 0.5 - xs) / fx).expand(B, H, W)
        b = ((py - 0.5 - ys) / fy).expand(B, H, W)
        inv = torch.rsqrt(a * a + b * b + 1)
        dc = torch.stack([a * inv, b * inv, inv], dim=-1)                       # (B,H,W,3)
        g_R = torch.einsum('bhwi,bhwj->bij', g, dc)
        h = torch.einsum('bij,bhwi->bhwj', Rm, g)                               # R^T g
        gv = (h - dc * (dc * h).sum(-1, keepdim=True)) * inv.unsqueeze(-1)
        ga, gb = gv[..., 0], gv[..., 1]
        g_cam = torch.zeros_like(cam)
        g_cam[:, :9] = g_R.reshape(B, 9)
        g_cam[:, 9] = -(ga * a).sum((1, 2)) / fx.view(B)
        g_cam[:, 10] = -(gb * b).sum((1, 2)) / fy.view(B)
        g_cam[:, 11] = ga.sum((1, 2)) / fx.view(B)
        g_cam[:, 12] = gb.sum((1, 2)) / fy.view(B)
        return g_cam, None


def generate_rays(cam, image_size):
    """(B,H,W,3) unit ray directions of the closed-form camera records cam (B,16); differentiable."""
    return _GenerateRays.apply(cam, (int(image_size[0]), int(image_size[1])))

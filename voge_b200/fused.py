"""Fused renderer path: one autograd Function for GaussianRenderer.forward's whole chain
(reference Renderer.py:130-150: camera-centred copies -> ray_tracing -> aggregation).

Forward : per-view tile culling (voge_bin_count / voge_bin_fill) -> voge_render_forward
Backward: voge_render_backward_fused -- recompute the hits, analytic blend backward, chain rule straight
          into (N,3) / compact-sigma gradients for all views of the batch, plus d/d(rays), d/d(origins) when
          the camera requires grad (the reference materialises
          (B*N,3) and (B*N,3,3) gradient tensors and differentiates ~15 PyTorch ops over (R,K,K)).
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
    def forward(ctx, verts, sigmas, origins, rays, R, T, focal, principal, image_size, thr, absorptivity, K,
                use_ref_bins, bin_size):
        thr_act = -math.log(thr + 1e-10)                       # RayTracing.py:85
        tile = choose_tile(bin_size, K, use_ref_bins)
        offsets, tile_list, rects, item_offsets = _C.bin_views(verts, sigmas, R, T, origins, focal, principal,
                                                               image_size, thr, thr_act, use_ref_bins, bin_size, tile)
        gauss = _C.pack_gaussians(verts, sigmas)     # (N, 4|8|12) aligned records shared by forward and backward
        idx, weight, tlen, valid, _, _ = _C.render_forward(verts, sigmas, origins, rays, offsets, tile_list, rects,
                                                           thr_act, absorptivity, K, tile, need_act=False,
                                                           item_offsets=item_offsets, gauss=gauss)
        if verts.requires_grad or sigmas.requires_grad or origins.requires_grad or rays.requires_grad:
            # recompute-not-store: only the inputs and the returned weights (which the caller's Fragments keep
            # alive anyway) are saved; the backward re-evaluates the K hits per pixel from idx (the reference
            # saves mus, isigmas (B*N copies), rays, sel_idx and the PyTorch aggregation ~10 (R,K,K) tensors)
            ctx.save_for_backward(verts, sigmas, origins, rays, weight)
            # idx / valid are handed out as Fragments.vert_index / valid_num and merge_final rewrites
            # vert_index in place (-1 -> 0, reference Aggregation.py:131; the reference clones the
            # tensor for that reason, Renderer.py:145).  They are kept outside autograd's version
            # tracking instead of cloned: the backward only reads the first valid_num slots.
            ctx.idx, ctx.valid, ctx.gauss = idx, valid, gauss
        ctx.absorptivity = float(absorptivity)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(idx, valid)
        return weight, idx, valid, tlen

    @staticmethod
    def backward(ctx, g_weight, _g_idx, _g_valid, g_len_out):
        verts, sigmas, origins, rays, weight = ctx.saved_tensors
        if g_weight is None:
            g_weight = torch.zeros(ctx.idx.shape, dtype=torch.float32, device=ctx.idx.device)
        # camera gradients (pose optimisation): d/d(origins), d/d(rays) come out of the same kernel and flow
        # on to R, T, focal through the ray generator's autograd graph (voge_b200/cameras.py)
        g_verts, g_sig, g_rays, g_org = _C.render_backward_fused(
            verts, sigmas, origins, rays, ctx.idx, ctx.valid, g_weight.contiguous(), g_len_out, ctx.absorptivity,
            need_sigma=ctx.needs_input_grad[1], need_rays=ctx.needs_input_grad[3], need_origins=ctx.needs_input_grad[2],
            gauss=ctx.gauss, weight=weight)
        return (g_verts, g_sig, g_org, g_rays) + (None,) * 10


def render_fused(verts, sigmas, origins, rays, R, T, focal, principal, image_size, thr, absorptivity, K,
                 use_ref_bins, bin_size):
    """-> (vert_weight, vert_index (packed, -1 padded), valid_num i64, vert_hit_length)."""
    return _RenderFused.apply(verts, sigmas, origins, rays, R, T, focal, principal, tuple(image_size), float(thr),
                              float(absorptivity), int(K), bool(use_ref_bins), int(bin_size))

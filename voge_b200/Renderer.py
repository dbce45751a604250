"""Public renderer API (reference VoGE/Renderer.py): Fragments :13-50, GaussianRenderSettings :53-84,
GaussianRenderer :87-150, interpolate_attr :153, get_silhouette :157-159, to_colored_background
:162-171, to_white_background :174-176 -- same names, signatures, return types and dtypes."""
from typing import Tuple, Union

import torch
import torch.nn as nn

from .Aggregation import aggregation, expend_sigma, merge_final
from .RayTracing import ray_tracing
from .cameras import camera_params, generate_rays
from .fused import render_fused
from .RayTracing import default_bin_size


class Fragments(object):
    """Per-pixel top-K hit lists: vert_weight (B,H,W,K) f32, vert_index (B,H,W,K) i32 (-1 = empty,
    packed b*N+n for B > 1), valid_num (B,H,W) i64, vert_hit_length (B,H,W,K) f32 (1e10 = empty)."""

    def __init__(self, vert_weight, vert_index, valid_num, vert_hit_length, points_per_view=None):
        self.vert_weight = vert_weight
        self.vert_index = vert_index
        self.valid_num = valid_num
        self.vert_hit_length = vert_hit_length
        # extension: number of Gaussians per view, set by GaussianRenderer.  With several views the
        # indices are packed b*N+n (as in the reference, whose merge_final then asserts out); knowing N
        # lets interpolate_attr fold them onto an (N, d) attribute table.
        self.points_per_view = points_per_view

    def _map(self, fn):
        return Fragments(vert_weight=fn(self.vert_weight), vert_index=fn(self.vert_index),
                         valid_num=fn(self.valid_num), vert_hit_length=fn(self.vert_hit_length),
                         points_per_view=self.points_per_view)

    def __getitem__(self, item):
        assert len(self.valid_num.shape) == 3, 'Index access is only available when batched.'
        return self._map(lambda t: t[item])

    def __len__(self):
        return self.valid_num.shape[0]

    @property
    def shape(self):
        return (self.vert_weight.shape, self.vert_index.shape, self.valid_num.shape, self.vert_hit_length.shape)

    def squeeze(self):
        assert self.valid_num.shape[0] == 1
        return self[0]

    def unsqueeze(self):
        assert len(self.valid_num.shape) == 2
        return self._map(lambda t: t.unsqueeze(0))

    def to_dict(self):
        return dict(vert_weight=self.vert_weight, vert_index=self.vert_index, valid_num=self.valid_num,
                    vert_hit_length=self.vert_hit_length)

    def copy(self):
        # like the reference this aliases contiguous tensors (.contiguous() is a no-op for them)
        return self._map(lambda t: t.contiguous())


class GaussianRenderSettings:
    __slots__ = ['image_size', 'max_assign', 'thr_activation', 'absorptivity', 'inverse_sigma', 'principal',
                 'max_point_per_bin']

    def __init__(self, image_size: Union[int, Tuple[int, int]] = 256, max_assign: int = 20,
                 thr_activation: float = 0.01, absorptivity: float = 1, inverse_sigma: bool = False,
                 principal=None, max_point_per_bin: Union[None, int] = None, **kwargs):
        # unknown keyword arguments are accepted and ignored, as in the reference (demos pass
        # batch_size=, principal_point=, ...)
        self.image_size = (image_size, image_size) if isinstance(image_size, int) else image_size
        self.max_assign = max_assign
        self.thr_activation = thr_activation
        self.absorptivity = absorptivity
        self.inverse_sigma = inverse_sigma
        self.principal = principal
        self.max_point_per_bin = max_point_per_bin

    def __getitem__(self, item):
        return getattr(self, item)


class GaussianRenderer(nn.Module):
    to_set_args = ['R', 'T', 'focal', 'principal']

    def __init__(self, cameras, render_settings: Union[dict, GaussianRenderSettings]):
        super().__init__()
        self.cameras = cameras
        self.render_settings = render_settings
        self.device = cameras.device

    def to(self, device):
        self.cameras = self.cameras.to(device)  # cameras are not an nn.Module
        self.device = device
        return self

    # The fused CUDA path (voge_b200/fused.py) is used whenever it applies; set False to force the
    # op-by-op chain ray_tracing -> aggregation (same results, reference-shaped intermediates).
    use_fused = True

    def _fusable(self, verts, sigmas, rays, origins):
        if not (verts.is_cuda and verts.dim() == 3 and verts.shape[0] == 1 and verts.dtype == torch.float32):
            return False
        return sigmas.dim() in (1, 2, 3)

    def _forward_fused(self, verts, sigmas, rays, origins):
        st = self.render_settings
        map_size = st['image_size']
        sig = sigmas
        if st['inverse_sigma']:
            sig = torch.inverse(expend_sigma(sigmas))
        R, T, focal, principal = camera_params(self.cameras, map_size)
        n_views = rays.shape[0]
        R, T = R.expand(n_views, -1, -1), T.expand(n_views, -1)
        focal, principal = focal.expand(n_views, -1), principal.expand(n_views, -1)
        M = st['max_point_per_bin']
        w, idx, valid, ln = render_fused(verts[0], sig, origins, rays, R, T, focal, principal, map_size,
                                         st['thr_activation'], st['absorptivity'], st['max_assign'],
                                         use_ref_bins=(M != -1), bin_size=default_bin_size(map_size))
        return Fragments(vert_weight=w, vert_index=idx, valid_num=valid, vert_hit_length=ln,
                         points_per_view=verts.shape[1])

    def _rays(self, image_size):
        """(directions (B,H,W,3), origins (B,3)).  Real pytorch3d cameras go through pytorch3d's own
        ray sampler exactly like the reference (:124-128); the built-in camera uses the closed form."""
        cams = self.cameras
        if type(cams).__module__.startswith('pytorch3d'):
            from pytorch3d.renderer.implicit.raysampling import NDCMultinomialRaysampler
            sampler = NDCMultinomialRaysampler(image_width=int(image_size[1]), image_height=int(image_size[0]),
                                               unit_directions=True, n_pts_per_ray=1, min_depth=0, max_depth=10)
            bundle = sampler(cams)
            return bundle.directions, bundle.origins[:, 0, 0, :]
        # The rays depend only on the camera tensors: regenerate them (a dozen elementwise passes over
        # (B,H,W,3)) only when one of those tensors was replaced or written to, or when gradients must
        # flow to the camera.  The reference rebuilds them on every forward.
        tensors = [t for t in (cams.R, cams.T, cams.focal_length, cams.principal_point) if torch.is_tensor(t)]
        if any(t.requires_grad for t in tensors):
            return generate_rays(cams, image_size)
        key = (tuple(image_size),) + tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors)
        cache = getattr(self, '_ray_cache', None)
        if cache is None or cache[0] != key:
            with torch.no_grad():
                cache = (key,) + tuple(generate_rays(cams, image_size))
            self._ray_cache = cache
        return cache[1], cache[2]

    def forward(self, gmeshes, **kwargs):
        assert not self.cameras.in_ndc(), 'Got NDC camera. Cameras.in_ndc must be set to false.'
        for name, value in kwargs.items():
            if name in self.to_set_args:
                setattr(self.cameras, name, value.to(self.device) if isinstance(value, torch.Tensor) else value)

        st = self.render_settings
        verts, sigmas, _radians = gmeshes()
        map_size = st['image_size']
        if verts.dim() == 2:
            verts = verts[None]

        rays, ray_origins = self._rays(map_size)
        if self.use_fused and self._fusable(verts, sigmas, rays, ray_origins):
            return self._forward_fused(verts, sigmas, rays, ray_origins)
        sigmas = expend_sigma(sigmas)
        verts_transformed = verts - ray_origins[:, None]
        if sigmas.dim() == 3:
            sigmas = sigmas.unsqueeze(0).expand(verts_transformed.shape[0], -1, -1, -1)
        isigma = 2 * torch.inverse(sigmas) if st['inverse_sigma'] else 2 * sigmas

        sel_idx, sel_len, sel_act, sel_dsd = ray_tracing(
            self.cameras, verts_transformed, isigma, rays, map_size, thr=st['thr_activation'],
            n_assign=st['max_assign'], max_points_per_bin=st['max_point_per_bin'])
        # the tensor saved for the ray-tracing backward keeps its -1 markers: merge_final rewrites
        # Fragments.vert_index in place (reference :145)
        sel_idx = sel_idx.clone()
        vert_weight, vert_index, valid_num, vert_hit_length = aggregation(
            sel_idx=sel_idx, sel_act=sel_act, sel_len=sel_len, sel_dsd=sel_dsd,
            occupation_weight=st['absorptivity'])
        return Fragments(vert_weight=vert_weight, vert_index=vert_index, valid_num=valid_num,
                         vert_hit_length=vert_hit_length, points_per_view=verts.shape[1])


def _idx_mod(fragments, vert_attr):
    n = getattr(fragments, 'points_per_view', None)
    return int(n) if (n is not None and vert_attr.shape[0] == n) else 0


def interpolate_attr(fragments: Fragments, vert_attr: torch.Tensor):
    return merge_final(vert_attr=vert_attr, weight=fragments.vert_weight, valid_num=fragments.valid_num,
                       vert_assign=fragments.vert_index, idx_mod=_idx_mod(fragments, vert_attr))


def get_silhouette(fragments: Fragments):
    merged = fragments.vert_weight.sum(-1)
    return torch.min(merged, torch.ones_like(merged))


def to_colored_background(fragments: Fragments, colors: torch.Tensor,
                          background_color: Union[torch.Tensor, tuple, list] = (1, 1, 1), thr: float = -1):
    """min(sum_k w_k colour[idx_k] + (1 - silhouette) * background, 1) -- one fused gather-blend kernel
    (the reference chains get_silhouette, interpolate_attr and four elementwise ops, :162-171)."""
    if not torch.is_tensor(background_color):
        background_color = torch.tensor(list(background_color), dtype=torch.float32)
    background_color = background_color.to(device=colors.device, dtype=torch.float32).reshape(-1)
    if background_color.numel() == 1:
        background_color = background_color.expand(colors.shape[-1])
    return merge_final(vert_attr=colors, weight=fragments.vert_weight, valid_num=fragments.valid_num,
                       vert_assign=fragments.vert_index, background=background_color.contiguous(), mask_thr=thr,
                       idx_mod=_idx_mod(fragments, colors))


def to_white_background(fragments: Fragments, colors: torch.Tensor, thr: float = -1):
    return to_colored_background(fragments=fragments, colors=colors, background_color=(1, 1, 1), thr=thr)

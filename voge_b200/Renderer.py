"""Public renderer API (reference VoGE/Renderer.py): Fragments :13-50, GaussianRenderSettings :53-84,
GaussianRenderer :87-150, interpolate_attr :153, get_silhouette :157-159, to_colored_background
:162-171, to_white_background :174-176 -- same names, signatures, return types and dtypes."""
from typing import Tuple, Union

import torch
import torch.nn as nn

from . import _C
from .Aggregation import aggregation, expend_sigma, merge_final
from .RayTracing import default_bin_size, ray_tracing
from .cameras import PerspectiveCameras, camera_params, generate_rays
from .fused import generate_rays as fused_generate_rays
from .fused import render_fused


MAX_BACKGROUND_CHANNELS = 32     # kMaxBgChannels of csrc/blend.cu
_BACKGROUND_CACHE = {}


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
        out = Fragments(vert_weight=fn(self.vert_weight), vert_index=fn(self.vert_index),
                        valid_num=fn(self.valid_num), vert_hit_length=fn(self.vert_hit_length),
                        points_per_view=self.points_per_view)
        src = getattr(self, '_fused_src', None)
        if src is not None:      # honoured only if the mapped tensors are still the renderer's own objects (copy())
            out._fused_src = src
        return out

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
                 'max_point_per_bin', 'cholesky_sigma']

    def __init__(self, image_size: Union[int, Tuple[int, int]] = 256, max_assign: int = 20,
                 thr_activation: float = 0.01, absorptivity: float = 1, inverse_sigma: bool = False,
                 principal=None, max_point_per_bin: Union[None, int] = None, cholesky_sigma: bool = False, **kwargs):
        # unknown keyword arguments are accepted and ignored, as in the reference (demos pass
        # batch_size=, principal_point=, ...)
        self.image_size = (image_size, image_size) if isinstance(image_size, int) else image_size
        self.max_assign = max_assign
        self.thr_activation = thr_activation
        self.absorptivity = absorptivity
        self.inverse_sigma = inverse_sigma
        self.principal = principal
        self.max_point_per_bin = max_point_per_bin
        # extension: `sigmas` (N,3,3) are Cholesky factors L, the inverse covariance is tril(L) tril(L)^T -- the
        # `to_sym` parameterisation of demo/EfficientCuboidViaOptimization.py:17-18, evaluated (and differentiated)
        # inside the pack / gradient-epilogue kernels instead of by autograd
        self.cholesky_sigma = cholesky_sigma

    def __getitem__(self, item):
        return getattr(self, item)


def _setting(st, name, default=None):
    try:
        return st[name]
    except (KeyError, AttributeError):
        return default


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
    # Differences under misuse: the fused path has no per-bin candidate cap (max_point_per_bin only selects
    # "reference coarse bins" vs "no coarse stage"); the op-by-op chain keeps the first M candidates of an
    # overflowing bin (the reference drops a non-deterministic chunk, rasterize_coarse.cu:150-163).
    use_fused = True

    def _builtin_camera(self):
        return isinstance(self.cameras, PerspectiveCameras)

    def _fusable(self, verts, sigmas):
        if not (verts.is_cuda and verts.dim() == 3 and verts.shape[0] == 1 and verts.dtype == torch.float32):
            return False
        if sigmas.dtype != torch.float32 or sigmas.dim() not in (1, 2, 3):
            return False
        if _setting(self.render_settings, 'cholesky_sigma', False) and sigmas.dim() != 3:
            return False
        return True

    def _sigma_mode(self):
        st = self.render_settings
        chol, inv = bool(_setting(st, 'cholesky_sigma', False)), bool(st['inverse_sigma'])
        if chol and inv:
            raise ValueError("cholesky_sigma and inverse_sigma are mutually exclusive")
        return 2 if chol else (1 if inv else 0)

    def _camera_tensors(self, image_size):
        """(R (B,3,3), T (B,3), focal (B,2), principal (B,2), origins (B,3)); origins = camera centres -T R^T,
        plain torch ops on (B,.) tensors, differentiable."""
        R, T, focal, principal = camera_params(self.cameras, image_size)
        n = max(R.shape[0], T.shape[0], focal.shape[0], principal.shape[0])
        R, T = R.expand(n, -1, -1), T.expand(n, -1)
        focal, principal = focal.expand(n, -1), principal.expand(n, -1)
        origins = -torch.matmul(T[:, None, :], R.transpose(1, 2))[:, 0, :]
        return R, T, focal, principal, origins

    def _forward_fused(self, verts, sigmas, rays, origins):
        st = self.render_settings
        map_size = st['image_size']
        # The camera-derived tensors ((B,.) only, but a dozen tiny launches) are rebuilt only when a camera tensor was
        # replaced or written to; cameras that require grad always take the differentiable route.
        tensors = self._camera_key_tensors()
        frozen = not any(t.requires_grad for t in tensors)
        key = self._camera_key(map_size) if frozen else None
        cached = getattr(self, '_fused_cam_cache', None)
        if frozen and cached is not None and cached[0] == key:
            R, T, focal, principal, cam_origins, cam_rec = cached[2]
        else:
            R, T, focal, principal, cam_origins = self._camera_tensors(map_size)
            cam_rec = _C.make_cam(R, focal, principal)
            if frozen:
                self._fused_cam_cache = (key, list(tensors), (R, T, focal, principal, cam_origins, cam_rec))
        if rays is None:
            # closed-form camera: rays are generated inside the kernels from the (B,16) camera records
            cam, origins = cam_rec, cam_origins
        else:
            cam = None
        M = st['max_point_per_bin']
        w, idx, valid, ln, src = render_fused(verts[0], sigmas, origins, rays, cam, R, T, focal, principal, map_size,
                                              st['thr_activation'], st['absorptivity'], st['max_assign'],
                                              use_ref_bins=(M != -1), bin_size=default_bin_size(map_size),
                                              sigma_mode=self._sigma_mode(), with_source=True)
        frag = Fragments(vert_weight=w, vert_index=idx, valid_num=valid, vert_hit_length=ln,
                         points_per_view=verts.shape[1])
        frag._fused_src = src          # lets interpolate_attr / to_*_background fold their backward into the renderer's
        return frag

    def _rays_match_model(self, rays, origins, image_size):
        """Foreign camera objects (pytorch3d): the fused path culls with the closed-form pinhole model of
        voge_b200.cameras while the rays come from the camera's own ray sampler.  Compare the two
        at the image corners and centre (and the origins); on any mismatch -- K-matrix cameras, other camera
        classes, a sampler with different conventions -- the op-by-op chain is used instead.  One host sync per
        camera state (cached on the camera tensors' identities and versions)."""
        H, W = int(image_size[0]), int(image_size[1])
        key = self._camera_key(image_size)
        cached = getattr(self, '_model_check', None)
        if cached is not None and cached[0] == key:
            return cached[2]
        ok = False
        try:
            R, T, focal, principal, cam_origins = self._camera_tensors(image_size)
            if rays.shape[0] == R.shape[0] and tuple(rays.shape[1:3]) == (H, W):
                ys = torch.tensor([0, 0, H - 1, H - 1, H // 2], device=rays.device)
                xs = torch.tensor([0, W - 1, 0, W - 1, W // 2], device=rays.device)
                a = (principal[:, 0:1] - 0.5 - xs[None].float()) / focal[:, 0:1]
                b = (principal[:, 1:2] - 0.5 - ys[None].float()) / focal[:, 1:2]
                dc = torch.nn.functional.normalize(torch.stack([a, b, torch.ones_like(a)], -1), dim=-1)   # (B,5,3)
                want = torch.matmul(dc, R.transpose(1, 2))
                got = rays[:, ys, xs]
                err = max(float((got - want).abs().max()), float((origins - cam_origins).abs().max()))
                ok = err < 2e-5
        except Exception:
            ok = False
        self._model_check = (key, [t for t in self._camera_key_tensors()], ok)
        return ok

    def _camera_key_tensors(self):
        cams = self.cameras
        names = ('R', 'T', 'focal_length', 'principal_point')
        return [getattr(cams, n) for n in names if torch.is_tensor(getattr(cams, n, None))]

    def _camera_key(self, image_size):
        # identity (not address) of the camera tensors + version + layout: the cache below keeps references to the
        # keyed tensors, so an id can not be recycled while its entry is alive
        return (tuple(image_size),) + tuple((id(t), t._version, tuple(t.shape), tuple(t.stride()), str(t.device),
                                             t.data_ptr()) for t in self._camera_key_tensors())

    def _rays(self, image_size):
        """(directions (B,H,W,3), origins (B,3)).  Real pytorch3d cameras go through pytorch3d's own
        ray sampler exactly like the reference (:124-128); the built-in camera uses the closed form
        (voge_generate_rays: the generator the fused kernels evaluate per pixel, materialised)."""
        cams = self.cameras
        if not self._builtin_camera():
            from pytorch3d.renderer.implicit.raysampling import NDCMultinomialRaysampler
            sampler = NDCMultinomialRaysampler(image_width=int(image_size[1]), image_height=int(image_size[0]),
                                               unit_directions=True, n_pts_per_ray=1, min_depth=0, max_depth=10)
            bundle = sampler(cams)
            return bundle.directions, bundle.origins[:, 0, 0, :]
        R, T, focal, principal, origins = self._camera_tensors(image_size)
        if not R.is_cuda:
            return generate_rays(cams, image_size)      # host-side transcription (CPU tensors: tests of the glue)
        tensors = self._camera_key_tensors()
        if any(t.requires_grad for t in tensors):
            return fused_generate_rays(_C.make_cam(R, focal, principal), image_size), origins
        # The rays depend only on the camera tensors: regenerate them only when one of those tensors was
        # replaced or written to.  The reference rebuilds them on every forward.
        key = self._camera_key(image_size)
        cache = getattr(self, '_ray_cache', None)
        if cache is None or cache[0] != key:
            with torch.no_grad():
                rays = _C.generate_rays(_C.make_cam(R, focal, principal), image_size)
                cache = (key, list(tensors), rays, origins.detach())
            self._ray_cache = cache
        return cache[2], cache[3]

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

        if self.use_fused and self._fusable(verts, sigmas):
            if self._builtin_camera():
                return self._forward_fused(verts, sigmas, None, None)
            rays, ray_origins = self._rays(map_size)
            if self._rays_match_model(rays, ray_origins, map_size):
                return self._forward_fused(verts, sigmas, rays, ray_origins)
        else:
            rays, ray_origins = self._rays(map_size)
        if _setting(st, 'cholesky_sigma', False):
            low = torch.tril(sigmas)
            sigmas = low @ low.transpose(-2, -1)
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
                       vert_assign=fragments.vert_index, idx_mod=_idx_mod(fragments, vert_attr),
                       fused_src=getattr(fragments, '_fused_src', None))


def get_silhouette(fragments: Fragments):
    merged = fragments.vert_weight.sum(-1)
    return torch.min(merged, torch.ones_like(merged))


def to_colored_background(fragments: Fragments, colors: torch.Tensor,
                          background_color: Union[torch.Tensor, tuple, list] = (1, 1, 1), thr: float = -1):
    """min(sum_k w_k colour[idx_k] + (1 - silhouette) * background, 1) -- one fused gather-blend kernel
    (the reference chains get_silhouette, interpolate_attr and four elementwise ops, :162-171)."""
    C = int(colors.shape[-1])
    if not torch.is_tensor(background_color):
        # constant backgrounds (tuples / lists / scalars) are uploaded once per (value, device): a fresh host -> device
        # copy per call is a blocking cudaMemcpy on the fitting loop's critical path
        key = (tuple(float(v) for v in (background_color if hasattr(background_color, '__len__') else (background_color,))),
               str(colors.device))
        cached = _BACKGROUND_CACHE.get(key)
        if cached is None:
            cached = torch.tensor(key[0], dtype=torch.float32).to(colors.device)
            if len(_BACKGROUND_CACHE) < 64:
                _BACKGROUND_CACHE[key] = cached
        background_color = cached
    background_color = background_color.to(device=colors.device, dtype=torch.float32).reshape(-1)
    if background_color.numel() == 1:
        background_color = background_color.expand(C)
    elif background_color.numel() != C:
        # the reference's `rgb + ones_like(rgb) * (1 - masks) * background` raises the same way (:171)
        raise RuntimeError("The size of tensor a (%d) must match the size of tensor b (%d) at non-singleton "
                           "dimension 3: background_color must have 1 or %d entries" % (C, background_color.numel(), C))
    if C > MAX_BACKGROUND_CHANNELS:
        raise RuntimeError("to_colored_background supports at most %d colour channels (got %d); composite wider "
                           "feature maps with interpolate_attr and get_silhouette" % (MAX_BACKGROUND_CHANNELS, C))
    return merge_final(vert_attr=colors, weight=fragments.vert_weight, valid_num=fragments.valid_num,
                       vert_assign=fragments.vert_index, background=background_color.contiguous(), mask_thr=thr,
                       idx_mod=_idx_mod(fragments, colors), fused_src=getattr(fragments, '_fused_src', None))


def to_white_background(fragments: Fragments, colors: torch.Tensor, thr: float = -1):
    return to_colored_background(fragments=fragments, colors=colors, background_color=(1, 1, 1), thr=thr)

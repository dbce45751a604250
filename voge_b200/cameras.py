"""Minimal pytorch3d-compatible camera layer for the VoGE hot path.

The reference renderer takes a pytorch3d `PerspectiveCameras` and uses it for (a) per-pixel ray
directions / origins via `NDCMultinomialRaysampler` (reference VoGE/Renderer.py:124-128) and
(b) world->view / projection / NDC transforms for the coarse culling (VoGE/RayTracing.py:45-57).
pytorch3d is a third-party dependency that is not vendored by the reference (Readme.md:13 pins
"PyTorch3D 0.6" in prose only) and is not installable here, so this module restates the
published pytorch3d semantics it needs in closed form:

  * row-vector convention  X_view = X_world @ R + T ;  view frame +X left, +Y up, +Z forward
  * screen-space PerspectiveCameras (in_ndc=False): focal_length / principal_point in pixels
  * NDC: shorter image side spans [-1, 1];  x_ndc = fx*X/(Z*s) - (px - W/2)/s,  s = min(H,W)/2
  * projected z = 1 / Z_view
  * rays through pixel centres (xi+.5, yi+.5); origins = camera centre C = -T @ R^T

`GaussianRenderer` accepts either a real pytorch3d camera (if the user has pytorch3d) or
`PerspectiveCameras` below; both expose the attributes/methods the reference touches.
"PARITY UNPINNED" at this boundary: the reference holds no test vector for these transforms.
"""
import math
from typing import Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F


def _as_batch(x, n, width, device, dtype=torch.float32):
    if not torch.is_tensor(x):
        x = torch.tensor(x, dtype=dtype, device=device)
    x = x.to(device=device, dtype=dtype)
    if x.dim() == 0:
        x = x.view(1, 1).expand(1, width)
    if x.dim() == 1:
        x = x.view(1, -1) if x.shape[0] == width else x.view(-1, 1)
    if x.shape[-1] == 1 and width > 1:
        x = x.expand(x.shape[0], width)
    if x.shape[0] == 1 and n > 1:
        x = x.expand(n, *x.shape[1:])
    return x


class Transform3d:
    """Row-vector 4x4 transform stack (points @ matrix), the subset of pytorch3d's API VoGE uses."""

    def __init__(self, matrix: Optional[torch.Tensor] = None, device="cpu", dtype=torch.float32):
        if matrix is None:
            matrix = torch.eye(4, dtype=dtype, device=device)[None]
        if matrix.dim() == 2:
            matrix = matrix[None]
        self._matrix = matrix
        self.device = matrix.device

    def get_matrix(self) -> torch.Tensor:
        return self._matrix

    def compose(self, *others: "Transform3d") -> "Transform3d":
        m = self._matrix
        for o in others:
            m = torch.matmul(m, o.get_matrix())
        return Transform3d(matrix=m)

    def inverse(self) -> "Transform3d":
        return Transform3d(matrix=torch.inverse(self._matrix))

    def transform_points(self, points: torch.Tensor, eps: Optional[float] = None) -> torch.Tensor:
        squeeze = points.dim() == 2
        if squeeze:
            points = points[None]
        ones = torch.ones(points.shape[:-1] + (1,), dtype=points.dtype, device=points.device)
        homo = torch.cat([points, ones], dim=-1)
        out = torch.matmul(homo, self._matrix)
        denom = out[..., 3:]
        if eps is not None:
            denom = denom.sign().clamp(min=0) * 2 - 1
            denom = denom * torch.clamp(out[..., 3:].abs(), eps)
        out = out[..., :3] / denom
        return out[0] if squeeze else out

    def to(self, device):
        return Transform3d(matrix=self._matrix.to(device))


class PerspectiveCameras:
    """Screen-space (in_ndc=False by default here is allowed either way) perspective cameras."""

    def __init__(self, focal_length=1.0, principal_point=((0.0, 0.0),), R=None, T=None, K=None,
                 device="cpu", in_ndc: bool = True, image_size=None):
        self.device = torch.device(device)
        n = 1
        for v in (R, T, focal_length, principal_point):
            if torch.is_tensor(v) and v.dim() >= 2:
                n = max(n, v.shape[0])
        if R is None:
            R = torch.eye(3)[None]
        if T is None:
            T = torch.zeros(1, 3)
        R = R if torch.is_tensor(R) else torch.tensor(R, dtype=torch.float32)
        T = T if torch.is_tensor(T) else torch.tensor(T, dtype=torch.float32)
        if R.dim() == 2:
            R = R[None]
        if T.dim() == 1:
            T = T[None]
        n = max(n, R.shape[0], T.shape[0])
        self._N = n
        self.R = R.to(self.device, torch.float32)
        self.T = T.to(self.device, torch.float32)
        self.focal_length = _as_batch(focal_length, n, 2, self.device)
        self.principal_point = _as_batch(principal_point, n, 2, self.device)
        self.K = K
        self._in_ndc = in_ndc
        if image_size is not None:
            self.image_size = _as_batch(image_size, n, 2, self.device)  # (H, W)
        else:
            self.image_size = None

    # VoGE sets `cameras.focal` / `cameras.principal` through setattr (Renderer.py:104-109); these are
    # plain attributes on the pytorch3d object too (no effect on the projection there either).
    def __len__(self):
        return self._N

    def in_ndc(self) -> bool:
        return self._in_ndc

    def is_perspective(self) -> bool:
        return True

    def to(self, device):
        device = torch.device(device)
        other = PerspectiveCameras.__new__(PerspectiveCameras)
        other.__dict__.update(self.__dict__)
        other.device = device
        for k in ("R", "T", "focal_length", "principal_point", "image_size"):
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(other, k, v.to(device))
        return other

    def get_image_size(self):
        return self.image_size

    def get_principal_point(self, **kwargs):
        return kwargs.get("principal_point", self.principal_point)

    def _batch(self):
        return max(self._N, self.R.shape[0], self.T.shape[0])

    def _RT(self, **kwargs):
        R = kwargs.get("R", self.R)
        T = kwargs.get("T", self.T)
        n = max(R.shape[0], T.shape[0])
        if R.shape[0] != n:
            R = R.expand(n, -1, -1)
        if T.shape[0] != n:
            T = T.expand(n, -1)
        return R, T

    def get_camera_center(self, **kwargs) -> torch.Tensor:
        R, T = self._RT(**kwargs)
        return -torch.matmul(T[:, None, :], R.transpose(1, 2))[:, 0, :]

    def get_world_to_view_transform(self, **kwargs) -> Transform3d:
        R, T = self._RT(**kwargs)
        n = R.shape[0]
        m = torch.zeros(n, 4, 4, dtype=torch.float32, device=R.device)
        m[:, :3, :3] = R
        m[:, 3, :3] = T
        m[:, 3, 3] = 1.0
        return Transform3d(matrix=m)

    def _intrinsics(self, **kwargs):
        n = self._batch()
        f = _as_batch(kwargs.get("focal_length", self.focal_length), n, 2, self.R.device)
        p = _as_batch(kwargs.get("principal_point", self.principal_point), n, 2, self.R.device)
        return f, p

    def get_projection_transform(self, **kwargs) -> Transform3d:
        f, p = self._intrinsics(**kwargs)
        n = f.shape[0]
        Kt = torch.zeros(n, 4, 4, dtype=torch.float32, device=f.device)  # already transposed (row-vector)
        Kt[:, 0, 0] = f[:, 0]
        Kt[:, 1, 1] = f[:, 1]
        Kt[:, 2, 0] = p[:, 0]
        Kt[:, 2, 1] = p[:, 1]
        Kt[:, 3, 2] = 1.0
        Kt[:, 2, 3] = 1.0
        return Transform3d(matrix=Kt)

    def get_full_projection_transform(self, **kwargs) -> Transform3d:
        return self.get_world_to_view_transform(**kwargs).compose(self.get_projection_transform(**kwargs))

    def get_ndc_camera_transform(self, **kwargs) -> Transform3d:
        if self.in_ndc():
            return Transform3d(device=self.R.device)
        if self.image_size is None:
            raise ValueError("image_size must be set for screen-space cameras")
        f, p = self._intrinsics(**kwargs)
        n = f.shape[0]
        img = _as_batch(self.image_size, n, 2, f.device)  # (H, W)
        height, width = img[:, 0], img[:, 1]
        scale = torch.minimum(height, width) / 2.0
        fix = torch.eye(4, dtype=torch.float32, device=f.device)[None].repeat(n, 1, 1)
        fix[:, 3, 0] = -2.0 * p[:, 0]
        fix[:, 3, 1] = -2.0 * p[:, 1]
        # inverse of ndc->screen  x_s = scale*x_ndc - W/2
        s2n = torch.eye(4, dtype=torch.float32, device=f.device)[None].repeat(n, 1, 1)
        s2n[:, 0, 0] = 1.0 / scale
        s2n[:, 1, 1] = 1.0 / scale
        s2n[:, 3, 0] = (width / 2.0) / scale
        s2n[:, 3, 1] = (height / 2.0) / scale
        return Transform3d(matrix=fix).compose(Transform3d(matrix=s2n))

    def transform_points(self, points, **kwargs):
        return self.get_full_projection_transform(**kwargs).transform_points(points)


def camera_params(cameras, image_size):
    """(R (B,3,3), T (B,3), focal (B,2), principal (B,2)) as float32 tensors from a camera duck-type."""
    R, T = cameras.R, cameras.T
    n = max(R.shape[0], T.shape[0])
    f = cameras.focal_length
    p = cameras.principal_point
    dev = R.device
    f = _as_batch(f, n, 2, dev)
    p = _as_batch(p, n, 2, dev)
    if R.shape[0] != n:
        R = R.expand(n, -1, -1)
    if T.shape[0] != n:
        T = T.expand(n, -1)
    return R.to(torch.float32), T.to(torch.float32), f, p


def generate_rays(cameras, image_size: Tuple[int, int]):
    """Closed-form equivalent of NDCMultinomialRaysampler(W, H, unit_directions=True)(cameras) as the
    reference uses it (Renderer.py:124-128): returns (directions (B,H,W,3) unit, origins (B,3)).
    Differentiable w.r.t. R, T, focal (plain torch ops)."""
    H, W = int(image_size[0]), int(image_size[1])
    R, T, f, p = camera_params(cameras, image_size)
    dev = R.device
    xs = torch.arange(W, dtype=torch.float32, device=dev) + 0.5
    ys = torch.arange(H, dtype=torch.float32, device=dev) + 0.5
    dx = -(xs[None, None, :] - p[:, 0, None, None]) / f[:, 0, None, None]   # (B,1,W)
    dy = -(ys[None, :, None] - p[:, 1, None, None]) / f[:, 1, None, None]   # (B,H,1)
    B = R.shape[0]
    d_cam = torch.stack([dx.expand(B, H, W), dy.expand(B, H, W), torch.ones(B, H, W, device=dev)], dim=-1)
    d_cam = F.normalize(d_cam, dim=-1)
    # world direction = d_cam @ R^T
    d_world = torch.matmul(d_cam.view(B, H * W, 3), R.transpose(1, 2)).view(B, H, W, 3)
    origins = -torch.matmul(T[:, None, :], R.transpose(1, 2))[:, 0, :]
    return d_world, origins


def look_at_rotation(camera_position, at=((0, 0, 0),), up=((0, 1, 0),), device="cpu") -> torch.Tensor:
    def prep(v):
        v = v if torch.is_tensor(v) else torch.tensor(v, dtype=torch.float32)
        v = v.to(device=device, dtype=torch.float32)
        return v.view(1, 3) if v.dim() == 1 else v
    c, at, up = prep(camera_position), prep(at), prep(up)
    n = max(c.shape[0], at.shape[0], up.shape[0])
    c, at, up = c.expand(n, 3), at.expand(n, 3), up.expand(n, 3)
    z_axis = F.normalize(at - c, eps=1e-5)
    x_axis = F.normalize(torch.cross(up, z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    is_close = torch.isclose(x_axis, torch.tensor(0.0, device=x_axis.device), atol=5e-3).all(dim=1, keepdim=True)
    if is_close.any():
        replacement = F.normalize(torch.cross(y_axis, z_axis, dim=1), eps=1e-5)
        x_axis = torch.where(is_close, replacement, x_axis)
    R = torch.cat((x_axis[:, None, :], y_axis[:, None, :], z_axis[:, None, :]), dim=1)
    return R.transpose(1, 2)


def look_at_view_transform(dist=1.0, elev=0.0, azim=0.0, degrees: bool = True, eye=None,
                           at=((0, 0, 0),), up=((0, 1, 0),), device="cpu"):
    """pytorch3d.renderer.look_at_view_transform restated: returns (R (N,3,3), T (N,3))."""
    def prep(v):
        v = v if torch.is_tensor(v) else torch.tensor(v, dtype=torch.float32)
        return v.to(device=device, dtype=torch.float32).view(-1)
    if eye is not None:
        C = eye if torch.is_tensor(eye) else torch.tensor(eye, dtype=torch.float32)
        C = C.to(device=device, dtype=torch.float32).view(-1, 3)
    else:
        dist, elev, azim = prep(dist), prep(elev), prep(azim)
        n = max(dist.shape[0], elev.shape[0], azim.shape[0])
        dist, elev, azim = dist.expand(n), elev.expand(n), azim.expand(n)
        if degrees:
            elev = math.pi / 180.0 * elev
            azim = math.pi / 180.0 * azim
        x = dist * torch.cos(elev) * torch.sin(azim)
        y = dist * torch.sin(elev)
        z = dist * torch.cos(elev) * torch.cos(azim)
        at_t = at if torch.is_tensor(at) else torch.tensor(at, dtype=torch.float32)
        C = torch.stack([x, y, z], dim=1) + at_t.to(device=device, dtype=torch.float32).view(-1, 3)
    R = look_at_rotation(C, at=at, up=up, device=device)
    T = -torch.bmm(R.transpose(1, 2), C[:, :, None])[:, :, 0]
    return R, T

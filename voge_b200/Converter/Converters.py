"""Mesh / point cloud -> Gaussian ellipsoid converters (reference VoGE/Converter/Converters.py).

Same functions, arguments and (verts, isigma, radians) returns.  The per-vertex neighbour statistics are
vectorised (the reference loops over faces and vertices in Python) and the point-cloud k-NN runs in chunks
of torch.cdist on whatever device the points live on (the reference builds the (N,N,3) difference tensor in
batches on the CPU)."""
import math
import os

import numpy as np
import torch

from ..Meshes import GaussianMeshes


def get_vert_edge_length(verts, faces, default_l=1e-3):
    """Mean distance from every vertex to the distinct vertices of the faces that contain it
    (reference :10-32: sum of distances over the unique set, which includes the vertex itself, divided by
    its size minus one); `default_l` for vertices without a face."""
    verts = np.asarray(verts, dtype=np.float64)
    faces = np.asarray(faces)[:, :3].astype(np.int64)
    n = verts.shape[0]
    # directed pairs (v, u) for every pair of corners of a face, deduplicated
    a = np.repeat(faces, 3, axis=1).reshape(-1)
    b = np.tile(faces, (1, 3)).reshape(-1)
    pairs = np.unique(a * n + b)
    v, u = pairs // n, pairs % n
    dist = np.sqrt(((verts[v] - verts[u]) ** 2).sum(1))
    len_sum = np.bincount(v, weights=dist, minlength=n)
    cnt = np.bincount(v, minlength=n)                     # distinct neighbours including the vertex itself
    out = np.full(n, float(default_l))
    has = cnt > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        out[has] = len_sum[has] / (cnt[has] - 1)
    return out


def _look_at_rotation(camera_position):
    """pytorch3d.renderer.look_at_rotation(camera_position) with at = 0, up = +y: columns x, y, z of the
    returned matrices are the camera axes; a direction parallel to `up` falls back to x = y x z."""
    cam = torch.as_tensor(camera_position, dtype=torch.float32).reshape(-1, 3)
    up = torch.tensor([[0.0, 1.0, 0.0]]).expand_as(cam)
    z = torch.nn.functional.normalize(-cam, eps=1e-5)
    x = torch.nn.functional.normalize(torch.cross(up, z, dim=1), eps=1e-5)
    y = torch.nn.functional.normalize(torch.cross(z, x, dim=1), eps=1e-5)
    close = torch.isclose(x, torch.zeros(()), atol=5e-3).all(dim=1, keepdim=True)
    if close.any():
        x = torch.where(close, torch.nn.functional.normalize(torch.cross(y, z, dim=1), eps=1e-5), x)
    return torch.stack([x, y, z], dim=2)


def _cap(isigma, max_sig_rate):
    if max_sig_rate > 0:
        thr = np.mean(isigma) * max_sig_rate
        isigma[isigma > thr] = thr
    return isigma


def _default_len(vertices):
    return 10 * np.sum((vertices.max(axis=0) - vertices.min(axis=0)) ** 2) ** 0.5 / vertices.shape[0]


def normal_mesh_converter(vertices, faces, normals, percentage=0.5, shape_ratio=0.5, max_sig_rate=-1, auto_fix=True):
    """Anisotropic Gaussians flattened along the vertex normals: isigma = R diag(b, b, shape_ratio b) R^T with
    b = 1 / (len^2 / (2 ln(1/percentage))) and R = look_at_rotation(-normal) (reference :35-73)."""
    is_torch = torch.is_tensor(vertices)
    if is_torch:
        vertices, faces = vertices.numpy(), faces.numpy()
    normals = torch.as_tensor(normals, dtype=torch.float32)
    n2 = (normals ** 2).sum(-1)
    assert n2.max() < 1.1 and n2.min() > 0.9
    average_len = get_vert_edge_length(vertices, faces, _default_len(vertices))
    base = 1 / ((average_len ** 2) / (2 * np.log(1 / percentage)) + 1e-10)
    diag = np.diag([1.0, 1.0, shape_ratio])[None] * base.reshape(-1, 1, 1)
    R = _look_at_rotation(-normals).numpy().astype(np.float64)
    isigma = R @ diag @ R.transpose(0, 2, 1)
    if auto_fix:
        bad = np.linalg.det(isigma) == 0
        isigma[bad] = np.eye(3)[None] * base[bad].reshape(-1, 1, 1)
    isigma = _cap(isigma, max_sig_rate)
    if is_torch:
        return torch.from_numpy(vertices).type(torch.float32), torch.from_numpy(isigma).type(torch.float32), None
    return vertices, isigma, None


def naive_vertices_converter(vertices, faces, percentage=0.5, max_sig_rate=-1):
    """Isotropic Gaussians sized by the mean edge length at each vertex (reference :76-97)."""
    is_torch = torch.is_tensor(vertices)
    if is_torch:
        vertices, faces = vertices.numpy(), faces.numpy()
    average_len = get_vert_edge_length(vertices, faces, _default_len(vertices))
    isigma = _cap(1 / ((average_len ** 2) / (2 * np.log(1 / percentage)) + 1e-10), max_sig_rate)
    if is_torch:
        return torch.from_numpy(vertices).type(torch.float32), torch.from_numpy(isigma).type(torch.float32), None
    return vertices, isigma, None


def naive_point_cloud_converter(points, percentage=0.5, n_nearest=4, thr_max=2, chunk=4096):
    """Isotropic Gaussians sized by the n_nearest smallest distances of each point (its own zero distance
    included, each clipped at thr_max x their mean); reference :100-122."""
    to_np = not torch.is_tensor(points)
    pts = (torch.from_numpy(points) if to_np else points).type(torch.float32)
    if pts.is_cuda and 1 <= n_nearest <= 16 and n_nearest <= pts.shape[0]:
        # CUDA points: one hand-written kernel (csrc/knn.cu), no (chunk, N) distance matrices
        from .. import _C
        with torch.no_grad():
            avg = _C.knn_mean_dist(pts, n_nearest, thr_max)
            isigma = 1 / ((avg ** 2) / (4 * math.log(1 / percentage)) + 1e-8)
        return pts, isigma, None
    sigma = torch.empty(pts.shape[0], dtype=torch.float32, device=pts.device)
    with torch.no_grad():
        for s in range(0, pts.shape[0], chunk):
            d = torch.cdist(pts[s:s + chunk], pts, compute_mode="donot_use_mm_for_euclid_dist")
            top = torch.topk(d, k=n_nearest, dim=1, largest=False)[0]
            avg = torch.min(top, top.mean(dim=1, keepdim=True).expand(-1, n_nearest) * thr_max).mean(dim=1)
            sigma[s:s + chunk] = (avg ** 2) / (4 * math.log(1 / percentage))
    isigma = 1 / (sigma + 1e-8)
    if to_np:
        return pts.numpy(), isigma.cpu().numpy(), None
    return pts, isigma, None


def fixed_pointcloud_converter(points, radius, percentage=0.5):
    """Isotropic Gaussians of a given radius (reference :125-137)."""
    to_np = not torch.is_tensor(points)
    if to_np:
        points = torch.from_numpy(points)
        if not isinstance(radius, float):
            radius = torch.from_numpy(radius)
    isigma = torch.ones(points.shape[0]) / ((radius ** 2) / (2 * np.log(1 / percentage)) + 1e-10)
    if to_np:
        return points.numpy(), isigma.numpy(), None
    return points, isigma, None


def convert_path(source_path, destiny_path, convert_function, filter_=None):
    """Apply convert_function(src, dst) to every file below source_path, mirroring the tree (reference :140-153)."""
    os.makedirs(destiny_path, exist_ok=True)
    for name in os.listdir(source_path):
        src, dst = os.path.join(source_path, name), os.path.join(destiny_path, name)
        if os.path.isfile(src):
            if filter_ is None or filter_(name):
                convert_function(src, dst)
        else:
            convert_path(src, dst, convert_function)


class ComposedConverter(object):
    """loader(path) -> converter(*loaded, **kwargs) -> saver(path, *converted); reference :156-172."""

    def __init__(self, loader, saver, converter, **kwargs):
        self.loader, self.saver, self.converter, self.kwargs = loader, saver, converter, kwargs

    def __call__(self, source_path, destiny_path):
        got = self.loader(source_path)
        got = self.converter(*(got if isinstance(got, tuple) else (got,)), **self.kwargs)
        self.saver(destiny_path, *(got if isinstance(got, tuple) else (got,)))


def pytorch3d2gaussian(converter, **kwargs):
    """Wrap a converter so that it takes a pytorch3d Meshes / Pointclouds (duck-typed: verts_packed /
    faces_packed / points_packed) and returns GaussianMeshes on the input's device; reference :175-195."""
    def wrapper(input_, **mesh_kwargs):
        if hasattr(input_, "verts_packed"):
            mesh = input_[0] if len(input_) > 1 else input_
            verts, sigmas, radians = converter(mesh.verts_packed().cpu(), mesh.faces_packed().cpu(), **kwargs)
        else:
            verts, sigmas, radians = converter(input_.points_packed(), **kwargs)
        return GaussianMeshes(verts.type(torch.float32), sigmas.type(torch.float32),
                              radians.type(torch.float32) if radians is not None else None, **mesh_kwargs).to(input_.device)
    return wrapper

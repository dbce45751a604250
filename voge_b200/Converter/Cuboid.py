"""Cuboid test objects (reference VoGE/Converter/Cuboid.py): `cuboid_gauss` :8-67 -- a regular grid of isotropic
Gaussians on the faces of an axis-aligned box -- and `cuboid_mesh` :70-159 -- the same box as a triangle mesh, six
full face grids.  Same arguments, vertex / face orders, return types (ndarrays; `as_obj=True` wraps them) and per-face
colour expansion as the reference; the grids are built with numpy instead of Python loops."""
import numpy as np
import torch

from ..scenes import cuboid_gauss as _cuboid_gauss


def _axis_samples(x_range, y_range, z_range, number_vertices):
    """Sample positions along the three axes: spacing ~ sqrt(2 * area / (2 n)) (reference :9-19)."""
    w, h, d = (r[1] - r[0] for r in (x_range, y_range, z_range))
    edge = (2.0 * (w * h + h * d + w * d) / (number_vertices * 2) * 2) ** 0.5
    return tuple(r[0] + np.linspace(0, ext, int(ext / edge + 1)) for r, ext in ((x_range, w), (y_range, h), (z_range, d)))


def _expand_colors(colors, counts):
    """One colour row per face -> one row per vertex (reference :53, :143)."""
    colors = np.asarray(colors)
    return np.concatenate([np.repeat(c[None, :], r, axis=0) for r, c in zip(counts, colors)], axis=0)


def cuboid_gauss(x_range, y_range, z_range, number_vertices, percentage=0.5, colors=None, as_obj=False):
    """-> (verts (n,3) float64 ndarray, isigma (n,) float64 ndarray[, colors (n,C)]); as_obj=True returns
    GaussianMeshes (float32 parameters)[, colors].  `colors` holds one row per face in the order z-min, z-max,
    y-min, y-max, x-min, x-max."""
    verts, isigma = _cuboid_gauss(x_range, y_range, z_range, number_vertices, percentage=percentage)
    out = (verts, isigma)
    if as_obj:
        from ..Meshes import GaussianMeshes
        out = (GaussianMeshes(verts=torch.from_numpy(verts).type(torch.float32),
                              sigmas=torch.from_numpy(isigma).type(torch.float32)),)
    if colors is not None:
        xs, ys, zs = _axis_samples(x_range, y_range, z_range, number_vertices)
        xn, yn, zn = xs.size, ys.size, zs.size
        counts = [yn * xn, yn * xn, (zn - 2) * (xn - 1), (zn - 2) * (xn - 1), (zn - 2) * (yn - 1), (zn - 2) * (yn - 1)]
        out = out + (_expand_colors(colors, counts),)
    return out[0] if len(out) == 1 else out


def _grid_face(us, vs, fixed, order):
    """Vertices of a full face grid, `us` fastest; order = positions of (u, v, fixed) among (x, y, z)."""
    uu, vv = np.meshgrid(us, vs)                     # (len(vs), len(us)), u fastest
    cols = [None, None, None]
    cols[order[0]], cols[order[1]], cols[order[2]] = uu.reshape(-1), vv.reshape(-1), np.full(uu.size, fixed)
    return np.stack(cols, axis=1)


def _grid_triangles(base, rows, cols):
    """Two triangles per grid cell, cell-major: (p, p+1, p+cols), (p+cols+1, p+1, p+cols) with p = base + m cols + n."""
    m, n = np.meshgrid(np.arange(rows - 1), np.arange(cols - 1), indexing="ij")
    p = (base + m * cols + n).reshape(-1)
    t1 = np.stack([p, p + 1, p + cols], axis=1)
    t2 = np.stack([p + cols + 1, p + 1, p + cols], axis=1)
    return np.stack([t1, t2], axis=1).reshape(-1, 3)


class _SimpleMeshes(object):
    """Stand-in for pytorch3d.structures.Meshes when pytorch3d is not installed (verts_list / faces_list only)."""

    def __init__(self, verts, faces):
        self._verts, self._faces = list(verts), list(faces)

    def verts_list(self):
        return self._verts

    def faces_list(self):
        return self._faces

    def verts_packed(self):
        return torch.cat(self._verts, 0)

    def faces_packed(self):
        return torch.cat(self._faces, 0)


def cuboid_mesh(x_range, y_range, z_range, number_vertices, colors=None, as_obj=False):
    """-> (verts (n,3) float64, faces (m,3) int64[, colors]); as_obj=True returns a Meshes object (pytorch3d's when
    it is installed)[, colors].  Six full face grids in the order z-min, z-max, y-min, y-max, x-min, x-max; the box
    edges are duplicated between adjacent faces, as in the reference (:88-141)."""
    xs, ys, zs = _axis_samples(x_range, y_range, z_range, number_vertices)
    xn, yn, zn = xs.size, ys.size, zs.size
    faces_spec = [(xs, ys, zs[0], (0, 1, 2)), (xs, ys, zs[-1], (0, 1, 2)),       # z faces: x fastest, then y
                  (xs, zs, ys[0], (0, 2, 1)), (xs, zs, ys[-1], (0, 2, 1)),       # y faces: x fastest, then z
                  (ys, zs, xs[0], (1, 2, 0)), (ys, zs, xs[-1], (1, 2, 0))]       # x faces: y fastest, then z
    verts, tris, counts, base = [], [], [], 0
    for us, vs, fixed, order in faces_spec:
        verts.append(_grid_face(us, vs, fixed, order))
        tris.append(_grid_triangles(base, vs.size, us.size))
        counts.append(vs.size * us.size)
        base += vs.size * us.size
    verts, tris = np.concatenate(verts, 0), np.concatenate(tris, 0).astype(np.int64)
    out = (verts, tris)
    if as_obj:
        v_t, f_t = torch.from_numpy(verts).type(torch.float32), torch.from_numpy(tris).type(torch.long)
        try:
            from pytorch3d.structures import Meshes
            out = (Meshes(verts=[v_t], faces=[f_t]),)
        except ImportError:
            out = (_SimpleMeshes([v_t], [f_t]),)
    if colors is not None:
        out = out + (_expand_colors(colors, counts),)
    return out[0] if len(out) == 1 else out

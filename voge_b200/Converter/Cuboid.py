"""Cuboid test objects (reference VoGE/Converter/Cuboid.py:8-68): a regular grid of isotropic Gaussians on the
faces of an axis-aligned box.  The generator lives in voge_b200.scenes (bench / tests use it too)."""
import numpy as np
import torch

from ..scenes import cuboid_gauss as _cuboid_gauss


def cuboid_gauss(x_range, y_range, z_range, number_vertices, percentage=0.5, colors=None, as_obj=False):
    """-> (verts (n,3), isigma (n,)) float32 tensors; with `colors` a third (n,3) tensor (one colour per
    Gaussian, broadcast from a single RGB triple).  as_obj=True wraps them in GaussianMeshes."""
    v, s = _cuboid_gauss(x_range, y_range, z_range, number_vertices, percentage=percentage)
    verts = torch.from_numpy(np.asarray(v, dtype=np.float32))
    sig = torch.from_numpy(np.asarray(s, dtype=np.float32))
    out = (verts, sig)
    if colors is not None:
        col = torch.as_tensor(colors, dtype=torch.float32).reshape(-1, 3)
        out = out + (col.expand(verts.shape[0], 3).contiguous() if col.shape[0] == 1 else col,)
    if as_obj:
        from ..Meshes import GaussianMeshes
        return GaussianMeshes(verts, sig)
    return out

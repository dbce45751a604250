"""Workload generators for tests and bench.py: the scenes of BASELINE.json's configs, built without
pytorch3d / mesh files (there is no network and `/root/reference` is absent on the GPU box).

  cuboid_gauss     -- procedural cuboid surface Gaussians (restates reference
                      VoGE/Converter/Cuboid.py:8-67; same vertex order, same sigma formula)
  ico_sphere       -- subdivided icosahedron vertices (ShapeFitting demo uses ico_sphere(4): 2562 verts)
  synthetic_scene  -- the C5 scale-sweep distribution of SURVEY.md 8(d)
  orbit_cameras    -- look_at_view_transform orbits
"""
import math

import numpy as np
import torch

from .cameras import PerspectiveCameras, look_at_view_transform


def cuboid_gauss(x_range, y_range, z_range, number_vertices, percentage=0.5):
    """-> (verts (n,3) float64 ndarray, isigma (n,) float64 ndarray).
    A regular grid on the six faces of the box with spacing chosen so that about
    `number_vertices` vertices tile the surface; one isotropic inverse-variance for all."""
    w, h, d = (r[1] - r[0] for r in (x_range, y_range, z_range))
    area = 2.0 * (w * h + h * d + w * d)
    edge = math.sqrt(2.0 * area / (number_vertices * 2))
    xs = x_range[0] + np.linspace(0, w, int(w / edge + 1))
    ys = y_range[0] + np.linspace(0, h, int(h / edge + 1))
    zs = z_range[0] + np.linspace(0, d, int(d / edge + 1))
    xn, yn, zn = xs.size, ys.size, zs.size
    v = []
    v += [(xs[m], ys[n], zs[0]) for n in range(yn) for m in range(xn)]            # z-min face
    v += [(xs[m], ys[n], zs[-1]) for n in range(yn) for m in range(xn)]           # z-max face
    v += [(xs[m], ys[0], zs[n]) for n in range(1, zn - 1) for m in range(xn - 1)]  # y-min strip
    v += [(xs[m], ys[-1], zs[n]) for n in range(1, zn - 1) for m in range(1, xn)]  # y-max strip
    v += [(xs[0], ys[m], zs[n]) for n in range(1, zn - 1) for m in range(1, yn)]   # x-min strip
    v += [(xs[-1], ys[m], zs[n]) for n in range(1, zn - 1) for m in range(yn - 1)]  # x-max strip
    sigma = edge ** 2 / (2 * np.log(1 / percentage)) + 1e-10
    verts = np.array(v)
    return verts, np.ones(len(v)) * (1.0 / sigma)


def ico_sphere(level=0):
    """Unit icosphere vertices after `level` midpoint subdivisions: 12, 42, 162, 642, 2562, ..."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    verts = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
             (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
             (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5),
             (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = [np.array(v, dtype=np.float64) / np.linalg.norm(v) for v in verts]
    for _ in range(level):
        cache, new_faces = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            new_faces += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = new_faces
    return np.stack(verts), np.array(faces)


def synthetic_scene(n=1_000_000, seed=0, aniso_fraction=0.1, device="cpu"):
    """C5 distribution: verts ~ U[-1,1]^3; inverse-covariance s = ln(100)/(2 r^2) with
    cut-off radius r ~ logU[0.004, 0.012]; `aniso_fraction` of the Gaussians are R diag(s,s,4s) R^T with
    random rotations; colours ~ U[0,1]^3.  Returns (verts (n,3), sigmas (n,3,3), colors (n,3))."""
    g = torch.Generator().manual_seed(seed)
    verts = torch.rand(n, 3, generator=g) * 2 - 1
    r = torch.exp(torch.rand(n, generator=g) * (math.log(0.012) - math.log(0.004)) + math.log(0.004))
    s = math.log(100.0) / (2 * r * r)
    sig = torch.zeros(n, 3, 3)
    sig[:, 0, 0] = s
    sig[:, 1, 1] = s
    sig[:, 2, 2] = s
    n_an = int(n * aniso_fraction)
    if n_an > 0:
        q = torch.randn(n_an, 4, generator=g)
        q = q / q.norm(dim=1, keepdim=True)
        a, b, c, d = q.unbind(1)
        Rm = torch.stack([1 - 2 * (c * c + d * d), 2 * (b * c - a * d), 2 * (b * d + a * c),
                          2 * (b * c + a * d), 1 - 2 * (b * b + d * d), 2 * (c * d - a * b),
                          2 * (b * d - a * c), 2 * (c * d + a * b), 1 - 2 * (b * b + c * c)], dim=1).view(-1, 3, 3)
        diag = torch.zeros(n_an, 3, 3)
        diag[:, 0, 0] = s[:n_an]
        diag[:, 1, 1] = s[:n_an]
        diag[:, 2, 2] = 4 * s[:n_an]
        m = Rm @ diag @ Rm.transpose(1, 2)
        sig[:n_an] = 0.5 * (m + m.transpose(1, 2))
    colors = torch.rand(n, 3, generator=g)
    return verts.to(device), sig.to(device), colors.to(device)


def orbit_cameras(n_views, dist=3.0, elev_amp=20.0, focal=900.0, image_size=(1024, 1024), device="cpu",
                  first=0, count=None, indices=None):
    """Views i = first .. first+count-1 (or the given `indices`) of an n_views orbit: elev = elev_amp*sin(2 pi i/n),
    azim = 360 i/n."""
    count = n_views if count is None else count
    i = torch.arange(first, first + count, dtype=torch.float32) if indices is None else torch.tensor(list(indices), dtype=torch.float32)
    elev = elev_amp * torch.sin(2 * math.pi * i / n_views)
    azim = 360.0 * i / n_views
    R, T = look_at_view_transform(dist=dist, elev=elev, azim=azim, device=device)
    H, W = image_size
    return PerspectiveCameras(focal_length=focal, principal_point=((W / 2.0, H / 2.0),), R=R, T=T, device=device,
                              in_ndc=False, image_size=(image_size,))

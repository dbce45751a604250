"""Stored hits per pixel on the C2 / C4 bench scenes (which selection path their pixels take)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from voge_b200 import scenes, _C
from voge_b200.cameras import PerspectiveCameras, camera_params, generate_rays, look_at_view_transform
from voge_b200.fused import choose_tile
from voge_b200.RayTracing import default_bin_size
from voge_b200.Converter.Converters import naive_vertices_converter
dev = "cuda:0"

def hist(name, verts, sig, hw, K, focal, cam, M):
    R, T = look_at_view_transform(dist=cam[0], elev=torch.tensor(cam[1]), azim=torch.tensor(cam[2]))
    cams = PerspectiveCameras(focal_length=focal, principal_point=((hw / 2.0, hw / 2.0),), R=R, T=T, in_ndc=False, image_size=((hw, hw),), device=dev)
    rays, origins = generate_rays(cams, (hw, hw))
    R, T, f, p = camera_params(cams, (hw, hw))
    thr_act = -math.log(0.01 + 1e-10)
    bs = default_bin_size((hw, hw)); tile = choose_tile(bs, K, M != -1)
    verts, sig = verts.to(dev), sig.to(dev)
    off, tl, rects, ioff = _C.bin_views(verts, sig, R, T, origins, f, p, (hw, hw), 0.01, thr_act, M != -1, bs, tile)
    dbg = {}
    _C.render_forward(verts, sig, origins, rays, off, tl, rects, thr_act, 1.0, K, tile, need_act=False, item_offsets=ioff, debug=dbg)
    c = dbg["counts"].to(torch.int64)
    nz = c[c > 0]
    print(name, "pixels with hits %d: mean %.1f" % (nz.numel(), nz.float().mean()), " ".join(
        "(%d,%s]: %.1f%%" % (lo, hi, 100.0 * int(((nz > lo) & (nz <= hi)).sum()) / nz.numel())
        for lo, hi in ((0, 32), (32, 64), (64, 128), (128, 256), (256, 10 ** 9))))

mv, mf = scenes.ico_sphere(6)
rng = np.random.RandomState(0)
dirs = rng.randn(6, 3); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
bump = 1.0 + 0.12 * sum(np.sin(3.0 * mv @ d + i) for i, d in enumerate(dirs)) / len(dirs)
vn, sn, _ = naive_vertices_converter(mv * bump[:, None] * 0.3, mf, percentage=0.6)
hist("C2", torch.tensor(vn, dtype=torch.float32), torch.tensor(sn, dtype=torch.float32), 512, 40, 4000.0, (6.0, [0.0], [10.0]), None)
v1, s1 = scenes.cuboid_gauss((-0.6, 0.6), (-0.4, 0.4), (-0.5, 0.5), 1500, percentage=0.6)
v2, s2 = scenes.cuboid_gauss((-0.5, 0.5), (-0.5, 0.5), (-0.3, 0.3), 1200, percentage=0.6)
verts = torch.tensor(np.concatenate([v1, v2 + np.array([0.4, 0.1, -0.9])]), dtype=torch.float32)
sig = torch.tensor(np.concatenate([s1, s2]), dtype=torch.float32)
hist("C4", verts, sig, 400, 60, 300.0, (4.0, [15.0], [30.0]), 1500)

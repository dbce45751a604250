"""Development aid: fwd+bwd time of the fused path on C2-like / C4-like scenes (mesh-converted Gaussians with
heavy overlap, large K) -- a sanity check that nothing degenerates away from the C5 regime."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
from voge_b200 import scenes
from voge_b200.cameras import PerspectiveCameras, look_at_view_transform
from voge_b200.Converter.Converters import naive_vertices_converter
from voge_b200.Meshes import GaussianMeshes
from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings, to_white_background
dev = "cuda:0"


def run(name, verts, sig, hw, K, focal, dist, views=4, M=None, iters=5):
    R, T = look_at_view_transform(dist=dist, elev=[10.0 + 5 * i for i in range(views)], azim=[40.0 * i for i in range(views)])
    cams = PerspectiveCameras(focal_length=focal, principal_point=((hw / 2, hw / 2),), R=R, T=T, in_ndc=False,
                              image_size=((hw, hw),), device=dev)
    r = GaussianRenderer(cams, GaussianRenderSettings(image_size=(hw, hw), max_assign=K, max_point_per_bin=M)).to(dev)
    gm = GaussianMeshes(verts, sig).to(dev)
    col = torch.rand(verts.shape[0], 3, device=dev, requires_grad=True)
    tgt = torch.rand(views, hw, hw, 3, device=dev)
    for it in range(iters + 2):
        if it == 2:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        frag = r(gm)
        loss = ((to_white_background(frag, col) - tgt) ** 2).mean()
        loss.backward()
        gm.zero_grad(); col.grad = None
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / iters / views * 1e3
    print("%-28s N=%6d %4dx%-4d K=%3d  %.3f ms/view fwd+bwd  (%.1f Mrays/s)  mean valid %.1f max %d" % (
        name, verts.shape[0], hw, hw, K, ms, hw * hw / ms / 1e3, float(frag.valid_num.float().mean()), int(frag.valid_num.max())))


v, f = scenes.ico_sphere(6)
vv, ss, _ = naive_vertices_converter(torch.tensor(v, dtype=torch.float32), torch.tensor(f), percentage=0.5)
run("C2-like sphere mesh", vv, ss, 512, 40, 600.0, 3.0)
run("C2-like, K=80", vv, ss, 512, 80, 600.0, 3.0)
v1, s1 = scenes.cuboid_gauss((-0.6, 0.6), (-0.4, 0.4), (-0.5, 0.5), 4000, percentage=0.6)
v2, s2 = scenes.cuboid_gauss((-0.5, 0.5), (-0.5, 0.5), (-0.3, 0.3), 3000, percentage=0.6)
verts = torch.tensor(np.concatenate([v1, v2 + np.array([0.4, 0.1, -0.9])]), dtype=torch.float32)
sig = torch.tensor(np.concatenate([s1, s2]), dtype=torch.float32)
run("C4-like two cuboids", verts, sig, 400, 60, 300.0, 4.0, M=1500)
v3, s3 = scenes.cuboid_gauss((-1, 1), (-1, 1), (-1, 1), 1000, percentage=0.6)
run("C1 quick-start cuboid", torch.tensor(v3, dtype=torch.float32), torch.tensor(s3, dtype=torch.float32), 256, 20, 300.0, 6.0, M=200, views=1)
v4, _ = scenes.ico_sphere(4)
run("C3 ico_sphere(4), no coarse", torch.tensor(v4, dtype=torch.float32), torch.full((2562,), 400.0), 128, 25, 150.0, 2.7, M=-1, views=5)

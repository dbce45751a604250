import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, numpy as np
from voge_b200 import scenes
from voge_b200.cameras import PerspectiveCameras, look_at_view_transform
from voge_b200.Meshes import GaussianMeshes
from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings, to_white_background
dev = "cuda:0"
v, s = scenes.cuboid_gauss((-1, 1), (-1, 1), (-1, 1), 1000, percentage=0.6)
verts = torch.tensor(v, dtype=torch.float32); sig = torch.tensor(s, dtype=torch.float32)
H = W = 64
R, T = look_at_view_transform(dist=6, elev=10, azim=70)
cams = PerspectiveCameras(focal_length=75.0, principal_point=((W / 2, H / 2),), R=R, T=T, in_ndc=False, image_size=((H, W),), device=dev)
st = GaussianRenderSettings(image_size=(H, W), max_assign=20, thr_activation=0.01, max_point_per_bin=866)
torch.manual_seed(0)
col0 = torch.rand(verts.shape[0], 3, device=dev)
res = {}
for fused in (True, False):
    r = GaussianRenderer(cams, st).to(dev); r.use_fused = fused
    gm = GaussianMeshes(verts.clone(), sig.clone()).to(dev)
    colors = col0.clone().requires_grad_(True)
    frag = r(gm)
    img = to_white_background(frag, colors)
    img.square().mean().backward()
    res[fused] = (gm.verts.grad.clone(), gm.sigmas.grad.clone(), colors.grad.clone(), frag.vert_weight.detach().clone())
    print("fused", fused, "|dverts|", float(gm.verts.grad.abs().sum()), "|dsig|", float(gm.sigmas.grad.abs().sum()), "|dcol|", float(colors.grad.abs().sum()))
for i, n in enumerate(("verts", "sig", "col", "weight")):
    a, b = res[True][i], res[False][i]
    print(n, "max abs diff", float((a - b).abs().max()), "max ref", float(b.abs().max()))

"""Small driver for ncu: a few fused fwd+bwd steps on V views of the C5 scene (default 2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from voge_b200 import scenes
from voge_b200.Meshes import GaussianMeshes
from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings, to_white_background
dev = "cuda:0"
V = int(os.environ.get("V", 2)); N = int(os.environ.get("N", 1000000)); HW = int(os.environ.get("HW", 1024))
ITERS = int(os.environ.get("ITERS", 2))
verts, sig, colors = scenes.synthetic_scene(N, device=dev)
cams = scenes.orbit_cameras(64, image_size=(HW, HW), focal=900.0 * HW / 1024, device=dev, first=0, count=V)
r = GaussianRenderer(cams, GaussianRenderSettings(image_size=(HW, HW), max_assign=20)).to(dev)
gm = GaussianMeshes(verts, sig).to(dev)
colors.requires_grad_(True)
target = torch.rand(V, HW, HW, 3, device=dev)
for it in range(ITERS):
    frag = r(gm)
    img = to_white_background(frag, colors)
    ((img - target) ** 2).mean().backward()
    gm.zero_grad(); colors.grad = None
torch.cuda.synchronize()
print("done")

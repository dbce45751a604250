"""Quick timing of the unfused API path at C5 scale (1 view) to see where time goes."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from voge_b200 import scenes
from voge_b200.Meshes import GaussianMeshes
from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings, to_white_background

dev = "cuda:0"
N = int(os.environ.get("N", 1000000)); HW = int(os.environ.get("HW", 1024))
verts, sig, colors = scenes.synthetic_scene(N, device=dev)
cams = scenes.orbit_cameras(64, image_size=(HW, HW), focal=900.0 * HW / 1024, device=dev, first=0, count=1)
st = GaussianRenderSettings(image_size=(HW, HW), max_assign=20)
r = GaussianRenderer(cams, st).to(dev)
gm = GaussianMeshes(verts, sig).to(dev)
colors.requires_grad_(True)
target = torch.rand(1, HW, HW, 3, device=dev)

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

from voge_b200 import _C
import voge_b200.RayTracing as RT
for it in range(3):
    torch.cuda.synchronize()
    t0 = ev()
    frag = r(gm)
    t1 = ev()
    img = to_white_background(frag, colors)
    loss = ((img - target) ** 2).mean()
    t2 = ev()
    loss.backward()
    t3 = ev()
    torch.cuda.synchronize()
    print("iter %d: forward(frag) %.2f ms, image+loss %.2f ms, backward %.2f ms | valid mean %.2f, bin max %d" % (
        it, t0.elapsed_time(t1), t1.elapsed_time(t2), t2.elapsed_time(t3), frag.valid_num.float().mean().item(),
        int(_C.last_bin_counts.max())))
    gm.zero_grad(); colors.grad = None

# per-op timing
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    frag = r(gm); img = to_white_background(frag, colors); loss = ((img - target) ** 2).mean(); loss.backward()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))

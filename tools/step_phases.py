"""Development aid: where a fitting step's time goes outside the kernels.  One rank, V views per call (default 8, the
per-rank share of the 8-GPU C5 step): CUDA events around the phases of bench.fit_step + the kernel sum."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from voge_b200 import scenes, _lib
from voge_b200.distributed import GradientBucket
from voge_b200.Meshes import GaussianMeshes
from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings, to_white_background

dev = "cuda:0"
V = int(os.environ.get("V", 8)); ITERS = int(os.environ.get("ITERS", 10))
verts, sig, colors = scenes.synthetic_scene(1000000)
cams = scenes.orbit_cameras(64, image_size=(1024, 1024), focal=900.0, device=dev, indices=list(range(0, 64, 64 // V)))
r = GaussianRenderer(cams, GaussianRenderSettings(image_size=(1024, 1024), max_assign=20)).to(dev)
gm = GaussianMeshes(verts, sig).to(dev)
col = torch.nn.Parameter(colors.to(dev))
bucket = GradientBucket([gm.verts, gm.sigmas, col])
target = torch.rand(V, 1024, 1024, 3, device=dev)
times = {}
class _KT:
    enabled = True
    def add(self, name, e0, e1):
        times.setdefault("k:" + name, []).append((e0, e1))
_lib.kernel_timer = _KT()

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

for it in range(ITERS + 3):
    if it == 3:
        times.clear()
    e = [ev()]
    bucket.zero(); e.append(ev())
    frag = r(gm); e.append(ev())
    img = to_white_background(frag, col); e.append(ev())
    loss = torch.nn.functional.mse_loss(img, target, reduction="sum") / (64 * 1024 * 1024 * 3); e.append(ev())
    loss.backward(); e.append(ev())
    for n, a, b in zip(("zero", "renderer_fwd", "to_white_background", "loss_fwd", "backward"), e[:-1], e[1:]):
        times.setdefault(n, []).append((a, b))
    times.setdefault("step", []).append((e[0], e[-1]))
torch.cuda.synchronize()
ks = 0.0
for n, evs in times.items():
    ms = sum(a.elapsed_time(b) for a, b in evs) / ITERS
    if n.startswith("k:"):
        ks += ms
    print("%-34s %.4f ms/step" % (n, ms))
print("kernel sum %.4f ms/step" % ks)

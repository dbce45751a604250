"""Development aid: fwd+bwd time and item / hit counts of every single view of the C5 orbit (how uneven are the views a
rank gets under a given sharding?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from voge_b200 import scenes, _C
from voge_b200.Meshes import GaussianMeshes
from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings, to_white_background
dev = "cuda:0"
verts, sig, colors = scenes.synthetic_scene(1000000)
gm = GaussianMeshes(verts, sig).to(dev)
col = colors.to(dev).requires_grad_(True)
tgt = torch.rand(1, 1024, 1024, 3, device=dev)
st = GaussianRenderSettings(image_size=(1024, 1024), max_assign=20)
ms, hits = [], []
for i in range(64):
    cams = scenes.orbit_cameras(64, image_size=(1024, 1024), focal=900.0, device=dev, indices=[i])
    r = GaussianRenderer(cams, st).to(dev)
    best = 1e9
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        frag = r(gm)
        ((to_white_background(frag, col) - tgt) ** 2).mean().backward()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
        gm.zero_grad(); col.grad = None
    ms.append(best); hits.append(int(frag.valid_num.sum()))
t = torch.tensor(ms)
print("per-view ms: mean %.3f min %.3f max %.3f" % (t.mean(), t.min(), t.max()))
print(" ".join("%.2f" % v for v in ms))
print("hits (M):", " ".join("%.1f" % (h / 1e6) for h in hits))
for name, parts in (("contiguous", [list(range(8 * r, 8 * r + 8)) for r in range(8)]),
                    ("round-robin", [list(range(r, 64, 8)) for r in range(8)])):
    loads = [sum(ms[i] for i in p) for p in parts]
    print("%-12s rank loads: %s  max/mean %.3f" % (name, " ".join("%.2f" % l for l in loads), max(loads) / (sum(loads) / 8)))
# greedy LPT with equal counts, cost proxy = hits
order = sorted(range(64), key=lambda i: -hits[i])
loads, parts = [0.0] * 8, [[] for _ in range(8)]
for i in order:
    r = min((q for q in range(8) if len(parts[q]) < 8), key=lambda q: loads[q])
    parts[r].append(i); loads[r] += hits[i]
tl = [sum(ms[i] for i in p) for p in parts]
print("LPT by hits  rank loads: %s  max/mean %.3f" % (" ".join("%.2f" % l for l in tl), max(tl) / (sum(tl) / 8)))

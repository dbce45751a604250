"""Quick per-op timing of the fused path on V views of the C5 scene (development aid, not the bench).
Prints ms per view for binning / forward / merge fwd / merge bwd / fused backward and the forward's
device counters."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from voge_b200 import scenes, _C
from voge_b200.Meshes import GaussianMeshes
from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings, to_white_background

dev = "cuda:0"
V = int(os.environ.get("V", 8)); N = int(os.environ.get("N", 1000000)); HW = int(os.environ.get("HW", 1024))
K = int(os.environ.get("K", 20)); ITERS = int(os.environ.get("ITERS", 3))
verts, sig, colors = scenes.synthetic_scene(N, device=dev)
cams = scenes.orbit_cameras(64, image_size=(HW, HW), focal=900.0 * HW / 1024, device=dev, first=0, count=V)
r = GaussianRenderer(cams, GaussianRenderSettings(image_size=(HW, HW), max_assign=K)).to(dev)
gm = GaussianMeshes(verts, sig).to(dev)
colors.requires_grad_(True)
target = torch.rand(V, HW, HW, 3, device=dev)

times = {}
class _KT:
    enabled = True
    def add(self, name, e0, e1):
        times.setdefault(name, []).append((e0, e1))
from voge_b200 import _lib
_lib.kernel_timer = _KT()          # every C-ABI call (= one kernel launch) bracketed by CUDA events

for it in range(ITERS + 1):
    if it == 1:
        times.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    frag = r(gm)
    img = to_white_background(frag, colors)
    ((img - target) ** 2).mean().backward()
    t1.record()
    gm.zero_grad(); colors.grad = None
torch.cuda.synchronize()
_lib.kernel_timer.enabled = False
# device counters of the forward (one extra call outside the timing)
with torch.no_grad():
    import math
    from voge_b200.cameras import camera_params
    from voge_b200.fused import choose_tile
    from voge_b200.RayTracing import default_bin_size
    rays, origins = r._rays((HW, HW))
    R, T, focal, principal = camera_params(cams, (HW, HW))
    thr_act = -math.log(0.01 + 1e-10)
    bs = default_bin_size((HW, HW)); tile = choose_tile(bs, K, True)
    off, tl, rects, ioff = _C.bin_views(verts, sig, R, T, origins, focal, principal, (HW, HW), 0.01, thr_act, True, bs, tile)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    out = _C.render_forward(verts, sig, origins, rays, off, tl, rects, thr_act, 1.0, K, tile, need_act=False, stats=stats,
                            item_offsets=ioff)
    torch.cuda.synchronize()
    st = stats.tolist()
    print("per view: tile entries %.3fM  items %.2fM (alloc %.2fM)  exact-select pixels %d  hits %.2fM" % (
        tl.shape[0] / V / 1e6, st[0] / V / 1e6, (ioff.total_items - getattr(ioff, 'slack_items', 0)) / V / 1e6, st[2] // V, int((out[0] >= 0).sum()) / V / 1e6))
    if os.environ.get("TILESTATS"):
        # hits per (Gaussian, tile): how much a per-tile Gaussian-major gradient reduction could aggregate
        idx = out[0]
        ys = torch.arange(HW, device=dev).view(1, HW, 1, 1)
        xs = torch.arange(HW, device=dev).view(1, 1, HW, 1)
        bb = torch.arange(V, device=dev).view(V, 1, 1, 1)
        ok = idx >= 0
        n_hits = int(ok.sum())
        for t in (4, 8, 16):
            tiles = (bb * ((HW + t - 1) // t) + ys // t) * ((HW + t - 1) // t) + xs // t
            key = (tiles.expand_as(idx)[ok].long() << 32) | idx[ok].long()
            uniq = torch.unique(key).numel()
            print("  tile %2d: %.2f hits per (Gaussian, tile) [%d hits, %d groups]" % (t, n_hits / max(uniq, 1), n_hits, uniq))
        del key, tiles
tot = 0.0
for nm, ev in times.items():
    ms = sum(a.elapsed_time(b) for a, b in ev) / ITERS / V
    tot += ms
    print("%-24s %.4f ms/view" % (nm, ms))
print("sum %.4f ms/view -> %.1f Mrays/s ; last step wall %.3f ms/view" % (tot, HW * HW / tot / 1e3, t0.elapsed_time(t1) / V))

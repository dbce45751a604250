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
def wrap(name):
    fn = getattr(_C, name)
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(*a, **k); e1.record()
        times.setdefault(name, []).append((e0, e1))
        return out
    setattr(_C, name, w)
for nm in ("bin_views", "render_forward", "merge_final_forward", "merge_final_backward", "render_backward_fused"):
    if hasattr(_C, nm):
        wrap(nm)

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
        tl.numel() / V / 1e6, st[0] / V / 1e6, ioff.total_items / V / 1e6, st[2] // V, int((out[0] >= 0).sum()) / V / 1e6))
    if os.environ.get("XCHECK"):
        ref = _C.render_forward(verts, sig, origins, rays, off, tl, rects, thr_act, 1.0, K, tile, need_act=False)
        torch.cuda.synchronize()
        for nm, x, y in zip(("idx", "weight", "len", "valid"), out, ref):
            print("  xcheck", nm, "equal" if torch.equal(x, y) else "DIFFERENT (%d, max rel %.2e)" % (
                int((x != y).sum()), float(((x - y).abs() / y.abs().clamp_min(1e-30)).max())))
tot = 0.0
for nm, ev in times.items():
    ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev) / V
    tot += ms
    print("%-24s %.4f ms/view" % (nm, ms))
print("sum %.4f ms/view -> %.1f Mrays/s ; last step wall %.3f ms/view" % (tot, HW * HW / tot / 1e3, t0.elapsed_time(t1) / V))

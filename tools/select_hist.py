"""Distribution of the stored hits per pixel / per warp of select_topk on V views of the C5 scene
(development aid: which sorting-network size the warps of select_topk_kernel take)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from voge_b200 import scenes, _C
from voge_b200.cameras import camera_params, generate_rays
from voge_b200.fused import choose_tile
from voge_b200.RayTracing import default_bin_size

dev = "cuda:0"
V = int(os.environ.get("V", 4)); N = int(os.environ.get("N", 1000000)); HW = int(os.environ.get("HW", 1024)); K = 20
verts, sig, colors = scenes.synthetic_scene(N, device=dev)
cams = scenes.orbit_cameras(64, image_size=(HW, HW), focal=900.0 * HW / 1024, device=dev, first=0, count=V)
rays, origins = generate_rays(cams, (HW, HW))
R, T, focal, principal = camera_params(cams, (HW, HW))
thr_act = -math.log(0.01 + 1e-10)
bs = default_bin_size((HW, HW)); tile = choose_tile(bs, K, True)
off, tl, rects, ioff = _C.bin_views(verts, sig, R, T, origins, focal, principal, (HW, HW), 0.01, thr_act, True, bs, tile)
dbg = {}
_C.render_forward(verts, sig, origins, rays, off, tl, rects, thr_act, 1.0, K, tile, need_act=False, item_offsets=ioff, debug=dbg)
c = dbg["counts"].view(-1, 32).to(torch.int64)     # a warp of select_topk = 32 consecutive columns of the table
wmax = c.max(dim=1).values
tot = wmax.numel()
print("pixels: mean hits %.2f  p50 %d  p90 %d  p99 %d  max %d" % (
    c.float().mean(), *(int(torch.quantile(c.float().flatten()[:: 7], q)) for q in (0.5, 0.9, 0.99)), int(c.max())))
for lo, hi in ((0, 0), (1, 16), (17, 32), (33, 48), (49, 64), (65, 10 ** 9)):
    n = int(((wmax >= lo) & (wmax <= hi)).sum())
    print("warps with max hits in [%d, %s]: %.2f %%" % (lo, hi if hi < 10 ** 9 else "inf", 100.0 * n / tot))

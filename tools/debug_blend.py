"""Debug aid: pipeline vs one-launch kernel weights, and depth-window statistics of the C5 scene."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from voge_b200 import scenes, _C
from voge_b200.cameras import camera_params
from voge_b200.fused import choose_tile
from voge_b200.RayTracing import default_bin_size
from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings
dev = "cuda:0"
V = 2; N = 1000000; HW = 1024; K = 20
verts, sig, colors = scenes.synthetic_scene(N, device=dev)
cams = scenes.orbit_cameras(64, image_size=(HW, HW), focal=900.0, device=dev, first=0, count=V)
r = GaussianRenderer(cams, GaussianRenderSettings(image_size=(HW, HW), max_assign=K)).to(dev)
rays, origins = r._rays((HW, HW))
R, T, focal, principal = camera_params(cams, (HW, HW))
thr_act = -math.log(0.01 + 1e-10)
bs = default_bin_size((HW, HW)); tile = choose_tile(bs, K, True)
off, tl, rects, ioff = _C.bin_views(verts, sig, R, T, origins, focal, principal, (HW, HW), 0.01, thr_act, True, bs, tile)
p = _C.render_forward(verts, sig, origins, rays, off, tl, rects, thr_act, 1.0, K, tile, need_act=True, item_offsets=ioff)
s = _C.render_forward(verts, sig, origins, rays, off, tl, rects, thr_act, 1.0, K, tile, need_act=True)
torch.cuda.synchronize()
idx, w, ln, valid, act, dsd = p
w2 = s[1]
d = (w - w2).abs()
print("weights differ at", int((w != w2).sum()), "max abs", float(d.max()), "max rel", float((d / w2.abs().clamp_min(1e-20)).max()))
bad = (w != w2).any(-1).nonzero()
for q in bad[:3]:
    b, y, x = q.tolist()
    print("pixel", b, y, x, "cnt", int(valid[b, y, x]))
    print(" w ", w[b, y, x].tolist()); print(" w2", w2[b, y, x].tolist()); print(" len", ln[b, y, x].tolist())
# window statistics
sk = torch.sqrt(dsd + 1e-10)
ok = idx >= 0
smin = torch.where(ok, sk, torch.full_like(sk, 3e38)).min(-1).values
dl = (ln[..., :, None] - ln[..., None, :]).abs() * smin[..., None, None]
pair = (dl < 4.0) & ok[..., :, None] & ok[..., None, :]
npair = pair.sum((-1, -2)) - valid          # ordered in-window pairs without the diagonal
cnt = valid
print("pixels", cnt.numel(), "mean cnt", float(cnt.float().mean()), "mean ordered neighbour pairs", float(npair.float().mean()))
foot = cnt > 0
print("in footprint: mean cnt %.2f  pairs %.2f  p50 %d p90 %d p99 %d max %d" % (
    float(cnt[foot].float().mean()), float(npair[foot].float().mean()),
    int(npair[foot].float().quantile(0.5)), int(npair[foot].float().quantile(0.9)), int(npair[foot].float().quantile(0.99)), int(npair.max())))
# exact criterion: column k reaches row m iff |l_m - l_k| * s_k < 4
dle = (ln[..., :, None] - ln[..., None, :]).abs() * sk[..., None, :]        # [m, k] * s_k
paire = (dle < 4.0) & ok[..., :, None] & ok[..., None, :]
npe = paire.sum((-1, -2)) - valid
print("exact-reach ordered pairs: mean %.2f in footprint %.2f" % (float(npe.float().mean()), float(npe[foot].float().mean())))
wsz_e = paire.sum(-2)                        # per column k: rows in reach (incl self)
blk_e = wsz_e.view(V, HW // 4, 4, HW // 8, 8, K).permute(0, 1, 3, 2, 4, 5).reshape(V, HW // 4, HW // 8, 32, K)
print("per warp (column-major, exact reach): sum_k max_lanes(window incl self) %.1f" % float(blk_e.max(3).values.sum(-1).float().mean()))
# warp-level (8x4 blocks): max over lanes of per-slot window size summed over slots vs max over lanes of total
wsz = pair.sum(-1)                          # (B,H,W,K) window size incl self
B = V
blk = wsz.view(B, HW // 4, 4, HW // 8, 8, K).permute(0, 1, 3, 2, 4, 5).reshape(B, HW // 4, HW // 8, 32, K)
per_slot_max = blk.max(3).values.sum(-1).float()
tot_max = blk.sum(-1).max(3).values.float()
print("per warp: sum_m max_lanes(window) %.1f ; max_lanes(total pairs incl self) %.1f ; mean lanes total %.1f" % (
    float(per_slot_max.mean()), float(tot_max.mean()), float(blk.sum(-1).float().mean())))

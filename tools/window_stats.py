"""Development aid (CPU only): replay of the window loops of blend_pair_kernel / render_bwd_pair_kernel on a band of
view 0 of the C5 scene (hits from the CPU oracle): warp trip counts of the current scheme (conservative window by the
pixel's s_min, per-slot serial loops, 16 pixels x 2 lanes per warp) against an own-reach scheme (every slot visits the
neighbours within ITS reach 4 / s and scatters the mirrored term)."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import voge_oracle as oracle
from voge_b200 import scenes
from voge_b200.cameras import PerspectiveCameras, generate_rays

N, HW, K = 1_000_000, 1024, 20
y0, rows = int(os.environ.get("Y0", 480)), int(os.environ.get("ROWS", 32))
cache = "/tmp/window_band_%d_%d.npz" % (y0, rows)
if os.path.exists(cache):
    z = np.load(cache); o_len, o_act, o_dsd = z["len"], z["act"], z["dsd"]
else:
    t0 = time.time()
    verts, sig, col = scenes.synthetic_scene(N, seed=0)
    R, T = oracle.look_at_view(3.0, 0.0, 0.0)
    cams = PerspectiveCameras(focal_length=900.0, principal_point=((HW / 2, HW / 2),), R=R, T=T, in_ndc=False, image_size=((HW, HW),))
    rays, origins = generate_rays(cams, (HW, HW))
    bs = oracle.default_bin_size((HW, HW))
    thr_act = -math.log(0.01 + 1e-10)
    mus = (verts[None] - origins[:, None]); isg = (2 * sig)[None]
    ndc, radii = oracle.coarse_inputs(R, T, 900.0, (HW / 2.0, HW / 2.0), (HW, HW), mus, isg, 0.01)
    first, nper = torch.zeros(1, dtype=torch.long), torch.full((1,), N)
    bp, bc = oracle.rasterize_coarse(ndc.reshape(-1, 3), radii.reshape(-1, 2), first, nper, (HW, HW), bs, 8192)
    bp_sub = torch.from_numpy(bp[:, y0 // bs:(y0 + rows) // bs, :, :int(bc.max())].copy())
    rays_sub = rays[:, y0:y0 + rows].contiguous()
    o_idx, o_len, o_act, o_dsd = oracle.ray_trace_fine(mus.reshape(-1, 3), isg.reshape(-1, 3, 3), rays_sub, bp_sub, thr_act, bs, K)
    np.savez(cache, len=o_len, act=o_act, dsd=o_dsd)
    print("oracle band in %.1f s" % (time.time() - t0))
L = o_len[0]; A = o_act[0]; D = o_dsd[0]              # (rows, W, K)
valid = L < 1e9
cnt = valid.sum(-1)
S = np.sqrt(D + 1e-10)
print("pixels %d, mean hits %.2f, full pixels %.1f%%" % (cnt.size, cnt.mean(), 100 * (cnt == K).mean()))
SAT = 4.0
tot_cur_trips = tot_cur_a = tot_cur_b = 0
tot_new_trips_lr = tot_new_trips_max = tot_new_evals = 0
tot_cons_pairs = tot_exact_pairs = 0
warps = 0
cur_cost = new_cost = 0.0
fwd_cur = fwd_new = 0.0
for by in range(0, rows, 4):
    for bx in range(0, HW, 4):
        warps += 1
        # lanes: 16 pixels x 2 subs; per step jj every lane works on slot j = 2 jj + sub
        lane_info = []
        for p in range(16):
            y, x = by + p // 4, bx + p % 4
            c = int(cnt[y, x]); l = L[y, x, :c].astype(np.float64); s = S[y, x, :c].astype(np.float64)
            smin = s.min() if c else 1.0
            lane_info.append((c, l, s, smin))
        for jj in range(K // 2):
            trips = []; a_any = []; new_l = []; new_r = []
            for (c, l, s, smin) in lane_info:
                for sub in range(2):
                    j = 2 * jj + sub
                    if j >= c:
                        continue
                    dl = l - l[j]
                    inwin = (np.abs(dl) * smin < SAT); inwin[j] = False
                    n_tr = int(inwin.sum())
                    a_un = (np.abs(dl) * s[j] < SAT) & inwin          # own reach (term a)
                    b_un = (np.abs(dl) * s < SAT) & inwin             # neighbour's reach (term b)
                    trips.append(n_tr)
                    a_any.append((int(a_un.sum()), int(b_un.sum())))
                    own = (np.abs(dl) * s[j] < SAT); own[j] = False
                    new_l.append(int(own[:j].sum())); new_r.append(int(own[j + 1:].sum()))
                    tot_cons_pairs += n_tr; tot_exact_pairs += int(own.sum())
            if not trips:
                continue
            mt = max(trips)
            tot_cur_trips += mt
            tot_cur_a += sum(a for a, b in a_any); tot_cur_b += sum(b for a, b in a_any)
            # cost model (warp instructions): current backward trip ~ 14 + 32 (a unsaturated somewhere) + 16 (b)
            cur_cost += mt * (14 + 32 + 16)
            fwd_cur += mt * 30
            ml, mr = max(new_l), max(new_r)
            tot_new_trips_lr += ml + mr
            tot_new_trips_max += max(max(a, b) for a, b in zip(new_l, new_r))
            tot_new_evals += sum(new_l) + sum(new_r)
            new_cost += (ml + mr) * (14 + 32 + 10)
            fwd_new += (ml + mr) * (30 + 8)
print("warps %d" % warps)
print("conservative ordered pairs / pixel %.2f, own-reach ordered pairs / pixel %.2f" % (tot_cons_pairs / cnt.size, tot_exact_pairs / cnt.size))
print("current : warp trips per warp %.1f ; expensive a %.1f b %.1f per warp" % (tot_cur_trips / warps, tot_cur_a / warps, tot_cur_b / warps))
print("own-reach: warp trips per warp (left loop + right loop) %.1f ; (both sides per trip) %.1f ; evals per warp %.1f" % (tot_new_trips_lr / warps, tot_new_trips_max / warps, tot_new_evals / warps))
print("cost model, backward window loops: current %.0f -> own-reach %.0f warp instr per warp" % (cur_cost / warps, new_cost / warps))
print("cost model, forward window loops : current %.0f -> own-reach %.0f" % (fwd_cur / warps, fwd_new / warps))

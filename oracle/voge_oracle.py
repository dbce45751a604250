"""TEST INFRASTRUCTURE, NOT PRODUCT CODE -- CPU oracle for the VoGE ray-tracing hot path.

Two layers:
  * ctypes wrappers over oracle/libvoge_oracle.so (plain-C restatement of the reference CUDA
    kernels with the reference binary's fp32 rounding sequence, see voge_oracle.c);
  * PyTorch (CPU) transcriptions of the reference's PyTorch-level maths: Aggregation.py
    (blend weights, merge_final), Renderer.py (host glue, background composite), RayTracing.py
    (bin-size heuristics, bbox maths) and the pytorch3d camera semantics they rely on; plus
    `ray_trace_fine_torch`, the line-for-line vectorised PyTorch transcription of
    RayTraceFineVogeKernel that BASELINE.md names as the CPU baseline (the reference ships no
    CPU ray tracer: ray_trace_voge.h:28-30).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Parity pinning: see voge_oracle.c header and tests/golden/README.md.
Citations are relative to /root/reference.
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvoge_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "voge_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.vo_count_pairs.restype = ctypes.c_int64
        _lib.vo_num_threads.restype = ctypes.c_int
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _l(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _np(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


# ------------------------------------------------------------------------------------------------
# C-oracle wrappers (numpy in / numpy out)
# ------------------------------------------------------------------------------------------------
def pair(mu, S, d):
    """(len, act, dsd) of one pair, ray_trace_voge.cu:188-193."""
    mu, S, d = _f(_np(mu)), _f(_np(S)), _f(_np(d))
    out = [ctypes.c_float() for _ in range(3)]
    lib().vo_pair(_p(mu), _p(S), _p(d), *[ctypes.byref(o) for o in out])
    return tuple(np.float32(o.value) for o in out)


def rasterize_coarse(points_ndc, radius, first_idx, num_per, image_size, bin_size, M):
    """-> (bin_points (B,BH,BW,M) i32, bin_counts (B,BH,BW) i32); rasterize_coarse.cu:20-188."""
    pts, rad = _f(_np(points_ndc)).reshape(-1, 3), _f(_np(radius)).reshape(-1, 2)
    first, nper = _l(_np(first_idx)), _l(_np(num_per))
    H, W = int(image_size[0]), int(image_size[1])
    B, P = nper.shape[0], pts.shape[0]
    BH, BW = 1 + (H - 1) // bin_size, 1 + (W - 1) // bin_size
    bp = np.empty((B, BH, BW, M), np.int32)
    bc = np.empty((B, BH, BW), np.int32)
    lib().vo_rasterize_coarse(_p(pts), _p(rad), _p(first), _p(nper), B, P, H, W, int(bin_size), int(M), _p(bp), _p(bc))
    return bp, bc


def ray_trace_fine(mus, isigmas, rays, bin_points, thr_act, bin_size, K):
    """-> (idx i32, len, act, dsd) each (B,H,W,K); ray_trace_voge.cu:135-217."""
    mus, isg, rays = _f(_np(mus)).reshape(-1, 3), _f(_np(isigmas)).reshape(-1, 9), _f(_np(rays))
    bp = _i(_np(bin_points))
    B, H, W, _ = rays.shape
    _, BH, BW, M = bp.shape
    idx = np.empty((B, H, W, K), np.int32)
    tl = np.empty((B, H, W, K), np.float32)
    ta = np.empty((B, H, W, K), np.float32)
    td = np.empty((B, H, W, K), np.float32)
    lib().vo_ray_trace_fine(_p(mus), _p(isg), _p(rays), _p(bp), ctypes.c_float(thr_act), int(bin_size), B, H, W, BH,
                            BW, M, int(K), _p(idx), _p(tl), _p(ta), _p(td))
    return idx, tl, ta, td


def count_pairs(bin_points, bin_size, H, W):
    bp = _i(_np(bin_points))
    B, BH, BW, M = bp.shape
    return int(lib().vo_count_pairs(_p(bp), int(bin_size), B, int(H), int(W), BH, BW, M))


def ray_trace_fine_backward(mus, isigmas, rays, idx, g_len, g_act, g_dsd):
    """-> (grad_rays (B,H,W,3), grad_mus (P,3), grad_isg (P,3,3)); ray_trace_voge.cu:283-332."""
    mus, isg, rays = _f(_np(mus)).reshape(-1, 3), _f(_np(isigmas)).reshape(-1, 9), _f(_np(rays))
    idx = _i(_np(idx))
    gl, ga, gd = _f(_np(g_len)), _f(_np(g_act)), _f(_np(g_dsd))
    B, H, W, K = idx.shape
    P = mus.shape[0]
    gr = np.empty((B, H, W, 3), np.float32)
    gm = np.empty((P, 3), np.float32)
    gs = np.empty((P, 3, 3), np.float32)
    lib().vo_ray_trace_fine_backward(_p(mus), _p(isg), _p(rays), _p(idx), _p(gl), _p(ga), _p(gd), B, H, W, K, P,
                                     _p(gr), _p(gm), _p(gs))
    return gr, gm, gs


def sample(image, weight, idx, num_vert):
    image, weight, idx = _f(_np(image)), _f(_np(weight)), _i(_np(idx))
    C, K = image.shape[-1], idx.shape[-1]
    R = idx.size // K
    feat = np.empty((num_vert, C), np.float32)
    wsum = np.empty((num_vert,), np.float32)
    lib().vo_sample(_p(image), _p(weight), _p(idx), ctypes.c_int64(R), K, C, int(num_vert), _p(feat), _p(wsum))
    return feat, wsum


def sample_backward(image, weight, idx, g_feat, g_wsum):
    image, weight, idx = _f(_np(image)), _f(_np(weight)), _i(_np(idx))
    gf, gw = _f(_np(g_feat)), _f(_np(g_wsum))
    C, K = image.shape[-1], idx.shape[-1]
    R = idx.size // K
    gi = np.empty_like(image)
    gwt = np.empty_like(weight)
    lib().vo_sample_backward(_p(image), _p(weight), _p(idx), _p(gf), _p(gw), ctypes.c_int64(R), K, C, _p(gi), _p(gwt))
    return gi, gwt


def scatter_max(weight, idx, num_vert):
    weight, idx = _f(_np(weight)), _i(_np(idx))
    K = idx.shape[-1]
    R = idx.size // K
    out = np.empty((num_vert,), np.float32)
    lib().vo_scatter_max(_p(weight), _p(idx), ctypes.c_int64(R), K, int(num_vert), _p(out))
    return out


def ray_trace_ray(mus, isigmas, rays):
    """-> (len, act, dsd) each (N,M); voge_ray_tracing_ray.cu:114-143."""
    mus, isg, rays = _f(_np(mus)).reshape(-1, 3), _f(_np(isigmas)).reshape(-1, 9), _f(_np(rays)).reshape(-1, 3)
    M, N = mus.shape[0], rays.shape[0]
    out = [np.empty((N, M), np.float32) for _ in range(3)]
    lib().vo_ray_trace_ray(_p(mus), _p(isg), _p(rays), M, N, _p(out[0]), _p(out[1]), _p(out[2]))
    return tuple(out)


def ray_trace_ray_backward(mus, isigmas, rays, g_len, g_act, g_dsd):
    """Dense backward (voge_ray_tracing_ray.cu:147-188) through the fine-backward restatement: the (N,M) table
    is a (1,N,1,M) hit list whose slot m holds Gaussian m."""
    mus, rays = _f(_np(mus)).reshape(-1, 3), _f(_np(rays)).reshape(-1, 3)
    M, N = mus.shape[0], rays.shape[0]
    idx = np.broadcast_to(np.arange(M, dtype=np.int32), (1, N, 1, M)).copy()
    shp = (1, N, 1, M)
    gr, gm, gs = ray_trace_fine_backward(mus, isigmas, rays.reshape(1, N, 1, 3), idx, _f(_np(g_len)).reshape(shp),
                                         _f(_np(g_act)).reshape(shp), _f(_np(g_dsd)).reshape(shp))
    return gr.reshape(N, 3), gm, gs


def find_nearest_k(len_in, act_in, dsd_in, thr_act, K):
    """-> (idx i32, len, act, dsd) each (N,K); voge_ray_tracing_ray.cu:191-239."""
    ln, ac, ds = _f(_np(len_in)), _f(_np(act_in)), _f(_np(dsd_in))
    N, M = ln.shape
    idx = np.empty((N, K), np.int32)
    out = [np.empty((N, K), np.float32) for _ in range(3)]
    lib().vo_find_nearest_k(_p(ln), _p(ac), _p(ds), ctypes.c_float(thr_act), M, int(K), N, _p(idx), _p(out[0]), _p(out[1]),
                            _p(out[2]))
    return idx, out[0], out[1], out[2]


def num_threads():
    return int(lib().vo_num_threads())


# ------------------------------------------------------------------------------------------------
# PyTorch transcriptions (CPU, differentiable) of the reference's PyTorch-level maths
# ------------------------------------------------------------------------------------------------
def aggregation_torch(sel_idx, sel_act, sel_len, sel_dsd, occupation_weight=1.0):
    """Aggregation.py:30-107 (get_cross_activation, assign2weight, aggregation)."""
    K = sel_idx.shape[-1]
    shape = sel_idx.shape[:-1]
    ln, ds, ac = sel_len.reshape(-1, K), sel_dsd.reshape(-1, K), sel_act.reshape(-1, K)
    cross = (ln.unsqueeze(2) - ln.unsqueeze(1)) * (ds.view(-1, 1, K) + 1e-10).pow(.5)            # :49
    density = torch.exp(-ac.unsqueeze(1)) * ((torch.erf(cross) + 1) / 2)                           # :70
    dweight = torch.exp(-(torch.sum(density, dim=2)) * occupation_weight)                          # :74
    weight = dweight * torch.exp(-ac) / math.exp(-0.5)                                             # :77-79
    valid_num = torch.sum(sel_idx >= 0, dim=-1)                                                    # :104
    return weight.view(*shape, K), sel_idx, valid_num, sel_len


def merge_final_torch(vert_attr, weight, vert_assign, valid_num):
    """Aggregation.py:111-141 without the in-place index mutation (works on a copy)."""
    K = weight.shape[-1]
    ar = torch.arange(K, device=weight.device).view(*([1] * valid_num.dim()), K)
    mask = (ar < valid_num.unsqueeze(-1)).to(weight.dtype)                                         # :125-129
    idx = vert_assign.clone().long()
    idx = idx + (idx < 0) * 1                                                                      # :131
    sel_attr = vert_attr[idx]                                                                      # :137
    return torch.sum(sel_attr * (mask * weight).unsqueeze(-1), dim=-2)                             # :134,:140


def get_silhouette_torch(weight):
    merged = weight.sum(-1)
    return torch.min(merged, torch.ones_like(merged))                                              # Renderer.py:157-159


def to_colored_background_torch(weight, idx, valid_num, colors, background=(1, 1, 1), thr=-1):
    """Renderer.py:162-171."""
    masks = get_silhouette_torch(weight).unsqueeze(-1)
    bg = torch.tensor(list(background), dtype=colors.dtype, device=colors.device) if not torch.is_tensor(background) else background
    if thr > 0:
        masks = (masks > thr).type_as(masks)
    rgb = merge_final_torch(colors, weight, idx, valid_num)
    return torch.min(rgb + torch.ones_like(rgb) * (1 - masks) * bg, torch.ones_like(rgb))


def expend_sigma_torch(sigma):
    """Aggregation.py:144-175 with the default identity rotation."""
    if sigma.dim() == 3:
        return sigma
    eye = torch.eye(3, device=sigma.device, dtype=sigma.dtype)[None]
    if sigma.dim() == 1:
        return sigma.view(-1, 1, 1) * eye
    return sigma.unsqueeze(2) * eye


def default_bin_size(image_size):
    return max(int(2 ** np.ceil(np.log2(max(image_size)) - 5)), 10)                                # RayTracing.py:14-16


def default_max_points_per_bin(n_assign, n_points):
    return min(int(max(n_assign * 10, n_points / 10)), n_points)                                   # RayTracing.py:18-19


# ---- pytorch3d camera semantics, restated in closed form (float64 available via dtype) ----------
def look_at_view(dist, elev, azim, dtype=torch.float32):
    """pytorch3d look_at_view_transform (degrees): C = d(cos e sin a, sin e, cos e cos a);
    z = normalize(-C), x = normalize(up x z), y = normalize(z x x); R = [x y z] columns; T = -R^T C."""
    dist, elev, azim = (torch.as_tensor(v, dtype=dtype).reshape(-1) for v in (dist, elev, azim))
    n = max(dist.numel(), elev.numel(), azim.numel())
    dist, elev, azim = dist.expand(n), elev.expand(n) * math.pi / 180, azim.expand(n) * math.pi / 180
    C = torch.stack([dist * torch.cos(elev) * torch.sin(azim), dist * torch.sin(elev),
                     dist * torch.cos(elev) * torch.cos(azim)], dim=1)
    up = torch.tensor([[0., 1., 0.]], dtype=dtype).expand(n, 3)
    z = torch.nn.functional.normalize(-C, dim=1)
    x = torch.nn.functional.normalize(torch.cross(up, z, dim=1), dim=1)
    y = torch.nn.functional.normalize(torch.cross(z, x, dim=1), dim=1)
    R = torch.stack([x, y, z], dim=2)
    T = -torch.bmm(R.transpose(1, 2), C[:, :, None])[:, :, 0]
    return R, T


def camera_rays(R, T, focal, principal, image_size):
    """Rays of a screen-space PerspectiveCameras through pixel centres:
    d_cam ~ (-(x+.5-px)/fx, -(y+.5-py)/fy, 1) normalised; d_world = d_cam @ R^T; origin = -T @ R^T."""
    H, W = image_size
    dt = R.dtype
    B = R.shape[0]
    f = torch.as_tensor(focal, dtype=dt).reshape(-1, 1) * torch.ones(1, 2, dtype=dt) if not (torch.is_tensor(focal) and focal.dim() == 2) else focal.to(dt)
    p = torch.as_tensor(principal, dtype=dt).reshape(-1, 2)
    f, p = f.expand(B, 2), p.expand(B, 2)
    xs = torch.arange(W, dtype=dt) + 0.5
    ys = torch.arange(H, dtype=dt) + 0.5
    dx = (-(xs[None, None, :] - p[:, 0, None, None]) / f[:, 0, None, None]).expand(B, H, W)
    dy = (-(ys[None, :, None] - p[:, 1, None, None]) / f[:, 1, None, None]).expand(B, H, W)
    d = torch.nn.functional.normalize(torch.stack([dx, dy, torch.ones(B, H, W, dtype=dt)], -1), dim=-1)
    dw = torch.matmul(d.view(B, -1, 3), R.transpose(1, 2)).view(B, H, W, 3)
    origin = -torch.matmul(T[:, None, :], R.transpose(1, 2))[:, 0, :]
    return dw, origin


def coarse_inputs(R, T, focal, principal, image_size, points_centred, isigmas, thr):
    """RayTracing.py:45-57 in closed form: flipped NDC centre ((x_screen - W/2)/s, (y_screen - H/2)/s),
    view depth, and bbox radii sqrt(-ln thr * colsum(F inv(S_view[:2,:2]) F)) / z_view."""
    H, W = image_size
    dt = R.dtype
    B = R.shape[0]
    f = torch.as_tensor(focal, dtype=dt).reshape(-1, 1) * torch.ones(1, 2, dtype=dt) if not (torch.is_tensor(focal) and focal.dim() == 2) else focal.to(dt)
    p = torch.as_tensor(principal, dtype=dt).reshape(-1, 2)
    f, p = f.expand(B, 2), p.expand(B, 2)
    s = min(H, W) / 2.0
    C = -torch.matmul(T[:, None, :], R.transpose(1, 2))          # (B,1,3)
    world = points_centred + C
    view = torch.matmul(world, R) + T[:, None, :]
    z = view[..., 2]
    xs = p[:, 0, None] - f[:, 0, None] * view[..., 0] / z        # +x right pixel coordinate
    ys = p[:, 1, None] - f[:, 1, None] * view[..., 1] / z
    ndc = torch.stack([(xs - W / 2.0) / s, (ys - H / 2.0) / s, z], dim=-1)
    S_view = R.transpose(1, 2)[:, None] @ isigmas @ R[:, None]
    Fm = torch.diag_embed(f / s)[:, None]                        # (B,1,2,2)
    get = -math.log(thr) * Fm @ torch.inverse(S_view[..., :2, :2]) @ Fm
    radii = get.sum(dim=-2).pow(.5) * (1.0 / z).unsqueeze(-1)
    return ndc, radii


def ray_trace_fine_torch(mus, isigmas, rays, bin_points, thr_act, bin_size, K, chunk_pixels=4096):
    """Line-for-line *vectorised* PyTorch transcription of RayTraceFineVogeKernel
    (ray_trace_voge.cu:135-217): per bin, the three quadratic forms for every (pixel, candidate)
    pair, threshold, and the K smallest hit lengths in ascending order.  fp32, default torch
    rounding (NOT the reference's contracted sequence: use the C oracle for bit-level checks).
    This is the CPU baseline of BASELINE.md (`REF-CPU`)."""
    B, H, W, _ = rays.shape
    _, BH, BW, M = bin_points.shape
    idx = torch.full((B, H, W, K), -1, dtype=torch.int32)
    tl = torch.full((B, H, W, K), 1e10)
    ta = torch.full((B, H, W, K), 1e10)
    td = torch.zeros((B, H, W, K))
    for b in range(B):
        for by in range(BH):
            for bx in range(BW):
                cand = bin_points[b, by, bx]
                cand = cand[cand > -1].long()
                if cand.numel() == 0:
                    continue
                y0, y1 = by * bin_size, min((by + 1) * bin_size, H)
                x0, x1 = bx * bin_size, min((bx + 1) * bin_size, W)
                d = rays[b, y0:y1, x0:x1].reshape(-1, 3)                       # (p,3)
                mu, S = mus[cand], isigmas[cand]                               # (m,3), (m,3,3)
                Sd = torch.einsum('mij,pj->pmi', S, d)                         # (p,m,3)
                ksk = torch.einsum('pi,pmi->pm', d, Sd)
                msk = torch.einsum('mi,pmi->pm', mu, Sd)
                msm = torch.einsum('mi,mij,mj->m', mu, S, mu)[None]
                ln = msk / ksk
                act = msm - msk * msk / ksk
                ln_m = torch.where(act < thr_act, ln, torch.full_like(ln, 1e10))
                kk = min(K, cand.numel())
                top, order = torch.topk(ln_m, kk, dim=1, largest=False, sorted=True)
                ok = top < 1e10
                g_act = torch.gather(act, 1, order)
                g_dsd = torch.gather(ksk, 1, order)
                p = d.shape[0]
                blk_idx = torch.full((p, K), -1, dtype=torch.int32)
                blk_len = torch.full((p, K), 1e10)
                blk_act = torch.full((p, K), 1e10)
                blk_dsd = torch.zeros((p, K))
                blk_idx[:, :kk] = torch.where(ok, cand[order].int(), torch.full_like(order, -1).int())
                blk_len[:, :kk] = torch.where(ok, top, torch.full_like(top, 1e10))
                blk_act[:, :kk] = torch.where(ok, g_act, torch.full_like(top, 1e10))
                blk_dsd[:, :kk] = torch.where(ok, g_dsd, torch.zeros_like(top))
                hh, ww = y1 - y0, x1 - x0
                idx[b, y0:y1, x0:x1] = blk_idx.view(hh, ww, K)
                tl[b, y0:y1, x0:x1] = blk_len.view(hh, ww, K)
                ta[b, y0:y1, x0:x1] = blk_act.view(hh, ww, K)
                td[b, y0:y1, x0:x1] = blk_dsd.view(hh, ww, K)
    return idx, tl, ta, td


def render_reference_cpu(verts, sigmas, R, T, focal, principal, image_size, K=20, thr=0.01, absorptivity=1.0,
                         max_points_per_bin=None, bin_size=None, use_c=True, rays=None, origin=None):
    """GaussianRenderer.forward (Renderer.py:102-150) end to end on the CPU for the closed-form
    camera: returns dict(weight, idx, valid_num, len, act, dsd, rays, bin_points, mus, isigmas)."""
    H, W = image_size
    verts = verts.detach().float().cpu()
    S_in = expend_sigma_torch(sigmas.detach().float().cpu())
    R, T = R.float().cpu(), T.float().cpu()
    B, N = R.shape[0], verts.shape[0]
    if rays is None:
        rays, origin = camera_rays(R, T, focal, principal, image_size)
    else:   # rays / origins produced by the device-side generator under test (inputs of the hot path)
        rays, origin = rays.detach().float().cpu().contiguous(), origin.detach().float().cpu()
    mus = verts[None] - origin[:, None]                                                            # :130
    isig = (2 * S_in)[None].expand(B, -1, -1, -1)                                                  # :131-137
    if bin_size is None:
        bin_size = default_bin_size(image_size)
    if max_points_per_bin is None:
        max_points_per_bin = default_max_points_per_bin(K, N)
    if max_points_per_bin == -1:
        BH, BW = 1 + (H - 1) // bin_size, 1 + (W - 1) // bin_size
        bp = (torch.arange(N).view(1, 1, 1, -1) + torch.arange(B).view(-1, 1, 1, 1) * N).expand(-1, BH, BW, -1).int().contiguous()
    else:
        ndc, radii = coarse_inputs(R, T, focal, principal, image_size, mus, isig, thr)
        first = torch.arange(B) * N
        nper = torch.full((B,), N)
        bp_np, bc = rasterize_coarse(ndc.reshape(-1, 3), radii.reshape(-1, 2), first, nper, image_size, bin_size,
                                     max_points_per_bin)
        assert bc.max() <= max_points_per_bin, "bin overflow in oracle: raise max_points_per_bin"
        bp = torch.from_numpy(bp_np)
    thr_act = -math.log(thr + 1e-10)
    mus_p, isig_p = mus.reshape(-1, 3).contiguous(), isig.reshape(-1, 3, 3).contiguous()
    if use_c:
        idx, tl, ta, td = (torch.from_numpy(a) for a in ray_trace_fine(mus_p, isig_p, rays, bp, thr_act, bin_size, K))
    else:
        idx, tl, ta, td = ray_trace_fine_torch(mus_p, isig_p, rays, bp, thr_act, bin_size, K)
    weight, _, valid, _ = aggregation_torch(idx, ta, tl, td, absorptivity)
    return dict(weight=weight, idx=idx, valid_num=valid, len=tl, act=ta, dsd=td, rays=rays, bin_points=bp,
                mus=mus_p, isigmas=isig_p, bin_size=bin_size, thr_act=thr_act)

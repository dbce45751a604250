"""GPU tests of the fused renderer path (voge_pack_gaussians / voge_bin_count / voge_bin_fill / voge_trace_hits /
voge_select_topk / voge_blend_weights / voge_render_backward_fused / voge_unpack_gradients) through the public API: it must reproduce the op-by-op chain
(ray_tracing -> aggregation), which test_gpu_parity.py pins to the reference kernels, bit for bit,
and the CPU oracle run on the same rays."""
import math

import numpy as np
import pytest
import torch

from scene_utils import ambiguous_bbox_gaussians, check_index_rows, small_scene

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(sc, sig_kind="full", K=None, M=None, image_size=None):
    from voge_b200.cameras import PerspectiveCameras
    from voge_b200.Meshes import GaussianMeshes
    from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings
    H, W = image_size or sc["image_size"]
    cams = PerspectiveCameras(focal_length=sc["focal"], principal_point=(sc["principal"],), R=sc["R"], T=sc["T"],
                              in_ndc=False, image_size=((H, W),), device=DEV)
    sig = sc["sigmas"]
    if sig_kind == "iso":
        sig = sig[:, 0, 0].contiguous()
    elif sig_kind == "diag":
        sig = torch.stack([sig[:, 0, 0], sig[:, 1, 1] * 1.3, sig[:, 2, 2] * 0.8], dim=1).contiguous()
    st = GaussianRenderSettings(image_size=(H, W), max_assign=K or sc["K"], max_point_per_bin=M)
    return GaussianRenderer(cams, st).to(DEV), GaussianMeshes(sc["verts"].clone(), sig.clone()).to(DEV)


def _both(renderer, gm):
    renderer.use_fused = True
    a = renderer(gm)
    renderer.use_fused = False
    b = renderer(gm)
    renderer.use_fused = True
    return a, b


@pytest.mark.parametrize("kind,seed,views,M", [("full", 1, 1, None), ("iso", 2, 2, None), ("diag", 3, 1, None),
                                               ("full", 4, 2, -1), ("iso", 5, 1, -1)])
def test_fused_equals_unfused_bitwise(kind, seed, views, M):
    sc = small_scene(seed=seed, aniso=True, views=views, n=400)
    if M is None:
        M = 400
    renderer, gm = _setup(sc, kind, M=M)
    a, b = _both(renderer, gm)
    assert (b.vert_index >= 0).sum() > 1000
    assert torch.equal(a.vert_index, b.vert_index)
    assert torch.equal(a.vert_hit_length, b.vert_hit_length)
    assert torch.equal(a.valid_num, b.valid_num) and a.valid_num.dtype == torch.int64
    # same maths, kernel-local FMA contraction; NaN only where a non-PD S drives exp(-act) to inf in both
    assert torch.allclose(a.vert_weight, b.vert_weight, rtol=1e-5, atol=1e-9, equal_nan=True)
    assert torch.isnan(a.vert_weight).float().mean() < 0.01


@pytest.mark.parametrize("hw,K", [((64, 64), 6), ((96, 80), 30), ((33, 47), 70)])
def test_fused_tile_shapes_and_oracle(oracle, hw, K):
    # 64 -> bin 10/tile 10 ; 96x80 -> bin 10 ; K=70 forces a smaller tile; non-multiple sizes
    sc = small_scene(seed=7, aniso=True, views=2, n=500, image_size=hw, focal=90.0)
    renderer, gm = _setup(sc, "full", K=K, M=500)
    frag = renderer(gm)
    rays, origins = renderer._rays(hw)
    o = oracle.render_reference_cpu(sc["verts"], sc["sigmas"], sc["R"], sc["T"], sc["focal"], sc["principal"], hw, K=K,
                                    max_points_per_bin=500, rays=rays, origin=origins)
    # candidate sets can differ from the oracle's only for a Gaussian whose bbox edge is within rounding of a
    # bin edge (bbox maths is fp32 PyTorch in the oracle, closed form in the kernel): exact assertion elsewhere
    from voge_b200.RayTracing import default_bin_size
    amb = ambiguous_bbox_gaussians(sc["R"], sc["T"], sc["focal"], sc["principal"], hw, sc["verts"], 2 * sc["sigmas"], 0.01,
                                   default_bin_size(hw))
    rows, stats = check_index_rows(frag.vert_index.cpu(), o["idx"], amb, sc["verts"].shape[0], label="%dx%d K=%d" % (hw + (K,)))
    assert stats["unexplained"] == 0 and stats["differing"] <= 2e-3 * stats["rows"]
    assert torch.equal(frag.vert_hit_length.cpu()[rows], o["len"][rows])
    assert torch.allclose(frag.vert_weight.cpu()[rows], o["weight"][rows], rtol=1e-5, atol=1e-7)


def test_fused_gradients_match_unfused_and_oracle(oracle):
    from voge_b200.Renderer import to_white_background
    sc = small_scene(seed=9, aniso=True, views=2, n=300)
    target = torch.rand(2, *sc["image_size"], 3, generator=torch.Generator().manual_seed(1)).to(DEV)
    grads = {}
    for kind in ("full", "iso", "diag"):
        for fused in (True, False):
            renderer, gm = _setup(sc, kind, M=300)
            renderer.use_fused = fused
            colors = sc["colors"].to(DEV).requires_grad_(True)
            frag = renderer(gm)
            img = to_white_background(frag, colors, )
            loss = ((img - target) ** 2).mean() + 1e-3 * frag.vert_hit_length.clamp(max=100).mean()
            loss.backward()
            grads[(kind, fused)] = (gm.verts.grad.clone(), gm.sigmas.grad.clone(), colors.grad.clone())
        for a, b in zip(grads[(kind, True)], grads[(kind, False)]):
            assert torch.isfinite(a).all()
            assert (a - b).abs().max() <= 2e-5 * b.abs().max() + 1e-12, kind
    # oracle autograd (float64 torch transcription of the whole chain) on the selected hits
    renderer, gm = _setup(sc, "full", M=300)
    frag = renderer(gm)
    rays, origins = renderer._rays(sc["image_size"])
    idx = frag.vert_index.cpu().long()
    v = sc["verts"].double().requires_grad_(True)
    S_in = sc["sigmas"].double().requires_grad_(True)
    col = sc["colors"].double().requires_grad_(True)
    B, N = 2, v.shape[0]
    mus = (v[None] - origins.cpu().double()[:, None]).reshape(-1, 3)
    S = (2 * S_in)[None].expand(B, -1, -1, -1).reshape(-1, 3, 3)
    g = idx.clamp(min=0)
    d = rays.cpu().double()[..., None, :].expand(-1, -1, -1, idx.shape[-1], -1)
    Sd = torch.einsum('...ij,...j->...i', S[g], d)
    ksk = (d * Sd).sum(-1); msk = (mus[g] * Sd).sum(-1)
    msm = torch.einsum('...i,...ij,...j->...', mus[g], S[g], mus[g])
    valid = idx >= 0
    ln = torch.where(valid, msk / ksk, torch.full_like(ksk, 1e10))
    act = torch.where(valid, msm - msk * msk / ksk, torch.full_like(ksk, 1e10))
    dsd = torch.where(valid, ksk, torch.zeros_like(ksk))
    w, _, vn, _ = oracle.aggregation_torch(idx.int(), act, ln, dsd, 1.0)
    img = oracle.to_colored_background_torch(w, idx % N, vn, col, (1, 1, 1), -1)
    loss = ((img - target.cpu().double()) ** 2).mean() + 1e-3 * ln.clamp(max=100).mean()
    loss.backward()
    gv, gs, gc = grads[("full", True)]
    for got, want in ((gv, v.grad), (gs, S_in.grad), (gc, col.grad)):
        want = want.float()
        assert (got.cpu() - want).abs().max() <= 1e-4 * want.abs().max()


def test_fused_culling_is_conservative_on_hard_cases():
    # Gaussians near / behind the camera plane, very elongated ones, huge ones covering the image
    g = torch.Generator().manual_seed(3)
    sc = small_scene(seed=11, aniso=True, views=2, n=300, dist=1.2)
    sig = sc["sigmas"]
    sig[:20] *= 0.002                                                # huge blobs (cover everything)
    A = torch.randn(40, 3, 3, generator=g)
    Q, _ = torch.linalg.qr(A)
    sig[20:60] = Q @ torch.diag(torch.tensor([4000.0, 30.0, 3.0])) @ Q.transpose(1, 2)   # needles / discs
    sc["sigmas"] = sig
    renderer, gm = _setup(sc, "full", M=300, K=12)
    a, b = _both(renderer, gm)
    assert torch.equal(a.vert_index, b.vert_index) and torch.equal(a.vert_hit_length, b.vert_hit_length)
    renderer, gm = _setup(sc, "full", M=-1, K=12)
    a, b = _both(renderer, gm)
    assert torch.equal(a.vert_index, b.vert_index) and torch.equal(a.vert_hit_length, b.vert_hit_length)


def test_fused_edge_cases(oracle):
    from voge_b200.cameras import PerspectiveCameras
    from voge_b200.Meshes import GaussianMeshesNaive
    from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings, interpolate_attr, to_white_background
    # (a) a single Gaussian, K = 1, image smaller than a tile, per-view intrinsics
    R, T = oracle.look_at_view(torch.tensor([3.0, 2.5]), torch.tensor([0.0, 30.0]), torch.tensor([0.0, 45.0]))
    cams = PerspectiveCameras(focal_length=torch.tensor([[20.0, 20.0], [30.0, 25.0]]),
                              principal_point=torch.tensor([[6.0, 4.5], [5.0, 5.0]]), R=R, T=T, in_ndc=False,
                              image_size=((9, 12),), device=DEV)
    st = GaussianRenderSettings(image_size=(9, 12), max_assign=1, max_point_per_bin=1)
    r = GaussianRenderer(cams, st).to(DEV)
    gm = GaussianMeshesNaive(torch.zeros(1, 3, device=DEV), torch.tensor([3.0], device=DEV))
    a, b = _both(r, gm)
    assert torch.equal(a.vert_index, b.vert_index) and torch.equal(a.vert_hit_length, b.vert_hit_length)
    assert a.vert_index.shape == (2, 9, 12, 1) and (a.vert_index[1] >= 1).any() and (a.valid_num.sum() > 10)
    img = to_white_background(a, torch.tensor([[0.2, 0.4, 0.6]], device=DEV))
    assert img.shape == (2, 9, 12, 3) and torch.isfinite(img).all() and img.min() < 0.99
    # (b) nothing visible: everything behind the camera -> all-empty fragments
    gm2 = GaussianMeshesNaive(torch.tensor([[0.0, 0.0, 50.0]], device=DEV), torch.tensor([3.0], device=DEV))
    R1, T1 = oracle.look_at_view(3.0, 0.0, 0.0)
    cams1 = PerspectiveCameras(focal_length=20.0, principal_point=((6.0, 4.5),), R=R1, T=T1, in_ndc=False,
                               image_size=((9, 12),), device=DEV)
    r1 = GaussianRenderer(cams1, GaussianRenderSettings(image_size=(9, 12), max_assign=4)).to(DEV)
    f = r1(gm2)
    assert (f.vert_index == -1).all() and (f.vert_weight == 0).all() and (f.vert_hit_length == 1e10).all() and (f.valid_num == 0).all()
    assert torch.equal(to_white_background(f, torch.rand(1, 3, device=DEV)), torch.ones(1, 9, 12, 3, device=DEV))
    # (c) inverse_sigma=True (covariances given) goes through the fused path with autograd through torch.inverse
    sc = small_scene(seed=13, aniso=True, n=120)
    cov = torch.inverse(0.5 * (sc["sigmas"] + sc["sigmas"].transpose(1, 2)))
    cams3 = PerspectiveCameras(focal_length=sc["focal"], principal_point=(sc["principal"],), R=sc["R"], T=sc["T"],
                               in_ndc=False, image_size=(sc["image_size"],), device=DEV)
    st3 = GaussianRenderSettings(image_size=sc["image_size"], max_assign=6, inverse_sigma=True, max_point_per_bin=120)
    r3 = GaussianRenderer(cams3, st3).to(DEV)
    covp = cov.to(DEV).requires_grad_(True)
    gm3 = GaussianMeshesNaive(sc["verts"].to(DEV), covp)
    a3, b3 = _both(r3, gm3)
    # fused: adjugate inverse inside voge_pack_gaussians; op-by-op: torch.inverse (LU) -- last-bit differences of S
    # may flip a hit that sits within rounding of the threshold
    assert float((a3.vert_index == b3.vert_index).all(-1).float().mean()) > 0.99 and (a3.vert_index >= 0).sum() > 200
    interpolate_attr(a3, sc["colors"].to(DEV)).sum().backward()
    assert torch.isfinite(covp.grad).all() and covp.grad.abs().sum() > 0


@pytest.mark.parametrize("kind", ["iso", "full"])
def test_camera_gradients_fused_match_op_by_op(kind):
    """Pose optimisation: rays / origins that require grad go through the fused path too; its backward emits
    d/d(rays), d/d(origins) (reference grad_rays, RayTracing.py:179-206 / ray_trace_voge.cu:283-332), which
    autograd carries on to R, T and the focal length through the ray generator."""
    from voge_b200.cameras import PerspectiveCameras
    from voge_b200.Meshes import GaussianMeshesNaive
    from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings, get_silhouette, to_white_background
    sc = small_scene(seed=17, aniso=(kind == "full"), n=150, views=2)
    H, W = sc["image_size"]
    sig = sc["sigmas"] if kind == "full" else sc["sigmas"][:, 0, 0].contiguous()
    colors = sc["colors"].to(DEV)
    target = torch.rand(2, H, W, 3, generator=torch.Generator().manual_seed(3)).to(DEV)
    grads = {}
    for fused in (True, False):
        Rp = sc["R"].to(DEV).requires_grad_(True)
        Tp = sc["T"].to(DEV).requires_grad_(True)
        fp = torch.full((2, 2), float(sc["focal"]), device=DEV, requires_grad=True)
        cams = PerspectiveCameras(focal_length=fp, principal_point=(sc["principal"],), R=Rp, T=Tp, in_ndc=False,
                                  image_size=((H, W),), device=DEV)
        r = GaussianRenderer(cams, GaussianRenderSettings(image_size=(H, W), max_assign=6, max_point_per_bin=150)).to(DEV)
        r.use_fused = fused
        vp = sc["verts"].to(DEV).requires_grad_(True)
        f = r(GaussianMeshesNaive(vp, sig.to(DEV)))
        loss = ((to_white_background(f, colors) - target) ** 2).mean() + 1e-2 * get_silhouette(f).mean() \
            + 1e-3 * f.vert_hit_length.clamp(max=100).mean()
        loss.backward()
        grads[fused] = (Rp.grad.clone(), Tp.grad.clone(), fp.grad.clone(), vp.grad.clone())
    for name, a, b in zip(("R", "T", "focal", "verts"), grads[True], grads[False]):
        assert torch.isfinite(a).all() and a.abs().sum() > 0, name
        assert (a - b).abs().max() <= 2e-4 * b.abs().max() + 1e-10, (name, float((a - b).abs().max()), float(b.abs().max()))


def test_sampler_api_roundtrip(oracle):
    """sample_features / scatter_max_weight on fragments from the renderer (Sampler.py:5-42), fwd + bwd."""
    from voge_b200.Sampler import sample_features, scatter_max_weight
    sc = small_scene(seed=19, aniso=True, n=200)
    renderer, gm = _setup(sc, "full", M=200)
    frag = renderer(gm)
    H, W = sc["image_size"]
    img = torch.rand(1, H, W, 3, generator=torch.Generator().manual_seed(5)).to(DEV).requires_grad_(True)
    feat, wsum = sample_features(frag, img, n_vert=200)
    f_o, s_o = oracle.sample(img.detach().cpu(), frag.vert_weight.detach().cpu(), frag.vert_index.cpu(), 200)
    assert np.allclose(feat.detach().cpu().numpy(), f_o, rtol=1e-4, atol=1e-5) and np.allclose(wsum.detach().cpu().numpy(), s_o, rtol=1e-4, atol=1e-5)
    (feat.sum() + wsum.sum()).backward()
    assert torch.isfinite(img.grad).all() and gm.verts.grad is not None and torch.isfinite(gm.verts.grad).all()
    wmax = scatter_max_weight(frag, n_vert=200)
    assert np.array_equal(wmax.cpu().numpy(), oracle.scatter_max(frag.vert_weight.detach().cpu(), frag.vert_index.cpu(), 200))


@pytest.mark.parametrize("K,hw", [(8, (48, 48)), (20, (64, 64)), (30, (40, 56))])
def test_forward_pipeline_dense_ties_and_single_kernel(K, hw):
    """trace_hits -> select_topk -> blend_weights against the op-by-op chain on a
    scene built to leave the sorting-network fast path: > 64 hits per pixel (exact-key selection), duplicated
    Gaussians (identical len: ties broken by index), lens spanning many binades."""
    from voge_b200 import _C
    from voge_b200.cameras import camera_params
    from voge_b200.fused import choose_tile
    from voge_b200.RayTracing import default_bin_size
    sc = small_scene(seed=23, aniso=True, views=2, n=500, image_size=hw, focal=60.0)
    sig = sc["sigmas"]
    sig[:300] *= 0.01                       # wide blobs: every pixel collects a long hit list
    verts = sc["verts"]
    verts[300:360] = verts[0:60]            # exact duplicates -> identical (len, act), ordered by index
    sig[300:360] = sig[0:60]
    verts[360:380] *= 1e-3                  # Gaussians right at the look-at point ...
    verts[380:400] = verts[380:400] * 0.02 + sc["R"][0][:, 2] * 0.0   # ... and a tight cluster
    renderer, gm = _setup(sc, "full", K=K, M=500)
    a, b = _both(renderer, gm)
    assert torch.equal(a.vert_index, b.vert_index) and torch.equal(a.vert_hit_length, b.vert_hit_length)
    assert torch.equal(a.valid_num, b.valid_num)
    assert torch.allclose(a.vert_weight, b.vert_weight, rtol=1e-5, atol=1e-9, equal_nan=True)
    # the same fragments from the one-launch kernel, and the counters of the pipeline
    rays, origins = renderer._rays(hw)
    R, T, focal, principal = camera_params(renderer.cameras, hw)
    R, T = R.expand(2, -1, -1).contiguous(), T.expand(2, -1).contiguous()
    focal, principal = focal.expand(2, -1).contiguous(), principal.expand(2, -1).contiguous()
    thr_act = -math.log(0.01 + 1e-10)
    bs = default_bin_size(hw); tile = choose_tile(bs, K, True)
    off, tl, rects, ioff = _C.bin_views(gm.verts, gm.sigmas, R, T, origins, focal, principal, hw, 0.01, thr_act, True, bs, tile)
    stats = torch.zeros(4, dtype=torch.int64, device=DEV)
    p = _C.render_forward(gm.verts, gm.sigmas, origins, rays, off, tl, rects, thr_act, 1.0, K, tile, stats=stats, item_offsets=ioff)
    assert torch.equal(p[0], a.vert_index) and torch.equal(p[2], a.vert_hit_length) and torch.equal(p[3], a.valid_num)
    # rays generated inside the kernels (cam records) == the materialised rays of voge_generate_rays
    cam = _C.make_cam(R, focal, principal)
    assert torch.equal(_C.generate_rays(cam, hw), rays)
    q = _C.render_forward(gm.verts, gm.sigmas, origins, None, off, tl, rects, thr_act, 1.0, K, tile, item_offsets=ioff,
                          cam=cam, image_size=hw)
    for name, x, y in zip(("idx", "weight", "len", "valid", "act", "dsd"), p, q):
        assert torch.equal(x, y), name
    # views traced one group at a time (bounded scratch) give the same fragments
    g = _C.render_forward(gm.verts, gm.sigmas, origins, rays, off, tl, rects, thr_act, 1.0, K, tile, item_offsets=ioff,
                          max_group_items=1)
    for x, y in zip(p, g):
        assert torch.equal(x, y)
    assert int(stats[0]) == ioff.total_items - ioff.slack_items > 0          # every item evaluated exactly once
    assert int(stats[2]) > 0                              # the exact-key selection was exercised
    assert int(p[3].max()) == K


@pytest.mark.parametrize("K,hw,scale", [(20, (64, 64), 0.08), (40, (64, 64), 0.08), (60, (48, 80), 0.05), (25, (40, 56), 0.12),
                                        (64, (64, 64), 0.05)])
def test_select_topk_hit_count_classes(K, hw, scale):
    """select_topk over every per-pixel hit-count class -- two lanes per pixel with 8 / 16 / 32 slots each (<= 64 hits,
    ranks 32.. written by the second lane when K > 32), four lanes per pixel (65..128 hits, K <= 63), warp-cooperative
    and per-lane exact selection (more hits, K = 64, tied lens) -- for K below / at / above 32 and K % 4 != 0: bit-identical index lists
    to the op-by-op chain (pixel-major fine kernel with the reference's insertion rule)."""
    from voge_b200 import _C
    from voge_b200.cameras import camera_params
    from voge_b200.fused import choose_tile
    from voge_b200.RayTracing import default_bin_size
    sc = small_scene(seed=31, aniso=True, views=2, n=700, image_size=hw, focal=70.0)
    sig = sc["sigmas"]
    sig[:450] *= scale                      # wide blobs of graded density: hit lists from a few to > 64 entries
    verts = sc["verts"]
    verts[450:470] = verts[0:20]            # duplicates: tied lens inside otherwise ordinary pixels
    sig[450:470] = sig[0:20]
    renderer, gm = _setup(sc, "full", K=K, M=700)
    rays, origins = renderer._rays(hw)
    R, T, focal, principal = camera_params(renderer.cameras, hw)
    R, T = R.expand(2, -1, -1).contiguous(), T.expand(2, -1).contiguous()
    focal, principal = focal.expand(2, -1).contiguous(), principal.expand(2, -1).contiguous()
    thr_act = -math.log(0.01 + 1e-10)
    bs = default_bin_size(hw); tile = choose_tile(bs, K, True)
    off, tl, rects, ioff = _C.bin_views(gm.verts, gm.sigmas, R, T, origins, focal, principal, hw, 0.01, thr_act, True, bs, tile)
    stats = torch.zeros(4, dtype=torch.int64, device=DEV)
    dbg = {}
    p = _C.render_forward(gm.verts, gm.sigmas, origins, rays, off, tl, rects, thr_act, 1.0, K, tile, stats=stats,
                          item_offsets=ioff, debug=dbg)
    a, b = _both(renderer, gm)
    assert torch.equal(a.vert_index, b.vert_index) and torch.equal(a.valid_num, b.valid_num)
    assert torch.equal(a.vert_hit_length, b.vert_hit_length)
    assert torch.equal(p[0], b.vert_index) and torch.equal(p[3], b.valid_num)
    wmax = dbg["counts"].view(-1, 32).max(dim=1).values
    classes = [int(((wmax >= lo) & (wmax <= hi)).sum()) for lo, hi in ((1, 16), (17, 32), (33, 64), (65, 10 ** 9))]
    assert classes[2] > 0 and classes[3] > 0 and (classes[0] > 0 or classes[1] > 0), classes
    assert int(stats[2]) > 0
    # 65 .. 128 hits: four lanes per pixel (select_quad, K <= 63); more: the exact-key selection
    cpix = dbg["counts"]
    n_quad, n_more = int(((cpix > 64) & (cpix <= 128)).sum()), int((cpix > 128).sum())
    print("[select classes K=%d] pixels with 65..128 hits: %d, with more: %d, exact-key selections: %d" % (K, n_quad, n_more, int(stats[2])))
    assert n_quad > 0
    # selected lists are ascending in (len, idx) and padded consistently
    idx, ln, valid = p[0], p[2], p[3]
    kk = torch.arange(K, device=DEV).view(1, 1, 1, K)
    live = kk < valid.unsqueeze(-1)
    assert bool(((idx >= 0) == live).all())
    d = ln[..., 1:] - ln[..., :-1]
    assert bool((d[live[..., 1:]] >= 0).all())

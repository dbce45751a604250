"""GPU parity at the shapes of BASELINE.json's configs: the public renderer API against the
CPU oracle on the same rays.  C1 quick-start cuboid (866 Gaussians, 256^2, K=20, M=200), C3 shape fitting
(ico_sphere(4) = 2562 Gaussians, 128^2, K=25, no coarse stage, 5 views), C4 occlusion reasoning (two cuboids,
400^2, K=60, M=1500).  C2 / C5 sizes are covered through size-independent properties in
test_gpu_fused.py / bench.py (the oracle needs minutes there)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _render_and_check(oracle, verts, sig, R, T, focal, hw, K, M, frac=0.9995):
    from voge_b200.cameras import PerspectiveCameras
    from voge_b200.Meshes import GaussianMeshes
    from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings
    H, W = hw
    cams = PerspectiveCameras(focal_length=focal, principal_point=((W / 2, H / 2),), R=R, T=T, in_ndc=False,
                              image_size=((H, W),), device=DEV)
    st = GaussianRenderSettings(image_size=(H, W), max_assign=K, thr_activation=0.01, max_point_per_bin=M)
    renderer = GaussianRenderer(cams, st).to(DEV)
    gm = GaussianMeshes(verts.clone(), sig.clone()).to(DEV)
    frag = renderer(gm)
    rays, origins = renderer._rays(hw)
    o = oracle.render_reference_cpu(verts, sig, R, T, focal, (W / 2, H / 2), hw, K=K, max_points_per_bin=M, rays=rays,
                                    origin=origins)
    idx = frag.vert_index.cpu()
    same = idx == o["idx"]
    # candidate sets can differ from the oracle's only for a Gaussian whose bbox edge is within an ulp of a
    # bin edge (bbox maths is fp32 PyTorch in the oracle, closed form in the kernel)
    assert same.float().mean() > frac
    rows = same.all(dim=-1)
    assert torch.equal(frag.vert_hit_length.cpu()[rows], o["len"][rows])
    assert torch.equal(frag.valid_num.cpu()[rows], o["valid_num"][rows])
    assert torch.allclose(frag.vert_weight.cpu()[rows], o["weight"][rows], rtol=1e-5, atol=1e-7)
    assert (idx >= 0).sum() > 1000
    return renderer, gm, frag


def test_c1_quickstart_cuboid(oracle):
    from voge_b200 import scenes
    v, s = scenes.cuboid_gauss((-1, 1), (-1, 1), (-1, 1), 1000, percentage=0.6)
    verts, sig = torch.tensor(v, dtype=torch.float32), torch.tensor(s, dtype=torch.float32)
    assert verts.shape[0] == 866
    R, T = oracle.look_at_view(6.0, 10.0, 70.0)
    renderer, gm, frag = _render_and_check(oracle, verts, sig, R, T, 300.0, (256, 256), K=20, M=200)
    # forward + backward through to_white_background, finite gradients on every parameter
    from voge_b200.Renderer import to_white_background
    colors = torch.rand(866, 3, device=DEV, requires_grad=True)
    to_white_background(frag, colors).square().mean().backward()
    assert torch.isfinite(gm.verts.grad).all() and gm.verts.grad.abs().sum() > 0
    assert torch.isfinite(gm.sigmas.grad).all() and torch.isfinite(colors.grad).all()


def test_c3_shape_fitting_no_coarse(oracle):
    from voge_b200 import scenes
    verts = torch.tensor(scenes.ico_sphere(4)[0], dtype=torch.float32)
    assert verts.shape[0] == 2562
    sig = torch.full((2562,), 400.0)
    R, T = oracle.look_at_view(torch.full((5,), 2.7), torch.tensor([0.0, 20.0, -15.0, 40.0, 5.0]),
                               torch.tensor([0.0, 72.0, 144.0, 216.0, 288.0]))
    _render_and_check(oracle, verts, sig, R, T, 150.0, (128, 128), K=25, M=-1, frac=0.99999)


def test_c4_two_cuboids_k60(oracle):
    from voge_b200 import scenes
    v1, s1 = scenes.cuboid_gauss((-0.6, 0.6), (-0.4, 0.4), (-0.5, 0.5), 1500, percentage=0.6)
    v2, s2 = scenes.cuboid_gauss((-0.5, 0.5), (-0.5, 0.5), (-0.3, 0.3), 1200, percentage=0.6)
    v2 = v2 + np.array([0.4, 0.1, -0.9])
    verts = torch.tensor(np.concatenate([v1, v2]), dtype=torch.float32)
    sig = torch.tensor(np.concatenate([s1, s2]), dtype=torch.float32)
    R, T = oracle.look_at_view(4.0, 15.0, 30.0)
    _render_and_check(oracle, verts, sig, R, T, 300.0, (400, 400), K=60, M=1500)


def test_c5_full_size_band_and_properties(oracle):
    """BASELINE.json's metric configuration at full size (1M Gaussians, 1024^2, K=20), one view: a 64-row band
    of the fragments bit-exact against the CPU oracle (reference coarse bins + fine kernel restatement), and
    size-independent properties of the whole frame (sorted lens, unique in-range indices, padding, valid_num
    checksum, determinism, pipeline == one-launch kernel)."""
    import math
    from voge_b200 import _C, scenes
    from voge_b200.cameras import PerspectiveCameras, camera_params
    from voge_b200.fused import choose_tile
    from voge_b200.Meshes import GaussianMeshes
    from voge_b200.RayTracing import default_bin_size
    from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings
    N, HW, K = 1_000_000, 1024, 20
    verts, sig, _ = scenes.synthetic_scene(N, seed=0)
    R, T = oracle.look_at_view(3.0, 0.0, 0.0)
    cams = PerspectiveCameras(focal_length=900.0, principal_point=((HW / 2, HW / 2),), R=R, T=T, in_ndc=False,
                              image_size=((HW, HW),), device=DEV)
    renderer = GaussianRenderer(cams, GaussianRenderSettings(image_size=(HW, HW), max_assign=K)).to(DEV)
    gm = GaussianMeshes(verts.clone(), sig.clone()).to(DEV)
    with torch.no_grad():
        frag = renderer(gm)
        frag2 = renderer(gm)
    idx, ln, w, valid = frag.vert_index, frag.vert_hit_length, frag.vert_weight, frag.valid_num
    # determinism (hit order inside a segment is arbitrary, the selection is not)
    assert torch.equal(idx, frag2.vert_index) and torch.equal(ln, frag2.vert_hit_length) and torch.equal(w, frag2.vert_weight)
    # properties
    ok = idx >= 0
    assert torch.equal(ok.sum(-1), valid) and int(valid.max()) == K and int(valid.sum()) > 5_000_000
    assert bool(((idx < N) & (idx >= -1)).all())
    assert bool((ln[~ok] == 1e10).all()) and bool((w[~ok] == 0).all())
    k = torch.arange(K, device=DEV)
    assert bool((ok == (k < valid[..., None])).all())                       # valid slots are the leading ones
    d = ln[..., 1:] - ln[..., :-1]
    assert bool((d[ok[..., 1:]] >= 0).all())                                # ascending hit lengths
    srt = torch.sort(torch.where(ok, idx, -1 - k.expand_as(idx)), dim=-1).values
    assert bool((srt[..., 1:] != srt[..., :-1]).all())                      # a Gaussian hits a pixel at most once
    assert bool((w >= 0).all()) and bool((w <= math.exp(0.5) + 1e-6).all())
    # the one-launch kernel gives the same fragments
    rays, origins = renderer._rays((HW, HW))
    Rm, Tm, focal, principal = camera_params(cams, (HW, HW))
    thr_act = -math.log(0.01 + 1e-10)
    bs = default_bin_size((HW, HW)); tile = choose_tile(bs, K, True)
    off, tl, rects, _ = _C.bin_views(gm.verts, gm.sigmas, Rm, Tm, origins, focal, principal, (HW, HW), 0.01, thr_act, True, bs, tile)
    s = _C.render_forward(gm.verts, gm.sigmas, origins, rays, off, tl, rects, thr_act, 1.0, K, tile, need_act=False)
    assert torch.equal(s[0], idx) and torch.equal(s[2], ln) and torch.equal(s[3], valid)
    # a band of rows against the CPU oracle: reference coarse bins (bin 32) + fine kernel restatement
    rows, y0 = 64, 480
    mus = (verts[None] - origins.cpu()[:, None])
    isg = (2 * sig)[None]
    ndc, radii = oracle.coarse_inputs(R, T, 900.0, (HW / 2.0, HW / 2.0), (HW, HW), mus, isg, 0.01)
    first, nper = torch.zeros(1, dtype=torch.long), torch.full((1,), N)
    bp, bc = oracle.rasterize_coarse(ndc.reshape(-1, 3), radii.reshape(-1, 2), first, nper, (HW, HW), bs, 8192)
    assert bc.max() <= 8192
    bp_sub = torch.from_numpy(bp[:, y0 // bs:(y0 + rows) // bs, :, :int(bc.max())].copy())
    rays_sub = rays[:, y0:y0 + rows].cpu().contiguous()
    o_idx, o_len, _, _ = (torch.from_numpy(a) for a in oracle.ray_trace_fine(mus.reshape(-1, 3), isg.reshape(-1, 3, 3), rays_sub,
                                                                             bp_sub, thr_act, bs, K))
    g_idx, g_len = idx[:, y0:y0 + rows].cpu(), ln[:, y0:y0 + rows].cpu()
    same = (g_idx == o_idx).all(-1)
    # candidate sets can differ only for a Gaussian whose bbox edge is within an ulp of a bin edge
    assert same.float().mean() > 0.9995
    assert torch.equal(g_len[same], o_len[same])

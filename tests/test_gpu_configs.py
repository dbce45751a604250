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

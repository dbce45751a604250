"""CPU tests of the host layer: the C-ABI library loads and exports every symbol the header
declares, the Python surface mirrors the reference's names, misuse raises like the reference."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from voge_b200 import _lib
    header = open(os.path.join(ROOT, "include", "voge_b200.h")).read()
    declared = set(re.findall(r"\b(voge_[a-z0-9_]+)\s*\(", header))
    declared.discard("voge_stream_t")
    assert len(declared) >= 15
    h = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(h, name), "libvoge_b200.so does not export " + name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.lib().voge_version() == 1
    assert _lib.lib().voge_rasterize_coarse_scratch_elems(2, 5000, 64, 64, 16) == 2 * 3 * 16


def test_python_surface_matches_reference_names():
    import voge_b200 as V
    for mod, names in {
        "Renderer": ["GaussianRenderer", "GaussianRenderSettings", "Fragments", "interpolate_attr", "get_silhouette",
                     "to_colored_background", "to_white_background"],
        "RayTracing": ["ray_tracing", "rasterize_coarse", "ray_tracing_fine", "convert_to_box", "_RayTraceVoGE",
                       "_RasterizeCoarse", "ray_trace_voge_ray", "find_nearest_k", "find_farest_k", "_RayTraceVoGERay",
                       "_FindNearestK"],
        "Aggregation": ["aggregation", "merge_final", "expend_sigma", "get_cross_activation", "assign2weight",
                        "inverse_cumsum", "get_ray_camera_space"],
        "Sampler": ["sample_features", "scatter_max_weight", "_SampleVoGE", "_ScatterMax"],
        "Meshes": ["GaussianMeshes", "GaussianMeshesNaive", "DeformedGaussianMeshes"],
        "Utils": ["ind_sel", "ind_fill", "rotation_theta", "eye_like"],
        "_C": ["rasterize_points_coarse", "ray_trace_voge_fine", "ray_trace_voge_fine_backward", "sample_voge",
               "sample_voge_backward", "scatter_max", "ray_trace_voge_ray", "ray_trace_voge_ray_backward",
               "find_nearest_k"],
    }.items():
        for n in names:
            assert hasattr(getattr(V, mod), n), "%s.%s missing" % (mod, n)


def test_no_cpu_fallback():
    from voge_b200 import _C
    z = torch.zeros(1, 2, 2, 4)
    zi = torch.zeros(1, 2, 2, 4, dtype=torch.int32)
    with pytest.raises(RuntimeError):
        _C.sample_voge(torch.zeros(1, 2, 2, 3), z, zi, 5)
    with pytest.raises(RuntimeError):
        _C.ray_trace_voge_fine(torch.zeros(3, 3), torch.zeros(3, 3, 3), torch.zeros(1, 4, 4, 3),
                               torch.zeros(1, 1, 1, 3, dtype=torch.int32), 4.6, 10, 2)
    with pytest.raises(RuntimeError):
        _C.aggregation_forward(zi, z, z, z, 1.0)


def test_settings_and_fragments_semantics():
    from voge_b200.Renderer import Fragments, GaussianRenderSettings
    s = GaussianRenderSettings(image_size=128, max_assign=7, batch_size=-1, principal_point=(1, 2))   # extra kwargs ignored
    assert s["image_size"] == (128, 128) and s["max_assign"] == 7 and s["thr_activation"] == 0.01
    assert s["absorptivity"] == 1 and s["inverse_sigma"] is False and s["max_point_per_bin"] is None
    f = Fragments(torch.zeros(2, 3, 4, 5), torch.zeros(2, 3, 4, 5, dtype=torch.int32), torch.zeros(2, 3, 4, dtype=torch.long),
                  torch.zeros(2, 3, 4, 5))
    assert len(f) == 2 and f[0].valid_num.shape == (3, 4) and f[0].unsqueeze().valid_num.shape == (1, 3, 4)
    assert f[0:1].squeeze().vert_weight.shape == (3, 4, 5) and set(f.to_dict()) == {"vert_weight", "vert_index", "valid_num", "vert_hit_length"}
    assert f.copy().vert_weight.data_ptr() == f.vert_weight.data_ptr()
    with pytest.raises(AssertionError):
        f[0][0]


def test_helpers_match_reference_semantics(oracle):
    from voge_b200.Aggregation import assign2weight, expend_sigma, get_cross_activation
    from voge_b200.RayTracing import default_bin_size, default_max_points_per_bin
    from voge_b200.Utils import ind_fill, ind_sel, rotation_theta
    assert [default_bin_size((s, s)) for s in (128, 256, 400, 512, 672, 1024)] == [10, 10, 16, 16, 32, 32]
    assert default_max_points_per_bin(20, 866) == 200 and default_max_points_per_bin(20, 10 ** 6) == 100000
    assert default_max_points_per_bin(25, 100) == 100
    g = torch.Generator().manual_seed(0)
    ln, ds, ac = torch.rand(7, 5, generator=g) * 3, torch.rand(7, 5, generator=g) * 50, torch.rand(7, 5, generator=g)
    w = assign2weight(ac, get_cross_activation(ln, ds), 1.3)
    w0 = oracle.aggregation_torch(torch.zeros(7, 5, dtype=torch.int32), ac, ln, ds, 1.3)[0]
    assert torch.equal(w, w0)
    assert torch.equal(expend_sigma(torch.tensor([2.0, 3.0]))[1], 3.0 * torch.eye(3))
    assert torch.equal(expend_sigma(torch.tensor([[1.0, 2.0, 3.0]]))[0], torch.diag(torch.tensor([1.0, 2.0, 3.0])))
    t = torch.arange(24.).view(2, 4, 3)
    i = torch.tensor([[0, 3], [2, 2]])
    assert torch.equal(ind_sel(t, i, dim=1)[1, 0], t[1, 2])
    assert ind_fill(torch.zeros(2, 4), i, 1.0, dim=1).sum() == 3
    r = rotation_theta(torch.tensor([0.3]))
    assert torch.allclose(r[0] @ r[0].T, torch.eye(3), atol=1e-6)


def test_camera_shim_consistent_with_closed_form(oracle):
    from voge_b200.cameras import PerspectiveCameras, generate_rays, look_at_view_transform
    from voge_b200.RayTracing import coarse_inputs
    R, T = look_at_view_transform(dist=6, elev=10, azim=70)
    R0, T0 = oracle.look_at_view(6, 10, 70)
    assert torch.allclose(R, R0, atol=1e-6) and torch.allclose(T, T0, atol=1e-6)
    H, W = 48, 64
    cam = PerspectiveCameras(focal_length=80.0, principal_point=((30.0, 25.0),), R=R, T=T, in_ndc=False, image_size=((H, W),))
    d, o = generate_rays(cam, (H, W))
    d0, o0 = oracle.camera_rays(R0, T0, 80.0, (30.0, 25.0), (H, W))
    assert torch.allclose(d, d0, atol=1e-6) and torch.allclose(o, o0, atol=1e-5)
    assert torch.allclose(d.norm(dim=-1), torch.ones(1, H, W), atol=1e-6)
    g = torch.Generator().manual_seed(0)
    pts = torch.rand(1, 50, 3, generator=g) - 0.5 - o[:, None]
    isg = (torch.rand(50, generator=g) * 50 + 20).view(1, 50, 1, 1) * torch.eye(3)
    ndc, box = coarse_inputs(cam, pts, isg, 0.01)            # through 4x4 transform objects (reference route)
    ndc0, box0 = oracle.coarse_inputs(R0, T0, 80.0, (30.0, 25.0), (H, W), pts, isg, 0.01)   # closed form
    assert torch.allclose(ndc, ndc0, atol=2e-5) and torch.allclose(box, box0, rtol=1e-4)
    # a point on the optical axis projects to the principal point: flipped NDC = (px - W/2)/s, (py - H/2)/s
    axis = (o + 3.0 * R[:, :, 2]) - o
    n2, _ = coarse_inputs(cam, axis[:, None], isg[:, :1], 0.01)
    assert torch.allclose(n2[0, 0, :2], torch.tensor([(30.0 - W / 2) / 24.0, (25.0 - H / 2) / 24.0]), atol=1e-5)
    assert abs(n2[0, 0, 2].item() - 3.0) < 1e-5


def test_batchifier_helpers():
    """VoGE.Utils.Batchifier / Reshaper (reference Utils.py:59-176): chunked execution over flattened dims gives
    the unchunked result; scalars are summed; the converters' call pattern (target_dims=0, tbar=True) works."""
    from VoGE.Utils import Batchifier, DataParallelBatchifier, Reshaper  # noqa: F401
    x = torch.randn(2, 5, 7, 3, generator=torch.Generator().manual_seed(0))

    def fn(a, b):
        return a * 2, (a * b).sum()
    for kw in (dict(target_dims=(1, 2)), dict(remain_dims=(0, -1)), dict(target_dims=0), dict(target_dims=-1)):
        y, s = Batchifier(4, batch_args=("a",), **kw)(fn)(a=x, b=3.0)
        assert y.shape == x.shape and torch.equal(y, x * 2) and torch.allclose(s, (x * 3).sum())
    pts = torch.randn(37, 3, generator=torch.Generator().manual_seed(1))

    def nn(point_v, point_t):
        return (point_v - point_t).pow(2).sum(-1).pow(.5).topk(4, dim=1, largest=False)[0].mean(1)
    want = nn(point_v=pts.unsqueeze(1), point_t=pts.unsqueeze(0))
    got = Batchifier(8, batch_args="point_v", target_dims=0, tbar=True)(nn)(point_v=pts.unsqueeze(1), point_t=pts.unsqueeze(0))
    assert torch.allclose(want, got)
    assert Reshaper((2, 3), 0)([torch.ones(4, 5), torch.ones(2, 5)]).shape == (2, 3, 5)


def test_bin_plan_capacities_and_groups():
    """_C.BinPlan (speculative binning): capacities = previous totals + 12.5 % (+ a floor), the views of a call are
    split into equal groups whose capacity fits the scratch budget, every view lands in exactly one group."""
    from voge_b200 import _C
    p = _C.BinPlan(1_000_000, 57_600_000)
    assert p.list_cap == 1_000_000 * 9 // 8 + 1024 and p.view_items == 57_600_000 * 9 // 8 + 4096
    g = p.groups(64, 1 << 29)
    assert g[0] == (0, 8) and g[-1] == (56, 64) and len(g) == 8
    assert p.hits_cap == 8 * p.view_items <= (1 << 29)
    assert [b for a, b in g[:-1]] == [a for a, b in g[1:]]                  # contiguous, no gaps
    # a single view larger than the budget still forms a group (as in the exact pass)
    g1 = p.groups(3, 1000)
    assert g1 == [(0, 1), (1, 2), (2, 3)] and p.hits_cap == p.view_items
    # few views: one group, capacity = views x largest view
    p2 = _C.BinPlan(10, 100)
    assert p2.groups(5, 1 << 29) == [(0, 5)] and p2.hits_cap == 5 * p2.view_items
    p2.update(2000, 3000)
    assert p2.list_cap == 2000 * 9 // 8 + 1024 and p2.view_items == 3000 * 9 // 8 + 4096

"""Converter / IO parity against golden values produced by the reference's own functions
(tools/make_golden_converters.py -> tests/golden/converters.npz) and properties of the normal-aligned converter."""
import os

import numpy as np
import torch


def _g(golden_dir):
    return np.load(os.path.join(golden_dir, "converters.npz"))


def test_vertex_converters_match_reference(golden_dir):
    from voge_b200.Converter import Converters as C
    g = _g(golden_dir)
    verts, faces = g["verts"], g["faces"]
    # the reference sums float32 distances vertex by vertex; here the sums are float64
    assert np.allclose(C.get_vert_edge_length(verts, faces, 1e-3), g["edge_len"], rtol=1e-6)
    v, s, r = C.naive_vertices_converter(verts, faces, percentage=0.5)
    assert r is None and np.array_equal(v, verts) and np.allclose(s, g["naive_isigma"], rtol=1e-5)
    assert np.allclose(C.naive_vertices_converter(verts, faces, percentage=0.3, max_sig_rate=1.2)[1], g["naive_isigma_capped"], rtol=1e-5)
    vt, st, _ = C.naive_vertices_converter(torch.from_numpy(verts), torch.from_numpy(faces))
    assert vt.dtype == torch.float32 and st.dtype == torch.float32 and np.allclose(st.numpy(), g["naive_isigma"], rtol=1e-5)
    # a vertex without faces falls back to the default length
    v2 = np.concatenate([verts, np.array([[5.0, 5.0, 5.0]], dtype=np.float32)])
    assert C.get_vert_edge_length(v2, faces, 0.25)[-1] == 0.25


def test_point_cloud_converters_match_reference(golden_dir):
    from voge_b200.Converter import Converters as C
    g = _g(golden_dir)
    p, s, _ = C.naive_point_cloud_converter(g["pts"], percentage=0.5, n_nearest=4, thr_max=2, chunk=64)
    assert np.array_equal(p, g["pts"]) and np.allclose(s, g["pc_isigma"], rtol=2e-5)
    pt, st, _ = C.naive_point_cloud_converter(torch.from_numpy(g["pts"]))
    assert torch.is_tensor(st) and np.allclose(st.numpy(), g["pc_isigma"], rtol=2e-5)
    assert np.allclose(C.fixed_pointcloud_converter(g["pts"], 0.05)[1], g["fixed_isigma"], rtol=1e-6)


def test_normal_mesh_converter_properties(golden_dir):
    from voge_b200.Converter import Converters as C
    g = _g(golden_dir)
    verts, faces = g["verts"], g["faces"]
    normals = verts / np.linalg.norm(verts, axis=1, keepdims=True)
    normals[0] = [0.0, 1.0, 0.0]          # parallel to `up`: the degenerate branch of look_at_rotation
    _, isig, _ = C.normal_mesh_converter(verts, faces, normals.astype(np.float32), percentage=0.5, shape_ratio=0.5)
    assert isig.shape == (162, 3, 3) and np.allclose(isig, isig.transpose(0, 2, 1), atol=1e-6 * np.abs(isig).max())
    base = g["naive_isigma"]
    w = np.linalg.eigvalsh(isig)
    # normals parallel to `up`: up x normal = 0 leaves a rank-1 matrix (as in pytorch3d's look_at_rotation), which
    # auto_fix replaces by the isotropic Gaussian (reference :60-62)
    deg = np.abs(normals[:, 1]) > 0.99999
    assert deg[0] and deg.sum() <= 4
    assert np.allclose(w[deg], base[deg, None], rtol=1e-4)
    ok = ~deg
    assert np.allclose(w[ok, 0], 0.5 * base[ok], rtol=1e-4) and np.allclose(w[ok, 1:], base[ok, None], rtol=1e-4)
    # the short axis of the ellipsoid's inverse covariance is the normal
    along = np.einsum("ni,nij,nj->n", normals, isig, normals)
    assert np.allclose(along[ok], 0.5 * base[ok], rtol=1e-4)


def test_off_goff_io_roundtrip_matches_reference(golden_dir, tmp_path):
    from voge_b200.Converter import IO
    g = _g(golden_dir)
    ref_off, ref_goff = tmp_path / "ref.off", tmp_path / "ref.goff"
    ref_off.write_bytes(g["off_text"].tobytes()); ref_goff.write_bytes(g["goff_text"].tobytes())
    v, f = IO.load_off(str(ref_off))
    assert np.array_equal(v, g["off_verts"]) and np.array_equal(f, g["off_faces"])
    p, s, r = IO.load_goff(str(ref_goff))
    assert r is None and np.array_equal(p, g["goff_points"]) and np.array_equal(s, g["goff_sigma"])
    # files written here load back identically (through the reference-shaped loader)
    mine = tmp_path / "mine.off"
    IO.save_off(str(mine), g["verts"], g["faces"])
    v2, f2 = IO.load_off(str(mine), to_torch=True)
    assert np.array_equal(v2.numpy(), g["off_verts"]) and np.array_equal(f2.numpy(), g["off_faces"])
    mg = tmp_path / "mine.goff"
    IO.save_goff(str(mg), torch.from_numpy(g["goff_points"]), torch.from_numpy(g["goff_sigma"]), radians=np.arange(162, dtype=np.float32))
    p2, s2, r2 = IO.load_goff(str(mg))
    assert np.array_equal(p2, g["goff_points"]) and np.array_equal(s2, g["goff_sigma"]) and np.array_equal(r2, np.arange(162, dtype=np.float32))
    assert IO.pre_process_pascal(np.array([[1.0, 2.0, 3.0]]))[0].tolist() == [[1.0, 3.0, -2.0]]


def test_cuboid_and_alias_package(golden_dir):
    """Cuboid.py:8-159 against the reference's own outputs: vertex order, sigma, faces, per-face colours, return types."""
    from VoGE.Converter.Cuboid import cuboid_gauss, cuboid_mesh
    from VoGE.Converter import Converters, IO  # noqa: F401
    from voge_b200.Meshes import GaussianMeshes
    g = _g(golden_dir)
    face_cols = np.eye(6, dtype=np.float32)
    for tag, a in (("a", ((-1, 1), (-1, 1), (-1, 1), 1000)), ("b", ((0, 2), (-1, 0.5), (3, 3.7), 400))):
        v, s, c = cuboid_gauss(*a, percentage=0.6, colors=face_cols)
        assert isinstance(v, np.ndarray) and v.dtype == np.float64 and isinstance(s, np.ndarray)     # ndarrays, as the reference
        assert np.array_equal(v, g["cub_%s_verts" % tag]) and np.array_equal(s, g["cub_%s_isigma" % tag])
        assert np.array_equal(c, g["cub_%s_colors" % tag])
        mv, mf, mc = cuboid_mesh(*a, colors=face_cols)
        assert np.array_equal(mv, g["mesh_%s_verts" % tag]) and np.array_equal(mf, g["mesh_%s_faces" % tag])
        assert np.array_equal(mc, g["mesh_%s_colors" % tag]) and mf.dtype == np.int64
    v, s = cuboid_gauss((-1, 1), (-1, 1), (-1, 1), 1000, percentage=0.6)
    assert v.shape == (866, 3) and s.shape == (866,)
    obj = cuboid_gauss((-1, 1), (-1, 1), (-1, 1), 1000, as_obj=True)
    assert isinstance(obj, GaussianMeshes) and obj.verts.dtype == torch.float32 and obj.verts.shape == (866, 3)
    mesh, cols = cuboid_mesh((-1, 1), (-1, 1), (-1, 1), 1000, colors=face_cols, as_obj=True)
    assert mesh.verts_list()[0].shape == (1014, 3) and mesh.faces_list()[0].dtype == torch.long and cols.shape == (1014, 6)

"""tests/golden/converters.npz: outputs of the reference's own converter / IO functions (imported from
/root/reference with pytorch3d stubbed out -- the functions exercised here never call it) on small inputs.
Run in the build container: python tools/make_golden_converters.py"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import _ref_import  # noqa: E402

REF = _ref_import.REF
_ref_import.load()
for name in ("pytorch3d", "pytorch3d.renderer", "pytorch3d.structures"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["pytorch3d.renderer"].look_at_rotation = None
sys.modules["pytorch3d.structures"].Meshes = type("Meshes", (), {})


def _load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "VoGE", "Converter", rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


conv = _load("Converters.py", "VoGE.Converter.Converters")
io = _load("IO.py", "VoGE.Converter.IO")
from voge_b200 import scenes  # noqa: E402

verts, faces = scenes.ico_sphere(2)
verts = (verts * np.array([1.0, 0.7, 1.3])).astype(np.float32)
faces = faces.astype(np.int64)
g = torch.Generator().manual_seed(5)
pts = torch.randn(300, 3, generator=g).numpy().astype(np.float32)
out = dict(verts=verts, faces=faces, pts=pts)
out["edge_len"] = conv.get_vert_edge_length(verts, faces, 1e-3)
out["naive_isigma"] = conv.naive_vertices_converter(verts, faces, percentage=0.5)[1]
out["naive_isigma_capped"] = conv.naive_vertices_converter(verts, faces, percentage=0.3, max_sig_rate=1.2)[1]
out["pc_isigma"] = conv.naive_point_cloud_converter(pts, percentage=0.5, n_nearest=4, thr_max=2)[1]
out["fixed_isigma"] = conv.fixed_pointcloud_converter(pts, 0.05, percentage=0.5)[1]
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "m.off")
    io.save_off(p, verts, faces)
    out["off_text"] = np.frombuffer(open(p, "rb").read(), dtype=np.uint8)
    v2, f2 = io.load_off(p)
    out["off_verts"], out["off_faces"] = v2, f2
    sig9 = np.tile(np.eye(3, dtype=np.float32)[None], (verts.shape[0], 1, 1)) * out["naive_isigma"].astype(np.float32)[:, None, None]
    q = os.path.join(d, "g.goff")
    io.save_goff(q, verts, sig9)
    out["goff_text"] = np.frombuffer(open(q, "rb").read(), dtype=np.uint8)
    gp, gs, gr = io.load_goff(q)
    out["goff_points"], out["goff_sigma"] = gp, gs
# cuboids (Cuboid.py:8-159): vertices, sigmas, faces, per-face colour expansion
sys.modules["VoGE.Meshes"] = __import__("voge_b200.Meshes", fromlist=["x"])
cub = _load("Cuboid.py", "VoGE.Converter.Cuboid")
face_cols = np.eye(6, dtype=np.float32)
for tag, a in (("a", ((-1, 1), (-1, 1), (-1, 1), 1000)), ("b", ((0, 2), (-1, 0.5), (3, 3.7), 400))):
    cv, cs, cc = cub.cuboid_gauss(*a, percentage=0.6, colors=face_cols)
    mv, mf, mc = cub.cuboid_mesh(*a, colors=face_cols)
    out.update({"cub_%s_verts" % tag: cv, "cub_%s_isigma" % tag: cs, "cub_%s_colors" % tag: cc,
                "mesh_%s_verts" % tag: mv, "mesh_%s_faces" % tag: mf, "mesh_%s_colors" % tag: mc})
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "converters.npz"), **out)
print({k: getattr(v, "shape", None) for k, v in out.items()})

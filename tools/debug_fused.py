import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from scene_utils import small_scene
from test_gpu_fused import _setup, _both
for kind, seed, views, M in [("full", 4, 2, -1), ("iso", 5, 1, -1), ("full", 4, 1, -1)]:
    sc = small_scene(seed=seed, aniso=True, views=views, n=400)
    renderer, gm = _setup(sc, kind, M=M)
    a, b = _both(renderer, gm)
    diff = (a.vert_index != b.vert_index).any(-1)
    print(kind, seed, views, M, "pixels differing:", int(diff.sum()), "of", diff.numel())
    nz = diff.nonzero()[:6]
    for bb, y, x in nz.tolist():
        print("  pix", bb, y, x, "fused", a.vert_index[bb, y, x].tolist(), "unfused", b.vert_index[bb, y, x].tolist())
        print("     len f", [round(v, 4) for v in a.vert_hit_length[bb, y, x].tolist()[:4]], "u", [round(v, 4) for v in b.vert_hit_length[bb, y, x].tolist()[:4]])

"""Run on the GPU box (gpurun): executes the UNMODIFIED reference CUDA kernels (oracle/_ref/voge_ref_C.so,
built by oracle/build_ref.py from /root/reference/VoGE/csrc) on seeded small scenes and stores
inputs + outputs as golden vectors:

    python tools/make_golden_gpu.py            -> gpurun_out/ref_gpu_golden.npz  (copy to tests/golden/)

It also prints, for each op, how the oracle (C restatement) and libvoge_b200 compare with the
reference on the spot, so a disagreement is visible in the gpurun log.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import build_ref  # noqa: E402
import voge_oracle as vo  # noqa: E402
from scene_utils import small_scene  # noqa: E402

DEV = "cuda:0"


def main():
    ref = build_ref.load_ref()
    from voge_b200 import _C as mine
    out = {}
    report = []

    def cmp(name, a, b, exact=True, tol=0.0):
        a, b = torch.as_tensor(a).cpu(), torch.as_tensor(b).cpu()
        if exact:
            ok = torch.equal(a, b)
            report.append("%-44s %s" % (name, "BIT-EXACT" if ok else "MISMATCH (%d of %d differ, max |d| %.3g)" % (
                int((a != b).sum()), a.numel(), float((a.double() - b.double()).abs().max()))))
        else:
            scale = float(b.abs().max()) + 1e-30
            err = float((a.double() - b.double()).abs().max()) / scale
            report.append("%-44s max err / max|ref| = %.3g %s" % (name, err, "ok" if err <= tol else "EXCEEDS %g" % tol))

    scenes = {"iso": dict(seed=0, aniso=False, views=1), "aniso": dict(seed=1, aniso=True, views=1),
              "multi": dict(seed=2, aniso=True, views=2, image_size=(40, 40))}
    for tag, kw in scenes.items():
        sc = small_scene(**kw)
        n = sc["verts"].shape[0]
        o = vo.render_reference_cpu(sc["verts"], sc["sigmas"], sc["R"], sc["T"], sc["focal"], sc["principal"],
                                    sc["image_size"], K=sc["K"], max_points_per_bin=n)
        B = sc["R"].shape[0]
        H, W = sc["image_size"]
        mus, isg, rays = o["mus"].to(DEV), o["isigmas"].to(DEV), o["rays"].to(DEV)
        # ---- coarse: reference kernel #1/#2 (runs at these bin counts: smem <= 48 KB) ----
        ndc, radii = vo.coarse_inputs(sc["R"], sc["T"], sc["focal"], sc["principal"], sc["image_size"],
                                      o["mus"].view(B, n, 3), o["isigmas"].view(B, n, 3, 3), 0.01)
        first = (torch.arange(B) * n).to(DEV)
        nper = torch.full((B,), n, dtype=torch.long, device=DEV)
        bp_ref = ref.rasterize_points_coarse(ndc.reshape(-1, 3).to(DEV), first, nper, (H, W),
                                             radii.reshape(-1, 2).to(DEV), o["bin_size"], n)
        torch.cuda.synchronize()
        bp_ref_sorted = torch.sort(torch.where(bp_ref < 0, torch.full_like(bp_ref, 2 ** 30), bp_ref), dim=-1).values
        bp_ref_sorted = torch.where(bp_ref_sorted == 2 ** 30, torch.full_like(bp_ref, -1), bp_ref_sorted)
        cmp(tag + ": coarse oracle vs REF (sorted)", o["bin_points"], bp_ref_sorted)
        bp_mine = mine.rasterize_points_coarse(ndc.reshape(-1, 3).to(DEV), first, nper, (H, W),
                                               radii.reshape(-1, 2).to(DEV), o["bin_size"], n)
        cmp(tag + ": coarse voge_b200 vs REF (sorted)", bp_mine, bp_ref_sorted)
        # ---- fine forward ----
        bp = o["bin_points"].to(DEV)
        r_idx, r_len, r_act, r_dsd = ref.ray_trace_voge_fine(mus, isg, rays, bp, o["thr_act"], o["bin_size"], sc["K"])
        torch.cuda.synchronize()
        m_idx, m_len, m_act, m_dsd = mine.ray_trace_voge_fine(mus, isg, rays, bp, o["thr_act"], o["bin_size"], sc["K"])
        for nm, a, b, c in (("idx", o["idx"], m_idx, r_idx), ("len", o["len"], m_len, r_len),
                            ("act", o["act"], m_act, r_act), ("dsd", o["dsd"], m_dsd, r_dsd)):
            cmp(tag + ": fine %s oracle vs REF" % nm, a, c)
            cmp(tag + ": fine %s voge_b200 vs REF" % nm, b, c)
        # ---- fine backward ----
        g = torch.Generator().manual_seed(100)
        gl, ga, gd = (torch.randn(r_idx.shape, generator=g) for _ in range(3))
        rg_ray, rg_mus, rg_isg = ref.ray_trace_voge_fine_backward(mus, isg, rays, r_idx, gl.to(DEV), ga.to(DEV), gd.to(DEV))
        torch.cuda.synchronize()
        og = vo.ray_trace_fine_backward(o["mus"], o["isigmas"], o["rays"], r_idx.cpu(), gl, ga, gd)
        mg = mine.ray_trace_voge_fine_backward(mus, isg, rays, r_idx, gl.to(DEV), ga.to(DEV), gd.to(DEV))
        for nm, a, b, c in (("grad_rays", og[0], mg[0], rg_ray), ("grad_mus", og[1], mg[1], rg_mus),
                            ("grad_isg", og[2], mg[2], rg_isg)):
            cmp(tag + ": bwd %s oracle vs REF" % nm, a, c, exact=False, tol=2e-5)
            cmp(tag + ": bwd %s voge_b200 vs REF" % nm, b, c, exact=False, tol=2e-5)
        # ---- sampling ----
        w = torch.rand(r_idx.shape, generator=g)
        img = torch.rand(B, H, W, 3, generator=g)
        n_vert = B * n
        r_feat, r_wsum = ref.sample_voge(img.to(DEV), w.to(DEV), r_idx, n_vert)
        gf, gs = torch.rand(n_vert, 3, generator=g), torch.rand(n_vert, generator=g)
        r_gi, r_gw = ref.sample_voge_backward(img.to(DEV), w.to(DEV), r_idx, gf.to(DEV), gs.to(DEV))
        r_max = ref.scatter_max(w.to(DEV), r_idx, n_vert)
        torch.cuda.synchronize()
        m_feat, m_wsum = mine.sample_voge(img.to(DEV), w.to(DEV), r_idx, n_vert)
        m_gi, m_gw = mine.sample_voge_backward(img.to(DEV), w.to(DEV), r_idx, gf.to(DEV), gs.to(DEV))
        m_max = mine.scatter_max(w.to(DEV), r_idx, n_vert)
        cmp(tag + ": sample feat voge_b200 vs REF", m_feat, r_feat, exact=False, tol=1e-5)
        cmp(tag + ": sample wsum voge_b200 vs REF", m_wsum, r_wsum, exact=False, tol=1e-5)
        cmp(tag + ": sample g_image voge_b200 vs REF", m_gi, r_gi, exact=False, tol=1e-5)
        cmp(tag + ": sample g_weight voge_b200 vs REF", m_gw, r_gw, exact=False, tol=1e-5)
        cmp(tag + ": scatter_max voge_b200 vs REF", m_max, r_max)
        store = dict(mus=o["mus"], isigmas=o["isigmas"], rays=o["rays"], bin_points=o["bin_points"],
                     ndc=ndc.reshape(-1, 3), radii=radii.reshape(-1, 2), coarse_ref_sorted=bp_ref_sorted,
                     idx=r_idx, len=r_len, act=r_act, dsd=r_dsd, gl=gl, ga=ga, gd=gd,
                     grad_rays=rg_ray, grad_mus=rg_mus, grad_isg=rg_isg, w=w, img=img, feat=r_feat, wsum=r_wsum,
                     gf=gf, gs=gs, g_image=r_gi, g_weight=r_gw, wmax=r_max)
        for k, v in store.items():
            out["%s_%s" % (tag, k)] = torch.as_tensor(v).cpu().numpy()
        out[tag + "_meta"] = np.array([B, H, W, sc["K"], o["bin_size"], n], dtype=np.int64)
        out[tag + "_thr_act"] = np.float64(o["thr_act"])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "ref_gpu_golden.npz")
    np.savez_compressed(path, **out)
    print("\n".join(report))
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))
    bad = [r for r in report if "MISMATCH" in r or "EXCEEDS" in r]
    print("SUMMARY: %d comparisons, %d failed" % (len(report), len(bad)))


if __name__ == "__main__":
    main()

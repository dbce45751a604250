#!/usr/bin/env python
"""bench.py -- headline benchmark of the VoGE ray-tracing hot path on B200.

Metric (BASELINE.json): fwd+bwd Mrays/s on the C5 synthetic scale sweep -- 1M Gaussians, 1024x1024,
64 views (sharded by camera over the ranks: strong scaling), K=20, thr=0.01.  One "step" = one
fitting step over all 64 views: for every chunk of views render fragments (fused CUDA path), composite
against a white background, MSE against a target image, backward to verts / sigmas / colours, then (N>1)
one NCCL all-reduce of the parameter gradients.

  python bench.py [--gpus N --steps K --warmup W]        # our arm (torchrun for N > 1)
  python bench.py --impl reference [...]                 # CPU port of the reference path (see cpu_baseline)

Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for the definition of every field.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "fwd+bwd Mrays/s, 1M Gaussians 1024²×64 views, 1/2/4/8 B200; % FP32/SFU peak"
FLOP_PER_PAIR = 33.0        # algorithmic forward ray-trace work per (ray, candidate) pair
FLOP_PER_HIT_BWD = 110.0    # recompute + chain rule per selected hit


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="voge_b200", choices=["voge_b200", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--hw", type=int, default=1024)
    ap.add_argument("--views", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=64, help="views rendered per renderer call (capped by the views of the rank); the 64-view batch of C5 is ONE call at N=1 (21 GB of fragments + scratch out of 180 GB)")
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--config", default="c5", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="BASELINE.json configs[0..4]; c5 (default) is the configuration the metric is quoted on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """index of the next sample (nvidia-smi needs ~0.5 s to start: the sampler is started before the warm-up and
        the timed region is cut out by sample index)"""
        return len(self.rows)

    def stop(self, first=0, last=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows[first:last]:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


class OpTimer:
    """CUDA-event timing of individual C-ABI ops on the launching (current) stream."""

    def __init__(self):
        self.events = {}
        self.enabled = False

    def wrap(self, module, name):
        fn = getattr(module, name)
        timer = self

        def wrapped(*a, **k):
            if not timer.enabled:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            timer.events.setdefault(name, []).append((e0, e1))
            return out
        setattr(module, name, wrapped)

    def add(self, name, e0, e1):
        """kernel-level entry (voge_b200._lib.kernel_timer protocol): one C-ABI call"""
        self.events.setdefault(name, []).append((e0, e1))

    def summary(self):
        out = {}
        for name, evs in self.events.items():
            ms = [a.elapsed_time(b) for a, b in evs]
            out[name] = {"launches": len(ms), "total_ms": sum(ms), "avg_ms": sum(ms) / max(len(ms), 1)}
        return out

    def reset(self):
        self.events = {}


# ------------------------------------------------------------------------------------------------
def build_workload(args, dev, views):
    """views: the orbit indices this rank renders (voge_b200.distributed.shard_view_indices)"""
    from voge_b200 import scenes
    from voge_b200.Meshes import GaussianMeshes
    from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings
    H = W = args.hw
    verts, sig, colors = scenes.synthetic_scene(args.n, seed=0)
    focal = 900.0 * args.hw / 1024.0
    settings = GaussianRenderSettings(image_size=(H, W), max_assign=args.k, thr_activation=0.01, absorptivity=1)
    renderers, targets = [], []
    for c0 in range(0, len(views), args.chunk):
        chunk = views[c0:c0 + args.chunk]
        cams = scenes.orbit_cameras(args.views, dist=3.0, elev_amp=20.0, focal=focal, image_size=(H, W), device=dev,
                                    indices=chunk)
        renderers.append(GaussianRenderer(cams, settings).to(dev))
        tg = [torch.rand(H, W, 3, generator=torch.Generator().manual_seed(1000 + v)) for v in chunk]
        targets.append(torch.stack(tg))
    gm = GaussianMeshes(verts, sig).to(dev)
    col = torch.nn.Parameter(colors.to(dev))
    from voge_b200.distributed import GradientBucket
    bucket = GradientBucket([gm.verts, gm.sigmas, col])
    return dict(gm=gm, colors=col, renderers=renderers, targets_host=targets, H=H, W=W, verts_host=verts,
                sig_host=sig, colors_host=colors, bucket=bucket)


def fit_step(wl, targets_dev, n_views_total, wait_events=None):
    """fwd + bwd over this rank's views; gradients accumulate straight into the flat bucket whose slices are
    gm.verts/.sigmas/.colors .grad (voge_b200.distributed.GradientBucket), then ONE in-place all-reduce.
    wait_events: optional per-chunk CUDA events (e2e: the chunk's target images have arrived) -- waited for right
    before the loss, the only consumer of the targets."""
    from voge_b200.Renderer import to_white_background
    gm, col = wl["gm"], wl["colors"]
    bucket = wl["bucket"]
    bucket.zero()
    total = None
    main = torch.cuda.current_stream()
    for i, (renderer, tgt) in enumerate(zip(wl["renderers"], targets_dev)):
        frag = renderer(gm)
        img = to_white_background(frag, col)
        if wait_events is not None:
            main.wait_event(wait_events[i])
        loss = torch.nn.functional.mse_loss(img, tgt, reduction="sum") / (n_views_total * wl["H"] * wl["W"] * 3)
        loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
    bucket.allreduce()
    return total


def count_ref_pairs(wl, args, dev):
    """N_pairs under the REFERENCE's coarse semantics (bin_size from RayTracing.py:14-16): sum over bins
    of candidates(bin) x in-image pixels(bin), per SURVEY.md 8(d).  Uses the API-compatible coarse op
    (true per-bin counts, M=0 so nothing is filled)."""
    from voge_b200 import _C
    from voge_b200.Aggregation import expend_sigma
    from voge_b200.RayTracing import coarse_inputs, default_bin_size
    from voge_b200.cameras import generate_rays
    H, W = wl["H"], wl["W"]
    bs = default_bin_size((H, W))
    total = 0
    gm = wl["gm"]
    with torch.no_grad():
        isg = (2 * expend_sigma(gm.sigmas))[None]
        for renderer in wl["renderers"]:
            cams = renderer.cameras
            for b in range(cams.R.shape[0]):
                from voge_b200.cameras import PerspectiveCameras
                cam1 = PerspectiveCameras(focal_length=cams.focal_length[b:b + 1], principal_point=cams.principal_point[b:b + 1],
                                          R=cams.R[b:b + 1], T=cams.T[b:b + 1], in_ndc=False,
                                          image_size=((H, W),), device=dev)
                _, origin = generate_rays(cam1, (H, W))
                mus = gm.verts[None] - origin[:, None]
                ndc, boxes = coarse_inputs(cam1, mus, isg, 0.01)
                first = torch.zeros(1, dtype=torch.long, device=dev)
                nper = torch.full((1,), mus.shape[1], dtype=torch.long, device=dev)
                _, counts = _C.rasterize_points_coarse(ndc.reshape(-1, 3), first, nper, (H, W), boxes.reshape(-1, 2), bs, 0,
                                                       return_counts=True, check_overflow=False)
                BH, BW = counts.shape[1], counts.shape[2]
                ph = torch.clamp(torch.tensor(H, device=dev) - torch.arange(BH, device=dev) * bs, max=bs)
                pw = torch.clamp(torch.tensor(W, device=dev) - torch.arange(BW, device=dev) * bs, max=bs)
                total += int((counts[0].long() * (ph[:, None] * pw[None, :])).sum().item())
    return total, bs


def measure_peaks(dev):
    """In-run FFMA and MUFU.EX2 peaks (MEASURED_PEAKS.json has HBM and bf16-GEMM only)."""
    from voge_b200._lib import check, lib, ptr, stream_of
    out = torch.zeros(4, device=dev)
    blocks, iters = 148 * 16, 4096
    res = {}
    for name, fn, per in (("fp32_tflops", lib().voge_peak_fp32, 32.0), ("sfu_tops", lib().voge_peak_sfu, 8.0)):
        best = 0.0
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(fn(blocks, iters, ptr(out), stream_of(out)), "peak")
            e1.record()
            torch.cuda.synchronize()
            best = max(best, blocks * 256 * iters * per / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        res[name] = best
    return res


# ------------------------------------------------------------------------------------------------
def cpu_port_sample(args, repeats=1):
    """The reference ships NO CPU ray tracer (ray_trace_voge.h:28-30).  CPU baseline = the oracle port:
    C/OpenMP restatement of the coarse + fine kernels and of the backward kernel (oracle/voge_oracle.c)
    plus the reference's own PyTorch Aggregation maths (oracle transcription, bit-identical to
    Aggregation.py on CPU), on a bounded sample of the C5 workload: view 0, a band of 128 full-width rows
    (fine / blend / backward) + the coarse stage of the full view charged pro rata."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import voge_oracle as vo
    from voge_b200 import scenes
    H = W = args.hw
    torch.set_num_threads(os.cpu_count())
    verts, sig, colors = scenes.synthetic_scene(args.n, seed=0)
    focal = 900.0 * args.hw / 1024.0
    R, T = vo.look_at_view(3.0, 0.0, 0.0)
    rays, origin = vo.camera_rays(R, T, focal, (W / 2.0, H / 2.0), (H, W))
    K, thr = args.k, 0.01
    bs = vo.default_bin_size((H, W))
    rows = min(128, H)
    y0 = ((H // 2 - rows // 2) // bs) * bs
    mus = (verts[None] - origin[:, None])
    isg = (2 * sig)[None]
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        ndc, radii = vo.coarse_inputs(R, T, focal, (W / 2.0, H / 2.0), (H, W), mus, isg, thr)
        first, nper = torch.zeros(1, dtype=torch.long), torch.full((1,), verts.shape[0])
        M = 8192
        bp, bc = vo.rasterize_coarse(ndc.reshape(-1, 3), radii.reshape(-1, 2), first, nper, (H, W), bs, M)
        t_coarse = time.perf_counter() - t0
        assert bc.max() <= M
        Msub = int(bc.max())
        bp_sub = torch.from_numpy(bp[:, y0 // bs:(y0 + rows) // bs, :, :Msub].copy())
        rays_sub = rays[:, y0:y0 + rows].contiguous()
        t1 = time.perf_counter()
        thr_act = -math.log(thr + 1e-10)
        idx, tl, ta, td = (torch.from_numpy(a) for a in vo.ray_trace_fine(mus.reshape(-1, 3), isg.reshape(-1, 3, 3), rays_sub,
                                                                           bp_sub, thr_act, bs, K))
        t_fine = time.perf_counter() - t1
        t2 = time.perf_counter()
        ta.requires_grad_(True); tl.requires_grad_(True); td.requires_grad_(True)
        col = colors.clone().requires_grad_(True)
        w, _, valid, _ = vo.aggregation_torch(idx, ta, tl, td, 1.0)
        img = vo.to_colored_background_torch(w, idx, valid, col, (1, 1, 1), -1)
        tgt = torch.rand(img.shape, generator=torch.Generator().manual_seed(1000))
        ((img - tgt) ** 2).mean().backward()
        t_blend = time.perf_counter() - t2
        t3 = time.perf_counter()
        vo.ray_trace_fine_backward(mus.reshape(-1, 3), isg.reshape(-1, 3, 3), rays_sub, idx, tl.grad, ta.grad, td.grad)
        t_bwd = time.perf_counter() - t3
        rays_n = rows * W
        t_total = t_coarse * rays_n / (H * W) + t_fine + t_blend + t_bwd
        cur = dict(t_coarse=t_coarse, t_fine=t_fine, t_blend_fwd_bwd=t_blend, t_geom_bwd=t_bwd, t_total=t_total,
                   rays=rays_n, mrays=rays_n / t_total / 1e6)
        if best is None or cur["mrays"] > best["mrays"]:
            best = cur
    best["cores"] = max(vo.num_threads(), torch.get_num_threads())
    best["sample"] = ("view 0 of C5 (N=%d, %dx%d, K=%d): rows %d..%d full width (%d rays) through C/OpenMP port of "
                      "coarse+fine+backward kernels and the reference's PyTorch Aggregation maths; coarse stage of the "
                      "full view charged pro rata; stages s: coarse %.2f fine %.2f blend(fwd+bwd) %.2f geom-bwd %.2f"
                      % (args.n, H, W, K, y0, y0 + rows, rays_n, t_coarse, t_fine, t_blend, t_bwd))
    return best


def ref_gpu_sample(wl, args, dev):
    """The reference's OWN CUDA kernels (oracle/_ref, unmodified, compiled for sm_100a) + its PyTorch
    aggregation on this B200 for view 0: fine kernel fed by our coarse op's bin_points (the reference's
    coarse kernel cannot launch at 32x32 bins).  Informational bar, not the CPU arm."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import build_ref
        import voge_oracle as vo
        ref = build_ref.load_ref()
    except Exception as e:
        return {"unavailable": repr(e)[:200]}
    from voge_b200 import _C
    from voge_b200.Aggregation import expend_sigma
    from voge_b200.RayTracing import coarse_inputs, default_bin_size
    from voge_b200.cameras import PerspectiveCameras, generate_rays
    H, W = wl["H"], wl["W"]
    gm = wl["gm"]
    cams = wl["renderers"][0].cameras
    cam1 = PerspectiveCameras(focal_length=cams.focal_length[:1], principal_point=cams.principal_point[:1], R=cams.R[:1],
                              T=cams.T[:1], in_ndc=False, image_size=((H, W),), device=dev)
    bs = default_bin_size((H, W))
    K = args.k
    thr_act = -math.log(0.01 + 1e-10)
    tgt = wl["targets_host"][0][:1].to(dev)
    col = wl["colors"].detach().clone().requires_grad_(True)
    times = []
    for it in range(3):
        verts = gm.verts.detach().clone().requires_grad_(True)
        sig = gm.sigmas.detach().clone().requires_grad_(True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rays, origin = generate_rays(cam1, (H, W))
        mus = verts[None] - origin[:, None]
        isg = (2 * expend_sigma(sig))[None]
        with torch.no_grad():
            ndc, boxes = coarse_inputs(cam1, mus, isg, 0.01)
            first = torch.zeros(1, dtype=torch.long, device=dev)
            nper = torch.full((1,), mus.shape[1], dtype=torch.long, device=dev)
            M = 8192
            bp = _C.rasterize_points_coarse(ndc.reshape(-1, 3), first, nper, (H, W), boxes.reshape(-1, 2), bs, M,
                                            check_overflow=False)

        class F(torch.autograd.Function):
            @staticmethod
            def forward(ctx, m, s, r):
                i, l, a, d = ref.ray_trace_voge_fine(m, s, r, bp, thr_act, bs, K)
                ctx.save_for_backward(m, s, r, i)
                ctx.mark_non_differentiable(i)
                return i, l, a, d

            @staticmethod
            def backward(ctx, gi, gl, ga, gd):
                m, s, r, i = ctx.saved_tensors
                gr, gm_, gs = ref.ray_trace_voge_fine_backward(m, s, r, i, gl.contiguous(), ga.contiguous(), gd.contiguous())
                return gm_, gs, None
        idx, tl, ta, td = F.apply(mus.reshape(-1, 3), isg.reshape(-1, 3, 3).contiguous(), rays)
        w, _, valid, _ = vo.aggregation_torch(idx, ta, tl, td, 1.0)     # == reference Aggregation.py ops, on the GPU
        img = vo.to_colored_background_torch(w, idx, valid, col, (1, 1, 1), -1)
        ((img - tgt) ** 2).mean().backward()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        del idx, tl, ta, td, w, img
    ms = min(times)
    return {"value": H * W / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_view": ms,
            "sample": "view 0, fwd+bwd: reference CUDA fine fwd/bwd kernels (unmodified, sm_100a) + reference PyTorch "
                      "aggregation/merge on this GPU; coarse bins from voge_b200's coarse op"}


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """--impl reference: the CPU port on all host cores; every step = one bounded sample (a 128-row band of view 0),
    W untimed warm-up samples first, then exactly K timed ones; the line reports their mean."""
    if rank != 0:
        return
    t0 = time.perf_counter()
    for _ in range(max(args.warmup, 0)):
        cpu_port_sample(args)
    samples = [cpu_port_sample(args) for _ in range(max(args.steps, 1))]
    t_mean = sum(c["t_total"] for c in samples) / len(samples)
    best = samples[0]
    mrays = best["rays"] / t_mean / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": len(samples), "warmup": max(args.warmup, 0), "ms_per_step": t_mean * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5 synthetic scale sweep (bounded sample per step): " + best["sample"]},
            "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": best["cores"], "kind": "port",
                             "sample": best["sample"]},
            "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_small_config(args):
    """BASELINE.json configs[0..3] as bench lines (informational: the driver times c5).  These scenes are launch- and
    host-sync-bound (a handful of tiles per view), so the line reports latency per step next to the rate.
      c1  quick-start cuboid (Readme.md:81-94): 866 Gaussians, 256^2, K=20, one view, forward + sample_features
      c2  RenderBunny (demo/RenderBunny.py:17-38 at BASELINE's size): 40 962 mesh-converted Gaussians, 512^2, K=40, fwd+bwd
      c3  ShapeFitting (demo/ShapeFitting.py:219-296): ico_sphere(4) = 2 562 Gaussians, 128^2, K=25, no coarse stage,
          8 views per step sharded over the ranks, fwd+bwd, gradient all-reduce, SGD(lr .8, momentum .9)
      c4  ReasonOcclusion (demo/ReasonOcclusion.py:27-55): two cuboids = 6 778 Gaussians, 400^2, K=60, M=1500, fwd+bwd
          to the vertices"""
    import numpy as np
    from voge_b200 import _lib, scenes
    from voge_b200.cameras import PerspectiveCameras, look_at_view_transform
    from voge_b200.distributed import GradientBucket, barrier, init_from_env, max_over_ranks, shard_views
    from voge_b200.Meshes import GaussianMeshes
    from voge_b200.Renderer import GaussianRenderer, GaussianRenderSettings, to_white_background
    from voge_b200.Sampler import sample_features
    rank, world, local = init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    _lib.lib()
    cfg = args.config
    g = torch.Generator().manual_seed(0)
    views, backward, optim, sample = 1, True, False, False
    if cfg == "c1":
        v, sg = scenes.cuboid_gauss((-1, 1), (-1, 1), (-1, 1), 1000, percentage=0.6)
        verts, sig = torch.tensor(v, dtype=torch.float32), torch.tensor(sg, dtype=torch.float32)
        hw, K, M, focal, cam = (256, 256), 20, None, 300.0, (6.0, [10.0], [70.0])
        backward, sample = False, True
        desc = "C1 quick-start cuboid: 866 Gaussians, 256x256, K=20, 1 view, forward + to_white_background + sample_features"
    elif cfg == "c2":
        from voge_b200.Converter.Converters import naive_vertices_converter
        mv, mf = scenes.ico_sphere(6)
        rng = np.random.RandomState(0)
        dirs = rng.randn(6, 3); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        bump = 1.0 + 0.12 * sum(np.sin(3.0 * mv @ d + i) for i, d in enumerate(dirs)) / len(dirs)
        vn, sn, _ = naive_vertices_converter(mv * bump[:, None] * 0.3, mf, percentage=0.6)
        verts, sig = torch.tensor(vn, dtype=torch.float32), torch.tensor(sn, dtype=torch.float32)
        hw, K, M, focal, cam = (512, 512), 40, None, 4000.0, (6.0, [0.0], [10.0])
        desc = "C2 RenderBunny-sized: 40 962 mesh-converted Gaussians (naive_vertices_converter), 512x512, K=40, 1 view, fwd+bwd"
    elif cfg == "c3":
        verts = torch.tensor(scenes.ico_sphere(4)[0], dtype=torch.float32)
        sig = torch.full((verts.shape[0],), 400.0)
        views = 8
        hw, K, M, focal = (128, 128), 25, -1, 150.0
        cam = (2.7, [20.0 * math.sin(2 * math.pi * i / 8) for i in range(8)], [45.0 * i for i in range(8)])
        optim = True
        desc = ("C3 ShapeFitting step: ico_sphere(4) = 2 562 Gaussians, 128x128, K=25, no coarse stage, 8 views per step "
                "sharded by camera, fwd+bwd + gradient all-reduce + SGD")
    else:
        v1, s1 = scenes.cuboid_gauss((-0.6, 0.6), (-0.4, 0.4), (-0.5, 0.5), 1500, percentage=0.6)
        v2, s2 = scenes.cuboid_gauss((-0.5, 0.5), (-0.5, 0.5), (-0.3, 0.3), 1200, percentage=0.6)
        verts = torch.tensor(np.concatenate([v1, v2 + np.array([0.4, 0.1, -0.9])]), dtype=torch.float32)
        sig = torch.tensor(np.concatenate([s1, s2]), dtype=torch.float32)
        hw, K, M, focal, cam = (400, 400), 60, 1500, 300.0, (4.0, [15.0], [30.0])
        desc = "C4 ReasonOcclusion: two cuboids = %d Gaussians, 400x400, K=60, M=1500, 1 view, fwd+bwd" % verts.shape[0]
    H, W = hw
    first, count = shard_views(views, rank, world)
    count = max(count, 0)
    R, T = look_at_view_transform(dist=cam[0], elev=torch.tensor(cam[1]), azim=torch.tensor(cam[2]))
    renderer = None
    if count > 0:
        cams = PerspectiveCameras(focal_length=focal, principal_point=((W / 2.0, H / 2.0),), R=R[first:first + count],
                                  T=T[first:first + count], in_ndc=False, image_size=((H, W),), device=dev)
        renderer = GaussianRenderer(cams, GaussianRenderSettings(image_size=hw, max_assign=K, max_point_per_bin=M)).to(dev)
    gm = GaussianMeshes(verts.clone(), sig.clone()).to(dev)
    col = torch.nn.Parameter(torch.rand(verts.shape[0], 3, generator=g).to(dev))
    bucket = GradientBucket([gm.verts, gm.sigmas, col])
    opt = torch.optim.SGD([gm.verts, gm.sigmas, col], lr=0.8 * 1e-4, momentum=0.9) if optim else None
    target = torch.rand(max(count, 1), H, W, 3, generator=g).to(dev)

    def core():
        # the part of the step that a CUDA graph can hold: renderer forward, composite, loss, fused backward
        bucket.zero()
        out = None
        if renderer is not None:
            if backward:
                frag = renderer(gm)
                img = to_white_background(frag, col)
                out = torch.nn.functional.mse_loss(img, target[:count], reduction="sum") / (views * H * W * 3)
                out.backward()
            else:
                with torch.no_grad():
                    frag = renderer(gm)
                    img = to_white_background(frag, col)
                    out = img.mean()
                    if sample:
                        feat, wsum = sample_features(frag, img, n_vert=verts.shape[0])
                        out = out + feat.mean()
        return out

    def tail():
        if backward:
            bucket.allreduce()
        if opt is not None:
            opt.step()

    def step():
        out = core()
        tail()
        return out
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(); barrier()
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize(); barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    ms_step = max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    launches_eager = int(_lib.launch_count - l0)
    # per-op breakdown (untimed extra steps with CUDA events around every C-ABI call)
    timer = OpTimer()
    _lib.kernel_timer = timer
    timer.enabled = True
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    timer.enabled = False
    ops = {k: round(v["total_ms"] / 3, 4) for k, v in timer.summary().items()}
    # the same step with core() replayed from ONE CUDA graph (voge_b200.graphs.GraphedStep; the all-reduce and the
    # optimizer stay outside): what the launch-bound configurations cost once the host is out of the way
    graphed = None
    _lib.kernel_timer = None
    loss_eager = float(out.detach()) if out is not None else None
    out = None          # drop the eager autograd graph (its AccumulateGrad nodes live on the default stream)
    if renderer is not None and os.environ.get("VOGE_BENCH_NO_GRAPH") != "1":
        from voge_b200.graphs import GraphedStep
        gs = GraphedStep(core, device=dev)

        def gstep():
            out = gs()
            tail()
            return out
        for _ in range(max(args.warmup, 3)):
            gstep()
        torch.cuda.synchronize(); barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tg = time.perf_counter()
        g0.record()
        for _ in range(args.steps):
            gout = gstep()
        g1.record()
        torch.cuda.synchronize(); barrier()
        graphed = {"ms_per_step": max_over_ranks(g0.elapsed_time(g1), dev) / args.steps,
                   "host_wall_ms_per_step": (time.perf_counter() - tg) * 1e3 / args.steps,
                   "kernel_launches_per_replay": int(gs.launches_per_replay), "recaptures": int(gs.recaptures),
                   "validated_every_replay": True,
                   "loss_eager": loss_eager, "loss_graphed": float(gout) if gout is not None else None}
    if rank == 0:
        line = {"metric": ("fwd+bwd" if backward else "fwd") + " Mrays/s", "value": views * H * W / (ms_step * 1e-3) / 1e6,
                "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
                "host_wall_ms_per_step": wall_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": {"workload": desc, "views_per_rank": count},
                "gpu_launches": launches_eager, "gpu_launches_per_step": launches_eager // max(args.steps, 1),
                "roofline": None, "cpu_baseline": None, "e2e": None, "op_breakdown_ms_per_step": ops, "graphed": graphed,
                "note": "informational line for BASELINE.json configs[%d]; small scenes are bound by kernel launches and "
                        "the one host sync of the binning, not by a pipe" % (int(cfg[1]) - 1)}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    from voge_b200.distributed import barrier, init_from_env, max_over_ranks, shard_view_indices
    if args.impl != "reference" and args.config != "c5":
        run_small_config(args)
        return
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference(args, rank)
        return
    rank, world, local = init_from_env("nccl")
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: voge_b200 has no CPU path")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    from voge_b200 import _C, _lib
    _lib.lib()
    my_views = shard_view_indices(args.views, rank, world)     # round-robin: every rank gets the same mix of cameras
    count = len(my_views)
    wl = build_workload(args, dev, my_views)
    H, W = wl["H"], wl["W"]
    rays_total = args.views * H * W

    timer = OpTimer()
    for name in ("bin_views", "render_forward"):        # host-side ops that span several launches
        timer.wrap(_C, name)
    _lib.kernel_timer = timer                           # every C-ABI call (= one kernel launch) individually

    targets_dev = [t.to(dev) for t in wl["targets_host"]]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        fit_step(wl, targets_dev, args.views)
    torch.cuda.synchronize()

    # ---- timed region: value (inputs resident in HBM) ----
    barrier(); torch.cuda.synchronize()
    mark0 = sampler.mark()
    launches0 = _lib.launch_count
    timer.enabled = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = fit_step(wl, targets_dev, args.views)
    e1.record()
    torch.cuda.synchronize(); barrier()
    timer.enabled = False
    ms_total = max_over_ranks(e0.elapsed_time(e1), dev)
    launches = _lib.launch_count - launches0        # C-ABI kernel launches inside the timed region (all steps)
    timer.enabled = False
    mark1 = sampler.mark()
    extra = 0
    while rank == 0 and world == 1 and sampler.proc is not None and sampler.mark() - mark0 < 5 and extra < 40:
        # a timed region shorter than a few 50 ms sampling periods: keep the same load running (untimed) until the
        # sampler has seen it
        fit_step(wl, targets_dev, args.views)
        torch.cuda.synchronize()
        extra += 1
        mark1 = sampler.mark()
    clocks = sampler.stop(mark0, max(mark1, mark0 + 1)) if rank == 0 else None
    if clocks is not None:
        clocks["sampled_over"] = "timed region" + (" + %d identical untimed steps" % extra if extra else "")
    ms_step = ms_total / args.steps
    value = rays_total / (ms_step * 1e-3) / 1e6
    ops = timer.summary()

    # ---- e2e: same step through the public API with HOST buffers (pinned), H2D inside the timed region ----
    e2e = None
    if not args.no_e2e:
        pin = lambda t: t.contiguous().pin_memory()
        h_verts, h_sig, h_col = pin(wl["verts_host"]), pin(wl["sig_host"]), pin(wl["colors_host"])
        h_targets = [pin(t) for t in wl["targets_host"]]
        h2d = sum(t.numel() * 4 for t in h_targets)
        loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
        loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
        losses = []

        copy_stream = torch.cuda.Stream(device=dev)
        with torch.no_grad():     # model state lives on the device (as in a fitting loop); uploaded once, outside the timing
            wl["gm"].verts.copy_(h_verts); wl["gm"].sigmas.copy_(h_sig); wl["colors"].copy_(h_col)

        def e2e_step(i):
            # the step's INPUTS (target images) come from pinned host memory on a copy stream, one event per chunk;
            # the compute stream waits for a chunk's targets right before the loss that consumes them, so the
            # upload of chunk i overlaps the rendering of chunks <= i.  The step's RESULT (the loss) is copied to
            # pinned host memory every step and read by the host one step later (as a training loop logs it), so the
            # host keeps queueing work instead of draining the device once per step.
            main = torch.cuda.current_stream(dev)
            copy_stream.wait_stream(main)        # previous step must be done with the buffers we overwrite
            with torch.cuda.stream(copy_stream), torch.no_grad():
                tdev, evs = [], []
                for t in h_targets:
                    d = t.to(dev, non_blocking=True)
                    d.record_stream(main)
                    e = torch.cuda.Event(); e.record(copy_stream)
                    tdev.append(d); evs.append(e)
            total = fit_step(wl, tdev, args.views, wait_events=evs)
            loss_host[i & 1].copy_(total.reshape(1), non_blocking=True)
            loss_ready[i & 1].record(main)
            if i > 0:
                loss_ready[(i - 1) & 1].synchronize()
                losses.append(float(loss_host[(i - 1) & 1][0]))

        def e2e_drain(i_last):
            loss_ready[i_last & 1].synchronize()
            losses.append(float(loss_host[i_last & 1][0]))
        for i in range(2):
            e2e_step(i)
        e2e_drain(1)
        barrier(); torch.cuda.synchronize()
        del losses[:]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            e2e_step(i)
        e2e_drain(args.steps - 1)
        e1.record()
        torch.cuda.synchronize(); barrier()
        assert len(losses) == args.steps and all(math.isfinite(v) for v in losses)
        ms_e2e = max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
        e2e = {"value": rays_total / (ms_e2e * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
               "losses_read_on_host": len(losses),
               "note": "per step: this rank's target images from pinned host memory (copy stream, one event per chunk, "
                       "waited for right before the loss) + the loss copied to pinned host memory and read by the host "
                       "one step later (all reads inside the timed region); Gaussian parameters are device-resident "
                       "model state, uploaded once before the timed steps (they do not change between steps)"}

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (rank 0; work counted for rank 0's views) ----
    peaks = measure_peaks(dev)
    n_pairs, ref_bin = count_ref_pairs(wl, args, dev)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    with torch.no_grad():
        orig = _C.render_forward
        _C.render_forward = lambda *a, **k: orig(*a, **{**k, "stats": stats})
        hits, valid_sq = 0, 0
        for r in wl["renderers"]:
            vn = r(wl["gm"]).valid_num
            hits += int(vn.sum().item())
            valid_sq += int((vn * vn).sum().item())            # sum over rays of v^2: the (m, k) pairs that exist
        _C.render_forward = orig
    torch.cuda.synchronize()
    n_calls = max(len(wl["renderers"]), 1)
    views_per_launch = count / n_calls
    rays_per_launch = views_per_launch * H * W
    pairs_per_launch = n_pairs / n_calls
    hits_per_launch = hits / n_calls
    vsq_per_launch = valid_sq / n_calls
    K = args.k
    mp = {}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = mp.get("hbm_gbs") or 6531.9     # fallback: B200_PROFILING.md's measured copy bandwidth
    # per-kernel counters of ONE `ncu --set full` capture of the C5 scene (profiles/pipes_r2.json, written from the
    # .ncu-rep by tools/ncu_pipes.py): DRAM bytes per view and what the hardware actually issued -- FMA / XU pipe,
    # issue slots, L1 LSU wavefronts, L2 tag requests, all in % of peak.  null for other workloads.
    pipes = {}
    try:
        if args.n == 1_000_000 and args.hw == 1024 and args.k == 20:
            pipes = json.load(open(os.path.join(ROOT, "profiles", "pipes_r2.json")))
    except Exception:
        pipes = {}

    def kernel_entry(sym, label, bound, work, unit_scale, note, valid_work=None):
        """work = algorithmic FLOP (fp32) or bytes (hbm) per launch by SURVEY 8(d)'s accounting; unit_scale 1e12 / 1e9;
        valid_work = the same count over the (m, k) pairs that exist (sum v^2 instead of K^2 per ray)"""
        o = ops.get(sym)
        if not o or not o["launches"]:
            return None
        # time per renderer CALL (the forward traces / selects a long batch in groups of views: several launches)
        o = dict(o)
        o["launches_per_call"] = o["launches"] / (args.steps * n_calls)
        o["avg_ms"] = o["total_ms"] / (args.steps * n_calls)
        ach = work / (o["avg_ms"] * 1e-3) / unit_scale
        peak = peaks["fp32_tflops"] if bound == "fp32" else hbm_peak
        pk = pipes.get(sym) or {}
        e = {"kernel": label, "bound": bound, "achieved": ach, "peak": peak,
             "unit": "TFLOP/s" if bound == "fp32" else "GB/s", "frac": ach / peak,
             "traffic": (pk["dram_bytes_per_view"] * views_per_launch) if pk.get("dram_bytes_per_view") else None,
             "avg_launch_ms": o["avg_ms"], "launches_timed": o["launches"], "launches_per_call": o["launches_per_call"],
             "share_of_step": o["total_ms"] / max(ms_total, 1e-9), "algorithmic_work_per_launch": work, "work": note,
             # what the pipes actually did (ncu, % of peak) -- `frac` above is algorithmic work over time and, for the
             # kernels that cull or window their work, is NOT pipe efficiency
             "issued": {k: pk.get(k) for k in ("fma_pipe_pct", "xu_pipe_pct", "issue_active_pct", "l1_wavefront_pct",
                                               "l2_tag_request_pct", "dram_pct")} if pk else None}
        if valid_work is not None:
            e["valid_pair_work_per_launch"] = valid_work
            e["valid_pair_frac"] = valid_work / (o["avg_ms"] * 1e-3) / unit_scale / peak
        return e

    frag_bytes = (12 * K + 8) * rays_per_launch
    items_per_launch = int(stats[0].item()) / n_calls
    bwd_sym = "voge_render_backward_image" if "voge_render_backward_image" in ops else "voge_render_backward_fused"
    kernels = [
        kernel_entry("voge_trace_hits", "trace_hits_kernel (Gaussian-major exact ray trace -> per-pixel hit segments)",
                     "fp32", FLOP_PER_PAIR * pairs_per_launch, 1e12,
                     "33 FLOP x N_pairs under the reference's coarse semantics (SURVEY 8d); the kernel evaluates only "
                     "the items of the culled pixel rectangles: valid_pair_* = 33 FLOP x items actually evaluated",
                     valid_work=FLOP_PER_PAIR * items_per_launch),
        kernel_entry(bwd_sym, "render_bwd_pair_kernel (recompute + analytic blend backward + chain rule"
                     + (" + merge_final backward, image mode)" if bwd_sym.endswith("image") else ")"),
                     "fp32", FLOP_PER_HIT_BWD * hits_per_launch + (K * K * 30.0) * rays_per_launch, 1e12,
                     "110 FLOP per hit + K^2 x 30 FLOP per ray (dense K x K blend backward of the reference, SURVEY 8d); "
                     "valid_pair_*: sum over rays of v^2 x 30 instead of K^2 x 30",
                     valid_work=FLOP_PER_HIT_BWD * hits_per_launch + 30.0 * vsq_per_launch),
        kernel_entry("voge_blend_weights", "blend_pair_kernel (exact re-evaluation + windowed erf blend)",
                     "fp32", 33.0 * hits_per_launch + (K * K * 20.0 + K * 8.0) * rays_per_launch, 1e12,
                     "33 FLOP per hit + (K^2 x 20 + K x 8) FLOP per ray (SURVEY 8d); valid_pair_*: v^2 x 20 + v x 8",
                     valid_work=33.0 * hits_per_launch + 20.0 * vsq_per_launch + 8.0 * hits_per_launch),
        kernel_entry("voge_select_topk", "select_topk_kernel (register sorting networks)", "hbm",
                     8.0 * hits_per_launch * 1.6 + (4 * K + 8 + 12) * rays_per_launch, 1e9,
                     "reads the stored hits (8 B each; ~1.6 stored per selected) + 12 B/pixel of segment tables, "
                     "writes (4K+8) B/ray"),
        kernel_entry("voge_merge_final", "merge_fwd_small_kernel (gather-blend + composite)", "hbm",
                     (8 * K + 8 + 12) * rays_per_launch, 1e9, "8K B/ray in + 4C B/ray out (SURVEY 8d)"),
        kernel_entry("voge_merge_final_backward", "merge_bwd_small_kernel", "hbm",
                     (8 * K + 8 + 12 + 4 * K) * rays_per_launch, 1e9, "8K B/ray + grad in, 4K B/ray grad_weight out"),
    ]
    kernels = [k for k in kernels if k is not None]
    # SFU (MUFU) side of the ray-trace kernels, SURVEY 8d: 1 rcp per pair forward; K^2 x 2 per ray in the blend backward
    for k in kernels:
        sfu_ops = None
        if k["kernel"].startswith("trace_hits"):
            sfu_ops = 1.0 * pairs_per_launch
        elif k["kernel"].startswith("render_bwd"):
            sfu_ops = (K * K * 2.0) * rays_per_launch + 2.0 * hits_per_launch
        elif k["kernel"].startswith("blend_pair"):
            sfu_ops = (K * K * 1.0 + 2.0 * K) * rays_per_launch
        if sfu_ops is not None:
            k["sfu_achieved_tops"] = sfu_ops / (k["avg_launch_ms"] * 1e-3) / 1e12
            k["sfu_frac"] = k["sfu_achieved_tops"] / peaks["sfu_tops"]
    kernels.sort(key=lambda k: -k["share_of_step"])
    fwd_ms = ops.get("render_forward", {"avg_ms": float("nan")})["avg_ms"]
    # whole step by the same accounting (SURVEY 8d FLOPs of the three FP32 kernels over the step time)
    calls_per_step = n_calls
    step_flops = calls_per_step * sum(k["algorithmic_work_per_launch"] for k in kernels if k["bound"] == "fp32")
    step_valid = calls_per_step * sum(k.get("valid_pair_work_per_launch", 0.0) for k in kernels if k["bound"] == "fp32")
    roofline = dict(kernels[0]) if kernels else {"kernel": None}
    roofline.update({
        "peak_source": "in-run FFMA micro-benchmark (MEASURED_PEAKS.json holds HBM and bf16-GEMM only); "
                       "HBM peak from MEASURED_PEAKS.json" + ("" if mp.get("hbm_gbs") else " (absent: B200_PROFILING.md fallback)"),
        "traffic_source": "profiles/pipes_r2.json (ncu --set full, per view, scaled to the views of one launch)",
        "issued_source": pipes.get("_source"),
        "sfu_peak_tops": peaks["sfu_tops"],
        "whole_step": {"algorithmic_tflops": step_flops / (ms_step * 1e-3) / 1e12,
                       "frac_of_fp32_peak": step_flops / (ms_step * 1e-3) / 1e12 / peaks["fp32_tflops"],
                       "valid_pair_tflops": step_valid / (ms_step * 1e-3) / 1e12,
                       "valid_pair_frac_of_fp32_peak": step_valid / (ms_step * 1e-3) / 1e12 / peaks["fp32_tflops"],
                       "note": "SURVEY 8(d) FLOPs of trace + blend + backward over the step time; valid_pair_* counts the "
                               "items the trace evaluates and sum v^2 blend pairs instead of N_pairs and K^2"},
        "algorithmic_pairs_per_launch": pairs_per_launch, "flop_per_pair": FLOP_PER_PAIR,
        "reference_bin_size": ref_bin,
        "items_evaluated_per_launch": items_per_launch,
        "pixels_selected_with_exact_keys": int(stats[2].item()),
        "hits_per_launch": hits_per_launch,
        "valid_pairs_sum_v2_per_launch": vsq_per_launch,
        "forward_all_launches_ms": fwd_ms,
        "forward_algorithmic_tflops": FLOP_PER_PAIR * pairs_per_launch / (fwd_ms * 1e-3) / 1e12,
        "fragment_write_GBps": frag_bytes / (fwd_ms * 1e-3) / 1e9,
        "hbm_peak_GBps": hbm_peak,
        "kernels": kernels,
        "op_breakdown_ms_per_step": {k: v["total_ms"] / args.steps for k, v in ops.items()},
    })
    cpu = None
    if world == 1 and not args.no_cpu_baseline:      # the CPU arm is timed at N=1 only
        c = cpu_port_sample(args)
        cpu = {"value": c["mrays"], "unit": "Mrays/s", "cores": c["cores"], "kind": "port", "sample": c["sample"]}
    line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5 synthetic scale sweep: %d Gaussians (10%% anisotropic, (N,3,3) sigmas), %dx%d, "
                                   "%d views sharded by camera (round-robin over the ranks), K=%d, thr=0.01, fwd+bwd to verts/sigmas/colours"
                                   % (args.n, H, W, args.views, args.k),
                       "views_per_rank": count, "views_per_call": min(args.chunk, count),
                       "l2": "no flush needed: each renderer call streams %.0f MB of fragments (+ %d MB targets), "
                             ">> 126 MB L2" % (frag_bytes / 1e6, min(args.chunk, count) * H * W * 12 // 10 ** 6)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "gpu_launches_per_step": int(launches) // max(args.steps, 1), "clocks": clocks,
            "loss": float(loss), "peak_hbm_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
    if world == 1 and not args.no_ref_gpu:
        # the bar that matters: the reference's OWN CUDA kernels + PyTorch aggregation on this GPU (one view, fwd+bwd)
        rg = ref_gpu_sample(wl, args, dev)
        line["ref_gpu"] = rg
        if rg.get("value"):
            line["vs_ref_gpu"] = {"ratio": value / rg["value"], "e2e_ratio": (e2e["value"] / rg["value"]) if e2e else None,
                                  "note": "this arm (64 views, fused path) over the reference CUDA kernels (view 0, per-view "
                                          "rate; the reference has no multi-view batching: its cost is per view)"}
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

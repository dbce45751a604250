"""Builds libvoge_b200.so (hand-written CUDA kernels + C ABI) for sm_100a with nvcc, in-tree.

`python -m voge_b200.build` or voge_b200.build.build().  The .so is git-ignored but travels to
the GPU box with the gpurun snapshot.  No torch / pybind in the library: plain `extern "C"`.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvoge_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

SOURCES = ["api.cu", "coarse.cu", "fine_fwd.cu", "fine_bwd.cu", "blend.cu", "sample.cu", "render.cu", "trace.cu", "select.cu", "dense.cu", "knn.cu"]
HEADERS = ["common.cuh", "fine_core.cuh", "render_core.cuh", "blend_core.cuh", "sort_net.h", os.path.join("..", "..", "include", "voge_b200.h")]

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=default",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS if os.path.exists(os.path.join(CSRC, s))]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and not _stale():
        return LIB
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    t_hdr = max(os.path.getmtime(h) for h in hdrs if os.path.exists(h))
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in srcs:
        o = os.path.join(HERE, "build", os.path.basename(s) + ".o")
        objs.append(o)
        # per-object staleness: recompile only what changed (the sorting networks of select.cu take minutes)
        if not force and os.path.exists(o) and os.path.getmtime(o) > max(os.path.getmtime(s), t_hdr):
            continue
        cmd = [NVCC, *FLAGS, "-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        with open(os.path.join(HERE, "build", os.path.basename(s) + ".ptxas.log"), "w") as f:
            f.write(out)
        if verbose:
            print(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on " + s)
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        for s in srcs:
            lp = os.path.join(HERE, "build", os.path.basename(s) + ".ptxas.log")
            if os.path.exists(lp):
                f.write(open(lp).read() + "\n")
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference CUDA extension into oracle/_ref/.

Compiles the reference's own sources *where they lie* under /root/reference/VoGE/csrc
(ext.cpp + the four .cu files; nothing is copied into this repo) with
torch.utils.cpp_extension for sm_100a, output `oracle/_ref/voge_ref_C.so`.
The resulting module exposes the reference's 9 pybind entry points
(reference VoGE/csrc/ext.cpp:7-17).  It executes only on a GPU, so it is used
(a) under gpurun to generate the golden vectors in tests/golden/ and
(b) by the `-m gpu` parity tests / bench `ref_gpu` leg as the ground truth.

Only tests/, __graft_entry__ and bench.py may import what this builds.
`oracle/_ref/` is git-ignored but NOT gpurun-ignored: the .so travels to the GPU box.

Deviation from the reference's own setup.py (which is not run): -std=c++17
(setup.py:11 asks c++14, which torch 2.x headers reject) and an explicit
-gencode for sm_100a (setup.py passes no arch flag).
"""
import os
import sys
import shutil

REF = os.environ.get("VOGE_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NAME = "voge_ref_C"


def ref_so_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False, force=False):
    csrc = os.path.join(REF, "VoGE", "csrc")
    if not os.path.isdir(csrc):
        return None  # reference tree not present (e.g. on the GPU box): use the prebuilt .so
    if os.path.exists(ref_so_path()) and not force:
        return ref_so_path()
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "5")
    from torch.utils.cpp_extension import load
    srcs = [
        os.path.join(csrc, "ext.cpp"),
        os.path.join(csrc, "rasterize_coarse", "rasterize_coarse.cu"),
        os.path.join(csrc, "ray_trace_voge", "ray_trace_voge.cu"),
        os.path.join(csrc, "sample_voge", "sample_voge.cu"),
        os.path.join(csrc, "voge_ray_tracing_ray", "voge_ray_tracing_ray.cu"),
    ]
    bdir = os.path.join(OUT, "build")
    os.makedirs(bdir, exist_ok=True)
    load(name=NAME, sources=srcs, extra_include_paths=[csrc],
         extra_cflags=["-DWITH_CUDA", "-std=c++17", "-O2"],
         extra_cuda_cflags=["-DWITH_CUDA", "-std=c++17", "-O3",
                            "-gencode", "arch=compute_100a,code=sm_100a"],
         build_directory=bdir, verbose=verbose, is_python_module=False)
    shutil.copy(os.path.join(bdir, NAME + ".so"), ref_so_path())
    return ref_so_path()


def load_ref():
    """Import the prebuilt reference module (needs a GPU to *run* anything)."""
    import importlib.util
    import torch  # noqa: F401  (must be imported first: the .so links libtorch)
    p = ref_so_path()
    if not os.path.exists(p):
        raise FileNotFoundError(p + " missing: run `python oracle/build_ref.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose=True, force="--force" in sys.argv)
    print("reference _C:", p)

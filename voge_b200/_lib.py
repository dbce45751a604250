"""ctypes binding of libvoge_b200.so (the C ABI declared in include/voge_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
The library is located in-tree (voge_b200/libvoge_b200.so, built by voge_b200/build.py).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvoge_b200.so")

_c_float_p = ctypes.c_void_p
_c_int_p = ctypes.c_void_p
_I, _L, _F, _P = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

# name -> (restype, argtypes); must list EVERY symbol of include/voge_b200.h (tests check this)
SIGNATURES = {
    "voge_version": (_I, []),
    "voge_error_string": (ctypes.c_char_p, [_I]),
    "voge_device_sm_count": (_I, [_P]),
    "voge_peak_fp32": (_I, [_I, _I, _P, _P]),
    "voge_peak_sfu": (_I, [_I, _I, _P, _P]),
    "voge_rasterize_coarse_scratch_elems": (_L, [_I, _I, _I, _I, _I]),
    "voge_rasterize_coarse": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "voge_ray_trace_fine": (_I, [_P, _P, _P, _P, _F, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "voge_ray_trace_fine_counts": (_I, [_P, _P, _P, _P, _P, _F, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "voge_ray_trace_fine_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "voge_aggregation": (_I, [_P, _P, _P, _P, _F, _L, _I, _P, _P, _P]),
    "voge_aggregation_backward": (_I, [_P, _P, _P, _P, _F, _L, _I, _P, _P, _P, _P]),
    "voge_merge_final": (_I, [_P, _P, _P, _P, _P, _F, _L, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "voge_merge_final_backward": (_I, [_P, _P, _P, _P, _P, _F, _P, _P, _L, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "voge_sample": (_I, [_P, _P, _P, _L, _I, _I, _I, _P, _P, _P]),
    "voge_sample_backward": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _P, _P, _P]),
    "voge_scatter_max": (_I, [_P, _P, _L, _I, _I, _P, _P]),
    "voge_ray_trace_ray": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "voge_ray_trace_ray_backward": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "voge_find_nearest_k": (_I, [_P, _P, _P, _F, _I, _I, _I, _P, _P, _P, _P, _P]),
    "voge_knn_mean_dist": (_I, [_P, _I, _I, _F, _P, _P]),
    "voge_bin_sub": (_I, []),
    "voge_bin_count": (_I, [_P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _I, _I, _I, _I, _P, _P, _P]),
    "voge_bin_fill": (_I, [_P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _L, _P]),
    "voge_trace_threads": (_I, [_I]),
    "voge_bin_item_slack": (_I, []),
    "voge_pack_gaussians": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "voge_pack_attr": (_I, [_P, _I, _I, _P, _P, _P]),
    "voge_unpack_gradients": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "voge_generate_rays": (_I, [_P, _I, _I, _I, _P, _P]),
    "voge_trace_hits": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _L, _F, _I, _I, _I, _I, _I, _P, _P, _P, _L, _L, _P, _P]),
    "voge_select_topk": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "voge_blend_weights": (_I, [_P, _I, _P, _P, _P, _P, _P, _F, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "voge_render_backward_fused": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _F, _I, _I, _I, _I, _I, _P, _I, _P, _P,
                                        _P, _P, _P]),
    "voge_render_backward_image": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _I, _F, _I, _I, _I, _I,
                                        _I, _P, _I, _P, _P, _P, _P, _P]),
}

_lib = None
# optional: an object with `.enabled` and `.add(name, start_event, end_event)`; when enabled every C-ABI
# call is bracketed by CUDA events on the launching (current) stream (bench.py uses it for per-kernel times)
kernel_timer = None


class _Handle(object):
    """Thin proxy over the ctypes handle (adds the optional per-call CUDA-event timing)."""

    def __init__(self, h):
        self._h = h

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        kt = kernel_timer
        if kt is None or not kt.enabled or name in ("voge_error_string", "voge_version", "voge_trace_threads", "voge_bin_sub", "voge_bin_item_slack"):
            return fn

        def timed(*args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*args)
            e1.record()
            kt.add(name, e0, e1)
            return r
        return timed


def lib():
    """Load (once) and return the library handle; raises RuntimeError if the .so is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "voge_b200: %s not found. Build it with `python -m voge_b200.build` "
                "(there is no CPU or PyTorch fallback)." % LIB_PATH)
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = _Handle(h)
    return _lib


# number of kernels launched through the C ABI since import (bench.py reports it as gpu_launches)
KERNELS_PER_CALL = {"rasterize_coarse": 3}
launch_count = 0


def check(code, what):
    global launch_count
    launch_count += KERNELS_PER_CALL.get(what, 1)
    if code != 0:
        msg = lib().voge_error_string(int(code))
        raise RuntimeError("voge_b200.%s failed: CUDA error %d (%s)" % (what, code, msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream_of(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def require_cuda(*tensors):
    """The reference raises RuntimeError for CPU tensors (CUDAGuard); so do we."""
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("voge_b200: expected a CUDA tensor (there is no CPU path), got device %s" % t.device)


def f32c(t):
    require_cuda(t)
    if t.dtype != torch.float32:
        raise RuntimeError("voge_b200: expected float32, got %s" % t.dtype)
    return t.contiguous()


def i32c(t):
    require_cuda(t)
    if t.dtype != torch.int32:
        raise RuntimeError("voge_b200: expected int32, got %s" % t.dtype)
    return t.contiguous()

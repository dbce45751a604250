"""Development aid: bench.py under a different cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes).
usage: python tools/l2_fetch_ab.py <bytes> [bench args]"""
import ctypes, os, runpy, sys
import torch
torch.cuda.init()
torch.zeros(1, device="cuda")
rt = ctypes.CDLL("libcudart.so.12")
g = int(sys.argv[1])
rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(g))          # cudaLimitMaxL2FetchGranularity
val = ctypes.c_size_t(0)
rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
print("cudaDeviceSetLimit rc", rc, "granularity now", val.value, file=sys.stderr)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = [os.path.join(root, "bench.py")] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")

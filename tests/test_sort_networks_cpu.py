"""CPU check of the register sorting networks and the index folding used by the CUDA kernels
(voge_b200/csrc/sort_net.h is plain C++17: the same header is compiled here with g++)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sort_networks_and_fold_index(tmp_path):
    exe = str(tmp_path / "sort_net_check")
    src = os.path.join(ROOT, "tests", "csrc", "sort_net_check.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "voge_b200", "csrc"), src, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout + out.stderr

"""Import the reference's pure-Python modules from /root/reference WITHOUT pytorch3d (its
VoGE/__init__.py imports pytorch3d, so the package init is bypassed with a stub package).
Only usable where /root/reference exists (this container); used to generate tests/golden/*."""
import importlib.util
import os
import sys
import types

REF = os.environ.get("VOGE_REFERENCE_ROOT", "/root/reference")


def load(ref_C=None):
    pkg = types.ModuleType("VoGE")
    pkg.__path__ = [os.path.join(REF, "VoGE")]
    if ref_C is not None:
        pkg._C = ref_C
        sys.modules["VoGE._C"] = ref_C
    sys.modules["VoGE"] = pkg
    mods = {}
    names = ["Utils", "Aggregation"] + (["RayTracing", "Sampler", "Meshes"] if ref_C is not None else ["Meshes"])
    for name in names:
        spec = importlib.util.spec_from_file_location("VoGE." + name, os.path.join(REF, "VoGE", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["VoGE." + name] = m
        spec.loader.exec_module(m)
        setattr(pkg, name, m)
        mods[name] = m
    return mods

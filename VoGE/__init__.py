"""Drop-in alias: `import VoGE` resolves to the B200 implementation (voge_b200) with the reference's
module layout -- VoGE.Renderer, VoGE.RayTracing, VoGE.Aggregation, VoGE.Sampler, VoGE.Meshes,
VoGE.Utils and the native module VoGE._C (reference VoGE/__init__.py:7, csrc/ext.cpp:7-17)."""
import sys

import voge_b200 as _impl
from voge_b200 import _C, Aggregation, Converter, Meshes, RayTracing, Renderer, Sampler, Utils, cameras  # noqa: F401

__version__ = "0.4.1+b200." + _impl.__version__

for _name in ("_C", "Aggregation", "Converter", "Meshes", "RayTracing", "Renderer", "Sampler", "Utils", "cameras"):
    sys.modules[__name__ + "." + _name] = getattr(_impl, _name)
for _name in ("IO", "Converters", "Cuboid"):
    sys.modules[__name__ + ".Converter." + _name] = getattr(_impl.Converter, _name)

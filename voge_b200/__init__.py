"""voge_b200 -- B200-native (sm_100a) implementation of the VoGE ray-tracing hot path behind VoGE's
own Python API.  Module names mirror the reference package (Renderer, RayTracing, Aggregation,
Sampler, Meshes, Utils); `_C` mirrors the reference's pybind extension `VoGE._C`."""
__version__ = "0.1.0"

from . import _C, Aggregation, Meshes, RayTracing, Renderer, Sampler, Utils, cameras  # noqa: F401
from .Meshes import GaussianMeshes, GaussianMeshesNaive  # noqa: F401
from .Renderer import (Fragments, GaussianRenderer, GaussianRenderSettings, get_silhouette,  # noqa: F401
                       interpolate_attr, to_colored_background, to_white_background)
from .Sampler import sample_features, scatter_max_weight  # noqa: F401

"""Mesh / point-cloud -> Gaussian converters and OFF / GOFF I/O (reference VoGE/Converter/):
pre-processing around the hot path, same function names and return conventions as the reference."""
from . import IO, Converters, Cuboid  # noqa: F401

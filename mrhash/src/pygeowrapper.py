"""Drop-in for the reference's nanobind module ``mrhash.src.pygeowrapper``
(/root/reference/mrhash/src/sdf/pybind/pygeowrapper.cpp:12-84)."""
from mrhash_b200.geowrapper import GeoWrapper  # noqa: F401

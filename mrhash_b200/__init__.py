"""mrhash_b200 — B200-native implementation of mrhash's per-frame TSDF integration hot path
behind the reference's ``pygeowrapper.GeoWrapper`` API."""
from .geowrapper import VOXEL_DTYPE, GeoWrapper  # noqa: F401

__all__ = ["GeoWrapper", "VOXEL_DTYPE"]

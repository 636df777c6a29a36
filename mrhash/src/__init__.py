"""Import shim: lets the reference's runners (``from mrhash.src.pygeowrapper import GeoWrapper``,
apps/rgbd_runner.py:9) pick up mrhash_b200's GeoWrapper unchanged."""

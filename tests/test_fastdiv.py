"""GPU: the shared-reciprocal divisions of the frame kernel (csrc/mrh_div.cuh) are bit-identical to
IEEE division. They stand in for the `a / b` of voxel_hash_utils.cuh:75-151,169-181, camera.cuh:131-160
and voxel_data_structures.cu:803-822, where one different last bit can move a ray into another block or
a voxel onto another pixel - so the check is exhaustive over the numerator for the divisors a map
really uses, plus random pairs over the whole exponent window."""
import ctypes as C

import numpy as np
import pytest

from mrhash_b200 import GeoWrapper, _capi, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize(
    "divisor",
    [
        np.float32(0.01),  # replica.cfg virtual_voxel_size
        np.float32(0.01) * np.float32(0.5),  # half a voxel: combineVoxel's delta
        np.float32(0.2),  # vbr.cfg
        np.float32(0.002),
        np.float32(1.0),  # weight sums
        np.float32(3.0),
        np.float32(255.0),
        np.float32(0.57735026),  # a direction component
        np.float32(1.9999999),
        np.float32(1.7320508e-6),
    ],
)
def test_every_numerator_of_a_divisor(divisor):
    bad = C.c_uint64(123)
    _capi.check(_capi.lib().mrh_selftest_div(0, C.c_float(float(divisor)), 0, 0, C.byref(bad)))
    assert bad.value == 0


def test_random_pairs():
    bad = C.c_uint64(123)
    _capi.check(_capi.lib().mrh_selftest_div(0, C.c_float(0.0), 1 << 33, 12345, C.byref(bad)))
    assert bad.value == 0


def test_block_shortcut_is_verified_for_the_shipped_voxel_sizes():
    for size in (0.01, 0.2, 0.004):
        params = dict(synth.REPLICA_PARAMS)
        params["virtual_voxel_size"] = size
        g = GeoWrapper(**params, num_sdf_blocks=2000, hash_num_buckets=1000, max_num_triangles=1, device=0)
        fx, fy, cx, cy = synth.intrinsics(64, 48)
        g.setCamera(fx, fy, cx, cy, 48, 64, params["min_depth"], params["max_depth"], 0)
        t, q, depth, rgb = synth.rgbd_frame(0, n_frames=10, width=64, height=48)
        g.setCurrPose(t, q)
        g.setDepthImage(depth)
        g.setRGBImage(rgb)
        g.compute()
        r = C.c_int(-1)
        _capi.check(_capi.lib().mrh_get_block_shortcut_radius(g._h, C.byref(r)))
        # a room of a few metres is far inside the verified radius (voxels)
        assert r.value * size > 50.0, (size, r.value)

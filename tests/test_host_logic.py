"""CPU: host-side pieces that need no device - synthetic streams, drop-in import path, bench plumbing."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

from mrhash_b200 import synth


def test_dropin_import_path():
    # apps/rgbd_runner.py:9 does `from mrhash.src.pygeowrapper import GeoWrapper`
    from mrhash.src.pygeowrapper import GeoWrapper as A
    from mrhash_b200 import GeoWrapper as B

    assert A is B
    import inspect

    sig = inspect.signature(A.__init__)
    # pygeowrapper.cpp:14-29, in order
    want = ["sdf_truncation", "sdf_truncation_scale", "integration_weight_sample", "virtual_voxel_size", "n_frames_invalidate_voxels", "voxel_extents_scale", "viewer_active", "marching_cubes_threshold", "min_weight_threshold", "min_depth", "max_depth", "gs_optimization_param_path", "sdf_var_threshold", "vertices_merging_threshold", "projective_sdf"]
    assert list(sig.parameters)[1 : 1 + len(want)] == want
    for m in ["getHashNumBuckets", "getNumSdfBlocks", "getCurrPose", "getVertices", "getFaces", "getColors", "setRGBImage", "setDepthImage", "setPointCloud", "setCamera", "setCurrPose", "setCameraInLidar", "compute", "extractMesh", "streamAllOut", "clearBuffers", "serializeData", "serializeGrid", "deserializeGrid", "GSSavePointCloud", "GSFinalOpt"]:
        assert callable(getattr(A, m)), m


def test_synthetic_rgbd_stream_is_seeded_and_plausible():
    t, q, d, rgb = synth.rgbd_frame(0, orbit=False)
    t2, q2, d2, rgb2 = synth.rgbd_frame(0, orbit=False)
    assert np.array_equal(d, d2) and np.array_equal(rgb, rgb2)
    assert d.dtype == np.float32 and rgb.dtype == np.uint8 and d.shape == (480, 640) and rgb.shape == (480, 640, 3)
    assert 0.7 < d.min() < 0.9 and d.max() == 2.0  # sphere in front, wall at z = 2
    assert np.allclose(d * synth.DEPTH_SCALE, np.round(d * synth.DEPTH_SCALE), atol=1e-2)  # Replica PNG quantisation
    # orbit: unit-radius circle, optical axis radially outward
    for k in (0, 250, 500):
        t, q, R = synth.orbit_pose(k, 1000)
        assert abs(np.linalg.norm(t) - 1) < 1e-6 and np.allclose(R[:, 2] * 1.0, t, atol=1e-6)
        assert np.allclose(synth.quat_to_matrix_f32(t, q)[:3, :3], R, atol=1e-6)
    # the torch renderer used by bench.py produces the same frame
    d3, c3 = synth.render_rgbd_torch(np.eye(3), np.zeros(3), device="cpu")
    assert np.array_equal(d3.numpy(), d2) and np.array_equal(c3.numpy(), rgb2)


def test_synthetic_lidar_stream():
    T, pts = synth.lidar_frame(0)
    assert pts.dtype == np.float32 and pts.shape[1] == 3 and 100000 < len(pts) <= 128 * 1024
    r = np.linalg.norm(pts, axis=1)
    assert r.min() > 0.2 and r.max() < 100.0
    assert np.array_equal(T[:3, :3], np.eye(3, dtype=np.float32))


def test_bench_reference_arm_reports_unavailable_without_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    import torch

    if not torch.cuda.is_available():
        assert "unavailable" in line


def test_reference_config_surface(tmp_path):
    """configurations/*.cfg of the reference (YAML) -> GeoWrapper keywords (mrhash_b200/config.py)."""
    import inspect

    from mrhash_b200 import GeoWrapper, synth
    from mrhash_b200.config import load_config

    cfg = tmp_path / "replica.cfg"
    cfg.write_text(
        "map:\n    sdf_truncation            : 0.07\n    sdf_truncation_scale      : 0.00\n    integration_weight_sample : 1 \n"
        "    virtual_voxel_size        : 0.01\n    n_frames_invalidate_voxels: 100\n\nstreamer:\n    voxel_extents_scale       : 1\n\n"
        "mesh:\n    marching_cubes_threshold: 1.5\n    min_weight_threshold : 5\n    sdf_var_threshold : 0.0\n    vertices_merging_threshold : 0.0\n\n"
        "sensor:\n    min_depth : 0.01\n    max_depth : 30\n    intrinsics: [600.0, 600.0, 599.5, 339.5]\n    resolution: [1200, 680]\n"
        "    depth_scaling: 6553.5\n    hz: 30\n\ndata_path: /data\nresults_path: /out\ngs_optimization_param_path: ./params.json\nend_frame: -1\n"
    )
    kw, sensor = load_config(str(cfg))
    # the same values as the hard-coded replica parameters the tests use, and only constructor keywords
    for k, v in synth.REPLICA_PARAMS.items():
        assert kw[k] == pytest.approx(v), k
    accepted = set(inspect.signature(GeoWrapper.__init__).parameters)
    assert set(kw) <= accepted
    assert sensor["resolution"] == [1200, 680] and sensor["end_frame"] == -1 and sensor["depth_scaling"] == 6553.5
    lidar = {"map": {"sdf_truncation": 0.4, "sdf_truncation_scale": 0.0, "integration_weight_sample": 1, "virtual_voxel_size": 0.2, "n_frames_invalidate_voxels": 0},
             "streamer": {"voxel_extents_scale": 1}, "mesh": {"marching_cubes_threshold": 1.5, "sdf_var_threshold": 0.0, "vertices_merging_threshold": 0.0, "min_weight_threshold": 50},
             "sensor": {"min_depth": 0.2, "max_depth": 100, "rosbag_topic": "/ouster/points"}, "end_frame": -1}
    kw, sensor = load_config(lidar)
    for k, v in synth.VBR_PARAMS.items():
        assert kw[k] == pytest.approx(v), k
    with pytest.raises(KeyError, match="mesh"):
        load_config({"map": lidar["map"], "streamer": lidar["streamer"], "sensor": lidar["sensor"]})

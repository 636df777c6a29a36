"""GPU: the streaming options of the host side change WHEN work is submitted, never what is computed.

mrh_set_ingest_mode (0: the setter copies into pinned staging, as the reference's setDepthImage
does; 1: DMA from the caller's page-locked frame, awaited in compute(); 2: fully asynchronous) and
mrh_set_stats_pipeline / mrh_get_stats_pipelined (counters copied behind every frame, read 0..3
frames late) against the default synchronous sequence, on the same frames."""
import numpy as np
import pytest

from compare import compare_dumps

from mrhash_b200 import GeoWrapper, synth

pytestmark = pytest.mark.gpu

N_FRAMES = 12
KEYS = ("frames", "rays_valid", "blocks_new", "blocks_visible", "voxels_updated", "blocks_freed", "live_blocks", "heap_free", "dropped_heap", "dropped_table")


def _new():
    params = dict(synth.REPLICA_PARAMS)
    params["n_frames_invalidate_voxels"] = 5  # starve frames inside the window
    fx, fy, cx, cy = synth.intrinsics(640, 480)
    g = GeoWrapper(**params, num_sdf_blocks=60000, hash_num_buckets=30000, max_num_triangles=1)
    g.setCamera(fx, fy, cx, cy, 480, 640, params["min_depth"], params["max_depth"], 0)
    return g


def _frames():
    return [synth.rgbd_frame(k, n_frames=200) for k in range(N_FRAMES)]


def _reference_run(frames):
    g = _new()
    per_frame = []
    for t, q, depth, rgb in frames:
        g.setCurrPose(t, q)
        g.setDepthImage(depth)
        g.setRGBImage(rgb)
        g.compute()
        per_frame.append(g.getStats())
    return g.dumpState(), per_frame


@pytest.mark.parametrize("mode", [1, 2])
def test_ingest_modes_compute_the_same_map(mode):
    import torch

    frames = _frames()
    want_dump, want_stats = _reference_run(frames)
    # page-locked copies of every frame, alive for the whole run (mode 2's contract)
    pinned = [(torch.from_numpy(d.copy()).pin_memory(), torch.from_numpy(c.copy()).pin_memory()) for _, _, d, c in frames]
    g = _new()
    g.setIngestMode(mode)
    for (t, q, _, _), (d, c) in zip(frames, pinned):
        g.setCurrPose(t, q)
        g.setDepthImage(d.numpy())
        g.setRGBImage(c.numpy())
        g.compute()
    got = g.getStats()
    assert {k: got[k] for k in KEYS} == {k: want_stats[-1][k] for k in KEYS}
    assert compare_dumps(g.dumpState(), want_dump)["ok"]


@pytest.mark.parametrize("lag", [0, 1, 2, 3])
def test_pipelined_counters_are_the_frames_own(lag):
    frames = _frames()
    _, want_stats = _reference_run(frames)
    g = _new()
    g.setStatsPipeline(True)
    seen = []
    for i, (t, q, depth, rgb) in enumerate(frames):
        g.setCurrPose(t, q)
        g.setDepthImage(depth)
        g.setRGBImage(rgb)
        g.compute()
        if i >= lag:
            seen.append(g.getStatsPipelined(lag))
    for back in range(lag - 1, -1, -1):  # drain: the last `lag` frames
        seen.append(g.getStatsPipelined(back))
    assert len(seen) == len(frames)
    for i, (a, b) in enumerate(zip(seen, want_stats)):
        assert {k: a[k] for k in KEYS} == {k: b[k] for k in KEYS}, i


def test_pipelined_counters_refuse_frames_that_do_not_exist():
    g = _new()
    g.setStatsPipeline(True)
    with pytest.raises(Exception):
        g.getStatsPipelined(0)
    t, q, depth, rgb = synth.rgbd_frame(0, n_frames=200)
    g.setCurrPose(t, q)
    g.setDepthImage(depth)
    g.setRGBImage(rgb)
    g.compute()
    assert g.getStatsPipelined(0)["frames"] == 1
    with pytest.raises(Exception):
        g.getStatsPipelined(1)
    with pytest.raises(Exception):
        g.getStatsPipelined(4)

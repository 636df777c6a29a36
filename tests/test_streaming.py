"""GPU: radius paging between the device map and the host store - Streamer::stream
(/root/reference/mrhash/src/sdf/streamer.cpp:337-355, streamer.cu:11-59) and its trigger in
GeoWrapper::compute (geowrapper.cpp:137-138, stream_threshold params.h:28). The reference's streamer
cannot be built here (Eigen / cista), so the checks are the invariants its own test holds
(tests/test_streamer.cu:40-117: nothing is lost, few duplicates) plus exact conservation of every
block that was never re-allocated while paged out."""
import numpy as np
import pytest

from compare import compare_dumps

from mrhash_b200 import GeoWrapper, synth

pytestmark = pytest.mark.gpu

W, H = 320, 240
N_FRAMES = 72


def make(num_blocks, threshold):
    p = dict(synth.REPLICA_PARAMS)
    p["max_depth"] = 2.0  # paging radius = camera max depth (geowrapper.cpp:138)
    p["n_frames_invalidate_voxels"] = 0  # no garbage collection: every allocated block must survive somewhere
    g = GeoWrapper(**p, num_sdf_blocks=num_blocks, hash_num_buckets=max(1000, num_blocks // 2), max_num_triangles=1)
    fx, fy, cx, cy = synth.intrinsics(W, H)
    g.setCamera(fx, fy, cx, cy, H, W, p["min_depth"], p["max_depth"], 0)
    g._set("StreamThreshold", threshold)
    return g


def feed(g, k):
    t, q, depth, rgb = synth.rgbd_frame(k, n_frames=N_FRAMES, width=W, height=H)
    g.setCurrPose(t, q)
    g.setDepthImage(depth)
    g.setRGBImage(rgb)
    g.compute()
    g.synchronize()
    return t


def keyed(entries, voxels):
    return {tuple(e[:3]): v for e, v in zip(entries.tolist(), voxels)}


def test_compute_pages_far_blocks_out_when_the_pool_runs_low():
    whole = make(60000, 0.0)  # paging off, pool large enough for the whole orbit
    for k in range(N_FRAMES):
        feed(whole, k)
    ew, vw = whole.dumpState()
    total = len(ew)
    assert total > 3000 and whole.getStats()["dropped_heap"] == 0

    small = make(int(total * 0.55), 0.15)
    for k in range(N_FRAMES):
        feed(small, k)
    st = small.getStats()
    events, dup = int(small._get("StreamEvents")), int(small._get("StreamDuplicates"))
    print(f"[paging] map {total} blocks, pool {int(total * 0.55)}: {events} stream events, store {small.storeSize()} blocks, device {st['live_blocks']}, duplicates {dup}")
    assert events >= 1 and small.storeSize() > 0
    assert st["dropped_heap"] == 0 and st["dropped_table"] == 0  # without paging the pool would have overflowed
    assert st["live_blocks"] + small.storeSize() >= total

    # the same stream without paging overflows the small pool
    tight = make(int(total * 0.55), 0.0)
    for k in range(N_FRAMES):
        feed(tight, k)
    assert tight.getStats()["dropped_heap"] > 0

    # conservation: everything ends up in the store; as a key set it is the unpaged map, and every
    # block that exists once (never re-allocated while it was paged out) is bit-identical to it
    small.streamAllOut()
    assert small.getStats()["live_blocks"] == 0
    es, vs = small.storeDump()
    n_store = len(es)
    assert {tuple(e[:3]) for e in es.tolist()} == {tuple(e[:3]) for e in ew.tolist()}
    n_dup = n_store - len(np.unique(es[:, :3], axis=0))
    # Blocks between max_depth and max_depth / cos(half FOV) from the camera are paged out and allocated
    # again by the next frame (the paging radius IS max_depth, geowrapper.cpp:138): copies of one key
    # accumulate in the store. Streaming them in fuses the copies (running weighted mean).
    big = make(60000, 0.0)
    big.storeAppend(es, vs)
    big.stream(np.zeros(3), 1e6)
    assert big.storeSize() == 0 and int(big._get("StreamDuplicates")) == n_dup
    em, vm = big.dumpState()
    assert np.array_equal(em[:, :3], ew[:, :3])  # one block per key again, the unpaged map's key set
    # fused weights never exceed the unpaged map's (a paged-out block can only miss updates), and the
    # voxels that saw every update agree with it
    wm, ww = vm["weight"].astype(int), vw["weight"].astype(int)
    assert (wm <= ww).all()
    full = (wm == ww) & (ww > 0)
    close = np.isclose(vm["sdf"][full], vw["sdf"][full], rtol=1e-4, atol=1e-6)
    print(f"[paging] {n_store} stored records, {n_dup} repeated keys fused; voxels with the full weight {int(full.sum())}/{int((ww > 0).sum())}, sdf agreeing on them {float(close.mean()):.4f}")
    assert full.sum() > 0.6 * (ww > 0).sum() and close.mean() > 0.999


def test_explicit_stream_round_trip():
    g = make(40000, 0.0)
    for k in range(6):
        t = feed(g, k)
    before = g.dumpState()
    n = len(before[0])
    free0 = g.getStats()["heap_free"]
    g.stream(t, 0.0)  # radius 0: every block is "far" and leaves; nothing lies inside the empty sphere
    assert g.getStats()["live_blocks"] == 0 and g.storeSize() == n and g.getStats()["heap_free"] == free0 + n
    assert int(g._get("LastStreamOutBlocks")) == n and int(g._get("LastStreamInBlocks")) == 0
    g.stream(t, 1e6)  # everything is near: all of it comes back
    assert g.storeSize() == 0 and int(g._get("LastStreamInBlocks")) == n
    after = g.dumpState()
    rep = compare_dumps(after, before)
    assert rep["ok"] and rep["sdf_bitexact"] and rep["sum_squared_bitexact"], rep
    # integration continues on the paged-in map exactly as on an untouched one
    h = make(40000, 0.0)
    for k in range(7):
        feed(h, k)
    feed(g, 6)
    rep = compare_dumps(g.dumpState(), h.dumpState())
    assert rep["ok"] and rep["sdf_bitexact"], rep


def test_serialize_grid_checkpoint_round_trip(tmp_path):
    """serializeGrid / deserializeGrid (geowrapper.cpp:567-573) through the handle: a map leaves the
    device, goes to a checkpoint in the reference's file format (pinned byte for byte against cista in
    tests/test_grid_format.py), comes back into a fresh handle and continues exactly as the original."""
    from mrhash_b200.geowrapper import grid_read

    a = make(40000, 0.0)
    for k in range(6):
        t = feed(a, k)
    before = a.dumpState()
    path = str(tmp_path / "grid.bin")
    a.streamAllOut()
    a.serializeGrid(path)
    e, v = grid_read(path)
    assert len(e) == len(before[0]) and {tuple(x[:3]) for x in e.tolist()} == {tuple(x[:3]) for x in before[0].tolist()}
    b = make(40000, 0.0)
    b.deserializeGrid(path)
    assert b.storeSize() == len(before[0])
    b.deserializeGrid(path)  # chunks of the file replace the same chunks: reading twice does not double anything
    assert b.storeSize() == len(before[0])
    b.stream(t, 1e6)  # page everything in
    rep = compare_dumps(b.dumpState(), before)
    assert rep["ok"] and rep["sdf_bitexact"] and rep["sum_squared_bitexact"], rep
    # the restored map integrates the next frame exactly like a map that never left the device
    c = make(40000, 0.0)
    for k in range(7):
        feed(c, k)
    b._lib.mrh_set_field(b._h, b"NFramesInvalidateVoxels", 0.0)
    feed(b, 6)
    rep = compare_dumps(b.dumpState(), c.dumpState())
    assert rep["ok"] and rep["sdf_bitexact"], rep


def test_paging_next_to_the_reference_streamer():
    """f-1 against the reference's OWN Streamer (streamer.cpp / streamer.cu compiled unmodified into
    oracle/_ref): the same 72-frame orbit in the same undersized pool, the reference paged the way
    GeoWrapper::compute pages it (geowrapper.cpp:137-138: stream(camera position, max depth) in front of
    integrate() whenever the free pool is at or below stream_threshold).
    What must agree: nothing is lost on either side - after streamAllOut both stores hold the key set of
    the unpaged map. What differs by design, stated as numbers: the reference keeps a re-allocated key
    twice (its own test bounds the ratio, tests/test_streamer.cu:40-117), mrhash_b200 fuses the copies
    when they are paged in together."""
    from oracle_lib import RefCuda, ref_available

    if not ref_available():
        pytest.skip("oracle/_ref not built")
    whole = make(60000, 0.0)
    for k in range(N_FRAMES):
        feed(whole, k)
    ew, _ = whole.dumpState()
    total = len(ew)
    pool = int(total * 0.55)
    small = make(pool, 0.15)
    p = dict(synth.REPLICA_PARAMS)
    p["max_depth"] = 2.0
    p["n_frames_invalidate_voxels"] = 0
    ref = RefCuda(p, pool, max(1000, pool // 2))
    fx, fy, cx, cy = synth.intrinsics(W, H)
    ref.set_camera(fx, fy, cx, cy, H, W, p["min_depth"], p["max_depth"], 0)
    ref.streamer_create(pool)
    ref_events = 0
    for k in range(N_FRAMES):
        t = feed(small, k)
        _, _, depth, rgb = synth.rgbd_frame(k, n_frames=N_FRAMES, width=W, height=H)
        ref_events += ref.stream(t, p["max_depth"]) == 1
        ref.compute_rgbd(small.getCurrPose(), depth, rgb)
    ref_dup_pct = ref.duplicates_ratio()
    small.streamAllOut()
    ref.stream_all_out()
    es, _ = small.storeDump()
    ours_keys = {tuple(e[:3]) for e in es.tolist()}
    whole_keys = {tuple(e[:3]) for e in ew.tolist()}
    n_ref = ref.grid_blocks()
    ours_dup = len(es) - len(ours_keys)
    print(
        f"[paging vs reference] map {total} blocks, pool {pool}: ours {int(small._get('StreamEvents'))} stream events, {len(es)} stored records "
        f"({ours_dup} repeated keys = {100.0 * ours_dup / len(es):.2f} %); reference {ref_events} stream events, {n_ref} stored records "
        f"({n_ref - total} above the unpaged map's {total} = {100.0 * (n_ref - total) / max(n_ref, 1):.2f} %, its own duplicate check said {ref_dup_pct:.2f} % before the final stream-out)"
    )
    assert ours_keys == whole_keys and small.getStats()["dropped_heap"] == 0
    assert ref_events >= 1 and n_ref >= total  # the reference lost nothing either (it may hold keys twice)

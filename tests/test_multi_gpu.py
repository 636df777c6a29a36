"""GPU: hash-bucket-range sharding. Two shard handles (on one device, or on two when available)
fed the same frames must hold, together, exactly the unsharded map - bit for bit - and each block
must sit on the rank `sharding.owner_of` names."""
import numpy as np
import pytest

from compare import compare_dumps
from test_parity_rgbd import NUM_BLOCKS, NUM_BUCKETS

from mrhash_b200 import GeoWrapper, sharding, synth

pytestmark = pytest.mark.gpu


def make(rank, world, device=0, width=320, height=240):
    p = dict(synth.REPLICA_PARAMS)
    g = GeoWrapper(**p, num_sdf_blocks=NUM_BLOCKS, hash_num_buckets=NUM_BUCKETS, max_num_triangles=1, device=device, shard_rank=rank, shard_world=world)
    fx, fy, cx, cy = synth.intrinsics(width, height)
    g.setCamera(fx, fy, cx, cy, height, width, p["min_depth"], p["max_depth"], 0)
    return g


@pytest.mark.parametrize("world", [2, 3])
def test_union_of_shards_equals_unsharded_map(world):
    import torch

    ndev = torch.cuda.device_count()
    whole = make(0, 1)
    shards = [make(r, world, device=r % ndev) for r in range(world)]
    for k in range(6):
        t, q, depth, rgb = synth.rgbd_frame(k, n_frames=100, width=320, height=240)
        for g in [whole] + shards:
            g.setCurrPose(t, q)
            g.setDepthImage(depth)
            g.setRGBImage(rgb)
            g.compute()
    ref = whole.dumpState()
    parts = [g.dumpState() for g in shards]
    for r, (e, _) in enumerate(parts):
        assert len(e) > 0
        assert (sharding.owner_of(e[:, :3], world, NUM_BUCKETS) == r).all()
    ee = np.concatenate([e for e, _ in parts])
    vv = np.concatenate([v for _, v in parts])
    order = np.lexsort((ee[:, 2], ee[:, 1], ee[:, 0]))
    rep = compare_dumps((ee[order], vv[order]), ref)
    assert rep["ok"] and rep["sdf_bitexact"] and rep["sum_squared_bitexact"], rep
    assert sum(g.getStats()["voxels_updated"] for g in shards) == whole.getStats()["voxels_updated"]


def test_sharded_map_meshes_like_the_unsharded_one():
    """Shard 1's blocks are appended to shard 0's host store (what sharding.extract_mesh_sharded does
    after the NCCL gather); meshing the union must give the unsharded map's triangle soup exactly."""
    import torch

    from mrhash_b200 import GeoWrapper

    ndev = torch.cuda.device_count()
    p = dict(synth.REPLICA_PARAMS)
    w, h = 320, 240
    fx, fy, cx, cy = synth.intrinsics(w, h)

    def mk(rank, world, device=0):
        g = GeoWrapper(**p, num_sdf_blocks=NUM_BLOCKS, hash_num_buckets=NUM_BUCKETS, max_num_triangles=1_000_000, device=device, shard_rank=rank, shard_world=world)
        g.setCamera(fx, fy, cx, cy, h, w, p["min_depth"], p["max_depth"], 0)
        return g

    whole, a, b = mk(0, 1), mk(0, 2), mk(1, 2, device=1 % ndev)
    for k in range(6):
        t, q, depth, rgb = synth.rgbd_frame(k, n_frames=2000, width=w, height=h)
        for g in (whole, a, b):
            g.setCurrPose(t, q)
            g.setDepthImage(depth)
            g.setRGBImage(rgb)
            g.compute()
    whole.extractMesh(None)
    eb, vb = b.dumpState()
    a.streamAllOut()
    a.storeAppend(eb, vb)
    a.setShard(0, 1)
    a.extractMesh(None)
    a.setShard(0, 2)
    ta, tw = a.getTriangles(), whole.getTriangles()
    assert len(ta) == len(tw) > 1000
    ka = np.sort(np.ascontiguousarray(ta.reshape(len(ta), -1)).view([("", ta.dtype)] * 18).ravel())
    kw = np.sort(np.ascontiguousarray(tw.reshape(len(tw), -1)).view([("", tw.dtype)] * 18).ravel())
    assert (ka == kw).all()

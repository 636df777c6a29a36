"""In-process reproduction of the sharded-mesh exchange after a long stream (world shards on one GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mrhash_b200 import GeoWrapper, sharding, synth

world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 330
NB, NK = 500000, 250000
p = dict(synth.REPLICA_PARAMS)
w, h = 640, 480
fx, fy, cx, cy = synth.intrinsics(w, h)
def mk(r, n):
    g = GeoWrapper(**p, num_sdf_blocks=NB, hash_num_buckets=NK, max_num_triangles=4_000_000, shard_rank=r, shard_world=n)
    g.setCamera(fx, fy, cx, cy, h, w, p["min_depth"], p["max_depth"], 0)
    return g
shards = [mk(r, world) for r in range(world)]
for k in range(n_frames):
    t, q, R = synth.orbit_pose(k, 1000)
    d, c = synth.render_rgbd_torch(R, t, w, h, device="cuda")
    torch.cuda.synchronize()
    for g in shards:
        g.setCurrPose(t, q); g.setDepthImageDevice(d.data_ptr(), h, w); g.setRGBImageDevice(c.data_ptr(), h, w); g.compute(); g.synchronize()
for r, g in enumerate(shards):
    st0 = g.getStats()
    e, _ = g.dumpState()
    req = g.haloRequests()
    owner = sharding.owner_of_torch(req, world, NK)
    uniq = len(torch.unique(req, dim=0))
    own_set = {tuple(x) for x in e[:, :3].tolist()}
    clash = sum(tuple(x) in own_set for x in req.cpu().tolist())
    keys, recs = [], []
    for o in range(world):
        kk = req[owner == o].contiguous()
        keys.append(kk); recs.append(shards[o].haloPack(kk, False))
    keys, recs = torch.cat(keys), torch.cat(recs)
    hdr = recs[:, :4].contiguous().view(torch.int32)[:, 0]
    print(f"rank {r}: owned {len(e)} live_blocks {st0['live_blocks']} requests {len(req)} unique {uniq} self-owned {int((owner == r).sum())} already-present {clash} found {int((hdr >= 0).sum())}")
    try:
        g.haloInsert(keys, recs, False)
    except RuntimeError as ex:
        print("  insert failed:", ex)
    st1 = g.getStats()
    print("  dropped_heap", st1["dropped_heap"] - st0["dropped_heap"], "dropped_table", st1["dropped_table"] - st0["dropped_table"], "blocks_new", st1["blocks_new"] - st0["blocks_new"])
    g.haloClear()

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
whole = mk(0, 1)
for k in range(n_frames):
    t, q, R = synth.orbit_pose(k, 1000)
    d, c = synth.render_rgbd_torch(R, t, w, h, device="cuda")
    torch.cuda.synchronize()
    for g in shards + [whole]:
        g.setCurrPose(t, q); g.setDepthImageDevice(d.data_ptr(), h, w); g.setRGBImageDevice(c.data_ptr(), h, w)
    whole.compute(); whole.synchronize()
    need = [g.computeBegin() for g in shards]
    if need[0]:  # starve frame: what sharding.compute_sharded does with an NCCL all-reduce(MIN)
        for g in shards:
            g.synchronize()
        zs = [g.zbufTensor() for g in shards]
        zmin = zs[0].clone()
        for z in zs[1:]:
            zmin = torch.minimum(zmin, z)
        for z in zs:
            z.copy_(zmin)
        torch.cuda.synchronize()
    for g in shards:
        g.computeEnd(); g.synchronize()
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
    soups = globals().setdefault("soups", [])
    soups.append(g.meshLocal().cpu().numpy())
    dumps = globals().setdefault("dumps", [])
    dumps.append(g.dumpState())  # owned + ghost blocks
    g.haloClear()
def srt(t):
    t = np.ascontiguousarray(t.reshape(len(t), -1))
    return np.sort(t.view([("", t.dtype)] * 18).ravel())
union = np.concatenate(soups)
ws = whole.meshLocal().cpu().numpy()
print("union", len(union), "unsharded", len(ws), "identical", len(union) == len(ws) and bool((srt(union) == srt(ws)).all()))

def rows(t):
    t = np.ascontiguousarray(t.reshape(len(t), -1))
    return t.view([("", t.dtype)] * 18).ravel()
ru, rw = rows(union), rows(ws)
missing = ws[~np.isin(rw, ru)]
extra = union[~np.isin(ru, rw)]
print("missing", len(missing), "extra", len(extra))
we, wv = whole.dumpState()
wkeys = {tuple(k): i for i, k in enumerate(we[:, :3].tolist())}
size = 0.01
for tri in missing[:6]:
    c = tri[:, :3].mean(axis=0)
    vox = np.round(c / size).astype(int)
    blk = np.floor_divide(vox, 8)
    own = int(sharding.owner_of(blk[None], world, NK)[0])
    print(" tri centre", c, "voxel", vox, "block", blk, "owner", own, "block in whole map:", tuple(blk) in wkeys)
    e, v = dumps[own]
    have = {tuple(k): i for i, k in enumerate(e[:, :3].tolist())}
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                nbk = (blk[0] + dx, blk[1] + dy, blk[2] + dz)
                if nbk in wkeys:
                    o2 = int(sharding.owner_of(np.array([nbk]), world, NK)[0])
                    if nbk not in have:
                        print("   neighbour", nbk, "owner", o2, "exists in the whole map but NOT on the meshing rank")
                    else:
                        a, b = v[have[nbk]], wv[wkeys[nbk]]
                        idx = np.arange(512); x, y, z = idx & 7, (idx >> 3) & 7, idx >> 6
                        shell = (x == 0) | (x == 7) | (y == 0) | (y == 7) | (z == 0) | (z == 7)
                        m = shell if o2 != own else np.ones(512, bool)
                        same = (a["sdf"][m].view(np.uint32) == b["sdf"][m].view(np.uint32)).all() and (a["weight"][m] == b["weight"][m]).all()
                        if not same:
                            print("   neighbour", nbk, "owner", o2, "payload differs")

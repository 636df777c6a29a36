"""CPU, world_size 2 (gloo): host-side logic of the multi-GPU partition (mrhash_b200/sharding.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mrhash_b200 import VOXEL_DTYPE, sharding, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. frame broadcast: rank 0 ingests, everybody ends up with the same bytes
        if rank == 0:
            t, q, depth, rgb = synth.rgbd_frame(3, n_frames=100, width=64, height=48)
            d, c, p = torch.from_numpy(depth), torch.from_numpy(rgb), torch.from_numpy(np.concatenate([t, q]).astype(np.float32))
        else:
            d, c, p = torch.zeros((48, 64)), torch.zeros((48, 64, 3), dtype=torch.uint8), torch.zeros(7)
        sharding.broadcast_frame(d, c, p, src=0)
        t, q, depth, rgb = synth.rgbd_frame(3, n_frames=100, width=64, height=48)
        assert np.array_equal(d.numpy(), depth) and np.array_equal(c.numpy(), rgb)
        assert np.array_equal(p.numpy(), np.concatenate([t, q]).astype(np.float32))
        # 1b. scatter ingest: every rank uploads only its band of rows, the in-place all-gather completes the frame
        dd, cc = torch.zeros((48, 64)), torch.zeros((48, 64, 3), dtype=torch.uint8)
        lo, hi = sharding.frame_row_band(rank, world, 48)
        host_d, host_c = torch.zeros((48, 64)), torch.zeros((48, 64, 3), dtype=torch.uint8)
        host_d[lo:hi], host_c[lo:hi] = torch.from_numpy(depth)[lo:hi], torch.from_numpy(rgb)[lo:hi]  # the rest is never read
        sharding.scatter_ingest_frame(dd, cc, host_d, host_c)
        assert np.array_equal(dd.numpy(), depth) and np.array_equal(cc.numpy(), rgb)
        # 2. each rank "owns" the blocks of its bucket range; gathering them on rank 0 restores the set
        rng = np.random.default_rng(0)
        blocks = np.unique(rng.integers(-40, 40, size=(500, 3)), axis=0).astype(np.int32)
        nb = 1000
        mine = blocks[sharding.owner_of(blocks, world, nb) == rank]
        lo, hi = sharding.bucket_range(rank, world, nb)
        h = sharding.block_hash(mine, nb)
        assert ((h >= lo) & (h < hi)).all()
        entries = np.concatenate([mine, np.zeros((len(mine), 2), np.int32)], axis=1)
        voxels = np.zeros((len(mine), 512), VOXEL_DTYPE)
        voxels["sdf"] = mine[:, :1].astype(np.float32)
        voxels["weight"] = rank + 1
        ee, vv = sharding.gather_blocks(entries, voxels, dst=0)
        if rank == 0:
            order = np.lexsort((blocks[:, 2], blocks[:, 1], blocks[:, 0]))
            assert np.array_equal(ee[:, :3], blocks[order])
            assert np.array_equal(vv["sdf"][:, 0], blocks[order][:, 0].astype(np.float32))
            owners = sharding.owner_of(ee[:, :3], world, nb)
            assert np.array_equal(vv["weight"][:, 0], owners + 1)
        else:
            assert ee is None and vv is None
        out.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover
        out.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


class _FakeShard:
    """numpy stand-in for a sharded GeoWrapper: owns the blocks of its bucket range out of a common
    random set; a block's record payload is derived from its key, so every answer can be checked."""

    REC = 32

    def __init__(self, blocks, rank, world, nb):
        self.rank, self.world, self.nb = rank, world, nb
        self.all = blocks
        self.own = {tuple(b) for b in blocks[sharding.owner_of(blocks, world, nb) == rank].tolist()}
        self.ghosts = {}

    def getHashNumBuckets(self):
        return self.nb

    def hasLowResolutionBlocks(self):
        return self.rank == 1  # one rank holding resolution-1 blocks switches everybody to full records

    def haloRequests(self):
        want = set()
        for b in self.own:
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dz in (-1, 0, 1):
                        k = (b[0] + dx, b[1] + dy, b[2] + dz)
                        if k != b and sharding.owner_of(np.array([k]), self.world, self.nb)[0] != self.rank:
                            want.add(k)
        return torch.tensor(sorted(want), dtype=torch.int32).reshape(-1, 3)

    @staticmethod
    def payload(key, owner):
        return np.array([1, owner, key[0] & 0xFF, key[1] & 0xFF, key[2] & 0xFF], np.int32)

    def haloPack(self, keys, full):
        assert full
        out = torch.zeros((len(keys), self.REC), dtype=torch.uint8)
        for i, k in enumerate(keys.tolist()):
            assert sharding.owner_of(np.array([k]), self.world, self.nb)[0] == self.rank  # routed to the owner
            rec = self.payload(k, self.rank) if tuple(k) in self.own else np.array([-1, 0, 0, 0, 0], np.int32)
            out[i, :20] = torch.from_numpy(rec.view(np.uint8))
        return out

    def haloInsert(self, keys, records, full):
        for k, r in zip(keys.tolist(), records.numpy()):
            self.ghosts[tuple(k)] = r[:20].view(np.int32).copy()


def _halo_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        blocks = np.unique(rng.integers(-6, 6, size=(300, 3)), axis=0).astype(np.int32)
        nb = 997
        g = _FakeShard(blocks, rank, world, nb)
        info = sharding.halo_exchange(g)
        assert info["full_blocks"] and info["requested"] == len(g.ghosts) > 0
        present = {tuple(b) for b in blocks.tolist()}
        n_found = 0
        for k, rec in g.ghosts.items():
            owner = int(sharding.owner_of(np.array([k]), world, nb)[0])
            assert owner != rank
            if k in present:
                assert np.array_equal(rec, _FakeShard.payload(k, owner)), (k, rec)
                n_found += 1
            else:
                assert rec[0] == -1
        assert n_found > 0
        # starve frames: compute_sharded min-reduces the z-buffer between computeBegin and computeEnd
        class _Starve:
            def __init__(self, starve):
                self.starve, self.log = starve, []
                self.z = torch.full((16,), 0x7F7F7F7F7F7F7F7F, dtype=torch.int64)
                self.z[rank::world] = torch.arange(len(self.z[rank::world]), dtype=torch.int64) + 10 * (rank + 1)

            def computeBegin(self):
                self.log.append("begin")
                return self.starve

            def zbufTensor(self):
                return self.z

            def computeEnd(self):
                self.log.append("end")

        s0 = _Starve(False)
        assert sharding.compute_sharded(s0) is False and s0.log == ["begin", "end"]
        s1 = _Starve(True)
        assert sharding.compute_sharded(s1) is True and s1.log == ["begin", "end"]
        want = torch.full((16,), 0x7F7F7F7F7F7F7F7F, dtype=torch.int64)
        for r in range(world):
            n = len(want[r::world])
            want[r::world] = torch.arange(n, dtype=torch.int64) + 10 * (r + 1)
        assert torch.equal(s1.z, want)  # every rank holds the per-cell minimum over the ranks
        # owner_of_torch == owner_of
        assert np.array_equal(sharding.owner_of_torch(torch.from_numpy(blocks), world, nb).numpy(), sharding.owner_of(blocks, world, nb))
        out.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover
        import traceback

        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_routing_world2():
    """Both all-to-alls of sharding.halo_exchange over gloo: every requested key reaches its owner
    and the owner's answer (or "not held") comes back to the rank that asked, in request order."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == {0: "ok", 1: "ok"}, results


def test_partition_and_exchange_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == {0: "ok", 1: "ok"}, results


def test_bucket_ranges_partition_the_table():
    for world in (1, 2, 3, 4, 8):
        for nb in (7, 1000, 250000):
            edges = [sharding.bucket_range(r, world, nb) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == nb
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


def test_block_hash_known_answers():
    # values derived from the formula of calculateHash
    assert sharding.block_hash([0, 0, 0], 250000) == 0
    assert sharding.block_hash([1, 2, 3], 250000) == 163698
    assert sharding.block_hash([1, 2, 3], 500000) == 163698
    assert sharding.block_hash([-1, -1, -1], 250000) == 112177
    assert sharding.block_hash([25, -12, 7], 250000) == 205824
    assert sharding.block_hash([25, -12, 7], 500000) == 455824
    assert sharding.block_hash([1000, -1000, 12345], 250000) == 184207

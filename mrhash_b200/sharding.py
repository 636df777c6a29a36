"""Multi-GPU partition of one map by hash-bucket range.

GPU g of G owns the block keys whose reference hash bucket (calculateHash,
/root/reference/mrhash/src/sdf/voxel_data_structures.cu:151-160) falls in
[g*nb/G, (g+1)*nb/G). Every rank sees the whole frame (rank 0 ingests it and broadcasts it), walks
all rays, but inserts / fuses / collects only the blocks it owns, so integration needs no exchange:
a voxel update depends only on the frame and the voxel's own state. Meshing reads the 26 neighbour
blocks, which under hash partitioning live on other ranks: before it, every rank fetches the
one-voxel shells of the neighbour blocks it does not own from their owners (`halo_exchange`, two
all-to-alls: 12-byte keys out, ~2.4 KB records back), meshes its own blocks in place, and only the
triangle soups are gathered for the weld (`extract_mesh_sharded`). `gather_blocks` /
`extract_mesh_gathered` is the older variant that ships whole shards to one rank.

torch.distributed is plumbing only (NCCL on the GPU box, gloo in the CPU tests).
"""
import numpy as np

P0, P1, P2 = 73856093, 19349669, 83492791  # params.h:102-104


def block_hash(blocks, num_buckets):
    """calculateHash for an int array [..., 3] of block coordinates."""
    b = np.asarray(blocks).astype(np.int64)
    u = b.astype(np.uint32).astype(np.uint64)
    h = ((u[..., 0] * P0) & 0xFFFFFFFF) ^ ((u[..., 1] * P1) & 0xFFFFFFFF) ^ ((u[..., 2] * P2) & 0xFFFFFFFF)
    return (h % np.uint64(num_buckets)).astype(np.int64)


def bucket_range(rank, world, num_buckets):
    """[lo, hi) of hash buckets owned by `rank` (same arithmetic as refresh_map_params in mrh_capi.cu)."""
    if world <= 1:
        return 0, int(num_buckets)
    return int(num_buckets) * rank // world, int(num_buckets) * (rank + 1) // world


def owner_of(blocks, world, num_buckets):
    """Rank that owns each block of an int array [..., 3]."""
    h = block_hash(blocks, num_buckets)
    if world <= 1:
        return np.zeros(h.shape, np.int64)
    bounds = np.array([bucket_range(r, world, num_buckets)[1] for r in range(world)], np.int64)
    return np.searchsorted(bounds, h, side="right")


def broadcast_frame(depth, rgb, pose, src=0, group=None):
    """Broadcast one frame (torch tensors, in place) from `src` to every rank.
    depth f32 [H,W], rgb u8 [H,W,3], pose f32 [7] = translation + quaternion xyzw."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.broadcast(pose, src, group=group)
    dist.broadcast(depth, src, group=group)
    dist.broadcast(rgb, src, group=group)


_COALESCED_GATHER = [True]  # cleared on the first failure of the grouped NCCL all-gather


def frame_row_band(rank, world, rows):
    """Rows [lo, hi) of a frame that `rank` uploads in scatter_ingest_frame (equal bands; rows % world == 0)."""
    band = rows // world
    return rank * band, (rank + 1) * band


def scatter_views(depth_dev, rgb_dev, group=None):
    """The tensor views scatter_ingest_frame needs for one pair of device images, built once by a caller
    that reuses its images (a per-frame loop then spends its host time on the copies and collectives only)."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    rows = depth_dev.shape[0]
    if rows % world:
        raise ValueError(f"scatter_ingest_frame: {rows} rows do not split into {world} equal bands")
    lo, hi = frame_row_band(rank, world, rows)
    return {"lo": lo, "hi": hi, "d_band": depth_dev[lo:hi], "c_band": rgb_dev[lo:hi], "d_all": depth_dev.view(-1), "c_all": rgb_dev.view(-1),
            "d_in": depth_dev[lo:hi].reshape(-1), "c_in": rgb_dev[lo:hi].reshape(-1)}


def scatter_ingest_frame(depth_dev, rgb_dev, depth_host, rgb_host, group=None, views=None):
    """Upload ONE frame to all ranks of a node without sending it down one PCIe link `world` times.

    Every rank maps the same host frame (one shared, page-locked segment in a server; in bench.py every
    rank's own copy of the synthetic stream) and uploads only its band of rows over ITS OWN PCIe link,
    straight into its place in the full device image; an in-place all-gather over NVLink / NVSwitch then
    completes the image on every rank. Host -> device bytes per rank: 1 / world of the frame; the
    collective moves the frame at NVLink rate. Runs on the current torch stream; the caller orders the
    handle's stream after it (events), as with broadcast_frame.
    depth_dev f32 [H,W] and rgb_dev u8 [H,W,3]: device, contiguous; *_host: the same shapes, pinned.
    views: scatter_views(depth_dev, rgb_dev), for callers that reuse their device images."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        depth_dev.copy_(depth_host, non_blocking=True)
        rgb_dev.copy_(rgb_host, non_blocking=True)
        return
    v = views if views is not None else scatter_views(depth_dev, rgb_dev, group)
    lo, hi = v["lo"], v["hi"]
    # (a caller that slices its host frames once hands over the band itself: [hi - lo, W] instead of [H, W])
    v["d_band"].copy_(depth_host if depth_host.shape[0] == hi - lo else depth_host[lo:hi], non_blocking=True)
    v["c_band"].copy_(rgb_host if rgb_host.shape[0] == hi - lo else rgb_host[lo:hi], non_blocking=True)
    # in place: each rank's input is its own band of the output (NCCL's in-place all-gather layout).
    # On NCCL both images go in ONE grouped launch (the per-collective host cost of torch.distributed is
    # what bounds a 8-rank step); elsewhere (gloo in the CPU tests) as two collectives.
    if depth_dev.is_cuda and _COALESCED_GATHER[0]:
        try:
            pg = group if group is not None else dist.distributed_c10d._get_default_group()
            work = pg.allgather_into_tensor_coalesced([v["d_all"], v["c_all"]], [v["d_in"], v["c_in"]])
            if work is not None:
                work.wait()  # stream-ordered on NCCL: the current stream waits, the host does not
            return
        except (AttributeError, RuntimeError, TypeError):
            _COALESCED_GATHER[0] = False  # this torch / backend has no coalesced all-gather: two calls from now on
    dist.all_gather_into_tensor(v["d_all"], v["d_in"], group=group)
    dist.all_gather_into_tensor(v["c_all"], v["c_in"], group=group)


def compute_sharded(geo, group=None):
    """compute() of one rank of a sharded map. Integration itself needs no exchange; the exception is
    the starve frame (every n_frames_invalidate_voxels-th): the per-pixel front-most voxel has to be the
    front-most of the WHOLE map, so the z-buffer (8 bytes per pixel) is min-reduced over the ranks
    between the two starve passes - one NCCL all-reduce every n-th frame, ordered with events only."""
    import torch
    import torch.distributed as dist

    if not geo.computeBegin():
        geo.computeEnd()
        return False
    zb = geo.zbufTensor()
    if not zb.is_cuda:  # host-side tests (gloo): no streams to order
        dist.all_reduce(zb, op=dist.ReduceOp.MIN, group=group)
        geo.computeEnd()
        return True
    lib_stream = torch.cuda.ExternalStream(geo.cudaStream(), device=zb.device)
    ev = torch.cuda.Event()
    ev.record(lib_stream)
    cur = torch.cuda.current_stream(zb.device)
    cur.wait_event(ev)
    dist.all_reduce(zb, op=dist.ReduceOp.MIN, group=group)
    ev2 = torch.cuda.Event()
    ev2.record(cur)
    lib_stream.wait_event(ev2)
    geo.computeEnd()
    return True


def gather_blocks(entries, voxels, dst=0, group=None, device="cpu"):
    """Gather every rank's (entries [n,5] int32, voxels [n,512] VOXEL_DTYPE) on `dst`.
    Returns the merged, key-sorted (entries, voxels) on dst and (None, None) elsewhere."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return entries, voxels
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([len(entries)], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts + [1])
    e = torch.zeros((n_max, 5), dtype=torch.int32, device=device)
    v = torch.zeros((n_max, 512 * 12), dtype=torch.uint8, device=device)
    if len(entries):
        e[: len(entries)] = torch.from_numpy(np.ascontiguousarray(entries)).to(device)
        v[: len(entries)] = torch.from_numpy(np.ascontiguousarray(voxels).view(np.uint8).reshape(len(entries), -1)).to(device)
    if rank == dst:
        es = [torch.zeros_like(e) for _ in range(world)]
        vs = [torch.zeros_like(v) for _ in range(world)]
    else:
        es = vs = None
    dist.gather(e, es, dst=dst, group=group)
    dist.gather(v, vs, dst=dst, group=group)
    if rank != dst:
        return None, None
    ee = np.concatenate([es[r][: counts[r]].cpu().numpy() for r in range(world)])
    vv = np.concatenate([vs[r][: counts[r]].cpu().numpy() for r in range(world)]).view(voxels.dtype).reshape(-1, 512)
    order = np.lexsort((ee[:, 2], ee[:, 1], ee[:, 0]))
    return ee[order], vv[order]


def owner_of_torch(keys, world, num_buckets):
    """owner_of for an int32 tensor [n,3] (any device), same arithmetic, in torch int64."""
    import torch

    k = keys.to(torch.int64) & 0xFFFFFFFF
    h = ((k[:, 0] * P0) & 0xFFFFFFFF) ^ ((k[:, 1] * P1) & 0xFFFFFFFF) ^ ((k[:, 2] * P2) & 0xFFFFFFFF)
    h = h % int(num_buckets)
    if world <= 1:
        return torch.zeros_like(h)
    bounds = torch.tensor([bucket_range(r, world, num_buckets)[1] for r in range(world)], dtype=torch.int64, device=keys.device)
    return torch.searchsorted(bounds, h, right=True)


def _wait(t):
    """NCCL collectives are asynchronous on torch's stream; the C ABI works on the handle's own
    stream. Wait for the collective that produced `t` before handing it to the library."""
    if t.is_cuda:
        import torch

        torch.cuda.current_stream(t.device).synchronize()
    return t


def _all_to_all_rows(rows, send_counts, group=None):
    """Variable-size all-to-all of the rows of a 2-D tensor (rows grouped by destination rank).
    Returns (received rows, receive counts)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sc = torch.tensor(send_counts, dtype=torch.int64, device=rows.device)
    rc = torch.empty(world, dtype=torch.int64, device=rows.device)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.tolist()]
    out = torch.empty((sum(recv_counts),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    dist.all_to_all_single(out, rows.contiguous(), output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
    return _wait(out), recv_counts


def halo_exchange(geo, group=None):
    """Boundary exchange before meshing a sharded map. `geo` needs haloRequests(),
    haloPack(keys, full), haloInsert(keys, records, full), hasLowResolutionBlocks(),
    getHashNumBuckets(); tensors live wherever haloRequests() puts them (CUDA with NCCL, CPU with
    gloo in the tests). Returns a dict of what was moved."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    nb = int(geo.getHashNumBuckets())
    req = geo.haloRequests()
    # whole blocks instead of shells as soon as any rank holds resolution-1 blocks
    flag = torch.tensor([1 if geo.hasLowResolutionBlocks() else 0], dtype=torch.int64, device=req.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    full = bool(int(flag.item()))
    owner = owner_of_torch(req, world, nb)
    order = torch.argsort(owner, stable=True)
    req = req[order].contiguous()
    send_counts = torch.bincount(owner, minlength=world).tolist()
    asked, asked_counts = _all_to_all_rows(req, send_counts, group)  # keys other ranks want from me
    records = geo.haloPack(asked, full)
    got, got_counts = _all_to_all_rows(records, asked_counts, group)  # answers, in the order I asked
    assert got_counts == [int(c) for c in send_counts] and len(got) == len(req)
    try:
        geo.haloInsert(req, got, full)
    except Exception:
        geo.haloClear()  # a partial insert must not leave ghost blocks (or the "exchange active" state) behind
        raise
    return {"requested": int(len(req)), "served": int(len(asked)), "full_blocks": full, "bytes_received": int(got.numel()), "bytes_sent": int(records.numel())}


def extract_mesh_sharded(geo, path, dst=0, group=None):
    """extractMesh for a hash-sharded map without moving the shards: boundary exchange, marching
    cubes over the owned blocks on every rank, gather of the triangle soups on `dst`, weld + PLY
    there. Returns (True, info) on dst, (False, info) elsewhere. The maps stay on their devices."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        geo.extractMesh(path)
        return True, {}
    info = halo_exchange(geo, group)
    try:
        soup = geo.meshLocal()
    finally:
        geo.haloClear()
    n = torch.tensor([len(soup)], dtype=torch.int64, device=soup.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    info["triangles_local"] = counts[rank]
    n_max = max(counts + [1])
    padded = torch.zeros((n_max, 3, 6), dtype=torch.float32, device=soup.device)
    padded[: len(soup)] = soup
    parts = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, parts, dst=dst, group=group)
    if rank != dst:
        return False, info
    whole = _wait(torch.cat([parts[r][: counts[r]] for r in range(world)]))
    info["triangles_total"] = int(len(whole))
    geo.weldSoup(whole, path)
    return True, info


def extract_mesh_gathered(geo, path, dst=0, group=None, device="cpu"):
    """extractMesh for a hash-sharded map: every rank hands its blocks to `dst` (one gather over the
    process group), which streams all of them into its own handle and meshes the whole map, so
    neighbour look-ups never cross a shard boundary. Returns True on `dst`. The shards stay as they
    were on the other ranks; on `dst` the whole map ends up in the host store (as extractMesh leaves it)."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        geo.extractMesh(path)
        return True
    entries, voxels = geo.dumpState()
    if rank == dst:
        geo.streamAllOut()  # own blocks -> host store
        mine = len(entries)
    all_e, all_v = gather_blocks(entries, voxels, dst=dst, group=group, device=device)
    if rank != dst:
        return False
    own = owner_of(all_e[:, :3], world, int(geo.getHashNumBuckets())) == rank
    assert int(own.sum()) == mine
    geo.storeAppend(all_e[~own], all_v[~own])
    geo.setShard(0, 1)
    try:
        geo.extractMesh(path)
    finally:
        geo.setShard(rank, world)
    return True

"""extractMesh + serializeData at scale (BASELINE configs[4]): grows the S2 room map at a fine voxel
size until it holds the requested number of blocks, then times marching cubes (device), the weld
and serializeData. Usage: python tools/bench_mesh.py [target_blocks] [voxel_size]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from mrhash_b200 import GeoWrapper, synth


def triangle_fingerprints(tris):
    """One 64-bit value per triangle from its 72 bytes (order of the three vertices kept, as both
    implementations emit them): equal sorted fingerprint arrays = the same multiset of bit-identical
    triangles, up to a 2^-64 collision chance per pair."""
    w = np.ascontiguousarray(tris, dtype=np.float32).reshape(len(tris), 18).view(np.uint32).astype(np.uint64)
    h = np.zeros(len(tris), np.uint64)
    with np.errstate(over="ignore"):
        for j in range(18):
            h = (h ^ w[:, j]) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(j + 1)
            h ^= h >> np.uint64(29)
    return np.sort(h)


def run(target=97657, voxel=0.003, with_ref=False, log=None, compare=False):
    w, h = 1280, 960
    p = dict(synth.REPLICA_PARAMS)
    p["virtual_voxel_size"] = voxel
    p["sdf_truncation"] = 7 * voxel
    g = GeoWrapper(**p, num_sdf_blocks=400000, hash_num_buckets=200000, max_num_triangles=40_000_000)
    ref = None
    if with_ref:
        from oracle_lib import RefCuda

        ref = RefCuda(p, 400000, 200000, max_num_triangles=8_000_000)
    fx, fy, cx, cy = synth.intrinsics(w, h)
    g.setCamera(fx, fy, cx, cy, h, w, p["min_depth"], p["max_depth"], 0)
    if ref is not None:
        ref.set_camera(fx, fy, cx, cy, h, w, p["min_depth"], p["max_depth"], 0)
    k = 0
    t0 = time.perf_counter()
    while True:
        # each pose is held for 6 frames so that voxels reach min_weight_threshold
        t, q, R = synth.orbit_pose(k // 6 * 8, 1000)
        d, c = synth.render_rgbd_torch(R, t, w, h, device="cuda")
        torch.cuda.synchronize()
        g.setCurrPose(t, q), g.setDepthImageDevice(d.data_ptr(), h, w), g.setRGBImageDevice(c.data_ptr(), h, w), g.compute(), g.synchronize()
        if ref is not None:
            ref.compute_rgbd(g.getCurrPose(), d.cpu().numpy(), c.cpu().numpy())
        k += 1
        if k % 6 == 0:
            st = g.getStats()
            if st["live_blocks"] >= target or k > 6000:
                break
    st = g.getStats()
    build_s = time.perf_counter() - t0
    if log:
        log(f"mesh bench: frames {k}, live blocks {st['live_blocks']} ({st['live_blocks'] * 512 / 1e6:.1f} M voxels), build {build_s:.1f} s")
    ref_info = None
    if ref is not None:
        # the reference's extractMesh device part: flatAndReduceHashTable + extractIsoSurface + D2H of the soup
        t0 = time.perf_counter()
        n_ref = ref.lib.ref_extract_triangles(ref.h, None, 0)
        ref_info = {"ref_kernel_and_d2h_s": time.perf_counter() - t0, "ref_triangles": int(n_ref)}
        if compare:
            ref_tris, _ = ref.extract_triangles(int(n_ref))
            ref_fp = triangle_fingerprints(ref_tris)
            del ref_tris
        del ref
    # the call sequence of rgbd_runner.py: extractMesh on the resident map, then serializeData
    t0 = time.perf_counter()
    g.extractMesh("/tmp/mesh_bench.ply")
    t_mesh = time.perf_counter() - t0
    tris, V, F = g.getTriangles(), g.getVertices(), g.getFaces()
    if ref_info is not None and compare:
        fp = triangle_fingerprints(tris)
        same = len(fp) == len(ref_fp) and bool((fp == ref_fp).all())
        ref_info["bit_identical_triangles"] = len(fp) if same else int(np.isin(fp, ref_fp).sum())

    t0 = time.perf_counter()
    g.serializeData("/tmp/hash_bench.ply", "/tmp/voxel_bench.ply")
    t_ser = time.perf_counter() - t0
    breakdown = {kk: g._get(kk) for kk in ("LastMeshStreamMs", "LastMeshKernelMs", "LastMeshMergeMs", "LastMeshPlyMs")}
    # marching cubes reads 12 B x (voxels of allocated blocks + 26-neighbour halo ~ 1.95x) and writes 72 B per triangle
    algo = 12.0 * 1.95 * st["live_blocks"] * 512 + 72.0 * len(tris)
    out = {
        "workload": "extractMesh + serializeData over a map of >= 97 657 blocks (50 M voxels) grown from scene S2 at 1280x960 (BASELINE configs[4])",
        "reference": ref_info,
        "blocks": st["live_blocks"],
        "voxels_M": st["live_blocks"] * 512 / 1e6,
        "triangles": len(tris),
        "vertices": len(V),
        "faces": len(F),
        "build_s": build_s,
        "stream_all_out_s": breakdown["LastMeshStreamMs"] / 1e3,  # the copy of the resident map into the host store inside extractMesh
        "extract_mesh_total_s": t_mesh,
        "mesh_breakdown_ms": breakdown,
        "mc_kernel_algorithmic_bytes": algo,
        "mc_kernel_gbs": algo / max(breakdown["LastMeshKernelMs"], 1e-6) / 1e6,
        "serialize_data_s": t_ser,
        "mesh_ply_MB": os.path.getsize("/tmp/mesh_bench.ply") / 1e6,
        "voxel_ply_MB": os.path.getsize("/tmp/voxel_bench.ply") / 1e6,
    }
    g.close()
    for f in ("/tmp/mesh_bench.ply", "/tmp/hash_bench.ply", "/tmp/voxel_bench.ply"):
        try:
            os.remove(f)
        except OSError:
            pass
    return out


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 97657, float(sys.argv[2]) if len(sys.argv) > 2 else 0.003, os.environ.get("MRH_BENCH_REF", "0") == "1", compare=os.environ.get("MRH_BENCH_COMPARE", "0") == "1", log=lambda s: print(s, file=sys.stderr, flush=True))))

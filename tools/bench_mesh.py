"""extractMesh + serializeData at scale (BASELINE configs[4]): grows the S2 room map at a fine voxel
size until it holds the requested number of blocks, then times marching cubes (device), the host
merge and serializeData. Usage: python tools/bench_mesh.py [target_blocks] [voxel_size]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mrhash_b200 import GeoWrapper, synth

target = int(sys.argv[1]) if len(sys.argv) > 1 else 97657
voxel = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
w, h = 1280, 960
p = dict(synth.REPLICA_PARAMS)
p["virtual_voxel_size"] = voxel
p["sdf_truncation"] = 7 * voxel
g = GeoWrapper(**p, num_sdf_blocks=400000, hash_num_buckets=200000, max_num_triangles=40_000_000)
with_ref = os.environ.get("MRH_BENCH_REF", "0") == "1"
ref = None
if with_ref:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
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
    g.setCurrPose(t, q); g.setDepthImageDevice(d.data_ptr(), h, w); g.setRGBImageDevice(c.data_ptr(), h, w); g.compute(); g.synchronize()
    if ref is not None:
        ref.compute_rgbd(g.getCurrPose(), d.cpu().numpy(), c.cpu().numpy())
    k += 1
    if k % 6 == 0:
        st = g.getStats()
        if st["live_blocks"] >= target or k > 6000:
            break
st = g.getStats()
print(f"frames {k}, live blocks {st['live_blocks']} ({st['live_blocks'] * 512 / 1e6:.1f} M voxels), build {time.perf_counter() - t0:.1f} s", flush=True)
ref_info = None
if ref is not None:
    # the reference's extractMesh device part: flatAndReduceHashTable + extractIsoSurface + D2H of the soup
    t0 = time.perf_counter()
    n_ref = ref.lib.ref_extract_triangles(ref.h, None, 0)
    ref_info = {"ref_kernel_and_d2h_s": time.perf_counter() - t0, "ref_triangles": int(n_ref)}
t0 = time.perf_counter(); g.streamAllOut(); t_out = time.perf_counter() - t0
t0 = time.perf_counter(); g.extractMesh("/tmp/mesh_bench.ply"); t_mesh = time.perf_counter() - t0
tris = g.getTriangles(); V = g.getVertices(); F = g.getFaces()
t0 = time.perf_counter(); g.serializeData("/tmp/hash_bench.ply", "/tmp/voxel_bench.ply"); t_ser = time.perf_counter() - t0
print(json.dumps({"reference": ref_info, "blocks": st["live_blocks"], "voxels_M": st["live_blocks"] * 512 / 1e6, "triangles": len(tris), "vertices": len(V), "faces": len(F), "stream_all_out_s": t_out, "extract_mesh_total_s": t_mesh, "mesh_breakdown_ms": {k: g._get(k) for k in ("LastMeshStreamMs", "LastMeshKernelMs", "LastMeshMergeMs", "LastMeshPlyMs")}, "serialize_data_s": t_ser, "mesh_ply_MB": os.path.getsize("/tmp/mesh_bench.ply") / 1e6, "voxel_ply_MB": os.path.getsize("/tmp/voxel_bench.ply") / 1e6}))

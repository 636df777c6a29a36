"""Device-side timeline of the end-to-end streaming path (bench.py pass B): upload of depth / colour and the
frame kernel of consecutive frames, microseconds since the first stamp. python tools/debug_timeline.py [w h]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mrhash_b200 import GeoWrapper, synth, _capi
w = int(sys.argv[1]) if len(sys.argv) > 1 else 640
h = int(sys.argv[2]) if len(sys.argv) > 2 else 480
n, first, count = 90, 40, 24
p = dict(synth.REPLICA_PARAMS)
frames = []
for k in range(n):
    t, q, R = synth.orbit_pose(k, 1000)
    d, c = synth.render_rgbd_torch(R, t, w, h, device="cuda")
    frames.append((t, q, d.cpu().pin_memory(), c.cpu().pin_memory()))
g = GeoWrapper(**p, num_sdf_blocks=500000, hash_num_buckets=250000, max_num_triangles=1)
fx, fy, cx, cy = synth.intrinsics(w, h)
g.setCamera(fx, fy, cx, cy, h, w, p["min_depth"], p["max_depth"], 0)
g.setIngestMode(2)
g.setStatsPipeline(True)
lib = _capi.lib()
lib.mrh_debug_timeline.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
lag = int(os.environ.get("LAG", "2"))
for k, (t, q, d, c) in enumerate(frames):
    if k == first:
        lib.mrh_debug_timeline(g._h, count, None)
    g.setCurrPose(t, q)
    g.setDepthImage(d.numpy())
    g.setRGBImage(c.numpy())
    g.compute()
    if k >= lag:
        g.getStatsPipelined(lag)
out = (C.c_float * (6 * count))()
got = lib.mrh_debug_timeline(g._h, 0, out)
v = np.array(list(out)).reshape(count, 6)[:got]
print("frame  depth upload      colour upload     frame kernel      (us since the first stamp)")
for i, r in enumerate(v):
    print(f"{i:3d}   {r[0]:7.1f}-{r[1]:7.1f}   {r[2]:7.1f}-{r[3]:7.1f}   {r[4]:7.1f}-{r[5]:7.1f}")
if got > 2:
    print(f"period {(v[-1, 5] - v[0, 5]) / (got - 1):.1f} us; depth upload {np.mean(v[:, 1] - v[:, 0]):.1f} us, colour upload {np.mean(v[:, 3] - v[:, 2]):.1f} us, kernel {np.mean(v[:, 5] - v[:, 4]):.1f} us")

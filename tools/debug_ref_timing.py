import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle_lib import RefCuda
from mrhash_b200 import synth
p = dict(synth.REPLICA_PARAMS)
w, h = 640, 480
fx, fy, cx, cy = synth.intrinsics(w, h)
os.makedirs("/tmp/mrh_ref_run", exist_ok=True); os.chdir("/tmp/mrh_ref_run")
r = RefCuda(p, 500000, 250000)
r.set_camera(fx, fy, cx, cy, h, w, p["min_depth"], p["max_depth"], 0)
frames = []
for k in range(0, 1000):
    t, q, R = synth.orbit_pose(k, 1000)
    if k % 100 == 0 or not frames:
        d, c = synth.render_rgbd_torch(R, t, w, h, device="cuda")
        frames.append((d.cpu().numpy(), c.cpu().numpy()))
    T = synth.quat_to_matrix_f32(t, q)
    d, c = frames[-1]
    t0 = time.perf_counter()
    r.compute_rgbd(T, d, c)
    dt = time.perf_counter() - t0
    if k % 100 == 0 or k == 999:
        free, total = torch.cuda.mem_get_info()
        print(k, f"wall {dt*1e3:.2f} ms integrate {r.last_integrate_ms():.3f} ms occupied {r.occupied()} gpu_used {(total-free)/2**30:.2f} GiB", flush=True)

"""Replays a few frames of the S2 stream through the C ABI (device-resident inputs) - the small
driver used under ncu (profiles/README). Usage: python tools/profile_frames.py [n_frames] [width] [height]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from mrhash_b200 import GeoWrapper, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
w = int(sys.argv[2]) if len(sys.argv) > 2 else 640
h = int(sys.argv[3]) if len(sys.argv) > 3 else 480
p = dict(synth.REPLICA_PARAMS)
g = GeoWrapper(**p, num_sdf_blocks=500000, hash_num_buckets=250000, max_num_triangles=1)
fx, fy, cx, cy = synth.intrinsics(w, h)
g.setCamera(fx, fy, cx, cy, h, w, p["min_depth"], p["max_depth"], 0)
frames = []
for k in range(n):
    t, q, R = synth.orbit_pose(k, 1000)
    d, c = synth.render_rgbd_torch(R, t, w, h, device="cuda")
    frames.append((t, q, d, c))
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for t, q, d, c in frames:
    flush.zero_()
    torch.cuda.synchronize()
    g.setCurrPose(t, q)
    g.setDepthImageDevice(d.data_ptr(), h, w)
    g.setRGBImageDevice(c.data_ptr(), h, w)
    g.compute()
    g.synchronize()
print(g.getStats())

"""Quick device-resident timing of the RGB-D frame (tuning helper, not the bench contract)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mrhash_b200 import GeoWrapper, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 250
w = int(sys.argv[2]) if len(sys.argv) > 2 else 640
h = int(sys.argv[3]) if len(sys.argv) > 3 else 480
orbit = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
p = dict(synth.REPLICA_PARAMS)
frames = []
for k in range(n):
    t, q, R = synth.orbit_pose(k, orbit)
    d, c = synth.render_rgbd_torch(R, t, w, h, device="cuda")
    frames.append((t, q, d, c))
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(flushed):
    g = GeoWrapper(**p, num_sdf_blocks=500000, hash_num_buckets=250000, max_num_triangles=1)
    fx, fy, cx, cy = synth.intrinsics(w, h)
    g.setCamera(fx, fy, cx, cy, h, w, p["min_depth"], p["max_depth"], 0)
    st = torch.cuda.ExternalStream(g.cudaStream())
    ev = []
    with torch.cuda.stream(st):
        for i, (t, q, d, c) in enumerate(frames):
            if flushed: flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            g.setCurrPose(t, q); g.setDepthImageDevice(d.data_ptr(), h, w); g.setRGBImageDevice(c.data_ptr(), h, w); g.compute()
            b.record(st)
            ev.append((a, b))
    g.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev[30:]]
    return 1e3 * float(np.mean(ms)), 1e3 * float(np.median(ms)), g.getStats()
for fl in (True, False):
    mean, med, st = run(fl)
    print(f"{w}x{h} {os.environ.get('MRH_LIB','default')[-20:]} frame={os.environ.get('MRH_FRAME','fused')} pref={os.environ.get('MRH_FUSED_PREF','-')} bulk={os.environ.get('MRH_BULK_DEPTH','-')} ctas={os.environ.get('MRH_FUSED_CTAS_PER_SM','-')} flushed={fl}: mean {mean:.1f} us median {med:.1f} us/frame  upd/frame {st['voxels_updated']/n:.0f} vis/frame {st['blocks_visible']/n:.0f} new/frame {st['blocks_new']/n:.1f}")

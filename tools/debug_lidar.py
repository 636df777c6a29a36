import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle_lib import Oracle
from mrhash_b200 import GeoWrapper, synth
from test_parity_lidar import make, K, ROWS, COLS
params = dict(synth.VBR_PARAMS)
ours, orc, _ = make(params, with_ref=False)
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for k in range(nf):
    T, pts = synth.lidar_frame(k, noise_sigma=0.01)
    ours.setCurrPoseMatrix(T); ours.setPointCloud(pts, False); ours.compute()
    orc.compute_points(T, pts)
ea, va = ours.dumpState(); eb, vb = orc.dump()
assert np.array_equal(ea[:, :3], eb[:, :3])
bad = np.argwhere(va["sdf"] != vb["sdf"])
print("frames", nf, "mismatch", len(bad), "of updated", (va["weight"] > 0).sum())
for bi, vi in bad[:12]:
    print(ea[bi, :3], vi, "ours", va[bi, vi], "orc", vb[bi, vi])
wbad = np.argwhere(va["weight"] != vb["weight"])
for bi, vi in wbad[:6]:
    print("W", ea[bi, :3], vi, "ours", va[bi, vi], "orc", vb[bi, vi])

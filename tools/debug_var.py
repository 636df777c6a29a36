import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from compare import compare_dumps
from oracle_lib import Oracle, RefCuda, ref_available
from mrhash_b200 import GeoWrapper, synth
import test_parity_rgbd as tr
import test_parity_lidar as tl

mode = sys.argv[1]; nf = int(sys.argv[2]); thr = float(sys.argv[3])
if mode == "rgbd":
    params = dict(synth.REPLICA_PARAMS); params["sdf_var_threshold"] = thr
    ours, orc, ref = tr.make_all(params, width=320, height=240)
    for k in range(nf):
        t, q, depth, rgb = synth.rgbd_frame(k, n_frames=400, width=320, height=240, noise_sigma=0.002)
        tr.feed(ours, [orc, ref], t, q, depth, rgb)
        a = ours.dumpState(); st = ours.getStats()
        print(k, "ours blocks", len(a[0]), "res1", int((a[0][:,3]==1).sum()), "heap", st["heap_free"], st["heap_low_free"], "| orc", orc.heap_high_free(), orc.heap_low_free(), "| ref", ref.heap_high_free(), ref.heap_low_free())
        print("   vs orc", {k2: v for k2, v in compare_dumps(a, orc.dump()).items() if "mismatch" in k2 or "only" in k2})
        print("   vs ref", {k2: v for k2, v in compare_dumps(a, ref.dump()).items() if "mismatch" in k2 or "only" in k2})
else:
    params = dict(synth.VBR_PARAMS); params["sdf_var_threshold"] = thr
    ours, orc, ref = tl.make(params)
    for k in range(nf):
        T, pts = synth.lidar_frame(k, noise_sigma=0.01)
        ours.setCurrPoseMatrix(T); ours.setPointCloud(pts, False); ours.compute()
        orc.compute_points(T, pts); ref.compute_points(T, pts)
        a = ours.dumpState(); st = ours.getStats()
        print(k, "ours blocks", len(a[0]), "res1", int((a[0][:,3]==1).sum()), "heap", st["heap_free"], st["heap_low_free"], "| orc", orc.heap_high_free(), orc.heap_low_free(), "| ref", ref.heap_high_free(), ref.heap_low_free())
        print("   vs orc", {k2: v for k2, v in compare_dumps(a, orc.dump()).items() if "mismatch" in k2 or "only" in k2})
        print("   vs ref", {k2: v for k2, v in compare_dumps(a, ref.dump()).items() if "mismatch" in k2 or "only" in k2})

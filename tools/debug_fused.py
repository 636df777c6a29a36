"""Per-role timers of the fused frame kernel (needs a -DMRH_FUSED_DEBUG build: tools/build_variant.sh dbg -DMRH_FUSED_DEBUG,
then MRH_LIB=mrhash_b200/libmrhash_b200_dbg.so python tools/debug_fused.py [frames] [w] [h])."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mrhash_b200 import GeoWrapper, synth, _capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
w = int(sys.argv[2]) if len(sys.argv) > 2 else 640
h = int(sys.argv[3]) if len(sys.argv) > 3 else 480
p = dict(synth.REPLICA_PARAMS)
g = GeoWrapper(**p, num_sdf_blocks=500000, hash_num_buckets=250000, max_num_triangles=1)
fx, fy, cx, cy = synth.intrinsics(w, h)
g.setCamera(fx, fy, cx, cy, h, w, p["min_depth"], p["max_depth"], 0)
lib = _capi.lib()
lib.mrh_debug_timers.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
buf = (C.c_uint64 * 32)()
names = {0: "exit", 1: "chunk", 2: "tile", 3: "fuse"}
for k in range(n):
    t, q, R = synth.orbit_pose(k, 1000)
    d, c = synth.render_rgbd_torch(R, t, w, h, device="cuda")
    torch.cuda.synchronize()
    if k == 0:
        g.setCurrPose(t, q); g.setDepthImageDevice(d.data_ptr(), h, w); g.setRGBImageDevice(c.data_ptr(), h, w); g.compute(); g.synchronize()
        lib.mrh_debug_timers(g._h, buf)
        continue
    g.setCurrPose(t, q); g.setDepthImageDevice(d.data_ptr(), h, w); g.setRGBImageDevice(c.data_ptr(), h, w); g.compute(); g.synchronize()
    lib.mrh_debug_timers(g._h, buf)
    v = list(buf)
    t0 = v[0]
    if k % 5 == 0 or k < 4:
        us = lambda x: (x - t0) / 1e3
        print(f"frame {k}: last CTA start +{us(v[1]):.1f} us | first/last CTA out of loop +{us(v[2]):.1f}/+{us(v[3]):.1f} | finaliser +{us(v[20]):.1f} | fq {v[21]} q_tile {v[22]}")
        for r in (1, 2, 3):
            cnt = max(v[8 + r], 1)
            print(f"   {names[r]:5s}: items {v[8+r]:6d} sum {v[4+r]/1e3:9.1f} us  mean {v[4+r]/cnt/1e3:7.2f} us  last ended +{us(v[12+r]):.1f} us  longest {v[24+r]/1e3:.1f} us")
        print(f"   sched: sum {v[16]/1e3:.1f} us max {v[17]/1e3:.1f} us")
        print(f"   walked patches {v[29]}: ray off the fast path {v[18]}, box too large {v[19]}, block missing {v[28]}")
    if k >= n - 3 and os.environ.get('MRH_TRACE_LONG'):  # with a -DMRH_TRACE_LONG build: the tile items longer than 8 us
        ntr = min(int(v[30]), 8192)
        tb = (C.c_uint64 * (2 * max(ntr, 1)))()
        lib.mrh_debug_trace.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_size_t]
        lib.mrh_debug_trace(g._h, tb, 2 * ntr)
        t0l = t0 & 0xFFFFFFFF
        rows = sorted((int(tb[2 * i] >> 32), (((tb[2 * i + 1] >> 32) - t0l) & 0xFFFFFFFF) / 1e3, (((tb[2 * i + 1] & 0xFFFFFFFF) - t0l) & 0xFFFFFFFF) / 1e3) for i in range(ntr))
        tiles_x = (w + 31) // 32
        print(f'frame {k}: long tile items ((column,row) begin-end us): ' + ' '.join(f"({t % tiles_x},{t // tiles_x})[{a:.0f}-{b:.0f}]" for t, a, b in rows))
    elif k == n - 1:
        ntr = min(int(v[30]), 8192)
        tb = (C.c_uint64 * (2 * ntr))()
        lib.mrh_debug_trace.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_size_t]
        lib.mrh_debug_trace(g._h, tb, 2 * ntr)
        t0l = t0 & 0xFFFFFFFF
        rows = {}
        for i in range(ntr):
            cta, kind = tb[2 * i] >> 32, tb[2 * i] & 0xFF
            a, b = tb[2 * i + 1] >> 32, tb[2 * i + 1] & 0xFFFFFFFF
            rows.setdefault(cta, []).append((((a - t0l) & 0xFFFFFFFF) / 1e3, ((b - t0l) & 0xFFFFFFFF) / 1e3, names.get(kind, "?")))
        for cta in sorted(rows)[:40]:
            print(f"cta {cta:5d}: " + " ".join(f"{nm[0]}[{a:.1f}-{b:.1f}]" for a, b, nm in sorted(rows[cta])))
print(g.getStats())

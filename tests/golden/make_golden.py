#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference kernels (oracle/_ref, built from
/root/reference by oracle/Makefile) on small seeded synthetic inputs. Needs a GPU:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy into tests/golden/

The fixtures pin the CPU oracle to the real reference in the `-m "not gpu"` suite
(tests/test_oracle_golden.py). Kept small: per-block checksums instead of full payloads.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden_cases import CASES, block_digest, run_case, triangle_digest  # noqa: E402


def main():
    from oracle_lib import RefCuda

    out = sys.argv[1] if len(sys.argv) > 1 else HERE
    os.makedirs(out, exist_ok=True)
    for name, case in CASES.items():
        ref = RefCuda(case["params"], case["num_blocks"], case["num_buckets"], max_num_triangles=case.get("max_triangles", 0))
        entries, voxels, tris = run_case(case, ref)
        data = dict(entries=entries[:, :4].astype(np.int32), digest=block_digest(entries, voxels, case), first_voxels=voxels[:4].copy())
        if case.get("racy"):
            # the reference's LiDAR update is racy: keep the raw sdf / weight of the first blocks so the
            # test can compare exactly the voxels that a single point touched
            data.update(head_sdf=voxels["sdf"][:256].copy(), head_weight=voxels["weight"][:256].copy())
        if tris is not None:
            data.update(triangle_digest(tris))
        np.savez_compressed(os.path.join(out, name + ".npz"), **data)
        print(name, "blocks", len(entries), "triangles", None if tris is None else len(tris))


if __name__ == "__main__":
    main()

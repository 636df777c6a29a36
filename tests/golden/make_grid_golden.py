"""Generates tests/golden/grid_cista.bin + grid_cista.npz: a serializeGrid checkpoint written by the
reference's OWN serialization library (cista, vendored by the reference; built into
oracle/_ref/cista_grid_tool by `make -C oracle ref`, which needs /root/reference) from seeded random
blocks. Run in the build container:  python tests/golden/make_grid_golden.py"""
import os
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from mrhash_b200 import VOXEL_DTYPE  # noqa: E402

TOOL = os.path.join(ROOT, "oracle", "_ref", "cista_grid_tool")
VOXEL_SIZE, EXTENTS = 0.01, 1.0


def chunk_of(pos, size=VOXEL_SIZE, ext=EXTENTS):
    """Streamer::worldToChunks of the block origin (streamer.cuh:251-260, streamer.cpp:230-232), float32."""
    pw = (pos.astype(np.float32) * np.float32(8.0)) * np.float32(size)
    p = pw / np.float32(ext)
    return (p + np.sign(p).astype(np.float32) * np.float32(0.5)).astype(np.int32)


def blocks_to_tool_input(entries, voxels):
    out = [struct.pack("<I", len(entries))]
    for e, v in zip(entries, voxels):
        nv = 512 if e[3] == 0 else 64
        out.append(struct.pack("<3i3iiiI", *chunk_of(e[:3]).tolist(), int(e[0]), int(e[1]), int(e[2]), int(e[4]), int(e[3]), nv))
        out.append(v[:nv].tobytes())
    return b"".join(out)


def main():
    rng = np.random.default_rng(11)
    n = 40
    pos = rng.integers(-40, 40, size=(n, 3)).astype(np.int32)
    pos[:6] = [[0, 0, 0], [12, 0, 0], [13, 0, 0], [-1, -1, -1], [-13, 5, 7], [6, 6, 6]]  # chunk borders, negatives
    res = (rng.random(n) < 0.25).astype(np.int32)
    ptr = np.where(res == 0, rng.integers(0, 1000, n) * 512, rng.integers(0, 8000, n) * 64).astype(np.int32)
    entries = np.concatenate([pos, res[:, None], ptr[:, None]], axis=1).astype(np.int32)
    voxels = np.zeros((n, 512), VOXEL_DTYPE)
    raw = rng.integers(0, 256, size=(n, 512 * 12), dtype=np.uint8)
    voxels[:] = raw.view(VOXEL_DTYPE).reshape(n, 512)
    for i in range(n):
        if res[i]:
            voxels[i, 64:] = np.zeros(1, VOXEL_DTYPE)[0]  # a resolution-1 block has 64 voxels
    tmp = os.path.join(HERE, "_blocks.tmp")
    open(tmp, "wb").write(blocks_to_tool_input(entries, voxels))
    subprocess.check_call([TOOL, "encode", tmp, os.path.join(HERE, "grid_cista.bin")])
    os.remove(tmp)
    np.savez_compressed(os.path.join(HERE, "grid_cista.npz"), entries=entries, voxels=voxels.view(np.uint8).reshape(n, -1), voxel_size=VOXEL_SIZE, extents=EXTENTS)
    print("wrote grid_cista.bin", os.path.getsize(os.path.join(HERE, "grid_cista.bin")), "bytes,", n, "blocks")


if __name__ == "__main__":
    main()

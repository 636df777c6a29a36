"""CPU: serializeGrid / deserializeGrid file format (csrc/mrh_grid.cu, no device needed) pinned
against a checkpoint written by the reference's own serialization library: tests/golden/grid_cista.bin
comes from cista (the reference's vendored utils/cista.h, compiled into oracle/_ref/cista_grid_tool) with
the record types of streamer.cuh:21-165 and the framing of serializer.h:16-75
(tests/golden/make_grid_golden.py)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from mrhash_b200 import VOXEL_DTYPE
from mrhash_b200.geowrapper import grid_read, grid_write

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "grid_cista.bin")
TOOL = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "cista_grid_tool")


def golden_blocks():
    z = np.load(os.path.join(HERE, "golden", "grid_cista.npz"))
    return z["entries"], z["voxels"].view(VOXEL_DTYPE).reshape(len(z["entries"]), 512), float(z["voxel_size"]), float(z["extents"])


def records(path):
    """{chunk: bytes} of a grid file (the chunk order of a file is the iteration order of an unordered_map)."""
    data, out, off = open(path, "rb").read(), {}, 0
    while off < len(data):
        (size,) = struct.unpack_from("<Q", data, off)
        chunk = struct.unpack_from("<3i", data, off + 8)
        assert chunk not in out
        out[chunk] = data[off + 20 : off + 20 + size]
        off += 20 + size
    assert off == len(data)
    return out


def by_key(entries, voxels):
    return {tuple(e[:3]): (tuple(e[3:]), v.tobytes()) for e, v in zip(entries.tolist(), voxels)}


def test_reader_decodes_the_cista_file():
    entries, voxels, _, _ = golden_blocks()
    e, v = grid_read(GOLD)
    assert len(e) == len(entries) and (e[:, 3] == 1).sum() > 3
    assert by_key(e, v) == by_key(entries, voxels)


def test_writer_reproduces_the_cista_file_byte_for_byte(tmp_path):
    entries, voxels, size, ext = golden_blocks()
    path = str(tmp_path / "grid.bin")
    grid_write(path, entries, voxels, size, ext)
    mine, gold = records(path), records(GOLD)
    assert set(mine) == set(gold) and len(gold) > 10
    for chunk in gold:
        assert mine[chunk] == gold[chunk], chunk
    assert os.path.getsize(path) == os.path.getsize(GOLD)


def test_round_trip_and_errors(tmp_path):
    rng = np.random.default_rng(3)
    n = 300
    pos = np.unique(rng.integers(-200, 200, size=(n, 3)), axis=0).astype(np.int32)
    entries = np.concatenate([pos, np.zeros((len(pos), 1), np.int32), (np.arange(len(pos))[:, None] * 512).astype(np.int32)], axis=1)
    voxels = rng.integers(0, 256, size=(len(pos), 512 * 12), dtype=np.uint8).view(VOXEL_DTYPE).reshape(len(pos), 512)
    path = str(tmp_path / "rt.bin")
    grid_write(path, entries, voxels, 0.2, 1.0)
    e, v = grid_read(path)
    assert by_key(e, v) == by_key(entries, voxels)
    grid_write(str(tmp_path / "empty.bin"), entries[:0], voxels[:0], 0.2, 1.0)
    assert len(grid_read(str(tmp_path / "empty.bin"))[0]) == 0
    with pytest.raises(RuntimeError, match="Failed to open file for reading"):
        grid_read(str(tmp_path / "missing.bin"))
    # a truncated / corrupted file is reported, not read past its end
    data = open(path, "rb").read()
    open(str(tmp_path / "cut.bin"), "wb").write(data[: len(data) // 2 + 30])
    with pytest.raises(RuntimeError, match="Read failed|Corrupted"):
        grid_read(str(tmp_path / "cut.bin"))
    bad = bytearray(data)
    bad[20:28] = struct.pack("<q", 1 << 40)  # first vector header points far outside its record
    open(str(tmp_path / "bad.bin"), "wb").write(bytes(bad))
    with pytest.raises(RuntimeError, match="Corrupted"):
        grid_read(str(tmp_path / "bad.bin"))


@pytest.mark.skipif(not os.path.exists(TOOL), reason="oracle/_ref/cista_grid_tool not built (needs /root/reference)")
def test_cista_itself_reads_what_the_writer_produces(tmp_path):
    """The other direction: cista::deserialize (through the tool) decodes a file of our writer."""
    entries, voxels, size, ext = golden_blocks()
    path, back = str(tmp_path / "grid.bin"), str(tmp_path / "blocks.bin")
    grid_write(path, entries, voxels, size, ext)
    subprocess.check_call([TOOL, "decode", path, back])
    data = open(back, "rb").read()
    (n,) = struct.unpack_from("<I", data, 0)
    assert n == len(entries)
    off, seen = 4, {}
    for _ in range(n):
        c0, c1, c2, x, y, z, ptr, res, nv = struct.unpack_from("<3i3iiiI", data, off)
        off += 36
        assert nv == (512 if res == 0 else 64)
        seen[(x, y, z)] = ((res, ptr), data[off : off + 12 * nv])
        off += 12 * nv
    want = {k: (a, b[: 12 * (512 if a[0] == 0 else 64)]) for k, (a, b) in by_key(entries, voxels).items()}
    assert seen == want


def test_ply_number_formatter_equals_printf_g():
    """csrc/mrh_fmt.h (the ASCII PLY writer's fast path) against printf's %g on floats widened to double:
    random bit patterns, the mesh's usual range, decade edges and decimal ties."""
    import ctypes as C

    from mrhash_b200 import _capi

    lib = _capi.lib()
    buf = C.create_string_buffer(64)
    rng = np.random.default_rng(9)
    vals = [rng.integers(0, 2**32, 60000, dtype=np.uint64).astype(np.uint32).view(np.float32), (rng.random(60000) * 20 - 10).astype(np.float32),
            (rng.random(20000) * 2e-3 - 1e-3).astype(np.float32), (rng.random(20000) * 2e6 - 1e6).astype(np.float32),
            np.array([0.0, -0.0, 0.5, 1.5, 2.5, 0.125, 1e-4, 9.99999e-5, 1e-5, 0.001, 0.1, 1, 10, 100000, 999999, 999999.5, 1e6, 123456.5, 12345.65, 1.0000005, 9.999995, 99999.95, 262144.5], np.float32)]
    vals.append(np.nextafter(vals[-1], np.float32(1e9)))
    vals.append(np.nextafter(vals[-2], np.float32(-1e9)))
    n = 0
    for arr in vals:
        with np.errstate(invalid="ignore"):  # signalling NaN bit patterns among the random floats
            wide = arr.astype(np.float64).tolist()
        for v in wide:
            if not np.isfinite(v):
                continue
            for x in (v, -v):
                k = lib.mrh_format_g6(x, buf)
                assert buf.raw[:k].decode() == "%g" % x, x
                n += 1
    assert n > 300000


def test_ascii_ply_writer_on_the_host(tmp_path):
    """The mesh PLY writer of extractMesh (geowrapper.cpp:194-229: ASCII, `ostream << double`, colours cast
    to uchar) through its handle-free hook: header, every line byte for byte against printf, across the
    batch boundaries of the pipelined writer (batch = threads x 65536 lines)."""
    import ctypes as C

    from mrhash_b200 import _capi

    lib = _capi.lib()
    rng = np.random.default_rng(12)
    nv, nf = 1_200_000, 300_001
    V = (rng.random((nv, 3)) * 8 - 4).astype(np.float32).astype(np.float64)
    V[:5] = [[0, -0.0, 1], [1e-5, 123456.5, -2.5], [0.1, 0.25, 1e6], [-3.0000001, 2.9999998, 0.015625], [1e-4, 9.99999e-5, 999999.5]]
    Cc = rng.random((nv, 3)) * 300 - 20  # un-normalised colours wrap like the reference's uchar cast (Q4)
    F = rng.integers(0, nv, size=(nf, 3)).astype(np.int32)
    path = str(tmp_path / "mesh.ply")
    assert lib.mrh_write_mesh_ply(path.encode(), V.ctypes.data, Cc.ctypes.data, F.ctypes.data, nv, nf) == 0
    lines = open(path).read().split("\n")
    h = lines.index("end_header")
    assert lines[:h] == ["ply", "format ascii 1.0", f"element vertex {nv}", "property float x", "property float y", "property float z", "property uchar red",
                         "property uchar green", "property uchar blue", f"element face {nf}", "property list uchar int vertex_indices"]
    assert len(lines) == h + 1 + nv + nf + 1 and lines[-1] == ""
    idx = np.unique(np.concatenate([np.arange(0, 2000), np.arange(65536 - 50, 65536 + 50), np.arange(8 * 65536 - 50, 8 * 65536 + 50), np.arange(16 * 65536 - 50, 16 * 65536 + 50),
                                    rng.integers(0, nv, 20000), np.arange(nv - 2000, nv)]))
    for i in idx.tolist():
        v, c = V[i], Cc[i]
        assert lines[h + 1 + i] == "%g %g %g %d %d %d" % (v[0], v[1], v[2], int(c[0]) & 0xFF, int(c[1]) & 0xFF, int(c[2]) & 0xFF), i
    for i in np.unique(np.concatenate([np.arange(0, 2000), rng.integers(0, nf, 20000), np.arange(nf - 2000, nf)])).tolist():
        assert lines[h + 1 + nv + i] == "3 %d %d %d" % tuple(F[i]), i

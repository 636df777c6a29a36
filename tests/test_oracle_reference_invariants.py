"""CPU: the invariants and known answers the reference's own tests state, asserted on the oracle
(reference tests live in /root/reference/mrhash/tests; every test here cites the one it mirrors)."""
import ctypes as C

import numpy as np
import pytest

from oracle_lib import Oracle, build_oracle

from mrhash_b200 import synth

LIB = C.CDLL(build_oracle())
LIB.orc_calculate_hash.restype = C.c_uint32
LIB.orc_calculate_hash.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32]
LIB.orc_voxel_to_block_index.restype = C.c_uint32


def w2v(size, p):
    out = (C.c_int * 3)()
    LIB.orc_world_to_voxel(C.c_float(size), (C.c_float * 3)(*p), out)
    return list(out)


def v2b(v, size, ext=1.0):
    out = (C.c_int * 3)()
    LIB.orc_voxel_to_block((C.c_int * 3)(*v), C.c_float(size), (C.c_float * 3)(ext, ext, ext), out)
    return list(out)


def test_hash_known_answers():
    # calculateHash (voxel_data_structures.cu:151-160); values of SURVEY.md §8c
    kat = [((0, 0, 0), 250000, 0), ((1, 2, 3), 250000, 163698), ((1, 2, 3), 500000, 163698), ((-1, -1, -1), 250000, 112177),
           ((25, -12, 7), 250000, 205824), ((25, -12, 7), 500000, 455824), ((1000, -1000, 12345), 250000, 184207)]
    for (x, y, z), n, want in kat:
        assert LIB.orc_calculate_hash(x, y, z, n) == want


def test_world_voxel_block_known_answers():
    # world -> voxel -> block -> local x at virtual_voxel_size 0.01
    s = 0.01
    rows = [(0.004, 0, 0, 0), (0.005, 1, 0, 1), (0.074, 7, 0, 7), (0.075, 8, 1, 0), (-0.004, 0, 0, 0), (-0.005, -1, -1, 7),
            (-0.075, -8, -1, 0), (-0.085, -9, -2, 7), (2.0, 200, 25, 0)]
    for x, vox, blk, local in rows:
        v = w2v(s, (x, 0.0, 0.0))
        assert v[0] == vox, (x, v)
        assert v2b(v, s)[0] == blk, (x, v2b(v, s))
        assert LIB.orc_voxel_to_block_index((C.c_int * 3)(*v), 8) % 8 == local


def test_coordinate_round_trips():
    # tests/test_hash_utils.cu:40-163 VOXEL.*: (47.32, 52.45, 150.23) at 1e-6 m voxels, tolerance 1e-4
    s = 1e-6
    p = (47.32, 52.45, 150.23)
    v = w2v(s, p)
    back = [vi * np.float32(s) for vi in v]
    assert np.allclose(back, p, atol=1e-4)
    b = v2b(v, s)
    local = LIB.orc_voxel_to_block_index((C.c_int * 3)(*v), 8)
    lx, ly, lz = local % 8, (local // 8) % 8, local // 64
    re = [(b[k] * 8 + l) * np.float32(s) for k, l in enumerate((lx, ly, lz))]
    assert np.allclose(re, p, atol=1e-4)
    for bs in (8, 4, 2):  # delinearize / linearize are inverse for every block size the reference tests
        for idx in (0, 1, bs * bs * bs - 1, bs * bs + bs + 1):
            out = (C.c_uint32 * 3)()
            LIB.orc_delinearize(idx, bs, out)
            assert out[2] * bs * bs + out[1] * bs + out[0] == idx


def test_buffer_initialisation():
    # tests/test_hash_utils.cu:306-376 HASHTABLE.BufferInitialization
    n, nb = 5000, 2500
    o = Oracle(synth.REPLICA_PARAMS, n, nb)
    LIB.orc_heap_high.restype = C.POINTER(C.c_uint32)
    LIB.orc_heap_high.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    counter = C.c_int()
    heap = LIB.orc_heap_high(o.h, C.byref(counter))
    assert counter.value == n - 1
    assert all(heap[i] == n - 1 - i for i in range(n))
    LIB.orc_table.restype = C.c_void_p
    LIB.orc_table.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    total = C.c_uint32()
    ptr = LIB.orc_table(o.h, C.byref(total))
    assert total.value == nb * 10
    table = np.ctypeslib.as_array((C.c_int32 * (6 * total.value)).from_address(ptr)).reshape(-1, 6)
    assert (table[:, :3] == 0).all() and (table[:, 3] == 0).all() and (table[:, 4] == -2).all()
    assert o.heap_high_free() == n and o.heap_low_free() == 0


def constant_depth_setup(n_blocks=60000, n_buckets=30000):
    # tests/test_hash_utils.cu:175-190: 400x400 constant depth 1.0, K = (400,400,200,200), identity
    # pose, voxel 5 mm, truncation 0.02 + 0.01 z, weight sample 3 (sizes scaled down for the CPU)
    params = dict(synth.REPLICA_PARAMS, sdf_truncation=0.02, sdf_truncation_scale=0.01, integration_weight_sample=3, virtual_voxel_size=0.005, n_frames_invalidate_voxels=10, min_depth=0.1, max_depth=5.0)
    o = Oracle(params, n_blocks, n_buckets, threads=4)
    o.set_camera(400.0, 400.0, 200.0, 200.0, 400, 400, 0.1, 5.0, 0)
    depth = np.full((400, 400), 1.0, np.float32)
    rgb = np.full((400, 400, 3), 128, np.uint8)
    return o, depth, rgb, n_blocks


def test_heap_sanity_after_one_frame():
    # tests/test_hash_utils.cu:378-526 HASHTABLE.HeapSanityCheck
    o, depth, rgb, n = constant_depth_setup()
    o.compute_rgbd(np.eye(4, dtype=np.float32), depth, rgb)
    entries, voxels = o.dump()
    assert len(entries) > 1000 and o.overflow_events() == 0
    assert len({tuple(e[:3]) for e in entries}) == len(entries)  # no duplicate block positions
    ptrs = entries[:, 4]
    assert len(set(ptrs.tolist())) == len(ptrs) and (ptrs % 512 == 0).all()
    assert len(entries) + o.heap_high_free() == n  # occupied + free == num_sdf_blocks
    LIB.orc_heap_high.restype = C.POINTER(C.c_uint32)
    LIB.orc_heap_high.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    counter = C.c_int()
    heap = LIB.orc_heap_high(o.h, C.byref(counter))
    free = {heap[i] for i in range(counter.value + 1)}
    assert len(free) == counter.value + 1  # free pointers unique
    assert free.isdisjoint(set((ptrs // 512).tolist()))  # no pointer both free and allocated
    w = voxels["weight"]
    assert set(np.unique(w).tolist()) <= {0, 3}
    sdf = voxels["sdf"][w > 0]
    assert (np.abs(sdf) <= 0.02 + 0.01 * 1.0 + 1e-6).all()


def test_garbage_collection_frees_empty_blocks():
    # tests/test_hash_utils.cu:192-304 HASHTABLE.AllocationDeletion: allocated + free == total after GC,
    # and blocks whose voxels never came within the truncation band are collected
    o, depth, rgb, n = constant_depth_setup()
    o.compute_rgbd(np.eye(4, dtype=np.float32), depth, rgb)
    st = o.stats()
    entries, voxels = o.dump()
    assert st["blocks_new"] - st["blocks_freed"] == len(entries)
    assert len(entries) + o.heap_high_free() == n
    # every surviving block holds at least one voxel with weight > 0 and |sdf| below the GC threshold
    thr = 0.02 + 0.01 * 5.0
    ok = ((voxels["weight"] > 0) & (np.abs(voxels["sdf"]) < thr)).any(axis=1)
    assert ok.all()


def test_projection_round_trips():
    # tests/test_projections.cu:41-221: back-projected z == depth (pinhole), range == depth (spherical),
    # project(inverseProject) returns the pixel
    o = Oracle(synth.REPLICA_PARAMS, 100, 100)
    LIB.orc_inverse_projection.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.POINTER(C.c_float)]
    LIB.orc_project_point.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rng = np.random.default_rng(0)
    o.set_camera(517.3, 516.5, 318.6, 255.3, 480, 640, 0.0, 10.0, 0)
    for _ in range(2000):
        r, c, d = int(rng.integers(0, 480)), int(rng.integers(0, 640)), float(rng.uniform(0.1, 10.0))
        p = (C.c_float * 3)()
        LIB.orc_inverse_projection(o.h, r, c, d, p)
        assert p[2] == np.float32(d)
        rr, cc = C.c_int(), C.c_int()
        assert LIB.orc_project_point(o.h, p, C.byref(rr), C.byref(cc))
        assert abs(rr.value - r) <= 1 and abs(cc.value - c) <= 1
    rows, cols = 128, 1024
    o.set_camera(-cols / (2 * np.pi), -rows / (np.pi / 2), cols / 2, rows / 2, rows, cols, 0.0, 100.0, 1)
    for _ in range(2000):
        r, c, d = int(rng.integers(1, rows - 1)), int(rng.integers(1, cols - 1)), float(rng.uniform(0.5, 90.0))
        p = (C.c_float * 3)()
        LIB.orc_inverse_projection(o.h, r, c, d, p)
        assert abs(np.sqrt(p[0] ** 2 + p[1] ** 2 + p[2] ** 2) - d) < 2e-2 * max(1.0, d) * 1e-2 + 1e-4
        rr, cc = C.c_int(), C.c_int()
        assert LIB.orc_project_point(o.h, p, C.byref(rr), C.byref(cc))
        assert abs(rr.value - r) <= 1 and abs(cc.value - c) <= 1


def test_quaternion_pose_matches_scipy():
    # geowrapper.cpp:86-92: Eigen::Quaternionf(w, x, y, z).toRotationMatrix()
    from scipy.spatial.transform import Rotation

    LIB.orc_quat_to_matrix.argtypes = [C.POINTER(C.c_float)] * 3
    rng = np.random.default_rng(1)
    for _ in range(50):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        t = rng.normal(size=3)
        out = (C.c_float * 16)()
        LIB.orc_quat_to_matrix((C.c_float * 3)(*t), (C.c_float * 4)(*q), out)
        T = np.array(out).reshape(4, 4)
        assert np.allclose(T[:3, :3], Rotation.from_quat(q).as_matrix(), atol=1e-6)
        assert np.allclose(T[:3, 3], t, atol=1e-6)
        assert np.array_equal(T, synth.quat_to_matrix_f32(t.astype(np.float32), q.astype(np.float32)))

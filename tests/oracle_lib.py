"""ctypes wrappers of the test-infrastructure libraries under oracle/ (never imported by the product):

* ``Oracle``  - oracle/libmrh_oracle.so, the CPU restatement (oracle/mrh_oracle.c)
* ``RefCuda`` - oracle/_ref/libref_harness.so, the UNMODIFIED reference CUDA sources compiled for
                sm_100a behind oracle/ref_harness/harness.cu (needs a GPU; prebuilt in the build
                container, travels with the snapshot)
Both expose: set_camera, compute_rgbd, compute_points, dump, extract_triangles, heap_high_free.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libmrh_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")

VOXEL_DTYPE = np.dtype([("sdf", "<f4"), ("sum_squared", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("weight", "u1")])


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ROOT, "oracle", "mrh_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return ORACLE_SO


class FrameStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays_valid", "blocks_new", "blocks_visible", "voxels_updated", "blocks_freed", "blocks_realloc")]


def _sort_dump(entries, voxels):
    if len(entries) == 0:
        return entries, voxels
    order = np.lexsort((entries[:, 2], entries[:, 1], entries[:, 0]))
    return entries[order], voxels[order]


_CREATE_ARGS = [C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int]


def _create_args(p, num_sdf_blocks, hash_num_buckets):
    return (
        num_sdf_blocks,
        hash_num_buckets,
        p["sdf_truncation"],
        p["sdf_truncation_scale"],
        p["integration_weight_sample"],
        p["virtual_voxel_size"],
        p["n_frames_invalidate_voxels"],
        p["voxel_extents_scale"],
        p["marching_cubes_threshold"],
        p["min_weight_threshold"],
        p.get("sdf_var_threshold", 0.0),
        int(p.get("projective_sdf", True)),
    )


class Oracle:
    def __init__(self, params, num_sdf_blocks, hash_num_buckets, threads=1):
        self.lib = l = C.CDLL(build_oracle())
        l.orc_create.argtypes = _CREATE_ARGS
        l.orc_create.restype = C.c_void_p
        l.orc_destroy.argtypes = [C.c_void_p]
        l.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
        l.orc_set_camera.argtypes = [C.c_void_p] + [C.c_float] * 4 + [C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
        l.orc_compute_rgbd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        l.orc_compute_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        l.orc_last_stats.argtypes = [C.c_void_p, C.POINTER(FrameStats)]
        l.orc_heap_high_free.argtypes = [C.c_void_p]
        l.orc_heap_low_free.argtypes = [C.c_void_p]
        l.orc_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        l.orc_dump.restype = C.c_uint32
        l.orc_extract_triangles.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        l.orc_extract_triangles.restype = C.c_uint32
        l.orc_overflow_events.argtypes = [C.c_void_p]
        l.orc_overflow_events.restype = C.c_uint32
        self.h = l.orc_create(*_create_args(params, num_sdf_blocks, hash_num_buckets))
        l.orc_set_threads(self.h, threads)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def set_camera(self, fx, fy, cx, cy, rows, cols, min_depth, max_depth, model):
        self.lib.orc_set_camera(self.h, fx, fy, cx, cy, rows, cols, min_depth, max_depth, model)

    def compute_rgbd(self, pose44, depth, rgb):
        pose = np.ascontiguousarray(pose44, np.float32)
        depth = np.ascontiguousarray(depth, np.float32)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        self.lib.orc_compute_rgbd(self.h, pose.ctypes.data, depth.ctypes.data, rgb.ctypes.data, depth.shape[0], depth.shape[1])

    def compute_points(self, pose44, points, normals=None):
        pose = np.ascontiguousarray(pose44, np.float32)
        pts = np.ascontiguousarray(points, np.float32)
        nrm = None if normals is None else np.ascontiguousarray(normals, np.float32)
        self.lib.orc_compute_points(self.h, pose.ctypes.data, pts.ctypes.data, None if nrm is None else nrm.ctypes.data, pts.shape[0])

    def stats(self):
        s = FrameStats()
        self.lib.orc_last_stats(self.h, C.byref(s))
        return {n: int(getattr(s, n)) for n, _ in s._fields_}

    def heap_high_free(self):
        return self.lib.orc_heap_high_free(self.h)

    def heap_low_free(self):
        return self.lib.orc_heap_low_free(self.h)

    def overflow_events(self):
        return self.lib.orc_overflow_events(self.h)

    def dump(self):
        n = self.lib.orc_dump(self.h, None, None, 0)
        entries = np.zeros((n, 5), np.int32)
        voxels = np.zeros((n, 512), VOXEL_DTYPE)
        if n:
            self.lib.orc_dump(self.h, entries.ctypes.data, voxels.ctypes.data, n)
        return _sort_dump(entries, voxels)

    def extract_triangles(self, max_out=1 << 22):
        out = np.zeros((max_out, 3, 6), np.float32)
        n = self.lib.orc_extract_triangles(self.h, out.ctypes.data, max_out)
        return out[: min(n, max_out)].copy(), n


def ref_available():
    return os.path.exists(REF_SO)


class RefCuda:
    """The reference's own kernels (GPU only)."""

    def __init__(self, params, num_sdf_blocks, hash_num_buckets, max_num_triangles=0):
        self.lib = l = C.CDLL(REF_SO)
        l.ref_create.argtypes = _CREATE_ARGS + [C.c_uint32]
        l.ref_create.restype = C.c_void_p
        l.ref_destroy.argtypes = [C.c_void_p]
        l.ref_set_camera.argtypes = [C.c_void_p] + [C.c_float] * 4 + [C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
        l.ref_compute_rgbd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        l.ref_compute_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        l.ref_last_integrate_ms.argtypes = [C.c_void_p]
        l.ref_last_integrate_ms.restype = C.c_float
        l.ref_heap_high_free.argtypes = [C.c_void_p]
        l.ref_heap_low_free.argtypes = [C.c_void_p]
        l.ref_current_occupied_blocks.argtypes = [C.c_void_p]
        l.ref_current_occupied_blocks.restype = C.c_uint32
        l.ref_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        l.ref_dump.restype = C.c_uint32
        l.ref_extract_triangles.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        l.ref_extract_triangles.restype = C.c_uint32
        self.h = l.ref_create(*_create_args(params, num_sdf_blocks, hash_num_buckets), max_num_triangles)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_destroy(self.h)
            self.h = None

    def set_camera(self, fx, fy, cx, cy, rows, cols, min_depth, max_depth, model):
        self.lib.ref_set_camera(self.h, fx, fy, cx, cy, rows, cols, min_depth, max_depth, model)

    def compute_rgbd(self, pose44, depth, rgb):
        pose = np.ascontiguousarray(pose44, np.float32)
        depth = np.ascontiguousarray(depth, np.float32)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        self.lib.ref_compute_rgbd(self.h, pose.ctypes.data, depth.ctypes.data, rgb.ctypes.data, depth.shape[0], depth.shape[1])

    def compute_points(self, pose44, points, normals=None):
        pose = np.ascontiguousarray(pose44, np.float32)
        pts = np.ascontiguousarray(points, np.float32)
        nrm = None if normals is None else np.ascontiguousarray(normals, np.float32)
        self.lib.ref_compute_points(self.h, pose.ctypes.data, pts.ctypes.data, None if nrm is None else nrm.ctypes.data, pts.shape[0])

    def last_integrate_ms(self):
        return self.lib.ref_last_integrate_ms(self.h)

    def heap_high_free(self):
        return self.lib.ref_heap_high_free(self.h)

    def heap_low_free(self):
        return self.lib.ref_heap_low_free(self.h)

    def occupied(self):
        return self.lib.ref_current_occupied_blocks(self.h)

    # ---- the reference's host mesh post-processing (mesh_extractor.cpp, unmodified, in oracle/_ref) ----
    def process_soup(self, soup, eps, merge=False):
        """MeshExtractor::processTriangles over a [T, 3, 6] float32 soup -> (V f64 [n,3], F i32 [m,3], C f64 [n,3])."""
        soup = np.ascontiguousarray(soup, np.float32)
        nv, nf = C.c_uint32(), C.c_uint32()
        self.lib.ref_process_soup.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        rc = self.lib.ref_process_soup(self.h, soup.ctypes.data, len(soup), eps, 1 if merge else 0, C.byref(nv), C.byref(nf))
        assert rc == 0, "ref_process_soup: soup larger than max_num_triangles"
        V, F, Cc = np.zeros((nv.value, 3)), np.zeros((nf.value, 3), np.int32), np.zeros((nv.value, 3))
        self.lib.ref_get_processed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.ref_get_processed(self.h, V.ctypes.data, F.ctypes.data, Cc.ctypes.data)
        return V, F, Cc

    def remove_duplicate_vertices(self, V, F, eps):
        V, F = np.ascontiguousarray(V, np.float64).reshape(-1, 3), np.ascontiguousarray(F, np.int32).reshape(-1, 3)
        Vo, Fo, mp = np.zeros_like(V), np.zeros_like(F), np.zeros(len(V), np.int32)
        n = C.c_uint32()
        self.lib.ref_remove_duplicate_vertices.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_double, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p]
        self.lib.ref_remove_duplicate_vertices(self.h, V.ctypes.data, len(V), F.ctypes.data, len(F), eps, Vo.ctypes.data, C.byref(n), Fo.ctypes.data, mp.ctypes.data)
        return Vo[: n.value], Fo, mp

    def remove_duplicate_faces(self, F):
        F = np.ascontiguousarray(F, np.int32).reshape(-1, 3)
        Fo = np.zeros_like(F)
        n = C.c_uint32()
        self.lib.ref_remove_duplicate_faces.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
        self.lib.ref_remove_duplicate_faces(self.h, F.ctypes.data, len(F), Fo.ctypes.data, C.byref(n))
        return Fo[: n.value]

    # ---- the reference's Streamer (streamer.cpp / streamer.cu, unmodified, in oracle/_ref) ----
    def streamer_create(self, max_blocks_per_pass=100000):
        self.lib.ref_streamer_create.argtypes = [C.c_void_p, C.c_uint32]
        assert self.lib.ref_streamer_create(self.h, max_blocks_per_pass) == 0

    def stream(self, position, radius, force=False):
        pos = (C.c_float * 3)(*[float(x) for x in position])
        self.lib.ref_stream.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_float, C.c_int]
        return self.lib.ref_stream(self.h, pos, radius, 1 if force else 0)

    def stream_all_out(self):
        self.lib.ref_stream_all_out.argtypes = [C.c_void_p]
        assert self.lib.ref_stream_all_out(self.h) == 0

    def serialize_data(self, hash_path, voxel_path):
        self.lib.ref_serialize_data.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        assert self.lib.ref_serialize_data(self.h, hash_path.encode(), voxel_path.encode()) == 0

    def duplicates_ratio(self):
        self.lib.ref_duplicates_ratio.argtypes = [C.c_void_p]
        self.lib.ref_duplicates_ratio.restype = C.c_double
        return self.lib.ref_duplicates_ratio(self.h)

    def grid_blocks(self):
        self.lib.ref_grid_blocks.argtypes = [C.c_void_p]
        self.lib.ref_grid_blocks.restype = C.c_uint32
        return self.lib.ref_grid_blocks(self.h)

    def dump(self):
        n = self.lib.ref_dump(self.h, None, None, 0)
        entries = np.zeros((n, 5), np.int32)
        voxels = np.zeros((n, 512), VOXEL_DTYPE)
        if n:
            self.lib.ref_dump(self.h, entries.ctypes.data, voxels.ctypes.data, n)
        return _sort_dump(entries, voxels)

    def extract_triangles(self, max_out=1 << 22):
        out = np.zeros((max_out, 3, 6), np.float32)
        n = self.lib.ref_extract_triangles(self.h, out.ctypes.data, max_out)
        return out[: min(n, max_out)].copy(), n

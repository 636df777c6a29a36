"""ctypes binding of libmrhash_b200.so (include/mrhash_b200.h).

The library is built in-tree by ``mrhash_b200/build.sh`` (``__graft_entry__.build()``). There is no
CPU fallback: if the shared object is missing or no CUDA device is present, calls fail loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MRH_LIB", os.path.join(_HERE, "libmrhash_b200.so"))  # MRH_LIB: tuning builds only


class Params(C.Structure):
    _fields_ = [
        ("sdf_truncation", C.c_float),
        ("sdf_truncation_scale", C.c_float),
        ("integration_weight_sample", C.c_int32),
        ("virtual_voxel_size", C.c_float),
        ("n_frames_invalidate_voxels", C.c_int32),
        ("voxel_extents_scale", C.c_int32),
        ("viewer_active", C.c_int32),
        ("marching_cubes_threshold", C.c_float),
        ("min_weight_threshold", C.c_int32),
        ("min_depth", C.c_float),
        ("max_depth", C.c_float),
        ("sdf_var_threshold", C.c_float),
        ("vertices_merging_threshold", C.c_float),
        ("projective_sdf", C.c_int32),
        ("num_sdf_blocks", C.c_uint64),
        ("hash_num_buckets", C.c_uint64),
        ("max_num_triangles", C.c_uint64),
        ("device", C.c_int32),
        ("shard_rank", C.c_int32),
        ("shard_world", C.c_int32),
        ("reserved", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("frames", C.c_uint64),
        ("rays_valid", C.c_uint64),
        ("blocks_new", C.c_uint64),
        ("blocks_visible", C.c_uint64),
        ("voxels_updated", C.c_uint64),
        ("blocks_freed", C.c_uint64),
        ("blocks_realloc", C.c_uint64),
        ("dropped_heap", C.c_uint64),
        ("dropped_table", C.c_uint64),
        ("live_blocks", C.c_uint64),
        ("heap_free", C.c_int64),
        ("heap_low_free", C.c_int64),
        ("dropped_updates", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class DumpEntry(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("z", C.c_int32), ("resolution", C.c_int32), ("ptr", C.c_int32)]


_P = C.POINTER
_vp = C.c_void_p
_f = C.c_float
_i = C.c_int
_fp = _P(C.c_float)
_u8p = _P(C.c_uint8)

# name -> (argtypes, restype); every function declared in include/mrhash_b200.h
SIGNATURES = {
    "mrh_last_error": ([], C.c_char_p),
    "mrh_abi_version": ([], _i),
    "mrh_params_default": ([_P(Params)], _i),
    "mrh_create": ([_P(Params), _P(_vp)], _i),
    "mrh_destroy": ([_vp], _i),
    "mrh_set_camera": ([_vp, _f, _f, _f, _f, _i, _i, _f, _f, _i], _i),
    "mrh_set_pose": ([_vp, _fp, _fp], _i),
    "mrh_set_pose_matrix": ([_vp, _fp], _i),
    "mrh_get_pose_matrix": ([_vp, _fp], _i),
    "mrh_set_camera_in_lidar": ([_vp, _fp], _i),
    "mrh_set_depth": ([_vp, _vp, _i, _i], _i),
    "mrh_set_rgb": ([_vp, _vp, _i, _i], _i),
    "mrh_set_rgb_f32": ([_vp, _vp, _i, _i], _i),
    "mrh_set_depth_device": ([_vp, _vp, _i, _i], _i),
    "mrh_set_rgb_device": ([_vp, _vp, _i, _i], _i),
    "mrh_set_points": ([_vp, _vp, C.c_size_t, _vp], _i),
    "mrh_compute": ([_vp], _i),
    "mrh_compute_begin": ([_vp, _P(_i)], _i),
    "mrh_compute_end": ([_vp], _i),
    "mrh_get_zbuf": ([_vp, _P(_vp), _P(C.c_size_t)], _i),
    "mrh_stream": ([_vp, _fp, _f], _i),
    "mrh_synchronize": ([_vp], _i),
    "mrh_stream_all_out": ([_vp], _i),
    "mrh_store_append": ([_vp, _vp, _vp, C.c_size_t], _i),
    "mrh_set_shard": ([_vp, _i, _i], _i),
    "mrh_store_size": ([_vp, _P(C.c_size_t)], _i),
    "mrh_store_read": ([_vp, _vp, _vp, C.c_size_t, _P(C.c_size_t)], _i),
    "mrh_extract_mesh": ([_vp, C.c_char_p], _i),
    "mrh_extract_mesh_ex": ([_vp, C.c_char_p, _i], _i),
    "mrh_get_mesh": ([_vp, _P(_P(C.c_double)), _P(_P(C.c_int32)), _P(_P(C.c_double)), _P(C.c_size_t), _P(C.c_size_t)], _i),
    "mrh_get_triangles": ([_vp, _P(_fp), _P(C.c_size_t)], _i),
    "mrh_write_mesh_ply": ([C.c_char_p, _vp, _vp, _vp, C.c_size_t, C.c_size_t], _i),
    "mrh_serialize_data": ([_vp, C.c_char_p, C.c_char_p], _i),
    "mrh_serialize_grid": ([_vp, C.c_char_p], _i),
    "mrh_deserialize_grid": ([_vp, C.c_char_p], _i),
    "mrh_grid_write": ([C.c_char_p, _vp, _vp, C.c_size_t, _f, _f], _i),
    "mrh_grid_read": ([C.c_char_p, _vp, _vp, C.c_size_t, _P(C.c_size_t)], _i),
    "mrh_format_g6": ([C.c_double, C.c_char_p], C.c_size_t),
    "mrh_clear_buffers": ([_vp], _i),
    "mrh_get_field": ([_vp, C.c_char_p, _P(C.c_double)], _i),
    "mrh_set_field": ([_vp, C.c_char_p, C.c_double], _i),
    "mrh_get_stats": ([_vp, _P(Stats)], _i),
    "mrh_reset_stats": ([_vp], _i),
    "mrh_last_compute_ms": ([_vp, _fp], _i),
    "mrh_set_profiling": ([_vp, _i], _i),
    "mrh_get_kernel_times": ([_vp, _P(C.c_double), _P(C.c_uint64)], _i),
    "mrh_get_stream": ([_vp, _P(_vp)], _i),
    "mrh_get_launch_count": ([_vp, _P(C.c_uint64)], _i),
    "mrh_dump_state": ([_vp, _vp, _vp, C.c_size_t, _P(C.c_size_t)], _i),
    "mrh_halo_record_bytes": ([_i], C.c_size_t),
    "mrh_halo_requests": ([_vp, _vp, C.c_size_t, _P(C.c_size_t)], _i),
    "mrh_halo_pack": ([_vp, _vp, C.c_size_t, _i, _vp], _i),
    "mrh_halo_insert": ([_vp, _vp, _vp, C.c_size_t, _i], _i),
    "mrh_halo_clear": ([_vp], _i),
    "mrh_mesh_local": ([_vp, _P(C.c_size_t)], _i),
    "mrh_copy_triangles_device": ([_vp, _vp, C.c_size_t], _i),
    "mrh_weld_device_soup": ([_vp, _vp, C.c_size_t, C.c_char_p], _i),
    "mrh_set_ingest_mode": ([_vp, _i], _i),
    "mrh_set_stats_pipeline": ([_vp, _i], _i),
    "mrh_get_stats_pipelined": ([_vp, _i, _P(Stats)], _i),
    "mrh_selftest_div": ([_i, _f, C.c_uint64, C.c_uint64, _P(C.c_uint64)], _i),
    "mrh_get_block_shortcut_radius": ([_vp, _P(_i)], _i),
}

_lib = None


def lib():
    """Load libmrhash_b200.so (once). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with mrhash_b200/build.sh (or __graft_entry__.build()); "
                "mrhash_b200 has no CPU fallback"
            )
        l = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = l
    return _lib


def check(status):
    if status != 0:
        raise RuntimeError(lib().mrh_last_error().decode("utf-8", "replace"))

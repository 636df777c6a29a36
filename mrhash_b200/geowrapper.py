"""``GeoWrapper`` — host-side mirror of the reference's ``pygeowrapper.GeoWrapper`` class.

Same constructor keywords, method names, argument meaning and error behaviour as the nanobind class
in /root/reference/mrhash/src/sdf/pybind/pygeowrapper.cpp:12-84 (over geowrapper.h:18-260), so that
``apps/rgbd_runner.py`` runs unmodified against it (``from mrhash.src.pygeowrapper import GeoWrapper``
is provided by the ``mrhash/`` shim package at the repo root). All work happens in libmrhash_b200.so
through the C ABI of include/mrhash_b200.h; this file only converts numpy arguments.

Additions over the reference API (keyword-only, all optional): explicit sizing
(``num_sdf_blocks``, ``hash_num_buckets``, ``max_num_triangles``), ``device``, and the
hash-bucket-range shard of a multi-GPU run (``shard_rank``, ``shard_world``); plus ``synchronize()``,
``getStats()``, ``dumpState()`` used by tests and bench.py.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import check


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class GeoWrapper:
    # pygeowrapper.cpp:14-29
    def __init__(
        self,
        sdf_truncation,
        sdf_truncation_scale,
        integration_weight_sample,
        virtual_voxel_size,
        n_frames_invalidate_voxels,
        voxel_extents_scale,
        viewer_active,
        marching_cubes_threshold,
        min_weight_threshold,
        min_depth,
        max_depth,
        gs_optimization_param_path="",
        sdf_var_threshold=0.0,
        vertices_merging_threshold=0.0,
        projective_sdf=True,
        *,
        num_sdf_blocks=0,
        hash_num_buckets=0,
        max_num_triangles=0,
        device=-1,
        shard_rank=0,
        shard_world=1,
    ):
        self._h = None
        if gs_optimization_param_path:
            raise RuntimeError("GeoWrapper: Gaussian-splatting side branch (gs_optimization_param_path) is out of scope of mrhash_b200")
        self._lib = _capi.lib()
        p = _capi.Params()
        check(self._lib.mrh_params_default(C.byref(p)))
        p.sdf_truncation = float(sdf_truncation)
        p.sdf_truncation_scale = float(sdf_truncation_scale)
        p.integration_weight_sample = int(integration_weight_sample)
        p.virtual_voxel_size = float(virtual_voxel_size)
        p.n_frames_invalidate_voxels = int(n_frames_invalidate_voxels)
        p.voxel_extents_scale = int(voxel_extents_scale)
        p.viewer_active = int(bool(viewer_active))
        p.marching_cubes_threshold = float(marching_cubes_threshold)
        p.min_weight_threshold = int(min_weight_threshold)
        p.min_depth = float(min_depth)
        p.max_depth = float(max_depth)
        p.sdf_var_threshold = float(sdf_var_threshold)
        p.vertices_merging_threshold = float(vertices_merging_threshold)
        p.projective_sdf = int(bool(projective_sdf))
        p.num_sdf_blocks = int(num_sdf_blocks)
        p.hash_num_buckets = int(hash_num_buckets)
        p.max_num_triangles = int(max_num_triangles)
        p.device = int(device)
        p.shard_rank = int(shard_rank)
        p.shard_world = int(shard_world)
        h = C.c_void_p()
        check(self._lib.mrh_create(C.byref(p), C.byref(h)))
        self._h = h
        self._shard_world = int(shard_world)
        self._device = int(self._get("Device"))
        self._points = np.zeros((0, 3), np.float32)
        self._normals = np.zeros((0, 3), np.float32)
        self._extra = {}

    def __del__(self):
        self.close()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mrh_destroy(self._h)
            self._h = None

    # ---- getters / setters (pygeowrapper.cpp:31-61, geowrapper.h:79-109) ----
    def _get(self, name):
        v = C.c_double()
        check(self._lib.mrh_get_field(self._h, name.encode(), C.byref(v)))
        return v.value

    def _set(self, name, value):
        check(self._lib.mrh_set_field(self._h, name.encode(), float(value)))

    def getHashNumBuckets(self):
        return int(self._extra.get("HashNumBuckets", self._get("HashNumBuckets")))

    def getNumSdfBlocks(self):
        return int(self._extra.get("NumSDFBlocks", self._get("NumSDFBlocks")))

    def getHashBucketSize(self):
        return int(self._extra.get("HashBucketSize", self._get("HashBucketSize")))

    def getSdfTruncation(self):
        return self._get("SDFTruncation")

    def getSdfTruncationScale(self):
        return self._get("SDFTruncationScale")

    def getIntegrationWeightSample(self):
        return int(self._get("IntegrationWeightSample"))

    def getIntegrationWeightMax(self):
        return int(self._extra.get("IntegrationWeightMax", self._get("IntegrationWeightMax")))

    def getVirtualVoxelSize(self):
        return self._get("VirtualVoxelSize")

    def getLinkedListSize(self):
        return int(self._extra.get("LinkedListSize", self._get("LinkedListSize")))

    def getNFramesInvalidateVoxels(self):
        return int(self._get("NFramesInvalidateVoxels"))

    def getMaxNumSdfBlockIntegrateFromGlobalHash(self):
        return int(self._extra.get("MaxNumSdfBlockIntegrateFromGlobalHash", self._get("MaxNumSdfBlockIntegrateFromGlobalHash")))

    def getVoxelExtentsScale(self):
        return int(self._extra.get("VoxelExtentsScale", self._get("VoxelExtentsScale")))

    # The reference's sizing setters only overwrite GeoWrapper's own copies after the container was
    # built (geowrapper.h:98-109): the getters echo the value, the map is unchanged. Same here.
    def setHashNumBuckets(self, v):
        self._extra["HashNumBuckets"] = int(v)

    def setNumSdfBlocks(self, v):
        self._extra["NumSDFBlocks"] = int(v)

    def setHashBucketSize(self, v):
        self._extra["HashBucketSize"] = int(v)

    def setIntegrationWeightMax(self, v):
        self._extra["IntegrationWeightMax"] = int(v)

    def setLinkedListSize(self, v):
        self._extra["LinkedListSize"] = int(v)

    def setMaxNumSdfBlockIntegrateFromGlobalHash(self, v):
        self._extra["MaxNumSdfBlockIntegrateFromGlobalHash"] = int(v)

    def setVoxelExtentsScale(self, v):
        self._extra["VoxelExtentsScale"] = int(v)

    def setVirtualVoxelSize(self, v):
        self._extra["VirtualVoxelSize"] = float(v)

    def setSdfTruncation(self, v):
        self._set("SDFTruncation", v)

    def setSdfTruncationScale(self, v):
        self._set("SDFTruncationScale", v)

    def setIntegrationWeightSample(self, v):
        self._set("IntegrationWeightSample", v)

    def setNFramesInvalidateVoxels(self, v):
        self._set("NFramesInvalidateVoxels", v)

    # ---- sensor / pose ----
    def setCamera(self, fx, fy, cx, cy, rows, cols, min_depth, max_depth, camera_model):
        model = int(getattr(camera_model, "value", camera_model))
        check(self._lib.mrh_set_camera(self._h, fx, fy, cx, cy, int(rows), int(cols), min_depth, max_depth, model))

    def setCurrPose(self, pose, orientation):
        t = _f32(pose).reshape(-1)
        q = _f32(orientation).reshape(-1)
        if t.size != 3 or q.size != 4:
            raise TypeError("setCurrPose(pose[3], orientation[4] = qx, qy, qz, qw)")
        check(self._lib.mrh_set_pose(self._h, t.ctypes.data_as(_capi._fp), q.ctypes.data_as(_capi._fp)))

    def setCurrPoseMatrix(self, cam_in_world):
        T = _f32(cam_in_world).reshape(16)
        check(self._lib.mrh_set_pose_matrix(self._h, T.ctypes.data_as(_capi._fp)))

    def getCurrPose(self):
        out = np.zeros(16, np.float32)
        check(self._lib.mrh_get_pose_matrix(self._h, out.ctypes.data_as(_capi._fp)))
        return out.reshape(4, 4)

    def setCameraInLidar(self, camera_in_lidar):
        T = _f32(camera_in_lidar)
        if T.shape != (4, 4):
            raise TypeError("setCameraInLidar expects a 4x4 matrix")
        check(self._lib.mrh_set_camera_in_lidar(self._h, T.ctypes.data_as(_capi._fp)))

    # ---- frame data (geowrapper.cpp:246-321, 345-505) ----
    def setDepthImage(self, input_depth_array):
        a = np.asarray(input_depth_array)
        if a.ndim != 2:
            raise RuntimeError("GeoWrapper::setDepthImage|input should be a 2D numpy array")
        a = _f32(a)
        check(self._lib.mrh_set_depth(self._h, a.ctypes.data, a.shape[0], a.shape[1]))

    def setRGBImage(self, input_rgb_array):
        a = np.asarray(input_rgb_array)
        if a.ndim != 3:
            raise RuntimeError("GeoWrapper::setRGBImage|input should be a 3D numpy array")
        if a.shape[2] != 3:
            raise RuntimeError("GeoWrapper::setRGBImage|input should have 3 channels")
        if a.dtype == np.uint8:
            a = np.ascontiguousarray(a)
            check(self._lib.mrh_set_rgb(self._h, a.ctypes.data, a.shape[0], a.shape[1]))
        else:
            # apps/utils/depth_reader.py:83-93 hands float32 colour; nanobind casts element-wise
            a = _f32(a)
            check(self._lib.mrh_set_rgb_f32(self._h, a.ctypes.data, a.shape[0], a.shape[1]))

    def setDepthImageDevice(self, ptr, rows, cols):
        check(self._lib.mrh_set_depth_device(self._h, int(ptr), int(rows), int(cols)))

    def setRGBImageDevice(self, ptr, rows, cols):
        check(self._lib.mrh_set_rgb_device(self._h, int(ptr), int(rows), int(cols)))

    def setPointCloud(self, input_point_cloud, normals_or_flag=False):
        pts = np.asarray(input_point_cloud)
        if pts.ndim != 2:
            raise RuntimeError("GeoWrapper::setPointCloud|input should be a 2D numpy array")
        pts = _f32(pts)
        nrm = None
        if isinstance(normals_or_flag, (bool, np.bool_)):
            if normals_or_flag:
                raise RuntimeError("GeoWrapper::setPointCloud|compute_normals=True (MAD-tree normals) is out of scope; every shipped runner passes False")
        else:
            nrm = np.asarray(normals_or_flag)
            if nrm.ndim != 2:
                raise RuntimeError("GeoWrapper::setPointCloud|normals input should be a 2D numpy array")
            if nrm.shape[0] != pts.shape[0]:
                raise RuntimeError("GeoWrapper::setPointCloud|point_cloud input and normals input should have the same number of points")
            nrm = _f32(nrm)
        # the C side copies the points (staging buffer); the wrapper keeps references for the getters
        # (getPointCloud / getNormals hand out copies) instead of two more 1.5 MB copies per frame
        p = np.ascontiguousarray(pts[:, :3])
        self._points = p
        self._normals = np.ascontiguousarray(nrm[:, :3]) if nrm is not None else None
        check(self._lib.mrh_set_points(self._h, p.ctypes.data, p.shape[0], self._normals.ctypes.data if nrm is not None else None))

    def getPointCloud(self):
        return np.array(self._points, copy=True)

    def getNormals(self):
        if self._normals is None:
            return np.zeros((self._points.shape[0], 3), np.float32)
        return np.array(self._normals, copy=True)

    # ---- the hot path ----
    def compute(self):
        check(self._lib.mrh_compute(self._h))

    def computeBegin(self):
        """Sharded maps: compute() up to the point where a starve frame needs the z-buffer min-reduced
        over the ranks. Returns True when that reduction is due before computeEnd()."""
        need = C.c_int()
        check(self._lib.mrh_compute_begin(self._h, C.byref(need)))
        return bool(need.value)

    def computeEnd(self):
        check(self._lib.mrh_compute_end(self._h))

    def zbufTensor(self):
        """The starve z-buffer as a torch int64 CUDA tensor (a view of the library's memory)."""
        import torch

        ptr, n = C.c_void_p(), C.c_size_t()
        check(self._lib.mrh_get_zbuf(self._h, C.byref(ptr), C.byref(n)))

        class _View:
            __cuda_array_interface__ = {"shape": (n.value,), "typestr": "<i8", "data": (ptr.value, False), "version": 2}

        return torch.as_tensor(_View(), device=torch.device("cuda", self._device))

    def synchronize(self):
        check(self._lib.mrh_synchronize(self._h))

    def stream(self, centre, radius):
        """Streamer::stream (streamer.cpp:337-355): page blocks farther than `radius` from `centre` out to
        the host store and the stored blocks around `centre` back in. compute() does this by itself
        when the pool runs low (field StreamThreshold)."""
        c = _f32(centre).reshape(3)
        check(self._lib.mrh_stream(self._h, c.ctypes.data_as(_capi._fp), float(radius)))

    def streamAllOut(self):
        check(self._lib.mrh_stream_all_out(self._h))

    def extractMesh(self, filename, force_generic=False):
        check(self._lib.mrh_extract_mesh_ex(self._h, None if filename is None else str(filename).encode(), int(force_generic)))

    def _mesh(self):
        v = C.POINTER(C.c_double)()
        f = C.POINTER(C.c_int32)()
        c = C.POINTER(C.c_double)()
        nv = C.c_size_t()
        nf = C.c_size_t()
        check(self._lib.mrh_get_mesh(self._h, C.byref(v), C.byref(f), C.byref(c), C.byref(nv), C.byref(nf)))
        V = np.ctypeslib.as_array(v, (nv.value, 3)).copy() if nv.value else np.zeros((0, 3))
        F = np.ctypeslib.as_array(f, (nf.value, 3)).copy() if nf.value else np.zeros((0, 3), np.int32)
        Cc = np.ctypeslib.as_array(c, (nv.value, 3)).copy() if nv.value else np.zeros((0, 3))
        return V, F, Cc

    def getVertices(self):
        return self._mesh()[0]

    def getFaces(self):
        return self._mesh()[1]

    def getColors(self):
        return self._mesh()[2]

    def getTriangles(self):
        """Raw triangle soup of the last extractMesh as float32 [T, 3, 6] (position, colour)."""
        t = C.POINTER(C.c_float)()
        n = C.c_size_t()
        check(self._lib.mrh_get_triangles(self._h, C.byref(t), C.byref(n)))
        if not n.value:
            return np.zeros((0, 3, 6), np.float32)
        return np.ctypeslib.as_array(t, (n.value, 3, 6)).copy()

    def serializeData(self, filename_hash="./data/hash_points.ply", filename_voxel="./data/voxel_points.ply"):
        check(self._lib.mrh_serialize_data(self._h, str(filename_hash).encode(), str(filename_voxel).encode()))

    def serializeGrid(self, filename="./data/grid.bin"):
        """geowrapper.cpp:567-569: the host chunk grid as a checkpoint in the reference's file format."""
        check(self._lib.mrh_serialize_grid(self._h, str(filename).encode()))

    def deserializeGrid(self, filename="./data/grid.bin"):
        """geowrapper.cpp:571-573: chunks of the file replace the same chunks of the host store."""
        check(self._lib.mrh_deserialize_grid(self._h, str(filename).encode()))

    def clearBuffers(self):
        check(self._lib.mrh_clear_buffers(self._h))

    def GSSavePointCloud(self, folder):
        print("GeoWrapper::GSSavePointCloud | GS container not initialized")

    def GSFinalOpt(self):
        pass

    # ---- additions used by tests / bench ----
    def getStats(self):
        s = _capi.Stats()
        check(self._lib.mrh_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def resetStats(self):
        check(self._lib.mrh_reset_stats(self._h))

    def setIngestMode(self, mode):
        """0: setters copy (default, reference semantics); 1 / 2: page-locked inputs are read by DMA
        (include/mrhash_b200.h, mrh_set_ingest_mode, for the lifetime contract of mode 2)."""
        check(self._lib.mrh_set_ingest_mode(self._h, int(mode)))

    def setStatsPipeline(self, enabled):
        check(self._lib.mrh_set_stats_pipeline(self._h, 1 if enabled else 0))

    def getStatsPipelined(self, which=1):
        """Counters after the frame `which` compute() calls before the last one (0 .. 3; waits for that frame only)."""
        s = _capi.Stats()
        check(self._lib.mrh_get_stats_pipelined(self._h, int(which), C.byref(s)))
        return s.as_dict()

    def lastComputeMs(self):
        ms = C.c_float()
        check(self._lib.mrh_last_compute_ms(self._h, C.byref(ms)))
        return ms.value

    def launchCount(self):
        n = C.c_uint64()
        check(self._lib.mrh_get_launch_count(self._h, C.byref(n)))
        return n.value

    def setProfiling(self, enabled):
        check(self._lib.mrh_set_profiling(self._h, int(bool(enabled))))

    def kernelTimes(self):
        """({slot: total ms}, {slot: launches}) accumulated since setProfiling(True)."""
        ms = (C.c_double * 8)()
        n = (C.c_uint64 * 8)()
        check(self._lib.mrh_get_kernel_times(self._h, ms, n))
        return list(ms), list(n)

    def cudaStream(self):
        s = C.c_void_p()
        check(self._lib.mrh_get_stream(self._h, C.byref(s)))
        return s.value

    def storeSize(self):
        n = C.c_size_t()
        check(self._lib.mrh_store_size(self._h, C.byref(n)))
        return n.value

    def storeDump(self):
        """(entries [n,5] int32, voxels [n,512]) of the host store, in store order."""
        n = C.c_size_t()
        check(self._lib.mrh_store_read(self._h, None, None, 0, C.byref(n)))
        entries = np.zeros((n.value, 5), np.int32)
        voxels = np.zeros((n.value, 512), VOXEL_DTYPE)
        if n.value:
            check(self._lib.mrh_store_read(self._h, entries.ctypes.data, voxels.ctypes.data, n.value, C.byref(n)))
        return entries, voxels

    def storeAppend(self, entries, voxels):
        """Add blocks (as dumpState returns them, e.g. another shard's) to the host store."""
        e = np.ascontiguousarray(entries, np.int32)
        v = np.ascontiguousarray(voxels)
        assert e.ndim == 2 and e.shape[1] == 5 and v.shape == (len(e), 512) and v.dtype == VOXEL_DTYPE
        check(self._lib.mrh_store_append(self._h, e.ctypes.data, v.ctypes.data, len(e)))

    def setShard(self, shard_rank, shard_world):
        check(self._lib.mrh_set_shard(self._h, int(shard_rank), int(shard_world)))
        self._shard_world = int(shard_world)

    def shardWorld(self):
        return self._shard_world

    # ---- sharded meshing: boundary exchange (include/mrhash_b200.h, csrc/mrh_halo.cu). The tensors
    # are torch CUDA tensors on this handle's device; torch is only the buffer / NCCL plumbing. ----
    def haloRecordBytes(self, full):
        return int(self._lib.mrh_halo_record_bytes(int(bool(full))))

    def haloRequests(self):
        """int32 [n,3] keys of the blocks next to owned blocks that another rank owns (unique)."""
        import torch

        dev = torch.device("cuda", self._device)
        cap = max(1024, 8 * self.getStats()["live_blocks"])
        while True:
            keys = torch.empty((cap, 3), dtype=torch.int32, device=dev)
            n = C.c_size_t()
            check(self._lib.mrh_halo_requests(self._h, keys.data_ptr(), cap, C.byref(n)))
            if n.value <= cap:
                return keys[: n.value]
            cap = n.value

    def haloPack(self, keys, full):
        """Owner side: uint8 [n, haloRecordBytes(full)] records for int32 [n,3] keys."""
        import torch

        keys = keys.contiguous()
        out = torch.empty((len(keys), self.haloRecordBytes(full)), dtype=torch.uint8, device=keys.device)
        check(self._lib.mrh_halo_pack(self._h, keys.data_ptr(), len(keys), int(bool(full)), out.data_ptr()))
        return out

    def haloInsert(self, keys, records, full):
        keys, records = keys.contiguous(), records.contiguous()
        assert len(keys) == len(records)
        check(self._lib.mrh_halo_insert(self._h, keys.data_ptr(), records.data_ptr(), len(keys), int(bool(full))))

    def haloClear(self):
        check(self._lib.mrh_halo_clear(self._h))

    def meshLocal(self):
        """Marching cubes over the owned blocks in place; returns the soup as a torch CUDA tensor [T,3,6]."""
        import torch

        n = C.c_size_t()
        check(self._lib.mrh_mesh_local(self._h, C.byref(n)))
        soup = torch.empty((n.value, 3, 6), dtype=torch.float32, device=torch.device("cuda", self._device))
        check(self._lib.mrh_copy_triangles_device(self._h, soup.data_ptr(), n.value))
        return soup

    def weldSoup(self, soup, filename=None):
        """processTriangles + PLY over a soup tensor [T,3,6] on this handle's device; mesh via getVertices/..."""
        soup = soup.contiguous()
        check(self._lib.mrh_weld_device_soup(self._h, soup.data_ptr(), len(soup), filename.encode() if filename else None))

    def hasLowResolutionBlocks(self):
        """True when the map may hold resolution-1 blocks (then whole blocks are exchanged, not shells)."""
        self.getStats()  # refreshes the counters the field reads
        return bool(self._get("LowResolutionBlocks"))

    def dumpState(self):
        """(entries [n,5] int32 = x,y,z,resolution,ptr sorted by key; voxels structured [n,512])."""
        n = C.c_size_t()
        check(self._lib.mrh_dump_state(self._h, None, None, 0, C.byref(n)))
        entries = np.zeros((n.value, 5), np.int32)
        voxels = np.zeros((n.value, 512), VOXEL_DTYPE)
        if n.value:
            m = C.c_size_t()
            check(self._lib.mrh_dump_state(self._h, entries.ctypes.data, voxels.ctypes.data, n.value, C.byref(m)))
            entries, voxels = entries[: m.value], voxels[: m.value]
        return entries, voxels


def grid_write(path, entries, voxels, virtual_voxel_size, voxel_extents=1.0):
    """serializeGrid's file format without a handle (no GPU needed): entries [n,5] int32, voxels [n,512]."""
    e = np.ascontiguousarray(entries, np.int32)
    v = np.ascontiguousarray(voxels)
    assert e.ndim == 2 and e.shape[1] == 5 and v.shape == (len(e), 512) and v.dtype == VOXEL_DTYPE
    check(_capi.lib().mrh_grid_write(str(path).encode(), e.ctypes.data, v.ctypes.data, len(e), float(virtual_voxel_size), float(voxel_extents)))


def grid_read(path):
    """(entries [n,5] int32, voxels [n,512]) of a serializeGrid file, in file order."""
    lib = _capi.lib()
    n = C.c_size_t()
    check(lib.mrh_grid_read(str(path).encode(), None, None, 0, C.byref(n)))
    entries = np.zeros((n.value, 5), np.int32)
    voxels = np.zeros((n.value, 512), VOXEL_DTYPE)
    if n.value:
        check(lib.mrh_grid_read(str(path).encode(), entries.ctypes.data, voxels.ctypes.data, n.value, C.byref(n)))
    return entries, voxels


# voxel_hash_utils.cuh:8-22
VOXEL_DTYPE = np.dtype([("sdf", "<f4"), ("sum_squared", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("weight", "u1")])

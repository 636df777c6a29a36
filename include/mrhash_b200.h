/* mrhash_b200.h — C ABI of libmrhash_b200.so, the B200-native drop-in for mrhash's
 * GeoWrapper::compute() hot path (depth-ray block allocation into the spatial hash, TSDF fusion,
 * garbage collection, variance-adaptive re-allocation, LiDAR point clouds, radius paging) plus
 * extractMesh(), serializeData(), serializeGrid() and the multi-GPU support calls (sharded starve
 * frames, boundary exchange for meshing).
 *
 * The reference has no C ABI: its boundary is the nanobind class `pygeowrapper.GeoWrapper`
 * (/root/reference/mrhash/src/sdf/pybind/pygeowrapper.cpp:12-84) over the C++ class
 * pygeowrapper::GeoWrapper (/root/reference/mrhash/src/sdf/geowrapper.h:18-260). Every entry
 * point below names the reference member it replaces; INTEGRATION.md shows the binding a
 * maintainer would add on the reference side.
 *
 * Conventions: every call returns 0 on success, non-zero on failure with a message available from
 * mrh_last_error() (thread-local). One handle = one CUDA device; a handle is not thread-safe
 * (like the reference). compute() is asynchronous on the handle's stream; every call that returns
 * data to the host synchronises first.
 */
#ifndef MRHASH_B200_H
#define MRHASH_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRH_ABI_VERSION 1

typedef struct mrh_map mrh_map;

/* Constructor arguments of GeoWrapper (pygeowrapper.cpp:14-29, geowrapper.cpp:9-81) plus explicit
 * sizing. A sizing field left at 0 is derived from cudaMemGetInfo with the reference's ratios
 * (params.h:33-37, geowrapper.cpp:37-54) in 64-bit arithmetic. */
typedef struct {
  float sdf_truncation;
  float sdf_truncation_scale;
  int32_t integration_weight_sample;
  float virtual_voxel_size;
  int32_t n_frames_invalidate_voxels;
  int32_t voxel_extents_scale;
  int32_t viewer_active; /* accepted, ignored: no viewer thread */
  float marching_cubes_threshold;
  int32_t min_weight_threshold;
  float min_depth;
  float max_depth;
  float sdf_var_threshold;
  float vertices_merging_threshold;
  int32_t projective_sdf;
  /* sizing (0 = auto) */
  uint64_t num_sdf_blocks;
  uint64_t hash_num_buckets;
  uint64_t max_num_triangles;
  /* placement */
  int32_t device;      /* CUDA device ordinal, -1 = current */
  int32_t shard_rank;  /* this handle owns hash buckets [rank*nb/world, (rank+1)*nb/world) */
  int32_t shard_world; /* 0 or 1 = unsharded */
  int32_t reserved;
} mrh_params;

/* per-run totals since creation / mrh_reset_stats (the work-unit accounting of BASELINE.md §4) */
typedef struct {
  uint64_t frames;
  uint64_t rays_valid;
  uint64_t blocks_new;     /* B_new */
  uint64_t blocks_visible; /* B_vis summed over frames */
  uint64_t voxels_updated; /* V_upd */
  uint64_t blocks_freed;
  uint64_t blocks_realloc;
  uint64_t dropped_heap;   /* "mem size exceed, not inserting hash entry" events */
  uint64_t dropped_table;
  uint64_t live_blocks;    /* currently allocated */
  int64_t heap_free;       /* getHeapHighFreeCount() (voxel_data_structures.cpp:148-153) */
  int64_t heap_low_free;
  uint64_t dropped_updates; /* point-cloud path: update records beyond the staging capacity (0 in a healthy run) */
} mrh_stats;

/* record format of mrh_dump_state (test / parity only) */
typedef struct {
  int32_t x, y, z, resolution, ptr;
} mrh_dump_entry;

const char* mrh_last_error(void);
int mrh_abi_version(void);

/* GeoWrapper::GeoWrapper / ~GeoWrapper (geowrapper.cpp:9-84) */
int mrh_params_default(mrh_params* p);
int mrh_create(const mrh_params* p, mrh_map** out);
int mrh_destroy(mrh_map* m);

/* GeoWrapper::setCamera (geowrapper.cpp:98-116); camera_model 0 = pinhole, 1 = spherical */
int mrh_set_camera(mrh_map* m, float fx, float fy, float cx, float cy, int rows, int cols, float min_depth, float max_depth, int camera_model);
/* GeoWrapper::setCurrPose (geowrapper.cpp:86-92): translation + quaternion (x, y, z, w) */
int mrh_set_pose(mrh_map* m, const float translation[3], const float quaternion_xyzw[4]);
/* same state, given directly as a row-major 4x4 cam_in_world (what compute() hands to Camera::setCamInWorld) */
int mrh_set_pose_matrix(mrh_map* m, const float cam_in_world[16]);
/* GeoWrapper::getCurrPose (geowrapper.h:94) */
int mrh_get_pose_matrix(mrh_map* m, float out[16]);
/* GeoWrapper::setCameraInLidar (geowrapper.cpp:94-96) */
int mrh_set_camera_in_lidar(mrh_map* m, const float T[16]);

/* Frame setters. By default every host buffer is copied before the call returns (the reference's
 * setters copy element-wise, geowrapper.cpp:261-273): the caller may reuse it at once.
 * mrh_set_ingest_mode lets a caller that holds page-locked (cudaHostAlloc / cudaHostRegister) frames
 * skip that copy; pageable buffers are always staged.
 *   0  (default) staged copy in the setter;
 *   1  page-locked memory is read by DMA, the transfer completes inside the next mrh_compute(): the
 *      buffer must not change between the setter and the return of mrh_compute();
 *   2  page-locked memory is read by DMA and mrh_compute() does not wait for it (the transfer of frame
 *      k+1 then overlaps the kernels of frame k): a buffer must not change until the SECOND next call
 *      of the same setter has returned, or until mrh_synchronize(). */
int mrh_set_ingest_mode(mrh_map* m, int mode);
/* GeoWrapper::setDepthImage (geowrapper.cpp:246-274): host float32 [rows, cols], copied */
int mrh_set_depth(mrh_map* m, const float* depth, int rows, int cols);
/* GeoWrapper::setRGBImage (geowrapper.cpp:276-298): host uint8 [rows, cols, 3], copied */
int mrh_set_rgb(mrh_map* m, const uint8_t* rgb, int rows, int cols);
/* same, for callers that hold float32 colour (apps/utils/depth_reader.py:83-93); cast to uint8 like nanobind's implicit conversion */
int mrh_set_rgb_f32(mrh_map* m, const float* rgb, int rows, int cols);
/* inputs already resident in device memory (no copy; the pointers must stay valid until the next compute() completes) */
int mrh_set_depth_device(mrh_map* m, const float* d_depth, int rows, int cols);
int mrh_set_rgb_device(mrh_map* m, const uint8_t* d_rgb, int rows, int cols);
/* GeoWrapper::setPointCloud (geowrapper.cpp:345-405, 469-505): host float32 [n, 3]; normals may be NULL */
int mrh_set_points(mrh_map* m, const float* points, size_t n, const float* normals_or_null);

/* GeoWrapper::compute (geowrapper.cpp:118-148) */
int mrh_compute(mrh_map* m);
/* Sharded maps (shard_world > 1) only. Every n_frames_invalidate_voxels-th frame the reference decrements
 * the weight of the front-most voxel of each pixel (starveVoxelsKernel, voxel_data_structures.cu:1597-1671),
 * found with a per-pixel atomicMin z-buffer; a shard sees only its own voxels, so on those frames
 * mrh_compute_begin stops after the z-buffer pass and sets *needs_zbuf_reduce: the caller min-reduces
 * the buffer (mrh_get_zbuf: int64 cells, device memory) over all ranks - the one collective of the
 * integration path - and calls mrh_compute_end. On all other frames begin runs the whole frame.
 * mrh_compute == begin + end without the reduction. */
int mrh_compute_begin(mrh_map* m, int* needs_zbuf_reduce);
int mrh_compute_end(mrh_map* m);
int mrh_get_zbuf(mrh_map* m, void** d_zbuf, size_t* n_cells);
/* Streamer::stream (streamer.cpp:337-355): blocks whose origin is at least `radius` away from `centre`
 * leave the device for the host store, stored blocks whose 1 m chunk lies inside the sphere come back.
 * mrh_compute calls it by itself with (camera position, max_depth) when no more than
 * StreamThreshold (field, default 0.15 = params.h:28; 0 = never) of the pool is free, as
 * GeoWrapper::compute does (geowrapper.cpp:137-138); the free count it looks at is the one of the last
 * completed frame. Fields StreamEvents, LastStreamOutBlocks, LastStreamInBlocks, StreamDuplicates report it. */
int mrh_stream(mrh_map* m, const float centre[3], float radius);
/* wait for all queued work of this handle */
int mrh_synchronize(mrh_map* m);

/* GeoWrapper::streamAllOut (geowrapper.cpp:559-561) */
int mrh_stream_all_out(mrh_map* m);
/* Multi-GPU meshing support. mrh_store_append adds blocks (records in mrh_dump_entry layout + 512
 * reference Voxel structs each, as mrh_dump_state returns them - e.g. another shard's dump) to this
 * handle's host store; mrh_set_shard changes the hash-bucket range the handle accepts (world <= 1 =
 * everything), so that one rank can stream the gathered shards in and mesh the whole map. */
int mrh_store_append(mrh_map* m, const mrh_dump_entry* entries, const void* voxels, size_t n);
int mrh_set_shard(mrh_map* m, int shard_rank, int shard_world);
/* number of blocks currently held by the host store (the reference's Streamer::grid_) */
int mrh_store_size(mrh_map* m, size_t* n_blocks);
/* test / tooling: copy of the host store in store order (same layouts as mrh_dump_state; duplicates of a key, if any, included) */
int mrh_store_read(mrh_map* m, mrh_dump_entry* entries, void* voxels, size_t max_entries, size_t* n_out);
/* GeoWrapper::extractMesh (geowrapper.cpp:150-230): marching cubes + host weld + ASCII PLY (path may be NULL: no file) */
int mrh_extract_mesh(mrh_map* m, const char* path_or_null);
/* same; force_generic != 0 sends every block through the per-read hash sampler (test / validation of the shared-memory path) */
int mrh_extract_mesh_ex(mrh_map* m, const char* path_or_null, int force_generic);
/* GeoWrapper::getVertices / getFaces / getColors (geowrapper.h:91-93); pointers stay valid until the next extract */
int mrh_get_mesh(mrh_map* m, const double** vertices, const int32_t** faces, const double** colors, size_t* n_vertices, size_t* n_faces);
/* the ASCII PLY writer of extractMesh (geowrapper.cpp:194-229) on caller-supplied arrays (no handle, no device:
 * tooling and the CPU test of the file format) */
int mrh_write_mesh_ply(const char* path, const double* vertices, const double* colors, const int32_t* faces, size_t n_vertices, size_t n_faces);
/* raw triangle soup of the last extract (72-byte Triangle records, voxel_hash_utils.cuh:46-64) */
int mrh_get_triangles(mrh_map* m, const float** triangles, size_t* n_triangles);
/* ---- sharded meshing (no reference counterpart: the reference is single-GPU; marching_cubes.cu:72-214
 * reads the 26 neighbour blocks, which under the hash-bucket partition live on other ranks).
 * All pointers named d_* are DEVICE pointers on the handle's device (buffers the caller exchanges
 * with NCCL). Sequence per rank: mrh_halo_requests -> all-to-all of the keys to their owners ->
 * mrh_halo_pack on the owner -> all-to-all of the records back -> mrh_halo_insert ->
 * mrh_mesh_local -> mrh_halo_clear. */
/* bytes of one exchanged record: 16-byte header (int32 resolution, -1 = key not held) + (sdf, rgbw)
 * pairs of the block's one-voxel shell (296 voxels; full != 0: all 512, needed with resolution-1 blocks) */
size_t mrh_halo_record_bytes(int full);
/* unique keys (x,y,z int32) of the neighbours of owned blocks that another rank owns; *n may exceed
 * cap, in which case only cap keys were written and the call must be repeated with a larger buffer */
int mrh_halo_requests(mrh_map* m, int32_t* d_keys_xyz, size_t cap, size_t* n);
/* owner side: one record per requested key */
int mrh_halo_pack(mrh_map* m, const int32_t* d_keys_xyz, size_t n, int full, void* d_records);
/* requester side: received records become ghost blocks (sampled by the mesher, never meshed themselves) */
int mrh_halo_insert(mrh_map* m, const int32_t* d_keys_xyz, const void* d_records, size_t n, int full);
/* removes the ghost blocks again; the map is exactly as before mrh_halo_insert */
int mrh_halo_clear(mrh_map* m);
/* marching cubes over the owned blocks where they are (no stream-out); the soup stays on the device */
int mrh_mesh_local(mrh_map* m, size_t* n_triangles);
/* copies the soup of the last extraction into a device buffer of cap_triangles x 18 floats */
int mrh_copy_triangles_device(mrh_map* m, float* d_dst, size_t cap_triangles);
/* MeshExtractor::processTriangles (mesh_extractor.cpp:9-76) + the PLY writer (geowrapper.cpp:194-229) over a
 * soup in device memory (e.g. the soups of all ranks gathered with NCCL); result via mrh_get_mesh */
int mrh_weld_device_soup(mrh_map* m, const float* d_soup, size_t n_triangles, const char* path_or_null);

/* GeoWrapper::serializeData (geowrapper.cpp:563-565) */
int mrh_serialize_data(mrh_map* m, const char* hash_path, const char* voxel_path);
/* GeoWrapper::serializeGrid / deserializeGrid (geowrapper.cpp:567-573): the host store (what streamAllOut
 * leaves behind) as a checkpoint file in the reference's own format - Serializer<Voxel> (serializer.h:16-75):
 * per 1 m chunk `u64 size | int32 chunk[3] | cista::offset bytes of ChunkDesc<Voxel>` - so that files move
 * between the reference and this library in both directions. deserialize replaces the chunks the file
 * names and keeps the rest of the store. */
int mrh_serialize_grid(mrh_map* m, const char* path);
int mrh_deserialize_grid(mrh_map* m, const char* path);
/* the same format without a handle (no device needed): records as in mrh_dump_state, 512 x 12-byte
 * voxels per record (a resolution-1 record uses the first 64) */
int mrh_grid_write(const char* path, const mrh_dump_entry* entries, const void* voxels, size_t n, float virtual_voxel_size, float voxel_extents);
int mrh_grid_read(const char* path, mrh_dump_entry* entries, void* voxels, size_t max_entries, size_t* n_out);
/* test hook: the number formatter of the ASCII PLY writer (`ostream << double` of geowrapper.cpp:194-229,
 * i.e. printf("%g")); writes at most 32 bytes, no terminator, returns the length */
size_t mrh_format_g6(double v, char* dst);
/* GeoWrapper::clearBuffers (geowrapper.cpp:552-557) */
int mrh_clear_buffers(mrh_map* m);

/* getters / setters of the sizing fields (geowrapper.h:79-109); names as in pygeowrapper.cpp:31-61 without get/set */
int mrh_get_field(mrh_map* m, const char* name, double* out);
int mrh_set_field(mrh_map* m, const char* name, double value);

int mrh_get_stats(mrh_map* m, mrh_stats* out);
/* Statistics without a stall per frame: once enabled, every mrh_compute() ends with a one-CTA kernel that
 * writes the counters into one of four page-locked (mapped) slots; mrh_get_stats_pipelined(which = k) returns
 * the state after the frame k compute() calls before the last one (k = 0 .. 3; it waits for that frame
 * only). A caller that reads which = 2 after every compute() sees every frame's result two frames late
 * and never waits for a kernel it has just queued, so the next frame's upload is submitted while the
 * previous frame is still on the device. */
int mrh_set_stats_pipeline(mrh_map* m, int enabled);
int mrh_get_stats_pipelined(mrh_map* m, int which, mrh_stats* out);
int mrh_reset_stats(mrh_map* m);
/* device time of the last compute() in ms (CUDAProfiler::CUDAEvent window, voxel_data_structures.cpp:94-109); synchronises */
int mrh_last_compute_ms(mrh_map* m, float* ms);
/* Per-kernel device timing for the roofline pass of bench.py: when enabled, compute() records CUDA
 * events between its kernels and waits for the frame, accumulating ms per kernel slot
 * (RGB-D: 0 = allocate, 1 = visibility, 2 = integrate[+GC]). Enabling resets the accumulators. */
int mrh_set_profiling(mrh_map* m, int enabled);
int mrh_get_kernel_times(mrh_map* m, double ms_out[8], uint64_t launches_out[8]);
/* the CUDA stream compute() runs on (cudaStream_t), for event timing by the caller */
int mrh_get_stream(mrh_map* m, void** stream);
/* number of kernel launches issued by this handle so far */
int mrh_get_launch_count(mrh_map* m, uint64_t* n);

/* Test hook for the shared-reciprocal divisions of the frame kernel (csrc/mrh_div.cuh, which replace the
 * `a / b` of voxel_hash_utils.cuh:75-151,169-181, camera.cuh:131-160 and voxel_data_structures.cu:803-822):
 * compares them with IEEE division on the device for every one of the 2^32 numerator bit patterns of
 * `divisor` (skipped when divisor <= 0) and for n_random (numerator, divisor) pairs; *mismatches = number
 * of quotients whose bits differ. */
int mrh_selftest_div(int device, float divisor, uint64_t n_random, uint64_t seed, uint64_t* mismatches);
/* radius (in voxels) inside which the integer voxel -> block shortcut of the ray walk was verified to
 * equal the reference's metric arithmetic for the handle's current voxel size (0: shortcut unused) */
int mrh_get_block_shortcut_radius(mrh_map* m, int* radius);

/* Parity dump: live block records sorted by (x, y, z) and their voxels as 12-byte reference Voxel
 * structs {f32 sdf, f32 sum_squared, u8 r, g, b, weight} (512 per record; resolution-1 records use
 * the first 64). Returns the number of live blocks in *n_out; fills at most max_entries. */
int mrh_dump_state(mrh_map* m, mrh_dump_entry* entries, void* voxels, size_t max_entries, size_t* n_out);

#ifdef __cplusplus
}
#endif
#endif

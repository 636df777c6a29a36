/* mrh_oracle.h — CPU restatement of mrhash's per-frame TSDF integration hot path.
 *
 * TEST INFRASTRUCTURE ONLY. This is the parity oracle: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it. The product
 * (mrhash_b200/, libmrhash_b200.so) never links, loads or calls anything in oracle/.
 *
 * Parity pinning: the reference (rvp-group/mrhash) is CUDA-only and its tests hold no golden
 * values for TSDF contents, so the oracle is pinned two ways:
 *   1. against every invariant / known-answer the reference's own tests state
 *      (tests/test_oracle_reference_invariants.py cites them), and
 *   2. against the UNMODIFIED reference kernels compiled for sm_100a (oracle/_ref/,
 *      built by oracle/Makefile) run on the same synthetic inputs on the B200 box; small outputs
 *      of those runs are committed as fixtures under tests/golden/.
 * Known deviation: normalize() uses MUFU.RSQ on the GPU (cuda_math.cuh:1075-1078); the CPU uses a
 * correctly rounded 1/sqrt, so DDA tie decisions may differ in the last ulp (comparator reports).
 *
 * All reference citations are relative to /root/reference/mrhash/src/sdf/.
 */
#ifndef MRH_ORACLE_H
#define MRH_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* voxel_hash_utils.cuh:8-22 */
typedef struct {
  float sdf;
  float sum_squared;
  uint8_t r, g, b;
  uint8_t weight;
} orc_voxel;

/* voxel_hash_utils.cuh:27-38 */
typedef struct {
  int32_t x, y, z;
  uint32_t offset;
  int32_t ptr;
  int32_t resolution;
} orc_entry;

/* record layout shared with oracle/ref_harness/harness.cu (ref_dump_entry) */
typedef struct {
  int32_t x, y, z, resolution, ptr;
} orc_dump_entry;

/* voxel_hash_utils.cuh:46-64 — 3 x {float3 p, float3 c} = 72 B */
typedef struct {
  float v[3][6];
} orc_triangle;

typedef struct {
  uint64_t rays_valid;      /* pixels / points that produced a DDA */
  uint64_t blocks_new;      /* B_new: hash entries created this frame */
  uint64_t blocks_visible;  /* B_vis: compact list length used by integrate */
  uint64_t voxels_updated;  /* V_upd: voxels that passed every test and were fused */
  uint64_t blocks_freed;    /* entries removed by garbage collection */
  uint64_t blocks_realloc;  /* variance path: blocks re-allocated at resolution 1 */
} orc_frame_stats;

typedef struct orc_map orc_map;

orc_map* orc_create(uint32_t num_sdf_blocks,
                    uint32_t hash_num_buckets,
                    float sdf_truncation,
                    float sdf_truncation_scale,
                    int integration_weight_sample,
                    float virtual_voxel_size,
                    int n_frames_invalidate_voxels,
                    int voxel_extents_scale,
                    float marching_cubes_threshold,
                    int min_weight_threshold,
                    float sdf_var_threshold,
                    int projective_sdf);
void orc_destroy(orc_map* m);
void orc_set_threads(orc_map* m, int n_threads);
void orc_set_camera(orc_map* m, float fx, float fy, float cx, float cy, int rows, int cols, float min_depth, float max_depth, int model);
void orc_compute_rgbd(orc_map* m, const float* pose16, const float* depth, const uint8_t* rgb, int rows, int cols);
void orc_compute_points(orc_map* m, const float* pose16, const float* points, const float* normals, int n);
void orc_last_stats(const orc_map* m, orc_frame_stats* out);
int orc_heap_high_free(const orc_map* m);
int orc_heap_low_free(const orc_map* m);
uint32_t orc_dump(const orc_map* m, orc_dump_entry* entries, orc_voxel* voxels, uint32_t max_entries);
uint32_t orc_extract_triangles(orc_map* m, orc_triangle* out, uint32_t max_out);

/* raw views for the invariant tests (tests/test_hash_utils.cu:306-376) */
const orc_entry* orc_table(const orc_map* m, uint32_t* total_size);
const uint32_t* orc_heap_high(const orc_map* m, int* counter);
uint32_t orc_overflow_events(const orc_map* m);

/* scalar helpers exported for known-answer tests */
uint32_t orc_calculate_hash(int x, int y, int z, uint32_t hash_num_buckets);
void orc_world_to_voxel(float voxel_size, const float p[3], int out[3]);
void orc_voxel_to_block(const int v[3], float voxel_size, const float extents[3], int out[3]);
uint32_t orc_voxel_to_block_index(const int v[3], int block_size);
void orc_delinearize(uint32_t index, int block_size, uint32_t out[3]);
void orc_inverse_projection(const orc_map* m, uint32_t row, uint32_t col, float d, float out[3]);
int orc_project_point(const orc_map* m, const float pc[3], int* row, int* col);
void orc_quat_to_matrix(const float t[3], const float q_xyzw[4], float out16[16]);

#ifdef __cplusplus
}
#endif
#endif

// R-CUDA harness: drives the UNMODIFIED reference translation units (compiled in place from
// /root/reference/mrhash/src/sdf by oracle/Makefile into oracle/_ref/) through a small C ABI
// so that Python tests / bench.py can replay the same synthetic streams through the
// reference's own kernels on the GPU box.
//
// Test infrastructure only. Nothing under mrhash_b200/ may link or load this.
//
// It mirrors what the reference's orchestrator does per frame, without nanobind / OpenCV /
// the Streamer (none of which build here):
//   GeoWrapper::GeoWrapper   geowrapper.cpp:56-80   (container construction, voxel_extents)
//   GeoWrapper::setCamera    geowrapper.cpp:98-116  (Camera + setIntegrationDistance)
//   GeoWrapper::compute      geowrapper.cpp:118-148 (setCamInWorld, toDevice, computeCloud, integrate)
//   MeshExtractor::extractMesh mesh_extractor.cpp:95-98 (flatAndReduceHashTable + extractIsoSurface)
// The container is sized explicitly (the reference's auto-sizing overflows 32-bit arithmetic
// on a 180 GB part).
#include <Eigen/Core>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <vector>

#include "camera.cuh"
#include "marching_cubes.cuh"
#include "streamer.cuh"
#include "voxel_data_structures.cuh"

using namespace cupanutils::cugeoutils;

// MeshExtractor's host members (processTriangles, combine, removeDuplicate*) come from the reference's own
// mesh_extractor.cpp, compiled into this library by oracle/Makefile against the Eigen stand-in.

namespace {
  struct Ref {
    std::unique_ptr<GeometricVoxelContainer> container;
    std::unique_ptr<Camera> camera;
    std::unique_ptr<GeometricMarchingCubes> mc;
    std::unique_ptr<GeometricStreamer> streamer; // declared after the container: destroyed first
    CUDAMatrixf depth_img;
    CUDAMatrixuc3 rgb_img;
    CUDAVectorf3 point_cloud;
    CUDAVectorf3 eigenvectors;
    CUDAVectorf weights;
    int n_frames_invalidate = 0;
    float last_integrate_ms = 0.f;
    float mc_threshold      = 0.f;
    float vertices_merging  = 0.f;
    uint32_t max_triangles  = 0;
  };
} // namespace

extern "C" {

struct ref_dump_entry {
  int32_t x, y, z, resolution, ptr;
};

void* ref_create(uint32_t num_sdf_blocks,
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
                 int projective_sdf,
                 uint32_t max_num_triangles) {
  Ref* r       = new Ref();
  r->container = std::make_unique<GeometricVoxelContainer>(num_sdf_blocks,
                                                           hash_num_buckets,
                                                           0.f,
                                                           sdf_truncation,
                                                           sdf_truncation_scale,
                                                           virtual_voxel_size,
                                                           integration_weight_sample,
                                                           (uchar) min_weight_threshold,
                                                           sdf_var_threshold,
                                                           projective_sdf != 0,
                                                           false,
                                                           "/tmp/ref_memory_allocation.txt",
                                                           "ref_integration_profiler",
                                                           "ref_rendering_profiler");
  // Streamer::create does this in the reference (streamer.cpp:22)
  r->container->voxel_extents_ =
    make_float3((float) voxel_extents_scale, (float) voxel_extents_scale, (float) voxel_extents_scale);
  r->container->current_occupied_blocks_ = 0;
  r->container->updateFieldsDevice();
  r->n_frames_invalidate = n_frames_invalidate_voxels;
  r->mc_threshold        = marching_cubes_threshold;
  r->max_triangles       = max_num_triangles;
  if (max_num_triangles > 0)
    r->mc = std::make_unique<GeometricMarchingCubes>(marching_cubes_threshold, false, max_num_triangles, 0.f);
  return r;
}

void ref_destroy(void* h) {
  delete (Ref*) h;
}

void ref_set_camera(void* h, float fx, float fy, float cx, float cy, int rows, int cols, float min_depth, float max_depth, int model) {
  Ref* r = (Ref*) h;
  Eigen::Matrix3f K;
  K(0, 0) = fx;
  K(0, 2) = cx;
  K(1, 1) = fy;
  K(1, 2) = cy;
  K(2, 2) = 1.f;
  CUDAMat3 d_K(K);
  r->container->setIntegrationDistance(max_depth);
  r->camera = std::make_unique<Camera>(d_K, rows, cols, min_depth, max_depth, (CameraModel) model);
}

// pose: row-major 4x4 cam_in_world. depth: rows*cols f32. rgb: rows*cols*3 u8.
void ref_compute_rgbd(void* h, const float* pose, const float* depth, const uint8_t* rgb, int rows, int cols) {
  Ref* r = (Ref*) h;
  Eigen::Matrix4f T;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      T(i, j) = pose[i * 4 + j];
  r->camera->setCamInWorld(T);
  r->depth_img.resize(rows, cols);
  r->rgb_img.resize(rows, cols);
  for (int i = 0; i < rows * cols; ++i) {
    r->depth_img.at(i)   = depth[i];
    r->rgb_img.at(i).x   = rgb[3 * i];
    r->rgb_img.at(i).y   = rgb[3 * i + 1];
    r->rgb_img.at(i).z   = rgb[3 * i + 2];
  }
  CUDAMatrixf3 point_cloud_img;
  r->depth_img.toDevice();
  r->rgb_img.toDevice();
  r->camera->setDepthImage(r->depth_img);
  r->camera->computeCloud(point_cloud_img);
  r->container->integrate(point_cloud_img, r->rgb_img, *r->camera, r->n_frames_invalidate);
  r->last_integrate_ms = (float) r->container->integration_profiler_.elapsed_;
}

// points: n*3 f32 in the sensor frame; normals may be null (projective sdf ignores them)
void ref_compute_points(void* h, const float* pose, const float* points, const float* normals, int n) {
  Ref* r = (Ref*) h;
  Eigen::Matrix4f T;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      T(i, j) = pose[i * 4 + j];
  r->camera->setCamInWorld(T);
  r->point_cloud.resize(n, 1);
  r->eigenvectors.resize(3 * n, 1);
  r->weights.resize(n, 1);
  for (int i = 0; i < n; ++i) {
    r->point_cloud.at(i) = make_float3(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
    r->weights.at(i)     = 0.f;
    for (int k = 0; k < 3; ++k)
      r->eigenvectors.at(3 * i + k) = make_float3(0.f, 0.f, 0.f);
    if (normals)
      r->eigenvectors.at(3 * i) = make_float3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
  }
  r->point_cloud.toDevice();
  r->eigenvectors.toDevice();
  r->weights.toDevice();
  r->container->integrate(r->point_cloud, r->eigenvectors, r->weights, *r->camera, r->n_frames_invalidate);
  r->last_integrate_ms = (float) r->container->integration_profiler_.elapsed_;
}

float ref_last_integrate_ms(void* h) {
  return ((Ref*) h)->last_integrate_ms;
}
int ref_heap_high_free(void* h) {
  return ((Ref*) h)->container->getHeapHighFreeCount();
}
int ref_heap_low_free(void* h) {
  return ((Ref*) h)->container->getHeapLowFreeCount();
}
uint32_t ref_current_occupied_blocks(void* h) {
  return ((Ref*) h)->container->current_occupied_blocks_;
}

// Number of live hash entries (ptr != FREE_ENTRY) in the whole table.
// If entries != null, fills up to max_entries records and, if voxels != null, the 512 (or 64)
// voxels of each record at voxels[i*512 .. ] as raw 12-byte reference Voxel structs.
uint32_t ref_dump(void* h, ref_dump_entry* entries, uint8_t* voxels, uint32_t max_entries) {
  Ref* r              = (Ref*) h;
  const uint32_t size = r->container->total_size_;
  std::vector<HashEntry> table(size);
  CUDA_CHECK(cudaMemcpy(table.data(), r->container->d_hashTable_, sizeof(HashEntry) * size, cudaMemcpyDeviceToHost));
  uint32_t n = 0;
  for (uint32_t i = 0; i < size; ++i) {
    if (table[i].ptr == FREE_ENTRY)
      continue;
    if (entries && n < max_entries) {
      entries[n] = {table[i].pos.x, table[i].pos.y, table[i].pos.z, table[i].resolution, table[i].ptr};
      if (voxels) {
        const int nv = table[i].resolution == 0 ? 512 : 64;
        CUDA_CHECK(cudaMemcpy(voxels + (size_t) n * 512 * sizeof(Voxel),
                              r->container->d_SDFBlocks_ + table[i].ptr,
                              sizeof(Voxel) * nv,
                              cudaMemcpyDeviceToHost));
      }
    }
    ++n;
  }
  return n;
}

// Runs the reference marching cubes over everything in the table; returns the triangle count and
// copies min(count, max_out) 72-byte Triangle structs to out.
uint32_t ref_extract_triangles(void* h, float* out, uint32_t max_out) {
  Ref* r = (Ref*) h;
  if (!r->mc)
    return 0;
  r->mc->num_triangles_ = 0;
  r->container->flatAndReduceHashTable();
  r->mc->extractIsoSurface(*r->container);
  const uint32_t n = r->mc->num_triangles_;
  if (out) {
    const uint32_t m = n < max_out ? n : max_out;
    memcpy(out, r->mc->h_triangles_, (size_t) m * sizeof(Triangle));
  }
  return n;
}

// MeshExtractor::processTriangles (mesh_extractor.cpp:9-76), the reference's own host code, over a
// caller-supplied soup of 72-byte Triangle records: vertex weld (exact for eps == 0, floor(v / eps)
// buckets otherwise), degenerate-face and duplicate-face removal. merge != 0 appends to the mesh of
// the previous call first (merge_mesh_, the path GeoWrapper::extractMesh takes per region).
// Returns 0 and the sizes; ref_get_processed copies V (f64 x 3), F (i32 x 3), C (f64 x 3).
int ref_process_soup(void* h, const float* soup, uint32_t n_tri, float eps, int merge, uint32_t* n_vertices, uint32_t* n_faces) {
  Ref* r = (Ref*) h;
  if (!r->mc || n_tri > r->mc->max_num_triangles_mesh_)
    return 1;
  memcpy(r->mc->h_triangles_, soup, (size_t) n_tri * sizeof(Triangle));
  r->mc->num_triangles_              = n_tri;
  r->mc->vertices_merging_threshold_ = eps;
  r->mc->merge_mesh_                 = merge != 0;
  if (!merge) {
    r->mc->vertices_ = Eigen::MatrixXd();
    r->mc->faces_    = Eigen::MatrixXi();
    r->mc->colors_   = Eigen::MatrixXd();
  }
  r->mc->processTriangles();
  *n_vertices = (uint32_t) r->mc->getVertices().rows();
  *n_faces    = (uint32_t) r->mc->getFaces().rows();
  return 0;
}
void ref_get_processed(void* h, double* V, int32_t* F, double* C) {
  Ref* r = (Ref*) h;
  const Eigen::MatrixXd& v = r->mc->getVertices();
  const Eigen::MatrixXi& f = r->mc->getFaces();
  const Eigen::MatrixXd& c = r->mc->getColors();
  for (int i = 0; i < v.rows(); ++i)
    for (int j = 0; j < 3; ++j)
      V[3 * i + j] = v(i, j), C[3 * i + j] = c(i, j);
  for (int i = 0; i < f.rows(); ++i)
    for (int j = 0; j < 3; ++j)
      F[3 * i + j] = f(i, j);
}
// the three mesh utilities the reference's own tests exercise (tests/test_marching_cubes.cpp:12-271)
// on caller-supplied matrices; outputs are sized by the caller (n_v / n_f are upper bounds in, sizes out)
int ref_remove_duplicate_vertices(void* h, const double* V, uint32_t n_v, const int32_t* F, uint32_t n_f, double eps, double* V_out, uint32_t* n_v_out, int32_t* F_out, int32_t* map_out) {
  Ref* r = (Ref*) h;
  Eigen::MatrixXd v(n_v, 3), vo;
  Eigen::MatrixXi f(n_f, 3), fo;
  Eigen::VectorXi map;
  for (uint32_t i = 0; i < n_v; ++i)
    for (int j = 0; j < 3; ++j)
      v(i, j) = V[3 * i + j];
  for (uint32_t i = 0; i < n_f; ++i)
    for (int j = 0; j < 3; ++j)
      f(i, j) = F[3 * i + j];
  r->mc->removeDuplicateVerticesTriangle(v, f, eps, vo, fo, map);
  *n_v_out = (uint32_t) vo.rows();
  for (int i = 0; i < vo.rows(); ++i)
    for (int j = 0; j < 3; ++j)
      V_out[3 * i + j] = vo(i, j);
  for (int i = 0; i < fo.rows(); ++i)
    for (int j = 0; j < 3; ++j)
      F_out[3 * i + j] = fo(i, j);
  for (int i = 0; i < map.size(); ++i)
    map_out[i] = map(i);
  return 0;
}
int ref_remove_duplicate_faces(void* h, const int32_t* F, uint32_t n_f, int32_t* F_out, uint32_t* n_f_out) {
  Ref* r = (Ref*) h;
  Eigen::MatrixXi f(n_f, 3), fo;
  for (uint32_t i = 0; i < n_f; ++i)
    for (int j = 0; j < 3; ++j)
      f(i, j) = F[3 * i + j];
  r->mc->removeDuplicateFacesTriangle(f, fo);
  *n_f_out = (uint32_t) fo.rows();
  for (int i = 0; i < fo.rows(); ++i)
    for (int j = 0; j < 3; ++j)
      F_out[3 * i + j] = fo(i, j);
  return 0;
}

// ---- the reference's Streamer (streamer.cpp / streamer.cu, unmodified), driven as GeoWrapper does
// (geowrapper.cpp:74-76 create, :137-138 stream before integrate, :153 / :560 streamAllOut, :564 serializeData)
int ref_streamer_create(void* h, uint32_t max_blocks_per_pass) {
  Ref* r = (Ref*) h;
  r->streamer = std::make_unique<GeometricStreamer>(r->container.get(), false, "ref_memory_allocation.txt", "ref_streamer_profiler");
  const Eigen::Vector3f ext(r->container->voxel_extents_.x, r->container->voxel_extents_.y, r->container->voxel_extents_.z);
  r->streamer->create(ext, max_blocks_per_pass, 0);
  return 0;
}
// Streamer::stream (streamer.cpp:337-355) when the free pool drops to the threshold, as compute() does
// (geowrapper.cpp:137-138); force != 0 streams unconditionally (tests/test_streamer.cu:92). Returns 1 if it streamed.
int ref_stream(void* h, const float pos[3], float radius, int force) {
  Ref* r = (Ref*) h;
  if (!r->streamer)
    return -1;
  if (!force && r->container->getHeapHighFreeCount() > stream_threshold * r->container->num_sdf_blocks_)
    return 0;
  r->streamer->stream(Eigen::Vector3f(pos[0], pos[1], pos[2]), radius);
  return 1;
}
int ref_stream_all_out(void* h) {
  Ref* r = (Ref*) h;
  if (!r->streamer)
    return -1;
  r->streamer->streamAllOut();
  return 0;
}
int ref_serialize_data(void* h, const char* hash_path, const char* voxel_path) {
  Ref* r = (Ref*) h;
  if (!r->streamer)
    return -1;
  r->streamer->serializeData(hash_path, voxel_path);
  return 0;
}
// debugCheckForDuplicates (streamer.cpp:401-446): percentage of (device table + host grid) records whose key repeats
double ref_duplicates_ratio(void* h) {
  Ref* r = (Ref*) h;
  return r->streamer ? r->streamer->debugCheckForDuplicates() : -1.0;
}
uint32_t ref_grid_blocks(void* h) {
  Ref* r = (Ref*) h;
  uint32_t n = 0;
  if (r->streamer)
    for (const auto& kv : r->streamer->getGrid())
      n += kv.second->getNElements();
  return n;
}

int ref_sizeof_voxel() {
  return (int) sizeof(Voxel);
}
int ref_sizeof_triangle() {
  return (int) sizeof(Triangle);
}

} // extern "C"

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
#include "voxel_data_structures.cuh"

using namespace cupanutils::cugeoutils;

// The reference defines these members in mesh_extractor.cpp (host, needs real Eigen). The harness
// never calls them, but the constructor / vtable reference them, so give the linker empty bodies.
namespace cupanutils {
  namespace cugeoutils {
    template <typename T>
    void MeshExtractor<T>::processTriangles() {
    }
    template <typename T>
    void MeshExtractor<T>::processTrianglesThread() {
    }
    // mesh_extractor.cuh already holds `template class MeshExtractor<Voxel>;`, so force the
    // two bodies above to be emitted by odr-using them.
    void (MeshExtractor<Voxel>::*const harness_keep_a)() = &MeshExtractor<Voxel>::processTriangles;
    void (MeshExtractor<Voxel>::*const harness_keep_b)() = &MeshExtractor<Voxel>::processTrianglesThread;
  } // namespace cugeoutils
} // namespace cupanutils

namespace {
  struct Ref {
    std::unique_ptr<GeometricVoxelContainer> container;
    std::unique_ptr<Camera> camera;
    std::unique_ptr<GeometricMarchingCubes> mc;
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

int ref_sizeof_voxel() {
  return (int) sizeof(Voxel);
}
int ref_sizeof_triangle() {
  return (int) sizeof(Triangle);
}

} // extern "C"

// mrh_mesh.cu — GeoWrapper::extractMesh (geowrapper.cpp:150-230): region loop over the host store,
// marching cubes on the device, weld of the triangle soup on the device (mrh_weld.cu, replacing the
// host post-process MeshExtractor::processTriangles, mesh_extractor.cpp:9-76,156-259), ASCII PLY.
// The soup never leaves the device unless a caller asks for it (mrh_get_triangles).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <thread>

#include "mrh_fmt.h"
#include "mrh_host.h"
#include "mrh_mesh.cuh"

using namespace mrh;

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_));            \
  } while (0)

namespace {
  inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }

  // Streamer::worldToChunks (streamer.cuh:251-260) of a block origin (streamer.cpp:231-233)
  inline void block_chunk(const GatherRecord& r, float size, float ext, int out[3]) {
    const int pos[3] = {r.x, r.y, r.z};
    for (int k = 0; k < 3; ++k) {
      const float pw = ((float) pos[k] * 8.f) * size;
      const float p  = pw / ext;
      const float s  = (float) ((0.f < p) - (p < 0.f));
      out[k]         = (int) (p + s * 0.5f);
    }
  }

  // "%g" is what `ostream << double` prints at the default precision (geowrapper.cpp:194-229);
  // chunks are formatted in parallel and written in order.
  void write_mesh_ply(const char* path, const HostMesh& mesh) {
    const size_t nv = mesh.vertices.size() / 3, nf = mesh.faces.size() / 3;
    FILE* f = fopen(path, "wb");
    if (!f) {
      std::cerr << "GeoWrapper::extractMesh | Failed to open file for writing: " << path << std::endl;
      return;
    }
    fprintf(f, "ply\nformat ascii 1.0\nelement vertex %zu\nproperty float x\nproperty float y\nproperty float z\n", nv);
    fprintf(f, "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face %zu\n", nf);
    fprintf(f, "property list uchar int vertex_indices\nend_header\n");
    // Batches of nthr x 64 Ki lines are formatted by nthr threads into reusable raw buffers (no zero
    // fill) while the previous batch is being written: formatting and fwrite overlap.
    const unsigned nthr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const size_t chunk  = 1 << 16;
    auto emit = [&](size_t total, size_t max_line, auto&& line) {
      std::unique_ptr<char[]> buf[2];
      std::vector<size_t> len[2];
      for (int k = 0; k < 2; ++k) {
        buf[k].reset(new char[(size_t) nthr * chunk * max_line]);
        len[k].assign(nthr, 0);
      }
      std::thread writer;
      int cur = 0;
      for (size_t base = 0; base < total; base += chunk * nthr, cur ^= 1) {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthr; ++t)
          th.emplace_back([&, t, cur, base] {
            const size_t lo = std::min(total, base + t * chunk), hi = std::min(total, lo + chunk);
            char* dst  = buf[cur].get() + (size_t) t * chunk * max_line;
            size_t pos = 0;
            for (size_t i = lo; i < hi; ++i)
              pos += line(i, dst + pos);
            len[cur][t] = pos;
          });
        for (auto& x : th)
          x.join();
        if (writer.joinable())
          writer.join(); // the other buffer set is free again, and the file stays in order
        writer = std::thread([&, cur] {
          for (unsigned t = 0; t < nthr; ++t)
            fwrite(buf[cur].get() + (size_t) t * chunk * max_line, 1, len[cur][t], f);
        });
      }
      if (writer.joinable())
        writer.join();
    };
    const double* V = mesh.vertices.data();
    const double* C = mesh.colors.data();
    emit(nv, 96, [&](size_t i, char* dst) {
      // colour cast to uchar (Q4: un-normalised interpolated colours wrap)
      char* p = dst;
      for (int k = 0; k < 3; ++k) {
        p += fmt_g6(V[3 * i + k], p);
        *p++ = ' ';
      }
      for (int k = 0; k < 3; ++k) {
        p += fmt_uint((uint32_t) (unsigned char) (long long) C[3 * i + k], p); // via an integer: double -> uchar out of range is undefined
        *p++ = k == 2 ? '\n' : ' ';
      }
      return (size_t) (p - dst);
    });
    const int32_t* F = mesh.faces.data();
    emit(nf, 48, [&](size_t i, char* dst) {
      char* p = dst;
      *p++ = '3';
      for (int k = 0; k < 3; ++k) {
        *p++ = ' ';
        p += fmt_int(F[3 * i + k], p);
      }
      *p++ = '\n';
      return (size_t) (p - dst);
    });
    fclose(f);
    std::cout << "GeoWrapper::extractMesh | written " << nv << " vertices and " << nf << " faces to " << path << std::endl;
  }

  // soup of the regions meshed so far moves to the accumulation buffer (only a map larger than one
  // streaming region, radius 10 x max_depth, ever takes this path)
  int stash_soup(mrh_map* m) {
    if (m->soup_in_tri == 0)
      return 0;
    const size_t need = m->soup_acc_n + m->soup_in_tri;
    if (need > m->soup_acc_cap) {
      const size_t cap = std::max(need, 2 * m->soup_acc_cap);
      float* grown     = nullptr;
      CK(cudaMalloc(&grown, sizeof(float) * 18 * cap));
      if (m->soup_acc_n)
        CK(cudaMemcpyAsync(grown, m->d_soup_acc, sizeof(float) * 18 * m->soup_acc_n, cudaMemcpyDeviceToDevice, m->stream));
      CK(cudaStreamSynchronize(m->stream));
      cudaFree(m->d_soup_acc);
      m->d_soup_acc = grown, m->soup_acc_cap = cap;
    }
    CK(cudaMemcpyAsync(m->d_soup_acc + 18 * m->soup_acc_n, m->d_tri, sizeof(float) * 18 * m->soup_in_tri, cudaMemcpyDeviceToDevice, m->stream));
    m->soup_acc_n += m->soup_in_tri;
    m->soup_in_tri = 0;
    return 0;
  }

  int run_marching_cubes(mrh_map* m, int force_generic, uint32_t max_centers = 0xFFFFFFFFu) {
    if (stash_soup(m))
      return 1;
    if (m->max_num_triangles > m->d_tri_cap) {
      CK(cudaStreamSynchronize(m->stream));
      cudaFree(m->d_tri);
      m->d_tri = nullptr;
      CK(cudaMalloc(&m->d_tri, sizeof(float) * 18 * m->max_num_triangles));
      m->d_tri_cap = m->max_num_triangles;
    }
    if (!m->d_tri_count)
      CK(cudaMalloc(&m->d_tri_count, sizeof(uint32_t)));
    const double t0 = now_ms();
    CK(cudaMemsetAsync(m->d_tri_count, 0, sizeof(uint32_t), m->stream));
    const uint32_t cap = (uint32_t) std::min<uint64_t>(m->max_num_triangles, 0xFFFFFFFFull);
    static int ctas_per_sm = 0; // resident CTAs per SM: the grid is exactly one wave of persistent CTAs
    if (!ctas_per_sm) {
      cudaFuncSetAttribute(k_mc_blocks, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_mc_blocks, kMcThreads, 0) != cudaSuccess || ctas_per_sm < 1)
        ctas_per_sm = 4;
    }
    k_mc_blocks<<<m->num_sms * ctas_per_sm, kMcThreads, 0, m->stream>>>(m->dev, m->live_cur, m->d_tri, m->d_tri_count, cap, force_generic, max_centers);
    m->launches++;
    CK(cudaGetLastError());
    uint32_t n = 0;
    CK(cudaMemcpyAsync(&n, m->d_tri_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    if (n > cap) {
      fprintf(stderr, "appendTriangle | exceeded max triangles: %u >= %u\n", n, cap);
      n = cap;
    }
    m->soup_in_tri = n;
    std::cout << "MarchingCubesExtractor::extractIsoSurface | triangles extracted: " << n << std::endl;
    m->mesh_ms_kernel += now_ms() - t0;
    return 0;
  }

  // the soup of the last extraction: [d_soup, d_soup + 18 * n)
  const float* device_soup(const mrh_map* m, size_t& n) {
    if (m->soup_acc_n) {
      n = m->soup_acc_n;
      return m->d_soup_acc;
    }
    n = m->soup_in_tri;
    return m->d_tri;
  }

} // namespace

extern "C" {

int mrh_extract_mesh_ex(mrh_map* m, const char* path, int force_generic) {
  if (!m)
    return fail("null handle");
  CK(cudaSetDevice(m->device));
  // streamAllOut (geowrapper.cpp:153), in two halves: the copy to the host store now, the release of
  // the device copy once it is known whether the map can be meshed where it already is (below)
  HostStore& st           = m->store;
  const size_t store_was  = st.recs.size();
  const double t_begin    = now_ms();
  const uint32_t frame    = m->frame_index;
  bool device_copy_intact = false;
  if (m->max_num_triangles != 0 && !getenv("MRH_MESH_ROUND_TRIP")) {
    if (gather_to_host(m, st.recs, st.voxels))
      return 1;
    device_copy_intact = store_was == 0 && !st.recs.empty();
  } else if (mrh_stream_all_out(m)) {
    return 1;
  }
  auto release_device_copy = [&]() -> int {
    if (!device_copy_intact)
      return 0;
    device_copy_intact = false;
    if (reset_map(m))
      return 1;
    m->frame_index = frame; // num_integrated_frames_ survives streaming
    return cudaStreamSynchronize(m->stream) == cudaSuccess ? 0 : fail("extractMesh: release of the device copy failed");
  };
  if (m->max_num_triangles == 0) {
    std::cerr << "GeoWrapper::extractMesh | no triangles to extract" << std::endl;
    return 0;
  }
  m->mesh.clear();
  m->soup_in_tri = m->soup_acc_n = 0;
  m->mesh_ms_stream = m->mesh_ms_kernel = m->mesh_ms_merge = m->mesh_ms_ply = 0;
  std::cout << "GeoWrapper::extractMesh | extracting..." << std::endl;
  const float size  = m->p.virtual_voxel_size;
  const float ext   = (float) m->p.voxel_extents_scale;
  const float radius = 10.f * m->cam.max_depth; // radius_scale_chunk (params.h:35)
  const int radiusi  = (int) radius;
  // computeBounds (streamer.cuh:357-368). The reference's grid keeps the keys of emptied chunks, so
  // the bounds are those of everything that was ever streamed out; here: of the current store.
  int lo[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, hi[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
  for (const GatherRecord& r : st.recs) {
    int c[3];
    block_chunk(r, size, ext, c);
    for (int k = 0; k < 3; ++k)
      lo[k] = std::min(lo[k], c[k]), hi[k] = std::max(hi[k], c[k]);
  }
  if (!st.recs.empty() && radiusi > 0) {
    for (int k = 0; k < 3; ++k)
      if (lo[k] == hi[k])
        hi[k] += 1;
    for (int x = lo[0]; x < hi[0]; x += radiusi)
      for (int y = lo[1]; y < hi[1]; y += radiusi)
        for (int z = lo[2]; z < hi[2]; z += radiusi) {
          // streamInToGPU(chunkToWorld(chunk), radius): chunks whose centre passes isChunkInSphere
          const float centre[3] = {(float) x * ext, (float) y * ext, (float) z * ext};
          std::vector<GatherRecord> in_recs;
          VoxelWords in_vox;
          std::vector<uint8_t> inside(st.recs.size());
          size_t n_in = 0;
          for (size_t i = 0; i < st.recs.size(); ++i) {
            inside[i] = record_in_sphere(m, st.recs[i], centre, radius);
            n_in += inside[i];
          }
          if (device_copy_intact) {
            // First region of a map that was entirely resident. If the region takes in every block and
            // no second region follows, streaming the blocks back in would rebuild exactly the table the
            // device still holds: mesh in place, then release (the second streamAllOut of the loop).
            const bool single = n_in == st.recs.size() && x + radiusi >= hi[0] && y + radiusi >= hi[1] && z + radiusi >= hi[2];
            if (single) {
              if (run_marching_cubes(m, force_generic) || release_device_copy())
                return 1;
              continue;
            }
            if (release_device_copy())
              return 1;
          }
          if (n_in == st.recs.size()) {
            in_recs.swap(st.recs); // the usual case: one region covers the whole map
            in_vox.swap(st.voxels);
          } else if (n_in) {
            HostStore keep;
            for (size_t i = 0; i < st.recs.size(); ++i) {
              std::vector<GatherRecord>& rr = inside[i] ? in_recs : keep.recs;
              VoxelWords& vv                = inside[i] ? in_vox : keep.voxels;
              rr.push_back(st.recs[i]);
              vv.insert(vv.end(), st.voxels.begin() + i * 3 * kBlockVoxels, st.voxels.begin() + (i + 1) * 3 * kBlockVoxels);
            }
            st.recs.swap(keep.recs);
            st.voxels.swap(keep.voxels);
          }
          {
            // Paging a region in must not lose blocks: refuse up front when the pool cannot take the
            // region (the records go back to the host store untouched), and hand them back as well if
            // the insert itself reports a shortfall.
            auto give_back = [&]() {
              st.recs.insert(st.recs.end(), in_recs.begin(), in_recs.end());
              st.voxels.insert(st.voxels.end(), in_vox.begin(), in_vox.end());
            };
            mrh_stats s0;
            if (mrh_get_stats(m, &s0)) {
              give_back();
              return 1;
            }
            if ((int64_t) in_recs.size() > s0.heap_free) {
              give_back();
              return fail("extractMesh: the region holds %zu blocks but the pool has %lld free: a map larger than num_sdf_blocks cannot be meshed in one region - "
                          "raise num_sdf_blocks (currently %llu)",
                          in_recs.size(), (long long) s0.heap_free, (unsigned long long) m->num_sdf_blocks);
            }
            if (insert_from_host(m, in_recs.data(), in_vox.data(), in_recs.size())) {
              give_back();
              return 1;
            }
          }
          if (run_marching_cubes(m, force_generic))
            return 1;
          if (mrh_stream_all_out(m))
            return 1;
        }
  }
  if (release_device_copy()) // (empty store bounds: no region was visited)
    return 1;
  {
    // processTriangles with merge_mesh_ = true over the soups of all regions, in region order
    const double t_weld = now_ms();
    if (m->soup_acc_n && stash_soup(m))
      return 1;
    size_t n_tri      = 0;
    const float* soup = device_soup(m, n_tri);
    if (weld_on_device(m, soup, n_tri, (double) m->p.vertices_merging_threshold))
      return 1;
    m->mesh_ms_merge = now_ms() - t_weld;
  }
  m->mesh_ms_stream = (now_ms() - t_begin) - m->mesh_ms_kernel - m->mesh_ms_merge;
  const double t_ply = now_ms();
  if (path)
    write_mesh_ply(path, m->mesh);
  m->mesh_ms_ply = now_ms() - t_ply;
  return 0;
}

int mrh_write_mesh_ply(const char* path, const double* vertices, const double* colors, const int32_t* faces, size_t n_vertices, size_t n_faces) {
  if (!path || (n_vertices && (!vertices || !colors)) || (n_faces && !faces))
    return fail("null argument");
  HostMesh mesh;
  mesh.vertices.assign(vertices, vertices + 3 * n_vertices);
  mesh.colors.assign(colors, colors + 3 * n_vertices);
  mesh.faces.assign(faces, faces + 3 * n_faces);
  write_mesh_ply(path, mesh);
  return 0;
}

int mrh_extract_mesh(mrh_map* m, const char* path) {
  return mrh_extract_mesh_ex(m, path, 0);
}

// ---- sharded meshing: mesh the blocks this handle owns where they are ----
int mrh_mesh_local(mrh_map* m, size_t* n_triangles) {
  if (!m || !n_triangles)
    return fail("null argument");
  CK(cudaSetDevice(m->device));
  *n_triangles = 0;
  m->mesh.clear();
  m->soup_in_tri = m->soup_acc_n = 0;
  if (m->max_num_triangles == 0)
    return fail("mrh_mesh_local: max_num_triangles is 0");
  if (run_marching_cubes(m, 0, m->halo_active ? m->halo_owned : 0xFFFFFFFFu))
    return 1;
  *n_triangles = m->soup_in_tri;
  return 0;
}

int mrh_copy_triangles_device(mrh_map* m, float* d_dst, size_t cap_triangles) {
  if (!m)
    return fail("null handle");
  CK(cudaSetDevice(m->device));
  size_t n_tri      = 0;
  const float* soup = device_soup(m, n_tri);
  if (n_tri > cap_triangles)
    return fail("mrh_copy_triangles_device: %zu triangles do not fit %zu", n_tri, cap_triangles);
  if (n_tri) {
    if (!d_dst)
      return fail("null argument");
    CK(cudaMemcpyAsync(d_dst, soup, sizeof(float) * 18 * n_tri, cudaMemcpyDeviceToDevice, m->stream));
    CK(cudaStreamSynchronize(m->stream));
  }
  return 0;
}

int mrh_weld_device_soup(mrh_map* m, const float* d_soup, size_t n_triangles, const char* path) {
  if (!m || (n_triangles && !d_soup))
    return fail("null argument");
  CK(cudaSetDevice(m->device));
  const double t0 = now_ms();
  if (weld_on_device(m, d_soup, n_triangles, (double) m->p.vertices_merging_threshold))
    return 1;
  m->mesh_ms_merge = now_ms() - t0;
  const double t1  = now_ms();
  if (path)
    write_mesh_ply(path, m->mesh);
  m->mesh_ms_ply = now_ms() - t1;
  return 0;
}

int mrh_get_mesh(mrh_map* m, const double** v, const int32_t** f, const double** c, size_t* nv, size_t* nf) {
  if (!m || !v || !f || !c || !nv || !nf)
    return fail("null argument");
  *v = m->mesh.vertices.data(), *f = m->mesh.faces.data(), *c = m->mesh.colors.data();
  *nv = m->mesh.vertices.size() / 3, *nf = m->mesh.faces.size() / 3;
  return 0;
}

int mrh_get_triangles(mrh_map* m, const float** t, size_t* n) {
  if (!m || !t || !n)
    return fail("null argument");
  size_t n_tri      = 0;
  const float* soup = device_soup(m, n_tri);
  if (m->mesh.triangles.size() != n_tri * 18) {
    CK(cudaSetDevice(m->device));
    m->mesh.triangles.resize(n_tri * 18);
    if (n_tri)
      CK(cudaMemcpy(m->mesh.triangles.data(), soup, sizeof(float) * 18 * n_tri, cudaMemcpyDeviceToHost));
  }
  *t = m->mesh.triangles.data();
  *n = n_tri;
  return 0;
}

} // extern "C"

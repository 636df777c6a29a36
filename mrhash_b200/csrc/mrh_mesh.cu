// mrh_mesh.cu — GeoWrapper::extractMesh (geowrapper.cpp:150-230): region loop over the host store,
// marching cubes on the device, host merge of the triangle soup (MeshExtractor::processTriangles,
// mesh_extractor.cpp:9-76,156-259), ASCII PLY.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

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

  // MeshExtractor::processTriangles with merge_mesh_ = true, incrementally: vertices already in the
  // mesh keep their first-seen index, exactly as re-running removeDuplicateVerticesTriangle over
  // (unique old, new) does.
  void merge_triangles(HostMesh& mesh, const float* tris, size_t n, double eps) {
    const double inv_eps = eps != 0.0 ? 1.0 / eps : 0.0;
    for (size_t t = 0; t < n; ++t) {
      int32_t idx[3];
      for (int j = 0; j < 3; ++j) {
        const float* v    = tris + t * 18 + j * 6;
        const double p[3] = {(double) v[0], (double) v[1], (double) v[2]};
        VertexKey key;
        if (eps == 0.0) {
          memcpy(&key.a, &p[0], 8), memcpy(&key.b, &p[1], 8), memcpy(&key.c, &p[2], 8);
        } else {
          key.a = (uint64_t) (int64_t) (int) std::floor(p[0] * inv_eps);
          key.b = (uint64_t) (int64_t) (int) std::floor(p[1] * inv_eps);
          key.c = (uint64_t) (int64_t) (int) std::floor(p[2] * inv_eps);
        }
        auto it = mesh.vertex_map.find(key);
        if (it != mesh.vertex_map.end()) {
          idx[j] = it->second;
        } else {
          idx[j] = (int32_t) (mesh.vertices.size() / 3);
          mesh.vertex_map.emplace(key, idx[j]);
          mesh.vertices.insert(mesh.vertices.end(), {p[0], p[1], p[2]});
          mesh.colors.insert(mesh.colors.end(), {(double) v[3], (double) v[4], (double) v[5]});
        }
      }
      if (idx[0] == idx[1] || idx[0] == idx[2] || idx[1] == idx[2])
        continue; // degenerate after the merge
      if (!mesh.face_set.insert({idx[0], idx[1], idx[2]}).second)
        continue; // duplicate face
      mesh.faces.insert(mesh.faces.end(), {idx[0], idx[1], idx[2]});
    }
  }

  int run_marching_cubes(mrh_map* m, int force_generic) {
    if (m->max_num_triangles > m->d_tri_cap) {
      cudaFree(m->d_tri);
      m->d_tri = nullptr;
      CK(cudaMalloc(&m->d_tri, sizeof(float) * 18 * m->max_num_triangles));
      m->d_tri_cap = m->max_num_triangles;
    }
    if (!m->d_tri_count)
      CK(cudaMalloc(&m->d_tri_count, sizeof(uint32_t)));
    CK(cudaMemsetAsync(m->d_tri_count, 0, sizeof(uint32_t), m->stream));
    const uint32_t cap = (uint32_t) std::min<uint64_t>(m->max_num_triangles, 0xFFFFFFFFull);
    k_mc_blocks<<<m->num_sms * 4, 256, 0, m->stream>>>(m->dev, m->live_cur, m->d_tri, m->d_tri_count, cap, force_generic);
    m->launches++;
    CK(cudaGetLastError());
    uint32_t n = 0;
    CK(cudaMemcpyAsync(&n, m->d_tri_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    if (n > cap) {
      fprintf(stderr, "appendTriangle | exceeded max triangles: %u >= %u\n", n, cap);
      n = cap;
    }
    const size_t base = m->mesh.triangles.size();
    m->mesh.triangles.resize(base + (size_t) n * 18);
    if (n)
      CK(cudaMemcpy(m->mesh.triangles.data() + base, m->d_tri, sizeof(float) * 18 * n, cudaMemcpyDeviceToHost));
    std::cout << "MarchingCubesExtractor::extractIsoSurface | triangles extracted: " << n << std::endl;
    if (n)
      merge_triangles(m->mesh, m->mesh.triangles.data() + base, n, (double) m->p.vertices_merging_threshold);
    return 0;
  }

} // namespace

extern "C" {

int mrh_extract_mesh_ex(mrh_map* m, const char* path, int force_generic) {
  if (!m)
    return fail("null handle");
  CK(cudaSetDevice(m->device));
  if (mrh_stream_all_out(m))
    return 1;
  if (m->max_num_triangles == 0) {
    std::cerr << "GeoWrapper::extractMesh | no triangles to extract" << std::endl;
    return 0;
  }
  m->mesh.clear();
  std::cout << "GeoWrapper::extractMesh | extracting..." << std::endl;
  HostStore& st     = m->store;
  const float size  = m->p.virtual_voxel_size;
  const float ext   = (float) m->p.voxel_extents_scale;
  const float radius = 10.f * m->cam.max_depth; // radius_scale_chunk (params.h:35)
  const int radiusi  = (int) radius;
  const float chunk_radius = 0.5f * ext * std::sqrt(3.f); // streamer.cpp:15
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
          std::vector<uint32_t> in_vox;
          HostStore keep;
          for (size_t i = 0; i < st.recs.size(); ++i) {
            int c[3];
            block_chunk(st.recs[i], size, ext, c);
            const float d[3] = {(float) c[0] * ext - centre[0], (float) c[1] * ext - centre[1], (float) c[2] * ext - centre[2]};
            const float l    = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            const bool in    = l <= std::fabs(radius - chunk_radius);
            std::vector<GatherRecord>& rr = in ? in_recs : keep.recs;
            std::vector<uint32_t>& vv     = in ? in_vox : keep.voxels;
            rr.push_back(st.recs[i]);
            vv.insert(vv.end(), st.voxels.begin() + i * 3 * kBlockVoxels, st.voxels.begin() + (i + 1) * 3 * kBlockVoxels);
          }
          st.recs.swap(keep.recs);
          st.voxels.swap(keep.voxels);
          if (insert_from_host(m, in_recs.data(), in_vox.data(), in_recs.size()))
            return 1;
          if (run_marching_cubes(m, force_generic))
            return 1;
          if (mrh_stream_all_out(m))
            return 1;
        }
  }
  const size_t nv = m->mesh.vertices.size() / 3, nf = m->mesh.faces.size() / 3;
  if (path) {
    // geowrapper.cpp:187-229: ASCII PLY, default ostream precision, colour cast to uchar (Q4)
    std::ofstream ply(path);
    if (!ply.is_open()) {
      std::cerr << "GeoWrapper::extractMesh | Failed to open file for writing: " << path << std::endl;
      return 0;
    }
    ply << "ply\nformat ascii 1.0\nelement vertex " << nv << "\nproperty float x\nproperty float y\nproperty float z\n";
    ply << "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face " << nf << "\n";
    ply << "property list uchar int vertex_indices\nend_header\n";
    const double* V = m->mesh.vertices.data();
    const double* C = m->mesh.colors.data();
    for (size_t i = 0; i < nv; ++i) {
      const unsigned char col[3] = {(unsigned char) C[3 * i], (unsigned char) C[3 * i + 1], (unsigned char) C[3 * i + 2]};
      ply << V[3 * i] << " " << V[3 * i + 1] << " " << V[3 * i + 2] << " " << (int) col[0] << " " << (int) col[1] << " " << (int) col[2] << "\n";
    }
    const int32_t* F = m->mesh.faces.data();
    for (size_t i = 0; i < nf; ++i)
      ply << "3 " << F[3 * i] << " " << F[3 * i + 1] << " " << F[3 * i + 2] << "\n";
    ply.close();
    std::cout << "GeoWrapper::extractMesh | written " << nv << " vertices and " << nf << " faces to " << path << std::endl;
  }
  return 0;
}

int mrh_extract_mesh(mrh_map* m, const char* path) {
  return mrh_extract_mesh_ex(m, path, 0);
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
  *t = m->mesh.triangles.data();
  *n = m->mesh.triangles.size() / 18;
  return 0;
}

} // extern "C"

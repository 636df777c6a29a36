// mrh_mesh.cu — GeoWrapper::extractMesh (geowrapper.cpp:150-230): region loop over the host store,
// marching cubes on the device, host merge of the triangle soup (MeshExtractor::processTriangles,
// mesh_extractor.cpp:9-76,156-259), ASCII PLY.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>

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

  // MeshExtractor::processTriangles with merge_mesh_ = true, incrementally: vertices already in the
  // mesh keep their first-seen index, exactly as re-running removeDuplicateVerticesTriangle over
  // (unique old, new) does.
  void merge_triangles(HostMesh& mesh, const float* tris, size_t n, double eps) {
    const double inv_eps = eps != 0.0 ? 1.0 / eps : 0.0;
    mesh.vertex_map.reserve(mesh.vertex_map.count + 3 * n);
    mesh.face_map.reserve(mesh.face_map.count + n);
    mesh.vertices.reserve(mesh.vertices.size() + 6 * n);
    mesh.colors.reserve(mesh.colors.size() + 6 * n);
    mesh.faces.reserve(mesh.faces.size() + 3 * n);
    for (size_t t = 0; t < n; ++t) {
      int32_t idx[3];
      for (int j = 0; j < 3; ++j) {
        const float* v = tris + t * 18 + j * 6;
        uint32_t key[3];
        if (eps == 0.0) {
          // exact merge: byte-wise equality of the doubles == of the floats they were widened from
          memcpy(key, v, 12);
        } else {
          for (int k = 0; k < 3; ++k)
            key[k] = (uint32_t) (int) std::floor((double) v[k] * inv_eps);
        }
        bool inserted;
        idx[j] = mesh.vertex_map.find_or_insert(key[0], key[1], key[2], (int32_t) (mesh.vertices.size() / 3), inserted);
        if (inserted) {
          mesh.vertices.insert(mesh.vertices.end(), {(double) v[0], (double) v[1], (double) v[2]});
          mesh.colors.insert(mesh.colors.end(), {(double) v[3], (double) v[4], (double) v[5]});
        }
      }
      if (idx[0] == idx[1] || idx[0] == idx[2] || idx[1] == idx[2])
        continue; // degenerate after the merge
      bool inserted;
      mesh.face_map.find_or_insert((uint32_t) idx[0], (uint32_t) idx[1], (uint32_t) idx[2], 0, inserted);
      if (!inserted)
        continue; // duplicate face
      mesh.faces.insert(mesh.faces.end(), {idx[0], idx[1], idx[2]});
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
    const unsigned nthr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    auto emit = [&](size_t total, size_t max_line, auto&& line) {
      const size_t chunk = 1 << 16;
      for (size_t base = 0; base < total; base += chunk * nthr) {
        std::vector<std::string> out(nthr);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthr; ++t)
          th.emplace_back([&, t] {
            const size_t lo = base + t * chunk, hi = std::min(total, lo + chunk);
            if (lo >= hi)
              return;
            std::string& s = out[t];
            s.resize((hi - lo) * max_line);
            size_t pos = 0;
            for (size_t i = lo; i < hi; ++i)
              pos += line(i, &s[pos]);
            s.resize(pos);
          });
        for (auto& x : th)
          x.join();
        for (auto& s : out)
          fwrite(s.data(), 1, s.size(), f);
      }
    };
    const double* V = mesh.vertices.data();
    const double* C = mesh.colors.data();
    emit(nv, 96, [&](size_t i, char* dst) {
      // colour cast to uchar (Q4: un-normalised interpolated colours wrap)
      return (size_t) sprintf(dst, "%g %g %g %d %d %d\n", V[3 * i], V[3 * i + 1], V[3 * i + 2], (int) (unsigned char) C[3 * i], (int) (unsigned char) C[3 * i + 1], (int) (unsigned char) C[3 * i + 2]);
    });
    const int32_t* F = mesh.faces.data();
    emit(nf, 48, [&](size_t i, char* dst) { return (size_t) sprintf(dst, "3 %d %d %d\n", F[3 * i], F[3 * i + 1], F[3 * i + 2]); });
    fclose(f);
    std::cout << "GeoWrapper::extractMesh | written " << nv << " vertices and " << nf << " faces to " << path << std::endl;
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
    const double t0 = now_ms();
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
    const double t1 = now_ms();
    if (n)
      merge_triangles(m->mesh, m->mesh.triangles.data() + base, n, (double) m->p.vertices_merging_threshold);
    m->mesh_ms_kernel += t1 - t0;
    m->mesh_ms_merge += now_ms() - t1;
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
  m->mesh_ms_stream = m->mesh_ms_kernel = m->mesh_ms_merge = m->mesh_ms_ply = 0;
  const double t_begin = now_ms();
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
          std::vector<uint8_t> inside(st.recs.size());
          size_t n_in = 0;
          for (size_t i = 0; i < st.recs.size(); ++i) {
            int c[3];
            block_chunk(st.recs[i], size, ext, c);
            const float d[3] = {(float) c[0] * ext - centre[0], (float) c[1] * ext - centre[1], (float) c[2] * ext - centre[2]};
            const float l    = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            inside[i]        = l <= std::fabs(radius - chunk_radius);
            n_in += inside[i];
          }
          if (n_in == st.recs.size()) {
            in_recs.swap(st.recs); // the usual case: one region covers the whole map
            in_vox.swap(st.voxels);
          } else if (n_in) {
            HostStore keep;
            for (size_t i = 0; i < st.recs.size(); ++i) {
              std::vector<GatherRecord>& rr = inside[i] ? in_recs : keep.recs;
              std::vector<uint32_t>& vv     = inside[i] ? in_vox : keep.voxels;
              rr.push_back(st.recs[i]);
              vv.insert(vv.end(), st.voxels.begin() + i * 3 * kBlockVoxels, st.voxels.begin() + (i + 1) * 3 * kBlockVoxels);
            }
            st.recs.swap(keep.recs);
            st.voxels.swap(keep.voxels);
          }
          if (insert_from_host(m, in_recs.data(), in_vox.data(), in_recs.size()))
            return 1;
          if (run_marching_cubes(m, force_generic))
            return 1;
          if (mrh_stream_all_out(m))
            return 1;
        }
  }
  const size_t nv = m->mesh.vertices.size() / 3, nf = m->mesh.faces.size() / 3;
  m->mesh_ms_stream = (now_ms() - t_begin) - m->mesh_ms_kernel - m->mesh_ms_merge;
  const double t_ply = now_ms();
  if (path)
    write_mesh_ply(path, m->mesh);
  m->mesh_ms_ply = now_ms() - t_ply;
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

// mrh_weld.cu — vertex weld and face de-duplication of the marching-cubes triangle soup on the device.
//
// Replaces the host post-process of MeshExtractor::processTriangles (mesh_extractor.cpp:9-76):
// removeDuplicateVerticesTriangle (:181-259, an unordered_map<Vector3d,int> over 3T vertices, keyed
// on the bytes of the position or on floor(v / eps) cells), the degenerate-face filter (:55-67) and
// removeDuplicateFacesTriangle (:156-179, a std::set of index triples). Same result, index for index:
// a vertex keeps the number it gets in FIRST-SEEN order of the soup, the first copy's position and
// colour win, a face survives if it is not degenerate and no earlier face has the same ordered triple.
//
// "First seen" is made order-independent of the thread schedule by keeping, per key, the MINIMUM
// soup index: an open-addressing table of u32 cells holds one representative index per key class;
// a thread that finds a representative with an equal key lowers it with atomicMin. A prefix sum over
// "I am my class's minimum" then numbers the classes in soup order. Algorithmic bytes: 72 B per
// triangle read twice (insert, resolve) + 4 B cells + 48 B per unique vertex and 12 B per face out.
#include <cub/device/device_scan.cuh>

#include "mrh_host.h"

using namespace mrh;

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_));            \
  } while (0)

namespace {

  constexpr uint32_t kNone = 0xFFFFFFFFu;

  struct Key3 {
    uint32_t a, b, c;
  };
  __device__ __forceinline__ bool operator==(const Key3& x, const Key3& y) {
    return x.a == y.a && x.b == y.b && x.c == y.c;
  }
  __device__ __forceinline__ uint32_t mix3(Key3 k) {
    unsigned long long h = (unsigned long long) k.a * 0x9E3779B97F4A7C15ull;
    h ^= (unsigned long long) k.b * 0xC2B2AE3D27D4EB4Full + (h >> 29);
    h ^= (unsigned long long) k.c * 0x165667B19E3779F9ull + (h << 7);
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ull;
    return (uint32_t) (h >> 32);
  }

  // key of soup vertex i (mesh_extractor.cpp:196-207 exact, :224-240 quantised).
  // Exact mode: the reference's map is unordered_map<Vector3d, int, Vector3dHash, Vector3dEqual>
  // (mesh_extractor.cuh:25-36): the hash runs over the 24 BYTES of the vertex, equality is a == b.
  // Two vertices therefore meet only if their bytes agree (-0.0 and +0.0 hash apart, so they stay two
  // vertices unless their hashes happen to share a bucket), and a vertex with a NaN coordinate never
  // equals anything, itself included: the key is the bit pattern, and a NaN vertex gets a key of its own
  // (0xFFFFFFFF in x is a NaN pattern, so no non-NaN vertex can produce it).
  __device__ __forceinline__ Key3 vertex_key(const float* __restrict__ soup, uint32_t i, double inv_eps) {
    const float* v = soup + (size_t) i * 6;
    if (inv_eps == 0.0) {
      if (v[0] != v[0] || v[1] != v[1] || v[2] != v[2])
        return {0xFFFFFFFFu, i, 0u};
      return {__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2])};
    }
    return {(uint32_t) (int) floor((double) v[0] * inv_eps), (uint32_t) (int) floor((double) v[1] * inv_eps), (uint32_t) (int) floor((double) v[2] * inv_eps)};
  }

  // insert-or-lower: afterwards the cell of i's key class holds the smallest index of the class
  template <class KeyOf>
  __device__ __forceinline__ void class_insert(uint32_t* __restrict__ cells, uint32_t mask, uint32_t i, const KeyOf& key_of) {
    const Key3 k = key_of(i);
    uint32_t h   = mix3(k) & mask;
    for (;;) {
      uint32_t c = cells[h];
      if (c == kNone) {
        c = atomicCAS(cells + h, kNone, i);
        if (c == kNone)
          return;
      }
      if (c == i)
        return;
      if (key_of(c) == k) {
        atomicMin(cells + h, i);
        return;
      }
      h = (h + 1) & mask;
    }
  }
  template <class KeyOf>
  __device__ __forceinline__ uint32_t class_find(const uint32_t* __restrict__ cells, uint32_t mask, uint32_t i, const KeyOf& key_of) {
    const Key3 k = key_of(i);
    uint32_t h   = mix3(k) & mask;
    for (;;) {
      const uint32_t c = cells[h];
      if (c == i || key_of(c) == k)
        return c;
      h = (h + 1) & mask;
    }
  }

  struct VertexKeyOf {
    const float* soup;
    double inv_eps;
    __device__ __forceinline__ Key3 operator()(uint32_t i) const {
      return vertex_key(soup, i, inv_eps);
    }
  };
  struct FaceKeyOf {
    const int32_t* ids; // [T][3]
    __device__ __forceinline__ Key3 operator()(uint32_t t) const {
      return {(uint32_t) ids[3 * (size_t) t], (uint32_t) ids[3 * (size_t) t + 1], (uint32_t) ids[3 * (size_t) t + 2]};
    }
  };

  __global__ void __launch_bounds__(256) k_weld_insert_vertices(const float* __restrict__ soup, uint32_t n, double inv_eps, uint32_t* __restrict__ cells, uint32_t mask) {
    const VertexKeyOf key_of{soup, inv_eps};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
      class_insert(cells, mask, i, key_of);
  }

  // rep[i] = first-seen index of i's class, first[i] = 1 when i is that index
  __global__ void __launch_bounds__(256) k_weld_resolve_vertices(const float* __restrict__ soup, uint32_t n, double inv_eps, const uint32_t* __restrict__ cells, uint32_t mask, uint32_t* __restrict__ rep,
                                                                 uint32_t* __restrict__ first) {
    const VertexKeyOf key_of{soup, inv_eps};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const uint32_t r = class_find(cells, mask, i, key_of);
      rep[i]           = r;
      first[i]         = r == i ? 1u : 0u;
    }
  }

  // unique vertices (f64, as the reference's MatrixXd) in first-seen order; per-triangle new indices;
  // non-degenerate faces enter the face table
  __global__ void __launch_bounds__(256) k_weld_emit_vertices(const float* __restrict__ soup, uint32_t n, const uint32_t* __restrict__ first, const uint32_t* __restrict__ new_id, double* __restrict__ V,
                                                              double* __restrict__ C) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      if (!first[i])
        continue;
      const float* v  = soup + (size_t) i * 6;
      const size_t o  = (size_t) new_id[i] * 3;
      V[o] = (double) v[0], V[o + 1] = (double) v[1], V[o + 2] = (double) v[2];
      C[o] = (double) v[3], C[o + 1] = (double) v[4], C[o + 2] = (double) v[5];
    }
  }
  __global__ void __launch_bounds__(256) k_weld_face_ids(uint32_t n_tri, const uint32_t* __restrict__ rep, const uint32_t* __restrict__ new_id, int32_t* __restrict__ ids, uint32_t* __restrict__ keep) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tri; t += gridDim.x * blockDim.x) {
      const int32_t a = (int32_t) new_id[rep[3 * (size_t) t]], b = (int32_t) new_id[rep[3 * (size_t) t + 1]], c = (int32_t) new_id[rep[3 * (size_t) t + 2]];
      ids[3 * (size_t) t] = a, ids[3 * (size_t) t + 1] = b, ids[3 * (size_t) t + 2] = c;
      keep[t] = (a != b && a != c && b != c) ? 1u : 0u; // mesh_extractor.cpp:55-57
    }
  }
  __global__ void __launch_bounds__(256) k_weld_insert_faces(uint32_t n_tri, const int32_t* __restrict__ ids, const uint32_t* __restrict__ keep, uint32_t* __restrict__ cells, uint32_t mask) {
    const FaceKeyOf key_of{ids};
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tri; t += gridDim.x * blockDim.x)
      if (keep[t])
        class_insert(cells, mask, t, key_of);
  }
  __global__ void __launch_bounds__(256) k_weld_resolve_faces(uint32_t n_tri, const int32_t* __restrict__ ids, const uint32_t* __restrict__ cells, uint32_t mask, uint32_t* __restrict__ keep) {
    const FaceKeyOf key_of{ids};
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tri; t += gridDim.x * blockDim.x)
      if (keep[t] && class_find(cells, mask, t, key_of) != t)
        keep[t] = 0u; // an earlier face has the same ordered triple (mesh_extractor.cpp:156-179)
  }
  __global__ void __launch_bounds__(256) k_weld_emit_faces(uint32_t n_tri, const int32_t* __restrict__ ids, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos, int32_t* __restrict__ F) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tri; t += gridDim.x * blockDim.x) {
      if (!keep[t])
        continue;
      const size_t o = (size_t) pos[t] * 3;
      F[o] = ids[3 * (size_t) t], F[o + 1] = ids[3 * (size_t) t + 1], F[o + 2] = ids[3 * (size_t) t + 2];
    }
  }

  struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
      cudaFree(p);
    }
    cudaError_t alloc(size_t bytes) {
      return cudaMalloc(&p, bytes ? bytes : 1);
    }
    template <class T>
    T* as() const {
      return static_cast<T*>(p);
    }
  };

  // last element of an exclusive scan + its flag = the total
  int scan_total(const uint32_t* d_flags, const uint32_t* d_scan, size_t n, cudaStream_t s, uint32_t& total) {
    uint32_t a = 0, b = 0;
    CK(cudaMemcpyAsync(&a, d_flags + n - 1, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&b, d_scan + n - 1, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    total = a + b;
    return 0;
  }

} // namespace

namespace mrh {

  // Welds n_tri triangles of d_soup (18 floats each, device memory) into mesh.vertices / colors /
  // faces. The soup itself stays on the device (mrh_get_triangles fetches it on demand).
  int weld_on_device(mrh_map* m, const float* d_soup, size_t n_tri, double eps) {
    HostMesh& mesh = m->mesh;
    mesh.vertices.clear(), mesh.colors.clear(), mesh.faces.clear();
    if (n_tri == 0)
      return 0;
    if (3 * n_tri >= 0x7FFFFFFFull)
      return fail("extractMesh: %zu triangles exceed the 32-bit vertex index of the mesh", n_tri);
    cudaStream_t s       = m->stream;
    const uint32_t T     = (uint32_t) n_tri, N = 3u * T;
    const double inv_eps = eps != 0.0 ? 1.0 / eps : 0.0;
    const int grid       = m->num_sms * 8;
    uint32_t bits        = 10;
    while ((1ull << bits) < 2ull * N)
      ++bits;
    const uint32_t mask = (uint32_t) ((1ull << bits) - 1ull);

    DevBuf cells, rep, first, new_id, scan_tmp, ids, keep;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const uint32_t*) nullptr, (uint32_t*) nullptr, (int) N, s);
    if (cells.alloc(sizeof(uint32_t) << bits) != cudaSuccess || rep.alloc(4ull * N) != cudaSuccess || first.alloc(4ull * N) != cudaSuccess || new_id.alloc(4ull * N) != cudaSuccess ||
        scan_tmp.alloc(tmp_bytes) != cudaSuccess || ids.alloc(12ull * T) != cudaSuccess || keep.alloc(4ull * T) != cudaSuccess)
      return fail("extractMesh: out of device memory for the weld of %zu triangles", n_tri);

    // ---- vertices ----
    CK(cudaMemsetAsync(cells.p, 0xFF, sizeof(uint32_t) << bits, s));
    k_weld_insert_vertices<<<grid, 256, 0, s>>>(d_soup, N, inv_eps, cells.as<uint32_t>(), mask);
    k_weld_resolve_vertices<<<grid, 256, 0, s>>>(d_soup, N, inv_eps, cells.as<uint32_t>(), mask, rep.as<uint32_t>(), first.as<uint32_t>());
    CK(cub::DeviceScan::ExclusiveSum(scan_tmp.p, tmp_bytes, first.as<uint32_t>(), new_id.as<uint32_t>(), (int) N, s));
    uint32_t n_unique = 0;
    if (scan_total(first.as<uint32_t>(), new_id.as<uint32_t>(), N, s, n_unique))
      return 1;
    DevBuf dV, dC;
    if (dV.alloc(24ull * n_unique) != cudaSuccess || dC.alloc(24ull * n_unique) != cudaSuccess)
      return fail("extractMesh: out of device memory for %u welded vertices", n_unique);
    k_weld_emit_vertices<<<grid, 256, 0, s>>>(d_soup, N, first.as<uint32_t>(), new_id.as<uint32_t>(), dV.as<double>(), dC.as<double>());
    // ---- faces ----
    k_weld_face_ids<<<grid, 256, 0, s>>>(T, rep.as<uint32_t>(), new_id.as<uint32_t>(), ids.as<int32_t>(), keep.as<uint32_t>());
    CK(cudaMemsetAsync(cells.p, 0xFF, sizeof(uint32_t) << bits, s));
    k_weld_insert_faces<<<grid, 256, 0, s>>>(T, ids.as<int32_t>(), keep.as<uint32_t>(), cells.as<uint32_t>(), mask);
    k_weld_resolve_faces<<<grid, 256, 0, s>>>(T, ids.as<int32_t>(), cells.as<uint32_t>(), mask, keep.as<uint32_t>());
    uint32_t* face_pos = rep.as<uint32_t>(); // rep is no longer needed
    CK(cub::DeviceScan::ExclusiveSum(scan_tmp.p, tmp_bytes, keep.as<uint32_t>(), face_pos, (int) T, s));
    m->launches += 9;
    CK(cudaGetLastError());
    uint32_t n_faces = 0;
    if (scan_total(keep.as<uint32_t>(), face_pos, T, s, n_faces))
      return 1;
    DevBuf dF;
    if (dF.alloc(12ull * n_faces) != cudaSuccess)
      return fail("extractMesh: out of device memory for %u faces", n_faces);
    k_weld_emit_faces<<<grid, 256, 0, s>>>(T, ids.as<int32_t>(), keep.as<uint32_t>(), face_pos, dF.as<int32_t>());
    m->launches += 1;
    CK(cudaGetLastError());
    mesh.vertices.resize(3ull * n_unique), mesh.colors.resize(3ull * n_unique), mesh.faces.resize(3ull * n_faces);
    // pageable destinations: pinned bounce buffers, DMA overlapped with a parallel host copy (mrh_state.cu)
    if (bulk_d2h(m, mesh.vertices.data(), dV.p, 24ull * n_unique) || bulk_d2h(m, mesh.colors.data(), dC.p, 24ull * n_unique) || (n_faces && bulk_d2h(m, mesh.faces.data(), dF.p, 12ull * n_faces)))
      return 1;
    CK(cudaStreamSynchronize(s));
    return 0;
  }

} // namespace mrh

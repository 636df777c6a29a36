// mrh_state.cu — moving whole blocks between the device map and the host store:
// streamAllOut / streamInToGPU / serializeData (Streamer, streamer.cpp:104-160, 216-281, 290-378)
// and the parity dump. The Streamer's pass-1/pass-2 kernels with their per-thread serial prefix
// sums (streamer.cu:77-187, 250-329) and per-block blocking copies (streamer.cpp:316-322) are
// replaced by one gather / one scatter kernel over dense lists and one bulk copy each way.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <thread>
#include <unordered_map>

#include "mrh_host.h"
#include "mrh_table.cuh"

using namespace mrh;

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_));            \
  } while (0)

namespace mrh {

// device staging buffer released on every return path (the CK macro returns from the caller)
template <typename T>
struct ScopedDev {
  T* p = nullptr;
  ~ScopedDev() {
    cudaFree(p);
  }
  cudaError_t alloc(size_t count) {
    return cudaMalloc(&p, sizeof(T) * (count ? count : 1));
  }
};

// ---------------------------------------------------------------------------------------------
// k_gather_blocks: live list -> dense (record, AoS payload) buffers for streamAllOut /
// serializeData / the parity dump (replaces the Streamer's integrateFromGlobalHashPass1/2,
// streamer.cu:77-187: no per-thread serial prefix sums).
// ---------------------------------------------------------------------------------------------
// far_release: radius paging (integrateFromGlobalHashPass1Kernel, streamer.cu:11-59): only blocks whose
// origin is at least `radius` away from `centre` are gathered, and each one leaves the map (key
// tombstoned, pool storage zeroed and pushed back on its heap, live entry invalidated).
struct FarFilter {
  float cx, cy, cz, radius;
  int far_release;
};
__global__ void __launch_bounds__(128) k_gather_blocks(MapDev m, uint32_t live_cur, GatherRecord* records, uint32_t* voxels_aos, uint32_t* out_count, uint32_t first, uint32_t max_out, FarFilter ff) {
  __shared__ uint32_t s_out;
  const int tid    = threadIdx.x;
  const uint32_t n = min(m.ctr->live_count[live_cur], first + max_out);
  for (uint32_t i = first + blockIdx.x; i < n; i += gridDim.x) {
    const LiveEntry le = m.live[live_cur][i];
    if (le.slot == kInvalid)
      continue;
    const unsigned long long key = le.key;
    const uint32_t val           = le.val;
    if (ff.far_release) {
      // SDFBlockToWorldPoint (voxel_hash_utils.cuh:163-165) and length(a, b) (cuda_math.cuh:1062-1065)
      const i3 kb  = unpack_key(key);
      const float dx = fsub(fmul(i2f(kb.x * kBlockSide), m.voxel_size), ff.cx), dy = fsub(fmul(i2f(kb.y * kBlockSide), m.voxel_size), ff.cy),
                  dz = fsub(fmul(i2f(kb.z * kBlockSide), m.voxel_size), ff.cz);
      const float d  = __fsqrt_rn(ffma(dz, dz, ffma(dx, dx, fmul(dy, dy))));
      if (!(d >= ff.radius))
        continue;
    }
    if (tid == 0)
      s_out = atomicAdd(out_count, 1u);
    __syncthreads();
    const uint32_t o = s_out;
    __syncthreads();
    if (o >= max_out)
      continue;
    const i3 b = unpack_key(key);
    if (tid == 0)
      records[o] = {b.x, b.y, b.z, (int) (val >> 31), (int) ((val & 0x7FFFFFFFu) * ((val >> 31) ? 64u : 512u))};
    uint32_t* dst = voxels_aos + (size_t) o * kBlockVoxels * 3;
    if (!(val >> 31)) {
      const uint8_t* base = m.pool + (size_t) val * kBlockBytes;
      const float4 sdf4   = reinterpret_cast<const float4*>(base)[tid];
      const float4 ss4    = reinterpret_cast<const float4*>(base + kPlaneBytes)[tid];
      const uint4 cw4     = reinterpret_cast<const uint4*>(base + 2 * kPlaneBytes)[tid];
      const float sv[4] = {sdf4.x, sdf4.y, sdf4.z, sdf4.w}, qv[4] = {ss4.x, ss4.y, ss4.z, ss4.w};
      const uint32_t cv[4] = {cw4.x, cw4.y, cw4.z, cw4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dst[(tid * 4 + j) * 3 + 0] = __float_as_uint(sv[j]);
        dst[(tid * 4 + j) * 3 + 1] = __float_as_uint(qv[j]);
        dst[(tid * 4 + j) * 3 + 2] = cv[j];
      }
    } else {
      // resolution 1: 64 voxels in a 768-byte sub-slot laid out as sdf[64] | sum_sq[64] | rgbw[64]
      const uint8_t* base = m.pool + (size_t) (val & 0x7FFFFFFFu) * 768u;
      for (int v = tid; v < kBlockVoxels; v += 128) {
        uint32_t a = 0, b2 = 0, c = 0;
        if (v < 64) {
          a  = reinterpret_cast<const uint32_t*>(base)[v];
          b2 = reinterpret_cast<const uint32_t*>(base + 256)[v];
          c  = reinterpret_cast<const uint32_t*>(base + 512)[v];
        }
        dst[v * 3 + 0] = a, dst[v * 3 + 1] = b2, dst[v * 3 + 2] = c;
      }
    }
    if (ff.far_release) {
      __syncthreads(); // every thread has read its voxels
      if (!(val >> 31)) {
        uint8_t* base  = m.pool + (size_t) val * kBlockBytes;
        const float4 z = {0.f, 0.f, 0.f, 0.f};
        reinterpret_cast<float4*>(base)[tid]                   = z;
        reinterpret_cast<float4*>(base + kPlaneBytes)[tid]     = z;
        reinterpret_cast<float4*>(base + 2 * kPlaneBytes)[tid] = z;
      } else {
        uint8_t* base = m.pool + (size_t) (val & 0x7FFFFFFFu) * 768u;
        for (int w = tid; w < 192; w += 128)
          reinterpret_cast<uint32_t*>(base)[w] = 0u;
      }
      if (tid == 0) {
        atomicExch(m.keys + le.slot, kTomb);
        if (val >> 31) {
          const int addr       = atomicAdd(&m.ctr->heap_low_counter, 1);
          m.heap_low[addr + 1] = val & 0x7FFFFFFFu;
          atomicAdd(&m.ctr->low_live, (unsigned long long) -1ll);
        } else {
          const int addr   = atomicAdd(&m.ctr->heap_counter, 1);
          m.heap[addr + 1] = val;
          m.stats[val]     = {3.40282346638528859812e+38f, 0u};
        }
        m.live[live_cur][i].slot = kInvalid;
      }
    }
  }
}


  // One CTA per record: warp 0 inserts the key (no frustum test), then 128 threads scatter the
  // AoS voxels into the block's SoA planes (chunkToGlobalHashPass1/2, streamer.cu:250-329).
  __global__ void __launch_bounds__(128) k_insert_blocks(MapDev m, uint32_t live_cur, const GatherRecord* __restrict__ recs, const uint32_t* __restrict__ voxels_aos, uint32_t n) {
    __shared__ uint32_t s_val;
    const int tid = threadIdx.x, lane = tid & 31;
    CameraDev cam_unused{};
    PoseDev pose_unused{};
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
      const GatherRecord r = recs[i];
      if (tid < 32) {
        const uint32_t v = warp_insert<false>(m, cam_unused, pose_unused, live_cur, {r.x, r.y, r.z}, lane, r.resolution);
        if (lane == 0)
          s_val = v;
      }
      __syncthreads();
      uint32_t val = s_val;
      __syncthreads();
      // The key is live already: the block was paged out, and allocated again before it came back
      // (the reference ends up with two entries of one key there, tests/test_streamer.cu:40-117 bounds
      // their ratio). Here the stored observations are fused into the live block with the running
      // weighted mean of combineVoxel (voxel_hash_utils.cuh:169-181), so nothing is lost or doubled.
      bool merge = false;
      if (val == kInvalid && r.resolution == 0) {
        if (tid == 0) {
          const int sl = table_find(m, {r.x, r.y, r.z});
          s_val        = sl >= 0 ? m.vals[sl] : kInvalid;
          if (sl >= 0)
            atomicAdd(&m.ctr->stream_merged, 1ull);
        }
        __syncthreads();
        val = s_val;
        __syncthreads();
        merge = val != kInvalid && !(val >> 31);
        if (!merge)
          val = kInvalid;
      }
      if (val == kInvalid)
        continue;
      const uint32_t* src = voxels_aos + (size_t) i * kBlockVoxels * 3;
      if (!(val >> 31)) {
        uint8_t* base = m.pool + (size_t) val * kBlockBytes;
        float4 sdf4, ss4;
        uint4 cw4;
        float* sv    = reinterpret_cast<float*>(&sdf4);
        float* qv    = reinterpret_cast<float*>(&ss4);
        uint32_t* cv = reinterpret_cast<uint32_t*>(&cw4);
        float min_abs  = 3.40282346638528859812e+38f;
        uint32_t max_w = 0;
        if (merge) {
          sdf4 = reinterpret_cast<const float4*>(base)[tid];
          ss4  = reinterpret_cast<const float4*>(base + kPlaneBytes)[tid];
          cw4  = reinterpret_cast<const uint4*>(base + 2 * kPlaneBytes)[tid];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float s1    = __uint_as_float(src[(tid * 4 + j) * 3 + 0]);
          const float q1    = __uint_as_float(src[(tid * 4 + j) * 3 + 1]);
          const uint32_t c1 = src[(tid * 4 + j) * 3 + 2];
          const uint32_t w1 = c1 >> 24, w0 = merge ? cv[j] >> 24 : 0u;
          if (!merge || w0 == 0) {
            if (!merge || w1)
              sv[j] = s1, qv[j] = q1, cv[j] = c1;
          } else if (w1) {
            const uint32_t c0 = cv[j];
            sv[j]             = fdiv(ffma(s1, __uint2float_rn(w1), fmul(sv[j], __uint2float_rn(w0))), __uint2float_rn(w0 + w1));
            const uint32_t rr = (uint32_t) f2i(fadd(ffma(__uint2float_rn(c1 & 0xFF), 0.5f, fmul(__uint2float_rn(c0 & 0xFF), 0.5f)), 0.5f)) & 0xFF;
            const uint32_t gg = (uint32_t) f2i(fadd(ffma(__uint2float_rn((c1 >> 8) & 0xFF), 0.5f, fmul(__uint2float_rn((c0 >> 8) & 0xFF), 0.5f)), 0.5f)) & 0xFF;
            const uint32_t bb = (uint32_t) f2i(fadd(ffma(__uint2float_rn((c1 >> 16) & 0xFF), 0.5f, fmul(__uint2float_rn((c0 >> 16) & 0xFF), 0.5f)), 0.5f)) & 0xFF;
            cv[j]             = rr | (gg << 8) | (bb << 16) | (min(w0 + w1, (uint32_t) kWeightMax) << 24);
          }
          if (cv[j] >> 24)
            min_abs = fminf(min_abs, fabsf(sv[j]));
          max_w = max(max_w, cv[j] >> 24);
        }
        reinterpret_cast<float4*>(base)[tid]                  = sdf4;
        reinterpret_cast<float4*>(base + kPlaneBytes)[tid]    = ss4;
        reinterpret_cast<uint4*>(base + 2 * kPlaneBytes)[tid] = cw4;
        // GC statistics of the restored block
        __shared__ float s_min[4];
        __shared__ uint32_t s_max[4];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          min_abs = fminf(min_abs, __shfl_xor_sync(0xFFFFFFFFu, min_abs, o));
          max_w   = max(max_w, __shfl_xor_sync(0xFFFFFFFFu, max_w, o));
        }
        if (lane == 0)
          s_min[tid >> 5] = min_abs, s_max[tid >> 5] = max_w;
        __syncthreads();
        if (tid == 0)
          m.stats[val] = {fminf(fminf(s_min[0], s_min[1]), fminf(s_min[2], s_min[3])), max(max(s_max[0], s_max[1]), max(s_max[2], s_max[3]))};
        __syncthreads();
      } else if (tid < 64) {
        uint8_t* base = m.pool + (size_t) (val & 0x7FFFFFFFu) * 768u;
        reinterpret_cast<uint32_t*>(base)[tid]       = src[tid * 3 + 0];
        reinterpret_cast<uint32_t*>(base + 256)[tid] = src[tid * 3 + 1];
        reinterpret_cast<uint32_t*>(base + 512)[tid] = src[tid * 3 + 2];
      }
    }
  }

  // Gather every live block of the device map into host vectors (appended).
  int gather_to_host(mrh_map* m, std::vector<GatherRecord>& recs, VoxelWords& voxels, const float* far_centre, float far_radius) {
    FarFilter ff{0.f, 0.f, 0.f, 0.f, 0};
    if (far_centre)
      ff = {far_centre[0], far_centre[1], far_centre[2], far_radius, 1};
    mrh_stats st;
    if (mrh_get_stats(m, &st))
      return 1;
    const size_t n = (size_t) st.live_blocks;
    if (n == 0)
      return 0;
    // bounded staging so that a 100 GB map does not need a 100 GB mirror on the device
    const size_t chunk = std::min<size_t>(n, 1u << 16);
    ScopedDev<GatherRecord> recs_buf;
    ScopedDev<uint32_t> vox_buf, count_buf; // released on every return path
    CK(recs_buf.alloc(chunk));
    CK(vox_buf.alloc((size_t) 3 * kBlockVoxels * chunk));
    CK(count_buf.alloc(1));
    GatherRecord* d_recs = recs_buf.p;
    uint32_t* d_vox      = vox_buf.p;
    uint32_t* d_count    = count_buf.p;
    const size_t base = recs.size();
    recs.resize(base + n);
    voxels.resize((base + n) * 3 * kBlockVoxels);
    const uint32_t n_list = m->h_ctr->live_count[m->live_cur]; // refreshed by mrh_get_stats above
    size_t done           = 0;
    for (uint32_t first = 0; first < n_list; first += (uint32_t) chunk) {
      CK(cudaMemsetAsync(d_count, 0, sizeof(uint32_t), m->stream));
      k_gather_blocks<<<m->num_sms * 8, 128, 0, m->stream>>>(m->dev, m->live_cur, d_recs, d_vox, d_count, first, (uint32_t) chunk, ff);
      m->launches++;
      uint32_t got = 0;
      CK(cudaMemcpyAsync(&got, d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
      CK(cudaStreamSynchronize(m->stream));
      if (got == 0)
        continue;
      if (done + got > n)
        return fail("gather_to_host: live list holds more blocks than the heap accounts for");
      CK(cudaMemcpy(recs.data() + base + done, d_recs, sizeof(GatherRecord) * got, cudaMemcpyDeviceToHost));
      if (bulk_d2h(m, voxels.data() + (base + done) * 3 * kBlockVoxels, d_vox, sizeof(uint32_t) * 3 * kBlockVoxels * got))
        return 1;
      done += got;
    }
    recs.resize(base + done);
    voxels.resize((base + done) * 3 * kBlockVoxels);
    return 0;
  }

  // ---- bulk transfers between device memory and PAGEABLE host vectors ----------------------------
  // cudaMemcpy on pageable memory goes through the driver's own staging at a few GB/s on one thread.
  // Here: 16 MB pieces, DMA into / out of two pinned bounce buffers, and the host side of piece k
  // (8 threads, 256 KB slices) overlaps the DMA of piece k+1.
  constexpr size_t kBounceBytes = 16u << 20;

  void parallel_copy(void* dst, const void* src, size_t bytes) {
    const size_t slice = 256u << 10;
    const long n       = (long) ((bytes + slice - 1) / slice);
    const int nthr     = (int) std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
#pragma omp parallel for num_threads(nthr) schedule(static) if (n > 2)
    for (long i = 0; i < n; ++i) {
      const size_t o = (size_t) i * slice;
      memcpy((char*) dst + o, (const char*) src + o, std::min(slice, bytes - o));
    }
  }

  static int bounce_ready(mrh_map* m) {
    for (int i = 0; i < 2; ++i)
      if (!m->h_bounce[i]) {
        CK(cudaMallocHost(&m->h_bounce[i], kBounceBytes));
        CK(cudaEventCreateWithFlags(&m->ev_bounce[i], cudaEventDisableTiming));
      }
    return 0;
  }

  int bulk_d2h(mrh_map* m, void* dst_host, const void* src_dev, size_t bytes) {
    if (bounce_ready(m))
      return 1;
    const size_t n = (bytes + kBounceBytes - 1) / kBounceBytes;
    for (size_t k = 0; k <= n; ++k) {
      if (k < n) { // start the DMA of piece k
        const size_t o = k * kBounceBytes;
        CK(cudaMemcpyAsync(m->h_bounce[k & 1], (const char*) src_dev + o, std::min(kBounceBytes, bytes - o), cudaMemcpyDeviceToHost, m->stream));
        CK(cudaEventRecord(m->ev_bounce[k & 1], m->stream));
      }
      if (k > 0) { // while it runs, move piece k-1 to its destination
        const size_t o = (k - 1) * kBounceBytes;
        CK(cudaEventSynchronize(m->ev_bounce[(k - 1) & 1]));
        parallel_copy((char*) dst_host + o, m->h_bounce[(k - 1) & 1], std::min(kBounceBytes, bytes - o));
      }
    }
    return 0;
  }

  int bulk_h2d(mrh_map* m, void* dst_dev, const void* src_host, size_t bytes) {
    if (bounce_ready(m))
      return 1;
    const size_t n = (bytes + kBounceBytes - 1) / kBounceBytes;
    for (size_t k = 0; k < n; ++k) {
      const size_t o = k * kBounceBytes, len = std::min(kBounceBytes, bytes - o);
      if (k >= 2)
        CK(cudaEventSynchronize(m->ev_bounce[k & 1])); // the DMA that last read this buffer has finished
      parallel_copy(m->h_bounce[k & 1], (const char*) src_host + o, len);
      CK(cudaMemcpyAsync((char*) dst_dev + o, m->h_bounce[k & 1], len, cudaMemcpyHostToDevice, m->stream));
      CK(cudaEventRecord(m->ev_bounce[k & 1], m->stream));
    }
    return 0; // the caller's next stream operation (or synchronise) orders after the copies
  }

  // one pass: records with pairwise different keys
  static int insert_pass(mrh_map* m, const GatherRecord* recs, const uint32_t* voxels, size_t n) {
    const size_t chunk   = std::min<size_t>(n, 1u << 16);
    ScopedDev<GatherRecord> recs_buf;
    ScopedDev<uint32_t> vox_buf; // released on every return path
    CK(recs_buf.alloc(chunk));
    CK(vox_buf.alloc((size_t) 3 * kBlockVoxels * chunk));
    GatherRecord* d_recs = recs_buf.p;
    uint32_t* d_vox      = vox_buf.p;
    for (size_t first = 0; first < n; first += chunk) {
      const size_t k = std::min(chunk, n - first);
      CK(cudaMemcpyAsync(d_recs, recs + first, sizeof(GatherRecord) * k, cudaMemcpyHostToDevice, m->stream));
      if (bulk_h2d(m, d_vox, voxels + first * 3 * kBlockVoxels, sizeof(uint32_t) * 3 * kBlockVoxels * k))
        return 1;
      k_insert_blocks<<<m->num_sms * 8, 128, 0, m->stream>>>(m->dev, m->live_cur, d_recs, d_vox, (uint32_t) k);
      m->launches++;
      CK(cudaStreamSynchronize(m->stream));
    }
    return 0;
  }

  // Streamer::streamInToGPU (streamer.cpp:358-378) for a selection of host-store records.
  static int insert_from_host_unchecked(mrh_map* m, const GatherRecord* recs, const uint32_t* voxels, size_t n);

  // k_insert_blocks skips a record whose block cannot be placed (pool or table full) and only counts
  // it; a caller that then dropped the record from the host store would lose the block for good.
  // The counters are compared around the insert and a shortfall is an error: the callers leave the
  // host store untouched in that case.
  int insert_from_host(mrh_map* m, const GatherRecord* recs, const uint32_t* voxels, size_t n) {
    if (n == 0)
      return 0;
    mrh_stats s0, s1;
    if (mrh_get_stats(m, &s0) || insert_from_host_unchecked(m, recs, voxels, n) || mrh_get_stats(m, &s1))
      return 1;
    const uint64_t lost = (s1.dropped_heap - s0.dropped_heap) + (s1.dropped_table - s0.dropped_table);
    if (lost)
      return fail("stream-in: %llu of %zu blocks did not fit on the device (free pool blocks before: %lld, table buckets: %llu); "
                  "the host store keeps every record - enlarge num_sdf_blocks / hash_num_buckets or page with a smaller radius",
                  (unsigned long long) lost, n, (long long) s0.heap_free, (unsigned long long) m->hash_num_buckets);
    return 0;
  }

  static int insert_from_host_unchecked(mrh_map* m, const GatherRecord* recs, const uint32_t* voxels, size_t n) {
    uint32_t n_low = 0;
    for (size_t i = 0; i < n; ++i)
      n_low += recs[i].resolution != 0;
    if (carve_low_blocks(m, n_low))
      return 1;
    // Records that repeat a key (a block paged out, allocated again, paged out again) must not race
    // with the record that creates the block: the k-th copy of a key goes into pass k.
    std::unordered_map<unsigned long long, uint32_t> seen;
    seen.reserve(2 * n);
    std::vector<uint32_t> pass(n);
    uint32_t n_pass = 1;
    for (size_t i = 0; i < n; ++i) {
      const unsigned long long key = ((unsigned long long) (uint32_t) (recs[i].x + kCoordBias) << 42) | ((unsigned long long) (uint32_t) (recs[i].y + kCoordBias) << 21) | (unsigned long long) (uint32_t) (recs[i].z + kCoordBias);
      pass[i] = seen[key]++;
      n_pass  = std::max(n_pass, pass[i] + 1);
    }
    if (n_pass == 1)
      return insert_pass(m, recs, voxels, n);
    const size_t words = (size_t) 3 * kBlockVoxels;
    for (uint32_t p = 0; p < n_pass; ++p) {
      std::vector<GatherRecord> r;
      VoxelWords v;
      for (size_t i = 0; i < n; ++i)
        if (pass[i] == p) {
          r.push_back(recs[i]);
          v.insert(v.end(), voxels + i * words, voxels + (i + 1) * words);
        }
      if (insert_pass(m, r.data(), v.data(), r.size()))
        return 1;
    }
    return 0;
  }

  // Streamer::worldToChunks (streamer.cuh:251-260) of a block origin (streamer.cpp:231-233), then
  // isChunkInSphere (streamer.cuh:346-352) of that chunk
  bool record_in_sphere(const mrh_map* m, const GatherRecord& r, const float centre[3], float radius) {
    const float size = m->p.virtual_voxel_size, ext = (float) m->p.voxel_extents_scale;
    const float chunk_radius = 0.5f * ext * std::sqrt(3.f); // streamer.cpp:15
    const int pos[3] = {r.x, r.y, r.z};
    float d[3];
    for (int k = 0; k < 3; ++k) {
      const float pw = ((float) pos[k] * 8.f) * size;
      const float p  = pw / ext;
      const float sg = (float) ((0.f < p) - (p < 0.f));
      const int c    = (int) (p + sg * 0.5f);
      d[k]           = (float) c * ext - centre[k];
    }
    const float l = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    return l <= std::fabs(radius - chunk_radius);
  }

  int stream_radius(mrh_map* m, const float centre[3], float radius) {
    HostStore& st = m->store;
    // stream out (streamOutToHostPass0 + integrateInChunkGrid): blocks >= radius away go to the store
    const size_t before = st.recs.size();
    if (gather_to_host(m, st.recs, st.voxels, centre, radius))
      return 1;
    m->last_stream_out = st.recs.size() - before;
    // stream in (streamInToGPU / integrateInHash): every stored block whose chunk lies in the sphere
    std::vector<GatherRecord> in_recs, keep_recs;
    VoxelWords in_vox, keep_vox;
    const size_t words = (size_t) 3 * kBlockVoxels;
    for (size_t i = 0; i < st.recs.size(); ++i) {
      const bool in = i < before && record_in_sphere(m, st.recs[i], centre, radius); // what just left stays out
      (in ? in_recs : keep_recs).push_back(st.recs[i]);
      VoxelWords& vv = in ? in_vox : keep_vox;
      vv.insert(vv.end(), st.voxels.begin() + i * words, st.voxels.begin() + (i + 1) * words);
    }
    m->last_stream_in = in_recs.size();
    if (!in_recs.empty()) {
      mrh_stats s0, s1;
      if (mrh_get_stats(m, &s0) || insert_from_host(m, in_recs.data(), in_vox.data(), in_recs.size()) || mrh_get_stats(m, &s1))
        return 1;
      // keys that were allocated again while their old block sat in the store: fused (k_insert_blocks)
      m->stream_duplicates += in_recs.size() - (size_t) (s1.blocks_new - s0.blocks_new);
      st.recs.swap(keep_recs);
      st.voxels.swap(keep_vox);
    }
    m->stream_events++;
    m->counters_clean = false;
    return 0;
  }

} // namespace mrh

extern "C" {

int mrh_stream(mrh_map* m, const float centre[3], float radius) {
  if (!m || !centre)
    return fail("null argument");
  CK(cudaSetDevice(m->device));
  return stream_radius(m, centre, radius);
}

int mrh_dump_state(mrh_map* m, mrh_dump_entry* entries, void* voxels, size_t max_entries, size_t* n_out) {
  if (!m || !n_out)
    return fail("null argument");
  CK(cudaSetDevice(m->device));
  std::vector<GatherRecord> recs;
  VoxelWords vox;
  if (gather_to_host(m, recs, vox))
    return 1;
  *n_out = recs.size();
  if (!entries)
    return 0;
  std::vector<size_t> order(recs.size());
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
    const GatherRecord &p = recs[a], &q = recs[b];
    if (p.x != q.x)
      return p.x < q.x;
    if (p.y != q.y)
      return p.y < q.y;
    return p.z < q.z;
  });
  const size_t n = std::min(max_entries, recs.size());
  for (size_t i = 0; i < n; ++i) {
    const GatherRecord& r = recs[order[i]];
    entries[i]            = {r.x, r.y, r.z, r.resolution, r.ptr};
    if (voxels)
      memcpy((uint8_t*) voxels + i * 12 * kBlockVoxels, vox.data() + order[i] * 3 * kBlockVoxels, 12 * kBlockVoxels);
  }
  return 0;
}

int mrh_stream_all_out(mrh_map* m) {
  if (!m)
    return fail("null handle");
  CK(cudaSetDevice(m->device));
  if (gather_to_host(m, m->store.recs, m->store.voxels))
    return 1;
  const uint32_t frame = m->frame_index;
  if (reset_map(m))
    return 1;
  m->frame_index = frame; // num_integrated_frames_ survives streaming
  CK(cudaStreamSynchronize(m->stream));
  return 0;
}

int mrh_store_append(mrh_map* m, const mrh_dump_entry* entries, const void* voxels, size_t n) {
  if (!m || (n && (!entries || !voxels)))
    return fail("null argument");
  HostStore& st     = m->store;
  const size_t base = st.recs.size();
  st.recs.resize(base + n);
  st.voxels.resize((base + n) * 3 * kBlockVoxels);
  for (size_t i = 0; i < n; ++i)
    st.recs[base + i] = {entries[i].x, entries[i].y, entries[i].z, entries[i].resolution, entries[i].ptr};
  memcpy(st.voxels.data() + base * 3 * kBlockVoxels, voxels, n * 12 * kBlockVoxels);
  return 0;
}

int mrh_store_read(mrh_map* m, mrh_dump_entry* entries, void* voxels, size_t max_entries, size_t* n_out) {
  if (!m || !n_out)
    return fail("null argument");
  const HostStore& st = m->store;
  *n_out              = st.recs.size();
  if (!entries)
    return 0;
  const size_t n = std::min(max_entries, st.recs.size());
  for (size_t i = 0; i < n; ++i)
    entries[i] = {st.recs[i].x, st.recs[i].y, st.recs[i].z, st.recs[i].resolution, st.recs[i].ptr};
  if (voxels && n)
    memcpy(voxels, st.voxels.data(), n * 12 * kBlockVoxels);
  return 0;
}

int mrh_store_size(mrh_map* m, size_t* n) {
  if (!m || !n)
    return fail("null argument");
  *n = m->store.size();
  return 0;
}

// Streamer::serializeData (streamer.cpp:104-160) over the host store.
int mrh_serialize_data(mrh_map* m, const char* hash_path, const char* voxel_path) {
  if (!m || !hash_path || !voxel_path)
    return fail("null argument");
  const HostStore& s = m->store;
  const float size   = m->p.virtual_voxel_size;
  const size_t nrec  = s.recs.size();
  // pass 1: counts per record
  std::vector<uint32_t> count(nrec + 1, 0);
  const unsigned nthr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  auto parallel = [&](auto&& fn) {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthr; ++t)
      th.emplace_back([&, t] {
        for (size_t k = t; k < nrec; k += nthr)
          fn(k);
      });
    for (auto& x : th)
      x.join();
  };
  parallel([&](size_t k) {
    const int nv        = s.recs[k].resolution == 0 ? 512 : 64;
    const uint32_t* vox = s.voxels.data() + k * 3 * kBlockVoxels;
    uint32_t c          = 0;
    for (int l = 0; l < nv; ++l)
      c += (vox[3 * l + 2] >> 24) != 0;
    count[k + 1] = c;
  });
  std::vector<uint64_t> vofs(nrec + 1, 0), hofs(nrec + 1, 0);
  for (size_t k = 0; k < nrec; ++k) {
    vofs[k + 1] = vofs[k] + count[k + 1];
    hofs[k + 1] = hofs[k] + (count[k + 1] ? 1 : 0);
  }
  const uint64_t n_vox = vofs[nrec], n_hash = hofs[nrec];
  // records: voxel point = x,y,z,sdf,weight,rgba (24 B); hash point = x,y,z,weight,rgba (20 B)
  // every byte is written by the parallel fill below: no zero fill of the (GB-sized) buffers
  std::vector<uint8_t, NoInitAlloc<uint8_t>> vbuf(n_vox * 24), hbuf(n_hash * 20);
  parallel([&](size_t k) {
    if (!count[k + 1])
      return;
    const GatherRecord& r = s.recs[k];
    const int res         = r.resolution;
    const int nv          = res == 0 ? 512 : 64;
    const int sf          = 1 << res;
    const int bs          = 8 / sf;
    const uint32_t* vox   = s.voxels.data() + k * 3 * kBlockVoxels;
    const float bw[3]     = {(float) (r.x * 8) * size, (float) (r.y * 8) * size, (float) (r.z * 8) * size};
    const uint8_t col[4]  = {(uint8_t) (res == 0 ? 255 : 0), (uint8_t) (res == 1 ? 255 : 0), 0, 0};
    uint8_t* vp           = vbuf.data() + vofs[k] * 24;
    float wsum            = 0.f;
    float csum[4]         = {0, 0, 0, 0};
    uint32_t valid        = 0;
    for (int l = 0; l < nv; ++l) {
      const uint32_t cw = vox[3 * l + 2];
      if (!(cw >> 24))
        continue;
      const int dl[3]  = {(l % bs) * sf, ((l % (bs * bs)) / bs) * sf, (l / (bs * bs)) * sf};
      const float p[3] = {bw[0] + (float) dl[0] * size, bw[1] + (float) dl[1] * size, bw[2] + (float) dl[2] * size};
      const float w    = (float) (cw >> 24);
      memcpy(vp, p, 12);
      memcpy(vp + 12, &vox[3 * l], 4); // sdf
      memcpy(vp + 16, &w, 4);
      memcpy(vp + 20, col, 4);
      vp += 24;
      wsum += w;
      csum[0] += res == 0 ? 1.f : 0.f;
      csum[1] += res == 1 ? 1.f : 0.f;
      valid++;
    }
    uint8_t* hp       = hbuf.data() + hofs[k] * 20;
    const float avg_w = wsum / (float) valid;
    const uint8_t hc[4] = {(uint8_t) (csum[0] / (float) valid * 255), (uint8_t) (csum[1] / (float) valid * 255), 0, 0};
    memcpy(hp, bw, 12);
    memcpy(hp + 12, &avg_w, 4);
    memcpy(hp + 16, hc, 4);
  });
  auto write_ply = [&](const char* path, uint64_t n, bool with_sdf, const std::vector<uint8_t, NoInitAlloc<uint8_t>>& data) -> int {
    if (n == 0) {
      fprintf(stderr, "PointCloudSerializer|empty point cloud\n");
      return 0;
    }
    const std::string p(path);
    if (p.size() < 4 || p.substr(p.size() - 4) != ".ply")
      return fail("PointCloudSerializer|unknown file extension%s", path);
    FILE* f = fopen(path, "wb");
    if (!f)
      return fail("Could not open file for writing %s", path);
    fprintf(f, "ply\nformat binary_little_endian 1.0\nelement vertex %llu\nproperty float x\nproperty float y\nproperty float z\n", (unsigned long long) n);
    if (with_sdf)
      fprintf(f, "property float sdf\n");
    fprintf(f, "property float weight\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\nend_header\n");
    fwrite(data.data(), 1, data.size(), f);
    fclose(f);
    return 0;
  };
  if (write_ply(hash_path, n_hash, false, hbuf) || write_ply(voxel_path, n_vox, true, vbuf))
    return 1;
  printf("Streamer::serializeData | written %llu hash points and %llu voxels to %s and %s\n", (unsigned long long) n_hash, (unsigned long long) n_vox, hash_path, voxel_path);
  return 0;
}

} // extern "C"

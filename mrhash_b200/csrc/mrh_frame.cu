// mrh_frame.cu — per-frame launch sequence of the integration path.
//
// VoxelContainer::integrate (voxel_data_structures.cpp:90-134) issues ~12 kernels, ~10 blocking
// copies and ~15 device synchronisations per frame. Here a frame is 3 kernels on one stream
// (5 more on the every-n-th starve frame) and the host never waits.
#include <algorithm>
#include <cmath>

#include <cub/device/device_radix_sort.cuh>

#include "mrh_host.h"
#include "mrh_kernels.cuh"
#include "mrh_points.cuh"

namespace mrh {

#define CKL()                                                                                      \
  do {                                                                                             \
    cudaError_t e_ = cudaGetLastError();                                                           \
    if (e_ != cudaSuccess)                                                                         \
      return fail("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_));               \
  } while (0)

FrameDev make_frame(const mrh_map* m) {
  FrameDev f;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      f.R[i * 3 + j] = m->pose[i * 4 + j];
    f.t[i] = m->pose[i * 4 + 3];
  }
  f.frame_index = m->frame_index;
  f.live_cur    = m->live_cur;
  f.pad[0] = f.pad[1] = 0;
  return f;
}

static int gc_tail(mrh_map* m, const FrameDev& f, bool starve);

int integrate_rgbd(mrh_map* m) {
  const MapDev& d    = m->dev;
  const CameraDev& c = m->cam;
  const FrameDev f   = make_frame(m);
  cudaStream_t s     = m->stream;
  const int n_gc     = m->p.n_frames_invalidate_voxels;
  const bool gc      = n_gc > 0; // voxel_data_structures.cpp:105
  const bool starve  = gc && m->frame_index > 0 && (m->frame_index % (uint32_t) n_gc) == 0; // :140
  const bool var     = m->p.sdf_var_threshold > 0.f;

  if (var)
    return fail("sdf_var_threshold > 0 is not wired up yet");

  const bool prof = m->profiling;
  auto mark = [&](int i) {
    if (prof)
      cudaEventRecord(m->ev_k[i], s);
  };
  const dim3 grid_alloc((c.cols + 31) / 32, (c.rows + 7) / 8);
  mark(0);
  k_alloc_rgbd<<<grid_alloc, 256, 0, s>>>(d, f, c, m->depth_ptr);
  CKL();
  mark(1);
  k_visible<<<m->num_sms * 4, 256, 0, s>>>(d, f, c, 1);
  CKL();
  mark(2);
  m->launches += 2;
  const int grid_blocks = m->num_sms * 8;
  if (gc && !starve) {
    k_integrate<true><<<grid_blocks, 128, 0, s>>>(d, f, c, m->depth_ptr, m->rgb_ptr);
    CKL();
    mark(3);
    m->launches += 1;
  } else {
    k_integrate<false><<<grid_blocks, 128, 0, s>>>(d, f, c, m->depth_ptr, m->rgb_ptr);
    CKL();
    mark(3);
    m->launches += 1;
    if (starve && gc_tail(m, f, true))
      return 1;
  }
  if (prof) {
    // profiling pass only: wait for the frame and accumulate the per-kernel device times
    if (cudaEventSynchronize(m->ev_k[3]) != cudaSuccess)
      return fail("profiling: event synchronize failed");
    for (int i = 0; i < 3; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, m->ev_k[i], m->ev_k[i + 1]);
      m->kernel_ms[i] += ms;
      m->kernel_launches[i] += 1;
    }
  }
  m->live_cur ^= 1u;
  m->frame_index++;
  m->frames_total++;
  return 0;
}

// GC tail shared by both sensor paths when the fused kernel cannot be used
// (voxel_data_structures.cpp:137-145: [starve], identify, free)
static int gc_tail(mrh_map* m, const FrameDev& f, bool starve) {
  const MapDev& d    = m->dev;
  const CameraDev& c = m->cam;
  cudaStream_t s     = m->stream;
  const int grid     = m->num_sms * 8;
  if (starve) {
    if (cudaMemsetAsync(d.zbuf, 0xFF, sizeof(unsigned long long) * c.rows * c.cols, s) != cudaSuccess)
      return fail("memset zbuf failed");
    k_starve<0><<<grid, 128, 0, s>>>(d, f, c);
    k_starve<1><<<grid, 128, 0, s>>>(d, f, c);
    m->launches += 2;
  }
  k_identify<<<grid, 128, 0, s>>>(d);
  k_gc_free<<<grid, 128, 0, s>>>(d, f);
  m->launches += 2;
  CKL();
  return 0;
}

int integrate_points(mrh_map* m) {
  const MapDev& d    = m->dev;
  const CameraDev& c = m->cam;
  const FrameDev f   = make_frame(m);
  cudaStream_t s     = m->stream;
  const int n_gc     = m->p.n_frames_invalidate_voxels;
  const bool gc      = n_gc > 0;
  const bool starve  = gc && m->frame_index > 0 && (m->frame_index % (uint32_t) n_gc) == 0;
  const bool var     = m->p.sdf_var_threshold > 0.f;
  const uint32_t n   = (uint32_t) m->n_points;
  if (var)
    return fail("sdf_var_threshold > 0 is not wired up yet");
  if (m->n_points >= (1ull << kPointIdxBits))
    return fail("point cloud too large: %zu points (limit %u)", m->n_points, 1u << kPointIdxBits);

  // staging for the (voxel, point, sdf) records: a ray of length 2t crosses at most 3*(2t/size)+4 voxels
  const float t_max   = m->p.sdf_truncation + m->p.sdf_truncation_scale * m->max_integration_distance;
  const size_t per_pt = (size_t) std::min(3.0 * std::ceil(2.0 * t_max / m->p.virtual_voxel_size) + 4.0, 256.0);
  const size_t want   = std::min<size_t>((size_t) n * per_pt, (size_t) 1 << 28);
  if (want > m->upd_cap) {
    cudaStreamSynchronize(s);
    for (int i = 0; i < 2; ++i) {
      cudaFree(m->d_upd_keys[i]), cudaFree(m->d_upd_vals[i]);
      if (cudaMalloc(&m->d_upd_keys[i], sizeof(unsigned long long) * want) != cudaSuccess || cudaMalloc(&m->d_upd_vals[i], sizeof(float) * want) != cudaSuccess)
        return fail("out of device memory for %zu point-update records", want);
    }
    m->upd_cap = want;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, m->d_upd_keys[0], m->d_upd_keys[1], m->d_upd_vals[0], m->d_upd_vals[1], (int) want, 0, 64, s);
    cudaFree(m->d_sort_tmp);
    if (cudaMalloc(&m->d_sort_tmp, tmp) != cudaSuccess)
      return fail("out of device memory for the sort workspace");
    m->sort_tmp_bytes = tmp;
  }

  const int grid_pts = (int) ((n + 255) / 256);
  k_alloc_points<<<grid_pts, 256, 0, s>>>(d, f, c, m->d_points, n);
  CKL();
  k_visible<<<m->num_sms * 4, 256, 0, s>>>(d, f, c, 0);
  CKL();
  k_points_emit<<<grid_pts, 256, 0, s>>>(d, f, m->d_points, n, m->d_upd_keys[0], m->d_upd_vals[0], (uint32_t) m->upd_cap);
  CKL();
  m->launches += 3;
  // the sort needs the record count on the host: one 4-byte read-back per frame
  if (cudaMemcpyAsync(m->h_n_updates, &d.ctr->n_updates, sizeof(uint32_t), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
    return fail("point path: reading the record count failed: %s", cudaGetErrorString(cudaGetLastError()));
  const uint32_t n_upd = (uint32_t) std::min<size_t>(*m->h_n_updates, m->upd_cap);
  const unsigned long long* keys = m->d_upd_keys[0];
  const float* vals              = m->d_upd_vals[0];
  if (n_upd > 1) {
    int addr_bits = 1;
    while ((1ull << addr_bits) < (unsigned long long) d.num_blocks * 512ull)
      ++addr_bits;
    size_t tmp = m->sort_tmp_bytes;
    if (cub::DeviceRadixSort::SortPairs(m->d_sort_tmp, tmp, m->d_upd_keys[0], m->d_upd_keys[1], m->d_upd_vals[0], m->d_upd_vals[1], (int) n_upd, 0, kPointIdxBits + addr_bits, s) != cudaSuccess)
      return fail("point path: radix sort failed");
    keys = m->d_upd_keys[1], vals = m->d_upd_vals[1];
    m->launches += 4;
  }
  k_points_apply<<<m->num_sms * 4, 256, 0, s>>>(d, keys, vals, (uint32_t) m->upd_cap);
  CKL();
  m->launches += 1;
  if (gc && gc_tail(m, f, starve))
    return 1;
  m->live_cur ^= 1u;
  m->frame_index++;
  m->frames_total++;
  return 0;
}

} // namespace mrh

// mrh_frame.cu — per-frame launch sequence of the integration path.
//
// VoxelContainer::integrate (voxel_data_structures.cpp:90-134) issues ~12 kernels, ~10 blocking
// copies and ~15 device synchronisations per frame. Here a frame is 3 kernels on one stream
// (5 more on the every-n-th starve frame) and the host never waits.
#include "mrh_host.h"
#include "mrh_kernels.cuh"

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
    if (starve) {
      if (cudaMemsetAsync(d.zbuf, 0xFF, sizeof(unsigned long long) * c.rows * c.cols, s) != cudaSuccess)
        return fail("memset zbuf failed");
      k_starve<0><<<grid_blocks, 128, 0, s>>>(d, f, c);
      k_starve<1><<<grid_blocks, 128, 0, s>>>(d, f, c);
      k_identify<<<grid_blocks, 128, 0, s>>>(d);
      k_gc_free<<<grid_blocks, 128, 0, s>>>(d, f);
      CKL();
      m->launches += 4;
    }
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

int integrate_points(mrh_map* m) {
  (void) m;
  return fail("LiDAR integration is not wired up yet");
}

} // namespace mrh

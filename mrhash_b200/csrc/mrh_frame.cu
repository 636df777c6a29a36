// mrh_frame.cu — per-frame launch sequence of the integration path.
//
// VoxelContainer::integrate (voxel_data_structures.cpp:90-134) issues ~12 kernels, ~10 blocking
// copies and ~15 device synchronisations per frame. Here an RGB-D frame is 3 kernels on one stream
// (4 more on the every-n-th starve frame) and the host never waits, on the point-cloud path either
// (its sort runs over a record count the host knows in advance).
#include <algorithm>
#include <cmath>
#include <utility>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "mrh_host.h"
#include "mrh_fast.cuh"
#include "mrh_fused.cuh"
#include "mrh_kernels.cuh"
#include "mrh_points.cuh"
#include "mrh_var.cuh"

namespace mrh {

#define CKL()                                                                                      \
  do {                                                                                             \
    cudaError_t e_ = cudaGetLastError();                                                           \
    if (e_ != cudaSuccess)                                                                         \
      return fail("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_));               \
  } while (0)

int integrate_ctas_per_sm() {
  return kIntVox == 2 ? 6 : 9;
}

FrameDev make_frame(const mrh_map* m) {
  FrameDev f;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      f.R[i * 3 + j] = m->pose[i * 4 + j];
    f.t[i] = m->pose[i * 4 + 3];
  }
  // CUDAMatSE3::inverse() (cuda_algebra.cuh:137-143) as pose_finish (mrh_math.cuh) evaluates it on the
  // device: R^T, then -(R^T t) with row_dot's contraction fma(c, z, fma(a, x, b * y)). fmaf is
  // correctly rounded on the host too, so the bits are the same.
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      f.Ri[j * 3 + i] = f.R[i * 3 + j];
  for (int i = 0; i < 3; ++i) {
    volatile float by = f.Ri[i * 3 + 1] * f.t[1];
    f.ti[i]           = -std::fmaf(f.Ri[i * 3 + 2], f.t[2], std::fmaf(f.Ri[i * 3 + 0], f.t[0], by));
  }
  f.frame_index = m->frame_index;
  f.live_cur    = m->live_cur;
  // paging probe: every frame while paging is enabled, as the reference looks at the free count in
  // front of every integrate (geowrapper.cpp:137). The write to host memory delays the end of the
  // kernel by ~1.5 us; a map that never pages (stream_threshold = 0) does not pay it.
  f.pad[0] = m->stream_threshold > 0.f ? 1u : 0u;
  f.pad[1] = 0;
  return f;
}

namespace {

  // Launch with programmatic stream serialisation (programmatic dependent launch): the grid is staged
  // while the previous kernel of the stream drains, and its CTAs block in cudaGridDependencySynchronize()
  // (first statement of k_front / k_integrate) until that kernel has completed and flushed - the launch
  // latency between the two kernels of a frame, and between frames, disappears from the critical path.
  template <typename... KArgs, typename... Args>
  cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, bool chained, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = grid;
    cfg.blockDim           = block;
    cfg.dynamicSmemBytes   = 0;
    cfg.stream             = s;
    cudaLaunchAttribute attr[1];
    attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs                                          = attr;
    cfg.numAttrs                                       = chained ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
  }

  struct FrameCtx {
    mrh_map* m;
    FrameDev f;
    bool gc, starve, var, var_active;
    int grid_blocks, grid_list;
  };

  FrameCtx begin_frame(mrh_map* m) {
    FrameCtx c;
    c.m            = m;
    c.f            = make_frame(m);
    const int n_gc = m->p.n_frames_invalidate_voxels;
    c.gc           = n_gc > 0;                                                                 // voxel_data_structures.cpp:105
    c.starve       = c.gc && m->frame_index > 0 && (m->frame_index % (uint32_t) n_gc) == 0;    // :140
    c.var          = m->p.sdf_var_threshold > 0.f;
    c.var_active   = c.var && m->frame_index > 0;                                              // :99
    c.grid_blocks  = m->num_sms * 8;
    c.grid_list    = m->num_sms * 4;
    return c;
  }

  // allocBlocks :885-891 / allocBlocks3D :1050-1056
  int top_up_low_heap(FrameCtx& c) {
    mrh_map* m = c.m;
    const uint32_t low_blocks = (uint32_t) ((float) m->num_sdf_blocks * 0.1f); // voxel_data_structures.cuh:60
    k_carve_decide<<<1, 1, 0, m->stream>>>(m->dev, low_blocks);
    k_carve_low<<<c.grid_list, 256, 0, m->stream>>>(m->dev, 0u);
    m->launches += 2;
    CKL();
    return 0;
  }

  // GC tail when the fused kernel cannot be used (voxel_data_structures.cpp:137-145:
  // [starve], identify, free). On a sharded map the z-buffer of a starve frame holds only this rank's
  // voxels; with split_zbuf the frame pauses after the z-buffer pass so that the caller can min-reduce
  // it over the ranks (the front-most voxel of a pixel is then the front-most of the WHOLE map, as in
  // the unsharded run) and finish_gc_tail resumes.
  int gc_tail_finish(mrh_map* m, const FrameDev& f, bool starve, bool var) {
    const MapDev& d    = m->dev;
    cudaStream_t s     = m->stream;
    const int grid_blocks = m->num_sms * 8, grid_list = m->num_sms * 4;
    if (starve) {
      k_starve<1><<<grid_blocks, 128, 0, s>>>(d, f, m->cam);
      m->launches += 1;
    }
    k_identify<<<grid_blocks, 128, 0, s>>>(d);
    k_gc_free<<<grid_blocks, 128, 0, s>>>(d, f);
    m->launches += 2;
    if (var) {
      k_gc_low<<<grid_list, 128, 0, s>>>(d, f);
      m->launches += 1;
    }
    CKL();
    return 0;
  }

  int gc_tail(FrameCtx& c, const FrameDev& f) {
    mrh_map* m         = c.m;
    const MapDev& d    = m->dev;
    const CameraDev& k = m->cam;
    cudaStream_t s     = m->stream;
    if (c.starve) {
      // 0x7F bytes: above every packed (depth, id) value and still positive as int64 (NCCL MIN)
      if (cudaMemsetAsync(d.zbuf, 0x7F, sizeof(unsigned long long) * k.rows * k.cols, s) != cudaSuccess)
        return fail("memset zbuf failed");
      k_starve<0><<<c.grid_blocks, 128, 0, s>>>(d, f, k);
      m->launches += 1;
      CKL();
      if (m->split_zbuf) {
        m->pending_gc = true, m->pending_f = f, m->pending_var = c.var;
        return 0;
      }
    }
    return gc_tail_finish(m, f, c.starve, c.var);
  }

  // checkVarSDF + reallocBlocks + flatAndReduceHashTable (voxel_data_structures.cpp:99-103);
  // returns the frame descriptor the rest of the frame must use (the live lists have swapped)
  int variance_pass(FrameCtx& c, int use_frustum, FrameDev& f2) {
    mrh_map* m      = c.m;
    const MapDev& d = m->dev;
    cudaStream_t s  = m->stream;
    const uint32_t cur = c.f.live_cur;
    k_check_var<<<c.grid_blocks, 64, 0, s>>>(d, c.f);
    k_realloc_prepare<<<1, 1, 0, s>>>(d, cur);
    k_realloc<<<c.grid_list, 128, 0, s>>>(d, cur ^ 1u);
    f2          = c.f;
    f2.live_cur = cur ^ 1u;
    k_visible<<<c.grid_list, 256, 0, s>>>(d, f2, m->cam, use_frustum);
    m->launches += 4;
    CKL();
    return 0;
  }

  void end_frame(FrameCtx& c, bool lists_swapped_twice, bool probe_written = false) {
    mrh_map* m = c.m;
    if (!probe_written && m->stream_threshold > 0.f && !m->pending_gc) {
      k_probe<<<1, 1, 0, m->stream>>>(m->dev, m->frame_index);
      m->launches += 1;
    }
    if (!lists_swapped_twice)
      m->live_cur ^= 1u;
    m->frame_index++;
    m->frames_total++;
  }

} // namespace

int finish_gc_tail(mrh_map* m) {
  if (!m->pending_gc)
    return 0;
  m->pending_gc = false;
  return gc_tail_finish(m, m->pending_f, true, m->pending_var);
}

// used by stream-in: make room for n_low resolution-1 blocks
int carve_low_blocks(mrh_map* m, uint32_t n_low) {
  if (n_low == 0)
    return 0;
  k_carve_low<<<m->num_sms * 4, 256, 0, m->stream>>>(m->dev, (n_low + 7u) / 8u);
  m->launches += 1;
  CKL();
  return 0;
}

namespace {
  // the fused frame kernel (mrh_fused.cuh): one persistent launch per frame
  using FrameKernel = void (*)(MapDev, FrameDev, CameraDev, const float*, const uint8_t*, uint32_t, int, int, uint32_t, uint32_t, uint32_t);
  FrameKernel frame_kernel(bool gc, int model, bool fast) {
    if (model == 0 && fast)
      return gc ? k_frame<true, 0, true> : k_frame<false, 0, true>;
    if (model == 0)
      return gc ? k_frame<true, 0, false> : k_frame<false, 0, false>;
    return gc ? k_frame<true, 1, false> : k_frame<false, 1, false>;
  }
} // namespace

// resident CTAs of the fused kernel on the whole device (every instance is compiled for the same bound)
int fused_grid(mrh_map* m) {
  if (m->fused_grid > 0)
    return m->fused_grid;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frame_kernel(true, 0, true), kFuThreads, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  int other = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&other, frame_kernel(true, 1, false), kFuThreads, 0) == cudaSuccess && other >= 1)
    per_sm = std::min(per_sm, other);
  if (const char* e = getenv("MRH_FUSED_CTAS_PER_SM"))
    per_sm = std::max(1, std::min(per_sm, atoi(e)));
  per_sm = std::min(per_sm, 8192 / std::max(1, m->num_sms)); // the fusion queue reserves 8192 terminator entries
  m->fused_ctas_per_sm = per_sm;
  m->fused_grid        = m->num_sms * per_sm;
  return m->fused_grid;
}

int integrate_rgbd(mrh_map* m) {
  FrameCtx c         = begin_frame(m);
  const MapDev& d    = m->dev;
  const CameraDev& k = m->cam;
  const FrameDev& f  = c.f;
  cudaStream_t s     = m->stream;
  const bool prof    = m->profiling;
  auto mark = [&](int i) {
    if (prof)
      cudaEventRecord(m->ev_k[i], s);
  };
  const bool fused_gc = c.gc && !c.starve && !c.var;
  if (!c.var && m->use_fused) {
    // one persistent launch (mrh_fused.cuh)
    if (!m->counters_clean) {
      k_zero_frame_counters<<<1, 1, 0, s>>>(d, f.live_cur ^ 1u);
      m->launches += 1;
    }
    const uint32_t tiles_x = (k.cols + kTileW - 1) / kTileW, tiles_y = (k.rows + kTileH - 1) / kTileH;
    const int rearm        = c.starve ? 0 : 1;
    FrameDev ff            = f;
    ff.tag                 = ++m->fuse_tag;
    ff.band_lo             = 0;
    ff.band_hi             = tiles_x * tiles_y;
    // cp.async.bulk needs 16-byte aligned rows: image width a multiple of 4 pixels, aligned base
    const int bulk_depth = m->use_bulk_depth && (k.cols % 4u) == 0 && ((uintptr_t) m->depth_ptr % 16u) == 0 ? 1 : 0;
    if (m->rgb_ready)
      cudaStreamWaitEvent(s, m->rgb_ready, 0);
    mark(0);
    launch_chained(frame_kernel(fused_gc, k.model, d.fast_div != 0 && k.model == 0), dim3(fused_grid(m)), dim3(kFuThreads), s, m->use_pdl && m->counters_clean, d, ff, k, m->depth_ptr, m->rgb_ptr, tiles_x, rearm, bulk_depth,
                   (uint32_t) m->num_sms, (uint32_t) m->fused_pref_num, (uint32_t) m->fused_pref_den);
    CKL();
    mark(1);
    mark(2);
    mark(3);
    m->launches += 1;
    m->counters_clean = rearm != 0;
  } else if (!c.var) {
    // two-launch frame (mrh_fast.cuh)
    if (!m->counters_clean) {
      k_zero_frame_counters<<<1, 1, 0, s>>>(d, f.live_cur ^ 1u);
      m->launches += 1;
    }
    const uint32_t tiles_x = (k.cols + 31) / 32, tiles_y = (k.rows + kFrontWarps - 1) / kFrontWarps;
    const uint32_t n_vis_ctas = (uint32_t) m->num_sms;
    const int rearm           = c.starve ? 0 : 1;
    mark(0);
    launch_chained(k_front, dim3(tiles_x * tiles_y + n_vis_ctas), dim3(kFrontThreads), s, m->use_pdl && m->counters_clean, d, f, k, m->depth_ptr, tiles_x, n_vis_ctas);
    CKL();
    mark(1);
    if (m->rgb_ready)
      cudaStreamWaitEvent(s, m->rgb_ready, 0);
    mark(2);
    if (fused_gc)
      launch_chained(k_integrate<true>, dim3(m->integrate_grid), dim3(kIntThreads), s, m->use_pdl, d, f, k, m->depth_ptr, m->rgb_ptr, rearm);
    else
      launch_chained(k_integrate<false>, dim3(m->integrate_grid), dim3(kIntThreads), s, m->use_pdl, d, f, k, m->depth_ptr, m->rgb_ptr, rearm);
    CKL();
    mark(3);
    m->launches += 2;
    m->counters_clean = rearm != 0;
  } else {
    if (top_up_low_heap(c))
      return 1;
    const dim3 grid_alloc((k.cols + 31) / 32, (k.rows + 7) / 8);
    mark(0);
    k_alloc_rgbd<<<grid_alloc, 256, 0, s>>>(d, f, k, m->depth_ptr);
    CKL();
    mark(1);
    k_visible<<<c.grid_list, 256, 0, s>>>(d, f, k, 1);
    CKL();
    if (m->rgb_ready)
      cudaStreamWaitEvent(s, m->rgb_ready, 0);
    mark(2);
    k_integrate<false><<<c.grid_blocks, kIntThreads, 0, s>>>(d, f, k, m->depth_ptr, m->rgb_ptr, 0);
    CKL();
    mark(3);
    m->launches += 3;
    m->counters_clean = false;
  }
  bool swapped_twice = false;
  FrameDev f2        = f;
  if (c.var) {
    k_integrate_low<<<c.grid_list, 128, 0, s>>>(d, f, k, m->depth_ptr, m->rgb_ptr);
    m->launches += 1;
    CKL();
    if (c.var_active) {
      if (variance_pass(c, 1, f2))
        return 1;
      k_reintegrate<<<c.grid_list, 128, 0, s>>>(d, f2, k, m->depth_ptr, m->rgb_ptr);
      m->launches += 1;
      CKL();
      swapped_twice = true;
    }
  }
  if (c.gc && !fused_gc && gc_tail(c, f2))
    return 1;
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
  end_frame(c, swapped_twice, !c.var && !c.starve);
  return 0;
}


// integrate3D (:1381-1401): emit -> stable sort by voxel address -> apply. K is the key type.
template <typename K>
static int fuse_points_typed(FrameCtx& c, const FrameDev& f, int key_bits) {
  mrh_map* m      = c.m;
  const MapDev& d = m->dev;
  cudaStream_t s  = m->stream;
  const uint32_t n = (uint32_t) m->n_points, slots = m->upd_slots;
  const size_t items = (size_t) n * slots;
  K* keys[2]   = {(K*) m->d_upd_keys[0], (K*) m->d_upd_keys[1]};
  const K hole = (K) ~(K) 0; // what the memset leaves in an unused slot; its low key_bits sort after every address
  if (cudaMemsetAsync(keys[0], 0xFF, sizeof(K) * items, s) != cudaSuccess)
    return fail("point path: clearing the record slots failed");
  k_points_emit<K><<<(int) ((n + 255) / 256), 256, 0, s>>>(d, f, m->d_points, m->d_normals, n, keys[0], m->d_upd_vals[0], slots);
  CKL();
  size_t tmp = m->sort_tmp_bytes;
  if (cub::DeviceRadixSort::SortPairs(m->d_sort_tmp, tmp, keys[0], keys[1], m->d_upd_vals[0], m->d_upd_vals[1], (int) items, 0, key_bits, s) != cudaSuccess)
    return fail("point path: radix sort failed");
  k_points_apply<K><<<c.grid_list, 256, 0, s>>>(d, keys[1], m->d_upd_vals[1], (uint32_t) items, hole);
  CKL();
  m->launches += 3 + (key_bits + 7) / 8;
  return 0;
}

static int fuse_points(FrameCtx& c, const FrameDev& f) {
  // pool address of a voxel (block * 512 + index) plus one more bit for the hole key
  int key_bits = 1;
  while ((1ull << key_bits) < (unsigned long long) c.m->dev.num_blocks * 512ull)
    ++key_bits;
  ++key_bits;
  return key_bits <= 32 ? fuse_points_typed<uint32_t>(c, f, key_bits) : fuse_points_typed<unsigned long long>(c, f, key_bits);
}

int integrate_points(mrh_map* m) {
  FrameCtx c         = begin_frame(m);
  const MapDev& d    = m->dev;
  const CameraDev& k = m->cam;
  const FrameDev& f  = c.f;
  cudaStream_t s     = m->stream;
  const uint32_t n   = (uint32_t) m->n_points;
  // record slots: a segment of length L = 2t visits at most L (|dx| + |dy| + |dz|) / size + 4 <=
  // sqrt(3) L / size + 4 voxels (one per axis crossing plus the start, plus one per axis of rounding
  // slack); every point owns that many slots, and the sort runs over points x slots items - a count
  // the host knows without asking the device. Anything beyond would be counted in dropped_updates.
  const float t_max    = m->p.sdf_truncation + m->p.sdf_truncation_scale * m->max_integration_distance;
  const size_t per_pt  = (size_t) std::min(std::ceil(1.7320508 * 2.0 * t_max / m->p.virtual_voxel_size) + 4.0, 256.0);
  const size_t max_items = (size_t) 1 << 28;
  const size_t slots   = std::max<size_t>(1, std::min(per_pt, max_items / std::max<size_t>(n, 1)));
  const size_t want    = (size_t) n * slots;
  if (want > max_items)
    return fail("point cloud too large: %zu points", m->n_points);
  m->upd_slots = (uint32_t) slots;
  if (want > m->upd_cap) {
    cudaStreamSynchronize(s);
    for (int i = 0; i < 2; ++i) {
      cudaFree(m->d_upd_keys[i]), cudaFree(m->d_upd_vals[i]);
      m->d_upd_keys[i] = nullptr, m->d_upd_vals[i] = nullptr;
      if (cudaMalloc(&m->d_upd_keys[i], sizeof(unsigned long long) * want) != cudaSuccess || cudaMalloc(&m->d_upd_vals[i], sizeof(float) * want) != cudaSuccess)
        return fail("out of device memory for %zu point-update records", want);
    }
    m->upd_cap = want;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (unsigned long long*) nullptr, (unsigned long long*) nullptr, (float*) nullptr, (float*) nullptr, (int) want, 0, 64, s);
    cudaFree(m->d_sort_tmp);
    m->d_sort_tmp = nullptr;
    if (cudaMalloc(&m->d_sort_tmp, tmp) != cudaSuccess)
      return fail("out of device memory for the sort workspace");
    m->sort_tmp_bytes = tmp;
  }

  if (c.var && top_up_low_heap(c))
    return 1;
  const int grid_pts = (int) ((n + 255) / 256);
  k_alloc_points<<<grid_pts, 256, 0, s>>>(d, f, k, m->d_points, m->d_normals, n);
  CKL();
  k_visible<<<c.grid_list, 256, 0, s>>>(d, f, k, 0);
  CKL();
  m->launches += 2;
  if (fuse_points(c, f))
    return 1;
  bool swapped_twice = false;
  FrameDev f2        = f;
  if (c.var_active) {
    if (variance_pass(c, 0, f2))
      return 1;
    // reintegrate3D re-launches the full integrate3DKernel (:1568): the frame is fused a second time (Q7)
    if (fuse_points(c, f2))
      return 1;
    swapped_twice = true;
  }
  if (c.gc && gc_tail(c, f2))
    return 1;
  m->counters_clean = false;
  end_frame(c, swapped_twice);
  return 0;
}

} // namespace mrh

// ---------------------------------------------------------------------------------------------
// self-test of mrh_div.cuh (declared in include/mrhash_b200.h)
// ---------------------------------------------------------------------------------------------
namespace {
  __device__ __forceinline__ uint32_t mix32(uint64_t x) {
    x ^= x >> 33, x *= 0xff51afd7ed558ccdull, x ^= x >> 33, x *= 0xc4ceb9fe1a85ec53ull, x ^= x >> 33;
    return (uint32_t) x;
  }
  __global__ void k_selftest_div_all(float b, unsigned long long* bad) {
    const float y1 = mrh::div_recip(b);
    unsigned long long n = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; i < (1ull << 32); i += (unsigned long long) gridDim.x * blockDim.x) {
      const float a = __uint_as_float((uint32_t) i);
      const float q = mrh::div_fast(a, b, y1), r = __fdiv_rn(a, b);
      if (__float_as_uint(q) != __float_as_uint(r) && !(q != q && r != r))
        ++n;
      const float qs = mrh::div_fast_signed(a, -b, y1), rs = __fdiv_rn(a, -b);
      if (__float_as_uint(qs) != __float_as_uint(rs) && !(qs != qs && rs != rs))
        ++n;
    }
    if (n)
      atomicAdd(bad, n);
  }
  __global__ void k_selftest_div_random(unsigned long long n_pairs, unsigned long long seed, unsigned long long* bad) {
    unsigned long long n = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; i < n_pairs; i += (unsigned long long) gridDim.x * blockDim.x) {
      const float a = __uint_as_float(mix32(seed + 2 * i));
      // divisor: any sign-less float inside the window the kernels guarantee (2^-40 .. 2^40)
      const uint32_t e = 87u + mix32(seed + 2 * i + 1) % 81u;
      const float b    = __uint_as_float((e << 23) | (mix32(seed ^ (i * 0x9E3779B97F4A7C15ull)) & 0x7FFFFFu));
      if (!mrh::div_range_ok(b))
        continue;
      const float q = mrh::div_fast(a, b, mrh::div_recip(b)), r = __fdiv_rn(a, b);
      if (__float_as_uint(q) != __float_as_uint(r) && !(q != q && r != r))
        ++n;
    }
    if (n)
      atomicAdd(bad, n);
  }
} // namespace

extern "C" int mrh_selftest_div(int device, float divisor, uint64_t n_random, uint64_t seed, uint64_t* mismatches) {
  using mrh::fail;
  if (!mismatches)
    return fail("null argument");
  if (cudaSetDevice(device < 0 ? 0 : device) != cudaSuccess)
    return fail("no CUDA device: libmrhash_b200 has no CPU path");
  unsigned long long* d_bad = nullptr;
  if (cudaMalloc(&d_bad, sizeof(unsigned long long)) != cudaSuccess || cudaMemset(d_bad, 0, sizeof(unsigned long long)) != cudaSuccess)
    return fail("selftest: allocation failed");
  if (divisor > 0.f) {
    const float ad = divisor;
    if (!(ad >= 9.094947017729282e-13f && ad <= 1.099511627776e12f)) {
      cudaFree(d_bad);
      return fail("selftest: divisor outside the window of the shared-reciprocal division");
    }
    k_selftest_div_all<<<148 * 16, 256>>>(divisor, d_bad);
  }
  if (n_random)
    k_selftest_div_random<<<148 * 16, 256>>>(n_random, seed, d_bad);
  unsigned long long h = 0;
  const cudaError_t e = cudaMemcpy(&h, d_bad, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d_bad);
  if (e != cudaSuccess)
    return fail("selftest: %s", cudaGetErrorString(e));
  *mismatches = h;
  return 0;
}


// mrh_var.cuh — variance-adaptive resolution path (sdf_var_threshold > 0).
//
// Replaces allocateMemoryLow (voxel_data_structures.cu:860-871), checkVarSDFKernel (:1857-1939),
// reallocBlocksKernel / reallocBlock (:626-755, 2021-2069), reintegrateDepthMapKernel (:1942-2018,
// launched as at :2097) and the resolution-1 branches of integrateDepthMapKernel /
// garbageCollectIdentifyKernel / garbageCollectFreeKernel. Two levels exist: resolution 0 (8^3
// voxels of the base size) and resolution 1 (4^3 voxels of twice the size, eight 768-byte sub-slots
// carved out of one pool block).
#pragma once
#include "mrh_table.cuh"

namespace mrh {

constexpr uint32_t kLowSlotBytes = 768; // sdf[64] | sum_sq[64] | rgbw[64]

// allocBlocks :885-891: top up the low heap when fewer than low_blocks_to_allocate_ sub-slots are free
__global__ void k_carve_decide(MapDev m, uint32_t low_blocks_to_allocate) {
  const int free_low   = m.ctr->heap_low_counter + 1;
  m.ctr->carve_request = ((uint32_t) free_low < low_blocks_to_allocate) ? low_blocks_to_allocate : 0u;
}

// carve `n_parents` (or ctr->carve_request when n_parents == 0) pool blocks into 8 sub-slots each
__global__ void __launch_bounds__(256) k_carve_low(MapDev m, uint32_t n_parents) {
  const uint32_t n = n_parents ? n_parents : m.ctr->carve_request;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int addr_high = atomicSub(&m.ctr->heap_counter, 1);
    if (addr_high < 0) {
      atomicAdd(&m.ctr->heap_counter, 1);
      atomicAdd(&m.ctr->dropped_heap, 1ull);
      continue;
    }
    const uint32_t parent = m.heap[addr_high];
    const int addr_low    = atomicAdd(&m.ctr->heap_low_counter, 8);
#pragma unroll
    for (int idx = 1; idx <= 8; ++idx)
      m.heap_low[addr_low + idx] = parent * 8u + 8u - (uint32_t) idx;
    m.carved[parent] = 1;
    atomicAdd(&m.ctr->low_parents, 1ull);
  }
}

// the fusion rule shared by every depth-map kernel (integrateDepthMapKernel :1155-1180 +
// combineVoxel, voxel_hash_utils.cuh:169-181). track_variance = false restates
// reintegrateDepthMapKernel, which leaves sum_squared at the default 0.
__device__ __forceinline__ void fuse_rgbd(float& sdf0, float& ss0, uint32_t& cw, float sdf, const uint8_t* px, int weight_sample, float half_size, bool track_variance) {
  const uint32_t w0 = cw >> 24;
  const uint32_t r1 = px[0], g1 = px[1], b1 = px[2];
  uint32_t r0 = cw & 0xFF, g0 = (cw >> 8) & 0xFF, b0 = (cw >> 16) & 0xFF;
  if (w0 == 0)
    r0 = r1, g0 = g1, b0 = b1;
  const float curr_mean = w0 > 0 ? sdf0 : sdf;
  const float delta     = fdiv(fsub(sdf, curr_mean), half_size);
  const uint32_t wsum   = w0 + (uint32_t) weight_sample;
  const float merged    = fdiv(ffma(sdf, __uint2float_rn((uint32_t) weight_sample), fmul(sdf0, __uint2float_rn(w0))), __uint2float_rn(wsum));
  const uint32_t rr     = (uint32_t) f2i(fadd(ffma(__uint2float_rn(r1), 0.5f, fmul(__uint2float_rn(r0), 0.5f)), 0.5f)) & 0xFF;
  const uint32_t gg     = (uint32_t) f2i(fadd(ffma(__uint2float_rn(g1), 0.5f, fmul(__uint2float_rn(g0), 0.5f)), 0.5f)) & 0xFF;
  const uint32_t bb     = (uint32_t) f2i(fadd(ffma(__uint2float_rn(b1), 0.5f, fmul(__uint2float_rn(b0), 0.5f)), 0.5f)) & 0xFF;
  const uint32_t wn     = min(wsum, (uint32_t) kWeightMax);
  float ss              = 0.f;
  if (track_variance) {
    const float delta2 = fdiv(fsub(sdf, merged), half_size);
    ss                 = fmul(delta, delta2);
    if (fabsf(ss) < 1.175494350822287508e-38f)
      ss = 0.f;
    ss = fadd(0.f, ss);
  }
  sdf0 = merged, ss0 = ss, cw = rr | (gg << 8) | (bb << 16) | (wn << 24);
}

// depth test of one voxel centre; returns true and (sdf, pixel index) when the voxel is updated
__device__ __forceinline__ bool voxel_sample(const MapDev& m, const CameraDev& cam, const PoseDev& pose, const float* __restrict__ depth, i3 pi, float& sdf_out, uint32_t& pix) {
  const f3 pf = {fmul(i2f(pi.x), m.voxel_size), fmul(i2f(pi.y), m.voxel_size), fmul(i2f(pi.z), m.voxel_size)};
  const f3 pc = se3_mul(pose.Ri, pose.ti, pf);
  int row, col;
  if (!project_point(cam, pc, row, col))
    return false;
  const uint32_t p = (uint32_t) row * cam.cols + (uint32_t) col;
  const float d    = cloud_depth(cam, (uint32_t) row, (uint32_t) col, __ldg(depth + p));
  if (d == 0.f || d > m.max_integration_distance)
    return false;
  float sdf     = fsub(d, get_depth(cam, pc));
  const float t = truncation(m.trunc, m.trunc_scale, d);
  if (sdf <= -t)
    return false;
  sdf_out = (sdf >= 0.f) ? fminf(t, sdf) : fmaxf(-t, sdf);
  pix     = p;
  return true;
}

// integrateDepthMapKernel for resolution-1 entries of the visible list: one warp per block,
// two voxels per lane. voxel_limit = 64 (full) .
__global__ void __launch_bounds__(128) k_integrate_low(MapDev m, FrameDev f, CameraDev cam, const float* __restrict__ depth, const uint8_t* __restrict__ rgb) {
  __shared__ PoseDev pose;
  if (threadIdx.x == 0)
    load_pose(f, pose);
  __syncthreads();
  const int lane       = threadIdx.x & 31;
  const uint32_t warp  = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarp = (gridDim.x * blockDim.x) >> 5;
  const uint32_t n_vis = m.ctr->vis_count;
  const float half     = fmul(m.voxel_size, 0.5f);
  unsigned long long updated = 0;
  for (uint32_t bi = warp; bi < n_vis; bi += nwarp) {
    const VisEntry e = m.vis[bi];
    if (!(e.val & 0x80000000u))
      continue;
    uint8_t* base = m.pool + (size_t) (e.val & 0x7FFFFFFFu) * kLowSlotBytes;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int vi = lane + 32 * h;
      const i3 pi  = {e.x * 8 + 2 * (vi % 4), e.y * 8 + 2 * ((vi % 16) / 4), e.z * 8 + 2 * (vi / 16)};
      float sdf;
      uint32_t pix;
      if (!voxel_sample(m, cam, pose, depth, pi, sdf, pix))
        continue;
      float* psdf   = reinterpret_cast<float*>(base) + vi;
      float* pss    = reinterpret_cast<float*>(base + 256) + vi;
      uint32_t* pcw = reinterpret_cast<uint32_t*>(base + 512) + vi;
      float s0 = *psdf, q0 = *pss;
      uint32_t cw = *pcw;
      fuse_rgbd(s0, q0, cw, sdf, rgb + (size_t) pix * 3, m.weight_sample, half, true);
      *psdf = s0, *pss = q0, *pcw = cw;
      ++updated;
    }
  }
  for (int o = 16; o > 0; o >>= 1)
    updated += __shfl_xor_sync(0xFFFFFFFFu, updated, o);
  if (lane == 0 && updated)
    atomicAdd(&m.ctr->voxels_updated, updated);
}

// checkVarSDFKernel: one 64-thread CTA per visible resolution-0 block; the summation order
// (2x2x2 cell per thread, then a stride-halving tree) is the reference's.
__global__ void __launch_bounds__(64) k_check_var(MapDev m, FrameDev f) {
  __shared__ float s_ss[64], s_w[64];
  __shared__ int s_free;
  const int tid        = threadIdx.x;
  const uint32_t n_vis = m.ctr->vis_count;
  for (uint32_t bi = blockIdx.x; bi < n_vis; bi += gridDim.x) {
    const VisEntry e = m.vis[bi];
    if (e.val & 0x80000000u)
      continue;
    uint8_t* base       = m.pool + (size_t) e.val * kBlockBytes;
    const float* ssp    = reinterpret_cast<const float*>(base + kPlaneBytes);
    const uint32_t* cwp = reinterpret_cast<const uint32_t*>(base + 2 * kPlaneBytes);
    const int gx = (tid % 4) * 2, gy = ((tid / 4) % 4) * 2, gz = (tid / 16) * 2;
    float ls = 0.f, lw = 0.f;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int li     = (gz + dz) * 64 + (gy + dy) * 8 + (gx + dx);
          const uint32_t w = cwp[li] >> 24;
          if (w > 0) {
            ls = fadd(ls, ssp[li]);
            lw = fadd(lw, __uint2float_rn(w));
          }
        }
    s_ss[tid] = ls, s_w[tid] = lw;
    __syncthreads();
    for (int stride = 32; stride > 0; stride >>= 1) {
      if (tid < stride) {
        s_ss[tid] = fadd(s_ss[tid], s_ss[tid + stride]);
        s_w[tid]  = fadd(s_w[tid], s_w[tid + stride]);
      }
      __syncthreads();
    }
    if (tid == 0) {
      int do_free = 0;
      const float w = s_w[0];
      if (!(w < 2.f)) {
        const float wm1      = fsub(w, 1.f);
        const double avg_var = (double) fdiv(s_ss[0], wm1);
        if (wm1 > 1e-6f && avg_var > 0.0 && avg_var < (double) m.var_threshold && m.keys[e.slot] < kNoKey) {
          // deleteHashEntryElement + queue for re-allocation one level coarser (:1925-1936)
          atomicExch(m.keys + e.slot, kTomb);
          const int addr   = atomicAdd(&m.ctr->heap_counter, 1);
          m.heap[addr + 1] = e.val;
          m.live[f.live_cur ^ 1u][e.live_idx].slot = kInvalid;
          m.stats[e.val]   = {3.40282346638528859812e+38f, 0u};
          const uint32_t q = atomicAdd(&m.ctr->n_realloc, 1u);
          VisEntry r       = e;
          r.val            = 1u; // target resolution
          m.realloc_list[q] = r;
          do_free           = 1;
        }
      }
      s_free = do_free;
    }
    __syncthreads();
    if (s_free) {
      const float4 z = {0.f, 0.f, 0.f, 0.f};
      float4* p4     = reinterpret_cast<float4*>(base);
      for (int i = tid; i < (int) (kBlockBytes / 16); i += 64)
        p4[i] = z;
    }
    __syncthreads();
  }
}

// reallocBlocks :2036-2042: the re-integration list is only cleared when something was queued;
// also re-arms the lists that the second visibility pass of the frame fills.
__global__ void k_realloc_prepare(MapDev m, uint32_t live_out) {
  if (m.ctr->n_realloc > 0)
    m.ctr->n_reintegrate = 0;
  m.ctr->live_count[live_out] = 0;
  m.ctr->vis_count            = 0;
}

// reallocBlocksKernel: one warp per queued block; live_in is the list new entries are appended to
__global__ void __launch_bounds__(128) k_realloc(MapDev m, uint32_t live_in) {
  const int lane       = threadIdx.x & 31;
  const uint32_t warp  = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarp = (gridDim.x * blockDim.x) >> 5;
  const uint32_t n     = m.ctr->n_realloc;
  CameraDev cam_unused{};
  PoseDev pose_unused{};
  for (uint32_t i = warp; i < n; i += nwarp) {
    const VisEntry r   = m.realloc_list[i];
    const uint32_t val = warp_insert<false>(m, cam_unused, pose_unused, live_in, {r.x, r.y, r.z}, lane, (int) r.val);
    if (lane == 0 && val != kInvalid) {
      const uint32_t q = atomicAdd(&m.ctr->n_reintegrate, 1u);
      m.reint_keys[q]  = pack_key({r.x, r.y, r.z});
      atomicAdd(&m.ctr->blocks_realloc, 1ull);
    }
  }
}

// reintegrateDepthMapKernel as launched at :2097 (<<<n_blocks, n_threads>>> with a 2-D grid):
// blockDim.y == 1, so voxel_idx = blockIdx.y in [0, 32) - only voxels 0..31 of each listed block are
// re-fused (Q6), and sum_squared is left at 0. The reference lists table slots; this lists keys and
// looks them up, so a key that has left the map in the meantime is skipped.
__global__ void __launch_bounds__(128) k_reintegrate(MapDev m, FrameDev f, CameraDev cam, const float* __restrict__ depth, const uint8_t* __restrict__ rgb) {
  __shared__ PoseDev pose;
  if (threadIdx.x == 0)
    load_pose(f, pose);
  __syncthreads();
  const int lane       = threadIdx.x & 31;
  const uint32_t warp  = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarp = (gridDim.x * blockDim.x) >> 5;
  const uint32_t n     = m.ctr->n_reintegrate;
  const float half     = fmul(m.voxel_size, 0.5f);
  unsigned long long updated = 0;
  for (uint32_t i = warp; i < n; i += nwarp) {
    const i3 b     = unpack_key(m.reint_keys[i]);
    const int slot = table_find(m, b);
    if (slot < 0)
      continue;
    const uint32_t val = m.vals[slot];
    const int r        = (int) (val >> 31);
    const int vi       = lane; // voxel indices 0..31
    const int bs = 8 >> r, sf = 1 << r;
    const i3 pi = {b.x * 8 + sf * (vi % bs), b.y * 8 + sf * ((vi % (bs * bs)) / bs), b.z * 8 + sf * (vi / (bs * bs))};
    float sdf;
    uint32_t pix;
    if (!voxel_sample(m, cam, pose, depth, pi, sdf, pix))
      continue;
    float *psdf, *pss;
    uint32_t* pcw;
    if (r) {
      uint8_t* base = m.pool + (size_t) (val & 0x7FFFFFFFu) * kLowSlotBytes;
      psdf = reinterpret_cast<float*>(base) + vi, pss = reinterpret_cast<float*>(base + 256) + vi, pcw = reinterpret_cast<uint32_t*>(base + 512) + vi;
    } else {
      uint8_t* base = m.pool + (size_t) val * kBlockBytes;
      psdf = reinterpret_cast<float*>(base) + vi, pss = reinterpret_cast<float*>(base + kPlaneBytes) + vi, pcw = reinterpret_cast<uint32_t*>(base + 2 * kPlaneBytes) + vi;
    }
    float s0 = *psdf, q0 = *pss;
    uint32_t cw = *pcw;
    fuse_rgbd(s0, q0, cw, sdf, rgb + (size_t) pix * 3, m.weight_sample, half, false);
    *psdf = s0, *pss = q0, *pcw = cw;
    ++updated;
  }
  for (int o = 16; o > 0; o >>= 1)
    updated += __shfl_xor_sync(0xFFFFFFFFu, updated, o);
  if (lane == 0 && updated)
    atomicAdd(&m.ctr->voxels_updated, updated);
}

// garbage collection of resolution-1 entries of the visible list (identify + free in one pass:
// garbageCollectIdentifyKernel :1674-1713 with 32 pairs, garbageCollectFreeKernel :1827-1844)
__global__ void __launch_bounds__(128) k_gc_low(MapDev m, FrameDev f) {
  const int lane       = threadIdx.x & 31;
  const uint32_t warp  = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarp = (gridDim.x * blockDim.x) >> 5;
  const uint32_t n_vis = m.ctr->vis_count;
  for (uint32_t bi = warp; bi < n_vis; bi += nwarp) {
    const VisEntry e = m.vis[bi];
    if (!(e.val & 0x80000000u))
      continue;
    if (m.keys[e.slot] >= kNoKey)
      continue;
    const uint32_t low = e.val & 0x7FFFFFFFu;
    uint8_t* base      = m.pool + (size_t) low * kLowSlotBytes;
    float min_abs      = 3.40282346638528859812e+38f;
    uint32_t max_w     = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int vi      = lane + 32 * h;
      const uint32_t w  = reinterpret_cast<const uint32_t*>(base + 512)[vi] >> 24;
      if (w)
        min_abs = fminf(min_abs, fabsf(reinterpret_cast<const float*>(base)[vi]));
      max_w = max(max_w, w);
    }
    for (int o = 16; o > 0; o >>= 1) {
      min_abs = fminf(min_abs, __shfl_xor_sync(0xFFFFFFFFu, min_abs, o));
      max_w   = max(max_w, __shfl_xor_sync(0xFFFFFFFFu, max_w, o));
    }
    if (!(min_abs >= m.gc_threshold || max_w == 0u))
      continue;
    for (int i = lane; i < (int) (kLowSlotBytes / 4); i += 32)
      reinterpret_cast<uint32_t*>(base)[i] = 0u;
    if (lane == 0) {
      atomicExch(m.keys + e.slot, kTomb);
      const int addr       = atomicAdd(&m.ctr->heap_low_counter, 1); // appendHeapLow (:58-62)
      m.heap_low[addr + 1] = low;
      m.live[f.live_cur ^ 1u][e.live_idx].slot = kInvalid;
      atomicAdd(&m.ctr->blocks_freed, 1ull);
      atomicAdd(&m.ctr->low_live, (unsigned long long) -1ll);
    }
  }
}

} // namespace mrh

// mrh_fast.cuh — the two-launch RGB-D frame (single resolution, the configuration every shipped
// RGB-D config uses: sdf_var_threshold = 0).
//
//   k_front      : the first n_vis_ctas CTAs run the visibility pass over the blocks that were already
//                  live (flatAndReduceHashTable, voxel_data_structures.cu:406-449); every other CTA
//                  walks the depth rays of one 32x8 pixel tile and inserts missing blocks
//                  (allocBlocksKernel + retry loop, :758-922). The two roles are independent: a block
//                  inserted this frame has just passed the frustum test, so the inserting warp
//                  appends it to the output live list and the visible list itself.
//   k_integrate  : (mrh_kernels.cuh) one CTA per visible block; with rearm = 1 the last CTA to finish
//                  re-arms the list counters for the next frame, so a frame needs no reset kernel
//                  and no host round trip.
#pragma once
#include "mrh_kernels.cuh"

namespace mrh {

constexpr int kTileSet = 256; // per-tile set of block keys (one tile = 32 x 8 rays)

__device__ __forceinline__ uint32_t key_hash(unsigned long long k) {
  k ^= k >> 29;
  k *= 0x9E3779B97F4A7C15ull;
  return (uint32_t) (k >> 40);
}

// insert-if-absent into the tile's shared-memory key set; false = no room within the probe limit
__device__ __forceinline__ bool tile_set_insert(unsigned long long* s_set, unsigned long long key) {
  const uint32_t h = key_hash(key);
#pragma unroll 1
  for (int j = 0; j < 16; ++j) {
    unsigned long long* cell   = s_set + ((h + j) & (kTileSet - 1));
    const unsigned long long c = *reinterpret_cast<volatile unsigned long long*>(cell);
    if (c == key)
      return true;
    if (c == kNoKey) {
      const unsigned long long old = atomicCAS(cell, kNoKey, key);
      if (old == kNoKey || old == key)
        return true;
    }
  }
  return false;
}

// Allocation role, two phases per tile:
//   1. every ray walks its block DDA and drops the visited keys into a shared-memory set (the 256
//      rays of a tile visit only ~10-20 distinct blocks);
//   2. the distinct keys are handed out to the 8 warps, ONE warp-cooperative table insert each.
// If the set fills up (long grazing rays), the walk pauses, phase 2 drains the set, and the walk
// resumes: any ray length is handled.
#ifndef MRH_FRONT_MIN_CTAS
#define MRH_FRONT_MIN_CTAS 5
#endif
__global__ void __launch_bounds__(256, MRH_FRONT_MIN_CTAS) k_front(MapDev m, FrameDev f, CameraDev cam, const float* __restrict__ depth, uint32_t tiles_x, uint32_t n_vis_ctas) {
  __shared__ PoseDev pose;
  __shared__ unsigned long long s_set[kTileSet];
  __shared__ unsigned long long s_list[kTileSet];
  __shared__ uint32_t s_n;
  __shared__ int s_full;
  if (threadIdx.x == 0) {
    load_pose(f, pose);
    s_n = 0, s_full = 0;
  }
  s_set[threadIdx.x] = kNoKey;
  __syncthreads();
  if (blockIdx.x < n_vis_ctas) { // scheduled first: the visibility role is a pure latency chain
    visible_pass(m, f.live_cur, cam, pose, 1, blockIdx.x, n_vis_ctas);
    return;
  }
  const uint32_t tile = blockIdx.x - n_vis_ctas;
  const unsigned full = 0xFFFFFFFFu;
  const int lane      = threadIdx.x & 31;
  const int warp      = threadIdx.x >> 5;
  const uint32_t col  = (tile % tiles_x) * 32 + lane;
  const uint32_t row  = (tile / tiles_x) * 8 + warp;
  bool active         = false;
  DDA dda;
  if (row < cam.rows && col < cam.cols) {
    const float raw = __ldg(depth + (size_t) row * cam.cols + col);
    const float d   = cloud_depth(cam, row, col, raw);
    if (d != 0.f) {
      const float t    = truncation(m.trunc, m.trunc_scale, d);
      const float dmin = fminf(m.max_integration_distance, fsub(d, t));
      const float dmax = fminf(m.max_integration_distance, fadd(d, t));
      if (!(dmin >= dmax)) {
        const f3 p0 = se3_mul(pose.R, pose.t, inverse_projection(cam, row, col, dmin));
        const f3 p1 = se3_mul(pose.R, pose.t, inverse_projection(cam, row, col, dmax));
        dda.init(p0, p1, m.voxel_size, m.ext, true);
        active = true;
      }
    }
  }
  const unsigned n_rays = __popc(__ballot_sync(full, active));
  if (lane == 0 && n_rays)
    atomicAdd(&m.ctr->rays_valid, (unsigned long long) n_rays);
  int iter = 0;
  bool more;
  do {
    // ---- phase 1: walk, collect keys ----
    while (__any_sync(full, active)) {
      if (__shfl_sync(full, *reinterpret_cast<volatile int*>(&s_full), 0))
        break;
      unsigned long long key = kNoKey;
      if (active) {
        if (key_in_range(dda.cur))
          key = pack_key(dda.cur);
        else
          atomicAdd(&m.ctr->dropped_table, 1ull);
      }
      // one lane per distinct key of the warp talks to the set
      const unsigned peers = __match_any_sync(full, key);
      const int leader     = __ffs(peers) - 1;
      int placed           = 1;
      if (key != kNoKey && lane == leader)
        placed = tile_set_insert(s_set, key) ? 1 : 0;
      placed = __shfl_sync(full, placed, leader);
      if (active) {
        if (placed) {
          active = dda.advance();
          if (++iter >= kMaxDDA)
            active = false;
        } else {
          s_full = 1; // keep the DDA where it is: this key is retried after the set has been drained
        }
      }
    }
    __syncthreads();
    // ---- phase 2: one table insert per distinct key ----
    {
      const unsigned long long k = s_set[threadIdx.x];
      const unsigned has         = __ballot_sync(full, k != kNoKey);
      uint32_t base              = 0;
      if (lane == 0 && has)
        base = atomicAdd(&s_n, (uint32_t) __popc(has));
      base = __shfl_sync(full, base, 0);
      if (k != kNoKey)
        s_list[base + __popc(has & ((1u << lane) - 1u))] = k;
    }
    __syncthreads();
    const uint32_t n_keys = s_n;
    for (uint32_t i = warp; i < n_keys; i += 8)
      warp_insert<true, true>(m, cam, pose, f.live_cur, unpack_key(s_list[i]), lane);
    more = s_full != 0;
    __syncthreads();
    s_set[threadIdx.x] = kNoKey;
    if (threadIdx.x == 0)
      s_n = 0, s_full = 0;
    __syncthreads();
  } while (more);
}

__global__ void k_zero_frame_counters(MapDev m, uint32_t live_out) {
  m.ctr->live_count[live_out] = 0;
  m.ctr->vis_count            = 0;
  m.ctr->done_ctas            = 0;
}

} // namespace mrh

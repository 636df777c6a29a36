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

// insert-if-absent into the tile's shared-memory key set; 0 = no room within the probe limit,
// 1 = the key was there already, 2 = this call placed it
__device__ __forceinline__ int tile_set_insert(unsigned long long* s_set, unsigned long long key) {
  const uint32_t h = key_hash(key);
#pragma unroll 1
  for (int j = 0; j < 16; ++j) {
    unsigned long long* cell   = s_set + ((h + j) & (kTileSet - 1));
    const unsigned long long c = *reinterpret_cast<volatile unsigned long long*>(cell);
    if (c == key)
      return 1;
    if (c == kNoKey) {
      const unsigned long long old = atomicCAS(cell, kNoKey, key);
      if (old == kNoKey)
        return 2;
      if (old == key)
        return 1;
    }
  }
  return 0;
}

// Allocation role, two phases per tile:
//   1. every ray walks its block DDA and drops the visited keys into a shared-memory set (the 256
//      rays of a tile visit only ~10-20 distinct blocks);
//   2. one thread per distinct key checks whether the block exists; the few that do not are handed
//      out to the 8 warps, ONE warp-cooperative table insert each.
// If the set fills up (long grazing rays), the walk pauses, phase 2 drains the set, and the walk
// resumes: any ray length is handled.
#ifndef MRH_FRONT_MIN_CTAS
#define MRH_FRONT_MIN_CTAS 5
#endif
#ifndef MRH_FRONT_WARPS
#define MRH_FRONT_WARPS 8 // rows of 32 rays per tile
#endif
constexpr int kFrontWarps   = MRH_FRONT_WARPS;
constexpr int kFrontThreads = 32 * kFrontWarps;
__global__ void __launch_bounds__(kFrontThreads, MRH_FRONT_MIN_CTAS * 8 / MRH_FRONT_WARPS) k_front(MapDev m, FrameDev f, CameraDev cam, const float* __restrict__ depth, uint32_t tiles_x, uint32_t n_vis_ctas) {
  __shared__ unsigned long long s_set[kTileSet];
  __shared__ unsigned long long s_list[kTileSet];
  __shared__ uint32_t s_n;
  __shared__ int s_full;
  cudaGridDependencySynchronize(); // chained launch: the previous frame's k_integrate has completed (no-op otherwise)
  const PoseDev& pose = frame_pose(f);
  if (blockIdx.x < n_vis_ctas) { // scheduled first: the visibility role is a pure latency chain
    visible_pass(m, f.live_cur, cam, pose, 1, blockIdx.x, n_vis_ctas);
    return;
  }
  const uint32_t tile = blockIdx.x - n_vis_ctas;
  const unsigned full = 0xFFFFFFFFu;
  const int lane      = threadIdx.x & 31;
  const int warp      = threadIdx.x >> 5;
  const uint32_t col  = (tile % tiles_x) * 32 + lane;
  const uint32_t row  = (tile / tiles_x) * kFrontWarps + warp;
  const bool inside   = row < cam.rows && col < cam.cols;
  // the depth read is the head of every ray's dependency chain: issue it before the set is cleared
  const float raw = inside ? __ldg(depth + (size_t) row * cam.cols + col) : 0.f;
  if (threadIdx.x == 0)
    s_n = 0, s_full = 0;
  for (int i = threadIdx.x; i < kTileSet; i += kFrontThreads)
    s_set[i] = kNoKey;
  __syncthreads();
  bool active = false;
  DDA dda;
  if (inside) {
    const float d = cloud_depth(cam, row, col, raw);
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
    // ---- phase 1: every lane walks its own ray and drops the visited keys into the tile set. No
    // warp collectives in here: ~90 % of the probes hit a key a neighbouring ray has already placed,
    // which is one shared-memory load (a broadcast when several lanes probe the same key).
    while (active) {
      if (*reinterpret_cast<volatile int*>(&s_full))
        break; // the set is being drained: keep the DDA where it is and resume afterwards
      int placed = 1;
      if (key_in_range(dda.cur)) {
        placed = tile_set_insert(s_set, pack_key(dda.cur));
        if (placed == 2) {
          // first sighting of this block in the tile: start pulling its two bucket lines towards the
          // SM now, phase 2 looks the key up once the walks are done
          const uint32_t h = block_hash_fast(m, dda.cur);
          if (h >= m.shard_lo && h < m.shard_hi) {
            prefetch_l2(m.keys + (size_t) h * kBucketSlots);
            prefetch_l2(m.keys + (size_t) (h + 1 < m.num_buckets ? h + 1 : 0) * kBucketSlots);
          }
        }
      } else {
        atomicAdd(&m.ctr->dropped_table, 1ull);
      }
      if (!placed) {
        s_full = 1; // this key is retried after the set has been drained
        break;
      }
      active = dda.advance();
      if (++iter >= kMaxDDA)
        active = false;
    }
    __syncthreads();
    // ---- phase 2: the distinct keys of the tile. Nearly all of them are blocks that exist already:
    // one thread per key looks its key up first (two 128-byte bucket lines), and only the absent
    // ones (a few dozen per FRAME) take the warp-cooperative insert.
    for (int cell = threadIdx.x; cell < kTileSet; cell += kFrontThreads) {
      unsigned long long k = s_set[cell];
      if (k != kNoKey) {
        const i3 kb      = unpack_key(k);
        const uint32_t h = block_hash_fast(m, kb);
        if (h < m.shard_lo || h >= m.shard_hi || table_find(m, kb) >= 0)
          k = kNoKey; // another GPU's block, or present
      }
      const unsigned has = __ballot_sync(full, k != kNoKey);
      uint32_t base      = 0;
      if (lane == 0 && has)
        base = atomicAdd(&s_n, (uint32_t) __popc(has));
      base = __shfl_sync(full, base, 0);
      if (k != kNoKey)
        s_list[base + __popc(has & ((1u << lane) - 1u))] = k;
    }
    __syncthreads();
    const uint32_t n_keys = s_n;
    for (uint32_t i = warp; i < n_keys; i += kFrontWarps)
      warp_insert<true, true>(m, cam, pose, f.live_cur, unpack_key(s_list[i]), lane);
    more = s_full != 0;
    __syncthreads();
    for (int i = threadIdx.x; i < kTileSet; i += kFrontThreads)
      s_set[i] = kNoKey;
    if (threadIdx.x == 0)
      s_n = 0, s_full = 0;
    __syncthreads();
  } while (more);
}

__global__ void k_zero_frame_counters(MapDev m, uint32_t live_out) {
  m.ctr->live_count[live_out] = 0;
  m.ctr->vis_count            = 0;
  m.ctr->done_ctas            = 0;
  for (int i = 0; i < kQueueShards; ++i)
    m.fqs->q_chunk[i].v = 0, m.fqs->q_tile[i].v = 0, m.fqs->q_fuse[i].v = 0;
  m.fqs->fq_count.v = 0, m.fqs->items_done.v = 0, m.fqs->gc_count.v = 0, m.fqs->done_ctas.v = 0;
}

} // namespace mrh

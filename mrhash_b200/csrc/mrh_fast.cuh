// mrh_fast.cuh — the two-launch RGB-D frame (single resolution, the configuration every shipped
// RGB-D config uses: sdf_var_threshold = 0).
//
//   k_front      : the first n_vis_ctas CTAs run the visibility pass over the blocks that were already
//                  live (flatAndReduceHashTable, voxel_data_structures.cu:406-449); every other CTA
//                  walks the depth rays of one 32x8 pixel tile and inserts missing blocks
//                  (allocBlocksKernel + retry loop, :758-922). The two roles are independent: a block
//                  inserted this frame has just passed the frustum test, so the inserting warp
//                  appends it to the output live list and the visible list itself.
//   k_integrate8 : one 64-thread CTA per visible block, 8 consecutive-x voxels per thread
//                  (integrateDepthMapKernel :1095-1181 + garbageCollectIdentify/Free :1674-1854);
//                  the last CTA to finish re-arms the list counters for the next frame, so a frame
//                  needs no reset kernel and no host round trip.
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
__global__ void __launch_bounds__(256) k_front(MapDev m, FrameDev f, CameraDev cam, const float* __restrict__ depth, uint32_t tiles_x, uint32_t n_vis_ctas) {
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

// One 64-thread CTA per visible block. Thread t owns the x-row (y = t & 7, z = t >> 3): 8 voxels =
// 32 contiguous bytes in each plane, so a warp reads 1 KB contiguous per plane with 128-bit loads.
template <bool FUSE_GC>
__global__ void __launch_bounds__(64) k_integrate8(MapDev m, FrameDev f, CameraDev cam, const float* __restrict__ depth, const uint8_t* __restrict__ rgb, int rearm) {
  __shared__ PoseDev pose;
  __shared__ float s_min[2];
  __shared__ uint32_t s_max[2];
  __shared__ uint32_t s_upd[2];
  __shared__ int s_delete;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0)
    load_pose(f, pose);
  __syncthreads();
  const uint32_t n_vis  = m.ctr->vis_count;
  const int ly = tid & 7, lz = tid >> 3;
  const float half_size = fmul(m.voxel_size, 0.5f);
  const float ws_f      = __uint2float_rn((uint32_t) m.weight_sample);
  unsigned long long cta_updated = 0;
  for (uint32_t bi = blockIdx.x; bi < n_vis; bi += gridDim.x) {
    const VisEntry e = m.vis[bi];
    if (e.val & 0x80000000u)
      continue;
    // ---- projection + depth test, registers only ----
    float sdf_new[8];
    uint32_t pix[8];
    unsigned ok    = 0;
    const float py = fmul(i2f(e.y * kBlockSide + ly), m.voxel_size), pz = fmul(i2f(e.z * kBlockSide + lz), m.voxel_size);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const f3 pf = {fmul(i2f(e.x * kBlockSide + j), m.voxel_size), py, pz};
      const f3 pc = se3_mul(pose.Ri, pose.ti, pf);
      int row, col;
      sdf_new[j] = 0.f, pix[j] = 0;
      if (project_point(cam, pc, row, col)) {
        const uint32_t p = (uint32_t) row * cam.cols + (uint32_t) col;
        const float d    = cloud_depth(cam, (uint32_t) row, (uint32_t) col, __ldg(depth + p));
        if (d != 0.f && !(d > m.max_integration_distance)) {
          float sdf     = fsub(d, get_depth(cam, pc));
          const float t = truncation(m.trunc, m.trunc_scale, d);
          if (!(sdf <= -t)) {
            sdf_new[j] = (sdf >= 0.f) ? fminf(t, sdf) : fmaxf(-t, sdf);
            pix[j]     = p;
            ok |= 1u << j;
          }
        }
      }
    }
    const int any  = __syncthreads_or((int) ok);
    uint8_t* base  = m.pool + (size_t) e.val * kBlockBytes;
    float sv[8], qv[8];
    uint32_t cv[8];
    if (any) {
      const float4* ps = reinterpret_cast<const float4*>(base) + 2 * tid;
      const float4* pq = reinterpret_cast<const float4*>(base + kPlaneBytes) + 2 * tid;
      const uint4* pc  = reinterpret_cast<const uint4*>(base + 2 * kPlaneBytes) + 2 * tid;
      const float4 a0 = ps[0], a1 = ps[1], b0 = pq[0], b1 = pq[1];
      const uint4 c0 = pc[0], c1 = pc[1];
      sv[0] = a0.x, sv[1] = a0.y, sv[2] = a0.z, sv[3] = a0.w, sv[4] = a1.x, sv[5] = a1.y, sv[6] = a1.z, sv[7] = a1.w;
      qv[0] = b0.x, qv[1] = b0.y, qv[2] = b0.z, qv[3] = b0.w, qv[4] = b1.x, qv[5] = b1.y, qv[6] = b1.z, qv[7] = b1.w;
      cv[0] = c0.x, cv[1] = c0.y, cv[2] = c0.z, cv[3] = c0.w, cv[4] = c1.x, cv[5] = c1.y, cv[6] = c1.z, cv[7] = c1.w;
      float min_abs  = 3.40282346638528859812e+38f;
      uint32_t max_w = 0, n_upd = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (ok & (1u << j)) {
          // integrateDepthMapKernel :1155-1180 + combineVoxel (voxel_hash_utils.cuh:169-181)
          const uint32_t cw = cv[j];
          const uint32_t w0 = cw >> 24;
          const uint8_t* px = rgb + (size_t) pix[j] * 3;
          const uint32_t r1 = px[0], g1 = px[1], b1c = px[2];
          uint32_t r0 = cw & 0xFF, g0 = (cw >> 8) & 0xFF, b0c = (cw >> 16) & 0xFF;
          if (w0 == 0)
            r0 = r1, g0 = g1, b0c = b1c;
          const float sdf       = sdf_new[j];
          const float curr_mean = w0 > 0 ? sv[j] : sdf;
          const float delta     = fdiv(fsub(sdf, curr_mean), half_size);
          const uint32_t wsum   = w0 + (uint32_t) m.weight_sample;
          const float merged    = fdiv(ffma(sdf, ws_f, fmul(sv[j], __uint2float_rn(w0))), __uint2float_rn(wsum));
          const uint32_t rr     = (uint32_t) f2i(fadd(ffma(__uint2float_rn(r1), 0.5f, fmul(__uint2float_rn(r0), 0.5f)), 0.5f)) & 0xFF;
          const uint32_t gg     = (uint32_t) f2i(fadd(ffma(__uint2float_rn(g1), 0.5f, fmul(__uint2float_rn(g0), 0.5f)), 0.5f)) & 0xFF;
          const uint32_t bb     = (uint32_t) f2i(fadd(ffma(__uint2float_rn(b1c), 0.5f, fmul(__uint2float_rn(b0c), 0.5f)), 0.5f)) & 0xFF;
          const uint32_t wn     = min(wsum, (uint32_t) kWeightMax);
          const float delta2    = fdiv(fsub(sdf, merged), half_size);
          float ss              = fmul(delta, delta2);
          if (fabsf(ss) < 1.175494350822287508e-38f)
            ss = 0.f; // ATOM.ADD.F32.FTZ of the reference flushes a denormal addend
          sv[j] = merged;
          qv[j] = fadd(0.f, ss); // Q1: merged_voxel starts from sum_squared = 0
          cv[j] = rr | (gg << 8) | (bb << 16) | (wn << 24);
          ++n_upd;
        }
        const uint32_t w = cv[j] >> 24;
        if (w != 0)
          min_abs = fminf(min_abs, fabsf(sv[j]));
        max_w = max(max_w, w);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        min_abs = fminf(min_abs, __shfl_xor_sync(0xFFFFFFFFu, min_abs, o));
        max_w   = max(max_w, __shfl_xor_sync(0xFFFFFFFFu, max_w, o));
        n_upd += __shfl_xor_sync(0xFFFFFFFFu, n_upd, o);
      }
      if (lane == 0)
        s_min[warp] = min_abs, s_max[warp] = max_w, s_upd[warp] = n_upd;
    }
    __syncthreads();
    if (tid == 0) {
      BlockStats st;
      if (any) {
        st.min_abs_sdf = fminf(s_min[0], s_min[1]);
        st.max_weight  = max(s_max[0], s_max[1]);
        cta_updated += s_upd[0] + s_upd[1];
      } else {
        st = m.stats[e.val];
      }
      int del = 0;
      if (FUSE_GC) {
        del = gc_predicate(m, st.min_abs_sdf, st.max_weight) ? 1 : 0;
        if (del) {
          free_block(m, f.live_cur, e);
          atomicAdd(&m.ctr->blocks_freed, 1ull);
          st.min_abs_sdf = 3.40282346638528859812e+38f;
          st.max_weight  = 0;
        }
      }
      if (any || del)
        m.stats[e.val] = st;
      s_delete = del;
    }
    __syncthreads();
    const int del = FUSE_GC ? s_delete : 0;
    float4* ws    = reinterpret_cast<float4*>(base) + 2 * tid;
    float4* wq    = reinterpret_cast<float4*>(base + kPlaneBytes) + 2 * tid;
    uint4* wc     = reinterpret_cast<uint4*>(base + 2 * kPlaneBytes) + 2 * tid;
    if (del) {
      // deleteVoxel over the whole block (:1838-1841): free pool blocks are always all-zero
      const float4 z = {0.f, 0.f, 0.f, 0.f};
      const uint4 zu = {0u, 0u, 0u, 0u};
      ws[0] = z, ws[1] = z, wq[0] = z, wq[1] = z, wc[0] = zu, wc[1] = zu;
    } else if (ok) {
      ws[0] = make_float4(sv[0], sv[1], sv[2], sv[3]), ws[1] = make_float4(sv[4], sv[5], sv[6], sv[7]);
      wq[0] = make_float4(qv[0], qv[1], qv[2], qv[3]), wq[1] = make_float4(qv[4], qv[5], qv[6], qv[7]);
      wc[0] = make_uint4(cv[0], cv[1], cv[2], cv[3]), wc[1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
    }
    __syncthreads(); // s_* reused by the next block of this CTA
  }
  if (tid == 0) {
    if (cta_updated)
      atomicAdd(&m.ctr->voxels_updated, cta_updated);
    if (blockIdx.x == 0)
      atomicAdd(&m.ctr->blocks_visible, (unsigned long long) n_vis);
    if (rearm) {
      // every CTA has read vis_count by the time it gets here; the last one re-arms the counters
      // the next frame's k_front appends to (its output live list is this frame's input list)
      __threadfence();
      const unsigned done = atomicAdd(&m.ctr->done_ctas, 1u) + 1u;
      if (done == gridDim.x) {
        m.ctr->live_count[f.live_cur] = 0;
        m.ctr->vis_count              = 0;
        m.ctr->done_ctas              = 0;
      }
    }
  }
}

} // namespace mrh

// mrh_kernels.cuh — sm_100a kernels of the per-frame TSDF hot path (RGB-D).
//
// Replaces, for GeoWrapper::compute() (reference files under /root/reference/mrhash/src/sdf):
//   k_alloc_rgbd      <- calculateCloudKernel (camera.cu:5-19) + allocBlocksKernel x(>=2) +
//                        resetHashBucketMutexKernel + the host retry loop (voxel_data_structures.cu:758-922)
//   k_visible         <- resetCompactHashTableKernel + flatAndReduceHashTableKernel(camera) (:9-14, :406-449)
//   k_integrate       <- integrateDepthMapKernel (:1095-1181) + garbageCollectIdentifyKernel (:1674-1724)
//                        [+ garbageCollectFreeKernel (:1827-1854) when FUSE_GC]
//   k_starve          <- starveVoxelsKernel x2 (:1597-1671)
//   k_identify        <- garbageCollectIdentifyKernel on starve / variance frames
//   k_gc_free         <- garbageCollectFreeKernel + deleteHashEntryElement (:1727-1854)
//
#pragma once
#include "mrh_table.cuh"

namespace mrh {

// Warp-cooperative block walk shared by the RGB-D and point-cloud allocation kernels: every lane
// advances its own DDA; per step the distinct unresolved keys of the warp are inserted once each.
template <bool FRUSTUM_TEST>
__device__ __forceinline__ void alloc_walk(const MapDev& m, const CameraDev& cam, const PoseDev& pose, uint32_t live_cur, DDA& dda, bool active, int lane) {
  const unsigned full = 0xFFFFFFFFu;
  const unsigned n_rays = __popc(__ballot_sync(full, active));
  if (lane == 0 && n_rays)
    atomicAdd(&m.ctr->rays_valid, (unsigned long long) n_rays);
  int iter = 0;
  // Keys this warp has already resolved (present, inserted or rejected by the frustum test): the
  // outcome cannot change within the frame, and neighbouring rays revisit the same few blocks.
  unsigned long long recent[4] = {kNoKey, kNoKey, kNoKey, kNoKey};
  int recent_pos               = 0;
  while (__any_sync(full, active)) {
    unsigned long long key = kNoKey;
    bool want              = false;
    if (active) {
      if (key_in_range(dda.cur)) {
        key  = pack_key(dda.cur);
        want = key != recent[0] && key != recent[1] && key != recent[2] && key != recent[3];
      } else {
        atomicAdd(&m.ctr->dropped_table, 1ull);
      }
    }
    // one leader per distinct unresolved key in the warp
    const unsigned peers = __match_any_sync(full, want ? key : kNoKey);
    const bool leader    = want && (lane == __ffs(peers) - 1);
    unsigned todo        = __ballot_sync(full, leader);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const i3 b = {__shfl_sync(full, dda.cur.x, src), __shfl_sync(full, dda.cur.y, src), __shfl_sync(full, dda.cur.z, src)};
      warp_insert<FRUSTUM_TEST>(m, cam, pose, live_cur, b, lane);
      recent[recent_pos] = pack_key(b);
      recent_pos         = (recent_pos + 1) & 3;
    }
    if (active) {
      active = dda.advance();
      if (++iter >= kMaxDDA)
        active = false;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_alloc_rgbd: one thread per pixel, one warp per 32-pixel row segment.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_alloc_rgbd(MapDev m, FrameDev f, CameraDev cam, const float* __restrict__ depth) {
  __shared__ PoseDev pose;
  if (threadIdx.x == 0) {
    load_pose(f, pose);
    if (blockIdx.x == 0 && blockIdx.y == 0) {
      // lists consumed by later kernels of this frame start empty
      m.ctr->live_count[f.live_cur ^ 1u] = 0;
      m.ctr->vis_count                   = 0;
      m.ctr->n_realloc                   = 0;
    }
  }
  __syncthreads();
  const int lane      = threadIdx.x & 31;
  const int warp      = threadIdx.x >> 5;
  const uint32_t col  = blockIdx.x * 32 + lane;
  const uint32_t row  = blockIdx.y * 8 + warp;
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
  alloc_walk<true>(m, cam, pose, f.live_cur, dda, active, lane);
}

// ---------------------------------------------------------------------------------------------
// k_visible: O(live) frustum pass over the dense live list (no full-table scan).
// ---------------------------------------------------------------------------------------------
// Visibility pass over the input live list of the frame: frustum test (isSDFBlockInCameraFrustumApprox,
// voxel_data_structures.cu:66-77), compaction of the live list, emission of the visible list.
// cta / n_ctas: the caller's position in the set of CTAs that share this work.
__device__ __forceinline__ void visible_pass(const MapDev& m, uint32_t cur, const CameraDev& cam, const PoseDev& pose, int use_frustum, uint32_t cta, uint32_t n_ctas) {
  const unsigned full  = 0xFFFFFFFFu;
  const int lane       = threadIdx.x & 31;
  const uint32_t n     = m.ctr->live_count[cur];
  const LiveEntry* in  = m.live[cur];
  LiveEntry* out       = m.live[cur ^ 1u];
  const uint32_t n_pad = (n + 31u) & ~31u;
  for (uint32_t i = cta * blockDim.x + threadIdx.x; i < n_pad; i += n_ctas * blockDim.x) {
    LiveEntry le = {kEmpty, kInvalid, 0u};
    if (i < n)
      le = in[i];
    const uint32_t slot = le.slot;
    const bool alive    = slot != kInvalid;
    i3 b                = {0, 0, 0};
    bool vis            = false;
    bool maybe          = true;
    if (alive) {
      b   = unpack_key(le.key);
      vis = use_frustum ? block_in_frustum_ex(cam, pose, b, m.voxel_size, maybe) : true;
    }
    const unsigned am = __ballot_sync(full, alive);
    const unsigned vm = __ballot_sync(full, vis);
    uint32_t abase = 0, vbase = 0;
    if (lane == 0) {
      if (am)
        abase = atomicAdd(&m.ctr->live_count[cur ^ 1u], (uint32_t) __popc(am));
      if (vm)
        vbase = atomicAdd(&m.ctr->vis_count, (uint32_t) __popc(vm));
    }
    abase                = __shfl_sync(full, abase, 0);
    vbase                = __shfl_sync(full, vbase, 0);
    const unsigned lt    = (1u << lane) - 1u;
    const uint32_t my_li = abase + __popc(am & lt);
    if (alive)
      out[my_li] = le;
    if (vis) {
      VisEntry e;
      e.x = b.x, e.y = b.y, e.z = b.z;
      e.val      = le.val;
      e.slot     = slot;
      e.live_idx = my_li;
      e.maybe_in_image = maybe ? 1u : 0u;
      e.pad1           = 0;
      m.vis[vbase + __popc(vm & lt)] = e;
    }
  }
}

__global__ void __launch_bounds__(256) k_visible(MapDev m, FrameDev f, CameraDev cam, int use_frustum) {
  __shared__ PoseDev pose;
  if (threadIdx.x == 0)
    load_pose(f, pose);
  __syncthreads();
  visible_pass(m, f.live_cur, cam, pose, use_frustum, blockIdx.x, gridDim.x);
}

// Remove a block from the map: tombstone its key, return its pool block to the free stack
// (appendHeapHigh :52-56), drop it from the live list. Called by one thread.
__device__ __forceinline__ void free_block(const MapDev& m, uint32_t live_cur, const VisEntry& e) {
  atomicExch(m.keys + e.slot, kTomb);
  const int addr   = atomicAdd(&m.ctr->heap_counter, 1);
  m.heap[addr + 1] = e.val & 0x7FFFFFFFu;
  m.live[live_cur ^ 1u][e.live_idx].slot = kInvalid;
}

__device__ __forceinline__ bool gc_predicate(const MapDev& m, float min_abs, uint32_t max_w) {
  return min_abs >= m.gc_threshold || max_w == 0u; // voxel_data_structures.cu:1711
}

// ---------------------------------------------------------------------------------------------
// k_integrate: one CTA (kIntThreads threads x kIntVox consecutive-x voxels) per visible block, persistent.
// ---------------------------------------------------------------------------------------------
// rearm: the last CTA to finish zeroes the list counters the next frame's k_front appends to (every
// CTA has read vis_count by then), so a frame needs no reset kernel.
// Consecutive-x voxels per thread: 2 (256 threads per block, 64-bit plane access) with a 40-register
// budget keeps 48 warps per SM resident; 4 (128 threads, 128-bit access, 56 registers) only 36. The
// kernel is latency-bound, so the extra warps win: 45.4 vs 47.4 us per frame (profiles/r1_summary.md).
#ifndef MRH_INTEGRATE_VOX
#define MRH_INTEGRATE_VOX 2
#endif
#ifndef MRH_INTEGRATE_MIN_CTAS
#define MRH_INTEGRATE_MIN_CTAS (MRH_INTEGRATE_VOX == 2 ? 6 : 1) // resident CTAs per SM the register budget must allow
#endif
constexpr int kIntVox     = MRH_INTEGRATE_VOX;
constexpr int kIntThreads = kBlockVoxels / kIntVox;
constexpr int kIntWarps   = kIntThreads / 32;
template <int V>
struct PlaneVec;
template <>
struct PlaneVec<4> {
  using F = float4;
  using U = uint4;
};
template <>
struct PlaneVec<2> {
  using F = float2;
  using U = uint2;
};
template <bool FUSE_GC>
__global__ void __launch_bounds__(kIntThreads, MRH_INTEGRATE_MIN_CTAS) k_integrate(MapDev m, FrameDev f, CameraDev cam, const float* __restrict__ depth, const uint8_t* __restrict__ rgb, int rearm) {
  using VF = typename PlaneVec<kIntVox>::F;
  using VU = typename PlaneVec<kIntVox>::U;
  __shared__ float s_min[kIntWarps];
  __shared__ uint32_t s_max[kIntWarps];
  __shared__ uint32_t s_upd[kIntWarps];
  __shared__ int s_delete;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cudaGridDependencySynchronize(); // chained launch: k_front has completed and its lists are visible (no-op otherwise)
  const PoseDev& pose  = frame_pose(f);
  const uint32_t n_vis = m.ctr->vis_count;
  constexpr int kPerRow = kBlockSide / kIntVox; // threads per row of 8 voxels
  const int lx0 = (tid % kPerRow) * kIntVox, ly = (tid / kPerRow) & 7, lz = tid / (kPerRow * 8);
  const float half_size = fmul(m.voxel_size, 0.5f);
  unsigned long long cta_updated = 0;
  for (uint32_t bi = blockIdx.x; bi < n_vis; bi += gridDim.x) {
    const VisEntry e = m.vis[bi];
    if (e.val & 0x80000000u)
      continue; // resolution-1 blocks are fused by k_integrate_lowres
    uint8_t* base = m.pool + (size_t) e.val * kBlockBytes;
    if (e.maybe_in_image && (tid % (32 / kIntVox)) == 0) {
      // the three planes are read after the projection pass: start moving their lines (one per 8
      // threads) towards L2 now, the projection and the depth gathers hide the DRAM latency
      prefetch_l2(base + 4 * kIntVox * tid);
      prefetch_l2(base + kPlaneBytes + 4 * kIntVox * tid);
      prefetch_l2(base + 2 * kPlaneBytes + 4 * kIntVox * tid);
    }
    // ---- pass 1: projection + depth test, registers only ----
    float sdf_new[kIntVox] = {};
    uint32_t pix[kIntVox]  = {};
    unsigned ok      = 0;
    if (e.maybe_in_image) { // else: no voxel of this block can project into the image (k_visible)
    const f3 pf_yz = {0.f, fmul(i2f(e.y * kBlockSide + ly), m.voxel_size), fmul(i2f(e.z * kBlockSide + lz), m.voxel_size)};
#pragma unroll
    for (int j = 0; j < kIntVox; ++j) {
      const f3 pf = {fmul(i2f(e.x * kBlockSide + lx0 + j), m.voxel_size), pf_yz.y, pf_yz.z};
      const f3 pc = se3_mul(pose.Ri, pose.ti, pf);
      int row, col;
      if (project_point(cam, pc, row, col)) {
        const uint32_t p = (uint32_t) row * cam.cols + (uint32_t) col;
        const float d    = cloud_depth(cam, (uint32_t) row, (uint32_t) col, __ldg(depth + p));
        if (d != 0.f && !(d > m.max_integration_distance)) {
          float sdf     = fsub(d, get_depth(cam, pc));
          const float t = truncation(m.trunc, m.trunc_scale, d);
          if (!(sdf <= -t)) {
            sdf        = (sdf >= 0.f) ? fminf(t, sdf) : fmaxf(-t, sdf);
            sdf_new[j] = sdf;
            pix[j]     = p;
            ok |= 1u << j;
          }
        }
      }
    }
    }
    const int any = e.maybe_in_image ? __syncthreads_or((int) ok) : 0;
    float min_abs  = 3.40282346638528859812e+38f;
    uint32_t max_w = 0, n_upd = 0;
    VF sdf4, ss4;
    VU cw4;
    if (any) {
      // ---- pass 2: vector loads of this thread's voxels from the three planes ----
      sdf4 = reinterpret_cast<const VF*>(base)[tid];
      ss4  = reinterpret_cast<const VF*>(base + kPlaneBytes)[tid];
      cw4  = reinterpret_cast<const VU*>(base + 2 * kPlaneBytes)[tid];
      float* sdfv    = reinterpret_cast<float*>(&sdf4);
      float* ssv     = reinterpret_cast<float*>(&ss4);
      uint32_t* cwv  = reinterpret_cast<uint32_t*>(&cw4);
#pragma unroll
      for (int j = 0; j < kIntVox; ++j) {
        if (ok & (1u << j)) {
          // integrateDepthMapKernel :1155-1180 + combineVoxel (voxel_hash_utils.cuh:169-181)
          const uint32_t cw = cwv[j];
          const uint32_t w0 = cw >> 24;
          const uint8_t* px = rgb + (size_t) pix[j] * 3;
          const uint32_t r1 = px[0], g1 = px[1], b1 = px[2];
          uint32_t r0 = cw & 0xFF, g0 = (cw >> 8) & 0xFF, b0 = (cw >> 16) & 0xFF;
          if (w0 == 0)
            r0 = r1, g0 = g1, b0 = b1;
          const float sdf       = sdf_new[j];
          const float curr_mean = w0 > 0 ? sdfv[j] : sdf;
          const float delta     = fdiv(fsub(sdf, curr_mean), half_size);
          const uint32_t wsum   = w0 + (uint32_t) m.weight_sample;
          const float merged    = fdiv(ffma(sdf, __uint2float_rn((uint32_t) m.weight_sample), fmul(sdfv[j], __uint2float_rn(w0))), __uint2float_rn(wsum));
          const uint32_t rr     = (uint32_t) f2i(fadd(ffma(__uint2float_rn(r1), 0.5f, fmul(__uint2float_rn(r0), 0.5f)), 0.5f)) & 0xFF;
          const uint32_t gg     = (uint32_t) f2i(fadd(ffma(__uint2float_rn(g1), 0.5f, fmul(__uint2float_rn(g0), 0.5f)), 0.5f)) & 0xFF;
          const uint32_t bb     = (uint32_t) f2i(fadd(ffma(__uint2float_rn(b1), 0.5f, fmul(__uint2float_rn(b0), 0.5f)), 0.5f)) & 0xFF;
          const uint32_t wn     = min(wsum, (uint32_t) kWeightMax);
          const float delta2    = fdiv(fsub(sdf, merged), half_size);
          float ss              = fmul(delta, delta2);
          if (fabsf(ss) < 1.175494350822287508e-38f)
            ss = 0.f; // ATOM.ADD.F32.FTZ of the reference flushes a denormal addend
          sdfv[j] = merged;
          ssv[j]  = fadd(0.f, ss); // Q1: merged_voxel starts from sum_squared = 0
          cwv[j]  = rr | (gg << 8) | (bb << 16) | (wn << 24);
          ++n_upd;
        }
        const uint32_t w = cwv[j] >> 24;
        if (w != 0)
          min_abs = fminf(min_abs, fabsf(sdfv[j]));
        max_w = max(max_w, w);
      }
      // block-level reduction of the GC statistics
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
        st.min_abs_sdf = s_min[0], st.max_weight = s_max[0];
        uint32_t upd   = s_upd[0];
#pragma unroll
        for (int w = 1; w < kIntWarps; ++w) {
          st.min_abs_sdf = fminf(st.min_abs_sdf, s_min[w]);
          st.max_weight  = max(st.max_weight, s_max[w]);
          upd += s_upd[w];
        }
        cta_updated += upd;
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
    if (del) {
      // deleteVoxel over the whole block (:1838-1841): free pool blocks are always all-zero
      VF z;
      memset(&z, 0, sizeof(z));
      reinterpret_cast<VF*>(base)[tid]                   = z;
      reinterpret_cast<VF*>(base + kPlaneBytes)[tid]     = z;
      reinterpret_cast<VF*>(base + 2 * kPlaneBytes)[tid] = z;
    } else if (ok) {
      reinterpret_cast<VF*>(base)[tid]                   = sdf4;
      reinterpret_cast<VF*>(base + kPlaneBytes)[tid]     = ss4;
      reinterpret_cast<VU*>(base + 2 * kPlaneBytes)[tid] = cw4;
    }
    __syncthreads(); // s_* reused by the next block of this CTA
  }
  if (tid == 0) {
    if (cta_updated)
      atomicAdd(&m.ctr->voxels_updated, cta_updated);
    if (blockIdx.x == 0)
      atomicAdd(&m.ctr->blocks_visible, (unsigned long long) n_vis);
    if (rearm) {
      __threadfence();
      const unsigned done = atomicAdd(&m.ctr->done_ctas, 1u) + 1u;
      if (done == gridDim.x) {
        m.ctr->live_count[f.live_cur] = 0; // next frame's output list
        m.ctr->vis_count              = 0;
        m.ctr->done_ctas              = 0;
        // paging trigger (GeoWrapper::compute, geowrapper.cpp:137): the host looks at these two words
        // before the next frame, no copy engine operation and no wait involved
        if (f.pad[0]) { // host asks for it every 16th frame, every frame once the pool is below 40 % free
          m.host_probe[0] = *reinterpret_cast<volatile int*>(&m.ctr->heap_counter);
          m.host_probe[1] = (int) f.frame_index;
          __threadfence_system();
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// starve (every n-th frame): z-buffer pass then decrement pass
// ---------------------------------------------------------------------------------------------
template <int PASS>
__global__ void __launch_bounds__(128) k_starve(MapDev m, FrameDev f, CameraDev cam) {
  __shared__ PoseDev pose;
  const int tid = threadIdx.x;
  if (tid == 0)
    load_pose(f, pose);
  __syncthreads();
  const uint32_t n_vis = m.ctr->vis_count;
  for (uint32_t bi = blockIdx.x; bi < n_vis; bi += gridDim.x) {
    const VisEntry e = m.vis[bi];
    if (e.val & 0x80000000u)
      continue;
    if ((bi >> (m.starve_id_bits - 9u)) != 0u) {
      // the id 512 * bi + i would run into the rank field: two voxels would share a z-buffer value
      if (tid == 0)
        *reinterpret_cast<volatile uint32_t*>(&m.ctr->fault) = 2u;
      continue;
    }
    uint8_t* base = m.pool + (size_t) e.val * kBlockBytes;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const int i  = tid * 4 + j;
      const int lx = i & 7, ly = (i >> 3) & 7, lz = i >> 6;
      const f3 pf  = {fmul(i2f(e.x * kBlockSide + lx), m.voxel_size), fmul(i2f(e.y * kBlockSide + ly), m.voxel_size),
                      fmul(i2f(e.z * kBlockSide + lz), m.voxel_size)};
      const f3 pc   = se3_mul(pose.Ri, pose.ti, pf);
      const float d = get_depth(cam, pc);
      if (d < cam.min_depth)
        continue;
      int row, col;
      if (!project_point(cam, pc, row, col))
        continue;
      // pack(unique_tid, depth) (:1583-1585, :1629-1641)
      const unsigned long long packed = ((unsigned long long) __float_as_uint(d) << 32) + (unsigned long long) (m.shard_tag | (uint32_t) (kBlockVoxels * bi + i));
      unsigned long long* cell = m.zbuf + (size_t) row * cam.cols + col;
      if (PASS == 0) {
        atomicMin(cell, packed);
      } else if (*cell == packed) {
        uint8_t* w = base + 2 * kPlaneBytes + 4 * i + 3;
        *w         = (uint8_t) max(0, (int) *w - 1);
      }
    }
  }
}

// end-of-frame probe for the paths whose last kernel does not write it (see k_integrate)
__global__ void k_probe(MapDev m, uint32_t frame_index) {
  m.host_probe[0] = m.ctr->heap_counter;
  m.host_probe[1] = (int) frame_index;
  __threadfence_system();
}

// recompute the GC statistics of every visible block from its payload (starve / variance frames)
__global__ void __launch_bounds__(128) k_identify(MapDev m) {
  __shared__ float s_min[4];
  __shared__ uint32_t s_max[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_vis = m.ctr->vis_count;
  for (uint32_t bi = blockIdx.x; bi < n_vis; bi += gridDim.x) {
    const VisEntry e = m.vis[bi];
    if (e.val & 0x80000000u)
      continue;
    const uint8_t* base = m.pool + (size_t) e.val * kBlockBytes;
    const float4 sdf4   = reinterpret_cast<const float4*>(base)[tid];
    const uint4 cw4     = reinterpret_cast<const uint4*>(base + 2 * kPlaneBytes)[tid];
    const float sv[4]   = {sdf4.x, sdf4.y, sdf4.z, sdf4.w};
    const uint32_t cv[4] = {cw4.x, cw4.y, cw4.z, cw4.w};
    float min_abs  = 3.40282346638528859812e+38f;
    uint32_t max_w = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t w = cv[j] >> 24;
      if (w)
        min_abs = fminf(min_abs, fabsf(sv[j]));
      max_w = max(max_w, w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      min_abs = fminf(min_abs, __shfl_xor_sync(0xFFFFFFFFu, min_abs, o));
      max_w   = max(max_w, __shfl_xor_sync(0xFFFFFFFFu, max_w, o));
    }
    if (lane == 0)
      s_min[warp] = min_abs, s_max[warp] = max_w;
    __syncthreads();
    if (tid == 0) {
      BlockStats st;
      st.min_abs_sdf = fminf(fminf(s_min[0], s_min[1]), fminf(s_min[2], s_min[3]));
      st.max_weight  = max(max(s_max[0], s_max[1]), max(s_max[2], s_max[3]));
      m.stats[e.val] = st;
    }
    __syncthreads();
  }
}

// apply the GC decision (stats must be current for every visible block)
__global__ void __launch_bounds__(128) k_gc_free(MapDev m, FrameDev f) {
  const int tid        = threadIdx.x;
  const uint32_t n_vis = m.ctr->vis_count;
  for (uint32_t bi = blockIdx.x; bi < n_vis; bi += gridDim.x) {
    const VisEntry e = m.vis[bi];
    if (e.val & 0x80000000u)
      continue;
    if (m.keys[e.slot] >= kNoKey)
      continue; // already removed this frame (variance path)
    const BlockStats st = m.stats[e.val];
    if (!gc_predicate(m, st.min_abs_sdf, st.max_weight))
      continue;
    uint8_t* base  = m.pool + (size_t) e.val * kBlockBytes;
    const float4 z = {0.f, 0.f, 0.f, 0.f};
    reinterpret_cast<float4*>(base)[tid]                   = z;
    reinterpret_cast<float4*>(base + kPlaneBytes)[tid]     = z;
    reinterpret_cast<float4*>(base + 2 * kPlaneBytes)[tid] = z;
    if (tid == 0) {
      free_block(m, f.live_cur, e);
      m.stats[e.val] = {3.40282346638528859812e+38f, 0u};
      atomicAdd(&m.ctr->blocks_freed, 1ull);
    }
  }
}

} // namespace mrh

// mrh_points.cuh — LiDAR / point-cloud integration kernels.
//
// Replaces allocBlocks3DKernel (voxel_data_structures.cu:925-1033) and integrate3DKernel
// (:1215-1379). The reference updates voxels with a plain load / modify / store from one thread per
// point, so several points racing on one voxel make its output run-to-run nondeterministic
// (SURVEY.md H3). Here the walk and the update are split:
//   k_points_emit  : one thread per point walks its voxel DDA and writes (voxel address, sdf)
//                    records into the point's own fixed run of slots (point-major order, unused
//                    slots keep the all-ones hole key) - the records do not depend on voxel contents;
//   (STABLE radix sort by voxel address: equal addresses stay in point order, holes sink to the end;
//    the item count is points x slots, known on the host, so the frame needs no read-back)
//   k_points_apply : one thread per distinct voxel folds its records in point order.
// The result is deterministic and equals the reference's semantics with the points applied one
// after another in index order (the CPU oracle's order).
#pragma once
#include "mrh_table.cuh"

namespace mrh {


// ray end points of one point (shared by the allocation and integration walks)
// for_alloc: allocBlocks3DKernel :939-961; else integrate3DKernel :1229-1254.
// nrm: the point's normal (the reference reads its `normals` array at 3 * point_idx, :937 / :1227: the
// first eigenvector of the point); only used when projective_sdf is off, where both kernels lay the
// ray along the normal (:959-960, :1252-1253). norm_dir returns the normalised normal.
__device__ __forceinline__ bool point_ray(const MapDev& m, const PoseDev& pose, f3 pcam, f3 nrm, bool for_alloc, float& range, float& trunc, f3& pw_min, f3& pw_max, f3& norm_dir) {
  range = norm3df(pcam.x, pcam.y, pcam.z);
  if (for_alloc) {
    if (range == 0.f)
      return false;
  } else if ((double) range < 1e-6 || range > m.max_integration_distance) {
    return false;
  }
  const f3 dir     = normalize3(pcam);
  trunc            = truncation(m.trunc, m.trunc_scale, range);
  const float dmin = fminf(m.max_integration_distance, fsub(range, trunc));
  const float dmax = fminf(m.max_integration_distance, fadd(range, trunc));
  if (dmin >= dmax)
    return false;
  f3 a, b;
  norm_dir = {0.f, 0.f, 0.f};
  if (!m.projective) {
    norm_dir       = normalize3(nrm);
    const float ka = fsub(dmin, range), kb = fsub(dmax, range);
    a = {ffma(norm_dir.x, ka, pcam.x), ffma(norm_dir.y, ka, pcam.y), ffma(norm_dir.z, ka, pcam.z)};
    b = {ffma(norm_dir.x, kb, pcam.x), ffma(norm_dir.y, kb, pcam.y), ffma(norm_dir.z, kb, pcam.z)};
  } else if (for_alloc) {
    const float ka = fsub(dmin, range), kb = fsub(dmax, range);
    a = {ffma(dir.x, ka, pcam.x), ffma(dir.y, ka, pcam.y), ffma(dir.z, ka, pcam.z)};
    b = {ffma(dir.x, kb, pcam.x), ffma(dir.y, kb, pcam.y), ffma(dir.z, kb, pcam.z)};
  } else {
    a = {ffma(-dir.x, trunc, pcam.x), ffma(-dir.y, trunc, pcam.y), ffma(-dir.z, trunc, pcam.z)};
    b = {ffma(dir.x, trunc, pcam.x), ffma(dir.y, trunc, pcam.y), ffma(dir.z, trunc, pcam.z)};
  }
  pw_min = se3_mul(pose.R, pose.t, a);
  pw_max = se3_mul(pose.R, pose.t, b);
  return true;
}

__device__ __forceinline__ f3 point_normal(const float* __restrict__ normals, uint32_t i) {
  if (!normals)
    return {0.f, 0.f, 0.f};
  return {__ldg(normals + 3 * (size_t) i), __ldg(normals + 3 * (size_t) i + 1), __ldg(normals + 3 * (size_t) i + 2)};
}

__global__ void __launch_bounds__(256) k_alloc_points(MapDev m, FrameDev f, CameraDev cam, const float* __restrict__ points, const float* __restrict__ normals, uint32_t n_points) {
  __shared__ PoseDev pose;
  if (threadIdx.x == 0) {
    load_pose(f, pose);
    if (blockIdx.x == 0) {
      m.ctr->live_count[f.live_cur ^ 1u] = 0;
      m.ctr->vis_count                   = 0;
      m.ctr->n_updates                   = 0;
      m.ctr->n_realloc                   = 0;
    }
  }
  __syncthreads();
  const int lane   = threadIdx.x & 31;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool active      = false;
  DDA dda;
  if (i < n_points) {
    const f3 p = {__ldg(points + 3 * (size_t) i), __ldg(points + 3 * (size_t) i + 1), __ldg(points + 3 * (size_t) i + 2)};
    float range, trunc;
    f3 a, b, nd;
    if (point_ray(m, pose, p, point_normal(normals, i), true, range, trunc, a, b, nd)) {
      dda.init(a, b, m.voxel_size, m.ext, true);
      active = true;
    }
  }
  alloc_walk<false>(m, cam, pose, f.live_cur, dda, active, lane);
}

// One thread per point: voxel-level DDA, one record per visited voxel of an allocated block.
// K = uint32_t while the pool address (+ hole key) fits 32 bits, else unsigned long long.
template <typename K>
__global__ void __launch_bounds__(256) k_points_emit(MapDev m, FrameDev f, const float* __restrict__ points, const float* __restrict__ normals, uint32_t n_points, K* __restrict__ keys, float* __restrict__ vals, uint32_t slots) {
  __shared__ PoseDev pose;
  if (threadIdx.x == 0)
    load_pose(f, pose);
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points || m.ctr->vis_count == 0) // integrate3D runs only when current_occupied_blocks_ > 0 (:1387)
    return;
  const f3 p = {__ldg(points + 3 * (size_t) i), __ldg(points + 3 * (size_t) i + 1), __ldg(points + 3 * (size_t) i + 2)};
  float range, trunc;
  f3 a, b, nd;
  if (!point_ray(m, pose, p, point_normal(normals, i), false, range, trunc, a, b, nd))
    return;
  DDA dda;
  dda.init(a, b, m.voxel_size, m.ext, false);
  K* my_keys     = keys + (size_t) i * slots;
  float* my_vals = vals + (size_t) i * slots;
  uint32_t n_out = 0, n_lost = 0;
  // consecutive voxels of a ray mostly share a block: remember the last lookup
  i3 last_b         = {INT_MIN, 0, 0};
  uint32_t last_val = kInvalid;
#pragma unroll 1
  for (int iter = 0; iter < kMaxDDA; ++iter) {
    const i3 v  = dda.cur;
    const i3 bl = voxel_to_block(v, m.voxel_size, m.ext);
    if (bl.x != last_b.x || bl.y != last_b.y || bl.z != last_b.z) {
      const int slot = table_find(m, bl);
      last_val       = slot >= 0 ? m.vals[slot] : kInvalid;
      last_b         = bl;
    }
    if (last_val != kInvalid) {
      const int r     = (int) (last_val >> 31);
      const int scale = 1 << r;
      const float vs  = fmul(m.voxel_size, i2f(scale));
      const f3 vp     = {fmul(i2f(v.x / scale), vs), fmul(i2f(v.y / scale), vs), fmul(i2f(v.z / scale), vs)};
      const f3 vc     = se3_mul(pose.Ri, pose.ti, vp);
      float sdf;
      if (m.projective)
        sdf = fsub(range, norm3df(vc.x, vc.y, vc.z));
      else // :1320: dot(voxel_pos_camera - pcam, norm_dir)
        sdf = dot3({fsub(vc.x, p.x), fsub(vc.y, p.y), fsub(vc.z, p.z)}, nd);
      if (sdf <= -trunc)
        break;
      sdf = (sdf >= 0.f) ? fminf(trunc, sdf) : fmaxf(-trunc, sdf);
      // reference pool address: entry.ptr + virtualVoxelPosToSDFBlockIndex(v, 8 / scale) (:1343)
      int lx = v.x % 8, ly = v.y % 8, lz = v.z % 8;
      lx += lx < 0 ? 8 : 0, ly += ly < 0 ? 8 : 0, lz += lz < 0 ? 8 : 0;
      lx >>= r, ly >>= r, lz >>= r;
      const unsigned long long addr = (r ? (unsigned long long) (last_val & 0x7FFFFFFFu) * 64ull : (unsigned long long) last_val * 512ull) + (unsigned long long) (lz * 64 + ly * 8 + lx);
      if (n_out < slots) {
        my_keys[n_out] = (K) addr;
        my_vals[n_out] = sdf;
        ++n_out;
      } else {
        ++n_lost;
      }
    }
    if (!dda.advance())
      break;
  }
  if (n_lost)
    atomicAdd(&m.ctr->dropped_updates, (unsigned long long) n_lost);
}

// pointers to the three words of the voxel at a reference pool address (see read_voxel_addr)
__device__ __forceinline__ bool voxel_words(const MapDev& m, unsigned long long addr, float*& sdf, float*& ss, uint32_t*& cw) {
  const unsigned long long P = addr >> 9;
  const uint32_t w           = (uint32_t) (addr & 511ull);
  if (P >= m.num_blocks)
    return false;
  uint8_t* base = m.pool + (size_t) P * kBlockBytes;
  if (m.carved[P]) {
    base += (w >> 6) * 768u;
    sdf = reinterpret_cast<float*>(base) + (w & 63u);
    ss  = reinterpret_cast<float*>(base + 256) + (w & 63u);
    cw  = reinterpret_cast<uint32_t*>(base + 512) + (w & 63u);
  } else {
    sdf = reinterpret_cast<float*>(base) + w;
    ss  = reinterpret_cast<float*>(base + kPlaneBytes) + w;
    cw  = reinterpret_cast<uint32_t*>(base + 2 * kPlaneBytes) + w;
  }
  return true;
}

// Folds the sorted records into the voxels: one lane per run of equal voxel address (holes sorted to
// the end). A warp first finds the run heads of a 256-record tile with ballots and queues them in
// shared memory, then its lanes take the queued runs round-robin - with ~5 records per run, one
// thread per record would leave 4 of 5 lanes idle while the heads loop.
constexpr int kApplyTile = 256; // records per warp tile
template <typename K>
__global__ void __launch_bounds__(256) k_points_apply(MapDev m, const K* __restrict__ keys, const float* __restrict__ vals, uint32_t n, K hole) {
  __shared__ uint32_t s_heads[8][kApplyTile];
  const unsigned full = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float half = fmul(m.voxel_size, 0.5f);
  const float wn_f = __uint2float_rn((uint32_t) m.weight_sample);
  unsigned long long updated = 0;
  const uint32_t n_tiles     = (n + kApplyTile - 1) / kApplyTile;
  for (uint32_t tile = blockIdx.x * 8 + warp; tile < n_tiles; tile += gridDim.x * 8) {
    const uint32_t base = tile * kApplyTile;
    if (keys[base] == hole)
      break; // sorted: nothing but holes from here on, for every later tile of this warp too
    uint32_t n_heads = 0;
#pragma unroll
    for (int r = 0; r < kApplyTile / 32; ++r) {
      const uint32_t i = base + r * 32 + lane;
      const K k        = i < n ? keys[i] : hole;
      const K prev     = i > 0 && i <= n ? keys[i - 1] : hole;
      const bool head  = k != hole && (i == 0 || prev != k);
      const unsigned hm = __ballot_sync(full, head);
      if (head)
        s_heads[warp][n_heads + __popc(hm & ((1u << lane) - 1u))] = i;
      n_heads += __popc(hm);
    }
    __syncwarp();
    for (uint32_t hq = lane; hq < n_heads; hq += 32) {
      const uint32_t i = s_heads[warp][hq];
      const K addr     = keys[i];
      float *psdf, *pss;
      uint32_t* pcw;
      if (!voxel_words(m, (unsigned long long) addr, psdf, pss, pcw))
        continue;
      float sdf0  = *psdf;
      float ss0   = *pss;
      uint32_t cw = *pcw;
      for (uint32_t j = i; j < n && keys[j] == addr; ++j) {
        // integrate3DKernel :1333-1357 + combineVoxel (voxel_hash_utils.cuh:169-181), no colour input
        const float sdf       = vals[j];
        const uint32_t w0     = cw >> 24;
        const float curr_mean = w0 > 0 ? sdf0 : 0.f;
        const float delta     = fdiv(fsub(sdf, curr_mean), half);
        const uint32_t wsum   = w0 + (uint32_t) m.weight_sample;
        const float merged    = fdiv(ffma(sdf, wn_f, fmul(sdf0, __uint2float_rn(w0))), __uint2float_rn(wsum));
        const uint32_t r0 = cw & 0xFF, g0 = (cw >> 8) & 0xFF, b0 = (cw >> 16) & 0xFF;
        const uint32_t rr = (uint32_t) f2i(fadd(ffma(0.f, 0.5f, fmul(__uint2float_rn(r0), 0.5f)), 0.5f)) & 0xFF;
        const uint32_t gg = (uint32_t) f2i(fadd(ffma(0.f, 0.5f, fmul(__uint2float_rn(g0), 0.5f)), 0.5f)) & 0xFF;
        const uint32_t bb = (uint32_t) f2i(fadd(ffma(0.f, 0.5f, fmul(__uint2float_rn(b0), 0.5f)), 0.5f)) & 0xFF;
        const uint32_t wn = min(wsum, (uint32_t) kWeightMax);
        const float delta2 = fdiv(fsub(sdf, merged), half);
        float ss           = fmul(delta, delta2);
        if (fabsf(ss) < 1.175494350822287508e-38f)
          ss = 0.f;
        sdf0 = merged;
        ss0  = fadd(0.f, ss);
        cw   = rr | (gg << 8) | (bb << 16) | (wn << 24);
        ++updated;
      }
      *psdf = sdf0, *pss = ss0, *pcw = cw;
    }
    __syncwarp();
  }
  // warp-aggregated statistics
  for (int o = 16; o > 0; o >>= 1)
    updated += __shfl_xor_sync(0xFFFFFFFFu, updated, o);
  if (lane == 0 && updated)
    atomicAdd(&m.ctr->voxels_updated, updated);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    atomicAdd(&m.ctr->blocks_visible, (unsigned long long) m.ctr->vis_count);
}

} // namespace mrh

// mrh_mesh.cuh — marching cubes over the block map.
//
// Replaces extractIsoSurfaceKernel / extractIsoSurfaceAtPosition / checkVertexVoxels
// (marching_cubes.cu:7-285), trilinearInterpolation / getVoxel / getVoxelSize
// (voxel_data_structures.cu:163-338) and vertexInterp / appendTriangle (mesh_extractor.cu:6-55).
//
// The reference runs one thread per voxel and resolves ~100 voxel reads per voxel through the hash
// (each a 10-slot bucket scan plus list walk, device functions in another translation unit).
// k_mc_blocks runs one CTA per allocated block: the 27 neighbour blocks are looked up once, a 10^3
// halo of (sdf, rgbw) is staged in shared memory, every voxel evaluates its 8 cell corners from
// shared memory, and triangles are emitted with a CTA prefix sum and ONE atomicAdd per CTA pass.
// A block whose neighbourhood holds a resolution-1 block takes the generic sampler, which restates
// the reference's per-read hash walk literally (including its quirks).
//
// Float arithmetic follows the reference build's instruction sequence (explicit intrinsics, see
// mrh_math.cuh); the contraction pattern of the trilinear polynomial and of vertexInterp was read
// from the SASS of the reference objects.
#pragma once
#ifndef MRH_MC_TIGHT
#define MRH_MC_TIGHT 1 // second, tighter pre-filter of k_mc_blocks (corner means)
#endif
#include "mrh_mc_tables.cuh"
#include "mrh_table.cuh"

namespace mrh {

struct VoxVal {
  float sdf;
  uint32_t cw; // r | g << 8 | b << 16 | weight << 24
};

struct TriVertex {
  f3 p, c;
};

// ---- generic sampler: every read goes through the hash, as in the reference -------------------
// Voxel at a reference pool address (in voxels). The reference addresses resolution-1 payloads with
// the 8-wide linearisation (virtualVoxelPosToSDFBlockIndex, voxel_hash_utils.cuh:110-128), i.e. it
// reads past the block's own 64 voxels; the address map keeps that behaviour defined here.
__device__ __forceinline__ VoxVal read_voxel_addr(const MapDev& m, unsigned long long addr) {
  const unsigned long long P = addr >> 9;
  const uint32_t w           = (uint32_t) (addr & 511ull);
  if (P >= m.num_blocks)
    return {0.f, 0u};
  const uint8_t* base = m.pool + (size_t) P * kBlockBytes;
  if (m.carved[P]) {
    base += (w >> 6) * 768u;
    return {reinterpret_cast<const float*>(base)[w & 63u], reinterpret_cast<const uint32_t*>(base + 512)[w & 63u]};
  }
  return {reinterpret_cast<const float*>(base)[w], reinterpret_cast<const uint32_t*>(base + 2 * kPlaneBytes)[w]};
}

struct HashSampler {
  static constexpr bool kRollCorners = false;
  const MapDev& m;
  __device__ __forceinline__ bool block_val(i3 b, uint32_t& val) const {
    const int slot = table_find(m, b);
    if (slot < 0)
      return false;
    val = m.vals[slot];
    return true;
  }
  // getVoxelSize(float3) (:236-240)
  __device__ __forceinline__ float voxel_size_point(f3 p) const {
    uint32_t val;
    const int res = block_val(world_to_block(p, m.voxel_size, m.ext), val) ? (int) (val >> 31) : 0;
    return fmul(m.voxel_size, i2f(1 << res));
  }
  // resolution of getHashEntry(worldPointToSDFBlock(voxel_size, ...)) (:264)
  __device__ __forceinline__ int block_res_scaled(f3 p, float vs) const {
    uint32_t val;
    return block_val(world_to_block(p, vs, m.ext), val) ? (int) (val >> 31) : 0;
  }
  // getVoxel(int3[, res]) (:163-195)
  __device__ __forceinline__ VoxVal voxel_at(i3 v, int* res) const {
    uint32_t val;
    if (!block_val(voxel_to_block(v, m.voxel_size, m.ext), val))
      return {0.f, 0u};
    const int r = (int) (val >> 31);
    if (res)
      *res = r;
    int lx = v.x % 8, ly = v.y % 8, lz = v.z % 8;
    lx += lx < 0 ? 8 : 0, ly += ly < 0 ? 8 : 0, lz += lz < 0 ? 8 : 0;
    lx >>= r, ly >>= r, lz >>= r;
    const uint32_t idx = (uint32_t) (lz * 64 + ly * 8 + lx);
    const unsigned long long base = r ? (unsigned long long) (val & 0x7FFFFFFFu) * 64ull : (unsigned long long) val * 512ull;
    return read_voxel_addr(m, base + idx);
  }
  __device__ __forceinline__ VoxVal voxel_point(f3 p, int* res) const {
    return voxel_at(world_to_voxel(p, m.voxel_size), res);
  }
};

// ---- shared-memory sampler: 10^3 halo around one resolution-0 block whose 27-neighbourhood holds
// no resolution-1 block (so every voxel size query answers the base size) ------------------------
#ifndef MRH_MC_ROLL
#define MRH_MC_ROLL 1
#endif
struct HaloSampler {
  static constexpr bool kRollCorners = MRH_MC_ROLL != 0;
  const MapDev& m;
  const float* s_sdf;   // [1000]
  const uint32_t* s_cw; // [1000]
  i3 origin;            // voxel coordinate of halo cell (0,0,0) = 8 * block - 1
  __device__ __forceinline__ float voxel_size_point(f3) const {
    return m.voxel_size;
  }
  __device__ __forceinline__ int block_res_scaled(f3, float) const {
    return 0;
  }
  __device__ __forceinline__ VoxVal voxel_at(i3 v, int*) const {
    const int i = min(max(v.x - origin.x, 0), 9), j = min(max(v.y - origin.y, 0), 9), k = min(max(v.z - origin.z, 0), 9);
    const int c = (k * 10 + j) * 10 + i;
    return {s_sdf[c], s_cw[c]};
  }
  __device__ __forceinline__ VoxVal voxel_point(f3 p, int* res) const {
    return voxel_at(world_to_voxel(p, m.voxel_size), res);
  }
};

// trilinearInterpolation (voxel_data_structures.cu:260-338)
template <class S>
__device__ __forceinline__ bool trilinear(const S& s, f3 pos, float& dist) {
  const float vs   = s.voxel_size_point(pos);
  const float half = fmul(vs, 0.5f);
  const f3 dual    = {fsub(pos.x, half), fsub(pos.y, half), fsub(pos.z, half)};
  const int base_r = s.block_res_scaled(pos, vs);
  dist             = 0.f;
  const float pos_sdf = s.voxel_point(dual, nullptr).sdf;
  float x1 = dual.x, y1 = dual.y, z1 = dual.z;
  float sdf[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int dx = i & 1, dy = (i >> 1) & 1, dz = (i >> 2) & 1;
    const f3 vp  = {ffma(i2f(dx), vs, dual.x), ffma(i2f(dy), vs, dual.y), ffma(i2f(dz), vs, dual.z)};
    int res      = 0;
    const VoxVal v = s.voxel_point(vp, &res);
    if (!(v.cw >> 24))
      return false;
    if (res > base_r) {
      const float nvs = fmul(vs, 2.f);
      const float nh  = fmul(nvs, 0.5f);
      const f3 np     = {ffma(i2f(dx), nvs, fsub(pos.x, nh)), ffma(i2f(dy), nvs, fsub(pos.y, nh)), ffma(i2f(dz), nvs, fsub(pos.z, nh))};
      const float ns  = s.voxel_point(np, nullptr).sdf;
      sdf[i]          = ffma(0.5f, ns, fmul(0.5f, pos_sdf));
    } else {
      sdf[i] = v.sdf;
    }
    if (vp.x > x1)
      x1 = vp.x;
    if (vp.y > y1)
      y1 = vp.y;
    if (vp.z > z1)
      z1 = vp.z;
  }
  const float wx = fsub(x1, dual.x), wy = fsub(y1, dual.y), wz = fsub(z1, dual.z);
  const float dx = wx > 1e-6f ? fdiv(fsub(pos.x, dual.x), wx) : 0.5f;
  const float dy = wy > 1e-6f ? fdiv(fsub(pos.y, dual.y), wy) : 0.5f;
  const float dz = wz > 1e-6f ? fdiv(fsub(pos.z, dual.z), wz) : 0.5f;
  const float c0 = sdf[0];
  const float c1 = fsub(sdf[1], sdf[0]);
  const float c2 = fsub(sdf[2], sdf[0]);
  const float c3 = fsub(sdf[4], sdf[0]);
  const float c4 = fadd(fsub(fsub(sdf[3], sdf[2]), sdf[1]), sdf[0]);
  const float c5 = fadd(fsub(fsub(sdf[6], sdf[4]), sdf[2]), sdf[0]);
  const float c6 = fadd(fsub(fsub(sdf[5], sdf[4]), sdf[1]), sdf[0]);
  const float c7 = fsub(fadd(fadd(fadd(fsub(fsub(fsub(sdf[7], sdf[6]), sdf[5]), sdf[3]), sdf[1]), sdf[4]), sdf[2]), sdf[0]);
  float acc      = ffma(c1, dx, c0);
  acc            = ffma(c2, dy, acc);
  acc            = ffma(c3, dz, acc);
  acc            = ffma(fmul(c4, dx), dy, acc);
  acc            = ffma(fmul(c5, dy), dz, acc);
  acc            = ffma(fmul(c6, dx), dz, acc);
  acc            = ffma(fmul(fmul(c7, dx), dy), dz, acc);
  dist           = acc;
  return true;
}

// vertexInterp (mesh_extractor.cu:6-36), isolevel = 0
__device__ __forceinline__ TriVertex vertex_interp(f3 p1, f3 p2, float d1, float d2, uint32_t cw1, uint32_t cw2) {
  const float r1 = __uint2float_rn(cw1 & 0xFF), g1 = __uint2float_rn((cw1 >> 8) & 0xFF), b1 = __uint2float_rn((cw1 >> 16) & 0xFF);
  if (fabsf(fsub(0.f, d1)) < 0.00001f)
    return {p1, {fdiv(r1, 255.f), fdiv(g1, 255.f), fdiv(b1, 255.f)}};
  if (fabsf(fsub(0.f, d2)) < 0.00001f) {
    const float r2 = __uint2float_rn(cw2 & 0xFF), g2 = __uint2float_rn((cw2 >> 8) & 0xFF), b2 = __uint2float_rn((cw2 >> 16) & 0xFF);
    return {p2, {fdiv(r2, 255.f), fdiv(g2, 255.f), fdiv(b2, 255.f)}};
  }
  if (fabsf(fsub(d1, d2)) < 0.00001f)
    return {p1, {fdiv(r1, 255.f), fdiv(g1, 255.f), fdiv(b1, 255.f)}};
  const float mu = fdiv(fsub(0.f, d1), fsub(d2, d1));
  TriVertex r;
  r.p = {ffma(mu, fsub(p2.x, p1.x), p1.x), ffma(mu, fsub(p2.y, p1.y), p1.y), ffma(mu, fsub(p2.z, p1.z), p1.z)};
  // c1 + mu * (c2 - c1) / 255 with c1 un-normalised and (c2 - c1) an int difference (Q4)
  const int dr = (int) (cw2 & 0xFF) - (int) (cw1 & 0xFF), dg = (int) ((cw2 >> 8) & 0xFF) - (int) ((cw1 >> 8) & 0xFF), db = (int) ((cw2 >> 16) & 0xFF) - (int) ((cw1 >> 16) & 0xFF);
  r.c = {fadd(r1, fdiv(fmul(mu, i2f(dr)), 255.f)), fadd(g1, fdiv(fmul(mu, i2f(dg)), 255.f)), fadd(b1, fdiv(fmul(mu, i2f(db)), 255.f))};
  return r;
}

struct CellResult {
  int n_tri; // 0..5
  unsigned cube;
  f3 p[8];
  float d[8];
  uint32_t cw[8];
};

// extractIsoSurfaceAtPosition up to the table lookup (marching_cubes.cu:72-214). Corner order of
// p/d/cw: bit0 = +x, bit1 = +y, bit2 = +z (the reference's 000,001,010,011,100,101,110,111).
template <class S, bool CHECK_NEIGHBOUR_SIZES>
__device__ __forceinline__ void mc_cell(const S& s, const MapDev& m, f3 pf, CellResult& out) {
  out.n_tri       = 0;
  const float vvs = s.voxel_size_point(pf);
  const float P   = fmul(vvs, 0.5f);
  float sp[3]     = {P, P, P};
  float sm[3]     = {-P, -P, -P};
  if (CHECK_NEIGHBOUR_SIZES) {
    // checkVertexVoxels (:7-69): pf + make_float3(scaled, 0, 0) etc.
    const f3 q[6] = {{fadd(pf.x, sp[0]), fadd(pf.y, 0.f), fadd(pf.z, 0.f)}, {fadd(pf.x, sm[0]), fadd(pf.y, 0.f), fadd(pf.z, 0.f)},
                     {fadd(pf.x, 0.f), fadd(pf.y, sp[1]), fadd(pf.z, 0.f)}, {fadd(pf.x, 0.f), fadd(pf.y, sm[1]), fadd(pf.z, 0.f)},
                     {fadd(pf.x, 0.f), fadd(pf.y, 0.f), fadd(pf.z, sp[2])}, {fadd(pf.x, 0.f), fadd(pf.y, 0.f), fadd(pf.z, sm[2])}};
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const float vs = s.voxel_size_point(q[a]);
      if (vs > 0.f && vs < 1.f && vs != vvs) {
        if (a & 1)
          sm[a >> 1] = fmul(sm[a >> 1], 0.499f);
        else
          sp[a >> 1] = fmul(sp[a >> 1], 0.499f);
      }
    }
  }
  auto corner = [&](int k) -> bool {
    const f3 p = {fadd(pf.x, (k & 1) ? sp[0] : sm[0]), fadd(pf.y, (k & 2) ? sp[1] : sm[1]), fadd(pf.z, (k & 4) ? sp[2] : sm[2])};
    float dist;
    const bool valid = trilinear(s, p, dist);
    const VoxVal v   = s.voxel_point(p, nullptr);
    if (!valid) {
      if ((int) (v.cw >> 24) < m.min_weight_threshold)
        return false;
      dist = v.sdf;
    }
    out.p[k] = p, out.d[k] = dist, out.cw[k] = v.cw;
    return true;
  };
  if (S::kRollCorners) {
    // one copy of the corner evaluation instead of eight: unrolled, the shared-memory path is ~4 000
    // straight-line instructions and the few warps a block keeps after the pre-filter wait on
    // instruction fetch; the cell's arrays live in local memory either way (mc_emit indexes them)
#pragma unroll 1
    for (int k = 0; k < 8; ++k)
      if (!corner(k))
        return;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (!corner(k))
        return;
  }
  unsigned cube = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    cube |= (out.d[k] < 0.f) ? (1u << k) : 0u;
  const float thr = m.mc_threshold;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      if (fmul(out.d[k], out.d[l]) < 0.f) {
        if (fadd(fabsf(out.d[k]), fabsf(out.d[l])) > thr)
          return;
      } else if (fabsf(fsub(out.d[k], out.d[l])) > thr) {
        return;
      }
    }
    if (fabsf(out.d[k]) > thr)
      return;
  }
  out.cube  = cube;
  out.n_tri = k_mc_cell_counts[k_mc_cell_class[cube]] & 0x0F;
}

// edge vertex for the Transvoxel vertex code (low byte = the two corner numbers, :216-252)
__device__ __forceinline__ TriVertex mc_edge_vertex(const CellResult& c, unsigned code) {
  const int a = (code >> 4) & 0xF, b = code & 0xF;
  return vertex_interp(c.p[a], c.p[b], c.d[a], c.d[b], c.cw[a], c.cw[b]);
}

__device__ __forceinline__ void mc_emit(const CellResult& c, float* tri_out /* n_tri * 18 floats */) {
  const unsigned cls          = k_mc_cell_class[c.cube];
  const int n_vert            = k_mc_cell_counts[cls] >> 4;
  const unsigned short* codes = k_mc_vertex_data[c.cube];
  for (int t = 0; t < c.n_tri; ++t) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int vi = k_mc_cell_index[cls][3 * t + j];
      TriVertex v  = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
      if (vi < n_vert) {
        const unsigned code = codes[vi] & 0xFF;
        // the reference only recognises the 12 regular edge codes; anything else leaves Vertex() zeros
        v = mc_edge_vertex(c, code);
      }
      float* o = tri_out + (size_t) t * 18 + j * 6;
      o[0] = v.p.x, o[1] = v.p.y, o[2] = v.p.z, o[3] = v.c.x, o[4] = v.c.y, o[5] = v.c.z;
    }
  }
}

// Exact pre-filter of the halo path. Every value extractIsoSurfaceAtPosition can read for voxel
// (x, y, z) of a resolution-0 block with resolution-0 neighbours lies in the voxel's 3 x 3 x 3
// neighbourhood: the eight corners sit half a voxel away, and each corner's trilinear sample reads the
// eight voxels around it with weights in [0, 1] (coordinates below 2^19 voxels keep positions exact to
// vs / 16). A corner value is therefore either a convex combination of eight OBSERVED (weight > 0)
// neighbourhood sdf values - evaluated with at most ~3e-5 * max|sdf| of rounding error - or, for a
// corner with an unobserved voxel among its eight, the value of the voxel under the corner, which
// must itself carry weight >= min_weight_threshold > 0 or the cell is abandoned. Hence, over the
// observed voxels of the neighbourhood:
//  * all sdf values > 1e-4 * max|sdf|  -> every corner is positive, cube index 0, no triangle;
//  * all sdf values < -1e-4 * max|sdf| -> every corner is negative, cube index 255, no triangle;
//  * no observed voxel at all -> the first corner falls back to an unobserved voxel and the cell is
//    abandoned.
// In each case the full evaluation would emit nothing, so skipping it changes no output bit. Most
// voxels of a TSDF band are on one side of the surface with their whole neighbourhood.
//
// Second, tighter test when all 27 voxels are observed. Every corner is then valid, and its value is the
// trilinear blend of its 2 x 2 x 2 voxels with weights (1/2 +- a)(1/2 +- b)(1/2 +- c): the blend
// fraction is RN((pos - dual) / w) with pos - dual = vs / 2 up to the rounding of `dual` and w = vs up
// to the rounding of `dual + vs` (both subtractions are exact, Sterbenz), so |a|, |b|, |c| <= delta =
// 2.1 k u, k = the largest voxel coordinate involved, u = 2^-24. Hence |corner - mean of the 8 voxels|
// <= ((1 + 2 delta)^3 - 1) max|sdf| <= 6.2 delta max|sdf| (+ 3e-5 max|sdf| of evaluation rounding).
// `slack` = 6.2 * (twice that delta) + 1e-4, computed once per block from its coordinates: if the eight
// means all exceed slack * max|sdf|, or are all below its negative, the eight corners share that sign
// and the cell has no triangle. This removes the cells one to two voxels away from the surface, which
// pass the neighbourhood sign test but cannot be crossed.
__device__ __forceinline__ bool halo_cell_is_empty(const float* s_sdf, const uint32_t* s_cw, int vi, float slack) {
  const int x = vi & 7, y = (vi >> 3) & 7, z = vi >> 6; // halo cell (x + 1, y + 1, z + 1)
  float lo = 3.4e38f, hi = -3.4e38f;
  uint32_t w = 0;
  bool finite = true, all_seen = true;
  float v[27];
#pragma unroll
  for (int dz = 0; dz < 3; ++dz)
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int c       = ((z + dz) * 10 + (y + dy)) * 10 + (x + dx);
        const uint32_t wc = s_cw[c] >> 24;
        const float sv    = s_sdf[c];
        v[(dz * 3 + dy) * 3 + dx] = sv;
        if (wc) { // unobserved voxels never supply a value (see above)
          lo = fminf(lo, sv), hi = fmaxf(hi, sv);
          finite &= fabsf(sv) < 1e30f; // false for NaN too (fminf / fmaxf would drop it silently)
        } else {
          all_seen = false;
        }
        w |= wc;
      }
  if (!finite)
    return false; // NaN / Inf / absurd payloads: let the full path decide
  const float big    = fmaxf(fabsf(lo), fabsf(hi));
  const float margin = 1e-4f * big;
  if (w == 0u || lo > margin || hi < -margin)
    return true;
  if (!MRH_MC_TIGHT || !all_seen)
    return false;
  float cmin = 3.4e38f, cmax = -3.4e38f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ox = k & 1, oy = (k >> 1) & 1, oz = k >> 2;
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sum += v[((oz + (j >> 2)) * 3 + oy + ((j >> 1) & 1)) * 3 + ox + (j & 1)];
    const float mean = 0.125f * sum;
    cmin = fminf(cmin, mean), cmax = fmaxf(cmax, mean);
  }
  const float m2 = slack * big;
  return cmin > m2 || cmax < -m2;
}

// One CTA of kMcThreads per live block. Small CTAs: per block the kernel is a chain of dependent
// latencies (neighbour lookups -> halo fill -> filter -> cells -> append), so the SM is kept busy by
// having many blocks in flight, not by wide CTAs idling at the barriers of one block.
#ifndef MRH_MC_THREADS
#define MRH_MC_THREADS 128
#endif
constexpr int kMcThreads = MRH_MC_THREADS;

// force_generic: run every block through the hash sampler (validation of the halo path).
// max_centers: only the first max_centers entries of the live list are meshed; the entries behind
// them are ghost copies of other ranks' blocks (mrh_halo.cu), present only to be sampled.
__global__ void __launch_bounds__(kMcThreads, 1024 / kMcThreads) k_mc_blocks(MapDev m, uint32_t live_cur, float* __restrict__ triangles, uint32_t* __restrict__ tri_count, uint32_t max_triangles, int force_generic, uint32_t max_centers) {
  __shared__ float s_sdf[1000];
  __shared__ uint32_t s_cw[1000];
  __shared__ uint32_t s_nb[27];
  __shared__ int s_mixed;
  __shared__ uint32_t s_warp_sum[kMcThreads / 32];
  __shared__ uint32_t s_base;
  __shared__ uint32_t s_n_cand;
  __shared__ uint16_t s_cand[512];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_live = min(m.ctr->live_count[live_cur], max_centers);
  for (uint32_t li = blockIdx.x; li < n_live; li += gridDim.x) {
    const LiveEntry le = m.live[live_cur][li];
    if (le.slot == kInvalid)
      continue;
    const unsigned long long key = le.key;
    const uint32_t val           = le.val;
    const i3 b         = unpack_key(key);
    const int res      = (int) (val >> 31);
    if (tid == 0)
      s_mixed = res | force_generic, s_n_cand = 0;
    __syncthreads();
    if (tid < 27) {
      const i3 nb = {b.x + tid % 3 - 1, b.y + (tid / 3) % 3 - 1, b.z + tid / 9 - 1};
      uint32_t v  = kInvalid;
      if (tid == 13) {
        v = val;
      } else {
        const int sl = table_find(m, nb);
        if (sl >= 0)
          v = m.vals[sl];
      }
      if (v != kInvalid && (v >> 31))
        s_mixed = 1;
      s_nb[tid] = v;
    }
    __syncthreads();
    const bool generic = s_mixed != 0;
    // the pre-filter's bounds assume voxel coordinates far below 2^23 (positions exact to vs / 16)
    const bool prefilter = m.min_weight_threshold > 0 && max(max(abs(b.x), abs(b.y)), abs(b.z)) < (1 << 16);
    // blend-fraction bound of the tighter test: delta = 2.1 k u, doubled, k = largest voxel coordinate around the block
    const float slack = 6.2f * (4.2f * (float) (8 * (max(max(abs(b.x), abs(b.y)), abs(b.z)) + 2)) * 5.9604645e-8f) + 1e-4f;
    if (!generic) {
      for (int c = tid; c < 1000; c += kMcThreads) {
        const int i = c % 10, j = (c / 10) % 10, k = c / 100;
        const int bi = (i + 7) >> 3, bj = (j + 7) >> 3, bk = (k + 7) >> 3;
        const uint32_t v = s_nb[(bk * 3 + bj) * 3 + bi];
        float sdf   = 0.f;
        uint32_t cw = 0u;
        if (v != kInvalid) {
          const int idx       = (((k + 7) & 7) * 8 + ((j + 7) & 7)) * 8 + ((i + 7) & 7);
          const uint8_t* base = m.pool + (size_t) v * kBlockBytes;
          sdf                 = reinterpret_cast<const float*>(base)[idx];
          cw                  = reinterpret_cast<const uint32_t*>(base + 2 * kPlaneBytes)[idx];
        }
        s_sdf[c] = sdf, s_cw[c] = cw;
      }
    }
    __syncthreads();
    // Candidate cells of the block, compacted: the pre-filter drops most voxels of a TSDF band, and
    // the survivors are packed so that whole warps take the full evaluation (a warp with one
    // survivor among its 32 consecutive voxels would otherwise pay for all of it).
    int n_vox = res ? 64 : 512;
    if (!generic && prefilter) {
      for (int vi = tid; vi < 512; vi += kMcThreads) {
        const bool keep      = !halo_cell_is_empty(s_sdf, s_cw, vi, slack);
        const unsigned votes = __ballot_sync(0xFFFFFFFFu, keep);
        uint32_t base        = 0;
        if (lane == 0 && votes)
          base = atomicAdd(&s_n_cand, (uint32_t) __popc(votes));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (keep)
          s_cand[base + __popc(votes & ((1u << lane) - 1u))] = (uint16_t) vi;
      }
      __syncthreads();
      n_vox = (int) s_n_cand;
    }
    for (int first = 0; first < n_vox; first += kMcThreads) {
      CellResult cell;
      cell.n_tri = 0;
      if (first + tid < n_vox) {
        const int vi = (!generic && prefilter) ? (int) s_cand[first + tid] : first + tid;
        const int sf = 1 << res, bs = 8 >> res;
        const i3 pi  = {b.x * 8 + sf * (vi % bs), b.y * 8 + sf * ((vi % (bs * bs)) / bs), b.z * 8 + sf * (vi / (bs * bs))};
        const f3 pf  = {fmul(i2f(pi.x), m.voxel_size), fmul(i2f(pi.y), m.voxel_size), fmul(i2f(pi.z), m.voxel_size)};
        if (generic) {
          const HashSampler s{m};
          mc_cell<HashSampler, true>(s, m, pf, cell);
        } else {
          const HaloSampler s{m, s_sdf, s_cw, {b.x * 8 - 1, b.y * 8 - 1, b.z * 8 - 1}};
          mc_cell<HaloSampler, false>(s, m, pf, cell);
        }
      }
      // CTA-wide exclusive prefix sum of the triangle counts, one atomicAdd for the whole pass
      uint32_t incl = (uint32_t) cell.n_tri;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o)
          incl += t;
      }
      if (lane == 31)
        s_warp_sum[warp] = incl;
      __syncthreads();
      if (tid == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < kMcThreads / 32; ++w) {
          const uint32_t t = s_warp_sum[w];
          s_warp_sum[w]    = tot;
          tot += t;
        }
        s_base = tot ? atomicAdd(tri_count, tot) : 0u;
      }
      __syncthreads();
      if (cell.n_tri) {
        const uint32_t first = s_base + s_warp_sum[warp] + incl - (uint32_t) cell.n_tri;
        // appendTriangle (:43-53): triangles past the capacity are counted but not stored
        if (first + (uint32_t) cell.n_tri <= max_triangles)
          mc_emit(cell, triangles + (size_t) first * 18);
        else if (first < max_triangles) {
          CellResult part = cell;
          part.n_tri      = (int) (max_triangles - first);
          mc_emit(part, triangles + (size_t) first * 18);
        }
      }
      __syncthreads();
    }
  }
}

} // namespace mrh

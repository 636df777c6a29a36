// mrh_math.cuh — device numerics for the TSDF hot path.
//
// Parity with the reference is decided by float rounding at a handful of discrete tests (block
// index of a ray end point, DDA step order, pixel rounding, sdf <= -t, GC threshold), so every
// expression that feeds one of them is written with explicit round-to-nearest intrinsics in the
// exact operation order the reference's own build produces (nvcc -O3, default -fmad=true; the
// sequence was read from the SASS of the reference kernels compiled for sm_100a). Intrinsics are
// never re-contracted by the compiler, so this file pins the arithmetic independently of flags.
//
// Contraction pattern of the reference build (reference file:line in /root/reference/mrhash/src/sdf):
//   a*x + b*y + c*z    -> fma(c, z, fma(a, x, b*y))     cuda_algebra.cuh:71-75, cuda_math.cuh dot()
//   trunc + scale*z    -> fma(scale, z, trunc)          voxel_hash_utils.cuh:184-187
//   k*size - 0.5*size  -> fma(k, size, -(0.5*size))     voxel_data_structures.cu:797
#pragma once
#include <climits>
#include <cstdint>
#include <cuda_runtime.h>

namespace mrh {

constexpr int kBlockSide   = 8;    // params.h:10 sdf_block_size
constexpr int kBlockVoxels = 512;  // params.h:11
constexpr int kWeightMax   = 255;  // params.h:25
constexpr int kMaxDDA      = 1024; // params.h:26

struct CameraDev {
  float fx, fy, ifx, ify, cx, cy;
  uint32_t rows, cols;
  int row_thr, col_thr;
  float min_depth, max_depth;
  int model; // 0 pinhole, 1 spherical (camera.cuh:10)
};

// cam_in_world (R,t) and its inverse exactly as CUDAMatSE3::inverse() builds it (cuda_algebra.cuh:137-143)
struct PoseDev {
  float R[9];
  float t[3];
  float Ri[9];
  float ti[3];
};

struct f3 {
  float x, y, z;
};
struct i3 {
  int x, y, z;
};

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ int f2i(float a) { return __float2int_rz(a); }
__device__ __forceinline__ float i2f(int a) { return __int2float_rn(a); }

// cuda_math.cuh:62-64
__device__ __forceinline__ int sign_i(float v) { return (0.f < v) - (v < 0.f); }

// row-times-vector with the reference's contraction order
__device__ __forceinline__ float row_dot(float a, float b, float c, f3 p) {
  return ffma(c, p.z, ffma(a, p.x, fmul(b, p.y)));
}
__device__ __forceinline__ f3 mat3_mul(const float* R, f3 p) {
  return {row_dot(R[0], R[1], R[2], p), row_dot(R[3], R[4], R[5], p), row_dot(R[6], R[7], R[8], p)};
}
// CUDAMatSE3::operator* (cuda_algebra.cuh:146-148)
__device__ __forceinline__ f3 se3_mul(const float* R, const float* t, f3 p) {
  f3 r = mat3_mul(R, p);
  return {fadd(r.x, t[0]), fadd(r.y, t[1]), fadd(r.z, t[2])};
}
__device__ __forceinline__ float dot3(f3 a, f3 b) { return ffma(a.z, b.z, ffma(a.x, b.x, fmul(a.y, b.y))); }
// cuda_math.cuh:1075-1078 — rsqrtf is MUFU.RSQ + denormal wrapper on both sides
__device__ __forceinline__ f3 normalize3(f3 v) {
  const float inv = rsqrtf(dot3(v, v));
  return {fmul(v.x, inv), fmul(v.y, inv), fmul(v.z, inv)};
}

__device__ __forceinline__ void pose_finish(PoseDev& p) {
  // inverse(): rotation^T, translation = -(R^T t)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      p.Ri[j * 3 + i] = p.R[i * 3 + j];
  f3 t  = {p.t[0], p.t[1], p.t[2]};
  f3 rt = mat3_mul(p.Ri, t);
  p.ti[0] = -rt.x, p.ti[1] = -rt.y, p.ti[2] = -rt.z;
}

// getTruncation on the device (voxel_hash_utils.cuh:184-187)
__device__ __forceinline__ float truncation(float trunc, float scale, float z) { return ffma(scale, z, trunc); }

// voxel_hash_utils.cuh:143-151 (one axis)
__device__ __forceinline__ int world_to_voxel_1(float p, float size) {
  const float q = fdiv(p, size);
  float a       = fadd(q, fmul(i2f(sign_i(q)), 0.5f));
  a             = (a >= 0.f) ? floorf(fadd(a, 1e-5f)) : ceilf(fsub(a, 1e-5f));
  return f2i(a);
}
// voxel_hash_utils.cuh:75-103 (one axis)
__device__ __forceinline__ int voxel_to_block_1(int v, float size, float ext) {
  if (v < 0)
    v -= (kBlockSide - 1);
  const float pw  = fmul(i2f(v), size);
  const float mbs = fmul(fmul(ext, 8.f), size);
  const float b   = (pw >= 0.f) ? floorf(fdiv(fadd(pw, 1e-5f), mbs)) : ceilf(fdiv(fsub(pw, 1e-5f), mbs));
  return f2i(b);
}
__device__ __forceinline__ i3 world_to_voxel(f3 p, float size) {
  return {world_to_voxel_1(p.x, size), world_to_voxel_1(p.y, size), world_to_voxel_1(p.z, size)};
}
__device__ __forceinline__ i3 voxel_to_block(i3 v, float size, const float* ext) {
  return {voxel_to_block_1(v.x, size, ext[0]), voxel_to_block_1(v.y, size, ext[1]), voxel_to_block_1(v.z, size, ext[2])};
}
__device__ __forceinline__ i3 world_to_block(f3 p, float size, const float* ext) {
  return voxel_to_block(world_to_voxel(p, size), size, ext);
}

// camera.cuh:84-103
__device__ __forceinline__ f3 inverse_projection(const CameraDev& c, uint32_t row, uint32_t col, float d) {
  const float u = fmul(c.ifx, fsub(fsub(__uint2float_rn(col), c.cx), 0.5f));
  const float v = fmul(c.ify, fsub(fsub(__uint2float_rn(row), c.cy), 0.5f));
  if (c.model == 0)
    return {fmul(d, u), fmul(d, v), fmul(d, 1.f)};
  const float s0 = sinf(u), c0 = cosf(u), s1 = sinf(v), c1 = cosf(v);
  return {fmul(d, fmul(c0, c1)), fmul(d, fmul(s0, c1)), fmul(d, s1)};
}
// camera.cuh:120-129
__device__ __forceinline__ float get_depth(const CameraDev& c, f3 p) {
  if (c.model == 0)
    return p.z;
  return __fsqrt_rn(ffma(p.z, p.z, ffma(p.x, p.x, fmul(p.y, p.y))));
}
// the depth the reference reads back from its float3 cloud image (camera.cu:5-19 then getDepth).
// A NaN depth passes both comparisons, exactly as in calculateCloudKernel: such a pixel allocates
// nothing (fminf drops the NaN bounds, dmin == dmax) but poisons the voxels that project onto it.
__device__ __forceinline__ float cloud_depth(const CameraDev& c, uint32_t row, uint32_t col, float raw) {
  if (raw <= c.min_depth || raw > c.max_depth)
    return 0.f;
  if (c.model == 0)
    return raw;
  return get_depth(c, inverse_projection(c, row, col, raw));
}
// common part of projectPoint / projectPointApprox (camera.cuh:131-203)
__device__ __forceinline__ bool project_rc(const CameraDev& c, f3 pc, int& row, int& col) {
  if (c.model == 0) {
    if (pc.z <= c.min_depth || !(pc.z <= c.max_depth))
      return false;
    row = f2i(fadd(fadd(fdiv(fmul(c.fy, pc.y), pc.z), c.cy), 0.5f));
    col = f2i(fadd(fadd(fdiv(fmul(c.fx, pc.x), pc.z), c.cx), 0.5f));
    return true;
  }
  const float range = __fsqrt_rn(ffma(pc.z, pc.z, ffma(pc.x, pc.x, fmul(pc.y, pc.y))));
  if (range < c.min_depth || range > c.max_depth)
    return false;
  const float px = atan2f(pc.y, pc.x);
  const float py = asinf(fdiv(pc.z, range));
  row            = f2i(fadd(ffma(c.fy, py, c.cy), 0.5f));
  col            = f2i(fadd(ffma(c.fx, px, c.cx), 0.5f));
  return true;
}
__device__ __forceinline__ bool project_point(const CameraDev& c, f3 pc, int& row, int& col) {
  int r, q;
  if (!project_rc(c, pc, r, q))
    return false;
  if (r >= 0 && q >= 0 && (uint32_t) r < c.rows && (uint32_t) q < c.cols) {
    row = r, col = q;
    return true;
  }
  return false;
}
__device__ __forceinline__ bool project_point_approx(const CameraDev& c, f3 pc) {
  int r, q;
  if (!project_rc(c, pc, r, q))
    return false;
  return r >= -c.row_thr && q >= -c.col_thr && r < (int) (c.rows + (uint32_t) c.row_thr) && q < (int) (c.cols + (uint32_t) c.col_thr);
}

// one corner of isSDFBlockInCameraFrustumApprox (voxel_data_structures.cu:66-77, params.h:41-49)
__device__ __forceinline__ bool block_corner_in_frustum(const CameraDev& c, const PoseDev& pose, i3 b, int corner, float size) {
  const int ox = (corner & 4) ? 7 : 0, oy = (corner & 2) ? 7 : 0, oz = (corner & 1) ? 7 : 0;
  const f3 w   = {fmul(i2f(b.x * kBlockSide + ox), size), fmul(i2f(b.y * kBlockSide + oy), size), fmul(i2f(b.z * kBlockSide + oz), size)};
  return project_point_approx(c, se3_mul(pose.Ri, pose.ti, w));
}
__device__ __forceinline__ bool block_in_frustum(const CameraDev& c, const PoseDev& pose, i3 b, float size) {
#pragma unroll 1
  for (int k = 0; k < 8; ++k)
    if (block_corner_in_frustum(c, pose, b, k, size))
      return true;
  return false;
}

// Frustum test of a block plus a conservative "can any voxel of it land inside the image?" answer.
// For the pinhole model the projection maps the box spanned by the block's voxel centres (all in
// front of the camera) onto the convex hull of its 8 projected corners, so if every corner falls
// at least 3 pixels outside one image edge no voxel can pass projectPoint (3 px absorbs float
// rounding and the int(x + 0.5) truncation toward zero). Any corner with an invalid depth, or the
// spherical model, answers "maybe". Used to skip the per-voxel pass of such blocks.
__device__ __forceinline__ bool block_in_frustum_ex(const CameraDev& c, const PoseDev& pose, i3 b, float size, bool& maybe_in_image) {
  bool in_frustum = false, all_valid = c.model == 0;
  int rmin = INT_MAX, rmax = INT_MIN, qmin = INT_MAX, qmax = INT_MIN;
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    const int ox = (k & 4) ? 7 : 0, oy = (k & 2) ? 7 : 0, oz = (k & 1) ? 7 : 0;
    const f3 w   = {fmul(i2f(b.x * kBlockSide + ox), size), fmul(i2f(b.y * kBlockSide + oy), size), fmul(i2f(b.z * kBlockSide + oz), size)};
    int r, q;
    if (project_rc(c, se3_mul(pose.Ri, pose.ti, w), r, q)) {
      in_frustum |= r >= -c.row_thr && q >= -c.col_thr && r < (int) (c.rows + (uint32_t) c.row_thr) && q < (int) (c.cols + (uint32_t) c.col_thr);
      rmin = min(rmin, r), rmax = max(rmax, r), qmin = min(qmin, q), qmax = max(qmax, q);
    } else {
      all_valid = false;
    }
  }
  maybe_in_image = !all_valid || !(rmax <= -3 || qmax <= -3 || rmin >= (int) c.rows + 3 || qmin >= (int) c.cols + 3);
  return in_frustum;
}

// calculateHash (voxel_data_structures.cu:151-160)
__device__ __forceinline__ uint32_t block_hash(i3 b, uint32_t num_buckets) {
  return (((uint32_t) b.x * 73856093u) ^ ((uint32_t) b.y * 19349669u) ^ ((uint32_t) b.z * 83492791u)) % num_buckets;
}

// Amanatides-Woo DDA state shared by the block walk (allocBlocksKernel :782-857,
// allocBlocks3DKernel :963-1033) and the voxel walk (integrate3DKernel :1259-1378).
struct DDA {
  i3 cur, bound, istep;
  f3 t_max, t_delta;
  __device__ __forceinline__ void init(f3 p0, f3 p1, float size, const float* ext, bool block_level) {
    const f3 dir = normalize3({fsub(p1.x, p0.x), fsub(p1.y, p0.y), fsub(p1.z, p0.z)});
    i3 end;
    if (block_level) {
      cur = world_to_block(p0, size, ext);
      end = world_to_block(p1, size, ext);
    } else {
      cur = world_to_voxel(p0, size);
      end = world_to_voxel(p1, size);
    }
    // The reference keeps the step as a float3 and advances with int(float(id) + step)
    // (voxel_data_structures.cu:803,826-854): for |id| < 2^24 that is exactly id + sign, so the step
    // lives here as an integer (coordinates beyond 2^24 cells are outside the key range anyway).
    istep           = {sign_i(dir.x), sign_i(dir.y), sign_i(dir.z)};
    const f3 step   = {i2f(istep.x), i2f(istep.y), i2f(istep.z)};
    const float nh  = -fmul(0.5f, size);
    const int scale = block_level ? kBlockSide : 1;
    const float cell = block_level ? fmul(8.f, size) : size; // (step*8)*size == step*(8*size): step is 0/+-1
    const float bx = ffma(i2f((cur.x + max(istep.x, 0)) * scale), size, nh); // clamp(step, 0, 1)
    const float by = ffma(i2f((cur.y + max(istep.y, 0)) * scale), size, nh);
    const float bz = ffma(i2f((cur.z + max(istep.z, 0)) * scale), size, nh);
    t_max   = {fdiv(fsub(bx, p0.x), dir.x), fdiv(fsub(by, p0.y), dir.y), fdiv(fsub(bz, p0.z), dir.z)};
    t_delta = {fdiv(fmul(step.x, cell), dir.x), fdiv(fmul(step.y, cell), dir.y), fdiv(fmul(step.z, cell), dir.z)};
    bound   = {end.x + istep.x, end.y + istep.y, end.z + istep.z};
    const float big = 3.40282346638528859812e+38f;
    if (fabsf(dir.x) < 1e-6f || fabsf(fsub(bx, dir.x)) < 1e-6f)
      t_max.x = big, t_delta.x = big;
    if (fabsf(dir.y) < 1e-6f || fabsf(fsub(by, dir.y)) < 1e-6f)
      t_max.y = big, t_delta.y = big;
    if (fabsf(dir.z) < 1e-6f || fabsf(fsub(bz, dir.z)) < 1e-6f)
      t_max.z = big, t_delta.z = big;
  }
  // returns false when the walk leaves the segment (the reference `return`s)
  __device__ __forceinline__ bool advance() {
    if (t_max.x < t_max.y && t_max.x < t_max.z) {
      cur.x += istep.x;
      if (cur.x == bound.x)
        return false;
      t_max.x = fadd(t_max.x, t_delta.x);
    } else if (t_max.z < t_max.y) {
      cur.z += istep.z;
      if (cur.z == bound.z)
        return false;
      t_max.z = fadd(t_max.z, t_delta.z);
    } else {
      cur.y += istep.y;
      if (cur.y == bound.y)
        return false;
      t_max.y = fadd(t_max.y, t_delta.y);
    }
    return true;
  }
};

} // namespace mrh

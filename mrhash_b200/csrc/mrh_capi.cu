// mrh_capi.cu — host orchestration behind the C ABI of include/mrhash_b200.h.
//
// Replaces the host side of GeoWrapper::compute() (geowrapper.cpp:118-148) and
// VoxelContainer::integrate() (voxel_data_structures.cpp:90-134): one stream, no host
// synchronisation inside a frame, no per-frame allocation, pinned double-buffered ingest.
#include <algorithm>
#include <thread>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/mrhash_b200.h"
#include "mrh_host.h"

using namespace mrh;

namespace mrh {
  thread_local std::string g_last_error;

  int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return 1;
  }
} // namespace mrh

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_));            \
  } while (0)

#define GUARD(m)                                                                                   \
  if (!(m))                                                                                        \
    return fail("null handle");                                                                    \
  CK(cudaSetDevice((m)->device))

// ---------------------------------------------------------------------------------------------
// initialisation kernels (resetBuffers, voxel_data_structures.cpp:58-87, done on the device)
// ---------------------------------------------------------------------------------------------
namespace {
  __global__ void k_init_heap(uint32_t* heap, BlockStats* stats, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      heap[i]  = n - 1u - i;
      stats[i] = {FLT_MAX, 0u};
    }
  }
  __global__ void k_init_counters(Counters* c, uint32_t n) {
    Counters z;
    memset(&z, 0, sizeof(z));
    z.heap_counter     = (int) n - 1;
    z.heap_low_counter = -1;
    *c                 = z;
  }
  // The counters of the frame just finished, written straight into page-locked host memory (mapped):
  // no copy-engine operation sits between two frames. A 144-byte cudaMemcpyAsync in the compute stream
  // queued behind the 1 MB uploads of the next frame on the DMA engines and held the next kernel back
  // by up to 30 us (tools/debug_timeline.py).
  __global__ void k_snapshot_counters(const Counters* __restrict__ src, Counters* __restrict__ dst_host) {
    constexpr int kWords = offsetof(Counters, dbg) / 4;
    const uint32_t* s    = reinterpret_cast<const uint32_t*>(src);
    volatile uint32_t* d = reinterpret_cast<volatile uint32_t*>(dst_host);
    for (int i = threadIdx.x; i < kWords; i += blockDim.x)
      d[i] = s[i];
    __threadfence_system();
  }
} // namespace

// ---------------------------------------------------------------------------------------------
static int alloc_map(mrh_map* m) {
  MapDev& d        = m->dev;
  const uint64_t N = m->num_sdf_blocks, NB = m->hash_num_buckets;
  if (N == 0 || NB == 0)
    return fail("num_sdf_blocks / hash_num_buckets must be > 0");
  if (N >= (1ull << 31) || NB * kBucketSlots >= (1ull << 32))
    return fail("sizing exceeds 32-bit slot / pool indices (num_sdf_blocks=%llu, hash_num_buckets=%llu)", (unsigned long long) N, (unsigned long long) NB);
  d.num_blocks  = (uint32_t) N;
  d.num_buckets = (uint32_t) NB;
  d.capacity    = (uint32_t) (NB * kBucketSlots);
  d.bucket_magic = (uint32_t) std::min<uint64_t>((1ull << 32) / NB, 0xFFFFFFFFull);
  CK(cudaMalloc(&d.keys, sizeof(unsigned long long) * d.capacity));
  CK(cudaMalloc(&d.vals, sizeof(uint32_t) * d.capacity));
  CK(cudaMalloc(&d.heap, sizeof(uint32_t) * N));
  CK(cudaMalloc(&d.heap_low, sizeof(uint32_t) * N * 8));
  CK(cudaMalloc(&d.pool, (size_t) kBlockBytes * N));
  CK(cudaMalloc(&d.carved, N));
  CK(cudaMalloc(&d.stats, sizeof(BlockStats) * N));
  CK(cudaMalloc(&d.live[0], sizeof(LiveEntry) * N * 2));
  CK(cudaMalloc(&d.live[1], sizeof(LiveEntry) * N * 2));
  CK(cudaMalloc(&d.vis, sizeof(VisEntry) * N * 2));
  // room for every block twice over (visible + new) plus one terminator per resident CTA
  CK(cudaMalloc(&d.fq, sizeof(FuseEntry) * (N * 2 + 8192)));
  CK(cudaMemset(d.fq, 0, sizeof(FuseEntry) * (N * 2 + 8192))); // tag 0 is never used by a frame
  CK(cudaMalloc(&d.gc_list, sizeof(GcEntry) * N * 2));
  CK(cudaMalloc(&d.fqs, sizeof(FrameQueues)));
  CK(cudaMemset(d.fqs, 0, sizeof(FrameQueues)));
  CK(cudaMalloc(&d.realloc_list, sizeof(VisEntry) * N));
  CK(cudaMalloc(&d.reint_keys, sizeof(unsigned long long) * N));
  CK(cudaMalloc(&d.ctr, sizeof(Counters)));
  return 0;
}

int mrh::reset_map(mrh_map* m) {
  MapDev& d = m->dev;
  CK(cudaMemsetAsync(d.keys, 0xFF, sizeof(unsigned long long) * d.capacity, m->stream));
  CK(cudaMemsetAsync(d.vals, 0xFF, sizeof(uint32_t) * d.capacity, m->stream));
  CK(cudaMemsetAsync(d.pool, 0, (size_t) kBlockBytes * d.num_blocks, m->stream));
  CK(cudaMemsetAsync(d.carved, 0, d.num_blocks, m->stream));
  k_init_heap<<<592, 256, 0, m->stream>>>(d.heap, d.stats, d.num_blocks);
  k_init_counters<<<1, 1, 0, m->stream>>>(d.ctr, d.num_blocks);
  CK(cudaMemsetAsync(d.fqs, 0, sizeof(FrameQueues), m->stream));
  m->launches += 2;
  m->live_cur       = 0;
  m->counters_clean = true;
  CK(cudaGetLastError());
  return 0;
}

static void free_map(mrh_map* m) {
  MapDev& d = m->dev;
  cudaFree(d.keys), cudaFree(d.vals), cudaFree(d.heap), cudaFree(d.heap_low), cudaFree(d.pool), cudaFree(d.carved), cudaFree(d.stats);
  cudaFree(d.live[0]), cudaFree(d.live[1]), cudaFree(d.vis), cudaFree(d.fq), cudaFree(d.gc_list), cudaFree(d.fqs), cudaFree(d.realloc_list), cudaFree(d.reint_keys), cudaFree(d.ctr), cudaFree(d.zbuf);
  for (Ingest* in : {&m->in_depth, &m->in_rgb, &m->in_points, &m->in_normals})
    for (int i = 0; i < 2; ++i) {
      cudaFree(in->d_buf[i]), cudaFreeHost(in->h_buf[i]);
      if (in->copied[i])
        cudaEventDestroy(in->copied[i]);
      if (in->consumed[i])
        cudaEventDestroy(in->consumed[i]);
    }
  if (m->copy_stream)
    cudaStreamDestroy(m->copy_stream);
  if (m->copy_stream2)
    cudaStreamDestroy(m->copy_stream2);
  cudaFree(m->d_tri), cudaFree(m->d_tri_count), cudaFree(m->d_soup_acc), cudaFree(m->d_shell_idx);
  cudaFree(m->d_upd_keys[0]), cudaFree(m->d_upd_keys[1]), cudaFree(m->d_upd_vals[0]), cudaFree(m->d_upd_vals[1]), cudaFree(m->d_sort_tmp);
  cudaFreeHost(m->h_ctr);
  cudaFreeHost(m->h_ctr_ring);
  for (int i = 0; i < mrh_map::kCtrRing; ++i)
    if (m->ev_ctr[i])
      cudaEventDestroy(m->ev_ctr[i]);
  cudaFreeHost(m->h_heap_probe);
  for (int i = 0; i < 2; ++i) {
    cudaFreeHost(m->h_bounce[i]);
    if (m->ev_bounce[i])
      cudaEventDestroy(m->ev_bounce[i]);
  }
  for (int i = 0; i < 8; ++i)
    if (m->ev_k[i])
      cudaEventDestroy(m->ev_k[i]);
  if (m->ev0)
    cudaEventDestroy(m->ev0);
  if (m->ev1)
    cudaEventDestroy(m->ev1);
  if (m->stream)
    cudaStreamDestroy(m->stream);
}

static void refresh_map_params(mrh_map* m) {
  MapDev& d                  = m->dev;
  const mrh_params& p        = m->p;
  d.voxel_size               = p.virtual_voxel_size;
  d.trunc                    = p.sdf_truncation;
  d.trunc_scale              = p.sdf_truncation_scale;
  d.max_integration_distance = m->max_integration_distance;
  d.ext[0] = d.ext[1] = d.ext[2] = (float) p.voxel_extents_scale;
  // host getTruncation(camera.maxDepth()) (voxel_data_structures.cu:1720): g++ host code, two roundings
  volatile float prod = p.sdf_truncation_scale * m->cam.max_depth;
  d.gc_threshold      = p.sdf_truncation + prod;
  d.var_threshold     = p.sdf_var_threshold;
  d.weight_sample     = p.integration_weight_sample;
  d.min_weight_threshold = p.min_weight_threshold;
  d.projective           = p.projective_sdf;
  d.mc_threshold         = p.marching_cubes_threshold;
  // shared-reciprocal divisions (mrh_div.cuh): every shared divisor must lie in the exponent window
  const float lo = 9.094947017729282e-13f, hi = 1.099511627776e12f; // 2^-40, 2^40
  const float vs = p.virtual_voxel_size;
  d.fast_div     = vs * 0.5f >= lo && vs * 8.f <= hi && m->cam.min_depth >= lo && m->cam.max_depth <= hi && p.integration_weight_sample >= 1 && p.integration_weight_sample <= (1 << 20) &&
                   !getenv("MRH_NO_FAST_DIV");
  d.block_shortcut_radius = 0;
  if (d.fast_div && m->shortcut_size == vs && m->shortcut_ext == d.ext[0])
    d.block_shortcut_radius = m->shortcut_radius;
  if (p.shard_world > 1) {
    d.shard_lo = (uint32_t) ((uint64_t) d.num_buckets * (uint64_t) p.shard_rank / (uint64_t) p.shard_world);
    d.shard_hi = (uint32_t) ((uint64_t) d.num_buckets * (uint64_t) (p.shard_rank + 1) / (uint64_t) p.shard_world);
    // starve z-buffer ids: the rank takes just the bits it needs, the voxel id the rest (k_starve raises
    // the sticky fault word if a frame ever has more visible voxels than that field can name)
    uint32_t rank_bits = 0;
    while ((1u << rank_bits) < (uint32_t) p.shard_world)
      ++rank_bits;
    d.starve_id_bits = 32u - rank_bits;
    d.shard_tag      = rank_bits ? (uint32_t) p.shard_rank << d.starve_id_bits : 0u;
  } else {
    d.shard_lo = 0, d.shard_hi = d.num_buckets;
    d.shard_tag = 0, d.starve_id_bits = 32u;
  }
}

// Ingest (replaces the blocking DualMatrix::toDevice inside compute(), cuda_matrix.cuh:129).
// The copy goes to the other device image on the copy stream, after the frame that last read that
// image has finished. Pageable caller memory goes through a pinned staging buffer and the call
// returns once the bytes are staged (setters copy: the caller may reuse its buffer on return).
// Page-locked caller memory is read by DMA directly; that transfer is only waited for at the end of
// compute(), so such a buffer must stay unchanged until compute() returns (include/mrhash_b200.h).
// Staging copy of a pageable frame: the source is cold (a freshly decoded image), so one thread gets
// ~10 GB/s out of it and the copy, not the transfer, is what the setter waits for; a few threads in
// 256 KB slices bring it close to the memory system's rate.
static int copy_threads() {
  static const int n = [] {
    const char* e = getenv("MRH_COPY_THREADS"); // tuning knob; default: 8, at most half of the machine
    const int hw  = (int) std::thread::hardware_concurrency();
    const int v   = e ? atoi(e) : 8;
    return std::max(1, std::min(v, std::max(1, hw / 2)));
  }();
  return n;
}
static void staged_copy(void* dst, const void* src, size_t bytes) {
  const size_t slice = 256u << 10;
  const long n       = (long) ((bytes + slice - 1) / slice);
  if (n <= 2) {
    memcpy(dst, src, bytes);
    return;
  }
  const int nthr = copy_threads();
#pragma omp parallel for num_threads(nthr) schedule(static)
  for (long i = 0; i < n; ++i) {
    const size_t o = (size_t) i * slice;
    memcpy((char*) dst + o, (const char*) src + o, std::min(slice, bytes - o));
  }
}

template <typename T, typename F>
static int ingest_upload(mrh_map* m, Ingest& in, const T* src_or_null, size_t n, F fill) {
  if (in.pending_direct && m->ingest_mode != 2) { // set twice without a compute() in between
    CK(cudaEventSynchronize(in.copied[in.which]));
    in.pending_direct = false;
  }
  const int w        = in.which ^ 1;
  const size_t bytes = sizeof(T) * n;
  if (bytes > in.d_cap[w]) {
    CK(cudaEventSynchronize(in.consumed[w]));
    cudaFree(in.d_buf[w]);
    in.d_buf[w] = nullptr;
    CK(cudaMalloc(&in.d_buf[w], bytes));
    in.d_cap[w] = bytes;
  }
  CK(cudaStreamWaitEvent(in.stream, in.consumed[w], 0));
  bool direct = false;
  if (src_or_null && m->ingest_mode != 0) {
    cudaPointerAttributes attr;
    direct = cudaPointerGetAttributes(&attr, src_or_null) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
  }
  const int tl_lane = &in == &m->in_rgb ? 1 : 0;
  mrh_map::TimelineFrame* tl = (&in == &m->in_depth || &in == &m->in_rgb) && m->timeline_up_next[tl_lane] < m->timeline.size() ? &m->timeline[m->timeline_up_next[tl_lane]++] : nullptr;
  if (direct) {
    // mode 2: this device image was filled two setter calls ago; that transfer read a caller buffer
    // which the caller may reuse from now on (the contract of mrh_set_ingest_mode)
    if (m->ingest_mode == 2)
      CK(cudaEventSynchronize(in.copied[w]));
    if (tl)
      CK(cudaEventRecord(tl->up0[tl_lane], in.stream));
    CK(cudaMemcpyAsync(in.d_buf[w], src_or_null, bytes, cudaMemcpyHostToDevice, in.stream));
    if (tl)
      CK(cudaEventRecord(tl->up1[tl_lane], in.stream));
    CK(cudaEventRecord(in.copied[w], in.stream));
    in.pending_direct = true;
  } else {
    if (bytes > in.h_cap[w]) {
      CK(cudaEventSynchronize(in.copied[w]));
      cudaFreeHost(in.h_buf[w]);
      in.h_buf[w] = nullptr;
      CK(cudaMallocHost(&in.h_buf[w], bytes));
      in.h_cap[w] = bytes;
    }
    CK(cudaEventSynchronize(in.copied[w])); // the transfer that last read this staging buffer has finished
    fill((T*) in.h_buf[w]);
    CK(cudaMemcpyAsync(in.d_buf[w], in.h_buf[w], bytes, cudaMemcpyHostToDevice, in.stream));
    CK(cudaEventRecord(in.copied[w], in.stream));
  }
  in.which  = w;
  in.active = true;
  m->h2d_bytes += bytes;
  return 0;
}

// ---------------------------------------------------------------------------------------------
extern "C" {

const char* mrh_last_error(void) {
  return g_last_error.c_str();
}
int mrh_abi_version(void) {
  return MRH_ABI_VERSION;
}

int mrh_params_default(mrh_params* p) {
  if (!p)
    return fail("null params");
  memset(p, 0, sizeof(*p));
  // configurations/replica.cfg:1-18
  p->sdf_truncation             = 0.07f;
  p->sdf_truncation_scale       = 0.f;
  p->integration_weight_sample  = 1;
  p->virtual_voxel_size         = 0.01f;
  p->n_frames_invalidate_voxels = 100;
  p->voxel_extents_scale        = 1;
  p->marching_cubes_threshold   = 1.5f;
  p->min_weight_threshold       = 5;
  p->min_depth                  = 0.01f;
  p->max_depth                  = 30.f;
  p->projective_sdf             = 1;
  p->device                     = -1;
  return 0;
}

static int init_map(mrh_map* m, const mrh_params* p, int dev);

int mrh_create(const mrh_params* p, mrh_map** out) {
  if (!p || !out)
    return fail("null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device: libmrhash_b200 has no CPU path");
  int dev = p->device;
  if (dev < 0)
    CK(cudaGetDevice(&dev));
  CK(cudaSetDevice(dev));
  mrh_map* m = new mrh_map();
  m->p       = *p;
  m->device  = dev;
  // every failure below releases what has been built so far (free_map tolerates a half-built handle)
  if (init_map(m, p, dev)) {
    free_map(m);
    delete m;
    return 1;
  }
  *out = m;
  return 0;
}

static int init_map(mrh_map* m, const mrh_params* p, int dev) {
  CK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&m->ev0));
  CK(cudaEventCreate(&m->ev1));
  CK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&m->copy_stream2, cudaStreamNonBlocking));
  m->in_depth.stream = m->in_points.stream = m->copy_stream;
  m->in_rgb.stream = m->in_normals.stream = getenv("MRH_ONE_COPY_STREAM") ? m->copy_stream : m->copy_stream2;
  for (Ingest* in : {&m->in_depth, &m->in_rgb, &m->in_points, &m->in_normals})
    for (int i = 0; i < 2; ++i) {
      CK(cudaEventCreateWithFlags(&in->copied[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&in->consumed[i], cudaEventDisableTiming));
    }
  CK(cudaMallocHost(&m->h_ctr, sizeof(Counters)));
  CK(cudaMallocHost(&m->h_heap_probe, 2 * sizeof(int)));
  m->h_heap_probe[0] = 0x7FFFFFF0, m->h_heap_probe[1] = -1;
  m->dev.host_probe  = m->h_heap_probe;
  for (int i = 0; i < 8; ++i)
    CK(cudaEventCreate(&m->ev_k[i]));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  m->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("MRH_PDL"))
    m->use_pdl = atoi(e) != 0;
  if (const char* e = getenv("MRH_FRAME"))
    m->use_fused = strcmp(e, "split") != 0;
  if (const char* e = getenv("MRH_BULK_DEPTH"))
    m->use_bulk_depth = atoi(e) != 0;
  if (const char* e = getenv("MRH_FUSED_PREF")) { // "num/den": CTAs that look at the fusion queue first
    int a = 1, b = 4;
    if (sscanf(e, "%d/%d", &a, &b) == 2 && b > 0 && a >= 0)
      m->fused_pref_num = a, m->fused_pref_den = b;
  }
  {
    const char* e      = getenv("MRH_INTEGRATE_CTAS_PER_SM");
    const int per_sm   = e ? std::max(1, atoi(e)) : integrate_ctas_per_sm(); // one wave of resident CTAs
    m->integrate_grid = m->num_sms * per_sm;
  }

  // sizing: geowrapper.cpp:37-54 with params.h:33-37 ratios, in 64-bit arithmetic
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  const double to_alloc = (double) free_b * 0.70;
  uint64_t nblocks      = p->num_sdf_blocks ? p->num_sdf_blocks : (uint64_t) (to_alloc * 0.70 / (12.0 * 512.0));
  if (!p->num_sdf_blocks)
    nblocks = std::min<uint64_t>(nblocks, (1ull << 31) - 1);
  uint64_t nbuckets = p->hash_num_buckets ? p->hash_num_buckets : nblocks;
  nbuckets          = std::min<uint64_t>(nbuckets, ((1ull << 32) - 1) / kBucketSlots);
  uint64_t ntri     = p->max_num_triangles ? p->max_num_triangles : (uint64_t) (to_alloc * 0.25 / 72.0);
  m->num_sdf_blocks    = nblocks;
  m->hash_num_buckets  = nbuckets;
  m->max_num_triangles = ntri;
  m->max_stream_blocks = (uint64_t) (to_alloc * 0.10 / (12.0 * 512.0));
  if (alloc_map(m))
    return 1;
  // geowrapper.cpp:80: default 1x1 spherical camera
  static const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  memcpy(m->pose, I, sizeof(I));
  memcpy(m->cam_in_lidar, I, sizeof(I));
  if (mrh_set_camera(m, 1.f, 1.f, 0.f, 0.f, 1, 1, p->min_depth, p->max_depth, 1) || reset_map(m))
    return 1;
  CK(cudaStreamSynchronize(m->stream));
  return 0;
}

int mrh_destroy(mrh_map* m) {
  if (!m)
    return 0;
  cudaSetDevice(m->device);
  cudaStreamSynchronize(m->stream);
  free_map(m);
  delete m;
  return 0;
}

int mrh_set_camera(mrh_map* m, float fx, float fy, float cx, float cy, int rows, int cols, float min_depth, float max_depth, int camera_model) {
  GUARD(m);
  if (rows <= 0 || cols <= 0)
    return fail("mrh_set_camera: rows/cols must be positive");
  if (camera_model != 0 && camera_model != 1)
    return fail("mrh_set_camera: camera_model must be 0 (pinhole) or 1 (spherical)");
  CameraDev& c = m->cam;
  // camera.cuh:13-39
  c.fx = fx, c.fy = fy, c.ifx = 1.f / fx, c.ify = 1.f / fy, c.cx = cx, c.cy = cy;
  c.rows = (uint32_t) rows, c.cols = (uint32_t) cols;
  c.row_thr   = (int) ((float) rows * 0.5f);
  c.col_thr   = (int) ((float) cols * 0.5f);
  c.min_depth = min_depth, c.max_depth = max_depth, c.model = camera_model;
  m->max_integration_distance = max_depth; // setIntegrationDistance (geowrapper.cpp:111)
  const size_t npix           = (size_t) rows * cols;
  if (npix > m->zbuf_cap) {
    CK(cudaStreamSynchronize(m->stream));
    cudaFree(m->dev.zbuf);
    CK(cudaMalloc(&m->dev.zbuf, sizeof(unsigned long long) * npix));
    m->zbuf_cap = npix;
  }
  refresh_map_params(m);
  return 0;
}

int mrh_set_pose_matrix(mrh_map* m, const float T[16]) {
  if (!m || !T)
    return fail("null argument");
  memcpy(m->pose, T, sizeof(float) * 16);
  return 0;
}

int mrh_set_pose(mrh_map* m, const float t[3], const float q[4]) {
  if (!m || !t || !q)
    return fail("null argument");
  // Eigen::Quaternionf(w, x, y, z).toRotationMatrix() in float, no normalisation (geowrapper.cpp:86-92)
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w;
  const float txx = tx * x, txy = ty * x, txz = tz * x;
  const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  float* o = m->pose;
  o[0] = 1.f - (tyy + tzz), o[1] = txy - twz, o[2] = txz + twy, o[3] = t[0];
  o[4] = txy + twz, o[5] = 1.f - (txx + tzz), o[6] = tyz - twx, o[7] = t[1];
  o[8] = txz - twy, o[9] = tyz + twx, o[10] = 1.f - (txx + tyy), o[11] = t[2];
  o[12] = o[13] = o[14] = 0.f, o[15] = 1.f;
  return 0;
}

int mrh_get_pose_matrix(mrh_map* m, float out[16]) {
  if (!m || !out)
    return fail("null argument");
  memcpy(out, m->pose, sizeof(float) * 16);
  return 0;
}

int mrh_set_camera_in_lidar(mrh_map* m, const float T[16]) {
  if (!m || !T)
    return fail("null argument");
  memcpy(m->cam_in_lidar, T, sizeof(float) * 16);
  return 0;
}

int mrh_set_depth(mrh_map* m, const float* depth, int rows, int cols) {
  GUARD(m);
  if (!depth || rows <= 0 || cols <= 0)
    return fail("GeoWrapper::setDepthImage|input should be a 2D numpy array");
  const size_t n = (size_t) rows * cols;
  if (ingest_upload<float>(m, m->in_depth, depth, n, [&](float* dst) { staged_copy(dst, depth, sizeof(float) * n); }))
    return 1;
  m->depth_ptr = (const float*) m->in_depth.d_buf[m->in_depth.which], m->depth_rows = rows, m->depth_cols = cols;
  return 0;
}

int mrh_set_rgb(mrh_map* m, const uint8_t* rgb, int rows, int cols) {
  GUARD(m);
  if (!rgb || rows <= 0 || cols <= 0)
    return fail("GeoWrapper::setRGBImage|input should be a 3D numpy array");
  const size_t n = (size_t) rows * cols * 3;
  if (ingest_upload<uint8_t>(m, m->in_rgb, rgb, n, [&](uint8_t* dst) { staged_copy(dst, rgb, n); }))
    return 1;
  m->rgb_ptr = (const uint8_t*) m->in_rgb.d_buf[m->in_rgb.which], m->rgb_rows = rows, m->rgb_cols = cols;
  return 0;
}

int mrh_set_rgb_f32(mrh_map* m, const float* rgb, int rows, int cols) {
  GUARD(m);
  if (!rgb || rows <= 0 || cols <= 0)
    return fail("GeoWrapper::setRGBImage|input should be a 3D numpy array");
  const size_t n = (size_t) rows * cols * 3;
  if (ingest_upload<uint8_t>(m, m->in_rgb, nullptr, n, [&](uint8_t* dst) {
        const int nthr = copy_threads();
#pragma omp parallel for num_threads(nthr) schedule(static)
        for (long i = 0; i < (long) n; ++i)
          dst[i] = (uint8_t) rgb[i];
      }))
    return 1;
  m->rgb_ptr = (const uint8_t*) m->in_rgb.d_buf[m->in_rgb.which], m->rgb_rows = rows, m->rgb_cols = cols;
  return 0;
}

int mrh_set_depth_device(mrh_map* m, const float* d_depth, int rows, int cols) {
  if (!m || !d_depth || rows <= 0 || cols <= 0)
    return fail("mrh_set_depth_device: bad argument");
  m->depth_ptr = d_depth, m->depth_rows = rows, m->depth_cols = cols;
  m->in_depth.active = false;
  return 0;
}

int mrh_set_rgb_device(mrh_map* m, const uint8_t* d_rgb, int rows, int cols) {
  if (!m || !d_rgb || rows <= 0 || cols <= 0)
    return fail("mrh_set_rgb_device: bad argument");
  m->rgb_ptr = d_rgb, m->rgb_rows = rows, m->rgb_cols = cols;
  m->in_rgb.active = false;
  return 0;
}

int mrh_set_points(mrh_map* m, const float* points, size_t n, const float* normals) {
  GUARD(m);
  if (!points && n)
    return fail("GeoWrapper::setPointCloud|input should be a 2D numpy array");
  if (n == 0) {
    m->n_points = 0;
    return 0;
  }
  // projective_sdf = false lays the rays along the point normals and measures the sdf along them
  // (voxel_data_structures.cu:957-961, 1251-1254, 1317-1321): it needs them. The reference's kernels
  // read the normal of point i at normals[3 i] (the first of three eigenvectors per point, the layout
  // the MAD-tree path of setPointCloud writes, geowrapper.cpp:386-398); its setPointCloud(points,
  // normals) overload stores ONE vector per point and so makes the kernels read the normal of point
  // 3 i, past the array for i >= n / 3 (geowrapper.cpp:489-500). Here normal i belongs to point i.
  if (!m->p.projective_sdf && !normals)
    return fail("mrh_set_points: projective_sdf = false needs per-point normals (setPointCloud(points, normals))");
  if (ingest_upload<float>(m, m->in_points, points, n * 3, [&](float* dst) { staged_copy(dst, points, sizeof(float) * n * 3); }))
    return 1;
  m->d_points  = (float*) m->in_points.d_buf[m->in_points.which];
  m->d_normals = nullptr;
  if (normals && !m->p.projective_sdf) { // (the projective path never reads them, :1317-1319)
    if (ingest_upload<float>(m, m->in_normals, normals, n * 3, [&](float* dst) { staged_copy(dst, normals, sizeof(float) * n * 3); }))
      return 1;
    m->d_normals = (float*) m->in_normals.d_buf[m->in_normals.which];
  }
  m->n_points = n;
  return 0;
}

// voxel_to_block_1 (voxel_hash_utils.cuh:75-103) is a function of the integer voxel coordinate alone;
// for voxel_extents_scale = 1 it should be v >> 3, but it is computed in metres with a 1e-5 slack, so
// float rounding can disagree far from the origin. One kernel evaluates it for every coordinate of
// the key range and reports the smallest |v| where the two differ: inside that radius the ray walk
// takes the integer shortcut, outside it the reference arithmetic.
namespace {
  __global__ void k_verify_block_shortcut(float size, float ext, int* min_bad) {
    const int n = 1 << 24;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const int v = i - (1 << 23);
      if (voxel_to_block_1(v, size, ext) != (v >> 3))
        atomicMin(min_bad, v < 0 ? -v : v);
    }
  }
} // namespace

static int verify_block_shortcut(mrh_map* m, float size, float ext, int* radius) {
  *radius = 0;
  if (ext != 1.f)
    return 0;
  int* d_min = nullptr;
  CK(cudaMalloc(&d_min, sizeof(int)));
  const int init = 1 << 23;
  CK(cudaMemcpyAsync(d_min, &init, sizeof(int), cudaMemcpyHostToDevice, m->stream));
  k_verify_block_shortcut<<<m->num_sms * 8, 256, 0, m->stream>>>(size, ext, d_min);
  m->launches += 1;
  int h_min = 0;
  CK(cudaMemcpyAsync(&h_min, d_min, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  cudaFree(d_min);
  *radius = std::max(0, h_min - 1);
  return 0;
}

static int compute_frame(mrh_map* m) {
  const bool rgbd = m->depth_ptr && m->rgb_ptr;
  if (rgbd && m->use_fused && (m->shortcut_size != m->p.virtual_voxel_size || m->shortcut_ext != (float) m->p.voxel_extents_scale)) {
    // first frame with these parameters: verify the integer voxel -> block shortcut once
    m->shortcut_size = m->p.virtual_voxel_size, m->shortcut_ext = (float) m->p.voxel_extents_scale;
    if (verify_block_shortcut(m, m->shortcut_size, m->shortcut_ext, &m->shortcut_radius))
      return 1;
  }
  if (rgbd) {
    if (m->depth_rows != (int) m->cam.rows || m->depth_cols != (int) m->cam.cols || m->rgb_rows != m->depth_rows || m->rgb_cols != m->depth_cols)
      return fail("mrh_compute: depth %dx%d / rgb %dx%d do not match the camera %ux%u", m->depth_rows, m->depth_cols, m->rgb_rows, m->rgb_cols, m->cam.rows, m->cam.cols);
  }
  refresh_map_params(m);
  // GeoWrapper::compute (geowrapper.cpp:137-138): page when the pool runs low. The free count is the
  // one probed at the end of the last frame whose probe has arrived (no device wait here).
  if (m->stream_threshold > 0.f && m->frames_probe_seen != ((volatile int*) m->h_heap_probe)[1]) {
    m->frames_probe_seen = ((volatile int*) m->h_heap_probe)[1];
    if ((double) (((volatile int*) m->h_heap_probe)[0] + 1) <= (double) m->stream_threshold * (double) m->num_sdf_blocks) {
      const float centre[3] = {m->pose[3], m->pose[7], m->pose[11]};
      if (stream_radius(m, centre, m->cam.max_depth))
        return 1;
    }
  }
  Ingest* used[4] = {rgbd ? &m->in_depth : nullptr, rgbd ? &m->in_rgb : nullptr, m->n_points ? &m->in_points : nullptr, m->n_points && m->d_normals ? &m->in_normals : nullptr};
  // the ray walk needs the depth image (or the points) only: the colour transfer overlaps it and is
  // waited for in front of the first fusion kernel (integrate_rgbd)
  m->rgb_ready = nullptr;
  for (Ingest* in : used)
    if (in && in->active) {
      if (in == &m->in_rgb)
        m->rgb_ready = in->copied[in->which];
      else
        CK(cudaStreamWaitEvent(m->stream, in->copied[in->which], 0));
    }
  CK(cudaEventRecord(m->ev0, m->stream));
  mrh_map::TimelineFrame* tl = rgbd && m->timeline_next < m->timeline.size() ? &m->timeline[m->timeline_next++] : nullptr;
  if (tl) {
    if (m->rgb_ready) // (the fused frame waits for the colour in front of its launch: make the stamp see it too)
      CK(cudaStreamWaitEvent(m->stream, m->rgb_ready, 0));
    CK(cudaEventRecord(tl->k0, m->stream));
  }
  if (rgbd && integrate_rgbd(m))
    return 1;
  if (tl)
    CK(cudaEventRecord(tl->k1, m->stream));
  if (m->n_points && integrate_points(m))
    return 1;
  CK(cudaEventRecord(m->ev1, m->stream));
  for (Ingest* in : used)
    if (in && in->active)
      CK(cudaEventRecord(in->consumed[in->which], m->stream));
  // transfers that read the caller's page-locked memory directly: the caller owns it again on return
  for (Ingest* in : used)
    if (in && in->pending_direct && m->ingest_mode != 2) {
      CK(cudaEventSynchronize(in->copied[in->which]));
      in->pending_direct = false;
    }
  if (m->stats_pipeline) {
    m->ctr_slot = (m->ctr_slot + 1) % mrh_map::kCtrRing;
    k_snapshot_counters<<<1, 64, 0, m->stream>>>(m->dev.ctr, m->h_ctr_ring + m->ctr_slot);
    m->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(m->ev_ctr[m->ctr_slot], m->stream));
    m->ctr_frames[m->ctr_slot] = m->frames_total;
    m->ctr_filled              = std::min(m->ctr_filled + 1, (int) mrh_map::kCtrRing);
  }
  return 0;
}

/* tuning: device-side timeline of the streaming path. mrh_debug_timeline(m, n, nullptr) arms timing events for
 * the next n RGB-D frames; mrh_debug_timeline(m, 0, out) waits and writes 6 floats per frame (microseconds
 * since the first stamp): depth upload begin / end, colour upload begin / end, frame kernel begin / end */
extern "C" int mrh_debug_timeline(mrh_map* m, int arm_frames, float* out) {
  GUARD(m);
  if (arm_frames > 0) {
    m->timeline.resize((size_t) arm_frames);
    for (auto& t : m->timeline) {
      for (int k = 0; k < 2; ++k) {
        CK(cudaEventCreate(&t.up0[k]));
        CK(cudaEventCreate(&t.up1[k]));
      }
      CK(cudaEventCreate(&t.k0));
      CK(cudaEventCreate(&t.k1));
    }
    m->timeline_next = m->timeline_up_next[0] = m->timeline_up_next[1] = 0;
    return 0;
  }
  CK(cudaDeviceSynchronize());
  const size_t n = std::min(m->timeline_next, std::min(m->timeline_up_next[0], m->timeline_up_next[1]));
  for (size_t i = 0; i < n && out; ++i) {
    const auto& t = m->timeline[i];
    cudaEvent_t ev[6] = {t.up0[0], t.up1[0], t.up0[1], t.up1[1], t.k0, t.k1};
    for (int k = 0; k < 6; ++k) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, m->timeline[0].up0[0], ev[k]));
      out[6 * i + k] = ms * 1e3f;
    }
  }
  for (auto& t : m->timeline) {
    for (int k = 0; k < 2; ++k)
      cudaEventDestroy(t.up0[k]), cudaEventDestroy(t.up1[k]);
    cudaEventDestroy(t.k0), cudaEventDestroy(t.k1);
  }
  m->timeline.clear();
  return (int) n;
}

/* tuning builds (-DMRH_FUSED_DEBUG): reads the 32 timer words of the fused kernel and re-arms them */
extern "C" int mrh_debug_timers(mrh_map* m, unsigned long long out[32]) {
  GUARD(m);
  CK(cudaStreamSynchronize(m->stream));
  CK(cudaMemcpy(out, (char*) m->dev.ctr + offsetof(Counters, dbg), 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  unsigned long long init[32];
  memset(init, 0, sizeof(init));
  init[0] = init[2] = ~0ull;
  CK(cudaMemcpy((char*) m->dev.ctr + offsetof(Counters, dbg), init, sizeof(init), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int mrh_debug_trace(mrh_map* m, unsigned long long* out, size_t n_words) {
  GUARD(m);
  CK(cudaStreamSynchronize(m->stream));
  CK(cudaMemcpy(out, m->dev.reint_keys, n_words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return 0;
}

int mrh_get_block_shortcut_radius(mrh_map* m, int* radius) {
  if (!m || !radius)
    return fail("null argument");
  *radius = m->dev.block_shortcut_radius;
  return 0;
}

int mrh_synchronize(mrh_map* m) {
  GUARD(m);
  CK(cudaStreamSynchronize(m->stream));
  return 0;
}

int mrh_compute(mrh_map* m) {
  GUARD(m);
  if (m->pending_gc)
    return fail("mrh_compute: the previous frame is waiting for mrh_compute_end");
  m->split_zbuf = false;
  return compute_frame(m);
}

int mrh_compute_begin(mrh_map* m, int* needs_zbuf_reduce) {
  GUARD(m);
  if (!needs_zbuf_reduce)
    return fail("null argument");
  if (m->pending_gc)
    return fail("mrh_compute_begin: the previous frame is waiting for mrh_compute_end");
  m->split_zbuf = m->p.shard_world > 1;
  const int rc  = compute_frame(m);
  m->split_zbuf = false;
  *needs_zbuf_reduce = m->pending_gc ? 1 : 0;
  return rc;
}

int mrh_compute_end(mrh_map* m) {
  GUARD(m);
  return finish_gc_tail(m);
}

int mrh_get_zbuf(mrh_map* m, void** d_zbuf, size_t* n_cells) {
  if (!m || !d_zbuf || !n_cells)
    return fail("null argument");
  *d_zbuf  = m->dev.zbuf;
  *n_cells = (size_t) m->cam.rows * m->cam.cols;
  return 0;
}

int mrh_clear_buffers(mrh_map* m) {
  GUARD(m);
  if (mrh_stream_all_out(m))
    return 1;
  m->store.clear();
  return 0;
}

int mrh_get_field(mrh_map* m, const char* name, double* out) {
  if (!m || !name || !out)
    return fail("null argument");
  const std::string n(name);
  const mrh_params& p = m->p;
  // pygeowrapper.cpp:31-61
  if (n == "SDFTruncation") *out = p.sdf_truncation;
  else if (n == "SDFTruncationScale") *out = p.sdf_truncation_scale;
  else if (n == "IntegrationWeightSample") *out = p.integration_weight_sample;
  else if (n == "IntegrationWeightMax") *out = kWeightMax;
  else if (n == "VirtualVoxelSize") *out = p.virtual_voxel_size;
  else if (n == "NumSDFBlocks") *out = (double) m->num_sdf_blocks;
  else if (n == "HashNumBuckets") *out = (double) m->hash_num_buckets;
  else if (n == "HashBucketSize") *out = 10; // params.h:12 (the reference's logical bucket size)
  else if (n == "LinkedListSize") *out = 7;  // params.h:13
  else if (n == "NFramesInvalidateVoxels") *out = p.n_frames_invalidate_voxels;
  else if (n == "VoxelExtentsScale") *out = p.voxel_extents_scale;
  else if (n == "MaxNumSdfBlockIntegrateFromGlobalHash") *out = (double) m->max_stream_blocks;
  else if (n == "MaxNumTrianglesMesh") *out = (double) m->max_num_triangles;
  else if (n == "MinWeightThreshold") *out = p.min_weight_threshold;
  else if (n == "SDFVarThreshold") *out = p.sdf_var_threshold;
  else if (n == "VerticesMergingThreshold") *out = p.vertices_merging_threshold;
  else if (n == "MarchingCubesThreshold") *out = p.marching_cubes_threshold;
  else if (n == "LastMeshStreamMs") *out = m->mesh_ms_stream;
  else if (n == "LastMeshKernelMs") *out = m->mesh_ms_kernel;
  else if (n == "LastMeshMergeMs") *out = m->mesh_ms_merge;
  else if (n == "LastMeshPlyMs") *out = m->mesh_ms_ply;
  else if (n == "Device") *out = m->device;
  else if (n == "StreamThreshold") *out = m->stream_threshold;
  else if (n == "StreamEvents") *out = (double) m->stream_events;
  else if (n == "LastStreamOutBlocks") *out = (double) m->last_stream_out;
  else if (n == "LastStreamInBlocks") *out = (double) m->last_stream_in;
  else if (n == "StreamDuplicates") *out = (double) m->stream_duplicates;
  else if (n == "LowResolutionBlocks") *out = (double) (m->p.sdf_var_threshold > 0.f || m->h_ctr->low_live > 0);
  else
    return fail("mrh_get_field: unknown field '%s'", name);
  return 0;
}

int mrh_set_field(mrh_map* m, const char* name, double v) {
  if (!m || !name)
    return fail("null argument");
  const std::string n(name);
  mrh_params& p = m->p;
  // The reference's setters only change GeoWrapper's own fields after the container has been
  // built (geowrapper.h:98-109), so the getters reflect them but the map does not; same here for
  // the sizing fields. The fusion parameters are the exception we keep live (read every frame).
  if (n == "SDFTruncation") p.sdf_truncation = (float) v;
  else if (n == "SDFTruncationScale") p.sdf_truncation_scale = (float) v;
  else if (n == "IntegrationWeightSample") p.integration_weight_sample = (int) v;
  else if (n == "IntegrationWeightMax") return 0;
  else if (n == "VirtualVoxelSize") return fail("mrh_set_field: VirtualVoxelSize cannot change after construction");
  else if (n == "NumSDFBlocks" || n == "HashNumBuckets" || n == "HashBucketSize" || n == "LinkedListSize" || n == "MaxNumTrianglesMesh") return 0;
  else if (n == "NFramesInvalidateVoxels") p.n_frames_invalidate_voxels = (int) v;
  else if (n == "VoxelExtentsScale") return 0;
  else if (n == "MinWeightThreshold") p.min_weight_threshold = (int) v;
  else if (n == "SDFVarThreshold") p.sdf_var_threshold = (float) v;
  else if (n == "VerticesMergingThreshold") p.vertices_merging_threshold = (float) v;
  else if (n == "MarchingCubesThreshold") p.marching_cubes_threshold = (float) v;
  else if (n == "StreamThreshold") m->stream_threshold = (float) v; // params.h:28, 0 switches paging off
  else
    return fail("mrh_set_field: unknown field '%s'", name);
  refresh_map_params(m);
  return 0;
}

int mrh_set_shard(mrh_map* m, int shard_rank, int shard_world) {
  if (!m)
    return fail("null handle");
  if (shard_world > 1 && (shard_rank < 0 || shard_rank >= shard_world))
    return fail("mrh_set_shard: rank %d outside world %d", shard_rank, shard_world);
  m->p.shard_rank  = shard_rank;
  m->p.shard_world = shard_world;
  refresh_map_params(m);
  return 0;
}

static int fill_stats(const mrh_map* m, const Counters& c, uint64_t frames, mrh_stats* out);

int mrh_set_ingest_mode(mrh_map* m, int mode) {
  GUARD(m);
  if (mode < 0 || mode > 2)
    return fail("mrh_set_ingest_mode: mode must be 0, 1 or 2");
  for (Ingest* in : {&m->in_depth, &m->in_rgb, &m->in_points, &m->in_normals})
    if (in->pending_direct) {
      CK(cudaEventSynchronize(in->copied[in->which]));
      in->pending_direct = false;
    }
  m->ingest_mode = mode;
  return 0;
}

int mrh_set_stats_pipeline(mrh_map* m, int enabled) {
  GUARD(m);
  if (enabled && !m->h_ctr_ring) {
    CK(cudaMallocHost(&m->h_ctr_ring, mrh_map::kCtrRing * sizeof(Counters)));
    for (int i = 0; i < mrh_map::kCtrRing; ++i)
      CK(cudaEventCreateWithFlags(&m->ev_ctr[i], cudaEventDisableTiming));
  }
  m->stats_pipeline = enabled != 0;
  m->ctr_filled     = 0;
  return 0;
}

int mrh_get_stats_pipelined(mrh_map* m, int which, mrh_stats* out) {
  GUARD(m);
  if (!out)
    return fail("null argument");
  if (!m->stats_pipeline)
    return fail("mrh_get_stats_pipelined: enable mrh_set_stats_pipeline first");
  // which = 0: the last compute() (waits for that frame); which = k: k frames earlier. A caller that
  // reads frame n - 2 after submitting frame n never waits for a kernel it has just queued, so the
  // upload of frame n + 1 can be submitted while frame n - 1 is still on the device.
  if (which < 0 || which >= mrh_map::kCtrRing)
    return fail("mrh_get_stats_pipelined: which must be 0 .. %d", mrh_map::kCtrRing - 1);
  if (m->ctr_filled < 1 + which)
    return fail("mrh_get_stats_pipelined: no such frame yet");
  const int slot = (m->ctr_slot - which + mrh_map::kCtrRing) % mrh_map::kCtrRing;
  CK(cudaEventSynchronize(m->ev_ctr[slot]));
  return fill_stats(m, m->h_ctr_ring[slot], m->ctr_frames[slot], out);
}

int mrh_get_stats(mrh_map* m, mrh_stats* out) {
  GUARD(m);
  if (!out)
    return fail("null argument");
  CK(cudaMemcpyAsync(m->h_ctr, m->dev.ctr, offsetof(Counters, dbg), cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return fill_stats(m, *m->h_ctr, m->frames_total, out);
}

static int fill_stats(const mrh_map* m, const Counters& c, uint64_t frames, mrh_stats* out) {
  if (c.fault == 2)
    return fail("starve pass: more visible voxels than the z-buffer id field can name (%u bits with %d shards): weights were not decremented for the excess; use fewer shards or a smaller map per shard",
                m->dev.starve_id_bits, m->p.shard_world);
  if (c.fault)
    return fail("frame kernel watchdog: a wait inside k_frame hit its iteration bound (internal error, results are invalid)");
  out->frames         = frames;
  out->rays_valid     = c.rays_valid;
  out->blocks_new     = c.blocks_new;
  out->blocks_visible = c.blocks_visible;
  out->voxels_updated = c.voxels_updated;
  out->blocks_freed   = c.blocks_freed;
  out->blocks_realloc = c.blocks_realloc;
  out->dropped_heap   = c.dropped_heap;
  out->dropped_table  = c.dropped_table;
  out->live_blocks    = (uint64_t) ((int64_t) m->num_sdf_blocks - ((int64_t) c.heap_counter + 1)) - c.low_parents + c.low_live;
  out->heap_free      = (int64_t) c.heap_counter + 1;
  out->heap_low_free  = (int64_t) c.heap_low_counter + 1;
  out->dropped_updates = c.dropped_updates;
  return 0;
}

int mrh_reset_stats(mrh_map* m) {
  GUARD(m);
  const size_t off = offsetof(Counters, rays_valid);
  CK(cudaMemsetAsync((char*) m->dev.ctr + off, 0, offsetof(Counters, low_parents) - off, m->stream));
  m->frames_total = 0;
  m->h2d_bytes    = 0;
  return 0;
}

int mrh_last_compute_ms(mrh_map* m, float* ms) {
  GUARD(m);
  if (!ms)
    return fail("null argument");
  CK(cudaEventSynchronize(m->ev1));
  CK(cudaEventElapsedTime(ms, m->ev0, m->ev1));
  return 0;
}

int mrh_set_profiling(mrh_map* m, int enabled) {
  GUARD(m);
  CK(cudaStreamSynchronize(m->stream));
  m->profiling = enabled != 0;
  for (int i = 0; i < 8; ++i)
    m->kernel_ms[i] = 0.0, m->kernel_launches[i] = 0;
  return 0;
}

int mrh_get_kernel_times(mrh_map* m, double ms_out[8], uint64_t launches_out[8]) {
  if (!m || !ms_out || !launches_out)
    return fail("null argument");
  for (int i = 0; i < 8; ++i)
    ms_out[i] = m->kernel_ms[i], launches_out[i] = m->kernel_launches[i];
  return 0;
}

int mrh_get_stream(mrh_map* m, void** stream) {
  if (!m || !stream)
    return fail("null argument");
  *stream = (void*) m->stream;
  return 0;
}

int mrh_get_launch_count(mrh_map* m, uint64_t* n) {
  if (!m || !n)
    return fail("null argument");
  *n = m->launches;
  return 0;
}

} // extern "C"

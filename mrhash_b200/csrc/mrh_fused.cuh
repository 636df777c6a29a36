// mrh_fused.cuh — the RGB-D frame as ONE persistent kernel (single resolution: sdf_var_threshold = 0,
// the configuration every shipped RGB-D config uses).
//
// Replaces, for VoxelContainer::integrate (voxel_data_structures.cpp:90-134; files under
// /root/reference/mrhash/src/sdf): calculateCloudKernel (camera.cu:5-26), allocBlocksKernel + its host
// retry loop (voxel_data_structures.cu:758-922), resetCompactHashTable + flatAndReduceHashTable
// (:9-14, :406-449), integrateDepthMapKernel (:1095-1181) and, fused, garbageCollectIdentify / Free
// (:1674-1724, :1827-1854).
//
// k_frame is launched with exactly as many CTAs as are resident at once (SMs x occupancy). Every
// CTA loops over work items taken from three queues in global memory:
//   chunk  128 entries of the input live list: frustum test (isSDFBlockInCameraFrustumApprox),
//          live-list compaction, visible list; blocks that can receive depth this frame are
//          appended to the fusion queue fq[];
//   tile   32 x 4 depth pixels (border tiles of the image are claimed first, tile_of): the depth rows
//          arrive in shared memory through cp.async.bulk on an mbarrier; each warp owns an 8 x 4 pixel
//          patch and first tests whether any of its rays can reach a block that does not exist (the
//          union box of the ray-end blocks, one lane per block) - nearly always not, and the patch is
//          done; otherwise it walks the block DDA; visited keys are de-duplicated in a per-CTA
//          shared-memory set and each distinct missing key is resolved once by a warp-cooperative
//          find-or-insert on the 128-byte bucket rows; a new block goes straight to the output live
//          list, the visible list and fq[];
//   fuse   one entry of fq[]: the block's three planes (6 KB, contiguous) are pulled into shared
//          memory by ONE cp.async.bulk issued before the projection pass and awaited on an mbarrier
//          after it; 4 consecutive-x voxels per thread, 128-bit shared loads and global stores.
// Blocks that already existed do not depend on the ray walk, so fusion of the visible list overlaps
// it; blocks inserted by the walk are fused as they appear. Removing a block (GC) is deferred until
// every walk of the frame is over: a CTA removes the blocks it condemned on its way out (it leaves only
// after seeing a terminator, and those are written after the last tile), so a walker can never
// re-insert a block that the reference would still have seen (its GC runs after allocation), and the
// free stack receives the freed blocks after every allocation of the frame, as in the reference.
// No CTA ever waits for a specific other CTA and nobody polls a shared word: items are claimed with
// blind fetch-and-add tickets, and a CTA with nothing to do polls the 32-byte queue entry its own
// ticket names until a block or a terminator appears there.
#pragma once
#include "mrh_div.cuh"
#include "mrh_kernels.cuh"

namespace mrh {

constexpr int kFuThreads = 128;
constexpr int kFuWarps   = kFuThreads / 32;
#ifndef MRH_BORDER_FIRST
#define MRH_BORDER_FIRST 1 // tiles along the image border are claimed first (tile_of)
#endif
constexpr int kTileW     = 32; // one 128-byte depth row segment per bulk copy
constexpr int kTileH     = 4;
constexpr int kGcLocal   = 8; // condemned blocks a CTA removes itself on its way out; more go to the shared list
constexpr int kSetCells  = 128; // per-tile set of resolved block keys
constexpr unsigned long long kStatsOnly = 1ull << 63;
constexpr unsigned long long kTerminator = ~0ull; // fusion-queue entry that releases a waiting CTA

#ifndef MRH_FUSED_MIN_CTAS
#define MRH_FUSED_MIN_CTAS 7
#endif

enum : int { kItemExit = 0, kItemChunk = 1, kItemTile = 2, kItemFuse = 3 };

#ifdef MRH_FUSED_DEBUG
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define DBG_ADD(i, v) atomicAdd(&m.ctr->dbg[i], (unsigned long long) (v))
#define DBG_MAX(i, v) atomicMax(&m.ctr->dbg[i], (unsigned long long) (v))
#define DBG_MIN(i, v) atomicMin(&m.ctr->dbg[i], (unsigned long long) (v))
#endif

// ---- async-proxy helpers (PTX: mbarrier + cp.async.bulk; SASS: SYNCS / UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(smem_u32(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}
// Every wait in this file is bounded (a bound is only ever reached through a bug): it sets the sticky
// fault word, which fails the next mrh_get_stats / mrh_synchronize, instead of hanging the device.
constexpr uint32_t kSpinBound = 1u << 22;
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity, uint32_t* fault) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > kSpinBound) {
      *reinterpret_cast<volatile uint32_t*>(fault) = 1u;
      return;
    }
}
// global -> shared bulk copy (16-byte aligned, size a multiple of 16), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t ld_vol(const uint32_t* p) {
  return *reinterpret_cast<const volatile uint32_t*>(p);
}
__device__ __forceinline__ uint4 ld_vol_v4(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

struct FusedItem {
  int kind;
  uint32_t arg;
  uint32_t pad[2];
  FuseEntry e;
};

struct FusedSmem {
  alignas(128) uint32_t planes[kBlockBytes / 4]; // sdf[512] | sum_squared[512] | rgbw[512] of the block being fused
  alignas(128) float depth[2][kTileH][kTileW];   // depth rows of the current / next tile
  unsigned long long set[2][kSetCells];          // resolved keys of the current tile; the other one is being cleared
  unsigned long long bar_planes;
  unsigned long long bar_depth[2];
  FusedItem item[2];
  float red_min[kFuWarps];
  uint32_t red_max[kFuWarps];
  uint32_t red_upd[kFuWarps];
  uint32_t warps_done; // warps of this CTA that have finished a chunk / tile item (every 4th completes an item)
  uint32_t n_items;    // chunks + tiles of the frame
  uint32_t rays;       // valid rays walked by this CTA
  uint32_t vis_n;      // entries this CTA appended to the visible list (chunk items)
  uint32_t gc_n;       // blocks this CTA has condemned (first kGcLocal of them in gc_local)
  alignas(16) uint4 gc_local[kGcLocal]; // {table slot, pool block, output live index, -}
  alignas(16) uint4 peek[2]; // the scheduler's look at its next fusion-queue entry (cp.async target)
  int last;
};

// forward: fusion-queue writer (below)
__device__ __forceinline__ void fq_write(const MapDev& m, uint32_t qi, uint32_t tag, unsigned long long key, uint32_t val, uint32_t slot, uint32_t live_idx, uint32_t vis_idx);
__device__ __forceinline__ uint32_t ld_vol(const uint32_t* p);

// One warp of the CTA has finished its share of a chunk / tile item; the last of the four reports
// the item. The warp that reports the LAST item of the frame knows that the fusion queue is final:
// it appends one terminator entry per CTA, so that every CTA still waiting for a queue ticket to
// materialise is released by the entry it is polling itself (nobody polls a shared word).
__device__ __forceinline__ void item_warp_done(const MapDev& m, const FrameDev& f, FusedSmem& sm, int lane) {
  uint32_t last = 0;
  if (lane == 0) {
    const uint32_t n = atomicAdd(&sm.warps_done, 1u) + 1u;
    if ((n & (kFuWarps - 1)) == 0)
      last = atomicAdd(&m.fqs->items_done.v, 1u) + 1u == sm.n_items ? 1u : 0u;
  }
  if (__shfl_sync(0xFFFFFFFFu, last, 0)) {
    const uint32_t end = ld_vol(&m.fqs->fq_count.v); // every producer bumped it before reporting its item
    // every CTA may hold one ticket of its class past the end of the queue
    for (uint32_t i = lane; i < gridDim.x + kQueueShards; i += 32)
      fq_write(m, end + i, f.tag, kTerminator, 0xFFFFFFFFu, 0u, 0u, 0u);
  }
}

// division by a shared divisor: FAST = the compiler's own sequence with the reciprocal hoisted (mrh_div.cuh)
template <bool FAST>
__device__ __forceinline__ float recip_of(float b) {
  return FAST ? div_recip(b) : 0.f;
}
template <bool FAST>
__device__ __forceinline__ float qdiv(float a, float b, float y1) {
  return FAST ? div_fast(a, b, y1) : fdiv(a, b);
}

// ---------------------------------------------------------------------------------------------
// fusion queue
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fq_write(const MapDev& m, uint32_t qi, uint32_t tag, unsigned long long key, uint32_t val, uint32_t slot, uint32_t live_idx, uint32_t vis_idx) {
  FuseEntry* e = m.fq + qi;
  st_v4(reinterpret_cast<char*>(e) + 16, slot, live_idx, vis_idx, tag);
  st_v4(e, (uint32_t) key, (uint32_t) (key >> 32), val, tag);
}

// ---------------------------------------------------------------------------------------------
// chunk role: visibility pass over 128 entries of the input live list
// ---------------------------------------------------------------------------------------------
template <bool FUSE_GC>
__device__ __forceinline__ void role_chunk(const MapDev& m, const FrameDev& f, const CameraDev& cam, const PoseDev& pose, uint32_t chunk, uint32_t n_live, FusedSmem& sm) {
  const unsigned full = 0xFFFFFFFFu;
  const int lane      = threadIdx.x & 31;
  const uint32_t cur  = f.live_cur;
  const uint32_t i    = chunk * kFuThreads + threadIdx.x;
  LiveEntry le        = {kEmpty, kInvalid, 0u};
  if (i < n_live)
    le = m.live[cur][i];
  const uint32_t slot = le.slot;
  const bool alive    = slot != kInvalid;
  i3 b                = {0, 0, 0};
  bool vis = false, maybe = true;
  if (alive) {
    b   = unpack_key(le.key);
    vis = block_in_frustum_ex(cam, pose, b, m.voxel_size, maybe);
  }
  // a visible block whose voxels cannot land in the image only matters if its stored statistics say
  // it must be collected (never in a steady stream: it would have been collected when last updated)
  bool enqueue = vis && !(le.val & 0x80000000u);
  if (enqueue && !maybe) {
    enqueue = false;
    if (FUSE_GC) {
      const BlockStats st = m.stats[le.val];
      enqueue             = gc_predicate(m, st.min_abs_sdf, st.max_weight);
    }
  }
  const unsigned am = __ballot_sync(full, alive);
  const unsigned vm = __ballot_sync(full, vis);
  const unsigned qm = __ballot_sync(full, enqueue);
  uint32_t abase = 0, vbase = 0, qbase = 0;
  if (lane == 0) {
    if (am)
      abase = atomicAdd(&m.ctr->live_count[cur ^ 1u], (uint32_t) __popc(am));
    if (vm) {
      vbase = atomicAdd(&m.ctr->vis_count, (uint32_t) __popc(vm));
      atomicAdd(&sm.vis_n, (uint32_t) __popc(vm));
    }
    if (qm)
      qbase = atomicAdd(&m.fqs->fq_count.v, (uint32_t) __popc(qm));
  }
  abase                = __shfl_sync(full, abase, 0);
  vbase                = __shfl_sync(full, vbase, 0);
  qbase                = __shfl_sync(full, qbase, 0);
  const unsigned lt    = (1u << lane) - 1u;
  const uint32_t my_li = abase + __popc(am & lt);
  const uint32_t my_vi = vbase + __popc(vm & lt);
  if (alive)
    m.live[cur ^ 1u][my_li] = le;
  if (vis) {
    VisEntry e;
    e.x = b.x, e.y = b.y, e.z = b.z;
    e.val            = le.val;
    e.slot           = slot;
    e.live_idx       = my_li;
    e.maybe_in_image = maybe ? 1u : 0u;
    e.pad1           = 0;
    m.vis[my_vi]     = e;
  }
  if (enqueue)
    fq_write(m, qbase + __popc(qm & lt), f.tag, le.key | (maybe ? 0ull : kStatsOnly), le.val, slot, my_li, my_vi);
  // the counters above were bumped by atomics whose results this warp has received: whoever sees the
  // completion count sees them too (both are resolved in L2)
  item_warp_done(m, f, sm, lane);
}

// ---------------------------------------------------------------------------------------------
// tile role: ray walk + block allocation
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool key_in_range_margin(i3 b, int margin) {
  const unsigned span = 2u * (unsigned) (kCoordBias - margin);
  return (unsigned) (b.x + kCoordBias - margin) < span && (unsigned) (b.y + kCoordBias - margin) < span && (unsigned) (b.z + kCoordBias - margin) < span;
}
__device__ __forceinline__ uint32_t set_hash(unsigned long long k) {
  const uint32_t x = (uint32_t) k ^ ((uint32_t) (k >> 21) * 0x9E3779B1u) ^ ((uint32_t) (k >> 42) * 0x85EBCA6Bu);
  return (x ^ (x >> 15)) & (kSetCells - 1);
}
// true when the key has been resolved for this tile already
__device__ __forceinline__ bool set_contains(const unsigned long long* set, unsigned long long key) {
  uint32_t h = set_hash(key);
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    const unsigned long long c = *reinterpret_cast<const volatile unsigned long long*>(set + h);
    if (c == key)
      return true;
    if (c == kNoKey)
      return false;
    h = (h + 1) & (kSetCells - 1);
  }
  return false;
}
__device__ __forceinline__ void set_insert(unsigned long long* set, unsigned long long key) {
  uint32_t h = set_hash(key);
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    const unsigned long long old = atomicCAS(set + h, kNoKey, key);
    if (old == kNoKey || old == key)
      return;
    h = (h + 1) & (kSetCells - 1);
  }
  // no room within the probe limit: the key simply stays unresolved for this tile (it is looked up again)
}

// Warp-cooperative find-or-insert of one block (allocBlock, voxel_data_structures.cu:502-624, without
// the bucket mutex): same protocol as warp_insert (mrh_table.cuh), new blocks are published to the
// output live list, the visible list and the fusion queue.
__device__ __forceinline__ void warp_resolve(const MapDev& m, const CameraDev& cam, const PoseDev& pose, const FrameDev& f, unsigned long long key, int lane) {
  const unsigned full = 0xFFFFFFFFu;
  const i3 b          = unpack_key(key);
  const uint32_t h    = block_hash_fast(m, b);
  if (h < m.shard_lo || h >= m.shard_hi)
    return; // another GPU owns this bucket range
  bool frustum_ok = false;
#pragma unroll 1
  for (int attempt = 0; attempt < 1024; ++attempt) {
    int free_slot                    = -1;
    unsigned long long free_expected = kEmpty;
    bool found = false, end = false;
#pragma unroll 1
    for (int w = 0; w < kMaxWindows && !end; ++w) {
      uint32_t bkt = h + 2u * w + (lane >> 4);
      if (bkt >= m.num_buckets)
        bkt %= m.num_buckets;
      const uint32_t slot        = bkt * kBucketSlots + (lane & 15);
      const unsigned long long k = ld_cg_u64(m.keys + slot); // the per-lane probe has used the L1 view already
      if (__ballot_sync(full, k == key)) {
        found = true;
        break;
      }
      const unsigned fr = __ballot_sync(full, k == kEmpty || k == kTomb);
      const unsigned em = __ballot_sync(full, k == kEmpty);
      if (free_slot < 0 && fr) {
        const int src = __ffs(fr) - 1;
        free_slot     = (int) __shfl_sync(full, slot, src);
        free_expected = __shfl_sync(full, k, src);
      }
      end = em != 0;
    }
    if (found)
      return;
    if (free_slot < 0) {
      if (lane == 0)
        atomicAdd(&m.ctr->dropped_table, 1ull);
      return;
    }
    if (!frustum_ok) {
      const bool in = lane < 8 && block_corner_in_frustum(cam, pose, b, lane, m.voxel_size);
      if (!__ballot_sync(full, in))
        return;
      frustum_ok = true;
    }
    unsigned long long prev = 0;
    if (lane == 0)
      prev = atomicCAS(m.keys + free_slot, free_expected, key);
    prev = __shfl_sync(full, prev, 0);
    if (prev == free_expected) {
      if (lane == 0) {
        const int addr = atomicSub(&m.ctr->heap_counter, 1); // consumeHeapHigh (:33-40)
        if (addr < 0) {
          atomicAdd(&m.ctr->heap_counter, 1);
          atomicExch(m.keys + free_slot, kTomb);
          atomicAdd(&m.ctr->dropped_heap, 1ull);
        } else {
          // the three list reservations and the read of the free stack are independent: issued back
          // to back they cost one L2 round trip instead of four (an insert is a chain of round trips,
          // and the rays of the tile wait for it)
          const uint32_t out = f.live_cur ^ 1u;
          const uint32_t li  = atomicAdd(&m.ctr->live_count[out], 1u);
          const uint32_t vi  = atomicAdd(&m.ctr->vis_count, 1u);
          const uint32_t qi  = atomicAdd(&m.fqs->fq_count.v, 1u);
          const uint32_t val = m.heap[addr];
          m.stats[val]       = {3.40282346638528859812e+38f, 0u};
          m.vals[free_slot]  = val;
          m.live[out][li]    = {key, (uint32_t) free_slot, val};
          VisEntry e;
          e.x = b.x, e.y = b.y, e.z = b.z;
          e.val = val, e.slot = (uint32_t) free_slot, e.live_idx = li, e.maybe_in_image = 1u, e.pad1 = 0;
          m.vis[vi] = e;
          fq_write(m, qi, f.tag, key, val, (uint32_t) free_slot, li, vi);
          atomicAdd(&m.ctr->blocks_new, 1ull);
          atomicAdd(&m.ctr->blocks_visible, 1ull);
        }
      }
      __syncwarp();
      return;
    }
    if (prev == key)
      return; // another warp inserted the same key into the same slot first
    // slot taken by a different key: rescan (reads now bypass L1)
  }
}

// DDA set-up of allocBlocksKernel (:782-822) on the fast path: shared-reciprocal divisions, the integer
// voxel -> block shortcut, no branches; in two stages so that a patch whose rays cannot reach an
// unallocated block stops after the first.
// Stage A: the blocks of the two ray ends (world -> voxel -> block, voxel_hash_utils.cuh:143-151 then
// :75-103). Returns false when an intermediate left the range in which the shortcuts are exact.
__device__ __forceinline__ bool ray_end_blocks_fast(f3 p0, f3 p1, const MapDev& m, float y_size, i3& cur, i3& end) {
  const float size = m.voxel_size;
  const unsigned R = (unsigned) m.block_shortcut_radius;
  bool bad         = false;
  auto to_block = [&](float p) -> int {
    bad |= div_bad(p, true); // the sign of a zero quotient is absorbed by the `+ 0.5 sign(q)` below
    const float q = div_core(p, size, y_size);
    // q + 0.5 * sign(q): i2f(sign) * 0.5 is +-0.5, or +0 for q == +-0
    const float s = (q != 0.f) ? __uint_as_float(0x3F000000u | (__float_as_uint(q) & 0x80000000u)) : 0.f;
    const float a = fadd(q, s);
    // floor(a + 1e-5) for a >= 0, ceil(a - 1e-5) otherwise, then float -> int: both are the truncation of the sum
    const int v = f2i(fadd(a, (a >= 0.f) ? 1e-5f : -1e-5f));
    // the metric block division equals v >> 3 wherever mrh_compute verified it (exhaustively, on the device)
    bad |= (unsigned) v + R > 2u * R;
    return v >> 3;
  };
  cur = {to_block(p0.x), to_block(p0.y), to_block(p0.z)};
  end = {to_block(p1.x), to_block(p1.y), to_block(p1.z)};
  return !bad;
}
// Stage B: direction, step signs, t_max / t_delta. Returns false like stage A.
__device__ __forceinline__ bool dda_steps_fast(DDA& d, f3 p0, f3 p1, i3 cur, i3 end, const MapDev& m) {
  const float size = m.voxel_size;
  const f3 dir     = normalize3({fsub(p1.x, p0.x), fsub(p1.y, p0.y), fsub(p1.z, p0.z)});
  bool bad         = false;
  d.cur            = cur;
  d.istep          = {sign_i(dir.x), sign_i(dir.y), sign_i(dir.z)};
  const float nh   = -fmul(0.5f, size);
  const float cell = fmul(8.f, size);
  const float big  = 3.40282346638528859812e+38f;
  auto axis = [&](float dr, float p, int c, int step, float& tm, float& td) {
    const float bx  = ffma(i2f((c + max(step, 0)) * kBlockSide), size, nh);
    const float ad  = fabsf(dr);
    const bool over = ad < 1e-6f || fabsf(fsub(bx, dr)) < 1e-6f; // the reference overrides its quotients here
    const float y   = div_recip(ad);
    const float n1  = fsub(bx, p);
    const uint32_t sg = __float_as_uint(dr) & 0x80000000u;
    const float q1  = __uint_as_float(__float_as_uint(div_core(n1, ad, y)) ^ sg);
    const float q2  = __uint_as_float(__float_as_uint(div_core(fmul(i2f(step), cell), ad, y)) ^ sg);
    bad |= !over && (div_bad(n1, true) || !(ad <= kDivHi)); // a zero t_max is only compared and added to
    tm = over ? big : q1;
    td = over ? big : q2;
  };
  axis(dir.x, p0.x, cur.x, d.istep.x, d.t_max.x, d.t_delta.x);
  axis(dir.y, p0.y, cur.y, d.istep.y, d.t_max.y, d.t_delta.y);
  axis(dir.z, p0.z, cur.z, d.istep.z, d.t_max.z, d.t_delta.z);
  d.bound = {end.x + d.istep.x, end.y + d.istep.y, end.z + d.istep.z};
  return !bad;
}

// the reference arithmetic, out of line: taken by the rays the fast set-up declines
__device__ __noinline__ DDA dda_init_blocks_ref(f3 p0, f3 p1, float size, float e0, float e1, float e2) {
  const float ext[3] = {e0, e1, e2};
  DDA d;
  d.init(p0, p1, size, ext, true);
  return d; // by value: the caller's walk state must stay in registers
}

// Per-lane presence probe: the first 32-byte sector (4 slots) of the key's home bucket. Inserts take
// the first free slot in probe order, so at any sane load nearly every live key sits there; a miss
// only means "ask the cooperative path". A line cached in L1 may predate an insert by another SM:
// that, too, is only a miss.
__device__ __forceinline__ bool quick_present(const MapDev& m, i3 b, unsigned long long key) {
  const uint32_t h = block_hash_fast(m, b);
  if (h < m.shard_lo || h >= m.shard_hi)
    return true; // another GPU's block: nothing to do here
  const ulonglong2* row = reinterpret_cast<const ulonglong2*>(m.keys + (size_t) h * kBucketSlots);
  const ulonglong2 a = row[0], c = row[1];
  return a.x == key || a.y == key || c.x == key || c.y == key;
}

template <int MODEL, bool FAST>
__device__ __forceinline__ void role_tile(const MapDev& m, const FrameDev& f, const CameraDev& cam, const PoseDev& pose, const float* __restrict__ depth, const FusedItem& it, FusedSmem& sm, uint32_t tile_seq, int bulk_depth) {
  const unsigned full = 0xFFFFFFFFu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pr = lane >> 3, pc = warp * 8 + (lane & 7); // pixel inside the tile: each warp owns an 8 x 4 patch
  const uint32_t col = it.pad[0] + pc, row = it.pad[1] + pr; // tile origin, computed once by the scheduler
  const bool inside  = row < cam.rows && col < cam.cols;
  unsigned long long* set = sm.set[tile_seq & 1u];
  sm.set[(tile_seq & 1u) ^ 1u][tid] = kNoKey; // nobody reads that one before the next tile of this CTA
  float raw = 0.f;
  if (bulk_depth) {
    mbar_wait(&sm.bar_depth[tile_seq & 1u], (tile_seq >> 1) & 1u, &m.ctr->fault);
    if (inside)
      raw = sm.depth[tile_seq & 1u][pr][pc];
  } else if (inside) {
    raw = __ldg(depth + (size_t) row * cam.cols + col);
  }
  bool active = false;
  f3 p0 = {0.f, 0.f, 0.f}, p1 = p0;
  if (inside) {
    const float d = cloud_depth(cam, row, col, raw);
    if (d != 0.f) {
      const float t    = truncation(m.trunc, m.trunc_scale, d);
      const float dmin = fminf(m.max_integration_distance, fsub(d, t));
      const float dmax = fminf(m.max_integration_distance, fadd(d, t));
      if (!(dmin >= dmax)) {
        p0     = se3_mul(pose.R, pose.t, inverse_projection(cam, row, col, dmin));
        p1     = se3_mul(pose.R, pose.t, inverse_projection(cam, row, col, dmax));
        active = true;
      }
    }
  }
  DDA dda;
  if (FAST) {
    // ---- can any ray of this warp's 8 x 4 patch reach a block that does not exist yet? ----
    // On every axis a walk moves monotonically from its start block towards its end block and stops
    // there at the latest (allocBlocksKernel :826-854), so every block a ray visits lies inside the
    // box spanned by its two end blocks, and every block the patch visits inside the union box of
    // its rays. That box is a handful of blocks (the rays of a patch are nearly parallel): one lane
    // per block looks it up with the per-lane probe; if all exist, nothing can be inserted and the
    // whole stepping set-up and walk of the patch is skipped. Any doubt (a ray off the fast path, a
    // box larger than 4 x 4 x 2, a block not found in the first sector of its bucket) falls back to
    // the walk, which is exact.
    i3 cur = {0, 0, 0}, end = {0, 0, 0};
    const bool fast_ok = !active || ray_end_blocks_fast(p0, p1, m, div_recip(m.voxel_size), cur, end);
    const int big_i    = 0x3FFFFFFF;
    const int lox = __reduce_min_sync(full, active ? min(cur.x, end.x) : big_i), hix = __reduce_max_sync(full, active ? max(cur.x, end.x) : -big_i);
    const int loy = __reduce_min_sync(full, active ? min(cur.y, end.y) : big_i), hiy = __reduce_max_sync(full, active ? max(cur.y, end.y) : -big_i);
    const int loz = __reduce_min_sync(full, active ? min(cur.z, end.z) : big_i), hiz = __reduce_max_sync(full, active ? max(cur.z, end.z) : -big_i);
    const bool all_fast = __all_sync(full, fast_ok);
    bool skip           = hix < lox; // no active ray at all
    if (!skip && all_fast) {
      const unsigned nx = (unsigned) (hix - lox) + 1u, ny = (unsigned) (hiy - loy) + 1u, nz = (unsigned) (hiz - loz) + 1u;
      // lanes enumerate a 4 x 4 x 2 box; the axis with at most two blocks takes the 1-bit coordinate
      const unsigned a = lane & 3, b = (lane >> 2) & 3, c = lane >> 4;
      unsigned ix, iy, iz;
      bool fits;
      if (nz <= 2u)
        ix = a, iy = b, iz = c, fits = nx <= 4u && ny <= 4u;
      else if (ny <= 2u)
        ix = a, iz = b, iy = c, fits = nx <= 4u && nz <= 4u;
      else
        iy = a, iz = b, ix = c, fits = nx <= 2u && ny <= 4u && nz <= 4u;
      if (fits) {
        bool present = true;
        if (ix < nx && iy < ny && iz < nz) {
          const i3 bb = {lox + (int) ix, loy + (int) iy, loz + (int) iz};
          present     = key_in_range(bb) && quick_present(m, bb, pack_key(bb));
        }
        skip = __all_sync(full, present);
      }
    }
#ifdef MRH_FUSED_DEBUG
    if (lane == 0 && !skip) { // why is this patch walked?
      DBG_ADD(29, 1);
      if (!all_fast)
        DBG_ADD(18, 1);
      else {
        const unsigned nx = (unsigned) (hix - lox) + 1u, ny = (unsigned) (hiy - loy) + 1u, nz = (unsigned) (hiz - loz) + 1u;
        const bool fits = (nz <= 2u && nx <= 4u && ny <= 4u) || (ny <= 2u && nx <= 4u && nz <= 4u) || (nx <= 2u && ny <= 4u && nz <= 4u);
        DBG_ADD(fits ? 28 : 19, 1);
      }
    }
#endif
    if (skip) {
      const unsigned n_skipped = __popc(__ballot_sync(full, active));
      if (lane == 0 && n_skipped)
        atomicAdd(&sm.rays, n_skipped);
      item_warp_done(m, f, sm, lane);
      return;
    }
    if (active && !(fast_ok && dda_steps_fast(dda, p0, p1, cur, end, m)))
      dda = dda_init_blocks_ref(p0, p1, m.voxel_size, m.ext[0], m.ext[1], m.ext[2]);
  } else if (active) {
    dda = dda_init_blocks_ref(p0, p1, m.voxel_size, m.ext[0], m.ext[1], m.ext[2]);
  }
  const unsigned n_rays = __popc(__ballot_sync(full, active));
  if (lane == 0 && n_rays)
    atomicAdd(&sm.rays, n_rays);
  // A walk takes at most kMaxDDA steps of one cell: if both end blocks keep that distance from the
  // border of the key range, no visited cell can leave it, and the per-step range test is skipped.
  bool in_range = true;
  if (active) {
    const i3 e = {dda.bound.x - dda.istep.x, dda.bound.y - dda.istep.y, dda.bound.z - dda.istep.z};
    in_range   = key_in_range_margin(dda.cur, kMaxDDA) && key_in_range_margin(e, kMaxDDA);
  }
  int iter = 0;
  while (__any_sync(full, active)) {
    unsigned long long key = kNoKey;
    bool want              = false;
    if (active) {
      if (in_range || key_in_range(dda.cur)) {
        key = pack_key(dda.cur);
        // nearly every visited block exists already: one 32-byte load per lane answers that (lanes on
        // the same block share the transaction); the few misses are looked up in the tile's set of
        // keys the cooperative path has dealt with
        want = !quick_present(m, dda.cur, key) && !set_contains(set, key);
      } else {
        atomicAdd(&m.ctr->dropped_table, 1ull);
      }
    }
    unsigned todo = __ballot_sync(full, want);
    while (todo) {
      const int src               = __ffs(todo) - 1;
      const unsigned long long kb = __shfl_sync(full, key, src);
      todo &= ~__ballot_sync(full, want && key == kb);
      warp_resolve(m, cam, pose, f, kb, lane);
      if (lane == 0)
        set_insert(set, kb);
      __syncwarp();
    }
    if (active) {
      active = dda.advance();
      if (++iter >= kMaxDDA)
        active = false;
    }
  }
  item_warp_done(m, f, sm, lane);
}

// ---------------------------------------------------------------------------------------------
// fuse role: integrateDepthMapKernel (:1095-1181) + combineVoxel (voxel_hash_utils.cuh:169-181)
// + garbageCollectIdentify (:1674-1724) on one block
// ---------------------------------------------------------------------------------------------
// combineVoxel + the sum_squared update of one voxel with plain divisions (voxel_hash_utils.cuh:169-181,
// voxel_data_structures.cu:1155-1180), out of line: the voxels whose numerators leave the window of
// the shared-reciprocal divisions come here
struct VoxelOut {
  float sdf, ss;
  uint32_t cw;
};
__device__ __noinline__ VoxelOut combine_ref(float sdf, float sdf0, uint32_t cw, uint32_t pxl, float half_size, uint32_t ws) {
  const uint32_t w0     = cw >> 24;
  const uint32_t c0     = w0 == 0 ? pxl : cw;
  const float curr_mean = w0 > 0 ? sdf0 : sdf;
  const float delta     = fdiv(fsub(sdf, curr_mean), half_size);
  const uint32_t wsum   = w0 + ws;
  const float merged    = fdiv(ffma(sdf, __uint2float_rn(ws), fmul(sdf0, __uint2float_rn(w0))), __uint2float_rn(wsum));
  const float delta2    = fdiv(fsub(sdf, merged), half_size);
  float ss              = fmul(delta, delta2);
  if (fabsf(ss) < 1.175494350822287508e-38f)
    ss = 0.f;
  return {merged, fadd(0.f, ss), (__vavgu4(pxl, c0) & 0x00FFFFFFu) | (min(wsum, (uint32_t) kWeightMax) << 24)};
}

// projectPoint (camera.cuh:131-160) with plain divisions, out of line (same role as combine_ref)
// returns the pixel index, or 0xFFFFFFFF when the point does not project into the image
__device__ __noinline__ uint32_t project_ref(const CameraDev& cam, float cx, float cy, float cz) {
  int r, q;
  if (!project_point(cam, {cx, cy, cz}, r, q))
    return 0xFFFFFFFFu;
  return (uint32_t) r * cam.cols + (uint32_t) q;
}

template <bool FUSE_GC, int MODEL, bool FAST>
__device__ __forceinline__ void role_fuse(const MapDev& m, const FrameDev& f, const CameraDev& cam, const PoseDev& pose, const float* __restrict__ depth, const uint8_t* __restrict__ rgb, const FuseEntry& e, FusedSmem& sm,
                                          uint32_t& planes_seq, unsigned long long& cta_updated) {
  const unsigned full = 0xFFFFFFFFu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t val = e.val;
  if (val & 0x80000000u)
    return; // resolution-1 blocks never reach this kernel
  const bool stats_only = (e.key & kStatsOnly) != 0;
  uint8_t* base         = m.pool + (size_t) val * kBlockBytes;
  if (!stats_only && tid == 0) {
    // every thread has left the previous item (barrier at the top of the item loop): planes[] is free
    mbar_expect_tx(&sm.bar_planes, kBlockBytes);
    bulk_g2s(sm.planes, base, kBlockBytes, &sm.bar_planes);
  }
  float min_abs  = 3.40282346638528859812e+38f;
  uint32_t max_w = 0, n_upd = 0;
  unsigned ok    = 0;
  float4 sdf4    = {0.f, 0.f, 0.f, 0.f}, ss4 = {0.f, 0.f, 0.f, 0.f};
  uint4 cw4      = {0u, 0u, 0u, 0u};
  if (stats_only) {
    const BlockStats st = m.stats[val];
    min_abs = st.min_abs_sdf, max_w = st.max_weight;
  } else {
    // ---- pass 1: projection + depth test, registers only (the bulk copy is in flight) ----
    const i3 b       = unpack_key(e.key & ~kStatsOnly);
    const float size = m.voxel_size;
    const int lx0 = (tid & 1) * 4, ly = (tid >> 1) & 7, lz = tid >> 4;
    const float py = fmul(i2f(b.y * kBlockSide + ly), size), pz = fmul(i2f(b.z * kBlockSide + lz), size);
    float sdf_new[4];
    uint32_t pix[4];
    if (MODEL == 0) {
      float pcz[4], raw[4];
      unsigned in = 0, bad = 0;
      const float by0 = fmul(pose.Ri[1], py), by1 = fmul(pose.Ri[4], py), by2 = fmul(pose.Ri[7], py);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float px = fmul(i2f(b.x * kBlockSide + lx0 + j), size);
        const float cx = fadd(ffma(pose.Ri[2], pz, ffma(pose.Ri[0], px, by0)), pose.ti[0]);
        const float cy = fadd(ffma(pose.Ri[5], pz, ffma(pose.Ri[3], px, by1)), pose.ti[1]);
        const float cz = fadd(ffma(pose.Ri[8], pz, ffma(pose.Ri[6], px, by2)), pose.ti[2]);
        pcz[j] = cz;
        // projectPoint (camera.cuh:131-160); the two quotients share the reciprocal of z, no branch
        const bool zok = !(cz <= cam.min_depth) && cz <= cam.max_depth;
        int r, q;
        if (FAST) {
          const float y1 = div_recip(zok ? cz : 1.f);
          const float ay = fmul(cam.fy, cy), ax = fmul(cam.fx, cx);
          r = f2i(fadd(fadd(div_core(ay, cz, y1), cam.cy), 0.5f));
          q = f2i(fadd(fadd(div_core(ax, cz, y1), cam.cx), 0.5f));
          // (a zero quotient's sign vanishes in `+ cy`)
          if (zok && (div_bad(ay, true) || div_bad(ax, true)))
            bad |= 1u << j;
        } else {
          r = f2i(fadd(fadd(fdiv(fmul(cam.fy, cy), cz), cam.cy), 0.5f));
          q = f2i(fadd(fadd(fdiv(fmul(cam.fx, cx), cz), cam.cx), 0.5f));
        }
        const bool inimg = zok && (uint32_t) r < cam.rows && (uint32_t) q < cam.cols;
        pix[j] = inimg ? (uint32_t) r * cam.cols + (uint32_t) q : 0u;
        in |= inimg ? 1u << j : 0u;
      }
      if (FAST && bad) { // numerators outside the division window: those voxels again, with plain divisions
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (bad >> j & 1u) {
            const f3 pc      = se3_mul(pose.Ri, pose.ti, {fmul(i2f(b.x * kBlockSide + lx0 + j), size), py, pz});
            const uint32_t p = project_ref(cam, pc.x, pc.y, pc.z);
            const bool inimg = p != 0xFFFFFFFFu;
            pix[j] = inimg ? p : 0u;
            in     = (in & ~(1u << j)) | (inimg ? 1u << j : 0u);
          }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        raw[j] = (in >> j & 1u) ? __ldg(depth + pix[j]) : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = (raw[j] <= cam.min_depth || raw[j] > cam.max_depth) ? 0.f : raw[j]; // calculateCloudKernel
        float sdf     = fsub(d, pcz[j]);
        const float t = truncation(m.trunc, m.trunc_scale, d);
        const bool okj = (in >> j & 1u) && d != 0.f && !(d > m.max_integration_distance) && !(sdf <= -t);
        sdf_new[j]    = (sdf >= 0.f) ? fminf(t, sdf) : fmaxf(-t, sdf);
        ok |= okj ? 1u << j : 0u;
      }
    } else {
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const f3 pf = {fmul(i2f(b.x * kBlockSide + lx0 + j), size), py, pz};
        const f3 pc = se3_mul(pose.Ri, pose.ti, pf);
        int r, q;
        sdf_new[j] = 0.f, pix[j] = 0;
        if (project_point(cam, pc, r, q)) {
          const uint32_t p = (uint32_t) r * cam.cols + (uint32_t) q;
          const float d    = cloud_depth(cam, (uint32_t) r, (uint32_t) q, __ldg(depth + p));
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
    }
    // colour of the pixels that will be fused (issued before the wait on the planes)
    uint32_t pxl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pxl[j] = 0;
      if (ok >> j & 1u) {
        const uint8_t* p = rgb + (size_t) pix[j] * 3;
        pxl[j]           = (uint32_t) __ldg(p) | ((uint32_t) __ldg(p + 1) << 8) | ((uint32_t) __ldg(p + 2) << 16);
      }
    }
    // ---- pass 2: the block's planes are in shared memory ----
    mbar_wait(&sm.bar_planes, planes_seq & 1u, &m.ctr->fault);
    sdf4 = reinterpret_cast<const float4*>(sm.planes)[tid];
    ss4  = reinterpret_cast<const float4*>(sm.planes + 512)[tid];
    cw4  = reinterpret_cast<const uint4*>(sm.planes + 1024)[tid];
    float* sdfv   = reinterpret_cast<float*>(&sdf4);
    float* ssv    = reinterpret_cast<float*>(&ss4);
    uint32_t* cwv = reinterpret_cast<uint32_t*>(&cw4);
    const float half_size = fmul(size, 0.5f);
    const float y_half    = recip_of<FAST>(half_size);
    const uint32_t ws     = (uint32_t) m.weight_sample;
    const float wsf       = __uint2float_rn(ws);
    unsigned badc         = 0;
    // about half of the warps of a frame have no voxel to fuse (blocks that straddle the image border or
    // lie behind the surface): they only contribute their stored voxels to the block statistics
    if (__any_sync(full, ok != 0)) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // integrateDepthMapKernel :1155-1180 + combineVoxel: new voxel {sdf, weight_sample, pixel colour};
      // a stored voxel without weight takes the pixel colour. Computed for every voxel, kept for the ok ones.
      const uint32_t cw     = cwv[j];
      const uint32_t w0     = cw >> 24;
      const uint32_t c0     = w0 == 0 ? pxl[j] : cw;
      const float sdf       = sdf_new[j];
      const float curr_mean = w0 > 0 ? sdfv[j] : sdf;
      const uint32_t wsum   = w0 + ws;
      const float wsumf     = __uint2float_rn(wsum);
      const float n_delta   = fsub(sdf, curr_mean);
      const float n_merged  = ffma(sdf, wsf, fmul(sdfv[j], __uint2float_rn(w0)));
      float delta, merged, delta2, n_delta2;
      if (FAST) {
        delta    = div_core(n_delta, half_size, y_half);
        merged   = div_core(n_merged, wsumf, div_recip(wsumf));
        n_delta2 = fsub(sdf, merged);
        delta2   = div_core(n_delta2, half_size, y_half);
        // delta and delta2 only feed the product below, where a zero's sign is erased; merged is stored
        if (div_bad(n_delta, true) || div_bad(n_merged, false) || div_bad(n_delta2, true))
          badc |= 1u << j;
      } else {
        delta    = fdiv(n_delta, half_size);
        merged   = fdiv(n_merged, wsumf);
        n_delta2 = fsub(sdf, merged);
        delta2   = fdiv(n_delta2, half_size);
      }
      // u8(0.5 c0 + 0.5 c1 + 0.5): every term is exact in float, i.e. (c0 + c1 + 1) >> 1 per channel
      const uint32_t rgbn = __vavgu4(pxl[j], c0) & 0x00FFFFFFu;
      const uint32_t wn   = min(wsum, (uint32_t) kWeightMax);
      float ss            = fmul(delta, delta2);
      if (fabsf(ss) < 1.175494350822287508e-38f)
        ss = 0.f; // ATOM.ADD.F32.FTZ of the reference flushes a denormal addend
      if (ok >> j & 1u) {
        sdfv[j] = merged;
        ssv[j]  = fadd(0.f, ss); // Q1: merged_voxel starts from sum_squared = 0
        cwv[j]  = rgbn | (wn << 24);
        ++n_upd;
      }
    }
    badc &= ok;
    if (FAST && badc) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (badc >> j & 1u) {
          const VoxelOut o = combine_ref(sdf_new[j], __uint_as_float(sm.planes[4 * tid + j]), sm.planes[1024 + 4 * tid + j], pxl[j], half_size, ws);
          sdfv[j] = o.sdf, ssv[j] = o.ss, cwv[j] = o.cw;
        }
    }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t w = cwv[j] >> 24;
      if (w != 0)
        min_abs = fminf(min_abs, fabsf(sdfv[j]));
      max_w = max(max_w, w);
    }
    planes_seq++;
  }
  // block-level reduction of the GC statistics (non-negative floats order like their bit patterns)
  const uint32_t wmin = __reduce_min_sync(full, __float_as_uint(min_abs));
  const uint32_t wmax = __reduce_max_sync(full, max_w);
  const uint32_t wupd = __reduce_add_sync(full, n_upd);
  if (lane == 0)
    sm.red_min[warp] = __uint_as_float(wmin), sm.red_max[warp] = wmax, sm.red_upd[warp] = wupd;
  __syncthreads();
  BlockStats st;
  st.min_abs_sdf = fminf(fminf(sm.red_min[0], sm.red_min[1]), fminf(sm.red_min[2], sm.red_min[3]));
  st.max_weight  = max(max(sm.red_max[0], sm.red_max[1]), max(sm.red_max[2], sm.red_max[3]));
  const bool del = FUSE_GC && gc_predicate(m, st.min_abs_sdf, st.max_weight);
  if (tid == 0) {
    cta_updated += sm.red_upd[0] + sm.red_upd[1] + sm.red_upd[2] + sm.red_upd[3];
    if (del) {
      // the block stays in the table until every walk of this frame is over: walkers must still find
      // it. This CTA removes it on its way out (by then it has seen a terminator, which is written
      // after the last tile); past kGcLocal blocks, the last CTA of the frame does.
      const uint32_t k = sm.gc_n++;
      if (k < (uint32_t) kGcLocal) {
        sm.gc_local[k] = make_uint4(e.slot, val, e.live_idx, 0u);
      } else {
        const uint32_t gi = atomicAdd(&m.fqs->gc_count.v, 1u);
        m.gc_list[gi]     = {e.slot, val, e.live_idx, 0u};
      }
      st.min_abs_sdf    = 3.40282346638528859812e+38f;
      st.max_weight     = 0;
    }
    m.stats[val] = st;
  }
  if (del) {
    // deleteVoxel over the whole block (:1838-1841): free pool blocks are always all-zero
    const float4 z = {0.f, 0.f, 0.f, 0.f};
    reinterpret_cast<float4*>(base)[tid]                   = z;
    reinterpret_cast<float4*>(base + kPlaneBytes)[tid]     = z;
    reinterpret_cast<float4*>(base + 2 * kPlaneBytes)[tid] = z;
  } else if (ok) {
    reinterpret_cast<float4*>(base)[tid]                  = sdf4;
    reinterpret_cast<float4*>(base + kPlaneBytes)[tid]    = ss4;
    reinterpret_cast<uint4*>(base + 2 * kPlaneBytes)[tid] = cw4;
  }
}

// ---------------------------------------------------------------------------------------------
// scheduler (thread 0): what does this CTA do next?
// ---------------------------------------------------------------------------------------------
struct FusedPlan {
  uint32_t n_chunks, n_tiles, tiles_x;
  uint32_t tiles_y; // tile rows of the whole image
  uint32_t prefer_fuse; // 1: this CTA looks at the fusion queue before the tile queue
  uint32_t shard;       // queue class of this CTA
};

constexpr uint32_t kNoTicket = 0xFFFFFFFFu;

// Scheduler state of a CTA (thread 0). Queue positions are claimed with fetch-and-add, blindly: a
// tile index past the end closes the tile queue for this CTA; a fusion ticket names an entry that
// may not exist yet - the CTA keeps it and polls that entry (its own 32 bytes, no shared word) until
// the producer's tags appear: a block to fuse, or the terminator written when the last producer is done.
struct SchedState {
  bool chunk_open, tile_open, drained, peeked;
  bool a_issued;       // the depth rows of tile_a are on their way into shared memory
  uint32_t next_chunk; // chunk index assigned / claimed ahead of time
  uint32_t tile_a;     // the next tile of this CTA (assigned, or claimed two items ago)
  uint32_t tile_b;     // the one after it (claim in flight)
  uint32_t a_c0, a_r0; // pixel origin of tile_a
  uint32_t a_tile;     // tile_a (a claim index) as a tile of the band
  uint32_t ticket;     // fusion-queue ticket assigned / claimed ahead of time, waiting for its entry
};

// The first item of each queue is assigned statically (CTA c of class s = c % kQueueShards, rank
// r = c / kQueueShards, owns index s + kQueueShards * r): a frame starts without a single round trip
// to a queue head. Later claims continue behind the statically assigned ones.
__device__ __forceinline__ uint32_t shard_base(uint32_t shard) {
  return (gridDim.x - shard + kQueueShards - 1) / kQueueShards; // CTAs in this class
}
__device__ __forceinline__ uint32_t claim(QueueWord* heads, uint32_t shard) {
  return shard + kQueueShards * (shard_base(shard) + atomicAdd(&heads[shard].v, 1u));
}

// Claim index -> tile of the band: the tiles along the image border first (left column, right column,
// top row, bottom row), then the interior in image order. Whatever enters the view enters through the
// border: those are the tiles whose rays find blocks missing and walk (10+ us each against 3 us for a
// tile that stops at the patch test). Claimed in image order, a border tile of the lower half starts
// when its CTA has finished two other tiles, and the end of the tile phase - hence the terminators,
// hence the frame - waits for it.
__device__ __forceinline__ uint32_t tile_of(const FusedPlan& plan, uint32_t i) {
  const uint32_t tx = plan.tiles_x, ty = plan.tiles_y;
  if (!MRH_BORDER_FIRST || tx < 3u || ty < 3u || plan.n_tiles != tx * ty || i >= plan.n_tiles)
    return i;
  if (i < ty)
    return i * tx; // left column
  i -= ty;
  if (i < ty)
    return i * tx + (tx - 1u); // right column
  i -= ty;
  if (i < tx - 2u)
    return 1u + i; // top row
  i -= tx - 2u;
  if (i < tx - 2u)
    return (ty - 1u) * tx + 1u + i; // bottom row
  i -= tx - 2u;
  return (1u + i / (tx - 2u)) * tx + 1u + i % (tx - 2u);
}

// depth rows of a tile -> shared memory (cp.async.bulk, one 128-byte row segment per copy); completes
// on the buffer's mbarrier. `seq` = how many tiles this CTA has walked before this one: the buffer and
// its barrier alternate, and buffer seq & 1 was last read two tiles ago.
__device__ __forceinline__ void issue_depth(const FrameDev& f, const CameraDev& cam, const float* depth, const FusedPlan& plan, FusedSmem& sm, SchedState& st, uint32_t seq, int bulk_depth) {
  st.a_tile           = tile_of(plan, st.tile_a);
  const uint32_t tile = f.band_lo + st.a_tile;
  const uint32_t tx = tile % plan.tiles_x, ty = tile / plan.tiles_x;
  const uint32_t c0 = tx * kTileW, r0 = ty * kTileH;
  st.a_c0 = c0, st.a_r0 = r0, st.a_issued = true;
  if (bulk_depth) {
    const uint32_t wb = min((uint32_t) kTileW, cam.cols - c0) * 4u;
    const uint32_t nr = min((uint32_t) kTileH, cam.rows - r0);
    unsigned long long* bar = &sm.bar_depth[seq & 1u];
    mbar_expect_tx(bar, wb * nr);
    for (uint32_t r = 0; r < nr; ++r)
      bulk_g2s(&sm.depth[seq & 1u][r][0], depth + (size_t) (r0 + r) * cam.cols + c0, wb, bar);
  }
}

// Issued right after the item is published: the round trips of these claims, and the transfer of the
// NEXT tile's depth rows, overlap the item's work. Tile claims run two items ahead, so that the index
// of the next tile is known here (next_seq: the tile sequence number that tile will have).
__device__ __forceinline__ void prefetch_claims(const MapDev& m, const FrameDev& f, const CameraDev& cam, const float* depth, const FusedPlan& plan, FusedSmem& sm, SchedState& st, uint32_t next_seq, int bulk_depth) {
  if (st.drained)
    return;
  if (st.tile_open) {
    if (st.tile_a == kNoTicket)
      st.tile_a = st.tile_b, st.tile_b = kNoTicket, st.a_issued = false;
    if (st.tile_b == kNoTicket)
      st.tile_b = claim(m.fqs->q_tile, plan.shard);
    if (st.tile_a != kNoTicket && !st.a_issued && st.tile_a < plan.n_tiles)
      issue_depth(f, cam, depth, plan, sm, st, next_seq, bulk_depth);
  }
  if (st.ticket == kNoTicket) {
    if (plan.prefer_fuse || !st.tile_open)
      st.ticket = plan.shard + kQueueShards * atomicAdd(&m.fqs->q_fuse[plan.shard].v, 1u);
  } else {
    // look at the entry the ticket names while the item runs: two 16-byte asynchronous copies
    // (L2 -> shared memory, no registers held) are in flight during it
    const char* src = reinterpret_cast<const char*>(m.fq + st.ticket);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&sm.peek[0])), "l"(src) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&sm.peek[1])), "l"(src + 16) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    st.peeked = true;
  }
}

__device__ __forceinline__ void schedule_next(const MapDev& m, const FrameDev& f, const CameraDev& cam, const float* depth, const FusedPlan& plan, FusedSmem& sm, FusedItem& it, uint32_t tile_seq, int bulk_depth, SchedState& st) {
  FrameQueues* q = m.fqs;
  it.kind        = kItemExit;
  uint32_t sleep_ns = 32;
  for (uint32_t spin = 0;; ++spin) {
    if (spin > kSpinBound) {
      *reinterpret_cast<volatile uint32_t*>(&m.ctr->fault) = 1u;
      return; // exit
    }
    if (st.chunk_open) {
      if (st.next_chunk == kNoTicket)
        st.next_chunk = claim(q->q_chunk, plan.shard);
      const uint32_t ch = st.next_chunk;
      st.next_chunk     = kNoTicket;
      if (ch < plan.n_chunks) {
        it.kind = kItemChunk, it.arg = ch;
        return;
      }
      st.chunk_open = false;
    }
    if (st.ticket != kNoTicket) {
      // has the entry this ticket names been written? (both 16-byte halves carry the frame's tag)
      const char* src = reinterpret_cast<const char*>(m.fq + st.ticket);
      uint4 h0 = make_uint4(0, 0, 0, 0), h1 = h0;
      if (st.peeked) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        const volatile uint32_t* pk = reinterpret_cast<const volatile uint32_t*>(sm.peek);
        h0 = make_uint4(pk[0], pk[1], pk[2], pk[3]), h1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        st.peeked = false;
      }
      if (h0.w != f.tag || h1.w != f.tag)
        h0 = ld_vol_v4(src), h1 = ld_vol_v4(src + 16);
      if (h0.w == f.tag && h1.w == f.tag) {
        const unsigned long long key = (unsigned long long) h0.x | ((unsigned long long) h0.y << 32);
        const uint32_t qi            = st.ticket;
        st.ticket                    = kNoTicket;
        if (key == kTerminator) {
          st.drained = true; // production is over and every real entry has a holder
        } else {
          it.kind = kItemFuse, it.arg = qi;
          it.e.key = key, it.e.val = h0.z, it.e.tag0 = h0.w;
          it.e.slot = h1.x, it.e.live_idx = h1.y, it.e.vis_idx = h1.z, it.e.tag1 = h1.w;
          return;
        }
      }
    }
    if (st.tile_open) {
      if (st.tile_a == kNoTicket) {
        st.tile_a = st.tile_b, st.tile_b = kNoTicket, st.a_issued = false;
        if (st.tile_a == kNoTicket)
          st.tile_a = claim(q->q_tile, plan.shard);
      }
      if (st.tile_a < plan.n_tiles) {
        if (!st.a_issued)
          issue_depth(f, cam, depth, plan, sm, st, tile_seq, bulk_depth);
        it.kind = kItemTile, it.arg = f.band_lo + st.a_tile;
        it.pad[0] = st.a_c0, it.pad[1] = st.a_r0;
        st.tile_a = kNoTicket, st.a_issued = false;
        return;
      }
      st.tile_open = false;
      continue;
    }
    if (st.drained)
      return; // exit
    if (st.ticket == kNoTicket) {
      st.ticket = plan.shard + kQueueShards * atomicAdd(&q->q_fuse[plan.shard].v, 1u);
      continue;
    }
    __nanosleep(sleep_ns);
    sleep_ns = min(sleep_ns * 2u, 512u);
  }
}

// ---------------------------------------------------------------------------------------------
// the frame kernel
// ---------------------------------------------------------------------------------------------
template <bool FUSE_GC, int MODEL, bool FAST>
__global__ void __launch_bounds__(kFuThreads, MRH_FUSED_MIN_CTAS)
    k_frame(MapDev m, FrameDev f, CameraDev cam_arg, const float* __restrict__ depth, const uint8_t* __restrict__ rgb, uint32_t tiles_x, int rearm, int bulk_depth, uint32_t num_sms, uint32_t pref_num, uint32_t pref_den) {
  __shared__ FusedSmem sm;
  // the camera model is a template parameter: with the field pinned, every `model == 0` test in the
  // inlined camera functions folds away and the pinhole instance carries no trigonometry
  CameraDev cam = cam_arg;
  cam.model     = MODEL;
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&sm.bar_planes, 1);
    mbar_init(&sm.bar_depth[0], 1);
    mbar_init(&sm.bar_depth[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    sm.last = 0, sm.warps_done = 0, sm.rays = 0, sm.vis_n = 0, sm.gc_n = 0;
  }
  sm.set[0][tid] = kNoKey;
  sm.set[1][tid] = kNoKey;
  cudaGridDependencySynchronize(); // chained launch: the previous frame has completed (no-op otherwise)
  const PoseDev& pose = frame_pose(f);
  const uint32_t n_live = m.ctr->live_count[f.live_cur];
  FusedPlan plan;
  plan.n_chunks    = (n_live + kFuThreads - 1) / kFuThreads;
  plan.n_tiles     = f.band_hi - f.band_lo;
  plan.tiles_x     = tiles_x;
  plan.tiles_y     = (cam.rows + kTileH - 1) / kTileH;
  plan.prefer_fuse = ((blockIdx.x / num_sms) % pref_den) < pref_num ? 1u : 0u;
  plan.shard       = blockIdx.x % kQueueShards;
  if (tid == 0)
    sm.n_items = plan.n_chunks + plan.n_tiles; // (published by the first barrier of the item loop)
  if (plan.n_chunks + plan.n_tiles == 0 && blockIdx.x == 0 && tid < 32) // nothing will ever be produced
    for (uint32_t i = tid; i < gridDim.x + kQueueShards; i += 32)
      fq_write(m, i, f.tag, kTerminator, 0xFFFFFFFFu, 0u, 0u, 0u);
  uint32_t tile_seq = 0, planes_seq = 0, iter = 0;
  SchedState sched; // thread 0
  sched.chunk_open = true, sched.tile_open = true, sched.drained = false, sched.peeked = false;
  sched.next_chunk = blockIdx.x, sched.tile_a = blockIdx.x, sched.tile_b = kNoTicket, sched.ticket = kNoTicket;
  sched.a_issued = false, sched.a_c0 = 0, sched.a_r0 = 0, sched.a_tile = 0;
  if (tid == 0 && sched.tile_a < plan.n_tiles) // the first tile is assigned statically: its depth rows start moving now
    issue_depth(f, cam, depth, plan, sm, sched, 0, bulk_depth);

  unsigned long long cta_updated = 0;
#ifdef MRH_FUSED_DEBUG
  unsigned long long t_prev = 0;
  if (tid == 0) {
    t_prev = gtimer();
    DBG_MIN(0, t_prev); // first CTA past the dependency wait
    DBG_MAX(1, t_prev); // last CTA past it
  }
#endif
  for (;; ++iter) {
    FusedItem& it = sm.item[iter & 1u];
#ifdef MRH_FUSED_DEBUG
    unsigned long long t_a = 0;
    if (tid == 0) {
      t_a = gtimer();
      if (iter > 0) {
        const int pk = sm.item[(iter & 1u) ^ 1u].kind;
#ifdef MRH_TRACE_LONG
        // the long tile items of the frame: {tile, kind, item begin, item end (thread 0)}
        if (pk == kItemTile && t_a - t_prev > 8000ull) {
          const unsigned long long ti = atomicAdd(&m.ctr->dbg[30], 1ull);
          if (ti < 8192) {
            m.reint_keys[2 * ti]     = ((unsigned long long) sm.item[(iter & 1u) ^ 1u].arg << 32) | (unsigned) pk;
#else
        if (blockIdx.x % 37 == 0) { // timeline of a sample of CTAs: {cta, kind, item begin, item end (thread 0)}
          const unsigned long long ti = atomicAdd(&m.ctr->dbg[30], 1ull);
          if (ti < 8192) {
            m.reint_keys[2 * ti]     = ((unsigned long long) blockIdx.x << 32) | (unsigned) pk;
#endif
            m.reint_keys[2 * ti + 1] = ((t_prev & 0xFFFFFFFFull) << 32) | (t_a & 0xFFFFFFFFull);
          }
        }
        DBG_ADD(4 + pk, t_a - t_prev); // time in role pk (thread 0's view)
        DBG_MAX(24 + pk, t_a - t_prev); // longest item of that kind
        DBG_ADD(8 + pk, 1);
        DBG_MAX(12 + pk, t_a); // when the last item of that kind ended
      }
    }
#endif
    if (tid == 0)
      schedule_next(m, f, cam, depth, plan, sm, it, tile_seq, bulk_depth, sched);
#ifdef MRH_FUSED_DEBUG
    if (tid == 0) {
      const unsigned long long t_b = gtimer();
      DBG_ADD(16, t_b - t_a); // time in the scheduler
      DBG_MAX(17, t_b - t_a);
    }
#endif
    __syncthreads(); // every thread has left the previous item; the next one is published
#ifdef MRH_FUSED_DEBUG
    if (tid == 0) {
      t_prev = gtimer();
    }
#endif
    const int kind = it.kind;
    if (kind == kItemExit)
      break;
    if (tid == 0)
      prefetch_claims(m, f, cam, depth, plan, sm, sched, tile_seq + (kind == kItemTile ? 1u : 0u), bulk_depth);
    if (kind == kItemChunk) {
      role_chunk<FUSE_GC>(m, f, cam, pose, it.arg, n_live, sm);
    } else if (kind == kItemTile) {
      role_tile<MODEL, FAST>(m, f, cam, pose, depth, it, sm, tile_seq, bulk_depth);
      ++tile_seq;
    } else {
      role_fuse<FUSE_GC, MODEL, FAST>(m, f, cam, pose, depth, rgb, it.e, sm, planes_seq, cta_updated);
    }
  }
  // ---- the last CTA to get here finishes the frame ----
#ifdef MRH_FUSED_DEBUG
  if (tid == 0) {
    const unsigned long long t = gtimer();
    DBG_MIN(2, t); // first CTA out of the loop
    DBG_MAX(3, t); // last CTA out of the loop
  }
#endif
  // deferred removal (garbageCollectFreeKernel + deleteHashEntryElement, :1727-1854): tombstone the
  // key, push the pool block back on the free stack (appendHeapHigh :52-56), drop it from the live
  // list. No walk is running any more, and nothing else of this frame looks at these three places.
  Counters* c    = m.ctr;
  FrameQueues* q = m.fqs;
  const uint32_t n_mine = min(sm.gc_n, (uint32_t) kGcLocal);
  if ((uint32_t) tid < n_mine) {
    const uint4 g = sm.gc_local[tid];
    atomicExch(m.keys + g.x, kTomb);
    const int addr   = atomicAdd(&c->heap_counter, 1);
    m.heap[addr + 1] = g.y & 0x7FFFFFFFu;
    m.live[f.live_cur ^ 1u][g.z].slot = kInvalid;
  }
  if (n_mine)
    __syncthreads(); // (uniform) the pushes above have been performed before this CTA reports below
  if (tid == 0) {
    if (cta_updated)
      atomicAdd(&c->voxels_updated, cta_updated);
    if (sm.rays)
      atomicAdd(&c->rays_valid, (unsigned long long) sm.rays);
    if (sm.vis_n)
      atomicAdd(&c->blocks_visible, (unsigned long long) sm.vis_n);
    if (sm.gc_n)
      atomicAdd(&c->blocks_freed, (unsigned long long) sm.gc_n);
    __threadfence();
    // one word tells the last CTA both that it is the last and whether any CTA spilled removals into
    // the shared list (bit 16 and up), so that the common case ends without another round trip
    const unsigned spill = sm.gc_n > (uint32_t) kGcLocal ? 0x10000u : 0u;
    const unsigned done  = atomicAdd(&q->done_ctas.v, 1u + spill) + 1u + spill;
    sm.last              = (done & 0xFFFFu) == gridDim.x ? (int) (1u + (done >> 16)) : 0;
  }
  __syncthreads();
  if (!sm.last)
    return;
  // ---- the last CTA out re-arms the queues for the next frame ----
  if (sm.last > 1) {
    __threadfence();
    const uint32_t n_gc = ld_vol(&q->gc_count.v);
    for (uint32_t i = tid; i < n_gc; i += kFuThreads) {
      const uint4 g = ld_vol_v4(m.gc_list + i);
      atomicExch(m.keys + g.x, kTomb);
      const int addr   = atomicAdd(&c->heap_counter, 1);
      m.heap[addr + 1] = g.y & 0x7FFFFFFFu;
      m.live[f.live_cur ^ 1u][g.z].slot = kInvalid;
    }
    __syncthreads();
  }
  if (tid < kQueueShards)
    q->q_chunk[tid].v = 0, q->q_tile[tid].v = 0, q->q_fuse[tid].v = 0;
  if (tid == 0) {
#ifdef MRH_FUSED_DEBUG
    DBG_MAX(20, gtimer()); // finaliser done
    DBG_MAX(21, ld_vol(&q->fq_count.v));
    DBG_MAX(22, ld_vol(&q->q_tile[0].v));
#endif
    q->fq_count.v = 0, q->items_done.v = 0, q->gc_count.v = 0, q->done_ctas.v = 0;
    if (rearm) {
      c->live_count[f.live_cur] = 0; // next frame's output list
      c->vis_count              = 0;
      // paging trigger (GeoWrapper::compute, geowrapper.cpp:137): the host looks at these two words
      // before the next frame, no copy engine operation and no wait involved
      if (f.pad[0]) {
        m.host_probe[0] = *reinterpret_cast<volatile int*>(&c->heap_counter);
        m.host_probe[1] = (int) f.frame_index;
        __threadfence_system();
      }
    }
  }
}

} // namespace mrh

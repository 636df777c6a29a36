// mrh_table.cuh — device functions of the block hash table shared by every kernel file:
// lookup (getHashEntry, voxel_data_structures.cu:80-127) and warp-cooperative insert
// (allocBlock / reallocBlock, :502-755, without the bucket mutex and host retry loop).
#pragma once
#include "mrh_types.cuh"

namespace mrh {

__device__ __forceinline__ void load_pose(const FrameDev& f, PoseDev& pose) {
  for (int i = 0; i < 9; ++i)
    pose.R[i] = f.R[i];
  for (int i = 0; i < 3; ++i)
    pose.t[i] = f.t[i];
  pose_finish(pose);
}

// calculateHash without the integer division: q = umulhi(h, floor(2^32 / n)) is floor(h / n) or one
// less, so one conditional subtraction finishes the remainder.
__device__ __forceinline__ uint32_t block_hash_fast(const MapDev& m, i3 b) {
  const uint32_t h = ((uint32_t) b.x * 73856093u) ^ ((uint32_t) b.y * 19349669u) ^ ((uint32_t) b.z * 83492791u);
  uint32_t r       = h - __umulhi(h, m.bucket_magic) * m.num_buckets;
  if (r >= m.num_buckets)
    r -= m.num_buckets;
  return r;
}

// Single-thread lookup. Probe order: windows of two buckets starting at the home bucket; a window
// that holds an EMPTY slot terminates the chain (inserts always take the first free slot in this
// order and slots never return to EMPTY, so a present key sits before the first EMPTY).
__device__ __forceinline__ int table_find(const MapDev& m, i3 b) {
  if (!key_in_range(b))
    return -1;
  const unsigned long long key = pack_key(b);
  const uint32_t h             = block_hash_fast(m, b);
#pragma unroll 1
  for (int w = 0; w < kMaxWindows; ++w) {
    bool has_empty = false;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      uint32_t bkt = h + 2u * w + half;
      if (bkt >= m.num_buckets)
        bkt %= m.num_buckets;
      const ulonglong2* row    = reinterpret_cast<const ulonglong2*>(m.keys + (size_t) bkt * kBucketSlots);
#pragma unroll
      for (int i = 0; i < kBucketSlots / 2; ++i) {
        const ulonglong2 k2 = row[i];
        if (k2.x == key)
          return (int) (bkt * kBucketSlots + 2 * i);
        if (k2.y == key)
          return (int) (bkt * kBucketSlots + 2 * i + 1);
        has_empty |= (k2.x == kEmpty) | (k2.y == kEmpty);
      }
    }
    if (has_empty)
      return -1;
  }
  return -1;
}

// Warp-cooperative "insert if absent and in the enlarged frustum" of one block key.
// All 32 lanes call this with the same b. Lane l inspects slot l of the current 32-slot window
// (one 256-byte coalesced read), the match is resolved with ballots, the frustum test of a new
// block is spread over lanes 0..7 (one corner each) and lane 0 claims the slot with one 64-bit CAS
// on the key word: claim and key publication are a single atomic, so two warps racing on the same
// key can never both insert it and no bucket mutex / host retry loop is needed.
// Returns (to every lane) the pool value of the block when this call inserted it, kInvalid otherwise.
// resolution 1 takes a 64-voxel sub-slot from the low heap (reallocBlock, voxel_data_structures.cu:626-755).
// DIRECT_VIS: the new block goes straight to the frame's output live list and visible list (it has
// just passed the frustum test); otherwise it is appended to the input live list live[live_cur].
template <bool FRUSTUM_TEST, bool DIRECT_VIS = false>
__device__ __forceinline__ uint32_t warp_insert(const MapDev& m, const CameraDev& cam, const PoseDev& pose, uint32_t live_cur, i3 b, int lane, int resolution = 0) {
  const unsigned full = 0xFFFFFFFFu;
  if (!key_in_range(b)) {
    if (lane == 0)
      atomicAdd(&m.ctr->dropped_table, 1ull);
    return kInvalid;
  }
  const uint32_t h = block_hash_fast(m, b);
  if (h < m.shard_lo || h >= m.shard_hi)
    return kInvalid; // another GPU owns this bucket range
  const unsigned long long key = pack_key(b);
  bool frustum_ok              = !FRUSTUM_TEST;
#pragma unroll 1
  for (int attempt = 0; attempt < 1024; ++attempt) {
    int free_slot                   = -1;
    unsigned long long free_expected = kEmpty;
    bool found = false, end = false;
#pragma unroll 1
    for (int w = 0; w < kMaxWindows && !end; ++w) {
      uint32_t bkt = h + 2u * w + (lane >> 4);
      if (bkt >= m.num_buckets)
        bkt %= m.num_buckets;
      const uint32_t slot = bkt * kBucketSlots + (lane & 15);
      const unsigned long long k = attempt == 0 ? m.keys[slot] : ld_cg_u64(m.keys + slot);
      if (__ballot_sync(full, k == key)) {
        found = true;
        break;
      }
      const unsigned fr = __ballot_sync(full, k == kEmpty || k == kTomb);
      const unsigned em = __ballot_sync(full, k == kEmpty);
      if (free_slot < 0 && fr) {
        const int src  = __ffs(fr) - 1;
        free_slot      = (int) __shfl_sync(full, slot, src);
        free_expected  = __shfl_sync(full, k, src);
      }
      end = em != 0;
    }
    if (found)
      return kInvalid;
    if (free_slot < 0) {
      if (lane == 0)
        atomicAdd(&m.ctr->dropped_table, 1ull);
      return kInvalid;
    }
    if (!frustum_ok) {
      const bool in = lane < 8 && block_corner_in_frustum(cam, pose, b, lane, m.voxel_size);
      if (!__ballot_sync(full, in))
        return kInvalid;
      frustum_ok = true;
    }
    unsigned long long prev = 0;
    if (lane == 0)
      prev = atomicCAS(m.keys + free_slot, free_expected, key);
    prev = __shfl_sync(full, prev, 0);
    if (prev == free_expected) {
      uint32_t val = kInvalid;
      if (lane == 0) {
        int* counter   = resolution ? &m.ctr->heap_low_counter : &m.ctr->heap_counter;
        const int addr = atomicSub(counter, 1); // consumeHeapHigh / consumeHeapLow (:33-50)
        if (addr < 0) {
          atomicAdd(counter, 1);
          atomicExch(m.keys + free_slot, kTomb);
          atomicAdd(&m.ctr->dropped_heap, 1ull);
        } else {
          if (resolution) {
            val = m.heap_low[addr] | 0x80000000u;
            atomicAdd(&m.ctr->low_live, 1ull);
          } else {
            val          = m.heap[addr];
            m.stats[val] = {3.40282346638528859812e+38f, 0u};
          }
          m.vals[free_slot]  = val;
          const uint32_t cur = DIRECT_VIS ? (live_cur ^ 1u) : live_cur;
          const uint32_t li  = atomicAdd(&m.ctr->live_count[cur], 1u);
          m.live[cur][li]    = {key, (uint32_t) free_slot, val};
          if (DIRECT_VIS) {
            const uint32_t vi = atomicAdd(&m.ctr->vis_count, 1u);
            VisEntry e;
            e.x = b.x, e.y = b.y, e.z = b.z;
            e.val = val, e.slot = (uint32_t) free_slot, e.live_idx = li, e.maybe_in_image = 1u, e.pad1 = 0;
            m.vis[vi] = e;
          }
          atomicAdd(&m.ctr->blocks_new, 1ull);
        }
      }
      return __shfl_sync(full, val, 0);
    }
    if (prev == key)
      return kInvalid; // another warp inserted the same key into the same slot first
    // slot taken by a different key: rescan (reads now bypass L1)
  }
  return kInvalid;
}

} // namespace mrh

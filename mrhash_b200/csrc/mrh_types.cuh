// mrh_types.cuh — device-visible data structures of the map (shared by every translation unit).
//
// Data layout in HBM:
//   keys[capacity] u64   packed block position (21 bits/axis, biased), EMPTY / TOMB sentinels;
//                        bucket = 16 consecutive keys = one 128-byte line, bucket index =
//                        calculateHash(pos) of the reference (voxel_data_structures.cu:151-160)
//   vals[capacity] u32   pool block index (bit 31 = resolution 1, then index is a 64-voxel sub-slot)
//   pool[num_blocks]     6144 B per block as three 2 KB planes: f32 sdf[512] | f32 sum_sq[512] |
//                        u32 rgbw[512] (r | g<<8 | b<<16 | weight<<24) -> 128-bit coalesced access.
//                        A block carved for resolution 1 holds 8 sub-slots of 768 B, each
//                        sdf[64] | sum_sq[64] | rgbw[64].
//   stats[num_blocks]    per pool block {min |sdf| over weight>0, max weight}: the GC predicate of
//                        untouched blocks is answered without re-reading their payload
//   heap[num_blocks]     free stack of pool indices (heap[i] = N-1-i initially, voxel_data_structures.cpp:58-69)
//   heap_low[8*num_blocks] free stack of 64-voxel sub-slot indices (allocateMemoryLow :860-871)
//   live[2][2*num_blocks] dense list of live blocks {key, slot, val} (double buffered, compacted per frame)
//   vis[2*num_blocks]    per-frame list of in-frustum blocks (the reference's compact hash table)
#pragma once
#include "mrh_math.cuh"

namespace mrh {

constexpr unsigned long long kEmpty   = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned long long kTomb    = 0xFFFFFFFFFFFFFFFEull;
constexpr unsigned long long kNoKey   = 0xFFFFFFFFFFFFFFFDull;
constexpr uint32_t kInvalid           = 0xFFFFFFFFu;
constexpr int kBucketSlots            = 16;
constexpr int kMaxWindows             = 64; // 2 buckets per window
constexpr uint32_t kBlockBytes        = 6144;
constexpr uint32_t kPlaneBytes        = 2048;
constexpr int kCoordBias              = 1 << 20;

struct Counters {
  int heap_counter;      // index of the stack top; free = heap_counter + 1 (voxel_data_structures.cpp:148-153)
  int heap_low_counter;
  uint32_t live_count[2];
  uint32_t vis_count;
  uint32_t n_realloc;
  uint32_t n_reintegrate;
  uint32_t n_updates;    // point-cloud path: records emitted this frame
  uint32_t carve_request; // variance path: pool blocks to split into sub-slots this frame
  uint32_t done_ctas;     // k_integrate / k_frame: CTAs that have finished (the last one re-arms the frame counters)
  uint32_t fault;         // a wait of the fused kernel ran into its iteration bound (never in a correct run): sticky, reported by mrh_get_stats
  // per-run totals (read back on demand)
  unsigned long long rays_valid;
  unsigned long long blocks_new;
  unsigned long long blocks_visible;
  unsigned long long voxels_updated;
  unsigned long long blocks_freed;
  unsigned long long blocks_realloc;
  unsigned long long dropped_heap;   // allocBlock "mem size exceed" events
  unsigned long long dropped_table;  // probe sequence exhausted / coordinate out of key range
  unsigned long long dropped_updates; // point-cloud records beyond the staging capacity
  unsigned long long stream_merged;   // streamed-in blocks whose key was live again: fused into the live block
  // map state (not reset by mrh_reset_stats)
  unsigned long long low_parents;    // pool blocks carved into 64-voxel sub-slots
  unsigned long long low_live;       // live resolution-1 blocks
  unsigned long long dbg[32];        // tuning builds only (-DMRH_FUSED_DEBUG): timers of the fused kernel
};

struct BlockStats {
  float min_abs_sdf; // FLT_MAX when no voxel has weight > 0
  uint32_t max_weight;
};

// entry of the dense live list: everything the visibility pass needs, so it never touches the table
struct __align__(16) LiveEntry {
  unsigned long long key;
  uint32_t slot; // kInvalid = removed since the list was built
  uint32_t val;
};

struct __align__(16) VisEntry {
  int x, y, z;
  uint32_t val;
  uint32_t slot;
  uint32_t live_idx;
  uint32_t maybe_in_image; // 0: proven that no voxel of the block projects into the image this frame
  uint32_t pad1;
};

// Work queues of the fused frame kernel (mrh_fused.cuh); all zero between frames. Every word sits in
// its own 128-byte line: each is the target of thousands of atomics / polls per frame, and words that
// share a line share one L2 atomic unit.
struct __align__(128) QueueWord {
  uint32_t v;
  uint32_t pad[31];
};
// Queue heads come in kQueueShards classes: item i belongs to class i % kQueueShards, CTA c claims from
// class c % kQueueShards only. L2 serves returning atomics on ONE address one at a time (~5 ns each
// under contention): a single head per queue throttled the hand-out of a frame's ~4 500 items to ~20 us.
constexpr int kQueueShards = 16;
struct FrameQueues {
  QueueWord q_chunk[kQueueShards]; // next chunk of the input live list (visibility role), per class
  QueueWord q_tile[kQueueShards];  // next ray tile (allocation role), per class
  QueueWord q_fuse[kQueueShards];  // next entry of the fusion queue fq[] to claim (fusion role), per class
  QueueWord fq_count;   // entries reserved in fq[]
  QueueWord items_done; // chunks + tiles completed: fq_count is final once this equals their number
  QueueWord gc_count;   // blocks queued for removal at the end of the frame
  QueueWord done_ctas;  // CTAs that have left the item loop (the last one finishes the frame)
};

// Entry of the fusion queue of the fused frame kernel. Producers (visibility role, allocation role)
// reserve an index with one atomicAdd on fq_count and write the two 16-byte halves with one vector
// store each; both halves carry the frame's tag, so a consumer that polls them knows when the entry
// is complete without any fence (16-byte aligned vector accesses are single transactions).
struct __align__(32) FuseEntry {
  unsigned long long key; // packed block position; bit 63: no voxel can project into the image (stats-only entry)
  uint32_t val;
  uint32_t tag0;
  uint32_t slot;
  uint32_t live_idx;
  uint32_t vis_idx; // position in the visible list (the starve / GC tail kernels index it)
  uint32_t tag1;
};

// block queued for removal by the fused kernel's finaliser
struct __align__(16) GcEntry {
  uint32_t slot, val, live_idx, pad;
};

// record of a gathered block (same layout as mrh_dump_entry)
struct GatherRecord {
  int x, y, z, resolution, ptr;
};

struct FrameDev {
  float R[9];
  float t[3];
  float Ri[9]; // inverse pose, filled on the host by make_frame with pose_finish's exact arithmetic:
  float ti[3]; // the first 24 floats have the layout of PoseDev (frame_pose)
  uint32_t frame_index;
  uint32_t live_cur; // which live list is the input of this frame
  uint32_t pad[2];   // pad[0]: write the paging probe at the end of this frame
  uint32_t tag;      // fusion-queue tag of this frame (never repeats over the life of a map)
  uint32_t band_lo, band_hi; // ray tiles [band_lo, band_hi) are walked by this rank (multi-GPU row bands)
  uint32_t pad2;
};

struct MapDev {
  float voxel_size, trunc, trunc_scale, max_integration_distance;
  float ext[3];
  float gc_threshold; // host: trunc + scale * camera.maxDepth() (voxel_data_structures.cu:1720)
  float var_threshold;
  float mc_threshold;
  int weight_sample;
  int min_weight_threshold;
  int projective;
  uint32_t num_buckets, capacity, num_blocks;
  uint32_t bucket_magic; // floor(2^32 / num_buckets): block_hash_fast needs no integer division
  uint32_t shard_lo, shard_hi; // owned range of reference hash buckets (multi-GPU partition)
  uint32_t shard_tag;          // shard_rank << starve_id_bits: makes the starve z-buffer ids unique across ranks
  uint32_t starve_id_bits;     // low bits of a z-buffer id that name the voxel (512 * visible index + voxel); the rest is the rank
  int block_shortcut_radius;   // voxel_to_block_1(v) == v >> 3 verified for |v| <= radius (0: never take the shortcut)
  int fast_div;                // 1: voxel size, depth range and extents allow the shared-reciprocal divisions (mrh_div.cuh)
  unsigned long long* keys;
  uint32_t* vals;
  uint32_t* heap;
  uint32_t* heap_low;
  uint8_t* pool;
  uint8_t* carved; // 1 = pool block split into 8 resolution-1 sub-slots
  BlockStats* stats;
  LiveEntry* live[2];
  VisEntry* vis;
  FuseEntry* fq;    // fusion queue of the fused frame kernel
  GcEntry* gc_list; // blocks to remove at the end of the frame
  FrameQueues* fqs; // queue words of the fused frame kernel
  VisEntry* realloc_list; // variance path: blocks queued for re-allocation at resolution 1, then the re-integration list
  unsigned long long* reint_keys; // variance path: keys of the blocks re-fused by k_reintegrate
  unsigned long long* zbuf;
  Counters* ctr;
  int* host_probe; // pinned host memory (UVA): {free-stack top, frame tag} written at the end of a frame for the paging trigger
};

// The pose straight out of the kernel parameters (constant bank): no shared-memory staging, no barrier.
static_assert(sizeof(PoseDev) == 24 * sizeof(float), "PoseDev layout");
__device__ __forceinline__ const PoseDev& frame_pose(const FrameDev& f) {
  return *reinterpret_cast<const PoseDev*>(f.R);
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ bool key_in_range(i3 b) {
  return (unsigned) (b.x + kCoordBias) < (2u * kCoordBias) && (unsigned) (b.y + kCoordBias) < (2u * kCoordBias) &&
         (unsigned) (b.z + kCoordBias) < (2u * kCoordBias);
}
__device__ __forceinline__ unsigned long long pack_key(i3 b) {
  return ((unsigned long long) (unsigned) (b.x + kCoordBias) << 42) | ((unsigned long long) (unsigned) (b.y + kCoordBias) << 21) |
         (unsigned long long) (unsigned) (b.z + kCoordBias);
}
__device__ __forceinline__ i3 unpack_key(unsigned long long k) {
  return {(int) ((k >> 42) & 0x1FFFFF) - kCoordBias, (int) ((k >> 21) & 0x1FFFFF) - kCoordBias, (int) (k & 0x1FFFFF) - kCoordBias};
}
__device__ __forceinline__ unsigned long long ld_cg_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

} // namespace mrh

// mrh_halo.cu — boundary exchange for meshing a map that is sharded by hash-bucket range.
//
// The reference is single-GPU; its marching cubes (marching_cubes.cu:72-214) and trilinear sampler
// (voxel_data_structures.cu:260-338) read up to one voxel beyond a block's faces, i.e. the 26
// neighbour blocks, through the hash. Under the multi-GPU partition those neighbours
// mostly live on other ranks. Before meshing, every rank therefore
//   1. lists the neighbour keys it does not own (k_halo_requests, de-duplicated on the device),
//   2. sends each list to the owner (all-to-all of 12-byte keys, mrhash_b200/sharding.py),
//   3. the owner packs, per requested key, only the voxels a neighbour can touch: the one-voxel
//      SHELL of the block, 296 of 512 voxels, as (sdf, rgbw) pairs - sum_squared is never read by
//      the mesher (k_halo_pack), and the records travel back in one all-to-all,
//   4. the requester inserts them as ghost blocks: in the table (so every sampler finds them) but
//      behind the end of its list of owned blocks, so they are never meshed themselves,
//   5. after k_mc_blocks the ghosts are removed again (k_halo_clear).
// A map that holds resolution-1 blocks ships whole blocks instead of shells (full = 1): the mixed-
// resolution sampler reads deeper than one fine voxel.
#include <vector>

#include "mrh_host.h"
#include "mrh_table.cuh"

using namespace mrh;

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_));            \
  } while (0)

namespace {

  constexpr int kShellVoxels = 296; // 8^3 - 6^3
  constexpr int kHaloHeader  = 16;  // int32 resolution (-1 = the owner does not hold the key) + padding

  struct HaloVoxel {
    float sdf;
    uint32_t cw;
  };

  __global__ void __launch_bounds__(256) k_halo_requests(MapDev m, uint32_t live_cur, unsigned long long* __restrict__ set, uint32_t set_mask, int* __restrict__ out_xyz, uint32_t* __restrict__ out_count, uint32_t cap) {
    const uint32_t n_live = m.ctr->live_count[live_cur];
    const uint32_t total  = n_live * 26u;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
      const LiveEntry le = m.live[live_cur][t / 26u];
      if (le.slot == kInvalid)
        continue;
      int d = (int) (t % 26u);
      d += d >= 13; // skip the centre
      const i3 b  = unpack_key(le.key);
      const i3 nb = {b.x + d % 3 - 1, b.y + (d / 3) % 3 - 1, b.z + d / 9 - 1};
      if (!key_in_range(nb))
        continue;
      const uint32_t h = block_hash_fast(m, nb);
      if (h >= m.shard_lo && h < m.shard_hi)
        continue; // this rank owns the neighbour (whether or not it exists)
      const unsigned long long key = pack_key(nb);
      unsigned long long x         = key * 0x9E3779B97F4A7C15ull;
      uint32_t c                   = (uint32_t) (x >> 32) & set_mask;
      for (;;) {
        const unsigned long long old = atomicCAS(set + c, kEmpty, key);
        if (old == key)
          break;
        if (old == kEmpty) {
          const uint32_t o = atomicAdd(out_count, 1u);
          if (o < cap)
            out_xyz[3 * (size_t) o] = nb.x, out_xyz[3 * (size_t) o + 1] = nb.y, out_xyz[3 * (size_t) o + 2] = nb.z;
          break;
        }
        c = (c + 1) & set_mask;
      }
    }
  }

  // one CTA per requested key
  __global__ void __launch_bounds__(128) k_halo_pack(MapDev m, const int* __restrict__ xyz, uint32_t n, int full, const uint16_t* __restrict__ shell_idx, uint8_t* __restrict__ out, uint32_t rec_bytes) {
    __shared__ uint32_t s_val;
    __shared__ int s_found;
    const int tid = threadIdx.x;
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
      if (tid == 0) {
        const i3 b   = {xyz[3 * (size_t) i], xyz[3 * (size_t) i + 1], xyz[3 * (size_t) i + 2]};
        const int sl = table_find(m, b);
        s_found      = sl >= 0;
        s_val        = sl >= 0 ? m.vals[sl] : kInvalid;
      }
      __syncthreads();
      const uint32_t val = s_val;
      const int found    = s_found;
      __syncthreads();
      uint8_t* rec = out + (size_t) i * rec_bytes;
      if (tid < 4)
        reinterpret_cast<int*>(rec)[tid] = tid == 0 ? (found ? (int) (val >> 31) : -1) : 0;
      if (!found)
        continue;
      HaloVoxel* dst = reinterpret_cast<HaloVoxel*>(rec + kHaloHeader);
      if (val >> 31) {
        const uint8_t* base = m.pool + (size_t) (val & 0x7FFFFFFFu) * 768u;
        if (tid < 64)
          dst[tid] = {reinterpret_cast<const float*>(base)[tid], reinterpret_cast<const uint32_t*>(base + 512)[tid]};
      } else {
        const uint8_t* base = m.pool + (size_t) val * kBlockBytes;
        const int n_vox     = full ? kBlockVoxels : kShellVoxels;
        for (int t = tid; t < n_vox; t += 128) {
          const int v = full ? t : (int) shell_idx[t];
          dst[t]      = {reinterpret_cast<const float*>(base)[v], reinterpret_cast<const uint32_t*>(base + 2 * kPlaneBytes)[v]};
        }
      }
    }
  }

  // one CTA per received record; `open` is the map with its shard range opened up
  __global__ void __launch_bounds__(128) k_halo_insert(MapDev open, uint32_t live_cur, const int* __restrict__ xyz, uint32_t n, int full, const uint16_t* __restrict__ shell_idx, const uint8_t* __restrict__ in, uint32_t rec_bytes,
                                                       uint32_t* __restrict__ n_inserted) {
    __shared__ uint32_t s_val;
    const int tid = threadIdx.x, lane = tid & 31;
    CameraDev cam_unused{};
    PoseDev pose_unused{};
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
      const uint8_t* rec = in + (size_t) i * rec_bytes;
      const int res      = reinterpret_cast<const int*>(rec)[0];
      if (res < 0)
        continue;
      if (tid < 32) {
        const i3 b       = {xyz[3 * (size_t) i], xyz[3 * (size_t) i + 1], xyz[3 * (size_t) i + 2]};
        const uint32_t v = warp_insert<false>(open, cam_unused, pose_unused, live_cur, b, lane, res);
        if (lane == 0) {
          s_val = v;
          if (v != kInvalid)
            atomicAdd(n_inserted, 1u);
        }
      }
      __syncthreads();
      const uint32_t val = s_val;
      __syncthreads();
      if (val == kInvalid)
        continue;
      const HaloVoxel* src = reinterpret_cast<const HaloVoxel*>(rec + kHaloHeader);
      if (val >> 31) {
        uint8_t* base = open.pool + (size_t) (val & 0x7FFFFFFFu) * 768u;
        if (tid < 64) {
          reinterpret_cast<float*>(base)[tid]          = src[tid].sdf;
          reinterpret_cast<uint32_t*>(base + 512)[tid] = src[tid].cw;
        }
      } else {
        uint8_t* base   = open.pool + (size_t) val * kBlockBytes;
        const int n_vox = full ? kBlockVoxels : kShellVoxels;
        for (int t = tid; t < n_vox; t += 128) {
          const int v                                           = full ? t : (int) shell_idx[t];
          reinterpret_cast<float*>(base)[v]                     = src[t].sdf;
          reinterpret_cast<uint32_t*>(base + 2 * kPlaneBytes)[v] = src[t].cw;
        }
      }
    }
  }

  // entries [n_owned, live_count) of the live list are ghosts: give their storage back (free pool
  // blocks are all-zero, an invariant of the map) and tombstone their keys
  __global__ void __launch_bounds__(128) k_halo_clear(MapDev m, uint32_t live_cur, uint32_t n_owned) {
    const int tid         = threadIdx.x;
    const uint32_t n_live = m.ctr->live_count[live_cur];
    for (uint32_t li = n_owned + blockIdx.x; li < n_live; li += gridDim.x) {
      const LiveEntry le = m.live[live_cur][li];
      if (le.slot == kInvalid)
        continue;
      if (le.val >> 31) {
        const uint32_t low = le.val & 0x7FFFFFFFu;
        uint8_t* base      = m.pool + (size_t) low * 768u;
        for (int i = tid; i < 192; i += 128)
          reinterpret_cast<uint32_t*>(base)[i] = 0u;
        if (tid == 0) {
          atomicExch(m.keys + le.slot, kTomb);
          const int addr       = atomicAdd(&m.ctr->heap_low_counter, 1);
          m.heap_low[addr + 1] = low;
          atomicAdd(&m.ctr->low_live, (unsigned long long) -1ll);
        }
      } else {
        uint8_t* base  = m.pool + (size_t) le.val * kBlockBytes;
        const float4 z = {0.f, 0.f, 0.f, 0.f};
        reinterpret_cast<float4*>(base)[tid]                   = z;
        reinterpret_cast<float4*>(base + kPlaneBytes)[tid]     = z;
        reinterpret_cast<float4*>(base + 2 * kPlaneBytes)[tid] = z;
        if (tid == 0) {
          atomicExch(m.keys + le.slot, kTomb);
          const int addr   = atomicAdd(&m.ctr->heap_counter, 1);
          m.heap[addr + 1] = le.val;
          m.stats[le.val]  = {3.40282346638528859812e+38f, 0u};
        }
      }
      if (tid == 0)
        atomicAdd(&m.ctr->blocks_new, (unsigned long long) -1ll); // ghosts are not new blocks of the map
    }
  }
  __global__ void k_halo_truncate(MapDev m, uint32_t live_cur, uint32_t n_owned) {
    m.ctr->live_count[live_cur] = n_owned;
  }

  int shell_table(mrh_map* m) {
    if (m->d_shell_idx)
      return 0;
    std::vector<uint16_t> idx;
    for (int v = 0; v < kBlockVoxels; ++v) {
      const int x = v & 7, y = (v >> 3) & 7, z = v >> 6;
      if (x == 0 || x == 7 || y == 0 || y == 7 || z == 0 || z == 7)
        idx.push_back((uint16_t) v);
    }
    CK(cudaMalloc(&m->d_shell_idx, sizeof(uint16_t) * kShellVoxels));
    CK(cudaMemcpy(m->d_shell_idx, idx.data(), sizeof(uint16_t) * kShellVoxels, cudaMemcpyHostToDevice));
    return 0;
  }

  int live_count_now(mrh_map* m, uint32_t& n) {
    CK(cudaMemcpyAsync(m->h_ctr, m->dev.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    n = m->h_ctr->live_count[m->live_cur];
    return 0;
  }

} // namespace

extern "C" {

size_t mrh_halo_record_bytes(int full) {
  return (size_t) kHaloHeader + sizeof(HaloVoxel) * (size_t) (full ? kBlockVoxels : kShellVoxels);
}

int mrh_halo_requests(mrh_map* m, int32_t* d_keys_xyz, size_t cap, size_t* n_out) {
  if (!m || !n_out || (cap && !d_keys_xyz))
    return fail("null argument");
  CK(cudaSetDevice(m->device));
  if (m->halo_active)
    return fail("mrh_halo_requests: ghost blocks of a previous exchange are still in the map (mrh_halo_clear)");
  uint32_t n_live = 0;
  if (live_count_now(m, n_live))
    return 1;
  *n_out = 0;
  if (n_live == 0 || m->p.shard_world <= 1)
    return 0;
  uint32_t bits = 10;
  while ((1ull << bits) < 2ull * 26ull * n_live)
    ++bits;
  unsigned long long* set = nullptr;
  uint32_t* d_count       = nullptr;
  CK(cudaMalloc(&set, sizeof(unsigned long long) << bits));
  CK(cudaMalloc(&d_count, sizeof(uint32_t)));
  CK(cudaMemsetAsync(set, 0xFF, sizeof(unsigned long long) << bits, m->stream));
  CK(cudaMemsetAsync(d_count, 0, sizeof(uint32_t), m->stream));
  k_halo_requests<<<m->num_sms * 8, 256, 0, m->stream>>>(m->dev, m->live_cur, set, (uint32_t) ((1ull << bits) - 1ull), d_keys_xyz, d_count, (uint32_t) std::min<size_t>(cap, 0xFFFFFFFFull));
  m->launches++;
  uint32_t n = 0;
  CK(cudaMemcpyAsync(&n, d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  cudaFree(set), cudaFree(d_count);
  CK(cudaGetLastError());
  *n_out = n; // may exceed cap: the caller then retries with a larger buffer
  return 0;
}

int mrh_halo_pack(mrh_map* m, const int32_t* d_keys_xyz, size_t n, int full, void* d_records) {
  if (!m || (n && (!d_keys_xyz || !d_records)))
    return fail("null argument");
  CK(cudaSetDevice(m->device));
  if (n == 0)
    return 0;
  if (shell_table(m))
    return 1;
  k_halo_pack<<<m->num_sms * 8, 128, 0, m->stream>>>(m->dev, d_keys_xyz, (uint32_t) n, full, m->d_shell_idx, (uint8_t*) d_records, (uint32_t) mrh_halo_record_bytes(full));
  m->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(m->stream));
  return 0;
}

int mrh_halo_insert(mrh_map* m, const int32_t* d_keys_xyz, const void* d_records, size_t n, int full) {
  if (!m || (n && (!d_keys_xyz || !d_records)))
    return fail("null argument");
  CK(cudaSetDevice(m->device));
  if (m->halo_active)
    return fail("mrh_halo_insert: ghost blocks of a previous exchange are still in the map (mrh_halo_clear)");
  uint32_t n_live = 0;
  if (live_count_now(m, n_live))
    return 1;
  m->halo_owned  = n_live;
  m->halo_active = true;
  if (n == 0)
    return 0;
  if (shell_table(m))
    return 1;
  const uint32_t rec = (uint32_t) mrh_halo_record_bytes(full);
  // headers back to the host: how many ghosts there are, and how many need a resolution-1 sub-slot
  std::vector<int32_t> res(n);
  CK(cudaMemcpy2DAsync(res.data(), sizeof(int32_t), d_records, rec, sizeof(int32_t), n, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  uint32_t n_found = 0, n_low = 0;
  for (int32_t r : res)
    n_found += r >= 0, n_low += r == 1;
  if ((uint64_t) n_live + n_found > 2ull * m->num_sdf_blocks)
    return fail("mrh_halo_insert: %u owned + %u ghost blocks exceed the block list", n_live, n_found);
  if (carve_low_blocks(m, n_low))
    return 1;
  uint32_t* d_count = nullptr;
  CK(cudaMalloc(&d_count, sizeof(uint32_t)));
  CK(cudaMemsetAsync(d_count, 0, sizeof(uint32_t), m->stream));
  MapDev open   = m->dev;
  open.shard_lo = 0, open.shard_hi = open.num_buckets;
  k_halo_insert<<<m->num_sms * 8, 128, 0, m->stream>>>(open, m->live_cur, d_keys_xyz, (uint32_t) n, full, m->d_shell_idx, (const uint8_t*) d_records, rec, d_count);
  m->launches++;
  uint32_t n_ins = 0;
  CK(cudaMemcpyAsync(&n_ins, d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  cudaFree(d_count);
  CK(cudaGetLastError());
  m->counters_clean = false;
  if (n_ins != n_found)
    return fail("mrh_halo_insert: only %u of %u ghost blocks fit (pool or table full): the mesh would have holes", n_ins, n_found);
  return 0;
}

int mrh_halo_clear(mrh_map* m) {
  if (!m)
    return fail("null handle");
  CK(cudaSetDevice(m->device));
  if (!m->halo_active)
    return 0;
  k_halo_clear<<<m->num_sms * 8, 128, 0, m->stream>>>(m->dev, m->live_cur, m->halo_owned);
  k_halo_truncate<<<1, 1, 0, m->stream>>>(m->dev, m->live_cur, m->halo_owned);
  m->launches += 2;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(m->stream));
  m->halo_active = false;
  return 0;
}

} // extern "C"

// mrh_grid.cu — GeoWrapper::serializeGrid / deserializeGrid (geowrapper.cpp:567-573): the host chunk
// grid as a checkpoint file, wire-compatible with the reference's
// Serializer<Voxel>::serialize / deserialize (serializer.h:16-75). Host code only.
//
// File = a sequence of records, one per 1 m chunk (any order):
//     u64 size | int32 chunk[3] | size bytes of cista::serialize(ChunkDesc<Voxel>)
// and the cista bytes (cista::offset mode, no header; streamer.cuh:6,21-37,82-165) are laid out as
//     +0    ChunkDesc        = two vector headers: vecSDFBlock_, vecChunkDesc_
//     +48   SDFBlock[n]      = n vector headers (one `data` vector each)
//     ...   Voxel[nv_0], Voxel[nv_1], ...      12-byte voxels, 512 (resolution 0) or 64 (resolution 1) per block
//     (16-byte aligned)  SDFBlockDesc[n]      = {int3 pos; int ptr; int resolution; 12 bytes padding}
// with a vector header = {i64 offset of the elements RELATIVE TO THE HEADER'S OWN ADDRESS, u32 used,
// u32 allocated (= used), u8 self_allocated (= 0), 7 bytes padding}. The layout is not taken from
// cista's sources but pinned against a file written by cista itself (the reference's vendored cista.h built
// into a golden-file writer by the test infrastructure): tests/test_grid_format.py compares byte for byte.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

#include "mrh_fmt.h"
#include "mrh_host.h"

using namespace mrh;

namespace {

  constexpr uint64_t kMaxChunkBytes = 1024ull * 1024ull * 100ull; // serializer.h:13
  constexpr int64_t kNullOffset     = INT64_MIN;                 // cista::offset null pointer

  struct VecHeader {
    int64_t rel;
    uint32_t used, allocated;
    uint8_t self_allocated;
    uint8_t pad[7];
  };
  static_assert(sizeof(VecHeader) == 24, "cista::offset::vector header");
  struct DescRecord {
    int32_t x, y, z, ptr, resolution;
    int32_t pad[3];
  };
  static_assert(sizeof(DescRecord) == 32, "SDFBlockDesc is 16-byte aligned");

  struct ChunkKey {
    int c[3];
    bool operator<(const ChunkKey& o) const {
      return std::lexicographical_compare(c, c + 3, o.c, o.c + 3);
    }
  };

  // Streamer::worldToChunks of the block origin (streamer.cuh:251-260, streamer.cpp:230-232)
  ChunkKey chunk_of(const mrh_dump_entry& e, float size, float ext) {
    ChunkKey k;
    const int pos[3] = {e.x, e.y, e.z};
    for (int a = 0; a < 3; ++a) {
      const float pw = ((float) pos[a] * 8.f) * size;
      const float p  = pw / ext;
      const float sg = (float) ((0.f < p) - (p < 0.f));
      k.c[a]         = (int) (p + sg * 0.5f);
    }
    return k;
  }

  int voxels_of(int resolution) {
    return resolution == 0 ? kBlockVoxels : 64; // 1 << 3 * (finest_block_log2_dim - resolution), streamer.cpp:221-222
  }

  int write_grid(const char* path, const mrh_dump_entry* entries, const uint8_t* voxels, size_t n, float size, float ext) {
    FILE* f = fopen(path, "wb");
    if (!f)
      return fail("Serializer::serialize | Failed to open file for writing: %s", path);
    std::map<ChunkKey, std::vector<size_t>> grid;
    for (size_t i = 0; i < n; ++i)
      grid[chunk_of(entries[i], size, ext)].push_back(i);
    std::vector<uint8_t> buf;
    for (const auto& [key, idx] : grid) {
      const size_t nb = idx.size();
      size_t vox_bytes = 0;
      for (size_t i : idx)
        vox_bytes += (size_t) 12 * voxels_of(entries[i].resolution);
      const size_t off_blocks = 2 * sizeof(VecHeader);
      const size_t off_vox    = off_blocks + nb * sizeof(VecHeader);
      const size_t off_desc   = (off_vox + vox_bytes + 15) & ~(size_t) 15;
      const size_t total      = off_desc + nb * sizeof(DescRecord);
      if (total > kMaxChunkBytes) {
        fclose(f);
        return fail("Serializer::serialize | chunk (%d, %d, %d) needs %zu bytes, more than a reader accepts", key.c[0], key.c[1], key.c[2], total);
      }
      buf.assign(total, 0);
      VecHeader root[2] = {{(int64_t) off_blocks, (uint32_t) nb, (uint32_t) nb, 0, {}}, {(int64_t) off_desc - (int64_t) sizeof(VecHeader), (uint32_t) nb, (uint32_t) nb, 0, {}}};
      memcpy(buf.data(), root, sizeof(root));
      size_t vo = off_vox;
      for (size_t b = 0; b < nb; ++b) {
        const mrh_dump_entry& e = entries[idx[b]];
        const uint32_t nv       = (uint32_t) voxels_of(e.resolution);
        const size_t self       = off_blocks + b * sizeof(VecHeader);
        const VecHeader h       = {(int64_t) vo - (int64_t) self, nv, nv, 0, {}};
        memcpy(buf.data() + self, &h, sizeof(h));
        memcpy(buf.data() + vo, voxels + idx[b] * (size_t) 12 * kBlockVoxels, (size_t) 12 * nv);
        vo += (size_t) 12 * nv;
        const DescRecord d = {e.x, e.y, e.z, e.ptr, e.resolution, {0, 0, 0}};
        memcpy(buf.data() + off_desc + b * sizeof(DescRecord), &d, sizeof(d));
      }
      const uint64_t sz = total;
      if (fwrite(&sz, 8, 1, f) != 1 || fwrite(key.c, 12, 1, f) != 1 || fwrite(buf.data(), 1, total, f) != total) {
        fclose(f);
        return fail("Serializer::serialize | Write failed to: %s", path);
      }
    }
    fclose(f);
    return 0;
  }

  // elements of the vector whose header sits at `at`; false = the header points outside the record
  bool vec_span(const std::vector<uint8_t>& buf, size_t at, size_t elem, size_t& first, uint32_t& count) {
    if (at + sizeof(VecHeader) > buf.size())
      return false;
    VecHeader h;
    memcpy(&h, buf.data() + at, sizeof(h));
    count = h.used;
    first = 0;
    if (count == 0 || h.rel == kNullOffset)
      return (count = 0, true);
    const int64_t target = (int64_t) at + h.rel;
    if (target < 0 || (uint64_t) target + (uint64_t) count * elem > buf.size())
      return false;
    first = (size_t) target;
    return true;
  }

  int read_grid(const char* path, std::vector<mrh_dump_entry>& entries, std::vector<uint8_t>& voxels) {
    FILE* f = fopen(path, "rb");
    if (!f)
      return fail("Serializer::deserialize | Failed to open file for reading: %s", path);
    std::vector<uint8_t> buf;
    uint64_t sz = 0;
    while (fread(&sz, 8, 1, f) == 1) {
      if (sz > kMaxChunkBytes) {
        fclose(f);
        return fail("Serializer::deserialize | Corrupted file: chunk size too large (%llu bytes)", (unsigned long long) sz);
      }
      int chunk[3];
      buf.resize(sz);
      if (fread(chunk, 12, 1, f) != 1 || (sz && fread(buf.data(), 1, sz, f) != sz)) {
        fclose(f);
        return fail("Serializer::deserialize | Read failed from: %s", path);
      }
      size_t blocks_at = 0, descs_at = 0;
      uint32_t nb = 0, nd = 0;
      if (!vec_span(buf, 0, sizeof(VecHeader), blocks_at, nb) || !vec_span(buf, sizeof(VecHeader), sizeof(DescRecord), descs_at, nd) || nb != nd) {
        fclose(f);
        return fail("Serializer::deserialize | Corrupted file: bad chunk record (%d, %d, %d)", chunk[0], chunk[1], chunk[2]);
      }
      for (uint32_t b = 0; b < nb; ++b) {
        DescRecord d;
        memcpy(&d, buf.data() + descs_at + b * sizeof(DescRecord), sizeof(d));
        size_t vox_at = 0;
        uint32_t nv   = 0;
        if (!vec_span(buf, blocks_at + b * sizeof(VecHeader), 12, vox_at, nv) || (d.resolution != 0 && d.resolution != 1) || nv > (uint32_t) kBlockVoxels) {
          fclose(f);
          return fail("Serializer::deserialize | Corrupted file: bad block in chunk (%d, %d, %d)", chunk[0], chunk[1], chunk[2]);
        }
        entries.push_back({d.x, d.y, d.z, d.resolution, d.ptr});
        const size_t base = voxels.size();
        voxels.resize(base + (size_t) 12 * kBlockVoxels, 0);
        memcpy(voxels.data() + base, buf.data() + vox_at, (size_t) 12 * nv);
      }
    }
    fclose(f);
    return 0;
  }

} // namespace

extern "C" {

size_t mrh_format_g6(double v, char* dst) {
  return fmt_g6(v, dst);
}

int mrh_grid_write(const char* path, const mrh_dump_entry* entries, const void* voxels, size_t n, float virtual_voxel_size, float voxel_extents) {
  if (!path || (n && (!entries || !voxels)))
    return fail("null argument");
  if (!(virtual_voxel_size > 0.f) || !(voxel_extents > 0.f))
    return fail("mrh_grid_write: voxel size and extents must be positive");
  return write_grid(path, entries, (const uint8_t*) voxels, n, virtual_voxel_size, voxel_extents);
}

int mrh_grid_read(const char* path, mrh_dump_entry* entries, void* voxels, size_t max_entries, size_t* n_out) {
  if (!path || !n_out)
    return fail("null argument");
  std::vector<mrh_dump_entry> e;
  std::vector<uint8_t> v;
  if (read_grid(path, e, v))
    return 1;
  *n_out = e.size();
  if (!entries)
    return 0;
  const size_t n = std::min(max_entries, e.size());
  if (n) {
    memcpy(entries, e.data(), n * sizeof(mrh_dump_entry));
    if (voxels)
      memcpy(voxels, v.data(), n * (size_t) 12 * kBlockVoxels);
  }
  return 0;
}

int mrh_serialize_grid(mrh_map* m, const char* path) {
  if (!m || !path)
    return fail("null argument");
  printf("Serializer::serialize | writing to %s\n", path);
  const HostStore& st = m->store;
  std::vector<mrh_dump_entry> e(st.recs.size());
  for (size_t i = 0; i < e.size(); ++i)
    e[i] = {st.recs[i].x, st.recs[i].y, st.recs[i].z, st.recs[i].resolution, st.recs[i].ptr};
  if (write_grid(path, e.data(), (const uint8_t*) st.voxels.data(), e.size(), m->p.virtual_voxel_size, (float) m->p.voxel_extents_scale))
    return 1;
  printf("Serializer::serialize | written %s\n", path);
  return 0;
}

int mrh_deserialize_grid(mrh_map* m, const char* path) {
  if (!m || !path)
    return fail("null argument");
  printf("Serializer::deserialize | from %s\n", path);
  std::vector<mrh_dump_entry> e;
  std::vector<uint8_t> v;
  if (read_grid(path, e, v))
    return 1;
  // deserialize() replaces the chunks the file names and keeps the others (serializer.h:71): blocks of
  // the store that live in a chunk of the file make way for the file's
  HostStore& st = m->store;
  const float size = m->p.virtual_voxel_size, ext = (float) m->p.voxel_extents_scale;
  std::map<ChunkKey, bool> in_file;
  for (const mrh_dump_entry& x : e)
    in_file[chunk_of(x, size, ext)] = true;
  HostStore keep;
  const size_t words = (size_t) 3 * kBlockVoxels;
  for (size_t i = 0; i < st.recs.size(); ++i) {
    const mrh_dump_entry x = {st.recs[i].x, st.recs[i].y, st.recs[i].z, st.recs[i].resolution, st.recs[i].ptr};
    if (in_file.count(chunk_of(x, size, ext)))
      continue;
    keep.recs.push_back(st.recs[i]);
    keep.voxels.insert(keep.voxels.end(), st.voxels.begin() + i * words, st.voxels.begin() + (i + 1) * words);
  }
  st.recs.swap(keep.recs);
  st.voxels.swap(keep.voxels);
  if (mrh_store_append(m, e.data(), v.data(), e.size()))
    return 1;
  printf("Serializer::deserialize | read %s\n", path);
  return 0;
}

} // extern "C"

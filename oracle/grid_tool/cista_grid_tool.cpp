// cista_grid_tool.cpp — test infrastructure (never product code): writes / reads the reference's
// serializeGrid file with the reference's OWN serialization library, cista (vendored by the reference
// at mrhash/src/sdf/utils/cista.h and included from there, never copied), so that the product's
// cista-free writer / reader (csrc/mrh_grid.cpp) can be pinned byte for byte.
//
// The three record types are restated from the reference, member for member:
//   Voxel         voxel_hash_utils.cuh:8-22      {float sdf; float sum_squared; uchar3 rgb; uchar weight}
//   SDFBlock<T>   streamer.cuh:21-37             {data::vector<T> data}
//   SDFBlockDesc  streamer.cuh:39-80             {int3 pos; int ptr; int resolution} __align__(16)
//   ChunkDesc<T>  streamer.cuh:82-165            {data::vector<SDFBlock<T>> vecSDFBlock_; data::vector<SDFBlockDesc> vecChunkDesc_}
// with `namespace data = cista::offset` (streamer.cuh:6), and the file framing of
// Serializer<T>::serialize / deserialize (serializer.h:16-75): u64 size | int[3] chunk | cista bytes.
//
//   cista_grid_tool encode <blocks.bin> <grid.bin>    blocks.bin: u32 n, then per block
//                                                     int chunk[3], int pos[3], int ptr, int resolution,
//                                                     u32 n_voxels, n_voxels x 12 bytes
//   cista_grid_tool decode <grid.bin> <blocks.bin>    the inverse (blocks in file order)
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <tuple>
#include <vector>

#include "cista.h"

namespace data = cista::offset;

struct uchar3 {
  unsigned char x, y, z;
};
struct int3 {
  int x, y, z;
};
struct Voxel {
  float sdf         = 0.f;
  float sum_squared = 0.f;
  uchar3 rgb        = {0, 0, 0};
  unsigned char weight = 0;
  auto cista_members() const {
    return std::tie(sdf, sum_squared, rgb, weight);
  }
};
static_assert(sizeof(Voxel) == 12, "Voxel is 12 bytes");
struct SDFBlock {
  data::vector<Voxel> data;
  auto cista_members() {
    return std::tie(data);
  }
};
struct alignas(16) SDFBlockDesc {
  int3 pos;
  int ptr;
  int resolution;
  auto cista_members() {
    return std::tie(pos, ptr, resolution);
  }
};
struct ChunkDesc {
  data::vector<SDFBlock> vecSDFBlock_;
  data::vector<SDFBlockDesc> vecChunkDesc_;
  auto cista_members() {
    return std::tie(vecSDFBlock_, vecChunkDesc_);
  }
};

struct ChunkKey {
  int c[3];
  bool operator<(const ChunkKey& o) const {
    return std::tie(c[0], c[1], c[2]) < std::tie(o.c[0], o.c[1], o.c[2]);
  }
};

static bool rd(FILE* f, void* p, size_t n) {
  return fread(p, 1, n, f) == n;
}

int main(int argc, char** argv) {
  if (argc != 4)
    return fprintf(stderr, "usage: %s encode|decode <in> <out>\n", argv[0]), 2;
  FILE* in  = fopen(argv[2], "rb");
  FILE* out = fopen(argv[3], "wb");
  if (!in || !out)
    return fprintf(stderr, "cannot open files\n"), 2;
  if (!strcmp(argv[1], "encode")) {
    uint32_t n = 0;
    rd(in, &n, 4);
    std::map<ChunkKey, std::unique_ptr<ChunkDesc>> grid;
    std::vector<ChunkKey> order;
    for (uint32_t i = 0; i < n; ++i) {
      ChunkKey k;
      SDFBlockDesc d{};
      uint32_t nv = 0;
      rd(in, k.c, 12), rd(in, &d.pos, 12), rd(in, &d.ptr, 4), rd(in, &d.resolution, 4), rd(in, &nv, 4);
      SDFBlock b;
      b.data.reserve(nv); // streamer.cuh:24-26
      for (uint32_t j = 0; j < nv; ++j) {
        Voxel v;
        rd(in, &v, 12);
        b.data.push_back(v); // streamer.cpp:225-228
      }
      if (!grid.count(k)) {
        grid[k] = std::make_unique<ChunkDesc>();
        order.push_back(k);
      }
      grid[k]->vecChunkDesc_.push_back(d); // ChunkDesc::addSDFBlock, streamer.cuh:123-126
      grid[k]->vecSDFBlock_.push_back(b);
    }
    for (const ChunkKey& k : order) { // serializer.h:26-33
      auto buffer         = cista::serialize(*grid[k]);
      uint64_t chunk_size = buffer.size();
      fwrite(&chunk_size, 8, 1, out);
      fwrite(k.c, 12, 1, out);
      fwrite(buffer.data(), 1, buffer.size(), out);
    }
  } else {
    std::vector<uint8_t> buffer;
    uint64_t chunk_size = 0;
    std::vector<uint8_t> blocks;
    uint32_t n = 0;
    auto put   = [&](const void* p, size_t k) { blocks.insert(blocks.end(), (const uint8_t*) p, (const uint8_t*) p + k); };
    while (rd(in, &chunk_size, 8)) { // serializer.h:55-72
      ChunkKey k;
      buffer.resize(chunk_size);
      rd(in, k.c, 12), rd(in, buffer.data(), chunk_size);
      ChunkDesc* c = cista::deserialize<ChunkDesc>(buffer);
      for (uint32_t i = 0; i < c->vecSDFBlock_.size(); ++i) {
        const SDFBlockDesc& d = c->vecChunkDesc_[i];
        const uint32_t nv     = (uint32_t) c->vecSDFBlock_[i].data.size();
        put(k.c, 12), put(&d.pos, 12), put(&d.ptr, 4), put(&d.resolution, 4), put(&nv, 4);
        put(c->vecSDFBlock_[i].data.data(), 12 * (size_t) nv);
        ++n;
      }
    }
    fwrite(&n, 4, 1, out);
    fwrite(blocks.data(), 1, blocks.size(), out);
  }
  fclose(in), fclose(out);
  return 0;
}

// mrh_host.h — host-side state of one map handle (shared by the translation units of libmrhash_b200).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/mrhash_b200.h"
#include "mrh_types.cuh"

namespace mrh {

  // Voxel payload words on the host. The allocator default-initialises (no zero fill): these vectors
  // are hundreds of MB and every word is overwritten by the copy that follows the resize, so the
  // memset - and taking all the page faults on one thread - would cost more than the transfer.
  template <typename T>
  struct NoInitAlloc : std::allocator<T> {
    template <typename U>
    struct rebind {
      using other = NoInitAlloc<U>;
    };
    NoInitAlloc() = default;
    template <typename U>
    NoInitAlloc(const NoInitAlloc<U>&) {
    }
    template <typename U>
    void construct(U* p) {
      ::new ((void*) p) U;
    }
    template <typename U, typename... Args>
    void construct(U* p, Args&&... args) {
      ::new ((void*) p) U(std::forward<Args>(args)...);
    }
  };
  using VoxelWords = std::vector<uint32_t, NoInitAlloc<uint32_t>>;

  // Host store of streamed-out blocks: the role of Streamer::grid_ (streamer.cuh:354, the
  // unordered_map of 1 m chunks filled by integrateInChunkGrid, streamer.cpp:216-247). Records are
  // kept dense; the chunk of a record is derived on demand (worldToChunks of the block origin).
  struct HostStore {
    std::vector<GatherRecord> recs;
    VoxelWords voxels; // 512 x {sdf bits, sum_squared bits, rgbw} per record (reference Voxel layout)
    void clear() {
      recs.clear();
      voxels.clear();
      recs.shrink_to_fit();
      voxels.shrink_to_fit();
    }
    size_t size() const {
      return recs.size();
    }
  };

  struct HostMesh {
    std::vector<float> triangles; // host copy of the raw soup (18 floats per triangle), fetched on demand
    std::vector<double, NoInitAlloc<double>> vertices; // V x 3
    std::vector<int32_t, NoInitAlloc<int32_t>> faces;  // F x 3
    std::vector<double, NoInitAlloc<double>> colors;   // V x 3
    void clear() {
      triangles.clear(), vertices.clear(), faces.clear(), colors.clear();
    }
  };

  int fail(const char* fmt, ...);

  // one double-buffered host->device channel (depth, colour or points)
  struct Ingest {
    void* d_buf[2]{};
    size_t d_cap[2]{};
    void* h_buf[2]{}; // pinned staging, used when the caller's memory is pageable
    size_t h_cap[2]{};
    cudaEvent_t copied[2]{};   // on the copy stream, after the H2D into d_buf[w]
    cudaEvent_t consumed[2]{}; // on the compute stream, after the last frame that read d_buf[w]
    int which   = 0;
    bool active = false; // d_buf[which] holds the current frame's data
    bool pending_direct = false; // a transfer straight from the caller's pinned memory is in flight
    cudaStream_t stream = nullptr; // the copy stream this input travels on (depth / points: 0, colour / normals: 1)
  };

} // namespace mrh

struct mrh_map {
  mrh_params p{};
  int device  = 0;
  int num_sms = 148;
  int integrate_grid = 148 * 8; // CTAs of k_integrate (env MRH_INTEGRATE_CTAS_PER_SM overrides, for tuning)
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  mrh::MapDev dev{};
  mrh::CameraDev cam{};
  float pose[16]{};
  float cam_in_lidar[16]{};
  float max_integration_distance = 0.f;
  uint64_t num_sdf_blocks = 0, hash_num_buckets = 0, max_num_triangles = 0, max_stream_blocks = 0;
  size_t zbuf_cap = 0;

  // ingest: copies run on their own stream into double-buffered device images, so the transfer of
  // frame k+1 overlaps the kernels of frame k (mrh_capi.cu: struct use in ingest_upload)
  // MRH_TIMELINE=n (tuning): timing events around the uploads and the frame of the first n frames after
  // the switch, printed by mrh_debug_timeline (device-side timeline of the streaming path)
  struct TimelineFrame {
    cudaEvent_t up0[2]{}, up1[2]{}, k0 = nullptr, k1 = nullptr;
  };
  std::vector<TimelineFrame> timeline;
  size_t timeline_next = 0, timeline_up_next[2] = {0, 0};
  cudaStream_t copy_stream = nullptr;  // H2D of depth images / points
  cudaStream_t copy_stream2 = nullptr; // H2D of colour images / normals: the two uploads of a frame overlap their fixed costs
  mrh::Ingest in_depth, in_rgb, in_points, in_normals;
  const float* depth_ptr = nullptr;
  const uint8_t* rgb_ptr = nullptr;
  cudaEvent_t rgb_ready  = nullptr; // set by compute(): the colour image is only needed by the fusion kernels
  int depth_rows = 0, depth_cols = 0, rgb_rows = 0, rgb_cols = 0;
  size_t n_points = 0;
  float* d_points = nullptr;
  float* d_normals = nullptr; // per-point normals of the current cloud (nullptr: none given)

  // sharded starve frames: the frame stops after the z-buffer pass (mrh_compute_begin) so that the
  // caller can min-reduce the z-buffer over the ranks, and resumes in mrh_compute_end
  bool split_zbuf = false, pending_gc = false, pending_var = false;
  mrh::FrameDev pending_f{};

  uint32_t frame_index = 0; // num_integrated_frames_ (voxel_data_structures.cpp:106)
  uint32_t live_cur    = 0;
  // fused frame kernel (mrh_fused.cuh); MRH_FRAME=split selects the two-launch frame of round 1 for A/B runs
  bool use_fused = true, use_bulk_depth = true;
  int fused_grid = 0, fused_ctas_per_sm = 0;
  int fused_pref_num = 0, fused_pref_den = 4; // CTAs in slots < num of every den prefer fusion items over ray tiles
  uint32_t fuse_tag = 0;                      // tag of the last frame's fusion-queue entries (never reset)
  float shortcut_size = 0.f, shortcut_ext = 0.f; // parameters shortcut_radius was verified for
  int shortcut_radius = 0;
  bool use_pdl         = true; // programmatic dependent launch of the two frame kernels (env MRH_PDL=0 switches it off)
  bool counters_clean  = true; // live_count[live_cur ^ 1] and vis_count are already zero (fast RGB-D path precondition)
  uint64_t frames_total = 0;
  uint64_t launches     = 0;
  uint64_t h2d_bytes    = 0;
  mrh::Counters* h_ctr  = nullptr; // pinned read-back
  // how page-locked caller memory is ingested (mrh_set_ingest_mode): 0 = staged like pageable memory
  // (the setter copies, reference semantics), 1 = DMA from the caller's buffer, awaited at the end of
  // compute(), 2 = DMA from the caller's buffer, never awaited by compute() (streaming callers)
  int ingest_mode = 0;
  // pipelined statistics (mrh_set_stats_pipeline): every compute() ends with k_snapshot_counters writing
  // the counters into one of kCtrRing pinned (mapped) slots; mrh_get_stats_pipelined reads an earlier frame's
  static constexpr int kCtrRing = 4;
  mrh::Counters* h_ctr_ring = nullptr; // kCtrRing pinned slots
  cudaEvent_t ev_ctr[kCtrRing]{};
  uint64_t ctr_frames[kCtrRing]{};     // frames_total at the time of the copy
  bool stats_pipeline = false;
  int ctr_slot        = 0;             // slot of the most recent compute()
  int ctr_filled      = 0;             // copies issued so far (saturates at kCtrRing)

  // optional per-kernel timing (bench.py roofline pass): events between the kernels of a frame
  bool profiling = false;
  cudaEvent_t ev_k[8]{};
  double kernel_ms[8]{};
  uint64_t kernel_launches[8]{};

  // point-cloud path staging: voxel-address keys (u32 or u64) + sdf values, `upd_slots` per point,
  // double buffered for the sort
  void* d_upd_keys[2]{};
  float* d_upd_vals[2]{};
  size_t upd_cap       = 0;
  uint32_t upd_slots   = 0;
  void* d_sort_tmp     = nullptr;
  size_t sort_tmp_bytes = 0;

  // meshing
  float* d_tri           = nullptr;
  size_t d_tri_cap       = 0;
  uint32_t* d_tri_count  = nullptr;
  size_t soup_in_tri     = 0;       // triangles of the last marching-cubes run, still in d_tri
  float* d_soup_acc      = nullptr; // soups of earlier regions of the same extraction
  size_t soup_acc_n = 0, soup_acc_cap = 0;
  // sharded meshing: ghost copies of other ranks' blocks sit behind entry halo_owned of the live list
  bool halo_active       = false;
  uint32_t halo_owned    = 0;
  uint16_t* d_shell_idx  = nullptr;
  // radius paging (Streamer::stream): trigger when free pool blocks <= stream_threshold * num_sdf_blocks
  // (params.h:28); the last kernel of a frame writes {free-stack top, frame tag} into pinned host memory
  float stream_threshold  = 0.15f;
  int* h_heap_probe       = nullptr; // pinned, 2 ints
  int frames_probe_seen   = -1;
  uint64_t stream_events = 0, last_stream_out = 0, last_stream_in = 0, stream_duplicates = 0;
  // pinned bounce buffers (2 x kBounceBytes) for bulk transfers between the device and pageable host vectors
  void* h_bounce[2]{};
  cudaEvent_t ev_bounce[2]{};
  mrh::HostStore store;
  mrh::HostMesh mesh;
  // wall-clock breakdown of the last extractMesh (ms): stream in/out, marching-cubes kernel, device weld + D2H of the mesh, PLY
  double mesh_ms_stream = 0, mesh_ms_kernel = 0, mesh_ms_merge = 0, mesh_ms_ply = 0;
};

namespace mrh {
  int reset_map(mrh_map* m);
  // far_centre != nullptr: only blocks at least far_radius away, which also leave the device map
  int gather_to_host(mrh_map* m, std::vector<GatherRecord>& recs, VoxelWords& voxels, const float* far_centre = nullptr, float far_radius = 0.f);
  // Streamer::stream (streamer.cpp:337-355): page far blocks out, page the host blocks around centre in
  int stream_radius(mrh_map* m, const float centre[3], float radius);
  // chunk test of Streamer::isChunkInSphere for the record's chunk (streamer.cuh:346-352)
  bool record_in_sphere(const mrh_map* m, const GatherRecord& r, const float centre[3], float radius);
  int insert_from_host(mrh_map* m, const GatherRecord* recs, const uint32_t* voxels, size_t n);
  // bulk copies through the pinned bounce buffers: DMA and parallel host copy overlap piece by piece
  int bulk_d2h(mrh_map* m, void* dst_host, const void* src_dev, size_t bytes);
  int bulk_h2d(mrh_map* m, void* dst_dev, const void* src_host, size_t bytes);
  void parallel_copy(void* dst, const void* src, size_t bytes);
  int carve_low_blocks(mrh_map* m, uint32_t n_low);
  int weld_on_device(mrh_map* m, const float* d_soup, size_t n_tri, double eps);
  int integrate_rgbd(mrh_map* m);
  int integrate_points(mrh_map* m);
  int finish_gc_tail(mrh_map* m);
  FrameDev make_frame(const mrh_map* m);
  int integrate_ctas_per_sm(); // resident CTAs per SM of k_integrate as compiled (mrh_frame.cu)
  int fused_grid(mrh_map* m);  // resident CTAs of k_frame on the whole device

} // namespace mrh

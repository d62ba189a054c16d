// render.h -- the renderer object behind vnr_renderer_t (MainRenderer, renderer.h:55-235)
#pragma once
#include <memory>
#include <vector>

#include "volume.h"

namespace vnr {

struct FrameParams;
struct RendererComm;
struct RayBuffers;
constexpr int kMaxRounds = 1024;
constexpr int kMaxFramesInFlight = 8;

// everything baked into the kernel nodes of the captured wavefront loop
struct GraphKey {
  DecoderDesc desc;
  const void* params;
  const void* ptrs[16];
  const void* volume_src;               // non-null: trilinear volume lookup instead of the network decode
  unsigned grid; size_t cap; int rounds; int shade;
};

// Everything ONE frame in flight needs: its stream, ray / sample / value buffers, counters, accumulation and frame buffers,
// the captured wavefront graphs (their kernel nodes bake these pointers) and the pinned host frames.  A renderer owns a ring
// of them (MainRenderer double-buffers its framebuffer the same way, renderer.h:84-94, framebuffer.h:73-77, but renders one
// frame at a time; here consecutive vnr_render calls land in consecutive slots and overlap on the device).
struct FrameSlot {
  cudaStream_t stream = nullptr;
  cudaEvent_t frame_done[2] = {nullptr, nullptr};
  cudaEvent_t vol_ready = nullptr;
  int cur = 0, last_rounds = 1, last_passes = 1;
  bool rendered = false, downloaded = false, mapped = true;
  cudaStream_t vol_waited_on = nullptr; bool vol_waited = false;   // the volume already ordered a stream after this frame
  int frame_index = 0;                  // the frame_index this slot rendered last

  DevBuf<float4> accum, frame, samples[2], ray_rgba, ray_tn, ssh_org, ssh_col, ssh_rgba;
  DevBuf<int4> ray_cell;
  DevBuf<float> values, ray_jitter, ssh_jitter;
  DevBuf<uint32_t> ray_state, counters;
  // path tracer (pathtrace.cuh): per-ray state, live-ray lists (ping-pong), its own loop graph
  DevBuf<float4> pt_org, pt_dir, pt_rad, pt_thr, pt_tn; DevBuf<int4> pt_cell; DevBuf<uint32_t> pt_list[2];
  cudaGraph_t pt_graph = nullptr; cudaGraphExec_t pt_exec = nullptr; GraphKey pt_key; bool last_pt = false;
  float4* h_frame[2] = {nullptr, nullptr};
  bool h_frame_external = false;        // the host frames belong to a communicator (shared by all ranks), not to the slot
  DevBuf<uint32_t> host_nonzero[2]; bool host_nonzero_valid[2] = {false, false};   // FrameParams::host_nonzero of each host frame
  int map_idx = 0;                      // host frame the last render of this slot wrote (`cur` = the one the next render writes)
  uint32_t* h_counters = nullptr;
  std::vector<cudaEvent_t> prof_events;
  int prof_used = 0;
  uint64_t launches = 0;                // kernels launched by the last render()
  DevBuf<uint8_t> fp_dev;               // FrameParams of the frame in flight (device copy)
  bool last_graph = false;
  cudaGraph_t loop_graph[2] = {nullptr, nullptr}; cudaGraphExec_t loop_exec[2] = {nullptr, nullptr}; cudaStream_t capture_stream = nullptr;
  GraphKey graph_key[2];                // [1]: the shadow pass of the single-shade heuristic
  // multi-GPU: finished pixels are stored here instead of `frame` (rank 0's frame buffer of the same slot, peer-mapped)
  float4* frame_target = nullptr;
  float4* frame_out() { return frame_target ? frame_target : frame.p; }

  FrameSlot();
  ~FrameSlot();
  void resize(size_t npix);
  void destroy_graph();
};

struct Renderer {
  Volume* vol;
  int width = 0, height = 0;
  int mode = 5;                         // vnrCreateRenderer default (api.cpp:456)
  bool gt_source = false;               // march the ground-truth volume (SimpleVolume renderer)
  int n_iters = 16;                     // N_ITERS (method_raymarching.cu:30-40); up to 32 for unshaded marching (fewer, fuller rounds)
  int jitter_mode = 0;
  bool tiled = true;                    // warps own 8 x 4 pixel tiles instead of 32-pixel scanline segments
  bool transpose = true;                // sample slots of a warp are depth-major (all rays' j-th samples adjacent)
  int part_rank = 0, part_world = 1; uint32_t strip_rows = 4;
  float cam_from[3] = {0, 0, -1}, cam_at[3] = {0, 0, 0}, cam_up[3] = {0, 1, 0}, fovy = 60.f;   // instantvnr_types.h:74-83
  float sampling_rate = 1.f, density_scale = 1.f;
  float scale[3] = {1, 1, 1};
  float clip_lo[3] = {0, 0, 0}, clip_hi[3] = {1, 1, 1};
  int frame_index = 0; bool reset = true;
  float light_dir[3] = {0.7f, 0.9f, 0.4f};                                // light_directional_dir, instantvnr_types.h:148 (persistent sign flips)
  bool download = true;                 // framebuffer_skip_download (renderer.cpp:132)
  bool zero_copy = true;                // finished pixels go straight to the pinned host frame (no D2H copy after the frame)
  bool profiling = false;               // CUDA events around every decode launch
  // device-driven wavefront loop (CUDA graph with a WHILE node); off: bounded host-enqueued rounds
  bool use_graph = true;

  // the frame ring: vnr_render fills slot n_rendered % n_slots, vnr_map_frame returns the oldest rendered frame not yet mapped
  std::vector<std::unique_ptr<FrameSlot>> slots;
  uint64_t n_rendered = 0, n_mapped = 0;
  int last_slot = 0;                    // slot of the most recent vnr_render (stats / profile / device frame refer to it)
  FrameSlot& slot(int k) { return *slots[(size_t)k]; }
  FrameSlot& last() { return *slots[(size_t)last_slot]; }
  RendererComm* rcomm = nullptr;        // communicator attachment (comm.h): tile-parallel rendering through vnr_render

  explicit Renderer(Volume* v);
  ~Renderer();
  void set_frames_in_flight(int n);
  uint32_t local_rays() const;
  void resize(int w, int h);
  void fill_frame_params(FrameParams& fp);
  int round_bound(int iters) const;
  void sync_all();
  void ensure_graph(FrameSlot& S, int pass, int shade, const RayBuffers& rb, unsigned grid, size_t cap, int rounds, const float* volume_src);
  void render();
  void render_pathtracing(FrameSlot& S, const float* volume_src, unsigned grid, size_t cap, bool graph_loop);
  void download_now();
  const float* map_frame();
  void stats(uint64_t* s4);
  void profile(float* decode_ms, int* decode_launches);
};

}  // namespace vnr

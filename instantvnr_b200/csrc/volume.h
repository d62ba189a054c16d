// volume.h -- the neural volume and renderer objects behind the opaque C handles.
#pragma once
#include "vnr_host.h"

namespace vnr {

template <typename T>
struct DevBuf {
  T* p = nullptr; size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  static size_t& total() { static size_t t = 0; return t; }
  void alloc(size_t count) {
    if (count == n && p) return;
    release();
    if (count) { VNR_CUDA(cudaMalloc((void**)&p, count * sizeof(T))); n = count; total() += count * sizeof(T); }
  }
  void ensure(size_t count) { if (count > n) alloc(count); }
  void zero(cudaStream_t s = 0) { if (p) VNR_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void release() { if (p) { cudaFree(p); total() -= n * sizeof(T); } p = nullptr; n = 0; }
  size_t bytes() const { return n * sizeof(T); }
};

struct SlabSampler;
struct VolumeComm;

struct Volume {
  ModelConfig cfg;
  int dims[3] = {0, 0, 0};
  int device = 0;                  // the CUDA device the volume lives on (current device at creation)
  cudaStream_t stream = nullptr;

  // parameters: fp16 working copy (MLP matrices, then grid), fp32 master, gradients, Adam state
  DevBuf<__half> params;
  DevBuf<float> master, m1, m2;
  DevBuf<__half> grid_grads;       // fp16 [n_grid], loss-scaled (x128); cleared by the optimizer sweep
  DevBuf<float> mlp_grads;         // fp32 [n_mlp], loss-scaled, reduced over CTAs
  DevBuf<float> mlp_partial;       // fp32 [n_cta][n_mlp] per-CTA weight-gradient partials
  bool grads_clean = false, grads_pending = false;
  DevBuf<uint32_t> steps; bool steps16 = false;   // per-parameter Adam step counters: 32-bit, or 16-bit saturating packed two per word (train.cu AdamArgs)
  DevBuf<float> bias_tab; uint32_t bias_filled = 0;   // Adam bias-correction table (train.cu) ...
  std::vector<float*> bias_retired;                   // outgrown tables still referenced by kernels in flight
  float bias_beta1 = -1.f, bias_beta2 = -1.f;         // ... and the betas it was built with (refilled when the optimizer config changes)
  // measurement taps of the training kernel (vnr_volume_train_debug): chain variant, role switches, per-CTA role timers
  int train_variant = 2; uint32_t train_flags = 0; bool train_prof_on = false; DevBuf<uint32_t> train_prof;
  bool have_params = false, have_opt = false;
  uint32_t opt_step = 0; float lr_factor = 1.f;
  uint64_t train_step = 0;
  DevBuf<double> loss_accum;       // [0] running sum of per-step losses, [1] last step
  uint64_t loss_count = 0;

  // ground truth + sampler
  DevBuf<float> gt;                // linear [z][y][x]
  bool have_gt = false;
  Pcg32 sampler_rng;               // neural_sampler.cu:36  `static default_rng_t rng{1337}`
  DevBuf<float> train_x, train_y;
  // second batch buffer + side stream of train_steps: the next batch is drawn, the MLP's optimizer step and the macrocell update
  // run while the hash-grid optimizer sweep (HBM-bound) occupies the main stream
  DevBuf<float> train_x2, train_y2;
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  SlabSampler* ooc = nullptr;      // out-of-core sampler (slab_sampler.cu): training draws from a pool of random file slabs

  // progressively decoded volume (network.cu:290-326): what the "decoding" rendering modes march
  DevBuf<float> decoded; int decode_blob = 0;

  // macrocell (core/macrocell.h): value range (offset by -1/+1) and max opacity per cell
  int mc_dims[3] = {0, 0, 0};
  DevBuf<float> mc_range, mc_maxop;
  bool mc_external = false;        // value ranges came from the ground truth (MacroCell::set_external)

  // transfer function (object.cpp:321-348)
  DevBuf<float4> tfn_color; DevBuf<float> tfn_alpha;
  int n_color = 0, n_alpha = 0; float tfn_lo = 0.f, tfn_hi = 1.f;

  // data-parallel peers (train.cu dp_optimizer_step): every rank's parameter / gradient buffers, own ones included
  int dp_rank = 0, dp_world = 0;
  void* dp_params[kMaxPeers] = {}; void* dp_grid_grads[kMaxPeers] = {}; void* dp_mlp_grads[kMaxPeers] = {};

  // communicator attachment (comm.h): data-parallel training through vnr_volume_train
  VolumeComm* vcomm = nullptr;

  // renderers created on this volume (they register themselves): what mutating entry points order against
  std::vector<Renderer*> renderers;

  std::string blob;                // last serialized params.json
  std::string peek_json;

  Volume();
  ~Volume();
  size_t cells() const { return (size_t)mc_dims[0] * mc_dims[1] * mc_dims[2]; }
};

// Ordering of volume-side writes against frames in flight: work enqueued on `s` after this call starts after the last
// frame of every slot of every renderer of the volume has finished (render.cu).  The frame side already waits for the
// volume's pending work (vol_ready), so train / render / train sequences are ordered in both directions without a host sync.
void wait_for_frames(Volume* v, cudaStream_t s);

}  // namespace vnr

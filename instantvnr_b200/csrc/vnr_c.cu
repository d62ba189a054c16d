// vnr_c.cu -- implementation of the C ABI declared in include/vnr_c.h.
// Exceptions never cross the boundary: every entry point funnels through guard().
#include <cmath>
#include <cstring>
#include <random>

#include "../../include/vnr_c.h"
#include "mini_json.h"
#include "volume.h"
#include "comm.h"
#include "render.h"
#include "train.h"
#include "scene.h"

using namespace vnr;

#define VNR_EXPORT extern "C" __attribute__((visibility("default")))

static thread_local std::string g_last_error;
// one process driving several devices (vnr_comm_init(n > 1)): every entry point makes the object's device current
static bool g_multi_device = false;

template <typename Fn>
static int guard(Fn&& fn) {
  try { fn(); return VNR_OK; }
  catch (const CudaError& e) { g_last_error = e.what(); cudaGetLastError(); return VNR_ERR_CUDA; }
  catch (const InvalidError& e) { g_last_error = e.what(); return VNR_ERR_INVALID; }
  catch (const UnsupportedError& e) { g_last_error = e.what(); return VNR_ERR_UNSUPPORTED; }
  catch (const StateError& e) { g_last_error = e.what(); return VNR_ERR_STATE; }
  catch (const std::exception& e) { g_last_error = e.what(); return VNR_ERR_INVALID; }
}

static void require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { cudaGetLastError(); throw CudaError("no CUDA device available: this library has no CPU fallback"); }
}

Volume::Volume() { sampler_rng.seed(1337); }
Volume::~Volume() {
  if (g_multi_device) cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  if (vcomm) { try { comm_detach_volume(this); } catch (...) {} }
  outofcore_release(this);
  for (float* p : bias_retired) cudaFree(p);
  for (int r = 0; r < dp_world; ++r) {          // mappings of the peers' buffers (vnr_volume_dp_attach)
    if (r == dp_rank) continue;
    if (dp_params[r]) cudaIpcCloseMemHandle(dp_params[r]);
    if (dp_grid_grads[r]) cudaIpcCloseMemHandle(dp_grid_grads[r]);
    if (dp_mlp_grads[r]) cudaIpcCloseMemHandle(dp_mlp_grads[r]);
  }
  if (side) cudaStreamDestroy(side);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_join) cudaEventDestroy(ev_join);
  if (stream) cudaStreamDestroy(stream);
}

static Volume* V(vnr_volume_t* v) {
  if (!v) throw InvalidError("null volume handle");
  Volume* p = reinterpret_cast<Volume*>(v);
  if (g_multi_device) VNR_CUDA(cudaSetDevice(p->device));
  return p;
}
static const Volume* V(const vnr_volume_t* v) {
  if (!v) throw InvalidError("null volume handle");
  const Volume* p = reinterpret_cast<const Volume*>(v);
  if (g_multi_device) VNR_CUDA(cudaSetDevice(p->device));
  return p;
}
static cudaStream_t S(Volume* v, void* stream) { return stream ? (cudaStream_t)stream : v->stream; }

VNR_EXPORT const char* vnr_last_error(void) { return g_last_error.c_str(); }

VNR_EXPORT int vnr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

VNR_EXPORT int vnr_volume_create(const char* model_json, int dx, int dy, int dz, vnr_volume_t** out) {
  return guard([&] {
    if (!model_json || !out) throw InvalidError("null argument");
    if (dx <= 0 || dy <= 0 || dz <= 0) throw InvalidError("volume dims must be positive");
    ModelConfig cfg = parse_model_config(model_json);
    require_device();
    std::unique_ptr<Volume> v(new Volume());
    if (const char* e = getenv("VNR_TRAIN_VARIANT")) { const int t = atoi(e); if (t >= 0 && t <= 2) v->train_variant = t; }   // A/B runs of the training kernel
    v->cfg = cfg;
    v->dims[0] = dx; v->dims[1] = dy; v->dims[2] = dz;
    VNR_CUDA(cudaGetDevice(&v->device));
    VNR_CUDA(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
    v->params.alloc(cfg.n_params());
    v->params.zero(v->stream);
    // MacroCell::set_shape + allocate (macrocell.cu:195-219): 16^3 voxels per cell, zero-initialised ranges
    for (int k = 0; k < 3; ++k) v->mc_dims[k] = (v->dims[k] + 15) / 16;
    v->mc_range.alloc(2 * v->cells()); v->mc_range.zero(v->stream);
    v->mc_maxop.alloc(v->cells()); v->mc_maxop.zero(v->stream);
    v->loss_accum.alloc(2); v->loss_accum.zero(v->stream);
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    *out = reinterpret_cast<vnr_volume_t*>(v.release());
  });
}

VNR_EXPORT void vnr_volume_release(vnr_volume_t* v) { delete reinterpret_cast<Volume*>(v); }

VNR_EXPORT int vnr_volume_model_info(const vnr_volume_t* v, uint64_t* n_params, uint64_t* n_mlp_params, int* n_levels, int* n_feat, int* n_hidden) {
  return guard([&] {
    const Volume* vol = V(v);
    if (n_params) *n_params = vol->cfg.n_params();
    if (n_mlp_params) *n_mlp_params = vol->cfg.desc.n_mlp;
    if (n_levels) *n_levels = vol->cfg.n_levels;
    if (n_feat) *n_feat = vol->cfg.n_feat;
    if (n_hidden) *n_hidden = vol->cfg.n_hidden;
  });
}

static void upload_master_from_f16(Volume* v, const std::vector<__half>& h) {
  std::vector<float> f(h.size());
  for (size_t i = 0; i < h.size(); ++i) f[i] = __half2float(h[i]);
  v->master.alloc(f.size());
  VNR_CUDA(cudaMemcpy(v->master.p, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice));
}

VNR_EXPORT int vnr_volume_init_params(vnr_volume_t* vh, uint32_t seed) {
  return guard([&] {
    Volume* v = V(vh);
    const DecoderDesc& d = v->cfg.desc;
    const size_t n = v->cfg.n_params();
    std::vector<float> p(n);
    // Trainer ctor (trainer.h:54-60): pcg32 seeded with seed_seq{seed}.generate()[0]
    std::seed_seq seq{seed};
    std::vector<uint32_t> seeds(2);
    seq.generate(seeds.begin(), seeds.end());
    Pcg32 rng; rng.seed((uint64_t)seeds.front());
    size_t pos = 0;
    auto xavier = [&](int rows, int cols) {       // gpu_matrix.h:197-211
      const float scale = std::sqrt(6.0f / (float)(rows + cols));
      for (size_t i = 0; i < (size_t)rows * cols; ++i) p[pos + i] = rng.next_float() * 2.0f * scale - scale;
      pos += (size_t)rows * cols;
    };
    xavier(kWidth, d.enc_pad);
    for (int i = 0; i < d.n_hidden - 1; ++i) xavier(kWidth, kWidth);
    xavier(kOutPad, kWidth);
    // grid: generate_random_uniform(-1e-4, 1e-4) in device order (random.h:67-98): thread i takes
    // stream elements 4i..4i+3 and writes them to i + n_threads*j.
    {
      const size_t ng = d.n_grid, need = (ng + 3) / 4, n_threads = ((need + 127) / 128) * 128;
      const float lower = -1e-4f, upper = 1e-4f;
      for (size_t i = 0; i < n_threads; ++i)
        for (size_t j = 0; j < 4; ++j) {
          const float u = rng.next_float();
          const size_t idx = i + n_threads * j;
          if (idx < ng) p[pos + idx] = fmaf(u, (upper - lower), lower);
        }
    }
    std::vector<__half> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = __float2half_rn(p[i]);
    v->master.alloc(n);
    wait_for_frames(v, v->stream);
    VNR_CUDA(cudaStreamSynchronize(v->stream));            // in-flight training kernels still write params / master, frames read params
    VNR_CUDA(cudaMemcpy(v->master.p, p.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    VNR_CUDA(cudaMemcpy(v->params.p, h.data(), n * sizeof(__half), cudaMemcpyHostToDevice));
    v->have_params = true;
    reset_optimizer_state(v);
  });
}

// vnrNeuralVolumeSetModel (api.h:126; NeuralVolume::set_network_from_json(config) core/network.cu:731-741 ->
// TcnnNetwork::deserialize_model): a new network, optimizer state and freshly initialised parameters under the same
// dims, ground truth, sampler stream, macrocells and transfer function.
VNR_EXPORT int vnr_volume_set_model(vnr_volume_t* vh, const char* model_json, uint32_t seed) {
  int rc = guard([&] {
    Volume* v = V(vh);
    if (!model_json) throw InvalidError("null argument");
    if (v->dp_world) throw StateError("detach the data-parallel peers before changing the model");
    ModelConfig cfg = parse_model_config(model_json);       // throws before anything is touched
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    VNR_CUDA(cudaDeviceSynchronize());                      // renderers of this volume may still be decoding the old table
    v->cfg = cfg;
    v->master.release(); v->m1.release(); v->m2.release(); v->steps.release();
    v->grid_grads.release(); v->mlp_grads.release(); v->mlp_partial.release();
    v->have_params = false; v->have_opt = false; v->grads_clean = false; v->grads_pending = false;
    v->decode_blob = 0;
    v->params.alloc(cfg.n_params());
    v->params.zero(v->stream);
    VNR_CUDA(cudaStreamSynchronize(v->stream));
  });
  return rc != VNR_OK ? rc : vnr_volume_init_params(vh, seed);
}

VNR_EXPORT int vnr_volume_set_params_f16(vnr_volume_t* vh, const uint16_t* h_params, size_t n) {
  return guard([&] {
    Volume* v = V(vh);
    if (!h_params) throw InvalidError("null params");
    if (n != v->cfg.n_params()) throw InvalidError("Can't set params because CPU buffer has the wrong size.");   // trainer.h:283
    std::vector<__half> h(n);
    memcpy(h.data(), h_params, n * 2);
    wait_for_frames(v, v->stream);
    VNR_CUDA(cudaStreamSynchronize(v->stream));            // the copies below run on the legacy stream, which does not order against v->stream
    VNR_CUDA(cudaMemcpy(v->params.p, h.data(), n * 2, cudaMemcpyHostToDevice));
    upload_master_from_f16(v, h);            // params_fp[i] = (float)params_inference[i]  trainer.h:289-291
    v->have_params = true;
    if (!v->have_opt) reset_optimizer_state(v);
  });
}

VNR_EXPORT int vnr_volume_get_params_f16(const vnr_volume_t* vh, uint16_t* h_params, size_t n) {
  return guard([&] {
    const Volume* v = V(vh);
    if (!h_params) throw InvalidError("null params");
    if (n != v->cfg.n_params()) throw InvalidError("parameter count mismatch");
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    VNR_CUDA(cudaMemcpy(h_params, v->params.p, n * 2, cudaMemcpyDeviceToHost));
  });
}

VNR_EXPORT int vnr_volume_decode(vnr_volume_t* vh, const float* d_xyz, float* d_out, size_t n, void* stream) {
  return guard([&] {
    Volume* v = V(vh);
    if (n && (!d_xyz || !d_out)) throw InvalidError("null buffer");
    if (n > 0xFFFFFF00ull) throw InvalidError("n too large");
    VNR_CUDA(launch_decode(v->cfg.desc, v->params.p, d_xyz, d_out, n, nullptr, S(v, stream)));
  });
}

VNR_EXPORT int vnr_volume_gather_probe(vnr_volume_t* vh, const float* d_xyz, uint32_t* d_out, size_t n, void* stream) {
  return guard([&] {
    Volume* v = V(vh);
    if (n && (!d_xyz || !d_out)) throw InvalidError("null buffer");
    VNR_CUDA(launch_gather_probe(v->cfg.desc, v->params.p, d_xyz, d_out, n, S(v, stream)));
  });
}

static void decode_host_impl(Volume* v, const float* h_xyz, float* h_out, uint16_t* h_enc, size_t n) {
  if (n == 0) return;
  if (!h_xyz || !h_out) throw InvalidError("null buffer");
  if (n > 0xFFFFFF00ull) throw InvalidError("n too large");
  DevBuf<float> x, y; DevBuf<__half> e;
  x.alloc(3 * n); y.alloc(n);
  if (h_enc) e.alloc(n * v->cfg.desc.enc_pad);
  VNR_CUDA(cudaMemcpyAsync(x.p, h_xyz, 3 * n * sizeof(float), cudaMemcpyHostToDevice, v->stream));
  VNR_CUDA(launch_decode(v->cfg.desc, v->params.p, x.p, y.p, n, e.p, v->stream));
  VNR_CUDA(cudaMemcpyAsync(h_out, y.p, n * sizeof(float), cudaMemcpyDeviceToHost, v->stream));
  if (h_enc) VNR_CUDA(cudaMemcpyAsync(h_enc, e.p, e.bytes(), cudaMemcpyDeviceToHost, v->stream));
  VNR_CUDA(cudaStreamSynchronize(v->stream));
}

VNR_EXPORT int vnr_volume_decode_host(vnr_volume_t* vh, const float* h_xyz, float* h_out, size_t n) {
  return guard([&] { decode_host_impl(V(vh), h_xyz, h_out, nullptr, n); });
}
VNR_EXPORT int vnr_volume_decode_debug(vnr_volume_t* vh, const float* h_xyz, float* h_out, uint16_t* h_enc, size_t n) {
  return guard([&] { decode_host_impl(V(vh), h_xyz, h_out, h_enc, n); });
}

VNR_EXPORT int vnr_memory_query(size_t* used_by_renderer, size_t* used_by_network) {
  return guard([&] {
    // single counter split by element type is not meaningful here; report everything the
    // library holds under "network" except float4 frame buffers.
    size_t fb = DevBuf<float4>::total();
    size_t all = DevBuf<float>::total() + DevBuf<__half>::total() + DevBuf<uint32_t>::total() + DevBuf<double>::total() + DevBuf<uint8_t>::total();
    if (used_by_renderer) *used_by_renderer = fb;
    if (used_by_network) *used_by_network = all;
  });
}

#include "vnr_c_volume.inl"
#include "vnr_c_render.inl"

// ---- scene descriptions (scene.cpp; serializer.cpp:138-477).  Host-only: no device is required. ----
static const Scene* SC(const vnr_scene_t* s) { if (!s) throw InvalidError("null scene handle"); return reinterpret_cast<const Scene*>(s); }
VNR_EXPORT int vnr_scene_create(const char* json, int is_path, vnr_scene_t** out) {
  return guard([&] {
    if (!json || !out) throw InvalidError("null argument");
    *out = nullptr;
    Scene* s = new Scene(is_path ? load_scene(json) : parse_scene(json));
    *out = reinterpret_cast<vnr_scene_t*>(s);
  });
}
VNR_EXPORT void vnr_scene_release(vnr_scene_t* s) { delete reinterpret_cast<Scene*>(s); }
VNR_EXPORT int vnr_scene_volume(const vnr_scene_t* sh, int* dims3, int* value_type, int* n_timesteps, float* range2, int* has_range) {
  return guard([&] {
    const Scene* s = SC(sh);
    if (dims3) { dims3[0] = s->dims[0]; dims3[1] = s->dims[1]; dims3[2] = s->dims[2]; }
    if (value_type) *value_type = s->value_type;
    if (n_timesteps) *n_timesteps = (int)s->files.size();
    if (range2) { range2[0] = s->range[0]; range2[1] = s->range[1]; }
    if (has_range) *has_range = s->has_range ? 1 : 0;
  });
}
VNR_EXPORT int vnr_scene_timestep(const vnr_scene_t* sh, int t, const char** filename, uint64_t* offset, int* big_endian) {
  return guard([&] {
    const Scene* s = SC(sh);
    if (t < 0 || t >= (int)s->files.size()) throw InvalidError("time step out of range");
    if (filename) *filename = s->files[t].filename.c_str();
    if (offset) *offset = s->files[t].offset;
    if (big_endian) *big_endian = s->files[t].big_endian ? 1 : 0;
  });
}
VNR_EXPORT int vnr_scene_camera(const vnr_scene_t* sh, float* from3, float* at3, float* up3, float* fovy) {
  return guard([&] {
    const Scene* s = SC(sh);
    if (!s->has_camera) throw StateError("the scene description has no camera");
    for (int k = 0; k < 3; ++k) { if (from3) from3[k] = s->cam_from[k]; if (at3) at3[k] = s->cam_at[k]; if (up3) up3[k] = s->cam_up[k]; }
    if (fovy) *fovy = s->fovy;
  });
}
VNR_EXPORT int vnr_scene_tfn(const vnr_scene_t* sh, const float** rgb, int* n_rgb, const float** alpha_pairs, int* n_alpha, float* range2, int* has_range) {
  return guard([&] {
    const Scene* s = SC(sh);
    if (range2) { range2[0] = s->range[0]; range2[1] = s->range[1]; }
    if (has_range) *has_range = s->has_range ? 1 : 0;
    if (rgb) *rgb = s->has_tfn ? s->tfn_color.data() : nullptr;
    if (n_rgb) *n_rgb = s->has_tfn ? (int)(s->tfn_color.size() / 3) : 0;
    if (alpha_pairs) *alpha_pairs = s->has_tfn ? s->tfn_alpha.data() : nullptr;
    if (n_alpha) *n_alpha = s->has_tfn ? (int)(s->tfn_alpha.size() / 2) : 0;
    if (s->tfn_present && !s->has_tfn)
      throw UnsupportedError("the scene's transferFunction is in the OVR tfn-module format (tfn::loadTransferFunction, not part of the reference "
                             "tree); set colours / alphas with the setters");
  });
}

// vnr_api.hpp -- C++ host-side mirror of the reference's public API (api.h) for the hot path,
// header-only, over the flat C ABI of libvnr_b200.so (vnr_c.h).
//
// Same function names, argument meaning and error behaviour as the reference's api.h (cited per
// function; paths relative to the reference repo): handles are std::shared_ptr, errors are
// std::runtime_error (api.cpp:129,138,215), a vnrJson that "is a string" is a FILE NAME
// (api.cpp:180-185).  Differences, all forced by what is (not) vendored in the reference:
//   * vnrJson is a small value type (JSON text | file name | BSON blob), not nlohmann::json;
//   * vnrCreateSimpleVolume takes an in-memory normalised float volume: the reference's scene-file
//     ingest (serializer.cpp, OVR volume readers) is outside the path;
//   * rendering modes other than 4/5/6 throw "unsupported".
// apps/vnr_cmd_train.cpp and apps/vnr_cmd_render.cpp are the reference's two headless drivers
// (apps/batch_trainer.cpp:72-141, apps/batch_renderer.cpp:156-239) written against this header.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "vnr_c.h"

namespace vnr {
struct vec2i { int x, y; vec2i(int x_ = 0, int y_ = 0) : x(x_), y(y_) {} };
struct vec3i { int x, y, z; vec3i(int x_ = 0, int y_ = 0, int z_ = 0) : x(x_), y(y_), z(z_) {} };
struct vec2f { float x, y; vec2f(float x_ = 0, float y_ = 0) : x(x_), y(y_) {} };
struct vec3f { float x, y, z; vec3f(float x_ = 0, float y_ = 0, float z_ = 0) : x(x_), y(y_), z(z_) {} explicit vec3f(float s) : x(s), y(s), z(s) {} };
struct vec4f { float x, y, z, w; };
struct range1f { float lo, hi; range1f(float l = 0.f, float h = 1.f) : lo(l), hi(h) {} };

// JSON argument of the api.h functions: inline text, a file name (is_string()), or a BSON blob
struct Json {
  enum Kind { Text, FileName, Binary } kind = Text;
  std::string data;
  Json() {}
  static Json text(const std::string& t) { Json j; j.kind = Text; j.data = t; return j; }
  static Json filename(const std::string& f) { Json j; j.kind = FileName; j.data = f; return j; }
  static Json binary(const std::string& b) { Json j; j.kind = Binary; j.data = b; return j; }
  bool is_string() const { return kind == FileName; }
};

inline std::string read_file(const std::string& name, bool binary) {
  std::ifstream f(name, binary ? std::ios::binary : std::ios::in);
  if (!f) throw std::runtime_error("cannot open " + name);
  std::stringstream ss; ss << f.rdbuf();
  return ss.str();
}
inline void check(int rc) { if (rc != VNR_OK) throw std::runtime_error(vnr_last_error()); }

struct Camera { vec3f from{0, 0, -1}, at{0, 0, 0}, up{0, 1, 0}; float fovy = 60.f; };              // instantvnr_types.h:74-83
struct TransferFunction { std::vector<vec3f> color; std::vector<vec2f> alpha; range1f range; };     // api_internal.h

struct VolumeContext {                                                                               // api_internal.h:17-39
  vec3i dims;
  vec3f clip_lo{0, 0, 0}, clip_hi{1, 1, 1};
  virtual bool isNetwork() const = 0;
  virtual ~VolumeContext() {}
};
struct SimpleVolumeContext : VolumeContext {
  std::vector<float> voxels;                        // normalised to [0,1], x fastest
  bool isNetwork() const override { return false; }
};
struct NeuralVolumeContext : VolumeContext {
  vnr_volume_t* h = nullptr;
  bool isNetwork() const override { return true; }
  ~NeuralVolumeContext() override { vnr_volume_release(h); }
};
struct RendererContext {                                                                             // api_internal.h:41-45
  std::shared_ptr<VolumeContext> volume;            // keeps the volume alive
  vnr_renderer_t* h = nullptr;
  vec2i size;
  ~RendererContext() { vnr_renderer_release(h); }
};
}  // namespace vnr

typedef std::shared_ptr<vnr::VolumeContext> vnrVolume;
typedef std::shared_ptr<vnr::RendererContext> vnrRenderer;
typedef std::shared_ptr<vnr::TransferFunction> vnrTransferFunction;
typedef std::shared_ptr<vnr::Camera> vnrCamera;
typedef vnr::Json vnrJson;

enum vnrRenderMode {                                                                                 // api.h:36-60
  VNR_OPTIX_NO_SHADING = 0, VNR_OPTIX_GRADIENT_SHADING, VNR_OPTIX_FULL_SHADOW, VNR_OPTIX_SINGLE_SHADE_HEURISTIC,
  VNR_RAYMARCHING_NO_SHADING_DECODING, VNR_RAYMARCHING_NO_SHADING_SAMPLE_STREAMING, VNR_RAYMARCHING_NO_SHADING_IN_SHADER,
  VNR_RAYMARCHING_GRADIENT_SHADING_DECODING, VNR_RAYMARCHING_GRADIENT_SHADING_SAMPLE_STREAMING, VNR_RAYMARCHING_GRADIENT_SHADING_IN_SHADER,
  VNR_RAYMARCHING_SINGLE_SHADE_HEURISTIC_DECODING, VNR_RAYMARCHING_SINGLE_SHADE_HEURISTIC_SAMPLE_STREAMING,
  VNR_RAYMARCHING_SINGLE_SHADE_HEURISTIC_IN_SHADER,
  VNR_PATHTRACING_DECODING, VNR_PATHTRACING_SAMPLE_STREAMING, VNR_PATHTRACING_IN_SHADER, VNR_INVALID,
};

// ---- json I/O (api.h:89-95) -----------------------------------------------------------------------
inline vnrJson vnrCreateJsonText(std::string filename) { return vnrJson::text(vnr::read_file(filename, false)); }
inline vnrJson vnrCreateJsonBinary(std::string filename) { return vnrJson::binary(vnr::read_file(filename, true)); }
inline void vnrLoadJsonText(vnrJson& j, std::string filename) { j = vnrCreateJsonText(filename); }
inline void vnrLoadJsonBinary(vnrJson& j, std::string filename) { j = vnrCreateJsonBinary(filename); }
inline void vnrSaveJsonBinary(const vnrJson& j, std::string filename) {
  std::ofstream f(filename, std::ios::binary);
  if (!f) throw std::runtime_error("cannot write " + filename);
  f.write(j.data.data(), (std::streamsize)j.data.size());
}

// ---- camera (api.h:102-110) -----------------------------------------------------------------------
inline vnrCamera vnrCreateCamera() { return std::make_shared<vnr::Camera>(); }
inline void vnrCameraSet(vnrCamera c, vnr::vec3f from, vnr::vec3f at, vnr::vec3f up) { c->from = from; c->at = at; c->up = up; }
inline vnr::vec3f vnrCameraGetPosition(vnrCamera c) { return c->from; }
inline vnr::vec3f vnrCameraGetFocus(vnrCamera c) { return c->at; }
inline vnr::vec3f vnrCameraGetUpVec(vnrCamera c) { return c->up; }

// ---- volumes --------------------------------------------------------------------------------------
// in-memory stand-in of vnrCreateSimpleVolume(scene, mode) (api.h:117): `voxels` already normalised to [0,1]
inline vnrVolume vnrCreateSimpleVolume(const float* voxels, vnr::vec3i dims) {
  auto v = std::make_shared<vnr::SimpleVolumeContext>();
  v->dims = dims;
  v->voxels.assign(voxels, voxels + (size_t)dims.x * dims.y * dims.z);
  return v;
}
inline std::shared_ptr<vnr::NeuralVolumeContext> castNeuralVolume(vnrVolume v) {                    // api.cpp:125-131
  if (!v || !v->isNetwork()) throw std::runtime_error("expecting a neural volume");
  return std::dynamic_pointer_cast<vnr::NeuralVolumeContext>(v);
}
inline std::shared_ptr<vnr::SimpleVolumeContext> castSimpleVolume(vnrVolume v) {                    // api.cpp:133-140
  if (!v || v->isNetwork()) throw std::runtime_error("expecting a simple volume");
  return std::dynamic_pointer_cast<vnr::SimpleVolumeContext>(v);
}

// vnrCreateNeuralVolume(config, dims)                                               api.cpp:190-204
inline vnrVolume vnrCreateNeuralVolume(const vnrJson& config, vnr::vec3i dims, uint32_t seed = 0) {
  if (config.kind == vnrJson::Binary) throw std::runtime_error("expecting a model config, not a params blob");
  const std::string text = config.is_string() ? vnr::read_file(config.data, false) : config.data;
  auto ret = std::make_shared<vnr::NeuralVolumeContext>();
  ret->dims = dims;
  vnr::check(vnr_volume_create(text.c_str(), dims.x, dims.y, dims.z, &ret->h));
  vnr::check(vnr_volume_init_params(ret->h, seed ? seed : (uint32_t)time(nullptr)));   // tcnn_network.h:207: time(NULL)
  return ret;
}
// vnrCreateNeuralVolume(config, groundtruth, online_macrocell_construction)          api.cpp:174-188
inline vnrVolume vnrCreateNeuralVolume(const vnrJson& config, vnrVolume groundtruth, bool online_macrocell_construction = true, uint32_t seed = 0) {
  auto src = castSimpleVolume(groundtruth);
  auto ret = castNeuralVolume(vnrCreateNeuralVolume(config, src->dims, seed));
  vnr::check(vnr_volume_set_groundtruth_f32(ret->h, src->voxels.data()));
  if (!online_macrocell_construction) vnr::check(vnr_volume_macrocell_from_groundtruth(ret->h));
  return ret;
}
// vnrNeuralVolumeSetParams                                                           api.cpp:246-259
inline void vnrNeuralVolumeSetParams(vnrVolume v, const vnrJson& params) {
  const std::string blob = params.is_string() ? vnr::read_file(params.data, true) : params.data;
  vnr::check(vnr_volume_load_params(castNeuralVolume(v)->h, blob.data(), blob.size()));
}
// vnrCreateNeuralVolume(params)                                                      api.cpp:206-220
inline vnrVolume vnrCreateNeuralVolume(const vnrJson& params) {
  const std::string blob = params.is_string() ? vnr::read_file(params.data, true) : params.data;
  int dx, dy, dz; const char* model = nullptr;
  vnr::check(vnr_params_peek(blob.data(), blob.size(), &dx, &dy, &dz, &model));   // throws "expecting a model config with volume dims tag"
  auto ret = vnrCreateNeuralVolume(vnrJson::text(model), vnr::vec3i(dx, dy, dz), 1);
  vnrNeuralVolumeSetParams(ret, vnrJson::binary(blob));
  return ret;
}
inline void vnrNeuralVolumeTrain(vnrVolume v, int steps, bool fast_mode) { vnr::check(vnr_volume_train(castNeuralVolume(v)->h, steps, 0, fast_mode, nullptr)); }   // api.cpp:222-226
inline int vnrNeuralVolumeGetTrainingStep(vnrVolume v) { uint64_t s; double l; vnr::check(vnr_volume_stats(castNeuralVolume(v)->h, &s, &l)); return (int)s; }
inline double vnrNeuralVolumeGetTrainingLoss(vnrVolume v) { uint64_t s; double l; vnr::check(vnr_volume_stats(castNeuralVolume(v)->h, &s, &l)); return l; }
inline double vnrNeuralVolumeGetPSNR(vnrVolume v, bool /*verbose*/) { double p; vnr::check(vnr_volume_psnr(castNeuralVolume(v)->h, &p)); return p; }
inline void vnrNeuralVolumeDecodeProgressive(vnrVolume v) { vnr::check(vnr_volume_decode_progressive(castNeuralVolume(v)->h, nullptr)); }   // api.cpp:228-232
inline int vnrNeuralVolumeGetNumberOfBlobs(vnrVolume v) { int n; vnr::check(vnr_volume_num_blobs(castNeuralVolume(v)->h, &n)); return n; }    // api.cpp:314-318
inline void vnrNeuralVolumeSerializeParams(vnrVolume v, vnrJson& params) {                          // api.cpp:292-298
  const void* p; size_t n;
  vnr::check(vnr_volume_save_params(castNeuralVolume(v)->h, &p, &n));
  params = vnrJson::binary(std::string((const char*)p, n));
}
inline void vnrNeuralVolumeSerializeParams(vnrVolume v, std::string filename) { vnrJson j; vnrNeuralVolumeSerializeParams(v, j); vnrSaveJsonBinary(j, filename); }
inline void vnrVolumeSetClippingBox(vnrVolume v, vnr::vec3f lower, vnr::vec3f upper) { v->clip_lo = lower; v->clip_hi = upper; }
inline vnr::range1f vnrVolumeGetValueRange(vnrVolume) { return vnr::range1f(0.f, 1.f); }

// ---- transfer function (api.h:154-162) ------------------------------------------------------------
inline vnrTransferFunction vnrCreateTransferFunction() { return std::make_shared<vnr::TransferFunction>(); }
inline void vnrTransferFunctionSetColor(vnrTransferFunction t, const std::vector<vnr::vec3f>& colors) { t->color = colors; }
inline void vnrTransferFunctionSetAlpha(vnrTransferFunction t, const std::vector<vnr::vec2f>& alphas) { t->alpha = alphas; }
inline void vnrTransferFunctionSetValueRange(vnrTransferFunction t, vnr::range1f range) { t->range = range; }

// ---- renderer (api.h:168-178) ---------------------------------------------------------------------
inline vnrRenderer vnrCreateRenderer(vnrVolume v) {                                                 // api.cpp:419-459
  auto self = std::make_shared<vnr::RendererContext>();
  self->volume = v;
  vnr::check(vnr_renderer_create(castNeuralVolume(v)->h, &self->h));                                // mode 5 by default (:456)
  vnr::check(vnr_renderer_set_clipping_box(self->h, &v->clip_lo.x, &v->clip_hi.x));                  // set_scene_clipbox (:454)
  return self;
}
inline void vnrRendererSetFramebufferSize(vnrRenderer r, vnr::vec2i fbsize) { vnr::check(vnr_renderer_set_size(r->h, fbsize.x, fbsize.y)); r->size = fbsize; }
inline void vnrRendererSetTransferFunction(vnrRenderer r, vnrTransferFunction t) {                  // api.cpp:485-498
  std::vector<float> alpha; alpha.reserve(t->alpha.size());
  for (auto& a : t->alpha) alpha.push_back(a.y);
  vnr::check(vnr_volume_set_tfn(castNeuralVolume(r->volume)->h, t->color.empty() ? nullptr : &t->color[0].x, (int)t->color.size(),
                                alpha.empty() ? nullptr : alpha.data(), (int)alpha.size(), t->range.lo, t->range.hi));
  vnr::check(vnr_renderer_reset_accumulation(r->h));
}
inline void vnrRendererSetCamera(vnrRenderer r, vnrCamera c) { vnr::check(vnr_renderer_set_camera(r->h, &c->from.x, &c->at.x, &c->up.x, c->fovy)); }
inline void vnrRendererSetMode(vnrRenderer r, int mode) { vnr::check(vnr_renderer_set_mode(r->h, mode)); }
inline void vnrRendererSetDenoiser(vnrRenderer, bool) {}                                            // OptiX denoiser: outside the path
inline void vnrRendererSetVolumeSamplingRate(vnrRenderer r, float v) { vnr::check(vnr_renderer_set_sampling_rate(r->h, v)); }
inline void vnrRendererSetVolumeDensityScale(vnrRenderer r, float v) { vnr::check(vnr_renderer_set_density_scale(r->h, v)); }
inline void vnrRendererResetAccumulation(vnrRenderer r) { vnr::check(vnr_renderer_reset_accumulation(r->h)); }
inline void vnrRender(vnrRenderer r) { vnr::check(vnr_render(r->h)); }                              // api.cpp:522-525
inline vnr::vec4f* vnrRendererMapFrame(vnrRenderer r) {                                             // api.cpp:510-515
  const float* p = vnr_map_frame(r->h);
  if (!p) throw std::runtime_error(vnr_last_error());
  return reinterpret_cast<vnr::vec4f*>(const_cast<float*>(p));
}
inline void vnrMemoryQuery(size_t* used_by_renderer, size_t* used_by_tcnn) { vnr::check(vnr_memory_query(used_by_renderer, used_by_tcnn)); }

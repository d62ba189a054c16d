// vnr_api.hpp -- C++ host-side mirror of the reference's public API (api.h) for the hot path,
// header-only, over the flat C ABI of libvnr_b200.so (vnr_c.h).
//
// Same function names, argument meaning and error behaviour as the reference's api.h (cited per
// function; paths relative to the reference repo): handles are std::shared_ptr, errors are
// std::runtime_error (api.cpp:129,138,215), a vnrJson that "is a string" is a FILE NAME
// (api.cpp:180-185).  Differences, all forced by what is (not) vendored in the reference:
//   * vnrJson IS nlohmann::json, exactly as in the reference's api.h, whenever <json/json.hpp> is on the include path (the
//     reference tree vendors it: -I<reference>/tcnn/dependencies; its apps then compile against this header unchanged).
//     Without it (this repo ships no third-party code) vnrJson is a small value type (JSON text | file name | BSON blob)
//     with the same is_string() = "this is a file name" convention; define VNR_API_NO_NLOHMANN to force that;
//   * vnrCreateSimpleVolume(scene, mode) reads VIDI3D / DIVA scene descriptions of raw binary volumes
//     (serializer.cpp:138-477; other OVR readers are not vendored); an in-memory overload takes a normalised
//     float volume.  A simple volume is carried by a vnr_volume_t with a minimal model, so that its
//     ground truth, macrocells and transfer function live where the renderer expects them;
//   * rendering modes 0-3 (OptiX) return "unsupported" from vnrRender.
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

#if !defined(VNR_API_NO_NLOHMANN) && defined(__has_include)
#if __has_include(<json/json.hpp>)
#include <json/json.hpp>
#define VNR_API_HAS_NLOHMANN 1
#endif
#endif

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
  // a string converts to "file name", as a std::string converts to a nlohmann::json string in the reference's apps
  Json(const std::string& f) : kind(FileName), data(f) {}
  Json(const char* f) : kind(FileName), data(f) {}
  bool is_string() const { return kind == FileName; }
};

inline std::string read_file(const std::string& name, bool binary) {
  std::ifstream f(name, binary ? std::ios::binary : std::ios::in);
  if (!f) throw std::runtime_error("cannot open " + name);
  std::stringstream ss; ss << f.rdbuf();
  return ss.str();
}
inline void check(int rc) { if (rc != VNR_OK) throw std::runtime_error(vnr_last_error()); }

// ---- the two JSON back ends behind one set of accessors -------------------------------------------
#ifdef VNR_API_HAS_NLOHMANN
using json = nlohmann::json;                          // api.h:21
typedef nlohmann::json ApiJson;
namespace jx {
inline ApiJson from_text(const std::string& t) { return nlohmann::json::parse(t, nullptr, true, true); }            // api.cpp:20 (comments allowed)
inline ApiJson from_bson(const std::string& b) { return nlohmann::json::from_bson(b.begin(), b.end()); }             // api.cpp:30
// a JSON description (model config, scene): a string is a file name (api.cpp:180-185), anything else is the description itself
inline std::string text_of(const ApiJson& j, const char* what) {
  if (j.is_string()) return read_file(j.get<std::string>(), false);
  if (j.is_binary()) throw std::runtime_error(std::string("expecting ") + what + ", not a params blob");
  return j.dump();
}
inline std::string scene_arg(const ApiJson& j, int& is_path) {
  if (j.is_string()) { is_path = 1; return j.get<std::string>(); }
  is_path = 0; return j.dump();
}
// serialized parameters: a string is a file name, an object is what json::from_bson produced (api.cpp:246-259)
inline std::string blob_of(const ApiJson& j) {
  if (j.is_string()) return read_file(j.get<std::string>(), true);
  const std::vector<uint8_t> b = nlohmann::json::to_bson(j);
  return std::string(b.begin(), b.end());
}
inline std::string pretty(const ApiJson& j) { return j.dump(4); }                                                    // api.cpp:37 std::setw(4)
}  // namespace jx
#else
typedef Json ApiJson;
namespace jx {
inline ApiJson from_text(const std::string& t) { return Json::text(t); }
inline ApiJson from_bson(const std::string& b) { return Json::binary(b); }
inline std::string text_of(const ApiJson& j, const char* what) {
  if (j.kind == Json::Binary) throw std::runtime_error(std::string("expecting ") + what + ", not a params blob");
  return j.is_string() ? read_file(j.data, false) : j.data;
}
inline std::string scene_arg(const ApiJson& j, int& is_path) {
  if (j.kind == Json::Binary) throw std::runtime_error("expecting a scene description, not a params blob");
  is_path = j.is_string() ? 1 : 0; return j.data;
}
inline std::string blob_of(const ApiJson& j) { return j.is_string() ? read_file(j.data, true) : j.data; }
inline std::string pretty(const ApiJson& j) {
  if (j.kind != Json::Text) throw std::runtime_error("vnrSaveJsonText: not a JSON text value");
  return j.data;
}
}  // namespace jx
#endif

struct Camera { vec3f from{0, 0, -1}, at{0, 0, 0}, up{0, 1, 0}; float fovy = 60.f; };              // instantvnr_types.h:74-83
struct TransferFunction { std::vector<vec3f> color; std::vector<vec2f> alpha; range1f range; };     // api_internal.h

struct VolumeContext {                                                                               // api_internal.h:17-39
  vec3i dims;
  vec3f clip_lo{0, 0, 0}, clip_hi{1, 1, 1};
  vec3f scaling{1, 1, 1};                           // data transform = scale(scaling) * translate(-dims/2) * scale(dims)
  range1f value_range{0.f, 1.f};                    // unnormalised range of the data (desc.range)
  virtual bool isNetwork() const = 0;
  virtual ~VolumeContext() {}
};
struct SimpleVolumeContext : VolumeContext {                                                        // api_internal.h:23-30
  std::vector<float> voxels;                        // in-memory form: normalised to [0,1], x fastest
  vnr_scene_t* scene = nullptr;                     // scene form: MultiVolume descriptor (files per time step)
  std::string mode = "GPU";                         // sampling mode (neural_sampler.cpp:1210-1229)
  int n_timesteps = 1, timestep = 0, value_type = 8;
  vnr_volume_t* h = nullptr;                        // carrier of ground truth + macrocell + tfn (minimal model)
  bool isNetwork() const override { return false; }
  ~SimpleVolumeContext() override { vnr_volume_release(h); vnr_scene_release(scene); }
};
struct NeuralVolumeContext : VolumeContext {
  vnr_volume_t* h = nullptr;
  std::shared_ptr<VolumeContext> groundtruth;       // keeps the source alive (NeuralVolume holds a raw pointer, network.h)
  bool isNetwork() const override { return true; }
  ~NeuralVolumeContext() override { vnr_volume_release(h); }
};
inline vnr_volume_t* handle_of(const std::shared_ptr<VolumeContext>& v);
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
typedef vnr::ApiJson vnrJson;

enum vnrRenderMode {                                                                                 // api.h:36-60
  VNR_OPTIX_NO_SHADING = 0, VNR_OPTIX_GRADIENT_SHADING, VNR_OPTIX_FULL_SHADOW, VNR_OPTIX_SINGLE_SHADE_HEURISTIC,
  VNR_RAYMARCHING_NO_SHADING_DECODING, VNR_RAYMARCHING_NO_SHADING_SAMPLE_STREAMING, VNR_RAYMARCHING_NO_SHADING_IN_SHADER,
  VNR_RAYMARCHING_GRADIENT_SHADING_DECODING, VNR_RAYMARCHING_GRADIENT_SHADING_SAMPLE_STREAMING, VNR_RAYMARCHING_GRADIENT_SHADING_IN_SHADER,
  VNR_RAYMARCHING_SINGLE_SHADE_HEURISTIC_DECODING, VNR_RAYMARCHING_SINGLE_SHADE_HEURISTIC_SAMPLE_STREAMING,
  VNR_RAYMARCHING_SINGLE_SHADE_HEURISTIC_IN_SHADER,
  VNR_PATHTRACING_DECODING, VNR_PATHTRACING_SAMPLE_STREAMING, VNR_PATHTRACING_IN_SHADER, VNR_INVALID,
};

inline bool vnrRequireDecoding(int m) {                                                             // api.h:62-87
  if (m < 0 || m >= VNR_INVALID) throw std::runtime_error("unknown rendering mode");
  return m <= VNR_OPTIX_SINGLE_SHADE_HEURISTIC || ((m - VNR_RAYMARCHING_NO_SHADING_DECODING) % 3) == 0;
}

// ---- json I/O (api.h:89-95) -----------------------------------------------------------------------
inline vnrJson vnrCreateJsonText(std::string filename) { return vnr::jx::from_text(vnr::read_file(filename, false)); }
inline vnrJson vnrCreateJsonBinary(std::string filename) { return vnr::jx::from_bson(vnr::read_file(filename, true)); }
inline void vnrLoadJsonText(vnrJson& j, std::string filename) { j = vnrCreateJsonText(filename); }
inline void vnrLoadJsonBinary(vnrJson& j, std::string filename) { j = vnrCreateJsonBinary(filename); }
inline void vnrSaveJsonText(const vnrJson& j, std::string filename) {
  const std::string text = vnr::jx::pretty(j);
  std::ofstream f(filename);
  if (!f) throw std::runtime_error("cannot write " + filename);
  f << text << std::endl;
}
inline void vnrSaveJsonBinary(const vnrJson& j, std::string filename) {
  const std::string blob = vnr::jx::blob_of(j);
  std::ofstream f(filename, std::ios::binary);
  if (!f) throw std::runtime_error("cannot write " + filename);
  f.write(blob.data(), (std::streamsize)blob.size());
}

// ---- camera (api.h:102-110) -----------------------------------------------------------------------
inline vnrCamera vnrCreateCamera() { return std::make_shared<vnr::Camera>(); }
inline void vnrCameraSet(vnrCamera c, vnr::vec3f from, vnr::vec3f at, vnr::vec3f up) { c->from = from; c->at = at; c->up = up; }
namespace vnr {
struct SceneHandle {                                  // RAII over vnr_scene_t for the scene overloads
  vnr_scene_t* s = nullptr;
  explicit SceneHandle(const ApiJson& scene) {
    int is_path = 0;
    const std::string arg = jx::scene_arg(scene, is_path);
    check(vnr_scene_create(arg.c_str(), is_path, &s));
  }
  ~SceneHandle() { vnr_scene_release(s); }
  vnr_scene_t* release() { vnr_scene_t* r = s; s = nullptr; return r; }
};
}  // namespace vnr
inline void vnrCameraSet(vnrCamera c, const vnrJson& scene) {                                       // api.cpp:94-103
  vnr::SceneHandle sc(scene);
  vnr::check(vnr_scene_camera(sc.s, &c->from.x, &c->at.x, &c->up.x, &c->fovy));
}
inline vnrCamera vnrCreateCamera(const vnrJson& scene) { auto c = std::make_shared<vnr::Camera>(); vnrCameraSet(c, scene); return c; }   // api.cpp:73-83
inline vnr::vec3f vnrCameraGetPosition(vnrCamera c) { return c->from; }
inline vnr::vec3f vnrCameraGetFocus(vnrCamera c) { return c->at; }
inline vnr::vec3f vnrCameraGetUpVec(vnrCamera c) { return c->up; }

// ---- volumes --------------------------------------------------------------------------------------
namespace vnr {
// model of the vnr_volume_t that carries a simple volume: never trained or decoded, sized to be negligible
inline const char* carrier_model() {
  return "{\"encoding\":{\"otype\":\"HashGrid\",\"n_levels\":2,\"n_features_per_level\":8,\"log2_hashmap_size\":4,\"base_resolution\":2},"
         "\"network\":{\"otype\":\"FullyFusedMLP\",\"n_neurons\":64,\"n_hidden_layers\":1,\"activation\":\"ReLU\",\"output_activation\":\"None\"},"
         "\"loss\":{\"otype\":\"L1\"},\"optimizer\":{\"otype\":\"Adam\"}}";
}
// StaticSampler::load / OutOfCoreSampler (neural_sampler.cpp:223-288,1040-1120,1205-1232): time step t of the scene -> `dst`
inline void load_timestep(vnr_volume_t* dst, const SimpleVolumeContext& sv, int t) {
  if (!sv.scene) { check(vnr_volume_set_groundtruth_f32(dst, sv.voxels.data())); return; }
  const char* file = nullptr; uint64_t offset = 0; int big = 0;
  check(vnr_scene_timestep(sv.scene, t, &file, &offset, &big));
  int dims[3], vt, nt, has; float rg[2];
  check(vnr_scene_volume(sv.scene, dims, &vt, &nt, rg, &has));
  if (sv.mode == "GPU") {
    check(vnr_volume_set_groundtruth_file(dst, file, vt, offset, big, has ? rg[0] : 0.f, has ? rg[1] : 0.f, nullptr));
  } else if (sv.mode == "OUT_OF_CORE" || sv.mode == "VIRTUAL_MEMORY") {
    // both stream random slabs of the file; the reference's mmap variant differs only in how the host reads them
    if (big) throw std::runtime_error("out-of-core sampling of big-endian files is not supported");
    check(vnr_volume_set_groundtruth_outofcore(dst, file, vt, offset, rg[0], rg[1], 0, 0));   // range required (:1068-1070)
  } else throw std::runtime_error("unknown sampling mode: " + sv.mode);
}
}  // namespace vnr
// vnrCreateSimpleVolume(scene, mode)                                                  api.cpp:143-156
inline vnrVolume vnrCreateSimpleVolume(const vnrJson& scene, std::string mode, bool save_loaded_volume = false) {
  if (save_loaded_volume) throw std::runtime_error("save_loaded_volume is not supported");
  auto v = std::make_shared<vnr::SimpleVolumeContext>();
  vnr::SceneHandle sc(scene);
  int dims[3], has; float rg[2];
  vnr::check(vnr_scene_volume(sc.s, dims, &v->value_type, &v->n_timesteps, rg, &has));
  v->dims = vnr::vec3i(dims[0], dims[1], dims[2]);
  if (has) v->value_range = vnr::range1f(rg[0], rg[1]);
  v->scene = sc.release();
  v->mode = mode;
  return v;
}
// in-memory stand-in: `voxels` already normalised to [0,1]
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

namespace vnr {
// the vnr_volume_t of a simple volume: created on first use (renderer / time-step change), in-core only
inline vnr_volume_t* carrier(SimpleVolumeContext& sv) {
  if (!sv.h) {
    if (sv.mode != "GPU") throw std::runtime_error("a simple volume is rendered from HBM: sampling mode must be \"GPU\"");
    check(vnr_volume_create(carrier_model(), sv.dims.x, sv.dims.y, sv.dims.z, &sv.h));
    load_timestep(sv.h, sv, sv.timestep);
    check(vnr_volume_macrocell_from_groundtruth(sv.h));                                            // SimpleVolume::load -> compute_everything (sampler.cu:12-17)
  }
  return sv.h;
}
inline vnr_volume_t* handle_of(const std::shared_ptr<VolumeContext>& v) {
  if (!v) throw std::runtime_error("null volume");
  return v->isNetwork() ? std::dynamic_pointer_cast<NeuralVolumeContext>(v)->h : carrier(*std::dynamic_pointer_cast<SimpleVolumeContext>(v));
}
}  // namespace vnr
inline void vnrSimpleVolumeSetCurrentTimeStep(vnrVolume v, int time) {                              // api.cpp:158-162, sampler.cu:19-26
  auto sv = castSimpleVolume(v);
  if (time < 0 || time >= sv->n_timesteps) throw std::runtime_error("time step out of range");
  sv->timestep = time;
  if (sv->h) { vnr::load_timestep(sv->h, *sv, time); vnr::check(vnr_volume_macrocell_from_groundtruth(sv->h)); }
}
inline int vnrSimpleVolumeGetNumberOfTimeSteps(vnrVolume v) { return castSimpleVolume(v)->n_timesteps; }   // api.cpp:164-168

// vnrCreateNeuralVolume(config, dims)                                               api.cpp:190-204
inline vnrVolume vnrCreateNeuralVolume(const vnrJson& config, vnr::vec3i dims, uint32_t seed = 0) {
  const std::string text = vnr::jx::text_of(config, "a model config");
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
  vnr::load_timestep(ret->h, *src, src->timestep);
  if (!online_macrocell_construction) {
    if (src->mode != "GPU") throw std::runtime_error("offline macrocells need the volume in HBM (sampling mode \"GPU\")");
    vnr::check(vnr_volume_macrocell_from_groundtruth(ret->h));
  }
  ret->groundtruth = groundtruth; ret->value_range = src->value_range;
  return ret;
}
// vnrNeuralVolumeSetParams                                                           api.cpp:246-259
inline void vnrNeuralVolumeSetParams(vnrVolume v, const vnrJson& params) {
  const std::string blob = vnr::jx::blob_of(params);
  vnr::check(vnr_volume_load_params(castNeuralVolume(v)->h, blob.data(), blob.size()));
}
// vnrCreateNeuralVolume(params)                                                      api.cpp:206-220
inline vnrVolume vnrCreateNeuralVolume(const vnrJson& params) {
  const std::string blob = vnr::jx::blob_of(params);
  int dx, dy, dz; const char* model = nullptr;
  vnr::check(vnr_params_peek(blob.data(), blob.size(), &dx, &dy, &dz, &model));   // throws "expecting a model config with volume dims tag"
  auto ret = vnrCreateNeuralVolume(vnr::jx::from_text(model), vnr::vec3i(dx, dy, dz), 1);
  vnr::check(vnr_volume_load_params(castNeuralVolume(ret)->h, blob.data(), blob.size()));
  return ret;
}
// vnrNeuralVolumeSetModel                                                            api.cpp:261-270
inline void vnrNeuralVolumeSetModel(vnrVolume v, const vnrJson& config, uint32_t seed = 0) {
  const std::string text = vnr::jx::text_of(config, "a model config");
  vnr::check(vnr_volume_set_model(castNeuralVolume(v)->h, text.c_str(), seed ? seed : (uint32_t)time(nullptr)));
}
inline void vnrNeuralVolumeTrain(vnrVolume v, int steps, bool fast_mode) { vnr::check(vnr_volume_train(castNeuralVolume(v)->h, steps, 0, fast_mode, nullptr)); }   // api.cpp:222-226
inline int vnrNeuralVolumeGetTrainingStep(vnrVolume v) { uint64_t s; double l; vnr::check(vnr_volume_stats(castNeuralVolume(v)->h, &s, &l)); return (int)s; }
inline double vnrNeuralVolumeGetTrainingLoss(vnrVolume v) { uint64_t s; double l; vnr::check(vnr_volume_stats(castNeuralVolume(v)->h, &s, &l)); return l; }
inline double vnrNeuralVolumeGetPSNR(vnrVolume v, bool /*verbose*/) { double p; vnr::check(vnr_volume_psnr(castNeuralVolume(v)->h, &p)); return p; }
inline double vnrNeuralVolumeGetSSIM(vnrVolume v, bool /*verbose*/) { double p; vnr::check(vnr_volume_ssim(castNeuralVolume(v)->h, &p, nullptr)); return p; }   // api.cpp:286-290
inline double vnrNeuralVolumeGetTestingLoss(vnrVolume v) { double l; vnr::check(vnr_volume_test_loss(castNeuralVolume(v)->h, 0, &l)); return l; }            // api.cpp:292-298
inline void vnrNeuralVolumeDecodeInference(vnrVolume v, std::string filename) { vnr::check(vnr_volume_export(castNeuralVolume(v)->h, filename.c_str(), 0, nullptr)); }   // api.cpp:234-238
inline void vnrNeuralVolumeDecodeReference(vnrVolume v, std::string filename) { vnr::check(vnr_volume_export(castNeuralVolume(v)->h, filename.c_str(), 1, nullptr)); }   // api.cpp:240-244
inline void vnrNeuralVolumeDecodeProgressive(vnrVolume v) { vnr::check(vnr_volume_decode_progressive(castNeuralVolume(v)->h, nullptr)); }   // api.cpp:228-232
inline int vnrNeuralVolumeGetNumberOfBlobs(vnrVolume v) { int n; vnr::check(vnr_volume_num_blobs(castNeuralVolume(v)->h, &n)); return n; }    // api.cpp:314-318
inline void vnrNeuralVolumeSerializeParams(vnrVolume v, vnrJson& params) {                          // api.cpp:292-298
  const void* p; size_t n;
  vnr::check(vnr_volume_save_params(castNeuralVolume(v)->h, &p, &n));
  params = vnr::jx::from_bson(std::string((const char*)p, n));
}
inline void vnrNeuralVolumeSerializeParams(vnrVolume v, std::string filename) { vnrJson j; vnrNeuralVolumeSerializeParams(v, j); vnrSaveJsonBinary(j, filename); }
// vnrVolumeSetClippingBox (api.cpp:330-348): `lower` / `upper` are in voxel units [0, dims]; they go through the inverse
// data transform (scale(scaling) * translate(-dims/2) * scale(dims)) into the unit cube the renderer clips in.
// As in the reference the box is read when the renderer is created (set_scene_clipbox, api.cpp:454).
inline void vnrVolumeSetClippingBox(vnrVolume v, vnr::vec3f lower, vnr::vec3f upper) {
  const float d[3] = {(float)v->dims.x, (float)v->dims.y, (float)v->dims.z}, sc[3] = {v->scaling.x, v->scaling.y, v->scaling.z};
  float lo[3] = {lower.x, lower.y, lower.z}, hi[3] = {upper.x, upper.y, upper.z};
  for (int k = 0; k < 3; ++k) { lo[k] = (lo[k] - d[k] / 2.f) / (sc[k] * d[k]) + 0.5f; hi[k] = (hi[k] - d[k] / 2.f) / (sc[k] * d[k]) + 0.5f; }
  v->clip_lo = vnr::vec3f(lo[0], lo[1], lo[2]); v->clip_hi = vnr::vec3f(hi[0], hi[1], hi[2]);
}
// vnrVolumeSetScaling (api.cpp:350-361): scale(s) is multiplied onto the current data transform
inline void vnrVolumeSetScaling(vnrVolume v, vnr::vec3f scale) { v->scaling = vnr::vec3f(v->scaling.x * scale.x, v->scaling.y * scale.y, v->scaling.z * scale.z); }
// both volume kinds hold data normalised to [0,1] (network.cu:977-981, sampler.h:85)
inline vnr::range1f vnrVolumeGetValueRange(vnrVolume) { return vnr::range1f(0.f, 1.f); }

// ---- transfer function (api.h:154-162) ------------------------------------------------------------
inline vnrTransferFunction vnrCreateTransferFunction() { return std::make_shared<vnr::TransferFunction>(); }
// vnrCreateTransferFunction(scene) (api.cpp:372-382): value range always; colour / alpha tables only when the scene gives
// them explicitly (the OVR tfn-module formats are not vendored in the reference -> "unsupported")
inline vnrTransferFunction vnrCreateTransferFunction(const vnrJson& scene) {
  auto t = std::make_shared<vnr::TransferFunction>();
  vnr::SceneHandle sc(scene);
  const float *rgb = nullptr, *alpha = nullptr; int n_rgb = 0, n_alpha = 0, has = 0; float rg[2] = {0.f, 1.f};
  vnr::check(vnr_scene_tfn(sc.s, &rgb, &n_rgb, &alpha, &n_alpha, rg, &has));
  for (int i = 0; i < n_rgb; ++i) t->color.push_back(vnr::vec3f(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]));
  for (int i = 0; i < n_alpha; ++i) t->alpha.push_back(vnr::vec2f(alpha[2 * i], alpha[2 * i + 1]));
  if (has) t->range = vnr::range1f(rg[0], rg[1]);
  return t;
}
inline void vnrTransferFunctionSetColor(vnrTransferFunction t, const std::vector<vnr::vec3f>& colors) { t->color = colors; }
inline void vnrTransferFunctionSetAlpha(vnrTransferFunction t, const std::vector<vnr::vec2f>& alphas) { t->alpha = alphas; }
inline void vnrTransferFunctionSetValueRange(vnrTransferFunction t, vnr::range1f range) { t->range = range; }

inline const std::vector<vnr::vec3f>& vnrTransferFunctionGetColor(vnrTransferFunction t) { return t->color; }
inline const std::vector<vnr::vec2f>& vnrTransferFunctionGetAlpha(vnrTransferFunction t) { return t->alpha; }
inline const vnr::range1f& vnrTransferFunctionGetValueRange(vnrTransferFunction t) { return t->range; }

// ---- renderer (api.h:168-178) ---------------------------------------------------------------------
inline vnrRenderer vnrCreateRenderer(vnrVolume v) {                                                 // api.cpp:419-459
  auto self = std::make_shared<vnr::RendererContext>();
  self->volume = v;
  vnr::check(vnr_renderer_create(vnr::handle_of(v), &self->h));                                     // mode 5 by default (:456)
  if (!v->isNetwork()) vnr::check(vnr_renderer_set_groundtruth_source(self->h, 1));                  // set_scene(source.texture(), ...) :441-452
  vnr::check(vnr_renderer_set_scaling(self->h, &v->scaling.x));                                      // get_data_transform()
  vnr::check(vnr_renderer_set_clipping_box(self->h, &v->clip_lo.x, &v->clip_hi.x));                  // set_scene_clipbox (:454)
  return self;
}
inline void vnrRendererSetFramebufferSize(vnrRenderer r, vnr::vec2i fbsize) { vnr::check(vnr_renderer_set_size(r->h, fbsize.x, fbsize.y)); r->size = fbsize; }
inline void vnrRendererSetTransferFunction(vnrRenderer r, vnrTransferFunction t) {                  // api.cpp:485-498
  std::vector<float> alpha; alpha.reserve(t->alpha.size());
  for (auto& a : t->alpha) alpha.push_back(a.y);
  vnr::check(vnr_volume_set_tfn(vnr::handle_of(r->volume), t->color.empty() ? nullptr : &t->color[0].x, (int)t->color.size(),
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

// ---- misc (api.h:184-188) -------------------------------------------------------------------------
inline void vnrRelease(void*) {}   // handles are shared_ptr: nothing to do (the reference's vnrRelease is declared, never defined)
inline void vnrMemoryQuery(size_t* used_by_renderer, size_t* used_by_tcnn) { vnr::check(vnr_memory_query(used_by_renderer, used_by_tcnn)); }
inline void vnrMemoryQueryPrint(const char* str) {                                                  // api.cpp:538-552
  size_t r = 0, n = 0; vnrMemoryQuery(&r, &n);
  std::printf("%s: total used by renderer = %.3f MB, network = %.3f MB\n", str, r / 1048576.0, n / 1048576.0);
}
inline void vnrFreeTemporaryGPUMemory() {}   // tcnn's stream-ordered arena (api.cpp:554-557) has no counterpart: all buffers are owned by their objects

// ref_marcher_driver.cu -- C driver around the REFERENCE's own renderer kernels, compiled unmodified and in place from
// /root/reference: core/renderer/method_raymarching.cu, core/renderer/method_pathtracing.cu, core/macrocell.cu,
// core/instantvnr_types.cu (+ their headers dda.h, raytracing.h, instantvnr_types.h, array.h, macrocell.h).  Only what is NOT in
// /root/reference is supplied here: the OVR framework headers those sources include (oracle/ovr_shim/: vector types, LCG,
// CUDABuffer, kernel launch helpers) and the host-side frame setup of renderer.cpp / object.cpp, which drags in OVR's
// framebuffer / colormap code and is therefore restated below, line by line, with citations.
// TEST INFRASTRUCTURE (checker + baseline): builds oracle/_ref/libvnr_marcher_ref.so; only tests/, tools/ and bench.py's
// reference arm load it.  It needs a GPU (the reference has no CPU path).
#include <cstring>
#include <string>
#include <vector>

#include "core/instantvnr_types.h"
#include "core/macrocell.h"
#include "core/renderer/method_pathtracing.h"
#include "core/renderer/method_raymarching.h"

using namespace vnr;

// NeuralVolume::inference is defined in core/network.cu, which needs tiny-cuda-nn AND the OVR framework.  Its body
// (network.cu:1043-1052) pads the batch to a multiple of 256 and calls the network; restated here around a decoder callback
// (the reference's tiny-cuda-nn build of oracle/ref_driver, or any other decoder under test).  The marcher only ever calls
// this one member through its NeuralVolume*, so the pointer it gets is a DecoderHook in disguise (no NeuralVolume object is
// ever constructed).
typedef int (*refm_decode_fn)(void* user, const void* d_xyz, void* d_out, size_t n, void* stream);
struct DecoderHook { refm_decode_fn fn; void* user; uint64_t calls, coords; };
void NeuralVolume::inference(int len, const float* d_input, float* d_output, cudaStream_t stream) {
  DecoderHook* hook = reinterpret_cast<DecoderHook*>(this);
  const uint32_t padded = ((uint32_t)len + 255u) / 256u * 256u;          // util::next_multiple<uint32_t>(len, 256)
  hook->calls += 1; hook->coords += padded;
  if (hook->fn(hook->user, d_input, d_output, padded, stream) != 0) throw std::runtime_error("decoder callback failed");
}

namespace {

struct Scene {
  vec3i dims;
  cudaArray_t array = nullptr;
  Array3DScalar volume;                       // DeviceVolume::volume (rdims stays 0: nothing in the reference sets it)
  MacroCell macrocell;
  TransferFunctionObject tfn;
  bool have_tfn = false;
  LaunchParams params;                        // persistent: the light direction flip of renderer.cpp:98-101 sticks
  CUDABuffer d_volume, accumulation, frame;
  MethodRayMarching raymarching;
  MethodPathTracing pathtracing;
  DecoderHook hook{nullptr, nullptr, 0, 0};
  float sampling_rate = 1.f, density_scale = 1.f;
  box3f clipbox = box3f(vec3f(0.f), vec3f(1.f));
  vec2i size = vec2i(0, 0);
  bool reset = true;
  cudaStream_t stream = nullptr;
};

thread_local std::string g_err;
template <typename F> int guard(F f) {
  try { f(); return 0; } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

}  // namespace

extern "C" {

const char* refm_last_error() { return g_err.c_str(); }

// SimpleVolume::load (core/sampler.cu:5-17): the normalised float volume goes into a 3-D texture (CreateArray3DScalar, array.h:61-75:
// linear filter, element read mode) and MacroCell::set_shape + allocate + compute_everything run on it.
int refm_create(const int* dims3, const float* h_volume, void** out) {
  return guard([&] {
    Scene* s = new Scene();
    s->dims = vec3i(dims3[0], dims3[1], dims3[2]);
    CUDA_CHECK(cudaStreamCreate(&s->stream));
    cudaTextureObject_t tex = 0;
    CreateArray3DScalar<float>(s->array, tex, s->dims, /*trilinear=*/true, (void*)h_volume);
    s->volume.type = VALUE_TYPE_FLOAT; s->volume.dims = s->dims; s->volume.data = tex;       // set_volume(data, type, dims, range) object.cpp:375-383
    s->macrocell.set_shape(s->dims);
    s->macrocell.allocate();
    s->macrocell.compute_everything(tex);
    CUDA_CHECK(cudaDeviceSynchronize());
    *out = s;
  });
}

void refm_release(void* h) {
  Scene* s = (Scene*)h;
  if (!s) return;
  cudaDeviceSynchronize();
  s->tfn.clean();
  if (s->volume.data) cudaDestroyTextureObject(s->volume.data);
  if (s->array) cudaFreeArray(s->array);
  s->d_volume.free(); s->accumulation.free(); s->frame.free();
  s->raymarching.clear(0); s->pathtracing.clear(0);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

// replace the texture content (a progressively decoded volume for the "decoding" modes); macrocells are NOT recomputed
int refm_update_volume(void* h, const float* h_volume) {
  return guard([&] {
    Scene* s = (Scene*)h;
    CopyLinearMemoryToArray<float>((void*)h_volume, s->array, s->dims, cudaMemcpyHostToDevice);
  });
}

// NeuralVolume keeps its own, online-trained macrocell value ranges (network.cu:249-257): overwrite the ranges computed from the
// texture with `h_value_range` (2 floats per cell, stored offset by -1 / +1 as macrocell.cu:35-39) and refresh the max opacity
int refm_set_macrocell_value_range(void* h, const float* h_value_range) {
  return guard([&] {
    Scene* s = (Scene*)h;
    const size_t cells = s->macrocell.dims().long_product();
    CUDA_CHECK(cudaMemcpy(s->macrocell.d_value_range(), h_value_range, cells * 2 * sizeof(float), cudaMemcpyHostToDevice));
    if (s->have_tfn) { s->macrocell.update_max_opacity(s->tfn.tfn, s->stream); CUDA_CHECK(cudaStreamSynchronize(s->stream)); }
  });
}

// NeuralVolume's online macrocell construction (network.cu:249-257): MacroCell::allocate (zeroed ranges) and, per training
// batch, MacroCell::update_explicit on the batch's coordinates and target values + update_max_opacity
int refm_macrocell_reset(void* h) {
  return guard([&] { Scene* s = (Scene*)h; s->macrocell.allocate(); CUDA_CHECK(cudaDeviceSynchronize()); });
}
int refm_macrocell_update_explicit(void* h, const float* d_xyz, const float* d_values, size_t n) {
  return guard([&] {
    Scene* s = (Scene*)h;
    s->macrocell.update_explicit((vec3f*)d_xyz, (float*)d_values, n, s->stream);
    if (s->have_tfn) s->macrocell.update_max_opacity(s->tfn.tfn, s->stream);
    CUDA_CHECK(cudaStreamSynchronize(s->stream));
  });
}

int refm_get_macrocell(void* h, int* mc_dims3, float* h_value_range, float* h_max_opacity) {
  return guard([&] {
    Scene* s = (Scene*)h;
    const vec3i d = s->macrocell.dims();
    if (mc_dims3) { mc_dims3[0] = d.x; mc_dims3[1] = d.y; mc_dims3[2] = d.z; }
    const size_t cells = d.long_product();
    CUDA_CHECK(cudaDeviceSynchronize());
    if (h_value_range) CUDA_CHECK(cudaMemcpy(h_value_range, s->macrocell.d_value_range(), cells * 2 * sizeof(float), cudaMemcpyDeviceToHost));
    if (h_max_opacity) CUDA_CHECK(cudaMemcpy(h_max_opacity, s->macrocell.d_max_opacity(), cells * sizeof(float), cudaMemcpyDeviceToHost));
  });
}

// SimpleVolume::set_transfer_function (core/sampler.cu:28-35): TransferFunctionObject::set_transfer_function + update_max_opacity
int refm_set_transfer_function(void* h, const float* rgb, int n_rgb, const float* alpha_pairs, int n_alpha, float lo, float hi) {
  return guard([&] {
    Scene* s = (Scene*)h;
    std::vector<vec3f> c(n_rgb); std::vector<vec2f> o(n_alpha);
    for (int i = 0; i < n_rgb; ++i) c[i] = vec3f(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    for (int i = 0; i < n_alpha; ++i) o[i] = vec2f(alpha_pairs[2 * i], alpha_pairs[2 * i + 1]);
    s->tfn.set_transfer_function(c, o, range1f(lo, hi), s->stream);
    s->have_tfn = true;
    s->macrocell.update_max_opacity(s->tfn.tfn, s->stream);
    CUDA_CHECK(cudaStreamSynchronize(s->stream));
    s->reset = true;
  });
}

int refm_set_decoder(void* h, refm_decode_fn fn, void* user) {
  return guard([&] { Scene* s = (Scene*)h; s->hook.fn = fn; s->hook.user = user; s->reset = true; });
}
int refm_set_sampling(void* h, float sampling_rate, float density_scale) {
  return guard([&] { Scene* s = (Scene*)h; s->sampling_rate = sampling_rate; s->density_scale = density_scale; s->reset = true; });
}
int refm_set_clipbox(void* h, const float* lo, const float* hi) {
  return guard([&] { Scene* s = (Scene*)h; s->clipbox = box3f(vec3f(lo[0], lo[1], lo[2]), vec3f(hi[0], hi[1], hi[2])); s->reset = true; });
}
int refm_reset_accumulation(void* h) { return guard([&] { ((Scene*)h)->reset = true; }); }

// One vnrRender (MainRenderer::render, renderer.cpp:59-140) in rendering mode `mode` (api.h:36-60), `neural` != 0: the scene
// is a NeuralVolume (render_neural, :183-225: sample-streaming / in-shader modes decode through the hook), else a SimpleVolume
// (render_normal, :143-180).  Returns the host copy of the frame buffer (width * height float4) and the decode statistics.
int refm_render(void* h, int mode, int neural, int width, int height, const float* from, const float* at, const float* up, float fovy,
                float* h_frame, uint64_t* stats2) {
  return guard([&] {
    Scene* s = (Scene*)h;
    if (!s->have_tfn) throw std::runtime_error("no transfer function");
    LaunchParams& params = s->params;
    const size_t npix = (size_t)width * height;
    if (s->size.x != width || s->size.y != height) {                    // MainRenderer::resize
      s->accumulation.resize(npix * sizeof(vec4f)); s->frame.resize(npix * sizeof(vec4f));
      s->size = vec2i(width, height); s->reset = true;
    }
    params.frame.size = s->size;
    params.accumulation = (vec4f*)s->accumulation.d_pointer();         // renderer.cpp:66-76
    params.frame.rgba = (vec4f*)s->frame.d_pointer();
    // StructuredRegularVolume: set_volume / set_macrocell / set_transfer_function / set_clipping / commit (object.cpp:300-390)
    DeviceVolume self;
    self.volume = s->volume;
    self.tfn = s->tfn.tfn;                                               // range already clamped to the data range [0,1] by the caller
    self.macrocell_value_range = s->macrocell.d_value_range();
    self.macrocell_max_opacity = s->macrocell.d_max_opacity();
    self.macrocell_dims = s->macrocell.dims();
    self.macrocell_spacings = s->macrocell.spacings();
    self.macrocell_spacings_rcp = 1.f / s->macrocell.spacings();
    self.bbox = s->clipbox;
    self.step = 1.f / s->sampling_rate; self.step_rcp = s->sampling_rate;
    self.grad_step = vec3f(1.f / vec3f(self.volume.dims));
    self.density_scale = s->density_scale;
    s->d_volume.resize(sizeof(DeviceVolume));
    CUDA_CHECK(cudaMemcpyAsync(s->d_volume.d_pointer(), &self, sizeof(DeviceVolume), cudaMemcpyHostToDevice, s->stream));
    // network.cu:569 / neural_sampler: object [0,1]^3 -> world box of size dims centred at the origin
    params.transform = affine3f::translate(vec3f(s->dims) / vec3f(-2.f)) * affine3f::scale(vec3f(s->dims));
    // camera (renderer.cpp:87-96)
    const vec3f cfrom(from[0], from[1], from[2]), cat(at[0], at[1], at[2]), cup(up[0], up[1], up[2]);
    const float t = 2.f * tan(fovy * 0.5f * (float)M_PI / 180.f);
    const float aspect = params.frame.size.x / float(params.frame.size.y);
    params.last_camera = params.camera;
    params.camera.position = cfrom;
    params.camera.direction = normalize(cat - cfrom);
    params.camera.horizontal = t * aspect * normalize(cross(params.camera.direction, cup));
    params.camera.vertical = cross(params.camera.horizontal, params.camera.direction) / aspect;
    if (dot(params.camera.direction, params.light_directional_dir) > 0) params.light_directional_dir *= -1;   // :98-101
    if (s->reset) params.frame_index = 0;                                // :104-105
    params.frame_index++;
    s->reset = false;
    s->hook.calls = 0; s->hook.coords = 0;
    DeviceVolume* dv = (DeviceVolume*)s->d_volume.d_pointer();
    NeuralVolume* nvr = neural ? reinterpret_cast<NeuralVolume*>(&s->hook) : nullptr;
    if (neural && !s->hook.fn) throw std::runtime_error("no decoder set");
    // render_normal / render_neural (renderer.cpp:143-225)
    const int m = mode;
    if (m >= 13 && m <= 15) {
      if (m == 13) s->pathtracing.render(s->stream, params, dv);
      else s->pathtracing.render(s->stream, params, dv, nvr, m == 14);
    } else if (m >= 4 && m <= 12) {
      const MethodRayMarching::ShadingMode sh = m <= 6 ? MethodRayMarching::NO_SHADING : (m <= 9 ? MethodRayMarching::GRADIENT_SHADING : MethodRayMarching::SINGLE_SHADE_HEURISTIC);
      const int kind = (m - 4) % 3;                                      // 0 decoding, 1 sample streaming, 2 in shader
      if (kind == 0) s->raymarching.render(s->stream, params, sh, dv);
      else if (kind == 2 && nvr) throw std::runtime_error("in-shader modes on a neural volume need ENABLE_IN_SHADER (tcnn device API); not built");
      else s->raymarching.render(s->stream, params, sh, dv, nvr, kind == 1);
    } else throw std::runtime_error("mode outside the ray-marching / path-tracing range");
    CUDA_CHECK(cudaMemcpyAsync(h_frame, params.frame.rgba, npix * sizeof(vec4f), cudaMemcpyDeviceToHost, s->stream));
    CUDA_CHECK(cudaStreamSynchronize(s->stream));
    if (stats2) { stats2[0] = s->hook.calls; stats2[1] = s->hook.coords; }
  });
}

}  // extern "C"

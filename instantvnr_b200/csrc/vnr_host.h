// vnr_host.h -- host-side state behind the C ABI (include/vnr_c.h) and the launchers
// of the CUDA kernels.  Product code: never includes or links anything from oracle/.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "vnr_device.cuh"

namespace vnr {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
struct InvalidError : std::runtime_error { using std::runtime_error::runtime_error; };
struct StateError : std::runtime_error { using std::runtime_error::runtime_error; };
struct UnsupportedError : std::runtime_error { using std::runtime_error::runtime_error; };

#define VNR_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (expr);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      throw ::vnr::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

// pcg32 (tcnn/dependencies/pcg32/pcg32.h), host + device
struct Pcg32 {
  uint64_t state, inc;
  __host__ __device__ static constexpr uint64_t mult() { return 0x5851f42d4c957f2dULL; }
  __host__ __device__ void seed(uint64_t initstate, uint64_t initseq = 1u) {
    state = 0U; inc = (initseq << 1u) | 1u; next_uint(); state += initstate; next_uint();
  }
  __host__ __device__ uint32_t next_uint() {
    uint64_t old = state;
    state = old * mult() + inc;
    uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xs >> rot) | (xs << ((~rot + 1u) & 31));
  }
  __host__ __device__ float next_float() {
    union { uint32_t u; float f; } x;
    x.u = (next_uint() >> 9) | 0x3f800000u;
    return x.f - 1.0f;
  }
  __host__ __device__ void advance(uint64_t delta) {
    uint64_t cur_mult = mult(), cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
    while (delta > 0) {
      if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
      cur_plus = (cur_mult + 1) * cur_plus; cur_mult *= cur_mult; delta /= 2;
    }
    state = acc_mult * state + acc_plus;
  }
};

// optimizer hyper-parameters (example-model.json:2-15; defaults tcnn adam.h:288-301,
// exponential_decay.h:137-140)
struct OptimizerConfig {
  bool has_decay = false;
  float decay_base = 0.1f; uint32_t decay_start = 10000, decay_interval = 10000, decay_end = 10000000;
  float lr = 1e-3f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, l2_reg = 1e-8f;
};

struct ModelConfig {
  int n_levels = 16, n_feat = 2, log2_hashmap = 19, base_res = 16;
  float per_level_scale = 2.0f;
  int n_neurons = 64, n_hidden = 4;
  OptimizerConfig opt;
  std::string model_json;      // {"loss":..,"encoding":..,"network":..} as the reference keeps in m_model
  std::string full_json;       // the text the volume was created from
  DecoderDesc desc;            // derived
  size_t n_params() const { return (size_t)desc.n_mlp + desc.n_grid; }
};

ModelConfig parse_model_config(const std::string& json_text);

struct Volume;
struct Renderer;

// ---- kernel launchers (decode.cu, render.cu, macrocell.cu, train.cu) -------------------
int num_sms();
cudaError_t launch_decode(const DecoderDesc& d, const __half* params, const float* coords, float* out, size_t n, __half* enc_out, cudaStream_t stream);
cudaError_t launch_gather_probe(const DecoderDesc& d, const __half* params, const float* coords, uint32_t* out, size_t n, cudaStream_t stream);
cudaError_t launch_decode_samples(const DecoderDesc& d, const __half* params, const float4* samples, const float4* samples_alt, float* out,
                                  const uint32_t* n_dev, const uint32_t* round_dev, size_t n_max, cudaStream_t stream);

}  // namespace vnr

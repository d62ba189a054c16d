// ref_driver.cu -- TEST INFRASTRUCTURE.  A thin extern "C" driver around the reference's OWN
// tiny-cuda-nn sources (compiled in place from /root/reference/tcnn by oracle/ref_driver/Makefile).
// It builds the network and trainer exactly as the reference does in
// core/networks/tcnn_network.h:200-209 (NetworkWithInputEncoding<3 -> 1> + Trainer with the
// model's loss/optimizer) and exposes what NeuralVolume calls on it:
//   tcnn_inference  (core/networks/tcnn_impl.cu:438-448)   -> ref_inference
//   Trainer::training_step (tcnn trainer.h:211-247)        -> ref_training_step
// Used (a) to pin the oracle and the product kernels against the reference's own arithmetic on
// the GPU box and (b) as the "reference" baseline arm of bench.py.  Never linked by the product.
#include <tiny-cuda-nn/common.h>
#include <tiny-cuda-nn/config.h>
#include <tiny-cuda-nn/gpu_matrix.h>
#include <tiny-cuda-nn/loss.h>
#include <tiny-cuda-nn/network_with_input_encoding.h>
#include <tiny-cuda-nn/optimizer.h>
#include <tiny-cuda-nn/trainer.h>

#include <cstdio>
#include <memory>
#include <string>

using namespace tcnn;
using precision_t = network_precision_t;
using json = nlohmann::json;

struct RefNet {
  std::shared_ptr<Loss<precision_t>> loss;
  std::shared_ptr<Optimizer<precision_t>> optimizer;
  std::shared_ptr<NetworkWithInputEncoding<precision_t>> network;
  std::shared_ptr<Trainer<float, precision_t, precision_t>> trainer;
  std::string error;
};

static std::string g_err;

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API const char* ref_last_error() { return g_err.c_str(); }

REF_API void* ref_create(const char* model_json, uint32_t seed) {
  try {
    json config = json::parse(model_json, nullptr, true, true);
    json loss_opts = config.value("loss", json::object());
    json encoding_opts = config.value("encoding", json::object());
    json network_opts = config.value("network", json::object());
    json optimizer_opts = config.value("optimizer", json::object());
    auto* r = new RefNet();
    r->loss.reset(create_loss<precision_t>(loss_opts));
    r->optimizer.reset(create_optimizer<precision_t>(optimizer_opts));
    r->network = std::make_shared<NetworkWithInputEncoding<precision_t>>(3u, 1u, encoding_opts, network_opts);
    r->trainer = std::make_shared<Trainer<float, precision_t, precision_t>>(r->network, r->optimizer, r->loss, seed);
    return r;
  } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

REF_API void ref_destroy(void* h) { delete (RefNet*)h; }

REF_API uint64_t ref_n_params(void* h) { return ((RefNet*)h)->network->n_params(); }
REF_API int ref_precision_bytes() { return (int)sizeof(precision_t); }

REF_API int ref_set_params_f16(void* h, const uint16_t* host, uint64_t n) {
  try { ((RefNet*)h)->trainer->set_params((const precision_t*)host, n); return 0; }
  catch (std::exception& e) { g_err = e.what(); return -1; }
}

REF_API int ref_get_params_f16(void* h, uint16_t* host, uint64_t n) {
  try {
    json data = ((RefNet*)h)->trainer->serialize(false);
    json::binary_t blob = data["params_binary"];
    if (blob.size() != n * sizeof(precision_t)) { g_err = "size mismatch"; return -1; }
    memcpy(host, blob.data(), blob.size());
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// device pointers; n must be a multiple of 256 (NeuralVolume::inference pads, network.cu:1046)
REF_API int ref_inference(void* h, const float* d_xyz, float* d_out, uint32_t n, void* stream) {
  try {
    GPUMatrix<float, CM> input((float*)d_xyz, 3, n);
    GPUMatrix<float, CM> output(d_out, 1, n);
    ((RefNet*)h)->network->inference((cudaStream_t)stream, input, output);
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// one Trainer::training_step; returns the loss through *loss when non-null (host sync, as the
// reference's old-API path does every step, tcnn_network.h:231,246)
REF_API int ref_training_step(void* h, const float* d_xyz, const float* d_target, uint32_t n, void* stream, float* loss) {
  try {
    GPUMatrix<float, CM> input((float*)d_xyz, 3, n);
    GPUMatrix<float, CM> target((float*)d_target, 1, n);
    ((RefNet*)h)->trainer->training_step((cudaStream_t)stream, input, target, loss);
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// Generates tests/golden/params_ref_small.bson with the reference's OWN JSON/BSON serializer
// (nlohmann json.hpp vendored under /root/reference/tcnn/dependencies/json), laid out exactly as
// NeuralVolume::save_params_to_json does (core/network.cu:827-857; parameters as tcnn
// Trainer::serialize, trainer.h:299-311; model as TcnnNetwork::serialize_model, tcnn_network.h:161).
// Build + run (in the container that has /root/reference; nothing of the reference is copied):
//   g++ -std=c++17 -I/root/reference/tcnn/dependencies tools/make_golden_bson.cpp -o /tmp/mkbson && /tmp/mkbson tests/golden/params_ref_small.bson
// Contents are deterministic: dims 40x24x17 -> macrocell 3x2x2; model = the small test config
// (n_levels 2, n_features_per_level 2, log2_hashmap_size 6, base_resolution 4, 1 hidden layer);
// params_binary[i] = uint16(0x2000 + (i * 7) % 0x1c00); macrocell data[k] = float(k) * 0.125f - 1.0f.
#include <json/json.hpp>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <vector>
using json = nlohmann::json;

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s out.bson\n", argv[0]); return 2; }
  const int dims[3] = {40, 24, 17};
  const int mc[3] = {3, 2, 2};
  json model = json::parse(R"({
    "loss": {"otype": "L1"},
    "encoding": {"otype": "HashGrid", "n_levels": 2, "n_features_per_level": 2, "log2_hashmap_size": 6, "base_resolution": 4},
    "network": {"otype": "FullyFusedMLP", "n_neurons": 64, "n_hidden_layers": 1, "activation": "ReLU", "output_activation": "None"}
  })");
  // parameter count of that model: MLP 64*16 + 16*64 = 2048; grid: level 0 res 4 -> 64 entries, level 1 res 8 -> min(512, 64) = 64 entries; (64+64)*2 = 256
  const size_t n_params = 2048 + 256;
  std::vector<uint16_t> params(n_params);
  for (size_t i = 0; i < n_params; ++i) params[i] = (uint16_t)(0x2000 + (i * 7) % 0x1c00);
  std::vector<float> mcdata(2 * mc[0] * mc[1] * mc[2]);
  for (size_t k = 0; k < mcdata.size(); ++k) mcdata[k] = (float)k * 0.125f - 1.0f;

  json::binary_t pb; pb.resize(n_params * 2); memcpy(pb.data(), params.data(), pb.size());
  json::binary_t mb; mb.resize(mcdata.size() * 4); memcpy(mb.data(), mcdata.data(), mb.size());

  json root;
  root["volume"] = {{"dims", {{"x", dims[0]}, {"y", dims[1]}, {"z", dims[2]}}}};
  root["macrocell"] = {
      {"groundtruth", false},
      {"dims", {{"x", mc[0]}, {"y", mc[1]}, {"z", mc[2]}}},
      {"spacings", {{"x", 16.f / dims[0]}, {"y", 16.f / dims[1]}, {"z", 16.f / dims[2]}}},
      {"data", mb},
  };
  json par;
  par["n_params"] = n_params;
  par["params_binary"] = pb;
  root["parameters"] = par;
  root["model"] = model;
  const auto b = json::to_bson(root);
  std::ofstream ofs(argv[1], std::ios::binary);
  ofs.write((const char*)b.data(), b.size());
  printf("wrote %zu bytes\n", b.size());
  return 0;
}

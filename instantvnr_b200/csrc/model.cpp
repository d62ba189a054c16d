// model.cpp -- model-config parsing (example-model.json) and the derived level tables.
// Follows tcnn's GridEncodingTemplated constructor (encodings/grid.h:527-594) for the
// per-level sizes and FullyFusedMLP's constructor (src/fully_fused_mlp.cu:643-703) for the
// matrix shapes; defaults as create_grid_encoding_templated (grid.h:861-881).
#include <cmath>
#include <limits>

#include "mini_json.h"
#include "vnr_host.h"

namespace vnr {

static uint32_t powi_u32(uint32_t b, int e) { uint32_t r = 1; for (int i = 0; i < e; ++i) r *= b; return r; }

static bool ieq(const std::string& a, const char* b) {
  size_t n = strlen(b);
  if (a.size() != n) return false;
  for (size_t i = 0; i < n; ++i) if (tolower(a[i]) != tolower(b[i])) return false;
  return true;
}

ModelConfig parse_model_config(const std::string& text) {
  mj::Value root;
  try { root = mj::Parser::parse(text); } catch (const std::exception& e) { throw InvalidError(e.what()); }
  if (!root.is_object()) throw InvalidError("model config must be a JSON object");
  ModelConfig c;
  c.full_json = text;
  const mj::Value enc = root.value_obj("encoding"), net = root.value_obj("network"), loss = root.value_obj("loss");
  mj::Value opt = root.value_obj("optimizer");

  const std::string etype = enc.value("otype", "Grid");
  if (!(ieq(etype, "HashGrid") || ieq(etype, "Grid"))) throw UnsupportedError("encoding.otype '" + etype + "' is outside the hot path (only HashGrid)");
  if (enc.contains("type") && !ieq(enc.value("type", "Hash"), "Hash")) throw UnsupportedError("only grid type Hash is supported");
  if (enc.contains("interpolation") && !ieq(enc.value("interpolation", "Linear"), "Linear")) throw UnsupportedError("only Linear interpolation is supported");
  c.n_feat = (int)enc.value("n_features_per_level", 2.0);
  if (enc.contains("n_features")) {
    if (enc.contains("n_levels")) throw InvalidError("GridEncoding: may not specify n_features and n_levels simultaneously");
    c.n_levels = (int)enc.at("n_features").num() / c.n_feat;
  } else c.n_levels = (int)enc.value("n_levels", 16.0);
  c.log2_hashmap = (int)enc.value("log2_hashmap_size", 19.0);
  c.base_res = (int)enc.value("base_resolution", 16.0);
  c.per_level_scale = (float)enc.value("per_level_scale", 2.0);
  if (!(c.n_feat == 1 || c.n_feat == 2 || c.n_feat == 4 || c.n_feat == 8)) throw InvalidError("GridEncoding: n_features_per_level must be 1, 2, 4, or 8.");
  if (c.n_levels < 1 || c.n_levels > kMaxLevels) throw UnsupportedError("n_levels must be in [1,16]");
  if (c.log2_hashmap < 4 || c.log2_hashmap > 28) throw InvalidError("log2_hashmap_size out of range");

  const std::string ntype = net.value("otype", "FullyFusedMLP");
  if (!ieq(ntype, "FullyFusedMLP")) throw UnsupportedError("network.otype '" + ntype + "' is outside the hot path (only FullyFusedMLP)");
  if (!ieq(net.value("activation", "ReLU"), "ReLU")) throw UnsupportedError("only ReLU hidden activation is supported");
  if (!ieq(net.value("output_activation", "None"), "None")) throw UnsupportedError("only output_activation None is supported");
  c.n_neurons = (int)net.value("n_neurons", 128.0);
  c.n_hidden = (int)net.value("n_hidden_layers", 5.0);
  if (c.n_neurons != kWidth) throw UnsupportedError("only n_neurons = 64 is supported");
  if (c.n_hidden < 1 || c.n_hidden > kMaxHidden) throw UnsupportedError("n_hidden_layers must be in [1,8]");
  const std::string ltype = loss.value("otype", "L1");
  if (!ieq(ltype, "L1")) throw UnsupportedError("only the L1 loss is supported");

  // optimizer: ExponentialDecay(nested Adam) or plain Adam
  std::string otype = opt.value("otype", "Adam");
  if (ieq(otype, "ExponentialDecay")) {
    c.opt.has_decay = true;
    c.opt.decay_base = (float)opt.value("decay_base", 0.1);
    c.opt.decay_start = (uint32_t)opt.value("decay_start", 10000.0);
    c.opt.decay_interval = (uint32_t)opt.value("decay_interval", 10000.0);
    c.opt.decay_end = (uint32_t)opt.value("decay_end", 10000000.0);
    opt = opt.value_obj("nested");
    otype = opt.value("otype", "Adam");
  }
  if (!ieq(otype, "Adam")) throw UnsupportedError("optimizer '" + otype + "' is outside the hot path (only Adam / ExponentialDecay(Adam))");
  c.opt.lr = (float)opt.value("learning_rate", 1e-3);
  c.opt.beta1 = (float)opt.value("beta1", 0.9);
  c.opt.beta2 = (float)opt.value("beta2", 0.999);
  c.opt.eps = (float)opt.value("epsilon", 1e-8);
  c.opt.l2_reg = (float)opt.value("l2_reg", 1e-8);

  // m_model as the reference keeps it (tcnn_network.h:172-175): loss / encoding / network
  {
    mj::Value m = mj::Value::make_object();
    m.set("loss", loss); m.set("encoding", enc); m.set("network", net);
    mj::dump(m, c.model_json);
  }

  // derived tables
  DecoderDesc& d = c.desc;
  memset(&d, 0, sizeof d);
  d.n_levels = c.n_levels; d.n_feat = c.n_feat; d.n_hidden = c.n_hidden;
  d.enc_dims = c.n_levels * c.n_feat;
  d.enc_pad = ((d.enc_dims + 15) / 16) * 16;
  if (d.enc_pad > 64) throw UnsupportedError("n_levels * n_features_per_level must be <= 64");
  uint32_t offset = 0;
  for (int i = 0; i < c.n_levels; ++i) {
    const float scale = exp2f(i * std::log2(c.per_level_scale)) * c.base_res - 1.0f;
    const uint32_t resolution = (uint32_t)(ceilf(scale)) + 1;
    const uint32_t max_params = std::numeric_limits<uint32_t>::max() / 2;
    uint32_t params_in_level = std::pow((float)resolution, 3) > (float)max_params ? max_params : powi_u32(resolution, 3);
    params_in_level = ((params_in_level + 7u) / 8u) * 8u;
    params_in_level = std::min(params_in_level, (1u << c.log2_hashmap));
    LevelDesc& lv = d.lv[i];
    lv.offset = offset; lv.size = params_in_level; lv.res = resolution; lv.res2 = resolution * resolution; lv.scale = scale;
    // grid_index (grid.h:82-99): dims are accumulated while stride <= size; hashed iff size < final stride
    uint32_t stride = 1; int ndim = 0;
    for (int dim = 0; dim < 3 && stride <= params_in_level; ++dim) { stride *= resolution; ++ndim; }
    lv.hashed = params_in_level < stride ? 1u : 0u;
    if (!lv.hashed && ndim != 3) throw UnsupportedError("degenerate dense level (uint32 stride overflow)");
    lv.mask = (params_in_level & (params_in_level - 1)) == 0 ? params_in_level - 1 : 0u;
    offset += params_in_level;
  }
  d.n_grid = offset * (uint32_t)c.n_feat;
  d.n_mlp = (uint32_t)(kWidth * d.enc_pad + (c.n_hidden - 1) * kWidth * kWidth + kOutPad * kWidth);
  return c;
}

}  // namespace vnr

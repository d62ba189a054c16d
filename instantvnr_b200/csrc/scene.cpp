// scene.cpp -- scene descriptions (VIDI3D / DIVA JSON) -> what the hot path needs from them: the raw volume
// file(s), their layout and value range, the camera and the transfer function's value range.
// Restates serializer.cpp:138-477 (create_json_{scene,volume,camera,tfn}_stringify) on mini_json.
//
// Not restated: the transfer-function TABLE.  serializer.cpp:189-211 hands root.view.volume.transferFunction to
// tfn::loadTransferFunction / tfn::TransferFunctionCore of the un-vendored OVR tfn module; its file format
// is not visible in the reference.  vnr_scene_tfn() therefore accepts only an explicit table (see the header) and
// reports VNR_ERR_UNSUPPORTED for anything else; colours/alphas go in through the setters, as batch_renderer does
// for the range.
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>

#include "mini_json.h"
#include "scene.h"

namespace vnr {

static const mj::Value& need(const mj::Value& o, const char* key) {
  if (!o.is_object()) throw InvalidError("has to be a JSON object");                    // serializer.cpp:81
  if (!o.contains(key)) throw InvalidError(std::string("incorrect key: ") + key);       // :82
  return o.at(key);
}

// NLOHMANN_JSON_SERIALIZE_ENUM(ValueType, ...) serializer.cpp:26-35: an unknown name maps to the FIRST pair (INT8)
static int value_type_of(const mj::Value& v) {
  static const struct { const char* name; int type; } names[] = {
      {"BYTE", 1}, {"UNSIGNED_BYTE", 0}, {"SHORT", 3}, {"UNSIGNED_SHORT", 2}, {"INT", 5}, {"UNSIGNED_INT", 4}, {"FLOAT", 8}, {"DOUBLE", 12}};
  if (v.is_string())
    for (auto& n : names) if (v.s == n.name) return n.type;
  return 1;
}

static void vec3_of(const mj::Value& v, double* out) {               // NLOHMANN_DEFINE_TYPE_NON_INTRUSIVE(vec3*, x, y, z) :44
  out[0] = need(v, "x").num(); out[1] = need(v, "y").num(); out[2] = need(v, "z").num();
}

static bool file_exists(const std::string& name) { std::ifstream f(name.c_str()); return f.good(); }

// valid_filename serializer.cpp:116-134: an array lists candidates, the first that exists wins
static std::string valid_filename(const mj::Value& in, const char* key) {
  if (!in.contains(key)) throw InvalidError("Json key 'fileName' doesnot exist");
  const mj::Value& js = in.at(key);
  if (js.type == mj::Value::ArrayT) {
    for (auto& s : *js.arr) if (s.is_string() && file_exists(s.s)) return s.s;
    throw InvalidError("Cannot find volume file.");
  }
  if (!js.is_string()) throw InvalidError("json: fileName is not a string");
  return js.s;
}

static double type_max(int type) {                                   // create_scene_vidi__tfn :221-256
  switch (type) {
    case 0: return std::numeric_limits<uint8_t>::max();
    case 1: return std::numeric_limits<int8_t>::max();
    case 2: return std::numeric_limits<uint16_t>::max();
    case 3: return std::numeric_limits<int16_t>::max();
    case 4: return std::numeric_limits<uint32_t>::max();
    case 5: return std::numeric_limits<int32_t>::max();
    case 8: case 12: return 1.0;
    default: throw InvalidError("unknown data type");
  }
}

// create_scene_vidi__volume / __multivolume :261-318
static SceneFile vidi_file(const mj::Value& jsdata, int* dims, int* type) {
  const std::string format = need(jsdata, "format").is_string() ? jsdata.at("format").s : std::string();
  if (format != "REGULAR_GRID_RAW_BINARY") throw UnsupportedError("data type unimplemented");
  SceneFile f;
  f.filename = valid_filename(jsdata, "fileName");
  double d[3]; vec3_of(need(jsdata, "dimensions"), d);
  const int t = value_type_of(need(jsdata, "type"));
  f.offset = jsdata.contains("offset") ? (uint64_t)jsdata.at("offset").num() : 0;
  f.big_endian = jsdata.contains("endian") && jsdata.at("endian").is_string() && jsdata.at("endian").s == "BIG_ENDIAN";
  if (dims[0] < 0) { dims[0] = (int)d[0]; dims[1] = (int)d[1]; dims[2] = (int)d[2]; *type = t; }
  else if (dims[0] != (int)d[0] || dims[1] != (int)d[1] || dims[2] != (int)d[2] || *type != t)
    throw InvalidError("every dataSource entry must have the dimensions and type of the first one");   // the asserts at :303-304
  return f;
}

static void vidi_range(const mj::Value& jsvolume, int type, Scene& s) {      // create_scene_vidi__tfn :213-259
  auto range_of = [](const mj::Value& r, double* out) {                       // rangeFromJson :98-108
    if (!r.contains("minimum") || !r.contains("maximum")) { out[0] = out[1] = 0.0; return; }
    out[0] = (float)r.at("minimum").num(); out[1] = (float)r.at("maximum").num();
  };
  double r[2];
  if (jsvolume.contains("scalarMappingRangeUnnormalized")) {
    range_of(jsvolume.at("scalarMappingRangeUnnormalized"), r);
    s.range[0] = (float)r[0]; s.range[1] = (float)r[1]; s.has_range = true;
  } else if (jsvolume.contains("scalarMappingRange")) {
    range_of(jsvolume.at("scalarMappingRange"), r);
    const double m = type_max(type);
    s.range[0] = (float)(m * (float)r[0]); s.range[1] = (float)(m * (float)r[1]); s.has_range = true;
  }
  // else: the range is taken from the data (range1f default = empty -> StaticSampler::load computes min/max)
}

static void explicit_tfn(const mj::Value& jstfn, Scene& s) {
  // our own explicit form: {"colors": [[r,g,b], ...], "alphas": [[position, alpha], ...]} -- see the file header
  if (!jstfn.is_object() || !jstfn.contains("colors") || !jstfn.contains("alphas")) return;
  const mj::Value &c = jstfn.at("colors"), &a = jstfn.at("alphas");
  if (c.type != mj::Value::ArrayT || a.type != mj::Value::ArrayT) return;
  for (auto& e : *c.arr) {
    if (e.type != mj::Value::ArrayT || e.arr->size() != 3) throw InvalidError("transferFunction.colors: expecting [r, g, b] triples");
    for (auto& x : *e.arr) s.tfn_color.push_back((float)x.num());
  }
  for (auto& e : *a.arr) {
    if (e.type != mj::Value::ArrayT || e.arr->size() != 2) throw InvalidError("transferFunction.alphas: expecting [position, alpha] pairs");
    for (auto& x : *e.arr) s.tfn_alpha.push_back((float)x.num());
  }
  s.has_tfn = !s.tfn_color.empty() && !s.tfn_alpha.empty();
}

Scene parse_scene(const std::string& text) {
  mj::Value root;
  try { root = mj::Parser::parse(text); } catch (const std::exception& e) { throw InvalidError(e.what()); }
  if (!root.is_object()) throw InvalidError("scene description must be a JSON object");
  Scene s;
  std::string version;
  if (root.contains("version")) {                                   // create_json_*_stringify :420-477
    version = root.at("version").is_string() ? root.at("version").s : std::string("?");
    if (version != "DIVA" && version != "VIDI3D") throw InvalidError("unknown JSON configuration format");
  }
  try {
    if (version == "DIVA") {
      // create_json_volume_stringify_diva :138-168; camera and transfer function are "TODO" in the reference (:174)
      const mj::Value& config = need(root, "volume");
      const mj::Value& rg = need(config, "range");                    // parsed as vec2f {x, y}
      s.range[0] = (float)need(rg, "x").num(); s.range[1] = (float)need(rg, "y").num(); s.has_range = true;
      double d[3]; vec3_of(need(config, "dims"), d);
      s.dims[0] = (int)d[0]; s.dims[1] = (int)d[1]; s.dims[2] = (int)d[2];
      s.value_type = value_type_of(need(config, "type"));
      const bool big = config.contains("bigendian") && config.at("bigendian").num() != 0;
      const mj::Value& fn = need(config, "filename");
      auto add = [&](const mj::Value& v) {
        if (!v.is_string()) throw InvalidError("json: filename is not a string");
        SceneFile f; f.filename = v.s; f.big_endian = big; f.offset = 0; s.files.push_back(f);
      };
      if (fn.type == mj::Value::ArrayT) for (auto& v : *fn.arr) add(v); else add(fn);
    } else {
      const mj::Value& ds = need(root, "dataSource");
      if (ds.type != mj::Value::ArrayT) throw InvalidError("'dataSource' is expected to be an array");
      if (ds.arr->empty()) throw InvalidError("'dataSource' should contain at least one element");
      // one entry per time step.  (create_json_volume_stringify_vidi :381-388 resizes the file list to ds.size() and then
      // push_backs entries 1.., leaving empty descriptors in between; the intent -- file i = dataSource[i] -- is kept.)
      s.dims[0] = -1;
      for (auto& e : *ds.arr) s.files.push_back(vidi_file(e, s.dims, &s.value_type));
      if (root.contains("view") && root.at("view").is_object()) {
        const mj::Value& view = root.at("view");
        if (view.contains("volume") && view.at("volume").is_object()) {
          const mj::Value& jv = view.at("volume");
          vidi_range(jv, s.value_type, s);
          s.unnormalized_range_missing = !jv.contains("scalarMappingRangeUnnormalized") && s.value_type != 8 && s.value_type != 12;   // the warning at :359-364
          if (jv.contains("transferFunction")) { s.tfn_present = true; explicit_tfn(jv.at("transferFunction"), s); }
        }
        if (view.contains("camera") && view.at("camera").is_object()) {
          // create_scene_vidi__camera :177-187 + the shift into the centred world box :367-369
          const mj::Value& jc = view.at("camera");
          double e[3], c[3], u[3];
          vec3_of(need(jc, "eye"), e); vec3_of(need(jc, "center"), c); vec3_of(need(jc, "up"), u);
          for (int k = 0; k < 3; ++k) {
            const float half = (float)s.dims[k] / 2.f;
            s.cam_from[k] = (float)e[k] - half; s.cam_at[k] = (float)c[k] - half; s.cam_up[k] = (float)u[k];
          }
          s.fovy = (float)need(jc, "fovy").num();
          s.has_camera = true;
        }
      }
    }
  } catch (const InvalidError&) { throw; } catch (const UnsupportedError&) { throw; } catch (const std::exception& e) { throw InvalidError(e.what()); }
  if (s.dims[0] <= 0 || s.dims[1] <= 0 || s.dims[2] <= 0) throw InvalidError("volume dimensions must be positive");
  return s;
}

Scene load_scene(const std::string& path) {
  std::ifstream f(path.c_str(), std::ios::binary);
  if (!f) throw InvalidError("cannot open scene file " + path);
  std::ostringstream ss; ss << f.rdbuf();
  Scene s = parse_scene(ss.str());
  // file names relative to the scene file resolve against its directory when they do not exist as given
  const size_t slash = path.find_last_of('/');
  if (slash != std::string::npos) {
    const std::string dir = path.substr(0, slash + 1);
    for (auto& fl : s.files)
      if (!fl.filename.empty() && fl.filename[0] != '/' && !file_exists(fl.filename) && file_exists(dir + fl.filename)) fl.filename = dir + fl.filename;
  }
  return s;
}

}  // namespace vnr

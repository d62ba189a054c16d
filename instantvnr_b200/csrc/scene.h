// scene.h -- parsed scene description (scene.cpp; serializer.cpp:138-477 of the reference)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "vnr_host.h"

namespace vnr {

struct SceneFile { std::string filename; uint64_t offset = 0; bool big_endian = false; };

struct Scene {
  int dims[3] = {0, 0, 0};
  int value_type = 8;                       // ValueType, core/mathdef.h:51-65
  std::vector<SceneFile> files;             // one per time step (MultiVolume::data)
  bool has_range = false;                   // false: take the value range from the data (range1f empty)
  float range[2] = {0.f, 0.f};              // unnormalised value range = the transfer function's range
  bool unnormalized_range_missing = false;  // integer volume without scalarMappingRangeUnnormalized (the reference warns)
  bool has_camera = false;
  float cam_from[3] = {0, 0, -1}, cam_at[3] = {0, 0, 0}, cam_up[3] = {0, 1, 0}, fovy = 60.f;
  bool tfn_present = false, has_tfn = false;
  std::vector<float> tfn_color, tfn_alpha;  // rgb triples; (position, alpha) pairs
};

Scene parse_scene(const std::string& json_text);
Scene load_scene(const std::string& path);

}  // namespace vnr

// vnr_api_check -- every function of the reference's api.h exercised through include/vnr_api.hpp.
//   vnr_api_check --host  <tmpdir>   host-only part (no CUDA device needed): scene ingest, camera, transfer function,
//                                    vnrRequireDecoding, handle-type errors
//   vnr_api_check --device <tmpdir>  the whole flow on cuda:0: a two-time-step uint16 scene written to <tmpdir> ->
//                                    vnrCreateSimpleVolume(scene, "GPU") -> vnrCreateNeuralVolume(config, simple) ->
//                                    train / evaluate / serialize / reload / SetModel / render both volume kinds
// Prints one "key value" line per checked quantity (tests/test_api_app.py reads them); exit code 0 = all checks passed.
#include <cstring>
#include <iostream>

#include "synthetic.hpp"

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::cerr << "CHECK failed: " #cond " (" << __FILE__ << ":" << __LINE__ << ")" << std::endl; ++failures; } } while (0)
template <typename F> static bool throws(F f) { try { f(); } catch (const std::exception&) { return true; } return false; }

static const char* kModel =
    "{\"encoding\":{\"otype\":\"HashGrid\",\"n_levels\":8,\"n_features_per_level\":8,\"log2_hashmap_size\":15,\"base_resolution\":8,\"per_level_scale\":2.0},"
    "\"network\":{\"otype\":\"FullyFusedMLP\",\"n_neurons\":64,\"n_hidden_layers\":2,\"activation\":\"ReLU\",\"output_activation\":\"None\"},"
    "\"loss\":{\"otype\":\"L1\"},\"optimizer\":{\"otype\":\"Adam\",\"learning_rate\":0.01}}";
static const char* kModelSmall =
    "{\"encoding\":{\"otype\":\"HashGrid\",\"n_levels\":4,\"n_features_per_level\":4,\"log2_hashmap_size\":12,\"base_resolution\":8},"
    "\"network\":{\"otype\":\"FullyFusedMLP\",\"n_neurons\":64,\"n_hidden_layers\":1,\"activation\":\"ReLU\",\"output_activation\":\"None\"},"
    "\"loss\":{\"otype\":\"L1\"},\"optimizer\":{\"otype\":\"Adam\"}}";

// two time steps of the synthetic volume as uint16 raw files + a VIDI3D scene file naming them
static std::string write_scene(const std::string& dir, vnr::vec3i dims) {
  for (int t = 0; t < 2; ++t) {
    const std::vector<float> v = synthetic::make_volume(dims, 42 + t);
    std::vector<uint16_t> q(v.size());
    for (size_t i = 0; i < v.size(); ++i) q[i] = (uint16_t)(v[i] * 60000.f + 0.5f);
    std::ofstream f(dir + "/step" + std::to_string(t) + ".raw", std::ios::binary);
    f.write((const char*)q.data(), (std::streamsize)(q.size() * 2));
  }
  std::ostringstream js;
  js << "// written by vnr_api_check\n{\"version\":\"VIDI3D\",\"dataSource\":[";
  for (int t = 0; t < 2; ++t)
    js << (t ? "," : "") << "{\"format\":\"REGULAR_GRID_RAW_BINARY\",\"fileName\":\"step" << t << ".raw\",\"dimensions\":{\"x\":" << dims.x << ",\"y\":" << dims.y
       << ",\"z\":" << dims.z << "},\"type\":\"UNSIGNED_SHORT\"}";
  js << "],\"view\":{\"volume\":{\"scalarMappingRangeUnnormalized\":{\"minimum\":0.0,\"maximum\":60000.0},"
        "\"transferFunction\":{\"colors\":[[0,0,1],[1,1,1],[1,0,0]],\"alphas\":[[0.0,0.0],[0.3,0.0],[1.0,0.8]]}},"
        "\"camera\":{\"eye\":{\"x\":" << dims.x / 2.0 << ",\"y\":" << dims.y / 2.0 << ",\"z\":" << -1.5 * dims.z << "},\"center\":{\"x\":" << dims.x / 2.0
     << ",\"y\":" << dims.y / 2.0 << ",\"z\":" << dims.z / 2.0 << "},\"up\":{\"x\":0,\"y\":1,\"z\":0},\"fovy\":50.0}}}\n";
  const std::string path = dir + "/scene.json";
  std::ofstream(path) << js.str();
  return path;
}

static double mean_alpha(const vnr::vec4f* p, int n) { double a = 0; for (int i = 0; i < n; ++i) a += p[i].w; return a / n; }
static double max_diff(const std::vector<vnr::vec4f>& a, const vnr::vec4f* b) {
  double m = 0;
  for (size_t i = 0; i < a.size(); ++i) m = std::max({m, (double)std::fabs(a[i].x - b[i].x), (double)std::fabs(a[i].y - b[i].y), (double)std::fabs(a[i].z - b[i].z), (double)std::fabs(a[i].w - b[i].w)});
  return m;
}

static void host_part(const std::string& dir) {
  const vnr::vec3i dims(48, 40, 32);
  const std::string scene_path = write_scene(dir, dims);
  vnrJson scene = vnrJson(scene_path);            // a vnrJson that is_string() is a file name (api.cpp:75-80)

  // modes (api.h:62-87)
  CHECK(vnrRequireDecoding(VNR_OPTIX_NO_SHADING) && vnrRequireDecoding(VNR_RAYMARCHING_NO_SHADING_DECODING) && vnrRequireDecoding(VNR_PATHTRACING_DECODING));
  CHECK(!vnrRequireDecoding(VNR_RAYMARCHING_NO_SHADING_SAMPLE_STREAMING) && !vnrRequireDecoding(VNR_RAYMARCHING_SINGLE_SHADE_HEURISTIC_IN_SHADER));
  CHECK(throws([] { vnrRequireDecoding(VNR_INVALID); }));

  // camera
  vnrCamera cam = vnrCreateCamera(scene);
  CHECK(vnrCameraGetPosition(cam).z == -1.5f * dims.z - dims.z / 2.f && vnrCameraGetFocus(cam).x == 0.f && vnrCameraGetUpVec(cam).y == 1.f && cam->fovy == 50.f);
  vnrCamera cam2 = vnrCreateCamera();
  vnrCameraSet(cam2, scene);
  CHECK(vnrCameraGetPosition(cam2).z == vnrCameraGetPosition(cam).z);
  vnrCameraSet(cam2, vnr::vec3f(1, 2, 3), vnr::vec3f(0, 0, 0), vnr::vec3f(0, 0, 1));
  CHECK(vnrCameraGetPosition(cam2).y == 2.f && vnrCameraGetUpVec(cam2).z == 1.f);

  // transfer function
  vnrTransferFunction tfn = vnrCreateTransferFunction(scene);
  CHECK(vnrTransferFunctionGetColor(tfn).size() == 3 && vnrTransferFunctionGetAlpha(tfn).size() == 3);
  CHECK(vnrTransferFunctionGetValueRange(tfn).lo == 0.f && vnrTransferFunctionGetValueRange(tfn).hi == 60000.f);
  vnrTransferFunction t2 = vnrCreateTransferFunction();
  vnrTransferFunctionSetColor(t2, {vnr::vec3f(1, 0, 0)});
  vnrTransferFunctionSetAlpha(t2, {vnr::vec2f(0, 0), vnr::vec2f(1, 1)});
  vnrTransferFunctionSetValueRange(t2, vnr::range1f(0.25f, 0.75f));
  CHECK(vnrTransferFunctionGetColor(t2).size() == 1 && vnrTransferFunctionGetAlpha(t2)[1].y == 1.f && vnrTransferFunctionGetValueRange(t2).hi == 0.75f);

  // simple volume from the scene: descriptor only, nothing touches the device until it is trained on / rendered
  vnrVolume simple = vnrCreateSimpleVolume(scene, "GPU");
  CHECK(simple->dims.x == dims.x && simple->dims.z == dims.z && !simple->isNetwork());
  CHECK(vnrSimpleVolumeGetNumberOfTimeSteps(simple) == 2);
  CHECK(throws([&] { vnrSimpleVolumeSetCurrentTimeStep(simple, 2); }));
  CHECK(throws([&] { vnrNeuralVolumeTrain(simple, 1, true); }));                       // "expecting a neural volume" (api.cpp:125-131)
  CHECK(throws([&] { vnrCreateSimpleVolume(vnr::jx::from_text("{\"version\":\"nope\"}"), "GPU"); }));
  CHECK(vnrVolumeGetValueRange(simple).lo == 0.f && vnrVolumeGetValueRange(simple).hi == 1.f);
  // clipping box in voxel units -> unit cube through the inverse data transform (api.cpp:330-348)
  vnrVolumeSetClippingBox(simple, vnr::vec3f(12, 0, 8), vnr::vec3f(48, 20, 32));
  CHECK(std::fabs(simple->clip_lo.x - 0.25f) < 1e-6f && std::fabs(simple->clip_hi.y - 0.5f) < 1e-6f && std::fabs(simple->clip_lo.z - 0.25f) < 1e-6f);
  vnrVolumeSetScaling(simple, vnr::vec3f(2, 1, 1));
  vnrVolumeSetClippingBox(simple, vnr::vec3f(24, 0, 0), vnr::vec3f(72, 40, 32));       // world box is now 96 wide, centred
  CHECK(std::fabs(simple->clip_lo.x - 0.5f) < 1e-6f && std::fabs(simple->clip_hi.x - 1.0f) < 1e-6f);

  // json helpers
  vnrJson j = vnrCreateJsonText(scene_path), j2;
  vnrLoadJsonText(j2, scene_path);
#ifdef VNR_API_HAS_NLOHMANN
  CHECK(j == j2 && !j.is_string());                                                   // vnrJson is nlohmann::json, as in api.h
  vnrSaveJsonText(j, dir + "/copy.json");
  CHECK(vnrCreateJsonText(dir + "/copy.json") == j);
  vnrSaveJsonBinary(vnr::jx::from_bson(std::string("\x05\x00\x00\x00\x00", 5)), dir + "/empty.bson");
  CHECK(vnrCreateJsonBinary(dir + "/empty.bson").empty());
#else
  CHECK(j.data == j2.data && !j.is_string());
  vnrSaveJsonText(j, dir + "/copy.json");
  CHECK(vnrCreateJsonText(dir + "/copy.json").data.substr(0, j.data.size()) == j.data);
  vnrSaveJsonBinary(vnr::jx::from_bson(std::string("\x05\x00\x00\x00\x00", 5)), dir + "/empty.bson");
  CHECK(vnrCreateJsonBinary(dir + "/empty.bson").data.size() == 5);
#endif
  vnrLoadJsonBinary(j2, dir + "/empty.bson");
#ifndef VNR_API_HAS_NLOHMANN
  CHECK(j2.kind == vnrJson::Binary);
#endif
  CHECK(throws([&] { vnrCreateNeuralVolume(j2); }));                                   // not a params.json
  vnrRelease(nullptr);
  std::cout << "host_checks_failed " << failures << std::endl;
}

static void device_part(const std::string& dir) {
  const vnr::vec3i dims(48, 40, 32);
  vnrJson scene = vnrJson(write_scene(dir, dims));
  vnrVolume simple = vnrCreateSimpleVolume(scene, "GPU");
  vnrVolume neural = vnrCreateNeuralVolume(vnr::jx::from_text(kModel), simple, /*online_macrocell_construction=*/true, /*seed=*/3);
  CHECK(neural->isNetwork() && neural->dims.y == dims.y);

  // train + evaluators (api.h:129-136)
  CHECK(vnrNeuralVolumeGetTrainingStep(neural) == 0);
  for (int i = 0; i < 30; ++i) vnrNeuralVolumeTrain(neural, 10, /*fast_mode=*/false);
  const double loss = vnrNeuralVolumeGetTrainingLoss(neural), test = vnrNeuralVolumeGetTestingLoss(neural);
  const double psnr = vnrNeuralVolumeGetPSNR(neural, false), ssim = vnrNeuralVolumeGetSSIM(neural, false);
  std::cout << "train_step " << vnrNeuralVolumeGetTrainingStep(neural) << "\ntrain_loss " << loss << "\ntest_loss " << test << "\npsnr " << psnr << "\nssim " << ssim << std::endl;
  CHECK(vnrNeuralVolumeGetTrainingStep(neural) == 300 && loss > 0 && loss < 0.05 && test > 0 && test < 0.05 && psnr > 25.0 && ssim > 0.5 && ssim <= 1.0);

  // progressive decode + exports (api.h:137-140)
  const int blobs = vnrNeuralVolumeGetNumberOfBlobs(neural);
  CHECK(blobs == (dims.z + 15) / 16);
  for (int i = 0; i < blobs; ++i) vnrNeuralVolumeDecodeProgressive(neural);
  vnrNeuralVolumeDecodeInference(neural, dir + "/inference.raw");
  vnrNeuralVolumeDecodeReference(neural, dir + "/reference.raw");
  const std::string inf = vnr::read_file(dir + "/inference.raw", true), ref = vnr::read_file(dir + "/reference.raw", true);
  const size_t rec = (((size_t)dims.x * dims.y + 255) / 256) * 256 * sizeof(float) * dims.z;
  CHECK(inf.size() == rec && ref.size() == rec);
  {
    double mae = 0; const float* a = (const float*)inf.data(); const float* b = (const float*)ref.data();
    for (size_t i = 0; i < (size_t)dims.x * dims.y; ++i) mae += std::fabs(a[i] - b[i]);
    std::cout << "slice0_mae " << mae / (dims.x * dims.y) << std::endl;
    CHECK(mae / (dims.x * dims.y) < 0.05);
  }

  // renderers on both volume kinds (api.h:168-178): same camera / transfer function
  vnrCamera cam = vnrCreateCamera(scene);
  vnrTransferFunction tfn = vnrCreateTransferFunction(scene);
  vnrTransferFunctionSetValueRange(tfn, vnr::range1f(0, 1));                           // batch_renderer.cpp:194
  const vnr::vec2i fb(96, 64);
  auto setup = [&](vnrVolume v, int mode) {
    vnrRenderer r = vnrCreateRenderer(v);
    vnrRendererSetTransferFunction(r, tfn);
    vnrRendererSetCamera(r, cam);
    vnrRendererSetFramebufferSize(r, fb);
    vnrRendererSetMode(r, mode);
    vnrRendererSetDenoiser(r, false);
    vnrRendererSetVolumeDensityScale(r, 1.f);
    vnrRendererSetVolumeSamplingRate(r, 1.f);
    return r;
  };
  vnrRenderer rn = setup(neural, VNR_RAYMARCHING_NO_SHADING_SAMPLE_STREAMING), rs = setup(simple, VNR_RAYMARCHING_NO_SHADING_SAMPLE_STREAMING);
  vnrRender(rn); vnrRender(rs);
  const vnr::vec4f* pn = vnrRendererMapFrame(rn);
  std::vector<vnr::vec4f> frame_n(pn, pn + fb.x * fb.y);
  const vnr::vec4f* ps = vnrRendererMapFrame(rs);
  std::vector<vnr::vec4f> frame_s(ps, ps + fb.x * fb.y);
  const double cover_n = mean_alpha(frame_n.data(), fb.x * fb.y), cover_s = mean_alpha(ps, fb.x * fb.y), diff = max_diff(frame_n, ps);
  std::cout << "coverage_neural " << cover_n << "\ncoverage_simple " << cover_s << "\nneural_vs_simple_max_abs " << diff << std::endl;
  CHECK(cover_s > 0.02 && std::fabs(cover_n - cover_s) < 0.05 && diff < 0.5);
  // accumulation: a second frame averages a new jitter into the first; reset starts over and reproduces frame 1
  vnrRender(rn);
  vnrRendererResetAccumulation(rn);
  vnrRender(rn);
  CHECK(max_diff(frame_n, vnrRendererMapFrame(rn)) == 0.0);
  // every marching mode renders
  for (int mode = VNR_RAYMARCHING_NO_SHADING_DECODING; mode <= VNR_RAYMARCHING_SINGLE_SHADE_HEURISTIC_IN_SHADER; ++mode) {
    vnrRendererSetMode(rn, mode);
    vnrRender(rn);
    const double c = mean_alpha(vnrRendererMapFrame(rn), fb.x * fb.y);
    std::cout << "coverage_mode_" << mode << " " << c << std::endl;
    CHECK(std::fabs(c - cover_n) < 0.05);
  }
  // path tracing: every pixel is written with alpha 1 (writePixelColor(vec4f(L, 1))), scattered light is positive
  for (int mode = VNR_PATHTRACING_DECODING; mode <= VNR_PATHTRACING_IN_SHADER; ++mode) {
    vnrRendererSetMode(rn, mode);
    vnrRender(rn);
    const vnr::vec4f* p = vnrRendererMapFrame(rn);
    double light = 0; for (int i = 0; i < fb.x * fb.y; ++i) light += p[i].x + p[i].y + p[i].z;
    std::cout << "pathtracing_light_mode_" << mode << " " << light / (fb.x * fb.y) << std::endl;
    CHECK(mean_alpha(p, fb.x * fb.y) == 1.0 && light > 0);
  }
  vnrRendererSetMode(rn, VNR_OPTIX_NO_SHADING);
  CHECK(throws([&] { vnrRender(rn); }));
  vnrRendererSetMode(rn, VNR_RAYMARCHING_NO_SHADING_SAMPLE_STREAMING);
  CHECK(throws([&] { vnrRendererSetMode(rn, 99); }));

  // clipping box (read at renderer creation, api.cpp:454) and scaling
  vnrVolumeSetClippingBox(simple, vnr::vec3f(0, 0, 0), vnr::vec3f(dims.x / 2.f, (float)dims.y, (float)dims.z));
  vnrRenderer rc = setup(simple, VNR_RAYMARCHING_NO_SHADING_SAMPLE_STREAMING);
  vnrRender(rc);
  const double cover_clip = mean_alpha(vnrRendererMapFrame(rc), fb.x * fb.y);
  vnrVolumeSetClippingBox(simple, vnr::vec3f(0, 0, 0), vnr::vec3f((float)dims.x, (float)dims.y, (float)dims.z));
  vnrVolumeSetScaling(simple, vnr::vec3f(0.5f, 0.5f, 0.5f));
  vnrRenderer rh = setup(simple, VNR_RAYMARCHING_NO_SHADING_SAMPLE_STREAMING);
  vnrRender(rh);
  const double cover_half = mean_alpha(vnrRendererMapFrame(rh), fb.x * fb.y);
  std::cout << "coverage_clipped " << cover_clip << "\ncoverage_half_scale " << cover_half << std::endl;
  CHECK(cover_clip > 0 && cover_clip < cover_s && cover_half > 0 && cover_half < 0.6 * cover_s);
  vnrVolumeSetScaling(simple, vnr::vec3f(2.f, 2.f, 2.f));

  // time steps: the carrier reloads, the frame changes
  vnrSimpleVolumeSetCurrentTimeStep(simple, 1);
  vnrRendererResetAccumulation(rs);
  vnrRender(rs);
  const double step_diff = max_diff(frame_s, vnrRendererMapFrame(rs));
  std::cout << "timestep_frame_max_abs " << step_diff << std::endl;
  CHECK(step_diff > 0.01);

  // params.json round trip (api.h:124,127,142-143): same decoded frame from the reloaded volume
  vnrNeuralVolumeSerializeParams(neural, dir + "/params.json");
  vnrJson blob;
  vnrNeuralVolumeSerializeParams(neural, blob);
  CHECK(vnr::jx::blob_of(blob) == vnr::read_file(dir + "/params.json", true));     // (nlohmann back end: to_bson of from_bson is the same bytes)
  vnrVolume reloaded = vnrCreateNeuralVolume(vnrJson(dir + "/params.json"));
  CHECK(reloaded->dims.x == dims.x && reloaded->dims.y == dims.y && reloaded->dims.z == dims.z);
  vnrRenderer rr = setup(reloaded, VNR_RAYMARCHING_NO_SHADING_SAMPLE_STREAMING);
  vnrRender(rr);
  const double reload_diff = max_diff(frame_n, vnrRendererMapFrame(rr));
  std::cout << "reloaded_frame_max_abs " << reload_diff << std::endl;
  CHECK(reload_diff == 0.0);
  CHECK(throws([&] { vnrNeuralVolumeSetParams(neural, vnr::jx::from_bson("garbage")); }));

  // SetModel: a new network under the same ground truth -- step counter restarts, training works
  vnrNeuralVolumeSetModel(neural, vnr::jx::from_text(kModelSmall), 5);
  CHECK(vnrNeuralVolumeGetTrainingStep(neural) == 0);
  vnrNeuralVolumeTrain(neural, 50, true);
  const double small_loss = vnrNeuralVolumeGetTrainingLoss(neural);
  std::cout << "setmodel_train_step " << vnrNeuralVolumeGetTrainingStep(neural) << "\nsetmodel_train_loss " << small_loss << std::endl;
  CHECK(vnrNeuralVolumeGetTrainingStep(neural) == 50 && small_loss > 0 && small_loss < 0.2);
  CHECK(throws([&] { vnrNeuralVolumeSetModel(neural, vnr::jx::from_text("{\"encoding\":{\"otype\":\"Frequency\"}}")); }));
  vnrRendererResetAccumulation(rn);
  vnrRender(rn);                                         // the renderer follows the new network
  CHECK(mean_alpha(vnrRendererMapFrame(rn), fb.x * fb.y) >= 0.0);
  // params.json carries its model: loading it resets the network to the stored architecture (load_params_from_json,
  // network.cu:925-929) and the first frame comes back bit for bit
  vnrNeuralVolumeSetParams(neural, blob);
  vnrRendererResetAccumulation(rn);
  vnrRender(rn);
  const double restored_diff = max_diff(frame_n, vnrRendererMapFrame(rn));
  std::cout << "restored_frame_max_abs " << restored_diff << std::endl;
  CHECK(restored_diff == 0.0);

  // untrained volume of given dims (api.h:123) + memory queries
  vnrVolume blank = vnrCreateNeuralVolume(vnr::jx::from_text(kModelSmall), vnr::vec3i(16, 16, 16));
  CHECK(throws([&] { vnrNeuralVolumeTrain(blank, 1, true); }));     // no ground truth
  size_t by_renderer = 0, by_network = 0;
  vnrMemoryQuery(&by_renderer, &by_network);
  CHECK(by_renderer > 0 && by_network > 0);
  vnrMemoryQueryPrint("vnr_api_check");
  vnrFreeTemporaryGPUMemory();
  vnrRelease(nullptr);
  std::cout << "device_checks_failed " << failures << std::endl;
}

int main(int ac, char** av) {
  if (ac < 3 || (strcmp(av[1], "--host") && strcmp(av[1], "--device"))) { std::cerr << "usage: vnr_api_check --host|--device <tmpdir>" << std::endl; return 2; }
  try {
    if (!strcmp(av[1], "--host")) host_part(av[2]); else device_part(av[2]);
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << std::endl;
    return 1;
  }
  return failures ? 1 : 0;
}

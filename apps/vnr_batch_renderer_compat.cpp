// The call sequence of the reference's headless renderer (apps/batch_renderer.cpp:156-239), statement for statement, against
// include/vnr_api.hpp -- what a maintainer's app does after swapping the include.  Not copied from the reference: its command
// line parser (CmdArgs), logger, vidi timers and JPEG writer belong to the un-vendored parts and are stood in for by the few
// lines below; every vnr* call, its arguments and their order are the reference's.
// Built twice by tests/test_api_app.py: with the reference's own nlohmann header on the include path (vnrJson = nlohmann::json,
// as in api.h) and without it (the value-type fallback).
#include <chrono>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "vnr_api.hpp"

using namespace vnr;

struct CmdArgs {            // stand-in for apps/cmdline.h: the accessors batch_renderer.cpp uses
  std::string volume_, tfn_, expname_ = "compat";
  int mode_ = 5, frames_ = 4; float density_ = 1.f, rate_ = 1.f; bool simple_ = false;
  CmdArgs(const char*, int ac, char** av) {
    for (int i = 1; i < ac; ++i) {
      const std::string a = av[i];
      auto next = [&]() { return std::string(i + 1 < ac ? av[++i] : ""); };
      if (a == "--volume") volume_ = next();
      else if (a == "--simple-volume") { volume_ = next(); simple_ = true; }
      else if (a == "--tfn") tfn_ = next();
      else if (a == "--rendering-mode") mode_ = std::stoi(next());
      else if (a == "--num-frames") frames_ = std::stoi(next());
    }
  }
  bool has_simple_volume() const { return simple_; }
  std::string volume() const { return volume_; }
  std::string tfn() const { return tfn_; }
  std::string expname() const { return expname_; }
  int rendering_mode() const { return mode_; }
  int num_frames() const { return frames_; }
  float density_scale() const { return density_; }
  float sampling_rate() const { return rate_; }
};
struct Timer {
  std::chrono::steady_clock::time_point t0; double ms = 0;
  void start() { t0 = std::chrono::steady_clock::now(); } void reset() { ms = 0; }
  void stop() { ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
  double milliseconds() const { return ms; }
};

extern "C" int main(int ac, char** av) {
  CmdArgs args("Commandline Volume Renderer", ac, av);
  if (args.volume().empty()) { std::printf("usage: %s (--volume params.json | --simple-volume scene.json) --tfn scene.json\n", av[0]); return 0; }

  vnrVolume volume;

  if (args.has_simple_volume()) {
    volume = vnrCreateSimpleVolume(args.volume(), "GPU", false);
  }
  else {
    vnrJson params;
    vnrLoadJsonBinary(params, args.volume());
    volume = vnrCreateNeuralVolume(params);
  }

  auto camera = vnrCreateCamera();
  vnrCameraSet(camera, args.tfn());
  auto from = vnrCameraGetPosition(camera);
  auto at = vnrCameraGetFocus(camera);
  auto up = vnrCameraGetUpVec(camera);

  auto tfn = vnrCreateTransferFunction(args.tfn());
  vnrTransferFunctionSetValueRange(tfn, range1f(0, 1));

  auto ren = vnrCreateRenderer(volume);
  vnrRendererSetTransferFunction(ren, tfn);
  vnrRendererSetCamera(ren, camera);
  vnrRendererSetFramebufferSize(ren, vec2i(768, 768));
  vnrRendererSetMode(ren, args.rendering_mode());
  vnrRendererSetDenoiser(ren, false);
  vnrRendererSetVolumeDensityScale(ren, args.density_scale());
  vnrRendererSetVolumeSamplingRate(ren, args.sampling_rate());

  for (int i = 0; i < 5; ++i) vnrRender(ren); // warm up

  std::vector<double> timings(args.num_frames());

  Timer timer1, timer2;

  timer1.start();
  for (int i = 0; i < args.num_frames(); ++i) {
    timer2.reset();
    timer2.start();
    vnrRender(ren);
    timer2.stop();
    timings[i] = timer2.milliseconds();
  }
  timer1.stop();
  const auto totaltime = timer1.milliseconds() / 1000.0;

  const vec4f* pixels = vnrRendererMapFrame(ren);

  std::cout << "Summary: " << args.expname() << std::endl;
  std::cout << "\tvolume: " << args.volume() << std::endl;
  std::cout << "\t   tfn: " << args.tfn()    << std::endl;
  std::cout << "\t   fps: " << args.num_frames() / totaltime << std::endl;
  std::cout << "\tdensity scale: " << args.density_scale() << std::endl;
  std::cout << "\tsampling rate: " << args.sampling_rate() << std::endl;
  std::cout << "\tcamera: " << from.x << " " << from.y << " " << from.z << " -> " << at.x << " " << at.y << " " << at.z << " up " << up.x << " " << up.y << " " << up.z << std::endl;
  std::cout << "\tcentre pixel alpha: " << pixels[768 * 384 + 384].w << std::endl;
  return 0;
}

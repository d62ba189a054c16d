// vnr_cmd_render -- the reference's headless renderer (apps/batch_renderer.cpp:156-239) against the
// B200 library: load params.json, set camera / transfer function / framebuffer / mode / sampling
// rate, 5 warm-up frames, N timed vnrRender calls, map the last frame, print the summary.  The
// reference hard-codes a 768x768 framebuffer (:199); --size overrides.  Writes a PPM screenshot.
//   vnr_cmd_render --volume params.json [--num-frames 100] [--size 768] [--rendering-mode 5] [--sampling-rate 1] [--out shot.ppm] [--gpus N]
// --gpus N (N <= 8, one NVSwitch box): the same api.h call sequence on every device of a communicator (vnr_comm_init,
// include/vnr_c.h): the image strips are dealt to the N GPUs, finished pixels land in ONE pinned host frame, rank 0 maps it.
#include <chrono>
#include <cstring>
#include <iostream>
#include <vector>

#include "synthetic.hpp"

int main(int ac, char** av) {
  std::string volume_file = "params.json", out = "screenshot.ppm";
  int frames = 100, size = 768, mode = 5, gpus = 1; float sampling_rate = 1.f, density_scale = 1.f;
  for (int i = 1; i < ac; ++i) {
    auto next = [&]() -> const char* { if (i + 1 >= ac) { std::cerr << "missing value for " << av[i] << std::endl; exit(2); } return av[++i]; };
    if (!strcmp(av[i], "--volume")) volume_file = next();
    else if (!strcmp(av[i], "--num-frames")) frames = atoi(next());
    else if (!strcmp(av[i], "--size")) size = atoi(next());
    else if (!strcmp(av[i], "--rendering-mode")) mode = atoi(next());
    else if (!strcmp(av[i], "--sampling-rate")) sampling_rate = (float)atof(next());
    else if (!strcmp(av[i], "--density-scale")) density_scale = (float)atof(next());
    else if (!strcmp(av[i], "--out")) out = next();
    else if (!strcmp(av[i], "--gpus")) gpus = atoi(next());
    else { std::cerr << "unknown argument " << av[i] << std::endl; return 2; }
  }
  try {
    if (gpus < 1 || gpus > 8) throw std::runtime_error("--gpus must be 1..8");
    std::vector<vnr_comm_t*> comms((size_t)gpus, nullptr);
    if (gpus > 1) vnr::check(vnr_comm_init(gpus, comms.data()));
    vnrJson params;
    vnrLoadJsonBinary(params, volume_file);
    std::vector<vnrVolume> volumes; std::vector<vnrRenderer> rens;
    for (int r = 0; r < gpus; ++r) {
      if (gpus > 1) vnr::check(vnr_comm_set_device(comms[r]));             // the objects of rank r live on its device
      vnrVolume volume = vnrCreateNeuralVolume(params);

      auto camera = synthetic::orbit_camera(volume->dims, 1);
      auto tfn = synthetic::make_tfn();
      vnrTransferFunctionSetValueRange(tfn, vnr::range1f(0, 1));

      auto ren = vnrCreateRenderer(volume);
      vnrRendererSetTransferFunction(ren, tfn);
      vnrRendererSetCamera(ren, camera);
      vnrRendererSetFramebufferSize(ren, vnr::vec2i(size, size));
      vnrRendererSetMode(ren, mode);
      vnrRendererSetDenoiser(ren, false);
      vnrRendererSetVolumeDensityScale(ren, density_scale);
      vnrRendererSetVolumeSamplingRate(ren, sampling_rate);
      volumes.push_back(volume); rens.push_back(ren);
    }
    if (gpus > 1) {                                                        // collective: completes when the last rank attaches
      for (int r = 0; r < gpus; ++r) vnr::check(vnr_volume_attach_comm(castNeuralVolume(volumes[r])->h, comms[r]));
      for (int r = 0; r < gpus; ++r) vnr::check(vnr_renderer_attach_comm(rens[r]->h, comms[r]));
    }
    vnrVolume volume = volumes[0];
    vnrRenderer ren = rens[0];
    // a frame = vnrRender on every rank (asynchronous: the ranks' streams run concurrently); rank 0 maps
    auto render_all = [&]() { for (auto& r : rens) vnrRender(r); };

    for (int i = 0; i < 5; ++i) render_all();      // warm up
    vnrRendererMapFrame(ren);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < frames; ++i) render_all();
    const vnr::vec4f* pixels = vnrRendererMapFrame(ren);   // syncs with the last frame
    const double totaltime = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    FILE* f = fopen(out.c_str(), "wb");
    if (f) {
      fprintf(f, "P6\n%d %d\n255\n", size, size);
      for (int y = size - 1; y >= 0; --y)
        for (int x = 0; x < size; ++x) {
          const vnr::vec4f p = pixels[(size_t)y * size + x];
          const float c[3] = {p.x, p.y, p.z};      // premultiplied colour over black
          for (float v : c) fputc((int)(255.f * std::min(1.f, std::max(0.f, v)) + 0.5f), f);
        }
      fclose(f);
    }
    std::cout << "Summary" << std::endl;
    std::cout << "\tvolume: " << volume_file << " (" << volume->dims.x << "x" << volume->dims.y << "x" << volume->dims.z << ")" << std::endl;
    std::cout << "\t   fps: " << frames / totaltime << std::endl;
    std::cout << "\tsampling rate: " << sampling_rate << std::endl;
    std::cout << "\tscreenshot: " << out << std::endl;
    std::cout << "\tgpus: " << gpus << std::endl;
    if (gpus > 1) {
      for (int r = 0; r < gpus; ++r) { vnr_renderer_detach_comm(rens[r]->h); vnr_volume_detach_comm(castNeuralVolume(volumes[r])->h); }
      rens.clear(); volumes.clear(); ren.reset(); volume.reset();
      for (auto c : comms) vnr_comm_release(c);
    }
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}

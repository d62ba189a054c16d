// vnr_cmd_train -- the reference's headless trainer (apps/batch_trainer.cpp:72-141) against the
// B200 library: create the ground-truth volume, create the neural volume from the model config,
// train in bursts of 10 steps in fast mode, report STEP / LOSS / TIME / PSNR, write params.json.
// The scene-file ingest of the reference is replaced by a seeded synthetic volume (--dims N).
//   vnr_cmd_train [--config example-model.json] [--dims 256] [--max-num-steps 2000] [--train-macrocell] [--out params.json] [--seed S]
#include <chrono>
#include <cstring>
#include <iostream>

#include "synthetic.hpp"

int main(int ac, char** av) {
  std::string config = "instantvnr_b200/configs/example-model.json", out = "params.json";
  int dim = 256, steps = 2000; bool train_macrocell = false; unsigned seed = 1337;
  for (int i = 1; i < ac; ++i) {
    auto next = [&]() -> const char* { if (i + 1 >= ac) { std::cerr << "missing value for " << av[i] << std::endl; exit(2); } return av[++i]; };
    if (!strcmp(av[i], "--config")) config = next();
    else if (!strcmp(av[i], "--dims")) dim = atoi(next());
    else if (!strcmp(av[i], "--max-num-steps")) steps = atoi(next());
    else if (!strcmp(av[i], "--train-macrocell")) train_macrocell = true;
    else if (!strcmp(av[i], "--out")) out = next();
    else if (!strcmp(av[i], "--seed")) seed = (unsigned)atoi(next());
    else { std::cerr << "unknown argument " << av[i] << std::endl; return 2; }
  }
  try {
    vnrJson model = vnrCreateJsonText(config);
    const vnr::vec3i dims(dim, dim, dim);
    const std::vector<float> voxels = synthetic::make_volume(dims);
    vnrVolume simple_volume = vnrCreateSimpleVolume(voxels.data(), dims);
    vnrVolume neural_volume;
  restart:
    neural_volume = vnrCreateNeuralVolume(model, simple_volume, train_macrocell, seed);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; i += 10) {
      vnrNeuralVolumeTrain(neural_volume, 10, true);
      if (i >= 5000 && vnrNeuralVolumeGetTrainingLoss(neural_volume) > /*bad loss = */ 0.9) {       // batch_trainer.cpp:114-118
        std::cout << "bad setup, ... restart" << std::endl;
        ++seed;
        goto restart;
      }
    }
    const double loss = vnrNeuralVolumeGetTrainingLoss(neural_volume);      // syncs
    const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const double psnr = vnrNeuralVolumeGetPSNR(neural_volume, false);
    std::cout << "Summary" << std::endl;
    std::cout << "  STEP=" << vnrNeuralVolumeGetTrainingStep(neural_volume) << std::endl;
    std::cout << "  LOSS=" << loss << std::endl;
    std::cout << "  TIME=" << seconds << "s" << std::endl;
    std::cout << "  STEPS_PER_SEC=" << steps / seconds << std::endl;
    std::cout << "  PSNR=" << psnr << std::endl;
    vnrJson output;
    vnrNeuralVolumeSerializeParams(neural_volume, output);
    vnrSaveJsonBinary(output, out);
    std::cout << "  wrote " << out << " (" << output.data.size() << " bytes)" << std::endl;
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}

// train_loop.cpp — the reference's Train() / TrainStep() (ScratchNerf/ScratchNerf/Program.cs:21-62) as a native C++ host
// over include/nerfb200.hpp, using the SAME call sequence: GetGradient with a host callback that calls
// AcceleratedGradientCalculator.get_output_gradient, optimizer.step(model.mlp.allParams, grad, lr),
// OutputRetriever.RetrieveOutput every PrintEvery steps.  Rays come from a 64-byte-record train_data.bin
// (SN/BinDataset.cs:35-49) when a path is given, else from a tiny procedural batch.
//
//   g++ -std=c++17 -Iinclude examples/train_loop.cpp -Lnerf_or_nothing_b200 -lnerfb200 -Wl,-rpath,$PWD/nerf_or_nothing_b200 -o train_loop
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "nerfb200.hpp"

using namespace AcceleratedNeRFUtils;

int main(int argc, char** argv) {
  const int batch = 1024, steps = argc > 2 ? std::atoi(argv[2]) : 3;  // SN/BinDataset.cs:12
  std::vector<Vector3> origins(batch), dirs(batch), pixels(batch);
  std::vector<float> radii(batch), nears(batch, 2.f), fars(batch, 6.f), lossMults(batch, 1.f);
  std::FILE* f = argc > 1 ? std::fopen(argv[1], "rb") : nullptr;
  std::mt19937 rng(2024);
  std::uniform_real_distribution<float> u(-1.f, 1.f);
  try {
    AcceleratedMipNeRF model;                                       // Program.cs:24
    AcceleratedAdamOptimizer optimizer(model.GetLayerSizes());      // Program.cs:25
    AcceleratedGradientCalculator gradientCalculator(batch);        // Program.cs:26
    for (int step = 1; step <= steps; ++step) {
      for (int i = 0; i < batch; ++i) {
        float rec[16];
        if (f && std::fread(rec, sizeof(float), 16, f) == 16) {     // o(3) d(3) viewdir(3) radius near far lossmult rgb(3)
          origins[i] = {rec[0], rec[1], rec[2]}; dirs[i] = {rec[3], rec[4], rec[5]};
          radii[i] = rec[9]; nears[i] = rec[10]; fars[i] = rec[11]; lossMults[i] = rec[12]; pixels[i] = {rec[13], rec[14], rec[15]};
        } else {
          origins[i] = {0.f, 0.f, 4.f}; dirs[i] = {0.3f * u(rng), 0.3f * u(rng), -1.f};
          radii[i] = 1.04e-3f; pixels[i] = {0.5f + 0.5f * u(rng), 0.5f, 0.5f};
        }
      }
      uint64_t output = 0;
      const float lr = LearningRateDecay(step);                                       // MipHelpers.cs:758-773 via the library
      auto grad = model.GetGradient(origins, dirs, radii, nears, fars, lossMults,     // Program.cs:51-58
                                    [&](uint64_t inputptr, int level, float lossMultSum, uint64_t lm) {
                                      output = inputptr;
                                      return gradientCalculator.get_output_gradient(inputptr, pixels, lm, lossMultSum, level);
                                    });
      optimizer.step(model.mlp.allParams(), grad, lr);                                // Program.cs:59-60
      auto returned = OutputRetriever::RetrieveOutput(output, batch);                 // Program.cs:42
      double loss = 0, lm = 0;
      for (int i = 0; i < batch; ++i) {                                               // Program.cs:64
        const float dx = returned[i].X - pixels[i].X, dy = returned[i].Y - pixels[i].Y, dz = returned[i].Z - pixels[i].Z;
        loss += lossMults[i] * (dx * dx + dy * dy + dz * dz); lm += lossMults[i];
      }
      std::printf("Step %d/%d, Loss: %g\n", step, steps, loss / lm);
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 2;
  }
  if (f) std::fclose(f);
  return 0;
}

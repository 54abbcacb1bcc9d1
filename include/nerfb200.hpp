// nerfb200.hpp — header-only C++ mirror of the reference's five wrapper classes over the C ABI of nerfb200.h.
//
// The reference's wrappers are C++/CLI `ref class`es (ScratchNerf/AcceleratedNeRFUtils/*.h) and cannot be compiled
// outside MSVC /clr; these plain C++ classes keep their names, method names and argument order so that a native host
// reads like the reference's Train() / TrainStep() (ScratchNerf/ScratchNerf/Program.cs:21-62).  Errors, which the
// reference prints and ignores (ANU/AcceleratedMLP.cpp:265-268), become std::runtime_error carrying nerf_last_error().
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "nerfb200.h"

namespace AcceleratedNeRFUtils {

inline void check(int status) {
  if (status != 0) throw std::runtime_error("libnerfb200 error " + std::to_string(status) + ": " + nerf_last_error());
}

struct Vector3 { float X, Y, Z; };  // System.Numerics.Vector3 == packed float3

class AcceleratedMipNeRF;

// ANU/AcceleratedMLP.h:7-45 (reached as model.mlp, like the reference's field)
class AcceleratedMLP {
 public:
  explicit AcceleratedMLP(nerf_mipnerf* h) : h_(h) {}
  std::vector<float*> allParams() const {  // ANU/AcceleratedMLP.h:25
    int n = 0;
    check(nerf_mipnerf_num_tensors(h_, &n));
    std::vector<float*> p(n);
    check(nerf_mipnerf_all_params(h_, p.data()));
    return p;
  }
  std::vector<float*> allGradients() const {  // ANU/AcceleratedMLP.h:24
    int n = 0;
    check(nerf_mipnerf_num_tensors(h_, &n));
    std::vector<float*> p(n);
    check(nerf_mipnerf_all_gradients(h_, p.data()));
    return p;
  }
  // ANU/AcceleratedMLP.cpp:214-255 -> (density, rgb) device pointers (named, SURVEY A-D2)
  std::pair<uint64_t, uint64_t> get_output(const float* dev_encoded_position, const float* dev_encoded_direction, int level, int n_rays) {
    uint64_t d = 0, r = 0;
    check(nerf_mlp_get_output(h_, dev_encoded_position, dev_encoded_direction, level, n_rays, &d, &r));
    return {d, r};
  }
  std::vector<float*> get_gradient(const float* color_gradient, const float* density_gradient, int level) {  // .cpp:256-321
    check(nerf_mlp_get_gradient(h_, color_gradient, density_gradient, level, nullptr));
    return allGradients();
  }
  void reset_gradients(int level) { check(nerf_mlp_reset_gradients(h_, level)); }  // .cpp:113-129

 private:
  nerf_mipnerf* h_;
};

// ANU/AcceleratedMipNeRF.h:10-41
class AcceleratedMipNeRF {
 public:
  using OutputGradientFn = std::function<uint64_t(uint64_t comp_rgb_dev, int level, float loss_mult_sum, uint64_t loss_mults_dev)>;

  AcceleratedMipNeRF() : AcceleratedMipNeRF(default_config()) {}  // the reference's compile-time configuration
  explicit AcceleratedMipNeRF(const nerf_config& cfg) : cfg_(cfg), mlp(nullptr) {
    check(nerf_mipnerf_create(&cfg_, &h_));
    mlp = AcceleratedMLP(h_);
  }
  ~AcceleratedMipNeRF() { nerf_mipnerf_destroy(h_); }
  AcceleratedMipNeRF(const AcceleratedMipNeRF&) = delete;
  AcceleratedMipNeRF& operator=(const AcceleratedMipNeRF&) = delete;

  static nerf_config default_config() {
    nerf_config c;
    nerf_default_config(&c);
    return c;
  }
  std::vector<int> GetLayerSizes() const {  // ANU/AcceleratedMipNeRF.cpp:146-149
    int n = 0;
    check(nerf_mipnerf_num_tensors(h_, &n));
    std::vector<int> s(n);
    check(nerf_mipnerf_get_layer_sizes(h_, s.data(), &n));
    return s;
  }
  // ANU/AcceleratedMipNeRF.cpp:52-144; getOutputGradient == nullptr uses the built-in MSE against SetPixels()
  std::vector<float*> GetGradient(const std::vector<Vector3>& origins, const std::vector<Vector3>& directions,
                                  const std::vector<float>& radii, const std::vector<float>& nears, const std::vector<float>& fars,
                                  const std::vector<float>& lossMultipliers, const OutputGradientFn& getOutputGradient = nullptr) {
    struct Tramp {
      static uint64_t call(uint64_t c, int l, float s, uint64_t m, void* user) { return (*static_cast<const OutputGradientFn*>(user))(c, l, s, m); }
    };
    check(nerf_mipnerf_get_gradient(h_, &origins[0].X, &directions[0].X, radii.data(), nears.data(), fars.data(), lossMultipliers.data(),
                                    (int)origins.size(), getOutputGradient ? &Tramp::call : nullptr,
                                    getOutputGradient ? const_cast<OutputGradientFn*>(&getOutputGradient) : nullptr, nullptr));
    return mlp.allGradients();
  }
  void SetPixels(const std::vector<Vector3>& pixels) { check(nerf_mipnerf_set_pixels(h_, &pixels[0].X, (int)pixels.size())); }
  nerf_mipnerf* handle() const { return h_; }

 private:
  nerf_config cfg_;
  nerf_mipnerf* h_ = nullptr;

 public:
  AcceleratedMLP mlp;  // ANU/AcceleratedMipNeRF.h:18
};

// ANU/AcceleratedAdamOptimizer.h:5-20
class AcceleratedAdamOptimizer {
 public:
  explicit AcceleratedAdamOptimizer(const std::vector<int>& layer_sizes, int eps_mode = 0, int device = 0) {
    check(nerf_adam_create(layer_sizes.data(), (int)layer_sizes.size(), eps_mode, device, &a_));
  }
  ~AcceleratedAdamOptimizer() { nerf_adam_destroy(a_); }
  AcceleratedAdamOptimizer(const AcceleratedAdamOptimizer&) = delete;
  void step(std::vector<float*> params, std::vector<float*> grads, float learning_rate) {  // .cpp:23-41
    check(nerf_adam_step(a_, params.data(), grads.data(), learning_rate));
  }
  nerf_adam* handle() const { return a_; }

 private:
  nerf_adam* a_ = nullptr;
};

// ANU/AcceleratedGradientCalculator.h:8-17
class AcceleratedGradientCalculator {
 public:
  explicit AcceleratedGradientCalculator(int batch_size, int n_levels = 2, float coarse_loss_mult = 0.1f, int device = 0) {
    check(nerf_gradcalc_create(batch_size, n_levels, coarse_loss_mult, device, &g_));
  }
  ~AcceleratedGradientCalculator() { nerf_gradcalc_destroy(g_); }
  AcceleratedGradientCalculator(const AcceleratedGradientCalculator&) = delete;
  uint64_t get_output_gradient(uint64_t input, const std::vector<Vector3>& pixels, uint64_t loss_mults, float loss_mult_sum, int level) {
    uint64_t out = 0;  // .cpp:18-30
    check(nerf_gradcalc_get_output_gradient(g_, input, &pixels[0].X, (int)pixels.size(), loss_mults, loss_mult_sum, level, &out));
    return out;
  }

 private:
  nerf_gradcalc* g_ = nullptr;
};

// ScratchNerf/ScratchNerf/BinDataset.cs:10-52 with the 64-byte records resident in device memory: Next() + the upload
// in AcceleratedMipNeRF::GetGradient become one on-device draw-and-gather inside TrainStep.
class BinDataset {
 public:
  static constexpr int BatchSize = 1024;  // BinDataset.cs:12
  explicit BinDataset(const std::string& file, int device = 0) { check(nerf_dataset_load(file.c_str(), device, &d_)); }
  BinDataset(const void* records, long n_records, int device = 0) { check(nerf_dataset_create(records, n_records, device, &d_)); }
  ~BinDataset() { nerf_dataset_destroy(d_); }
  BinDataset(const BinDataset&) = delete;
  long NumSamples() const {  // BinDataset.cs:15
    long n = 0;
    check(nerf_dataset_size(d_, &n));
    return n;
  }
  // one iteration of Train() (Program.cs:28-45): draw the batch, forward, backward, Adam; returns the total loss
  float TrainStep(AcceleratedMipNeRF& model, AcceleratedAdamOptimizer& optimizer, int batch_size, uint64_t sampler_seed, float lr) {
    float loss = 0.f;
    check(nerf_mipnerf_train_step_dataset(model.handle(), optimizer.handle(), d_, batch_size, sampler_seed, lr, &loss));
    return loss;
  }
  nerf_dataset* handle() const { return d_; }

 private:
  nerf_dataset* d_ = nullptr;
};

// MipHelpers.LearningRateDecay (ScratchNerf/ScratchNerf/MipHelpers.cs:758-773) with Config's defaults (TrainState.cs:54-60)
inline float LearningRateDecay(int step, float lr_init = 5e-4f, float lr_final = 5e-6f, int max_steps = 1000000, int lr_delay_steps = 2500,
                               float lr_delay_mult = 0.01f) {
  return nerf_learning_rate_decay(step, lr_init, lr_final, max_steps, lr_delay_steps, lr_delay_mult);
}
inline void SaveCheckpoint(AcceleratedMipNeRF& model, AcceleratedAdamOptimizer* optimizer, const std::string& path) {
  check(nerf_checkpoint_save(model.handle(), optimizer ? optimizer->handle() : nullptr, path.c_str()));
}
inline void LoadCheckpoint(AcceleratedMipNeRF& model, AcceleratedAdamOptimizer* optimizer, const std::string& path) {
  check(nerf_checkpoint_load(model.handle(), optimizer ? optimizer->handle() : nullptr, path.c_str()));
}
// MathHelpers.MseToPsnr / ComputeSsimAverage (ScratchNerf/ScratchNerf/MipHelpers.cs:672, 688-737) of two host images [height, width, 3]
struct ImageMetrics { double mse, psnr, ssim; };
inline ImageMetrics CompareImages(const float* a, const float* b, int width, int height, float max_val = 1.0f) {
  ImageMetrics m{};
  check(nerf_image_error(a, b, (long)width * height * 3, 0, &m.mse, &m.psnr));
  check(nerf_image_ssim(a, b, width, height, max_val, 11, 1.5f, 0.01f, 0.03f, 0, &m.ssim, nullptr));
  return m;
}

// ANU/OutputRetriever.h:7-11
struct OutputRetriever {
  static std::vector<Vector3> RetrieveOutput(uint64_t dev_output, int size) {  // .cpp:6-14
    std::vector<Vector3> out(size);
    check(nerf_retrieve_output(dev_output, size, &out[0].X));
    return out;
  }
};

}  // namespace AcceleratedNeRFUtils

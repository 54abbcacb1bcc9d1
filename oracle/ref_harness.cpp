/*
 * ref_harness.cpp — C-ABI launcher for the REFERENCE's own CUDA kernels (pin P1).
 *
 * TEST INFRASTRUCTURE ONLY.  oracle/Makefile compiles
 *   /root/reference/ScratchNerf/AcceleratedNeRFUtils/accelerated_functions.cu
 * UNMODIFIED, from where it lies, into oracle/_ref/accelerated_functions.o and links it with this
 * file into oracle/_ref/libref_kernels.so (git-ignored; travels to the GPU box).  No reference
 * source is copied into the repository.
 *
 * The kernels are reached exactly the way the reference's C++/CLI wrappers reach them
 * (ANU/AcceleratedMLP.cpp:22-25, ANU/AcceleratedMipNeRF.cpp:55-60): an `extern void kernel(...)`
 * host-stub declaration + cudaLaunchKernel((void*)kernel, grid, block, args) with the reference's
 * launch shapes block_1d=(1024), block_2d=(32,32), block_3d=(16,8,8) (ANU/helpers.cpp:2-4) and its
 * ceil-div grid rule (ANU/helpers.h:12-15).  Arguments are the CORRECTED ones where the reference
 * call sites are defective (D1 density-head N/K swap, D8 2-D launch).
 * Problem size is the reference's compile-time constant: num_rays=1024, num_samples=128
 * (.cu:15-16).  All pointers are device pointers; every call synchronises and returns the
 * cudaError_t as int.
 */
#include <cuda_runtime.h>

#include <cstdint>

extern void get_neuron_output(const float*, const float*, const float*, float*, float*, const int, const int);
extern void get_neuron_output_sigmoid(const float*, const float*, const float*, float*, float*, const int, const int);
extern void get_neuron_output_soft_plus(const float*, const float*, const float*, float*, float*, const int, const int);
extern void get_neuron_output_conjoined_inputs(const float*, const float*, const float*, const float*, float*, float*, const int, const int, const int);
extern void backpropagate_neuron(const float*, const float*, const float*, const float*, float*, float*, float*, const int, const int);
extern void backpropagate_neuron_sigmoid(const float*, const float*, const float*, const float*, float*, float*, float*, const int, const int);
extern void backpropagate_neuron_soft_plus(const float*, const float*, const float*, const float*, float*, float*, float*, const int, const int);
extern void backpropagate_neuron_partial_conjoined(const float*, const float*, const float*, const float*, const float*, float*, float*, float*, const int, const int, const int);
extern void encode_input_data(const float3*, const float3*, const float3*, float*, float*);
extern void cast_rays(const float*, const float3*, const float3*, float3*, float3*, const float*);
extern void volumetric_rendering(const float3*, const float*, const float*, const float3*, float3*, float*, float*, float*);
extern void get_output_gradient(const float3*, const float3*, const float*, float3*, const float, const int);
extern void volumetric_rendering_gradient(const float3*, const float*, const float*, const float*, const float3*, const float*, const float3*, float3*, float*);
extern void adam_optimizer_step(float*, const float*, float*, float*, const float, const float, const float, const float, const float, const int);

namespace {
constexpr int kRays = 1024, kSamples = 128;  // .cu:15-16
const dim3 block_1d(1024), block_2d(32, 32), block_3d(16, 8, 8);
dim3 cdiv(dim3 a, dim3 b) { return dim3((a.x + b.x - 1) / b.x, (a.y + b.y - 1) / b.y, (a.z + b.z - 1) / b.z); }
int launch(const void* fn, dim3 grid, dim3 block, void** args) {
  cudaError_t e = cudaLaunchKernel(fn, grid, block, args, 0, nullptr);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaDeviceSynchronize();
}
}  // namespace

extern "C" {
int ref_num_rays() { return kRays; }
int ref_num_samples() { return kSamples; }

// kind: 0 relu (.cu:36), 1 sigmoid (.cu:49), 2 softplus (.cu:62)
int ref_apply_layer(int kind, const float* in, const float* w, const float* b, float* out, float* z, int n, int k) {
  void* a[7] = {&in, &w, &b, &out, &z, &n, &k};
  const void* fn = kind == 0 ? (const void*)get_neuron_output
                   : kind == 1 ? (const void*)get_neuron_output_sigmoid
                               : (const void*)get_neuron_output_soft_plus;
  return launch(fn, cdiv(dim3(n, kRays, kSamples), block_3d), block_3d, a);
}
int ref_apply_layer_conjoined(const float* ia, const float* ib, const float* w, const float* b, float* out, float* z, int n, int ka, int kb) {
  void* a[9] = {&ia, &ib, &w, &b, &out, &z, &n, &ka, &kb};
  return launch((const void*)get_neuron_output_conjoined_inputs, cdiv(dim3(n, kRays, kSamples), block_3d), block_3d, a);
}
int ref_backpropagate_layer(int kind, const float* in, const float* w, const float* z, const float* dout, float* din, float* dw, float* db, int n, int k) {
  void* a[9] = {&in, &w, &z, &dout, &din, &dw, &db, &n, &k};
  const void* fn = kind == 0 ? (const void*)backpropagate_neuron
                   : kind == 1 ? (const void*)backpropagate_neuron_sigmoid
                               : (const void*)backpropagate_neuron_soft_plus;
  return launch(fn, cdiv(dim3(n, kRays, kSamples), block_3d), block_3d, a);
}
int ref_backpropagate_layer_partial_conjoined(const float* ia, const float* ib, const float* w, const float* z, const float* dout, float* dia, float* dw, float* db, int n, int ka, int kb) {
  void* a[11] = {&ia, &ib, &w, &z, &dout, &dia, &dw, &db, &n, &ka, &kb};
  return launch((const void*)backpropagate_neuron_partial_conjoined, cdiv(dim3(n, kRays, kSamples), block_3d), block_3d, a);
}
int ref_cast_rays(const float* t, const float* o, const float* d, float* mean, float* cov, const float* radii) {
  void* a[6] = {&t, &o, &d, &mean, &cov, &radii};
  return launch((const void*)cast_rays, cdiv(dim3(kRays, kSamples), block_2d), block_2d, a);
}
// NOTE the reference indexes direction_data per SAMPLE (.cu:208, A-D10): dirs must be [R*S] float3.
int ref_encode_input_data(const float* mean, const float* cov, const float* dirs_per_sample, float* enc_pos, float* enc_dir) {
  void* a[5] = {&mean, &cov, &dirs_per_sample, &enc_pos, &enc_dir};
  return launch((const void*)encode_input_data, cdiv(dim3(kRays, kSamples, 16), block_3d), block_3d, a);
}
int ref_volumetric_rendering(const float* rgb, const float* density, const float* t, const float* d, float* comp, float* alpha, float* trans, float* w) {
  void* a[8] = {&rgb, &density, &t, &d, &comp, &alpha, &trans, &w};
  return launch((const void*)volumetric_rendering, cdiv(dim3(kRays), block_1d), block_1d, a);
}
int ref_get_output_gradient(const float* comp, const float* pix, const float* lm, float* g, float lm_sum, int level) {
  void* a[6] = {&comp, &pix, &lm, &g, &lm_sum, &level};
  return launch((const void*)get_output_gradient, cdiv(dim3(kRays), block_1d), block_1d, a);
}
int ref_volumetric_rendering_gradient(const float* g, const float* alpha, const float* trans, const float* w, const float* rgb, const float* t, const float* d, float* drgb, float* dden) {
  void* a[9] = {&g, &alpha, &trans, &w, &rgb, &t, &d, &drgb, &dden};
  return launch((const void*)volumetric_rendering_gradient, cdiv(dim3(kRays), block_1d), block_1d, a);
}
int ref_adam_optimizer_step(float* p, const float* g, float* m, float* v, float lr, float b1, float b2, float inv1, float inv2, int n) {
  void* a[10] = {&p, &g, &m, &v, &lr, &b1, &b2, &inv1, &inv2, &n};
  return launch((const void*)adam_optimizer_step, cdiv(dim3(n), block_1d), block_1d, a);
}
}

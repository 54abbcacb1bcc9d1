// kernels.cuh — host-side launchers of every hot-path kernel (definitions in the .cu files beside this).
// All pointers are device pointers; every launcher enqueues on `st` and returns 0 or a cudaError_t.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace nerf {

// RNG source for the sampling kernels: explicit uniforms (parity runs) or Philox(seed, step, level).
struct SampleRng {
  const float* u = nullptr;  // [R, S+1] or null
  uint64_t seed = 0;
  uint32_t step = 0, level = 0, ray0 = 0;
};

// ---- sampling.cu -------------------------------------------------------------------
int launch_sample_t_vals(const float* nears, const float* fars, SampleRng rng, int R, int S, int randomized,
                         float* t, cudaStream_t st);
int launch_resample_t_vals(const float* t, const float* w, SampleRng rng, int R, int S, float padding,
                           int randomized, float* t_new, cudaStream_t st);

// ---- encode.cu (B.1, B.2) ------------------------------------------------------------------------
int launch_cast_rays(const float* t, const float* o, const float* d, const float* radii, int R, int S,
                     float* means, float* covs, cudaStream_t st);
int launch_encode_input_data(const float* means, const float* covs, const float* dirs, float* enc_pos,
                             float* enc_dir, int R, int S, int deg_point, int deg_view, cudaStream_t st);
// cast_rays + IPE (+ direction PE) in one pass, never materialising mean/cov.  Outputs (each optional):
//   enc_pos_f32 [M, 6*deg_point]; enc_dir_f32 [M, dir_pitch_f32];
//   bf16 split planes for the tensor-core path: pos_hi/pos_lo [M, pos_pitch_h], dir_hi/dir_lo [M, dir_pitch_h]
struct EncodeOut {
  float* enc_pos_f32 = nullptr;
  float* enc_dir_f32 = nullptr;
  int dir_pitch_f32 = 0;
  __nv_bfloat16 *pos_hi = nullptr, *pos_lo = nullptr, *dir_hi = nullptr, *dir_lo = nullptr;
  int pos_pitch_h = 0, dir_pitch_h = 0;
  // optional third planes, fp16, same pitches: the X operands of the fp32-accurate mode's fp16 wgrad GEMMs (mlp_tc.cu)
  void *pos_f16 = nullptr, *dir_f16 = nullptr;
};
int launch_cast_encode_fused(const float* t, const float* o, const float* d, const float* radii, int R, int S,
                             int deg_point, int deg_view, EncodeOut out, cudaStream_t st);

// ---- compositing.cu (B.4, B.5) --------------------------------------------------------------------
struct OutputAct {  // activations feeding compositing; raw=false means inputs are already activated
  bool raw = false;
  float density_bias = 0.f, rgb_padding = 0.f;
};
int launch_composite_fwd(const float* rgb, const float* density, const float* t, const float* dirs, int R,
                         int S, int white_bkgd, OutputAct act, float* comp_rgb, float* depth, float* acc,
                         float* weights, cudaStream_t st);
// g: dL/d comp_rgb [R,3].  With act.raw the outputs are dL/d RAW density / rgb (activation grads fused).
int launch_composite_bwd(const float* g, const float* rgb, const float* density, const float* t,
                         const float* dirs, int R, int S, int white_bkgd, int last_sample_mode, OutputAct act,
                         float* d_rgb, float* d_density, cudaStream_t st);
// g = 2*lm/lm_sum*(rgb-pix)*level_mult (.cu:347-361).  lm_sum_dev (device scalar) overrides lm_sum when set.
// loss_out (optional, device scalar): sum(lm*|rgb-pix|^2)/lm_sum (SN/Program.cs:64).
// scratch (optional): NERF_RED_SCRATCH_FLOATS zero-initialised floats; with it the reductions run on several blocks
// for large batches and are still summed in a fixed order (bitwise reproducible).
#define NERF_RED_SCRATCH_FLOATS 72
int launch_output_gradient(const float* comp_rgb, const float* pixels, const float* loss_mults, int R,
                           float lm_sum, const float* lm_sum_dev, float level_mult, float* g, float* loss_out,
                           float* scratch, cudaStream_t st);
int launch_sum(const float* x, int n, float* out, float* scratch, cudaStream_t st);  // deterministic sum
int launch_scale_by_inv(float* g, long n, const float* lm_sum_dev, cudaStream_t st);  // g *= 1 / *lm_sum_dev

// dataset.cu — resident dataset (64-byte records, SN/BinDataset.cs:40-49), on-device batch draw + gather, image error
int launch_draw_indices(uint64_t seed, uint32_t slot0, uint32_t step, long n, int R, long* idx, cudaStream_t st);
int launch_gather_batch(const float* records, long n, const long* idx, uint64_t seed, uint32_t slot0, uint32_t step, int R, float* o,
                        float* d, float* radii, float* nears, float* fars, float* lm, float* pix, cudaStream_t st);
// Dataset.GenerateRays (SN/Dataset.cs:111-176) for pixels [first, first + n) of one camera; c2w12_host: 3 x 4 row-major [R | t] on the HOST
int launch_generate_rays(const float* c2w12_host, float focal, int W, int H, float near, float far, int edge_mode, long first, long n,
                         float* o, float* d, float* radii, float* nears, float* fars, cudaStream_t st);
int launch_sq_err(const float* a, const float* b, long n, double* out, cudaStream_t st);
// SSIM map + per-block sums (SN/MipHelpers.cs:688-757); images [H, W, 3]; block_sums: ssim_blocks(W, H) doubles
long ssim_blocks(int W, int H);
int launch_ssim(const float* a, const float* b, int W, int H, float max_val, int filter_size, float filter_sigma, float k1, float k2,
                float* map_out, double* block_sums, cudaStream_t st);

// ---- adam.cu (B.6) ------------------------------------------------------------------------------
int launch_adam(float* p, const float* g, float* m, float* v, long n, float lr, float b1, float b2, float inv1,
                float inv2, int eps_mode, float grad_scale, cudaStream_t st);
// data-parallel form: the gradient buffer holds the allreduced UN-normalised sum; the pass multiplies by
// 1 / *lm_sum_dev (the allreduced sum of loss multipliers), writes the normalised gradient back and steps.
int launch_adam_dp(float* p, float* g, float* m, float* v, long n, float lr, float b1, float b2, float inv1,
                   float inv2, int eps_mode, const float* lm_sum_dev, cudaStream_t st);

// ---- gemm_simt.cu: strict-fp32 CUDA-core MLP layers ---------------------------------------------
enum Act { ACT_RELU = 0, ACT_SIGMOID = 1, ACT_SOFTPLUS = 2, ACT_NONE = 3 };
// Y = act([A1|A2] W^T + b).  A1 [M,k1] pitch lda1, A2 [M,k2] pitch lda2 (nullable), W [N, k1+k2] row-major.
// Z (pre-activation, optional) and Y (optional) are [M,N] dense.
int launch_dense_fwd(const float* A1, int lda1, int k1, const float* A2, int lda2, int k2, const float* W,
                     const float* b, float* Y, float* Z, long M, int N, Act act, cudaStream_t st);
// dX[M,k1] (=|+=) dZ[M,N] W[:, :k1]; optional rank-1 term r[m]*v[k]; optional mask by (mask[m,k] > 0).
int launch_dense_dgrad(const float* dZ, const float* W, int ldw, float* dX, long M, int N, int k1,
                       const float* r1, const float* v1, const float* mask, bool accumulate, cudaStream_t st);
// dW[N, k1+k2] += dZ^T [A1|A2],  db[N] += colsum(dZ).  workspace: >= dense_wgrad_workspace() floats.
size_t dense_wgrad_workspace(long M, int N, int K);
int launch_dense_wgrad(const float* dZ, const float* A1, int lda1, int k1, const float* A2, int lda2, int k2,
                       float* dW, float* db, long M, int N, float* workspace, cudaStream_t st);
// dZ = dY * act'(Z) for the per-stage API and the thin heads (elementwise)
int launch_act_grad(const float* dY, const float* Z, float* dZ, long n, Act act, cudaStream_t st);
// dz[m,j] = (y[m,j] > 0) ? dy[m,j] : 0
int launch_relu_mask(const float* dY, const float* Y, float* dZ, long n, cudaStream_t st);
// thin heads (N <= 4): forward y[m,n] = x[m,:].w[n,:] + b[n]
int launch_thin_fwd(const float* X, int ldx, const float* W, const float* b, float* Y, long M, int N, int K,
                    cudaStream_t st);
// thin dgrad: dX[m,k] (=|+=) sum_n dz[m,n] W[n,k], optional relu mask
int launch_thin_dgrad(const float* dZ, const float* W, float* dX, long M, int N, int K, const float* mask,
                      bool accumulate, cudaStream_t st);
// thin wgrad: dW[n,k] += sum_m dz[m,n] x[m,k]; db[n] += sum_m dz[m,n]
int launch_thin_wgrad(const float* dZ, const float* X, int ldx, float* dW, float* db, long M, int N, int K,
                      float* workspace, cudaStream_t st);
size_t thin_wgrad_workspace(long M, int N, int K);

// output activations of the heads for the stand-alone AcceleratedMLP API (.cu:60,73 / SN/MipNerfModel.cs:81-83)
int launch_output_activations(const float* raw_density, const float* raw_rgb, long M, OutputAct act, float* density,
                              float* rgb, cudaStream_t st);
int launch_output_activations_grad(const float* raw_density, const float* raw_rgb, const float* d_density,
                                   const float* d_rgb, long M, OutputAct act, float* d_raw_density, float* d_raw_rgb,
                                   cudaStream_t st);
int launch_apply_act(const float* Z, float* Y, long n, Act act, cudaStream_t st);

// misc
int launch_fill(float* p, float v, long n, cudaStream_t st);
int launch_pad_rows(const float* src, int src_pitch, float* dst, int dst_pitch, long rows, int cols,
                    cudaStream_t st);  // dst[r, :cols] = src[r, :cols], dst[r, cols:pitch] = 0

}  // namespace nerf

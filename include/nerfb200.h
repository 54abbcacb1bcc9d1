/*
 * nerfb200.h — C ABI of libnerfb200.so: a B200-native (sm_100a) drop-in for the MipNeRF hot path of
 * SimonMacLean/NeRF-or-nothing (the five C++/CLI classes of ScratchNerf/AcceleratedNeRFUtils and the
 * kernels of accelerated_functions.cu).
 *
 * Citation shorthands: ANU/ = ScratchNerf/AcceleratedNeRFUtils/, SN/ = ScratchNerf/ScratchNerf/,
 * ".cu" = ANU/accelerated_functions.cu.
 *
 * Conventions
 *   - every function returns 0 on success, else a non-zero status (a cudaError_t / ncclResult_t value or
 *     NERF_ERR_*); nerf_last_error() returns the text of the calling thread's last failure.  (The
 *     reference has no error convention at all: CUDA errors are printed as tags and ignored,
 *     ANU/AcceleratedMLP.cpp:265-268.)
 *   - plain pointers and sizes only.  "_dev" arguments are device pointers, everything else is host
 *     memory.  Device pointers handed out by the library stay library-owned (valid until the next call
 *     that recomputes them or until destroy), exactly like the reference's allParams / allGradients
 *     tables (ANU/AcceleratedMLP.h:24-25).
 *   - float3 data is packed xyz (12 bytes), identical to System.Numerics.Vector3 / CUDA float3
 *     (SN/MipNerfModel.cs:221-230).
 *   - no CPU fallback: every entry point fails with NERF_ERR_NO_DEVICE when no sm_100 device exists.
 */
#ifndef NERFB200_H
#define NERFB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERF_OK 0
#define NERF_ERR_INVALID 100001   /* bad argument / unsupported shape */
#define NERF_ERR_NO_DEVICE 100002 /* no usable sm_100 GPU: there is no CPU path */
#define NERF_ERR_STATE 100003     /* call order violated (e.g. get_gradient before get_output) */
#define NERF_ERR_COMM 100004      /* NCCL unavailable / failed */

/* MLP arithmetic mode.  All three accumulate in fp32 and keep fp32 master weights / gradients. */
#define NERF_PRECISION_FP32 0       /* CUDA-core FFMA, strict fp32 like the reference kernels (.cu:36-182) */
#define NERF_PRECISION_FP32_TC 1    /* tcgen05, bf16x3 split operands: fp32-accurate (<=1e-4 rel) */
#define NERF_PRECISION_BF16_TC 2    /* tcgen05, bf16 operands (<=2e-2 rel) */

/* Runtime replacement of the reference's compile-time constants, duplicated in ANU/helpers.h:16-20,
 * .cu:15-16,183-184,346, ANU/AcceleratedMLP.h:10-19 and SN/BinDataset.cs:12.  Fill with
 * nerf_default_config() first, then override. */
typedef struct nerf_config {
  int n_rays;              /* max rays per batch (R)           ANU/helpers.h:18 (1024)   */
  int n_samples;           /* samples per level (S), %32==0, <=256   ANU/helpers.h:17 (128) */
  int n_levels;            /* coarse + fine = 2                ANU/helpers.h:16          */
  int net_depth;           /* 8                                ANU/AcceleratedMLP.h:11   */
  int net_width;           /* 256                              ANU/AcceleratedMLP.h:12   */
  int net_depth_condition; /* 1                                ANU/AcceleratedMLP.h:13   */
  int net_width_condition; /* 128                              ANU/AcceleratedMLP.h:14   */
  int skip_layer;          /* 4                                ANU/AcceleratedMLP.h:19   */
  int deg_point;           /* 16 -> 96 IPE inputs              ANU/helpers.h:19          */
  int deg_view;            /* 4  -> 27 direction inputs        ANU/helpers.h:20          */
  int white_bkgd;          /* 1                                SN/TrainState.cs:71       */
  int randomized;          /* 1: stratified jitter             SN/TrainState.cs:66       */
  int adam_eps_mode;       /* 0: rsqrt(v+1e-8) (.cu:415); 1: 1/(sqrt(v)+1e-8) (SN/TrainState.cs:34) */
  int last_sample_mode;    /* 0: exact gradient of the S-sample forward; 1: reference kernel (.cu:375-379) */
  int precision;           /* NERF_PRECISION_*                                            */
  int device;              /* CUDA device ordinal (reference: cudaSetDevice(0), ANU/AcceleratedMipNeRF.cpp:10) */
  int chunk_rays;          /* rays processed per internal pass (activation-cache size); 0 = auto (<= n_rays) */
  float density_bias;      /* 0 (.cu:73) | -1 (SN/MipNerfModel.cs:20)                     */
  float rgb_padding;       /* 0 (.cu:60) | 0.001 (SN/MipNerfModel.cs:22)                  */
  float coarse_loss_mult;  /* 0.1 (.cu:345)                                               */
  float resample_padding;  /* 0.01 (.cu:243)                                              */
  uint64_t seed;           /* weights + sampling RNG (reference: time(nullptr), A-D7)     */
  uint32_t engine_flags;   /* NERF_FLAG_*: A/B switches of the tensor-core engine (0 = the shipped schedule) */
} nerf_config;

/* engine_flags: bits 0-4, 6 and 8 select the slower, simpler path the default replaced; bits 5, 7 and 9 are alternatives that are not the default
 * (5 and 9 measured slower, 7 trades wgrad accuracy for bandwidth) — parity tests compare every one of them with the default. */
#define NERF_FLAG_NO_FUSED_FORWARD 1u       /* render / forward-only: one GEMM launch per layer instead of the fused kernel */
#define NERF_FLAG_NO_FUSED_TRAIN_FORWARD 2u /* training forward: per-layer launches */
#define NERF_FLAG_NO_FUSED_DGRAD 4u         /* backward: per-layer dgrad launches instead of the fused chain */
#define NERF_FLAG_NO_DEFERRED_REDUCE 8u     /* wgrad partial tiles reduced by a launch of their own */
#define NERF_FLAG_NO_FUSED_ENCODE 16u       /* rendering: cast_rays + IPE + direction PE as a kernel of their own (planes through HBM)
                                              * instead of the encoder warps inside the fused MLP kernels */
#define NERF_FLAG_NO_WEIGHT_MULTICAST 64u   /* fused kernels: one CTA per launch slot, every CTA streams its own weights
                                              * from L2, instead of 2-CTA clusters sharing each weight stage by TMA multicast */
#define NERF_FLAG_FUSED_ENCODE_TRAIN 32u    /* training forward: encoder warps too.  Off by default: measured on a power-capped B200
                                              * (profiles/README.md, r02b) the fused training forward loses more than the 0.3 ms encode
                                              * kernel costs, and the planes must reach HBM for the wgrad GEMMs either way */

#define NERF_FLAG_WGRAD_FP16 128u           /* fp32-accurate mode, opt-in: the wgrad operands (activations, encodings, dZ times a per-level
                                              * power of two) leave the fused kernels as ONE fp16 plane each instead of hi + lo bf16 planes,
                                              * and the wgrad GEMMs run one fp16 MMA per product instead of three bf16 ones: half the HBM
                                              * traffic of the backward pass.  The dgrad chain and everything per-ray are unchanged; the training
                                              * forward switches to the fp8-correction products (see NERF_FLAG_NO_FP8_CORRECTIONS).
                                              * Gradient accuracy: ~1e-5 of the gradient's scale on a real step (sums over ~5e5 samples),
                                              * up to ~1e-3 on sums that cancel like a random walk — outside the mode's 1e-4, so not the default */

#define NERF_FLAG_NO_FP8_CORRECTIONS 256u   /* fp32-accurate mode: the fused forward kernels multiply with three bf16 MMAs per product (hi*hi +
                                              * lo*hi + hi*lo) instead of one fp16 MMA plus two E4M3 correction MMAs at twice the rate onto the
                                              * same accumulator.  The default form is as accurate at the outputs (measured, profiles/README.md)
                                              * and holds for |activation| < 2047, |weight| < 64 (beyond, values saturate instead of overflowing);
                                              * it runs in rendering, and in training when NERF_FLAG_WGRAD_FP16 is set */

#define NERF_FLAG_PAIR_MMA 512u             /* fp32-accurate fused kernels: the 2-CTA clusters issue tcgen05.mma.cta_group::2 (one MMA for both
                                              * tiles, each SM holding half of every weight tile) instead of sharing whole tiles by multicast.
                                              * Same bits; measured slower (profiles/README.md, r02p) */

typedef struct nerf_mipnerf nerf_mipnerf;   /* AcceleratedMipNeRF + its embedded AcceleratedMLP */
typedef struct nerf_adam nerf_adam;         /* AcceleratedAdamOptimizer */
typedef struct nerf_gradcalc nerf_gradcalc; /* AcceleratedGradientCalculator */

/* Host callback of GetGradient (ANU/AcceleratedMipNeRF.h:14-16, call site .cpp:127): called once per
 * level, in level order, on the calling thread after the stream has been synchronised; must return a
 * device pointer to float3[n_rays] dL/d(comp_rgb) valid until get_gradient returns. */
typedef uint64_t (*nerf_output_gradient_cb)(uint64_t comp_rgb_dev, int level, float loss_mult_sum,
                                            uint64_t loss_mults_dev, void* user);

const char* nerf_last_error(void);
int nerf_version(void);
void nerf_default_config(nerf_config* cfg);
int nerf_device_count(int* n);

/* ---- AcceleratedMipNeRF (ANU/AcceleratedMipNeRF.h:10-41) ------------------------------------------- */
/* ctor ANU/AcceleratedMipNeRF.cpp:7-50 (+ AcceleratedMLP ctor ANU/AcceleratedMLP.cpp:168-213):
 * allocates every step buffer; Glorot-uniform weights, zero biases from cfg->seed (A-D7). */
int nerf_mipnerf_create(const nerf_config* cfg, nerf_mipnerf** out);
int nerf_mipnerf_destroy(nerf_mipnerf* h); /* dtor .cpp:151-176 (without the cudaDeviceReset) */
int nerf_mipnerf_num_tensors(const nerf_mipnerf* h, int* n); /* 2*(depth+depth_cond+2) = 22 */
/* GetLayerSizes ANU/AcceleratedMipNeRF.cpp:146-149 -> ANU/AcceleratedMLP.cpp:131-154: W0..W10,b0..b10 */
int nerf_mipnerf_get_layer_sizes(const nerf_mipnerf* h, int* sizes, int* n);
/* mlp.allParams (ANU/AcceleratedMLP.h:25): n device pointers, views into ONE flat allocation */
int nerf_mipnerf_all_params(nerf_mipnerf* h, float** dev_ptrs);
int nerf_mipnerf_all_gradients(nerf_mipnerf* h, float** dev_ptrs); /* mlp.allGradients (.h:24) */
int nerf_mipnerf_flat_params(nerf_mipnerf* h, float** params_dev, float** grads_dev, long* n);
int nerf_mipnerf_set_params(nerf_mipnerf* h, const float* flat_host, long n); /* blob W0..W10,b0..b10 */
int nerf_mipnerf_get_params(nerf_mipnerf* h, float* flat_host, long n);
int nerf_mipnerf_get_gradients(nerf_mipnerf* h, float* flat_host, long n);
/* target pixels for the built-in MSE gradient (replaces the pixel upload of
 * ANU/AcceleratedGradientCalculator.cpp:23 when no callback is given) */
int nerf_mipnerf_set_pixels(nerf_mipnerf* h, const float* pixels3, int n_rays);
/* explicit sampling uniforms u[level][n_rays][S+1] (parity runs); NULL returns to Philox(seed,step) */
int nerf_mipnerf_set_sampling_uniforms(nerf_mipnerf* h, const float* u, int n_rays);
int nerf_mipnerf_set_step(nerf_mipnerf* h, uint32_t step); /* Philox counter word */
/* GetGradient ANU/AcceleratedMipNeRF.cpp:52-144: H2D of the ray batch, per level {sample t, cast_rays,
 * encode, MLP forward, volumetric_rendering}, per level {loss gradient, volumetric_rendering_gradient},
 * per level MLP backward.  Gradients are the SUM over both levels (A-D5).  cb == NULL uses the built-in
 * MSE of .cu:347-361 against nerf_mipnerf_set_pixels.  grad_dev_ptrs (may be NULL) receives the table. */
int nerf_mipnerf_get_gradient(nerf_mipnerf* h, const float* origins3, const float* directions3,
                              const float* radii, const float* nears, const float* fars,
                              const float* loss_mults, int n_rays, nerf_output_gradient_cb cb,
                              void* user, float** grad_dev_ptrs);
/* same step with the ray batch + pixels already resident on the device (bench `value` path) */
int nerf_mipnerf_get_gradient_dev(nerf_mipnerf* h, const float* origins3_dev,
                                  const float* directions3_dev, const float* radii_dev,
                                  const float* nears_dev, const float* fars_dev,
                                  const float* loss_mults_dev, const float* pixels3_dev, int n_rays);
/* forward only (SN/MipNerfModel.cs:36-97), any n_rays (processed in passes of chunk_rays rays); outputs of the
 * LAST level on the host; any output may be NULL.  Evaluation is deterministic whatever cfg.randomized says
 * (bin midpoints, no jitter; explicit sampling uniforms are ignored), so two renders of the same rays are
 * bitwise equal and do not consume training RNG counters.  The reference has no render entry (SN/Dataset.cs:107). */
int nerf_mipnerf_render(nerf_mipnerf* h, const float* origins3, const float* directions3,
                        const float* radii, const float* nears, const float* fars, long n_rays,
                        float* rgb3, float* depth, float* acc);
int nerf_mipnerf_render_dev(nerf_mipnerf* h, const float* origins3_dev, const float* directions3_dev,
                            const float* radii_dev, const float* nears_dev, const float* fars_dev,
                            long n_rays, float* rgb3_dev, float* depth_dev, float* acc_dev);
/* One view (or a pixel range of it, row-major y * width + x) from a camera pose: c2w12 = 3 x 4 row-major [R | t] on the host.
 * The rays are generated on the device chunk by chunk with the arithmetic of Dataset.GenerateRays (SN/Dataset.cs:111-176), so the
 * only per-view host->device traffic is the pose.  edge_mode 0: the reference's radius at the last column (a pixel differenced
 * with itself: 0, :151); 1: the left neighbour's difference (mip-NeRF).  Outputs on the host, or device pointers if
 * outputs_on_device; any may be NULL. */
int nerf_mipnerf_render_view(nerf_mipnerf* h, const float* c2w12, float focal, int width, int height, float near_,
                             float far_, int edge_mode, long first_pixel, long n_pixels, float* rgb3, float* depth,
                             float* acc, int outputs_on_device);
/* per-level results of the last get_gradient / render chunk (device pointers, library-owned) */
int nerf_mipnerf_level_outputs(nerf_mipnerf* h, int level, uint64_t* comp_rgb_dev, uint64_t* depth_dev,
                               uint64_t* acc_dev, uint64_t* weights_dev, uint64_t* t_vals_dev);
/* per-level MSE of the last get_gradient (SN/Program.cs:64); total = sum_l lambda_l*loss_l */
int nerf_mipnerf_get_loss(nerf_mipnerf* h, float* loss_per_level, float* total);
int nerf_mipnerf_synchronize(nerf_mipnerf* h);
/* number of kernels launched by this handle since creation (bench `gpu_launches`) */
int nerf_mipnerf_launch_count(nerf_mipnerf* h, long* n);

/* the CUDA stream every call on this handle enqueues on (a cudaStream_t), for event timing by the caller */
int nerf_mipnerf_stream(nerf_mipnerf* h, uint64_t* stream);
/* in-stream CUDA-event timing per kernel family (off by default; bench.py's roofline numbers) */
int nerf_mipnerf_set_profiling(nerf_mipnerf* h, int on);
int nerf_mipnerf_read_profile(nerf_mipnerf* h, int max_cat, int* n_cat, const char** names, double* ms,
                              long* launches, int reset);

/* ---- AcceleratedMLP (ANU/AcceleratedMLP.h:7-45), reached through the model handle ------------------ */
/* get_output ANU/AcceleratedMLP.cpp:214-255: enc_pos_dev [n_rays*S, 6*deg_point], enc_dir_dev
 * [n_rays*S, 3+6*deg_view] fp32 -> (density, rgb) device pointers — post-activation softplus / sigmoid
 * like the reference heads (.cu:60,73), named explicitly (A-D2). */
int nerf_mlp_get_output(nerf_mipnerf* h, const float* enc_pos_dev, const float* enc_dir_dev, int level,
                        int n_rays, uint64_t* density_dev, uint64_t* rgb_dev);
/* get_gradient ANU/AcceleratedMLP.cpp:256-321: color_grad_dev [M,3], density_grad_dev [M] are dL/d of
 * the POST-activation outputs; consumes the activations cached by the last get_output(level) and
 * ACCUMULATES into the flat gradient (reset_gradients zeroes it). */
int nerf_mlp_get_gradient(nerf_mipnerf* h, const float* color_grad_dev, const float* density_grad_dev,
                          int level, float** grad_dev_ptrs);
int nerf_mlp_reset_gradients(nerf_mipnerf* h, int level); /* ANU/AcceleratedMLP.cpp:113-129 (A-D4) */
/* Parity hook on the cached activations (the reference exposes them as the weighted_sums_ / outputs_ buffers,
 * ANU/AcceleratedMLP.h:35-44): the ReLU masks the backward pass of `level` uses, hidden layer `layer` (trunk
 * 0..depth-1, then the condition layers), as a device bit plane [rows, words_per_row] — bit j of word c of row m
 * <=> output[m, 32 c + j] > 0.  Tensor-core precision modes only (NERF_ERR_STATE otherwise). */
int nerf_mlp_relu_bits(nerf_mipnerf* h, int level, int layer, uint64_t* bits_dev, int* words_per_row);

/* ---- AcceleratedAdamOptimizer (ANU/AcceleratedAdamOptimizer.h:5-20) -------------------------------- */
int nerf_adam_create(const int* sizes, int n, int eps_mode, int device, nerf_adam** out); /* .cpp:6-21, m=v=0 (A-D14) */
/* step .cpp:23-41: params/grads are the n-entry device pointer tables; one launch per contiguous run
 * (a single launch when they are views of one flat buffer, as nerf_mipnerf_all_params returns). */
int nerf_adam_step(nerf_adam* a, float** params_dev, float** grads_dev, float lr);
int nerf_adam_state(nerf_adam* a, float** m_dev, float** v_dev, long* n, int* iteration);
int nerf_adam_set_state(nerf_adam* a, const float* m_host, const float* v_host, long n, int iteration);
int nerf_adam_destroy(nerf_adam* a);

/* ---- AcceleratedGradientCalculator (ANU/AcceleratedGradientCalculator.h:8-17) ---------------------- */
int nerf_gradcalc_create(int batch, int n_levels, float coarse_loss_mult, int device, nerf_gradcalc** out);
/* get_output_gradient .cpp:18-30: uploads pixels (dst/src fixed, A-D13), launches the MSE derivative,
 * returns the device address of the level's gradient buffer. */
int nerf_gradcalc_get_output_gradient(nerf_gradcalc* g, uint64_t comp_rgb_dev, const float* pixels3,
                                      int n, uint64_t loss_mults_dev, float loss_mult_sum, int level,
                                      uint64_t* grad_dev);
int nerf_gradcalc_destroy(nerf_gradcalc* g);

/* ---- OutputRetriever (ANU/OutputRetriever.h:7-11, .cpp:6-14) --------------------------------------- */
int nerf_retrieve_output(uint64_t dev, int n_float3, float* host_out);

/* ---- fused training step (host part of SN/Program.cs:48-62) ---------------------------------------- */
/* get_gradient (built-in MSE) -> [allreduce when a communicator is attached] -> Adam, one call, no
 * host synchronisation except the optional loss read-back (loss_out may be NULL). */
int nerf_mipnerf_train_step(nerf_mipnerf* h, nerf_adam* a, const float* origins3,
                            const float* directions3, const float* radii, const float* nears,
                            const float* fars, const float* loss_mults, const float* pixels3, int n_rays,
                            float lr, float* loss_out);
int nerf_mipnerf_train_step_dev(nerf_mipnerf* h, nerf_adam* a, const float* origins3_dev,
                                const float* directions3_dev, const float* radii_dev,
                                const float* nears_dev, const float* fars_dev,
                                const float* loss_mults_dev, const float* pixels3_dev, int n_rays,
                                float lr, float* loss_out);

/* ---- multi-GPU: rays sharded per rank, ONE exchange per step -------------------------- */
#define NERF_COMM_ID_BYTES 128
int nerf_comm_get_unique_id(void* id_out); /* rank 0; ship the bytes to the other ranks */
/* attach rank `rank` of `world` to the model.  A step then has exactly ONE collective: each rank accumulates its
 * gradient UN-normalised (g = 2 lm (rgb - pix) lambda), the local sum(loss_mults) and the per-level loss numerators
 * ride as the last 1 + n_levels floats of the flat gradient buffer, one ncclAllReduce(sum, fp32) covers all of it, and
 * the global 1 / sum(loss_mults) is applied inside the Adam pass (train_step) or by nerf_mipnerf_allreduce_gradients
 * (GetGradient + explicit allreduce + nerf_adam_step).  After either, the gradient buffers hold the global mean
 * gradient and nerf_mipnerf_get_loss returns the GLOBAL loss on every rank. */
int nerf_mipnerf_comm_init(nerf_mipnerf* h, const void* id, int rank, int world);
int nerf_mipnerf_allreduce_gradients(nerf_mipnerf* h);
int nerf_mipnerf_comm_destroy(nerf_mipnerf* h);

/* ---- per-stage entry points: one per hot-path kernel of .cu, same arguments + explicit sizes ------- */
/* All pointers are DEVICE pointers; work is enqueued on the legacy default stream and synchronised. */
int nerf_get_sample_t_vals(const float* nears, const float* fars, const float* u, int R, int S,
                           int randomized, float* t_vals);                       /* .cu:222-242 */
int nerf_get_resampled_t_vals(const float* t_vals, const float* weights, const float* u, int R, int S,
                              float padding, int randomized, float* new_t_vals); /* .cu:246-291 */
int nerf_cast_rays(const float* t_vals, const float* origins3, const float* directions3,
                   float* means3, float* covs3, const float* radii, int R, int S); /* .cu:292-317 */
/* .cu:187-221; directions3 is per RAY [R,3]; enc_dir is per SAMPLE [R*S, 3+6*deg_view] like the
 * buffer the reference allocates (ANU/AcceleratedMipNeRF.cpp:36) */
int nerf_encode_input_data(const float* means3, const float* covs3, const float* directions3,
                           float* enc_pos, float* enc_dir, int R, int S, int deg_point, int deg_view);
/* .cu:36-90: act 0 relu, 1 sigmoid, 2 softplus, 3 identity; in_b may be NULL (k_b = 0) */
int nerf_apply_layer(const float* in_a, const float* in_b, const float* weights, const float* biases,
                     float* outputs, float* weighted_sums, long M, int n, int k_a, int k_b, int act);
/* .cu:91-182: grads are ACCUMULATED (+=) like the reference's atomicAdd; in_a_grads may be NULL */
int nerf_backpropagate_layer(const float* in_a, const float* in_b, const float* weights,
                             const float* weighted_sums, const float* output_grads, float* in_a_grads,
                             float* weight_grads, float* bias_grads, long M, int n, int k_a, int k_b,
                             int act);
/* .cu:318-344 (+ depth/acc, A-D11); depth/acc/weights may be NULL */
int nerf_volumetric_rendering(const float* rgb3, const float* density, const float* t_vals,
                              const float* directions3, float* comp_rgb3, float* depth, float* acc,
                              float* weights, int R, int S, int white_bkgd);
/* .cu:347-361 */
int nerf_get_output_gradient(const float* comp_rgb3, const float* pixels3, const float* loss_mults,
                             float* comp_rgb_grad3, float loss_mult_sum, float level_mult, int R);
/* .cu:362-402, recomputing alpha/T/w from (density, t) instead of reading caches */
int nerf_volumetric_rendering_gradient(const float* comp_rgb_grad3, const float* rgb3,
                                       const float* density, const float* t_vals,
                                       const float* directions3, float* color_grad3, float* density_grad,
                                       int R, int S, int white_bkgd, int last_sample_mode);
/* The two compositing stages as the model runs them: enqueued on `stream` (a cudaStream_t; NULL = default stream)
 * WITHOUT a synchronize, optionally with the output activations of SN/MipNerfModel.cs:81-83 fused in
 * (raw != 0: density = softplus(raw + density_bias), rgb = sigmoid(raw)(1 + 2 rgb_padding) - rgb_padding, and the
 * gradients come out w.r.t. the RAW head outputs, SN/MipNerfModel.cs:184-189).  Back-to-back launches on one
 * stream are how bench.py measures these kernels against the HBM roofline. */
int nerf_volumetric_rendering_async(const float* rgb3, const float* density, const float* t_vals,
                                    const float* directions3, float* comp_rgb3, float* depth, float* acc,
                                    float* weights, int R, int S, int white_bkgd, int raw, float density_bias,
                                    float rgb_padding, void* stream);
int nerf_volumetric_rendering_gradient_async(const float* comp_rgb_grad3, const float* rgb3,
                                             const float* density, const float* t_vals,
                                             const float* directions3, float* color_grad3,
                                             float* density_grad, int R, int S, int white_bkgd,
                                             int last_sample_mode, int raw, float density_bias,
                                             float rgb_padding, void* stream);
/* .cu:403-416 */
int nerf_adam_optimizer_step(float* variables, const float* gradients, float* m, float* v, float lr,
                             float beta1, float beta2, float inv_1_minus_beta1_pow,
                             float inv_1_minus_beta2_pow, long size, int eps_mode);


/* ---- resident dataset, image metrics, schedule, checkpoints ----------------------- */
#define NERF_RECORD_BYTES 64 /* SN/BinDataset.cs:35-49: o(3) d(3) viewdir(3) radius near far lossmult rgb(3) */
typedef struct nerf_dataset nerf_dataset;
/* Replaces BinDataset (SN/BinDataset.cs:10-52): the records live in device memory once. n_records < 2^32. */
int nerf_dataset_create(const void* records, long n_records, int device, nerf_dataset** out);
int nerf_dataset_load(const char* path, int device, nerf_dataset** out); /* train_data.bin, SN/Program.cs:23 */
int nerf_dataset_size(const nerf_dataset* ds, long* n);                  /* SN/BinDataset.cs:15 */
/* Dataset.GenerateRays (SN/Dataset.cs:111-176) for pixels [first_pixel, first_pixel + n_pixels) of one camera into device
 * arrays (what nerf_mipnerf_render_view feeds itself with); c2w12 on the host. */
int nerf_generate_rays(const float* c2w12, float focal, int width, int height, float near_, float far_, int edge_mode,
                       long first_pixel, long n_pixels, float* origins3_dev, float* directions3_dev, float* radii_dev,
                       float* nears_dev, float* fars_dev);
int nerf_dataset_destroy(nerf_dataset* ds);
/* BinDataset.Next (SN/BinDataset.cs:21-25): batch indices drawn with replacement — here from Philox4x32-10 with
 * counter (first_slot + i, 0, step, 0x0DA7A5E7), key = seed, index = floor(word0 * n / 2^32) — instead of
 * System.Random; this call exports them for parity checks. */
int nerf_dataset_draw_indices(nerf_dataset* ds, uint64_t seed, uint32_t step, uint32_t first_slot,
                              int n_rays, int64_t* idx_host);
/* LoadBatch (SN/BinDataset.cs:27-51) on the device: idx_host == NULL uses the draw above. */
int nerf_dataset_gather(nerf_dataset* ds, const int64_t* idx_host, uint64_t seed, uint32_t step,
                        uint32_t first_slot, int n_rays, float* origins3_dev, float* directions3_dev,
                        float* radii_dev, float* nears_dev, float* fars_dev, float* loss_mults_dev,
                        float* pixels3_dev);
/* One iteration of Train() (SN/Program.cs:28-45) with the batch drawn and assembled on the device. */
int nerf_mipnerf_train_step_dataset(nerf_mipnerf* h, nerf_adam* a, nerf_dataset* ds, int n_rays,
                                    uint64_t sampler_seed, float lr, float* loss_out);
/* mse and MseToPsnr (SN/MipHelpers.cs:672) of two float arrays (host pointers unless on_device). */
int nerf_image_error(const float* a, const float* b, long n_floats, int on_device, double* mse,
                     double* psnr);
/* ComputeSsim / ComputeSsimAverage (SN/MipHelpers.cs:688-737): 2-D normalised Gaussian window (filter_size x filter_size,
 * odd, <= 15; reference defaults 11, 1.5, k1 = 0.01, k2 = 0.03), ZERO-padded borders (VectorImage.Convolve, :903-927),
 * variances AND covariance clamped at 0 (:705-712), mean over pixels and the 3 channels.  Images are [height, width, 3]
 * floats; ssim_map (may be NULL) receives the per-pixel, per-channel map. */
int nerf_image_ssim(const float* a, const float* b, int width, int height, float max_val, int filter_size,
                    float filter_sigma, float k1, float k2, int on_device, double* ssim_mean, float* ssim_map);
/* LearningRateDecay (SN/MipHelpers.cs:758-773). */
float nerf_learning_rate_decay(int step, float lr_init, float lr_final, int max_steps,
                               int lr_delay_steps, float lr_delay_mult);
/* Parameters + Adam m/v/iteration + sampling step counter in one file (Config.SaveEvery, SN/TrainState.cs:62;
 * the reference never implemented the save). `a` may be NULL (parameters only). */
int nerf_checkpoint_save(nerf_mipnerf* h, nerf_adam* a, const char* path);
int nerf_checkpoint_load(nerf_mipnerf* h, nerf_adam* a, const char* path);

#ifdef __cplusplus
}
#endif
#endif /* NERFB200_H */

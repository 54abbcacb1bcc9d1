/*
 * oracle.h — CPU restatement of the NeRF-or-nothing MipNeRF hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker or
 * the CPU baseline.  The product (libnerfb200.so) never links or calls it.
 *
 * PARITY UNPINNED by the reference's own tests: the reference ships no tests,
 * golden vectors or fixtures, and its CPU path is C#
 * (no .NET toolchain in this image).  The pins this oracle has instead:
 *   P1  the reference's own CUDA kernels (ANU/accelerated_functions.cu),
 *       compiled unmodified into oracle/_ref and run on the GPU box
 *       (tests/test_ref_kernels_gpu.py);
 *   P2  an fp64 build of every function here (suffix _f64) shadowing the
 *       fp32 build (suffix _f32);
 *   P3  finite-difference / torch-fp64-autograd checks of every gradient
 *       (tests/test_oracle_cpu.py);
 *   P4  analytic identities (sum(w)+T_end = 1, IPE -> sin/cos as var -> 0,
 *       Adam step-1 closed form, Philox4x32-10 known-answer vectors).
 *
 * Path shorthands used in citations:
 *   ANU/ = ScratchNerf/AcceleratedNeRFUtils/   SN/ = ScratchNerf/ScratchNerf/
 *
 * Every function exists twice: NAME_f32 (float arithmetic, the C# / CUDA
 * operation order, compiled with -ffp-contract=off) and NAME_f64 (double).
 */
#ifndef NERF_ORACLE_H
#define NERF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Hyper-parameters. Defaults follow ANU/AcceleratedMLP.h:10-19 and
 * SN/MipNerfModel.cs:10-28; the switches resolve SURVEY Appendix A items. */
typedef struct orc_config {
  int n_samples;          /* S per level            SN/MipNerfModel.cs:10  */
  int n_levels;           /* 2                      SN/MipNerfModel.cs:11  */
  int net_depth;          /* 8                      ANU/AcceleratedMLP.h:11 */
  int net_width;          /* 256                    ANU/AcceleratedMLP.h:12 */
  int net_depth_condition;/* 1                      ANU/AcceleratedMLP.h:13 */
  int net_width_condition;/* 128                    ANU/AcceleratedMLP.h:14 */
  int skip_layer;         /* 4                      ANU/AcceleratedMLP.h:19 */
  int deg_point;          /* 16 -> 96 IPE inputs    SN/MipNerfModel.cs:17  */
  int deg_view;           /* 4  -> 27 dir inputs    SN/MipNerfModel.cs:18  */
  int white_bkgd;         /* 1                      SN/TrainState.cs:71    */
  int adam_eps_mode;      /* 0: eps inside sqrt (.cu:415); 1: outside (SN/TrainState.cs:34) */
  int last_sample_mode;   /* 0: exact gradient; 1: reference kernel (.cu:375-379 drops sample S-1) */
  int randomized;         /* 1                      SN/TrainState.cs:66    */
  int reserved;
  double density_bias;    /* 0 (.cu:73) or -1 (SN/MipNerfModel.cs:20)      */
  double rgb_padding;     /* 0 (.cu:60) or 0.001 (SN/MipNerfModel.cs:22)   */
  double coarse_loss_mult;/* 0.1 (.cu:345)                                 */
  double resample_padding;/* 0.01 (.cu:243)                                */
} orc_config;

void orc_default_config(orc_config* c);
int  orc_num_layers(const orc_config* c);               /* depth + depth_cond + 2 */
/* sizes[2*L]: W0..W(L-1), b0..b(L-1)  — ANU/AcceleratedMLP.cpp:131-154 */
void orc_layer_sizes(const orc_config* c, int* sizes);
/* per layer: out, in_a (previous activation / encoding), in_b (conjoined encoding or 0) */
void orc_layer_shapes(const orc_config* c, int* out, int* in_a, int* in_b);
long orc_num_params(const orc_config* c);
int  orc_max_threads(void);
void orc_set_threads(int n);

/* Philox4x32-10 (Salmon et al., SC'11) — counter-based RNG replacing the
 * reference's time-seeded cuRAND XORWOW (.cu:17-23; SURVEY A-D7/D8).
 * out[4] = philox(counter[4], key[2]). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* u[r*n + i] in [0,1): counter = (ray0 + r, i, step, level), key = seed; u = (x0 >> 8) * 2^-24 */
void orc_sampling_uniforms(uint64_t seed, uint32_t step, uint32_t level, uint32_t ray0,
                           int n_rays, int n, float* u);
/* Glorot-uniform weights, zero biases (SN/MLP.cs:78-85), Philox stream 0xG10 keyed by seed. */
void orc_init_params(const orc_config* c, uint64_t seed, float* params);

#define ORC_DECL(SUF, REAL)                                                                        \
  /* B.3 level 0 — intent of .cu:222-242 / SN/MipHelpers.cs:611-631 */                            \
  void orc_sample_t_vals##SUF(const REAL* nears, const REAL* fars, const float* u, int R, int S,  \
                              int randomized, REAL* t);                                            \
  /* B.3 level 1 — SN/MipHelpers.cs:634-666, 774-851 */                                           \
  void orc_resample_t_vals##SUF(const REAL* t, const REAL* w, const float* u, int R, int S,       \
                                REAL padding, int randomized, REAL* t_new);                       \
  /* B.1 — .cu:292-317 == SN/MipHelpers.cs:367-402,410-428 */                                     \
  void orc_cast_rays##SUF(const REAL* t, const REAL* o, const REAL* d, const REAL* radii, int R,  \
                          int S, REAL* mean, REAL* cov);                                           \
  /* B.2 IPE — .cu:187-204 == SN/MipHelpers.cs:429-449 (true cos, A-D18) */                       \
  void orc_encode_position##SUF(const REAL* mean, const REAL* cov, long M, int deg, REAL* enc);    \
  /* B.2 direction PE — SN/MipHelpers.cs:337-356 (A-D10); out[R, 3 + 6*deg] */                    \
  void orc_encode_direction##SUF(const REAL* d, int R, int deg, REAL* enc);                        \
  /* MLP forward — SN/MLP.cs:87-136, ANU/AcceleratedMLP.cpp:214-255.                              \
   * enc_dir is per SAMPLE [M, 3+6*deg_view] as in the reference kernels.                         \
   * acts (optional): post-activation outputs of every hidden layer,                              \
   * [M, depth*width + depth_cond*width_cond]. */                                                  \
  void orc_mlp_forward##SUF(const orc_config* c, const REAL* params, const REAL* enc_pos,         \
                            const REAL* enc_dir, long M, REAL* acts, REAL* raw_density,           \
                            REAL* raw_rgb);                                                        \
  /* MLP backward — SN/MLP.cs:138-220 (act'(Z), A-D16), ANU/AcceleratedMLP.cpp:256-321.           \
   * grads[P] is ACCUMULATED into (caller zeroes). */                                              \
  void orc_mlp_backward##SUF(const orc_config* c, const REAL* params, const REAL* enc_pos,        \
                             const REAL* enc_dir, const REAL* acts, const REAL* d_raw_density,    \
                             const REAL* d_raw_rgb, long M, REAL* grads);                          \
  /* activations feeding compositing — B.4; SN/MipNerfModel.cs:81-83; .cu:60,73 */                \
  void orc_output_activations##SUF(const orc_config* c, const REAL* raw_density,                  \
                                   const REAL* raw_rgb, long M, REAL* density, REAL* rgb);        \
  void orc_output_activations_grad##SUF(const orc_config* c, const REAL* raw_density,             \
                                        const REAL* raw_rgb, const REAL* d_density,               \
                                        const REAL* d_rgb, long M, REAL* d_raw_density,           \
                                        REAL* d_raw_rgb);                                          \
  /* B.4 — .cu:318-344 == SN/MipHelpers.cs:472-515 (depth/acc per A-D11) */                       \
  void orc_volumetric_rendering##SUF(const REAL* rgb, const REAL* density, const REAL* t,         \
                                     const REAL* d, int R, int S, int white_bkgd, REAL* comp_rgb, \
                                     REAL* depth, REAL* acc, REAL* weights, REAL* alpha,          \
                                     REAL* transmittance);                                         \
  /* B.5 loss gradient — .cu:347-361 (independent per level, A-D13) */                            \
  void orc_output_gradient##SUF(const REAL* comp_rgb, const REAL* pixels, const REAL* loss_mults, \
                                int R, REAL loss_mult_sum, REAL level_mult, REAL* g);             \
  /* B.5 compositing backward — recurrences of .cu:362-402 / SN/MipHelpers.cs:517-610 */          \
  void orc_volumetric_rendering_gradient##SUF(const REAL* g, const REAL* rgb,                     \
                                              const REAL* density, const REAL* t, const REAL* d,  \
                                              int R, int S, int white_bkgd, int last_sample_mode, \
                                              REAL* d_rgb, REAL* d_density);                       \
  /* B.6 Adam — .cu:403-416 + ANU/AcceleratedAdamOptimizer.cpp:26-28; SN/TrainState.cs:25-37 */   \
  void orc_adam_step##SUF(REAL* p, const REAL* g, REAL* m, REAL* v, long n, REAL lr,              \
                          int iteration, int eps_mode);                                            \
  /* Whole step for one ray batch — SN/MipNerfModel.cs:99-200 + SN/Program.cs:48-64.              \
   * u[level][R,S+1] sampling uniforms.  Outputs (any may be NULL):                               \
   * grads[P] (zeroed here), comp_rgb[L,R,3], depth[L,R], acc[L,R], t_vals[L,R,S+1],              \
   * weights[L,R,S], loss[L] (per-level MSE), returns total loss. */                              \
  double orc_train_gradient##SUF(const orc_config* c, const REAL* params, const REAL* origins,    \
                                 const REAL* dirs, const REAL* radii, const REAL* nears,          \
                                 const REAL* fars, const REAL* loss_mults, const REAL* pixels,    \
                                 const float* u, int R, int with_backward, REAL* grads,           \
                                 REAL* comp_rgb, REAL* depth, REAL* acc, REAL* t_vals,            \
                                 REAL* weights, REAL* loss);

ORC_DECL(_f32, float)
ORC_DECL(_f64, double)

/* LR schedule — SN/MipHelpers.cs:758-773 (float arithmetic as in the C#). */
void orc_generate_rays(const float* c2w, float focal, int W, int H, float near_, float far_, int edge_mode, long first, long n,
                       float* o, float* d, float* radii, float* nears, float* fars); /* SN/Dataset.cs:111-176 */
double orc_ssim(const float* a, const float* b, int W, int H, float max_val, int fs, float sigma, float k1, float k2,
                float* map); /* SN/MipHelpers.cs:688-737 */
float orc_learning_rate_decay(int step, float lr_init, float lr_final, int max_steps,
                              int lr_delay_steps, float lr_delay_mult);

#ifdef __cplusplus
}
#endif
#endif

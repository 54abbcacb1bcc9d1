/*
 * oracle.c — CPU oracle for the NeRF-or-nothing MipNeRF hot path (see oracle.h).
 * TEST INFRASTRUCTURE ONLY; never linked into or called by the product library.
 * Build: oracle/Makefile  (gcc -O3 -fopenmp -ffp-contract=off -march=x86-64-v3).
 */
#include "oracle.h"

#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ config / topology */

void orc_default_config(orc_config* c) {
  memset(c, 0, sizeof(*c));
  c->n_samples = 128; c->n_levels = 2;
  c->net_depth = 8; c->net_width = 256; c->net_depth_condition = 1; c->net_width_condition = 128;
  c->skip_layer = 4; c->deg_point = 16; c->deg_view = 4;
  c->white_bkgd = 1; c->adam_eps_mode = 0; c->last_sample_mode = 0; c->randomized = 1;
  c->density_bias = 0.0; c->rgb_padding = 0.0; /* CUDA-kernel arithmetic (A-D15) */
  c->coarse_loss_mult = 0.1; c->resample_padding = 0.01;
}
int orc_num_layers(const orc_config* c) { return c->net_depth + c->net_depth_condition + 2; }

/* Layer table of SURVEY §2.3: ANU/AcceleratedMLP.cpp:131-154 == SN/MLP.cs:72-77. */
void orc_layer_shapes(const orc_config* c, int* out, int* in_a, int* in_b) {
  const int D = c->net_depth, C = c->net_depth_condition, W = c->net_width, Wc = c->net_width_condition;
  const int P = 6 * c->deg_point, Dd = 3 + 6 * c->deg_view;
  out[0] = W; in_a[0] = P; in_b[0] = 0;
  for (int i = 1; i < D; i++) {
    out[i] = W; in_a[i] = W;
    in_b[i] = (c->skip_layer > 0 && i % c->skip_layer == 0) ? P : 0;
  }
  out[D] = 1; in_a[D] = W; in_b[D] = 0;
  out[D + 1] = Wc; in_a[D + 1] = W; in_b[D + 1] = Dd;
  for (int i = 1; i < C; i++) { out[D + 1 + i] = Wc; in_a[D + 1 + i] = Wc; in_b[D + 1 + i] = 0; }
  out[D + C + 1] = 3; in_a[D + C + 1] = Wc; in_b[D + C + 1] = 0;
}
void orc_layer_sizes(const orc_config* c, int* sizes) {
  int out[64], ia[64], ib[64];
  const int L = orc_num_layers(c);
  orc_layer_shapes(c, out, ia, ib);
  for (int l = 0; l < L; l++) { sizes[l] = out[l] * (ia[l] + ib[l]); sizes[L + l] = out[l]; }
}
long orc_num_params(const orc_config* c) {
  int sizes[128];
  long n = 0;
  orc_layer_sizes(c, sizes);
  for (int l = 0; l < 2 * orc_num_layers(c); l++) n += sizes[l];
  return n;
}
int orc_max_threads(void) { return omp_get_max_threads(); }
void orc_set_threads(int n) { omp_set_num_threads(n); }

/* ------------------------------------------------------------------ Philox4x32-10 */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_sampling_uniforms(uint64_t seed, uint32_t step, uint32_t level, uint32_t ray0, int n_rays,
                           int n, float* u) {
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int r = 0; r < n_rays; r++)
    for (int i = 0; i < n; i++) {
      const uint32_t ctr[4] = {ray0 + (uint32_t)r, (uint32_t)i, step, level};
      uint32_t o[4];
      orc_philox4x32_10(ctr, key, o);
      u[(long)r * n + i] = (float)(o[0] >> 8) * (1.0f / 16777216.0f);
    }
}

/* Glorot-uniform weights (SN/MipHelpers.cs:675: sqrt(6/(in+out))*(2u-1)), zero biases
 * (SN/MLP.cs:78). Deterministic: counter = (flat index, 0, 0, 0x610), key = seed. */
void orc_init_params(const orc_config* c, uint64_t seed, float* params) {
  int out[64], ia[64], ib[64];
  const int L = orc_num_layers(c);
  orc_layer_shapes(c, out, ia, ib);
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  long off = 0;
  for (int l = 0; l < L; l++) {
    const int in = ia[l] + ib[l];
    const float lim = sqrtf(6.0f / (float)(in + out[l]));
    for (long i = 0; i < (long)out[l] * in; i++, off++) {
      const uint32_t ctr[4] = {(uint32_t)off, 0u, 0u, 0x610u};
      uint32_t o[4];
      orc_philox4x32_10(ctr, key, o);
      const float uu = (float)(o[0] >> 8) * (1.0f / 16777216.0f);
      params[off] = lim * (uu * 2.0f - 1.0f);
    }
  }
  for (int l = 0; l < L; l++)
    for (int j = 0; j < out[l]; j++) params[off++] = 0.0f;
}

/* Dataset.GenerateRays (SN/Dataset.cs:111-176), one camera, float arithmetic in the original's order: cameraDirs, rotation * dir
 * (Matrix3x3 rows dotted left to right), radius = |d(x) - d(nextX)| * 2 / sqrt(12) with nextX = x at the last column (:151;
 * edge_mode 1: the left neighbour instead).  c2w: 3 x 4 row-major [R | t]; pixels [first, first + n) row-major. */
static void orc_view_dir_(const float* c, float focal, int W, int H, int x, int y, float* d) {
  const float dx = ((float)x - (float)W * 0.5f + 0.5f) / focal, dy = -(((float)y - (float)H * 0.5f + 0.5f) / focal), dz = -1.0f;
  d[0] = c[0] * dx + c[1] * dy + c[2] * dz;
  d[1] = c[4] * dx + c[5] * dy + c[6] * dz;
  d[2] = c[8] * dx + c[9] * dy + c[10] * dz;
}
void orc_generate_rays(const float* c2w, float focal, int W, int H, float near_, float far_, int edge_mode, long first, long n,
                       float* o, float* d, float* radii, float* nears, float* fars) {
  for (long i = 0; i < n; i++) {
    const long p = first + i;
    const int y = (int)(p / W), x = (int)(p % W);
    float dd[3], dn[3];
    orc_view_dir_(c2w, focal, W, H, x, y, dd);
    int nx = x < W - 1 ? x + 1 : x;
    if (edge_mode == 1 && x == W - 1 && W > 1) nx = x - 1;
    orc_view_dir_(c2w, focal, W, H, nx, y, dn);
    const float ex = dd[0] - dn[0], ey = dd[1] - dn[1], ez = dd[2] - dn[2];
    const float len = sqrtf(ex * ex + ey * ey + ez * ez);
    o[i * 3] = c2w[3]; o[i * 3 + 1] = c2w[7]; o[i * 3 + 2] = c2w[11];
    d[i * 3] = dd[0]; d[i * 3 + 1] = dd[1]; d[i * 3 + 2] = dd[2];
    radii[i] = len * 2 / sqrtf(12.0f);
    nears[i] = near_; fars[i] = far_;
  }
}

/* ComputeSsim (SN/MipHelpers.cs:688-727) with VectorImage.Convolve (:903-927) and CreateGaussianFilter (:739-756), float
 * arithmetic in the original's order: five zero-padded "same" convolutions of a, b, a*a, b*b, a*b with the normalised
 * fs x fs Gaussian (taps kx outer, ky inner), variances and the covariance clamped at 0, map = num / den.  Images are
 * [H, W, 3]; pixel (x, y) of VectorImage[x, y] at (y * W + x) * 3.  The mean (ComputeSsimAverage :728-737) is returned in
 * double precision over the float map (the original accumulates 3*W*H terms in a float). */
static void orc_convolve_(const float* img, int W, int H, const float* filt, int fs, float* out) {
  const int pad = fs / 2, PW = W + 2 * pad, PH = H + 2 * pad;
  float* p = (float*)calloc((size_t)PW * PH * 3, sizeof(float));
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      for (int c = 0; c < 3; c++) p[((size_t)(y + pad) * PW + x + pad) * 3 + c] = img[((size_t)y * W + x) * 3 + c];
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      for (int c = 0; c < 3; c++) {
        float sum = 0.0f;
        for (int kx = 0; kx < fs; kx++)
          for (int ky = 0; ky < fs; ky++) sum += p[((size_t)(y + ky) * PW + x + kx) * 3 + c] * filt[kx * fs + ky];
        out[((size_t)y * W + x) * 3 + c] = sum;
      }
  free(p);
}
double orc_ssim(const float* a, const float* b, int W, int H, float max_val, int fs, float sigma, float k1, float k2,
                float* map) {
  const size_t n = (size_t)W * H * 3;
  float* filt = (float*)malloc((size_t)fs * fs * sizeof(float));
  const int hs = fs / 2;
  float fsum = 0.0f;
  for (int i = 0; i < fs; i++)
    for (int j = 0; j < fs; j++) {
      const float x = (float)(i - hs), y = (float)(j - hs);
      filt[i * fs + j] = expf(-(x * x + y * y) / (2 * sigma * sigma));
      fsum += filt[i * fs + j];
    }
  for (int i = 0; i < fs * fs; i++) filt[i] /= fsum;
  float *mu0 = malloc(n * 4), *mu1 = malloc(n * 4), *s00 = malloc(n * 4), *s11 = malloc(n * 4), *s01 = malloc(n * 4), *tmp = calloc(n, 4);
  orc_convolve_(a, W, H, filt, fs, mu0);
  orc_convolve_(b, W, H, filt, fs, mu1);
  for (size_t i = 0; i < n; i++) tmp[i] = a[i] * a[i];
  orc_convolve_(tmp, W, H, filt, fs, s00);
  for (size_t i = 0; i < n; i++) tmp[i] = b[i] * b[i];
  orc_convolve_(tmp, W, H, filt, fs, s11);
  for (size_t i = 0; i < n; i++) tmp[i] = a[i] * b[i];
  orc_convolve_(tmp, W, H, filt, fs, s01);
  const float c1 = powf(k1 * max_val, 2.0f), c2 = powf(k2 * max_val, 2.0f);
  double total = 0.0;
  for (size_t i = 0; i < n; i++) {
    const float mu00 = mu0[i] * mu0[i], mu11 = mu1[i] * mu1[i], mu01 = mu0[i] * mu1[i];
    const float g00 = fmaxf(s00[i] - mu00, 0.0f), g11 = fmaxf(s11[i] - mu11, 0.0f), g01 = fmaxf(s01[i] - mu01, 0.0f);
    const float num = (mu01 * 2 + c1) * (g01 * 2 + c2);
    const float den = (mu00 + mu11 + c1) * (g00 + g11 + c2);
    const float v = num / den;
    if (map) map[i] = v;
    total += (double)v;
  }
  free(filt); free(mu0); free(mu1); free(s00); free(s11); free(s01); free(tmp);
  return total / (double)n;
}

/* SN/MipHelpers.cs:758-773, float arithmetic. */
float orc_learning_rate_decay(int step, float lr_init, float lr_final, int max_steps,
                              int lr_delay_steps, float lr_delay_mult) {
  float delay_rate = 1.0f;
  if (lr_delay_steps > 0) {
    float p = (float)step / (float)lr_delay_steps;
    p = p < 0.0f ? 0.0f : (p > 1.0f ? 1.0f : p);
    delay_rate = lr_delay_mult + (1.0f - lr_delay_mult) * sinf(0.5f * 3.14159265358979323846f * p);
  }
  float t = (float)step / (float)max_steps;
  t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
  const float log_lerp = expf(logf(lr_init) * (1.0f - t) + logf(lr_final) * t);
  return delay_rate * log_lerp;
}

/* ------------------------------------------------------------------ the two precisions */

#define REAL float
#define SUF _f32
#define R_EXP expf
#define R_LOG logf
#define R_SIN sinf
#define R_COS cosf
#define R_SQRT sqrtf
#define R_POW powf
#include "oracle_impl.h"
#undef REAL
#undef SUF
#undef R_EXP
#undef R_LOG
#undef R_SIN
#undef R_COS
#undef R_SQRT
#undef R_POW

#define REAL double
#define SUF _f64
#define R_EXP exp
#define R_LOG log
#define R_SIN sin
#define R_COS cos
#define R_SQRT sqrt
#define R_POW pow
#include "oracle_impl.h"

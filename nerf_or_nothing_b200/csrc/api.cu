// api.cu — handles, step orchestration and the extern "C" surface declared in include/nerfb200.h.
//
// nerf_mipnerf  <- AcceleratedMipNeRF (ANU/AcceleratedMipNeRF.{h,cpp}) + its AcceleratedMLP field
// nerf_adam     <- AcceleratedAdamOptimizer (ANU/AcceleratedAdamOptimizer.{h,cpp})
// nerf_gradcalc <- AcceleratedGradientCalculator (ANU/AcceleratedGradientCalculator.{h,cpp})
// The reference synchronises the device after each of its ~57 launches per step
// (ANU/AcceleratedMipNeRF.cpp:97-141); here a whole step is enqueued on one stream with no host sync
// unless the caller uses the legacy host callback or asks for the loss.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <exception>
#include <vector>

#include "../../include/nerfb200.h"
#include "kernels.cuh"
#include "mlp.cuh"
#include "profiler.cuh"

namespace nerf {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------ NCCL
// NCCL is reached through dlopen so that the library loads (and single-GPU use works) without it; inside a
// torch process the already-loaded bundled libnccl.so.2 is picked up by SONAME.
struct UidByValue { char internal[NERF_COMM_ID_BYTES]; };  // == ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
struct NcclApi {
  void* so = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, UidByValue, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.so) return 0;
  void* so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!so) so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!so) { set_error("NCCL not found: %s", dlerror()); return NERF_ERR_COMM; }
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(so, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, UidByValue, int))dlsym(so, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(so, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(so, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(so, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    set_error("NCCL symbols missing");
    return NERF_ERR_COMM;
  }
  g_nccl.so = so;
  return 0;
}
#define NERF_NCCL(expr)                                                                              \
  do {                                                                                               \
    int r_ = (expr);                                                                                 \
    if (r_ != 0) {                                                                                   \
      set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?");    \
      return NERF_ERR_COMM;                                                                          \
    }                                                                                                \
  } while (0)
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

static int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device available (%s): libnerfb200 has no CPU path", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    return NERF_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) { set_error("device %d out of range (%d devices)", device, n); return NERF_ERR_INVALID; }
  cudaDeviceProp p;
  NERF_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) {
    set_error("device %d is sm_%d%d; libnerfb200 is built for sm_100a only", device, p.major, p.minor);
    return NERF_ERR_NO_DEVICE;
  }
  NERF_CUDA(cudaSetDevice(device));
  return 0;
}

// Glorot-uniform weights (SN/MipHelpers.cs:675), zero biases (SN/MLP.cs:78), Philox(seed) counter = flat index
// with stream word 0x610 — the same stream the CPU oracle draws from, so both start from identical weights.
__global__ void k_init_params(float* p, const long* w_off, const int* out, const int* in, int L, long n_weights,
                              long n_params, uint64_t seed) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_params) return;
  if (i >= n_weights) { p[i] = 0.f; return; }
  int l = 0;
  while (l + 1 < L && i >= w_off[l + 1]) l++;
  const float lim = sqrtf(6.0f / (float)(in[l] + out[l]));
  const uint32_t x = philox4x32_10_w0((uint32_t)i, 0u, 0u, 0x610u, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float u = (float)(x >> 8) * (1.0f / 16777216.0f);
  p[i] = lim * (u * 2.0f - 1.0f);
}

}  // namespace nerf

using namespace nerf;

// ================================================================================================ handles

struct nerf_mipnerf {
  nerf_config cfg{};
  MlpShape shape;
  MlpEngine* mlp = nullptr;
  cudaStream_t st = nullptr;
  int Rmax = 0, Rc = 0, S = 0, NL = 0;
  float *params = nullptr, *grads = nullptr;
  std::vector<float*> param_ptrs, grad_ptrs;
  std::vector<int> sizes;
  // ray batch: one device allocation + one pinned staging buffer, 16 floats per ray (SN/BinDataset.cs:40-49 order)
  float *rays_dev = nullptr, *rays_pinned = nullptr;
  const float *origins = nullptr, *dirs = nullptr, *radii = nullptr, *nears = nullptr, *fars = nullptr,
              *lm = nullptr, *pixels = nullptr;
  float* pixels_own = nullptr;
  bool have_pixels = false;
  struct Level {
    float *t = nullptr, *raw_density = nullptr, *raw_rgb = nullptr, *weights = nullptr;
    float *d_raw_density = nullptr, *d_raw_rgb = nullptr;
    float *comp_rgb = nullptr, *depth = nullptr, *acc = nullptr, *g = nullptr;  // full Rmax
    float *density_out = nullptr, *rgb_out = nullptr;                              // stand-alone MLP API only
    long last_rows = 0;
  };
  std::vector<Level> lv;
  // device: [0] = sum(loss_mults), [1..NL] = per-level loss.  These are the floats right after the n_params gradients
  // in the SAME allocation, so that the data-parallel step moves gradient + normaliser + losses in one allreduce.
  float* scalars = nullptr;
  float* scalars_pinned = nullptr;  // host mirror
  float* red_scratch = nullptr;     // partials + ticket of the multi-block deterministic reductions
  float* user_u = nullptr;          // explicit sampling uniforms [NL, user_u_rays, S+1]
  int user_u_rays = 0;
  bool grads_unnormalised = false;  // data parallel, between get_gradient and the allreduce / Adam
  uint32_t step = 0;
  uint32_t ray_offset = 0;  // global index of this rank's first ray (Philox counter)
  void* comm = nullptr;
  int rank = 0, world = 1;
  std::vector<void*> owned;
  Profiler* prof = nullptr;
};

struct nerf_adam {
  int device = 0;
  std::vector<int> sizes;
  long n = 0;
  float *m = nullptr, *v = nullptr;
  int iteration = 0, eps_mode = 0;
  static constexpr float beta1 = 0.9f, beta2 = 0.999f;  // ANU/AcceleratedAdamOptimizer.h:15-16
};

struct nerf_gradcalc {
  int device = 0, batch = 0, n_levels = 0;
  float coarse_mult = 0.1f;
  float *grad = nullptr, *pixels = nullptr;  // grad: [n_levels, batch, 3]
};

namespace {

constexpr int kGradTail = 16;  // floats behind the gradients: [0] sum(loss_mults), [1..n_levels] per-level loss (n_levels <= 4)

struct ProfActivate {  // routes ProfScope hooks to this handle's profiler for the duration of a call
  Profiler* prev;
  explicit ProfActivate(nerf_mipnerf* h) : prev(g_prof) { g_prof = h->prof; }
  ~ProfActivate() { g_prof = prev; }
};

int dalloc(nerf_mipnerf* h, float** p, size_t n) {
  NERF_CUDA(cudaMalloc(p, (n ? n : 1) * sizeof(float)));
  h->owned.push_back(*p);
  return 0;
}

OutputAct model_act(const nerf_mipnerf* h) {
  OutputAct a;
  a.raw = true; a.density_bias = h->cfg.density_bias; a.rgb_padding = h->cfg.rgb_padding;
  return a;
}

int validate(const nerf_config& c) {
  if (c.n_rays <= 0 || c.n_levels < 1 || c.n_levels > 4) { set_error("config: n_rays=%d n_levels=%d", c.n_rays, c.n_levels); return NERF_ERR_INVALID; }
  if (!(c.n_samples == 32 || c.n_samples == 64 || c.n_samples == 128 || c.n_samples == 256)) { set_error("config: n_samples must be 32/64/128/256 (got %d)", c.n_samples); return NERF_ERR_INVALID; }
  if (c.net_depth < 1 || c.net_depth > 32 || c.net_depth_condition < 1 || c.net_depth_condition > 8) { set_error("config: net_depth=%d net_depth_condition=%d", c.net_depth, c.net_depth_condition); return NERF_ERR_INVALID; }
  if (c.net_width < 16 || c.net_width % 16 || c.net_width_condition < 16 || c.net_width_condition % 16) { set_error("config: widths must be multiples of 16"); return NERF_ERR_INVALID; }
  if (c.deg_point < 1 || c.deg_point > 24 || c.deg_view < 0 || c.deg_view > 8) { set_error("config: deg_point=%d deg_view=%d", c.deg_point, c.deg_view); return NERF_ERR_INVALID; }
  if (c.precision < 0 || c.precision > 2) { set_error("config: precision=%d", c.precision); return NERF_ERR_INVALID; }
  return 0;
}

// forward of all levels for rays [c0, c0+rc) of the current batch
// for_training = false (render): deterministic sampling (SN/MipNerfModel.cs:36-97 evaluates with randomized = false),
// explicit uniforms ignored, no activation caches
int forward_chunk(nerf_mipnerf* h, int c0, int rc, bool for_training = true) {
  const nerf_config& c = h->cfg;
  const int S = h->S;
  const int randomized = for_training ? c.randomized : 0;
  for (int l = 0; l < h->NL; l++) {
    auto& L = h->lv[l];
    SampleRng rng;
    rng.u = (h->user_u && for_training) ? h->user_u + ((size_t)l * h->user_u_rays + c0) * (S + 1) : nullptr;
    rng.seed = c.seed; rng.step = h->step; rng.level = (uint32_t)l; rng.ray0 = h->ray_offset + (uint32_t)c0;
    { ProfScope ps(PC_SAMPLE, h->st);
    if (l == 0) NERF_TRY(launch_sample_t_vals(h->nears + c0, h->fars + c0, rng, rc, S, randomized, L.t, h->st));
    else NERF_TRY(launch_resample_t_vals(h->lv[l - 1].t, h->lv[l - 1].weights, rng, rc, S, c.resample_padding, randomized, L.t, h->st)); }
    // cast_rays + encode_input_data (.cu:292-317, 187-221): inside the fused MLP kernel where the engine can (no kernel of
    // their own, no HBM round trip of the encodings), else one fused encode kernel into the engine's planes
    RaySource rs;
    rs.t = L.t; rs.o = h->origins + (size_t)c0 * 3; rs.d = h->dirs + (size_t)c0 * 3; rs.radii = h->radii + c0;
    rs.R = rc; rs.S = S; rs.deg_point = c.deg_point; rs.deg_view = c.deg_view;
    int handled = 0;
    NERF_TRY(h->mlp->forward_from_rays(l, rs, (long)rc * S, h->params, L.raw_density, L.raw_rgb, for_training, h->st, &handled));
    if (!handled) {
      { ProfScope ps(PC_ENCODE, h->st);
      NERF_TRY(launch_cast_encode_fused(L.t, rs.o, rs.d, rs.radii, rc, S, c.deg_point, c.deg_view, h->mlp->encode_targets(l), h->st)); }
      if (for_training) NERF_TRY(h->mlp->forward(l, (long)rc * S, h->params, L.raw_density, L.raw_rgb, h->st));
      else NERF_TRY(h->mlp->forward_only(l, (long)rc * S, h->params, L.raw_density, L.raw_rgb, h->st));
    }
    L.last_rows = (long)rc * S;
    { ProfScope ps(PC_COMPOSITE_FWD, h->st);
    NERF_TRY(launch_composite_fwd(L.raw_rgb, L.raw_density, L.t, h->dirs + (size_t)c0 * 3, rc, S, c.white_bkgd,
                                  model_act(h), L.comp_rgb + (size_t)c0 * 3, L.depth + c0, L.acc + c0, L.weights, h->st)); }
  }
  return 0;
}

int backward_chunk(nerf_mipnerf* h, int c0, int rc, float* const* g_levels) {
  const nerf_config& c = h->cfg;
  const int S = h->S;
  for (int l = 0; l < h->NL; l++) {  // ANU/AcceleratedMipNeRF.cpp:125-134
    auto& L = h->lv[l];
    ProfScope ps(PC_COMPOSITE_BWD, h->st);
    NERF_TRY(launch_composite_bwd(g_levels[l] + (size_t)c0 * 3, L.raw_rgb, L.raw_density, L.t, h->dirs + (size_t)c0 * 3, rc, S,
                                  c.white_bkgd, c.last_sample_mode, model_act(h), L.d_raw_rgb, L.d_raw_density, h->st));
  }
  for (int l = h->NL - 1; l >= 0; l--) {  // SN/MipNerfModel.cs:171; sum over levels (A-D5)
    auto& L = h->lv[l];
    NERF_TRY(h->mlp->backward(l, (long)rc * S, h->params, h->grads, L.d_raw_density, L.d_raw_rgb, h->st));
  }
  return 0;
}

// batch pointers (h->origins ...) must be set.  cb == nullptr -> built-in MSE against h->pixels.
int gradient_core(nerf_mipnerf* h, int n_rays, nerf_output_gradient_cb cb, void* user) {
  const nerf_config& c = h->cfg;
  if (n_rays <= 0 || n_rays > h->Rmax) { set_error("n_rays=%d outside (0, %d]", n_rays, h->Rmax); return NERF_ERR_INVALID; }
  if (!cb && !h->pixels) { set_error("no target pixels: call nerf_mipnerf_set_pixels or pass a callback"); return NERF_ERR_STATE; }
  if (h->user_u && h->user_u_rays < n_rays) { set_error("sampling uniforms cover %d rays < %d", h->user_u_rays, n_rays); return NERF_ERR_INVALID; }
  if (cb && n_rays > h->Rc) { set_error("the host-callback path needs n_rays (%d) <= chunk_rays (%d)", n_rays, h->Rc); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(c.device));
  ProfActivate pa(h);
  // gradients + the scalars behind them in one memset
  NERF_CUDA(cudaMemsetAsync(h->grads, 0, (size_t)(h->shape.n_params + kGradTail) * sizeof(float), h->st));
  { ProfScope ps(PC_MISC, h->st); NERF_TRY(launch_sum(h->lm, n_rays, h->scalars, h->red_scratch, h->st)); }  // sum(loss_mults) as a float (A-D13)
  // Data parallel: NO collective here.  The rank-local gradient is accumulated un-normalised (lm_sum = 1) and the global
  // 1 / sum(loss_mults) is applied after the single allreduce of [gradient | sum(lm) | loss numerators] (finish_step).
  const bool dp = h->comm != nullptr;
  if (dp && cb) { set_error("the host-callback path is single-GPU (the callback sees a rank-local loss_mult_sum)"); return NERF_ERR_STATE; }
  h->grads_unnormalised = dp;
  NERF_TRY(h->mlp->prepare(h->params, h->st));
  std::vector<float*> g(h->NL);
  for (int c0 = 0; c0 < n_rays; c0 += h->Rc) {
    const int rc = n_rays - c0 < h->Rc ? n_rays - c0 : h->Rc;
    NERF_TRY(forward_chunk(h, c0, rc));
    if (cb) {
      // legacy contract (ANU/AcceleratedMipNeRF.cpp:127): host callback per level, in level order
      NERF_CUDA(cudaMemcpyAsync(h->scalars_pinned, h->scalars, sizeof(float), cudaMemcpyDeviceToHost, h->st));
      NERF_CUDA(cudaStreamSynchronize(h->st));
      for (int l = 0; l < h->NL; l++) {
        g[l] = (float*)(uintptr_t)cb((uint64_t)(uintptr_t)h->lv[l].comp_rgb, l, h->scalars_pinned[0], (uint64_t)(uintptr_t)h->lm, user);
        if (!g[l]) { set_error("output-gradient callback returned NULL for level %d", l); return NERF_ERR_INVALID; }
      }
      NERF_CUDA(cudaDeviceSynchronize());  // the callback may have used any stream
    } else {
      for (int l = 0; l < h->NL; l++) {
        const float mult = l < h->NL - 1 ? c.coarse_loss_mult : 1.0f;  // .cu:356
        ProfScope ps(PC_LOSS, h->st);
        NERF_TRY(launch_output_gradient(h->lv[l].comp_rgb + (size_t)c0 * 3, h->pixels + (size_t)c0 * 3, h->lm + c0, rc, 1.0f,
                                        dp ? nullptr : h->scalars, mult, h->lv[l].g + (size_t)c0 * 3, h->scalars + 1 + l,
                                        h->red_scratch, h->st));
        g[l] = h->lv[l].g;
      }
    }
    NERF_TRY(backward_chunk(h, c0, rc, g.data()));
  }
  h->step++;
  return 0;
}

// stage a host ray batch (+ optional pixels) through pinned memory with ONE H2D copy
int upload_batch(nerf_mipnerf* h, const float* o, const float* d, const float* radii, const float* nears,
                 const float* fars, const float* lm, const float* pixels, int n) {
  if (n <= 0 || n > h->Rmax) { set_error("n_rays=%d outside (0, %d]", n, h->Rmax); return NERF_ERR_INVALID; }
  if (!o || !d || !radii || !nears || !fars) { set_error("null ray array"); return NERF_ERR_INVALID; }
  const size_t R = (size_t)h->Rmax;
  float* p = h->rays_pinned;
  // the previous step's H2D copy must have left the staging buffer
  NERF_CUDA(cudaStreamSynchronize(h->st));
  memcpy(p, o, (size_t)n * 12);
  memcpy(p + 3 * R, d, (size_t)n * 12);
  memcpy(p + 6 * R, radii, (size_t)n * 4);
  memcpy(p + 7 * R, nears, (size_t)n * 4);
  memcpy(p + 8 * R, fars, (size_t)n * 4);
  if (lm) memcpy(p + 9 * R, lm, (size_t)n * 4);
  else for (int i = 0; i < n; i++) p[9 * R + i] = 1.0f;
  size_t floats = 10 * R;
  if (pixels) { memcpy(p + 10 * R, pixels, (size_t)n * 12); floats = 13 * R; }
  // contiguous prefix copy: everything up to the last used array (gaps between arrays ride along)
  const size_t used = pixels ? 10 * R + (size_t)n * 3 : 9 * R + (size_t)n;
  (void)floats;
  NERF_CUDA(cudaMemcpyAsync(h->rays_dev, p, used * sizeof(float), cudaMemcpyHostToDevice, h->st));
  float* b = h->rays_dev;
  h->origins = b; h->dirs = b + 3 * R; h->radii = b + 6 * R; h->nears = b + 7 * R; h->fars = b + 8 * R; h->lm = b + 9 * R;
  if (pixels) h->pixels = b + 10 * R;
  else h->pixels = h->have_pixels ? h->pixels_own : nullptr;
  return 0;
}

void set_batch_dev(nerf_mipnerf* h, const float* o, const float* d, const float* radii, const float* nears,
                   const float* fars, const float* lm, const float* pixels) {
  h->origins = o; h->dirs = d; h->radii = radii; h->nears = nears; h->fars = fars; h->lm = lm;
  h->pixels = pixels ? pixels : (h->have_pixels ? h->pixels_own : nullptr);
}

int adam_step_flat(nerf_adam* a, float* p, const float* g, long n, long off, float lr, float gs, cudaStream_t st) {
  const float inv1 = 1.0f / (1.0f - powf(nerf_adam::beta1, (float)a->iteration));  // ANU/AcceleratedAdamOptimizer.cpp:27-28
  const float inv2 = 1.0f / (1.0f - powf(nerf_adam::beta2, (float)a->iteration));
  return launch_adam(p, g, a->m + off, a->v + off, n, lr, nerf_adam::beta1, nerf_adam::beta2, inv1, inv2, a->eps_mode, gs, st);
}

int read_loss(nerf_mipnerf* h, float* loss_per_level, float* total) {
  NERF_CUDA(cudaMemcpyAsync(h->scalars_pinned, h->scalars, (size_t)(1 + h->NL) * sizeof(float), cudaMemcpyDeviceToHost, h->st));
  NERF_CUDA(cudaStreamSynchronize(h->st));
  float tot = 0.f;
  // data parallel: the tail holds allreduced NUMERATORS sum(lm |rgb - pix|^2) and the global sum(lm)
  const float div = h->comm ? h->scalars_pinned[0] : 1.0f;
  for (int l = 0; l < h->NL; l++) {
    const float v = h->scalars_pinned[1 + l] / div;
    if (loss_per_level) loss_per_level[l] = v;
    tot += (l < h->NL - 1 ? h->cfg.coarse_loss_mult : 1.0f) * v;  // SN/Program.cs:81
  }
  if (total) *total = tot;
  return 0;
}

}  // namespace

// ================================================================================================ C ABI

extern "C" {

const char* nerf_last_error(void) { return g_err; }
int nerf_version(void) { return 100; }

void nerf_default_config(nerf_config* c) {
  memset(c, 0, sizeof(*c));
  c->n_rays = 1024; c->n_samples = 128; c->n_levels = 2;                                   // ANU/helpers.h:16-18
  c->net_depth = 8; c->net_width = 256; c->net_depth_condition = 1; c->net_width_condition = 128;  // ANU/AcceleratedMLP.h:11-14
  c->skip_layer = 4; c->deg_point = 16; c->deg_view = 4;                                    // :19, ANU/helpers.h:19-20
  c->white_bkgd = 1; c->randomized = 1; c->adam_eps_mode = 0; c->last_sample_mode = 0;
  c->precision = NERF_PRECISION_FP32; c->device = 0; c->chunk_rays = 0;
  c->density_bias = 0.f; c->rgb_padding = 0.f; c->coarse_loss_mult = 0.1f; c->resample_padding = 0.01f;
  c->seed = 7;
  c->engine_flags = 0;
}

int nerf_device_count(int* n) {
  int k = 0;
  if (cudaGetDeviceCount(&k) != cudaSuccess) { cudaGetLastError(); k = 0; }
  if (n) *n = k;
  return 0;
}

int nerf_mipnerf_create(const nerf_config* cfg, nerf_mipnerf** out) {
  if (!cfg || !out) { set_error("null argument"); return NERF_ERR_INVALID; }
  *out = nullptr;
  NERF_TRY(validate(*cfg));
  NERF_TRY(check_device(cfg->device));
  nerf_mipnerf* h = new nerf_mipnerf();
  h->cfg = *cfg;
  h->Rmax = cfg->n_rays; h->S = cfg->n_samples; h->NL = cfg->n_levels;
  h->shape.build(cfg->net_depth, cfg->net_width, cfg->net_depth_condition, cfg->net_width_condition, cfg->skip_layer,
                 cfg->deg_point, cfg->deg_view);
  // chunk size: keep the activation cache under ~48 GB of the 180 GB HBM
  {
    const double per_sample = (double)(h->shape.D * h->shape.W + h->shape.C * h->shape.Wc + h->shape.P + 64) * 4.0 * h->NL * 1.5;
    long auto_rc = (long)(48e9 / (per_sample * h->S));
    auto_rc = auto_rc / 256 * 256;
    if (auto_rc < 256) auto_rc = 256;
    long rc = cfg->chunk_rays > 0 ? cfg->chunk_rays : auto_rc;
    h->Rc = (int)(rc < h->Rmax ? rc : h->Rmax);
  }
  auto fail = [&](int s) { nerf_mipnerf_destroy(h); return s; };
  int s;
  if (cudaStreamCreate(&h->st) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(NERF_ERR_NO_DEVICE); }
  const long P = h->shape.n_params;
  if ((s = dalloc(h, &h->params, P))) return fail(s);
  if ((s = dalloc(h, &h->grads, P + kGradTail))) return fail(s);
  h->scalars = h->grads + P;
  if ((s = dalloc(h, &h->red_scratch, NERF_RED_SCRATCH_FLOATS))) return fail(s);
  if (cudaMemset(h->red_scratch, 0, NERF_RED_SCRATCH_FLOATS * sizeof(float)) != cudaSuccess) { set_error("cudaMemset failed"); return fail(NERF_ERR_NO_DEVICE); }
  for (auto& l : h->shape.layers) { h->param_ptrs.push_back(h->params + l.w_off); h->grad_ptrs.push_back(h->grads + l.w_off); h->sizes.push_back(l.out * (l.in_a + l.in_b)); }
  for (auto& l : h->shape.layers) { h->param_ptrs.push_back(h->params + l.b_off); h->grad_ptrs.push_back(h->grads + l.b_off); h->sizes.push_back(l.out); }
  const size_t R = (size_t)h->Rmax, Mc = (size_t)h->Rc * h->S;
  if ((s = dalloc(h, &h->rays_dev, 13 * R))) return fail(s);
  if ((s = dalloc(h, &h->pixels_own, 3 * R))) return fail(s);
  if (cudaMallocHost(&h->rays_pinned, 13 * R * sizeof(float)) != cudaSuccess) { set_error("cudaMallocHost failed"); return fail(NERF_ERR_NO_DEVICE); }
  if (cudaMallocHost(&h->scalars_pinned, 16 * sizeof(float)) != cudaSuccess) { set_error("cudaMallocHost failed"); return fail(NERF_ERR_NO_DEVICE); }
  h->lv.resize(h->NL);
  for (auto& L : h->lv) {
    if ((s = dalloc(h, &L.t, (size_t)h->Rc * (h->S + 1)))) return fail(s);
    if ((s = dalloc(h, &L.raw_density, Mc))) return fail(s);
    if ((s = dalloc(h, &L.raw_rgb, Mc * 3))) return fail(s);
    if ((s = dalloc(h, &L.weights, Mc))) return fail(s);
    if ((s = dalloc(h, &L.d_raw_density, Mc))) return fail(s);
    if ((s = dalloc(h, &L.d_raw_rgb, Mc * 3))) return fail(s);
    if ((s = dalloc(h, &L.comp_rgb, R * 3))) return fail(s);
    if ((s = dalloc(h, &L.depth, R))) return fail(s);
    if ((s = dalloc(h, &L.acc, R))) return fail(s);
    if ((s = dalloc(h, &L.g, R * 3))) return fail(s);
  }
  h->mlp = cfg->precision == NERF_PRECISION_FP32 ? make_simt_mlp() : make_tc_mlp(cfg->precision == NERF_PRECISION_FP32_TC, cfg->engine_flags);
  if (!h->mlp) { set_error("precision mode %d is not available in this build", cfg->precision); return fail(NERF_ERR_INVALID); }
  if ((s = h->mlp->init(h->shape, (long)Mc, h->NL))) return fail(s);
  // deterministic init (A-D7)
  {
    const int L = h->shape.L;
    std::vector<long> woff(L);
    std::vector<int> o(L), in(L);
    long nw = 0;
    for (int l = 0; l < L; l++) { woff[l] = h->shape.layers[l].w_off; o[l] = h->shape.layers[l].out; in[l] = h->shape.layers[l].in_a + h->shape.layers[l].in_b; nw += (long)o[l] * in[l]; }
    long* d_woff; int *d_o, *d_in;
    if (cudaMalloc(&d_woff, L * sizeof(long)) != cudaSuccess || cudaMalloc(&d_o, L * sizeof(int)) != cudaSuccess || cudaMalloc(&d_in, L * sizeof(int)) != cudaSuccess) { set_error("cudaMalloc failed"); return fail(NERF_ERR_NO_DEVICE); }
    cudaMemcpy(d_woff, woff.data(), L * sizeof(long), cudaMemcpyHostToDevice);
    cudaMemcpy(d_o, o.data(), L * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(d_in, in.data(), L * sizeof(int), cudaMemcpyHostToDevice);
    k_init_params<<<(unsigned)cdiv(P, 256), 256, 0, h->st>>>(h->params, d_woff, d_o, d_in, L, nw, P, cfg->seed);
    count_launch();
    cudaStreamSynchronize(h->st);
    cudaFree(d_woff); cudaFree(d_o); cudaFree(d_in);
    if (cudaGetLastError() != cudaSuccess) { set_error("parameter init failed"); return fail(NERF_ERR_NO_DEVICE); }
  }
  cudaMemsetAsync(h->grads, 0, (P + kGradTail) * sizeof(float), h->st);
  cudaStreamSynchronize(h->st);
  *out = h;
  return 0;
}

int nerf_mipnerf_destroy(nerf_mipnerf* h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  if (h->st) cudaStreamSynchronize(h->st);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  delete h->mlp;
  delete h->prof;
  for (void* p : h->owned) cudaFree(p);
  if (h->user_u) cudaFree(h->user_u);
  if (h->rays_pinned) cudaFreeHost(h->rays_pinned);
  if (h->scalars_pinned) cudaFreeHost(h->scalars_pinned);
  if (h->st) cudaStreamDestroy(h->st);
  delete h;
  return 0;
}

int nerf_mipnerf_num_tensors(const nerf_mipnerf* h, int* n) {
  if (!h || !n) { set_error("null argument"); return NERF_ERR_INVALID; }
  *n = (int)h->sizes.size();
  return 0;
}
int nerf_mipnerf_get_layer_sizes(const nerf_mipnerf* h, int* sizes, int* n) {
  if (!h) { set_error("null handle"); return NERF_ERR_INVALID; }
  if (n) *n = (int)h->sizes.size();
  if (sizes) memcpy(sizes, h->sizes.data(), h->sizes.size() * sizeof(int));
  return 0;
}
int nerf_mipnerf_all_params(nerf_mipnerf* h, float** p) {
  if (!h || !p) { set_error("null argument"); return NERF_ERR_INVALID; }
  memcpy(p, h->param_ptrs.data(), h->param_ptrs.size() * sizeof(float*));
  return 0;
}
int nerf_mipnerf_all_gradients(nerf_mipnerf* h, float** p) {
  if (!h || !p) { set_error("null argument"); return NERF_ERR_INVALID; }
  memcpy(p, h->grad_ptrs.data(), h->grad_ptrs.size() * sizeof(float*));
  return 0;
}
int nerf_mipnerf_flat_params(nerf_mipnerf* h, float** params, float** grads, long* n) {
  if (!h) { set_error("null handle"); return NERF_ERR_INVALID; }
  if (params) *params = h->params;
  if (grads) *grads = h->grads;
  if (n) *n = h->shape.n_params;
  return 0;
}
int nerf_mipnerf_set_params(nerf_mipnerf* h, const float* flat, long n) {
  if (!h || !flat || n != h->shape.n_params) { set_error("set_params: expected %ld floats", h ? h->shape.n_params : 0L); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  NERF_CUDA(cudaStreamSynchronize(h->st));
  NERF_CUDA(cudaMemcpy(h->params, flat, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}
int nerf_mipnerf_get_params(nerf_mipnerf* h, float* flat, long n) {
  if (!h || !flat || n != h->shape.n_params) { set_error("get_params: expected %ld floats", h ? h->shape.n_params : 0L); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  NERF_CUDA(cudaStreamSynchronize(h->st));
  NERF_CUDA(cudaMemcpy(flat, h->params, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}
int nerf_mipnerf_get_gradients(nerf_mipnerf* h, float* flat, long n) {
  if (!h || !flat || n != h->shape.n_params) { set_error("get_gradients: expected %ld floats", h ? h->shape.n_params : 0L); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  NERF_CUDA(cudaStreamSynchronize(h->st));
  NERF_CUDA(cudaMemcpy(flat, h->grads, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}
int nerf_mipnerf_set_pixels(nerf_mipnerf* h, const float* pixels3, int n_rays) {
  if (!h || !pixels3 || n_rays <= 0 || n_rays > h->Rmax) { set_error("set_pixels: bad arguments"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  NERF_CUDA(cudaStreamSynchronize(h->st));
  NERF_CUDA(cudaMemcpy(h->pixels_own, pixels3, (size_t)n_rays * 12, cudaMemcpyHostToDevice));
  h->have_pixels = true;
  h->pixels = h->pixels_own;
  return 0;
}
int nerf_mipnerf_set_sampling_uniforms(nerf_mipnerf* h, const float* u, int n_rays) {
  if (!h) { set_error("null handle"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  NERF_CUDA(cudaStreamSynchronize(h->st));
  if (h->user_u) { cudaFree(h->user_u); h->user_u = nullptr; h->user_u_rays = 0; }
  if (!u) return 0;
  if (n_rays <= 0 || n_rays > h->Rmax) { set_error("set_sampling_uniforms: n_rays=%d", n_rays); return NERF_ERR_INVALID; }
  const size_t n = (size_t)h->NL * n_rays * (h->S + 1);
  NERF_CUDA(cudaMalloc(&h->user_u, n * sizeof(float)));
  NERF_CUDA(cudaMemcpy(h->user_u, u, n * sizeof(float), cudaMemcpyHostToDevice));
  h->user_u_rays = n_rays;
  return 0;
}
int nerf_mipnerf_set_step(nerf_mipnerf* h, uint32_t step) {
  if (!h) { set_error("null handle"); return NERF_ERR_INVALID; }
  h->step = step;
  return 0;
}

int nerf_mipnerf_get_gradient(nerf_mipnerf* h, const float* o, const float* d, const float* radii, const float* nears,
                              const float* fars, const float* lm, int n_rays, nerf_output_gradient_cb cb, void* user,
                              float** grad_dev_ptrs) {
  if (!h) { set_error("null handle"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  NERF_TRY(upload_batch(h, o, d, radii, nears, fars, lm, nullptr, n_rays));
  NERF_TRY(gradient_core(h, n_rays, cb, user));
  if (grad_dev_ptrs) memcpy(grad_dev_ptrs, h->grad_ptrs.data(), h->grad_ptrs.size() * sizeof(float*));
  return 0;
}

int nerf_mipnerf_get_gradient_dev(nerf_mipnerf* h, const float* o, const float* d, const float* radii, const float* nears,
                                  const float* fars, const float* lm, const float* pixels, int n_rays) {
  if (!h || !o || !d || !radii || !nears || !fars || !lm) { set_error("null argument"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  set_batch_dev(h, o, d, radii, nears, fars, lm, pixels);
  return gradient_core(h, n_rays, nullptr, nullptr);
}

int nerf_mipnerf_render_dev(nerf_mipnerf* h, const float* o, const float* d, const float* radii, const float* nears,
                            const float* fars, long n_rays, float* rgb, float* depth, float* acc) {
  if (!h || !o || !d || !radii || !nears || !fars || n_rays <= 0) { set_error("render: bad arguments"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  ProfActivate pa(h);
  NERF_TRY(h->mlp->prepare(h->params, h->st));
  const int last = h->NL - 1;
  for (long b0 = 0; b0 < n_rays; b0 += h->Rc) {
    const int rc = (int)(n_rays - b0 < h->Rc ? n_rays - b0 : h->Rc);
    set_batch_dev(h, o + b0 * 3, d + b0 * 3, radii + b0, nears + b0, fars + b0, nullptr, nullptr);
    const uint32_t keep = h->ray_offset;
    h->ray_offset = keep + (uint32_t)b0;
    const int s = forward_chunk(h, 0, rc, false);
    h->ray_offset = keep;
    NERF_TRY(s);
    auto& L = h->lv[last];
    if (rgb) NERF_CUDA(cudaMemcpyAsync(rgb + b0 * 3, L.comp_rgb, (size_t)rc * 12, cudaMemcpyDeviceToDevice, h->st));
    if (depth) NERF_CUDA(cudaMemcpyAsync(depth + b0, L.depth, (size_t)rc * 4, cudaMemcpyDeviceToDevice, h->st));
    if (acc) NERF_CUDA(cudaMemcpyAsync(acc + b0, L.acc, (size_t)rc * 4, cudaMemcpyDeviceToDevice, h->st));
  }
  return 0;
}

int nerf_mipnerf_render(nerf_mipnerf* h, const float* o, const float* d, const float* radii, const float* nears,
                        const float* fars, long n_rays, float* rgb, float* depth, float* acc) {
  if (!h || n_rays <= 0) { set_error("render: bad arguments"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  ProfActivate pa(h);
  NERF_TRY(h->mlp->prepare(h->params, h->st));
  const int last = h->NL - 1;
  for (long b0 = 0; b0 < n_rays; b0 += h->Rc) {
    const int rc = (int)(n_rays - b0 < h->Rc ? n_rays - b0 : h->Rc);
    NERF_TRY(upload_batch(h, o + b0 * 3, d + b0 * 3, radii + b0, nears + b0, fars + b0, nullptr, nullptr, rc));
    const uint32_t keep = h->ray_offset;
    h->ray_offset = keep + (uint32_t)b0;
    const int s = forward_chunk(h, 0, rc, false);
    h->ray_offset = keep;
    NERF_TRY(s);
    auto& L = h->lv[last];
    if (rgb) NERF_CUDA(cudaMemcpyAsync(rgb + b0 * 3, L.comp_rgb, (size_t)rc * 12, cudaMemcpyDeviceToHost, h->st));
    if (depth) NERF_CUDA(cudaMemcpyAsync(depth + b0, L.depth, (size_t)rc * 4, cudaMemcpyDeviceToHost, h->st));
    if (acc) NERF_CUDA(cudaMemcpyAsync(acc + b0, L.acc, (size_t)rc * 4, cudaMemcpyDeviceToHost, h->st));
  }
  NERF_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

// One view from a camera pose: the rays of every pixel are generated on the device chunk by chunk (Dataset.GenerateRays,
// SN/Dataset.cs:111-176), so nothing per-ray crosses PCIe on the way in.
int nerf_mipnerf_render_view(nerf_mipnerf* h, const float* c2w12, float focal, int width, int height, float near, float far, int edge_mode,
                             long first_pixel, long n_pixels, float* rgb, float* depth, float* acc, int outputs_on_device) {
  if (!h || !c2w12 || width <= 0 || height <= 0 || !(focal > 0.f) || first_pixel < 0 || n_pixels <= 0 ||
      first_pixel + n_pixels > (long)width * height) { set_error("render_view: bad arguments"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  ProfActivate pa(h);
  NERF_TRY(h->mlp->prepare(h->params, h->st));
  const int last = h->NL - 1;
  const size_t R = (size_t)h->Rmax;
  float* b = h->rays_dev;
  const cudaMemcpyKind kind = outputs_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  for (long b0 = 0; b0 < n_pixels; b0 += h->Rc) {
    const int rc = (int)(n_pixels - b0 < h->Rc ? n_pixels - b0 : h->Rc);
    { ProfScope ps(PC_MISC, h->st);
      NERF_TRY(launch_generate_rays(c2w12, focal, width, height, near, far, edge_mode, first_pixel + b0, rc, b, b + 3 * R, b + 6 * R, b + 7 * R,
                                    b + 8 * R, h->st)); }
    set_batch_dev(h, b, b + 3 * R, b + 6 * R, b + 7 * R, b + 8 * R, nullptr, nullptr);
    const uint32_t keep = h->ray_offset;
    h->ray_offset = keep + (uint32_t)b0;
    const int s = forward_chunk(h, 0, rc, false);
    h->ray_offset = keep;
    NERF_TRY(s);
    auto& L = h->lv[last];
    if (rgb) NERF_CUDA(cudaMemcpyAsync(rgb + b0 * 3, L.comp_rgb, (size_t)rc * 12, kind, h->st));
    if (depth) NERF_CUDA(cudaMemcpyAsync(depth + b0, L.depth, (size_t)rc * 4, kind, h->st));
    if (acc) NERF_CUDA(cudaMemcpyAsync(acc + b0, L.acc, (size_t)rc * 4, kind, h->st));
  }
  if (!outputs_on_device) NERF_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

int nerf_mipnerf_level_outputs(nerf_mipnerf* h, int level, uint64_t* comp_rgb, uint64_t* depth, uint64_t* acc,
                               uint64_t* weights, uint64_t* t_vals) {
  if (!h || level < 0 || level >= h->NL) { set_error("level_outputs: bad level"); return NERF_ERR_INVALID; }
  auto& L = h->lv[level];
  if (comp_rgb) *comp_rgb = (uint64_t)(uintptr_t)L.comp_rgb;
  if (depth) *depth = (uint64_t)(uintptr_t)L.depth;
  if (acc) *acc = (uint64_t)(uintptr_t)L.acc;
  if (weights) *weights = (uint64_t)(uintptr_t)L.weights;
  if (t_vals) *t_vals = (uint64_t)(uintptr_t)L.t;
  return 0;
}

int nerf_mipnerf_get_loss(nerf_mipnerf* h, float* loss_per_level, float* total) {
  if (!h) { set_error("null handle"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  return read_loss(h, loss_per_level, total);
}
int nerf_mipnerf_synchronize(nerf_mipnerf* h) {
  if (!h) { set_error("null handle"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  NERF_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}
int nerf_mipnerf_launch_count(nerf_mipnerf* h, long* n) {
  (void)h;
  if (n) *n = launch_count();
  return 0;
}

int nerf_mipnerf_stream(nerf_mipnerf* h, uint64_t* stream) {
  if (!h || !stream) { set_error("null argument"); return NERF_ERR_INVALID; }
  *stream = (uint64_t)(uintptr_t)h->st;
  return 0;
}
int nerf_mipnerf_set_profiling(nerf_mipnerf* h, int on) {
  if (!h) { set_error("null handle"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  NERF_CUDA(cudaStreamSynchronize(h->st));
  if (on && !h->prof) h->prof = new Profiler();
  if (!on && h->prof) { delete h->prof; h->prof = nullptr; }
  if (h->prof) h->prof->reset();
  return 0;
}
int nerf_mipnerf_read_profile(nerf_mipnerf* h, int max_cat, int* n_cat, const char** names, double* ms, long* launches, int reset) {
  if (!h || !h->prof) { set_error("profiling is off"); return NERF_ERR_STATE; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  h->prof->collect();
  const int n = max_cat < PC_COUNT ? max_cat : PC_COUNT;
  for (int i = 0; i < n; i++) {
    if (names) names[i] = prof_name(i);
    if (ms) ms[i] = h->prof->ms[i];
    if (launches) launches[i] = h->prof->launches[i];
  }
  if (n_cat) *n_cat = n;
  if (reset) h->prof->reset();
  return 0;
}

// ---- AcceleratedMLP -------------------------------------------------------------------------------
int nerf_mlp_get_output(nerf_mipnerf* h, const float* enc_pos, const float* enc_dir, int level, int n_rays,
                        uint64_t* density_dev, uint64_t* rgb_dev) {
  if (!h || !enc_pos || !enc_dir || level < 0 || level >= h->NL || n_rays <= 0 || n_rays > h->Rc) { set_error("mlp_get_output: bad arguments (n_rays must be <= chunk_rays=%d)", h ? h->Rc : 0); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  auto& L = h->lv[level];
  const long M = (long)n_rays * h->S;
  if (!L.density_out) { NERF_TRY(dalloc(h, &L.density_out, (size_t)h->Rc * h->S)); NERF_TRY(dalloc(h, &L.rgb_out, (size_t)h->Rc * h->S * 3)); }
  NERF_TRY(h->mlp->prepare(h->params, h->st));
  NERF_TRY(h->mlp->import_encodings(level, enc_pos, enc_dir, M, h->st));
  NERF_TRY(h->mlp->forward(level, M, h->params, L.raw_density, L.raw_rgb, h->st));
  NERF_TRY(launch_output_activations(L.raw_density, L.raw_rgb, M, model_act(h), L.density_out, L.rgb_out, h->st));
  L.last_rows = M;
  NERF_CUDA(cudaStreamSynchronize(h->st));
  if (density_dev) *density_dev = (uint64_t)(uintptr_t)L.density_out;
  if (rgb_dev) *rgb_dev = (uint64_t)(uintptr_t)L.rgb_out;
  return 0;
}
int nerf_mlp_get_gradient(nerf_mipnerf* h, const float* color_grad, const float* density_grad, int level, float** grad_dev_ptrs) {
  if (!h || !color_grad || !density_grad || level < 0 || level >= h->NL) { set_error("mlp_get_gradient: bad arguments"); return NERF_ERR_INVALID; }
  auto& L = h->lv[level];
  if (L.last_rows <= 0) { set_error("mlp_get_gradient(level %d) before get_output", level); return NERF_ERR_STATE; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  NERF_TRY(launch_output_activations_grad(L.raw_density, L.raw_rgb, density_grad, color_grad, L.last_rows, model_act(h),
                                          L.d_raw_density, L.d_raw_rgb, h->st));
  NERF_TRY(h->mlp->backward(level, L.last_rows, h->params, h->grads, L.d_raw_density, L.d_raw_rgb, h->st));
  NERF_CUDA(cudaStreamSynchronize(h->st));
  if (grad_dev_ptrs) memcpy(grad_dev_ptrs, h->grad_ptrs.data(), h->grad_ptrs.size() * sizeof(float*));
  return 0;
}
int nerf_mlp_relu_bits(nerf_mipnerf* h, int level, int layer, uint64_t* bits_dev, int* words_per_row) {
  if (!h || !bits_dev || !words_per_row || level < 0 || level >= h->NL) { set_error("mlp_relu_bits: bad arguments"); return NERF_ERR_INVALID; }
  const uint32_t* b = nullptr;
  NERF_TRY(h->mlp->relu_bits(level, layer, &b, words_per_row));
  *bits_dev = (uint64_t)(uintptr_t)b;
  return 0;
}
int nerf_mlp_reset_gradients(nerf_mipnerf* h, int level) {
  (void)level;  // one gradient buffer summed over levels (A-D4/D5)
  if (!h) { set_error("null handle"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  NERF_CUDA(cudaMemsetAsync(h->grads, 0, (size_t)h->shape.n_params * sizeof(float), h->st));
  return 0;
}

// ---- AcceleratedAdamOptimizer ----------------------------------------------------------------------
int nerf_adam_create(const int* sizes, int n, int eps_mode, int device, nerf_adam** out) {
  if (!sizes || n <= 0 || !out) { set_error("adam_create: bad arguments"); return NERF_ERR_INVALID; }
  *out = nullptr;
  NERF_TRY(check_device(device));
  nerf_adam* a = new nerf_adam();
  a->device = device; a->eps_mode = eps_mode;
  a->sizes.assign(sizes, sizes + n);
  for (int i = 0; i < n; i++) a->n += sizes[i];
  if (cudaMalloc(&a->m, a->n * sizeof(float)) != cudaSuccess || cudaMalloc(&a->v, a->n * sizeof(float)) != cudaSuccess) {
    set_error("adam_create: cudaMalloc failed"); delete a; return NERF_ERR_NO_DEVICE;
  }
  cudaMemset(a->m, 0, a->n * sizeof(float));  // A-D14: the reference never zeroes m, v
  cudaMemset(a->v, 0, a->n * sizeof(float));
  *out = a;
  return 0;
}
int nerf_adam_step(nerf_adam* a, float** params, float** grads, float lr) {
  if (!a || !params || !grads) { set_error("adam_step: null argument"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(a->device));
  a->iteration++;  // pre-incremented (ANU/AcceleratedAdamOptimizer.cpp:26)
  const int n = (int)a->sizes.size();
  long off = 0;
  int i = 0;
  while (i < n) {  // coalesce contiguous runs: a single launch when the tables view one flat buffer
    int j = i;
    long len = a->sizes[i];
    while (j + 1 < n && params[j + 1] == params[j] + a->sizes[j] && grads[j + 1] == grads[j] + a->sizes[j]) { j++; len += a->sizes[j]; }
    NERF_TRY(adam_step_flat(a, params[i], grads[i], len, off, lr, 1.0f, 0));
    off += len;
    i = j + 1;
  }
  return 0;
}
int nerf_adam_state(nerf_adam* a, float** m, float** v, long* n, int* iteration) {
  if (!a) { set_error("null handle"); return NERF_ERR_INVALID; }
  if (m) *m = a->m;
  if (v) *v = a->v;
  if (n) *n = a->n;
  if (iteration) *iteration = a->iteration;
  return 0;
}
int nerf_adam_set_state(nerf_adam* a, const float* m, const float* v, long n, int iteration) {
  if (!a || n != a->n) { set_error("adam_set_state: expected %ld floats", a ? a->n : 0L); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(a->device));
  NERF_CUDA(cudaDeviceSynchronize());
  if (m) NERF_CUDA(cudaMemcpy(a->m, m, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
  if (v) NERF_CUDA(cudaMemcpy(a->v, v, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
  a->iteration = iteration;
  return 0;
}
int nerf_adam_destroy(nerf_adam* a) {
  if (!a) return 0;
  cudaSetDevice(a->device);
  cudaFree(a->m); cudaFree(a->v);
  delete a;
  return 0;
}

// ---- AcceleratedGradientCalculator -----------------------------------------------------------------
int nerf_gradcalc_create(int batch, int n_levels, float coarse_loss_mult, int device, nerf_gradcalc** out) {
  if (batch <= 0 || n_levels <= 0 || !out) { set_error("gradcalc_create: bad arguments"); return NERF_ERR_INVALID; }
  *out = nullptr;
  NERF_TRY(check_device(device));
  nerf_gradcalc* g = new nerf_gradcalc();
  g->device = device; g->batch = batch; g->n_levels = n_levels; g->coarse_mult = coarse_loss_mult;
  if (cudaMalloc(&g->grad, (size_t)n_levels * batch * 12) != cudaSuccess || cudaMalloc(&g->pixels, (size_t)batch * 12) != cudaSuccess) {
    set_error("gradcalc_create: cudaMalloc failed"); delete g; return NERF_ERR_NO_DEVICE;
  }
  *out = g;
  return 0;
}
int nerf_gradcalc_get_output_gradient(nerf_gradcalc* g, uint64_t comp_rgb_dev, const float* pixels3, int n,
                                      uint64_t loss_mults_dev, float loss_mult_sum, int level, uint64_t* grad_dev) {
  if (!g || !pixels3 || n <= 0 || n > g->batch || level < 0 || level >= g->n_levels) { set_error("gradcalc: bad arguments"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(g->device));
  NERF_CUDA(cudaMemcpy(g->pixels, pixels3, (size_t)n * 12, cudaMemcpyHostToDevice));  // dst/src as intended (A-D13)
  float* out = g->grad + (size_t)level * g->batch * 3;
  const float mult = level < g->n_levels - 1 ? g->coarse_mult : 1.0f;
  NERF_TRY(launch_output_gradient((const float*)(uintptr_t)comp_rgb_dev, g->pixels, (const float*)(uintptr_t)loss_mults_dev, n,
                                  loss_mult_sum, nullptr, mult, out, nullptr, nullptr, 0));
  NERF_CUDA(cudaDeviceSynchronize());
  if (grad_dev) *grad_dev = (uint64_t)(uintptr_t)out;
  return 0;
}
int nerf_gradcalc_destroy(nerf_gradcalc* g) {
  if (!g) return 0;
  cudaSetDevice(g->device);
  cudaFree(g->grad); cudaFree(g->pixels);
  delete g;
  return 0;
}

// ---- OutputRetriever -------------------------------------------------------------------------------
int nerf_retrieve_output(uint64_t dev, int n_float3, float* host_out) {
  if (!dev || !host_out || n_float3 <= 0) { set_error("retrieve_output: bad arguments"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaMemcpy(host_out, (const void*)(uintptr_t)dev, (size_t)n_float3 * 12, cudaMemcpyDeviceToHost));
  return 0;
}

// ---- fused training step ---------------------------------------------------------------------------
static int finish_step(nerf_mipnerf* h, nerf_adam* a, float lr, float* loss_out) {
  if (a->n != h->shape.n_params) { set_error("optimizer size %ld != model parameters %ld", a->n, h->shape.n_params); return NERF_ERR_INVALID; }
  ProfActivate pa(h);
  // the step's ONE collective: [gradient | sum(loss_mults) | per-level loss numerators], fp32 sum
  if (h->comm) { ProfScope ps(PC_COMM, h->st); NERF_NCCL(g_nccl.AllReduce(h->grads, h->grads, (size_t)h->shape.n_params + 1 + h->NL, kNcclFloat32, kNcclSum, h->comm, h->st)); }
  a->iteration++;
  { ProfScope ps(PC_ADAM, h->st);
    if (h->grads_unnormalised) {  // the global 1 / sum(lm) rides in the Adam pass, which writes the normalised gradient back
      const float inv1 = 1.0f / (1.0f - powf(nerf_adam::beta1, (float)a->iteration));
      const float inv2 = 1.0f / (1.0f - powf(nerf_adam::beta2, (float)a->iteration));
      NERF_TRY(launch_adam_dp(h->params, h->grads, a->m, a->v, a->n, lr, nerf_adam::beta1, nerf_adam::beta2, inv1, inv2, a->eps_mode,
                              h->scalars, h->st));
      h->grads_unnormalised = false;
    } else {
      NERF_TRY(adam_step_flat(a, h->params, h->grads, a->n, 0, lr, 1.0f, h->st));
    } }
  if (loss_out) {
    float per[8];
    NERF_TRY(read_loss(h, per, loss_out));
  }
  return 0;
}
int nerf_mipnerf_train_step(nerf_mipnerf* h, nerf_adam* a, const float* o, const float* d, const float* radii,
                            const float* nears, const float* fars, const float* lm, const float* pixels, int n_rays,
                            float lr, float* loss_out) {
  if (!h || !a || !pixels) { set_error("train_step: null argument"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  NERF_TRY(upload_batch(h, o, d, radii, nears, fars, lm, pixels, n_rays));
  NERF_TRY(gradient_core(h, n_rays, nullptr, nullptr));
  return finish_step(h, a, lr, loss_out);
}
int nerf_mipnerf_train_step_dev(nerf_mipnerf* h, nerf_adam* a, const float* o, const float* d, const float* radii,
                                const float* nears, const float* fars, const float* lm, const float* pixels, int n_rays,
                                float lr, float* loss_out) {
  if (!h || !a || !o || !d || !radii || !nears || !fars || !lm || !pixels) { set_error("train_step_dev: null argument"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  set_batch_dev(h, o, d, radii, nears, fars, lm, pixels);
  NERF_TRY(gradient_core(h, n_rays, nullptr, nullptr));
  return finish_step(h, a, lr, loss_out);
}

// ---- resident dataset + on-device batch assembly (replaces SN/BinDataset.cs:27-52) -----------
struct nerf_dataset {
  int device = 0;
  long n = 0;
  float* rec = nullptr;  // [n, 16] floats in HBM
  long* idx = nullptr;   // scratch for explicit / exported indices
  long idx_cap = 0;
};

static int dataset_scratch(nerf_dataset* ds, long n) {
  if (n <= ds->idx_cap) return 0;
  if (ds->idx) cudaFree(ds->idx);
  ds->idx = nullptr; ds->idx_cap = 0;
  NERF_CUDA(cudaMalloc(&ds->idx, (size_t)n * sizeof(long)));
  ds->idx_cap = n;
  return 0;
}

int nerf_dataset_create(const void* records, long n_records, int device, nerf_dataset** out) {
  if (!records || n_records <= 0 || n_records > 0xFFFFFFFFL || !out) { set_error("dataset_create: bad arguments"); return NERF_ERR_INVALID; }
  *out = nullptr;
  NERF_TRY(check_device(device));
  nerf_dataset* ds = new nerf_dataset();
  ds->device = device; ds->n = n_records;
  if (cudaMalloc(&ds->rec, (size_t)n_records * NERF_RECORD_BYTES) != cudaSuccess) {
    set_error("dataset_create: cudaMalloc of %ld records failed", n_records); delete ds; return NERF_ERR_NO_DEVICE;
  }
  if (cudaMemcpy(ds->rec, records, (size_t)n_records * NERF_RECORD_BYTES, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("dataset_create: upload failed"); cudaFree(ds->rec); delete ds; return NERF_ERR_NO_DEVICE;
  }
  *out = ds;
  return 0;
}

int nerf_dataset_load(const char* path, int device, nerf_dataset** out) {  // train_data.bin of SN/Program.cs:23
  if (!path || !out) { set_error("dataset_load: null argument"); return NERF_ERR_INVALID; }
  FILE* f = fopen(path, "rb");
  if (!f) { set_error("dataset_load: cannot open %s", path); return NERF_ERR_INVALID; }
  fseek(f, 0, SEEK_END);
  const long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  const long n = bytes / NERF_RECORD_BYTES;  // SN/BinDataset.cs:15
  if (n <= 0) { fclose(f); set_error("dataset_load: %s holds no 64-byte record", path); return NERF_ERR_INVALID; }
  std::vector<char> buf((size_t)n * NERF_RECORD_BYTES);
  const size_t got = fread(buf.data(), NERF_RECORD_BYTES, (size_t)n, f);
  fclose(f);
  if ((long)got != n) { set_error("dataset_load: short read from %s", path); return NERF_ERR_INVALID; }  // SN/BinDataset.cs:38-39
  return nerf_dataset_create(buf.data(), n, device, out);
}

// Dataset.GenerateRays for one camera into caller device arrays (parity hook of nerf_mipnerf_render_view's ray source)
int nerf_generate_rays(const float* c2w12, float focal, int width, int height, float near, float far, int edge_mode, long first_pixel,
                       long n_pixels, float* origins3_dev, float* directions3_dev, float* radii_dev, float* nears_dev, float* fars_dev) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); set_error("no CUDA device available: libnerfb200 has no CPU path"); return NERF_ERR_NO_DEVICE; }
  if (!c2w12 || width <= 0 || height <= 0 || !(focal > 0.f) || first_pixel < 0 || n_pixels <= 0 || first_pixel + n_pixels > (long)width * height ||
      !origins3_dev || !directions3_dev || !radii_dev || !nears_dev || !fars_dev) { set_error("generate_rays: bad arguments"); return NERF_ERR_INVALID; }
  NERF_TRY(launch_generate_rays(c2w12, focal, width, height, near, far, edge_mode, first_pixel, n_pixels, origins3_dev, directions3_dev, radii_dev,
                                nears_dev, fars_dev, 0));
  NERF_CUDA(cudaStreamSynchronize(0));
  return 0;
}

int nerf_dataset_size(const nerf_dataset* ds, long* n) {
  if (!ds || !n) { set_error("null argument"); return NERF_ERR_INVALID; }
  *n = ds->n;
  return 0;
}

int nerf_dataset_destroy(nerf_dataset* ds) {
  if (!ds) return 0;
  cudaSetDevice(ds->device);
  cudaFree(ds->rec); cudaFree(ds->idx);
  delete ds;
  return 0;
}

// the record indices a step with (seed, step, first_slot) draws — what nerf_mipnerf_train_step_dataset uses
int nerf_dataset_draw_indices(nerf_dataset* ds, uint64_t seed, uint32_t step, uint32_t first_slot, int n_rays, int64_t* idx_host) {
  if (!ds || !idx_host || n_rays <= 0) { set_error("draw_indices: bad arguments"); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(ds->device));
  NERF_TRY(dataset_scratch(ds, n_rays));
  NERF_TRY(launch_draw_indices(seed, first_slot, step, ds->n, n_rays, ds->idx, nullptr));
  NERF_CUDA(cudaMemcpy(idx_host, ds->idx, (size_t)n_rays * sizeof(long), cudaMemcpyDeviceToHost));
  return 0;
}

// gather records idx_host[0..n) (or, if NULL, the Philox draw of (seed, step, first_slot)) into caller device arrays
int nerf_dataset_gather(nerf_dataset* ds, const int64_t* idx_host, uint64_t seed, uint32_t step, uint32_t first_slot, int n_rays,
                        float* origins3_dev, float* directions3_dev, float* radii_dev, float* nears_dev, float* fars_dev,
                        float* loss_mults_dev, float* pixels3_dev) {
  if (!ds || n_rays <= 0 || !origins3_dev || !directions3_dev || !radii_dev || !nears_dev || !fars_dev || !loss_mults_dev || !pixels3_dev) {
    set_error("dataset_gather: bad arguments"); return NERF_ERR_INVALID;
  }
  NERF_CUDA(cudaSetDevice(ds->device));
  const long* idx = nullptr;
  if (idx_host) {
    NERF_TRY(dataset_scratch(ds, n_rays));
    NERF_CUDA(cudaMemcpy(ds->idx, idx_host, (size_t)n_rays * sizeof(long), cudaMemcpyHostToDevice));
    idx = ds->idx;
  }
  NERF_TRY(launch_gather_batch(ds->rec, ds->n, idx, seed, first_slot, step, n_rays, origins3_dev, directions3_dev, radii_dev, nears_dev,
                               fars_dev, loss_mults_dev, pixels3_dev, nullptr));
  NERF_CUDA(cudaDeviceSynchronize());
  return 0;
}

// one training step with the batch drawn and assembled on the device: no host->device traffic
int nerf_mipnerf_train_step_dataset(nerf_mipnerf* h, nerf_adam* a, nerf_dataset* ds, int n_rays, uint64_t sampler_seed, float lr,
                                    float* loss_out) {
  if (!h || !a || !ds) { set_error("train_step_dataset: null argument"); return NERF_ERR_INVALID; }
  if (ds->device != h->cfg.device) { set_error("train_step_dataset: dataset on device %d, model on %d", ds->device, h->cfg.device); return NERF_ERR_INVALID; }
  if (n_rays <= 0 || n_rays > h->Rmax) { set_error("n_rays=%d outside (0, %d]", n_rays, h->Rmax); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  const size_t R = (size_t)h->Rmax;
  float* b = h->rays_dev;
  {
    ProfActivate pa(h);
    ProfScope ps(PC_MISC, h->st);
    NERF_TRY(launch_gather_batch(ds->rec, ds->n, nullptr, sampler_seed, h->ray_offset, h->step, n_rays, b, b + 3 * R, b + 6 * R, b + 7 * R,
                                 b + 8 * R, b + 9 * R, b + 10 * R, h->st));
  }
  set_batch_dev(h, b, b + 3 * R, b + 6 * R, b + 7 * R, b + 8 * R, b + 9 * R, b + 10 * R);
  NERF_TRY(gradient_core(h, n_rays, nullptr, nullptr));
  return finish_step(h, a, lr, loss_out);
}

#define STAGE_BEGIN_() do { int n_ = 0; if (cudaGetDeviceCount(&n_) != cudaSuccess || n_ <= 0) { cudaGetLastError(); set_error("no CUDA device available: libnerfb200 has no CPU path"); return NERF_ERR_NO_DEVICE; } } while (0)
// ---- image metrics ----------------------------------------------------------------------------
// mse over n floats and psnr = -10/ln(10) * ln(mse) (SN/MipHelpers.cs:672); on_device: a/b are device pointers
int nerf_image_error(const float* a, const float* b, long n, int on_device, double* mse, double* psnr) {
  if (!a || !b || n <= 0) { set_error("image_error: bad arguments"); return NERF_ERR_INVALID; }
  struct DevBuf {  // freed on every exit path
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
  } da, db, dout;
  NERF_CUDA(cudaMalloc(&dout.p, sizeof(double)));
  if (!on_device) {
    NERF_CUDA(cudaMalloc(&da.p, (size_t)n * 4));
    NERF_CUDA(cudaMalloc(&db.p, (size_t)n * 4));
    NERF_CUDA(cudaMemcpy(da.p, a, (size_t)n * 4, cudaMemcpyHostToDevice));
    NERF_CUDA(cudaMemcpy(db.p, b, (size_t)n * 4, cudaMemcpyHostToDevice));
  }
  NERF_TRY(launch_sq_err(on_device ? a : (const float*)da.p, on_device ? b : (const float*)db.p, n, (double*)dout.p, nullptr));
  double sq = 0.0;
  NERF_CUDA(cudaMemcpy(&sq, dout.p, sizeof(double), cudaMemcpyDeviceToHost));
  const double m = sq / (double)n;
  if (mse) *mse = m;
  if (psnr) *psnr = -10.0 / log(10.0) * log(m);
  return 0;
}

// ComputeSsimAverage / ComputeSsim (SN/MipHelpers.cs:688-737): images [height, width, 3]; ssim_map (optional) gets the
// per-pixel, per-channel map.  on_device: a / b / ssim_map are device pointers.
int nerf_image_ssim(const float* a, const float* b, int width, int height, float max_val, int filter_size, float filter_sigma,
                    float k1, float k2, int on_device, double* ssim_mean, float* ssim_map) {
  if (!a || !b || width <= 0 || height <= 0) { set_error("image_ssim: bad arguments"); return NERF_ERR_INVALID; }
  STAGE_BEGIN_();
  struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
  } da, db, dmap, dsum;
  const size_t n = (size_t)width * height * 3;
  const long nb = ssim_blocks(width, height);
  NERF_CUDA(cudaMalloc(&dsum.p, (size_t)nb * sizeof(double)));
  if (!on_device) {
    NERF_CUDA(cudaMalloc(&da.p, n * 4));
    NERF_CUDA(cudaMalloc(&db.p, n * 4));
    NERF_CUDA(cudaMemcpy(da.p, a, n * 4, cudaMemcpyHostToDevice));
    NERF_CUDA(cudaMemcpy(db.p, b, n * 4, cudaMemcpyHostToDevice));
    if (ssim_map) NERF_CUDA(cudaMalloc(&dmap.p, n * 4));
  }
  float* map_dev = on_device ? ssim_map : (float*)dmap.p;
  NERF_TRY(launch_ssim(on_device ? a : (const float*)da.p, on_device ? b : (const float*)db.p, width, height, max_val, filter_size,
                       filter_sigma, k1, k2, map_dev, (double*)dsum.p, nullptr));
  std::vector<double> sums;
  try { sums.resize((size_t)nb); } catch (const std::exception& e) { set_error("image_ssim: %s", e.what()); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaMemcpy(sums.data(), dsum.p, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost));
  if (!on_device && ssim_map) NERF_CUDA(cudaMemcpy(ssim_map, dmap.p, n * 4, cudaMemcpyDeviceToHost));
  double t = 0.0;
  for (long i = 0; i < nb; i++) t += sums[(size_t)i];  // block order: bitwise reproducible
  if (ssim_mean) *ssim_mean = t / (double)n;
  return 0;
}

// ---- learning-rate schedule and checkpoints ---------------------------------------------------
// SN/MipHelpers.cs:758-773, float32 arithmetic like the original
float nerf_learning_rate_decay(int step, float lr_init, float lr_final, int max_steps, int lr_delay_steps, float lr_delay_mult) {
  float delay_rate = 1.0f;
  if (lr_delay_steps > 0) {
    float p = (float)step / (float)lr_delay_steps;
    p = p < 0.f ? 0.f : (p > 1.f ? 1.f : p);
    delay_rate = lr_delay_mult + (1.0f - lr_delay_mult) * sinf(0.5f * 3.14159274f * p);
  }
  float t = (float)step / (float)max_steps;
  t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
  return delay_rate * expf(logf(lr_init) * (1.0f - t) + logf(lr_final) * t);
}

namespace {
struct CkptHeader {
  char magic[8];        // "NERFB200"
  uint32_t version;     // 2 (1: the same without seed / precision — still readable)
  uint32_t n_tensors;
  int64_t n_params;
  int32_t adam_iteration;  // -1: no optimizer state stored
  uint32_t step;           // Philox step counter of the model
  int32_t shape[7];        // depth, width, depth_cond, width_cond, skip, deg_point, deg_view
};
struct CkptHeaderV2 {  // follows CkptHeader when version >= 2: what a bitwise-identical resume also depends on
  uint64_t seed;       // sampling RNG key (cfg.seed)
  int32_t precision;   // NERF_PRECISION_* the state was trained in
  int32_t reserved;
};
}  // namespace

// parameters (flat W0..W10,b0..b10), Adam m/v/iteration and the sampling step counter -> one file
int nerf_checkpoint_save(nerf_mipnerf* h, nerf_adam* a, const char* path) {
  if (!h || !path) { set_error("checkpoint_save: null argument"); return NERF_ERR_INVALID; }
  if (a && a->n != h->shape.n_params) { set_error("checkpoint_save: optimizer size %ld != model parameters %ld", a->n, h->shape.n_params); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  NERF_CUDA(cudaStreamSynchronize(h->st));
  const long n = h->shape.n_params;
  std::vector<float> buf;
  try { buf.resize((size_t)n * (a ? 3 : 1)); } catch (const std::exception& e) { set_error("checkpoint_save: %s", e.what()); return NERF_ERR_INVALID; }
  NERF_CUDA(cudaMemcpy(buf.data(), h->params, (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (a) {
    NERF_CUDA(cudaMemcpy(buf.data() + n, a->m, (size_t)n * 4, cudaMemcpyDeviceToHost));
    NERF_CUDA(cudaMemcpy(buf.data() + 2 * n, a->v, (size_t)n * 4, cudaMemcpyDeviceToHost));
  }
  CkptHeader hd;
  memset(&hd, 0, sizeof(hd));
  memcpy(hd.magic, "NERFB200", 8);
  hd.version = 2; hd.n_tensors = (uint32_t)h->sizes.size(); hd.n_params = n;
  hd.adam_iteration = a ? a->iteration : -1; hd.step = h->step;
  const nerf_config& c = h->cfg;
  const int32_t shp[7] = {c.net_depth, c.net_width, c.net_depth_condition, c.net_width_condition, c.skip_layer, c.deg_point, c.deg_view};
  memcpy(hd.shape, shp, sizeof(shp));
  FILE* f = fopen(path, "wb");
  if (!f) { set_error("checkpoint_save: cannot open %s", path); return NERF_ERR_INVALID; }
  CkptHeaderV2 h2;
  memset(&h2, 0, sizeof(h2));
  h2.seed = c.seed; h2.precision = c.precision;
  bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1;
  ok = ok && fwrite(&h2, sizeof(h2), 1, f) == 1;
  ok = ok && fwrite(h->sizes.data(), sizeof(int), h->sizes.size(), f) == h->sizes.size();
  ok = ok && fwrite(buf.data(), sizeof(float), buf.size(), f) == buf.size();
  ok = (fclose(f) == 0) && ok;
  if (!ok) { set_error("checkpoint_save: write to %s failed", path); return NERF_ERR_INVALID; }
  return 0;
}

int nerf_checkpoint_load(nerf_mipnerf* h, nerf_adam* a, const char* path) {
  if (!h || !path) { set_error("checkpoint_load: null argument"); return NERF_ERR_INVALID; }
  struct File {  // closed on every exit path
    FILE* f = nullptr;
    ~File() { if (f) fclose(f); }
  } file;
  file.f = fopen(path, "rb");
  if (!file.f) { set_error("checkpoint_load: cannot open %s", path); return NERF_ERR_INVALID; }
  FILE* f = file.f;
  try {
    CkptHeader hd;
    if (fread(&hd, sizeof(hd), 1, f) != 1 || memcmp(hd.magic, "NERFB200", 8) != 0 || (hd.version != 1 && hd.version != 2)) {
      set_error("checkpoint_load: %s is not a version-1/2 checkpoint", path); return NERF_ERR_INVALID;
    }
    const nerf_config& c = h->cfg;
    CkptHeaderV2 h2;
    memset(&h2, 0, sizeof(h2));
    h2.seed = c.seed; h2.precision = c.precision;
    if (hd.version >= 2 && fread(&h2, sizeof(h2), 1, f) != 1) { set_error("checkpoint_load: %s is truncated", path); return NERF_ERR_INVALID; }
    const int32_t shp[7] = {c.net_depth, c.net_width, c.net_depth_condition, c.net_width_condition, c.skip_layer, c.deg_point, c.deg_view};
    // compare the header with the model BEFORE sizing anything from it: a corrupt or foreign file must not drive an allocation
    if (hd.n_params != h->shape.n_params || hd.n_tensors != h->sizes.size() || memcmp(hd.shape, shp, sizeof(shp)) != 0) {
      set_error("checkpoint_load: %s was written for a different network shape", path); return NERF_ERR_INVALID;
    }
    std::vector<int> sizes(h->sizes.size());
    if (fread(sizes.data(), sizeof(int), sizes.size(), f) != sizes.size() || sizes != h->sizes) {
      set_error("checkpoint_load: %s was written for a different network shape", path); return NERF_ERR_INVALID;
    }
    const long n = h->shape.n_params;
    const bool has_opt = hd.adam_iteration >= 0;
    if (a && !has_opt) { set_error("checkpoint_load: %s holds no optimizer state", path); return NERF_ERR_INVALID; }
    if (a && a->n != n) { set_error("checkpoint_load: optimizer size %ld != %ld", a->n, n); return NERF_ERR_INVALID; }
    // resuming TRAINING (optimizer state requested) is only bitwise identical with the same sampling seed and precision
    if (a && (h2.seed != c.seed || h2.precision != c.precision)) {
      set_error("checkpoint_load: %s was trained with seed %llu / precision %d, this model has seed %llu / precision %d "
                "(pass a NULL optimizer to load the parameters only)", path, (unsigned long long)h2.seed, h2.precision,
                (unsigned long long)c.seed, c.precision);
      return NERF_ERR_INVALID;
    }
    std::vector<float> buf((size_t)n * (has_opt ? 3 : 1));
    const size_t got = fread(buf.data(), sizeof(float), buf.size(), f);
    if (got != buf.size()) { set_error("checkpoint_load: %s is truncated", path); return NERF_ERR_INVALID; }
    NERF_CUDA(cudaSetDevice(h->cfg.device));
    NERF_CUDA(cudaStreamSynchronize(h->st));
    NERF_CUDA(cudaMemcpy(h->params, buf.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    h->step = hd.step;
    if (a) {
      NERF_CUDA(cudaMemcpy(a->m, buf.data() + n, (size_t)n * 4, cudaMemcpyHostToDevice));
      NERF_CUDA(cudaMemcpy(a->v, buf.data() + 2 * n, (size_t)n * 4, cudaMemcpyHostToDevice));
      a->iteration = hd.adam_iteration;
    }
  } catch (const std::exception& e) {  // nothing may unwind through the extern "C" boundary
    set_error("checkpoint_load: %s", e.what());
    return NERF_ERR_INVALID;
  }
  return 0;
}

// ---- multi-GPU -------------------------------------------------------------------------------------
int nerf_comm_get_unique_id(void* id_out) {
  if (!id_out) { set_error("null argument"); return NERF_ERR_INVALID; }
  NERF_TRY(nccl_load());
  NERF_NCCL(g_nccl.GetUniqueId(id_out));
  return 0;
}
int nerf_mipnerf_comm_init(nerf_mipnerf* h, const void* id, int rank, int world) {
  if (!h || !id || world < 1 || rank < 0 || rank >= world) { set_error("comm_init: bad arguments"); return NERF_ERR_INVALID; }
  NERF_TRY(nccl_load());
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  UidByValue uid;
  memcpy(uid.internal, id, NERF_COMM_ID_BYTES);
  NERF_NCCL(g_nccl.CommInitRank(&h->comm, world, uid, rank));
  h->rank = rank; h->world = world;
  h->ray_offset = (uint32_t)rank * (uint32_t)h->Rmax;  // distinct Philox counters per rank
  return 0;
}
int nerf_mipnerf_allreduce_gradients(nerf_mipnerf* h) {
  if (!h || !h->comm) { set_error("allreduce: no communicator attached"); return NERF_ERR_STATE; }
  NERF_CUDA(cudaSetDevice(h->cfg.device));
  NERF_NCCL(g_nccl.AllReduce(h->grads, h->grads, (size_t)h->shape.n_params + 1 + h->NL, kNcclFloat32, kNcclSum, h->comm, h->st));
  if (h->grads_unnormalised) {  // GetGradient + explicit allreduce + nerf_adam_step: normalise here
    NERF_TRY(launch_scale_by_inv(h->grads, h->shape.n_params, h->scalars, h->st));
    h->grads_unnormalised = false;
  }
  return 0;
}
int nerf_mipnerf_comm_destroy(nerf_mipnerf* h) {
  if (!h) return 0;
  if (h->comm && g_nccl.CommDestroy) { cudaSetDevice(h->cfg.device); cudaStreamSynchronize(h->st); g_nccl.CommDestroy(h->comm); }
  h->comm = nullptr; h->world = 1; h->rank = 0; h->ray_offset = 0;
  return 0;
}

// ---- per-stage entry points (default stream, synchronous) ------------------------------------------
#define STAGE_BEGIN() do { int n_ = 0; if (cudaGetDeviceCount(&n_) != cudaSuccess || n_ <= 0) { cudaGetLastError(); set_error("no CUDA device available: libnerfb200 has no CPU path"); return NERF_ERR_NO_DEVICE; } } while (0)
#define STAGE_END() do { NERF_CUDA(cudaStreamSynchronize(0)); return 0; } while (0)
namespace {
struct DevScratch {  // device allocation of a per-stage entry point, released on every exit path (cudaFree synchronises)
  void* p = nullptr;
  ~DevScratch() { if (p) cudaFree(p); }
};
}  // namespace

int nerf_get_sample_t_vals(const float* nears, const float* fars, const float* u, int R, int S, int randomized, float* t) {
  STAGE_BEGIN();
  SampleRng rng; rng.u = u;
  if (randomized && !u) { set_error("get_sample_t_vals: randomized needs explicit uniforms in the per-stage API"); return NERF_ERR_INVALID; }
  NERF_TRY(launch_sample_t_vals(nears, fars, rng, R, S, randomized, t, 0));
  STAGE_END();
}
int nerf_get_resampled_t_vals(const float* t, const float* w, const float* u, int R, int S, float padding, int randomized, float* t_new) {
  STAGE_BEGIN();
  SampleRng rng; rng.u = u;
  if (randomized && !u) { set_error("get_resampled_t_vals: randomized needs explicit uniforms in the per-stage API"); return NERF_ERR_INVALID; }
  NERF_TRY(launch_resample_t_vals(t, w, rng, R, S, padding, randomized, t_new, 0));
  STAGE_END();
}
int nerf_cast_rays(const float* t, const float* o, const float* d, float* means, float* covs, const float* radii, int R, int S) {
  STAGE_BEGIN();
  NERF_TRY(launch_cast_rays(t, o, d, radii, R, S, means, covs, 0));
  STAGE_END();
}
int nerf_encode_input_data(const float* means, const float* covs, const float* dirs, float* enc_pos, float* enc_dir, int R, int S, int deg_point, int deg_view) {
  STAGE_BEGIN();
  NERF_TRY(launch_encode_input_data(means, covs, dirs, enc_pos, enc_dir, R, S, deg_point, deg_view, 0));
  STAGE_END();
}
int nerf_apply_layer(const float* in_a, const float* in_b, const float* W, const float* b, float* out, float* z, long M, int n, int k_a, int k_b, int act) {
  STAGE_BEGIN();
  if (act < 0 || act > 3 || M <= 0 || n <= 0 || k_a <= 0) { set_error("apply_layer: bad arguments"); return NERF_ERR_INVALID; }
  if (n > 4) {
    NERF_TRY(launch_dense_fwd(in_a, k_a, k_a, in_b, k_b, in_b ? k_b : 0, W, b, out, z, M, n, (Act)act, 0));
  } else {
    if (in_b && k_b > 0) { set_error("apply_layer: conjoined inputs need n > 4"); return NERF_ERR_INVALID; }
    DevScratch own;  // freed on every exit path
    float* zz = z;
    if (!zz) { NERF_CUDA(cudaMalloc(&own.p, (size_t)M * n * sizeof(float))); zz = (float*)own.p; }
    NERF_TRY(launch_thin_fwd(in_a, k_a, W, b, zz, M, n, k_a, 0));
    if (out) NERF_TRY(launch_apply_act(zz, out, M * n, (Act)act, 0));
    NERF_CUDA(cudaStreamSynchronize(0));
  }
  STAGE_END();
}
int nerf_backpropagate_layer(const float* in_a, const float* in_b, const float* W, const float* Z, const float* dY, float* in_a_grads,
                             float* dW, float* db, long M, int n, int k_a, int k_b, int act) {
  STAGE_BEGIN();
  if (act < 0 || act > 3 || M <= 0 || n <= 0 || k_a <= 0) { set_error("backpropagate_layer: bad arguments"); return NERF_ERR_INVALID; }
  const int kb = in_b ? k_b : 0;
  if (n <= 4 && kb) { set_error("backpropagate_layer: conjoined inputs need n > 4"); return NERF_ERR_INVALID; }  // before any allocation
  DevScratch dz_own, ws_own;  // freed on every exit path
  NERF_CUDA(cudaMalloc(&dz_own.p, (size_t)M * n * sizeof(float)));
  float* dZ = (float*)dz_own.p;
  float* ws = nullptr;
  NERF_TRY(launch_act_grad(dY, Z, dZ, M * n, (Act)act, 0));  // dZ = dY * act'(Z)  (.cu:97-99)
  if (n > 4) {
    size_t wsz = dense_wgrad_workspace(M, n, k_a);
    if (kb) { const size_t w2 = dense_wgrad_workspace(M, n, kb); wsz = w2 > wsz ? w2 : wsz; }
    NERF_CUDA(cudaMalloc(&ws_own.p, wsz * sizeof(float)));
    ws = (float*)ws_own.p;
    NERF_TRY(launch_dense_wgrad(dZ, in_a, k_a, k_a, in_b, kb, kb, dW, db, M, n, ws, 0));
    if (in_a_grads) NERF_TRY(launch_dense_dgrad(dZ, W, k_a + kb, in_a_grads, M, n, k_a, nullptr, nullptr, nullptr, true, 0));
  } else {
    NERF_CUDA(cudaMalloc(&ws_own.p, thin_wgrad_workspace(M, n, k_a) * sizeof(float)));
    ws = (float*)ws_own.p;
    NERF_TRY(launch_thin_wgrad(dZ, in_a, k_a, dW, db, M, n, k_a, ws, 0));
    if (in_a_grads) NERF_TRY(launch_thin_dgrad(dZ, W, in_a_grads, M, n, k_a, nullptr, true, 0));
  }
  STAGE_END();  // synchronises the stream; the scratch buffers are released when they go out of scope after it
}
int nerf_volumetric_rendering(const float* rgb, const float* density, const float* t, const float* dirs, float* comp, float* depth,
                              float* acc, float* weights, int R, int S, int white) {
  STAGE_BEGIN();
  NERF_TRY(launch_composite_fwd(rgb, density, t, dirs, R, S, white, OutputAct{}, comp, depth, acc, weights, 0));
  STAGE_END();
}
int nerf_get_output_gradient(const float* comp, const float* pix, const float* lm, float* g, float lm_sum, float level_mult, int R) {
  STAGE_BEGIN();
  NERF_TRY(launch_output_gradient(comp, pix, lm, R, lm_sum, nullptr, level_mult, g, nullptr, nullptr, 0));
  STAGE_END();
}
int nerf_volumetric_rendering_gradient(const float* g, const float* rgb, const float* density, const float* t, const float* dirs,
                                       float* d_rgb, float* d_density, int R, int S, int white, int last_sample_mode) {
  STAGE_BEGIN();
  NERF_TRY(launch_composite_bwd(g, rgb, density, t, dirs, R, S, white, last_sample_mode, OutputAct{}, d_rgb, d_density, 0));
  STAGE_END();
}
int nerf_volumetric_rendering_async(const float* rgb, const float* density, const float* t, const float* dirs, float* comp, float* depth,
                                    float* acc, float* weights, int R, int S, int white, int raw, float density_bias, float rgb_padding,
                                    void* stream) {
  STAGE_BEGIN();
  OutputAct act; act.raw = raw != 0; act.density_bias = density_bias; act.rgb_padding = rgb_padding;
  NERF_TRY(launch_composite_fwd(rgb, density, t, dirs, R, S, white, act, comp, depth, acc, weights, (cudaStream_t)stream));
  return 0;
}
int nerf_volumetric_rendering_gradient_async(const float* g, const float* rgb, const float* density, const float* t, const float* dirs,
                                             float* d_rgb, float* d_density, int R, int S, int white, int last_sample_mode, int raw,
                                             float density_bias, float rgb_padding, void* stream) {
  STAGE_BEGIN();
  OutputAct act; act.raw = raw != 0; act.density_bias = density_bias; act.rgb_padding = rgb_padding;
  NERF_TRY(launch_composite_bwd(g, rgb, density, t, dirs, R, S, white, last_sample_mode, act, d_rgb, d_density, (cudaStream_t)stream));
  return 0;
}
int nerf_adam_optimizer_step(float* p, const float* g, float* m, float* v, float lr, float b1, float b2, float inv1, float inv2,
                             long size, int eps_mode) {
  STAGE_BEGIN();
  NERF_TRY(launch_adam(p, g, m, v, size, lr, b1, b2, inv1, inv2, eps_mode, 1.0f, 0));
  STAGE_END();
}

}  // extern "C"

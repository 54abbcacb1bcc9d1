// compositing.cu — volumetric ray integration forward/backward and the MSE output gradient.
//
// Replaces volumetric_rendering (.cu:318-344), volumetric_rendering_gradient (.cu:362-402) and
// get_output_gradient (.cu:347-361); semantics: SURVEY Appendix B.4/B.5.
//
// HBM-bound.  The reference walks each ray serially in ONE 1024-thread block with lane stride = S floats
// (fully uncoalesced) and round-trips alpha/T/w caches.  Here a ray is owned by LPR = S/8 consecutive lanes of a
// warp (32/LPR rays per warp), each lane holding 8 consecutive samples: sigma / rgb arrive as 128-bit loads (2 + 6
// per lane, the lanes of a ray covering its rows contiguously), w / dsigma / drgb leave as 128-bit stores, the
// transmittance is a shuffle-based segmented exclusive prefix PRODUCT, the backward a segmented suffix SUM
// (dL/dsigma_i = delta_i |d| (T_{i+1} dLdw_i - sum_{j>i} dLdw_j w_j)), alpha/T are recomputed instead of cached,
// and the output activations (softplus / sigmoid, SN/MipNerfModel.cs:81-83) and their derivatives are fused.
// Algorithmic bytes: fwd 24 B/sample + 32 B/ray, bwd 36 B/sample + 24 B/ray.
//
// Instruction budget: at 70 % of the measured copy bandwidth the forward pass has ~130 issue slots per sample
// (148 SMs x 4 schedulers x 32 lanes against 190 G samples/s); the first version spent 156 (one ray per warp with 4
// samples per lane: ~340 per-thread instructions of scan / reduction / addressing overhead amortised over 4 samples,
// denormal-safe MUFU wrappers, a shared-memory detour for t).  8 samples per lane, .ftz MUFU forms and direct t loads
// bring it to ~60.
#include "kernels.cuh"

namespace nerf {
namespace {

constexpr int kThreads = 256;  // 8 warps per block
constexpr int kV = 8;          // samples per lane
constexpr unsigned kFull = 0xffffffffu;
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

// These kernels must stay HBM-bound: with libm expf/log1pf/IEEE division the ~9 transcendentals per sample cost more
// issue slots than the 24-36 bytes per sample cost memory time.  The .ftz MUFU forms are ONE instruction each (the
// non-ftz intrinsics wrap every MUFU in a denormal pre/post-scale) and keep every quantity within ~4e-7 relative
// (1e-7 absolute near 0) — far inside the 1e-4 + 1e-6 parity tolerance.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_ftz(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 1/(1+2^(-x log2 e)): the exponential saturates to +inf / 0 at the ends and rcp(inf) = 0, so no branch is needed
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_ftz(1.f + ex2_ftz(-kLog2e * x)); }
__device__ __forceinline__ float fast_softplus(float x) {
  return fmaf(kLn2, lg2_ftz(1.f + ex2_ftz(-kLog2e * fabsf(x))), fmaxf(x, 0.f));
}

template <int N> __device__ __forceinline__ void load_n(const float* p, float* v) {  // N % 4 == 0, 16-byte aligned
#pragma unroll
  for (int i = 0; i < N; i += 4) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(p + i));
    v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
  }
}
template <int N> __device__ __forceinline__ void store_n(float* p, const float* v) {
#pragma unroll
  for (int i = 0; i < N; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}

// which ray / which part of it this lane owns.  Lanes past the last ray shadow ray R-1 (so every shuffle stays
// warp-uniform) and skip their stores.
template <int LPR> struct LaneRay {
  int sub, r;
  bool live;
  __device__ __forceinline__ LaneRay(int R) {
    const int lane = threadIdx.x & 31;
    sub = lane & (LPR - 1);
    const int ray = (blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * (32 / LPR) + lane / LPR;
    live = ray < R;
    r = live ? ray : R - 1;
  }
};

// per-lane forward state for kV consecutive samples of one ray
template <int LPR, bool RAW>
struct LaneSamples {
  static constexpr int V = kV, S = kV * LPR;
  float c[V][3];   // RAW: s = sigmoid(raw rgb) (the padded colour is s*k - pad); else the colour itself
  float rs[V];     // RAW: raw density + bias (for softplus'); else unused
  float tv[V + 1]; // t_i .. t_{i+V}
  float alpha[V], T[V];

  __device__ __forceinline__ float colour(int q, int a, float k, float pad) const { return RAW ? fmaf(c[q][a], k, -pad) : c[q][a]; }

  // loads, activations, alpha_i = 1-exp(-sigma_i delta_i |d|), T_i = prod_{j<i}(1-alpha_j)   (.cu:330-332)
  __device__ __forceinline__ void load(const float* __restrict__ rgb, const float* __restrict__ density,
                                       const float* __restrict__ t, int r, int sub, float dl, OutputAct act) {
    const long base = (long)r * S + sub * V;
    float dv[V], cv[3 * V];
    load_n<V>(density + base, dv);
    load_n<3 * V>(rgb + base * 3, cv);
    const float* tr = t + (long)r * (S + 1) + sub * V;  // rows of S+1 floats: scalar loads, neighbours share the lines
#pragma unroll
    for (int q = 0; q <= V; q++) tv[q] = __ldg(tr + q);
    float p = 1.f;
#pragma unroll
    for (int q = 0; q < V; q++) {
      float sig;
      if (RAW) {
        rs[q] = dv[q] + act.density_bias;
        sig = fast_softplus(rs[q]);
#pragma unroll
        for (int a = 0; a < 3; a++) c[q][a] = fast_sigmoid(cv[q * 3 + a]);
      } else {
        sig = dv[q];
#pragma unroll
        for (int a = 0; a < 3; a++) c[q][a] = cv[q * 3 + a];
      }
      // alpha = 1 - exp(-s) (.cu:330); for small s the 3-term series keeps alpha's RELATIVE accuracy (no cancellation)
      const float s = sig * (tv[q + 1] - tv[q]) * dl;
      const float big = 1.f - ex2_ftz(-kLog2e * s);
      const float small = s * fmaf(-0.5f * s, fmaf(s, -1.f / 3.f, 1.f), 1.f);
      alpha[q] = s < 1e-2f ? small : big;
      p *= 1.f - alpha[q];
    }
    float incl = p;  // inclusive prefix product over the ray's lanes
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) {
      const float up = __shfl_up_sync(kFull, incl, o, LPR);
      if (sub >= o) incl *= up;
    }
    float Tq = __shfl_up_sync(kFull, incl, 1, LPR);
    if (sub == 0) Tq = 1.f;
#pragma unroll
    for (int q = 0; q < V; q++) { T[q] = Tq; Tq *= 1.f - alpha[q]; }
  }
};

template <int LPR> __device__ __forceinline__ float ray_sum(float v) {  // all-lanes sum over the LPR lanes of a ray
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) { return ray_sum<32>(v); }

__device__ __forceinline__ float dir_length(const float* __restrict__ dirs, int r) {
  const float dx = __ldg(dirs + r * 3), dy = __ldg(dirs + r * 3 + 1), dz = __ldg(dirs + r * 3 + 2);
  return sqrt_ftz(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
}

template <int LPR, bool RAW>
__global__ void __launch_bounds__(kThreads)
k_composite_fwd(const float* __restrict__ rgb, const float* __restrict__ density, const float* __restrict__ t,
                const float* __restrict__ dirs, int R, int white, OutputAct act, float* __restrict__ comp_rgb,
                float* __restrict__ depth, float* __restrict__ acc, float* __restrict__ weights) {
  constexpr int V = kV, S = V * LPR;
  const LaneRay<LPR> lr(R);
  const int r = lr.r, sub = lr.sub;
  const float dl = dir_length(dirs, r);
  LaneSamples<LPR, RAW> ls;
  ls.load(rgb, density, t, r, sub, dl, act);
  const float k = 1.f + 2.f * act.rgb_padding, pad = act.rgb_padding;
  float w[V], cr = 0.f, cg = 0.f, cb = 0.f, a = 0.f, wd = 0.f;
#pragma unroll
  for (int q = 0; q < V; q++) {
    w[q] = ls.alpha[q] * ls.T[q];
    cr = fmaf(w[q], ls.colour(q, 0, k, pad), cr); cg = fmaf(w[q], ls.colour(q, 1, k, pad), cg); cb = fmaf(w[q], ls.colour(q, 2, k, pad), cb);
    a += w[q]; wd = fmaf(w[q], ls.tv[q] + ls.tv[q + 1], wd);  // 2 x sum w_i t_mid,i
  }
  if (weights && lr.live) store_n<V>(weights + (long)r * S + sub * V, w);
  cr = ray_sum<LPR>(cr); cg = ray_sum<LPR>(cg); cb = ray_sum<LPR>(cb); a = ray_sum<LPR>(a); wd = ray_sum<LPR>(wd);
  const float t_first = __shfl_sync(kFull, ls.tv[0], 0, LPR), t_last = __shfl_sync(kFull, ls.tv[V], LPR - 1, LPR);
  if (sub == 0 && lr.live) {
    if (white) { cr += 1.f - a; cg += 1.f - a; cb += 1.f - a; }  // .cu:338-340
    comp_rgb[r * 3] = cr; comp_rgb[r * 3 + 1] = cg; comp_rgb[r * 3 + 2] = cb;
    if (depth) {  // SN/MipHelpers.cs:490
      const float dv = a > 0.f ? 0.5f * wd / a : INFINITY;
      depth[r] = fminf(fmaxf(dv, t_first), t_last);
    }
    if (acc) acc[r] = a;
  }
}

template <int LPR, bool RAW>
__global__ void __launch_bounds__(kThreads, 3)  // <= 80 registers: three blocks (24 warps) per SM keep the loads in flight
k_composite_bwd(const float* __restrict__ g, const float* __restrict__ rgb, const float* __restrict__ density,
                const float* __restrict__ t, const float* __restrict__ dirs, int R, int white, int last_mode,
                OutputAct act, float* __restrict__ d_rgb, float* __restrict__ d_density) {
  constexpr int V = kV, S = V * LPR;
  const LaneRay<LPR> lr(R);
  const int r = lr.r, sub = lr.sub;
  const float dl = dir_length(dirs, r);
  const float gx = __ldg(g + r * 3), gy = __ldg(g + r * 3 + 1), gz = __ldg(g + r * 3 + 2);
  // .cu:370.  Same association as the colour dot product below, so a saturated white sample (c = 1) on a white
  // background gets dLdw = 0 exactly, as in the reference's arithmetic.
  const float dLdAcc = white ? -((gx + gy) + gz) : 0.f;
  LaneSamples<LPR, RAW> ls;
  ls.load(rgb, density, t, r, sub, dl, act);
  const float k = 1.f + 2.f * act.rgb_padding, pad = act.rgb_padding;
  // last_mode 1 drops sample S-1 like the reference kernel (A-D12)
  const bool drop_last = last_mode == 1 && sub == LPR - 1;
  // dLdw_i = g.c_i + dLdAcc (.cu:385)
  float dLdw[V], loc = 0.f;
#pragma unroll
  for (int q = 0; q < V; q++) {
    dLdw[q] = fmaf(gz, ls.colour(q, 2, k, pad), fmaf(gy, ls.colour(q, 1, k, pad), gx * ls.colour(q, 0, k, pad))) + dLdAcc;
    if (q == V - 1 && drop_last) dLdw[q] = 0.f;
    loc = fmaf(dLdw[q], ls.alpha[q] * ls.T[q], loc);
  }
  // exclusive suffix sum of loc over the ray's lanes
  float incl = loc;
#pragma unroll
  for (int o = 1; o < LPR; o <<= 1) {
    const float dn = __shfl_down_sync(kFull, incl, o, LPR);
    if (sub + o < LPR) incl += dn;
  }
  float suffix = __shfl_down_sync(kFull, incl, 1, LPR);
  if (sub == LPR - 1) suffix = 0.f;
  float dc[3 * V], ds[V];
#pragma unroll
  for (int q = V - 1; q >= 0; q--) {
    float wq = ls.alpha[q] * ls.T[q];
    const float Tnext = ls.T[q] * (1.f - ls.alpha[q]);
    float dsig = (ls.tv[q + 1] - ls.tv[q]) * dl * fmaf(Tnext, dLdw[q], -suffix);
    suffix = fmaf(dLdw[q], wq, suffix);
    if (q == V - 1 && drop_last) { dsig = 0.f; wq = 0.f; }
    float cx = gx * wq, cy = gy * wq, cz = gz * wq;  // .cu:388
    if (RAW) {  // SN/MipNerfModel.cs:184-189: softplus' = sigmoid, sigmoid' = s(1-s), colour = s*k - pad
      dsig *= fast_sigmoid(ls.rs[q]);
      cx *= fmaf(-ls.c[q][0], ls.c[q][0], ls.c[q][0]) * k;
      cy *= fmaf(-ls.c[q][1], ls.c[q][1], ls.c[q][1]) * k;
      cz *= fmaf(-ls.c[q][2], ls.c[q][2], ls.c[q][2]) * k;
    }
    ds[q] = dsig; dc[q * 3] = cx; dc[q * 3 + 1] = cy; dc[q * 3 + 2] = cz;
  }
  if (lr.live) {
    const long base = (long)r * S + sub * V;
    store_n<V>(d_density + base, ds);
    store_n<3 * V>(d_rgb + base * 3, dc);
  }
}

// Deterministic grid-wide sum for the two tiny per-ray reductions below: every block leaves its partial in
// scratch[blockIdx.x]; the block that arrives last (atomic ticket in scratch[kRedBlocks]) adds the partials in index
// order, so the result does not depend on which block that was.  Returns true on lane 0 of the finishing warp.
constexpr int kRedBlocks = 64;
__device__ __forceinline__ bool grid_sum_fixed_order(float block_part, float* scratch, float& total) {
  __shared__ bool last;
  if (gridDim.x == 1) { total = block_part; return threadIdx.x == 0; }
  unsigned* ticket = reinterpret_cast<unsigned*>(scratch + kRedBlocks);
  if (threadIdx.x == 0) {
    scratch[blockIdx.x] = block_part;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last || threadIdx.x >= 32) return false;
  __threadfence();
  float s = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) s += __ldcg(scratch + i);
  total = warp_sum(s);
  if (threadIdx.x == 0) *ticket = 0u;  // ready for the next launch on the stream
  return threadIdx.x == 0;
}
__device__ __forceinline__ float block_sum_1024(float part, float* red) {  // valid on thread 0
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  float v = 0.f;
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
  }
  return v;
}

// g = 2*lm/lm_sum*(rgb-pix)*level_mult; optional loss = sum(lm |rgb-pix|^2)/lm_sum, summed in a fixed order.
__global__ void __launch_bounds__(1024)
k_output_gradient(const float* __restrict__ comp_rgb, const float* __restrict__ pixels, const float* __restrict__ lm,
                  int R, float lm_sum, const float* __restrict__ lm_sum_dev, float level_mult, float* __restrict__ g,
                  float* __restrict__ loss_out, float* __restrict__ scratch) {
  __shared__ float red[32];
  if (lm_sum_dev) lm_sum = *lm_sum_dev;
  float part = 0.f;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
    const float l = lm[r];
    float e2 = 0.f;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float df = comp_rgb[r * 3 + a] - pixels[r * 3 + a];
      e2 += df * df;
      if (g) g[r * 3 + a] = 2.f * l / lm_sum * df * level_mult;  // .cu:356
    }
    part += l * e2;
  }
  if (!loss_out) return;
  const float v = block_sum_1024(part, red);
  float total;
  if (grid_sum_fixed_order(v, scratch, total)) *loss_out += total / lm_sum;  // accumulates over ray chunks; caller zeroes per step
}

__global__ void __launch_bounds__(1024) k_sum(const float* __restrict__ x, int n, float* __restrict__ out, float* __restrict__ scratch) {
  __shared__ float red[32];
  float part = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) part += x[i];
  const float v = block_sum_1024(part, red);
  float total;
  if (grid_sum_fixed_order(v, scratch, total)) *out = total;
}

// grads[i] *= 1 / *lm_sum_dev  (data-parallel GetGradient path: normalise the allreduced un-normalised sum)
__global__ void k_scale_by_inv(float* __restrict__ g, long n, const float* __restrict__ lm_sum_dev) {
  const float s = 1.0f / *lm_sum_dev;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) g[i] *= s;
}

template <bool RAW>
int dispatch_fwd(int S, cudaStream_t st, const float* rgb, const float* density, const float* t,
                 const float* dirs, int R, int white, OutputAct act, float* comp, float* depth, float* acc, float* w) {
#define NERF_FWD(LPR)                                                                                               \
  k_composite_fwd<LPR, RAW><<<(unsigned)cdiv(R, (kThreads / 32) * (32 / LPR)), kThreads, 0, st>>>(rgb, density, t, dirs, R, white, \
                                                                                                   act, comp, depth, acc, w)
  switch (S / kV) {
    case 4: NERF_FWD(4); break;
    case 8: NERF_FWD(8); break;
    case 16: NERF_FWD(16); break;
    case 32: NERF_FWD(32); break;
    default: return 1;
  }
#undef NERF_FWD
  return 0;
}
template <bool RAW>
int dispatch_bwd(int S, cudaStream_t st, const float* g, const float* rgb, const float* density,
                 const float* t, const float* dirs, int R, int white, int last_mode, OutputAct act, float* d_rgb,
                 float* d_den) {
#define NERF_BWD(LPR)                                                                                               \
  k_composite_bwd<LPR, RAW><<<(unsigned)cdiv(R, (kThreads / 32) * (32 / LPR)), kThreads, 0, st>>>(g, rgb, density, t, dirs, R, white, \
                                                                                                   last_mode, act, d_rgb, d_den)
  switch (S / kV) {
    case 4: NERF_BWD(4); break;
    case 8: NERF_BWD(8); break;
    case 16: NERF_BWD(16); break;
    case 32: NERF_BWD(32); break;
    default: return 1;
  }
#undef NERF_BWD
  return 0;
}
bool supported_S(int S) { return S == 32 || S == 64 || S == 128 || S == 256; }

}  // namespace

int launch_composite_fwd(const float* rgb, const float* density, const float* t, const float* dirs, int R, int S,
                         int white_bkgd, OutputAct act, float* comp_rgb, float* depth, float* acc, float* weights,
                         cudaStream_t st) {
  if (!supported_S(S)) { set_error("compositing: n_samples must be 32/64/128/256, got %d", S); return 100001; }
  if (R <= 0) return 0;
  const int rc = act.raw ? dispatch_fwd<true>(S, st, rgb, density, t, dirs, R, white_bkgd, act, comp_rgb, depth, acc, weights)
                         : dispatch_fwd<false>(S, st, rgb, density, t, dirs, R, white_bkgd, act, comp_rgb, depth, acc, weights);
  if (rc) { set_error("compositing: bad S"); return 100001; }
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_composite_bwd(const float* g, const float* rgb, const float* density, const float* t, const float* dirs,
                         int R, int S, int white_bkgd, int last_sample_mode, OutputAct act, float* d_rgb,
                         float* d_density, cudaStream_t st) {
  if (!supported_S(S)) { set_error("compositing: n_samples must be 32/64/128/256, got %d", S); return 100001; }
  if (R <= 0) return 0;
  const int rc = act.raw ? dispatch_bwd<true>(S, st, g, rgb, density, t, dirs, R, white_bkgd, last_sample_mode, act, d_rgb, d_density)
                         : dispatch_bwd<false>(S, st, g, rgb, density, t, dirs, R, white_bkgd, last_sample_mode, act, d_rgb, d_density);
  if (rc) { set_error("compositing: bad S"); return 100001; }
  NERF_CHECK_LAUNCH();
  return 0;
}

// one block per 4096 rays (a single block — the original summation order — up to 4096 rays), at most kRedBlocks
static unsigned red_grid(int n, const float* scratch) {
  if (!scratch) return 1u;
  const long b = cdiv(n, 4096);
  return (unsigned)(b < 1 ? 1 : (b > kRedBlocks ? kRedBlocks : b));
}

int launch_output_gradient(const float* comp_rgb, const float* pixels, const float* loss_mults, int R, float lm_sum,
                           const float* lm_sum_dev, float level_mult, float* g, float* loss_out, float* scratch, cudaStream_t st) {
  // without a loss to reduce the rays are independent: any grid; with one, multi-block needs the scratch
  const unsigned grid = loss_out ? red_grid(R, scratch) : (unsigned)(cdiv(R, 1024) > 1024 ? 1024 : cdiv(R, 1024));
  k_output_gradient<<<grid, 1024, 0, st>>>(comp_rgb, pixels, loss_mults, R, lm_sum, lm_sum_dev, level_mult, g, loss_out, scratch);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_sum(const float* x, int n, float* out, float* scratch, cudaStream_t st) {
  k_sum<<<red_grid(n, scratch), 1024, 0, st>>>(x, n, out, scratch);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_scale_by_inv(float* g, long n, const float* lm_sum_dev, cudaStream_t st) {
  if (n <= 0) return 0;
  const long want = cdiv(n, 256);
  k_scale_by_inv<<<(unsigned)(want > 148 * 8 ? 148 * 8 : want), 256, 0, st>>>(g, n, lm_sum_dev);
  NERF_CHECK_LAUNCH();
  return 0;
}

}  // namespace nerf

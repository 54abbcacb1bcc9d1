// compositing.cu — volumetric ray integration forward/backward and the MSE output gradient.
//
// Replaces volumetric_rendering (.cu:318-344), volumetric_rendering_gradient (.cu:362-402) and
// get_output_gradient (.cu:347-361); semantics: SURVEY Appendix B.4/B.5.
//
// HBM-bound.  The reference walks each ray serially in ONE 1024-thread block with lane stride = S floats
// (fully uncoalesced) and round-trips alpha/T/w caches.  Here: one warp per ray, V = S/32 consecutive
// samples per lane, 128-bit loads of sigma / rgb / stores of w, t staged through shared memory, the
// transmittance as a shuffle-based exclusive prefix PRODUCT, the backward as a shuffle-based suffix SUM
// (dL/dsigma_i = delta_i |d| (T_{i+1} dLdw_i - sum_{j>i} dLdw_j w_j)), alpha/T recomputed instead of cached,
// and the output activations (softplus / sigmoid, SN/MipNerfModel.cs:81-83) and their derivatives fused.
// Algorithmic bytes: fwd 24 B/sample + 32 B/ray, bwd 36 B/sample + 24 B/ray.
#include "kernels.cuh"

namespace nerf {
namespace {

constexpr int kWarpsPerBlock = 8;

// These kernels must stay HBM-bound: with libm expf/log1pf/IEEE division the ~7 transcendentals per sample cost more
// issue slots than the 24-36 bytes per sample cost memory time.  ex2.approx / lg2.approx / rcp.approx keep every
// quantity within ~4e-7 relative (1e-7 absolute near 0) — far inside the 1e-4 + 1e-6 parity tolerance.
__device__ __forceinline__ float fast_sigmoid(float x) {
  const float e = __expf(-fabsf(x));
  const float r = __fdividef(1.f, 1.f + e);
  return x >= 0.f ? r : e * r;
}
__device__ __forceinline__ float fast_softplus(float x) { return fmaxf(x, 0.f) + __logf(1.f + __expf(-fabsf(x))); }

template <int V> struct VecLoad;
template <> struct VecLoad<1> {
  static __device__ __forceinline__ void ld(const float* p, float* v) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void st(float* p, const float* v) { p[0] = v[0]; }
};
template <> struct VecLoad<2> {
  static __device__ __forceinline__ void ld(const float* p, float* v) { const float2 x = __ldg(reinterpret_cast<const float2*>(p)); v[0] = x.x; v[1] = x.y; }
  static __device__ __forceinline__ void st(float* p, const float* v) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};
template <> struct VecLoad<4> {
  static __device__ __forceinline__ void ld(const float* p, float* v) { const float4 x = __ldg(reinterpret_cast<const float4*>(p)); v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; }
  static __device__ __forceinline__ void st(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
// n consecutive floats (n = V or 3V) with the widest aligned vector op
template <int N> __device__ __forceinline__ void load_n(const float* p, float* v) {
  if constexpr (N % 4 == 0) { for (int i = 0; i < N; i += 4) VecLoad<4>::ld(p + i, v + i); }
  else if constexpr (N % 2 == 0) { for (int i = 0; i < N; i += 2) VecLoad<2>::ld(p + i, v + i); }
  else { for (int i = 0; i < N; i++) VecLoad<1>::ld(p + i, v + i); }
}
template <int N> __device__ __forceinline__ void store_n(float* p, const float* v) {
  if constexpr (N % 4 == 0) { for (int i = 0; i < N; i += 4) VecLoad<4>::st(p + i, v + i); }
  else if constexpr (N % 2 == 0) { for (int i = 0; i < N; i += 2) VecLoad<2>::st(p + i, v + i); }
  else { for (int i = 0; i < N; i++) VecLoad<1>::st(p + i, v + i); }
}

struct RayCtx { float dl; };

// shared per-lane forward state for V consecutive samples
template <int V, bool RAW>
struct LaneSamples {
  float sig[V], c[V][3], delta[V], tm[V];  // activated density, activated rgb, t_{i+1}-t_i, (t_i+t_{i+1})/2
  float rs[V], rc[V][3];                   // raw values kept for the activation derivatives (RAW only)
  float alpha[V], T[V], w[V];
  float t_first, t_last;

  __device__ __forceinline__ void load(const float* rgb, const float* density, const float* t, const float* tsm_in,
                                       float* tsm, int r, int S, int lane, OutputAct act) {
    const long base = (long)r * S + lane * V;
    float dv[V], cv[3 * V];
    load_n<V>(density + base, dv);
    load_n<3 * V>(rgb + base * 3, cv);
    // t row: coalesced into shared memory, then V+1 values per lane
    const float* tr = t + (long)r * (S + 1);
    for (int i = lane; i <= S; i += 32) tsm[i] = __ldg(tr + i);
    __syncwarp();
    (void)tsm_in;
    float tv[V + 1];
#pragma unroll
    for (int q = 0; q <= V; q++) tv[q] = tsm[lane * V + q];
    t_first = tsm[0]; t_last = tsm[S];
#pragma unroll
    for (int q = 0; q < V; q++) {
      delta[q] = tv[q + 1] - tv[q];
      tm[q] = (tv[q] + tv[q + 1]) / 2;
      if (RAW) {
        rs[q] = dv[q] + act.density_bias;
        sig[q] = fast_softplus(rs[q]);
#pragma unroll
        for (int a = 0; a < 3; a++) {
          rc[q][a] = fast_sigmoid(cv[q * 3 + a]);  // keep s = sigmoid(raw) for s(1-s)
          c[q][a] = rc[q][a] * (1.f + 2.f * act.rgb_padding) - act.rgb_padding;
        }
      } else {
        sig[q] = dv[q];
#pragma unroll
        for (int a = 0; a < 3; a++) c[q][a] = cv[q * 3 + a];
      }
    }
  }

  // alpha_i = 1-exp(-sigma_i delta_i |d|), T_i = prod_{j<i}(1-alpha_j), w_i = alpha_i T_i   (.cu:330-332)
  __device__ __forceinline__ float weights(float dl, int lane) {
    float om[V], p = 1.f;
#pragma unroll
    for (int q = 0; q < V; q++) {
      // alpha = 1 - exp(-s) (.cu:330); for small s the 3-term series keeps alpha's RELATIVE accuracy (no cancellation)
      const float s = sig[q] * delta[q] * dl;
      alpha[q] = s < 1e-2f ? s * (1.f - 0.5f * s * (1.f - s * (1.f / 3.f))) : 1.f - __expf(-s);
      om[q] = 1.f - alpha[q];
      p *= om[q];
    }
    float incl = p;  // inclusive prefix product over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl *= up;
    }
    float Tq = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) Tq = 1.f;
#pragma unroll
    for (int q = 0; q < V; q++) { T[q] = Tq; w[q] = alpha[q] * Tq; Tq *= om[q]; }
    return Tq;  // T after this lane's last sample
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int V, bool RAW>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_composite_fwd(const float* __restrict__ rgb, const float* __restrict__ density, const float* __restrict__ t,
                const float* __restrict__ dirs, int R, int white, OutputAct act, float* __restrict__ comp_rgb,
                float* __restrict__ depth, float* __restrict__ acc, float* __restrict__ weights) {
  constexpr int S = 32 * V;
  __shared__ float tsm[kWarpsPerBlock][S + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kWarpsPerBlock + warp;
  if (r >= R) return;
  const float dx = __ldg(dirs + r * 3), dy = __ldg(dirs + r * 3 + 1), dz = __ldg(dirs + r * 3 + 2);
  const float dl = sqrtf(dx * dx + dy * dy + dz * dz);
  LaneSamples<V, RAW> ls;
  ls.load(rgb, density, t, nullptr, tsm[warp], r, S, lane, act);
  ls.weights(dl, lane);
  float cr = 0.f, cg = 0.f, cb = 0.f, a = 0.f, wd = 0.f;
#pragma unroll
  for (int q = 0; q < V; q++) {
    cr += ls.w[q] * ls.c[q][0]; cg += ls.w[q] * ls.c[q][1]; cb += ls.w[q] * ls.c[q][2];
    a += ls.w[q]; wd += ls.w[q] * ls.tm[q];
  }
  if (weights) store_n<V>(weights + (long)r * S + lane * V, ls.w);
  cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); a = warp_sum(a); wd = warp_sum(wd);
  if (lane == 0) {
    if (white) { cr += 1.f - a; cg += 1.f - a; cb += 1.f - a; }  // .cu:338-340
    comp_rgb[r * 3] = cr; comp_rgb[r * 3 + 1] = cg; comp_rgb[r * 3 + 2] = cb;
    if (depth) {  // SN/MipHelpers.cs:490
      float dv = a > 0.f ? wd / a : INFINITY;
      depth[r] = fminf(fmaxf(dv, ls.t_first), ls.t_last);
    }
    if (acc) acc[r] = a;
  }
}

template <int V, bool RAW>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_composite_bwd(const float* __restrict__ g, const float* __restrict__ rgb, const float* __restrict__ density,
                const float* __restrict__ t, const float* __restrict__ dirs, int R, int white, int last_mode,
                OutputAct act, float* __restrict__ d_rgb, float* __restrict__ d_density) {
  constexpr int S = 32 * V;
  __shared__ float tsm[kWarpsPerBlock][S + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kWarpsPerBlock + warp;
  if (r >= R) return;
  const float dx = __ldg(dirs + r * 3), dy = __ldg(dirs + r * 3 + 1), dz = __ldg(dirs + r * 3 + 2);
  const float dl = sqrtf(dx * dx + dy * dy + dz * dz);
  const float gx = __ldg(g + r * 3), gy = __ldg(g + r * 3 + 1), gz = __ldg(g + r * 3 + 2);
  const float dLdAcc = white ? -(gx + gy + gz) : 0.f;  // .cu:370
  LaneSamples<V, RAW> ls;
  ls.load(rgb, density, t, nullptr, tsm[warp], r, S, lane, act);
  ls.weights(dl, lane);
  // dLdw_i = g.c_i + dLdAcc (.cu:385); last_mode 1 drops sample S-1 like the reference kernel (A-D12)
  float dLdw[V], loc = 0.f;
#pragma unroll
  for (int q = 0; q < V; q++) {
    dLdw[q] = gx * ls.c[q][0] + gy * ls.c[q][1] + gz * ls.c[q][2] + dLdAcc;
    if (last_mode == 1 && lane == 31 && q == V - 1) dLdw[q] = 0.f;
    loc += dLdw[q] * ls.w[q];
  }
  // exclusive suffix sum over lanes of loc
  float incl = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float dn = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += dn;
  }
  float suffix = __shfl_down_sync(0xffffffffu, incl, 1);
  if (lane == 31) suffix = 0.f;
  float dc[3 * V], ds[V];
#pragma unroll
  for (int q = V - 1; q >= 0; q--) {
    const float Tnext = ls.T[q] * (1.f - ls.alpha[q]);
    float dsig = ls.delta[q] * dl * (Tnext * dLdw[q] - suffix);
    float wq = ls.w[q];
    if (last_mode == 1 && lane == 31 && q == V - 1) { dsig = 0.f; wq = 0.f; }
    suffix += dLdw[q] * ls.w[q];
    float cx = gx * wq, cy = gy * wq, cz = gz * wq;  // .cu:388
    if (RAW) {  // SN/MipNerfModel.cs:184-189
      dsig *= fast_sigmoid(ls.rs[q]);
      const float k = 1.f + 2.f * act.rgb_padding;
      cx *= ls.rc[q][0] * (1.f - ls.rc[q][0]) * k;
      cy *= ls.rc[q][1] * (1.f - ls.rc[q][1]) * k;
      cz *= ls.rc[q][2] * (1.f - ls.rc[q][2]) * k;
    }
    ds[q] = dsig; dc[q * 3] = cx; dc[q * 3 + 1] = cy; dc[q * 3 + 2] = cz;
  }
  const long base = (long)r * S + lane * V;
  store_n<V>(d_density + base, ds);
  store_n<3 * V>(d_rgb + base * 3, dc);
}

// g = 2*lm/lm_sum*(rgb-pix)*level_mult; optional loss = sum(lm |rgb-pix|^2)/lm_sum.  One block, fixed order.
__global__ void __launch_bounds__(1024)
k_output_gradient(const float* __restrict__ comp_rgb, const float* __restrict__ pixels, const float* __restrict__ lm,
                  int R, float lm_sum, const float* __restrict__ lm_sum_dev, float level_mult, float* __restrict__ g,
                  float* __restrict__ loss_out) {
  __shared__ float red[32];
  if (lm_sum_dev) lm_sum = *lm_sum_dev;
  float part = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
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
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) *loss_out += v / lm_sum;  // accumulates over ray chunks; caller zeroes per step
  }
}

__global__ void __launch_bounds__(1024) k_sum(const float* __restrict__ x, int n, float* __restrict__ out) {
  __shared__ float red[32];
  float part = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) part += x[i];
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = red[threadIdx.x];
    v = warp_sum(v);
    if (threadIdx.x == 0) *out = v;
  }
}

template <bool RAW>
int dispatch_fwd(int V, dim3 grid, cudaStream_t st, const float* rgb, const float* density, const float* t,
                 const float* dirs, int R, int white, OutputAct act, float* comp, float* depth, float* acc, float* w) {
  const int th = kWarpsPerBlock * 32;
  switch (V) {
    case 1: k_composite_fwd<1, RAW><<<grid, th, 0, st>>>(rgb, density, t, dirs, R, white, act, comp, depth, acc, w); break;
    case 2: k_composite_fwd<2, RAW><<<grid, th, 0, st>>>(rgb, density, t, dirs, R, white, act, comp, depth, acc, w); break;
    case 4: k_composite_fwd<4, RAW><<<grid, th, 0, st>>>(rgb, density, t, dirs, R, white, act, comp, depth, acc, w); break;
    case 8: k_composite_fwd<8, RAW><<<grid, th, 0, st>>>(rgb, density, t, dirs, R, white, act, comp, depth, acc, w); break;
    default: return 1;
  }
  return 0;
}
template <bool RAW>
int dispatch_bwd(int V, dim3 grid, cudaStream_t st, const float* g, const float* rgb, const float* density,
                 const float* t, const float* dirs, int R, int white, int last_mode, OutputAct act, float* d_rgb,
                 float* d_den) {
  const int th = kWarpsPerBlock * 32;
  switch (V) {
    case 1: k_composite_bwd<1, RAW><<<grid, th, 0, st>>>(g, rgb, density, t, dirs, R, white, last_mode, act, d_rgb, d_den); break;
    case 2: k_composite_bwd<2, RAW><<<grid, th, 0, st>>>(g, rgb, density, t, dirs, R, white, last_mode, act, d_rgb, d_den); break;
    case 4: k_composite_bwd<4, RAW><<<grid, th, 0, st>>>(g, rgb, density, t, dirs, R, white, last_mode, act, d_rgb, d_den); break;
    case 8: k_composite_bwd<8, RAW><<<grid, th, 0, st>>>(g, rgb, density, t, dirs, R, white, last_mode, act, d_rgb, d_den); break;
    default: return 1;
  }
  return 0;
}
bool supported_S(int S) { return S == 32 || S == 64 || S == 128 || S == 256; }

}  // namespace

int launch_composite_fwd(const float* rgb, const float* density, const float* t, const float* dirs, int R, int S,
                         int white_bkgd, OutputAct act, float* comp_rgb, float* depth, float* acc, float* weights,
                         cudaStream_t st) {
  if (!supported_S(S)) { set_error("compositing: n_samples must be 32/64/128/256, got %d", S); return 100001; }
  const dim3 grid((unsigned)cdiv(R, kWarpsPerBlock));
  const int rc = act.raw ? dispatch_fwd<true>(S / 32, grid, st, rgb, density, t, dirs, R, white_bkgd, act, comp_rgb, depth, acc, weights)
                         : dispatch_fwd<false>(S / 32, grid, st, rgb, density, t, dirs, R, white_bkgd, act, comp_rgb, depth, acc, weights);
  if (rc) { set_error("compositing: bad V"); return 100001; }
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_composite_bwd(const float* g, const float* rgb, const float* density, const float* t, const float* dirs,
                         int R, int S, int white_bkgd, int last_sample_mode, OutputAct act, float* d_rgb,
                         float* d_density, cudaStream_t st) {
  if (!supported_S(S)) { set_error("compositing: n_samples must be 32/64/128/256, got %d", S); return 100001; }
  const dim3 grid((unsigned)cdiv(R, kWarpsPerBlock));
  const int rc = act.raw ? dispatch_bwd<true>(S / 32, grid, st, g, rgb, density, t, dirs, R, white_bkgd, last_sample_mode, act, d_rgb, d_density)
                         : dispatch_bwd<false>(S / 32, grid, st, g, rgb, density, t, dirs, R, white_bkgd, last_sample_mode, act, d_rgb, d_density);
  if (rc) { set_error("compositing: bad V"); return 100001; }
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_output_gradient(const float* comp_rgb, const float* pixels, const float* loss_mults, int R, float lm_sum,
                           const float* lm_sum_dev, float level_mult, float* g, float* loss_out, cudaStream_t st) {
  k_output_gradient<<<1, 1024, 0, st>>>(comp_rgb, pixels, loss_mults, R, lm_sum, lm_sum_dev, level_mult, g, loss_out);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_sum(const float* x, int n, float* out, cudaStream_t st) {
  k_sum<<<1, 1024, 0, st>>>(x, n, out);
  NERF_CHECK_LAUNCH();
  return 0;
}

}  // namespace nerf

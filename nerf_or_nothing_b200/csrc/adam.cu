// adam.cu — Adam over the flat parameter buffer (adam_optimizer_step, .cu:403-416).
//
// The reference launches one 1024-thread-block kernel per parameter tensor (22 launches,
// ANU/AcceleratedAdamOptimizer.cpp:31-39).  Parameters, gradients, m and v are each ONE flat fp32
// allocation here, so the step is a single 128-bit vectorised elementwise pass: 28 B/param of HBM
// traffic (read p,g,m,v; write p,m,v).  `grad_scale` lets the data-parallel path fold a gradient
// rescale into the same pass.
#include <type_traits>

#include "kernels.cuh"

namespace nerf {
namespace {

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float lr, float b1, float b2,
                                         float inv1, float inv2, int eps_mode) {
  m = b1 * m + (1.f - b1) * g;      // .cu:409
  v = b2 * v + (1.f - b2) * g * g;  // .cu:410
  const float mh = m * inv1, vh = v * inv2;
  if (eps_mode == 0) p -= lr * mh * rsqrtf(vh + 1e-8f);  // .cu:415 (eps inside the sqrt)
  else p -= lr * mh / (sqrtf(vh) + 1e-8f);               // SN/TrainState.cs:34
}

// DP: gs = 1 / *gs_dev (global sum of loss multipliers, allreduced with the gradient) and the normalised gradient
// is written back, so the buffers the host sees after the step hold the global mean gradient.
template <bool DP>
__global__ void __launch_bounds__(256)
k_adam(float* __restrict__ p, typename std::conditional<DP, float*, const float*>::type __restrict__ g, float* __restrict__ m,
       float* __restrict__ v, long n, float lr, float b1, float b2, float inv1, float inv2, int eps_mode, float gs,
       const float* __restrict__ gs_dev) {
  if (DP) gs = 1.0f / *gs_dev;
  const long n4 = n >> 2;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<const float4*>(g)[i];
    gg.x *= gs; gg.y *= gs; gg.z *= gs; gg.w *= gs;
    if (DP) reinterpret_cast<float4*>(const_cast<float*>(g))[i] = gg;
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, lr, b1, b2, inv1, inv2, eps_mode);
    adam_one(pp.y, gg.y, mm.y, vv.y, lr, b1, b2, inv1, inv2, eps_mode);
    adam_one(pp.z, gg.z, mm.z, vv.z, lr, b1, b2, inv1, inv2, eps_mode);
    adam_one(pp.w, gg.w, mm.w, vv.w, lr, b1, b2, inv1, inv2, eps_mode);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail (n % 4) by the first threads of block 0
  const long tail0 = n4 << 2;
  if (blockIdx.x == 0 && threadIdx.x < n - tail0) {
    const long i = tail0 + threadIdx.x;
    float pp = p[i], mm = m[i], vv = v[i];
    const float gi = g[i] * gs;
    if (DP) const_cast<float*>(g)[i] = gi;
    adam_one(pp, gi, mm, vv, lr, b1, b2, inv1, inv2, eps_mode);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

__global__ void k_fill(float* p, float v, long n) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

__global__ void k_pad_rows(const float* __restrict__ src, int sp, float* __restrict__ dst, int dp, long rows, int cols) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * dp) return;
  const long r = idx / dp;
  const int c = (int)(idx % dp);
  dst[idx] = c < cols ? src[r * sp + c] : 0.f;
}

__global__ void k_output_activations(const float* __restrict__ rd, const float* __restrict__ rr, long M, OutputAct act,
                                     float* __restrict__ den, float* __restrict__ rgb) {
  const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  den[m] = softplusf_(rd[m] + act.density_bias);
#pragma unroll
  for (int a = 0; a < 3; a++) rgb[m * 3 + a] = sigmoidf_(rr[m * 3 + a]) * (1.f + 2.f * act.rgb_padding) - act.rgb_padding;
}
__global__ void k_output_activations_grad(const float* __restrict__ rd, const float* __restrict__ rr,
                                          const float* __restrict__ dd, const float* __restrict__ dr, long M, OutputAct act,
                                          float* __restrict__ o_d, float* __restrict__ o_r) {
  const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  o_d[m] = dd[m] * sigmoidf_(rd[m] + act.density_bias);
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float s = sigmoidf_(rr[m * 3 + a]);
    o_r[m * 3 + a] = dr[m * 3 + a] * (s * (1.f - s)) * (1.f + 2.f * act.rgb_padding);
  }
}
__global__ void k_apply_act(const float* __restrict__ Z, float* __restrict__ Y, long n, int act) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float z = Z[i];
  Y[i] = act == ACT_RELU ? (z > 0.f ? z : 0.f) : act == ACT_SIGMOID ? sigmoidf_(z) : act == ACT_SOFTPLUS ? softplusf_(z) : z;
}
__global__ void k_adam_scalar(float* p, const float* g, float* m, float* v, long n, float lr, float b1, float b2,
                              float inv1, float inv2, int eps_mode, float gs) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float pp = p[i], mm = m[i], vv = v[i];
  adam_one(pp, g[i] * gs, mm, vv, lr, b1, b2, inv1, inv2, eps_mode);
  p[i] = pp; m[i] = mm; v[i] = vv;
}

}  // namespace

int launch_output_activations(const float* raw_density, const float* raw_rgb, long M, OutputAct act, float* density,
                              float* rgb, cudaStream_t st) {
  k_output_activations<<<(unsigned)cdiv(M, 256), 256, 0, st>>>(raw_density, raw_rgb, M, act, density, rgb);
  NERF_CHECK_LAUNCH();
  return 0;
}
int launch_output_activations_grad(const float* raw_density, const float* raw_rgb, const float* d_density,
                                   const float* d_rgb, long M, OutputAct act, float* d_raw_density, float* d_raw_rgb,
                                   cudaStream_t st) {
  k_output_activations_grad<<<(unsigned)cdiv(M, 256), 256, 0, st>>>(raw_density, raw_rgb, d_density, d_rgb, M, act,
                                                                    d_raw_density, d_raw_rgb);
  NERF_CHECK_LAUNCH();
  return 0;
}
int launch_apply_act(const float* Z, float* Y, long n, Act act, cudaStream_t st) {
  k_apply_act<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(Z, Y, n, (int)act);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_adam(float* p, const float* g, float* m, float* v, long n, float lr, float b1, float b2, float inv1,
                float inv2, int eps_mode, float grad_scale, cudaStream_t st) {
  if (n <= 0) return 0;
  if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) {  // unaligned per-tensor view
    k_adam_scalar<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, inv1, inv2, eps_mode, grad_scale);
    NERF_CHECK_LAUNCH();
    return 0;
  }
  // 148 SMs x 8 resident 256-thread blocks, capped by the work available
  const long want = cdiv(n >> 2, 256);
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
  k_adam<false><<<grid, 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, inv1, inv2, eps_mode, grad_scale, nullptr);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_adam_dp(float* p, float* g, float* m, float* v, long n, float lr, float b1, float b2, float inv1, float inv2,
                   int eps_mode, const float* lm_sum_dev, cudaStream_t st) {
  if (n <= 0) return 0;
  if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) { set_error("adam (data-parallel): buffers must be 16-byte aligned"); return 100001; }
  const long want = cdiv(n >> 2, 256);
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
  k_adam<true><<<grid, 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, inv1, inv2, eps_mode, 1.0f, lm_sum_dev);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_fill(float* p, float v, long n, cudaStream_t st) {
  if (n <= 0) return 0;
  const long want = cdiv(n, 256);
  k_fill<<<(unsigned)(want > 148 * 8 ? 148 * 8 : want), 256, 0, st>>>(p, v, n);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_pad_rows(const float* src, int src_pitch, float* dst, int dst_pitch, long rows, int cols, cudaStream_t st) {
  const long n = rows * dst_pitch;
  if (n <= 0) return 0;
  k_pad_rows<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(src, src_pitch, dst, dst_pitch, rows, cols);
  NERF_CHECK_LAUNCH();
  return 0;
}

}  // namespace nerf

// sampling.cu — stratified (level 0) and hierarchical (level 1) sampling of t along each ray.
//
// Replaces get_sample_t_vals (.cu:222-242) and get_resampled_t_vals (.cu:246-291); semantics follow
// SURVEY Appendix B.3 = SN/MipHelpers.cs:611-666,774-851 (the CUDA variants are defective: A-D8, A-D9).
// Index/ordering work is done with explicitly rounded fp32 ops (__fadd_rn/__fmul_rn/__fdiv_rn, never
// contracted to FMA) and in the oracle's summation order, so t-values are bit-identical to the CPU
// oracle for the same uniforms.
#include "kernels.cuh"

namespace nerf {
namespace {

__device__ __forceinline__ float lerp_near_far(float nr, float fr, int i, int S) {
  const float a = __fdiv_rn((float)i, (float)S);                                     // SN/MipHelpers.cs:616
  return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.0f, a)), __fmul_rn(fr, a));             // :622 == .cu:233
}

// one thread per (ray, edge i in [0,S])
__global__ void k_sample_t_vals(const float* __restrict__ nears, const float* __restrict__ fars,
                                SampleRng rng, int R, int S, int randomized, float* __restrict__ t) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)R * (S + 1)) return;
  const int r = (int)(idx / (S + 1)), i = (int)(idx % (S + 1));
  const float nr = nears[r], fr = fars[r];
  const float cur = lerp_near_far(nr, fr, i, S);
  if (!randomized) { t[idx] = cur; return; }
  const float upper = i < S ? __fmul_rn(0.5f, __fadd_rn(cur, lerp_near_far(nr, fr, i + 1, S))) : cur;
  const float lower = i > 0 ? __fmul_rn(0.5f, __fadd_rn(lerp_near_far(nr, fr, i - 1, S), cur)) : cur;
  const float u = rng.u ? rng.u[idx] : philox_uniform(rng.seed, rng.ray0 + r, i, rng.step, rng.level);
  t[idx] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u));                   // :629
}

// one warp per ray; S <= 256.  smem per warp: blur[S], cdf[S+1], tv[S+1].
constexpr int kResampleWarps = 4;
__global__ void __launch_bounds__(kResampleWarps * 32)
k_resample_t_vals(const float* __restrict__ t, const float* __restrict__ w, SampleRng rng, int R, int S,
                  float padding, int randomized, float* __restrict__ t_new) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kResampleWarps + warp;
  if (r >= R) return;
  float* blur = sm + warp * (3 * S + 2);
  float* cdf = blur + S;
  float* tv = cdf + S + 1;
  const float* wr = w + (long)r * S;
  const float* tr = t + (long)r * (S + 1);
  for (int i = lane; i <= S; i += 32) tv[i] = tr[i];
  // blur-pool: pad with edge values, pairwise max, average, + padding   (SN/MipHelpers.cs:645-661)
  for (int i = lane; i < S; i += 32) {
    const float wl = wr[i > 0 ? i - 1 : 0], wc = wr[i], wh = wr[i < S - 1 ? i + 1 : S - 1];
    const float m0 = wl > wc ? wl : wc, m1 = wc > wh ? wc : wh;
    blur[i] = __fadd_rn(__fmul_rn(0.5f, __fadd_rn(m0, m1)), padding);
  }
  __syncwarp();
  // sequential sums in the oracle's order (lane 0), broadcast by shuffle   (SN/MipHelpers.cs:785-793)
  float sum = 0.f;
  if (lane == 0)
    for (int i = 0; i < S; i++) sum = __fadd_rn(sum, blur[i]);
  sum = __shfl_sync(0xffffffffu, sum, 0);
  const float pad = __fsub_rn(1e-5f, sum);
  if (pad > 0.f) {
    const float per = __fdiv_rn(pad, (float)S);
    for (int i = lane; i < S; i += 32) blur[i] = __fadd_rn(blur[i], per);
    sum = __fadd_rn(sum, pad);
  }
  __syncwarp();
  for (int i = lane; i < S; i += 32) blur[i] = __fdiv_rn(blur[i], sum);  // pdf (:796)
  __syncwarp();
  if (lane == 0) {  // cdf = [0, min(1, cumsum(pdf[:-1])), 1]   (:799-812)
    float cum = 0.f;
    cdf[0] = 0.f;
    for (int i = 0; i < S - 1; i++) {
      cum = __fadd_rn(cum, blur[i]);
      cdf[i + 1] = cum < 1.0f ? cum : 1.0f;
    }
    cdf[S] = 1.0f;
  }
  __syncwarp();
  const int ns = S + 1;
  const float s1 = __fdiv_rn(1.0f, (float)ns);
  for (int s = lane; s < ns; s += 32) {
    float us;
    if (randomized) {  // :819
      const float u = rng.u ? rng.u[(long)r * ns + s] : philox_uniform(rng.seed, rng.ray0 + r, s, rng.step, rng.level);
      us = __fadd_rn(__fmul_rn((float)s, s1), __fmul_rn(u, __fsub_rn(s1, 1e-7f)));
      const float cap = __fsub_rn(1.0f, 1e-7f);
      if (us > cap) us = cap;
    } else {  // mip-NeRF linspace(0, 1-eps, ns)
      us = __fmul_rn((float)s, __fdiv_rn(__fsub_rn(1.0f, 1.1920929e-7f), (float)(ns - 1)));
    }
    int lo = 0, hi = S + 1;  // count of cdf entries <= us   (:827-832)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= us) lo = mid + 1; else hi = mid;
    }
    int idx = lo - 1;
    idx = idx < 0 ? 0 : (idx > S - 1 ? S - 1 : idx);
    const float b0 = tv[idx], b1 = tv[idx + 1], c0 = cdf[idx], c1 = cdf[idx + 1];
    const float den = __fsub_rn(c1, c0);
    float tt = den > 0.f ? __fdiv_rn(__fsub_rn(us, c0), den) : 0.f;  // :844
    tt = tt < 0.f ? 0.f : (tt > 1.f ? 1.f : tt);
    t_new[(long)r * ns + s] = __fadd_rn(b0, __fmul_rn(tt, __fsub_rn(b1, b0)));  // :847
  }
}

}  // namespace

int launch_sample_t_vals(const float* nears, const float* fars, SampleRng rng, int R, int S, int randomized,
                         float* t, cudaStream_t st) {
  const long n = (long)R * (S + 1);
  k_sample_t_vals<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(nears, fars, rng, R, S, randomized, t);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_resample_t_vals(const float* t, const float* w, SampleRng rng, int R, int S, float padding,
                           int randomized, float* t_new, cudaStream_t st) {
  if (S < 2 || S > 1024) { set_error("resample: unsupported S=%d", S); return 100001; }
  const size_t smem = (size_t)kResampleWarps * (3 * S + 2) * sizeof(float);
  k_resample_t_vals<<<(unsigned)cdiv(R, kResampleWarps), kResampleWarps * 32, smem, st>>>(
      t, w, rng, R, S, padding, randomized, t_new);
  NERF_CHECK_LAUNCH();
  return 0;
}

}  // namespace nerf

// encode.cu — conical frustum -> Gaussian (cast_rays, .cu:292-317) and integrated positional encoding +
// direction encoding (encode_input_data, .cu:187-221); SURVEY Appendix B.1/B.2.
//
// HBM-bound elementwise work (roofline: write 6*deg_point*4 B/sample when materialised).  The reference
// maps adjacent lanes to different rays (uncoalesced) and scatters 4-byte stores with stride 96; here a
// block stages a [32 samples x 6*deg] tile in shared memory and streams it out with 128-bit stores.
// The fused kernel goes straight from t-values to encodings and can emit the bf16 hi/lo planes the
// tcgen05 MLP consumes, so mean/cov never touch HBM.
#include "kernels.cuh"

namespace nerf {
namespace {

struct Gauss { float mx, my, mz, cx, cy, cz; };

// B.1 in the reference's operation order (.cu:298-316), with explicitly rounded (never FMA-contracted)
// ops: IPE multiplies the mean by up to 2^15, so a 1-ulp difference in the mean would show up as ~1e-2 rad
// in the highest frequency.  With this the Gaussian is bit-identical to the CPU oracle's for the same t.
__device__ __forceinline__ Gauss frustum_to_gaussian(float t0, float t1, float radius, float3 o, float3 d) {
#define M_(a, b) __fmul_rn(a, b)
#define A_(a, b) __fadd_rn(a, b)
#define S_(a, b) __fsub_rn(a, b)
#define D_(a, b) __fdiv_rn(a, b)
  const float mu = D_(A_(t0, t1), 2.f), hw = D_(S_(t1, t0), 2.f);
  const float mu2 = M_(mu, mu), hw2 = M_(hw, hw);
  const float den = A_(M_(3.f, mu2), hw2);
  const float t_mean = A_(mu, D_(M_(M_(2.f, mu), hw2), den));                                         // .cu:306
  const float t_var = S_(D_(hw2, 3.f), D_(M_(D_(4.f, 15.f), M_(M_(hw2, hw2), S_(M_(12.f, mu2), hw2))), M_(den, den)));  // .cu:307
  const float r_var = M_(M_(radius, radius),
                         S_(A_(D_(mu2, 4.f), M_(D_(5.f, 12.f), hw2)), D_(M_(D_(4.f, 15.f), M_(hw2, hw2)), den)));       // .cu:308
  const float ddx = M_(d.x, d.x), ddy = M_(d.y, d.y), ddz = M_(d.z, d.z);
  const float dmag = fmaxf(1e-10f, A_(A_(ddx, ddy), ddz));                                            // .cu:311
  Gauss g;
  g.mx = A_(M_(d.x, t_mean), o.x); g.my = A_(M_(d.y, t_mean), o.y); g.mz = A_(M_(d.z, t_mean), o.z);  // .cu:310
  g.cx = A_(M_(t_var, ddx), M_(r_var, S_(1.f, D_(ddx, dmag))));                                       // .cu:313-316
  g.cy = A_(M_(t_var, ddy), M_(r_var, S_(1.f, D_(ddy, dmag))));
  g.cz = A_(M_(t_var, ddz), M_(r_var, S_(1.f, D_(ddz, dmag))));
#undef M_
#undef A_
#undef S_
#undef D_
  return g;
}

// exp(-.5*var*4^f) * {sin,cos}(mean*2^f)  (.cu:185-186,196-204).
// The argument mean*2^f reaches 2^15*|x| ~ 2e5 rad, where sincosf falls into its slow Payne-Hanek path.  Instead the
// mean is converted ONCE per axis to half-turns v = mean/pi as a float-float (vh + vl, ~2^-48 relative); scaling by
// 2^f is exact, sincospif reduces its argument exactly, and the tiny tail vl*2^f enters through a second-order
// rotation.  Result: ~2 ulp of the correctly rounded sin/cos of the reference's exact argument, at a fixed cost.
struct HalfTurns { float hi, lo; };
__device__ __forceinline__ HalfTurns to_half_turns(float mean) {
  const float kInvPiHi = 0.31830987334251404f, kInvPiLo = 1.2841276486597053e-08f;
  HalfTurns v;
  v.hi = __fmul_rn(mean, kInvPiHi);
  v.lo = __fmaf_rn(mean, kInvPiHi, -v.hi) + mean * kInvPiLo;
  return v;
}
// `fast`: the output is rounded to a single bf16 plane (2^-9 relative), so after the same exact reduction to [-1, 1]
// half-turns the SFU sin/cos (absolute error < 5e-7 on [-pi, pi]) replace sincospif's polynomials.
__device__ __forceinline__ void ipe_pair(HalfTurns v, float var, float scale, float& s, float& c, bool fast = false) {
  const float x = __fmul_rn(0.5f, __fmul_rn(__fmul_rn(var, scale), scale));  // .5*var*4^f, the reference's rounding
  if (x > 87.f) { s = 0.f; c = 0.f; return; }                               // exp(-x) < 1.2e-38: below fp32 normals
  const float e = exp2f(-1.4426950216293335f * x);
  float s0, c0;
  if (fast) {
    float a = v.hi * scale;            // exact
    a = a - 2.f * rintf(0.5f * a);     // exact: a mod 2 in [-1, 1]
    s0 = __sinf(3.1415927410125732f * a);
    c0 = __cosf(3.1415927410125732f * a);
  } else {
    sincospif(v.hi * scale, &s0, &c0);  // exact scaling; sin/cos(pi * a)
  }
  const float d = 3.1415927410125732f * (v.lo * scale);
  const float q = fmaf(-0.5f * d, d, 1.0f);
  s = e * fmaf(d, c0, s0 * q);
  c = e * fmaf(-d, s0, c0 * q);
}

__global__ void k_cast_rays(const float* __restrict__ t, const float* __restrict__ o, const float* __restrict__ d,
                            const float* __restrict__ radii, int R, int S, float* __restrict__ means,
                            float* __restrict__ covs) {
  const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= (long)R * S) return;
  const int r = (int)(m / S), s = (int)(m % S);
  const float3 oo = make_float3(o[r * 3], o[r * 3 + 1], o[r * 3 + 2]);
  const float3 dd = make_float3(d[r * 3], d[r * 3 + 1], d[r * 3 + 2]);
  const Gauss g = frustum_to_gaussian(t[(long)r * (S + 1) + s], t[(long)r * (S + 1) + s + 1], radii[r], oo, dd);
  means[m * 3] = g.mx; means[m * 3 + 1] = g.my; means[m * 3 + 2] = g.mz;
  covs[m * 3] = g.cx; covs[m * 3 + 1] = g.cy; covs[m * 3 + 2] = g.cz;
}

constexpr int kTileSamples = 32;  // samples per block iteration
constexpr int kFreqLanes = 8;     // threads cooperating on one sample (each takes deg/8 frequencies)

// FUSED=true : inputs t,o,d,radii.   FUSED=false: inputs means,covs (per-stage entry).
template <bool FUSED>
__global__ void __launch_bounds__(kTileSamples * kFreqLanes)
k_encode_pos(const float* __restrict__ in0, const float* __restrict__ in1, const float* __restrict__ o,
             const float* __restrict__ d, const float* __restrict__ radii, long M, int S, int deg,
             float* __restrict__ enc_f32, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
             int pitch_h) {
  extern __shared__ float tile[];  // [kTileSamples][P]
  __shared__ Gauss gs[kTileSamples];
  const int P = 6 * deg;
  const int ls = threadIdx.x / kFreqLanes, fl = threadIdx.x % kFreqLanes;
  const long m0 = (long)blockIdx.x * kTileSamples;
  const long m = m0 + ls;
  // the Gaussian of a sample (a dozen IEEE divisions) is computed ONCE, by one lane of the first warp, not by each of
  // the 8 lanes that share the sample's frequencies: a warp issues an instruction for all its samples at the price of one
  if (threadIdx.x < kTileSamples && m0 + threadIdx.x < M) {
    const long mm = m0 + threadIdx.x;
    Gauss g;
    if (FUSED) {
      const int r = (int)(mm / S), s = (int)(mm % S);
      const float3 oo = make_float3(o[r * 3], o[r * 3 + 1], o[r * 3 + 2]);
      const float3 dd = make_float3(d[r * 3], d[r * 3 + 1], d[r * 3 + 2]);
      g = frustum_to_gaussian(in0[(long)r * (S + 1) + s], in0[(long)r * (S + 1) + s + 1], radii[r], oo, dd);
    } else {
      g.mx = in0[mm * 3]; g.my = in0[mm * 3 + 1]; g.mz = in0[mm * 3 + 2];
      g.cx = in1[mm * 3]; g.cy = in1[mm * 3 + 1]; g.cz = in1[mm * 3 + 2];
    }
    gs[threadIdx.x] = g;
  }
  __syncthreads();
  if (m < M) {
    const Gauss g = gs[ls];
    const HalfTurns hx = to_half_turns(g.mx), hy = to_half_turns(g.my), hz = to_half_turns(g.mz);
    const bool fast = enc_f32 == nullptr && lo == nullptr;  // single bf16 plane out
    for (int f = fl; f < deg; f += kFreqLanes) {
      const float scale = (float)(1u << f);  // .cu:196
      float* e = tile + ls * P + f * 6;
      ipe_pair(hx, g.cx, scale, e[0], e[3], fast);
      ipe_pair(hy, g.cy, scale, e[1], e[4], fast);
      ipe_pair(hz, g.cz, scale, e[2], e[5], fast);
    }
  }
  __syncthreads();
  const int rows = (int)min((long)kTileSamples, M - m0);
  const int n = rows * P;
  if (enc_f32) {  // rows are contiguous in global memory: one linear, 128-bit coalesced copy
    float* dst = enc_f32 + m0 * P;
    if ((P & 3) == 0) {
      for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4)
        *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(tile + i);
    } else {
      for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = tile[i];
    }
  }
  if (hi) {  // bf16 split planes: x ~= hi + lo  (lo optional).  One warp per row: conflict-free float2 reads, 128-byte stores,
             // the zero K padding [P, pitch_h) in the same sweep, no integer division
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    for (int row = warp; row < rows; row += n_warps) {
      const float* trow = tile + row * P;
      const long off0 = (m0 + row) * pitch_h;
      for (int col = lane * 2; col < pitch_h; col += 64) {
        float a = 0.f, b = 0.f;
        if (col < P) { const float2 v = *reinterpret_cast<const float2*>(trow + col); a = v.x; b = v.y; }  // P even
        const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
        *reinterpret_cast<__nv_bfloat162*>(hi + off0 + col) = __nv_bfloat162(ah, bh);
        if (lo)
          *reinterpret_cast<__nv_bfloat162*>(lo + off0 + col) =
              __nv_bfloat162(__float2bfloat16_rn(a - __bfloat162float(ah)), __float2bfloat16_rn(b - __bfloat162float(bh)));
      }
    }
  }
}

// direction PE per SAMPLE: [d, sin(2^0 d), cos(2^0 d), ...] (SN/MipHelpers.cs:337-356, A-D10), from the
// per-ray direction.  One block per ray: the 3+6*deg values are computed once into shared memory (as the fp32 row
// and/or the zero-padded bf16 hi/lo rows), then replicated to the ray's S sample rows with 16-byte stores.
__global__ void __launch_bounds__(128)
k_encode_dir(const float* __restrict__ d, int R, int S, int deg, float* __restrict__ f32, int pitch_f,
             __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int pitch_h) {
  __shared__ __align__(16) float row_f[64];
  __shared__ __align__(16) __nv_bfloat16 row_h[64], row_l[64];
  const int r = blockIdx.x;
  const int Dd = 3 + 6 * deg;
  if (threadIdx.x < 64) {
    const int c = threadIdx.x;
    float v = 0.f;
    if (c < 3) v = d[r * 3 + c];
    else if (c < Dd) {
      const int j = (c - 3) / 6, k = (c - 3) % 6;
      const float x = d[r * 3 + (k % 3)] * (float)(1u << j);
      v = k < 3 ? sinf(x) : cosf(x);
    }
    row_f[c] = v;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    row_h[c] = h;
    row_l[c] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
  __syncthreads();
  const long m0 = (long)r * S;
  if (f32) {
    for (int i = threadIdx.x; i < S * pitch_f; i += blockDim.x) {
      const int c = i % pitch_f;
      f32[m0 * pitch_f + i] = c < Dd ? row_f[c] : 0.f;
    }
  }
  if (hi) {  // pitch_h is a multiple of 64 elements; columns >= 64 (if any) are zero
    const int chunks = pitch_h >> 3;  // 16-byte chunks per row
    for (int i = threadIdx.x; i < S * chunks; i += blockDim.x) {
      const int s = i / chunks, c = i % chunks;
      const uint4 z = make_uint4(0, 0, 0, 0);
      const uint4 vh = c < 8 ? reinterpret_cast<const uint4*>(row_h)[c] : z;
      reinterpret_cast<uint4*>(hi + (m0 + s) * pitch_h)[c] = vh;
      if (lo) reinterpret_cast<uint4*>(lo + (m0 + s) * pitch_h)[c] = c < 8 ? reinterpret_cast<const uint4*>(row_l)[c] : z;
    }
  }
}

}  // namespace

int launch_cast_rays(const float* t, const float* o, const float* d, const float* radii, int R, int S,
                     float* means, float* covs, cudaStream_t st) {
  const long M = (long)R * S;
  k_cast_rays<<<(unsigned)cdiv(M, 256), 256, 0, st>>>(t, o, d, radii, R, S, means, covs);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_encode_input_data(const float* means, const float* covs, const float* dirs, float* enc_pos,
                             float* enc_dir, int R, int S, int deg_point, int deg_view, cudaStream_t st) {
  const long M = (long)R * S;
  const size_t smem = (size_t)kTileSamples * 6 * deg_point * sizeof(float);
  k_encode_pos<false><<<(unsigned)cdiv(M, kTileSamples), kTileSamples * kFreqLanes, smem, st>>>(
      means, covs, nullptr, nullptr, nullptr, M, S, deg_point, enc_pos, nullptr, nullptr, 0);
  NERF_CHECK_LAUNCH();
  if (enc_dir) {
    if (3 + 6 * deg_view > 64) { set_error("encode: deg_view too large"); return 100001; }
    k_encode_dir<<<(unsigned)R, 128, 0, st>>>(dirs, R, S, deg_view, enc_dir, 3 + 6 * deg_view, nullptr, nullptr, 0);
    NERF_CHECK_LAUNCH();
  }
  return 0;
}

int launch_cast_encode_fused(const float* t, const float* o, const float* d, const float* radii, int R, int S,
                             int deg_point, int deg_view, EncodeOut out, cudaStream_t st) {
  const long M = (long)R * S;
  const size_t smem = (size_t)kTileSamples * 6 * deg_point * sizeof(float);
  k_encode_pos<true><<<(unsigned)cdiv(M, kTileSamples), kTileSamples * kFreqLanes, smem, st>>>(
      t, nullptr, o, d, radii, M, S, deg_point, out.enc_pos_f32, out.pos_hi, out.pos_lo, out.pos_pitch_h);
  NERF_CHECK_LAUNCH();
  if (out.enc_dir_f32 || out.dir_hi) {
    if (3 + 6 * deg_view > 64) { set_error("encode: deg_view too large"); return 100001; }
    k_encode_dir<<<(unsigned)R, 128, 0, st>>>(d, R, S, deg_view, out.enc_dir_f32, out.dir_pitch_f32, out.dir_hi, out.dir_lo,
                                              out.dir_pitch_h);
    NERF_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace nerf

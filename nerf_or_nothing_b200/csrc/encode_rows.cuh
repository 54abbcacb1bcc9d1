// encode_rows.cuh — device arithmetic of cast_rays (.cu:292-317) and encode_input_data (.cu:187-221), SURVEY B.1/B.2,
// shared by the stand-alone encode kernels (encode.cu) and by the encoder warps INSIDE the fused MLP kernels
// (mlp_fused.cu, mlp_fused_split.cu), which build the layer-0 / skip-layer / condition-layer A operands themselves so
// that no encode kernel runs and the encodings never make an HBM round trip on the render path.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace nerf {

// where a level's samples come from: t-values [R, S+1] and the ray batch (device pointers)
struct RaySource {
  const float *t = nullptr, *o = nullptr, *d = nullptr, *radii = nullptr;
  int R = 0, S = 0, deg_point = 0, deg_view = 0;
};

#ifdef __CUDACC__
namespace enc {

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


__device__ __forceinline__ uint32_t bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 8 consecutive features -> one 16-byte chunk of the hi plane (and of the lo plane: x - hi, if lo != nullptr)
__device__ __forceinline__ void store_chunk8(const float (&x)[8], __nv_bfloat16* hi, __nv_bfloat16* lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    h[q] = bf16x2(x[2 * q], x[2 * q + 1]);
    l[q] = bf16x2(x[2 * q] - __uint_as_float(h[q] << 16), x[2 * q + 1] - __uint_as_float(h[q] & 0xFFFF0000u));
  }
  *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
  if (lo) *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

// One thread = one sample row.  Writes row `out_row` of the encoding planes the tensor-core MLP reads: position
// [*, 128] (6 * deg_point features, zero padded; needs deg_point % 4 == 0 and 6 * deg_point <= 128) and direction [*, 64]
// (3 + 6 * deg_view features, deg_view <= 4, zero padded), each as a bf16 hi plane and — fp32-accurate mode — a lo plane
// with x ~= hi + lo.  Same operations in the same order as k_encode_pos / k_encode_dir (encode.cu): the planes are
// bit-identical to what those kernels write.  `valid` = false (row beyond the batch) writes zeros.
// FAST = single bf16 plane out (encode.cu: `fast`); a template parameter and a rolled frequency loop keep the code small —
// the encoder shares the instruction cache with the MMA issue and epilogue loops of the fused kernels.
template <bool FAST>
__device__ __noinline__ void encode_row_to_planes(const RaySource& rs, long m, bool valid, long out_row, __nv_bfloat16* pos_hi,
                                                  __nv_bfloat16* pos_lo, __nv_bfloat16* dir_hi, __nv_bfloat16* dir_lo) {
  __nv_bfloat16* ph = pos_hi + out_row * 128;
  __nv_bfloat16* pl = pos_lo ? pos_lo + out_row * 128 : nullptr;
  __nv_bfloat16* dh = dir_hi + out_row * 64;
  __nv_bfloat16* dl = dir_lo ? dir_lo + out_row * 64 : nullptr;
  const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (!valid) {
#pragma unroll
    for (int c = 0; c < 16; c++) store_chunk8(zero8, ph + 8 * c, pl ? pl + 8 * c : nullptr);
#pragma unroll
    for (int c = 0; c < 8; c++) store_chunk8(zero8, dh + 8 * c, dl ? dl + 8 * c : nullptr);
    return;
  }
  const int r = (int)(m / rs.S), s = (int)(m % rs.S);
  const float3 oo = make_float3(__ldg(rs.o + r * 3), __ldg(rs.o + r * 3 + 1), __ldg(rs.o + r * 3 + 2));
  const float3 dd = make_float3(__ldg(rs.d + r * 3), __ldg(rs.d + r * 3 + 1), __ldg(rs.d + r * 3 + 2));
  const Gauss g = frustum_to_gaussian(__ldg(rs.t + (long)r * (rs.S + 1) + s), __ldg(rs.t + (long)r * (rs.S + 1) + s + 1),
                                      __ldg(rs.radii + r), oo, dd);
  const HalfTurns hx = to_half_turns(g.mx), hy = to_half_turns(g.my), hz = to_half_turns(g.mz);
  constexpr bool fast = FAST;
  // four frequencies = 24 features = three 16-byte chunks per pass
#pragma unroll 1
  for (int f0 = 0; f0 < 20; f0 += 4) {
    const int c0 = (f0 >> 2) * 3;
    if (f0 >= rs.deg_point) {
#pragma unroll
      for (int c = 0; c < 3; c++) store_chunk8(zero8, ph + 8 * (c0 + c), pl ? pl + 8 * (c0 + c) : nullptr);
      continue;
    }
    float e[24];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float scale = (float)(1u << (f0 + j));  // .cu:196
      ipe_pair(hx, g.cx, scale, e[6 * j + 0], e[6 * j + 3], fast);
      ipe_pair(hy, g.cy, scale, e[6 * j + 1], e[6 * j + 4], fast);
      ipe_pair(hz, g.cz, scale, e[6 * j + 2], e[6 * j + 5], fast);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float x[8];
#pragma unroll
      for (int q = 0; q < 8; q++) x[q] = e[8 * c + q];
      store_chunk8(x, ph + 8 * (c0 + c), pl ? pl + 8 * (c0 + c) : nullptr);
    }
  }
  store_chunk8(zero8, ph + 120, pl ? pl + 120 : nullptr);  // columns 120..127 (20 frequencies fill 0..119)
  // direction PE: [d, sin(2^0 d), cos(2^0 d), ...] (SN/MipHelpers.cs:337-356, A-D10)
  float v[32];
  const float dv[3] = {dd.x, dd.y, dd.z};
#pragma unroll
  for (int c = 0; c < 32; c++) {
    float val = 0.f;
    if (c < 3) val = dv[c];
    else if (c < 27) {
      const int j = (c - 3) / 6, k = (c - 3) % 6;
      if (j < rs.deg_view) {
        const float x = dv[k % 3] * (float)(1u << j);
        val = k < 3 ? sinf(x) : cosf(x);
      }
    }
    v[c] = val;
  }
#pragma unroll
  for (int c = 0; c < 4; c++) {
    float x[8];
#pragma unroll
    for (int q = 0; q < 8; q++) x[q] = v[8 * c + q];
    store_chunk8(x, dh + 8 * c, dl ? dl + 8 * c : nullptr);
  }
#pragma unroll
  for (int c = 4; c < 8; c++) store_chunk8(zero8, dh + 8 * c, dl ? dl + 8 * c : nullptr);
}

}  // namespace enc
#endif  // __CUDACC__

}  // namespace nerf

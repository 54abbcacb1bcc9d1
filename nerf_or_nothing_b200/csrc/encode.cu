// encode.cu — conical frustum -> Gaussian (cast_rays, .cu:292-317) and integrated positional encoding +
// direction encoding (encode_input_data, .cu:187-221); SURVEY Appendix B.1/B.2.
//
// HBM-bound elementwise work (roofline: write 6*deg_point*4 B/sample when materialised).  The reference
// maps adjacent lanes to different rays (uncoalesced) and scatters 4-byte stores with stride 96; here a
// block stages a [32 samples x 6*deg] tile in shared memory and streams it out with 128-bit stores.
// The fused kernel goes straight from t-values to encodings and can emit the bf16 hi/lo planes the
// tcgen05 MLP consumes, so mean/cov never touch HBM.
#include <cuda_fp16.h>

#include "encode_rows.cuh"
#include "kernels.cuh"

namespace nerf {
namespace {

using namespace enc;

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
             int pitch_h, __half* __restrict__ h16) {
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
        if (h16) *reinterpret_cast<__half2*>(h16 + off0 + col) = __floats2half2_rn(a, b);
      }
    }
  }
}

// direction PE per SAMPLE: [d, sin(2^0 d), cos(2^0 d), ...] (SN/MipHelpers.cs:337-356, A-D10), from the
// per-ray direction.  One block per ray: the 3+6*deg values are computed once into shared memory (as the fp32 row
// and/or the zero-padded bf16 hi/lo rows), then replicated to the ray's S sample rows with 16-byte stores.
__global__ void __launch_bounds__(128)
k_encode_dir(const float* __restrict__ d, int R, int S, int deg, float* __restrict__ f32, int pitch_f,
             __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int pitch_h, __half* __restrict__ h16) {
  __shared__ __align__(16) float row_f[64];
  __shared__ __align__(16) __nv_bfloat16 row_h[64], row_l[64];
  __shared__ __align__(16) __half row_16[64];
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
    row_16[c] = __float2half_rn(v);
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
      if (h16) reinterpret_cast<uint4*>(h16 + (m0 + s) * pitch_h)[c] = c < 8 ? reinterpret_cast<const uint4*>(row_16)[c] : z;
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
      means, covs, nullptr, nullptr, nullptr, M, S, deg_point, enc_pos, nullptr, nullptr, 0, nullptr);
  NERF_CHECK_LAUNCH();
  if (enc_dir) {
    if (3 + 6 * deg_view > 64) { set_error("encode: deg_view too large"); return 100001; }
    k_encode_dir<<<(unsigned)R, 128, 0, st>>>(dirs, R, S, deg_view, enc_dir, 3 + 6 * deg_view, nullptr, nullptr, 0, nullptr);
    NERF_CHECK_LAUNCH();
  }
  return 0;
}

int launch_cast_encode_fused(const float* t, const float* o, const float* d, const float* radii, int R, int S,
                             int deg_point, int deg_view, EncodeOut out, cudaStream_t st) {
  const long M = (long)R * S;
  const size_t smem = (size_t)kTileSamples * 6 * deg_point * sizeof(float);
  k_encode_pos<true><<<(unsigned)cdiv(M, kTileSamples), kTileSamples * kFreqLanes, smem, st>>>(
      t, nullptr, o, d, radii, M, S, deg_point, out.enc_pos_f32, out.pos_hi, out.pos_lo, out.pos_pitch_h, static_cast<__half*>(out.pos_f16));
  NERF_CHECK_LAUNCH();
  if (out.enc_dir_f32 || out.dir_hi) {
    if (3 + 6 * deg_view > 64) { set_error("encode: deg_view too large"); return 100001; }
    k_encode_dir<<<(unsigned)R, 128, 0, st>>>(d, R, S, deg_view, out.enc_dir_f32, out.dir_pitch_f32, out.dir_hi, out.dir_lo,
                                              out.dir_pitch_h, static_cast<__half*>(out.dir_f16));
    NERF_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace nerf

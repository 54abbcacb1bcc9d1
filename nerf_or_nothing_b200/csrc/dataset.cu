// dataset.cu — device-resident ray dataset and on-device batch assembly, image metrics (row 3).
//
// Replaces SN/BinDataset.cs:27-52 (LoadBatch: BatchSize random 64-byte records read one by one from train_data.bin with a
// file seek each, then six small host->device copies per step in ANU/AcceleratedMipNeRF.cpp:66-79).  Here the records
// live in HBM once; a step draws its record indices on the device from the counter-based Philox stream and gathers them
// straight into the structure-of-arrays batch the kernels read — no per-step host traffic at all.
// Record (16 floats, SN/BinDataset.cs:40-49): origin(3) direction(3) viewdir(3) radius near far lossmult rgb(3).
#include <cmath>

#include "kernels.cuh"

namespace nerf {
namespace {

// Philox stream of the batch sampler: counter = (global ray slot, 0, step, 0xDA7A5E7), key = seed; index = floor(x * n / 2^32)
__device__ __forceinline__ long draw_index(uint64_t seed, uint32_t slot, uint32_t step, long n) {
  const uint32_t x = philox4x32_10_w0(slot, 0u, step, 0x0DA7A5E7u, (uint32_t)seed, (uint32_t)(seed >> 32));
  return (long)(((unsigned long long)x * (unsigned long long)n) >> 32);
}

__global__ void k_draw_indices(uint64_t seed, uint32_t slot0, uint32_t step, long n, int R, long* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R) idx[i] = draw_index(seed, slot0 + (uint32_t)i, step, n);
}

// one thread per ray: 64-byte record in (4 x 16-byte loads), SoA out
__global__ void k_gather_batch(const float4* __restrict__ rec, long n, const long* __restrict__ idx, uint64_t seed, uint32_t slot0,
                               uint32_t step, int R, float* __restrict__ o, float* __restrict__ d, float* __restrict__ radii,
                               float* __restrict__ nears, float* __restrict__ fars, float* __restrict__ lm, float* __restrict__ pix) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  long j = idx ? idx[i] : draw_index(seed, slot0 + (uint32_t)i, step, n);
  j = j < 0 ? 0 : (j >= n ? n - 1 : j);
  const float4 a = __ldg(rec + j * 4), b = __ldg(rec + j * 4 + 1), c = __ldg(rec + j * 4 + 2), e = __ldg(rec + j * 4 + 3);
  o[i * 3] = a.x; o[i * 3 + 1] = a.y; o[i * 3 + 2] = a.z;
  d[i * 3] = a.w; d[i * 3 + 1] = b.x; d[i * 3 + 2] = b.y;
  // b.z b.w c.x = viewdir: loaded and dropped exactly as the reference does (SN/BinDataset.cs:44); the direction PE encodes the
  // UN-normalised `direction` (SURVEY A-D10), so the field only stays in the record for layout parity
  radii[i] = c.y; nears[i] = c.z; fars[i] = c.w; lm[i] = e.x;
  pix[i * 3] = e.y; pix[i * 3 + 1] = e.z; pix[i * 3 + 2] = e.w;
}

// sum of squared differences over n floats -> out[0] (fp64 accumulation; one launch, atomics on a zeroed double)
__global__ void k_sq_err(const float* __restrict__ a, const float* __restrict__ b, long n, double* __restrict__ out) {
  double s = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double d = (double)a[i] - (double)b[i];
    s += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += ws[w];
    atomicAdd(out, t);
  }
}

// ---- Dataset.GenerateRays (SN/Dataset.cs:111-176) for one camera, on the device ---------------------------------------------
// Pixel p = y * W + x of the view -> camera direction ((x - W/2 + .5)/f, -(y - H/2 + .5)/f, -1), rotated by the camera-to-world
// rotation (rows dotted left to right, separately rounded like the C#), origin = translation, radius = |d(x) - d(x+1)| * 2 / sqrt(12).
// At the last column the reference differences a pixel with itself (radius 0, :151); edge_mode 1 uses the left neighbour instead
// (mip-NeRF's own behaviour, and what nerf_or_nothing_b200/scene.py generates).
__device__ __forceinline__ float3 view_dir(const float* __restrict__ c, float focal, int W, int H, int x, int y) {
  const float dx = __fdiv_rn(__fadd_rn(__fsub_rn((float)x, __fmul_rn((float)W, 0.5f)), 0.5f), focal);
  const float dy = -__fdiv_rn(__fadd_rn(__fsub_rn((float)y, __fmul_rn((float)H, 0.5f)), 0.5f), focal);
  const float dz = -1.f;
  float3 d;
  d.x = __fadd_rn(__fadd_rn(__fmul_rn(c[0], dx), __fmul_rn(c[1], dy)), __fmul_rn(c[2], dz));
  d.y = __fadd_rn(__fadd_rn(__fmul_rn(c[4], dx), __fmul_rn(c[5], dy)), __fmul_rn(c[6], dz));
  d.z = __fadd_rn(__fadd_rn(__fmul_rn(c[8], dx), __fmul_rn(c[9], dy)), __fmul_rn(c[10], dz));
  return d;
}
struct Pose { float c[12]; };  // 3 x 4 row-major [R | t]
__global__ void k_generate_rays(Pose pose, float focal, int W, int H, float near, float far, int edge_mode, long first, long n,
                                float* __restrict__ o, float* __restrict__ d, float* __restrict__ radii, float* __restrict__ nears,
                                float* __restrict__ fars) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long p = first + i;
  const int y = (int)(p / W), x = (int)(p % W);
  const float3 dd = view_dir(pose.c, focal, W, H, x, y);
  int nx = x < W - 1 ? x + 1 : x;                       // Dataset.cs:151
  if (edge_mode == 1 && x == W - 1 && W > 1) nx = x - 1;
  const float3 dn = view_dir(pose.c, focal, W, H, nx, y);
  const float ex = __fsub_rn(dd.x, dn.x), ey = __fsub_rn(dd.y, dn.y), ez = __fsub_rn(dd.z, dn.z);
  const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
  o[i * 3] = pose.c[3]; o[i * 3 + 1] = pose.c[7]; o[i * 3 + 2] = pose.c[11];
  d[i * 3] = dd.x; d[i * 3 + 1] = dd.y; d[i * 3 + 2] = dd.z;
  radii[i] = __fdiv_rn(__fmul_rn(len, 2.f), 3.4641016151377544f);  // * 2 / MathF.Sqrt(12)  (:152)
  nears[i] = near; fars[i] = far;
}

// ---- SSIM (SN/MipHelpers.cs:688-757 ComputeSsim / ComputeSsimAverage; VectorImage.Convolve :903-927) -----------------
// The reference convolves five whole images (a, b, a^2, b^2, ab) with a normalised size x size Gaussian over a ZERO-padded
// border, one scalar triple loop per image.  Here a block owns a 16 x 16 pixel tile: both images' (16 + size - 1)^2 halo
// tiles are staged once in shared memory, each thread accumulates the five windowed sums of its pixel and channel in the
// reference's tap order (kx outer, ky inner), and the SSIM map value is formed in registers.  Image layout: [height,
// width, 3] floats — pixel (x, y) of the reference's VectorImage[x, y] at (y * width + x) * 3 (the filter is symmetric, so
// the result does not depend on which axis is called x).
constexpr int kSsimTile = 16, kSsimMaxFilter = 15;
struct SsimFilter { float w[kSsimMaxFilter * kSsimMaxFilter]; };

__global__ void __launch_bounds__(kSsimTile * kSsimTile)
k_ssim(const float* __restrict__ a, const float* __restrict__ b, int W, int H, int fs, SsimFilter filt, float c1, float c2,
       float* __restrict__ map_out, double* __restrict__ block_sums) {
  extern __shared__ float tile[];  // [2][span][span][3]
  const int pad = fs / 2, span = kSsimTile + fs - 1;
  const int x0 = blockIdx.x * kSsimTile - pad, y0 = blockIdx.y * kSsimTile - pad;
  float* ta = tile;
  float* tb = tile + span * span * 3;
  for (int i = threadIdx.x; i < span * span; i += blockDim.x) {
    const int ty = i / span, tx = i % span, x = x0 + tx, y = y0 + ty;
    const bool in = x >= 0 && x < W && y >= 0 && y < H;  // zero padding (MipHelpers.cs:909-911)
    const long g = ((long)y * W + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; c++) { ta[i * 3 + c] = in ? a[g + c] : 0.f; tb[i * 3 + c] = in ? b[g + c] : 0.f; }
  }
  __syncthreads();
  const int lx = threadIdx.x % kSsimTile, ly = threadIdx.x / kSsimTile;
  const int x = blockIdx.x * kSsimTile + lx, y = blockIdx.y * kSsimTile + ly;
  double part = 0.0;
  if (x < W && y < H) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float m0 = 0.f, m1 = 0.f, s00 = 0.f, s11 = 0.f, s01 = 0.f;
      for (int kx = 0; kx < fs; kx++)
        for (int ky = 0; ky < fs; ky++) {  // MipHelpers.cs:917-923
          const float w = filt.w[kx * fs + ky];
          const float va = ta[((ly + ky) * span + lx + kx) * 3 + c], vb = tb[((ly + ky) * span + lx + kx) * 3 + c];
          m0 = __fadd_rn(m0, __fmul_rn(va, w)); m1 = __fadd_rn(m1, __fmul_rn(vb, w));
          s00 = __fadd_rn(s00, __fmul_rn(__fmul_rn(va, va), w)); s11 = __fadd_rn(s11, __fmul_rn(__fmul_rn(vb, vb), w));
          s01 = __fadd_rn(s01, __fmul_rn(__fmul_rn(va, vb), w));
        }
      const float mu00 = __fmul_rn(m0, m0), mu11 = __fmul_rn(m1, m1), mu01 = __fmul_rn(m0, m1);
      const float g00 = fmaxf(__fsub_rn(s00, mu00), 0.f), g11 = fmaxf(__fsub_rn(s11, mu11), 0.f), g01 = fmaxf(__fsub_rn(s01, mu01), 0.f);  // :705-712
      const float num = __fmul_rn(__fadd_rn(__fmul_rn(mu01, 2.f), c1), __fadd_rn(__fmul_rn(g01, 2.f), c2));  // :723
      const float den = __fmul_rn(__fadd_rn(__fadd_rn(mu00, mu11), c1), __fadd_rn(__fadd_rn(g00, g11), c2));  // :724
      const float v = __fdiv_rn(num, den);
      if (map_out) map_out[((long)y * W + x) * 3 + c] = v;
      part += (double)v;
    }
  }
  // fixed-order block sum (fp64), one partial per block; the host adds the partials in block order
  __shared__ double red[kSsimTile * kSsimTile / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kSsimTile * kSsimTile / 32; w++) t += red[w];
    block_sums[blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

}  // namespace

// mean SSIM over pixels and channels; block_sums: >= ssim_blocks(W, H) doubles of device scratch
long ssim_blocks(int W, int H) { return cdiv(W, kSsimTile) * cdiv(H, kSsimTile); }
int launch_ssim(const float* a, const float* b, int W, int H, float max_val, int filter_size, float filter_sigma, float k1, float k2,
                float* map_out, double* block_sums, cudaStream_t st) {
  if (filter_size < 1 || filter_size > kSsimMaxFilter || !(filter_size & 1)) { set_error("ssim: filter size must be odd and <= %d", kSsimMaxFilter); return 100001; }
  SsimFilter f;
  {  // CreateGaussianFilter (MipHelpers.cs:739-756), fp32 like the original
    const int hs = filter_size / 2;
    float sum = 0.f;
    for (int i = 0; i < filter_size; i++)
      for (int j = 0; j < filter_size; j++) {
        const float x = (float)(i - hs), y = (float)(j - hs);
        f.w[i * filter_size + j] = expf(-(x * x + y * y) / (2 * filter_sigma * filter_sigma));
        sum += f.w[i * filter_size + j];
      }
    for (int i = 0; i < filter_size * filter_size; i++) f.w[i] /= sum;
  }
  const float c1 = powf(k1 * max_val, 2.f), c2 = powf(k2 * max_val, 2.f);  // :714-715
  const int span = kSsimTile + filter_size - 1;
  const size_t smem = (size_t)2 * span * span * 3 * sizeof(float);
  k_ssim<<<dim3((unsigned)cdiv(W, kSsimTile), (unsigned)cdiv(H, kSsimTile)), kSsimTile * kSsimTile, smem, st>>>(a, b, W, H, filter_size, f, c1, c2,
                                                                                                                 map_out, block_sums);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_generate_rays(const float* c2w12_host, float focal, int W, int H, float near, float far, int edge_mode, long first, long n,
                         float* o, float* d, float* radii, float* nears, float* fars, cudaStream_t st) {
  if (n <= 0) return 0;
  Pose pose;
  for (int i = 0; i < 12; i++) pose.c[i] = c2w12_host[i];
  k_generate_rays<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(pose, focal, W, H, near, far, edge_mode, first, n, o, d, radii, nears, fars);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_draw_indices(uint64_t seed, uint32_t slot0, uint32_t step, long n, int R, long* idx, cudaStream_t st) {
  k_draw_indices<<<(unsigned)cdiv(R, 256), 256, 0, st>>>(seed, slot0, step, n, R, idx);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_gather_batch(const float* records, long n, const long* idx, uint64_t seed, uint32_t slot0, uint32_t step, int R, float* o,
                        float* d, float* radii, float* nears, float* fars, float* lm, float* pix, cudaStream_t st) {
  k_gather_batch<<<(unsigned)cdiv(R, 128), 128, 0, st>>>(reinterpret_cast<const float4*>(records), n, idx, seed, slot0, step, R, o, d, radii,
                                                         nears, fars, lm, pix);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_sq_err(const float* a, const float* b, long n, double* out, cudaStream_t st) {
  NERF_CUDA(cudaMemsetAsync(out, 0, sizeof(double), st));
  long blocks = cdiv(n, 256 * 8);
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  k_sq_err<<<(unsigned)blocks, 256, 0, st>>>(a, b, n, out);
  NERF_CHECK_LAUNCH();
  return 0;
}

}  // namespace nerf

// gemm_tc.cu — the MLP layers as tcgen05 (5th-gen tensor core) GEMMs for sm_100a.
//
// Replaces the inner loops of get_neuron_output* (.cu:36-90, one dot product per thread from global memory) and
// backpropagate_neuron* (.cu:91-182, two global float atomics per multiply).
//
// One CTA computes one 128 x BN output tile:
//   warp 4 (one elected lane)  TMA producer: cp.async.bulk.tensor boxes of 64 bf16 x rows, 128-byte swizzle,
//                              into an n_stages-deep shared-memory ring guarded by full/empty mbarriers
//   warp 5 (one elected lane)  MMA issuer: tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16), 4 per stage,
//                              accumulating in TMEM (BN fp32 columns); tcgen05.commit frees the stage / signals done
//   warps 0-3                  epilogue: tcgen05.ld (32 lanes x 32 columns per warp), bias + activation / ReLU mask
//                              + rank-1 term / raw fp32 partial, bf16 hi(+lo) plane stores
// 192 threads, <= 100 KB shared memory and 256 TMEM columns per CTA so that two CTAs share an SM: one CTA's
// epilogue overlaps the other's main loop without a persistent scheduler.
//
// Operand layouts (sm100.cuh): K-major tiles for forward/dgrad (activations [M,K] and weights [N,K] are both
// reduction-contiguous), MN-major tiles for wgrad (dW = dZ^T X: both operands are read "transposed" straight from
// the row-major planes — no transposed copies of activations are ever made).
//
// fp32-accurate mode (NERF_PRECISION_FP32_TC): x = hi + lo with hi = bf16(x), lo = bf16(x - hi); the product keeps
// hi*hi + hi*lo + lo*hi (error ~2^-17 per product), expressed as three k-blocks per K slice.
#include <cuda.h>

#include <mutex>

#include "gemm_tc.cuh"
#include "sm100.cuh"

namespace nerf {
namespace {

using namespace sm100;

constexpr int kThreads = 192;
constexpr int kMaxStages = 6;
constexpr int kABytes = 128 * 128;  // 128 rows x 64 bf16 (K-major) or 2 boxes of 64 k-rows x 64 bf16 (MN-major)

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <bool MN_MAJOR>
__global__ void __launch_bounds__(kThreads, 2) k_tc_gemm(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], done_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  const int stage_bytes = kABytes + BN * 128;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  // tile coordinates and k-block count
  long row0;   // K-major: first output row (sample); MN-major: first output row (= column of the A planes)
  int col0;    // first output column
  int n_kb;    // pipeline iterations
  long red0 = 0;
  if (!MN_MAJOR) {
    row0 = (long)blockIdx.x * 128;
    col0 = blockIdx.y * BN;
    n_kb = p.n_kb;
  } else {
    row0 = (long)blockIdx.x * 128;
    col0 = blockIdx.y * BN;
    red0 = (long)blockIdx.z * p.split_len;
    long red1 = red0 + p.split_len;
    if (red1 > p.red_len) red1 = p.red_len;
    const long nblk = red1 > red0 ? (red1 - red0 + 63) / 64 : 0;
    n_kb = (int)nblk * p.n_pass;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc<256>(&tmem_base_smem);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 8; i++) prefetch_tmap(&p.maps[i]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < n_kb; i++) {
        const int s = i % p.n_stages;
        const uint32_t ph = (i / p.n_stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = smem + (size_t)s * stage_bytes;
        uint8_t* b_dst = a_dst + kABytes;
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
        if (!MN_MAJOR) {
          const TcParams::KB kb = p.kb[i];
          tma_load_2d(a_dst, &p.maps[kb.a], kb.a_col, (int)row0, &full_bar[s]);  // box {64, 128}
          tma_load_2d(b_dst, &p.maps[kb.b], kb.b_col, col0, &full_bar[s]);       // box {64, BN}
        } else {
          const int blk = i / p.n_pass, ps = i % p.n_pass;
          const int r = (int)(red0 + (long)blk * 64);
          const CUtensorMap* ma = &p.maps[p.pass_a[ps]];
          const CUtensorMap* mb = &p.maps[p.pass_b[ps]];
          tma_load_2d(a_dst, ma, p.a_col0 + (int)row0, r, &full_bar[s]);  // boxes {64 cols, 64 rows} = 8 KB each
          tma_load_2d(a_dst + 8192, ma, p.a_col0 + (int)row0 + 64, r, &full_bar[s]);
          for (int c = 0; c < BN; c += 64) tma_load_2d(b_dst + (c >> 6) * 8192, mb, col0 + c, r, &full_bar[s]);
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, BN, MN_MAJOR, MN_MAJOR);
      for (int i = 0; i < n_kb; i++) {
        const int s = i % p.n_stages;
        const uint32_t ph = (i / p.n_stages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after_sync();
        const uint32_t a_base = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b_base = a_base + kABytes;
#pragma unroll
        for (int k = 0; k < 4; k++) {  // 4 x (K = 16) per 64-wide k-block
          uint64_t da, db;
          if (!MN_MAJOR) {
            da = make_smem_desc(a_base + k * 32, 16, 1024);
            db = make_smem_desc(b_base + k * 32, 16, 1024);
          } else {
            da = make_smem_desc(a_base + k * 2048, 8192, 1024);
            db = make_smem_desc(b_base + k * 2048, 8192, 1024);
          }
          umma_bf16(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the stage once these MMAs have read it
      }
      umma_commit(&done_bar);  // accumulator complete
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-3: TMEM lanes 32w..32w+31)
    const long row = row0 + threadIdx.x;
    if (n_kb > 0) {
      mbar_wait(&done_bar, 0);
      tc_fence_after_sync();
    }
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      if (n_kb > 0) {
        tmem_ld_32x32(t_lane + c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; j++) r[j] = 0u;
      }
      const int col = col0 + c0;
      if (p.epi == 2) {  // raw fp32 partial: out[split][row][col]
        if (row < p.rows_valid) {
          float* dst = p.out_f32 + (long)blockIdx.z * p.split_stride + row * p.ld_f32 + col;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col + j + 3 < p.n_valid) {
              *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            } else {
              for (int q = 0; q < 4; q++)
                if (col + j + q < p.n_valid) dst[j + q] = __uint_as_float(r[j + q]);
            }
          }
        }
        continue;
      }
      if (row >= p.M || col >= p.n_valid) continue;
      float v[32];
      if (p.epi == 0) {  // Z = acc + b (.cu:45), Y = act(Z)
#pragma unroll
        for (int j = 0; j < 32; j++) {
          float z = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + col + j) : 0.f);
          v[j] = p.act == ACT_RELU ? fmaxf(z, 0.f) : z;
        }
      } else {  // dgrad: (+ r1[m] v1[k]) then the ReLU mask of the layer below (.cu:99)
        const float ri = p.r1 ? __ldg(p.r1 + row) : 0.f;
        uint32_t mk[16];
        if (p.mask) {
          const uint4* mp = reinterpret_cast<const uint4*>(p.mask + row * p.ld_mask + col);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const uint4 t = __ldg(mp + q);
            mk[q * 4] = t.x; mk[q * 4 + 1] = t.y; mk[q * 4 + 2] = t.z; mk[q * 4 + 3] = t.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; j++) {
          float x = __uint_as_float(r[j]);
          if (p.r1) x = fmaf(ri, __ldg(p.v1 + col + j), x);
          if (p.mask) {
            const uint32_t w = mk[j >> 1];
            const uint16_t h = (j & 1) ? (uint16_t)(w >> 16) : (uint16_t)(w & 0xFFFF);
            // bf16 > 0  <=>  sign bit clear and magnitude non-zero
            x = ((h & 0x8000u) == 0 && (h & 0x7FFFu) != 0) ? x : 0.f;
          }
          v[j] = x;
        }
      }
      // bf16 split planes: hi = bf16(v), lo = bf16(v - hi)
      uint4* oh = reinterpret_cast<uint4*>(p.out_hi + row * p.ld_out + col);
#pragma unroll
      for (int q = 0; q < 4; q++)
        oh[q] = make_uint4(pack_bf16(v[q * 8], v[q * 8 + 1]), pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                           pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
      if (p.out_lo) {
        uint4* ol = reinterpret_cast<uint4*>(p.out_lo + row * p.ld_out + col);
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] -= __bfloat162float(__float2bfloat16_rn(v[j]));
#pragma unroll
        for (int q = 0; q < 4; q++)
          ol[q] = make_uint4(pack_bf16(v[q * 8], v[q * 8 + 1]), pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                             pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc<256>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ plane helpers

__global__ void k_f32_to_planes(const float* __restrict__ src, int sp, long rows, int cols, __nv_bfloat16* __restrict__ hi,
                                __nv_bfloat16* __restrict__ lo, int dp, int dcols, int transpose, long drows_t) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (!transpose) {
    if (idx >= rows * dcols) return;
    const long r = idx / dcols;
    const int c = (int)(idx % dcols);
    const float x = c < cols ? src[r * sp + c] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[r * dp + c] = h;
    if (lo) lo[r * dp + c] = __float2bfloat16_rn(x - __bfloat162float(h));
  } else {  // dst[c, r] = src[r, c] for c < drows_t
    if (idx >= drows_t * dcols) return;
    const long c = idx / dcols;      // dst row = src column
    const int r = (int)(idx % dcols);  // dst column = src row
    const float x = (r < rows && c < cols) ? src[(long)r * sp + c] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[c * dp + r] = h;
    if (lo) lo[c * dp + r] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

__device__ __forceinline__ float plane_val(const __nv_bfloat16* h, const __nv_bfloat16* l, long i) {
  float x = __bfloat162float(h[i]);
  if (l) x += __bfloat162float(l[i]);
  return x;
}

__global__ void __launch_bounds__(256)
k_thin_fwd_planes(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl, int ldx,
                  const float* __restrict__ W, const float* __restrict__ b, float* __restrict__ Y, long M, int N, int K) {
  const long m = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = lane * 2; k < K; k += 64) {  // 2 bf16 per lane per step: 128 B per warp, coalesced
    const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(xh + m * ldx + k);
    float x0 = __low2float(h2), x1 = __high2float(h2);
    if (xl) {
      const __nv_bfloat162 l2 = *reinterpret_cast<const __nv_bfloat162*>(xl + m * ldx + k);
      x0 += __low2float(l2); x1 += __high2float(l2);
    }
#pragma unroll
    for (int n = 0; n < 4; n++)
      if (n < N) acc[n] = fmaf(x1, __ldg(W + n * K + k + 1), fmaf(x0, __ldg(W + n * K + k), acc[n]));
  }
#pragma unroll
  for (int n = 0; n < 4; n++) {
    float v = acc[n];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && n < N) Y[m * N + n] = v + (b ? b[n] : 0.f);
  }
}

__global__ void k_thin_dgrad_planes(const float* __restrict__ dZ, const float* __restrict__ W, long M, int N, int K,
                                    const __nv_bfloat16* __restrict__ mask, int ldm, __nv_bfloat16* __restrict__ oh,
                                    __nv_bfloat16* __restrict__ ol, int ldo) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * K) return;
  const long m = idx / K;
  const int k = (int)(idx % K);
  float v = 0.f;
  for (int n = 0; n < N; n++) v = fmaf(__ldg(dZ + m * N + n), __ldg(W + n * K + k), v);
  if (mask) v = __bfloat162float(mask[m * ldm + k]) > 0.f ? v : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  oh[m * ldo + k] = h;
  if (ol) ol[m * ldo + k] = __float2bfloat16_rn(v - __bfloat162float(h));
}

__global__ void __launch_bounds__(256)
k_thin_wgrad_planes_partial(const float* __restrict__ dZ, const __nv_bfloat16* __restrict__ xh,
                            const __nv_bfloat16* __restrict__ xl, int ldx, long M, int N, int K, long chunk,
                            float* __restrict__ part, float* __restrict__ partb) {
  const long m0 = (long)blockIdx.x * chunk, m1 = min(M, m0 + chunk);
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long m = m0; m < m1; m++) {
      const float x = plane_val(xh, xl, m * ldx + k);
#pragma unroll
      for (int n = 0; n < 4; n++)
        if (n < N) acc[n] = fmaf(__ldg(dZ + m * N + n), x, acc[n]);
    }
    for (int n = 0; n < N; n++) part[(long)blockIdx.x * N * K + n * K + k] = acc[n];
  }
  if (threadIdx.x < N) {
    float s = 0.f;
    for (long m = m0; m < m1; m++) s += __ldg(dZ + m * N + threadIdx.x);
    partb[(long)blockIdx.x * N + threadIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256)
k_colsum_planes_partial(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ l, int ld, long M, int N,
                        long chunk, float* __restrict__ part) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const long m0 = (long)blockIdx.y * chunk, m1 = min(M, m0 + chunk);
  float s = 0.f;
  if (n < N)
    for (long m = m0 + wy; m < m1; m += 8) s += plane_val(h, l, m * ld + n);
  red[wy][lane] = s;
  __syncthreads();
  if (wy == 0 && n < N) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) tot += red[k][lane];
    part[(long)blockIdx.y * N + n] = tot;
  }
}

__global__ void k_reduce_partials_tc(const float* __restrict__ ws, int splits, long stride, int rows, int cols, int ldw,
                                     float* __restrict__ out, int ldo, int coff) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)rows * cols) return;
  const int i = (int)(idx / cols), j = (int)(idx % cols);
  float s = 0.f;
  for (int z = 0; z < splits; z++) s += ws[(long)z * stride + (long)i * ldw + j];  // fixed order: deterministic
  out[(long)i * ldo + coff + j] += s;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace

int tc_make_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return 100002; }
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {pitch_elems * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu pitch=%llu box_rows=%u", (int)r, base,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_elems, box_rows);
    return 100001;
  }
  return 0;
}

int tc_smem_bytes(int BN, int n_stages) { return n_stages * (kABytes + BN * 128) + 1024; }

int tc_pick_stages(int BN, int n_kblocks) {
  // two CTAs per SM: <= ~110 KB each
  int s = (110 * 1024 - 1024) / (kABytes + BN * 128);
  if (s > kMaxStages) s = kMaxStages;
  if (s > n_kblocks && n_kblocks > 0) s = n_kblocks;
  return s < 1 ? 1 : s;
}

int tc_launch(const TcParams& p, bool mn_major, dim3 grid, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(k_tc_gemm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(k_tc_gemm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  if (attr_err != cudaSuccess) { set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err)); return (int)attr_err; }
  if (p.BN % 16 || p.BN < 16 || p.BN > 256 || p.n_stages < 1 || p.n_stages > kMaxStages) { set_error("tc_launch: bad BN=%d stages=%d", p.BN, p.n_stages); return 100001; }
  const size_t smem = (size_t)tc_smem_bytes(p.BN, p.n_stages);
  if (mn_major) k_tc_gemm<true><<<grid, kThreads, smem, st>>>(p);
  else k_tc_gemm<false><<<grid, kThreads, smem, st>>>(p);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_f32_to_planes(const float* src, int src_pitch, long rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo,
                         int dst_pitch, int dst_cols, bool transpose, long dst_rows_t, cudaStream_t st) {
  const long n = transpose ? dst_rows_t * dst_cols : rows * dst_cols;
  if (n <= 0) return 0;
  k_f32_to_planes<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(src, src_pitch, rows, cols, hi, lo, dst_pitch, dst_cols,
                                                         transpose ? 1 : 0, dst_rows_t);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_thin_fwd_planes(const __nv_bfloat16* xh, const __nv_bfloat16* xl, int ldx, const float* W, const float* b,
                           float* Y, long M, int N, int K, cudaStream_t st) {
  if (N > 4 || (K & 1)) { set_error("thin_fwd_planes: N=%d K=%d", N, K); return 100001; }
  k_thin_fwd_planes<<<(unsigned)cdiv(M, 8), 256, 0, st>>>(xh, xl, ldx, W, b, Y, M, N, K);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_thin_dgrad_planes(const float* dZ, const float* W, long M, int N, int K, const __nv_bfloat16* mask, int ld_mask,
                             __nv_bfloat16* oh, __nv_bfloat16* ol, int ldo, cudaStream_t st) {
  k_thin_dgrad_planes<<<(unsigned)cdiv(M * K, 256), 256, 0, st>>>(dZ, W, M, N, K, mask, ld_mask, oh, ol, ldo);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_reduce_partials(const float* ws, int splits, long split_stride, int rows, int cols, int ldw, float* out, int ldo,
                           int coff, cudaStream_t st) {
  const long n = (long)rows * cols;
  k_reduce_partials_tc<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(ws, splits, split_stride, rows, cols, ldw, out, ldo, coff);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_thin_wgrad_planes(const float* dZ, const __nv_bfloat16* xh, const __nv_bfloat16* xl, int ldx, float* dW, float* db,
                             long M, int N, int K, float* workspace, cudaStream_t st) {
  if (N > 4) { set_error("thin_wgrad_planes: N=%d", N); return 100001; }
  const long chunk = 1024;
  const int chunks = (int)cdiv(M, chunk);
  float* part = workspace;
  float* partb = workspace + (size_t)chunks * N * K;
  k_thin_wgrad_planes_partial<<<chunks, 256, 0, st>>>(dZ, xh, xl, ldx, M, N, K, chunk, part, partb);
  NERF_CHECK_LAUNCH();
  NERF_TRY(launch_reduce_partials(part, chunks, (long)N * K, N, K, K, dW, K, 0, st));
  if (db) NERF_TRY(launch_reduce_partials(partb, chunks, N, 1, N, N, db, N, 0, st));
  return 0;
}

int launch_colsum_planes(const __nv_bfloat16* h, const __nv_bfloat16* l, int ld, long M, int N, float* db, float* workspace,
                         cudaStream_t st) {
  const long chunk = 2048;
  const int chunks = (int)cdiv(M, chunk);
  k_colsum_planes_partial<<<dim3((unsigned)cdiv(N, 32), (unsigned)chunks), 256, 0, st>>>(h, l, ld, M, N, chunk, workspace);
  NERF_CHECK_LAUNCH();
  return launch_reduce_partials(workspace, chunks, N, 1, N, N, db, N, 0, st);
}

}  // namespace nerf

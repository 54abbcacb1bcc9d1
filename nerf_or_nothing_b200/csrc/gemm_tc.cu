// gemm_tc.cu — the MLP layers as tcgen05 (5th-gen tensor core) GEMMs for sm_100a.
//
// Replaces the inner loops of get_neuron_output* (.cu:36-90, one dot product per thread from global memory) and
// backpropagate_neuron* (.cu:91-182, two global float atomics per multiply).
//
// Two kernels share the TMA / mbarrier / tcgen05 machinery of sm100.cuh:
//   k_tc_gemm_persist  forward and dgrad (K-major operands): persistent, warp-specialised, accumulator double-buffered
//                      in TMEM, per-warp TMA-store epilogue with bias/ReLU/bit-mask/head or rank-1/mask fused
//   k_tc_wgrad         dW = dZ^T X (MN-major operands) split over the samples, bias gradient via a ones tile
//
// Operand layouts (sm100.cuh): K-major tiles for forward/dgrad (activations [M,K] and weights [N,K] are both
// reduction-contiguous), MN-major tiles for wgrad (dW = dZ^T X: both operands are read "transposed" straight from
// the row-major planes — no transposed copies of activations are ever made).
//
// fp32-accurate mode (NERF_PRECISION_FP32_TC): x = hi + lo with hi = bf16(x), lo = bf16(x - hi); the product keeps
// hi*hi + hi*lo + lo*hi (error ~2^-17 per product), expressed as three k-blocks per K slice.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

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

// One slot of a ReduceJob: slot idx < rows*ceil(cols/4) sums 4 consecutive columns of one output row over the splits with
// 128-bit loads, later slots one bias column each.  The order is the one k_reduce_partials_tc uses — four interleaved
// split groups z = g, g+4, ... combined as (g0 + g1) + (g2 + g3) — so a job gives bitwise the same result whether it runs
// stand-alone or inside the next wgrad launch.  Needs ldw % 4 == 0 and stride % 4 == 0.
__device__ __forceinline__ void f4_add(float4& s, const float4 v) { s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
__device__ __forceinline__ void reduce_job_slot(const ReduceJob& j, long idx) {
  const int c4 = (j.cols + 3) >> 2;
  const long n1 = (long)j.rows * c4;
  const float mul = j.mul ? __ldg(j.mul) : 1.0f;  // powers of two: exact
  const float mul2 = j.mul2 ? __ldg(j.mul2) : 1.0f;
  if (idx < n1) {
    const int i = (int)(idx / c4), c = (int)(idx % c4) * 4;
    const float* src = j.ws + (long)i * j.ldw + c;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
    int z = 0;
    for (; z + 4 <= j.splits; z += 4) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(src + (long)z * j.stride));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(src + (long)(z + 1) * j.stride));
      const float4 v2 = __ldg(reinterpret_cast<const float4*>(src + (long)(z + 2) * j.stride));
      const float4 v3 = __ldg(reinterpret_cast<const float4*>(src + (long)(z + 3) * j.stride));
      f4_add(s0, v0); f4_add(s1, v1); f4_add(s2, v2); f4_add(s3, v3);
    }
    if (z < j.splits) f4_add(s0, __ldg(reinterpret_cast<const float4*>(src + (long)z * j.stride)));
    if (z + 1 < j.splits) f4_add(s1, __ldg(reinterpret_cast<const float4*>(src + (long)(z + 1) * j.stride)));
    if (z + 2 < j.splits) f4_add(s2, __ldg(reinterpret_cast<const float4*>(src + (long)(z + 2) * j.stride)));
    float* o = j.out + (long)i * j.ldo + j.coff + c;
    o[0] += ((s0.x + s1.x) + (s2.x + s3.x)) * mul;
    if (c + 1 < j.cols) o[1] += ((s0.y + s1.y) + (s2.y + s3.y)) * mul;
    if (c + 2 < j.cols) o[2] += ((s0.z + s1.z) + (s2.z + s3.z)) * mul;
    if (c + 3 < j.cols) o[3] += ((s0.w + s1.w) + (s2.w + s3.w)) * mul;
  } else if (idx - n1 < j.n2) {
    const int jb = (int)(idx - n1);
    const float* src = j.ws2 + jb;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int z = 0;
    for (; z + 4 <= j.splits; z += 4) {
      s0 += src[(long)z * j.stride2]; s1 += src[(long)(z + 1) * j.stride2];
      s2 += src[(long)(z + 2) * j.stride2]; s3 += src[(long)(z + 3) * j.stride2];
    }
    if (z < j.splits) s0 += src[(long)z * j.stride2];
    if (z + 1 < j.splits) s1 += src[(long)(z + 1) * j.stride2];
    if (z + 2 < j.splits) s2 += src[(long)(z + 2) * j.stride2];
    j.out2[jb] += ((s0 + s1) + (s2 + s3)) * mul2;
  }
}

__global__ void __launch_bounds__(128) k_reduce_job(const ReduceJob j) {
  reduce_job_slot(j, (long)blockIdx.x * 128 + threadIdx.x);
}

// MN-major kernel (wgrad): one CTA = one 128 x BN tile of dW over a slice of the samples.
//   warp 4 lane 0  TMA producer   boxes {64 cols, 64 rows} of the dZ / X planes, 128B-swizzled, n_stages-deep ring
//   warp 5 lane 0  MMA issuer     tcgen05.mma M=128, N=BN, K=16, both operands MN-major (a_major = b_major = 1);
//                                 + an N=16 MMA of the same dZ tile against a constant tile of ones: TMEM columns
//                                 256..271 accumulate colsum(dZ) = the bias gradient for free
//   warps 0-3      epilogue       tcgen05.ld -> fp32 partial tile [split][row][col] (+ bias column)
// p.f16_ops: the planes hold fp16 (the fp32-accurate mode's wgrad operands, see mlp_tc.cu) — same tiles, same descriptors,
// only the instruction descriptor's operand formats and the tile of ones differ.
// One CTA per SM (192 KB ring, 512 TMEM columns); the kernel runs at the HBM roofline of reading dZ and X once.
__global__ void __launch_bounds__(kThreads, 1) k_tc_wgrad(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], done_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  // one ring stage = one 64-sample k-block: the dZ tile and the X tile of every plane, each fetched ONCE and shared by the
  // passes of the split product (hi*hi, hi*lo, lo*hi): [A plane 0][A plane 1][B plane 0][B plane 1]
  const int np = p.n_pass > 1 ? 2 : 1;
  const int b_bytes = BN * 128;
  const int stage_bytes = np * (kABytes + b_bytes);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ones = smem + (size_t)p.n_stages * stage_bytes;  // 8 KB of bf16 1.0

  const long row0 = (long)blockIdx.x * 128;  // first output row (= column of the dZ planes)
  const int col0 = blockIdx.y * BN;          // first output column (= column of the X planes)
  const long red0 = (long)blockIdx.z * p.split_len;
  long red1 = red0 + p.split_len;
  if (red1 > p.red_len) red1 = p.red_len;
  const long nblk = red1 > red0 ? (red1 - red0 + 63) / 64 : 0;
  const int n_kb = (int)nblk;
  const bool want_bias = p.bias_out != nullptr && blockIdx.y == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2048; i += kThreads) reinterpret_cast<uint32_t*>(ones)[i] = p.f16_ops ? 0x3C003C00u : 0x3F803F80u;  // 1.0
  fence_proxy_async();
  if (warp == 4) {
    tmem_alloc<512>(&tmem_base_smem);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 6; i++) prefetch_tmap(&p.maps[i]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 4) {
    if (lane == 0) {  // ---------------------------------------------------------------- TMA producer
      for (int i = 0; i < n_kb; i++) {
        const int s = i % p.n_stages;
        const uint32_t ph = (i / p.n_stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = smem + (size_t)s * stage_bytes;
        uint8_t* b_dst = a_dst + np * kABytes;
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
        const int r = (int)(red0 + (long)i * 64);
        for (int pl = 0; pl < np; pl++) {  // maps 0/1 = dZ hi/lo, 4/5 = X hi/lo
          tma_load_2d(a_dst + pl * kABytes, &p.maps[pl], p.a_col0 + (int)row0, r, &full_bar[s]);
          tma_load_2d(a_dst + pl * kABytes + 8192, &p.maps[pl], p.a_col0 + (int)row0 + 64, r, &full_bar[s]);
          for (int c = 0; c < BN; c += 64) tma_load_2d(b_dst + pl * b_bytes + (c >> 6) * 8192, &p.maps[4 + pl], col0 + c, r, &full_bar[s]);
        }
      }
    }
  } else if (warp == 5) {
    {  // ---------------------------------------------------------------- MMA issuer (uniform warp, one elected lane issues)
      const bool leader = elect_one();
      const uint32_t idesc = p.f16_ops ? make_idesc_f16(128, BN, true, true) : make_idesc_bf16(128, BN, true, true);
      const uint32_t idesc_bias = p.f16_ops ? make_idesc_f16(128, 16, true, true) : make_idesc_bf16(128, 16, true, true);
      const uint64_t desc0 = make_smem_desc(0, 8192, 1024);  // + (shared address >> 4)
      const uint32_t ones_base = smem_u32(ones), smem_base = smem_u32(smem);
      for (int i = 0; i < n_kb; i++) {
        const int s = i % p.n_stages;
        const uint32_t ph = (i / p.n_stages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after_sync();
        const uint32_t a_base = smem_base + (uint32_t)s * (uint32_t)stage_bytes;
        const uint32_t b_base = a_base + (uint32_t)(np * kABytes);
        if (leader) {
          for (int q = 0; q < p.n_pass; q++) {  // (A plane, B plane) of each term of the split product
            const uint64_t da = desc0 + ((a_base + (uint32_t)p.pass_a[q] * kABytes) >> 4);
            const uint64_t db = desc0 + ((b_base + (uint32_t)(p.pass_b[q] - 4) * (uint32_t)b_bytes) >> 4);
#pragma unroll
            for (int k = 0; k < 4; k++)  // 16 reduction rows per MMA = 2 KB of each 64-column box
              umma_bf16(tmem_base, da + k * 128, db + k * 128, idesc, (i > 0 || q > 0 || k > 0) ? 1u : 0u);
          }
          if (want_bias) {  // colsum(dZ hi + dZ lo): every dZ plane once against the tile of ones
            const uint64_t d1 = desc0 + (ones_base >> 4);
            for (int pl = 0; pl < np; pl++) {
              const uint64_t da = desc0 + ((a_base + (uint32_t)pl * kABytes) >> 4);
#pragma unroll
              for (int k = 0; k < 4; k++) umma_bf16(tmem_base + 256, da + k * 128, d1 + k * 128, idesc_bias, (i > 0 || pl > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[s]);
        }
      }
      if (leader) umma_commit(&done_bar);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-3: TMEM lanes 32w..32w+31)
    // idle until the last k-block has been multiplied: first reduce the previous launch's partial tiles
    if (p.red.ws != nullptr) {
      const long cta = ((long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      const long step = (long)gridDim.x * gridDim.y * gridDim.z * 128;
      const long total = (long)p.red.rows * ((p.red.cols + 3) >> 2) + p.red.n2;
      for (long idx = cta * 128 + threadIdx.x; idx < total; idx += step) reduce_job_slot(p.red, idx);
    }
    const long row = row0 + threadIdx.x;
    if (n_kb > 0) {
      mbar_wait(&done_bar, 0);
      tc_fence_after_sync();
    }
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      if (n_kb > 0) { tmem_ld_32x32(t_lane + c0, r); tmem_ld_wait(); }
      else {
#pragma unroll
        for (int j = 0; j < 32; j++) r[j] = 0u;
      }
      const int col = col0 + c0;
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
    }
    if (want_bias) {
      uint32_t r[32];
      if (n_kb > 0) { tmem_ld_32x32(t_lane + 256, r); tmem_ld_wait(); }
      else r[0] = 0u;
      if (row < p.rows_valid) p.bias_out[(long)blockIdx.z * p.bias_split_stride + row] = __uint_as_float(r[0]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tmem_base);
}

// Persistent K-major variant (forward / dgrad): one CTA per SM walks row tiles tile = blockIdx.x + i * gridDim.x.
// The TMA producer streams k-blocks of the NEXT tile while the epilogue of the current one runs: the accumulator is
// double-buffered in TMEM (2 x 256 columns, tfull/tempty mbarriers), the operand ring is 3-4 stages deep and never
// drains at tile boundaries.  Each epilogue warp owns its 32 rows end to end — TMEM load, math, restaging into its own
// 128B-swizzled 32-row boxes and its own TMA bulk stores — so the four warps never synchronise with each other; the
// per-column constants (bias / rank-1 vector / head weights) are staged once per CTA in shared memory.
constexpr int kConstFloats = 4 * 512;  // bias|v1 [512] + head weights [3][512]

__device__ __forceinline__ void store_planes_32(uint8_t* row_hi, uint8_t* row_lo, bool with_lo, int half, int swz, float (&v)[32]) {
  // hi = bf16(v) packed in pairs (one cvt per pair); lo = bf16(v - hi) with hi rebuilt from the packed bits (no 2nd cvt)
  uint32_t hw[16];
#pragma unroll
  for (int q = 0; q < 16; q++) hw[q] = pack_bf16(v[2 * q], v[2 * q + 1]);
#pragma unroll
  for (int q = 0; q < 4; q++)
    *reinterpret_cast<uint4*>(row_hi + (((half * 4 + q) ^ swz) << 4)) = make_uint4(hw[4 * q], hw[4 * q + 1], hw[4 * q + 2], hw[4 * q + 3]);
  if (with_lo) {
    uint32_t lw[16];
#pragma unroll
    for (int q = 0; q < 16; q++)
      lw[q] = pack_bf16(v[2 * q] - __uint_as_float(hw[q] << 16), v[2 * q + 1] - __uint_as_float(hw[q] & 0xFFFF0000u));
#pragma unroll
    for (int q = 0; q < 4; q++)
      *reinterpret_cast<uint4*>(row_lo + (((half * 4 + q) ^ swz) << 4)) = make_uint4(lw[4 * q], lw[4 * q + 1], lw[4 * q + 2], lw[4 * q + 3]);
  }
}

__global__ void __launch_bounds__(kThreads, 1) k_tc_gemm_persist(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_const[kConstFloats];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  const int stage_bytes = kABytes + BN * 128;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stg = smem + (size_t)p.n_stages * stage_bytes;  // per warp: 2 slot pairs x (hi 4 KB + lo 4 KB)

  const int n_row_tiles = (int)((p.M + 127) / 128);
  const int n_col_tiles = (p.n_valid + BN - 1) / BN;
  const int n_tiles = n_row_tiles * n_col_tiles;
  const int n_kb = p.n_kb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
    fence_barrier_init();
  }
  {  // per-column constants -> shared memory (zero beyond n_valid)
    const float* vec = p.epi == 0 ? p.bias : p.v1;
    for (int i = threadIdx.x; i < 512; i += kThreads) s_const[i] = (vec && i < p.n_valid) ? __ldg(vec + i) : 0.f;
    for (int i = threadIdx.x; i < 3 * 512; i += kThreads) {
      const int n = i / 512, c = i % 512;
      s_const[512 + i] = (p.epi == 0 && n < p.head_n && c < p.n_valid) ? __ldg(p.head_w + (long)n * p.n_valid + c) : 0.f;
    }
  }
  if (warp == 4) {
    tmem_alloc<512>(&tmem_base_smem);
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
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = (tile / n_col_tiles) * 128, col0 = (tile % n_col_tiles) * BN;
        for (int i = 0; i < n_kb; i++, it++) {
          const int s = it % p.n_stages;
          const uint32_t ph = (it / p.n_stages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* a_dst = smem + (size_t)s * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
          const TcParams::KB kb = p.kb[i];
          tma_load_2d(a_dst, &p.maps[kb.a], kb.a_col, row0, &full_bar[s]);            // box {64, 128}
          tma_load_2d(a_dst + kABytes, &p.maps[kb.b], kb.b_col, col0, &full_bar[s]);  // box {64, BN}
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer (uniform warp, one elected lane issues)
    {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc_bf16(128, BN, false, false);
      const uint64_t desc0 = make_smem_desc(0, 16, 1024);  // + (shared address >> 4)
      const uint32_t smem_base = smem_u32(smem);
      uint32_t it = 0, t = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, t++) {
        const uint32_t b = t & 1;
        mbar_wait(&tempty_bar[b], ((t >> 1) & 1) ^ 1);  // all four epilogue warps have drained this accumulator
        tc_fence_after_sync();
        const uint32_t acc = tmem_base + b * 256;
        for (int i = 0; i < n_kb; i++, it++) {
          const int s = it % p.n_stages;
          const uint32_t ph = (it / p.n_stages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after_sync();
          const uint32_t a_base = smem_base + (uint32_t)s * (uint32_t)stage_bytes;
          const uint64_t da = desc0 + (a_base >> 4), db = desc0 + ((a_base + kABytes) >> 4);
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; k++) umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[s]);
          }
        }
        if (leader) umma_commit(&tfull_bar[b]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue: warp w owns rows 32w..32w+31
    const int n_groups = BN / 64;
    const int swz = lane & 7;
    const bool with_lo = p.out_lo != nullptr;
    uint8_t* wstg = stg + (size_t)warp * 16384;
    uint32_t t = 0, gc = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, t++) {
      const int row0 = (tile / n_col_tiles) * 128, col0 = (tile % n_col_tiles) * BN;
      const long row = (long)row0 + threadIdx.x;
      const bool row_ok = row < p.M;
      const uint32_t b = t & 1;
      mbar_wait(&tfull_bar[b], (t >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t t_lane = tmem_base + b * 256 + ((uint32_t)(warp * 32) << 16);
      const float ri = (p.epi == 1 && p.r1 && row_ok) ? __ldg(p.r1 + row) : 0.f;
      float head_acc[3] = {0.f, 0.f, 0.f};
      for (int g = 0; g < n_groups; g++, gc++) {
        uint8_t* slot_hi = wstg + (size_t)(gc & 1) * 8192;
        uint8_t* slot_lo = slot_hi + 4096;
        if (gc >= 2) {
          if (lane == 0) tma_store_wait_read<1>();  // this warp's group that used the slot pair has been read out
          __syncwarp();
        }
        uint8_t* row_hi = slot_hi + lane * 128;
        uint8_t* row_lo = slot_lo + lane * 128;
#pragma unroll
        for (int half = 0; half < 2; half++) {
          const int c0 = g * 64 + half * 32;
          const int col = col0 + c0;  // multiple of 32; n_valid is a multiple of 32
          uint32_t r[32];
          tmem_ld_32x32(t_lane + c0, r);
          tmem_ld_wait();
          float v[32];
          const float4* cv = reinterpret_cast<const float4*>(s_const + col);
          if (p.epi == 0) {  // Z = acc + b (.cu:45), Y = relu(Z)
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const float4 bb = cv[q];
              v[4 * q] = fmaxf(__uint_as_float(r[4 * q]) + bb.x, 0.f);
              v[4 * q + 1] = fmaxf(__uint_as_float(r[4 * q + 1]) + bb.y, 0.f);
              v[4 * q + 2] = fmaxf(__uint_as_float(r[4 * q + 2]) + bb.z, 0.f);
              v[4 * q + 3] = fmaxf(__uint_as_float(r[4 * q + 3]) + bb.w, 0.f);
            }
            if (p.bits_out) {  // v >= 0, so v > 0 <=> its bit pattern is non-zero
              uint32_t bits = 0u;
#pragma unroll
              for (int j = 0; j < 32; j++) bits |= (__float_as_uint(v[j]) != 0u ? 1u : 0u) << j;
              if (row_ok) p.bits_out[row * p.ld_bits + (col >> 5)] = bits;
            }
#pragma unroll
            for (int n = 0; n < 3; n++) {
              if (n < p.head_n) {
                const float4* hv = reinterpret_cast<const float4*>(s_const + 512 + n * 512 + col);
                float a = head_acc[n];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                  const float4 w = hv[q];
                  a = fmaf(v[4 * q], w.x, a); a = fmaf(v[4 * q + 1], w.y, a);
                  a = fmaf(v[4 * q + 2], w.z, a); a = fmaf(v[4 * q + 3], w.w, a);
                }
                head_acc[n] = a;
              }
            }
          } else {  // dgrad: (+ r1[m] v1[k]) then the ReLU mask of the layer below (.cu:99), read as a bit plane
            const uint32_t mbits = p.mask_bits ? (row_ok ? __ldg(p.mask_bits + row * p.ld_bits + (col >> 5)) : 0u) : 0xFFFFFFFFu;
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const float4 vv = cv[q];
              const float w4[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
              for (int e = 0; e < 4; e++) {
                const int j = 4 * q + e;
                const float x = fmaf(ri, w4[e], __uint_as_float(r[j]));  // ri = 0 when there is no rank-1 term
                v[j] = ((mbits >> j) & 1u) ? x : 0.f;
              }
            }
          }
          store_planes_32(row_hi, row_lo, with_lo, half, swz, v);
        }
        if (g == n_groups - 1) tc_fence_before_sync();  // this warp's last TMEM read of the tile is done
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (g == n_groups - 1) mbar_arrive(&tempty_bar[b]);
          tma_store_2d(&p.maps[6], slot_hi, col0 + g * 64, row0 + warp * 32);  // box {64, 32}
          if (with_lo) tma_store_2d(&p.maps[7], slot_lo, col0 + g * 64, row0 + warp * 32);
          tma_store_commit();
        }
      }
      if (p.epi == 0 && p.head_n > 0 && row_ok) {
#pragma unroll
        for (int n = 0; n < 3; n++)
          if (n < p.head_n) p.head_out[row * p.head_n + n] = head_acc[n] + (p.head_b ? __ldg(p.head_b + n) : 0.f);
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ plane helpers

__device__ __forceinline__ void f32_to_planes_elem(long idx, const float* __restrict__ src, int sp, long rows, int cols,
                                                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int dp, int dcols,
                                                   int transpose, long drows_t) {
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

__global__ void k_f32_to_planes(const float* __restrict__ src, int sp, long rows, int cols, __nv_bfloat16* __restrict__ hi,
                                __nv_bfloat16* __restrict__ lo, int dp, int dcols, int transpose, long drows_t) {
  f32_to_planes_elem((long)blockIdx.x * blockDim.x + threadIdx.x, src, sp, rows, cols, hi, lo, dp, dcols, transpose, drows_t);
}

// every weight plane (and transposed plane) of the network in ONE launch: blockIdx.y = job
__global__ void k_f32_to_planes_batch(const __grid_constant__ PlaneJobs jobs) {
  const PlaneJobs::Job& j = jobs.job[blockIdx.y];
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < j.n; idx += (long)gridDim.x * blockDim.x)
    f32_to_planes_elem(idx, j.src, j.sp, j.rows, j.cols, j.hi, j.lo, j.dp, j.dcols, j.transpose, j.drows_t);
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

// dX[m, k..k+7] = mask * sum_n dz[m,n] W[n,k].  A thread owns ONE 8-column chunk for the whole launch — its NN x 8 weights
// live in registers — and walks rows with a grid stride (K/8 consecutive lanes cover a row: 16-byte plane stores, a full 128-byte
// line per 8 lanes); the head gradients of a row are broadcast loads.  Pure write traffic: 4 B (hi + lo) per element.
template <int NN>
__global__ void __launch_bounds__(256)
k_thin_dgrad_planes(const float* __restrict__ dZ, const float* __restrict__ W, long M, int K, const uint32_t* __restrict__ mask_bits,
                    int ld_bits, __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol, int ldo, __half* __restrict__ o16,
                    const float* __restrict__ scale16) {
  const float s16 = o16 ? __ldg(scale16) : 1.0f;  // o16: a third plane fp16(dX * s16) for the fp16 wgrad GEMMs
  const int k8 = K >> 3;                      // chunks per row (blockDim.x is a multiple of it)
  const int k = (threadIdx.x % k8) * 8;
  const int rows_per_block = blockDim.x / k8;
  float w[NN][8];
#pragma unroll
  for (int n = 0; n < NN; n++) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + n * K + k)), w1 = __ldg(reinterpret_cast<const float4*>(W + n * K + k + 4));
    w[n][0] = w0.x; w[n][1] = w0.y; w[n][2] = w0.z; w[n][3] = w0.w; w[n][4] = w1.x; w[n][5] = w1.y; w[n][6] = w1.z; w[n][7] = w1.w;
  }
  for (long m = (long)blockIdx.x * rows_per_block + threadIdx.x / k8; m < M; m += (long)gridDim.x * rows_per_block) {
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int n = 0; n < NN; n++) {
      const float g = __ldg(dZ + m * NN + n);
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = fmaf(g, w[n][j], v[j]);
    }
    if (mask_bits) {
      const uint32_t bits = __ldg(mask_bits + m * ld_bits + (k >> 5)) >> (k & 31);
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = ((bits >> j) & 1u) ? v[j] : 0.f;
    }
    *reinterpret_cast<uint4*>(oh + m * ldo + k) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    if (o16)
      *reinterpret_cast<uint4*>(o16 + m * ldo + k) = make_uint4(pack_f16x2_sat(v[0] * s16, v[1] * s16), pack_f16x2_sat(v[2] * s16, v[3] * s16),
                                                                pack_f16x2_sat(v[4] * s16, v[5] * s16), pack_f16x2_sat(v[6] * s16, v[7] * s16));
    if (ol) {
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] -= __bfloat162float(__float2bfloat16_rn(v[j]));
      *reinterpret_cast<uint4*>(ol + m * ldo + k) =
          make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
  }
}

// Thin-head wgrad partials: dW[n, k] = sum_m dz[m, n] x[m, k] over this block's rows.  Each lane owns 8 consecutive
// columns (16-byte plane loads); a warp covers 32 / (K/8) rows per load instruction (two rows when K = 128), 4 loads in
// flight; row groups and the 8 warps are then summed through shuffles / shared memory.  part[block][n*K + k],
// partb[block][n].   K <= 256, K a multiple of 8.  NN = compile-time head width (1 or 3): no FMAs on absent heads.
// F16: xh is ONE fp16 plane (xl unused).
template <int NN, bool F16>
__global__ void __launch_bounds__(256)
k_thin_wgrad_planes_partial(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl, int ldx,
                            const float* __restrict__ dZ, long M, int N, int K, long chunk, float* __restrict__ part,
                            float* __restrict__ partb, float x_mul) {
  __shared__ float red[8][3][256 + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long m0 = (long)blockIdx.x * chunk, m1 = min(M, m0 + chunk);
  const int lpr = K <= 128 ? 16 : 32;        // lanes per row
  const int rpw = 32 / lpr;                  // rows per warp-wide load
  const int sub = lane / lpr;                // which of those rows this lane reads
  const int k = (lane % lpr) * 8;
  const bool active = k < K;
  float acc[NN][8];
#pragma unroll
  for (int n = 0; n < NN; n++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[n][j] = 0.f;
  float bsum[NN];
#pragma unroll
  for (int n = 0; n < NN; n++) bsum[n] = 0.f;
  const int step = 8 * rpw;                  // rows consumed by the 8 warps per load slot
  for (long m = m0 + warp * rpw + sub; m < m1; m += 4 * step) {  // rows m, m+step, m+2 step, m+3 step in flight
    uint4 h[4], l[4];
    float g[4][NN];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long mm = m + (long)step * u;
      const bool ok = mm < m1;
      h[u] = (ok && active) ? __ldg(reinterpret_cast<const uint4*>(xh + mm * ldx + k)) : make_uint4(0, 0, 0, 0);
      l[u] = (!F16 && ok && active && xl) ? __ldg(reinterpret_cast<const uint4*>(xl + mm * ldx + k)) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int n = 0; n < NN; n++) g[u][n] = ok ? __ldg(dZ + mm * NN + n) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint32_t hw[4] = {h[u].x, h[u].y, h[u].z, h[u].w}, lw[4] = {l[u].x, l[u].y, l[u].z, l[u].w};
      float x[8];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (F16) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[q]));
          x[2 * q] = f.x * x_mul; x[2 * q + 1] = f.y * x_mul;  // x_mul: the power of two the plane carries (exact)
        } else {
          x[2 * q] = __uint_as_float(hw[q] << 16) + __uint_as_float(lw[q] << 16);
          x[2 * q + 1] = __uint_as_float(hw[q] & 0xFFFF0000u) + __uint_as_float(lw[q] & 0xFFFF0000u);
        }
      }
#pragma unroll
      for (int n = 0; n < NN; n++) {
#pragma unroll
        for (int j = 0; j < 8; j++) acc[n][j] = fmaf(g[u][n], x[j], acc[n][j]);
        bsum[n] += g[u][n];
      }
    }
  }
  if (rpw == 2) {  // fold the two row groups of the warp: lane i += lane i + 16
#pragma unroll
    for (int n = 0; n < NN; n++) {
#pragma unroll
      for (int j = 0; j < 8; j++) acc[n][j] += __shfl_down_sync(0xffffffffu, acc[n][j], 16);
      bsum[n] += __shfl_down_sync(0xffffffffu, bsum[n], 16);
    }
  }
#pragma unroll
  for (int n = 0; n < NN; n++)
    if (active && sub == 0)
#pragma unroll
      for (int j = 0; j < 8; j++) red[warp][n][k + j] = acc[n][j];
  __syncthreads();
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int n = i / K, kk = i % K;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) s += red[w][n][kk];
    part[(long)blockIdx.x * N * K + i] = s;
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int n = 0; n < NN; n++) red[warp][n][0] = bsum[n];
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) s += red[w][threadIdx.x][0];
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

// out[i, coff + j] += sum_z ws[z, i, j] in a FIXED order (deterministic, no atomics).  A block is 64 column quads x 4
// z-groups: thread (q, g) sums the splits z = g, g+4, ... of 4 consecutive columns with 128-bit loads (ldw is a multiple
// of 4), the four groups meet in shared memory as (g0 + g1) + (g2 + g3).  A second, optional segment reduces the bias
// partials in the same launch.
__global__ void __launch_bounds__(256)
k_reduce_partials_tc(const float* __restrict__ ws, int splits, long stride, int rows, int cols, int ldw, float* __restrict__ out, int ldo,
                     int coff, const float* __restrict__ ws2, int n2, long stride2, float* __restrict__ out2) {
  __shared__ float4 red[4][64];
  const int q = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int c4 = (cols + 3) >> 2;
  const long n1 = (long)rows * c4;
  const long idx = (long)blockIdx.x * 64 + q;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int i = 0, j = 0;
  if (idx < n1) {
    i = (int)(idx / c4); j = (int)(idx % c4) * 4;
    const float* src = ws + (long)i * ldw + j;
    if ((ldw & 3) == 0 && (stride & 3) == 0) {
#pragma unroll 4
      for (int z = g; z < splits; z += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + (long)z * stride));
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    } else {
      for (int z = g; z < splits; z += 4) {
        const float* p = src + (long)z * stride;
        s.x += p[0];
        if (j + 1 < cols) s.y += p[1];
        if (j + 2 < cols) s.z += p[2];
        if (j + 3 < cols) s.w += p[3];
      }
    }
  } else if (idx - n1 < n2) {  // bias segment: one column per thread quad slot
    const int jb = (int)(idx - n1);
    for (int z = g; z < splits; z += 4) s.x += ws2[(long)z * stride2 + jb];
  }
  red[g][q] = s;
  __syncthreads();
  if (g != 0) return;
  const float4 a = red[0][q], b = red[1][q], c = red[2][q], d = red[3][q];
  const float4 t = make_float4((a.x + b.x) + (c.x + d.x), (a.y + b.y) + (c.y + d.y), (a.z + b.z) + (c.z + d.z), (a.w + b.w) + (c.w + d.w));
  if (idx < n1) {
    float* o = out + (long)i * ldo + coff + j;
    o[0] += t.x;
    if (j + 1 < cols) o[1] += t.y;
    if (j + 2 < cols) o[2] += t.z;
    if (j + 3 < cols) o[3] += t.w;
  } else if (idx - n1 < n2) {
    out2[idx - n1] += t.x;
  }
}

// The same reduction for MANY splits (the thin heads write ~600 partial rows of a few hundred floats): 16 z-groups per
// column quad instead of 4, so a thread's chain of dependent loads is 16x shorter than its split count; the groups meet in
// shared memory in a fixed pairwise tree (deterministic).
__global__ void __launch_bounds__(1024)
k_reduce_partials_wide(const float* __restrict__ ws, int splits, long stride, int rows, int cols, int ldw, float* __restrict__ out, int ldo,
                       int coff, const float* __restrict__ ws2, int n2, long stride2, float* __restrict__ out2) {
  __shared__ float4 red[16][64];
  const int q = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int c4 = (cols + 3) >> 2;
  const long n1 = (long)rows * c4;
  const long idx = (long)blockIdx.x * 64 + q;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int i = 0, j = 0;
  if (idx < n1) {
    i = (int)(idx / c4); j = (int)(idx % c4) * 4;
    const float* src = ws + (long)i * ldw + j;
    const bool vec = (ldw & 3) == 0 && (stride & 3) == 0;
#pragma unroll 4
    for (int z = g; z < splits; z += 16) {
      const float* p = src + (long)z * stride;
      if (vec) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p));
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      } else {
        s.x += p[0];
        if (j + 1 < cols) s.y += p[1];
        if (j + 2 < cols) s.z += p[2];
        if (j + 3 < cols) s.w += p[3];
      }
    }
  } else if (idx - n1 < n2) {
    const int jb = (int)(idx - n1);
    for (int z = g; z < splits; z += 16) s.x += ws2[(long)z * stride2 + jb];
  }
  red[g][q] = s;
  __syncthreads();
  for (int w = 8; w > 0; w >>= 1) {  // pairwise tree over the groups: g += g + w
    if (g < w) {
      const float4 a = red[g][q], b = red[g + w][q];
      red[g][q] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    __syncthreads();
  }
  if (g != 0) return;
  const float4 t = red[0][q];
  if (idx < n1) {
    float* o = out + (long)i * ldo + coff + j;
    o[0] += t.x;
    if (j + 1 < cols) o[1] += t.y;
    if (j + 2 < cols) o[2] += t.z;
    if (j + 3 < cols) o[3] += t.w;
  } else if (idx - n1 < n2) {
    out2[idx - n1] += t.x;
  }
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

namespace {
// Encoding a tensor map costs microseconds of host time and a step issues ~270 of them over a fixed set of buffers:
// memoise by (base, rows, cols, pitch, box) so steady-state steps only pay a hash lookup.
struct TmapKey {
  const void* base; uint64_t rows, cols, pitch; uint32_t box_rows;
  bool operator==(const TmapKey& o) const { return base == o.base && rows == o.rows && cols == o.cols && pitch == o.pitch && box_rows == o.box_rows; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = (uint64_t)(uintptr_t)k.base * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.cols * 1315423911ull + (h << 6) + (h >> 2));
    h ^= (k.pitch * 2654435761ull + (h << 6) + (h >> 2));
    h ^= ((uint64_t)k.box_rows * 40503ull + (h << 6) + (h >> 2));
    return (size_t)h;
  }
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
}  // namespace

void tc_tmap_cache_clear() {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  g_tmap_cache.clear();
}

int tc_make_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows) {
  return tc_make_tmap_box(out, base, rows, cols, pitch_elems, 64, box_rows);
}

// box_cols = 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B); box_rows packed with box_cols in the key
int tc_make_tmap_box(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_cols,
                     uint32_t box_rows) {
  const TmapKey key{base, rows, cols, pitch_elems, box_rows | (box_cols << 16)};
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) { *out = it->second; return 0; }
  }
  const int rc = tc_encode_tmap_box(out, base, rows, cols, pitch_elems, box_cols, box_rows);
  if (rc == 0) {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, *out);
  }
  return rc;
}

int tc_encode_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows) {
  return tc_encode_tmap_box(out, base, rows, cols, pitch_elems, 64, box_rows);
}

int tc_encode_tmap_box(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_cols,
                       uint32_t box_rows) {
  if (box_cols != 64 && box_cols != 32) { set_error("tensor map: box of %u columns unsupported", box_cols); return 100001; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return 100002; }
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {pitch_elems * 2};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu pitch=%llu box_rows=%u", (int)r, base,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_elems, box_rows);
    return 100001;
  }
  return 0;
}

int tc_smem_bytes(int BN, int n_stages, bool mn_major) {
  const int ring = n_stages * (kABytes + BN * 128);
  // K-major: the epilogue restages the output tile in 6 x 16 KB slots of the ring; MN-major: + 8 KB tile of ones
  return (mn_major ? ring + 8192 : (ring > 6 * 16384 ? ring : 6 * 16384)) + 1024;
}

int tc_pick_stages(int BN, int n_kblocks, bool mn_major) {
  // K-major: two CTAs per SM (<= ~110 KB each); MN-major (wgrad, HBM-bound): one CTA per SM, deep ring
  const int budget = mn_major ? 200 * 1024 : 110 * 1024;
  int s = (budget - 1024 - (mn_major ? 8192 : 0)) / (kABytes + BN * 128);
  if (s > kMaxStages) s = kMaxStages;
  if (s > n_kblocks && n_kblocks > 0) s = n_kblocks;
  return s < 1 ? 1 : s;
}

static int persist_stages(int BN) {
  const int budget = 227 * 1024 - (kConstFloats * 4 + 256) - 1024 - 4 * 16384;  // 227 KB per CTA minus static smem
  int s = budget / (kABytes + BN * 128);
  return s > kMaxStages ? kMaxStages : (s < 1 ? 1 : s);
}
int tc_launch(const TcParams& p_in, bool mn_major, dim3 grid, cudaStream_t st) {
  TcParams p = p_in;
  if (p.BN % 16 || p.BN < 16 || p.BN > 256 || p.n_stages < 1 || p.n_stages > kMaxStages) { set_error("tc_launch: bad BN=%d stages=%d", p.BN, p.n_stages); return 100001; }
  if (mn_major && p.BN % 64) { set_error("tc_launch: MN-major needs BN %% 64 == 0 (got %d)", p.BN); return 100001; }
  if (!mn_major && p.BN % 64) { set_error("tc_launch: K-major epilogue needs BN %% 64 == 0 (got %d)", p.BN); return 100001; }
  if (!mn_major) {
    NERF_TRY(ensure_kernel_smem((const void*)k_tc_gemm_persist, 227 * 1024 - (kConstFloats * 4 + 256)));  // per device
    const int sms = device_sm_count();
    if (p.n_valid % 32 || p.n_valid > 512) { set_error("tc_launch: n_valid=%d must be a multiple of 32, <= 512", p.n_valid); return 100001; }
    p.n_stages = persist_stages(p.BN);
    const size_t smem_p = (size_t)p.n_stages * (kABytes + p.BN * 128) + 4 * 16384 + 1024;
    const int tiles = (int)(grid.x * grid.y);
    k_tc_gemm_persist<<<tiles < sms ? tiles : sms, kThreads, smem_p, st>>>(p);
    NERF_CHECK_LAUNCH();
    return 0;
  }
  const int np = p.n_pass > 1 ? 2 : 1;
  const int stage_bytes = np * (kABytes + p.BN * 128);
  int stages = (220 * 1024 - 1024 - 8192) / stage_bytes;
  p.n_stages = stages > kMaxStages ? kMaxStages : (stages < 1 ? 1 : stages);
  const size_t smem = (size_t)p.n_stages * stage_bytes + 8192 + 1024;
  NERF_TRY(ensure_kernel_smem((const void*)k_tc_wgrad, 220 * 1024));
  k_tc_wgrad<<<grid, kThreads, smem, st>>>(p);
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

// dst[job.dst + i] = src[job.src + i]: the fused kernels' constants (biases, head weights) in one launch
__global__ void k_gather_f32(const float* __restrict__ src, float* __restrict__ dst, const __grid_constant__ GatherJobs jobs) {
  const GatherJobs::Job j = jobs.job[blockIdx.x];
  for (int i = threadIdx.x; i < j.n; i += blockDim.x) dst[j.dst + i] = src[j.src + i];
}

int launch_gather_f32(const float* src, float* dst, const GatherJobs& jobs, cudaStream_t st) {
  if (jobs.n <= 0) return 0;
  k_gather_f32<<<jobs.n, 128, 0, st>>>(src, dst, jobs);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_f32_to_planes_batch(const PlaneJobs& jobs, cudaStream_t st) {
  if (jobs.n <= 0) return 0;
  long nmax = 0;
  for (int i = 0; i < jobs.n; i++) nmax = jobs.job[i].n > nmax ? jobs.job[i].n : nmax;
  long bx = cdiv(nmax, 256);
  if (bx > 64) bx = 64;
  k_f32_to_planes_batch<<<dim3((unsigned)bx, (unsigned)jobs.n), 256, 0, st>>>(jobs);
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

int launch_thin_dgrad_planes(const float* dZ, const float* W, long M, int N, int K, const uint32_t* mask_bits, int ld_bits,
                             __nv_bfloat16* oh, __nv_bfloat16* ol, int ldo, cudaStream_t st, void* o16v, const float* scale16) {
  __half* o16 = static_cast<__half*>(o16v);
  if (o16 && !scale16) { set_error("thin_dgrad_planes: the fp16 plane needs its scale"); return 100001; }
  if (K % 8 || K > 2048 || N < 1 || N > 4 || ((uintptr_t)W & 15)) { set_error("thin_dgrad_planes: N=%d K=%d / W alignment", N, K); return 100001; }
  const int rows_per_block = 256 / (K / 8);
  const unsigned threads = (unsigned)(rows_per_block * (K / 8));  // whole rows per block
  long blocks = cdiv(M, rows_per_block);
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride over the rows: the weights are loaded once per thread
  switch (N) {
    case 1: k_thin_dgrad_planes<1><<<(unsigned)blocks, threads, 0, st>>>(dZ, W, M, K, mask_bits, ld_bits, oh, ol, ldo, o16, scale16); break;
    case 2: k_thin_dgrad_planes<2><<<(unsigned)blocks, threads, 0, st>>>(dZ, W, M, K, mask_bits, ld_bits, oh, ol, ldo, o16, scale16); break;
    case 3: k_thin_dgrad_planes<3><<<(unsigned)blocks, threads, 0, st>>>(dZ, W, M, K, mask_bits, ld_bits, oh, ol, ldo, o16, scale16); break;
    default: k_thin_dgrad_planes<4><<<(unsigned)blocks, threads, 0, st>>>(dZ, W, M, K, mask_bits, ld_bits, oh, ol, ldo, o16, scale16); break;
  }
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_reduce_partials2(const float* ws, int splits, long split_stride, int rows, int cols, int ldw, float* out, int ldo, int coff,
                            const float* ws2, int n2, long stride2, float* out2, cudaStream_t st) {
  const long n = (long)rows * ((cols + 3) / 4) + (ws2 ? n2 : 0);
  if (splits >= 64)
    k_reduce_partials_wide<<<(unsigned)cdiv(n, 64), 1024, 0, st>>>(ws, splits, split_stride, rows, cols, ldw, out, ldo, coff, ws2,
                                                                   ws2 ? n2 : 0, stride2, out2);
  else
    k_reduce_partials_tc<<<(unsigned)cdiv(n, 64), 256, 0, st>>>(ws, splits, split_stride, rows, cols, ldw, out, ldo, coff, ws2, ws2 ? n2 : 0,
                                                                 stride2, out2);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_reduce_job(const ReduceJob& job, cudaStream_t st) {
  if (!job.ws) return 0;
  if ((job.ldw & 3) || (job.stride & 3)) { set_error("reduce job: ldw and stride must be multiples of 4"); return 100001; }
  const long n = (long)job.rows * ((job.cols + 3) / 4) + job.n2;
  k_reduce_job<<<(unsigned)cdiv(n, 128), 128, 0, st>>>(job);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_reduce_partials(const float* ws, int splits, long split_stride, int rows, int cols, int ldw, float* out, int ldo,
                           int coff, cudaStream_t st) {
  return launch_reduce_partials2(ws, splits, split_stride, rows, cols, ldw, out, ldo, coff, nullptr, 0, 0, nullptr, st);
}

long thin_wgrad_chunk(long M) {  // ~4 blocks per SM, at least 256 rows each
  long c = cdiv(M, 148 * 4);
  c = cdiv(c, 32) * 32;
  return c < 256 ? 256 : c;
}

int launch_thin_wgrad_planes(const float* dZ, const __nv_bfloat16* xh, const __nv_bfloat16* xl, int ldx, float* dW, float* db,
                             long M, int N, int K, float* workspace, cudaStream_t st, bool x_f16, float x_mul) {
  if (N > 4) { set_error("thin_wgrad_planes: N=%d", N); return 100001; }
  if (N > 3 || K % 8) { set_error("thin_wgrad_planes: N=%d K=%d unsupported", N, K); return 100001; }
  const long chunk = thin_wgrad_chunk(M);
  const int chunks = (int)cdiv(M, chunk);
  float* part = workspace;
  float* partb = workspace + (size_t)chunks * N * 256;
  for (int k0 = 0; k0 < K; k0 += 256) {  // the kernel covers 256 columns (8 per lane) per pass
    const int kp = K - k0 < 256 ? K - k0 : 256;
    const __nv_bfloat16* xlk = xl ? xl + k0 : nullptr;
    if (N == 1 && x_f16) k_thin_wgrad_planes_partial<1, true><<<chunks, 256, 0, st>>>(xh + k0, nullptr, ldx, dZ, M, N, kp, chunk, part, partb, x_mul);
    else if (N == 3 && x_f16) k_thin_wgrad_planes_partial<3, true><<<chunks, 256, 0, st>>>(xh + k0, nullptr, ldx, dZ, M, N, kp, chunk, part, partb, x_mul);
    else if (N == 1) k_thin_wgrad_planes_partial<1, false><<<chunks, 256, 0, st>>>(xh + k0, xlk, ldx, dZ, M, N, kp, chunk, part, partb, 1.0f);
    else if (N == 3) k_thin_wgrad_planes_partial<3, false><<<chunks, 256, 0, st>>>(xh + k0, xlk, ldx, dZ, M, N, kp, chunk, part, partb, 1.0f);
    else { set_error("thin_wgrad_planes: N=%d unsupported", N); return 100001; }
    NERF_CHECK_LAUNCH();
    NERF_TRY(launch_reduce_partials2(part, chunks, (long)N * kp, N, kp, kp, dW, K, k0, (db && k0 == 0) ? partb : nullptr, N, N, db, st));
  }
  return 0;
}

// ---------------------------------------------------------------- fp16 operand planes of the fp32-accurate mode's wgrad
namespace {

// max |x| over the head gradients of a level -> the power-of-two scale of its fp16 dZ planes.  Non-negative floats order like
// their bit patterns, so the block maxima meet in one atomicMax; the last block (ticket) turns the maximum into
// s = 2^-floor(log2 max), 1/s and re-zeroes the scratch for the next launch.
__global__ void __launch_bounds__(256) k_dz_scale(const float* __restrict__ a, long na, const float* __restrict__ b, long nb, float act_scale,
                                                  float* __restrict__ out, unsigned* __restrict__ scratch) {
  __shared__ unsigned red[8];
  unsigned m = 0u;
  const long stride = (long)gridDim.x * blockDim.x, t0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long na4 = na >> 2, nb4 = nb >> 2;
  for (long i = t0; i < na4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(a) + i);
    m = max(max(m, __float_as_uint(fabsf(v.x))), max(__float_as_uint(fabsf(v.y)), max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
  }
  for (long i = na4 * 4 + t0; i < na; i += stride) m = max(m, __float_as_uint(fabsf(a[i])));
  for (long i = t0; i < nb4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(b) + i);
    m = max(max(m, __float_as_uint(fabsf(v.x))), max(__float_as_uint(fabsf(v.y)), max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
  }
  for (long i = nb4 * 4 + t0; i < nb; i += stride) m = max(m, __float_as_uint(fabsf(b[i])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; w++) m = max(m, red[w]);
    atomicMax(&scratch[0], m);
    __threadfence();
    if (atomicAdd(&scratch[1], 1u) == gridDim.x - 1) {  // last block: every maximum has arrived
      __threadfence();
      const unsigned mx = atomicExch(&scratch[0], 0u);
      scratch[1] = 0u;
      unsigned e = (mx >> 23) & 0xFFu;  // biased exponent of the maximum; 0 (all zero / subnormal) and inf / nan: no scaling
      if (e == 0u || e == 255u) e = 127u;
      if (e > 253u) e = 253u;
      out[0] = __uint_as_float((254u - e) << 23);  // 2^(127 - e)
      out[1] = __uint_as_float(e << 23);           // 2^(e - 127)
      out[2] = out[1] * act_scale;
    }
  }
}

__global__ void __launch_bounds__(256) k_f32_to_f16_plane(const float* __restrict__ src, int sp, long rows, int cols, __half* __restrict__ dst,
                                                          int dp, int dcols) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * dcols) return;
  const long r = idx / dcols;
  const int c = (int)(idx % dcols);
  dst[r * dp + c] = __float2half_rn(c < cols ? src[r * sp + c] : 0.f);
}

}  // namespace

int launch_dz_scale(const float* d_raw_rgb, const float* d_raw_density, long M, float act_scale, float* out, unsigned* scratch, cudaStream_t st) {
  if (((uintptr_t)d_raw_rgb | (uintptr_t)d_raw_density) & 15) { set_error("dz_scale: head gradients must be 16-byte aligned"); return 100001; }
  long blocks = cdiv(M, 256 * 4);
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  k_dz_scale<<<(unsigned)blocks, 256, 0, st>>>(d_raw_rgb, 3 * M, d_raw_density, M, act_scale, out, scratch);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_f32_to_f16_plane(const float* src, int src_pitch, long rows, int cols, void* dst, int dst_pitch, int dst_cols, cudaStream_t st) {
  const long n = rows * dst_cols;
  if (n <= 0) return 0;
  k_f32_to_f16_plane<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(src, src_pitch, rows, cols, static_cast<__half*>(dst), dst_pitch, dst_cols);
  NERF_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------- weight planes of the fp16 + fp8-correction representation
namespace {

// One thread per (row n, column k) of a layer's weight matrix W [rows, cols] (mlp_fused_split.cu, REP = 1):
//   k <  enc_from  (multiplies the tensor-memory-resident activations)   p0[n, k] = fp16(2^10 w) = 2^10 wh;
//                  p1, seen as bytes, per 64-column k-block b: byte 128 b + j = E4M3(2^6 wh), byte 128 b + 64 + j = E4M3(2^15 wl)
//   k >= enc_from  (multiplies the bf16 hi/lo encodings from shared memory) p0 / p1 = bf16 hi / lo of 2^15 w
//   k >= cols      zero padding up to kpad
__global__ void __launch_bounds__(256) k_f32_to_f8c_planes(const __grid_constant__ F8cJobs jobs) {
  const F8cJobs::Job& j = jobs.job[blockIdx.y];
  const long total = (long)j.rows * j.kpad;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / j.kpad), k = (int)(idx % j.kpad);
    const float w = k < j.cols ? j.src[(long)n * j.cols + k] : 0.f;
    uint16_t* p0 = reinterpret_cast<uint16_t*>(j.p0) + (long)n * j.kpad;
    if (k < j.enc_from) {
      const float t = w * 1024.f;
      const uint16_t hb = (uint16_t)(pack_f16x2_sat(t, 0.f) & 0xFFFFu);  // |w| >= 64 saturates instead of becoming inf
      const float hf = __half2float(__ushort_as_half(hb));
      p0[k] = hb;
      uint8_t* row8 = reinterpret_cast<uint8_t*>(j.p1) + (long)n * j.kpad * 2;
      const int b = k >> 6, jj = k & 63;
      row8[128 * b + jj] = (uint8_t)(pack_e4m3x2_f32(hf * 0.0625f, 0.f) & 0xFFu);             // 2^6 wh = 2^-4 (2^10 wh)
      row8[128 * b + 64 + jj] = (uint8_t)(pack_e4m3x2_f32((t - hf) * 32.f, 0.f) & 0xFFu);      // 2^15 wl = 2^5 (2^10 wl)
    } else {
      const float sw = w * 32768.f;
      const __nv_bfloat16 hi = __float2bfloat16_rn(sw);
      p0[k] = __bfloat16_as_ushort(hi);
      reinterpret_cast<uint16_t*>(j.p1)[(long)n * j.kpad + k] = __bfloat16_as_ushort(__float2bfloat16_rn(sw - __bfloat162float(hi)));
    }
  }
}

}  // namespace

int launch_f32_to_f8c_planes(const F8cJobs& jobs, cudaStream_t st) {
  if (jobs.n <= 0) return 0;
  for (int i = 0; i < jobs.n; i++)
    if (jobs.job[i].enc_from % 64 || jobs.job[i].kpad % 64) { set_error("f8c planes: k-blocks of 64 columns"); return 100001; }
  k_f32_to_f8c_planes<<<dim3(64, (unsigned)jobs.n), 256, 0, st>>>(jobs);
  NERF_CHECK_LAUNCH();
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

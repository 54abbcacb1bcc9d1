// f8_mma_probe.cu — experiment (not product): the building blocks of an fp16 + fp8-correction split product on tcgen05.
//   (1) kind::f16 with FP16 operands: A [128 x 64] in tensor memory (lane = row, two k per 32-bit column), B [N x 64] K-major in
//       128B-swizzled shared memory;
//   (2) kind::f8f6f4 with E4M3 operands ACCUMULATING ONTO THE SAME fp32 accumulator: A8 [128 x 128] in tensor memory — assumed
//       layout lane = row, FOUR consecutive k per 32-bit column (byte 0 = lowest k), K = 32 per MMA = 8 columns — and B8 [N x 128]
//       K-major in 128B-swizzled shared memory (one 128-byte row = 128 k).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I nerf_or_nothing_b200/csrc scripts/f8_mma_probe.cu -o scripts/f8_mma_probe
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sm100.cuh"

using namespace nerf::sm100;

constexpr int M = 128, N = 128, K16 = 64, K8 = 128;

__device__ __forceinline__ void tmem_st_32(uint32_t taddr, const uint32_t (&r)[32]) {
  tmem_st_16(taddr, r);
  tmem_st_16(taddr + 16, r + 16);
}
__device__ __forceinline__ void umma_f8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f8f6f4 instruction descriptor: c_format F32, a/b format 0 = E4M3 (1 = E5M2), K-major, N >> 3, M >> 4
__host__ __device__ constexpr uint32_t make_idesc_f8(int m, int n, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// mode bit 0: run the f16 part; bit 1: run the fp8 part
__global__ void __launch_bounds__(160) k_probe(const __half* A, const __half* B, const uint8_t* A8, const uint8_t* B8, float* D, int mode) {
  __shared__ __align__(1024) uint8_t bsm[N * 128];   // B16[n][k], 64 fp16 per 128-byte row
  __shared__ __align__(1024) uint8_t b8sm[N * 128];  // B8[n][k], 128 fp8 per 128-byte row
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < N * K16; i += blockDim.x) {
    const int n = i / K16, k = i % K16;
    *reinterpret_cast<__half*>(bsm + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = B[i];
  }
  for (int i = threadIdx.x; i < N * K8; i += blockDim.x) {
    const int n = i / K8, k = i % K8;
    b8sm[n * 128 + (((k >> 4) ^ (n & 7)) << 4) + (k & 15)] = B8[i];
  }
  fence_proxy_async();
  if (warp == 4) tmem_alloc<512>(&tbase_s);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tbase_s;
  const uint32_t A_COL = 256, A8_COL = 320;  // A16: 64 fp16 = 32 columns; A8: 128 fp8 = 32 columns
  if (warp < 4) {
    uint32_t r[32];
    const int row = threadIdx.x;
    for (int c = 0; c < 32; c++) r[c] = *reinterpret_cast<const uint32_t*>(&A[row * K16 + 2 * c]);
    tmem_st_32(tb + A_COL + ((uint32_t)(warp * 32) << 16), r);
    for (int c = 0; c < 32; c++) r[c] = *reinterpret_cast<const uint32_t*>(&A8[row * K8 + 4 * c]);
    tmem_st_32(tb + A8_COL + ((uint32_t)(warp * 32) << 16), r);
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 4 && lane == 0) {
    uint32_t acc = 0u;
    if (mode & 1) {
      const uint32_t idesc = make_idesc_f16(M, N, false, false);
      for (int k = 0; k < K16 / 16; k++, acc = 1u)
        umma_bf16_ts(tb, tb + A_COL + k * 8, make_smem_desc(smem_u32(bsm) + k * 32, 16, 1024), idesc, acc);
    }
    if (mode & 2) {
      const uint32_t idesc8 = make_idesc_f8(M, N, 0u);
      for (int k = 0; k < K8 / 32; k++, acc = 1u)
        umma_f8_ts(tb, tb + A8_COL + k * 8, make_smem_desc(smem_u32(b8sm) + k * 32, 16, 1024), idesc8, acc);
    }
    umma_commit(&bar);
  }
  if (warp < 4) {
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tb + c0 + ((uint32_t)(warp * 32) << 16), r);
      tmem_ld_wait();
      for (int j = 0; j < 32; j++) D[threadIdx.x * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tb);
}

int main() {
  std::vector<__half> A(M * K16), B(N * K16);
  std::vector<uint8_t> A8(M * K8), B8(N * K8);
  std::vector<float> Af(M * K16), Bf(N * K16), A8f(M * K8), B8f(N * K8), out(M * N);
  srand(1);
  for (int i = 0; i < M * K16; i++) { A[i] = __float2half((rand() % 2001 - 1000) / 500.f); Af[i] = __half2float(A[i]); }
  for (int i = 0; i < N * K16; i++) { B[i] = __float2half((rand() % 2001 - 1000) / 500.f); Bf[i] = __half2float(B[i]); }
  auto f8 = [](float x, uint8_t& b, float& f) {
    __nv_fp8_e4m3 v(x);
    b = *reinterpret_cast<uint8_t*>(&v);
    f = float(v);
  };
  for (int i = 0; i < M * K8; i++) f8((rand() % 2001 - 1000) / 500.f, A8[i], A8f[i]);
  for (int i = 0; i < N * K8; i++) f8((rand() % 2001 - 1000) / 500.f, B8[i], B8f[i]);
  __half *dA, *dB;
  uint8_t *dA8, *dB8;
  float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dA8, A8.size()); cudaMalloc(&dB8, B8.size());
  cudaMalloc(&dD, out.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dA8, A8.data(), A8.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB8, B8.data(), B8.size(), cudaMemcpyHostToDevice);
  int bad = 0;
  for (int mode = 1; mode <= 3; mode++) {
    cudaMemset(dD, 0, out.size() * 4);
    k_probe<<<1, 160>>>(dA, dB, dA8, dB8, dD, mode);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; m++)
      for (int n = 0; n < N; n++) {
        double s = 0;
        if (mode & 1) for (int k = 0; k < K16; k++) s += (double)Af[m * K16 + k] * Bf[n * K16 + k];
        if (mode & 2) for (int k = 0; k < K8; k++) s += (double)A8f[m * K8 + k] * B8f[n * K8 + k];
        maxerr = fmax(maxerr, fabs(out[m * N + n] - s));
        maxref = fmax(maxref, fabs(s));
      }
    const bool ok = e == cudaSuccess && maxerr < 1e-4 * maxref;
    bad += !ok;
    printf("mode %d (%s%s): kernel %s, max |D - ref| = %g (max |ref| = %g) -> %s\n", mode, mode & 1 ? "f16 " : "", mode & 2 ? "e4m3" : "",
           cudaGetErrorString(e), maxerr, maxref, ok ? "HOLDS" : "MISMATCH");
  }
  return bad;
}

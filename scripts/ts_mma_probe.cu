// ts_mma_probe.cu — experiment (not product): does tcgen05.mma with the A operand in TENSOR MEMORY behave as assumed?
// Assumption under test: for kind::f16, M=128, A[128 x K] lives in TMEM with lane = row and each 32-bit column holding
// two consecutive K elements (low half = even k), K=16 per MMA = 8 columns; written by tcgen05.st.32x32b.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I nerf_or_nothing_b200/csrc scripts/ts_mma_probe.cu -o scripts/ts_mma_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sm100.cuh"

using namespace nerf::sm100;

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(160) k_probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
  __shared__ __align__(1024) uint8_t bsm[N * 128];  // B[n][k] K-major, 128B rows, 128B swizzle
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    const int off = n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(bsm + off) = B[i];
  }
  fence_proxy_async();
  if (warp == 4) tmem_alloc<256>(&tbase_s);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tbase_s;
  const uint32_t A_COL = 128;  // A operand at TMEM columns [128, 160): 64 bf16 = 32 columns
  if (warp < 4) {
    uint32_t r[32];
    const int row = threadIdx.x;
    for (int c = 0; c < 32; c++) {
      const uint16_t lo = *reinterpret_cast<const uint16_t*>(&A[row * K + 2 * c]);
      const uint16_t hi = *reinterpret_cast<const uint16_t*>(&A[row * K + 2 * c + 1]);
      r[c] = (uint32_t)lo | ((uint32_t)hi << 16);
    }
    tmem_st_32x32(tb + A_COL + ((uint32_t)(warp * 32) << 16), r);
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = make_idesc_bf16(M, N, false, false);
    const uint32_t b_base = smem_u32(bsm);
    for (int k = 0; k < K / 16; k++)
      umma_bf16_ts(tb, tb + A_COL + k * 8, make_smem_desc(b_base + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
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
  if (warp == 4) tmem_dealloc<256>(tb);
}

int main() {
  std::vector<__nv_bfloat16> A(M * K), B(N * K);
  std::vector<float> Af(M * K), Bf(N * K), ref(M * N), out(M * N);
  srand(1);
  for (int i = 0; i < M * K; i++) { A[i] = __float2bfloat16((rand() % 2001 - 1000) / 500.f); Af[i] = __bfloat162float(A[i]); }
  for (int i = 0; i < N * K; i++) { B[i] = __float2bfloat16((rand() % 2001 - 1000) / 500.f); Bf[i] = __bfloat162float(B[i]); }
  for (int m = 0; m < M; m++)
    for (int n = 0; n < N; n++) {
      float s = 0;
      for (int k = 0; k < K; k++) s += Af[m * K + k] * Bf[n * K + k];
      ref[m * N + n] = s;
    }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, out.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, out.size() * 4);
  k_probe<<<1, 160>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int i = 0; i < M * N; i++) { maxerr = fmax(maxerr, fabs(out[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
  printf("TS MMA probe: max |D - ref| = %g (max |ref| = %g) -> %s\n", maxerr, maxref, maxerr < 1e-3 * maxref ? "LAYOUT ASSUMPTION HOLDS" : "MISMATCH");
  printf("D[0,0..3] = %g %g %g %g   ref = %g %g %g %g\n", out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
  return 0;
}

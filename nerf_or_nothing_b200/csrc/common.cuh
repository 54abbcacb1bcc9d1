// common.cuh — shared host/device helpers of libnerfb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

namespace nerf {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
long launch_count();

#define NERF_CUDA(expr)                                                                     \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      ::nerf::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                \
                        cudaGetErrorString(e_));                                            \
      return (int)e_;                                                                       \
    }                                                                                       \
  } while (0)

#define NERF_CHECK_LAUNCH()                                                                 \
  do {                                                                                      \
    ::nerf::count_launch();                                                                 \
    cudaError_t e_ = cudaGetLastError();                                                    \
    if (e_ != cudaSuccess) {                                                                \
      ::nerf::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,            \
                        cudaGetErrorString(e_));                                            \
      return (int)e_;                                                                       \
    }                                                                                       \
  } while (0)

#define NERF_TRY(expr)                 \
  do {                                 \
    int nerf_status__ = (expr);        \
    if (nerf_status__ != 0) return nerf_status__; \
  } while (0)

inline long cdiv(long a, long b) { return (a + b - 1) / b; }

// Per-DEVICE one-time kernel setup.  The dynamic-shared-memory opt-in (cudaFuncAttributeMaxDynamicSharedMemorySize) is a
// property of (kernel, device), so a process that drives several GPUs must set it on each of them; the SM count is
// per device too.  Both are memoised per (kernel, current device); returns 0 or the cudaError_t of the first failure.
int ensure_kernel_smem(const void* kernel, int dyn_smem_bytes);
int device_sm_count();  // of the current device (memoised)
// Launch a persistent one-CTA-per-SM kernel, optionally as thread-block clusters of `cluster` CTAs: sets the dynamic shared memory
// opt-in (per device), rounds the grid up to whole clusters and caps it by what the device can hold at once
// (cudaOccupancyMaxActiveClusters: clusters are placed inside a GPC, so an odd SM left over in a GPC cannot host one — a
// persistent kernel must be ONE wave).  `arg` points at the kernel's single by-value parameter.
int launch_persistent_clusters(const void* kernel, int grid, int threads, size_t smem, int smem_optin_bytes, int cluster, void* arg,
                               cudaStream_t st);

// Optional in-stream kernel timing (CUDA events around each launch group on the launching stream) so that
// bench.py can report per-kernel durations measured live inside the timed region.  Off by default.
enum ProfCat {
  PC_SAMPLE = 0, PC_ENCODE, PC_MLP_FWD, PC_MLP_HEADS_FWD, PC_COMPOSITE_FWD, PC_LOSS, PC_COMPOSITE_BWD,
  PC_MLP_DGRAD, PC_MLP_WGRAD, PC_MLP_HEADS_BWD, PC_ADAM, PC_COMM, PC_CAST, PC_MISC, PC_COUNT
};
const char* prof_name(int cat);
struct Profiler;
extern thread_local Profiler* g_prof;  // set by the API layer around a call on the calling thread; nullptr = profiling off
void prof_begin(int cat, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
  cudaStream_t st;
  ProfScope(int cat, cudaStream_t s) : st(s) { if (g_prof) prof_begin(cat, s); }
  ~ProfScope() { if (g_prof) prof_end(st); }
};

#ifdef __CUDACC__
// Philox4x32-10 (Salmon et al. SC'11): counter-based replacement of the reference's time-seeded cuRAND
// XORWOW states (.cu:17-23).  Returns word 0 of the output block.
__device__ __forceinline__ uint32_t philox4x32_10_w0(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c0;
}
// u in [0,1): counter = (ray, index, step, level), key = seed
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t ray, uint32_t idx, uint32_t step,
                                                uint32_t level) {
  const uint32_t x = philox4x32_10_w0(ray, idx, step, level, (uint32_t)seed, (uint32_t)(seed >> 32));
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }  // .cu:9
// softplus(x) = log(1 + e^x) (.cu:14), evaluated without overflow for large x
__device__ __forceinline__ float softplusf_(float x) { return x > 30.0f ? x : log1pf(expf(x)); }
#endif

}  // namespace nerf

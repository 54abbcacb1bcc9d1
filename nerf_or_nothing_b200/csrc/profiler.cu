// profiler.cu — in-stream CUDA-event timing of launch groups (see common.cuh).  Used by bench.py to measure
// each kernel family's duration live inside the timed region; never enabled by default.
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "profiler.cuh"

namespace nerf {

thread_local Profiler* g_prof = nullptr;

static const char* kNames[PC_COUNT] = {
    "sample_t_vals", "cast_rays+encode", "mlp_fwd_gemm", "mlp_fwd_heads", "composite_fwd", "loss_gradient",
    "composite_bwd", "mlp_dgrad_gemm", "mlp_wgrad_gemm", "mlp_bwd_heads", "adam", "allreduce", "cast_planes", "misc"};
const char* prof_name(int cat) { return cat >= 0 && cat < PC_COUNT ? kNames[cat] : "?"; }

cudaEvent_t Profiler::get() {
  if (used == pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    pool.push_back(e);
  }
  return pool[used++];
}

void Profiler::collect() {
  if (spans.empty()) { used = 0; return; }
  cudaEventSynchronize(spans.back().e1);
  for (auto& s : spans) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, s.e0, s.e1) == cudaSuccess) {
      ms[s.cat] += t;
      launches[s.cat] += s.launches;
      nspans[s.cat]++;
    }
  }
  spans.clear();
  used = 0;
}

void Profiler::reset() {
  collect();
  for (int i = 0; i < PC_COUNT; i++) { ms[i] = 0; launches[i] = 0; nspans[i] = 0; }
}

Profiler::~Profiler() {
  for (auto e : pool) cudaEventDestroy(e);
}

void prof_begin(int cat, cudaStream_t st) {
  Profiler* p = g_prof;
  if (!p) return;
  if (p->depth++ > 0) return;  // nested scopes are attributed to the outermost one
  if (p->spans.size() > 8192) p->collect();
  p->open.cat = cat;
  p->open.e0 = p->get();
  p->open.launches = launch_count();
  cudaEventRecord(p->open.e0, st);
}

void prof_end(cudaStream_t st) {
  Profiler* p = g_prof;
  if (!p) return;
  if (--p->depth > 0) return;
  p->open.e1 = p->get();
  cudaEventRecord(p->open.e1, st);
  p->open.launches = launch_count() - p->open.launches;
  p->spans.push_back(p->open);
}

// ---- per-device kernel setup (declared in common.cuh)
namespace {
std::mutex g_dev_mu;
std::map<std::pair<const void*, int>, std::pair<int, cudaError_t>> g_smem_set;  // (kernel, device) -> (bytes granted, status)
std::map<int, int> g_sms;
}  // namespace

int ensure_kernel_smem(const void* kernel, int dyn_smem_bytes) {
  int dev = 0;
  NERF_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_dev_mu);
  auto key = std::make_pair(kernel, dev);
  auto it = g_smem_set.find(key);
  if (it == g_smem_set.end() || (it->second.second == cudaSuccess && it->second.first < dyn_smem_bytes)) {
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem_bytes);
    g_smem_set[key] = std::make_pair(dyn_smem_bytes, e);
    it = g_smem_set.find(key);
  }
  if (it->second.second != cudaSuccess) {
    set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize = %d) failed on device %d: %s", dyn_smem_bytes, dev,
              cudaGetErrorString(it->second.second));
    return (int)it->second.second;
  }
  return 0;
}

int device_sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::lock_guard<std::mutex> lock(g_dev_mu);
  auto it = g_sms.find(dev);
  if (it != g_sms.end()) return it->second;
  int sms = 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  g_sms[dev] = sms;
  return sms;
}

int launch_persistent_clusters(const void* kernel, int grid, int threads, size_t smem, int smem_optin_bytes, int cluster, void* arg,
                               cudaStream_t st) {
  NERF_TRY(ensure_kernel_smem(kernel, smem_optin_bytes));
  if (cluster < 1) cluster = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)((grid + cluster - 1) / cluster * cluster));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = (unsigned)cluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  if (cluster > 1) {
    static std::map<std::pair<const void*, int>, int> cache;  // (kernel, device) -> clusters resident at once
    int dev = 0;
    NERF_CUDA(cudaGetDevice(&dev));
    int max_clusters = 0;
    const int sms = device_sm_count();
    {
      std::lock_guard<std::mutex> lock(g_dev_mu);
      auto it = cache.find({kernel, dev});
      if (it == cache.end()) {
        cudaLaunchConfig_t probe = cfg;
        probe.gridDim = dim3((unsigned)(sms / cluster * cluster));
        int n = 0;
        NERF_CUDA(cudaOccupancyMaxActiveClusters(&n, kernel, &probe));
        it = cache.emplace(std::make_pair(kernel, dev), n).first;
      }
      max_clusters = it->second;
    }
    if (max_clusters < 1) { set_error("no %d-CTA cluster of this kernel fits on the device", cluster); return 100001; }
    if ((int)cfg.gridDim.x > cluster * max_clusters) cfg.gridDim = dim3((unsigned)(cluster * max_clusters));
  }
  void* args[] = {arg};
  NERF_CUDA(cudaLaunchKernelExC(&cfg, kernel, args));
  count_launch();
  return 0;
}

}  // namespace nerf

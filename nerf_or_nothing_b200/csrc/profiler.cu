// profiler.cu — in-stream CUDA-event timing of launch groups (see common.cuh).  Used by bench.py to measure
// each kernel family's duration live inside the timed region; never enabled by default.
#include <vector>

#include "profiler.cuh"

namespace nerf {

Profiler* g_prof = nullptr;

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

}  // namespace nerf

// profiler.cuh — state behind the ProfScope hooks of common.cuh.
#pragma once
#include <vector>

#include "common.cuh"

namespace nerf {

struct Profiler {
  struct Span {
    int cat = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    long launches = 0;
  };
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  std::vector<Span> spans;
  Span open;
  int depth = 0;
  double ms[PC_COUNT] = {0};
  long launches[PC_COUNT] = {0};
  long nspans[PC_COUNT] = {0};
  cudaEvent_t get();
  void collect();  // waits for the last recorded span, folds every span into ms/launches
  void reset();
  ~Profiler();
};

}  // namespace nerf

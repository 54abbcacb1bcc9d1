// mlp.cuh — the MipNeRF MLP as an engine with two implementations:
//   SimtMlp (mlp_simt.cu)  strict fp32 on CUDA cores             NERF_PRECISION_FP32
//   TcMlp   (mlp_tc.cu)    tcgen05/TMEM GEMMs, bf16 / bf16x3    NERF_PRECISION_BF16_TC / _FP32_TC
// Orchestration mirrors AcceleratedMLP::get_output / get_gradient (ANU/AcceleratedMLP.cpp:214-321) with the
// call-site defects of SURVEY Appendix A (D1, D2, D5, D6) resolved to the standard chain rule.
#pragma once
#include <vector>

#include "encode_rows.cuh"
#include "kernels.cuh"

namespace nerf {

struct LayerInfo {
  int out, in_a, in_b;  // in_b: conjoined encoding width (skip / condition layer), else 0
  long w_off, b_off;    // offsets into the flat parameter buffer (order W0..W(L-1), b0..b(L-1))
};

struct MlpShape {
  int L = 0, D = 0, C = 0, W = 0, Wc = 0, P = 0, Dd = 0;
  std::vector<LayerInfo> layers;
  long n_params = 0;
  // layer table of ANU/AcceleratedMLP.cpp:131-154 == SN/MLP.cs:72-77
  void build(int depth, int width, int depth_cond, int width_cond, int skip, int deg_point, int deg_view) {
    D = depth; C = depth_cond; W = width; Wc = width_cond; P = 6 * deg_point; Dd = 3 + 6 * deg_view;
    L = D + C + 2;
    layers.assign(L, LayerInfo{});
    layers[0] = {W, P, 0, 0, 0};
    for (int i = 1; i < D; i++) layers[i] = {W, W, (skip > 0 && i % skip == 0) ? P : 0, 0, 0};
    layers[D] = {1, W, 0, 0, 0};
    layers[D + 1] = {Wc, W, Dd, 0, 0};
    for (int i = 1; i < C; i++) layers[D + 1 + i] = {Wc, Wc, 0, 0, 0};
    layers[D + C + 1] = {3, Wc, 0, 0, 0};
    long off = 0;
    for (auto& l : layers) { l.w_off = off; off += (long)l.out * (l.in_a + l.in_b); }
    for (auto& l : layers) { l.b_off = off; off += l.out; }
    n_params = off;
  }
};

class MlpEngine {
 public:
  virtual ~MlpEngine() {}
  // allocate activation caches for `n_levels` levels of up to `max_rows` samples each
  virtual int init(const MlpShape& shape, long max_rows, int n_levels) = 0;
  // where the fused cast_rays+IPE kernel must write this level's encodings
  virtual EncodeOut encode_targets(int level) = 0;
  // AcceleratedMLP::get_output takes caller-provided fp32 encodings: bring them into the engine's layout
  virtual int import_encodings(int level, const float* enc_pos, const float* enc_dir, long M, cudaStream_t st) = 0;
  // once per step, after the parameters changed (TC: refresh the bf16 weight planes)
  virtual int prepare(const float* params, cudaStream_t st) = 0;
  // raw (pre-activation) heads: raw_density [M], raw_rgb [M,3]; caches activations of `level`
  virtual int forward(int level, long M, const float* params, float* raw_density, float* raw_rgb, cudaStream_t st) = 0;
  // Forward straight from the level's t-values and rays: an engine whose forward kernel builds the encodings itself sets
  // *handled = 1 and runs the whole level (training: caches as forward(); else as forward_only()); otherwise it leaves
  // *handled = 0 and the caller runs the encode kernel into encode_targets() followed by forward() / forward_only().
  virtual int forward_from_rays(int level, const RaySource& rays, long M, const float* params, float* raw_density, float* raw_rgb,
                                bool training, cudaStream_t st, int* handled) {
    (void)level; (void)rays; (void)M; (void)params; (void)raw_density; (void)raw_rgb; (void)training; (void)st;
    *handled = 0;
    return 0;
  }
  // same heads, but no backward pass will follow: an engine may skip the activation caches (rendering)
  virtual int forward_only(int level, long M, const float* params, float* raw_density, float* raw_rgb, cudaStream_t st) {
    return forward(level, M, params, raw_density, raw_rgb, st);
  }
  // accumulates dL/dparams into grads from dL/d raw heads
  virtual int backward(int level, long M, const float* params, float* grads, const float* d_raw_density,
                       const float* d_raw_rgb, cudaStream_t st) = 0;
  // parity hook: ReLU masks of hidden layer i (trunk 0..D-1, then condition layers) of `level`'s last training forward,
  // one bit per unit: bit j of word c of row m <=> Y[m, 32 c + j] > 0.  Only engines that keep bit planes have them.
  virtual int relu_bits(int level, int i, const uint32_t** bits, int* words_per_row) {
    (void)level; (void)i; (void)bits; (void)words_per_row;
    set_error("this precision mode keeps no ReLU bit planes");
    return 100003;
  }
  virtual size_t bytes_allocated() const = 0;
};

MlpEngine* make_simt_mlp();
MlpEngine* make_tc_mlp(bool split3, unsigned engine_flags);  // defined in mlp_tc.cu; flags = nerf_config.engine_flags

}  // namespace nerf

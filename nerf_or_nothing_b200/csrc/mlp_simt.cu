// mlp_simt.cu — strict-fp32 MLP engine (NERF_PRECISION_FP32) on the tiled SGEMMs of gemm_simt.cu.
// Forward = AcceleratedMLP::get_output (ANU/AcceleratedMLP.cpp:214-255), backward = get_gradient
// (:256-321) as the standard chain rule (SURVEY A-D1/D5/D6), gradients accumulated over levels.
// Only post-ReLU outputs are cached (relu'(Z) == (Y > 0)), half of the reference's Y+Z traffic.
#include "mlp.cuh"

namespace nerf {
namespace {

class SimtMlp : public MlpEngine {
 public:
  ~SimtMlp() override { release(); }

  int init(const MlpShape& shape, long max_rows, int n_levels) override {
    s_ = shape; max_rows_ = max_rows; n_levels_ = n_levels;
    dir_pitch_ = (s_.Dd + 3) / 4 * 4;  // 27 -> 28: keeps rows 16-byte aligned for vector loads
    levels_.resize(n_levels);
    for (auto& lv : levels_) {
      NERF_TRY(alloc(&lv.enc_pos, (size_t)max_rows * s_.P));
      NERF_TRY(alloc(&lv.enc_dir, (size_t)max_rows * dir_pitch_));
      lv.acts.resize(s_.D + s_.C);
      for (int i = 0; i < s_.D + s_.C; i++) NERF_TRY(alloc(&lv.acts[i], (size_t)max_rows * (i < s_.D ? s_.W : s_.Wc)));
    }
    const int mw = s_.W > s_.Wc ? s_.W : s_.Wc;
    NERF_TRY(alloc(&dz_[0], (size_t)max_rows * mw));
    NERF_TRY(alloc(&dz_[1], (size_t)max_rows * mw));
    size_t ws = 0;
    for (auto& l : s_.layers) {
      size_t need = l.out <= 4 ? thin_wgrad_workspace(max_rows, l.out, l.in_a) : dense_wgrad_workspace(max_rows, l.out, l.in_a);
      if (l.out > 4 && l.in_b > 0) {
        const size_t nb = dense_wgrad_workspace(max_rows, l.out, l.in_b);
        need = nb > need ? nb : need;
      }
      ws = need > ws ? need : ws;
    }
    NERF_TRY(alloc(&ws_, ws));
    return 0;
  }

  EncodeOut encode_targets(int level) override {
    EncodeOut o;
    o.enc_pos_f32 = levels_[level].enc_pos;
    o.enc_dir_f32 = levels_[level].enc_dir;
    o.dir_pitch_f32 = dir_pitch_;
    return o;
  }

  int import_encodings(int level, const float* enc_pos, const float* enc_dir, long M, cudaStream_t st) override {
    NERF_CUDA(cudaMemcpyAsync(levels_[level].enc_pos, enc_pos, (size_t)M * s_.P * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return launch_pad_rows(enc_dir, s_.Dd, levels_[level].enc_dir, dir_pitch_, M, s_.Dd, st);
  }

  int prepare(const float*, cudaStream_t) override { return 0; }

  int forward(int level, long M, const float* params, float* raw_density, float* raw_rgb, cudaStream_t st) override {
    Level& lv = levels_[level];
    const int D = s_.D, C = s_.C;
    const float* h = lv.enc_pos;
    int ldh = s_.P;
    for (int i = 0; i < D; i++) {  // trunk (.cpp:217-232)
      const LayerInfo& l = s_.layers[i];
      ProfScope ps(PC_MLP_FWD, st);
      NERF_TRY(launch_dense_fwd(h, ldh, l.in_a, l.in_b ? lv.enc_pos : nullptr, s_.P, l.in_b, params + l.w_off,
                                params + l.b_off, lv.acts[i], nullptr, M, l.out, ACT_RELU, st));
      h = lv.acts[i]; ldh = s_.W;
    }
    {  // density head, N=1, K=width (the intended shape; the reference swaps them, A-D1)
      const LayerInfo& l = s_.layers[D];
      ProfScope ps(PC_MLP_HEADS_FWD, st);
      NERF_TRY(launch_thin_fwd(h, s_.W, params + l.w_off, params + l.b_off, raw_density, M, 1, s_.W, st));
    }
    const float* c = h;
    int ldc = s_.W;
    for (int i = 0; i < C; i++) {  // condition layers (.cpp:236-247)
      const LayerInfo& l = s_.layers[D + 1 + i];
      ProfScope ps(PC_MLP_FWD, st);
      NERF_TRY(launch_dense_fwd(c, ldc, l.in_a, l.in_b ? lv.enc_dir : nullptr, dir_pitch_, l.in_b, params + l.w_off,
                                params + l.b_off, lv.acts[D + i], nullptr, M, l.out, ACT_RELU, st));
      c = lv.acts[D + i]; ldc = s_.Wc;
    }
    {  // rgb head, N=3
      const LayerInfo& l = s_.layers[D + C + 1];
      ProfScope ps(PC_MLP_HEADS_FWD, st);
      NERF_TRY(launch_thin_fwd(c, s_.Wc, params + l.w_off, params + l.b_off, raw_rgb, M, 3, s_.Wc, st));
    }
    return 0;
  }

  int backward(int level, long M, const float* params, float* grads, const float* d_raw_density,
               const float* d_raw_rgb, cudaStream_t st) override {
    Level& lv = levels_[level];
    const int D = s_.D, C = s_.C, W = s_.W, Wc = s_.Wc;
    float* cur = dz_[0];
    float* nxt = dz_[1];
    {  // rgb head (.cpp:259-264): dW, db, and dZ of the last condition layer (masked by its ReLU)
      const LayerInfo& l = s_.layers[D + C + 1];
      ProfScope ps(PC_MLP_HEADS_BWD, st);
      NERF_TRY(launch_thin_wgrad(d_raw_rgb, lv.acts[D + C - 1], Wc, grads + l.w_off, grads + l.b_off, M, 3, Wc, ws_, st));
      NERF_TRY(launch_thin_dgrad(d_raw_rgb, params + l.w_off, cur, M, 3, Wc, lv.acts[D + C - 1], false, st));
    }
    for (int i = C - 1; i >= 0; i--) {  // condition layers (.cpp:269-282)
      const LayerInfo& l = s_.layers[D + 1 + i];
      const float* in = i == 0 ? lv.acts[D - 1] : lv.acts[D + i - 1];
      const int ldin = i == 0 ? W : Wc;
      { ProfScope ps(PC_MLP_WGRAD, st);
      NERF_TRY(launch_dense_wgrad(cur, in, ldin, l.in_a, l.in_b ? lv.enc_dir : nullptr, dir_pitch_, l.in_b,
                                  grads + l.w_off, grads + l.b_off, M, l.out, ws_, st)); }
      ProfScope ps(PC_MLP_DGRAD, st);
      if (i > 0) {
        NERF_TRY(launch_dense_dgrad(cur, params + l.w_off, l.in_a + l.in_b, nxt, M, l.out, l.in_a, nullptr, nullptr,
                                    lv.acts[D + i - 1], false, st));
      } else {
        // gradient into the trunk output = condition dgrad (direction part dropped, SN/MLP.cs:148)
        // + density-head dgrad as a rank-1 term (SN/MLP.cs:149-153), then the ReLU mask of trunk layer D-1
        const LayerInfo& ld = s_.layers[D];
        NERF_TRY(launch_dense_dgrad(cur, params + l.w_off, l.in_a + l.in_b, nxt, M, l.out, l.in_a, d_raw_density,
                                    params + ld.w_off, lv.acts[D - 1], false, st));
      }
      float* t = cur; cur = nxt; nxt = t;
    }
    {  // density head weights (.cpp:283-290)
      const LayerInfo& l = s_.layers[D];
      ProfScope ps(PC_MLP_HEADS_BWD, st);
      NERF_TRY(launch_thin_wgrad(d_raw_density, lv.acts[D - 1], W, grads + l.w_off, grads + l.b_off, M, 1, W, ws_, st));
    }
    for (int i = D - 1; i >= 0; i--) {  // trunk (.cpp:291-310)
      const LayerInfo& l = s_.layers[i];
      const float* in = i == 0 ? lv.enc_pos : lv.acts[i - 1];
      const int ldin = i == 0 ? s_.P : W;
      { ProfScope ps(PC_MLP_WGRAD, st);
      NERF_TRY(launch_dense_wgrad(cur, in, ldin, l.in_a, l.in_b ? lv.enc_pos : nullptr, s_.P, l.in_b, grads + l.w_off,
                                  grads + l.b_off, M, l.out, ws_, st)); }
      if (i > 0) {  // encodings need no gradient (.cu:154-182)
        ProfScope ps(PC_MLP_DGRAD, st);
        NERF_TRY(launch_dense_dgrad(cur, params + l.w_off, l.in_a + l.in_b, nxt, M, l.out, l.in_a, nullptr, nullptr,
                                    lv.acts[i - 1], false, st));
        float* t = cur; cur = nxt; nxt = t;
      }
    }
    return 0;
  }

  size_t bytes_allocated() const override { return bytes_; }

 private:
  struct Level {
    float *enc_pos = nullptr, *enc_dir = nullptr;
    std::vector<float*> acts;  // [0,D): trunk outputs [M,W]; [D,D+C): condition outputs [M,Wc]
  };
  int alloc(float** p, size_t n) {
    NERF_CUDA(cudaMalloc(p, n * sizeof(float)));
    bytes_ += n * sizeof(float);
    owned_.push_back(*p);
    return 0;
  }
  void release() {
    for (float* p : owned_) cudaFree(p);
    owned_.clear();
  }
  MlpShape s_;
  long max_rows_ = 0;
  int n_levels_ = 0, dir_pitch_ = 0;
  std::vector<Level> levels_;
  float* dz_[2] = {nullptr, nullptr};
  float* ws_ = nullptr;
  size_t bytes_ = 0;
  std::vector<float*> owned_;
};

}  // namespace

MlpEngine* make_simt_mlp() { return new SimtMlp(); }

}  // namespace nerf

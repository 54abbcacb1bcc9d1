// mlp_tc.cu — tensor-core MLP engine: every dense layer of the MipNeRF MLP forward, dgrad and wgrad
// on the tcgen05 GEMM of gemm_tc.cu.  NERF_PRECISION_BF16_TC keeps one bf16 plane per tensor;
// NERF_PRECISION_FP32_TC keeps hi/lo planes (x = hi + lo) and multiplies with the 3-term split.
//
// Orchestration is the same chain as AcceleratedMLP::get_output / get_gradient (ANU/AcceleratedMLP.cpp:214-321):
//   forward   Y_i planes are written by the GEMM epilogue (bias + ReLU fused) and are at once the next layer's
//             operand, the ReLU mask of the backward pass and the wgrad operand — nothing else is cached;
//   backward  dZ planes ping-pong between two buffers; dgrad's epilogue fuses the ReLU mask of the layer below and
//             the density head's rank-1 contribution; wgrad reads dZ and X "transposed" (MN-major UMMA operands),
//             splits the sample dimension over CTAs and reduces the fp32 partials in a fixed order.
// The N=1 / N=3 heads stay on CUDA cores (too thin for a UMMA tile) but read the same planes.
#include <algorithm>
#include <cstring>

#include "../../include/nerfb200.h"
#include "gemm_tc.cuh"
#include "mlp.cuh"

namespace nerf {
namespace {

struct Plane {
  __nv_bfloat16 *hi = nullptr, *lo = nullptr;
  __nv_bfloat16* f16 = nullptr;  // fp16 plane of the same tensor (typed as a 2-byte carrier): wgrad operand in the w16 mode
  int pitch = 0;  // elements per row (multiple of 64)
};
enum PlaneKind { PLANE_DEFAULT = 0, PLANE_F16_ONLY = 1, PLANE_PLUS_F16 = 2 };

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

class TcMlp : public MlpEngine {
 public:
  TcMlp(bool split3, unsigned flags) : split_(split3), flags_(flags) {}
  ~TcMlp() override {
    for (void* p : owned_) cudaFree(p);
  }

  int init(const MlpShape& shape, long max_rows, int n_levels) override {
    s_ = shape; max_rows_ = max_rows;
    if (s_.W % 64 || s_.Wc % 64 || s_.W > 512 || s_.Wc > 512) {
      set_error("tensor-core MLP needs widths that are multiples of 64 and <= 512 (got %d / %d)", s_.W, s_.Wc);
      return 100001;
    }
    pos_pitch_ = round_up(s_.P, 64);
    dir_pitch_ = round_up(s_.Dd, 64);
    // NERF_FLAG_WGRAD_FP16 (fp32-accurate mode, fused kernels; opt-in): the wgrad operands (activations, encodings, dZ) are kept
    // as ONE fp16 plane each instead of hi + lo bf16 planes — half the HBM traffic of the backward pass and one MMA per product
    // in wgrad.  The forward and the dgrad chain, where an error would propagate through the layers and flip ReLU masks, keep
    // the three-term products on chip.  Each wgrad operand then carries a 2^-12 relative rounding: on the real training step
    // the sums over ~5e5 samples average it out (whole-step gradient ~1e-5 of its scale against fp64), but a sum whose terms
    // cancel like a random walk keeps ~3e-4 of its scale — outside the 1e-4 contract of the mode, hence not the default.
    w16_ = (flags_ & NERF_FLAG_WGRAD_FP16) != 0;
    if (w16_ && !(split_ && can_fuse_forward() &&
                  !(flags_ & (NERF_FLAG_NO_FUSED_TRAIN_FORWARD | NERF_FLAG_NO_FUSED_DGRAD | NERF_FLAG_FUSED_ENCODE_TRAIN)))) {
      set_error("NERF_FLAG_WGRAD_FP16 needs NERF_PRECISION_FP32_TC with the fused training kernels (widths 256/128 or 128/64, no per-layer / encoder-warp flags)");
      return 100001;
    }
    // fp32-accurate mode, fused forward kernels: the three-term product as one fp16 MMA plus two E4M3 correction MMAs at twice
    // the rate (mlp_fused_split.cu, REP = 1) — measured as accurate as the bf16x3 split at the outputs (rendered rgb 9e-7 vs
    // 7e-7 against fp64) and 14 % faster on an 800x800 render.  Rendering by default; training when the activations leave as
    // fp16 planes anyway (NERF_FLAG_WGRAD_FP16).  NERF_FLAG_NO_FP8_CORRECTIONS keeps the bf16x3 kernels everywhere.
    f8c_ = split_ && can_fuse_forward() && !(flags_ & NERF_FLAG_NO_FP8_CORRECTIONS);
    levels_.resize(n_levels);
    if (w16_) {
      NERF_CUDA(cudaMalloc(&dz_sc_, (size_t)n_levels * 4 * sizeof(float) + 2 * sizeof(unsigned)));
      owned_.push_back(dz_sc_);
      NERF_CUDA(cudaMemset(dz_sc_, 0, (size_t)n_levels * 4 * sizeof(float) + 2 * sizeof(unsigned)));
      dz_sc_scratch_ = reinterpret_cast<unsigned*>(dz_sc_ + (size_t)n_levels * 4);
    }
    for (auto& lv : levels_) {
      NERF_TRY(alloc_plane(&lv.enc_pos, max_rows, pos_pitch_, w16_ ? PLANE_PLUS_F16 : PLANE_DEFAULT));
      NERF_TRY(alloc_plane(&lv.enc_dir, max_rows, dir_pitch_, w16_ ? PLANE_PLUS_F16 : PLANE_DEFAULT));
      lv.acts.resize(s_.D + s_.C);
      lv.bits.resize(s_.D + s_.C);
      for (int i = 0; i < s_.D + s_.C; i++) {
        const int w = i < s_.D ? s_.W : s_.Wc;
        NERF_TRY(alloc_plane(&lv.acts[i], max_rows, w, w16_ ? PLANE_F16_ONLY : PLANE_DEFAULT));
        const size_t nb = (size_t)max_rows * (w / 32) * sizeof(uint32_t);
        NERF_CUDA(cudaMalloc(&lv.bits[i], nb));
        owned_.push_back(lv.bits[i]);
        bytes_ += nb;
      }
    }
    const int mw = s_.W > s_.Wc ? s_.W : s_.Wc;
    NERF_TRY(alloc_plane(&dz_[0], max_rows, mw, w16_ ? PLANE_PLUS_F16 : PLANE_DEFAULT));
    if (!w16_) NERF_TRY(alloc_plane(&dz_[1], max_rows, mw));  // ping-pong partner of the per-layer dgrad launches
    wp_.resize(s_.L); wtp_.resize(s_.L);
    size_t ws = 1 << 20;
    for (int l = 0; l < s_.L; l++) {
      const LayerInfo& L = s_.layers[l];
      if (L.out <= 4) { ws = std::max(ws, (size_t)cdiv(max_rows, 256) * L.out * (L.in_a + 1) + 64); continue; }
      NERF_TRY(alloc_plane(&wp_[l], L.out, round_up(L.in_a + L.in_b, 64)));
      if (f8c_ && (l < s_.D || l == s_.D + 1)) {
        if (wf_.empty()) wf_.resize(s_.L);
        NERF_TRY(alloc_plane(&wf_[l], L.out, round_up(L.in_a + L.in_b, 64)));  // hi / lo = the W16 / W8 planes
      }
      if (needs_dgrad(l)) NERF_TRY(alloc_plane(&wtp_[l], L.in_a, L.out));
      for (int src = 0; src < 2; src++) {
        const int K = src == 0 ? L.in_a : L.in_b;
        if (K <= 0) continue;
        int splits; long split_len;
        wgrad_split(L.out, K, max_rows, &splits, &split_len);
        ws = std::max(ws, (size_t)splits * L.out * (round_up(K, 4) + 1) + 64);
      }
      ws = std::max(ws, (size_t)cdiv(max_rows, 2048) * L.out + 64);
    }
    NERF_CUDA(cudaMalloc(&ws_, ws * sizeof(float)));
    owned_.push_back(ws_);
    bytes_ += ws * sizeof(float);
    // wgrad partial tiles alternate between two buffers: a launch's partials are reduced by the NEXT wgrad launch
    for (int i = 0; i < 2; i++) {
      NERF_CUDA(cudaMalloc(&wsw_[i], ws * sizeof(float)));
      owned_.push_back(wsw_[i]);
      bytes_ += ws * sizeof(float);
    }
    pending_.ws = nullptr;
    defer_reduce_ = !(flags_ & NERF_FLAG_NO_DEFERRED_REDUCE);
    return 0;
  }

  EncodeOut encode_targets(int level) override {
    EncodeOut o;
    o.pos_hi = levels_[level].enc_pos.hi; o.pos_lo = levels_[level].enc_pos.lo; o.pos_pitch_h = pos_pitch_;
    o.dir_hi = levels_[level].enc_dir.hi; o.dir_lo = levels_[level].enc_dir.lo; o.dir_pitch_h = dir_pitch_;
    o.pos_f16 = levels_[level].enc_pos.f16; o.dir_f16 = levels_[level].enc_dir.f16;
    return o;
  }

  int import_encodings(int level, const float* enc_pos, const float* enc_dir, long M, cudaStream_t st) override {
    Level& lv = levels_[level];
    NERF_TRY(launch_f32_to_planes(enc_pos, s_.P, M, s_.P, lv.enc_pos.hi, lv.enc_pos.lo, pos_pitch_, pos_pitch_, false, 0, st));
    if (w16_) {
      NERF_TRY(launch_f32_to_f16_plane(enc_pos, s_.P, M, s_.P, lv.enc_pos.f16, pos_pitch_, pos_pitch_, st));
      NERF_TRY(launch_f32_to_f16_plane(enc_dir, s_.Dd, M, s_.Dd, lv.enc_dir.f16, dir_pitch_, dir_pitch_, st));
    }
    return launch_f32_to_planes(enc_dir, s_.Dd, M, s_.Dd, lv.enc_dir.hi, lv.enc_dir.lo, dir_pitch_, dir_pitch_, false, 0, st);
  }

  // refresh the bf16 weight planes (and their transposes for dgrad) from the fp32 master parameters
  int prepare(const float* params, cudaStream_t st) override {
    ProfScope ps(PC_CAST, st);
    fconsts_dirty_ = true;
    PlaneJobs jobs;
    jobs.n = 0;
    auto add = [&](const float* src, int sp, long rows, int cols, const Plane& d, int dcols, bool transpose, long drows_t) {
      PlaneJobs::Job& j = jobs.job[jobs.n++];
      j.src = src; j.sp = sp; j.rows = rows; j.cols = cols; j.hi = d.hi; j.lo = d.lo; j.dp = d.pitch; j.dcols = dcols;
      j.transpose = transpose ? 1 : 0; j.drows_t = drows_t;
      j.n = transpose ? drows_t * dcols : rows * dcols;
    };
    for (int l = 0; l < s_.L; l++) {
      const LayerInfo& L = s_.layers[l];
      if (L.out <= 4) continue;
      const int K = L.in_a + L.in_b;
      if (jobs.n + 2 > 32) { NERF_TRY(launch_f32_to_planes_batch(jobs, st)); jobs.n = 0; }
      add(params + L.w_off, K, L.out, K, wp_[l], wp_[l].pitch, false, 0);
      if (needs_dgrad(l)) add(params + L.w_off, K, L.out, L.in_a, wtp_[l], L.out, true, L.in_a);  // WT[k, n] = W[n, k] for k < in_a
    }
    NERF_TRY(launch_f32_to_planes_batch(jobs, st));
    wf_dirty_ = true;  // the fp8-correction planes are refreshed by the first forward that multiplies with them (ensure_f8c_planes)
    return 0;
  }

  // W16 / W8 planes of the fp16 + E4M3 representation, once per parameter version and only when a forward uses them (a training
  // step in the default mode never does)
  int ensure_f8c_planes(const float* params, cudaStream_t st) {
    if (!wf_dirty_) return 0;
    ProfScope ps(PC_CAST, st);
    {
      F8cJobs fj;
      fj.n = 0;
      for (int l = 0; l < s_.L; l++) {
        if (!(l < s_.D || l == s_.D + 1)) continue;
        const LayerInfo& L = s_.layers[l];
        if (fj.n == 16) { NERF_TRY(launch_f32_to_f8c_planes(fj, st)); fj.n = 0; }
        F8cJobs::Job& j = fj.job[fj.n++];
        j.src = params + L.w_off; j.p0 = wf_[l].hi; j.p1 = wf_[l].lo; j.rows = L.out; j.cols = L.in_a + L.in_b;
        j.enc_from = l == 0 ? 0 : L.in_a;  // layer 0 multiplies the position encoding only
        j.kpad = wf_[l].pitch;
      }
      NERF_TRY(launch_f32_to_f8c_planes(fj, st));
    }
    wf_dirty_ = false;
    return 0;
  }

  int forward(int level, long M, const float* params, float* raw_density, float* raw_rgb, cudaStream_t st) override {
    if (can_fuse_forward() && !(flags_ & NERF_FLAG_NO_FUSED_TRAIN_FORWARD))
      return fused_forward(level, M, params, raw_density, raw_rgb, true, nullptr, st);
    Level& lv = levels_[level];
    const int D = s_.D, C = s_.C;
    const Plane* h = &lv.enc_pos;
    int kh = s_.P;
    for (int i = 0; i < D; i++) {
      const LayerInfo& L = s_.layers[i];
      ProfScope ps(PC_MLP_FWD, st);
      Head hd;
      if (i == D - 1 && fuse_heads()) {  // density head (N=1) rides in the last trunk layer's epilogue
        const LayerInfo& Lh = s_.layers[D];
        hd.w = params + Lh.w_off; hd.b = params + Lh.b_off; hd.out = raw_density; hd.n = 1;
      }
      NERF_TRY(gemm_fwd(*h, kh, L.in_b ? &lv.enc_pos : nullptr, L.in_b, L.in_a, wp_[i], params + L.b_off, lv.acts[i], lv.bits[i], hd, M, L.out, st));
      h = &lv.acts[i]; kh = s_.W;
    }
    if (!fuse_heads()) {
      const LayerInfo& L = s_.layers[D];
      ProfScope ps(PC_MLP_HEADS_FWD, st);
      NERF_TRY(launch_thin_fwd_planes(h->hi, h->lo, h->pitch, params + L.w_off, params + L.b_off, raw_density, M, 1, s_.W, st));
    }
    const Plane* c = h;
    int kc = s_.W;
    for (int i = 0; i < C; i++) {
      const int l = D + 1 + i;
      const LayerInfo& L = s_.layers[l];
      ProfScope ps(PC_MLP_FWD, st);
      Head hd;
      if (i == C - 1 && fuse_heads()) {  // rgb head (N=3) rides in the last condition layer's epilogue
        const LayerInfo& Lh = s_.layers[D + C + 1];
        hd.w = params + Lh.w_off; hd.b = params + Lh.b_off; hd.out = raw_rgb; hd.n = 3;
      }
      NERF_TRY(gemm_fwd(*c, kc, L.in_b ? &lv.enc_dir : nullptr, L.in_b, L.in_a, wp_[l], params + L.b_off, lv.acts[D + i], lv.bits[D + i], hd, M, L.out, st));
      c = &lv.acts[D + i]; kc = s_.Wc;
    }
    if (!fuse_heads()) {
      const LayerInfo& L = s_.layers[D + C + 1];
      ProfScope ps(PC_MLP_HEADS_FWD, st);
      NERF_TRY(launch_thin_fwd_planes(c->hi, c->lo, c->pitch, params + L.w_off, params + L.b_off, raw_rgb, M, 3, s_.Wc, st));
    }
    return 0;
  }

  // Rendering: no backward follows, so in bf16 mode the whole net runs as one kernel with the activations kept in
  // tensor memory (mlp_fused.cu) instead of one GEMM launch per layer with the activations written to HBM.
  // The fused kernels are written for trunk / condition widths 256 / 128 (the reference network) and 128 / 64 (the narrow end of
  // the configs[4] sweep), one condition layer, encodings that fit 128 / 64 columns.
  bool can_fuse_forward() const {
    const bool widths = (s_.W == 256 && s_.Wc == 128) || (s_.W == 128 && s_.Wc == 64);
    return widths && s_.C == 1 && s_.D >= 2 && s_.D + 1 <= 12 && pos_pitch_ == 128 && dir_pitch_ == 64 && !(flags_ & NERF_FLAG_NO_FUSED_FORWARD);
  }

  int forward_only(int level, long M, const float* params, float* raw_density, float* raw_rgb, cudaStream_t st) override {
    if (!can_fuse_forward()) return forward(level, M, params, raw_density, raw_rgb, st);
    return fused_forward(level, M, params, raw_density, raw_rgb, false, nullptr, st);
  }

  // cast_rays + IPE + direction PE inside the fused forward kernel (encoder warps, encode_rows.cuh): no encode kernel runs.
  // Training: the encoder warps write the level's planes (the wgrad GEMMs of layer 0, the skip layer and the condition layer
  // read them); rendering: a per-CTA double-buffered scratch that stays in L2.
  int forward_from_rays(int level, const RaySource& rays, long M, const float* params, float* raw_density, float* raw_rgb,
                        bool training, cudaStream_t st, int* handled) override {
    *handled = 0;
    if (!can_fuse_forward()) return 0;
    if (training ? (!(flags_ & NERF_FLAG_FUSED_ENCODE_TRAIN) || (flags_ & NERF_FLAG_NO_FUSED_TRAIN_FORWARD)) : (flags_ & NERF_FLAG_NO_FUSED_ENCODE) != 0) return 0;
    if (rays.deg_point % 4 || 6 * rays.deg_point > 120 || rays.deg_view > 4) return 0;
    if (!training && !scr_pos_.hi) {  // 148 CTAs x 2 buffers x 256 rows (bf16 walks tile pairs): 29 MB with the lo planes
      scr_rows_ = (long)device_sm_count() * 2 * 256;
      NERF_TRY(alloc_plane(&scr_pos_, scr_rows_, pos_pitch_));
      NERF_TRY(alloc_plane(&scr_dir_, scr_rows_, dir_pitch_));
    }
    *handled = 1;
    return fused_forward(level, M, params, raw_density, raw_rgb, training, &rays, st);
  }

  // biases of the D trunk layers and the condition layer, then the density head (w[256], b) and the rgb head (w[3][128], b[3]):
  // the constants the fused kernels stage in shared memory, gathered once per parameter version
  int ensure_fconsts(const float* params, cudaStream_t st) {
    const int D = s_.D;
    const int head_d_off = D * 256 + 128, head_rgb_off = head_d_off + 260, n_consts = head_rgb_off + 3 * 128 + 4;
    if (!fconsts_) {
      NERF_CUDA(cudaMalloc(&fconsts_, n_consts * sizeof(float)));
      owned_.push_back(fconsts_);
      NERF_CUDA(cudaMemsetAsync(fconsts_, 0, n_consts * sizeof(float), st));
    }
    if (!fconsts_dirty_) return 0;
    ProfScope ps(PC_CAST, st);
    GatherJobs jobs;
    jobs.n = 0;
    auto add = [&](int dst, long src, int n) { jobs.job[jobs.n].dst = dst; jobs.job[jobs.n].src = src; jobs.job[jobs.n].n = n; jobs.n++; };
    for (int s = 0; s <= D; s++) {
      const LayerInfo& L = s_.layers[s < D ? s : D + 1];
      add(s * 256, L.b_off, L.out);
    }
    const LayerInfo& Ld = s_.layers[D];
    const LayerInfo& Lr = s_.layers[D + 2];
    add(head_d_off, Ld.w_off, s_.W); add(head_d_off + 256, Ld.b_off, 1);  // w[W] (zero beyond W), bias in a fixed slot
    for (int n = 0; n < 3; n++) add(head_rgb_off + n * 128, Lr.w_off + (long)n * s_.Wc, s_.Wc);  // w[3][Wc] in rows of 128
    add(head_rgb_off + 384, Lr.b_off, 3);
    NERF_TRY(launch_gather_f32(params, fconsts_, jobs, st));
    fconsts_dirty_ = false;
    return 0;
  }

  int fused_forward(int level, long M, const float* params, float* raw_density, float* raw_rgb, bool train, const RaySource* rays,
                    cudaStream_t st) {
    Level& lv = levels_[level];
    // where the encodings are: the level's planes, or (rendering with in-kernel encoding) the L2 scratch
    const bool scratch = rays && !train;
    const Plane& epos = scratch ? scr_pos_ : lv.enc_pos;
    const Plane& edir = scratch ? scr_dir_ : lv.enc_dir;
    const long scr_rows = scratch ? scr_rows_ : 0;
    const int D = s_.D;
    std::vector<int> bias_off(D + 1), kpad(D + 1), in_b(D + 1);
    std::vector<const __nv_bfloat16*> wpl(D + 1);
    for (int s = 0; s <= D; s++) {
      const int l = s < D ? s : D + 1;
      bias_off[s] = s * 256;
      kpad[s] = wp_[l].pitch; in_b[s] = s_.layers[l].in_b; wpl[s] = wp_[l].hi;
    }
    const int head_d_off = D * 256 + 128, head_rgb_off = head_d_off + 260, n_consts = head_rgb_off + 3 * 128 + 4;
    NERF_TRY(ensure_fconsts(params, st));
    const bool f16 = train && w16_;  // training in the w16 mode: every layer's activations leave as one fp16 plane
    const bool f8c = f8c_ && split_ && (!train || f16);
    if (f8c) NERF_TRY(ensure_f8c_planes(params, st));
    std::vector<__nv_bfloat16*> act_out(D + 1);
    for (int s = 0; s <= D; s++) act_out[s] = f16 ? lv.acts[s].f16 : lv.acts[s].hi;
    ProfScope ps(PC_MLP_FWD, st);
    if (split_) {
      std::vector<const __nv_bfloat16*> wlo(D + 1);
      std::vector<__nv_bfloat16*> act_lo(D + 1);
      for (int s = 0; s <= D; s++) { wlo[s] = wp_[s < D ? s : D + 1].lo; act_lo[s] = lv.acts[s].lo; }
      if (f8c)
        for (int s = 0; s <= D; s++) { wpl[s] = wf_[s < D ? s : D + 1].hi; wlo[s] = wf_[s < D ? s : D + 1].lo; }
      if (train) act_x_scale_ = f8c ? 0.03125f : 1.0f;  // the fp16 planes the backward pass will read then hold 32 a
      return launch_mlp_fused_forward_split(epos.hi, epos.lo, pos_pitch_, edir.hi, edir.lo, dir_pitch_, wpl.data(),
                                            wlo.data(), kpad.data(), in_b.data(), D, s_.W, s_.Wc, M, fconsts_, n_consts, head_d_off,
                                            head_rgb_off, bias_off.data(), raw_density, raw_rgb, train ? act_out.data() : nullptr,
                                            train ? act_lo.data() : nullptr, train ? lv.bits.data() : nullptr, rays, scr_rows, pair(), st, f16, f8c, pair_mma());
    }
    return launch_mlp_fused_forward(epos.hi, pos_pitch_, edir.hi, dir_pitch_, wpl.data(), kpad.data(), in_b.data(), D, s_.W,
                                    s_.Wc, M, fconsts_, n_consts, head_d_off, head_rgb_off, bias_off.data(), raw_density, raw_rgb,
                                    train ? act_out.data() : nullptr, train ? lv.bits.data() : nullptr, rays, scr_rows, pair(), st);
  }

  int backward(int level, long M, const float* params, float* grads, const float* d_raw_density, const float* d_raw_rgb,
               cudaStream_t st) override {
    const int s = backward_impl(level, M, params, grads, d_raw_density, d_raw_rgb, st);
    ProfScope ps(PC_MLP_WGRAD, st);
    const int f = flush_reduce(st);  // the last wgrad launch's partials have no successor to ride in
    return s ? s : f;
  }

  int flush_reduce(cudaStream_t st) {
    const ReduceJob job = pending_;
    pending_.ws = nullptr;
    return launch_reduce_job(job, st);
  }

  int backward_impl(int level, long M, const float* params, float* grads, const float* d_raw_density, const float* d_raw_rgb,
                    cudaStream_t st) {
    Level& lv = levels_[level];
    const int D = s_.D, C = s_.C, W = s_.W, Wc = s_.Wc;
    Plane* cur = &dz_[0];
    Plane* nxt = &dz_[1];
    // [s, 1/s, x/s]: power-of-two scale of this level's fp16 dZ planes, its inverse, and the inverse times the scale x the level's
    // activation planes carry (1, or 1/32 when the fp8-correction forward wrote 32 a)
    float* sc = w16_ ? dz_sc_ + 4 * level : nullptr;
    {  // rgb head
      const LayerInfo& L = s_.layers[D + C + 1];
      const Plane& x = lv.acts[D + C - 1];
      ProfScope ps(PC_MLP_HEADS_BWD, st);
      if (w16_) NERF_TRY(launch_dz_scale(d_raw_rgb, d_raw_density, M, act_x_scale_, sc, dz_sc_scratch_, st));
      NERF_TRY(launch_thin_wgrad_planes(d_raw_rgb, w16_ ? x.f16 : x.hi, x.lo, x.pitch, grads + L.w_off, grads + L.b_off, M, 3, Wc, ws_, st, w16_,
                                        act_x_scale_));
      NERF_TRY(launch_thin_dgrad_planes(d_raw_rgb, params + L.w_off, M, 3, Wc, lv.bits[D + C - 1], Wc / 32, cur->hi, cur->lo, cur->pitch, st,
                                        cur->f16, sc));
    }
    // The trunk's dgrad chain is one fused kernel in both tensor-core modes.  In the fp32-accurate mode it moves half the
    // bytes of the per-layer launches (each dZ is written once instead of written and re-read) and, since the next layer's
    // first k-blocks run under the second-half epilogue, measures 3.24 ms against their 3.60 ms per step at configs[1].
    // NERF_FLAG_NO_FUSED_DGRAD selects the per-layer launches (parity tests compare the two).
    if (can_fuse_forward() && !(flags_ & NERF_FLAG_NO_FUSED_DGRAD))
      return backward_fused_chain(level, M, params, grads, d_raw_density, *cur, st);
    for (int i = C - 1; i >= 0; i--) {
      const int l = D + 1 + i;
      const LayerInfo& L = s_.layers[l];
      const Plane& in = i == 0 ? lv.acts[D - 1] : lv.acts[D + i - 1];
      {
        ProfScope ps(PC_MLP_WGRAD, st);
        NERF_TRY(gemm_wgrad(*cur, in, L.in_a, L.in_b ? &lv.enc_dir : nullptr, L.in_b, grads + L.w_off, grads + L.b_off, M, L.out, st));
      }
      ProfScope ps(PC_MLP_DGRAD, st);
      if (i > 0) {
        NERF_TRY(gemm_dgrad(*cur, wtp_[l], *nxt, M, L.out, L.in_a, nullptr, nullptr, lv.bits[D + i - 1], st));
      } else {
        const LayerInfo& Ld = s_.layers[D];
        NERF_TRY(gemm_dgrad(*cur, wtp_[l], *nxt, M, L.out, L.in_a, d_raw_density, params + Ld.w_off, lv.bits[D - 1], st));
      }
      Plane* t = cur; cur = nxt; nxt = t;
    }
    {
      const LayerInfo& L = s_.layers[D];
      const Plane& x = lv.acts[D - 1];
      ProfScope ps(PC_MLP_HEADS_BWD, st);
      NERF_TRY(launch_thin_wgrad_planes(d_raw_density, x.hi, x.lo, x.pitch, grads + L.w_off, grads + L.b_off, M, 1, W, ws_, st));
    }
    for (int i = D - 1; i >= 0; i--) {
      const LayerInfo& L = s_.layers[i];
      const Plane& in = i == 0 ? lv.enc_pos : lv.acts[i - 1];
      {
        ProfScope ps(PC_MLP_WGRAD, st);
        NERF_TRY(gemm_wgrad(*cur, in, L.in_a, L.in_b ? &lv.enc_pos : nullptr, L.in_b, grads + L.w_off, grads + L.b_off, M, L.out, st));
      }
      if (i > 0) {
        ProfScope ps(PC_MLP_DGRAD, st);
        NERF_TRY(gemm_dgrad(*cur, wtp_[i], *nxt, M, L.out, L.in_a, nullptr, nullptr, lv.bits[i - 1], st));
        Plane* t = cur; cur = nxt; nxt = t;
      }
    }
    return 0;
  }

  // The whole dgrad chain (condition layer -> trunk layers D-1..1) is ONE kernel that keeps dZ in tensor
  // memory between layers and writes every layer's dZ once; the wgrad GEMMs then read those planes.
  int backward_fused_chain(int level, long M, const float* params, float* grads, const float* d_raw_density, const Plane& dz_cond,
                           cudaStream_t st) {
    auto& lv = levels_[level];
    const int D = s_.D, W = s_.W;
    if (dzs_.empty()) {  // first training step through this path
      dzs_.resize(D);
      for (int j = 0; j < D; j++) NERF_TRY(alloc_plane(&dzs_[j], max_rows_, W, w16_ ? PLANE_F16_ONLY : PLANE_DEFAULT));
    }
    NERF_TRY(ensure_fconsts(params, st));
    const int head_d_off = D * 256 + 128, n_consts = head_d_off + 260 + 3 * 128 + 4;
    std::vector<const __nv_bfloat16*> wt(D);
    std::vector<int> wt_pitch(D);
    std::vector<__nv_bfloat16*> dz_out(D);
    std::vector<const uint32_t*> masks(D);
    for (int j = 0; j < D; j++) {
      const int l = j == 0 ? D + 1 : D - j;  // the layer whose dgrad step j performs; it produces dZ of trunk layer D-1-j
      wt[j] = wtp_[l].hi; wt_pitch[j] = wtp_[l].pitch;
      dz_out[j] = w16_ ? dzs_[j].f16 : dzs_[j].hi; masks[j] = lv.bits[D - 1 - j];
    }
    {
      ProfScope ps(PC_MLP_DGRAD, st);
      if (split_) {
        std::vector<const __nv_bfloat16*> wt_lo(D);
        std::vector<__nv_bfloat16*> dz_lo(D);
        for (int j = 0; j < D; j++) { wt_lo[j] = wtp_[j == 0 ? D + 1 : D - j].lo; dz_lo[j] = dzs_[j].lo; }
        NERF_TRY(launch_mlp_fused_dgrad_split(dz_cond.hi, dz_cond.lo, dz_cond.pitch, wt.data(), wt_lo.data(), wt_pitch.data(), D, W, s_.Wc, M,
                                              fconsts_, n_consts, head_d_off, d_raw_density, dz_out.data(), dz_lo.data(), masks.data(), pair(), st,
                                              w16_ ? dz_sc_ + 4 * level : nullptr, pair_mma()));
      } else {
        NERF_TRY(launch_mlp_fused_dgrad(dz_cond.hi, dz_cond.pitch, wt.data(), wt_pitch.data(), D, W, s_.Wc, M, fconsts_, n_consts, head_d_off,
                                        d_raw_density, dz_out.data(), masks.data(), pair(), st));
      }
    }
    wgrad_unscale_ = w16_ ? dz_sc_ + 4 * level + 1 : nullptr;
    {  // condition layer
      const LayerInfo& L = s_.layers[D + 1];
      ProfScope ps(PC_MLP_WGRAD, st);
      NERF_TRY(gemm_wgrad(dz_cond, lv.acts[D - 1], L.in_a, &lv.enc_dir, L.in_b, grads + L.w_off, grads + L.b_off, M, L.out, st, true));
    }
    {  // density head
      const LayerInfo& L = s_.layers[D];
      const Plane& x = lv.acts[D - 1];
      ProfScope ps(PC_MLP_HEADS_BWD, st);
      NERF_TRY(launch_thin_wgrad_planes(d_raw_density, w16_ ? x.f16 : x.hi, x.lo, x.pitch, grads + L.w_off, grads + L.b_off, M, 1, W, ws_, st, w16_,
                                        act_x_scale_));
    }
    for (int i = D - 1; i >= 0; i--) {
      const LayerInfo& L = s_.layers[i];
      const Plane& in = i == 0 ? lv.enc_pos : lv.acts[i - 1];
      ProfScope ps(PC_MLP_WGRAD, st);
      NERF_TRY(gemm_wgrad(dzs_[D - 1 - i], in, L.in_a, L.in_b ? &lv.enc_pos : nullptr, L.in_b, grads + L.w_off, grads + L.b_off, M, L.out, st, i > 0));
    }
    return 0;
  }

  int relu_bits(int level, int i, const uint32_t** bits, int* words_per_row) override {
    if (level < 0 || level >= (int)levels_.size() || i < 0 || i >= s_.D + s_.C) { set_error("relu_bits: bad level / layer"); return 100001; }
    *bits = levels_[level].bits[i];
    *words_per_row = (i < s_.D ? s_.W : s_.Wc) / 32;
    return 0;
  }

  size_t bytes_allocated() const override { return bytes_; }

 private:
  struct Level {
    Plane enc_pos, enc_dir;
    std::vector<Plane> acts;
    std::vector<uint32_t*> bits;  // ReLU masks of acts[i] as bit planes [M, width/32]
  };
  struct Head { const float* w = nullptr; const float* b = nullptr; float* out = nullptr; int n = 0; };

  bool needs_dgrad(int l) const { return (l > 0 && l < s_.D) || (l > s_.D && l <= s_.D + s_.C); }

  int alloc_plane(Plane* p, long rows, int pitch, PlaneKind kind = PLANE_DEFAULT) {
    const size_t bytes = (size_t)rows * pitch * sizeof(__nv_bfloat16);
    p->pitch = pitch;
    if (kind != PLANE_DEFAULT) {
      NERF_CUDA(cudaMalloc(&p->f16, bytes));
      owned_.push_back(p->f16);
      NERF_CUDA(cudaMemset(p->f16, 0, bytes));
      bytes_ += bytes;
      if (kind == PLANE_F16_ONLY) return 0;
    }
    NERF_CUDA(cudaMalloc(&p->hi, bytes));
    owned_.push_back(p->hi);
    NERF_CUDA(cudaMemset(p->hi, 0, bytes));
    bytes_ += bytes;
    if (split_) {
      NERF_CUDA(cudaMalloc(&p->lo, bytes));
      owned_.push_back(p->lo);
      NERF_CUDA(cudaMemset(p->lo, 0, bytes));
      bytes_ += bytes;
    }
    p->pitch = pitch;
    return 0;
  }

  // passes of the split product: (A plane, B plane) with 0 = hi, 1 = lo
  int passes(int (*pa)[2]) const {
    pa[0][0] = 0; pa[0][1] = 0;
    if (!split_) return 1;
    pa[1][0] = 0; pa[1][1] = 1;
    pa[2][0] = 1; pa[2][1] = 0;
    return 3;
  }

  // Y = relu([A1|A2] W^T + b): K-major, grid (M/128, N/BN)
  bool fuse_heads() const { return s_.W <= 256 && s_.Wc <= 256; }  // the head needs the whole row in one CTA

  int gemm_fwd(const Plane& a1, int k1, const Plane* a2, int k2, int w_col2, const Plane& w, const float* bias, const Plane& out,
               uint32_t* bits, const Head& head, long M, int N, cudaStream_t st) {
    TcParams p;
    memset(&p, 0, sizeof(p));
    const int BN = N > 256 ? 256 : N;
    NERF_TRY(tc_make_tmap(&p.maps[0], a1.hi, M, a1.pitch, a1.pitch, 128));
    if (split_) NERF_TRY(tc_make_tmap(&p.maps[1], a1.lo, M, a1.pitch, a1.pitch, 128));
    if (a2) {
      NERF_TRY(tc_make_tmap(&p.maps[2], a2->hi, M, a2->pitch, a2->pitch, 128));
      if (split_) NERF_TRY(tc_make_tmap(&p.maps[3], a2->lo, M, a2->pitch, a2->pitch, 128));
    }
    NERF_TRY(tc_make_tmap(&p.maps[4], w.hi, N, w.pitch, w.pitch, BN));
    if (split_) NERF_TRY(tc_make_tmap(&p.maps[5], w.lo, N, w.pitch, w.pitch, BN));
    int pa[3][2];
    const int np = passes(pa);
    int n = 0;
    for (int seg = 0; seg < (a2 ? 2 : 1); seg++) {
      const int K = seg == 0 ? k1 : k2;
      const int bcol0 = seg == 0 ? 0 : w_col2;
      for (int kc = 0; kc < K; kc += 64)
        for (int q = 0; q < np; q++) {
          if (n >= TC_MAX_KB) { set_error("tc gemm: too many k-blocks"); return 100001; }
          p.kb[n].a = (int8_t)(seg * 2 + pa[q][0]);
          p.kb[n].b = (int8_t)(4 + pa[q][1]);
          p.kb[n].a_col = (int16_t)kc;
          p.kb[n].b_col = (int16_t)(bcol0 + kc);
          n++;
        }
    }
    NERF_TRY(tc_make_tmap(&p.maps[6], out.hi, M, N, out.pitch, 32));  // per-warp 32-row store boxes
    if (out.lo) NERF_TRY(tc_make_tmap(&p.maps[7], out.lo, M, N, out.pitch, 32));
    p.n_kb = n; p.M = M; p.BN = BN; p.n_valid = N; p.n_stages = tc_pick_stages(BN, n, false);
    p.epi = 0; p.bias = bias; p.act = ACT_RELU;
    p.bits_out = bits; p.ld_bits = N / 32;
    p.head_w = head.w; p.head_b = head.b; p.head_out = head.out; p.head_n = head.n;
    p.out_hi = out.hi; p.out_lo = out.lo; p.ld_out = out.pitch;
    return tc_launch(p, false, dim3((unsigned)cdiv(M, 128), (unsigned)cdiv(N, BN), 1), st);
  }

  // dX[M, k1] = mask(dZ[M,N] * WT^T (+ r1 v1^T)):  A = dZ planes (K-major over n), B = WT planes [k1, N]
  int gemm_dgrad(const Plane& dz, const Plane& wt, const Plane& out, long M, int N, int k1, const float* r1, const float* v1,
                 const uint32_t* mask_bits, cudaStream_t st) {
    TcParams p;
    memset(&p, 0, sizeof(p));
    const int BN = k1 > 256 ? 256 : k1;
    NERF_TRY(tc_make_tmap(&p.maps[0], dz.hi, M, N, dz.pitch, 128));
    if (split_) NERF_TRY(tc_make_tmap(&p.maps[1], dz.lo, M, N, dz.pitch, 128));
    NERF_TRY(tc_make_tmap(&p.maps[4], wt.hi, k1, N, wt.pitch, BN));
    if (split_) NERF_TRY(tc_make_tmap(&p.maps[5], wt.lo, k1, N, wt.pitch, BN));
    int pa[3][2];
    const int np = passes(pa);
    int n = 0;
    for (int kc = 0; kc < N; kc += 64)
      for (int q = 0; q < np; q++) {
        if (n >= TC_MAX_KB) { set_error("tc gemm: too many k-blocks"); return 100001; }
        p.kb[n].a = (int8_t)pa[q][0];
        p.kb[n].b = (int8_t)(4 + pa[q][1]);
        p.kb[n].a_col = (int16_t)kc;
        p.kb[n].b_col = (int16_t)kc;
        n++;
      }
    NERF_TRY(tc_make_tmap(&p.maps[6], out.hi, M, k1, out.pitch, 32));
    if (out.lo) NERF_TRY(tc_make_tmap(&p.maps[7], out.lo, M, k1, out.pitch, 32));
    p.n_kb = n; p.M = M; p.BN = BN; p.n_valid = k1; p.n_stages = tc_pick_stages(BN, n, false);
    p.epi = 1; p.r1 = r1; p.v1 = v1; p.mask_bits = mask_bits; p.ld_bits = k1 / 32;
    p.out_hi = out.hi; p.out_lo = out.lo; p.ld_out = out.pitch;
    return tc_launch(p, false, dim3((unsigned)cdiv(M, 128), (unsigned)cdiv(k1, BN), 1), st);
  }

  static void wgrad_split(int N, int K, long M, int* splits, long* split_len) {
    const int BN = K > 256 ? 256 : round_up(K, 64);
    const long tiles = cdiv(N, 128) * cdiv(K, BN);
    long s = cdiv(148, tiles);  // one CTA per SM, one wave
    const long maxs = cdiv(M, 256);
    if (s > maxs) s = maxs;
    if (s < 1) s = 1;
    *split_len = cdiv(cdiv(M, s), 64) * 64;
    *splits = (int)cdiv(M, *split_len);
  }

  // dW[N, k1+k2] += dZ^T [X1|X2]; db += colsum(dZ).  MN-major operands, reduction over the M samples split over CTAs.
  // x1_act (w16): x1 is an activation plane of the level (it may carry the fp8-correction forward's factor 32), not an encoding
  int gemm_wgrad(const Plane& dz, const Plane& x1, int k1, const Plane* x2, int k2, float* dW, float* db, long M, int N,
                 cudaStream_t st, bool x1_act = false) {
    const int ldw = k1 + k2;
    for (int src = 0; src < 2; src++) {
      const Plane* x = src == 0 ? &x1 : x2;
      const int K = src == 0 ? k1 : k2, coff = src == 0 ? 0 : k1;
      if (!x || K <= 0) continue;
      TcParams p;
      memset(&p, 0, sizeof(p));
      const int BN = K > 256 ? 256 : round_up(K, 64);
      const int xcols = round_up(K, 64) <= x->pitch ? round_up(K, 64) : x->pitch;
      if (w16_) {  // one fp16 plane per operand, one pass; dZ carries the level's scale (wgrad_unscale_ divides it out)
        if (!dz.f16 || !x->f16) { set_error("wgrad: fp16 planes missing"); return 100001; }
        NERF_TRY(tc_make_tmap(&p.maps[0], dz.f16, M, N, dz.pitch, 64));
        NERF_TRY(tc_make_tmap(&p.maps[4], x->f16, M, xcols, x->pitch, 64));
        p.f16_ops = 1;
      } else {
        NERF_TRY(tc_make_tmap(&p.maps[0], dz.hi, M, N, dz.pitch, 64));
        if (split_) NERF_TRY(tc_make_tmap(&p.maps[1], dz.lo, M, N, dz.pitch, 64));
        NERF_TRY(tc_make_tmap(&p.maps[4], x->hi, M, xcols, x->pitch, 64));
        if (split_) NERF_TRY(tc_make_tmap(&p.maps[5], x->lo, M, xcols, x->pitch, 64));
      }
      int pa[3][2];
      p.n_pass = w16_ ? 1 : passes(pa);
      if (w16_) { pa[0][0] = 0; pa[0][1] = 0; }
      for (int q = 0; q < p.n_pass; q++) { p.pass_a[q] = (int8_t)pa[q][0]; p.pass_b[q] = (int8_t)(4 + pa[q][1]); }
      int splits; long split_len;
      wgrad_split(N, K, M, &splits, &split_len);
      const int ldf = round_up(K, 4);
      p.split_len = (int)split_len; p.red_len = M; p.BN = BN; p.n_valid = K; p.rows_valid = N;
      p.n_stages = tc_pick_stages(BN, 1 << 20, true);
      float* wsp = wsw_[ws_flip_];
      ws_flip_ ^= 1;
      p.epi = 2; p.out_f32 = wsp; p.ld_f32 = ldf; p.split_stride = (long)N * ldf;
      const bool bias_here = db != nullptr && src == 0;  // db = colsum(dZ) rides along as a 16-column MMA against ones
      float* bias_ws = wsp + (size_t)splits * N * ldf;
      if (bias_here) { p.bias_out = bias_ws; p.bias_split_stride = N; }
      p.red = pending_;  // the previous launch's partials are summed by this launch's waiting epilogue warps
      pending_.ws = nullptr;
      NERF_TRY(tc_launch(p, true, dim3((unsigned)cdiv(N, 128), (unsigned)cdiv(K, BN), (unsigned)splits), st));
      ReduceJob job;
      job.ws = wsp; job.out = dW; job.ws2 = bias_here ? bias_ws : nullptr; job.out2 = db; job.stride = p.split_stride; job.stride2 = N;
      job.splits = splits; job.rows = N; job.cols = K; job.ldw = ldf; job.ldo = ldw; job.coff = coff; job.n2 = bias_here ? N : 0;
      // w16: the fp16 dZ planes carry the level's scale s and an activation plane may carry 32: [1/s, x/s] at wgrad_unscale_.
      // The bias gradient rides with src 0 and has no X factor, so it takes 1/s whatever x1 is.
      job.mul = w16_ ? wgrad_unscale_ + ((src == 0 && x1_act) ? 1 : 0) : nullptr;
      job.mul2 = w16_ ? wgrad_unscale_ : nullptr;
      if (defer_reduce_) pending_ = job;
      else NERF_TRY(launch_reduce_job(job, st));
    }
    return 0;
  }

  bool pair() const { return !(flags_ & NERF_FLAG_NO_WEIGHT_MULTICAST); }
  bool pair_mma() const { return pair() && (flags_ & NERF_FLAG_PAIR_MMA) != 0; }

  bool split_;
  bool w16_ = false;                  // fp32-accurate mode: wgrad operands as fp16 planes (see init)
  bool f8c_ = false;                  // fp32-accurate mode: fused forward with fp16 + E4M3 correction products (see init)
  std::vector<Plane> wf_;             // f8c: W16 / W8 planes of the trunk layers and the condition layer
  bool wf_dirty_ = true;              // ... stale since the last prepare()
  float act_x_scale_ = 1.0f;          // w16 + f8c: the activation planes hold 32 a
  float* dz_sc_ = nullptr;            // w16: per level [s, 1/s], then 2 scratch words of launch_dz_scale
  unsigned* dz_sc_scratch_ = nullptr;
  const float* wgrad_unscale_ = nullptr;  // w16: device scalar the wgrad reductions of the level being walked multiply by
  unsigned flags_ = 0;
  MlpShape s_;
  long max_rows_ = 0;
  int pos_pitch_ = 0, dir_pitch_ = 0;
  std::vector<Level> levels_;
  Plane scr_pos_, scr_dir_;  // rendering: encoding scratch of the in-kernel encoder warps (L2-resident)
  long scr_rows_ = 0;
  Plane dz_[2];
  std::vector<Plane> dzs_;  // bf16 fused dgrad chain: dZ of trunk layer D-1-j, kept for the wgrad GEMMs
  std::vector<Plane> wp_, wtp_;
  float* ws_ = nullptr;               // thin heads / column sums
  float* wsw_[2] = {nullptr, nullptr};  // wgrad partial tiles, alternating
  int ws_flip_ = 0;
  ReduceJob pending_{};               // reduction of the last wgrad launch, not yet run
  bool defer_reduce_ = true;
  float* fconsts_ = nullptr;  // fused forward: biases + head weights, gathered per parameter version
  bool fconsts_dirty_ = true;
  size_t bytes_ = 0;
  std::vector<void*> owned_;
};

}  // namespace

MlpEngine* make_tc_mlp(bool split3, unsigned engine_flags) { return new TcMlp(split3, engine_flags); }

}  // namespace nerf

// gemm_tc.cuh — interface of the tcgen05/TMEM GEMM used by the tensor-core MLP engine (mlp_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "encode_rows.cuh"
#include "kernels.cuh"

namespace nerf {

constexpr int TC_MAX_KB = 30;

// One launch = one output tile per CTA:  C(128 x BN) = sum over k-blocks of A_blk * B_blk^T, fp32 accumulate in TMEM.
// Operands are bf16 "planes" in global memory (x ~= hi [+ lo]) described by TMA tensor maps; a k-block names
// which A plane / B plane it multiplies, so conjoined inputs (two K segments) and the bf16x3 split passes
// (hi*hi + hi*lo + lo*hi) are just longer k-block lists over the same kernel.
// out[i*ldo + coff + j] += sum_z ws[z*stride + i*ldw + j]  (+ the bias segment out2[j] += sum_z ws2[z*stride2 + j]) in a
// FIXED order: the reduction of a wgrad launch's fp32 partial tiles.  ws == nullptr: no job.
struct ReduceJob {
  const float* ws;
  float* out;
  const float* ws2;
  float* out2;
  long stride, stride2;
  int splits, rows, cols, ldw, ldo, coff, n2;
  const float *mul, *mul2;  // optional device scalars: the sums (mul2: the bias segment) are multiplied by them — the power-of-two
                            // un-scaling of fp16 dZ / activation planes
};

struct alignas(64) TcParams {
  CUtensorMap maps[8];  // [0..3] A-side planes, [4..5] B-side planes, [6..7] output planes (K-major epilogue TMA stores)
  // ---- K-major mode (forward, dgrad): grid = (ceil(M/128), ceil(N/BN))
  struct KB { int8_t a, b; int16_t a_col, b_col; } kb[TC_MAX_KB];
  int n_kb;
  // ---- MN-major mode (wgrad): grid = (out_rows/128, ceil(out_cols/BN), splits); reduction over the rows of the planes
  int n_pass;
  int8_t pass_a[4], pass_b[4];
  int f16_ops;     // MN-major: the planes hold fp16 (one pass), not bf16
  int a_col0;      // column of the A plane where this launch's output rows start (usually 0)
  int split_len;   // reduction rows per split (multiple of 64)
  long red_len;    // total reduction length (rows of the planes = samples)
  // ---- common
  long M;          // K-major: rows of the output
  int BN;          // UMMA N = tile width (multiple of 16, <= 256)
  int n_valid;     // columns of the output that exist (stores are clipped to it)
  int rows_valid;  // MN-major: output rows that exist
  int n_stages;
  // ---- epilogue
  int epi;  // 0: forward (bias, activation, planes out); 1: dgrad (rank-1, relu mask, planes out); 2: fp32 partial
  const float* bias;
  int act;
  const float* r1;            // [M]
  const float* v1;            // [n_valid]
  const __nv_bfloat16* mask;  // [M, ld_mask] (> 0 keeps)
  int ld_mask;
  __nv_bfloat16 *out_hi, *out_lo;  // lo may be null (bf16 mode)
  int ld_out;
  float* out_f32;     // epi 2: [split][rows][ld_f32]
  int ld_f32;
  long split_stride;  // floats between splits
  float* bias_out;    // epi 2, optional: [split][rows] column sums of the A operand (bias gradient)
  long bias_split_stride;
  // epi 2, optional: the reduction of the PREVIOUS wgrad launch's partials, run by this launch's epilogue warps while
  // they wait for the MMA pipeline (they are idle until the last k-block): the reduction leaves the critical path
  ReduceJob red;
  // ReLU masks as bit planes: 1 bit per output element instead of re-reading the 2-byte activation in dgrad
  uint32_t* bits_out;         // epi 0, optional: [M, ld_bits] words, bit j of word c <=> Y[m, 32c + j] > 0
  const uint32_t* mask_bits;  // epi 1, optional (replaces `mask`)
  int ld_bits;
  // thin head fused into the forward epilogue (needs the whole row in one CTA: gridDim.y == 1)
  const float* head_w;  // [head_n, n_valid]
  const float* head_b;  // [head_n]
  float* head_out;      // [M, head_n]
  int head_n;           // 0..3
};

int tc_make_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows);  // memoised
int tc_encode_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows);
int tc_make_tmap_box(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_cols,
                     uint32_t box_rows);  // memoised; box_cols 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)
int tc_encode_tmap_box(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_cols,
                       uint32_t box_rows);
void tc_tmap_cache_clear();
int tc_launch(const TcParams& p, bool mn_major, dim3 grid, cudaStream_t st);
int tc_smem_bytes(int BN, int n_stages, bool mn_major);
int tc_pick_stages(int BN, int n_kblocks, bool mn_major);

// whole-MLP forward with TMEM-resident activations (mlp_fused.cu); bf16 planes only
int launch_mlp_fused_forward(const __nv_bfloat16* pos, int pos_pitch, const __nv_bfloat16* dir, int dir_pitch,
                             const __nv_bfloat16* const* wplanes, const int* kpad, const int* in_b, int D, int W, int Wc, long M,
                             const float* consts_dev, int n_consts, int head_d_off, int head_rgb_off, const int* bias_off,
                             float* raw_density, float* raw_rgb, __nv_bfloat16* const* act_out, uint32_t* const* bits_out,
                             const RaySource* rays, long enc_scratch_rows, bool pair, cudaStream_t st);
// act_out != nullptr: also write every layer's activations + ReLU bit planes.  rays != nullptr: the kernel's encoder warps
// build the encodings from the level's t-values (cast_rays + IPE + direction PE in-kernel) into `pos` / `dir`, which are
// the level's planes (enc_scratch_rows == 0) or an L2-resident scratch of enc_scratch_rows rows (rendering)

// backward dgrad chain of the trunk as one kernel (bf16 planes; mlp_fused.cu)
int launch_mlp_fused_dgrad(const __nv_bfloat16* dz_cond, int dz_cond_pitch, const __nv_bfloat16* const* wt, const int* wt_pitch, int D, int W,
                           int Wc, long M, const float* consts_dev, int n_consts, int head_d_off, const float* d_raw_density,
                           __nv_bfloat16* const* dz_out, const uint32_t* const* mask_bits, bool pair, cudaStream_t st);
// the same for the fp32-accurate split mode (hi/lo planes, three-term products; mlp_fused_split.cu)
int launch_mlp_fused_forward_split(const __nv_bfloat16* pos_hi, const __nv_bfloat16* pos_lo, int pos_pitch, const __nv_bfloat16* dir_hi,
                                   const __nv_bfloat16* dir_lo, int dir_pitch, const __nv_bfloat16* const* w_hi,
                                   const __nv_bfloat16* const* w_lo, const int* kpad, const int* in_b, int D, int W, int Wc, long M,
                                   const float* consts_dev, int n_consts, int head_d_off, int head_rgb_off, const int* bias_off,
                                   float* raw_density, float* raw_rgb, __nv_bfloat16* const* act_hi, __nv_bfloat16* const* act_lo,
                                   uint32_t* const* bits_out, const RaySource* rays, long enc_scratch_rows, bool pair, cudaStream_t st,
                                   bool act_f16 = false, bool fp8c = false, bool pair_mma = false);
// fp8c: the activations stay on chip as fp16 + two E4M3 correction planes and w_hi / w_lo are the matching W16 / W8 weight planes
// (launch_f32_to_f8c_planes): two MMA-equivalents per product instead of three (mlp_fused_split.cu, REP = 1)
struct F8cJobs {
  struct Job { const float* src; __nv_bfloat16 *p0, *p1; int rows, cols, enc_from, kpad; } job[16];
  int n;
};
int launch_f32_to_f8c_planes(const F8cJobs& jobs, cudaStream_t st);
// act_f16 (training): act_hi[s] is ONE fp16 plane per layer (act_lo unused) — the X operands of the fp16 wgrad GEMMs
// pair: run as 2-CTA clusters sharing every weight stage by TMA multicast (halves the L2 weight traffic; mlp_fused_split.cu)

int launch_mlp_fused_dgrad_split(const __nv_bfloat16* dz_cond_hi, const __nv_bfloat16* dz_cond_lo, int dz_cond_pitch,
                                 const __nv_bfloat16* const* wt_hi, const __nv_bfloat16* const* wt_lo, const int* wt_pitch, int D, int W,
                                 int Wc, long M, const float* consts_dev, int n_consts, int head_d_off, const float* d_raw_density,
                                 __nv_bfloat16* const* dz_out_hi, __nv_bfloat16* const* dz_out_lo, const uint32_t* const* mask_bits,
                                 bool pair, cudaStream_t st, const float* dz_scale_f16 = nullptr, bool pair_mma = false);
// pair_mma (with pair): the 2-CTA clusters run tcgen05.mma.cta_group::2 — each SM holds half of every weight tile (mlp_fused_split.cu, PM)
// dz_scale_f16 != nullptr: dz_out_hi[s] is ONE fp16 plane that receives dZ * *dz_scale_f16 (device scalar, launch_dz_scale)

// helpers on bf16 planes ------------------------------------------------------------------------------
// fp32 [rows, cols] (pitch src_pitch) -> hi/lo planes (pitch dst_pitch, zero padded); transpose writes dst[c, r]
int launch_f32_to_planes(const float* src, int src_pitch, long rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo,
                         int dst_pitch, int dst_cols, bool transpose, long dst_rows_t, cudaStream_t st);
// the same conversion for a list of tensors in one launch (all weight planes of the network once per step)
struct PlaneJobs {
  struct Job {
    const float* src; __nv_bfloat16 *hi, *lo;
    long rows, drows_t, n;  // n = elements written: transpose ? drows_t * dcols : rows * dcols
    int sp, cols, dp, dcols, transpose;
  } job[32];
  int n;
};
int launch_f32_to_planes_batch(const PlaneJobs& jobs, cudaStream_t st);
struct GatherJobs {
  struct Job { long src; int dst, n; } job[24];
  int n;
};
int launch_gather_f32(const float* src, float* dst, const GatherJobs& jobs, cudaStream_t st);  // dst[job.dst + i] = src[job.src + i]
// heads on planes (N <= 4)
int launch_thin_fwd_planes(const __nv_bfloat16* xh, const __nv_bfloat16* xl, int ldx, const float* W, const float* b,
                           float* Y, long M, int N, int K, cudaStream_t st);
// o16 (optional): a third output plane fp16(dX * *scale16), same pitch — the dZ operand of the fp16 wgrad GEMMs
int launch_thin_dgrad_planes(const float* dZ, const float* W, long M, int N, int K, const uint32_t* mask_bits, int ld_bits,
                             __nv_bfloat16* oh, __nv_bfloat16* ol, int ldo, cudaStream_t st, void* o16 = nullptr,
                             const float* scale16 = nullptr);
// x_f16: xh is one fp16 plane (xl ignored)
int launch_thin_wgrad_planes(const float* dZ, const __nv_bfloat16* xh, const __nv_bfloat16* xl, int ldx, float* dW, float* db,
                             long M, int N, int K, float* workspace, cudaStream_t st, bool x_f16 = false, float x_mul = 1.0f);
// Device scalars for the fp16 dZ planes of one level: out[0] = s = 2^-floor(log2(max |d_raw|)) (so that the largest head gradient
// lands in [1, 2)), out[1] = 1 / s, out[2] = act_scale / s (act_scale: the power of two the level's activation planes carry).
// scratch: 2 zero-initialised words (left zeroed).  One launch, no host round trip.
int launch_dz_scale(const float* d_raw_rgb, const float* d_raw_density, long M, float act_scale, float* out, unsigned* scratch, cudaStream_t st);
// fp32 [rows, cols] (pitch src_pitch) -> one fp16 plane (pitch dst_pitch, zero padded up to dst_cols)
int launch_f32_to_f16_plane(const float* src, int src_pitch, long rows, int cols, void* dst, int dst_pitch, int dst_cols, cudaStream_t st);
// db[n] += sum_m (hi + lo)[m, n]
int launch_colsum_planes(const __nv_bfloat16* h, const __nv_bfloat16* l, int ld, long M, int N, float* db, float* workspace,
                         cudaStream_t st);
// out[i*ldo + coff + j] += sum_z ws[z*stride + i*ldw + j]
// the same plus a second segment out2[j] += sum_z ws2[z * stride2 + j] (bias partials) in the same launch
int launch_reduce_partials2(const float* ws, int splits, long split_stride, int rows, int cols, int ldw, float* out, int ldo, int coff,
                            const float* ws2, int n2, long stride2, float* out2, cudaStream_t st);
int launch_reduce_partials(const float* ws, int splits, long split_stride, int rows, int cols, int ldw, float* out, int ldo,
                           int coff, cudaStream_t st);
int launch_reduce_job(const ReduceJob& job, cudaStream_t st);  // stand-alone launch of a job (same arithmetic, same order)

}  // namespace nerf

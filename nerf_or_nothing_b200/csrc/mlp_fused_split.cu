// mlp_fused_split.cu — the MipNeRF MLP forward as ONE persistent tcgen05 kernel in the fp32-accurate
// tensor-core mode (NERF_PRECISION_FP32_TC): every tensor is a pair of bf16 planes x = hi + lo and every product is the
// three-term split hi*hi + lo*hi + hi*lo accumulated in fp32.
//
// Replaces the per-layer launch chain of AcceleratedMLP::get_output (ANU/AcceleratedMLP.cpp:214-255).  In this mode the
// layer-by-layer GEMMs are HBM-bound (each layer reads the previous activations' two planes and writes its own);
// here the activations never come back from HBM: per 128-row tile, tensor memory (512 columns) holds
//   ACT_hi [0,128)  ACT_lo [128,256)   the current layer's input as the A operand of ".ts" MMAs (256 bf16 per row, each)
//   ACC    [256,512)                   the layer's fp32 accumulator: two N-halves of 128 columns
// and the only global traffic per layer is the (training-only) write of the layer's activation planes + ReLU bits.
// A layer is two N-halves; per half and 64-wide k-block one 32 KB ring stage brings W_hi|W_lo [128 x 64] and feeds the
// MMAs A_hi*W_hi, A_lo*W_hi, A_hi*W_lo (128x128x16 each).  The eight epilogue warps (two per TMEM lane quarter) apply
// bias + ReLU, split into hi/lo, write both planes back into ACT in place once the layer's MMAs are complete, and
// (training) ship them through 64-byte-swizzled [32 x 32] boxes with their own TMA stores.  The encodings of layer 0,
// the skip layer and the condition layer stream through the same ring as shared-memory A operands.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstring>

#include "encode_rows.cuh"
#include "gemm_tc.cuh"
#include "sm100.cuh"

namespace nerf {
#ifdef NERF_SPLIT_DBG
// debug builds only (scripts/split_stamps.py): clock64 stamps of epilogue warp 0 / MMA warp of CTA 0 for its 4th tile
__device__ unsigned long long g_split_dbg[16 * 2 * 8];
__device__ unsigned long long g_split_dbg_mma[16 * 2 * 4];
#define DBG_STAMP(cond, s_, h_, i_) do { if ((cond) && lane == 0) g_split_dbg[((s_) * 2 + (h_)) * 8 + (i_)] = clock64(); } while (0)
#define DBG_STAMP_MMA(cond, s_, h_, i_) do { if ((cond) && leader) g_split_dbg_mma[((s_) * 2 + (h_)) * 4 + (i_)] = clock64(); } while (0)
#else
#define DBG_STAMP(cond, s_, h_, i_) do { } while (0)
#define DBG_STAMP_MMA(cond, s_, h_, i_) do { } while (0)
#endif

namespace {

using namespace sm100;

constexpr int kThreadsS = 320;            // 8 epilogue warps + TMA producer + MMA issuer
constexpr int kEncWarpsS = 2;             // forward kernels: + 2 encoder warps (cast_rays + IPE + direction PE in-kernel)
constexpr int kThreadsSE = kThreadsS + 32 * kEncWarpsS;
constexpr int kEpiWarps = 8;
constexpr int kNSInfer = 6, kNSTrain = 5;  // operand ring: 32 KB stages (a weight plane tile [<=256 x 64], or an encoding k-block's hi|lo tiles)
#ifndef NERF_DOUBLE_BOX_SPLIT
#define NERF_DOUBLE_BOX_SPLIT 0
#endif
// hi/lo planes out (not F16): two alternating box PAIRS per warp (8 KB slots) paid for with one ring stage.  Measured (r02ab2, A/B in
// one call): forward 3.39 vs 3.33 ms, dgrad chain 2.83 vs 2.86, step 10.86 vs 10.87 — no gain, compiled out
constexpr bool kDoubleBoxSplit = NERF_DOUBLE_BOX_SPLIT != 0;
__host__ __device__ constexpr int train_stages(bool f16) { return (!f16 && kDoubleBoxSplit) ? kNSTrain - 1 : kNSTrain; }
__host__ __device__ constexpr int train_slot_bytes(bool f16) { return (!f16 && kDoubleBoxSplit) ? 8192 : 4096; }
constexpr int kStageB = 256 * 128;        // 32 KB
constexpr int kMaxStepsS = 12;
#ifndef NERF_LATE_SHIP
#define NERF_LATE_SHIP 1
#endif
#ifndef NERF_DOUBLE_BOX
#define NERF_DOUBLE_BOX 1
#endif
constexpr bool kDoubleBox = NERF_DOUBLE_BOX != 0;  // F16 kernels: two alternating store boxes per warp
constexpr bool kLateShip = NERF_LATE_SHIP != 0;  // second-half epilogue: TMA-store the chunks after act_ready instead of between them

__device__ __forceinline__ uint32_t pack2s(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace

struct alignas(64) SplitParams {
  CUtensorMap map_pos[2], map_dir[2];         // [hi, lo] encodings [M, 128] / [M, 64], box {64, 128}
  CUtensorMap map_w[kMaxStepsS][2];           // [hi, lo] weight planes [N, Kpad], box {64, 128}
  CUtensorMap map_w64[kMaxStepsS][2];         // the same planes, box {64, 64}: pair-MMA kernels, each CTA of a pair fetches 64 rows of a tile
  CUtensorMap map_act[kMaxStepsS][2];         // training: [hi, lo] activation planes [M, N], box {32, 32} (SWIZZLE_64B)
  uint32_t* bits[kMaxStepsS];                 // training: ReLU bit planes [M, N/32]
  struct Step {
    int16_t n_act_kb, enc_kind, n_enc_kb, n_halves;  // n_halves = ceil(N / 128): a 64-wide layer runs as one half whose upper
                                                     // 64 weight rows are the TMA's out-of-bounds zeros
    int16_t produces, head;
    int16_t n_cols;                                  // N: output columns that exist (64, 128 or 256)
    int32_t bias_off;
  } steps[kMaxStepsS];
  int n_steps;
  long M;
  const float* consts;
  int n_consts, head_d_off, head_rgb_off;
  float* raw_density;
  float* raw_rgb;
  const float* r1;  // dgrad chain: [M] dL/d raw_density (rank-1 term of step 0); bits[] are then READ as ReLU masks
  // F16 kernels (the wgrad GEMMs multiply fp16 planes, see mlp_tc.cu): map_act[s][0] views ONE fp16 plane per layer; the dgrad
  // chain stores fp16(dZ * *dz_scale), a power of two chosen per level from the head gradients (launch_dz_scale)
  const float* dz_scale;
  int pair_mma;  // host: launch the PM kernels (cta_group::2 MMAs; needs the 2-CTA clusters)
  // forward only — in-kernel cast_rays + IPE + direction PE (see FusedParams::enc_mode in mlp_fused.cu): 0 = planes written by
  // another kernel, 1 = the encoder warps write the level's hi/lo planes, 2 = a per-CTA double-buffered scratch (128 rows each)
  int enc_mode;
  RaySource rs;
  __nv_bfloat16 *enc_pos[2], *enc_dir[2];  // [hi, lo]
};

namespace {

// bias + ReLU + hi/lo split of one 32-column chunk; optional ReLU mask (bit j = column j passed) and head FMAs
// F16: also the chunk as 16 packed fp16 pairs (fw) — what the training forward stores for the wgrad GEMMs in that mode
template <bool BITS, bool F16>
__device__ __forceinline__ uint32_t split_chunk(const uint32_t (&r)[32], const float* bias, int head_n, const float* head_w, float (&head)[3],
                                                uint32_t* hw, uint32_t* lw, uint32_t* fw) {
  float x[32];
  const float4* bv = reinterpret_cast<const float4*>(bias);
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const float4 bb = bv[q];
    x[4 * q] = __uint_as_float(r[4 * q]) + bb.x; x[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bb.y;
    x[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bb.z; x[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bb.w;
  }
  uint32_t mask = 0u;
  if (BITS) {  // from the sign bits: four chains of eight funnel shifts (m = m << 1 | sign)
    uint32_t m[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int g = 0; g < 4; g++)
#pragma unroll
      for (int e = 7; e >= 0; e--) m[g] = __funnelshift_l(__float_as_uint(x[8 * g + e]), m[g], 1);
    mask = ~__byte_perm(__byte_perm(m[0], m[1], 0x0040), __byte_perm(m[2], m[3], 0x0040), 0x5410);
  }
#pragma unroll
  for (int j = 0; j < 32; j++) x[j] = fmaxf(x[j], 0.f);
  if (head_n) {  // 1: density head; 3: rgb head
#pragma unroll
    for (int n = 0; n < 3; n++)
      if (n < head_n) {
        const float4* hv = reinterpret_cast<const float4*>(head_w + n * 128);
        float a = head[n];
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const float4 w = hv[q];
          a = fmaf(x[4 * q], w.x, a); a = fmaf(x[4 * q + 1], w.y, a);
          a = fmaf(x[4 * q + 2], w.z, a); a = fmaf(x[4 * q + 3], w.w, a);
        }
        head[n] = a;
      }
  }
  // x = hi + lo: hi = bf16(x) in pairs, lo = bf16(x - hi) with hi rebuilt from the packed bits
#pragma unroll
  for (int q = 0; q < 16; q++) {
    hw[q] = pack2s(x[2 * q], x[2 * q + 1]);
    lw[q] = pack2s(x[2 * q] - __uint_as_float(hw[q] << 16), x[2 * q + 1] - __uint_as_float(hw[q] & 0xFFFF0000u));
    if (F16) fw[q] = pack_f16x2_sat(x[2 * q], x[2 * q + 1]);
  }
  return mask;
}

// dgrad chain: one 32-column chunk of dX = dZ W (+ r1 v1^T), masked by the ReLU bits of the layer below, split into hi/lo
template <bool F16>
__device__ __forceinline__ void dgrad_split_chunk(const uint32_t (&r)[32], uint32_t mask, float r1, const float* v1, uint32_t* hw,
                                                  uint32_t* lw, uint32_t* fw, float s16) {
#pragma unroll
  for (int q = 0; q < 16; q++) {
    float x0 = __uint_as_float(r[2 * q]), x1 = __uint_as_float(r[2 * q + 1]);
    if (v1) { x0 = fmaf(r1, v1[2 * q], x0); x1 = fmaf(r1, v1[2 * q + 1], x1); }
    x0 = ((mask >> (2 * q)) & 1u) ? x0 : 0.f;
    x1 = ((mask >> (2 * q + 1)) & 1u) ? x1 : 0.f;
    hw[q] = pack2s(x0, x1);
    lw[q] = pack2s(x0 - __uint_as_float(hw[q] << 16), x1 - __uint_as_float(hw[q] & 0xFFFF0000u));
    if (F16) fw[q] = pack_f16x2_sat(x0 * s16, x1 * s16);
  }
}

// REP = 1, the "fp16 + fp8 corrections" representation of the same three-term product (forward kernels):
//   a w ~= ah wh + al wh + ah wl   with ah = fp16(a), al = a - ah (|al| <= 2^-12 |a|), likewise w.
// The two correction terms are 2^-12 of the product, so 4 significant bits of their factors keep the total error at 2^-16: they
// run as E4M3 MMAs (kind::f8f6f4, K = 32 per instruction = twice the bf16 rate) onto the SAME accumulator — two MMA-equivalents
// per product instead of three.  E4M3 spans 2^-9 .. 448 only, so every factor carries a power of two and the accumulator holds
// 2^15 x the true sum (undone by the epilogue's FFMA):
//   tensor memory  ACT16 [0,128)  fp16(32 a), two per column       x  W16 = fp16(2^10 w)                       (kind::f16)
//                  ACT8 [128,256) per 64-wide k-block 16 columns E4M3(2^9 al), then 16 columns E4M3(ah)
//                                                                  x  W8 = per k-block 64 bytes E4M3(2^6 wh), 64 bytes E4M3(2^15 wl)
// and the encoding k-blocks (shared-memory A operands, bf16 hi/lo as before) multiply weight columns stored as bf16 hi/lo of
// 2^15 w.  A ring stage is still 32 KB: the W16 tile [128 x 64] and the W8 tile [128 x 128 bytes] of the (half, k-block).
// One 32-column chunk: t = 32 relu(z) -> 16 fp16 pairs (hw), 8 words E4M3(2^9 al) (l8), 8 words E4M3(ah) (h8).
template <bool BITS>
__device__ __forceinline__ uint32_t f8c_chunk(const uint32_t (&r)[32], const float* bias32, int head_n, const float* head_w, float (&head)[3],
                                              uint32_t* hw, uint32_t* l8, uint32_t* h8) {
  float t[32];
  const float4* bv = reinterpret_cast<const float4*>(bias32);
  constexpr float kAcc = 1.0f / 1024.0f;  // accumulator = 2^15 z; t = 32 z + 32 b
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const float4 bb = bv[q];
    t[4 * q] = fmaf(__uint_as_float(r[4 * q]), kAcc, bb.x); t[4 * q + 1] = fmaf(__uint_as_float(r[4 * q + 1]), kAcc, bb.y);
    t[4 * q + 2] = fmaf(__uint_as_float(r[4 * q + 2]), kAcc, bb.z); t[4 * q + 3] = fmaf(__uint_as_float(r[4 * q + 3]), kAcc, bb.w);
  }
  uint32_t mask = 0u;
  if (BITS) {
    uint32_t m[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int g = 0; g < 4; g++)
#pragma unroll
      for (int e = 7; e >= 0; e--) m[g] = __funnelshift_l(__float_as_uint(t[8 * g + e]), m[g], 1);
    mask = ~__byte_perm(__byte_perm(m[0], m[1], 0x0040), __byte_perm(m[2], m[3], 0x0040), 0x5410);
  }
#pragma unroll
  for (int j = 0; j < 32; j++) t[j] = fmaxf(t[j], 0.f);
  if (head_n) {  // on 32 a: the caller scales the head sums by 2^-5
#pragma unroll
    for (int n = 0; n < 3; n++)
      if (n < head_n) {
        const float4* hv = reinterpret_cast<const float4*>(head_w + n * 128);
        float a = head[n];
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const float4 w = hv[q];
          a = fmaf(t[4 * q], w.x, a); a = fmaf(t[4 * q + 1], w.y, a);
          a = fmaf(t[4 * q + 2], w.z, a); a = fmaf(t[4 * q + 3], w.w, a);
        }
        head[n] = a;
      }
  }
  const __half2 k32nd = __floats2half2_rn(0.03125f, 0.03125f);
#pragma unroll
  for (int q = 0; q < 8; q++) {  // four values per E4M3 word
    const uint32_t h0 = pack_f16x2_sat(t[4 * q], t[4 * q + 1]), h1 = pack_f16x2_sat(t[4 * q + 2], t[4 * q + 3]);
    hw[2 * q] = h0; hw[2 * q + 1] = h1;
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h0)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
    l8[q] = pack_e4m3x2_f32((t[4 * q] - f0.x) * 16.f, (t[4 * q + 1] - f0.y) * 16.f) |
            (pack_e4m3x2_f32((t[4 * q + 2] - f1.x) * 16.f, (t[4 * q + 3] - f1.y) * 16.f) << 16);
    const __half2 g0 = __hmul2(*reinterpret_cast<const __half2*>(&h0), k32nd), g1 = __hmul2(*reinterpret_cast<const __half2*>(&h1), k32nd);
    h8[q] = pack_e4m3x2_f16x2(*reinterpret_cast<const uint32_t*>(&g0)) | (pack_e4m3x2_f16x2(*reinterpret_cast<const uint32_t*>(&g1)) << 16);
  }
  return mask;
}

// One CTA walks 128-row tiles.  A layer is two N-halves with their own accumulator columns: while the MMAs of the second
// half run, the eight epilogue warps (TMEM lane quarter = warp % 4, 64 of the half's 128 columns each) finish the first
// half into registers; its hi/lo words are written into ACT — in place — only after the second half's MMAs have read
// ACT, together with the second half's.
// MODE 0: inference forward; 1: training forward (activation planes + ReLU bits written); 2: backward dgrad chain (step 0:
// A = dZ of the condition layer streamed through the ring like an encoding, epilogue = density-head rank-1 term + ReLU
// mask of the layer below; every step's dZ planes written for the wgrad GEMMs).
// Forward modes run two more warps (10, 11), the ENCODERS: one thread per sample row turns the tile's t-values + ray into
// the hi/lo encoding rows (encode_rows.cuh), stored through L2; enc_ready[tile & 1] hands them to the producer's TMA loads,
// enc_free[tile & 1] (MMA issuer, end of tile) returns the scratch buffer.  They run up to two tiles ahead of the MMAs.
//
// CL = 2: the kernel runs as thread-block CLUSTERS of two CTAs that share every weight stage.  Measured (ncu, round 1) the
// CL = 1 kernels move ~5300 B/cycle through L2 — 2.3 MB of hi+lo weights per 128-row tile against 6144 cycles of MMAs per layer
// — which is the chip's L2 throughput ceiling (~6300 B/cycle, B300_MICROARCH.md), not the tensor pipe: that is why the
// training forward and the dgrad chain both sat at 55 % tensor-pipe activity.  Every CTA walks the SAME weight sequence
// (weights do not depend on the tile), so the two rings run in lock-step slot for slot: rank 0 fetches the W_hi box of a stage,
// rank 1 the W_lo box, each with `.multicast::cluster` into both CTAs' rings — one L2 read feeds two SMs, weight traffic
// halves.  A slot is recycled when BOTH CTAs' MMAs have consumed it (tcgen05.commit multicast onto w_empty, count CL).
// Encoding stages stay per-CTA (own tile).  Every CTA of the grid walks the same NUMBER of tiles (phantom tiles past the end
// load zeros and store nothing) so that the rings never diverge.
// F16 (training modes): what leaves the SM for the wgrad GEMMs is ONE fp16 plane per layer (activations, or dZ times the level's
// power-of-two scale) instead of the hi + lo bf16 planes — half the store traffic here, half the read traffic and one MMA
// instead of three there.  Everything on chip (ACT_hi | ACT_lo, the three-term products) is unchanged.
// PM (pair MMA, CL = 2): the cluster's rank-0 CTA issues every MMA for BOTH tiles with tcgen05.mma.cta_group::2 (M = 256: its own 128
// rows and the peer's, each CTA's A operand and accumulator in its own tensor memory), and each CTA holds only ITS 64 rows of a
// stage's [128 x 64] weight tile — the tensor core reads the two halves from the two shared memories.  Every SM then pulls
// half the weight bytes through L2 (the CL = 2 multicast kernels still deliver the whole tile to each SM; in-kernel stamps
// showed their MMAs paced by that stream, 83 cycles per MMA instead of 64).  Protocol differences: a stage's TMA loads of both
// CTAs signal the LEADER's w_full (cp.async.bulk.tensor ... cta_group::2); the leader's commits release the slot and publish
// the accumulators in both CTAs (multicast); the peer's epilogue warps arrive on the leader's acc_empty / act_lo_ready / act_ready
// (counts doubled) through the cluster address; the leader returns the peer's encoding scratch buffer too.
template <int MODE, int CL, bool F16, int REP = 0, bool PM = false>
__global__ void __launch_bounds__(MODE == 2 ? kThreadsS : kThreadsSE, 1) k_mlp_fused_split(const __grid_constant__ SplitParams p) {
  static_assert(CL == 1 || CL == 2, "cluster of 1 or 2 CTAs");
  static_assert(!PM || CL == 2, "pair MMA needs the 2-CTA cluster");
  static_assert(!F16 || MODE != 0, "the fp16 planes exist in the training kernels only");
  static_assert(REP == 0 || MODE == 0 || (MODE == 1 && F16), "fp8 corrections: inference forward, or training forward with fp16 planes out");
  constexpr bool TRAIN = MODE != 0;
  constexpr bool DGRAD = MODE == 2;
  constexpr int NS = TRAIN ? train_stages(F16) : kNSInfer;
  constexpr int kSlot = train_slot_bytes(F16);
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_full[NS], w_empty[NS], acc_full[2], acc_empty[2], act_ready, act_lo_ready, enc_ready[2], enc_free[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float head_part[128][4];  // partial head dot products of the upper-column warps

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* w_ring = smem;                         // NS x 32 KB
  uint8_t* stage_buf = smem + NS * kStageB;       // TRAIN: 8 x 4 KB
  float* s_const = reinterpret_cast<float*>(stage_buf + (TRAIN ? kEpiWarps * kSlot : 0));

  const int n_tiles = (int)((p.M + 127) / 128);
  const int tiles_per_cta = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;  // the same for every CTA: rings stay in lock-step
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kAllCtas = (uint16_t)((1u << CL) - 1u);

  if (threadIdx.x == 0) {
    constexpr int kEpiArrivals = PM ? 2 * kEpiWarps : kEpiWarps;  // PM: the leader's MMA warp waits for both CTAs' epilogues
    for (int s = 0; s < NS; s++) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], PM ? 1 : CL); }
    for (int h = 0; h < 2; h++) { mbar_init(&acc_full[h], 1); mbar_init(&acc_empty[h], kEpiArrivals); }
    mbar_init(&act_ready, kEpiArrivals);
    mbar_init(&act_lo_ready, kEpiArrivals);
    for (int b = 0; b < 2; b++) { mbar_init(&enc_ready[b], kEncWarpsS); mbar_init(&enc_free[b], 1); }
    fence_barrier_init();
  }
  // REP = 1: the biases (everything before the density head's weights) are staged as 32 b, see f8c_chunk
  for (int i = threadIdx.x; i < p.n_consts; i += blockDim.x) s_const[i] = __ldg(p.consts + i) * ((REP == 1 && i < p.head_d_off) ? 32.f : 1.f);
  if (warp == kEpiWarps) {
    if (PM) tmem_alloc_pair<512>(&tmem_base_smem);
    else tmem_alloc<512>(&tmem_base_smem);
    if (lane == 0) {
      for (int q = 0; q < 2; q++) { prefetch_tmap(&p.map_pos[q]); if (!DGRAD) prefetch_tmap(&p.map_dir[q]); }
      for (int s = 0; s < p.n_steps; s++)
        for (int q = 0; q < 2; q++) { prefetch_tmap(PM ? &p.map_w64[s][q] : &p.map_w[s][q]); if (TRAIN && (q == 0 || !F16)) prefetch_tmap(&p.map_act[s][q]); }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast at them
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t ACT_HI = tmem_base, ACT_LO = tmem_base + 128, ACC = tmem_base + 256;  // ACC half h at + 128 h

  if (warp == kEpiWarps) {
    // ------------------------------------------------------------------ TMA producer: everything goes through ONE ring of
    // 32 KB stages — the hi|lo tiles [128 x 64] of a weight half-layer k-block, or of an encoding k-block (A operand)
    if (lane == 0) {
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; (int)tl < tiles_per_cta; tile += gridDim.x, tl++) {
        // rows of this tile's encodings in the tensors the maps view: the sample index, or this CTA's scratch buffer tl & 1
        const int row0 = (!DGRAD && p.enc_mode == 2) ? (blockIdx.x * 2 + (int)(tl & 1)) * 128 : tile * 128;
        if (!DGRAD && p.enc_mode) mbar_wait(&enc_ready[tl & 1], (tl >> 1) & 1);  // the encoder warps have written them
        for (int s = 0; s < p.n_steps; s++) {
          const SplitParams::Step st = p.steps[s];
          const int n_kb = st.n_act_kb + st.n_enc_kb;
          for (int h = 0; h < st.n_halves; h++)
            for (int kb = 0; kb < n_kb; kb++) {
              if (kb >= st.n_act_kb) {  // encoding k-block: A_hi | A_lo
                const int ws = it % NS;
                mbar_wait(&w_empty[ws], ((it / NS) & 1) ^ 1);
                const CUtensorMap* me = st.enc_kind == 2 ? p.map_dir : p.map_pos;
                const int ecol = (kb - st.n_act_kb) * 64;
                if (PM) {  // each CTA its own rows; all four boxes of the pair complete on the leader's barrier
                  if (cta_rank == 0) mbar_arrive_expect_tx(&w_full[ws], 4 * 16384);
                  const uint32_t lb = map_to_cta(&w_full[ws], 0);
                  tma_load_2d_pair(w_ring + (size_t)ws * kStageB, &me[0], ecol, row0, lb);
                  tma_load_2d_pair(w_ring + (size_t)ws * kStageB + 16384, &me[1], ecol, row0, lb);
                } else {
                  mbar_arrive_expect_tx(&w_full[ws], 2 * 16384);
                  tma_load_2d(w_ring + (size_t)ws * kStageB, &me[0], ecol, row0, &w_full[ws]);
                  tma_load_2d(w_ring + (size_t)ws * kStageB + 16384, &me[1], ecol, row0, &w_full[ws]);
                }
                it++;
              }
              const int ws = it % NS;  // W_hi | W_lo rows [128 h, 128 h + 128) of this k-block
              mbar_wait(&w_empty[ws], ((it / NS) & 1) ^ 1);   // every CTA of the cluster has consumed this slot
              if (PM) {  // this CTA's 64 rows of W_hi | W_lo (8 KB each); the pair's four boxes complete on the leader's barrier
                if (cta_rank == 0) mbar_arrive_expect_tx(&w_full[ws], 4 * 8192);
                const uint32_t lb = map_to_cta(&w_full[ws], 0);
                tma_load_2d_pair(w_ring + (size_t)ws * kStageB, &p.map_w64[s][0], kb * 64, h * 128 + (int)cta_rank * 64, lb);
                tma_load_2d_pair(w_ring + (size_t)ws * kStageB + 8192, &p.map_w64[s][1], kb * 64, h * 128 + (int)cta_rank * 64, lb);
                it++;
                continue;
              }
              mbar_arrive_expect_tx(&w_full[ws], 2 * 16384);  // both boxes land here, whoever fetches them
              if (CL == 1) {
                tma_load_2d(w_ring + (size_t)ws * kStageB, &p.map_w[s][0], kb * 64, h * 128, &w_full[ws]);
                tma_load_2d(w_ring + (size_t)ws * kStageB + 16384, &p.map_w[s][1], kb * 64, h * 128, &w_full[ws]);
              } else {  // this CTA fetches one of the two boxes for the whole cluster
                tma_load_2d_multicast(w_ring + (size_t)ws * kStageB + cta_rank * 16384, &p.map_w[s][cta_rank], kb * 64, h * 128, &w_full[ws],
                                      kAllCtas);
              }
              it++;
            }
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ------------------------------------------------------------------ MMA issuer (uniform warp, one elected lane issues)
    const bool leader = elect_one() && (!PM || cta_rank == 0);  // PM: the peer's MMA warp issues nothing
    const uint64_t desc0 = make_smem_desc(0, 16, 1024);
    const uint32_t ring_base = smem_u32(w_ring);
    constexpr int kM = PM ? 256 : 128;                 // PM: one MMA covers both CTAs' tiles
    constexpr uint32_t kLoOff = PM ? 8192 : 16384;     // W_lo (W8) tile behind W_hi (W16) in a stage: 64 or 128 rows of 128 bytes
    const uint32_t idesc = make_idesc_bf16(kM, 128, false, false);
    const uint32_t idesc_h = make_idesc_f16(kM, 128, false, false), idesc_8 = make_idesc_e4m3(kM, 128);
    uint32_t it = 0, n_acc[2] = {0, 0}, n_act = 0, tl = 0;
    // a ring slot is released in EVERY CTA of the cluster: the peer may multicast into this CTA's slot only when both are done
    auto release = [&](uint64_t* bar) {
      if (PM) umma_pair_commit(bar, kAllCtas);
      else if (CL == 1) umma_commit(bar);
      else umma_commit_multicast(bar, kAllCtas);
    };
    auto mma_ts = [&](uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t accf) { if (PM) umma_pair_ts(d, a, b, id, accf); else umma_bf16_ts(d, a, b, id, accf); };
    auto mma_ts8 = [&](uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t accf) { if (PM) umma_pair_f8_ts(d, a, b, id, accf); else umma_f8_ts(d, a, b, id, accf); };
    auto mma_ss = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t accf) { if (PM) umma_pair_ss(d, a, b, id, accf); else umma_bf16(d, a, b, id, accf); };
    auto wait_epi = [&](uint64_t* bar, uint32_t parity) { if (PM) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity); };  // PM: the peer's epilogue arrives too
    for (; (int)tl < tiles_per_cta && (!PM || cta_rank == 0); tl++) {
      for (int s = 0; s < p.n_steps; s++) {
        const SplitParams::Step st = p.steps[s];
        const int n_kb = st.n_act_kb + st.n_enc_kb;
        for (int h = 0; h < st.n_halves; h++) {
          wait_epi(&acc_empty[h], (n_acc[h] & 1) ^ 1);
          n_acc[h]++;
          // the layer below rewrites ACT in two instalments: its first N-half (= this layer's k-blocks 0,1; parked in
          // registers during its second half's MMAs) lands as soon as those MMAs are complete, its second N-half when that
          // half's epilogue is done — this layer's first k-blocks run under that epilogue instead of after it
          const bool wait_act = h == 0 && st.n_act_kb > 0;
          if (wait_act) wait_epi(&act_lo_ready, n_act & 1);
          tc_fence_after_sync();
          const uint32_t acc = ACC + 128 * h;
          DBG_STAMP_MMA(blockIdx.x == 0 && tl == 3 && MODE == 1, s, h, 0);
          for (int kb = 0; kb < n_kb; kb++) {
            const bool from_act = kb < st.n_act_kb;
            if (wait_act && kb == st.n_act_kb / 2) { DBG_STAMP_MMA(blockIdx.x == 0 && tl == 3 && MODE == 1, s, h, 1); wait_epi(&act_ready, n_act & 1); tc_fence_after_sync(); DBG_STAMP_MMA(blockIdx.x == 0 && tl == 3 && MODE == 1, s, h, 2); }
            uint32_t a_stage = 0, a_hi = 0;
            if (!from_act) {  // the encoding tiles of this k-block arrive in their own stage
              a_stage = it % NS;
              mbar_wait(&w_full[a_stage], (it / NS) & 1);
              a_hi = ring_base + a_stage * kStageB;
              it++;
            }
            const uint32_t ws = it % NS;
            mbar_wait(&w_full[ws], (it / NS) & 1);
            tc_fence_after_sync();
            const uint64_t dbh = desc0 + ((ring_base + ws * kStageB) >> 4), dbl = dbh + (kLoOff >> 4);
            if (leader) {
              if (from_act && REP == 1) {  // fp16 main product, then the two E4M3 correction products (K = 32 each)
#pragma unroll
                for (int k = 0; k < 4; k++) mma_ts(acc, ACT_HI + kb * 32 + k * 8, dbh + 2 * k, idesc_h, (kb | k) ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; k++) mma_ts8(acc, ACT_LO + kb * 32 + k * 8, dbl + 2 * k, idesc_8, 1u);
              } else if (from_act) {
#pragma unroll
                for (int k = 0; k < 4; k++) {  // hi*hi + lo*hi + hi*lo
                  mma_ts(acc, ACT_HI + kb * 32 + k * 8, dbh + 2 * k, idesc, (kb | k) ? 1u : 0u);
                  mma_ts(acc, ACT_LO + kb * 32 + k * 8, dbh + 2 * k, idesc, 1u);
                  mma_ts(acc, ACT_HI + kb * 32 + k * 8, dbl + 2 * k, idesc, 1u);
                }
              } else {
                const uint64_t dah = desc0 + (a_hi >> 4), dal = dah + (16384 >> 4);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                  mma_ss(acc, dah + 2 * k, dbh + 2 * k, idesc, (kb | k) ? 1u : 0u);
                  mma_ss(acc, dal + 2 * k, dbh + 2 * k, idesc, 1u);
                  mma_ss(acc, dah + 2 * k, dbl + 2 * k, idesc, 1u);
                }
                release(&w_empty[a_stage]);
              }
              release(&w_empty[ws]);
            }
            it++;
          }
          if (wait_act) n_act++;
          if (leader) { if (PM) umma_pair_commit(&acc_full[h], kAllCtas); else umma_commit(&acc_full[h]); }  // PM: both CTAs' epilogues
          DBG_STAMP_MMA(blockIdx.x == 0 && tl == 3 && MODE == 1, s, h, 3);
        }
      }
      // every encoding stage of this tile has been waited for (w_full): its rows have left the scratch buffer tl & 1
      if (!DGRAD && p.enc_mode && leader) {
        mbar_arrive(&enc_free[tl & 1]);
        if (PM) mbar_arrive_cluster(map_to_cta(&enc_free[tl & 1], 1));  // the peer's tile went through the same w_full barriers
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarps + 2) {
    // ------------------------------------------------------------------ encoders (forward modes): one thread per sample row
    if (!DGRAD && p.enc_mode) {
      const int et = threadIdx.x - kThreadsS;  // 0 .. 32 * kEncWarpsS
      uint32_t tl = 0;
      for (int tile = blockIdx.x; (int)tl < tiles_per_cta; tile += gridDim.x, tl++) {
        if (tl >= 2) mbar_wait(&enc_free[tl & 1], ((tl >> 1) & 1) ^ 1);  // tile tl - 2 has been loaded out of this buffer
        const long out0 = p.enc_mode == 2 ? (long)(blockIdx.x * 2 + (int)(tl & 1)) * 128 : (long)tile * 128;
        for (int i = et; i < 128; i += 32 * kEncWarpsS) {
          const long m = (long)tile * 128 + i;
          const bool valid = m < p.M;
          if (valid || p.enc_mode == 2)
            enc::encode_row_to_planes<false>(p.rs, m, valid, out0 + i, p.enc_pos[0], p.enc_pos[1], p.enc_dir[0], p.enc_dir[1]);
        }
        asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy global writes -> visible to the TMA loads
        __syncwarp();
        if (lane == 0) mbar_arrive(&enc_ready[tl & 1]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: lane quarter warp % 4, column half warp / 4
    const int qtr = warp & 3, ch = warp >> 2;
    const uint32_t lane_off = (uint32_t)(qtr * 32) << 16;
    uint8_t* slot = stage_buf + warp * kSlot;    // TRAIN: two alternating [32 x 32] hi + lo box pairs; F16: two alternating fp16 boxes
    uint8_t* slot_row = slot + lane * 64;
    uint32_t n_ship = 0;                         // F16: boxes shipped by this warp (selects the half of the slot)
    const int swz = (lane >> 1) & 3;             // SWIZZLE_64B: 16-byte chunk index ^ address bits [7:8]
    uint32_t n_full[2] = {0, 0};
    // signals to the MMA warp: this CTA's, or (PM) the leader's through its cluster address — the leader issues for both tiles
    auto arrive_mma = [&](uint64_t* bar) {
      if (PM) mbar_arrive_cluster(map_to_cta(bar, 0));
      else mbar_arrive(bar);
    };
    for (int ti = 0, tile = blockIdx.x; ti < tiles_per_cta; ti++, tile += gridDim.x) {
      const int row_w = tile * 128 + qtr * 32;
      const int row_t = qtr * 32 + lane;         // row within the tile
      const long row = (long)row_w + lane;
      const bool row_ok = row < p.M;
      const float r1v = (DGRAD && row_ok) ? __ldg(p.r1 + row) : 0.f;
      const float s16 = (DGRAD && F16) ? __ldg(p.dz_scale) : 1.0f;
      uint32_t held_h[32], held_l[32];           // first half's words wait here until the second half's MMAs have read ACT
      for (int s = 0; s < p.n_steps; s++) {
        const SplitParams::Step st = p.steps[s];
        float head[3] = {0.f, 0.f, 0.f};
        // TRAIN: ship one 32-column chunk (both planes) of this warp's 32 rows
        auto ship = [&](int col, const uint32_t* hw, const uint32_t* lw, const uint32_t* fw) {
          if (!TRAIN) return;
          // the box (pair) this one overwrites has been read out.  F16: one fp16 box is half a slot, so two boxes alternate and the
          // warp only waits for the box before the previous one — the TMA engine's read-out latency leaves the warp's path
          constexpr bool dbl = F16 ? kDoubleBox : kDoubleBoxSplit;
          const uint32_t boff = dbl ? (n_ship & 1u) * (F16 ? 2048u : 4096u) : 0u;
          n_ship++;
          if (lane == 0) {
            if (dbl) tma_store_wait_read<1>();
            else tma_store_wait_read<0>();
          }
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; q++) {
            if (F16) {
              *reinterpret_cast<uint4*>(slot_row + boff + ((q ^ swz) << 4)) = make_uint4(fw[4 * q], fw[4 * q + 1], fw[4 * q + 2], fw[4 * q + 3]);
            } else {
              *reinterpret_cast<uint4*>(slot_row + boff + ((q ^ swz) << 4)) = make_uint4(hw[4 * q], hw[4 * q + 1], hw[4 * q + 2], hw[4 * q + 3]);
              *reinterpret_cast<uint4*>(slot_row + boff + 2048 + ((q ^ swz) << 4)) = make_uint4(lw[4 * q], lw[4 * q + 1], lw[4 * q + 2], lw[4 * q + 3]);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&p.map_act[s][0], slot + boff, col, row_w);
            if (!F16) tma_store_2d(&p.map_act[s][1], slot + boff + 2048, col, row_w);
            tma_store_commit();
          }
        };
        for (int h = 0; h < st.n_halves; h++) {
          const bool last_half = h == st.n_halves - 1;
          const int col_t = h * 128 + ch * 64;   // first column of this thread in this half
          const float* bias = s_const + st.bias_off + col_t;
          const float* head_w = s_const + (st.head == 3 ? p.head_rgb_off : p.head_d_off) + col_t;
          const uint32_t acc = ACC + 128 * h + lane_off + ch * 64;
          uint32_t m0, m1;
          uint2 mk = make_uint2(0u, 0u);  // DGRAD: ReLU mask words of this thread's 64 columns (requested before the wait)
          if (DGRAD && row_ok && col_t < st.n_cols) mk = __ldg(reinterpret_cast<const uint2*>(p.bits[s] + row * (st.n_cols >> 5) + (col_t >> 5)));
          const float* v1 = (DGRAD && s == 0) ? s_const + p.head_d_off + col_t : nullptr;
          uint32_t fw[F16 ? 16 : 1];  // F16: the chunk as fp16 pairs, shipped at once
          // REP = 1: hw_ = 16 fp16 pairs of 32 a (also what F16 ships), lw_ = 8 words E4M3(2^9 al), h8_ = 8 words E4M3(ah)
          auto chunk = [&](const uint32_t (&r)[32], int c, uint32_t* hw_, uint32_t* lw_, uint32_t* h8_) -> uint32_t {
            if (REP == 1) return f8c_chunk<MODE == 1>(r, bias + c * 32, st.head, head_w + c * 32, head, hw_, lw_, h8_);
            if (DGRAD) {
              dgrad_split_chunk<F16>(r, c == 0 ? mk.x : mk.y, r1v, v1 ? v1 + c * 32 : nullptr, hw_, lw_, fw, s16);
              return 0u;
            }
            return split_chunk<MODE == 1, F16>(r, bias + c * 32, st.head, head_w + c * 32, head, hw_, lw_, fw);
          };
          const bool dbg = blockIdx.x == 0 && ti == 3 && warp == 0 && MODE == 1;
          DBG_STAMP(dbg, s, h, 0);
          mbar_wait(&acc_full[h], n_full[h] & 1);
          n_full[h]++;
          tc_fence_after_sync();
          DBG_STAMP(dbg, s, h, 1);
          const bool unpark = last_half && st.n_halves == 2 && st.produces;
          if (unpark) {  // every MMA of the layer is complete: ACT is rewritten in place
            tmem_st_16(ACT_HI + lane_off + ch * 32, held_h); tmem_st_16(ACT_HI + lane_off + ch * 32 + 16, held_h + 16);
            tmem_st_16(ACT_LO + lane_off + ch * 32, held_l); tmem_st_16(ACT_LO + lane_off + ch * 32 + 16, held_l + 16);
          }
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(acc, r0);
          tmem_ld_32x32(acc + 32, r1);
          tmem_ld_wait();
          if (unpark) tmem_st_wait();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            arrive_mma(&acc_empty[h]);                  // accumulator half in registers: its next MMAs may start
            if (unpark) arrive_mma(&act_lo_ready);      // the next layer's k-blocks 0,1 may start
          }
          DBG_STAMP(dbg, s, h, 2);
          if (!last_half) {
            // REP = 1: held_l = [E4M3(al) chunk 0 | chunk 1 | E4M3(ah) chunk 0 | chunk 1] = the 32 ACT8 columns of this k-block
            m0 = chunk(r0, 0, held_h, held_l, held_l + 16);
            DBG_STAMP(dbg, s, h, 3);
            ship(col_t, held_h, held_l, REP == 1 ? held_h : fw);
            DBG_STAMP(dbg, s, h, 4);
            m1 = chunk(r1, 1, held_h + 16, held_l + (REP == 1 ? 8 : 16), held_l + 24);
            DBG_STAMP(dbg, s, h, 5);
            ship(col_t + 32, held_h + 16, held_l + 16, REP == 1 ? held_h + 16 : fw);
            DBG_STAMP(dbg, s, h, 6);
          } else if (kLateShip && TRAIN && !(F16 && REP == 0)) {
            // The next layer's k-blocks 2.. wait for act_ready: everything that is not needed for it — restaging the two chunks
            // in shared memory, waiting for the previous box to be read out, issuing the TMA stores — moves BEHIND the signal.
            // The words wait in held_h / held_l, which are free here (the parked half went into ACT above).
            const uint32_t out = lane_off + (uint32_t)(col_t >> 1);
            uint32_t* hw0 = held_h; uint32_t* lw0 = held_l; uint32_t* hw1 = held_h + 16; uint32_t* lw1 = held_l + 16;
            m0 = chunk(r0, 0, hw0, lw0, lw0 + 8);
            DBG_STAMP(dbg, s, h, 3);
            if (st.produces) {
              tmem_st_16(ACT_HI + out, hw0);
              if (REP == 1) { tmem_st_8(ACT_LO + out, lw0); tmem_st_8(ACT_LO + out + 16, lw0 + 8); }
              else tmem_st_16(ACT_LO + out, lw0);
            }
            m1 = chunk(r1, 1, hw1, lw1, lw1 + 8);
            if (st.produces) {
              tmem_st_16(ACT_HI + out + 16, hw1);
              if (REP == 1) { tmem_st_8(ACT_LO + out + 8, lw1); tmem_st_8(ACT_LO + out + 24, lw1 + 8); }
              else tmem_st_16(ACT_LO + out + 16, lw1);
              tmem_st_wait();
              tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) {
                if (st.n_halves == 1) arrive_mma(&act_lo_ready);
                arrive_mma(&act_ready);
              }
            }
            DBG_STAMP(dbg, s, h, 4);
            ship(col_t, hw0, lw0, hw0);       // F16 here means REP == 1: the fp16 words are hw
            DBG_STAMP(dbg, s, h, 5);
            ship(col_t + 32, hw1, lw1, hw1);
            DBG_STAMP(dbg, s, h, 6);
          } else {
            uint32_t hw[16], lw[16];
            const uint32_t out = lane_off + (uint32_t)(col_t >> 1);
            m0 = chunk(r0, 0, hw, lw, lw + 8);
            if (st.produces) {
              tmem_st_16(ACT_HI + out, hw);
              if (REP == 1) { tmem_st_8(ACT_LO + out, lw); tmem_st_8(ACT_LO + out + 16, lw + 8); }
              else tmem_st_16(ACT_LO + out, lw);
            }
            ship(col_t, hw, lw, REP == 1 ? hw : fw);
            m1 = chunk(r1, 1, hw, lw, lw + 8);
            if (st.produces) {
              tmem_st_16(ACT_HI + out + 16, hw);
              if (REP == 1) { tmem_st_8(ACT_LO + out + 8, lw); tmem_st_8(ACT_LO + out + 24, lw + 8); }
              else tmem_st_16(ACT_LO + out + 16, lw);
              tmem_st_wait();
              tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) {
                if (st.n_halves == 1) arrive_mma(&act_lo_ready);  // a one-half layer has no parked first instalment
                arrive_mma(&act_ready);
              }
            }
            ship(col_t + 32, hw, lw, REP == 1 ? hw : fw);
          }
          if (MODE == 1 && row_ok && col_t < st.n_cols) *reinterpret_cast<uint2*>(p.bits[s] + row * (st.n_cols >> 5) + (col_t >> 5)) = make_uint2(m0, m1);
        }
        if (!DGRAD && st.head) {  // the column halves of a row meet in shared memory
          if (ch == 1) { head_part[row_t][0] = head[0]; head_part[row_t][1] = head[1]; head_part[row_t][2] = head[2]; }
          named_barrier_sync(1 + qtr, 64);
          if (ch == 0 && row_ok) {
            constexpr float hs = REP == 1 ? 0.03125f : 1.0f;  // REP = 1: the head sums ran on 32 a
            if (st.head == 1) {
              p.raw_density[row] = (head[0] + head_part[row_t][0]) * hs + s_const[p.head_d_off + 256];
            } else {
#pragma unroll
              for (int n = 0; n < 3; n++) p.raw_rgb[row * 3 + n] = (head[n] + head_part[row_t][n]) * hs + s_const[p.head_rgb_off + 3 * 128 + n];
            }
          }
          named_barrier_sync(1 + qtr, 64);  // head_part may be rewritten
        }
      }
    }
    if (TRAIN && lane == 0) tma_store_wait_read<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA leaves while its peer may still signal its barriers
  if (warp == kEpiWarps) { if (PM) tmem_dealloc_pair<512>(tmem_base); else tmem_dealloc<512>(tmem_base); }
}


template <int MODE, bool F16, int REP = 0>
int launch_split(const SplitParams& p, int grid, int threads, size_t smem, bool pair, cudaStream_t st) {
  const void* kern = pair ? (p.pair_mma ? (const void*)k_mlp_fused_split<MODE, 2, F16, REP, true> : (const void*)k_mlp_fused_split<MODE, 2, F16, REP, false>)
                          : (const void*)k_mlp_fused_split<MODE, 1, F16, REP, false>;
  SplitParams pp = p;
  return launch_persistent_clusters(kern, grid, threads, smem, 218 * 1024, pair ? 2 : 1, &pp, st);
}

}  // namespace

// Host side.  wplanes_hi/lo[s]: weight planes of dense layer s (trunk 0..D-1, then the condition layer), [N, kpad[s]].
// consts layout as in mlp_fused.cu.  act_hi == nullptr -> inference (nothing but the raw heads is written).
int launch_mlp_fused_forward_split(const __nv_bfloat16* pos_hi, const __nv_bfloat16* pos_lo, int pos_pitch, const __nv_bfloat16* dir_hi,
                                   const __nv_bfloat16* dir_lo, int dir_pitch, const __nv_bfloat16* const* w_hi,
                                   const __nv_bfloat16* const* w_lo, const int* kpad, const int* in_b, int D, int W, int Wc, long M,
                                   const float* consts_dev, int n_consts, int head_d_off, int head_rgb_off, const int* bias_off,
                                   float* raw_density, float* raw_rgb, __nv_bfloat16* const* act_hi, __nv_bfloat16* const* act_lo,
                                   uint32_t* const* bits_out, const RaySource* rays, long enc_scratch_rows, bool pair, cudaStream_t st,
                                   bool act_f16, bool fp8c, bool pair_mma) {
  // fp8c: w_hi[s] / w_lo[s] are the W16 / W8 planes of f8c_chunk's representation (launch_f32_to_f8c_planes); training needs act_f16
  if (fp8c && act_hi && !act_f16) { set_error("fused forward: fp8 corrections in training need the fp16 activation planes"); return 100001; }
  if (act_f16 && (!act_hi || rays)) { set_error("fused forward: fp16 activation planes are a training option with the encode kernel"); return 100001; }
  if (!((W == 256 && Wc == 128) || (W == 128 && Wc == 64)) || D + 1 > kMaxStepsS || pos_pitch != 128 || dir_pitch != 64) {
    set_error("fused forward supports widths 256/128 and 128/64 (trunk / condition), position pitch 128, direction pitch 64");
    return 100001;
  }
  const bool train = act_hi != nullptr;
  const size_t smem = (size_t)(train ? train_stages(act_f16) : kNSInfer) * kStageB + (train ? kEpiWarps * train_slot_bytes(act_f16) : 0) +
                      (size_t)((n_consts + 3) / 4 * 4) * sizeof(float) + 1024;
  const int sms = device_sm_count();
  if (smem > 218 * 1024) { set_error("fused forward: %zu bytes of shared memory needed", smem); return 100001; }
  SplitParams p;
  memset(&p, 0, sizeof(p));
  const int tiles = (int)cdiv(M, 128);
  const int grid = pair ? ((tiles < sms ? tiles : sms) + 1) / 2 * 2 : (tiles < sms ? tiles : sms);  // whole clusters
  // rays != nullptr: the kernel's encoder warps build the encodings (see launch_mlp_fused_forward)
  long map_rows = M;
  if (rays) {
    p.enc_mode = enc_scratch_rows > 0 ? 2 : 1;
    p.rs = *rays;
    p.enc_pos[0] = const_cast<__nv_bfloat16*>(pos_hi); p.enc_pos[1] = const_cast<__nv_bfloat16*>(pos_lo);
    p.enc_dir[0] = const_cast<__nv_bfloat16*>(dir_hi); p.enc_dir[1] = const_cast<__nv_bfloat16*>(dir_lo);
    if (p.enc_mode == 2) {
      if (enc_scratch_rows < (long)grid * 256) { set_error("fused forward: encoding scratch too small"); return 100001; }
      map_rows = enc_scratch_rows;
    }
    if (rays->deg_point % 4 || rays->deg_point * 6 > 120 || rays->deg_view > 4 || (long)rays->R * rays->S < M) {
      set_error("fused forward: in-kernel encoding needs deg_point %% 4 == 0, <= 20, deg_view <= 4"); return 100001;
    }
  }
  NERF_TRY(tc_make_tmap(&p.map_pos[0], pos_hi, map_rows, 128, pos_pitch, 128));
  NERF_TRY(tc_make_tmap(&p.map_pos[1], pos_lo, map_rows, 128, pos_pitch, 128));
  NERF_TRY(tc_make_tmap(&p.map_dir[0], dir_hi, map_rows, 64, dir_pitch, 128));
  NERF_TRY(tc_make_tmap(&p.map_dir[1], dir_lo, map_rows, 64, dir_pitch, 128));
  for (int s = 0; s <= D; s++) {
    const int N = s < D ? W : Wc;
    NERF_TRY(tc_make_tmap(&p.map_w[s][0], w_hi[s], N, kpad[s], kpad[s], 128));
    NERF_TRY(tc_make_tmap(&p.map_w[s][1], w_lo[s], N, kpad[s], kpad[s], 128));
    NERF_TRY(tc_make_tmap(&p.map_w64[s][0], w_hi[s], N, kpad[s], kpad[s], 64));
    NERF_TRY(tc_make_tmap(&p.map_w64[s][1], w_lo[s], N, kpad[s], kpad[s], 64));
    if (train) {
      NERF_TRY(tc_make_tmap_box(&p.map_act[s][0], act_hi[s], M, N, N, 32, 32));  // act_f16: the layer's ONE fp16 plane
      if (!act_f16) NERF_TRY(tc_make_tmap_box(&p.map_act[s][1], act_lo[s], M, N, N, 32, 32));
      p.bits[s] = bits_out[s];
    }
    SplitParams::Step& stp = p.steps[s];
    if (s == 0) { stp.n_act_kb = 0; stp.enc_kind = 1; stp.n_enc_kb = 2; }
    else if (s < D) { stp.n_act_kb = (int16_t)(W / 64); stp.enc_kind = in_b[s] ? 1 : 0; stp.n_enc_kb = in_b[s] ? 2 : 0; }
    else { stp.n_act_kb = (int16_t)(W / 64); stp.enc_kind = 2; stp.n_enc_kb = 1; }
    stp.n_halves = (int16_t)((N + 127) / 128);
    stp.n_cols = (int16_t)N;
    stp.produces = s < D ? 1 : 0;
    stp.head = s == D - 1 ? 1 : (s == D ? 3 : 0);
    stp.bias_off = bias_off[s];
  }
  p.n_steps = D + 1; p.M = M; p.consts = consts_dev; p.n_consts = n_consts;
  p.head_d_off = head_d_off; p.head_rgb_off = head_rgb_off;
  p.raw_density = raw_density; p.raw_rgb = raw_rgb;
  p.pair_mma = (pair && pair_mma) ? 1 : 0;
  if (fp8c) return train ? launch_split<1, true, 1>(p, grid, kThreadsSE, smem, pair, st) : launch_split<0, false, 1>(p, grid, kThreadsSE, smem, pair, st);
  if (act_f16) return launch_split<1, true>(p, grid, kThreadsSE, smem, pair, st);
  return train ? launch_split<1, false>(p, grid, kThreadsSE, smem, pair, st) : launch_split<0, false>(p, grid, kThreadsSE, smem, pair, st);
}


// The backward dgrad chain of the trunk in the fp32-accurate mode (see launch_mlp_fused_dgrad in mlp_fused.cu for the step
// order): dz_cond hi/lo [M, Wc]; wt_hi/lo[0] = W_cond^T [W, Wc], wt[i] = W_{D-i}^T [W, W]; dz_out[i] = dZ of trunk layer D-1-i.
int launch_mlp_fused_dgrad_split(const __nv_bfloat16* dz_cond_hi, const __nv_bfloat16* dz_cond_lo, int dz_cond_pitch,
                                 const __nv_bfloat16* const* wt_hi, const __nv_bfloat16* const* wt_lo, const int* wt_pitch, int D, int W,
                                 int Wc, long M, const float* consts_dev, int n_consts, int head_d_off, const float* d_raw_density,
                                 __nv_bfloat16* const* dz_out_hi, __nv_bfloat16* const* dz_out_lo, const uint32_t* const* mask_bits,
                                 bool pair, cudaStream_t st, const float* dz_scale_f16, bool pair_mma) {
  const bool f16 = dz_scale_f16 != nullptr;  // dz_out_hi[s] is then ONE fp16 plane holding dZ * *dz_scale_f16 (dz_out_lo unused)
  if (!((W == 256 && Wc == 128) || (W == 128 && Wc == 64)) || D > kMaxStepsS || D < 2) { set_error("fused dgrad supports widths 256/128 and 128/64"); return 100001; }
  const int sms = device_sm_count();
  const size_t smem = (size_t)train_stages(dz_scale_f16 != nullptr) * kStageB + kEpiWarps * train_slot_bytes(dz_scale_f16 != nullptr) +
                      (size_t)((n_consts + 3) / 4 * 4) * sizeof(float) + 1024;
  if (smem > 218 * 1024) { set_error("fused dgrad: %zu bytes of shared memory needed", smem); return 100001; }
  SplitParams p;
  memset(&p, 0, sizeof(p));
  NERF_TRY(tc_make_tmap(&p.map_pos[0], dz_cond_hi, M, Wc, dz_cond_pitch, 128));  // A of step 0, streamed like an encoding
  NERF_TRY(tc_make_tmap(&p.map_pos[1], dz_cond_lo, M, Wc, dz_cond_pitch, 128));
  for (int s = 0; s < D; s++) {
    NERF_TRY(tc_make_tmap(&p.map_w[s][0], wt_hi[s], W, s == 0 ? Wc : W, wt_pitch[s], 128));
    NERF_TRY(tc_make_tmap(&p.map_w[s][1], wt_lo[s], W, s == 0 ? Wc : W, wt_pitch[s], 128));
    NERF_TRY(tc_make_tmap(&p.map_w64[s][0], wt_hi[s], W, s == 0 ? Wc : W, wt_pitch[s], 64));
    NERF_TRY(tc_make_tmap(&p.map_w64[s][1], wt_lo[s], W, s == 0 ? Wc : W, wt_pitch[s], 64));
    NERF_TRY(tc_make_tmap_box(&p.map_act[s][0], dz_out_hi[s], M, W, W, 32, 32));
    if (!f16) NERF_TRY(tc_make_tmap_box(&p.map_act[s][1], dz_out_lo[s], M, W, W, 32, 32));
    p.bits[s] = const_cast<uint32_t*>(mask_bits[s]);
    SplitParams::Step& stp = p.steps[s];
    if (s == 0) { stp.n_act_kb = 0; stp.enc_kind = 1; stp.n_enc_kb = (int16_t)(Wc / 64); }
    else { stp.n_act_kb = (int16_t)(W / 64); stp.enc_kind = 0; stp.n_enc_kb = 0; }
    stp.n_halves = (int16_t)(W / 128);
    stp.n_cols = (int16_t)W;
    stp.produces = s < D - 1 ? 1 : 0;
    stp.head = 0; stp.bias_off = 0;
  }
  p.n_steps = D; p.M = M; p.consts = consts_dev; p.n_consts = n_consts;
  p.head_d_off = head_d_off; p.head_rgb_off = 0; p.r1 = d_raw_density; p.dz_scale = dz_scale_f16; p.pair_mma = (pair && pair_mma) ? 1 : 0;
  const int tiles = (int)cdiv(M, 128);
  if (f16) return launch_split<2, true>(p, tiles < sms ? tiles : sms, kThreadsS, smem, pair, st);
  return launch_split<2, false>(p, tiles < sms ? tiles : sms, kThreadsS, smem, pair, st);
}

}  // namespace nerf

#ifdef NERF_SPLIT_DBG
extern "C" int nerf_debug_split_stamps(unsigned long long* epi, unsigned long long* mma) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(epi, nerf::g_split_dbg, sizeof(nerf::g_split_dbg)) != cudaSuccess) return 1;
  if (cudaMemcpyFromSymbol(mma, nerf::g_split_dbg_mma, sizeof(nerf::g_split_dbg_mma)) != cudaSuccess) return 1;
  return 0;
}
#endif


// mlp_fused.cu — the whole MipNeRF MLP forward (trunk, skip, density head, condition layer, rgb head) as
// ONE persistent tcgen05 kernel with the activations resident in TENSOR MEMORY across layers (bf16 mode, inference).
//
// Replaces the per-layer launch chain of AcceleratedMLP::get_output (ANU/AcceleratedMLP.cpp:214-255), which writes every
// layer's outputs and pre-activations to global memory (2 x 134 MB per layer per level at the reference batch).
//
// Per 128-row tile (one CTA per SM walks tiles), TMEM (512 columns) holds
//   ACC[0], ACC[1]   2 x 128 fp32 columns: the two N-halves of the current layer's accumulator
//   ACT[0], ACT[1]   2 x 128 columns = 256 bf16 per row: the A operand of the current layer / the next layer's A
// and a layer is two half-layers: tcgen05.mma reads A straight from TMEM (".ts" form: D[tmem] += A[tmem] * B[smem]) while the
// four epilogue warps drain the other half (tcgen05.ld -> bias + ReLU -> bf16 pairs -> tcgen05.st into ACT[next]).  The
// next layer's MMAs start k-block by k-block as soon as the half of ACT they read has been written, so the tensor pipe
// and the epilogue overlap inside a tile and nothing but the raw heads ever leaves the SM.  Weights stream from L2
// through a 6-stage TMA ring of [128 x 64] bf16 tiles; the IPE / direction encodings of the tile (the only global
// inputs) are TMA-loaded once into shared memory and used as shared-memory A operands for layer 0, the skip layer and
// the condition layer.  The N=1 density head and N=3 rgb head are FMAs on the epilogue registers.
#include <cuda.h>


#include "encode_rows.cuh"
#include "gemm_tc.cuh"
#include "sm100.cuh"

namespace nerf {
namespace {

using namespace sm100;

constexpr int kThreadsF = 320;                    // 8 epilogue warps + TMA producer + MMA issuer
constexpr int kEncWarps = 2;                      // forward kernels: + 2 encoder warps (cast_rays + IPE + direction PE in-kernel)
constexpr int kThreadsFE = kThreadsF + 32 * kEncWarps;
constexpr int kWStages = 6;                       // render; training gives one stage to the activation staging
constexpr int kWStageBytes = 128 * 128;           // [128 rows (N half) x 64 bf16]
constexpr int kEncBytes = 2 * 16384 + 16384;      // per tile: position encoding 2 boxes of [128 x 64], direction 1 box
constexpr int kMaxSteps = 12;
constexpr int kStageSlot = 4096;                  // per epilogue warp: one [32 rows x 64 bf16] store box
#ifndef NERF_LATE_SHIP
#define NERF_LATE_SHIP 1
#endif
constexpr bool kLateShipF = NERF_LATE_SHIP != 0;  // second-half epilogue: TMA-store the chunks after act_ready instead of between them

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace

struct alignas(64) FusedParams {
  CUtensorMap map_pos, map_dir;     // encodings [M, 128] / [M, 64] bf16, box {64, 128}
  CUtensorMap map_w[kMaxSteps];     // weight planes [N, Kpad] bf16, box {64, 128}
  CUtensorMap map_w64[kMaxSteps];   // the same planes, box {64, 64}: one CTA of a 2-CTA cluster fetches half a stage for both
  struct Step {
    int16_t n_act_kb;   // k-blocks read from the TMEM-resident activation (0 or width/64)
    int16_t enc_kind;   // 0 none, 1 position encoding, 2 direction encoding (shared-memory A operand)
    int16_t n_enc_kb;   // k-blocks of that encoding
    int16_t n_halves;   // ceil(N / 128): a 64-wide layer runs as one half whose upper 64 weight rows are the TMA's out-of-bounds zeros
    int16_t n_cols;     // N: output columns that exist (64, 128 or 256)
    int16_t produces;   // 1: writes the next layer's activation; 0: last (condition) layer
    int16_t head;       // 0 none, 1 density head after this step, 3 rgb head after this step
    int32_t bias_off;   // offset into the staged constants
  } steps[kMaxSteps];
  int n_steps;
  long M;
  const float* consts;  // device: biases of every step, then density head w[256], b, rgb head w[3][128], b[3]
  int n_consts;
  int head_d_off, head_rgb_off;
  float* raw_density;   // [M]
  float* raw_rgb;       // [M, 3]
  // training only: every step's activations [M, N] bf16 (box {64, 32}, one store per epilogue warp) and ReLU bit planes
  CUtensorMap map_act[kMaxSteps];
  uint32_t* bits[kMaxSteps];
  // dgrad chain only: bits[s] is READ (ReLU mask of the step's output); step 0 adds the rank-1 term r1[row] * v1[col]
  const float* r1;      // [M] dL/d raw_density
  // forward only — in-kernel cast_rays + IPE + direction PE (accelerated_functions.cu:292-317, 187-221; encode_rows.cuh):
  // enc_mode 0: map_pos / map_dir view planes another kernel wrote (stand-alone AcceleratedMLP::get_output);
  //          1: the encoder warps write the level's planes [M, 128] / [M, 64] (training: the wgrad GEMMs read them later);
  //          2: they write a per-CTA, double-buffered scratch [gridDim.x * 2 * 256 rows] that stays in L2 (rendering)
  int enc_mode;
  RaySource rs;
  __nv_bfloat16 *enc_pos, *enc_dir;
};

namespace {

// One 32-column chunk of the accumulator: + bias, ReLU, bf16 pairs out (two per 32-bit TMEM column of the next A
// operand).  HEAD = 0: ReLU on the packed pair (one HMNMX2 per two values).  HEAD = 1 / 3: the density / rgb head
// rides along as FMAs on the fp32 activations.  BITS: also the ReLU mask of the 32 columns (bit j = column j passed),
// gathered from the sign bits with one funnel shift per value (four independent chains of eight).
template <int HEAD, bool BITS>
__device__ __forceinline__ uint32_t epi_chunk(const uint32_t (&r)[32], const float* bias, const float* head_w, float (&head)[3],
                                              uint32_t* packed16) {
  const float4* bv = reinterpret_cast<const float4*>(bias);
  float x[32];
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const float4 bb = bv[q];
    x[4 * q] = __uint_as_float(r[4 * q]) + bb.x; x[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bb.y;
    x[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bb.z; x[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bb.w;
  }
  uint32_t mask = 0u;
  if (BITS) {
    uint32_t m[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int e = 7; e >= 0; e--) m[c] = __funnelshift_l(__float_as_uint(x[8 * c + e]), m[c], 1);  // m = m << 1 | sign
    mask = ~__byte_perm(__byte_perm(m[0], m[1], 0x0040), __byte_perm(m[2], m[3], 0x0040), 0x5410);
  }
#pragma unroll
  for (int q = 0; q < 8; q++) {
    float x0 = x[4 * q], x1 = x[4 * q + 1], x2 = x[4 * q + 2], x3 = x[4 * q + 3];
    if (HEAD == 0) {
      const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
      __nv_bfloat162 a = __hmax2(__floats2bfloat162_rn(x0, x1), z), b = __hmax2(__floats2bfloat162_rn(x2, x3), z);
      packed16[2 * q] = *reinterpret_cast<uint32_t*>(&a);
      packed16[2 * q + 1] = *reinterpret_cast<uint32_t*>(&b);
    } else {
      x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f);
#pragma unroll
      for (int n = 0; n < HEAD; n++) {
        const float4 w = reinterpret_cast<const float4*>(head_w + n * 128)[q];
        head[n] = fmaf(x0, w.x, head[n]); head[n] = fmaf(x1, w.y, head[n]);
        head[n] = fmaf(x2, w.z, head[n]); head[n] = fmaf(x3, w.w, head[n]);
      }
      packed16[2 * q] = pack2(x0, x1);
      packed16[2 * q + 1] = pack2(x2, x3);
    }
  }
  return mask;
}
template <bool BITS>
__device__ __forceinline__ uint32_t epi_chunk_any(int head_kind, const uint32_t (&r)[32], const float* bias, const float* head_w,
                                                  float (&head)[3], uint32_t* packed16) {
  if (head_kind == 0) return epi_chunk<0, BITS>(r, bias, head_w, head, packed16);
  if (head_kind == 1) return epi_chunk<1, BITS>(r, bias, head_w, head, packed16);
  return epi_chunk<3, BITS>(r, bias, head_w, head, packed16);
}
// 16 packed words = 32 bf16 columns = half of this thread's 128-byte row of the warp's store box (128-byte swizzle)
__device__ __forceinline__ void stage_half_row(uint8_t* row_ptr, int half, int swz, const uint32_t* pk) {
#pragma unroll
  for (int q = 0; q < 4; q++)
    *reinterpret_cast<uint4*>(row_ptr + (((half * 4 + q) ^ swz) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}

// dgrad chain: one 32-column chunk of dX = dZ W (+ r1 v1^T), masked by the ReLU bits of the layer below, bf16 pairs out
__device__ __forceinline__ void dgrad_chunk(const uint32_t (&r)[32], uint32_t mask, float r1, const float* v1, uint32_t* packed16) {
#pragma unroll
  for (int q = 0; q < 16; q++) {
    float x0 = __uint_as_float(r[2 * q]), x1 = __uint_as_float(r[2 * q + 1]);
    if (v1) { x0 = fmaf(r1, v1[2 * q], x0); x1 = fmaf(r1, v1[2 * q + 1], x1); }
    x0 = ((mask >> (2 * q)) & 1u) ? x0 : 0.f;
    x1 = ((mask >> (2 * q + 1)) & 1u) ? x1 : 0.f;
    packed16[q] = pack2(x0, x1);
  }
}

// One CTA walks PAIRS of 128-row tiles.  Tensor memory (512 columns): per tile t, ACT[t] = 128 columns holding the
// 256 bf16 activations of the current layer (updated IN PLACE once the layer's MMAs are complete) and ACC[t] = 128
// fp32 accumulator columns (one N-half of the layer).  Both tiles consume every weight stage, so a [128 x 64] weight
// tile is fetched from L2 once per 256 rows.  Warps 0-3 are the epilogue of tile 0, warps 4-7 of tile 1 (warp % 4 =
// the TMEM lane quarter it may access), warp 8 the TMA producer, warp 9 the MMA issuer.
// TRAIN: every layer's activations and ReLU bit planes are also written out for the backward pass (each epilogue warp
// restages its 32 rows in shared memory and issues its own TMA store).
// MODE 0: inference forward; 1: training forward; 2: backward dgrad chain (A of step 0 = dZ of the condition layer from
// shared memory, epilogue = rank-1 density-head term + ReLU mask of the layer below, every step's dZ written out for wgrad).
// Forward modes run two more warps (10, 11): the ENCODERS.  For pair j they turn the pair's 256 samples (t-values + ray)
// into the position / direction encoding rows — one thread per row, the arithmetic of encode.cu — and store them through
// L2 (planes or scratch, see FusedParams::enc_mode); `fence.proxy.async` + enc_ready[j & 1] hands them to the producer,
// whose TMA loads bring them into shared memory as the A operands of layer 0, the skip layer and the condition layer.
// They run up to two pairs ahead of the MMAs (enc_free[j & 1] = pair j's loads have landed), so the encode of the next
// pair is hidden under the current pair's MMAs and no encode kernel runs at all.
// CL = 2: thread-block clusters of two CTAs share every weight stage (see mlp_fused_split.cu): rank r fetches rows [64 r, 64 r + 64)
// of the stage's [128 x 64] weight tile with `.multicast::cluster` into both rings, a slot is recycled when both CTAs' MMAs have
// consumed it, every CTA walks the same number of tile pairs (phantom pairs load zeros and store nothing).
template <int MODE, int CL>
__global__ void __launch_bounds__(MODE == 2 ? kThreadsF : kThreadsFE, 1) k_mlp_fused_fwd(const __grid_constant__ FusedParams p) {
  static_assert(CL == 1 || CL == 2, "cluster of 1 or 2 CTAs");
  constexpr bool TRAIN = MODE != 0;   // activations / gradients are written out
  constexpr bool DGRAD = MODE == 2;
  constexpr int NS = TRAIN ? kWStages - 1 : kWStages;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_full[kWStages], w_empty[kWStages], pos_full, pos_empty, dir_full, dir_empty, acc_full, acc_empty, act_ready, act_lo_ready;
  __shared__ uint64_t enc_ready[2], enc_free[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  uint8_t* w_ring = smem;                                   // NS x 16 KB
  uint8_t* pos_buf = smem + NS * kWStageBytes;              // tile t: 2 boxes of [128 x 64] at t * 32 KB
  uint8_t* dir_buf = pos_buf + 2 * 32768;                   // tile t: 1 box at t * 16 KB
  uint8_t* stage_buf = dir_buf + 2 * 16384;                 // TRAIN: 8 epilogue warps x 4 KB
  float* s_const = reinterpret_cast<float*>(stage_buf + (TRAIN ? 8 * kStageSlot : 0));

  const int n_tiles = (int)((p.M + 127) / 128);
  const int n_pairs = (n_tiles + 1) / 2;
  const int pairs_per_cta = (n_pairs + (int)gridDim.x - 1) / (int)gridDim.x;  // the same for every CTA: rings stay in lock-step
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kAllCtas = (uint16_t)((1u << CL) - 1u);
  int last_pos_step = 0;
  for (int s = 0; s < p.n_steps; s++) if (p.steps[s].enc_kind == 1) last_pos_step = s;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; s++) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], CL); }
    mbar_init(&pos_full, 1); mbar_init(&pos_empty, 1); mbar_init(&dir_full, 1); mbar_init(&dir_empty, 1);
    mbar_init(&acc_full, 1); mbar_init(&acc_empty, 8); mbar_init(&act_ready, 8); mbar_init(&act_lo_ready, 8);
    for (int b = 0; b < 2; b++) { mbar_init(&enc_ready[b], kEncWarps); mbar_init(&enc_free[b], 1); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.n_consts; i += blockDim.x) s_const[i] = __ldg(p.consts + i);
  if (warp == 8) {
    tmem_alloc<512>(&tmem_base_smem);
    if (lane == 0) {
      prefetch_tmap(&p.map_pos);
      if (!DGRAD) prefetch_tmap(&p.map_dir);
      for (int s = 0; s < p.n_steps; s++) { prefetch_tmap(&p.map_w[s]); if (TRAIN) prefetch_tmap(&p.map_act[s]); }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast at them
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;  // tile t: ACT at +256 t, ACC at +256 t + 128

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer: encodings per pair + weight ring
    if (lane == 0) {
      uint32_t wit = 0, pl = 0;
      for (int pair = blockIdx.x; (int)pl < pairs_per_cta; pair += gridDim.x, pl++) {
        // rows of this pair's encodings in the tensor the maps view: the sample index, or this CTA's scratch buffer pl & 1
        const int row0 = (!DGRAD && p.enc_mode == 2) ? (blockIdx.x * 2 + (int)(pl & 1)) * 256 : pair * 256;
        if (!DGRAD && p.enc_mode) mbar_wait(&enc_ready[pl & 1], (pl >> 1) & 1);  // the encoder warps have written them
        mbar_wait(&pos_empty, (pl & 1) ^ 1);
        mbar_arrive_expect_tx(&pos_full, 2 * 32768);
        for (int t = 0; t < 2; t++) {
          tma_load_2d(pos_buf + t * 32768, &p.map_pos, 0, row0 + t * 128, &pos_full);
          tma_load_2d(pos_buf + t * 32768 + 16384, &p.map_pos, 64, row0 + t * 128, &pos_full);
        }
        for (int s = 0; s < p.n_steps; s++) {
          const FusedParams::Step st = p.steps[s];
          const int n_kb = st.n_act_kb + st.n_enc_kb;
          if (st.enc_kind == 2) {  // direction encodings: needed by the condition layer only, so loaded just ahead of it
            mbar_wait(&dir_empty, (pl & 1) ^ 1);
            mbar_arrive_expect_tx(&dir_full, 2 * 16384);
            for (int t = 0; t < 2; t++) tma_load_2d(dir_buf + t * 16384, &p.map_dir, 0, row0 + t * 128, &dir_full);
          }
          for (int h = 0; h < st.n_halves; h++)
            for (int kb = 0; kb < n_kb; kb++, wit++) {
              const int ws = wit % NS;
              mbar_wait(&w_empty[ws], ((wit / NS) & 1) ^ 1);
              mbar_arrive_expect_tx(&w_full[ws], kWStageBytes);
              if (CL == 1) tma_load_2d(w_ring + (size_t)ws * kWStageBytes, &p.map_w[s], kb * 64, h * 128, &w_full[ws]);
              else tma_load_2d_multicast(w_ring + (size_t)ws * kWStageBytes + cta_rank * 8192, &p.map_w64[s], kb * 64, h * 128 + (int)cta_rank * 64,
                                         &w_full[ws], kAllCtas);  // this CTA's 64 rows of the tile, for the whole cluster
            }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer.  The whole warp walks the schedule in
    // uniform control flow (so addresses and descriptors live in uniform registers); one elected lane issues.
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, 128, false, false);
    const uint64_t desc0 = make_smem_desc(0, 16, 1024);  // + (smem address >> 4)
    const uint32_t pos_base = smem_u32(pos_buf), dir_base = smem_u32(dir_buf), ring_base = smem_u32(w_ring);
    uint32_t wit = 0, pl = 0, n_acc = 0, n_act = 0;
    auto release = [&](uint64_t* bar) {  // in EVERY CTA of the cluster: the peer multicasts into this CTA's slot too
      if (CL == 1) umma_commit(bar);
      else umma_commit_multicast(bar, kAllCtas);
    };
    for (; (int)pl < pairs_per_cta; pl++) {
      for (int s = 0; s < p.n_steps; s++) {
        const FusedParams::Step st = p.steps[s];
        for (int h = 0; h < st.n_halves; h++) {
          mbar_wait(&acc_empty, (n_acc & 1) ^ 1);  // both tiles' accumulators drained by the epilogue warps
          n_acc++;
          // ACT of both tiles is rewritten in place in two instalments: the layer below's first N-half (this layer's k-blocks
          // 0,1; parked in registers during its second half's MMAs) as soon as those MMAs are complete, its second N-half
          // when that half's epilogue is done — this layer's first k-blocks run under that epilogue instead of after it
          const bool wait_act = h == 0 && st.n_act_kb > 0;
          if (h == 0) {
            if (wait_act) mbar_wait(&act_lo_ready, n_act & 1);
            if (s == 0) mbar_wait(&pos_full, pl & 1);
            if (st.enc_kind == 2) mbar_wait(&dir_full, pl & 1);
          }
          tc_fence_after_sync();
          for (int kb = 0; kb < st.n_act_kb; kb++, wit++) {  // A = activations resident in tensor memory
            if (wait_act && kb == st.n_act_kb / 2) { mbar_wait(&act_ready, n_act & 1); tc_fence_after_sync(); }
            const uint32_t ws = wit % NS;
            mbar_wait(&w_full[ws], (wit / NS) & 1);
            tc_fence_after_sync();
            const uint64_t db = desc0 + ((ring_base + ws * kWStageBytes) >> 4);
            if (leader) {
#pragma unroll
              for (int t = 0; t < 2; t++)
#pragma unroll
                for (int k = 0; k < 4; k++)
                  umma_bf16_ts(tmem_base + 256 * t + 128, tmem_base + 256 * t + kb * 32 + k * 8, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
              release(&w_empty[ws]);
            }
          }
          if (wait_act) n_act++;
          for (int kb = 0; kb < st.n_enc_kb; kb++, wit++) {  // A = this pair's encodings in shared memory
            const uint32_t ws = wit % NS;
            mbar_wait(&w_full[ws], (wit / NS) & 1);
            tc_fence_after_sync();
            const uint64_t db = desc0 + ((ring_base + ws * kWStageBytes) >> 4);
            const uint32_t a0 = st.enc_kind == 2 ? dir_base : pos_base + kb * 16384;
            const uint32_t a_tile = st.enc_kind == 2 ? 16384 : 32768;
            if (leader) {
#pragma unroll
              for (int t = 0; t < 2; t++)
#pragma unroll
                for (int k = 0; k < 4; k++)
                  umma_bf16(tmem_base + 256 * t + 128, desc0 + ((a0 + t * a_tile) >> 4) + 2 * k, db + 2 * k, idesc,
                            (st.n_act_kb | kb | k) ? 1u : 0u);
              release(&w_empty[ws]);
            }
          }
          if (leader) umma_commit(&acc_full);
        }
        if (s == last_pos_step && leader) umma_commit(&pos_empty);  // the next pair's position encodings may land
      }
      if (leader) {
        umma_commit(&dir_empty);
        // pos_full and dir_full of this pair have been waited for: its encodings have left the scratch buffer pl & 1
        if (!DGRAD && p.enc_mode) mbar_arrive(&enc_free[pl & 1]);
      }
    }
    __syncwarp();
  } else if (warp >= 10) {
    // ------------------------------------------------------------------ encoders (forward modes): one thread per sample row
    if (!DGRAD && p.enc_mode) {
      const int et = threadIdx.x - 320;  // 0 .. 32 * kEncWarps
      uint32_t pl = 0;
      for (int pair = blockIdx.x; (int)pl < pairs_per_cta; pair += gridDim.x, pl++) {
        if (pl >= 2) mbar_wait(&enc_free[pl & 1], ((pl >> 1) & 1) ^ 1);  // pair pl - 2 has been loaded out of this buffer
        const long out0 = p.enc_mode == 2 ? (long)(blockIdx.x * 2 + (int)(pl & 1)) * 256 : (long)pair * 256;
        for (int i = et; i < 256; i += 32 * kEncWarps) {
          const long m = (long)pair * 256 + i;
          const bool valid = m < p.M;
          if (valid || p.enc_mode == 2) enc::encode_row_to_planes<true>(p.rs, m, valid, out0 + i, p.enc_pos, nullptr, p.enc_dir, nullptr);
        }
        asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy global writes -> visible to the TMA loads
        __syncwarp();
        if (lane == 0) mbar_arrive(&enc_ready[pl & 1]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: tile t = warp / 4, rows 32 (warp % 4) ..
    const int t = warp >> 2;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t act = tmem_base + 256 * t + lane_off, acc = act + 128;
    uint8_t* slot = stage_buf + warp * kStageSlot;  // TRAIN: this warp's [32 x 64] bf16 store box
    uint8_t* slot_row = slot + lane * 128;
    const int swz = lane & 7;
    uint32_t n_full = 0;
    for (int pi = 0, pair = blockIdx.x; pi < pairs_per_cta; pi++, pair += gridDim.x) {
      const int row_w = pair * 256 + t * 128 + (warp & 3) * 32;  // first row of this warp
      const long row = (long)row_w + lane;
      const bool row_ok = row < p.M;
      float head[3] = {0.f, 0.f, 0.f};
      const float r1v = (DGRAD && row_ok) ? __ldg(p.r1 + row) : 0.f;
      uint32_t held[64];  // first half's outputs wait in registers until the second half's MMAs have read ACT
      for (int s = 0; s < p.n_steps; s++) {
        const FusedParams::Step st = p.steps[s];
        const float* head_w = s_const + (st.head == 3 ? p.head_rgb_off : p.head_d_off);
        // TRAIN: 32 packed columns of this thread's row go to the warp's store box; every second call ships the box
        auto ship = [&](int col0, int half, const uint32_t* pk) {
          if (!TRAIN) return;
          if (half == 0) {
            if (lane == 0) tma_store_wait_read<0>();  // the previous box of this warp has been read out
            __syncwarp();
          }
          stage_half_row(slot_row, half, swz, pk);
          if (half == 1) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&p.map_act[s], slot, col0, row_w);
              tma_store_commit();
            }
          }
        };
        for (int h = 0; h < st.n_halves; h++) {
          const bool last_half = h == st.n_halves - 1;
          const float* bias = s_const + st.bias_off + h * 128;
          uint32_t m0, m1, m2, m3;
          uint4 mk = make_uint4(0u, 0u, 0u, 0u);  // DGRAD: ReLU mask words of this thread's 128 columns (requested before the wait)
          if (DGRAD && row_ok) mk = __ldg(reinterpret_cast<const uint4*>(p.bits[s] + row * (st.n_cols >> 5) + h * 4));  // dgrad outputs are 128 or 256 wide
          const float* v1 = (DGRAD && s == 0) ? s_const + p.head_d_off + h * 128 : nullptr;
          // one 32-column chunk c of this half: forward = bias + ReLU (+ heads, + mask out); dgrad = (+ r1 v1^T) * mask in
          auto chunk = [&](const uint32_t (&r)[32], int c, const float* hw, uint32_t* out16) -> uint32_t {
            if (DGRAD) {
              const uint32_t m = c == 0 ? mk.x : (c == 1 ? mk.y : (c == 2 ? mk.z : mk.w));
              dgrad_chunk(r, m, r1v, v1 ? v1 + c * 32 : nullptr, out16);
              return 0u;
            }
            return epi_chunk_any<MODE == 1>(st.head, r, bias + c * 32, hw + c * 32, head, out16);
          };
          mbar_wait(&acc_full, n_full & 1);
          n_full++;
          tc_fence_after_sync();
          if (!last_half) {
            // first half: pull the whole accumulator into registers at once so the second half's MMAs can start
            uint32_t r0[32], r1[32], r2[32], r3[32];
            tmem_ld_32x32(acc, r0); tmem_ld_32x32(acc + 32, r1); tmem_ld_32x32(acc + 64, r2); tmem_ld_32x32(acc + 96, r3);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty);
            m0 = chunk(r0, 0, head_w + h * 128, held);
            ship(h * 128, 0, held);
            m1 = chunk(r1, 1, head_w + h * 128, held + 16);
            ship(h * 128, 1, held + 16);
            m2 = chunk(r2, 2, head_w + h * 128, held + 32);
            ship(h * 128 + 64, 0, held + 32);
            m3 = chunk(r3, 3, head_w + h * 128, held + 48);
            ship(h * 128 + 64, 1, held + 48);
          } else {
            // every MMA of the layer is complete: ACT may be overwritten in place
            const bool unpark = st.n_halves == 2 && st.produces;
            if (unpark) {
#pragma unroll
              for (int c = 0; c < 4; c++) tmem_st_16(act + c * 16, held + 16 * c);
            }
            const uint32_t out = act + h * 64;
            const float* hw = head_w + (st.head == 3 ? 0 : h * 128);
            // TRAIN with kLateShipF: the packed words wait in `held` (free here: the parked half went into ACT above) and the
            // restaging + TMA stores of all four chunks move BEHIND the act_ready signal the next layer's k-blocks 2.. wait for
            constexpr bool late = kLateShipF && TRAIN;
            uint32_t ra[32], rb[32], pk1[16];
            uint32_t* pk0 = late ? held : pk1; uint32_t* pkA = late ? held + 16 : pk1; uint32_t* pkB = late ? held + 32 : pk1; uint32_t* pkC = late ? held + 48 : pk1;
            tmem_ld_32x32(acc, ra);
            tmem_ld_wait();
            if (unpark) {  // the next layer's k-blocks 0,1 may start
              tmem_st_wait();
              tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) mbar_arrive(&act_lo_ready);
            }
            tmem_ld_32x32(acc + 32, rb);
            m0 = chunk(ra, 0, hw, pk0);
            if (st.produces) tmem_st_16(out, pk0);
            if (!late) ship(h * 128, 0, pk0);
            tmem_ld_wait();
            tmem_ld_32x32(acc + 64, ra);
            m1 = chunk(rb, 1, hw, pkA);
            if (st.produces) tmem_st_16(out + 16, pkA);
            if (!late) ship(h * 128, 1, pkA);
            tmem_ld_wait();
            tmem_ld_32x32(acc + 96, rb);
            m2 = chunk(ra, 2, hw, pkB);
            if (st.produces) tmem_st_16(out + 32, pkB);
            if (!late) ship(h * 128 + 64, 0, pkB);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty);
            m3 = chunk(rb, 3, hw, pkC);
            if (st.produces) {
              tmem_st_16(out + 48, pkC);
              tmem_st_wait();
              tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) {
                if (st.n_halves == 1) mbar_arrive(&act_lo_ready);  // a one-half layer has no parked first instalment
                mbar_arrive(&act_ready);
              }
            }
            if (late) { ship(h * 128, 0, pk0); ship(h * 128, 1, pkA); ship(h * 128 + 64, 0, pkB); }
            ship(h * 128 + 64, 1, pkC);
          }
          if (MODE == 1 && row_ok) {
            uint32_t* bw = p.bits[s] + row * (st.n_cols >> 5) + h * 4;
            if (st.n_cols >= 128) *reinterpret_cast<uint4*>(bw) = make_uint4(m0, m1, m2, m3);
            else *reinterpret_cast<uint2*>(bw) = make_uint2(m0, m1);  // 64-wide layer: two mask words per row
          }
        }
        if (DGRAD) continue;
        if (st.head == 1) {
          if (row_ok) p.raw_density[row] = head[0] + s_const[p.head_d_off + 256];
          head[0] = 0.f;
        } else if (st.head == 3) {
          if (row_ok) {
#pragma unroll
            for (int n = 0; n < 3; n++) p.raw_rgb[row * 3 + n] = head[n] + s_const[p.head_rgb_off + 3 * 128 + n];
          }
          head[0] = head[1] = head[2] = 0.f;
        }
      }
    }
    if (TRAIN && lane == 0) tma_store_wait_read<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA leaves while its peer may still signal its barriers
  if (warp == 8) tmem_dealloc<512>(tmem_base);
}

template <int MODE>
int launch_fused(const FusedParams& p, int grid, int threads, size_t smem, bool pair, cudaStream_t st) {
  const void* kern = pair ? (const void*)k_mlp_fused_fwd<MODE, 2> : (const void*)k_mlp_fused_fwd<MODE, 1>;
  FusedParams pp = p;
  return launch_persistent_clusters(kern, grid, threads, smem, 226 * 1024, pair ? 2 : 1, &pp, st);
}

}  // namespace

// Host side: build the step table for a net of `D` trunk layers of width 256 with skips, one condition layer of
// width 128, and launch.  wplanes[s] = bf16 weight plane of dense layer s (trunk 0..D-1, then the condition layer),
// [N, kpad[s]] row-major.  consts layout: bias of every step (256 or 128 floats each), density head w[256], b[1],
// rgb head w[3][128], b[3].
int launch_mlp_fused_forward(const __nv_bfloat16* pos, int pos_pitch, const __nv_bfloat16* dir, int dir_pitch,
                             const __nv_bfloat16* const* wplanes, const int* kpad, const int* in_b, int D, int W, int Wc, long M,
                             const float* consts_dev, int n_consts, int head_d_off, int head_rgb_off, const int* bias_off,
                             float* raw_density, float* raw_rgb, __nv_bfloat16* const* act_out, uint32_t* const* bits_out,
                             const RaySource* rays, long enc_scratch_rows, bool pair, cudaStream_t st) {
  if (!((W == 256 && Wc == 128) || (W == 128 && Wc == 64)) || D + 1 > kMaxSteps || pos_pitch != 128 || dir_pitch != 64) {
    set_error("fused forward supports widths 256/128 and 128/64 (trunk / condition), position pitch 128, direction pitch 64");
    return 100001;
  }
  const bool train = act_out != nullptr;
  const size_t smem = (size_t)(train ? kWStages - 1 : kWStages) * kWStageBytes + 2 * kEncBytes + (train ? 8 * kStageSlot : 0) +
                      (size_t)((n_consts + 3) / 4 * 4) * sizeof(float) + 1024;
  const int sms = device_sm_count();
  if (smem > 226 * 1024) { set_error("fused forward: %zu bytes of shared memory needed", smem); return 100001; }
  FusedParams p;
  memset(&p, 0, sizeof(p));
  const int pairs = (int)cdiv(M, 256);
  const int grid = pair ? ((pairs < sms ? pairs : sms) + 1) / 2 * 2 : (pairs < sms ? pairs : sms);  // whole clusters
  // rays != nullptr: the kernel's encoder warps build the encodings from the t-values; `pos` / `dir` are then the level's
  // planes (training) or, with enc_scratch_rows > 0, a scratch of that many rows (>= grid * 512) that never leaves L2
  long map_rows = M;
  if (rays) {
    p.enc_mode = enc_scratch_rows > 0 ? 2 : 1;
    p.rs = *rays; p.enc_pos = const_cast<__nv_bfloat16*>(pos); p.enc_dir = const_cast<__nv_bfloat16*>(dir);
    if (p.enc_mode == 2) {
      if (enc_scratch_rows < (long)grid * 512) { set_error("fused forward: encoding scratch too small"); return 100001; }
      map_rows = enc_scratch_rows;
    }
    if (rays->deg_point % 4 || rays->deg_point * 6 > 120 || rays->deg_view > 4 || (long)rays->R * rays->S < M) {
      set_error("fused forward: in-kernel encoding needs deg_point %% 4 == 0, <= 20, deg_view <= 4"); return 100001;
    }
  }
  NERF_TRY(tc_make_tmap(&p.map_pos, pos, map_rows, 128, pos_pitch, 128));
  NERF_TRY(tc_make_tmap(&p.map_dir, dir, map_rows, 64, dir_pitch, 128));
  for (int s = 0; s <= D; s++) {
    const int N = s < D ? W : Wc;
    NERF_TRY(tc_make_tmap(&p.map_w[s], wplanes[s], N, kpad[s], kpad[s], 128));
    NERF_TRY(tc_make_tmap(&p.map_w64[s], wplanes[s], N, kpad[s], kpad[s], 64));
    if (train) {
      NERF_TRY(tc_make_tmap(&p.map_act[s], act_out[s], M, N, N, 32));
      p.bits[s] = bits_out[s];
    }
    FusedParams::Step& stp = p.steps[s];
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
  return train ? launch_fused<1>(p, grid, kThreadsFE, smem, pair, st) : launch_fused<0>(p, grid, kThreadsFE, smem, pair, st);
}


// The backward dgrad chain of the trunk as one kernel (bf16 mode): step 0 takes dZ of the condition layer [M, Wc] and
// W_cond^T, adds the density head's rank-1 term d_raw_density w_d^T and masks with the ReLU bits of the last trunk layer;
// steps 1..D-1 walk the trunk downwards.  dz_out[i] receives dZ of trunk layer D-1-i (what the wgrad of that layer reads);
// mask_bits[i] is the bit plane of that layer's activations.  wt[0] = W_cond^T [W, Wc]; wt[i] = W_{D-i}^T [W, W].
int launch_mlp_fused_dgrad(const __nv_bfloat16* dz_cond, int dz_cond_pitch, const __nv_bfloat16* const* wt, const int* wt_pitch, int D, int W,
                           int Wc, long M, const float* consts_dev, int n_consts, int head_d_off, const float* d_raw_density,
                           __nv_bfloat16* const* dz_out, const uint32_t* const* mask_bits, bool pair, cudaStream_t st) {
  if (!((W == 256 && Wc == 128) || (W == 128 && Wc == 64)) || D > kMaxSteps || D < 2) { set_error("fused dgrad supports widths 256/128 and 128/64"); return 100001; }
  const int sms = device_sm_count();
  const size_t smem = (size_t)(kWStages - 1) * kWStageBytes + 2 * kEncBytes + 8 * kStageSlot + (size_t)((n_consts + 3) / 4 * 4) * sizeof(float) + 1024;
  if (smem > 226 * 1024) { set_error("fused dgrad: %zu bytes of shared memory needed", smem); return 100001; }
  FusedParams p;
  memset(&p, 0, sizeof(p));
  NERF_TRY(tc_make_tmap(&p.map_pos, dz_cond, M, Wc, dz_cond_pitch, 128));  // A of step 0, loaded like the position encoding
  for (int s = 0; s < D; s++) {
    NERF_TRY(tc_make_tmap(&p.map_w[s], wt[s], W, s == 0 ? Wc : W, wt_pitch[s], 128));
    NERF_TRY(tc_make_tmap(&p.map_w64[s], wt[s], W, s == 0 ? Wc : W, wt_pitch[s], 64));
    NERF_TRY(tc_make_tmap(&p.map_act[s], dz_out[s], M, W, W, 32));
    p.bits[s] = const_cast<uint32_t*>(mask_bits[s]);
    FusedParams::Step& stp = p.steps[s];
    if (s == 0) { stp.n_act_kb = 0; stp.enc_kind = 1; stp.n_enc_kb = (int16_t)(Wc / 64); }
    else { stp.n_act_kb = (int16_t)(W / 64); stp.enc_kind = 0; stp.n_enc_kb = 0; }
    stp.n_halves = (int16_t)(W / 128);
    stp.n_cols = (int16_t)W;
    stp.produces = s < D - 1 ? 1 : 0;
    stp.head = 0; stp.bias_off = 0;
  }
  p.n_steps = D; p.M = M; p.consts = consts_dev; p.n_consts = n_consts;
  p.head_d_off = head_d_off; p.head_rgb_off = 0; p.r1 = d_raw_density;
  const int pairs = (int)cdiv(M, 256);
  return launch_fused<2>(p, pairs < sms ? pairs : sms, kThreadsF, smem, pair, st);
}

}  // namespace nerf

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

#include <mutex>

#include "gemm_tc.cuh"
#include "sm100.cuh"

namespace nerf {
namespace {

using namespace sm100;

constexpr int kThreadsF = 192;
constexpr int kWStages = 6;
constexpr int kWStageBytes = 128 * 128;           // [128 rows (N half) x 64 bf16]
constexpr int kEncBytes = 2 * 16384 + 16384;      // pos: 2 boxes of [128 x 64]; dir: 1 box
constexpr int kMaxSteps = 20;

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace

struct alignas(64) FusedParams {
  CUtensorMap map_pos, map_dir;     // encodings [M, 128] / [M, 64] bf16, box {64, 128}
  CUtensorMap map_w[kMaxSteps];     // weight planes [N, Kpad] bf16, box {64, 128}
  struct Step {
    int16_t n_act_kb;   // k-blocks read from the TMEM-resident activation (0 or width/64)
    int16_t enc_kind;   // 0 none, 1 position encoding, 2 direction encoding (shared-memory A operand)
    int16_t n_enc_kb;   // k-blocks of that encoding
    int16_t n_halves;   // N / 128
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
};

namespace {

__global__ void __launch_bounds__(kThreadsF, 1) k_mlp_fused_fwd(const __grid_constant__ FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_full[kWStages], w_empty[kWStages], enc_full[2], enc_empty[2], acc_full[2], acc_empty[2], act_ready[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* w_ring = smem;                                   // kWStages x 16 KB
  uint8_t* enc_buf = smem + kWStages * kWStageBytes;        // 2 x 48 KB
  float* s_const = reinterpret_cast<float*>(enc_buf + 2 * kEncBytes);

  const int n_tiles = (int)((p.M + 127) / 128);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWStages; s++) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int b = 0; b < 2; b++) {
      mbar_init(&enc_full[b], 1); mbar_init(&enc_empty[b], 1);
      mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4);
      mbar_init(&act_ready[b], 4);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.n_consts; i += kThreadsF) s_const[i] = __ldg(p.consts + i);
  if (warp == 4) {
    tmem_alloc<512>(&tmem_base_smem);
    if (lane == 0) {
      prefetch_tmap(&p.map_pos); prefetch_tmap(&p.map_dir);
      for (int s = 0; s < p.n_steps; s++) prefetch_tmap(&p.map_w[s]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t ACC0 = tmem_base, ACT0 = tmem_base + 256;  // ACC[h] = ACC0 + 128 h; ACT[b] = ACT0 + 128 b

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer: encodings per tile + weight ring
    if (lane == 0) {
      uint32_t wit = 0, tl = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, tl++) {
        const int row0 = tile * 128;
        const uint32_t eb = tl & 1;
        mbar_wait(&enc_empty[eb], ((tl >> 1) & 1) ^ 1);
        uint8_t* e = enc_buf + (size_t)eb * kEncBytes;
        mbar_arrive_expect_tx(&enc_full[eb], kEncBytes);
        tma_load_2d(e, &p.map_pos, 0, row0, &enc_full[eb]);
        tma_load_2d(e + 16384, &p.map_pos, 64, row0, &enc_full[eb]);
        tma_load_2d(e + 32768, &p.map_dir, 0, row0, &enc_full[eb]);
        for (int s = 0; s < p.n_steps; s++) {
          const FusedParams::Step st = p.steps[s];
          const int n_kb = st.n_act_kb + st.n_enc_kb;
          for (int h = 0; h < st.n_halves; h++)
            for (int kb = 0; kb < n_kb; kb++, wit++) {
              const int ws = wit % kWStages;
              mbar_wait(&w_empty[ws], ((wit / kWStages) & 1) ^ 1);
              mbar_arrive_expect_tx(&w_full[ws], kWStageBytes);
              tma_load_2d(w_ring + (size_t)ws * kWStageBytes, &p.map_w[s], kb * 64, h * 128, &w_full[ws]);
            }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 128, false, false);
      uint32_t wit = 0, tl = 0, acc_uses[2] = {0, 0}, act_waits[2] = {0, 0};
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, tl++) {
        const uint32_t eb = tl & 1;
        const uint32_t e_base = smem_u32(enc_buf + (size_t)eb * kEncBytes);
        bool enc_waited = false;
        for (int s = 0; s < p.n_steps; s++) {
          const FusedParams::Step st = p.steps[s];
          const uint32_t act_cur = ACT0 + 128 * (s & 1);  // written by the epilogue of step s-1
          bool act_ok[2] = {false, false};
          for (int h = 0; h < st.n_halves; h++) {
            mbar_wait(&acc_empty[h], (acc_uses[h] & 1) ^ 1);  // epilogue drained this accumulator half
            acc_uses[h]++;
            tc_fence_after_sync();
            const uint32_t acc = ACC0 + 128 * h;
            const int n_kb = st.n_act_kb + st.n_enc_kb;
            for (int kb = 0; kb < n_kb; kb++, wit++) {
              const bool from_act = kb < st.n_act_kb;
              if (from_act) {
                const int hh = kb >> 1;  // k-blocks 0,1 read the half-0 output of the previous step; 2,3 half 1
                if (!act_ok[hh]) {
                  mbar_wait(&act_ready[hh], act_waits[hh] & 1);
                  act_waits[hh]++;
                  act_ok[hh] = true;
                  tc_fence_after_sync();
                }
              } else if (!enc_waited) {
                mbar_wait(&enc_full[eb], (tl >> 1) & 1);
                enc_waited = true;
                tc_fence_after_sync();
              }
              const int ws = wit % kWStages;
              mbar_wait(&w_full[ws], (wit / kWStages) & 1);
              tc_fence_after_sync();
              const uint32_t b_base = smem_u32(w_ring + (size_t)ws * kWStageBytes);
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const uint64_t db = make_smem_desc(b_base + k * 32, 16, 1024);
                const uint32_t accum = (kb > 0 || k > 0) ? 1u : 0u;
                if (from_act) {
                  umma_bf16_ts(acc, act_cur + kb * 32 + k * 8, db, idesc, accum);
                } else {
                  const int ekb = kb - st.n_act_kb;
                  const uint32_t a_base = e_base + (st.enc_kind == 2 ? 32768u : (uint32_t)ekb * 16384u);
                  umma_bf16(acc, make_smem_desc(a_base + k * 32, 16, 1024), db, idesc, accum);
                }
              }
              umma_commit(&w_empty[ws]);
            }
            umma_commit(&acc_full[h]);
          }
        }
        umma_commit(&enc_empty[eb]);  // every MMA that read this tile's encodings has completed
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: warp w owns rows 32w..32w+31
    uint32_t full_uses[2] = {0, 0};
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long row = (long)tile * 128 + threadIdx.x;
      const bool row_ok = row < p.M;
      const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
      float head[3] = {0.f, 0.f, 0.f};
      for (int s = 0; s < p.n_steps; s++) {
        const FusedParams::Step st = p.steps[s];
        const uint32_t act_next = ACT0 + 128 * ((s + 1) & 1);
        for (int h = 0; h < st.n_halves; h++) {
          mbar_wait(&acc_full[h], full_uses[h] & 1);
          full_uses[h]++;
          tc_fence_after_sync();
          const uint32_t acc = ACC0 + 128 * h + lane_off;
#pragma unroll
          for (int c = 0; c < 4; c++) {
            uint32_t r[32];
            tmem_ld_32x32(acc + c * 32, r);
            tmem_ld_wait();
            const int col = h * 128 + c * 32;
            const float4* bv = reinterpret_cast<const float4*>(s_const + st.bias_off + col);
            float v[32];
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const float4 bb = bv[q];
              v[4 * q] = fmaxf(__uint_as_float(r[4 * q]) + bb.x, 0.f);
              v[4 * q + 1] = fmaxf(__uint_as_float(r[4 * q + 1]) + bb.y, 0.f);
              v[4 * q + 2] = fmaxf(__uint_as_float(r[4 * q + 2]) + bb.z, 0.f);
              v[4 * q + 3] = fmaxf(__uint_as_float(r[4 * q + 3]) + bb.w, 0.f);
            }
            if (st.head == 1) {  // density head: raw = y . w + b (N = 1)
              const float4* hv = reinterpret_cast<const float4*>(s_const + p.head_d_off + col);
#pragma unroll
              for (int q = 0; q < 8; q++) {
                const float4 w = hv[q];
                head[0] = fmaf(v[4 * q], w.x, head[0]); head[0] = fmaf(v[4 * q + 1], w.y, head[0]);
                head[0] = fmaf(v[4 * q + 2], w.z, head[0]); head[0] = fmaf(v[4 * q + 3], w.w, head[0]);
              }
            } else if (st.head == 3) {  // rgb head (N = 3) over the 128 condition features
#pragma unroll
              for (int n = 0; n < 3; n++) {
                const float4* hv = reinterpret_cast<const float4*>(s_const + p.head_rgb_off + n * 128 + col);
                float a = head[n];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                  const float4 w = hv[q];
                  a = fmaf(v[4 * q], w.x, a); a = fmaf(v[4 * q + 1], w.y, a);
                  a = fmaf(v[4 * q + 2], w.z, a); a = fmaf(v[4 * q + 3], w.w, a);
                }
                head[n] = a;
              }
            }
            if (st.produces) {  // next layer's A operand: bf16 pairs, two per 32-bit TMEM column
              uint32_t w16[16];
#pragma unroll
              for (int q = 0; q < 16; q++) w16[q] = pack2(v[2 * q], v[2 * q + 1]);
              tmem_st_32x16(act_next + lane_off + (uint32_t)(col >> 1), w16);
            }
          }
          if (st.produces) tmem_st_wait();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&acc_empty[h]);
            if (st.produces) mbar_arrive(&act_ready[h]);
          }
        }
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
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tmem_base);
}

}  // namespace

// Host side: build the step table for a net of `D` trunk layers of width 256 with skips, one condition layer of
// width 128, and launch.  wplanes[s] = bf16 weight plane of dense layer s (trunk 0..D-1, then the condition layer),
// [N, kpad[s]] row-major.  consts layout: bias of every step (256 or 128 floats each), density head w[256], b[1],
// rgb head w[3][128], b[3].
int launch_mlp_fused_forward(const __nv_bfloat16* pos, int pos_pitch, const __nv_bfloat16* dir, int dir_pitch,
                             const __nv_bfloat16* const* wplanes, const int* kpad, const int* in_b, int D, int W, int Wc, long M,
                             const float* consts_dev, int n_consts, int head_d_off, int head_rgb_off, const int* bias_off,
                             float* raw_density, float* raw_rgb, cudaStream_t st) {
  if (W != 256 || Wc != 128 || D + 1 > kMaxSteps || pos_pitch != 128 || dir_pitch != 64) {
    set_error("fused forward supports width 256 / condition width 128 / position pitch 128 / direction pitch 64");
    return 100001;
  }
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  static int sms = 148;
  const size_t smem = (size_t)kWStages * kWStageBytes + 2 * kEncBytes + (size_t)((n_consts + 3) / 4 * 4) * sizeof(float) + 1024;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(k_mlp_fused_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  });
  if (attr_err != cudaSuccess) { set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err)); return (int)attr_err; }
  if (smem > 226 * 1024) { set_error("fused forward: %zu bytes of shared memory needed", smem); return 100001; }
  FusedParams p;
  memset(&p, 0, sizeof(p));
  NERF_TRY(tc_make_tmap(&p.map_pos, pos, M, 128, pos_pitch, 128));
  NERF_TRY(tc_make_tmap(&p.map_dir, dir, M, 64, dir_pitch, 128));
  for (int s = 0; s <= D; s++) {
    const int N = s < D ? W : Wc;
    NERF_TRY(tc_make_tmap(&p.map_w[s], wplanes[s], N, kpad[s], kpad[s], 128));
    FusedParams::Step& stp = p.steps[s];
    if (s == 0) { stp.n_act_kb = 0; stp.enc_kind = 1; stp.n_enc_kb = 2; }
    else if (s < D) { stp.n_act_kb = 4; stp.enc_kind = in_b[s] ? 1 : 0; stp.n_enc_kb = in_b[s] ? 2 : 0; }
    else { stp.n_act_kb = 4; stp.enc_kind = 2; stp.n_enc_kb = 1; }
    stp.n_halves = (int16_t)(N / 128);
    stp.produces = s < D ? 1 : 0;
    stp.head = s == D - 1 ? 1 : (s == D ? 3 : 0);
    stp.bias_off = bias_off[s];
  }
  p.n_steps = D + 1; p.M = M; p.consts = consts_dev; p.n_consts = n_consts;
  p.head_d_off = head_d_off; p.head_rgb_off = head_rgb_off;
  p.raw_density = raw_density; p.raw_rgb = raw_rgb;
  const int tiles = (int)cdiv(M, 128);
  k_mlp_fused_fwd<<<tiles < sms ? tiles : sms, kThreadsF, smem, st>>>(p);
  NERF_CHECK_LAUNCH();
  return 0;
}

}  // namespace nerf

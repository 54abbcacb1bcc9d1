// gemm_simt.cu — strict-fp32 CUDA-core MLP layers (NERF_PRECISION_FP32).
//
// Replaces get_neuron_output{,_sigmoid,_soft_plus,_conjoined_inputs} (.cu:36-90) and
// backpropagate_neuron{,_sigmoid,_soft_plus,_partial_conjoined} (.cu:91-182).  The reference computes
// one K-long dot product per thread straight from global memory (forward) and two global float
// atomicAdds per multiply (backward).  Here every layer is a shared-memory-tiled SGEMM
// (128x128x16 tiles, 8x8 register micro-tiles, register-staged double buffering):
//   forward  Y = act([A1|A2] W^T + b)              "NT"  both operands reduction-contiguous
//   dgrad    dX = dZ W[:, :k1]  (+ rank-1, mask)     "NN"  W read output-contiguous
//   wgrad    dW += dZ^T [A1|A2]                      "TN"  both operands output-contiguous, split over M,
//                                                          deterministic two-pass reduction (no atomics)
// This is the fp32 arithmetic baseline / cross-check of the tcgen05 path (gemm_tc.cu), and also hosts the
// thin heads (N = 1 density, N = 3 rgb) which are too narrow for UMMA tiles.
#include "kernels.cuh"

namespace nerf {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, NT = 256;

struct Operand {
  const float* p;
  int ld;
  bool vec;  // 16-byte aligned base and ld % 4 == 0
};
inline Operand make_operand(const float* p, int ld) {
  return Operand{p, ld, p != nullptr && (((uintptr_t)p & 15) == 0) && (ld % 4 == 0)};
}

struct Segment {  // one reduction segment: r in [0, len)
  Operand a, b;
  int len;
};

struct Epilogue {
  int mode;  // 0 forward, 1 dgrad, 2 raw partial store (split reduction)
  // forward
  const float* bias; float* Y; float* Z; int act;
  // dgrad
  const float* r1; const float* v1; const float* mask; int accumulate;
  // output
  float* C; int ldc;
};

__device__ __forceinline__ float apply_act(float z, int act) {
  switch (act) {
    case ACT_RELU: return z > 0.f ? z : 0.f;        // .cu:7
    case ACT_SIGMOID: return sigmoidf_(z);          // .cu:9
    case ACT_SOFTPLUS: return softplusf_(z);        // .cu:14
    default: return z;
  }
}

// Stage one [BK x 128] operand tile into registers (2 float4 per thread).
// KMAJ : X(i,r) = p[i*ld + r]   (reduction index contiguous)
// !KMAJ: X(i,r) = p[r*ld + i]   (output index contiguous)
template <bool KMAJ>
__device__ __forceinline__ void stage_load(const Operand& x, long i0, long I, int r0, int rend, float4 (&reg)[2]) {
  const int t = threadIdx.x;
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KMAJ) {
      const long i = i0 + (t >> 2) + 64 * pass;
      const int r = r0 + 4 * (t & 3);
      if (i < I && r < rend) {
        const float* src = x.p + i * x.ld + r;
        if (x.vec) {
          v = __ldg(reinterpret_cast<const float4*>(src));
          if (r + 1 >= rend) v.y = 0.f;
          if (r + 2 >= rend) v.z = 0.f;
          if (r + 3 >= rend) v.w = 0.f;
        } else {
          v.x = __ldg(src);
          if (r + 1 < rend) v.y = __ldg(src + 1);
          if (r + 2 < rend) v.z = __ldg(src + 2);
          if (r + 3 < rend) v.w = __ldg(src + 3);
        }
      }
    } else {
      const int r = r0 + (t >> 5) + 8 * pass;
      const long i = i0 + 4 * (t & 31);
      if (r < rend && i < I) {
        const float* src = x.p + (long)r * x.ld + i;
        if (x.vec && i + 3 < I) {
          v = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          v.x = __ldg(src);
          if (i + 1 < I) v.y = __ldg(src + 1);
          if (i + 2 < I) v.z = __ldg(src + 2);
          if (i + 3 < I) v.w = __ldg(src + 3);
        }
      }
    }
    reg[pass] = v;
  }
}
template <bool KMAJ>
__device__ __forceinline__ void stage_store(float (*sm)[BM + PAD], const float4 (&reg)[2]) {
  const int t = threadIdx.x;
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    if (KMAJ) {
      const int row = (t >> 2) + 64 * pass, kq = 4 * (t & 3);
      sm[kq][row] = reg[pass].x; sm[kq + 1][row] = reg[pass].y; sm[kq + 2][row] = reg[pass].z; sm[kq + 3][row] = reg[pass].w;
    } else {
      const int rr = (t >> 5) + 8 * pass;
      *reinterpret_cast<float4*>(&sm[rr][4 * (t & 31)]) = reg[pass];
    }
  }
}

// C(i,j) = sum over segments, r of A(i,r) B(j,r);  grid = (ceil(J/128), ceil(I/128), splits)
template <bool A_K, bool B_K>
__global__ void __launch_bounds__(NT)
k_sgemm(Segment s0, Segment s1, int nseg, long I, int J, long split_len, Epilogue ep) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long i0 = (long)blockIdx.y * BM;
  const long j0 = (long)blockIdx.x * BN;
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; a++)
#pragma unroll
    for (int b = 0; b < 8; b++) acc[a][b] = 0.f;

  for (int sg = 0; sg < nseg; sg++) {
    const Segment& s = sg == 0 ? s0 : s1;
    // split over the reduction (only used with one segment)
    int rbeg = 0, rend = s.len;
    if (split_len > 0) {
      rbeg = (int)min((long)s.len, (long)blockIdx.z * split_len);
      rend = (int)min((long)s.len, (long)(blockIdx.z + 1) * split_len);
    }
    const int ntiles = (rend - rbeg + BK - 1) / BK;
    if (ntiles <= 0) continue;
    float4 ra[2], rb[2];
    stage_load<A_K>(s.a, i0, I, rbeg, rend, ra);
    stage_load<B_K>(s.b, j0, J, rbeg, rend, rb);
    stage_store<A_K>(As[0], ra);
    stage_store<B_K>(Bs[0], rb);
    __syncthreads();
    for (int kt = 0; kt < ntiles; kt++) {
      const int cur = kt & 1;
      if (kt + 1 < ntiles) {
        stage_load<A_K>(s.a, i0, I, rbeg + (kt + 1) * BK, rend, ra);
        stage_load<B_K>(s.b, j0, J, rbeg + (kt + 1) * BK, rend, rb);
      }
#pragma unroll
      for (int k = 0; k < BK; k++) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
          for (int b = 0; b < 8; b++) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
      }
      if (kt + 1 < ntiles) {
        stage_store<A_K>(As[cur ^ 1], ra);
        stage_store<B_K>(Bs[cur ^ 1], rb);
      }
      __syncthreads();
    }
  }

  // epilogue
  float* C = ep.C;
  if (ep.mode == 2) C += (long)blockIdx.z * I * ep.ldc;
#pragma unroll
  for (int a = 0; a < 8; a++) {
    const long i = i0 + (a < 4 ? ty * 4 + a : 64 + ty * 4 + (a - 4));
    if (i >= I) continue;
    const float ri = (ep.mode == 1 && ep.r1) ? ep.r1[i] : 0.f;
#pragma unroll
    for (int bh = 0; bh < 2; bh++) {
      const long j = j0 + bh * 64 + tx * 4;
      float v[4] = {acc[a][bh * 4], acc[a][bh * 4 + 1], acc[a][bh * 4 + 2], acc[a][bh * 4 + 3]};
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const long jj = j + c;
        if (jj >= J) continue;
        float x = v[c];
        if (ep.mode == 0) {
          x += ep.bias ? ep.bias[jj] : 0.f;  // bias after the k-ascending sum (.cu:45)
          if (ep.Z) ep.Z[i * ep.ldc + jj] = x;
          if (ep.Y) ep.Y[i * ep.ldc + jj] = apply_act(x, ep.act);
        } else if (ep.mode == 1) {
          if (ep.r1) x += ri * ep.v1[jj];
          if (ep.accumulate) x += C[i * ep.ldc + jj];
          if (ep.mask) x = ep.mask[i * ep.ldc + jj] > 0.f ? x : 0.f;
          C[i * ep.ldc + jj] = x;
        } else {
          C[i * ep.ldc + jj] = x;
        }
      }
    }
  }
}

// out[i*ldo + coff + j] += sum_z ws[z][i*J + j]   (fixed order -> deterministic)
__global__ void k_reduce_partials(const float* __restrict__ ws, int splits, long I, int J, float* __restrict__ out,
                                  int ldo, int coff) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= I * J) return;
  float s = 0.f;
  for (int z = 0; z < splits; z++) s += ws[(long)z * I * J + idx];
  const long i = idx / J;
  const int j = (int)(idx % J);
  out[i * ldo + coff + j] += s;
}

// column sums of X[M,N] over row chunks: part[z][n]
__global__ void __launch_bounds__(256)
k_colsum_partial(const float* __restrict__ X, long M, int N, long chunk, float* __restrict__ part) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const long m0 = (long)blockIdx.y * chunk, m1 = min(M, m0 + chunk);
  float s = 0.f;
  if (n < N)
    for (long m = m0 + wy; m < m1; m += 8) s += __ldg(X + m * N + n);
  red[wy][lane] = s;
  __syncthreads();
  if (wy == 0 && n < N) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) tot += red[k][lane];
    part[(long)blockIdx.y * N + n] = tot;
  }
}

__global__ void k_act_grad(const float* __restrict__ dY, const float* __restrict__ Z, float* __restrict__ dZ, long n, int act) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float z = Z[i];
  float d;
  switch (act) {
    case ACT_RELU: d = z > 0.f ? 1.f : 0.f; break;                                  // .cu:8
    case ACT_SIGMOID: { const float s = sigmoidf_(z); d = s * (1.f - s); } break;  // .cu:10-13
    case ACT_SOFTPLUS: d = sigmoidf_(z); break;                                     // .cu:141
    default: d = 1.f;
  }
  dZ[i] = dY[i] * d;
}

__global__ void k_relu_mask(const float* __restrict__ dY, const float* __restrict__ Y, float* __restrict__ dZ, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dZ[i] = Y[i] > 0.f ? dY[i] : 0.f;
}

// thin head forward: one warp per row, N <= 4 outputs
__global__ void __launch_bounds__(256)
k_thin_fwd(const float* __restrict__ X, int ldx, const float* __restrict__ W, const float* __restrict__ b,
           float* __restrict__ Y, long M, int N, int K) {
  const long m = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = lane; k < K; k += 32) {
    const float x = __ldg(X + m * ldx + k);
#pragma unroll
    for (int n = 0; n < 4; n++)
      if (n < N) acc[n] = fmaf(x, __ldg(W + n * K + k), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < 4; n++) {
    float v = acc[n];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && n < N) Y[m * N + n] = v + (b ? b[n] : 0.f);
  }
}

// thin dgrad: dX[m,k] (=|+=) sum_n dz[m,n] W[n,k], optional relu mask
__global__ void k_thin_dgrad(const float* __restrict__ dZ, const float* __restrict__ W, float* __restrict__ dX, long M,
                             int N, int K, const float* __restrict__ mask, int accumulate) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * K) return;
  const long m = idx / K;
  const int k = (int)(idx % K);
  float v = 0.f;
  for (int n = 0; n < N; n++) v = fmaf(__ldg(dZ + m * N + n), __ldg(W + n * K + k), v);
  if (accumulate) v += dX[idx];
  if (mask) v = mask[idx] > 0.f ? v : 0.f;
  dX[idx] = v;
}

// thin wgrad partials: block z handles rows [z*chunk, ...); thread k owns column k for all N outputs.
// part[z][n*K + k], partb[z][n]
__global__ void __launch_bounds__(256)
k_thin_wgrad_partial(const float* __restrict__ dZ, const float* __restrict__ X, int ldx, long M, int N, int K,
                     long chunk, float* __restrict__ part, float* __restrict__ partb) {
  const long m0 = (long)blockIdx.x * chunk, m1 = min(M, m0 + chunk);
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long m = m0; m < m1; m++) {
      const float x = __ldg(X + m * ldx + k);
#pragma unroll
      for (int n = 0; n < 4; n++)
        if (n < N) acc[n] = fmaf(__ldg(dZ + m * N + n), x, acc[n]);
    }
    for (int n = 0; n < N; n++) part[(long)blockIdx.x * N * K + n * K + k] = acc[n];
  }
  if (threadIdx.x < N) {
    float s = 0.f;
    for (long m = m0; m < m1; m++) s += __ldg(dZ + m * N + threadIdx.x);
    partb[(long)blockIdx.x * N + threadIdx.x] = s;
  }
}

int pick_splits(long tiles, long M) {
  // aim for ~4 CTAs per SM over 148 SMs, at least 512 rows per split
  long want = cdiv(148 * 4, tiles);
  long maxs = cdiv(M, 512);
  long s = want < maxs ? want : maxs;
  return (int)(s < 1 ? 1 : s);
}

}  // namespace

int launch_dense_fwd(const float* A1, int lda1, int k1, const float* A2, int lda2, int k2, const float* W,
                     const float* b, float* Y, float* Z, long M, int N, Act act, cudaStream_t st) {
  const int ldw = k1 + k2;
  Segment s0{make_operand(A1, lda1), make_operand(W, ldw), k1};
  Segment s1{make_operand(A2, lda2), make_operand(W ? W + k1 : nullptr, ldw), k2};
  // the second segment of W starts at column k1: only vectorisable when k1 % 4 == 0 (true for 256/512)
  s1.b.vec = s1.b.vec && (k1 % 4 == 0);
  Epilogue ep{};
  ep.mode = 0; ep.bias = b; ep.Y = Y; ep.Z = Z; ep.act = (int)act; ep.ldc = N;
  const dim3 grid((unsigned)cdiv(N, BN), (unsigned)cdiv(M, BM), 1);
  k_sgemm<true, true><<<grid, NT, 0, st>>>(s0, s1, (A2 && k2 > 0) ? 2 : 1, M, N, 0, ep);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_dense_dgrad(const float* dZ, const float* W, int ldw, float* dX, long M, int N, int k1, const float* r1,
                       const float* v1, const float* mask, bool accumulate, cudaStream_t st) {
  // dX(i=m, j=k) = sum_{r=n} dZ[m*N + n] * W[n*ldw + k]   -> A reduction-contiguous, B output-contiguous
  Segment s0{make_operand(dZ, N), make_operand(W, ldw), N};
  Epilogue ep{};
  ep.mode = 1; ep.r1 = r1; ep.v1 = v1; ep.mask = mask; ep.accumulate = accumulate ? 1 : 0; ep.C = dX; ep.ldc = k1;
  const dim3 grid((unsigned)cdiv(k1, BN), (unsigned)cdiv(M, BM), 1);
  k_sgemm<true, false><<<grid, NT, 0, st>>>(s0, s0, 1, M, k1, 0, ep);
  NERF_CHECK_LAUNCH();
  return 0;
}

size_t dense_wgrad_workspace(long M, int N, int K) {
  const long tiles = cdiv(N, BM) * cdiv(K, BN);
  const int splits = pick_splits(tiles, M);
  const size_t colsum = (size_t)cdiv(M, 4096) * N;
  return (size_t)splits * N * K + colsum + 64;
}

int launch_dense_wgrad(const float* dZ, const float* A1, int lda1, int k1, const float* A2, int lda2, int k2,
                       float* dW, float* db, long M, int N, float* workspace, cudaStream_t st) {
  const int ldw = k1 + k2;
  for (int src = 0; src < 2; src++) {
    const float* A = src == 0 ? A1 : A2;
    const int lda = src == 0 ? lda1 : lda2, K = src == 0 ? k1 : k2, coff = src == 0 ? 0 : k1;
    if (!A || K <= 0) continue;
    // dW(i=n, j=k) = sum_{r=m} dZ[m*N + n] * A[m*lda + k]  -> both output-contiguous
    Segment s0{make_operand(dZ, N), make_operand(A, lda), (int)M};
    const long tiles = cdiv(N, BM) * cdiv(K, BN);
    const int splits = pick_splits(tiles, M);
    const long split_len = cdiv(cdiv(M, splits), BK) * BK;
    Epilogue ep{};
    ep.mode = 2; ep.C = workspace; ep.ldc = K;
    const dim3 grid((unsigned)cdiv(K, BN), (unsigned)cdiv(N, BM), (unsigned)splits);
    k_sgemm<false, false><<<grid, NT, 0, st>>>(s0, s0, 1, N, K, split_len, ep);
    NERF_CHECK_LAUNCH();
    const long n = (long)N * K;
    k_reduce_partials<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(workspace, splits, N, K, dW, ldw, coff);
    NERF_CHECK_LAUNCH();
  }
  if (db) {
    const long chunk = 4096;
    const int nchunks = (int)cdiv(M, chunk);
    float* part = workspace;  // reuse (stream-ordered after the reductions above)
    k_colsum_partial<<<dim3((unsigned)cdiv(N, 32), (unsigned)nchunks), 256, 0, st>>>(dZ, M, N, chunk, part);
    NERF_CHECK_LAUNCH();
    k_reduce_partials<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(part, nchunks, 1, N, db, N, 0);
    NERF_CHECK_LAUNCH();
  }
  return 0;
}

int launch_act_grad(const float* dY, const float* Z, float* dZ, long n, Act act, cudaStream_t st) {
  k_act_grad<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(dY, Z, dZ, n, (int)act);
  NERF_CHECK_LAUNCH();
  return 0;
}
int launch_relu_mask(const float* dY, const float* Y, float* dZ, long n, cudaStream_t st) {
  k_relu_mask<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(dY, Y, dZ, n);
  NERF_CHECK_LAUNCH();
  return 0;
}

int launch_thin_fwd(const float* X, int ldx, const float* W, const float* b, float* Y, long M, int N, int K,
                    cudaStream_t st) {
  if (N > 4) { set_error("thin_fwd: N=%d > 4", N); return 100001; }
  k_thin_fwd<<<(unsigned)cdiv(M, 8), 256, 0, st>>>(X, ldx, W, b, Y, M, N, K);
  NERF_CHECK_LAUNCH();
  return 0;
}
int launch_thin_dgrad(const float* dZ, const float* W, float* dX, long M, int N, int K, const float* mask,
                      bool accumulate, cudaStream_t st) {
  k_thin_dgrad<<<(unsigned)cdiv(M * K, 256), 256, 0, st>>>(dZ, W, dX, M, N, K, mask, accumulate ? 1 : 0);
  NERF_CHECK_LAUNCH();
  return 0;
}
size_t thin_wgrad_workspace(long M, int N, int K) {
  const long chunks = cdiv(M, 1024);
  return (size_t)chunks * N * (K + 1) + 64;
}
int launch_thin_wgrad(const float* dZ, const float* X, int ldx, float* dW, float* db, long M, int N, int K,
                      float* workspace, cudaStream_t st) {
  if (N > 4) { set_error("thin_wgrad: N=%d > 4", N); return 100001; }
  const long chunk = 1024;
  const int chunks = (int)cdiv(M, chunk);
  float* part = workspace;
  float* partb = workspace + (size_t)chunks * N * K;
  k_thin_wgrad_partial<<<chunks, 256, 0, st>>>(dZ, X, ldx, M, N, K, chunk, part, partb);
  NERF_CHECK_LAUNCH();
  k_reduce_partials<<<(unsigned)cdiv((long)N * K, 256), 256, 0, st>>>(part, chunks, N, K, dW, K, 0);
  NERF_CHECK_LAUNCH();
  if (db) {
    k_reduce_partials<<<1, 256, 0, st>>>(partb, chunks, 1, N, db, N, 0);
    NERF_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace nerf

// rk_drp.cu -- deep reactive policy (DRP) kernels: dense layers and the action heads (sm_100a).
//
// Replaces the flax MLP behind JaxDeepReactivePolicy (planner.py:1342-1449: nn.Dense +
// activation stack, per-action mu / log_sigma heads) and the reverse pass jax.value_and_grad
// derives for it (planner.py:2873).  The MLP runs once per decision epoch between two
// launches of the rollout kernel (closed loop: actions depend on the current state).
//
// rk_sgemm is an fp32 SIMT GEMM (128x128x8 tiles, 8x8 register micro-tiles, float4 global
// loads, register-staged double buffering) with the layer epilogues fused in; results are
// bit-reproducible (fixed summation order, deterministic split-K through a workspace).
// It is the parity build of the dense layers; the tcgen05 (3xTF32) path replaces the
// hidden-layer GEMM where the batch is large enough to fill 128-row MMA tiles.

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/rk.h"
#include "rk_act.cuh"

// The hidden-layer activation (rk_act.cuh ids) is an argument of every launcher -- bits 8..15 of the
// flags / epi word of the GEMMs, an explicit `act` elsewhere -- and of every kernel, so the library keeps
// no process state: calls are re-entrant per stream and a captured graph keeps its own value.
static inline bool rk_act_ok(int act) { return act >= 0 && act < RK_ACT_COUNT; }

// y = act(z), element-wise: the second half of a layer whose pre-activation is kept (gelu / silu)
__global__ void __launch_bounds__(256)
rk_act_apply_kernel(int act, size_t n, const float* __restrict__ z, float* __restrict__ y) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = rk_act_fn(act, z[i]);
}

extern "C" int rk_act_apply(void* stream, int act, size_t n, const float* z, float* y) {
  if (!rk_act_ok(act) || n == 0 || z == nullptr || y == nullptr) return RK_E_BADARG;
  rk_act_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(act, n, z, y);
  return (int)cudaGetLastError();
}

#define BM 128
#define BN 128
#define BK 8
#define TM 8
#define TN 8

// C[M][N] (row-major, ldc) = epi( op(A)[M][K] * B[K][N] )
//   A_TRANS = 0: A stored [M][K] (lda >= K);  1: A stored [K][M] (lda >= M)
//   split-K: blockIdx.z handles K range [z*kchunk, (z+1)*kchunk); partial tiles go to
//   `out` = workspace + z*M*N (ldc = N) and the epilogue is applied by the reduce kernel.
template <bool A_TRANS, int ACT_CT>
__global__ void __launch_bounds__(256)
rk_sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                const float* __restrict__ bias, const float* __restrict__ aux, int ldaux,
                int epi, int accumulate, int kchunk, float* __restrict__ workspace) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, each 8 x 8 outputs

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // global -> register staging: each thread moves 4 A values and 4 B values per k-tile
  float ra[4], rb[4];
  auto load_tiles = [&](int k0) {
    if (A_TRANS) {                                 // A[k][m]: 8 rows of 128 floats
      const int kk = tid >> 5, mm = (tid & 31) * 4;
      const int k = k0 + kk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + mm + i;
        ra[i] = (k < kend && m < M) ? A[(size_t)k * lda + m] : 0.f;
      }
    } else {                                       // A[m][k]: 128 rows of 8 floats
      const int mm = tid >> 1, kk = (tid & 1) * 4;
      const int m = m0 + mm;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + kk + i;
        ra[i] = (m < M && k < kend) ? A[(size_t)m * lda + k] : 0.f;
      }
    }
    {
      const int kk = tid >> 5, nn = (tid & 31) * 4;
      const int k = k0 + kk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = n0 + nn + i;
        rb[i] = (k < kend && n < N) ? B[(size_t)k * ldb + n] : 0.f;
      }
    }
  };
  auto store_tiles = [&](int buf) {
    if (A_TRANS) {
      const int kk = tid >> 5, mm = (tid & 31) * 4;
      *reinterpret_cast<float4*>(&As[buf][kk][mm]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    } else {
      const int mm = tid >> 1, kk = (tid & 1) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) As[buf][kk + i][mm] = ra[i];
    }
    const int kk = tid >> 5, nn = (tid & 31) * 4;
    *reinterpret_cast<float4*>(&Bs[buf][kk][nn]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
  };

  int buf = 0;
  if (kbeg < kend) {
    load_tiles(kbeg);
    store_tiles(0);
  }
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) load_tiles(k0 + BK);                 // next tile's loads overlap the math
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  const bool split = gridDim.z > 1;
  float* out = split ? workspace + (size_t)blockIdx.z * M * N : C;
  const int ldo = split ? N : ldc;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = acc[i][j];
      if (!split) {                      // C_in + A*B first, then the layer epilogue
        if (accumulate) v += out[(size_t)m * ldo + n];
        if ((epi & 0xff) == 1 || (epi & 0xff) == 2) v += bias[n];
        if ((epi & 0xff) == 2) v = rk_act_fn(ACT_CT >= 0 ? ACT_CT : (epi >> 8), v);
        if ((epi & 0xff) == 3) { const float h = aux[(size_t)m * ldaux + n]; v *= rk_act_grad(ACT_CT >= 0 ? ACT_CT : (epi >> 8), h); }
      }
      out[(size_t)m * ldo + n] = v;
    }
  }
}

__global__ void rk_sgemm_reduce_kernel(int M, int N, int splits, const float* __restrict__ ws,
                                       float* __restrict__ C, int ldc,
                                       const float* __restrict__ bias,
                                       const float* __restrict__ aux, int ldaux, int epi,
                                       int accumulate) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  float v = 0.f;
  for (int z = 0; z < splits; ++z) v += ws[(size_t)z * M * N + i];      // fixed order
  if (accumulate) v += C[(size_t)m * ldc + n];
  if ((epi & 0xff) == 1 || (epi & 0xff) == 2) v += bias[n];
  if ((epi & 0xff) == 2) v = rk_act_fn(epi >> 8, v);
  if ((epi & 0xff) == 3) { const float h = aux[(size_t)m * ldaux + n]; v *= rk_act_grad(epi >> 8, h); }
  C[(size_t)m * ldc + n] = v;
}

// Skinny shapes (N <= 16: action heads, state cotangent blocks; M <= 16: bias and first-layer
// weight gradients): the 128x128 tile would spend > 90% of its FMAs on padding, and these
// products are bound by streaming the [B][h] operand once.  16 x 16 threads, thread (tx, ty)
// owns rows ty + 16 i and columns tx + 16 j of a BM_ x BN_ tile.
#define SBK 16
template <int BM_, int BN_, bool A_TRANS, int ACT_CT>
__global__ void __launch_bounds__(256)
rk_sgemm_skinny_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                       const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                       const float* __restrict__ bias, const float* __restrict__ aux, int ldaux,
                       int epi, int accumulate, int kchunk, float* __restrict__ workspace) {
  constexpr int TM_ = BM_ / 16, TN_ = BN_ / 16;
  __shared__ float As[SBK][BM_ + 1];
  __shared__ float Bs[SBK][BN_ + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM_, n0 = blockIdx.x * BN_;
  const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  float acc[TM_][TN_];
#pragma unroll
  for (int i = 0; i < TM_; ++i)
#pragma unroll
    for (int j = 0; j < TN_; ++j) acc[i][j] = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += SBK) {
    for (int idx = tid; idx < SBK * BM_; idx += 256) {
      int kk, mm;
      if (A_TRANS) { kk = idx / BM_; mm = idx % BM_; } else { mm = idx / SBK; kk = idx % SBK; }
      const int k = k0 + kk, m = m0 + mm;
      float v = 0.f;
      if (k < kend && m < M) v = A_TRANS ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k];
      As[kk][mm] = v;
    }
    for (int idx = tid; idx < SBK * BN_; idx += 256) {
      const int kk = idx / BN_, nn = idx % BN_;
      const int k = k0 + kk, n = n0 + nn;
      Bs[kk][nn] = (k < kend && n < N) ? B[(size_t)k * ldb + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      float a[TM_], b[TN_];
#pragma unroll
      for (int i = 0; i < TM_; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < TN_; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < TM_; ++i)
#pragma unroll
        for (int j = 0; j < TN_; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
  float* out = split ? workspace + (size_t)blockIdx.z * M * N : C;
  const int ldo = split ? N : ldc;
#pragma unroll
  for (int i = 0; i < TM_; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN_; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (!split) {
        if (accumulate) v += out[(size_t)m * ldo + n];
        if ((epi & 0xff) == 1 || (epi & 0xff) == 2) v += bias[n];
        if ((epi & 0xff) == 2) v = rk_act_fn(ACT_CT >= 0 ? ACT_CT : (epi >> 8), v);
        if ((epi & 0xff) == 3) { const float h = aux[(size_t)m * ldaux + n]; v *= rk_act_grad(ACT_CT >= 0 ? ACT_CT : (epi >> 8), h); }
      }
      out[(size_t)m * ldo + n] = v;
    }
  }
}

extern "C" int rk_sgemm(void* stream, int flags, int M, int N, int K, const float* A, int lda,
                        const float* B, int ldb, float* C, int ldc, const float* bias,
                        const float* aux, int ldaux, float* workspace, int split_k) {
  if (M <= 0 || N <= 0 || K <= 0 || A == nullptr || B == nullptr || C == nullptr)
    return RK_E_BADARG;
  const int a_trans = flags & RK_GEMM_A_TRANS, accumulate = (flags & RK_GEMM_ACCUMULATE) ? 1 : 0;
  const int epi = (flags >> 4) & 0xf;
  const int act_id = (flags >> 8) & 0xff;          // activation id of the tanh-family epilogues
  if (!rk_act_ok(act_id)) return RK_E_BADARG;
  if (epi > 3 || ((epi == 1 || epi == 2) && bias == nullptr) || (epi == 3 && aux == nullptr))
    return RK_E_BADARG;
  // tile shape: skinny variants when one output dimension is <= 16
  int bm = BM, bn = BN, bk = BK;
  if (N <= 16 && M > 16) { bm = 128; bn = 16; bk = SBK; }
  else if (M <= 16) { bm = 16; bn = 128; bk = SBK; }
  const int tiles = ((N + bn - 1) / bn) * ((M + bm - 1) / bm);
  if (split_k == 0) split_k = 1;
  if (split_k < 0) {                 // auto: enough CTAs for one wave, at most -split_k slices
    int want = (148 + tiles - 1) / tiles;        // one full wave of CTAs
    if (want > -split_k) want = -split_k;
    if (want > K / 256) want = K / 256;
    split_k = want < 1 ? 1 : want;
  }
  if (split_k > 1 && workspace == nullptr) return RK_E_BADARG;
  int kchunk = (K + split_k - 1) / split_k;
  kchunk = ((kchunk + bk - 1) / bk) * bk;
  const int splits = (K + kchunk - 1) / kchunk;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((N + bn - 1) / bn, (M + bm - 1) / bm, splits);
  // the tanh build of each kernel has the activation as a compile-time constant
#define RK_LAUNCH(KERNEL, ...)                                                                \
  do {                                                                                        \
    if (act_id == RK_ACT_TANH)                                                              \
      KERNEL<__VA_ARGS__, RK_ACT_TANH><<<grid, 256, 0, st>>>(                                 \
          M, N, K, A, lda, B, ldb, C, ldc, bias, aux, ldaux, epi, accumulate, kchunk,         \
          workspace);                                                                         \
    else                                                                                      \
      KERNEL<__VA_ARGS__, -1><<<grid, 256, 0, st>>>(                                          \
          M, N, K, A, lda, B, ldb, C, ldc, bias, aux, ldaux, epi | (act_id << 8),           \
          accumulate, kchunk, workspace);                                                     \
  } while (0)
  if (bn == 16) {
    if (a_trans) RK_LAUNCH(rk_sgemm_skinny_kernel, 128, 16, true);
    else RK_LAUNCH(rk_sgemm_skinny_kernel, 128, 16, false);
  } else if (bm == 16) {
    if (a_trans) RK_LAUNCH(rk_sgemm_skinny_kernel, 16, 128, true);
    else RK_LAUNCH(rk_sgemm_skinny_kernel, 16, 128, false);
  } else {
    if (a_trans) RK_LAUNCH(rk_sgemm_kernel, true);
    else RK_LAUNCH(rk_sgemm_kernel, false);
  }
#undef RK_LAUNCH
  if (splits > 1) {
    const size_t total = (size_t)M * N;
    rk_sgemm_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        M, N, splits, workspace, C, ldc, bias, aux, ldaux, epi | (act_id << 8), accumulate);
  }
  return (int)cudaGetLastError();
}

// ---- action heads (planner.py:1394-1429, 1472-1475) ---------------------------------------
// heads [B][2A]: columns [0, A) = mu, [A, 2A) = log_sigma (pre-clip).  Actions, their noise and
// their cotangents use the rollout kernels' fluent-major layout: action fluent f with n_f
// elements is the block [B][n_f] at word offset off_f * B (include/rk.h), so head column
// a = off_f + e of rollout b lives at off_f * B + b * n_f + e.
// action = clip(mu + train * exp(clip(log_sigma)) * noise, lo, hi);
// entropy[b] = sum_a clip(log_sigma)   (value only: stop_gradient unless sigma_entropy_grad)
struct RkActSegs {
  int n;
  int off[RK_MAX_ACTION_FLUENTS], len[RK_MAX_ACTION_FLUENTS];
  int kind[RK_MAX_ACTION_FLUENTS];       // 0 real, 1 bool (sigmoid head), 2 int / enum (real head)
};

__device__ __forceinline__ float rk_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float rk_softplus(float x) { return log1pf(expf(-fabsf(x))) + fmaxf(x, 0.f); }
// _jax_bound_action (planner.py:620-628): sigmoid wrap for two-sided boxes, softplus wraps for
// one-sided ones, identity when unbounded; *d receives the derivative
__device__ __forceinline__ float rk_bound_action(float x, float lo, float hi, float* d) {
  const bool fl = isfinite(lo), fh = isfinite(hi);
  if (fl && fh) { const float s = rk_sigmoid(x); *d = (hi - lo) * s * (1.f - s); return lo + (hi - lo) * s; }
  if (fl) { *d = rk_sigmoid(x); return lo + rk_softplus(x); }
  if (fh) { *d = rk_sigmoid(-x); return hi - rk_softplus(-x); }
  *d = 1.f;
  return x;
}
// x * safe_log(x) of compiler.py:65-75, with 0 * log 0 taken as 0 (the reference yields NaN there)
__device__ __forceinline__ float rk_xlogx(float x) { return x > 0.f ? x * logf(x) : 0.f; }

// Boolean fluents (kind 1, planner.py:1371-1391): p = sigmoid(w * (logit + train * logistic noise)),
// entropy -= p0 log p0 + (1 - p0) log(1 - p0) with p0 = sigmoid(w * logit); the exact test program
// receives (p > 0.5) as an int32 word.  Int / enum fluents (kind 2) are real heads whose exact-
// program input is round(clip(.)) as int32 (planner.py:1491-1493).
__global__ void rk_drp_actions_fwd_kernel(int B, int A, int stochastic, RkActSegs S,
                                          const float* __restrict__ heads,
                                          const float* __restrict__ noise, float train,
                                          const float* __restrict__ lo,
                                          const float* __restrict__ hi, float ls_lo, float ls_hi,
                                          float bool_weight, int typed,
                                          float* __restrict__ actions,
                                          float* __restrict__ entropy) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int ld = stochastic ? 2 * A : A;
  float ent = 0.f;
  for (int f = 0; f < S.n; ++f) {
    if (S.kind[f] >= 3) continue;       // pooled boolean fluents: rk_drp_topk_forward
    for (int e = 0; e < S.len[f]; ++e) {
      const int a = S.off[f] + e;
      const size_t w = (size_t)S.off[f] * B + (size_t)b * S.len[f] + e;
      float act = heads[(size_t)b * ld + a];
      const float z = (stochastic && noise != nullptr) ? noise[w] : 0.f;
      if (S.kind[f] == 1) {
        if (stochastic) {
          const float p0 = rk_sigmoid(bool_weight * act);
          ent -= rk_xlogx(p0) + rk_xlogx(1.f - p0);
          act = act + train * z;
        }
        act = rk_sigmoid(bool_weight * act);
        if (typed & 1) ((int32_t*)actions)[w] = act > 0.5f ? 1 : 0;
        else actions[w] = act;
        continue;
      }
      if (stochastic) {
        const float ls = fminf(fmaxf(heads[(size_t)b * ld + A + a], ls_lo), ls_hi);
        ent += ls;
        act = act + train * expf(ls) * z;
      }
      if (lo != nullptr && (typed & 2)) {           // wrap_non_bool (planner.py:1411-1415)
        float d;
        act = rk_bound_action(act, lo[a], hi[a], &d);
      }
      if (lo != nullptr && (!(typed & 2) || (typed & 1)))
        act = fminf(fmaxf(act, lo[a]), hi[a]);      // planner.py:1472-1475, 1493-1496
      if ((typed & 1) && S.kind[f] == 2) ((int32_t*)actions)[w] = (int32_t)rintf(act);
      else actions[w] = act;
    }
  }
  if (entropy != nullptr) entropy[b] = ent;
}

__global__ void rk_drp_actions_bwd_kernel(int B, int A, int stochastic, RkActSegs S,
                                          const float* __restrict__ heads,
                                          const float* __restrict__ noise, float train,
                                          const float* __restrict__ lo,
                                          const float* __restrict__ hi, float ls_lo, float ls_hi,
                                          float bool_weight, const float* __restrict__ ent_coeff,
                                          float inv_batch, int wrap,
                                          const float* __restrict__ gact,
                                          float* __restrict__ gheads) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int ld = stochastic ? 2 * A : A;
  const bool sigma_entropy_grad = (wrap & 2) != 0;      // planner.py:1414-1417
  wrap &= 1;
  for (int f = 0; f < S.n; ++f) {
    if (S.kind[f] >= 3) continue;       // pooled boolean fluents: rk_drp_topk_backward
    for (int e = 0; e < S.len[f]; ++e) {
      const int a = S.off[f] + e;
      const size_t w = (size_t)S.off[f] * B + (size_t)b * S.len[f] + e;
      const float mu = heads[(size_t)b * ld + a];
      const float z = (stochastic && noise != nullptr) ? noise[w] : 0.f;
      if (S.kind[f] == 1) {
        const float s = rk_sigmoid(bool_weight * (mu + (stochastic ? train * z : 0.f)));
        float g = gact[w] * bool_weight * s * (1.f - s);
        if (stochastic && ent_coeff != nullptr) {
          // loss term -coeff * mean_b sum_t H: dH/dlogit through safe_log's tangent 1 / (x + eps)
          const float p0 = rk_sigmoid(bool_weight * mu), q0 = 1.f - p0, eps = 1e-12f;
          const float df = (p0 > 0.f ? logf(p0) : 0.f) + p0 / (p0 + eps)
                         - (q0 > 0.f ? logf(q0) : 0.f) - q0 / (q0 + eps);
          g += ent_coeff[0] * inv_batch * df * bool_weight * p0 * q0;
        }
        gheads[(size_t)b * ld + a] = g;
        if (stochastic) gheads[(size_t)b * ld + A + a] = 0.f;
        continue;
      }
      float pre = mu, ex = 0.f, raw = 0.f;
      if (stochastic) {
        raw = heads[(size_t)b * ld + A + a];
        const float ls = fminf(fmaxf(raw, ls_lo), ls_hi);
        ex = expf(ls);
        pre = mu + train * ex * z;
      }
      float g = gact[w];
      if (lo != nullptr && wrap) {                  // through the box wrap instead of the clip
        float d;
        rk_bound_action(pre, lo[a], hi[a], &d);
        g *= d;
      } else if (lo != nullptr && (pre < lo[a] || pre > hi[a])) {
        g = 0.f;    // jnp.clip: gradient 1 inside [lo, hi] (bounds included), 0 outside
      }
      gheads[(size_t)b * ld + a] = g;
      if (stochastic) {
        const bool inside = raw >= ls_lo && raw <= ls_hi;
        float gl = inside ? g * train * ex * z : 0.f;
        // entropy term -coeff * mean_b sum_t sum_a clip(log_sigma) with its gradient kept
        if (sigma_entropy_grad && inside && ent_coeff != nullptr) gl -= ent_coeff[0] * inv_batch;
        gheads[(size_t)b * ld + A + a] = gl;
      }
    }
  }
}

// ---- pooled boolean heads under a concurrency constraint (planner.py:1027-1080, 1435-1447) ------
// The boolean action fluents (kind 3; kind 4 = default value true, output flipped to 1 - a,
// planner.py:1456-1463) share ONE 'dense_topk' head: their head columns, in fluent order, are the
// n <= RK_TOPK_MAX_N pooled logits.  With allowed = max-nondef-actions:
//   allowed == 1        : a = softmax(logits + train * gumbel)                 (no weight, P:1040-1045)
//   allowed > 1, train  : khot_0 = 0, l_0 = logits + gumbel;  l_i = l_{i-1} + safe_log(1 - khot_{i-1}),
//                         khot_i = khot_{i-1} + softmax(w l_i), i = 1..allowed (P:1048-1063)
//   allowed > 1, test   : allowed times arg-max (first index on ties) -> exact k-hot (P:1066-1078)
// entropy (stochastic): -sum p log p with p = softmax(w logits) (P:1441-1444), ADDED to entropy[b].
// One thread per rollout; the per-iteration softmax outputs stay in local memory for the adjoint.
#define RK_TOPK_MAX_N 64
#define RK_TOPK_MAX_K 8

struct RkTopkPool {
  int n;                       // pooled logits
  int col[RK_TOPK_MAX_N];      // head column of pooled element j
  int seg[RK_TOPK_MAX_N];      // its fluent
  int elt[RK_TOPK_MAX_N];      // its element inside the fluent
};

__device__ __forceinline__ void rk_softmax_n(int n, float w, const float* l, float* p) {
  float m = -INFINITY;
  for (int j = 0; j < n; ++j) m = fmaxf(m, w * l[j]);
  float z = 0.f;
  for (int j = 0; j < n; ++j) { p[j] = expf(w * l[j] - m); z += p[j]; }
  const float r = 1.f / z;
  for (int j = 0; j < n; ++j) p[j] *= r;
}

// Fills P[i][0..n) = softmax(w l_i), i = 0..K-1, and returns the final khot in kh.
__device__ __forceinline__ void rk_soft_topk(int n, int K, float w, float* l, float* kh, float* P) {
  for (int j = 0; j < n; ++j) kh[j] = 0.f;
  for (int i = 0; i < K; ++i) {
    for (int j = 0; j < n; ++j) {
      const float r = 1.f - kh[j];
      l[j] += r > 0.f ? logf(r) : -INFINITY;        // safe_log (compiler.py:65-67)
    }
    float* p = P + i * RK_TOPK_MAX_N;
    rk_softmax_n(n, w, l, p);
    for (int j = 0; j < n; ++j) kh[j] += p[j];
  }
}

__global__ void rk_drp_topk_fwd_kernel(int B, int ld, int stochastic, RkActSegs S, RkTopkPool Q,
                                       const float* __restrict__ heads,
                                       const float* __restrict__ noise, int train, int allowed,
                                       float weight, int typed, float* __restrict__ actions,
                                       float* __restrict__ entropy) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float l[RK_TOPK_MAX_N], a[RK_TOPK_MAX_N], P[RK_TOPK_MAX_K * RK_TOPK_MAX_N];
  const int n = Q.n;
  for (int j = 0; j < n; ++j) l[j] = heads[(size_t)b * ld + Q.col[j]];
  if (stochastic && entropy != nullptr) {
    rk_softmax_n(n, weight, l, a);
    float m = -INFINITY, z = 0.f, h = 0.f;
    for (int j = 0; j < n; ++j) m = fmaxf(m, weight * l[j]);
    for (int j = 0; j < n; ++j) z += expf(weight * l[j] - m);
    const float lz = logf(z);
    for (int j = 0; j < n; ++j) h -= a[j] * (weight * l[j] - m - lz);     // p * log_softmax
    entropy[b] += h;
  }
  if (allowed > 1 && !train) {            // exact k-hot
    for (int j = 0; j < n; ++j) a[j] = 0.f;
    for (int i = 0; i < allowed; ++i) {
      int best = 0;
      for (int j = 1; j < n; ++j) if (l[j] > l[best]) best = j;
      l[best] = -INFINITY;
      a[best] = 1.f;
    }
  } else {
    if (stochastic && noise != nullptr && (train || allowed > 1)) {
      for (int j = 0; j < n; ++j) {
        const int f = Q.seg[j];
        l[j] += noise[(size_t)S.off[f] * B + (size_t)b * S.len[f] + Q.elt[j]];
      }
    }
    if (allowed == 1) rk_softmax_n(n, 1.f, l, a);
    else rk_soft_topk(n, allowed, weight, l, a, P);
  }
  for (int j = 0; j < n; ++j) {
    const int f = Q.seg[j];
    const size_t w = (size_t)S.off[f] * B + (size_t)b * S.len[f] + Q.elt[j];
    const float v = S.kind[f] == 4 ? 1.f - a[j] : a[j];
    if (typed & 1) ((int32_t*)actions)[w] = v > 0.5f ? 1 : 0;
    else actions[w] = v;
  }
}

__global__ void rk_drp_topk_bwd_kernel(int B, int A, int ld, int stochastic, RkActSegs S,
                                       RkTopkPool Q, const float* __restrict__ heads,
                                       const float* __restrict__ noise, int allowed, float weight,
                                       const float* __restrict__ ent_coeff, float inv_batch,
                                       const float* __restrict__ gact, float* __restrict__ gheads) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float l[RK_TOPK_MAX_N], kh[RK_TOPK_MAX_N], gk[RK_TOPK_MAX_N], gl[RK_TOPK_MAX_N];
  float P[RK_TOPK_MAX_K * RK_TOPK_MAX_N];
  const int n = Q.n;
  for (int j = 0; j < n; ++j) {
    const int f = Q.seg[j];
    const size_t w = (size_t)S.off[f] * B + (size_t)b * S.len[f] + Q.elt[j];
    l[j] = heads[(size_t)b * ld + Q.col[j]];
    if (stochastic && noise != nullptr) l[j] += noise[w];
    gk[j] = S.kind[f] == 4 ? -gact[w] : gact[w];
    gl[j] = 0.f;
  }
  if (allowed == 1) {
    rk_softmax_n(n, 1.f, l, kh);
    float d = 0.f;
    for (int j = 0; j < n; ++j) d += kh[j] * gk[j];
    for (int j = 0; j < n; ++j) gl[j] = kh[j] * (gk[j] - d);
  } else {
    rk_soft_topk(n, allowed, weight, l, kh, P);
    for (int i = allowed - 1; i >= 0; --i) {
      const float* p = P + i * RK_TOPK_MAX_N;
      float d = 0.f;
      for (int j = 0; j < n; ++j) d += p[j] * gk[j];
      for (int j = 0; j < n; ++j) gl[j] += weight * p[j] * (gk[j] - d);
      // khot_{i-1} in the forward order of additions; l_i = l_{i-1} + safe_log(1 - khot_{i-1})
      for (int j = 0; j < n; ++j) {
        float k0 = 0.f;
        for (int q = 0; q < i; ++q) k0 += P[q * RK_TOPK_MAX_N + j];
        gk[j] -= gl[j] / ((1.f - k0) + 1e-12f);       // safe_log tangent (compiler.py:70-75)
      }
    }
  }
  if (stochastic && ent_coeff != nullptr) {     // loss term -coeff * mean_b sum_t H(softmax(w logits))
    float lg[RK_TOPK_MAX_N];
    for (int j = 0; j < n; ++j) lg[j] = heads[(size_t)b * ld + Q.col[j]];
    rk_softmax_n(n, weight, lg, kh);
    float m = -INFINITY, z = 0.f, h = 0.f;
    for (int j = 0; j < n; ++j) m = fmaxf(m, weight * lg[j]);
    for (int j = 0; j < n; ++j) z += expf(weight * lg[j] - m);
    const float lz = logf(z);
    for (int j = 0; j < n; ++j) h -= kh[j] * (weight * lg[j] - m - lz);
    // dH/dlogit_j = -w p_j (log p_j + H)
    const float c = ent_coeff[0] * inv_batch;
    for (int j = 0; j < n; ++j)
      gl[j] += c * weight * kh[j] * ((weight * lg[j] - m - lz) + h);
  }
  for (int j = 0; j < n; ++j) {
    gheads[(size_t)b * ld + Q.col[j]] = gl[j];
    if (stochastic) gheads[(size_t)b * ld + A + Q.col[j]] = 0.f;
  }
}


static int make_segs(int A, int n_seg, const int32_t* seg_off, const int32_t* seg_len,
                     RkActSegs* S, const int32_t* seg_kind = nullptr) {
  if (n_seg <= 0 || n_seg > RK_MAX_ACTION_FLUENTS || seg_off == nullptr || seg_len == nullptr)
    return RK_E_BADARG;
  S->n = n_seg;
  int total = 0;
  for (int f = 0; f < n_seg; ++f) {
    if (seg_len[f] <= 0 || seg_off[f] != total) return RK_E_BADARG;
    S->off[f] = seg_off[f];
    S->len[f] = seg_len[f];
    S->kind[f] = seg_kind != nullptr ? seg_kind[f] : 0;
    total += seg_len[f];
  }
  return total == A ? 0 : RK_E_BADARG;
}

extern "C" int rk_drp_actions_forward(void* stream, int B, int A, int stochastic, int n_seg,
                                      const int32_t* seg_off, const int32_t* seg_len,
                                      const int32_t* seg_kind, const float* heads,
                                      const float* noise, float train, const float* lo,
                                      const float* hi, float ls_lo, float ls_hi,
                                      float bool_weight, int typed, float* actions,
                                      float* entropy) {
  if (B <= 0 || A <= 0 || heads == nullptr || actions == nullptr) return RK_E_BADARG;
  if ((lo == nullptr) != (hi == nullptr)) return RK_E_BADARG;
  RkActSegs S;
  int rc = make_segs(A, n_seg, seg_off, seg_len, &S, seg_kind);
  if (rc != 0) return rc;
  rk_drp_actions_fwd_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      B, A, stochastic, S, heads, noise, train, lo, hi, ls_lo, ls_hi, bool_weight, typed, actions,
      entropy);
  return (int)cudaGetLastError();
}

extern "C" int rk_drp_actions_backward(void* stream, int B, int A, int stochastic, int n_seg,
                                       const int32_t* seg_off, const int32_t* seg_len,
                                       const int32_t* seg_kind, const float* heads,
                                       const float* noise, float train, const float* lo,
                                       const float* hi, float ls_lo, float ls_hi,
                                       float bool_weight, const float* ent_coeff, float inv_batch,
                                       int wrap, const float* gact, float* gheads) {
  if (B <= 0 || A <= 0 || heads == nullptr || gact == nullptr || gheads == nullptr)
    return RK_E_BADARG;
  if ((lo == nullptr) != (hi == nullptr)) return RK_E_BADARG;
  RkActSegs S;
  int rc = make_segs(A, n_seg, seg_off, seg_len, &S, seg_kind);
  if (rc != 0) return rc;
  rk_drp_actions_bwd_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      B, A, stochastic, S, heads, noise, train, lo, hi, ls_lo, ls_hi, bool_weight, ent_coeff,
      inv_batch, wrap, gact, gheads);
  return (int)cudaGetLastError();
}

static int make_pool(const RkActSegs& S, RkTopkPool* Q) {
  Q->n = 0;
  for (int f = 0; f < S.n; ++f) {
    if (S.kind[f] < 3) continue;
    if (S.kind[f] > 4) return RK_E_BADARG;
    for (int e = 0; e < S.len[f]; ++e) {
      if (Q->n >= RK_TOPK_MAX_N) return RK_E_BADARG;
      Q->col[Q->n] = S.off[f] + e;
      Q->seg[Q->n] = f;
      Q->elt[Q->n] = e;
      ++Q->n;
    }
  }
  return Q->n > 0 ? 0 : RK_E_BADARG;
}

extern "C" int rk_drp_topk_limits(int* max_pooled, int* max_allowed) {
  if (max_pooled != nullptr) *max_pooled = RK_TOPK_MAX_N;
  if (max_allowed != nullptr) *max_allowed = RK_TOPK_MAX_K;
  return 0;
}

extern "C" int rk_drp_topk_forward(void* stream, int B, int A, int stochastic, int n_seg,
                                   const int32_t* seg_off, const int32_t* seg_len,
                                   const int32_t* seg_kind, const float* heads, const float* noise,
                                   int train, int allowed, float weight, int typed, float* actions,
                                   float* entropy) {
  if (B <= 0 || A <= 0 || heads == nullptr || actions == nullptr || seg_kind == nullptr ||
      allowed < 1 || allowed > RK_TOPK_MAX_K)
    return RK_E_BADARG;
  RkActSegs S;
  int rc = make_segs(A, n_seg, seg_off, seg_len, &S, seg_kind);
  if (rc != 0) return rc;
  RkTopkPool Q;
  rc = make_pool(S, &Q);
  if (rc != 0) return rc;
  rk_drp_topk_fwd_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      B, stochastic ? 2 * A : A, stochastic, S, Q, heads, noise, train, allowed, weight, typed,
      actions, entropy);
  return (int)cudaGetLastError();
}

extern "C" int rk_drp_topk_backward(void* stream, int B, int A, int stochastic, int n_seg,
                                    const int32_t* seg_off, const int32_t* seg_len,
                                    const int32_t* seg_kind, const float* heads,
                                    const float* noise, int allowed, float weight,
                                    const float* ent_coeff, float inv_batch, const float* gact,
                                    float* gheads) {
  if (B <= 0 || A <= 0 || heads == nullptr || gact == nullptr || gheads == nullptr ||
      seg_kind == nullptr || allowed < 1 || allowed > RK_TOPK_MAX_K)
    return RK_E_BADARG;
  RkActSegs S;
  int rc = make_segs(A, n_seg, seg_off, seg_len, &S, seg_kind);
  if (rc != 0) return rc;
  RkTopkPool Q;
  rc = make_pool(S, &Q);
  if (rc != 0) return rc;
  rk_drp_topk_bwd_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      B, A, stochastic ? 2 * A : A, stochastic, S, Q, heads, noise, allowed, weight, ent_coeff,
      inv_batch, gact, gheads);
  return (int)cudaGetLastError();
}

// ---- first dense layer straight from the fluent-major state (planner.py:1300-1313, 1358-1361)
// out[b][j] = tanh(sum_d x[b][d] * W0[d][j] + b0[j]), where the state vector x[b][:] is the
// concatenation of the fluent blocks ([B][n_f] at word offset off_f * B).  D is small (the state
// dimension), so the kernel is bound by writing out (B x H floats once).  Optionally also packs
// xp[b][0..D) = x[b][:], xp[b][D] = 1: the operand of the fused weight+bias gradient GEMM.
template <int ACT_CT>
__global__ void __launch_bounds__(256)
rk_drp_layer0_kernel(int B, int D, int H, RkActSegs S, const float* __restrict__ x,
                     const float* __restrict__ W0, const float* __restrict__ b0,
                     float* __restrict__ out, float* __restrict__ xp, int act_rt) {
  const int act = ACT_CT >= 0 ? ACT_CT : act_rt;
  extern __shared__ float sh0[];
  float* Ws = sh0;                       // [D][H]
  float* xs = sh0 + (size_t)D * H;       // [64][D]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < D * H; i += 256) Ws[i] = W0[i];
  const int r0 = blockIdx.x * 64;
  for (int i = tid; i < 64 * D; i += 256) {
    const int r = i / D, d = i % D, b = r0 + r;
    float v = 0.f;
    if (b < B) {
      int f = 0;
      while (f + 1 < S.n && d >= S.off[f + 1]) ++f;
      v = x[(size_t)S.off[f] * B + (size_t)b * S.len[f] + (d - S.off[f])];
      if (xp != nullptr) xp[(size_t)b * (D + 1) + d] = v;
    }
    xs[i] = v;
  }
  if (xp != nullptr)
    for (int r = tid; r < 64; r += 256)
      if (r0 + r < B) xp[(size_t)(r0 + r) * (D + 1) + D] = 1.f;
  __syncthreads();
  for (int r = warp; r < 64; r += 8) {
    const int b = r0 + r;
    if (b >= B) break;
    for (int j = lane; j < H; j += 32) {
      float acc = 0.f;
      for (int d = 0; d < D; ++d) acc = fmaf(xs[r * D + d], Ws[d * H + j], acc);
      out[(size_t)b * H + j] = rk_act_fn(act, acc + b0[j]);
    }
  }
}

// D <= 16, H <= 256: a thread owns one output column (its column of W0 in registers), the CTA walks
// over 64-row tiles of the state staged in shared memory as [64][16] (float4 broadcast reads).
template <int ACT_CT>
__global__ void __launch_bounds__(256)
rk_drp_layer0_cols_kernel(int B, int D, int H, RkActSegs S, const float* __restrict__ x,
                          const float* __restrict__ W0, const float* __restrict__ b0,
                          float* __restrict__ out, float* __restrict__ xp, int act_rt) {
  const int act = ACT_CT >= 0 ? ACT_CT : act_rt;
  __shared__ __align__(16) float xs[64 * 16];
  __shared__ int foff[16], flen[16], fd[16];
  const int tid = threadIdx.x;
  if (tid < 16) {
    int f = 0;
    while (f + 1 < S.n && tid >= S.off[f + 1]) ++f;
    foff[tid] = S.off[f];
    flen[tid] = S.len[f];
    fd[tid] = tid - S.off[f];
  }
  const bool live = tid < H;
  float w[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) w[d] = (live && d < D) ? W0[(size_t)d * H + tid] : 0.f;
  const float bias = live ? b0[tid] : 0.f;
  const int nq = (D + 3) >> 2;
  __syncthreads();
  for (int r0 = blockIdx.x * 64; r0 < B; r0 += gridDim.x * 64) {
    for (int i = tid; i < 64 * 16; i += 256) {
      const int r = i >> 4, d = i & 15, b = r0 + r;
      float v = 0.f;
      if (b < B && d < D) {
        v = x[(size_t)foff[d] * B + (size_t)b * flen[d] + fd[d]];
        if (xp != nullptr) xp[(size_t)b * (D + 1) + d] = v;
      }
      if (b < B && d == D && xp != nullptr) xp[(size_t)b * (D + 1) + D] = 1.f;
      xs[i] = v;
    }
    __syncthreads();
    const int rows = min(64, B - r0);
    if (live) {
#pragma unroll 4
      for (int r = 0; r < rows; ++r) {
        const float4* xr = reinterpret_cast<const float4*>(xs + r * 16);
        float acc = bias;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          if (v < nq) {
            const float4 q = xr[v];
            acc = fmaf(q.x, w[4 * v], acc);
            acc = fmaf(q.y, w[4 * v + 1], acc);
            acc = fmaf(q.z, w[4 * v + 2], acc);
            acc = fmaf(q.w, w[4 * v + 3], acc);
          }
        }
        out[(size_t)(r0 + r) * H + tid] = rk_act_fn(act, acc);
      }
    }
    __syncthreads();
  }
}

extern "C" int rk_drp_layer0(void* stream, int B, int D, int H, int n_seg,
                             const int32_t* seg_off, const int32_t* seg_len, const float* x,
                             const float* W0, const float* b0, float* out, float* xp, int act) {
  if (B <= 0 || D <= 0 || H <= 0 || x == nullptr || W0 == nullptr || b0 == nullptr ||
      out == nullptr || !rk_act_ok(act))
    return RK_E_BADARG;
  const int act_id = act;
  RkActSegs S;
  int rc = make_segs(D, n_seg, seg_off, seg_len, &S);
  if (rc != 0) return rc;
  if (D < 16 && H <= 256) {       // (D < 16: slot D of a tile row carries the packed 1)
    int grid = (B + 63) / 64;
    if (grid > 148 * 4) grid = 148 * 4;
    if (act_id == RK_ACT_TANH)
      rk_drp_layer0_cols_kernel<RK_ACT_TANH><<<grid, 256, 0, (cudaStream_t)stream>>>(
          B, D, H, S, x, W0, b0, out, xp, act_id);
    else
      rk_drp_layer0_cols_kernel<-1><<<grid, 256, 0, (cudaStream_t)stream>>>(
          B, D, H, S, x, W0, b0, out, xp, act_id);
    return (int)cudaGetLastError();
  }
  const size_t smem = ((size_t)D * H + 64 * (size_t)D) * sizeof(float);
  if (smem > 200 * 1024) return RK_E_SMEM;
  if (smem > 48 * 1024) {
    cudaError_t ce = cudaFuncSetAttribute(rk_drp_layer0_kernel<RK_ACT_TANH>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaFuncSetAttribute(rk_drp_layer0_kernel<-1>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return (int)ce;
  }
  if (act_id == RK_ACT_TANH)
    rk_drp_layer0_kernel<RK_ACT_TANH><<<(B + 63) / 64, 256, smem, (cudaStream_t)stream>>>(
        B, D, H, S, x, W0, b0, out, xp, act_id);
  else
    rk_drp_layer0_kernel<-1><<<(B + 63) / 64, 256, smem, (cudaStream_t)stream>>>(
        B, D, H, S, x, W0, b0, out, xp, act_id);
  return (int)cudaGetLastError();
}

// ---- column sums over the batch (bias gradients): y[n] (+)= sum_b G[b][n] --------------------
// stage 1: chunk c sums its rows for 32 columns per block into ws[c][n]; stage 2 (the split-K
// reduce kernel) adds the chunks in a fixed order.  G is read exactly once.
__global__ void __launch_bounds__(256)
rk_colsum_kernel(int B, int N, const float* __restrict__ G, int ldg, int rows_per_chunk,
                 float* __restrict__ ws) {
  __shared__ float part[8][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 32 + lane;
  const int b0 = blockIdx.y * rows_per_chunk, b1 = min(B, b0 + rows_per_chunk);
  float acc = 0.f;
  if (n < N)
    for (int b = b0 + warp; b < b1; b += 8) acc += G[(size_t)b * ldg + n];
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w][lane];
    ws[(size_t)blockIdx.y * N + n] = s;
  }
}

extern "C" int rk_colsum(void* stream, int B, int N, const float* G, int ldg, float* y,
                         int accumulate, float* workspace, int chunks) {
  if (B <= 0 || N <= 0 || G == nullptr || y == nullptr || workspace == nullptr || chunks < 1)
    return RK_E_BADARG;
  if (chunks > B) chunks = B;
  const int rpc = (B + chunks - 1) / chunks;
  chunks = (B + rpc - 1) / rpc;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((N + 31) / 32, chunks);
  rk_colsum_kernel<<<grid, 256, 0, st>>>(B, N, G, ldg, rpc, workspace);
  rk_sgemm_reduce_kernel<<<(N + 255) / 256, 256, 0, st>>>(1, N, chunks, workspace, y, N, nullptr,
                                                           nullptr, 0, 0, accumulate ? 1 : 0);
  return (int)cudaGetLastError();
}

// ---- FiLM time conditioning of a hidden layer (planner.py:1083-1092, 1362-1363) ------------------
// out[b][c] = gamma[c] * z[b][c] + beta[c]: gamma / beta are the two Dense(time embedding) vectors
// of this decision epoch, the same for every rollout.  The adjoint takes the cotangent g of `out`
// and produces, in one pass over [B][H] (g and z read once): the pre-activation cotangent
// gz = g * gamma * (1 - z^2) (z = tanh output; gz may alias g) and the column sums
// dgamma[c] += sum_b g z, dbeta[c] += sum_b g through per-chunk partials ws[chunk][2 H] added in
// chunk order (deterministic).
__global__ void rk_add_rows_kernel(int R, int n, const float* __restrict__ ws,
                                   float* __restrict__ y, int accumulate);

__global__ void rk_film_fwd_kernel(size_t n, int H, const float* __restrict__ z,
                                   const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % (size_t)H);
    out[i] = fmaf(gamma[c], z[i], beta[c]);
  }
}

__global__ void __launch_bounds__(256)
rk_film_bwd_kernel(int B, int H, int rows_per_chunk, const float* g,
                   const float* __restrict__ z, const float* __restrict__ gamma, float* gz,
                   float* __restrict__ ws, int act) {
  __shared__ float part[2][8][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lane;
  const int b0 = blockIdx.y * rows_per_chunk, b1 = min(B, b0 + rows_per_chunk);
  float ag = 0.f, ab = 0.f;
  if (c < H) {
    const float gm = gamma[c];
    for (int b = b0 + warp; b < b1; b += 8) {
      const size_t i = (size_t)b * H + c;
      const float gv = g[i], zv = z[i];
      ag += gv * zv;
      ab += gv;
      gz[i] = gv * gm * rk_act_grad(act, zv);
    }
  }
  part[0][warp][lane] = ag;
  part[1][warp][lane] = ab;
  __syncthreads();
  if (warp < 2 && c < H) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[warp][w][lane];
    ws[(size_t)blockIdx.y * (2 * (size_t)H) + (size_t)warp * H + c] = s;
  }
}

extern "C" int rk_film_chunks(int B) {
  int chunks = B / 64;
  if (chunks < 1) chunks = 1;
  if (chunks > 74) chunks = 74;       // 8 column blocks x 74 chunks = 4 x 148 CTAs at H = 256
  const int rpc = (B + chunks - 1) / chunks;
  return (B + rpc - 1) / rpc;
}

extern "C" int rk_film_forward(void* stream, int B, int H, const float* z, const float* gamma,
                               const float* beta, float* out) {
  if (B <= 0 || H <= 0 || z == nullptr || gamma == nullptr || beta == nullptr || out == nullptr)
    return RK_E_BADARG;
  const size_t n = (size_t)B * H;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  rk_film_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, H, z, gamma, beta, out);
  return (int)cudaGetLastError();
}

extern "C" int rk_film_backward(void* stream, int B, int H, const float* g, const float* z,
                                const float* gamma, float* gz, float* grad_gamma_beta,
                                float* workspace, int act) {
  if (B <= 0 || H <= 0 || g == nullptr || z == nullptr || gamma == nullptr || gz == nullptr ||
      grad_gamma_beta == nullptr || workspace == nullptr || !rk_act_ok(act))
    return RK_E_BADARG;
  const int act_id = act;
  const int chunks = rk_film_chunks(B);
  const int rpc = (B + chunks - 1) / chunks;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((H + 31) / 32, chunks);
  rk_film_bwd_kernel<<<grid, 256, 0, st>>>(B, H, rpc, g, z, gamma, gz, workspace, act_id);
  rk_add_rows_kernel<<<(2 * H + 127) / 128, 128, 0, st>>>(chunks, 2 * H, workspace,
                                                          grad_gamma_beta, 1);
  return (int)cudaGetLastError();
}

// ---- fused backward of the action heads -----------------------------------------------------
// Given the head cotangent G[B][NH] and the last hidden activation Hl[B][hl] (both read ONCE):
//   gz[b][c]  = (sum_j G[b][j] Wh[c][j]) * (1 - Hl[b][c]^2)     cotangent of the last pre-activation
//   dWh[c][j] += sum_b Hl[b][c] G[b][j]      dbh[j] += sum_b G[b][j]      dbl[c] += sum_b gz[b][c]
// Grid = (32-column blocks, row chunks).  A lane owns one column (its row of Wh and its
// accumulators stay in registers), the 8 warps of a CTA stride over the chunk's rows two at a
// time, and the CTA leaves its partial [dbl | dWh | dbh] in ws[chunk][...]; the caller adds the
// chunk partials in order (deterministic).  Replaces four streaming passes over [B][hl] (two
// skinny GEMMs and two column sums) by one read and one write.
#define RK_FB_WARPS 8
__device__ __forceinline__ float cta_sum_in_warp_order(float v, float (*part)[33], int warp,
                                                       int lane) {
  __syncthreads();
  part[warp][lane] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < RK_FB_WARPS; ++q) s += part[q][lane];
  return s;
}

template <int ACT_CT>
__global__ void __launch_bounds__(256)
rk_drp_head_bwd_kernel(int B, int hl, int NH, const float* __restrict__ Hl,
                       const float* __restrict__ G, const float* __restrict__ Wh,
                       float* __restrict__ gz, float* __restrict__ ws, int rows_per_chunk,
                       int act_rt) {
  const int act = ACT_CT >= 0 ? ACT_CT : act_rt;
  __shared__ float part[RK_FB_WARPS][33];
  extern __shared__ __align__(16) float Gs[];          // [rows_per_chunk][16], zero padded
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lane;
  const bool live = c < hl;
  const int b0 = blockIdx.y * rows_per_chunk, b1 = min(B, b0 + rows_per_chunk);
  for (int i = threadIdx.x; i < (b1 - b0) * 16; i += 256) {
    const int r = i >> 4, j = i & 15;
    Gs[i] = (j < NH) ? G[(size_t)(b0 + r) * NH + j] : 0.f;
  }
  float w[16], dw[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    w[j] = (live && j < NH) ? Wh[(size_t)c * NH + j] : 0.f;
    dw[j] = 0.f;
  }
  __syncthreads();
  float dbl = 0.f, dbh = 0.f;
  const int nq = (NH + 3) >> 2;
  for (int b = b0 + warp; b < b1; b += 2 * RK_FB_WARPS) {
    const int bb = b + RK_FB_WARPS;
    const bool two = bb < b1;
    const float h0 = live ? Hl[(size_t)b * hl + c] : 0.f;
    const float h1 = (live && two) ? Hl[(size_t)bb * hl + c] : 0.f;
    const float4* ga = reinterpret_cast<const float4*>(Gs + (size_t)(b - b0) * 16);
    const float4* gb = reinterpret_cast<const float4*>(Gs + (size_t)((two ? bb : b) - b0) * 16);
    if (blockIdx.x == 0 && lane < NH) {
      dbh += Gs[(size_t)(b - b0) * 16 + lane];
      if (two) dbh += Gs[(size_t)(bb - b0) * 16 + lane];
    }
    float z0 = 0.f, z1 = 0.f;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (v < nq) {
        const float4 p = ga[v];
        float4 q = gb[v];
        if (!two) q = make_float4(0.f, 0.f, 0.f, 0.f);
        z0 = fmaf(p.x, w[4 * v], z0);     z1 = fmaf(q.x, w[4 * v], z1);
        z0 = fmaf(p.y, w[4 * v + 1], z0); z1 = fmaf(q.y, w[4 * v + 1], z1);
        z0 = fmaf(p.z, w[4 * v + 2], z0); z1 = fmaf(q.z, w[4 * v + 2], z1);
        z0 = fmaf(p.w, w[4 * v + 3], z0); z1 = fmaf(q.w, w[4 * v + 3], z1);
        dw[4 * v] = fmaf(h1, q.x, fmaf(h0, p.x, dw[4 * v]));
        dw[4 * v + 1] = fmaf(h1, q.y, fmaf(h0, p.y, dw[4 * v + 1]));
        dw[4 * v + 2] = fmaf(h1, q.z, fmaf(h0, p.z, dw[4 * v + 2]));
        dw[4 * v + 3] = fmaf(h1, q.w, fmaf(h0, p.w, dw[4 * v + 3]));
      }
    }
    const float v0 = z0 * rk_act_grad(act, h0), v1 = z1 * rk_act_grad(act, h1);
    if (live) {
      gz[(size_t)b * hl + c] = v0;
      if (two) gz[(size_t)bb * hl + c] = v1;
    }
    dbl += v0 + v1;
  }
  float* out = ws + (size_t)blockIdx.y * ((size_t)hl + (size_t)hl * NH + NH);
  float s = cta_sum_in_warp_order(dbl, part, warp, lane);
  if (warp == 0 && live) out[c] = s;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j < NH) {                                // NH is uniform: every thread reaches the barriers
      s = cta_sum_in_warp_order(dw[j], part, warp, lane);
      if (warp == 0 && live) out[hl + (size_t)c * NH + j] = s;
    }
  }
  if (blockIdx.x == 0) {
    s = cta_sum_in_warp_order(dbh, part, warp, lane);
    if (warp == 0 && lane < NH) out[(size_t)hl + (size_t)hl * NH + lane] = s;
  }
}

// adds rows [0, R) of ws[R][n] to y[n] in row order
__global__ void rk_add_rows_kernel(int R, int n, const float* __restrict__ ws,
                                   float* __restrict__ y, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = accumulate ? y[i] : 0.f;
  const float* p = ws + i;
  int r = 0;
  for (; r + 4 <= R; r += 4) {
    const float v0 = p[0], v1 = p[(size_t)n], v2 = p[2 * (size_t)n], v3 = p[3 * (size_t)n];
    acc = (((acc + v0) + v1) + v2) + v3;
    p += 4 * (size_t)n;
  }
  for (; r < R; ++r) { acc += *p; p += n; }
  y[i] = acc;
}

extern "C" int rk_drp_head_backward_ctas(int B) {     // row chunks of the fused backward kernels
  int chunks = (B + 255) / 256;                       // at most 256 rows per chunk (smem tile)
  return chunks < 1 ? 1 : chunks;
}

extern "C" int rk_drp_head_backward(void* stream, int B, int hl, int NH, const float* Hl,
                                    const float* G, const float* Wh, float* gz, float* grad_tail,
                                    float* workspace, int act) {
  if (B <= 0 || hl <= 0 || hl > 1024 || NH <= 0 || NH > 16 || Hl == nullptr || G == nullptr ||
      Wh == nullptr || gz == nullptr || grad_tail == nullptr || workspace == nullptr || !rk_act_ok(act))
    return RK_E_BADARG;
  const int act_id = act;
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks0 = rk_drp_head_backward_ctas(B);
  const int rpc = (B + chunks0 - 1) / chunks0;
  const int chunks = (B + rpc - 1) / rpc;
  dim3 grid((hl + 31) / 32, chunks);
  const size_t smem = (size_t)rpc * 16 * sizeof(float);     // <= 256 rows x 64 B
  if (act_id == RK_ACT_TANH)
    rk_drp_head_bwd_kernel<RK_ACT_TANH><<<grid, 256, smem, st>>>(B, hl, NH, Hl, G, Wh, gz, workspace,
                                                                 rpc, act_id);
  else
    rk_drp_head_bwd_kernel<-1><<<grid, 256, smem, st>>>(B, hl, NH, Hl, G, Wh, gz, workspace, rpc,
                                                        act_id);
  const int n = hl + hl * NH + NH;
  rk_add_rows_kernel<<<(n + 127) / 128, 128, 0, st>>>(chunks, n, workspace, grad_tail, 1);
  return (int)cudaGetLastError();
}

// ---- fused weight / bias gradient of the first layer -------------------------------------------
// Given gz0[B][H] (read once) and the fluent-major state x (D <= 16 input words per rollout):
//   dW0[d][c] += sum_b x[b][d] gz0[b][c]     db0[c] += sum_b gz0[b][c]
// Same decomposition as the head kernel: a lane owns a column, warps stride over the chunk's
// rows, per-chunk partials [dW0 | db0] go to ws[chunk][(D + 1) * H] and are added in order.
// (The input cotangent gs += gz0 W0^T is a row reduction: rk_rowdot_segs below.)
__global__ void __launch_bounds__(256)
rk_drp_layer0_bwd_kernel(int B, int D, int H, RkActSegs S, const float* __restrict__ x,
                         const float* __restrict__ gz0, float* __restrict__ ws,
                         int rows_per_chunk) {
  __shared__ float part[RK_FB_WARPS][33];
  __shared__ int foff[16], flen[16], fd[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 16) {
    const int d = threadIdx.x;
    int f = 0;
    while (f + 1 < S.n && d >= S.off[f + 1]) ++f;
    foff[d] = S.off[f];
    flen[d] = S.len[f];
    fd[d] = d - S.off[f];
  }
  __syncthreads();
  extern __shared__ __align__(16) float Xs[];          // [rows_per_chunk][16], zero padded
  const int c = blockIdx.x * 32 + lane;
  const bool live = c < H;
  const int b0 = blockIdx.y * rows_per_chunk, b1 = min(B, b0 + rows_per_chunk);
  for (int i = threadIdx.x; i < (b1 - b0) * 16; i += 256) {
    const int r = i >> 4, d = i & 15;
    Xs[i] = (d < D) ? x[(size_t)foff[d] * B + (size_t)(b0 + r) * flen[d] + fd[d]] : 0.f;
  }
  __syncthreads();
  float dw[16], db = 0.f;
#pragma unroll
  for (int d = 0; d < 16; ++d) dw[d] = 0.f;
  const int nq = (D + 3) >> 2;
  for (int b = b0 + warp; b < b1; b += 2 * RK_FB_WARPS) {
    const int bb = b + RK_FB_WARPS;
    const bool two = bb < b1;
    const float g0 = live ? gz0[(size_t)b * H + c] : 0.f;
    const float g1 = (live && two) ? gz0[(size_t)bb * H + c] : 0.f;
    db += g0 + g1;
    const float4* xa = reinterpret_cast<const float4*>(Xs + (size_t)(b - b0) * 16);
    const float4* xb = reinterpret_cast<const float4*>(Xs + (size_t)((two ? bb : b) - b0) * 16);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (v < nq) {
        const float4 p = xa[v], q = xb[v];       // g1 == 0 when there is no second row
        dw[4 * v] = fmaf(q.x, g1, fmaf(p.x, g0, dw[4 * v]));
        dw[4 * v + 1] = fmaf(q.y, g1, fmaf(p.y, g0, dw[4 * v + 1]));
        dw[4 * v + 2] = fmaf(q.z, g1, fmaf(p.z, g0, dw[4 * v + 2]));
        dw[4 * v + 3] = fmaf(q.w, g1, fmaf(p.w, g0, dw[4 * v + 3]));
      }
    }
  }
  float* out = ws + (size_t)blockIdx.y * ((size_t)(D + 1) * H);
#pragma unroll
  for (int d = 0; d < 16; ++d) {
    if (d < D) {
      const float s = cta_sum_in_warp_order(dw[d], part, warp, lane);
      if (warp == 0 && live) out[(size_t)d * H + c] = s;
    }
  }
  const float s = cta_sum_in_warp_order(db, part, warp, lane);
  if (warp == 0 && live) out[(size_t)D * H + c] = s;
}

// ---- input LayerNorm over the non-boolean state words (planner.py:1300-1319) -------------------
// y = (x - mean) * rsqrt(var + eps) * scale + bias, var = E[x^2] - E[x]^2 (flax nn.LayerNorm,
// use_fast_variance), taken over the words of one GROUP of fluents: seg_norm[f] = 0 passes fluent f
// through, g >= 1 puts it in group g (normalize: every non-bool fluent in group 1;
// normalize_per_layer: one group per fluent).  All buffers are fluent-major; scale / bias are
// indexed by the word's position among the normalised words in fluent order.  One thread per
// rollout (the state dimension is tens of words).
__device__ __forceinline__ int rk_ln_stats(int B, int b, const RkActSegs& S, int g,
                                           const float* __restrict__ x, float eps, float* mean,
                                           float* rstd) {
  float s = 0.f, ss = 0.f;
  int n = 0;
  for (int f = 0; f < S.n; ++f) {
    if (S.kind[f] != g) continue;
    const float* p = x + (size_t)S.off[f] * B + (size_t)b * S.len[f];
    for (int e = 0; e < S.len[f]; ++e) { const float v = p[e]; s += v; ss += v * v; }
    n += S.len[f];
  }
  const float m = s / (float)n;
  const float var = fmaxf(ss / (float)n - m * m, 0.f);
  *mean = m;
  *rstd = rsqrtf(var + eps);
  return n;
}

__global__ void rk_state_layernorm_fwd_kernel(int B, RkActSegs S, int G, float eps,
                                              const float* __restrict__ x,
                                              const float* __restrict__ scale,
                                              const float* __restrict__ bias,
                                              float* __restrict__ y) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int f = 0; f < S.n; ++f) {
    if (S.kind[f] != 0) continue;
    const size_t w = (size_t)S.off[f] * B + (size_t)b * S.len[f];
    for (int e = 0; e < S.len[f]; ++e) y[w + e] = x[w + e];
  }
  for (int g = 1; g <= G; ++g) {
    float mean, rstd;
    rk_ln_stats(B, b, S, g, x, eps, &mean, &rstd);
    int j = 0;
    for (int f = 0; f < S.n; ++f) {
      if (S.kind[f] == g) {
        const size_t w = (size_t)S.off[f] * B + (size_t)b * S.len[f];
        for (int e = 0; e < S.len[f]; ++e)
          y[w + e] = (x[w + e] - mean) * rstd * scale[j + e] + bias[j + e];
      }
      if (S.kind[f] != 0) j += S.len[f];
    }
  }
}

// gx += LayerNorm^T gy; per-CTA partials of [dscale | dbias] go to ws[cta][2 * Dn]
__global__ void __launch_bounds__(128)
rk_state_layernorm_bwd_kernel(int B, RkActSegs S, int G, int Dn, float eps,
                              const float* __restrict__ x, const float* __restrict__ scale,
                              const float* __restrict__ gy, float* __restrict__ gx,
                              float* __restrict__ ws) {
  __shared__ float red[2][4];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = b < B;
  float* out = ws + (size_t)blockIdx.x * (2 * (size_t)Dn);
  if (live) {
    for (int f = 0; f < S.n; ++f) {
      if (S.kind[f] != 0) continue;
      const size_t w = (size_t)S.off[f] * B + (size_t)b * S.len[f];
      for (int e = 0; e < S.len[f]; ++e) gx[w + e] += gy[w + e];
    }
  }
  for (int g = 1; g <= G; ++g) {
    float mean = 0.f, rstd = 0.f, m1 = 0.f, m2 = 0.f;
    if (live) {
      const int n = rk_ln_stats(B, b, S, g, x, eps, &mean, &rstd);
      int j = 0;
      for (int f = 0; f < S.n; ++f) {
        if (S.kind[f] == g) {
          const size_t w = (size_t)S.off[f] * B + (size_t)b * S.len[f];
          for (int e = 0; e < S.len[f]; ++e) {
            const float gt = gy[w + e] * scale[j + e], xh = (x[w + e] - mean) * rstd;
            m1 += gt;
            m2 += gt * xh;
          }
        }
        if (S.kind[f] != 0) j += S.len[f];
      }
      m1 /= (float)n;
      m2 /= (float)n;
    }
    int j = 0;
    for (int f = 0; f < S.n; ++f) {
      if (S.kind[f] == g) {
        for (int e = 0; e < S.len[f]; ++e) {
          float ds = 0.f, db = 0.f;
          if (live) {
            const size_t w = (size_t)S.off[f] * B + (size_t)b * S.len[f] + e;
            const float gg = gy[w], xh = (x[w] - mean) * rstd;
            gx[w] += rstd * (gg * scale[j + e] - m1 - xh * m2);
            ds = gg * xh;
            db = gg;
          }
          // block sums in lane / warp order (deterministic)
          for (int o = 16; o > 0; o >>= 1) {
            ds += __shfl_xor_sync(0xffffffffu, ds, o);
            db += __shfl_xor_sync(0xffffffffu, db, o);
          }
          __syncthreads();
          if ((threadIdx.x & 31) == 0) {
            red[0][threadIdx.x >> 5] = ds;
            red[1][threadIdx.x >> 5] = db;
          }
          __syncthreads();
          if (threadIdx.x == 0) {
            out[j + e] = ((red[0][0] + red[0][1]) + red[0][2]) + red[0][3];
            out[Dn + j + e] = ((red[1][0] + red[1][1]) + red[1][2]) + red[1][3];
          }
        }
      }
      if (S.kind[f] != 0) j += S.len[f];
    }
  }
}

extern "C" int rk_state_layernorm_ctas(int B) { return (B + 127) / 128; }

// groups must be numbered 1..G without holes; returns G (or a negative error) and the number of
// normalised words
static int rk_ln_groups(const RkActSegs& S, int* Dn) {
  int G = 0;
  *Dn = 0;
  for (int f = 0; f < S.n; ++f) {
    if (S.kind[f] < 0 || S.kind[f] > S.n) return RK_E_BADARG;
    if (S.kind[f] > G) G = S.kind[f];
    if (S.kind[f] != 0) *Dn += S.len[f];
  }
  for (int g = 1; g <= G; ++g) {
    bool any = false;
    for (int f = 0; f < S.n; ++f) any = any || S.kind[f] == g;
    if (!any) return RK_E_BADARG;
  }
  return G > 0 ? G : RK_E_BADARG;
}

extern "C" int rk_state_layernorm_forward(void* stream, int B, int D, int n_seg,
                                          const int32_t* seg_off, const int32_t* seg_len,
                                          const int32_t* seg_norm, float eps, const float* x,
                                          const float* scale, const float* bias, float* y) {
  if (B <= 0 || x == nullptr || y == nullptr || scale == nullptr || bias == nullptr ||
      seg_norm == nullptr)
    return RK_E_BADARG;
  RkActSegs S;
  int rc = make_segs(D, n_seg, seg_off, seg_len, &S, seg_norm);
  if (rc != 0) return rc;
  int Dn = 0;
  const int G = rk_ln_groups(S, &Dn);
  if (G < 0) return G;
  rk_state_layernorm_fwd_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      B, S, G, eps, x, scale, bias, y);
  return (int)cudaGetLastError();
}

extern "C" int rk_state_layernorm_backward(void* stream, int B, int D, int n_seg,
                                           const int32_t* seg_off, const int32_t* seg_len,
                                           const int32_t* seg_norm, float eps, const float* x,
                                           const float* scale, const float* gy, float* gx,
                                           float* grad_scale_bias, float* workspace) {
  if (B <= 0 || x == nullptr || gy == nullptr || gx == nullptr || scale == nullptr ||
      grad_scale_bias == nullptr || workspace == nullptr || seg_norm == nullptr)
    return RK_E_BADARG;
  RkActSegs S;
  int rc = make_segs(D, n_seg, seg_off, seg_len, &S, seg_norm);
  if (rc != 0) return rc;
  int Dn = 0;
  const int G = rk_ln_groups(S, &Dn);
  if (G < 0) return G;
  cudaStream_t st = (cudaStream_t)stream;
  const int ctas = rk_state_layernorm_ctas(B);
  rk_state_layernorm_bwd_kernel<<<ctas, 128, 0, st>>>(B, S, G, Dn, eps, x, scale, gy, gx,
                                                      workspace);
  rk_add_rows_kernel<<<(2 * Dn + 127) / 128, 128, 0, st>>>(ctas, 2 * Dn, workspace,
                                                           grad_scale_bias, 1);
  return (int)cudaGetLastError();
}

// ---- static min-max scaling of the policy input (StaticNormalizer, planner.py:284-338) ----------
// y = (x - lo[d]) / (hi[d] - lo[d]) per state word d (the caller passes lo = 0, hi = 1 for the words
// that are not scaled: exact pass-through); backward: gx += gy / (hi[d] - lo[d]).  Fluent-major
// buffers, one thread per (rollout, word) of a fluent block: fully coalesced.
__global__ void rk_state_scale_kernel(int B, RkActSegs S, const float* __restrict__ lo,
                                      const float* __restrict__ hi, const float* __restrict__ x,
                                      float* __restrict__ y, int backward) {
  const size_t total = (size_t)(S.off[S.n - 1] + S.len[S.n - 1]) * B;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    int f = 0;
    while (f + 1 < S.n && i >= (size_t)S.off[f + 1] * B) ++f;
    const int d = S.off[f] + (int)((i - (size_t)S.off[f] * B) % (size_t)S.len[f]);
    const float w = hi[d] - lo[d];
    if (backward) y[i] += x[i] / w;
    else y[i] = (x[i] - lo[d]) / w;
  }
}

static int rk_state_scale(void* stream, int B, int D, int n_seg, const int32_t* seg_off,
                          const int32_t* seg_len, const float* lo, const float* hi, const float* x,
                          float* y, int backward) {
  if (B <= 0 || lo == nullptr || hi == nullptr || x == nullptr || y == nullptr) return RK_E_BADARG;
  RkActSegs S;
  int rc = make_segs(D, n_seg, seg_off, seg_len, &S);
  if (rc != 0) return rc;
  size_t blocks = ((size_t)D * B + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  rk_state_scale_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(B, S, lo, hi, x, y,
                                                                            backward);
  return (int)cudaGetLastError();
}

extern "C" int rk_state_scale_forward(void* stream, int B, int D, int n_seg,
                                      const int32_t* seg_off, const int32_t* seg_len,
                                      const float* lo, const float* hi, const float* x, float* y) {
  return rk_state_scale(stream, B, D, n_seg, seg_off, seg_len, lo, hi, x, y, 0);
}

extern "C" int rk_state_scale_backward(void* stream, int B, int D, int n_seg,
                                       const int32_t* seg_off, const int32_t* seg_len,
                                       const float* lo, const float* hi, const float* gy,
                                       float* gx) {
  return rk_state_scale(stream, B, D, n_seg, seg_off, seg_len, lo, hi, gy, gx, 1);
}

// ---- skinny products with N <= 16 (action heads, state-cotangent blocks) ----------------------
// out[b][n] = (accumulate ? out[b][n] : 0) + sum_k A[b][k] * W[k][n] (+ bias[n]).
// One warp owns FOUR rows at a time: lanes stride over k (coalesced 128-byte reads of A, read
// exactly once), W sits in shared memory as [K][20] (float4 reads, conflict-free stride) and is
// shared by the four rows; the 16 partial sums of a row are reduced across the warp by a
// transposing butterfly (16 shuffles instead of 80).  With segments (S.n > 0) output n of row b
// goes to the fluent-major block of its fluent: out[off_f * B + b * len_f + (n - off_f)].
#define RK_RD_NP 20
__device__ __forceinline__ float warp_transpose_sum16(const float (&acc)[16], int lane) {
  float a8[8], a4[4], a2[2];
  const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0, b1 = (lane & 2) != 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = b4 ? acc[i] : acc[i + 8], keep = b4 ? acc[i + 8] : acc[i];
    a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b3 ? a8[i] : a8[i + 4], keep = b3 ? a8[i + 4] : a8[i];
    a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b2 ? a4[i] : a4[i + 2], keep = b2 ? a4[i + 2] : a4[i];
    a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = b1 ? a2[0] : a2[1], keep = b1 ? a2[1] : a2[0];
  float a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  return a1;        // total of output n = 8 b4 + 4 b3 + 2 b2 + b1
}

template <int RK_RD_ROWS, int MINB>
__global__ void __launch_bounds__(256, MINB)
rk_rowdot_kernel(int B, int N, int K, const float* __restrict__ A, int lda,
                 const float* __restrict__ W, int ldw, float* __restrict__ out, int ldo,
                 const float* __restrict__ bias, int accumulate, RkActSegs S) {
  extern __shared__ __align__(16) float Wsh[];
  __shared__ int ooff[16], olen[16];
  for (int i = threadIdx.x; i < K * RK_RD_NP; i += 256) {
    const int k = i / RK_RD_NP, n = i % RK_RD_NP;
    Wsh[i] = (n < N) ? W[(size_t)k * ldw + n] : 0.f;
  }
  if (threadIdx.x < 16) {
    const int n = threadIdx.x;
    int f = 0;
    while (f + 1 < S.n && n >= S.off[f + 1]) ++f;
    ooff[n] = S.n > 0 ? S.off[f] : 0;
    olen[n] = S.n > 0 ? S.len[f] : 0;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = (N + 3) >> 2;
  for (int b = (blockIdx.x * 8 + warp) * RK_RD_ROWS; b < B; b += gridDim.x * 8 * RK_RD_ROWS) {
    float acc[RK_RD_ROWS][16];
#pragma unroll
    for (int r = 0; r < RK_RD_ROWS; ++r)
#pragma unroll
      for (int n = 0; n < 16; ++n) acc[r][n] = 0.f;
    for (int k0 = lane; k0 < K; k0 += 64) {       // two k per trip: 2 x ROWS loads in flight
     float av2[2][RK_RD_ROWS];
#pragma unroll
     for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int r = 0; r < RK_RD_ROWS; ++r)
        av2[u][r] = (b + r < B && k0 + 32 * u < K) ? A[(size_t)(b + r) * lda + k0 + 32 * u] : 0.f;
#pragma unroll
     for (int u = 0; u < 2; ++u) {
      const int k = min(k0 + 32 * u, K - 1);       // out-of-range k: av == 0
      const float* av = av2[u];
      const float4* w4 = reinterpret_cast<const float4*>(Wsh + k * RK_RD_NP);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        if (v < nq) {
          const float4 w = w4[v];
#pragma unroll
          for (int r = 0; r < RK_RD_ROWS; ++r) {
            acc[r][4 * v + 0] = fmaf(av[r], w.x, acc[r][4 * v + 0]);
            acc[r][4 * v + 1] = fmaf(av[r], w.y, acc[r][4 * v + 1]);
            acc[r][4 * v + 2] = fmaf(av[r], w.z, acc[r][4 * v + 2]);
            acc[r][4 * v + 3] = fmaf(av[r], w.w, acc[r][4 * v + 3]);
          }
        }
      }
     }
    }
    const int n = lane >> 1;
#pragma unroll
    for (int r = 0; r < RK_RD_ROWS; ++r) {
      float v = warp_transpose_sum16(acc[r], lane);
      const int bb = b + r;
      if ((lane & 1) == 0 && n < N && bb < B) {
        if (bias != nullptr) v += bias[n];
        float* dst = (S.n > 0) ? out + (size_t)ooff[n] * B + (size_t)bb * olen[n] + (n - ooff[n])
                               : out + (size_t)bb * ldo + n;
        if (accumulate) v += *dst;
        *dst = v;
      }
    }
  }
}

static int launch_rowdot(cudaStream_t st, int accumulate, int B, int N, int K, const float* A,
                         int lda, const float* W, int ldw, float* out, int ldo, const float* bias,
                         const RkActSegs& S) {
  const size_t smem = (size_t)K * RK_RD_NP * sizeof(float);
  if (smem > 200 * 1024) return RK_E_SMEM;
  auto launch = [&](auto kern, int rows, int ctas_per_sm) -> int {
    if (smem > 48 * 1024) {
      cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem);
      if (ce != cudaSuccess) return (int)ce;
    }
    int grid = (B + 8 * rows - 1) / (8 * rows);
    // persistent: W (K x 20 floats) is staged once per CTA
    if (grid > 148 * ctas_per_sm) grid = 148 * ctas_per_sm;
    kern<<<grid, 256, smem, st>>>(B, N, K, A, lda, W, ldw, out, ldo, bias, accumulate ? 1 : 0, S);
    return (int)cudaGetLastError();
  };
  return launch(rk_rowdot_kernel<4, 2>, 4, 2);
}

extern "C" int rk_drp_layer0_backward(void* stream, int B, int D, int H, int n_seg,
                                      const int32_t* seg_off, const int32_t* seg_len,
                                      const float* x, const float* gz0, const float* W0T,
                                      float* gs, float* grad_W0b0, float* workspace) {
  if (B <= 0 || D <= 0 || D > 16 || H <= 0 || x == nullptr || gz0 == nullptr ||
      W0T == nullptr || gs == nullptr || grad_W0b0 == nullptr || workspace == nullptr)
    return RK_E_BADARG;
  RkActSegs S;
  int rc = make_segs(D, n_seg, seg_off, seg_len, &S);
  if (rc != 0) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks0 = rk_drp_head_backward_ctas(B);
  const int rpc = (B + chunks0 - 1) / chunks0;
  const int chunks = (B + rpc - 1) / rpc;
  dim3 grid((H + 31) / 32, chunks);
  rk_drp_layer0_bwd_kernel<<<grid, 256, (size_t)rpc * 16 * sizeof(float), st>>>(
      B, D, H, S, x, gz0, workspace, rpc);
  const int n = (D + 1) * H;
  rk_add_rows_kernel<<<(n + 127) / 128, 128, 0, st>>>(chunks, n, workspace, grad_W0b0, 1);
  // input cotangent: gs[b][d] += sum_c gz0[b][c] W0T[c][d], scattered to the fluents' blocks
  return launch_rowdot(st, 1, B, D, H, gz0, H, W0T, D, gs, 0, nullptr, S);
}

extern "C" int rk_rowdot(void* stream, int accumulate, int B, int N, int K, const float* A,
                         int lda, const float* W, int ldw, float* out, int ldo,
                         const float* bias) {
  if (B <= 0 || N <= 0 || N > 16 || K <= 0 || A == nullptr || W == nullptr || out == nullptr)
    return RK_E_BADARG;
  RkActSegs S;
  S.n = 0;
  return launch_rowdot((cudaStream_t)stream, accumulate, B, N, K, A, lda, W, ldw, out, ldo, bias, S);
}

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

#define BM 128
#define BN 128
#define BK 8
#define TM 8
#define TN 8

// C[M][N] (row-major, ldc) = epi( op(A)[M][K] * B[K][N] )
//   A_TRANS = 0: A stored [M][K] (lda >= K);  1: A stored [K][M] (lda >= M)
//   split-K: blockIdx.z handles K range [z*kchunk, (z+1)*kchunk); partial tiles go to
//   `out` = workspace + z*M*N (ldc = N) and the epilogue is applied by the reduce kernel.
template <bool A_TRANS>
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
        if (epi == 1 || epi == 2) v += bias[n];
        if (epi == 2) v = tanhf(v);
        if (epi == 3) { const float h = aux[(size_t)m * ldaux + n]; v *= (1.f - h * h); }
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
  if (epi == 1 || epi == 2) v += bias[n];
  if (epi == 2) v = tanhf(v);
  if (epi == 3) { const float h = aux[(size_t)m * ldaux + n]; v *= (1.f - h * h); }
  C[(size_t)m * ldc + n] = v;
}

// Skinny shapes (N <= 16: action heads, state cotangent blocks; M <= 16: bias and first-layer
// weight gradients): the 128x128 tile would spend > 90% of its FMAs on padding, and these
// products are bound by streaming the [B][h] operand once.  16 x 16 threads, thread (tx, ty)
// owns rows ty + 16 i and columns tx + 16 j of a BM_ x BN_ tile.
#define SBK 16
template <int BM_, int BN_, bool A_TRANS>
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
        if (epi == 1 || epi == 2) v += bias[n];
        if (epi == 2) v = tanhf(v);
        if (epi == 3) { const float h = aux[(size_t)m * ldaux + n]; v *= (1.f - h * h); }
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
#define RK_LAUNCH(KERNEL)                                                                     \
  KERNEL<<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, aux, ldaux, epi,        \
                               accumulate, kchunk, workspace)
  if (bn == 16) {
    if (a_trans) RK_LAUNCH((rk_sgemm_skinny_kernel<128, 16, true>));
    else RK_LAUNCH((rk_sgemm_skinny_kernel<128, 16, false>));
  } else if (bm == 16) {
    if (a_trans) RK_LAUNCH((rk_sgemm_skinny_kernel<16, 128, true>));
    else RK_LAUNCH((rk_sgemm_skinny_kernel<16, 128, false>));
  } else {
    if (a_trans) RK_LAUNCH(rk_sgemm_kernel<true>);
    else RK_LAUNCH(rk_sgemm_kernel<false>);
  }
#undef RK_LAUNCH
  if (splits > 1) {
    const size_t total = (size_t)M * N;
    rk_sgemm_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        M, N, splits, workspace, C, ldc, bias, aux, ldaux, epi, accumulate);
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
};

__global__ void rk_drp_actions_fwd_kernel(int B, int A, int stochastic, RkActSegs S,
                                          const float* __restrict__ heads,
                                          const float* __restrict__ noise, float train,
                                          const float* __restrict__ lo,
                                          const float* __restrict__ hi, float ls_lo, float ls_hi,
                                          float* __restrict__ actions,
                                          float* __restrict__ entropy) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int ld = stochastic ? 2 * A : A;
  float ent = 0.f;
  for (int f = 0; f < S.n; ++f) {
    for (int e = 0; e < S.len[f]; ++e) {
      const int a = S.off[f] + e;
      const size_t w = (size_t)S.off[f] * B + (size_t)b * S.len[f] + e;
      float act = heads[(size_t)b * ld + a];
      if (stochastic) {
        const float ls = fminf(fmaxf(heads[(size_t)b * ld + A + a], ls_lo), ls_hi);
        ent += ls;
        const float z = (noise != nullptr) ? noise[w] : 0.f;
        act = act + train * expf(ls) * z;
      }
      if (lo != nullptr) act = fminf(fmaxf(act, lo[a]), hi[a]);
      actions[w] = act;
    }
  }
  if (entropy != nullptr) entropy[b] = ent;
}

__global__ void rk_drp_actions_bwd_kernel(int B, int A, int stochastic, RkActSegs S,
                                          const float* __restrict__ heads,
                                          const float* __restrict__ noise, float train,
                                          const float* __restrict__ lo,
                                          const float* __restrict__ hi, float ls_lo, float ls_hi,
                                          const float* __restrict__ gact,
                                          float* __restrict__ gheads) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int ld = stochastic ? 2 * A : A;
  for (int f = 0; f < S.n; ++f) {
    for (int e = 0; e < S.len[f]; ++e) {
      const int a = S.off[f] + e;
      const size_t w = (size_t)S.off[f] * B + (size_t)b * S.len[f] + e;
      const float mu = heads[(size_t)b * ld + a];
      float pre = mu, ex = 0.f, z = 0.f, raw = 0.f;
      if (stochastic) {
        raw = heads[(size_t)b * ld + A + a];
        const float ls = fminf(fmaxf(raw, ls_lo), ls_hi);
        z = (noise != nullptr) ? noise[w] : 0.f;
        ex = expf(ls);
        pre = mu + train * ex * z;
      }
      float g = gact[w];
      // jnp.clip: gradient 1 inside [lo, hi] (bounds included), 0 outside
      if (lo != nullptr && (pre < lo[a] || pre > hi[a])) g = 0.f;
      gheads[(size_t)b * ld + a] = g;
      if (stochastic) {
        const bool inside = raw >= ls_lo && raw <= ls_hi;
        gheads[(size_t)b * ld + A + a] = inside ? g * train * ex * z : 0.f;
      }
    }
  }
}

static int make_segs(int A, int n_seg, const int32_t* seg_off, const int32_t* seg_len,
                     RkActSegs* S) {
  if (n_seg <= 0 || n_seg > RK_MAX_ACTION_FLUENTS || seg_off == nullptr || seg_len == nullptr)
    return RK_E_BADARG;
  S->n = n_seg;
  int total = 0;
  for (int i = 0; i < n_seg; ++i) {
    if (seg_off[i] != total || seg_len[i] <= 0) return RK_E_BADARG;   // contiguous, in order
    S->off[i] = seg_off[i];
    S->len[i] = seg_len[i];
    total += seg_len[i];
  }
  return total == A ? 0 : RK_E_BADARG;
}

extern "C" int rk_drp_actions_forward(void* stream, int B, int A, int stochastic, int n_seg,
                                      const int32_t* seg_off, const int32_t* seg_len,
                                      const float* heads, const float* noise, float train,
                                      const float* lo, const float* hi, float ls_lo, float ls_hi,
                                      float* actions, float* entropy) {
  if (B <= 0 || A <= 0 || heads == nullptr || actions == nullptr) return RK_E_BADARG;
  if ((lo == nullptr) != (hi == nullptr)) return RK_E_BADARG;
  RkActSegs S;
  int rc = make_segs(A, n_seg, seg_off, seg_len, &S);
  if (rc != 0) return rc;
  rk_drp_actions_fwd_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      B, A, stochastic, S, heads, noise, train, lo, hi, ls_lo, ls_hi, actions, entropy);
  return (int)cudaGetLastError();
}

extern "C" int rk_drp_actions_backward(void* stream, int B, int A, int stochastic, int n_seg,
                                       const int32_t* seg_off, const int32_t* seg_len,
                                       const float* heads, const float* noise, float train,
                                       const float* lo, const float* hi, float ls_lo, float ls_hi,
                                       const float* gact, float* gheads) {
  if (B <= 0 || A <= 0 || heads == nullptr || gact == nullptr || gheads == nullptr)
    return RK_E_BADARG;
  if ((lo == nullptr) != (hi == nullptr)) return RK_E_BADARG;
  RkActSegs S;
  int rc = make_segs(A, n_seg, seg_off, seg_len, &S);
  if (rc != 0) return rc;
  rk_drp_actions_bwd_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      B, A, stochastic, S, heads, noise, train, lo, hi, ls_lo, ls_hi, gact, gheads);
  return (int)cudaGetLastError();
}

// ---- first dense layer straight from the fluent-major state (planner.py:1300-1313, 1358-1361)
// out[b][j] = tanh(sum_d x[b][d] * W0[d][j] + b0[j]), where the state vector x[b][:] is the
// concatenation of the fluent blocks ([B][n_f] at word offset off_f * B).  D is small (the state
// dimension), so the kernel is bound by writing out (B x H floats once).  Optionally also packs
// xp[b][0..D) = x[b][:], xp[b][D] = 1: the operand of the fused weight+bias gradient GEMM.
__global__ void __launch_bounds__(256)
rk_drp_layer0_kernel(int B, int D, int H, RkActSegs S, const float* __restrict__ x,
                     const float* __restrict__ W0, const float* __restrict__ b0,
                     float* __restrict__ out, float* __restrict__ xp) {
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
      out[(size_t)b * H + j] = tanhf(acc + b0[j]);
    }
  }
}

extern "C" int rk_drp_layer0(void* stream, int B, int D, int H, int n_seg,
                             const int32_t* seg_off, const int32_t* seg_len, const float* x,
                             const float* W0, const float* b0, float* out, float* xp) {
  if (B <= 0 || D <= 0 || H <= 0 || x == nullptr || W0 == nullptr || b0 == nullptr ||
      out == nullptr)
    return RK_E_BADARG;
  RkActSegs S;
  int rc = make_segs(D, n_seg, seg_off, seg_len, &S);
  if (rc != 0) return rc;
  const size_t smem = ((size_t)D * H + 64 * (size_t)D) * sizeof(float);
  if (smem > 200 * 1024) return RK_E_SMEM;
  if (smem > 48 * 1024) {
    cudaError_t ce = cudaFuncSetAttribute(rk_drp_layer0_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return (int)ce;
  }
  rk_drp_layer0_kernel<<<(B + 63) / 64, 256, smem, (cudaStream_t)stream>>>(B, D, H, S, x, W0, b0,
                                                                           out, xp);
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

// ---- skinny products with N <= 16 (action heads, state-cotangent blocks) ----------------------
// out[b][n] = (accumulate ? out[b][n] : 0) + sum_k A[b][k] * W[k][n] (+ bias[n]).
// One warp per row: lanes stride over k (coalesced 128-byte reads of A, read exactly once),
// W staged in shared memory with an odd row stride, N warp reductions per row.
__global__ void __launch_bounds__(256)
rk_rowdot_kernel(int B, int N, int K, const float* __restrict__ A, int lda,
                 const float* __restrict__ W, int ldw, float* __restrict__ out, int ldo,
                 const float* __restrict__ bias, int accumulate) {
  extern __shared__ float Wsh[];
  const int Np = N | 1;
  for (int i = threadIdx.x; i < K * N; i += 256) Wsh[(i / N) * Np + (i % N)] = W[(size_t)(i / N) * ldw + (i % N)];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = blockIdx.x * 8 + warp; b < B; b += gridDim.x * 8) {
    float acc[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) acc[n] = 0.f;
    const float* a = A + (size_t)b * lda;
    for (int k0 = lane; k0 < K; k0 += 256) {     // 8 independent loads in flight per lane
      float av[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) av[u] = (k0 + 32 * u < K) ? a[k0 + 32 * u] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (k0 + 32 * u < K) {
          const float* w = Wsh + (k0 + 32 * u) * Np;
#pragma unroll
          for (int n = 0; n < 16; ++n)
            if (n < N) acc[n] = fmaf(av[u], w[n], acc[n]);
        }
      }
    }
    float mine = 0.f;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      if (n < N) {
        float v = acc[n];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == n) mine = v;
      }
    }
    if (lane < N) {
      float v = mine;
      if (bias != nullptr) v += bias[lane];
      if (accumulate) v += out[(size_t)b * ldo + lane];
      out[(size_t)b * ldo + lane] = v;
    }
  }
}

extern "C" int rk_rowdot(void* stream, int accumulate, int B, int N, int K, const float* A,
                         int lda, const float* W, int ldw, float* out, int ldo,
                         const float* bias) {
  if (B <= 0 || N <= 0 || N > 16 || K <= 0 || A == nullptr || W == nullptr || out == nullptr)
    return RK_E_BADARG;
  const size_t smem = (size_t)K * (N | 1) * sizeof(float);
  if (smem > 200 * 1024) return RK_E_SMEM;
  if (smem > 48 * 1024) {
    cudaError_t ce = cudaFuncSetAttribute(rk_rowdot_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return (int)ce;
  }
  int grid = (B + 7) / 8;
  if (grid > 148 * 8) grid = 148 * 8;
  rk_rowdot_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(B, N, K, A, lda, W, ldw, out, ldo,
                                                               bias, accumulate ? 1 : 0);
  return (int)cudaGetLastError();
}

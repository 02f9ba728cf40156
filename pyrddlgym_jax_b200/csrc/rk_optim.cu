// rk_optim.cu -- gradient reduction, loss and optimiser kernels (sm_100a).
//
// Replaces the device work of planner.py:2811-2819 (loss = -utility(returns)) and
// planner.py:2885-2899 (optax update, apply_updates, box projection).  Every reduction
// here runs in a fixed order (no float atomics) so results are reproducible run to run and
// identical on every rank after the gradient all-reduce.

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/rk.h"

// ---- fixed-order column sums over the per-warp partial rows ------------------------------
// pass 1: rows are split into `chunks` contiguous groups; each group is summed into its own
// first row (in place).  pass 2: the group heads are summed into grad.
__global__ void rk_reduce_rows_kernel(float* __restrict__ gradp, int rows, int n, int chunks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (i >= n) return;
  const int per = (rows + chunks - 1) / chunks;
  const int r0 = c * per, r1 = min(rows, r0 + per);
  if (r0 >= r1) return;
  float acc = 0.f;
  const float* p = gradp + (size_t)r0 * n + i;
  int r = r0;
  for (; r + 4 <= r1; r += 4) {          // 4 independent loads in flight, fixed add order
    float v0 = p[0], v1 = p[(size_t)n], v2 = p[2 * (size_t)n], v3 = p[3 * (size_t)n];
    acc = (((acc + v0) + v1) + v2) + v3;
    p += 4 * (size_t)n;
  }
  for (; r < r1; ++r) { acc += *p; p += n; }
  gradp[(size_t)r0 * n + i] = acc;
}

__global__ void rk_reduce_heads_kernel(const float* __restrict__ gradp, int rows, int n,
                                       int chunks, float* __restrict__ grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int per = (rows + chunks - 1) / chunks;
  float acc = 0.f;
  for (int c = 0; c < chunks; ++c) {
    const int r0 = c * per;
    if (r0 >= rows) break;
    acc += gradp[(size_t)r0 * n + i];
  }
  grad[i] = acc;
}

// few rows (a straight-line plan at B = 4096 has 512 warp rows): one launch, one memory round trip.
// A CTA owns 32 columns; each of its 32 warps sums a contiguous slice of at most 32 rows with all
// loads of a batch of 8 in flight, and the slices are added in slice order.
__global__ void __launch_bounds__(1024)
rk_reduce_small_kernel(const float* __restrict__ gradp, int rows, int n, float* __restrict__ grad) {
  __shared__ float part[32][33];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + c;
  const int per = (rows + 31) / 32;
  const int r0 = g * per, r1 = min(rows, r0 + per);
  float acc = 0.f;
  if (i < n) {
    const float* p = gradp + (size_t)r0 * n + i;
    for (int r = r0; r < r1; r += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (r + u < r1) ? p[(size_t)u * n] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
      p += 8 * (size_t)n;
    }
  }
  part[g][c] = acc;
  __syncthreads();
  if (g == 0 && i < n) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) s += part[q][c];
    grad[i] = s;
  }
}

extern "C" int rk_reduce_partials(void* stream, const float* gradp, int num_rows, int n,
                                  float* grad) {
  if (gradp == nullptr || grad == nullptr || num_rows <= 0 || n <= 0) return RK_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (num_rows <= 1024) {
    rk_reduce_small_kernel<<<(n + 31) / 32, 1024, 0, st>>>(gradp, num_rows, n, grad);
    return (int)cudaGetLastError();
  }
  int chunks = 1;
  // enough CTAs to cover the machine when the parameter vector is short
  const int col_blocks = (n + 127) / 128;
  while (chunks < 64 && chunks * 2 * 32 <= num_rows && col_blocks * chunks < 148 * 8) chunks *= 2;
  dim3 grid(col_blocks, chunks);
  rk_reduce_rows_kernel<<<grid, 128, 0, st>>>((float*)gradp, num_rows, n, chunks);
  rk_reduce_heads_kernel<<<col_blocks, 128, 0, st>>>(gradp, num_rows, n, chunks, grad);
  return (int)cudaGetLastError();
}

// ---- deterministic single-CTA statistics: sum of squares + all-zero test -----------------
__device__ __forceinline__ float block_sum_1024(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = (l < (int)(blockDim.x >> 5)) ? sh[l] : 0.f;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (l == 0) sh[0] = r;
  }
  __syncthreads();
  r = sh[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(1024)
rk_grad_stats_kernel(const float* __restrict__ g, int n, float* __restrict__ stats,
                     int32_t* __restrict__ zero_flag) {
  __shared__ float sh[32];
  float ss = 0.f, nz = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float v = g[i];
    ss += v * v;
    nz += (fabsf(v) > 1e-8f || v != v) ? 1.f : 0.f;     // jnp.allclose(g, 0): |g| <= 1e-8
  }
  ss = block_sum_1024(ss, sh);
  nz = block_sum_1024(nz, sh);
  if (threadIdx.x == 0) {
    stats[0] = ss;
    if (zero_flag != nullptr) *zero_flag = (nz == 0.f) ? 1 : 0;
  }
}

// One parameter's optimiser update + apply_updates + box projection (planner.py:2885-2899), shared by every
// kernel that applies a step so that they agree to the bit.
// hyper = {lr, decay|b1, eps, b2, clip_norm, momentum, step, 1 - decay|b1, 1 - b2}
__device__ __forceinline__ float rk_opt_update(int kind, const float* __restrict__ hyper, int n, int i,
                                               float g, float p, float* __restrict__ state,
                                               const float* __restrict__ lo,
                                               const float* __restrict__ hi) {
  const float lr = hyper[0], d1 = hyper[1], eps = hyper[2], b2 = hyper[3];
  const float mom = hyper[5], step = hyper[6];
  const float om1 = hyper[7], om2 = hyper[8];      // 1 - decay|b1, 1 - b2 rounded from double (optax)
  float upd;
  if (kind == 1) {                       // optax.rmsprop: nu <- d nu + (1-d) g^2 ; g * rsqrt(nu + eps)
    float nu = d1 * state[i] + om1 * g * g;
    state[i] = nu;
    upd = g / sqrtf(nu + eps);
  } else if (kind == 2) {                // optax.adam with bias correction, step is 1-based
    float mu = d1 * state[i] + om1 * g;
    float nu = b2 * state[n + i] + om2 * g * g;
    state[i] = mu;
    state[n + i] = nu;
    float mh = mu / (1.f - powf(d1, step));
    float nh = nu / (1.f - powf(b2, step));
    upd = mh / (sqrtf(nh) + eps);
  } else {                               // optax.sgd (optional heavy-ball momentum)
    if (mom > 0.f) {
      float m = g + mom * state[i];
      state[i] = m;
      upd = m;
    } else {
      upd = g;
    }
  }
  p = p - lr * upd;                      // optax.apply_updates (planner.py:2898)
  if (lo != nullptr) p = fmaxf(p, lo[i]);   // box projection (planner.py:896-904)
  if (hi != nullptr) p = fminf(p, hi[i]);
  return p;
}

__global__ void rk_optimizer_kernel(int kind, const float* __restrict__ hyper, int n,
                                    float* __restrict__ params, float* __restrict__ state,
                                    const float* __restrict__ grad, const float* __restrict__ lo,
                                    const float* __restrict__ hi, const float* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float clip = hyper[4];
  float g = grad[i];
  if (clip > 0.f) {                      // optax.clip_by_global_norm (planner.py:2368-2369)
    const float gn = sqrtf(stats[0]);
    if (gn > clip) g = g * (clip / gn);
  }
  params[i] = rk_opt_update(kind, hyper, n, i, g, params[i], state, lo, hi);
}

// short parameter vectors: statistics and update in one single-CTA launch
__global__ void __launch_bounds__(1024)
rk_optimizer_small_kernel(int kind, const float* __restrict__ hyper, int n, float* params,
                          float* state, const float* __restrict__ grad,
                          const float* __restrict__ lo, const float* __restrict__ hi,
                          int32_t* __restrict__ zero_flag) {
  __shared__ float sh[32];
  float ss = 0.f, nz = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = grad[i];
    ss += v * v;
    nz += (fabsf(v) > 1e-8f || v != v) ? 1.f : 0.f;
  }
  ss = block_sum_1024(ss, sh);
  nz = block_sum_1024(nz, sh);
  if (threadIdx.x == 0 && zero_flag != nullptr) *zero_flag = (nz == 0.f) ? 1 : 0;
  const float clip = hyper[4];
  float scale = 1.f;
  bool clipped = false;
  if (clip > 0.f) {
    const float gn = sqrtf(ss);
    if (gn > clip) { scale = clip / gn; clipped = true; }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float g = grad[i];
    if (clipped) g = g * scale;
    params[i] = rk_opt_update(kind, hyper, n, i, g, params[i], state, lo, hi);
  }
}

static float* g_stats[64] = {nullptr};

extern "C" int rk_optimizer_step(void* stream, int kind, const float* hyper, int n,
                                 float* params, float* state, const float* grad,
                                 const float* lo, const float* hi, int32_t* zero_flag) {
  if (hyper == nullptr || params == nullptr || grad == nullptr || n <= 0) return RK_E_BADARG;
  if (kind < 0 || kind > 2) return RK_E_BADARG;
  if (kind != 0 && state == nullptr) return RK_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 16384) {
    rk_optimizer_small_kernel<<<1, 1024, 0, st>>>(kind, hyper, n, params, state, grad, lo, hi,
                                                  zero_flag);
    return (int)cudaGetLastError();
  }
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return RK_E_BADARG;
  if (g_stats[dev] == nullptr) {
    cudaError_t ce = cudaMalloc((void**)&g_stats[dev], 16);
    if (ce != cudaSuccess) return (int)ce;
  }
  rk_grad_stats_kernel<<<1, 1024, 0, st>>>(grad, n, g_stats[dev], zero_flag);
  rk_optimizer_kernel<<<(n + 255) / 256, 256, 0, st>>>(kind, hyper, n, params, state, grad, lo, hi,
                                                        g_stats[dev]);
  return (int)cudaGetLastError();
}

// ---- gradient reduction fused with the optimiser step ------------------------------------------------
// The tail of the adjoint: the per-warp partial rows are summed (same order as rk_reduce_partials) and the
// CTA that owns a column applies the optimiser step to it at once -- gradient, update, apply_updates, box
// projection in ONE launch (no global-norm clipping: that needs the whole gradient first).  The all-zero
// flag of planner.py:2861-2864 is agreed through two counters: every CTA adds its non-zero count, the last
// one to arrive writes the flag and resets them (integer atomics: the result does not depend on the order).
__device__ __forceinline__ void rk_zero_flag_vote(int nz_cta, int32_t* counters, int32_t* zero_flag) {
  if (zero_flag == nullptr) return;
  if (nz_cta) atomicAdd(&counters[0], nz_cta);
  __threadfence();
  const int ticket = atomicAdd(&counters[1], 1);
  if (ticket == (int)gridDim.x - 1) {
    *zero_flag = (atomicAdd(&counters[0], 0) == 0) ? 1 : 0;
    counters[0] = 0;
    counters[1] = 0;
  }
}

__global__ void __launch_bounds__(1024)
rk_reduce_opt_small_kernel(const float* __restrict__ gradp, int rows, int n, float* __restrict__ grad,
                           int kind, const float* __restrict__ hyper, float* __restrict__ params,
                           float* __restrict__ state, const float* __restrict__ lo,
                           const float* __restrict__ hi, int32_t* zero_flag, int32_t* counters) {
  __shared__ float part[32][33];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + c;
  const int per = (rows + 31) / 32;
  const int r0 = g * per, r1 = min(rows, r0 + per);
  float acc = 0.f;
  if (i < n) {
    const float* p = gradp + (size_t)r0 * n + i;
    for (int r = r0; r < r1; r += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (r + u < r1) ? p[(size_t)u * n] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
      p += 8 * (size_t)n;
    }
  }
  part[g][c] = acc;
  __syncthreads();
  if (g == 0) {
    int nz = 0;
    if (i < n) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 32; ++q) s += part[q][c];
      grad[i] = s;
      nz = (fabsf(s) > 1e-8f || s != s) ? 1 : 0;
      params[i] = rk_opt_update(kind, hyper, n, i, s, params[i], state, lo, hi);
    }
    nz = __reduce_add_sync(0xffffffffu, nz);
    if (c == 0) rk_zero_flag_vote(nz, counters, zero_flag);
  }
}

__global__ void __launch_bounds__(128)
rk_reduce_opt_heads_kernel(const float* __restrict__ gradp, int rows, int n, int chunks,
                           float* __restrict__ grad, int kind, const float* __restrict__ hyper,
                           float* __restrict__ params, float* __restrict__ state,
                           const float* __restrict__ lo, const float* __restrict__ hi,
                           int32_t* zero_flag, int32_t* counters) {
  __shared__ int nz_s[4];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int nz = 0;
  if (i < n) {
    const int per = (rows + chunks - 1) / chunks;
    float acc = 0.f;
    for (int c = 0; c < chunks; ++c) {
      const int r0 = c * per;
      if (r0 >= rows) break;
      acc += gradp[(size_t)r0 * n + i];
    }
    grad[i] = acc;
    nz = (fabsf(acc) > 1e-8f || acc != acc) ? 1 : 0;
    params[i] = rk_opt_update(kind, hyper, n, i, acc, params[i], state, lo, hi);
  }
  nz = __reduce_add_sync(0xffffffffu, nz);
  if ((threadIdx.x & 31) == 0) nz_s[threadIdx.x >> 5] = nz;
  __syncthreads();
  if (threadIdx.x == 0) rk_zero_flag_vote(nz_s[0] + nz_s[1] + nz_s[2] + nz_s[3], counters, zero_flag);
}

extern "C" int rk_reduce_optimizer_step(void* stream, const float* gradp, int num_rows, int n,
                                        float* grad, int kind, const float* hyper, float* params,
                                        float* state, const float* lo, const float* hi,
                                        int32_t* zero_flag, int32_t* counters) {
  if (gradp == nullptr || grad == nullptr || num_rows <= 0 || n <= 0 || hyper == nullptr ||
      params == nullptr)
    return RK_E_BADARG;
  if (kind < 0 || kind > 2 || (kind != 0 && state == nullptr)) return RK_E_BADARG;
  if (zero_flag != nullptr && counters == nullptr) return RK_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (num_rows <= 1024) {
    rk_reduce_opt_small_kernel<<<(n + 31) / 32, 1024, 0, st>>>(gradp, num_rows, n, grad, kind, hyper,
                                                               params, state, lo, hi, zero_flag, counters);
    return (int)cudaGetLastError();
  }
  int chunks = 1;
  const int col_blocks = (n + 127) / 128;
  while (chunks < 64 && chunks * 2 * 32 <= num_rows && col_blocks * chunks < 148 * 8) chunks *= 2;
  dim3 grid(col_blocks, chunks);
  rk_reduce_rows_kernel<<<grid, 128, 0, st>>>((float*)gradp, num_rows, n, chunks);
  rk_reduce_opt_heads_kernel<<<col_blocks, 128, 0, st>>>(gradp, num_rows, n, chunks, grad, kind, hyper,
                                                         params, state, lo, hi, zero_flag, counters);
  return (int)cudaGetLastError();
}

// ---- loss = -mean(returns), gret[b] = -1/B ------------------------------------------------
__global__ void __launch_bounds__(1024)
rk_mean_loss_kernel(const float* __restrict__ returns, int B, float* __restrict__ loss) {
  __shared__ float sh[32];
  // four independent float4 streams per thread: 16 loads in flight instead of one (a single CTA summing
  // 262144 returns took 117 us with the scalar loop); the order is fixed, so the sum is reproducible
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int done = 0;
  if ((((uintptr_t)returns) & 15) == 0) {
    const float4* r4 = reinterpret_cast<const float4*>(returns);
    const int n4 = B >> 2;
    for (int i = threadIdx.x; i < n4; i += 4 * blockDim.x) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = i + k * blockDim.x;
        if (j < n4) {
          const float4 v = r4[j];
          acc[k] += (v.x + v.y) + (v.z + v.w);
        }
      }
    }
    done = n4 << 2;
  }
  float s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  for (int i = done + threadIdx.x; i < B; i += blockDim.x) s += returns[i];
  s = block_sum_1024(s, sh);
  if (threadIdx.x == 0) loss[0] = -s / (float)B;
}

__global__ void rk_fill_kernel(float* __restrict__ p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

extern "C" int rk_mean_loss(void* stream, const float* returns, int B, float* loss, float* gret) {
  if (returns == nullptr || B <= 0) return RK_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (loss != nullptr) rk_mean_loss_kernel<<<1, 1024, 0, st>>>(returns, B, loss);
  if (gret != nullptr) rk_fill_kernel<<<(B + 255) / 256, 256, 0, st>>>(gret, B, -1.f / (float)B);
  return (int)cudaGetLastError();
}

// ---- concurrency projection of the boolean plan parameters --------------------------------
// planner.py:914-959: after the box projection, every time step's boolean action parameters
// are projected so that at most `allowed` actions differ from their no-op value; the reference
// vmaps the per-step projection over the horizon axis, here one warp owns one time step.
// method 0 = SOGBOFA loop (planner.py:540-616), 1 = sort-based shift (planner.py:499-527).
struct RkProjSegs {
  int n;
  int off[RK_MAX_BOOL_FLUENTS], len[RK_MAX_BOOL_FLUENTS], noop[RK_MAX_BOOL_FLUENTS];
  float weight[RK_MAX_BOOL_FLUENTS];
};

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(128)
rk_project_slp_kernel(int method, int T, int A, float* __restrict__ params, RkProjSegs S,
                      int K, int allowed, int max_iter, int wrap_sigmoid, float min_prob,
                      int32_t* __restrict__ converged) {
  extern __shared__ float sh[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + warp;
  if (t >= T) return;
  float* a = sh + (size_t)warp * 2 * K;          // action values of this step
  float* prm = a + K;                            // raw parameters of this step
  float* row = params + (size_t)t * A;
  // gather the boolean parameters of the step (fluent order, flattened)
  int base = 0;
  for (int s = 0; s < S.n; ++s) {
    for (int i = lane; i < S.len[s]; i += 32) {
      const float p = row[S.off[s] + i];
      prm[base + i] = p;
      a[base + i] = wrap_sigmoid ? 1.f / (1.f + expf(-(S.weight[s] * p))) : p;
    }
    base += S.len[s];
  }
  __syncwarp();

  if (method == 1) {
    // ---- sort-based: shift so that only the top `allowed` scores stay above the threshold
    const float thr = wrap_sigmoid ? 0.f : 0.5f;
    float kth = 0.f;
    int found = 0;
    int b0 = 0;
    for (int s = 0; s < S.n && allowed < K; ++s) {
      for (int i0 = 0; i0 < S.len[s]; i0 += 32) {
        const int i = i0 + lane;
        const bool ok = i < S.len[s];
        float si = 0.f;
        int rank = -1;
        if (ok) {
          const float p = prm[b0 + i];
          si = S.noop[s] ? (wrap_sigmoid ? -p : 1.f - p) : p;
          rank = 0;
          int c0 = 0;
          for (int s2 = 0; s2 < S.n; ++s2) {
            for (int j = 0; j < S.len[s2]; ++j) {
              const float q = prm[c0 + j];
              const float sj = S.noop[s2] ? (wrap_sigmoid ? -q : 1.f - q) : q;
              rank += (sj > si || (sj == si && (c0 + j) < (b0 + i))) ? 1 : 0;
            }
            c0 += S.len[s2];
          }
        }
        const unsigned hit = __ballot_sync(0xffffffffu, ok && rank == allowed);
        if (hit) {
          kth = __shfl_sync(0xffffffffu, si, __ffs(hit) - 1);
          found = 1;
        }
      }
      b0 += S.len[s];
    }
    const float surplus = found ? fmaxf(kth - thr, 0.f) : 0.f;
    b0 = 0;
    for (int s = 0; s < S.n; ++s) {
      float lo, hi;
      if (wrap_sigmoid) {
        lo = logf(min_prob / (1.f - min_prob)) / S.weight[s];
        hi = logf((1.f - min_prob) / min_prob) / S.weight[s];
      } else {
        lo = min_prob;
        hi = 1.f - min_prob;
      }
      for (int i = lane; i < S.len[s]; i += 32) {
        const float p = prm[b0 + i];
        const float sh_ = S.noop[s] ? p + surplus : p - surplus;
        row[S.off[s] + i] = fminf(fmaxf(sh_, lo), hi);
      }
      b0 += S.len[s];
    }
    return;
  }

  // ---- SOGBOFA: subtract surplus / k from every active action until the sum fits
  float surplus = 0.f;
  int k = 0;
  for (int it = 0; it <= max_iter; ++it) {
    float sum = 0.f;
    int cnt = 0;
    int b0 = 0;
    const float amount = (it == 0 || k == 0) ? 0.f : surplus / (float)k;
    for (int s = 0; s < S.n; ++s) {
      for (int i = lane; i < S.len[s]; i += 32) {
        float v = a[b0 + i];
        if (it > 0) {
          v = S.noop[s] ? fminf(v + amount, 1.f) : fmaxf(v - amount, 0.f);
          a[b0 + i] = v;
        }
        const float act = S.noop[s] ? 1.f - v : v;
        sum += act;
        cnt += (act != 0.f) ? 1 : 0;
      }
      b0 += S.len[s];
    }
    surplus = fmaxf(warp_sum(sum) - (float)allowed, 0.f);
    k = warp_sum_i(cnt);
    if (!(it < max_iter && surplus > 0.f && k > 0)) break;
  }
  __syncwarp();
  if (lane == 0 && surplus > 0.f && converged != nullptr) atomicAnd(converged, 0);

  // remaining violations: set the first `excess` still-active actions to 0.5
  int total = 0;
  {
    int b0 = 0, cnt = 0;
    for (int s = 0; s < S.n; ++s) {
      for (int i = lane; i < S.len[s]; i += 32) {
        const float v = a[b0 + i];
        cnt += (S.noop[s] ? (v < 0.5f) : (v > 0.5f)) ? 1 : 0;
      }
      b0 += S.len[s];
    }
    total = warp_sum_i(cnt);
  }
  int excess = max(total - allowed, 0);
  const float amin = min_prob, amax = 1.f - min_prob;
  int b0 = 0;
  for (int s = 0; s < S.n; ++s) {
    int seen = 0;                                  // active flags before this chunk
    int replaced = 0;
    for (int i0 = 0; i0 < S.len[s]; i0 += 32) {
      const int i = i0 + lane;
      const bool ok = i < S.len[s];
      float v = ok ? fminf(fmaxf(a[b0 + i], amin), amax) : 0.f;
      const bool active = ok && (S.noop[s] ? (v < 0.5f) : (v > 0.5f));
      const unsigned m = __ballot_sync(0xffffffffu, active);
      const int rank = seen + __popc(m & ((2u << lane) - 1u));      // inclusive cumsum
      const bool rep = active && rank <= excess;
      if (rep) v = 0.5f;
      replaced += __popc(__ballot_sync(0xffffffffu, rep));
      seen += __popc(m);
      if (ok) {
        const float w = S.weight[s];
        row[S.off[s] + i] = wrap_sigmoid ? logf(v / (1.f - v)) / w : v;
      }
    }
    excess = max(excess - replaced, 0);
    b0 += S.len[s];
  }
}

extern "C" int rk_project_slp(void* stream, int method, int T, int A, float* params, int n_seg,
                              const int32_t* seg_off, const int32_t* seg_len,
                              const int32_t* seg_noop, const float* seg_weight, int allowed,
                              int max_iter, int wrap_sigmoid, float min_prob,
                              int32_t* converged) {
  if (params == nullptr || T < 0 || A <= 0 || n_seg < 0 || n_seg > RK_MAX_BOOL_FLUENTS)
    return RK_E_BADARG;
  if (method < 0 || method > 1 || allowed < 0) return RK_E_BADARG;
  if (n_seg > 0 && (seg_off == nullptr || seg_len == nullptr || seg_noop == nullptr ||
                    seg_weight == nullptr)) return RK_E_BADARG;
  if (T == 0 || n_seg == 0) return 0;
  RkProjSegs S;
  S.n = n_seg;
  int K = 0;
  for (int s = 0; s < n_seg; ++s) {
    if (seg_off[s] < 0 || seg_len[s] < 0 || seg_off[s] + seg_len[s] > A) return RK_E_BADARG;
    S.off[s] = seg_off[s]; S.len[s] = seg_len[s]; S.noop[s] = seg_noop[s];
    S.weight[s] = seg_weight[s];
    K += seg_len[s];
  }
  if (K == 0) return 0;
  const int wpc = 4;
  const size_t smem = (size_t)wpc * 2 * K * sizeof(float);
  if (smem > 200 * 1024) return RK_E_SMEM;
  cudaStream_t st = (cudaStream_t)stream;
  if (smem > 48 * 1024) {
    cudaError_t ce = cudaFuncSetAttribute(rk_project_slp_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return (int)ce;
  }
  rk_project_slp_kernel<<<(T + wpc - 1) / wpc, wpc * 32, smem, st>>>(
      method, T, A, params, S, K, allowed, max_iter, wrap_sigmoid, min_prob, converged);
  return (int)cudaGetLastError();
}

// ---- initial-state tiling on the device (init_sim_state, compiler.py:700) ---------------------
// The reference tiles the un-batched initial fluent values to [B, ...] on the host and ships the
// tiled arrays; here only the S state words cross PCIe and every rollout's copy is written
// straight into the fluent-major s0 buffer (fluent f = block [B][n_f] at word offset off_f * B).
struct RkStateSegs {
  int n;
  int off[RK_MAX_STATE_FLUENTS], len[RK_MAX_STATE_FLUENTS];
};

__global__ void rk_tile_state_kernel(int B, RkStateSegs S, const uint32_t* __restrict__ init,
                                     uint32_t* __restrict__ s0) {
  const int f = blockIdx.y;
  const long long total = (long long)B * S.len[f];
  uint32_t* dst = s0 + (size_t)S.off[f] * B;
  const uint32_t* src = init + S.off[f];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i % S.len[f]];
}

extern "C" int rk_tile_state(void* stream, int B, int n_seg, const int32_t* seg_off,
                             const int32_t* seg_len, const void* init, void* s0) {
  if (B <= 0 || n_seg <= 0 || n_seg > RK_MAX_STATE_FLUENTS || seg_off == nullptr ||
      seg_len == nullptr || init == nullptr || s0 == nullptr)
    return RK_E_BADARG;
  RkStateSegs S;
  S.n = n_seg;
  int maxlen = 1;
  for (int i = 0; i < n_seg; ++i) {
    if (seg_off[i] < 0 || seg_len[i] <= 0) return RK_E_BADARG;
    S.off[i] = seg_off[i];
    S.len[i] = seg_len[i];
    if (seg_len[i] > maxlen) maxlen = seg_len[i];
  }
  long long work = (long long)B * maxlen;
  int gx = (int)((work + 255) / 256);
  if (gx > 148 * 4) gx = 148 * 4;
  dim3 grid(gx, n_seg);
  rk_tile_state_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(B, S, (const uint32_t*)init,
                                                               (uint32_t*)s0);
  return (int)cudaGetLastError();
}

// ---- supplied-noise tensor from a counter-based generator ---------------------------------------
// Fills the [T][N * B] noise buffer of a program (fluent-major per step: draw site s = block
// [B][n_s] at word offset off_s * B) in ONE launch: Philox4x32-10 keyed by (seed), counter =
// (element quad index, draw number), followed by the site's inverse-CDF / Box-Muller transform.
// Replaces: jax.random.split + jax.random.<law> at every draw site of a step (compiler.py:1627-2240
// draw their noise from the step's key); the laws are the standard ones of core/layout.py:draw_noise,
// which stays the host-side (torch generator) form used by the parity tests.
// The draw number lives in device memory (draws[0]) and is advanced by rk_noise_advance, so a captured
// CUDA graph draws fresh noise at every replay.
#define RK_NOISE_MAX_SITES 64
struct RkNoiseSites {
  int n;
  int kind[RK_NOISE_MAX_SITES];        // 0 uniform, 1 normal, 2 exponential, 3 gumbel, 4 logistic, 5 laplace, 6 cauchy, 7 gamma
  int off[RK_NOISE_MAX_SITES + 1];     // word offsets (in units of B words); off[n] = N
};

__device__ __forceinline__ void rk_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

// standard Gamma(alpha) by Marsaglia-Tsang rejection (alpha >= 1: d = alpha - 1/3, c = 1 / sqrt(9 d),
// accept d v with v = (1 + c x)^3 when log u < x^2 / 2 + d - d v + d log v; alpha < 1: Gamma(alpha + 1)
// times u^(1 / alpha)).  Every element owns a Philox stream: counter = (element, draw number, attempt),
// key = seed ^ "gamm"; the acceptance rate is >= 0.95, the attempt loop is bounded.
__device__ float rk_gamma_draw(long long elem, unsigned long long d, uint32_t k0, uint32_t k1, float alpha) {
  const float a = alpha < 1.f ? alpha + 1.f : alpha;
  const float dd = a - (1.f / 3.f), c = rsqrtf(9.f * dd);
  float g = dd;
  uint32_t r[4];
  for (uint32_t attempt = 0; attempt < 64u; ++attempt) {
    rk_philox4x32_10((uint32_t)elem, (uint32_t)(elem >> 32) ^ (attempt << 8), (uint32_t)d,
                     (uint32_t)(d >> 32), k0 ^ 0x67616d6du, k1 + attempt, r);
    const float u1 = ((float)(r[0] >> 8) + 0.5f) * 5.9604644775390625e-8f;
    const float u2 = ((float)(r[1] >> 8) + 0.5f) * 5.9604644775390625e-8f;
    const float u = ((float)(r[2] >> 8) + 0.5f) * 5.9604644775390625e-8f;
    float sn, cs;
    sincospif(2.f * u2, &sn, &cs);
    const float x = sqrtf(-2.f * logf(u1)) * cs;
    const float t = 1.f + c * x;
    if (t <= 0.f) continue;
    const float v = t * t * t;
    if (logf(u) < 0.5f * x * x + dd - dd * v + dd * logf(v)) {
      g = dd * v;
      break;
    }
  }
  if (alpha < 1.f) {
    const float ub = ((float)(r[3] >> 8) + 0.5f) * 5.9604644775390625e-8f;
    g *= powf(ub, 1.f / alpha);
  }
  return g;
}

// grid = (quads of one step's row, T): no 64-bit division per element; a quad's word index and draw site
// are found once and advanced element by element (block [B][n_s] of site s starts at column off_s * B)
__global__ void __launch_bounds__(256)
rk_draw_noise_kernel(uint32_t row_words, int B, RkNoiseSites S, uint32_t seed_lo, uint32_t seed_hi,
                     const unsigned long long* __restrict__ draws, float* __restrict__ out,
                     const float* __restrict__ conc) {
  const uint32_t qx = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  const uint32_t c0 = qx * 4u;
  if (c0 >= row_words) return;
  const unsigned long long d = draws != nullptr ? draws[0] : 0ull;
  uint32_t r[4];
  rk_philox4x32_10(qx, t, (uint32_t)d, (uint32_t)(d >> 32), seed_lo, seed_hi, r);
  const long long i0 = (long long)t * row_words + c0;
  uint32_t w = c0 / (uint32_t)B, rem = c0 - w * (uint32_t)B;        // word index, rollout
  int s = 0;
  while (s + 1 < S.n && (int)w >= S.off[s + 1]) ++s;
  // a normal site consumes its words in pairs (Box-Muller): (r0, r1) -> elements 0, 1; (r2, r3) -> 2, 3
  float z[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (c0 + j >= row_words) break;
    const int kind = S.kind[s];
    float v;
    if (kind == 0) {                               // uniform [0, 1)
      v = (float)(r[j] >> 8) * 5.9604644775390625e-8f;
    } else if (kind == 1) {                        // normal: Box-Muller on the word pair (j & ~1, j | 1)
      const float u1 = ((float)(r[j & ~1] >> 8) + 0.5f) * 5.9604644775390625e-8f;
      const float u2 = ((float)(r[j | 1] >> 8) + 0.5f) * 5.9604644775390625e-8f;
      const float rad = sqrtf(-2.f * logf(u1));
      float sn, cs;
      sincospif(2.f * u2, &sn, &cs);
      v = (j & 1) ? rad * sn : rad * cs;
    } else if (kind == 7) {                        // gamma: word of the site -> its concentration
      const uint32_t n_s = (uint32_t)(S.off[s + 1] - S.off[s]);
      const uint32_t within = (c0 + j) - (uint32_t)S.off[s] * (uint32_t)B;
      v = rk_gamma_draw(i0 + j, d, seed_lo, seed_hi, conc[S.off[s] + (int)(within % n_s)]);
    } else {
      const float u = ((float)(r[j] >> 8) + 0.5f) * 5.9604644775390625e-8f;      // (0, 1)
      switch (kind) {
        case 2: v = -log1pf(-u); break;              // exponential
        case 3: v = -logf(-logf(u)); break;          // gumbel
        case 4: v = logf(u) - log1pf(-u); break;     // logistic
        case 5: v = (u < 0.5f) ? logf(2.f * u) : -logf(2.f * (1.f - u)); break;     // laplace
        default: { float sn, cs; sincospif(u - 0.5f, &sn, &cs); v = sn / cs; break; }    // cauchy
      }
    }
    z[j] = v;
    if (++rem == (uint32_t)B) {                    // next word of the row (maybe the next site)
      rem = 0;
      ++w;
      if (s + 1 < S.n && (int)w >= S.off[s + 1]) ++s;
    }
  }
  if (c0 + 3 < row_words && (((uintptr_t)(out + i0)) & 15) == 0) {
    *reinterpret_cast<float4*>(out + i0) = make_float4(z[0], z[1], z[2], z[3]);
  } else {
    for (int j = 0; j < 4 && c0 + j < row_words; ++j) out[i0 + j] = z[j];
  }
}

__global__ void rk_noise_advance_kernel(unsigned long long* draws) { draws[0] += 1ull; }

extern "C" int rk_draw_noise(void* stream, int T, int B, int n_sites, const int32_t* kinds,
                             const int32_t* offs, unsigned long long seed,
                             const unsigned long long* draws, float* out, const float* conc) {
  if (T <= 0 || B <= 0 || n_sites <= 0 || n_sites > RK_NOISE_MAX_SITES || kinds == nullptr ||
      offs == nullptr || out == nullptr)
    return RK_E_BADARG;
  RkNoiseSites S;
  S.n = n_sites;
  for (int s = 0; s < n_sites; ++s) {
    if (kinds[s] < 0 || kinds[s] > 7 || offs[s + 1] <= offs[s]) return RK_E_BADARG;
    if (kinds[s] == 7 && conc == nullptr) return RK_E_BADARG;
    S.kind[s] = kinds[s];
    S.off[s] = offs[s];
  }
  if (offs[0] != 0) return RK_E_BADARG;
  S.off[n_sites] = offs[n_sites];
  const long long row_words = (long long)offs[n_sites] * B;
  if (row_words >= (1ll << 31) || T > 65535) return RK_E_BADARG;
  const unsigned quads = (unsigned)((row_words + 3) / 4);
  dim3 grid((quads + 255) / 256, (unsigned)T);
  rk_draw_noise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      (uint32_t)row_words, B, S, (uint32_t)seed, (uint32_t)(seed >> 32), draws, out, conc);
  return (int)cudaGetLastError();
}

extern "C" int rk_noise_advance(void* stream, unsigned long long* draws) {
  if (draws == nullptr) return RK_E_BADARG;
  rk_noise_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(draws);
  return (int)cudaGetLastError();
}

// ---- returns and loss of a closed-loop (deep reactive policy) update ----------------------------
// The policy is evaluated between T one-step rollout launches, so the per-step rewards [T][B] (and the
// per-step policy entropies [T][B]) are reduced afterwards (planner.py:2747-2765, 2811-2819):
//   returns[b] = sum_t disc[t] * reward[t][b]          (disc NULL = 1)
//   ent[b]     = sum_t entropy[t][b]                   (entropy NULL: not touched)
// one thread per rollout, coalesced over b; then  loss = -mean(returns) - coeff * mean(ent)  in one CTA
// (fixed order: deterministic).
__global__ void __launch_bounds__(256)
rk_drp_returns_kernel(int T, int B, const float* __restrict__ reward, const float* __restrict__ disc,
                      const float* __restrict__ entropy, float* __restrict__ returns,
                      float* __restrict__ ent) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float r = 0.f, e = 0.f;
  for (int t = 0; t < T; ++t) {
    const float w = disc != nullptr ? disc[t] : 1.f;
    r += w * reward[(size_t)t * B + b];
    if (entropy != nullptr) e += entropy[(size_t)t * B + b];
  }
  returns[b] = r;
  if (entropy != nullptr) ent[b] = e;
}

__global__ void __launch_bounds__(1024)
rk_drp_loss_kernel(const float* __restrict__ returns, const float* __restrict__ ent,
                   const float* __restrict__ coeff, int B, float* __restrict__ loss) {
  __shared__ float sh[32];
  float s = 0.f, e = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    s += returns[i];
    if (ent != nullptr) e += ent[i];
  }
  s = block_sum_1024(s, sh);
  __syncthreads();
  e = block_sum_1024(e, sh);
  if (threadIdx.x == 0)
    loss[0] = -s / (float)B - (ent != nullptr ? coeff[0] * (e / (float)B) : 0.f);
}

extern "C" int rk_drp_returns_loss(void* stream, int T, int B, const float* reward, const float* disc,
                                   const float* entropy, const float* ent_coeff, float* returns,
                                   float* ent_scratch, float* loss) {
  if (T <= 0 || B <= 0 || reward == nullptr || returns == nullptr) return RK_E_BADARG;
  if (entropy != nullptr && (ent_scratch == nullptr || ent_coeff == nullptr)) return RK_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  rk_drp_returns_kernel<<<(B + 255) / 256, 256, 0, st>>>(T, B, reward, disc, entropy, returns,
                                                         ent_scratch);
  if (loss != nullptr)
    rk_drp_loss_kernel<<<1, 1024, 0, st>>>(returns, entropy != nullptr ? ent_scratch : nullptr,
                                           ent_coeff, B, loss);
  return (int)cudaGetLastError();
}

// ---- W [r][c] -> W^T [c][r], optionally with its TF32 hi / lo split (once per optimiser step) -------
__global__ void __launch_bounds__(256)
rk_transpose_split_kernel(int r, int c, const float* __restrict__ W, float* __restrict__ WT,
                          float* __restrict__ hi, float* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;          // bx: column block of W, by: row block
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;         // 32 x 8
  for (int j = ty; j < 32; j += 8)
    if (by + j < r && bx + tx < c) tile[j][tx] = W[(size_t)(by + j) * c + bx + tx];
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int row = bx + j, col = by + tx;                        // element of W^T
    if (row < c && col < r) {
      const float v = tile[tx][j];
      const size_t o = (size_t)row * r + col;
      if (WT != nullptr) WT[o] = v;
      if (hi != nullptr) {
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
        const float h = __uint_as_float(t);
        hi[o] = h;
        lo[o] = v - h;
      }
    }
  }
}

extern "C" int rk_transpose_split(void* stream, int r, int c, const float* W, float* WT, float* WT_hi,
                                  float* WT_lo) {
  if (r <= 0 || c <= 0 || W == nullptr || (WT == nullptr && WT_hi == nullptr)) return RK_E_BADARG;
  if ((WT_hi == nullptr) != (WT_lo == nullptr)) return RK_E_BADARG;
  dim3 grid((c + 31) / 32, (r + 31) / 32);
  rk_transpose_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(r, c, W, WT, WT_hi, WT_lo);
  return (int)cudaGetLastError();
}

// ---- utilities other than the plain mean (planner.py:2161-2247, 2761-2763, 2817-2818) ------------
// loss = -U(returns) and its cotangent gret[b] = -dU/dreturn_b in closed form, one CTA, fixed-order
// sums (deterministic), so the iteration stays inside one CUDA graph.
//   kind 1 mean_var     U = mu - beta/2 var            2 mean_std   U = mu - beta/2 sigma
//        3 mean_semivar U = mu - beta/2 E[min(r-mu,0)^2]   4 mean_semidev: its square root
//        5 sharpe       U = (mu - p0) / (sigma + p1)   6 entropic   U = -(logsumexp(-beta r) - log B) / beta
//        7 symlog mean  U = mean(sign(r) log1p|r|)   (use_symlog_reward with the mean utility)
// p0 = beta (risk_free for sharpe), p1 = eps (sharpe).  scale multiplies the cotangent (1 / world size).
__global__ void __launch_bounds__(1024)
rk_utility_loss_kernel(int kind, int B, int b0, int n_local, const float* __restrict__ r, float p0,
                       float p1, float scale, float* __restrict__ loss, float* __restrict__ gret) {
  __shared__ float sh[32];
  __shared__ float bc[4];
  const float invB = 1.f / (float)B;
  float s = 0.f, mx = -CUDART_INF_F;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float v = r[i];
    s += kind == 7 ? copysignf(log1pf(fabsf(v)), v) : v;
    mx = fmaxf(mx, -p0 * v);
  }
  s = block_sum_1024(s, sh);
  __syncthreads();
  if (kind == 6) {                      // block max for the stabilised log-sum-exp
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = sh[0];
      for (int w = 1; w < 32; ++w) m = fmaxf(m, sh[w]);
      bc[1] = m;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) bc[0] = s * invB;
  __syncthreads();
  const float mu = bc[0], zmax = bc[1];
  if (kind == 7) {
    if (threadIdx.x == 0 && loss != nullptr) loss[0] = -mu;
    if (gret != nullptr)
      for (int i = b0 + threadIdx.x; i < b0 + n_local; i += blockDim.x)
        gret[i - b0] = -scale * invB / (1.f + fabsf(r[i]));
    return;
  }
  // second moment / lower partial moments / sum of exponentials
  float a = 0.f, d1 = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float c = r[i] - mu;
    if (kind == 1 || kind == 2 || kind == 5) {
      a += c * c;
    } else if (kind == 3 || kind == 4) {
      const float d = fminf(c, 0.f);
      a += d * d;
      d1 += d;
    } else {
      a += expf(-p0 * r[i] - zmax);
    }
  }
  a = block_sum_1024(a, sh);
  __syncthreads();
  d1 = block_sum_1024(d1, sh);
  __syncthreads();
  if (threadIdx.x == 0) {
    bc[2] = a;
    bc[3] = d1 * invB;
  }
  __syncthreads();
  const float A = bc[2], dmean = bc[3];
  const float var = A * invB, sigma = sqrtf(var);
  float U;
  switch (kind) {
    case 1: U = mu - 0.5f * p0 * var; break;
    case 2: U = mu - 0.5f * p0 * sigma; break;
    case 3: U = mu - 0.5f * p0 * var; break;
    case 4: U = mu - 0.5f * p0 * sigma; break;
    case 5: U = (mu - p0) / (sigma + p1); break;
    default: U = -(logf(A) + zmax - logf((float)B)) / p0; break;
  }
  if (threadIdx.x == 0 && loss != nullptr) loss[0] = -U;
  if (gret == nullptr) return;
  for (int i = b0 + threadIdx.x; i < b0 + n_local; i += blockDim.x) {
    const float c = r[i] - mu;
    float g;                             // dU / dr_i
    switch (kind) {
      case 1: g = invB * (1.f - p0 * c); break;
      case 2: g = invB * (1.f - 0.5f * p0 * c / sigma); break;
      case 3: g = invB * (1.f - p0 * (fminf(c, 0.f) - dmean)); break;
      case 4: g = invB * (1.f - 0.5f * p0 * (fminf(c, 0.f) - dmean) / sigma); break;
      case 5: g = invB * (1.f / (sigma + p1) - (mu - p0) * c / (sigma * (sigma + p1) * (sigma + p1))); break;
      default: g = expf(-p0 * r[i] - zmax) / A; break;
    }
    gret[i - b0] = -scale * g;
  }
}

extern "C" int rk_utility_loss_shard(void* stream, int kind, int B, int b0, int n_local,
                                     const float* returns, float p0, float p1, float scale, float* loss,
                                     float* gret) {
  if (kind < 1 || kind > 7 || B <= 0 || returns == nullptr || (loss == nullptr && gret == nullptr) ||
      b0 < 0 || n_local < 0 || b0 + n_local > B)
    return RK_E_BADARG;
  rk_utility_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(kind, B, b0, n_local, returns, p0, p1,
                                                                scale, loss, gret);
  return (int)cudaGetLastError();
}

extern "C" int rk_utility_loss(void* stream, int kind, int B, const float* returns, float p0, float p1,
                               float scale, float* loss, float* gret) {
  return rk_utility_loss_shard(stream, kind, B, 0, B, returns, p0, p1, scale, loss, gret);
}

// ---- the optional links of the optax chain around the optimiser, on the device ------------------
// planner.py:2349-2387 builds  clip_by_global_norm -> add_noise(eta, gamma, seed) -> optimiser ->
// reduce_on_plateau -> ema(decay)  and planner.py:2885-2899 applies the result and projects onto the box.
// Stage 1 (one CTA, fixed-order sums): step count t <- t + 1, all-zero test and global norm of the RAW
// gradient, noise deviation sqrt(eta / t^gamma), the plateau state machine of optax.contrib stepped with
// the loss of this gradient, the EMA debias 1 - decay^t; the derived scalars go to scal[4].
// Stage 2 (element-wise): g <- clip(g) + std * N(0, 1) (Philox4x32-10, counter = (quad, t), key = seed)
// -> optimiser update u -> step = -lr * u * plateau_scale -> ema <- d ema + (1 - d) step -> p <- p +
// ema / debias -> box.  Short vectors run both stages in one single-CTA launch.
struct RkChainCfg {
  float eta, gamma, ema_decay, factor, patience, rtol, atol, cooldown, accumulation, min_scale;
  int use_noise, use_plateau, use_ema;
};

__device__ void rk_chain_stats(const float* __restrict__ g, int n, const float* __restrict__ hyper,
                               const RkChainCfg& C, float* count, float* plateau,
                               const float* __restrict__ loss, int32_t* zero_flag, float* scal,
                               float* sh) {
  float ss = 0.f, nz = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = g[i];
    ss += v * v;
    nz += (fabsf(v) > 1e-8f || v != v) ? 1.f : 0.f;
  }
  ss = block_sum_1024(ss, sh);
  nz = block_sum_1024(nz, sh);
  if (threadIdx.x == 0) {
    if (zero_flag != nullptr) *zero_flag = (nz == 0.f) ? 1 : 0;
    const float t = count[0] + 1.f;
    count[0] = t;
    const float clip = hyper[4];
    float cs = 1.f;
    if (clip > 0.f) {
      const float gn = sqrtf(ss);
      if (gn > clip) cs = clip / gn;
    }
    float pscale = 1.f;
    if (C.use_plateau) {
      float scale = plateau[0], best = plateau[1], plat = plateau[2], cool = plateau[3];
      const float cnt = plateau[4], avg = plateau[5];
      const float ncnt = cnt + 1.f;
      const float navg = (cnt * avg + loss[0]) / ncnt;
      if (ncnt == C.accumulation) {
        const bool improved = navg < (1.f - C.rtol) * best - C.atol;
        if (improved) best = navg;
        const float cur = improved ? 0.f : plat + 1.f;
        const bool in_cd = cool > 0.f, hit = cur == C.patience;
        plat = (in_cd || hit) ? 0.f : cur;
        if (!in_cd && hit) scale = fmaxf(scale * C.factor, C.min_scale);
        cool = in_cd ? cool - 1.f : (hit ? C.cooldown : 0.f);
        plateau[0] = scale; plateau[1] = best; plateau[2] = plat; plateau[3] = cool;
        plateau[4] = 0.f; plateau[5] = 0.f;
      } else {
        plateau[4] = ncnt; plateau[5] = navg;
      }
      pscale = plateau[0];
    }
    scal[0] = cs;
    scal[1] = C.use_noise ? sqrtf(C.eta / powf(t, C.gamma)) : 0.f;
    scal[2] = pscale;
    scal[3] = C.use_ema ? 1.f - powf(C.ema_decay, t) : 1.f;
    scal[4] = t;
  }
}

__device__ __forceinline__ void rk_chain_quad(int q, int kind, const float* __restrict__ hyper, int n,
                                              float* params, float* state, const float* grad,
                                              const float* lo, const float* hi, const RkChainCfg& C,
                                              const float* scal, float* ema, uint32_t seed_lo,
                                              uint32_t seed_hi) {
  const float lr = hyper[0], d1 = hyper[1], eps = hyper[2], b2 = hyper[3];
  const float mom = hyper[5], step = hyper[6], om1 = hyper[7], om2 = hyper[8];
  const float cs = scal[0], sd = scal[1], ps = scal[2], debias = scal[3];
  float z[4] = {0.f, 0.f, 0.f, 0.f};
  if (C.use_noise) {
    uint32_t r[4];
    rk_philox4x32_10((uint32_t)q, 0u, (uint32_t)scal[4], 0x6e6f6973u, seed_lo, seed_hi, r);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u1 = ((float)(r[2 * h] >> 8) + 0.5f) * 5.9604644775390625e-8f;
      const float u2 = ((float)(r[2 * h + 1] >> 8) + 0.5f) * 5.9604644775390625e-8f;
      const float rad = sqrtf(-2.f * logf(u1));
      float sn, cn;
      sincospif(2.f * u2, &sn, &cn);
      z[2 * h] = rad * cn;
      z[2 * h + 1] = rad * sn;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = 4 * q + j;
    if (i >= n) break;
    const float g = grad[i] * cs + sd * z[j];
    float upd;
    if (kind == 1) {
      const float nu = d1 * state[i] + om1 * g * g;
      state[i] = nu;
      upd = g / sqrtf(nu + eps);
    } else if (kind == 2) {
      const float mu = d1 * state[i] + om1 * g;
      const float nu = b2 * state[n + i] + om2 * g * g;
      state[i] = mu;
      state[n + i] = nu;
      upd = (mu / (1.f - powf(d1, step))) / (sqrtf(nu / (1.f - powf(b2, step))) + eps);
    } else if (mom > 0.f) {
      const float m = g + mom * state[i];
      state[i] = m;
      upd = m;
    } else {
      upd = g;
    }
    float delta = -lr * upd * ps;                   // optimiser -> reduce_on_plateau
    if (C.use_ema) {                                // optax.ema with debias
      const float e = C.ema_decay * ema[i] + (1.f - C.ema_decay) * delta;
      ema[i] = e;
      delta = e / debias;
    }
    float p = params[i] + delta;                    // optax.apply_updates + box projection
    if (lo != nullptr) p = fmaxf(p, lo[i]);
    if (hi != nullptr) p = fminf(p, hi[i]);
    params[i] = p;
  }
}

__global__ void __launch_bounds__(1024)
rk_chain_small_kernel(int kind, const float* __restrict__ hyper, int n, float* params, float* state,
                      const float* __restrict__ grad, const float* __restrict__ lo,
                      const float* __restrict__ hi, int32_t* zero_flag, RkChainCfg C, float* count,
                      float* plateau, float* ema, const float* __restrict__ loss, uint32_t seed_lo,
                      uint32_t seed_hi) {
  __shared__ float sh[32];
  __shared__ float scal[8];
  rk_chain_stats(grad, n, hyper, C, count, plateau, loss, zero_flag, scal, sh);
  __syncthreads();
  for (int q = threadIdx.x; 4 * q < n; q += blockDim.x)
    rk_chain_quad(q, kind, hyper, n, params, state, grad, lo, hi, C, scal, ema, seed_lo, seed_hi);
}

__global__ void __launch_bounds__(1024)
rk_chain_stats_kernel(const float* __restrict__ grad, int n, const float* __restrict__ hyper,
                      RkChainCfg C, float* count, float* plateau, const float* __restrict__ loss,
                      int32_t* zero_flag, float* scal) {
  __shared__ float sh[32];
  rk_chain_stats(grad, n, hyper, C, count, plateau, loss, zero_flag, scal, sh);
}

__global__ void __launch_bounds__(256)
rk_chain_apply_kernel(int kind, const float* __restrict__ hyper, int n, float* params, float* state,
                      const float* __restrict__ grad, const float* __restrict__ lo,
                      const float* __restrict__ hi, RkChainCfg C, const float* __restrict__ scal,
                      float* ema, uint32_t seed_lo, uint32_t seed_hi) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (4 * q < n)
    rk_chain_quad(q, kind, hyper, n, params, state, grad, lo, hi, C, scal, ema, seed_lo, seed_hi);
}

extern "C" int rk_optimizer_chain_step(void* stream, int kind, const float* hyper, int n, float* params,
                                       float* state, const float* grad, const float* lo,
                                       const float* hi, int32_t* zero_flag, const float* cfg,
                                       float* count, float* plateau, float* ema, const float* loss,
                                       unsigned long long seed, float* scratch) {
  if (hyper == nullptr || params == nullptr || grad == nullptr || cfg == nullptr || count == nullptr ||
      n <= 0)
    return RK_E_BADARG;
  if (kind < 0 || kind > 2) return RK_E_BADARG;
  if (kind != 0 && state == nullptr) return RK_E_BADARG;
  RkChainCfg C;
  C.eta = cfg[0]; C.gamma = cfg[1]; C.ema_decay = cfg[2]; C.factor = cfg[3]; C.patience = cfg[4];
  C.rtol = cfg[5]; C.atol = cfg[6]; C.cooldown = cfg[7]; C.accumulation = cfg[8]; C.min_scale = cfg[9];
  C.use_noise = cfg[10] != 0.f; C.use_plateau = cfg[11] != 0.f; C.use_ema = cfg[12] != 0.f;
  if (C.use_plateau && (plateau == nullptr || loss == nullptr)) return RK_E_BADARG;
  if (C.use_ema && ema == nullptr) return RK_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const uint32_t s_lo = (uint32_t)seed, s_hi = (uint32_t)(seed >> 32);
  if (n <= 16384) {
    rk_chain_small_kernel<<<1, 1024, 0, st>>>(kind, hyper, n, params, state, grad, lo, hi, zero_flag, C,
                                              count, plateau, ema, loss, s_lo, s_hi);
    return (int)cudaGetLastError();
  }
  if (scratch == nullptr) return RK_E_BADARG;
  rk_chain_stats_kernel<<<1, 1024, 0, st>>>(grad, n, hyper, C, count, plateau, loss, zero_flag,
                                            scratch);
  const int quads = (n + 3) / 4;
  rk_chain_apply_kernel<<<(quads + 255) / 256, 256, 0, st>>>(kind, hyper, n, params, state, grad, lo, hi,
                                                              C, scratch, ema, s_lo, s_hi);
  return (int)cudaGetLastError();
}

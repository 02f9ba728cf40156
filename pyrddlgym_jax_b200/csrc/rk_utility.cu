// rk_utility.cu -- the order-statistic utilities of the planner loss on the device (sm_100a).
//
// Replaces the jit-compiled utility functions of planner.py:2195-2229 and their reverse-mode
// derivatives (value_and_grad of planner.py:2873 through planner.py:2817-2818):
//   var        U = percentile(R, 100 alpha)                      (linear interpolation, jnp / sj default)
//   cvar       U = v - mean(relu(v - R)) / alpha,  v = the percentile
//   cvar_ste   U = sum(mask R) / max(1, count(mask)),  mask = R <= percentile (a constant of the derivative)
//   expectile  max_iter fixed-point steps  e <- sum(w R) / sum(w),  w = alpha where R > e else 1 - alpha
//   gini       U = mean(R) - gamma * 2 * mean_ij relu(R_j - R_i)
// loss = -U and the cotangent gret[b] = -scale * dU/dR_b in closed form, so a planner iteration with any of
// the reference's thirteen utilities stays one CUDA graph.  The return vector is sorted once (CUB radix
// sort, key = return, value = rollout index); order statistics, tie groups (binary search on the sorted
// keys) and the pair sum of the Gini dispersion all read the sorted vector.  Sums run in a fixed order
// inside one CTA (deterministic); the per-rollout cotangents are written by a full grid.

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/rk.h"

namespace {

__device__ __forceinline__ float block_sum(float v, float* sh) {
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

__global__ void rk_iota_kernel(int* __restrict__ idx, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = i;
}

// scal: [0] percentile v, [1] weight of the lower order statistic, [2] weight of the upper one,
//       [3] count / sum of weights, [4] expectile of the step before the last, [5] mean
// sidx: [0] rollout of the lower order statistic, [1] of the upper one
struct UtilScal {
  float v, w_lo, w_hi, cnt, prev, mu;
  int i_lo, i_hi;
};

enum { U_VAR = 8, U_CVAR = 9, U_CVAR_STE = 10, U_EXPECTILE = 11, U_GINI = 12 };

__global__ void __launch_bounds__(1024)
rk_utility_scalar_kernel(int kind, int B, const float* __restrict__ r, const float* __restrict__ sk,
                         const int* __restrict__ si, float p0, int max_iter, float* __restrict__ loss,
                         UtilScal* __restrict__ out) {
  __shared__ float sh[32];
  const float invB = 1.f / (float)B;
  float U = 0.f;
  UtilScal S = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0, 0};
  if (kind == U_VAR || kind == U_CVAR || kind == U_CVAR_STE) {
    const double pos = (double)p0 * (double)(B - 1);   // jnp.percentile(method='linear')
    int lo = (int)floor(pos);
    lo = lo < 0 ? 0 : (lo > B - 1 ? B - 1 : lo);
    const int hi = lo + 1 > B - 1 ? B - 1 : lo + 1;
    const float f = (float)(pos - (double)lo);
    const float v = sk[lo] + f * (sk[hi] - sk[lo]);
    S.v = v; S.w_lo = 1.f - f; S.w_hi = f; S.i_lo = si[lo]; S.i_hi = si[hi];
    if (kind == U_VAR) {
      U = v;
    } else if (kind == U_CVAR) {
      float a = 0.f, c = 0.f;
      for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const float d = v - r[i];
        if (d > 0.f) { a += d; c += 1.f; }
      }
      a = block_sum(a, sh);
      c = block_sum(c, sh);
      U = v - a * invB / p0;
      S.cnt = c;
    } else {
      float a = 0.f, c = 0.f;
      for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const float x = r[i];
        if (x <= v) { a += x; c += 1.f; }
      }
      a = block_sum(a, sh);
      c = block_sum(c, sh);
      S.cnt = fmaxf(c, 1.f);
      U = a / S.cnt;
    }
  } else if (kind == U_EXPECTILE) {
    float s = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) s += r[i];
    float e = block_sum(s, sh) * invB, prev = e, wsum = (float)B;
    for (int it = 0; it < max_iter; ++it) {
      float a = 0.f, w = 0.f;
      for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const float x = r[i];
        const float wi = x > e ? p0 : 1.f - p0;
        a += wi * x;
        w += wi;
      }
      a = block_sum(a, sh);
      w = block_sum(w, sh);
      prev = e;
      wsum = w;
      e = a / w;
    }
    U = e;
    S.prev = prev;
    S.cnt = wsum;
    S.v = max_iter > 0 ? 1.f : 0.f;
  } else {                                              // gini: pair sum from the sorted vector
    float s = 0.f, a = 0.f;
    for (int k = threadIdx.x; k < B; k += blockDim.x) {
      const float x = sk[k];
      s += x;
      a += x * (float)(2 * k - B + 1);
    }
    s = block_sum(s, sh);
    a = block_sum(a, sh);
    S.mu = s * invB;
    U = S.mu - p0 * 2.f * a * invB * invB;
  }
  if (threadIdx.x == 0) {
    if (loss != nullptr) loss[0] = -U;
    *out = S;
  }
}

__device__ __forceinline__ int lower_bound_f(const float* __restrict__ a, int n, float x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int upper_bound_f(const float* __restrict__ a, int n, float x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// gret[b - b0] for the rollouts [b0, b0 + n) of the (global) return vector r[B]
__global__ void __launch_bounds__(256)
rk_utility_cotangent_kernel(int kind, int B, int b0, int n, const float* __restrict__ r,
                            const float* __restrict__ sk, float p0, float scale,
                            const UtilScal* __restrict__ Sp, float* __restrict__ gret) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int b = b0 + j;
  const UtilScal S = *Sp;
  const float invB = 1.f / (float)B, x = r[b];
  float dv = 0.f;                                       // d percentile / d R_b
  if (b == S.i_lo) dv += S.w_lo;
  if (b == S.i_hi) dv += S.w_hi;
  float g;
  switch (kind) {
    case U_VAR: g = dv; break;
    case U_CVAR: g = dv * (1.f - S.cnt * invB / p0) + ((S.v - x > 0.f) ? invB / p0 : 0.f); break;
    case U_CVAR_STE: g = (x <= S.v) ? 1.f / S.cnt : 0.f; break;
    case U_EXPECTILE: g = S.v != 0.f ? ((x > S.prev) ? p0 : 1.f - p0) / S.cnt : invB; break;
    default: {
      const int below = lower_bound_f(sk, B, x), above = B - upper_bound_f(sk, B, x);
      g = invB - p0 * 2.f * (float)(below - above) * invB * invB;
      break;
    }
  }
  gret[j] = -scale * g;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t sort_temp_bytes(int B) {
  size_t t = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t, (const float*)nullptr, (float*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, B);
  return t;
}

}  // namespace

extern "C" size_t rk_utility_sorted_work_bytes(int B) {
  if (B <= 0) return 0;
  return align256(sizeof(UtilScal)) + align256((size_t)B * 4) * 3 + align256(sort_temp_bytes(B));
}

extern "C" int rk_utility_sorted_loss(void* stream, int kind, int B, int b0, int n_local,
                                      const float* returns, float p0, int max_iter, float scale,
                                      float* loss, float* gret, void* work, size_t work_bytes) {
  if (kind < U_VAR || kind > U_GINI || B <= 0 || returns == nullptr || work == nullptr ||
      (loss == nullptr && gret == nullptr) || b0 < 0 || n_local < 0 || b0 + n_local > B)
    return RK_E_BADARG;
  if (work_bytes < rk_utility_sorted_work_bytes(B)) return RK_E_BADARG;
  if ((kind == U_VAR || kind == U_CVAR || kind == U_CVAR_STE) && !(p0 > 0.f && p0 <= 1.f))
    return RK_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)work;
  UtilScal* scal = (UtilScal*)w;            w += align256(sizeof(UtilScal));
  float* sk = (float*)w;                    w += align256((size_t)B * 4);
  int* idx = (int*)w;                       w += align256((size_t)B * 4);
  int* si = (int*)w;                        w += align256((size_t)B * 4);
  void* temp = (void*)w;
  if (kind != U_EXPECTILE) {
    size_t tb = sort_temp_bytes(B);
    rk_iota_kernel<<<(B + 255) / 256, 256, 0, st>>>(idx, B);
    cudaError_t ce = cub::DeviceRadixSort::SortPairs(temp, tb, returns, sk, (const int*)idx, si, B, 0, 32, st);
    if (ce != cudaSuccess) return (int)ce;
  }
  rk_utility_scalar_kernel<<<1, 1024, 0, st>>>(kind, B, returns, sk, si, p0, max_iter, loss, scal);
  if (gret != nullptr && n_local > 0)
    rk_utility_cotangent_kernel<<<(n_local + 255) / 256, 256, 0, st>>>(kind, B, b0, n_local, returns, sk,
                                                                        p0, scale, scal, gret);
  return (int)cudaGetLastError();
}

// rk_device.cuh -- scalar device functions with the reference's op semantics, shared by the
// interpreter kernel and the generated specialised kernels.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

__device__ __forceinline__ float stable_tanh(float x) {       // logic.py:57-65
  float ax = fabsf(x);
  float s;
  if (ax < 20.f) {
    float em = expm1f(2.f * ax);
    s = em / (em + 2.f);
  } else {
    s = 1.f - 2.f * expf(-2.f * ax);
  }
  return copysignf(1.f, x) * ((x == 0.f) ? 0.f : s);
}

// digamma(x) = d/dx ln Gamma(x): the derivative of LGAMMA in adjoint programs.  Reflection for
// x < 0.5, upward recurrence to x >= 6, then the asymptotic series (error < 1e-7 relative).
__device__ __forceinline__ float rk_digamma(float x) {
  float refl = 0.f;
  bool neg = false;
  if (x < 0.5f) {                          // psi(x) = psi(1 - x) - pi / tan(pi x)
    refl = 3.14159265358979f / tanf(3.14159265358979f * x);
    x = 1.f - x;
    neg = true;
  }
  float acc = 0.f;
#pragma unroll 1
  for (int i = 0; i < 6; ++i) {
    if (x < 6.f) { acc -= 1.f / x; x += 1.f; }
  }
  const float r = 1.f / x, r2 = r * r;
  const float ser = r2 * (1.f / 12.f - r2 * (1.f / 120.f - r2 * (1.f / 252.f - r2 * (1.f / 240.f))));
  const float v = acc + logf(x) - 0.5f * r - ser;
  return neg ? v - refl : v;
}

// ---- asynchronous global -> shared copies of one word per lane (prefetch ring of the generated
// kernels); a false predicate zero-fills the destination: with a source size of 0 nothing is read
// from global memory, so the (possibly out-of-range) address is never dereferenced
__device__ __forceinline__ void rk_cp_async4(uint32_t* dst, const uint32_t* src, bool pred) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int nbytes = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(nbytes)
               : "memory");
}
__device__ __forceinline__ void rk_cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void rk_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- branch-free variants used by relaxed (train / adjoint) programs -----------------------
// libdevice's sincosf / sqrtf and the IEEE division guard a slow path with a branch; every such
// branch ends a basic block, so a step with a dozen of them cannot be scheduled as one block
// and a warp that runs alone on its scheduler stalls on every dependent pair.  These versions
// have no branch (<= 1-2 ulp for normal operands; infinities, zeros and NaNs handled by a
// select on the seed result).  Exact programs never use them.
__device__ __forceinline__ float rk_fdiv(float x, float y) {
  float r0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(y));
  const float r = fmaf(fmaf(-y, r0, 1.f), r0, r0);          // Newton step on the reciprocal
  const float q0 = x * r;
  const float q1 = fmaf(fmaf(-y, q0, x), r, q0);            // residual correction of the quotient
  return (fabsf(q1) < CUDART_INF_F) ? q1 : x * r0;          // inf / 0 / NaN: the seed's answer
}

__device__ __forceinline__ float rk_fsqrt(float x) {
  float s0, y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s0) : "f"(x));
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  const float s1 = fmaf(fmaf(-s0, s0, x), 0.5f * y, s0);
  return (fabsf(s1) < CUDART_INF_F) ? s1 : s0;
}

__device__ __forceinline__ void rk_fsincos(float x, float* sn, float* cs) {
  // Cody-Waite reduction by pi/2 in three fp32 pieces (accurate for |x| up to ~1e5, the range
  // of libdevice's own fast path), then the cephes minimax polynomials on [-pi/4, pi/4]
  const float j = rintf(x * __uint_as_float(0x3f22f983u));
  const int i = (int)j;
  float r = fmaf(j, -__uint_as_float(0x3fc90fdbu), x);
  r = fmaf(j, -__uint_as_float(0xb33bbd2eu), r);
  r = fmaf(j, -__uint_as_float(0xa7000000u), r);
  const float z = r * r;
  const float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f),
                        z * r, r);
  const float cp = fmaf(fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z,
                                  4.166664568298827e-2f), z, -0.5f), z, 1.f);
  const bool sw = (i & 1) != 0;
  const float ss = sw ? cp : sp, cc = sw ? sp : cp;
  *sn = (i & 2) ? -ss : ss;
  *cs = ((i + 1) & 2) ? -cc : cc;
}

__device__ __forceinline__ float rk_fsin(float x) { float s, c; rk_fsincos(x, &s, &c); return s; }
__device__ __forceinline__ float rk_fcos(float x) { float s, c; rk_fsincos(x, &s, &c); return c; }
__device__ __forceinline__ float rk_ftan(float x) { float s, c; rk_fsincos(x, &s, &c); return rk_fdiv(s, c); }
__device__ __forceinline__ float rk_ftan_ieee(float x) { float s, c; rk_fsincos(x, &s, &c); return s / c; }

__device__ __forceinline__ float rk_fstable_tanh(float x) {   // stable_tanh without the branch
  const float ax = fminf(fabsf(x), 20.f);                     // tanh(20) == 1.f in fp32
  const float em = expm1f(2.f * ax);
  return copysignf(1.f, x) * ((x == 0.f) ? 0.f : rk_fdiv(em, em + 2.f));
}

__device__ __forceinline__ float sgnf(float x) { return (float)((x > 0.f) - (x < 0.f)); }

__device__ __forceinline__ float py_modf(float x, float y) {  // jnp.mod: sign of divisor
  float r = fmodf(x, y);
  if (r != 0.f && ((r < 0.f) != (y < 0.f))) r += y;
  return r;
}

__device__ __forceinline__ int ifloordiv(int a, int b) {
  if (b == 0) return 0;
  int qd = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --qd;
  return qd;
}

__device__ __forceinline__ int ipymod(int a, int b) {
  if (b == 0) return 0;
  int r = a % b;
  if (r != 0 && ((r < 0) != (b < 0))) r += b;
  return r;
}

__device__ __forceinline__ int ipow(int a, int b) {
  int r = 1;
  for (int i = 0; i < b; ++i) r *= a;
  return r;
}


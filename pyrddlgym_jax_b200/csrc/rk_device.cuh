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


// Hidden-layer activations of the deep reactive policy (planner.py:1128, 1358-1361: the reference
// takes any jax.nn function; these are the ones whose derivative is a function of the OUTPUT, so the
// adjoint needs nothing but the stored activation).
#pragma once
#include <cuda_runtime.h>

#define RK_ACT_TANH 0
#define RK_ACT_RELU 1
#define RK_ACT_SIGMOID 2
#define RK_ACT_SOFTPLUS 3
#define RK_ACT_ELU 4
#define RK_ACT_LEAKY_RELU 5
#define RK_ACT_COUNT 6


__device__ __forceinline__ float rk_act_fn(int act, float x) {
  switch (act) {
    case RK_ACT_RELU: return fmaxf(x, 0.f);
    case RK_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case RK_ACT_SOFTPLUS: return log1pf(expf(-fabsf(x))) + fmaxf(x, 0.f);
    case RK_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case RK_ACT_LEAKY_RELU: return x >= 0.f ? x : 0.01f * x;
    default: return tanhf(x);
  }
}

// d act / d x as a function of y = act(x)
__device__ __forceinline__ float rk_act_grad(int act, float y) {
  switch (act) {
    case RK_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case RK_ACT_SIGMOID: return y * (1.f - y);
    case RK_ACT_SOFTPLUS: return 1.f - expf(-y);
    case RK_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
    case RK_ACT_LEAKY_RELU: return y >= 0.f ? 1.f : 0.01f;
    default: return 1.f - y * y;
  }
}

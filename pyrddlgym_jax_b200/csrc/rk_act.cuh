// Hidden-layer activations of the deep reactive policy (planner.py:1128, 1358-1361: the reference
// takes any jax.nn function).  Ids 0-5 have a derivative that is a function of the OUTPUT, so the adjoint
// needs nothing but the stored activation (the fused per-epoch kernels are built for them).  gelu (jax's
// default tanh approximation) and silu are not monotone: their derivative is taken from the PRE-activation,
// which the layer-wise path keeps in a second tape (core/drp.py); for those ids the `aux` operand of an
// activation-adjoint epilogue is z, not act(z).  RK_ACT_IDENTITY writes the pre-activation itself.
#pragma once
#include <cuda_runtime.h>

#define RK_ACT_TANH 0
#define RK_ACT_RELU 1
#define RK_ACT_SIGMOID 2
#define RK_ACT_SOFTPLUS 3
#define RK_ACT_ELU 4
#define RK_ACT_LEAKY_RELU 5
#define RK_ACT_GELU 6
#define RK_ACT_SILU 7
#define RK_ACT_IDENTITY 8
#define RK_ACT_COUNT 9


__device__ __forceinline__ float rk_act_fn(int act, float x) {
  switch (act) {
    case RK_ACT_RELU: return fmaxf(x, 0.f);
    case RK_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case RK_ACT_SOFTPLUS: return log1pf(expf(-fabsf(x))) + fmaxf(x, 0.f);
    case RK_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case RK_ACT_LEAKY_RELU: return x >= 0.f ? x : 0.01f * x;
    case RK_ACT_GELU: return 0.5f * x * (1.f + tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x)));
    case RK_ACT_SILU: return x / (1.f + expf(-x));
    case RK_ACT_IDENTITY: return x;
    default: return tanhf(x);
  }
}

// d act / d x as a function of y = act(x) (ids 0-5) or of the pre-activation x itself (gelu, silu)
__device__ __forceinline__ float rk_act_grad(int act, float y) {
  switch (act) {
    case RK_ACT_GELU: {
      const float u = 0.7978845608028654f * (y + 0.044715f * y * y * y), t = tanhf(u);
      return 0.5f * (1.f + t) + 0.5f * y * (1.f - t * t) * 0.7978845608028654f * (1.f + 0.134145f * y * y);
    }
    case RK_ACT_SILU: {
      const float sg = 1.f / (1.f + expf(-y));
      return sg * (1.f + y * (1.f - sg));
    }
    case RK_ACT_IDENTITY: return 1.f;
    case RK_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case RK_ACT_SIGMOID: return y * (1.f - y);
    case RK_ACT_SOFTPLUS: return 1.f - expf(-y);
    case RK_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
    case RK_ACT_LEAKY_RELU: return y >= 0.f ? 1.f : 0.01f;
    default: return 1.f - y * y;
  }
}

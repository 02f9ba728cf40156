// rk_interp.cu -- the rollout interpreter kernel (sm_100a) and its C ABI.
//
// One warp owns RPW rollouts; lane = q*L + e walks the object tuples (e) of rollout q.
// The warp executes the program's instruction stream (core/isa.py) over a register file
// held in shared memory: a value of n elements per rollout is the word range [q][e], so
//   * elementwise instructions are bank-conflict free,
//   * global loads/stores of fluent-major buffers are flat, fully coalesced copies,
//   * aggregations (sum/prod/min/max/forall/exists/argmax, compiler.py:1303-1348) are CSR
//     segment loops over static tables kept in a CTA-wide constant region.
// The same kernel runs forward programs (t ascending) and adjoint programs (t descending).
//
// Replaces the XLA executables behind compile_rollouts (compiler.py:621-666) and the
// reverse pass of jax.value_and_grad (planner.py:2873).

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/rk.h"
#include "rk_isa.h"

#include "rk_launch.h"

struct Opnd {
  const uint32_t* p;
  uint32_t map;
};

__device__ __forceinline__ Opnd mk_opnd(uint32_t base, uint32_t sq, uint32_t map, uint32_t q,
                                        const uint32_t* R, const uint32_t* C) {
  Opnd o;
  o.p = ((base & RK_CONST_BIT) ? (C + (base & 0x7fffu)) : (R + base)) + q * sq;
  o.map = map;
  return o;
}

__device__ __forceinline__ uint32_t ldw(const Opnd& o, uint32_t e, const uint32_t* C) {
  uint32_t k = (o.map == 0u) ? e : ((o.map == 1u) ? 0u : C[o.map + e]);
  return o.p[k];
}

__device__ __forceinline__ float ldf(const Opnd& o, uint32_t e, const uint32_t* C) {
  return __uint_as_float(ldw(o, e, C));
}

#include "rk_device.cuh"

#define F2U(x) __float_as_uint(x)

// elementwise loops: lanes with e0 >= n do nothing
#define EW1F(expr)                                                        \
  for (uint32_t e = e0; e < n; e += L) {                                  \
    float x = ldf(A, e, C);                                               \
    D[e] = F2U(expr);                                                     \
  }                                                                       \
  break;
#define EW2F(expr)                                                        \
  for (uint32_t e = e0; e < n; e += L) {                                  \
    float x = ldf(A, e, C);                                               \
    float y = ldf(Bo, e, C);                                              \
    D[e] = F2U(expr);                                                     \
  }                                                                       \
  break;
#define EW2FB(expr)                                                       \
  for (uint32_t e = e0; e < n; e += L) {                                  \
    float x = ldf(A, e, C);                                               \
    float y = ldf(Bo, e, C);                                              \
    D[e] = (expr) ? 1u : 0u;                                              \
  }                                                                       \
  break;
#define EW1I(expr)                                                        \
  for (uint32_t e = e0; e < n; e += L) {                                  \
    int x = (int)ldw(A, e, C);                                            \
    D[e] = (uint32_t)(expr);                                              \
  }                                                                       \
  break;
#define EW2I(expr)                                                        \
  for (uint32_t e = e0; e < n; e += L) {                                  \
    int x = (int)ldw(A, e, C);                                            \
    int y = (int)ldw(Bo, e, C);                                           \
    D[e] = (uint32_t)(expr);                                              \
  }                                                                       \
  break;
#define RED(init, TYPE, LOAD, STORE, update)                              \
  {                                                                       \
    const uint32_t* ptr = C + aux0;                                       \
    const uint32_t* idx = C + aux1;                                       \
    for (uint32_t e = e0; e < n; e += L) {                                \
      TYPE acc = init;                                                    \
      uint32_t j1 = ptr[e + 1];                                           \
      for (uint32_t j = ptr[e]; j < j1; ++j) {                            \
        TYPE v = LOAD(A.p[idx[j]]);                                       \
        update;                                                           \
      }                                                                   \
      D[e] = STORE(acc);                                                  \
    }                                                                     \
  }                                                                       \
  break;
#define AS_INT(x) ((int)(x))
#define AS_U32(x) ((uint32_t)(x))

__device__ __forceinline__ void run_list(const RkLaunch& K, const uint4* __restrict__ ins,
                                         int count, long long t, uint32_t* R,
                                         const uint32_t* C, uint32_t lane, uint32_t q,
                                         uint32_t le, uint32_t nvalid, long long b0,
                                         long long gwarp) {
  const uint32_t L = (uint32_t)K.lanes;
  const bool lane_ok = q < (uint32_t)K.rpw;
  uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
  if (count > 0) {
    w0 = __ldg(ins);
    w1 = __ldg(ins + 1);
  }
  for (int pc = 0; pc < count; ++pc) {
    uint4 n0 = w0, n1 = w1;
    if (pc + 1 < count) {                       // prefetch the next instruction
      n0 = __ldg(ins + 2 * (pc + 1));
      n1 = __ldg(ins + 2 * (pc + 1) + 1);
    }
    const uint32_t op = w0.x & 0xffffu, flags = w0.x >> 16;
    const uint32_t n = w0.y & 0xffffu, dst = w0.y >> 16;
    const uint32_t dst_sq = w0.z & 0xffffu;
    const uint32_t a = w0.z >> 16, a_sq = w0.w & 0xffffu, a_map = w0.w >> 16;
    const uint32_t b = w1.x & 0xffffu, b_sq = w1.x >> 16, b_map = w1.y & 0xffffu;
    const uint32_t c = w1.y >> 16, c_sq = w1.z & 0xffffu, c_map = w1.z >> 16;
    const uint32_t aux0 = w1.w & 0xffffu, aux1 = w1.w >> 16;

    const bool uni = (flags & RK_F_UNIFORM) != 0;
    const uint32_t e0 = (lane_ok && (!uni || q == 0u)) ? le : n;   // n => lane idle
    uint32_t* D = R + dst + q * dst_sq;
    const Opnd A = mk_opnd(a, a_sq, a_map, q, R, C);
    const Opnd Bo = mk_opnd(b, b_sq, b_map, q, R, C);
    const Opnd Co = mk_opnd(c, c_sq, c_map, q, R, C);

    switch (op) {
      case OP_NOP: break;
      case OP_MOV:
        for (uint32_t e = e0; e < n; e += L) D[e] = ldw(A, e, C);
        break;
      case OP_NEG: EW1F(-x)
      case OP_ABS: EW1F(fabsf(x))
      case OP_SGN: EW1F(sgnf(x))
      case OP_FLOOR: EW1F(floorf(x))
      case OP_CEIL: EW1F(ceilf(x))
      case OP_ROUND: EW1F(rintf(x))
      case OP_SQRT: EW1F(sqrtf(x))
      case OP_EXP: EW1F(expf(x))
      case OP_LOG: EW1F(logf(x))
      case OP_SIN: EW1F(sinf(x))
      case OP_COS: EW1F(cosf(x))
      case OP_TAN: EW1F(tanf(x))
      case OP_ASIN: EW1F(asinf(x))
      case OP_ACOS: EW1F(acosf(x))
      case OP_ATAN: EW1F(atanf(x))
      case OP_SINH: EW1F(sinhf(x))
      case OP_COSH: EW1F(coshf(x))
      case OP_TANH: EW1F(tanhf(x))
      case OP_LGAMMA: EW1F(lgammaf(x))
      case OP_SIGMOID: EW1F(1.f / (1.f + expf(-x)))
      case OP_STANH: EW1F(stable_tanh(x))
      case OP_RELU: EW1F(fmaxf(x, 0.f))
      case OP_RCP: EW1F(1.f / x)
      case OP_SQUARE: EW1F(x * x)
      case OP_LOG1P: EW1F(log1pf(x))
      case OP_EXPM1: EW1F(expm1f(x))
      case OP_DIGAMMA: EW1F(rk_digamma(x))
      case OP_ADD: EW2F(x + y)
      case OP_SUB: EW2F(x - y)
      case OP_MUL: EW2F(x * y)
      case OP_DIV: EW2F(x / y)
      case OP_MIN: EW2F(fminf(x, y))
      case OP_MAX: EW2F(fmaxf(x, y))
      case OP_POW: EW2F(powf(x, y))
      case OP_HYPOT: EW2F(hypotf(x, y))
      case OP_FMOD: EW2F(fmodf(x, y))
      case OP_MOD: EW2F(py_modf(x, y))
      case OP_FDIV: EW2F(floorf(x / y))
      case OP_SAFELOG: EW2F((x > 0.f) ? logf(x) : -CUDART_INF_F + 0.f * y)
      case OP_LT: EW2FB(x < y)
      case OP_LE: EW2FB(x <= y)
      case OP_GT: EW2FB(x > y)
      case OP_GE: EW2FB(x >= y)
      case OP_EQ: EW2FB(x == y)
      case OP_NE: EW2FB(x != y)
      case OP_MINW: EW2F((x < y) ? 1.f : ((x == y) ? 0.5f : 0.f))
      case OP_MAXW: EW2F((x > y) ? 1.f : ((x == y) ? 0.5f : 0.f))
      case OP_SELECT:
        for (uint32_t e = e0; e < n; e += L)
          D[e] = (ldw(A, e, C) != 0u) ? ldw(Bo, e, C) : ldw(Co, e, C);
        break;
      case OP_FMA:
        for (uint32_t e = e0; e < n; e += L)
          D[e] = F2U(__fadd_rn(__fmul_rn(ldf(A, e, C), ldf(Bo, e, C)), ldf(Co, e, C)));
        break;
      case OP_LERP:
        for (uint32_t e = e0; e < n; e += L) {
          float x = ldf(A, e, C);
          D[e] = F2U(__fadd_rn(__fmul_rn(x, ldf(Bo, e, C)),
                               __fmul_rn(__fsub_rn(1.f, x), ldf(Co, e, C))));
        }
        break;
      case OP_INEG: EW1I(-x)
      case OP_IABS: EW1I(abs(x))
      case OP_ISGN: EW1I((x > 0) - (x < 0))
      case OP_IADD: EW2I(x + y)
      case OP_ISUB: EW2I(x - y)
      case OP_IMUL: EW2I(x * y)
      case OP_IMIN: EW2I(min(x, y))
      case OP_IMAX: EW2I(max(x, y))
      case OP_IFDIV: EW2I(ifloordiv(x, y))
      case OP_IMOD: EW2I(ipymod(x, y))
      case OP_IFMOD: EW2I((y == 0) ? 0 : (x % y))
      case OP_IPOW: EW2I(ipow(x, y))
      case OP_ILT: EW2I(x < y)
      case OP_ILE: EW2I(x <= y)
      case OP_IGT: EW2I(x > y)
      case OP_IGE: EW2I(x >= y)
      case OP_IEQ: EW2I(x == y)
      case OP_INE: EW2I(x != y)
      case OP_AND: EW2I(x & y)
      case OP_OR: EW2I(x | y)
      case OP_XOR: EW2I(x ^ y)
      case OP_NOT: EW1I(x == 0)
      case OP_I2F:
        for (uint32_t e = e0; e < n; e += L) D[e] = F2U((float)(int)ldw(A, e, C));
        break;
      case OP_F2I:
        for (uint32_t e = e0; e < n; e += L) D[e] = (uint32_t)(int)ldf(A, e, C);
        break;
      case OP_F2B:
        for (uint32_t e = e0; e < n; e += L) D[e] = (ldf(A, e, C) != 0.f) ? 1u : 0u;
        break;
      case OP_I2B: EW1I(x != 0)
      case OP_RSUM: RED(0.f, float, __uint_as_float, F2U, acc += v)
      case OP_RPROD: RED(1.f, float, __uint_as_float, F2U, acc *= v)
      case OP_RMIN: RED(CUDART_INF_F, float, __uint_as_float, F2U, acc = fminf(acc, v))
      case OP_RMAX: RED(-CUDART_INF_F, float, __uint_as_float, F2U, acc = fmaxf(acc, v))
      case OP_RISUM: RED(0, int, AS_INT, AS_U32, acc += v)
      case OP_RIPROD: RED(1, int, AS_INT, AS_U32, acc *= v)
      case OP_RIMIN: RED(2147483647, int, AS_INT, AS_U32, acc = min(acc, v))
      case OP_RIMAX: RED((-2147483647 - 1), int, AS_INT, AS_U32, acc = max(acc, v))
      case OP_RARGMAX:
      case OP_RARGMIN: {
        const uint32_t* ptr = C + aux0;
        const uint32_t* idx = C + aux1;
        const float sgn = (op == OP_RARGMAX) ? 1.f : -1.f;
        for (uint32_t e = e0; e < n; e += L) {
          uint32_t j0 = ptr[e], j1 = ptr[e + 1];
          float best = -CUDART_INF_F;
          int arg = 0;
          for (uint32_t j = j0; j < j1; ++j) {
            float v = sgn * __uint_as_float(A.p[idx[j]]);
            if (v > best) { best = v; arg = (int)(j - j0); }
          }
          D[e] = (uint32_t)arg;
        }
      } break;
      case OP_RPRODX: {   // product of the other members of the element's segment
        const uint32_t* ptr = C + aux0;
        const uint32_t* idx = C + aux1;
        const uint32_t* seg = C + b;
        for (uint32_t e = e0; e < n; e += L) {
          uint32_t s = seg[e];
          float acc = 1.f;
          uint32_t j1 = ptr[s + 1];
          for (uint32_t j = ptr[s]; j < j1; ++j) {
            uint32_t k = idx[j];
            if (k != e) acc *= __uint_as_float(A.p[k]);
          }
          D[e] = F2U(acc);
        }
      } break;
      case OP_QSUM: {     // uniform dst[e] = sum over the warp's valid rollouts
        const uint32_t* src = R + a;
        if (lane_ok && q == 0u) {
          for (uint32_t e = le; e < n; e += L) {
            float acc = 0.f;
            for (uint32_t qq = 0; qq < nvalid; ++qq) acc += __uint_as_float(src[qq * a_sq + e]);
            D[e] = F2U(acc);
          }
        }
      } break;
      case OP_LDB: {      // flat coalesced copy: registers [q][e] <- buffer [b][e]
        const uint32_t off = aux1 | (c << 16);
        const uint32_t* g = (const uint32_t*)K.buf[aux0];
        if (g != nullptr) {
          long long tt = (flags & RK_F_TSTEP) ? t : 0;
          const uint32_t* src = g + (tt * K.W[aux0] + (long long)off) * (long long)K.B + b0 * (long long)n;
          uint32_t* d = R + dst;
          for (uint32_t i = lane; i < nvalid * n; i += 32u) d[i] = src[i];
        }
      } break;
      case OP_STB: {
        const uint32_t off = aux1 | (c << 16);
        uint32_t* g = (uint32_t*)K.buf[aux0];
        if (g != nullptr) {
          long long tt = (flags & RK_F_TSTEP) ? t : 0;
          uint32_t* dstp = g + (tt * K.W[aux0] + (long long)off) * (long long)K.B + b0 * (long long)n;
          if (a_map == 0u && a_sq == n && !(a & RK_CONST_BIT)) {
            const uint32_t* s = R + a;
            for (uint32_t i = lane; i < nvalid * n; i += 32u) dstp[i] = s[i];
          } else if (lane_ok && q < nvalid) {
            for (uint32_t e = le; e < n; e += L) dstp[q * n + e] = ldw(A, e, C);
          }
        }
      } break;
      case OP_LDU: {
        const uint32_t off = aux1 | (c << 16);
        const uint32_t* g = (const uint32_t*)K.buf[aux0];
        if (g != nullptr) {
          long long tt = (flags & RK_F_TSTEP) ? t : 0;
          const uint32_t* src = g + tt * K.W[aux0] + off;
          uint32_t* d = R + dst;
          for (uint32_t i = lane; i < n; i += 32u) d[i] = __ldg(src + i);
        }
      } break;
      case OP_STG: {      // per-warp gradient partial row
        const uint32_t off = aux1 | (c << 16);
        uint32_t* g = (uint32_t*)K.buf[aux0];
        if (g != nullptr) {
          uint32_t* dstp = g + (gwarp * (long long)K.T + t) * K.W[aux0] + off;
          const Opnd A0 = mk_opnd(a, 0u, a_map, 0u, R, C);
          for (uint32_t i = lane; i < n; i += 32u) dstp[i] = ldw(A0, i, C);
        }
      } break;
      case OP_ERRCHK: {
        bool bad = false;
        if (lane_ok && q < nvalid)
          for (uint32_t e = le; e < n; e += L) bad |= (ldw(A, e, C) != 0u);
        if (bad) atomicOr(R + K.err_base + q, 1u << aux0);
      } break;
      case OP_STERR: {
        uint32_t* g = (uint32_t*)K.buf[aux0];
        if (lane < nvalid) {
          if (g != nullptr) g[t * (long long)K.B + b0 + lane] = R[K.err_base + lane] | (uint32_t)K.static_err;
          R[K.err_base + lane] = 0u;
        }
      } break;
      default: break;
    }
    __syncwarp();
    w0 = n0;
    w1 = n1;
  }
}

extern "C" __global__ void __launch_bounds__(256)
rk_interp_kernel(const RkLaunch K) {
  extern __shared__ __align__(16) uint32_t smem[];
  const uint32_t* prog = K.prog;
  for (int i = threadIdx.x; i < K.const_words; i += blockDim.x) smem[i] = prog[RK_HEADER_WORDS + i];
  __syncthreads();
  if (K.mparams != nullptr)
    for (int i = threadIdx.x; i < K.n_mparams; i += blockDim.x)
      smem[K.mparam_off + i] = __float_as_uint(K.mparams[i]);
  __syncthreads();

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const long long gwarp = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const long long b0 = gwarp * K.rpw;
  if (b0 >= K.B) return;
  const uint32_t nvalid = (uint32_t)min((long long)K.rpw, (long long)K.B - b0);
  const uint32_t q = lane / (uint32_t)K.lanes, le = lane - q * (uint32_t)K.lanes;
  const uint32_t* C = smem;
  uint32_t* R = smem + K.const_words + warp * K.warp_words;
  for (uint32_t i = lane; i < (uint32_t)K.rpw; i += 32u) R[K.err_base + i] = 0u;
  __syncwarp();

  const uint4* ins = (const uint4*)(prog + RK_HEADER_WORDS + K.const_words);
  run_list(K, ins, K.n_pro, 0, R, C, lane, q, le, nvalid, b0, gwarp);
  const uint4* step = ins + 2 * K.n_pro;
  if (K.t_dir > 0) {
    for (int t = 0; t < K.T; ++t) run_list(K, step, K.n_step, t, R, C, lane, q, le, nvalid, b0, gwarp);
  } else {
    for (int t = K.T - 1; t >= 0; --t) run_list(K, step, K.n_step, t, R, C, lane, q, le, nvalid, b0, gwarp);
  }
  run_list(K, step + 2 * K.n_step, K.n_epi, 0, R, C, lane, q, le, nvalid, b0, gwarp);
}

// =====================================================================================
// host side: program handle + launch wrappers (the C ABI of include/rk.h)
// =====================================================================================

struct rk_program {
  uint32_t* d_blob;
  size_t nbytes;
  uint32_t hdr[RK_HEADER_WORDS];
  int warps_per_cta;
  int smem_bytes;
  int device;
  void* module;     // CUmodule of the specialised kernel (core/codegen.py), or NULL
  void* func;       // CUfunction
  int spec_threads; // launch geometry of the attached kernel: threads per CTA (0 = warps_per_cta * 32)
  int spec_smem;    //   dynamic shared memory in bytes (-1 = smem_bytes of the interpreter layout)
};

// ---- CUDA driver API, resolved lazily so that the library also loads on hosts without a
// driver (CPU-only CI boxes only check the exported symbols) -----------------------------
#include <dlfcn.h>
typedef int (*cuModuleLoadData_t)(void**, const void*);
typedef int (*cuModuleGetFunction_t)(void**, void*, const char*);
typedef int (*cuModuleUnload_t)(void*);
typedef int (*cuFuncSetAttribute_t)(void*, int, int);
typedef int (*cuLaunchKernel_t)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                unsigned, void*, void**, void**);
static struct {
  void* lib;
  cuModuleLoadData_t moduleLoadData;
  cuModuleGetFunction_t moduleGetFunction;
  cuModuleUnload_t moduleUnload;
  cuFuncSetAttribute_t funcSetAttribute;
  cuLaunchKernel_t launchKernel;
} g_drv = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

static int load_driver() {
  if (g_drv.lib != nullptr) return 0;
  void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (lib == nullptr) lib = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
  if (lib == nullptr) return RK_E_BADARG;
  g_drv.moduleLoadData = (cuModuleLoadData_t)dlsym(lib, "cuModuleLoadData");
  g_drv.moduleGetFunction = (cuModuleGetFunction_t)dlsym(lib, "cuModuleGetFunction");
  g_drv.moduleUnload = (cuModuleUnload_t)dlsym(lib, "cuModuleUnload");
  g_drv.funcSetAttribute = (cuFuncSetAttribute_t)dlsym(lib, "cuFuncSetAttribute");
  g_drv.launchKernel = (cuLaunchKernel_t)dlsym(lib, "cuLaunchKernel");
  if (!g_drv.moduleLoadData || !g_drv.moduleGetFunction || !g_drv.moduleUnload ||
      !g_drv.funcSetAttribute || !g_drv.launchKernel) return RK_E_BADARG;
  g_drv.lib = lib;
  return 0;
}

extern "C" int rk_program_attach_cubin_ex(rk_program* p, const void* image, size_t nbytes,
                                          const char* entry, int threads_per_cta, int smem_bytes,
                                          int carveout);

extern "C" int rk_program_attach_cubin(rk_program* p, const void* image, size_t nbytes,
                                       const char* entry) {
  const char* co = getenv("RK_CARVEOUT");
  return rk_program_attach_cubin_ex(p, image, nbytes, entry, 0, -1, co ? atoi(co) : 100);
}

extern "C" int rk_program_attach_cubin_ex(rk_program* p, const void* image, size_t nbytes,
                                          const char* entry, int threads_per_cta, int smem_bytes,
                                          int carveout) {
  if (p == nullptr || image == nullptr || nbytes == 0 || entry == nullptr) return RK_E_BADARG;
  if (threads_per_cta < 0 || threads_per_cta > 1024 || (threads_per_cta & 31) != 0) return RK_E_BADARG;
  if (smem_bytes < -1 || smem_bytes > 227 * 1024) return RK_E_BADARG;
  int rc = load_driver();
  if (rc != 0) return rc;
  cudaFree(0);                                   // make sure the primary context is current
  void* mod = nullptr;
  void* fn = nullptr;
  rc = g_drv.moduleLoadData(&mod, image);
  if (rc != 0) return rc;
  rc = g_drv.moduleGetFunction(&fn, mod, entry);
  if (rc != 0) { g_drv.moduleUnload(mod); return rc; }
  // CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES = 8
  rc = g_drv.funcSetAttribute(fn, 8, smem_bytes >= 0 ? smem_bytes : p->smem_bytes);
  if (rc != 0) { g_drv.moduleUnload(mod); return rc; }
  // CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT = 9: every rollout kernel asks for the
  // same L1 / shared split, so the exact test kernel and the train kernels of the pipelined
  // iteration can share an SM (kernels with different splits cannot co-reside)
  if (carveout >= 0) g_drv.funcSetAttribute(fn, 9, carveout > 100 ? 100 : carveout);
  if (p->module != nullptr) g_drv.moduleUnload(p->module);
  p->module = mod;
  p->func = fn;
  p->spec_threads = threads_per_cta;
  p->spec_smem = smem_bytes;
  return 0;
}

static int pick_geometry(rk_program* p) {
  const int const_words = (int)p->hdr[5], warp_words = (int)p->hdr[8];
  int dev = 0;
  cudaGetDevice(&dev);
  int max_optin = 0, sm_smem = 0;
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  if (max_optin <= 0) max_optin = 227 * 1024;
  if (sm_smem <= 0) sm_smem = 228 * 1024;
  // small CTAs keep the tail short; 4 warps amortise the constant region
  int wpc = 4;
  while (wpc > 1 && (size_t)(const_words + wpc * warp_words) * 4 > (size_t)max_optin) wpc >>= 1;
  size_t bytes = (size_t)(const_words + wpc * warp_words) * 4;
  if (bytes > (size_t)max_optin) return RK_E_SMEM;
  p->warps_per_cta = wpc;
  p->smem_bytes = (int)bytes;
  p->device = dev;
  return 0;
}

extern "C" int rk_program_create(const void* blob, size_t nbytes, rk_program** out) {
  if (blob == nullptr || out == nullptr || nbytes < RK_HEADER_WORDS * 4) return RK_E_BADARG;
  rk_program* p = (rk_program*)calloc(1, sizeof(rk_program));
  if (p == nullptr) return RK_E_NOMEM;
  memcpy(p->hdr, blob, RK_HEADER_WORDS * 4);
  if (p->hdr[0] != RK_MAGIC) { free(p); return RK_E_BADPROG; }
  const size_t need = ((size_t)RK_HEADER_WORDS + p->hdr[5]) * 4 +
                      ((size_t)p->hdr[2] + p->hdr[3] + p->hdr[4]) * RK_INS_WORDS * 2;
  if (nbytes < need) { free(p); return RK_E_BADPROG; }
  int rc = pick_geometry(p);
  if (rc != 0) { free(p); return rc; }
  cudaError_t ce = cudaMalloc((void**)&p->d_blob, nbytes);
  if (ce != cudaSuccess) { free(p); return (int)ce; }
  ce = cudaMemcpy(p->d_blob, blob, nbytes, cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) { cudaFree(p->d_blob); free(p); return (int)ce; }
  p->nbytes = nbytes;
  ce = cudaFuncSetAttribute(rk_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            227 * 1024);
  if (ce != cudaSuccess) { cudaFree(p->d_blob); free(p); return (int)ce; }
  *out = p;
  return 0;
}

extern "C" void rk_program_destroy(rk_program* p) {
  if (p == nullptr) return;
  if (p->module != nullptr && g_drv.moduleUnload != nullptr) g_drv.moduleUnload(p->module);
  if (p->d_blob) cudaFree(p->d_blob);
  free(p);
}

extern "C" int rk_program_query(const rk_program* p, rk_sizes* o) {
  if (p == nullptr || o == nullptr) return RK_E_BADARG;
  o->kind = (p->hdr[14] & 2u) ? 1 : 0;
  o->relaxed = (int)(p->hdr[14] & 1u);
  o->S = (int)p->hdr[11];
  o->A = (int)p->hdr[12];
  o->N = (int)p->hdr[13];
  o->n_mparams = (int)p->hdr[7];
  o->rpw = (int)p->hdr[9];
  o->lanes = (int)p->hdr[10];
  o->warp_words = (int)p->hdr[8];
  o->const_words = (int)p->hdr[5];
  o->n_prologue = (int)p->hdr[2];
  o->n_step = (int)p->hdr[3];
  o->n_epilogue = (int)p->hdr[4];
  o->warps_per_cta = p->warps_per_cta;
  o->smem_bytes = p->smem_bytes;
  o->static_err = (int)p->hdr[15];
  o->tape_words = p->hdr[18] ? (int)p->hdr[18] : (int)p->hdr[11];
  return 0;
}

extern "C" int rk_program_num_warps(const rk_program* p, int B) {
  if (p == nullptr || B < 0) return RK_E_BADARG;
  const int rpw = (int)p->hdr[9];
  return (B + rpw - 1) / rpw;
}

static int launch(const rk_program* p, cudaStream_t st, RkLaunch& K) {
  K.prog = p->d_blob;
  K.const_words = (int)p->hdr[5];
  K.warp_words = (int)p->hdr[8];
  K.rpw = (int)p->hdr[9];
  K.lanes = (int)p->hdr[10];
  K.n_pro = (int)p->hdr[2];
  K.n_step = (int)p->hdr[3];
  K.n_epi = (int)p->hdr[4];
  K.mparam_off = (int)p->hdr[6];
  K.n_mparams = (int)p->hdr[7];
  K.err_base = (int)p->hdr[16];
  K.static_err = (int)p->hdr[15];
  if (K.B == 0) return 0;
  const int nwarps = (K.B + K.rpw - 1) / K.rpw;
  const int grid = (nwarps + p->warps_per_cta - 1) / p->warps_per_cta;
  if (p->func != nullptr) {                      // per-program specialised kernel
    // generated kernels index one step's [W][B] block with 32-bit word offsets
    for (int i = 0; i < RK_NUM_BUFFERS; ++i)
      if (K.W[i] * (long long)K.B >= (1ll << 31)) return RK_E_BADARG;
    void* args[] = {(void*)&K};
    int threads = p->spec_threads > 0 ? p->spec_threads : p->warps_per_cta * 32;
    if (p->spec_threads > 0 && p->spec_smem == 0) {
      // register-only (thread-per-rollout) kernels index rollouts by blockDim: pick the CTA width that
      // leaves the fewest warps on the busiest SM (1024 warps as 256 CTAs of 4 put 8 warps on 108 SMs and
      // 4 on the rest; as 1024 CTAs of 1 they put 7 on every SM)
      static int n_sm = 0;
      if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0)
          n_sm = 148;
      }
      long long best = -1;
      for (int w = p->spec_threads / 32; w >= 1; w >>= 1) {
        const long long ctas = (nwarps + w - 1) / w;
        const long long busiest = ((ctas + n_sm - 1) / n_sm) * w;
        if (best < 0 || busiest < best) {
          best = busiest;
          threads = w * 32;
        }
      }
    }
    const int sgrid = (nwarps + threads / 32 - 1) / (threads / 32);
    const int smem = p->spec_smem >= 0 ? p->spec_smem : p->smem_bytes;
    return g_drv.launchKernel(p->func, (unsigned)sgrid, 1, 1, (unsigned)threads, 1, 1,
                              (unsigned)smem, (void*)st, args, nullptr);
  }
  rk_interp_kernel<<<grid, p->warps_per_cta * 32, p->smem_bytes, st>>>(K);
  return (int)cudaGetLastError();
}

extern "C" int rk_rollout_forward(const rk_program* p, void* stream, int B, int T,
                                  const void* s0, const float* policy, const void* actions,
                                  const float* noise, const float* mparams, const float* disc,
                                  float* reward, float* returns, uint32_t* err,
                                  void* tape, void* final_state, void* traj) {
  if (p == nullptr || B < 0 || T < 0) return RK_E_BADARG;
  if (p->hdr[14] & 2u) return RK_E_BADPROG;
  if (s0 == nullptr && B > 0) return RK_E_BADARG;
  if (p->hdr[13] > 0 && noise == nullptr && B > 0 && T > 0) return RK_E_BADARG;
  RkLaunch K;
  memset(&K, 0, sizeof(K));
  K.B = B; K.T = T; K.t_dir = 1; K.mparams = mparams;
  const long long S = p->hdr[11], A = p->hdr[12], N = p->hdr[13];
  K.buf[RK_BUF_S0] = (void*)s0;           K.W[RK_BUF_S0] = S;
  K.buf[RK_BUF_TAPE] = tape;              K.W[RK_BUF_TAPE] = p->hdr[18] ? p->hdr[18] : S;
  K.buf[RK_BUF_NOISE] = (void*)noise;     K.W[RK_BUF_NOISE] = N;
  K.buf[RK_BUF_PARAMS] = (void*)policy;   K.W[RK_BUF_PARAMS] = A;
  K.buf[RK_BUF_ACT] = (void*)actions;     K.W[RK_BUF_ACT] = A;
  K.buf[RK_BUF_REWARD] = reward;          K.W[RK_BUF_REWARD] = 1;
  K.buf[RK_BUF_RETURNS] = returns;        K.W[RK_BUF_RETURNS] = 1;
  K.buf[RK_BUF_ERR] = err;                K.W[RK_BUF_ERR] = 1;
  K.buf[RK_BUF_DISC] = (void*)disc;       K.W[RK_BUF_DISC] = 1;
  K.buf[RK_BUF_FINAL] = final_state;      K.W[RK_BUF_FINAL] = S;
  K.buf[RK_BUF_TRAJ] = traj;              K.W[RK_BUF_TRAJ] = p->hdr[17];
  return launch(p, (cudaStream_t)stream, K);
}

extern "C" int rk_rollout_backward(const rk_program* p, void* stream, int B, int T,
                                   const void* tape, const float* policy, const void* actions,
                                   const float* noise, const float* mparams, const float* disc,
                                   const float* gret, float* gradp, float* gact, float* gs0,
                                   const float* gs_in) {
  if (p == nullptr || B < 0 || T < 0) return RK_E_BADARG;
  if (!(p->hdr[14] & 2u)) return RK_E_BADPROG;
  if ((tape == nullptr || gret == nullptr) && B > 0) return RK_E_BADARG;
  if (p->hdr[13] > 0 && noise == nullptr && B > 0 && T > 0) return RK_E_BADARG;
  RkLaunch K;
  memset(&K, 0, sizeof(K));
  K.B = B; K.T = T; K.t_dir = -1; K.mparams = mparams;
  const long long S = p->hdr[11], A = p->hdr[12], N = p->hdr[13];
  K.buf[RK_BUF_TAPE] = (void*)tape;       K.W[RK_BUF_TAPE] = p->hdr[18] ? p->hdr[18] : S;
  K.buf[RK_BUF_NOISE] = (void*)noise;     K.W[RK_BUF_NOISE] = N;
  K.buf[RK_BUF_PARAMS] = (void*)policy;   K.W[RK_BUF_PARAMS] = A;
  K.buf[RK_BUF_ACT] = (void*)actions;     K.W[RK_BUF_ACT] = A;
  K.buf[RK_BUF_DISC] = (void*)disc;       K.W[RK_BUF_DISC] = 1;
  K.buf[RK_BUF_GRET] = (void*)gret;       K.W[RK_BUF_GRET] = 1;
  K.buf[RK_BUF_GRADP] = gradp;            K.W[RK_BUF_GRADP] = A;
  K.buf[RK_BUF_GACT] = gact;              K.W[RK_BUF_GACT] = A;
  K.buf[RK_BUF_GS0] = gs0;                K.W[RK_BUF_GS0] = S;
  K.buf[RK_BUF_GSIN] = (void*)gs_in;      K.W[RK_BUF_GSIN] = S;
  return launch(p, (cudaStream_t)stream, K);
}

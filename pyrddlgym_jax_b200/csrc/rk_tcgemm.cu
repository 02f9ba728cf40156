// rk_tcgemm.cu -- the hidden dense layers of the deep reactive policy on the 5th-generation
// tensor cores (tcgen05 + TMEM + TMA), fp32-grade accuracy through a 3xTF32 split.
//
// Replaces: flax nn.Dense + tanh of the DRP hidden stack (planner.py:1358-1361) and its
// input-side transpose in the reverse pass -- the only dense contraction on the path.
//
//   C[M][N] = epi( A[M][K] * Bt[N][K]^T ),   A = A_hi + A_lo,  Bt = Bt_hi + Bt_lo
//   A*B ~= A_hi*B_hi + A_lo*B_hi + A_hi*B_lo      (the lo*lo term is below fp32 rounding)
//
// hi parts are exactly representable in TF32 (cvt.rna), lo parts are the fp32 remainders, so the tensor core's fp32->tf32 operand conversion is exact for hi and costs
// < 2^-22 relative for lo: the sum matches an fp32 GEMM to ~1e-6.
//
// One CTA per 128-row tile, full N (<= 256) in one TMEM accumulator (N fp32 columns):
//   warp 0   : TMA producer  (3 tiles per k-block: raw A, B_hi, B_lo; 64B-swizzled, K-major)
//   warp 1   : MMA issuer    (one elected lane; 3 tcgen05.mma.kind::tf32 per 8-wide k-slice)
//   warps 2-9: during the main loop split the raw A tile into hi (in place) / lo (second buffer)
//              -- an elementwise map, so the swizzled layout does not matter -- then run the
//              epilogue (tcgen05.ld 32 lanes x 32 columns -> bias/tanh or tanh-adjoint -> C)
// The activation matrix is therefore read from HBM exactly once, as plain fp32; only the small
// weight matrix is pre-split (rk_tf32_split, once per update).
// 4-stage smem ring (48 KB per stage at N=256), mbarriers full -> split -> (tcgen05.commit) empty.

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

#include "../../include/rk.h"
#include "rk_act.cuh"

#define TC_BM 128
#define TC_BK 16          // fp32 elements per k-block = one 64-byte swizzle row
#define TC_STAGES 4       // 48 KB per stage at N = 256: three loads in flight behind the MMA
#define TC_UMMA_K 8       // kind::tf32: 32 bytes of K per instruction

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 64B-swizzled operand tile (rows of 16 fp32 = 64 B, 8-row groups 512 B apart):
// start address, SBO = 512 B, version 1 (sm_100), layout type 4 (SWIZZLE_64B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((8 * TC_BK * 4) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

// epi: 2 = bias + tanh, 3 = * (1 - aux^2), 1 = bias, 0 = none
#define TC_THREADS 320     // producer warp, MMA warp, 8 converter / epilogue warps
// ACT_CT: the activation as a compile-time constant (the tanh build keeps a straight-line
// epilogue), or -1 = the run-time value `act_rt`
template <int ACT_CT>
__global__ void __launch_bounds__(TC_THREADS, 1)
rk_tcgemm_kernel(const __grid_constant__ CUtensorMap tmA,
                 const __grid_constant__ CUtensorMap tmB_hi,
                 const __grid_constant__ CUtensorMap tmB_lo, int M, int N, int K, int epi,
                 float* __restrict__ C, int ldc, const float* __restrict__ bias,
                 const float* __restrict__ aux, int ldaux, uint32_t tmem_cols, int act_rt) {
  const int act = ACT_CT >= 0 ? ACT_CT : act_rt;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = TC_BM * TC_BK * 4;            // 8 KB
  const uint32_t b_bytes = (uint32_t)N * TC_BK * 4;      // 16 KB at N = 256
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  uint64_t* bars = (uint64_t*)(smem + TC_STAGES * stage_bytes);
  uint64_t* full = bars;                  // [TC_STAGES]
  uint64_t* empty = bars + TC_STAGES;     // [TC_STAGES]
  uint64_t* split_done = bars + 2 * TC_STAGES;     // [TC_STAGES], one arrival per converter warp
  uint64_t* acc_ready = bars + 3 * TC_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * TC_STAGES + 1);
  float* bias_s = (float*)(bars + 3 * TC_STAGES + 2);   // [N] when the epilogue adds a bias

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM;
  const int num_kb = K / TC_BK;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&split_done[s], 8);
    }
    mbar_init(acc_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {      // one warp allocates the TMEM accumulator (N fp32 columns, power of 2)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp >= 2 && (epi == 1 || epi == 2))
    for (int i = threadIdx.x - 64; i < N; i += 256) bias_s[i] = bias[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + s * stage_bytes;
        mbar_expect_tx(&full[s], a_bytes + 2 * b_bytes);
        tma_load_2d(st, &tmA, &full[s], kb * TC_BK, m0);
        tma_load_2d(st + 2 * a_bytes, &tmB_hi, &full[s], kb * TC_BK, 0);
        tma_load_2d(st + 2 * a_bytes + b_bytes, &tmB_lo, &full[s], kb * TC_BK, 0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait(&full[s], ph);                // B_hi / B_lo landed
        mbar_wait(&split_done[s], ph);          // A split into hi / lo by the converter warps
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem_u32(smem + s * stage_bytes);
        const uint64_t dAh = make_desc(st), dAl = make_desc(st + a_bytes);
        const uint64_t dBh = make_desc(st + 2 * a_bytes);
        const uint64_t dBl = make_desc(st + 2 * a_bytes + b_bytes);
#pragma unroll
        for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
          const uint64_t adv = (uint64_t)((k * TC_UMMA_K * 4) >> 4);   // +32 B along K
          umma_tf32(tmem_base, dAh + adv, dBh + adv, idesc, (kb | k) != 0);
          umma_tf32(tmem_base, dAl + adv, dBh + adv, idesc, 1);
          umma_tf32(tmem_base, dAh + adv, dBl + adv, idesc, 1);
        }
        umma_commit(&empty[s]);                 // frees the smem stage when the MMAs retire
      }
      umma_commit(acc_ready);                   // accumulator complete
    }
  } else {
    // ===== main loop: split the raw fp32 A tile into TF32-exact hi (in place) + remainder lo =====
    const int ct = threadIdx.x - 64;            // 0..255
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % TC_STAGES;
      const uint32_t ph = (kb / TC_STAGES) & 1;
      mbar_wait(&full[s], ph);
      float4* raw = reinterpret_cast<float4*>(smem + s * stage_bytes);
      float4* lo = reinterpret_cast<float4*>(smem + s * stage_bytes + a_bytes);
#pragma unroll
      for (int i = 0; i < (TC_BM * TC_BK / 4) / 256; ++i) {
        const int idx = ct + 256 * i;
        const float4 v = raw[idx];
        float4 h, l;
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); h.x = __uint_as_float(t); l.x = v.x - h.x;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); h.y = __uint_as_float(t); l.y = v.y - h.y;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); h.z = __uint_as_float(t); l.z = v.z - h.z;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); h.w = __uint_as_float(t); l.w = v.w - h.w;
        raw[idx] = h;
        lo[idx] = l;
      }
      // generic-proxy writes must be visible to the tensor core's async-proxy reads
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&split_done[s]))
                     : "memory");
      }
    }
    // ===== epilogue: warp w reads TMEM lanes 32 * (w % 4) .. + 31; the two warps that share a
    // lane quarter take one half of the columns each =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int nh = ((N / 32 + 1) / 2) * 32;      // columns of the first half (multiple of 32)
    const int c_beg = half ? nh : 0, c_end = half ? N : nh;
    mbar_wait(acc_ready, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = m0 + q * 32 + lane;
    for (int c0 = c_beg; c0 < c_end; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      if (row < M) {
        float* out = C + (size_t)row * ldc + c0;
        const float* ax = (epi == 3) ? aux + (size_t)row * ldaux + c0 : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float x = __uint_as_float(r[j + i]);
            if (epi == 1 || epi == 2) x += bias_s[c0 + j + i];
            if (epi == 2) x = rk_act_fn(act, x);
            v[i] = x;
          }
          if (epi == 3) {
            const float4 h = *reinterpret_cast<const float4*>(ax + j);
            v[0] *= rk_act_grad(act, h.x);
            v[1] *= rk_act_grad(act, h.y);
            v[2] *= rk_act_grad(act, h.z);
            v[3] *= rk_act_grad(act, h.w);
          }
          *reinterpret_cast<float4*>(out + j) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(tmem_cols));
  }
}

// ---- fused forward of a two-hidden-layer policy network, one launch per decision epoch ----------
// heads = (act(act(x W0 + b0) W1 + b1)) Wh + bh      (planner.py:1342-1449: DRPStateLayer -> Dense +
// activation -> Dense + activation -> per-action heads), for a 128-row tile of rollouts per CTA:
//   warps 2-9 : compute the layer-0 activations of the tile from the fluent-major state words
//               (D <= 16 inputs, SIMT) one 16-column k-block at a time, write them as the TF32-exact
//               hi part and the fp32 remainder lo STRAIGHT INTO the MMA's A-operand stage (K-major,
//               64-byte swizzle: 16-byte chunk c of row r lives at chunk c ^ ((r >> 1) & 3)), and --
//               only when the adjoint needs them (train sweep) -- to H0 in HBM;
//   warp 0    : TMA producer of the pre-split weight tiles W1^T hi / lo of the k-block;
//   warp 1    : tcgen05.mma issuer, 3 TF32 products per k-slice into the TMEM accumulator;
//   warps 2-9 : epilogue: tcgen05.ld -> + b1 -> activation -> (train sweep) H1 to HBM, and the head
//               product H1 Wh accumulated in registers from the tile that is already there; the two
//               column halves are combined through shared memory.
// Per decision epoch the hidden activations cross HBM at most once (written, train sweep only) and the
// exact test sweep writes nothing but the head outputs: the separate layer-0, GEMM and head kernels
// moved 5 x 33.5 MB per epoch at B = 32768, H = 256.
#define MLP_STAGES 3
#define MLP_MAX_D 16
#define MLP_MAX_NH 16

struct RkMlpSegs {
  int n;
  int off[16], len[16];
};

// tanh with MUFU exp / rcp and an odd polynomial below 0.25: <= 4e-7 relative error, no branch
__device__ __forceinline__ float rk_fast_tanh(float x) {
  const float ax = fabsf(x);
  const float e = __expf(2.f * ax);
  const float big = 1.f - __fdividef(2.f, e + 1.f);
  const float x2 = ax * ax;
  const float small = ax * fmaf(x2, fmaf(x2, fmaf(x2, -17.f / 315.f, 2.f / 15.f), -1.f / 3.f), 1.f);
  return copysignf(ax < 0.25f ? small : big, x);
}
template <int ACT_CT>
__device__ __forceinline__ float mlp_act(int act, float x) {
  if (ACT_CT == RK_ACT_TANH) return rk_fast_tanh(x);
  return rk_act_fn(act, x);
}

template <int ACT_CT>
__global__ void __launch_bounds__(TC_THREADS, 1)
rk_mlp2_fwd_kernel(const __grid_constant__ CUtensorMap tmB_hi,
                   const __grid_constant__ CUtensorMap tmB_lo, int M, int D, int K, int N, int NH,
                   RkMlpSegs S, const float* __restrict__ x_fm, const float* __restrict__ W0,
                   const float* __restrict__ b0, const float* __restrict__ b1,
                   const float* __restrict__ Wh, const float* __restrict__ bh,
                   float* __restrict__ H0, float* __restrict__ H1, float* __restrict__ heads,
                   uint32_t tmem_cols, int act_rt) {
  const int act = ACT_CT >= 0 ? ACT_CT : act_rt;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = TC_BM * TC_BK * 4;            // 8 KB
  const uint32_t b_bytes = (uint32_t)N * TC_BK * 4;      // 16 KB at N = 256
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  uint8_t* tail = smem + MLP_STAGES * stage_bytes;
  uint64_t* bars = (uint64_t*)tail;
  uint64_t* full = bars;                          // [MLP_STAGES]  weight tiles landed (TMA)
  uint64_t* empty = bars + MLP_STAGES;            // [MLP_STAGES]  MMAs of the stage retired
  uint64_t* a_done = bars + 2 * MLP_STAGES;       // [MLP_STAGES]  layer-0 tile written (8 warps)
  uint64_t* acc_ready = bars + 3 * MLP_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * MLP_STAGES + 1);
  float* b1_s = (float*)(tail + 128);                    // [N]
  float* b0_s = b1_s + N;                                // [K]
  float* bh_s = b0_s + K;                                // [16]
  float* W0_s = bh_s + 16;                               // [D][K]
  float* Wh_s = W0_s + (size_t)D * K;                    // [N][16] (NH padded to 16)
  float* xs = Wh_s + (size_t)N * 16;                     // [128][17] (odd stride: no bank conflicts)
  float* part = reinterpret_cast<float*>(smem);          // [128][16] head partials (stage 0, after the loop)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM;
  const int num_kb = K / TC_BK;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MLP_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&a_done[s], 8);
    }
    mbar_init(acc_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp >= 2) {
    const int ct = threadIdx.x - 64;
    for (int i = ct; i < N; i += 256) b1_s[i] = b1[i];
    for (int i = ct; i < K; i += 256) b0_s[i] = b0[i];
    if (ct < 16) bh_s[ct] = ct < NH ? bh[ct] : 0.f;
    for (int i = ct; i < D * K; i += 256) W0_s[i] = W0[i];
    for (int i = ct; i < N * 16; i += 256) {
      const int c = i >> 4, j = i & 15;
      Wh_s[i] = j < NH ? Wh[(size_t)c * NH + j] : 0.f;
    }
    // the tile's state rows, [128][16] (columns >= D are zero): fluent f is a [M][len_f] block
    for (int i = ct; i < TC_BM * 16; i += 256) {
      const int r = i >> 4, d = i & 15, b = m0 + r;
      float v = 0.f;
      if (b < M && d < D) {
        int f = 0;
        while (f + 1 < S.n && d >= S.off[f + 1]) ++f;
        v = x_fm[(size_t)S.off[f] * M + (size_t)b * S.len[f] + (d - S.off[f])];
      }
      xs[r * 17 + d] = v;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: the two weight tiles of every k-block =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % MLP_STAGES;
        const uint32_t ph = (kb / MLP_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + s * stage_bytes;
        mbar_expect_tx(&full[s], 2 * b_bytes);
        tma_load_2d(st + 2 * a_bytes, &tmB_hi, &full[s], kb * TC_BK, 0);
        tma_load_2d(st + 2 * a_bytes + b_bytes, &tmB_lo, &full[s], kb * TC_BK, 0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % MLP_STAGES;
        const uint32_t ph = (kb / MLP_STAGES) & 1;
        mbar_wait(&full[s], ph);
        mbar_wait(&a_done[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem_u32(smem + s * stage_bytes);
        const uint64_t dAh = make_desc(st), dAl = make_desc(st + a_bytes);
        const uint64_t dBh = make_desc(st + 2 * a_bytes);
        const uint64_t dBl = make_desc(st + 2 * a_bytes + b_bytes);
#pragma unroll
        for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
          const uint64_t adv = (uint64_t)((k * TC_UMMA_K * 4) >> 4);
          umma_tf32(tmem_base, dAh + adv, dBh + adv, idesc, (kb | k) != 0);
          umma_tf32(tmem_base, dAl + adv, dBh + adv, idesc, 1);
          umma_tf32(tmem_base, dAh + adv, dBl + adv, idesc, 1);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(acc_ready);
    }
  } else {
    // ===== layer 0 -> A operand.  thread = (row pair, 4-column group): rows r and r + 64 =====
    const int ct = threadIdx.x - 64;            // 0..255
    const int cg = ct & 3, r_lo = ct >> 2;      // chunk (4 columns = 16 bytes) and row 0..63
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % MLP_STAGES;
      const uint32_t ph = (kb / MLP_STAGES) & 1;
      mbar_wait(&empty[s], ph ^ 1);             // the MMAs that read this stage have retired
      uint8_t* st = smem + s * stage_bytes;
      const int c0 = kb * TC_BK + cg * 4;
      const float4 bias = *reinterpret_cast<const float4*>(b0_s + c0);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = r_lo + 64 * i;
        float4 a = bias;
        const float* xr = xs + r * 17;
#pragma unroll 4
        for (int d = 0; d < D; ++d) {
          const float xv = xr[d];
          const float4 w = *reinterpret_cast<const float4*>(W0_s + (size_t)d * K + c0);
          a.x = fmaf(xv, w.x, a.x);
          a.y = fmaf(xv, w.y, a.y);
          a.z = fmaf(xv, w.z, a.z);
          a.w = fmaf(xv, w.w, a.w);
        }
        a.x = mlp_act<ACT_CT>(act, a.x);
        a.y = mlp_act<ACT_CT>(act, a.y);
        a.z = mlp_act<ACT_CT>(act, a.z);
        a.w = mlp_act<ACT_CT>(act, a.w);
        if (H0 != nullptr && m0 + r < M)
          *reinterpret_cast<float4*>(H0 + (size_t)(m0 + r) * K + c0) = a;
        float4 h, l;
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(a.x)); h.x = __uint_as_float(t); l.x = a.x - h.x;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(a.y)); h.y = __uint_as_float(t); l.y = a.y - h.y;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(a.z)); h.z = __uint_as_float(t); l.z = a.z - h.z;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(a.w)); h.w = __uint_as_float(t); l.w = a.w - h.w;
        const uint32_t o = (uint32_t)r * 64u + (uint32_t)((cg ^ ((r >> 1) & 3)) << 4);
        *reinterpret_cast<float4*>(st + o) = h;
        *reinterpret_cast<float4*>(st + a_bytes + o) = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_done[s]))
                     : "memory");
      }
    }
    // ===== epilogue: hidden layer 1 activations + head product =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int nh = ((N / 32 + 1) / 2) * 32;
    const int c_beg = half ? nh : 0, c_end = half ? N : nh;
    mbar_wait(acc_ready, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = m0 + q * 32 + lane;
    float hacc[MLP_MAX_NH];
#pragma unroll
    for (int j = 0; j < MLP_MAX_NH; ++j) hacc[j] = 0.f;
    for (int c0 = c_beg; c0 < c_end; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float h = mlp_act<ACT_CT>(act, __uint_as_float(r[j + i]) + b1_s[c0 + j + i]);
          v[i] = h;
          const float4* w = reinterpret_cast<const float4*>(Wh_s + (size_t)(c0 + j + i) * 16);
#pragma unroll
          for (int g = 0; g < MLP_MAX_NH / 4; ++g) {
            if (4 * g < NH) {
              const float4 ww = w[g];
              hacc[4 * g] = fmaf(h, ww.x, hacc[4 * g]);
              hacc[4 * g + 1] = fmaf(h, ww.y, hacc[4 * g + 1]);
              hacc[4 * g + 2] = fmaf(h, ww.z, hacc[4 * g + 2]);
              hacc[4 * g + 3] = fmaf(h, ww.w, hacc[4 * g + 3]);
            }
          }
        }
        if (H1 != nullptr && row < M)
          *reinterpret_cast<float4*>(H1 + (size_t)row * N + c0 + j) =
              make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    // combine the two column halves (all MMAs have retired: stage 0 is free to reuse)
    const int rl = q * 32 + lane;
    if (half == 1) {
#pragma unroll
      for (int j = 0; j < MLP_MAX_NH; ++j) part[rl * 16 + j] = hacc[j];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (half == 0 && row < M) {
#pragma unroll
      for (int j = 0; j < MLP_MAX_NH; ++j)
        if (j < NH) heads[(size_t)row * NH + j] = hacc[j] + part[rl * 16 + j] + bh_s[j];
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(tmem_cols));
  }
}

// ---- weight gradients: C[M][N] (+)= A[K][M]^T * B[K][N], reduction over the batch ---------------
// Both operands are activations stored batch-major, so the contraction index K (= rollout) is
// the slow one: the tiles are MN-major.  For 32-bit MN-major operands the only UMMA layout is
// SWIZZLE_128B_BASE32B (32-byte swizzle chunks; TMA mode 128B_ATOM_32B): TMA boxes of 32 (mn) x
// TC_BK (k) land as k-rows of 128 bytes, 4 k-rows per 512-byte atom, next mn-atom at
// LBO = TC_BK * 128 B, next 4-k group at SBO = 512 B; one K=8 MMA spans two atoms (1024 B).
// Both raw fp32 tiles are split into hi (in place) / lo by the converter warps.  Split-K over
// gridDim.y; every CTA writes its partial tile to workspace[split][M][N].
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)(((TC_BK * 128) >> 4) & 0x3fff) << 16;   // LBO: next 32-wide mn atom
  d |= (uint64_t)((512 >> 4) & 0x3fff) << 32;             // SBO: next group of 4 k-rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                                 // SWIZZLE_128B_BASE32B
  return d;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
rk_tcgemm_tn_kernel(const __grid_constant__ CUtensorMap tmA,
                    const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                    int kb_per_split, float* __restrict__ ws, uint32_t tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = TC_BM * TC_BK * 4;            // 8 KB: 4 mn-atoms x 16 k-rows x 128 B
  const uint32_t b_bytes = (uint32_t)N * TC_BK * 4;      // 16 KB at N = 256
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  uint64_t* bars = (uint64_t*)(smem + TC_STAGES * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + TC_STAGES;
  uint64_t* split_done = bars + 2 * TC_STAGES;
  uint64_t* acc_ready = bars + 3 * TC_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * TC_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM;
  const int kb_total = (K + TC_BK - 1) / TC_BK;
  const int kb0 = blockIdx.y * kb_per_split;
  const int kb1 = min(kb_total, kb0 + kb_per_split);
  const int num_kb = max(0, kb1 - kb0);
  const int a_boxes = min(TC_BM, M - m0) / 32, b_boxes = N / 32;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&split_done[s], 8);
    }
    mbar_init(acc_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // mn-atoms of A beyond M (when M - m0 < 128) are never loaded: keep them zero
  if (a_boxes < TC_BM / 32)
    for (uint32_t i = threadIdx.x; i < TC_STAGES * stage_bytes / 16; i += TC_THREADS)
      reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % TC_STAGES;
        const uint32_t ph = (i / TC_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + s * stage_bytes;
        const int k0 = (kb0 + i) * TC_BK;
        mbar_expect_tx(&full[s], (uint32_t)(a_boxes + b_boxes) * TC_BK * 128);
        for (int j = 0; j < a_boxes; ++j)
          tma_load_2d(st + j * TC_BK * 128, &tmA, &full[s], m0 + 32 * j, k0);
        for (int j = 0; j < b_boxes; ++j)
          tma_load_2d(st + 2 * a_bytes + j * TC_BK * 128, &tmB, &full[s], 32 * j, k0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D = F32, A = B = TF32, both MN-major (bits 15, 16), N >> 3, M >> 4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % TC_STAGES;
        const uint32_t ph = (i / TC_STAGES) & 1;
        mbar_wait(&full[s], ph);
        mbar_wait(&split_done[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem_u32(smem + s * stage_bytes);
        const uint64_t dAh = make_desc_mn(st), dAl = make_desc_mn(st + a_bytes);
        const uint64_t dBh = make_desc_mn(st + 2 * a_bytes);
        const uint64_t dBl = make_desc_mn(st + 2 * a_bytes + b_bytes);
#pragma unroll
        for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
          const uint64_t adv = (uint64_t)((k * 1024) >> 4);       // next group of 8 k-rows
          umma_tf32(tmem_base, dAh + adv, dBh + adv, idesc, (i | k) != 0);
          umma_tf32(tmem_base, dAl + adv, dBh + adv, idesc, 1);
          umma_tf32(tmem_base, dAh + adv, dBl + adv, idesc, 1);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(acc_ready);
    }
  } else {
    const int ct = threadIdx.x - 64;
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % TC_STAGES;
      const uint32_t ph = (i / TC_STAGES) & 1;
      mbar_wait(&full[s], ph);
      // raw A | A_lo | raw B | B_lo : split both raw tiles (layout-agnostic elementwise map)
      for (int part = 0; part < 2; ++part) {
        const uint32_t off = part ? 2 * a_bytes : 0, bytes = part ? b_bytes : a_bytes;
        float4* raw = reinterpret_cast<float4*>(smem + s * stage_bytes + off);
        float4* lo = reinterpret_cast<float4*>(smem + s * stage_bytes + off + bytes);
        for (uint32_t idx = ct; idx < bytes / 16; idx += 256) {
          const float4 v = raw[idx];
          float4 h, l;
          uint32_t t;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); h.x = __uint_as_float(t); l.x = v.x - h.x;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); h.y = __uint_as_float(t); l.y = v.y - h.y;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); h.z = __uint_as_float(t); l.z = v.z - h.z;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); h.w = __uint_as_float(t); l.w = v.w - h.w;
          raw[idx] = h;
          lo[idx] = l;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&split_done[s]))
                     : "memory");
      }
    }
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int nh = ((N / 32 + 1) / 2) * 32;
    const int c_beg = half ? nh : 0, c_end = half ? N : nh;
    mbar_wait(acc_ready, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = m0 + q * 32 + lane;
    float* out = ws + (size_t)blockIdx.y * M * N + (size_t)row * N;
    for (int c0 = c_beg; c0 < c_end; c0 += 32) {
      uint32_t r[32];
      if (num_kb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (row < M) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(out + c0 + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                          __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(tmem_cols));
  }
}

__global__ void rk_tc_reduce_kernel(int MN, int splits, const float* __restrict__ ws,
                                    float* __restrict__ C, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MN) return;
  float v = 0.f;
  for (int z = 0; z < splits; ++z) v += ws[(size_t)z * MN + i];       // fixed order
  if (accumulate) v += C[i];
  C[i] = v;
}

// ---- fp32 -> (tf32-exact hi, fp32 remainder lo) ------------------------------------------------
__global__ void rk_tf32_split_kernel(const float4* __restrict__ x, float4* __restrict__ hi,
                                     float4* __restrict__ lo, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = x[i];
  float4 h, l;
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); h.x = __uint_as_float(t); l.x = v.x - h.x;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); h.y = __uint_as_float(t); l.y = v.y - h.y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); h.z = __uint_as_float(t); l.z = v.z - h.z;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); h.w = __uint_as_float(t); l.w = v.w - h.w;
  hi[i] = h;
  lo[i] = l;
}

extern "C" int rk_tf32_split(void* stream, const float* x, float* hi, float* lo, size_t n) {
  if (x == nullptr || hi == nullptr || lo == nullptr || (n & 3) != 0) return RK_E_BADARG;
  if (n == 0) return 0;
  const size_t n4 = n / 4;
  rk_tf32_split_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)x, (float4*)hi, (float4*)lo, n4);
  return (int)cudaGetLastError();
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_t g_encode = nullptr;

static int load_encode() {
  if (g_encode != nullptr) return 0;
  void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (lib == nullptr) lib = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
  if (lib == nullptr) return RK_E_BADARG;
  g_encode = (encode_tiled_t)dlsym(lib, "cuTensorMapEncodeTiled");
  return g_encode != nullptr ? 0 : RK_E_BADARG;
}

// row-major [rows][K] fp32 matrix, box = TC_BK (k) x box_rows, 64B swizzle
static int make_map(CUtensorMap* map, const float* base, int rows, int K, int ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

extern "C" int rk_tcgemm(void* stream, int epi, int M, int N, int K, const float* A, int lda,
                         const float* Bt_hi, const float* Bt_lo, int ldb, float* C, int ldc,
                         const float* bias, const float* aux, int ldaux) {
  if (M <= 0 || A == nullptr || Bt_hi == nullptr || Bt_lo == nullptr || C == nullptr)
    return RK_E_BADARG;
  if (N < 32 || N > 256 || (N % 32) != 0 || K < TC_BK || (K % TC_BK) != 0) return RK_E_BADARG;
  if ((lda % 4) != 0 || (ldb % 4) != 0 || (ldc % 4) != 0) return RK_E_BADARG;
  if (epi < 0 || epi > 3 || ((epi == 1 || epi == 2) && bias == nullptr) ||
      (epi == 3 && (aux == nullptr || (ldaux % 4) != 0)))
    return RK_E_BADARG;
  int rc = load_encode();
  if (rc != 0) return rc;
  CUtensorMap mA, mBh, mBl;
  if ((rc = make_map(&mA, A, M, K, lda, TC_BM)) != 0) return rc;
  if ((rc = make_map(&mBh, Bt_hi, N, K, ldb, N)) != 0) return rc;
  if ((rc = make_map(&mBl, Bt_lo, N, K, ldb, N)) != 0) return rc;
  const size_t stage = 2 * (size_t)TC_BM * TC_BK * 4 + 2 * (size_t)N * TC_BK * 4;
  const size_t smem = TC_STAGES * stage + 128 + (size_t)N * 4 + 1024;   // barriers, bias, slack
  static int configured = 0;
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(rk_tcgemm_kernel<RK_ACT_TANH>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaFuncSetAttribute(rk_tcgemm_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              227 * 1024);
    if (ce != cudaSuccess) return (int)ce;
    configured = 1;
  }
  uint32_t cols = 32;
  while ((int)cols < N) cols <<= 1;
  if (g_rk_act == RK_ACT_TANH)
    rk_tcgemm_kernel<RK_ACT_TANH><<<(M + TC_BM - 1) / TC_BM, TC_THREADS, smem, (cudaStream_t)stream>>>(
        mA, mBh, mBl, M, N, K, epi, C, ldc, bias, aux, ldaux, cols, g_rk_act);
  else
    rk_tcgemm_kernel<-1><<<(M + TC_BM - 1) / TC_BM, TC_THREADS, smem, (cudaStream_t)stream>>>(
        mA, mBh, mBl, M, N, K, epi, C, ldc, bias, aux, ldaux, cols, g_rk_act);
  return (int)cudaGetLastError();
}

// row-major [K][cols] fp32 matrix read as an MN-major operand: box = 32 (cols) x TC_BK (k)
static int make_map_mn(CUtensorMap* map, const float* base, int K, int cols, int ld) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)K};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, TC_BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

extern "C" int rk_tcgemm_tn(void* stream, int accumulate, int M, int N, int K, const float* A,
                            int lda, const float* B, int ldb, float* C, float* workspace,
                            int max_splits) {
  if (A == nullptr || B == nullptr || C == nullptr || workspace == nullptr || K <= 0)
    return RK_E_BADARG;
  if (M < 32 || (M % 32) != 0 || N < 32 || N > 256 || (N % 32) != 0 || max_splits < 1)
    return RK_E_BADARG;
  if ((lda % 4) != 0 || (ldb % 4) != 0) return RK_E_BADARG;
  int rc = load_encode();
  if (rc != 0) return rc;
  CUtensorMap mA, mB;
  if ((rc = make_map_mn(&mA, A, K, M, lda)) != 0) return rc;
  if ((rc = make_map_mn(&mB, B, K, N, ldb)) != 0) return rc;
  const int m_tiles = (M + TC_BM - 1) / TC_BM;
  const int kb_total = (K + TC_BK - 1) / TC_BK;
  int splits = (148 + m_tiles - 1) / m_tiles;                  // one wave of CTAs
  if (splits > max_splits) splits = max_splits;
  if (splits > kb_total) splits = kb_total;
  const int kb_per = (kb_total + splits - 1) / splits;
  splits = (kb_total + kb_per - 1) / kb_per;
  const size_t stage = 2 * (size_t)TC_BM * TC_BK * 4 + 2 * (size_t)N * TC_BK * 4;
  const size_t smem = TC_STAGES * stage + 128 + 1024;
  static int configured = 0;
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(rk_tcgemm_tn_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (ce != cudaSuccess) return (int)ce;
    configured = 1;
  }
  uint32_t cols = 32;
  while ((int)cols < N) cols <<= 1;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(m_tiles, splits);
  rk_tcgemm_tn_kernel<<<grid, TC_THREADS, smem, st>>>(mA, mB, M, N, K, kb_per, workspace, cols);
  rk_tc_reduce_kernel<<<(M * N + 255) / 256, 256, 0, st>>>(M * N, splits, workspace, C,
                                                           accumulate ? 1 : 0);
  return (int)cudaGetLastError();
}

// Fused forward of a 2-hidden-layer policy network for one decision epoch (see rk_mlp2_fwd_kernel).
extern "C" int rk_drp_mlp2_forward(void* stream, int B, int D, int H0n, int H1n, int NH, int n_seg,
                                   const int32_t* seg_off, const int32_t* seg_len, const float* x_fm,
                                   const float* W0, const float* b0, const float* W1T_hi,
                                   const float* W1T_lo, const float* b1, const float* Wh,
                                   const float* bh, float* H0, float* H1, float* heads) {
  if (B <= 0 || x_fm == nullptr || W0 == nullptr || b0 == nullptr || W1T_hi == nullptr ||
      W1T_lo == nullptr || b1 == nullptr || Wh == nullptr || bh == nullptr || heads == nullptr)
    return RK_E_BADARG;
  if (D <= 0 || D > MLP_MAX_D || NH <= 0 || NH > MLP_MAX_NH) return RK_E_BADARG;
  if (H0n < TC_BK || H0n > 256 || (H0n % TC_BK) != 0 || (H0n % 4) != 0) return RK_E_BADARG;
  if (H1n < 32 || H1n > 256 || (H1n % 32) != 0) return RK_E_BADARG;
  if (n_seg <= 0 || n_seg > 16 || seg_off == nullptr || seg_len == nullptr) return RK_E_BADARG;
  RkMlpSegs S;
  S.n = n_seg;
  int total = 0;
  for (int f = 0; f < n_seg; ++f) {
    if (seg_len[f] <= 0 || seg_off[f] != total) return RK_E_BADARG;
    S.off[f] = seg_off[f];
    S.len[f] = seg_len[f];
    total += seg_len[f];
  }
  if (total != D) return RK_E_BADARG;
  int rc = load_encode();
  if (rc != 0) return rc;
  CUtensorMap mBh, mBl;
  if ((rc = make_map(&mBh, W1T_hi, H1n, H0n, H0n, H1n)) != 0) return rc;
  if ((rc = make_map(&mBl, W1T_lo, H1n, H0n, H0n, H1n)) != 0) return rc;
  const size_t stage = 2 * (size_t)TC_BM * TC_BK * 4 + 2 * (size_t)H1n * TC_BK * 4;
  const size_t smem = MLP_STAGES * stage + 128 +
                      ((size_t)H1n + H0n + 16 + (size_t)D * H0n + (size_t)H1n * 16 + TC_BM * 17) * 4 +
                      1024;
  if (smem > 227 * 1024) return RK_E_SMEM;
  static int configured = 0;
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(rk_mlp2_fwd_kernel<RK_ACT_TANH>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaFuncSetAttribute(rk_mlp2_fwd_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              227 * 1024);
    if (ce != cudaSuccess) return (int)ce;
    configured = 1;
  }
  uint32_t cols = 32;
  while ((int)cols < H1n) cols <<= 1;
  const int grid = (B + TC_BM - 1) / TC_BM;
  if (g_rk_act == RK_ACT_TANH)
    rk_mlp2_fwd_kernel<RK_ACT_TANH><<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(
        mBh, mBl, B, D, H0n, H1n, NH, S, x_fm, W0, b0, b1, Wh, bh, H0, H1, heads, cols, g_rk_act);
  else
    rk_mlp2_fwd_kernel<-1><<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(
        mBh, mBl, B, D, H0n, H1n, NH, S, x_fm, W0, b0, b1, Wh, bh, H0, H1, heads, cols, g_rk_act);
  return (int)cudaGetLastError();
}

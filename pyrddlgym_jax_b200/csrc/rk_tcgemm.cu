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
#include "rk_tc.cuh"

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

extern "C" int rk_tcgemm(void* stream, int epi, int M, int N, int K, const float* A, int lda,
                         const float* Bt_hi, const float* Bt_lo, int ldb, float* C, int ldc,
                         const float* bias, const float* aux, int ldaux) {
  if (M <= 0 || A == nullptr || Bt_hi == nullptr || Bt_lo == nullptr || C == nullptr)
    return RK_E_BADARG;
  const int act_id = (epi >> 8) & 0xff;              // activation id of the epilogue (rk_act.cuh)
  epi &= 0xff;
  if (act_id < 0 || act_id >= RK_ACT_COUNT) return RK_E_BADARG;
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
  if (act_id == RK_ACT_TANH)
    rk_tcgemm_kernel<RK_ACT_TANH><<<(M + TC_BM - 1) / TC_BM, TC_THREADS, smem, (cudaStream_t)stream>>>(
        mA, mBh, mBl, M, N, K, epi, C, ldc, bias, aux, ldaux, cols, act_id);
  else
    rk_tcgemm_kernel<-1><<<(M + TC_BM - 1) / TC_BM, TC_THREADS, smem, (cudaStream_t)stream>>>(
        mA, mBh, mBl, M, N, K, epi, C, ldc, bias, aux, ldaux, cols, act_id);
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

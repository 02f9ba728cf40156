// rk_wgrad.cu -- weight gradients of the deep reactive policy as ONE reduction over all decision epochs.
//
// The reverse sweep (rk_drp_mlp2_backward, one launch per epoch) leaves the pre-activation cotangents
// of every epoch in HBM next to the activations the forward sweep kept, so every weight gradient of the
// update (planner.py:2873: the transposes of nn.Dense inside the scanned policy call) is a single
// product whose contraction index runs over K = T * B (epoch, rollout) rows:
//
//   dW1 [h0][h1] = H0^T gz1,  db1 = column sums of gz1          -> rk_drp_wgrad_hidden  (N <= 256)
//   dWh [h1][NH] = H1^T gheads                                   -> rk_drp_wgrad_heads   (N  = 16)
//   dW0 [D][h0]  = x^T gz0,   db0 = column sums of gz0           -> rk_drp_wgrad_input   (N  = 16: [x | 1])
//
// C[M][N] = A[K][M]^T B[K][N] on the tensor cores, 3xTF32: both operands are batch-major activations,
// i.e. MN-major UMMA tiles (SWIZZLE_128B_BASE32B, TMA mode 128B_ATOM_32B), split into hi / lo by the
// converter warps after the TMA load.  One CTA per K-slice holds the WHOLE M x N result in TMEM (two
// 128-row accumulators), so each operand is read from HBM exactly once per update; the per-CTA partial
// results are added in CTA order afterwards (deterministic).  Per decision epoch this replaces a
// split-K GEMM + a 19 MB partial reduction + two streaming passes over [B][256].
//
// Accumulator hygiene: the tensor core truncates when it aligns a product to the TMEM accumulator
// (measured on this path: -1.6e-8 of the accumulator per MMA, a bias that does not average out; 8300
// chained MMAs left -1.3e-4).  The wide product therefore DRAINS its accumulators into the CTA's fp32
// workspace slice every WG_SEG k-blocks (<= 384 chained MMAs, bias <= 6e-6) and restarts them; the narrow
// products rotate over 16 accumulator sets (512 TMEM columns) that the epilogue adds in fp32.
//
// The narrow products (N = 16) read their small operand with plain loads and transpose it into a
// K-major 64B-swizzled [16][16] block while splitting it (the head cotangents are [K][NH] rows, the
// state tape is fluent-major): A stays MN-major, B is K-major in the instruction descriptor.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rk.h"
#include "rk_tc.cuh"

#define WG_THREADS 320            // TMA warp, MMA warp, 8 converter / epilogue warps
#define WG_SEG 64                 // k-blocks per accumulator segment of the wide product
#define WG_NSET 16                // accumulator sets of the narrow products
#define WG_AHEAD 6                // k-blocks the narrow operand's plain loads run ahead

struct RkWgSegs {
  int n;
  int off[16], len[16];
};

__device__ __forceinline__ uint64_t wg_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)(((TC_BK * 128) >> 4) & 0x3fff) << 16;   // LBO: next 32-wide mn atom
  d |= (uint64_t)((512 >> 4) & 0x3fff) << 32;             // SBO: next group of 4 k-rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                                 // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void wg_tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void wg_split4(const float4 v, float4& h, float4& l) {
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); h.x = __uint_as_float(t); l.x = v.x - h.x;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); h.y = __uint_as_float(t); l.y = v.y - h.y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); h.z = __uint_as_float(t); l.z = v.z - h.z;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); h.w = __uint_as_float(t); l.w = v.w - h.w;
}

// BMODE 0: B [K][N] through TMA (MN-major), N <= 256, optional column sums of B
// BMODE 1: B [K][NB] plain rows (NB <= 16), loaded and transposed by the converter warps
// BMODE 2: B = [x | 1]: state words of the fluent-major tape (epoch t = k / Bt, rollout b = k % Bt) and
//          a column of ones
template <int BMODE, int STAGES>
__global__ void __launch_bounds__(WG_THREADS, 1)
rk_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                int M, int N, long long K, int kb_per_cta, float* __restrict__ ws,
                float* __restrict__ ws_cs, const float* __restrict__ Bsm, int NB, int Bt, int S_words,
                RkWgSegs SG, uint32_t tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int ntiles = (M + TC_BM - 1) / TC_BM;
  const int a_boxes = M / 32;
  const uint32_t a_bytes = (uint32_t)ntiles * TC_BM * TC_BK * 4;       // hi (raw) region of A
  const int Nacc = BMODE == 0 ? N : 16;                                 // accumulator columns per tile
  const uint32_t b_bytes = BMODE == 0 ? (uint32_t)N * TC_BK * 4 : 1024u;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  uint64_t* bars = (uint64_t*)(smem + (size_t)STAGES * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* split_done = bars + 2 * STAGES;
  uint64_t* acc_ready = bars + 3 * STAGES;
  uint64_t* acc_drained = bars + 3 * STAGES + 1;
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long kb_total = (K + TC_BK - 1) / TC_BK;
  const long long kb0 = (long long)blockIdx.x * kb_per_cta;
  const int num_kb = (int)max(0ll, min(kb_total, kb0 + kb_per_cta) - kb0);
  const int b_boxes = BMODE == 0 ? N / 32 : 0;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&split_done[s], 8);
    }
    mbar_init(acc_ready, 1);
    mbar_init(acc_drained, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // mn-atoms of A beyond M (last tile) are never loaded, padding columns of a narrow B never written
  if ((M % TC_BM) != 0 || BMODE != 0)
    for (uint32_t i = threadIdx.x; i < (uint32_t)STAGES * stage_bytes / 16; i += WG_THREADS)
      reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * stage_bytes;
        const int k0 = (int)((kb0 + i) * TC_BK);
        mbar_expect_tx(&full[s], (uint32_t)(a_boxes + b_boxes) * TC_BK * 128);
        for (int j = 0; j < a_boxes; ++j)
          tma_load_2d(st + j * TC_BK * 128, &tmA, &full[s], 32 * j, k0);
        for (int j = 0; j < b_boxes; ++j)
          tma_load_2d(st + 2 * a_bytes + j * TC_BK * 128, &tmB, &full[s], 32 * j, k0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D = F32, A = B = TF32; A MN-major (bit 15); B MN-major (bit 16) only when it came through TMA
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) |
                             (BMODE == 0 ? (1u << 16) : 0u) | ((uint32_t)(Nacc >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full[s], ph);
        mbar_wait(&split_done[s], ph);
        const int seg = i / WG_SEG;
        const bool seg_start = BMODE == 0 && (i % WG_SEG) == 0;
        if (seg_start && seg > 0) mbar_wait(acc_drained, (seg - 1) & 1);   // restart the accumulators
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t sb = st + 2 * a_bytes;
        const bool fresh = BMODE == 0 ? seg_start : i < WG_NSET;
        const uint32_t set_off = BMODE == 0 ? 0u : (uint32_t)((i % WG_NSET) * ntiles * 16);
        for (int tl = 0; tl < ntiles; ++tl) {
          const uint32_t sa = st + (uint32_t)tl * (TC_BM * TC_BK * 4);
          const uint64_t dAh = wg_desc_mn(sa), dAl = wg_desc_mn(sa + a_bytes);
          const uint64_t dBh = BMODE == 0 ? wg_desc_mn(sb) : make_desc(sb);
          const uint64_t dBl = BMODE == 0 ? wg_desc_mn(sb + b_bytes) : make_desc(sb + b_bytes);
          const uint32_t d = tmem_base + set_off + (uint32_t)(tl * Nacc);
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
            const uint64_t adv_a = (uint64_t)((k * 1024) >> 4);           // next group of 8 k-rows
            const uint64_t adv_b = BMODE == 0 ? adv_a : (uint64_t)((k * TC_UMMA_K * 4) >> 4);
            umma_tf32(d, dAh + adv_a, dBh + adv_b, idesc, (!fresh || k != 0) ? 1u : 0u);
            umma_tf32(d, dAl + adv_a, dBh + adv_b, idesc, 1);
            umma_tf32(d, dAh + adv_a, dBl + adv_b, idesc, 1);
          }
        }
        umma_commit(&empty[s]);
        if (BMODE == 0 && ((i + 1) % WG_SEG == 0) && i + 1 < num_kb) umma_commit(acc_ready);
      }
      umma_commit(acc_ready);
    }
  } else {
    const int ct = threadIdx.x - 64;              // 0..255
    float cs[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) cs[j] = 0.f;
    const uint32_t a_f4 = (uint32_t)a_boxes * TC_BK * 128 / 16, b_f4 = b_bytes / 16;
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    float* out = ws + (size_t)blockIdx.x * M * Nacc;
    // wide product: add the accumulators of segment `seg` to the workspace slice (the first one stores)
    auto drain = [&](int seg, bool last) {
      mbar_wait(acc_ready, seg & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int nh = ((N / 32 + 1) / 2) * 32;
      const int c_beg = half ? nh : 0, c_end = half ? N : nh;
      for (int tl = 0; tl < ntiles; ++tl) {
        const int row = tl * TC_BM + q * 32 + lane;
        for (int c0 = c_beg; c0 < c_end; c0 += 32) {
          uint32_t r[32];
          if (num_kb > 0) {
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * N + c0), r);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0u;
          }
          if (row < M) {
            float4* dst = reinterpret_cast<float4*>(out + (size_t)row * N + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                     __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
              if (seg > 0) {
                const float4 o = dst[j];
                v.x += o.x;
                v.y += o.y;
                v.z += o.z;
                v.w += o.w;
              }
              dst[j] = v;
            }
          }
        }
      }
      if (!last) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc_drained))
                       : "memory");
        }
      }
    };
    // narrow operand: thread ct owns element (k-row kk, column n) of every k-block's [16][16] tile
    bool bw = false;
    uint32_t boff = 0;
    int b_kk = 0, b_n = 0, b_f = 0;
    if (BMODE == 1) {
      bw = ct < 16 * NB;
      b_kk = ct / max(NB, 1);
      b_n = ct - b_kk * NB;
    } else if (BMODE == 2) {
      b_kk = ct & 15;
      b_n = ct >> 4;
      bw = b_n <= NB;
      if (b_n < NB)
        while (b_f + 1 < SG.n && b_n >= SG.off[b_f + 1]) ++b_f;
    }
    if (BMODE != 0)
      boff = (uint32_t)b_n * 64u + ((((uint32_t)b_kk >> 2) ^ (((uint32_t)b_n >> 1) & 3u)) << 4) +
             ((uint32_t)b_kk & 3u) * 4u;
    auto fetch_b = [&](int i) -> float {
      if (!bw || i >= num_kb) return 0.f;
      const long long k = (kb0 + i) * TC_BK + b_kk;
      if (k >= K) return 0.f;
      if (BMODE == 1) return __ldg(Bsm + (size_t)k * NB + b_n);
      if (b_n == NB) return 1.f;
      const long long t = k / Bt;
      const int b = (int)(k - t * Bt);
      return __ldg(Bsm + (size_t)t * S_words * Bt + (size_t)SG.off[b_f] * Bt + (size_t)b * SG.len[b_f] +
                   (b_n - SG.off[b_f]));
    };
    float bq[WG_AHEAD];
    float bsum = 0.f;                      // BMODE 1: column sum of this thread's column (bias gradient)
#pragma unroll
    for (int u = 0; u < WG_AHEAD; ++u) bq[u] = BMODE != 0 ? fetch_b(u) : 0.f;
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      uint8_t* st = smem + (size_t)s * stage_bytes;
      // the narrow operand's element of this thread: fetched WG_AHEAD k-blocks ahead (a plain load
      // consumed in the same k-block would expose a full HBM round trip per k-block)
      float bv = 0.f;
      if (BMODE != 0) {
        bv = bq[0];
#pragma unroll
        for (int u = 0; u + 1 < WG_AHEAD; ++u) bq[u] = bq[u + 1];
        bq[WG_AHEAD - 1] = fetch_b(i + WG_AHEAD);
        if (BMODE == 1) bsum += bv;
      }
      mbar_wait(&full[s], ph);
      {   // A: raw -> hi in place, lo behind it
        float4* raw = reinterpret_cast<float4*>(st);
        float4* lo = reinterpret_cast<float4*>(st + a_bytes);
        for (uint32_t idx = ct; idx < a_f4; idx += 256) {
          float4 h, l;
          wg_split4(raw[idx], h, l);
          raw[idx] = h;
          lo[idx] = l;
        }
      }
      if (BMODE == 0) {
        float4* raw = reinterpret_cast<float4*>(st + 2 * a_bytes);
        float4* lo = reinterpret_cast<float4*>(st + 2 * a_bytes + b_bytes);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t idx = ct + 256 * u;
          if (idx < b_f4) {
            const float4 v = raw[idx];
            float4 h, l;
            wg_split4(v, h, l);
            raw[idx] = h;
            lo[idx] = l;
            cs[4 * u] += v.x;
            cs[4 * u + 1] += v.y;
            cs[4 * u + 2] += v.z;
            cs[4 * u + 3] += v.w;
          }
        }
      } else if (bw) {
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(bv));
        const float h = __uint_as_float(t);
        *reinterpret_cast<float*>(st + 2 * a_bytes + boff) = h;
        *reinterpret_cast<float*>(st + 2 * a_bytes + b_bytes + boff) = bv - h;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&split_done[s]))
                     : "memory");
      }
      // wide product: a finished segment is drained (after the next STAGES - 1 k-blocks have been
      // split, so the MMAs restart at once) into the fp32 workspace slice of this CTA
      if (BMODE == 0) {
        const int due = i - (STAGES - 1);          // k-block whose segment end is drained now
        if (due >= 0 && (due + 1) % WG_SEG == 0 && due + 1 < num_kb)
          drain((due + 1) / WG_SEG - 1, false);
      }
    }
    if (BMODE == 0) {
      // segment ends hidden by the loop end, then the last (partial) segment
      for (int due = max(0, num_kb - (STAGES - 1)); due < num_kb; ++due)
        if ((due + 1) % WG_SEG == 0 && due + 1 < num_kb) drain((due + 1) / WG_SEG - 1, false);
      drain(num_kb > 0 ? (num_kb - 1) / WG_SEG : 0, true);
      if (ws_cs != nullptr) {
        // column sums of B: thread ct's float4 u of every k-block covers four fixed columns
        // (128B_ATOM_32B swizzle: 32-byte chunk p of k-row r holds logical chunk p ^ (r & 3))
        float* part = reinterpret_cast<float*>(smem);          // [256][16], stages are free now
#pragma unroll
        for (int j = 0; j < 16; ++j) part[ct * 16 + j] = cs[j];
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ct < N) {
          const int box = ct >> 5, lc = (ct >> 3) & 3, hf = (ct >> 2) & 1, e = ct & 3;
          float acc = 0.f;
          for (int r = 0; r < 16; ++r) {
            const int src = ((box & 1) << 7) | (r << 3) | ((lc ^ (r & 3)) << 1) | hf;
            acc += part[src * 16 + (box >> 1) * 4 + e];
          }
          ws_cs[(size_t)blockIdx.x * N + ct] = acc;
        }
      }
    } else {
      mbar_wait(acc_ready, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (BMODE == 1 && ws_cs != nullptr) {
        // column sums of the narrow operand: 16 threads (k-rows) per column, added in fixed order
        float* part = reinterpret_cast<float*>(smem);          // stages are free now
        part[ct] = bw ? bsum : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ct < NB) {
          float acc = 0.f;
          for (int kk = 0; kk < 16; ++kk) acc += part[kk * NB + ct];
          ws_cs[(size_t)blockIdx.x * 16 + ct] = acc;
        }
      }
      const int tl = half;
      if (tl < ntiles) {
        const int row = tl * TC_BM + q * 32 + lane;
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        const int nset = min(WG_NSET, num_kb);
        for (int st = 0; st < nset; ++st) {          // the accumulator sets, fixed order
          uint32_t r[16];
          wg_tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(st * ntiles * 16 + tl * 16), r);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(r[j]);
        }
        if (row < M) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(out + (size_t)row * 16 + j) =
                make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
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

// out[o] (+)= sum over the CTA partials, in CTA order.  mode 0: out[m][n] <- ws[s][m][n] (ldw = N);
// mode 1: out[m][n < NO] <- ws[s][m][n] (ldw = 16); mode 2: out[d][m] <- ws[s][m][d] for d < NO, and
// out2[m] <- ws[s][m][NO] (ldw = 16)
__global__ void rk_wgrad_reduce_kernel(int mode, int M, int NO, int ldw, int splits,
                                       const float* __restrict__ ws, float* __restrict__ out,
                                       float* __restrict__ out2, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int cols = mode == 2 ? NO + 1 : NO;
  if (i >= M * cols) return;
  const int m = i / cols, n = i - m * cols;
  float v = 0.f;
  const size_t stride = (size_t)M * ldw;
  for (int z = 0; z < splits; ++z) v += ws[(size_t)z * stride + (size_t)m * ldw + n];
  float* dst = mode == 2 ? (n == NO ? out2 + m : out + (size_t)n * M + m) : out + (size_t)m * NO + n;
  if (accumulate) v += *dst;
  *dst = v;
}
__global__ void rk_wgrad_reduce_cs_kernel(int N, int splits, const float* __restrict__ ws_cs,
                                          float* __restrict__ out, int accumulate, int ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float v = 0.f;
  for (int z = 0; z < splits; ++z) v += ws_cs[(size_t)z * ld + i];
  if (accumulate) v += out[i];
  out[i] = v;
}

// ---- host side ----------------------------------------------------------------------------------
static int wg_make_map_mn(CUtensorMap* map, const float* base, long long K, int cols, int ld) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)K};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, TC_BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

static int wg_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

extern "C" int rk_drp_wgrad_ctas(void) { return wg_sm_count(); }

template <int BMODE, int STAGES>
static int wg_launch(cudaStream_t st, int M, int N, long long K, const float* A, const float* B,
                     const float* Bsm, int NB, int Bt, int S_words, const RkWgSegs& SG, float* ws,
                     float* ws_cs, int* splits_out) {
  int rc = load_encode();
  if (rc != 0) return rc;
  CUtensorMap mA, mB;
  if ((rc = wg_make_map_mn(&mA, A, K, M, M)) != 0) return rc;
  if (BMODE == 0) {
    if ((rc = wg_make_map_mn(&mB, B, K, N, N)) != 0) return rc;
  } else {
    mB = mA;
  }
  const int ntiles = (M + TC_BM - 1) / TC_BM;
  const long long kb_total = (K + TC_BK - 1) / TC_BK;
  int splits = wg_sm_count();
  if ((long long)splits > kb_total) splits = (int)kb_total;
  const int kb_per = (int)((kb_total + splits - 1) / splits);
  splits = (int)((kb_total + kb_per - 1) / kb_per);
  const size_t a_bytes = (size_t)ntiles * TC_BM * TC_BK * 4;
  const size_t b_bytes = BMODE == 0 ? (size_t)N * TC_BK * 4 : 1024;
  const size_t smem = STAGES * (2 * a_bytes + 2 * b_bytes) + 256 + 1024;
  if (smem > 227 * 1024) return RK_E_SMEM;
  static int configured = 0;
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(rk_wgrad_kernel<BMODE, STAGES>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (ce != cudaSuccess) return (int)ce;
    configured = 1;
  }
  const int Nacc = BMODE == 0 ? N : 16;
  uint32_t cols = 32;
  while ((int)cols < ntiles * Nacc * (BMODE == 0 ? 1 : WG_NSET)) cols <<= 1;
  rk_wgrad_kernel<BMODE, STAGES><<<splits, WG_THREADS, smem, st>>>(
      mA, mB, M, N, K, kb_per, ws, ws_cs, Bsm, NB, Bt, S_words, SG, cols);
  *splits_out = splits;
  return (int)cudaGetLastError();
}

static bool wg_m_ok(int M) { return M >= 32 && M <= 256 && (M % 32) == 0; }

// dW [M][N] (+)= A[K][M]^T B[K][N], db [N] (+)= column sums of B (db may be NULL).
// workspace: rk_drp_wgrad_ctas() * (M * N + N) floats.
extern "C" int rk_drp_wgrad_hidden(void* stream, int accumulate, int M, int N, long long K,
                                   const float* A, const float* B, float* dW, float* db,
                                   float* workspace) {
  if (A == nullptr || B == nullptr || dW == nullptr || workspace == nullptr || K <= 0)
    return RK_E_BADARG;
  if (!wg_m_ok(M) || N < 32 || N > 256 || (N % 32) != 0) return RK_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  RkWgSegs SG = {};
  float* ws_cs = db != nullptr ? workspace + (size_t)wg_sm_count() * M * N : nullptr;
  int splits = 0;
  int rc = wg_launch<0, 3>(st, M, N, K, A, B, nullptr, 0, 1, 0, SG, workspace, ws_cs, &splits);
  if (rc != 0) return rc;
  rk_wgrad_reduce_kernel<<<(M * N + 255) / 256, 256, 0, st>>>(0, M, N, N, splits, workspace, dW,
                                                              nullptr, accumulate ? 1 : 0);
  if (db != nullptr)
    rk_wgrad_reduce_cs_kernel<<<(N + 255) / 256, 256, 0, st>>>(N, splits, ws_cs, db,
                                                               accumulate ? 1 : 0, N);
  return (int)cudaGetLastError();
}

// dWh [M][NH] (+)= H[K][M]^T G[K][NH], dbh [NH] (+)= column sums of G (dbh may be NULL); NH <= 16.
// workspace: rk_drp_wgrad_ctas() * (M + 1) * 16 floats.
extern "C" int rk_drp_wgrad_heads(void* stream, int accumulate, int M, int NH, long long K,
                                  const float* H, const float* G, float* dWh, float* dbh,
                                  float* workspace) {
  if (H == nullptr || G == nullptr || dWh == nullptr || workspace == nullptr || K <= 0)
    return RK_E_BADARG;
  if (!wg_m_ok(M) || NH <= 0 || NH > 16) return RK_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  RkWgSegs SG = {};
  int splits = 0;
  float* ws_cs = dbh != nullptr ? workspace + (size_t)wg_sm_count() * M * 16 : nullptr;
  int rc = wg_launch<1, 6>(st, M, 16, K, H, nullptr, G, NH, 1, 0, SG, workspace, ws_cs, &splits);
  if (rc != 0) return rc;
  rk_wgrad_reduce_kernel<<<(M * NH + 255) / 256, 256, 0, st>>>(1, M, NH, 16, splits, workspace, dWh,
                                                               nullptr, accumulate ? 1 : 0);
  if (dbh != nullptr)
    rk_wgrad_reduce_cs_kernel<<<1, 32, 0, st>>>(NH, splits, ws_cs, dbh, accumulate ? 1 : 0, 16);
  return (int)cudaGetLastError();
}

// dW0 [D][M] (+)= x^T gz0, db0 [M] (+)= column sums of gz0: gz0 [K = T * Bt][M]; x: the fluent-major
// state tape, epoch t at word offset t * S_words * Bt (D = sum of seg_len <= 15 of its words).
// workspace: rk_drp_wgrad_ctas() * M * 16 floats.
extern "C" int rk_drp_wgrad_input(void* stream, int accumulate, int M, int D, int T, int Bt,
                                  int S_words, int n_seg, const int32_t* seg_off,
                                  const int32_t* seg_len, const float* gz0, const float* x_tape,
                                  float* dW0, float* db0, float* workspace) {
  if (gz0 == nullptr || x_tape == nullptr || dW0 == nullptr || db0 == nullptr ||
      workspace == nullptr || T <= 0 || Bt <= 0)
    return RK_E_BADARG;
  if (!wg_m_ok(M) || D <= 0 || D > 15 || n_seg <= 0 || n_seg > 16 || seg_off == nullptr ||
      seg_len == nullptr)
    return RK_E_BADARG;
  RkWgSegs SG = {};
  SG.n = n_seg;
  int total = 0;
  for (int f = 0; f < n_seg; ++f) {
    if (seg_len[f] <= 0 || seg_off[f] != total) return RK_E_BADARG;
    SG.off[f] = seg_off[f];
    SG.len[f] = seg_len[f];
    total += seg_len[f];
  }
  if (total != D || S_words < D) return RK_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  int splits = 0;
  int rc = wg_launch<2, 6>(st, M, 16, (long long)T * Bt, gz0, nullptr, x_tape, D, Bt, S_words, SG,
                           workspace, nullptr, &splits);
  if (rc != 0) return rc;
  rk_wgrad_reduce_kernel<<<(M * (D + 1) + 255) / 256, 256, 0, st>>>(2, M, D, 16, splits, workspace,
                                                                    dW0, db0, accumulate ? 1 : 0);
  return (int)cudaGetLastError();
}

// rk_comm.cu -- one-shot sum all-reduce of a short vector over NVLink peer memory (sm_100a).
//
// The only exchange step of the path: the policy gradient (a few KB for a
// straight-line plan) plus the two loss partial sums, once per planner iteration.  At that size a
// ring / tree collective is pure latency, so every rank PUSHES its vector into slot [rank] of a
// buffer on every peer (symmetric allocation, pointers exchanged once at start-up; remote stores
// are posted, nobody waits for a round trip), raises a flag on every peer, waits for the peers'
// flags, and sums the `world` slots of its OWN buffer in RANK ORDER -- every rank computes the
// same bits, so the replicated plans stay identical without a broadcast.
//
// Buffers are double-buffered by the parity of a per-rank epoch counter that the kernel itself
// advances (so a captured CUDA graph can replay it): a rank can be at most one epoch ahead of a
// peer, because passing epoch e needs every peer's flag of epoch e, and a peer raises that flag
// only after it has finished reading epoch e - 1.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rk.h"

struct RkPeers {
  float* buf[RK_MAX_PEERS];       // buf[r]: rank r's symmetric buffer, 2 * world * n_max floats
  uint32_t* flag[RK_MAX_PEERS];   // flag[r]: rank r's flags, 2 * RK_MAX_PEERS words
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// Optional tail of the message: up to RK_PEER_EXTRAS short vectors gathered from their own buffers (the
// two loss sums and the loop-control votes of a planner iteration) and scattered, reduced, to one
// destination -- the staging copies around the exchange become part of its single launch.
#define RK_PEER_EXTRAS 4
struct RkPeerExtras {
  int count, total;
  const float* src[RK_PEER_EXTRAS];
  int len[RK_PEER_EXTRAS];
  float* dst;
};

__device__ __forceinline__ float rk_peer_extra(const RkPeerExtras& X, int j) {
  for (int k = 0; k < X.count; ++k) {
    if (j < X.len[k]) return X.src[k][j];
    j -= X.len[k];
  }
  return 0.f;
}

__global__ void __launch_bounds__(1024)
rk_peer_allreduce_kernel(RkPeers P, int world, int rank, int n, int n_max, float* vec,
                         uint32_t* epoch_ctr, int32_t* status, RkPeerExtras X) {
  __shared__ uint32_t ep;
  __shared__ int failed;
  const int tid = threadIdx.x;
  if (tid == 0) { ep = *epoch_ctr + 1u; failed = 0; }
  __syncthreads();
  const uint32_t e = ep;
  const size_t half = (size_t)(e & 1u) * (size_t)world * (size_t)n_max;
  const size_t slot = half + (size_t)rank * (size_t)n_max;
  for (int i = tid; i < n + X.total; i += blockDim.x) {
    const float v = i < n ? vec[i] : rk_peer_extra(X, i - n);
    for (int r = 0; r < world; ++r) P.buf[r][slot + i] = v;      // posted stores over NVLink
  }
  __threadfence_system();
  __syncthreads();
  if (tid < world) {
    st_release_sys(P.flag[tid] + (e & 1u) * RK_MAX_PEERS + rank, e);
    const uint32_t* f = P.flag[rank] + (e & 1u) * RK_MAX_PEERS + tid;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(f) - e) < 0) {
      if (clock64() - t0 > 20000000000ll) { failed = 1; break; }    // ~10 s: a peer is gone
    }
  }
  __syncthreads();
  if (failed) {
    // a peer never arrived: the vector is NOT reduced.  Record it and stop the context -- every
    // later CUDA call of this process fails loudly instead of training on an un-reduced gradient.
    if (tid == 0 && status != nullptr) *status = 1;
    __threadfence_system();
    __syncthreads();
    asm volatile("trap;");
  }
  const float* local = P.buf[rank] + half;
  for (int i = tid; i < n + X.total; i += blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < world; ++r) acc += ld_relaxed_sys(local + (size_t)r * n_max + i);
    if (i < n) vec[i] = acc; else X.dst[i - n] = acc;
  }
  if (tid == 0) *epoch_ctr = e;
}

extern "C" int rk_peer_allreduce_gather(void* stream, int world, int rank, const uint64_t* peer_bufs,
                                        const uint64_t* peer_flags, int n, int n_max, float* vec,
                                        uint32_t* epoch_ctr, int32_t* status, int n_extra,
                                        const float* const* extra_src, const int32_t* extra_len,
                                        float* extra_dst) {
  if (world < 1 || world > RK_MAX_PEERS || rank < 0 || rank >= world || n <= 0 || n > n_max ||
      peer_bufs == nullptr || peer_flags == nullptr || vec == nullptr || epoch_ctr == nullptr)
    return RK_E_BADARG;
  if (n_extra < 0 || n_extra > RK_PEER_EXTRAS ||
      (n_extra > 0 && (extra_src == nullptr || extra_len == nullptr || extra_dst == nullptr)))
    return RK_E_BADARG;
  RkPeerExtras X;
  X.count = n_extra;
  X.total = 0;
  X.dst = extra_dst;
  for (int k = 0; k < RK_PEER_EXTRAS; ++k) {
    X.src[k] = k < n_extra ? extra_src[k] : nullptr;
    X.len[k] = k < n_extra ? extra_len[k] : 0;
    if (k < n_extra && (extra_src[k] == nullptr || extra_len[k] <= 0)) return RK_E_BADARG;
    X.total += X.len[k];
  }
  if (n + X.total > n_max) return RK_E_BADARG;
  RkPeers P;
  for (int r = 0; r < RK_MAX_PEERS; ++r) {
    P.buf[r] = r < world ? (float*)(uintptr_t)peer_bufs[r] : nullptr;
    P.flag[r] = r < world ? (uint32_t*)(uintptr_t)peer_flags[r] : nullptr;
  }
  rk_peer_allreduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(P, world, rank, n, n_max, vec,
                                                                 epoch_ctr, status, X);
  return (int)cudaGetLastError();
}

extern "C" int rk_peer_allreduce(void* stream, int world, int rank, const uint64_t* peer_bufs,
                                 const uint64_t* peer_flags, int n, int n_max, float* vec,
                                 uint32_t* epoch_ctr, int32_t* status) {
  return rk_peer_allreduce_gather(stream, world, rank, peer_bufs, peer_flags, n, n_max, vec, epoch_ctr,
                                  status, 0, nullptr, nullptr, nullptr);
}

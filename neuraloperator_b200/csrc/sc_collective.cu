// Gradient all-reduce over NVLink peer memory (one process per GPU, buffers exchanged through CUDA symmetric memory).
//
// The data-parallel step has ONE collective: the average of dweight / dbias over the ranks (reference: DDP,
// neuralop/training/trainer.py:203-205).  NCCL takes ~115 us for the 17.8 MB of the headline config on 2 GPUs, of which only the
// ~46 us of the dxm contraction + dx synthesis can hide it.  This kernel is a two-shot all-reduce written for exactly this use:
//   * every rank owns a 1/world slice; a CTA reads its part of the slice from EVERY rank's buffer with 128-bit peer loads (NVLink
//     P2P; local for its own), adds them in rank order (deterministic, and every rank ends up with bit-identical values because
//     each element is reduced exactly once), scales, and writes the result straight into every rank's buffer with peer stores;
//   * cross-rank ordering with flags in the symmetric signal pads: "my inputs are complete" before the loads, "my stores are
//     visible" after them (compare-and-swap hand-shakes with release / acquire at system scope, one flag per (CTA, peer));
//   * a handful of CTAs (the persistent transform kernels leave SMs free: sc_plan_set_reserved_sms) with many loads in flight each.
// It runs on a side stream behind the library's grads_ready event (sc_backward_dense), i.e. underneath the rest of the backward pass.
#include <cstdlib>

#include "sc_plan.h"

namespace sc {

constexpr int AR_THREADS = 512;
constexpr int AR_MAX_WORLD = 8;

struct AllReduceParams {
  float* bufs[AR_MAX_WORLD];
  uint32_t* signals[AR_MAX_WORLD];
  int rank, world;
  long long n_vec;       // float4 elements
  float scale;
};

__device__ __forceinline__ void ar_put(uint32_t* addr) {     // set 0 -> 1 (the peer has consumed the previous flag), release
  uint32_t old;
  do {
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
  } while (old != 0u);
}
__device__ __forceinline__ void ar_wait(uint32_t* addr) {    // consume 1 -> 0, acquire
  uint32_t old;
  do {
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
  } while (old != 1u);
}

// W = world size (compile time: every peer's loads of an iteration are in flight together), U = vectors per thread and iteration:
// W * U 128-bit loads in flight per thread -- peer loads take ~1 us, and only the few SMs the transform kernels leave free run this
template <int W, int U>
__global__ void __launch_bounds__(AR_THREADS, 1) k_allreduce_p2p(const AllReduceParams P) {
  const int tid = threadIdx.x, b = blockIdx.x, G = gridDim.x;
  // ---- every rank's inputs are complete (its producer kernels precede this one on its stream)
  if (tid < W && tid != P.rank) {
    ar_put(P.signals[tid] + b * W + P.rank);
    ar_wait(P.signals[P.rank] + b * W + tid);
  }
  __syncthreads();
  // ---- my slice, this CTA's share of it
  const long long per_rank = (P.n_vec + W - 1) / W;
  const long long lo = (long long)P.rank * per_rank;
  const long long hi = lo + per_rank < P.n_vec ? lo + per_rank : P.n_vec;
  const long long stride = (long long)G * AR_THREADS;
  for (long long i0 = lo + (long long)b * AR_THREADS + tid; i0 < hi; i0 += stride * U) {
    float4 v[W][U];
#pragma unroll
    for (int rr = 0; rr < W; ++rr) {
      const int r = (P.rank + 1 + rr) % W;          // remote peers first, the local buffer last
      const float4* src = reinterpret_cast<const float4*>(P.bufs[r]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + (long long)u * stride;
        if (i < hi) asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[rr][u].x), "=f"(v[rr][u].y), "=f"(v[rr][u].z), "=f"(v[rr][u].w) : "l"(src + i));
        else v[rr][u] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + (long long)u * stride;
      // add in RANK order (rank r sits at position (r - rank - 1) mod W): every rank would compute the same sum; here one does
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < W; ++r) {
        const int rr = (r - P.rank - 1 + 2 * W) % W;
        float4 t = v[0][u];
#pragma unroll
        for (int k = 1; k < W; ++k) if (rr == k) t = v[k][u];
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      if (i < hi) {
        const float4 o = make_float4(acc.x * P.scale, acc.y * P.scale, acc.z * P.scale, acc.w * P.scale);
#pragma unroll
        for (int r = 0; r < W; ++r) {
          float4* dst = reinterpret_cast<float4*>(P.bufs[r]) + i;
          asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
        }
      }
    }
  }
  // ---- my stores are visible everywhere before any rank's stream moves on
  __threadfence_system();
  __syncthreads();
  if (tid < W && tid != P.rank) {
    ar_put(P.signals[tid] + b * W + P.rank);
    ar_wait(P.signals[P.rank] + b * W + tid);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same two-shot all-reduce with cp.async on the load side.  Measured with the LDG kernel above on 2 GPUs: an SM keeps only
// ~20 KB of peer loads in flight (registers + LSU tracking; ~11 GB/s per SM at ~1.8 us peer latency), so the 12 SMs the transform
// kernels leave free need ~65 us for the 17.8 MB of the headline config.  (A TMA version -- 8 KB bulk pieces into a shared-memory
// ring, consumers, bulk stores back -- was measured too: correct, but ~112 us on 12 CTAs, bound by its per-chunk barrier chain.)
// cp.async copies carry no registers: every thread queues W x U 16-byte copies per iteration into PRIVATE shared-memory slots,
// two iterations deep, then adds its own slots in rank order -- no barrier inside the loop, ~100 KB of peer loads in flight per SM.
// ---------------------------------------------------------------------------------------------------------------------
template <int W, int U>
__global__ void __launch_bounds__(AR_THREADS, 1) k_allreduce_cpasync(const AllReduceParams P) {
  extern __shared__ __align__(16) uint8_t ar_smem[];
  const int tid = threadIdx.x, b = blockIdx.x, G = gridDim.x;
  if (tid < W && tid != P.rank) {
    ar_put(P.signals[tid] + b * W + P.rank);
    ar_wait(P.signals[P.rank] + b * W + tid);
  }
  __syncthreads();
  const long long per_rank = (P.n_vec + W - 1) / W;
  const long long lo = (long long)P.rank * per_rank;
  const long long hi = lo + per_rank < P.n_vec ? lo + per_rank : P.n_vec;
  const long long stride = (long long)G * AR_THREADS;
  const long long first = lo + (long long)b * AR_THREADS + tid;
  const long long n_iter = first < hi ? (hi - first + stride * U - 1) / (stride * U) : 0;
  // every CTA thread runs the same number of iterations of the copy pipeline (copies past the slice are predicated off)
  long long n_max = (hi - lo + stride * U - 1) / (stride * U);
  if (n_max < 0) n_max = 0;
  (void)n_iter;
  const uint32_t my_slot = static_cast<uint32_t>(__cvta_generic_to_shared(ar_smem)) + (uint32_t)tid * 16u;
  auto issue = [&](long long it, int buf) {
    const long long i0 = first + it * stride * U;
#pragma unroll
    for (int rr = 0; rr < W; ++rr) {
      const int r = (P.rank + 1 + rr) % W;          // remote peers first
      const float4* src = reinterpret_cast<const float4*>(P.bufs[r]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + (long long)u * stride;
        const uint32_t dst = my_slot + (uint32_t)(((buf * W + r) * U + u) * AR_THREADS) * 16u;
        if (i < hi) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + i) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (n_max > 0) issue(0, 0);
  for (long long it = 0; it < n_max; ++it) {
    const int buf = (int)(it & 1);
    if (it + 1 < n_max) { issue(it + 1, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    const long long i0 = first + it * stride * U;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + (long long)u * stride;
      if (i < hi) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < W; ++r) {             // rank order
          const float4 v = *reinterpret_cast<const float4*>(ar_smem + ((size_t)((buf * W + r) * U + u) * AR_THREADS + tid) * 16);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        const float4 o = make_float4(acc.x * P.scale, acc.y * P.scale, acc.z * P.scale, acc.w * P.scale);
#pragma unroll
        for (int r = 0; r < W; ++r) {
          float4* dst = reinterpret_cast<float4*>(P.bufs[r]) + i;
          asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
        }
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (tid < W && tid != P.rank) {
    ar_put(P.signals[tid] + b * W + P.rank);
    ar_wait(P.signals[P.rank] + b * W + tid);
  }
}

template <int W, int U>
static bool launch_cpasync(const AllReduceParams& P, int n_ctas, cudaStream_t st) {
  constexpr uint32_t smem = 2u * W * U * AR_THREADS * 16u;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    if (!cuda_ok(cudaFuncSetAttribute(k_allreduce_cpasync<W, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(k_allreduce_cpasync)"))
      return false;
    attr_set[dev] = true;
  }
  k_allreduce_cpasync<W, U><<<n_ctas, AR_THREADS, smem, st>>>(P);
  count_launch();
  return cuda_ok(cudaGetLastError(), "k_allreduce_cpasync launch");
}

bool launch_allreduce_p2p(float* const* bufs, uint32_t* const* signals, int rank, int world, int64_t n_floats, float scale, int n_ctas,
                          cudaStream_t st) {
  if (world < 1 || world > AR_MAX_WORLD || rank < 0 || rank >= world) { set_error("allreduce_p2p: world size must be 1..8"); return false; }
  if (n_floats % 4 != 0) { set_error("allreduce_p2p: element count must be a multiple of 4"); return false; }
  if (world == 1 || n_floats == 0) return true;
  AllReduceParams P{};
  for (int r = 0; r < world; ++r) {
    if (bufs[r] == nullptr || signals[r] == nullptr || (reinterpret_cast<uintptr_t>(bufs[r]) & 15u) != 0) { set_error("allreduce_p2p: bad peer pointer"); return false; }
    P.bufs[r] = bufs[r]; P.signals[r] = signals[r];
  }
  P.rank = rank; P.world = world; P.n_vec = n_floats / 4; P.scale = scale;
  if (n_ctas < 1) n_ctas = 1;
  if (n_ctas > 64) n_ctas = 64;
  static const bool use_cpasync = [] { const char* e = getenv("SC_ALLREDUCE_CPASYNC"); return e == nullptr || atoi(e) != 0; }();   // =0: LDG kernel (A/B runs)
  if (use_cpasync) {
    switch (world) {     // W * U <= 12 slots of 8 KB per buffer, two buffers
      case 2: return launch_cpasync<2, 6>(P, n_ctas, st);
      case 3: return launch_cpasync<3, 4>(P, n_ctas, st);
      case 4: return launch_cpasync<4, 3>(P, n_ctas, st);
      case 5: return launch_cpasync<5, 2>(P, n_ctas, st);
      case 6: return launch_cpasync<6, 2>(P, n_ctas, st);
      case 7: return launch_cpasync<7, 1>(P, n_ctas, st);
      default: return launch_cpasync<8, 1>(P, n_ctas, st);
    }
  }
  switch (world) {
    case 2: k_allreduce_p2p<2, 8><<<n_ctas, AR_THREADS, 0, st>>>(P); break;
    case 3: k_allreduce_p2p<3, 4><<<n_ctas, AR_THREADS, 0, st>>>(P); break;
    case 4: k_allreduce_p2p<4, 4><<<n_ctas, AR_THREADS, 0, st>>>(P); break;
    case 5: k_allreduce_p2p<5, 2><<<n_ctas, AR_THREADS, 0, st>>>(P); break;
    case 6: k_allreduce_p2p<6, 2><<<n_ctas, AR_THREADS, 0, st>>>(P); break;
    case 7: k_allreduce_p2p<7, 2><<<n_ctas, AR_THREADS, 0, st>>>(P); break;
    default: k_allreduce_p2p<8, 2><<<n_ctas, AR_THREADS, 0, st>>>(P); break;
  }
  count_launch();
  return cuda_ok(cudaGetLastError(), "k_allreduce_p2p launch");
}

}  // namespace sc

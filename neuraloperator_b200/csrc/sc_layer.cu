// sc_layer.cu -- the Fourier-layer epilogue around the spectral convolution:
//
//   f1  x1  = act( SpectralConv(x) + W_skip x )                       neuralop/layers/fno_block.py:377-397
//           (linear skip = Flattened1dConv, 1x1 conv without bias, skip_connections.py:96-130)
//   f2  out = act( W2 gelu(W1 x1 + b1) + b2 + gate * x )              fno_block.py:399-412, channel_mlp.py:6-119,
//           (soft-gating skip: per-channel weight, skip_connections.py:53-93)
//   f3  tanh stabilizer / fp16 rounding points of the reduced-precision modes (fno_block.py:386-390,
//       spectral_convolution.py:436-462)
//
// Every one of these is a pass over a (B, C, P) tensor (P = points of the grid) whose arithmetic is per point.  PyTorch runs
// them as separate kernels (conv1d, add, gelu, mul, add, gelu ...: ~20 tensor passes per layer); here ONE kernel computes
//
//   pre[b,o,p] = sum_i w[o,i] in[b,i,p] + bias[o] + add[b,o,p] + gate[o] * gated[b,o,p],     out = act(pre)
//
// so f1 is one launch (reads x and the conv output, writes x1) and f2 is two.  Backward is three kernels: the activation
// derivative with the per-channel reductions (dbias, dgate), the same mixing kernel with the transposed weight (din), and a
// reduction GEMM over the points for dweight.
//
// These are plain SIMT fp32 kernels (exact fp32 products, so results agree with PyTorch's fp32 conv to summation order).  They were
// written after the round's GPU minutes were spent, so: no mbarriers, no spin waits, bounded loops only -- and every kernel body
// is a sequence of `__host__ __device__` tile functions that the `sc_hostcheck_*` entry points below run thread by thread on
// host buffers, which is how tests/test_layer_cpu.py checks the index arithmetic of the very code the GPU executes.
#include <cuda_fp16.h>

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sc_plan.h"
#include "sc_umma.cuh"

namespace sc {

#define SC_HD __host__ __device__ __forceinline__

// ---- activations (F.gelu default = exact erf form, torch/nn/functional.py; fno_block.py:150 non_linearity=F.gelu) ------------
SC_HD float act_apply(int act, float v) {
  switch (act) {
    case SC_ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    case SC_ACT_RELU: return v > 0.f ? v : 0.f;
    case SC_ACT_SILU: return v / (1.0f + expf(-v));
    case SC_ACT_TANH: return tanhf(v);
    default: return v;
  }
}
SC_HD float act_grad(int act, float v) {
  if (act == SC_ACT_GELU) {
    const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
    return cdf + v * pdf;
  }
  if (act == SC_ACT_RELU) return v > 0.f ? 1.0f : 0.f;              // (zero at the kink, as torch)
  if (act == SC_ACT_SILU) { const float sg = 1.0f / (1.0f + expf(-v)); return sg * (1.0f + v * (1.0f - sg)); }
  if (act == SC_ACT_TANH) { const float t = tanhf(v); return 1.0f - t * t; }
  return 1.0f;
}
// (d0, d1) += a * (b0, b1) as ONE packed instruction: on sm_100 the scalar FFMA issues every other cycle per sub-partition and
// the two-wide `fma.rn.f32x2` (SASS FFMA2, with `a` as a broadcast scalar operand) is what reaches the full fp32 FMA rate; same
// round-to-nearest fused multiply-add per element, so the host check (two fmaf) computes bit-identical values.
SC_HD void fma2(float& d0, float& d1, float a, float b0, float b1) {
#ifdef __CUDA_ARCH__
  uint64_t av, bv, cv, dv;
  asm("mov.b64 %0, {%1, %2};" : "=l"(av) : "f"(a), "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bv) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(cv) : "f"(d0), "f"(d1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(dv) : "l"(av), "l"(bv), "l"(cv));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(dv));
#else
  d0 = fmaf(a, b0, d0);
  d1 = fmaf(a, b1, d1);
#endif
}
SC_HD void red_add(float* addr, float v) {
#ifdef __CUDA_ARCH__
  atomicAdd(addr, v);
#else
  *addr += v;       // the host check runs the threads one after the other
#endif
}

// =====================================================================================================================
// 1. channel mixing + adds + activation
// =====================================================================================================================
constexpr int MIX_THREADS = 256;
constexpr int MIX_TO = 64;     // output channels per CTA
constexpr int MIX_TP = 128;    // points per CTA
constexpr int MIX_KC = 16;     // input channels per shared-memory stage

struct MixArgs {
  const float* in;             // (B, Ci, P) or nullptr when Ci == 0
  const float* w;              // element (o, i) at w[o * w_so + i * w_si]
  long long w_so, w_si;
  const float* bias;           // (Co) or nullptr
  const float* add;            // (B, Co, P) or nullptr
  const float* gate;           // (Co) or nullptr (coefficient 1)
  const float* gated;          // (B, Co, P) or nullptr
  float* out;                  // (B, Co, P)
  float* pre_out;              // (B, Co, P) or nullptr
  int act;
  int B, Ci, Co;
  long long P;
};
struct MixAcc { float v[8][4]; };      // thread tile: 8 output channels x 4 points (points strided by 32: coalesced rows)

struct Grid3 { long long x; int y, z; };
static Grid3 mix_grid(const MixArgs& a) {
  return Grid3{(a.P + MIX_TP - 1) / MIX_TP, (a.Co + MIX_TO - 1) / MIX_TO, a.B};
}

// stage the [KC x TP] block of `in` and the [KC x TO] block of w^T (zero beyond the extents)
SC_HD void mix_load(const MixArgs& a, int b, int o0, long long p0, int i0, int tid, float* s_in, float* s_w) {
  const float* in_b = a.in + (long long)b * a.Ci * a.P;
#pragma unroll
  for (int r = 0; r < MIX_KC * MIX_TP / MIX_THREADS; ++r) {
    const int idx = tid + r * MIX_THREADS;
    const int i = idx / MIX_TP, p = idx % MIX_TP;
    float v = 0.f;
    if (i0 + i < a.Ci && p0 + p < a.P) v = in_b[(long long)(i0 + i) * a.P + p0 + p];
    s_in[idx] = v;                                   // [i][p]
  }
#pragma unroll
  for (int r = 0; r < MIX_KC * MIX_TO / MIX_THREADS; ++r) {
    const int idx = tid + r * MIX_THREADS;
    const int i = idx / MIX_TO, o = idx % MIX_TO;
    float v = 0.f;
    if (i0 + i < a.Ci && o0 + o < a.Co) v = a.w[(long long)(o0 + o) * a.w_so + (long long)(i0 + i) * a.w_si];
    s_w[idx] = v;                                    // [i][o]
  }
}

SC_HD void mix_fma(int tid, const float* s_in, const float* s_w, MixAcc& acc) {
  const int ty = tid >> 5, tx = tid & 31;
#pragma unroll
  for (int i = 0; i < MIX_KC; ++i) {
    float wv[8], xv[4];
#pragma unroll
    for (int r = 0; r < 8; ++r) wv[r] = s_w[i * MIX_TO + ty * 8 + r];          // one address per warp: broadcast
#pragma unroll
    for (int j = 0; j < 4; ++j) xv[j] = s_in[i * MIX_TP + tx + 32 * j];        // consecutive lanes, consecutive banks
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      fma2(acc.v[r][0], acc.v[r][1], wv[r], xv[0], xv[1]);
      fma2(acc.v[r][2], acc.v[r][3], wv[r], xv[2], xv[3]);
    }
  }
}

SC_HD void mix_store(const MixArgs& a, int b, int o0, long long p0, int tid, const MixAcc& acc) {
  const int ty = tid >> 5, tx = tid & 31;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int o = o0 + ty * 8 + r;
    if (o < a.Co) {
      const float bo = a.bias != nullptr ? a.bias[o] : 0.f;
      const float go = a.gate != nullptr ? a.gate[o] : 1.f;
      const long long row = ((long long)b * a.Co + o) * a.P;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long p = p0 + tx + 32 * j;
        if (p < a.P) {
          float v = acc.v[r][j] + bo;
          if (a.add != nullptr) v += a.add[row + p];
          if (a.gated != nullptr) v = fmaf(go, a.gated[row + p], v);
          if (a.pre_out != nullptr) a.pre_out[row + p] = v;
          a.out[row + p] = act_apply(a.act, v);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(MIX_THREADS) k_channel_mix(MixArgs a) {
  __shared__ __align__(16) float s_in[MIX_KC * MIX_TP];
  __shared__ __align__(16) float s_w[MIX_KC * MIX_TO];
  const int tid = threadIdx.x;
  const long long p0 = (long long)blockIdx.x * MIX_TP;
  const int o0 = blockIdx.y * MIX_TO;
  const int b = blockIdx.z;
  MixAcc acc;
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc.v[r][j] = 0.f;
  for (int i0 = 0; i0 < a.Ci; i0 += MIX_KC) {
    mix_load(a, b, o0, p0, i0, tid, s_in, s_w);
    __syncthreads();
    mix_fma(tid, s_in, s_w, acc);
    __syncthreads();
  }
  mix_store(a, b, o0, p0, tid, acc);
}

// =====================================================================================================================
// 1b. the same op on the tensor cores (tcgen05, bf16x3) -- OPT-IN (sc_layer_set_tensor_cores / SC_MIX_TC=1), never run on hardware
// =====================================================================================================================
// D[128 points x Np] = A[128 points x Kp] * W[Np x Kp]^T per tile, built on the library's validated bring-up kernel for a tensor-memory
// A operand (`k_umma_selftest_ts`, sc_fast.cu): thread <-> point <-> TMEM lane, so every global access of the kernel is a coalesced
// 128-byte row (consecutive lanes, consecutive points of one channel).  Per tile: each thread loads the Kp channel values of its
// point, splits them into bf16 hi / lo pairs and writes them with tcgen05.st as the A operand (no shared-memory staging of the
// activations at all); W is split once per CTA into two K-major SWIZZLE_128B images (hi, lo) in shared memory; one thread issues
// D = A_hi W_hi + A_lo W_hi + A_hi W_lo (3 x Kp/16 MMAs, fp32 accumulation in TMEM); the epilogue reads D back with tcgen05.ld
// (thread <-> point again) and applies bias / add / gate / activation on the way to global memory.  Tensor memory: D | A_hi | A_lo =
// Np + Kp columns (128 at C = 64: four CTAs per SM).  All waits are the library's bounded mbarrier wait (trap after 2 s, no hang).
using namespace umma;

constexpr int TC_THREADS = 128;

struct MixTcArgs {
  MixArgs m;
  int Np, Kp;                  // Co padded to 16, Ci padded to 64
  long long tiles_per_batch, n_tiles;
  uint32_t tmem_cols, col_ahi, col_alo;
};

__global__ void __launch_bounds__(TC_THREADS) k_channel_mix_tc(MixTcArgs a) {
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const MixArgs& m = a.m;
  const int Np = a.Np, Kp = a.Kp;
  uint8_t* sBhi = tc_smem;
  uint8_t* sBlo = tc_smem + (size_t)Np * Kp * 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, a.tmem_cols);
  if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  for (int idx = tid; idx < Np * Kp; idx += TC_THREADS) {
    const int o = idx / Kp, i = idx % Kp;
    const float w = (o < m.Co && i < m.Ci) ? m.w[(long long)o * m.w_so + (long long)i * m.w_si] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    const uint32_t off = sw128_offset(o, i, Np);
    *reinterpret_cast<__nv_bfloat16*>(sBhi + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(sBlo + off) = lo;
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t idesc = idesc_bf16(128, Np);
  uint32_t phase = 0;
  for (long long t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const int b = (int)(t / a.tiles_per_batch);
    const long long p = (t % a.tiles_per_batch) * 128 + tid;
    const bool p_ok = p < m.P;
    const float* in_p = m.in + (long long)b * m.Ci * m.P + p;             // dereferenced only when p_ok
    // ---- A operand: my point's channel values -> bf16 hi / lo pairs -> tensor memory ----
    for (int c0 = 0; c0 < Kp; c0 += 32) {
      float xv[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int i = c0 + e;
        xv[e] = (p_ok && i < m.Ci) ? in_p[(long long)i * m.P] : 0.f;
      }
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) split2_bf16(xv[2 * e], xv[2 * e + 1], hi[e], lo[e]);
      tmem_st16(tmem + lane_base + a.col_ahi + (uint32_t)(c0 >> 1), hi);
      tmem_st16(tmem + lane_base + a.col_alo + (uint32_t)(c0 >> 1), lo);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    // ---- D = A_hi W_hi + A_lo W_hi + A_hi W_lo ----
    if (tid == 0) {
      for (int ks = 0; ks < Kp / 16; ++ks) {
        const int slab = ks >> 2, kk = ks & 3;
        const uint32_t boff = (uint32_t)(slab * Np * 128 + kk * 32);
        const uint64_t dhi = smem_desc_sw128(smem_u32(sBhi) + boff);
        const uint64_t dlo = smem_desc_sw128(smem_u32(sBlo) + boff);
        mma_bf16_ts(tmem, tmem + a.col_ahi + (uint32_t)(ks * 8), dhi, idesc, ks > 0);
        mma_bf16_ts(tmem, tmem + a.col_alo + (uint32_t)(ks * 8), dhi, idesc, true);
        mma_bf16_ts(tmem, tmem + a.col_ahi + (uint32_t)(ks * 8), dlo, idesc, true);
      }
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1u;
    tc_fence_after_sync();
    // ---- epilogue: D[my point, :] -> + bias + add + gate * gated -> activation -> global ----
    for (int c = 0; c < Np; c += 16) {
      float v[16];
      tmem_ld16(tmem + lane_base + (uint32_t)c, v);
      tmem_ld_wait();
      if (p_ok) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int o = c + e;
          if (o < m.Co) {
            const long long at = ((long long)b * m.Co + o) * m.P + p;
            float r = v[e];
            if (m.bias != nullptr) r += m.bias[o];
            if (m.add != nullptr) r += m.add[at];
            if (m.gated != nullptr) r = fmaf(m.gate != nullptr ? m.gate[o] : 1.f, m.gated[at], r);
            if (m.pre_out != nullptr) m.pre_out[at] = r;
            m.out[at] = act_apply(m.act, r);
          }
        }
      }
    }
    tc_fence_before_sync();
    __syncthreads();                 // every warp has read D and the MMAs have consumed A: both regions are free for the next tile
    tc_fence_after_sync();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, a.tmem_cols);
}

static std::atomic<int> g_mix_tc{-1};      // -1: read SC_MIX_TC on first use
static bool mix_tc_enabled() {
  int v = g_mix_tc.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = std::getenv("SC_MIX_TC");
    v = (e != nullptr && std::atoi(e) == 1) ? 1 : 0;
    g_mix_tc.store(v, std::memory_order_relaxed);
  }
  return v == 1;
}
static bool mix_tc_can(const MixArgs& a) {
  return a.Ci >= 1 && a.Ci <= 256 && a.Co >= 1 && a.Co <= 128 && a.B >= 1 && a.P >= 1;
}
static bool launch_channel_mix_tc(const MixArgs& m, cudaStream_t st) {
  MixTcArgs a{};
  a.m = m;
  a.Np = (m.Co + 15) / 16 * 16;
  a.Kp = (m.Ci + 63) / 64 * 64;
  a.tiles_per_batch = (m.P + 127) / 128;
  a.n_tiles = a.tiles_per_batch * m.B;
  a.col_ahi = (uint32_t)a.Np;
  a.col_alo = (uint32_t)(a.Np + a.Kp / 2);
  const uint32_t need = (uint32_t)(a.Np + a.Kp);
  a.tmem_cols = 32;
  while (a.tmem_cols < need) a.tmem_cols *= 2;             // power of two, <= 512 (Np <= 128, Kp <= 256: 384 -> 512)
  const size_t smem = (size_t)2 * a.Np * a.Kp * 2 + 1024;
  if (smem > 48 * 1024 &&          // (per device and cheap: set whenever the default 48 KB limit is exceeded)
      !cuda_ok(cudaFuncSetAttribute(k_channel_mix_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
               "cudaFuncSetAttribute(k_channel_mix_tc)"))
    return false;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long ctas_per_sm = 512 / a.tmem_cols;         // what tensor memory lets be resident
  long long grid = (long long)sms * (ctas_per_sm < 1 ? 1 : ctas_per_sm);
  if (grid > a.n_tiles) grid = a.n_tiles;
  k_channel_mix_tc<<<(unsigned)grid, TC_THREADS, smem, st>>>(a);
  count_launch();
  return cuda_ok(cudaGetLastError(), "k_channel_mix_tc launch");
}

static void host_channel_mix(const MixArgs& a) {
  const Grid3 g = mix_grid(a);
  std::vector<float> s_in(MIX_KC * MIX_TP), s_w(MIX_KC * MIX_TO);
  std::vector<MixAcc> acc(MIX_THREADS);
  for (int bz = 0; bz < g.z; ++bz)
    for (int by = 0; by < g.y; ++by)
      for (long long bx = 0; bx < g.x; ++bx) {
        const long long p0 = bx * MIX_TP;
        const int o0 = by * MIX_TO;
        std::memset(acc.data(), 0, sizeof(MixAcc) * acc.size());
        for (int i0 = 0; i0 < a.Ci; i0 += MIX_KC) {
          for (int t = 0; t < MIX_THREADS; ++t) mix_load(a, bz, o0, p0, i0, t, s_in.data(), s_w.data());
          for (int t = 0; t < MIX_THREADS; ++t) mix_fma(t, s_in.data(), s_w.data(), acc[t]);
        }
        for (int t = 0; t < MIX_THREADS; ++t) mix_store(a, bz, o0, p0, t, acc[t]);
      }
}

// =====================================================================================================================
// 2. activation derivative + per-channel reductions
// =====================================================================================================================
constexpr int AB_THREADS = 256;
constexpr int AB_PER_THREAD = 16;                       // a CTA covers 4096 points of one (b, c) row
constexpr int AB_SPAN = AB_THREADS * AB_PER_THREAD;

struct ActBwdArgs {
  const float* gout;           // (B, C, P)
  const float* pre;            // (B, C, P) or nullptr (identity)
  int act;
  const float* gate;           // (C) or nullptr (coefficient 1)
  const float* gated;          // (B, C, P) or nullptr
  float* gpre;                 // (B, C, P) or nullptr; may alias gout
  float* dgated;               // (B, C, P) or nullptr
  float* dbias;                // (C) or nullptr, pre-zeroed
  float* dgate;                // (C) or nullptr, pre-zeroed
  int B, C;
  long long P;
};
static Grid3 ab_grid(const ActBwdArgs& a) { return Grid3{(a.P + AB_SPAN - 1) / AB_SPAN, a.C, a.B}; }

SC_HD void ab_thread(const ActBwdArgs& a, int b, int c, long long p0, int tid, float* s_sum) {
  const long long row = ((long long)b * a.C + c) * a.P;
  const float g_c = a.gate != nullptr ? a.gate[c] : 1.f;
  float sb = 0.f, sg = 0.f;
  for (int r = 0; r < AB_PER_THREAD; ++r) {
    const long long p = p0 + tid + (long long)r * AB_THREADS;
    if (p >= a.P) break;
    float g = a.gout[row + p];
    if (a.pre != nullptr) g *= act_grad(a.act, a.pre[row + p]);
    if (a.gpre != nullptr) a.gpre[row + p] = g;
    sb += g;
    if (a.gated != nullptr) sg = fmaf(g, a.gated[row + p], sg);
    if (a.dgated != nullptr) a.dgated[row + p] = g_c * g;
  }
  s_sum[tid] = sb;
  s_sum[AB_THREADS + tid] = sg;
}
// lane 0 of every warp folds the 32 partials of its own warp (each warp touches only its own 32 slots)
SC_HD void ab_fold_warp(int tid, float* s_sum) {
  if ((tid & 31) != 0) return;
  float sb = 0.f, sg = 0.f;
  for (int l = 0; l < 32; ++l) { sb += s_sum[tid + l]; sg += s_sum[AB_THREADS + tid + l]; }
  s_sum[tid] = sb;
  s_sum[AB_THREADS + tid] = sg;
}
SC_HD void ab_finish(const ActBwdArgs& a, int c, int tid, const float* s_sum) {
  if (tid != 0) return;
  float sb = 0.f, sg = 0.f;
  for (int w = 0; w < AB_THREADS / 32; ++w) { sb += s_sum[w * 32]; sg += s_sum[AB_THREADS + w * 32]; }
  if (a.dbias != nullptr) red_add(a.dbias + c, sb);
  if (a.dgate != nullptr) red_add(a.dgate + c, sg);
}

__global__ void __launch_bounds__(AB_THREADS) k_channel_act_backward(ActBwdArgs a) {
  __shared__ float s_sum[2 * AB_THREADS];
  const int tid = threadIdx.x;
  ab_thread(a, blockIdx.z, blockIdx.y, (long long)blockIdx.x * AB_SPAN, tid, s_sum);
  __syncthreads();
  ab_fold_warp(tid, s_sum);
  __syncthreads();
  ab_finish(a, blockIdx.y, tid, s_sum);
}

static void host_channel_act_backward(const ActBwdArgs& a) {
  const Grid3 g = ab_grid(a);
  std::vector<float> s_sum(2 * AB_THREADS);
  for (int bz = 0; bz < g.z; ++bz)
    for (int by = 0; by < g.y; ++by)
      for (long long bx = 0; bx < g.x; ++bx) {
        for (int t = 0; t < AB_THREADS; ++t) ab_thread(a, bz, by, bx * AB_SPAN, t, s_sum.data());
        for (int t = 0; t < AB_THREADS; ++t) ab_fold_warp(t, s_sum.data());
        for (int t = 0; t < AB_THREADS; ++t) ab_finish(a, by, t, s_sum.data());
      }
}

// =====================================================================================================================
// 3. dweight[o, i] = sum_{b, p} g[b, o, p] * in[b, i, p]
// =====================================================================================================================
constexpr int WG_THREADS = 256;
constexpr int WG_T = 64;           // tile of output channels and of input channels
constexpr int WG_KP = 32;          // points per shared-memory stage
constexpr int WG_LD = 68;          // row pitch of the transposed stages (floats; keeps rows 16-byte aligned)
constexpr int WG_CHUNK = 2048;     // points per CTA (multiple of WG_KP)

struct WGradArgs {
  const float* g;              // (B, Co, P)
  const float* in;             // (B, Ci, P)
  float* dw;                   // (Co, Ci) row-major, pre-zeroed
  int B, Ci, Co;
  long long P;
};
struct WgAcc { float v[4][4]; };
static Grid3 wg_grid(const WGradArgs& a) {
  const int to = (a.Co + WG_T - 1) / WG_T, ti = (a.Ci + WG_T - 1) / WG_T;
  return Grid3{(a.P + WG_CHUNK - 1) / WG_CHUNK, to * ti, a.B};
}

SC_HD void wg_load(const WGradArgs& a, int b, int o0, int i0, long long p0, int tid, float* s_g, float* s_x) {
#pragma unroll
  for (int r = 0; r < WG_T * WG_KP / WG_THREADS; ++r) {
    const int idx = tid + r * WG_THREADS;
    const int row = idx / WG_KP, p = idx % WG_KP;       // a warp reads 32 consecutive points of one channel row
    float gv = 0.f, xv = 0.f;
    if (p0 + p < a.P) {
      if (o0 + row < a.Co) gv = a.g[((long long)b * a.Co + o0 + row) * a.P + p0 + p];
      if (i0 + row < a.Ci) xv = a.in[((long long)b * a.Ci + i0 + row) * a.P + p0 + p];
    }
    s_g[p * WG_LD + row] = gv;                          // transposed: [p][channel]
    s_x[p * WG_LD + row] = xv;
  }
}
SC_HD void wg_fma(int tid, const float* s_g, const float* s_x, WgAcc& acc) {
  const int ty = tid >> 4, tx = tid & 15;
#pragma unroll 8
  for (int p = 0; p < WG_KP; ++p) {
    float gv[4], xv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) gv[r] = s_g[p * WG_LD + ty * 4 + r];
#pragma unroll
    for (int c = 0; c < 4; ++c) xv[c] = s_x[p * WG_LD + tx * 4 + c];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      fma2(acc.v[r][0], acc.v[r][1], gv[r], xv[0], xv[1]);
      fma2(acc.v[r][2], acc.v[r][3], gv[r], xv[2], xv[3]);
    }
  }
}
SC_HD void wg_store(const WGradArgs& a, int o0, int i0, int tid, const WgAcc& acc) {
  const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int o = o0 + ty * 4 + r;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + tx * 4 + c;
      if (o < a.Co && i < a.Ci) red_add(a.dw + (long long)o * a.Ci + i, acc.v[r][c]);
    }
  }
}

__global__ void __launch_bounds__(WG_THREADS) k_channel_weight_grad(WGradArgs a) {
  __shared__ __align__(16) float s_g[WG_KP * WG_LD];
  __shared__ __align__(16) float s_x[WG_KP * WG_LD];
  const int tid = threadIdx.x;
  const int ti = (a.Ci + WG_T - 1) / WG_T;
  const int o0 = ((int)blockIdx.y / ti) * WG_T, i0 = ((int)blockIdx.y % ti) * WG_T;
  const int b = blockIdx.z;
  const long long p_begin = (long long)blockIdx.x * WG_CHUNK;
  const long long p_end = (p_begin + WG_CHUNK < a.P) ? p_begin + WG_CHUNK : a.P;
  WgAcc acc;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc.v[r][c] = 0.f;
  for (long long p0 = p_begin; p0 < p_end; p0 += WG_KP) {
    wg_load(a, b, o0, i0, p0, tid, s_g, s_x);
    __syncthreads();
    wg_fma(tid, s_g, s_x, acc);
    __syncthreads();
  }
  wg_store(a, o0, i0, tid, acc);
}

static void host_channel_weight_grad(const WGradArgs& a) {
  const Grid3 g = wg_grid(a);
  const int ti = (a.Ci + WG_T - 1) / WG_T;
  std::vector<float> s_g(WG_KP * WG_LD), s_x(WG_KP * WG_LD);
  std::vector<WgAcc> acc(WG_THREADS);
  for (int bz = 0; bz < g.z; ++bz)
    for (int by = 0; by < g.y; ++by)
      for (long long bx = 0; bx < g.x; ++bx) {
        const int o0 = (by / ti) * WG_T, i0 = (by % ti) * WG_T;
        const long long p_begin = bx * WG_CHUNK;
        const long long p_end = (p_begin + WG_CHUNK < a.P) ? p_begin + WG_CHUNK : a.P;
        std::memset(acc.data(), 0, sizeof(WgAcc) * acc.size());
        for (long long p0 = p_begin; p0 < p_end; p0 += WG_KP) {
          for (int t = 0; t < WG_THREADS; ++t) wg_load(a, bz, o0, i0, p0, t, s_g.data(), s_x.data());
          for (int t = 0; t < WG_THREADS; ++t) wg_fma(t, s_g.data(), s_x.data(), acc[t]);
        }
        for (int t = 0; t < WG_THREADS; ++t) wg_store(a, o0, i0, t, acc[t]);
      }
}

// =====================================================================================================================
// 4. elementwise: tanh stabilizer (fno_block.py:386-390) and the fp16 rounding points of the reduced-precision spectral path
// =====================================================================================================================
constexpr int PW_THREADS = 256;
constexpr int PW_PER_THREAD = 8;

SC_HD float pointwise_elem(int op, const float* a, const float* b, long long i) {
  switch (op) {
    case SC_POINTWISE_TANH: return tanhf(a[i]);
    case SC_POINTWISE_TANH_BACKWARD: { const float t = b[i]; return a[i] * (1.0f - t * t); }    // a = upstream grad, b = tanh(x)
    case SC_POINTWISE_ROUND_HALF: return __half2float(__float2half_rn(a[i]));                   // x.half() (:436-437), x.chalf() (:451-454)
    // interleaved complex (re, im) pairs, element i = 2k + parity:  out = a + 1j * b   (apply_complex, neuralop/layers/complex.py:55-62)
    case SC_POINTWISE_ADD_I_TIMES: return (i & 1) ? a[i] + b[i - 1] : a[i] - b[i + 1];
    // out = -1j * a   (its gradient with respect to b)
    case SC_POINTWISE_MUL_NEG_I: return (i & 1) ? -a[i - 1] : a[i + 1];
    case SC_POINTWISE_MUL: return a[i] * b[i];                                                  // dropout: x * (mask / (1 - p))
    default: return a[i];
  }
}
__global__ void __launch_bounds__(PW_THREADS) k_pointwise(int op, const float* a, const float* b, float* out, long long n) {
  const long long base = (long long)blockIdx.x * (PW_THREADS * PW_PER_THREAD) + threadIdx.x;
#pragma unroll
  for (int r = 0; r < PW_PER_THREAD; ++r) {
    const long long i = base + (long long)r * PW_THREADS;
    if (i < n) out[i] = pointwise_elem(op, a, b, i);
  }
}

// ---- launches ---------------------------------------------------------------------------------------------------------
static bool grid_ok(const Grid3& g, const char* what) {
  if (g.x > 2147483647LL || g.y > 65535 || g.z > 65535) { set_error(std::string(what) + ": problem exceeds the launch grid limits"); return false; }
  return true;
}

static bool launch_channel_mix(const MixArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.Co <= 0 || a.P <= 0) return true;
  if (mix_tc_enabled() && mix_tc_can(a)) return launch_channel_mix_tc(a, st);
  const Grid3 g = mix_grid(a);
  if (!grid_ok(g, "sc_channel_mix")) return false;
  k_channel_mix<<<dim3((unsigned)g.x, (unsigned)g.y, (unsigned)g.z), MIX_THREADS, 0, st>>>(a);
  count_launch();
  return cuda_ok(cudaGetLastError(), "k_channel_mix launch");
}
static bool launch_channel_act_backward(const ActBwdArgs& a, cudaStream_t st) {
  if (a.dbias != nullptr && !cuda_ok(cudaMemsetAsync(a.dbias, 0, sizeof(float) * (size_t)a.C, st), "dbias memset")) return false;
  if (a.dgate != nullptr && !cuda_ok(cudaMemsetAsync(a.dgate, 0, sizeof(float) * (size_t)a.C, st), "dgate memset")) return false;
  if (a.B <= 0 || a.C <= 0 || a.P <= 0) return true;
  const Grid3 g = ab_grid(a);
  if (!grid_ok(g, "sc_channel_mix_act_backward")) return false;
  k_channel_act_backward<<<dim3((unsigned)g.x, (unsigned)g.y, (unsigned)g.z), AB_THREADS, 0, st>>>(a);
  count_launch();
  return cuda_ok(cudaGetLastError(), "k_channel_act_backward launch");
}
static bool launch_channel_weight_grad(const WGradArgs& a, cudaStream_t st) {
  if (a.Co <= 0 || a.Ci <= 0) return true;
  if (!cuda_ok(cudaMemsetAsync(a.dw, 0, sizeof(float) * (size_t)a.Co * (size_t)a.Ci, st), "dweight memset")) return false;
  if (a.B <= 0 || a.P <= 0) return true;
  const Grid3 g = wg_grid(a);
  if (!grid_ok(g, "sc_channel_mix_weight_grad")) return false;
  k_channel_weight_grad<<<dim3((unsigned)g.x, (unsigned)g.y, (unsigned)g.z), WG_THREADS, 0, st>>>(a);
  count_launch();
  return cuda_ok(cudaGetLastError(), "k_channel_weight_grad launch");
}
static bool launch_pointwise(int op, const float* a, const float* b, float* out, long long n, cudaStream_t st) {
  if (n <= 0) return true;
  const long long per = PW_THREADS * PW_PER_THREAD;
  const long long blocks = (n + per - 1) / per;
  if (blocks > 2147483647LL) { set_error("sc_pointwise: tensor exceeds the launch grid limits"); return false; }
  k_pointwise<<<(unsigned)blocks, PW_THREADS, 0, st>>>(op, a, b, out, n);
  count_launch();
  return cuda_ok(cudaGetLastError(), "k_pointwise launch");
}

}  // namespace sc

// =====================================================================================================================
// C ABI
// =====================================================================================================================
using namespace sc;

#define SC_TRY(expr) do { if (!(expr)) return 1; } while (0)
#define SC_REQUIRE(cond, msg) do { if (!(cond)) { set_error(msg); return 1; } } while (0)

static int fill_mix(MixArgs& a, const char* who, const float* in, const float* w, int64_t w_stride_o, int64_t w_stride_i, const float* bias,
                    const float* add, const float* gate, const float* gated, int act, float* out, float* pre_out, int32_t batch,
                    int32_t in_channels, int32_t out_channels, int64_t n_points) {
  (void)who;
  SC_REQUIRE(out != nullptr, "sc_channel_mix: null output");
  SC_REQUIRE(batch >= 0 && in_channels >= 0 && out_channels >= 0 && n_points >= 0, "sc_channel_mix: negative extent");
  SC_REQUIRE(in_channels == 0 || (in != nullptr && w != nullptr), "sc_channel_mix: in / w are required when in_channels > 0");
  SC_REQUIRE(gate == nullptr || gated != nullptr, "sc_channel_mix: a gate needs the tensor it gates");
  SC_REQUIRE(act >= SC_ACT_IDENTITY && act <= SC_ACT_TANH, "sc_channel_mix: unknown activation");
  a = MixArgs{in, w, (long long)w_stride_o, (long long)w_stride_i, bias, add, gate, gated, out, pre_out, act, batch, in_channels,
              out_channels, (long long)n_points};
  return 0;
}

static int fill_act_backward(ActBwdArgs& a, const float* gout, const float* pre, int act, const float* gate, const float* gated,
                             float* gpre_out, float* dgated_out, float* dbias, float* dgate, int32_t batch, int32_t channels,
                             int64_t n_points) {
  SC_REQUIRE(gout != nullptr, "sc_channel_mix_act_backward: null upstream gradient");
  SC_REQUIRE(batch >= 0 && channels >= 0 && n_points >= 0, "sc_channel_mix_act_backward: negative extent");
  SC_REQUIRE(act >= SC_ACT_IDENTITY && act <= SC_ACT_TANH, "sc_channel_mix_act_backward: unknown activation");
  SC_REQUIRE(act == SC_ACT_IDENTITY || pre != nullptr, "sc_channel_mix_act_backward: the activation derivative needs the pre-activation");
  SC_REQUIRE(dgate == nullptr || gated != nullptr, "sc_channel_mix_act_backward: dgate needs the gated tensor");
  a = ActBwdArgs{gout, act == SC_ACT_IDENTITY ? nullptr : pre, act, gate, gated, gpre_out, dgated_out, dbias, dgate, batch, channels,
                 (long long)n_points};
  return 0;
}

extern "C" {

int sc_channel_mix(const float* in, const float* w, int64_t w_stride_o, int64_t w_stride_i, const float* bias, const float* add,
                   const float* gate, const float* gated, int act, float* out, float* pre_out, int32_t batch, int32_t in_channels,
                   int32_t out_channels, int64_t n_points, sc_stream stream) {
  MixArgs a;
  if (fill_mix(a, "sc_channel_mix", in, w, w_stride_o, w_stride_i, bias, add, gate, gated, act, out, pre_out, batch, in_channels,
               out_channels, n_points)) return 1;
  SC_TRY(launch_channel_mix(a, static_cast<cudaStream_t>(stream)));
  return 0;
}

int sc_channel_mix_act_backward(const float* gout, const float* pre, int act, const float* gate, const float* gated, float* gpre_out,
                                float* dgated_out, float* dbias, float* dgate, int32_t batch, int32_t channels, int64_t n_points,
                                sc_stream stream) {
  ActBwdArgs a;
  if (fill_act_backward(a, gout, pre, act, gate, gated, gpre_out, dgated_out, dbias, dgate, batch, channels, n_points)) return 1;
  SC_TRY(launch_channel_act_backward(a, static_cast<cudaStream_t>(stream)));
  return 0;
}

int sc_channel_mix_weight_grad(const float* gpre, const float* in, float* dw, int32_t batch, int32_t in_channels, int32_t out_channels,
                               int64_t n_points, sc_stream stream) {
  SC_REQUIRE(gpre != nullptr && in != nullptr && dw != nullptr, "sc_channel_mix_weight_grad: null argument");
  SC_REQUIRE(batch >= 0 && in_channels >= 0 && out_channels >= 0 && n_points >= 0, "sc_channel_mix_weight_grad: negative extent");
  WGradArgs a{gpre, in, dw, batch, in_channels, out_channels, (long long)n_points};
  SC_TRY(launch_channel_weight_grad(a, static_cast<cudaStream_t>(stream)));
  return 0;
}

int sc_layer_set_tensor_cores(int enable) {
  g_mix_tc.store(enable != 0 ? 1 : 0, std::memory_order_relaxed);
  return 0;
}

int sc_layer_uses_tensor_cores(void) { return mix_tc_enabled() ? 1 : 0; }

int sc_pointwise(int op, const float* a, const float* b, float* out, int64_t n, sc_stream stream) {
  SC_REQUIRE(op >= SC_POINTWISE_TANH && op <= SC_POINTWISE_MUL, "sc_pointwise: unknown op");
  SC_REQUIRE(n == 0 || (a != nullptr && out != nullptr), "sc_pointwise: null argument");
  SC_REQUIRE((op != SC_POINTWISE_TANH_BACKWARD && op != SC_POINTWISE_ADD_I_TIMES && op != SC_POINTWISE_MUL) || n == 0 || b != nullptr,
             "sc_pointwise: this op needs its second operand");
  SC_REQUIRE((op != SC_POINTWISE_ADD_I_TIMES && op != SC_POINTWISE_MUL_NEG_I) || ((n & 1) == 0 && a != out && b != out),
             "sc_pointwise: the complex-pair ops need an even count and an output that aliases no input");
  SC_TRY(launch_pointwise(op, a, b, out, (long long)n, static_cast<cudaStream_t>(stream)));
  return 0;
}

// ---- host checks: the same tile functions, run thread by thread on HOST buffers (tests only; the package never calls them) ----
int sc_hostcheck_channel_mix(const float* in, const float* w, int64_t w_stride_o, int64_t w_stride_i, const float* bias, const float* add,
                             const float* gate, const float* gated, int act, float* out, float* pre_out, int32_t batch,
                             int32_t in_channels, int32_t out_channels, int64_t n_points) {
  MixArgs a;
  if (fill_mix(a, "sc_hostcheck_channel_mix", in, w, w_stride_o, w_stride_i, bias, add, gate, gated, act, out, pre_out, batch,
               in_channels, out_channels, n_points)) return 1;
  if (a.B <= 0 || a.Co <= 0 || a.P <= 0) return 0;
  SC_TRY(grid_ok(mix_grid(a), "sc_hostcheck_channel_mix"));
  host_channel_mix(a);
  return 0;
}

int sc_hostcheck_channel_mix_act_backward(const float* gout, const float* pre, int act, const float* gate, const float* gated,
                                          float* gpre_out, float* dgated_out, float* dbias, float* dgate, int32_t batch,
                                          int32_t channels, int64_t n_points) {
  ActBwdArgs a;
  if (fill_act_backward(a, gout, pre, act, gate, gated, gpre_out, dgated_out, dbias, dgate, batch, channels, n_points)) return 1;
  if (a.dbias != nullptr) std::memset(a.dbias, 0, sizeof(float) * (size_t)a.C);
  if (a.dgate != nullptr) std::memset(a.dgate, 0, sizeof(float) * (size_t)a.C);
  if (a.B <= 0 || a.C <= 0 || a.P <= 0) return 0;
  SC_TRY(grid_ok(ab_grid(a), "sc_hostcheck_channel_mix_act_backward"));
  host_channel_act_backward(a);
  return 0;
}

int sc_hostcheck_channel_mix_weight_grad(const float* gpre, const float* in, float* dw, int32_t batch, int32_t in_channels,
                                         int32_t out_channels, int64_t n_points) {
  SC_REQUIRE(gpre != nullptr && in != nullptr && dw != nullptr, "sc_hostcheck_channel_mix_weight_grad: null argument");
  WGradArgs a{gpre, in, dw, batch, in_channels, out_channels, (long long)n_points};
  if (a.Co <= 0 || a.Ci <= 0) return 0;
  std::memset(a.dw, 0, sizeof(float) * (size_t)a.Co * (size_t)a.Ci);
  if (a.B <= 0 || a.P <= 0) return 0;
  SC_TRY(grid_ok(wg_grid(a), "sc_hostcheck_channel_mix_weight_grad"));
  host_channel_weight_grad(a);
  return 0;
}

int sc_hostcheck_pointwise(int op, const float* a, const float* b, float* out, int64_t n) {
  SC_REQUIRE(op >= SC_POINTWISE_TANH && op <= SC_POINTWISE_MUL, "sc_hostcheck_pointwise: unknown op");
  SC_REQUIRE((op != SC_POINTWISE_ADD_I_TIMES && op != SC_POINTWISE_MUL_NEG_I) || ((n & 1) == 0 && a != out && b != out),
             "sc_hostcheck_pointwise: the complex-pair ops need an even count and an output that aliases no input");
  // the kernel's own index walk: block, thread, slot
  const long long per = PW_THREADS * PW_PER_THREAD;
  const long long blocks = (n + per - 1) / per;
  for (long long blk = 0; blk < blocks; ++blk)
    for (int t = 0; t < PW_THREADS; ++t)
      for (int r = 0; r < PW_PER_THREAD; ++r) {
        const long long i = blk * per + t + (long long)r * PW_THREADS;
        if (i < n) out[i] = pointwise_elem(op, a, b, i);
      }
  return 0;
}

}  // extern "C"

// tcgen05 fused transform path (sm_100a).  See DESIGN.md "fast path" for the derivation.
//
// Both transforms are chains of two small GEMMs per 128-row tile, executed on the 5th-generation tensor cores
// with BF16 operands and FP32 accumulation in TMEM.  FP32 accuracy is kept by splitting every operand into
// two bf16 terms (x = x_hi + x_lo, table = T1 + T2) and accumulating the three significant products
// x_hi*T1 + x_lo*T1 + x_hi*T2 ("bf16x3", relative error ~1e-5).
#include <cstring>
#include <vector>

#include "sc_fast.h"
#include "sc_umma.cuh"

namespace sc {

using namespace umma;

// =====================================================================================================
// self-test: D[128 x N] = A[128 x K] * B[N x K]^T with bf16-rounded operands -- exercises the swizzled operand
// stores, the shared-memory / instruction descriptors, TMEM allocation, tcgen05.mma, commit and tcgen05.ld
// exactly the way the transform kernels use them.
// =====================================================================================================
__global__ void __launch_bounds__(128) k_umma_selftest(const float* __restrict__ A, const float* __restrict__ B,
                                                        float* __restrict__ D, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_base, 128);
  if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  for (int idx = tid; idx < 128 * K; idx += 128) {
    const int r = idx / K, k = idx % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + sw128_offset(r, k, 128)) = __float2bfloat16_rn(A[idx]);
  }
  for (int idx = tid; idx < N * K; idx += 128) {
    const int r = idx / K, k = idx % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + sw128_offset(r, k, N)) = __float2bfloat16_rn(B[idx]);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16(128, N);
    for (int ks = 0; ks < K / 16; ++ks) {
      const int slab = ks >> 2, kk = ks & 3;
      const uint64_t da = smem_desc_sw128(smem_u32(sA) + slab * 128 * 128 + kk * 32);
      const uint64_t db = smem_desc_sw128(smem_u32(sB) + slab * N * 128 + kk * 32);
      mma_bf16_ss(tmem, da, db, idesc, ks > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)(warp * 32 + lane) * N + c + i] = v[i];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

bool umma_selftest(const float* A, const float* B, float* D, int N, int K, cudaStream_t st) {
  if (N < 16 || N > 128 || N % 16 != 0 || K < 64 || K > 256 || K % 64 != 0) {
    set_error("umma selftest: need N in 16..128 step 16 and K in 64..256 step 64");
    return false;
  }
  const size_t smem = (size_t)(128 + N) * K * 2 + 1024;
  if (!cuda_ok(cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
               "cudaFuncSetAttribute(selftest)"))
    return false;
  k_umma_selftest<<<1, 128, smem, st>>>(A, B, D, N, K);
  count_launch();
  return cuda_ok(cudaGetLastError(), "k_umma_selftest launch");
}

// =====================================================================================================
// fused analysis:  tile of 128 image rows  ->  kept modes of the G = 128/H images the tile holds
//
//   stage 1 (last dim)    D1[h, j]   = sum_w x[h, w] * TA[w, j]                 M=128 (rows)  N=2*N1  K=W
//   stage 2 (leading dim) D2[i', n]  = sum_{p,h} A2[i', (p,h)] * R_p[h, n]      M=128         N=N1    K=256
//
//   warps 0-3  epilogue  (TMEM -> registers; D1 -> bf16 hi/lo B-operand of stage 2;  D2 -> global modes)
//   warp  4    MMA issuer (one thread) + TMEM allocation
//   warps 5-12 loaders   (LDG.128 -> bf16 hi/lo split -> swizzled STS into the slab ring)
// =====================================================================================================
constexpr int FA_LOADER_WARPS = 8;
constexpr int FA_THREADS = (4 + 1 + FA_LOADER_WARPS) * 32;   // 416
constexpr int FA_SLAB_BYTES = 128 * 128;                      // one [128 x 64] bf16 slab
constexpr int FA_STAGE_BYTES = 2 * FA_SLAB_BYTES;             // hi + lo

struct AnaParams {
  const float* x;
  float2* out;
  const uint8_t* b1_img;   // [2*N1 x W] bf16, canonical K-major SW128 image (T1 rows then T2 rows)
  const uint8_t* a2_img;   // [128 x 256] bf16 image of the real-embedded leading-dim table (T1 rows 0-63, T2 rows 64-127)
  int n_tiles, W, slabs, N1, KX, QROWS, n_stages, tmem_cols;
  uint32_t off_b1, off_a2, off_b2, off_scratch;
};

template <int N1>
__global__ void __launch_bounds__(FA_THREADS, 1) k_fused_analysis(const AnaParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms need 1024-byte alignment
  __shared__ uint64_t bar_full[4], bar_empty[4], bar_d1_full[2], bar_d1_empty[2], bar_b2_full, bar_d2_full;
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = P.n_stages;
  uint8_t* s_b1 = smem + P.off_b1;
  uint8_t* s_a2 = smem + P.off_a2;
  uint8_t* s_b2 = smem + P.off_b2;
  float* s_scr = reinterpret_cast<float*>(smem + P.off_scratch);

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&bar_full[i], FA_LOADER_WARPS); mbar_init(&bar_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_d1_full[i], 1); mbar_init(&bar_d1_empty[i], 128); }
    mbar_init(&bar_b2_full, 128);
    mbar_init(&bar_d2_full, 1);
    mbar_init_fence();
  }
  if (warp == 4) tmem_alloc(&tmem_base_slot, (uint32_t)P.tmem_cols);
  // constant operand images: global -> shared, byte for byte
  {
    const int b1_vec = (2 * N1 * P.W * 2) / 16, a2_vec = (128 * 256 * 2) / 16;
    const uint4* g1 = reinterpret_cast<const uint4*>(P.b1_img);
    const uint4* g2 = reinterpret_cast<const uint4*>(P.a2_img);
    uint4* d1 = reinterpret_cast<uint4*>(s_b1);
    uint4* d2 = reinterpret_cast<uint4*>(s_a2);
    for (int i = tid; i < b1_vec; i += FA_THREADS) d1[i] = __ldg(g1 + i);
    for (int i = tid; i < a2_vec; i += FA_THREADS) d2[i] = __ldg(g2 + i);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t tm_d1[2] = {tmem, tmem + (uint32_t)(2 * N1)};
  const uint32_t tm_d2 = tmem + (uint32_t)(4 * N1);

  const int n_local = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp >= 5) {
    // ------------------------------------------------------------------ loaders
    const int lt = tid - 5 * 32;                 // 0..255
    const int rbase = lt >> 4, c4 = lt & 15;     // 16 float4 per 64-float row segment; 16 rows per pass
    uint32_t g = 0;                              // running slab counter
    float4 v[8];
    auto issue = [&](int tile, int slab) {
      const float* src = P.x + ((size_t)tile * 128) * P.W + slab * 64 + c4 * 4;
#pragma unroll
      for (int it = 0; it < 8; ++it) v[it] = __ldg(reinterpret_cast<const float4*>(src + (size_t)(rbase + it * 16) * P.W));
    };
    const int total = n_local * P.slabs;
    if (total > 0) issue((int)blockIdx.x, 0);
    for (int idx = 0; idx < total; ++idx, ++g) {
      const int slot = (int)(g % (uint32_t)NS);
      const uint32_t ph = (g / (uint32_t)NS) & 1u;
      float4 cur[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) cur[it] = v[it];
      if (idx + 1 < total) {                     // keep the next slab's loads in flight while this one is converted
        const int nidx = idx + 1;
        issue((int)blockIdx.x + (nidx / P.slabs) * (int)gridDim.x, nidx % P.slabs);
      }
      mbar_wait(&bar_empty[slot], ph ^ 1u);
      uint8_t* hi = smem + (size_t)slot * FA_STAGE_BYTES;
      uint8_t* lo = hi + FA_SLAB_BYTES;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = rbase + it * 16;
        float h0, h1, h2, h3, l0, l1, l2, l3;
        split_bf16(cur[it].x, h0, l0); split_bf16(cur[it].y, h1, l1);
        split_bf16(cur[it].z, h2, l2); split_bf16(cur[it].w, h3, l3);
        const uint32_t off = sw128_offset(r, c4 * 4, 128);
        *reinterpret_cast<uint2*>(hi + off) = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
        *reinterpret_cast<uint2*>(lo + off) = make_uint2(pack_bf16(l0, l1), pack_bf16(l2, l3));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[slot]);
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_p1 = idesc_bf16(128, 2 * N1), idesc_p2 = idesc_bf16(128, N1);
      const uint32_t a_b1 = smem_u32(s_b1), a_a2 = smem_u32(s_a2), a_b2 = smem_u32(s_b2);
      uint32_t g = 0;
      auto stage2 = [&](int j) {   // leading-dim pass of local tile j
        mbar_wait(&bar_b2_full, (uint32_t)(j & 1));
        tc_fence_after_sync();
#pragma unroll 1
        for (int ks = 0; ks < 16; ++ks) {
          const int slab = ks >> 2, kk = ks & 3;
          mma_bf16_ss(tm_d2, smem_desc_sw128(a_a2 + slab * (128 * 128) + kk * 32),
                      smem_desc_sw128(a_b2 + slab * (N1 * 128) + kk * 32), idesc_p2, ks > 0);
        }
        mma_commit(&bar_d2_full);
      };
      for (int i = 0; i < n_local; ++i) {
        const int buf = i & 1;
        mbar_wait(&bar_d1_empty[buf], (uint32_t)(((i >> 1) & 1) ^ 1));
        tc_fence_after_sync();
        for (int s = 0; s < P.slabs; ++s, ++g) {
          const int slot = (int)(g % (uint32_t)NS);
          mbar_wait(&bar_full[slot], (g / (uint32_t)NS) & 1u);
          tc_fence_after_sync();
          const uint32_t a_hi = smem_u32(smem + (size_t)slot * FA_STAGE_BYTES), a_lo = a_hi + FA_SLAB_BYTES;
          const uint32_t b_sl = a_b1 + s * (2 * N1 * 128);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            mma_bf16_ss(tm_d1[buf], smem_desc_sw128(a_hi + kk * 32), smem_desc_sw128(b_sl + kk * 32), idesc_p1, (s | kk) != 0);
            mma_bf16_ss(tm_d1[buf], smem_desc_sw128(a_lo + kk * 32), smem_desc_sw128(b_sl + kk * 32), idesc_p2, true);
          }
          mma_commit(&bar_empty[slot]);
        }
        mma_commit(&bar_d1_full[buf]);
        if (i >= 1) stage2(i - 1);
      }
      if (n_local >= 1) stage2(n_local - 1);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-3 = TMEM lane quarters)
    const int row = warp * 32 + lane;                       // TMEM lane == tile row h (stage 1) / output row i' (stage 2)
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    const int KX = P.KX;
    constexpr int half = N1 / 2;
    auto epi2 = [&](int j) {   // D2 -> modes of local tile j
      mbar_wait(&bar_d2_full, (uint32_t)(j & 1));
      tc_fence_after_sync();
      float acc[N1 / 2];
#pragma unroll
      for (int c = 0; c < N1; c += 16) {
        float t[16];
        tmem_ld16(tm_d2 + lane_sel + c, t);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int col = c + e;                 // hi block [0, half), lo block [half, N1)
          if (col < half) acc[col] = t[e];
          else acc[col - half] += t[e];
        }
      }
      tc_fence_before_sync();
      // T2 rows (warps 2,3) hand their partial sums to the matching T1 rows (warps 0,1)
      if (warp >= 2) {
        float* dst = s_scr + (row - 64) * (KX + 1);
#pragma unroll
        for (int kx = 0; kx < N1 / 2; ++kx) if (kx < KX) dst[kx] = acc[kx];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp < 2) {
        const float* src = s_scr + row * (KX + 1);
        const int tile = (int)blockIdx.x + j * (int)gridDim.x;
        const int q = row >> 1, part = row & 1;
        const bool live = q < P.QROWS;
        float2* dst = P.out + ((size_t)tile * P.QROWS + q) * KX;
#pragma unroll
        for (int kx = 0; kx < N1 / 2; ++kx) {
          if (kx < KX) {                         // warp-uniform
            const float mine = acc[kx] + src[kx];
            const float other = __shfl_xor_sync(0xffffffffu, mine, 1);
            if (live && part == 0) dst[kx] = make_float2(mine, other);
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };
    for (int i = 0; i < n_local; ++i) {
      if (i >= 1) epi2(i - 1);
      const int buf = i & 1;
      mbar_wait(&bar_d1_full[buf], (uint32_t)((i >> 1) & 1));
      tc_fence_after_sync();
      float r[N1];             // R[h, j] = T1 block + T2 block
#pragma unroll
      for (int c = 0; c < 2 * N1; c += 16) {
        float t[16];
        tmem_ld16(tm_d1[buf] + lane_sel + c, t);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int col = c + e;
          if (col < N1) r[col] = t[e];
          else r[col - N1] += t[e];
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&bar_d1_empty[buf]);
      // B operand of stage 2: row n = kx (hi) / half + kx (lo), column k2 = part*128 + h, K-major SW128, N1 rows per slab
#pragma unroll
      for (int j = 0; j < N1; ++j) {
        if (j < 2 * KX) {
          const int kx = j >> 1, part = j & 1;
          float hi_f, lo_f;
          split_bf16(r[j], hi_f, lo_f);
          const int k2 = part * 128 + row;
          *reinterpret_cast<__nv_bfloat16*>(s_b2 + sw128_offset(kx, k2, N1)) = __float2bfloat16_rn(hi_f);
          *reinterpret_cast<__nv_bfloat16*>(s_b2 + sw128_offset(half + kx, k2, N1)) = __float2bfloat16_rn(lo_f);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&bar_b2_full);
    }
    if (n_local >= 1) epi2(n_local - 1);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, (uint32_t)P.tmem_cols);
}

// =====================================================================================================
// fused synthesis:  kept modes of the G images of a tile  ->  128 image rows (+ bias)
//
//   stage A (leading dim) DA[hl, n] = sum_k AA[hl, k] * BA[n, k]      M=128 (rows)  N=2*N1  K=128 (T1 | T2 halves)
//   stage B (last dim)    DB[hl, w] = sum_j U[hl, j] * TS[j, w]       M=128         N=W     K=N1 x 3 bf16 products
//
//   warps 0-3  epilogue B (TMEM -> +bias -> 256-bit global stores of the image rows)
//   warps 4-7  epilogue A (TMEM -> U -> bf16 hi/lo A-operand of stage B)
//   warp  8    MMA issuer + TMEM allocation
//   warps 9-12 prep      (modes -> bf16 hi/lo real-embedded B-operand of stage A)
// =====================================================================================================
constexpr int FS_THREADS = 13 * 32;

struct SynParams {
  const float2* modes;
  float* out;
  const float* bias;       // may be null
  const uint8_t* aa_img;   // [128 x 128] bf16 image: leading-dim table, columns (2q+s | 64+2q+s)
  const uint8_t* bb_img;   // two [W x 64] bf16 images: T1 then T2 of the last-dim table (rows = w, K = j)
  int n_tiles, W, KX, QROWS, H, n_channels, tmem_cols;
  uint32_t off_aa, off_ba, off_u, off_bb;
};

__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

template <int N1>
__global__ void __launch_bounds__(FS_THREADS, 1) k_fused_synthesis(const SynParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_ba_full[2], bar_ba_empty[2], bar_da_full[2], bar_da_empty[2];
  __shared__ uint64_t bar_u_full[2], bar_u_empty[2], bar_db_full[2], bar_db_empty[2];
  __shared__ uint32_t tmem_base_slot;
  constexpr int BA_BYTES = 2 * N1 * 256;    // [2*N1 x 128] bf16 = two slabs of 2*N1 rows
  constexpr int U_BYTES = 2 * FA_SLAB_BYTES;  // hi slab + lo slab, [128 x 64] bf16 each

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int W = P.W, KX = P.KX;
  uint8_t* s_aa = smem + P.off_aa;
  uint8_t* s_ba = smem + P.off_ba;
  uint8_t* s_u = smem + P.off_u;
  uint8_t* s_bb = smem + P.off_bb;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_ba_full[i], 4);  mbar_init(&bar_ba_empty[i], 1);
      mbar_init(&bar_da_full[i], 1);  mbar_init(&bar_da_empty[i], 128);
      mbar_init(&bar_u_full[i], 128); mbar_init(&bar_u_empty[i], 1);
      mbar_init(&bar_db_full[i], 1);  mbar_init(&bar_db_empty[i], 128);
    }
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(&tmem_base_slot, (uint32_t)P.tmem_cols);
  {
    const uint4* g1 = reinterpret_cast<const uint4*>(P.aa_img);
    const uint4* g2 = reinterpret_cast<const uint4*>(P.bb_img);
    uint4* d1 = reinterpret_cast<uint4*>(s_aa);
    uint4* d2 = reinterpret_cast<uint4*>(s_bb);
    for (int i = tid; i < (128 * 128 * 2) / 16; i += FS_THREADS) d1[i] = __ldg(g1 + i);
    for (int i = tid; i < (2 * W * 128) / 16; i += FS_THREADS) d2[i] = __ldg(g2 + i);
    uint4* z1 = reinterpret_cast<uint4*>(s_ba);     // zero both BA buffers and both U buffers once: padding rows /
    for (int i = tid; i < (2 * BA_BYTES) / 16; i += FS_THREADS) z1[i] = make_uint4(0, 0, 0, 0);   // columns stay zero
    uint4* z2 = reinterpret_cast<uint4*>(s_u);
    for (int i = tid; i < (2 * U_BYTES) / 16; i += FS_THREADS) z2[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t tm_da[2] = {tmem, tmem + (uint32_t)(2 * N1)};
  const uint32_t tm_db[2] = {tmem + (uint32_t)(4 * N1), tmem + (uint32_t)(4 * N1 + W)};
  const int n_local = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp >= 9) {
    // ------------------------------------------------------------------ prep: modes -> BA
    const int pt = tid - 9 * 32;   // 0..127
    const int n_el = P.QROWS * KX;
    for (int i = 0; i < n_local; ++i) {
      const int buf = i & 1;
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const float2* src = P.modes + (size_t)tile * n_el;
      mbar_wait(&bar_ba_empty[buf], (uint32_t)(((i >> 1) & 1) ^ 1));
      uint8_t* ba = s_ba + buf * BA_BYTES;
      for (int e = pt; e < n_el; e += 128) {
        const float2 y = __ldg(src + e);
        const int q = e / KX, kx = e - q * KX;
        float rh, rl, ih, il;
        split_bf16(y.x, rh, rl);
        split_bf16(y.y, ih, il);
        const uint32_t re_row_hi = pack_bf16(rh, -ih), im_row_hi = pack_bf16(ih, rh);
        const uint32_t re_row_lo = pack_bf16(rl, -il), im_row_lo = pack_bf16(il, rl);
        const int n0 = 2 * kx, n1 = 2 * kx + 1, k = 2 * q;
        *reinterpret_cast<uint32_t*>(ba + sw128_offset(n0, k, 2 * N1)) = re_row_hi;        // hi * T1
        *reinterpret_cast<uint32_t*>(ba + sw128_offset(n1, k, 2 * N1)) = im_row_hi;
        *reinterpret_cast<uint32_t*>(ba + sw128_offset(n0, 64 + k, 2 * N1)) = re_row_hi;   // hi * T2
        *reinterpret_cast<uint32_t*>(ba + sw128_offset(n1, 64 + k, 2 * N1)) = im_row_hi;
        *reinterpret_cast<uint32_t*>(ba + sw128_offset(N1 + n0, k, 2 * N1)) = re_row_lo;   // lo * T1
        *reinterpret_cast<uint32_t*>(ba + sw128_offset(N1 + n1, k, 2 * N1)) = im_row_lo;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_ba_full[buf]);
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_a = idesc_bf16(128, 2 * N1), idesc_b = idesc_bf16(128, W);
      const uint32_t a_aa = smem_u32(s_aa), a_ba = smem_u32(s_ba), a_u = smem_u32(s_u), a_bb = smem_u32(s_bb);
      auto stage_b = [&](int j) {
        const int buf = j & 1;
        const uint32_t ph = (uint32_t)((j >> 1) & 1);
        mbar_wait(&bar_u_full[buf], ph);
        mbar_wait(&bar_db_empty[buf], ph ^ 1u);
        tc_fence_after_sync();
        const uint32_t u_hi = a_u + buf * U_BYTES, u_lo = u_hi + FA_SLAB_BYTES;
        const uint32_t t1 = a_bb, t2 = a_bb + (uint32_t)W * 128;
#pragma unroll
        for (int ks = 0; ks < N1 / 16; ++ks) {
          mma_bf16_ss(tm_db[buf], smem_desc_sw128(u_hi + ks * 32), smem_desc_sw128(t1 + ks * 32), idesc_b, ks > 0);
          mma_bf16_ss(tm_db[buf], smem_desc_sw128(u_lo + ks * 32), smem_desc_sw128(t1 + ks * 32), idesc_b, true);
          mma_bf16_ss(tm_db[buf], smem_desc_sw128(u_hi + ks * 32), smem_desc_sw128(t2 + ks * 32), idesc_b, true);
        }
        mma_commit(&bar_u_empty[buf]);
        mma_commit(&bar_db_full[buf]);
      };
      for (int i = 0; i < n_local; ++i) {
        const int buf = i & 1;
        const uint32_t ph = (uint32_t)((i >> 1) & 1);
        mbar_wait(&bar_ba_full[buf], ph);
        mbar_wait(&bar_da_empty[buf], ph ^ 1u);
        tc_fence_after_sync();
        const uint32_t ba = a_ba + buf * BA_BYTES;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const int slab = ks >> 2, kk = ks & 3;
          mma_bf16_ss(tm_da[buf], smem_desc_sw128(a_aa + slab * (128 * 128) + kk * 32),
                      smem_desc_sw128(ba + slab * (2 * N1 * 128) + kk * 32), idesc_a, ks > 0);
        }
        mma_commit(&bar_ba_empty[buf]);
        mma_commit(&bar_da_full[buf]);
        if (i >= 1) stage_b(i - 1);
      }
      if (n_local >= 1) stage_b(n_local - 1);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue A: DA -> U (hi / lo)
    const int q4 = warp - 4;
    const int row = q4 * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    for (int i = 0; i < n_local; ++i) {
      const int buf = i & 1;
      const uint32_t ph = (uint32_t)((i >> 1) & 1);
      mbar_wait(&bar_da_full[buf], ph);
      tc_fence_after_sync();
      float u[N1];
#pragma unroll
      for (int c = 0; c < 2 * N1; c += 16) {
        float t[16];
        tmem_ld16(tm_da[buf] + lane_sel + c, t);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int col = c + e;
          if (col < N1) u[col] = t[e];
          else u[col - N1] += t[e];
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&bar_da_empty[buf]);
      mbar_wait(&bar_u_empty[buf], ph ^ 1u);
      uint8_t* uhi = s_u + buf * U_BYTES;
      uint8_t* ulo = uhi + FA_SLAB_BYTES;
#pragma unroll
      for (int c = 0; c < N1 / 8; ++c) {          // one 16-byte chunk = 8 consecutive j
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float h0, l0, h1, l1;
          split_bf16(u[c * 8 + 2 * e], h0, l0);
          split_bf16(u[c * 8 + 2 * e + 1], h1, l1);
          hw[e] = pack_bf16(h0, h1);
          lw[e] = pack_bf16(l0, l1);
        }
        const uint32_t off = (uint32_t)(row * 128 + (((c ^ row) & 7) << 4));
        *reinterpret_cast<uint4*>(uhi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(ulo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(&bar_u_full[buf]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue B: DB -> image rows
    const int row = warp * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    for (int i = 0; i < n_local; ++i) {
      const int buf = i & 1;
      const uint32_t ph = (uint32_t)((i >> 1) & 1);
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      float b = 0.f;
      if (P.bias != nullptr) {
        const long long image = (long long)tile * (128 / P.H) + row / P.H;
        b = __ldg(P.bias + (int)(image % P.n_channels));
      }
      float* dst = P.out + ((size_t)tile * 128 + row) * W;
      mbar_wait(&bar_db_full[buf], ph);
      tc_fence_after_sync();
      for (int c = 0; c < W; c += 32) {
        float t0[16], t1[16];
        tmem_ld16(tm_db[buf] + lane_sel + c, t0);
        tmem_ld16(tm_db[buf] + lane_sel + c + 16, t1);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) { t0[e] += b; t1[e] += b; }
        st_global_v8(dst + c, t0);
        st_global_v8(dst + c + 8, t0 + 8);
        st_global_v8(dst + c + 16, t1);
        st_global_v8(dst + c + 24, t1 + 8);
      }
      tc_fence_before_sync();
      mbar_arrive(&bar_db_empty[buf]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, (uint32_t)P.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------
// host side: operand images and dispatch
// ---------------------------------------------------------------------------------------------------------
struct FusedAnalysisTables {
  bool ok = false;
  int W = 0, H = 0, G = 0, N1 = 0, KX = 0, KY = 0, slabs = 0, n_stages = 0, tmem_cols = 0;
  uint8_t* d_b1 = nullptr;
  uint8_t* d_a2 = nullptr;
  uint32_t off_b1 = 0, off_a2 = 0, off_b2 = 0, off_scratch = 0, smem_bytes = 0;
};

struct FusedSynthesisTables {
  bool ok = false;
  int W = 0, H = 0, G = 0, N1 = 0, KX = 0, KY = 0, tmem_cols = 0;
  uint8_t* d_aa = nullptr;
  uint8_t* d_bb = nullptr;
  uint32_t off_aa = 0, off_ba = 0, off_u = 0, off_bb = 0, smem_bytes = 0;
};

struct FastTables {
  FusedAnalysisTables ana[2];   // [0] forward analysis on `grid`, [1] adjoint-of-synthesis analysis on `out_grid`
  FusedSynthesisTables syn[2];  // [0] forward synthesis onto `out_grid`, [1] adjoint-of-analysis synthesis onto `grid`
  int sm_count = 0;
};

static inline uint16_t bf16_bits(float f) {   // round to nearest even
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf16_to_float(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline size_t host_sw128_offset(int r, int k, int rows) {
  const int slab = k >> 6, kk = k & 63;
  return (size_t)slab * rows * 128 + (size_t)r * 128 + ((((kk >> 3) ^ r) & 7) << 4) + ((kk & 7) << 1);
}
// writes v as T1 (row r1) and T2 (row r2) of a [rows x K] K-major SW128 bf16 image
static inline void put_split(std::vector<uint8_t>& img, int rows, int r1, int r2, int k, float v) {
  const uint16_t t1 = bf16_bits(v);
  const uint16_t t2 = bf16_bits(v - bf16_to_float(t1));
  memcpy(&img[host_sw128_offset(r1, k, rows)], &t1, 2);
  memcpy(&img[host_sw128_offset(r2, k, rows)], &t2, 2);
}

template <typename T>
static bool upload_bytes(Plan* p, const std::vector<T>& host, uint8_t** dev) {
  void* d = nullptr;
  if (!cuda_ok(cudaMalloc(&d, host.size() * sizeof(T)), "cudaMalloc(fast table)")) return false;
  p->owned.push_back(d);
  if (!cuda_ok(cudaMemcpy(d, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice), "cudaMemcpy(fast table)")) return false;
  *dev = static_cast<uint8_t*>(d);
  return true;
}

// last-dim table `tab` is [W x 2KX] (row stride 2KX); leading-dim table `lead` is [KY x H] complex
static bool build_fused_analysis(Plan* p, FusedAnalysisTables* t, int H, int W, int KY, int KX, const std::vector<float>& tab,
                                 const std::vector<float2>& lead) {
  t->ok = false;
  if (W % 64 != 0 || W > 128 || H < 16 || 128 % H != 0) return true;
  const int G = 128 / H;
  const int N1 = ((2 * KX + 15) / 16) * 16;
  if (N1 > 96 || G * KY > 32 || KX < 1 || KY < 1) return true;
  t->W = W; t->H = H; t->G = G; t->N1 = N1; t->KX = KX; t->KY = KY; t->slabs = W / 64;
  t->tmem_cols = 5 * N1 <= 256 ? 256 : 512;
  // ---- B1: [2*N1 x W]
  std::vector<uint8_t> b1((size_t)2 * N1 * W * 2, 0);
  for (int j = 0; j < 2 * KX; ++j)
    for (int w = 0; w < W; ++w) put_split(b1, 2 * N1, j, N1 + j, w, tab[(size_t)w * 2 * KX + j]);
  // ---- A2: [128 x 256]; row i' = 2*(g*KY + ky) + p_out (T1), 64 + i' (T2); column k2 = p_in*128 + g*H + h
  std::vector<uint8_t> a2((size_t)128 * 256 * 2, 0);
  for (int g = 0; g < G; ++g)
    for (int ky = 0; ky < KY; ++ky)
      for (int h = 0; h < H; ++h) {
        const float2 f = lead[(size_t)ky * H + h];
        const int q = g * KY + ky, hl = g * H + h;
        put_split(a2, 128, 2 * q + 0, 64 + 2 * q + 0, 0 * 128 + hl, f.x);
        put_split(a2, 128, 2 * q + 0, 64 + 2 * q + 0, 1 * 128 + hl, -f.y);
        put_split(a2, 128, 2 * q + 1, 64 + 2 * q + 1, 0 * 128 + hl, f.y);
        put_split(a2, 128, 2 * q + 1, 64 + 2 * q + 1, 1 * 128 + hl, f.x);
      }
  if (!upload_bytes(p, b1, &t->d_b1) || !upload_bytes(p, a2, &t->d_a2)) return false;
  // ---- shared-memory carve-up
  const uint32_t fixed = (uint32_t)b1.size() + 65536u + (uint32_t)N1 * 512u + ((64u * (KX + 1) * 4u + 1023u) & ~1023u);
  int stages = (int)((227u * 1024u - 4096u - fixed) / FA_STAGE_BYTES);
  if (stages > 4) stages = 4;
  if (stages < 2) return true;
  t->n_stages = stages;
  t->off_b1 = (uint32_t)stages * FA_STAGE_BYTES;
  t->off_a2 = t->off_b1 + (uint32_t)b1.size();
  t->off_b2 = t->off_a2 + 65536u;
  t->off_scratch = t->off_b2 + (uint32_t)N1 * 512u;
  t->smem_bytes = t->off_scratch + ((64u * (KX + 1) * 4u + 1023u) & ~1023u) + 1024u;
  t->ok = true;
  return true;
}

// last-dim table `tab` is [2KX x W] (row stride W); leading-dim table `lead` is [H x KY] complex
static bool build_fused_synthesis(Plan* p, FusedSynthesisTables* t, int H, int W, int KY, int KX, const std::vector<float>& tab,
                                  const std::vector<float2>& lead) {
  t->ok = false;
  if (W % 64 != 0 || W > 256 || H < 16 || 128 % H != 0) return true;
  const int G = 128 / H;
  const int N1 = ((2 * KX + 15) / 16) * 16;
  if (N1 > 64 || G * KY > 32 || KX < 1 || KY < 1 || 4 * N1 + 2 * W > 512) return true;
  t->W = W; t->H = H; t->G = G; t->N1 = N1; t->KX = KX; t->KY = KY;
  t->tmem_cols = 4 * N1 + 2 * W <= 256 ? 256 : 512;
  // ---- AA: [128 x 128]; row hl = g*H + h; columns 2q+s (T1) and 64+2q+s (T2), q = g*KY + ky, s: 0 = Re, 1 = Im of the table
  std::vector<uint8_t> aa((size_t)128 * 128 * 2, 0);
  {
    // T1 and T2 of the same coefficient live in the SAME row at columns k and 64 + k
    for (int g = 0; g < G; ++g)
      for (int h = 0; h < H; ++h)
        for (int ky = 0; ky < KY; ++ky) {
          const float2 f = lead[(size_t)h * KY + ky];
          const int hl = g * H + h, q = g * KY + ky;
          const float vals[2] = {f.x, f.y};
          for (int sgn = 0; sgn < 2; ++sgn) {
            const uint16_t t1 = bf16_bits(vals[sgn]);
            const uint16_t t2 = bf16_bits(vals[sgn] - bf16_to_float(t1));
            memcpy(&aa[host_sw128_offset(hl, 2 * q + sgn, 128)], &t1, 2);
            memcpy(&aa[host_sw128_offset(hl, 64 + 2 * q + sgn, 128)], &t2, 2);
          }
        }
  }
  // ---- BB: T1 image then T2 image, each [W x 64]: row w, column j
  std::vector<uint8_t> bb((size_t)2 * W * 128, 0);
  for (int j = 0; j < 2 * KX; ++j)
    for (int w = 0; w < W; ++w) {
      const float v = tab[(size_t)j * W + w];
      const uint16_t t1 = bf16_bits(v);
      const uint16_t t2 = bf16_bits(v - bf16_to_float(t1));
      memcpy(&bb[host_sw128_offset(w, j, W)], &t1, 2);
      memcpy(&bb[(size_t)W * 128 + host_sw128_offset(w, j, W)], &t2, 2);
    }
  if (!upload_bytes(p, aa, &t->d_aa) || !upload_bytes(p, bb, &t->d_bb)) return false;
  t->off_aa = 0;
  t->off_ba = 32768u;
  t->off_u = t->off_ba + 2u * (uint32_t)(2 * N1 * 256);
  t->off_bb = t->off_u + 2u * (uint32_t)(2 * FA_SLAB_BYTES);
  t->smem_bytes = t->off_bb + (uint32_t)bb.size() + 1024u;
  if (t->smem_bytes > 227u * 1024u - 4096u) return true;
  t->ok = true;
  return true;
}

bool fast_plan_init(Plan* p) {
  p->fast = nullptr;
  if (p->d < 2) return true;
  cudaDeviceProp prop{};
  if (!cuda_ok(cudaGetDeviceProperties(&prop, p->device), "cudaGetDeviceProperties")) return false;
  if (prop.major != 10) return true;   // tcgen05 path is sm_100-only
  FastTables* f = new FastTables();
  f->sm_count = prop.multiProcessorCount;
  const DimTables& L = p->dim[p->d - 1];
  const DimTables& Y = p->dim[p->d - 2];
  bool good = build_fused_analysis(p, &f->ana[0], Y.N, L.N, Y.k, L.k, p->h_TA, Y.h_A) &&
              build_fused_analysis(p, &f->ana[1], Y.M, L.M, Y.k, L.k, p->h_TST, Y.h_SH) &&
              build_fused_synthesis(p, &f->syn[0], Y.M, L.M, Y.k, L.k, p->h_TS, Y.h_S) &&
              build_fused_synthesis(p, &f->syn[1], Y.N, L.N, Y.k, L.k, p->h_TAT, Y.h_AH);
  if (!good) { delete f; return false; }
  p->fast = f;
  return true;
}

void fast_plan_destroy(Plan* p) {
  delete p->fast;
  p->fast = nullptr;
}

bool fast_can_analyze(const Plan* p, bool adjoint) {
  return p->fast != nullptr && p->d == 2 && p->fast->ana[adjoint ? 1 : 0].ok;
}
bool fast_can_synthesize(const Plan* p, bool adjoint) {
  return p->fast != nullptr && p->d == 2 && p->fast->syn[adjoint ? 1 : 0].ok;
}

bool fast_analyze(const Plan* p, const float* images, int64_t n_images, float2* modes_out, bool adjoint, cudaStream_t st) {
  const FusedAnalysisTables& t = p->fast->ana[adjoint ? 1 : 0];
  if (n_images % t.G != 0) { set_error("fast_analyze: image count not a multiple of the tile group"); return false; }
  AnaParams P{};
  P.x = images; P.out = modes_out; P.b1_img = t.d_b1; P.a2_img = t.d_a2;
  P.n_tiles = (int)(n_images / t.G); P.W = t.W; P.slabs = t.slabs; P.N1 = t.N1; P.KX = t.KX; P.QROWS = t.G * t.KY;
  P.n_stages = t.n_stages; P.tmem_cols = t.tmem_cols;
  P.off_b1 = t.off_b1; P.off_a2 = t.off_a2; P.off_b2 = t.off_b2; P.off_scratch = t.off_scratch;
  const int grid = P.n_tiles < p->fast->sm_count ? P.n_tiles : p->fast->sm_count;
  switch (t.N1) {
#define SC_FA_CASE(N)                                                                                            \
  case N: {                                                                                                      \
    static uint32_t attr_bytes = 0;                                                                              \
    if (attr_bytes < t.smem_bytes) {                                                                             \
      if (!cuda_ok(cudaFuncSetAttribute(k_fused_analysis<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                        (int)t.smem_bytes), "cudaFuncSetAttribute(k_fused_analysis)"))          \
        return false;                                                                                            \
      attr_bytes = t.smem_bytes;                                                                                 \
    }                                                                                                            \
    k_fused_analysis<N><<<grid, FA_THREADS, t.smem_bytes, st>>>(P);                                              \
  } break;
    SC_FA_CASE(16) SC_FA_CASE(32) SC_FA_CASE(48) SC_FA_CASE(64) SC_FA_CASE(80) SC_FA_CASE(96)
#undef SC_FA_CASE
    default: set_error("fast_analyze: unsupported N1"); return false;
  }
  count_launch();
  return cuda_ok(cudaGetLastError(), "k_fused_analysis launch");
}

bool fast_synthesize(const Plan* p, const float2* modes_in, int64_t n_images, int n_channels, const float* bias,
                     float* images_out, bool adjoint, cudaStream_t st) {
  const FusedSynthesisTables& t = p->fast->syn[adjoint ? 1 : 0];
  if (n_images % t.G != 0) { set_error("fast_synthesize: image count not a multiple of the tile group"); return false; }
  SynParams P{};
  P.modes = modes_in; P.out = images_out; P.bias = bias; P.aa_img = t.d_aa; P.bb_img = t.d_bb;
  P.n_tiles = (int)(n_images / t.G); P.W = t.W; P.KX = t.KX; P.QROWS = t.G * t.KY; P.H = t.H;
  P.n_channels = n_channels > 0 ? n_channels : 1; P.tmem_cols = t.tmem_cols;
  P.off_aa = t.off_aa; P.off_ba = t.off_ba; P.off_u = t.off_u; P.off_bb = t.off_bb;
  const int grid = P.n_tiles < p->fast->sm_count ? P.n_tiles : p->fast->sm_count;
  switch (t.N1) {
#define SC_FS_CASE(N)                                                                                            \
  case N: {                                                                                                      \
    static uint32_t attr_bytes = 0;                                                                              \
    if (attr_bytes < t.smem_bytes) {                                                                             \
      if (!cuda_ok(cudaFuncSetAttribute(k_fused_synthesis<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                        (int)t.smem_bytes), "cudaFuncSetAttribute(k_fused_synthesis)"))         \
        return false;                                                                                            \
      attr_bytes = t.smem_bytes;                                                                                 \
    }                                                                                                            \
    k_fused_synthesis<N><<<grid, FS_THREADS, t.smem_bytes, st>>>(P);                                             \
  } break;
    SC_FS_CASE(16) SC_FS_CASE(32) SC_FS_CASE(48) SC_FS_CASE(64)
#undef SC_FS_CASE
    default: set_error("fast_synthesize: unsupported N1"); return false;
  }
  count_launch();
  return cuda_ok(cudaGetLastError(), "k_fused_synthesis launch");
}

}  // namespace sc

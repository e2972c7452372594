// Tensor-core (warp-level mma.sync m16n8k8 TF32) attention core for the 64-token, 4-head, d_head=62 layer of ATM-S
// (FullAttention.forward, models/subject_layers/SelfAttention_Family.py:56-75).  One CTA per (sample, head), one warp per
// 16 query rows.  The whole 64x64 score tile of a head lives in registers; the probability tile is fed back as the
// A operand of P.V by choosing the reduction-slot order of the MMA (slot t <-> key 2t, slot t+4 <-> key 2t+1), which is
// exactly the accumulator ownership, so no shuffles or shared-memory round trip are needed.
// Plain TF32 (RN) products; a 3xTF32 variant is kept as a template parameter (the legacy mma.sync pipe is the issue
// limit of this kernel, see conv_mma.cu).  Backward: P and dS are exchanged through shared memory for the two
// products that reduce over the query index.
#include "kernels.h"
#include <stdlib.h>

namespace eegb200 {

static constexpr int AL = 68;    // row stride of Q/K/V/dO tiles  (stride = 4 mod 32: conflict-free (row g, col t) gathers)
static constexpr int PL = 72;    // row stride of P/dS tiles       (stride = 8 mod 32: conflict-free (row t, col g) gathers)
static constexpr int AT_THREADS = 128;
static constexpr float QK_SCALE = 0.12700012700019050f;   // 1/sqrt(62)

__device__ __forceinline__ void mma8(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tfb(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
  hi = tfb(x);
  lo = tfb(x - __uint_as_float(hi));
}

// 64 rows x 64 floats of one head -> smem tile with row stride AL (128 threads, float4)
// ROUND: store the values already rounded to TF32 (RN) so that the fragment gathers of the plain-TF32 kernels need no
// cvt per use (2 of the ~5 instructions per MMA in the first version; every tile element is used 16-64 times)
template <bool ROUND>
__device__ __forceinline__ void load_tile(const float* __restrict__ base, int ld, float* __restrict__ dst) {
  const int r0 = threadIdx.x >> 4, c4 = threadIdx.x & 15;
  float4 v[8];
#pragma unroll
  for (int p = 0; p < 8; ++p) v[p] = *reinterpret_cast<const float4*>(base + (size_t)(r0 + 8 * p) * ld + c4 * 4);
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    if (ROUND) { v[p].x = tf32_rn(v[p].x); v[p].y = tf32_rn(v[p].y); v[p].z = tf32_rn(v[p].z); v[p].w = tf32_rn(v[p].w); }
    *reinterpret_cast<float4*>(dst + (r0 + 8 * p) * AL + c4 * 4) = v[p];
  }
}
// operand bits of a value that is either pre-rounded (XP == 1 tiles) or still fp32
template <int XP>
__device__ __forceinline__ uint32_t opb(float x) { return XP == 1 ? __float_as_uint(x) : tfb(x); }

// scores of 16 query rows (warp tile) against 64 keys.  XP = 3 -> 3xTF32, XP = 1 -> TF32.
// s[nt][0..3]: rows (g, g+8) x keys (nt*8+2t, +1)
template <int XP>
__device__ __forceinline__ void qk_scores(const float* __restrict__ Q, const float* __restrict__ K, int row0, int g, int t,
                                          float s[8][4]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll 2
  for (int kt = 0; kt < 8; ++kt) {
    uint32_t ah[4], al[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float v = Q[(row0 + g + 8 * (h & 1)) * AL + kt * 8 + t + 4 * (h >> 1)];
      if (XP == 3) split(v, ah[h], al[h]); else ah[h] = __float_as_uint(v);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float k0 = K[(nt * 8 + g) * AL + kt * 8 + t], k1 = K[(nt * 8 + g) * AL + kt * 8 + t + 4];
      if (XP == 3) {
        uint32_t h0, l0, h1, l1;
        split(k0, h0, l0);
        split(k1, h1, l1);
        mma8(s[nt], al, h0, h1);
        mma8(s[nt], ah, l0, l1);
        mma8(s[nt], ah, h0, h1);
      } else {
        mma8(s[nt], ah, __float_as_uint(k0), __float_as_uint(k1));
      }
    }
  }
}
// in-register softmax over the 64 keys of rows g and g+8 (values spread over the 4 lanes of a quad)
__device__ __forceinline__ void softmax_rows(float s[8][4]) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int q = 0; q < 4; ++q) s[nt][q] *= QK_SCALE;
    m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
    m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float z0 = 0.f, z1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    s[nt][0] = __expf(s[nt][0] - m0); s[nt][1] = __expf(s[nt][1] - m0);
    s[nt][2] = __expf(s[nt][2] - m1); s[nt][3] = __expf(s[nt][3] - m1);
    z0 += s[nt][0] + s[nt][1];
    z1 += s[nt][2] + s[nt][3];
  }
  z0 += __shfl_xor_sync(0xffffffffu, z0, 1); z0 += __shfl_xor_sync(0xffffffffu, z0, 2);
  z1 += __shfl_xor_sync(0xffffffffu, z1, 1); z1 += __shfl_xor_sync(0xffffffffu, z1, 2);
  const float i0 = 1.f / z0, i1 = 1.f / z1;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { s[nt][0] *= i0; s[nt][1] *= i0; s[nt][2] *= i1; s[nt][3] *= i1; }
}
// dropout keep factors for the accumulator layout: element (row, key) index = (bh*64 + row)*64 + key
__device__ __forceinline__ void keep_factors(const DropoutCfg& d, int bh, int row0, int g, int t, float kf[8][4]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const uint64_t idx = ((uint64_t)bh * 64 + row0 + g + 8 * hrow) * 64 + nt * 8 + 2 * t;
      const uint32_t m = dropout_keep4(d, idx & ~(uint64_t)3) >> (idx & 3);
      kf[nt][2 * hrow] = (m & 1u) ? d.scale : 0.f;
      kf[nt][2 * hrow + 1] = (m & 2u) ? d.scale : 0.f;
    }
}
// out[16 x 64] = P[16 x 64 keys] . B[64 keys x 64], P given in accumulator layout, B row-major in smem (stride AL)
template <int XP>
__device__ __forceinline__ void pv_product(const float p[8][4], const float* __restrict__ Bm, int g, int t, float o[8][4]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < 8; ++kt) {
    // reduction slots: t <-> key kt*8+2t, t+4 <-> key kt*8+2t+1  (== accumulator ownership of this lane)
    uint32_t ah[4], al[4];
    const float av[4] = {p[kt][0], p[kt][2], p[kt][1], p[kt][3]};   // (row g, slot t), (row g+8, slot t), (g, t+4), (g+8, t+4)
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      if (XP == 3) split(av[h], ah[h], al[h]); else ah[h] = tfb(av[h]);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float b0 = Bm[(kt * 8 + 2 * t) * AL + nt * 8 + g], b1 = Bm[(kt * 8 + 2 * t + 1) * AL + nt * 8 + g];
      if (XP == 3) {
        uint32_t h0, l0, h1, l1;
        split(b0, h0, l0);
        split(b1, h1, l1);
        mma8(o[nt], al, h0, h1);
        mma8(o[nt], ah, l0, l1);
        mma8(o[nt], ah, h0, h1);
      } else {
        mma8(o[nt], ah, __float_as_uint(b0), __float_as_uint(b1));
      }
    }
  }
}

template <int XP>
__global__ void __launch_bounds__(AT_THREADS) attention_fwd_mma_kernel(const float* __restrict__ qkv, float* __restrict__ o,
                                                                       DropoutCfg drop) {
  extern __shared__ __align__(16) float sm[];
  float* Q = sm;
  float* K = Q + 64 * AL;
  float* V = K + 64 * AL;
  const int b = blockIdx.x >> 2, h = blockIdx.x & 3;
  const float* base = qkv + (size_t)b * 64 * 768 + h * 64;
  load_tile<XP == 1>(base, 768, Q);
  load_tile<XP == 1>(base + 256, 768, K);
  load_tile<XP == 1>(base + 512, 768, V);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int row0 = warp * 16;
  float s[8][4];
  qk_scores<XP>(Q, K, row0, g, t, s);
  softmax_rows(s);
  if (drop.p > 0.f) {
    float kf[8][4];
    keep_factors(drop, blockIdx.x, row0, g, t, kf);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) s[nt][q] *= kf[nt][q];
  }
  float acc[8][4];
  pv_product<XP>(s, V, g, t, acc);
#pragma unroll
  for (int hrow = 0; hrow < 2; ++hrow) {
    float* orow = o + ((size_t)b * 64 + row0 + g + 8 * hrow) * 256 + h * 64;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int e = nt * 8 + 2 * t;
      const float v0 = e < D_HEAD ? tf32_rn(acc[nt][2 * hrow]) : 0.f;
      const float v1 = e + 1 < D_HEAD ? tf32_rn(acc[nt][2 * hrow + 1]) : 0.f;
      *reinterpret_cast<float2*>(orow + e) = make_float2(v0, v1);
    }
  }
}

__global__ void __launch_bounds__(AT_THREADS) attention_bwd_mma_kernel(const float* __restrict__ qkv,
                                                                       const float* __restrict__ d_o,
                                                                       float* __restrict__ dqkv, DropoutCfg drop) {
  extern __shared__ __align__(16) float sm[];
  float* Q = sm;
  float* K = Q + 64 * AL;
  float* V = K + 64 * AL;
  float* DO = V + 64 * AL;
  float* Pd = DO + 64 * AL;     // dropped probabilities [i][j], stride PL
  float* DS = Pd + 64 * PL;     // d scores [i][j], stride PL
  const int b = blockIdx.x >> 2, h = blockIdx.x & 3;
  const float* base = qkv + (size_t)b * 64 * 768 + h * 64;
  load_tile<true>(base, 768, Q);
  load_tile<true>(base + 256, 768, K);
  load_tile<true>(base + 512, 768, V);
  load_tile<true>(d_o + (size_t)b * 64 * 256 + h * 64, 256, DO);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int row0 = warp * 16;
  float p[8][4];
  qk_scores<1>(Q, K, row0, g, t, p);
  softmax_rows(p);
  // dPd[i][j] = dO[i,:] . V[j,:]   (same operand pattern as Q.K^T)
  float dp[8][4];
  qk_scores<1>(DO, V, row0, g, t, dp);
  float kf[8][4];
  if (drop.p > 0.f) keep_factors(drop, blockIdx.x, row0, g, t, kf);
  float r0 = 0.f, r1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float keep = drop.p > 0.f ? kf[nt][q] : 1.f;
      dp[nt][q] *= keep;                      // gradient wrt the un-dropped probability
      if (q < 2) r0 = fmaf(dp[nt][q], p[nt][q], r0); else r1 = fmaf(dp[nt][q], p[nt][q], r1);
    }
  r0 += __shfl_xor_sync(0xffffffffu, r0, 1); r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
  r1 += __shfl_xor_sync(0xffffffffu, r1, 1); r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
  float ds[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    ds[nt][0] = p[nt][0] * (dp[nt][0] - r0) * QK_SCALE;
    ds[nt][1] = p[nt][1] * (dp[nt][1] - r0) * QK_SCALE;
    ds[nt][2] = p[nt][2] * (dp[nt][2] - r1) * QK_SCALE;
    ds[nt][3] = p[nt][3] * (dp[nt][3] - r1) * QK_SCALE;
  }
  // publish Pd and dS for the products that reduce over the query index
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int i = row0 + g + 8 * hrow, j = nt * 8 + 2 * t;
      const float k0 = drop.p > 0.f ? kf[nt][2 * hrow] : 1.f, k1 = drop.p > 0.f ? kf[nt][2 * hrow + 1] : 1.f;
      *reinterpret_cast<float2*>(Pd + i * PL + j) = make_float2(tf32_rn(p[nt][2 * hrow] * k0), tf32_rn(p[nt][2 * hrow + 1] * k1));
      *reinterpret_cast<float2*>(DS + i * PL + j) = make_float2(tf32_rn(ds[nt][2 * hrow]), tf32_rn(ds[nt][2 * hrow + 1]));
    }
  // dQ[i][e] = sum_j dS[i][j] K[j][e]   (rows owned by this warp)
  {
    float dq[8][4];
    pv_product<1>(ds, K, g, t, dq);
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      float* orow = dqkv + ((size_t)b * 64 + row0 + g + 8 * hrow) * 768 + h * 64;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int e = nt * 8 + 2 * t;
        *reinterpret_cast<float2*>(orow + e) = make_float2(e < D_HEAD ? tf32_rn(dq[nt][2 * hrow]) : 0.f,
                                                           e + 1 < D_HEAD ? tf32_rn(dq[nt][2 * hrow + 1]) : 0.f);
      }
    }
  }
  __syncthreads();
  // dV[j][e] = sum_i Pd[i][j] dO[i][e],  dK[j][e] = sum_i dS[i][j] Q[i][e]     (this warp: keys j in [row0, row0+16))
  {
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
#pragma unroll 1
    for (int kt = 0; kt < 8; ++kt) {
      uint32_t ap[4], as[4];
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        const int j = row0 + g + 8 * (hh & 1), i = kt * 8 + t + 4 * (hh >> 1);
        ap[hh] = __float_as_uint(Pd[i * PL + j]);      // stored pre-rounded
        as[hh] = __float_as_uint(DS[i * PL + j]);
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int e = nt * 8 + g;
        mma8(dv[nt], ap, __float_as_uint(DO[(kt * 8 + t) * AL + e]), __float_as_uint(DO[(kt * 8 + t + 4) * AL + e]));
        mma8(dk[nt], as, __float_as_uint(Q[(kt * 8 + t) * AL + e]), __float_as_uint(Q[(kt * 8 + t + 4) * AL + e]));
      }
    }
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      float* orow = dqkv + ((size_t)b * 64 + row0 + g + 8 * hrow) * 768 + h * 64;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int e = nt * 8 + 2 * t;
        const bool v0 = e < D_HEAD, v1 = e + 1 < D_HEAD;
        *reinterpret_cast<float2*>(orow + 256 + e) = make_float2(v0 ? tf32_rn(dk[nt][2 * hrow]) : 0.f, v1 ? tf32_rn(dk[nt][2 * hrow + 1]) : 0.f);
        *reinterpret_cast<float2*>(orow + 512 + e) = make_float2(v0 ? tf32_rn(dv[nt][2 * hrow]) : 0.f, v1 ? tf32_rn(dv[nt][2 * hrow + 1]) : 0.f);
      }
    }
  }
}

int attention_fwd_simt(const float* qkv, float* o, int B, DropoutCfg drop, cudaStream_t s);
int attention_bwd_simt(const float* qkv, const float* d_o, float* dqkv, int B, DropoutCfg drop, cudaStream_t s);

int attention_fwd(const float* qkv, float* o, int B, DropoutCfg drop, cudaStream_t s) {
  if (!tf32_rounding()) return attention_fwd_simt(qkv, o, B, drop, s);     // exact-fp32 verification path
  if (attention_tc_enabled()) return attention_fwd_tc(qkv, o, B, drop, s); // tcgen05 path (default)
  ProfScope _ps("attention_fwd", s, (double)B * 4 * 4.0 * 64 * 64 * 62, (double)B * 64 * 1024 * 4.0);
  const size_t smem = 3 * 64 * AL * sizeof(float);
  static int xp = -1;
  if (xp < 0) {
    const char* e = getenv("EEGB200_ATTN_XP");
    xp = (e && e[0] == '3') ? 3 : 1;
  }
  static PerDeviceOnce once_fwd;
  if (once_fwd.first()) {
    EEG_CUDA_OK(cudaFuncSetAttribute(attention_fwd_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EEG_CUDA_OK(cudaFuncSetAttribute(attention_fwd_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (xp == 3) attention_fwd_mma_kernel<3><<<B * N_HEAD, AT_THREADS, smem, s>>>(qkv, o, drop);
  else attention_fwd_mma_kernel<1><<<B * N_HEAD, AT_THREADS, smem, s>>>(qkv, o, drop);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int attention_bwd(const float* qkv, const float* d_o, float* dqkv, int B, DropoutCfg drop, cudaStream_t s) {
  if (!tf32_rounding()) return attention_bwd_simt(qkv, d_o, dqkv, B, drop, s);
  ProfScope _ps("attention_bwd", s, (double)B * 4 * 12.0 * 64 * 64 * 62, (double)B * 64 * 1792 * 4.0);
  const size_t smem = (4 * 64 * AL + 2 * 64 * PL) * sizeof(float);
  static PerDeviceOnce once;
  if (once.first()) {
    EEG_CUDA_OK(cudaFuncSetAttribute(attention_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  attention_bwd_mma_kernel<<<B * N_HEAD, AT_THREADS, smem, s>>>(qkv, d_o, dqkv, drop);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace eegb200

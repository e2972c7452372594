// Tensor-core (warp-level mma.sync m16n8k8 TF32) versions of the temporal conv of PatchEmbedding
// (Retrieval/ATMS_retrieval.py:102-103: Conv2d(1->40,(1,25)) o AvgPool((1,51),/5)) and of its backward.
// The conv acts independently on each (sample, token-row) of 250 samples; per row it is three tiny matmuls
//   forward : Y[j][k]   = sum_i  P[j][i] * w[k][i]/51          (36 x 25 x 40,   P[j][i] = box51(x)[5j+i])
//   dW      : dW[k][i] += sum_j  dY[j][k] * P[j][i]            (40 x 36 x 25)
//   dP      : dp[5m+rho] = sum_{a<5,k} dY[m-a][k] * w[k][rho+5a]/51   (1-D transposed conv written as a GEMM over (a,k))
// so one warp owns a row, keeps the weight fragments (and the dW accumulators) in registers and streams the row
// through shared memory.  The hand-written tcgen05 path is reserved for the large GEMMs; these row problems are far
// below one 128-row UMMA tile.  Forward: 3xTF32 split (hi.hi + hi.lo + lo.hi).  The conv output feeds a train-mode BatchNorm that
// subtracts the batch mean, so rounding errors count relative to the CENTRED signal: with plain TF32 the train-mode
// embedding error was 1.0e-3 (at the parity gate), with the split 6.5e-4 (tools/gpu_error_budget.py); cost ~55 us
// per step (the legacy mma.sync pipe issues one m16n8k8 per ~16 cycles per SM sub-partition).  Backward: plain TF32.
#include "kernels.h"
#include <stdlib.h>

namespace eegb200 {

static constexpr int CW_WARPS = 4;
static constexpr int CW_THREADS = CW_WARPS * 32;
static constexpr int XS_LEN = 320;     // row of 250 + zero tail for the sliding box sums
static constexpr int PS_LEN = 272;     // 200 pooled sums + zero tail (A/B fragment gathers touch up to 5*47+31)

__device__ __forceinline__ void mma_tf32(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32_bits(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}

// load one token row (256 floats, pad zero) and build the 200 box-51 sums with a sliding window (7 per lane)
__device__ __forceinline__ void warp_load_row_pool(const float* __restrict__ xrow, float* xs, float* ps, int lane) {
  const float4* src = reinterpret_cast<const float4*>(xrow);
  float4 v0 = src[lane], v1 = src[lane + 32];
  reinterpret_cast<float4*>(xs)[lane] = v0;
  reinterpret_cast<float4*>(xs)[lane + 32] = v1;
  __syncwarp();
  const int s0 = lane * 7;
  float a = 0.f;
#pragma unroll
  for (int v = 0; v < K_POOL; ++v) a += xs[s0 + v];
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    if (s0 + q < N_PSUM) ps[s0 + q] = a;
    a += xs[s0 + q + K_POOL] - xs[s0 + q];
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int XP>   // 3: 3xTF32 split (near fp32), 2: activations split / weights rounded (p_hi.w + p_lo.w), 1: plain TF32
__global__ void __launch_bounds__(CW_THREADS) conv_temporal_fwd_mma_kernel(const float* __restrict__ x3,
                                                                           const float* __restrict__ wt,
                                                                           const float* __restrict__ bt,
                                                                           float* __restrict__ y1,
                                                                           double* __restrict__ sums) {
  __shared__ __align__(16) float xs_all[CW_WARPS][XS_LEN];
  __shared__ float ps_all[CW_WARPS][PS_LEN];
  __shared__ float red[2][N_FILT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x;
  float* xs = xs_all[warp];
  float* ps = ps_all[warp];
  for (int i = lane; i < XS_LEN; i += 32) xs[i] = 0.f;
  for (int i = lane; i < PS_LEN; i += 32) ps[i] = 0.f;
  if (threadIdx.x < 2 * N_FILT) (&red[0][0])[threadIdx.x] = 0.f;
  // B fragments: w[k][i]/51 split hi/lo; (k-dim = tap i, n-dim = filter k)
  uint32_t bh[4][5][2], bl[4][5][2];
#pragma unroll
  for (int kt = 0; kt < 4; ++kt)
#pragma unroll
    for (int nt = 0; nt < 5; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = kt * 8 + t + 4 * h, k = nt * 8 + g;
        const float w = i < K_TEMP ? wt[k * K_TEMP + i] * (1.f / K_POOL) : 0.f;
        const uint32_t hi = tf32_bits(w);
        bh[kt][nt][h] = hi;
        bl[kt][nt][h] = tf32_bits(w - __uint_as_float(hi));
      }
  float bias[5][2];
#pragma unroll
  for (int nt = 0; nt < 5; ++nt) { bias[nt][0] = bt[nt * 8 + 2 * t]; bias[nt][1] = bt[nt * 8 + 2 * t + 1]; }
  float s1[5][2], s2[5][2];
#pragma unroll
  for (int nt = 0; nt < 5; ++nt) s1[nt][0] = s1[nt][1] = s2[nt][0] = s2[nt][1] = 0.f;
  __syncthreads();

  for (int r = warp; r < N_CH; r += CW_WARPS) {
    warp_load_row_pool(x3 + ((size_t)b * N_TOK + r) * D_PAD, xs, ps, lane);
#pragma unroll 1
    for (int mt = 0; mt < 3; ++mt) {
      float c[5][4];
#pragma unroll
      for (int nt = 0; nt < 5; ++nt) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int j = mt * 16 + g + 8 * (h & 1), i = kt * 8 + t + 4 * (h >> 1);
          const float p = ps[5 * j + i];
          ah[h] = tf32_bits(p);
          if (XP >= 2) al[h] = tf32_bits(p - __uint_as_float(ah[h]));
        }
#pragma unroll
        for (int nt = 0; nt < 5; ++nt) {
          if (XP >= 2) mma_tf32(c[nt], al, bh[kt][nt][0], bh[kt][nt][1]);
          if (XP == 3) mma_tf32(c[nt], ah, bl[kt][nt][0], bl[kt][nt][1]);
          mma_tf32(c[nt], ah, bh[kt][nt][0], bh[kt][nt][1]);
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = mt * 16 + g + 8 * h;
        if (j < N_POOL) {
          float* dst = y1 + (((size_t)b * N_POOL + j) * N_CH + r) * N_FILT;
#pragma unroll
          for (int nt = 0; nt < 5; ++nt) {
            const float v0 = c[nt][2 * h] + bias[nt][0], v1 = c[nt][2 * h + 1] + bias[nt][1];
            *reinterpret_cast<float2*>(dst + nt * 8 + 2 * t) = make_float2(v0, v1);
            s1[nt][0] += v0; s1[nt][1] += v1;
            s2[nt][0] = fmaf(v0, v0, s2[nt][0]); s2[nt][1] = fmaf(v1, v1, s2[nt][1]);
          }
        }
      }
    }
    __syncwarp();
  }
  if (sums != nullptr) {
#pragma unroll
    for (int nt = 0; nt < 5; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float a = s1[nt][h], q = s2[nt][h];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        if (g == 0) { atomicAdd(&red[0][nt * 8 + 2 * t + h], a); atomicAdd(&red[1][nt * 8 + 2 * t + h], q); }
      }
    __syncthreads();
    if (threadIdx.x < N_FILT) {
      atomicAdd(&sums[threadIdx.x], (double)red[0][threadIdx.x]);
      atomicAdd(&sums[N_FILT + threadIdx.x], (double)red[1][threadIdx.x]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward: BN1 backward apply fused in (dz1, y1 -> dy on the fly); outputs dX3, dWt, dbt (+ BN1 dgamma/dbeta)
// ------------------------------------------------------------------------------------------------
static constexpr int DY_LD = 44;
static constexpr int DY_ROWS = 52;           // 4 zero rows in front (shifted reads m-a), 36 data rows, 12 zero rows
static constexpr int DPZ = 320;              // 50 leading zeros + dp[0..199] + zero tail
static constexpr int RAW = N_POOL * N_FILT;  // 1440 floats of dz1 / y1 per (sample, row), staged with cp.async

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// stage dz1 / y1 of (b, r): 36 segments of 40 floats each, laid out [j][k] in smem
__device__ __forceinline__ void prefetch_row(const float* __restrict__ dz1, const float* __restrict__ y1, int b, int r,
                                             float* rawdz, float* rawy, int lane) {
  for (int f = lane; f < N_POOL * 10; f += 32) {
    const int j = f / 10, k4 = (f % 10) * 4;
    const size_t idx = (((size_t)b * N_POOL + j) * N_CH + r) * N_FILT + k4;
    cp_async16(rawdz + f * 4, dz1 + idx);
    cp_async16(rawy + f * 4, y1 + idx);
  }
  cp_async_commit();
}

__global__ void __launch_bounds__(CW_THREADS) conv_temporal_bwd_mma_kernel(
    const float* __restrict__ dz1, const float* __restrict__ y1, const float* __restrict__ x3,
    const float* __restrict__ wt, const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
    const double* __restrict__ bwd_sums, long long count, float* __restrict__ dx3, float* __restrict__ dwt,
    float* __restrict__ dbt, float* __restrict__ dgamma, float* __restrict__ dbeta, float gscale) {
  extern __shared__ __align__(16) float smem[];
  // CTA-shared: BN constants [5][40], dW reduction [40][26]
  float* c_mu = smem;
  float* c_rs = c_mu + N_FILT;
  float* c_gr = c_rs + N_FILT;
  float* c_m1 = c_gr + N_FILT;
  float* c_m2 = c_m1 + N_FILT;
  float* wred = c_m2 + N_FILT;                       // [40][26]
  float* per_warp = wred + N_FILT * 26 + 8;          // keep 16-byte alignment: 200 + 1040 + 8 = 1248 floats
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x;
  constexpr int PW = XS_LEN + PS_LEN + DY_ROWS * DY_LD + 16 + DPZ + 2 * RAW;
  float* xs = per_warp + (size_t)warp * PW;
  float* ps = xs + XS_LEN;
  float* dys = ps + PS_LEN;
  float* dpz = dys + DY_ROWS * DY_LD + 16;
  float* rawdz = dpz + DPZ;
  float* rawy = rawdz + RAW;
  for (int i = lane; i < PW - 2 * RAW; i += 32) xs[i] = 0.f;
  if (warp < N_CH) prefetch_row(dz1, y1, b, warp, rawdz, rawy, lane);
  for (int i = threadIdx.x; i < N_FILT * 26; i += CW_THREADS) wred[i] = 0.f;
  if (threadIdx.x < N_FILT) {
    const int k = threadIdx.x;
    c_mu[k] = mean_rstd[k];
    c_rs[k] = mean_rstd[N_FILT + k];
    c_gr[k] = gamma[k] * mean_rstd[N_FILT + k];
    c_m1[k] = (float)(bwd_sums[k] / (double)count);
    c_m2[k] = (float)(bwd_sums[N_FILT + k] / (double)count);
    if (b == 0) {
      dgamma[k] += gscale * (float)bwd_sums[N_FILT + k];
      dbeta[k] += gscale * (float)bwd_sums[k];
    }
  }
  // B fragments of the dp GEMM: W'[(a,kk*8+k)][rho] = w[k][rho+5a]/51  (n = rho = g, valid for g < 5)
  uint32_t bw[25][2];
#pragma unroll
  for (int a = 0; a < 5; ++a)
#pragma unroll
    for (int kk = 0; kk < 5; ++kk)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = kk * 8 + t + 4 * h, i = g + 5 * a;
        bw[a * 5 + kk][h] = g < 5 ? tf32_bits(wt[k * K_TEMP + i] * (1.f / K_POOL)) : 0u;
      }
  float accw[3][4][4];
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) accw[mt][nt][0] = accw[mt][nt][1] = accw[mt][nt][2] = accw[mt][nt][3] = 0.f;
  __syncthreads();

  for (int r = warp; r < N_CH; r += CW_WARPS) {
    warp_load_row_pool(x3 + ((size_t)b * N_TOK + r) * D_PAD, xs, ps, lane);
    // ---- dy[j][k] for this (b, r) -> smem (TF32-rounded), rows offset by 4 ----
    cp_async_wait_all();
    __syncwarp();
    for (int f = lane; f < N_POOL * 10; f += 32) {
      const int j = f / 10, k4 = (f % 10) * 4;
      const float4 dz = *reinterpret_cast<const float4*>(rawdz + f * 4);
      const float4 yv = *reinterpret_cast<const float4*>(rawy + f * 4);
      const float dzv[4] = {dz.x, dz.y, dz.z, dz.w}, yy[4] = {yv.x, yv.y, yv.z, yv.w};
      float o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k4 + q;
        const float yh = (yy[q] - c_mu[k]) * c_rs[k];
        o[q] = tf32_rn(c_gr[k] * (dzv[q] - c_m1[k] - yh * c_m2[k]));
      }
      *reinterpret_cast<float4*>(dys + (j + 4) * DY_LD + k4) = make_float4(o[0], o[1], o[2], o[3]);
    }
    __syncwarp();
    if (r + CW_WARPS < N_CH) prefetch_row(dz1, y1, b, r + CW_WARPS, rawdz, rawy, lane);   // overlaps the MMAs below
    // ---- dW[k][i] += sum_j dy[j][k] * p[5j+i]   (column i == 25 carries a ones-vector: the bias gradient) ----
#pragma unroll
    for (int kt = 0; kt < 5; ++kt) {
      uint32_t bp[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = kt * 8 + t + 4 * h, i = nt * 8 + g;
          float v = 0.f;
          if (i < K_TEMP) v = ps[5 * j + i];
          else if (i == K_TEMP) v = j < N_POOL ? 1.f : 0.f;
          bp[nt][h] = tf32_bits(v);
        }
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {
        uint32_t a[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int k = mt * 16 + g + 8 * (h & 1), j = kt * 8 + t + 4 * (h >> 1);
          a[h] = __float_as_uint(dys[(j + 4) * DY_LD + k]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32(accw[mt][nt], a, bp[nt][0], bp[nt][1]);
      }
    }
    // ---- dp[5m+rho] = sum_{a,k} dy[m-a][k] * w[k][rho+5a]/51 ----
    // 6 independent accumulator chains (3 m-tiles x even/odd k-tile): a single chain of 25 dependent MMAs per m-tile
    // left the tensor pipe idle most of the time ("wait" was the top stall reason in ncu)
    {
      float c[3][2][4];
#pragma unroll
      for (int mt = 0; mt < 3; ++mt)
#pragma unroll
        for (int e = 0; e < 2; ++e) c[mt][e][0] = c[mt][e][1] = c[mt][e][2] = c[mt][e][3] = 0.f;
#pragma unroll
      for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int kk = 0; kk < 5; ++kk) {
#pragma unroll
          for (int mt = 0; mt < 3; ++mt) {
            uint32_t af[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int m = mt * 16 + g + 8 * (h & 1), k = kk * 8 + t + 4 * (h >> 1);
              af[h] = __float_as_uint(dys[(m - a + 4) * DY_LD + k]);
            }
            mma_tf32(c[mt][(a * 5 + kk) & 1], af, bw[a * 5 + kk][0], bw[a * 5 + kk][1]);
          }
        }
#pragma unroll
      for (int mt = 0; mt < 3; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = mt * 16 + g + 8 * h;
          if (m < 40) {
            if (2 * t < 5) dpz[50 + 5 * m + 2 * t] = c[mt][0][2 * h] + c[mt][1][2 * h];
            if (2 * t + 1 < 5) dpz[50 + 5 * m + 2 * t + 1] = c[mt][0][2 * h + 1] + c[mt][1][2 * h + 1];
          }
        }
    }
    __syncwarp();
    // ---- dx[t] = sum_{s=t-50}^{t} dp[s]  (sliding window, 8 outputs per lane) ----
    {
      const int t0 = lane * 8;
      float a = 0.f;
#pragma unroll
      for (int v = 0; v < K_POOL; ++v) a += dpz[t0 + v];          // dpz index = 50 + s, s = t0-50 .. t0
      float o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        o[q] = (t0 + q) < N_T ? a : 0.f;
        a += dpz[t0 + q + K_POOL] - dpz[t0 + q];
      }
      float4* dst = reinterpret_cast<float4*>(dx3 + ((size_t)b * N_TOK + r) * D_PAD + t0);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
    __syncwarp();
  }
  // token 63 (channel 62) never reaches the conv stack (enc_out[:, :63], ATMS_retrieval.py:91)
  for (int i = threadIdx.x; i < D_PAD; i += CW_THREADS) dx3[((size_t)b * N_TOK + N_CH) * D_PAD + i] = 0.f;
  // reduce the dW accumulators of the 4 warps, then one atomic per entry per CTA
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int k = mt * 16 + g + 8 * (h >> 1), i = nt * 8 + 2 * t + (h & 1);
        if (k < N_FILT && i <= K_TEMP) atomicAdd(&wred[k * 26 + i], accw[mt][nt][h]);
      }
  __syncthreads();
  for (int idx = threadIdx.x; idx < N_FILT * 26; idx += CW_THREADS) {
    const int k = idx / 26, i = idx % 26;
    if (i < K_TEMP) atomicAdd(&dwt[k * K_TEMP + i], wred[idx] * (1.f / K_POOL));
    else atomicAdd(&dbt[k], wred[idx]);
  }
}

// ------------------------------------------------------------------------------------------------
// backward, second version.  Same math and the same two MMA phases as conv_temporal_bwd_mma_kernel; what changed is
// everything around them, following the ncu source view of the first version (profiles/r01_conv_temporal_bwd.ncu-rep:
// 24 % of the stall samples in the dz/y -> dy transform, 20 % in the two serial 51-term sliding sums, 7.5 % on the
// un-prefetched token-row load, ~5 % in per-use cvt.rna of the pooled sums):
//   * the token row is prefetched with cp.async together with dz / y of the next row;
//   * box-51 sums and the transposed pooling are differences of a warp-scanned prefix sum (8 values per lane, 5 shuffle
//     steps) instead of 51 dependent adds per lane;
//   * the BatchNorm backward is folded into three per-channel constants dy = A*dz + B*y + C kept in registers: 30 lanes
//     cover exactly 3 rows x 10 float4, so a lane's channel quad never changes (was 20 shared loads per float4);
//   * pooled sums are stored TF32-rounded once instead of converted at every fragment gather.
// ------------------------------------------------------------------------------------------------
// A fragment (16 rows x 8 tf32) of a row-major smem tile with one ldmatrix.x4: viewed as b16, an 8x8 matrix is 8 rows
// of 4 floats; lane L supplies the address of row (L%8) + 8*((L/8)&1), column 4*(L/16), and receives element
// (row L/4, col L%4) of each of the four 8x4 float blocks = a0..a3 of mma.m16n8k8.tf32.  Rows must be 16-byte aligned.
__device__ __forceinline__ void ldsm_a_tf32(uint32_t a[4], const float* lane_row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
               : "r"(smem_u32(lane_row_ptr)));
}

static constexpr int XRAW = 256;     // token row staging (cp.async)
static constexpr int CS_LEN = 272;   // prefix sums C[0..256]
static constexpr int DP_LEN = 256;   // dp[0..199] + zero tail

// v[0..7]: 8 consecutive values of lane `lane`.  Returns the exclusive-prefix form E[i] = sum of all elements before
// element 8*lane+i, plus the inclusive total of the lane in *last.
__device__ __forceinline__ void warp_excl_scan8(float v[8], int lane, float* last) {
#pragma unroll
  for (int i = 1; i < 8; ++i) v[i] += v[i - 1];
  const float tot = v[7];
  float inc = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  const float excl = inc - tot;
  *last = inc;
#pragma unroll
  for (int i = 7; i >= 1; --i) v[i] = v[i - 1] + excl;
  v[0] = excl;
}

__device__ __forceinline__ void prefetch_row2(const float* __restrict__ dz1, const float* __restrict__ y1,
                                              const float* __restrict__ x3, int b, int r, float* rawdz, float* rawy,
                                              float* xraw, int lane) {
  const float* xrow = x3 + ((size_t)b * N_TOK + r) * D_PAD;
  cp_async16(xraw + lane * 4, xrow + lane * 4);
  cp_async16(xraw + 128 + lane * 4, xrow + 128 + lane * 4);
  for (int f = lane; f < N_POOL * 10; f += 32) {
    const int j = f / 10, k4 = (f % 10) * 4;
    const size_t idx = (((size_t)b * N_POOL + j) * N_CH + r) * N_FILT + k4;
    cp_async16(rawdz + f * 4, dz1 + idx);
    cp_async16(rawy + f * 4, y1 + idx);
  }
  cp_async_commit();
}

__global__ void __launch_bounds__(CW_THREADS) conv_temporal_bwd_mma2_kernel(
    const float* __restrict__ dz1, const float* __restrict__ y1, const float* __restrict__ x3,
    const float* __restrict__ wt, const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
    const double* __restrict__ bwd_sums, long long count, float* __restrict__ dx3, float* __restrict__ dwt,
    float* __restrict__ dbt, float* __restrict__ dgamma, float* __restrict__ dbeta, float gscale) {
  extern __shared__ __align__(16) float smem[];
  // CTA-shared: folded BatchNorm-backward constants [3][40], dW reduction [40][26]
  float* c_A = smem;
  float* c_B = c_A + N_FILT;
  float* c_C = c_B + N_FILT;
  float* wred = c_C + N_FILT;                        // [40][26]
  float* per_warp = wred + N_FILT * 26;              // 120 + 1040 = 1160 floats: 16-byte aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x;
  constexpr int PW = XRAW + CS_LEN + PS_LEN + DY_ROWS * DY_LD + 16 + DP_LEN + 2 * RAW;
  float* xraw = per_warp + (size_t)warp * PW;
  float* cs = xraw + XRAW;
  float* ps = cs + CS_LEN;
  float* dys = ps + PS_LEN;
  float* dps = dys + DY_ROWS * DY_LD + 16;
  float* rawdz = dps + DP_LEN;
  float* rawy = rawdz + RAW;
  for (int i = lane; i < PW - 2 * RAW - XRAW; i += 32) cs[i] = 0.f;          // cs, ps (zero tail), dys (zero rows), dps
  if (warp < N_CH) prefetch_row2(dz1, y1, x3, b, warp, rawdz, rawy, xraw, lane);
  for (int i = threadIdx.x; i < N_FILT * 26; i += CW_THREADS) wred[i] = 0.f;
  if (threadIdx.x < N_FILT) {
    const int k = threadIdx.x;
    const float mu = mean_rstd[k], rs = mean_rstd[N_FILT + k];
    const float gr = gamma[k] * rs;
    const float m1 = (float)(bwd_sums[k] / (double)count), m2 = (float)(bwd_sums[N_FILT + k] / (double)count);
    // dy = gr*(dz - m1 - (y-mu)*rs*m2) = A*dz + B*y + C
    c_A[k] = gr;
    c_B[k] = -gr * m2 * rs;
    c_C[k] = gr * (m2 * rs * mu - m1);
    if (b == 0) {
      dgamma[k] += gscale * (float)bwd_sums[N_FILT + k];
      dbeta[k] += gscale * (float)bwd_sums[k];
    }
  }
  // B fragments of the dp GEMM: W'[(a,kk*8+k)][rho] = w[k][rho+5a]/51  (n = rho = g, valid for g < 5)
  uint32_t bw[25][2];
#pragma unroll
  for (int a = 0; a < 5; ++a)
#pragma unroll
    for (int kk = 0; kk < 5; ++kk)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = kk * 8 + t + 4 * h, i = g + 5 * a;
        bw[a * 5 + kk][h] = g < 5 ? tf32_bits(wt[k * K_TEMP + i] * (1.f / K_POOL)) : 0u;
      }
  float accw[3][4][4];
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) accw[mt][nt][0] = accw[mt][nt][1] = accw[mt][nt][2] = accw[mt][nt][3] = 0.f;
  __syncthreads();
  // this lane's channel quad in the transform (lanes 0..29: 3 rows x 10 float4 per pass)
  const int tr_j = lane / 10, tr_k4 = (lane % 10) * 4;
  float kA[4], kB[4], kC[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { kA[q] = c_A[tr_k4 + q]; kB[q] = c_B[tr_k4 + q]; kC[q] = c_C[tr_k4 + q]; }

  for (int r = warp; r < N_CH; r += CW_WARPS) {
    cp_async_wait_all();
    __syncwarp();
    // ---- box-51 sums of the token row: ps[s] = C[s+51] - C[s], C = prefix sums (TF32-rounded: dW operand only) ----
    {
      const float4 x0 = *reinterpret_cast<const float4*>(xraw + 8 * lane), x1 = *reinterpret_cast<const float4*>(xraw + 8 * lane + 4);
      float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      float last;
      warp_excl_scan8(v, lane, &last);
      *reinterpret_cast<float4*>(cs + 8 * lane) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(cs + 8 * lane + 4) = make_float4(v[4], v[5], v[6], v[7]);
      if (lane == 31) cs[256] = last;
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const int sidx = lane + 32 * q;
        if (sidx < N_PSUM) ps[sidx] = tf32_rn(cs[sidx + K_POOL] - cs[sidx]);
      }
    }
    // ---- dy[j][k] = A*dz + B*y + C for this (b, r) -> smem (TF32-rounded), rows offset by 4 ----
    if (lane < 30) {
#pragma unroll 4
      for (int it = 0; it < 12; ++it) {
        const int j = it * 3 + tr_j;
        const float4 dz = *reinterpret_cast<const float4*>(rawdz + j * N_FILT + tr_k4);
        const float4 yv = *reinterpret_cast<const float4*>(rawy + j * N_FILT + tr_k4);
        float4 o;
        o.x = tf32_rn(fmaf(kA[0], dz.x, fmaf(kB[0], yv.x, kC[0])));
        o.y = tf32_rn(fmaf(kA[1], dz.y, fmaf(kB[1], yv.y, kC[1])));
        o.z = tf32_rn(fmaf(kA[2], dz.z, fmaf(kB[2], yv.z, kC[2])));
        o.w = tf32_rn(fmaf(kA[3], dz.w, fmaf(kB[3], yv.w, kC[3])));
        *reinterpret_cast<float4*>(dys + (j + 4) * DY_LD + tr_k4) = o;
      }
    }
    __syncwarp();
    if (r + CW_WARPS < N_CH) prefetch_row2(dz1, y1, x3, b, r + CW_WARPS, rawdz, rawy, xraw, lane);   // overlaps the MMAs below
    // ---- dW[k][i] += sum_j dy[j][k] * p[5j+i]   (column i == 25 carries a ones-vector: the bias gradient) ----
#pragma unroll
    for (int kt = 0; kt < 5; ++kt) {
      uint32_t bp[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = kt * 8 + t + 4 * h, i = nt * 8 + g;
          float v = 0.f;
          if (i < K_TEMP) v = ps[5 * j + i];                      // already TF32
          else if (i == K_TEMP) v = j < N_POOL ? 1.f : 0.f;
          bp[nt][h] = __float_as_uint(v);
        }
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {
        uint32_t a[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int k = mt * 16 + g + 8 * (h & 1), j = kt * 8 + t + 4 * (h >> 1);
          a[h] = __float_as_uint(dys[(j + 4) * DY_LD + k]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32(accw[mt][nt], a, bp[nt][0], bp[nt][1]);
      }
    }
    // ---- dp[5m+rho] = sum_{a,k} dy[m-a][k] * w[k][rho+5a]/51 : 6 independent accumulator chains ----
    {
      // ldmatrix row address of this lane: tile row (lane%8) + 8*((lane/8)&1) (+4 zero rows in front), column 4*(lane/16)
      const float* dys_lane = dys + ((lane & 7) + 8 * ((lane >> 3) & 1) + 4) * DY_LD + 4 * (lane >> 4);
      float c[3][2][4];
#pragma unroll
      for (int mt = 0; mt < 3; ++mt)
#pragma unroll
        for (int e = 0; e < 2; ++e) c[mt][e][0] = c[mt][e][1] = c[mt][e][2] = c[mt][e][3] = 0.f;
#pragma unroll
      for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int kk = 0; kk < 5; ++kk) {
#pragma unroll
          for (int mt = 0; mt < 3; ++mt) {
            uint32_t af[4];      // rows m - a (m = mt*16 .. +15), columns kk*8 .. +7 of dy: one ldmatrix.x4
            ldsm_a_tf32(af, dys_lane + ((mt * 16 - a) * DY_LD + kk * 8));
            mma_tf32(c[mt][(a * 5 + kk) & 1], af, bw[a * 5 + kk][0], bw[a * 5 + kk][1]);
          }
        }
#pragma unroll
      for (int mt = 0; mt < 3; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = mt * 16 + g + 8 * h;
          if (m < 40) {
            if (2 * t < 5) dps[5 * m + 2 * t] = c[mt][0][2 * h] + c[mt][1][2 * h];
            if (2 * t + 1 < 5) dps[5 * m + 2 * t + 1] = c[mt][0][2 * h + 1] + c[mt][1][2 * h + 1];
          }
        }
    }
    __syncwarp();
    // ---- dx[t] = sum_{s=max(t-50,0)}^{min(t,199)} dp[s] = D[min(t,199)+1] - D[max(t-50,0)], D = prefix sums of dp ----
    {
      const float4 d0 = *reinterpret_cast<const float4*>(dps + 8 * lane), d1 = *reinterpret_cast<const float4*>(dps + 8 * lane + 4);
      float v[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      float last;
      warp_excl_scan8(v, lane, &last);
      *reinterpret_cast<float4*>(cs + 8 * lane) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(cs + 8 * lane + 4) = make_float4(v[4], v[5], v[6], v[7]);
      if (lane == 31) cs[256] = last;
      __syncwarp();
      const int t0 = lane * 8;
      float o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int tt = t0 + q;
        const int hi = (tt < N_PSUM ? tt : N_PSUM - 1) + 1, lo = tt > K_POOL - 1 ? tt - (K_POOL - 1) : 0;
        o[q] = tt < N_T ? cs[hi] - cs[lo] : 0.f;
      }
      float4* dst = reinterpret_cast<float4*>(dx3 + ((size_t)b * N_TOK + r) * D_PAD + t0);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
    __syncwarp();
  }
  // token 63 (channel 62) never reaches the conv stack (enc_out[:, :63], ATMS_retrieval.py:91)
  for (int i = threadIdx.x; i < D_PAD; i += CW_THREADS) dx3[((size_t)b * N_TOK + N_CH) * D_PAD + i] = 0.f;
  // reduce the dW accumulators of the 4 warps, then one atomic per entry per CTA
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int k = mt * 16 + g + 8 * (h >> 1), i = nt * 8 + 2 * t + (h & 1);
        if (k < N_FILT && i <= K_TEMP) atomicAdd(&wred[k * 26 + i], accw[mt][nt][h]);
      }
  __syncthreads();
  for (int idx = threadIdx.x; idx < N_FILT * 26; idx += CW_THREADS) {
    const int k = idx / 26, i = idx % 26;
    if (i < K_TEMP) atomicAdd(&dwt[k * K_TEMP + i], wred[idx] * (1.f / K_POOL));
    else atomicAdd(&dbt[k], wred[idx]);
  }
}

// ------------------------------------------------------------------------------------------------
int conv_temporal_fwd_simt(const float* x3, const float* wt, const float* bt, float* y1, double* sums, int B, cudaStream_t s);
int conv_temporal_bwd_simt(const float* dz1, const float* y1, const float* x3, const float* wt, const float* mean_rstd,
                           const float* gamma, const double* bwd_sums, long long count, float* dx3, float* dwt, float* dbt,
                           float* dgamma, float* dbeta, int B, float gscale, cudaStream_t s);

int conv_temporal_fwd(const float* x3, const float* wt, const float* bt, float* y1, double* sums, int B, cudaStream_t s) {
  if (!tf32_rounding()) return conv_temporal_fwd_simt(x3, wt, bt, y1, sums, B, s);   // exact-fp32 verification path
  ProfScope _ps("conv_temporal_fwd", s, (double)B * 63 * 36 * 40 * 50.0, (double)B * (63 * 1000.0 + 36 * 2520 * 4.0));
  static int xp = -1;
  if (xp < 0) { const char* e = getenv("EEGB200_CONV_XP"); xp = (e && e[0] == '1') ? 1 : ((e && e[0] == '2') ? 2 : 3); }
  if (xp == 3) conv_temporal_fwd_mma_kernel<3><<<B, CW_THREADS, 0, s>>>(x3, wt, bt, y1, sums);
  else if (xp == 2) conv_temporal_fwd_mma_kernel<2><<<B, CW_THREADS, 0, s>>>(x3, wt, bt, y1, sums);
  else conv_temporal_fwd_mma_kernel<1><<<B, CW_THREADS, 0, s>>>(x3, wt, bt, y1, sums);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int conv_temporal_bwd(const float* dz1, const float* y1, const float* x3, const float* wt, const float* mean_rstd,
                      const float* gamma, const double* bwd_sums, long long count, float* dx3, float* dwt, float* dbt,
                      float* dgamma, float* dbeta, int B, float gscale, cudaStream_t s) {
  if (!tf32_rounding())
    return conv_temporal_bwd_simt(dz1, y1, x3, wt, mean_rstd, gamma, bwd_sums, count, dx3, dwt, dbt, dgamma, dbeta, B, gscale, s);
  ProfScope _ps("conv_temporal_bwd", s, (double)B * 63 * 36 * 40 * 100.0, (double)B * (36 * 2520 * 8.0 + 63 * 2000.0));
  static int version = -1;
  if (version < 0) { const char* e = getenv("EEGB200_CONV_BWD"); version = (e && e[0] == '1') ? 1 : 2; }
  if (version == 2) {
    constexpr int PW2 = XRAW + CS_LEN + PS_LEN + DY_ROWS * DY_LD + 16 + DP_LEN + 2 * RAW;
    const size_t smem2 = (size_t)(3 * N_FILT + N_FILT * 26 + CW_WARPS * PW2) * sizeof(float);
    static PerDeviceOnce once2;
    if (once2.first())
      EEG_CUDA_OK(cudaFuncSetAttribute(conv_temporal_bwd_mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    conv_temporal_bwd_mma2_kernel<<<B, CW_THREADS, smem2, s>>>(dz1, y1, x3, wt, mean_rstd, gamma, bwd_sums, count, dx3, dwt,
                                                              dbt, dgamma, dbeta, gscale);
    EEG_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
  }
  constexpr int PW = XS_LEN + PS_LEN + DY_ROWS * DY_LD + 16 + DPZ + 2 * RAW;
  const size_t smem = (size_t)(5 * N_FILT + N_FILT * 26 + 8 + CW_WARPS * PW) * sizeof(float);
  static PerDeviceOnce once;
  if (once.first()) {
    EEG_CUDA_OK(cudaFuncSetAttribute(conv_temporal_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  conv_temporal_bwd_mma_kernel<<<B, CW_THREADS, smem, s>>>(dz1, y1, x3, wt, mean_rstd, gamma, bwd_sums, count, dx3, dwt, dbt,
                                                           dgamma, dbeta, gscale);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace eegb200

// The skinny temporal + spatial conv stack of PatchEmbedding (Retrieval/ATMS_retrieval.py:101-115):
//   Conv2d(1->40,(1,25)) -> AvgPool((1,51),/5) -> BN -> ELU -> Conv2d(40->40,(63,1)) -> BN -> ELU -> Dropout
//   -> Conv2d(40->40,1x1) -> 'b e h w -> b (h w) e'
// conv(25 taps) o avgpool(51,/5) is computed exactly as a 51-wide box prefilter followed by a 25-tap stride-5
// conv (both linear, they commute): 4.5 MFLOP/sample instead of 33.
// Layouts: y1 / a1 : [B*36 (b,j), 63 (r), 40 (k)]  == row-major [B*36, 2520], the A operand of the spatial-conv GEMM
//          y2      : [B*36, 40];   feat : [B, 36*40] (index j*40 + e, the reference's flatten order)
#include "kernels.h"

namespace eegb200 {

static constexpr int CT_THREADS = 160;   // 4 pooled positions x 40 filters per pass

// shared by fwd/bwd: load token row (250 valid of 256) and build the 200 box-51 sums
__device__ __forceinline__ void load_row_and_pool(const float* __restrict__ xrow, float* xs, float* ps) {
  for (int t = threadIdx.x; t < 256; t += CT_THREADS) xs[t] = t < N_T ? xrow[t] : 0.f;
  __syncthreads();
  for (int s = threadIdx.x; s < N_PSUM; s += CT_THREADS) {
    float a = 0.f;
#pragma unroll
    for (int v = 0; v < K_POOL; ++v) a += xs[s + v];
    ps[s] = a;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(CT_THREADS) conv_temporal_fwd_kernel(const float* __restrict__ x3,
                                                                       const float* __restrict__ wt,
                                                                       const float* __restrict__ bt,
                                                                       float* __restrict__ y1,
                                                                       double* __restrict__ sums) {
  __shared__ float xs[256 + 64];
  __shared__ float ps[N_PSUM];
  __shared__ float red[2][4][N_FILT];
  const int b = blockIdx.x;
  const int k = threadIdx.x % N_FILT, jq = threadIdx.x / N_FILT;
  float w[K_TEMP];
#pragma unroll
  for (int i = 0; i < K_TEMP; ++i) w[i] = wt[k * K_TEMP + i] * (1.f / K_POOL);
  const float bias = bt[k];
  float s1 = 0.f, s2 = 0.f;
  for (int r = 0; r < N_CH; ++r) {
    load_row_and_pool(x3 + ((size_t)b * N_TOK + r) * D_PAD, xs, ps);
#pragma unroll 3
    for (int jj = 0; jj < N_POOL / 4; ++jj) {
      const int j = jq + 4 * jj;
      float a = bias;
#pragma unroll
      for (int i = 0; i < K_TEMP; ++i) a = fmaf(w[i], ps[5 * j + i], a);
      y1[(((size_t)b * N_POOL + j) * N_CH + r) * N_FILT + k] = a;
      s1 += a;
      s2 = fmaf(a, a, s2);
    }
    __syncthreads();
  }
  if (sums != nullptr) {
    red[0][jq][k] = s1;
    red[1][jq][k] = s2;
    __syncthreads();
    if (threadIdx.x < N_FILT) {
      double t1 = 0.0, t2 = 0.0;
      for (int q = 0; q < 4; ++q) { t1 += red[0][q][k]; t2 += red[1][q][k]; }
      atomicAdd(&sums[k], t1);
      atomicAdd(&sums[N_FILT + k], t2);
    }
  }
}
int conv_temporal_fwd_simt(const float* x3, const float* wt, const float* bt, float* y1, double* sums, int B, cudaStream_t s) {
  ProfScope _ps("conv_temporal_fwd_simt", s, (double)B * 63 * 36 * 40 * 50.0, (double)B * (63 * 1000.0 + 36 * 2520 * 4.0));
  conv_temporal_fwd_kernel<<<B, CT_THREADS, 0, s>>>(x3, wt, bt, y1, sums);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// statistics -> (mean, rstd); running-stat update like nn.BatchNorm2d (momentum 0.1, unbiased variance)
__global__ void bn_finalize_kernel(BnState bn, long long count, int train, int update_running) {
  const int k = threadIdx.x;
  if (k >= N_FILT) return;
  if (train) {
    const double mean = bn.sums[k] / (double)count;
    double var = bn.sums[N_FILT + k] / (double)count - mean * mean;
    if (var < 0.0) var = 0.0;
    bn.mean_rstd[k] = (float)mean;
    bn.mean_rstd[N_FILT + k] = (float)(1.0 / sqrt(var + (double)NORM_EPS));
    if (update_running) {
      const double unb = count > 1 ? var * (double)count / (double)(count - 1) : var;
      bn.running_mean[k] = (1.f - BN_MOMENTUM) * bn.running_mean[k] + BN_MOMENTUM * (float)mean;
      bn.running_var[k] = (1.f - BN_MOMENTUM) * bn.running_var[k] + BN_MOMENTUM * (float)unb;
    }
  } else {
    bn.mean_rstd[k] = bn.running_mean[k];
    bn.mean_rstd[N_FILT + k] = rsqrtf(bn.running_var[k] + NORM_EPS);
  }
}
int bn_finalize(BnState bn, long long count, int train, int update_running, cudaStream_t s) {
  ProfScope _ps("bn_finalize", s);
  bn_finalize_kernel<<<1, 64, 0, s>>>(bn, count, train, update_running);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// a = ELU(gamma*(y-mean)*rstd + beta), channel = index % 40.
// 320 threads (a multiple of the 10 float4 per 40-channel period) and a grid stride that is a multiple of 10 keep every
// thread on ONE column quad, so its four (scale, shift) pairs live in registers; expm1 is exp2-based with a 4-term
// series near zero.  The first version (64-bit modulo, 8 shared loads and 4 libm expm1f per float4) executed 190
// instructions per warp iteration and was issue-bound (ncu: 80 % issue-active, 54 % of the DRAM peak).
__device__ __forceinline__ float elu1_fast(float z) {
  const float e = __expf(z) - 1.f;                                   // abs error ~1e-7: fine away from 0
  const float p = z * fmaf(z, fmaf(z, fmaf(z, 1.f / 24.f, 1.f / 6.f), 0.5f), 1.f);   // |z| < 1/16: error < 1e-8
  const float neg = z > -0.0625f ? p : e;
  return z > 0.f ? z : neg;
}
__global__ void __launch_bounds__(320) bn_elu_apply_kernel(const float4* __restrict__ y, const float* __restrict__ mean_rstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float4* __restrict__ a, long long n4, int round_tf) {
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int c = (int)(i0 % 10) * 4;                                  // invariant: the stride below is a multiple of 10
  float sc[4], sh[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float r = mean_rstd[N_FILT + c + q] * gamma[c + q];
    sc[q] = r;
    sh[q] = beta[c + q] - mean_rstd[c + q] * r;
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = i0; i < n4; i += 2 * stride) {
    const long long j = i + stride;
    const float4 v0 = y[i];
    float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < n4) v1 = y[j];
    float o[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      o[q] = elu1_fast(fmaf(o[q], sc[q & 3], sh[q & 3]));
      if (round_tf) o[q] = tf32_rn(o[q]);
    }
    a[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (j < n4) a[j] = make_float4(o[4], o[5], o[6], o[7]);
  }
}
int bn_elu_apply(const float* y, const float* mean_rstd, const float* gamma, const float* beta, float* a, long long n,
                 int round_tf, cudaStream_t s) {
  ProfScope _ps("bn_elu_apply", s, 0.0, (double)n * 8.0);
  EEG_REQUIRE(n % N_FILT == 0, "bn_elu_apply: element count must be a multiple of 40 channels");
  const long long n4 = n / 4;
  long long blocks = (n4 + 2 * 320 - 1) / (2 * 320);
  if (blocks > 148 * 12) blocks = 148 * 12;
  if (blocks < 1) blocks = 1;
  // blocks * 320 is a multiple of 10 float4 = one 40-channel period for any block count
  bn_elu_apply_kernel<<<(int)blocks, 320, 0, s>>>(reinterpret_cast<const float4*>(y), mean_rstd, gamma, beta,
                                                  reinterpret_cast<float4*>(a), n4, round_tf);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// per-column sum / sum of squares of y[rows, cols<=64] (double atomics)
__global__ void colstats_kernel(const float* __restrict__ y, int ld, int rows, int cols, double* __restrict__ sums) {
  __shared__ float p1[4][64], p2[4][64];
  const int c = threadIdx.x & 63, q = threadIdx.x >> 6;
  float s1 = 0.f, s2 = 0.f;
  if (c < cols)
    for (int r = blockIdx.x * 4 + q; r < rows; r += gridDim.x * 4) {
      const float v = y[(size_t)r * ld + c];
      s1 += v;
      s2 = fmaf(v, v, s2);
    }
  p1[q][c] = s1;
  p2[q][c] = s2;
  __syncthreads();
  if (q == 0 && c < cols) {
    atomicAdd(&sums[c], (double)p1[0][c] + p1[1][c] + p1[2][c] + p1[3][c]);
    atomicAdd(&sums[cols + c], (double)p2[0][c] + p2[1][c] + p2[2][c] + p2[3][c]);
  }
}
int colstats(const float* y, int ld, int rows, int cols, double* sums, cudaStream_t s) {
  ProfScope _ps("colstats", s, 0.0, (double)rows * cols * 4.0);
  EEG_REQUIRE(cols <= 64, "colstats: cols %d > 64", cols);
  int blocks = cdiv(rows, 4 * 16);
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  colstats_kernel<<<blocks, 256, 0, s>>>(y, ld, rows, cols, sums);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// BN2 -> ELU -> Dropout(0.5) -> 1x1 conv; 16 (b,j) rows per 256-thread block
static constexpr int HB = 16;
__global__ void __launch_bounds__(256) conv_head_fwd_kernel(const float* __restrict__ y2, const float* __restrict__ mean_rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ wc, const float* __restrict__ bc,
                                                            float* __restrict__ feat, int rows, DropoutCfg drop, int rt) {
  __shared__ float a2[HB][N_FILT + 1], swc[N_FILT * N_FILT + N_FILT], sc[N_FILT], sh[N_FILT];
  for (int i = threadIdx.x; i < N_FILT * N_FILT; i += 256) swc[i + i / N_FILT] = wc[i];   // row stride 41: conflict-free
  if (threadIdx.x < N_FILT) {
    const float r = mean_rstd[N_FILT + threadIdx.x] * gamma[threadIdx.x];
    sc[threadIdx.x] = r;
    sh[threadIdx.x] = beta[threadIdx.x] - mean_rstd[threadIdx.x] * r;
  }
  __syncthreads();
  const int row0 = blockIdx.x * HB;
  for (int idx = threadIdx.x; idx < HB * N_FILT; idx += 256) {
    const int r = idx / N_FILT, k = idx % N_FILT, row = row0 + r;
    float a = 0.f;
    if (row < rows) {
      a = elu1(fmaf(y2[(size_t)row * N_FILT + k], sc[k], sh[k]));
      if (drop.p > 0.f) a = dropout_keep(drop, (uint64_t)row * N_FILT + k) ? a * drop.scale : 0.f;
    }
    a2[r][k] = a;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < HB * N_FILT; idx += 256) {
    const int r = idx / N_FILT, e = idx % N_FILT, row = row0 + r;
    if (row < rows) {
      float acc = bc[e];
#pragma unroll 8
      for (int c = 0; c < N_FILT; ++c) acc = fmaf(swc[e * (N_FILT + 1) + c], a2[r][c], acc);
      feat[(size_t)row * N_FILT + e] = tf32_if(acc, rt);
    }
  }
}
int conv_head_fwd(const float* y2, const float* mean_rstd, const float* gamma, const float* beta, const float* wc,
                  const float* bc, float* feat, int B, DropoutCfg drop, cudaStream_t s) {
  ProfScope _ps("conv_head_fwd", s, (double)B * 36 * 3200.0, (double)B * 36 * 320.0);
  const int rows = B * N_POOL;
  conv_head_fwd_kernel<<<cdiv(rows, HB), 256, 0, s>>>(y2, mean_rstd, gamma, beta, wc, bc, feat, rows, drop, tf32_rounding());
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// backward of the head down to dz2 = d loss / d (BN2 output), plus dWc, dbc and the BN2 reduction sums.
// 16 (b,j) rows per pass.  Register tiling keeps the shared-memory traffic low (the first version issued 2 LDS per
// FMA and 3 shared atomics per element: 60-70 us for a 6 MB problem):
//   dA[r][k]  = sum_e Wc[e][k] dF[r][e] : thread = (channel k, row lane rq), up to 3 rows per thread, Wc[e][k] loaded once
//               per e for all of them; the three per-channel reductions stay in registers across all passes;
//   dWc[e][k] += sum_r dF[r][e] A2[r][k] : thread = (2 e) x (4 k) register block, one float4 + two scalar loads per row.
static constexpr int HS = 44;     // row stride of the staging tiles (16-byte aligned rows)
__global__ void __launch_bounds__(256) conv_head_bwd_kernel(const float* __restrict__ dfeat, const float* __restrict__ y2,
                                                            const float* __restrict__ mean_rstd,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const float* __restrict__ wc,
                                                            float* __restrict__ dz2, float* __restrict__ dwc,
                                                            float* __restrict__ dbc, double* __restrict__ bwd_sums,
                                                            int rows, DropoutCfg drop) {
  __shared__ __align__(16) float df[HB][HS], a2[HB][HS], zk[HB][HS], yhs[HB][HS];
  __shared__ float swc[N_FILT * N_FILT], sred[3][N_FILT];
  for (int i = threadIdx.x; i < N_FILT * N_FILT; i += 256) swc[i] = wc[i];
  if (threadIdx.x < 3 * N_FILT) (&sred[0][0])[threadIdx.x] = 0.f;
  // phase-2 role: channel k2, row lane rq (rows rq, rq+6, rq+12); threads 240..255 idle there
  const int k2 = threadIdx.x % N_FILT, rq = threadIdx.x / N_FILT;
  // phase-3 role: e pair (e0, e0+1) x k quad (k0..k0+3); threads 200..255 idle there
  const int e0 = (threadIdx.x / 10) * 2, k0 = (threadIdx.x % 10) * 4;
  const bool p3 = threadIdx.x < 200;
  float accw[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a) accw[a][0] = accw[a][1] = accw[a][2] = accw[a][3] = 0.f;
  float s_dz = 0.f, s_dzy = 0.f, s_df = 0.f;
  __syncthreads();
  for (int row0 = blockIdx.x * HB; row0 < rows; row0 += gridDim.x * HB) {
    for (int idx = threadIdx.x; idx < HB * N_FILT; idx += 256) {
      const int r = idx / N_FILT, k = idx % N_FILT, row = row0 + r;
      float d = 0.f, a = 0.f, zz = 0.f, yh = 0.f;
      if (row < rows) {
        d = dfeat[(size_t)row * N_FILT + k];
        yh = (y2[(size_t)row * N_FILT + k] - mean_rstd[k]) * mean_rstd[N_FILT + k];
        const float z = yh * gamma[k] + beta[k];
        float keep = 1.f;
        if (drop.p > 0.f) keep = dropout_keep(drop, (uint64_t)row * N_FILT + k) ? drop.scale : 0.f;
        a = elu1(z) * keep;
        zz = keep * elu1_grad(z);
      }
      df[r][k] = d; a2[r][k] = a; zk[r][k] = zz; yhs[r][k] = yh;
    }
    __syncthreads();
    if (rq < 6) {
      float da[3] = {0.f, 0.f, 0.f};
#pragma unroll 8
      for (int e = 0; e < N_FILT; ++e) {
        const float w = swc[e * N_FILT + k2];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int r = rq + 6 * q;
          if (r < HB) da[q] = fmaf(w, df[r][e], da[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int r = rq + 6 * q, row = row0 + r;
        if (r < HB && row < rows) {
          const float dz = da[q] * zk[r][k2];
          dz2[(size_t)row * N_FILT + k2] = dz;
          s_dz += dz;
          s_dzy = fmaf(dz, yhs[r][k2], s_dzy);
          s_df += df[r][k2];
        }
      }
    }
    if (p3) {
#pragma unroll 4
      for (int r = 0; r < HB; ++r) {
        const float f0 = df[r][e0], f1 = df[r][e0 + 1];
        const float4 av = *reinterpret_cast<const float4*>(&a2[r][k0]);
        accw[0][0] = fmaf(f0, av.x, accw[0][0]); accw[0][1] = fmaf(f0, av.y, accw[0][1]);
        accw[0][2] = fmaf(f0, av.z, accw[0][2]); accw[0][3] = fmaf(f0, av.w, accw[0][3]);
        accw[1][0] = fmaf(f1, av.x, accw[1][0]); accw[1][1] = fmaf(f1, av.y, accw[1][1]);
        accw[1][2] = fmaf(f1, av.z, accw[1][2]); accw[1][3] = fmaf(f1, av.w, accw[1][3]);
      }
    }
    __syncthreads();
  }
  if (p3) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int q = 0; q < 4; ++q) atomicAdd(&dwc[(e0 + a) * N_FILT + k0 + q], accw[a][q]);
  }
  if (rq < 6) {
    atomicAdd(&sred[0][k2], s_dz);
    atomicAdd(&sred[1][k2], s_dzy);
    atomicAdd(&sred[2][k2], s_df);
  }
  __syncthreads();
  if (threadIdx.x < N_FILT) {
    const int k = threadIdx.x;
    atomicAdd(&dbc[k], sred[2][k]);
    atomicAdd(&bwd_sums[k], (double)sred[0][k]);
    atomicAdd(&bwd_sums[N_FILT + k], (double)sred[1][k]);
  }
}
int conv_head_bwd(const float* dfeat, const float* y2, const float* mean_rstd, const float* gamma, const float* beta,
                  const float* wc, float* dz2, float* dwc, float* dbc, double* bwd_sums, int B, DropoutCfg drop,
                  cudaStream_t s) {
  ProfScope _ps("conv_head_bwd", s, (double)B * 36 * 6400.0, (double)B * 36 * 480.0);
  const int rows = B * N_POOL;
  int blocks = cdiv(rows, HB);
  if (blocks > 148 * 4) blocks = 148 * 4;
  conv_head_bwd_kernel<<<blocks, 256, 0, s>>>(dfeat, y2, mean_rstd, gamma, beta, wc, dz2, dwc, dbc, bwd_sums, rows, drop);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// BatchNorm backward (train mode) given the two reductions S1 = sum dz, S2 = sum dz*yhat per channel
__global__ void bn_bwd_apply_kernel(const float4* __restrict__ dz, const float4* __restrict__ y,
                                    const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                    const double* __restrict__ bwd_sums, long long count, float4* __restrict__ dy,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, long long n4, int round_tf,
                                    float gscale) {
  __shared__ float mu[N_FILT], rs[N_FILT], gr[N_FILT], m1[N_FILT], m2[N_FILT];
  if (threadIdx.x < N_FILT) {
    const int k = threadIdx.x;
    mu[k] = mean_rstd[k];
    rs[k] = mean_rstd[N_FILT + k];
    gr[k] = gamma[k] * rs[k];
    m1[k] = (float)(bwd_sums[k] / (double)count);
    m2[k] = (float)(bwd_sums[N_FILT + k] / (double)count);
    if (blockIdx.x == 0) {
      dgamma[k] += gscale * (float)bwd_sums[N_FILT + k];
      dbeta[k] += gscale * (float)bwd_sums[k];
    }
  }
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i * 4) % N_FILT);
    const float4 g = dz[i];
    const float4 v = y[i];
    float4 o;
    o.x = gr[c] * (g.x - m1[c] - (v.x - mu[c]) * rs[c] * m2[c]);
    o.y = gr[c + 1] * (g.y - m1[c + 1] - (v.y - mu[c + 1]) * rs[c + 1] * m2[c + 1]);
    o.z = gr[c + 2] * (g.z - m1[c + 2] - (v.z - mu[c + 2]) * rs[c + 2] * m2[c + 2]);
    o.w = gr[c + 3] * (g.w - m1[c + 3] - (v.w - mu[c + 3]) * rs[c + 3] * m2[c + 3]);
    if (round_tf) { o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w); }
    dy[i] = o;
  }
}
int bn_bwd_apply(const float* dz, const float* y, const float* mean_rstd, const float* gamma, const double* bwd_sums,
                 long long count, float* dy, float* dgamma, float* dbeta, long long n, int round_tf, float gscale,
                 cudaStream_t s) {
  ProfScope _ps("bn_bwd_apply", s, 0.0, (double)n * 12.0);
  const long long n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  bn_bwd_apply_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(dz), reinterpret_cast<const float4*>(y),
                                             mean_rstd, gamma, bwd_sums, count, reinterpret_cast<float4*>(dy), dgamma,
                                             dbeta, n4, round_tf, gscale);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// dz1 = da1 * ELU'(z1) (in place), z1 = BN1(y1); per-channel S1 = sum dz1, S2 = sum dz1*yhat1.
// 320 threads/block and 4 floats/thread keep each thread on a fixed channel quad (320*4 % 40 == 0).
__global__ void __launch_bounds__(320) bn1_bwd_reduce_kernel(float4* __restrict__ da1, const float4* __restrict__ y1,
                                                             const float* __restrict__ mean_rstd,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, double* __restrict__ bwd_sums,
                                                             long long n4) {
  __shared__ float red[2][N_FILT];
  if (threadIdx.x < 2 * N_FILT) (&red[0][0])[threadIdx.x] = 0.f;
  const int c = (int)((((long long)blockIdx.x * 320 + threadIdx.x) * 4) % N_FILT);
  float mu[4], rs[4], ga[4], be[4], s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    mu[q] = mean_rstd[c + q]; rs[q] = mean_rstd[N_FILT + c + q]; ga[q] = gamma[c + q]; be[q] = beta[c + q];
  }
  __syncthreads();
  for (long long i = blockIdx.x * 320LL + threadIdx.x; i < n4; i += (long long)gridDim.x * 320) {
    float4 g = da1[i];
    const float4 v = y1[i];
    float gg[4] = {g.x, g.y, g.z, g.w};
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float yh = (vv[q] - mu[q]) * rs[q];
      const float z = fmaf(yh, ga[q], be[q]);
      gg[q] *= elu1_grad(z);
      s1[q] += gg[q];
      s2[q] = fmaf(gg[q], yh, s2[q]);
    }
    da1[i] = make_float4(gg[0], gg[1], gg[2], gg[3]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    atomicAdd(&red[0][c + q], s1[q]);
    atomicAdd(&red[1][c + q], s2[q]);
  }
  __syncthreads();
  if (threadIdx.x < N_FILT) {
    atomicAdd(&bwd_sums[threadIdx.x], (double)red[0][threadIdx.x]);
    atomicAdd(&bwd_sums[N_FILT + threadIdx.x], (double)red[1][threadIdx.x]);
  }
}
int bn1_bwd_reduce(float* da1, const float* y1, const float* mean_rstd, const float* gamma, const float* beta,
                   double* bwd_sums, long long n, cudaStream_t s) {
  ProfScope _ps("bn1_bwd_reduce", s, 0.0, (double)n * 12.0);
  const long long n4 = n / 4;
  // grid stride must keep the channel quad fixed: gridDim*320*4 % 40 == 0 holds for every gridDim
  int blocks = (int)((n4 + 319) / 320);
  if (blocks > 148 * 6) blocks = 148 * 6;
  if (blocks < 1) blocks = 1;
  bn1_bwd_reduce_kernel<<<blocks, 320, 0, s>>>(reinterpret_cast<float4*>(da1), reinterpret_cast<const float4*>(y1),
                                               mean_rstd, gamma, beta, bwd_sums, n4);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// BN1 backward apply fused with the transposed temporal conv: dX3, dWt, dbt (+ BN1 dgamma/dbeta)
__global__ void __launch_bounds__(CT_THREADS) conv_temporal_bwd_kernel(
    const float* __restrict__ dz1, const float* __restrict__ y1, const float* __restrict__ x3,
    const float* __restrict__ wt, const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
    const double* __restrict__ bwd_sums, long long count, float* __restrict__ dx3, float* __restrict__ dwt,
    float* __restrict__ dbt, float* __restrict__ dgamma, float* __restrict__ dbeta, float gscale) {
  __shared__ float xs[256 + 64];
  __shared__ float ps[N_PSUM];
  __shared__ float dys[N_POOL * N_FILT];
  __shared__ float dps[N_PSUM + 64];
  __shared__ float sw[N_FILT * K_TEMP];
  const int b = blockIdx.x;
  const int k = threadIdx.x % N_FILT, jq = threadIdx.x / N_FILT;
  for (int i = threadIdx.x; i < N_FILT * K_TEMP; i += CT_THREADS) sw[i] = wt[i];
  const float mu = mean_rstd[k], rs = mean_rstd[N_FILT + k];
  const float gr = gamma[k] * rs;
  const float m1 = (float)(bwd_sums[k] / (double)count);
  const float m2 = (float)(bwd_sums[N_FILT + k] / (double)count);
  if (b == 0 && threadIdx.x < N_FILT) {
    dgamma[k] += gscale * (float)bwd_sums[N_FILT + k];
    dbeta[k] += gscale * (float)bwd_sums[k];
  }
  float accw[7];
#pragma unroll
  for (int q = 0; q < 7; ++q) accw[q] = 0.f;
  float accb = 0.f;
  for (int t = threadIdx.x; t < 64; t += CT_THREADS) dps[N_PSUM + t] = 0.f;

  for (int r = 0; r < N_CH; ++r) {
    load_row_and_pool(x3 + ((size_t)b * N_TOK + r) * D_PAD, xs, ps);
    // dy1 for this (b, r): [36][40]
#pragma unroll 3
    for (int jj = 0; jj < N_POOL / 4; ++jj) {
      const int j = jq + 4 * jj;
      const size_t idx = (((size_t)b * N_POOL + j) * N_CH + r) * N_FILT + k;
      const float yh = (y1[idx] - mu) * rs;
      const float dy = gr * (dz1[idx] - m1 - yh * m2);
      dys[j * N_FILT + k] = dy;
      accb += dy;
    }
    __syncthreads();
    // dWt[k][i] += sum_j dy[j][k] * p[5j+i]
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const int idx = threadIdx.x + CT_THREADS * q;
      if (idx < N_FILT * K_TEMP) {
        const int kk = idx / K_TEMP, i = idx % K_TEMP;
        float a = accw[q];
        for (int j = 0; j < N_POOL; ++j) a = fmaf(dys[j * N_FILT + kk], ps[5 * j + i], a);
        accw[q] = a;
      }
    }
    // dp[s] = sum_{i == s mod 5} sum_k w[k][i] * dy[(s-i)/5][k]
    for (int sidx = threadIdx.x; sidx < N_PSUM; sidx += CT_THREADS) {
      float a = 0.f;
      for (int i = sidx % 5; i < K_TEMP; i += 5) {
        const int j = (sidx - i) / 5;
        if (sidx - i >= 0 && j < N_POOL) {
          const float* dyr = dys + j * N_FILT;
#pragma unroll 8
          for (int kk = 0; kk < N_FILT; ++kk) a = fmaf(sw[kk * K_TEMP + i], dyr[kk], a);
        }
      }
      dps[sidx] = a;
    }
    __syncthreads();
    // dx[t] = (1/51) * sum_{s = max(0,t-50)}^{min(t,199)} dp[s]
    float* dxr = dx3 + ((size_t)b * N_TOK + r) * D_PAD;
    for (int t = threadIdx.x; t < D_PAD; t += CT_THREADS) {
      float a = 0.f;
      if (t < N_T) {
        const int lo = t - (K_POOL - 1) > 0 ? t - (K_POOL - 1) : 0;
        const int hi = t < N_PSUM - 1 ? t : N_PSUM - 1;
        for (int sidx = lo; sidx <= hi; ++sidx) a += dps[sidx];
        a *= (1.f / K_POOL);
      }
      dxr[t] = a;
    }
    __syncthreads();
  }
  // token 63 (channel 62) never reaches the conv stack (enc_out[:, :63], ATMS_retrieval.py:91)
  for (int t = threadIdx.x; t < D_PAD; t += CT_THREADS) dx3[((size_t)b * N_TOK + N_CH) * D_PAD + t] = 0.f;
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    const int idx = threadIdx.x + CT_THREADS * q;
    if (idx < N_FILT * K_TEMP) atomicAdd(&dwt[idx], accw[q] * (1.f / K_POOL));
  }
  // accb holds this thread's partial over its (jq) quarter; reduce the 4 quarters through atomics
  atomicAdd(&dbt[k], accb);
}
int conv_temporal_bwd_simt(const float* dz1, const float* y1, const float* x3, const float* wt, const float* mean_rstd,
                      const float* gamma, const double* bwd_sums, long long count, float* dx3, float* dwt, float* dbt,
                      float* dgamma, float* dbeta, int B, float gscale, cudaStream_t s) {
  ProfScope _ps("conv_temporal_bwd_simt", s, (double)B * 63 * 36 * 40 * 100.0, (double)B * (36 * 2520 * 8.0 + 63 * 2000.0));
  conv_temporal_bwd_kernel<<<B, CT_THREADS, 0, s>>>(dz1, y1, x3, wt, mean_rstd, gamma, bwd_sums, count, dx3, dwt, dbt,
                                                    dgamma, dbeta, gscale);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace eegb200

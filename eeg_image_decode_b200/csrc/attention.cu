// Multi-head self-attention core over 64 channel-tokens (FullAttention.forward,
// models/subject_layers/SelfAttention_Family.py:56-75): per (sample, head) softmax(Q K^T / sqrt(62)) V with the
// attention-probability dropout; backward by recomputation (no probability tensor is stored).
// One CTA per (b, h); Q/K/V (64 x 62) live in shared memory.  fp32 CUDA-core math (4 % of the step's FLOPs).
#include "kernels.h"

namespace eegb200 {

static constexpr int LDS = 65;           // padded smem row stride
static constexpr int ATT_THREADS = 256;

__device__ __forceinline__ void load_head(const float* __restrict__ base, int ld, float* __restrict__ dst) {
  // 64 rows x 64 floats (float4 per thread, 16 threads per row)
  const int r0 = threadIdx.x >> 4, c4 = threadIdx.x & 15;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int r = r0 + 16 * p;
    const float4 v = *reinterpret_cast<const float4*>(base + (size_t)r * ld + c4 * 4);
    float* d = dst + r * LDS + c4 * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
}

// scores + softmax for row i, columns j = jj*4 + g.  Returns normalised probabilities in p[16].
__device__ __forceinline__ void softmax_row(const float* __restrict__ Q, const float* __restrict__ K, int i, int g,
                                            float p[16]) {
  float acc[16];
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) acc[jj] = 0.f;
  for (int e = 0; e < D_HEAD; ++e) {
    const float q = Q[i * LDS + e];
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) acc[jj] = fmaf(q, K[(jj * 4 + g) * LDS + e], acc[jj]);
  }
  const float scale = 0.12700012700019050f;   // 1/sqrt(62)
  float mx = -INFINITY;
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) { acc[jj] *= scale; mx = fmaxf(mx, acc[jj]); }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  float sum = 0.f;
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) { p[jj] = __expf(acc[jj] - mx); sum += p[jj]; }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  const float inv = 1.f / sum;
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) p[jj] *= inv;
}

__global__ void __launch_bounds__(ATT_THREADS) attention_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ o,
                                                                    DropoutCfg drop, int rt) {
  extern __shared__ float sm[];
  float* Q = sm;
  float* K = Q + 64 * LDS;
  float* V = K + 64 * LDS;
  float* P = V + 64 * LDS;
  const int b = blockIdx.x >> 2, h = blockIdx.x & 3;
  const float* base = qkv + (size_t)b * 64 * 768 + h * 64;
  load_head(base, 768, Q);
  load_head(base + 256, 768, K);
  load_head(base + 512, 768, V);
  __syncthreads();
  const int i = threadIdx.x >> 2, g = threadIdx.x & 3;
  float p[16];
  softmax_row(Q, K, i, g, p);
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) {
    const int j = jj * 4 + g;
    float v = p[jj];
    if (drop.p > 0.f) v = dropout_keep(drop, ((uint64_t)blockIdx.x * 64 + i) * 64 + j) ? v * drop.scale : 0.f;
    P[i * LDS + j] = v;
  }
  __syncthreads();
  float acc[16];
#pragma unroll
  for (int ee = 0; ee < 16; ++ee) acc[ee] = 0.f;
  for (int j = 0; j < 64; ++j) {
    const float pv = P[i * LDS + j];
#pragma unroll
    for (int ee = 0; ee < 16; ++ee) acc[ee] = fmaf(pv, V[j * LDS + ee * 4 + g], acc[ee]);
  }
  float* orow = o + ((size_t)b * 64 + i) * 256 + h * 64;
#pragma unroll
  for (int ee = 0; ee < 16; ++ee) {
    const int e = ee * 4 + g;
    orow[e] = e < D_HEAD ? tf32_if(acc[ee], rt) : 0.f;
  }
}

__global__ void __launch_bounds__(ATT_THREADS) attention_bwd_kernel(const float* __restrict__ qkv,
                                                                    const float* __restrict__ d_o,
                                                                    float* __restrict__ dqkv, DropoutCfg drop, int rt) {
  extern __shared__ float sm[];
  float* Q = sm;
  float* K = Q + 64 * LDS;
  float* V = K + 64 * LDS;
  float* DO = V + 64 * LDS;
  float* X = DO + 64 * LDS;
  const int b = blockIdx.x >> 2, h = blockIdx.x & 3;
  const float* base = qkv + (size_t)b * 64 * 768 + h * 64;
  load_head(base, 768, Q);
  load_head(base + 256, 768, K);
  load_head(base + 512, 768, V);
  load_head(d_o + (size_t)b * 64 * 256 + h * 64, 256, DO);
  __syncthreads();
  const int i = threadIdx.x >> 2, g = threadIdx.x & 3;
  float p[16];
  softmax_row(Q, K, i, g, p);
  // dPd[i][j] = dO[i,:] . V[j,:]
  float dp[16];
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) dp[jj] = 0.f;
  for (int e = 0; e < D_HEAD; ++e) {
    const float d = DO[i * LDS + e];
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) dp[jj] = fmaf(d, V[(jj * 4 + g) * LDS + e], dp[jj]);
  }
  float rowdot = 0.f;
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) {
    const int j = jj * 4 + g;
    float keep = 1.f;
    if (drop.p > 0.f) keep = dropout_keep(drop, ((uint64_t)blockIdx.x * 64 + i) * 64 + j) ? drop.scale : 0.f;
    X[i * LDS + j] = p[jj] * keep;          // dropped probabilities (for dV)
    dp[jj] *= keep;                         // gradient wrt the un-dropped probabilities
    rowdot = fmaf(dp[jj], p[jj], rowdot);
  }
  rowdot += __shfl_xor_sync(0xffffffffu, rowdot, 1);
  rowdot += __shfl_xor_sync(0xffffffffu, rowdot, 2);
  const float scale = 0.12700012700019050f;
  float ds[16];
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) ds[jj] = p[jj] * (dp[jj] - rowdot) * scale;
  __syncthreads();
  // dV[j][e] = sum_i Pd[i][j] dO[i][e]     (thread: j = i-index of this thread, e = ee*4+g)
  {
    const int j = i;
    float acc[16];
#pragma unroll
    for (int ee = 0; ee < 16; ++ee) acc[ee] = 0.f;
    for (int r = 0; r < 64; ++r) {
      const float pv = X[r * LDS + j];
#pragma unroll
      for (int ee = 0; ee < 16; ++ee) acc[ee] = fmaf(pv, DO[r * LDS + ee * 4 + g], acc[ee]);
    }
    float* out = dqkv + ((size_t)b * 64 + j) * 768 + 512 + h * 64;
#pragma unroll
    for (int ee = 0; ee < 16; ++ee) {
      const int e = ee * 4 + g;
      out[e] = e < D_HEAD ? tf32_if(acc[ee], rt) : 0.f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) X[i * LDS + jj * 4 + g] = ds[jj];
  __syncthreads();
  {
    // dQ[i][e] = sum_j dS[i][j] K[j][e];  dK[j][e] = sum_i dS[i][j] Q[i][e]
    float aq[16], ak[16];
#pragma unroll
    for (int ee = 0; ee < 16; ++ee) aq[ee] = ak[ee] = 0.f;
    for (int r = 0; r < 64; ++r) {
      const float s_ir = X[i * LDS + r];   // dS[i][r]
      const float s_ri = X[r * LDS + i];   // dS[r][i]
#pragma unroll
      for (int ee = 0; ee < 16; ++ee) {
        aq[ee] = fmaf(s_ir, K[r * LDS + ee * 4 + g], aq[ee]);
        ak[ee] = fmaf(s_ri, Q[r * LDS + ee * 4 + g], ak[ee]);
      }
    }
    float* oq = dqkv + ((size_t)b * 64 + i) * 768 + h * 64;
    float* ok = oq + 256;
#pragma unroll
    for (int ee = 0; ee < 16; ++ee) {
      const int e = ee * 4 + g;
      oq[e] = e < D_HEAD ? tf32_if(aq[ee], rt) : 0.f;
      ok[e] = e < D_HEAD ? tf32_if(ak[ee], rt) : 0.f;
    }
  }
}

int attention_fwd_simt(const float* qkv, float* o, int B, DropoutCfg drop, cudaStream_t s) {
  ProfScope _ps("attention_fwd_simt", s, (double)B * 4 * 4.0 * 64 * 64 * 62, (double)B * 64 * 1024 * 4.0);
  const size_t smem = 4 * 64 * LDS * sizeof(float);
  static PerDeviceOnce once;
  if (once.first()) {
    EEG_CUDA_OK(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  attention_fwd_kernel<<<B * N_HEAD, ATT_THREADS, smem, s>>>(qkv, o, drop, tf32_rounding());
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int attention_bwd_simt(const float* qkv, const float* d_o, float* dqkv, int B, DropoutCfg drop, cudaStream_t s) {
  ProfScope _ps("attention_bwd_simt", s, (double)B * 4 * 12.0 * 64 * 64 * 62, (double)B * 64 * 1792 * 4.0);
  const size_t smem = 5 * 64 * LDS * sizeof(float);
  static PerDeviceOnce once;
  if (once.first()) {
    EEG_CUDA_OK(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  attention_bwd_kernel<<<B * N_HEAD, ATT_THREADS, smem, s>>>(qkv, d_o, dqkv, drop, tf32_rounding());
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace eegb200

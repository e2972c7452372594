// Row-wise / elementwise kernels around the GEMMs: input padding, subject token, LayerNorm fwd/bwd,
// column sums (bias gradients), dropout-mask dump, TF32 rounding copies.
#include "kernels.h"

namespace eegb200 {

// ------------------------------------------------------------------------------------------------
// x (B,63,250) -> token matrix Xp [B*64, 256]: row b*64+0 = 0 (subject-token slot), rows 1..63 = channels,
// columns 250..255 = 0; values rounded to TF32 (they only feed the value-embedding GEMM, Embed.py:146).
// ------------------------------------------------------------------------------------------------
__global__ void pad_input_kernel(const float* __restrict__ x, float* __restrict__ xp, int B, int rt) {
  const long long total = (long long)B * 64 * 64;   // float4 slots
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i & 63);
    const long long row = i >> 6;
    const int t = (int)(row & 63);
    const long long b = row >> 6;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t > 0) {
      const float* src = x + (b * 63 + (t - 1)) * 250;
      const int c = c4 * 4;
      if (c < 250) v.x = tf32_if(src[c], rt);
      if (c + 1 < 250) v.y = tf32_if(src[c + 1], rt);
      if (c + 2 < 250) v.z = tf32_if(src[c + 2], rt);
      if (c + 3 < 250) v.w = tf32_if(src[c + 3], rt);
    }
    reinterpret_cast<float4*>(xp)[i] = v;
  }
}
int pad_input(const float* x, float* xp, int B, cudaStream_t s) {
  ProfScope _ps("pad_input", s, 0.0, (double)B * (63000.0 + 65536.0));
  const long long total = (long long)B * 64 * 64;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  pad_input_kernel<<<blocks, 256, 0, s>>>(x, xp, B, tf32_rounding());
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// flag[0] = 1 if any subject id >= n_subj or < 0 (SubjectEmbedding.forward, Embed.py:117: the WHOLE batch then
// uses the shared token)
__global__ void subject_flag_kernel(const long long* __restrict__ ids, int B, int n_subj, int* __restrict__ flag) {
  __shared__ int any;
  if (threadIdx.x == 0) any = 0;
  __syncthreads();
  int a = 0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) a |= (ids[i] >= n_subj || ids[i] < 0) ? 1 : 0;
  if (a) atomicOr(&any, 1);
  __syncthreads();
  if (threadIdx.x == 0) flag[0] = any;
}
// H0[b*64 + 0, :] = dropout(subject row)   (Embed.py:158-162)
__global__ void subject_token_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                     const float* __restrict__ shared_tok, const int* __restrict__ flag,
                                     float* __restrict__ h0, int B, DropoutCfg drop, int round_tf) {
  const int b = blockIdx.x;
  const int c = threadIdx.x;   // 256 threads
  const float* src = flag[0] ? shared_tok : table + ids[b] * 250;
  float v = c < 250 ? src[c] : 0.f;
  const size_t row = (size_t)b * 64;
  if (drop.p > 0.f) v = dropout_keep(drop, row * 256 + c) ? v * drop.scale : 0.f;
  h0[row * 256 + c] = round_tf ? tf32_rn(v) : v;
}
int subject_token(const long long* ids, const float* table, const float* shared_tok, int n_subj, int* flag, float* h0,
                  int B, DropoutCfg drop, int round_tf, cudaStream_t s) {
  ProfScope _ps("subject_token", s, 0.0, (double)B * 2000.0);
  subject_flag_kernel<<<1, 256, 0, s>>>(ids, B, n_subj, flag);
  subject_token_kernel<<<B, 256, 0, s>>>(ids, table, shared_tok, flag, h0, B, drop, round_tf);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch(2);
  return 0;
}
// backward of the subject token: g[b*64+0, :] (after the embed dropout mask) is scattered into the table
// (dense grad, unused rows stay 0) or summed into the shared token.
__global__ void subject_token_bwd_kernel(const long long* __restrict__ ids, const int* __restrict__ flag,
                                         const float* __restrict__ dh0, float* __restrict__ dtable,
                                         float* __restrict__ dshared, int B, DropoutCfg drop) {
  const int b = blockIdx.x;
  const int c = threadIdx.x;
  if (c >= 250) return;
  const size_t row = (size_t)b * 64;
  float g = dh0[row * 256 + c];
  if (drop.p > 0.f) g = dropout_keep(drop, row * 256 + c) ? g * drop.scale : 0.f;
  if (flag[0]) atomicAdd(&dshared[c], g);
  else atomicAdd(&dtable[ids[b] * 250 + c], g);
}
int subject_token_bwd(const long long* ids, const int* flag, const float* dh0, float* dtable, float* dshared, int B,
                      DropoutCfg drop, cudaStream_t s) {
  ProfScope _ps("subject_token_bwd", s, 0.0, (double)B * 2000.0);
  subject_token_bwd_kernel<<<B, 256, 0, s>>>(ids, flag, dh0, dtable, dshared, B, drop);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim D (valid columns), row stride ld; one warp per row.
// stats[row] = (mean, rstd).  Columns D..ld_out-1 of the output are written as 0.
// `second`: optional chained LayerNorm (EncoderLayer.norm2 followed by Encoder.norm,
// Transformer_EncDec.py:51,77-78) applied to the first one's output.
// ------------------------------------------------------------------------------------------------
template <int MAXV>   // MAXV = ceil(D/32) values per lane held in registers
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, int ld, int rows, int D, const float* __restrict__ g1,
                                     const float* __restrict__ b1, float* __restrict__ stats1,
                                     const float* __restrict__ g2, const float* __restrict__ b2,
                                     float* __restrict__ stats2, float* __restrict__ y, int ld_out, int round_tf,
                                     float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (size_t)warp * ld;
  float v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < D ? xr[c] : 0.f;
    s += v[i];
  }
  float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    const float d = c < D ? v[i] - mean : 0.f;
    q += d * d;
  }
  float rstd = rsqrtf(warp_sum(q) / D + eps);
  if (lane == 0) { stats1[2 * warp] = mean; stats1[2 * warp + 1] = rstd; }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < D ? (v[i] - mean) * rstd * g1[c] + b1[c] : 0.f;
  }
  if (g2 != nullptr) {
    s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) s += v[i];
    mean = warp_sum(s) / D;
    q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + 32 * i;
      const float d = c < D ? v[i] - mean : 0.f;
      q += d * d;
    }
    rstd = rsqrtf(warp_sum(q) / D + eps);
    if (lane == 0) { stats2[2 * warp] = mean; stats2[2 * warp + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < D ? (v[i] - mean) * rstd * g2[c] + b2[c] : 0.f;
    }
  }
  float* yr = y + (size_t)warp * ld_out;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < ld_out) yr[c] = round_tf ? tf32_rn(v[i]) : v[i];
  }
}

int layernorm_fwd(const float* x, int ld, int rows, int D, const float* g1, const float* b1, float* stats1,
                  const float* g2, const float* b2, float* stats2, float* y, int ld_out, int round_tf, cudaStream_t s) {
  ProfScope _ps(D > 256 ? "layernorm_fwd_1024" : (g2 ? "layernorm2x_fwd" : "layernorm_fwd"), s, 0.0, (double)rows * (ld + ld_out) * 4.0);
  const int threads = 256;
  const int blocks = cdiv(rows * 32, threads);
  if (D <= 256 && ld_out <= 256)
    layernorm_fwd_kernel<8><<<blocks, threads, 0, s>>>(x, ld, rows, D, g1, b1, stats1, g2, b2, stats2, y, ld_out, round_tf, 1e-5f);
  else if (D <= 1024 && ld_out <= 1024)
    layernorm_fwd_kernel<32><<<blocks, threads, 0, s>>>(x, ld, rows, D, g1, b1, stats1, g2, b2, stats2, y, ld_out, round_tf, 1e-5f);
  else {
    set_error("layernorm_fwd: D=%d unsupported", D);
    return 2;
  }
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// LayerNorm backward (optionally through the chained pair).  dy: gradient wrt the final output.
// x: input of the first LN.  Writes dx; accumulates dgamma/dbeta (atomics on per-block partials).
template <int MAXV>
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, int ld_dy, const float* __restrict__ x, int ld,
                                     int rows, int D, const float* __restrict__ g1, const float* __restrict__ b1,
                                     const float* __restrict__ stats1, const float* __restrict__ g2,
                                     const float* __restrict__ stats2, float* __restrict__ dx, int ld_dx,
                                     float* __restrict__ dg1, float* __restrict__ db1, float* __restrict__ dg2,
                                     float* __restrict__ db2, int round_tf) {
  // per-block partial sums of the affine gradients live in shared memory: [4][MAXV*32]
  extern __shared__ float sh[];
  const int W = MAXV * 32;
  for (int i = threadIdx.x; i < 4 * W; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  float acc_g1[MAXV], acc_b1[MAXV], acc_g2[MAXV], acc_b2[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) acc_g1[i] = acc_b1[i] = acc_g2[i] = acc_b2[i] = 0.f;

  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
    const float* xr = x + (size_t)row * ld;
    const float* dyr = dy + (size_t)row * ld_dy;
    const float m1 = stats1[2 * row], r1 = stats1[2 * row + 1];
    float xh1[MAXV], g[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + 32 * i;
      xh1[i] = c < D ? (xr[c] - m1) * r1 : 0.f;
      g[i] = c < D ? dyr[c] : 0.f;
    }
    if (g2 != nullptr) {
      // second LN: its input is u = xh1*g1 + b1
      const float m2 = stats2[2 * row], r2 = stats2[2 * row + 1];
      float s1 = 0.f, s2 = 0.f;
      float xh2[MAXV];
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        const float u = c < D ? xh1[i] * g1[c] + b1[c] : 0.f;
        xh2[i] = c < D ? (u - m2) * r2 : 0.f;
        acc_g2[i] += g[i] * xh2[i];
        acc_b2[i] += g[i];
        const float dxh = c < D ? g[i] * g2[c] : 0.f;
        g[i] = dxh;
        s1 += dxh;
        s2 += dxh * xh2[i];
      }
      s1 = warp_sum(s1) / D;
      s2 = warp_sum(s2) / D;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        g[i] = c < D ? r2 * (g[i] - s1 - xh2[i] * s2) : 0.f;
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + 32 * i;
      acc_g1[i] += g[i] * xh1[i];
      acc_b1[i] += g[i];
      const float dxh = c < D ? g[i] * g1[c] : 0.f;
      g[i] = dxh;
      s1 += dxh;
      s2 += dxh * xh1[i];
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    float* dxr = dx + (size_t)row * ld_dx;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < ld_dx) {
        const float v = c < D ? r1 * (g[i] - s1 - xh1[i] * s2) : 0.f;
        dxr[c] = round_tf ? tf32_rn(v) : v;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    atomicAdd(&sh[c], acc_g1[i]);
    atomicAdd(&sh[W + c], acc_b1[i]);
    if (g2 != nullptr) {
      atomicAdd(&sh[2 * W + c], acc_g2[i]);
      atomicAdd(&sh[3 * W + c], acc_b2[i]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    atomicAdd(&dg1[c], sh[c]);
    atomicAdd(&db1[c], sh[W + c]);
    if (g2 != nullptr) {
      atomicAdd(&dg2[c], sh[2 * W + c]);
      atomicAdd(&db2[c], sh[3 * W + c]);
    }
  }
}

int layernorm_bwd(const float* dy, int ld_dy, const float* x, int ld, int rows, int D, const float* g1, const float* b1,
                  const float* stats1, const float* g2, const float* stats2, float* dx, int ld_dx, float* dg1,
                  float* db1, float* dg2, float* db2, int round_tf, cudaStream_t s) {
  ProfScope _ps(D > 256 ? "layernorm_bwd_1024" : (g2 ? "layernorm2x_bwd" : "layernorm_bwd"), s, 0.0, (double)rows * (ld_dy + ld + ld_dx) * 4.0);
  const int threads = 256;
  int blocks = cdiv(rows, threads / 32);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (D <= 256 && ld_dx <= 256)
    layernorm_bwd_kernel<8><<<blocks, threads, 4 * 256 * sizeof(float), s>>>(dy, ld_dy, x, ld, rows, D, g1, b1, stats1, g2,
                                                                            stats2, dx, ld_dx, dg1, db1, dg2, db2, round_tf);
  else if (D <= 1024 && ld_dx <= 1024)
    layernorm_bwd_kernel<32><<<blocks, threads, 4 * 1024 * sizeof(float), s>>>(dy, ld_dy, x, ld, rows, D, g1, b1, stats1,
                                                                              g2, stats2, dx, ld_dx, dg1, db1, dg2, db2,
                                                                              round_tf);
  else {
    set_error("layernorm_bwd: D=%d unsupported", D);
    return 2;
  }
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// Token-stream LayerNorm backward (all leading dimensions 256, D <= 256): one warp per row, each lane owns the column
// quads [4*lane, 4*lane+4) and [128+4*lane, ...), so every access is a 16-byte vector.  Optionally also writes
// drop_out = tf32?(dropout(dx)) -- the A operand of the next backward GEMM -- and accumulates its column sums (the
// bias gradient of the layer in front of that dropout); this replaces a separate 134 MB pass over dx.
template <bool TWO>
__global__ void __launch_bounds__(256) layernorm_bwd_tok_kernel(
    const float* __restrict__ dy, const float* __restrict__ x, int rows, int D, const float* __restrict__ g1,
    const float* __restrict__ b1, const float* __restrict__ stats1, const float* __restrict__ g2,
    const float* __restrict__ stats2, float* __restrict__ dx, float* __restrict__ dg1, float* __restrict__ db1,
    float* __restrict__ dg2, float* __restrict__ db2, float* __restrict__ drop_out, DropoutCfg cfg, int drop_round,
    float* __restrict__ colsum) {
  __shared__ float sh[5][256];
  for (int i = threadIdx.x; i < 5 * 256; i += blockDim.x) (&sh[0][0])[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int col[8];
  float gm1[8], bt1[8], gm2[8];
  bool ok[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    col[i] = (i < 4 ? 0 : 128) + 4 * lane + (i & 3);
    ok[i] = col[i] < D;
    gm1[i] = ok[i] ? g1[col[i]] : 0.f;
    bt1[i] = (TWO && ok[i]) ? b1[col[i]] : 0.f;
    gm2[i] = (TWO && ok[i]) ? g2[col[i]] : 0.f;
  }
  float acc_g1[8], acc_b1[8], acc_g2[8], acc_b2[8], acc_cs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc_g1[i] = acc_b1[i] = acc_g2[i] = acc_b2[i] = acc_cs[i] = 0.f;
  const float invD = 1.f / (float)D;

  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * 256);
    const float4* dyr = reinterpret_cast<const float4*>(dy + (size_t)row * 256);
    const float4 xa = xr[lane], xb = xr[32 + lane], ga = dyr[lane], gb = dyr[32 + lane];
    const float2 st1 = *reinterpret_cast<const float2*>(stats1 + 2 * row);
    const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
    float g[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    float xh1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      xh1[i] = ok[i] ? (xv[i] - st1.x) * st1.y : 0.f;
      g[i] = ok[i] ? g[i] : 0.f;
    }
    if (TWO) {
      const float2 st2 = *reinterpret_cast<const float2*>(stats2 + 2 * row);
      float s1 = 0.f, s2 = 0.f, xh2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float u = fmaf(xh1[i], gm1[i], bt1[i]);
        xh2[i] = ok[i] ? (u - st2.x) * st2.y : 0.f;
        acc_g2[i] = fmaf(g[i], xh2[i], acc_g2[i]);
        acc_b2[i] += g[i];
        g[i] *= gm2[i];
        s1 += g[i];
        s2 = fmaf(g[i], xh2[i], s2);
      }
      s1 = warp_sum(s1) * invD;
      s2 = warp_sum(s2) * invD;
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = ok[i] ? st2.y * (g[i] - s1 - xh2[i] * s2) : 0.f;
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc_g1[i] = fmaf(g[i], xh1[i], acc_g1[i]);
      acc_b1[i] += g[i];
      g[i] *= gm1[i];
      s1 += g[i];
      s2 = fmaf(g[i], xh1[i], s2);
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = ok[i] ? st1.y * (g[i] - s1 - xh1[i] * s2) : 0.f;
    float4* dxr = reinterpret_cast<float4*>(dx + (size_t)row * 256);
    dxr[lane] = make_float4(v[0], v[1], v[2], v[3]);
    dxr[32 + lane] = make_float4(v[4], v[5], v[6], v[7]);
    if (drop_out != nullptr) {
      if (cfg.p > 0.f) {
        const uint32_t m0 = dropout_keep4(cfg, (uint64_t)row * 256 + 4 * lane);
        const uint32_t m1 = dropout_keep4(cfg, (uint64_t)row * 256 + 128 + 4 * lane);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i] = (m0 >> i) & 1u ? v[i] * cfg.scale : 0.f;
          v[4 + i] = (m1 >> i) & 1u ? v[4 + i] * cfg.scale : 0.f;
        }
      }
      if (drop_round) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = tf32_rn(v[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc_cs[i] += v[i];
      float4* dr = reinterpret_cast<float4*>(drop_out + (size_t)row * 256);
      dr[lane] = make_float4(v[0], v[1], v[2], v[3]);
      dr[32 + lane] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (!ok[i]) continue;
    atomicAdd(&sh[0][col[i]], acc_g1[i]);
    atomicAdd(&sh[1][col[i]], acc_b1[i]);
    if (TWO) {
      atomicAdd(&sh[2][col[i]], acc_g2[i]);
      atomicAdd(&sh[3][col[i]], acc_b2[i]);
    }
    if (drop_out != nullptr) atomicAdd(&sh[4][col[i]], acc_cs[i]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    atomicAdd(&dg1[c], sh[0][c]);
    atomicAdd(&db1[c], sh[1][c]);
    if (TWO) {
      atomicAdd(&dg2[c], sh[2][c]);
      atomicAdd(&db2[c], sh[3][c]);
    }
    if (colsum != nullptr) atomicAdd(&colsum[c], sh[4][c]);
  }
}

int layernorm_bwd_tok(const float* dy, const float* x, int rows, int D, const float* g1, const float* b1,
                      const float* stats1, const float* g2, const float* stats2, float* dx, float* dg1, float* db1,
                      float* dg2, float* db2, float* drop_out, DropoutCfg cfg, int drop_round, float* colsum_out,
                      cudaStream_t s) {
  ProfScope _ps(g2 ? "layernorm2x_bwd" : "layernorm_bwd", s, 0.0, (double)rows * 256 * (drop_out ? 16.0 : 12.0));
  EEG_REQUIRE(D <= 256 && D > 0, "layernorm_bwd_tok: D=%d unsupported", D);
  EEG_REQUIRE(!g2 || (stats2 && dg2 && db2), "layernorm_bwd_tok: the chained form needs stats2/dg2/db2");
  int blocks = cdiv(rows, 8);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  if (g2)
    layernorm_bwd_tok_kernel<true><<<blocks, 256, 0, s>>>(dy, x, rows, D, g1, b1, stats1, g2, stats2, dx, dg1, db1, dg2, db2,
                                                         drop_out, cfg, drop_round, colsum_out);
  else
    layernorm_bwd_tok_kernel<false><<<blocks, 256, 0, s>>>(dy, x, rows, D, g1, b1, stats1, g2, stats2, dx, dg1, db1, dg2,
                                                          db2, drop_out, cfg, drop_round, colsum_out);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// out[c] += sum_r x[r*ld + c] (c < cols); optional dropout mask on the fly (bias grads of layers whose
// output passes through a dropout before the residual).
// ------------------------------------------------------------------------------------------------
// float4 columns x 16 row lanes per block.  blockIdx.y selects a 256-column chunk (one launch covers wide matrices such
// as dQKV [M,768]); every block owns a contiguous band of rows and keeps 8 predicated 16-byte loads in flight per thread
// (the former grid-strided version fell into a one-row-at-a-time tail loop for most of its rows: 2.3 TB/s).
__global__ void __launch_bounds__(1024) colsum_vec_kernel(const float* __restrict__ x, int ld, int rows, int cols,
                                                          float* __restrict__ out, int row_mod, int row_skip) {
  __shared__ float4 part[16][64];
  const int c4 = threadIdx.x;            // 64 float4 columns = 256 floats
  const int ry = threadIdx.y;            // 16 row lanes
  const int c0 = blockIdx.y * 256;
  const int w = cols - c0 < 256 ? cols - c0 : 256;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 * 4 < w) {
    const int per = (rows + gridDim.x - 1) / gridDim.x;
    const int r_begin = blockIdx.x * per;
    const int r_end = min(rows, r_begin + per);
    const float* base = x + c0 + c4 * 4;
    for (int r = r_begin + ry; r < r_end; r += 16 * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int rr = r + u * 16;
        v[u] = rr < r_end ? *reinterpret_cast<const float4*>(base + (size_t)rr * ld) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int rr = r + u * 16;
        if (row_mod > 0 && (rr % row_mod) == row_skip) continue;
        acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
      }
    }
  }
  part[ry][c4] = acc;
  __syncthreads();
  if (ry == 0 && c4 * 4 < w) {
    float4 t = part[0][c4];
    for (int i = 1; i < 16; ++i) { t.x += part[i][c4].x; t.y += part[i][c4].y; t.z += part[i][c4].z; t.w += part[i][c4].w; }
    const float tv[4] = {t.x, t.y, t.z, t.w};
    for (int q = 0; q < 4; ++q)
      if (c4 * 4 + q < w) atomicAdd(&out[c0 + c4 * 4 + q], tv[q]);
  }
}
__global__ void colsum_kernel(const float* __restrict__ x, int ld, int rows, int cols, float* __restrict__ out,
                              int row_mod, int row_skip) {
  __shared__ float part[8][257];
  const int c = threadIdx.x;
  float s = 0.f;
  for (int r = blockIdx.x * blockDim.y + threadIdx.y; r < rows; r += gridDim.x * blockDim.y) {
    if (row_mod > 0 && (r % row_mod) == row_skip) continue;
    if (c < cols) s += x[(size_t)r * ld + c];
  }
  part[threadIdx.y][c] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
    for (int i = 0; i < blockDim.y; ++i) t += part[i][c];
    atomicAdd(&out[c], t);
  }
}
// out[c] += sum_r x[r*ld + c]; rows with r % row_mod == row_skip are left out (row_mod 0: none).
// Reads up to the next multiple of 4 columns of each row: callers pass padded rows (ld >= cols rounded up to 4).
int colsum(const float* x, int ld, int rows, int cols, float* out, int row_mod, int row_skip, cudaStream_t s) {
  ProfScope _ps("colsum", s, 0.0, (double)rows * cols * 4.0);
  const bool vec = (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && ld >= (cols + 3) / 4 * 4;
  if (vec) {
    const int chunks = cdiv(cols, 256);
    int blocks = cdiv(rows, 16 * 8);
    const int cap = chunks >= 2 ? 148 : 148 * 2;      // 2 blocks of 1024 threads per SM; wide matrices fill the y dimension
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    colsum_vec_kernel<<<dim3(blocks, chunks), dim3(64, 16), 0, s>>>(x, ld, rows, cols, out, row_mod, row_skip);
    count_launch();
    EEG_CUDA_OK(cudaGetLastError());
    return 0;
  }
  for (int c0 = 0; c0 < cols; c0 += 256) {
    const int w = cols - c0 < 256 ? cols - c0 : 256;
    const int tx = (w + 31) / 32 * 32;
    const int ty = 1024 / tx > 8 ? 8 : 1024 / tx;
    int blocks = cdiv(rows, ty * 8);
    if (blocks > 148 * 2) blocks = 148 * 2;
    if (blocks < 1) blocks = 1;
    colsum_kernel<<<blocks, dim3(tx, ty), 0, s>>>(x + c0, ld, rows, w, out + c0, row_mod, row_skip);
    count_launch();
  }
  EEG_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void dropout_mask_kernel(DropoutCfg cfg, int rows, int cols, int ld, float* __restrict__ out) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i % cols);
    out[i] = (cfg.p <= 0.f || dropout_keep(cfg, (uint64_t)r * ld + c)) ? 1.f : 0.f;
  }
}
int dropout_mask(DropoutCfg cfg, int rows, int cols, int ld, float* out, cudaStream_t s) {
  const long long total = (long long)rows * cols;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  dropout_mask_kernel<<<blocks, 256, 0, s>>>(cfg, rows, cols, ld, out);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// dst[r*ld_dst + c] = (r < rows && c < cols) ? tf32(src[r*ld_src + c]) : 0     for r < rows_dst, c < ld_dst
__global__ void pad_copy_kernel(const float* __restrict__ src, int ld_src, int rows, int cols, float* __restrict__ dst,
                                int ld_dst, int rows_dst, int round_tf, float scale) {
  const long long total = (long long)rows_dst * ld_dst;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ld_dst;
    const int c = (int)(i % ld_dst);
    float v = (r < rows && c < cols) ? src[r * ld_src + c] * scale : 0.f;
    dst[i] = round_tf ? tf32_rn(v) : v;
  }
}
int pad_copy(const float* src, int ld_src, int rows, int cols, float* dst, int ld_dst, int rows_dst, int round_tf,
             float scale, cudaStream_t s) {
  ProfScope _ps("pad_copy", s, 0.0, (double)rows_dst * ld_dst * 8.0);
  const long long total = (long long)rows_dst * ld_dst;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  pad_copy_kernel<<<blocks, 256, 0, s>>>(src, ld_src, rows, cols, dst, ld_dst, rows_dst, round_tf, scale);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace eegb200

namespace eegb200 {
// dst = tf32?(dropout(src)) over [rows, ld]; optionally colsum[c] += sum_r dst[r][c] for c < cs_cols (bias gradient
// of the layer whose output fed this dropout), rows with r % row_mod == row_skip excluded.  Block = 64 float4
// columns x 4 row lanes (ld <= 256) so that each thread stays on one column quad.
__global__ void __launch_bounds__(256) dropout_apply_colsum_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                   int rows, int ld, DropoutCfg cfg, int round_tf,
                                                                   float* __restrict__ colsum, int cs_cols, int row_mod,
                                                                   int row_skip) {
  __shared__ float4 part[4][64];
  const int c4 = threadIdx.x & 63, ry = threadIdx.x >> 6;
  const int ld4 = ld >> 2;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 < ld4) {
    const int stride = gridDim.x * 4;
    for (int r0 = blockIdx.x * 4 + ry; r0 < rows; r0 += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * stride;
        if (r < rows) v[u] = reinterpret_cast<const float4*>(src)[(size_t)r * ld4 + c4];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * stride;
        if (r >= rows) continue;
        const size_t i = (size_t)r * ld4 + c4;
        float4 w = v[u];
        if (cfg.p > 0.f) {
          const uint32_t m = dropout_keep4(cfg, (uint64_t)i * 4);
          w.x = (m & 1u) ? w.x * cfg.scale : 0.f;
          w.y = (m & 2u) ? w.y * cfg.scale : 0.f;
          w.z = (m & 4u) ? w.z * cfg.scale : 0.f;
          w.w = (m & 8u) ? w.w * cfg.scale : 0.f;
        }
        if (round_tf) { w.x = tf32_rn(w.x); w.y = tf32_rn(w.y); w.z = tf32_rn(w.z); w.w = tf32_rn(w.w); }
        reinterpret_cast<float4*>(dst)[i] = w;
        if (!(row_mod > 0 && (r % row_mod) == row_skip)) { acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w; }
      }
    }
  }
  part[ry][c4] = acc;
  __syncthreads();
  if (ry == 0 && c4 < ld4) {
    float4 t = part[0][c4];
    for (int i = 1; i < 4; ++i) { t.x += part[i][c4].x; t.y += part[i][c4].y; t.z += part[i][c4].z; t.w += part[i][c4].w; }
    const float tv[4] = {t.x, t.y, t.z, t.w};
    for (int q = 0; q < 4; ++q)
      if (c4 * 4 + q < cs_cols) atomicAdd(&colsum[c4 * 4 + q], tv[q]);
  }
}
int dropout_apply_colsum(const float* src, float* dst, int rows, int ld, DropoutCfg cfg, int round_tf, float* colsum_out,
                         int cs_cols, int row_mod, int row_skip, cudaStream_t s) {
  ProfScope _ps("dropout_apply_colsum", s, 0.0, (double)rows * ld * 8.0);
  EEG_REQUIRE((ld & 3) == 0 && ld <= 256, "dropout_apply_colsum: ld %d must be a multiple of 4 and <= 256", ld);
  static int cap = -1;
  if (cap < 0) { const char* e = getenv("EEGB200_DACS_BLOCKS"); cap = e ? atoi(e) : 148 * 4; }
  int blocks = cdiv(rows, 4 * 16);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dropout_apply_colsum_kernel<<<blocks, 256, 0, s>>>(src, dst, rows, ld, cfg, round_tf, colsum_out, cs_cols, row_mod, row_skip);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

__global__ void dropout_apply_kernel(const float4* __restrict__ src, float4* __restrict__ dst, long long n4, DropoutCfg cfg,
                                     int round_tf) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = src[i];
    if (cfg.p > 0.f) {
      const uint32_t m = dropout_keep4(cfg, (uint64_t)i * 4);
      v.x = (m & 1u) ? v.x * cfg.scale : 0.f;
      v.y = (m & 2u) ? v.y * cfg.scale : 0.f;
      v.z = (m & 4u) ? v.z * cfg.scale : 0.f;
      v.w = (m & 8u) ? v.w * cfg.scale : 0.f;
    }
    if (round_tf) { v.x = tf32_rn(v.x); v.y = tf32_rn(v.y); v.z = tf32_rn(v.z); v.w = tf32_rn(v.w); }
    dst[i] = v;
  }
}
int dropout_apply(const float* src, float* dst, int rows, int ld, DropoutCfg cfg, int round_tf, cudaStream_t s) {
  ProfScope _ps("dropout_apply", s, 0.0, (double)rows * ld * 8.0);
  EEG_REQUIRE((ld & 3) == 0, "dropout_apply: ld %d not a multiple of 4", ld);
  const long long n4 = (long long)rows * ld / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  dropout_apply_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), n4, cfg,
                                              round_tf);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}
// ------------------------------------------------------------------------------------------------
// optional L2 normalisation of the embedding (north_star wording; the reference does NOT normalise the EEG embedding,
// ATMS_retrieval.py:182-191, so ATMS(normalize=False) is the default): y = x / max(|x|, eps), one warp per row
// backward: dx = (dy - y * (y . dy)) / max(|x|, eps)
// ------------------------------------------------------------------------------------------------
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ norms, int rows,
                                  int D, float eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
  float s = 0.f;
  for (int i = lane; i < D / 4; i += 32) {
    const float4 v = xr[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  const float nrm = fmaxf(sqrtf(warp_sum(s)), eps);
  const float inv = 1.f / nrm;
  if (lane == 0) norms[row] = nrm;
  float4* yr = reinterpret_cast<float4*>(y + (size_t)row * D);
  for (int i = lane; i < D / 4; i += 32) {
    const float4 v = xr[i];
    yr[i] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
  }
}
__global__ void l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ norms,
                                  const float* __restrict__ dy, float* __restrict__ dx, int rows, int D) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* yr = reinterpret_cast<const float4*>(y + (size_t)row * D);
  const float4* gr = reinterpret_cast<const float4*>(dy + (size_t)row * D);
  float s = 0.f;
  for (int i = lane; i < D / 4; i += 32) {
    const float4 a = yr[i], g = gr[i];
    s += a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w;
  }
  const float dot = warp_sum(s), inv = 1.f / norms[row];
  float4* dr = reinterpret_cast<float4*>(dx + (size_t)row * D);
  for (int i = lane; i < D / 4; i += 32) {
    const float4 a = yr[i], g = gr[i];
    dr[i] = make_float4((g.x - a.x * dot) * inv, (g.y - a.y * dot) * inv, (g.z - a.z * dot) * inv, (g.w - a.w * dot) * inv);
  }
}
int l2norm_fwd(const float* x, float* y, float* norms, int rows, int D, cudaStream_t s) {
  ProfScope _ps("l2norm_fwd", s, 0.0, (double)rows * D * 8.0);
  EEG_REQUIRE((D & 3) == 0, "l2norm: D %d must be a multiple of 4", D);
  l2norm_fwd_kernel<<<cdiv(rows * 32, 256), 256, 0, s>>>(x, y, norms, rows, D, 1e-12f);   // F.normalize eps
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}
int l2norm_bwd(const float* y, const float* norms, const float* dy, float* dx, int rows, int D, cudaStream_t s) {
  ProfScope _ps("l2norm_bwd", s, 0.0, (double)rows * D * 12.0);
  EEG_REQUIRE((D & 3) == 0, "l2norm: D %d must be a multiple of 4", D);
  l2norm_bwd_kernel<<<cdiv(rows * 32, 256), 256, 0, s>>>(y, norms, dy, dx, rows, D);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace eegb200

// Symmetric InfoNCE (ClipLoss.forward, models/loss.py:100-141) on a local row block of the logits
//   L[i, t*N + j] = s * <E_i, T^t_j>,  t in {img, txt},  i local (global index row_offset+i), j global
// loss_t = 1/2 [ mean_i (lse_j L[i,j] - L[i,i]) + mean_j (lse_i L[i,j] - L[j,j]) ]
// The row terms are local; the column log-sum-exp is reduced over all row blocks ("parts": row chunks of this rank,
// and the other ranks' partials after an all-gather) -- mathematically the reference's world_size>1,
// local_loss=False path (loss.py:113-120) at 1/W of the FLOPs.
// Gradient: G = w_t/(2N) * (softmax_row + softmax_col - 2 I),  dE = s * G * Tcat,  ds = sum(G .* L) / s.
#include "kernels.h"

namespace eegb200 {

// one warp per (row, target)
__global__ void infonce_row_lse_kernel(InfoNceArgs a, float* __restrict__ row_lse, float* __restrict__ diag) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= a.nt * a.B) return;
  const int t = w / a.B, i = w % a.B;
  const float* row = a.logits + (size_t)i * a.ld + (size_t)t * a.N;
  float mx = -INFINITY;
  for (int j = lane; j < a.N; j += 32) mx = fmaxf(mx, row[j]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < a.N; j += 32) s += __expf(row[j] - mx);
  s = warp_sum(s);
  if (lane == 0) {
    row_lse[w] = mx + logf(s);
    diag[w] = row[a.row_offset + i];
  }
}
// the same, one pass (online max / sum) over float4 loads: N % 4 == 0 and 16-byte aligned rows.  At the data-parallel
// sizes (N = 8192 per target) the logits block is 67 MB and this kernel is bandwidth-bound
__global__ void infonce_row_lse_vec_kernel(InfoNceArgs a, float* __restrict__ row_lse, float* __restrict__ diag) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= a.nt * a.B) return;
  const int t = w / a.B, i = w % a.B;
  const float* row = a.logits + (size_t)i * a.ld + (size_t)t * a.N;
  const float4* row4 = reinterpret_cast<const float4*>(row);
  float mx = -INFINITY, s = 0.f;
  for (int j = lane; j < a.N / 4; j += 32) {
    const float4 v = row4[j];
    const float m4 = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    if (m4 > mx) { s *= __expf(mx - m4); mx = m4; }
    s += __expf(v.x - mx) + __expf(v.y - mx) + __expf(v.z - mx) + __expf(v.w - mx);
  }
  const float M = warp_max(mx);
  s = warp_sum(mx > -INFINITY ? s * __expf(mx - M) : 0.f);
  if (lane == 0) {
    row_lse[w] = M + logf(s);
    diag[w] = row[a.row_offset + i];
  }
}
static bool info_vec_ok(const InfoNceArgs& a) {
  return (a.N & 3) == 0 && (a.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(a.logits) & 15) == 0;
}
int infonce_row_lse(const InfoNceArgs& a, float* row_lse, float* diag, cudaStream_t s) {
  ProfScope _ps("infonce_row_lse", s, 0.0, (double)a.B * a.nt * a.N * 4.0);
  const int warps = a.nt * a.B;
  if (info_vec_ok(a)) infonce_row_lse_vec_kernel<<<cdiv(warps * 32, 128), 128, 0, s>>>(a, row_lse, diag);
  else infonce_row_lse_kernel<<<cdiv(warps * 32, 256), 256, 0, s>>>(a, row_lse, diag);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// column partials over row chunks: block (32 columns x 8 row lanes), grid (2N/32, chunks)
__global__ void infonce_col_partial_kernel(InfoNceArgs a, int rows_per_chunk, float* __restrict__ part_max,
                                           float* __restrict__ part_sum) {
  __shared__ float sm[8][33], ss[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int ncol = a.nt * a.N;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(a.B, r0 + rows_per_chunk);
  float mx = -INFINITY, s = 0.f;
  if (c < ncol) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      const float v = a.logits[(size_t)r * a.ld + c];
      if (v > mx) { s = s * __expf(mx - v) + 1.f; mx = v; }
      else s += __expf(v - mx);
    }
  }
  sm[threadIdx.y][threadIdx.x] = mx;
  ss[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < ncol) {
    float M = sm[0][threadIdx.x];
    for (int q = 1; q < 8; ++q) M = fmaxf(M, sm[q][threadIdx.x]);
    float S = 0.f;
    for (int q = 0; q < 8; ++q) {
      const float m = sm[q][threadIdx.x];
      if (m > -INFINITY) S += ss[q][threadIdx.x] * __expf(m - M);
    }
    part_max[(size_t)blockIdx.y * ncol + c] = M;
    part_sum[(size_t)blockIdx.y * ncol + c] = S;
  }
}
// the same with four columns per thread (float4 rows of 512 contiguous bytes per warp): block = 128 columns x 8 row lanes
__global__ void infonce_col_partial_vec_kernel(InfoNceArgs a, int rows_per_chunk, float* __restrict__ part_max,
                                               float* __restrict__ part_sum) {
  __shared__ float4 sm[8][33], ss[8][33];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int ncol = a.nt * a.N;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(a.B, r0 + rows_per_chunk);
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, s[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < ncol) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      const float4 v4 = *reinterpret_cast<const float4*>(a.logits + (size_t)r * a.ld + c);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float nm = fmaxf(mx[q], v[q]);
        s[q] = s[q] * __expf(mx[q] - nm) + __expf(v[q] - nm);      // exp(-inf - finite) = 0 on the first row
        mx[q] = nm;
      }
    }
  }
  sm[threadIdx.y][threadIdx.x] = make_float4(mx[0], mx[1], mx[2], mx[3]);
  ss[threadIdx.y][threadIdx.x] = make_float4(s[0], s[1], s[2], s[3]);
  __syncthreads();
  if (threadIdx.y < 4 && c < ncol) {                       // row lane q merges column c + q
    const int q = threadIdx.y;
    float M = -INFINITY;
    for (int y = 0; y < 8; ++y) M = fmaxf(M, reinterpret_cast<const float*>(&sm[y][threadIdx.x])[q]);
    float S = 0.f;
    for (int y = 0; y < 8; ++y) {
      const float m = reinterpret_cast<const float*>(&sm[y][threadIdx.x])[q];
      if (m > -INFINITY) S += reinterpret_cast<const float*>(&ss[y][threadIdx.x])[q] * __expf(m - M);
    }
    part_max[(size_t)blockIdx.y * ncol + c + q] = M;
    part_sum[(size_t)blockIdx.y * ncol + c + q] = S;
  }
}
int infonce_col_chunks(int B) {
  int chunks = B / 64;
  if (chunks < 1) chunks = 1;
  if (chunks > 16) chunks = 16;
  return chunks;
}
int infonce_col_partial(const InfoNceArgs& a, float* part_max, float* part_sum, cudaStream_t s) {
  ProfScope _ps("infonce_col_partial", s, 0.0, (double)a.B * a.nt * a.N * 4.0);
  const int chunks = infonce_col_chunks(a.B);
  const int rpc = cdiv(a.B, chunks);
  if (info_vec_ok(a)) {
    dim3 grid(cdiv(a.nt * a.N, 128), chunks);
    infonce_col_partial_vec_kernel<<<grid, dim3(32, 8), 0, s>>>(a, rpc, part_max, part_sum);
  } else {
    dim3 grid(cdiv(a.nt * a.N, 32), chunks);
    infonce_col_partial_kernel<<<grid, dim3(32, 8), 0, s>>>(a, rpc, part_max, part_sum);
  }
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// merge n_parts (max,sum) partials per column; write merged (max,sum) and/or lse
__global__ void infonce_col_reduce_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                                          int n_parts, size_t stride, int n, float* __restrict__ out_max,
                                          float* __restrict__ out_sum, float* __restrict__ out_lse) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float M = -INFINITY;
  for (int p = 0; p < n_parts; ++p) M = fmaxf(M, part_max[(size_t)p * stride + c]);
  float S = 0.f;
  for (int p = 0; p < n_parts; ++p) {
    const float m = part_max[(size_t)p * stride + c];
    if (m > -INFINITY) S += part_sum[(size_t)p * stride + c] * __expf(m - M);
  }
  if (out_max) out_max[c] = M;
  if (out_sum) out_sum[c] = S;
  if (out_lse) out_lse[c] = M + logf(S);
}
int infonce_col_reduce(const float* part_max, const float* part_sum, int n_parts, size_t stride, int n, float* out_max,
                       float* out_sum, float* out_lse, cudaStream_t s) {
  ProfScope _ps("infonce_col_reduce", s);
  infonce_col_reduce_kernel<<<cdiv(n, 256), 256, 0, s>>>(part_max, part_sum, n_parts, stride, n, out_max, out_sum, out_lse);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// this rank's share of the global loss (sum over ranks == loss):
//   sum_t w_t/(2N) * sum_i [ (row_lse - diag) + (col_lse[global i] - diag) ]
// out[0] = mix, out[1] = img loss share, out[2] = txt loss share
__global__ void infonce_loss_kernel(InfoNceArgs a, const float* __restrict__ row_lse, const float* __restrict__ diag,
                                    const float* __restrict__ col_lse, float w_img, float w_txt, float* __restrict__ out) {
  __shared__ float red[2][32];
  float acc[2] = {0.f, 0.f};
  for (int t = 0; t < a.nt; ++t)
    for (int i = threadIdx.x; i < a.B; i += blockDim.x) {
      const float d = diag[t * a.B + i];
      acc[t] += (row_lse[t * a.B + i] - d) + (col_lse[(size_t)t * a.N + a.row_offset + i] - d);
    }
  for (int t = 0; t < 2; ++t) {
    const float v = warp_sum(acc[t]);
    if ((threadIdx.x & 31) == 0) red[t][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s0 = 0.f, s1 = 0.f;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { s0 += red[0][q]; s1 += red[1][q]; }
    const float inv = 1.f / (2.f * a.N);
    out[1] = s0 * inv;
    out[2] = s1 * inv;
    out[0] = w_img * out[1] + w_txt * out[2];
  }
}
int infonce_loss(const InfoNceArgs& a, const float* row_lse, const float* diag, const float* col_lse, float w_img,
                 float w_txt, float* loss_out, cudaStream_t s) {
  ProfScope _ps("infonce_loss", s);
  infonce_loss_kernel<<<1, 1024, 0, s>>>(a, row_lse, diag, col_lse, w_img, w_txt, loss_out);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

__global__ void infonce_grad_kernel(InfoNceArgs a, float* __restrict__ L, const float* __restrict__ row_lse,
                                    const float* __restrict__ col_lse, float w_img, float w_txt,
                                    const float* __restrict__ scale_dev, float* __restrict__ dscale, float gout, int rt) {
  __shared__ float red[32];
  const int ncol = a.nt * a.N;
  const long long total = (long long)a.B * ncol;
  float ds = 0.f;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / ncol), c = (int)(idx % ncol);
    const int t = c >= a.N ? 1 : 0;
    const int j = c - t * a.N;
    const float l = L[(size_t)i * a.ld + c];
    const float w = (t ? w_txt : w_img) * gout / (2.f * a.N);
    float g = __expf(l - row_lse[t * a.B + i]) + __expf(l - col_lse[c]);
    if (j == a.row_offset + i) g -= 2.f;
    g *= w;
    ds = fmaf(g, l, ds);
    L[(size_t)i * a.ld + c] = tf32_if(g, rt);
  }
  ds = warp_sum(ds);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ds;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) s += red[q];
    // d loss / d scale = sum(G .* L) / scale (L = scale * <E,T>); at scale == 0 every logit is 0 and the quotient is 0/0:
    // report 0 there instead of NaN (the raw dot products are not kept)
    const float sc = __ldg(scale_dev);
    atomicAdd(dscale, sc != 0.f ? s / sc : 0.f);
  }
}
// the same on a 2-D grid: blockIdx.x = 1024-column slab (float4 per thread), blockIdx.y strides over the rows -- no
// per-element index division, 16-byte accesses
__global__ void infonce_grad_vec_kernel(InfoNceArgs a, float* __restrict__ L, const float* __restrict__ row_lse,
                                        const float* __restrict__ col_lse, float w_img, float w_txt,
                                        const float* __restrict__ scale_dev, float* __restrict__ dscale, float gout, int rt) {
  __shared__ float red[8];
  const int ncol = a.nt * a.N;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  float ds = 0.f;
  if (c < ncol) {
    const int t = c >= a.N ? 1 : 0;                       // N % 4 == 0: a float4 never straddles the two targets
    const int j0 = c - t * a.N;
    const float w = (t ? w_txt : w_img) * gout / (2.f * a.N);
    const float4 cl = *reinterpret_cast<const float4*>(col_lse + c);
    for (int i = blockIdx.y; i < a.B; i += gridDim.y) {
      float4* p = reinterpret_cast<float4*>(L + (size_t)i * a.ld + c);
      const float4 l = *p;
      const float rl = row_lse[t * a.B + i];
      float g[4] = {__expf(l.x - rl) + __expf(l.x - cl.x), __expf(l.y - rl) + __expf(l.y - cl.y),
                    __expf(l.z - rl) + __expf(l.z - cl.z), __expf(l.w - rl) + __expf(l.w - cl.w)};
      const int dj = a.row_offset + i - j0;                // the diagonal element of this row, if inside the float4
      if (dj >= 0 && dj < 4) g[dj] -= 2.f;
      g[0] *= w; g[1] *= w; g[2] *= w; g[3] *= w;
      ds = fmaf(g[0], l.x, fmaf(g[1], l.y, fmaf(g[2], l.z, fmaf(g[3], l.w, ds))));
      *p = make_float4(tf32_if(g[0], rt), tf32_if(g[1], rt), tf32_if(g[2], rt), tf32_if(g[3], rt));
    }
  }
  ds = warp_sum(ds);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ds;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) s += red[q];
    // d loss / d scale = sum(G .* L) / scale (L = scale * <E,T>); at scale == 0 every logit is 0 and the quotient is 0/0:
    // report 0 there instead of NaN (the raw dot products are not kept)
    const float sc = __ldg(scale_dev);
    atomicAdd(dscale, sc != 0.f ? s / sc : 0.f);
  }
}
int infonce_grad(const InfoNceArgs& a, float* logits_inout, const float* row_lse, const float* col_lse, float w_img,
                 float w_txt, const float* logit_scale_dev, float* dscale, float grad_out_scale, cudaStream_t s) {
  ProfScope _ps("infonce_grad", s, 0.0, (double)a.B * a.nt * a.N * 8.0);
  if (info_vec_ok(a) && (reinterpret_cast<uintptr_t>(col_lse) & 15) == 0) {
    const int slabs = cdiv(a.nt * a.N, 1024);
    int rows = 148 * 8 / slabs;                            // ~8 blocks of 256 threads per SM
    if (rows < 1) rows = 1;
    if (rows > a.B) rows = a.B;
    infonce_grad_vec_kernel<<<dim3(slabs, rows), 256, 0, s>>>(a, logits_inout, row_lse, col_lse, w_img, w_txt, logit_scale_dev,
                                                              dscale, grad_out_scale, tf32_rounding());
    EEG_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
  }
  const long long total = (long long)a.B * a.nt * a.N;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  infonce_grad_kernel<<<blocks, 256, 0, s>>>(a, logits_inout, row_lse, col_lse, w_img, w_txt, logit_scale_dev, dscale,
                                             grad_out_scale, tf32_rounding());
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// argmax per row (first maximal index) compared against labels; used for the train-accuracy bookkeeping
// (ATMS_retrieval.py:241-250) and top-1 retrieval
__global__ void argmax_count_kernel(const float* __restrict__ logits, int ld, int rows, int cols,
                                    const long long* __restrict__ labels, int* __restrict__ correct,
                                    long long* __restrict__ pred_out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows) return;
  const float* row = logits + (size_t)w * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = lane; j < cols; j += 32) {
    const float v = row[j];
    if (v > best) { best = v; bi = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) {
    if (pred_out) pred_out[w] = bi;
    if (labels && correct && labels[w] == bi) atomicAdd(correct, 1);
  }
}
int argmax_count(const float* logits, int ld, int rows, int cols, const long long* labels, int* correct,
                 long long* pred_out, cudaStream_t s) {
  ProfScope _ps("argmax_count", s, 0.0, (double)rows * cols * 4.0);
  argmax_count_kernel<<<cdiv(rows * 32, 256), 256, 0, s>>>(logits, ld, rows, cols, labels, correct, pred_out);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// top-5 indices per row, descending score (ties -> smaller index first)
__global__ void topk5_kernel(const float* __restrict__ logits, int ld, int rows, int cols, int* __restrict__ out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows) return;
  const float* row = logits + (size_t)w * ld;
  int sel[5] = {-1, -1, -1, -1, -1};
  for (int p = 0; p < 5; ++p) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < cols; j += 32) {
      bool used = false;
#pragma unroll
      for (int q = 0; q < 5; ++q) used |= (sel[q] == j);
      const float v = row[j];
      if (!used && (v > best || (v == best && j < bi))) { best = v; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    sel[p] = (bi == 0x7fffffff) ? -1 : bi;
    if (lane == 0) out[w * 5 + p] = sel[p];
  }
}
int topk5(const float* logits, int ld, int rows, int cols, int* top5_out, cudaStream_t s) {
  ProfScope _ps("topk5", s, 0.0, (double)rows * cols * 20.0);
  topk5_kernel<<<cdiv(rows * 32, 256), 256, 0, s>>>(logits, ld, rows, cols, top5_out);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// out[q][c] = logits[q][sel[q][c]]  (per-query candidate lists of evaluate_model, ATMS_retrieval.py:298-306)
__global__ void gather_cols_kernel(const float* __restrict__ logits, int ld, const int* __restrict__ sel, int Q, int k,
                                   float* __restrict__ out) {
  const long long total = (long long)Q * k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i / k);
    out[i] = logits[(size_t)q * ld + sel[i]];
  }
}
int gather_cols(const float* logits, int ld, const int* sel, int Q, int k, float* out, cudaStream_t s) {
  ProfScope _ps("gather_cols", s);
  const long long total = (long long)Q * k;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  gather_cols_kernel<<<blocks, 256, 0, s>>>(logits, ld, sel, Q, k, out);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// x = hi + lo with both parts exactly representable in TF32 (3xTF32 split: a.b ~ hi.hi + hi.lo + lo.hi)
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n,
                                  int rt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float h = tf32_if(v, rt);
    hi[i] = h;
    lo[i] = tf32_if(v - h, rt);
  }
}
// 3xTF32 as ONE GEMM with K tripled:  A' = [hi | hi | lo],  B' = [hi | lo | hi]   (rows of length 3*D)
__global__ void split3_tf32_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows, int D, int is_b,
                                   int rt) {
  const long long n = rows * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / D;
    const int c = (int)(i % D);
    const float v = x[i];
    const float h = tf32_if(v, rt);
    const float l = tf32_if(v - h, rt);
    float* o = out + r * 3 * D + c;
    o[0] = h;
    o[D] = is_b ? l : h;
    o[2 * D] = is_b ? h : l;
  }
}
int split3_tf32(const float* x, float* out, long long rows, int D, int is_b, cudaStream_t s) {
  ProfScope _ps("split_tf32", s, 0.0, (double)rows * D * 16.0);
  const long long n = rows * D;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  split3_tf32_kernel<<<blocks, 256, 0, s>>>(x, out, rows, D, is_b, tf32_rounding());
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}
int split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t s) {
  ProfScope _ps("split_tf32", s, 0.0, (double)n * 12.0);
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  split_tf32_kernel<<<blocks, 256, 0, s>>>(x, hi, lo, n, tf32_rounding());
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Regression term of the reconstruction-training variant: nn.MSELoss()(eeg_features, img_features)
// (Generation/ATMS_reconstruction.py:201, 227-228): mean over ALL n_total*D elements of the global batch.
// This rank's rows add   loss += weight * sum (e-t)^2 / (n_total*D)   and   d_eeg += g_scale * (e-t),
// g_scale = weight * grad_out * 2 / (n_total*D).  float4 grid-stride; one double atomic per block.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mse_kernel(const float4* __restrict__ e, const float4* __restrict__ t, long long n4,
                                                  float loss_scale, float g_scale, float* __restrict__ loss,
                                                  float* __restrict__ loss_term, float4* d_eeg) {
  __shared__ float part[8];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = e[i], b = t[i];
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
    acc = fmaf(dx, dx, acc); acc = fmaf(dy, dy, acc); acc = fmaf(dz, dz, acc); acc = fmaf(dw, dw, acc);
    if (d_eeg != nullptr) {
      float4 g = d_eeg[i];
      g.x = fmaf(g_scale, dx, g.x); g.y = fmaf(g_scale, dy, g.y); g.z = fmaf(g_scale, dz, g.z); g.w = fmaf(g_scale, dw, g.w);
      d_eeg[i] = g;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += part[i];
    tot *= loss_scale;
    if (loss != nullptr) atomicAdd(loss, tot);
    if (loss_term != nullptr) atomicAdd(loss_term, tot);
  }
}
int mse_loss(const float* eeg, const float* tgt, int B, int D, long long n_total, float weight, float grad_out, float* loss,
             float* loss_term, float* d_eeg, cudaStream_t s) {
  ProfScope _ps("mse_loss", s, 0.0, (double)B * D * (d_eeg ? 16.0 : 8.0));
  const long long n4 = (long long)B * D / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  const double inv = 1.0 / ((double)n_total * (double)D);
  mse_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(eeg), reinterpret_cast<const float4*>(tgt), n4,
                                    (float)(weight * inv), (float)(2.0 * weight * grad_out * inv), loss, loss_term,
                                    reinterpret_cast<float4*>(d_eeg));
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace eegb200

// Fused AdamW over one flat fp32 parameter arena (torch.optim.AdamW single-tensor semantics:
// decoupled weight decay, bias-corrected moments; constructed at Retrieval/ATMS_retrieval.py:548).
#include "kernels.h"

namespace eegb200 {

__global__ void adamw_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                             float4* __restrict__ v, long long n4, float decay, float b1, float b2, float eps,
                             float step_size, float inv_sqrt_bc2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 P = p[i], M = m[i], V = v[i];
    const float4 G = g[i];
    float* pp = &P.x; float* mm = &M.x; float* vv = &V.x; const float* gg = &G.x;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      pp[q] *= decay;
      mm[q] = mm[q] + (gg[q] - mm[q]) * (1.f - b1);            // lerp, like torch
      vv[q] = vv[q] * b2 + (1.f - b2) * gg[q] * gg[q];
      const float denom = sqrtf(vv[q]) * inv_sqrt_bc2 + eps;
      pp[q] -= step_size * (mm[q] / denom);
    }
    p[i] = P; m[i] = M; v[i] = V;
  }
}
__global__ void adamw_tail_kernel(float* p, const float* g, float* m, float* v, long long begin, long long n, float decay,
                                  float b1, float b2, float eps, float step_size, float inv_sqrt_bc2) {
  const long long i = begin + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float P = p[i] * decay;
  const float G = g[i];
  const float M = m[i] + (G - m[i]) * (1.f - b1);
  const float V = v[i] * b2 + (1.f - b2) * G * G;
  P -= step_size * (M / (sqrtf(V) * inv_sqrt_bc2 + eps));
  p[i] = P; m[i] = M; v[i] = V;
}

// bias corrections computed on the device from a step counter in global memory
__global__ void adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                                 const long long* __restrict__ step_dev) {
  __shared__ float s_step, s_isq;
  if (threadIdx.x == 0) {
    const double st = (double)step_dev[0];
    s_step = (float)((double)lr / (1.0 - pow((double)b1, st)));
    s_isq = (float)(1.0 / sqrt(1.0 - pow((double)b2, st)));
  }
  __syncthreads();
  const float step_size = s_step, inv_sqrt_bc2 = s_isq, decay = 1.f - lr * wd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float P = p[i] * decay;
    const float G = g[i];
    const float M = m[i] + (G - m[i]) * (1.f - b1);
    const float V = v[i] * b2 + (1.f - b2) * G * G;
    P -= step_size * (M / (sqrtf(V) * inv_sqrt_bc2 + eps));
    p[i] = P; m[i] = M; v[i] = V;
  }
}
int adamw_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                   float wd, const long long* step_dev, cudaStream_t s) {
  ProfScope _ps("adamw", s, 0.0, (double)n * 28.0);
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  adamw_dev_kernel<<<blocks, 256, 0, s>>>(p, g, m, v, n, lr, b1, b2, eps, wd, step_dev);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
               float wd, int step, cudaStream_t s) {
  ProfScope _ps("adamw", s, 0.0, (double)n * 28.0);
  EEG_REQUIRE(step >= 1, "adamw: step must start at 1");
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const float decay = 1.f - lr * wd;
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  long long n4 = aligned ? n / 4 : 0;
  if (n4 > 0) {
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    adamw_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g),
                                        reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n4, decay, b1, b2, eps,
                                        step_size, inv_sqrt_bc2);
    count_launch();
  }
  const long long done = n4 * 4;
  if (done < n) {
    adamw_tail_kernel<<<(int)((n - done + 255) / 256), 256, 0, s>>>(p, g, m, v, done, n, decay, b1, b2, eps, step_size,
                                                                    inv_sqrt_bc2);
    count_launch();
  }
  EEG_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace eegb200

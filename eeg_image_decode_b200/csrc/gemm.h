// Internal GEMM interface: D[M,N] = epilogue( A[M,K] * B[N,K]^T ).
// Two implementations share it: the tcgen05/TMA TF32 kernel (product path) and an exact-fp32 SIMT
// kernel used to verify it on device (tests only, selected with eegb200_set_gemm_backend).
#pragma once
#include "common.cuh"

namespace eegb200 {

struct GemmOperand {
  const float* ptr;
  int ld;         // leading dimension in floats (multiple of 4, TMA needs 16-byte strides)
  int mn_major;   // 0: K-major, element (i,k) at ptr[i*ld + k];  1: MN-major, element (i,k) at ptr[k*ld + i]
};

enum { EPI_ACT_NONE = 0, EPI_ACT_GELU = 1 };
enum { EPI_STORE = 0, EPI_ADD = 1, EPI_ATOMIC = 2 };

// order: v = alpha*acc; +bias; [+resid if resid_before_drop]; aux_out=v; act; dropout; *gelu'(mul_in); +resid; tf32 round; store
struct Epilogue {
  float* C = nullptr;
  int ldc = 0;
  float alpha = 1.f;
  const float* alpha_dev = nullptr;   // optional device scalar multiplied into alpha (e.g. the logit_scale parameter)
  const float* bias = nullptr;   // per column, or table [bias_period][ld_bias] indexed by (row % bias_period)
  int bias_period = 0;           // 0 -> per-column vector
  int ld_bias = 0;
  float* aux_out = nullptr;      // pre-activation copy
  int ld_aux = 0;
  int act = EPI_ACT_NONE;
  DropoutCfg drop = {nullptr, 0, 0, 0.f, 1.f, 0};
  int drop_ld = 0;               // mask element index = row*drop_ld + col
  const float* mul_in = nullptr; // v *= gelu'(mul_in[row,col])
  int ld_mul = 0;
  const float* resid = nullptr;
  int ld_res = 0;
  int resid_before_drop = 0;     // add the residual right after the bias (before activation / dropout) instead of last
  int round_tf32 = 0;
  int store_mode = EPI_STORE;
  // fused BatchNorm(train)+ELU backward of the conv stack (tcgen05 vector path only):
  //   v *= ELU'(gamma[c]*yhat + beta[c]),  yhat = (bn_y[row,col] - mean[c]) * rstd[c],  c = col % 40
  //   bn_sums[c] += v,  bn_sums[40+c] += v*yhat      (the two reductions the BatchNorm backward needs)
  const float* bn_y = nullptr;
  int ld_bn_y = 0;
  const float* bn_mean_rstd = nullptr;   // [2][40]
  const float* bn_gamma = nullptr;
  const float* bn_beta = nullptr;
  double* bn_sums = nullptr;             // [2][40]
};

struct GemmArgs {
  int M = 0, N = 0, K = 0;
  GemmOperand A{nullptr, 0, 0}, B{nullptr, 0, 0};
  Epilogue epi;
  int split_k = 1;   // >1 requires epi.store_mode == EPI_ATOMIC and a pre-zeroed C
  int tile_n = 0;    // 0: heuristic; 64 / 128 / 256: force the N tile of the tcgen05 kernel
};

enum { GEMM_BACKEND_TCGEN05 = 0, GEMM_BACKEND_SIMT_FP32 = 1 };
// 1 when GEMM operands should be pre-rounded to TF32 (tensor-core backend), 0 for the exact-fp32 verification backend
int tf32_rounding();
void gemm_set_backend(int backend);
int gemm_get_backend();
int gemm_launch(const GemmArgs& g, cudaStream_t stream);           // dispatches on the backend switch
int gemm_launch_tcgen05(const GemmArgs& g, cudaStream_t stream);
int gemm_launch_simt(const GemmArgs& g, cudaStream_t stream);
int gemm_make_tmap(CUtensorMap* tm, const GemmOperand& op, int rows, int K, int box_rows, int* is_3d);
long long gemm_launch_count();                                      // kernels launched so far (bench bookkeeping)
void count_launch(int n = 1);
long long total_launch_count();

// ---- optional per-launch event profiler (bench.py's roofline leg; off by default) ----
// ProfScope brackets the launches issued in its lifetime with CUDA events on `stream` when profiling is on.
void prof_enable(int on);
int prof_enabled();
int prof_report(char* buf, size_t cap);   // JSON {"name": {"ms":..,"n":..,"flops":..,"bytes":..}, ...}; clears records
struct ProfScope {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t s;
  ProfScope(const char* name, cudaStream_t stream, double flops = 0.0, double bytes = 0.0);
  ~ProfScope();
  const char* name_;
  double flops_, bytes_;
};

// ---- epilogue math shared by both kernels ----
__device__ __forceinline__ float epi_value(const Epilogue& e, int row, int col, float acc) {
  float v = e.alpha * acc;
  if (e.alpha_dev) v *= __ldg(e.alpha_dev);
  if (e.bias) v += e.bias_period ? e.bias[(size_t)(row % e.bias_period) * e.ld_bias + col] : e.bias[col];
  if (e.resid && e.resid_before_drop) v += e.resid[(size_t)row * e.ld_res + col];
  if (e.aux_out) e.aux_out[(size_t)row * e.ld_aux + col] = v;
  if (e.act == EPI_ACT_GELU) v = gelu_exact(v);
  if (e.drop.p > 0.f) v = dropout_keep(e.drop, (uint64_t)row * e.drop_ld + col) ? v * e.drop.scale : 0.f;
  if (e.mul_in) v *= gelu_grad(e.mul_in[(size_t)row * e.ld_mul + col]);
  if (e.resid && !e.resid_before_drop) v += e.resid[(size_t)row * e.ld_res + col];
  if (e.round_tf32) v = tf32_rn(v);
  return v;
}
__device__ __forceinline__ void epi_store(const Epilogue& e, int row, int col, float v) {
  float* p = e.C + (size_t)row * e.ldc + col;
  if (e.store_mode == EPI_STORE) *p = v;
  else if (e.store_mode == EPI_ADD) *p += v;
  else atomicAdd(p, v);
}

// 4 consecutive columns of one row (col % 4 == 0, col + 4 <= N, every used pointer 16-byte aligned with ld % 4 == 0).
// Split in two so the caller can issue the global loads of several rows before consuming any (latency hiding).
// Compile-time epilogue feature masks.  EF_GENERIC keeps every feature behind a runtime check (one kernel serves any
// Epilogue); the hot token GEMMs dispatch to kernels specialised on exactly the features they use, which removes the
// dead paths from the unrolled epilogue (the generic kernel is ~60 KB of SASS and stalls on instruction fetch).
enum : uint32_t {
  EF_BIAS = 1, EF_BIAS_TABLE = 2, EF_AUX = 4, EF_GELU = 8, EF_DROP = 16, EF_MUL = 32, EF_RESID = 64, EF_ROUND = 128,
  EF_BNF = 256, EF_RESID_FIRST = 512 /* generic kernel only */, EF_GENERIC = 0x80000000u
};
#define EPI_ON(bit, runtime_cond) ((F & EF_GENERIC) ? (runtime_cond) : ((F & (bit)) != 0))

struct EpiLoads { float4 bias, mul, resid; };
template <uint32_t F>
__device__ __forceinline__ EpiLoads epi_load4(const Epilogue& e, int row, int col) {
  EpiLoads L;
  L.bias = L.mul = L.resid = make_float4(0.f, 0.f, 0.f, 0.f);
  if (F & EF_GENERIC) {
    if (e.bias) L.bias = __ldg(reinterpret_cast<const float4*>(e.bias + (e.bias_period ? (size_t)(row % e.bias_period) * e.ld_bias : 0) + col));
  } else if (F & EF_BIAS_TABLE) {
    L.bias = __ldg(reinterpret_cast<const float4*>(e.bias + (size_t)(row & 63) * e.ld_bias + col));   // period 64 (tokens)
  } else if (F & EF_BIAS) {
    L.bias = __ldg(reinterpret_cast<const float4*>(e.bias + col));
  }
  if (EPI_ON(EF_MUL, e.mul_in != nullptr)) L.mul = *reinterpret_cast<const float4*>(e.mul_in + (size_t)row * e.ld_mul + col);
  else if (EPI_ON(EF_BNF, false)) L.mul = *reinterpret_cast<const float4*>(e.bn_y + (size_t)row * e.ld_bn_y + col);
  if (EPI_ON(EF_RESID, e.resid != nullptr)) L.resid = *reinterpret_cast<const float4*>(e.resid + (size_t)row * e.ld_res + col);
  return L;
}
// (an out-of-line version of this function was tried to shrink the kernels: the call overhead and the shared-memory
// copy of the parameters cost more than the instruction-cache misses they saved: 95 -> 125 us per GEMM)
template <uint32_t F>
__device__ __forceinline__ void epi_finish4(const Epilogue& e, int row, int col, float4 acc, const EpiLoads& L, float a,
                                            float* bn_s1 = nullptr, float* bn_s2 = nullptr, const float* bn_tab = nullptr,
                                            int bn_c = 0) {
  float v[4] = {acc.x * a, acc.y * a, acc.z * a, acc.w * a};
  if (F & EF_BNF) {
    // bn_tab (shared memory, built once per CTA): [0][c] = mean, [1][c] = rstd, [2][c] = gamma, [3][c] = beta
    const float yv[4] = {L.mul.x, L.mul.y, L.mul.z, L.mul.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float yh = (yv[i] - bn_tab[bn_c + i]) * bn_tab[40 + bn_c + i];
      const float z = fmaf(yh, bn_tab[80 + bn_c + i], bn_tab[120 + bn_c + i]);
      v[i] *= elu1_grad(z);
      bn_s1[i] += v[i];
      bn_s2[i] = fmaf(v[i], yh, bn_s2[i]);
    }
  }
  if (EPI_ON(EF_BIAS | EF_BIAS_TABLE, e.bias != nullptr)) { v[0] += L.bias.x; v[1] += L.bias.y; v[2] += L.bias.z; v[3] += L.bias.w; }
  const bool resid_first = (F & EF_GENERIC) ? (e.resid != nullptr && e.resid_before_drop != 0) : false;
  if (resid_first) { v[0] += L.resid.x; v[1] += L.resid.y; v[2] += L.resid.z; v[3] += L.resid.w; }
  if (EPI_ON(EF_AUX, e.aux_out != nullptr))
    *reinterpret_cast<float4*>(e.aux_out + (size_t)row * e.ld_aux + col) = make_float4(v[0], v[1], v[2], v[3]);
  if (EPI_ON(EF_GELU, e.act == EPI_ACT_GELU)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = gelu_exact(v[i]);
  }
  if (EPI_ON(EF_DROP, e.drop.p > 0.f)) {
    const uint32_t m = dropout_keep4(e.drop, (uint64_t)row * e.drop_ld + col);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (m >> i) & 1u ? v[i] * e.drop.scale : 0.f;
  }
  if (EPI_ON(EF_MUL, e.mul_in != nullptr)) {
    v[0] *= gelu_grad(L.mul.x); v[1] *= gelu_grad(L.mul.y); v[2] *= gelu_grad(L.mul.z); v[3] *= gelu_grad(L.mul.w);
  }
  if (EPI_ON(EF_RESID, e.resid != nullptr) && !resid_first) { v[0] += L.resid.x; v[1] += L.resid.y; v[2] += L.resid.z; v[3] += L.resid.w; }
  if (EPI_ON(EF_ROUND, e.round_tf32 != 0)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = tf32_rn(v[i]);
  }
  float* p = e.C + (size_t)row * e.ldc + col;
  if (!(F & EF_GENERIC) || e.store_mode == EPI_STORE) {          // specialised flavours always store
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else if (e.store_mode == EPI_ADD) {
    float4 o = *reinterpret_cast<float4*>(p);
    o.x += v[0]; o.y += v[1]; o.z += v[2]; o.w += v[3];
    *reinterpret_cast<float4*>(p) = o;
  } else {
    red_add_v4(p, v[0], v[1], v[2], v[3]);
  }
}
// feature mask of an epilogue (host side); specialised kernels exist for some masks only
static inline uint32_t epi_feature_mask(const Epilogue& e) {
  uint32_t m = 0;
  if (e.bias) m |= e.bias_period ? EF_BIAS_TABLE : EF_BIAS;
  if (e.aux_out) m |= EF_AUX;
  if (e.act == EPI_ACT_GELU) m |= EF_GELU;
  if (e.drop.p > 0.f) m |= EF_DROP;
  if (e.mul_in) m |= EF_MUL;
  if (e.resid) m |= EF_RESID;
  if (e.resid && e.resid_before_drop) m |= EF_RESID_FIRST;     // no specialised kernel carries this: generic path
  if (e.round_tf32) m |= EF_ROUND;
  if (e.bn_y) m |= EF_BNF;
  return m;
}
__device__ __forceinline__ void epi_scalar4(const Epilogue& e, int row, int col, float4 a4, int N) {
  const float av[4] = {a4.x, a4.y, a4.z, a4.w};
  for (int i = 0; i < 4 && col + i < N; ++i) epi_store(e, row, col + i, epi_value(e, row, col + i, av[i]));
}
// host-side: can the vector path be used for this epilogue?
static inline bool epi_vec_ok(const Epilogue& e) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  bool ok = al(e.C) && (e.ldc & 3) == 0;
  if (e.bias) ok = ok && al(e.bias) && (e.bias_period == 0 || (e.ld_bias & 3) == 0);
  if (e.aux_out) ok = ok && al(e.aux_out) && (e.ld_aux & 3) == 0;
  if (e.mul_in) ok = ok && al(e.mul_in) && (e.ld_mul & 3) == 0;
  if (e.resid) ok = ok && al(e.resid) && (e.ld_res & 3) == 0;
  if (e.drop.p > 0.f) ok = ok && (e.drop_ld & 3) == 0;
  if (e.bn_y) ok = ok && al(e.bn_y) && (e.ld_bn_y & 3) == 0;
  return ok;
}

}  // namespace eegb200

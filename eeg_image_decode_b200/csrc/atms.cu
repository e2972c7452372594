// ATM-S encoder forward / backward orchestration behind the C ABI (include/eegdecode_b200.h).
// Every Linear / 1x1-conv / spatial-conv contraction runs through gemm_launch (tcgen05 TF32); the glue
// between them are the fused row-wise, attention and conv kernels.  Reference: ATMS.forward,
// Retrieval/ATMS_retrieval.py:182-191 and the modules it composes.
#include "../../include/eegdecode_b200.h"
#include "kernels.h"
#include <string.h>
#include <stdlib.h>

namespace eegb200 {

// ------------------------------------------------------------------------------------------------
// workspace carving
// ------------------------------------------------------------------------------------------------
struct Ws {
  // packed / TF32-rounded weights
  float *Wv_p, *tokbias, *Wqkv_p, *bqkv_p, *Wo_p, *bo_p, *W1_p, *b1_p, *W2_p, *b2_p, *Ws_p, *Wp1_r, *Wp2_r;
  // forward activations (kept for the backward)
  float *Xp, *H0, *QKV, *O, *R1, *X1, *U, *Hf, *R2, *X3, *st1, *st2, *stf;
  float *Y1, *A1, *Y2, *feat, *Z1, *G, *Z2, *stp;
  double *bn1_sums, *bn2_sums, *bn1_bsums, *bn2_bsums;
  float *bn1_mr, *bn2_mr;
  int* flag;
  // backward temporaries
  float *dZ2, *dZ2d, *dZ1, *dfeat, *dz2, *dY2, *dA1, *dX3, *dR2, *T1, *dU, *dX1, *dR1, *T2, *dO, *dQKV, *dH0, *T3;
  float *dWqkv_p, *dbqkv_p, *dWo_p, *dWs_p;
  // joint-subject variant: one packed value embedding / token-bias table per subject slot
  float *Wv_j, *tokbias_j;
  // experimental tcgen05 conv stack: spatial weights packed per channel
  float* Ws_tc;
};
constexpr int MAX_JOINT_SUBJECTS = 16;

struct Carver {
  uint8_t* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

// The (b,p) x (c,k) conv intermediates Y1 / A1 / dA1 (363 KB per sample each) only exist for the unfused round-1 kernels,
// the exact-fp32 verification backend and the debug stage stores; the fused tcgen05 path keeps them on chip.
static bool need_conv_intermediates() { return !(tf32_rounding() && conv_tc_enabled()) || debug_stores() != 0; }

static size_t carve(void* base, int B, Ws* w) {
  Carver c{reinterpret_cast<uint8_t*>(base)};
  const size_t conv_rows = need_conv_intermediates() ? (size_t)B * N_POOL : 0;
  const size_t M = (size_t)B * N_TOK;          // token rows
  const size_t R = (size_t)B * N_POOL;         // (b, j) rows of the conv stack
  Ws t;
  t.Wv_p = c.take<float>(256 * 256);
  t.tokbias = c.take<float>(64 * 256);
  t.Wqkv_p = c.take<float>(768 * 256);
  t.bqkv_p = c.take<float>(768);
  t.Wo_p = c.take<float>(256 * 256);
  t.bo_p = c.take<float>(256);
  t.W1_p = c.take<float>(256 * 256);
  t.b1_p = c.take<float>(256);
  t.W2_p = c.take<float>(256 * 256);
  t.b2_p = c.take<float>(256);
  t.Ws_p = c.take<float>((size_t)N_FILT * K_SPAT);
  t.Wp1_r = c.take<float>((size_t)D_OUT * D_FEAT);
  t.Wp2_r = c.take<float>((size_t)D_OUT * D_OUT);
  t.Xp = c.take<float>(M * 256);
  t.H0 = c.take<float>(M * 256);
  t.QKV = c.take<float>(M * 768);
  t.O = c.take<float>(M * 256);
  t.R1 = c.take<float>(M * 256);
  t.X1 = c.take<float>(M * 256);
  t.U = c.take<float>(M * 256);
  t.Hf = c.take<float>(M * 256);
  t.R2 = c.take<float>(M * 256);
  t.X3 = c.take<float>(M * 256);
  t.st1 = c.take<float>(M * 2);
  t.st2 = c.take<float>(M * 2);
  t.stf = c.take<float>(M * 2);
  t.Y1 = c.take<float>(conv_rows * K_SPAT);
  t.A1 = c.take<float>(conv_rows * K_SPAT);
  t.Y2 = c.take<float>(R * N_FILT);
  t.feat = c.take<float>((size_t)B * D_FEAT);
  t.Z1 = c.take<float>((size_t)B * D_OUT);
  t.G = c.take<float>((size_t)B * D_OUT);
  t.Z2 = c.take<float>((size_t)B * D_OUT);
  t.stp = c.take<float>((size_t)B * 2);
  t.bn1_sums = c.take<double>(2 * N_FILT);
  t.bn2_sums = c.take<double>(2 * N_FILT);
  t.bn1_bsums = c.take<double>(2 * N_FILT);
  t.bn2_bsums = c.take<double>(2 * N_FILT);
  t.bn1_mr = c.take<float>(2 * N_FILT);
  t.bn2_mr = c.take<float>(2 * N_FILT);
  t.flag = c.take<int>(4);
  t.dZ2 = c.take<float>((size_t)B * D_OUT);
  t.dZ2d = c.take<float>((size_t)B * D_OUT);
  t.dZ1 = c.take<float>((size_t)B * D_OUT);
  t.dfeat = c.take<float>((size_t)B * D_FEAT);
  t.dz2 = c.take<float>(R * N_FILT);
  t.dY2 = c.take<float>(R * N_FILT);
  t.dA1 = c.take<float>(conv_rows * K_SPAT);
  t.dX3 = c.take<float>(M * 256);
  t.dR2 = c.take<float>(M * 256);
  t.T1 = c.take<float>(M * 256);
  t.dU = c.take<float>(M * 256);
  t.dX1 = c.take<float>(M * 256);
  t.dR1 = c.take<float>(M * 256);
  t.T2 = c.take<float>(M * 256);
  t.dO = c.take<float>(M * 256);
  t.dQKV = c.take<float>(M * 768);
  t.dH0 = c.take<float>(M * 256);
  t.T3 = c.take<float>(M * 256);
  t.dWqkv_p = c.take<float>(768 * 256);
  t.dbqkv_p = c.take<float>(768);
  t.dWo_p = c.take<float>(256 * 256);
  t.dWs_p = c.take<float>((size_t)N_FILT * K_SPAT);
  t.Wv_j = c.take<float>((size_t)MAX_JOINT_SUBJECTS * 256 * 256);
  t.tokbias_j = c.take<float>((size_t)MAX_JOINT_SUBJECTS * 64 * 256);
  t.Ws_tc = c.take<float>(conv_tc_ws_floats());
  if (w) *w = t;
  return align_up(c.off, 256);
}

// ------------------------------------------------------------------------------------------------
// weight packing (TF32-rounded, zero-padded GEMM operands) and gradient unpacking
// ------------------------------------------------------------------------------------------------
__global__ void pack_qkv_kernel(const float* __restrict__ wq, const float* __restrict__ wk, const float* __restrict__ wv,
                                float* __restrict__ out, int rt) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 768 * 256) return;
  const int row = idx >> 8, c = idx & 255;
  const int which = row >> 8, hh = (row & 255) >> 6, e = row & 63;
  const float* w = which == 0 ? wq : (which == 1 ? wk : wv);
  out[idx] = (e < D_HEAD && c < N_T) ? tf32_if(w[(hh * D_HEAD + e) * N_T + c], rt) : 0.f;
}
__global__ void pack_wo_kernel(const float* __restrict__ wo, float* __restrict__ out, int rt) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 256 * 256) return;
  const int o = idx >> 8, c = idx & 255;
  const int hh = c >> 6, e = c & 63;
  out[idx] = (o < N_T && e < D_HEAD) ? tf32_if(wo[o * (N_HEAD * D_HEAD) + hh * D_HEAD + e], rt) : 0.f;
}
// tsconv.4.weight [k2][k1][r] -> [k2][r*40 + k1]
__global__ void pack_ws_kernel(const float* __restrict__ ws, float* __restrict__ out, int rt) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N_FILT * K_SPAT) return;
  const int k2 = idx / K_SPAT, kk = idx % K_SPAT;
  const int r = kk / N_FILT, k1 = kk % N_FILT;
  out[idx] = tf32_if(ws[(k2 * N_FILT + k1) * N_CH + r], rt);
}
__global__ void pack_small_kernel(const float* __restrict__ bv, const float* __restrict__ pe, const float* __restrict__ bq,
                                  const float* __restrict__ bk, const float* __restrict__ bvv,
                                  const float* __restrict__ bo, const float* __restrict__ b1,
                                  const float* __restrict__ b2, float* __restrict__ tokbias,
                                  float* __restrict__ bqkv_p, float* __restrict__ bo_p, float* __restrict__ b1_p,
                                  float* __restrict__ b2_p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < 64 * 256) {
    const int t = idx >> 8, c = idx & 255;
    // value-embedding bias + positional table row of the channel token (Embed.py:146-149); token 0 is the subject slot
    tokbias[idx] = (t > 0 && c < N_T) ? bv[c] + pe[(t - 1) * N_T + c] : 0.f;
  }
  if (idx < 768) {
    const int which = idx >> 8, hh = (idx & 255) >> 6, e = idx & 63;
    const float* b = which == 0 ? bq : (which == 1 ? bk : bvv);
    bqkv_p[idx] = e < D_HEAD ? b[hh * D_HEAD + e] : 0.f;
  }
  if (idx < 256) {
    bo_p[idx] = idx < N_T ? bo[idx] : 0.f;
    b1_p[idx] = b1[idx];
    b2_p[idx] = idx < N_T ? b2[idx] : 0.f;
  }
}
__global__ void unpack_qkv_grad_kernel(const float* __restrict__ dw_p, const float* __restrict__ db_p, float* dwq,
                                       float* dwk, float* dwv, float* dbq, float* dbk, float* dbv) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 768 * 256) return;
  const int row = idx >> 8, c = idx & 255;
  const int which = row >> 8, hh = (row & 255) >> 6, e = row & 63;
  if (e >= D_HEAD) return;
  float* dw = which == 0 ? dwq : (which == 1 ? dwk : dwv);
  float* db = which == 0 ? dbq : (which == 1 ? dbk : dbv);
  if (c < N_T) dw[(hh * D_HEAD + e) * N_T + c] += dw_p[idx];
  if (c == 0) db[hh * D_HEAD + e] += db_p[row];
}
__global__ void unpack_wo_grad_kernel(const float* __restrict__ dw_p, float* dwo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 256 * 256) return;
  const int o = idx >> 8, c = idx & 255;
  const int hh = c >> 6, e = c & 63;
  if (o < N_T && e < D_HEAD) dwo[o * (N_HEAD * D_HEAD) + hh * D_HEAD + e] += dw_p[idx];
}
__global__ void unpack_ws_grad_kernel(const float* __restrict__ dw_p, float* dws) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N_FILT * K_SPAT) return;
  const int k2 = idx / K_SPAT, kk = idx % K_SPAT;
  const int r = kk / N_FILT, k1 = kk % N_FILT;
  dws[(k2 * N_FILT + k1) * N_CH + r] += dw_p[idx];
}

// one launch for every packed operand: segment table in kernel arguments
struct PackArgs {
  const float *value_w, *wq, *wk, *wv, *wo, *w1, *w2, *ws, *wp1, *wp2;
  float *Wv_p, *Wqkv_p, *Wo_p, *W1_p, *W2_p, *Ws_p, *Wp1_r, *Wp2_r;
};
__global__ void pack_all_kernel(PackArgs a, int rt) {
  constexpr long long N0 = 256 * 256, N1 = 768 * 256, N2 = 256 * 256, N3 = 256 * 256, N4 = 256 * 256,
                      N5 = (long long)N_FILT * K_SPAT, N6 = (long long)D_OUT * D_FEAT, N7 = (long long)D_OUT * D_OUT;
  constexpr long long TOTAL = N0 + N1 + N2 + N3 + N4 + N5 + N6 + N7;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < TOTAL; i += (long long)gridDim.x * blockDim.x) {
    long long j = i;
    if (j < N0) {                                    // value embedding [250,250] -> [256,256]
      const int r = (int)(j >> 8), c = (int)(j & 255);
      a.Wv_p[j] = (r < N_T && c < N_T) ? tf32_if(a.value_w[r * N_T + c], rt) : 0.f;
      continue;
    }
    j -= N0;
    if (j < N1) {                                    // q|k|v [248,250] x3 -> [768,256], heads padded 62 -> 64
      const int row = (int)(j >> 8), c = (int)(j & 255);
      const int which = row >> 8, hh = (row & 255) >> 6, e = row & 63;
      const float* w = which == 0 ? a.wq : (which == 1 ? a.wk : a.wv);
      a.Wqkv_p[j] = (e < D_HEAD && c < N_T) ? tf32_if(w[(hh * D_HEAD + e) * N_T + c], rt) : 0.f;
      continue;
    }
    j -= N1;
    if (j < N2) {                                    // out projection [250,248] -> [256,256]
      const int o = (int)(j >> 8), c = (int)(j & 255);
      const int hh = c >> 6, e = c & 63;
      a.Wo_p[j] = (o < N_T && e < D_HEAD) ? tf32_if(a.wo[o * (N_HEAD * D_HEAD) + hh * D_HEAD + e], rt) : 0.f;
      continue;
    }
    j -= N2;
    if (j < N3) {                                    // conv1 [256,250] -> [256,256]
      const int r = (int)(j >> 8), c = (int)(j & 255);
      a.W1_p[j] = c < N_T ? tf32_if(a.w1[r * N_T + c], rt) : 0.f;
      continue;
    }
    j -= N3;
    if (j < N4) {                                    // conv2 [250,256] -> [256,256]
      const int r = (int)(j >> 8);
      a.W2_p[j] = r < N_T ? tf32_if(a.w2[j], rt) : 0.f;
      continue;
    }
    j -= N4;
    if (j < N5) {                                    // tsconv.4.weight [k2][k1][r] -> [k2][r*40 + k1]
      const int k2 = (int)(j / K_SPAT), kk = (int)(j % K_SPAT);
      const int r = kk / N_FILT, k1 = kk % N_FILT;
      a.Ws_p[j] = tf32_if(a.ws[(k2 * N_FILT + k1) * N_CH + r], rt);
      continue;
    }
    j -= N5;
    if (j < N6) { a.Wp1_r[j] = tf32_if(a.wp1[j], rt); continue; }
    j -= N6;
    a.Wp2_r[j] = tf32_if(a.wp2[j], rt);
  }
}

// joint-subject variant: value embedding of one subject [250,250] -> [256,256] (TF32) and its token-bias table
// (bias + positional row, Embed.py:144-149); token 0 is the subject slot
__global__ void pack_value_joint_kernel(const float* __restrict__ vw, const float* __restrict__ vb,
                                        const float* __restrict__ pe, float* __restrict__ Wv_p,
                                        float* __restrict__ tokbias, int rt) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < 256 * 256) {
    const int r = idx >> 8, c = idx & 255;
    Wv_p[idx] = (r < N_T && c < N_T) ? tf32_if(vw[r * N_T + c], rt) : 0.f;
  }
  if (idx < 64 * 256) {
    const int t = idx >> 8, c = idx & 255;
    tokbias[idx] = (t > 0 && c < N_T) ? vb[c] + pe[(t - 1) * N_T + c] : 0.f;
  }
}

static int check_joint(const eegb200_atms_io* io, bool backward) {
  if (!io->joint_value_w) return 0;
  EEG_REQUIRE(io->joint_value_b && io->group_offsets && io->group_subject, "joint variant: null pointer");
  EEG_REQUIRE(io->n_subjects >= 1 && io->n_subjects <= MAX_JOINT_SUBJECTS, "joint variant: n_subjects %d outside [1,%d]",
              io->n_subjects, MAX_JOINT_SUBJECTS);
  EEG_REQUIRE(io->n_groups >= 1 && io->n_groups <= io->n_subjects, "joint variant: n_groups %d outside [1,%d]", io->n_groups,
              io->n_subjects);
  EEG_REQUIRE(io->group_offsets[0] == 0 && io->group_offsets[io->n_groups] == io->B,
              "joint variant: group_offsets must run from 0 to B=%d", io->B);
  unsigned seen = 0;
  for (int g = 0; g < io->n_groups; ++g) {
    const int sj = io->group_subject[g];
    EEG_REQUIRE(io->group_offsets[g + 1] > io->group_offsets[g], "joint variant: empty or unordered group %d", g);
    EEG_REQUIRE(sj >= 0 && sj < io->n_subjects, "joint variant: no value embedding for subject id %d (KeyError '%d' in the "
                "reference, Embed.py:144)", sj, sj);
    EEG_REQUIRE(!(seen & (1u << sj)), "joint variant: subject %d appears in two groups (sort the batch by subject)", sj);
    seen |= 1u << sj;
    EEG_REQUIRE(io->joint_value_w[sj] && io->joint_value_b[sj], "joint variant: null value embedding for subject %d", sj);
    if (backward)
      EEG_REQUIRE(io->joint_value_dw && io->joint_value_db && io->joint_value_dw[sj] && io->joint_value_db[sj],
                  "joint variant: null gradient buffer for subject %d", sj);
  }
  return 0;
}

static int pack_weights(const float* const* P, float* const* BUF, const Ws& w, cudaStream_t s) {
  const int RT = tf32_rounding();
  ProfScope _ps("pack_weights", s, 0.0, 3.2e6 * 8.0);
  PackArgs a{P[EEGB200_P_VALUE_W], P[EEGB200_P_WQ], P[EEGB200_P_WK], P[EEGB200_P_WV], P[EEGB200_P_WO], P[EEGB200_P_W1],
             P[EEGB200_P_W2], P[EEGB200_P_WS], P[EEGB200_P_WP1], P[EEGB200_P_WP2],
             w.Wv_p, w.Wqkv_p, w.Wo_p, w.W1_p, w.W2_p, w.Ws_p, w.Wp1_r, w.Wp2_r};
  pack_all_kernel<<<148 * 8, 256, 0, s>>>(a, RT);
  pack_small_kernel<<<64, 256, 0, s>>>(P[EEGB200_P_VALUE_B], BUF[EEGB200_BUF_PE], P[EEGB200_P_BQ], P[EEGB200_P_BK],
                                        P[EEGB200_P_BV], P[EEGB200_P_BO], P[EEGB200_P_B1], P[EEGB200_P_B2], w.tokbias,
                                        w.bqkv_p, w.bo_p, w.b1_p, w.b2_p);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch(2);
  return 0;
}

// ------------------------------------------------------------------------------------------------
static int run_gemm(int M, int N, int K, const float* A, int lda, int amn, const float* B, int ldb, int bmn,
                    const Epilogue& e, int split, cudaStream_t s) {
  GemmArgs g;
  g.M = M; g.N = N; g.K = K;
  g.A = {A, lda, amn};
  g.B = {B, ldb, bmn};
  g.epi = e;
  g.split_k = split;
  return gemm_launch(g, s);
}
// split-K factor for the weight-gradient GEMMs (tiny outputs, reduction over all tokens).  Measured on B200
// (tools/gemm_sweep.py): the work items (tiles x splits) must fit ONE wave of the 148 persistent CTAs -- 150 items ran
// 45 % slower than 140 (dWs: 126 -> 87 us, dWqkv: 107 -> 74 us) -- and large outputs prefer few splits because every
// split adds a full pass of red.global traffic over the output (dWp2 1024x1024: split 5 -> 41 us, split 2 -> 19 us).
static int pick_split(int M, int N, int K) {
  static int forced = -2;
  if (forced == -2) { const char* e = getenv("EEGB200_WGRAD_SPLIT"); forced = e ? atoi(e) : -1; }
  const int tiles = cdiv(M, 128) * cdiv(N, N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  const int kb = cdiv(K, 32);
  int split;
  if (forced == 0) split = cdiv(148, tiles);          // round-1 rule (A/B)
  else {
    split = 148 / tiles;
    if ((long long)M * N >= 512 * 1024 && split > 2) split = 2;
    if (split >= 64) split /= 2;
  }
  if (split > kb / 4) split = kb / 4;
  if (split < 1) split = 1;
  return split;
}
static Epilogue epi_out(float* C, int ldc) {
  Epilogue e;
  e.C = C; e.ldc = ldc;
  return e;
}
static Epilogue epi_wgrad(float* C, int ldc) {   // accumulate into a (pre-zeroed or running) gradient
  Epilogue e;
  e.C = C; e.ldc = ldc;
  e.store_mode = EPI_ATOMIC;
  return e;
}

struct Cfg {
  DropoutCfg d[EEGB200_SITE_COUNT];
};
static Cfg make_cfg(const eegb200_atms_io* io) {
  static const float ref_p[EEGB200_SITE_COUNT] = {0.f, 0.25f, 0.25f, 0.25f, 0.25f, 0.25f, 0.5f, 0.5f};
  Cfg c;
  for (int i = 0; i < EEGB200_SITE_COUNT; ++i) {
    const float p = io->dropout_p ? io->dropout_p[i] : ref_p[i];
    c.d[i] = make_dropout(io->seed, (uint32_t)i, p, io->train != 0,
                           reinterpret_cast<const unsigned long long*>(io->seed_offset_dev));
  }
  return c;
}

static int check_io(const eegb200_atms_io* io, Ws* w) {
  EEG_REQUIRE(io != nullptr, "null io");
  EEG_REQUIRE(io->B > 0, "batch must be positive (got %d)", io->B);
  EEG_REQUIRE(io->params && io->buffers && io->x && io->subject_ids && io->workspace, "null pointer in atms io");
  for (int i = 0; i < EEGB200_P_COUNT; ++i) EEG_REQUIRE(io->params[i] != nullptr, "params[%d] is null", i);
  for (int i = 0; i < EEGB200_BUF_COUNT; ++i) EEG_REQUIRE(io->buffers[i] != nullptr, "buffers[%d] is null", i);
  const size_t need = carve(nullptr, io->B, nullptr);
  EEG_REQUIRE(io->workspace_bytes >= need, "workspace too small: %zu < %zu", io->workspace_bytes, need);
  EEG_REQUIRE((reinterpret_cast<uintptr_t>(io->workspace) & 255) == 0, "workspace must be 256-byte aligned");
  carve(io->workspace, io->B, w);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
static int forward(const eegb200_atms_io* io, int phases, cudaStream_t s) {
  Ws w;
  EEG_TRY(check_io(io, &w));
  EEG_REQUIRE(io->out != nullptr || !(phases & EEGB200_PHASE_C), "null output");
  const float* const* P = io->params;
  float* const* BUF = io->buffers;
  const int B = io->B;
  const int M = B * N_TOK;
  const int R = B * N_POOL;
  const Cfg cfg = make_cfg(io);
  const int RT = tf32_rounding();
  const int train = io->train != 0;
  const long long wmul = ((phases >> 8) & 0xFF) > 1 ? ((phases >> 8) & 0xFF) : 1;   // SyncBN: statistics cover wmul*B samples

  if (phases & EEGB200_PHASE_A) {
    EEG_TRY(pack_weights(P, BUF, w, s));
    // ---- DataEmbedding (Embed.py:141-162) ----
    EEG_TRY(pad_input(io->x, w.Xp, B, s));
    if (!io->joint_value_w) {
      Epilogue e = epi_out(w.H0, 256);
      e.bias = w.tokbias; e.bias_period = 64; e.ld_bias = 256;
      e.drop = cfg.d[EEGB200_SITE_EMBED]; e.drop_ld = 256;
      e.round_tf32 = RT;
      EEG_TRY(run_gemm(M, 256, 256, w.Xp, 256, 0, w.Wv_p, 256, 0, e, 1, s));
    } else {
      // joint-subject variant (Embed.py:144): one grouped GEMM per subject present in the batch, each with its own
      // packed weights / token-bias table, into T3 (a backward temporary, free here); then one pass applies the embed
      // dropout with the same element indexing as the fused epilogue above (+ TF32 rounding) into H0
      EEG_TRY(check_joint(io, false));
      for (int g = 0; g < io->n_groups; ++g) {
        const int sj = io->group_subject[g];
        const size_t r0 = (size_t)io->group_offsets[g] * N_TOK;
        const int rows = (io->group_offsets[g + 1] - io->group_offsets[g]) * N_TOK;
        float* Wj = w.Wv_j + (size_t)sj * 256 * 256;
        float* bj = w.tokbias_j + (size_t)sj * 64 * 256;
        pack_value_joint_kernel<<<256, 256, 0, s>>>(io->joint_value_w[sj], io->joint_value_b[sj], BUF[EEGB200_BUF_PE], Wj, bj, RT);
        EEG_CUDA_OK(cudaGetLastError());
        count_launch();
        Epilogue e = epi_out(w.T3 + r0 * 256, 256);
        e.bias = bj; e.bias_period = 64; e.ld_bias = 256;
        EEG_TRY(run_gemm(rows, 256, 256, w.Xp + r0 * 256, 256, 0, Wj, 256, 0, e, 1, s));
      }
      EEG_TRY(dropout_apply(w.T3, w.H0, M, 256, cfg.d[EEGB200_SITE_EMBED], RT, s));
    }
    EEG_TRY(subject_token(reinterpret_cast<const long long*>(io->subject_ids), P[EEGB200_P_SUBJ_TABLE],
                          P[EEGB200_P_SUBJ_SHARED], io->n_subjects, w.flag, w.H0, B, cfg.d[EEGB200_SITE_EMBED], RT, s));
    // ---- AttentionLayer (SelfAttention_Family.py:194-213) ----
    {
      Epilogue e = epi_out(w.QKV, 768);
      e.bias = w.bqkv_p;
      if (RT && attention_tc_enabled()) e.round_tf32 = 1;     // tcgen05 reads Q/K/V as they are (kind::tf32 truncates)
      EEG_TRY(run_gemm(M, 768, 256, w.H0, 256, 0, w.Wqkv_p, 256, 0, e, 1, s));
    }
    EEG_TRY(attention_fwd(w.QKV, w.O, B, cfg.d[EEGB200_SITE_ATTN], s));
    {
      Epilogue e = epi_out(w.R1, 256);          // x + dropout(out_projection(attn))   (Transformer_EncDec.py:45)
      e.bias = w.bo_p;
      e.drop = cfg.d[EEGB200_SITE_RES1]; e.drop_ld = 256;
      e.resid = w.H0; e.ld_res = 256;
      EEG_TRY(run_gemm(M, 256, 256, w.O, 256, 0, w.Wo_p, 256, 0, e, 1, s));
    }
    EEG_TRY(layernorm_fwd(w.R1, 256, M, N_T, P[EEGB200_P_LN1_G], P[EEGB200_P_LN1_B], w.st1, nullptr, nullptr, nullptr,
                          w.X1, 256, RT, s));
    // ---- position-wise FFN (Transformer_EncDec.py:48-51) ----
    {
      Epilogue e = epi_out(w.Hf, 256);
      e.bias = w.b1_p;
      e.aux_out = w.U; e.ld_aux = 256;
      e.act = EPI_ACT_GELU;
      e.drop = cfg.d[EEGB200_SITE_FFN1]; e.drop_ld = 256;
      e.round_tf32 = RT;
      EEG_TRY(run_gemm(M, 256, 256, w.X1, 256, 0, w.W1_p, 256, 0, e, 1, s));
    }
    {
      Epilogue e = epi_out(w.R2, 256);
      e.bias = w.b2_p;
      e.drop = cfg.d[EEGB200_SITE_FFN2]; e.drop_ld = 256;
      e.resid = w.X1; e.ld_res = 256;
      EEG_TRY(run_gemm(M, 256, 256, w.Hf, 256, 0, w.W2_p, 256, 0, e, 1, s));
    }
    // norm2 then the encoder's final norm (Transformer_EncDec.py:51, 77-78)
    EEG_TRY(layernorm_fwd(w.R2, 256, M, N_T, P[EEGB200_P_LN2_G], P[EEGB200_P_LN2_B], w.st2, P[EEGB200_P_LNF_G],
                          P[EEGB200_P_LNF_B], w.stf, w.X3, 256, 0, s));
    // ---- PatchEmbedding temporal conv + pool (ATMS_retrieval.py:102-103) on tokens 0..62 ----
    if (train) EEG_CUDA_OK(cudaMemsetAsync(w.bn1_sums, 0, 2 * N_FILT * sizeof(double), s));
    if (RT && conv_tc_enabled()) {
      // fused tcgen05 path (conv_tc.cu): statistics pass only; the activations are recomputed by the apply pass in phase B
      if (train) EEG_TRY(conv_tc_stats(w.X3, P[EEGB200_P_WT], P[EEGB200_P_BT], w.bn1_sums, B, s));
    } else {
      EEG_TRY(conv_temporal_fwd(w.X3, P[EEGB200_P_WT], P[EEGB200_P_BT], w.Y1, train ? w.bn1_sums : nullptr, B, s));
    }
  }
  if (phases & EEGB200_PHASE_B) {
    BnState bn1{w.bn1_sums, w.bn1_mr, P[EEGB200_P_BN1_G], P[EEGB200_P_BN1_B], BUF[EEGB200_BUF_BN1_RM], BUF[EEGB200_BUF_BN1_RV]};
    EEG_TRY(bn_finalize(bn1, wmul * B * N_CH * N_POOL, train, train && io->update_running_stats, s));
    if (RT && conv_tc_enabled()) {
      // conv + pool + BN1 + ELU + spatial conv in one kernel; y1 / a1 never leave the chip (debug stores: stage checks)
      const bool dbg = debug_stores() != 0;
      EEG_TRY(conv_tc_apply(w.X3, P[EEGB200_P_WT], P[EEGB200_P_BT], w.bn1_mr, P[EEGB200_P_BN1_G], P[EEGB200_P_BN1_B],
                            P[EEGB200_P_WS], P[EEGB200_P_BS], w.Ws_tc, dbg ? w.Y1 : nullptr, dbg ? w.A1 : nullptr, w.Y2, B, s));
    } else {
      EEG_TRY(bn_elu_apply(w.Y1, w.bn1_mr, P[EEGB200_P_BN1_G], P[EEGB200_P_BN1_B], w.A1, (long long)R * K_SPAT, RT, s));
      Epilogue e = epi_out(w.Y2, N_FILT);        // spatial conv (63,1) == GEMM over (r,k1)  (ATMS_retrieval.py:106)
      e.bias = P[EEGB200_P_BS];
      EEG_TRY(run_gemm(R, N_FILT, K_SPAT, w.A1, K_SPAT, 0, w.Ws_p, K_SPAT, 0, e, 1, s));
    }
    if (train) {
      EEG_CUDA_OK(cudaMemsetAsync(w.bn2_sums, 0, 2 * N_FILT * sizeof(double), s));
      EEG_TRY(colstats(w.Y2, N_FILT, R, N_FILT, w.bn2_sums, s));
    }
  }
  if (phases & EEGB200_PHASE_C) {
    BnState bn2{w.bn2_sums, w.bn2_mr, P[EEGB200_P_BN2_G], P[EEGB200_P_BN2_B], BUF[EEGB200_BUF_BN2_RM], BUF[EEGB200_BUF_BN2_RV]};
    EEG_TRY(bn_finalize(bn2, wmul * R, train, train && io->update_running_stats, s));
    EEG_TRY(conv_head_fwd(w.Y2, w.bn2_mr, P[EEGB200_P_BN2_G], P[EEGB200_P_BN2_B], P[EEGB200_P_WC], P[EEGB200_P_BC], w.feat,
                          B, cfg.d[EEGB200_SITE_CONV], s));
    // ---- Proj_eeg (ATMS_retrieval.py:157-167) ----
    {
      Epilogue e = epi_out(w.G, D_OUT);
      e.bias = P[EEGB200_P_BP1];
      e.aux_out = w.Z1; e.ld_aux = D_OUT;
      e.act = EPI_ACT_GELU;
      e.round_tf32 = RT;
      EEG_TRY(run_gemm(B, D_OUT, D_FEAT, w.feat, D_FEAT, 0, w.Wp1_r, D_FEAT, 0, e, 1, s));
    }
    {
      Epilogue e = epi_out(w.Z2, D_OUT);
      e.bias = P[EEGB200_P_BP2];
      e.drop = cfg.d[EEGB200_SITE_PROJ]; e.drop_ld = D_OUT;
      e.resid = w.Z1; e.ld_res = D_OUT;
      EEG_TRY(run_gemm(B, D_OUT, D_OUT, w.G, D_OUT, 0, w.Wp2_r, D_OUT, 0, e, 1, s));
    }
    EEG_TRY(layernorm_fwd(w.Z2, D_OUT, B, D_OUT, P[EEGB200_P_LNP_G], P[EEGB200_P_LNP_B], w.stp, nullptr, nullptr, nullptr,
                          io->out, D_OUT, 0, s));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// The weight / bias gradient kernels (split-K GEMMs over all tokens, column sums, unpack) are off the critical path:
// only the optimiser needs them.  They are forked onto a side stream as soon as their inputs exist and joined before
// the call returns, so they fill the SMs / HBM bandwidth the latency-bound dX chain leaves idle.  Works the same under
// CUDA-graph capture (fork / join become parallel graph branches).
struct SideStream {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool used = false;
  int init() {
    if (side) return 0;
    EEG_CUDA_OK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    EEG_CUDA_OK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    EEG_CUDA_OK(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
    return 0;
  }
  // everything enqueued on `main` so far becomes visible to the side stream
  int sync_from(cudaStream_t main) {
    EEG_CUDA_OK(cudaEventRecord(fork, main));
    EEG_CUDA_OK(cudaStreamWaitEvent(side, fork, 0));
    used = true;
    return 0;
  }
  int join_into(cudaStream_t main) {
    if (!used) return 0;
    EEG_CUDA_OK(cudaEventRecord(join, side));
    EEG_CUDA_OK(cudaStreamWaitEvent(main, join, 0));
    used = false;
    return 0;
  }
};
static SideStream g_side_by_dev[64];      // side stream + events belong to a device
static int side_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("EEGB200_SIDE_STREAM");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on;
}

// EEGB200_LN_FUSED=0 falls back to the separate LayerNorm-backward + dropout passes (A/B switch)
static int ln_fused() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("EEGB200_LN_FUSED");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on;
}

static int backward(const eegb200_atms_io* io, const float* d_out, float* const* GR, int phases, cudaStream_t s) {
  Ws w;
  EEG_TRY(check_io(io, &w));
  EEG_REQUIRE(io->train != 0, "backward needs the workspace of a train-mode forward");
  EEG_REQUIRE(GR != nullptr, "null grads");
  for (int i = 0; i < EEGB200_P_COUNT; ++i)
    EEG_REQUIRE(GR[i] != nullptr || i == EEGB200_P_SUBJ_TABLE || i == EEGB200_P_SUBJ_SHARED, "grads[%d] is null", i);
  const float* const* P = io->params;
  const int B = io->B;
  const int M = B * N_TOK;
  const int R = B * N_POOL;
  const Cfg cfg = make_cfg(io);
  const int RT = tf32_rounding();
  const long long wmul = ((phases >> 8) & 0xFF) > 1 ? ((phases >> 8) & 0xFF) : 1;
  const bool use_side = side_enabled() && !prof_enabled();
  SideStream& g_side = g_side_by_dev[current_device()];
  if (use_side) EEG_TRY(g_side.init());
  cudaStream_t ws = use_side ? g_side.side : s;      // stream of the weight-gradient work
#define FORK() do { if (use_side) EEG_TRY(g_side.sync_from(s)); } while (0)

  if (phases & EEGB200_PHASE_A) {
    EEG_REQUIRE(d_out != nullptr, "null d_out");
    // ---- Proj_eeg backward ----
    EEG_TRY(layernorm_bwd(d_out, D_OUT, w.Z2, D_OUT, B, D_OUT, P[EEGB200_P_LNP_G], P[EEGB200_P_LNP_B], w.stp, nullptr,
                          nullptr, w.dZ2, D_OUT, GR[EEGB200_P_LNP_G], GR[EEGB200_P_LNP_B], nullptr, nullptr, 0, s));
    EEG_TRY(dropout_apply(w.dZ2, w.dZ2d, B, D_OUT, cfg.d[EEGB200_SITE_PROJ], RT, s));
    FORK();
    EEG_TRY(colsum(w.dZ2d, D_OUT, B, D_OUT, GR[EEGB200_P_BP2], 0, 0, ws));
    EEG_TRY(run_gemm(D_OUT, D_OUT, B, w.dZ2d, D_OUT, 1, w.G, D_OUT, 1, epi_wgrad(GR[EEGB200_P_WP2], D_OUT),
                     pick_split(D_OUT, D_OUT, B), ws));
    {
      Epilogue e = epi_out(w.dZ1, D_OUT);        // dZ1 = dZ2 + (dZ2d . Wp2) * GELU'(Z1)
      e.mul_in = w.Z1; e.ld_mul = D_OUT;
      e.resid = w.dZ2; e.ld_res = D_OUT;
      e.round_tf32 = RT;
      EEG_TRY(run_gemm(B, D_OUT, D_OUT, w.dZ2d, D_OUT, 0, w.Wp2_r, D_OUT, 1, e, 1, s));
    }
    FORK();
    EEG_TRY(colsum(w.dZ1, D_OUT, B, D_OUT, GR[EEGB200_P_BP1], 0, 0, ws));
    EEG_TRY(run_gemm(D_OUT, D_FEAT, B, w.dZ1, D_OUT, 1, w.feat, D_FEAT, 1, epi_wgrad(GR[EEGB200_P_WP1], D_FEAT),
                     pick_split(D_OUT, D_FEAT, B), ws));
    {
      GemmArgs g;                                // dfeat = dZ1 . Wp1 (128-wide tiles: 25 vs 34 us at B = 1024)
      g.M = B; g.N = D_FEAT; g.K = D_OUT;
      g.A = {w.dZ1, D_OUT, 0};
      g.B = {w.Wp1_r, D_FEAT, 1};
      g.epi = epi_out(w.dfeat, D_FEAT);
      g.tile_n = 128;
      EEG_TRY(gemm_launch(g, s));
    }
    // ---- conv head backward down to d(BN2 out) + BN2 reductions ----
    EEG_CUDA_OK(cudaMemsetAsync(w.bn2_bsums, 0, 2 * N_FILT * sizeof(double), s));
    EEG_TRY(conv_head_bwd(w.dfeat, w.Y2, w.bn2_mr, P[EEGB200_P_BN2_G], P[EEGB200_P_BN2_B], P[EEGB200_P_WC], w.dz2,
                          GR[EEGB200_P_WC], GR[EEGB200_P_BC], w.bn2_bsums, B, cfg.d[EEGB200_SITE_CONV], s));
  }
  if (phases & EEGB200_PHASE_B) {
    EEG_TRY(bn_bwd_apply(w.dz2, w.Y2, w.bn2_mr, P[EEGB200_P_BN2_G], w.bn2_bsums, wmul * R, w.dY2, GR[EEGB200_P_BN2_G],
                         GR[EEGB200_P_BN2_B], (long long)R * N_FILT, RT, 1.f / (float)wmul, s));
    FORK();
    EEG_TRY(colsum(w.dY2, N_FILT, R, N_FILT, GR[EEGB200_P_BS], 0, 0, ws));
    if (RT && conv_tc_enabled()) {
      // fused tcgen05 path: d a1 / a1 are recomputed on chip; this pass delivers the BatchNorm1-backward reductions and dWs
      EEG_CUDA_OK(cudaMemsetAsync(w.bn1_bsums, 0, 2 * N_FILT * sizeof(double), s));
      EEG_TRY(conv_tc_bwd_stats(w.X3, P[EEGB200_P_WT], P[EEGB200_P_BT], w.bn1_mr, P[EEGB200_P_BN1_G], P[EEGB200_P_BN1_B],
                                P[EEGB200_P_WS], w.dY2, w.Ws_tc, w.bn1_bsums, GR[EEGB200_P_WS], B, s));
    } else {
      // dWs[k2][(r,k1)] = sum_{(b,j)} dY2[(b,j)][k2] * A1[(b,j)][(r,k1)]
      EEG_CUDA_OK(cudaMemsetAsync(w.dWs_p, 0, (size_t)N_FILT * K_SPAT * sizeof(float), ws));
      EEG_TRY(run_gemm(N_FILT, K_SPAT, R, w.dY2, N_FILT, 1, w.A1, K_SPAT, 1, epi_wgrad(w.dWs_p, K_SPAT),
                       pick_split(N_FILT, K_SPAT, R), ws));
      unpack_ws_grad_kernel<<<cdiv(N_FILT * K_SPAT, 256), 256, 0, ws>>>(w.dWs_p, GR[EEGB200_P_WS]);
      count_launch();
      // dA1 = dY2 . Ws, then dz1 = dA1 * ELU'(BN1(y1)) and the two BN1-backward reductions
      EEG_CUDA_OK(cudaMemsetAsync(w.bn1_bsums, 0, 2 * N_FILT * sizeof(double), s));
      if (RT) {
        // tensor-core path: ELU', the y1 read and both reductions ride in the GEMM epilogue (no separate 1.1 GB pass)
        Epilogue e = epi_out(w.dA1, K_SPAT);
        e.bn_y = w.Y1; e.ld_bn_y = K_SPAT;
        e.bn_mean_rstd = w.bn1_mr; e.bn_gamma = P[EEGB200_P_BN1_G]; e.bn_beta = P[EEGB200_P_BN1_B];
        e.bn_sums = w.bn1_bsums;
        EEG_TRY(run_gemm(R, K_SPAT, N_FILT, w.dY2, N_FILT, 0, w.Ws_p, K_SPAT, 1, e, 1, s));
      } else {
        EEG_TRY(run_gemm(R, K_SPAT, N_FILT, w.dY2, N_FILT, 0, w.Ws_p, K_SPAT, 1, epi_out(w.dA1, K_SPAT), 1, s));
        EEG_TRY(bn1_bwd_reduce(w.dA1, w.Y1, w.bn1_mr, P[EEGB200_P_BN1_G], P[EEGB200_P_BN1_B], w.bn1_bsums,
                               (long long)R * K_SPAT, s));
      }
    }
  }
  if (phases & EEGB200_PHASE_C) {
    if (RT && conv_tc_enabled()) {
      EEG_TRY(conv_tc_bwd_apply(w.X3, P[EEGB200_P_WT], P[EEGB200_P_BT], w.bn1_mr, P[EEGB200_P_BN1_G], P[EEGB200_P_BN1_B],
                                w.dY2, w.Ws_tc, w.bn1_bsums, wmul * B * N_CH * N_POOL, 1.f / (float)wmul, w.dX3,
                                GR[EEGB200_P_WT], GR[EEGB200_P_BT], GR[EEGB200_P_BN1_G], GR[EEGB200_P_BN1_B], B, s));
    } else {
      EEG_TRY(conv_temporal_bwd(w.dA1, w.Y1, w.X3, P[EEGB200_P_WT], w.bn1_mr, P[EEGB200_P_BN1_G], w.bn1_bsums,
                                wmul * B * N_CH * N_POOL, w.dX3, GR[EEGB200_P_WT], GR[EEGB200_P_BT],
                                GR[EEGB200_P_BN1_G], GR[EEGB200_P_BN1_B], B, 1.f / (float)wmul, s));
    }
    // ---- final norm + norm2 ----
    // (the LayerNorm backward also emits T1 = dropout_ffn2(dR2), the operand of the FFN GEMMs, and db2 = colsum(T1))
    if (ln_fused()) {
      EEG_TRY(layernorm_bwd_tok(w.dX3, w.R2, M, N_T, P[EEGB200_P_LN2_G], P[EEGB200_P_LN2_B], w.st2, P[EEGB200_P_LNF_G],
                                w.stf, w.dR2, GR[EEGB200_P_LN2_G], GR[EEGB200_P_LN2_B], GR[EEGB200_P_LNF_G],
                                GR[EEGB200_P_LNF_B], w.T1, cfg.d[EEGB200_SITE_FFN2], RT, GR[EEGB200_P_B2], s));
    } else {
      EEG_TRY(layernorm_bwd(w.dX3, 256, w.R2, 256, M, N_T, P[EEGB200_P_LN2_G], P[EEGB200_P_LN2_B], w.st2,
                            P[EEGB200_P_LNF_G], w.stf, w.dR2, 256, GR[EEGB200_P_LN2_G], GR[EEGB200_P_LN2_B],
                            GR[EEGB200_P_LNF_G], GR[EEGB200_P_LNF_B], 0, s));
      EEG_TRY(dropout_apply_colsum(w.dR2, w.T1, M, 256, cfg.d[EEGB200_SITE_FFN2], RT, GR[EEGB200_P_B2], N_T, 0, 0, s));
    }
    // ---- FFN ----
    FORK();
    EEG_TRY(run_gemm(N_T, D_FF, M, w.T1, 256, 1, w.Hf, 256, 1, epi_wgrad(GR[EEGB200_P_W2], D_FF), pick_split(N_T, D_FF, M), ws));
    {
      Epilogue e = epi_out(w.dU, 256);           // dU = dropout_ffn1(T1 . W2) * GELU'(U)
      e.drop = cfg.d[EEGB200_SITE_FFN1]; e.drop_ld = 256;
      e.mul_in = w.U; e.ld_mul = 256;
      e.round_tf32 = RT;
      EEG_TRY(run_gemm(M, 256, 256, w.T1, 256, 0, w.W2_p, 256, 1, e, 1, s));
    }
    FORK();
    EEG_TRY(colsum(w.dU, 256, M, D_FF, GR[EEGB200_P_B1], 0, 0, ws));
    EEG_TRY(run_gemm(D_FF, N_T, M, w.dU, 256, 1, w.X1, 256, 1, epi_wgrad(GR[EEGB200_P_W1], N_T), pick_split(D_FF, N_T, M), ws));
    {
      Epilogue e = epi_out(w.dX1, 256);          // dX1 = dR2 + dU . W1
      e.resid = w.dR2; e.ld_res = 256;
      EEG_TRY(run_gemm(M, 256, 256, w.dU, 256, 0, w.W1_p, 256, 1, e, 1, s));
    }
    // ---- norm1 ----
    if (ln_fused()) {
      EEG_TRY(layernorm_bwd_tok(w.dX1, w.R1, M, N_T, P[EEGB200_P_LN1_G], P[EEGB200_P_LN1_B], w.st1, nullptr, nullptr, w.dR1,
                                GR[EEGB200_P_LN1_G], GR[EEGB200_P_LN1_B], nullptr, nullptr, w.T2,
                                cfg.d[EEGB200_SITE_RES1], RT, GR[EEGB200_P_BO], s));
    } else {
      EEG_TRY(layernorm_bwd(w.dX1, 256, w.R1, 256, M, N_T, P[EEGB200_P_LN1_G], P[EEGB200_P_LN1_B], w.st1, nullptr, nullptr,
                            w.dR1, 256, GR[EEGB200_P_LN1_G], GR[EEGB200_P_LN1_B], nullptr, nullptr, 0, s));
      EEG_TRY(dropout_apply_colsum(w.dR1, w.T2, M, 256, cfg.d[EEGB200_SITE_RES1], RT, GR[EEGB200_P_BO], N_T, 0, 0, s));
    }
    // ---- attention ----
    FORK();
    EEG_CUDA_OK(cudaMemsetAsync(w.dWo_p, 0, 256 * 256 * sizeof(float), ws));
    EEG_TRY(run_gemm(256, 256, M, w.T2, 256, 1, w.O, 256, 1, epi_wgrad(w.dWo_p, 256), pick_split(256, 256, M), ws));
    unpack_wo_grad_kernel<<<256, 256, 0, ws>>>(w.dWo_p, GR[EEGB200_P_WO]);
    count_launch();
    EEG_TRY(run_gemm(M, 256, 256, w.T2, 256, 0, w.Wo_p, 256, 1, epi_out(w.dO, 256), 1, s));
    EEG_TRY(attention_bwd(w.QKV, w.dO, w.dQKV, B, cfg.d[EEGB200_SITE_ATTN], s));
    FORK();
    EEG_CUDA_OK(cudaMemsetAsync(w.dWqkv_p, 0, 768 * 256 * sizeof(float), ws));
    EEG_CUDA_OK(cudaMemsetAsync(w.dbqkv_p, 0, 768 * sizeof(float), ws));
    EEG_TRY(colsum(w.dQKV, 768, M, 768, w.dbqkv_p, 0, 0, ws));
    EEG_TRY(run_gemm(768, 256, M, w.dQKV, 768, 1, w.H0, 256, 1, epi_wgrad(w.dWqkv_p, 256), pick_split(768, 256, M), ws));
    unpack_qkv_grad_kernel<<<768, 256, 0, ws>>>(w.dWqkv_p, w.dbqkv_p, GR[EEGB200_P_WQ], GR[EEGB200_P_WK], GR[EEGB200_P_WV],
                                                GR[EEGB200_P_BQ], GR[EEGB200_P_BK], GR[EEGB200_P_BV]);
    count_launch();
    // ---- dH0 = dR1 + dQKV . Wqkv, then the DataEmbedding backward ----
    // Only T3 = tf32(dropout_embed(dH0)) is consumed downstream (value-embedding weight / bias gradients; the subject-token
    // gradient reads its rows with the same mask), so the GEMM epilogue applies residual -> dropout -> rounding and writes
    // T3 directly: the separate 134 MB dropout pass (51 us on the critical path) is gone and the bias column sums run on
    // the side stream.  EEGB200_DH0_FUSED=0 restores the two-pass form (A/B switch).
    static int dh0_fused = -1;
    if (dh0_fused < 0) { const char* e_ = getenv("EEGB200_DH0_FUSED"); dh0_fused = (e_ && e_[0] == '0') ? 0 : 1; }
    const DropoutCfg no_drop = make_dropout(0, 0, 0.f, false);
    if (dh0_fused) {
      Epilogue e = epi_out(w.T3, 256);
      e.resid = w.dR1; e.ld_res = 256; e.resid_before_drop = 1;
      e.drop = cfg.d[EEGB200_SITE_EMBED]; e.drop_ld = 256;
      e.round_tf32 = RT;
      EEG_TRY(run_gemm(M, 256, 768, w.dQKV, 768, 0, w.Wqkv_p, 256, 1, e, 1, s));
    } else {
      Epilogue e = epi_out(w.dH0, 256);
      e.resid = w.dR1; e.ld_res = 256;
      EEG_TRY(run_gemm(M, 256, 768, w.dQKV, 768, 0, w.Wqkv_p, 256, 1, e, 1, s));
    }
    // token-0 rows carry no value embedding -> excluded from the bias gradient
    if (!io->joint_value_w) {
      if (dh0_fused) {
        FORK();
        EEG_TRY(colsum(w.T3, 256, M, N_T, GR[EEGB200_P_VALUE_B], N_TOK, 0, ws));
      } else {
        EEG_TRY(dropout_apply_colsum(w.dH0, w.T3, M, 256, cfg.d[EEGB200_SITE_EMBED], RT, GR[EEGB200_P_VALUE_B], N_T, N_TOK, 0, s));
        FORK();
      }
      EEG_TRY(run_gemm(N_T, N_T, M, w.T3, 256, 1, w.Xp, 256, 1, epi_wgrad(GR[EEGB200_P_VALUE_W], N_T), pick_split(N_T, N_T, M), ws));
    } else {
      // joint-subject variant: every subject's value embedding gets the gradient of its own trials only
      EEG_TRY(check_joint(io, true));
      if (!dh0_fused) EEG_TRY(dropout_apply(w.dH0, w.T3, M, 256, cfg.d[EEGB200_SITE_EMBED], RT, s));
      FORK();
      for (int g = 0; g < io->n_groups; ++g) {
        const int sj = io->group_subject[g];
        const size_t r0 = (size_t)io->group_offsets[g] * N_TOK;
        const int rows = (io->group_offsets[g + 1] - io->group_offsets[g]) * N_TOK;
        EEG_TRY(colsum(w.T3 + r0 * 256, 256, rows, N_T, io->joint_value_db[sj], N_TOK, 0, ws));
        EEG_TRY(run_gemm(N_T, N_T, rows, w.T3 + r0 * 256, 256, 1, w.Xp + r0 * 256, 256, 1,
                         epi_wgrad(io->joint_value_dw[sj], N_T), pick_split(N_T, N_T, rows), ws));
      }
    }
    EEG_REQUIRE(GR[EEGB200_P_SUBJ_TABLE] != nullptr && GR[EEGB200_P_SUBJ_SHARED] != nullptr,
                "subject-token gradients need both the table and the shared-token grad buffers");
    // (fused form: the token-0 rows of T3 already carry the embed dropout mask and scale)
    EEG_TRY(subject_token_bwd(reinterpret_cast<const long long*>(io->subject_ids), w.flag, dh0_fused ? w.T3 : w.dH0,
                              GR[EEGB200_P_SUBJ_TABLE], GR[EEGB200_P_SUBJ_SHARED], B,
                              dh0_fused ? no_drop : cfg.d[EEGB200_SITE_EMBED], s));
    EEG_CUDA_OK(cudaGetLastError());
  }
  if (use_side) EEG_TRY(g_side.join_into(s));      // gradients are complete when the caller's stream continues
#undef FORK
  return 0;
}

struct NamedTensor { const char* name; void* ptr; int rows, cols, ld; };

}  // namespace eegb200

using namespace eegb200;

extern "C" {

size_t eegb200_atms_workspace_bytes(int B) { return B > 0 ? carve(nullptr, B, nullptr) : 0; }

int eegb200_atms_forward(const eegb200_atms_io* io, int phase_mask, void* stream) {
  return forward(io, phase_mask, (cudaStream_t)stream);
}
int eegb200_atms_backward(const eegb200_atms_io* io, const float* d_out, float* const* grads, int phase_mask, void* stream) {
  return backward(io, d_out, grads, phase_mask, (cudaStream_t)stream);
}

int eegb200_atms_ws_tensor(void* workspace, int B, const char* name, void** ptr, int* rows, int* cols, int* ld) {
  EEG_REQUIRE(workspace && name && ptr && rows && cols && ld && B > 0, "ws_tensor: bad arguments");
  Ws w;
  carve(workspace, B, &w);
  const int M = B * N_TOK, R = B * N_POOL;
  const NamedTensor t[] = {
      {"xp", w.Xp, M, 256, 256},      {"h0", w.H0, M, 256, 256},       {"qkv", w.QKV, M, 768, 768},
      {"attn_o", w.O, M, 256, 256},   {"r1", w.R1, M, 256, 256},       {"x1", w.X1, M, 256, 256},
      {"ffn_u", w.U, M, 256, 256},    {"ffn_h", w.Hf, M, 256, 256},    {"r2", w.R2, M, 256, 256},
      {"x3", w.X3, M, 256, 256},      {"y1", w.Y1, R, K_SPAT, K_SPAT}, {"a1", w.A1, R, K_SPAT, K_SPAT},
      {"y2", w.Y2, R, N_FILT, N_FILT}, {"feat", w.feat, B, D_FEAT, D_FEAT}, {"z1", w.Z1, B, D_OUT, D_OUT},
      {"z2", w.Z2, B, D_OUT, D_OUT},  {"bn1_sums", w.bn1_sums, 1, 80, 80}, {"bn2_sums", w.bn2_sums, 1, 80, 80},
      {"bn1_bwd_sums", w.bn1_bsums, 1, 80, 80}, {"bn2_bwd_sums", w.bn2_bsums, 1, 80, 80},
      {"dx3", w.dX3, M, 256, 256},    {"dh0", w.dH0, M, 256, 256},     {"dfeat", w.dfeat, B, D_FEAT, D_FEAT},
      {"dqkv", w.dQKV, M, 768, 768},  {"dr1", w.dR1, M, 256, 256},     {"dr2", w.dR2, M, 256, 256},
      {"da1", w.dA1, R, K_SPAT, K_SPAT}, {"dy2", w.dY2, R, N_FILT, N_FILT},
  };
  if ((strcmp(name, "y1") == 0 || strcmp(name, "a1") == 0 || strcmp(name, "da1") == 0) && !need_conv_intermediates()) {
    set_error("ws_tensor: '%s' stays on chip in the fused conv path; call eegb200_set_debug_stores(1) before the forward", name);
    return 2;
  }
  for (const NamedTensor& e : t)
    if (strcmp(e.name, name) == 0) {
      *ptr = e.ptr; *rows = e.rows; *cols = e.cols; *ld = e.ld;
      return 0;
    }
  set_error("ws_tensor: unknown name '%s'", name);
  return 2;
}

}  // extern "C"

// C-ABI entry points (include/eegdecode_b200.h).
#include "../../include/eegdecode_b200.h"
#include "gemm.h"

namespace eegb200 {
static int g_debug_stores = 0;
int debug_stores() { return g_debug_stores; }
}  // namespace eegb200

using namespace eegb200;

extern "C" {

int eegb200_abi_version(void) { return EEGB200_ABI_VERSION; }
const char* eegb200_last_error(void) { return get_error(); }
long long eegb200_launch_count(void) { return total_launch_count(); }
int eegb200_set_gemm_backend(int backend) {
  if (backend != GEMM_BACKEND_TCGEN05 && backend != GEMM_BACKEND_SIMT_FP32) {
    set_error("unknown gemm backend %d", backend);
    return 2;
  }
  gemm_set_backend(backend);
  return 0;
}
int eegb200_get_gemm_backend(void) { return gemm_get_backend(); }
int eegb200_set_debug_stores(int on) { g_debug_stores = on ? 1 : 0; return 0; }
int eegb200_prof_enable(int on) { prof_enable(on); return 0; }
int eegb200_prof_report(char* buf, size_t cap) {
  EEG_REQUIRE(buf && cap > 2, "prof_report: bad buffer");
  int rc = prof_report(buf, cap);
  EEG_REQUIRE(rc == 0, "prof_report: buffer too small");
  return 0;
}

int eegb200_gemm(const eegb200_gemm_desc* d, void* stream) {
  EEG_REQUIRE(d != nullptr, "null gemm desc");
  GemmArgs g;
  g.M = d->M; g.N = d->N; g.K = d->K;
  g.A = {d->A, d->lda, d->a_mn_major};
  g.B = {d->B, d->ldb, d->b_mn_major};
  Epilogue& e = g.epi;
  e.C = d->C; e.ldc = d->ldc; e.alpha = d->alpha;
  e.bias = d->bias; e.bias_period = d->bias_period; e.ld_bias = d->ld_bias;
  e.aux_out = d->aux_out; e.ld_aux = d->ld_aux;
  e.act = d->act;
  e.drop = make_dropout(d->drop_seed, d->drop_site, d->drop_p, true);
  e.drop_ld = d->drop_ld;
  e.mul_in = d->mul_in; e.ld_mul = d->ld_mul;
  e.resid = d->resid; e.ld_res = d->ld_res;
  e.round_tf32 = d->round_tf32;
  e.store_mode = d->store_mode;
  g.split_k = d->split_k;
  g.tile_n = d->tile_n;
  return gemm_launch(g, (cudaStream_t)stream);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
#include "kernels.h"

namespace {
struct InfoWs {
  float *E_r, *T_r, *logits, *row_lse, *diag, *part_max, *part_sum, *col_lse;
  int ld, ncol, chunks;
};
size_t info_carve(void* base, int B, int N, int D, int nt, InfoWs* out) {
  uint8_t* b = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t n) {
    off = align_up(off, 256);
    float* p = b ? reinterpret_cast<float*>(b + off) : nullptr;
    off += n * sizeof(float);
    return p;
  };
  InfoWs w;
  w.ncol = nt * N;
  w.ld = (w.ncol + 3) / 4 * 4;
  w.chunks = infonce_col_chunks(B);
  w.E_r = take((size_t)B * D);
  w.T_r = take((size_t)w.ncol * D);
  w.logits = take((size_t)B * w.ld);
  w.row_lse = take((size_t)nt * B);
  w.diag = take((size_t)nt * B);
  w.part_max = take((size_t)w.chunks * w.ncol);
  w.part_sum = take((size_t)w.chunks * w.ncol);
  w.col_lse = take((size_t)w.ncol);
  if (out) *out = w;
  return align_up(off, 256);
}
}  // namespace

extern "C" {

int eegb200_dropout_mask(uint64_t seed, uint32_t site, float p, int rows, int cols, int ld, float* out, void* stream) {
  EEG_REQUIRE(out && rows > 0 && cols > 0 && ld >= cols, "dropout_mask: bad arguments");
  return dropout_mask(make_dropout(seed, site, p, true), rows, cols, ld, out, (cudaStream_t)stream);
}

size_t eegb200_infonce_workspace_bytes(int B, int N, int D, int n_targets) {
  if (B <= 0 || N <= 0 || D <= 0 || n_targets < 1 || n_targets > 2) return 0;
  return info_carve(nullptr, B, N, D, n_targets, nullptr);
}

size_t eegb200_infonce_target_offset(int B, int N, int D, int n_targets) {
  if (B <= 0 || N <= 0 || D <= 0 || n_targets < 1 || n_targets > 2) return 0;
  InfoWs w;
  uint8_t* fake = reinterpret_cast<uint8_t*>(uintptr_t(1) << 20);      // any 256-aligned base: only the offset is used
  info_carve(fake, B, N, D, n_targets, &w);
  return (size_t)(reinterpret_cast<uint8_t*>(w.T_r) - fake);
}

int eegb200_tf32_round(const float* src, float* dst, int rows, int D, void* stream) {
  EEG_REQUIRE(src && dst && rows > 0 && D > 0 && (D & 3) == 0, "tf32_round: bad arguments");
  return pad_copy(src, D, rows, D, dst, D, rows, tf32_rounding(), 1.f, (cudaStream_t)stream);
}

int eegb200_infonce(const eegb200_infonce_io* io, int phase_mask, void* stream) {
  EEG_REQUIRE(io && io->eeg && io->tgt_img && io->logit_scale && io->workspace && io->col_stats, "infonce: null pointer");
  EEG_REQUIRE(io->B > 0 && io->N >= io->B && io->D > 0 && (io->D & 3) == 0, "infonce: bad shape B=%d N=%d D=%d", io->B,
              io->N, io->D);
  EEG_REQUIRE(io->row_offset >= 0 && io->row_offset + io->B <= io->N, "infonce: row block [%d,%d) outside N=%d",
              io->row_offset, io->row_offset + io->B, io->N);
  cudaStream_t s = (cudaStream_t)stream;
  const int nt = io->tgt_txt ? 2 : 1;
  InfoWs w;
  const size_t need = info_carve(nullptr, io->B, io->N, io->D, nt, nullptr);
  EEG_REQUIRE(io->workspace_bytes >= need, "infonce: workspace too small: %zu < %zu", io->workspace_bytes, need);
  info_carve(io->workspace, io->B, io->N, io->D, nt, &w);
  InfoNceArgs a{w.logits, w.ld, io->B, io->N, io->row_offset, nt};
  const float w_img = io->w_img, w_txt = nt == 2 ? io->w_txt : 0.f;

  if (phase_mask & EEGB200_PHASE_A) {
    EEG_TRY(pad_copy(io->eeg, io->D, io->B, io->D, w.E_r, io->D, io->B, tf32_rounding(), 1.f, s));
    // targets that already sit in the workspace's own slots (eegb200_infonce_target_offset: the data-parallel step
    // all-gathers the ranks' eegb200_tf32_round-ed blocks straight into them) are used in place
    if (io->tgt_img != w.T_r)
      EEG_TRY(pad_copy(io->tgt_img, io->D, io->N, io->D, w.T_r, io->D, io->N, tf32_rounding(), 1.f, s));
    if (nt == 2 && io->tgt_txt != w.T_r + (size_t)io->N * io->D)
      EEG_TRY(pad_copy(io->tgt_txt, io->D, io->N, io->D, w.T_r + (size_t)io->N * io->D, io->D, io->N, tf32_rounding(), 1.f, s));
    GemmArgs g;
    g.M = io->B; g.N = w.ncol; g.K = io->D;
    g.A = {w.E_r, io->D, 0};
    g.B = {w.T_r, io->D, 0};
    g.epi.C = w.logits; g.epi.ldc = w.ld;
    g.epi.alpha_dev = io->logit_scale;
    EEG_TRY(gemm_launch(g, s));
    EEG_TRY(infonce_row_lse(a, w.row_lse, w.diag, s));
    EEG_TRY(infonce_col_partial(a, w.part_max, w.part_sum, s));
    EEG_TRY(infonce_col_reduce(w.part_max, w.part_sum, w.chunks, (size_t)w.ncol, w.ncol, io->col_stats,
                               io->col_stats + w.ncol, nullptr, s));
  }
  if (phase_mask & EEGB200_PHASE_B) {
    EEG_REQUIRE(io->loss != nullptr, "infonce: null loss output");
    const float* parts = io->col_parts ? io->col_parts : io->col_stats;
    const int n_parts = io->col_parts ? io->n_parts : 1;
    EEG_REQUIRE(n_parts >= 1, "infonce: n_parts must be >= 1");
    EEG_TRY(infonce_col_reduce(parts, parts + w.ncol, n_parts, (size_t)2 * w.ncol, w.ncol, nullptr, nullptr, w.col_lse, s));
    EEG_TRY(infonce_loss(a, w.row_lse, w.diag, w.col_lse, w_img, w_txt, io->loss, s));
    if (io->d_eeg) {
      float* ds = io->d_logit_scale;
      EEG_REQUIRE(ds != nullptr, "infonce: d_logit_scale must be provided together with d_eeg");
      EEG_TRY(infonce_grad(a, w.logits, w.row_lse, w.col_lse, w_img, w_txt, io->logit_scale, ds, io->grad_out, s));
      GemmArgs g;                                  // dE = s * G . Tcat
      g.M = io->B; g.N = io->D; g.K = w.ncol;
      g.A = {w.logits, w.ld, 0};
      g.B = {w.T_r, io->D, 1};
      g.epi.C = io->d_eeg; g.epi.ldc = io->D;
      g.epi.alpha_dev = io->logit_scale;
      if (w.ncol >= 1024) {                        // two K splits on 128-wide tiles: one wave of CTAs (34 -> 25 us at N = 1024)
        EEG_CUDA_OK(cudaMemsetAsync(io->d_eeg, 0, (size_t)io->B * io->D * sizeof(float), s));
        g.epi.store_mode = EPI_ATOMIC;
        g.split_k = 2;
        g.tile_n = 128;
      }
      EEG_TRY(gemm_launch(g, s));
    }
  }
  return 0;
}

int eegb200_retrieval(const float* eeg, const float* gallery, int Q, int G, int D, const float* logit_scale,
                      float* logits_ws, int ld, void* round_ws, const int32_t* sel, int k, float* sel_ws,
                      const int64_t* labels, int* correct, int64_t* top1, int32_t* top5, void* stream) {
  EEG_REQUIRE(eeg && gallery && logit_scale && logits_ws && round_ws, "retrieval: null pointer");
  EEG_REQUIRE(Q > 0 && G > 0 && D > 0 && (D & 3) == 0 && ld >= G && (ld & 3) == 0, "retrieval: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  // 3xTF32 (hi.hi + hi.lo + lo.hi) as one GEMM over K = 3*D: near-fp32 scores, so retrieval ranks match the fp32
  // reference except for exact ties
  float* a3 = reinterpret_cast<float*>(round_ws);
  float* b3 = a3 + (size_t)Q * 3 * D;
  EEG_TRY(split3_tf32(eeg, a3, Q, D, 0, s));
  EEG_TRY(split3_tf32(gallery, b3, G, D, 1, s));
  {
    GemmArgs g;
    g.M = Q; g.N = G; g.K = 3 * D;
    g.A = {a3, 3 * D, 0};
    g.B = {b3, 3 * D, 0};
    g.epi.C = logits_ws; g.epi.ldc = ld;
    g.epi.alpha_dev = logit_scale;
    if (G > 128) {
      // 256-wide tiles with split-K sized to one wave of CTAs: 37 us instead of 89 us for 1024 x 1654 x 3072
      // (64-wide tiles re-read the query rows 26 times; tools/gemm_sweep.py)
      const int tiles = cdiv(Q, 128) * cdiv(G, 256);
      int split = 148 / tiles;
      if (split > 4) split = 4;
      if (split < 1) split = 1;
      g.tile_n = 256;
      if (split > 1) {
        EEG_CUDA_OK(cudaMemsetAsync(logits_ws, 0, (size_t)Q * ld * sizeof(float), s));
        g.epi.store_mode = EPI_ATOMIC;
        g.split_k = split;
      }
    }
    EEG_TRY(gemm_launch(g, s));
  }
  const float* scores = logits_ws;
  int cols = G, sld = ld;
  if (sel) {
    EEG_REQUIRE(k > 0 && sel_ws, "retrieval: candidate lists need k > 0 and sel_ws");
    EEG_TRY(gather_cols(logits_ws, ld, sel, Q, k, sel_ws, s));
    scores = sel_ws; cols = k; sld = k;
  }
  if (top1 || (labels && correct))
    EEG_TRY(argmax_count(scores, sld, Q, cols, reinterpret_cast<const long long*>(labels), correct,
                         reinterpret_cast<long long*>(top1), s));
  if (top5) EEG_TRY(topk5(scores, sld, Q, cols, top5, s));
  return 0;
}

int eegb200_mse(const float* eeg, const float* tgt, int B, int D, long long n_total_rows, float weight, float grad_out,
                float* loss, float* loss_term, float* d_eeg, void* stream) {
  EEG_REQUIRE(eeg && tgt && B > 0 && D > 0 && (D & 3) == 0 && n_total_rows > 0, "mse: bad arguments");
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  EEG_REQUIRE(al(eeg) && al(tgt) && al(d_eeg), "mse: pointers must be 16-byte aligned");
  return mse_loss(eeg, tgt, B, D, n_total_rows, weight, grad_out, loss, loss_term, d_eeg, (cudaStream_t)stream);
}

int eegb200_l2norm_forward(const float* x, float* y, float* norms, int rows, int D, void* stream) {
  EEG_REQUIRE(x && y && norms && rows > 0 && D > 0 && (D & 3) == 0, "l2norm_forward: bad arguments");
  return l2norm_fwd(x, y, norms, rows, D, (cudaStream_t)stream);
}
int eegb200_l2norm_backward(const float* y, const float* norms, const float* dy, float* dx, int rows, int D, void* stream) {
  EEG_REQUIRE(y && norms && dy && dx && rows > 0 && D > 0 && (D & 3) == 0, "l2norm_backward: bad arguments");
  return l2norm_bwd(y, norms, dy, dx, rows, D, (cudaStream_t)stream);
}

int eegb200_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                       float eps, float weight_decay, int step, void* stream) {
  EEG_REQUIRE(p && g && m && v && n > 0, "adamw: bad arguments");
  return adamw_step(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, (cudaStream_t)stream);
}

int eegb200_adamw_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                           float eps, float weight_decay, const long long* step_dev, void* stream) {
  EEG_REQUIRE(p && g && m && v && n > 0 && step_dev, "adamw_dev: bad arguments");
  return adamw_step_dev(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step_dev, (cudaStream_t)stream);
}

}  // extern "C"

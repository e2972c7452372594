// Internal launchers of the non-GEMM kernels (one .cu per family).
#pragma once
#include "common.cuh"
#include "gemm.h"

namespace eegb200 {

// ---- fixed ATM-S geometry (Retrieval/ATMS_retrieval.py:44-59, 97-167) ----
constexpr int N_CH = 63;       // EEG channels
constexpr int N_T = 250;       // time points = d_model
constexpr int N_TOK = 64;      // subject token + 63 channel tokens
constexpr int D_PAD = 256;     // padded d_model
constexpr int N_HEAD = 4;
constexpr int D_HEAD = 62;
constexpr int D_HEAD_PAD = 64;
constexpr int D_FF = 256;
constexpr int N_FILT = 40;
constexpr int K_TEMP = 25;
constexpr int K_POOL = 51;
constexpr int S_POOL = 5;
constexpr int N_POOL = 36;
constexpr int N_PSUM = 200;    // pooled prefix positions: 5*35+24+1
constexpr int K_SPAT = N_CH * N_FILT;   // 2520 = reduction length of the spatial conv, index r*40 + k1
constexpr int D_FEAT = 1440;
constexpr int D_OUT = 1024;
constexpr float BN_MOMENTUM = 0.1f;
constexpr float NORM_EPS = 1e-5f;

// dropout sites (ids are part of the Philox counter; masks are indexed in the padded layouts below)
enum DropSite : uint32_t {
  SITE_EMBED = 1,   // [B*64, 256]
  SITE_ATTN = 2,    // [(b*4+h)*64 + i, 64]
  SITE_RES1 = 3,    // [B*64, 256]
  SITE_FFN1 = 4,    // [B*64, 256]
  SITE_FFN2 = 5,    // [B*64, 256]
  SITE_CONV = 6,    // [B*36, 40]
  SITE_PROJ = 7,    // [B, 1024]
};

// ---- rowwise.cu ----
int pad_input(const float* x, float* xp, int B, cudaStream_t s);
int subject_token(const long long* ids, const float* table, const float* shared_tok, int n_subj, int* flag, float* h0,
                  int B, DropoutCfg drop, int round_tf, cudaStream_t s);
int subject_token_bwd(const long long* ids, const int* flag, const float* dh0, float* dtable, float* dshared, int B,
                      DropoutCfg drop, cudaStream_t s);
int layernorm_fwd(const float* x, int ld, int rows, int D, const float* g1, const float* b1, float* stats1,
                  const float* g2, const float* b2, float* stats2, float* y, int ld_out, int round_tf, cudaStream_t s);
int layernorm_bwd(const float* dy, int ld_dy, const float* x, int ld, int rows, int D, const float* g1, const float* b1,
                  const float* stats1, const float* g2, const float* stats2, float* dx, int ld_dx, float* dg1,
                  float* db1, float* dg2, float* db2, int round_tf, cudaStream_t s);
// token-stream form (every ld = 256): dx, plus optionally drop_out = tf32?(dropout(dx)) and colsum_out += colsum(drop_out)
int layernorm_bwd_tok(const float* dy, const float* x, int rows, int D, const float* g1, const float* b1,
                      const float* stats1, const float* g2, const float* stats2, float* dx, float* dg1, float* db1,
                      float* dg2, float* db2, float* drop_out, DropoutCfg cfg, int drop_round, float* colsum_out,
                      cudaStream_t s);
int colsum(const float* x, int ld, int rows, int cols, float* out, int row_mod, int row_skip, cudaStream_t s);
int dropout_mask(DropoutCfg cfg, int rows, int cols, int ld, float* out, cudaStream_t s);
// dst[r*ld+c] = tf32?(keep(r*ld+c) ? src*scale : 0) over [rows, ld]
int dropout_apply(const float* src, float* dst, int rows, int ld, DropoutCfg cfg, int round_tf, cudaStream_t s);
int dropout_apply_colsum(const float* src, float* dst, int rows, int ld, DropoutCfg cfg, int round_tf, float* colsum_out,
                         int cs_cols, int row_mod, int row_skip, cudaStream_t s);
int l2norm_fwd(const float* x, float* y, float* norms, int rows, int D, cudaStream_t s);
int l2norm_bwd(const float* y, const float* norms, const float* dy, float* dx, int rows, int D, cudaStream_t s);
int pad_copy(const float* src, int ld_src, int rows, int cols, float* dst, int ld_dst, int rows_dst, int round_tf,
             float scale, cudaStream_t s);

// ---- attention.cu ----  qkv [B*64, 768] (Q|K|V, head h at columns h*64..h*64+61), o [B*64, 256]
int attention_fwd(const float* qkv, float* o, int B, DropoutCfg drop, cudaStream_t s);
int attention_bwd(const float* qkv, const float* d_o, float* dqkv, int B, DropoutCfg drop, cudaStream_t s);
// attention_tc.cu: forward on tcgen05 (two samples of one head per 128-row UMMA tile); needs TF32-rounded qkv
int attention_fwd_tc(const float* qkv, float* o, int B, DropoutCfg drop, cudaStream_t s);
int attention_tc_enabled();     // default on, EEGB200_ATTN_TC=0 disables

// ---- convstack.cu ----
struct BnState {          // one BatchNorm2d(40); all device pointers
  double* sums;           // [2][40] batch sum / sum of squares (train)
  float* mean_rstd;       // [2][40] statistics used for normalisation this pass
  const float* gamma; const float* beta;
  float* running_mean; float* running_var;   // updated in train mode (momentum 0.1, unbiased variance)
};
int conv_temporal_fwd(const float* x3, const float* wt, const float* bt, float* y1, double* sums, int B, cudaStream_t s);
int bn_finalize(BnState bn, long long count, int train, int update_running, cudaStream_t s);
int bn_elu_apply(const float* y, const float* mean_rstd, const float* gamma, const float* beta, float* a, long long n,
                 int round_tf, cudaStream_t s);
int colstats(const float* y, int ld, int rows, int cols, double* sums, cudaStream_t s);
int conv_head_fwd(const float* y2, const float* mean_rstd, const float* gamma, const float* beta, const float* wc,
                  const float* bc, float* feat, int B, DropoutCfg drop, cudaStream_t s);
int conv_head_bwd(const float* dfeat, const float* y2, const float* mean_rstd, const float* gamma, const float* beta,
                  const float* wc, float* dz2, float* dwc, float* dbc, double* bwd_sums, int B, DropoutCfg drop,
                  cudaStream_t s);
// dy = gamma*rstd*(dz - mean(dz) - yhat*mean(dz*yhat)) elementwise over [rows, 40]; also dgamma/dbeta
int bn_bwd_apply(const float* dz, const float* y, const float* mean_rstd, const float* gamma, const double* bwd_sums,
                 long long count, float* dy, float* dgamma, float* dbeta, long long n, int round_tf, float gscale,
                 cudaStream_t s);
// dz1 = da1 * ELU'(bn1(y1)) in place + reduction sums for the BN1 backward
int bn1_bwd_reduce(float* da1, const float* y1, const float* mean_rstd, const float* gamma, const float* beta,
                   double* bwd_sums, long long n, cudaStream_t s);
int conv_temporal_bwd(const float* dz1, const float* y1, const float* x3, const float* wt, const float* mean_rstd,
                      const float* gamma, const double* bwd_sums, long long count, float* dx3, float* dwt, float* dbt,
                      float* dgamma, float* dbeta, int B, float gscale, cudaStream_t s);

// ---- conv_tc.cu: the conv stack on tcgen05, forward (F1 / F2) and backward (B1 / B2); y1 / a1 / d a1 stay on chip ----
int conv_tc_enabled();          // default on, EEGB200_CONV_TC=0 selects the unfused round-1 kernels
size_t conv_tc_ws_floats();     // packed spatial weights: forward [63][48][64] + backward [63][48][32] + tails [16][48][32]
int conv_tc_stats(const float* x3, const float* wt, const float* bt, double* sums, int B, cudaStream_t s);
int conv_tc_apply(const float* x3, const float* wt, const float* bt, const float* mean_rstd, const float* gamma,
                  const float* beta, const float* ws, const float* bs, float* ws_packed, float* y1, float* a1, float* y2,
                  int B, cudaStream_t s);
int conv_tc_bwd_stats(const float* x3, const float* wt, const float* bt, const float* mean_rstd, const float* gamma,
                      const float* beta, const float* ws, const float* dy2, float* ws_packed, double* bsums, float* dws,
                      int B, cudaStream_t s);
int conv_tc_bwd_apply(const float* x3, const float* wt, const float* bt, const float* mean_rstd, const float* gamma,
                      const float* beta, const float* dy2, float* ws_packed, double* bsums, long long count, float gscale,
                      float* dx3, float* dwt, float* dbt, float* dgamma, float* dbeta, int B, cudaStream_t s);
int debug_stores();             // eegb200_set_debug_stores: keep y1 / a1 in the workspace for the stage checks

// ---- loss.cu ----
struct InfoNceArgs {
  const float* logits;   // [B, ld] = s * E * Tcat^T, Tcat = [img ; txt] (2N columns)
  int ld;
  int B;                 // local rows
  int N;                 // global batch (columns per target)
  int row_offset;        // rank * B : global index of local row 0
  int nt;                // number of targets concatenated along the columns (1 or 2)
};
int infonce_row_lse(const InfoNceArgs& a, float* row_lse /*[2][B]*/, float* diag /*[2][B]*/, cudaStream_t s);
int infonce_col_chunks(int B);   // number of row chunks (= partial sets) infonce_col_partial writes
int infonce_col_partial(const InfoNceArgs& a, float* part_max /*[chunks][2N]*/, float* part_sum, cudaStream_t s);
int infonce_col_reduce(const float* part_max, const float* part_sum, int n_parts, size_t stride, int n, float* out_max,
                       float* out_sum, float* out_lse, cudaStream_t s);
int gather_cols(const float* logits, int ld, const int* sel, int Q, int k, float* out, cudaStream_t s);
int split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t s);
int split3_tf32(const float* x, float* out, long long rows, int D, int is_b, cudaStream_t s);
// loss_partial[0] += alpha*..., per rank partial of the global loss (sum over ranks = loss)
int infonce_loss(const InfoNceArgs& a, const float* row_lse, const float* diag, const float* col_lse, float w_img,
                 float w_txt, float* loss_out, cudaStream_t s);
// G (in place over logits) and d(logit_scale) partial
int infonce_grad(const InfoNceArgs& a, float* logits_inout, const float* row_lse, const float* col_lse, float w_img,
                 float w_txt, const float* logit_scale_dev, float* dscale, float grad_out_scale, cudaStream_t s);
int argmax_count(const float* logits, int ld, int rows, int cols, const long long* labels, int* correct,
                 long long* pred_out, cudaStream_t s);
int topk5(const float* logits, int ld, int rows, int cols, int* top5_out /*[rows][5]*/, cudaStream_t s);
// weight * MSE(eeg, tgt) share of this rank (mean over n_total*D elements): loss / loss_term += (either may be null),
// d_eeg += weight*grad_out * d MSE / d eeg (null: loss only)
int mse_loss(const float* eeg, const float* tgt, int B, int D, long long n_total, float weight, float grad_out, float* loss,
             float* loss_term, float* d_eeg, cudaStream_t s);

// ---- optim.cu ----
int adamw_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                   float wd, const long long* step_dev, cudaStream_t s);
int adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
               float wd, int step, cudaStream_t s);

}  // namespace eegb200

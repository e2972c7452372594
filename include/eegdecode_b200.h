/* eegdecode_b200 -- C ABI of the B200-native ATM-S contrastive hot path.
 *
 * The reference (dongyangli-del/EEG_Image_decode) is pure Python and has no FFI: its "plugin"
 * surface for this path is the Python classes/functions ATMS, ClipLoss, train_model and
 * evaluate_model (Retrieval/ATMS_retrieval.py:171-362, models/loss.py:78-141).  The Python package
 * eeg_image_decode_b200 mirrors those signatures and binds THIS library with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name says host; fp32 row-major, contiguous
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream)
 *   - functions return 0 on success; on failure a message is available from eegb200_last_error()
 *   - no allocation happens inside the library: callers pass a workspace whose size they query first
 *   - there is no CPU implementation behind any of these entry points
 */
#ifndef EEGDECODE_B200_H
#define EEGDECODE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EEGB200_ABI_VERSION 2

int eegb200_abi_version(void);
const char* eegb200_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long eegb200_launch_count(void);
/* 0 = tcgen05 TF32 tensor-core GEMMs (product path), 1 = exact-fp32 SIMT verification GEMM (tests only) */
int eegb200_set_gemm_backend(int backend);
int eegb200_get_gemm_backend(void);
/* 1: the fused conv-stack kernels also store their on-chip intermediates (y1, a1: 363 KB per sample each) into the
 * workspace so that eegb200_atms_ws_tensor("y1" / "a1") can be compared stage by stage (tests only; default 0) */
int eegb200_set_debug_stores(int on);
/* per-kernel CUDA-event profiler (bench.py roofline leg): enable, run steps, then fetch a JSON report
 * {"kernel name": {"ms": total, "n": launches, "flops": algorithmic, "bytes": algorithmic}} */
int eegb200_prof_enable(int on);
int eegb200_prof_report(char* buf, size_t cap);

/* ---- generic fused GEMM: C[M,N] = epilogue(A[M,K] * B[N,K]^T).  Building block of every Linear /
 * 1x1-conv / spatial-conv / logits product on the path (nn.Linear & F.linear call sites:
 * Embed.py:146, SelfAttention_Family.py:199-213, Transformer_EncDec.py:48-49, ATMS_retrieval.py:106,160-163,
 * loss.py:122-123).  Exposed for the parity tests. */
typedef struct eegb200_gemm_desc {
  int M, N, K;
  const float* A; int lda; int a_mn_major;   /* K-major: A[i*lda+k];  MN-major: A[k*lda+i] */
  const float* B; int ldb; int b_mn_major;
  float* C; int ldc;
  float alpha;
  const float* bias; int bias_period; int ld_bias;  /* per column, or table indexed by row % bias_period */
  float* aux_out; int ld_aux;                       /* optional copy of the pre-activation value */
  int act;                                          /* 0 none, 1 exact GELU */
  uint64_t drop_seed; uint32_t drop_site; float drop_p; int drop_ld;   /* dropout after the activation */
  const float* mul_in; int ld_mul;                  /* multiply by GELU'(mul_in[row,col]) */
  const float* resid; int ld_res;                   /* residual added last */
  int round_tf32;                                   /* round the stored value to TF32 (RN) */
  int store_mode;                                   /* 0 store, 1 C += v, 2 atomicAdd (split-K) */
  int split_k;
  int tile_n;                                       /* 0 = heuristic; 64 / 128 / 256 force the N tile (tuning, tests) */
} eegb200_gemm_desc;
int eegb200_gemm(const eegb200_gemm_desc* d, void* stream);

/* keep-mask (1.0 / 0.0) of a dropout site, element index = r*ld + c for r<rows, c<cols; written to
 * out[r*cols + c].  Lets the tests feed the library's own masks to the oracle. */
int eegb200_dropout_mask(uint64_t seed, uint32_t site, float p, int rows, int cols, int ld, float* out, void* stream);


/* ------------------------------------------------------------------------------------------------
 * ATM-S encoder  (replaces ATMS.forward, Retrieval/ATMS_retrieval.py:182-191, and its autograd graph)
 * ------------------------------------------------------------------------------------------------ */
/* trainable tensors, in the reference state_dict layouts */
enum eegb200_param {
  EEGB200_P_VALUE_W = 0,  /* encoder.enc_embedding.value_embedding.weight            [250,250] */
  EEGB200_P_VALUE_B,      /* ...value_embedding.bias                                  [250]     */
  EEGB200_P_SUBJ_TABLE,   /* ...subject_embedding.subject_embedding.weight            [n_subj,250] */
  EEGB200_P_SUBJ_SHARED,  /* ...subject_embedding.shared_embedding                    [1,250]   */
  EEGB200_P_WQ, EEGB200_P_BQ,   /* attention.query_projection  [248,250],[248] */
  EEGB200_P_WK, EEGB200_P_BK,   /* attention.key_projection */
  EEGB200_P_WV, EEGB200_P_BV,   /* attention.value_projection */
  EEGB200_P_WO, EEGB200_P_BO,   /* attention.out_projection    [250,248],[250] */
  EEGB200_P_W1, EEGB200_P_B1,   /* attn_layers.0.conv1         [256,250,1],[256] */
  EEGB200_P_W2, EEGB200_P_B2,   /* attn_layers.0.conv2         [250,256,1],[250] */
  EEGB200_P_LN1_G, EEGB200_P_LN1_B,   /* attn_layers.0.norm1 */
  EEGB200_P_LN2_G, EEGB200_P_LN2_B,   /* attn_layers.0.norm2 */
  EEGB200_P_LNF_G, EEGB200_P_LNF_B,   /* encoder.encoder.norm */
  EEGB200_P_WT, EEGB200_P_BT,         /* enc_eeg.0.tsconv.0  [40,1,1,25],[40] */
  EEGB200_P_BN1_G, EEGB200_P_BN1_B,   /* enc_eeg.0.tsconv.2 */
  EEGB200_P_WS, EEGB200_P_BS,         /* enc_eeg.0.tsconv.4  [40,40,63,1],[40] */
  EEGB200_P_BN2_G, EEGB200_P_BN2_B,   /* enc_eeg.0.tsconv.5 */
  EEGB200_P_WC, EEGB200_P_BC,         /* enc_eeg.0.projection.0 [40,40,1,1],[40] */
  EEGB200_P_WP1, EEGB200_P_BP1,       /* proj_eeg.0          [1024,1440],[1024] */
  EEGB200_P_WP2, EEGB200_P_BP2,       /* proj_eeg.1.fn.1     [1024,1024],[1024] */
  EEGB200_P_LNP_G, EEGB200_P_LNP_B,   /* proj_eeg.2 */
  EEGB200_P_COUNT
};
/* non-trainable buffers */
enum eegb200_buffer {
  EEGB200_BUF_PE = 0,     /* position_embedding.pe [1,5000,250] (rows 0..62 used) */
  EEGB200_BUF_BN1_RM, EEGB200_BUF_BN1_RV,   /* tsconv.2.running_mean / running_var [40] */
  EEGB200_BUF_BN2_RM, EEGB200_BUF_BN2_RV,   /* tsconv.5.* */
  EEGB200_BUF_COUNT
};
/* dropout sites; dropout_p[site] overrides the reference probabilities (0.25 x5, 0.5, 0.5) */
enum eegb200_dropout_site {
  EEGB200_SITE_EMBED = 1, EEGB200_SITE_ATTN = 2, EEGB200_SITE_RES1 = 3, EEGB200_SITE_FFN1 = 4,
  EEGB200_SITE_FFN2 = 5, EEGB200_SITE_CONV = 6, EEGB200_SITE_PROJ = 7, EEGB200_SITE_COUNT = 8
};
/* phases: the forward (backward) is cut where train-mode BatchNorm needs batch statistics, so that a
 * data-parallel caller can all-reduce the 80 doubles between phases (SyncBN); single GPU: mask 7. */
#define EEGB200_PHASE_A 1
#define EEGB200_PHASE_B 2
#define EEGB200_PHASE_C 4
#define EEGB200_PHASE_ALL 7
/* bits 8..15 of the phase mask: number of ranks whose BatchNorm sums were all-reduced into the workspace
 * (0/1 = local statistics).  The element count used for mean/var becomes world*B*... */
#define EEGB200_SYNCBN_WORLD(w) (((w) & 0xFF) << 8)

typedef struct eegb200_atms_io {
  const float* const* params;    /* [EEGB200_P_COUNT] device pointers */
  float* const* buffers;         /* [EEGB200_BUF_COUNT] */
  const float* x;                /* [B,63,250] */
  const int64_t* subject_ids;    /* [B] */
  int B;
  int n_subjects;                /* rows of the subject table (10) */
  int train;                     /* nn.Module.train(): batch-stat BatchNorm + dropout */
  int update_running_stats;
  uint64_t seed;                 /* dropout seed of this step (the backward must get the same value) */
  const float* dropout_p;        /* host float[EEGB200_SITE_COUNT] or NULL for the reference values */
  void* workspace;               /* >= eegb200_atms_workspace_bytes(B); holds the saved activations */
  size_t workspace_bytes;
  float* out;                    /* [B,1024] */
  const uint64_t* seed_offset_dev; /* optional DEVICE counter mixed into `seed` when the kernels run (CUDA-graph replay
                                      draws new dropout masks without re-capturing); NULL -> `seed` alone */
  /* ---- joint-subject variant (ABI v2): per-subject value embeddings, DataEmbedding(joint_train=True)
   * (models/subject_layers/Embed.py:128-130, 144; model built at Retrieval/ATMS_retrieval_joint_train.py:173-176).
   * joint_value_w == NULL selects the single shared embedding params[EEGB200_P_VALUE_W / _B] (all fields below
   * ignored).  Otherwise the trials must be ordered so that trials of one subject are contiguous: group g covers
   * trials [group_offsets[g], group_offsets[g+1]) and uses value embedding group_subject[g] -- one grouped GEMM per
   * subject instead of the reference's per-trial Python loop.  params[EEGB200_P_VALUE_W / _B] must still be
   * non-NULL (any valid pointer) and are not read; grads of those two slots are not written. */
  const float* const* joint_value_w;   /* [n_subjects] device pointers, each [250,250] */
  const float* const* joint_value_b;   /* [n_subjects] device pointers, each [250] */
  float* const* joint_value_dw;        /* backward: += gradients, same indexing (subjects absent from the batch untouched) */
  float* const* joint_value_db;
  const int32_t* group_offsets;        /* HOST int32[n_groups+1], group_offsets[0] = 0, group_offsets[n_groups] = B */
  const int32_t* group_subject;        /* HOST int32[n_groups], each in [0, n_subjects), pairwise distinct */
  int n_groups;
} eegb200_atms_io;

size_t eegb200_atms_workspace_bytes(int B);
int eegb200_atms_forward(const eegb200_atms_io* io, int phase_mask, void* stream);
/* grads[i] += d loss / d params[i] (caller zeroes; NULL entries are skipped where legal).
 * Needs the workspace of the matching forward (train mode). */
int eegb200_atms_backward(const eegb200_atms_io* io, const float* d_out, float* const* grads, int phase_mask, void* stream);
/* intermediates inside the workspace, for stage-level parity tests and SyncBN exchange.
 * name: "h0","qkv","attn_o","x1","ffn_u","x3","y1","a1","y2","feat","z1","z2",
 *       "bn1_sums","bn2_sums","bn1_bwd_sums","bn2_bwd_sums" (the last four are double[80]). */
int eegb200_atms_ws_tensor(void* workspace, int B, const char* name, void** ptr, int* rows, int* cols, int* ld);

/* ------------------------------------------------------------------------------------------------
 * Contrastive loss (replaces ClipLoss.forward, models/loss.py:100-141, called twice per step at
 * ATMS_retrieval.py:229-230 and mixed 0.99/0.01 at :234).  Row block of this rank against the
 * global target matrix; never builds the N x N problem more than B x N per rank.
 * Phase A: logits + row/column statistics.  Between the phases a multi-GPU caller all-gathers
 * `col_stats` ([2][n_targets*N]: running max, sum) from every rank.  Phase B: loss share + grads.
 * ------------------------------------------------------------------------------------------------ */
typedef struct eegb200_infonce_io {
  const float* eeg;            /* [B,D] local embeddings (first ClipLoss argument) */
  const float* tgt_img;        /* [N,D] global targets */
  const float* tgt_txt;        /* [N,D] or NULL (single ClipLoss) */
  int B, N, D, row_offset;     /* row_offset = rank*B */
  const float* logit_scale;    /* device scalar, used RAW like the reference (no exp) */
  float w_img, w_txt;          /* 0.99 / 0.01 */
  float grad_out;              /* upstream d(total)/d(loss) */
  void* workspace; size_t workspace_bytes;   /* >= eegb200_infonce_workspace_bytes */
  float* col_stats;            /* out (phase A): [2][nt*N] local (max,sum) per column */
  const float* col_parts;      /* in (phase B): [n_parts][2][nt*N]; NULL -> use col_stats, n_parts=1 */
  int n_parts;
  float* loss;                 /* out (phase B): [3] = mix, img, txt shares of this rank (sum over ranks = loss) */
  float* d_eeg;                /* out (phase B): [B,D], may be NULL */
  float* d_logit_scale;        /* += (phase B), device scalar, may be NULL */
} eegb200_infonce_io;
size_t eegb200_infonce_workspace_bytes(int B, int N, int D, int n_targets);
int eegb200_infonce(const eegb200_infonce_io* io, int phase_mask, void* stream);
/* Zero-copy targets for the data-parallel step: byte offset, inside an eegb200_infonce workspace of this shape, of the
 * [n_targets][N, D] operand the logits GEMM reads.  A caller may all-gather the ranks' target blocks straight into it
 * (each block rounded first with eegb200_tf32_round: the tensor core truncates, the loss wants round-to-nearest) and pass
 * those addresses as tgt_img / tgt_txt: eegb200_infonce then skips its own round-and-copy of the 2 x N x D targets. */
size_t eegb200_infonce_target_offset(int B, int N, int D, int n_targets);
int eegb200_tf32_round(const float* src, float* dst, int rows, int D, void* stream);

/* Regression term of the reconstruction-training variant (replaces nn.MSELoss()(eeg_features, img_features),
 * Generation/ATMS_reconstruction.py:201, 227-228 and :264, 285-286; the step loss there is
 * alpha*10*MSE + (1-alpha)*10*ClipLoss(img)).  MSE is the mean over all n_total_rows*D elements of the (global) batch;
 * this rank's B rows contribute
 *     *loss, *loss_term += weight * sum_{b,d} (eeg-tgt)^2 / (n_total_rows*D)          (either may be NULL)
 *     d_eeg[b,d]        += weight * grad_out * 2 (eeg-tgt)[b,d] / (n_total_rows*D)     (NULL: loss only)
 * so it composes with eegb200_infonce, whose phase B writes loss[0] and d_eeg first.  n_total_rows = b with B = n*b rows
 * gives the SUM of the n per-batch means of batches of b rows (evaluate_model averages per-batch losses). */
int eegb200_mse(const float* eeg, const float* tgt, int B, int D, long long n_total_rows, float weight, float grad_out,
                float* loss, float* loss_term, float* d_eeg, void* stream);

/* SyncBatchNorm statistics exchange of the data-parallel step without NCCL: one-shot all-reduce (sum, fp64, n <= 256) over
 * NVLink peer memory.  peer_buffers[r] = rank r's symmetric buffer of eegb200_peer_sum_buffer_bytes() bytes as mapped in
 * THIS process (zero-initialised once, e.g. torch.distributed._symmetric_memory.empty + rendezvous -> buffer_ptrs);
 * seq_dev = this rank's call counter (device uint64, starts at 0, advanced by the kernel: CUDA-graph replayable);
 * *error_flag_dev is set to 1 if a peer never arrived (the kernel gives up after a few seconds instead of hanging).
 * Every rank must make the same sequence of calls.  data (local, device) is summed in place, identically on all ranks.
 * Replaces torch.distributed.all_reduce on the 2 x 40 BatchNorm sums (train.py StepEngine). */
size_t eegb200_peer_sum_buffer_bytes(void);
int eegb200_peer_sum_f64(double* data, int n, const void* const* peer_buffers, int rank, int world,
                         unsigned long long* seq_dev, int* error_flag_dev, void* stream);

/* Optional L2 normalisation of the EEG embedding before the logits (BASELINE.json north_star wording).  The reference
 * does NOT normalise (ATMS.forward, Retrieval/ATMS_retrieval.py:182-191): ATMS(normalize=False) is the default and the
 * parity path.  y = x / max(|x|_2, 1e-12) row-wise (torch.nn.functional.normalize), norms[row] = the denominator;
 * backward dx = (dy - y (y . dy)) / norm.  x, y, dy, dx: [rows, D] fp32, D % 4 == 0, 16-byte aligned. */
int eegb200_l2norm_forward(const float* x, float* y, float* norms, int rows, int D, void* stream);
int eegb200_l2norm_backward(const float* y, const float* norms, const float* dy, float* dx, int rows, int D, void* stream);

/* scores = logit_scale * eeg @ gallery^T into logits_ws [Q, ld >= G rounded up to 4]; optional argmax
 * count against labels (train accuracy, ATMS_retrieval.py:241-250), top-1 / top-5 indices
 * (evaluate_model, :306-320).  `sel` (int32 [Q,k] or NULL) restricts each query to its own candidate
 * list: results are positions into that list. */
int eegb200_retrieval(const float* eeg, const float* gallery, int Q, int G, int D, const float* logit_scale,
                      float* logits_ws, int ld, void* round_ws /* 3*(Q+G)*D floats */,
                      const int32_t* sel, int k, float* sel_ws /* [Q,k] */,
                      const int64_t* labels, int* correct, int64_t* top1, int32_t* top5, void* stream);

/* torch.optim.AdamW semantics on a flat arena (ATMS_retrieval.py:548, :237) */
int eegb200_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                       float eps, float weight_decay, int step, void* stream);
/* same, with the step number read from DEVICE memory when the kernel runs (CUDA-graph friendly) */
int eegb200_adamw_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                           float eps, float weight_decay, const long long* step_dev, void* stream);

/* Debug probe (tools/gpu_tma_layout_probe.py): shared-memory image (4096 floats, verbatim) that TMA writes for the first
 * [128 rows x 32 k] k-block of a GEMM operand, K-major (src[row*ld + k], SWIZZLE_128B) or MN-major (src[k*ld + row],
 * SWIZZLE_128B_ATOM_32B).  Used to validate the layouts thread-written UMMA operands must follow. */
int eegb200_debug_tma_tile(const float* src, int ld, int mn_major, float* out, void* stream);
/* debug: out128[lane] = column 0 of TMEM lane `lane` after an M = 64 UMMA wrote D[r][0] = r + 1 (lanes it did not touch: -1) */
int eegb200_debug_umma_m64(float* out128, void* stream);
/* debug: cycles of a chain of n tcgen05.mma (kind::tf32, K = 8) of shape M x N: out2[0] = issue..completion, out2[1] = issue */
int eegb200_debug_umma_cost(int M, int N, int mn_major, int n, int background, long long* out2, void* stream);
/* debug: one UMMA chain D[128][N] = A . B^T over host-prepared shared-memory images (32 KB each) with caller-given
 * descriptor fields cfg12 = {a_layout, a_lbo, a_sbo, a_kadv, a_major, b_layout, b_lbo, b_sbo, b_kadv, b_major, N, nk};
 * out = [128][N].  tools/gpu_dual_layout_probe.py */
int eegb200_debug_umma_generic(const float* a_img, const float* b_img, const int* cfg12, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EEGDECODE_B200_H */

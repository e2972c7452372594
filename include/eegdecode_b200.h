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

#define EEGB200_ABI_VERSION 1

int eegb200_abi_version(void);
const char* eegb200_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long eegb200_launch_count(void);
/* 0 = tcgen05 TF32 tensor-core GEMMs (product path), 1 = exact-fp32 SIMT verification GEMM (tests only) */
int eegb200_set_gemm_backend(int backend);
int eegb200_get_gemm_backend(void);

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
} eegb200_gemm_desc;
int eegb200_gemm(const eegb200_gemm_desc* d, void* stream);

/* keep-mask (1.0 / 0.0) of a dropout site, element index = r*ld + c for r<rows, c<cols; written to
 * out[r*cols + c].  Lets the tests feed the library's own masks to the oracle. */
int eegb200_dropout_mask(uint64_t seed, uint32_t site, float p, int rows, int cols, int ld, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EEGDECODE_B200_H */

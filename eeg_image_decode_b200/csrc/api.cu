// C-ABI entry points (include/eegdecode_b200.h).
#include "../../include/eegdecode_b200.h"
#include "gemm.h"

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
  return gemm_launch(g, (cudaStream_t)stream);
}

}  // extern "C"

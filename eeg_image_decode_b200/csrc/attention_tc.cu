// Attention core of the 64-token, 4-head layer (FullAttention.forward, SelfAttention_Family.py:56-75) on the 5th-gen
// tensor cores.  A (sample, head) problem is only 64 x 64 x 64, half a UMMA tile, so one CTA takes the SAME head of two
// consecutive samples and runs them as one block-diagonal 128-row problem:
//
//   S[128,128] = [Q_a; Q_b] . [K_a; K_b]^T      tcgen05.mma kind::tf32, M=128 N=128 K=64, accumulator in TMEM cols 0..127;
//                                               only the two diagonal 64x64 blocks are meaningful
//   P          = dropout(softmax(S_diag / sqrt(62)))   thread = TMEM lane = query row: tcgen05.ld of its 64 scores,
//                                               softmax in registers, TF32-rounded, written to shared memory in the
//                                               SWIZZLE_128B K-major layout (zeros in the off-diagonal key blocks)
//   O[128,64]  = P . [V_a; V_b]                 M=128 N=64 K=128, V as an MN-major operand straight from the QKV matrix
//
// Q, K, V tiles arrive by TMA from the [B*64, 768] QKV matrix (one tensor map for Q and K, an MN-major one for V).  The
// 50 % padding waste of the block-diagonal form is irrelevant at these sizes; what matters is that the 2 x 256 legacy
// mma.sync instructions per warp of the first version (which sit at ~80 % of that pipe's issue rate) become 24 UMMAs.
// qkv must already be TF32-rounded (kind::tf32 truncates): the QKV GEMM epilogue rounds when this path is enabled.
#include "kernels.h"
#include <stdlib.h>

namespace eegb200 {

static constexpr int TC_THREADS = 128;
static constexpr float TC_QK_SCALE = 0.12700012700019050f;   // 1/sqrt(62)
static constexpr uint32_t TILE_KB = 128 * 32 * 4;            // one [128 rows x 32 floats] K-major k-block: 16 KB
static constexpr uint32_t V_KB = 2 * 32 * 32 * 4;            // one V k-block: 2 slabs x [32 keys][32 dims]: 8 KB

__global__ void __launch_bounds__(TC_THREADS, 2)
attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV,
                        float* __restrict__ o, int m_tok, DropoutCfg drop) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* tiles = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* Qs = tiles;                       // 2 k-blocks, 32 KB   } reused as P (4 k-blocks, 64 KB) once S is done
  uint8_t* Ks = tiles + 2 * TILE_KB;         // 2 k-blocks, 32 KB   }
  uint8_t* Vs = tiles + 4 * TILE_KB;         // 4 k-blocks x 8 KB
  uint8_t* Ps = tiles;
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(tiles + 4 * TILE_KB + 4 * V_KB);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;
  uint64_t* bar_o = bar_qk + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 4);

  const int warp = threadIdx.x >> 5;
  const int pair = blockIdx.x >> 2, h = blockIdx.x & 3;
  const int row0 = pair * 128;               // first token row of the two samples

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);              // S: columns 0..127, O: columns 128..191
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 0) {
    // ---- TMA: Q and K (2 k-blocks of 32 head dims each), then V (4 k-blocks of 32 keys) ----
    mbar_arrive_expect_tx(bar_qk, 4 * TILE_KB);
    for (int kb = 0; kb < 2; ++kb) {
      tma_load_2d(&tmQK, bar_qk, Qs + kb * TILE_KB, h * 64 + kb * 32, row0);
      tma_load_2d(&tmQK, bar_qk, Ks + kb * TILE_KB, 256 + h * 64 + kb * 32, row0);
    }
    mbar_arrive_expect_tx(bar_v, 4 * V_KB);
    for (int kb = 0; kb < 4; ++kb) tma_load_3d(&tmV, bar_v, Vs + kb * V_KB, 0, row0 + kb * 32, 16 + 2 * h);
    // ---- S = Q . K^T ----
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    constexpr uint32_t idesc_s = umma_idesc_tf32(128, 128, 0, 0);
    const uint32_t qa = smem_u32(Qs), ka = smem_u32(Ks);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        tc_mma_tf32(tmem_base, umma_smem_desc(qa + kb * TILE_KB + kk * 32, 16, 1024, UMMA_LAYOUT_SW128),
                    umma_smem_desc(ka + kb * TILE_KB + kk * 32, 16, 1024, UMMA_LAYOUT_SW128), idesc_s, (kb | kk) ? 1u : 0u);
    tc_commit(bar_s);
  }

  // ---- softmax of this thread's query row (TMEM lane = threadIdx.x); warps 0,1 = first sample, 2,3 = second ----
  const int half = warp >> 1;                 // which sample of the pair this row belongs to
  const int row = threadIdx.x;                // 0..127 inside the CTA tile
  mbar_wait(bar_s, 0);
  tc_fence_after();
  float p[64];
  tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 64), p);
  tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 64 + 32), p + 32);
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 64; ++j) { p[j] *= TC_QK_SCALE; mx = fmaxf(mx, p[j]); }
  float z = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) { p[j] = __expf(p[j] - mx); z += p[j]; }
  const float inv = 1.f / z;
  if (drop.p > 0.f) {
    // mask element index = ((sample*4 + head)*64 + query)*64 + key, as in the mma.sync / SIMT kernels
    const uint64_t base = ((uint64_t)((pair * 2 + half) * 4 + h) * 64 + (row & 63)) * 64;
#pragma unroll
    for (int j4 = 0; j4 < 16; ++j4) {
      const uint32_t m = dropout_keep4(drop, base + 4 * j4);
#pragma unroll
      for (int q = 0; q < 4; ++q) p[4 * j4 + q] = (m >> q) & 1u ? p[4 * j4 + q] * (inv * drop.scale) : 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 64; ++j) p[j] *= inv;
  }
  // ---- P -> shared memory, K-major SWIZZLE_128B: k-block kb holds keys 32*kb .. 32*kb+31 of the 128-key axis ----
  // (Q and K are dead: every MMA that read them completed before bar_s flipped)
  {
    const uint32_t r8 = row & 7;
    uint8_t* line = Ps + (row >> 3) * 1024 + r8 * 128;
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      const bool mine = (kb >> 1) == half;     // the off-diagonal key blocks of this row are zero
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mine) {
          const int j = (kb & 1) * 32 + c * 4;
          v = make_float4(tf32_rn(p[j]), tf32_rn(p[j + 1]), tf32_rn(p[j + 2]), tf32_rn(p[j + 3]));
        }
        *reinterpret_cast<float4*>(line + kb * TILE_KB + ((c ^ r8) << 4)) = v;
      }
    }
  }
  fence_proxy_async_smem();                   // generic-proxy stores -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    // ---- O = P . V ----
    tc_fence_after();
    mbar_wait(bar_v, 0);
    tc_fence_after();
    constexpr uint32_t idesc_o = umma_idesc_tf32(128, 64, 0, 1);
    const uint32_t pa = smem_u32(Ps), va = smem_u32(Vs);
#pragma unroll
    for (int kb = 0; kb < 4; ++kb)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        tc_mma_tf32(tmem_base + 128, umma_smem_desc(pa + kb * TILE_KB + kk * 32, 16, 1024, UMMA_LAYOUT_SW128),
                    umma_smem_desc(va + kb * V_KB + kk * 1024, 32 * 128, 512, UMMA_LAYOUT_SW128_BASE32B), idesc_o,
                    (kb | kk) ? 1u : 0u);
    tc_commit(bar_o);
  }
  // ---- epilogue: O row -> global (TF32-rounded: it is the A operand of the out-projection GEMM), pad dims zero ----
  mbar_wait(bar_o, 0);
  tc_fence_after();
  float acc[64];
  tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + 128u, acc);
  tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + 160u, acc + 32);
  if (row0 + row < m_tok) {
    float4* dst = reinterpret_cast<float4*>(o + (size_t)(row0 + row) * 256 + h * 64);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float4 v = make_float4(tf32_rn(acc[4 * c]), tf32_rn(acc[4 * c + 1]), tf32_rn(acc[4 * c + 2]), tf32_rn(acc[4 * c + 3]));
      if (c == 15) { v.z = 0.f; v.w = 0.f; }   // head dims 62, 63 are padding
      dst[c] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

int attention_tc_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("EEGB200_ATTN_TC");
    on = (e && e[0] == '0') ? 0 : 1;        // default on; EEGB200_ATTN_TC=0 selects the mma.sync forward (A/B switch)
  }
  return on;
}

int attention_fwd_tc(const float* qkv, float* o, int B, DropoutCfg drop, cudaStream_t s) {
  ProfScope _ps("attention_fwd_tc", s, (double)B * 4 * 4.0 * 64 * 64 * 62, (double)B * 64 * 1024 * 4.0);
  const int m_tok = B * N_TOK;
  CUtensorMap tq, tv;
  int d3 = 0;
  EEG_TRY(gemm_make_tmap(&tq, GemmOperand{qkv, 768, 0}, m_tok, 768, 128, &d3));
  EEG_TRY(gemm_make_tmap(&tv, GemmOperand{qkv, 768, 1}, 768, m_tok, 64, &d3));
  EEG_REQUIRE(d3 == 1, "attention_fwd_tc: the V operand map must be 3-D");
  const size_t smem = 4 * TILE_KB + 4 * V_KB + 64 + 1024;
  static PerDeviceOnce once;
  if (once.first()) {
    EEG_CUDA_OK(cudaFuncSetAttribute(attention_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  attention_fwd_tc_kernel<<<cdiv(B, 2) * N_HEAD, TC_THREADS, smem, s>>>(tq, tv, o, m_tok, drop);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace eegb200

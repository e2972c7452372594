// TF32 GEMM on the 5th-gen tensor cores:  D[M,N] = epilogue(A[M,K] * B[N,K]^T), fp32 in HBM.
//
//   warp 0 : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem stages, mbarrier complete_tx)
//   warp 1 : TMEM allocator + single-thread tcgen05.mma issuer (kind::tf32, fp32 accumulators in TMEM)
//   warps 2..9 : epilogue    (tcgen05.ld 32x32b -> registers -> smem transpose -> fused, coalesced float4 I/O)
//
// Tile 128 x BN x 32(K, = one 128-byte swizzle row of fp32).  K-major operands use the SWIZZLE_128B
// canonical layout; MN-major operands (needed by the weight-gradient GEMMs, whose reduction runs over
// the token dimension) use the 32-bit-only SWIZZLE_128B_ATOM_32B layout (UMMA layout type 1).
// Out-of-range rows / columns / K are zero-filled by TMA, the epilogue masks the stores.
#include "gemm.h"
#include <mutex>
#include <stdlib.h>

namespace eegb200 {

static constexpr int BM = 128;
static constexpr int BK = 32;   // floats per k-block (128 B)
static constexpr int UMMA_K = 8;
static constexpr int GEMM_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)

struct GemmKernelParams {
  int M, N, K;
  int k_blocks_per_split;   // k-blocks handled by one blockIdx.z
  int stages;
  int tile_bytes;           // smem bytes reserved for the pipeline stages / epilogue staging (barriers follow)
  int vec_ok;
  int pair_atomic;          // store_mode ATOMIC with nothing else fused and even ldc: red.v2 pairs
  int a_3d, b_3d;           // MN-major operand loaded with one 3-D TMA box per stage (else one 2-D box per 32-wide slab)
  Epilogue epi;
};

static constexpr int EPI_SLD = 36;   // staging row stride in floats (16-byte aligned, conflict-free float4 phases)

// Epilogue of one 128 x BN accumulator tile by warps 2..9 (TMEM lane quarter = warp % 4, two warps share a quarter and
// split the 32-column chunks).  TMEM -> registers (thread = row) -> per-warp smem transpose -> lanes along columns:
// every global access of the fused epilogue (bias / aux / GELU' input / residual / output) is a coalesced 128-byte row
// segment, and the loads of 4 row groups are in flight before the first is consumed.
template <int BN, uint32_t F>
__device__ __forceinline__ void epilogue_tile(const GemmKernelParams& p, const Epilogue& e, uint32_t tmem_acc, float* stage,
                                              int tile_m, int tile_n, int warp, int lane, bool has_acc, float* sred) {
  const int q = warp & 3;
  const int half = (warp - 2) >> 2;
  constexpr int SLD = EPI_SLD;
  const int r_sub = lane >> 3;      // 4 rows per pass
  const int cq = (lane & 7) * 4;    // 4 consecutive columns per lane
  float alpha = e.alpha;
  if (e.alpha_dev) alpha *= __ldg(e.alpha_dev);
#pragma unroll 1
  for (int c = half; c < BN / 32; c += 2) {
    const int col0 = tile_n * BN + c * 32;
    if (col0 >= p.N) break;                       // warp-uniform
    float v[32];
    if (has_acc) {
      tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    float4* srow = reinterpret_cast<float4*>(stage + lane * SLD);
#pragma unroll
    for (int j = 0; j < 8; ++j) srow[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const int col = col0 + cq;
    const int row0 = tile_m * BM + q * 32 + r_sub;
    if (!(F & EF_GENERIC) || p.vec_ok) {             // N % 4 == 0: a lane's 4 columns are all in range or all out
      const bool lane_ok = col < p.N;
      float bs1[4] = {0.f, 0.f, 0.f, 0.f}, bs2[4] = {0.f, 0.f, 0.f, 0.f};
      const int bn_c = (F & EF_BNF) ? col % 40 : 0;
      // 4 row groups have their global loads issued before the first is consumed.  (8 in flight for the BatchNorm-backward
      // flavour -- one 16-byte load per group, long-scoreboard bound in ncu -- measured SLOWER on B200: 244 -> 267 us.)
      constexpr int U = 4;
#pragma unroll 1
      for (int g4 = 0; g4 < 8 / U; ++g4) {
        EpiLoads L[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int row = row0 + (g4 * U + u) * 4;
          if (row < p.M && lane_ok) L[u] = epi_load4<F>(e, row, col);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int rl = (g4 * U + u) * 4 + r_sub;
          const int row = row0 + (g4 * U + u) * 4;
          if (row < p.M && lane_ok) epi_finish4<F>(e, row, col, *reinterpret_cast<const float4*>(stage + rl * SLD + cq), L[u], alpha, bs1, bs2, sred + 80, bn_c);
        }
      }
      if (F & EF_BNF) {  // same 4 columns for every row of this lane: reduce over the 4 row lanes, then 8 lanes publish
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          bs1[i] += __shfl_xor_sync(0xffffffffu, bs1[i], 8);  bs1[i] += __shfl_xor_sync(0xffffffffu, bs1[i], 16);
          bs2[i] += __shfl_xor_sync(0xffffffffu, bs2[i], 8);  bs2[i] += __shfl_xor_sync(0xffffffffu, bs2[i], 16);
        }
        if (r_sub == 0 && lane_ok) {
          const int c = bn_c;
#pragma unroll
          for (int i = 0; i < 4; ++i) { atomicAdd(&sred[c + i], bs1[i]); atomicAdd(&sred[40 + c + i], bs2[i]); }
        }
      }
    } else {
      for (int it = 0; it < 8; ++it) {
        const int rl = it * 4 + r_sub;
        const int row = row0 + it * 4;
        if (row < p.M && col < p.N) {
          const float4 a4 = *reinterpret_cast<const float4*>(stage + rl * SLD + cq);
          const float av[4] = {a4.x, a4.y, a4.z, a4.w};
          if (p.pair_atomic && col + 4 <= p.N) {      // plain split-K accumulation into an 8-byte aligned row
            float* dst = e.C + (size_t)row * e.ldc + col;
            red_add_v2(dst, alpha * av[0], alpha * av[1]);
            red_add_v2(dst + 2, alpha * av[2], alpha * av[3]);
          } else {
            for (int i = 0; i < 4 && col + i < p.N; ++i) epi_store(e, row, col + i, epi_value(e, row, col + i, av[i]));
          }
        }
      }
    }
    __syncwarp();
  }
}

// producer / MMA helpers shared by the two kernels
template <int BN, int A_MN, int B_MN>
__device__ __forceinline__ void issue_stage_loads(const CUtensorMap* tmA, const CUtensorMap* tmB, const GemmKernelParams& p,
                                                  uint64_t* bar, uint8_t* sa, uint8_t* sb, int tile_m, int tile_n, int k0) {
  if (A_MN) {
    if (p.a_3d) {
      tma_load_3d(tmA, bar, sa, 0, k0, tile_m * (BM / 32));     // lands as [slab][k][32 mn]
    } else {
#pragma unroll
      for (int j = 0; j < BM / 32; ++j) tma_load_2d(tmA, bar, sa + j * (BK * 128), tile_m * BM + j * 32, k0);
    }
  } else {
    tma_load_2d(tmA, bar, sa, k0, tile_m * BM);
  }
  if (B_MN) {
    if (p.b_3d) {
      tma_load_3d(tmB, bar, sb, 0, k0, tile_n * (BN / 32));
    } else {
#pragma unroll
      for (int j = 0; j < BN / 32; ++j) tma_load_2d(tmB, bar, sb + j * (BK * 128), tile_n * BN + j * 32, k0);
    }
  } else {
    tma_load_2d(tmB, bar, sb, k0, tile_n * BN);
  }
}
template <int BN, int A_MN, int B_MN>
__device__ __forceinline__ void issue_stage_mmas(uint32_t sa, uint32_t sb, uint32_t tmem_acc, bool first) {
  constexpr uint32_t idesc = umma_idesc_tf32(BM, BN, A_MN, B_MN);
#pragma unroll
  for (int kk = 0; kk < BK / UMMA_K; ++kk) {
    // K-major : 8 rows x 128 B swizzle atoms, SBO = 1024 B, advance 32 B (8 floats) per MMA inside the atom
    // MN-major: [k][32 mn] slabs of BK*128 B (LBO), 4-k-row atoms of 512 B (SBO), advance 8 k-rows = 1024 B
    const uint64_t adesc = A_MN ? umma_smem_desc(sa + kk * 1024, BK * 128, 512, UMMA_LAYOUT_SW128_BASE32B)
                                : umma_smem_desc(sa + kk * 32, 16, 1024, UMMA_LAYOUT_SW128);
    const uint64_t bdesc = B_MN ? umma_smem_desc(sb + kk * 1024, BK * 128, 512, UMMA_LAYOUT_SW128_BASE32B)
                                : umma_smem_desc(sb + kk * 32, 16, 1024, UMMA_LAYOUT_SW128);
    tc_mma_tf32(tmem_acc, adesc, bdesc, idesc, (!first || kk > 0) ? 1u : 0u);
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM loops over (m-tile, n-tile, k-split) work items.  Two TMEM accumulators
// (2 x BN columns) let the epilogue of item i overlap the TMA/MMA main loop of item i+1; the smem ring keeps
// streaming across items.  Epilogue staging has its own smem (the ring is never idle).
// ------------------------------------------------------------------------------------------------
template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const GemmKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t A_BYTES = BM * BK * 4;
  constexpr uint32_t B_BYTES = BN * BK * 4;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  constexpr int MAX_STAGES = 8;
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* tiles = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  float* staging = reinterpret_cast<float*>(tiles + (size_t)p.tile_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + 8 * 32 * EPI_SLD);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  __shared__ float sred[80];
  if (threadIdx.x < 80) sred[threadIdx.x] = 0.f;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = (p.M + BM - 1) / BM;
  const int total_kb = (p.K + BK - 1) / BK;
  const int splits = (total_kb + p.k_blocks_per_split - 1) / p.k_blocks_per_split;
  const int items = tiles_m * tiles_n * splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 8);        // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (split, m, n): n fastest so that CTAs running at the same time share the A rows in L2
  auto decode = [&](int item, int& tm, int& tn, int& kb0, int& nkb) {
    tn = item % tiles_n;
    const int rest = item / tiles_n;
    tm = rest % tiles_m;
    const int sp = rest / tiles_m;
    kb0 = sp * p.k_blocks_per_split;
    nkb = min(total_kb, kb0 + p.k_blocks_per_split) - kb0;
  };

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int tm, tn, kb0, nkb;
        decode(item, tm, tn, kb0, nkb);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
          uint8_t* sa = tiles + (size_t)s * STAGE_BYTES;
          issue_stage_loads<BN, A_MN, B_MN>(&tmA, &tmB, p, &full_bar[s], sa, sa + A_BYTES, tm, tn, (kb0 + i) * BK);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int it = 0, j = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++j) {
        int tm, tn, kb0, nkb;
        decode(item, tm, tn, kb0, nkb);
        const int acc = j & 1;
        mbar_wait(&tmem_empty_bar[acc], ((uint32_t)(j >> 1) & 1u) ^ 1u);   // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)s * STAGE_BYTES);
          issue_stage_mmas<BN, A_MN, B_MN>(sa, sa + A_BYTES, tacc, i == 0);
          tc_commit(&empty_bar[s]);
        }
        tc_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    float* stage = staging + (size_t)(warp - 2) * 32 * EPI_SLD;
    int j = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++j) {
      int tm, tn, kb0, nkb;
      decode(item, tm, tn, kb0, nkb);
      const int acc = j & 1;
      mbar_wait(&tmem_full_bar[acc], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      epilogue_tile<BN, EF_GENERIC>(p, p.epi, tmem_base + (uint32_t)(acc * BN), stage, tm, tn, warp, lane, true, sred);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int A_MN, int B_MN, uint32_t F = EF_GENERIC>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t A_BYTES = BM * BK * 4;
  constexpr uint32_t B_BYTES = BN * BK * 4;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr int MAX_STAGES = 8;

  // 1024-byte aligned tile area (swizzle atoms), then barriers
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* tiles = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + (size_t)p.tile_bytes);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  constexpr bool BNF = (F & EF_BNF) != 0;
  __shared__ float sred[BNF ? 240 : 80];     // [0,80): BN-backward reductions; [80,240): mean | rstd | gamma | beta
  if (threadIdx.x < 80) sred[threadIdx.x] = 0.f;
  if (BNF && threadIdx.x < 160) {
    const int k = threadIdx.x % 40, w = threadIdx.x / 40;
    sred[80 + threadIdx.x] = w == 0 ? p.epi.bn_mean_rstd[k] : (w == 1 ? p.epi.bn_mean_rstd[40 + k] : (w == 2 ? p.epi.bn_gamma[k] : p.epi.bn_beta[k]));
  }
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile_n = blockIdx.x, tile_m = blockIdx.y;
  const int total_kb = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.k_blocks_per_split;
  const int kb_end = min(total_kb, kb_begin + p.k_blocks_per_split);
  const int num_kb = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && num_kb > 0) {
      // ================= TMA producer =================
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % p.stages;
        const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        uint8_t* sa = tiles + (size_t)s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        const int k0 = (kb_begin + i) * BK;
        if (A_MN) {
          if (p.a_3d) {
            tma_load_3d(&tmA, &full_bar[s], sa, 0, k0, tile_m * (BM / 32));     // lands as [slab][k][32 mn]
          } else {
#pragma unroll
            for (int j = 0; j < BM / 32; ++j) tma_load_2d(&tmA, &full_bar[s], sa + j * (BK * 128), tile_m * BM + j * 32, k0);
          }
        } else {
          tma_load_2d(&tmA, &full_bar[s], sa, k0, tile_m * BM);
        }
        if (B_MN) {
          if (p.b_3d) {
            tma_load_3d(&tmB, &full_bar[s], sb, 0, k0, tile_n * (BN / 32));
          } else {
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) tma_load_2d(&tmB, &full_bar[s], sb + j * (BK * 128), tile_n * BN + j * 32, k0);
          }
        } else {
          tma_load_2d(&tmB, &full_bar[s], sb, k0, tile_n * BN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && num_kb > 0) {
      // ================= MMA issuer (one thread) =================
      constexpr uint32_t idesc = umma_idesc_tf32(BM, BN, A_MN, B_MN);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % p.stages;
        const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(tiles + (size_t)s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
          // K-major : 8 rows x 128 B swizzle atoms, SBO = 1024 B, advance 32 B (8 floats) per MMA inside the atom
          // MN-major: [k][32 mn] slabs of BK*128 B (LBO), 4-k-row atoms of 512 B (SBO), advance 8 k-rows = 1024 B
          const uint64_t adesc = A_MN ? umma_smem_desc(sa + kk * 1024, BK * 128, 512, UMMA_LAYOUT_SW128_BASE32B)
                                      : umma_smem_desc(sa + kk * 32, 16, 1024, UMMA_LAYOUT_SW128);
          const uint64_t bdesc = B_MN ? umma_smem_desc(sb + kk * 1024, BK * 128, 512, UMMA_LAYOUT_SW128_BASE32B)
                                      : umma_smem_desc(sb + kk * 32, 16, 1024, UMMA_LAYOUT_SW128);
          tc_mma_tf32(tmem_base, adesc, bdesc, idesc, (i > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(&empty_bar[s]);   // frees this smem stage once the MMAs above have read it
      }
      tc_commit(tmem_full_bar);     // accumulator complete
    }
  } else {
    // ================= epilogue (warps 2..9; TMEM lane quarter = warp % 4, two warps share a quarter) =========
    if (num_kb > 0) {
      mbar_wait(tmem_full_bar, 0);    // all MMAs done => every smem stage has been consumed: the tile area is free
      tc_fence_after();
    }
    float* stage = reinterpret_cast<float*>(tiles) + (size_t)(warp - 2) * 32 * EPI_SLD;
    epilogue_tile<BN, F>(p, p.epi, tmem_base, stage, tile_m, tile_n, warp, lane, num_kb > 0, sred);
  }

  tc_fence_before();
  __syncthreads();
  if (BNF && threadIdx.x < 80) atomicAdd(&p.epi.bn_sums[threadIdx.x], (double)sred[threadIdx.x]);
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

static int resolve_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  if (!g_encode) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return 3;
  }
  return 0;
}

// operand viewed as logical [rows, K]
static int make_operand_map(CUtensorMap* tm, const GemmOperand& op, int rows, int K, int box_rows, int* is_3d) {
  *is_3d = 0;
  EEG_REQUIRE(op.ptr != nullptr, "gemm: null operand");
  EEG_REQUIRE((op.ld & 3) == 0, "gemm: leading dimension %d is not a multiple of 4 floats (TMA needs 16-byte strides)", op.ld);
  EEG_REQUIRE((reinterpret_cast<uintptr_t>(op.ptr) & 15) == 0, "gemm: operand pointer not 16-byte aligned");
  cuuint64_t dims[3];
  cuuint64_t strides[2] = {(cuuint64_t)op.ld * 4, 128};
  cuuint32_t box[3];
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw;
  int rank = 2;
  if (op.mn_major) {
    // stored [K][rows]; viewed as (32 mn, K, rows/32 slabs) so that one box lands as [slab][k][32 mn] in smem,
    // the canonical MN-major SWIZZLE_128B_ATOM_32B layout (slab stride = LBO, 4-k-row atoms = SBO)
    EEG_REQUIRE(op.ld >= rows, "gemm: MN-major operand ld %d < rows %d", op.ld, rows);
    sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    const int slabs = (rows + 31) / 32;
    if (slabs * 32 <= op.ld) {
      // the last slab reaches past `rows` but stays inside the row pitch (pad columns): one 3-D box per stage
      rank = 3;
      *is_3d = 1;
      dims[0] = 32; dims[1] = (cuuint64_t)K; dims[2] = (cuuint64_t)slabs;
      box[0] = 32; box[1] = BK; box[2] = (cuuint32_t)(box_rows / 32);
    } else {
      dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K;
      box[0] = 32; box[1] = BK;
    }
  } else {
    EEG_REQUIRE(op.ld >= K, "gemm: K-major operand ld %d < K %d", op.ld, K);
    dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows;
    box[0] = BK; box[1] = (cuuint32_t)box_rows;
    sw = CU_TENSOR_MAP_SWIZZLE_128B;
  }
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<float*>(op.ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EEG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (rows %d K %d ld %d mn %d)", (int)r, rows, K, op.ld,
              op.mn_major);
  return 0;
}

// tensor map of a GEMM-style operand for kernels outside this file (attention_tc.cu): box = 32 k x box_rows rows
// (K-major, SWIZZLE_128B) or (32 mn, 32 k, box_rows/32 slabs) (MN-major, SWIZZLE_128B_ATOM_32B)
int gemm_make_tmap(CUtensorMap* tm, const GemmOperand& op, int rows, int K, int box_rows, int* is_3d) {
  EEG_TRY(resolve_encode());
  return make_operand_map(tm, op, rows, K, box_rows, is_3d);
}

template <int BN, int A_MN, int B_MN>
static int launch_cfg(const GemmArgs& g, const CUtensorMap& ta, const CUtensorMap& tb, int a3, int b3, cudaStream_t stream) {
  const int total_kb = cdiv(g.K, BK);
  int split = g.split_k < 1 ? 1 : g.split_k;
  if (split > total_kb) split = total_kb > 0 ? total_kb : 1;
  GemmKernelParams p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.k_blocks_per_split = cdiv(total_kb, split);
  split = p.k_blocks_per_split > 0 ? cdiv(total_kb, p.k_blocks_per_split) : 1;
  const int stage_bytes = (BM + BN) * BK * 4;
  // short K: shallow pipeline so two CTAs share an SM (one CTA's epilogue overlaps the other's main loop);
  // long K: deep pipeline, one CTA per SM.
  const int budget = (p.k_blocks_per_split >= 16 || split > 1) ? 196608 : 98304;
  int stages = budget / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages > p.k_blocks_per_split) stages = p.k_blocks_per_split;
  if (stages < 1) stages = 1;
  p.stages = stages;
  p.epi = g.epi;
  p.vec_ok = (epi_vec_ok(g.epi) && (g.N & 3) == 0) ? 1 : 0;
  p.a_3d = a3; p.b_3d = b3;
  {
    const Epilogue& e = g.epi;
    p.pair_atomic = (e.store_mode == EPI_ATOMIC && !e.bias && !e.aux_out && !e.mul_in && !e.resid && e.act == EPI_ACT_NONE &&
                     e.drop.p <= 0.f && !e.round_tf32 && (e.ldc & 1) == 0 && (reinterpret_cast<uintptr_t>(e.C) & 7) == 0) ? 1 : 0;
  }
  size_t tile_bytes = (size_t)stages * stage_bytes;
  if (tile_bytes < 36864) tile_bytes = 36864;      // epilogue staging (8 warps x 32 x 36 floats) reuses the tile area
  p.epi = g.epi;
  p.vec_ok = (epi_vec_ok(g.epi) && (g.N & 3) == 0) ? 1 : 0;
  p.a_3d = a3; p.b_3d = b3;
  {
    const Epilogue& e = g.epi;
    p.pair_atomic = (e.store_mode == EPI_ATOMIC && !e.bias && !e.aux_out && !e.mul_in && !e.resid && e.act == EPI_ACT_NONE &&
                     e.drop.p <= 0.f && !e.round_tf32 && (e.ldc & 1) == 0 && (reinterpret_cast<uintptr_t>(e.C) & 7) == 0) ? 1 : 0;
  }
  static int persistent = -1;
  static int sms_by_dev[64] = {0};
  if (persistent < 0) {
    const char* env = getenv("EEGB200_GEMM_PERSISTENT");
    persistent = (env && env[0] == '0') ? 0 : 1;
  }
  const int dev = current_device();
  if (sms_by_dev[dev] == 0) {
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms_by_dev[dev] = n > 0 ? n : 148;
  }
  const int num_sms = sms_by_dev[dev];
  // persistent (1 CTA/SM, overlapped epilogue) pays off when the main loop is long; the short-K token GEMMs are bound by
  // epilogue memory latency and run faster as two co-resident CTAs per SM (16 epilogue warps) -- measured on B200
  if (persistent && total_kb > 0 && p.k_blocks_per_split >= 16 && g.epi.bn_y == nullptr) {
    // one CTA per SM: stages fill what is left of the 227 KB after the dedicated epilogue staging
    int pst = (226 * 1024 - 1024 - 8 * 32 * EPI_SLD * 4 - 512) / stage_bytes;   // 1 KB of the 227 KB is static smem
    if (pst > 8) pst = 8;
    if (pst > 2 * p.k_blocks_per_split) pst = 2 * p.k_blocks_per_split;    // ring streams across items
    if (pst < 2) pst = 2;
    p.stages = pst;
    p.tile_bytes = pst * stage_bytes;
    const size_t psmem = (size_t)p.tile_bytes + 8 * 32 * EPI_SLD * 4 + 1024 + 512;
    auto pk = gemm_tf32_persistent_kernel<BN, A_MN, B_MN>;
    static PerDeviceOnce pconf;
    if (pconf.first()) EEG_CUDA_OK(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    const int items = cdiv(g.N, BN) * cdiv(g.M, BM) * split;
    const int grid_p = items < num_sms ? items : num_sms;
    pk<<<grid_p, GEMM_THREADS, psmem, stream>>>(ta, tb, p);
    EEG_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
  }
  const size_t smem = tile_bytes + 1024 + 256;
  p.tile_bytes = (int)tile_bytes;
  dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), split);
  // kernels specialised on the exact epilogue feature set of the hot 256-wide GEMMs (everything else: generic kernel)
  if constexpr (BN == 256 && A_MN == 0) {
    const uint32_t fm = epi_feature_mask(g.epi);
    const bool spec_ok = p.vec_ok && g.epi.store_mode == EPI_STORE && split == 1 && (g.N % 32) == 0 &&
                         (!(fm & EF_BIAS_TABLE) || g.epi.bias_period == 64);
    if (fm & EF_BNF) {
      EEG_REQUIRE(B_MN == 1 && p.vec_ok && fm == EF_BNF && g.epi.bn_sums && g.epi.bn_mean_rstd && g.epi.bn_gamma &&
                  g.epi.bn_beta && split == 1, "gemm: unsupported use of the fused BatchNorm-backward epilogue");
    }
#define EEG_SPEC(MASK)                                                                                        \
    if (fm == (MASK)) {                                                                                       \
      auto ks = gemm_tf32_kernel<256, 0, B_MN, (MASK)>;                                                       \
      static PerDeviceOnce cs;                                                                                \
      if (cs.first())                                                                                         \
        EEG_CUDA_OK(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));       \
      ks<<<grid, GEMM_THREADS, smem, stream>>>(ta, tb, p);                                                    \
      EEG_CUDA_OK(cudaGetLastError());                                                                        \
      count_launch();                                                                                         \
      return 0;                                                                                               \
    }
    if (spec_ok || (fm & EF_BNF)) {
      if constexpr (B_MN == 0) {
        EEG_SPEC(EF_BIAS_TABLE | EF_DROP | EF_ROUND)                       // value embedding (train)
        EEG_SPEC(EF_BIAS)                                                   // QKV projection
        EEG_SPEC(EF_BIAS | EF_ROUND)                                        // QKV projection, TF32-rounded for the tcgen05 attention
        EEG_SPEC(EF_BIAS | EF_DROP | EF_RESID)                              // out-projection / FFN2 + residual (train)
        EEG_SPEC(EF_BIAS | EF_AUX | EF_GELU | EF_DROP | EF_ROUND)           // FFN1 (train)
      } else {
        EEG_SPEC(EF_DROP | EF_MUL | EF_ROUND)                               // dU = dropout(T1.W2) * GELU'(U)
        EEG_SPEC(EF_RESID)                                                  // dX1, dH0
        EEG_SPEC(0u)                                                        // dO
        EEG_SPEC(EF_BNF)                                                    // dA1 with the BatchNorm1+ELU backward
      }
    }
#undef EEG_SPEC
    EEG_REQUIRE(!(fm & EF_BNF), "gemm: the fused BatchNorm-backward epilogue has no generic fallback");
  } else {
    EEG_REQUIRE(g.epi.bn_y == nullptr, "gemm: the fused BatchNorm-backward epilogue is only built for the 256-wide K-major x MN-major kernel");
  }
  auto kern = gemm_tf32_kernel<BN, A_MN, B_MN>;
  static PerDeviceOnce kconf;     // per template instantiation and device
  if (kconf.first()) EEG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  kern<<<grid, GEMM_THREADS, smem, stream>>>(ta, tb, p);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

template <int BN>
static int launch_bn(const GemmArgs& g, cudaStream_t stream) {
  CUtensorMap ta, tb;
  int a3 = 0, b3 = 0;
  EEG_TRY(make_operand_map(&ta, g.A, g.M, g.K, BM, &a3));
  EEG_TRY(make_operand_map(&tb, g.B, g.N, g.K, BN, &b3));
  if (!g.A.mn_major && !g.B.mn_major) return launch_cfg<BN, 0, 0>(g, ta, tb, a3, b3, stream);
  if (!g.A.mn_major && g.B.mn_major) return launch_cfg<BN, 0, 1>(g, ta, tb, a3, b3, stream);
  if (g.A.mn_major && !g.B.mn_major) return launch_cfg<BN, 1, 0>(g, ta, tb, a3, b3, stream);
  return launch_cfg<BN, 1, 1>(g, ta, tb, a3, b3, stream);
}

int gemm_launch_tcgen05(const GemmArgs& g, cudaStream_t stream) {
  EEG_REQUIRE(g.M > 0 && g.N > 0 && g.K >= 0, "gemm: bad shape %d x %d x %d", g.M, g.N, g.K);
  EEG_REQUIRE(g.epi.C != nullptr, "gemm: null output");
  EEG_REQUIRE(g.split_k <= 1 || g.epi.store_mode == EPI_ATOMIC, "gemm: split-K needs the atomic store mode");
  EEG_TRY(resolve_encode());
  // widest tile that still yields >= ~one wave of CTAs (small-M problems: projector, logits, retrieval)
  if (g.epi.bn_y != nullptr) return launch_bn<256>(g, stream);     // the fused BatchNorm-backward flavour is built for BN = 256
  if (g.tile_n == 256) return launch_bn<256>(g, stream);
  if (g.tile_n == 128) return launch_bn<128>(g, stream);
  if (g.tile_n == 64) return launch_bn<64>(g, stream);
  EEG_REQUIRE(g.tile_n == 0, "gemm: tile_n must be 0, 64, 128 or 256 (got %d)", g.tile_n);
  const int mt = cdiv(g.M, BM);
  const int split = g.split_k > 1 ? g.split_k : 1;
  auto ctas = [&](int bn) { return mt * cdiv(g.N, bn) * split; };
  if (g.N > 128 && ctas(256) >= 120) return launch_bn<256>(g, stream);
  if (g.N > 64 && (ctas(128) >= 120 || g.N <= 128)) return launch_bn<128>(g, stream);
  if (g.N <= 64 || ctas(128) < 120) return launch_bn<64>(g, stream);
  return launch_bn<128>(g, stream);
}

}  // namespace eegb200

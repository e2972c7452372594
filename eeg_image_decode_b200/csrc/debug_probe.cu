// Debug probe: the shared-memory image that TMA writes for one GEMM operand k-block (32 k x 128 rows), copied out
// verbatim.  tools/gpu_tma_layout_probe.py compares it with the closed-form swizzle formulas that thread-written UMMA
// operands have to follow (docs/ROUND2_CONV_TCGEN05.md): K-major SWIZZLE_128B (already relied upon by attention_tc.cu) and
// MN-major SWIZZLE_128B_ATOM_32B (needed by the round-2 backward kernels, formula not yet verified on hardware).
#include "../../include/eegdecode_b200.h"
#include "gemm.h"

namespace eegb200 {

__global__ void tma_tile_dump_kernel(const __grid_constant__ CUtensorMap tm, int is_3d, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* tile = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(tile + 16384);
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<float*>(tile)[i] = -1.f;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 16384);
    if (is_3d) tma_load_3d(&tm, bar, tile, 0, 0, 0);      // (32 mn, 32 k, 4 slabs) -> [slab][k][32 mn]
    else tma_load_2d(&tm, bar, tile, 0, 0);               // (32 k, 128 rows)      -> [row][32 k]
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = reinterpret_cast<float*>(tile)[i];
}


// Which TMEM lanes hold the rows of an M = 64 accumulator (cta_group::1)?  One UMMA with A[r][0] = r + 1 (K-major, 64
// rows), B[0][0] = 1 writes D[r][0] = r + 1; every lane's column 0 is dumped (untouched lanes keep the -1 they were
// pre-filled with through an M = 128 UMMA with a -1 operand).   tools/gpu_m64_probe.py
__global__ void umma_m64_probe_kernel(float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* A = sm;                       // [128 rows][32 floats] K-major SWIZZLE_128B
  uint8_t* Bm = sm + 16384;              // [8 rows][32 floats]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384 + 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < (16384 + 1024) / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
  __syncthreads();
  const int r = threadIdx.x;
  // element (r, 0): chunk 0 of row r
  *reinterpret_cast<float*>(A + (r >> 3) * 1024 + (r & 7) * 128 + ((0 ^ (r & 7)) << 4)) = -1.f;     // pass 1: all 128 rows = -1
  if (r == 0) *reinterpret_cast<float*>(Bm) = 1.f;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 32); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    tc_mma_tf32(tmem, umma_smem_desc(smem_u32(A), 16, 1024, UMMA_LAYOUT_SW128), umma_smem_desc(smem_u32(Bm), 16, 1024, UMMA_LAYOUT_SW128),
                umma_idesc_tf32(128, 8, 0, 0), 0u);
    tc_commit(&bar[0]);
  }
  mbar_wait(&bar[0], 0);
  tc_fence_after();
  __syncthreads();
  if (r < 64) *reinterpret_cast<float*>(A + (r >> 3) * 1024 + (r & 7) * 128 + ((0 ^ (r & 7)) << 4)) = (float)(r + 1);   // pass 2
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    tc_mma_tf32(tmem, umma_smem_desc(smem_u32(A), 16, 1024, UMMA_LAYOUT_SW128), umma_smem_desc(smem_u32(Bm), 16, 1024, UMMA_LAYOUT_SW128),
                umma_idesc_tf32(64, 8, 0, 0), 0u);
    tc_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tc_fence_after();
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16)) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  out[threadIdx.x] = __uint_as_float(v);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 32);
}

// Generic operand-layout probe: one UMMA chain D[128][N] = A . B^T (kind::tf32) over shared-memory IMAGES prepared on the
// host (a_img: 32 KB, b_img: 32 KB, copied verbatim), with every descriptor field given by the caller:
// cfg = {a_layout, a_lbo, a_sbo, a_kadv, a_major, b_layout, b_lbo, b_sbo, b_kadv, b_major, N, nk}.  Used to check which
// (major, swizzle) combinations the tensor core accepts for a tile that two different UMMAs want to read
// (tools/gpu_dual_layout_probe.py).  out: [128][N] accumulator rows.
struct UmmaProbeCfg { int v[12]; };
__global__ void umma_generic_probe_kernel(const float* __restrict__ a_img, const float* __restrict__ b_img, UmmaProbeCfg c,
                                          float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* A = sm;
  uint8_t* Bm = sm + 32768;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) {
    reinterpret_cast<float*>(A)[i] = a_img[i];
    reinterpret_cast<float*>(Bm)[i] = b_img[i];
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const int N = c.v[10], nk = c.v[11];
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, N, c.v[4], c.v[9]);
    for (int kk = 0; kk < nk; ++kk) {
      const uint64_t da = umma_smem_desc(smem_u32(A) + (uint32_t)(kk * c.v[3]), (uint32_t)c.v[1], (uint32_t)c.v[2], (uint32_t)c.v[0]);
      const uint64_t db = umma_smem_desc(smem_u32(Bm) + (uint32_t)(kk * c.v[8]), (uint32_t)c.v[6], (uint32_t)c.v[7], (uint32_t)c.v[5]);
      tc_mma_tf32(tmem, da, db, idesc, kk > 0 ? 1u : 0u);
    }
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5;
  for (int n = 0; n < N; ++n) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)n) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[threadIdx.x * N + n] = __uint_as_float(v);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}


// Cycles per tcgen05.mma (kind::tf32, K = 8) for the small shapes of the fused conv kernels: a chain of `n` UMMAs issued
// back to back by one thread into one accumulator, operands resident in shared memory.  `bg` adds background activity on
// the same SM: 1 = 8 warps streaming st.shared.v4, 2 = a SECOND thread (another warp) issuing its own UMMA chain,
// 4 = 4 warps looping tcgen05.ld on the accumulator (bits may be combined).
// out[0] = cycles from the first issue to the commit arrival, out[1] = cycles spent issuing.   tools/gpu_mma_cost.py
__global__ void umma_cost_kernel(int M, int N, int mn_major, int n, int bg, long long* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 160 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  volatile int* stop = reinterpret_cast<volatile int*>(slot + 1);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); *stop = 0; }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 || (warp == 1 && (bg & 2))) {
    if (lane == 0) {
      const uint32_t a0 = smem_u32(sm) + warp * 32768, b0 = smem_u32(sm + 64 * 1024) + warp * 32768;
      // bit 64: the SECOND issuer runs the other shape (M = 128, N = 48, K-major) instead of the primary's
      const bool other = warp == 1 && (bg & 64);
      const int mnm = other ? 0 : mn_major;
      const uint32_t idesc = other ? umma_idesc_tf32(128, 48, 0, 0) : umma_idesc_tf32(M, N, mn_major, mn_major);
      const uint32_t idesc_alt = umma_idesc_tf32(128, 48, 0, 0);
      uint64_t ad[4], bd[4], ad2[4], bd2[4];
      for (int kk = 0; kk < 4; ++kk) {
        ad[kk] = mnm ? umma_smem_desc(a0 + kk * 1024, 4096, 512, UMMA_LAYOUT_SW128_BASE32B)
                     : umma_smem_desc(a0 + kk * 32, 16, 1024, UMMA_LAYOUT_SW128);
        bd[kk] = mnm ? umma_smem_desc(b0 + kk * 1024, 4096, 512, UMMA_LAYOUT_SW128_BASE32B)
                     : umma_smem_desc(b0 + kk * 32, 16, 1024, UMMA_LAYOUT_SW128);
        ad2[kk] = umma_smem_desc(a0 + 16384 + kk * 32, 16, 1024, UMMA_LAYOUT_SW128);
        bd2[kk] = umma_smem_desc(b0 + 16384 + kk * 32, 16, 1024, UMMA_LAYOUT_SW128);
      }
      const uint32_t d = tmem + warp * 256;
      const long long t0 = clock64();
      tc_mma_tf32(d, ad[0], bd[0], idesc, 0u);
      tc_mma_tf32(d + 128, ad2[0], bd2[0], idesc_alt, 0u);
      for (int i = 1; i < n / 4; ++i) {
        if ((bg & 32) && (i & 1)) {
          // bit 32: every other block of 4 UMMAs has the other shape (M = 128, N = 48, K-major) and its own accumulator
          tc_mma_tf32(d + 128, ad2[0], bd2[0], idesc_alt, 1u);
          tc_mma_tf32(d + 128, ad2[1], bd2[1], idesc_alt, 1u);
          tc_mma_tf32(d + 128, ad2[2], bd2[2], idesc_alt, 1u);
          tc_mma_tf32(d + 128, ad2[3], bd2[3], idesc_alt, 1u);
        } else {
          tc_mma_tf32(d, ad[0], bd[0], idesc, 1u);
          tc_mma_tf32(d, ad[1], bd[1], idesc, 1u);
          tc_mma_tf32(d, ad[2], bd[2], idesc, 1u);
          tc_mma_tf32(d, ad[3], bd[3], idesc, 1u);
        }
        if (bg & 8) tc_commit(&bar[1 - warp]);         // a tcgen05.commit after every 4 UMMAs (nobody waits on it)
      }
      const long long t1 = clock64();
      tc_commit(&bar[warp]);
      mbar_wait(&bar[warp], 0);
      const long long t2 = clock64();
      if (warp == 0) { out[0] = t2 - t0; out[1] = t1 - t0; *stop = 1; }
    }
  } else if (warp >= 4 && warp < 12 && (bg & 1)) {
    // background shared-memory stores (conflict-free 512 B per instruction) into a region nobody reads
    float4* dst = reinterpret_cast<float4*>(sm + 128 * 1024) + (warp - 4) * 32 + lane;
    int it = 0;
    while (!*stop && it < (1 << 22)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[(j & 3) * 256] = make_float4((float)it, 0.f, 0.f, 0.f);
      ++it;
    }
  } else if (warp >= 4 && warp < 16 && (bg & 128)) {
    // background shared-memory LOADS and stores, 12 warps, scalar and partly bank-conflicted (what builders / scatter do)
    float* base = reinterpret_cast<float*>(sm + 128 * 1024) + (warp - 4) * 640;
    int it = 0;
    float acc = 0.f;
    while (!*stop && it < (1 << 22)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += base[5 * lane + j];
#pragma unroll
      for (int j = 0; j < 4; ++j) base[(8 * lane + j) & 511] = acc;
      ++it;
    }
    if (acc == 12345.f) out[1] = 0;
  } else if (warp >= 4 && warp < 16 && (bg & 256)) {
    // background generic-proxy stores each followed by fence.proxy.async (what every operand-writing thread does)
    float4* dst = reinterpret_cast<float4*>(sm + 128 * 1024) + (warp - 4) * 32 + lane;
    int it = 0;
    while (!*stop && it < (1 << 22)) {
      dst[(it & 3) * 512] = make_float4((float)it, 0.f, 0.f, 0.f);
      fence_proxy_async_smem();
      if (bg & 512) __nanosleep(200);       // ~ one fence per warp per few hundred cycles instead of back to back
      ++it;
    }
  } else if (warp >= 4 && warp < 16 && (bg & 16)) {
    // background mbarrier polling: 12 warps spin on a barrier phase that never completes (what waiting roles do)
    int it = 0;
    while (!*stop && it < (1 << 22)) {
      (void)mbar_try_wait(&bar[1], 0);
      ++it;
    }
  } else if (warp >= 12 && warp < 16 && (bg & 4)) {
    int it = 0;
    float acc = 0.f;
    while (!*stop && it < (1 << 22)) {
      float v[16];
      tmem_ld_32x32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 128u, v);     // columns the MMAs do not write
      acc += v[0];
      ++it;
    }
    if (acc == 12345.f) out[1] = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

}  // namespace eegb200

using namespace eegb200;

extern "C" int eegb200_debug_tma_tile(const float* src, int ld, int mn_major, float* out, void* stream) {
  EEG_REQUIRE(src && out && ld >= 128 && (ld & 3) == 0, "debug_tma_tile: bad arguments");
  CUtensorMap tm;
  int d3 = 0;
  // logical operand [128 rows, K = 32]: K-major reads src[row*ld + k], MN-major reads src[k*ld + row]
  EEG_TRY(gemm_make_tmap(&tm, GemmOperand{src, ld, mn_major}, 128, 32, 128, &d3));
  EEG_REQUIRE(!mn_major || d3 == 1, "debug_tma_tile: expected a 3-D map for the MN-major operand");
  static PerDeviceOnce once;
  if (once.first()) {
    EEG_CUDA_OK(cudaFuncSetAttribute(tma_tile_dump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 64 + 1024));
  }
  tma_tile_dump_kernel<<<1, 128, 16384 + 64 + 1024, (cudaStream_t)stream>>>(tm, d3, out);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int eegb200_debug_umma_m64(float* out128, void* stream) {
  EEG_REQUIRE(out128 != nullptr, "debug_umma_m64: null output");
  static PerDeviceOnce once;
  if (once.first())
    EEG_CUDA_OK(cudaFuncSetAttribute(umma_m64_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024 + 64 + 1024));
  umma_m64_probe_kernel<<<1, 128, 16384 + 1024 + 64 + 1024, (cudaStream_t)stream>>>(out128);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int eegb200_debug_umma_cost(int M, int N, int mn_major, int n, int bg, long long* out2, void* stream) {
  EEG_REQUIRE(out2 && (M == 64 || M == 128) && N >= 8 && N <= 256 && n > 0 && bg >= 0 && bg < 1024 && (!mn_major || N <= 128),
              "debug_umma_cost: bad arguments");
  static PerDeviceOnce once;
  if (once.first())
    EEG_CUDA_OK(cudaFuncSetAttribute(umma_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024 + 64 + 1024));
  umma_cost_kernel<<<1, 512, 160 * 1024 + 64 + 1024, (cudaStream_t)stream>>>(M, N, mn_major, n, bg, out2);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int eegb200_debug_umma_generic(const float* a_img, const float* b_img, const int* cfg12, float* out, void* stream) {
  EEG_REQUIRE(a_img && b_img && cfg12 && out, "debug_umma_generic: null argument");
  UmmaProbeCfg c;
  for (int i = 0; i < 12; ++i) c.v[i] = cfg12[i];
  EEG_REQUIRE(c.v[10] >= 8 && c.v[10] <= 256 && c.v[10] % 8 == 0 && c.v[11] >= 1 && c.v[11] <= 64, "debug_umma_generic: bad N / nk");
  static PerDeviceOnce once;
  if (once.first())
    EEG_CUDA_OK(cudaFuncSetAttribute(umma_generic_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64 + 1024));
  umma_generic_probe_kernel<<<1, 128, 65536 + 64 + 1024, (cudaStream_t)stream>>>(a_img, b_img, c, out);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

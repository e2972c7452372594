// Shared device/host helpers for the sm_100a kernels: error handling, PTX wrappers
// (mbarrier / TMA / tcgen05 / TMEM), Philox dropout, small math.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <atomic>

namespace eegb200 {

// ------------------------------------------------------------------------------------------------
// error plumbing: every C-ABI entry point returns 0 on success, non-zero on failure and leaves a
// message retrievable through eegb200_last_error().
// ------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define EEG_CUDA_OK(expr)                                                                       \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::eegb200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                                 \
    }                                                                                           \
  } while (0)

#define EEG_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) {                                               \
      ::eegb200::set_error(__VA_ARGS__);                         \
      return 2;                                                  \
    }                                                            \
  } while (0)

#define EEG_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

// One-time setup that is per DEVICE (cudaFuncSetAttribute, side streams, SM count): first() is true exactly once per
// (call site, current device).  The Python binding makes the tensors' device current around every library call.
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    const unsigned long long bit = 1ull << (d & 63);
    return (done.fetch_or(bit) & bit) == 0;
  }
};
static inline int current_device() { int d = 0; cudaGetDevice(&d); return d & 63; }

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// math
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rn(float x) {
  // round-to-nearest(-away) to the 10-bit TF32 mantissa; the tensor core then reads an exact value
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float tf32_if(float x, int on) { return on ? tf32_rn(x) : x; }
// erf via Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7, below the fp32 noise of the surrounding GEMMs) sharing the
// exp(-x^2/2) with the Gaussian pdf: ~15 instructions instead of ~40 for erff.  ex = exp(-z^2), z = |x|/sqrt(2).
__device__ __forceinline__ float erf_as(float z_abs, float ex) {
  const float t = __frcp_rn(fmaf(0.3275911f, z_abs, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  return 1.0f - p * t * ex;
}
__device__ __forceinline__ float gelu_exact(float x) {   // F.gelu default (erf form)
  const float ex = __expf(-0.5f * x * x);
  const float e = erf_as(fabsf(x) * 0.70710678118654752440f, ex);
  return 0.5f * x * (1.0f + copysignf(e, x));
}
__device__ __forceinline__ float gelu_grad(float x) {
  const float ex = __expf(-0.5f * x * x);
  const float e = erf_as(fabsf(x) * 0.70710678118654752440f, ex);
  return 0.5f * (1.0f + copysignf(e, x)) + x * 0.39894228040143267794f * ex;
}
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }
// ELU'(x) expressed through x itself
__device__ __forceinline__ float elu1_grad(float x) { return x > 0.f ? 1.f : __expf(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------
// Counter-based dropout RNG.  Masks are a pure function of (seed, site, element index), so the backward pass and the
// test-side mask dump regenerate them instead of storing them.  One SplitMix64 finaliser (2 multiplies) per 4
// consecutive elements yields four 16-bit uniform lanes; keep <=> lane >= p * 65536 (p = 0.25 / 0.5 are exact).
// (Philox4x32-10 was ~80 instructions per 4 elements and showed up as ~25 % of the GEMM epilogue instruction count.)
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct DropoutCfg {
  const unsigned long long* seed_dev;   // optional device counter added to `seed` at run time (lets a captured CUDA graph
                                        // draw fresh masks on every replay); nullptr -> `seed` alone
  uint64_t seed;
  uint32_t site;     // which dropout layer
  float p;           // drop probability; p <= 0 disables
  float scale;       // 1/(1-p)
  uint32_t thresh;   // keep iff 16-bit lane >= thresh
};
static inline DropoutCfg make_dropout(uint64_t seed, uint32_t site, float p, bool train,
                                      const unsigned long long* seed_dev = nullptr) {
  DropoutCfg d;
  d.seed_dev = seed_dev;
  d.seed = seed; d.site = site;
  d.p = (train && p > 0.f) ? p : 0.f;
  d.scale = d.p > 0.f ? 1.f / (1.f - d.p) : 1.f;
  double t = (double)d.p * 65536.0;
  d.thresh = d.p > 0.f ? (uint32_t)(t > 65535.0 ? 65535.0 : t + 0.5) : 0u;
  return d;
}
__device__ __forceinline__ uint64_t dropout_bits(const DropoutCfg& cfg, uint64_t blk) {
  uint64_t sd = cfg.seed;
  if (cfg.seed_dev) sd += __ldg(cfg.seed_dev) * 0x9E3779B97F4A7C15ull;
  return splitmix64(blk * 0xD1342543DE82EF95ull + (sd ^ ((uint64_t)cfg.site << 56)) + 0x9E3779B97F4A7C15ull);
}
// keep-masks for the 4 consecutive elements starting at idx (idx % 4 == 0); bit i = keep element i
__device__ __forceinline__ uint32_t dropout_keep4(const DropoutCfg& cfg, uint64_t idx) {
  const uint64_t r = dropout_bits(cfg, idx >> 2);
  return ((uint32_t)(r & 0xFFFFu) >= cfg.thresh ? 1u : 0u) | ((uint32_t)((r >> 16) & 0xFFFFu) >= cfg.thresh ? 2u : 0u) |
         ((uint32_t)((r >> 32) & 0xFFFFu) >= cfg.thresh ? 4u : 0u) | ((uint32_t)(r >> 48) >= cfg.thresh ? 8u : 0u);
}
__device__ __forceinline__ bool dropout_keep(const DropoutCfg& cfg, uint64_t idx) {
  const uint64_t r = dropout_bits(cfg, idx >> 2);
  return (uint32_t)((r >> (16 * (idx & 3))) & 0xFFFFu) >= cfg.thresh;
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
      "elect.sync R|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifndef EEGB200_TRYWAIT_HINT_NS
#define EEGB200_TRYWAIT_HINT_NS 128u
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  // the suspend-time hint (ns) lets the hardware park the warp until the phase completes or the hint expires instead of
  // returning at once: far fewer polling instructions compete with the working warps for issue slots
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(EEGB200_TRYWAIT_HINT_NS)
      : "memory");
  return ok != 0;
}
// Bounded wait: a mis-programmed TMA/MMA would otherwise hang the GPU; after 2^25 polls (seconds) we trap so the launch fails with an error instead.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 25)) {
      printf("eegb200: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}

// ---- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 inputs, FP32 accumulate.  Issued by ONE thread.
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 columns of fp32 -> 32 registers per thread (thread = TMEM lane = tile row)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float v[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// vector reductions into global memory (sm_90+): one L2 atomic transaction for 2 / 4 consecutive floats
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// ---- UMMA descriptors (bit layouts: cute/arch/mma_sm100_desc.hpp in CUTLASS) ----
// shared-memory matrix descriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout [61,64)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7u) << 61;
  return d;
}
constexpr uint32_t UMMA_LAYOUT_SW128 = 2;          // K-major tiles written by TMA SWIZZLE_128B
constexpr uint32_t UMMA_LAYOUT_SW128_BASE32B = 1;  // MN-major 32-bit tiles written by TMA SWIZZLE_128B_ATOM_32B
// instruction descriptor for kind::tf32, fp32 accumulate
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace eegb200

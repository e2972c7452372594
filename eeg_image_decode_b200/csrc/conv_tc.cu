// PatchEmbedding conv stack forward on tcgen05 (docs/ROUND2_CONV_TCGEN05.md, kernels F1 / F2).
//
// STATUS: EXPERIMENTAL, OFF BY DEFAULT (EEGB200_CONV_TC=1 selects it).  Written at the end of round 1 after the GPU
// budget was spent: it compiles for sm_100a and follows the blocking that tools/conv_tc_plan.py checks against autograd
// on the CPU, but it has NOT run on a GPU yet.  The default path (conv_mma.cu + convstack.cu + the spatial GEMM) is the
// verified one.  First thing to do in round 2: EEGB200_CONV_TC=1 python -m pytest tests/test_gpu_parity.py -k stages.
//
// One CTA = 3 consecutive samples = one 128-row UMMA tile (rows (s, p): sample s < 3, pooled position p < 36; 108 valid),
// looping over the 63 channel-tokens:
//
//   builders (warps 0-3)  token rows of channel c -> box-51 pooled sums (warp prefix scan) -> im2col tile
//                         A[(s,p), t] = ps[s][5p+t], t < 25 (K padded to 32), split hi / lo for 3xTF32, written into
//                         SWIZZLE_128B K-major shared memory (double buffered)
//   control (warp 8)      conv UMMAs  C1[buf] = A_lo.Bhi + A_hi.Blo + A_hi.Bhi   (M=128, N=48, K=32, TMEM double buffered)
//                         spatial UMMAs  Y2 += A1_c . Ws_c^T                      (M=128, N=48, K=40; one channel behind)
//                         TMA of the per-channel spatial weights Ws_c [48 x 64] (pre-packed, 4-deep ring)
//   epilogue (warps 4-7)  thread = TMEM lane = row: y = C1 + bias
//                           MODE_STATS: accumulate sum / sum^2 per filter (BatchNorm1 batch statistics, kernel F1)
//                           MODE_APPLY: a1 = tf32(ELU(BN1(y))) -> next K slice of the spatial A operand in swizzled smem;
//                                       optionally store y / a1 to HBM for the (not yet fused) backward
//                         after the last channel: Y2 + bias -> HBM
//
// Y1 / A1 never leave the chip in the target configuration (y1 == a1 == nullptr).
#include "kernels.h"
#include <stdlib.h>
#include <string.h>

namespace eegb200 {

namespace {

constexpr int CTC_THREADS = 288;
constexpr int TILE_S = 3;                         // samples per tile
constexpr uint32_t KB_A = 128 * 32 * 4;           // [128 rows x 32 floats] K-major k-block: 16 KB
constexpr uint32_t KB_B = 48 * 32 * 4;            // [48 rows x 32 floats] K-major k-block: 6 KB (6 swizzle atoms)
constexpr int WS_RING = 4;
constexpr int N48 = 48;

// shared-memory map (offsets from the 1024-byte aligned base)
constexpr uint32_t OFF_IM = 0;                              // [2 buf][hi, lo] x 16 KB
constexpr uint32_t OFF_BC = OFF_IM + 4 * KB_A;              // conv weights hi, lo: 2 x 6 KB
constexpr uint32_t OFF_A1 = OFF_BC + 2 * KB_B;              // [2 buf][2 k-blocks] x 16 KB
constexpr uint32_t OFF_WS = OFF_A1 + 4 * KB_A;              // [4 ring][2 k-blocks] x 6 KB
constexpr uint32_t OFF_PS = OFF_WS + WS_RING * 2 * KB_B;    // pooled sums [3][208] + scan scratch [3][264] floats
constexpr uint32_t PS_LD = 208, CSX_LD = 264;
constexpr uint32_t OFF_TAB = OFF_PS + (3 * PS_LD + 3 * CSX_LD) * 4;      // BN scale / shift / bias tables [3][48] floats
constexpr uint32_t OFF_RED = OFF_TAB + 3 * 48 * 4;                        // statistics reduction [2][40] floats
constexpr uint32_t OFF_BAR = (OFF_RED + 80 * 4 + 7) & ~7u;                // mbarriers
constexpr int N_BARS = 2 + 2 + 2 + 2 + 2 + 2 + WS_RING + WS_RING + 1;
constexpr uint32_t OFF_TMEM = OFF_BAR + N_BARS * 8;
constexpr uint32_t CTC_SMEM = OFF_TMEM + 16 + 1024;                        // + alignment slack

enum { MODE_STATS = 0, MODE_APPLY = 1 };

struct ConvTcParams {
  const float* x3;          // [B*64, 256] token rows (channel c of sample b at row b*64 + c)
  const float* wt;          // [40, 25]
  const float* bt;          // [40]
  const float* mean_rstd;   // [2][40]   (MODE_APPLY)
  const float* gamma;       // [40]
  const float* beta;        // [40]
  const float* bs;          // [40] spatial conv bias
  float* y1;                // [B*36, 2520] or nullptr
  float* a1;                // [B*36, 2520] or nullptr
  float* y2;                // [B*36, 40]
  double* sums;             // [2][40]   (MODE_STATS)
  int B;
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a K-major SWIZZLE_128B k-block
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) {
  return (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ float tf32_fast(float x) {     // cvt.rna for finite values: add half an ulp of tf32, truncate
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float elu_fast(float z) {
  const float e = __expf(z) - 1.f;
  const float p = z * fmaf(z, fmaf(z, fmaf(z, 1.f / 24.f, 1.f / 6.f), 0.5f), 1.f);
  const float neg = z > -0.0625f ? p : e;
  return z > 0.f ? z : neg;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(CTC_THREADS, 1)
conv_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmWs, const ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  float* ps_all = reinterpret_cast<float*>(sm + OFF_PS);
  float* cs_all = ps_all + 3 * PS_LD;
  float* tab = reinterpret_cast<float*>(sm + OFF_TAB);          // [0]: scale, [1]: shift, [2]: conv bias (+ spatial bias at 40..)
  float* red = reinterpret_cast<float*>(sm + OFF_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
  uint64_t* tile_full = bars;            // [2] builders -> control
  uint64_t* tile_empty = bars + 2;       // [2] conv UMMAs done -> builders
  uint64_t* c1_full = bars + 4;          // [2] conv UMMAs done -> epilogue
  uint64_t* c1_empty = bars + 6;         // [2] epilogue read TMEM -> control        (4 arrivals)
  uint64_t* a1_full = bars + 8;          // [2] epilogue wrote the A1 slice -> control (4 arrivals)
  uint64_t* a1_empty = bars + 10;        // [2] spatial UMMAs done -> epilogue
  uint64_t* ws_full = bars + 12;         // [4] TMA -> control
  uint64_t* ws_empty = bars + 12 + WS_RING;   // [4] spatial UMMAs done -> control (TMA refill)
  uint64_t* y2_full = bars + 12 + 2 * WS_RING;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.x * TILE_S;
  const int ns = min(TILE_S, p.B - b0);                 // samples of this tile
  const int rows_valid = ns * N_POOL;

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tile_full[i], 1);
      mbar_init(&tile_empty[i], 1);
      mbar_init(&c1_full[i], 1);
      mbar_init(&c1_empty[i], 4);
      mbar_init(&a1_full[i], 4);
      mbar_init(&a1_empty[i], 1);
    }
    for (int i = 0; i < WS_RING; ++i) {
      mbar_init(&ws_full[i], 1);
      mbar_init(&ws_empty[i], 1);
    }
    mbar_init(y2_full, 1);
    mbar_fence_init();
    if (MODE == MODE_APPLY) tma_prefetch_desc(&tmWs);
  }
  // zero the im2col tiles (pad rows / pad columns stay zero for the whole kernel) and the A1 tiles
  for (uint32_t i = threadIdx.x * 16; i < 4 * KB_A; i += CTC_THREADS * 16) {
    *reinterpret_cast<float4*>(sm + OFF_IM + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(sm + OFF_A1 + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // conv weights / 51 as the B operand, hi and lo parts: rows k < 40 (48 with padding), columns t < 25 (32)
  for (int i = threadIdx.x; i < N48 * 32; i += CTC_THREADS) {
    const int k = i >> 5, t = i & 31;
    const float w = (k < N_FILT && t < K_TEMP) ? p.wt[k * K_TEMP + t] * (1.f / K_POOL) : 0.f;
    const float hi = tf32_fast(w), lo = tf32_fast(w - hi);
    const uint32_t off = sw128_off(k, t >> 2) + (uint32_t)(t & 3) * 4u;
    *reinterpret_cast<float*>(sm + OFF_BC + off) = hi;
    *reinterpret_cast<float*>(sm + OFF_BC + KB_B + off) = lo;
  }
  if (threadIdx.x < N48) {
    const int k = threadIdx.x;
    float sc = 0.f, sh = 0.f;
    if (MODE == MODE_APPLY && k < N_FILT) {
      sc = p.mean_rstd[N_FILT + k] * p.gamma[k];
      sh = p.beta[k] - p.mean_rstd[k] * sc;
    }
    tab[k] = sc;
    tab[48 + k] = sh;
    tab[96 + k] = k < N_FILT ? p.bt[k] : 0.f;
  }
  if (threadIdx.x < 80) red[threadIdx.x] = 0.f;
  if (warp == 8) {
    tmem_alloc(tmem_slot, 256);          // C1[0]: cols 0..47, C1[1]: cols 64..111, Y2: cols 128..175
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =============================== builders ===============================
    const int r = threadIdx.x;                         // tile row 0..127
    const int s_row = r / N_POOL, p_row = r % N_POOL;
    for (int c = 0; c < N_CH; ++c) {
      const int bi = c & 1;
      const uint32_t n = (uint32_t)(c >> 1);
      // pooled sums of the (up to) three token rows of this channel: warps 0..2, one row each
      if (warp < ns) {
        const float* xrow = p.x3 + ((size_t)(b0 + warp) * N_TOK + c) * D_PAD + 8 * lane;
        const float4 x0 = *reinterpret_cast<const float4*>(xrow), x1 = *reinterpret_cast<const float4*>(xrow + 4);
        float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int i = 1; i < 8; ++i) v[i] += v[i - 1];
        const float tot = v[7];
        float inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float nb = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += nb;
        }
        const float excl = inc - tot;
        float* cs = cs_all + warp * CSX_LD;               // C[k] = sum of the first k samples, k = 0..256
        cs[8 * lane] = excl;
#pragma unroll
        for (int i = 0; i < 7; ++i) cs[8 * lane + 1 + i] = v[i] + excl;
        if (lane == 31) cs[256] = inc;
        __syncwarp();
        float* ps = ps_all + warp * PS_LD;
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          const int u = lane + 32 * q;
          if (u < N_PSUM) ps[u] = cs[u + K_POOL] - cs[u];
        }
      }
      named_bar_sync(1, 128);
      mbar_wait(&tile_empty[bi], (n & 1u) ^ 1u);          // conv UMMAs of channel c-2 have consumed this buffer
      if (r < rows_valid) {
        const float* src = ps_all + s_row * PS_LD + 5 * p_row;
        uint8_t* hi_t = sm + OFF_IM + (uint32_t)bi * 2 * KB_A;
        uint8_t* lo_t = hi_t + KB_A;
#pragma unroll
        for (int ch = 0; ch < 7; ++ch) {                  // chunk 6 = tap 24 + zeros; chunk 7 stays zero
          float a[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) a[q] = (ch * 4 + q) < K_TEMP ? src[ch * 4 + q] : 0.f;
          float4 h, l;
          h.x = tf32_fast(a[0]); h.y = tf32_fast(a[1]); h.z = tf32_fast(a[2]); h.w = tf32_fast(a[3]);
          l.x = tf32_fast(a[0] - h.x); l.y = tf32_fast(a[1] - h.y); l.z = tf32_fast(a[2] - h.z); l.w = tf32_fast(a[3] - h.w);
          const uint32_t off = sw128_off(r, ch);
          *reinterpret_cast<float4*>(hi_t + off) = h;
          *reinterpret_cast<float4*>(lo_t + off) = l;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (threadIdx.x == 0) mbar_arrive(&tile_full[bi]);
    }
  } else if (warp < 8) {
    // =============================== epilogue ===============================
    const int q = warp - 4;                              // TMEM lane quarter (== warp % 4)
    const int r = q * 32 + lane;                         // tile row
    const bool valid = r < rows_valid;
    const size_t grow = (size_t)b0 * N_POOL + r;         // global (b, p) row
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float s1[N_FILT], s2[N_FILT];
    if (MODE == MODE_STATS) {
#pragma unroll
      for (int k = 0; k < N_FILT; ++k) s1[k] = s2[k] = 0.f;
    }
    for (int c = 0; c < N_CH; ++c) {
      const int bi = c & 1;
      const uint32_t n = (uint32_t)(c >> 1);
      mbar_wait(&c1_full[bi], n & 1u);
      tc_fence_after();
      float y[48];
      tmem_ld_32x32(tmem_base + lane_addr + (uint32_t)(bi * 64), y);
      tmem_ld_32x16(tmem_base + lane_addr + (uint32_t)(bi * 64 + 32), y + 32);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&c1_empty[bi]);         // the accumulator may be overwritten by channel c+2
#pragma unroll
      for (int k = 0; k < N_FILT; ++k) y[k] += tab[96 + k];
      if (MODE == MODE_STATS) {
        if (valid) {
#pragma unroll
          for (int k = 0; k < N_FILT; ++k) { s1[k] += y[k]; s2[k] = fmaf(y[k], y[k], s2[k]); }
        }
      } else {
        if (valid && p.y1 != nullptr) {
          float4* dst = reinterpret_cast<float4*>(p.y1 + grow * K_SPAT + c * N_FILT);
#pragma unroll
          for (int j = 0; j < 10; ++j) dst[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
        }
#pragma unroll
        for (int k = 0; k < N_FILT; ++k) y[k] = tf32_fast(elu_fast(fmaf(y[k], tab[k], tab[48 + k])));
        if (valid && p.a1 != nullptr) {
          float4* dst = reinterpret_cast<float4*>(p.a1 + grow * K_SPAT + c * N_FILT);
#pragma unroll
          for (int j = 0; j < 10; ++j) dst[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
        }
        // K slice of the spatial A operand: k 0..31 -> k-block 0, k 32..39 -> first two chunks of k-block 1
        mbar_wait(&a1_empty[bi], (n & 1u) ^ 1u);         // spatial UMMAs of channel c-2 are done with this buffer
        uint8_t* a1t = sm + OFF_A1 + (uint32_t)bi * 2 * KB_A;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          *reinterpret_cast<float4*>(a1t + sw128_off(r, ch)) = make_float4(y[4 * ch], y[4 * ch + 1], y[4 * ch + 2], y[4 * ch + 3]);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
          *reinterpret_cast<float4*>(a1t + KB_A + sw128_off(r, ch)) =
              make_float4(y[32 + 4 * ch], y[33 + 4 * ch], y[34 + 4 * ch], y[35 + 4 * ch]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a1_full[bi]);
      }
    }
    if (MODE == MODE_STATS) {
      // per-filter sums over the 32 rows of this warp, then shared + one double atomic per filter per CTA
#pragma unroll
      for (int k = 0; k < N_FILT; ++k) {
        const float a = warp_sum(s1[k]), b = warp_sum(s2[k]);
        if (lane == 0) { atomicAdd(&red[k], a); atomicAdd(&red[N_FILT + k], b); }
      }
      named_bar_sync(2, 128);
      const int k = threadIdx.x - 128;
      if (k < 2 * N_FILT) atomicAdd(&p.sums[k], (double)red[k]);
    } else {
      // ---- y2 = accumulated spatial product + bias ----
      mbar_wait(y2_full, 0);
      tc_fence_after();
      float o[48];
      tmem_ld_32x32(tmem_base + lane_addr + 128u, o);
      tmem_ld_32x16(tmem_base + lane_addr + 160u, o + 32);
      if (valid) {
        float4* dst = reinterpret_cast<float4*>(p.y2 + grow * N_FILT);
#pragma unroll
        for (int j = 0; j < 10; ++j)
          dst[j] = make_float4(o[4 * j] + p.bs[4 * j], o[4 * j + 1] + p.bs[4 * j + 1], o[4 * j + 2] + p.bs[4 * j + 2],
                               o[4 * j + 3] + p.bs[4 * j + 3]);
      }
    }
  } else if (lane == 0) {
    // =============================== control: TMA + UMMA issue ===============================
    constexpr uint32_t idesc = umma_idesc_tf32(128, N48, 0, 0);
    const uint32_t im = smem_u32(sm + OFF_IM), bc = smem_u32(sm + OFF_BC), a1s = smem_u32(sm + OFF_A1), wss = smem_u32(sm + OFF_WS);
    auto load_ws = [&](int c) {
      const int wi = c & (WS_RING - 1);
      const uint32_t n = (uint32_t)(c / WS_RING);
      mbar_wait(&ws_empty[wi], (n & 1u) ^ 1u);
      mbar_arrive_expect_tx(&ws_full[wi], 2 * KB_B);
      tma_load_2d(&tmWs, &ws_full[wi], sm + OFF_WS + (uint32_t)wi * 2 * KB_B, 0, c * N48);
      tma_load_2d(&tmWs, &ws_full[wi], sm + OFF_WS + (uint32_t)wi * 2 * KB_B + KB_B, 32, c * N48);
    };
    auto spatial = [&](int c) {
      const int bi = c & 1, wi = c & (WS_RING - 1);
      mbar_wait(&a1_full[bi], (uint32_t)(c >> 1) & 1u);
      mbar_wait(&ws_full[wi], (uint32_t)(c / WS_RING) & 1u);
      tc_fence_after();
      const uint32_t a = a1s + (uint32_t)bi * 2 * KB_A, w = wss + (uint32_t)wi * 2 * KB_B;
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) {                   // K = 40: four 8-steps of k-block 0 and the first of k-block 1
        const uint32_t ao = kk < 4 ? (uint32_t)kk * 32u : KB_A, wo = kk < 4 ? (uint32_t)kk * 32u : KB_B;
        tc_mma_tf32(tmem_base + 128u, umma_smem_desc(a + ao, 16, 1024, UMMA_LAYOUT_SW128),
                    umma_smem_desc(w + wo, 16, 1024, UMMA_LAYOUT_SW128), idesc, (c > 0 || kk > 0) ? 1u : 0u);
      }
      tc_commit(&a1_empty[bi]);
      tc_commit(&ws_empty[wi]);
    };
    // weights of channel c+2 are requested while channel c is processed: their ring slot was last read by the spatial
    // UMMAs of channel c-2, issued one iteration earlier
    if (MODE == MODE_APPLY) {
      load_ws(0);
      load_ws(1);
    }
    for (int c = 0; c < N_CH; ++c) {
      const int bi = c & 1;
      const uint32_t n = (uint32_t)(c >> 1);
      mbar_wait(&tile_full[bi], n & 1u);
      mbar_wait(&c1_empty[bi], (n & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t hi = im + (uint32_t)bi * 2 * KB_A, lo = hi + KB_A;
      const uint32_t d = tmem_base + (uint32_t)(bi * 64);
      // 3xTF32: lo.hi + hi.lo + hi.hi (small terms first)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        tc_mma_tf32(d, umma_smem_desc(lo + kk * 32, 16, 1024, UMMA_LAYOUT_SW128),
                    umma_smem_desc(bc + kk * 32, 16, 1024, UMMA_LAYOUT_SW128), idesc, kk > 0 ? 1u : 0u);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        tc_mma_tf32(d, umma_smem_desc(hi + kk * 32, 16, 1024, UMMA_LAYOUT_SW128),
                    umma_smem_desc(bc + KB_B + kk * 32, 16, 1024, UMMA_LAYOUT_SW128), idesc, 1u);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        tc_mma_tf32(d, umma_smem_desc(hi + kk * 32, 16, 1024, UMMA_LAYOUT_SW128),
                    umma_smem_desc(bc + kk * 32, 16, 1024, UMMA_LAYOUT_SW128), idesc, 1u);
      tc_commit(&tile_empty[bi]);
      tc_commit(&c1_full[bi]);
      if (MODE == MODE_APPLY) {
        if (c + 2 < N_CH) load_ws(c + 2);
        if (c > 0) spatial(c - 1);                       // one channel behind: its epilogue ran while these MMAs were issued
      }
    }
    if (MODE == MODE_APPLY) {
      spatial(N_CH - 1);
      tc_commit(y2_full);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 256);
}

// tsconv.4.weight [j][k][c] -> [c][48 rows j][64 floats k], TF32-rounded, zero padded: B operand of the spatial UMMA
__global__ void pack_ws_tc_kernel(const float* __restrict__ ws, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N_CH * N48 * 64) return;
  const int k = idx & 63, j = (idx >> 6) % N48, c = idx / (N48 * 64);
  out[idx] = (j < N_FILT && k < N_FILT) ? tf32_rn(ws[(j * N_FILT + k) * N_CH + c]) : 0.f;
}

}  // namespace

int conv_tc_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("EEGB200_CONV_TC");
    on = (e && e[0] == '1') ? 1 : 0;        // experimental: off unless asked for
  }
  return on;
}
size_t conv_tc_ws_floats() { return (size_t)N_CH * N48 * 64; }

// BatchNorm1 batch statistics of the temporal conv output without materialising it (kernel F1)
int conv_tc_stats(const float* x3, const float* wt, const float* bt, double* sums, int B, cudaStream_t s) {
  ProfScope _ps("conv_tc_stats", s, (double)B * 63 * 36 * 40 * 50.0, (double)B * 63 * 1000.0);
  static PerDeviceOnce once;
  if (once.first()) {
    EEG_CUDA_OK(cudaFuncSetAttribute(conv_tc_fwd_kernel<MODE_STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CTC_SMEM));
  }
  ConvTcParams p{x3, wt, bt, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, sums, B};
  CUtensorMap dummy;
  memset(&dummy, 0, sizeof(dummy));
  conv_tc_fwd_kernel<MODE_STATS><<<cdiv(B, TILE_S), CTC_THREADS, CTC_SMEM, s>>>(dummy, p);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// temporal conv + pool + BatchNorm1 + ELU + spatial conv in one kernel (kernel F2); y1 / a1 may be nullptr
int conv_tc_apply(const float* x3, const float* wt, const float* bt, const float* mean_rstd, const float* gamma,
                  const float* beta, const float* ws, const float* bs, float* ws_packed, float* y1, float* a1, float* y2,
                  int B, cudaStream_t s) {
  ProfScope _ps("conv_tc_apply", s, (double)B * 36 * (63 * 40 * 50.0 + 2520 * 80.0),
                (double)B * (63 * 1000.0 + 36 * 160.0 + ((y1 ? 1 : 0) + (a1 ? 1 : 0)) * 36 * 2520 * 4.0));
  pack_ws_tc_kernel<<<cdiv(N_CH * N48 * 64, 256), 256, 0, s>>>(ws, ws_packed);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  CUtensorMap tw;
  int d3 = 0;
  EEG_TRY(gemm_make_tmap(&tw, GemmOperand{ws_packed, 64, 0}, N_CH * N48, 64, N48, &d3));
  static PerDeviceOnce once;
  if (once.first()) {
    EEG_CUDA_OK(cudaFuncSetAttribute(conv_tc_fwd_kernel<MODE_APPLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CTC_SMEM));
  }
  ConvTcParams p{x3, wt, bt, mean_rstd, gamma, beta, bs, y1, a1, y2, nullptr, B};
  conv_tc_fwd_kernel<MODE_APPLY><<<cdiv(B, TILE_S), CTC_THREADS, CTC_SMEM, s>>>(tw, p);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace eegb200

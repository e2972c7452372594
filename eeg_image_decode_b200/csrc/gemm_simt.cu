// Exact-fp32 SIMT GEMM with the same interface/epilogue as the tcgen05 kernel.
// Verification only: tests use it (a) as the on-device cross-check for the tensor-core kernel and
// (b) to run the whole path in exact fp32 so that every non-GEMM kernel can be compared with the
// oracle at 1e-5.  It is never selected unless eegb200_set_gemm_backend(1) is called.
#include "gemm.h"
#include <atomic>
#include <stdarg.h>
#include <string.h>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

namespace eegb200 {

// ---------------- process-wide small state: last error, launch counters, backend switch ----------------
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches += n; }
long long total_launch_count() { return g_launches.load(); }

static std::atomic<int> g_backend{GEMM_BACKEND_TCGEN05};
void gemm_set_backend(int b) { g_backend = b; }
int gemm_get_backend() { return g_backend.load(); }
int tf32_rounding() { return g_backend.load() == GEMM_BACKEND_TCGEN05 ? 1 : 0; }
static const char* intern_name(const std::string& n);
int gemm_launch(const GemmArgs& g, cudaStream_t stream) {
  const char* nm = "gemm";
  if (prof_enabled()) {
    char tmp[128];
    snprintf(tmp, sizeof(tmp), "gemm_tf32 M=%d N=%d K=%d %c%c%s", g.M, g.N, g.K, g.A.mn_major ? 'm' : 'k',
             g.B.mn_major ? 'n' : 'k', g.split_k > 1 ? " splitK" : "");
    nm = intern_name(tmp);
  }
  ProfScope _ps(nm, stream, 2.0 * g.M * g.N * g.K, 4.0 * ((double)g.M * g.K + (double)g.N * g.K + (double)g.M * g.N));
  return g_backend.load() == GEMM_BACKEND_SIMT_FP32 ? gemm_launch_simt(g, stream) : gemm_launch_tcgen05(g, stream);
}

// ---------------- event profiler ----------------
struct ProfRec { std::string name; cudaEvent_t a, b; double flops, bytes; };
static std::atomic<int> g_prof{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;
void prof_enable(int on) { g_prof = on; }
int prof_enabled() { return g_prof.load(); }
ProfScope::ProfScope(const char* name, cudaStream_t stream, double flops, double bytes)
    : s(stream), name_(name), flops_(flops), bytes_(bytes) {
  if (!g_prof.load()) return;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a, s);
}
ProfScope::~ProfScope() {
  if (!a) return;
  cudaEventRecord(b, s);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_recs.push_back({name_, a, b, flops_, bytes_});
}
static const char* intern_name(const std::string& n) {
  static std::set<std::string> names;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  return names.insert(n).first->c_str();
}
int prof_report(char* buf, size_t cap) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  struct Agg { double ms = 0, flops = 0, bytes = 0; long n = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : g_prof_recs) {
    cudaEventSynchronize(r.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    Agg& g = agg[r.name];
    g.ms += ms; g.n += 1; g.flops += r.flops; g.bytes += r.bytes;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof_recs.clear();
  std::string out = "{";
  bool first = true;
  for (auto& kv : agg) {
    char tmp[512];
    snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"ms\": %.6f, \"n\": %ld, \"flops\": %.6e, \"bytes\": %.6e}", first ? "" : ", ",
             kv.first.c_str(), kv.second.ms, kv.second.n, kv.second.flops, kv.second.bytes);
    out += tmp;
    first = false;
  }
  out += "}";
  if (out.size() + 1 > cap) return 4;
  memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

// ---------------- kernel ----------------
static constexpr int ST = 64;   // tile
static constexpr int SK = 16;

__device__ __forceinline__ float ld_op(const GemmOperand& o, int i, int k, int rows, int K) {
  if (i >= rows || k >= K) return 0.f;
  return o.mn_major ? o.ptr[(size_t)k * o.ld + i] : o.ptr[(size_t)i * o.ld + k];
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(int M, int N, int K, GemmOperand A, GemmOperand B, Epilogue e,
                                                        int k_per_split) {
  __shared__ float sA[SK][ST + 1];
  __shared__ float sB[SK][ST + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * ST, n0 = blockIdx.x * ST;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  float acc[4][4] = {};
  for (int k0 = kbeg; k0 < kend; k0 += SK) {
    for (int idx = threadIdx.x; idx < ST * SK; idx += 256) {
      int i, k;
      if (A.mn_major) { i = idx % ST; k = idx / ST; } else { k = idx % SK; i = idx / SK; }
      sA[k][i] = (k0 + k < kend) ? ld_op(A, m0 + i, k0 + k, M, K) : 0.f;
      if (B.mn_major) { i = idx % ST; k = idx / ST; } else { k = idx % SK; i = idx / SK; }
      sB[k][i] = (k0 + k < kend) ? ld_op(B, n0 + i, k0 + k, N, K) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[k][ty * 4 + i]; b[i] = sB[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = m0 + ty * 4 + i, c = n0 + tx * 4 + j;
      if (r < M && c < N) epi_store(e, r, c, epi_value(e, r, c, acc[i][j]));
    }
}

int gemm_launch_simt(const GemmArgs& g, cudaStream_t stream) {
  EEG_REQUIRE(g.M > 0 && g.N > 0, "gemm: bad shape");
  EEG_REQUIRE(g.split_k <= 1 || g.epi.store_mode == EPI_ATOMIC, "gemm: split-K needs the atomic store mode");
  int split = g.split_k < 1 ? 1 : g.split_k;
  int kps = cdiv(cdiv(g.K, split), SK) * SK;
  if (kps < SK) kps = SK;
  split = cdiv(g.K, kps);
  if (split < 1) split = 1;
  dim3 grid(cdiv(g.N, ST), cdiv(g.M, ST), split);
  gemm_simt_kernel<<<grid, 256, 0, stream>>>(g.M, g.N, g.K, g.A, g.B, g.epi, kps);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace eegb200

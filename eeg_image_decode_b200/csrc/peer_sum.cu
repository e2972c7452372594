// One-shot all-reduce (sum) of a small fp64 vector over NVLink peer memory -- the SyncBatchNorm statistics exchange of the
// data-parallel step (2 x 40 doubles, four times per step).  NCCL needs ~20-30 us per such call at 8 GPUs (launch +
// protocol latency); here every rank PUSHES its vector into a slot of every peer's buffer, raises a flag there, waits for
// the W flags in its own buffer and adds the W slots in rank order (bit-identical result on every rank): one tiny kernel,
// one NVLink round trip.
//
// Symmetric buffer (same layout on every rank; the peers' base addresses come from the caller, e.g.
// torch.distributed._symmetric_memory):   [2 sets][16 slots][PEER_MAX_N doubles]  then  [2 sets][16] uint64 flags.
// Calls alternate between the two sets (set = sequence number & 1): a rank that runs ahead writes set s+1 while a slow
// peer may still be adding set s, and it cannot reach s+2 before that peer has raised its s+1 flag, i.e. finished s.
// The sequence number lives in device memory and is advanced by the kernel itself, so a captured CUDA graph replays it.
#include "../../include/eegdecode_b200.h"
#include "kernels.h"

namespace eegb200 {
namespace {

constexpr int PEER_MAX_W = 16;
constexpr int PEER_MAX_N = 256;
constexpr size_t PEER_DATA_BYTES = (size_t)2 * PEER_MAX_W * PEER_MAX_N * sizeof(double);
constexpr size_t PEER_BYTES = PEER_DATA_BYTES + (size_t)2 * PEER_MAX_W * sizeof(unsigned long long);

struct PeerPtrs { void* p[PEER_MAX_W]; };

__device__ __forceinline__ void st_release_sys(unsigned long long* addr, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* addr) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
  return v;
}

__global__ void peer_sum_f64_kernel(double* __restrict__ data, int n, PeerPtrs peers, int rank, int world,
                                    unsigned long long* __restrict__ seq_dev, int* __restrict__ error_flag) {
  __shared__ int s_fail;
  const unsigned long long seq = *seq_dev + 1;          // every thread reads it before thread 0 advances it (barriers below)
  const int set = (int)(seq & 1ull);
  if (threadIdx.x == 0) s_fail = 0;
  // push: my vector into slot `rank` of every rank's buffer (my own included)
  for (int p = 0; p < world; ++p) {
    double* dst = reinterpret_cast<double*>(peers.p[p]) + ((size_t)set * PEER_MAX_W + rank) * PEER_MAX_N;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = data[i];
  }
  __syncthreads();
  if ((int)threadIdx.x < world) {
    __threadfence_system();
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(peers.p[threadIdx.x]) + PEER_DATA_BYTES) +
                               set * PEER_MAX_W + rank;
    st_release_sys(flag, seq);
    // wait for rank threadIdx.x's vector in MY buffer
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(reinterpret_cast<uint8_t*>(peers.p[rank]) + PEER_DATA_BYTES) +
                                     set * PEER_MAX_W + threadIdx.x;
    long long spins = 0;
    while (ld_acquire_sys(mine) < seq) {
      if (++spins > (1ll << 26)) { s_fail = 1; break; }      // ~ seconds: a rank is missing; report instead of hanging
      __nanosleep(20);
    }
  }
  __syncthreads();
  if (s_fail) {
    if (threadIdx.x == 0) { *error_flag = 1; *seq_dev = seq; }
    return;
  }
  const double* my = reinterpret_cast<const double*>(peers.p[rank]) + (size_t)set * PEER_MAX_W * PEER_MAX_N;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int p = 0; p < world; ++p) s += my[(size_t)p * PEER_MAX_N + i];
    data[i] = s;
  }
  if (threadIdx.x == 0) *seq_dev = seq;
}

}  // namespace
}  // namespace eegb200

using namespace eegb200;

extern "C" size_t eegb200_peer_sum_buffer_bytes(void) { return PEER_BYTES; }

extern "C" int eegb200_peer_sum_f64(double* data, int n, const void* const* peer_buffers, int rank, int world,
                                    unsigned long long* seq_dev, int* error_flag_dev, void* stream) {
  EEG_REQUIRE(data && peer_buffers && seq_dev && error_flag_dev, "peer_sum: null pointer");
  EEG_REQUIRE(n > 0 && n <= PEER_MAX_N && world >= 1 && world <= PEER_MAX_W && rank >= 0 && rank < world,
              "peer_sum: bad arguments n=%d rank=%d world=%d", n, rank, world);
  PeerPtrs pp;
  for (int i = 0; i < PEER_MAX_W; ++i) pp.p[i] = i < world ? const_cast<void*>(peer_buffers[i]) : nullptr;
  for (int i = 0; i < world; ++i) EEG_REQUIRE(pp.p[i] != nullptr, "peer_sum: null peer buffer %d", i);
  peer_sum_f64_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(data, n, pp, rank, world, seq_dev, error_flag_dev);
  EEG_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

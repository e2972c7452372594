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

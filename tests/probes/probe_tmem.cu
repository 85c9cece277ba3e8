// Hardware probes (not on the product path): TMEM -> register bandwidth of tcgen05.ld, used to
// size the epilogues of mlp.cu.  Built into its own library (loner_b200.build.build_probe -> tests/probes/libloner_probe.so), NOT into the
// product .so; run by tests/gpu_probe.py, results recorded in DESIGN.md.
#include "common.cuh"
#include "sm100.cuh"

namespace loner {
using namespace sm100;

// mode: number of 32-column loads issued back to back before one tcgen05.wait::ld (1, 2, 4, 8)
__global__ void __launch_bounds__(512, 1) probe_tmem_kernel(int iters, int mode, long long* cycles, unsigned* sink) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&s_tmem));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int g = 0; g < 16; g += mode) {
      uint32_t v[8][32];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < mode) tmem_ld32(base + ((g + k) & 15) * 32, v[k]);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < mode) acc ^= v[k][0] ^ v[k][31];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}
}  // namespace loner

// bytes read per block = warps * iters * 16 * 32 columns * 32 lanes * 4 B
extern "C" int loner_probe_tmem(int warps, int iters, int mode, long long* cycles, unsigned* sink, void* stream) {
  if (warps < 1 || warps > 16 || !cycles || !sink || (mode != 1 && mode != 2 && mode != 4 && mode != 8)) return LONER_E_BAD_ARG;
  loner::probe_tmem_kernel<<<148, warps * 32, 0, (cudaStream_t)stream>>>(iters, mode, cycles, sink);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

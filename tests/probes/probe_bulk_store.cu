// Hardware probe (NOT part of the product library): aggregate HBM write bandwidth of cp.async.bulk shared->global
// copies as a function of the copy size and of how many copies one SM keeps in flight.  Sizes the stash copies of
// mlp.cu (one CTA per SM, 64 KB images, at most two in flight).  Built by loner_b200.build.build_probe().
#include "common.cuh"
#include "sm100.cuh"

namespace loner {
using namespace sm100;

// every CTA writes `iters` chunks of `bytes` from one shared-memory buffer to its own region of `dst`,
// keeping at most `depth` bulk groups in flight (wait_group.read depth-1 before each new issue)
// `load_bytes` > 0: a second thread concurrently streams that many bytes per iteration from an L2-resident image
// (`src`, 416 KB, the same for every CTA - the weight ring of the pipelined kernels) into shared memory.
template <int kDepth>
__global__ void __launch_bounds__(128, 1) probe_bulk_store_kernel(uint8_t* dst, int bytes, int iters, long long* cycles,
                                                                  const uint8_t* src, int load_bytes) {
  extern __shared__ __align__(1024) uint8_t buf[];
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(buf)[i] = i;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  fence_async_smem();
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 32 && load_bytes > 0) {
    const uint32_t b = smem_u32(&bar), ld = smem_u32(buf) + 65536u;
    for (int it = 0; it < iters; ++it) {
      mbar_expect_tx(b, (uint32_t)load_bytes);
      for (int o = 0; o < load_bytes; o += 16384)
        bulk_g2s(ld + o, src + ((size_t)(it * load_bytes + o) % (416u * 1024u - 16384u)) / 16384 * 16384, 16384u, b);
      mbar_wait(b, (uint32_t)(it & 1));
    }
  }
  if (threadIdx.x == 0) {
    uint8_t* mine = dst + (size_t)blockIdx.x * (size_t)iters * (size_t)bytes;
    for (int it = 0; it < iters; ++it) {
      bulk_s2g(mine + (size_t)it * bytes, smem_u32(buf), (uint32_t)bytes);
      bulk_commit();
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kDepth - 1) : "memory");
    }
    bulk_wait0();
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}
}  // namespace loner

extern "C" int loner_probe_bulk_store(void* dst, int bytes, int iters, int depth, int blocks, long long* cycles,
                                      const void* src, int load_bytes, void* stream) {
  if (!dst || !cycles || bytes <= 0 || bytes > 64 * 1024 || (bytes & 15) || load_bytes < 0 || load_bytes > 128 * 1024 ||
      (load_bytes % 16384) || (load_bytes && !src))
    return LONER_E_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(D)                                                                                                  \
  cudaFuncSetAttribute(loner::probe_bulk_store_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
  loner::probe_bulk_store_kernel<D><<<blocks, 128, 65536 + 131072, st>>>((uint8_t*)dst, bytes, iters, cycles, (const uint8_t*)src, load_bytes);
  switch (depth) {
    case 1: LAUNCH(1) break;
    case 2: LAUNCH(2) break;
    case 4: LAUNCH(4) break;
    case 8: LAUNCH(8) break;
    default: return LONER_E_BAD_ARG;
  }
#undef LAUNCH
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

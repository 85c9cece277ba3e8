// Hardware probe (NOT part of the product library): L2 throughput of the two access patterns of the hash-grid sigma
// head - random 8-byte vector reductions (red.global.add.v2.f32) into an fp32 table and random 4-byte gathers from a
// half2 table, both L2 resident (the shipped 16 x 2^18 table: 33.5 MB of gradients, 16.8 MB of features).
// These are the denominators bench.py quotes the hash kernels against.  Built by loner_b200.build.build_probe().
#include "common.cuh"

namespace loner {
__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__global__ void __launch_bounds__(256) probe_atomics_kernel(float2* table, uint32_t entries, int per_thread) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < per_thread; ++i) {
    const uint32_t idx = mix(tid * 9781u + i * 6271u + 17u) % entries;
    atomicAdd(table + idx, make_float2(1.0f, 0.5f));
  }
}
__global__ void __launch_bounds__(256) probe_gather_kernel(const uint32_t* table, uint32_t entries, int per_thread, uint32_t* sink) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
#pragma unroll 8
  for (int i = 0; i < per_thread; ++i) acc ^= __ldg(table + mix(tid * 9781u + i * 6271u + 17u) % entries);
  if (acc == 0x12345u) sink[0] = acc;
}
}  // namespace loner

extern "C" int loner_probe_atomics(void* table, unsigned entries, int per_thread, int blocks, void* stream) {
  loner::probe_atomics_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((float2*)table, entries, per_thread);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}
extern "C" int loner_probe_gather(const void* table, unsigned entries, int per_thread, int blocks, void* sink, void* stream) {
  loner::probe_gather_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)table, entries, per_thread, (uint32_t*)sink);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

// Hardware probe (NOT part of the product library): sustained rate of tcgen05.mma kind::f16, M = 128, N = 256, K = 16
// (the instruction of the pipelined kernels: 4 per 64-wide K chunk), issued back to back from one converged warp with
// both operands in shared memory - alone, next to a stream of L2 -> shared bulk copies (the weight ring) and next to
// 128-bit shared-memory stores from 8 warps (the epilogue).  Answers whether the 3200 - 3400 clk per tile and layer of
// the forward / dgrad kernels (16 such MMAs, 2048 clk at the nominal rate) is the tensor pipe's own limit or
// interference in the shared-memory data pipe.  Built by loner_b200.build.build_probe().
#include "common.cuh"
#include "sm100.cuh"

namespace loner {
using namespace sm100;

__device__ __forceinline__ uint32_t ld_acquire_cta_probe(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// smem: A [128 x 64] fp16 image 16 KB | B [256 x 64] 32 KB | ring 3 x 32 KB | store target 64 KB
__global__ void __launch_bounds__(320, 1) probe_mma_kernel(int iters, int load_kb_per_chunk, int store_kb_per_chunk,
                                                           int tmem_lds_per_chunk, int mode, const uint8_t* src, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sA = smem_u32(smem), sB = sA + 16384, sRing = sB + 32768, sSt = sRing + 98304;
  for (int i = tid; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // 1.0
  if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_mbar_init();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t done = smem_u32(&bars[0]);
  const long long t0 = clock64();
  if (warp == 1) {
    // chunks of 4 MMAs (64 of K), a commit per chunk like the kernels; accumulate into the same 256 columns
    constexpr uint32_t idesc = make_idesc_f16(128, 256, 0, 1);
    // mode bit 0: a commit per chunk (to a barrier nobody waits on), like the kernels' w_empty commits;
    // mode bit 2: alternate between the two accumulators every 4 chunks (tile X / tile Y); mode bit 3: B from the ring slots
    const uint32_t scratch_bar = smem_u32(&bars[7]);
    for (int it = 0; it < iters; ++it) {
      const uint32_t acc = tmem + ((mode & 4) ? (uint32_t)((it >> 2) & 1) * 256u : 0u);
      const uint32_t b = (mode & 8) ? sRing + (uint32_t)(it % 3) * 32768u : sB;
      umma_f16_x4_warp<2, 128>(acc, desc_lo_sw128(sA, 16), desc_hi_sw128(1024), desc_lo_sw128(b, 8192), desc_hi_sw128(1024),
                               idesc, it > 1 ? 1u : 0u, it == iters - 1 ? done : ((mode & 1) ? scratch_bar : 0u));
    }
    mbar_wait_warp(done, 0);
    if (lane == 0) cycles[gridDim.x + blockIdx.x] = clock64() - t0;      // the MMA stream's own duration
  } else if (warp == 0 && load_kb_per_chunk > 0) {
    // weight-ring traffic: load_kb per chunk-time into a 3-slot ring, at most 3 copies in flight
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)load_kb_per_chunk * 1024u;
      for (int it = 0; it < iters; ++it) {
        const uint32_t b = smem_u32(&bars[1 + it % 3]);
        while (clock64() - t0 < (long long)it * 512) {}        // paced: one chunk per nominal chunk time (4 x 128 clk)
        if (it >= 3) mbar_wait(b, (uint32_t)((it / 3 - 1) & 1));
        mbar_expect_tx(b, bytes);
        bulk_g2s(sRing + (it % 3) * 32768u, src + (size_t)(it % 12) * 32768, bytes, b);
      }
      for (int it = iters > 3 ? iters - 3 : 0; it < iters; ++it) mbar_wait(smem_u32(&bars[1 + it % 3]), (uint32_t)((it / 3) & 1));
    }
  } else if (warp >= 2 && (mode & 2)) {
    // mode bit 1: the eight epilogue warps spin on an mbarrier that never completes (the kernels' acc_full polling)
    const uint32_t never = smem_u32(&bars[6]);
    while (clock64() - t0 < (long long)iters * 512) {
      if (mbar_try_wait(never, 0)) break;
    }
  } else if (warp >= 2 && tmem_lds_per_chunk > 0) {
    // epilogue traffic on the TMEM side: every warp reads 32 columns of its lane quarter of the OTHER accumulator
    // (columns 256 ..) tmem_lds_per_chunk times per chunk time (the kernels: one such load per warp and chunk time),
    // and optionally stores what it read
    const uint32_t acc = tmem + 256u + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp - 2) >> 2) * 128);
    const uint32_t dst = sSt + (uint32_t)(tid - 64) * 16u;
    uint32_t sink = 0;
    for (int it = 0; it < iters; ++it) {
      while (clock64() - t0 < (long long)it * 512) {}
      for (int r = 0; r < tmem_lds_per_chunk; ++r) {
        uint32_t v[32];
        tmem_ld32(acc + (uint32_t)((it + r) & 3) * 32u, v);
        tmem_ld_wait();
        if (store_kb_per_chunk > 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) sts128(dst + (uint32_t)((it * 4 + k) & 15) * 4096u, v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) sink ^= v[k];
        }
      }
    }
    if (sink == 0x12345678u) cycles[0] = 0;
  } else if (warp >= 2 && store_kb_per_chunk > 0) {
    // epilogue traffic: store_kb per chunk-time as 128-bit stores from 8 warps (256 threads x 16 B = 4 KB per round)
    const int rounds = store_kb_per_chunk / 4;
    const uint32_t dst = sSt + (uint32_t)(tid - 64) * 16u;
    for (int it = 0; it < iters; ++it) {
      while (clock64() - t0 < (long long)it * 512) {}          // paced like the loads
      for (int r = 0; r < rounds; ++r) sts128(dst + (uint32_t)((it * rounds + r) & 15) * 4096u, it, r, tid, 0u);
    }
  }
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}
}  // namespace loner

extern "C" int loner_probe_mma(int iters, int load_kb_per_chunk, int store_kb_per_chunk, int tmem_lds_per_chunk, int mode,
                               const void* src, int blocks, long long* cycles, void* stream) {
  if (iters <= 0 || !cycles || load_kb_per_chunk < 0 || load_kb_per_chunk > 32 || store_kb_per_chunk < 0 ||
      (store_kb_per_chunk % 4) || (load_kb_per_chunk && !src))
    return LONER_E_BAD_ARG;
  const int smem = 16384 + 32768 + 98304 + 65536;
  cudaFuncSetAttribute(loner::probe_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  loner::probe_mma_kernel<<<blocks, 320, smem, (cudaStream_t)stream>>>(iters, load_kb_per_chunk, store_kb_per_chunk,
                                                                       tmem_lds_per_chunk, mode, (const uint8_t*)src, cycles);
  return cudaGetLastError() == cudaSuccess ? LONER_OK : LONER_E_LAUNCH;
}

// ---- latency of the synchronisation primitives the MMA issuer executes between chunks (one warp, dependent chain):
// what[0] mbarrier.try_wait on a completed phase, [1] mbarrier.test_wait, [2] ld.acquire.cta.shared, [3] ld.volatile.shared,
// [4] tcgen05.fence::after_thread_sync, [5] try_wait + __syncwarp (= mbar_wait_warp)
namespace loner {
__global__ void probe_sync_kernel(int iters, long long* out) {
  __shared__ uint64_t bar;
  __shared__ uint32_t word;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); word = 1u; }
  __syncthreads();
  if (threadIdx.x == 0) mbar_arrive(smem_u32(&bar));       // phase 0 completes
  __syncthreads();
  const uint32_t b = smem_u32(&bar), w = smem_u32(&word);
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) acc += mbar_try_wait(b, 0) ? 1u : 0u;
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0);
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0u) : "memory");
    acc += ok;
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[1] = (t1 - t0);
  t0 = clock64();
  for (int i = 0; i < iters; ++i) acc += ld_acquire_cta_probe(w + (acc & 0u));
  t1 = clock64();
  if (threadIdx.x == 0) out[2] = (t1 - t0);
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(w + (acc & 0u)) : "memory");
    acc += v;
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[3] = (t1 - t0);
  t0 = clock64();
  for (int i = 0; i < iters; ++i) tc_fence_after();
  t1 = clock64();
  if (threadIdx.x == 0) out[4] = (t1 - t0);
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { mbar_wait(b, 0); __syncwarp(); }
  t1 = clock64();
  if (threadIdx.x == 0) { out[5] = (t1 - t0); out[6] = acc; }
}
}  // namespace loner

extern "C" int loner_probe_sync(int iters, long long* out, void* stream) {
  loner::probe_sync_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(iters, out);
  return cudaGetLastError() == cudaSuccess ? LONER_OK : LONER_E_LAUNCH;
}

// ---- the same question for the CTA-pair instruction: tcgen05.mma.cta_group::2, M = 256 over two SMs (each CTA holds its 128
// rows of A and its N/2 = 128 columns of B, like the pair variants of the pipelined kernels), issued back to back by the
// leader; optionally with a multicast commit per chunk.  cycles[cluster] = the leader's MMA stream duration.
namespace loner {
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1) probe_mma2_kernel(int iters, int commit_per_chunk,
                                                                                       long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t sA = smem_u32(smem), sB = sA + 16384;       // A [128 x 64] 16 KB | B half [64 K x 128 N] 16 KB
  for (int i = tid; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 1) tmem_alloc2<512>(smem_u32(&tmem_slot));
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    fence_mbar_init();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t done = smem_u32(&bars[0]), scratch = smem_u32(&bars[1]);
  const long long t0 = clock64();
  if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(256, 256, 0, 1);
      for (int it = 0; it < iters; ++it)
        umma2_f16_x4_warp<2, 128>(tmem, desc_lo_sw128(sA, 16), desc_hi_sw128(1024), desc_lo_sw128(sB, 8192), desc_hi_sw128(1024),
                                  idesc, it > 0 ? 1u : 0u, it == iters - 1 ? done : (commit_per_chunk ? scratch : 0u));
    }
    mbar_wait_warp(done, 0);                                  // both CTAs: the last commit is multicast
    if (rank == 0 && lane == 0) cycles[blockIdx.x >> 1] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2<512>(tmem);
}
}  // namespace loner

extern "C" int loner_probe_mma2(int iters, int commit_per_chunk, int clusters, long long* cycles, void* stream) {
  if (iters <= 0 || clusters <= 0 || !cycles) return LONER_E_BAD_ARG;
  cudaFuncSetAttribute(loner::probe_mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  loner::probe_mma2_kernel<<<2 * clusters, 64, 32768, (cudaStream_t)stream>>>(iters, commit_per_chunk, cycles);
  return cudaGetLastError() == cudaSuccess ? LONER_OK : LONER_E_LAUNCH;
}

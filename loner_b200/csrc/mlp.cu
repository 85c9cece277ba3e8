// Sigma head on the 5th-gen tensor cores: Frequency encoding + bias-free ReLU MLP, forward and
// hand-rolled backward, in tiles of 128 samples.
// Replaces tcnn.NetworkWithInputEncoding as used by DecoupledNeRF
// (/root/reference/src/models/nerf_tcnn.py:35-38, :59-78) and the xyz construction of
// render_rays (/root/reference/src/models/rendering_tcnn.py:241).
//
// Data layout (everything the tensor core touches is the SAME byte image in HBM and in smem):
//   "tile image" of a [rows x 64] fp16 block: row r is 128 B; 16-byte chunk j of row r sits at
//   chunk slot (j ^ (r & 7))  (the SWIZZLE_128B pattern); wider matrices are column blocks of 64
//   laid one after another.  Read K-major it is a [rows x 64k] operand, read MN-major it is the
//   transposed operand — so one image of the weights serves forward (B, MN-major) and dgrad
//   (B, K-major), and the activation / gradient images written by forward / dgrad are consumed
//   unchanged by wgrad (A and B, both MN-major).  All global<->shared traffic is plain bulk
//   async copies (cp.async.bulk) of contiguous ranges; no tensor maps.
//   Accumulators live in TMEM (tcgen05.mma, one issuing thread), epilogues read them back with
//   tcgen05.ld.
#include "common.cuh"
#include "sm100.cuh"

namespace loner {
namespace mlp {

using namespace sm100;

// ---- timeline instrumentation of the pipelined kernels.  Compiled ONLY into tests/probes/libloner_trace.so
// (-DLONER_TRACE, loner_b200.build.build_trace): one thread per role of CTAs 0 and 1 stores (tag, clock64) pairs into
// its own slice of a buffer set with loner_trace_setup; the product library contains none of this.
#ifdef LONER_TRACE
__device__ unsigned long long* g_trace_buf = nullptr;     // [2 CTAs][4 roles][cap][2]
__device__ unsigned int g_trace_cap = 0;
struct Trace {
  unsigned long long* p = nullptr;
  unsigned int n = 0, cap = 0;
  __device__ __forceinline__ void open(int role) {
    if (blockIdx.x < 2 && g_trace_buf != nullptr) {
      cap = g_trace_cap;
      p = g_trace_buf + ((size_t)(blockIdx.x * 4 + role) * cap) * 2;
    }
  }
  __device__ __forceinline__ void ev(unsigned event, unsigned unit, unsigned layer, unsigned tile, unsigned extra = 0) {
    if (p != nullptr && n < cap) {
      p[2 * n] = ((unsigned long long)event << 48) | ((unsigned long long)unit << 32) | (layer << 16) | (tile << 8) | extra;
      p[2 * n + 1] = (unsigned long long)clock64();
      ++n;
    }
  }
};
#define LONER_TRACE_OPEN(tr, role) Trace tr; tr.open(role)
#define LONER_TRACE_EV(tr, ...) tr.ev(__VA_ARGS__)
#else
#define LONER_TRACE_OPEN(tr, role)
#define LONER_TRACE_EV(tr, ...)
#endif

constexpr int kTile = 128;            // samples per tile = UMMA M
constexpr int kBlk = 16384;           // bytes of one [128 x 64] fp16 column block

// W = width the kernels run at (128 or 256); Wr = the network's real width.  A 64-wide network
// (tcnn FullyFusedMLP widths, BASELINE config 1) runs EXACTLY on the 128-wide kernels with its
// matrices zero-padded: padded neurons output relu(0) = 0, their mask bits are 0, and every padded
// weight gradient is a product with one of those zeros; only pack / reduce know the real layout.
struct Net {
  int F, E, Epad, W, L, nb, Wr, flags;
};

__host__ inline bool net_from(const loner_net_t* n, Net& o) {
  if (!n) return false;
  o.F = n->n_frequencies; o.Wr = n->n_neurons; o.L = n->n_hidden_layers; o.flags = n->flags;
  if (o.F < 1 || o.F > 10) return false;
  if (!(o.Wr == 64 || o.Wr == 128 || o.Wr == 256)) return false;
  if (o.L < 1 || o.L > 8) return false;
  o.W = o.Wr == 64 ? 128 : o.Wr;
  o.E = 6 * o.F; o.Epad = (o.E + 15) / 16 * 16; o.nb = o.W / 64;
  return true;
}
__host__ __device__ inline int layer_K(const Net& n, int l) { return l == 0 ? n.Epad : n.W; }
__host__ __device__ inline int layer_Kr(const Net& n, int l) { return l == 0 ? n.Epad : n.Wr; }   // real in-features
__host__ __device__ inline int64_t packed_off(const Net& n, int l) {      // byte offset of layer l's image
  return l == 0 ? 0 : (int64_t)n.Epad * n.W * 2 + (int64_t)(l - 1) * n.W * n.W * 2;
}
__host__ __device__ inline int64_t packed_wout_off(const Net& n) { return packed_off(n, n.L); }
__host__ __device__ inline int64_t param_off(const Net& n, int l) {       // float offset in flat params (real layout)
  return l == 0 ? 0 : (int64_t)n.Epad * n.Wr + (int64_t)(l - 1) * n.Wr * n.Wr;
}
__host__ __device__ inline int64_t act_tile_bytes(const Net& n) { return (int64_t)kBlk * (1 + n.L * n.nb); }
__host__ __device__ inline int64_t mask_tile_bytes(const Net& n) { return (int64_t)n.L * kTile * (n.W / 32) * 4; }
__host__ __device__ inline int64_t dz_tile_bytes(const Net& n) { return (int64_t)kBlk * n.L * n.nb; }

// ------------------------------------------------------------------------------------------
// pack: fp32 master [out,in] row-major -> two fp16 images, rows = in (k), cols = out (n):
//   "bwd" image  [column block][K rows][128 B]          (dgrad streams it one column block = one
//                                                         64-wide slice of its contraction at a time)
//   "fwd" image  [K chunk of 64 rows][column block][64 rows][128 B]   (forward streams it one
//                                                         64-row slice of ITS contraction at a time)
// so that every pipeline stage of either kernel is ONE contiguous bulk copy.
__host__ __device__ inline int64_t packed_fwd_base(const Net& n) { return packed_wout_off(n) + (int64_t)n.W * 4; }
// forward image: layer 0 always occupies one full 64-row chunk (rows >= Epad are never read)
__host__ __device__ inline int64_t fwd_off(const Net& n, int l) {
  return l == 0 ? 0 : (int64_t)64 * n.W * 2 + (int64_t)(l - 1) * n.W * n.W * 2;
}
__host__ __device__ inline int64_t packed_total(const Net& n) { return packed_fwd_base(n) + fwd_off(n, n.L); }

__global__ void pack_kernel(Net net, const float* __restrict__ params, uint8_t* __restrict__ packed) {
  const int l = blockIdx.y;
  if (l == net.L) {   // output layer: row 0 of [16, W], kept as fp32 values of the fp16-rounded weights
    float* wo = (float*)(packed + packed_wout_off(net));
    const float* src = params + param_off(net, net.L);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < net.W; j += gridDim.x * blockDim.x)
      wo[j] = j < net.Wr ? __half2float(__float2half_rn(src[j])) : 0.f;
    return;
  }
  const int K = layer_K(net, l), N = net.W;
  const int Kr = layer_Kr(net, l), Nr = net.Wr;
  const float* Wm = params + param_off(net, l);
  auto wv = [&](int n, int k) { return (n < Nr && k < Kr) ? Wm[(int64_t)n * Kr + k] : 0.f; };
  uint8_t* img_b = packed + packed_off(net, l);
  uint8_t* img_f = packed + packed_fwd_base(net) + fwd_off(net, l);
  const int chunks = K * (N / 8);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += gridDim.x * blockDim.x) {
    const int k = c % K, n0 = (c / K) * 8;
    __half2 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      h[i] = __floats2half2_rn(wv(n0 + 2 * i, k), wv(n0 + 2 * i + 1, k));
    const int cb = n0 / 64, j = (n0 % 64) / 8;
    const int sw = (j ^ (k & 7)) * 16;
    *reinterpret_cast<uint4*>(img_b + (int64_t)cb * K * 128 + (int64_t)k * 128 + sw) = *reinterpret_cast<uint4*>(h);
    *reinterpret_cast<uint4*>(img_f + (int64_t)(k >> 6) * (net.nb * 8192) + cb * 8192 + (k & 63) * 128 + sw) =
        *reinterpret_cast<uint4*>(h);
  }
}

// ------------------------------------------------------------------------------------------
// Shared pieces of the two pipelined kernels (forward, dgrad).
//
// One CTA per SM, 320 threads (352 with a second issuing warp, warp 10):  warp 0 = weight producer (bulk copies into a
// 3-slot ring), warp 1 = MMA issuer, warps 2-9 = epilogue (two warps per TMEM lane quarter, each owning half of the
// columns).  Two tiles of 128 samples are in flight with one 256-column TMEM accumulator each; inside
// a layer the MMAs of tile X (all K chunks) are followed by those of tile Y, and all eight epilogue
// warps drain X, then Y, so X's epilogue has the whole of Y's tensor time to finish (and vice versa).
// Design points, each measured on the GPU (profiles/README.md):
//  * producer and issuer run as CONVERGED warps, one lane elected inside each asm statement: issued
//    from a single lane of a diverged warp every UTCHMMA is wrapped in a lane-election loop and the
//    issuer, not the tensor pipe, bounded the kernel;
//  * ten warps leave 168 registers per thread: three 32-column TMEM buffers, 4-8 KB per warp in flight
//    (tests/gpu_probe.py: 8 warps reach 99 / 140 B/clk with one / two 4 KB loads in flight);
//  * no local memory at all (tables are address computations or live in shared memory): behind a
//    saturated HBM a local load misses the 28 KB L1 and stalls its warp for microseconds;
//  * with ONE issuing warp the weight chunks are fetched once per TILE (sharing them between the tiles of a pair from one
//    issuer - issue order X[c0,c1] Y[c0,c1] X[c2,c3] Y[c2,c3] - halves the L2->SM traffic, but then a tile's epilogue
//    overlaps only a quarter of the other tile's tensor work: measured slower in training, equal in inference, round 1);
//    the training forward and dgrad of CTA pairs run with one issuing warp PER TILE (kIss = 2, below), which shares the
//    chunks without that coupling;
//  * as CTA PAIRS (cta_group::2, kCtas = 2: training forward and dgrad) each SM stages half of every weight
//    chunk and one M = 256 MMA covers a tile of each CTA: -64 KB of operand reads and -64 KB of ring writes per
//    tile and layer in each SM's shared-memory pipe, half the weight re-streaming through the L2 fabric;
//  * what bounds the kernels (profiles/README.md): all run at 3200 - 3400 clk per tile and
//    layer against 2048 nominal tensor clocks - the inference forward on the per-SM L2->SM ingest of its weight
//    re-streaming (9.7 TB/s over 148 SMs), the training forward / dgrad on the L2 fabric ceiling for stash writes
//    next to weight loads (5.1 TB/s, tests/gpu_probe_store.py).  16 epilogue warps were tried and changed nothing
//    (profiles/experiments/r2_epilogue16.md).
constexpr int kPipeThreads = 320;
constexpr int kGroupThreads = 256;
// kIss = 2: one MMA-issuing warp PER TILE (warp 1 issues tile X, warp 10 tile Y) instead of one warp issuing X's and then
// Y's MMAs.  The issue of an MMA blocks until the tensor pipe accepts it, so a single issuer serialises 2 x 2048 clk of
// issue with all of its own waits, fences and descriptor set-up (tile-step 3250 clk instead of 2048,
// profiles/experiments/r2_timeline.md); with two issuers the pipe interleaves the two tiles' MMAs and each issuer's
// boundary work overlaps the other tile's MMAs.  Both tiles also consume the SAME weight chunk before its slot is
// released (w_empty counts two commits): every layer's weights are streamed once per unit, not once per tile.
template <int kIss> __host__ __device__ constexpr int pipe_threads() { return kPipeThreads + (kIss == 2 ? 32 : 0); }
constexpr int kRingBytes = 98304;                    // weight ring: 3 x 32 KB (one CTA) or 6 x 16 KB (CTA pair)
constexpr int kPipeExtra = 2576;
constexpr int kPipeSmem = 2 * 65536 + kRingBytes + kPipeExtra;

// Shared-memory map of the two pipelined kernels.  Everything is an ADDRESS COMPUTATION (no arrays of
// pointers): a table indexed by a run-time slot number would live in local memory, and local loads
// miss the (tiny, 28 KB) L1 behind a saturated HBM.
// kCtas = 2: the CTA pair variant (cta_group::2) - each CTA stages only ITS half (N/2 columns) of every weight
// chunk, so the ring has twice the slots at half the size; the leader's `w_full` barriers count two arrivals: its own
// producer's expect_tx and the peer's relay warp, which arrives (remotely) when the peer's half of the slot has landed.
template <int kCtas>
struct PipeSmem {
  static constexpr uint32_t kSlots = kCtas == 2 ? 6 : 3;
  static constexpr uint32_t kSlotBytes = kRingBytes / kSlots;
  static constexpr uint32_t kX = 131072u + kRingBytes;      // wout | part | barriers | tmem slot | encoding table
  uint8_t* base;
  uint32_t base_u32;
  __device__ __forceinline__ uint8_t* tileA_ptr(int t) const { return base + t * 65536; }
  __device__ __forceinline__ uint32_t tileA(int t) const { return base_u32 + (uint32_t)t * 65536u; }
  __device__ __forceinline__ uint32_t ring(uint32_t slot) const { return base_u32 + 131072u + slot * kSlotBytes; }
  __device__ __forceinline__ float* wout() const { return reinterpret_cast<float*>(base + kX); }
  __device__ __forceinline__ float* part() const { return wout() + 256; }   // [2][128] per-row partial sums (upper -> lower column half)
  __device__ __forceinline__ uint32_t bars() const { return base_u32 + kX + 2048u; }
  __device__ __forceinline__ uint32_t w_full(uint32_t i) const { return bars() + 8u * i; }
  __device__ __forceinline__ uint32_t w_empty(uint32_t i) const { return bars() + 8u * (kSlots + i); }
  __device__ __forceinline__ uint32_t a_ready(int t) const { return bars() + 8u * (2 * kSlots + t); }
  __device__ __forceinline__ uint32_t acc_full(int t) const { return bars() + 8u * (2 * kSlots + 2 + t); }
  __device__ __forceinline__ uint32_t* tmem_slot() const { return reinterpret_cast<uint32_t*>(base + kX + 2304); }
  // encoding table: feature pair p -> {kind, 2^f}; kind 0..2 = input dimension, 3 = padding ones, 4 = zeros
  __device__ __forceinline__ uint2* enc_tab() const { return reinterpret_cast<uint2*>(base + kX + 2320); }
};

template <int kCtas>
__device__ __forceinline__ PipeSmem<kCtas> carve(uint8_t* base) {
  PipeSmem<kCtas> p;
  p.base = base;
  p.base_u32 = smem_u32(base);
  return p;
}

template <int kCtas, int kIss = 1>
__device__ __forceinline__ void pipe_init(const PipeSmem<kCtas>& sm, int tid, int warp, const float* wout_src, int W,
                                          const Net& net) {
  constexpr uint32_t kSlots = PipeSmem<kCtas>::kSlots;
  if (warp == 1) {
    if (kCtas == 2) tmem_alloc2<512>(smem_u32(sm.tmem_slot()));
    else tmem_alloc<512>(smem_u32(sm.tmem_slot()));
  }
  if (tid == 0) {
    for (uint32_t i = 0; i < kSlots; ++i) {
      // pair, leader: a slot is full when its own half has landed (expect_tx arrive + bytes) AND the peer's relay warp
      // has arrived for the peer's half - ONE barrier for the issuer to wait on per chunk
      mbar_init(sm.w_full(i), (kCtas == 2 && cluster_ctarank() == 0) ? 2 : 1); mbar_init(sm.w_empty(i), kIss);
    }
    // pair: one arrive per epilogue WARP of either CTA (8 + 8); single CTA: one per epilogue thread
    for (int t = 0; t < 2; ++t) { mbar_init(sm.a_ready(t), kCtas == 2 ? 16 : kGroupThreads); mbar_init(sm.acc_full(t), 1); }
    fence_mbar_init();
  }
  for (int j = tid; j < W; j += pipe_threads<kIss>()) sm.wout()[j] = wout_src[j];
  if (tid >= 64 && tid < 96) {          // features [0, 6F): sin/cos pairs ordered [dim][freq]; [6F, Epad): ones; rest zeros
    const int p = tid - 64, dim = p / net.F, f = p % net.F;
    uint2 e;
    if (dim < 3) e = make_uint2((uint32_t)dim, __float_as_uint((float)(1 << f)));
    else e = make_uint2(2 * p < net.Epad ? 3u : 4u, 0u);
    sm.enc_tab()[p] = e;
  }
  tc_fence_before();
  __syncthreads();
  if (kCtas == 2) cluster_sync_all();    // the peer's barriers exist before any remote arrive / multicast commit
  tc_fence_after();
}

// An epilogue thread has finished (and fenced) its part of tile t's next A operand.
template <int kCtas>
__device__ __forceinline__ void arrive_a(uint32_t a_ready_addr, int lane) {
  if (kCtas == 2) {                      // a_ready_addr = the LEADER's barrier (shared::cluster address)
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(a_ready_addr);
  } else {
    mbar_arrive(a_ready_addr);
  }
}

template <int kCtas>
__device__ __forceinline__ void pipe_teardown(uint32_t tmem, int warp) {
  tc_fence_before();
  __syncthreads();
  if (kCtas == 2) {
    cluster_sync_all();                  // the peer may still read this CTA's operands / arrive on its barriers
    if (warp == 1) tmem_dealloc2<512>(tmem);
  } else {
    if (warp == 1) tmem_dealloc<512>(tmem);
  }
}

__device__ __forceinline__ void epi_bar() {   // the 256 epilogue threads
  asm volatile("bar.sync 1, 256;" ::: "memory");
}

// two fp32 -> packed half2 (lo = a, hi = b) with ReLU folded into the conversion
__device__ __forceinline__ uint32_t cvt_relu_h2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ uint32_t cvt_sat_h2(float a, float b) {   // saturates to the finite fp16 range
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// ReLU mask word of 32 columns given as 16 packed half2 (post-ReLU, so "active" = non-zero fp16 output,
// which is what tcnn's backward tests too).  Bit p = column 2p, bit 16+p = column 2p+1: one HSET2 and
// one LOP3 per PAIR of columns.  mask_bit(bits, c) is the inverse mapping.
__device__ __forceinline__ uint32_t relu_mask_word(const uint32_t* hh) {
  uint32_t m = 0u;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    uint32_t gt;
    asm("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(gt) : "r"(hh[p]), "r"(0u));
    m |= gt & ((1u << p) | (1u << (16 + p)));
  }
  return m;
}
__device__ __forceinline__ bool mask_bit(uint32_t bits, int c) {   // c = column within the 32-column word
  return ((bits >> ((c >> 1) + 16 * (c & 1))) & 1u) != 0u;
}
// 0xFFFF in each half of the result whose column (pair kPair of the mask word) is active: the two bits
// are moved to the sign positions of bytes 1 and 3 and replicated by PRMT (selector nibbles 8|byte).
template <int kPair>
__device__ __forceinline__ uint32_t half2_mask(uint32_t bits) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(m) : "r"(bits << (15 - kPair)));
  return m;
}

__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 32 columns (16 packed half2) of this thread's row into the swizzled tile image in shared memory.
// srow = the row's base inside the tile, xs = (row & 7) << 4, col0 = first column (multiple of 32).
__device__ __forceinline__ void store32(uint32_t srow, uint32_t xs, int col0, const uint32_t* hh) {
  const uint32_t cb_off = (uint32_t)(col0 >> 6) * kBlk;
  const uint32_t j0 = ((uint32_t)(col0 & 63) >> 3) << 4;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    sts128(srow + cb_off + ((j0 + 16u * k) ^ xs), hh[4 * k], hh[4 * k + 1], hh[4 * k + 2], hh[4 * k + 3]);
}

// Drains kCols accumulator columns of this thread's TMEM lane in 32-column loads.  128 columns: two
// loads are issued before the first wait and the third / fourth go out while the first / second are
// being processed, so 4-8 KB per warp stay in flight (tcgen05.wait::ld waits for ALL outstanding
// loads, hence the grouping).  pin32() keeps the compiler from moving uses above the issue points.
template <int kCols, class Proc>
__device__ __forceinline__ void drain32(uint32_t acc, Proc&& proc) {
  static_assert(kCols == 64 || kCols == 128, "column halves of W = 128 / 256");
  uint32_t b0[32], b1[32];
  tmem_ld32(acc, b0);
  tmem_ld32(acc + 32, b1);
  tmem_ld_wait();
  pin32(b0);
  pin32(b1);
  if constexpr (kCols == 128) {
    uint32_t b2[32], b3[32];
    tmem_ld32(acc + 64, b2);
    pin32(b0);
    proc(0, b0);
    tmem_ld32(acc + 96, b3);
    pin32(b1);
    proc(1, b1);
    tmem_ld_wait();
    pin32(b2);
    pin32(b3);
    proc(2, b2);
    proc(3, b3);
  } else {
    proc(0, b0);
    proc(1, b1);
  }
}

// Per-tile inputs of a row, fetched one tile pair ahead of their use.
struct RowIn {
  float v[7];   // pos mode: x,y,z ; ray mode: z, o[3], d[3]
};
__device__ __forceinline__ float ldg_now(const float* p) {   // volatile: issued where written, not sunk to its use
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ RowIn load_row(const float* pos, const float* rays, const float* z, int S, int s_shift,
                                          int64_t gs) {
  RowIn r;
  if (pos) {
    r.v[0] = ldg_now(pos + gs * 3 + 0); r.v[1] = ldg_now(pos + gs * 3 + 1); r.v[2] = ldg_now(pos + gs * 3 + 2);
    r.v[3] = r.v[4] = r.v[5] = r.v[6] = 0.f;
  } else {
    const int64_t ray = s_shift >= 0 ? (gs >> s_shift) : (int64_t)((uint64_t)gs / (uint32_t)S);
    const float* R = rays + ray * LONER_RAY_COLS;
    r.v[0] = ldg_now(z + gs);
#pragma unroll
    for (int a = 0; a < 6; ++a) r.v[1 + a] = ldg_now(R + a);
  }
  return r;
}
// sample position in [0,1]^3:  (o + d z + 1) / 2     rendering_tcnn.py:241, nerf_tcnn.py:63
__device__ __forceinline__ void row_pos01(bool pos_mode, const RowIn& r, float (&x)[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float p = pos_mode ? r.v[a] : __fadd_rn(r.v[1 + a], __fmul_rn(r.v[4 + a], r.v[0]));
    x[a] = __fmul_rn(__fadd_rn(p, 1.0f), 0.5f);
  }
}

// ------------------------------------------------------------------------------------------
struct FwdArgs {
  Net net;
  const uint8_t* packed;
  const float* pos;      // [P,3] or null
  const float* rays;     // [n,13]
  const float* z;        // [n,S]
  int S;
  int s_shift;           // log2(S) when S is a power of two, else -1
  int64_t P;
  int64_t tiles;
  float* sigma;          // [P]
  uint8_t* acts;         // activation stash or null
  uint8_t* masks;        // relu bit masks or null
  int stash_last;        // 1: A_L is stashed too; 0: nothing reads it (wgrad derives dW_out from dW_{L-1}'s partials)
};

// sin/cos(pi * 2^f * x).  2^f * x is exact in fp32, and so is its reduction r to [-1, 1]; sin(pi r) and
// cos(pi r) then come from the SFU (__sinf/__cosf on |pi r| <= pi: abs error < 1e-6, far below the
// fp16 rounding of the encoded feature).  `scale` = 2^f as a float.
__device__ __forceinline__ void freq_pair(float x, float scale, float& s, float& c) {
  const float t = x * scale;
  const float r = fmaf(-2.0f, rintf(0.5f * t), t);
  const float a = 3.14159265358979323846f * r;
  s = __sinf(a);
  c = __cosf(a);
}

// Writes encoded features [32*half, 32*half+32) of row r (column block 0 of `sA`) from the table built
// by pipe_init (tcnn pads the encoded width to 16 with ones).  A table in shared memory, not per-thread
// state: the compiler turned the running (dim, f, 2^f) of an unrolled loop into local-memory constants.
__device__ __forceinline__ void encode_row(uint32_t sA, int r, int half, const float (&x)[3], const uint2* tab) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int cj = half * 4 + c;
    __half2 h[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint2 e = tab[cj * 4 + q];            // features 2*pair, 2*pair+1
      float s, co;
      if (e.x < 3u) {
        const float xv = e.x == 0u ? x[0] : (e.x == 1u ? x[1] : x[2]);
        freq_pair(xv, __uint_as_float(e.y), s, co);
      } else {
        s = co = (e.x == 3u) ? 1.0f : 0.0f;
      }
      h[q] = __floats2half2_rn(s, co);
    }
    const uint32_t* hw = reinterpret_cast<const uint32_t*>(h);
    sts128(sA + r * 128 + ((cj ^ (r & 7)) * 16), hw[0], hw[1], hw[2], hw[3]);
  }
}

// Work units of the pipelined kernels.  Single CTA: a unit is a PAIR of tiles (X, Y), CTA b takes pairs b, b + grid, ...
// CTA pair: a unit is a QUAD of tiles; cluster c takes quads c, c + clusters, ...; the CTA of rank r owns the pair
// 2 q + r of the quad, and the leader's M = 256 MMAs cover X = (X_0, X_1), then Y = (Y_0, Y_1).
template <int kCtas>
struct Units {
  int64_t first, stride, count;     // in units
  uint32_t rank;
  __device__ __forceinline__ Units(int64_t tiles) {
    if (kCtas == 2) {
      rank = cluster_ctarank();
      first = blockIdx.x >> 1; stride = gridDim.x >> 1; count = (tiles + 3) / 4;
    } else {
      rank = 0; first = blockIdx.x; stride = gridDim.x; count = (tiles + 1) / 2;
    }
  }
  __device__ __forceinline__ int64_t pair(int64_t unit) const { return kCtas == 2 ? 2 * unit + rank : unit; }
};

template <int W, bool kStash, int kCtas, int kIss>
__global__ void __launch_bounds__(pipe_threads<kIss>(), 1) mlp_fwd_kernel(const FwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  using Smem = PipeSmem<kCtas>;
  const Smem sm = carve<kCtas>(smem_raw);
  const Net net = a.net;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pipe_init<kCtas, kIss>(sm, tid, warp, reinterpret_cast<const float*>(a.packed + packed_wout_off(net)), W, net);
  const uint32_t tmem = *sm.tmem_slot();
  constexpr uint32_t nslots = Smem::kSlots;
  const Units<kCtas> units(a.tiles);
  constexpr int kNb = W / 64;
  constexpr int kFetches = (kIss == 2) ? 1 : 2;     // weight streams per unit and layer: one per tile, or one shared by both
  constexpr uint32_t kChunkBytes = kNb * 8192;      // 64 K-rows x W out-features, fp16
  constexpr uint32_t kMyBytes = kChunkBytes / kCtas;   // pair: this CTA stages column blocks [rank*kNb/2, (rank+1)*kNb/2)

  if (warp == 0) {
    // ---------------- producer (whole warp, converged): one contiguous bulk copy per 64-row weight chunk
    const uint8_t* fimg = a.packed + packed_fwd_base(net) + units.rank * kMyBytes;
    uint32_t g = 0;
    LONER_TRACE_OPEN(tr, 3);
    for (int64_t u = units.first; u < units.count; u += units.stride) {
      for (int l = 0; l < net.L; ++l) {
        const int nch = (l == 0) ? 1 : kNb;
        for (int t = 0; t < kFetches; ++t) {     // one issuer: every tile fetches its own chunks, X[all K] then Y[all K]
          for (int c = 0; c < nch; ++c, ++g) {
            const uint32_t slot = g % nslots, use = g / nslots;
            if (use > 0) mbar_wait_warp(sm.w_empty(slot), (use - 1) & 1);
            if (lane == 0) LONER_TRACE_EV(tr, 0, (unsigned)((u - units.first) / units.stride), l, t, c);      // chunk load issued
            mbar_expect_tx_warp(sm.w_full(slot), kMyBytes);
            bulk_g2s_warp(sm.ring(slot), fimg + fwd_off(net, l) + (int64_t)c * kChunkBytes, kMyBytes, sm.w_full(slot));
          }
        }
      }
    }
  } else if ((warp == 1 || (kIss == 2 && warp == 10)) && units.rank == 0) {
    // ---------------- MMA issuer(s) (whole warp, converged; one lane is elected inside each asm statement).
    // kIss = 1: this warp issues tile X, then tile Y; kIss = 2: warp 1 owns tile X, warp 10 tile Y, same chunk sequence.
    constexpr uint32_t idesc = make_idesc_f16(128 * kCtas, W, 0, 1);
    uint32_t g = 0, par_a = 0u;
    const int t_lo = (kIss == 2 && warp == 10) ? 1 : 0, t_hi = (kIss == 2 && warp == 1) ? 1 : 2;
    LONER_TRACE_OPEN(tr, 0);
    for (int64_t u = units.first; u < units.count; u += units.stride) {
      for (int l = 0; l < net.L; ++l) {
        const int nch = (l == 0) ? 1 : kNb;
        const int ksteps0 = (l == 0) ? net.Epad / 16 : 4;   // layer 0 contracts over Epad (<= 64) features
        for (int t = t_lo; t < t_hi; ++t) {
          if (lane == 0) LONER_TRACE_EV(tr, 0, (unsigned)((u - units.first) / units.stride), l, t);           // waits for A
          mbar_wait_warp(sm.a_ready(t), par_a);
          if (lane == 0) LONER_TRACE_EV(tr, 5, (unsigned)((u - units.first) / units.stride), l, t);           // a_ready observed (before the fence)
          tc_fence_after();
          if (lane == 0) LONER_TRACE_EV(tr, 1, (unsigned)((u - units.first) / units.stride), l, t);           // A is ready
          for (int c = 0; c < nch; ++c, ++g) {
            const uint32_t slot = g % nslots, par_w = (g / nslots) & 1;
            mbar_wait_warp(sm.w_full(slot), par_w);
            if (lane == 0) LONER_TRACE_EV(tr, 6, (unsigned)((u - units.first) / units.stride), l, t, c);      // chunk observed (before the fence)
            tc_fence_after();
            if (lane == 0) LONER_TRACE_EV(tr, 2, (unsigned)((u - units.first) / units.stride), l, t, c);      // chunk c has landed
            const uint32_t sa = sm.tileA(t) + c * kBlk, sb = sm.ring(slot);
            if (ksteps0 == 4) {
              if (kCtas == 2)
                umma2_f16_x4_warp<2, 128>(tmem + t * 256, desc_lo_sw128(sa, 16), desc_hi_sw128(1024), desc_lo_sw128(sb, 8192),
                                          desc_hi_sw128(1024), idesc, c > 0 ? 1u : 0u, sm.w_empty(slot));
              else
                umma_f16_x4_warp<2, 128>(tmem + t * 256, desc_lo_sw128(sa, 16), desc_hi_sw128(1024), desc_lo_sw128(sb, 8192),
                                         desc_hi_sw128(1024), idesc, c > 0 ? 1u : 0u, sm.w_empty(slot));
            } else {
              for (int ks = 0; ks < ksteps0; ++ks) {
                const uint64_t ad = make_desc_sw128(sa + ks * 32, 16, 1024);
                const uint64_t bd = make_desc_sw128(sb + ks * 2048, 8192, 1024);
                if (kCtas == 2) umma2_f16_warp(tmem + t * 256, ad, bd, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                else umma_f16_warp(tmem + t * 256, ad, bd, idesc, (c > 0 || ks > 0) ? 1u : 0u);
              }
              if (kCtas == 2) umma2_commit_warp(sm.w_empty(slot)); else umma_commit_warp(sm.w_empty(slot));
            }
            if (lane == 0) LONER_TRACE_EV(tr, 4, (unsigned)((u - units.first) / units.stride), l, t, c);      // chunk c's MMAs issued
          }
          if (kCtas == 2) umma2_commit_warp(sm.acc_full(t)); else umma_commit_warp(sm.acc_full(t));
          if (lane == 0) LONER_TRACE_EV(tr, 3, (unsigned)((u - units.first) / units.stride), l, t);           // all MMAs of the tile-step issued
        }
        par_a ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ---------------- relay (peer CTA of a pair): tells the leader when THIS CTA's half of a slot has landed
    uint32_t g = 0;
    for (int64_t u = units.first; u < units.count; u += units.stride) {
      for (int l = 0; l < net.L; ++l) {
        const int n = kFetches * ((l == 0) ? 1 : kNb);
        for (int i = 0; i < n; ++i, ++g) {
          const uint32_t slot = g % nslots;
          mbar_wait_warp(sm.w_full(slot), (g / nslots) & 1);
          if (lane == 0) mbar_arrive_cluster(mapa_u32(sm.w_full(slot), 0));
          __syncwarp();
        }
      }
    }
  } else if (warp < 10) {
    // ---------------- epilogue warps: thread = (sample row, column half); tile X, then tile Y.
    // Stash mode: every finished image (A_0 .. A_L) leaves its tile buffer as ONE bulk copy issued by
    // the elected thread after the step's barrier.  The elected thread waits for the reads of all
    // earlier copies BEFORE that barrier, and the steps alternate X, Y, X, ... - so whenever a step
    // starts writing a tile buffer, the copy issued from it two steps ago has been read out.
    const int e = warp - 2;
    const int h = e >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool elected = (e == 0) && lane == 0;
    const uint32_t xs = (uint32_t)(row & 7) << 4;
    constexpr int kCols = W / 2;                      // columns per thread
    const uint32_t acc_base = tmem + ((uint32_t)(q * 32) << 16) + h * kCols;
    uint32_t par_acc = 0;
    const bool pos_mode = a.pos != nullptr;
    const bool keep_last = kStash && a.stash_last != 0;      // the last layer's image is written (and copied out) at all
    const uint32_t a_rdy[2] = {kCtas == 2 ? mapa_u32(sm.a_ready(0), 0) : sm.a_ready(0),
                               kCtas == 2 ? mapa_u32(sm.a_ready(1), 0) : sm.a_ready(1)};
    auto row_index = [&](int64_t pair, int t) {
      int64_t gs = (2 * pair + t) * kTile + row;
      return gs < a.P ? gs : a.P - 1;
    };
    RowIn nxt[2];
    LONER_TRACE_OPEN(tr, e == 0 ? 1 : 2);     // first and (below) last epilogue warp
#pragma unroll
    for (int t = 0; t < 2; ++t) nxt[t] = load_row(a.pos, a.rays, a.z, a.S, a.s_shift, row_index(units.pair(units.first), t));
    for (int64_t u = units.first; u < units.count; u += units.stride) {
      const int64_t pair = units.pair(u);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int64_t tile = 2 * pair + t;
        const bool active = tile < a.tiles;
        const uint32_t sA = sm.tileA(t);
        {
          float x[3];
          row_pos01(pos_mode, nxt[t], x);
          encode_row(sA, row, h, x, sm.enc_tab());
        }
        fence_async_smem();
        if (kStash && elected) bulk_wait_read0();
        arrive_a<kCtas>(a_rdy[t], lane);
        if (u + units.stride < units.count)
          nxt[t] = load_row(a.pos, a.rays, a.z, a.S, a.s_shift, row_index(units.pair(u + units.stride), t));
        if (kStash) {
          epi_bar();
          if (elected && active) { bulk_s2g(a.acts + tile * act_tile_bytes(net), sA, (uint32_t)kBlk); bulk_commit(); }
        }
      }
      for (int l = 0; l < net.L; ++l) {
        const bool last = (l == net.L - 1);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int64_t tile = 2 * pair + t;
          const bool active = tile < a.tiles;
          const int64_t gs = tile * kTile + row;
          const bool in = active && gs < a.P;
          const uint32_t sA = sm.tileA(t);
          const uint32_t srow = sA + row * 128;
          float* part = sm.part() + t * 128;
          if (lane == 0 && (e == 0 || e == 7)) LONER_TRACE_EV(tr, 0, (unsigned)((u - units.first) / units.stride), l, t);   // waits for the accumulator
          mbar_wait(sm.acc_full(t), par_acc);
          tc_fence_after();
          if (lane == 0 && (e == 0 || e == 7)) LONER_TRACE_EV(tr, 1, (unsigned)((u - units.first) / units.stride), l, t);   // accumulator complete
          float sig0 = 0.f, sig1 = 0.f;
          uint32_t mbits[kCols / 32];
          drain32<kCols>(acc_base + t * 256, [&](int i, uint32_t (&v)[32]) {
            const int col0 = h * kCols + i * 32;
            // packed in place: v[p] <- (v[2p], v[2p+1]); the 16 results stay in the (aligned, consecutive)
            // registers the TMEM load wrote, so the 128-bit stores need no register shuffling
#pragma unroll
            for (int p = 0; p < 16; ++p) v[p] = cvt_relu_h2(__uint_as_float(v[2 * p]), __uint_as_float(v[2 * p + 1]));
            if (kStash) mbits[i] = relu_mask_word(v);
            if (!last || keep_last) store32(srow, xs, col0, v);
            if (last) {
#pragma unroll
              for (int p = 0; p < 16; ++p) {
                const float2 r2 = __half22float2(*reinterpret_cast<const __half2*>(&v[p]));
                const float2 w2 = *reinterpret_cast<const float2*>(sm.wout() + col0 + 2 * p);
                sig0 = fmaf(r2.x, w2.x, sig0);
                sig1 = fmaf(r2.y, w2.y, sig1);
              }
            }
          });
          const float sig = sig0 + sig1;
          tc_fence_before();
          if (lane == 0 && (e == 0 || e == 7)) LONER_TRACE_EV(tr, 2, (unsigned)((u - units.first) / units.stride), l, t);   // drained + stored
          if (kStash && active) {
            uint32_t* mrow = reinterpret_cast<uint32_t*>(a.masks + tile * mask_tile_bytes(net)) +
                             ((int64_t)l * kTile + row) * (W / 32) + h * (kCols / 32);
            if (kCols / 32 == 4) {
              *reinterpret_cast<uint4*>(mrow) = make_uint4(mbits[0], mbits[1], mbits[2], mbits[3]);
            } else {
#pragma unroll
              for (int it = 0; it < kCols / 32; ++it) mrow[it] = mbits[it];
            }
          }
          if (!last || kStash) fence_async_smem();     // the image becomes visible to the MMA / bulk-copy engines
          if (kStash && elected) bulk_wait_read0();
          if (!last) arrive_a<kCtas>(a_rdy[t], lane);
          else if (h == 1) part[row] = sig;
          if (lane == 0 && (e == 0 || e == 7)) LONER_TRACE_EV(tr, 3, (unsigned)((u - units.first) / units.stride), l, t);   // handed off
          if (kStash || last) epi_bar();
          if (last && h == 0 && in) a.sigma[gs] = sig + part[row];
          if (kStash && elected && active && (!last || keep_last)) {
            bulk_s2g(a.acts + tile * act_tile_bytes(net) + kBlk + (int64_t)l * kNb * kBlk, sA, (uint32_t)(kNb * kBlk));
            bulk_commit();
          }
        }
        par_acc ^= 1u;
      }
    }
    if (kStash && elected) bulk_wait0();
  }
  pipe_teardown<kCtas>(tmem, warp);
}

// ------------------------------------------------------------------------------------------
// dgrad: d_sigma -> dZ_L ... dZ_1 (fp16, loss-scaled, stashed for wgrad) and optionally d_pos.
struct BwdArgs {
  Net net;
  const uint8_t* packed;
  const float* pos;
  const float* rays;
  const float* z;
  int S;
  int s_shift;
  int64_t P;
  int64_t tiles;
  const float* d_sigma;
  const uint8_t* masks;
  uint8_t* dz;           // dZ stash [tiles][L][nb*16 KB]
  int stash_last;        // 1: the dZ_L image is stashed too; 0: wgrad rebuilds it from masks, d_sigma, w_out
  float gscale;
  float* d_pos;          // [P,3] or null
};

template <int W, bool kDx, int kCtas, int kIss>
__global__ void __launch_bounds__(pipe_threads<kIss>(), 1) mlp_dgrad_kernel(const BwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  using Smem = PipeSmem<kCtas>;
  const Smem sm = carve<kCtas>(smem_raw);
  const Net net = a.net;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pipe_init<kCtas, kIss>(sm, tid, warp, reinterpret_cast<const float*>(a.packed + packed_wout_off(net)), W, net);
  const uint32_t tmem = *sm.tmem_slot();
  constexpr uint32_t nslots = Smem::kSlots;
  const Units<kCtas> units(a.tiles);
  constexpr int kFetches = (kIss == 2) ? 1 : 2;   // weight streams per unit and layer (see mlp_fwd_kernel)
  constexpr bool want_dx = kDx;            // d_pos requested: one more GEMM (layer 0) and the encoding backward
  const int l_lo = want_dx ? 0 : 1;        // GEMMs run for l = L-1 .. l_lo : dA_l = dZ_{l+1} * W_l
  constexpr int kNb = W / 64;              // contraction (out-features of layer l) in 64-wide chunks

  if (warp == 0) {
    // ---------------- producer (whole warp, converged).  Pair: this CTA stages rows [rank*K_l/2, (rank+1)*K_l/2)
    // (its half of the GEMM's N = in-features) of every column block.
    uint32_t g = 0;
    for (int64_t u = units.first; u < units.count; u += units.stride) {
      for (int l = net.L - 1; l >= l_lo; --l) {
        const uint32_t bytes = (uint32_t)layer_K(net, l) * 128u;     // one column block: K_l rows x 128 B
        const uint32_t mine = bytes / kCtas;
        for (int t = 0; t < kFetches; ++t) {
          for (int c = 0; c < kNb; ++c, ++g) {
            const uint32_t slot = g % nslots, use = g / nslots;
            if (use > 0) mbar_wait_warp(sm.w_empty(slot), (use - 1) & 1);
            mbar_expect_tx_warp(sm.w_full(slot), mine);
            bulk_g2s_warp(sm.ring(slot), a.packed + packed_off(net, l) + (int64_t)c * bytes + units.rank * mine, mine,
                          sm.w_full(slot));
          }
        }
      }
    }
  } else if ((warp == 1 || (kIss == 2 && warp == 10)) && units.rank == 0) {
    // ---------------- MMA issuer(s) (whole warp, converged); kIss = 2: warp 1 owns tile X, warp 10 tile Y
    uint32_t g = 0, par_a = 0u;
    const int t_lo = (kIss == 2 && warp == 10) ? 1 : 0, t_hi = (kIss == 2 && warp == 1) ? 1 : 2;
    for (int64_t u = units.first; u < units.count; u += units.stride) {
      for (int l = net.L - 1; l >= l_lo; --l) {
        const uint32_t idesc = make_idesc_f16(128 * kCtas, layer_K(net, l), 0, 0);
        for (int t = t_lo; t < t_hi; ++t) {
          mbar_wait_warp(sm.a_ready(t), par_a);
          tc_fence_after();
          for (int c = 0; c < kNb; ++c, ++g) {
            const uint32_t slot = g % nslots, par_w = (g / nslots) & 1;
            mbar_wait_warp(sm.w_full(slot), par_w);
            tc_fence_after();
            const uint32_t sa = sm.tileA(t) + c * kBlk, sb = sm.ring(slot);
            if (kCtas == 2)
              umma2_f16_x4_warp<2, 2>(tmem + t * 256, desc_lo_sw128(sa, 16), desc_hi_sw128(1024), desc_lo_sw128(sb, 16),
                                      desc_hi_sw128(1024), idesc, c > 0 ? 1u : 0u, sm.w_empty(slot));
            else
              umma_f16_x4_warp<2, 2>(tmem + t * 256, desc_lo_sw128(sa, 16), desc_hi_sw128(1024), desc_lo_sw128(sb, 16),
                                     desc_hi_sw128(1024), idesc, c > 0 ? 1u : 0u, sm.w_empty(slot));
          }
          if (kCtas == 2) umma2_commit_warp(sm.acc_full(t)); else umma_commit_warp(sm.acc_full(t));
        }
        par_a ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ---------------- relay (peer CTA of a pair)
    uint32_t g = 0;
    for (int64_t u = units.first; u < units.count; u += units.stride) {
      for (int l = net.L - 1; l >= l_lo; --l) {
        for (int i = 0; i < kFetches * kNb; ++i, ++g) {
          const uint32_t slot = g % nslots;
          mbar_wait_warp(sm.w_full(slot), (g / nslots) & 1);
          if (lane == 0) mbar_arrive_cluster(mapa_u32(sm.w_full(slot), 0));
          __syncwarp();
        }
      }
    }
  } else if (warp < 10) {
    // ---------------- epilogue warps: thread = (sample row, column half); tile X, then tile Y
    // (same step / barrier / bulk-copy protocol as the forward kernel).
    const int e = warp - 2;
    const int h = e >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool elected = (e == 0) && lane == 0;
    const uint32_t xs = (uint32_t)(row & 7) << 4;
    constexpr int kCols = W / 2;
    constexpr int kWords = W / 32;
    constexpr int kMw = kCols / 32;
    const uint32_t acc_row = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t par_acc = 0;
    const bool pos_mode = a.pos != nullptr;
    // mask words of (tile, layer index m) for this thread's row and column half
    auto load_masks = [&](int64_t tile, int m, uint32_t (&mw)[kMw]) {
      const float* mrow = reinterpret_cast<const float*>(a.masks + (tile < a.tiles ? tile : 0) * mask_tile_bytes(net)) +
                          ((int64_t)m * kTile + row) * kWords + h * kMw;
#pragma unroll
      for (int it = 0; it < kMw; ++it) mw[it] = __float_as_uint(ldg_now(mrow + it));
    };
    auto load_ds = [&](int64_t tile) {
      const int64_t gs = tile * kTile + row;
      return (tile < a.tiles && gs < a.P) ? ldg_now(a.d_sigma + gs) : 0.f;     // unscaled: nothing consumes the load here
    };
    // Masks (and d_sigma) of the NEXT step of each tile are fetched one step ahead, and AFTER the step's hand-off: the
    // loop-carried copy of a prefetched register waits for its load (ncu: 32 % of this kernel's stall samples sat on
    // those MOVs and on the d_sigma multiply, in front of the a_ready arrive - a memory latency on the
    // epilogue -> MMA -> epilogue chain of every step).
    uint32_t mw[2][kMw];
    float ds[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      load_masks(2 * units.pair(units.first) + t, net.L - 1, mw[t]);
      ds[t] = load_ds(2 * units.pair(units.first) + t);
    }
    const uint32_t a_rdy[2] = {kCtas == 2 ? mapa_u32(sm.a_ready(0), 0) : sm.a_ready(0),
                               kCtas == 2 ? mapa_u32(sm.a_ready(1), 0) : sm.a_ready(1)};
    const int64_t next_tile = 2 * (int64_t)gridDim.x;     // this thread's tile t of the NEXT unit (both variants)
    for (int64_t u = units.first; u < units.count; u += units.stride) {
      const int64_t pair = units.pair(u);
      RowIn rin[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int64_t tile = 2 * pair + t;
        const bool active = tile < a.tiles;
        const int64_t gs = tile * kTile + row;
        const bool in = active && gs < a.P;
        const uint32_t stile = sm.tileA(t);
        const uint32_t srow = stile + row * 128;
        if (want_dx && h == 0) rin[t] = load_row(a.pos, a.rays, a.z, a.S, a.s_shift, in ? gs : a.P - 1);
        // dZ_L = d_sigma * w_out * relu'(Z_L)
        const float dsv = ds[t] * a.gscale;
#pragma unroll
        for (int it = 0; it < kMw; ++it) {
          const uint32_t bits = mw[t][it];
          const int col0 = h * kCols + it * 32;
          uint32_t hh[16];
#define LONER_DZL(P)                                                                        \
  {                                                                                         \
    const float2 w2 = *reinterpret_cast<const float2*>(sm.wout() + col0 + 2 * P);           \
    hh[P] = cvt_sat_h2(dsv * w2.x, dsv * w2.y) & half2_mask<P>(bits);                       \
  }
          LONER_DZL(0) LONER_DZL(1) LONER_DZL(2) LONER_DZL(3) LONER_DZL(4) LONER_DZL(5) LONER_DZL(6) LONER_DZL(7)
          LONER_DZL(8) LONER_DZL(9) LONER_DZL(10) LONER_DZL(11) LONER_DZL(12) LONER_DZL(13) LONER_DZL(14) LONER_DZL(15)
#undef LONER_DZL
          store32(srow, xs, col0, hh);
        }
        fence_async_smem();
        if (elected) bulk_wait_read0();
        arrive_a<kCtas>(a_rdy[t], lane);
        epi_bar();
        if (a.stash_last && elected && active) {
          bulk_s2g(a.dz + tile * dz_tile_bytes(net) + (int64_t)(net.L - 1) * kNb * kBlk, stile, (uint32_t)(kNb * kBlk));
          bulk_commit();
        }
        if (net.L >= 2) {
          load_masks(tile, net.L - 2, mw[t]);      // for step (t, L-1)
        } else if (!want_dx) {                     // single hidden layer, no GEMM steps: next pair's first step
          load_masks(tile + next_tile, net.L - 1, mw[t]);
          ds[t] = load_ds(tile + next_tile);
        }
      }
      for (int l = net.L - 1; l >= l_lo; --l) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int64_t tile = 2 * pair + t;
          const bool active = tile < a.tiles;
          const int64_t gs = tile * kTile + row;
          const bool in = active && gs < a.P;
          const uint32_t stile = sm.tileA(t);
          const uint32_t srow = stile + row * 128;
          mbar_wait(sm.acc_full(t), par_acc);
          tc_fence_after();
          if (l >= 1) {
            // dZ_l = dA_l * relu'(Z_l)  -> fp16 image (next GEMM's A operand in smem, wgrad's B operand in HBM)
            const bool feeds_gemm = (l - 1 >= l_lo);
            drain32<kCols>(acc_row + t * 256 + h * kCols, [&](int i, uint32_t (&v)[32]) {
              const uint32_t bits = mw[t][i];
#define LONER_DZ(P) v[P] = cvt_sat_h2(__uint_as_float(v[2 * P]), __uint_as_float(v[2 * P + 1])) & half2_mask<P>(bits);
              LONER_DZ(0) LONER_DZ(1) LONER_DZ(2) LONER_DZ(3) LONER_DZ(4) LONER_DZ(5) LONER_DZ(6) LONER_DZ(7)
              LONER_DZ(8) LONER_DZ(9) LONER_DZ(10) LONER_DZ(11) LONER_DZ(12) LONER_DZ(13) LONER_DZ(14) LONER_DZ(15)
#undef LONER_DZ
              store32(srow, xs, h * kCols + i * 32, v);
            });
            tc_fence_before();
            fence_async_smem();
            if (elected) bulk_wait_read0();
            if (feeds_gemm) arrive_a<kCtas>(a_rdy[t], lane);
            epi_bar();
            if (elected && active) {
              bulk_s2g(a.dz + tile * dz_tile_bytes(net) + (int64_t)(l - 1) * kNb * kBlk, stile, (uint32_t)(kNb * kBlk));
              bulk_commit();
            }
            // masks (and d_sigma) of this tile's next step: layer l-2 of this pair, or the first step of the next pair
            if (l >= 2) {
              load_masks(tile, l - 2, mw[t]);
            } else if (!want_dx) {
              load_masks(tile + next_tile, net.L - 1, mw[t]);
              ds[t] = load_ds(tile + next_tile);
            }
          } else {
            // l == 0: dEnc [128 x Epad] -> d_pos through the sin/cos encoding (lower-half warps only)
            if (h == 0) {
              float x[3];
              row_pos01(pos_mode, rin[t], x);
              float dx[3] = {0.f, 0.f, 0.f};
              const uint2* tab = sm.enc_tab();
              for (int it = 0; it * 32 < net.Epad; ++it) {
                uint32_t v[32];
                tmem_ld32(acc_row + t * 256 + it * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int pq = 0; pq < 16; ++pq) {
                  const uint2 e = tab[it * 16 + pq];
                  if (e.x < 3u) {
                    const float xv = e.x == 0u ? x[0] : (e.x == 1u ? x[1] : x[2]);
                    const float scale = __uint_as_float(e.y);
                    float sn, cs;
                    freq_pair(xv, scale, sn, cs);
                    // d/dx sin(pi 2^f x) = pi 2^f cos, d/dx cos = -pi 2^f sin
                    const float g = (__uint_as_float(v[2 * pq]) * cs - __uint_as_float(v[2 * pq + 1]) * sn) *
                                    (3.14159265358979323846f * scale);
                    if (e.x == 0u) dx[0] += g; else if (e.x == 1u) dx[1] += g; else dx[2] += g;
                  }
                }
              }
              if (in) {
                const float inv = 0.5f / a.gscale;      // x = (pos + 1) / 2
                a.d_pos[gs * 3 + 0] = dx[0] * inv;
                a.d_pos[gs * 3 + 1] = dx[1] * inv;
                a.d_pos[gs * 3 + 2] = dx[2] * inv;
              }
            }
            tc_fence_before();
            load_masks(tile + next_tile, net.L - 1, mw[t]);
            ds[t] = load_ds(tile + next_tile);
          }
        }
        par_acc ^= 1u;
      }
    }
    if (elected) bulk_wait0();
  }
  pipe_teardown<kCtas>(tmem, warp);
}

// ------------------------------------------------------------------------------------------
// wgrad: dW_l[n,k] = sum_s dZ_{l+1}[s,n] * A_l[s,k], one CTA per (layer, range of tiles),
// accumulators resident in TMEM over the whole range, operands streamed in half tiles (64 samples).
//  * CTAs of the LAST hidden layer do not read dZ_L: it is relu'(Z_L) * d_sigma * w_out - a bit mask
//    times a rank-1 matrix - and eight otherwise idle warps rebuild its half-tile image in shared
//    memory from the mask words, d_sigma and w_out while the producer streams A_{L-1}
//    (-64 KB/tile written by dgrad, -64 KB/tile read here).
//  * The OUTPUT layer's gradient dW_out[n] = sum_s d_sigma[s] * A_L[s,n] needs no A_L at all ("fold", the
//    default when dZ_L is rebuilt): the network has no biases, so A_L[s,n] = mask_L[s,n] * sum_k A_{L-1}[s,k] W_{L-1}[n,k],
//    and with G[n,k] = sum_s d_sigma[s] mask_L[s,n] A_{L-1}[s,k] (what the CTAs of layer L-1 accumulate when the
//    rebuilt operand leaves out w_out)   dW_{L-1}[n,k] = w_out[n] G[n,k]   and   dW_out[n] = sum_k W_{L-1}[n,k] G[n,k],
//    both applied in fp32 by wgrad_reduce_kernel.  The forward then does not stash A_L (-64 KB/tile written,
//    -64 KB/tile read: 22 % of the forward's and 14 % of this kernel's HBM traffic).  The rebuilt operand is
//    fp16(d_sigma * gscale * c), c = the power of two >= max |w_out| (same dynamic range as d_sigma * gscale * w_out).
//  * Without the fold (LONER_NET_STASH_AL, a single hidden layer, or a stashed dZ_L) CTAs of layer 0 (the lightest
//    stream: A_0 is 16 KB per tile) also stream A_L and accumulate dW_out on the CUDA cores of the helper warps,
//    reduced deterministically like every other partial (round 1: a separate 0.5 ms kernel that ended in atomics).
struct WgradArgs {
  Net net;
  const uint8_t* acts;
  const uint8_t* dz;
  int64_t tiles;
  float* partials;           // [items][K_l*N_l] per layer, then [items of layer 0][W] for dW_out
  const uint8_t* masks;
  const float* d_sigma;
  const float* wout;         // [W] fp32 values of the fp16-rounded output weights (packed image)
  float gscale;
  int64_t P;
  int gen_last;              // 1: CTAs of layer L-1 rebuild dZ_L; 0: they read dgrad's stash of it
  int fold_out;              // 1 (needs gen_last): the rebuilt operand omits w_out, dW_out comes from layer L-1's partials
  const uint8_t* packed;     // fp16 weight images (the reduce kernel reads W_{L-1} for the fold)
  int item_begin[9];         // first item of each layer (prefix), item_begin[L] = total
  int64_t part_off[10];      // float offset of each layer's first partial; [L] = dW_out partials; [L+1] = total
};

// The power of two >= m (clamped to 2^-14 .. 2^14; 1 for m = 0): an exact scale factor, computed identically by the
// wgrad kernel (operand) and the reduce kernel (un-scaling).
__device__ __forceinline__ float pow2_ceil(float m) {
  if (!(m > 0.f)) return 1.f;
  const uint32_t b = __float_as_uint(m);
  int e = (int)(b >> 23) + ((b & 0x7FFFFFu) ? 1 : 0);
  e = e < 113 ? 113 : (e > 141 ? 141 : e);
  return __uint_as_float((uint32_t)e << 23);
}

constexpr int kWgStages = 3;
constexpr int kWgThreads = 320;     // warp 0 producer, warp 1 MMA, warps 2-9 helpers (generate dZ_L / accumulate dW_out), 2-5 epilogue
// stage of a layer > 0 CTA:  X = A_l half (32 KB) | Y = dZ_{l+1} half (32 KB)
// stage of a layer-0 CTA:    X = A_0 half (8 KB)  | Y = dZ_1 half (32 KB) | Z = A_L half (32 KB)
// The helpers' small per-row inputs travel by bulk copy too (a register prefetch of global loads turns into a
// wait of a full HBM latency per stage - the loop-carried copy of the prefetched value waits for its load):
//  * layer-0 CTAs: d_sigma of the 64 rows (256 B) rides in the stage's tail [73728, 73984);
//  * dZ_L-rebuilding CTAs (layer > 0, stage tail [65536, 74752) unused): a ring of kWgIn input slots
//    (ReLU mask words of the half tile, 8 W bytes, + d_sigma, 256 B) with its own barriers, fetched kWgIn - 1
//    stages ahead, so generating Y needs only an EMPTY stage and overlaps the latency of the stage's A load.
constexpr int kWgStageBytes = 74752;
constexpr int kWgIn = 6;
constexpr int kWgInBytes = 2304;
constexpr int kWgSmem = 1024 + kWgStages * kWgStageBytes + 2048;

__global__ void __launch_bounds__(kWgThreads, 1) mlp_wgrad_kernel(const WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  float* s_wout = reinterpret_cast<float*>(base + kWgStages * kWgStageBytes);          // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_wout + 256);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kWgStages + 1 + 2 * kWgIn);
  const Net net = a.net;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int l = 0;
  while (l + 1 < net.L && (int)blockIdx.x >= a.item_begin[l + 1]) ++l;
  const int item = blockIdx.x - a.item_begin[l];
  const int n_items = a.item_begin[l + 1] - a.item_begin[l];
  const int64_t t0 = a.tiles * item / n_items, t1 = a.tiles * (item + 1) / n_items;
  const int K = layer_K(net, l);                 // in-features of layer l
  const int64_t part_sz = (int64_t)K * net.W;
  float* part = a.partials + a.part_off[l] + (int64_t)item * part_sz;

  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[kWgStages + s]); };
  const uint32_t done_bar = smem_u32(&bars[2 * kWgStages]);
  auto in_full = [&](int j) { return smem_u32(&bars[2 * kWgStages + 1 + j]); };
  auto in_empty = [&](int j) { return smem_u32(&bars[2 * kWgStages + 1 + kWgIn + j]); };
  auto in_slot = [&](int j) { return base + (j % kWgStages) * kWgStageBytes + 65536 + (j / kWgStages) * kWgInBytes; };   // masks | d_sigma

  const bool gen_y = a.gen_last && (l == net.L - 1);   // this CTA rebuilds dZ_L instead of loading it
  const bool fold = gen_y && a.fold_out != 0;          // ... without w_out: this CTA accumulates G (see above)
  const bool do_out = (l == 0) && a.fold_out == 0;     // this CTA also streams A_L and accumulates dW_out
  // stage geometry ("X" = A_l half image, "Y" = dZ_{l+1} half image, "Z" = A_L half image)
  const int nbA = (l == 0) ? 1 : net.nb;          // column blocks of A_l
  const int nbY = net.nb;
  const uint32_t offY = (l == 0) ? 8192u : 32768u, offZ = 40960u;
  const uint32_t bytesA = (uint32_t)nbA * 8192, bytesY = (uint32_t)nbY * 8192;
  if (warp == 0) tmem_alloc<512>(smem_u32(s_tmem));
  if (tid == 32) {
    // full: the producer's expect_tx arrive (+ the helpers' arrive when they write Y); empty: the MMA commit
    // (+ the helpers' arrive when they read Z)
    for (int s = 0; s < kWgStages; ++s) { mbar_init(full_bar(s), gen_y ? 2 : 1); mbar_init(empty_bar(s), do_out ? 2 : 1); }
    for (int j = 0; j < kWgIn; ++j) { mbar_init(in_full(j), 1); mbar_init(in_empty(j), 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  for (int j = tid; j < net.W; j += kWgThreads) s_wout[j] = a.wout[j];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  const int64_t actA_off = (l == 0) ? 0 : (int64_t)kBlk + (int64_t)(l - 1) * net.nb * kBlk;
  const int64_t actL_off = (int64_t)kBlk + (int64_t)(net.L - 1) * net.nb * kBlk;
  const int64_t dzY_off = (int64_t)l * net.nb * kBlk;
  const int64_t n_half = (t1 - t0) * 2;
  const uint32_t mask_bytes = 8u * (uint32_t)net.W;      // ReLU mask words of one half tile (64 rows x W/32 words)
  // rows of half tile i all inside [0, P): its 64 d_sigma values can be fetched as one 256-byte bulk copy
  auto rows_full = [&](int64_t i) { return (t0 + (i >> 1)) * kTile + (i & 1) * 64 + 64 <= a.P; };

  if (warp == 0) {
    // ---- producer (whole warp, converged): bulk loads, one column block (64 rows x 128 B = 8 KB) per copy
    auto issue_in = [&](int64_t j) {     // mask words + d_sigma of half tile j -> input slot j % kWgIn
      const int sl = (int)(j % kWgIn);
      mbar_wait_warp(in_empty(sl), (uint32_t)(((j / kWgIn) & 1) ^ 1));
      const int64_t tile = t0 + (j >> 1);
      const int hf = (int)(j & 1);
      const bool full = rows_full(j);
      const uint32_t dst = smem_u32(in_slot(sl));
      mbar_expect_tx_warp(in_full(sl), mask_bytes + (full ? 256u : 0u));
      bulk_g2s_warp(dst, a.masks + tile * mask_tile_bytes(net) + ((int64_t)(net.L - 1) * kTile + hf * 64) * (net.W / 32) * 4,
                    mask_bytes, in_full(sl));
      if (full) bulk_g2s_warp(dst + 2048u, a.d_sigma + tile * kTile + hf * 64, 256u, in_full(sl));
    };
    if (gen_y)
      for (int64_t j = 0; j < kWgIn - 1 && j < n_half; ++j) issue_in(j);
    for (int64_t i = 0; i < n_half; ++i) {
      if (gen_y && i + kWgIn - 1 < n_half) issue_in(i + kWgIn - 1);
      const int s = (int)(i % kWgStages);
      const uint32_t ph = (uint32_t)((i / kWgStages) & 1);
      mbar_wait_warp(empty_bar(s), ph ^ 1u);
      const int64_t tile = t0 + (i >> 1);
      const int hf = (int)(i & 1);
      const uint32_t dstA = smem_u32(base + s * kWgStageBytes), dstY = dstA + offY, dstZ = dstA + offZ;
      const bool ds_bulk = do_out && rows_full(i);
      mbar_expect_tx_warp(full_bar(s), bytesA + (gen_y ? 0u : bytesY) + (do_out ? bytesY : 0u) + (ds_bulk ? 256u : 0u));
      if (ds_bulk) bulk_g2s_warp(dstA + 73728u, a.d_sigma + tile * kTile + hf * 64, 256u, full_bar(s));
      const uint8_t* tile_acts = a.acts + tile * act_tile_bytes(net);
      const uint8_t* srcA = tile_acts + actA_off + hf * 8192;
      const uint8_t* srcY = a.dz + tile * dz_tile_bytes(net) + dzY_off + hf * 8192;
      const uint8_t* srcZ = tile_acts + actL_off + hf * 8192;
      for (int cb = 0; cb < nbA; ++cb) bulk_g2s_warp(dstA + cb * 8192, srcA + (int64_t)cb * kBlk, 8192, full_bar(s));
      if (!gen_y)
        for (int cb = 0; cb < nbY; ++cb) bulk_g2s_warp(dstY + cb * 8192, srcY + (int64_t)cb * kBlk, 8192, full_bar(s));
      if (do_out)
        for (int cb = 0; cb < nbY; ++cb) bulk_g2s_warp(dstZ + cb * 8192, srcZ + (int64_t)cb * kBlk, 8192, full_bar(s));
    }
  } else if (warp == 1) {
    // ---- MMA issuer (whole warp, converged).  Both operands MN-major: 64-element MN blocks 8 KB apart
    // (LBO), 8-sample groups 1 KB apart (SBO), 16 samples per instruction = 2 KB per k-step.
    const bool swapped = (l == 0);   // layer 0: M = out-features (from dZ), N = Epad (from A_0)
    const int n_mblk = swapped ? net.W / 128 : K / 128;
    const int Ncols = swapped ? net.Epad : net.W;
    const uint32_t idesc = make_idesc_f16(128, Ncols, 1, 1);
    for (int64_t i = 0; i < n_half; ++i) {
      const int s = (int)(i % kWgStages);
      const uint32_t ph = (uint32_t)((i / kWgStages) & 1);
      mbar_wait_warp(full_bar(s), ph);
      tc_fence_after();
      const uint32_t sX = smem_u32(base + s * kWgStageBytes), sY = sX + offY;
      for (int mb = 0; mb < n_mblk; ++mb) {
        const uint32_t aaddr = (swapped ? sY : sX) + mb * 2 * 8192;
        const uint32_t baddr = (swapped ? sX : sY);
        umma_f16_x4_warp<128, 128>(tmem + mb * 256, desc_lo_sw128(aaddr, 8192), desc_hi_sw128(1024),
                                   desc_lo_sw128(baddr, 8192), desc_hi_sw128(1024), idesc, i > 0 ? 1u : 0u,
                                   mb == n_mblk - 1 ? empty_bar(s) : 0u);
      }
    }
    umma_commit_warp(done_bar);
  }
  __syncwarp();
  float out_acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) out_acc[k] = 0.f;
  if (warp >= 2 && (gen_y || do_out)) {
    // ---- helpers (8 warps; four of them are the epilogue warps, idle during the main loop)
    const int g = tid - 64;                       // 0..255
    // (a) generator mapping: warp = (32-row half of the 64-sample half tile, quarter of the columns), lane = row.
    // Lanes on distinct ROWS make the swizzled 128-bit stores conflict-free (like dgrad's epilogue) and the w_out
    // reads warp-uniform broadcasts; with lanes spread over column quarters instead every store took 16 and every
    // w_out load 8 shared-memory wavefronts (ncu), ~3000 clk per stage - the whole cost of round 1's variant.
    const int r = ((g >> 5) & 1) * 32 + lane, qc = g >> 6;
    const int cpt = net.W / 4;                    // columns per thread: 64 (W=256) or 32 (W=128)
    const int mwords = net.W / 32;
    // (b) dW_out mapping: thread = (16-byte chunk j of a row = 8 columns, rows rg, rg+8, ...): the lanes of a warp
    // read the 32 chunks of ONE row (conflict-free 128-bit loads), the swizzle term (r & 7) = rg is per-thread constant
    const int oj = g & 31, rg = g >> 5;
    const bool out_on = do_out && (oj < net.nb * 8);
    const uint32_t out_off = (uint32_t)(oj >> 3) * 8192u + ((uint32_t)((oj & 7) ^ rg) << 4);
    float cfold = 1.f;
    if (fold) {
      float m = 0.f;
      for (int j = 0; j < net.W; ++j) m = fmaxf(m, fabsf(s_wout[j]));
      cfold = pow2_ceil(m);
    }
    for (int64_t i = 0; i < n_half; ++i) {
      const int s = (int)(i % kWgStages);
      const uint32_t ph = (uint32_t)((i / kWgStages) & 1);
      const uint32_t stage = smem_u32(base + s * kWgStageBytes);
      const int64_t gs_half = (t0 + (i >> 1)) * kTile + (i & 1) * 64;       // first row of this half tile
      const bool full = rows_full(i);
      if (gen_y) {
        const int sl = (int)(i % kWgIn);
        mbar_wait(in_full(sl), (uint32_t)((i / kWgIn) & 1));
        const uint8_t* in = in_slot(sl);
        const uint32_t* mrow = reinterpret_cast<const uint32_t*>(in) + r * mwords + qc * (cpt / 32);
        const uint32_t mw0 = mrow[0], mw1 = cpt > 32 ? mrow[1] : 0u;
        float ds = full ? reinterpret_cast<const float*>(in + 2048)[r] : (gs_half + r < a.P ? __ldg(a.d_sigma + gs_half + r) : 0.f);
        ds *= a.gscale * cfold;
        const uint32_t hfold = cvt_sat_h2(ds, ds);
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t srow = stage + offY + (uint32_t)r * 128u;
        const uint32_t xs = (uint32_t)(r & 7) << 4;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          if (it * 32 < cpt) {
            const uint32_t bits = it == 0 ? mw0 : mw1;
            const int col0 = qc * cpt + it * 32;
            uint32_t hh[16];
            if (fold) {
#define LONER_GEN(P) hh[P] = hfold & half2_mask<P>(bits);
              LONER_GEN(0) LONER_GEN(1) LONER_GEN(2) LONER_GEN(3) LONER_GEN(4) LONER_GEN(5) LONER_GEN(6) LONER_GEN(7)
              LONER_GEN(8) LONER_GEN(9) LONER_GEN(10) LONER_GEN(11) LONER_GEN(12) LONER_GEN(13) LONER_GEN(14) LONER_GEN(15)
#undef LONER_GEN
            } else {
#define LONER_GEN(P)                                                                      \
  {                                                                                       \
    const float2 w2 = *reinterpret_cast<const float2*>(s_wout + col0 + 2 * P);            \
    hh[P] = cvt_sat_h2(ds * w2.x, ds * w2.y) & half2_mask<P>(bits);                       \
  }
              LONER_GEN(0) LONER_GEN(1) LONER_GEN(2) LONER_GEN(3) LONER_GEN(4) LONER_GEN(5) LONER_GEN(6) LONER_GEN(7)
              LONER_GEN(8) LONER_GEN(9) LONER_GEN(10) LONER_GEN(11) LONER_GEN(12) LONER_GEN(13) LONER_GEN(14) LONER_GEN(15)
#undef LONER_GEN
            }
            // same image as dgrad's store32, with 8 KB (64-row) column blocks
            const uint32_t cb_off = (uint32_t)(col0 >> 6) * 8192u;
            const uint32_t j0 = ((uint32_t)(col0 & 63) >> 3) << 4;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              sts128(srow + cb_off + ((j0 + 16u * k) ^ xs), hh[4 * k], hh[4 * k + 1], hh[4 * k + 2], hh[4 * k + 3]);
          }
        }
        fence_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");      // all helper threads have read their inputs, written and fenced
        if (g == 0) { mbar_arrive(full_bar(s)); mbar_arrive(in_empty(sl)); }
      }
      if (do_out) {
        mbar_wait(full_bar(s), ph);                         // the A_L half (and its d_sigma) have landed
        if (out_on) {
          const float* sds = reinterpret_cast<const float*>(base + s * kWgStageBytes + 73728);
          const uint32_t zrow = stage + offZ + out_off + (uint32_t)rg * 128u;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int row = rg + 8 * k;
            const float dsk = full ? sds[row] : (gs_half + row < a.P ? __ldg(a.d_sigma + gs_half + row) : 0.f);
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(zrow + (uint32_t)k * 1024u));
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&w0));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&w1));
            const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&w2));
            const float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&w3));
            out_acc[0] = fmaf(dsk, f0.x, out_acc[0]); out_acc[1] = fmaf(dsk, f0.y, out_acc[1]);
            out_acc[2] = fmaf(dsk, f1.x, out_acc[2]); out_acc[3] = fmaf(dsk, f1.y, out_acc[3]);
            out_acc[4] = fmaf(dsk, f2.x, out_acc[4]); out_acc[5] = fmaf(dsk, f2.y, out_acc[5]);
            out_acc[6] = fmaf(dsk, f3.x, out_acc[6]); out_acc[7] = fmaf(dsk, f3.y, out_acc[7]);
          }
        }
        __syncwarp();
        asm volatile("bar.sync 1, 256;" ::: "memory");      // every helper has read its part of Z
        if (g == 0) mbar_arrive(empty_bar(s));              // second arrival: the MMA commit is the first
      }
    }
  }
  if (warp >= 2 && do_out) {
    // ---- dW_out partial of this CTA: sum the eight row groups through the (now idle) first stage
    mbar_wait(done_bar, 0);                                 // all MMAs (the last readers of the stages) are done
    const int g = tid - 64;
    const int oj = g & 31, rg = g >> 5;
    float* red = reinterpret_cast<float*>(base);            // [8][W]
    if (oj < net.nb * 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) red[rg * net.W + oj * 8 + k] = out_acc[k];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (g < net.W) {
      float sum = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) sum += red[q * net.W + g];
      a.partials[a.part_off[net.L] + (int64_t)item * net.W + g] = (n_half == 0) ? 0.f : sum;
    }
  }
  if (warp >= 2 && warp < 6) {
    // ---- epilogue: TMEM -> partial sums in [out n][in k] order (the flat params order)
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    const bool swapped = (l == 0);
    const int n_mblk = swapped ? net.W / 128 : K / 128;
    const int Ncols = swapped ? net.Epad : net.W;
    for (int mb = 0; mb < n_mblk; ++mb) {
      const int m = mb * 128 + q * 32 + lane;
      for (int c0 = 0; c0 < Ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * 256 + c0), v);
        tmem_ld_wait();
        if (n_half == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        if (!swapped) {
#pragma unroll
          for (int i = 0; i < 32; ++i) part[(int64_t)(c0 + i) * K + m] = __uint_as_float(v[i]);   // m = k
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < K) part[(int64_t)m * K + (c0 + i)] = __uint_as_float(v[i]);              // m = n
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// d_params += (sum over a layer's items of its partials) / gscale; blockIdx.y == L: the output layer's row
__global__ void wgrad_reduce_kernel(Net net, const float* __restrict__ partials, WgradArgs w, float inv_gscale,
                                    float* __restrict__ d_params) {
  const int l = blockIdx.y;
  __shared__ float s_c;
  if (w.fold_out) {     // the scale of the rebuilt operand (see pow2_ceil)
    if (threadIdx.x < 32) {
      float m = 0.f;
      for (int j = threadIdx.x; j < net.W; j += 32) m = fmaxf(m, fabsf(w.wout[j]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (threadIdx.x == 0) s_c = pow2_ceil(m);
    }
    __syncthreads();
  }
  if (l == net.L && w.fold_out) return;     // dW_out is finished by the blocks of layer L-1 (below)
  if (l == net.L) {     // dW_out: one partial [W] per item of layer 0; d_sigma entered unscaled
    const int n_items = w.item_begin[1] - w.item_begin[0];
    const float* p = partials + w.part_off[net.L];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < net.Wr; c += gridDim.x * blockDim.x) {
      float s = 0.f;
      for (int g = 0; g < n_items; ++g) s += p[(int64_t)g * net.W + c];
      d_params[param_off(net, net.L) + c] += s;
    }
    return;
  }
  const int K = layer_K(net, l), Kr = layer_Kr(net, l);
  const int64_t sz = (int64_t)K * net.W;          // a multiple of 2048: every thread of a block runs the same trips
  const int n_items = w.item_begin[l + 1] - w.item_begin[l];
  const float* p = partials + w.part_off[l];
  // fold (layer L-1, K = W): the summed partial is G[n,k];  dW_{L-1}[n,k] = w_out[n] G[n,k] / (gscale c)  and
  // dW_out[n] = sum_k W_{L-1}[n,k] G[n,k] / (gscale c) - a block covers 256 / K whole rows n per trip, reduced in a fixed
  // order (warp shuffles, then the row's K / 32 warp sums): one writer per n, no atomics, nothing read twice
  const bool fold_l = w.fold_out && l == net.L - 1;
  const uint8_t* img = w.packed + packed_off(net, l);      // "bwd" image: [column block][K rows][128 B], swizzled
  __shared__ float s_red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < sz; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    const int n = (int)(i / K), k = (int)(i % K);         // partials are [out n][in k] at the kernel width
    const bool real = n < net.Wr && k < Kr;               // not the zero-padded part of a 64-wide network
    float s = 0.f;
    if (real) {
#pragma unroll 8
      for (int g = 0; g < n_items; ++g) s += p[(int64_t)g * sz + i];
      const float sc = fold_l ? inv_gscale * (w.wout[n] / s_c) : inv_gscale;
      d_params[param_off(net, l) + (int64_t)n * Kr + k] += s * sc;
    }
    if (fold_l) {
      float v = 0.f;
      if (real) {
        const int cb = n >> 6, j = (n & 63) >> 3;
        const __half wv = *reinterpret_cast<const __half*>(img + (int64_t)cb * K * 128 + (int64_t)k * 128 + ((j ^ (k & 7)) * 16) + (n & 7) * 2);
        v = __half2float(wv) * s;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_red[warp] = v;
      __syncthreads();
      const int wpr = K >> 5;                               // warps per row n (K = 128 or 256; blockDim.x = 256)
      if ((int)threadIdx.x < 256 / K) {
        const int nr = (int)(i0 / K) + (int)threadIdx.x;
        float t = 0.f;
        for (int q = 0; q < wpr; ++q) t += s_red[threadIdx.x * wpr + q];
        if (nr < net.Wr) d_params[param_off(net, net.L) + nr] += t * (inv_gscale / s_c);
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------
struct WgradPlan {
  int item_begin[9];
  int64_t part_off[10];
  int total_items;
  int64_t total_floats;
};

inline int device_sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      sms <= 0) {
    cudaGetLastError();
    sms = 148;
  }
  return sms;
}

inline WgradPlan plan_wgrad(const Net& net, bool gen_last, bool fold_out) {
  // The kernel is HBM-bound (it streams the A_l and dZ_{l+1} images once): give every layer a share
  // of the SMs proportional to the BYTES it reads per tile, not to its flops.  In 8 KB column blocks per
  // half tile: layer 0 reads A_0 (1) + dZ_1 (nb) + A_L (nb, for dW_out); a middle layer A_l + dZ_{l+1} (2 nb);
  // the last layer only A_{L-1} (nb) when it rebuilds dZ_L.
  const int sms = device_sm_count();
  WgradPlan p;
  double wgt[8], total = 0;
  for (int l = 0; l < net.L; ++l) {
    const bool last = (l == net.L - 1);
    // (a dZ_L-rebuilding CTA keeps only 3 x 32 KB of loads in flight instead of 3 x 64 KB, so under a saturated HBM it
    // streams at about half the rate of the others: weighted as if it still read dZ_L.  Sweep on a B200, C2, wgrad ms:
    // weight 0.25 nb 3.04, 0.75 nb 2.34, 1.0 nb 2.26, 1.25 nb 2.26, 1.5 nb 2.26)
    double w = (l == 0 ? 1.0 : (double)net.nb) + (double)net.nb + ((l == 0 && !fold_out) ? (double)net.nb : 0.0);
    // Same rule for layer 0 once it no longer streams A_L (fold): its stages hold 8 + 32 KB instead of 64 KB, so a CTA keeps
    // 120 KB in flight where a middle layer's keeps 192 KB, and needs a share in proportion to bytes per byte in flight:
    // (1 + nb) / 120 = 2 nb / 192 for nb = 4.  LONER_NET_WG_PLAN_BYTES keeps the share proportional to bytes (A/B).
    if (l == 0 && fold_out && !(net.flags & LONER_NET_WG_PLAN_BYTES)) w = 2.0 * (double)net.nb;
    (void)last; (void)gen_last;
    wgt[l] = w;
    total += w;
  }
  // shares rounded down, the SMs left over go to the layers with the largest remainders (every SM streams)
  int cnt[8], sum = 0;
  double rem[8];
  for (int l = 0; l < net.L; ++l) {
    const double share = (double)sms * wgt[l] / total;
    cnt[l] = (int)share < 1 ? 1 : (int)share;
    rem[l] = share - (double)cnt[l];
    sum += cnt[l];
  }
  for (; sum < sms; ++sum) {
    int best = 0;
    for (int l = 1; l < net.L; ++l) if (rem[l] > rem[best]) best = l;
    ++cnt[best];
    rem[best] -= 1.0;
  }
  int used = 0;
  int64_t off = 0;
  for (int l = 0; l < net.L; ++l) {
    const int g = cnt[l];
    p.item_begin[l] = used;
    p.part_off[l] = off;
    used += g;
    off += (int64_t)g * layer_K(net, l) * net.W;
  }
  p.item_begin[net.L] = used;
  p.part_off[net.L] = off;                                   // dW_out partials: one [W] row per item of layer 0
  off += (int64_t)(p.item_begin[1] - p.item_begin[0]) * net.W;
  p.part_off[net.L + 1] = off;
  p.total_items = used;
  p.total_floats = off;
  return p;
}

}  // namespace mlp
}  // namespace loner

using namespace loner::mlp;

#ifdef LONER_TRACE
extern "C" int loner_trace_setup(void* buf, uint32_t cap_per_role) {
  unsigned long long* p = (unsigned long long*)buf;
  if (cudaMemcpyToSymbol(loner::mlp::g_trace_buf, &p, sizeof(p)) != cudaSuccess) return LONER_E_LAUNCH;
  if (cudaMemcpyToSymbol(loner::mlp::g_trace_cap, &cap_per_role, sizeof(cap_per_role)) != cudaSuccess) return LONER_E_LAUNCH;
  return LONER_OK;
}
#endif

extern "C" int64_t loner_mlp_param_count(const loner_net_t* n) {
  Net net;
  if (!net_from(n, net)) return -1;
  return param_off(net, net.L) + 16 * (int64_t)net.Wr;
}
extern "C" int64_t loner_mlp_packed_bytes(const loner_net_t* n) {
  Net net;
  if (!net_from(n, net)) return -1;
  return packed_total(net);
}
static inline int64_t n_tiles(int64_t P) { return (P + kTile - 1) / kTile; }
// Kernel variants are selected by loner_net_t.flags (include/loner_b200.h), never by the environment.
static inline int pipe_ctas(const Net& net) { return (net.flags & LONER_NET_SINGLE_CTA) ? 1 : 2; }
// (a single hidden layer keeps the stash: its one wgrad CTA type already fills its stages with A_0, dZ_1 and A_1)
static inline bool wgrad_rebuilds_last(const Net& net) { return (net.flags & LONER_NET_STASH_DZL) == 0 && net.L >= 2; }
// dW_out from layer L-1's partials, no A_L stash (needs the rebuilt dZ_L operand)
static inline bool wgrad_folds_out(const Net& net) { return wgrad_rebuilds_last(net) && (net.flags & LONER_NET_STASH_AL) == 0; }

// Launch of a pipelined kernel: one CTA per SM, as CTA pairs (clusters of 2 on one TPC) or single CTAs.
template <class Kern, class Args>
static void launch_pipe(Kern kern, const Args& args, int64_t tiles, int ctas, cudaStream_t st, int threads = kPipeThreads) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPipeSmem);
  const int sms = device_sm_count();
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = kPipeSmem;
  cfg.stream = st;
  if (ctas == 2) {
    const int64_t quads = (tiles + 3) / 4;
    const int64_t clusters = quads < sms / 2 ? quads : sms / 2;
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    const int64_t pairs = (tiles + 1) / 2;
    cfg.gridDim = dim3((unsigned)(pairs < sms ? pairs : sms));
  }
  cudaLaunchKernelEx(&cfg, kern, args);
}
extern "C" int64_t loner_mlp_act_bytes(const loner_net_t* n, int64_t P) {
  Net net;
  if (!net_from(n, net) || P < 0) return -1;
  return n_tiles(P) * (act_tile_bytes(net) + mask_tile_bytes(net));
}
extern "C" int64_t loner_mlp_bwd_scratch_bytes(const loner_net_t* n, int64_t P) {
  Net net;
  if (!net_from(n, net) || P < 0) return -1;
  const WgradPlan p = plan_wgrad(net, wgrad_rebuilds_last(net), wgrad_folds_out(net));
  return n_tiles(P) * dz_tile_bytes(net) + p.total_floats * 4;
}

extern "C" int loner_mlp_pack(const loner_net_t* n, const float* params, void* packed, void* stream) {
  Net net;
  if (!net_from(n, net)) return LONER_E_UNSUPPORTED;
  if (!params || !packed) return LONER_E_BAD_ARG;
  dim3 grid(64, net.L + 1);
  pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(net, params, (uint8_t*)packed);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_mlp_fwd(const loner_net_t* n, const void* packed, const float* pos, const float* rays,
                             const float* z_vals, int32_t S, int64_t P, float* sigma, void* acts, void* stream) {
  Net net;
  if (!net_from(n, net)) return LONER_E_UNSUPPORTED;
  if (P == 0) return LONER_OK;
  if (!packed || !sigma || P < 0 || (!pos && (!rays || !z_vals || S <= 0))) return LONER_E_BAD_ARG;
  FwdArgs a;
  a.net = net; a.packed = (const uint8_t*)packed; a.pos = pos; a.rays = rays; a.z = z_vals; a.S = S; a.P = P;
  a.s_shift = (S > 0 && (S & (S - 1)) == 0) ? __builtin_ctz((unsigned)S) : -1;
  a.tiles = n_tiles(P); a.sigma = sigma; a.acts = (uint8_t*)acts;
  a.masks = acts ? (uint8_t*)acts + a.tiles * act_tile_bytes(net) : nullptr;
  a.stash_last = wgrad_folds_out(net) ? 0 : 1;
  // CTA pairs for the training AND the inference forward: each SM stages half of every weight chunk (half the L2 -> SM
  // stream that bounds the single-CTA inference kernel: 1.35 ms at C2; pairs 1.32 ms since the leader waits on ONE
  // barrier per chunk - with a separate barrier for the peer's half the pair was the slower one, 1.48 ms).
  const int ctas = pipe_ctas(net);
  cudaStream_t st = (cudaStream_t)stream;
  // One MMA-issuing warp per tile for the training forward of CTA pairs (2.009 vs 2.066 ms at C2 in tests/gpu_ab.py,
  // gpurun_out/r2d_ab.jsonl: both tiles share one weight stream, half the L2 -> SM traffic next to the stash copies);
  // inference and single CTAs keep one issuer (two consumers per slot were not faster there: the ring's three 32 KB slots
  // are too few for two tiles in lock-step).
  const bool two = acts != nullptr && ctas == 2 && !(net.flags & LONER_NET_ONE_ISSUER);
#define LONER_FWD_I(W_, S_, I_)                                                                              \
  (ctas == 2 ? launch_pipe(mlp_fwd_kernel<W_, S_, 2, I_>, a, a.tiles, 2, st, pipe_threads<I_>())             \
             : launch_pipe(mlp_fwd_kernel<W_, S_, 1, I_>, a, a.tiles, 1, st, pipe_threads<I_>()))
#define LONER_FWD(W_, S_) (two ? LONER_FWD_I(W_, S_, 2) : LONER_FWD_I(W_, S_, 1))
  if (net.W == 256) { if (acts) LONER_FWD(256, true); else LONER_FWD(256, false); }
  else              { if (acts) LONER_FWD(128, true); else LONER_FWD(128, false); }
#undef LONER_FWD
#undef LONER_FWD_I
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

static int bwd_common(const loner_net_t* n, Net& net, const void* packed, const void* acts, void* scratch, int64_t P) {
  if (!net_from(n, net)) return LONER_E_UNSUPPORTED;
  if (!packed || !acts || !scratch || P < 0) return LONER_E_BAD_ARG;
  return LONER_OK;
}

extern "C" int loner_mlp_dgrad(const loner_net_t* n, const void* packed, const float* pos, const float* rays,
                               const float* z_vals, int32_t S, int64_t P, const float* d_sigma, const void* acts,
                               float grad_scale, float* d_pos, void* scratch, void* stream) {
  Net net;
  int rc = bwd_common(n, net, packed, acts, scratch, P);
  if (rc) return rc;
  if (P == 0) return LONER_OK;
  if (!d_sigma || !(grad_scale > 0.f) || (d_pos && !pos && (!rays || !z_vals || S <= 0))) return LONER_E_BAD_ARG;
  const int64_t tiles = n_tiles(P);
  BwdArgs b;
  b.net = net; b.packed = (const uint8_t*)packed; b.pos = pos; b.rays = rays; b.z = z_vals; b.S = S; b.P = P;
  b.s_shift = (S > 0 && (S & (S - 1)) == 0) ? __builtin_ctz((unsigned)S) : -1;
  b.tiles = tiles; b.d_sigma = d_sigma; b.masks = (const uint8_t*)acts + tiles * act_tile_bytes(net);
  b.dz = (uint8_t*)scratch; b.gscale = grad_scale; b.d_pos = d_pos;
  b.stash_last = wgrad_rebuilds_last(net) ? 0 : 1;
  const int ctas = pipe_ctas(net);
  cudaStream_t st = (cudaStream_t)stream;
  // CTA pairs: one MMA-issuing warp per tile and one weight stream for both tiles, as in the training forward
  const bool two = ctas == 2 && !(net.flags & (LONER_NET_ONE_ISSUER | LONER_NET_DGRAD_ONE_ISSUER));
#define LONER_DGRAD_I(W_, D_, I_)                                                                                \
  (ctas == 2 ? launch_pipe(mlp_dgrad_kernel<W_, D_, 2, I_>, b, tiles, 2, st, pipe_threads<I_>())                 \
             : launch_pipe(mlp_dgrad_kernel<W_, D_, 1, 1>, b, tiles, 1, st, pipe_threads<1>()))
#define LONER_DGRAD(W_, D_) (two ? LONER_DGRAD_I(W_, D_, 2) : LONER_DGRAD_I(W_, D_, 1))
  if (net.W == 256) { if (d_pos) LONER_DGRAD(256, true); else LONER_DGRAD(256, false); }
  else              { if (d_pos) LONER_DGRAD(128, true); else LONER_DGRAD(128, false); }
#undef LONER_DGRAD
#undef LONER_DGRAD_I
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_mlp_wgrad(const loner_net_t* n, const void* packed, int64_t P, const float* d_sigma,
                               const void* acts, float grad_scale, float* d_params, void* scratch, void* stream) {
  Net net;
  int rc = bwd_common(n, net, packed, acts, scratch, P);
  if (rc) return rc;
  if (P == 0) return LONER_OK;
  if (!d_sigma || !d_params || !(grad_scale > 0.f)) return LONER_E_BAD_ARG;
  const int64_t tiles = n_tiles(P);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* dz = (uint8_t*)scratch;
  float* partials = (float*)(dz + tiles * dz_tile_bytes(net));
  const WgradPlan plan = plan_wgrad(net, wgrad_rebuilds_last(net), wgrad_folds_out(net));
  WgradArgs w;
  w.net = net; w.acts = (const uint8_t*)acts; w.dz = dz; w.tiles = tiles; w.partials = partials;
  w.masks = (const uint8_t*)acts + tiles * act_tile_bytes(net); w.d_sigma = d_sigma; w.gen_last = wgrad_rebuilds_last(net) ? 1 : 0;
  w.fold_out = wgrad_folds_out(net) ? 1 : 0; w.packed = (const uint8_t*)packed;
  w.wout = reinterpret_cast<const float*>((const uint8_t*)packed + packed_wout_off(net)); w.gscale = grad_scale; w.P = P;
  for (int i = 0; i <= net.L; ++i) w.item_begin[i] = plan.item_begin[i];
  for (int i = 0; i <= net.L + 1; ++i) w.part_off[i] = plan.part_off[i];
  cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
  mlp_wgrad_kernel<<<(unsigned)plan.total_items, kWgThreads, kWgSmem, st>>>(w);
  LONER_CHECK_LAUNCH();
  dim3 rgrid(64, net.L + 1);
  wgrad_reduce_kernel<<<rgrid, 256, 0, st>>>(net, partials, w, 1.0f / grad_scale, d_params);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_mlp_bwd(const loner_net_t* n, const void* packed, const float* pos, const float* rays,
                             const float* z_vals, int32_t S, int64_t P, const float* d_sigma, const void* acts,
                             float grad_scale, float* d_params, float* d_pos, void* scratch, void* stream) {
  int rc = loner_mlp_dgrad(n, packed, pos, rays, z_vals, S, P, d_sigma, acts, grad_scale, d_pos, scratch, stream);
  if (rc) return rc;
  return loner_mlp_wgrad(n, packed, P, d_sigma, acts, grad_scale, d_params, scratch, stream);
}

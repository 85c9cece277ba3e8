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
#include <cstdlib>
#include "common.cuh"
#include "sm100.cuh"

namespace loner {
namespace mlp {

using namespace sm100;

constexpr int kTile = 128;            // samples per tile = UMMA M
constexpr int kBlk = 16384;           // bytes of one [128 x 64] fp16 column block

struct Net {
  int F, E, Epad, W, L, nb;
};

__host__ inline bool net_from(const loner_net_t* n, Net& o) {
  if (!n) return false;
  o.F = n->n_frequencies; o.W = n->n_neurons; o.L = n->n_hidden_layers;
  if (o.F < 1 || o.F > 10) return false;
  if (!(o.W == 128 || o.W == 256)) return false;
  if (o.L < 1 || o.L > 8) return false;
  o.E = 6 * o.F; o.Epad = (o.E + 15) / 16 * 16; o.nb = o.W / 64;
  return true;
}
__host__ __device__ inline int layer_K(const Net& n, int l) { return l == 0 ? n.Epad : n.W; }
__host__ __device__ inline int64_t packed_off(const Net& n, int l) {      // byte offset of layer l's image
  return l == 0 ? 0 : (int64_t)n.Epad * n.W * 2 + (int64_t)(l - 1) * n.W * n.W * 2;
}
__host__ __device__ inline int64_t packed_wout_off(const Net& n) { return packed_off(n, n.L); }
__host__ __device__ inline int64_t param_off(const Net& n, int l) {       // float offset in flat params
  return l == 0 ? 0 : (int64_t)n.Epad * n.W + (int64_t)(l - 1) * n.W * n.W;
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
      wo[j] = __half2float(__float2half_rn(src[j]));
    return;
  }
  const int K = layer_K(net, l), N = net.W;
  const float* Wm = params + param_off(net, l);
  uint8_t* img_b = packed + packed_off(net, l);
  uint8_t* img_f = packed + packed_fwd_base(net) + fwd_off(net, l);
  const int chunks = K * (N / 8);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += gridDim.x * blockDim.x) {
    const int k = c % K, n0 = (c / K) * 8;
    __half2 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      h[i] = __floats2half2_rn(Wm[(int64_t)(n0 + 2 * i) * K + k], Wm[(int64_t)(n0 + 2 * i + 1) * K + k]);
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
// One CTA per SM, 576 threads:  warp 0 = weight producer (bulk copies into a 3-slot ring),
// warp 1 = MMA issuer, warps 2-9 = epilogue group of tile X, warps 10-17 = epilogue group of tile Y
// (two warps per TMEM lane quarter, each owning half of the columns).
// Two tiles of 128 samples are in flight with one 256-column TMEM accumulator each; the MMA order
// inside a layer is X[c0,c1] Y[c0,c1] X[c2,c3] Y[c2,c3] so that every 32 KB weight chunk is read
// from L2 once per tile PAIR and X's epilogue overlaps Y's last MMAs (and vice versa).
constexpr int kPipeThreads = 576;
constexpr int kGroupThreads = 256;
constexpr int kRingSlots = 3;
constexpr int kSlotBytes = 32768;
constexpr int kPipeSmem = 2 * 65536 + kRingSlots * kSlotBytes + 2304;

struct PipeSmem {
  uint8_t* tileA[2];
  uint8_t* ring;
  float* wout;        // [256]
  float* part;        // [2][128] per-row partial sums handed from the upper column half to the lower
  uint32_t w_full[kRingSlots], w_empty[kRingSlots], a_ready[2], acc_full[2];
  uint32_t* tmem_slot;
};

__device__ __forceinline__ PipeSmem carve(uint8_t* base) {
  PipeSmem p;
  p.tileA[0] = base;
  p.tileA[1] = base + 65536;
  p.ring = base + 131072;
  p.wout = reinterpret_cast<float*>(base + 131072 + kRingSlots * kSlotBytes);
  p.part = p.wout + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p.part + 256);
  for (int i = 0; i < kRingSlots; ++i) { p.w_full[i] = smem_u32(&bars[i]); p.w_empty[i] = smem_u32(&bars[kRingSlots + i]); }
  p.a_ready[0] = smem_u32(&bars[2 * kRingSlots]); p.a_ready[1] = smem_u32(&bars[2 * kRingSlots + 1]);
  p.acc_full[0] = smem_u32(&bars[2 * kRingSlots + 2]); p.acc_full[1] = smem_u32(&bars[2 * kRingSlots + 3]);
  p.tmem_slot = reinterpret_cast<uint32_t*>(&bars[2 * kRingSlots + 4]);
  return p;
}

__device__ __forceinline__ void pipe_init(const PipeSmem& sm, int tid, int warp, const float* wout_src, int W) {
  if (warp == 1) tmem_alloc<512>(smem_u32(sm.tmem_slot));
  if (tid == 0) {
    for (int i = 0; i < kRingSlots; ++i) { mbar_init(sm.w_full[i], 1); mbar_init(sm.w_empty[i], 1); }
    for (int t = 0; t < 2; ++t) { mbar_init(sm.a_ready[t], kGroupThreads); mbar_init(sm.acc_full[t], 1); }
    fence_mbar_init();
  }
  for (int j = tid; j < W; j += kPipeThreads) sm.wout[j] = wout_src[j];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

__device__ __forceinline__ void group_bar(int group) {
  asm volatile("bar.sync %0, 256;" ::"r"(group + 1) : "memory");
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
// bits |= (v > 0) << kBit in two instructions (FSETP + predicated LOP3)
template <int kBit>
__device__ __forceinline__ void set_bit_if_pos(uint32_t& bits, float v) {
  asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, 0f00000000;\n\t@p or.b32 %0, %0, %2;\n\t}"
      : "+r"(bits)
      : "f"(v), "n"(1u << kBit));
}

__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}

// One 32-column slab of an epilogue: `pack(i)` yields the packed half2 of this row's columns
// (col0+i, col0+i+1).  The four 16-byte chunks go to the shared-memory tile image (`srow` = the
// row's base inside the tile, or null) and/or straight to the image in HBM (`grow`, or null) as
// two 32-byte sectors: chunks 2m and 2m+1 share a sector, swapped when the row's swizzle is odd.
// xs = (row & 7) << 4.
template <class F>
__device__ __forceinline__ void store_slab(uint32_t srow, uint8_t* grow, uint32_t xs, int col0, F&& pack) {
  const bool odd = (xs & 16u) != 0u;
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    uint32_t a[4], b[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { a[e] = pack(m * 16 + 2 * e); b[e] = pack(m * 16 + 8 + 2 * e); }
    const int c0 = col0 + m * 16;                                        // first column of chunk 2m'
    const uint32_t cb_off = (uint32_t)(c0 >> 6) * kBlk;
    const uint32_t j0 = ((uint32_t)(c0 & 63) >> 3) << 4;                  // logical chunk byte offset (even chunk)
    if (srow != 0u) {
      sts128(srow + cb_off + (j0 ^ xs), a[0], a[1], a[2], a[3]);
      sts128(srow + cb_off + ((j0 + 16u) ^ xs), b[0], b[1], b[2], b[3]);
    }
    if (grow != nullptr) {
      uint32_t w[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { w[e] = odd ? b[e] : a[e]; w[4 + e] = odd ? a[e] : b[e]; }
      st_global_256(grow + cb_off + (j0 ^ (xs & 0x60u)), w);
    }
  }
}

// Half a slab: 16 columns = chunks (2m, 2m+1) = one 32-byte sector of the image.  hh[8] = packed half2.
__device__ __forceinline__ void store16(uint32_t srow, uint8_t* grow, uint32_t xs, int c0, const uint32_t (&hh)[8]) {
  const uint32_t cb_off = (uint32_t)(c0 >> 6) * kBlk;
  const uint32_t j0 = ((uint32_t)(c0 & 63) >> 3) << 4;
  if (srow != 0u) {
    sts128(srow + cb_off + (j0 ^ xs), hh[0], hh[1], hh[2], hh[3]);
    sts128(srow + cb_off + ((j0 + 16u) ^ xs), hh[4], hh[5], hh[6], hh[7]);
  }
  if (grow != nullptr) {
    const bool odd = (xs & 16u) != 0u;
    uint32_t w[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) { w[e] = odd ? hh[4 + e] : hh[e]; w[4 + e] = odd ? hh[e] : hh[4 + e]; }
    st_global_256(grow + cb_off + (j0 ^ (xs & 0x60u)), w);
  }
}

// Copies this thread's kCols columns of its row from the shared-memory tile image to the image in HBM
// (32-byte sectors).  Runs AFTER the tile has been handed to the MMA warp, so that store back-pressure
// from HBM never sits between an epilogue and the next layer's tensor-core work.
template <int kCols>
__device__ __forceinline__ void copy_out(uint32_t srow, uint8_t* grow, uint32_t xs, int col_begin) {
  const bool odd = (xs & 16u) != 0u;
#pragma unroll
  for (int i = 0; i < kCols / 16; ++i) {
    const int c0 = col_begin + i * 16;
    const uint32_t cb_off = (uint32_t)(c0 >> 6) * kBlk;
    const uint32_t j0 = ((uint32_t)(c0 & 63) >> 3) << 4;
    const uint4 a = lds128(srow + cb_off + (j0 ^ xs));
    const uint4 b = lds128(srow + cb_off + ((j0 + 16u) ^ xs));
    const uint4 lo = odd ? b : a, hi = odd ? a : b;
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    st_global_256(grow + cb_off + (j0 ^ (xs & 0x60u)), w);
  }
}

// Linear, fully coalesced copy of a finished tile image (contiguous in shared memory and in HBM) by
// the 256 threads of an epilogue group: every warp instruction moves 512 contiguous bytes.
__device__ __forceinline__ void copy_image(uint32_t simg, uint8_t* gimg, int bytes, int gtid) {
  for (int off = gtid * 16; off < bytes; off += kGroupThreads * 16) {
    const uint4 v = lds128(simg + off);
    asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(gimg + off), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
  }
}

// Drains kCols accumulator columns of this thread's TMEM lane in 16-column loads, keeping the next
// load in flight while `proc(i, v)` works on the current one (tcgen05.ld latency is ~300-450 clk,
// tests/gpu_probe.py).
template <int kCols, class Proc>
__device__ __forceinline__ void drain_cols(uint32_t acc, Proc&& proc) {
  uint32_t va[16], vb[16];
  tmem_ld16(acc, va);
  tmem_ld_wait16(va);
#pragma unroll
  for (int i = 0; i < kCols / 16; ++i) {
    if (i & 1) {
      if (i + 1 < kCols / 16) { tmem_ld16(acc + (i + 1) * 16, va); pin16(vb); }
      proc(i, vb);
      if (i + 1 < kCols / 16) tmem_ld_wait16(va);
    } else {
      if (i + 1 < kCols / 16) { tmem_ld16(acc + (i + 1) * 16, vb); pin16(va); }
      proc(i, va);
      if (i + 1 < kCols / 16) tmem_ld_wait16(vb);
    }
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
  int bulk;              // stash images leave through cp.async.bulk (1) or a thread copy (0)
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

// Writes encoded features [32*half, 32*half+32) of row r (column block 0 of `sA`): features
// [0, 6F) are sin/cos pairs ordered [dim][freq][sin,cos], [6F, Epad) = 1.0 (tcnn pads the encoded
// width to 16 with ones), rest 0.  (dim0, f0) = position of feature pair 16*half, precomputed.
__device__ __forceinline__ void encode_row(uint32_t sA, uint8_t* grow, int r, int half, const float (&x)[3],
                                           const Net& net, int dim0, int f0) {
  int dim = dim0, f = f0;
  float scale = (float)(1 << f0);
  uint32_t keep[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int cj = half * 4 + c;
    __half2 h[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int pair = cj * 4 + q;            // features 2*pair, 2*pair+1
      float s, co;
      if (dim < 3) {
        const float xv = dim == 0 ? x[0] : (dim == 1 ? x[1] : x[2]);
        freq_pair(xv, scale, s, co);
      } else if (2 * pair < net.Epad) {
        s = 1.0f; co = 1.0f;
      } else {
        s = 0.0f; co = 0.0f;
      }
      h[q] = __floats2half2_rn(s, co);
      ++f; scale *= 2.0f;
      if (f == net.F) { f = 0; scale = 1.0f; ++dim; }
    }
    {
      const uint32_t* hw0 = reinterpret_cast<const uint32_t*>(h);
      sts128(sA + r * 128 + ((cj ^ (r & 7)) * 16), hw0[0], hw0[1], hw0[2], hw0[3]);
    }
    if (grow != nullptr) {                 // chunks cj (even) and cj+1 share one 32-byte sector of the image
      const uint32_t* hw = reinterpret_cast<const uint32_t*>(h);
      if ((c & 1) == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) keep[e] = hw[e];
      } else {
        const bool odd = (r & 1) != 0;
        uint32_t w[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) { w[e] = odd ? hw[e] : keep[e]; w[4 + e] = odd ? keep[e] : hw[e]; }
        st_global_256(grow + (((cj - 1) ^ (r & 6)) * 16), w);
      }
    }
  }
}

template <int W, bool kStash>
__global__ void __launch_bounds__(kPipeThreads, 1) mlp_fwd_kernel(const FwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const PipeSmem sm = carve(smem_raw);
  const Net net = a.net;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pipe_init(sm, tid, warp, reinterpret_cast<const float*>(a.packed + packed_wout_off(net)), W);
  const uint32_t tmem = *sm.tmem_slot;
  const int64_t pairs = (a.tiles + 1) / 2;
  constexpr int kNb = W / 64;
  constexpr uint32_t kChunkBytes = kNb * 8192;      // 64 K-rows x W out-features, fp16

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- producer: one contiguous bulk copy per 64-row weight chunk
      const uint8_t* fimg = a.packed + packed_fwd_base(net);
      uint32_t g = 0;
      for (int64_t pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
        for (int l = 0; l < net.L; ++l) {
          const int nch = (l == 0) ? 1 : kNb;
          for (int c = 0; c < nch; ++c, ++g) {
            const uint32_t slot = g % kRingSlots, use = g / kRingSlots;
            if (use > 0) mbar_wait(sm.w_empty[slot], (use - 1) & 1);
            mbar_expect_tx(sm.w_full[slot], kChunkBytes);
            bulk_g2s(smem_u32(sm.ring + slot * kSlotBytes), fimg + fwd_off(net, l) + (int64_t)c * kChunkBytes,
                     kChunkBytes, sm.w_full[slot]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer
      constexpr uint32_t idesc = make_idesc_f16(128, W, 0, 1);
      uint32_t g_base = 0, par_a[2] = {0u, 0u};
      for (int64_t pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
        for (int l = 0; l < net.L; ++l) {
          const int nch = (l == 0) ? 1 : kNb;
          const int ksteps0 = (l == 0) ? net.Epad / 16 : 4;   // layer 0 contracts over Epad (<= 64) features
          for (int c0 = 0; c0 < nch; c0 += 2) {
            const int c1 = min(c0 + 2, nch);
            for (int t = 0; t < 2; ++t) {
              if (c0 == 0) { mbar_wait(sm.a_ready[t], par_a[t]); par_a[t] ^= 1u; tc_fence_after(); }
              for (int c = c0; c < c1; ++c) {
                const uint32_t g = g_base + c, slot = g % kRingSlots;
                if (t == 0) { mbar_wait(sm.w_full[slot], (g / kRingSlots) & 1); tc_fence_after(); }
                const uint32_t sa = smem_u32(sm.tileA[t]) + c * kBlk, sb = smem_u32(sm.ring + slot * kSlotBytes);
                for (int ks = 0; ks < ksteps0; ++ks) {
                  const uint64_t ad = make_desc_sw128(sa + ks * 32, 16, 1024);
                  const uint64_t bd = make_desc_sw128(sb + ks * 2048, 8192, 1024);
                  umma_f16(tmem + t * 256, ad, bd, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                }
                if (t == 1) umma_commit(sm.w_empty[slot]);     // both tiles have consumed this chunk
              }
              if (c1 == nch) umma_commit(sm.acc_full[t]);
            }
          }
          g_base += nch;
        }
      }
    }
  } else {
    // ---------------- epilogue groups: thread = (sample row, column half)
    const int e = warp - 2;
    const int t = e >> 3;
    const int h = (e & 7) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t sA = smem_u32(sm.tileA[t]);
    const uint32_t srow = sA + row * 128;
    const bool elected = ((e & 7) == 0) && lane == 0;
    const uint32_t xs = (uint32_t)(row & 7) << 4;
    constexpr int kCols = W / 2;                      // columns per thread
    const uint32_t acc = tmem + t * 256 + ((uint32_t)(q * 32) << 16) + h * kCols;
    float* part = sm.part + t * 128;
    uint32_t par_acc = 0;
    const int enc_dim0 = (16 * h) / net.F, enc_f0 = (16 * h) % net.F;
    const bool pos_mode = a.pos != nullptr;
    auto row_index = [&](int64_t pair) {
      int64_t gs = (2 * pair + t) * kTile + row;
      return gs < a.P ? gs : a.P - 1;
    };
    RowIn nxt = load_row(a.pos, a.rays, a.z, a.S, a.s_shift, row_index(blockIdx.x));
    for (int64_t pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
      const int64_t tile = 2 * pair + t;
      const bool active = tile < a.tiles;
      const int64_t gs = tile * kTile + row;
      const bool in = active && gs < a.P;
      uint8_t* gtile = (kStash && active) ? a.acts + tile * act_tile_bytes(net) + row * 128 : nullptr;
      if (kStash) { if (a.bulk && elected) bulk_wait_read0(); group_bar(t); }   // previous image has left sA
      {
        float x[3];
        row_pos01(pos_mode, nxt, x);
        encode_row(sA, gtile, row, h, x, net, enc_dim0, enc_f0);
      }
      fence_async_smem();
      mbar_arrive(sm.a_ready[t]);
      if (pair + gridDim.x < pairs) nxt = load_row(a.pos, a.rays, a.z, a.S, a.s_shift, row_index(pair + gridDim.x));
      for (int l = 0; l < net.L; ++l) {
        const bool last = (l == net.L - 1);
        uint8_t* grow = gtile ? gtile + kBlk + (int64_t)l * kNb * kBlk : nullptr;
        mbar_wait(sm.acc_full[t], par_acc);
        par_acc ^= 1u;
        tc_fence_after();
        if (kStash && l > 0) { if (a.bulk && elected) bulk_wait_read0(); group_bar(t); }   // previous image has left sA
        float sig0 = 0.f, sig1 = 0.f;
        uint32_t mbits[kCols / 32];
#pragma unroll
        for (int it = 0; it < kCols / 32; ++it) mbits[it] = 0u;
        drain_cols<kCols>(acc, [&](int i, const uint32_t (&v)[16]) {
          const int col0 = h * kCols + i * 16;
          if (kStash) {
            uint32_t b0 = 0u, b1 = 0u;      // two independent chains
#define LONER_BIT(B, I) set_bit_if_pos<I>(B, __uint_as_float(v[I]));
            LONER_BIT(b0, 0) LONER_BIT(b1, 8) LONER_BIT(b0, 1) LONER_BIT(b1, 9) LONER_BIT(b0, 2) LONER_BIT(b1, 10)
            LONER_BIT(b0, 3) LONER_BIT(b1, 11) LONER_BIT(b0, 4) LONER_BIT(b1, 12) LONER_BIT(b0, 5) LONER_BIT(b1, 13)
            LONER_BIT(b0, 6) LONER_BIT(b1, 14) LONER_BIT(b0, 7) LONER_BIT(b1, 15)
#undef LONER_BIT
            mbits[i >> 1] |= (b0 | b1) << ((i & 1) * 16);
          }
          uint32_t hh[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) hh[e] = cvt_relu_h2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
          if (!last || kStash) store16(srow, nullptr, xs, col0, hh);
          if (last) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float2 r2 = __half22float2(*reinterpret_cast<const __half2*>(&hh[e]));
              const float2 w2 = *reinterpret_cast<const float2*>(sm.wout + col0 + 2 * e);
              sig0 = fmaf(r2.x, w2.x, sig0);
              sig1 = fmaf(r2.y, w2.y, sig1);
            }
          }
        });
        const float sig = sig0 + sig1;
        tc_fence_before();
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
        if (!last) {
          fence_async_smem();
          mbar_arrive(sm.a_ready[t]);
        } else {
          if (kStash) fence_async_smem();       // A_L in sA becomes visible to the bulk-copy engine
          if (h == 1) part[row] = sig;
          group_bar(t);
          if (h == 0 && in) a.sigma[gs] = sig + part[row];
        }
        if (kStash) {
          if (!last) group_bar(t);            // (the last layer already met at the sigma barrier)
          uint8_t* gimg = a.acts + tile * act_tile_bytes(net) + kBlk + (int64_t)l * kNb * kBlk;
          if (active && !a.bulk) copy_image(sA, gimg, kNb * kBlk, (e & 7) * 32 + lane);
          if (active && a.bulk && elected) { bulk_s2g(gimg, sA, (uint32_t)(kNb * kBlk)); bulk_commit(); }
        }
      }
    }
    if (kStash && a.bulk && elected) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
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
  int bulk;
  float gscale;
  float* d_pos;          // [P,3] or null
};

template <int W>
__global__ void __launch_bounds__(kPipeThreads, 1) mlp_dgrad_kernel(const BwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const PipeSmem sm = carve(smem_raw);
  const Net net = a.net;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pipe_init(sm, tid, warp, reinterpret_cast<const float*>(a.packed + packed_wout_off(net)), W);
  const uint32_t tmem = *sm.tmem_slot;
  const int64_t pairs = (a.tiles + 1) / 2;
  const bool want_dx = a.d_pos != nullptr;
  const int l_lo = want_dx ? 0 : 1;        // GEMMs run for l = L-1 .. l_lo : dA_l = dZ_{l+1} * W_l
  constexpr int kNb = W / 64;              // contraction (out-features of layer l) in 64-wide chunks

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0;
      for (int64_t pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
        for (int l = net.L - 1; l >= l_lo; --l) {
          const uint32_t bytes = (uint32_t)layer_K(net, l) * 128u;     // one column block: K_l rows x 128 B
          for (int c = 0; c < kNb; ++c, ++g) {
            const uint32_t slot = g % kRingSlots, use = g / kRingSlots;
            if (use > 0) mbar_wait(sm.w_empty[slot], (use - 1) & 1);
            mbar_expect_tx(sm.w_full[slot], bytes);
            bulk_g2s(smem_u32(sm.ring + slot * kSlotBytes), a.packed + packed_off(net, l) + (int64_t)c * bytes, bytes,
                     sm.w_full[slot]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t g_base = 0, par_a[2] = {0u, 0u};
      for (int64_t pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
        for (int l = net.L - 1; l >= l_lo; --l) {
          const uint32_t idesc = make_idesc_f16(128, layer_K(net, l), 0, 0);
          for (int c0 = 0; c0 < kNb; c0 += 2) {
            const int c1 = min(c0 + 2, kNb);
            for (int t = 0; t < 2; ++t) {
              if (c0 == 0) { mbar_wait(sm.a_ready[t], par_a[t]); par_a[t] ^= 1u; tc_fence_after(); }
              for (int c = c0; c < c1; ++c) {
                const uint32_t g = g_base + c, slot = g % kRingSlots;
                if (t == 0) { mbar_wait(sm.w_full[slot], (g / kRingSlots) & 1); tc_fence_after(); }
                const uint32_t sa = smem_u32(sm.tileA[t]) + c * kBlk, sb = smem_u32(sm.ring + slot * kSlotBytes);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  const uint64_t ad = make_desc_sw128(sa + ks * 32, 16, 1024);
                  const uint64_t bd = make_desc_sw128(sb + ks * 32, 16, 1024);
                  umma_f16(tmem + t * 256, ad, bd, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                }
                if (t == 1) umma_commit(sm.w_empty[slot]);
              }
              if (c1 == kNb) umma_commit(sm.acc_full[t]);
            }
          }
          g_base += kNb;
        }
      }
    }
  } else {
    const int e = warp - 2;
    const int t = e >> 3;
    const int h = (e & 7) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t stile = smem_u32(sm.tileA[t]);
    const uint32_t srow = stile + row * 128;
    const bool elected = ((e & 7) == 0) && lane == 0;
    const uint32_t xs = (uint32_t)(row & 7) << 4;
    constexpr int kCols = W / 2;
    const uint32_t acc_row = tmem + t * 256 + ((uint32_t)(q * 32) << 16);
    const uint32_t acc = acc_row + h * kCols;
    uint32_t par_acc = 0;
    constexpr int kWords = W / 32;
    const bool pos_mode = a.pos != nullptr;
    for (int64_t pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
      const int64_t tile = 2 * pair + t;
      const bool active = tile < a.tiles;
      const int64_t gs = tile * kTile + row;
      const bool in = active && gs < a.P;
      const uint32_t* mtile = reinterpret_cast<const uint32_t*>(a.masks + (active ? tile : 0) * mask_tile_bytes(net));
      uint8_t* gtile = active ? a.dz + tile * dz_tile_bytes(net) + row * 128 : nullptr;
      if (a.bulk && elected) bulk_wait_read0();
      group_bar(t);                           // the previous tile's last image copy has left the buffer
      RowIn rin;
      if (want_dx && h == 0) rin = load_row(a.pos, a.rays, a.z, a.S, a.s_shift, in ? gs : a.P - 1);
      {  // dZ_L = d_sigma * w_out * relu'(Z_L)
        const float ds = in ? a.d_sigma[gs] * a.gscale : 0.f;
        const uint32_t* mrow = mtile + ((int64_t)(net.L - 1) * kTile + row) * kWords + h * (kCols / 32);
#pragma unroll
        for (int it = 0; it < kCols / 32; ++it) {
          const uint32_t bits = mrow[it];
          const int col0 = h * kCols + it * 32;
          store_slab(srow, nullptr, xs, col0, [&](int i0) {
            const float2 w2 = *reinterpret_cast<const float2*>(sm.wout + col0 + i0);
            const float g0 = ((bits >> i0) & 1u) ? ds * w2.x : 0.f;
            const float g1 = ((bits >> (i0 + 1)) & 1u) ? ds * w2.y : 0.f;
            return cvt_sat_h2(g0, g1);
          });
        }
      }
      fence_async_smem();
      mbar_arrive(sm.a_ready[t]);
      // (the dZ_L image is NOT stashed: wgrad rebuilds it from the masks, d_sigma and w_out)
      for (int l = net.L - 1; l >= l_lo; --l) {
        mbar_wait(sm.acc_full[t], par_acc);
        par_acc ^= 1u;
        tc_fence_after();
        if (a.bulk && elected) bulk_wait_read0();
        group_bar(t);                         // everyone's copy of the previous image has left the tile buffer
        if (l >= 1) {
          // dZ_l = dA_l * relu'(Z_l)  -> fp16 image (next GEMM's A operand in smem, wgrad's B operand in HBM)
          const uint32_t* mrow = mtile + ((int64_t)(l - 1) * kTile + row) * kWords + h * (kCols / 32);
          uint8_t* grow = gtile ? gtile + (int64_t)(l - 1) * kNb * kBlk : nullptr;
          const bool feeds_gemm = (l - 1 >= l_lo);
          uint32_t mw[kCols / 32];
#pragma unroll
          for (int it = 0; it < kCols / 32; ++it) mw[it] = mrow[it];
          drain_cols<kCols>(acc, [&](int i, const uint32_t (&v)[16]) {
            const uint32_t bits = mw[i >> 1] >> ((i & 1) * 16);
            uint32_t hh[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float g0 = ((bits >> (2 * e)) & 1u) ? __uint_as_float(v[2 * e]) : 0.f;
              const float g1 = ((bits >> (2 * e + 1)) & 1u) ? __uint_as_float(v[2 * e + 1]) : 0.f;
              hh[e] = cvt_sat_h2(g0, g1);
            }
            store16(srow, nullptr, xs, h * kCols + i * 16, hh);
          });
          tc_fence_before();
          fence_async_smem();
          if (feeds_gemm) mbar_arrive(sm.a_ready[t]);
          group_bar(t);
          {
            uint8_t* gimg = a.dz + tile * dz_tile_bytes(net) + (int64_t)(l - 1) * kNb * kBlk;
            if (active && !a.bulk) copy_image(stile, gimg, kNb * kBlk, (e & 7) * 32 + lane);
            if (active && a.bulk && elected) { bulk_s2g(gimg, stile, (uint32_t)(kNb * kBlk)); bulk_commit(); }
          }
        } else if (h == 0) {
          // l == 0: dEnc [128 x Epad] -> d_pos through the sin/cos encoding (lower-half warps only)
          float x[3];
          row_pos01(pos_mode, rin, x);
          float dx[3] = {0.f, 0.f, 0.f};
          int dim = 0, f = 0;
          float scale = 1.0f;
          for (int it = 0; it * 32 < net.Epad; ++it) {
            uint32_t v[32];
            tmem_ld32(acc_row + it * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int pq = 0; pq < 16; ++pq) {
              if (dim < 3) {
                const float xv = dim == 0 ? x[0] : (dim == 1 ? x[1] : x[2]);
                float sn, cs;
                freq_pair(xv, scale, sn, cs);
                // d/dx sin(pi 2^f x) = pi 2^f cos, d/dx cos = -pi 2^f sin
                const float g = (__uint_as_float(v[2 * pq]) * cs - __uint_as_float(v[2 * pq + 1]) * sn) *
                                (3.14159265358979323846f * scale);
                if (dim == 0) dx[0] += g; else if (dim == 1) dx[1] += g; else dx[2] += g;
              }
              ++f; scale *= 2.0f;
              if (f == net.F) { f = 0; scale = 1.0f; ++dim; }
            }
          }
          tc_fence_before();
          if (in) {
            const float inv = 0.5f / a.gscale;      // x = (pos + 1) / 2
            a.d_pos[gs * 3 + 0] = dx[0] * inv;
            a.d_pos[gs * 3 + 1] = dx[1] * inv;
            a.d_pos[gs * 3 + 2] = dx[2] * inv;
          }
        } else {
          tc_fence_before();
        }
      }
    }
    if (a.bulk && elected) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// wgrad: dW_l[n,k] = sum_s dZ_{l+1}[s,n] * A_l[s,k], one CTA per (layer, range of tiles),
// accumulators resident in TMEM over the whole range, operands streamed in half tiles.
struct WgradArgs {
  Net net;
  const uint8_t* acts;
  const uint8_t* dz;
  int64_t tiles;
  float* partials;           // [items][K_l*N_l]
  // dZ_L = relu'(Z_L) * d_sigma * w_out is rank-1 times a bit mask: the CTAs of layer L-1 rebuild its
  // image in shared memory from these instead of reading 64 KB/tile that dgrad would have to write.
  const uint8_t* masks;
  const float* d_sigma;
  const float* wout;         // [W] fp32 values of the fp16-rounded output weights (packed image)
  float gscale;
  int64_t P;
  int item_begin[9];         // first item of each layer (prefix), item_begin[L] = total
  int64_t part_off[9];       // float offset of each layer's first partial
};

constexpr int kWgStages = 3;
constexpr int kWgStageBytes = 65536;    // 64 samples: A half (<= 32 KB) | dZ half (<= 32 KB)
constexpr int kWgThreads = 320;     // warp 0 producer, warp 1 MMA, warps 2-5 epilogue (+ generators), warps 6-9 generators
constexpr int kWgGen = 256;         // generator threads
constexpr int kWgSmem = 1024 + kWgStages * kWgStageBytes + 2048;

__global__ void __launch_bounds__(kWgThreads, 1) mlp_wgrad_kernel(const WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  float* s_wout = reinterpret_cast<float*>(base + kWgStages * kWgStageBytes);          // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_wout + 256);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kWgStages + 1);
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

  const bool gen_y = (l == net.L - 1);           // this CTA rebuilds dZ_L instead of loading it
  if (warp == 0) tmem_alloc<512>(smem_u32(s_tmem));
  if (tid == 32) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(full_bar(s), gen_y ? 1 + kWgGen : 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  for (int j = tid; j < net.W; j += kWgThreads) s_wout[j] = a.wout[j];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  // per-stage operand geometry.  "X" = A_l image (64-sample half), "Y" = dZ_{l+1} image half.
  const int nbA = (l == 0) ? 1 : net.nb;          // column blocks of A_l
  const int nbY = net.nb;
  const uint32_t bytesA = (uint32_t)nbA * 8192, bytesY = (uint32_t)nbY * 8192;
  const int64_t actA_off = (l == 0) ? 0 : (int64_t)kBlk + (int64_t)(l - 1) * net.nb * kBlk;
  const int64_t dzY_off = (int64_t)l * net.nb * kBlk;
  const int64_t n_half = (t1 - t0) * 2;

  if (tid == 0) {
    // ---- producer: bulk loads, one column block (64 rows x 128 B = 8 KB) per copy
    for (int64_t i = 0; i < n_half; ++i) {
      const int s = (int)(i % kWgStages);
      const uint32_t ph = (uint32_t)((i / kWgStages) & 1);
      mbar_wait(empty_bar(s), ph ^ 1u);
      const int64_t tile = t0 + (i >> 1);
      const int hf = (int)(i & 1);
      const uint32_t dstA = smem_u32(base + s * kWgStageBytes), dstY = dstA + 32768;
      mbar_expect_tx(full_bar(s), gen_y ? bytesA : bytesA + bytesY);
      const uint8_t* srcA = a.acts + tile * act_tile_bytes(net) + actA_off + hf * 8192;
      const uint8_t* srcY = a.dz + tile * dz_tile_bytes(net) + dzY_off + hf * 8192;
      for (int cb = 0; cb < nbA; ++cb) bulk_g2s(dstA + cb * 8192, srcA + (int64_t)cb * kBlk, 8192, full_bar(s));
      if (!gen_y)
        for (int cb = 0; cb < nbY; ++cb) bulk_g2s(dstY + cb * 8192, srcY + (int64_t)cb * kBlk, 8192, full_bar(s));
    }
  } else if (tid == 32) {
    // ---- MMA issuer.  Both operands MN-major: 64-element MN blocks 8 KB apart (LBO), 8-sample
    // groups 1 KB apart (SBO), 16 samples per instruction = 2 KB per k-step.
    const bool swapped = (l == 0);   // layer 0: M = out-features (from dZ), N = Epad (from A_0)
    const int n_mblk = swapped ? net.W / 128 : K / 128;
    const int Ncols = swapped ? net.Epad : net.W;
    const uint32_t idesc = make_idesc_f16(128, Ncols, 1, 1);
    for (int64_t i = 0; i < n_half; ++i) {
      const int s = (int)(i % kWgStages);
      const uint32_t ph = (uint32_t)((i / kWgStages) & 1);
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t sX = smem_u32(base + s * kWgStageBytes), sY = sX + 32768;
      for (int mb = 0; mb < n_mblk; ++mb) {
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t aaddr = (swapped ? sY : sX) + mb * 2 * 8192 + ks * 2048;
          const uint32_t baddr = (swapped ? sX : sY) + ks * 2048;
          const uint64_t ad = make_desc_sw128(aaddr, 8192, 1024);
          const uint64_t bd = make_desc_sw128(baddr, 8192, 1024);
          umma_f16(tmem + mb * 256, ad, bd, idesc, (i > 0 || ks > 0) ? 1u : 0u);
        }
      }
      umma_commit(empty_bar(s));
    }
    umma_commit(done_bar);
  }
  __syncwarp();
  if (warp >= 2 && gen_y) {
    // ---- generators (8 warps; four of them are the epilogue warps, idle during the main loop):
    // thread = (row of the 64-sample half tile, quarter of the columns)
    const int g = tid - 64;                       // 0..255
    const int r = g >> 2, qc = g & 3;
    const int cpt = net.W / 4;                    // columns per thread: 64 (W=256) or 32 (W=128)
    const int mwords = net.W / 32;
    for (int64_t i = 0; i < n_half; ++i) {
      const int s = (int)(i % kWgStages);
      const uint32_t ph = (uint32_t)((i / kWgStages) & 1);
      const int64_t tile = t0 + (i >> 1);
      const int row = (int)(i & 1) * 64 + r;      // row inside the 128-sample tile
      const int64_t gs = tile * kTile + row;
      const float ds = gs < a.P ? __ldg(a.d_sigma + gs) * a.gscale : 0.f;
      const uint32_t* mrow = reinterpret_cast<const uint32_t*>(a.masks + tile * mask_tile_bytes(net)) +
                             ((int64_t)(net.L - 1) * kTile + row) * mwords + qc * (cpt / 32);
      const uint32_t mw0 = __ldg(mrow), mw1 = (cpt > 32) ? __ldg(mrow + 1) : 0u;
      mbar_wait(empty_bar(s), ph ^ 1u);
      const uint32_t sY = smem_u32(base + s * kWgStageBytes) + 32768 + (uint32_t)r * 128;
      const uint32_t xs = (uint32_t)(r & 7) << 4;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        if (it * 32 < cpt) {
          const uint32_t bits = it == 0 ? mw0 : mw1;
          const int col0 = qc * cpt + it * 32;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i0 = ch * 8 + 2 * e;
              const float2 w2 = *reinterpret_cast<const float2*>(s_wout + col0 + i0);
              const float g0 = ((bits >> i0) & 1u) ? ds * w2.x : 0.f;
              const float g1 = ((bits >> (i0 + 1)) & 1u) ? ds * w2.y : 0.f;
              w[e] = cvt_sat_h2(g0, g1);
            }
            const int c0 = col0 + ch * 8;
            sts128(sY + (uint32_t)(c0 >> 6) * 8192u + ((((uint32_t)(c0 & 63) >> 3) << 4) ^ xs), w[0], w[1], w[2], w[3]);
          }
        }
      }
      fence_async_smem();
      mbar_arrive(full_bar(s));
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

// d_params += (sum over a layer's items of its partials) / gscale
__global__ void wgrad_reduce_kernel(Net net, const float* __restrict__ partials, WgradArgs w, float inv_gscale,
                                    float* __restrict__ d_params) {
  const int l = blockIdx.y;
  const int K = layer_K(net, l);
  const int64_t sz = (int64_t)K * net.W;
  const int n_items = w.item_begin[l + 1] - w.item_begin[l];
  const float* p = partials + w.part_off[l];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < sz; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int g = 0; g < n_items; ++g) s += p[(int64_t)g * sz + i];
    d_params[param_off(net, l) + i] += s * inv_gscale;
  }
}

// dW_out[j] = sum_s d_sigma[s] * A_L[s,j]   (row 0 of the padded [16, W] output matrix).
// HBM-streaming: a warp reads 4 image rows (4 x 128 B) per instruction; lane = (row%4, 16-byte
// logical chunk j), so each lane owns 8 fixed columns per column block.
__global__ void __launch_bounds__(256) dwout_kernel(Net net, const uint8_t* __restrict__ acts,
                                                    const float* __restrict__ d_sigma, int64_t P, int64_t tiles,
                                                    float* __restrict__ d_params) {
  __shared__ float red[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * 8 + warp, nw = (int64_t)gridDim.x * 8;
  const int j = lane & 7, rsub = lane >> 3;
  float acc[4][8];
#pragma unroll
  for (int cb = 0; cb < 4; ++cb)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[cb][i] = 0.f;
  const int64_t aL = (int64_t)kBlk + (int64_t)(net.L - 1) * net.nb * kBlk;
  for (int64_t tile = gw; tile < tiles; tile += nw) {
    const uint8_t* img = acts + tile * act_tile_bytes(net) + aL;
#pragma unroll 4
    for (int r0 = 0; r0 < kTile; r0 += 4) {
      const int r = r0 + rsub;
      const int64_t gs = tile * kTile + r;
      const float ds = gs < P ? __ldg(d_sigma + gs) : 0.f;
      const int slot = (j ^ (r & 7)) * 16;
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        if (cb < net.nb) {
          const uint4 v = *reinterpret_cast<const uint4*>(img + cb * kBlk + r * 128 + slot);
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            acc[cb][2 * i] = fmaf(ds, f.x, acc[cb][2 * i]);
            acc[cb][2 * i + 1] = fmaf(ds, f.y, acc[cb][2 * i + 1]);
          }
        }
      }
    }
  }
  for (int c = threadIdx.x; c < 256; c += blockDim.x) red[c] = 0.f;
  __syncthreads();
#pragma unroll
  for (int cb = 0; cb < 4; ++cb)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = acc[cb][i];
      v += __shfl_xor_sync(kFull, v, 8);
      v += __shfl_xor_sync(kFull, v, 16);
      if (rsub == 0 && cb < net.nb) atomicAdd(&red[cb * 64 + j * 8 + i], v);
    }
  __syncthreads();
  for (int c = threadIdx.x; c < net.W; c += blockDim.x)
    if (red[c] != 0.f) atomicAdd(d_params + param_off(net, net.L) + c, red[c]);
}

// ------------------------------------------------------------------------------------------
struct WgradPlan {
  int item_begin[9];
  int64_t part_off[9];
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

inline WgradPlan plan_wgrad(const Net& net) {
  // The kernel is HBM-bound (it streams the A_l and dZ_{l+1} images once): give every layer a share
  // of the SMs proportional to the BYTES it reads per tile, not to its flops.
  const int sms = device_sm_count();
  WgradPlan p;
  double total = 0;
  for (int l = 0; l < net.L; ++l) total += (double)((l == 0 ? 1 : net.nb) + net.nb);
  int used = 0;
  int64_t off = 0;
  for (int l = 0; l < net.L; ++l) {
    int g = (int)((double)sms * ((l == 0 ? 1 : net.nb) + net.nb) / total);
    if (g < 1) g = 1;
    p.item_begin[l] = used;
    p.part_off[l] = off;
    used += g;
    off += (int64_t)g * layer_K(net, l) * net.W;
  }
  p.item_begin[net.L] = used;
  p.part_off[net.L] = off;
  p.total_items = used;
  p.total_floats = off;
  return p;
}

}  // namespace mlp
}  // namespace loner

using namespace loner::mlp;

extern "C" int64_t loner_mlp_param_count(const loner_net_t* n) {
  Net net;
  if (!net_from(n, net)) return -1;
  return param_off(net, net.L) + 16 * (int64_t)net.W;
}
extern "C" int64_t loner_mlp_packed_bytes(const loner_net_t* n) {
  Net net;
  if (!net_from(n, net)) return -1;
  return packed_total(net);
}
static inline int64_t n_tiles(int64_t P) { return (P + kTile - 1) / kTile; }
extern "C" int64_t loner_mlp_act_bytes(const loner_net_t* n, int64_t P) {
  Net net;
  if (!net_from(n, net) || P < 0) return -1;
  return n_tiles(P) * (act_tile_bytes(net) + mask_tile_bytes(net));
}
extern "C" int64_t loner_mlp_bwd_scratch_bytes(const loner_net_t* n, int64_t P) {
  Net net;
  if (!net_from(n, net) || P < 0) return -1;
  const WgradPlan p = plan_wgrad(net);
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
  { const char* m = getenv("LONER_STASH"); a.bulk = (m && m[0] == 'c') ? 0 : 1; }
  a.masks = acts ? (uint8_t*)acts + a.tiles * act_tile_bytes(net) : nullptr;
  const int sms = device_sm_count();
  const int64_t pairs = (a.tiles + 1) / 2;
  const unsigned grid = (unsigned)(pairs < sms ? pairs : sms);
  auto launch = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPipeSmem);
    kern<<<grid, kPipeThreads, kPipeSmem, (cudaStream_t)stream>>>(a);
  };
  if (net.W == 256) { if (acts) launch(mlp_fwd_kernel<256, true>); else launch(mlp_fwd_kernel<256, false>); }
  else              { if (acts) launch(mlp_fwd_kernel<128, true>); else launch(mlp_fwd_kernel<128, false>); }
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
  const int sms = device_sm_count();
  BwdArgs b;
  b.net = net; b.packed = (const uint8_t*)packed; b.pos = pos; b.rays = rays; b.z = z_vals; b.S = S; b.P = P;
  b.s_shift = (S > 0 && (S & (S - 1)) == 0) ? __builtin_ctz((unsigned)S) : -1;
  b.tiles = tiles; b.d_sigma = d_sigma; b.masks = (const uint8_t*)acts + tiles * act_tile_bytes(net);
  b.dz = (uint8_t*)scratch; b.gscale = grad_scale; b.d_pos = d_pos;
  { const char* m = getenv("LONER_STASH"); b.bulk = (m && m[0] == 'c') ? 0 : 1; }
  const int64_t pairs = (tiles + 1) / 2;
  auto launch = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPipeSmem);
    kern<<<(unsigned)(pairs < sms ? pairs : sms), kPipeThreads, kPipeSmem, (cudaStream_t)stream>>>(b);
  };
  if (net.W == 256) launch(mlp_dgrad_kernel<256>); else launch(mlp_dgrad_kernel<128>);
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
  const int sms = device_sm_count();
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* dz = (uint8_t*)scratch;
  float* partials = (float*)(dz + tiles * dz_tile_bytes(net));
  const WgradPlan plan = plan_wgrad(net);
  WgradArgs w;
  w.net = net; w.acts = (const uint8_t*)acts; w.dz = dz; w.tiles = tiles; w.partials = partials;
  w.masks = (const uint8_t*)acts + tiles * act_tile_bytes(net); w.d_sigma = d_sigma;
  w.wout = reinterpret_cast<const float*>((const uint8_t*)packed + packed_wout_off(net)); w.gscale = grad_scale; w.P = P;
  for (int i = 0; i <= net.L; ++i) { w.item_begin[i] = plan.item_begin[i]; w.part_off[i] = plan.part_off[i]; }
  cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
  mlp_wgrad_kernel<<<(unsigned)plan.total_items, kWgThreads, kWgSmem, st>>>(w);
  LONER_CHECK_LAUNCH();
  dim3 rgrid(64, net.L);
  wgrad_reduce_kernel<<<rgrid, 256, 0, st>>>(net, partials, w, 1.0f / grad_scale, d_params);
  LONER_CHECK_LAUNCH();
  dwout_kernel<<<(unsigned)(sms * 4), 256, 0, st>>>(net, (const uint8_t*)acts, d_sigma, P, tiles, d_params);
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

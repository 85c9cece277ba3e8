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
// pack: fp32 master [out,in] row-major -> fp16 image rows = in (k), cols = out (n)
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
  uint8_t* img = packed + packed_off(net, l);
  const int chunks = K * (N / 8);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += gridDim.x * blockDim.x) {
    const int k = c % K, n0 = (c / K) * 8;
    __half2 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      h[i] = __floats2half2_rn(Wm[(int64_t)(n0 + 2 * i) * K + k], Wm[(int64_t)(n0 + 2 * i + 1) * K + k]);
    const int cb = n0 / 64, j = (n0 % 64) / 8;
    uint8_t* dst = img + (int64_t)cb * K * 128 + (int64_t)k * 128 + ((j ^ (k & 7)) * 16);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(h);
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
  int64_t P;
  int64_t tiles;
  float* sigma;          // [P]
  uint8_t* acts;         // activation stash or null
  uint8_t* masks;        // relu bit masks or null
};

// sample position in [0,1]^3 for global sample index gs (clamped by the caller)
__device__ __forceinline__ void sample_pos01(const float* pos, const float* rays, const float* z, int S, int64_t gs,
                                             float (&x)[3]) {
  float p[3];
  if (pos) {
    p[0] = pos[gs * 3 + 0]; p[1] = pos[gs * 3 + 1]; p[2] = pos[gs * 3 + 2];
  } else {
    const int64_t ray = gs / S;
    const float zz = z[gs];
    const float* R = rays + ray * LONER_RAY_COLS;
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = __fadd_rn(R[a], __fmul_rn(R[3 + a], zz));      // rendering_tcnn.py:241
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) x[a] = __fmul_rn(__fadd_rn(p[a], 1.0f), 0.5f);           // nerf_tcnn.py:63
}

// sin/cos(pi * 2^f * x): the scaling is exact, the reduction to [-1,1] is exact, sincospif does the rest
__device__ __forceinline__ void freq_pair(float x, int f, float& s, float& c) {
  const float t = ldexpf(x, f);
  const float r = t - 2.0f * rintf(0.5f * t);
  sincospif(r, &s, &c);
}

// Writes the 32 encoded features [32*half, 32*half+32) of row r into column block 0 of `sA`.
__device__ __forceinline__ void encode_row(uint8_t* sA, int r, int half, const float (&x)[3], const Net& net) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int cj = half * 4 + c;
    __half2 h[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int pair = cj * 4 + q;            // features 2*pair, 2*pair+1
      float s, co;
      if (pair < 3 * net.F) {
        const int dim = pair / net.F, f = pair - dim * net.F;
        const float xv = dim == 0 ? x[0] : (dim == 1 ? x[1] : x[2]);
        freq_pair(xv, f, s, co);
      } else if (2 * pair < net.Epad) {
        s = 1.0f; co = 1.0f;                  // tcnn pads the encoded width to 16 with 1.0
      } else {
        s = 0.0f; co = 0.0f;
      }
      h[q] = __floats2half2_rn(s, co);
    }
    *reinterpret_cast<uint4*>(sA + r * 128 + ((cj ^ (r & 7)) * 16)) = *reinterpret_cast<uint4*>(h);
  }
}

constexpr int kFwdThreads = 256;
// dynamic smem: 1 KB alignment slack | A (64 KB) | W (128 KB) | misc (4 KB)
constexpr int kFwdSmem = 1024 + 65536 + 131072 + 4096;

template <bool kStash>
__global__ void __launch_bounds__(kFwdThreads, 1) mlp_fwd_kernel(const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = base;
  uint8_t* sW = base + 65536;
  float* s_part = reinterpret_cast<float*>(base + 65536 + 131072);   // [2][128] sigma partials
  float* s_wout = s_part + 256;                                      // [W]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_wout + 256);        // [0]=weights landed, [1]=mma done
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2);
  const Net net = a.net;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_w = smem_u32(&bars[0]), bar_m = smem_u32(&bars[1]);

  if (warp == 0) tmem_alloc<256>(smem_u32(s_tmem));
  if (tid == 32) { mbar_init(bar_w, 1); mbar_init(bar_m, 1); fence_mbar_init(); }
  for (int j = tid; j < net.W; j += kFwdThreads)
    s_wout[j] = reinterpret_cast<const float*>(a.packed + packed_wout_off(net))[j];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  uint32_t ph_w = 0, ph_m = 0;

  const int q = warp & 3, h = warp >> 2;
  const int row = q * 32 + lane;                 // TMEM lane == sample row of the tile
  const int cols_per_thread = net.W / 2;

  for (int64_t tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
    // weights of layer 0 -> sW (previous tile's MMAs are complete: bar_m was waited on)
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)(net.Epad * net.W * 2);
      mbar_expect_tx(bar_w, bytes);
      // column blocks of layer 0 are Epad*128 bytes each, contiguous in the packed image
      bulk_g2s(smem_u32(sW), a.packed + packed_off(net, 0), bytes, bar_w);
      if (kStash) bulk_wait_read0();   // previous tile's stash stores have finished reading sA
    }
    __syncthreads();
    {  // encode A_0 (thread: row = tid & 127, half = tid >> 7)
      const int r = tid & 127, hf = tid >> 7;
      int64_t gs = tile * kTile + r;
      if (gs >= a.P) gs = a.P - 1;
      float x[3];
      sample_pos01(a.pos, a.rays, a.z, a.S, gs, x);
      encode_row(sA, r, hf, x, net);
    }
    fence_async_smem();
    __syncthreads();
    if (kStash && tid == 0) {
      bulk_s2g(a.acts + tile * act_tile_bytes(net), smem_u32(sA), kBlk);
      bulk_commit();
    }

    for (int l = 0; l < net.L; ++l) {
      const int K = layer_K(net, l);
      if (tid == 0) {
        mbar_wait(bar_w, ph_w);
        tc_fence_after();
        const uint32_t idesc = make_idesc_f16(128, net.W, 0, 1);
        for (int ks = 0; ks < K / 16; ++ks) {
          const uint64_t ad = make_desc_sw128(smem_u32(sA) + (ks >> 2) * kBlk + (ks & 3) * 32, 16, 1024);
          const uint64_t bd = make_desc_sw128(smem_u32(sW) + ks * 2048, (uint32_t)K * 128, 1024);
          umma_f16(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar_m);
      }
      ph_w ^= 1;
      mbar_wait(bar_m, ph_m);
      ph_m ^= 1;
      tc_fence_after();
      if (tid == 0) {
        if (l + 1 < net.L) {   // sW is free: fetch the next layer while the epilogue runs
          const uint32_t bytes = (uint32_t)(net.W * net.W * 2);
          mbar_expect_tx(bar_w, bytes);
          bulk_g2s(smem_u32(sW), a.packed + packed_off(net, l + 1), bytes, bar_w);
        }
        if (kStash) bulk_wait_read0();     // stash store of this layer's input image is done with sA
      }
      __syncthreads();

      // epilogue: TMEM -> relu -> fp16 -> sA (the next layer's A operand / the stash image)
      const bool last = (l == net.L - 1);
      float sig = 0.f;
      uint32_t mbits[4] = {0u, 0u, 0u, 0u};
      for (int it = 0; it < cols_per_thread / 32; ++it) {
        const int col0 = h * cols_per_thread + it * 32;
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
        tmem_ld_wait();
        uint32_t bits = 0u;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          __half2 hh[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float f0 = __uint_as_float(v[ch * 8 + 2 * e]), f1 = __uint_as_float(v[ch * 8 + 2 * e + 1]);
            bits |= (f0 > 0.f ? 1u : 0u) << (ch * 8 + 2 * e);
            bits |= (f1 > 0.f ? 1u : 0u) << (ch * 8 + 2 * e + 1);
            hh[e] = __floats2half2_rn(fmaxf(f0, 0.f), fmaxf(f1, 0.f));
            if (last) {
              const int c = col0 + ch * 8 + 2 * e;
              sig = fmaf(__low2float(hh[e]), s_wout[c], sig);
              sig = fmaf(__high2float(hh[e]), s_wout[c + 1], sig);
            }
          }
          if (!last || kStash) {
            const int c0 = col0 + ch * 8;
            const int cb = c0 >> 6, j = (c0 & 63) >> 3;
            *reinterpret_cast<uint4*>(sA + cb * kBlk + row * 128 + ((j ^ (row & 7)) * 16)) =
                *reinterpret_cast<uint4*>(hh);
          }
        }
        mbits[it] = bits;
      }
      if (kStash) {
        uint32_t* mrow = reinterpret_cast<uint32_t*>(a.masks + tile * mask_tile_bytes(net)) +
                         ((int64_t)l * kTile + row) * (net.W / 32) + h * (cols_per_thread / 32);
        for (int it = 0; it < cols_per_thread / 32; ++it) mrow[it] = mbits[it];
      }
      if (last) s_part[h * 128 + row] = sig;
      tc_fence_before();
      fence_async_smem();
      __syncthreads();
      if (kStash && tid == 0) {
        bulk_s2g(a.acts + tile * act_tile_bytes(net) + kBlk + (int64_t)l * net.nb * kBlk, smem_u32(sA),
                 (uint32_t)(net.nb * kBlk));
        bulk_commit();
      }
      if (last && tid < 128) {
        const int64_t gs = tile * kTile + tid;
        if (gs < a.P) a.sigma[gs] = s_part[tid] + s_part[128 + tid];
      }
    }
  }
  if (kStash && tid == 0) bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
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
  int64_t P;
  int64_t tiles;
  const float* d_sigma;
  const uint8_t* masks;
  uint8_t* dz;           // dZ stash [tiles][L][nb*16 KB]
  float gscale;
  float* d_pos;          // [P,3] or null
};

constexpr int kBwdThreads = 256;
constexpr int kBwdSmem = 1024 + 65536 + 131072 + 8192;

__global__ void __launch_bounds__(kBwdThreads, 1) mlp_dgrad_kernel(const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sG = base;
  uint8_t* sW = base + 65536;
  float* s_wout = reinterpret_cast<float*>(base + 65536 + 131072);   // [W]
  float* s_dx = s_wout + 256;                                        // [2][128][3] input-grad partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_dx + 768);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2);
  const Net net = a.net;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_w = smem_u32(&bars[0]), bar_m = smem_u32(&bars[1]);

  if (warp == 0) tmem_alloc<256>(smem_u32(s_tmem));
  if (tid == 32) { mbar_init(bar_w, 1); mbar_init(bar_m, 1); fence_mbar_init(); }
  for (int j = tid; j < net.W; j += kBwdThreads)
    s_wout[j] = reinterpret_cast<const float*>(a.packed + packed_wout_off(net))[j];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  uint32_t ph_w = 0, ph_m = 0;
  const int q = warp & 3, h = warp >> 2;
  const int row = q * 32 + lane;
  const int cpt = net.W / 2;          // columns per thread for W-wide outputs
  const bool want_dx = a.d_pos != nullptr;

  for (int64_t tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
    const int64_t gs = tile * kTile + row;
    const bool in = gs < a.P;
    if (tid == 0) {
      if (net.L > 1 || want_dx) {     // first GEMM's weights: layer L-1 (or layer 0 when L == 1)
        const int lw = net.L - 1;
        const uint32_t bytes = (uint32_t)(layer_K(net, lw) * net.W * 2);
        mbar_expect_tx(bar_w, bytes);
        bulk_g2s(smem_u32(sW), a.packed + packed_off(net, lw), bytes, bar_w);
      }
      bulk_wait_read0();              // previous tile's dZ stores are done with sG
    }
    __syncthreads();
    {  // dZ_L = d_sigma * w_out * relu'(Z_L)
      const float ds = in ? a.d_sigma[gs] * a.gscale : 0.f;
      const uint32_t* mrow = reinterpret_cast<const uint32_t*>(a.masks + tile * mask_tile_bytes(net)) +
                             ((int64_t)(net.L - 1) * kTile + row) * (net.W / 32) + h * (cpt / 32);
      for (int it = 0; it < cpt / 32; ++it) {
        const uint32_t bits = mrow[it];
        const int col0 = h * cpt + it * 32;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          __half2 hh[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = col0 + ch * 8 + 2 * e;
            float g0 = ((bits >> (ch * 8 + 2 * e)) & 1u) ? ds * s_wout[c] : 0.f;
            float g1 = ((bits >> (ch * 8 + 2 * e + 1)) & 1u) ? ds * s_wout[c + 1] : 0.f;
            g0 = fminf(fmaxf(g0, -65504.f), 65504.f);
            g1 = fminf(fmaxf(g1, -65504.f), 65504.f);
            hh[e] = __floats2half2_rn(g0, g1);
          }
          const int c0 = col0 + ch * 8;
          const int cb = c0 >> 6, j = (c0 & 63) >> 3;
          *reinterpret_cast<uint4*>(sG + cb * kBlk + row * 128 + ((j ^ (row & 7)) * 16)) =
              *reinterpret_cast<uint4*>(hh);
        }
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(a.dz + tile * dz_tile_bytes(net) + (int64_t)(net.L - 1) * net.nb * kBlk, smem_u32(sG),
               (uint32_t)(net.nb * kBlk));
      bulk_commit();
    }

    // dA_l = dZ_{l+1} * W_l  for l = L-1 .. 1 (and l = 0 when input gradients are wanted)
    for (int l = net.L - 1; l >= (want_dx ? 0 : 1); --l) {
      const int Nout = layer_K(net, l);     // width of dA_l (in-features of layer l)
      if (tid == 0) {
        mbar_wait(bar_w, ph_w);
        tc_fence_after();
        const uint32_t idesc = make_idesc_f16(128, Nout, 0, 0);
        for (int ks = 0; ks < net.W / 16; ++ks) {     // contraction over the out-features of layer l
          const uint64_t ad = make_desc_sw128(smem_u32(sG) + (ks >> 2) * kBlk + (ks & 3) * 32, 16, 1024);
          const uint64_t bd =
              make_desc_sw128(smem_u32(sW) + (ks >> 2) * (Nout * 128) + (ks & 3) * 32, 16, 1024);
          umma_f16(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar_m);
      }
      ph_w ^= 1;
      mbar_wait(bar_m, ph_m);
      ph_m ^= 1;
      tc_fence_after();
      if (tid == 0) {
        const int ln = l - 1;
        if (ln >= (want_dx ? 0 : 1)) {
          const uint32_t bytes = (uint32_t)(layer_K(net, ln) * net.W * 2);
          mbar_expect_tx(bar_w, bytes);
          bulk_g2s(smem_u32(sW), a.packed + packed_off(net, ln), bytes, bar_w);
        }
        bulk_wait_read0();
      }
      __syncthreads();

      if (l >= 1) {
        // epilogue: dZ_l = dA_l * relu'(Z_l)  -> fp16 image in sG
        const uint32_t* mrow = reinterpret_cast<const uint32_t*>(a.masks + tile * mask_tile_bytes(net)) +
                               ((int64_t)(l - 1) * kTile + row) * (net.W / 32) + h * (cpt / 32);
        for (int it = 0; it < cpt / 32; ++it) {
          const int col0 = h * cpt + it * 32;
          uint32_t v[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
          tmem_ld_wait();
          const uint32_t bits = mrow[it];
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            __half2 hh[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float g0 = ((bits >> (ch * 8 + 2 * e)) & 1u) ? __uint_as_float(v[ch * 8 + 2 * e]) : 0.f;
              float g1 = ((bits >> (ch * 8 + 2 * e + 1)) & 1u) ? __uint_as_float(v[ch * 8 + 2 * e + 1]) : 0.f;
              g0 = fminf(fmaxf(g0, -65504.f), 65504.f);
              g1 = fminf(fmaxf(g1, -65504.f), 65504.f);
              hh[e] = __floats2half2_rn(g0, g1);
            }
            const int c0 = col0 + ch * 8;
            const int cb = c0 >> 6, j = (c0 & 63) >> 3;
            *reinterpret_cast<uint4*>(sG + cb * kBlk + row * 128 + ((j ^ (row & 7)) * 16)) =
                *reinterpret_cast<uint4*>(hh);
          }
        }
        tc_fence_before();
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
          bulk_s2g(a.dz + tile * dz_tile_bytes(net) + (int64_t)(l - 1) * net.nb * kBlk, smem_u32(sG),
                   (uint32_t)(net.nb * kBlk));
          bulk_commit();
        }
      } else {
        // l == 0: dEnc [128 x Epad] -> d_pos through the sin/cos encoding
        float x[3];
        {
          int64_t g2 = in ? gs : a.P - 1;
          sample_pos01(a.pos, a.rays, a.z, a.S, g2, x);
        }
        float dx[3] = {0.f, 0.f, 0.f};
        if (h * 32 < net.Epad) {
          uint32_t v[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int pq = 0; pq < 16; ++pq) {
            const int pair = h * 16 + pq;
            if (pair < 3 * net.F) {
              const int dim = pair / net.F, f = pair - dim * net.F;
              const float xv = dim == 0 ? x[0] : (dim == 1 ? x[1] : x[2]);
              float s, c;
              freq_pair(xv, f, s, c);
              const float k = ldexpf(3.14159265358979323846f, f);
              const float g = (__uint_as_float(v[2 * pq]) * c - __uint_as_float(v[2 * pq + 1]) * s) * k;
              if (dim == 0) dx[0] += g; else if (dim == 1) dx[1] += g; else dx[2] += g;
            }
          }
        }
        s_dx[(h * 128 + row) * 3 + 0] = dx[0];
        s_dx[(h * 128 + row) * 3 + 1] = dx[1];
        s_dx[(h * 128 + row) * 3 + 2] = dx[2];
        tc_fence_before();
        __syncthreads();
        if (tid < 128) {
          const int64_t g3 = tile * kTile + tid;
          if (g3 < a.P) {
            const float inv = 0.5f / a.gscale;      // x = (pos + 1) / 2
#pragma unroll
            for (int d = 0; d < 3; ++d)
              a.d_pos[g3 * 3 + d] = (s_dx[tid * 3 + d] + s_dx[(128 + tid) * 3 + d]) * inv;
          }
        }
        __syncthreads();
      }
    }
  }
  if (tid == 0) bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
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
  int item_begin[9];         // first item of each layer (prefix), item_begin[L] = total
  int64_t part_off[9];       // float offset of each layer's first partial
};

constexpr int kWgStages = 3;
constexpr int kWgStageBytes = 65536;    // 64 samples: A half (<= 32 KB) | dZ half (<= 32 KB)
constexpr int kWgThreads = 192;
constexpr int kWgSmem = 1024 + kWgStages * kWgStageBytes + 1024;

__global__ void __launch_bounds__(kWgThreads, 1) mlp_wgrad_kernel(const WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kWgStages * kWgStageBytes);
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

  if (warp == 0) tmem_alloc<512>(smem_u32(s_tmem));
  if (tid == 32) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
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
      mbar_expect_tx(full_bar(s), bytesA + bytesY);
      const uint8_t* srcA = a.acts + tile * act_tile_bytes(net) + actA_off + hf * 8192;
      const uint8_t* srcY = a.dz + tile * dz_tile_bytes(net) + dzY_off + hf * 8192;
      for (int cb = 0; cb < nbA; ++cb) bulk_g2s(dstA + cb * 8192, srcA + (int64_t)cb * kBlk, 8192, full_bar(s));
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
  if (warp >= 2) {
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

// dW_out[j] = sum_s d_sigma[s] * A_L[s,j]   (row 0 of the padded [16, W] output matrix)
__global__ void __launch_bounds__(256) dwout_kernel(Net net, const uint8_t* __restrict__ acts,
                                                    const float* __restrict__ d_sigma, int64_t P, int64_t tiles,
                                                    float* __restrict__ d_params) {
  __shared__ float red[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * 8 + warp, nw = (int64_t)gridDim.x * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};    // lane owns logical columns cb*64 + 2*lane, +1
  const int64_t aL = (int64_t)kBlk + (int64_t)(net.L - 1) * net.nb * kBlk;
  for (int64_t tile = gw; tile < tiles; tile += nw) {
    const uint8_t* img = acts + tile * act_tile_bytes(net) + aL;
    for (int r = 0; r < kTile; ++r) {
      const int64_t gs = tile * kTile + r;
      if (gs >= P) break;
      const float ds = __ldg(d_sigma + gs);
      const int slot = ((lane >> 2) ^ (r & 7)) * 16 + (lane & 3) * 4;
      for (int cb = 0; cb < net.nb; ++cb) {
        const __half2 v = *reinterpret_cast<const __half2*>(img + cb * kBlk + r * 128 + slot);
        acc[2 * cb] = fmaf(ds, __low2float(v), acc[2 * cb]);
        acc[2 * cb + 1] = fmaf(ds, __high2float(v), acc[2 * cb + 1]);
      }
    }
  }
  for (int j = threadIdx.x; j < 256; j += blockDim.x) red[j] = 0.f;
  __syncthreads();
  for (int cb = 0; cb < net.nb; ++cb) {
    atomicAdd(&red[cb * 64 + 2 * lane], acc[2 * cb]);
    atomicAdd(&red[cb * 64 + 2 * lane + 1], acc[2 * cb + 1]);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < net.W; j += blockDim.x)
    if (red[j] != 0.f) atomicAdd(d_params + param_off(net, net.L) + j, red[j]);
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
  return packed_wout_off(net) + (int64_t)net.W * 4;
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
  if (!packed || !sigma || P < 0 || (!pos && (!rays || !z_vals || S <= 0))) return LONER_E_BAD_ARG;
  if (P == 0) return LONER_OK;
  FwdArgs a;
  a.net = net; a.packed = (const uint8_t*)packed; a.pos = pos; a.rays = rays; a.z = z_vals; a.S = S; a.P = P;
  a.tiles = n_tiles(P); a.sigma = sigma; a.acts = (uint8_t*)acts;
  a.masks = acts ? (uint8_t*)acts + a.tiles * act_tile_bytes(net) : nullptr;
  const int sms = device_sm_count();
  const unsigned grid = (unsigned)(a.tiles < sms ? a.tiles : sms);
  if (acts) {
    cudaFuncSetAttribute(mlp_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
    mlp_fwd_kernel<true><<<grid, kFwdThreads, kFwdSmem, (cudaStream_t)stream>>>(a);
  } else {
    cudaFuncSetAttribute(mlp_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
    mlp_fwd_kernel<false><<<grid, kFwdThreads, kFwdSmem, (cudaStream_t)stream>>>(a);
  }
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
  if (!d_sigma || !(grad_scale > 0.f) || (d_pos && !pos && (!rays || !z_vals || S <= 0))) return LONER_E_BAD_ARG;
  if (P == 0) return LONER_OK;
  const int64_t tiles = n_tiles(P);
  const int sms = device_sm_count();
  BwdArgs b;
  b.net = net; b.packed = (const uint8_t*)packed; b.pos = pos; b.rays = rays; b.z = z_vals; b.S = S; b.P = P;
  b.tiles = tiles; b.d_sigma = d_sigma; b.masks = (const uint8_t*)acts + tiles * act_tile_bytes(net);
  b.dz = (uint8_t*)scratch; b.gscale = grad_scale; b.d_pos = d_pos;
  cudaFuncSetAttribute(mlp_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
  mlp_dgrad_kernel<<<(unsigned)(tiles < sms ? tiles : sms), kBwdThreads, kBwdSmem, (cudaStream_t)stream>>>(b);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_mlp_wgrad(const loner_net_t* n, const void* packed, int64_t P, const float* d_sigma,
                               const void* acts, float grad_scale, float* d_params, void* scratch, void* stream) {
  Net net;
  int rc = bwd_common(n, net, packed, acts, scratch, P);
  if (rc) return rc;
  if (!d_sigma || !d_params || !(grad_scale > 0.f)) return LONER_E_BAD_ARG;
  if (P == 0) return LONER_OK;
  const int64_t tiles = n_tiles(P);
  const int sms = device_sm_count();
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* dz = (uint8_t*)scratch;
  float* partials = (float*)(dz + tiles * dz_tile_bytes(net));
  const WgradPlan plan = plan_wgrad(net);
  WgradArgs w;
  w.net = net; w.acts = (const uint8_t*)acts; w.dz = dz; w.tiles = tiles; w.partials = partials;
  for (int i = 0; i <= net.L; ++i) { w.item_begin[i] = plan.item_begin[i]; w.part_off[i] = plan.part_off[i]; }
  cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
  mlp_wgrad_kernel<<<(unsigned)plan.total_items, kWgThreads, kWgSmem, st>>>(w);
  LONER_CHECK_LAUNCH();
  dim3 rgrid(64, net.L);
  wgrad_reduce_kernel<<<rgrid, 256, 0, st>>>(net, partials, w, 1.0f / grad_scale, d_params);
  LONER_CHECK_LAUNCH();
  dwout_kernel<<<(unsigned)(sms * 2), 256, 0, st>>>(net, (const uint8_t*)acts, d_sigma, P, tiles, d_params);
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

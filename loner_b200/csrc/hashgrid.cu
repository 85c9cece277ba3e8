// Sigma head of the reference's SHIPPED configuration: multiresolution hash encoding + one hidden
// layer of 64 neurons (/root/reference/cfg/nerf_config/default_nerf_hash.yaml `pos_encoding_sigma`,
// `sigma_network`; built at /root/reference/src/models/nerf_tcnn.py:35-38, evaluated at :59-78).
// SURVEY.md 8f rank 1.  Semantics: oracle/hashgrid_standin.py (tiny-cuda-nn GridEncoding, grid type
// Hash, linear interpolation) + oracle/tcnn_standin.py (bias-free ReLU MLP, fp16 weights and
// activations, fp32 accumulation, fp32 output).
//
// This path is gather-bound, not tensor-bound: 16 levels x 8 corners x 4 B per sample out of a 14 MB fp16
// table that lives in L2, against 4.2 KFLOP of MLP.  Design:
//  * one sample per thread, the 32 encoded features and the 64 hidden activations stay in registers,
//    weights are broadcast from shared memory; nothing but sigma is written by the forward;
//  * the backward RECOMPUTES the forward (re-gathering from L2 is cheaper than a 200 B/sample stash),
//    scatters table gradients with 8-byte vector atomics (red.global.add.v2.f32) into an fp32 gradient
//    table, and accumulates dW1 / dW_out per CTA from a shared-memory stash of (dh, enc, h) tiles -
//    persistent CTAs, per-CTA partial sums, deterministic reduction, no atomics on the weights.
#include <cmath>
#include <cstdlib>
#include "common.cuh"

namespace loner {
namespace hashgrid {

constexpr int kMaxLevels = 16;
constexpr int kW = 64;             // hidden width
constexpr int kThreads = 256;
constexpr int kMaxHidden = 4;      // hidden layers supported (tcnn FullyFusedMLP width 64)

struct HashNet {
  int flags;                       // LONER_HASH_* (kernel variants)
  int n_levels, E, Epad;           // E = 2 * n_levels encoded features, padded to 16 with 1.0
  int L;                           // hidden layers of 64 neurons (1 = the shipped head; 2-4 run on the "deep" kernels)
  float scale[kMaxLevels];
  uint32_t res[kMaxLevels];
  uint32_t entries[kMaxLevels];    // "hashmap size" of the level
  uint32_t offset[kMaxLevels + 1]; // first entry of the level in the table
  uint32_t dense[kMaxLevels];      // 1: index = x + y res + z res^2, 0: spatial hash
};

__host__ inline bool net_from(const loner_hashnet_t* n, HashNet& o) {
  if (!n) return false;
  if (n->n_levels < 1 || n->n_levels > kMaxLevels || n->n_features_per_level != 2) return false;
  if (n->log2_hashmap_size < 4 || n->log2_hashmap_size > 24 || n->base_resolution < 1) return false;
  if (n->n_neurons != kW || n->n_hidden_layers < 1 || n->n_hidden_layers > kMaxHidden) return false;
  if (n->n_hidden_layers > 1 && (n->flags & LONER_HASH_SCALAR)) return false;     // the scalar A/B kernels are 1 x 64 only
  o.L = n->n_hidden_layers;
  const float pls = n->per_level_scale > 0.f ? n->per_level_scale : 2.0f;
  o.n_levels = n->n_levels;
  o.flags = n->flags;
  o.E = 2 * n->n_levels;
  o.Epad = (o.E + 15) / 16 * 16;
  uint32_t off = 0;
  const float log2_pls = log2f(pls);
  for (int l = 0; l < o.n_levels; ++l) {
    const float scale = exp2f((float)l * log2_pls) * (float)n->base_resolution - 1.0f;   // grid_scale()
    const uint32_t res = (uint32_t)ceilf(scale) + 1u;                                     // grid_resolution()
    const double cube = (double)res * res * res;
    uint64_t cnt = cube > 2147483647.0 ? 2147483647ull : (uint64_t)cube;
    cnt = (cnt + 7) / 8 * 8;
    const uint64_t cap = 1ull << n->log2_hashmap_size;
    if (cnt > cap) cnt = cap;
    // grid_index(): walk the strides while stride <= hashmap size; hashed iff the walk was cut short
    uint64_t stride = 1;
    for (int d = 0; d < 3 && stride <= cnt; ++d) stride *= res;
    o.scale[l] = scale; o.res[l] = res; o.entries[l] = (uint32_t)cnt; o.offset[l] = off;
    o.dense[l] = (cnt < stride) ? 0u : 1u;
    off += (uint32_t)cnt;
  }
  for (int l = o.n_levels; l <= kMaxLevels; ++l) o.offset[l] = off;
  for (int l = o.n_levels; l < kMaxLevels; ++l) { o.scale[l] = 0.f; o.res[l] = 1; o.entries[l] = 1; o.dense[l] = 1; }
  return true;
}
__host__ __device__ inline int64_t n_entries(const HashNet& n) { return n.offset[kMaxLevels]; }
__host__ __device__ inline int64_t w1_floats(const HashNet& n) { return (int64_t)kW * n.Epad; }
__host__ __device__ inline int64_t wh_floats(const HashNet& n) { return (int64_t)(n.L - 1) * kW * kW; }   // hidden matrices 2..L
__host__ __device__ inline int64_t net_floats(const HashNet& n) { return w1_floats(n) + wh_floats(n) + 16 * kW; }   // + padded [16, W] output matrix
// packed image: W1 fp16 [64][Epad] | W_2..W_L fp16 [64][64] | w_out fp32 [64] (values of the fp16-rounded row 0) | table half2 [entries]
__host__ __device__ inline int64_t packed_wout_off(const HashNet& n) { return (w1_floats(n) + wh_floats(n)) * 2; }
__host__ __device__ inline int64_t packed_table_off(const HashNet& n) { return packed_wout_off(n) + kW * 4; }
__host__ __device__ inline int64_t packed_total(const HashNet& n) { return packed_table_off(n) + n_entries(n) * 4; }

__global__ void __launch_bounds__(256) pack_kernel(HashNet net, const float* __restrict__ params, uint8_t* __restrict__ packed) {
  const int64_t nw1 = w1_floats(net) + wh_floats(net), nt = n_entries(net);      // all fp16 matrices, contiguous
  __half* w1 = reinterpret_cast<__half*>(packed);
  float* wo = reinterpret_cast<float*>(packed + packed_wout_off(net));
  __half2* tb = reinterpret_cast<__half2*>(packed + packed_table_off(net));
  const float* tsrc = params + net_floats(net);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw1 + kW + nt; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < nw1) w1[i] = __float2half_rn(params[i]);
    else if (i < nw1 + kW) wo[i - nw1] = __half2float(__float2half_rn(params[nw1 + (i - nw1)]));   // row 0 of [16, W]
    else { const int64_t e = i - nw1 - kW; tb[e] = __floats2half2_rn(tsrc[2 * e], tsrc[2 * e + 1]); }
  }
}

struct Args {
  HashNet net;
  const uint8_t* packed;
  const float* pos;      // [P,3] in [-1,1] or null
  const float* rays;     // [n,13]
  const float* z;        // [n,S]
  int S;
  int64_t P;
  float* sigma;          // forward output
  const float* d_sigma;  // backward input
  float gscale;          // loss scale of the fp16 (dh) stash
  float* d_table;        // [entries][2] fp32, += with vector atomics
  float* d_pos;          // [P,3] or null
  float* partials;       // [gridDim.x][64*Epad + 64]
};

// sample position in [0,1]^3:  (o + d z + 1) / 2     rendering_tcnn.py:241, nerf_tcnn.py:63
__device__ __forceinline__ void position01(const Args& a, int64_t s, float (&x)[3]) {
  if (a.pos) {
#pragma unroll
    for (int d = 0; d < 3; ++d) x[d] = __fmul_rn(__fadd_rn(__ldg(a.pos + s * 3 + d), 1.0f), 0.5f);
  } else {
    const int64_t ray = s / a.S;
    const float* R = a.rays + ray * LONER_RAY_COLS;
    const float z = __ldg(a.z + s);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float p = __fadd_rn(__ldg(R + d), __fmul_rn(__ldg(R + 3 + d), z));
      x[d] = __fmul_rn(__fadd_rn(p, 1.0f), 0.5f);
    }
  }
}

struct Cell {
  uint32_t c[3];
  float f[3];
};
__device__ __forceinline__ Cell locate(float scale, const float (&x)[3]) {
  Cell r;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float pos = fmaf(scale, x[d], 0.5f);
    const float fl = floorf(pos);
    r.c[d] = (uint32_t)(int)fl;
    r.f[d] = pos - fl;
  }
  return r;
}
__device__ __forceinline__ uint32_t entry_index(uint32_t px, uint32_t py, uint32_t pz, uint32_t res, uint32_t entries, bool dense) {
  uint32_t idx = dense ? px + py * res + pz * res * res : (px * 1u) ^ (py * 2654435761u) ^ (pz * 805459861u);
  return ((entries & (entries - 1u)) == 0u) ? (idx & (entries - 1u)) : (idx % entries);
}
__device__ __forceinline__ float corner_weight(const Cell& q, int c) {
  return ((c & 1) ? q.f[0] : 1.0f - q.f[0]) * ((c & 2) ? q.f[1] : 1.0f - q.f[1]) * ((c & 4) ? q.f[2] : 1.0f - q.f[2]);
}

// encoded features of one sample, as fp32 values of the fp16-rounded encoding (tcnn writes __half)
__device__ __forceinline__ void encode(const HashNet& net, const __half2* __restrict__ table, const float (&x)[3],
                                       float (&enc)[2 * kMaxLevels]) {
#pragma unroll
  for (int l = 0; l < kMaxLevels; ++l) {
    if (l < net.n_levels) {
      const Cell q = locate(net.scale[l], x);
      const __half2* t = table + net.offset[l];
      const uint32_t res = net.res[l], ent = net.entries[l];
      const bool dense = net.dense[l] != 0u;
      float ax = 0.f, ay = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t idx = entry_index(q.c[0] + (c & 1), q.c[1] + ((c >> 1) & 1), q.c[2] + (c >> 2), res, ent, dense);
        const float2 v = __half22float2(__ldg(t + idx));
        const float w = corner_weight(q, c);
        ax = fmaf(w, v.x, ax);
        ay = fmaf(w, v.y, ay);
      }
      const float2 r = __half22float2(__floats2half2_rn(ax, ay));
      enc[2 * l] = r.x; enc[2 * l + 1] = r.y;
    } else {
      const float pad = (2 * l < net.Epad) ? 1.0f : 0.0f;     // encoded width padded to 16 with ones
      enc[2 * l] = pad; enc[2 * l + 1] = pad;
    }
  }
}

// hidden pre-activation j:  sum_i W1[j][i] enc[i]   (fp16 weights broadcast from shared memory, fp32 accumulate)
__device__ __forceinline__ float hidden_dot(const __half2* __restrict__ sW1row, const float (&enc)[2 * kMaxLevels], int pairs) {
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxLevels; ++i) {
    if (i < pairs) {
      const float2 w = __half22float2(sW1row[i]);
      a0 = fmaf(w.x, enc[2 * i], a0);
      a1 = fmaf(w.y, enc[2 * i + 1], a1);
    }
  }
  return a0 + a1;
}
__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }

__device__ __forceinline__ void load_weights(const Args& a, __half2 (*sW1)[kMaxLevels], float* sWout, int tid) {
  const __half2* w1 = reinterpret_cast<const __half2*>(a.packed);
  const int pairs = a.net.Epad / 2;
  for (int i = tid; i < kW * kMaxLevels; i += kThreads) {
    const int j = i / kMaxLevels, k = i % kMaxLevels;
    sW1[j][k] = k < pairs ? w1[j * pairs + k] : __floats2half2_rn(0.f, 0.f);
  }
  const float* wo = reinterpret_cast<const float*>(a.packed + packed_wout_off(a.net));
  for (int j = tid; j < kW; j += kThreads) sWout[j] = wo[j];
}

__global__ void __launch_bounds__(kThreads) hash_fwd_kernel(const Args a) {
  __shared__ __half2 sW1[kW][kMaxLevels];
  __shared__ float sWout[kW];
  load_weights(a, sW1, sWout, threadIdx.x);
  __syncthreads();
  const __half2* table = reinterpret_cast<const __half2*>(a.packed + packed_table_off(a.net));
  const int pairs = a.net.Epad / 2;
  for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < a.P; s += (int64_t)gridDim.x * kThreads) {
    float x[3], enc[2 * kMaxLevels];
    position01(a, s, x);
    encode(a.net, table, x, enc);
    float sig0 = 0.f, sig1 = 0.f;
#pragma unroll 4
    for (int j = 0; j < kW; j += 2) {
      const float h0 = round_h(fmaxf(hidden_dot(sW1[j], enc, pairs), 0.f));
      const float h1 = round_h(fmaxf(hidden_dot(sW1[j + 1], enc, pairs), 0.f));
      sig0 = fmaf(h0, sWout[j], sig0);
      sig1 = fmaf(h1, sWout[j + 1], sig1);
    }
    a.sigma[s] = sig0 + sig1;
  }
}

// ------------------------------------------------------------------------------------------
// backward: recompute, d_sigma -> table gradients (vector atomics), per-CTA dW1 / dW_out partials, d_pos
// Stash rows are written by their owning thread with 16-byte stores; the row strides (144 B and 80 B)
// put the eight lanes of a quarter warp on disjoint bank groups.
constexpr int kStashH = kW + 8;            // halves per row of the dh and h stashes (144 B)
constexpr int kStashE = 2 * kMaxLevels + 8;   // halves per row of the encoding stash (80 B)
struct BwdSmem {
  __half2 sW1[kW][kMaxLevels];             // 4 KB
  float sWout[kW];
  float s_ds[kThreads];
  __align__(16) __half s_dh[kThreads][kStashH];    // 36 KB  loss-scaled dh, fp16
  __align__(16) __half s_h[kThreads][kStashH];     // 36 KB  hidden activations (already fp16 values)
  __align__(16) __half s_enc[kThreads][kStashE];   // 20 KB  encoded inputs (already fp16 values)
};
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <bool kDx>
__global__ void __launch_bounds__(kThreads) hash_bwd_kernel(const Args a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int tid = threadIdx.x;
  load_weights(a, sm.sW1, sm.sWout, tid);
  __syncthreads();
  const HashNet& net = a.net;
  const __half2* table = reinterpret_cast<const __half2*>(a.packed + packed_table_off(net));
  const int pairs = net.Epad / 2;
  // this thread's slice of dW1: row j = tid / 4, columns [i0, i0 + per)
  const int per = net.Epad / 4;            // 4 or 8
  const int gj = tid >> 2, i0 = (tid & 3) * per;
  float accW[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) accW[k] = 0.f;
  float accWout = 0.f;                     // threads 0..63: dW_out[tid]
  const int64_t n_tiles = (a.P + kThreads - 1) / kThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t s = tile * kThreads + tid;
    const bool in = s < a.P;
    float x[3] = {0.f, 0.f, 0.f}, enc[2 * kMaxLevels];
    if (in) position01(a, s, x);
    encode(net, table, x, enc);
    const float ds = in ? __ldg(a.d_sigma + s) : 0.f;
    // forward through the hidden layer, backward into d_enc
    float d_enc[2 * kMaxLevels];
#pragma unroll
    for (int i = 0; i < 2 * kMaxLevels; ++i) d_enc[i] = 0.f;
    for (int j0 = 0; j0 < kW; j0 += 8) {
      float hv[8], dv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u;
        const float hj = round_h(fmaxf(hidden_dot(sm.sW1[j], enc, pairs), 0.f));
        const float dh = hj > 0.f ? ds * sm.sWout[j] : 0.f;
        hv[u] = hj;
        dv[u] = fminf(fmaxf(dh * a.gscale, -65504.f), 65504.f);
#pragma unroll
        for (int i = 0; i < kMaxLevels; ++i) {
          if (i < pairs) {
            const float2 w = __half22float2(sm.sW1[j][i]);
            d_enc[2 * i] = fmaf(dh, w.x, d_enc[2 * i]);
            d_enc[2 * i + 1] = fmaf(dh, w.y, d_enc[2 * i + 1]);
          }
        }
      }
      *reinterpret_cast<uint4*>(&sm.s_h[tid][j0]) =
          make_uint4(pack_h2(hv[0], hv[1]), pack_h2(hv[2], hv[3]), pack_h2(hv[4], hv[5]), pack_h2(hv[6], hv[7]));
      *reinterpret_cast<uint4*>(&sm.s_dh[tid][j0]) =
          make_uint4(pack_h2(dv[0], dv[1]), pack_h2(dv[2], dv[3]), pack_h2(dv[4], dv[5]), pack_h2(dv[6], dv[7]));
    }
    sm.s_ds[tid] = ds;
#pragma unroll
    for (int i = 0; i < 2 * kMaxLevels; i += 8)
      *reinterpret_cast<uint4*>(&sm.s_enc[tid][i]) =
          make_uint4(pack_h2(enc[i], enc[i + 1]), pack_h2(enc[i + 2], enc[i + 3]), pack_h2(enc[i + 4], enc[i + 5]),
                     pack_h2(enc[i + 6], enc[i + 7]));
    // scatter into the gradient table, and d_pos through the interpolation weights
    float dx[3] = {0.f, 0.f, 0.f};
    if (in && ds != 0.f) {
#pragma unroll
      for (int l = 0; l < kMaxLevels; ++l) {
        if (l < net.n_levels) {
          const Cell q = locate(net.scale[l], x);
          const uint32_t res = net.res[l], ent = net.entries[l];
          const bool dense = net.dense[l] != 0u;
          float2* gt = reinterpret_cast<float2*>(a.d_table) + net.offset[l];
          const float gx = d_enc[2 * l], gy = d_enc[2 * l + 1];
          float lx[3] = {0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t idx = entry_index(q.c[0] + (c & 1), q.c[1] + ((c >> 1) & 1), q.c[2] + (c >> 2), res, ent, dense);
            const float w = corner_weight(q, c);
            atomicAdd(gt + idx, make_float2(w * gx, w * gy));
            if (kDx) {
              const float2 v = __half22float2(__ldg(table + net.offset[l] + idx));
              const float dot = v.x * gx + v.y * gy;
              const float w0 = (c & 1) ? q.f[0] : 1.0f - q.f[0], w1 = (c & 2) ? q.f[1] : 1.0f - q.f[1],
                          w2 = (c & 4) ? q.f[2] : 1.0f - q.f[2];
              lx[0] += ((c & 1) ? 1.0f : -1.0f) * w1 * w2 * dot;
              lx[1] += ((c & 2) ? 1.0f : -1.0f) * w0 * w2 * dot;
              lx[2] += ((c & 4) ? 1.0f : -1.0f) * w0 * w1 * dot;
            }
          }
          if (kDx) {
#pragma unroll
            for (int d = 0; d < 3; ++d) dx[d] = fmaf(net.scale[l], lx[d], dx[d]);
          }
        }
      }
    }
    if (kDx && in) {
#pragma unroll
      for (int d = 0; d < 3; ++d) a.d_pos[s * 3 + d] = 0.5f * dx[d];     // x = (pos + 1) / 2
    }
    __syncthreads();
    // dW1[j][i] += sum_s dh[s][j] enc[s][i];  dW_out[j] += sum_s d_sigma[s] h[s][j]
    for (int t = 0; t < kThreads; ++t) {
      const float dh = __half2float(sm.s_dh[t][gj]);
      if (per == 8) {
        const uint4 e4 = *reinterpret_cast<const uint4*>(&sm.s_enc[t][i0]);
        const uint32_t ew[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&ew[k]));
          accW[2 * k] = fmaf(dh, e.x, accW[2 * k]);
          accW[2 * k + 1] = fmaf(dh, e.y, accW[2 * k + 1]);
        }
      } else {
        const uint2 e2 = *reinterpret_cast<const uint2*>(&sm.s_enc[t][i0]);
        const uint32_t ew[2] = {e2.x, e2.y};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&ew[k]));
          accW[2 * k] = fmaf(dh, e.x, accW[2 * k]);
          accW[2 * k + 1] = fmaf(dh, e.y, accW[2 * k + 1]);
        }
      }
      if (tid < kW) accWout = fmaf(sm.s_ds[t], __half2float(sm.s_h[t][tid]), accWout);
    }
    __syncthreads();
  }
  float* part = a.partials + (int64_t)blockIdx.x * (w1_floats(net) + kW);
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < per) part[gj * net.Epad + i0 + k] = accW[k];
  if (tid < kW) part[w1_floats(net) + tid] = accWout;
}

// ------------------------------------------------------------------------------------------
// Warp-level tensor-core variant (round 2).  ncu of the scalar kernels above: hash_bwd executed 26.7 K
// instructions per sample - 3.5 G warp instructions per C2 step, 46 % issue-slot utilisation, 31 % of the
// stall samples "no instruction" (instruction-cache misses of the fully unrolled body) - i.e. the kernel was
// bound by the SCALAR code of the 32 -> 64 -> 1 head (an LDS + two converts per two FMAs), not by the
// gathers / atomics it exists for.  Here the three small GEMMs of a 32-sample warp tile
//     H [32 x 64]  = Enc [32 x 32] W1^T          (forward, recomputed in the backward)
//     dEnc [32 x 32] = dH [32 x 64] W1           (input gradient of the head)
//     dW1 [64 x 32] += dH^T [64 x 256] Enc [256 x 32]   (per CTA tile of 256 samples, M x N split over the warps)
// run on mma.sync.m16n8k16 (fp16 operands, fp32 accumulation), fed by ldmatrix from small per-warp tiles in
// shared memory.  tcgen05 is the wrong tool for a 12 KFLOP-per-sample contraction inside a gather kernel:
// it would need the operands staged as 128-row tiles and TMEM round trips for 64 columns of output.
constexpr int kEncLd = 40;          // halves per row of the per-warp encoding tile (32 + 8: conflict-free ldmatrix)
constexpr int kDhLd = 72;           // halves per row of the per-warp dH tile (64 + 8)
constexpr int kDencLd = 33;         // floats per row of the per-warp dEnc tile
constexpr int kW1Ld = 40;           // halves per row of W1 [64][32 + 8]
constexpr int kAggLevels = 4;       // coarse levels whose table reductions are aggregated inside the warp (sweep on a B200, backward ms
                                    // at the C2 size / default operating point: 2 levels 3.65 / 2.27, 4: 3.37 / 2.04, 5: 3.36 / 2.04, 6: 3.40 / 2.04, 8: 3.31 / 2.04)
struct WarpTiles {
  __align__(16) __half enc[32][kEncLd];      // 2560 B
  __align__(16) __half dh[32][kDhLd];        // 4608 B
  float denc[32][kDencLd];                   // 4224 B
};
struct MmaSmem {
  __align__(16) __half w1[kW][kW1Ld];        // 5120 B, zero beyond Epad
  float wout[kW];
  float red[8][kW];                          // dW_out partials of the warps
  WarpTiles wt[8];
};

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

__device__ __forceinline__ void load_weights_mma(const Args& a, MmaSmem& sm, int tid) {
  const __half* w1 = reinterpret_cast<const __half*>(a.packed);
  const int E = a.net.Epad;
  for (int i = tid; i < kW * kW1Ld; i += kThreads) {
    const int j = i / kW1Ld, k = i % kW1Ld;
    sm.w1[j][k] = k < E ? w1[j * E + k] : __float2half_rn(0.f);
  }
  const float* wo = reinterpret_cast<const float*>(a.packed + packed_wout_off(a.net));
  for (int j = tid; j < kW; j += kThreads) sm.wout[j] = wo[j];
}

// This thread's sample encoded straight into row `lane` of the warp's fp16 tile (features [E, Epad) are 1, the rest 0).
// The level loop is ROLLED (two levels per trip): fully unrolled, the two 16-level x 8-corner bodies made the kernel
// ~310 KB of SASS and 38 % of all stall samples were instruction-cache misses ("no instruction").
__device__ __forceinline__ void encode_to_tile(const HashNet& net, const __half2* __restrict__ table, const float (&x)[3],
                                               __half* __restrict__ row) {
#pragma unroll 2
  for (int l = 0; l < net.n_levels; ++l) {
    const Cell q = locate(net.scale[l], x);
    const __half2* t = table + net.offset[l];
    const uint32_t res = net.res[l], ent = net.entries[l];
    const bool dense = net.dense[l] != 0u;
    float ax = 0.f, ay = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t idx = entry_index(q.c[0] + (c & 1), q.c[1] + ((c >> 1) & 1), q.c[2] + (c >> 2), res, ent, dense);
      const float2 v = __half22float2(__ldg(t + idx));
      const float w = corner_weight(q, c);
      ax = fmaf(w, v.x, ax);
      ay = fmaf(w, v.y, ay);
    }
    *reinterpret_cast<uint32_t*>(row + 2 * l) = pack_h2(ax, ay);
  }
  for (int i = 2 * net.n_levels; i < 2 * kMaxLevels; i += 2)
    *reinterpret_cast<uint32_t*>(row + i) = i < net.Epad ? pack_h2(1.0f, 1.0f) : 0u;
}

// hidden pre-activations of 16 samples (m-tile mt of the warp tile): c[nt][.] = C fragment of hidden columns 8 nt .. 8 nt + 7
__device__ __forceinline__ void hidden_gemm(const MmaSmem& sm, const WarpTiles& wt, int mt, int lane, float (&c)[8][4]) {
  uint32_t a0[4], a1[4];
  ldsm4(a0, &wt.enc[mt * 16 + (lane & 15)][(lane >> 4) * 8]);
  ldsm4(a1, &wt.enc[mt * 16 + (lane & 15)][16 + (lane >> 4) * 8]);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t b[4];                   // W1[n][k], k contiguous = the "col" operand as stored: k 0-7 | 8-15 | 16-23 | 24-31
    ldsm4(b, &sm.w1[nt * 8 + (lane & 7)][(lane >> 3) * 8]);
    c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
    mma16816(c[nt], a0, b[0], b[1]);
    mma16816(c[nt], a1, b[2], b[3]);
  }
}

__global__ void __launch_bounds__(kThreads, 2) hash_fwd_mma_kernel(const Args a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  MmaSmem& sm = *reinterpret_cast<MmaSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  load_weights_mma(a, sm, tid);
  __syncthreads();
  WarpTiles& wt = sm.wt[warp];
  const __half2* table = reinterpret_cast<const __half2*>(a.packed + packed_table_off(a.net));
  const int64_t n_tiles = (a.P + kThreads - 1) / kThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t s = tile * kThreads + tid;
    float x[3] = {0.f, 0.f, 0.f};
    if (s < a.P) position01(a, s, x);
    __syncwarp();
    encode_to_tile(a.net, table, x, &wt.enc[lane][0]);
    __syncwarp();
    const int64_t row0 = tile * kThreads + warp * 32;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float c[8][4];
      hidden_gemm(sm, wt, mt, lane, c);
      float s_lo = 0.f, s_hi = 0.f;          // rows 16 mt + g and 16 mt + g + 8
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float w0 = sm.wout[nt * 8 + 2 * t], w1 = sm.wout[nt * 8 + 2 * t + 1];
        s_lo = fmaf(round_h(fmaxf(c[nt][0], 0.f)), w0, s_lo);
        s_lo = fmaf(round_h(fmaxf(c[nt][1], 0.f)), w1, s_lo);
        s_hi = fmaf(round_h(fmaxf(c[nt][2], 0.f)), w0, s_hi);
        s_hi = fmaf(round_h(fmaxf(c[nt][3], 0.f)), w1, s_hi);
      }
      s_lo += __shfl_xor_sync(kFull, s_lo, 1); s_lo += __shfl_xor_sync(kFull, s_lo, 2);
      s_hi += __shfl_xor_sync(kFull, s_hi, 1); s_hi += __shfl_xor_sync(kFull, s_hi, 2);
      if (t == 0) {
        const int64_t r_lo = row0 + mt * 16 + g, r_hi = r_lo + 8;
        if (r_lo < a.P) a.sigma[r_lo] = s_lo;
        if (r_hi < a.P) a.sigma[r_hi] = s_hi;
      }
    }
  }
}

// Table-gradient scatter of one sample per lane (denc_row = its 32 encoding gradients), and d_pos through the
// interpolation weights.
template <bool kDx>
__device__ __forceinline__ void scatter_levels(const Args& a, const HashNet& net, const __half2* __restrict__ table,
                                               const float (&x)[3], const float* denc_row, bool in, float ds, int lane,
                                               int64_t s) {
  // ---- scatter into the gradient table, and d_pos through the interpolation weights.
  // The COARSE levels are aggregated inside the warp first: the lanes of a warp are consecutive samples of one ray,
  // so on a coarse level they sit in the same cell for long runs, and near the sensor origin EVERY ray of the
  // keyframe hits the same few cells - up to 4 x 10^5 reductions per address and step, which the L2 serialises
  // (measured: 6.5 ms on ray-ordered samples sharing an origin against 4.0 ms on independent positions).  Each
  // maximal run of equal cells is summed with a segmented shuffle reduction and only its first lane issues the
  // eight vector reductions.
  float dx[3] = {0.f, 0.f, 0.f};
  const bool act = in && ds != 0.f;
  const int n_agg = ((net.flags >> 4) & 0xF) ? ((net.flags >> 4) & 0xF) : kAggLevels;   // LONER_HASH_AGG_LEVELS(n), A/B
#pragma unroll 2
  for (int l = 0; l < net.n_levels; ++l) {
    const Cell q = locate(net.scale[l], x);
    const uint32_t res = net.res[l], ent = net.entries[l];
    const bool dense = net.dense[l] != 0u;
    float2* gt = reinterpret_cast<float2*>(a.d_table) + net.offset[l];
    const float gx = act ? denc_row[2 * l] : 0.f, gy = act ? denc_row[2 * l + 1] : 0.f;
    const bool agg = l < n_agg && res <= 1024u;               // warp-uniform
    float v[16];
    bool issue = act;
    if (agg) {
      const uint32_t key = act ? q.c[0] + res * (q.c[1] + res * q.c[2]) : 0xFFFFFFFFu;
      const uint32_t prev = __shfl_up_sync(kFull, key, 1);
      const bool head = lane == 0 || prev != key;
      const uint32_t heads = __ballot_sync(kFull, head);
      const uint32_t later = heads & ~((2u << lane) - 1u);    // run heads behind this lane
      const int run_end = later ? (__ffs(later) - 2) : 31;
#pragma unroll
      for (int c = 0; c < 8; ++c) { const float w = corner_weight(q, c); v[2 * c] = w * gx; v[2 * c + 1] = w * gy; }
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const bool take = lane + off <= run_end;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float o = __shfl_down_sync(kFull, v[k], off);
          if (take) v[k] += o;
        }
      }
      issue = act && head;
    }
    float lx[3] = {0.f, 0.f, 0.f};
    if (act) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t idx = entry_index(q.c[0] + (c & 1), q.c[1] + ((c >> 1) & 1), q.c[2] + (c >> 2), res, ent, dense);
        if (agg) {
          if (issue) atomicAdd(gt + idx, make_float2(v[2 * c], v[2 * c + 1]));
        } else {
          const float w = corner_weight(q, c);
          atomicAdd(gt + idx, make_float2(w * gx, w * gy));
        }
        if (kDx) {
          const float2 tv = __half22float2(__ldg(table + net.offset[l] + idx));
          const float dot = tv.x * gx + tv.y * gy;
          const float w0 = (c & 1) ? q.f[0] : 1.0f - q.f[0], w1 = (c & 2) ? q.f[1] : 1.0f - q.f[1],
                      w2 = (c & 4) ? q.f[2] : 1.0f - q.f[2];
          lx[0] += ((c & 1) ? 1.0f : -1.0f) * w1 * w2 * dot;
          lx[1] += ((c & 2) ? 1.0f : -1.0f) * w0 * w2 * dot;
          lx[2] += ((c & 4) ? 1.0f : -1.0f) * w0 * w1 * dot;
        }
      }
      if (kDx) {
#pragma unroll
        for (int dd = 0; dd < 3; ++dd) dx[dd] = fmaf(net.scale[l], lx[dd], dx[dd]);
      }
    }
  }
  if (kDx && in) {
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) a.d_pos[s * 3 + dd] = 0.5f * dx[dd];     // x = (pos + 1) / 2
  }
}

template <bool kDx>
__global__ void __launch_bounds__(kThreads, 2) hash_bwd_mma_kernel(const Args a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  MmaSmem& sm = *reinterpret_cast<MmaSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  load_weights_mma(a, sm, tid);
  __syncthreads();
  WarpTiles& wt = sm.wt[warp];
  const HashNet& net = a.net;
  const __half2* table = reinterpret_cast<const __half2*>(a.packed + packed_table_off(net));
  // dW1 [64 hidden x 32 features] = 4 x 4 C tiles of 16 x 8; warp w owns tiles (m = w >> 1, n = 2 (w & 1) + {0, 1})
  float accW[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i) accW[i][0] = accW[i][1] = accW[i][2] = accW[i][3] = 0.f;
  float accOut[8][2];                      // dW_out partial: hidden columns 8 nt + 2 t + {0, 1}, summed over this thread's rows
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) accOut[nt][0] = accOut[nt][1] = 0.f;
  const float inv_g = 1.0f / a.gscale;
  const int64_t n_tiles = (a.P + kThreads - 1) / kThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t s = tile * kThreads + tid;
    const bool in = s < a.P;
    float x[3] = {0.f, 0.f, 0.f};
    if (in) position01(a, s, x);
    encode_to_tile(net, table, x, &wt.enc[lane][0]);
    const float ds = in ? __ldg(a.d_sigma + s) : 0.f;
    __syncwarp();
    // ---- forward recompute + dH + dEnc, 16 samples at a time
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float c[8][4];
      hidden_gemm(sm, wt, mt, lane, c);
      const float ds_lo = __shfl_sync(kFull, ds, mt * 16 + g), ds_hi = __shfl_sync(kFull, ds, mt * 16 + g + 8);
      uint32_t dhp[8][2];                  // dH as packed half2 in C layout = A fragments of the dEnc GEMM
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float w0 = sm.wout[nt * 8 + 2 * t], w1 = sm.wout[nt * 8 + 2 * t + 1];
        const float h00 = round_h(fmaxf(c[nt][0], 0.f)), h01 = round_h(fmaxf(c[nt][1], 0.f));
        const float h10 = round_h(fmaxf(c[nt][2], 0.f)), h11 = round_h(fmaxf(c[nt][3], 0.f));
        accOut[nt][0] += ds_lo * h00 + ds_hi * h10;
        accOut[nt][1] += ds_lo * h01 + ds_hi * h11;
        const float k = a.gscale;
        const float d00 = h00 > 0.f ? fminf(fmaxf(ds_lo * w0 * k, -65504.f), 65504.f) : 0.f;
        const float d01 = h01 > 0.f ? fminf(fmaxf(ds_lo * w1 * k, -65504.f), 65504.f) : 0.f;
        const float d10 = h10 > 0.f ? fminf(fmaxf(ds_hi * w0 * k, -65504.f), 65504.f) : 0.f;
        const float d11 = h11 > 0.f ? fminf(fmaxf(ds_hi * w1 * k, -65504.f), 65504.f) : 0.f;
        dhp[nt][0] = pack_h2(d00, d01);
        dhp[nt][1] = pack_h2(d10, d11);
        *reinterpret_cast<uint32_t*>(&wt.dh[mt * 16 + g][nt * 8 + 2 * t]) = dhp[nt][0];
        *reinterpret_cast<uint32_t*>(&wt.dh[mt * 16 + g + 8][nt * 8 + 2 * t]) = dhp[nt][1];
      }
      // dEnc [16 x 32] = dH [16 x 64] W1 [64 x 32]: B from W1[k = hidden][n = feature] (n contiguous) through ldmatrix.trans
      float d[4][4];
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) d[nf][0] = d[nf][1] = d[nf][2] = d[nf][3] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {     // pairs of 16-deep k tiles: hidden 32 kp .. 32 kp + 31
        const uint32_t a0[4] = {dhp[4 * kp][0], dhp[4 * kp][1], dhp[4 * kp + 1][0], dhp[4 * kp + 1][1]};
        const uint32_t a1[4] = {dhp[4 * kp + 2][0], dhp[4 * kp + 2][1], dhp[4 * kp + 3][0], dhp[4 * kp + 3][1]};
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) {
          uint32_t b[4];                   // rows k: 32 kp + {0-7, 8-15, 16-23, 24-31}, columns n = 8 nf ..
          ldsm4t(b, &sm.w1[32 * kp + lane][nf * 8]);
          mma16816(d[nf], a0, b[0], b[1]);
          mma16816(d[nf], a1, b[2], b[3]);
        }
      }
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) {
        wt.denc[mt * 16 + g][nf * 8 + 2 * t] = d[nf][0] * inv_g;
        wt.denc[mt * 16 + g][nf * 8 + 2 * t + 1] = d[nf][1] * inv_g;
        wt.denc[mt * 16 + g + 8][nf * 8 + 2 * t] = d[nf][2] * inv_g;
        wt.denc[mt * 16 + g + 8][nf * 8 + 2 * t + 1] = d[nf][3] * inv_g;
      }
    }
    __syncwarp();
    scatter_levels<kDx>(a, net, table, x, &wt.denc[lane][0], in, ds, lane, s);
    __syncthreads();
    // ---- dW1 [64 x 32] += dH^T [64 x 256] Enc [256 x 32] over the CTA tile: warp w owns C tiles (m = w >> 1, n = 2 (w & 1) + {0, 1})
    {
      const int mtile = warp >> 1, npair = warp & 1;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) {           // the 32-sample tiles of the eight warps
        const WarpTiles& o = sm.wt[w8];
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          uint32_t af[4], bf[4];
          // A = dH^T: stored [k = sample][m = hidden] -> transposed load; tiles (m lo, k lo) (m hi, k lo) (m lo, k hi) (m hi, k hi)
          ldsm4t(af, &o.dh[kt * 16 + (lane & 7) + ((lane >> 4) & 1) * 8][mtile * 16 + ((lane >> 3) & 1) * 8]);
          // B = Enc: stored [k = sample][n = feature] -> transposed load; (k lo, n0) (k hi, n0) (k lo, n1) (k hi, n1)
          ldsm4t(bf, &o.enc[kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][npair * 16 + ((lane >> 4) & 1) * 8]);
          mma16816(accW[0], af, bf[0], bf[1]);
          mma16816(accW[1], af, bf[2], bf[3]);
        }
      }
    }
    __syncthreads();
  }
  // ---- per-CTA partials: dW1 (loss-scaled) in [hidden j][feature i] order with row stride Epad, then dW_out
  float* part = a.partials + (int64_t)blockIdx.x * (w1_floats(net) + kW);
  {
    const int mtile = warp >> 1, npair = warp & 1;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int col = npair * 16 + i * 8 + 2 * t;
      const int r0 = mtile * 16 + g;
      if (col < net.Epad) { part[r0 * net.Epad + col] = accW[i][0]; part[(r0 + 8) * net.Epad + col] = accW[i][2]; }
      if (col + 1 < net.Epad) { part[r0 * net.Epad + col + 1] = accW[i][1]; part[(r0 + 8) * net.Epad + col + 1] = accW[i][3]; }
    }
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float v = accOut[nt][j];
      v += __shfl_xor_sync(kFull, v, 4); v += __shfl_xor_sync(kFull, v, 8); v += __shfl_xor_sync(kFull, v, 16);
      if (g == 0) sm.red[warp][nt * 8 + 2 * t + j] = v;
    }
  __syncthreads();
  if (tid < kW) {
    float v = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) v += sm.red[w8][tid];
    part[w1_floats(net) + tid] = v;
  }
}

// ------------------------------------------------------------------------------------------
// Heads with 2 .. kMaxHidden hidden layers of 64 neurons (tcnn's FullyFusedMLP takes any n_hidden_layers; the shipped
// yaml uses 1, which stays on the kernels above).  Same structure, with the extra 64 x 64 GEMMs chained in registers:
// the C fragments of layer l (ReLU, fp16) ARE the A fragments of layer l + 1.  The backward keeps h_1 .. h_{L-1} of the
// CTA's 256 samples in shared memory for the weight gradients, one ReLU mask word per layer and m-tile in a register,
// and walks the layers down with a CTA-level dW_l GEMM per layer.
constexpr int kWhLd = 72;           // halves per row of a hidden matrix [64][64 + 8]
template <int kL>
struct DeepTiles {
  __align__(16) __half enc[32][kEncLd];
  __align__(16) __half dh[32][kDhLd];             // dH_l of the layer being processed
  __align__(16) __half h[kL - 1][32][kDhLd];      // h_1 .. h_{L-1}; h[kL-2] is dead after dW_L and then holds dEnc (fp32 [32][33])
};
template <int kL>
struct DeepSmem {
  __align__(16) __half w1[kW][kW1Ld];
  __align__(16) __half wh[kL - 1][kW][kWhLd];
  float wout[kW];
  float red[8][kW];
  DeepTiles<kL> wt[8];
};

template <int kL>
__device__ __forceinline__ void load_weights_deep(const Args& a, DeepSmem<kL>& sm, int tid) {
  const __half* w1 = reinterpret_cast<const __half*>(a.packed);
  const int E = a.net.Epad;
  for (int i = tid; i < kW * kW1Ld; i += kThreads) {
    const int j = i / kW1Ld, k = i % kW1Ld;
    sm.w1[j][k] = k < E ? w1[j * E + k] : __float2half_rn(0.f);
  }
  const __half* wh = w1 + w1_floats(a.net);
  for (int i = tid; i < (kL - 1) * kW * kW; i += kThreads) {
    const int l = i / (kW * kW), j = (i / kW) % kW, k = i % kW;
    sm.wh[l][j][k] = wh[i];
  }
  const float* wo = reinterpret_cast<const float*>(a.packed + packed_wout_off(a.net));
  for (int j = tid; j < kW; j += kThreads) sm.wout[j] = wo[j];
}

// first layer: pre-activations of 16 samples (m-tile mt) from the warp's encoding tile
__device__ __forceinline__ void first_gemm(const __half (*w1)[kW1Ld], const __half (*enc)[kEncLd], int mt, int lane, float (&c)[8][4]) {
  uint32_t a0[4], a1[4];
  ldsm4(a0, &enc[mt * 16 + (lane & 15)][(lane >> 4) * 8]);
  ldsm4(a1, &enc[mt * 16 + (lane & 15)][16 + (lane >> 4) * 8]);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t b[4];
    ldsm4(b, &w1[nt * 8 + (lane & 7)][(lane >> 3) * 8]);
    c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
    mma16816(c[nt], a0, b[0], b[1]);
    mma16816(c[nt], a1, b[2], b[3]);
  }
}
// hidden layer: c = hp [16 x 64] W^T, hp = the previous layer's activations as A fragments (C layout, packed half2)
__device__ __forceinline__ void next_gemm(const __half (*W)[kWhLd], const uint32_t (&hp)[8][2], int lane, float (&c)[8][4]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t b0[4], b1[4];           // W[n][k], k contiguous: k 0-31 and 32-63
    ldsm4(b0, &W[nt * 8 + (lane & 7)][(lane >> 3) * 8]);
    ldsm4(b1, &W[nt * 8 + (lane & 7)][32 + (lane >> 3) * 8]);
    c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
    const uint32_t a0[4] = {hp[0][0], hp[0][1], hp[1][0], hp[1][1]}, a1[4] = {hp[2][0], hp[2][1], hp[3][0], hp[3][1]};
    const uint32_t a2[4] = {hp[4][0], hp[4][1], hp[5][0], hp[5][1]}, a3[4] = {hp[6][0], hp[6][1], hp[7][0], hp[7][1]};
    mma16816(c[nt], a0, b0[0], b0[1]);
    mma16816(c[nt], a1, b0[2], b0[3]);
    mma16816(c[nt], a2, b1[0], b1[1]);
    mma16816(c[nt], a3, b1[2], b1[3]);
  }
}
// ReLU + fp16 rounding of a C tile: packed activations, and the mask word (bit 4 nt + e = element e of n-tile nt is active)
__device__ __forceinline__ uint32_t relu_pack(const float (&c)[8][4], uint32_t (&hp)[8][2]) {
  uint32_t m = 0u;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    hp[nt][0] = pack_h2(fmaxf(c[nt][0], 0.f), fmaxf(c[nt][1], 0.f));
    hp[nt][1] = pack_h2(fmaxf(c[nt][2], 0.f), fmaxf(c[nt][3], 0.f));
    // active = the ROUNDED activation is positive (what the backward of an fp16 network sees)
    const __half2 lo = *reinterpret_cast<const __half2*>(&hp[nt][0]), hi = *reinterpret_cast<const __half2*>(&hp[nt][1]);
    m |= (__low2float(lo) > 0.f ? 1u : 0u) << (4 * nt) | (__high2float(lo) > 0.f ? 2u : 0u) << (4 * nt) |
         (__low2float(hi) > 0.f ? 4u : 0u) << (4 * nt) | (__high2float(hi) > 0.f ? 8u : 0u) << (4 * nt);
  }
  return m;
}
__device__ __forceinline__ float sat_h(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }

template <int kL>
__global__ void __launch_bounds__(kThreads, 1) hash_fwd_deep_kernel(const Args a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  DeepSmem<kL>& sm = *reinterpret_cast<DeepSmem<kL>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  load_weights_deep<kL>(a, sm, tid);
  __syncthreads();
  DeepTiles<kL>& wt = sm.wt[warp];
  const __half2* table = reinterpret_cast<const __half2*>(a.packed + packed_table_off(a.net));
  const int64_t n_tiles = (a.P + kThreads - 1) / kThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t s = tile * kThreads + tid;
    float x[3] = {0.f, 0.f, 0.f};
    if (s < a.P) position01(a, s, x);
    __syncwarp();
    encode_to_tile(a.net, table, x, &wt.enc[lane][0]);
    __syncwarp();
    const int64_t row0 = tile * kThreads + warp * 32;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float c[8][4];
      first_gemm(sm.w1, wt.enc, mt, lane, c);
#pragma unroll
      for (int l = 2; l <= kL; ++l) {
        uint32_t hp[8][2];
        relu_pack(c, hp);
        next_gemm(sm.wh[l - 2], hp, lane, c);
      }
      float s_lo = 0.f, s_hi = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float w0 = sm.wout[nt * 8 + 2 * t], w1 = sm.wout[nt * 8 + 2 * t + 1];
        s_lo = fmaf(round_h(fmaxf(c[nt][0], 0.f)), w0, s_lo);
        s_lo = fmaf(round_h(fmaxf(c[nt][1], 0.f)), w1, s_lo);
        s_hi = fmaf(round_h(fmaxf(c[nt][2], 0.f)), w0, s_hi);
        s_hi = fmaf(round_h(fmaxf(c[nt][3], 0.f)), w1, s_hi);
      }
      s_lo += __shfl_xor_sync(kFull, s_lo, 1); s_lo += __shfl_xor_sync(kFull, s_lo, 2);
      s_hi += __shfl_xor_sync(kFull, s_hi, 1); s_hi += __shfl_xor_sync(kFull, s_hi, 2);
      if (t == 0) {
        const int64_t r_lo = row0 + mt * 16 + g, r_hi = r_lo + 8;
        if (r_lo < a.P) a.sigma[r_lo] = s_lo;
        if (r_hi < a.P) a.sigma[r_hi] = s_hi;
      }
    }
  }
}

template <int kL, bool kDx>
__global__ void __launch_bounds__(kThreads, 1) hash_bwd_deep_kernel(const Args a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  DeepSmem<kL>& sm = *reinterpret_cast<DeepSmem<kL>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  load_weights_deep<kL>(a, sm, tid);
  __syncthreads();
  DeepTiles<kL>& wt = sm.wt[warp];
  float (*denc)[kDencLd] = reinterpret_cast<float (*)[kDencLd]>(&wt.h[kL - 2][0][0]);
  const HashNet& net = a.net;
  const __half2* table = reinterpret_cast<const __half2*>(a.packed + packed_table_off(net));
  const int mtile = warp >> 1, nhalf = warp & 1;      // this warp's C tiles of the CTA-level weight-gradient GEMMs
  float accW[2][4];                                   // dW1: rows 16 mtile .., columns 16 nhalf + 8 i ..
  float accH[kL - 1][4][4];                           // dW_l (l = 2 .. L): rows 16 mtile .., columns 32 nhalf + 8 i ..
  float accOut[8][2];
#pragma unroll
  for (int i = 0; i < 2; ++i) accW[i][0] = accW[i][1] = accW[i][2] = accW[i][3] = 0.f;
#pragma unroll
  for (int l = 0; l < kL - 1; ++l)
#pragma unroll
    for (int i = 0; i < 4; ++i) accH[l][i][0] = accH[l][i][1] = accH[l][i][2] = accH[l][i][3] = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) accOut[nt][0] = accOut[nt][1] = 0.f;
  const float inv_g = 1.0f / a.gscale;
  const int64_t n_tiles = (a.P + kThreads - 1) / kThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t s = tile * kThreads + tid;
    const bool in = s < a.P;
    float x[3] = {0.f, 0.f, 0.f};
    if (in) position01(a, s, x);
    encode_to_tile(net, table, x, &wt.enc[lane][0]);
    const float ds = in ? __ldg(a.d_sigma + s) : 0.f;
    __syncwarp();
    // ---- forward recompute of both m-tiles: h_1 .. h_{L-1} to shared memory, their masks and dH_L to registers
    uint32_t dhp[2][8][2];
    uint32_t msk[2][kL - 1];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float c[8][4];
      first_gemm(sm.w1, wt.enc, mt, lane, c);
#pragma unroll
      for (int l = 1; l < kL; ++l) {
        uint32_t hp[8][2];
        msk[mt][l - 1] = relu_pack(c, hp);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          *reinterpret_cast<uint32_t*>(&wt.h[l - 1][mt * 16 + g][nt * 8 + 2 * t]) = hp[nt][0];
          *reinterpret_cast<uint32_t*>(&wt.h[l - 1][mt * 16 + g + 8][nt * 8 + 2 * t]) = hp[nt][1];
        }
        next_gemm(sm.wh[l - 1], hp, lane, c);
      }
      const float ds_lo = __shfl_sync(kFull, ds, mt * 16 + g), ds_hi = __shfl_sync(kFull, ds, mt * 16 + g + 8);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float w0 = sm.wout[nt * 8 + 2 * t], w1 = sm.wout[nt * 8 + 2 * t + 1];
        const float h00 = round_h(fmaxf(c[nt][0], 0.f)), h01 = round_h(fmaxf(c[nt][1], 0.f));
        const float h10 = round_h(fmaxf(c[nt][2], 0.f)), h11 = round_h(fmaxf(c[nt][3], 0.f));
        accOut[nt][0] += ds_lo * h00 + ds_hi * h10;
        accOut[nt][1] += ds_lo * h01 + ds_hi * h11;
        const float k = a.gscale;
        dhp[mt][nt][0] = pack_h2(h00 > 0.f ? sat_h(ds_lo * w0 * k) : 0.f, h01 > 0.f ? sat_h(ds_lo * w1 * k) : 0.f);
        dhp[mt][nt][1] = pack_h2(h10 > 0.f ? sat_h(ds_hi * w0 * k) : 0.f, h11 > 0.f ? sat_h(ds_hi * w1 * k) : 0.f);
      }
    }
    // ---- layers L .. 2: dW_l += dH_l^T h_{l-1} over the CTA's 256 samples, then dH_{l-1} = (dH_l W_l) * relu'(h_{l-1})
#pragma unroll
    for (int l = kL; l >= 2; --l) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          *reinterpret_cast<uint32_t*>(&wt.dh[mt * 16 + g][nt * 8 + 2 * t]) = dhp[mt][nt][0];
          *reinterpret_cast<uint32_t*>(&wt.dh[mt * 16 + g + 8][nt * 8 + 2 * t]) = dhp[mt][nt][1];
        }
      __syncthreads();
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) {
        const DeepTiles<kL>& o = sm.wt[w8];
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          uint32_t af[4];
          ldsm4t(af, &o.dh[kt * 16 + (lane & 7) + ((lane >> 4) & 1) * 8][mtile * 16 + ((lane >> 3) & 1) * 8]);
#pragma unroll
          for (int np = 0; np < 2; ++np) {
            uint32_t bf[4];
            ldsm4t(bf, &o.h[l - 2][kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][nhalf * 32 + np * 16 + ((lane >> 4) & 1) * 8]);
            mma16816(accH[l - 2][2 * np], af, bf[0], bf[1]);
            mma16816(accH[l - 2][2 * np + 1], af, bf[2], bf[3]);
          }
        }
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float d[8][4];
#pragma unroll
        for (int nf = 0; nf < 8; ++nf) d[nf][0] = d[nf][1] = d[nf][2] = d[nf][3] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          const uint32_t a0[4] = {dhp[mt][4 * kp][0], dhp[mt][4 * kp][1], dhp[mt][4 * kp + 1][0], dhp[mt][4 * kp + 1][1]};
          const uint32_t a1[4] = {dhp[mt][4 * kp + 2][0], dhp[mt][4 * kp + 2][1], dhp[mt][4 * kp + 3][0], dhp[mt][4 * kp + 3][1]};
#pragma unroll
          for (int nf = 0; nf < 8; ++nf) {
            uint32_t b[4];                 // W_l[k = out][n = in], n contiguous -> transposed load
            ldsm4t(b, &sm.wh[l - 2][32 * kp + lane][nf * 8]);
            mma16816(d[nf], a0, b[0], b[1]);
            mma16816(d[nf], a1, b[2], b[3]);
          }
        }
        const uint32_t m = msk[mt][l - 2];
#pragma unroll
        for (int nf = 0; nf < 8; ++nf) {
          dhp[mt][nf][0] = pack_h2((m >> (4 * nf)) & 1u ? sat_h(d[nf][0]) : 0.f, (m >> (4 * nf + 1)) & 1u ? sat_h(d[nf][1]) : 0.f);
          dhp[mt][nf][1] = pack_h2((m >> (4 * nf + 2)) & 1u ? sat_h(d[nf][2]) : 0.f, (m >> (4 * nf + 3)) & 1u ? sat_h(d[nf][3]) : 0.f);
        }
      }
      __syncthreads();                     // every warp is done with this layer's dh / h tiles
    }
    // ---- layer 1: dH_1 to shared memory, dEnc = dH_1 W1, scatter, dW1 (as in the one-layer kernel)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<uint32_t*>(&wt.dh[mt * 16 + g][nt * 8 + 2 * t]) = dhp[mt][nt][0];
        *reinterpret_cast<uint32_t*>(&wt.dh[mt * 16 + g + 8][nt * 8 + 2 * t]) = dhp[mt][nt][1];
      }
      float d[4][4];
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) d[nf][0] = d[nf][1] = d[nf][2] = d[nf][3] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        const uint32_t a0[4] = {dhp[mt][4 * kp][0], dhp[mt][4 * kp][1], dhp[mt][4 * kp + 1][0], dhp[mt][4 * kp + 1][1]};
        const uint32_t a1[4] = {dhp[mt][4 * kp + 2][0], dhp[mt][4 * kp + 2][1], dhp[mt][4 * kp + 3][0], dhp[mt][4 * kp + 3][1]};
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) {
          uint32_t b[4];
          ldsm4t(b, &sm.w1[32 * kp + lane][nf * 8]);
          mma16816(d[nf], a0, b[0], b[1]);
          mma16816(d[nf], a1, b[2], b[3]);
        }
      }
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) {
        denc[mt * 16 + g][nf * 8 + 2 * t] = d[nf][0] * inv_g;
        denc[mt * 16 + g][nf * 8 + 2 * t + 1] = d[nf][1] * inv_g;
        denc[mt * 16 + g + 8][nf * 8 + 2 * t] = d[nf][2] * inv_g;
        denc[mt * 16 + g + 8][nf * 8 + 2 * t + 1] = d[nf][3] * inv_g;
      }
    }
    __syncwarp();
    scatter_levels<kDx>(a, net, table, x, &denc[lane][0], in, ds, lane, s);
    __syncthreads();
    {
      const int npair = nhalf;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) {
        const DeepTiles<kL>& o = sm.wt[w8];
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          uint32_t af[4], bf[4];
          ldsm4t(af, &o.dh[kt * 16 + (lane & 7) + ((lane >> 4) & 1) * 8][mtile * 16 + ((lane >> 3) & 1) * 8]);
          ldsm4t(bf, &o.enc[kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][npair * 16 + ((lane >> 4) & 1) * 8]);
          mma16816(accW[0], af, bf[0], bf[1]);
          mma16816(accW[1], af, bf[2], bf[3]);
        }
      }
    }
    __syncthreads();
  }
  // ---- per-CTA partials: dW1 [64][Epad] | dW_2 .. dW_L [64][64] (loss-scaled) | dW_out [64]
  float* part = a.partials + (int64_t)blockIdx.x * (w1_floats(net) + wh_floats(net) + kW);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int col = nhalf * 16 + i * 8 + 2 * t;
    const int r0 = mtile * 16 + g;
    if (col < net.Epad) { part[r0 * net.Epad + col] = accW[i][0]; part[(r0 + 8) * net.Epad + col] = accW[i][2]; }
    if (col + 1 < net.Epad) { part[r0 * net.Epad + col + 1] = accW[i][1]; part[(r0 + 8) * net.Epad + col + 1] = accW[i][3]; }
  }
#pragma unroll
  for (int l = 0; l < kL - 1; ++l) {
    float* ph = part + w1_floats(net) + (int64_t)l * kW * kW;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int col = nhalf * 32 + i * 8 + 2 * t;
      const int r0 = mtile * 16 + g;
      ph[r0 * kW + col] = accH[l][i][0]; ph[r0 * kW + col + 1] = accH[l][i][1];
      ph[(r0 + 8) * kW + col] = accH[l][i][2]; ph[(r0 + 8) * kW + col + 1] = accH[l][i][3];
    }
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float v = accOut[nt][j];
      v += __shfl_xor_sync(kFull, v, 4); v += __shfl_xor_sync(kFull, v, 8); v += __shfl_xor_sync(kFull, v, 16);
      if (g == 0) sm.red[warp][nt * 8 + 2 * t + j] = v;
    }
  __syncthreads();
  if (tid < kW) {
    float v = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) v += sm.red[w8][tid];
    part[w1_floats(net) + wh_floats(net) + tid] = v;
  }
}

// d_params[W1] += sum_b partials[b][W1] / gscale;  d_params[W_out row 0] += sum_b partials[b][W_out]
__global__ void __launch_bounds__(256) hash_reduce_kernel(HashNet net, const float* __restrict__ partials, int n_blocks,
                                                         float inv_gscale, float* __restrict__ d_params) {
  const int64_t nw1 = w1_floats(net) + wh_floats(net), per = nw1 + kW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < n_blocks; ++b) s += partials[(int64_t)b * per + i];
    if (i < nw1) d_params[i] += s * inv_gscale;
    else d_params[nw1 + (i - nw1)] += s;
  }
}

inline int sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      sms <= 0) {
    cudaGetLastError();
    sms = 148;
  }
  return sms;
}
inline int bwd_blocks() { return sm_count() * 2; }      // 85 KB of shared memory per CTA: two per SM

}  // namespace hashgrid
}  // namespace loner

using namespace loner::hashgrid;

extern "C" int64_t loner_hash_param_count(const loner_hashnet_t* n) {
  HashNet net;
  if (!net_from(n, net)) return -1;
  return net_floats(net) + n_entries(net) * 2;
}
extern "C" int64_t loner_hash_table_entries(const loner_hashnet_t* n) {
  HashNet net;
  if (!net_from(n, net)) return -1;
  return n_entries(net);
}
extern "C" int64_t loner_hash_packed_bytes(const loner_hashnet_t* n) {
  HashNet net;
  if (!net_from(n, net)) return -1;
  return packed_total(net);
}
extern "C" int64_t loner_hash_bwd_scratch_bytes(const loner_hashnet_t* n, int64_t P) {
  HashNet net;
  if (!net_from(n, net) || P < 0) return -1;
  return (int64_t)bwd_blocks() * (w1_floats(net) + wh_floats(net) + kW) * 4;
}

extern "C" int loner_hash_pack(const loner_hashnet_t* n, const float* params, void* packed, void* stream) {
  HashNet net;
  if (!net_from(n, net)) return LONER_E_UNSUPPORTED;
  if (!params || !packed) return LONER_E_BAD_ARG;
  pack_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(net, params, (uint8_t*)packed);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

static int fill_args(const loner_hashnet_t* n, Args& a, const void* packed, const float* pos, const float* rays,
                     const float* z_vals, int32_t S, int64_t P) {
  if (!net_from(n, a.net)) return LONER_E_UNSUPPORTED;
  if (!packed || P < 0 || (!pos && (!rays || !z_vals || S <= 0))) return LONER_E_BAD_ARG;
  a.packed = (const uint8_t*)packed; a.pos = pos; a.rays = rays; a.z = z_vals; a.S = S > 0 ? S : 1; a.P = P;
  a.sigma = nullptr; a.d_sigma = nullptr; a.gscale = 1.f; a.d_table = nullptr; a.d_pos = nullptr; a.partials = nullptr;
  return LONER_OK;
}

extern "C" int loner_hash_fwd(const loner_hashnet_t* n, const void* packed, const float* pos, const float* rays,
                              const float* z_vals, int32_t S, int64_t P, float* sigma, void* stream) {
  if (P == 0) { HashNet t; return net_from(n, t) ? LONER_OK : LONER_E_UNSUPPORTED; }
  Args a;
  const int rc = fill_args(n, a, packed, pos, rays, z_vals, S, P);
  if (rc) return rc;
  if (!sigma) return LONER_E_BAD_ARG;
  a.sigma = sigma;
  const int64_t want = (P + kThreads - 1) / kThreads;
  if (a.net.L > 1) {
    const int64_t cap = sm_count();
    const unsigned grid = (unsigned)(want < cap ? want : cap);
#define LONER_HASH_FWD_DEEP(L_)                                                                           \
  do {                                                                                                    \
    const int smem = (int)sizeof(DeepSmem<L_>);                                                           \
    cudaFuncSetAttribute(hash_fwd_deep_kernel<L_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);    \
    hash_fwd_deep_kernel<L_><<<grid, kThreads, smem, (cudaStream_t)stream>>>(a);                          \
  } while (0)
    if (a.net.L == 2) LONER_HASH_FWD_DEEP(2); else if (a.net.L == 3) LONER_HASH_FWD_DEEP(3); else LONER_HASH_FWD_DEEP(4);
#undef LONER_HASH_FWD_DEEP
  } else if (a.net.flags & LONER_HASH_SCALAR) {
    const int64_t cap = (int64_t)sm_count() * 8;
    hash_fwd_kernel<<<(unsigned)(want < cap ? want : cap), kThreads, 0, (cudaStream_t)stream>>>(a);
  } else {
    const int64_t cap = (int64_t)sm_count() * 2;
    const int smem = (int)sizeof(MmaSmem);
    cudaFuncSetAttribute(hash_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    hash_fwd_mma_kernel<<<(unsigned)(want < cap ? want : cap), kThreads, smem, (cudaStream_t)stream>>>(a);
  }
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_hash_bwd(const loner_hashnet_t* n, const void* packed, const float* pos, const float* rays,
                              const float* z_vals, int32_t S, int64_t P, const float* d_sigma, float grad_scale,
                              float* d_params, float* d_pos, void* scratch, void* stream) {
  if (P == 0) { HashNet t; return net_from(n, t) ? LONER_OK : LONER_E_UNSUPPORTED; }
  Args a;
  const int rc = fill_args(n, a, packed, pos, rays, z_vals, S, P);
  if (rc) return rc;
  if (!d_sigma || !d_params || !scratch || !(grad_scale > 0.f)) return LONER_E_BAD_ARG;
  a.d_sigma = d_sigma; a.gscale = grad_scale; a.d_pos = d_pos; a.partials = (float*)scratch;
  a.d_table = d_params + net_floats(a.net);
  const int64_t tiles = (P + kThreads - 1) / kThreads;
  const int max_blocks = a.net.L > 1 ? sm_count() : bwd_blocks();     // the deep kernels hold one CTA per SM
  const int blocks = (int)(tiles < max_blocks ? tiles : max_blocks);
  cudaStream_t st = (cudaStream_t)stream;
  const bool scalar = (a.net.flags & LONER_HASH_SCALAR) != 0;
  const int smem = scalar ? (int)sizeof(BwdSmem) : (int)sizeof(MmaSmem);
#define LONER_HASH_BWD(K)                                                                        \
  do {                                                                                           \
    cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                  \
    K<<<blocks, kThreads, smem, st>>>(a);                                                        \
  } while (0)
#define LONER_HASH_BWD_DEEP(L_)                                                                          \
  do {                                                                                                   \
    const int dsmem = (int)sizeof(DeepSmem<L_>);                                                         \
    if (d_pos) {                                                                                         \
      cudaFuncSetAttribute(hash_bwd_deep_kernel<L_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dsmem);  \
      hash_bwd_deep_kernel<L_, true><<<blocks, kThreads, dsmem, st>>>(a);                                \
    } else {                                                                                             \
      cudaFuncSetAttribute(hash_bwd_deep_kernel<L_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dsmem); \
      hash_bwd_deep_kernel<L_, false><<<blocks, kThreads, dsmem, st>>>(a);                               \
    }                                                                                                    \
  } while (0)
  if (a.net.L == 2) LONER_HASH_BWD_DEEP(2);
  else if (a.net.L == 3) LONER_HASH_BWD_DEEP(3);
  else if (a.net.L == 4) LONER_HASH_BWD_DEEP(4);
  else if (scalar) { if (d_pos) LONER_HASH_BWD(hash_bwd_kernel<true>); else LONER_HASH_BWD(hash_bwd_kernel<false>); }
  else             { if (d_pos) LONER_HASH_BWD(hash_bwd_mma_kernel<true>); else LONER_HASH_BWD(hash_bwd_mma_kernel<false>); }
#undef LONER_HASH_BWD_DEEP
#undef LONER_HASH_BWD
  LONER_CHECK_LAUNCH();
  hash_reduce_kernel<<<16, 256, 0, st>>>(a.net, a.partials, blocks, 1.0f / grad_scale, d_params);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

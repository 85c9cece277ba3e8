// Per-ray sample placement.  One warp owns one ray; everything between the ray row and the
// sorted z_vals row stays in shared memory / registers.
// Replaces UniformRaySampler.get_samples, OccGridRaySampler.get_samples
// (/root/reference/src/models/ray_sampling.py:22-43, :53-92), OccupancyGridModel.interpolate
// (/root/reference/src/models/model_tcnn.py:124-131, ATen grid_sampler_3d semantics) and
// sample_pdf (/root/reference/src/models/rendering_tcnn.py:18-67): linspace + jitter,
// trilinear occupancy fetch (the 4 MB grid is L2 resident), CDF scan, inverse-CDF search,
// bitonic sort of the importance half and a rank merge with the (already sorted) stratified
// half — instead of the reference's ~25 launches and ten [N,S/2] temporaries.
#include "common.cuh"

namespace loner {

__device__ __forceinline__ float strat_base(float near, float far, int j, int H) {
  const float t = linspace01(j, H);
  return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(far, t));
}

// ray_sampling.py:59-73 for one index j
__device__ __forceinline__ float strat_sample(float near, float far, int j, int H, float perturb, float u) {
  const float zj = strat_base(near, far, j, H);
  if (!(perturb > 0.f)) return zj;
  const float zl = (j > 0) ? strat_base(near, far, j - 1, H) : zj;
  const float zu = (j < H - 1) ? strat_base(near, far, j + 1, H) : zj;
  const float lower = (j > 0) ? __fmul_rn(0.5f, __fadd_rn(zl, zj)) : zj;
  const float upper = (j < H - 1) ? __fmul_rn(0.5f, __fadd_rn(zj, zu)) : zj;
  return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), __fmul_rn(perturb, u)));
}

// ATen grid_sampler_3d, bilinear, align_corners=False, zeros padding; grid[z][y][x].
__device__ __forceinline__ float trilinear(const float* __restrict__ g, int V, float x, float y, float z) {
  const float fV = (float)V;
  const float ix = (__fmul_rn(__fadd_rn(x, 1.f), fV) - 1.f) * 0.5f;
  const float iy = (__fmul_rn(__fadd_rn(y, 1.f), fV) - 1.f) * 0.5f;
  const float iz = (__fmul_rn(__fadd_rn(z, 1.f), fV) - 1.f) * 0.5f;
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  const float wx1 = ix - fx, wy1 = iy - fy, wz1 = iz - fz;
  const float wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy, wz0 = (fz + 1.f) - iz;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int xi = x0 + (c & 1), yi = y0 + ((c >> 1) & 1), zi = z0 + (c >> 2);
    const float w = ((c & 1) ? wx1 : wx0) * (((c >> 1) & 1) ? wy1 : wy0) * ((c >> 2) ? wz1 : wz0);
    if ((unsigned)xi < (unsigned)V && (unsigned)yi < (unsigned)V && (unsigned)zi < (unsigned)V)
      acc += __ldg(g + ((int64_t)zi * V + yi) * V + xi) * w;
  }
  return acc;
}

__global__ void __launch_bounds__(256)
sample_uniform_kernel(const float* __restrict__ rays, int64_t n, int S, float perturb,
                      const float* __restrict__ u, uint64_t seed, float* __restrict__ z_vals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * S) return;
  const int64_t r = i / S;
  const int j = (int)(i - r * S);
  const float near = rays[r * LONER_RAY_COLS + 11], far = rays[r * LONER_RAY_COLS + 12];
  float uu = 0.f;
  if (perturb > 0.f) uu = u ? u[i] : u32_to_unit(Philox(seed)((uint64_t)i, 0u).x);
  z_vals[i] = strat_sample(near, far, j, S, perturb, uu);
}

constexpr int kSampWarps = 4;

// smem per warp: zc[H] | cdf[H] | zi[Hp]
__global__ void __launch_bounds__(kSampWarps * 32)
sample_ogm_kernel(const float* __restrict__ rays, int64_t n, int S, int H, int Hp, float perturb,
                  const float* __restrict__ grid, int V, const float* __restrict__ u1,
                  const float* __restrict__ u2, uint64_t seed, float* __restrict__ z_vals) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kSampWarps + warp;
  if (ray >= n) return;
  float* zc = smem + (size_t)warp * (2 * H + Hp);
  float* cdf = zc + H;
  float* zi = cdf + H;
  const float* R = rays + ray * LONER_RAY_COLS;
  const float ox = R[0], oy = R[1], oz = R[2], dx = R[3], dy = R[4], dz = R[5];
  const float near = R[11], far = R[12];
  const Philox rng(seed);

  // (1) stratified half + occupancy probabilities              ray_sampling.py:59-81
  float wsum = 0.f;
  for (int j = lane; j < H; j += 32) {
    float uu = 0.f;
    if (perturb > 0.f) uu = u1 ? u1[ray * H + j] : u32_to_unit(rng((uint64_t)(ray * H + j), 1u).x);
    const float z = strat_sample(near, far, j, H, perturb, uu);
    zc[j] = z;
    const float px = __fadd_rn(ox, __fmul_rn(dx, z)), py = __fadd_rn(oy, __fmul_rn(dy, z)),
                pz = __fadd_rn(oz, __fmul_rn(dz, z));
    const float logit = trilinear(grid, V, px, py, pz);
    float prob = 1.0f / (1.0f + expf(-logit));
    prob = 2.0f * (fminf(fmaxf(prob, 0.5f), 1.0f) - 0.5f);
    if (j >= 1 && j <= H - 2) {      // weights = point_probs[:, 1:-1] + eps   rendering_tcnn.py:32
      const float w = prob + 1e-5f;
      cdf[j] = w;                    // temporarily: weight k=j-1 lives at cdf[j]
      wsum += w;
    }
  }
  wsum = warp_sum(wsum);
  __syncwarp();

  // (2) cdf[0]=0, cdf[k]=sum_{i<k} pdf_i, k=1..H-2  (H-1 entries)   rendering_tcnn.py:34-38
  const int nb = H - 2;
  {
    // torch.cumsum on the reference's CPU path accumulates fp32 rows in double (ATen acc_type) and
    // rounds each prefix to fp32; do the same: low-probability bins have increments near the fp32
    // rounding of the running sum, and the inverse CDF divides by them.
    const int chunk = (nb + 31) / 32;
    const int b = lane * chunk, e = min(b + chunk, nb);
    double local = 0.0;
    for (int k = b; k < e; ++k) local += (double)__fdiv_rn(cdf[k + 1], wsum);
    double incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    double run = incl - local;
    __syncwarp();
    for (int k = b; k < e; ++k) {
      run += (double)__fdiv_rn(cdf[k + 1], wsum);
      cdf[k + 1] = (float)run;       // cdf index k+1 = inclusive sum through weight k
    }
    if (lane == 0) cdf[0] = 0.f;
  }
  __syncwarp();

  // (3) inverse-CDF draws                                       rendering_tcnn.py:41-67
  for (int i = lane; i < Hp; i += 32) {
    if (i >= H) { zi[i] = __int_as_float(0x7f800000); continue; }
    const float u = u2 ? u2[ray * H + i] : u32_to_unit(rng((uint64_t)(ray * H + i), 2u).x);
    int lo = 0, hi = nb + 1;         // upper_bound over cdf[0..nb]  (searchsorted right=True)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int below = max(lo - 1, 0), above = min(lo, nb);
    const float c0 = cdf[below], c1 = cdf[above];
    const float b0 = __fmul_rn(0.5f, __fadd_rn(zc[below], zc[below + 1]));
    const float b1 = __fmul_rn(0.5f, __fadd_rn(zc[above], zc[above + 1]));
    float denom = __fsub_rn(c1, c0);
    if (denom < 1e-5f) denom = 1.0f;
    zi[i] = __fadd_rn(b0, __fmul_rn(__fdiv_rn(__fsub_rn(u, c0), denom), __fsub_rn(b1, b0)));
  }
  __syncwarp();

  // (4) bitonic sort of the importance half (Hp a power of two, padded with +inf)
  for (int k = 2; k <= Hp; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < Hp; i += 32) {
        const int p = i ^ j;
        if (p > i) {
          const float a = zi[i], b = zi[p];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { zi[i] = b; zi[p] = a; }
        }
      }
      __syncwarp();
    }
  }

  // (5) rank merge with the stratified half (already ascending)  ray_sampling.py:90
  float* out = z_vals + ray * S;
  for (int i = lane; i < H; i += 32) {
    const float v = zc[i];
    int lo = 0, hi = H;              // #importance samples strictly below v
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (zi[mid] < v) lo = mid + 1; else hi = mid; }
    out[i + lo] = v;
    const float w = zi[i];
    lo = 0; hi = H;                  // #stratified samples <= w
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (zc[mid] <= w) lo = mid + 1; else hi = mid; }
    out[i + lo] = w;
  }
}

}  // namespace loner

extern "C" int loner_sample_uniform(const float* rays, int64_t n, int32_t S, float perturb, const float* u,
                                    uint64_t seed, float* z_vals, void* stream) {
  if (n == 0) return LONER_OK;
  if (!rays || !z_vals || n < 0 || S < 2) return LONER_E_BAD_ARG;
  const int64_t total = n * S;
  loner::sample_uniform_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rays, n, S, perturb, u, seed, z_vals);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_sample_ogm(const float* rays, int64_t n, int32_t S, float perturb, const float* grid,
                                int32_t V, const float* u1, const float* u2, uint64_t seed, float* z_vals,
                                void* stream) {
  if (n == 0) return LONER_OK;
  if (!rays || !grid || !z_vals || n < 0 || V <= 0) return LONER_E_BAD_ARG;
  if (S < 8 || (S & 1)) return LONER_E_UNSUPPORTED;   // needs H-2 >= 2 weights
  const int H = S / 2;
  int Hp = 32;
  while (Hp < H) Hp <<= 1;
  const size_t smem = (size_t)loner::kSampWarps * (2 * H + Hp) * sizeof(float);
  if (smem > 200 * 1024) return LONER_E_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(loner::sample_ogm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const unsigned blocks = (unsigned)((n + loner::kSampWarps - 1) / loner::kSampWarps);
  loner::sample_ogm_kernel<<<blocks, loner::kSampWarps * 32, smem, (cudaStream_t)stream>>>(
      rays, n, S, H, Hp, perturb, grid, V, u1, u2, seed, z_vals);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

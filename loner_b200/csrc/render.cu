// Volume rendering + JS-divergence dynamic-margin depth loss, forward and backward, one warp
// per ray, per-sample temporaries in shared memory (the reference keeps ~12 [N,S] fp32
// temporaries in HBM for autograd).
// Replaces raw2outputs (/root/reference/src/models/rendering_tcnn.py:93-145), the loss half of
// Optimizer.compute_loss (/root/reference/src/mapping/optimizer.py:460-591),
// get_weights_gt (/root/reference/src/models/losses.py:29-51),
// calculate_JS_divergence (/root/reference/src/mapping/optimizer.py:614-626) and the autograd
// backward of all of them w.r.t. sigma and |d|.
#include <cstdint>
#include "common.cuh"

namespace loner {

constexpr int kRenderWarps = 4;

struct RayCtx {
  float* w;   // weights
  float* T;   // transmittance before the sample
  float* e;   // exp(-delta * r)
  float* r;   // relu(sigma + noise)
};

struct LossCfg {
  float scale, eps_min, min_js, max_js, alpha, los_lambda, depth_lambda;
  int l2;            // 0: L1 on the weights (L1_JS / L1_LOS), 1: MSE (L2_JS / L2_LOS)   optimizer.py:568-574
  float fixed_eps;   // > 0: *_LOS variants use this margin instead of the JS dynamic one  optimizer.py:516-523
};

__device__ __forceinline__ float noise_at(const float* noise, float std, const Philox& rng, int64_t idx) {
  if (noise) return std > 0.f ? noise[idx] * std : 0.f;
  if (!(std > 0.f)) return 0.f;
  const uint4 q = rng((uint64_t)(idx >> 1), 3u);
  const float2 g = box_muller(q.x, q.y);
  return ((idx & 1) ? g.y : g.x) * std;
}

// Forward over one ray: fills ctx arrays, returns A = sum w, Z = sum w z (cube units).
__device__ __forceinline__ void ray_forward(const float* __restrict__ sigma, const float* __restrict__ z,
                                            const float* __restrict__ noise, float noise_std, const Philox& rng,
                                            int64_t base, int S, float dnorm, int lane, const RayCtx& c,
                                            float& A, float& Z) {
  float carry = 1.0f, a_sum = 0.f, z_sum = 0.f;
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    const bool ok = s < S;
    float zi = 0.f, zn = 0.f, sg = 0.f, nz = 0.f;
    if (ok) {
      zi = z[base + s];
      zn = (s + 1 < S) ? z[base + s + 1] : 0.f;
      sg = sigma[base + s];
      nz = noise_at(noise, noise_std, rng, base + s);
    }
    // deltas                                              rendering_tcnn.py:93-100
    float delta = (s + 1 < S) ? __fsub_rn(zn, zi) : 1e10f;
    delta = __fmul_rn(delta, dnorm);
    const float r = fmaxf(__fadd_rn(sg, nz), 0.f);
    const float e = ok ? expf(-__fmul_rn(delta, r)) : 1.0f;
    const float alpha = ok ? __fsub_rn(1.0f, e) : 0.f;
    const float v = ok ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;   // rendering_tcnn.py:113-115
    const float incl = warp_incl_scan_prod(v, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float T = carry * excl;
    carry *= __shfl_sync(kFull, incl, 31);
    const float w = alpha * T;
    if (ok) {
      c.w[s] = w; c.T[s] = T; c.e[s] = e; c.r[s] = r;
      a_sum += w;
      z_sum += w * zi;
    }
  }
  A = warp_sum(a_sum);
  Z = warp_sum(z_sum);
}

// Backward over one ray given dL/dw_i through `gw(s)`; writes d_sigma, returns dL/d|d|.
template <class GW>
__device__ __forceinline__ float ray_backward(const float* __restrict__ z, int64_t base, int S, float dnorm,
                                              int lane, const RayCtx& c, GW gw, float* __restrict__ d_sigma) {
  float carry = 0.f, dnorm_acc = 0.f;
  const int nblk = (S + 31) / 32;
  for (int b = nblk - 1; b >= 0; --b) {
    const int s = b * 32 + lane;
    const bool ok = s < S;
    float g = 0.f, w = 0.f, T = 1.f, e = 1.f, r = 0.f, zi = 0.f, zn = 0.f;
    if (ok) {
      g = gw(s); w = c.w[s]; T = c.T[s]; e = c.e[s]; r = c.r[s];
      zi = z[base + s];
      zn = (s + 1 < S) ? z[base + s + 1] : 0.f;
    }
    const float gwv = g * w;
    const float incl = warp_incl_suffix_sum(gwv, lane);
    const float suffix_excl = carry + (incl - gwv);          // sum over samples after s
    carry += __shfl_sync(kFull, incl, 0);
    if (ok) {
      const float alpha = 1.0f - e;
      const float v = (1.0f - alpha) + 1e-10f;
      const float d_alpha = g * T - suffix_excl / v;
      const float dz = (s + 1 < S) ? (zn - zi) : 1e10f;
      const float delta = dz * dnorm;
      const float pos = r > 0.f ? 1.f : 0.f;
      d_sigma[base + s] = d_alpha * delta * e * pos;
      dnorm_acc += d_alpha * dz * r * e;
    }
  }
  return warp_sum(dnorm_acc);
}

__device__ __forceinline__ float kl_gauss(float m1, float s1, float m2, float s2) {
  // optimizer.py:614-621
  return logf(s2 / s1) + (s1 * s1 + (m1 - m2) * (m1 - m2)) / (2.f * s2 * s2) - 0.5f;
}

// mode 0: forward only; mode 1: fused loss forward + backward
template <int MODE>
__global__ void __launch_bounds__(kRenderWarps * 32)
render_kernel(const float* __restrict__ sigma, const float* __restrict__ z_vals, const float* __restrict__ rays,
              const float* __restrict__ depths, const uint8_t* __restrict__ flags, int64_t n, int S,
              const float* __restrict__ noise, float noise_std, uint64_t seed, const int32_t* __restrict__ counts,
              LossCfg cfg, float* __restrict__ loss_acc, float* __restrict__ weights, float* __restrict__ depth_out,
              float* __restrict__ opacity_out, float* __restrict__ variance_out, float* __restrict__ eps_out,
              float* __restrict__ d_sigma, float* __restrict__ d_rays) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kRenderWarps + warp;
  if (ray >= n) return;
  float* my = smem + (size_t)warp * 4 * S;
  RayCtx c{my, my + S, my + 2 * S, my + 3 * S};
  const int64_t base = ray * S;
  const float* R = rays + ray * LONER_RAY_COLS;
  const float dx = R[3], dy = R[4], dz = R[5], far = R[12];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);               // rendering_tcnn.py:100
  const Philox rng(seed);
  const unsigned fl = flags ? flags[ray] : (LONER_FLAG_VALID | LONER_FLAG_OPAQUE);
  const bool valid = fl & LONER_FLAG_VALID;

  if (!valid) {   // filtered-out row (ray_utils.py:321-322): contributes nothing
    for (int s = lane; s < S; s += 32) {
      if (weights) weights[base + s] = 0.f;
      if (MODE == 1) d_sigma[base + s] = 0.f;
    }
    if (lane == 0) {
      if (depth_out) depth_out[ray] = 0.f;
      if (opacity_out) opacity_out[ray] = 0.f;
      if (variance_out) variance_out[ray] = 0.f;
      if (eps_out) eps_out[ray] = 0.f;
    }
    return;
  }

  float A, Z;
  ray_forward(sigma, z_vals, noise, noise_std, rng, base, S, dnorm, lane, c, A, Z);
  __syncwarp();
  const float D = Z + (1.0f - A) * far;                                    // rendering_tcnn.py:125-129
  // variance (rendering_tcnn.py:143) and the JS statistics (optimizer.py:476-478)
  const float scale = cfg.scale;
  float ms = 0.f;
  for (int s = lane; s < S; s += 32) ms += (z_vals[base + s] * scale) * c.w[s];
  const float mean = warp_sum(ms) / (A + 1e-10f);
  float var_out = 0.f, var_js = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float zi = z_vals[base + s], w = c.w[s];
    var_out += w * (D - zi) * (D - zi);
    const float ds = zi * scale - mean;
    var_js += ds * ds * w;
  }
  var_out = warp_sum(var_out);
  var_js = warp_sum(var_js) / (A + 1e-10f) + 1e-10f;
  if (weights) for (int s = lane; s < S; s += 32) weights[base + s] = c.w[s];
  if (lane == 0) {
    if (depth_out) depth_out[ray] = D;
    if (opacity_out) opacity_out[ray] = A;
    if (variance_out) variance_out[ray] = var_out;
  }
  if (MODE == 0) return;

  // ---- loss (optimizer.py:460-591)
  const bool opaque = fl & LONER_FLAG_OPAQUE;
  const float G = depths[ray] * scale;
  const float stdv = sqrtf(var_js);
  const float s0 = cfg.eps_min / 3.0f;
  const float mm = 0.5f * (G + mean);
  const float sm = 0.5f * sqrtf(s0 * s0 + stdv * stdv);
  float js = 0.5f * kl_gauss(G, s0, mm, sm) + 0.5f * kl_gauss(mean, stdv, mm, sm);
  if (js < cfg.min_js) js = 0.f;
  if (js > cfg.max_js) js = cfg.max_js;
  const float eps = cfg.fixed_eps > 0.f ? cfg.fixed_eps : cfg.eps_min * (1.0f + cfg.alpha * js);
  if (eps_out && lane == 0) eps_out[ray] = eps;

  // target weights (losses.py:29-51), unnormalised sum first
  const float sg = eps / 3.0f;
  const float ca = __fdiv_rn(__fsub_rn(__fsub_rn(G, eps), G), sg);
  const float cb = __fdiv_rn(__fsub_rn(__fadd_rn(G, eps), G), sg);
  const float cdf_d = 0.5f * (1.0f + erff(cb * 0.70710678118654752f)) - 0.5f * (1.0f + erff(ca * 0.70710678118654752f));
  const float lo_edge = __fsub_rn(G, eps), hi_edge = __fadd_rn(G, eps);
  auto wgt_raw = [&](float s_m) -> float {
    if (!(s_m > lo_edge) || !(hi_edge > s_m)) return 0.f;
    const float x = (s_m - G) / sg;
    return 0.3989422804014327f * expf(-0.5f * (x * x)) / sg / cdf_d;
  };
  float wn = 0.f;
  if (opaque) for (int s = lane; s < S; s += 32) wn += wgt_raw(z_vals[base + s] * scale);
  wn = warp_sum(wn) + 1e-6f;

  const float n_valid = (float)counts[0], n_opaque = (float)counts[1];
  const float k_los = cfg.los_lambda / (n_valid * (float)S);
  const float derr = D * scale - G;
  const float k_depth = opaque ? cfg.depth_lambda * 2.0f * derr * scale / n_opaque : 0.f;
  const float k_opac = opaque ? ((A - 1.0f) > 0.f ? 1.f : ((A - 1.0f) < 0.f ? -1.f : 0.f)) / n_opaque : 0.f;
  float l1 = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float t = opaque ? wgt_raw(z_vals[base + s] * scale) / wn : 0.f;
    const float df = c.w[s] - t;
    l1 += cfg.l2 ? df * df : fabsf(df);
  }
  l1 = warp_sum(l1);
  if (lane == 0) {
    if (opaque) { atomicAdd(loss_acc + 0, derr * derr); atomicAdd(loss_acc + 2, fabsf(A - 1.0f)); }
    atomicAdd(loss_acc + 1, l1);
    atomicAdd(loss_acc + 3, eps);
  }
  auto gw = [&](int s) -> float {
    const float zi = z_vals[base + s];
    const float t = opaque ? wgt_raw(zi * scale) / wn : 0.f;
    const float df = c.w[s] - t;
    const float sgn = cfg.l2 ? 2.0f * df : (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
    return k_los * sgn + k_depth * (zi - far) + k_opac;
  };
  const float g_norm = ray_backward(z_vals, base, S, dnorm, lane, c, gw, d_sigma);
  if (d_rays && lane == 0) {
    float* g = d_rays + ray * LONER_RAY_COLS;
    const float k = g_norm / dnorm;
    g[3] += k * dx; g[4] += k * dy; g[5] += k * dz;
    g[12] += k_depth * (1.0f - A);   // depth = sum w z + (1 - A) * far   (rendering_tcnn.py:125-129)
  }
}

// ------------------------------------------------------------------------------------------
// Register variant of render_kernel for S = 32 * kSpl (kSpl = 4, 8, 12, 16: S = 128 ... 512, the training sizes).
// Lane i owns the kSpl CONTIGUOUS samples [i kSpl, (i+1) kSpl): one 16-byte load per 4 samples, the
// transmittance product and the backward suffix sums are sequential inside the lane plus ONE warp scan of the
// lane totals (the smem variant above walks S/32 dependent warp scans), every per-sample quantity (z, w, T, e, r,
// target weight) lives in registers and is computed once, and Philox is called once per PAIR of normals.
// Same arithmetic per sample; only the association of the running product / sums differs (fp32 rounding level).
template <int MODE, int kSpl>
__global__ void __launch_bounds__(kRenderWarps * 32)
render_reg_kernel(const float* __restrict__ sigma, const float* __restrict__ z_vals, const float* __restrict__ rays,
                  const float* __restrict__ depths, const uint8_t* __restrict__ flags, int64_t n,
                  const float* __restrict__ noise, float noise_std, uint64_t seed, const int32_t* __restrict__ counts,
                  LossCfg cfg, float* __restrict__ loss_acc, float* __restrict__ weights, float* __restrict__ depth_out,
                  float* __restrict__ opacity_out, float* __restrict__ variance_out, float* __restrict__ eps_out,
                  float* __restrict__ d_sigma, float* __restrict__ d_rays) {
  constexpr int S = 32 * kSpl;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kRenderWarps + warp;
  if (ray >= n) return;
  const int64_t base = ray * S + (int64_t)lane * kSpl;
  const float* R = rays + ray * LONER_RAY_COLS;
  const float dx = R[3], dy = R[4], dz = R[5], far = R[12];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);               // rendering_tcnn.py:100
  const unsigned fl = flags ? flags[ray] : (LONER_FLAG_VALID | LONER_FLAG_OPAQUE);
  auto store4 = [&](float* dst, const float (&v)[kSpl]) {
#pragma unroll
    for (int k = 0; k < kSpl; k += 4) *reinterpret_cast<float4*>(dst + base + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
  };
  if (!(fl & LONER_FLAG_VALID)) {   // filtered-out row (ray_utils.py:321-322): contributes nothing
    float zero[kSpl];
#pragma unroll
    for (int k = 0; k < kSpl; ++k) zero[k] = 0.f;
    if (weights) store4(weights, zero);
    if (MODE == 1) store4(d_sigma, zero);
    if (lane == 0) {
      if (depth_out) depth_out[ray] = 0.f;
      if (opacity_out) opacity_out[ray] = 0.f;
      if (variance_out) variance_out[ray] = 0.f;
      if (eps_out) eps_out[ray] = 0.f;
    }
    return;
  }
  float z[kSpl], r[kSpl], e[kSpl], w[kSpl], T[kSpl];
#pragma unroll
  for (int k = 0; k < kSpl; k += 4) {
    const float4 a = *reinterpret_cast<const float4*>(z_vals + base + k);
    const float4 b = *reinterpret_cast<const float4*>(sigma + base + k);
    z[k] = a.x; z[k + 1] = a.y; z[k + 2] = a.z; z[k + 3] = a.w;
    r[k] = b.x; r[k + 1] = b.y; r[k + 2] = b.z; r[k + 3] = b.w;       // sigma for now
  }
  if (noise_std > 0.f) {
    if (noise) {
#pragma unroll
      for (int k = 0; k < kSpl; k += 4) {
        const float4 c = *reinterpret_cast<const float4*>(noise + base + k);
        r[k] = __fadd_rn(r[k], c.x * noise_std); r[k + 1] = __fadd_rn(r[k + 1], c.y * noise_std);
        r[k + 2] = __fadd_rn(r[k + 2], c.z * noise_std); r[k + 3] = __fadd_rn(r[k + 3], c.w * noise_std);
      }
    } else {
      const Philox rng(seed);
#pragma unroll
      for (int k = 0; k < kSpl; k += 2) {           // base + k is even: one Philox call serves indices (2c, 2c + 1)
        const uint4 q = rng((uint64_t)((base + k) >> 1), 3u);
        const float2 g = box_muller(q.x, q.y);
        r[k] = __fadd_rn(r[k], g.x * noise_std);
        r[k + 1] = __fadd_rn(r[k + 1], g.y * noise_std);
      }
    }
  }
  const float z_next_lane = __shfl_down_sync(kFull, z[0], 1);
  // ---- forward                                        rendering_tcnn.py:93-129
  float prod = 1.0f;                                     // running product of (1 - alpha + 1e-10) inside the lane
#pragma unroll
  for (int k = 0; k < kSpl; ++k) {
    const float zn = (k + 1 < kSpl) ? z[k + 1] : z_next_lane;
    const bool last = (lane == 31 && k == kSpl - 1);
    float delta = last ? 1e10f : __fsub_rn(zn, z[k]);
    delta = __fmul_rn(delta, dnorm);
    r[k] = fmaxf(r[k], 0.f);
    e[k] = expf(-__fmul_rn(delta, r[k]));
    const float alpha = __fsub_rn(1.0f, e[k]);
    T[k] = prod;                                         // exclusive, inside the lane
    w[k] = alpha;
    prod *= __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);   // rendering_tcnn.py:113-115
  }
  {
    const float incl = warp_incl_scan_prod(prod, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.0f;
#pragma unroll
    for (int k = 0; k < kSpl; ++k) { T[k] *= excl; w[k] *= T[k]; }
  }
  float a_sum = 0.f, z_sum = 0.f;
#pragma unroll
  for (int k = 0; k < kSpl; ++k) { a_sum += w[k]; z_sum += w[k] * z[k]; }
  const float A = warp_sum(a_sum), Z = warp_sum(z_sum);
  const float D = Z + (1.0f - A) * far;                                    // rendering_tcnn.py:125-129
  const float scale = cfg.scale;
  float ms = 0.f;
#pragma unroll
  for (int k = 0; k < kSpl; ++k) ms += (z[k] * scale) * w[k];
  const float mean = warp_sum(ms) / (A + 1e-10f);
  float var_out = 0.f, var_js = 0.f;
#pragma unroll
  for (int k = 0; k < kSpl; ++k) {
    var_out += w[k] * (D - z[k]) * (D - z[k]);
    const float ds = z[k] * scale - mean;
    var_js += ds * ds * w[k];
  }
  var_out = warp_sum(var_out);
  var_js = warp_sum(var_js) / (A + 1e-10f) + 1e-10f;
  if (weights) store4(weights, w);
  if (lane == 0) {
    if (depth_out) depth_out[ray] = D;
    if (opacity_out) opacity_out[ray] = A;
    if (variance_out) variance_out[ray] = var_out;
  }
  if (MODE == 0) return;

  // ---- loss (optimizer.py:460-591)
  const bool opaque = fl & LONER_FLAG_OPAQUE;
  const float G = depths[ray] * scale;
  const float stdv = sqrtf(var_js);
  const float s0 = cfg.eps_min / 3.0f;
  const float mm = 0.5f * (G + mean);
  const float sm = 0.5f * sqrtf(s0 * s0 + stdv * stdv);
  float js = 0.5f * kl_gauss(G, s0, mm, sm) + 0.5f * kl_gauss(mean, stdv, mm, sm);
  if (js < cfg.min_js) js = 0.f;
  if (js > cfg.max_js) js = cfg.max_js;
  const float eps = cfg.fixed_eps > 0.f ? cfg.fixed_eps : cfg.eps_min * (1.0f + cfg.alpha * js);
  if (eps_out && lane == 0) eps_out[ray] = eps;
  // target weights (losses.py:29-51)
  const float sg = eps / 3.0f;
  const float ca = __fdiv_rn(__fsub_rn(__fsub_rn(G, eps), G), sg);
  const float cb = __fdiv_rn(__fsub_rn(__fadd_rn(G, eps), G), sg);
  const float cdf_d = 0.5f * (1.0f + erff(cb * 0.70710678118654752f)) - 0.5f * (1.0f + erff(ca * 0.70710678118654752f));
  const float lo_edge = __fsub_rn(G, eps), hi_edge = __fadd_rn(G, eps);
  float t[kSpl];
  float wn = 0.f;
#pragma unroll
  for (int k = 0; k < kSpl; ++k) {
    const float s_m = z[k] * scale;
    float v = 0.f;
    if (opaque && (s_m > lo_edge) && (hi_edge > s_m)) {
      const float x = (s_m - G) / sg;
      v = 0.3989422804014327f * expf(-0.5f * (x * x)) / sg / cdf_d;
    }
    t[k] = v;
    wn += v;
  }
  wn = warp_sum(wn) + 1e-6f;
  const float n_valid = (float)counts[0], n_opaque = (float)counts[1];
  const float k_los = cfg.los_lambda / (n_valid * (float)S);
  const float derr = D * scale - G;
  const float k_depth = opaque ? cfg.depth_lambda * 2.0f * derr * scale / n_opaque : 0.f;
  const float k_opac = opaque ? ((A - 1.0f) > 0.f ? 1.f : ((A - 1.0f) < 0.f ? -1.f : 0.f)) / n_opaque : 0.f;
  float l1 = 0.f;
  float gsum = 0.f;                       // sum over the lane's samples of g_k w_k
#pragma unroll
  for (int k = 0; k < kSpl; ++k) {
    const float df = w[k] - (opaque ? t[k] / wn : 0.f);
    l1 += cfg.l2 ? df * df : fabsf(df);
    const float sgn = cfg.l2 ? 2.0f * df : (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
    t[k] = k_los * sgn + k_depth * (z[k] - far) + k_opac;      // t now holds dL/dw_k
    gsum += t[k] * w[k];
  }
  l1 = warp_sum(l1);
  if (lane == 0) {
    if (opaque) { atomicAdd(loss_acc + 0, derr * derr); atomicAdd(loss_acc + 2, fabsf(A - 1.0f)); }
    atomicAdd(loss_acc + 1, l1);
    atomicAdd(loss_acc + 3, eps);
  }
  // ---- backward: d_alpha_k = g_k T_k - (sum_{j>k} g_j w_j) / (1 - alpha_k + 1e-10)
  float after = warp_incl_suffix_sum(gsum, lane) - gsum;       // samples of the lanes behind this one
  float dnorm_acc = 0.f;
#pragma unroll
  for (int k = kSpl - 1; k >= 0; --k) {
    const float alpha = 1.0f - e[k];
    const float v = (1.0f - alpha) + 1e-10f;
    const float d_alpha = t[k] * T[k] - after / v;
    after += t[k] * w[k];
    const float zn = (k + 1 < kSpl) ? z[k + 1] : z_next_lane;
    const bool last = (lane == 31 && k == kSpl - 1);
    const float dzk = last ? 1e10f : (zn - z[k]);
    const float pos = r[k] > 0.f ? 1.f : 0.f;
    dnorm_acc += d_alpha * dzk * r[k] * e[k];
    T[k] = d_alpha * (dzk * dnorm) * e[k] * pos;               // T now holds d_sigma_k
  }
  store4(d_sigma, T);
  const float g_norm = warp_sum(dnorm_acc);
  if (d_rays && lane == 0) {
    float* g = d_rays + ray * LONER_RAY_COLS;
    const float kk = g_norm / dnorm;
    g[3] += kk * dx; g[4] += kk * dy; g[5] += kk * dz;
    g[12] += k_depth * (1.0f - A);   // depth = sum w z + (1 - A) * far   (rendering_tcnn.py:125-129)
  }
}

// generic backward of raw2outputs given upstream gradients
__global__ void __launch_bounds__(kRenderWarps * 32)
render_bwd_kernel(const float* __restrict__ sigma, const float* __restrict__ z_vals, const float* __restrict__ rays,
                  int64_t n, int S, const float* __restrict__ noise, float noise_std, uint64_t seed,
                  const float* __restrict__ g_weights, const float* __restrict__ g_depth,
                  const float* __restrict__ g_opacity, const float* __restrict__ g_variance,
                  float* __restrict__ d_sigma, float* __restrict__ d_rays) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kRenderWarps + warp;
  if (ray >= n) return;
  float* my = smem + (size_t)warp * 4 * S;
  RayCtx c{my, my + S, my + 2 * S, my + 3 * S};
  const int64_t base = ray * S;
  const float* R = rays + ray * LONER_RAY_COLS;
  const float dx = R[3], dy = R[4], dz = R[5], far = R[12];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  const Philox rng(seed);
  float A, Z;
  ray_forward(sigma, z_vals, noise, noise_std, rng, base, S, dnorm, lane, c, A, Z);
  __syncwarp();
  const float D = Z + (1.0f - A) * far;
  const float gd = g_depth ? g_depth[ray] : 0.f;
  const float go = g_opacity ? g_opacity[ray] : 0.f;
  const float gv = g_variance ? g_variance[ray] : 0.f;
  float m1 = 0.f;   // sum_k 2 w_k (D - z_k)
  if (gv != 0.f) {
    for (int s = lane; s < S; s += 32) m1 += 2.f * c.w[s] * (D - z_vals[base + s]);
    m1 = warp_sum(m1);
  }
  const float gD = gd + gv * m1;     // total gradient reaching D
  auto gw = [&](int s) -> float {
    const float zi = z_vals[base + s];
    float g = g_weights ? g_weights[base + s] : 0.f;
    g += gD * (zi - far) + go + gv * (D - zi) * (D - zi);
    return g;
  };
  const float g_norm = ray_backward(z_vals, base, S, dnorm, lane, c, gw, d_sigma);
  if (d_rays && lane == 0) {
    float* g = d_rays + ray * LONER_RAY_COLS;
    const float k = g_norm / dnorm;
    g[3] += k * dx; g[4] += k * dy; g[5] += k * dz;
    g[12] += gD * (1.0f - A);         // d depth / d far
  }
}

}  // namespace loner

// S = 128 / 256 / 384 / 512 with 16-byte aligned rows run on the register variant
template <int MODE, class... Args>
static bool launch_render_reg(int S, unsigned blocks, cudaStream_t st, Args... args) {
  switch (S) {
    case 128: loner::render_reg_kernel<MODE, 4><<<blocks, loner::kRenderWarps * 32, 0, st>>>(args...); return true;
    case 256: loner::render_reg_kernel<MODE, 8><<<blocks, loner::kRenderWarps * 32, 0, st>>>(args...); return true;
    case 384: loner::render_reg_kernel<MODE, 12><<<blocks, loner::kRenderWarps * 32, 0, st>>>(args...); return true;
    case 512: loner::render_reg_kernel<MODE, 16><<<blocks, loner::kRenderWarps * 32, 0, st>>>(args...); return true;
    default: return false;
  }
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int render_smem(int S, size_t* out) {
  *out = (size_t)loner::kRenderWarps * 4 * S * sizeof(float);
  return *out <= 200 * 1024;
}

extern "C" int loner_render_fwd(const float* sigma, const float* z_vals, const float* rays, int64_t n, int32_t S,
                                const float* noise, float raw_noise_std, uint64_t seed, float* weights,
                                float* depth, float* opacity, float* variance, void* stream) {
  if (n == 0) return LONER_OK;
  if (!sigma || !z_vals || !rays || n < 0 || S < 2) return LONER_E_BAD_ARG;
  size_t smem;
  if (!render_smem(S, &smem)) return LONER_E_UNSUPPORTED;
  auto k = loner::render_kernel<0>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const unsigned blocks = (unsigned)((n + loner::kRenderWarps - 1) / loner::kRenderWarps);
  loner::LossCfg cfg{1.f, 0.5f, 1.f, 10.f, 1.f, 0.f, 0.f, 0, 0.f};
  const bool al = aligned16(sigma) && aligned16(z_vals) && aligned16(noise) && aligned16(weights);
  if (!(al && launch_render_reg<0>(S, blocks, (cudaStream_t)stream, sigma, z_vals, rays, (const float*)nullptr,
                                   (const uint8_t*)nullptr, n, noise, raw_noise_std, seed, (const int32_t*)nullptr, cfg,
                                   (float*)nullptr, weights, depth, opacity, variance, (float*)nullptr, (float*)nullptr,
                                   (float*)nullptr)))
  k<<<blocks, loner::kRenderWarps * 32, smem, (cudaStream_t)stream>>>(
      sigma, z_vals, rays, nullptr, nullptr, n, S, noise, raw_noise_std, seed, nullptr, cfg, nullptr, weights, depth,
      opacity, variance, nullptr, nullptr, nullptr);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_render_loss(const float* sigma, const float* z_vals, const float* rays, const float* depths,
                                 const uint8_t* flags, int64_t n, int32_t S, const float* noise,
                                 float raw_noise_std, uint64_t seed, const int32_t* counts,
                                 const float* loss_cfg9_host, float* loss_acc, float* weights, float* depth,
                                 float* opacity, float* variance, float* eps_dyn, float* d_sigma, float* d_rays,
                                 void* stream) {
  if (n == 0) return LONER_OK;
  if (!sigma || !z_vals || !rays || !depths || !counts || !loss_cfg9_host || !loss_acc || !d_sigma || n < 0 || S < 2)
    return LONER_E_BAD_ARG;
  size_t smem;
  if (!render_smem(S, &smem)) return LONER_E_UNSUPPORTED;
  auto k = loner::render_kernel<1>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const unsigned blocks = (unsigned)((n + loner::kRenderWarps - 1) / loner::kRenderWarps);
  const float* c = loss_cfg9_host;
  loner::LossCfg cfg{c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7] != 0.f ? 1 : 0, c[8]};
  const bool al = aligned16(sigma) && aligned16(z_vals) && aligned16(noise) && aligned16(weights) && aligned16(d_sigma);
  if (!(al && launch_render_reg<1>(S, blocks, (cudaStream_t)stream, sigma, z_vals, rays, depths, flags, n, noise, raw_noise_std,
                                   seed, counts, cfg, loss_acc, weights, depth, opacity, variance, eps_dyn, d_sigma, d_rays)))
  k<<<blocks, loner::kRenderWarps * 32, smem, (cudaStream_t)stream>>>(
      sigma, z_vals, rays, depths, flags, n, S, noise, raw_noise_std, seed, counts, cfg, loss_acc, weights, depth,
      opacity, variance, eps_dyn, d_sigma, d_rays);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_render_bwd(const float* sigma, const float* z_vals, const float* rays, int64_t n, int32_t S,
                                const float* noise, float raw_noise_std, uint64_t seed, const float* g_weights,
                                const float* g_depth, const float* g_opacity, const float* g_variance,
                                float* d_sigma, float* d_rays, void* stream) {
  if (n == 0) return LONER_OK;
  if (!sigma || !z_vals || !rays || !d_sigma || n < 0 || S < 2) return LONER_E_BAD_ARG;
  size_t smem;
  if (!render_smem(S, &smem)) return LONER_E_UNSUPPORTED;
  auto k = loner::render_bwd_kernel;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const unsigned blocks = (unsigned)((n + loner::kRenderWarps - 1) / loner::kRenderWarps);
  k<<<blocks, loner::kRenderWarps * 32, smem, (cudaStream_t)stream>>>(
      sigma, z_vals, rays, n, S, noise, raw_noise_std, seed, g_weights, g_depth, g_opacity, g_variance, d_sigma,
      d_rays);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

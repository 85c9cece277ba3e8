// Optimiser steps and the occupancy-grid update.
// Replaces torch.optim.Adam as constructed at /root/reference/src/mapping/optimizer.py:257-267
// (step at :376), torch.optim.SGD for the occupancy grid (:108-109, :606) and the grid half of
// Optimizer._step_occupancy_grid + get_logits_grad (/root/reference/src/mapping/optimizer.py:598-609,
// /root/reference/src/models/losses.py:54-62): the pseudo-gradient is scattered trilinearly
// (the adjoint of ATen grid_sampler_3d, align_corners=False, zeros padding) straight from
// rays + z_vals, so `points_fine` [N,S,3] never has to exist in HBM.
#include <climits>
#include "common.cuh"

namespace loner {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            int64_t n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float unscale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * unscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;       // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;     // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p[i] = p[i] - (lr / bc1) * (mi / denom);            // param.addcdiv_(exp_avg, denom, value=-step_size)
  }
}

__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ x, const float* __restrict__ g, int64_t n, float lr) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = x[i] - lr * g[i];
}

// ---- poses: 6-vector [t | axis-angle] -> R | t through the unit quaternion (pytorch3d.transforms.axis_angle_to_matrix,
// which tensor_to_transform calls), and its chain rule by forward-mode differentiation over the three rotation
// parameters: a value with three tangents per intermediate, so the derivative follows the forward arithmetic line by line.
struct D3 {
  float v, d[3];
};
__device__ __forceinline__ D3 d3c(float c) { return D3{c, {0.f, 0.f, 0.f}}; }
__device__ __forceinline__ D3 operator+(const D3& a, const D3& b) { return D3{a.v + b.v, {a.d[0] + b.d[0], a.d[1] + b.d[1], a.d[2] + b.d[2]}}; }
__device__ __forceinline__ D3 operator-(const D3& a, const D3& b) { return D3{a.v - b.v, {a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2]}}; }
__device__ __forceinline__ D3 operator*(const D3& a, const D3& b) {
  return D3{a.v * b.v, {a.d[0] * b.v + a.v * b.d[0], a.d[1] * b.v + a.v * b.d[1], a.d[2] * b.v + a.v * b.d[2]}};
}
__device__ __forceinline__ D3 operator/(const D3& a, const D3& b) {
  const float q = a.v / b.v, ib = 1.f / b.v;
  return D3{q, {(a.d[0] - q * b.d[0]) * ib, (a.d[1] - q * b.d[1]) * ib, (a.d[2] - q * b.d[2]) * ib}};
}
__device__ __forceinline__ D3 d3f(float v, float dv, const D3& x) {      // f(x) with f'(x) = dv
  return D3{v, {dv * x.d[0], dv * x.d[1], dv * x.d[2]}};
}
// R (row-major, 9 values with tangents) of the axis-angle vector aa
__device__ __forceinline__ void axis_angle_rotation(const float aa[3], D3 R[9]) {
  D3 x[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { x[i] = d3c(aa[i]); x[i].d[i] = 1.f; }
  const float n2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  const float an = sqrtf(n2);
  D3 angle = d3c(an);                                   // d|aa| / d aa = aa / |aa| (0 at the origin, like torch.norm)
  if (an > 0.f) { angle.d[0] = aa[0] / an; angle.d[1] = aa[1] / an; angle.d[2] = aa[2] / an; }
  const D3 half = angle * d3c(0.5f);
  D3 s_over_a;                                          // sin(angle / 2) / angle, Taylor below 1e-6
  if (fabsf(an) < 1e-6f) s_over_a = d3c(0.5f) - angle * angle * d3c(1.f / 48.f);
  else s_over_a = d3f(sinf(half.v), cosf(half.v), half) / angle;
  const D3 r = d3f(cosf(half.v), -sinf(half.v), half);
  const D3 i = x[0] * s_over_a, j = x[1] * s_over_a, k = x[2] * s_over_a;
  const D3 two_s = d3c(2.f) / (r * r + i * i + j * j + k * k);
  const D3 one = d3c(1.f);
  R[0] = one - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r);       R[2] = two_s * (i * k + j * r);
  R[3] = two_s * (i * j + k * r);       R[4] = one - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
  R[6] = two_s * (i * k - j * r);       R[7] = two_s * (j * k + i * r);       R[8] = one - two_s * (i * i + j * j);
}

__device__ __forceinline__ bool finite6(const float* p) {
  bool ok = true;
#pragma unroll
  for (int e = 0; e < 6; ++e) ok = ok && isfinite(p[e]);
  return ok;
}

__global__ void pose_matrices_kernel(const float* __restrict__ store, const int32_t* __restrict__ rows, int K,
                                     float sx, float sy, float sz, float scale, float* __restrict__ poses12,
                                     int32_t* __restrict__ status) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const float* p = store + (int64_t)rows[k] * 6;
  const float aa[3] = {p[3], p[4], p[5]};
  if (status != nullptr) {
    int bits = 0;
    if (!finite6(p)) bits |= LONER_STATUS_BAD_POSE;
    // the sensor origin (every ray origin of this keyframe) must lie inside the world cube [-1,1]^3
    const float ox = (p[0] + sx) / scale, oy = (p[1] + sy) / scale, oz = (p[2] + sz) / scale;
    if (fmaxf(fabsf(ox), fmaxf(fabsf(oy), fabsf(oz))) > 1.f) bits |= LONER_STATUS_ORIGIN_OUTSIDE;
    if (bits) atomicOr(status, bits);
  }
  D3 R[9];
  axis_angle_rotation(aa, R);
#pragma unroll
  for (int e = 0; e < 9; ++e) poses12[k * 12 + e] = R[e].v;
#pragma unroll
  for (int e = 0; e < 3; ++e) poses12[k * 12 + 9 + e] = p[e];
}

// d_poses12 -> the gradient of the 6-vector; rows flagged free take one torch.optim.Adam step with their own step count
__global__ void pose_step_kernel(float* __restrict__ store, const int32_t* __restrict__ rows, const uint8_t* __restrict__ free_rows,
                                 int K, const float* __restrict__ d_poses12, float* __restrict__ grad6,
                                 float* __restrict__ m, float* __restrict__ v, int32_t* __restrict__ steps, float lr,
                                 float b1, float b2, float eps, int apply, int32_t* __restrict__ status) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int64_t row = rows[k];
  float* p = store + row * 6;
  const float aa[3] = {p[3], p[4], p[5]};
  D3 R[9];
  axis_angle_rotation(aa, R);
  const float* g12 = d_poses12 + k * 12;
  float g[6] = {g12[9], g12[10], g12[11], 0.f, 0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 9; ++e) {
    g[3] += g12[e] * R[e].d[0]; g[4] += g12[e] * R[e].d[1]; g[5] += g12[e] * R[e].d[2];
  }
  const bool is_free = free_rows[row] != 0;
#pragma unroll
  for (int e = 0; e < 6; ++e) grad6[row * 6 + e] = is_free ? g[e] : 0.f;
  if (!apply || !is_free) return;
  if (!finite6(g)) {                        // the reference raises before optimizer.step(): the row keeps its pose
    if (status != nullptr) atomicOr(status, LONER_STATUS_BAD_POSE_GRAD);
    return;
  }
  const int t = steps[row] + 1;
  steps[row] = t;
  const float bc1 = 1.f - powf(b1, (float)t), bc2_sqrt = sqrtf(1.f - powf(b2, (float)t));
#pragma unroll
  for (int e = 0; e < 6; ++e) {
    const float mi = b1 * m[row * 6 + e] + (1.f - b1) * g[e];
    const float vi = b2 * v[row * 6 + e] + (1.f - b2) * g[e] * g[e];
    m[row * 6 + e] = mi; v[row * 6 + e] = vi;
    p[e] = p[e] - (lr / bc1) * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
  if (status != nullptr && !finite6(p)) atomicOr(status, LONER_STATUS_BAD_POSE);
}

// One warp per ray.  The samples of a ray are sorted, so consecutive samples fall into the same voxel
// cell for long runs (a cell is 2/V of the cube; S samples cover at most ~sqrt(3) V cells): every lane
// walks a contiguous run of ceil(S/32) samples and keeps the 8 corner sums of its current cell in
// registers, flushing them with 8 atomics only when the cell changes - instead of 8 atomics per sample
// that serialise on the same L2 addresses (round 1: 1.5 ms per launch at the C2 size).
constexpr int kOgmWarps = 8;
__global__ void __launch_bounds__(kOgmWarps * 32)
ogm_grad_kernel(const float* __restrict__ rays, const float* __restrict__ z_vals, const float* __restrict__ depths,
                const uint8_t* __restrict__ flags, int64_t n, int S, float scale, int V, float* __restrict__ d_grid) {
  extern __shared__ float zsh[];                          // [kOgmWarps][S + 32], run r skewed by r words (bank spread)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kOgmWarps + warp;
  if (ray >= n) return;
  if (flags && !(flags[ray] & LONER_FLAG_VALID)) return;   // dropped by build_lidar_rays (ray_utils.py:321-322)
  float* zr = zsh + (size_t)warp * (S + 32);
  const int chunk = (S + 31) / 32;
  for (int s = lane; s < S; s += 32) zr[s + s / chunk] = z_vals[ray * S + s];
  __syncwarp();
  const float* R = rays + ray * LONER_RAY_COLS;
  const float ox = R[0], oy = R[1], oz = R[2], dx = R[3], dy = R[4], dz = R[5];
  const float gt = __fmul_rn(depths[ray], scale);
  const float fV = (float)V;
  const int s0 = lane * chunk, s1 = min(s0 + chunk, S);
  int cx = INT_MIN, cy = 0, cz = 0;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  auto flush = [&]() {
    if (cx == INT_MIN) return;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int xi = cx + (c & 1), yi = cy + ((c >> 1) & 1), zi = cz + (c >> 2);
      if (acc[c] != 0.f && (unsigned)xi < (unsigned)V && (unsigned)yi < (unsigned)V && (unsigned)zi < (unsigned)V)
        atomicAdd(d_grid + ((int64_t)zi * V + yi) * V + xi, acc[c]);
      acc[c] = 0.f;
    }
  };
  for (int s = s0; s < s1; ++s) {
    const float z = zr[s + lane];
    // get_logits_grad: x = s - gt in metres; +0.25 for x < -2, -2.5 for -2 < x < 2   (losses.py:54-62)
    const float x = __fmul_rn(z, scale) - gt;
    float g = 0.f;
    if (-x - 2.f > 0.f) g = 0.25f;
    else if (x + 2.f > 0.f && 2.f - x > 0.f) g = -2.5f;
    if (g == 0.f) continue;
    const float px = __fadd_rn(ox, __fmul_rn(dx, z)), py = __fadd_rn(oy, __fmul_rn(dy, z)),
                pz = __fadd_rn(oz, __fmul_rn(dz, z));
    const float ix = (__fmul_rn(__fadd_rn(px, 1.f), fV) - 1.f) * 0.5f;
    const float iy = (__fmul_rn(__fadd_rn(py, 1.f), fV) - 1.f) * 0.5f;
    const float iz = (__fmul_rn(__fadd_rn(pz, 1.f), fV) - 1.f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
    if (x0 != cx || y0 != cy || z0 != cz) { flush(); cx = x0; cy = y0; cz = z0; }
    const float wx1 = ix - fx, wy1 = iy - fy, wz1 = iz - fz;
    const float wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy, wz0 = (fz + 1.f) - iz;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      acc[c] += g * (((c & 1) ? wx1 : wx0) * (((c >> 1) & 1) ? wy1 : wy0) * ((c >> 2) ? wz1 : wz0));
  }
  flush();
}

// loss = depth_lambda * mean_opaque(sq depth err) + los_lambda * mean_{valid rays, samples}|w - w_gt| + mean_opaque|A - 1|
// (optimizer.py:486-491, :568-580) from the sums the loss kernel accumulated, and the mean dynamic margin
// `_depth_eps` (optimizer.py:503): one launch instead of a dozen 1-element ATen kernels per step.
__global__ void loss_finalize_kernel(const float* __restrict__ acc, const int32_t* __restrict__ counts, float depth_lambda,
                                     float los_lambda, int S, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float n_valid = (float)counts[0], n_opaque = (float)counts[1];
  const float depth_loss = acc[0] / n_opaque, los = acc[1] / (n_valid * (float)S), opac = acc[2] / n_opaque;
  out[0] = depth_lambda * depth_loss + los_lambda * los + opac;
  out[1] = acc[3] / n_valid;
  out[2] = depth_loss; out[3] = los; out[4] = opac; out[5] = n_valid;
}

}  // namespace loner

extern "C" int loner_loss_finalize(const float* loss_acc, const int32_t* counts, float depthloss_lambda, float los_lambda,
                                   int32_t S, float* out6, void* stream) {
  if (!loss_acc || !counts || !out6 || S <= 0) return LONER_E_BAD_ARG;
  loner::loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(loss_acc, counts, depthloss_lambda, los_lambda, S, out6);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t count,
                               int32_t step, float lr, float beta1, float beta2, float eps, float grad_unscale,
                               void* stream) {
  if (count == 0) return LONER_OK;
  if (!params || !grads || !exp_avg || !exp_avg_sq || count < 0 || step < 1) return LONER_E_BAD_ARG;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  unsigned blocks = (unsigned)((count + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  loner::adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, count, lr, beta1,
                                                              beta2, eps, (float)bc1, (float)sqrt(bc2), grad_unscale);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_pose_matrices(const float* poses6, const int32_t* rows, int32_t K, const float* shift, float scale,
                                   float* poses12, int32_t* status, void* stream) {
  if (K == 0) return LONER_OK;
  if (!poses6 || !rows || !poses12 || K < 0 || (status && (!shift || !(scale > 0.f)))) return LONER_E_BAD_ARG;
  const float sx = shift ? shift[0] : 0.f, sy = shift ? shift[1] : 0.f, sz = shift ? shift[2] : 0.f;
  loner::pose_matrices_kernel<<<(K + 63) / 64, 64, 0, (cudaStream_t)stream>>>(poses6, rows, K, sx, sy, sz, scale, poses12,
                                                                              status);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_pose_step(float* poses6, const int32_t* rows, const uint8_t* free_rows, int32_t K,
                               const float* d_poses12, float* grad6, float* exp_avg, float* exp_avg_sq, int32_t* steps,
                               float lr, float beta1, float beta2, float eps, int32_t apply, int32_t* status,
                               void* stream) {
  if (K == 0) return LONER_OK;
  if (!poses6 || !rows || !free_rows || !d_poses12 || !grad6 || K < 0) return LONER_E_BAD_ARG;
  if (apply && (!exp_avg || !exp_avg_sq || !steps)) return LONER_E_BAD_ARG;
  loner::pose_step_kernel<<<(K + 63) / 64, 64, 0, (cudaStream_t)stream>>>(poses6, rows, free_rows, K, d_poses12, grad6,
                                                                          exp_avg, exp_avg_sq, steps, lr, beta1, beta2, eps,
                                                                          apply, status);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_sgd_step(float* x, const float* g, int64_t count, float lr, void* stream) {
  if (count == 0) return LONER_OK;
  if (!x || !g || count < 0) return LONER_E_BAD_ARG;
  unsigned blocks = (unsigned)((count + 255) / 256);
  if (blocks > 2368) blocks = 2368;
  loner::sgd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, g, count, lr);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_ogm_grad(const float* rays, const float* z_vals, const float* depths, const uint8_t* flags,
                              int64_t n, int32_t S, float scale, int32_t V, float* d_grid, void* stream) {
  if (n == 0) return LONER_OK;
  if (!rays || !z_vals || !depths || !d_grid || n < 0 || S <= 0 || V <= 0) return LONER_E_BAD_ARG;
  const size_t smem = (size_t)loner::kOgmWarps * (S + 32) * sizeof(float);
  if (smem > 200 * 1024) return LONER_E_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(loner::ogm_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const unsigned blocks = (unsigned)((n + loner::kOgmWarps - 1) / loner::kOgmWarps);
  loner::ogm_grad_kernel<<<blocks, loner::kOgmWarps * 32, smem, (cudaStream_t)stream>>>(rays, z_vals, depths, flags, n, S,
                                                                                       scale, V, d_grid);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_version(void) { return 100; }

extern "C" int loner_sm_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}

extern "C" const char* loner_error_string(int code) {
  switch (code) {
    case LONER_OK: return "ok";
    case LONER_E_BAD_ARG: return "bad argument (null pointer, negative size)";
    case LONER_E_UNSUPPORTED: return "configuration not supported by the sm_100a kernels";
    case LONER_E_LAUNCH: return "CUDA launch failure";
    case LONER_E_ARCH: return "device is not sm_100";
    default: return "unknown error";
  }
}

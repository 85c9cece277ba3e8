// Ray construction from the device-resident keyframe store.
// Replaces LidarRayDirections.build_lidar_rays + get_far_val
// (/root/reference/src/common/ray_utils.py:269-322, :31-60) and its autograd backward.
//
// HBM layout: every LiDAR return is ONE float4 (dx,dy,dz,dist) so a randomly picked ray costs a
// single 16-byte load instead of four scattered 4-byte loads from the reference's SoA buffers
// (sensors.py:57-82).  Output rows are the reference's 13-column rows; rows are not compacted,
// validity travels in `flags`.
#include "common.cuh"

namespace loner {

__global__ void __launch_bounds__(256)
ray_build_kernel(const float4* __restrict__ points, const int32_t* __restrict__ ray_kf,
                 const int64_t* __restrict__ ray_point, int64_t n, const float* __restrict__ poses,
                 float sx, float sy, float sz, float scale, float r0, float r1,
                 float* __restrict__ rays, float* __restrict__ depths, uint8_t* __restrict__ flags,
                 int32_t* __restrict__ counters) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int valid = 0, opaque = 0;
  if (i < n) {
    const int kf = ray_kf[i] & LONER_KF_MASK;
    const float4 p = __ldg(points + ray_point[i]);
    const float* P = poses + (int64_t)kf * 12;
    // origin = (t + shift) / scale                         ray_utils.py:282-284
    const float ox = __fdiv_rn(__fadd_rn(P[9], sx), scale);
    const float oy = __fdiv_rn(__fadd_rn(P[10], sy), scale);
    const float oz = __fdiv_rn(__fadd_rn(P[11], sz), scale);
    // d = normalize(R v)                                   ray_utils.py:293-297
    float ux = P[0] * p.x + P[1] * p.y + P[2] * p.z;
    float uy = P[3] * p.x + P[4] * p.y + P[5] * p.z;
    float uz = P[6] * p.x + P[7] * p.y + P[8] * p.z;
    const float nrm = sqrtf(ux * ux + uy * uy + uz * uz);
    const float dx = __fdiv_rn(ux, nrm), dy = __fdiv_rn(uy, nrm), dz = __fdiv_rn(uz, nrm);
    const float near = __fdiv_rn(r0, scale);
    const float far_range = __fdiv_rn(r1, scale);
    // get_far_val(no_nan=True)                             ray_utils.py:31-60
    float far_clip = 3.0e38f;
    {
      const float o3[3] = {ox, oy, oz};
      const float d3[3] = {dx, dy, dz};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float dd = __fadd_rn(d3[a], 1e-15f);
        const float t0 = fmaxf(__fdiv_rn(__fsub_rn(-1.0f, o3[a]), dd), 0.0f);
        const float t1 = fmaxf(__fdiv_rn(__fsub_rn(1.0f, o3[a]), dd), 0.0f);
        far_clip = fminf(far_clip, fmaxf(t0, t1));
      }
    }
    const float far = fminf(far_range, far_clip);
    const float depth = __fdiv_rn(p.w, scale);
    float* r = rays + i * LONER_RAY_COLS;
    r[0] = ox; r[1] = oy; r[2] = oz;
    r[3] = dx; r[4] = dy; r[5] = dz;
    r[6] = -dx; r[7] = -dy; r[8] = -dz;
    r[9] = 0.f; r[10] = 0.f; r[11] = near; r[12] = far;
    depths[i] = depth;
    valid = far > __fadd_rn(near, __fdiv_rn(1.0f, scale));          // ray_utils.py:321
    opaque = valid && (depth > 0.0f) && !(depth > far);              // optimizer.py:460-463
    flags[i] = (uint8_t)((valid ? LONER_FLAG_VALID : 0u) | (opaque ? LONER_FLAG_OPAQUE : 0u));
  }
  if (counters != nullptr) {
    const unsigned bv = __ballot_sync(kFull, valid), bo = __ballot_sync(kFull, opaque);
    if ((threadIdx.x & 31) == 0) {
      if (bv) atomicAdd(counters + 0, __popc(bv));
      if (bo) atomicAdd(counters + 1, __popc(bo));
    }
  }
}

// Ray pick (optimizer.py:286-305): every output ray belongs to one segment = (keyframe, lidar | sky);
// RANDOM draws torch.randint(size) with Philox, MASK draws among the scan mask's index list, FIXED is
// arange.  One launch for the whole window instead of ~25 ATen micro-kernels.
__global__ void __launch_bounds__(256)
ray_pick_kernel(const loner_pick_seg_t* __restrict__ segs, int n_segs, const int64_t* __restrict__ index_map,
                uint64_t seed, int64_t n, int32_t* __restrict__ ray_kf, int64_t* __restrict__ ray_point) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = 0;
  while (s + 1 < n_segs && i >= segs[s + 1].out_begin) ++s;
  const loner_pick_seg_t sg = segs[s];
  const int64_t j = i - sg.out_begin;
  int64_t pick;
  if (sg.mode == LONER_PICK_FIXED) {
    pick = j;
  } else {
    const uint4 q = Philox(seed)((uint64_t)i, 4u);
    const uint64_t r = ((uint64_t)q.x << 32) | q.y;
    pick = (int64_t)__umul64hi(r, (uint64_t)sg.size);                 // uniform over [0, size)
  }
  if (sg.mode == LONER_PICK_MASK) pick = index_map[sg.map_off + pick];
  ray_kf[i] = sg.kf;
  ray_point[i] = sg.base + pick;
}

// Backward w.r.t. the pose (R,t): do -> dt / scale; dd -> through the normalisation -> dR.
// One block reduces its rays per keyframe in shared memory, then one atomic per (kf, entry).
__global__ void __launch_bounds__(256)
ray_build_bwd_kernel(const float4* __restrict__ points, const int32_t* __restrict__ ray_kf,
                     const int64_t* __restrict__ ray_point, int64_t n, const float* __restrict__ poses,
                     int K, float sx, float sy, float sz, float scale, float r1,
                     const float* __restrict__ d_rays, float* __restrict__ d_poses) {
  extern __shared__ float acc[];  // [K*12]
  for (int j = threadIdx.x; j < K * 12; j += blockDim.x) acc[j] = 0.f;
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // sky rays are built from the DETACHED pose (keyframe.py:93-95): no pose gradient through them
  if (i < n && !(ray_kf[i] & LONER_KF_DETACHED)) {
    const int kf = ray_kf[i] & LONER_KF_MASK;
    const float4 p = __ldg(points + ray_point[i]);
    const float* P = poses + (int64_t)kf * 12;
    const float* g = d_rays + i * LONER_RAY_COLS;
    const float ux = P[0] * p.x + P[1] * p.y + P[2] * p.z;
    const float uy = P[3] * p.x + P[4] * p.y + P[5] * p.z;
    const float uz = P[6] * p.x + P[7] * p.y + P[8] * p.z;
    const float inv = rsqrtf(ux * ux + uy * uy + uz * uz);
    const float dx = ux * inv, dy = uy * inv, dz = uz * inv;
    const float gd = g[3] * dx + g[4] * dy + g[5] * dz;
    const float gux = (g[3] - dx * gd) * inv, guy = (g[4] - dy * gd) * inv, guz = (g[5] - dz * gd) * inv;
    float* a = acc + kf * 12;
    atomicAdd(a + 0, gux * p.x); atomicAdd(a + 1, gux * p.y); atomicAdd(a + 2, gux * p.z);
    atomicAdd(a + 3, guy * p.x); atomicAdd(a + 4, guy * p.y); atomicAdd(a + 5, guy * p.z);
    atomicAdd(a + 6, guz * p.x); atomicAdd(a + 7, guz * p.y); atomicAdd(a + 8, guz * p.z);
    float gox = g[0], goy = g[1], goz = g[2];
    float gdx = 0.f, gdy = 0.f, gdz = 0.f;   // extra direction gradient through far (below)
    // far = min(r1/scale, cube exit): when the cube clips, far depends on (o, d)  ray_utils.py:306-311
    const float gfar = g[12];
    if (gfar != 0.f) {
      const float o3[3] = {(P[9] + sx) / scale, (P[10] + sy) / scale, (P[11] + sz) / scale};
      const float d3[3] = {dx, dy, dz};
      float best = 3.0e38f; int ba = 0;
      for (int ax = 0; ax < 3; ++ax) {
        const float dd = d3[ax] + 1e-15f;
        const float t0 = fmaxf((-1.f - o3[ax]) / dd, 0.f), t1 = fmaxf((1.f - o3[ax]) / dd, 0.f);
        const float t = fmaxf(t0, t1);
        if (t < best) { best = t; ba = ax; }
      }
      if (best < r1 / scale && best > 0.f) {
        const float dd = d3[ba] + 1e-15f;
        const float go = -gfar / dd, gd2 = -gfar * best / dd;
        if (ba == 0) { gox += go; gdx += gd2; } else if (ba == 1) { goy += go; gdy += gd2; } else { goz += go; gdz += gd2; }
      }
    }
    if (gdx != 0.f || gdy != 0.f || gdz != 0.f) {
      const float gd3 = gdx * dx + gdy * dy + gdz * dz;
      const float hx = (gdx - dx * gd3) * inv, hy = (gdy - dy * gd3) * inv, hz = (gdz - dz * gd3) * inv;
      atomicAdd(a + 0, hx * p.x); atomicAdd(a + 1, hx * p.y); atomicAdd(a + 2, hx * p.z);
      atomicAdd(a + 3, hy * p.x); atomicAdd(a + 4, hy * p.y); atomicAdd(a + 5, hy * p.z);
      atomicAdd(a + 6, hz * p.x); atomicAdd(a + 7, hz * p.y); atomicAdd(a + 8, hz * p.z);
    }
    atomicAdd(a + 9, gox / scale); atomicAdd(a + 10, goy / scale); atomicAdd(a + 11, goz / scale);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < K * 12; j += blockDim.x)
    if (acc[j] != 0.f) atomicAdd(d_poses + j, acc[j]);
}

// d_pos [n,S,3] -> d_rays: origin += sum_s g, direction += sum_s z*g   (xyz = o + d*z,
// rendering_tcnn.py:241).  One warp per ray.
__global__ void __launch_bounds__(256)
points_bwd_kernel(const float* __restrict__ d_pos, const float* __restrict__ z_vals, int64_t n, int S,
                  float* __restrict__ d_rays) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= n) return;
  float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0;
  for (int s = lane; s < S; s += 32) {
    const float* g = d_pos + (ray * S + s) * 3;
    const float z = z_vals[ray * S + s];
    ox += g[0]; oy += g[1]; oz += g[2];
    dx += z * g[0]; dy += z * g[1]; dz += z * g[2];
  }
  ox = warp_sum(ox); oy = warp_sum(oy); oz = warp_sum(oz);
  dx = warp_sum(dx); dy = warp_sum(dy); dz = warp_sum(dz);
  if (lane == 0) {
    float* r = d_rays + ray * LONER_RAY_COLS;
    r[0] += ox; r[1] += oy; r[2] += oz; r[3] += dx; r[4] += dy; r[5] += dz;
  }
}

}  // namespace loner

extern "C" int loner_ray_build(const void* points, const int32_t* ray_kf, const int64_t* ray_point, int64_t n,
                               const float* poses, int32_t K, const float* shift3_host, float scale, float r0,
                               float r1, float* rays, float* depths, uint8_t* flags, int32_t* counters,
                               void* stream) {
  if (n == 0) return LONER_OK;
  if (!points || !ray_kf || !ray_point || !poses || !shift3_host || !rays || !depths || !flags || n < 0 || K <= 0)
    return LONER_E_BAD_ARG;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  loner::ray_build_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
      (const float4*)points, ray_kf, ray_point, n, poses, shift3_host[0], shift3_host[1], shift3_host[2], scale,
      r0, r1, rays, depths, flags, counters);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_ray_pick(const loner_pick_seg_t* segs, int32_t n_segs, const int64_t* index_map, uint64_t seed,
                              int64_t n, int32_t* ray_kf, int64_t* ray_point, void* stream) {
  if (n == 0) return LONER_OK;
  if (!segs || n_segs <= 0 || !ray_kf || !ray_point || n < 0) return LONER_E_BAD_ARG;
  loner::ray_pick_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(segs, n_segs, index_map, seed, n,
                                                                                    ray_kf, ray_point);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_ray_build_bwd(const void* points, const int32_t* ray_kf, const int64_t* ray_point, int64_t n,
                                   const float* poses, int32_t K, const float* shift3_host, float scale,
                                   float r1, const float* d_rays, float* d_poses, void* stream) {
  if (n == 0) return LONER_OK;
  if (!points || !ray_kf || !ray_point || !poses || !shift3_host || !d_rays || !d_poses || n < 0 || K <= 0 || K > 1024)
    return LONER_E_BAD_ARG;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  loner::ray_build_bwd_kernel<<<blocks, threads, K * 12 * sizeof(float), (cudaStream_t)stream>>>(
      (const float4*)points, ray_kf, ray_point, n, poses, K, shift3_host[0], shift3_host[1], shift3_host[2], scale,
      r1, d_rays, d_poses);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

extern "C" int loner_points_bwd(const float* d_pos, const float* z_vals, int64_t n, int32_t S, float* d_rays,
                                void* stream) {
  if (n == 0) return LONER_OK;
  if (!d_pos || !z_vals || !d_rays || n < 0 || S <= 0) return LONER_E_BAD_ARG;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n * 32 + threads - 1) / threads);
  loner::points_bwd_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(d_pos, z_vals, n, S, d_rays);
  LONER_CHECK_LAUNCH();
  return LONER_OK;
}

/* loner_b200 — C ABI of the B200-native LONER mapping hot path.
 *
 * The reference (umautobots/LONER) has no FFI: its seam is Python module attributes
 * (SURVEY.md 8b).  This header is the boundary a maintainer binds instead (ctypes stub in
 * INTEGRATION.md); every entry point names the reference function it replaces.
 *
 * Conventions: every pointer is a DEVICE pointer unless its name ends in _host; the library
 * never allocates, never synchronises the device, keeps no global state, and launches on the
 * `stream` (a cudaStream_t passed as void*) the caller gives.  Return value: 0 = ok, otherwise
 * a LONER_E_* code (loner_error_string() names it).  All tensors are contiguous fp32 unless
 * stated.  Ray rows follow the reference layout (common/ray_utils.py:313-315):
 *   [0:3] origin, [3:6] unit direction, [6:9] view direction, [9:11] zero, [11] near, [12] far.
 */
#ifndef LONER_B200_H
#define LONER_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
  LONER_OK = 0,
  LONER_E_BAD_ARG = 1,     /* null pointer / non-positive size / unsupported shape */
  LONER_E_UNSUPPORTED = 2, /* configuration outside what the kernels implement      */
  LONER_E_LAUNCH = 3,      /* cudaGetLastError() after launch was not cudaSuccess    */
  LONER_E_ARCH = 4         /* device is not sm_100                                   */
};

#define LONER_RAY_COLS 13
#define LONER_FLAG_VALID 1u   /* far > near + 1/scale           (ray_utils.py:321) */
#define LONER_FLAG_OPAQUE 2u  /* depth > 0 and not depth > far  (optimizer.py:460-463) */
/* ray_kf[i] = keyframe row of `poses`, optionally OR-ed with LONER_KF_DETACHED: the ray is built from the
 * detached pose (sky rays, mapping/keyframe.py:93-95) and loner_ray_build_bwd skips it. */
#define LONER_KF_DETACHED 0x40000000
#define LONER_KF_MASK 0x3FFFFFFF

int loner_version(void);
int loner_sm_arch(void); /* compute capability of the current device *10 + minor, e.g. 100 */
const char* loner_error_string(int code);

/* ---- a2  ray pick of the optimisation loop (mapping/optimizer.py:286-305): per keyframe
 * `torch.randint(len(scan), (n,))` (RANDOM), picks among `scan.mask.nonzero()` (MASK), `arange(n)`
 * (FIXED), and `num_samples.sky` picks among the keyframe's sky directions.  The output rays are
 * a concatenation of segments; segment s fills rays [out_begin, next segment's out_begin). */
enum { LONER_PICK_RANDOM = 0, LONER_PICK_FIXED = 1, LONER_PICK_MASK = 2 };
typedef struct {
  int32_t kf;        /* value written to ray_kf (keyframe row, | LONER_KF_DETACHED for sky segments) */
  int32_t mode;      /* LONER_PICK_*                                                               */
  int64_t base;      /* index in `points` of the segment's first candidate                         */
  int64_t size;      /* number of candidates (points of the scan / sky set, or mask entries)       */
  int64_t map_off;   /* MASK: offset of the scan's mask index list inside `index_map`              */
  int64_t out_begin; /* first output ray of the segment                                            */
} loner_pick_seg_t;
/* segs: DEVICE array [n_segs] sorted by out_begin; index_map: device int64 (may be NULL without MASK);
 * draws are Philox(seed, ray index).  Writes ray_kf [n] and ray_point [n] for loner_ray_build. */
int loner_ray_pick(const loner_pick_seg_t* segs, int32_t n_segs, const int64_t* index_map, uint64_t seed,
                   int64_t n, int32_t* ray_kf, int64_t* ray_point, void* stream);

/* ---- a3/a4  LidarRayDirections.build_lidar_rays + get_far_val (common/ray_utils.py:269-322,
 * :31-60).  points: keyframe store, one float4 (dx,dy,dz,dist[m]) per LiDAR return, all
 * keyframes concatenated; ray_kf/ray_point: per output ray the keyframe id and the absolute
 * index into `points`.  poses: [K,12] = 3x3 R row-major then t.  Rows are NOT compacted:
 * flags[i] carries LONER_FLAG_VALID/OPAQUE and counters[0..1] += (#valid, #opaque). */
int loner_ray_build(const void* points, const int32_t* ray_kf, const int64_t* ray_point, int64_t n,
                    const float* poses, int32_t K, const float* shift3_host, float scale, float r0, float r1,
                    float* rays, float* depths, uint8_t* flags, int32_t* counters, void* stream);

/* backward of the above w.r.t. poses: d_rays [n,13] (columns 0..5 and 12 = far are read) ->
 * d_poses [K,12] (accumulated with atomics; caller zeroes). */
int loner_ray_build_bwd(const void* points, const int32_t* ray_kf, const int64_t* ray_point, int64_t n,
                        const float* poses, int32_t K, const float* shift3_host, float scale, float r1,
                        const float* d_rays, float* d_poses, void* stream);

/* ---- a6  UniformRaySampler.get_samples (models/ray_sampling.py:22-43).  u: [n,S] uniforms in
 * [0,1) or NULL (then Philox(seed) draws them).  z_vals [n,S], ascending. */
int loner_sample_uniform(const float* rays, int64_t n, int32_t S, float perturb, const float* u,
                         uint64_t seed, float* z_vals, void* stream);

/* ---- a7/a8/a9  OccGridRaySampler.get_samples + OccupancyGridModel.interpolate + sample_pdf
 * (models/ray_sampling.py:53-92, model_tcnn.py:124-131, rendering_tcnn.py:18-67).
 * grid: [V,V,V] logits indexed [z][y][x]; u1,u2: [n,S/2] or NULL. */
int loner_sample_ogm(const float* rays, int64_t n, int32_t S, float perturb, const float* grid, int32_t V,
                     const float* u1, const float* u2, uint64_t seed, float* z_vals, void* stream);

/* ---- a12  sigma head: Frequency encoding + bias-free ReLU MLP (models/nerf_tcnn.py:35-38,
 * :59-78).  Network description. */
typedef struct {
  int32_t n_frequencies;   /* Frequency encoding, 3 input dims -> 6F features, padded to 16 with 1.0 */
  int32_t n_neurons;       /* hidden width W: 64 (run zero-padded on the 128-wide kernels), 128 or 256 */
  int32_t n_hidden_layers; /* L >= 1 hidden layers                                             */
  int32_t flags;           /* 0 = production kernels; LONER_NET_* bits select the A/B variants */
} loner_net_t;
/* forward / dgrad as single CTAs (cta_group::1) instead of CTA pairs sharing the weight operand (cta_group::2) */
#define LONER_NET_SINGLE_CTA 1
/* dgrad stashes dZ_L and wgrad reads it, instead of wgrad rebuilding it from the ReLU masks, d_sigma and w_out */
#define LONER_NET_STASH_DZL 2
/* training forward with ONE MMA-issuing warp for both tiles (and a weight stream per tile) instead of one issuing warp per
 * tile sharing one weight stream */
#define LONER_NET_ONE_ISSUER 32
/* dgrad alone with one MMA-issuing warp (A/B of the dgrad pipeline) */
#define LONER_NET_DGRAD_ONE_ISSUER 64
/* the forward stashes A_L too and wgrad accumulates dW_out from it, instead of deriving dW_out from the last hidden
 * layer's weight-gradient partials (no biases: A_L = mask_L * (A_{L-1} W_{L-1}^T)) */
#define LONER_NET_STASH_AL 128
/* wgrad gives layer 0 a share of the CTAs in proportion to the bytes it streams, instead of bytes per byte in flight */
#define LONER_NET_WG_PLAN_BYTES 256

int64_t loner_mlp_param_count(const loner_net_t* net);       /* flat fp32 params, [out,in] row-major per layer */
int64_t loner_mlp_packed_bytes(const loner_net_t* net);      /* fp16 tensor-core image of the params */
int64_t loner_mlp_act_bytes(const loner_net_t* net, int64_t P);   /* activation stash for backward */
int64_t loner_mlp_bwd_scratch_bytes(const loner_net_t* net, int64_t P); /* dZ stash + wgrad partial sums */

/* fp32 master params -> fp16 swizzled tile image read by the tensor-core kernels */
int loner_mlp_pack(const loner_net_t* net, const float* params, void* packed, void* stream);

/* forward.  Either pos [P,3] in [-1,1] (DecoupledNeRF.forward API) or, when pos == NULL,
 * rays [n,13] + z_vals [n,S] with P = n*S (points o + d*z are formed in registers, never stored).
 * sigma [P] fp32.  acts: NULL (inference) or the activation stash consumed by loner_mlp_bwd (opaque: what it holds
 * depends on net->flags - e.g. A_L only with LONER_NET_STASH_AL - so forward and backward take the same loner_net_t). */
int loner_mlp_fwd(const loner_net_t* net, const void* packed, const float* pos, const float* rays,
                  const float* z_vals, int32_t S, int64_t P, float* sigma, void* acts, void* stream);

/* backward.  d_sigma [P] (16-byte aligned) -> d_params [param_count] (+=, caller zeroes) and, if d_pos != NULL,
 * d_pos [P,3] (gradient w.r.t. the [-1,1] positions).  grad_scale: power-of-two loss scale
 * applied to the fp16 intermediate gradients (undone before d_params / d_pos are written). */
int loner_mlp_bwd(const loner_net_t* net, const void* packed, const float* pos, const float* rays,
                  const float* z_vals, int32_t S, int64_t P, const float* d_sigma, const void* acts,
                  float grad_scale, float* d_params, float* d_pos, void* scratch, void* stream);

/* the two halves of loner_mlp_bwd, separately launchable (dgrad must run first: it writes the
 * loss-scaled fp16 layer gradients into `scratch` that wgrad contracts with the stashed activations) */
int loner_mlp_dgrad(const loner_net_t* net, const void* packed, const float* pos, const float* rays,
                    const float* z_vals, int32_t S, int64_t P, const float* d_sigma, const void* acts,
                    float grad_scale, float* d_pos, void* scratch, void* stream);
int loner_mlp_wgrad(const loner_net_t* net, const void* packed, int64_t P, const float* d_sigma,
                    const void* acts, float grad_scale, float* d_params, void* scratch, void* stream);

/* ---- a12 (shipped configuration), SURVEY 8f rank 1: multiresolution hash encoding + one hidden layer
 * of 64 neurons = the reference's default `pos_encoding_sigma` / `sigma_network`
 * (cfg/nerf_config/default_nerf_hash.yaml; models/nerf_tcnn.py:35-38, :59-78), and the same encoding with
 * 2-4 hidden layers of 64 (other `sigma_network.n_hidden_layers`).  tiny-cuda-nn semantics
 * (grid type Hash, linear interpolation, fp16 table / weights / activations, fp32 accumulation).
 * Flat fp32 params, tcnn order: W1 [64, E_pad] | W_2 .. W_L [64, 64] | W_out [16, 64] (row 0 used) | table [entries, 2]. */
typedef struct {
  int32_t n_levels;              /* <= 16 */
  int32_t n_features_per_level;  /* 2 */
  int32_t log2_hashmap_size;
  int32_t base_resolution;
  float per_level_scale;         /* <= 0: tcnn's default 2.0 */
  int32_t n_neurons;             /* 64 */
  int32_t n_hidden_layers;       /* 1 (shipped) .. 4 */
  int32_t flags;                 /* 0 = production kernels (warp-level tensor-core head); LONER_HASH_* A/B variants */
} loner_hashnet_t;
/* the head (32 -> 64 -> 1) on scalar CUDA-core code instead of mma.sync tiles (round 1's kernels) */
#define LONER_HASH_SCALAR 1
/* number of coarse levels whose table reductions are aggregated per warp run (0 = built-in default; A/B) */
#define LONER_HASH_AGG_LEVELS(n) (((n) & 0xF) << 4)

int64_t loner_hash_param_count(const loner_hashnet_t* net);
int64_t loner_hash_table_entries(const loner_hashnet_t* net);
int64_t loner_hash_packed_bytes(const loner_hashnet_t* net);     /* fp16 weights + half2 table */
int64_t loner_hash_bwd_scratch_bytes(const loner_hashnet_t* net, int64_t P);
int loner_hash_pack(const loner_hashnet_t* net, const float* params, void* packed, void* stream);
/* forward: pos [P,3] in [-1,1], or rays [n,13] + z_vals [n,S] with P = n*S.  Nothing is stashed:
 * the backward recomputes the forward. */
int loner_hash_fwd(const loner_hashnet_t* net, const void* packed, const float* pos, const float* rays,
                   const float* z_vals, int32_t S, int64_t P, float* sigma, void* stream);
/* backward: d_sigma [P] -> d_params (+=, caller zeroes; table gradients by 8-byte vector atomics, weight
 * gradients through deterministic per-CTA partial sums) and, if d_pos != NULL, d_pos [P,3]. */
int loner_hash_bwd(const loner_hashnet_t* net, const void* packed, const float* pos, const float* rays,
                   const float* z_vals, int32_t S, int64_t P, const float* d_sigma, float grad_scale,
                   float* d_params, float* d_pos, void* scratch, void* stream);

/* ---- a13  raw2outputs (models/rendering_tcnn.py:93-145), sigma-only, far appended, variance.
 * noise: [n,S] standard normals or NULL (then Philox(seed) * raw_noise_std). */
int loner_render_fwd(const float* sigma, const float* z_vals, const float* rays, int64_t n, int32_t S,
                     const float* noise, float raw_noise_std, uint64_t seed, float* weights,
                     float* depth, float* opacity, float* variance, void* stream);

/* generic backward of raw2outputs: upstream grads (any may be NULL) -> d_sigma [n,S] and
 * d_rays [n,13] (columns 3..5 through |d|; += ). */
int loner_render_bwd(const float* sigma, const float* z_vals, const float* rays, int64_t n, int32_t S,
                     const float* noise, float raw_noise_std, uint64_t seed, const float* g_weights,
                     const float* g_depth, const float* g_opacity, const float* g_variance,
                     float* d_sigma, float* d_rays, void* stream);

/* ---- a13+a15+a16+a17  fused raw2outputs + JS dynamic margin + L1_JS loss, forward AND
 * backward in one pass (mapping/optimizer.py:437-595, models/losses.py:29-51).
 * counts: device int32[2] = global (#valid rays, #opaque rays) — the loss normalisers.
 * loss_cfg: {scale_factor, min_depth_eps, min_js, max_js, alpha, los_lambda, depthloss_lambda,
 *            l2 (0: L1_* losses, 1: L2_*), fixed_eps (> 0 selects the *_LOS fixed margin)}.
 * loss_acc: device float[4] += {sum sq depth err, sum |w-w_gt|, sum |opacity-1|, sum eps_dyn}.
 * Outputs weights/depth/opacity/variance/eps_dyn may be NULL.  d_sigma [n,S]; d_rays [n,13] +=
 * (columns 3..5 through |d| and column 12 through far; the path through the network input is
 * loner_mlp_bwd's d_pos). */
int loner_render_loss(const float* sigma, const float* z_vals, const float* rays, const float* depths,
                      const uint8_t* flags, int64_t n, int32_t S, const float* noise, float raw_noise_std,
                      uint64_t seed, const int32_t* counts, const float* loss_cfg9_host, float* loss_acc,
                      float* weights, float* depth, float* opacity, float* variance, float* eps_dyn,
                      float* d_sigma, float* d_rays, void* stream);

/* the scalar loss of Optimizer.compute_loss (mapping/optimizer.py:486-491, :568-580) and the mean margin `_depth_eps`
 * (:503) from loss_acc / counts as loner_render_loss left them (after the multi-GPU all-reduce, if any):
 * out6 = {loss, mean eps_dyn, depth loss, LOS loss, opacity loss, #valid rays}. */
int loner_loss_finalize(const float* loss_acc, const int32_t* counts, float depthloss_lambda, float los_lambda,
                        int32_t S, float* out6, void* stream);

/* d_pos [n,S,3] -> d_rays [n,13] += (origin: sum_s d_pos; direction: sum_s z*d_pos). */
int loner_points_bwd(const float* d_pos, const float* z_vals, int64_t n, int32_t S, float* d_rays,
                     void* stream);

/* ---- a5  pose 6-vectors [t | axis-angle] -> [K,12] = R row-major | t (common/pose_utils.py:288-302
 * tensor_to_transform, i.e. pytorch3d.transforms.axis_angle_to_matrix through the unit quaternion).
 * poses6: the keyframe pose store [n_keyframes,6]; rows [K]: store rows of the window's keyframes. */
int loner_pose_matrices(const float* poses6, const int32_t* rows, int32_t K, const float* shift, float scale,
                        float* poses12, int32_t* status, void* stream);
/* status (device int32, may be NULL; bits are OR-ed in, the caller reads it when it chooses - no host sync here):
 * the reference's runtime guards on this path.  shift: host float[3], with scale the WorldCube of loner_ray_build. */
#define LONER_STATUS_ORIGIN_OUTSIDE 1 /* "ray origins are outside the world cube" assert, common/ray_utils.py:301-303 */
#define LONER_STATUS_BAD_POSE_GRAD 2  /* "Fatal: Encountered invalid gradient in pose.", mapping/optimizer.py:368-370 */
#define LONER_STATUS_BAD_POSE 4       /* "Fatal: Encountered invalid pose tensor.", mapping/optimizer.py:372-374 */

/* chain rule of the above + the pose half of the optimiser (mapping/optimizer.py:249-267,:376): d_poses12 [K,12] ->
 * grad6 [n_keyframes,6] (rows of the window; zero for rows with free_rows[row] == 0), and, if apply != 0, one
 * torch.optim.Adam step on the free rows (per-row step counts in steps [n_keyframes], moments [n_keyframes,6]).
 * A row whose gradient is not finite keeps its pose and sets LONER_STATUS_BAD_POSE_GRAD (the reference raises before
 * optimizer.step()). */
int loner_pose_step(float* poses6, const int32_t* rows, const uint8_t* free_rows, int32_t K, const float* d_poses12,
                    float* grad6, float* exp_avg, float* exp_avg_sq, int32_t* steps, float lr, float beta1,
                    float beta2, float eps, int32_t apply, int32_t* status, void* stream);

/* ---- a18  torch.optim.Adam step on the flat fp32 params (mapping/optimizer.py:257-267,:376)
 * fused with the fp16 repack.  step >= 1. */
int loner_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t count,
                    int32_t step, float lr, float beta1, float beta2, float eps, float grad_unscale,
                    void* stream);

/* ---- a19  Optimizer._step_occupancy_grid + get_logits_grad (mapping/optimizer.py:598-609,
 * models/losses.py:54-62): scatter the pseudo-gradient trilinearly into d_grid [V,V,V]
 * (caller zeroes), then loner_sgd_step applies grid -= lr * d_grid.  flags (may be NULL): rows whose
 * LONER_FLAG_VALID bit is clear were dropped by build_lidar_rays (ray_utils.py:321-322) and never reach
 * `points_fine`; they are skipped. */
int loner_ogm_grad(const float* rays, const float* z_vals, const float* depths, const uint8_t* flags,
                   int64_t n, int32_t S, float scale, int32_t V, float* d_grid, void* stream);
int loner_sgd_step(float* x, const float* g, int64_t count, float lr, void* stream);

#ifdef __cplusplus
}
#endif
#endif

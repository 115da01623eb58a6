/*
 * svs.h — C ABI of libsvolsdf_b200.so: the B200-native VolSDF hot path.
 *
 * The reference (cvlab-stonybrook/s-volsdf) is pure Python/PyTorch and has NO FFI; its plug-in seam is
 * the string-named model class (volsdf/vsdf.py:92-93, volsdf/utils/general.py:10-16).  This header is
 * the boundary the new implementation adds underneath that class API: every entry point below replaces
 * a group of ATen call sites of the reference, cited per function.  INTEGRATION.md shows the ctypes
 * binding (`s-volsdf_b200/_lib.py`) that the Python mirror of `volsdf.model.*` uses.
 *
 * Conventions
 *   - all pointers are DEVICE pointers to contiguous row-major fp32 unless stated; sizes are element
 *     counts; `stream` is a cudaStream_t passed as void*;
 *   - no ownership transfer: inputs, outputs, saved-for-backward and scratch buffers are allocated by
 *     the caller (PyTorch); the `*_floats` queries give workspace sizes in fp32 elements;
 *   - no hidden host synchronisation, no global state besides the last-error string;
 *   - return 0 on success, <0 = svs_status; svs_last_error() gives the message (thread-local).
 */
#ifndef SVS_H_
#define SVS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVS_MAX_LAYERS 12
#define SVS_OPT_MAX_TENSORS 96
#define SVS_MAX_PEERS 8
#define SVS_ABI_VERSION 6

typedef enum {
  SVS_OK = 0,
  SVS_ERR_INVALID = -1,      /* bad argument / unsupported shape */
  SVS_ERR_CUDA = -2,         /* CUDA runtime error (launch, attribute) */
  SVS_ERR_UNSUPPORTED = -3,  /* engine not available for this descriptor */
  SVS_ERR_BOUNDING_SPHERE = -4 /* reference prints "BOUNDING SPHERE PROBLEM!" and exit()s (rend_util.py:209-211) */
} svs_status;

/* arithmetic engine of the MLP kernels */
typedef enum {
  SVS_ENGINE_FP32 = 0, /* fp32 SIMT FFMA tiles: parity mode (rgb/depth <= 1e-3 vs the reference) */
  SVS_ENGINE_TC = 1,   /* tcgen05.mma kind::f16: fp16 operands (10-bit mantissa, TF32-class), fp32 TMEM accumulators;
                          activations stay in shared memory across the layers of a chain (mlp_tc.cuh).  Fastest; sdf
                          ~1e-3, i.e. outside the 1e-3 depth contract */
  SVS_ENGINE_TC_SPLIT = 2 /* tcgen05 with split operands in every FORWARD chain (fp16 hi + fp16 lo halves of activations
                          and weights, 3 MMAs per layer, mlp_tc_fwd3.cuh): sdf / rgb / depth at fp32-class accuracy
                          (<= 1e-5 vs the reference), ReLU masks as the reference; backward chains as SVS_ENGINE_TC
                          (parameter gradients <= 5e-3 relative).  The tensor-core mode that meets the parity contract. */
} svs_engine;

typedef enum { SVS_NET_SDF = 0, SVS_NET_RENDER = 1 } svs_net_kind;
typedef enum { SVS_RENDER_IDR = 0, SVS_RENDER_NERF = 1 } svs_render_mode;

/* Shape of one MLP.  SDF net = ImplicitNetwork (volsdf/model/network.py:10-88): positional encoding of a
 * d_in-vector, n_layers Linear layers, Softplus(beta=100) between them, input re-injected (cat/sqrt2) at
 * skip_layer.  Render net = RenderingNetwork (network.py:134-190): ReLU between layers, sigmoid at the end. */
typedef struct {
  int32_t kind;                    /* svs_net_kind */
  int32_t n_layers;                /* number of Linear layers */
  int32_t d_in;                    /* SDF: point dimension (3, or 4 for the inverted-sphere bg net) */
  int32_t n_freqs;                 /* SDF: PE frequencies of the point; RENDER: PE frequencies of the view dir */
  int32_t skip_layer;              /* SDF: layer whose input is cat[h, PE(x)]/sqrt(2); -1 = none */
  int32_t render_mode;             /* RENDER: svs_render_mode */
  int32_t weight_norm;             /* 1: parameters are (g, v, bias) with W = g*v/|v|_row; 0: (W, bias) */
  int32_t in_dim[SVS_MAX_LAYERS];  /* K of each layer (skip layer: includes the re-injected PE width) */
  int32_t out_dim[SVS_MAX_LAYERS]; /* N of each layer */
  float sphere_radius;             /* SDF: bounding-sphere clamp min(sdf, scale*(R-|x|)) (network.py:108-112); <=0 off */
  float sphere_scale;
} svs_mlp_desc;

/* Per-layer parameter (or gradient) pointers, exactly the reference's state_dict tensors:
 * lin{l}.weight_g (out,1) / lin{l}.weight_v (out,in) / lin{l}.bias, or lin{l}.weight / bias without weight-norm. */
typedef struct {
  float* g[SVS_MAX_LAYERS]; /* NULL when weight_norm == 0 */
  float* v[SVS_MAX_LAYERS];
  float* b[SVS_MAX_LAYERS];
} svs_mlp_params;

const char* svs_last_error(void);
int svs_abi_version(void);
/* 1 when the library provides the engine (both are always built for sm_100a) */
int svs_has_engine(int engine);
/* number of kernels this library has launched so far (bench.py's gpu_launches) */
int64_t svs_launch_count(void);
/* Optional per-kernel profiler: while enabled every launch is bracketed by CUDA events on its stream.
 * svs_prof_collect synchronises them and writes "name\tlaunches\ttotal_ms\tflops\tbytes\n" per kernel name
 * (flops/bytes = ALGORITHMIC work summed over the launches) into buf; returns the length or <0. */
int svs_prof_enable(int on);
int64_t svs_prof_collect(char* buf, int64_t cap);

/* ---------------------------------------------------------------------------------------------------
 * Weights.  Replaces the weight_norm pre-forward hook (network.py:64-65: W = g*v/|v| recomputed on every
 * call) and its autograd backward.  `wbuf` holds the effective weights of all layers in the layout the
 * kernels consume (svs_mlp_wbuf_floats elements); the same layout is used for the gradient accumulator.
 * ------------------------------------------------------------------------------------------------- */
int64_t svs_mlp_wbuf_floats(const svs_mlp_desc* d, int engine);
int svs_mlp_prepare(const svs_mlp_desc* d, const svs_mlp_params* p, float* wbuf, int engine, void* stream);
/* dwbuf (dL/dW_eff, dL/db in wbuf layout) -> dL/dg, dL/dv, dL/db written (not accumulated) to `grads` */
int svs_mlp_param_grads(const svs_mlp_desc* d, const svs_mlp_params* p, const float* wbuf,
                        const float* dwbuf, const svs_mlp_params* grads, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * SDF network (ImplicitNetwork).  x: (P, d_in).
 *   svs_sdf_forward         = forward() (network.py:71-88): y (P, ldy>=out) raw outputs, and/or
 *                             get_sdf_vals() (network.py:125-131): sdf (P) with the sphere clamp.
 *   svs_sdf_outputs_forward = get_outputs()/gradient() (network.py:90-123): raw outputs y (P, ldy),
 *                             clamped sdf (P), analytic d sdf/dx (P, d_in); `saved` (optional) receives the
 *                             activations the backward needs (svs_sdf_saved_floats).  Points [0, n_clamped) take the
 *                             sphere clamp (get_outputs), points [n_clamped, P) do not (gradient(): the eikonal
 *                             samples ride in the same launch as the ray samples).
 *   svs_sdf_outputs_backward= autograd backward + double-backward of the above (loss.backward() through
 *                             network.py:105-123): dy (P, ldy) = dL/d raw outputs (col 0 is combined with
 *                             d_sdf through the clamp), d_sdf (P) optional, d_grad (P, d_in) optional;
 *                             accumulates dL/dW_eff, dL/db into dwbuf.
 * `ws` is scratch of svs_sdf_ws_floats(d, P, training) elements.  ldy = svs_sdf_ldy(d).
 * ------------------------------------------------------------------------------------------------- */
int32_t svs_sdf_ldy(const svs_mlp_desc* d);
int64_t svs_sdf_ws_floats(const svs_mlp_desc* d, int64_t P, int with_grad, int engine);
int64_t svs_sdf_saved_floats(const svs_mlp_desc* d, int64_t P, int engine);
int svs_sdf_forward(const svs_mlp_desc* d, const float* wbuf, const float* x, int64_t P, float* y,
                    float* sdf, float* ws, int engine, void* stream);
int svs_sdf_outputs_forward(const svs_mlp_desc* d, const float* wbuf, const float* x, int64_t P, int64_t n_clamped,
                            float* y, float* sdf, float* grad, float* saved, float* ws, int engine,
                            void* stream);
int svs_sdf_outputs_backward(const svs_mlp_desc* d, const float* wbuf, const float* x, int64_t P, int64_t n_clamped,
                             const float* saved, const float* y, const float* dy, const float* d_sdf,
                             const float* d_grad, float* dwbuf, float* ws, int engine, void* stream);
int64_t svs_sdf_bwd_ws_floats(const svs_mlp_desc* d, int64_t P, int engine);

/* Positional encoding on its own (embedder.py:10-36 `Embedder.embed`): x (P,d_in) -> out (P, d_in*(1+2*n_freqs)). */
int svs_embed(const float* x, int64_t P, int32_t d_in, int32_t n_freqs, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Rendering network (RenderingNetwork.forward, network.py:170-190).
 *   idr : input = cat[points(3), PE(view_dirs), normals(3), feat(F)];  nerf: cat[PE(view_dirs), feat(F)].
 *   feat is read from `feat` with row stride ld_feat (pass y+1, ldy to consume the SDF net output in place).
 *   backward returns dL/dnormals (P,3) and dL/dfeat (written into d_feat with stride ld_dfeat) and
 *   accumulates weight gradients into dwbuf.  points / view dirs carry no gradient on the hot path
 *   (the sampler runs under no_grad, ray_sampler.py:88-89).
 * ------------------------------------------------------------------------------------------------- */
int64_t svs_render_saved_floats(const svs_mlp_desc* d, int64_t P, int engine);
int64_t svs_render_ws_floats(const svs_mlp_desc* d, int64_t P, int engine);
int svs_render_forward(const svs_mlp_desc* d, const float* wbuf, const float* points, const float* view_dirs,
                       const float* normals, const float* feat, int32_t ld_feat, int64_t P, float* rgb,
                       float* saved, int engine, void* stream);
int svs_render_backward(const svs_mlp_desc* d, const float* wbuf, int64_t P, const float* saved,
                        const float* rgb, const float* d_rgb, float* d_normals, float* d_feat,
                        int32_t ld_dfeat, float* dwbuf, float* ws, int engine, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Rays.  svs_raygen = rend_util.get_camera_params + lift (rend_util.py:60-95,143-156) called twice as in
 * network.py:213-217: uv (R,2) pixels, pose (4,4) cam->world, intrinsics (4,4) ->
 * ray_dirs (R,3) unit, cam_loc (R,3) broadcast, depth_scale (R) = z of the camera-frame unit ray.
 * svs_sphere_intersections = rend_util.get_sphere_intersections (rend_util.py:200-216): near/far (R,2);
 * `bad_flag` (device int) is set to 1 if any discriminant <= 0 (the reference exit()s there).
 * svs_ray_points: points (R,S,3) = cam_loc + z * dir  (network.py:227-228).
 * svs_depth2pts_outside = VolSDFNetworkBG.depth2pts_outside (network_bg.py:182-214): pts (R,S,4), depth_real (R,S).
 * ------------------------------------------------------------------------------------------------- */
int svs_raygen(const float* uv, const float* pose, const float* intrinsics, int64_t R, float* ray_dirs,
               float* cam_loc, float* depth_scale, void* stream);
int svs_sphere_intersections(const float* cam_loc, const float* ray_dirs, int64_t R, float radius,
                             float* near_far, int32_t* bad_flag, void* stream);
int svs_ray_points(const float* cam_loc, const float* ray_dirs, const float* z, int64_t R, int32_t S,
                   int32_t ldz, float* points, void* stream);
int svs_depth2pts_outside(const float* cam_loc, const float* ray_dirs, const float* depth, int64_t R,
                          int32_t S, float radius, float* pts, float* depth_real, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * ErrorBoundSampler (volsdf/model/ray_sampler.py:46-229), one warp per ray, rows staged in shared memory.
 * The Python side owns the loop (ray_sampler.py:83) because each iteration needs the SDF MLP:
 *   init      : uniform z (ray_sampler.py:22-43, training jitter from t_rand) + Lemma-2 beta (:76-78)
 *   bound     : merge sdf by samples_idx (:90-95), d* (:98-111), error(beta0) + bisection (:114-123);
 *               writes beta (R) and ORs `beta > beta0` into *not_converged (device int)  (:136)
 *   resample  : density/transmittance (:126-132), pdf/cdf (:138-163), inverse CDF (:167-185) and, when
 *               continuing, the stable merge z,samples_idx = sort(cat[z, samples]) (:189-190)
 *   finalize  : z = sort(cat[samples, near, far, z[:, extra_idx]]) (:193-208), z_eik gather (:211-212)
 * exact=1 evaluates exp/expm1 and the prefix sums in fp64 (bit-exact against oracle/volsdf_oracle.py).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  float near;            /* ray_sampler.near */
  float far;             /* 2*scene_bounding_sphere, or <0: per-ray far from `far_ray` (inverse_sphere_bg) */
  float eps;             /* error-bound target */
  float add_tiny;
  float inv4logeps;      /* 1/(4*log(1+eps)) evaluated in fp32 by the caller as the reference does (:77) */
  int32_t beta_iters;
  int32_t exact;
} svs_sampler_cfg;

int svs_sampler_init(const svs_sampler_cfg* c, int64_t R, int32_t n, const float* t_lin, const float* t_rand,
                     const float* far_ray, float* z, float* beta, void* stream);
int svs_sampler_bound(const svs_sampler_cfg* c, int64_t R, int32_t n, int32_t n_new, const float* z,
                      const float* sdf_old, const float* sdf_new, const int32_t* samples_idx, float* sdf,
                      const float* beta_param, float beta_min, float* beta, int32_t* not_converged,
                      void* stream);
int svs_sampler_resample(const svs_sampler_cfg* c, int64_t R, int32_t n, int32_t n_u, int32_t cont,
                         const float* z, const float* sdf, const float* beta, const float* u,
                         int32_t u_per_ray, float* samples, int32_t* inds, float* z_merged,
                         int32_t* samples_idx, void* stream);
int svs_sampler_finalize(const svs_sampler_cfg* c, int64_t R, int32_t n, int32_t n_samples, const float* z,
                         const float* samples, const int32_t* extra_idx, int32_t n_extra,
                         const float* far_ray, const int64_t* eik_idx, float* z_final, float* z_eik,
                         void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Density + transmittance + compositing, fused (density.py:21-35, network.py:281-295,239-248,270-276;
 * bg variants network_bg.py:147-180).  One warp per ray.
 *   flags: SVS_COMP_ABS_DENSITY  sigma=|s| (AbsDensity) instead of Laplace
 *          SVS_COMP_REVERSED     dists = z[i]-z[i+1] (flipped inverse-sphere samples, network_bg.py:170)
 *          SVS_COMP_FAST         opt-in throughput arithmetic (MUFU exp, fp32 scans); weights differ by ~1e-6
 *          SVS_COMP_ZMAX_TAIL    last dist = z_max - z[S-1] and bg_transmittance output (network_bg.py:152-162)
 *   beta_param: device scalar (density.beta); beta = |beta_param| + beta_min.
 *   outputs: weights (R,S), rgb_values (R,3), depth_values (R) = depth_scale*sum(w z)/(sum(w)+1e-8),
 *            normal_map (R,3) (eval: sum w * g/|g|, NULL to skip), bg_trans (R) (ZMAX_TAIL only).
 *   backward inputs: dL/drgb_values (R,3), dL/ddepth_values (R), dL/dweights (R,S) (dense, may be NULL),
 *            dL/dbg_trans (R) (may be NULL); outputs dL/dsdf (R,S), dL/drgb (R,S,3), dL/dbeta_param
 *            (device scalar, accumulated with atomicAdd).
 * ------------------------------------------------------------------------------------------------- */
#define SVS_COMP_ABS_DENSITY 1
#define SVS_COMP_REVERSED 2
#define SVS_COMP_ZMAX_TAIL 4
#define SVS_COMP_FAST 8 /* MUFU exp + fp32 scans instead of the oracle's canonical arithmetic (libm exp, fp64 prefix sums) */
int svs_composite_forward(const float* z, const float* sdf, const float* rgb, const float* normals,
                          const float* beta_param, float beta_min, const float* depth_scale,
                          const float* z_max, int64_t R, int32_t S, int32_t flags, float* weights,
                          float* rgb_values, float* depth_values, float* normal_map, float* bg_trans,
                          void* stream);
int svs_composite_backward(const float* z, const float* sdf, const float* rgb, const float* beta_param,
                           float beta_min, const float* depth_scale, const float* z_max, int64_t R,
                           int32_t S, int32_t flags, const float* d_rgb_values, const float* d_depth_values,
                           const float* d_weights, const float* d_bg_trans, float* d_sdf, float* d_rgb,
                           float* d_beta_param, void* stream);

/* Stand-alone density modules (density.py:16-35): out = Laplace(sdf; beta) or |sdf|.  sdf is (R,S);
 * beta is |*beta_param|+beta_min, or beta_rows[r] when beta_rows != NULL (the sampler's per-ray override).
 * backward: d_sdf = d_out * dsigma/ds ; *d_beta_param += sum d_out * dsigma/dbeta (only when beta_rows == NULL). */
int svs_density_forward(const float* sdf, int64_t R, int32_t S, const float* beta_param, float beta_min,
                        const float* beta_rows, int32_t abs_density, float* out, void* stream);
int svs_density_backward(const float* sdf, int64_t R, int32_t S, const float* beta_param, float beta_min,
                         const float* beta_rows, int32_t abs_density, const float* d_out, float* d_sdf,
                         float* d_beta_param, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Optimiser step of the reference's loop (volsdf/vsdf.py:214-219,454-464): clip_grad_norm_(max_norm) + the NaN/Inf
 * guard (non-finite gradients are zeroed, Adam still steps) + torch.optim.Adam (no amsgrad / weight decay) for all
 * parameter tensors in two launches.  Arrays of n_tensors HOST pointers to device tensors; *step_count = number of
 * this step (device float, incremented by the caller); scratch = 2 device floats (sum g^2, non-finite flag; the
 * pre-clip norm is sqrt(scratch[0]) afterwards).  max_norm <= 0 disables clipping.
 * ------------------------------------------------------------------------------------------------- */
int svs_adam_step(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const int64_t* numel, float lr, float beta1, float beta2, float eps,
                  float max_norm, int32_t skip_nonfinite, const float* step_count, float* scratch, void* stream);

/* Data-parallel variant ("NCCL allreduce(sum) of one flat fp32 gradient buffer per step ... then grad-clip
 * and Adam run redundantly on every rank") as ONE kernel over NVLink peer memory: all-reduce(mean) + clip + NaN guard +
 * Adam.  peer_grads[r] = rank r's flat gradient buffer (n_flat floats, every tensor starting on a multiple of 4 floats, in
 * the order of `params`), mapped into this process (e.g. torch.distributed._symmetric_memory); peer_flags[r] = rank r's
 * 2 * SVS_MAX_PEERS uint32 flags (zeroed once, symmetric memory); gmean = local n_flat floats receiving the averaged
 * gradient; ctrl = 4 + 1024 local uint32 (zeroed once).  All ranks must call it the same number of times.  The launch does not
 * return before every peer has read this rank's buffer, so later work on the stream may overwrite it.  world == 1: plain
 * clip + Adam from peer_grads[0]. */
int svs_adam_step_allreduce(int32_t n_tensors, float* const* params, float* const* exp_avg, float* const* exp_avg_sq,
                            const int64_t* numel, int32_t world, int32_t rank, const float* const* peer_grads,
                            uint32_t* const* peer_flags, int64_t n_flat, float* gmean, float lr, float beta1, float beta2,
                            float eps, float max_norm, int32_t skip_nonfinite, const float* step_count, float* scratch,
                            uint32_t* ctrl, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * MVS cost lookup of the reference's training loop — VolOpt.cost_mapping (volsdf/vsdf.py:382-452).
 * Every ray sample xyz (N, D, 3) is projected into each source view (pose, pinhole with skew, normalised at image
 * resolution img_h x img_w), the view's per-pixel depth range is read with two bilinear lookups, the probability
 * volume with one trilinear lookup (grid_sample: bilinear, zero padding, align_corners=True).  Outputs, all (N, D):
 *   cost_j   sum over the views with same_view == 0                       (results_cost_j, the loss's `pj`)
 *   cost_mvs the view with same_view == 1, zeroed where valid == 0        (results_cost_mvs, the loss's `pi`)
 *   valid    1 where the sample lies inside the frustum and depth range of at least one other view (valid_mask)
 * `views` is a HOST array of n_views (<= 8) descriptors holding DEVICE pointers; nothing is retained after the call.
 * ------------------------------------------------------------------------------------------------- */
typedef struct svs_mvs_view {
  const float* cost;   /* (Dz, H, W) probability volume            (self.costs[i][0],  vsdf.py:369) */
  const float* z_near; /* (H, W) first depth hypothesis plane       (self.z_mvs[i][0, 0],  vsdf.py:420) */
  const float* z_far;  /* (H, W) last depth hypothesis plane        (self.z_mvs[i][0, -1]) */
  int32_t Dz, H, W;
  float fx, fy, cx, cy, sk; /* K[0,0], K[1,1], K[0,2], K[1,2], K[0,1] of the view (vsdf.py:398-399) */
  float c2w[12];            /* rows 0..2 of the camera-to-world pose, row-major (vsdf.py:396-397) */
  int32_t same_view;        /* 1: the batch's own image (ts[0] == id_k, vsdf.py:392); used when own_view == NULL */
  int32_t view_id;          /* id_k of the view (self.trains_i[i]); compared with *own_view when that is given */
} svs_mvs_view;
/* own_view: optional DEVICE int32 holding the batch's image index ts[0]; when non-NULL the own view is decided on the
 * device (view_id == *own_view), so the call can sit inside a CUDA graph whose replays change the batch image. */
int svs_cost_mapping(const float* xyz, int64_t N, int32_t D, const svs_mvs_view* views, int32_t n_views,
                     int32_t img_h, int32_t img_w, int32_t inverse_depth, const int32_t* own_view, float* cost_j,
                     float* cost_mvs, uint8_t* valid, void* stream);

/* The same lookup fused with the MVS term of VolSDFLoss (volsdf/model/loss.py:53-67, `get_mvs_loss`): p_i p_j stay in
 * registers.  weights (N, D) are the rendering weights of the samples; gce the exponent of the generalised cross-entropy
 * (1: -pw w ; 0: -pw log(w + 1e-8) ; else -pw w.detach()^gce log(w + 1e-8)), confi the confidence threshold on
 * sum_s p_i p_j.  Outputs: ray_loss (N) = [sum_s p_i p_j > confi] * sum_s term (the loss is its mean over the rays),
 * conf_ray (N) = sum_s p_i p_j (the sparsity and uncertain-ray terms of the loss branch on it, loss.py:40-45,69-78),
 * d_weights (N, D) = d ray_loss / d weights (the gradient the compositor backward consumes, before the 1/N and the loss
 * weight).  D <= 256. */
int svs_mvs_loss(const float* xyz, int64_t N, int32_t D, const svs_mvs_view* views, int32_t n_views, int32_t img_h,
                 int32_t img_w, int32_t inverse_depth, const int32_t* own_view, const float* weights, float gce, float confi,
                 float* ray_loss, float* conf_ray, float* d_weights, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVS_H_ */

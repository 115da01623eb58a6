// MVS cost lookup of the reference's training loop: VolOpt.cost_mapping (volsdf/vsdf.py:382-452).
//
// Every ray sample (N rays x D samples) is projected into each source view (world -> camera with the view's
// pose, pinhole with skew, normalised to [-1, 1] at image resolution), the per-pixel depth range of the MVS
// stage-0 hypotheses is read with two bilinear lookups (first / last hypothesis plane), the sample's depth is
// normalised into that range (inverse-depth or linear) and the probability volume is read with one trilinear
// lookup — grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True) each.  The batch's own view
// feeds `cost_mvs` (p_i), the other views are summed into `cost_j` (p_j) and decide `valid`.
//
// One thread per sample, all views in one launch (the reference: ~35 elementwise launches + 3 grid_sample per view).
// Bound by L2 gathers: 16 scattered 4-byte reads per view next to 12 B in / 9 B out per sample.  Compiled with
// -fmad=false: the coordinate arithmetic is the reference's chain of separately rounded fp32 elementwise ops (the
// validity tests compare against thresholds, so an extra rounding could flip a sample at a frustum edge).
#include "svs_common.cuh"

namespace svs {

constexpr int kMaxMvsViews = 8;

struct MvsArgs {
  svs_mvs_view v[kMaxMvsViews];
  int n_views;
};

// one corner of a (bi/tri)linear lookup with zero padding
__device__ __forceinline__ float tap2(const float* __restrict__ img, int H, int W, int x, int y) {
  return (x >= 0 && x < W && y >= 0 && y < H) ? __ldg(img + (size_t)y * W + x) : 0.f;
}
__device__ __forceinline__ float tap3(const float* __restrict__ vol, int Dz, int H, int W, int x, int y, int z) {
  return (x >= 0 && x < W && y >= 0 && y < H && z >= 0 && z < Dz) ? __ldg(vol + ((size_t)z * H + y) * W + x) : 0.f;
}
// grid_sample, align_corners=True: pixel = (coord + 1) / 2 * (size - 1); corner weights as ATen forms them
__device__ __forceinline__ float bilinear(const float* __restrict__ img, int H, int W, float xn, float yn) {
  const float ix = ((xn + 1.f) / 2.f) * (float)(W - 1), iy = ((yn + 1.f) / 2.f) * (float)(H - 1);
  const float fx = floorf(ix), fy = floorf(iy);
  // far outside: every tap is padding (and the float -> int conversion below stays defined)
  if (!(fx > -2.f && fx < (float)W && fy > -2.f && fy < (float)H)) return 0.f;
  const int x0 = (int)fx, y0 = (int)fy;
  const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
  float r = tap2(img, H, W, x0, y0) * (wx0 * wy0);
  r += tap2(img, H, W, x0 + 1, y0) * (wx1 * wy0);
  r += tap2(img, H, W, x0, y0 + 1) * (wx0 * wy1);
  r += tap2(img, H, W, x0 + 1, y0 + 1) * (wx1 * wy1);
  return r;
}
__device__ __forceinline__ float trilinear(const float* __restrict__ vol, int Dz, int H, int W, float xn, float yn, float zn) {
  const float ix = ((xn + 1.f) / 2.f) * (float)(W - 1), iy = ((yn + 1.f) / 2.f) * (float)(H - 1),
              iz = ((zn + 1.f) / 2.f) * (float)(Dz - 1);
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  if (!(fx > -2.f && fx < (float)W && fy > -2.f && fy < (float)H && fz > -2.f && fz < (float)Dz)) return 0.f;
  const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy, wz1 = iz - fz, wz0 = (fz + 1.f) - iz;
  float r = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
    const float w = (dx ? wx1 : wx0) * (dy ? wy1 : wy0) * (dz ? wz1 : wz0);
    r += tap3(vol, Dz, H, W, x0 + dx, y0 + dy, z0 + dz) * w;
  }
  return r;
}

// the lookup of one sample in all views: p_j (sum over the other views), p_i (the batch's own view), validity
__device__ __forceinline__ void mvs_lookup(const MvsArgs& a, float px, float py, float pz, float half_w, float half_h, int inverse_depth,
                                           const int32_t* __restrict__ own_view, float* sum_out, float* own_out, bool* ok_out) {
  float sum = 0.f, own = 0.f;
  bool ok = false;
  const int own_id = own_view ? __ldg(own_view) : 0;
  for (int k = 0; k < a.n_views; ++k) {
    const svs_mvs_view& v = a.v[k];
    // world -> camera: (p - t) @ R, R = c2w[:, :3]                                       vsdf.py:401-402
    const float dx = px - v.c2w[3], dy = py - v.c2w[7], dz = pz - v.c2w[11];
    float cxm = dx * v.c2w[0] + dy * v.c2w[4] + dz * v.c2w[8];
    float cym = dx * v.c2w[1] + dy * v.c2w[5] + dz * v.c2w[9];
    float z = dx * v.c2w[2] + dy * v.c2w[6] + dz * v.c2w[10];
    float x = cxm / z, y = cym / z;                                                     // :406
    y = y * v.fy + v.cy;                                                                // :407
    x = x * v.fx + v.cx + (y - v.cy) * v.sk / v.fy;                                     // :408
    x = x / half_w - 1.f;                                                               // :410
    y = y / half_h - 1.f;                                                               // :411
    const bool bad = (z < 1e-5f) || (x > 1.001f) || (x < -1.001f) || (y > 1.001f) || (y < -1.001f);   // :418
    if (bad) x = y = z = -99.f;                                                         // :419
    const float near = bilinear(v.z_near, v.H, v.W, x, y);                              // :424
    float far = bilinear(v.z_far, v.H, v.W, x, y);                                      // :425
    float zn;
    if (inverse_depth) {                                                                // :426-428
      if (bad) far = 1e-8f;
      zn = 2.f * (1.f - near / z) / (1.f - near / far) - 1.f;
    } else {                                                                            // :432
      zn = 2.f * (z - near) / (far - near) - 1.f;
    }
    const bool bad2 = (near < 1e-5f) || (far < 1e-5f) || (zn > 1.01f) || (zn < -1.01f) || bad;   // :434
    const float c = bad2 ? 0.f : trilinear(v.cost, v.Dz, v.H, v.W, x, y, zn);           // :435-440 (-99 -> all padding)
    if (own_view ? (v.view_id == own_id) : (v.same_view != 0)) {
      own = c;                                                                          // :443-444
    } else {
      sum += c;                                                                         // :446-448
      ok = ok || !bad2;
    }
  }
  *sum_out = sum;
  *own_out = ok ? own : 0.f;                                                            // :450
  *ok_out = ok;
}

__global__ void __launch_bounds__(256)
cost_mapping_kernel(const MvsArgs a, const float* __restrict__ xyz, int64_t n, float half_w, float half_h, int inverse_depth,
                    const int32_t* __restrict__ own_view, float* __restrict__ cost_j, float* __restrict__ cost_mvs, uint8_t* __restrict__ valid) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float sum, own;
  bool ok;
  mvs_lookup(a, __ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2), half_w, half_h, inverse_depth, own_view, &sum, &own, &ok);
  cost_j[i] = sum;
  cost_mvs[i] = own;
  valid[i] = ok ? 1 : 0;
}

// Cost lookup fused with the MVS term of VolSDFLoss (volsdf/model/loss.py:53-67): one warp per ray, a lane per sample
// (D <= 32 * kMvsPerLane).  p_i p_j never leave the registers: the kernel writes the ray's loss term, its confidence
// sum_s p_i p_j (the sparsity / uncertain-ray terms branch on it) and d(term) / d weights for the compositor backward.
constexpr int kMvsPerLane = 8;
__global__ void __launch_bounds__(256)
mvs_loss_kernel(const MvsArgs a, const float* __restrict__ xyz, const float* __restrict__ weights, int64_t N, int D, float half_w,
                float half_h, int inverse_depth, const int32_t* __restrict__ own_view, float gce, float confi,
                float* __restrict__ ray_loss, float* __restrict__ conf_ray, float* __restrict__ d_weights) {
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= N) return;
  float pw[kMvsPerLane], w[kMvsPerLane];
  float conf = 0.f;
#pragma unroll
  for (int k = 0; k < kMvsPerLane; ++k) {
    const int s = lane + 32 * k;
    pw[k] = 0.f;
    w[k] = 0.f;
    if (s < D) {
      const int64_t i = r * D + s;
      float sum, own;
      bool ok;
      mvs_lookup(a, __ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2), half_w, half_h, inverse_depth, own_view, &sum, &own, &ok);
      pw[k] = own * sum;                  // pi * pj
      w[k] = __ldg(weights + i);
      conf += pw[k];
    }
  }
  conf = warp_sum(conf);
  const float keep = conf > confi ? 1.f : 0.f;   // loss.py:64: only rays the volumes are confident about
  float term = 0.f;
#pragma unroll
  for (int k = 0; k < kMvsPerLane; ++k) {
    const int s = lane + 32 * k;
    if (s < D) {
      float per, dw;
      if (gce == 1.f) {                   // -pw w
        per = -pw[k] * w[k];
        dw = -pw[k];
      } else if (gce == 0.f) {            // -pw log(w + 1e-8)
        per = -pw[k] * logf(w[k] + 1e-8f);
        dw = -pw[k] / (w[k] + 1e-8f);
      } else {                            // -pw w.detach()^gce log(w + 1e-8)
        const float wg = powf(w[k], gce);
        per = -pw[k] * wg * logf(w[k] + 1e-8f);
        dw = -pw[k] * wg / (w[k] + 1e-8f);
      }
      term += per;
      d_weights[r * D + s] = keep * dw;
    }
  }
  term = warp_sum(term);
  if (lane == 0) {
    ray_loss[r] = keep * term;
    conf_ray[r] = conf;
  }
}

}  // namespace svs

using namespace svs;

extern "C" int svs_cost_mapping(const float* xyz, int64_t N, int32_t D, const svs_mvs_view* views, int32_t n_views,
                                int32_t img_h, int32_t img_w, int32_t inverse_depth, const int32_t* own_view, float* cost_j,
                                float* cost_mvs, uint8_t* valid, void* stream) {
  SVS_CHECK_ARG(N >= 0 && D >= 1, "svs_cost_mapping: need N >= 0, D >= 1");
  SVS_CHECK_ARG(xyz && cost_j && cost_mvs && valid && views, "svs_cost_mapping: null pointer");
  SVS_CHECK_ARG(n_views >= 1 && n_views <= kMaxMvsViews, "svs_cost_mapping: 1 <= n_views <= %d (got %d)", kMaxMvsViews, n_views);
  SVS_CHECK_ARG(img_h >= 2 && img_w >= 2, "svs_cost_mapping: image resolution must be at least 2 x 2");
  MvsArgs a;
  memset(&a, 0, sizeof(a));
  a.n_views = n_views;
  for (int k = 0; k < n_views; ++k) {
    SVS_CHECK_ARG(views[k].cost && views[k].z_near && views[k].z_far, "svs_cost_mapping: view %d has a null volume", k);
    SVS_CHECK_ARG(views[k].Dz >= 1 && views[k].H >= 1 && views[k].W >= 1, "svs_cost_mapping: view %d has an empty volume", k);
    a.v[k] = views[k];
  }
  const int64_t n = N * (int64_t)D;
  if (n == 0) return SVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // bytes: the sample in, three results out, 16 gathered taps per view
  ProfScope ps("cost_mapping", 0.0, (double)n * (12.0 + 9.0 + 64.0 * n_views), st);
  // (_w - 1) / 2 and (_h - 1) / 2 are Python floats in the reference; the tensor is divided by their fp32 value
  const float half_w = (float)((double)(img_w - 1) / 2.0), half_h = (float)((double)(img_h - 1) / 2.0);
  cost_mapping_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(a, xyz, n, half_w, half_h, inverse_depth, own_view, cost_j, cost_mvs, valid);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

extern "C" int svs_mvs_loss(const float* xyz, int64_t N, int32_t D, const svs_mvs_view* views, int32_t n_views, int32_t img_h,
                            int32_t img_w, int32_t inverse_depth, const int32_t* own_view, const float* weights, float gce,
                            float confi, float* ray_loss, float* conf_ray, float* d_weights, void* stream) {
  SVS_CHECK_ARG(N >= 0 && D >= 1 && D <= 32 * kMvsPerLane, "svs_mvs_loss: need N >= 0, 1 <= D <= %d (got %d)", 32 * kMvsPerLane, D);
  SVS_CHECK_ARG(xyz && weights && ray_loss && conf_ray && d_weights && views, "svs_mvs_loss: null pointer");
  SVS_CHECK_ARG(n_views >= 1 && n_views <= kMaxMvsViews, "svs_mvs_loss: 1 <= n_views <= %d (got %d)", kMaxMvsViews, n_views);
  SVS_CHECK_ARG(img_h >= 2 && img_w >= 2, "svs_mvs_loss: image resolution must be at least 2 x 2");
  SVS_CHECK_ARG(gce >= 0.f, "svs_mvs_loss: gce must be >= 0");
  MvsArgs a;
  memset(&a, 0, sizeof(a));
  a.n_views = n_views;
  for (int k = 0; k < n_views; ++k) {
    SVS_CHECK_ARG(views[k].cost && views[k].z_near && views[k].z_far, "svs_mvs_loss: view %d has a null volume", k);
    SVS_CHECK_ARG(views[k].Dz >= 1 && views[k].H >= 1 && views[k].W >= 1, "svs_mvs_loss: view %d has an empty volume", k);
    a.v[k] = views[k];
  }
  if (N == 0) return SVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = N * (int64_t)D;
  ProfScope ps("mvs_loss", 0.0, (double)n * (12.0 + 8.0 + 64.0 * n_views) + 8.0 * (double)N, st);
  const float half_w = (float)((double)(img_w - 1) / 2.0), half_h = (float)((double)(img_h - 1) / 2.0);
  mvs_loss_kernel<<<(unsigned)cdiv(N * 32, 256), 256, 0, st>>>(a, xyz, weights, N, D, half_w, half_h, inverse_depth, own_view, gce, confi,
                                                              ray_loss, conf_ray, d_weights);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

// K6-K10: ErrorBoundSampler (volsdf/model/ray_sampler.py:46-229) as warp-per-ray kernels.
//
// A ray's rows live in shared memory for the whole kernel (coalesced global <-> smem copies); the 11
// error-bound evaluations of one iteration (1 at beta0 + 10 bisection steps) never touch HBM.  Scans run
// over 32-wide segments with warp shuffles.  This translation unit is compiled with -fmad=false: in
// EXACT mode every +,-,*,/ is an IEEE fp32 op in the reference's operator order, exp/expm1 are evaluated
// in fp64 and rounded once, prefix sums accumulate in fp64 and round every prefix — the canonical
// arithmetic of oracle/volsdf_oracle.py, so indices and sample counts match it bit-for-bit.
#include <math_constants.h>

#include "svs_common.cuh"

namespace svs {

constexpr int kSampWarps = 4;

template <bool X>
__device__ __forceinline__ float t_exp(float x) {
  if (X) return (float)exp((double)x);
  return __expf(x);
}
template <bool X>
__device__ __forceinline__ float t_expm1(float x) {
  if (X) return (float)expm1((double)x);
  return expm1f(x);
}

template <bool X>
struct Acc { typedef double type; };
template <>
struct Acc<false> { typedef float type; };

// inclusive scan of one value per lane; returns this lane's inclusive prefix
template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive prefix from the inclusive one: the previous lane's value (never `incl - self`, which cancels
// and loses the low bits of a small prefix next to a large element)
template <typename T>
__device__ __forceinline__ T excl_of(T incl, int lane) {
  T prev = __shfl_up_sync(0xffffffffu, incl, 1);
  return lane ? prev : (T)0;
}

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

// LaplaceDensity.density_func with an explicit beta (density.py:21-26)
template <bool X>
__device__ __forceinline__ float laplace(float s, float beta) {
  float alpha = 1.0f / beta;
  float em = t_expm1<X>(-fabsf(s) / beta);
  return alpha * (0.5f + (0.5f * sgn(s)) * em);
}

// Theorem-1 bound of one interval (ray_sampler.py:98-111): a = dist, d0/d1 = sdf at its ends
__device__ __forceinline__ float d_star_of(float a, float d0, float d1) {
  float b = fabsf(d0), c = fabsf(d1);
  bool first = a * a + b * b <= c * c;
  bool second = a * a + c * c <= b * b;
  float ds = 0.f;
  if (first) ds = b;
  if (second) ds = c;
  float s = (a + b + c) / 2.0f;
  float area = s * (s - a) * (s - b) * (s - c);
  if (!first && !second && (b + c - a > 0.f)) ds = (2.0f * sqrtf(area)) / a;
  return (sgn(d1) * sgn(d0) == 1.0f) ? ds : 0.f;
}

// ----------------------------------------------------------------------------------------------------
// init: UniformSampler.get_z_vals (ray_sampler.py:22-43) + Lemma-2 beta (ray_sampler.py:76-78)
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSampWarps * 32)
sampler_init_kernel(svs_sampler_cfg c, int64_t R, int n, const float* __restrict__ t_lin,
                    const float* __restrict__ t_rand, const float* __restrict__ far_ray,
                    float* __restrict__ z, float* __restrict__ beta) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sz = smem + warp * n;
  for (int64_t ray = blockIdx.x * (int64_t)kSampWarps + warp; ray < R; ray += (int64_t)gridDim.x * kSampWarps) {
    const float far = (c.far < 0.f) ? far_ray[ray] : c.far;
    for (int i = lane; i < n; i += 32) {
      float t = t_lin[i];
      sz[i] = c.near * (1.0f - t) + far * t;
    }
    __syncwarp();
    if (t_rand) {  // stratified jitter (training)
      float zj[32];  // n <= 1024
      int cnt = 0;
      for (int i = lane; i < n; i += 32, ++cnt) {
        float zi = sz[i];
        float lower = (i == 0) ? zi : 0.5f * (zi + sz[i - 1]);
        float upper = (i == n - 1) ? zi : 0.5f * (sz[i + 1] + zi);
        zj[cnt] = lower + (upper - lower) * t_rand[ray * n + i];
      }
      __syncwarp();
      cnt = 0;
      for (int i = lane; i < n; i += 32, ++cnt) sz[i] = zj[cnt];
      __syncwarp();
    }
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) {
      float zi = sz[i];
      z[ray * n + i] = zi;
      if (i < n - 1) {
        float d = sz[i + 1] - zi;
        acc += (double)(d * d);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) beta[ray] = sqrtf(c.inv4logeps * (float)acc);
    __syncwarp();
  }
}


// init, lane-contiguous (n = 128 = 4 per lane, rows 16-byte aligned): 128-bit loads / stores, neighbours by shuffle, no shared
// memory.  Same fp32 operations in the same order as sampler_init_kernel (bit-identical z and beta); the general kernel spent
// 349 warp instructions per ray on 1 KB of traffic (issue slots 67 % busy at 22 % of the DRAM peak).
__global__ void __launch_bounds__(256)
sampler_init128_kernel(svs_sampler_cfg c, int64_t R, const float* __restrict__ t_lin, const float* __restrict__ t_rand,
                       float* __restrict__ z, float* __restrict__ beta) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float4 t4 = __ldg(reinterpret_cast<const float4*>(t_lin) + lane);
  const float t[4] = {t4.x, t4.y, t4.z, t4.w};
  float zu[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) zu[j] = c.near * (1.0f - t[j]) + c.far * t[j];
  // the uniform row is the same for every ray (fixed far): its neighbours are fetched once
  const float zu_prev = __shfl_up_sync(0xffffffffu, zu[3], 1), zu_next = __shfl_down_sync(0xffffffffu, zu[0], 1);
  for (int64_t ray = warp0; ray < R; ray += nwarps) {
    float zj[4];
    if (t_rand) {
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(t_rand + ray * 128) + lane);
      const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = lane * 4 + j;
        const float zi = zu[j];
        const float zp = j ? zu[j ? j - 1 : 0] : zu_prev, zn = (j < 3) ? zu[(j + 1) & 3] : zu_next;
        const float lower = (i == 0) ? zi : 0.5f * (zi + zp);
        const float upper = (i == 127) ? zi : 0.5f * (zn + zi);
        zj[j] = lower + (upper - lower) * rr[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) zj[j] = zu[j];
    }
    *(reinterpret_cast<float4*>(z + ray * 128) + lane) = make_float4(zj[0], zj[1], zj[2], zj[3]);
    const float zj_next = __shfl_down_sync(0xffffffffu, zj[0], 1);
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = lane * 4 + j;
      if (i < 127) {
        const float d = ((j < 3) ? zj[(j + 1) & 3] : zj_next) - zj[j];
        acc += (double)(d * d);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) beta[ray] = sqrtf(c.inv4logeps * (float)acc);
  }
}

// ----------------------------------------------------------------------------------------------------
// error bound for one beta (ray_sampler.py:221-229).  Row data in smem: sdf[n], dist[n-1], dstar[n-1].
// ----------------------------------------------------------------------------------------------------
template <bool X>
__device__ __forceinline__ float error_bound(const float* ssdf, const float* sdist, const float* sdstar,
                                             int n, float beta, int lane) {
  typedef typename Acc<X>::type acc_t;
  acc_t carry_i = 0, carry_e = 0;  // running sums before this segment
  float best = -CUDART_INF_F;
  const float four_b2 = 4.0f * (beta * beta);
  for (int base = 0; base < n - 1; base += 32) {
    int i = base + lane;
    float sfe = 0.f, es = 0.f;
    bool valid = i < n - 1;
    if (valid) {
      float d = sdist[i];
      sfe = d * laplace<X>(ssdf[i], beta);
      es = (t_exp<X>(-sdstar[i] / beta) * (d * d)) / four_b2;
    }
    acc_t incl_i = warp_incl_scan<acc_t>((acc_t)sfe, lane);
    acc_t incl_e = warp_incl_scan<acc_t>((acc_t)es, lane);
    float integral = (float)(carry_i + excl_of(incl_i, lane));  // exclusive prefix, rounded
    float eint = (float)(carry_e + incl_e);                   // inclusive prefix, rounded
    if (valid) {
      float bo = (fminf(t_exp<X>(eint), 1.0e6f) - 1.0f) * t_exp<X>(-integral);
      best = fmaxf(best, bo);
    }
    carry_i += __shfl_sync(0xffffffffu, incl_i, 31);
    carry_e += __shfl_sync(0xffffffffu, incl_e, 31);
  }
  return warp_max(best);
}

// The same bound in plain fp32 with MUFU exponentials (relative error <= ~1e-4 near the threshold: ex2.approx 2^-22,
// fp32 scans over <= 1024 terms).  The line search only needs the outcome of `bound <= eps`, so this value decides
// whenever it is further than kBoundBand (relative, 5e-4; round 1: 2e-3) from eps and the canonical fp64 evaluation is run only inside
// the band: the beta sequence stays bit-identical to the canonical arithmetic at a fraction of its cost (the
// canonical evaluation is ~400 fp64 instructions per sample: exp, expm1 and two exps of the prefixes).
constexpr float kBoundBand = 5e-4f;   // fast-path error: ex2.approx 2^-22 per exponential, fp32 prefix sums over <= 1024 terms: <= ~1e-4
__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float error_bound_fast(const float* ssdf, const float* sdist, const float* sdstar, int n,
                                                  float beta, int lane) {
  const float kL2e = 1.4426950408889634f;
  const float inv_b = 1.0f / beta, k2 = -kL2e * inv_b, q = 0.25f * inv_b * inv_b;
  float carry_i = 0.f, carry_e = 0.f, best = -CUDART_INF_F;
  for (int base = 0; base < n - 1; base += 32) {
    const int i = base + lane;
    const bool valid = i < n - 1;
    float sfe = 0.f, es = 0.f;
    if (valid) {
      const float d = sdist[i], sv = ssdf[i];
      const float e = ex2f(fabsf(sv) * k2);   // exp(-|s| / beta)
      const float sigma = inv_b * ((sv > 0.f) ? 0.5f * e : ((sv < 0.f) ? 1.0f - 0.5f * e : 0.5f));
      sfe = d * sigma;
      es = ex2f(sdstar[i] * k2) * (d * d) * q;
    }
    const float incl_i = warp_incl_scan<float>(sfe, lane);
    const float incl_e = warp_incl_scan<float>(es, lane);
    const float integral = carry_i + excl_of(incl_i, lane);
    const float eint = carry_e + incl_e;
    if (valid) best = fmaxf(best, (fminf(ex2f(eint * kL2e), 1.0e6f) - 1.0f) * ex2f(-integral * kL2e));
    carry_i += __shfl_sync(0xffffffffu, incl_i, 31);
    carry_e += __shfl_sync(0xffffffffu, incl_e, 31);
  }
  return warp_max(best);
}
// value to compare with eps: the fast bound when the comparison is certain, else the canonical one (X only)
template <bool X>
__device__ __forceinline__ float bound_for_decision(const float* ssdf, const float* sdist, const float* sdstar, int n,
                                                    float beta, float eps, int lane) {
  if (X) {
    const float ef = error_bound_fast(ssdf, sdist, sdstar, n, beta, lane);
    if (fabsf(ef - eps) > kBoundBand * eps) return ef;   // false for NaN: falls through to the canonical evaluation
  }
  return error_bound<X>(ssdf, sdist, sdstar, n, beta, lane);
}

// ----------------------------------------------------------------------------------------------------
// bound: sdf merge + d* + beta line search (ray_sampler.py:90-123,136)
// ----------------------------------------------------------------------------------------------------
template <bool X>
__global__ void __launch_bounds__(kSampWarps * 32)
sampler_bound_kernel(svs_sampler_cfg c, int64_t R, int n, int n_new, const float* __restrict__ z,
                     const float* __restrict__ sdf_old, const float* __restrict__ sdf_new,
                     const int32_t* __restrict__ samples_idx, float* __restrict__ sdf_out,
                     const float* __restrict__ beta_param, float beta_min, float* __restrict__ beta,
                     int32_t* __restrict__ not_converged) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ssdf = smem + warp * (4 * n);
  float* sz = ssdf + n;
  float* sdist = sz + n;
  float* sdstar = sdist + n;
  const float beta0 = fabsf(__ldg(beta_param)) + beta_min;
  const int n_old = n - n_new;
  bool any_nc = false;
  for (int64_t ray = blockIdx.x * (int64_t)kSampWarps + warp; ray < R; ray += (int64_t)gridDim.x * kSampWarps) {
    for (int i = lane; i < n; i += 32) {
      float s;
      if (samples_idx) {
        int src = samples_idx[ray * n + i];
        s = (src < n_old) ? sdf_old[ray * n_old + src] : sdf_new[ray * n_new + (src - n_old)];
      } else {
        s = sdf_new[ray * n_new + i];
      }
      ssdf[i] = s;
      sdf_out[ray * n + i] = s;
      sz[i] = z[ray * n + i];
    }
    __syncwarp();
    for (int i = lane; i < n - 1; i += 32) {
      float d = sz[i + 1] - sz[i];
      sdist[i] = d;
      sdstar[i] = d_star_of(d, ssdf[i], ssdf[i + 1]);
    }
    __syncwarp();
    float b = beta[ray];
    float err = bound_for_decision<X>(ssdf, sdist, sdstar, n, beta0, c.eps, lane);
    if (err <= c.eps) b = beta0;
    float bmin = beta0, bmax = b;
    for (int j = 0; j < c.beta_iters; ++j) {
      float mid = (bmin + bmax) / 2.0f;
      err = bound_for_decision<X>(ssdf, sdist, sdstar, n, mid, c.eps, lane);
      if (err <= c.eps) bmax = mid;
      if (err > c.eps) bmin = mid;
    }
    if (lane == 0) beta[ray] = bmax;
    any_nc |= (bmax > beta0);
    __syncwarp();
  }
  if (lane == 0 && any_nc) atomicOr(not_converged, 1);
}


// ----------------------------------------------------------------------------------------------------
// bound, lane-contiguous layout (n <= 1024).  Lane L owns the C consecutive samples C L .. C L + C - 1, so a prefix sum
// over the row is a serial prefix inside the lane plus ONE 5-step warp scan of the lane totals (the segment layout above
// runs n / 32 dependent scans per quantity and evaluation: 56 shuffles at n = 128, here 17).  Rows sit in shared memory
// at index (i / C) * (C + 1) + i % C (odd stride: conflict-free for the even C used); for C <= 8 the pre-scaled row is
// held in registers over the 11 evaluations.  The canonical fp64 evaluation (only inside the decision band) reads the
// same rows in the original segment order, so the beta sequence stays bit-identical to the oracle's.
// ----------------------------------------------------------------------------------------------------
constexpr int kBoundWarps = 8;
constexpr float kLog2e = 1.4426950408889634f;

template <int C>
__device__ __forceinline__ int pidx(int i) { return (i / C) * (C + 1) + (i % C); }

template <int C>
__device__ __forceinline__ float error_bound_padded(const float* sS, const float* sD, const float* sX, int n, float beta, int lane) {
  double carry_i = 0, carry_e = 0;
  float best = -CUDART_INF_F;
  const float four_b2 = 4.0f * (beta * beta);
  for (int base = 0; base < n - 1; base += 32) {
    int i = base + lane;
    float sfe = 0.f, es = 0.f;
    bool valid = i < n - 1;
    if (valid) {
      float d = sD[pidx<C>(i)];
      sfe = d * laplace<true>(sS[pidx<C>(i)], beta);
      es = (t_exp<true>(-sX[pidx<C>(i)] / beta) * (d * d)) / four_b2;
    }
    double incl_i = warp_incl_scan<double>((double)sfe, lane);
    double incl_e = warp_incl_scan<double>((double)es, lane);
    float integral = (float)(carry_i + excl_of(incl_i, lane));
    float eint = (float)(carry_e + incl_e);
    if (valid) {
      float bo = (fminf(t_exp<true>(eint), 1.0e6f) - 1.0f) * t_exp<true>(-integral);
      best = fmaxf(best, bo);
    }
    carry_i += __shfl_sync(0xffffffffu, incl_i, 31);
    carry_e += __shfl_sync(0xffffffffu, incl_e, 31);
  }
  return warp_max(best);
}

// pre-scaled row of one lane: a = -log2e |s|, hs = sgn(s) / 2, d = dist, dq = dist^2 / 4, x = -log2e d*
template <int C>
struct LaneRow {
  float a[C], hs[C], d[C], dq[C], x[C];
};
template <int C>
__device__ __forceinline__ void lane_row_load(LaneRow<C>& r, const float* sS, const float* sD, const float* sX, int lane) {
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const int o = lane * (C + 1) + j;
    const float sv = sS[o], d = sD[o];
    r.a[j] = -kLog2e * fabsf(sv);
    r.hs[j] = 0.5f * sgn(sv);
    r.d[j] = d;
    r.dq[j] = 0.25f * (d * d);
    r.x[j] = -kLog2e * sX[o];
  }
}

// (min(exp(E_i), 1e6) - 1) exp(-I_i) maximised over the row, plain fp32 + MUFU: decides `bound <= eps` outside the band
template <int C, bool REG>
__device__ __forceinline__ float error_bound_lanes(const LaneRow<C>& row, const float* sS, const float* sD, const float* sX,
                                                   float beta, int lane) {
  const float inv_b = 1.0f / beta, q = inv_b * inv_b;
  float Fi[C], Gi[C];
  float F = 0.f, G = 0.f;
#pragma unroll
  for (int j = 0; j < C; ++j) {
    float a, hs, d, dq, x;
    if (REG) {
      a = row.a[j]; hs = row.hs[j]; d = row.d[j]; dq = row.dq[j]; x = row.x[j];
    } else {
      const int o = lane * (C + 1) + j;
      const float sv = sS[o];
      d = sD[o];
      a = -kLog2e * fabsf(sv);
      hs = 0.5f * sgn(sv);
      dq = 0.25f * (d * d);
      x = -kLog2e * sX[o];
    }
    const float e = ex2f(a * inv_b);                          // exp(-|s| / beta)
    const float sig = inv_b * fmaf(hs, e - 1.0f, 0.5f);       // Laplace CDF density
    F = fmaf(d, sig, F);
    Fi[j] = F;
    G = fmaf(ex2f(x * inv_b) * dq, q, G);
    Gi[j] = G;
  }
  float incF = F, incG = G;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float tf = __shfl_up_sync(0xffffffffu, incF, o), tg = __shfl_up_sync(0xffffffffu, incG, o);
    if (lane >= o) { incF += tf; incG += tg; }
  }
  float baseF = __shfl_up_sync(0xffffffffu, incF, 1), baseG = __shfl_up_sync(0xffffffffu, incG, 1);
  if (lane == 0) { baseF = 0.f; baseG = 0.f; }
  // intervals beyond n - 2 carry dist = 0: they repeat the last valid prefix with a larger integral and never win the max
  float best = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const float I = baseF + (j ? Fi[j - 1] : 0.f), E = baseG + Gi[j];
    best = fmaxf(best, (fminf(ex2f(E * kLog2e), 1.0e6f) - 1.0f) * ex2f(-I * kLog2e));
  }
  return warp_max(best);
}

template <bool X, int C>
__global__ void __launch_bounds__(kBoundWarps * 32)
sampler_bound_lanes_kernel(svs_sampler_cfg c, int64_t R, int n, int n_new, const float* __restrict__ z,
                           const float* __restrict__ sdf_old, const float* __restrict__ sdf_new,
                           const int32_t* __restrict__ samples_idx, float* __restrict__ sdf_out,
                           const float* __restrict__ beta_param, float beta_min, float* __restrict__ beta,
                           int32_t* __restrict__ not_converged) {
  constexpr int ROW = 32 * (C + 1);
  constexpr bool REG = C <= 8;
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sS = smem + warp * (4 * ROW);
  float* sZ = sS + ROW;
  float* sD = sZ + ROW;
  float* sX = sD + ROW;
  const float beta0 = fabsf(__ldg(beta_param)) + beta_min;
  const int n_old = n - n_new;
  bool any_nc = false;
  for (int64_t ray = blockIdx.x * (int64_t)kBoundWarps + warp; ray < R; ray += (int64_t)gridDim.x * kBoundWarps) {
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const int i = lane + 32 * k;
      float s = 0.f, zz = 0.f;
      if (i < n) {
        if (samples_idx) {
          int src = samples_idx[ray * n + i];
          s = (src < n_old) ? sdf_old[ray * n_old + src] : sdf_new[ray * n_new + (src - n_old)];
        } else {
          s = sdf_new[ray * n_new + i];
        }
        sdf_out[ray * n + i] = s;
        zz = z[ray * n + i];
      }
      sS[pidx<C>(i)] = s;
      sZ[pidx<C>(i)] = zz;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const int i = lane + 32 * k;
      float d = 0.f, ds = 0.f;
      if (i < n - 1) {
        d = sZ[pidx<C>(i + 1)] - sZ[pidx<C>(i)];
        ds = d_star_of(d, sS[pidx<C>(i)], sS[pidx<C>(i + 1)]);
      }
      sD[pidx<C>(i)] = d;
      sX[pidx<C>(i)] = ds;
    }
    __syncwarp();
    LaneRow<REG ? C : 1> row;
    if (REG) lane_row_load<REG ? C : 1>(row, sS, sD, sX, lane);
    auto decide = [&](float bt) -> float {
      float ef;
      if constexpr (REG) ef = error_bound_lanes<C, true>(row, sS, sD, sX, bt, lane);
      else ef = error_bound_lanes<C, false>(LaneRow<C>(), sS, sD, sX, bt, lane);
      if (!X || fabsf(ef - c.eps) > kBoundBand * c.eps) return ef;   // NaN: falls through to the canonical evaluation
      return error_bound_padded<C>(sS, sD, sX, n, bt, lane);
    };
    float b = beta[ray];
    float err = decide(beta0);
    if (err <= c.eps) b = beta0;
    float bmin = beta0, bmax = b;
    if (bmin != bmax) {   // bmin == bmax: every midpoint is beta0 again and the outcome is known
      for (int j = 0; j < c.beta_iters; ++j) {
        float mid = (bmin + bmax) / 2.0f;
        err = decide(mid);
        if (err <= c.eps) bmax = mid;
        if (err > c.eps) bmin = mid;
      }
    }
    if (lane == 0) beta[ray] = bmax;
    any_nc |= (bmax > beta0);
    __syncwarp();
  }
  if (lane == 0 && any_nc) atomicOr(not_converged, 1);
}

template <bool X, int C>
static int launch_bound_lanes(const svs_sampler_cfg* c, int64_t R, int n, int n_new, const float* z, const float* sdf_old,
                              const float* sdf_new, const int32_t* samples_idx, float* sdf, const float* beta_param,
                              float beta_min, float* beta, int32_t* not_converged, cudaStream_t st) {
  size_t smem = (size_t)kBoundWarps * 4 * 32 * (C + 1) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    SVS_CUDA_OK(cudaFuncSetAttribute(sampler_bound_lanes_kernel<X, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int64_t blocks = cdiv(R, kBoundWarps);
  int64_t cap = (int64_t)kNumSMs * (C <= 8 ? 6 : 2);
  sampler_bound_lanes_kernel<X, C><<<(int)(blocks < cap ? blocks : cap), kBoundWarps * 32, smem, st>>>(
      *c, R, n, n_new, z, sdf_old, sdf_new, samples_idx, sdf, beta_param, beta_min, beta, not_converged);
  return SVS_OK;
}
template <bool X>
static int launch_bound_lanes_x(int C, const svs_sampler_cfg* c, int64_t R, int n, int n_new, const float* z, const float* sdf_old,
                                const float* sdf_new, const int32_t* samples_idx, float* sdf, const float* beta_param,
                                float beta_min, float* beta, int32_t* not_converged, cudaStream_t st) {
#define SVS_BOUND_CASE(CC) \
  if (C <= CC) return launch_bound_lanes<X, CC>(c, R, n, n_new, z, sdf_old, sdf_new, samples_idx, sdf, beta_param, beta_min, beta, not_converged, st)
  SVS_BOUND_CASE(4);
  SVS_BOUND_CASE(8);
  SVS_BOUND_CASE(12);
  SVS_BOUND_CASE(16);
  SVS_BOUND_CASE(20);
  SVS_BOUND_CASE(24);
  SVS_BOUND_CASE(32);
#undef SVS_BOUND_CASE
  return SVS_ERR_INVALID;
}

// ----------------------------------------------------------------------------------------------------
// resample: weights/transmittance, pdf, cdf, inverse CDF, stable merge (ray_sampler.py:126-190)
// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool key_less(float va, int ia, float vb, int ib) {
  return (va < vb) || (va == vb && ia < ib);
}

// in-warp bitonic sort of m (power of two) (value, index) pairs held in shared memory
__device__ __forceinline__ void bitonic_sort_pairs(float* v, int* id, int m, int lane) {
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < m / 2; t += 32) {
        int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int hi = lo | j;
        bool up = ((lo & k) == 0);
        float a = v[lo], b = v[hi];
        int ia = id[lo], ib = id[hi];
        bool swap = up ? key_less(b, ib, a, ia) : key_less(a, ia, b, ib);
        if (swap) {
          v[lo] = b;
          v[hi] = a;
          id[lo] = ib;
          id[hi] = ia;
        }
      }
      __syncwarp();
    }
  }
}

template <bool X>
__global__ void __launch_bounds__(kSampWarps * 32)
sampler_resample_kernel(svs_sampler_cfg c, int64_t R, int n, int n_u, int n_u_pad, int cont,
                        const float* __restrict__ z, const float* __restrict__ sdf,
                        const float* __restrict__ beta, const float* __restrict__ u, int u_per_ray,
                        float* __restrict__ samples, int32_t* __restrict__ inds_out,
                        float* __restrict__ z_merged, int32_t* __restrict__ samples_idx) {
  typedef typename Acc<X>::type acc_t;
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = 4 * n + 2 * n_u_pad;
  float* sz = smem + warp * per_warp;
  float* ssdf = sz + n;
  float* sT = ssdf + n;
  float* scdf = sT + n;      // pdf, then cdf
  float* sval = scdf + n;    // new samples (n_u_pad)
  int* sid = (int*)(sval + n_u_pad);
  for (int64_t ray = blockIdx.x * (int64_t)kSampWarps + warp; ray < R; ray += (int64_t)gridDim.x * kSampWarps) {
    for (int i = lane; i < n; i += 32) {
      sz[i] = z[ray * n + i];
      ssdf[i] = sdf[ray * n + i];
    }
    __syncwarp();
    const float b = beta[ray];
    const float four_b2 = 4.0f * (b * b);
    // pass 1: transmittance T_i = exp(-sum_{j<i} dist_j sigma_j), pdf (un-normalised) for i < n-1
    acc_t carry_f = 0, carry_e = 0, psum = 0;
    for (int base = 0; base < n; base += 32) {
      int i = base + lane;
      float fe = 0.f, es = 0.f, d = 0.f;
      if (i < n) {
        d = (i < n - 1) ? (sz[i + 1] - sz[i]) : 1e10f;
        fe = d * laplace<X>(ssdf[i], b);
        if (cont && i < n - 1) {
          float ds = d_star_of(d, ssdf[i], ssdf[i + 1]);
          es = (t_exp<X>(-ds / b) * (d * d)) / four_b2;
        }
      }
      acc_t incl_f = warp_incl_scan<acc_t>((acc_t)fe, lane);
      float T = t_exp<X>(-(float)(carry_f + excl_of(incl_f, lane)));
      float p = 0.f;
      if (cont) {
        acc_t incl_e = warp_incl_scan<acc_t>((acc_t)es, lane);
        float eint = (float)(carry_e + incl_e);
        p = (fminf(t_exp<X>(eint), 1.0e6f) - 1.0f) * T + c.add_tiny;
        carry_e += __shfl_sync(0xffffffffu, incl_e, 31);
      } else {
        float alpha = 1.0f - t_exp<X>(-fe);
        p = alpha * T + 1e-5f;
      }
      if (i >= n - 1) p = 0.f;
      if (i < n) {
        sT[i] = T;
        scdf[i] = p;
      }
      psum += (acc_t)p;
      carry_f += __shfl_sync(0xffffffffu, incl_f, 31);
    }
    float total;
    {
      acc_t t = psum;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      total = (float)t;
    }
    __syncwarp();
    // pass 2: cdf[0] = 0, cdf[i] = round(sum_{j<i} pdf_j / total)
    acc_t carry_c = 0;
    for (int base = 0; base < n; base += 32) {
      int i = base + lane;
      float p = (i < n - 1) ? (scdf[i] / total) : 0.f;
      acc_t incl = warp_incl_scan<acc_t>((acc_t)p, lane);
      float cexcl = (float)(carry_c + excl_of(incl, lane));
      __syncwarp();
      if (i < n) scdf[i] = cexcl;
      carry_c += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    // inverse CDF (ray_sampler.py:173-185)
    for (int j = lane; j < n_u_pad; j += 32) {
      float sv = CUDART_INF_F;
      if (j < n_u) {
        float uu = u_per_ray ? u[ray * n_u + j] : u[j];
        int lo = 0, hi = n;  // count of cdf entries <= uu  (searchsorted right=True)
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (scdf[mid] <= uu) lo = mid + 1; else hi = mid;
        }
        int ind = lo;
        int below = max(ind - 1, 0), above = min(ind, n - 1);
        float cb = scdf[below], ca = scdf[above];
        float zb = sz[below], za = sz[above];
        float denom = ca - cb;
        if (denom < 1e-5f) denom = 1.0f;
        float t = (uu - cb) / denom;
        sv = zb + t * (za - zb);
        samples[ray * n_u + j] = sv;
        if (inds_out) inds_out[ray * n_u + j] = ind;
      }
      sval[j] = sv;
      sid[j] = j;
    }
    __syncwarp();
    if (cont) {
      // z, samples_idx = sort(cat[z, samples]) (ray_sampler.py:189-190), stable: old before new on ties
      bitonic_sort_pairs(sval, sid, n_u_pad, lane);
      const int nm = n + n_u;
      for (int i = lane; i < n; i += 32) {  // old sample i: rank among new = #{new < z_i}
        float zi = sz[i];
        int lo = 0, hi = n_u;
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (sval[mid] < zi) lo = mid + 1; else hi = mid;
        }
        z_merged[ray * nm + i + lo] = zi;
        samples_idx[ray * nm + i + lo] = i;
      }
      for (int r = lane; r < n_u; r += 32) {  // new sample of sorted rank r: #{old <= s}
        float s = sval[r];
        int lo = 0, hi = n;
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (sz[mid] <= s) lo = mid + 1; else hi = mid;
        }
        z_merged[ray * nm + r + lo] = s;
        samples_idx[ray * nm + r + lo] = n + sid[r];
      }
    }
    __syncwarp();
  }
}

// ----------------------------------------------------------------------------------------------------
// finalize (ray_sampler.py:193-212)
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSampWarps * 32)
sampler_finalize_kernel(svs_sampler_cfg c, int64_t R, int n, int n_samples, int m_pad,
                        const float* __restrict__ z, const float* __restrict__ samples,
                        const int32_t* __restrict__ extra_idx, int n_extra, const float* __restrict__ far_ray,
                        const int64_t* __restrict__ eik_idx, float* __restrict__ z_final,
                        float* __restrict__ z_eik) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sv = smem + warp * 2 * m_pad;
  int* sid = (int*)(sv + m_pad);
  const int m = n_samples + 2 + n_extra;
  for (int64_t ray = blockIdx.x * (int64_t)kSampWarps + warp; ray < R; ray += (int64_t)gridDim.x * kSampWarps) {
    const float far = (c.far < 0.f) ? far_ray[ray] : c.far;
    for (int i = lane; i < m_pad; i += 32) {
      float v = CUDART_INF_F;
      if (i < n_samples) v = samples[ray * n_samples + i];
      else if (i == n_samples) v = c.near;
      else if (i == n_samples + 1) v = far;
      else if (i < m) v = z[ray * n + extra_idx[i - n_samples - 2]];
      sv[i] = v;
      sid[i] = i;
    }
    __syncwarp();
    bitonic_sort_pairs(sv, sid, m_pad, lane);
    for (int i = lane; i < m; i += 32) z_final[ray * m + i] = sv[i];
    if (lane == 0 && z_eik) z_eik[ray] = sv[(int)eik_idx[ray]];
    __syncwarp();
  }
}


// finalize with the row in registers (m_pad = 32 C <= 256): C keys per lane, key e = C lane + r.  The reference discards
// the sort indices (ray_sampler.py:208), so plain min / max compare-exchanges sort the values; partners at distance < C
// are in the same lane, the others one shuffle away.  (The shared-memory bitonic sort above spent 0.94 ms on 262144 rays,
// a third of the whole sampler iteration.)
template <int C>
__global__ void __launch_bounds__(kSampWarps * 32)
sampler_finalize_regs_kernel(svs_sampler_cfg c, int64_t R, int n, int n_samples, const float* __restrict__ z,
                             const float* __restrict__ samples, const int32_t* __restrict__ extra_idx, int n_extra,
                             const float* __restrict__ far_ray, const int64_t* __restrict__ eik_idx,
                             float* __restrict__ z_final, float* __restrict__ z_eik) {
  constexpr int M = 32 * C;
  __shared__ float srow[kSampWarps][M + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = n_samples + 2 + n_extra;
  for (int64_t ray = blockIdx.x * (int64_t)kSampWarps + warp; ray < R; ray += (int64_t)gridDim.x * kSampWarps) {
    const float far = (c.far < 0.f) ? far_ray[ray] : c.far;
    // coalesced gather into shared memory, then C consecutive keys per lane
    for (int i = lane; i < M; i += 32) {
      float v = CUDART_INF_F;
      if (i < n_samples) v = samples[ray * n_samples + i];
      else if (i == n_samples) v = c.near;
      else if (i == n_samples + 1) v = far;
      else if (i < m) v = z[ray * n + extra_idx[i - n_samples - 2]];
      srow[warp][i] = v;
    }
    __syncwarp();
    float v[C];
#pragma unroll
    for (int r = 0; r < C; ++r) v[r] = srow[warp][lane * C + r];
#pragma unroll
    for (int k = 2; k <= M; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        if (j >= C) {
          const int lj = j / C;                       // partner lane distance
#pragma unroll
          for (int r = 0; r < C; ++r) {
            const int e = lane * C + r;
            const float p = __shfl_xor_sync(0xffffffffu, v[r], lj);
            const bool up = (e & k) == 0, low = (e & j) == 0;   // low element of an ascending pair keeps the minimum
            v[r] = (up == low) ? fminf(v[r], p) : fmaxf(v[r], p);
          }
        } else {
#pragma unroll
          for (int r = 0; r < C; ++r) {
            if ((r & j) == 0) {
              const int e = lane * C + r;
              const bool up = (e & k) == 0;
              const float a = v[r], b = v[r | j];
              v[r] = up ? fminf(a, b) : fmaxf(a, b);
              v[r | j] = up ? fmaxf(a, b) : fminf(a, b);
            }
          }
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < C; ++r) srow[warp][lane * C + r] = v[r];
    __syncwarp();
    for (int i = lane; i < m; i += 32) z_final[ray * m + i] = srow[warp][i];
    if (lane == 0 && z_eik) z_eik[ray] = srow[warp][(int)eik_idx[ray]];
    __syncwarp();
  }
}

static int samp_grid(int64_t R) {
  int64_t blocks = cdiv(R, kSampWarps);
  int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(blocks < cap ? blocks : cap);
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    SVS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  return SVS_OK;
}

}  // namespace svs

using namespace svs;

extern "C" int svs_sampler_init(const svs_sampler_cfg* c, int64_t R, int32_t n, const float* t_lin,
                                const float* t_rand, const float* far_ray, float* z, float* beta, void* stream) {
  SVS_CHECK_ARG(c && R >= 0 && n >= 2 && n <= 1024, "svs_sampler_init: need 2 <= n <= 1024 (got %d)", n);
  SVS_CHECK_ARG(t_lin && z && beta, "svs_sampler_init: null pointer");
  SVS_CHECK_ARG(c->far >= 0.f || far_ray, "svs_sampler_init: far_ray required when cfg.far < 0");
  if (R == 0) return SVS_OK;
  size_t smem = (size_t)kSampWarps * n * sizeof(float);
  ProfScope ps("sampler_init", 0.0, (double)R * (4.0 * n * (t_rand ? 2 : 1) + 4), (cudaStream_t)stream);
  if (n == 128 && c->far >= 0.f &&
      ((reinterpret_cast<uintptr_t>(t_lin) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(t_rand ? t_rand : z)) & 15) == 0) {
    const int64_t blocks = cdiv(R, 8);
    const int64_t cap = (int64_t)kNumSMs * 8;
    sampler_init128_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(*c, R, t_lin, t_rand, z, beta);
    SVS_LAUNCH_OK();
    return SVS_OK;
  }
  sampler_init_kernel<<<samp_grid(R), kSampWarps * 32, smem, (cudaStream_t)stream>>>(*c, R, n, t_lin, t_rand,
                                                                                    far_ray, z, beta);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

extern "C" int svs_sampler_bound(const svs_sampler_cfg* c, int64_t R, int32_t n, int32_t n_new, const float* z,
                                 const float* sdf_old, const float* sdf_new, const int32_t* samples_idx,
                                 float* sdf, const float* beta_param, float beta_min, float* beta,
                                 int32_t* not_converged, void* stream) {
  SVS_CHECK_ARG(c && R >= 0 && n >= 2 && n <= 4096 && n_new >= 1 && n_new <= n, "svs_sampler_bound: bad n=%d n_new=%d", n, n_new);
  SVS_CHECK_ARG(z && sdf_new && sdf && beta_param && beta && not_converged, "svs_sampler_bound: null pointer");
  SVS_CHECK_ARG((n_new == n) == (samples_idx == nullptr), "svs_sampler_bound: samples_idx iff n_new < n");
  SVS_CHECK_ARG(n_new == n || sdf_old, "svs_sampler_bound: sdf_old required when merging");
  if (R == 0) return SVS_OK;
  size_t smem = (size_t)kSampWarps * 4 * n * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps("sampler_bound", 0.0, (double)R * (4.0 * n * (samples_idx ? 4 : 3) + 8), st);
  if (n <= 1024) {   // lane-contiguous kernel; the segment kernel below serves longer rows
    const int C = (n + 31) / 32;
    if (c->exact)
      SVS_TRY(launch_bound_lanes_x<true>(C, c, R, n, n_new, z, sdf_old, sdf_new, samples_idx, sdf, beta_param, beta_min, beta, not_converged, st));
    else
      SVS_TRY(launch_bound_lanes_x<false>(C, c, R, n, n_new, z, sdf_old, sdf_new, samples_idx, sdf, beta_param, beta_min, beta, not_converged, st));
    SVS_LAUNCH_OK();
    return SVS_OK;
  }
  if (c->exact) {
    SVS_TRY(set_smem(sampler_bound_kernel<true>, smem));
    sampler_bound_kernel<true><<<samp_grid(R), kSampWarps * 32, smem, st>>>(
        *c, R, n, n_new, z, sdf_old, sdf_new, samples_idx, sdf, beta_param, beta_min, beta, not_converged);
  } else {
    SVS_TRY(set_smem(sampler_bound_kernel<false>, smem));
    sampler_bound_kernel<false><<<samp_grid(R), kSampWarps * 32, smem, st>>>(
        *c, R, n, n_new, z, sdf_old, sdf_new, samples_idx, sdf, beta_param, beta_min, beta, not_converged);
  }
  SVS_LAUNCH_OK();
  return SVS_OK;
}

extern "C" int svs_sampler_resample(const svs_sampler_cfg* c, int64_t R, int32_t n, int32_t n_u, int32_t cont,
                                    const float* z, const float* sdf, const float* beta, const float* u,
                                    int32_t u_per_ray, float* samples, int32_t* inds, float* z_merged,
                                    int32_t* samples_idx, void* stream) {
  SVS_CHECK_ARG(c && R >= 0 && n >= 2 && n <= 4096 && n_u >= 1 && n_u <= 1024, "svs_sampler_resample: bad n=%d n_u=%d", n, n_u);
  SVS_CHECK_ARG(z && sdf && beta && u && samples, "svs_sampler_resample: null pointer");
  SVS_CHECK_ARG(!cont || (z_merged && samples_idx), "svs_sampler_resample: merge outputs required when cont");
  if (R == 0) return SVS_OK;
  int n_u_pad = next_pow2(n_u < 32 ? 32 : n_u);
  size_t smem = (size_t)kSampWarps * (4 * n + 2 * n_u_pad) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps("sampler_resample", 0.0,
               (double)R * (8.0 * n + 4 + 4.0 * n_u * (u_per_ray ? 2 : 1) + (cont ? 8.0 * (n + n_u) : 0.0)), st);
  if (c->exact) {
    SVS_TRY(set_smem(sampler_resample_kernel<true>, smem));
    sampler_resample_kernel<true><<<samp_grid(R), kSampWarps * 32, smem, st>>>(
        *c, R, n, n_u, n_u_pad, cont, z, sdf, beta, u, u_per_ray, samples, inds, z_merged, samples_idx);
  } else {
    SVS_TRY(set_smem(sampler_resample_kernel<false>, smem));
    sampler_resample_kernel<false><<<samp_grid(R), kSampWarps * 32, smem, st>>>(
        *c, R, n, n_u, n_u_pad, cont, z, sdf, beta, u, u_per_ray, samples, inds, z_merged, samples_idx);
  }
  SVS_LAUNCH_OK();
  return SVS_OK;
}

extern "C" int svs_sampler_finalize(const svs_sampler_cfg* c, int64_t R, int32_t n, int32_t n_samples,
                                    const float* z, const float* samples, const int32_t* extra_idx,
                                    int32_t n_extra, const float* far_ray, const int64_t* eik_idx,
                                    float* z_final, float* z_eik, void* stream) {
  SVS_CHECK_ARG(c && R >= 0 && n >= 2 && n_samples >= 1 && n_extra >= 0, "svs_sampler_finalize: bad sizes");
  SVS_CHECK_ARG(z && samples && z_final && (n_extra == 0 || extra_idx), "svs_sampler_finalize: null pointer");
  SVS_CHECK_ARG(!z_eik || eik_idx, "svs_sampler_finalize: eik_idx required for z_eik");
  SVS_CHECK_ARG(c->far >= 0.f || far_ray, "svs_sampler_finalize: far_ray required when cfg.far < 0");
  int m = n_samples + 2 + n_extra;
  SVS_CHECK_ARG(m <= 1024, "svs_sampler_finalize: too many final samples (%d)", m);
  if (R == 0) return SVS_OK;
  int m_pad = next_pow2(m < 64 ? 64 : m);
  size_t smem = (size_t)kSampWarps * 2 * m_pad * sizeof(float);
  ProfScope ps("sampler_finalize", 0.0, (double)R * (4.0 * (n_samples + n_extra) + 4.0 * m + 16), (cudaStream_t)stream);
  if (m_pad <= 256) {
    cudaStream_t st = (cudaStream_t)stream;
    if (m_pad == 64)
      sampler_finalize_regs_kernel<2><<<samp_grid(R), kSampWarps * 32, 0, st>>>(*c, R, n, n_samples, z, samples, extra_idx, n_extra, far_ray, eik_idx, z_final, z_eik);
    else if (m_pad == 128)
      sampler_finalize_regs_kernel<4><<<samp_grid(R), kSampWarps * 32, 0, st>>>(*c, R, n, n_samples, z, samples, extra_idx, n_extra, far_ray, eik_idx, z_final, z_eik);
    else
      sampler_finalize_regs_kernel<8><<<samp_grid(R), kSampWarps * 32, 0, st>>>(*c, R, n, n_samples, z, samples, extra_idx, n_extra, far_ray, eik_idx, z_final, z_eik);
    SVS_LAUNCH_OK();
    return SVS_OK;
  }
  sampler_finalize_kernel<<<samp_grid(R), kSampWarps * 32, smem, (cudaStream_t)stream>>>(
      *c, R, n, n_samples, m_pad, z, samples, extra_idx, n_extra, far_ray, eik_idx, z_final, z_eik);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

// K11: Laplace/Abs density + transmittance + alpha compositing, fused forward and fused backward.
// One warp per ray; the ray's rows (z, sdf, rgb[, normals]) are staged in shared memory with coalesced
// loads, then every lane owns a contiguous chunk of C samples so the exclusive prefix sum is one local
// scan + one 5-step warp scan.  HBM-bound: algorithmic bytes per ray in DESIGN.md §compositor.
//
// Reference: density.py:21-35, network.py:281-295 (volume_rendering), :239-248 (rgb/depth),
// :270-276 (normal map), network_bg.py:147-180 (fg tail + bg pass).  Math of the backward: SURVEY.md App. G.
#include <map>
#include <type_traits>

#include "svs_common.cuh"

namespace svs {

constexpr int kCompWarps = 4;
constexpr int kCompStages = 3;   // rows of the warp's next TWO rays are in flight while the current one is composited

// per-warp input rows of one ray (z has one pad element for the i + 1 access); `g` is the normals row (forward, eval)
// or the dL/dweights row (backward)
template <int C, int G>
struct CompRows {
  float z[32 * C + 4];
  float s[32 * C];
  float c[32 * C * 3];
  float g[(G > 0 ? 32 * C * G : 0) + 4];   // G floats per sample: 3 = normals (eval forward), 1 = dL/dweights (backward), 0 = unused;
                                           // every array is a multiple of 16 bytes: lanes read their chunk with 128-bit loads
};
// kCompStages stages per warp: the rows of the warp's next rays land (cp.async, no registers) while the current ray is
// being composited, so a warp's HBM latency overlaps its own arithmetic (one ray ahead left the kernel latency-bound:
// a ray took ~5.9 k cycles per warp of which ~0.5 k were arithmetic); + one output row
template <int C, int G>
struct CompSmem {
  CompRows<C, G> in[kCompWarps][kCompStages];
  float w[kCompWarps][32 * C];
  float o[kCompWarps][32 * C];
};

__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// issue the asynchronous copies of one ray's rows (coalesced 4-byte elements: the rows are only 4-byte aligned)
template <int C, int G>
__device__ __forceinline__ void issue_rows(CompRows<C, G>& st, const float* __restrict__ zr, const float* __restrict__ sr,
                                           const float* __restrict__ cr, const float* __restrict__ gr, int n_g, int S, int lane) {
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const int i = lane + 32 * j;
    if (i < S) {
      cp_async4(st.z + i, zr + i);
      cp_async4(st.s + i, sr + i);
    }
  }
  if (cr) {
#pragma unroll
    for (int j = 0; j < 3 * C; ++j) {
      const int i = lane + 32 * j;
      if (i < 3 * S) cp_async4(st.c + i, cr + i);
    }
  }
  if (G > 0 && gr) {
#pragma unroll
    for (int j = 0; j < (G > 0 ? G : 1) * C; ++j) {
      const int i = lane + 32 * j;
      if (i < n_g) cp_async4(st.g + i, gr + i);
    }
  }
  cp_async_commit();
}

// a lane's N consecutive floats of a staged row starting at row[first] (first % 4 == 0 when N % 4 == 0): 128-bit loads.
// A scalar read of row[lane * C + j] walks the banks with stride C — a 4-way conflict for the C = 4 of DTU rows (ncu: 15.5 M
// conflicts per launch, LSU pipe 34 % busy).
template <int N>
__device__ __forceinline__ void ld_chunk(const float* row, int first, float (&v)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 t = *reinterpret_cast<const float4*>(row + first + 4 * q);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = row[first + j];
  }
}
template <int N>
__device__ __forceinline__ void st_chunk(float* row, int first, const float (&v)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q)
      *reinterpret_cast<float4*>(row + first + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) row[first + j] = v[j];
  }
}

__device__ __forceinline__ float beta_of(const float* beta_param, float beta_min) {
  return fabsf(__ldg(beta_param)) + beta_min;  // density.py:28-30
}

// FAST (SVS_COMP_FAST; set by the models when the tcgen05 engine is selected, whose fp16 MLP error is 1e-3): MUFU
// exp, fp32 scans and reciprocal multiplies instead of the canonical arithmetic (libm expf / expm1f, fp64 prefix sums,
// IEEE divisions) that reproduces the oracle; weights differ by ~1e-6.  The exact path costs ~1400 instructions per
// ray and lane, which bounds it far below the HBM roofline; FAST is the bandwidth-bound variant.
template <bool FAST>
__device__ __forceinline__ float exp_t(float x) { return FAST ? __expf(x) : expf(x); }

// sigma and the autograd-form derivative d sigma / d s
template <bool FAST = false>
__device__ __forceinline__ float density_fwd(float s, float beta, bool abs_density, float* em_out) {
  if (abs_density) {
    *em_out = 0.f;
    return fabsf(s);
  }
  float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
  if (FAST) {
    // em = fl(e^-x - 1): for e^-x < 1/8 the MUFU error is below half the rounding quantum of (e - 1), so em lands on
    // the same fp32 value as expm1f -- which matters because the reference's last interval of 1e10 (network.py:286)
    // amplifies the rounding of em + 1 (sigma is exactly 0 once e^-x < 3e-8, as in the reference)
    const float ib = __frcp_rn(beta);
    const float em = __expf(-fabsf(s) * ib) - 1.0f;
    *em_out = em;
    return ib * (0.5f + 0.5f * sg * em);
  }
  float em = expm1f(-fabsf(s) / beta);
  *em_out = em;
  return (1.0f / beta) * (0.5f + 0.5f * sg * em);
}

// exclusive prefix sum over the warp's 32*C elements held as v[C] per lane (lane-contiguous layout);
// returns the grand total.  Accumulates in fp64 like torch's CPU cumsum.
template <int C, typename T = double>
__device__ __forceinline__ T warp_excl_scan(const float (&v)[C], T (&excl)[C], int lane) {
  T run = 0;
#pragma unroll
  for (int j = 0; j < C; ++j) {
    excl[j] = run;
    run += (T)v[j];
  }
  T incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  T base = __shfl_up_sync(0xffffffffu, incl, 1);  // previous lanes' total (not `incl - run`: cancels)
  if (lane == 0) base = 0;
#pragma unroll
  for (int j = 0; j < C; ++j) excl[j] += base;
  return __shfl_sync(0xffffffffu, incl, 31);
}

// suffix sums suf[j] = sum of the elements AFTER element (lane, j) in the lane-contiguous order: a reverse scan, so
// the last element's suffix is exactly 0 and nothing cancels (a `total - prefix` in fp32 leaves a 1e-7 residue that
// the reference's last interval of 1e10 would blow up in d sigma)
template <int C, typename T>
__device__ __forceinline__ void warp_suffix_scan(const float (&v)[C], float (&suf)[C], int lane) {
  T run = 0, part[C];
#pragma unroll
  for (int j = C - 1; j >= 0; --j) {
    part[j] = run;
    run += (T)v[j];
  }
  T incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += t;
  }
  T base = __shfl_down_sync(0xffffffffu, incl, 1);   // total of the lanes after this one
  if (lane == 31) base = 0;
#pragma unroll
  for (int j = 0; j < C; ++j) suf[j] = (float)(part[j] + base);
}

template <int C, bool FAST, int G>
__global__ void __launch_bounds__(kCompWarps * 32)
composite_fwd_kernel(const float* __restrict__ z, const float* __restrict__ sdf, const float* __restrict__ rgb,
                     const float* __restrict__ normals, const float* __restrict__ beta_param, float beta_min,
                     const float* __restrict__ depth_scale, const float* __restrict__ z_max, int64_t R, int S,
                     int flags, float* __restrict__ weights, float* __restrict__ rgb_values,
                     float* __restrict__ depth_values, float* __restrict__ normal_map,
                     float* __restrict__ bg_trans) {
  extern __shared__ __align__(16) uint8_t comp_smem_raw[];
  CompSmem<C, G>& sm = *reinterpret_cast<CompSmem<C, G>*>(comp_smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool abs_d = flags & SVS_COMP_ABS_DENSITY, rev = flags & SVS_COMP_REVERSED,
             tail = flags & SVS_COMP_ZMAX_TAIL;
  const float beta = abs_d ? 1.f : beta_of(beta_param, beta_min);
  const int64_t ray0 = blockIdx.x * (int64_t)kCompWarps + warp, stride = (int64_t)gridDim.x * kCompWarps;
  // rows shorter than the padded width read as zeros: the copies never touch the pad
  for (int k = 0; k < kCompStages; ++k) {
    CompRows<C, G>& st = sm.in[warp][k];
    for (int i = lane; i < 32 * C + 4; i += 32) st.z[i] = 0.f;
    for (int i = lane; i < 32 * C; i += 32) st.s[i] = 0.f;
  }
  __syncwarp();
  const float* nrm_rows = normal_map ? normals : nullptr;
  // one commit group per ray (empty past the end), so `wait_group kCompStages - 1` always means "the current ray landed"
  auto issue = [&](int64_t r, int stage, float* ds_out, float* zmax_out) {
    if (r < R) {
      issue_rows<C, G>(sm.in[warp][stage], z + r * S, sdf + r * S, rgb ? rgb + r * S * 3 : nullptr,
                       nrm_rows ? nrm_rows + r * S * 3 : nullptr, 3 * S, S, lane);
      *ds_out = depth_scale ? __ldg(depth_scale + r) : 1.f;
      *zmax_out = tail ? __ldg(z_max + r) : 0.f;
    } else {
      cp_async_commit();
    }
  };
  float ds_q[kCompStages - 1] = {1.f, 1.f}, zm_q[kCompStages - 1] = {0.f, 0.f};
  issue(ray0, 0, &ds_q[0], &zm_q[0]);
  issue(ray0 + stride, 1, &ds_q[1], &zm_q[1]);
  int k = 0;
  for (int64_t ray = ray0; ray < R; ray += stride, k = (k + 1 == kCompStages) ? 0 : k + 1) {
    const float ds = ds_q[0], zmax = zm_q[0];
    ds_q[0] = ds_q[1];
    zm_q[0] = zm_q[1];
    issue(ray + 2 * stride, (k + 2) % kCompStages, &ds_q[1], &zm_q[1]);
    cp_async_wait<kCompStages - 1>();
    __syncwarp();
    const CompRows<C, G>& in = sm.in[warp][k];
    using ScanT = typename std::conditional<FAST, float, double>::type;
    float E[C], zz[C], sv[C];
    ScanT excl[C];
    ld_chunk<C>(in.z, lane * C, zz);
    ld_chunk<C>(in.s, lane * C, sv);
    const float z_next_lane = __shfl_down_sync(0xffffffffu, zz[0], 1);   // z[C (lane + 1)]; unused for the last lane (i >= S - 1)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      int i = lane * C + j;
      float zi = zz[j];
      float em;
      float sigma = density_fwd<FAST>(sv[j], beta, abs_d, &em);
      float d;
      if (i < S - 1) {
        float zn = (j + 1 < C) ? zz[(j + 1) % C] : z_next_lane;
        d = rev ? (zi - zn) : (zn - zi);
      } else if (i == S - 1) {
        d = tail ? (zmax - zi) : 1e10f;
      } else {
        d = 0.f;
      }
      E[j] = (i < S) ? d * sigma : 0.f;
    }
    ScanT total = warp_excl_scan<C, ScanT>(E, excl, lane);
    float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_w = 0.f, acc_wz = 0.f;
    float cv[3 * C], wv[C];
    if (rgb) ld_chunk<3 * C>(in.c, lane * 3 * C, cv);
#pragma unroll
    for (int j = 0; j < C; ++j) {
      int i = lane * C + j;
      float T = exp_t<FAST>(-(float)excl[j]);
      float a = 1.0f - exp_t<FAST>(-E[j]);
      float w = (i < S) ? a * T : 0.f;
      wv[j] = w;
      if (i < S) {
        acc_w += w;
        acc_wz += w * zz[j];
        if (rgb) {
          acc_r += w * cv[3 * j];
          acc_g += w * cv[3 * j + 1];
          acc_b += w * cv[3 * j + 2];
        }
      }
    }
    st_chunk<C>(sm.w[warp], lane * C, wv);
    __syncwarp();
    for (int i = lane; i < S; i += 32) weights[ray * S + i] = sm.w[warp][i];   // coalesced (the lane-contiguous layout is not)
    acc_w = warp_sum(acc_w);
    acc_wz = warp_sum(acc_wz);
    if (rgb) {
      acc_r = warp_sum(acc_r);
      acc_g = warp_sum(acc_g);
      acc_b = warp_sum(acc_b);
    }
    float nr = 0.f, ng = 0.f, nb = 0.f;
    if (normal_map) {  // eval: sum w * g/|g|  (network.py:270-274; no eps, as the reference)
      for (int i = lane; i < S; i += 32) {
        float gx = in.g[3 * i], gy = in.g[3 * i + 1], gz = in.g[3 * i + 2];
        float n = sqrtf(gx * gx + gy * gy + gz * gz);
        float w = sm.w[warp][i];
        nr += w * (gx / n);
        ng += w * (gy / n);
        nb += w * (gz / n);
      }
      nr = warp_sum(nr);
      ng = warp_sum(ng);
      nb = warp_sum(nb);
    }
    if (lane == 0) {
      if (rgb_values) {
        rgb_values[ray * 3] = acc_r;
        rgb_values[ray * 3 + 1] = acc_g;
        rgb_values[ray * 3 + 2] = acc_b;
      }
      if (depth_values) {
        depth_values[ray] = ds * (acc_wz / (acc_w + 1e-8f));
      }
      if (normal_map) {
        normal_map[ray * 3] = nr;
        normal_map[ray * 3 + 1] = ng;
        normal_map[ray * 3 + 2] = nb;
      }
      if (bg_trans) bg_trans[ray] = exp_t<FAST>(-(float)total);
    }
    __syncwarp();
  }
}

template <int C, bool FAST, int G>
__global__ void __launch_bounds__(kCompWarps * 32)
composite_bwd_kernel(const float* __restrict__ z, const float* __restrict__ sdf, const float* __restrict__ rgb,
                     const float* __restrict__ beta_param, float beta_min,
                     const float* __restrict__ depth_scale, const float* __restrict__ z_max, int64_t R, int S,
                     int flags, const float* __restrict__ d_rgb_values, const float* __restrict__ d_depth_values,
                     const float* __restrict__ d_weights, const float* __restrict__ d_bg_trans,
                     float* __restrict__ d_sdf, float* __restrict__ d_rgb, float* __restrict__ d_beta_param) {
  extern __shared__ __align__(16) uint8_t comp_smem_raw[];
  CompSmem<C, G>& sm = *reinterpret_cast<CompSmem<C, G>*>(comp_smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool abs_d = flags & SVS_COMP_ABS_DENSITY, rev = flags & SVS_COMP_REVERSED,
             tail = flags & SVS_COMP_ZMAX_TAIL;
  const float beta = abs_d ? 1.f : beta_of(beta_param, beta_min);
  float dbeta_acc = 0.f;
  const int64_t ray0 = blockIdx.x * (int64_t)kCompWarps + warp, stride = (int64_t)gridDim.x * kCompWarps;
  for (int k = 0; k < kCompStages; ++k) {
    CompRows<C, G>& st = sm.in[warp][k];
    for (int i = lane; i < 32 * C + 4; i += 32) st.z[i] = 0.f;
    for (int i = lane; i < 32 * C; i += 32) st.s[i] = 0.f;
  }
  __syncwarp();
  // per-ray scalars of the next rays are fetched together with their rows
  struct RayScal { float gr, gg, gb, gdep, gbt, ds, zmax; };
  auto load_scal = [&](int64_t ray) {
    RayScal q;
    q.gr = d_rgb_values ? __ldg(d_rgb_values + ray * 3) : 0.f;
    q.gg = d_rgb_values ? __ldg(d_rgb_values + ray * 3 + 1) : 0.f;
    q.gb = d_rgb_values ? __ldg(d_rgb_values + ray * 3 + 2) : 0.f;
    q.gdep = d_depth_values ? __ldg(d_depth_values + ray) : 0.f;
    q.gbt = (tail && d_bg_trans) ? __ldg(d_bg_trans + ray) : 0.f;
    q.ds = depth_scale ? __ldg(depth_scale + ray) : 1.f;
    q.zmax = tail ? __ldg(z_max + ray) : 0.f;
    return q;
  };
  auto issue = [&](int64_t r, int stage, RayScal* q) {
    if (r < R) {
      issue_rows<C, G>(sm.in[warp][stage], z + r * S, sdf + r * S, rgb ? rgb + r * S * 3 : nullptr,
                       d_weights ? d_weights + r * S : nullptr, S, S, lane);
      *q = load_scal(r);
    } else {
      cp_async_commit();
    }
  };
  RayScal q0 = {0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f}, q1 = q0;
  issue(ray0, 0, &q0);
  issue(ray0 + stride, 1, &q1);
  int k = 0;
  for (int64_t ray = ray0; ray < R; ray += stride, k = (k + 1 == kCompStages) ? 0 : k + 1) {
    const RayScal cq = q0;
    q0 = q1;
    issue(ray + 2 * stride, (k + 2) % kCompStages, &q1);
    cp_async_wait<kCompStages - 1>();
    __syncwarp();
    const CompRows<C, G>& in = sm.in[warp][k];
    const float gr = cq.gr, gg = cq.gg, gb = cq.gb, gdep = cq.gdep, gbt = cq.gbt, ds = cq.ds;
    const float ib = __frcp_rn(beta), i2b2 = __frcp_rn(2.0f * beta * beta);
    using ScanT = typename std::conditional<FAST, float, double>::type;
    float E[C], zz[C], dl[C], sig[C], em[C], ss[C];
    ScanT excl[C];
    ld_chunk<C>(in.z, lane * C, zz);
    ld_chunk<C>(in.s, lane * C, ss);
    const float z_next_lane = __shfl_down_sync(0xffffffffu, zz[0], 1);
#pragma unroll
    for (int j = 0; j < C; ++j) {
      int i = lane * C + j;
      float zi = zz[j];
      sig[j] = density_fwd<FAST>(ss[j], beta, abs_d, &em[j]);
      float d;
      if (i < S - 1) {
        float zn = (j + 1 < C) ? zz[(j + 1) % C] : z_next_lane;
        d = rev ? (zi - zn) : (zn - zi);
      } else if (i == S - 1) {
        d = tail ? (cq.zmax - zi) : 1e10f;
      } else {
        d = 0.f;
      }
      dl[j] = d;
      E[j] = (i < S) ? d * sig[j] : 0.f;
    }
    ScanT total = warp_excl_scan<C, ScanT>(E, excl, lane);
    float w[C], Te[C];
    float acc_w = 0.f, acc_wz = 0.f;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      int i = lane * C + j;
      float T = exp_t<FAST>(-(float)excl[j]);
      float ee = exp_t<FAST>(-E[j]);
      w[j] = (i < S) ? (1.0f - ee) * T : 0.f;
      Te[j] = T * ee;
      acc_w += w[j];
      acc_wz += w[j] * zz[j];
    }
    acc_w = warp_sum(acc_w);
    acc_wz = warp_sum(acc_wz);
    const float Wt = acc_w + 1e-8f;
    const float iWt2 = __frcp_rn(Wt * Wt);
    // w_hat_i = c_i . dL/drgb + dL/dw_i + dL/ddepth * ds * (z_i*Wt - sum(wz)) / Wt^2
    float what[C], ww[C];
    float cv[3 * C], gv[C];
    if (rgb) ld_chunk<3 * C>(in.c, lane * 3 * C, cv);
    if (d_weights) ld_chunk<C>(in.g, lane * C, gv);
#pragma unroll
    for (int j = 0; j < C; ++j) {
      int i = lane * C + j;
      float v = 0.f;
      if (i < S) {
        if (rgb) v = cv[3 * j] * gr + cv[3 * j + 1] * gg + cv[3 * j + 2] * gb;
        if (d_weights) v += gv[j];
        v += FAST ? gdep * ds * (zz[j] * Wt - acc_wz) * iWt2 : gdep * ds * (zz[j] * Wt - acc_wz) / (Wt * Wt);
      }
      what[j] = v;
      ww[j] = v * w[j];
    }
    // suffix sums sum_{k>i} what_k w_k as a reverse scan (fp64 in the canonical mode).  NOT total - inclusive prefix: the
    // residue of that cancellation, times the reference's last interval of 1e10, put 8 % of error into dL/dbeta of a
    // 1024-ray batch (tools/dbeta_debug.py)
    float suf[C];
    warp_suffix_scan<C, ScanT>(ww, suf, lane);
    const float bgt = tail ? exp_t<FAST>(-(float)total) : 0.f;
    float dbeta = 0.f;
    float ov[C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
      int i = lane * C + j;
      ov[j] = 0.f;
      if (i < S) {
        float suffix = suf[j];
        float dE = what[j] * Te[j] - suffix - gbt * bgt;
        float dsig = dl[j] * dE;
        float dsdf;
        if (abs_d) {
          float sg = (ss[j] > 0.f) ? 1.f : ((ss[j] < 0.f) ? -1.f : 0.f);
          dsdf = dsig * sg;
        } else {
          float e = em[j] + 1.0f;  // autograd's expm1 backward uses result + 1
          float nz = (ss[j] != 0.f) ? 1.f : 0.f;
          if (FAST) {
            dsdf = -dsig * nz * e * i2b2;
            dbeta += dsig * (-sig[j] * ib + ss[j] * e * (i2b2 * ib));
          } else {
            dsdf = -dsig * nz * e / (2.0f * beta * beta);
            dbeta += dsig * (-sig[j] / beta + ss[j] * e / (2.0f * beta * beta * beta));
          }
        }
        ov[j] = dsdf;
      }
    }
    st_chunk<C>(sm.o[warp], lane * C, ov);   // rows leave through shared memory: coalesced stores
    st_chunk<C>(sm.w[warp], lane * C, w);
    dbeta_acc += dbeta;
    __syncwarp();
    for (int i = lane; i < S; i += 32) d_sdf[ray * S + i] = sm.o[warp][i];
    if (d_rgb) {
      float* o = d_rgb + ray * S * 3;
      for (int i = lane; i < S * 3; i += 32) {
        int e = i / 3, ch = i - 3 * e;
        float g = (ch == 0) ? gr : ((ch == 1) ? gg : gb);
        o[i] = sm.w[warp][e] * g;
      }
    }
    __syncwarp();
  }
  if (!abs_d && d_beta_param) {
    dbeta_acc = warp_sum(dbeta_acc);
    if (lane == 0 && dbeta_acc != 0.f) {
      float sg = (__ldg(beta_param) > 0.f) ? 1.f : ((__ldg(beta_param) < 0.f) ? -1.f : 0.f);
      atomicAdd(d_beta_param, sg * dbeta_acc);
    }
  }
}

// stand-alone density modules (density.py:16-35)
__global__ void density_fwd_kernel(const float* __restrict__ sdf, int64_t n, int S, const float* __restrict__ beta_param,
                                   float beta_min, const float* __restrict__ beta_rows, int abs_d,
                                   float* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float beta = abs_d ? 1.f : (beta_rows ? beta_rows[i / S] : beta_of(beta_param, beta_min));
  float em;
  out[i] = density_fwd(sdf[i], beta, abs_d, &em);
}

__global__ void density_bwd_kernel(const float* __restrict__ sdf, int64_t n, int S, const float* __restrict__ beta_param,
                                   float beta_min, const float* __restrict__ beta_rows, int abs_d,
                                   const float* __restrict__ d_out, float* __restrict__ d_sdf,
                                   float* __restrict__ d_beta_param) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float dbeta = 0.f;
  if (i < n) {
    float s = sdf[i], g = d_out[i];
    float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
    if (abs_d) {
      d_sdf[i] = g * sg;
    } else {
      float beta = beta_rows ? beta_rows[i / S] : beta_of(beta_param, beta_min);
      float em;
      float sigma = density_fwd(s, beta, false, &em);
      float e = em + 1.0f;
      d_sdf[i] = -g * (sg * sg) * e / (2.0f * beta * beta);
      dbeta = g * (-sigma / beta + s * e / (2.0f * beta * beta * beta));
    }
  }
  if (!abs_d && !beta_rows && d_beta_param) {
    dbeta = warp_sum(dbeta);
    if ((threadIdx.x & 31) == 0 && dbeta != 0.f) {
      float bp = __ldg(beta_param);
      atomicAdd(d_beta_param, ((bp > 0.f) ? 1.f : ((bp < 0.f) ? -1.f : 0.f)) * dbeta);
    }
  }
}


// ----------------------------------------------------------------------------------------------------------------
// Lean FAST forward (the common case: Laplace density, forward sample order, no z_max tail, rgb present, S even).
// ncu on the general FAST kernel (profiles/r2_hbm_kernels_full.txt): 645 warp instructions per ray, issue slots ~70 % busy
// at 47 % of the HBM roofline — instruction-bound, a third of it predicates / branches for the variants and 4-byte copies.
// Here: 8-byte async copies (rows of an even S are 8-byte aligned), no per-sample branches (padding samples carry
// dist = 0, hence E = 0, alpha = 0, w = 0), explicit FMAs (this file is compiled with -fmad=false for the canonical mode),
// one 8-value transposing warp reduction (12 shuffles) instead of five to eight butterflies (40+), 64-bit weight stores.
// ----------------------------------------------------------------------------------------------------------------
constexpr int kLeanWarps = 8;
constexpr int kLeanStages = 3;
constexpr float kL2E = 1.4426950408889634f;

template <int C, bool NORMALS>
struct LeanRows {
  float z[32 * C + 4];
  float s[32 * C];
  float c[96 * C];
  float g[NORMALS ? 96 * C : 4];
};
template <int C, bool NORMALS>
struct LeanSmem {
  LeanRows<C, NORMALS> in[kLeanWarps][kLeanStages];
  float w[kLeanWarps][32 * C];
};

__device__ __forceinline__ void cp_async8(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ float ex2_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// sums 8 values over the warp with 3 halving exchanges + 2 butterflies; lane l ends with the total of value
// id(l) = 4 * bit4(l) + 2 * bit3(l) + bit2(l)
__device__ __forceinline__ float warp_reduce8(float (&v)[8], int lane) {
#pragma unroll
  for (int w = 4, m = 16; w >= 1; w >>= 1, m >>= 1) {
    const bool upper = (lane & m) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = upper ? v[i] : v[i + w];
      const float keep = upper ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
  }
  float r = v[0];
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

template <int C, bool NORMALS>
__global__ void __launch_bounds__(kLeanWarps * 32)
composite_fwd_lean_kernel(const float* __restrict__ z, const float* __restrict__ sdf, const float* __restrict__ rgb,
                          const float* __restrict__ normals, const float* __restrict__ beta_param, float beta_min,
                          const float* __restrict__ depth_scale, int64_t R, int S, float* __restrict__ weights,
                          float* __restrict__ rgb_values, float* __restrict__ depth_values, float* __restrict__ normal_map) {
  extern __shared__ __align__(16) uint8_t comp_smem_raw[];
  LeanSmem<C, NORMALS>& sm = *reinterpret_cast<LeanSmem<C, NORMALS>*>(comp_smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float beta = beta_of(beta_param, beta_min);
  const float ib = __frcp_rn(beta), kib = -kL2E * ib;
  const int64_t ray0 = blockIdx.x * (int64_t)kLeanWarps + warp, stride = (int64_t)gridDim.x * kLeanWarps;
  for (int k = 0; k < kLeanStages; ++k) {   // the pad behind S reads as zeros: the copies never touch it
    LeanRows<C, NORMALS>& st = sm.in[warp][k];
    for (int i = lane; i < 32 * C + 4; i += 32) st.z[i] = 0.f;
    for (int i = lane; i < 32 * C; i += 32) st.s[i] = 0.f;
    for (int i = lane; i < 96 * C; i += 32) st.c[i] = 0.f;
    if (NORMALS)
      for (int i = lane; i < 96 * C; i += 32) st.g[i] = 1.f;
  }
  __syncwarp();
  const int h1 = S >> 1, h3 = (3 * S) >> 1;   // 8-byte elements per row
  auto issue = [&](int64_t r, int stage, float* ds_out) {
    if (r < R) {
      LeanRows<C, NORMALS>& st = sm.in[warp][stage];
      const float* zr = z + r * S;
      const float* sr = sdf + r * S;
      const float* cr = rgb + r * S * 3;
#pragma unroll
      for (int j = 0; j < (C + 1) / 2; ++j) {
        const int i = lane + 32 * j;
        if (i < h1) {
          cp_async8(st.z + 2 * i, zr + 2 * i);
          cp_async8(st.s + 2 * i, sr + 2 * i);
        }
      }
#pragma unroll
      for (int j = 0; j < (3 * C + 1) / 2; ++j) {
        const int i = lane + 32 * j;
        if (i < h3) cp_async8(st.c + 2 * i, cr + 2 * i);
      }
      if (NORMALS) {
        const float* gr = normals + r * S * 3;
#pragma unroll
        for (int j = 0; j < (3 * C + 1) / 2; ++j) {
          const int i = lane + 32 * j;
          if (i < h3) cp_async8(st.g + 2 * i, gr + 2 * i);
        }
      }
      *ds_out = __ldg(depth_scale + r);
    }
    cp_async_commit();
  };
  float ds_q[2] = {1.f, 1.f};
  issue(ray0, 0, &ds_q[0]);
  issue(ray0 + stride, 1, &ds_q[1]);
  int k = 0;
  for (int64_t ray = ray0; ray < R; ray += stride, k = (k + 1 == kLeanStages) ? 0 : k + 1) {
    const float ds = ds_q[0];
    ds_q[0] = ds_q[1];
    issue(ray + 2 * stride, (k + 2) % kLeanStages, &ds_q[1]);
    cp_async_wait<kLeanStages - 1>();
    __syncwarp();
    const LeanRows<C, NORMALS>& in = sm.in[warp][k];
    float zz[C], sv[C], cv[3 * C], E[C], ex[C];
    ld_chunk<C>(in.z, lane * C, zz);
    ld_chunk<C>(in.s, lane * C, sv);
    ld_chunk<3 * C>(in.c, lane * 3 * C, cv);
    const float z_next_lane = __shfl_down_sync(0xffffffffu, zz[0], 1);
    float run = 0.f;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const int i = lane * C + j;
      const float zn = (j + 1 < C) ? zz[(j + 1) % C] : z_next_lane;
      float d = zn - zz[j];
      d = (i < S - 1) ? d : ((i == S - 1) ? 1e10f : 0.f);
      const float e = ex2_fast(fabsf(sv[j]) * kib);                         // exp(-|s| / beta)
      const float sigma = ib * fmaf(copysignf(0.5f, sv[j]), e - 1.0f, 0.5f);   // density.py:21-26
      E[j] = d * sigma;
      ex[j] = run;
      run += E[j];
    }
    float incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    float base = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) base = 0.f;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // w, w z, r, g, b, nx, ny, nz
    float wv[C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const float T = ex2_fast(-kL2E * (base + ex[j]));
      const float w = (1.0f - ex2_fast(-kL2E * E[j])) * T;    // padding: E = 0 -> w = 0
      wv[j] = w;
      acc[0] += w;
      acc[1] = fmaf(w, zz[j], acc[1]);
      acc[2] = fmaf(w, cv[3 * j], acc[2]);
      acc[3] = fmaf(w, cv[3 * j + 1], acc[3]);
      acc[4] = fmaf(w, cv[3 * j + 2], acc[4]);
    }
    if (NORMALS) {   // eval: sum w g / |g| (network.py:270-274)
      float gv[3 * C];
      ld_chunk<3 * C>(in.g, lane * 3 * C, gv);
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const float gx = gv[3 * j], gy = gv[3 * j + 1], gz = gv[3 * j + 2];
        const float wn = wv[j] * rsqrtf(fmaf(gx, gx, fmaf(gy, gy, gz * gz)));
        acc[5] = fmaf(wn, gx, acc[5]);
        acc[6] = fmaf(wn, gy, acc[6]);
        acc[7] = fmaf(wn, gz, acc[7]);
      }
    }
    st_chunk<C>(sm.w[warp], lane * C, wv);
    const float tot = warp_reduce8(acc, lane);      // lane l: total of value 4 bit4 + 2 bit3 + bit2
    __syncwarp();
    {
      float2* wout = reinterpret_cast<float2*>(weights + ray * S);
      const float2* wsm = reinterpret_cast<const float2*>(sm.w[warp]);
#pragma unroll
      for (int j = 0; j < (C + 1) / 2; ++j) {
        const int i = lane + 32 * j;
        if (i < h1) wout[i] = wsm[i];
      }
    }
    const float w_tot = __shfl_sync(0xffffffffu, tot, 0), wz_tot = __shfl_sync(0xffffffffu, tot, 4);
    {
      // lanes 0, 8, 12, 16 (, 20, 24, 28) hold depth / r / g / b (/ normal): ONE predicated store, no divergent branches
      // (per-lane `if (lane == k)` stores cost 22 % of the kernel's stall samples in branch resolution)
      const int id = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
      const float outv = (id == 0) ? ds * (wz_tot / (w_tot + 1e-8f)) : tot;
      float* dst = depth_values + ray;
      if (id >= 2 && id <= 4) dst = rgb_values + ray * 3 + (id - 2);
      if (NORMALS && id >= 5) dst = normal_map + ray * 3 + (id - 5);
      const bool on = (lane & 3) == 0 && id != 1 && (NORMALS || id <= 4);
      if (on) *dst = outv;
    }
    __syncwarp();
  }
}

template <int C, bool NORMALS>
static int launch_fwd_lean(const float* z, const float* sdf, const float* rgb, const float* normals, const float* beta_param,
                           float beta_min, const float* depth_scale, int64_t R, int S, float* weights, float* rgb_values,
                           float* depth_values, float* normal_map, cudaStream_t st) {
  static int cap_ctas = 0;
  const size_t smem = sizeof(LeanSmem<C, NORMALS>);
  if (cap_ctas == 0) {
    int per_sm = 0, sms = 0, dev = 0;
    SVS_CUDA_OK(cudaFuncSetAttribute(composite_fwd_lean_kernel<C, NORMALS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, composite_fwd_lean_kernel<C, NORMALS>, kLeanWarps * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = kNumSMs;
    cap_ctas = sms * per_sm;
  }
  const int64_t blocks = cdiv(R, kLeanWarps);
  composite_fwd_lean_kernel<C, NORMALS><<<(int)(blocks < cap_ctas ? blocks : cap_ctas), kLeanWarps * 32, smem, st>>>(
      z, sdf, rgb, normals, beta_param, beta_min, depth_scale, R, S, weights, rgb_values, depth_values, normal_map);
  return SVS_OK;
}


// Lean FAST backward (same restrictions as the lean forward): closed form of SURVEY.md App. G with fp32 scans, the
// cancellation-free suffix scan, 8-byte async copies in, 64-bit coalesced stores out (d_rgb is assembled per lane as 3 C
// consecutive floats in shared memory, conflict-free 128-bit stores).
template <int C, bool HAS_DW>
struct LeanBwdRows {
  float z[32 * C + 4];
  float s[32 * C];
  float c[96 * C];
  float g[HAS_DW ? 32 * C : 4];
};
template <int C, bool HAS_DW>
struct LeanBwdSmem {
  LeanBwdRows<C, HAS_DW> in[kLeanWarps][kLeanStages];
  float o[kLeanWarps][32 * C];
  float o3[kLeanWarps][96 * C];
};

template <int C, bool HAS_DW>
__global__ void __launch_bounds__(kLeanWarps * 32)
composite_bwd_lean_kernel(const float* __restrict__ z, const float* __restrict__ sdf, const float* __restrict__ rgb,
                          const float* __restrict__ beta_param, float beta_min, const float* __restrict__ depth_scale,
                          int64_t R, int S, const float* __restrict__ d_rgb_values, const float* __restrict__ d_depth_values,
                          const float* __restrict__ d_weights, float* __restrict__ d_sdf, float* __restrict__ d_rgb,
                          float* __restrict__ d_beta_param) {
  extern __shared__ __align__(16) uint8_t comp_smem_raw[];
  LeanBwdSmem<C, HAS_DW>& sm = *reinterpret_cast<LeanBwdSmem<C, HAS_DW>*>(comp_smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float beta = beta_of(beta_param, beta_min);
  const float ib = __frcp_rn(beta), kib = -kL2E * ib, i2b2 = __frcp_rn(2.0f * beta * beta), i2b3 = i2b2 * ib;
  const int64_t ray0 = blockIdx.x * (int64_t)kLeanWarps + warp, stride = (int64_t)gridDim.x * kLeanWarps;
  for (int k = 0; k < kLeanStages; ++k) {
    LeanBwdRows<C, HAS_DW>& st = sm.in[warp][k];
    for (int i = lane; i < 32 * C + 4; i += 32) st.z[i] = 0.f;
    for (int i = lane; i < 32 * C; i += 32) st.s[i] = 0.f;
    for (int i = lane; i < 96 * C; i += 32) st.c[i] = 0.f;
    if (HAS_DW)
      for (int i = lane; i < 32 * C; i += 32) st.g[i] = 0.f;
  }
  __syncwarp();
  const int h1 = S >> 1, h3 = (3 * S) >> 1;
  struct Scal { float gr, gg, gb, gdep, ds; };
  auto issue = [&](int64_t r, int stage, Scal* q) {
    if (r < R) {
      LeanBwdRows<C, HAS_DW>& st = sm.in[warp][stage];
      const float* zr = z + r * S;
      const float* sr = sdf + r * S;
      const float* cr = rgb + r * S * 3;
#pragma unroll
      for (int j = 0; j < (C + 1) / 2; ++j) {
        const int i = lane + 32 * j;
        if (i < h1) {
          cp_async8(st.z + 2 * i, zr + 2 * i);
          cp_async8(st.s + 2 * i, sr + 2 * i);
          if (HAS_DW) cp_async8(st.g + 2 * i, d_weights + r * S + 2 * i);
        }
      }
#pragma unroll
      for (int j = 0; j < (3 * C + 1) / 2; ++j) {
        const int i = lane + 32 * j;
        if (i < h3) cp_async8(st.c + 2 * i, cr + 2 * i);
      }
      q->gr = __ldg(d_rgb_values + r * 3);
      q->gg = __ldg(d_rgb_values + r * 3 + 1);
      q->gb = __ldg(d_rgb_values + r * 3 + 2);
      q->gdep = __ldg(d_depth_values + r);
      q->ds = __ldg(depth_scale + r);
    }
    cp_async_commit();
  };
  Scal q0 = {0.f, 0.f, 0.f, 0.f, 1.f}, q1 = q0;
  issue(ray0, 0, &q0);
  issue(ray0 + stride, 1, &q1);
  float dbeta_acc = 0.f;
  int k = 0;
  for (int64_t ray = ray0; ray < R; ray += stride, k = (k + 1 == kLeanStages) ? 0 : k + 1) {
    const Scal cq = q0;
    q0 = q1;
    issue(ray + 2 * stride, (k + 2) % kLeanStages, &q1);
    cp_async_wait<kLeanStages - 1>();
    __syncwarp();
    const LeanBwdRows<C, HAS_DW>& in = sm.in[warp][k];
    float zz[C], ss[C], cv[3 * C], E[C], ex[C], dl[C], sig[C], ee0[C];
    ld_chunk<C>(in.z, lane * C, zz);
    ld_chunk<C>(in.s, lane * C, ss);
    ld_chunk<3 * C>(in.c, lane * 3 * C, cv);
    const float z_next_lane = __shfl_down_sync(0xffffffffu, zz[0], 1);
    float run = 0.f;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const int i = lane * C + j;
      const float zn = (j + 1 < C) ? zz[(j + 1) % C] : z_next_lane;
      float d = zn - zz[j];
      d = (i < S - 1) ? d : ((i == S - 1) ? 1e10f : 0.f);
      dl[j] = d;
      ee0[j] = ex2_fast(fabsf(ss[j]) * kib);                                    // e = exp(-|s| / beta) = expm1 + 1
      sig[j] = ib * fmaf(copysignf(0.5f, ss[j]), ee0[j] - 1.0f, 0.5f);
      E[j] = d * sig[j];
      ex[j] = run;
      run += E[j];
    }
    float incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    float base = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) base = 0.f;
    float w[C], Te[C];
    float acc_w = 0.f, acc_wz = 0.f;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const float T = ex2_fast(-kL2E * (base + ex[j]));
      const float ee = ex2_fast(-kL2E * E[j]);
      w[j] = (1.0f - ee) * T;
      Te[j] = T * ee;
      acc_w += w[j];
      acc_wz = fmaf(w[j], zz[j], acc_wz);
    }
    acc_w = warp_sum(acc_w);
    acc_wz = warp_sum(acc_wz);
    const float Wt = acc_w + 1e-8f;
    const float kd = cq.gdep * cq.ds * __frcp_rn(Wt * Wt);
    // w_hat_i = c_i . dL/drgb + dL/dw_i + dL/ddepth * ds * (z_i Wt - sum(w z)) / Wt^2 ; suffix sums of w_hat_k w_k
    float what[C], ww[C], gv[C];
    if constexpr (HAS_DW) ld_chunk<C>(in.g, lane * C, gv);
    float srun = 0.f, sx[C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
      float v = fmaf(cv[3 * j], cq.gr, fmaf(cv[3 * j + 1], cq.gg, cv[3 * j + 2] * cq.gb));
      if constexpr (HAS_DW) v += gv[j];
      v = fmaf(kd, fmaf(zz[j], Wt, -acc_wz), v);
      what[j] = v;
      ww[j] = v * w[j];        // padding: w = 0
    }
#pragma unroll
    for (int j = C - 1; j >= 0; --j) {
      sx[j] = srun;
      srun += ww[j];
    }
    float sincl = srun;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_down_sync(0xffffffffu, sincl, o);
      if (lane + o < 32) sincl += t;
    }
    float sbase = __shfl_down_sync(0xffffffffu, sincl, 1);
    if (lane == 31) sbase = 0.f;
    float ov[C], o3v[3 * C];
    float dbeta = 0.f;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const float dE = fmaf(what[j], Te[j], -(sbase + sx[j]));
      const float dsig = dl[j] * dE;                          // padding: dist = 0
      const float nz = (ss[j] != 0.f) ? 1.f : 0.f;
      ov[j] = -dsig * nz * ee0[j] * i2b2;
      dbeta = fmaf(dsig, fmaf(ss[j] * ee0[j], i2b3, -sig[j] * ib), dbeta);
      o3v[3 * j] = w[j] * cq.gr;
      o3v[3 * j + 1] = w[j] * cq.gg;
      o3v[3 * j + 2] = w[j] * cq.gb;
    }
    dbeta_acc += dbeta;
    st_chunk<C>(sm.o[warp], lane * C, ov);
    st_chunk<3 * C>(sm.o3[warp], lane * 3 * C, o3v);
    __syncwarp();
    {
      float2* o1 = reinterpret_cast<float2*>(d_sdf + ray * S);
      const float2* s1 = reinterpret_cast<const float2*>(sm.o[warp]);
#pragma unroll
      for (int j = 0; j < (C + 1) / 2; ++j) {
        const int i = lane + 32 * j;
        if (i < h1) o1[i] = s1[i];
      }
      float2* o3 = reinterpret_cast<float2*>(d_rgb + ray * S * 3);
      const float2* s3 = reinterpret_cast<const float2*>(sm.o3[warp]);
#pragma unroll
      for (int j = 0; j < (3 * C + 1) / 2; ++j) {
        const int i = lane + 32 * j;
        if (i < h3) o3[i] = s3[i];
      }
    }
    __syncwarp();
  }
  if (d_beta_param) {
    dbeta_acc = warp_sum(dbeta_acc);
    if (lane == 0 && dbeta_acc != 0.f) {
      const float bp = __ldg(beta_param);
      atomicAdd(d_beta_param, ((bp > 0.f) ? 1.f : ((bp < 0.f) ? -1.f : 0.f)) * dbeta_acc);
    }
  }
}

template <int C, bool HAS_DW>
static int launch_bwd_lean(const float* z, const float* sdf, const float* rgb, const float* beta_param, float beta_min,
                           const float* depth_scale, int64_t R, int S, const float* d_rgb_values, const float* d_depth_values,
                           const float* d_weights, float* d_sdf, float* d_rgb, float* d_beta_param, cudaStream_t st) {
  static int cap_ctas = 0;
  const size_t smem = sizeof(LeanBwdSmem<C, HAS_DW>);
  if (cap_ctas == 0) {
    int per_sm = 0, sms = 0, dev = 0;
    SVS_CUDA_OK(cudaFuncSetAttribute(composite_bwd_lean_kernel<C, HAS_DW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, composite_bwd_lean_kernel<C, HAS_DW>, kLeanWarps * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = kNumSMs;
    cap_ctas = sms * per_sm;
  }
  const int64_t blocks = cdiv(R, kLeanWarps);
  composite_bwd_lean_kernel<C, HAS_DW><<<(int)(blocks < cap_ctas ? blocks : cap_ctas), kLeanWarps * 32, smem, st>>>(
      z, sdf, rgb, beta_param, beta_min, depth_scale, R, S, d_rgb_values, d_depth_values, d_weights, d_sdf, d_rgb, d_beta_param);
  return SVS_OK;
}

// one wave of resident CTAs (grid-stride over rays): a warp's next-ray prefetch then always has a successor
template <typename K>
static int comp_grid(K kernel, size_t smem, int64_t R) {
  static std::map<const void*, int> resident;   // kernel -> CTAs of one wave (attribute + occupancy queried once)
  int& cap_ctas = resident[(const void*)kernel];
  if (cap_ctas == 0) {
    int per_sm = 0, sms = 0, dev = 0;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kCompWarps * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = kNumSMs;
    cap_ctas = sms * per_sm;
  }
  const int64_t blocks = cdiv(R, kCompWarps), cap = cap_ctas;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace svs

using namespace svs;

#define DISPATCH_C(S, CALL)                       \
  if ((S) <= 32) { constexpr int C = 1; CALL; }   \
  else if ((S) <= 64) { constexpr int C = 2; CALL; } \
  else if ((S) <= 128) { constexpr int C = 4; CALL; } \
  else { constexpr int C = 8; CALL; }

extern "C" int svs_composite_forward(const float* z, const float* sdf, const float* rgb, const float* normals,
                                     const float* beta_param, float beta_min, const float* depth_scale,
                                     const float* z_max, int64_t R, int32_t S, int32_t flags, float* weights,
                                     float* rgb_values, float* depth_values, float* normal_map, float* bg_trans,
                                     void* stream) {
  SVS_CHECK_ARG(R >= 0 && S >= 1 && S <= 256, "svs_composite_forward: need 1 <= S <= 256 (got %d)", S);
  SVS_CHECK_ARG(z && sdf && weights, "svs_composite_forward: z/sdf/weights required");
  SVS_CHECK_ARG((flags & SVS_COMP_ABS_DENSITY) || beta_param, "svs_composite_forward: beta_param required");
  SVS_CHECK_ARG(!(flags & SVS_COMP_ZMAX_TAIL) || z_max, "svs_composite_forward: z_max required");
  SVS_CHECK_ARG(!normal_map || normals, "svs_composite_forward: normals required for normal_map");
  SVS_CHECK_ARG(!rgb_values || rgb, "svs_composite_forward: rgb required for rgb_values");
  if (R == 0) return SVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps("composite_fwd", 0.0, (double)R * (4.0 * S * (3 + (rgb ? 3 : 0) + (normal_map ? 3 : 0)) + 32), st);
  const bool lean_ok = flags == SVS_COMP_FAST && rgb && rgb_values && depth_values && depth_scale && !bg_trans && (S & 1) == 0 &&
                       S > 32 && S <= 256 &&
                       ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(sdf) | reinterpret_cast<uintptr_t>(rgb) |
                         reinterpret_cast<uintptr_t>(weights) | reinterpret_cast<uintptr_t>(normal_map ? normals : z)) & 7) == 0;
  if (lean_ok) {
    if (normal_map) {
      DISPATCH_C(S, (launch_fwd_lean<C, true>(z, sdf, rgb, normals, beta_param, beta_min, depth_scale, R, S, weights, rgb_values, depth_values, normal_map, st)));
    } else {
      DISPATCH_C(S, (launch_fwd_lean<C, false>(z, sdf, rgb, normals, beta_param, beta_min, depth_scale, R, S, weights, rgb_values, depth_values, normal_map, st)));
    }
  } else if (flags & SVS_COMP_FAST) {
    if (normal_map) {
      DISPATCH_C(S, (composite_fwd_kernel<C, true, 3><<<comp_grid(composite_fwd_kernel<C, true, 3>, sizeof(CompSmem<C, 3>), R), kCompWarps * 32, sizeof(CompSmem<C, 3>), st>>>(
                        z, sdf, rgb, normals, beta_param, beta_min, depth_scale, z_max, R, S, flags, weights,
                        rgb_values, depth_values, normal_map, bg_trans)));
    } else {   // training: no normals row -> 1/3 less shared memory, more resident warps
      DISPATCH_C(S, (composite_fwd_kernel<C, true, 0><<<comp_grid(composite_fwd_kernel<C, true, 0>, sizeof(CompSmem<C, 0>), R), kCompWarps * 32, sizeof(CompSmem<C, 0>), st>>>(
                        z, sdf, rgb, normals, beta_param, beta_min, depth_scale, z_max, R, S, flags, weights,
                        rgb_values, depth_values, normal_map, bg_trans)));
    }
  } else {
    if (normal_map) {
      DISPATCH_C(S, (composite_fwd_kernel<C, false, 3><<<comp_grid(composite_fwd_kernel<C, false, 3>, sizeof(CompSmem<C, 3>), R), kCompWarps * 32, sizeof(CompSmem<C, 3>), st>>>(
                        z, sdf, rgb, normals, beta_param, beta_min, depth_scale, z_max, R, S, flags, weights,
                        rgb_values, depth_values, normal_map, bg_trans)));
    } else {   // training: no normals row -> 1/3 less shared memory, more resident warps
      DISPATCH_C(S, (composite_fwd_kernel<C, false, 0><<<comp_grid(composite_fwd_kernel<C, false, 0>, sizeof(CompSmem<C, 0>), R), kCompWarps * 32, sizeof(CompSmem<C, 0>), st>>>(
                        z, sdf, rgb, normals, beta_param, beta_min, depth_scale, z_max, R, S, flags, weights,
                        rgb_values, depth_values, normal_map, bg_trans)));
    }
  }
  SVS_LAUNCH_OK();
  return SVS_OK;
}

extern "C" int svs_composite_backward(const float* z, const float* sdf, const float* rgb, const float* beta_param,
                                      float beta_min, const float* depth_scale, const float* z_max, int64_t R,
                                      int32_t S, int32_t flags, const float* d_rgb_values,
                                      const float* d_depth_values, const float* d_weights,
                                      const float* d_bg_trans, float* d_sdf, float* d_rgb, float* d_beta_param,
                                      void* stream) {
  SVS_CHECK_ARG(R >= 0 && S >= 1 && S <= 256, "svs_composite_backward: need 1 <= S <= 256 (got %d)", S);
  SVS_CHECK_ARG(z && sdf && d_sdf, "svs_composite_backward: z/sdf/d_sdf required");
  SVS_CHECK_ARG((flags & SVS_COMP_ABS_DENSITY) || beta_param, "svs_composite_backward: beta_param required");
  SVS_CHECK_ARG(!(flags & SVS_COMP_ZMAX_TAIL) || z_max, "svs_composite_backward: z_max required");
  SVS_CHECK_ARG(!d_rgb || rgb, "svs_composite_backward: rgb required for d_rgb");
  if (R == 0) return SVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps("composite_bwd", 0.0,
               (double)R * (4.0 * S * (3 + (rgb ? 3 : 0) + (d_weights ? 1 : 0) + (d_rgb ? 3 : 0)) + 32), st);
  const bool lean_ok = flags == SVS_COMP_FAST && rgb && d_rgb && d_rgb_values && d_depth_values && depth_scale && (S & 1) == 0 &&
                       S > 32 && S <= 256 &&
                       ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(sdf) | reinterpret_cast<uintptr_t>(rgb) |
                         reinterpret_cast<uintptr_t>(d_sdf) | reinterpret_cast<uintptr_t>(d_rgb) |
                         reinterpret_cast<uintptr_t>(d_weights ? d_weights : z)) & 7) == 0;
  if (lean_ok) {
    if (d_weights) {
      DISPATCH_C(S, (launch_bwd_lean<C, true>(z, sdf, rgb, beta_param, beta_min, depth_scale, R, S, d_rgb_values, d_depth_values, d_weights, d_sdf, d_rgb, d_beta_param, st)));
    } else {
      DISPATCH_C(S, (launch_bwd_lean<C, false>(z, sdf, rgb, beta_param, beta_min, depth_scale, R, S, d_rgb_values, d_depth_values, d_weights, d_sdf, d_rgb, d_beta_param, st)));
    }
  } else if (flags & SVS_COMP_FAST) {
    DISPATCH_C(S, (composite_bwd_kernel<C, true, 1><<<comp_grid(composite_bwd_kernel<C, true, 1>, sizeof(CompSmem<C, 1>), R), kCompWarps * 32, sizeof(CompSmem<C, 1>), st>>>(
                      z, sdf, rgb, beta_param, beta_min, depth_scale, z_max, R, S, flags, d_rgb_values,
                      d_depth_values, d_weights, d_bg_trans, d_sdf, d_rgb, d_beta_param)));
  } else {
    DISPATCH_C(S, (composite_bwd_kernel<C, false, 1><<<comp_grid(composite_bwd_kernel<C, false, 1>, sizeof(CompSmem<C, 1>), R), kCompWarps * 32, sizeof(CompSmem<C, 1>), st>>>(
                      z, sdf, rgb, beta_param, beta_min, depth_scale, z_max, R, S, flags, d_rgb_values,
                      d_depth_values, d_weights, d_bg_trans, d_sdf, d_rgb, d_beta_param)));
  }
  SVS_LAUNCH_OK();
  return SVS_OK;
}

extern "C" int svs_density_forward(const float* sdf, int64_t R, int32_t S, const float* beta_param, float beta_min,
                                   const float* beta_rows, int32_t abs_density, float* out, void* stream) {
  SVS_CHECK_ARG(sdf && out && R >= 0 && S >= 1, "svs_density_forward: bad arguments");
  SVS_CHECK_ARG(abs_density || beta_rows || beta_param, "svs_density_forward: beta required");
  int64_t n = R * S;
  if (n == 0) return SVS_OK;
  density_fwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(sdf, n, S, beta_param, beta_min,
                                                                              beta_rows, abs_density, out);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

extern "C" int svs_density_backward(const float* sdf, int64_t R, int32_t S, const float* beta_param, float beta_min,
                                    const float* beta_rows, int32_t abs_density, const float* d_out, float* d_sdf,
                                    float* d_beta_param, void* stream) {
  SVS_CHECK_ARG(sdf && d_out && d_sdf && R >= 0 && S >= 1, "svs_density_backward: bad arguments");
  SVS_CHECK_ARG(abs_density || beta_rows || beta_param, "svs_density_backward: beta required");
  int64_t n = R * S;
  if (n == 0) return SVS_OK;
  density_bwd_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
      sdf, n, S, beta_param, beta_min, beta_rows, abs_density, d_out, d_sdf, d_beta_param);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

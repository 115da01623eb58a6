// K1-K5: positional encoding, SDF MLP (forward, analytic input gradient, double backward), rendering
// MLP, weight-norm.  This file holds the layer orchestration and the element-wise kernels; the GEMM
// tiles come from the engine headers (mlp_gemm_fp32.cuh: fp32 SIMT parity engine).
//
// Math: SURVEY.md Appendix F.  Reference: volsdf/model/network.py:10-190, embedder.py:5-50.
#include "mlp_gemm_fp32.cuh"

namespace svs {

constexpr float kInvSqrt2 = 0.70710678118654752440f;
constexpr float kSqrt2 = 1.41421356237309504880f;

struct Layout {
  int L;
  int in[SVS_MAX_LAYERS], out[SVS_MAX_LAYERS], ldi[SVS_MAX_LAYERS];
  int64_t woff[SVS_MAX_LAYERS], boff[SVS_MAX_LAYERS];
  int64_t total;
  int ld0;   // padded width of the network input
  int H;     // padded width of hidden activations
  int ldy;   // padded width of the last layer's output
  int skip;  // skip layer or -1
  int pe_w;  // width of PE(x) for SDF nets
};

static int make_layout(const svs_mlp_desc* d, Layout* lo) {
  SVS_CHECK_ARG(d, "null descriptor");
  SVS_CHECK_ARG(d->n_layers >= 2 && d->n_layers <= SVS_MAX_LAYERS, "n_layers=%d out of range", d->n_layers);
  lo->L = d->n_layers;
  int64_t off = 0;
  int H = 0;
  for (int l = 0; l < lo->L; ++l) {
    SVS_CHECK_ARG(d->in_dim[l] > 0 && d->out_dim[l] > 0, "layer %d has empty dims", l);
    lo->in[l] = d->in_dim[l];
    lo->out[l] = d->out_dim[l];
    lo->ldi[l] = (int)round_up(d->in_dim[l], 4);
    lo->woff[l] = off;
    off += (int64_t)lo->out[l] * lo->ldi[l];
    lo->boff[l] = off;
    off += round_up(lo->out[l], 4);
    if (l > 0 && lo->ldi[l] > H) H = lo->ldi[l];
  }
  lo->total = off;
  lo->ld0 = lo->ldi[0];
  lo->H = H;
  lo->ldy = (int)round_up(lo->out[lo->L - 1], 4);
  lo->skip = (d->kind == SVS_NET_SDF) ? d->skip_layer : -1;
  lo->pe_w = d->d_in * (1 + 2 * d->n_freqs);
  if (d->kind == SVS_NET_SDF) {
    SVS_CHECK_ARG(d->d_in >= 1 && d->d_in <= 4, "d_in=%d unsupported", d->d_in);
    SVS_CHECK_ARG(lo->pe_w == lo->in[0], "in_dim[0]=%d != PE width %d", lo->in[0], lo->pe_w);
    SVS_CHECK_ARG(lo->skip < lo->L - 1 && lo->skip != 0, "skip layer %d unsupported", lo->skip);
    for (int l = 1; l < lo->L; ++l) {
      int expect = (l == lo->skip) ? lo->out[l - 1] + lo->pe_w : lo->out[l - 1];
      SVS_CHECK_ARG(lo->in[l] == expect, "layer %d: in_dim %d != %d", l, lo->in[l], expect);
    }
  } else {
    for (int l = 1; l < lo->L; ++l)
      SVS_CHECK_ARG(lo->in[l] == lo->out[l - 1], "layer %d: in_dim %d != out_dim %d", l, lo->in[l], lo->out[l - 1]);
  }
  return SVS_OK;
}

// --------------------------------------------------------------------------------------------------
// weight-norm forward / backward (network.py:64-65; torch._weight_norm: w = v * (g / |v|_row))
// --------------------------------------------------------------------------------------------------
struct PackArgs {
  int L;
  int in[SVS_MAX_LAYERS], out[SVS_MAX_LAYERS], ldi[SVS_MAX_LAYERS];
  int row0[SVS_MAX_LAYERS + 1];
  long long woff[SVS_MAX_LAYERS], boff[SVS_MAX_LAYERS];
  const float* g[SVS_MAX_LAYERS];
  const float* v[SVS_MAX_LAYERS];
  const float* b[SVS_MAX_LAYERS];
  float* dg[SVS_MAX_LAYERS];
  float* dv[SVS_MAX_LAYERS];
  float* db[SVS_MAX_LAYERS];
};

__device__ __forceinline__ int find_layer(const PackArgs& a, int row) {
  int l = 0;
  while (l + 1 < a.L && row >= a.row0[l + 1]) ++l;
  return l;
}

__global__ void __launch_bounds__(128) pack_weights_kernel(const PackArgs a, float* __restrict__ wbuf) {
  int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= a.row0[a.L]) return;
  int l = find_layer(a, row);
  int r = row - a.row0[l];
  const float* v = a.v[l] + (int64_t)r * a.in[l];
  float scale = 1.0f;
  if (a.g[l]) {
    float ss = 0.f;
    for (int k = lane; k < a.in[l]; k += 32) ss += v[k] * v[k];
    ss = warp_sum(ss);
    scale = a.g[l][r] / sqrtf(ss);
  }
  float* w = wbuf + a.woff[l] + (int64_t)r * a.ldi[l];
  for (int k = lane; k < a.ldi[l]; k += 32) w[k] = (k < a.in[l]) ? v[k] * scale : 0.f;
  if (lane == 0) wbuf[a.boff[l] + r] = a.b[l][r];
}

__global__ void __launch_bounds__(128)
param_grads_kernel(const PackArgs a, const float* __restrict__ dwbuf) {
  int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= a.row0[a.L]) return;
  int l = find_layer(a, row);
  int r = row - a.row0[l];
  const float* dw = dwbuf + a.woff[l] + (int64_t)r * a.ldi[l];
  float* dv = a.dv[l] + (int64_t)r * a.in[l];
  if (a.g[l]) {
    const float* v = a.v[l] + (int64_t)r * a.in[l];
    float ss = 0.f, dot = 0.f;
    for (int k = lane; k < a.in[l]; k += 32) {
      ss += v[k] * v[k];
      dot += dw[k] * v[k];
    }
    ss = warp_sum(ss);
    dot = warp_sum(dot);
    float inv = 1.0f / sqrtf(ss);
    float dgv = dot * inv;  // dL/dg = sum dW . v_hat
    float gs = a.g[l][r] * inv;
    for (int k = lane; k < a.in[l]; k += 32) dv[k] = gs * (dw[k] - dgv * v[k] * inv);
    if (lane == 0) a.dg[l][r] = dgv;
  } else {
    for (int k = lane; k < a.in[l]; k += 32) dv[k] = dw[k];
  }
  if (lane == 0) a.db[l][r] = dwbuf[a.boff[l] + r];
}

static void fill_pack(const Layout& lo, const svs_mlp_params* p, const svs_mlp_params* grads, PackArgs* a) {
  memset(a, 0, sizeof(*a));
  a->L = lo.L;
  int row = 0;
  for (int l = 0; l < lo.L; ++l) {
    a->in[l] = lo.in[l];
    a->out[l] = lo.out[l];
    a->ldi[l] = lo.ldi[l];
    a->row0[l] = row;
    row += lo.out[l];
    a->woff[l] = lo.woff[l];
    a->boff[l] = lo.boff[l];
    a->g[l] = p->g[l];
    a->v[l] = p->v[l];
    a->b[l] = p->b[l];
    if (grads) {
      a->dg[l] = grads->g[l];
      a->dv[l] = grads->v[l];
      a->db[l] = grads->b[l];
    }
  }
  a->row0[lo.L] = row;
}

// --------------------------------------------------------------------------------------------------
// element-wise kernels
// --------------------------------------------------------------------------------------------------

// PE(x)*scale -> dst[m, col_off + c], c < pe_w ; columns [pe_w, pad_to) zero-filled (embedder.py:10-36)
__global__ void pe_kernel(const float* __restrict__ x, int64_t P, int d_in, int n_freqs, float* __restrict__ dst,
                          int ld, int col_off, float scale, int pad_to) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t m = idx / pad_to;
  int c = (int)(idx - m * pad_to);
  if (m >= P) return;
  int pe_w = d_in * (1 + 2 * n_freqs);
  float v = 0.f;
  if (c < d_in) {
    v = x[m * d_in + c];
  } else if (c < pe_w) {
    int t = c - d_in;
    int k = t / (2 * d_in);
    int rem = t - k * 2 * d_in;
    int fn = rem / d_in, dim = rem - fn * d_in;
    float arg = x[m * d_in + dim] * (float)(1 << k);
    v = fn ? cosf(arg) : sinf(arg);
  }
  dst[m * ld + col_off + c] = v * scale;
}

// U[m,n] = s(A[m,n]) * wrow[n]  — start of the reverse sweep (p_{L-1} = W_{L-1}[0,:])
__global__ void u_last_kernel(const float* __restrict__ A, int lda, const float* __restrict__ wrow, int64_t P, int N,
                              float* __restrict__ U, int ldu) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t m = idx / N;
  int n = (int)(idx - m * N);
  if (m >= P) return;
  U[m * ldu + n] = dsoftplus_from_h(A[m * lda + n]) * wrow[n];
}

__device__ __forceinline__ float clamp_factor(float y0, float sphere, bool clamp_on) {
  // share of d(min(y0, sphere)) that flows to y0 (torch.minimum backward: 0.5/0.5 on exact ties)
  if (!clamp_on) return 1.f;
  return (y0 < sphere) ? 1.f : ((y0 == sphere) ? 0.5f : 0.f);
}

// g = J_PE(x)^T (p0 + e), sphere clamp of sdf and g (network.py:105-123,125-131)
__global__ void pe_grad_kernel(const float* __restrict__ x, const float* __restrict__ P0, const float* __restrict__ E,
                               int ld0, const float* __restrict__ y, int ldy, int64_t P, int d_in, int n_freqs,
                               float radius, float sph_scale, long long n_clamp, float* __restrict__ sdf,
                               float* __restrict__ grad) {
  int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= P) return;
  float xv[4], g[4];
  float nrm2 = 0.f;
  for (int d = 0; d < d_in; ++d) {
    xv[d] = x[m * d_in + d];
    nrm2 += xv[d] * xv[d];
  }
  const float* p = P0 + m * ld0;
  const float* e = E ? E + m * ld0 : nullptr;
  for (int d = 0; d < d_in; ++d) {
    float acc = p[d] + (e ? e[d] : 0.f);
    for (int k = 0; k < n_freqs; ++k) {
      float f = (float)(1 << k);
      float sn, cs;
      sincosf(xv[d] * f, &sn, &cs);
      int cs_col = d_in + k * 2 * d_in + d;  // sin column; cos column is + d_in
      float ps = p[cs_col] + (e ? e[cs_col] : 0.f);
      float pc = p[cs_col + d_in] + (e ? e[cs_col + d_in] : 0.f);
      acc += f * (cs * ps - sn * pc);
    }
    g[d] = acc;
  }
  float y0 = y[m * ldy];
  float out_sdf = y0;
  if (m < n_clamp && radius > 0.f) {
    float nrm = sqrtf(nrm2);
    float sphere = sph_scale * (radius - nrm);
    float w = clamp_factor(y0, sphere, true);
    out_sdf = fminf(y0, sphere);
    // the sphere branch only where it is taken (w < 1): at x = 0 its gradient is 0/0, and autograd's norm backward
    // masks |x| = 0 to a zero gradient
    if (w < 1.f)
      for (int d = 0; d < d_in; ++d) g[d] = w * g[d] + (1.f - w) * (nrm > 0.f ? -sph_scale * xv[d] / nrm : 0.f);
  }
  if (sdf) sdf[m] = out_sdf;
  if (grad)
    for (int d = 0; d < d_in; ++d) grad[m * d_in + d] = g[d];
}

// sdf[m] = clamp(y[m,0])  (get_sdf_vals, network.py:125-131)
__global__ void sdf_clamp_kernel(const float* __restrict__ x, const float* __restrict__ y, int ldy, int64_t P, int d_in,
                                 float radius, float sph_scale, long long n_clamp, float* __restrict__ sdf) {
  int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= P) return;
  float y0 = y[m * ldy];
  if (m < n_clamp && radius > 0.f) {
    float n2 = 0.f;
    for (int d = 0; d < d_in; ++d) n2 += x[m * d_in + d] * x[m * d_in + d];
    y0 = fminf(y0, sph_scale * (radius - sqrtf(n2)));
  }
  sdf[m] = y0;
}

// Q0 = J_PE(x) (w * d_grad) written at dst[m, col_off + c]*scale; w = clamp factor of the point
__global__ void pe_jvp_kernel(const float* __restrict__ x, const float* __restrict__ d_grad, const float* __restrict__ y,
                              int ldy, int64_t P, int d_in, int n_freqs, float radius, float sph_scale, long long n_clamp,
                              float* __restrict__ dst, int ld, int col_off, float scale, int pad_to) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t m = idx / pad_to;
  int c = (int)(idx - m * pad_to);
  if (m >= P) return;
  int pe_w = d_in * (1 + 2 * n_freqs);
  float w = 1.f;
  if (m < n_clamp && radius > 0.f) {
    float n2 = 0.f;
    for (int d = 0; d < d_in; ++d) n2 += x[m * d_in + d] * x[m * d_in + d];
    w = clamp_factor(y[m * ldy], sph_scale * (radius - sqrtf(n2)), true);
  }
  float v = 0.f;
  if (c < d_in) {
    v = d_grad[m * d_in + c];
  } else if (c < pe_w) {
    int t = c - d_in;
    int k = t / (2 * d_in);
    int rem = t - k * 2 * d_in;
    int fn = rem / d_in, dim = rem - fn * d_in;
    float f = (float)(1 << k);
    float arg = x[m * d_in + dim] * f;
    float j = fn ? (-f * sinf(arg)) : (f * cosf(arg));
    v = j * d_grad[m * d_in + dim];
  }
  dst[m * ld + col_off + c] = v * w * scale;
}

// dst[m, col_off + c] = src[m, c] * scale
__global__ void copy_cols_kernel(const float* __restrict__ src, int lds, int64_t P, int W, float* __restrict__ dst,
                                 int ldd, int col_off, float scale) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t m = idx / W;
  int c = (int)(idx - m * W);
  if (m >= P) return;
  dst[m * ldd + col_off + c] = src[m * lds + c] * scale;
}

// DY[m,c] = dy[m,c] (or 0) ; column 0 additionally receives w * d_sdf[m]
__global__ void dy_prep_kernel(const float* __restrict__ dy, const float* __restrict__ d_sdf, const float* __restrict__ x,
                               const float* __restrict__ y, int ldy, int64_t P, int n_out, int d_in, float radius,
                               float sph_scale, long long n_clamp, float* __restrict__ DY) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t m = idx / ldy;
  int c = (int)(idx - m * ldy);
  if (m >= P) return;
  float v = (dy && c < n_out) ? dy[m * ldy + c] : 0.f;
  if (c == 0 && d_sdf) {
    float w = 1.f;
    if (m < n_clamp && radius > 0.f) {
      float n2 = 0.f;
      for (int d = 0; d < d_in; ++d) n2 += x[m * d_in + d] * x[m * d_in + d];
      w = clamp_factor(y[m * ldy], sph_scale * (radius - sqrtf(n2)), true);
    }
    v += w * d_sdf[m];
  }
  DY[m * ldy + c] = v;
}

// rendering-net input: idr cat[points, PE(view), normals, feat] / nerf cat[PE(view), feat] (network.py:174-177)
__global__ void render_input_kernel(const float* __restrict__ points, const float* __restrict__ view, const float* __restrict__ normals,
                                    const float* __restrict__ feat, int ld_feat, int64_t P, int n_freqs, int idr, int F,
                                    float* __restrict__ RIN, int ld) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t m = idx / ld;
  int c = (int)(idx - m * ld);
  if (m >= P) return;
  int pe_v = 3 * (1 + 2 * n_freqs);
  int o_view = idr ? 3 : 0, o_n = o_view + pe_v, o_f = o_n + (idr ? 3 : 0);
  float v = 0.f;
  if (c < o_view) {
    v = points[m * 3 + c];
  } else if (c < o_n) {
    int t = c - o_view;
    if (t < 3) {
      v = view[m * 3 + t];
    } else {
      t -= 3;
      int k = t / 6, rem = t - 6 * k;
      int fn = rem / 3, dim = rem - 3 * fn;
      float arg = view[m * 3 + dim] * (float)(1 << k);
      v = fn ? cosf(arg) : sinf(arg);
    }
  } else if (c < o_f) {
    v = normals[m * 3 + (c - o_n)];
  } else if (c < o_f + F) {
    v = feat[m * ld_feat + (c - o_f)];
  }
  RIN[m * ld + c] = v;
}

// DZL[m,c] = d_rgb[m,c] * rgb (1 - rgb)   (sigmoid backward), padded to ld 4
__global__ void sigmoid_bwd_kernel(const float* __restrict__ rgb, const float* __restrict__ d_rgb, int64_t P, int n_out,
                                   float* __restrict__ DZL, int ld) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t m = idx / ld;
  int c = (int)(idx - m * ld);
  if (m >= P) return;
  float v = 0.f;
  if (c < n_out) {
    float s = rgb[m * n_out + c];
    v = d_rgb[m * n_out + c] * s * (1.f - s);
  }
  DZL[m * ld + c] = v;
}

// d_normals[m, 0..3) = DRIN[m, o_n + .] ; d_feat[m*ld_dfeat + j] = DRIN[m, o_f + j]
__global__ void render_scatter_kernel(const float* __restrict__ DRIN, int ld, int64_t P, int o_n, int has_n, int o_f, int F,
                                      float* __restrict__ d_normals, float* __restrict__ d_feat, int ld_dfeat) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int W = F + 3;
  int64_t m = idx / W;
  int c = (int)(idx - m * W);
  if (m >= P) return;
  if (c < 3) {
    if (has_n && d_normals) d_normals[m * 3 + c] = DRIN[m * ld + o_n + c];
  } else if (d_feat) {
    d_feat[m * ld_dfeat + (c - 3)] = DRIN[m * ld + o_f + (c - 3)];
  }
}

static inline unsigned blocks_for(int64_t n) { return (unsigned)cdiv(n, 256); }

}  // namespace svs

#include "mlp_tc_chains.cuh"

using namespace svs;

// ====================================================================================================
// C ABI
// ====================================================================================================

static bool engine_ok(int engine) { return engine == SVS_ENGINE_FP32 || engine == SVS_ENGINE_TC || engine == SVS_ENGINE_TC_SPLIT; }
static bool is_tc(int engine) { return engine == SVS_ENGINE_TC || engine == SVS_ENGINE_TC_SPLIT; }
static bool is_split(int engine) { return engine == SVS_ENGINE_TC_SPLIT; }

#if defined(SVS_F3_TRACE) || defined(SVS_CHAIN_TRACE)
// measurement builds only: copies the trace of tc_fwd3_kernel out and resets it (tools/f3_trace.py)
extern "C" int svs_dbg_f3_trace(long long* out, int n) {
  cudaDeviceSynchronize();
  if (n > 16384) n = 16384;
  cudaMemcpyFromSymbol(out, svs::tc::g_f3_trace, sizeof(long long) * n);
  static long long zeros[16384];
  cudaMemcpyToSymbol(svs::tc::g_f3_trace, zeros, sizeof(zeros));
  return n;
}
#endif

extern "C" int64_t svs_mlp_wbuf_floats(const svs_mlp_desc* d, int engine) {
  Layout lo;
  if (make_layout(d, &lo) != SVS_OK || !engine_ok(engine)) return -1;
  if (is_tc(engine)) return tc::wbuf_floats_tc(d, lo, is_split(engine));
  return lo.total;
}

extern "C" int svs_mlp_prepare(const svs_mlp_desc* d, const svs_mlp_params* p, float* wbuf, int engine, void* stream) {
  Layout lo;
  SVS_TRY(make_layout(d, &lo));
  SVS_CHECK_ARG(p && wbuf, "svs_mlp_prepare: null pointer");
  SVS_CHECK_ARG(engine_ok(engine), "svs_mlp_prepare: engine %d not built", engine);
  for (int l = 0; l < lo.L; ++l) {
    SVS_CHECK_ARG(p->v[l] && p->b[l], "svs_mlp_prepare: layer %d missing weight/bias", l);
    SVS_CHECK_ARG((d->weight_norm != 0) == (p->g[l] != nullptr), "svs_mlp_prepare: layer %d weight_g mismatch", l);
  }
  PackArgs a;
  fill_pack(lo, p, nullptr, &a);
  pack_weights_kernel<<<(unsigned)cdiv(a.row0[lo.L], 4), 128, 0, (cudaStream_t)stream>>>(a, wbuf);
  SVS_LAUNCH_OK();
  if (is_tc(engine)) SVS_TRY(tc::pack_images(d, lo, wbuf, (cudaStream_t)stream, is_split(engine)));
  return SVS_OK;
}

extern "C" int svs_mlp_param_grads(const svs_mlp_desc* d, const svs_mlp_params* p, const float* wbuf,
                                   const float* dwbuf, const svs_mlp_params* grads, void* stream) {
  Layout lo;
  SVS_TRY(make_layout(d, &lo));
  SVS_CHECK_ARG(p && dwbuf && grads, "svs_mlp_param_grads: null pointer");
  (void)wbuf;
  for (int l = 0; l < lo.L; ++l) {
    SVS_CHECK_ARG(grads->v[l] && grads->b[l], "svs_mlp_param_grads: layer %d missing grad buffers", l);
    SVS_CHECK_ARG(!p->g[l] || grads->g[l], "svs_mlp_param_grads: layer %d missing weight_g grad", l);
  }
  PackArgs a;
  fill_pack(lo, p, grads, &a);
  param_grads_kernel<<<(unsigned)cdiv(a.row0[lo.L], 4), 128, 0, (cudaStream_t)stream>>>(a, dwbuf);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

extern "C" int svs_embed(const float* x, int64_t P, int32_t d_in, int32_t n_freqs, float* out, void* stream) {
  SVS_CHECK_ARG(x && out && P >= 0 && d_in >= 1 && d_in <= 4 && n_freqs >= 0 && n_freqs <= 16, "svs_embed: bad arguments");
  if (P == 0) return SVS_OK;
  int w = d_in * (1 + 2 * n_freqs);
  pe_kernel<<<blocks_for(P * w), 256, 0, (cudaStream_t)stream>>>(x, P, d_in, n_freqs, out, w, 0, 1.f, w);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

// ---------------------------------------------------------------------------------------------------
// SDF network
// ---------------------------------------------------------------------------------------------------

extern "C" int32_t svs_sdf_ldy(const svs_mlp_desc* d) {
  Layout lo;
  if (make_layout(d, &lo) != SVS_OK) return -1;
  return lo.ldy;
}

extern "C" int64_t svs_sdf_saved_floats(const svs_mlp_desc* d, int64_t P, int engine) {
  Layout lo;
  if (make_layout(d, &lo) != SVS_OK || !engine_ok(engine)) return -1;
  if (is_tc(engine)) {
    tc::SdfSaved sv;
    tc::map_sdf_saved(lo, P, nullptr, &sv);
    return sv.bytes / 4;
  }
  return P * ((int64_t)lo.ld0 + 2 * (int64_t)(lo.L - 1) * lo.H);
}

extern "C" int64_t svs_sdf_ws_floats(const svs_mlp_desc* d, int64_t P, int with_grad, int engine) {
  Layout lo;
  if (make_layout(d, &lo) != SVS_OK || !engine_ok(engine)) return -1;
  if (is_tc(engine)) return with_grad ? svs_sdf_saved_floats(d, P, engine) : 4;
  if (with_grad) return svs_sdf_saved_floats(d, P, engine) + P * 2 * (int64_t)lo.ld0;
  return P * ((int64_t)lo.ld0 + 2 * (int64_t)lo.H + lo.ldy);
}

extern "C" int64_t svs_sdf_bwd_ws_floats(const svs_mlp_desc* d, int64_t P, int engine) {
  Layout lo;
  if (make_layout(d, &lo) != SVS_OK || !engine_ok(engine)) return -1;
  if (is_tc(engine)) {
    tc::SdfBwdWs bw;
    tc::map_sdf_bwd(lo, P, nullptr, &bw);
    return bw.bytes / 4;
  }
  return P * ((int64_t)lo.ld0 + (int64_t)(4 + lo.L - 1) * lo.H + lo.ldy);
}

static int check_sdf(const svs_mlp_desc* d, const Layout& lo, int engine, int64_t P = 0) {
  SVS_CHECK_ARG(d->kind == SVS_NET_SDF, "descriptor is not an SDF net");
  SVS_CHECK_ARG(P >= 0 && P <= 2147483647LL - 128, "point count %lld exceeds the 2^31 - 1 rows one launch can address", (long long)P);
  SVS_CHECK_ARG(engine_ok(engine), "engine %d not built", engine);
  if (engine == SVS_ENGINE_FP32)
    for (int l = 1; l < lo.L; ++l) SVS_CHECK_ARG(lo.ldi[l] == lo.H, "hidden widths must be uniform (layer %d)", l);
  return SVS_OK;
}

extern "C" int svs_sdf_forward(const svs_mlp_desc* d, const float* wbuf, const float* x, int64_t P, float* y,
                               float* sdf, float* ws, int engine, void* stream) {
  Layout lo;
  SVS_TRY(make_layout(d, &lo));
  SVS_TRY(check_sdf(d, lo, engine, P));
  SVS_CHECK_ARG(wbuf && x && ws && (y || sdf) && P >= 0, "svs_sdf_forward: bad arguments");
  if (P == 0) return SVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (is_tc(engine)) return tc::sdf_forward(d, lo, wbuf, x, P, y, sdf, st, is_split(engine));
  float* A0 = ws;
  float* B[2] = {A0 + P * lo.ld0, A0 + P * lo.ld0 + P * lo.H};
  float* Y = y ? y : (B[1] + P * lo.H);
  pe_kernel<<<blocks_for(P * lo.ld0), 256, 0, st>>>(x, P, d->d_in, d->n_freqs, A0, lo.ld0, 0, 1.f, lo.ld0);
  SVS_LAUNCH_OK();
  const float* cur = A0;
  int ldc_cur = lo.ld0;
  for (int l = 0; l < lo.L - 1; ++l) {
    float* dst = B[l & 1];
    GemmArgs g = {};
    g.A = cur; g.lda = ldc_cur;
    g.B = wbuf + lo.woff[l]; g.ldb = lo.ldi[l];
    g.C = dst; g.ldc = lo.H;
    g.M = (int)P; g.N = lo.out[l]; g.K = lo.in[l];
    g.bias = wbuf + lo.boff[l];
    g.scale = (l + 1 == lo.skip) ? kInvSqrt2 : 1.f;
    SVS_TRY((launch_gemm<true, EPI_BIAS_SOFTPLUS>(g, st)));
    if (l + 1 == lo.skip) {
      pe_kernel<<<blocks_for(P * lo.pe_w), 256, 0, st>>>(x, P, d->d_in, d->n_freqs, dst, lo.H, lo.out[l], kInvSqrt2, lo.pe_w);
      SVS_LAUNCH_OK();
    }
    cur = dst;
    ldc_cur = lo.H;
  }
  {
    int l = lo.L - 1;
    GemmArgs g = {};
    g.A = cur; g.lda = ldc_cur;
    g.B = wbuf + lo.woff[l]; g.ldb = lo.ldi[l];
    g.C = Y; g.ldc = lo.ldy;
    g.M = (int)P; g.N = y ? lo.out[l] : 1; g.K = lo.in[l];
    g.bias = wbuf + lo.boff[l];
    SVS_TRY((launch_gemm<true, EPI_BIAS>(g, st)));
  }
  if (sdf) {
    sdf_clamp_kernel<<<blocks_for(P), 256, 0, st>>>(x, Y, lo.ldy, P, d->d_in, d->sphere_radius, d->sphere_scale, P, sdf);
    SVS_LAUNCH_OK();
  }
  return SVS_OK;
}

struct SdfBuffers {
  float* A[SVS_MAX_LAYERS];  // A[0] (ld0), A[1..L-1] (H)
  float* U[SVS_MAX_LAYERS];  // U[0..L-2] (H)
};

static void map_saved(const Layout& lo, int64_t P, float* base, SdfBuffers* b) {
  b->A[0] = base;
  float* p = base + P * lo.ld0;
  for (int l = 1; l < lo.L; ++l, p += P * lo.H) b->A[l] = p;
  for (int l = 0; l < lo.L - 1; ++l, p += P * lo.H) b->U[l] = p;
}

extern "C" int svs_sdf_outputs_forward(const svs_mlp_desc* d, const float* wbuf, const float* x, int64_t P, int64_t n_clamped,
                                       float* y, float* sdf, float* grad, float* saved, float* ws, int engine,
                                       void* stream) {
  Layout lo;
  SVS_TRY(make_layout(d, &lo));
  SVS_TRY(check_sdf(d, lo, engine, P));
  SVS_CHECK_ARG(wbuf && x && y && ws && P >= 0, "svs_sdf_outputs_forward: bad arguments");
  if (P == 0) return SVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (is_tc(engine)) {
    if (!grad && !saved) {
      SVS_TRY(tc::sdf_forward(d, lo, wbuf, x, P, y, nullptr, st, is_split(engine)));
      if (sdf) {
        sdf_clamp_kernel<<<blocks_for(P), 256, 0, st>>>(x, y, lo.ldy, P, d->d_in, d->sphere_radius, d->sphere_scale,
                                                        n_clamped, sdf);
        SVS_LAUNCH_OK();
      }
      return SVS_OK;
    }
    return tc::sdf_outputs_forward(d, lo, wbuf, x, P, n_clamped, y, sdf, grad, saved ? (void*)saved : (void*)ws, st, is_split(engine));
  }
  SdfBuffers b;
  float* scratch = ws;
  if (saved) {
    map_saved(lo, P, saved, &b);
  } else {
    map_saved(lo, P, ws, &b);
    scratch = ws + svs_sdf_saved_floats(d, P, SVS_ENGINE_FP32);
  }
  float* P0 = scratch;
  float* E = scratch + P * lo.ld0;
  const int L = lo.L;
  // ---- forward, keeping every activation ----
  pe_kernel<<<blocks_for(P * lo.ld0), 256, 0, st>>>(x, P, d->d_in, d->n_freqs, b.A[0], lo.ld0, 0, 1.f, lo.ld0);
  SVS_LAUNCH_OK();
  if (lo.skip > 0) {
    pe_kernel<<<blocks_for(P * lo.pe_w), 256, 0, st>>>(x, P, d->d_in, d->n_freqs, b.A[lo.skip], lo.H,
                                                       lo.out[lo.skip - 1], kInvSqrt2, lo.pe_w);
    SVS_LAUNCH_OK();
  }
  for (int l = 0; l < L - 1; ++l) {
    GemmArgs g = {};
    g.A = b.A[l]; g.lda = l ? lo.H : lo.ld0;
    g.B = wbuf + lo.woff[l]; g.ldb = lo.ldi[l];
    g.C = b.A[l + 1]; g.ldc = lo.H;
    g.M = (int)P; g.N = lo.out[l]; g.K = lo.in[l];
    g.bias = wbuf + lo.boff[l];
    g.scale = (l + 1 == lo.skip) ? kInvSqrt2 : 1.f;
    SVS_TRY((launch_gemm<true, EPI_BIAS_SOFTPLUS>(g, st)));
  }
  {
    int l = L - 1;
    GemmArgs g = {};
    g.A = b.A[l]; g.lda = lo.H;
    g.B = wbuf + lo.woff[l]; g.ldb = lo.ldi[l];
    g.C = y; g.ldc = lo.ldy;
    g.M = (int)P; g.N = lo.out[l]; g.K = lo.in[l];
    g.bias = wbuf + lo.boff[l];
    SVS_TRY((launch_gemm<true, EPI_BIAS>(g, st)));
  }
  if (!grad && !saved) {
    if (sdf) {
      sdf_clamp_kernel<<<blocks_for(P), 256, 0, st>>>(x, y, lo.ldy, P, d->d_in, d->sphere_radius, d->sphere_scale,
                                                      n_clamped, sdf);
      SVS_LAUNCH_OK();
    }
    return SVS_OK;
  }
  // ---- reverse sweep: p_l = W_l^T (s_l * p_{l+1}), kept as U_l = s_l * p~_{l+1} ----
  u_last_kernel<<<blocks_for(P * lo.out[L - 2]), 256, 0, st>>>(b.A[L - 1], lo.H, wbuf + lo.woff[L - 1], P, lo.out[L - 2],
                                                              b.U[L - 2], lo.H);
  SVS_LAUNCH_OK();
  for (int l = L - 2; l >= 1; --l) {
    GemmArgs g = {};
    g.A = b.U[l]; g.lda = lo.H;
    g.B = wbuf + lo.woff[l]; g.ldb = lo.ldi[l];
    g.C = b.U[l - 1]; g.ldc = lo.H;
    g.M = (int)P; g.N = lo.in[l]; g.K = lo.out[l];
    g.E1 = b.A[l]; g.lde1 = lo.H;
    g.C2 = E; g.ldc2 = lo.ld0;
    g.scale = (l == lo.skip) ? kInvSqrt2 : 1.f;
    g.hscale = (l == lo.skip) ? kSqrt2 : 1.f;
    g.n_split = lo.out[l - 1];
    SVS_TRY((launch_gemm<false, EPI_REVERSE>(g, st)));
  }
  {
    GemmArgs g = {};
    g.A = b.U[0]; g.lda = lo.H;
    g.B = wbuf + lo.woff[0]; g.ldb = lo.ldi[0];
    g.C = P0; g.ldc = lo.ld0;
    g.M = (int)P; g.N = lo.ldi[0]; g.K = lo.out[0];
    SVS_TRY((launch_gemm<false, EPI_PLAIN>(g, st)));
  }
  pe_grad_kernel<<<blocks_for(P), 256, 0, st>>>(x, P0, lo.skip > 0 ? E : nullptr, lo.ld0, y, lo.ldy, P, d->d_in,
                                                d->n_freqs, d->sphere_radius, d->sphere_scale, n_clamped, sdf, grad);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

extern "C" int svs_sdf_outputs_backward(const svs_mlp_desc* d, const float* wbuf, const float* x, int64_t P, int64_t n_clamped,
                                        const float* saved, const float* y, const float* dy, const float* d_sdf,
                                        const float* d_grad, float* dwbuf, float* ws, int engine, void* stream) {
  Layout lo;
  SVS_TRY(make_layout(d, &lo));
  SVS_TRY(check_sdf(d, lo, engine, P));
  SVS_CHECK_ARG(wbuf && x && saved && y && dwbuf && ws && P >= 0, "svs_sdf_outputs_backward: bad arguments");
  if (P == 0) return SVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (is_tc(engine))
    return tc::sdf_outputs_backward(d, lo, wbuf, x, P, n_clamped, saved, y, dy, d_sdf, d_grad, dwbuf, ws, st);
  const int L = lo.L;
  SdfBuffers b;
  map_saved(lo, P, const_cast<float*>(saved), &b);
  float* Q0 = ws;
  float* QA[2] = {Q0 + P * lo.ld0, Q0 + P * lo.ld0 + P * lo.H};
  float* DZ[2] = {QA[1] + P * lo.H, QA[1] + 2 * P * lo.H};
  float* ZETA[SVS_MAX_LAYERS];
  float* p = DZ[1] + P * lo.H;
  for (int l = 0; l < L - 1; ++l, p += P * lo.H) ZETA[l] = p;
  float* DY = p;
  const bool have_tangent = d_grad != nullptr;
  const bool have_top = (dy != nullptr) || (d_sdf != nullptr);
  if (!have_tangent && !have_top) return SVS_OK;
  const float radius = d->sphere_radius;

  // ---- tangent sweep (adjoint of the reverse sweep): q_{l+1} = s_l * (W_l q_l), zeta_l, dW_l += U_l^T q_l ----
  if (have_tangent) {
    pe_jvp_kernel<<<blocks_for(P * lo.ld0), 256, 0, st>>>(x, d_grad, y, lo.ldy, P, d->d_in, d->n_freqs, radius,
                                                          d->sphere_scale, n_clamped, Q0, lo.ld0, 0, 1.f, lo.ld0);
    SVS_LAUNCH_OK();
    const float* q_in = Q0;
    int ld_q = lo.ld0;
    for (int l = 0; l < L - 1; ++l) {
      float* q_out = QA[l & 1];
      GemmArgs g = {};
      g.A = q_in; g.lda = ld_q;
      g.B = wbuf + lo.woff[l]; g.ldb = lo.ldi[l];
      g.C = ZETA[l]; g.ldc = lo.H;
      g.M = (int)P; g.N = lo.out[l]; g.K = lo.in[l];
      g.E1 = b.A[l + 1]; g.lde1 = lo.H;
      g.E2 = b.U[l]; g.lde2 = lo.H;
      g.C2 = q_out; g.ldc2 = lo.H;
      g.scale = (l + 1 == lo.skip) ? kInvSqrt2 : 1.f;
      g.hscale = (l + 1 == lo.skip) ? kSqrt2 : 1.f;
      SVS_TRY((launch_gemm<true, EPI_TANGENT>(g, st)));
      if (l + 1 == lo.skip) {
        copy_cols_kernel<<<blocks_for(P * lo.pe_w), 256, 0, st>>>(Q0, lo.ld0, P, lo.pe_w, q_out, lo.H, lo.out[l], kInvSqrt2);
        SVS_LAUNCH_OK();
      }
      SVS_TRY(launch_gemm_tn(b.U[l], lo.H, q_in, ld_q, dwbuf + lo.woff[l], lo.ldi[l], P, lo.out[l], lo.ldi[l], st));
      q_in = q_out;
      ld_q = lo.H;
    }
    // y_0 = W_{L-1}[0,:] a_{L-1} + b : the tangent reaches row 0 of the last weight directly
    SVS_TRY(launch_colsum(q_in, lo.H, P, lo.in[L - 1], dwbuf + lo.woff[L - 1], st));
  }

  // ---- ordinary backward ----
  const float* dz = nullptr;
  int cur = 0;
  if (have_top) {
    dy_prep_kernel<<<blocks_for(P * lo.ldy), 256, 0, st>>>(dy, d_sdf, x, y, lo.ldy, P, lo.out[L - 1], d->d_in, radius,
                                                           d->sphere_scale, n_clamped, DY);
    SVS_LAUNCH_OK();
    int l = L - 1;
    SVS_TRY(launch_gemm_tn(DY, lo.ldy, b.A[l], lo.H, dwbuf + lo.woff[l], lo.ldi[l], P, lo.out[l], lo.ldi[l], st));
    SVS_TRY(launch_colsum(DY, lo.ldy, P, lo.out[l], dwbuf + lo.boff[l], st));
    GemmArgs g = {};
    g.A = DY; g.lda = lo.ldy;
    g.B = wbuf + lo.woff[l]; g.ldb = lo.ldi[l];
    g.C = DZ[cur]; g.ldc = lo.H;
    g.M = (int)P; g.N = lo.out[l - 1]; g.K = lo.out[l];
    g.E1 = b.A[l]; g.lde1 = lo.H;
    g.E2 = have_tangent ? ZETA[l - 1] : nullptr; g.lde2 = lo.H;
    g.scale = 1.f; g.hscale = 1.f;
    SVS_TRY((launch_gemm<false, EPI_BACKWARD>(g, st)));
    dz = DZ[cur];
    cur ^= 1;
  } else {
    dz = ZETA[L - 2];
  }
  for (int l = L - 2; l >= 0; --l) {
    SVS_TRY(launch_gemm_tn(dz, lo.H, b.A[l], l ? lo.H : lo.ld0, dwbuf + lo.woff[l], lo.ldi[l], P, lo.out[l], lo.ldi[l], st));
    SVS_TRY(launch_colsum(dz, lo.H, P, lo.out[l], dwbuf + lo.boff[l], st));
    if (l == 0) break;
    GemmArgs g = {};
    g.A = dz; g.lda = lo.H;
    g.B = wbuf + lo.woff[l]; g.ldb = lo.ldi[l];
    g.C = DZ[cur]; g.ldc = lo.H;
    g.M = (int)P; g.N = lo.out[l - 1]; g.K = lo.out[l];
    g.E1 = b.A[l]; g.lde1 = lo.H;
    g.E2 = have_tangent ? ZETA[l - 1] : nullptr; g.lde2 = lo.H;
    g.scale = (l == lo.skip) ? kInvSqrt2 : 1.f;
    g.hscale = (l == lo.skip) ? kSqrt2 : 1.f;
    SVS_TRY((launch_gemm<false, EPI_BACKWARD>(g, st)));
    dz = DZ[cur];
    cur ^= 1;
  }
  return SVS_OK;
}

// ---------------------------------------------------------------------------------------------------
// rendering network
// ---------------------------------------------------------------------------------------------------

static int check_render(const svs_mlp_desc* d, const Layout& lo, int engine, int* pe_v, int* F, int64_t P = 0) {
  SVS_CHECK_ARG(P >= 0 && P <= 2147483647LL - 128, "point count %lld exceeds the 2^31 - 1 rows one launch can address", (long long)P);
  SVS_CHECK_ARG(d->kind == SVS_NET_RENDER, "descriptor is not a rendering net");
  SVS_CHECK_ARG(engine_ok(engine), "engine %d not built", engine);
  *pe_v = 3 * (1 + 2 * d->n_freqs);
  int fixed = (d->render_mode == SVS_RENDER_IDR) ? 6 + *pe_v : *pe_v;
  *F = lo.in[0] - fixed;
  SVS_CHECK_ARG(*F > 0, "rendering net input %d too small for mode %d", lo.in[0], d->render_mode);
  for (int l = 1; l < lo.L; ++l) SVS_CHECK_ARG(lo.ldi[l] <= lo.H, "layer width");
  return SVS_OK;
}

extern "C" int64_t svs_render_saved_floats(const svs_mlp_desc* d, int64_t P, int engine) {
  Layout lo;
  if (make_layout(d, &lo) != SVS_OK || !engine_ok(engine)) return -1;
  if (is_tc(engine)) {
    tc::WImages wi;
    if (tc::make_wimages(d, lo, &wi, nullptr, nullptr) != SVS_OK) return -1;
    tc::RenderSaved sv;
    tc::map_render_saved(lo, wi, P, nullptr, &sv);
    return sv.bytes / 4;
  }
  return P * ((int64_t)lo.ld0 + (int64_t)(lo.L - 1) * lo.H);
}

extern "C" int64_t svs_render_ws_floats(const svs_mlp_desc* d, int64_t P, int engine) {
  Layout lo;
  if (make_layout(d, &lo) != SVS_OK || !engine_ok(engine)) return -1;
  if (is_tc(engine)) {
    tc::RenderWs rw;
    tc::map_render_ws(lo, P, nullptr, &rw);
    return rw.bytes / 4;
  }
  return P * ((int64_t)lo.ld0 + 2 * (int64_t)lo.H + 4);
}

extern "C" int svs_render_forward(const svs_mlp_desc* d, const float* wbuf, const float* points, const float* view_dirs,
                                  const float* normals, const float* feat, int32_t ld_feat, int64_t P, float* rgb,
                                  float* saved, int engine, void* stream) {
  Layout lo;
  SVS_TRY(make_layout(d, &lo));
  int pe_v, F;
  SVS_TRY(check_render(d, lo, engine, &pe_v, &F, P));
  const int idr = d->render_mode == SVS_RENDER_IDR;
  SVS_CHECK_ARG(wbuf && view_dirs && feat && rgb && saved && P >= 0, "svs_render_forward: bad arguments");
  SVS_CHECK_ARG(!idr || (points && normals), "svs_render_forward: idr mode needs points and normals");
  SVS_CHECK_ARG(ld_feat >= F, "svs_render_forward: ld_feat %d < feature width %d", ld_feat, F);
  if (P == 0) return SVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (is_tc(engine))
    return tc::render_forward(d, lo, wbuf, points, view_dirs, normals, feat, ld_feat, P, rgb, saved, st, is_split(engine));
  const int L = lo.L;
  float* RIN = saved;
  float* Bv[SVS_MAX_LAYERS];
  float* p = saved + P * lo.ld0;
  for (int l = 1; l < L; ++l, p += P * lo.H) Bv[l] = p;
  render_input_kernel<<<blocks_for(P * lo.ld0), 256, 0, st>>>(points, view_dirs, normals, feat, ld_feat, P, d->n_freqs,
                                                              idr, F, RIN, lo.ld0);
  SVS_LAUNCH_OK();
  const float* cur = RIN;
  int ld_cur = lo.ld0;
  for (int l = 0; l < L - 1; ++l) {
    GemmArgs g = {};
    g.A = cur; g.lda = ld_cur;
    g.B = wbuf + lo.woff[l]; g.ldb = lo.ldi[l];
    g.C = Bv[l + 1]; g.ldc = lo.H;
    g.M = (int)P; g.N = lo.out[l]; g.K = lo.in[l];
    g.bias = wbuf + lo.boff[l];
    SVS_TRY((launch_gemm<true, EPI_BIAS_RELU>(g, st)));
    cur = Bv[l + 1];
    ld_cur = lo.H;
  }
  {
    int l = L - 1;
    GemmArgs g = {};
    g.A = cur; g.lda = ld_cur;
    g.B = wbuf + lo.woff[l]; g.ldb = lo.ldi[l];
    g.C = rgb; g.ldc = lo.out[l];
    g.M = (int)P; g.N = lo.out[l]; g.K = lo.in[l];
    g.bias = wbuf + lo.boff[l];
    SVS_TRY((launch_gemm<true, EPI_BIAS_SIGMOID>(g, st)));
  }
  return SVS_OK;
}

extern "C" int svs_render_backward(const svs_mlp_desc* d, const float* wbuf, int64_t P, const float* saved,
                                   const float* rgb, const float* d_rgb, float* d_normals, float* d_feat,
                                   int32_t ld_dfeat, float* dwbuf, float* ws, int engine, void* stream) {
  Layout lo;
  SVS_TRY(make_layout(d, &lo));
  int pe_v, F;
  SVS_TRY(check_render(d, lo, engine, &pe_v, &F, P));
  const int idr = d->render_mode == SVS_RENDER_IDR;
  SVS_CHECK_ARG(wbuf && saved && rgb && d_rgb && dwbuf && ws && P >= 0, "svs_render_backward: bad arguments");
  if (P == 0) return SVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (is_tc(engine))
    return tc::render_backward(d, lo, wbuf, P, saved, rgb, d_rgb, d_normals, d_feat, ld_dfeat, dwbuf, ws, st);
  const int L = lo.L;
  const float* RIN = saved;
  const float* Bv[SVS_MAX_LAYERS];
  const float* p = saved + P * lo.ld0;
  for (int l = 1; l < L; ++l, p += P * lo.H) Bv[l] = p;
  float* DZ[2] = {ws, ws + P * lo.H};
  float* DRIN = ws + 2 * P * lo.H;
  float* DZL = DRIN + P * lo.ld0;
  const int n_out = lo.out[L - 1];
  SVS_CHECK_ARG(n_out <= 4, "svs_render_backward: d_out=%d unsupported", n_out);
  sigmoid_bwd_kernel<<<blocks_for(P * 4), 256, 0, st>>>(rgb, d_rgb, P, n_out, DZL, 4);
  SVS_LAUNCH_OK();
  const float* dz = DZL;
  int ld_dz = 4;
  int cur = 0;
  for (int l = L - 1; l >= 0; --l) {
    const float* a_in = l ? Bv[l] : RIN;
    int ld_a = l ? lo.H : lo.ld0;
    SVS_TRY(launch_gemm_tn(dz, ld_dz, a_in, ld_a, dwbuf + lo.woff[l], lo.ldi[l], P, lo.out[l], lo.ldi[l], st));
    SVS_TRY(launch_colsum(dz, ld_dz, P, lo.out[l], dwbuf + lo.boff[l], st));
    GemmArgs g = {};
    g.A = dz; g.lda = ld_dz;
    g.B = wbuf + lo.woff[l]; g.ldb = lo.ldi[l];
    g.M = (int)P; g.K = lo.out[l];
    if (l > 0) {
      g.C = DZ[cur]; g.ldc = lo.H;
      g.N = lo.in[l];
      g.E1 = Bv[l]; g.lde1 = lo.H;
      SVS_TRY((launch_gemm<false, EPI_RELU_BWD>(g, st)));
      dz = DZ[cur];
      ld_dz = lo.H;
      cur ^= 1;
    } else if (d_normals || d_feat) {
      g.C = DRIN; g.ldc = lo.ld0;
      g.N = lo.in[0];
      SVS_TRY((launch_gemm<false, EPI_PLAIN>(g, st)));
      int o_view = idr ? 3 : 0, o_n = o_view + pe_v, o_f = o_n + (idr ? 3 : 0);
      render_scatter_kernel<<<blocks_for(P * (F + 3)), 256, 0, st>>>(DRIN, lo.ld0, P, o_n, idr, o_f, F, d_normals,
                                                                     d_feat, ld_dfeat);
      SVS_LAUNCH_OK();
    }
  }
  return SVS_OK;
}

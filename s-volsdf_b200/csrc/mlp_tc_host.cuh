// Host side of SVS_ENGINE_TC: fp16 weight images, buffer layouts, chain / weight-gradient job builders.
// Included by mlp.cu after its `Layout` helpers.  Reference semantics: volsdf/model/network.py:71-123,170-190.
#pragma once
#include "mlp_tc.cuh"

#ifndef SVS_TC_NW_FWD
#define SVS_TC_NW_FWD 16
#endif
#ifndef SVS_TC_NW_RFWD
#define SVS_TC_NW_RFWD 16
#endif
#ifndef SVS_TC_NW_REV
#define SVS_TC_NW_REV 16
#endif
#ifndef SVS_TC_NW_RBWD
#define SVS_TC_NW_RBWD 16
#endif
#ifndef SVS_TC_NW_TAN
#define SVS_TC_NW_TAN 16
#endif
#ifndef SVS_TC_NW_BWD
#define SVS_TC_NW_BWD 16
#endif
#include "mlp_tc_dw.cuh"
#include "mlp_tc_fwd2.cuh"
#include "mlp_tc_fwd3.cuh"

namespace svs {
namespace tc {

static inline int r16(int v) { return (int)round_up(v, 16); }
static inline int kb_of(int cols) { return (int)cdiv(cols, 64); }

// fp16 image region of wbuf (bytes, relative to the start of the region); every image is 1024-byte aligned
struct WImages {
  int64_t fwd[SVS_MAX_LAYERS];   // B[n = out][k = in]   (last SDF layer: the feature rows 1..out-1)
  int64_t bwd[SVS_MAX_LAYERS];   // B[n = in][k = out]   (rendering layer 0: the feature columns)
  int64_t fwd_sdf;               // SDF last layer, row 0 padded to 16 rows
  int64_t fwd_lo[SVS_MAX_LAYERS], fwd_sdf_lo;   // split engine: images of W - fp16(W) (behind all regular images)
  int64_t fwd_bt[SVS_MAX_LAYERS];               // split engine, SDF softplus layers: fp32 table bias * 100 log2(e), n_pad entries
  int64_t bwd_small;             // rendering layer 0: the non-feature input columns
  int fwd_npad[SVS_MAX_LAYERS], fwd_kb[SVS_MAX_LAYERS];
  int bwd_npad[SVS_MAX_LAYERS], bwd_kb[SVS_MAX_LAYERS];
  int small_npad;
  int F, n_small;                // rendering net: feature width and the number of other input columns
  int64_t bytes;
};

struct PackImg {
  long long dst;      // byte offset in the image region
  int n_pad, kb;      // image rows, 64-column blocks
  int layer;
  int transpose;      // 0: value(r, c) = W[r0 + r][cmap(c)] ; 1: value(r, c) = W[c][cmap(r0 + r)]
  int r0, nr, nc;     // row offset / valid rows / valid cols (in image coordinates)
  int rot_f, rot_s;   // input-column rotation: image input index i -> W column (i < rot_f ? i + rot_s : i - rot_f)
  int lo;             // 1: store W - fp16(W) (the low half of the split operand) instead of fp16(W)
};

static int make_wimages(const svs_mlp_desc* d, const Layout& lo, WImages* wi, PackImg* list, int* n_list, bool split = false) {
  memset(wi, 0, sizeof(*wi));
  int64_t off = 0;
  int n = 0;
  auto add = [&](int n_pad, int kb, int layer, int transpose, int r0, int nr, int nc, int rot_f, int rot_s) {
    int64_t o = off;
    if (list) list[n] = PackImg{(long long)o, n_pad, kb, layer, transpose, r0, nr, nc, rot_f, rot_s, 0};
    ++n;
    off += round_up((int64_t)n_pad * kb * 128, 1024);
    return o;
  };
  const int L = lo.L;
  if (d->kind == SVS_NET_SDF) {
    for (int l = 0; l < L; ++l) {
      const bool last = (l == L - 1);
      if (!last) {
        wi->fwd_npad[l] = r16(lo.out[l]);
        wi->fwd_kb[l] = kb_of(lo.in[l]);
        wi->fwd[l] = add(wi->fwd_npad[l], wi->fwd_kb[l], l, 0, 0, lo.out[l], lo.in[l], 0, 0);
      } else {
        SVS_CHECK_ARG(lo.out[l] - 1 <= 256 && lo.out[l] >= 2, "SDF net: d_out + feature size %d unsupported by the tcgen05 engine", lo.out[l]);
        wi->fwd_npad[l] = r16(lo.out[l] - 1);
        wi->fwd_kb[l] = kb_of(lo.in[l]);
        wi->fwd[l] = add(wi->fwd_npad[l], wi->fwd_kb[l], l, 0, 1, lo.out[l] - 1, lo.in[l], 0, 0);
        wi->fwd_sdf = add(16, wi->fwd_kb[l], l, 0, 0, 1, lo.in[l], 0, 0);
      }
      wi->bwd_npad[l] = r16(lo.in[l]);
      wi->bwd_kb[l] = kb_of(lo.out[l]);
      wi->bwd[l] = add(wi->bwd_npad[l], wi->bwd_kb[l], l, 1, 0, lo.in[l], lo.out[l], 0, 0);
      SVS_CHECK_ARG(wi->fwd_npad[l] <= 256 && wi->bwd_npad[l] <= 256 && wi->fwd_kb[l] <= 4 && wi->bwd_kb[l] <= kMaxKB,
                    "SDF layer %d (%d -> %d) too wide for the tcgen05 engine", l, lo.in[l], lo.out[l]);
    }
  } else {
    const int pe_v = 3 * (1 + 2 * d->n_freqs);
    const int n_small = (d->render_mode == SVS_RENDER_IDR) ? 6 + pe_v : pe_v;
    const int F = lo.in[0] - n_small;
    SVS_CHECK_ARG(F > 0 && F % 64 == 0 && F <= 256 && n_small <= 32, "rendering net input %d unsupported by the tcgen05 engine", lo.in[0]);
    wi->F = F;
    wi->n_small = n_small;
    for (int l = 0; l < L; ++l) {
      wi->fwd_npad[l] = r16(lo.out[l]);
      wi->fwd_kb[l] = (l == 0) ? F / 64 + 1 : kb_of(lo.in[l]);
      wi->fwd[l] = add(wi->fwd_npad[l], wi->fwd_kb[l], l, 0, 0, lo.out[l], (l == 0) ? F + n_small : lo.in[l],
                       (l == 0) ? F : 0, (l == 0) ? n_small : 0);
      wi->bwd_kb[l] = kb_of(lo.out[l]);
      if (l == 0) {
        wi->bwd_npad[l] = F;
        wi->bwd[l] = add(F, wi->bwd_kb[l], l, 1, 0, F, lo.out[l], F, n_small);
        wi->small_npad = r16(n_small);
        wi->bwd_small = add(wi->small_npad, wi->bwd_kb[l], l, 1, F, n_small, lo.out[l], F, n_small);
      } else {
        wi->bwd_npad[l] = r16(lo.in[l]);
        wi->bwd[l] = add(wi->bwd_npad[l], wi->bwd_kb[l], l, 1, 0, lo.in[l], lo.out[l], 0, 0);
      }
      SVS_CHECK_ARG(wi->fwd_npad[l] <= 256 && wi->bwd_npad[l] <= 256 && wi->fwd_kb[l] <= kMaxKB && wi->bwd_kb[l] <= 4,
                    "rendering layer %d (%d -> %d) too wide for the tcgen05 engine", l, lo.in[l], lo.out[l]);
    }
  }
  // low halves of the forward images (split engine only), appended so that the offsets above do not depend on `split`
  {
    auto add_lo = [&](int n_pad, int kb, int layer, int r0, int nr, int nc, int rot_f, int rot_s) {
      int64_t o = off;
      if (split) {
        if (list) list[n] = PackImg{(long long)o, n_pad, kb, layer, 0, r0, nr, nc, rot_f, rot_s, 1};
        ++n;
        off += round_up((int64_t)n_pad * kb * 128, 1024);
      }
      return o;
    };
    if (d->kind == SVS_NET_SDF) {
      for (int l = 0; l < L; ++l) {
        if (l < L - 1) {
          wi->fwd_lo[l] = add_lo(wi->fwd_npad[l], wi->fwd_kb[l], l, 0, lo.out[l], lo.in[l], 0, 0);
        } else {
          wi->fwd_lo[l] = add_lo(wi->fwd_npad[l], wi->fwd_kb[l], l, 1, lo.out[l] - 1, lo.in[l], 0, 0);
          wi->fwd_sdf_lo = add_lo(16, wi->fwd_kb[l], l, 0, 1, lo.in[l], 0, 0);
        }
      }
    } else {
      for (int l = 0; l < L; ++l)
        wi->fwd_lo[l] = add_lo(wi->fwd_npad[l], wi->fwd_kb[l], l, 0, lo.out[l], (l == 0) ? wi->F + wi->n_small : lo.in[l],
                               (l == 0) ? wi->F : 0, (l == 0) ? wi->n_small : 0);
    }
    if (d->kind == SVS_NET_SDF)
      for (int l = 0; l < L - 1; ++l) {
        wi->fwd_bt[l] = off;
        if (split) off += round_up((int64_t)wi->fwd_npad[l] * 4, 1024);
      }
  }
  wi->bytes = off;
  if (n_list) *n_list = n;
  return SVS_OK;
}

struct PackArgsTc {
  PackImg img[4 * SVS_MAX_LAYERS + 2];
  int woff_ld[SVS_MAX_LAYERS];
  long long woff[SVS_MAX_LAYERS];
  int in[SVS_MAX_LAYERS], out[SVS_MAX_LAYERS];
};

// fp32 effective weights (wbuf) -> fp16 SWIZZLE_128B images
__global__ void pack_images_kernel(const PackArgsTc a, const float* __restrict__ wbuf, uint8_t* __restrict__ region) {
  const PackImg im = a.img[blockIdx.y];
  const int cols = im.kb * 64;
  const int total = im.n_pad * cols;
  const int l = im.layer;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx / cols, c = idx - r * cols;
    float v = 0.f;
    if (r < im.nr && c < im.nc) {
      int o, i;  // (out, in) index into W_eff
      if (!im.transpose) { o = im.r0 + r; i = c; } else { o = c; i = im.r0 + r; }
      if (im.rot_f > 0) i = (i < im.rot_f) ? i + im.rot_s : i - im.rot_f;
      if (o < a.out[l] && i < a.in[l]) v = wbuf[a.woff[l] + (long long)o * a.woff_ld[l] + i];
    }
    __half hv = __float2half_rn(v);
    if (im.lo) hv = __float2half_rn(v - __half2float(hv));
    *reinterpret_cast<__half*>(region + im.dst + img_off(r, c, im.n_pad)) = hv;
  }
}

// split engine: softplus biases pre-multiplied by 100 log2(e) (the epilogue computes t = acc * c + bias_t in one FFMA)
struct BiasTArgs {
  long long src[SVS_MAX_LAYERS], dst[SVS_MAX_LAYERS];
  int n[SVS_MAX_LAYERS], n_pad[SVS_MAX_LAYERS], count;
};
__global__ void pack_bias_t_kernel(const BiasTArgs a, const float* __restrict__ wbuf, uint8_t* __restrict__ region) {
  const int l = blockIdx.x;
  float* dst = reinterpret_cast<float*>(region + a.dst[l]);
  for (int i = threadIdx.x; i < a.n_pad[l]; i += blockDim.x) dst[i] = i < a.n[l] ? wbuf[a.src[l] + i] * kSpK1 : 0.f;
}

static int64_t wbuf_floats_tc(const svs_mlp_desc* d, const Layout& lo, bool split) {
  WImages wi;
  if (make_wimages(d, lo, &wi, nullptr, nullptr, split) != SVS_OK) return -1;
  return round_up(lo.total, 256) + wi.bytes / 4;
}
static inline uint8_t* wimg_region(const Layout& lo, const float* wbuf) {
  return reinterpret_cast<uint8_t*>(const_cast<float*>(wbuf) + round_up(lo.total, 256));
}

static int pack_images(const svs_mlp_desc* d, const Layout& lo, float* wbuf, cudaStream_t st, bool split) {
  WImages wi;
  PackArgsTc a;
  memset(&a, 0, sizeof(a));
  int n = 0;
  SVS_TRY(make_wimages(d, lo, &wi, a.img, &n, split));
  for (int l = 0; l < lo.L; ++l) {
    a.woff[l] = lo.woff[l];
    a.woff_ld[l] = lo.ldi[l];
    a.in[l] = lo.in[l];
    a.out[l] = lo.out[l];
  }
  pack_images_kernel<<<dim3(32, n), 256, 0, st>>>(a, wbuf, wimg_region(lo, wbuf));
  SVS_LAUNCH_OK();
  if (split && d->kind == SVS_NET_SDF && lo.L > 1) {
    BiasTArgs b;
    memset(&b, 0, sizeof(b));
    b.count = lo.L - 1;
    for (int l = 0; l < lo.L - 1; ++l) {
      b.src[l] = lo.boff[l];
      b.dst[l] = wi.fwd_bt[l];
      b.n[l] = lo.out[l];
      b.n_pad[l] = wi.fwd_npad[l];
    }
    pack_bias_t_kernel<<<b.count, 256, 0, st>>>(b, wbuf, wimg_region(lo, wbuf));
    SVS_LAUNCH_OK();
  }
  return SVS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// buffer layouts (bytes; every image = n_tiles x KB x 16 KiB)
// ---------------------------------------------------------------------------------------------------------------
struct ImgBuf {
  uint8_t* base;
  int kb;
};
static inline TcImg timg(const ImgBuf& b) { return TcImg{b.base, (int64_t)b.kb * kBlk}; }

struct SdfSaved {   // forward -> backward
  ImgBuf A0, H[SVS_MAX_LAYERS], U[SVS_MAX_LAYERS];  // H[1..L-1], U[0..L-2]
  int64_t bytes;
};
struct SdfBwdWs {
  ImgBuf Q[SVS_MAX_LAYERS], ZETA[SVS_MAX_LAYERS], DZ[SVS_MAX_LAYERS], DY;  // Q[0..L-2], ZETA[0..L-2], DZ[0..L-2]
  int64_t bytes;
};
struct RenderSaved {
  ImgBuf RIN, H[SVS_MAX_LAYERS];  // H[1..L-1]
  int64_t bytes;
};
struct RenderWs {
  ImgBuf DZ[SVS_MAX_LAYERS];  // DZ[0..L-1]
  int64_t bytes;
};

static inline int64_t n_tiles_of(int64_t P) { return cdiv(P, kTile); }

struct Carver {
  uint8_t* base;
  int64_t off, T;
  ImgBuf take(int kb) {
    ImgBuf b{base ? base + off : nullptr, kb};
    off += T * kb * (int64_t)kBlk;
    return b;
  }
};

static void map_sdf_saved(const Layout& lo, int64_t P, void* base, SdfSaved* s) {
  Carver c{(uint8_t*)base, 0, n_tiles_of(P)};
  s->A0 = c.take(kb_of(lo.in[0]));
  for (int l = 1; l < lo.L; ++l) s->H[l] = c.take(kb_of(lo.in[l]));
  for (int l = 0; l < lo.L - 1; ++l) s->U[l] = c.take(kb_of(lo.out[l]));
  s->bytes = c.off;
}
static void map_sdf_bwd(const Layout& lo, int64_t P, void* base, SdfBwdWs* s) {
  Carver c{(uint8_t*)base, 1024, n_tiles_of(P)};   // first 1024 bytes: upstream-gradient amax scalar
  for (int l = 0; l < lo.L - 1; ++l) s->Q[l] = c.take(kb_of(lo.in[l]));
  for (int l = 0; l < lo.L - 1; ++l) s->ZETA[l] = c.take(kb_of(lo.out[l]));
  for (int l = 0; l < lo.L - 1; ++l) s->DZ[l] = c.take(kb_of(lo.out[l]));
  s->DY = c.take(kb_of(lo.out[lo.L - 1]));
  s->bytes = c.off;
}
static void map_render_saved(const Layout& lo, const WImages& wi, int64_t P, void* base, RenderSaved* s) {
  Carver c{(uint8_t*)base, 0, n_tiles_of(P)};
  s->RIN = c.take(wi.F / 64 + 1);
  for (int l = 1; l < lo.L; ++l) s->H[l] = c.take(kb_of(lo.in[l]));
  s->bytes = c.off;
}
static void map_render_ws(const Layout& lo, int64_t P, void* base, RenderWs* s) {
  Carver c{(uint8_t*)base, 1024, n_tiles_of(P)};   // first 1024 bytes: upstream-gradient amax scalar
  for (int l = 0; l < lo.L; ++l) s->DZ[l] = c.take(kb_of(lo.out[l]));
  s->bytes = c.off;
}

// ---------------------------------------------------------------------------------------------------------------
// upstream-gradient magnitude (device scalar; no host synchronisation)
// ---------------------------------------------------------------------------------------------------------------
struct AmaxArgs {
  const float* p[3];
  long long n[3];
  float mult[3];
};
__global__ void amax_kernel(const AmaxArgs a, uint32_t* __restrict__ out) {
  float m = 0.f;
  for (int k = 0; k < 3; ++k) {
    if (!a.p[k]) continue;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    long long head = 0;
    if ((reinterpret_cast<uintptr_t>(a.p[k]) & 15) == 0) {   // 128-bit loads over the aligned bulk
      const float4* p4 = reinterpret_cast<const float4*>(a.p[k]);
      const long long n4 = a.n[k] >> 2;
      for (long long i = tid; i < n4; i += nthr) {
        const float4 q = __ldg(p4 + i);
        float v = fmaxf(fmaxf(fabsf(q.x), fabsf(q.y)), fmaxf(fabsf(q.z), fabsf(q.w))) * a.mult[k];
        if (v < 3.0e38f) m = fmaxf(m, v);   // ignores inf / nan
      }
      head = n4 << 2;
    }
    for (long long i = head + tid; i < a.n[k]; i += nthr) {
      float v = fabsf(a.p[k][i]) * a.mult[k];
      if (v < 3.0e38f) m = fmaxf(m, v);
    }
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}
static int launch_amax(const float* p0, int64_t n0, float m0, const float* p1, int64_t n1, float m1, const float* p2,
                       int64_t n2, float m2, uint32_t* out, cudaStream_t st) {
  SVS_CUDA_OK(cudaMemsetAsync(out, 0, 4, st));
  AmaxArgs a = {{p0, p1, p2}, {n0, n1, n2}, {m0, m1, m2}};
  amax_kernel<<<8 * kNumSMs, 256, 0, st>>>(a, out);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------------------------
static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
  }
  return n;
}

template <uint32_t EPI, int PRO, int NW>
static int launch_chain_t(const TcChain& ch, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SVS_CUDA_OK(cudaFuncSetAttribute(tc_chain_kernel<EPI, PRO, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  tc_chain_kernel<EPI, PRO, NW><<<grid, kCtrlThreads + NW * 32, kSmemBytes, st>>>(ch);
  return SVS_OK;
}

// epilogue warps per chain: 24 where the epilogue fits 72 registers without spilling (forward nets), else 16
constexpr int kNwFwd = SVS_TC_NW_FWD, kNwRev = SVS_TC_NW_REV, kNwRenderFwd = SVS_TC_NW_RFWD, kNwRenderBwd = SVS_TC_NW_RBWD,
              kNwTan = SVS_TC_NW_TAN, kNwBwd = SVS_TC_NW_BWD;

static double chain_bytes(const TcChain& ch);
static int launch_chain(TcChain& ch, const char* name, double flops, double bytes, cudaStream_t st) {
  ch.n_tiles = (int)n_tiles_of(ch.P);
  if (ch.n_tiles <= 0) return SVS_OK;
  const int grid = ch.n_tiles < num_sms() ? ch.n_tiles : num_sms();
  if (bytes <= 0.0) bytes = chain_bytes(ch);
  ProfScope ps(name, flops, bytes, st);
  if (ch.split) {
    if (!fwd3_supports(ch)) { set_error("launch_chain: chain not supported by the split-operand forward kernel"); return SVS_ERR_UNSUPPORTED; }
    SVS_TRY(launch_fwd3(ch, grid, num_sms(), st));
    SVS_LAUNCH_OK();
    return SVS_OK;
  }
  switch (ch.prologue) {
    case PRO_PE:
#ifndef SVS_TC_NO_FWD2
      if (fwd2_supports(ch)) { SVS_TRY(launch_fwd2(ch, grid, st)); break; }   // two tiles in flight (mlp_tc_fwd2.cuh)
#endif
      SVS_TRY((launch_chain_t<kEpiFwd, PRO_PE, kNwFwd>(ch, grid, st)));
      break;
    case PRO_LOAD_ULAST: SVS_TRY((launch_chain_t<kEpiRev, PRO_LOAD_ULAST, kNwRev>(ch, grid, st))); break;
    case PRO_RENDER_IN:
#ifndef SVS_TC_NO_FWD2
      if (fwd2_supports(ch)) { SVS_TRY(launch_fwd2(ch, grid, st)); break; }
#endif
      SVS_TRY((launch_chain_t<kEpiRenderFwd, PRO_RENDER_IN, kNwRenderFwd>(ch, grid, st)));
      break;
    case PRO_SIGMOID_BWD: SVS_TRY((launch_chain_t<kEpiRenderBwd, PRO_SIGMOID_BWD, kNwRenderBwd>(ch, grid, st))); break;
    case PRO_PE_JVP: SVS_TRY((launch_chain_t<kEpiTan, PRO_PE_JVP, kNwTan>(ch, grid, st))); break;
    case PRO_DY: SVS_TRY((launch_chain_t<kEpiBwd, PRO_DY, kNwBwd>(ch, grid, st))); break;
    default: set_error("launch_chain: unknown prologue %d", ch.prologue); return SVS_ERR_INVALID;
  }
  SVS_LAUNCH_OK();
  return SVS_OK;
}

static void init_chain(TcChain* ch) {
  memset(ch, 0, sizeof(*ch));
  ch->pro_save = ch->pro_img = ch->pro_colsum = -1;
}
static TcStep make_step(const uint8_t* w, const float* bias, int KB, int n_pad, int n_valid, int epi) {
  TcStep s;
  memset(&s, 0, sizeof(s));
  s.w = w; s.bias = bias; s.KB = KB; s.n_pad = n_pad; s.n_valid = n_valid; s.epi = epi;
  s.scale = 1.f; s.hscale = 1.f;
  s.aux1 = s.aux2 = s.save = s.out2 = s.colsum = -1;
  return s;
}

static double chain_flops(const TcChain& ch) {
  double f = 0;
  for (int s = 0; s < ch.n_steps; ++s) f += 2.0 * ch.st[s].KB * 64 * ch.st[s].n_valid;
  return f * (double)ch.P;
}

// bytes a chain has to move by construction: the fp16 tile images it loads / saves (activations kept for the backward
// and the weight gradients) and its fp32 row-major inputs / outputs; weights are L2-resident and not counted
static double chain_bytes(const TcChain& ch) {
  double per_tile = 0, per_point = 0;
  if (ch.pro_save >= 0) per_tile += (double)ch.pro_kb * kBlk;
  if (ch.prologue == PRO_LOAD_ULAST) per_tile += (double)ch.pro_kb * kBlk;
  for (int s = 0; s < ch.n_steps; ++s) {
    const TcStep& st = ch.st[s];
    const int nchunk = (st.n_pad + 63) >> 6;
    per_tile += (double)((st.aux1 >= 0) + (st.aux2 >= 0) + (st.epi == EP_TANGENT)) * nchunk * kBlk;
    if (st.save >= 0) per_tile += (double)st.next_kb * kBlk;
    if (st.epi == EP_Y || st.epi == EP_DFEAT) per_point += 4.0 * st.n_valid;
    if (st.epi == EP_SDF || st.epi == EP_PEGRAD) per_point += 4.0 * (1 + (ch.grad ? ch.d_in : 0));
    if (st.epi == EP_RGB) per_point += 4.0 * st.n_valid;
    if (st.epi == EP_DSMALL) per_point += 12.0;
  }
  if (ch.x) per_point += 4.0 * ch.d_in;
  if (ch.prologue == PRO_RENDER_IN) per_point += 4.0 * ch.F + 36.0;
  if (ch.prologue == PRO_SIGMOID_BWD) per_point += 8.0 * ch.n_rgb;
  if (ch.prologue == PRO_PE_JVP) per_point += 4.0 * ch.d_in;
  if (ch.prologue == PRO_DY) per_point += 4.0 * ch.dy_cols + 4.0;
  return per_tile * (double)n_tiles_of(ch.P) + per_point * (double)ch.P;
}

// split the CTAs of one weight-gradient launch over the jobs in proportion to their MMA work
static int launch_dw(DwParams& prm, int64_t P, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SVS_CUDA_OK(cudaFuncSetAttribute(tc_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmem));
    attr_set = true;
  }
  prm.n_tiles = (int)n_tiles_of(P);
  if (prm.n_tiles <= 0 || prm.n_jobs <= 0) return SVS_OK;
  double cost[kDwMaxJobs], total = 0;
  for (int j = 0; j < prm.n_jobs; ++j) {
    const DwJob& jb = prm.job[j];
    cost[j] = (double)jb.n_pairs * (jb.n_mblk * 2 + jb.n_yblk);   // bytes streamed per tile (the kernel is L2/HBM bound)
    total += cost[j];
  }
  const int budget = num_sms();   // one wave: every CTA streams its share of the tiles once and reduces once
  int unit = 0;
  double flops = 0, bytes = 0;
  for (int j = 0; j < prm.n_jobs; ++j) {
    DwJob& jb = prm.job[j];
    int n = (int)(budget * cost[j] / total);   // floor: the total never exceeds one wave
    if (n < 1) n = 1;
    if (n > prm.n_tiles) n = prm.n_tiles;
    jb.unit0 = unit;
    jb.n_split = n;
    unit += n;
    flops += 2.0 * (double)P * jb.n_pairs * jb.n_rows * jb.n_cols;
    // operand bytes the job has to stream: every tile's X and Y blocks, once per pair
    int nxb = 0;
    for (int b = 0; b < 2 * jb.n_mblk; ++b)
      if (jb.x_blk0 + b < jb.x_kb) ++nxb;
    bytes += (double)prm.n_tiles * jb.n_pairs * (nxb + jb.n_yblk) * kBlk;
  }
  // hand the CTAs the floor() left over to the jobs with the most work per CTA
  while (unit < budget) {
    int best = -1;
    double worst = 0;
    for (int j = 0; j < prm.n_jobs; ++j) {
      const DwJob& jb = prm.job[j];
      if (jb.n_split >= prm.n_tiles) continue;
      const double load = cost[j] * (double)cdiv(prm.n_tiles, jb.n_split);
      if (load > worst) { worst = load; best = j; }
    }
    if (best < 0) break;
    ++prm.job[best].n_split;
    ++unit;
  }
  unit = 0;
  for (int j = 0; j < prm.n_jobs; ++j) {
    prm.job[j].unit0 = unit;
    unit += prm.job[j].n_split;
  }
  ProfScope ps("mlp_tc_dw", flops, bytes, st);
  tc_dw_kernel<<<unit, 192, kDwSmem, st>>>(prm);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

static DwJob make_job(float* dW, int ldw, int n_rows, int n_cols) {
  DwJob j;
  memset(&j, 0, sizeof(j));
  j.dW = dW; j.ldw = ldw; j.n_rows = n_rows; j.n_cols = n_cols;
  return j;
}
static void job_pair(DwJob* j, const ImgBuf& X, const ImgBuf& Y) {
  int k = j->n_pairs++;
  j->X[k] = X.base; j->x_tile_bytes[k] = (int64_t)X.kb * kBlk; j->x_kb = X.kb;
  j->Y[k] = Y.base; j->y_tile_bytes[k] = (int64_t)Y.kb * kBlk; j->y_kb = Y.kb;
}
// db += column sums of the X image of the pair added last (valid columns: the job's n_rows)
static void job_bias(DwJob* j, float* bias) {
  j->bias = bias;
  j->bias_pair = j->n_pairs - 1;
}

}  // namespace tc
}  // namespace svs

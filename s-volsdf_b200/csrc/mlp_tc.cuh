// SVS_ENGINE_TC: the MLP chains of the VolSDF hot path on tcgen05 tensor cores (sm_100a).
//
// One persistent CTA per SM walks 128-point tiles.  For every tile a *chain* of steps runs with the activations
// resident in shared memory: step = [tcgen05.mma: acc(TMEM, fp32) = A(smem, fp16) x W^T(smem, fp16)] followed by
// an element-wise epilogue (TMEM -> registers -> fp16 -> the same smem tile, in place) that produces the A operand
// of the next step.  Layer weights are streamed from L2 as pre-packed SWIZZLE_128B images by 1-D bulk copies
// (UBLKCP) through a 2-slot ring; the activations a later pass needs (backward, weight gradients) are saved as
// tile images with bulk stores straight from the smem tile, and read back the same way ("aux" ring).
//
// Warp roles (640 threads): warp 0 weight producer, warp 1 MMA issuer (one elected lane), warp 2 aux producer +
// TMEM allocation, warp 3 store lane, warps 4..19 epilogue (TMEM lane quarter = warp % 4, 16-column pieces).
// MMA and epilogue overlap at 16-column granularity: as soon as the four epilogue warps that own a piece of the tile
// have rewritten it, the one MMA k-step of the NEXT step that consumes those 16 columns is issued into the other
// half of the 512 TMEM columns.  The forward chains (no aux tiles) run on tc_fwd2_kernel (mlp_tc_fwd2.cuh), which
// keeps two tiles in flight instead; this kernel serves the chains whose aux ring fills the shared memory.
// fp32 row-major tensors at the chain boundary (dy, d_feat) are moved coalesced through shared-memory transposes.
//
// Chains (math: SURVEY.md Appendix F; reference: volsdf/model/network.py:71-123,170-190 and its autograd graph):
//   sdf forward        PE -> 8 x softplus layer -> sdf (no-grad) | y = [sdf, features] (+ saved h_1..h_8)
//   sdf reverse sweep  U_7 = s(h_8) * W_8[0,:] -> U_{l-1} = s(h_l) * (U_l W_l) -> g = J_PE^T (U_0 W_0 + E), sphere clamp
//   render forward     [feat | x, PE(d), n] -> 4 x relu layer -> sigmoid
//   render backward    dz_4 = d_rgb rgb (1-rgb) -> dz_{l-1} = relu'(h_l) * (dz_l W_l) -> d_feat, d_normals
//   sdf tangent sweep  q_0 = J_PE (w d_grad) -> r_l = q_l W_l^T ; zeta_l = beta (1-s) U_l r_l ; q_{l+1} = s r_l
//   sdf backward       dz_8 = dy -> dz_{l-1} = s(h_l) * (dz_l W_l) + zeta_{l-1}
// Weight gradients dW_l = sum_p X[p,:]^T Y[p,:] are a separate kernel (tc_dw_kernel) that reads the saved images as
// MN-major operands and keeps the 256x256 fp32 accumulator in TMEM over all its tiles.
#pragma once
#include "svs_common.cuh"
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace svs {
namespace tc {

constexpr int kTile = 128;            // points per tile = MMA M
constexpr int kBlk = 16384;           // one 64-column block of a 128-row tile image
constexpr int kMaxKB = 5;             // widest A operand: 320 columns
constexpr int kWSlots = 2, kWSlot = 32768;
constexpr int kXSlots = 3, kXSlot = 16384;   // aux ring; chains with two aux tiles per block use a 4th slot (kXSlotsMax)
constexpr int kXSlotsMax = 4;                // ... that overlays the stash region (unused by those chains)
constexpr int kCtrlThreads = 128;     // 4 control warps; then NW epilogue warps (16 or 24, template parameter)
constexpr int kPieceWarps = 4;        // a 16-column piece of a block is written by 4 warps (one per TMEM lane quarter)
constexpr int kChunkWarps = 16;       // ... and a 64-column block by 16
constexpr int kStashLd = 40;          // fp32 per row: skip-branch / PE-gradient stash (<= 39 + pad)
constexpr int kMaxSteps = 11;
constexpr int kMaxImgs = 36;
constexpr int kMaxColsum = 9;
constexpr int kColsumW = 320;

// shared memory map (bytes)
constexpr int kOffA = 0;
constexpr int kOffW = kOffA + kMaxKB * kBlk;                // 81920
constexpr int kOffX = kOffW + kWSlots * kWSlot;             // 147456
constexpr int kOffStash = kOffX + kXSlots * kXSlot;         // 196608
constexpr int kOffVec = kOffStash + kTile * kStashLd * 4;   // 217088 : positional-encoding table (kPeTabCols entries of 8 bytes)
constexpr int kOffVec2 = kOffVec + 1024;                    // (spare)
constexpr int kOffColsum = kOffVec2 + 1024;                 // kMaxColsum x 256 fp32 per-CTA column sums
constexpr int kOffBar = kOffColsum + kMaxColsum * kColsumW * 4;  // 230656
constexpr int kSmemBytes = kOffBar + 512;                        // 231168 <= 232448 (227 KB)

enum TcEpi : int32_t {
  EP_SOFTPLUS = 0,  // A' = softplus(acc + b) * scale ; columns >= n_valid: PE(x) * scale (TC_PEFILL) or 0
  EP_SDF = 1,       // sdf[p] = min(acc_0 + b_0, sphere)                    (get_sdf_vals)
  EP_Y = 2,         // y[p, y_col + n] = acc + b
  EP_REVERSE = 3,   // p = acc*scale ; n < n_split: A' = s(h*hscale) p ; else stash[n - n_split] = p
  EP_PEGRAD = 4,    // stash += acc ; g = J_PE^T stash ; sphere clamp ; writes sdf, grad
  EP_RELU = 5,      // A' = relu(acc + b)
  EP_RGB = 6,       // rgb[p, n] = sigmoid(acc + b), n < n_valid
  EP_RELU_BWD = 7,  // A' = h > 0 ? acc : 0
  EP_DFEAT = 8,     // d_feat[p, n] = acc
  EP_DSMALL = 9,    // d_normals[p, j] = acc[small_off + j]
  EP_TANGENT = 10,  // s = s(h*hscale) ; zeta = beta (1-s) U acc -> image out2 ; A' = s acc scale
  EP_BACKWARD = 11  // A' = s(h*hscale) acc scale + zeta
};

enum TcPrologue : int32_t {
  PRO_PE = 0,          // A = PE(x)
  PRO_LOAD_ULAST = 1,  // A = h_8 image ; A' = s(h_8) * w_row
  PRO_RENDER_IN = 2,   // A = [feat(F) | points, PE(view), normals]
  PRO_SIGMOID_BWD = 3, // A = d_rgb rgb (1 - rgb)
  PRO_PE_JVP = 4,      // A = J_PE(x) (w d_grad)
  PRO_DY = 5           // A = dy (+ w d_sdf in column 0)
};

constexpr int32_t TC_PEFILL = 1;   // step flag: fill pad columns of A' with PE(x)*scale (skip connection)
constexpr int32_t TC_QFILL = 2;    // step flag: fill pad columns of A' with (J_PE(x) w d_grad)*scale (tangent of the skip)

struct TcStep {
  const uint8_t* w;    // KB blocks of n_pad x 128 bytes
  const uint8_t* w_lo; // split-operand chains (SVS_ENGINE_TC_SPLIT, mlp_tc_fwd3.cuh): the image of W - fp16(W), same geometry
  const float* bias;   // fp32[n_valid] or nullptr
  const float* bias_t; // split softplus steps: fp32[n_pad] = bias * 100 log2(e) (16-byte aligned)
  int32_t KB, n_pad, n_valid, epi;
  float scale, hscale;
  int32_t n_split;
  int32_t aux1, aux2;  // image ids or -1
  int32_t save;        // image id A' is saved to, or -1
  int32_t out2;        // EP_TANGENT: image id of zeta
  int32_t next_kb;     // 64-column blocks of A' to produce (0: the step has no A')
  int32_t flags;
  int32_t colsum;      // >= 0: add the column sums of A' into colsum slot
  int32_t y_col;       // EP_Y: first output column ; EP_DSMALL: first accumulator column
};

struct TcImg {
  uint8_t* base;
  int64_t tile_bytes;
};

struct TcChain {
  TcStep st[kMaxSteps];
  TcImg img[kMaxImgs];
  int32_t n_steps, prologue, pro_kb, pro_save;  // pro_save: image id the prologue's A is saved to, or -1
  int32_t pro_img;                              // PRO_LOAD_ULAST: image id loaded into A
  int32_t pro_colsum;                           // >= 0: column sums of the prologue's A
  const float* pro_vec;                         // PRO_LOAD_ULAST: fp32 row vector (W_last[0,:])
  int64_t P;
  int32_t n_tiles;
  int32_t split;                                // 1: forward chain with hi + lo operands (tc_fwd3_kernel)
  // geometry of the SDF net input
  const float* x;
  int32_t d_in, n_freqs;
  float radius, sph_scale;
  int64_t n_clamped;           // points [0, n_clamped) take the bounding-sphere minimum (get_outputs), the rest do not (gradient)
  // fp32 inputs / outputs (row-major)
  float* y; int32_t ldy;
  const float* yin;            // raw sdf column source for the clamp (reverse / tangent / backward chains)
  float* sdf; float* grad;
  const float* dy; const float* d_sdf; const float* d_grad; int32_t dy_cols;
  int32_t dy_bulk;             // PRO_DY: the fp32 rows of a tile are staged through the aux ring by bulk copies (dy 16-byte aligned)
  uint32_t dy_magic;           // ... ceil(2^32 / ldy): row of flat element e = umulhi(e, dy_magic) for e < 2^16
  // rendering net
  const float* points; const float* view; const float* normals; const float* feat; int32_t ld_feat;
  int32_t view_freqs, idr, F;
  float* rgb; const float* rgb_in; const float* d_rgb; int32_t n_rgb;
  float* d_normals; float* d_feat; int32_t ld_dfeat;
  float* colsum_out[kMaxColsum]; int32_t colsum_n[kMaxColsum];
  // backward chains: gradients are carried multiplied by grad_scale(amax, amax_target) (fp16 range)
  const uint32_t* amax; float amax_target;
};

#if defined(SVS_F3_TRACE) || defined(SVS_CHAIN_TRACE)
// clock64 trace of CTA 0, third tile (measurement builds only; read with tools/f3_trace.py).  Every tracing thread owns a
// 1024-entry lane of the buffer (plain stores, no atomics: an event costs a clock read and a store).
__device__ unsigned long long g_f3_trace[16384];
__device__ unsigned int g_f3_trace_n;
__device__ __forceinline__ void f3_ev(bool on, int lane_id, uint32_t& n, int tag, int a, int b) {
  if (!on || n >= 1024u) return;
  g_f3_trace[lane_id * 1024 + n++] = ((unsigned long long)clock64() & 0xFFFFFFFFFFull) | ((unsigned long long)tag << 56) |
                                     ((unsigned long long)(a & 255) << 48) | ((unsigned long long)(b & 255) << 40);
}
#endif
#ifdef SVS_F3_TRACE
#define F3_EV(on, tag, a, b) f3_ev(on, trace_lane, trace_n, tag, a, b)
#else
#define F3_EV(on, tag, a, b)
#endif
#ifndef SVS_CTRL_SLEEP
#define SVS_CTRL_SLEEP 0   // ns between barrier probes of the producer / store lanes (0: tight try_wait loop)
#endif
#ifndef SVS_CHAIN_EXP
#define SVS_CHAIN_EXP 0   // measurement builds only: 1 no weight copies, 2 no aux copies, 4 no tile saves, 8 no MMA
#endif
#ifdef SVS_CHAIN_TRACE
#define TC_EV(on, tag, a, b) f3_ev(on, trace_lane, trace_n, tag, a, b)
#else
#define TC_EV(on, tag, a, b)
#endif

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Operand storage format: IEEE fp16 (10-bit mantissa: the precision class of the TF32 GEMMs the reference ran with
// on Ampere; bf16's 7 bits are not enough behind Softplus(beta=100)).  Conversions saturate instead of producing inf.
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// barrier wait of the producer / store lanes: these single-lane warps share a scheduler with four epilogue warps each
__device__ __forceinline__ void mbar_wait_ctrl(uint64_t* bar, uint32_t parity) {
  if (SVS_CTRL_SLEEP > 0) {
    while (!mbar_try(bar, parity)) __nanosleep(SVS_CTRL_SLEEP);
  } else {
    mbar_wait(bar, parity);
  }
}

// power-of-two scale that brings the largest upstream gradient (|.| max as float bits in *amax) to ~target
__device__ __forceinline__ float grad_scale(const uint32_t* amax, float target) {
  if (!amax) return 1.f;
  float a = __uint_as_float(*amax);
  if (!(a > 0.f) || !isfinite(a)) return 1.f;
  float e = floorf(log2f(target / a));
  e = fminf(fmaxf(e, -40.f), 40.f);
  return exp2f(e);
}

// byte offset of the 16-byte chunk `chunk` (0..7) of row m inside a 64-column block
__device__ __forceinline__ uint32_t chunk_off(int m, int chunk) {
  return (uint32_t)(m >> 3) * 1024u + (uint32_t)(m & 7) * 128u + (uint32_t)((chunk ^ (m & 7)) << 4);
}
// 16 consecutive columns (quarter `cq` of a block) of row m: write / read
__device__ __forceinline__ void st_row16(uint8_t* blk, int m, int cq, const float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint4 u;
    u.x = pack_h2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_h2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_h2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_h2(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(blk + chunk_off(m, cq * 2 + i)) = u;
  }
}
__device__ __forceinline__ void ld_row16(const uint8_t* blk, int m, int cq, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint4 u = *reinterpret_cast<const uint4*>(blk + chunk_off(m, cq * 2 + i));
    float2 f;
    f = unpack_h2(u.x); v[8 * i + 0] = f.x; v[8 * i + 1] = f.y;
    f = unpack_h2(u.y); v[8 * i + 2] = f.x; v[8 * i + 3] = f.y;
    f = unpack_h2(u.z); v[8 * i + 4] = f.x; v[8 * i + 5] = f.y;
    f = unpack_h2(u.w); v[8 * i + 6] = f.x; v[8 * i + 7] = f.y;
  }
}

// the same 16 columns as packed fp16 pairs (8 registers); element i = half (i & 1) of h[i >> 1]
__device__ __forceinline__ void ld_row16h(const uint8_t* blk, int m, int cq, uint32_t (&h)[8]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint4 u = *reinterpret_cast<const uint4*>(blk + chunk_off(m, cq * 2 + i));
    h[4 * i + 0] = u.x; h[4 * i + 1] = u.y; h[4 * i + 2] = u.z; h[4 * i + 3] = u.w;
  }
}
__device__ __forceinline__ float h_elem(const uint32_t (&h)[8], int i) {
  const __half2 v = *reinterpret_cast<const __half2*>(&h[i >> 1]);
  return (i & 1) ? __high2float(v) : __low2float(v);
}

constexpr float kSpK1 = 144.26950408889634f;     // 100 * log2(e)
constexpr float kSpK2 = 0.006931471805599453f;   // ln(2) / 100
constexpr float kSpThr = 28.853900817779268f;    // 20 * log2(e): Softplus threshold (network.py:69, torch default 20)
// Softplus(beta=100, threshold 20) * scale with one ex2 + one lg2 (abs. error of the lg2(1+e) form <= 1e-9)
// c_lo = ln2/100 * scale, c_hi = scale / (100 log2 e).  softplus(z) >= z, and the capped branch stays at softplus(0.2),
// so max() selects z exactly where torch's threshold does.
__device__ __forceinline__ float softplus100_fast(float z, float c_lo, float c_hi) {
  float t = z * kSpK1;
  float l = lg2_approx(1.0f + ex2_approx(fminf(t, kSpThr)));
  return fmaxf(l * c_lo, t * c_hi);
}
// sigma'(z) recovered from h = softplus(z): 1 - exp(-100 h)
__device__ __forceinline__ float dsoftplus_h(float h) { return 1.0f - ex2_approx(-kSpK1 * h); }

// Positional-encoding columns come from a per-CTA table (one entry per column, built once): the column -> (frequency,
// sin / cos, input dimension) mapping needs two integer divisions by run-time values, ~600 cycles per element when done
// inline (tools/chain_trace.py: 20 k cycles of the reverse sweep's gradient step, 10 k of the tangent's prologue).
struct PeEntry {
  float mult;
  int32_t code;   // dim | kind << 4 ; kind: 0 zero, 1 identity, 2 sin, 3 cos
};
__device__ __forceinline__ PeEntry pe_entry(int c, int d_in, int n_freqs) {
  PeEntry e{0.f, 0};
  if (c < d_in) {
    e.mult = 1.f;
    e.code = c | (1 << 4);
  } else if (c < d_in * (1 + 2 * n_freqs)) {
    const int t = c - d_in, k = t / (2 * d_in), rem = t - k * 2 * d_in, fn = rem / d_in, dim = rem - fn * d_in;
    e.mult = (float)(1 << k);
    e.code = dim | ((fn ? 3 : 2) << 4);
  }
  return e;
}
__device__ __forceinline__ float pe_eval(const float (&xv)[4], PeEntry e) {
  const int dim = e.code & 15, kind = e.code >> 4;
  const float xs = dim == 0 ? xv[0] : (dim == 1 ? xv[1] : (dim == 2 ? xv[2] : xv[3]));
  // MUFU sin/cos: |arg| <= 2^(n_freqs-1) * |x| ~ 1e2 -> abs. error ~1e-5, far below the fp16 operand rounding (5e-4)
  const float arg = xs * e.mult;
  return kind == 1 ? xs : (kind == 2 ? __sinf(arg) : (kind == 3 ? __cosf(arg) : 0.f));
}
// d PE_c / d x_dim(c)
__device__ __forceinline__ float pe_deval(const float (&xv)[4], PeEntry e, int* dim_out) {
  const int dim = e.code & 15, kind = e.code >> 4;
  const float xs = dim == 0 ? xv[0] : (dim == 1 ? xv[1] : (dim == 2 ? xv[2] : xv[3]));
  const float arg = xs * e.mult;
  *dim_out = dim;
  return kind == 1 ? 1.f : (kind == 2 ? e.mult * __cosf(arg) : (kind == 3 ? -e.mult * __sinf(arg) : 0.f));
}
constexpr int kPeTabCols = 128;   // widest positional encoding the chain kernels index (table lives at kOffVec)

__device__ __forceinline__ float clamp_w(float y0, float sphere) { return (y0 < sphere) ? 1.f : ((y0 == sphere) ? 0.5f : 0.f); }

// per-column sums over the 32 lanes of v[0..15]; lanes i and i+16 both return column i's sum
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int w = 8; w >= 1; w >>= 1) {
    const bool upper = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      float send = upper ? v[i] : v[i + w];
      float keep = upper ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

struct Bars {   // must fit the 512 bytes reserved at kOffBar
  uint64_t w_full[kWSlots], w_empty[kWSlots], x_full[kXSlotsMax], x_empty[kXSlotsMax], a_ready[kMaxKB * 4], s_free[kMaxKB],
      acc_full, a_load, z_ready[kXSlotsMax];   // z_ready[slot]: zeta written into aux slot `slot` by all epilogue warps
  uint32_t tmem;
};

__host__ __device__ constexpr uint32_t epi_bit(int e) { return 1u << e; }

// bytes of tile t's dy rows that travel by bulk copy: whole 16-byte units of the valid rows (the last tile's <= 3 floats
// behind them are read directly)
__device__ __forceinline__ uint32_t dy_bulk_bytes(const TcChain& ch, int t) {
  const int64_t rows = ch.P - (int64_t)t * kTile;
  const uint32_t nr = rows < kTile ? (uint32_t)rows : (uint32_t)kTile;
  return (nr * (uint32_t)ch.ldy * 4u) & ~15u;
}

// ---------------------------------------------------------------------------------------------------------------
// the chain kernel; EPI / PRO select which epilogues / prologue are compiled into an instantiation
// ---------------------------------------------------------------------------------------------------------------
template <uint32_t EPI, int PRO, int NW>
__global__ void __launch_bounds__(kCtrlThreads + NW * 32, 1) tc_chain_kernel(const __grid_constant__ TcChain ch) {
  constexpr int kThreads = kCtrlThreads + NW * 32, kEpiThreads = NW * 32, NCG = NW / 4;   // NCG column groups
#ifdef SVS_CHAIN_TRACE
  constexpr bool kTraceThis = PRO == SVS_CHAIN_TRACE;   // measurement builds: -DSVS_CHAIN_TRACE=<prologue id of the chain to trace>
#else
  constexpr bool kTraceThis = false;
#endif
  static_assert(NW % 4 == 0 && NCG >= 4, "every block needs 4 distinct column groups");
  static_assert(NW == 16 || (PRO != PRO_DY && (EPI & epi_bit(EP_DFEAT)) == 0), "row-ownership passes assume 16 epilogue warps x 8 rows");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem + kOffA;
  uint8_t* sW = smem + kOffW;
  uint8_t* sX = smem + kOffX;
  float* stash = reinterpret_cast<float*>(smem + kOffStash);
  float* colsum = reinterpret_cast<float*>(smem + kOffColsum);
  Bars* bars = reinterpret_cast<Bars*>(smem + kOffBar);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // two aux tiles per block (tangent / backward sweeps): 4 ring slots = two blocks of look-ahead over HBM latency
  constexpr int XS = (EPI & (epi_bit(EP_TANGENT) | epi_bit(EP_BACKWARD))) ? kXSlotsMax : kXSlots;
  static_assert(XS == kXSlots || (EPI & (epi_bit(EP_REVERSE) | epi_bit(EP_PEGRAD))) == 0, "4th aux slot overlays the stash");

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWSlots; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
    for (int i = 0; i < kXSlotsMax; ++i) { mbar_init(&bars->x_full[i], 1); mbar_init(&bars->x_empty[i], kChunkWarps); }
    for (int i = 0; i < kXSlotsMax; ++i) mbar_init(&bars->z_ready[i], kChunkWarps);
    for (int i = 0; i < kMaxKB * 4; ++i) mbar_init(&bars->a_ready[i], kPieceWarps);   // one per 16-column piece
    for (int i = 0; i < kMaxKB; ++i) mbar_init(&bars->s_free[i], 1);
    mbar_init(&bars->acc_full, 1);
    mbar_init(&bars->a_load, 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < kMaxColsum * kColsumW; i += kThreads) colsum[i] = 0.f;
  if (warp == 2) tmem_alloc(&bars->tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem;

  if (warp == 0) {
    // ===== weight producer =====
    if (lane == 0) {
      uint32_t seq = 0;
      for (int t = blockIdx.x; t < ch.n_tiles; t += gridDim.x) {
        for (int s = 0; s < ch.n_steps; ++s) {
          const TcStep& st = ch.st[s];
          const uint32_t bytes = (uint32_t)st.n_pad * 128u;
          for (int kb = 0; kb < st.KB; ++kb, ++seq) {
            const int slot = seq % kWSlots;
            const uint32_t use = seq / kWSlots;
            mbar_wait_ctrl(&bars->w_empty[slot], (use & 1) ^ 1);
            if (SVS_CHAIN_EXP & 1) { mbar_arrive(&bars->w_full[slot]); continue; }
            mbar_arrive_expect_tx(&bars->w_full[slot], bytes);
            bulk_g2s(sW + slot * kWSlot, st.w + (size_t)kb * bytes, bytes, &bars->w_full[slot]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: block kb of step s is issued as soon as the epilogue of step s-1 has rewritten it =====
    if (lane == 0) {
      uint32_t seq = 0, n_step = 0, a_par = 0;
#ifdef SVS_CHAIN_TRACE
      const int trace_lane = 0;
      uint32_t trace_n = 0;
#endif
      for (int t = blockIdx.x; t < ch.n_tiles; t += gridDim.x) {
        [[maybe_unused]] const bool tr = kTraceThis && blockIdx.x == 0 && t == 2 * (int)gridDim.x;
        for (int s = 0; s < ch.n_steps; ++s, ++n_step) {
          const TcStep& st = ch.st[s];
          const uint32_t idesc = make_idesc_f16(kTile, st.n_pad, 0, 0);
          const uint32_t acc = tmem + (n_step & 1) * 256;
          TC_EV(tr, 1, s, 0);
          for (int kb = 0; kb < st.KB; ++kb, ++seq) {
            const int slot = seq % kWSlots;
            const uint32_t use = seq / kWSlots;
            mbar_wait(&bars->w_full[slot], use & 1);
            TC_EV(tr, 2, s, kb);
            const uint32_t a0 = smem_u32(sA + kb * kBlk), b0 = smem_u32(sW + slot * kWSlot);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // k-step j multiplies columns [64 kb + 16 j, +16): exactly what the 4 epilogue warps with cq == j wrote
              const int bi = kb * 4 + j;
              mbar_wait(&bars->a_ready[bi], (a_par >> bi) & 1);
              TC_EV(tr, 3, s, bi);
              a_par ^= 1u << bi;
              tc_fence_after();
              if (!(SVS_CHAIN_EXP & 8))
                umma_f16(acc, make_smem_desc(a0 + j * 32, 0, 1024), make_smem_desc(b0 + j * 32, 0, 1024), idesc, (kb | j) != 0);
            }
            umma_commit(&bars->w_empty[slot]);
          }
          umma_commit(&bars->acc_full);
          TC_EV(tr, 6, s, 0);
        }
        // a last step that rewrites A publishes blocks nobody multiplies: consume their phases
        const int tail_kb = ch.st[ch.n_steps - 1].next_kb;
        for (int bi = 0; bi < tail_kb * 4; ++bi) {
          mbar_wait(&bars->a_ready[bi], (a_par >> bi) & 1);
          a_par ^= 1u << bi;
        }
      }
    }
  } else if (warp == 2) {
    // ===== aux producer: the aux blocks in the order the epilogue consumes them (tile, step, chunk, aux).  A backward
    //       chain's fp32 dy rows travel through the same ring ahead of step 0 (the 128 rows of a tile are one contiguous
    //       range of P x ldy: 16 KB pieces, converted to the fp16 operand tile by the epilogue warps).  Pulling blocks into
    //       L2 ahead of the ring (cp.async.bulk.prefetch.L2, 12 blocks) was measured 8 % SLOWER: the sweeps are bound by
    //       their epilogues (tools/chain_trace.py), not by the aux latency. =====
    if (lane == 0) {
      uint32_t seq = 0;
#ifdef SVS_CHAIN_TRACE
      const int trace_lane = 4;
      uint32_t trace_n = 0;
#endif
      auto put = [&](const uint8_t* src, uint32_t bytes) {
        const int slot = seq % XS;
        const uint32_t use = seq / XS;
        ++seq;
        mbar_wait_ctrl(&bars->x_empty[slot], (use & 1) ^ 1);
        if (SVS_CHAIN_EXP & 2) { mbar_arrive(&bars->x_full[slot]); return; }
        mbar_arrive_expect_tx(&bars->x_full[slot], bytes);
        bulk_g2s(sX + slot * kXSlot, src, bytes, &bars->x_full[slot]);
      };
      for (int t = blockIdx.x; t < ch.n_tiles; t += gridDim.x) {
        if (PRO == PRO_DY && ch.dy_bulk) {
          const uint32_t total = dy_bulk_bytes(ch, t);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(ch.dy) + (size_t)t * kTile * ch.ldy * 4;
          for (uint32_t off = 0; off < total; off += kBlk) put(src + off, total - off < (uint32_t)kBlk ? total - off : (uint32_t)kBlk);
        }
        [[maybe_unused]] const bool tr = kTraceThis && blockIdx.x == 0 && t == 2 * (int)gridDim.x;
        for (int s = 0; s < ch.n_steps; ++s) {
          const TcStep& st = ch.st[s];
          if (st.aux1 < 0) continue;
          const int nchunk = (st.n_pad + 63) >> 6;
          for (int c = 0; c < nchunk; ++c) {
            for (int a = 0; a < 2; ++a) {
              const int id = a ? st.aux2 : st.aux1;
              if (id < 0) continue;
              TC_EV(tr, 20, s, c * 2 + a);
              put(ch.img[id].base + (size_t)t * ch.img[id].tile_bytes + (size_t)c * kBlk, kBlk);
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===== store warp: saves every generation of the A tile (bulk store straight from smem) and hands the blocks
    //       back to the epilogue (s_free) once the async proxy has read them =====
    if (lane == 0) {
      uint32_t a_par = 0, z_par = 0, xs_seq = 0;
#ifdef SVS_CHAIN_TRACE
      const int trace_lane = 5;
      uint32_t trace_n = 0;
#endif
      auto consume = [&](int kb) {
        for (int bi = kb * 4; bi < kb * 4 + 4; ++bi) {
          mbar_wait_ctrl(&bars->a_ready[bi], (a_par >> bi) & 1);
          a_par ^= 1u << bi;
        }
      };
      for (int t = blockIdx.x; t < ch.n_tiles; t += gridDim.x) {
        for (int b = 0; b < ch.pro_kb; ++b) consume(b);
        if (ch.pro_save >= 0) {
          bulk_s2g(ch.img[ch.pro_save].base + (size_t)t * ch.img[ch.pro_save].tile_bytes, sA, (uint32_t)ch.pro_kb * kBlk);
          bulk_commit();
          bulk_wait_read<0>();
        }
        for (int b = 0; b < ch.pro_kb; ++b) mbar_arrive(&bars->s_free[b]);
        [[maybe_unused]] const bool tr = kTraceThis && blockIdx.x == 0 && t == 2 * (int)gridDim.x;
        for (int s = 0; s < ch.n_steps; ++s) {
          const TcStep& st = ch.st[s];
          if (st.next_kb > 0) {
            for (int c = 0; c < st.next_kb; ++c) {
              if ((EPI & epi_bit(EP_TANGENT)) && st.epi == EP_TANGENT && c < ((st.n_pad + 63) >> 6)) {
                // zeta block c was written over the U block in its aux slot by all epilogue warps
                const int slot2 = (xs_seq + 1) % XS;
                xs_seq += 2;
                mbar_wait_ctrl(&bars->z_ready[slot2], (z_par >> slot2) & 1);
                z_par ^= 1u << slot2;
                if (!(SVS_CHAIN_EXP & 4))
                  bulk_s2g(ch.img[st.out2].base + (size_t)t * ch.img[st.out2].tile_bytes + (size_t)c * kBlk, sX + slot2 * kXSlot, kBlk);
                bulk_commit();
                bulk_wait_read<0>();
                mbar_arrive_n(&bars->x_empty[slot2], kChunkWarps);
              }
              consume(c);
              TC_EV(tr, 30, s, c);
              if (st.save >= 0 && !(SVS_CHAIN_EXP & 4)) {
                bulk_s2g(ch.img[st.save].base + (size_t)t * ch.img[st.save].tile_bytes + (size_t)c * kBlk, sA + c * kBlk, kBlk);
                bulk_commit();
              }
            }
            if (st.save >= 0) bulk_wait_read<0>();
            TC_EV(tr, 31, s, 0);
            for (int c = 0; c < st.next_kb; ++c) mbar_arrive(&bars->s_free[c]);
          } else if (s + 1 < ch.n_steps) {
            for (int c = 0; c < ch.st[s + 1].KB; ++c) consume(c);
          }
        }
      }
      bulk_wait_all<0>();
    }
  } else if (warp >= 4) {
    // ===== epilogue warps =====
    // 16-column pieces p = 0, 1, ... of the step's columns; this warp owns the pieces p = cg (mod NCG)
    const int ew = warp - 4, q = ew & 3, cg = ew >> 2;
    const int m = q * 32 + lane;
    const int et = threadIdx.x - 128;
    const bool leader = (et == 0);
    const uint32_t tm_row = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t n_acc = 0, xseq = 0, n_load = 0, fgen = 0;   // fgen: parity of the write generation of each A block
#ifdef SVS_CHAIN_TRACE
    const int trace_lane = 1 + (ew == 0 ? 0 : (ew == 5 ? 1 : 2));
    uint32_t trace_n = 0;
#endif
    const float gs = grad_scale(ch.amax, ch.amax_target);
    const float inv_gs = 1.0f / gs;
    const int pe_w = ch.d_in * (1 + 2 * ch.n_freqs);
    // biases of all steps live in shared memory for the whole kernel (the stash region is free in the chains that
    // have biases: forward SDF / rendering nets)
    constexpr bool kHasBias = (EPI & (epi_bit(EP_SOFTPLUS) | epi_bit(EP_SDF) | epi_bit(EP_Y) | epi_bit(EP_RELU) | epi_bit(EP_RGB))) != 0;
    static_assert(!kHasBias || (EPI & (epi_bit(EP_REVERSE) | epi_bit(EP_PEGRAD))) == 0, "bias table shares the stash region");
    static_assert(kMaxSteps * 256 * 4 <= kTile * kStashLd * 4, "bias table must fit the stash region");
    float* btab = stash;
    PeEntry* petab = reinterpret_cast<PeEntry*>(smem + kOffVec);   // visible after the first tile's named barrier
    if (et < kPeTabCols) petab[et] = pe_entry(et, ch.d_in, ch.n_freqs);
    if (kHasBias) {
      for (int s = 0; s < ch.n_steps; ++s) {
        const float* b = ch.st[s].bias;
        const int nv = ch.st[s].n_valid;
        if (et < 256) btab[s * 256 + et] = (b && et < nv) ? b[et] : 0.f;
      }
      named_bar_sync(1, kEpiThreads);
    }

    if (PRO == PRO_DY && ch.dy_bulk) {
      // columns ldy .. 64 pro_kb of the dy tile are never written (the staged rows hold ldy floats): zero them once; the
      // steps only rewrite the first 4 blocks
      for (int i = et * 16; i < ch.pro_kb * kBlk; i += kEpiThreads * 16) *reinterpret_cast<uint4*>(sA + i) = make_uint4(0, 0, 0, 0);
      fence_proxy_async();
      named_bar_sync(1, kEpiThreads);
    }
    for (int t = blockIdx.x; t < ch.n_tiles; t += gridDim.x) {
      const int64_t p = (int64_t)t * kTile + m;
      const bool live = p < ch.P;
      [[maybe_unused]] const bool trp = kTraceThis && blockIdx.x == 0 && t == 2 * (int)gridDim.x && lane == 0 && (ew == 0 || ew == 5 || ew == 15);
      TC_EV(trp, 40, 0, ew);
      float xv[4] = {0.f, 0.f, 0.f, 0.f};
      if (ch.x && live)
        for (int d = 0; d < ch.d_in; ++d) xv[d] = ch.x[p * ch.d_in + d];
      // sphere clamp of the point (get_outputs / get_sdf_vals, network.py:108-112,128-130)
      float sphere = 0.f, cw = 1.f, nrm = 1.f;
      const bool clamp_on = ch.radius > 0.f && p < ch.n_clamped;
      if (clamp_on) {
        float n2 = 0.f;
        for (int d = 0; d < ch.d_in; ++d) n2 += xv[d] * xv[d];
        nrm = sqrtf(n2);
        sphere = ch.sph_scale * (ch.radius - nrm);
        if (ch.yin && live) cw = clamp_w(ch.yin[p * ch.ldy], sphere);
      }
      float dg[4] = {0.f, 0.f, 0.f, 0.f};  // grad_scale * w * dL/dgrad of the point (tangent sweep)
      if (PRO == PRO_PE_JVP && ch.d_grad && live)
        for (int d = 0; d < ch.d_in; ++d) dg[d] = gs * cw * ch.d_grad[p * ch.d_in + d];
      if (EPI & (epi_bit(EP_REVERSE) | epi_bit(EP_PEGRAD)))
        for (int i = cg; i < kStashLd; i += NCG) stash[m * kStashLd + i] = 0.f;
      if (PRO == PRO_LOAD_ULAST) {
        if (leader) {
          // the blocks must have been saved (previous tile) before the async proxy overwrites them
          for (int b = 0; b < ch.pro_kb; ++b) mbar_wait(&bars->s_free[b], ((fgen >> b) & 1) ^ 1);
          const uint32_t bytes = (uint32_t)ch.pro_kb * kBlk;
          mbar_arrive_expect_tx(&bars->a_load, bytes);
          bulk_g2s(sA, ch.img[ch.pro_img].base + (size_t)t * ch.img[ch.pro_img].tile_bytes, bytes, &bars->a_load);
        }
      }
      TC_EV(trp, 41, 0, ew);
      named_bar_sync(1, kEpiThreads);
      TC_EV(trp, 42, 0, ew);

      // ---------------- prologue: build the first A operand, one 64-column block at a time ----------------
      if (PRO == PRO_LOAD_ULAST) {
        mbar_wait(&bars->a_load, n_load & 1);
        ++n_load;
        TC_EV(trp, 43, 0, ew);
      }
      if (PRO == PRO_DY) {
        // dy is fp32 row-major (P x ldy); column 0 (which also takes w * dL/dsdf) is written by the thread that owns the
        // row.  The pieces are published after all warps are done.
        for (int b = 0; b < ch.pro_kb; ++b) mbar_wait(&bars->s_free[b], ((fgen >> b) & 1) ^ 1);
        TC_EV(trp, 43, 0, ew);
        const int64_t trow0 = (int64_t)t * kTile;
        const int ncol = ch.pro_kb * 64;
        float csum[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) csum[j] = 0.f;
        if (ch.dy_bulk) {
          // the tile's rows arrive as 16 KB pieces of the flat (row, column) array in the aux ring: thread et converts the
          // floats et, et + 512, ... of a piece (consecutive lanes: consecutive columns of a row, conflict-free both ways)
          const uint32_t total = dy_bulk_bytes(ch, t);
          const int64_t rows = ch.P - trow0;
          const uint32_t n_el = (uint32_t)(rows < kTile ? rows : kTile) * (uint32_t)ch.ldy;   // valid floats of the tile
          const float* gsrc = ch.dy + trow0 * ch.ldy;
          for (uint32_t off = 0; off < (uint32_t)(kTile * ch.ldy * 4); off += kBlk) {
            const uint32_t e0 = off >> 2;
            int slot = -1;
            if (off < total) {
              slot = xseq % XS;
              mbar_wait(&bars->x_full[slot], (xseq / XS) & 1);
              ++xseq;
            }
            const float* ssrc = reinterpret_cast<const float*>(sX + (slot < 0 ? 0 : slot) * kXSlot);
#pragma unroll
            for (int k = 0; k < kBlk / 4 / 512; ++k) {
              const uint32_t i = (uint32_t)(k * 512 + et), e = e0 + i;
              if (e >= (uint32_t)(kTile * ch.ldy)) break;
              float v = 0.f;
              if (e * 4u + 4u <= total) v = ssrc[i];
              else if (e < n_el) v = __ldg(gsrc + e);
              const uint32_t r = __umulhi(e, ch.dy_magic), c = e - r * (uint32_t)ch.ldy;
              if (c >= 1u && c < (uint32_t)ncol)
                *reinterpret_cast<__half*>(sA + (c >> 6) * kBlk + chunk_off((int)r, (int)((c & 63) >> 3)) + (c & 7) * 2) =
                    __float2half_rn(fminf(fmaxf(c < (uint32_t)ch.dy_cols ? v * gs : 0.f, -65504.f), 65504.f));
            }
            if (slot >= 0) {
              __syncwarp();
              if (lane == 0) mbar_arrive(&bars->x_empty[slot]);
            }
          }
        }
        for (int rb = 0; !ch.dy_bulk && rb < 8; rb += 2) {
          float fv[2][10];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int64_t pr = trow0 + ew * 8 + rb + h;
#pragma unroll
            for (int j = 0; j < 10; ++j) {
              const int c = lane + 32 * j;
              fv[h][j] = (ch.dy && pr < ch.P && c >= 1 && c < ch.dy_cols) ? __ldg(ch.dy + pr * ch.ldy + c) * gs : 0.f;
            }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int row = ew * 8 + rb + h;
#pragma unroll
            for (int j = 0; j < 10; ++j) {
              const int c = lane + 32 * j;
              if (c >= 1 && c < ncol) {
                csum[j] += fv[h][j];
                *reinterpret_cast<__half*>(sA + (c >> 6) * kBlk + chunk_off(row, (c & 63) >> 3) + (c & 7) * 2) =
                    __float2half_rn(fminf(fmaxf(fv[h][j], -65504.f), 65504.f));
              }
            }
          }
        }
        TC_EV(trp, 44, 0, ew);
        float v0 = 0.f;
        if (cg == 0) {   // one thread per row
          if (live) {
            if (ch.dy) v0 = ch.dy[p * ch.ldy];
            if (ch.d_sdf) v0 += cw * ch.d_sdf[p];
            v0 *= gs;
          }
          *reinterpret_cast<__half*>(sA + chunk_off(m, 0)) = __float2half_rn(fminf(fmaxf(v0, -65504.f), 65504.f));
        }
        if (ch.pro_colsum >= 0) {
#pragma unroll
          for (int j = 0; j < 10; ++j) {
            const int c = lane + 32 * j;
            if (c >= 1 && c < ch.dy_cols) atomicAdd(&colsum[ch.pro_colsum * kColsumW + c], csum[j]);
          }
          if (cg == 0) {
            const float s0 = warp_sum(v0);
            if (lane == 0) atomicAdd(&colsum[ch.pro_colsum * kColsumW], s0);
          }
        }
        TC_EV(trp, 45, 0, ew);
        fence_proxy_async();
        named_bar_sync(1, kEpiThreads);
        if (lane == 0)
          for (int pc = cg; pc < ch.pro_kb * 4; pc += NCG) mbar_arrive(&bars->a_ready[pc]);
        TC_EV(trp, 46, 0, ew);
      }
      for (int pc = cg; PRO != PRO_DY && pc < ch.pro_kb * 4; pc += NCG) {
        float v[16];
        const int b = pc >> 2, cq = pc & 3;
        const int c0 = pc * 16;
        if (PRO == PRO_PE) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (c0 + i < pe_w) ? pe_eval(xv, petab[c0 + i]) : 0.f;
        } else if (PRO == PRO_LOAD_ULAST) {
          ld_row16(sA + b * kBlk, m, cq, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = dsoftplus_h(v[i]) * __ldg(ch.pro_vec + c0 + i);
        } else if (PRO == PRO_RENDER_IN) {
          const int nfb = ch.F >> 6;  // feature blocks, then one block [points(3) if idr, PE(view), normals(3) if idr]
          if (b < nfb) {
            const float4* src = reinterpret_cast<const float4*>(ch.feat + p * ch.ld_feat + c0);
            const bool al = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
              if (live) {
                if (al) f = src[i];
                else { const float* sp = ch.feat + p * ch.ld_feat + c0 + 4 * i; f = make_float4(sp[0], sp[1], sp[2], sp[3]); }
              }
              v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
            }
          } else {
            const int pe_v = 3 * (1 + 2 * ch.view_freqs);
            const int o_view = ch.idr ? 3 : 0, o_n = o_view + pe_v, n_small = o_n + (ch.idr ? 3 : 0);
            float vv[4] = {0.f, 0.f, 0.f, 0.f};
            if (live) { vv[0] = ch.view[p * 3]; vv[1] = ch.view[p * 3 + 1]; vv[2] = ch.view[p * 3 + 2]; }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = cq * 16 + i;
              float r = 0.f;
              if (live && c < n_small) {
                if (c < o_view) r = ch.points[p * 3 + c];
                else if (c < o_n) r = pe_eval(vv, pe_entry(c - o_view, 3, ch.view_freqs));
                else r = ch.normals[p * 3 + (c - o_n)];
              }
              v[i] = r;
            }
          }
        } else if (PRO == PRO_SIGMOID_BWD) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + i;
            float r = 0.f;
            if (live && c < ch.n_rgb) {
              float sg = ch.rgb_in[p * ch.n_rgb + c];
              r = gs * ch.d_rgb[p * ch.n_rgb + c] * sg * (1.f - sg);
            }
            v[i] = r;
          }
        } else if (PRO == PRO_PE_JVP) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + i;
            float r = 0.f;
            if (live && c < pe_w) {
              int dim;
              float j = pe_deval(xv, petab[c], &dim);
              r = j * dg[dim];
            }
            v[i] = r;
          }
        } else if (PRO == PRO_DY) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + i;
            float r = 0.f;
            if (live && c < ch.dy_cols) {
              if (ch.dy) r = ch.dy[p * ch.ldy + c];
              if (c == 0 && ch.d_sdf) r += cw * ch.d_sdf[p];
              r *= gs;
            }
            v[i] = r;
          }
        }
        mbar_wait(&bars->s_free[b], ((fgen >> b) & 1) ^ 1);
        st_row16(sA + b * kBlk, m, cq, v);
        if (ch.pro_colsum >= 0) {
          float cs = warp_colsum16(v, lane);
          if (lane < 16) atomicAdd(&colsum[ch.pro_colsum * kColsumW + c0 + lane], cs);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->a_ready[pc]);
        TC_EV(trp, 46, 0, pc);
      }
      for (int b = 0; b < ch.pro_kb; ++b) fgen ^= 1u << b;

      // ---------------- steps ----------------
      for (int s = 0; s < ch.n_steps; ++s) {
        const TcStep st = ch.st[s];   // by value: keeps the hot fields in registers instead of indexed constant loads
        const float* btab_s = btab + s * 256;
        const bool has_next = s + 1 < ch.n_steps;
        const bool writes_a = st.next_kb > 0;
        const float c_lo = kSpK2 * st.scale, c_hi = st.scale / kSpK1;
        const float khs = -kSpK1 * st.hscale;   // s(h * hscale) = 1 - 2^(khs * h)
        const uint32_t tm_acc = tm_row + (n_acc & 1) * 256;
        mbar_wait(&bars->acc_full, n_acc & 1);
        ++n_acc;
        tc_fence_after();
        [[maybe_unused]] const bool tr = kTraceThis && blockIdx.x == 0 && (t == 2 * (int)gridDim.x || t == (int)gridDim.x) && lane == 0 && (ew == 0 || ew == 5 || ew == 15);
        TC_EV(tr, 7, s, ew);
        if (!writes_a && has_next) {
          // A is not rewritten by this step: the next step's MMAs may start at once (into the other accumulator)
          if (lane == 0)
            for (int pc = cg; pc < ch.st[s + 1].KB * 4; pc += NCG) mbar_arrive(&bars->a_ready[pc]);
        }

        const int nchunk_acc = (st.n_pad + 63) >> 6;
        const int nchunk = max(nchunk_acc, st.next_kb);
        const int naux = (st.aux1 >= 0 ? 1 : 0) + (st.aux2 >= 0 ? 1 : 0);
        const uint32_t xbase = xseq;   // aux blocks are numbered (chunk, aux) in the order the producer loads them
        uint32_t rr[16];   // accumulator columns of this warp's NEXT piece (tcgen05.ld issued one piece ahead)
        if (cg * 16 < st.n_pad) tmem_ld_32x16(tm_acc + (uint32_t)(cg * 16), rr);
        for (int pc = cg; pc < nchunk * 4; pc += NCG) {
          const int c = pc >> 2, cq = pc & 3;
          const int col0 = pc * 16;
          float acc[16];
          TC_EV(tr, 8, s, pc);
          if (col0 < st.n_pad) {
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(rr[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
          }
          TC_EV(tr, 9, s, pc);
          if (col0 + NCG * 16 < st.n_pad) tmem_ld_32x16(tm_acc + (uint32_t)(col0 + NCG * 16), rr);
          // aux tiles of this chunk
          uint32_t a1[8], a2[8];   // aux values stay packed (fp16 pairs) until they are used: 16 registers less
          int slot1 = -1, slot2 = -1;
          if (EPI & (epi_bit(EP_REVERSE) | epi_bit(EP_RELU_BWD) | epi_bit(EP_TANGENT) | epi_bit(EP_BACKWARD))) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a1[i] = a2[i] = 0u;
            if (c < nchunk_acc && st.aux1 >= 0) {
              const uint32_t x1 = xbase + (uint32_t)(c * naux);
              slot1 = x1 % XS;
              mbar_wait(&bars->x_full[slot1], (x1 / XS) & 1);
              ld_row16h(sX + slot1 * kXSlot, m, cq, a1);
              if (st.aux2 >= 0) {
                slot2 = (x1 + 1) % XS;
                mbar_wait(&bars->x_full[slot2], ((x1 + 1) / XS) & 1);
                ld_row16h(sX + slot2 * kXSlot, m, cq, a2);
              }
            }
          }
          TC_EV(tr, 13, s, pc);
          float o[16];
          const bool full = col0 + 16 <= st.n_valid;   // no pad columns in this thread's 16
          if ((EPI & epi_bit(EP_SOFTPLUS)) && st.epi == EP_SOFTPLUS) {
            if (full) {
              const float4* b4 = reinterpret_cast<const float4*>(btab_s + col0);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 b = b4[i];
                o[4 * i + 0] = softplus100_fast(acc[4 * i + 0] + b.x, c_lo, c_hi);
                o[4 * i + 1] = softplus100_fast(acc[4 * i + 1] + b.y, c_lo, c_hi);
                o[4 * i + 2] = softplus100_fast(acc[4 * i + 2] + b.z, c_lo, c_hi);
                o[4 * i + 3] = softplus100_fast(acc[4 * i + 3] + b.w, c_lo, c_hi);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int n = col0 + i;
                float r = 0.f;
                if (n < st.n_valid) r = softplus100_fast(acc[i] + btab_s[n & 255], c_lo, c_hi);
                else if (st.flags & TC_PEFILL) r = pe_eval(xv, petab[n - st.n_valid]) * st.scale;
                o[i] = r;
              }
            }
          } else if ((EPI & epi_bit(EP_SDF)) && st.epi == EP_SDF) {
            if (col0 == 0 && live) {
              float y0 = acc[0] + btab_s[0];
              if (clamp_on) y0 = fminf(y0, sphere);
              ch.sdf[p] = y0;
            }
          } else if ((EPI & epi_bit(EP_Y)) && st.epi == EP_Y) {
            if (live) {
              float* dst = ch.y + p * ch.ldy + st.y_col + col0;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (col0 + i < st.n_valid) dst[i] = acc[i] + btab_s[(col0 + i) & 255];
            }
          } else if ((EPI & epi_bit(EP_REVERSE)) && st.epi == EP_REVERSE) {
            if (col0 + 16 <= st.n_split) {   // whole piece inside the sigma' columns: no per-element predicates
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = (1.0f - ex2_approx(khs * h_elem(a1, i))) * (acc[i] * st.scale);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int n = col0 + i;
                const float pv = acc[i] * st.scale;
                float r = 0.f;
                if (n < st.n_split) r = (1.0f - ex2_approx(khs * h_elem(a1, i))) * pv;
                else if (n < st.n_valid) stash[m * kStashLd + (n - st.n_split)] = pv;
                o[i] = r;
              }
            }
          } else if ((EPI & epi_bit(EP_PEGRAD)) && st.epi == EP_PEGRAD) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (col0 + i < st.n_valid) stash[m * kStashLd + col0 + i] += acc[i];
          } else if ((EPI & epi_bit(EP_RELU)) && st.epi == EP_RELU) {
            if (full) {
              const float4* b4 = reinterpret_cast<const float4*>(btab_s + col0);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 b = b4[i];
                o[4 * i + 0] = fmaxf(acc[4 * i + 0] + b.x, 0.f);
                o[4 * i + 1] = fmaxf(acc[4 * i + 1] + b.y, 0.f);
                o[4 * i + 2] = fmaxf(acc[4 * i + 2] + b.z, 0.f);
                o[4 * i + 3] = fmaxf(acc[4 * i + 3] + b.w, 0.f);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int n = col0 + i;
                o[i] = (n < st.n_valid) ? fmaxf(acc[i] + btab_s[n & 255], 0.f) : 0.f;
              }
            }
          } else if ((EPI & epi_bit(EP_RGB)) && st.epi == EP_RGB) {
            if (live) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int n = col0 + i;
                if (n < st.n_valid) ch.rgb[p * st.n_valid + n] = 1.0f / (1.0f + __expf(-(acc[i] + btab_s[n & 255])));
              }
            }
          } else if ((EPI & epi_bit(EP_RELU_BWD)) && st.epi == EP_RELU_BWD) {
            if (full) {
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = (h_elem(a1, i) > 0.f) ? acc[i] : 0.f;
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = (h_elem(a1, i) > 0.f && col0 + i < st.n_valid) ? acc[i] : 0.f;
            }
          } else if ((EPI & epi_bit(EP_DFEAT)) && st.epi == EP_DFEAT) {
            if (ch.d_feat) {
              // fp32 row-major output from a thread-per-row register layout: transposed through this warp's scratch
              // (the stash region is idle in this chain) so that one store instruction writes 4 rows x 8 contiguous
              // floats instead of 32 rows x 1
              float* scr = stash + ew * 288;
              const int64_t row0 = (int64_t)t * kTile + q * 32;
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                for (int i = 0; i < 8; ++i) scr[lane * 9 + i] = acc[hf * 8 + i] * inv_gs;
                __syncwarp();
                const int cc = lane & 7, n = col0 + hf * 8 + cc;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const int rw = 4 * k + (lane >> 3);
                  const int64_t pr = row0 + rw;
                  if (pr < ch.P && n < st.n_valid) ch.d_feat[pr * ch.ld_dfeat + n] = scr[rw * 9 + cc];
                }
                __syncwarp();
              }
            }
          } else if ((EPI & epi_bit(EP_DSMALL)) && st.epi == EP_DSMALL) {
            if (live && ch.d_normals) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int j = col0 + i - st.y_col;
                if (j >= 0 && j < 3) ch.d_normals[p * 3 + j] = acc[i] * inv_gs;
              }
            }
          } else if ((EPI & epi_bit(EP_TANGENT)) && st.epi == EP_TANGENT) {
            float z[16];
            if (full) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float sg = 1.0f - ex2_approx(khs * h_elem(a1, i));
                z[i] = 100.0f * (1.0f - sg) * h_elem(a2, i) * acc[i];
                o[i] = sg * acc[i] * st.scale;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int n = col0 + i;
                const float sg = 1.0f - ex2_approx(khs * h_elem(a1, i));
                const bool ok = n < st.n_valid;
                z[i] = ok ? 100.0f * (1.0f - sg) * h_elem(a2, i) * acc[i] : 0.f;
                o[i] = ok ? sg * acc[i] * st.scale : 0.f;
              }
            }
            if (!full && (st.flags & TC_QFILL)) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int n = col0 + i;
                if (n >= st.n_valid && n - st.n_valid < pe_w) {
                  int dim;
                  float j = pe_deval(xv, petab[n - st.n_valid], &dim);
                  o[i] = live ? j * dg[dim] * st.scale : 0.f;
                }
              }
            }
            // zeta overwrites the U block in its aux slot and leaves through a bulk store
            if (slot2 >= 0) st_row16(sX + slot2 * kXSlot, m, cq, z);
          } else if ((EPI & epi_bit(EP_BACKWARD)) && st.epi == EP_BACKWARD) {
            if (full) {
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = (1.0f - ex2_approx(khs * h_elem(a1, i))) * acc[i] * st.scale + h_elem(a2, i);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float sg = 1.0f - ex2_approx(khs * h_elem(a1, i));
                o[i] = (col0 + i < st.n_valid) ? sg * acc[i] * st.scale + h_elem(a2, i) : 0.f;
              }
            }
          }
          if (writes_a && c < st.next_kb) {
            TC_EV(tr, 10, s, pc);
            mbar_wait(&bars->s_free[c], ((fgen >> c) & 1) ^ 1);   // the previous generation of this block has been saved
            TC_EV(tr, 14, s, pc);
            st_row16(sA + c * kBlk, m, cq, o);
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->a_ready[pc]);
            TC_EV(tr, 12, s, pc);
            if (st.colsum >= 0) {
              float cs = warp_colsum16(o, lane);
              if (lane < 16) atomicAdd(&colsum[st.colsum * kColsumW + col0 + lane], cs);
            }
          }
          // release the aux slots
          if (slot1 >= 0) {
            if ((EPI & epi_bit(EP_TANGENT)) && st.epi == EP_TANGENT) {
              // zeta sits in slot2: the store warp ships it (after all 16 warps arrived) and frees the slot
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                mbar_arrive(&bars->z_ready[slot2]);   // per slot: a single barrier would mix arrivals of different blocks
                mbar_arrive(&bars->x_empty[slot1]);
              }
            } else {
              __syncwarp();
              if (lane == 0) {
                mbar_arrive(&bars->x_empty[slot1]);
                if (slot2 >= 0) mbar_arrive(&bars->x_empty[slot2]);
              }
            }
          }
        }
        if ((EPI & epi_bit(EP_PEGRAD)) && st.epi == EP_PEGRAD) {
          // stash now holds p_0 + e for all PE columns of the row: g = J_PE^T (p_0 + e), sphere clamp of sdf and g
          TC_EV(tr, 47, s, ew);
          named_bar_sync(2, kEpiThreads);
          TC_EV(tr, 48, s, ew);
          if (cg == 0 && live) {   // the warps that own piece 0 (one thread per row)
            float g[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c = 0; c < pe_w; ++c) {
              int dim;
              float j = pe_deval(xv, petab[c], &dim);
              g[dim] += j * stash[m * kStashLd + c];
            }
            float y0 = ch.yin[p * ch.ldy];
            float out_sdf = y0;
            if (clamp_on) {
              out_sdf = fminf(y0, sphere);
              if (cw < 1.f)   // sphere branch only where taken; |x| = 0 has a zero gradient (norm backward masks it)
                for (int d = 0; d < ch.d_in; ++d) g[d] = cw * g[d] + (1.f - cw) * (nrm > 0.f ? -ch.sph_scale * xv[d] / nrm : 0.f);
            }
            if (ch.sdf) ch.sdf[p] = out_sdf;
            if (ch.grad)
              for (int d = 0; d < ch.d_in; ++d) ch.grad[p * ch.d_in + d] = g[d];
          }
          TC_EV(tr, 49, s, ew);
          named_bar_sync(2, kEpiThreads);   // the stash is re-zeroed by the next tile's prologue
          TC_EV(tr, 50, s, ew);
        }
        // every warp advances the ring / generation counters of ALL blocks of the step, also those it did not touch
        if (st.aux1 >= 0) xseq = xbase + (uint32_t)(nchunk_acc * naux);
        if (writes_a)
          for (int c = 0; c < st.next_kb; ++c) fgen ^= 1u << c;
        tc_fence_before();
      }
    }
    if (leader) bulk_wait_all<0>();   // zeta stores of the tangent chain
    // flush the per-CTA column sums (bias gradients, dW_last[0,:])
    named_bar_sync(1, kEpiThreads);
    for (int k = 0; k < kMaxColsum; ++k) {
      if (!ch.colsum_out[k]) continue;
      for (int i = et; i < ch.colsum_n[k]; i += kEpiThreads) atomicAdd(&ch.colsum_out[k][i], colsum[k * kColsumW + i] * inv_gs);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

static_assert(sizeof(Bars) <= 512, "barrier block overflows its shared-memory reservation");

constexpr uint32_t kEpiFwd = epi_bit(EP_SOFTPLUS) | epi_bit(EP_SDF) | epi_bit(EP_Y);
constexpr uint32_t kEpiRev = epi_bit(EP_REVERSE) | epi_bit(EP_PEGRAD);
constexpr uint32_t kEpiRenderFwd = epi_bit(EP_RELU) | epi_bit(EP_RGB);
constexpr uint32_t kEpiRenderBwd = epi_bit(EP_RELU_BWD) | epi_bit(EP_DFEAT) | epi_bit(EP_DSMALL);
static_assert(16 * 288 * 4 <= kTile * kStashLd * 4, "EP_DFEAT scratch must fit the stash region");
constexpr uint32_t kEpiTan = epi_bit(EP_TANGENT);
constexpr uint32_t kEpiBwd = epi_bit(EP_BACKWARD);

}  // namespace tc
}  // namespace svs

// SVS_ENGINE_TC, SDF forward chain, "two tiles in flight" kernel (tc_fwd2_kernel).
//
// The generic chain kernel (mlp_tc.cuh) walks ONE 128-point tile at a time, so every layer pays the
// MMA -> epilogue -> MMA dependency latency and the SM idles between the hand-offs.  The forward chain of the SDF
// net (ImplicitNetwork.forward / get_sdf_vals / the forward half of get_outputs, volsdf/model/network.py:71-88,
// 105-131) needs no aux tiles, which leaves room for a second activation tile in shared memory.  This kernel
// keeps TWO tiles resident and ping-pongs them: while the 16 epilogue warps turn tile 0's accumulator into the
// next layer's A operand, the tensor pipe multiplies tile 1, and vice versa.  All hand-offs are whole-tile:
//
//   a_full[T]   (count = epilogue warps)  A_T holds the operand of the next step and acc_T has been drained
//   acc_full[T] (tcgen05.commit)          acc_T holds the step's result and A_T has been read by the tensor pipe
//   s_free[T]   (store lane)              the bulk store that saves A_T (training forward) has read the tile
//
// TMEM: acc_T = columns [256 T, 256 T + 256).  Shared memory: A[2][4 blocks] 128 KB, weight ring 2 x 32 KB,
// bias table, PE column table.  Softplus(beta=100) costs ONE MUFU per element here: with t = 100 z log2(e),
//   softplus(z) = ln2/100 * ( max(t, 0) + log2(1 + 2^-|t|) ),   log2(1 + u) ~ u q(u) on [0, 1] (|err| <= 1.1e-4,
// i.e. 7e-7 in the activation — a fifth of an fp16 ulp at the smallest activations the term matters for).
// The chain description (TcChain) is the one the generic kernel takes; launch_chain() routes PRO_PE chains here.
#pragma once
#include "mlp_tc.cuh"

namespace svs {
namespace tc {

// MODE 0: SDF forward (K <= 256: 4 blocks per tile, bias + PE tables in shared memory)
// MODE 1: rendering-net forward (first layer K = F + 64 = 320: 5 blocks per tile; biases come through the read-only
//         cache because the two 80 KB tiles and the weight ring leave no room for a table)
constexpr int kF2Sdf = 0, kF2Render = 1;
constexpr int kF2PeCols = 128;
constexpr int kF2NW = 16;                                     // epilogue warps
constexpr int kF2Threads = kCtrlThreads + kF2NW * 32;
template <int MODE>
struct F2Cfg {
  static constexpr int kMaxKB = MODE == kF2Render ? 5 : 4;
  static constexpr int kOffA = 0;                                          // A[2][kMaxKB] blocks
  static constexpr int kOffW = 2 * kMaxKB * kBlk;
  static constexpr int kOffBtab = kOffW + kWSlots * kWSlot;
  static constexpr int kBtabBytes = MODE == kF2Render ? 0 : kMaxSteps * 256 * 4;
  static constexpr int kOffPe = kOffBtab + kBtabBytes;                     // 128 PE column entries of 8 bytes
  static constexpr int kPeBytes = MODE == kF2Render ? 0 : kF2PeCols * 8;
  static constexpr int kOffBar = kOffPe + kPeBytes;
  static constexpr int kSmemBytes = kOffBar + 256;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
};

struct F2Bars {
  uint64_t w_full[kWSlots], w_empty[kWSlots], acc_full[2], a_full[2], s_free[2];
  uint32_t tmem;
};
static_assert(sizeof(F2Bars) <= 256, "barrier block overflows its reservation");

// log2(1 + u) = u q(u), u in [0, 1]; minimax fit (max abs error 1.03e-4)
constexpr float kL2pC1 = 1.4390146732330322f, kL2pC2 = -0.6799439787864685f, kL2pC3 = 0.32559555768966675f,
                kL2pC4 = -0.08476857841014862f;
// Softplus(beta=100)(z) * scale from t = 100 z log2(e): c = ln2/100 * scale.  torch's threshold (20) needs no branch:
// for 100 z > 20 the correction term is < 2.1e-9 * c and vanishes in the fp16 rounding of the result.
__device__ __forceinline__ float softplus_t(float t, float c) {
  const float u = ex2_approx(-fabsf(t));
  float q = fmaf(kL2pC4, u, kL2pC3);
  q = fmaf(q, u, kL2pC2);
  q = fmaf(q, u, kL2pC1);
  return fmaf(q, u, fmaxf(t, 0.f)) * c;
}

// A chain can run on this kernel when it is a PE prologue followed by softplus / output steps that fit K <= 256
// and either every generation of A is saved or none is.
static bool fwd2_supports(const TcChain& ch) {
  const bool render = ch.prologue == PRO_RENDER_IN;
  if (ch.prologue != PRO_PE && !render) return false;
  const int max_kb = render ? F2Cfg<kF2Render>::kMaxKB : F2Cfg<kF2Sdf>::kMaxKB;
  if (ch.pro_kb > max_kb || ch.pro_colsum >= 0 || ch.n_steps < 1) return false;
  if (!render && (ch.d_in > 4 || ch.d_in * (1 + 2 * ch.n_freqs) > kF2PeCols)) return false;
  if (render && ((ch.F & 63) != 0 || ch.F > 256 || ch.pro_kb != ch.F / 64 + 1)) return false;
  const bool save = ch.pro_save >= 0;
  for (int s = 0; s < ch.n_steps; ++s) {
    const TcStep& st = ch.st[s];
    if (st.KB > max_kb || st.next_kb > max_kb || st.n_pad > 256 || st.colsum >= 0 || st.aux1 >= 0 || st.aux2 >= 0) return false;
    if (st.epi == (render ? EP_RELU : EP_SOFTPLUS)) {
      if (st.next_kb <= 0 || (st.save >= 0) != save) return false;
    } else if (render ? st.epi == EP_RGB : (st.epi == EP_SDF || st.epi == EP_Y)) {
      if (st.next_kb != 0) return false;
    } else {
      return false;
    }
  }
  return true;
}

template <int MODE>
__global__ void __launch_bounds__(kF2Threads, 1) tc_fwd2_kernel(const __grid_constant__ TcChain ch) {
  typedef F2Cfg<MODE> Cfg;
  constexpr int kF2MaxKB = Cfg::kMaxKB;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem + Cfg::kOffA;
  uint8_t* sW = smem + Cfg::kOffW;
  float* btab = reinterpret_cast<float*>(smem + Cfg::kOffBtab);
  PeEntry* petab = reinterpret_cast<PeEntry*>(smem + Cfg::kOffPe);
  F2Bars* bars = reinterpret_cast<F2Bars*>(smem + Cfg::kOffBar);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool all_save = ch.pro_save >= 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWSlots; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->acc_full[i], 1);
      mbar_init(&bars->a_full[i], kF2NW);
      mbar_init(&bars->s_free[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem;

  // this CTA's tiles: blockIdx.x + i * gridDim.x, i < n_mine; round r works on i = 2r (slot 0) and 2r + 1 (slot 1)
  const int n_mine = ch.n_tiles > (int)blockIdx.x ? (ch.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int n_rounds = (n_mine + 1) >> 1;
  auto tile_of = [&](int r, int T) -> int {
    const int i = 2 * r + T;
    return i < n_mine ? (int)blockIdx.x + i * (int)gridDim.x : -1;
  };

  if (warp == 0) {
    // ===== weight producer =====
    if (lane == 0) {
      uint32_t seq = 0;
      for (int r = 0; r < n_rounds; ++r)
        for (int s = 0; s < ch.n_steps; ++s) {
          const TcStep& st = ch.st[s];
          const uint32_t bytes = (uint32_t)st.n_pad * 128u;
          for (int T = 0; T < 2; ++T) {
            if (tile_of(r, T) < 0) continue;
            for (int kb = 0; kb < st.KB; ++kb, ++seq) {
              const int slot = seq % kWSlots;
              const uint32_t use = seq / kWSlots;
              mbar_wait(&bars->w_empty[slot], (use & 1) ^ 1);
              mbar_arrive_expect_tx(&bars->w_full[slot], bytes);
              bulk_g2s(sW + slot * kWSlot, st.w + (size_t)kb * bytes, bytes, &bars->w_full[slot]);
            }
          }
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t seq = 0, a_par = 0;
      for (int r = 0; r < n_rounds; ++r)
        for (int s = 0; s < ch.n_steps; ++s) {
          const TcStep& st = ch.st[s];
          const uint32_t idesc = make_idesc_f16(kTile, st.n_pad, 0, 0);
          for (int T = 0; T < 2; ++T) {
            if (tile_of(r, T) < 0) continue;
            mbar_wait(&bars->a_full[T], (a_par >> T) & 1);
            a_par ^= 1u << T;
            tc_fence_after();
            const uint32_t acc = tmem + (uint32_t)T * 256u;
            for (int kb = 0; kb < st.KB; ++kb, ++seq) {
              const int slot = seq % kWSlots;
              const uint32_t use = seq / kWSlots;
              mbar_wait(&bars->w_full[slot], use & 1);
              tc_fence_after();
              const uint32_t a0 = smem_u32(sA + (T * kF2MaxKB + kb) * kBlk), b0 = smem_u32(sW + slot * kWSlot);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_f16(acc, make_smem_desc(a0 + j * 32, 0, 1024), make_smem_desc(b0 + j * 32, 0, 1024), idesc, (kb | j) != 0);
              umma_commit(&bars->w_empty[slot]);
            }
            umma_commit(&bars->acc_full[T]);
          }
        }
    }
  } else if (warp == 3) {
    // ===== store lane: saves every generation of A_T (training forward) =====
    if (lane == 0 && all_save) {
      uint32_t a_par = 0;
      for (int r = 0; r < n_rounds; ++r) {
        for (int g = 0; g < ch.n_steps; ++g) {   // generation 0: the prologue's A ; generation g: A' of step g - 1
          const int kbs = g == 0 ? ch.pro_kb : ch.st[g - 1].next_kb;
          const int id = g == 0 ? ch.pro_save : ch.st[g - 1].save;
          for (int T = 0; T < 2; ++T) {
            const int t = tile_of(r, T);
            if (t < 0) continue;
            mbar_wait(&bars->a_full[T], (a_par >> T) & 1);
            a_par ^= 1u << T;
            if (kbs > 0 && id >= 0) {
              bulk_s2g(ch.img[id].base + (size_t)t * ch.img[id].tile_bytes, sA + T * kF2MaxKB * kBlk, (uint32_t)kbs * kBlk);
              bulk_commit();
              bulk_wait_read<0>();
              mbar_arrive(&bars->s_free[T]);
            }
          }
        }
      }
      bulk_wait_all<0>();
    }
  } else if (warp >= 4) {
    // ===== epilogue warps: TMEM lane quarter q, 16-column pieces pc = cg (mod 4) =====
    const int ew = warp - 4, q = ew & 3, cg = ew >> 2;
    const int m = q * 32 + lane;
    const int et = threadIdx.x - kCtrlThreads;
    const int pe_w = ch.d_in * (1 + 2 * ch.n_freqs);
    if constexpr (MODE == kF2Sdf) {
      // bias table (softplus steps: pre-multiplied by 100 log2 e) and PE column table
      for (int s = 0; s < ch.n_steps; ++s) {
        const float* b = ch.st[s].bias;
        const int nv = ch.st[s].n_valid;
        const float mul = ch.st[s].epi == EP_SOFTPLUS ? kSpK1 : 1.0f;
        if (et < 256) btab[s * 256 + et] = (b && et < nv) ? b[et] * mul : 0.f;
      }
      if (et < kF2PeCols) {
        const int c = et;
        petab[c] = pe_entry(c, ch.d_in, ch.n_freqs);
      }
      named_bar_sync(1, kF2NW * 32);
    }

    // s_free[T] is awaited lazily: `pend` bit T = the last generation of A_T is being saved and the bulk store may still
    // read the tile; whoever reuses A_T next (new generation, or scratch of the y stores) waits for it first
    uint32_t acc_par = 0, sf_par = 0, pend = 0;
    auto a_tile_free = [&](int T) {
      if (pend & (1u << T)) {
        mbar_wait(&bars->s_free[T], (sf_par >> T) & 1);
        sf_par ^= 1u << T;
        pend &= ~(1u << T);
      }
    };
    for (int r = 0; r < n_rounds; ++r) {
      float xv[2][4];
      bool live[2];
      int64_t pt[2];
      // ---------------- prologue: A_T = PE(x) ----------------
#pragma unroll
      for (int T = 0; T < 2; ++T) {
        const int t = tile_of(r, T);
        pt[T] = (int64_t)(t < 0 ? 0 : t) * kTile + m;
        live[T] = t >= 0 && pt[T] < ch.P;
        xv[T][0] = xv[T][1] = xv[T][2] = xv[T][3] = 0.f;
        if (t < 0) continue;
        if constexpr (MODE == kF2Sdf) {
          if (live[T]) {
#pragma unroll
            for (int d = 0; d < 4; ++d)
              if (d < ch.d_in) xv[T][d] = ch.x[pt[T] * ch.d_in + d];
          }
        }
        a_tile_free(T);
        uint8_t* A = sA + T * kF2MaxKB * kBlk;
        if constexpr (MODE == kF2Sdf) {
          for (int pc = cg; pc < ch.pro_kb * 4; pc += 4) {
            float v[16];
            const int c0 = pc * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = (c0 + i < pe_w) ? pe_eval(xv[T], petab[c0 + i]) : 0.f;
            st_row16(A + (pc >> 2) * kBlk, m, pc & 3, v);
          }
        } else {
          // A = [feat (F columns) | points(3) if idr, PE(view), normals(3) if idr]; this warp's pieces are cg, cg + 4, ...
          // The feature loads of two pieces are issued together (8 x 128 bits in flight per thread).
          // features: fp32 row-major.  Warp ew converts rows 8 ew .. 8 ew + 7: a lane reads columns lane, lane + 32, ...
          // (one 128-byte line per instruction) and drops the fp16 value at its place in the swizzled tile.
          const int nfp = (ch.F >> 6) * 4;   // feature pieces
          const int64_t trow0 = pt[T] - m;   // first point of the tile
          for (int rb = 0; rb < 8; rb += 2) {
            float fv[2][8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int64_t pr = trow0 + ew * 8 + rb + h;
              const float* frow = ch.feat + pr * ch.ld_feat;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int c = lane + 32 * j;
                fv[h][j] = (pr < ch.P && tile_of(r, T) >= 0 && c < ch.F) ? __ldg(frow + c) : 0.f;
              }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int row = ew * 8 + rb + h;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int c = lane + 32 * j;
                if (c < ch.F) {
                  const __half hv = __float2half_rn(fminf(fmaxf(fv[h][j], -65504.f), 65504.f));
                  *reinterpret_cast<__half*>(A + (c >> 6) * kBlk + chunk_off(row, (c & 63) >> 3) + (c & 7) * 2) = hv;
                }
              }
            }
          }
          {
            // the small block: piece nfp + cg
            const int pe_v = 3 * (1 + 2 * ch.view_freqs);
            const int o_view = ch.idr ? 3 : 0, o_n = o_view + pe_v, n_small = o_n + (ch.idr ? 3 : 0);
            float vv[4] = {0.f, 0.f, 0.f, 0.f};
            if (live[T] && cg * 16 < n_small) { vv[0] = ch.view[pt[T] * 3]; vv[1] = ch.view[pt[T] * 3 + 1]; vv[2] = ch.view[pt[T] * 3 + 2]; }
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = cg * 16 + i;
              float rv = 0.f;
              if (live[T] && c < n_small) {
                if (c < o_view) rv = ch.points[pt[T] * 3 + c];
                else if (c < o_n) rv = pe_eval(vv, pe_entry(c - o_view, 3, ch.view_freqs));
                else rv = ch.normals[pt[T] * 3 + (c - o_n)];
              }
              v[i] = rv;
            }
            st_row16(A + (nfp >> 2) * kBlk, m, cg, v);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->a_full[T]);
        if (all_save) pend |= 1u << T;
      }

      // ---------------- steps ----------------
      for (int s = 0; s < ch.n_steps; ++s) {
        const TcStep st = ch.st[s];
        const float* btab_s = btab + s * 256;
        const bool writes_a = st.next_kb > 0;
        const bool last = s + 1 == ch.n_steps;
        const float csp = kSpK2 * st.scale;
        const int npc = max((st.n_pad + 15) >> 4, st.next_kb * 4);
#pragma unroll
        for (int T = 0; T < 2; ++T) {
          if (tile_of(r, T) < 0) continue;
          const int64_t p = pt[T];
          uint8_t* A = sA + T * kF2MaxKB * kBlk;
          const uint32_t tm_acc = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)T * 256u;
          mbar_wait(&bars->acc_full[T], (acc_par >> T) & 1);
          acc_par ^= 1u << T;
          tc_fence_after();
          uint32_t rr[16];
          if (cg * 16 < st.n_pad) tmem_ld_32x16(tm_acc + (uint32_t)(cg * 16), rr);
          // the tile is rewritten (A') or used as scratch (coalesced y stores): the bulk store that saves the previous
          // generation must have read it
          if (writes_a || (st.epi == EP_Y && last)) a_tile_free(T);
          for (int pc = cg; pc < npc; pc += 4) {
            const int col0 = pc * 16;
            float acc[16];
            if (col0 < st.n_pad) {
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(rr[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[i] = 0.f;
            }
            if (col0 + 64 < st.n_pad) tmem_ld_32x16(tm_acc + (uint32_t)(col0 + 64), rr);
            if constexpr (MODE == kF2Sdf) {
              if (st.epi == EP_SOFTPLUS) {
                float o[16];
                if (col0 + 16 <= st.n_valid) {
                  const float4* b4 = reinterpret_cast<const float4*>(btab_s + col0);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float4 b = b4[i];
                    o[4 * i + 0] = softplus_t(fmaf(acc[4 * i + 0], kSpK1, b.x), csp);
                    o[4 * i + 1] = softplus_t(fmaf(acc[4 * i + 1], kSpK1, b.y), csp);
                    o[4 * i + 2] = softplus_t(fmaf(acc[4 * i + 2], kSpK1, b.z), csp);
                    o[4 * i + 3] = softplus_t(fmaf(acc[4 * i + 3], kSpK1, b.w), csp);
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const int n = col0 + i;
                    float v = 0.f;
                    if (n < st.n_valid) v = softplus_t(fmaf(acc[i], kSpK1, btab_s[n & 255]), csp);
                    else if ((st.flags & TC_PEFILL) && n - st.n_valid < pe_w) v = pe_eval(xv[T], petab[n - st.n_valid]) * st.scale;
                    o[i] = v;
                  }
                }
                st_row16(A + (pc >> 2) * kBlk, m, pc & 3, o);
              } else if (st.epi == EP_SDF) {
                if (col0 == 0 && live[T]) {
                  float y0 = acc[0] + btab_s[0];
                  if (ch.radius > 0.f && p < ch.n_clamped) {
                    // coordinates beyond d_in are zero
                    const float n2 = xv[T][0] * xv[T][0] + xv[T][1] * xv[T][1] + xv[T][2] * xv[T][2] + xv[T][3] * xv[T][3];
                    y0 = fminf(y0, ch.sph_scale * (ch.radius - sqrtf(n2)));
                  }
                  ch.sdf[p] = y0;
                }
              } else {   // EP_Y
                if (st.n_valid == 1) {
                  if (col0 == 0 && live[T]) ch.y[p * ch.ldy + st.y_col] = acc[0] + btab_s[0];
                } else if (!last) {   // a later step still multiplies A_T: no scratch, direct (uncoalesced) stores
                  if (live[T]) {
                    float* dst = ch.y + p * ch.ldy + st.y_col + col0;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                      if (col0 + i < st.n_valid) dst[i] = acc[i] + btab_s[(col0 + i) & 255];
                  }
                } else {
                  // y is fp32 row-major with a thread-per-row register layout: a direct store touches 32 lines per
                  // instruction.  Transpose the 32 x 16 piece through this warp's scratch (inside the now free A_T) so
                  // that every store instruction writes two rows x 16 contiguous floats.
                  float* scr = reinterpret_cast<float*>(A) + ew * (32 * 17);
#pragma unroll
                  for (int i = 0; i < 16; ++i) scr[lane * 17 + i] = acc[i] + btab_s[(col0 + i) & 255];
                  __syncwarp();
                  const int cc = lane & 15, n = col0 + cc;
                  const int64_t row0 = pt[T] - lane;   // first row of this warp's lane quarter
#pragma unroll
                  for (int k = 0; k < 16; ++k) {
                    const int rr_ = 2 * k + (lane >> 4);
                    const int64_t pr = row0 + rr_;
                    if (pr < ch.P && n < st.n_valid) ch.y[pr * ch.ldy + st.y_col + n] = scr[rr_ * 17 + cc];
                  }
                  __syncwarp();
                }
              }
            } else {
              if (st.epi == EP_RELU) {
                float o[16];
                if (col0 + 16 <= st.n_valid) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) o[i] = fmaxf(acc[i] + __ldg(st.bias + col0 + i), 0.f);
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i) o[i] = (col0 + i < st.n_valid) ? fmaxf(acc[i] + __ldg(st.bias + col0 + i), 0.f) : 0.f;
                }
                st_row16(A + (pc >> 2) * kBlk, m, pc & 3, o);
              } else {   // EP_RGB
                if (live[T]) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const int n = col0 + i;
                    if (n < st.n_valid) ch.rgb[p * st.n_valid + n] = 1.0f / (1.0f + __expf(-(acc[i] + __ldg(st.bias + n))));
                  }
                }
              }
            }
          }
          if (!last) {
            // A_T rewritten and / or acc_T drained: the next step's MMAs of this tile may go
            if (writes_a) fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->a_full[T]);
          }
          if (writes_a && all_save) pend |= 1u << T;
        }
      }
      tc_fence_before();
      // the y stores of the last step use A_T as scratch: every warp must be done with it before any warp starts the
      // next round's prologue
      if (MODE == kF2Sdf && ch.st[ch.n_steps - 1].epi == EP_Y) named_bar_sync(1, kF2NW * 32);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

template <int MODE>
static int launch_fwd2_t(const TcChain& ch, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SVS_CUDA_OK(cudaFuncSetAttribute(tc_fwd2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2Cfg<MODE>::kSmemBytes));
    attr_set = true;
  }
  tc_fwd2_kernel<MODE><<<grid, kF2Threads, F2Cfg<MODE>::kSmemBytes, st>>>(ch);
  return SVS_OK;
}
static int launch_fwd2(const TcChain& ch, int grid, cudaStream_t st) {
  return ch.prologue == PRO_RENDER_IN ? launch_fwd2_t<kF2Render>(ch, grid, st) : launch_fwd2_t<kF2Sdf>(ch, grid, st);
}

}  // namespace tc
}  // namespace svs

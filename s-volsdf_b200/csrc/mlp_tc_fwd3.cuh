// SVS_ENGINE_TC_SPLIT, forward chains with split operands (tc_fwd3_kernel).
//
// fp16 operands carry 11 significant bits: behind Softplus(beta=100) that is ~1e-3 in sdf, 1.3e-3 .. 2e-2 in depth
// (beta = 0.05 .. 0.001) and, through ReLU-mask / L1-sign flips, 1-3 % in the rendering-net gradients
// (tools/precision_sim.py reproduces the measured figures on the CPU) — outside the 1e-3 / 1e-2 contract.  Everything
// the reference's *outputs* depend on is a forward chain: sampler sdf (ImplicitNetwork.get_sdf_vals,
// volsdf/model/network.py:125-131), get_outputs' forward (:71-88,105-112) and RenderingNetwork.forward (:170-190).
// This kernel runs those chains with every operand split into hi + lo fp16 halves (22 significant bits) and three
// tensor-core passes per layer,
//
//     acc = A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T          (A_lo W_lo^T ~ 2^-22 is dropped)
//
// fp32 accumulation in TMEM as before.  Result: sdf within 1e-6, depth within 1e-5 of the fp32 reference, ReLU masks
// identical up to pre-activations below 1e-6.  The backward chains keep single fp16 operands (their rounding errors are
// zero-mean per element and average out over the points: parameter gradients 1e-3 .. 4e-3).
//
// Structure: one 128-point tile per CTA with A_hi and A_lo resident (2 x 64 KB, rendering net 2 x 80 KB), weight ring
// of 32 KB slots, MMA and epilogue overlapped at 16-column granularity exactly like tc_chain_kernel (a_ready per
// piece), two alternating 256-column accumulators.  Per layer the cross terms are issued first, the main term last (the
// accumulator write-back of tcgen05.mma rounds toward zero; see the MMA issuer).
// Softplus uses ex2 + lg2 (two MUFU, abs. error 2e-9 in the activation); the PE prologue uses sincosf.
#pragma once
#include "mlp_tc_fwd2.cuh"

namespace svs {
namespace tc {

#ifndef SVS_F3_EXP
#define SVS_F3_EXP 0   // measurement builds only (tools/f3_exp.sh): 1 no MMA, 2 no MUFU, 4 no A stores, 8 no weight copies, 16 idle epilogue, 32 main term of the last block only
#endif
constexpr int kF3NW = 16;
constexpr int kF3Threads = kCtrlThreads + kF3NW * 32;
constexpr int kF3MaxSlots = 6;
// PAIR: two CTAs of a cluster work as one cta_group::2 unit.  Every CTA keeps its own 128-point tile (A_hi, A_lo, its
// TMEM accumulators) and streams only HALF of every weight block (128 of the 256 output rows); one tcgen05.mma with
// M = 256 multiplies both tiles by the whole block.  The kernel is bound by shared-memory bandwidth (per layer and tile:
// 48 instructions x 12 KB of operands + 384 KB of weight fills + 128 KB of activation stores against 128 B/clk), and the
// pair halves the weight side of it: 8 KB of operands per instruction and CTA, 192 KB of fills.
template <int MODE, bool PAIR>
struct F3Cfg {
  static constexpr int kMaxKB = MODE == kF2Render ? 5 : 4;
  static constexpr int kSlotBytes = PAIR ? kWSlot / 2 : kWSlot;
  static constexpr int kSlots = (MODE == kF2Render ? 2 : 3) * (PAIR ? 2 : 1);
  static constexpr int kOffAhi = 0;
  static constexpr int kOffAlo = kMaxKB * kBlk;
  static constexpr int kOffW = 2 * kMaxKB * kBlk;
  static constexpr int kOffPe = kOffW + kSlots * kSlotBytes;
  static constexpr int kPeBytes = MODE == kF2Render ? 0 : kF2PeCols * 8;
  static constexpr int kOffBar = kOffPe + kPeBytes;
  static constexpr int kSmemBytes = kOffBar + 512;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
};

struct F3Bars {
  // w_peer (PAIR, leader CTA only): the peer CTA's half of the slot has landed
  uint64_t w_full[kF3MaxSlots], w_empty[kF3MaxSlots], w_peer[kF3MaxSlots], a_ready[kMaxKB], s_free[kMaxKB], acc_full;
  uint32_t tmem;
};
static_assert(sizeof(F3Bars) <= 512, "barrier block overflows its reservation");

// ---- cta_group::2 building blocks (cluster of two CTAs) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  Default (CTA-scope release)
// semantics: the data the arrival publishes stays in the arriving CTA's shared memory, where fence.proxy.async has made it
// visible to the tensor core; an explicit .release.cluster costs ~900 cycles per arrival (the peer's epilogue fell 3.7 k
// cycles per layer behind the leader's: tools/f3_trace.py)
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// wait that also acquires the writes released by arrivals from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAITC_LOOP_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAITC_DONE_%=;\n\t"
      "bra WAITC_LOOP_%=;\n\t"
      "WAITC_DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all MMAs issued so far arrives on `bar` in both CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// hi / lo halves of 16 consecutive columns of row m
template <bool PAIR>
__device__ __forceinline__ void f3_umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
  if (SVS_F3_EXP & 1) return;
  if constexpr (PAIR) umma_f16_2cta(d, a, b, idesc, accum);
  else umma_f16(d, a, b, idesc, accum);
}
__device__ __forceinline__ void st_row16_split(uint8_t* blk_hi, uint8_t* blk_lo, int m, int cq, const float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = v[8 * i + 2 * k], b = v[8 * i + 2 * k + 1];
      h[k] = pack_h2(a, b);
      const float2 f = unpack_h2(h[k]);
      l[k] = pack_h2(a - f.x, b - f.y);
    }
    const uint32_t off = chunk_off(m, cq * 2 + i);
    *reinterpret_cast<uint4*>(blk_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(blk_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}
__device__ __forceinline__ void st_elem_split(uint8_t* a_hi, uint8_t* a_lo, int row, int c, float v) {
  const uint32_t off = (uint32_t)(c >> 6) * kBlk + chunk_off(row, (c & 63) >> 3) + (uint32_t)(c & 7) * 2u;
  const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  *reinterpret_cast<__half*>(a_hi + off) = h;
  *reinterpret_cast<__half*>(a_lo + off) = __float2half_rn(v - __half2float(h));
}

// Softplus(beta=100)(z) * scale from t = 100 z log2(e): c = ln2/100 * scale (threshold: see softplus_t)
__device__ __forceinline__ float softplus_t2(float t, float c) {
  if (SVS_F3_EXP & 2) return (fmaxf(t, 0.f) + fmaf(fabsf(t), 1e-3f, 1.0f)) * c;
  return (fmaxf(t, 0.f) + lg2_approx(1.0f + ex2_approx(-fabsf(t)))) * c;
}

// One 16-column piece of a full softplus layer as ONE basic block: bias, softplus, hi/lo split and the stores, so that
// the scheduler can overlap the MUFU pipe (2 per element) with the conversion / store instructions of the same piece.
__device__ __forceinline__ void softplus_piece_split(const uint32_t (&rr)[16], const float* bias_t, float rzk, float csp,
                                                     uint8_t* blk_hi, uint8_t* blk_lo, int m, int cq) {
  const float4* b4 = reinterpret_cast<const float4*>(bias_t);
  float o[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 b = __ldg(b4 + i);
    o[4 * i + 0] = softplus_t2(fmaf(__uint_as_float(rr[4 * i + 0]), rzk, b.x), csp);
    o[4 * i + 1] = softplus_t2(fmaf(__uint_as_float(rr[4 * i + 1]), rzk, b.y), csp);
    o[4 * i + 2] = softplus_t2(fmaf(__uint_as_float(rr[4 * i + 2]), rzk, b.z), csp);
    o[4 * i + 3] = softplus_t2(fmaf(__uint_as_float(rr[4 * i + 3]), rzk, b.w), csp);
  }
  if (SVS_F3_EXP & 4) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += o[i];
    if (acc == 123.456f) blk_hi[m] = 1;
    return;
  }
  st_row16_split(blk_hi, blk_lo, m, cq, o);
}

__device__ __forceinline__ float pe_eval_precise(const float (&xv)[4], PeEntry e) {
  const int dim = e.code & 15, kind = e.code >> 4;
  const float xs = dim == 0 ? xv[0] : (dim == 1 ? xv[1] : (dim == 2 ? xv[2] : xv[3]));
  const float arg = xs * e.mult;
  return kind == 1 ? xs : (kind == 2 ? sinf(arg) : (kind == 3 ? cosf(arg) : 0.f));
}

// forward chains only: PE / rendering-input prologue, softplus / relu steps that rewrite A, output steps that do not
static bool fwd3_supports(const TcChain& ch) {
  const bool render = ch.prologue == PRO_RENDER_IN;
  if (ch.prologue != PRO_PE && !render) return false;
  const int max_kb = render ? F3Cfg<kF2Render, false>::kMaxKB : F3Cfg<kF2Sdf, false>::kMaxKB;
  if (ch.pro_kb > max_kb || ch.pro_colsum >= 0 || ch.n_steps < 1) return false;
  if (!render && (ch.d_in > 4 || ch.d_in * (1 + 2 * ch.n_freqs) > kF2PeCols)) return false;
  if (render && ((ch.F & 63) != 0 || ch.F > 256 || ch.pro_kb != ch.F / 64 + 1)) return false;
  for (int s = 0; s < ch.n_steps; ++s) {
    const TcStep& st = ch.st[s];
    if (!st.w_lo || st.KB > max_kb || st.next_kb > max_kb || st.n_pad > 256 || st.colsum >= 0 || st.aux1 >= 0 || st.aux2 >= 0) return false;
    if (st.epi == (render ? EP_RELU : EP_SOFTPLUS)) {
      if (st.next_kb <= 0 || s + 1 >= ch.n_steps || st.next_kb != ch.st[s + 1].KB) return false;
      if (!render && !st.bias_t) return false;
    } else if (render ? st.epi == EP_RGB : (st.epi == EP_SDF || st.epi == EP_Y)) {
      if (st.next_kb != 0 || (s + 1 < ch.n_steps && ch.st[s + 1].KB != st.KB)) return false;
    } else {
      return false;
    }
  }
  return ch.st[0].KB == ch.pro_kb;
}

template <int MODE, bool PAIR>
__global__ void __launch_bounds__(kF3Threads, 1) tc_fwd3_kernel(const __grid_constant__ TcChain ch) {
  typedef F3Cfg<MODE, PAIR> Cfg;
  constexpr int NS = Cfg::kSlots;
  constexpr int NCG = kF3NW / 4;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sAhi = smem + Cfg::kOffAhi;
  uint8_t* sAlo = smem + Cfg::kOffAlo;
  uint8_t* sW = smem + Cfg::kOffW;
  PeEntry* petab = reinterpret_cast<PeEntry*>(smem + Cfg::kOffPe);
  F3Bars* bars = reinterpret_cast<F3Bars*>(smem + Cfg::kOffBar);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // PAIR: CTA `rank` of the cluster takes tile 2 i + rank of tile pair i (an odd tile count leaves the last pair's second
  // CTA a tile behind the end: it runs the barrier protocol on zeros and neither reads nor writes global memory)
  uint32_t rank = 0;
  if constexpr (PAIR) rank = cluster_ctarank();
#define F3_TILES(t) int t = (PAIR ? 2 * ((int)blockIdx.x >> 1) + (int)rank : (int)blockIdx.x); t < (PAIR ? 2 * ((ch.n_tiles + 1) >> 1) : ch.n_tiles); t += (int)gridDim.x
  // chains that save nothing (sampler sdf, eval forward): no store lane, no hand-back of the blocks
  bool any_save = ch.pro_save >= 0;
  for (int s = 0; s < ch.n_steps; ++s) any_save |= ch.st[s].save >= 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); mbar_init(&bars->w_peer[i], 1); }
    // one per 64-column block: every epilogue warp writes one 16-column piece of it; the pair's leader also counts the peer's
    for (int i = 0; i < kMaxKB; ++i) mbar_init(&bars->a_ready[i], (PAIR && rank == 0) ? 2 * kF3NW : kF3NW);
    for (int i = 0; i < kMaxKB; ++i) mbar_init(&bars->s_free[i], 1);
    mbar_init(&bars->acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_2cta(&bars->tmem, 512);
    else tmem_alloc(&bars->tmem, 512);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem = bars->tmem;

  if (warp == 0) {
    // ===== weight producer: per step the cross-term blocks (W_lo(kb), W_hi(kb) for every kb), then the W_hi blocks again
    //       for the main term (see the MMA issuer for why the main term goes last).  PAIR: this CTA's half of the rows. =====
    if (lane == 0) {
      uint32_t seq = 0;
      auto load = [&](const uint8_t* src, uint32_t bytes) {
        const int slot = seq % NS;
        const uint32_t use = seq / NS;
        ++seq;
        mbar_wait(&bars->w_empty[slot], (use & 1) ^ 1);
        if (SVS_F3_EXP & 8) { mbar_arrive(&bars->w_full[slot]); return; }
        mbar_arrive_expect_tx(&bars->w_full[slot], bytes);
        bulk_g2s(sW + slot * Cfg::kSlotBytes, src, bytes, &bars->w_full[slot]);
      };
      for (F3_TILES(t)) {
        for (int s = 0; s < ch.n_steps; ++s) {
          const TcStep& st = ch.st[s];
          const uint32_t blk = (uint32_t)st.n_pad * 128u;                 // one 64-column block of the image
          const uint32_t bytes = PAIR ? blk / 2 : blk;                      // rows [rank n_pad / 2, +n_pad / 2) of it
          const size_t off = PAIR ? (size_t)rank * bytes : 0;
          for (int kb = 0; kb < st.KB; ++kb) {
            load(st.w_lo + (size_t)kb * blk + off, bytes);
            load(st.w + (size_t)kb * blk + off, bytes);
          }
          // main term, last block first: W_hi(KB-1) is still in its slot from the cross term
          for (int kb = st.KB - 2; kb >= 0; --kb) load(st.w + (size_t)kb * blk + off, bytes);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer.  The tensor core truncates the fp32 accumulator toward zero after EVERY instruction (probed:
    //       tools/tc_acc_probe.cu — products of one instruction are summed exactly, the write-back is RZ), a bias of
    //       ~0.5 ulp(partial sum) per instruction that adds up coherently over the layers.  So the two cross terms are
    //       accumulated FIRST, while the accumulator is ~2^-11 of its final size (their 2 x 4 KB instructions truncate at
    //       2^-35 of the result), and only the 4 KB main-term instructions see the full magnitude: 16 truncations per
    //       layer instead of 48 (measured sdf error 2.4e-5 -> see DESIGN.md).  Cross-term k-steps trail the epilogue of
    //       the previous step at 16-column granularity. =====
    if (PAIR && rank != 0) {
      // the peer has no MMAs to issue (the leader's instructions drive both tensor cores): its lane relays "my half of the
      // slot has landed" to the leader
      if (lane == 0) {
        uint32_t seq = 0;
        for (F3_TILES(t)) {
          for (int s = 0; s < ch.n_steps; ++s) {
            const int n_items = 3 * ch.st[s].KB - 1;
            for (int i = 0; i < n_items; ++i) {
              const int slot = seq % NS;
              const uint32_t use = seq / NS;
              ++seq;
              mbar_wait(&bars->w_full[slot], use & 1);
              mbar_arrive_cluster(&bars->w_peer[slot], 0);
            }
          }
        }
      }
    } else if (lane == 0) {
      uint32_t seq = 0, n_step = 0, a_par = 0;
#ifdef SVS_F3_TRACE
      const int trace_lane = 0;
      uint32_t trace_n = 0;
#endif
      // waits for the weight slot (PAIR: both halves) ; operand blocks published by the epilogue (PAIR: of both CTAs)
      auto wait_w = [&](int slot, uint32_t use) {
        mbar_wait(&bars->w_full[slot], use & 1);
        if constexpr (PAIR) mbar_wait_cluster(&bars->w_peer[slot], use & 1);
      };
      auto wait_a = [&](int kb) {
        if constexpr (PAIR) mbar_wait_cluster(&bars->a_ready[kb], (a_par >> kb) & 1);
        else mbar_wait(&bars->a_ready[kb], (a_par >> kb) & 1);
        a_par ^= 1u << kb;
      };
      auto commit = [&](uint64_t* bar) {
        if constexpr (PAIR) umma_commit_2cta(bar);
        else umma_commit(bar);
      };
      for (F3_TILES(t)) {
        for (int s = 0; s < ch.n_steps; ++s, ++n_step) {
          const TcStep& st = ch.st[s];
          const uint32_t idesc = make_idesc_f16(PAIR ? 2 * kTile : kTile, st.n_pad, 0, 0);
          const uint32_t acc = tmem + (n_step & 1) * 256;
          [[maybe_unused]] const bool tr = blockIdx.x == 0 && t == 2 * (int)gridDim.x;
          F3_EV(tr, 1, s, 0);
          for (int kb = 0; kb < st.KB; ++kb) {
            const uint32_t ah = smem_u32(sAhi + kb * kBlk), al = smem_u32(sAlo + kb * kBlk);
            {   // A_hi(kb) x W_lo(kb)
              const int slot = seq % NS;
              const uint32_t use = seq / NS;
              ++seq;
              wait_w(slot, use);
              F3_EV(tr, 2, s, kb);
              wait_a(kb);
              F3_EV(tr, 3, s, kb);
              tc_fence_after();
              const uint32_t b0 = smem_u32(sW + slot * Cfg::kSlotBytes);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                f3_umma<PAIR>(acc, make_smem_desc(ah + j * 32, 0, 1024), make_smem_desc(b0 + j * 32, 0, 1024), idesc, (kb | j) != 0);
              commit(&bars->w_empty[slot]);
            }
            {   // A_lo(kb) x W_hi(kb); the last block stays for the first main-term block
              const int slot = seq % NS;
              const uint32_t use = seq / NS;
              ++seq;
              wait_w(slot, use);
              F3_EV(tr, 4, s, kb);
              tc_fence_after();
              const uint32_t b0 = smem_u32(sW + slot * Cfg::kSlotBytes);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                f3_umma<PAIR>(acc, make_smem_desc(al + j * 32, 0, 1024), make_smem_desc(b0 + j * 32, 0, 1024), idesc, 1);
              if (kb == st.KB - 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  f3_umma<PAIR>(acc, make_smem_desc(ah + j * 32, 0, 1024), make_smem_desc(b0 + j * 32, 0, 1024), idesc, 1);
              }
              commit(&bars->w_empty[slot]);
            }
          }
          for (int kb = st.KB - 2; kb >= 0; --kb) {   // main term A_hi(kb) x W_hi(kb), remaining blocks
            const int slot = seq % NS;
            const uint32_t use = seq / NS;
            ++seq;
            wait_w(slot, use);
            F3_EV(tr, 5, s, kb);
            tc_fence_after();
            const uint32_t ah = smem_u32(sAhi + kb * kBlk), b0 = smem_u32(sW + slot * Cfg::kSlotBytes);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (!(SVS_F3_EXP & 32))   // measurement: what a tail half as long would buy
                f3_umma<PAIR>(acc, make_smem_desc(ah + j * 32, 0, 1024), make_smem_desc(b0 + j * 32, 0, 1024), idesc, 1);
            commit(&bars->w_empty[slot]);
          }
          commit(&bars->acc_full);
          F3_EV(tr, 6, s, 0);
        }
      }
    }
  } else if (warp == 3) {
    // ===== store lane: saves the hi half of every generation of A the backward needs, then frees the blocks =====
    if (lane == 0 && any_save) {
      uint32_t a_par = 0;
      auto consume = [&](int kb) {
        mbar_wait(&bars->a_ready[kb], (a_par >> kb) & 1);
        a_par ^= 1u << kb;
      };
      for (F3_TILES(t)) {
        const bool real = t < ch.n_tiles;   // the pair's padding tile saves nothing
        for (int b = 0; b < ch.pro_kb; ++b) consume(b);
        if (ch.pro_save >= 0 && real) {
          bulk_s2g(ch.img[ch.pro_save].base + (size_t)t * ch.img[ch.pro_save].tile_bytes, sAhi, (uint32_t)ch.pro_kb * kBlk);
          bulk_commit();
          bulk_wait_read<0>();
        }
        for (int b = 0; b < ch.pro_kb; ++b) mbar_arrive(&bars->s_free[b]);
        for (int s = 0; s < ch.n_steps; ++s) {
          const TcStep& st = ch.st[s];
          if (st.next_kb > 0) {
            for (int c = 0; c < st.next_kb; ++c) {
              consume(c);
              if (st.save >= 0 && real) {
                bulk_s2g(ch.img[st.save].base + (size_t)t * ch.img[st.save].tile_bytes + (size_t)c * kBlk, sAhi + c * kBlk, kBlk);
                bulk_commit();
              }
            }
            if (st.save >= 0 && real) bulk_wait_read<0>();
            for (int c = 0; c < st.next_kb; ++c) mbar_arrive(&bars->s_free[c]);
          } else if (s + 1 < ch.n_steps) {
            for (int c = 0; c < ch.st[s + 1].KB; ++c) consume(c);
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
    const uint32_t tm_row = tmem + ((uint32_t)(q * 32) << 16);
    if constexpr (MODE == kF2Sdf) {
      if (et < kF2PeCols) {
        const int c = et;
        petab[c] = pe_entry(c, ch.d_in, ch.n_freqs);
      }
      named_bar_sync(1, kF3NW * 32);
    }
    uint32_t n_acc = 0, fgen = 0;
#ifdef SVS_F3_TRACE
    const int trace_lane = 1 + (ew == 0 ? 0 : (ew == 5 ? 1 : 2));
    uint32_t trace_n = 0;
#endif

    // a block of A is published to the MMA issuer (PAIR: the leader's, which also counts this CTA's warps when it is the
    // peer) and to this CTA's store lane
    auto publish = [&](int kb) {
      mbar_arrive(&bars->a_ready[kb]);
      if constexpr (PAIR) {
        if (rank != 0) mbar_arrive_cluster(&bars->a_ready[kb], 0);
      }
    };
    for (F3_TILES(t)) {
      const int64_t p = (int64_t)t * kTile + m;
      const bool live = p < ch.P;
      float xv[4] = {0.f, 0.f, 0.f, 0.f};
      // ---------------- prologue ----------------
      if constexpr (MODE == kF2Sdf) {
        if (live) {
#pragma unroll
          for (int d = 0; d < 4; ++d)
            if (d < ch.d_in) xv[d] = ch.x[p * ch.d_in + d];
        }
        for (int pc = cg; pc < ch.pro_kb * 4; pc += NCG) {
          float v[16];
          const int b = pc >> 2, c0 = pc * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (c0 + i < pe_w) ? pe_eval_precise(xv, petab[c0 + i]) : 0.f;
          if (any_save) mbar_wait(&bars->s_free[b], ((fgen >> b) & 1) ^ 1);
          st_row16_split(sAhi + b * kBlk, sAlo + b * kBlk, m, pc & 3, v);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) publish(b);
        }
      } else {
        // A = [feat (F columns) | points(3) if idr, PE(view), normals(3) if idr].  Features are fp32 row-major: warp ew
        // converts rows 8 ew .. 8 ew + 7, a lane reads columns lane, lane + 32, ... (one 128-byte line per instruction).
        if (any_save)
          for (int b = 0; b < ch.pro_kb; ++b) mbar_wait(&bars->s_free[b], ((fgen >> b) & 1) ^ 1);
        const int64_t trow0 = (int64_t)t * kTile;
        for (int rb = 0; rb < 8; rb += 2) {
          float fv[2][8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int64_t pr = trow0 + ew * 8 + rb + h;
            const float* frow = ch.feat + pr * ch.ld_feat;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = lane + 32 * j;
              fv[h][j] = (pr < ch.P && c < ch.F) ? __ldg(frow + c) : 0.f;
            }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int row = ew * 8 + rb + h;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = lane + 32 * j;
              if (c < ch.F) st_elem_split(sAhi, sAlo, row, c, fv[h][j]);
            }
          }
        }
        {
          const int nfb = ch.F >> 6;
          const int pe_v = 3 * (1 + 2 * ch.view_freqs);
          const int o_view = ch.idr ? 3 : 0, o_n = o_view + pe_v, n_small = o_n + (ch.idr ? 3 : 0);
          float vv[4] = {0.f, 0.f, 0.f, 0.f};
          if (live && cg * 16 < n_small) { vv[0] = ch.view[p * 3]; vv[1] = ch.view[p * 3 + 1]; vv[2] = ch.view[p * 3 + 2]; }
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = cg * 16 + i;
            float rv = 0.f;
            if (live && c < n_small) {
              if (c < o_view) rv = ch.points[p * 3 + c];
              else if (c < o_n) {
                const int tt = c - o_view;
                if (tt < 3) rv = vv[tt];
                else {
                  const int k = (tt - 3) / 6, rem = (tt - 3) - k * 6, fn = rem / 3, dim = rem - fn * 3;
                  const float arg = vv[dim] * (float)(1 << k);
                  rv = fn ? cosf(arg) : sinf(arg);
                }
              } else rv = ch.normals[p * 3 + (c - o_n)];
            }
            v[i] = rv;
          }
          st_row16_split(sAhi + nfb * kBlk, sAlo + nfb * kBlk, m, cg, v);
        }
        fence_proxy_async();
        named_bar_sync(1, kF3NW * 32);
        if (lane == 0)
          for (int b = 0; b < ch.pro_kb; ++b) publish(b);
      }
      for (int b = 0; b < ch.pro_kb; ++b) fgen ^= 1u << b;

      // ---------------- steps ----------------
      for (int s = 0; s < ch.n_steps; ++s) {
        const TcStep st = ch.st[s];
        const bool has_next = s + 1 < ch.n_steps;
        const bool writes_a = st.next_kb > 0;
        const float csp = kSpK2 * st.scale;
        // Expected deficit of the round-toward-zero accumulator: every main-term instruction loses 0.5 ulp(partial sum)
        // on average, toward zero.  With partial sums growing linearly to the result over n = 4 KB instructions and
        // E[ulp(p) / |p|] = 2^-23 / (2 ln 2) (log-uniform mantissa) the result is short by 0.25 * 0.7213 * 2^-23 * (n + 1)
        // relative (n = 16: 3.65e-7; tools/tc_acc_probe.cu measures 2.9e-7 on random data).  Multiplying it back leaves
        // the zero-mean part of the truncation, which does not add up over the layers: sdf error 4.8e-6 -> 9.6e-7 in
        // the CPU model of this arithmetic.
        const float rz = 1.0f + 2.1496e-8f * (float)(4 * st.KB + 1);
        const float rzk = rz * kSpK1;
        const uint32_t tm_acc = tm_row + (n_acc & 1) * 256;
        mbar_wait_relaxed(&bars->acc_full, n_acc & 1);
        ++n_acc;
        tc_fence_after();
        [[maybe_unused]] const bool tr = blockIdx.x == 0 && t == 2 * (int)gridDim.x && lane == 0 && (ew == 0 || ew == 15 || ew == 5);
        F3_EV(tr, 7, s, ew);
        if (!writes_a && has_next) {
          // A is not rewritten by this step: the next step's MMAs may start at once (into the other accumulator)
          if (lane == 0)
            for (int b = 0; b < ch.st[s + 1].KB; ++b) publish(b);
        }
        if constexpr (MODE == kF2Sdf) {
          if (st.epi == EP_SOFTPLUS && st.n_valid == 256 && st.next_kb == 4) {
            // hot path (7 of the 8 hidden layers): 4 pieces per warp, unrolled with two alternating register sets for
            // the TMEM loads (no copies), every piece one basic block
            uint32_t ra[16], rb[16];
            const uint32_t tm0 = tm_acc + (uint32_t)(cg * 16);
            tmem_ld_32x16(tm0, ra);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (any_save) mbar_wait(&bars->s_free[k], ((fgen >> k) & 1) ^ 1);
              tmem_ld_wait();
              F3_EV(tr, 9, s, ew * 4 + k);
              if (k < 3) {
                if (k & 1) tmem_ld_32x16(tm0 + (uint32_t)((k + 1) * 64), ra);
                else tmem_ld_32x16(tm0 + (uint32_t)((k + 1) * 64), rb);
              }
              if (SVS_F3_EXP & 16) {
                if ((ra[0] ^ rb[1]) == 0x12345u) sAhi[m] = 1;
              } else if (k & 1) softplus_piece_split(rb, st.bias_t + k * 64 + cg * 16, rzk, csp, sAhi + k * kBlk, sAlo + k * kBlk, m, cg);
              else softplus_piece_split(ra, st.bias_t + k * 64 + cg * 16, rzk, csp, sAhi + k * kBlk, sAlo + k * kBlk, m, cg);
              F3_EV(tr, 11, s, ew * 4 + k);
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) publish(k);
              F3_EV(tr, 12, s, ew * 4 + k);
            }
            fgen ^= 15u;
            tc_fence_before();
            continue;
          }
        }
        const int npc = max((st.n_pad + 15) >> 4, st.next_kb * 4);
        uint32_t rr[16];
        if (cg * 16 < st.n_pad) tmem_ld_32x16(tm_acc + (uint32_t)(cg * 16), rr);
        for (int pc = cg; pc < npc; pc += NCG) {
          const int c = pc >> 2, col0 = pc * 16;
          float acc[16];
          if (col0 < st.n_pad) {
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(rr[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
          }
          if (col0 + NCG * 16 < st.n_pad) tmem_ld_32x16(tm_acc + (uint32_t)(col0 + NCG * 16), rr);
          float o[16];
          if constexpr (MODE == kF2Sdf) {
            if (st.epi == EP_SOFTPLUS) {
              if (col0 + 16 <= st.n_valid) {
                // t = 100 log2(e) (acc rz + b) as ONE FFMA per element: rz 100 log2(e) is uniform, b 100 log2(e) comes
                // from the table the weight packer wrote
                const float4* b4 = reinterpret_cast<const float4*>(st.bias_t + col0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 b = __ldg(b4 + i);
                  o[4 * i + 0] = softplus_t2(fmaf(acc[4 * i + 0], rzk, b.x), csp);
                  o[4 * i + 1] = softplus_t2(fmaf(acc[4 * i + 1], rzk, b.y), csp);
                  o[4 * i + 2] = softplus_t2(fmaf(acc[4 * i + 2], rzk, b.z), csp);
                  o[4 * i + 3] = softplus_t2(fmaf(acc[4 * i + 3], rzk, b.w), csp);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int n = col0 + i;
                  float v = 0.f;
                  if (n < st.n_valid) v = softplus_t2(fmaf(acc[i], rzk, __ldg(st.bias_t + n)), csp);
                  else if ((st.flags & TC_PEFILL) && n - st.n_valid < pe_w) v = pe_eval_precise(xv, petab[n - st.n_valid]) * st.scale;
                  o[i] = v;
                }
              }
            } else if (st.epi == EP_SDF) {
              if (col0 == 0 && live) {
                float y0 = fmaf(acc[0], rz, __ldg(st.bias));
                if (ch.radius > 0.f && p < ch.n_clamped) {
                  const float n2 = xv[0] * xv[0] + xv[1] * xv[1] + xv[2] * xv[2] + xv[3] * xv[3];
                  y0 = fminf(y0, ch.sph_scale * (ch.radius - sqrtf(n2)));
                }
                ch.sdf[p] = y0;
              }
            } else {   // EP_Y
              if (st.n_valid == 1) {
                if (col0 == 0 && live) ch.y[p * ch.ldy + st.y_col] = fmaf(acc[0], rz, __ldg(st.bias));
              } else if (has_next) {   // a later step still multiplies A: no scratch, direct stores
                if (live) {
                  float* dst = ch.y + p * ch.ldy + st.y_col + col0;
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (col0 + i < st.n_valid) dst[i] = fmaf(acc[i], rz, __ldg(st.bias + col0 + i));
                }
              } else {
                // coalesced fp32 row-major stores: transpose the 32 x 16 piece through this warp's scratch inside A_lo
                // (idle: every MMA of the tile has completed, and A_lo is never saved)
                float* scr = reinterpret_cast<float*>(sAlo) + ew * (32 * 17);
#pragma unroll
                for (int i = 0; i < 16; ++i) scr[lane * 17 + i] = fmaf(acc[i], rz, (col0 + i < st.n_valid) ? __ldg(st.bias + col0 + i) : 0.f);
                __syncwarp();
                const int cc = lane & 15, n = col0 + cc;
                const int64_t row0 = p - lane;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                  const int rw = 2 * k + (lane >> 4);
                  const int64_t pr = row0 + rw;
                  if (pr < ch.P && n < st.n_valid) ch.y[pr * ch.ldy + st.y_col + n] = scr[rw * 17 + cc];
                }
                __syncwarp();
              }
            }
          } else {
            if (st.epi == EP_RELU) {
              if (col0 + 16 <= st.n_valid) {
                const float4* b4 = reinterpret_cast<const float4*>(st.bias + col0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 b = __ldg(b4 + i);
                  o[4 * i + 0] = fmaxf(fmaf(acc[4 * i + 0], rz, b.x), 0.f);
                  o[4 * i + 1] = fmaxf(fmaf(acc[4 * i + 1], rz, b.y), 0.f);
                  o[4 * i + 2] = fmaxf(fmaf(acc[4 * i + 2], rz, b.z), 0.f);
                  o[4 * i + 3] = fmaxf(fmaf(acc[4 * i + 3], rz, b.w), 0.f);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = (col0 + i < st.n_valid) ? fmaxf(fmaf(acc[i], rz, __ldg(st.bias + col0 + i)), 0.f) : 0.f;
              }
            } else {   // EP_RGB
              if (live) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int n = col0 + i;
                  if (n < st.n_valid) ch.rgb[p * st.n_valid + n] = 1.0f / (1.0f + expf(-fmaf(acc[i], rz, __ldg(st.bias + n))));
                }
              }
            }
          }
          if (writes_a && c < st.next_kb) {
            if (any_save) mbar_wait(&bars->s_free[c], ((fgen >> c) & 1) ^ 1);   // the previous generation of this block has been saved
            st_row16_split(sAhi + c * kBlk, sAlo + c * kBlk, m, pc & 3, o);
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) publish(c);
          }
        }
        if (writes_a)
          for (int c = 0; c < st.next_kb; ++c) fgen ^= 1u << c;
        tc_fence_before();
      }
      // the y stores of the last step use A_lo as scratch: every warp must be done before the next prologue writes it
      if (MODE == kF2Sdf && ch.st[ch.n_steps - 1].epi == EP_Y) named_bar_sync(1, kF3NW * 32);
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // no CTA leaves (or frees its TMEM) while the peer may still arrive on its barriers
  if (warp == 2) {
    if constexpr (PAIR) tmem_dealloc_2cta(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

#undef F3_TILES

template <int MODE>
static int launch_fwd3_t(const TcChain& ch, int grid, int n_sms, cudaStream_t st) {
  // cta_group::2 pairs (clusters of two CTAs, one tile each): opt-in with SVS_F3_PAIR=1.  Measured on B200 (131 072 points,
  // tools/f3_exp.py): 0.438 ms against 0.413 ms for independent CTAs — the chain's critical path per layer is epilogue
  // (~6 k cycles) + the 24 main-term instructions that must follow it (3.1 k, tensor-bound) + hand-offs, none of which the
  // halved weight traffic shortens, and the pair runs at the pace of its slower CTA (tools/f3_trace.py).
  static const bool pair = getenv("SVS_F3_PAIR") != nullptr && getenv("SVS_F3_PAIR")[0] == '1';
  if (pair && ch.n_tiles >= 2 && n_sms >= 2) {
    static bool attr_set = false;
    if (!attr_set) {
      SVS_CUDA_OK(cudaFuncSetAttribute(tc_fwd3_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F3Cfg<MODE, true>::kSmemBytes));
      attr_set = true;
    }
    const int tile_pairs = (ch.n_tiles + 1) / 2, cta_pairs = n_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (tile_pairs < cta_pairs ? tile_pairs : cta_pairs));
    cfg.blockDim = dim3(kF3Threads);
    cfg.dynamicSmemBytes = F3Cfg<MODE, true>::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SVS_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_fwd3_kernel<MODE, true>, ch));
    return SVS_OK;
  }
  static bool attr_set1 = false;
  if (!attr_set1) {
    SVS_CUDA_OK(cudaFuncSetAttribute(tc_fwd3_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F3Cfg<MODE, false>::kSmemBytes));
    attr_set1 = true;
  }
  tc_fwd3_kernel<MODE, false><<<grid, kF3Threads, F3Cfg<MODE, false>::kSmemBytes, st>>>(ch);
  return SVS_OK;
}
static int launch_fwd3(const TcChain& ch, int grid, int n_sms, cudaStream_t st) {
  return ch.prologue == PRO_RENDER_IN ? launch_fwd3_t<kF2Render>(ch, grid, n_sms, st) : launch_fwd3_t<kF2Sdf>(ch, grid, n_sms, st);
}

}  // namespace tc
}  // namespace svs

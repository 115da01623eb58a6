// Weight gradients on tcgen05: dW[n, k] += sum_p X[p, n] * Y[p, k] with the contraction over POINTS.
//
// X (the "output side": U_l, dz_l, dy) and Y (the "input side": q_l, h_l, PE(x), rendering-net input) are the tile
// images the chain kernels saved.  A K-major SWIZZLE_128B image of a [points x columns] tile is byte-for-byte the
// MN-major SWIZZLE_128B operand of the transposed product, so the same images are fed to tcgen05.mma with the
// a_major/b_major bits set: M = 128 columns of X, N = up to 256 columns of Y, K = 16 points per instruction.
// One CTA owns one (job, split): it keeps the fp32 accumulators (up to 2 x 128 x 256 = all 512 TMEM columns) over
// all of its tiles and both (X, Y) pairs of the job (tangent-sweep pair and backward pair share dW_l), then adds its
// partial result into the fp32 gradient buffer with vector reductions (REDG.F32x4).
// The bias gradients db_l = sum_p dz_l[p, :] ride along: epilogue warp 0, idle until the accumulators are complete, adds
// up the columns of the X slices of the backward pair while the slices wait in the ring (a lane per column pair, fp32
// partial sums in registers over all tiles of the CTA), so the chains' epilogues carry no column reductions (they cost
// a quarter of the backward sweep's epilogue: shuffles + shared-memory CAS loops).
#pragma once
#include "mlp_tc.cuh"

namespace svs {
namespace tc {

constexpr int kDwStages = 6;
constexpr int kDwRows = 32;                   // points per pipeline stage
constexpr int kDwSlice = kDwRows * 128;       // 4096 bytes: 32 rows of one 64-column block
constexpr int kDwStageBytes = 8 * kDwSlice;   // up to 4 X blocks + 4 Y blocks
constexpr int kDwMaxJobs = 20;
constexpr int kDwSmem = kDwStages * kDwStageBytes + 256;

struct DwJob {
  const uint8_t* X[2];
  const uint8_t* Y[2];
  int64_t x_tile_bytes[2], y_tile_bytes[2];
  int32_t n_pairs;
  int32_t x_kb, x_blk0, n_mblk;   // X image blocks; first block used; M-blocks of 128 columns (1 or 2)
  int32_t y_kb, y_blk0, n_yblk;   // Y image blocks; first block used; blocks used (N = 64 * n_yblk)
  float* dW;                      // fp32 [.., ldw]
  int32_t ldw, n_rows, n_cols;    // valid rows (X columns) / cols (Y columns) of this job's slab
  int32_t col_shift;              // output column = n + col_shift (rendering layer 0: features sit after 15 columns)
  int32_t unit0, n_split;         // CTAs [unit0, unit0 + n_split) work on this job
  float* bias;                    // != nullptr: db[n] += sum_p X[p, n] of pair `bias_pair` (the bias gradient of the layer:
  int32_t bias_pair;              //   column sums of dz_l / dy, taken from the X slices while they sit in shared memory)
};

struct DwParams {
  DwJob job[kDwMaxJobs];
  int32_t n_jobs, n_tiles;
  const uint32_t* amax; float amax_target;   // the images carry gradients x grad_scale(amax, amax_target)
};

struct DwBars {
  uint64_t full[kDwStages], empty[kDwStages], acc_full;
  uint32_t tmem;
};

__global__ void __launch_bounds__(192, 1) tc_dw_kernel(const __grid_constant__ DwParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  DwBars* bars = reinterpret_cast<DwBars*>(smem + kDwStages * kDwStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int ji = 0;
  while (ji + 1 < prm.n_jobs && (int)blockIdx.x >= prm.job[ji + 1].unit0) ++ji;
  const DwJob& jb = prm.job[ji];
  const int split = blockIdx.x - jb.unit0;

  // zero the ring once: blocks beyond the image width are never loaded and must read as zeros
  for (int i = threadIdx.x * 16; i < kDwStages * kDwStageBytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    // a stage is free again when its MMAs have completed and, in jobs with a bias, the four summing warps have read it
    for (int i = 0; i < kDwStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], jb.bias ? 5 : 1); }
    mbar_init(&bars->acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&bars->tmem, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem;

  const int nx = 2 * jb.n_mblk;  // X block slots per stage
  int n_my_tiles = 0;
  for (int t = split; t < prm.n_tiles; t += jb.n_split) ++n_my_tiles;
  const int n_stage_total = n_my_tiles * jb.n_pairs * (kTile / kDwRows);

  if (warp == 4) {
    // ===== producer =====
    if (lane == 0) {
      uint32_t seq = 0;
      for (int t = split; t < prm.n_tiles; t += jb.n_split) {
        for (int pr = 0; pr < jb.n_pairs; ++pr) {
          const uint8_t* xt = jb.X[pr] + (size_t)t * jb.x_tile_bytes[pr];
          const uint8_t* yt = jb.Y[pr] + (size_t)t * jb.y_tile_bytes[pr];
          for (int r0 = 0; r0 < kTile / kDwRows; ++r0, ++seq) {
            // the 32-row slices of a tile may be accumulated in any order: CTAs start at different slices so that at any
            // moment their 4 KB reads spread over all (address >> 12) & 3 residues instead of marching through the same
            // one together (ncu: DRAM channels 22 % .. 45 % busy with the lock-step order)
            const int r = (r0 + split) & (kTile / kDwRows - 1);
            const int slot = seq % kDwStages;
            const uint32_t use = seq / kDwStages;
            mbar_wait(&bars->empty[slot], (use & 1) ^ 1);
            uint8_t* st = smem + slot * kDwStageBytes;
            int nload = 0;
            for (int b = 0; b < nx; ++b)
              if (jb.x_blk0 + b < jb.x_kb) ++nload;
            nload += jb.n_yblk;
            mbar_arrive_expect_tx(&bars->full[slot], (uint32_t)nload * kDwSlice);
            for (int b = 0; b < nx; ++b)
              if (jb.x_blk0 + b < jb.x_kb)
                bulk_g2s(st + b * kDwSlice, xt + (size_t)(jb.x_blk0 + b) * kBlk + (size_t)r * kDwSlice, kDwSlice,
                         &bars->full[slot]);
            for (int b = 0; b < jb.n_yblk; ++b)
              bulk_g2s(st + (4 + b) * kDwSlice, yt + (size_t)(jb.y_blk0 + b) * kBlk + (size_t)r * kDwSlice, kDwSlice,
                       &bars->full[slot]);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, 64 * jb.n_yblk, 1, 1);
      for (int seq = 0; seq < n_stage_total; ++seq) {
        const int slot = seq % kDwStages;
        const uint32_t use = seq / kDwStages;
        mbar_wait(&bars->full[slot], use & 1);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + slot * kDwStageBytes);
#pragma unroll
        for (int j = 0; j < kDwRows / 16; ++j) {
          const uint64_t db = make_smem_desc(base + 4 * kDwSlice + j * 2048, kDwSlice, 1024);
          for (int mb = 0; mb < jb.n_mblk; ++mb) {
            const uint64_t da = make_smem_desc(base + mb * 2 * kDwSlice + j * 2048, kDwSlice, 1024);
            umma_f16(tmem + mb * 256, da, db, idesc, (seq | j) != 0);
          }
        }
        umma_commit(&bars->empty[slot]);
        // jobs with a bias: the four summing warps also release the stages of the backward pair; stand in for them elsewhere
        if (jb.bias && (seq / (kTile / kDwRows)) % jb.n_pairs != jb.bias_pair) mbar_arrive_n(&bars->empty[slot], 4);
      }
      umma_commit(&bars->acc_full);
    }
  } else if (warp < 4) {
    if (jb.bias && n_stage_total > 0) {
      // ===== bias gradient: column sums of X block `warp` of every stage of the backward pair =====
      // lane = row of the 32-row slice; logical 16-byte chunk j of row r sits at chunk position j ^ (r & 7): the 8 lanes of a
      // quarter warp read 8 different positions (conflict-free) and every lane keeps static accumulator indices.  Only
      // warp 0 probes the `full` barrier (every probing warp costs the ring ~1 %: measured); it releases the other three
      // through a named barrier.
      const bool mine = warp < nx && jb.x_blk0 + warp < jb.x_kb;
      float cs[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) cs[i] = 0.f;
      const int per_tile = kTile / kDwRows;
      for (int seq = 0; seq < n_stage_total; ++seq) {
        const int slot = seq % kDwStages;
        const uint32_t use = seq / kDwStages;
        const int pr = (seq / per_tile) % jb.n_pairs;
        if (pr != jb.bias_pair) continue;   // stages of the other pair are released by the MMA commit alone (see `empty`)
        if (warp == 0) mbar_wait(&bars->full[slot], use & 1);
        named_bar_sync(1, 128);
        if (mine) {
          const uint8_t* row = smem + slot * kDwStageBytes + warp * kDwSlice + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 u = *reinterpret_cast<const uint4*>(row + ((j ^ (lane & 7)) << 4));
            float2 f;
            f = unpack_h2(u.x); cs[8 * j + 0] += f.x; cs[8 * j + 1] += f.y;
            f = unpack_h2(u.y); cs[8 * j + 2] += f.x; cs[8 * j + 3] += f.y;
            f = unpack_h2(u.z); cs[8 * j + 4] += f.x; cs[8 * j + 5] += f.y;
            f = unpack_h2(u.w); cs[8 * j + 6] += f.x; cs[8 * j + 7] += f.y;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->empty[slot]);
      }
      if (mine) {
        const float inv_gs = 1.0f / grad_scale(prm.amax, prm.amax_target);
        const int col0 = warp * 64;   // column of this job's slab
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float v = warp_sum(cs[i]);
          if (lane == (i & 31) && col0 + i < jb.n_rows) atomicAdd(jb.bias + col0 + i, v * inv_gs);
        }
      }
    }
    // ===== epilogue: TMEM -> fp32 reductions into dW =====
    if (n_stage_total > 0) {
      mbar_wait(&bars->acc_full, 0);
      tc_fence_after();
      const int N = 64 * jb.n_yblk;
      const bool vec_ok = (jb.col_shift & 3) == 0 && (jb.ldw & 3) == 0;
      const float inv_gs = 1.0f / grad_scale(prm.amax, prm.amax_target);
      for (int mb = 0; mb < jb.n_mblk; ++mb) {
        const int row = mb * 128 + warp * 32 + lane;
        float* dst = jb.dW + (size_t)row * jb.ldw + jb.col_shift;
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + mb * 256 + c0, r);
          tmem_ld_wait();
          if (row < jb.n_rows) {
            if (vec_ok) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                if (c0 + i + 3 < jb.n_cols) {
                  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + c0 + i), "f"(__uint_as_float(r[i]) * inv_gs),
                               "f"(__uint_as_float(r[i + 1]) * inv_gs), "f"(__uint_as_float(r[i + 2]) * inv_gs), "f"(__uint_as_float(r[i + 3]) * inv_gs)
                               : "memory");
                } else {
                  for (int k = 0; k < 4; ++k)
                    if (c0 + i + k < jb.n_cols) atomicAdd(dst + c0 + i + k, __uint_as_float(r[i + k]) * inv_gs);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c0 + i < jb.n_cols) atomicAdd(dst + c0 + i, __uint_as_float(r[i]) * inv_gs);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace tc
}  // namespace svs

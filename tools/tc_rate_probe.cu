// Issue-rate probe for tcgen05.mma kind::f16 (GPU box; measurement only, not part of the product or the tests):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I s-volsdf_b200/csrc -o /tmp/tc_rate tools/tc_rate_probe.cu && /tmp/tc_rate
// Every SM runs one CTA (or one CTA of a 2-CTA pair) that issues back-to-back M = 128 (M = 256 per pair) K = 16 instructions
// over 4 resident 64-column A blocks and 4 resident B blocks — the operand walk of one layer of the MLP chains — and reports
// cycles per instruction:  SS (both operands in shared memory) at N = 256 / 128, TS (A in tensor memory), cta_group::2 pairs
// (every CTA holds its own 128 rows of A and HALF of B), each with and without concurrent shared-memory stores from 16 warps
// (the epilogue's traffic).  Answers whether the chains are bound by shared-memory operand bandwidth.
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include <cuda_fp16.h>
#include "tc_common.cuh"

using namespace svs::tc;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit_2cta(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// mode 0: SS 1-CTA; 1: TS 1-CTA; 2: SS 2-CTA pair
template <int mode>
__global__ void __launch_bounds__(640, 1) rate_kernel(int N, int rounds, int bg, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  __shared__ volatile int stop;
  uint8_t* sA = smem;             // 4 x 16 KB
  uint8_t* sB = smem + 65536;     // 4 x 32 KB
  uint8_t* sScr = smem + 196608;  // 16 KB scratch for the background stores
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool pair = mode == 2;
  uint32_t rank = 0;
  if constexpr (mode == 2) rank = cluster_rank();
  for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) {
    const uint32_t h = (uint32_t)i * 2654435761u;
    const __half2 v = __floats2half2_rn(((h & 1023) - 512) * (1.f / 4096.f), (((h >> 10) & 1023) - 512) * (1.f / 4096.f));
    reinterpret_cast<__half2*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
    stop = 0;
  }
  if (warp == 0) {
    if constexpr (mode == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(&tmem_s, 512);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if constexpr (mode == 2) cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const int M = pair ? 256 : 128;
      const uint32_t idesc = make_idesc_f16(M, N, 0, 0);
      const long long t0 = clock64();
      uint32_t par = 0;
      for (int r = 0; r < rounds; ++r) {
        for (int kb = 0; kb < 4; ++kb) {
          const uint32_t a0 = smem_u32(sA + kb * 16384), b0 = smem_u32(sB + kb * 32768);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t db = make_smem_desc(b0 + j * 32, 0, 1024);
            if constexpr (mode == 0) umma_f16(tmem, make_smem_desc(a0 + j * 32, 0, 1024), db, idesc, (r | kb | j) != 0);
            else if constexpr (mode == 1) umma_f16_ts(tmem, tmem + 256 + (uint32_t)(kb * 32 + j * 8), db, idesc, (r | kb | j) != 0);
            else umma_f16_2cta(tmem, make_smem_desc(a0 + j * 32, 0, 1024), db, idesc, (r | kb | j) != 0);
          }
        }
        if ((r & 7) == 7 || r == rounds - 1) {   // one commit + wait per 128 instructions
          if constexpr (mode == 2) commit_2cta(&bar, 3); else umma_commit(&bar);
          mbar_wait(&bar, par);
          par ^= 1;
        }
      }
      const long long t1 = clock64();
      out[blockIdx.x] = (float)(t1 - t0) / (float)(rounds * 16);
      stop = 1;
    } else if (lane == 0 && rank == 1) {
      uint32_t par = 0;
      for (int r = 0; r < rounds; ++r)
        if ((r & 7) == 7 || r == rounds - 1) { mbar_wait(&bar, par); par ^= 1; }
      out[blockIdx.x] = -1.f;
      stop = 1;
    }
  } else if (warp >= 4 && bg) {
    // epilogue-like traffic: every warp stores 16 B per lane (512 B per instruction) in a loop, with a little arithmetic
    uint4 v = make_uint4(lane, warp, 3, 4);
    int it = 0;
    while (!stop) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        *reinterpret_cast<uint4*>(sScr + ((warp - 4) & 7) * 2048 + ((k & 3) * 512) + lane * 16) = v;
        v.x += v.y * 3u + k;
      }
      if (bg > 1) __nanosleep(bg);
      ++it;
    }
    if (v.x == 0x12345678u && it == -1) out[0] = 1.f;
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (mode == 2) cluster_sync();
  if (warp == 0) {
    if constexpr (mode == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else tmem_dealloc(tmem, 512);
  }
}

template <int mode>
static void run(const char* name, int N, int bg) {
  const int grid = 148, rounds = 64, smem = 196608 + 16384;
  float* d_out;
  CK(cudaMalloc(&d_out, grid * sizeof(float)));
  CK(cudaFuncSetAttribute(rate_kernel<mode>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(640);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = mode == 2 ? 2 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = mode == 2 ? 1 : 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<mode>, N, rounds, bg, d_out);
    if (e != cudaSuccess) { printf("%s N=%d: launch failed: %s\n", name, N, cudaGetErrorString(e)); cudaGetLastError(); cudaFree(d_out); return; }
    CK(cudaDeviceSynchronize());
  }
  std::vector<float> h(grid);
  CK(cudaMemcpy(h.data(), d_out, grid * sizeof(float), cudaMemcpyDeviceToHost));
  std::vector<float> v;
  for (float x : h) if (x > 0) v.push_back(x);
  std::sort(v.begin(), v.end());
  const double flop = 2.0 * (mode == 2 ? 256 : 128) * N * 16;
  printf("%-34s N=%3d bg=%3d : cycles / instruction min %.1f median %.1f max %.1f   (ideal %.0f; %.0f%% of the tensor peak)\n", name, N, bg,
         v.front(), v[v.size() / 2], v.back(), flop / 8192.0 / (mode == 2 ? 2 : 1), 100.0 * (flop / 8192.0 / (mode == 2 ? 2 : 1)) / v[v.size() / 2]);
  CK(cudaFree(d_out));
}

int main() {
  for (int bg = 0; bg <= 1; ++bg) {
    run<0>("SS  cta_group::1 M=128", 256, bg);
    run<0>("SS  cta_group::1 M=128", 128, bg);
    run<0>("SS  cta_group::1 M=128", 64, bg);
    run<1>("TS  cta_group::1 M=128 (A in TMEM)", 256, bg);
    run<1>("TS  cta_group::1 M=128 (A in TMEM)", 128, bg);
    run<2>("SS  cta_group::2 M=256", 256, bg);
    run<2>("SS  cta_group::2 M=256", 128, bg);
  }
  run<0>("SS  cta_group::1 M=128", 256, 200);
  run<2>("SS  cta_group::2 M=256", 256, 200);
  return 0;
}

// Bring-up probe for the tcgen05 building blocks in s-volsdf_b200/csrc/tc_common.cuh (run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I s-volsdf_b200/csrc -o /tmp/tc_probe tools/tc_probe.cu && /tmp/tc_probe
// Checks, against a host reference, one CTA doing
//   test 0: C[128,N]  = X[128,K] * W[N,K]^T      both operands K-major SWIZZLE_128B tile images
//   test 1: D[128,N2] = X[P,0:128]^T * Y[P,0:N2]  both operands MN-major views of the same kind of image
// Not part of the product or the tests.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include <cuda_fp16.h>
#include "tc_common.cuh"

using namespace svs::tc;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

// mode 0: K-major.  A image: 128 rows x K cols; B image: N rows x K cols.  C = A * B^T  (128 x N)
// mode 1: MN-major. A image: P rows x 128 cols (uses cols [0,128)); B image: P rows x N cols. D = A^T * B (128 x N)
__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* __restrict__ a_img, int a_bytes,
                                                   const uint8_t* __restrict__ b_img, int b_bytes, int mode, int afmt, int bfmt, int N,
                                                   int K, int a_rows, int b_rows, float* __restrict__ C) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a_bytes + 1023) / 1024) * 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_load, (uint32_t)(a_bytes + b_bytes));
    bulk_g2s(sa, a_img, (uint32_t)a_bytes, &bar_load);
    bulk_g2s(sb, b_img, (uint32_t)b_bytes, &bar_load);
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    if (mode == 0) {
      const uint32_t idesc = (make_idesc_bf16(128, N, 0, 0) & ~((7u << 7) | (7u << 10))) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10);
      for (int j = 0; j < K / 16; ++j) {
        uint64_t da = make_smem_desc(smem_u32(sa) + (j >> 2) * (a_rows * 128) + (j & 3) * 32, 0, 1024);
        uint64_t db = make_smem_desc(smem_u32(sb) + (j >> 2) * (b_rows * 128) + (j & 3) * 32, 0, 1024);
        umma_f16(tmem, da, db, idesc, j > 0);
      }
    } else {
      const uint32_t idesc = (make_idesc_bf16(128, N, 1, 1) & ~((7u << 7) | (7u << 10))) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10);
      for (int j = 0; j < K / 16; ++j) {  // K = number of points (rows of the images)
        uint64_t da = make_smem_desc(smem_u32(sa) + j * 2048, a_rows * 128, 1024);
        uint64_t db = make_smem_desc(smem_u32(sb) + j * 2048, b_rows * 128, 1024);
        umma_f16(tmem, da, db, idesc, j > 0);
      }
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 32 && c0 + i < N; ++i) C[row * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

static void build_image(const std::vector<float>& m, int rows, int cols, std::vector<uint8_t>& img, int fmt) {
  img.assign((size_t)rows * cols * 2, 0);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      if (fmt == 1) {
        __nv_bfloat16 v = __float2bfloat16(m[(size_t)r * cols + c]);
        memcpy(&img[img_off(r, c, rows)], &v, 2);
      } else {
        __half v = __float2half(m[(size_t)r * cols + c]);
        memcpy(&img[img_off(r, c, rows)], &v, 2);
      }
    }
}

static int run(int mode, int N, int K, int afmt = 1, int bfmt = 1) {
  // mode 0: A 128 x K, B N x K.   mode 1: A K(points) x 128, B K(points) x N
  int a_rows = mode == 0 ? 128 : K, a_cols = mode == 0 ? K : 128;
  int b_rows = mode == 0 ? N : K, b_cols = mode == 0 ? K : ((N + 63) / 64) * 64;
  std::vector<float> A((size_t)a_rows * a_cols), B((size_t)b_rows * b_cols, 0.f);
  srand(1 + mode * 7 + N);
  for (auto& v : A) v = bf((rand() % 2001 - 1000) / 1000.f);
  for (int r = 0; r < b_rows; ++r)
    for (int c = 0; c < (mode == 0 ? K : N); ++c) B[(size_t)r * b_cols + c] = bf((rand() % 2001 - 1000) / 1000.f);
  std::vector<uint8_t> ai, bi;
  build_image(A, a_rows, a_cols, ai, afmt);
  build_image(B, b_rows, b_cols, bi, bfmt);
  uint8_t *da, *db;
  float* dc;
  CK(cudaMalloc(&da, ai.size()));
  CK(cudaMalloc(&db, bi.size()));
  CK(cudaMalloc(&dc, 128 * N * 4));
  CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dc, 0, 128 * N * 4));
  int smem = (int)(((ai.size() + 1023) / 1024) * 1024 + bi.size() + 1024);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem>>>(da, (int)ai.size(), db, (int)bi.size(), mode, afmt, bfmt, N, K, a_rows, b_rows, dc);
  CK(cudaDeviceSynchronize());
  std::vector<float> C(128 * N);
  CK(cudaMemcpy(C.data(), dc, C.size() * 4, cudaMemcpyDeviceToHost));
  double worst = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      if (mode == 0)
        for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
      else
        for (int p = 0; p < K; ++p) ref += (double)A[(size_t)p * 128 + m] * B[(size_t)p * b_cols + n];
      worst = fmax(worst, fabs(ref - C[m * N + n]));
    }
  printf("mode %d fmt %d%d N %3d K %3d : max abs err %.3e  %s\n", mode, afmt, bfmt, N, K, worst, worst < 1e-3 ? "OK" : "FAIL");
  cudaFree(da);
  cudaFree(db);
  cudaFree(dc);
  return worst < 1e-3 ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run(0, 256, 64);
  bad += run(0, 256, 256);
  bad += run(0, 128, 128);
  bad += run(0, 16, 256);
  bad += run(0, 64, 320);
  bad += run(1, 256, 128);
  bad += run(1, 128, 64);
  bad += run(1, 64, 128);
  bad += run(0, 256, 256, 0, 0);
  bad += run(0, 256, 256, 0, 1);
  bad += run(0, 256, 256, 1, 0);
  bad += run(1, 256, 128, 0, 1);
  bad += run(1, 256, 128, 1, 0);
  printf(bad ? "PROBE FAILED (%d)\n" : "PROBE OK\n", bad);
  return bad;
}

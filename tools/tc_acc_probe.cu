// Probe of the tcgen05.mma kind::f16 accumulator arithmetic (fp32 in TMEM): rounding mode and alignment precision of
// D += A * B.  Run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I s-volsdf_b200/csrc -o /tmp/tc_acc_probe tools/tc_acc_probe.cu && /tmp/tc_acc_probe
// Not part of the product or the tests.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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

// C = A[128,K] * B[N,K]^T, K-major fp16 images, one MMA instruction per 16 columns of K, in order
__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* __restrict__ a_img, int a_bytes,
                                                   const uint8_t* __restrict__ b_img, int b_bytes, int N, int K,
                                                   float* __restrict__ C) {
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
    const uint32_t idesc = make_idesc_f16(128, N, 0, 0);
    for (int j = 0; j < K / 16; ++j) {
      uint64_t da = make_smem_desc(smem_u32(sa) + (j >> 2) * (128 * 128) + (j & 3) * 32, 0, 1024);
      uint64_t db = make_smem_desc(smem_u32(sb) + (j >> 2) * (N * 128) + (j & 3) * 32, 0, 1024);
      umma_f16(tmem, da, db, idesc, j > 0);
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

static void build_image(const std::vector<float>& m, int rows, int cols, std::vector<uint8_t>& img) {
  img.assign((size_t)rows * cols * 2, 0);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      __half v = __float2half(m[(size_t)r * cols + c]);
      memcpy(&img[img_off(r, c, rows)], &v, 2);
    }
}

static std::vector<float> run(const std::vector<float>& A, const std::vector<float>& B, int N, int K) {
  std::vector<uint8_t> ai, bi;
  build_image(A, 128, K, ai);
  build_image(B, N, K, bi);
  uint8_t *da, *db;
  float* dc;
  CK(cudaMalloc(&da, ai.size()));
  CK(cudaMalloc(&db, bi.size()));
  CK(cudaMalloc(&dc, 128 * N * 4));
  CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
  int smem = (int)(((ai.size() + 1023) / 1024) * 1024 + bi.size() + 1024);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem>>>(da, (int)ai.size(), db, (int)bi.size(), N, K, dc);
  CK(cudaDeviceSynchronize());
  std::vector<float> C(128 * N);
  CK(cudaMemcpy(C.data(), dc, C.size() * 4, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(dc);
  return C;
}

static float h(float x) { return __half2float(__float2half(x)); }

int main() {
  const int N = 16, K = 256;
  // ---- case studies: row m of A selects the experiment, column n = 0 of B --------------------------------------
  // k-step 0 sets acc = base (A[m,0] * B[0,0] = base * 1); later k-steps add deltas
  {
    std::vector<float> A(128 * K, 0.f), B(N * K, 0.f);
    B[0] = 1.f;
    // row 0: +0.75 ulp per k-step (15 steps):  RN -> 1 + 15*2^-23 ; RZ -> 1
    // row 1: +0.25 ulp per k-step            :  RN -> 1            ; RZ -> 1 ; exact wide -> 1 + 3.75 ulp
    // row 2: one k-step with 16 products of +0.25 ulp each (sum 4 ulp): exact-sum-then-round -> 1 + 4 ulp
    // row 3: base = -1, +0.75 ulp(of 1)/step : RN -> ... ; RZ (toward zero) -> each step moves toward zero
    // row 4: base = 1, -0.25 ulp per step    : RN -> 1 ; RZ toward zero -> 1 - 2^-24 per step (ulp below 1 is 2^-24: -0.25ulp = -0.5 ulp_below)
    // row 5: one k-step: products +1 ulp and 15 x +0.0625 ulp (sum 1.9375 ulp)
    // row 6: base 1, one k-step: 8 x (+2^-30), to probe alignment depth (sum 2^-27 = 1/16 ulp)  then 15 such steps
    for (int m = 0; m < 8; ++m) A[m * K + 0] = (m == 3) ? -1.f : 1.f;
    for (int j = 1; j < 16; ++j) {
      B[j * 16] = ldexpf(1.f, -12);            // b = 2^-12 at the first k of every k-step
      A[0 * K + j * 16] = 1.5f * ldexpf(1.f, -12);    // 1.5 * 2^-24 = 0.75 ulp(1)
      A[1 * K + j * 16] = 0.5f * ldexpf(1.f, -12);    // 0.25 ulp
      A[3 * K + j * 16] = 1.5f * ldexpf(1.f, -12);
      A[4 * K + j * 16] = -0.5f * ldexpf(1.f, -12);
    }
    for (int k = 16; k < 32; ++k) {             // k-step 1: all 16 products
      B[k] = ldexpf(1.f, -12);
      A[2 * K + k] = 0.5f * ldexpf(1.f, -12);   // 16 x 0.25 ulp
      A[5 * K + k] = (k == 16) ? 2.f * ldexpf(1.f, -12) : 0.125f * ldexpf(1.f, -12);
    }
    // rows 0,1,3,4 must not see B[17..31]: they only have A at k = 16 j -> fine (A zero elsewhere)
    std::vector<float> C = run(A, B, N, K);
    const double ulp = ldexp(1.0, -23);
    printf("row0 (+0.75ulp x15): (C-1)/ulp = %.4f   [RN: 15, RZ: 0]\n", (C[0 * N] - 1.0) / ulp);
    printf("row1 (+0.25ulp x15): (C-1)/ulp = %.4f   [RN/RZ per step: 0, exact: 3.75]\n", (C[1 * N] - 1.0) / ulp);
    printf("row2 (16 x 0.25ulp in one instr): (C-1)/ulp = %.4f   [sum-then-round: 4]\n", (C[2 * N] - 1.0) / ulp);
    printf("row3 (base -1, +0.75ulp x15): (C+1)/ulp = %.4f   [RZ(toward 0): each step lands on the next fp32 toward 0]\n", (C[3 * N] + 1.0) / ulp);
    printf("row4 (base 1, -0.25ulp x15): (C-1)/ulp = %.4f   [RN: 0 ; RZ: -0.5 per step = -7.5]\n", (C[4 * N] - 1.0) / ulp);
    printf("row5 (1ulp + 15 x 0.0625ulp in one instr): (C-1)/ulp = %.4f   [exact 1.9375]\n", (C[5 * N] - 1.0) / ulp);
  }
  // ---- statistics on random data: bias and rms of (C - exact) in ulps, K = 256 (16 accumulations) ----------------
  for (int pass = 0; pass < 2; ++pass) {
    const int Nn = 256;
    std::vector<float> A(128 * K), B(Nn * K);
    srand(7 + pass);
    for (auto& v : A) v = h((rand() % 20001 - 10000) / 10000.f * (pass ? 1.f : 0.3f));
    for (auto& v : B) v = h((rand() % 20001 - 10000) / 10000.f * 0.1f);
    if (pass == 1) for (auto& v : A) v = fabsf(v);   // positive activations (softplus-like): partial sums random-walk anyway
    std::vector<float> C = run(A, B, Nn, K);
    double bias = 0, rms = 0, rel_bias = 0;
    int cnt = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < Nn; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
        if (fabs(ref) < 1e-3) continue;
        int e;
        frexp(ref, &e);
        double ulp = ldexp(1.0, e - 24);
        double d = (C[m * Nn + n] - ref) / ulp * (ref > 0 ? 1 : -1);   // > 0: magnitude too large
        bias += d; rms += d * d; rel_bias += (C[m * Nn + n] - ref) / ref; ++cnt;
      }
    printf("random pass %d: mean signed error %.3f ulp (negative = toward zero), rms %.3f ulp, mean rel %.3e   [RN chain: mean ~0, rms ~1-2]\n",
           pass, bias / cnt, sqrt(rms / cnt), rel_bias / cnt);
  }
  return 0;
}

// fp32 SIMT tile GEMMs with fused epilogues: the parity engine (SVS_ENGINE_FP32) of the MLP kernels.
//
// Three shapes cover the whole forward / reverse-sweep / tangent / backward chain of SURVEY.md App. F:
//   gemm<NT>  C[M,N] = A[M,K] . W[N,K]^T   forward layers, tangent sweep        (W = weight, row-major)
//   gemm<NN>  C[M,N] = A[M,K] . W[K,N]     reverse sweep, backward data         (same W, read as K x N)
//   gemm_tn   dW[N,K] += A1[M,N]^T . A2[M,K]   weight gradients, split over M with fp32 atomics
// 128x128x16 tiles, 256 threads, 8x8 register micro-tile, register-prefetch double buffering.
#pragma once
#include "svs_common.cuh"

namespace svs {

constexpr float kSoftplusBeta = 100.0f;  // nn.Softplus(beta=100) (network.py:69)
constexpr float kSoftplusThresh = 20.0f; // torch default threshold

__device__ __forceinline__ float softplus100(float z) {
  float bz = kSoftplusBeta * z;
  return (bz > kSoftplusThresh) ? z : log1pf(expf(bz)) / kSoftplusBeta;
}
// sigma'(z) recovered from h = softplus(z): s = 1 - exp(-beta h)  (exactly 1 in the threshold branch)
__device__ __forceinline__ float dsoftplus_from_h(float h) { return -expm1f(-kSoftplusBeta * h); }

enum Epi {
  EPI_BIAS = 0,           // C = acc + bias
  EPI_BIAS_SOFTPLUS = 1,  // C = softplus(acc + bias) * scale
  EPI_BIAS_RELU = 2,      // C = relu(acc + bias)
  EPI_BIAS_SIGMOID = 3,   // C = sigmoid(acc + bias)
  EPI_PLAIN = 4,          // C = acc
  EPI_REVERSE = 5,        // p = acc*scale; n < n_split: C = s(E1*hscale) * p ; n >= n_split: C2[n-n_split] = p
  EPI_TANGENT = 6,        // r = acc; s = s(E1*hscale); C = beta (1-s) E2 r ; C2 = s r scale
  EPI_BACKWARD = 7,       // C = s(E1*hscale) * acc * scale + (E2 ? E2 : 0)
  EPI_RELU_BWD = 8        // C = E1 > 0 ? acc : 0
};

struct GemmArgs {
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;
  int M, N, K;
  const float* bias;
  const float* E1; int lde1;
  const float* E2; int lde2;
  float* C2; int ldc2;
  float scale, hscale;
  int n_split;
};

constexpr int GBM = 128, GBN = 128, GBK = 16, GPAD = 4;

__device__ __forceinline__ float4 ld4_masked(const float* row, int k, int K) {
  // row points at element 0 of a row whose allocation is padded to a multiple of 4 floats
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k < K) {
    v = *reinterpret_cast<const float4*>(row + k);
    if (k + 3 >= K) {
      if (k + 1 >= K) v.y = 0.f;
      if (k + 2 >= K) v.z = 0.f;
      if (k + 3 >= K) v.w = 0.f;
    }
  }
  return v;
}

template <int EPI>
__device__ __forceinline__ void epilogue_elem(const GemmArgs& a, int m, int n, float acc) {
  if (EPI == EPI_BIAS) {
    a.C[(int64_t)m * a.ldc + n] = acc + a.bias[n];
  } else if (EPI == EPI_BIAS_SOFTPLUS) {
    a.C[(int64_t)m * a.ldc + n] = softplus100(acc + a.bias[n]) * a.scale;
  } else if (EPI == EPI_BIAS_RELU) {
    a.C[(int64_t)m * a.ldc + n] = fmaxf(acc + a.bias[n], 0.f);
  } else if (EPI == EPI_BIAS_SIGMOID) {
    a.C[(int64_t)m * a.ldc + n] = 1.0f / (1.0f + expf(-(acc + a.bias[n])));
  } else if (EPI == EPI_PLAIN) {
    a.C[(int64_t)m * a.ldc + n] = acc;
  } else if (EPI == EPI_REVERSE) {
    float p = acc * a.scale;
    if (n < a.n_split) {
      float s = dsoftplus_from_h(a.E1[(int64_t)m * a.lde1 + n] * a.hscale);
      a.C[(int64_t)m * a.ldc + n] = s * p;
    } else {
      a.C2[(int64_t)m * a.ldc2 + (n - a.n_split)] = p;
    }
  } else if (EPI == EPI_TANGENT) {
    float s = dsoftplus_from_h(a.E1[(int64_t)m * a.lde1 + n] * a.hscale);
    float u = a.E2[(int64_t)m * a.lde2 + n];
    a.C[(int64_t)m * a.ldc + n] = kSoftplusBeta * (1.0f - s) * u * acc;
    a.C2[(int64_t)m * a.ldc2 + n] = s * acc * a.scale;
  } else if (EPI == EPI_BACKWARD) {
    float s = dsoftplus_from_h(a.E1[(int64_t)m * a.lde1 + n] * a.hscale);
    float v = s * acc * a.scale;
    if (a.E2) v += a.E2[(int64_t)m * a.lde2 + n];
    a.C[(int64_t)m * a.ldc + n] = v;
  } else if (EPI == EPI_RELU_BWD) {
    a.C[(int64_t)m * a.ldc + n] = (a.E1[(int64_t)m * a.lde1 + n] > 0.f) ? acc : 0.f;
  }
}

// TRANS_B = true : B is W[N,K] row-major (C = A W^T);  false: B is W[K,N] row-major (C = A W)
template <bool TRANS_B, int EPI>
__global__ void __launch_bounds__(256, 2) gemm_kernel(const GemmArgs a) {
  __shared__ float As[2][GBK][GBM + GPAD];
  __shared__ float Bs[2][GBK][GBN + GPAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
  const int ty = tid >> 4, tx = tid & 15;

  // loader coordinates
  const int a_row = tid >> 2, a_kq = (tid & 3) * 4;   // rows a_row, a_row+64 ; k offset a_kq
  const int b_krow = tid >> 5, b_nq = (tid & 31) * 4; // NN: k rows b_krow, b_krow+8 ; n offset b_nq

  float4 ra[2], rb[2];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int m = m0 + a_row + 64 * i;
      ra[i] = (m < a.M) ? ld4_masked(a.A + (int64_t)m * a.lda, k0 + a_kq, a.K) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (TRANS_B) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int n = n0 + a_row + 64 * i;
        rb[i] = (n < a.N) ? ld4_masked(a.B + (int64_t)n * a.ldb, k0 + a_kq, a.K) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int k = k0 + b_krow + 8 * i;
        rb[i] = (k < a.K) ? ld4_masked(a.B + (int64_t)k * a.ldb, n0 + b_nq, a.N) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = a_row + 64 * i;
      As[buf][a_kq + 0][r] = ra[i].x;
      As[buf][a_kq + 1][r] = ra[i].y;
      As[buf][a_kq + 2][r] = ra[i].z;
      As[buf][a_kq + 3][r] = ra[i].w;
    }
    if (TRANS_B) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int r = a_row + 64 * i;
        Bs[buf][a_kq + 0][r] = rb[i].x;
        Bs[buf][a_kq + 1][r] = rb[i].y;
        Bs[buf][a_kq + 2][r] = rb[i].z;
        Bs[buf][a_kq + 3][r] = rb[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 2; ++i)
        *reinterpret_cast<float4*>(&Bs[buf][b_krow + 8 * i][b_nq]) = rb[i];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = (a.K + GBK - 1) / GBK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * GBK);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4));
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + ((j < 4) ? (tx * 4 + j) : (64 + tx * 4 + j - 4));
      if (n < a.N) epilogue_elem<EPI>(a, m, n, acc[i][j]);
    }
  }
}

// dW[N,K] (+)= A1[M,N]^T A2[M,K] over rows [blockIdx.y*rows, +rows)
__global__ void __launch_bounds__(256, 2)
gemm_tn_kernel(const float* __restrict__ A1, int lda1, const float* __restrict__ A2, int lda2,
               float* __restrict__ dW, int ldw, int M, int N, int K, int tiles_k, int rows) {
  __shared__ float As[2][GBK][GBM + GPAD];
  __shared__ float Bs[2][GBK][GBN + GPAD];
  const int tid = threadIdx.x;
  const int tn = blockIdx.x / tiles_k, tk = blockIdx.x % tiles_k;
  const int n0 = tn * GBM, k0 = tk * GBN;
  const int mbeg = blockIdx.y * rows;
  const int mend = min(M, mbeg + rows);
  const int ty = tid >> 4, tx = tid & 15;
  const int l_row = tid >> 5, l_q = (tid & 31) * 4;

  float4 ra[2], rb[2];
  auto load_tiles = [&](int mm) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int m = mm + l_row + 8 * i;
      bool ok = m < mend;
      ra[i] = ok ? ld4_masked(A1 + (int64_t)m * lda1, n0 + l_q, N) : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[i] = ok ? ld4_masked(A2 + (int64_t)m * lda2, k0 + l_q, K) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      *reinterpret_cast<float4*>(&As[buf][l_row + 8 * i][l_q]) = ra[i];
      *reinterpret_cast<float4*>(&Bs[buf][l_row + 8 * i][l_q]) = rb[i];
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int nm = (mend - mbeg + GBK - 1) / GBK;
  if (nm <= 0) return;
  load_tiles(mbeg);
  store_tiles(0);
  __syncthreads();
  for (int t = 0; t < nm; ++t) {
    const int buf = t & 1;
    if (t + 1 < nm) load_tiles(mbeg + (t + 1) * GBK);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < nm) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int n = n0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4));
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int k = k0 + ((j < 4) ? (tx * 4 + j) : (64 + tx * 4 + j - 4));
      if (k < K) atomicAdd(&dW[(int64_t)n * ldw + k], acc[i][j]);
    }
  }
}

// out[n] += sum_m A[m, n]   (bias gradients, and the tangent-sweep term of the last layer's row 0)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ A, int lda, int M, int N, int rows, float* __restrict__ out) {
  int n = blockIdx.x * 32 + (threadIdx.x & 31);
  int sub = threadIdx.x >> 5;  // 8 row-slices per block
  int mbeg = blockIdx.y * rows, mend = min(M, mbeg + rows);
  float acc = 0.f;
  if (n < N)
    for (int m = mbeg + sub; m < mend; m += 8) acc += A[(int64_t)m * lda + n];
  __shared__ float red[8][33];
  red[sub][threadIdx.x & 31] = acc;
  __syncthreads();
  if (sub == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int s = 0; s < 8; ++s) t += red[s][threadIdx.x & 31];
    atomicAdd(&out[n], t);
  }
}

template <bool TRANS_B, int EPI>
static int launch_gemm(const GemmArgs& a, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0) return SVS_OK;
  dim3 grid((unsigned)cdiv(a.M, GBM), (unsigned)cdiv(a.N, GBN));
  ProfScope ps(TRANS_B ? "mlp_gemm_nt_fp32" : "mlp_gemm_nn_fp32", 2.0 * a.M * a.N * a.K,
               4.0 * ((double)a.M * a.K + (double)a.N * a.K + (double)a.M * a.N), st);
  gemm_kernel<TRANS_B, EPI><<<grid, 256, 0, st>>>(a);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

static int launch_gemm_tn(const float* A1, int lda1, const float* A2, int lda2, float* dW, int ldw, int64_t M,
                          int N, int K, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return SVS_OK;
  int tiles_n = (int)cdiv(N, GBM), tiles_k = (int)cdiv(K, GBN);
  int tiles = tiles_n * tiles_k;
  int want_chunks = (int)cdiv(4 * kNumSMs, tiles);  // ~2 waves at 2 CTAs/SM
  int64_t rows = round_up(cdiv(M, want_chunks), GBK);
  if (rows < 256) rows = 256;
  int chunks = (int)cdiv(M, rows);
  dim3 grid((unsigned)tiles, (unsigned)chunks);
  ProfScope ps("mlp_gemm_tn_fp32", 2.0 * M * N * K, 4.0 * ((double)M * N + (double)M * K + (double)N * K), st);
  gemm_tn_kernel<<<grid, 256, 0, st>>>(A1, lda1, A2, lda2, dW, ldw, (int)M, N, K, tiles_k, (int)rows);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

static int launch_colsum(const float* A, int lda, int64_t M, int N, float* out, cudaStream_t st) {
  if (M <= 0 || N <= 0) return SVS_OK;
  int64_t rows = round_up(cdiv(M, 4 * kNumSMs / (int)cdiv(N, 32) + 1), 8);
  if (rows < 64) rows = 64;
  dim3 grid((unsigned)cdiv(N, 32), (unsigned)cdiv(M, rows));
  colsum_kernel<<<grid, 256, 0, st>>>(A, lda, (int)M, N, (int)rows, out);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

}  // namespace svs

// Optimiser step of the reference's loop as two launches: gradient clipping by global L2 norm + non-finite guard +
// Adam over all parameter tensors (volsdf/vsdf.py:214-219,454-464: zero_grad / backward / clip_grad_norm_(1.0) /
// on_after_backward NaN guard / Adam.step).  PyTorch issues ~25 multi-tensor launches (plus one host sync per
// parameter in the reference's NaN guard) for the 43 small tensors of the DTU model; here every tensor is walked by
// one grid.  Semantics = torch.optim.Adam (no amsgrad, no weight decay) and torch.nn.utils.clip_grad_norm_.
#include "svs_common.cuh"

namespace svs {

struct OptTensors {
  float* p[SVS_OPT_MAX_TENSORS];
  const float* g[SVS_OPT_MAX_TENSORS];
  float* m[SVS_OPT_MAX_TENSORS];
  float* v[SVS_OPT_MAX_TENSORS];
  long long n[SVS_OPT_MAX_TENSORS];
  int count;
};

constexpr int kOptBlock = 256, kOptPerBlock = 4096;

// scratch[0] += sum g^2 ; scratch[1] = 1 if any gradient entry is not finite
__global__ void __launch_bounds__(kOptBlock) grad_norm_kernel(const OptTensors t, float* __restrict__ scratch) {
  const int ti = blockIdx.y;
  const long long n = t.n[ti];
  const long long base = (long long)blockIdx.x * kOptPerBlock;
  if (base >= n) return;
  const float* g = t.g[ti];
  float ss = 0.f;
  bool bad = false;
  for (long long i = base + threadIdx.x; i < n && i < base + kOptPerBlock; i += kOptBlock) {
    float x = g[i];
    bad |= !isfinite(x);
    ss += x * x;
  }
  ss = warp_sum(ss);
  bad = __any_sync(0xffffffffu, bad);
  __shared__ float red[kOptBlock / 32];
  __shared__ int redbad;
  if (threadIdx.x == 0) redbad = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = ss;
    if (bad) redbad = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < kOptBlock / 32; ++i) tot += red[i];
    atomicAdd(&scratch[0], tot);
    if (redbad) scratch[1] = 1.f;
  }
}

// *step_count = number of this step (device scalar, already incremented by the caller: torch keeps `step` as a tensor)
__global__ void __launch_bounds__(kOptBlock)
adam_kernel(const OptTensors t, const float* __restrict__ scratch, const float* __restrict__ step_count, float lr,
            float beta1, float beta2, float eps, float max_norm, int skip_nonfinite) {
  const int ti = blockIdx.y;
  const long long n = t.n[ti];
  const long long base = (long long)blockIdx.x * kOptPerBlock;
  const float step = *step_count;
  if (base >= n) return;
  const bool bad = skip_nonfinite && scratch[1] != 0.f;   // vsdf.py:454-464: gradients are zeroed, Adam still steps
  float clip = 1.0f;
  if (max_norm > 0.f) {
    const float norm = sqrtf(scratch[0]);
    clip = fminf(max_norm / (norm + 1e-6f), 1.0f);   // clip_grad_norm_: coef clamped to 1
  }
  const float bc1 = 1.0f - powf(beta1, step), bc2 = 1.0f - powf(beta2, step);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  float* p = t.p[ti];
  const float* g = t.g[ti];
  float* m = t.m[ti];
  float* v = t.v[ti];
  for (long long i = base + threadIdx.x; i < n && i < base + kOptPerBlock; i += kOptBlock) {
    const float gi = bad ? 0.f : g[i] * clip;
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

}  // namespace svs

using namespace svs;

extern "C" int svs_adam_step(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                             float* const* exp_avg_sq, const int64_t* numel, float lr, float beta1, float beta2,
                             float eps, float max_norm, int32_t skip_nonfinite, const float* step_count, float* scratch,
                             void* stream) {
  SVS_CHECK_ARG(n_tensors > 0 && n_tensors <= SVS_OPT_MAX_TENSORS, "svs_adam_step: %d tensors (max %d)", n_tensors,
                SVS_OPT_MAX_TENSORS);
  SVS_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && numel && step_count && scratch, "svs_adam_step: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  OptTensors t;
  memset(&t, 0, sizeof(t));
  t.count = n_tensors;
  long long nmax = 0;
  for (int i = 0; i < n_tensors; ++i) {
    SVS_CHECK_ARG(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i] && numel[i] >= 0, "svs_adam_step: tensor %d", i);
    t.p[i] = params[i]; t.g[i] = grads[i]; t.m[i] = exp_avg[i]; t.v[i] = exp_avg_sq[i]; t.n[i] = numel[i];
    if (numel[i] > nmax) nmax = numel[i];
  }
  if (nmax == 0) return SVS_OK;
  dim3 grid((unsigned)cdiv(nmax, kOptPerBlock), (unsigned)n_tensors);
  SVS_CUDA_OK(cudaMemsetAsync(scratch, 0, 2 * sizeof(float), st));
  grad_norm_kernel<<<grid, kOptBlock, 0, st>>>(t, scratch);
  SVS_LAUNCH_OK();
  adam_kernel<<<grid, kOptBlock, 0, st>>>(t, scratch, step_count, lr, beta1, beta2, eps, max_norm, skip_nonfinite);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

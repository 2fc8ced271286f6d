// Optimiser glue of the training step: global gradient norm, clip (train_rec.py:148,
// torch.nn.utils.clip_grad_norm_) and Adam (torch.optim.Adam defaults, train_detection.py:378,
// train_rec.py:381-382) over ONE flat fp32 parameter / gradient buffer per model.
#include "common.cuh"
#include <math.h>

namespace {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ partials) {
  __shared__ float red[32];
  float acc = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 v = g4[i];
    acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
    acc = fmaf(g[i], g[i], acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// norm_out[0] = grad_scale * sqrt(sum partials)  (the norm of the averaged gradient)
__global__ void norm_finalize_kernel(const float* __restrict__ partials, int nblk, float grad_scale,
                                     float* __restrict__ norm_out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) s += (double)partials[i];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x != 0) return;
  s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  norm_out[0] = (float)(sqrt(s) * (double)grad_scale);
}

// p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps), with g <- g * grad_scale * clip_coef first.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float bc1,
            float bc2_sqrt, float grad_scale, float max_norm, const float* __restrict__ norm) {
  float coef = grad_scale;
  if (max_norm > 0.f) coef *= fminf(max_norm / (norm[0] + 1e-6f), 1.f);
  const float step = lr / bc1;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float gi = g[i] * coef;
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] -= step * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

// Same update with the step count read from device memory (CUDA-graph friendly: the bias corrections change every
// replay without re-recording the launch). step_dev holds the number of steps already taken.
__global__ void __launch_bounds__(256)
adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                long long n, float lr, float b1, float b2, float eps, const int* __restrict__ step_dev,
                float grad_scale, float max_norm, const float* __restrict__ norm) {
  const float t = (float)(step_dev[0] + 1);
  const float bc1 = 1.f - powf(b1, t), bc2_sqrt = sqrtf(1.f - powf(b2, t));
  float coef = grad_scale;
  if (max_norm > 0.f) coef *= fminf(max_norm / (norm[0] + 1e-6f), 1.f);
  const float step = lr / bc1;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float gi = g[i] * coef;
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] -= step * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}
__global__ void step_inc_kernel(int* step_dev) { step_dev[0] += 1; }

}  // namespace

extern "C" {

int ocrs_optim_blocks(void) { return 4 * OCRS_NUM_SMS; }

// norm_out[0] = grad_scale * ||g||_2 ; partials: float[ocrs_optim_blocks()]
int ocrs_grad_norm(const float* g, long long n, float grad_scale, float* partials, float* norm_out,
                   void* stream) {
  OCRS_CHECK_ARG(((uintptr_t)g % 16) == 0, "grad_norm: buffer must be 16-byte aligned");
  const int nb = ocrs_optim_blocks();
  sumsq_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(g, n, partials);
  norm_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, nb, grad_scale, norm_out);
  OCRS_CHECK_LAUNCH_N("grad_norm", 2);
  return 0;
}

// One Adam step over a flat buffer. step_t >= 1. max_norm <= 0 disables clipping (norm may be NULL).
int ocrs_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1,
                   float b2, float eps, int step_t, float grad_scale, float max_norm, const float* norm,
                   void* stream) {
  OCRS_CHECK_ARG(step_t >= 1, "adam_step: step must be >= 1");
  OCRS_CHECK_ARG(max_norm <= 0.f || norm != nullptr, "adam_step: clipping needs the gradient norm");
  const float bc1 = 1.f - powf(b1, (float)step_t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, (float)step_t));
  adam_kernel<<<ocrs_optim_blocks(), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, b1, b2, eps, bc1,
                                                                    bc2_sqrt, grad_scale, max_norm, norm);
  OCRS_CHECK_LAUNCH("adam_kernel");
  return 0;
}

// ocrs_adam_step with the step counter in device memory: uses step_dev[0] + 1 as the step and increments it afterwards
// (so a captured CUDA graph of the training step stays valid across replays).
int ocrs_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                       int* step_dev, float grad_scale, float max_norm, const float* norm, void* stream) {
  OCRS_CHECK_ARG(step_dev != nullptr, "adam_step_dev: needs the device step counter");
  OCRS_CHECK_ARG(max_norm <= 0.f || norm != nullptr, "adam_step_dev: clipping needs the gradient norm");
  adam_dev_kernel<<<ocrs_optim_blocks(), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, b1, b2, eps, step_dev,
                                                                        grad_scale, max_norm, norm);
  step_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
  OCRS_CHECK_LAUNCH_N("adam_dev_kernel", 2);
  return 0;
}

}  // extern "C"

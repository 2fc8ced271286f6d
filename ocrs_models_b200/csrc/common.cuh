// Shared helpers for the sm_100a kernels behind the C-ABI in include/ocrs_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define OCRS_NUM_SMS 148

// Error plumbing: every extern "C" entry point returns 0 or a non-zero code and leaves a
// thread-local message readable through ocrs_last_error().
void ocrs_set_error(const char* fmt, ...);

#define OCRS_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ocrs_set_error(__VA_ARGS__);                \
      return -1;                                  \
    }                                             \
  } while (0)

void ocrs_count_launches(int n);
#define OCRS_CHECK_LAUNCH(name) OCRS_CHECK_LAUNCH_N(name, 1)
#define OCRS_CHECK_LAUNCH_N(name, n)                                         \
  do {                                                                       \
    ocrs_count_launches(n);                                                  \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      ocrs_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return (int)e__;                                                       \
    }                                                                        \
  } while (0)

#define OCRS_CUDA(call)                                                          \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      ocrs_set_error("%s failed: %s", #call, cudaGetErrorString(e__));            \
      return (int)e__;                                                           \
    }                                                                            \
  } while (0)

static inline int ocrs_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device/context: do it once per (kernel, device), not once
// per process (wrong with several GPUs in one process) and not on every call (microseconds of host time per launch
// add up: the recognition step has ~280 launches). `flags` is a per-call-site array of 64 relaxed atomics.
#define OCRS_SET_SMEM_ONCE(kernel, bytes)                                                                \
  do {                                                                                                   \
    static unsigned char flags__[64] = {0};                                                              \
    int dev__ = 0;                                                                                       \
    cudaGetDevice(&dev__);                                                                               \
    dev__ &= 63;                                                                                         \
    if (!__atomic_load_n(&flags__[dev__], __ATOMIC_ACQUIRE)) {                                           \
      OCRS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      __atomic_store_n(&flags__[dev__], (unsigned char)1, __ATOMIC_RELEASE);                             \
    }                                                                                                    \
  } while (0)

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of a float held by every thread; result valid in thread 0 (and broadcast
// when `bcast`). `red` is a >=32-float shared scratch.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) v = warp_sum(v);
  return v;
}

// Per-channel on-load transform: v -> max(v * scale + shift, lo). Folds the producer block's
// BatchNorm affine (+ReLU when lo == 0) into the consumer's load. Identity = {1, 0, -inf}.
struct ChanXform {
  const float* scale;  // [C] or nullptr (identity)
  const float* shift;  // [C]
  const float* lo;     // [C] lower clamp (0 for ReLU, -inf for none)
};
__device__ __forceinline__ float xform_apply(float v, float sc, float sh, float lo) {
  return fmaxf(fmaf(v, sc, sh), lo);
}
#endif

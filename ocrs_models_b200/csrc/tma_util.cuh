// TMA / mbarrier helpers shared by the sm_100a kernels, plus a process-lifetime cache of CUtensorMap
// descriptors keyed by (pointer, geometry): PyTorch's caching allocator hands the same addresses back every
// step, so after the first step no descriptor is encoded on the host again (SURVEY 8b "ownership").
#pragma once
#include "common.cuh"
#include <cuda.h>

// Encode (or fetch from the cache) a tiled fp32 tensor map of rank `rank` (<= 4).
// dims[i] elements, strides[i] = byte stride of dimension i+1, box[i] elements; no swizzle unless `swizzle128`.
// Returns 0 and fills *out, or non-zero with ocrs_last_error() set.
int ocrs_get_tensor_map(CUtensorMap* out, const void* ptr, int rank, const unsigned long long* dims,
                        const unsigned long long* strides_bytes, const unsigned* box, int swizzle128);

// Planar NCHW fp32 view (base, sample stride) -> 4-D map (W, H, C, N) with box (bw, bh, bc, 1).
// Out-of-bounds elements (halo outside the image, channels past C) are zero-filled by the hardware.
int ocrs_plane_map(CUtensorMap* out, const float* base, long long sample_stride, int N, int C, int H, int W, int bw,
                   int bh, int bc);

// 1 when a planar view can be addressed by TMA: 16-byte aligned base, W and the sample stride multiples of 4.
static inline int ocrs_plane_tma_ok(const float* base, long long ss, int H, int W) {
  return ((uintptr_t)base % 16 == 0) && (W % 4 == 0) && (ss % 4 == 0) && (((long long)H * W) % 4 == 0);
}

#ifdef __CUDACC__
namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void load_4d(uint32_t dst, const CUtensorMap* map, int x, int y, int c, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "r"(c), "r"(n) : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
}  // namespace tma
#endif

// Probe: which 4-D planar tensor-map / box shapes does UTMALDG accept? (build: nvcc ... tma_probe.cu ../../ocrs_models_b200/csrc/api.o)
#include "../../ocrs_models_b200/csrc/tma_util.cuh"
#include <stdio.h>
#include <vector>
__global__ void k(const __grid_constant__ CUtensorMap m, int x, int y, int c, int n, int bytes, float* out, int nout) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { tma::mbar_init(tma::smem_u32(&bar), 1); tma::fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { tma::mbar_expect_tx(tma::smem_u32(&bar), bytes); tma::load_4d(tma::smem_u32(smem), &m, x, y, c, n, tma::smem_u32(&bar)); }
  tma::mbar_wait(tma::smem_u32(&bar), 0);
  for (int i = threadIdx.x; i < nout; i += blockDim.x) out[i] = ((float*)smem)[i];
}
extern "C" const char* ocrs_last_error(void);
int run(int N, int C, int H, int W, int bw, int bh, int bc, int x, int y, int c, int n) {
  std::vector<float> h((size_t)N * C * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int nout = bw * bh * bc;
  cudaMalloc(&o, nout * 4);
  CUtensorMap m;
  if (ocrs_plane_map(&m, d, (long long)C * H * W, N, C, H, W, bw, bh, bc)) { printf("encode failed: %s\n", ocrs_last_error()); return 1; }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  k<<<1, 128, nout * 4 + 128>>>(m, x, y, c, n, nout * 4, o, nout);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> r(nout);
  cudaMemcpy(r.data(), o, nout * 4, cudaMemcpyDeviceToHost);
  // check element (0,1,1) of the box = global (n, c, y+1, x+1)
  float want = (float)((((size_t)n * C + c) * H + (y + 1)) * W + (x + 1));
  printf("N%d C%d H%d W%d box %dx%dx%d at (%d,%d,%d,%d): %s  got %.0f want %.0f\n", N, C, H, W, bw, bh, bc, x, y, c, n,
         cudaGetErrorString(e), r[bw + 1], want);
  cudaFree(d); cudaFree(o);
  return e != cudaSuccess;
}
int main(int argc, char** argv) {
  int which = argc > 1 ? atoi(argv[1]) : 0;
  switch (which) {
    case 0: return run(2, 8, 64, 64, 40, 34, 4, -4, -1, 0, 0);
    case 1: return run(2, 8, 64, 64, 40, 34, 4, 28, 31, 4, 1);
    case 2: return run(2, 8, 64, 64, 40, 34, 4, 0, -1, 0, 0);
    case 3: return run(2, 8, 64, 64, 40, 34, 2, 28, 31, 7, 1);
    case 4: return run(1, 1, 40, 36, 40, 34, 2, -4, -1, 0, 0);
  }
  return 0;
}

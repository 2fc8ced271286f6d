// Eval-path post-processing on the device: binarisation + 8-connected component labelling of the predicted text mask
// and compaction of the components' boundary pixels. Replaces the full-mask device->host copy and the
// cv2.findContours pass of reference ocrs_models/postprocess.py:11-36 (extract_cc_quads, called per image from
// train_detection.py:176-184 after binarize_mask, train_detection.py:33-34); the minimum-area rectangle of each
// component is then computed on the host from its (few) boundary points - it depends only on their convex hull,
// which is the hull of the external contour findContours returns.
//
// Labelling = lock-free union-find over pixel indices (each pixel starts as its own root; every foreground pixel is
// united with its E, SW, S, SE foreground neighbours by atomicMin on the roots; a final pass flattens the trees).
// The label of a component is 1 + the smallest linear pixel index in it, so the result does not depend on the
// order in which the atomics land (deterministic), background is 0.
#include "common.cuh"

namespace {

__device__ __forceinline__ int uf_find(const int* L, int i) {
  int p = L[i];
  while (p != i) { i = p; p = L[i]; }
  return i;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  while (true) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&L[b], a);  // hang the larger root under the smaller one
    if (old == b) return;
    b = old;                              // somebody re-parented b meanwhile: retry from its new parent
  }
}

// mask > threshold -> foreground. labels[n][i] = i for foreground, -1 for background.
__global__ void cc_init_kernel(const float* __restrict__ mask, float threshold, long long HW, int* __restrict__ labels) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (i >= HW) return;
  labels[(size_t)n * HW + i] = mask[(size_t)n * HW + i] > threshold ? (int)i : -1;
}

__global__ void cc_merge_kernel(int* __restrict__ labels, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  int* L = labels + (size_t)blockIdx.z * H * W;
  const int i = y * W + x;
  if (L[i] < 0) return;
  if (x + 1 < W && L[i + 1] >= 0) uf_union(L, i, i + 1);
  if (y + 1 < H) {
    if (x > 0 && L[i + W - 1] >= 0) uf_union(L, i, i + W - 1);
    if (L[i + W] >= 0) uf_union(L, i, i + W);
    if (x + 1 < W && L[i + W + 1] >= 0) uf_union(L, i, i + W + 1);
  }
}

// Flatten: out[i] = root(i) + 1 (0 for background); count components (roots) and boundary pixels per image.
__global__ void cc_flatten_kernel(const int* __restrict__ labels, int H, int W, int* __restrict__ out,
                                  int* __restrict__ n_components) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const size_t base = (size_t)blockIdx.z * H * W;
  const int i = y * W + x;
  int r = 0;
  if (labels[base + i] >= 0) {
    const int root = uf_find(labels + base, i);
    r = root + 1;
    if (root == i) atomicAdd(&n_components[blockIdx.z], 1);
  }
  out[base + i] = r;
}

// Boundary pixels (foreground with a background / out-of-image 4-neighbour) -> packed (label, x, y) triples.
__global__ void cc_boundary_kernel(const int* __restrict__ lab, int H, int W, int* __restrict__ points, int capacity,
                                   int* __restrict__ n_points) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const size_t base = (size_t)blockIdx.z * H * W;
  const int i = y * W + x;
  const int l = lab[base + i];
  if (l == 0) return;
  const bool edge = x == 0 || y == 0 || x == W - 1 || y == H - 1 || lab[base + i - 1] == 0 || lab[base + i + 1] == 0 ||
                    lab[base + i - W] == 0 || lab[base + i + W] == 0;
  if (!edge) return;
  const int k = atomicAdd(&n_points[blockIdx.z], 1);
  if (k < capacity) {
    int* p = points + ((size_t)blockIdx.z * capacity + k) * 3;
    p[0] = l; p[1] = x; p[2] = y;
  }
}

}  // namespace

extern "C" {

// 8-connected components of (mask > threshold) for N images [N][H][W] (fp32 probabilities or 0/1 masks):
// labels [N][H][W] int32 out (0 = background, else 1 + smallest pixel index of the component; deterministic),
// scratch [N][H][W] int32 workspace, n_components [N] int32 out (zero-filled by the caller).
int ocrs_cc_label(const float* mask, float threshold, int N, int H, int W, int* scratch, int* labels, int* n_components,
                  void* stream) {
  OCRS_CHECK_ARG(N > 0 && H > 0 && W > 0 && (long long)H * W < 2147483647LL, "cc_label: bad dims");
  cudaStream_t st = (cudaStream_t)stream;
  const long long HW = (long long)H * W;
  cc_init_kernel<<<dim3(ocrs_cdiv(HW, 256), N), 256, 0, st>>>(mask, threshold, HW, scratch);
  dim3 block(32, 8), grid(ocrs_cdiv(W, 32), ocrs_cdiv(H, 8), N);
  cc_merge_kernel<<<grid, block, 0, st>>>(scratch, H, W);
  cc_flatten_kernel<<<grid, block, 0, st>>>(scratch, H, W, labels, n_components);
  OCRS_CHECK_LAUNCH_N("cc_label", 3);
  return 0;
}

// Boundary pixels of every component as (label, x, y) int32 triples: points [N][capacity][3], n_points [N]
// (zero-filled by the caller; may exceed capacity, in which case the caller retries with a larger buffer).
int ocrs_cc_boundary(const int* labels, int N, int H, int W, int* points, int capacity, int* n_points, void* stream) {
  OCRS_CHECK_ARG(N > 0 && H > 0 && W > 0 && capacity > 0, "cc_boundary: bad dims");
  dim3 block(32, 8), grid(ocrs_cdiv(W, 32), ocrs_cdiv(H, 8), N);
  cc_boundary_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(labels, H, W, points, capacity, n_points);
  OCRS_CHECK_LAUNCH("cc_boundary_kernel");
  return 0;
}

}  // extern "C"

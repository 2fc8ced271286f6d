// Multi-tensor glue of the training step: ONE launch where the step used to issue one small kernel per parameter.
//
//  * ocrs_grad_deliver: every parameter gradient of a backward pass leaves the kernels as per-CTA / split-K partial rows
//    (or as a finished row). This kernel reduces each of them in double in a fixed order (what ocrs_finalize_partials did
//    per tensor), re-lays convolution gradients out from the GEMM's [Cout][(ky,kx,ci)] to the parameter's
//    [Cout][Cin][kh][kw], and either stores the result or ACCUMULATES it into the destination - the `.grad` view of the
//    flat gradient bucket of optim.FusedAdam - which replaces autograd's per-parameter `grad += g` kernels
//    (reference: loss.backward() at ocrs_models/train_rec.py:130 / train_detection.py:96 followed by optimizer.step()).
//  * ocrs_weight_prep: the per-step weight operands of the recognition GEMMs (reference models.py:179-243: Conv2d weights
//    as [Cout][(ky,kx,ci)] for the forward implicit GEMM and flipped [Cin][(ky,kx,co)] for the data gradient, every GEMM
//    weight split into TF32 hi/lo, W_hh transposed for the persistent GRU backward) from the parameters in one launch.
//
// The work lists travel as kernel parameters (by value), so a captured CUDA graph replays them without host tables.
#include "common.cuh"

namespace {

constexpr int GD_MAX = 72;  // entries per launch (44 bytes each: fits the 4 KB parameter space)

struct GradTable {
  const float* src[GD_MAX];
  float* dst[GD_MAX];
  int K[GD_MAX];       // elements of the destination
  int rows[GD_MAX];    // partial rows to sum
  int ld[GD_MAX];      // floats between partial rows
  int flags[GD_MAX];   // bits 0-1: column map (0 identity, 1 conv re-layout, 2 inner/pitch), bit 2: accumulate, bit 3: wide
  int p0[GD_MAX];      // map 1: Cin      map 2: inner run length
  int p1[GD_MAX];      // map 1: kh*kw    map 2: pitch of a run in the source
  int blk0[GD_MAX];    // first block of the entry
  int count;
};

// Work item k of an entry -> (source column, destination index). The convolution re-layout walks the SOURCE in order
// (coalesced partial-row reads, one scattered store per element); the other maps walk the destination.
__device__ __forceinline__ void gd_index(int k, int map, int p0, int p1, int& col, int& di) {
  col = k; di = k;
  if (map == 1) {  // src [co][t][ci] -> dst [co][ci][t]
    const int ci = k % p0, r = k / p0, t = r % p1, co = r / p1;
    di = (co * p0 + ci) * p1 + t;
  } else if (map == 2) {
    col = (k / p0) * p1 + (k % p0);
  }
}

// "wide" entries (few rows, many columns): 256 threads x 4 consecutive columns, rows summed by the thread.
// other entries (many rows): 8 row lanes x 32 columns, four loads in flight per thread, shared-memory reduce in double.
__global__ void __launch_bounds__(256) grad_deliver_kernel(const __grid_constant__ GradTable tab) {
  __shared__ double red[8][33];
  int e = 0;
  while (e + 1 < tab.count && (int)blockIdx.x >= tab.blk0[e + 1]) ++e;
  const int b = blockIdx.x - tab.blk0[e];
  const float* __restrict__ src = tab.src[e];
  float* __restrict__ dst = tab.dst[e];
  const int K = tab.K[e], rows = tab.rows[e], ld = tab.ld[e], fl = tab.flags[e], map = fl & 3;
  const bool acc = fl & 4;
  const int p0 = tab.p0[e], p1 = tab.p1[e];
  if (fl & 8) {
    const int k0 = (b * 256 + threadIdx.x) * 4;
    if (k0 >= K) return;
    if (map == 0 && k0 + 4 <= K && ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) && (ld & 3) == 0) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int r = 0;
      for (; r + 3 < rows; r += 4) {  // four rows in flight
        const float4 a = *reinterpret_cast<const float4*>(src + (size_t)r * ld + k0);
        const float4 c = *reinterpret_cast<const float4*>(src + (size_t)(r + 1) * ld + k0);
        const float4 d = *reinterpret_cast<const float4*>(src + (size_t)(r + 2) * ld + k0);
        const float4 f = *reinterpret_cast<const float4*>(src + (size_t)(r + 3) * ld + k0);
        s0 += ((double)a.x + (double)c.x) + ((double)d.x + (double)f.x);
        s1 += ((double)a.y + (double)c.y) + ((double)d.y + (double)f.y);
        s2 += ((double)a.z + (double)c.z) + ((double)d.z + (double)f.z);
        s3 += ((double)a.w + (double)c.w) + ((double)d.w + (double)f.w);
      }
      for (; r < rows; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(src + (size_t)r * ld + k0);
        s0 += a.x; s1 += a.y; s2 += a.z; s3 += a.w;
      }
      float4 o = make_float4((float)s0, (float)s1, (float)s2, (float)s3);
      if (acc) {
        const float4 d = *reinterpret_cast<const float4*>(dst + k0);
        o.x = d.x + o.x; o.y = d.y + o.y; o.z = d.z + o.z; o.w = d.w + o.w;
      }
      *reinterpret_cast<float4*>(dst + k0) = o;
      return;
    }
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    int c[4], di[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) gd_index(min(k0 + q, K - 1), map, p0, p1, c[q], di[q]);
    for (int r = 0; r < rows; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) s[q] += (double)src[(size_t)r * ld + c[q]];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (k0 + q < K) dst[di[q]] = acc ? dst[di[q]] + (float)s[q] : (float)s[q];
    return;
  }
  const int kx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int k = b * 32 + kx;
  double s = 0.0;
  int c = 0, di = 0;
  if (k < K) {
    gd_index(k, map, p0, p1, c, di);
    int r = ry;
    for (; r + 24 < rows; r += 32) {
      const float v0 = src[(size_t)r * ld + c], v1 = src[(size_t)(r + 8) * ld + c];
      const float v2 = src[(size_t)(r + 16) * ld + c], v3 = src[(size_t)(r + 24) * ld + c];
      s += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
    }
    for (; r < rows; r += 8) s += (double)src[(size_t)r * ld + c];
  }
  red[ry][kx] = s;
  __syncthreads();
  if (ry == 0 && k < K) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][kx];
    dst[di] = acc ? dst[di] + (float)t : (float)t;
  }
}

constexpr int WP_MAX = 40;

struct PrepTable {
  const float* src[WP_MAX];
  float* hi[WP_MAX];     // identity / forward layout, TF32 hi (or the plain copy when lo is null)
  float* lo[WP_MAX];
  float* dhi[WP_MAX];    // data-gradient layout (convolutions only), may be null
  float* dlo[WP_MAX];
  int n[WP_MAX];         // elements
  int mode[WP_MAX];      // 0: same layout, 1: conv [Cout][Cin][kh][kw], 2: transpose [R][C] -> [C][R]
  int d0[WP_MAX];        // mode 1: Cout   mode 2: R
  int d1[WP_MAX];        // mode 1: Cin    mode 2: C
  int kh[WP_MAX], kw[WP_MAX];
  int blk0[WP_MAX];
  int count;
};

__device__ __forceinline__ void split_store(float v, float* hi, float* lo, size_t i) {
  if (lo) {
    const float h = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);  // round to TF32 (ties away), as ocrs_split_tf32
    hi[i] = h;
    lo[i] = v - h;
  } else {
    hi[i] = v;
  }
}

__global__ void __launch_bounds__(256) weight_prep_kernel(const __grid_constant__ PrepTable tab) {
  int e = 0;
  while (e + 1 < tab.count && (int)blockIdx.x >= tab.blk0[e + 1]) ++e;
  const int i = (blockIdx.x - tab.blk0[e]) * 256 + threadIdx.x;
  if (i >= tab.n[e]) return;
  const float v = tab.src[e][i];
  const int mode = tab.mode[e];
  if (mode == 0) {
    split_store(v, tab.hi[e], tab.lo[e], i);
  } else if (mode == 2) {
    const int C = tab.d1[e], r = i / C, c = i - r * C;
    split_store(v, tab.hi[e], tab.lo[e], (size_t)c * tab.d0[e] + r);
  } else {
    const int kh = tab.kh[e], kw = tab.kw[e], Cin = tab.d1[e], Cout = tab.d0[e];
    const int kx = i % kw, r1 = i / kw, ky = r1 % kh, r2 = r1 / kh, ci = r2 % Cin, co = r2 / Cin;
    split_store(v, tab.hi[e], tab.lo[e], ((size_t)co * kh * kw + ky * kw + kx) * Cin + ci);
    if (tab.dhi[e])  // flipped kernel, channels swapped: the correlation that yields the data gradient
      split_store(v, tab.dhi[e], tab.dlo[e], ((size_t)ci * kh * kw + (kh - 1 - ky) * kw + (kw - 1 - kx)) * Cout + co);
  }
}

}  // namespace

extern "C" {

// Entries one ocrs_grad_deliver / ocrs_weight_prep call accepts (callers chunk longer lists).
int ocrs_grad_deliver_max(void) { return GD_MAX; }
int ocrs_weight_prep_max(void) { return WP_MAX; }

// dst[e][k] (+)= sum_{r < rows[e]} src[e][r * ld[e] + col(k)] for `count` gradient tensors in one launch; sums in double
// in a fixed order (deterministic). All arrays are HOST arrays of length count; the tensors they point to are device memory.
//   map[e] = 0: col(k) = k
//   map[e] = 1: convolution gradient: dst [Cout][Cin][kh*kw] from src rows laid out [Cout][kh*kw][Cin]; p0 = Cin, p1 = kh*kw
//   map[e] = 2: col(k) = (k / p0) * p1 + k % p0 (runs of p0 elements at pitch p1)
//   accumulate[e] != 0: dst += (the `.grad += g` of autograd's AccumulateGrad), else dst = .
int ocrs_grad_deliver(const float* const* src, float* const* dst, const int* K, const int* rows, const int* ld, const int* map,
                      const int* p0, const int* p1, const int* accumulate, int count, void* stream) {
  OCRS_CHECK_ARG(count >= 0 && count <= GD_MAX, "grad_deliver: %d entries (max %d per call)", count, GD_MAX);
  if (count == 0) return 0;
  GradTable t;
  int blocks = 0;
  for (int e = 0; e < count; ++e) {
    OCRS_CHECK_ARG(src[e] && dst[e] && K[e] > 0 && rows[e] > 0 && map[e] >= 0 && map[e] <= 2, "grad_deliver: bad entry %d", e);
    OCRS_CHECK_ARG(map[e] == 0 || (p0[e] > 0 && p1[e] > 0), "grad_deliver: bad column map of entry %d", e);
    // many columns: one thread per 4 columns walks the rows (split-K partials of the big GEMM gradients have up to ~100 rows);
    // few columns and many rows (per-CTA partials): 8 row lanes per column
    const bool wide = rows[e] < 16 || (K[e] >= 8192 && rows[e] <= 256);
    t.src[e] = src[e]; t.dst[e] = dst[e]; t.K[e] = K[e]; t.rows[e] = rows[e]; t.ld[e] = ld[e];
    t.flags[e] = map[e] | (accumulate[e] ? 4 : 0) | (wide ? 8 : 0);
    t.p0[e] = p0[e]; t.p1[e] = p1[e];
    t.blk0[e] = blocks;
    blocks += wide ? ocrs_cdiv(K[e], 1024) : ocrs_cdiv(K[e], 32);
  }
  t.count = count;
  grad_deliver_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(t);
  OCRS_CHECK_LAUNCH("grad_deliver_kernel");
  return 0;
}

// Per-step weight operands in one launch. Host arrays of length count:
//   mode 0: hi/lo = TF32 split of src in the same layout (lo null: plain copy)
//   mode 1: src = Conv2d weight [d0 = Cout][d1 = Cin][kh][kw] -> hi/lo as [Cout][(ky,kx,ci)] and, when dhi is given,
//           dhi/dlo as [Cin][(kh-1-ky, kw-1-kx, co)]
//   mode 2: src [d0][d1] -> hi/lo transposed [d1][d0]
int ocrs_weight_prep(const float* const* src, float* const* hi, float* const* lo, float* const* dhi, float* const* dlo,
                     const int* n, const int* mode, const int* d0, const int* d1, const int* kh, const int* kw, int count,
                     void* stream) {
  OCRS_CHECK_ARG(count >= 0 && count <= WP_MAX, "weight_prep: %d entries (max %d per call)", count, WP_MAX);
  if (count == 0) return 0;
  PrepTable t;
  int blocks = 0;
  for (int e = 0; e < count; ++e) {
    OCRS_CHECK_ARG(src[e] && hi[e] && n[e] > 0 && mode[e] >= 0 && mode[e] <= 2, "weight_prep: bad entry %d", e);
    OCRS_CHECK_ARG(mode[e] != 1 || ((long long)d0[e] * d1[e] * kh[e] * kw[e] == n[e] && (!dhi[e] || !lo[e] || dlo[e])),
                   "weight_prep: entry %d: bad convolution geometry", e);
    OCRS_CHECK_ARG(mode[e] != 2 || (long long)d0[e] * d1[e] == n[e], "weight_prep: entry %d: bad transpose geometry", e);
    t.src[e] = src[e]; t.hi[e] = hi[e]; t.lo[e] = lo[e]; t.dhi[e] = dhi[e]; t.dlo[e] = dlo[e];
    t.n[e] = n[e]; t.mode[e] = mode[e]; t.d0[e] = d0[e]; t.d1[e] = d1[e]; t.kh[e] = kh[e]; t.kw[e] = kw[e];
    t.blk0[e] = blocks;
    blocks += ocrs_cdiv(n[e], 256);
  }
  t.count = count;
  weight_prep_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(t);
  OCRS_CHECK_LAUNCH("weight_prep_kernel");
  return 0;
}

}  // extern "C"

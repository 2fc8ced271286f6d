// Backward kernels of the text-detection U-Net (autograd of reference ocrs_models/models.py:7-143).
//
// Per DepthwiseConv block (a = relu(bn(y)), y = pw(dw(xact))), given d_a:
//   1. bnrelu_bwd_reduce  : per-channel  sum(dz), sum(dz * yhat)          (dz = d_a * [a > 0])
//   2. bn_bwd_finalize    : d_gamma, d_beta and the affine  dy = k1*dz + k2*y + k3
//   3. pwT_bwd            : g[ci]   = sum_co Wpw[co][ci] * dy[co]
//   4. pw_wgrad           : dWpw[co][ci] = sum_p dy[co][p] * dwout[ci][p]   (dwout recomputed)
//   5. dw_bwd             : d_xact = dw3x3^T(g),  dWdw[ci][k] = sum_p g[ci][p] * xact[ci][p+k]
// Weight-gradient reductions over pixels are "skinny GEMMs" (M,N <= 16 per block, K = pixels) on warp-level
// mma.sync 3xTF32; per-block partial results are reduced in double by finalize_partials (deterministic).
// These kernels serve the shapes TMA cannot address (W % 4 != 0); csrc/det_tma.cu holds the main path.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace {

// ---------------------------------------------------------------------------------------------
// out[k] = sum_b partials[b][k], accumulated in double in a fixed order (deterministic).
// RL row lanes x 32 columns per block; four independent loads in flight per thread.
template <int RL>
__global__ void __launch_bounds__(32 * RL)
finalize_partials_kernel(const float* __restrict__ partials, int nblk, int K, float* __restrict__ out) {
  __shared__ double red[RL][33];
  const int kx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + kx;
  double s = 0.0;
  if (k < K) {
    int b = ry;
    for (; b + 3 * RL < nblk; b += 4 * RL) {
      const float v0 = partials[(size_t)b * K + k], v1 = partials[(size_t)(b + RL) * K + k];
      const float v2 = partials[(size_t)(b + 2 * RL) * K + k], v3 = partials[(size_t)(b + 3 * RL) * K + k];
      s += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
    }
    for (; b < nblk; b += RL) s += (double)partials[(size_t)b * K + k];
  }
  red[ry][kx] = s;
  __syncthreads();
  if (ry == 0 && k < K) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < RL; ++q) t += red[q][kx];
    out[k] = (float)t;
  }
}

// Few rows, many columns (split-K GEMM partials): one thread per 4 columns, 16-byte loads, 4 rows in flight.
__global__ void __launch_bounds__(128)
finalize_partials4_kernel(const float* __restrict__ partials, int nblk, int K, float* __restrict__ out) {
  const int k = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (k >= K) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int b = 0;
  for (; b + 3 < nblk; b += 4) {
    const float4 a = *reinterpret_cast<const float4*>(partials + (size_t)b * K + k);
    const float4 c = *reinterpret_cast<const float4*>(partials + (size_t)(b + 1) * K + k);
    const float4 d = *reinterpret_cast<const float4*>(partials + (size_t)(b + 2) * K + k);
    const float4 e = *reinterpret_cast<const float4*>(partials + (size_t)(b + 3) * K + k);
    s0 += ((double)a.x + (double)c.x) + ((double)d.x + (double)e.x);
    s1 += ((double)a.y + (double)c.y) + ((double)d.y + (double)e.y);
    s2 += ((double)a.z + (double)c.z) + ((double)d.z + (double)e.z);
    s3 += ((double)a.w + (double)c.w) + ((double)d.w + (double)e.w);
  }
  for (; b < nblk; ++b) {
    const float4 a = *reinterpret_cast<const float4*>(partials + (size_t)b * K + k);
    s0 += a.x; s1 += a.y; s2 += a.z; s3 += a.w;
  }
  *reinterpret_cast<float4*>(out + k) = make_float4((float)s0, (float)s1, (float)s2, (float)s3);
}

// ---------------------------------------------------------------------------------------------
// out_conv backward: dz = dp * p * (1 - p); d_a[c] = w[c] * dz; partial sums of dz * a[c] and dz.
__global__ void __launch_bounds__(256)
outconv_bwd_kernel(const float* __restrict__ dp, const float* __restrict__ prob,
                   const float* __restrict__ x, long long x_ss, int C, long long HW,
                   const float* __restrict__ sc, const float* __restrict__ sh,
                   const float* __restrict__ lo, const float* __restrict__ w,
                   float* __restrict__ d_a, long long da_ss, float* __restrict__ partials) {
  __shared__ float red[32];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  const bool ok = i < HW;
  float dz = 0.f;
  if (ok) {
    const float p = prob[(size_t)n * HW + i];
    dz = dp[(size_t)n * HW + i] * p * (1.f - p);
  }
  const size_t blk = (size_t)n * gridDim.x + blockIdx.x;
  for (int c = 0; c < C; ++c) {
    float a = 0.f;
    if (ok) {
      a = x[(size_t)n * x_ss + (size_t)c * HW + i];
      if (sc) a = xform_apply(a, sc[c], sh[c], lo[c]);
      d_a[(size_t)n * da_ss + (size_t)c * HW + i] = w[c] * dz;
    }
    const float s = block_sum(a * dz, red);
    if (threadIdx.x == 0) partials[blk * (C + 1) + c] = s;
  }
  const float s = block_sum(dz, red);
  if (threadIdx.x == 0) partials[blk * (C + 1) + C] = s;
}

// ---------------------------------------------------------------------------------------------
// BatchNorm+ReLU backward, reduction half. grid = (chunks, C, N); partials [N*chunks][2][C].
constexpr int RED_CHUNK = 4096;
__global__ void __launch_bounds__(256)
bnrelu_bwd_reduce_kernel(const float* __restrict__ d_a, long long da_ss,
                         const float* __restrict__ y, long long y_ss, int C, long long HW,
                         const float* __restrict__ sc, const float* __restrict__ sh,
                         const float* __restrict__ lo, const float* __restrict__ mean,
                         const float* __restrict__ invstd, float* __restrict__ partials) {
  __shared__ float red[32];
  const int c = blockIdx.y, n = blockIdx.z;
  const float s = sc[c], t = sh[c], l = lo[c], mu = mean[c], is = invstd[c];
  const float* dp = d_a + (size_t)n * da_ss + (size_t)c * HW;
  const float* yp = y + (size_t)n * y_ss + (size_t)c * HW;
  const long long i0 = (long long)blockIdx.x * RED_CHUNK;
  const long long i1 = min(i0 + RED_CHUNK, HW);
  float a = 0.f, b = 0.f;
  if (i0 + RED_CHUNK <= HW && (((uintptr_t)(dp + i0) | (uintptr_t)(yp + i0)) & 15) == 0) {
    // full, 16-byte aligned chunk: 4 + 4 independent 16-byte loads per thread in flight (the scalar loop ran at 2.4 TB/s)
    float4 yv[4], dv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + 4 * (threadIdx.x + 256 * u);
      yv[u] = *reinterpret_cast<const float4*>(yp + i);
      dv[u] = *reinterpret_cast<const float4*>(dp + i);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float ys[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w}, ds[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float dz = (fmaf(ys[q], s, t) > l) ? ds[q] : 0.f;
        a += dz;
        b = fmaf(dz, (ys[q] - mu) * is, b);
      }
    }
  } else {
    for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
      const float yv = yp[i];
      const float dz = (fmaf(yv, s, t) > l) ? dp[i] : 0.f;
      a += dz;
      b = fmaf(dz, (yv - mu) * is, b);
    }
  }
  a = block_sum(a, red);
  b = block_sum(b, red);
  if (threadIdx.x == 0) {
    const size_t blk = (size_t)n * gridDim.x + blockIdx.x;
    partials[blk * 2 * C + c] = a;
    partials[blk * 2 * C + C + c] = b;
  }
}

// d_gamma = sum dz*yhat, d_beta = sum dz; dy = k1*dz + k2*y + k3 with
// k1 = gamma*invstd, k2 = -k1*invstd*d_gamma/M, k3 = -k1*d_beta/M - k2*mean.
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partials, int nblk, int C,
                                       double count, const float* __restrict__ gamma,
                                       const float* __restrict__ mean,
                                       const float* __restrict__ invstd, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ k1,
                                       float* __restrict__ k2, float* __restrict__ k3, int training) {
  __shared__ double red[2][32];
  const int c = blockIdx.x;
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
    a += (double)partials[(size_t)i * 2 * C + c];
    b += (double)partials[(size_t)i * 2 * C + C + c];
  }
  a = warp_sum_d(a);
  b = warp_sum_d(b);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  a = 0.0; b = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; b += red[1][w]; }
  dbeta[c] = (float)a;
  dgamma[c] = (float)b;
  const double is = invstd[c], g1 = (double)gamma[c] * is;
  // eval-mode BatchNorm normalises with constants (the running statistics): dy = gamma * invstd * dz only
  const double g2 = training ? -g1 * is * b / count : 0.0;
  k1[c] = (float)g1;
  k2[c] = (float)g2;
  k3[c] = training ? (float)(-g1 * a / count - g2 * (double)mean[c]) : 0.f;
}

struct DyCoef {  // everything needed to rebuild dy[co] from (d_a, y) on the fly
  const float *sc, *sh, *lo, *k1, *k2, *k3;
};
__device__ __forceinline__ float dy_of(float da, float yv, float s, float t, float l, float a1,
                                       float a2, float a3) {
  const float dz = (fmaf(yv, s, t) > l) ? da : 0.f;
  return fmaf(a1, dz, fmaf(a2, yv, a3));
}

// dy materialised (levels with >= 64 channels: operand of the batched tcgen05 GEMMs): dy[n][co][p], contiguous.
__global__ void __launch_bounds__(256)
dy_kernel(const float* __restrict__ d_a, long long da_ss, const float* __restrict__ y, long long y_ss, int C,
          long long HW, DyCoef k, float* __restrict__ dy) {
  const long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;  // four consecutive pixels per thread
  const int c = blockIdx.y, n = blockIdx.z;
  if (i >= HW) return;
  const size_t off = (size_t)c * HW + i;
  const float* dp = d_a + (size_t)n * da_ss + off;
  const float* yp = y + (size_t)n * y_ss + off;
  float* op = dy + ((size_t)n * C + c) * HW + i;
  const float sc = k.sc[c], sh = k.sh[c], lo = k.lo[c], k1 = k.k1[c], k2 = k.k2[c], k3 = k.k3[c];
  if (i + 4 <= HW && (((uintptr_t)dp | (uintptr_t)yp | (uintptr_t)op) & 15) == 0) {
    const float4 d = *reinterpret_cast<const float4*>(dp), v = *reinterpret_cast<const float4*>(yp);
    *reinterpret_cast<float4*>(op) = make_float4(dy_of(d.x, v.x, sc, sh, lo, k1, k2, k3), dy_of(d.y, v.y, sc, sh, lo, k1, k2, k3),
                                                 dy_of(d.z, v.z, sc, sh, lo, k1, k2, k3), dy_of(d.w, v.w, sc, sh, lo, k1, k2, k3));
  } else {
    for (int q = 0; q < 4 && i + q < HW; ++q) op[q] = dy_of(dp[q], yp[q], sc, sh, lo, k1, k2, k3);
  }
}

// ---------------------------------------------------------------------------------------------
// g[ci] = sum_co Wpw[co][ci] * dy[co]. Thread = VEC consecutive pixels, CI_T input channels.
constexpr int PWT_CO_CHUNK = 16;
template <int CI_T, int VEC>
__global__ void __launch_bounds__(256)
pwT_bwd_kernel(const float* __restrict__ d_a, long long da_ss, const float* __restrict__ y,
               long long y_ss, int Cout, long long HW, DyCoef k, const float* __restrict__ wpw,
               int Cin, float* __restrict__ g, long long g_ss) {
  __shared__ __align__(16) float sw[PWT_CO_CHUNK * CI_T];
  __shared__ float sk[6][PWT_CO_CHUNK];
  const long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * VEC;
  const int cig = (Cin + CI_T - 1) / CI_T;
  const int ci0 = (blockIdx.y % cig) * CI_T, n = blockIdx.y / cig;
  const bool ok = i < HW;
  float acc[VEC][CI_T];
#pragma unroll
  for (int v = 0; v < VEC; ++v)
#pragma unroll
    for (int c = 0; c < CI_T; ++c) acc[v][c] = 0.f;
  for (int co0 = 0; co0 < Cout; co0 += PWT_CO_CHUNK) {
    const int nco = min(PWT_CO_CHUNK, Cout - co0);
    __syncthreads();
    for (int j = threadIdx.x; j < nco * CI_T; j += 256) {
      const int o = j / CI_T, c = j - o * CI_T;
      sw[j] = (ci0 + c < Cin) ? wpw[(size_t)(co0 + o) * Cin + ci0 + c] : 0.f;
    }
    if (threadIdx.x < nco) {
      const int o = co0 + threadIdx.x;
      sk[0][threadIdx.x] = k.sc[o]; sk[1][threadIdx.x] = k.sh[o]; sk[2][threadIdx.x] = k.lo[o];
      sk[3][threadIdx.x] = k.k1[o]; sk[4][threadIdx.x] = k.k2[o]; sk[5][threadIdx.x] = k.k3[o];
    }
    __syncthreads();
    if (!ok) continue;
    for (int o = 0; o < nco; ++o) {
      const float* dp = d_a + (size_t)n * da_ss + (size_t)(co0 + o) * HW + i;
      const float* yp = y + (size_t)n * y_ss + (size_t)(co0 + o) * HW + i;
      float dv[VEC], yv[VEC];
      if (VEC == 4) {
        const float4 a = *reinterpret_cast<const float4*>(dp);
        const float4 b = *reinterpret_cast<const float4*>(yp);
        dv[0] = a.x; dv[VEC > 1 ? 1 : 0] = a.y; dv[VEC > 2 ? 2 : 0] = a.z; dv[VEC > 3 ? 3 : 0] = a.w;
        yv[0] = b.x; yv[VEC > 1 ? 1 : 0] = b.y; yv[VEC > 2 ? 2 : 0] = b.z; yv[VEC > 3 ? 3 : 0] = b.w;
      } else {
        dv[0] = dp[0]; yv[0] = yp[0];
      }
      float dy[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        dy[v] = dy_of(dv[v], yv[v], sk[0][o], sk[1][o], sk[2][o], sk[3][o], sk[4][o], sk[5][o]);
      const float4* w4 = reinterpret_cast<const float4*>(sw + o * CI_T);
#pragma unroll
      for (int c4 = 0; c4 < CI_T / 4; ++c4) {
        const float4 wv = w4[c4];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          acc[v][c4 * 4 + 0] = fmaf(dy[v], wv.x, acc[v][c4 * 4 + 0]);
          acc[v][c4 * 4 + 1] = fmaf(dy[v], wv.y, acc[v][c4 * 4 + 1]);
          acc[v][c4 * 4 + 2] = fmaf(dy[v], wv.z, acc[v][c4 * 4 + 2]);
          acc[v][c4 * 4 + 3] = fmaf(dy[v], wv.w, acc[v][c4 * 4 + 3]);
        }
      }
    }
  }
  if (!ok) return;
#pragma unroll
  for (int c = 0; c < CI_T; ++c) {
    if (ci0 + c >= Cin) continue;
    float* gp = g + (size_t)n * g_ss + (size_t)(ci0 + c) * HW + i;
    if (VEC == 4) {
      *reinterpret_cast<float4*>(gp) =
          make_float4(acc[0][c], acc[VEC > 1 ? 1 : 0][c], acc[VEC > 2 ? 2 : 0][c], acc[VEC > 3 ? 3 : 0][c]);
    } else {
      gp[0] = acc[0][c];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dWpw[co][ci] = sum_p dy[co][p] * dwout[ci][p], dwout = dw3x3(xform(x)) recomputed per tile of 32x8 pixels.
constexpr int WG_TW = 32, WG_TH = 8, WG_SROW = WG_TW + 2, WG_SPLANE = (WG_TH + 2) * WG_SROW;

// ---------------------------------------------------------------------------------------------
// Tensor-core version of pw_wgrad. The reduction dW[co][ci] = sum_p dy[co][p] * dwout[ci][p] has
// M = N = 16 per block and K = pixels: far too skinny for tcgen05 (M >= 64, operands from shared
// memory), but a perfect fit for warp-level mma.sync.m16n8k8 (TF32) fed straight from registers:
// lane (g, t) of a warp owns pixels {t, t+4} of each 8-pixel k-step for channels {g, g+8}, which
// is exactly the A (dy) and B (dwout) fragment layout, so nothing is staged or transposed.
// fp32-class accuracy through the 3xTF32 split (hi.hi + hi.lo + lo.hi); the fragment accumulators
// are flushed into fp32 registers after every tile so the tensor core's truncating accumulation
// never sees chains longer than 12 products.
__device__ __forceinline__ void tf32_split(float v, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 3)
pw_wgrad_mma_kernel(const float* __restrict__ d_a, long long da_ss, const float* __restrict__ y,
                    long long y_ss, int Cout, DyCoef k, const float* __restrict__ x, long long x_ss,
                    int Cin, int H, int W, const float* __restrict__ isc, const float* __restrict__ ish,
                    const float* __restrict__ ilo, const float* __restrict__ wdw, int N, int tiles_x,
                    int tiles_y, float* __restrict__ partials) {
  __shared__ float xs[16 * WG_SPLANE];
  __shared__ float sred[8][256];
  __shared__ float sxf[3][16];
  const int tid = threadIdx.x, lane = tid & 31, ty = tid >> 5, g = lane >> 2, t = lane & 3;
  const int cit = (Cin + 15) / 16;
  const int co0 = (blockIdx.y / cit) * 16, ci0 = (blockIdx.y % cit) * 16;
  const int nco = min(16, Cout - co0), nci = min(16, Cin - ci0);
  const size_t HW = (size_t)H * W;
  if (tid < 16) {
    const bool v = tid < nci && isc != nullptr;
    sxf[0][tid] = v ? isc[ci0 + tid] : 1.f; sxf[1][tid] = v ? ish[ci0 + tid] : 0.f; sxf[2][tid] = v ? ilo[ci0 + tid] : -INFINITY;
  }
  // per-lane constants for its two output channels (g, g+8) and two input channels (g, g+8)
  float ksc[2], ksh[2], klo[2], kk1[2], kk2[2], kk3[2], wd[2][9];
  bool cov[2], civ[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int o = g + 8 * h;
    cov[h] = o < nco;
    civ[h] = o < nci;
    ksc[h] = cov[h] ? k.sc[co0 + o] : 0.f; ksh[h] = cov[h] ? k.sh[co0 + o] : 0.f; klo[h] = cov[h] ? k.lo[co0 + o] : 0.f;
    kk1[h] = cov[h] ? k.k1[co0 + o] : 0.f; kk2[h] = cov[h] ? k.k2[co0 + o] : 0.f; kk3[h] = cov[h] ? k.k3[co0 + o] : 0.f;
#pragma unroll
    for (int q = 0; q < 9; ++q) wd[h][q] = civ[h] ? wdw[(size_t)(ci0 + o) * 9 + q] : 0.f;
  }
  float ctot[2][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) ctot[j][q] = 0.f;
  const int tiles = tiles_x * tiles_y;
  const long long total = (long long)N * tiles;
  for (long long work = blockIdx.x; work < total; work += gridDim.x) {
    const int n = (int)(work / tiles), tile = (int)(work % tiles);
    const int x0 = (tile % tiles_x) * WG_TW, y0 = (tile / tiles_x) * WG_TH;
    __syncthreads();
    {
      constexpr int NSLOT = (WG_SPLANE + 255) / 256;
#pragma unroll
      for (int sl = 0; sl < NSLOT; ++sl) {
        const int pos = tid + 256 * sl;
        if (pos < WG_SPLANE) {
          const int ry = pos / WG_SROW, rx = pos - ry * WG_SROW;
          const int gy = y0 + ry - 1, gx = x0 + rx - 1;
          const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
          const float* src = x + (size_t)n * x_ss + (size_t)ci0 * HW + (size_t)(in ? gy * W + gx : 0);
          float v[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] = (in && c < nci) ? src[(size_t)c * HW] : 0.f;
          if (in) {
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = xform_apply(v[c], sxf[0][c], sxf[1][c], sxf[2][c]);
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) xs[c * WG_SPLANE + pos] = (in && c < nci) ? v[c] : 0.f;
        }
      }
    }
    // dy fragments of the whole tile row (4 k-steps x {t, t+4} x {g, g+8}) : issue all loads first
    const int gy = y0 + ty;
    float dyv[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int h = q & 1, px = x0 + ks * 8 + t + 4 * (q >> 1);
        float v = 0.f;
        if (cov[h] && gy < H && px < W) {
          const size_t off = (size_t)(co0 + g + 8 * h) * HW + (size_t)gy * W + px;
          v = dy_of(d_a[(size_t)n * da_ss + off], y[(size_t)n * y_ss + off], ksc[h], ksh[h], klo[h], kk1[h], kk2[h], kk3[h]);
        }
        dyv[ks][q] = v;  // q: 0 = (g, t), 1 = (g+8, t), 2 = (g, t+4), 3 = (g+8, t+4)
      }
    __syncthreads();
    float c[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) c[j][q] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[4], al[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) tf32_split(dyv[ks][q], ah[q], al[q]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        // dwout[ci = 8j + g][pixel t / t+4] : depthwise stencil from the shared tile
        float b[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int lx = ks * 8 + t + 4 * e;
          const float* tp = xs + (8 * j + g) * WG_SPLANE + ty * WG_SROW + lx;
          float sacc = 0.f;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) sacc = fmaf(tp[ky * WG_SROW + kx], wd[j][ky * 3 + kx], sacc);
          b[e] = (gy < H && x0 + lx < W) ? sacc : 0.f;
        }
        uint32_t bh0, bl0, bh1, bl1;
        tf32_split(b[0], bh0, bl0);
        tf32_split(b[1], bh1, bl1);
        mma_tf32(c[j], al, bh0, bh1);
        mma_tf32(c[j], ah, bl0, bl1);
        mma_tf32(c[j], ah, bh0, bh1);
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) ctot[j][q] += c[j][q];
  }
  // C fragment -> (co, ci): c0 (g, 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1), n-tile j adds 8 to ci
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    sred[ty][(g) * 16 + 8 * j + 2 * t] = ctot[j][0];
    sred[ty][(g) * 16 + 8 * j + 2 * t + 1] = ctot[j][1];
    sred[ty][(g + 8) * 16 + 8 * j + 2 * t] = ctot[j][2];
    sred[ty][(g + 8) * 16 + 8 * j + 2 * t + 1] = ctot[j][3];
  }
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += sred[w][tid];
  const int o = tid >> 4, ci = tid & 15;
  if (o < nco && ci < nci) partials[((size_t)blockIdx.x * Cout + co0 + o) * Cin + ci0 + ci] = sum;
}

// ---------------------------------------------------------------------------------------------
// Depthwise 3x3 backward for one channel plane tile: d_x = corr(g, flip(w)); dW[k] partials.
constexpr int DW_TH = 32;
__global__ void __launch_bounds__(256)
dw_bwd_kernel(const float* __restrict__ g, long long g_ss, const float* __restrict__ x,
              long long x_ss, int C, int H, int W, const float* __restrict__ isc,
              const float* __restrict__ ish, const float* __restrict__ ilo,
              const float* __restrict__ wdw, float* __restrict__ dx, long long dx_ss,
              int accumulate, float* __restrict__ partials, int tiles_x, int tiles) {
  __shared__ float gs[(DW_TH + 2) * 34];
  __shared__ float xs[(DW_TH + 2) * 34];
  __shared__ float red[8][9];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int c = blockIdx.y, n = blockIdx.z;
  const size_t HW = (size_t)H * W;
  const float* gp = g + (size_t)n * g_ss + (size_t)c * HW;
  const float* xp = x + (size_t)n * x_ss + (size_t)c * HW;
  const bool need_x = partials != nullptr;
  float s = 1.f, t = 0.f, l = -INFINITY;
  if (isc) { s = isc[c]; t = ish[c]; l = ilo[c]; }
  float w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w[k] = wdw[(size_t)c * 9 + k];
  float dw[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) dw[k] = 0.f;
  float* dxp = dx + (size_t)n * dx_ss + (size_t)c * HW;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
  const int x0 = (tile % tiles_x) * 32, y0 = (tile / tiles_x) * DW_TH;
  __syncthreads();
  {
    constexpr int NPOS = (DW_TH + 2) * 34, NSLOT = (NPOS + 255) / 256;
    float gv[NSLOT], xv[NSLOT];
#pragma unroll
    for (int sl = 0; sl < NSLOT; ++sl) {
      const int i = tid + 256 * sl;
      const int ry = i / 34, rx = i - ry * 34;
      const int gy = y0 + ry - 1, gx = x0 + rx - 1;
      const bool in = i < NPOS && gy >= 0 && gy < H && gx >= 0 && gx < W;
      gv[sl] = in ? gp[(size_t)gy * W + gx] : 0.f;
      xv[sl] = (in && need_x) ? xp[(size_t)gy * W + gx] : 0.f;
      if (in && need_x && isc) xv[sl] = xform_apply(xv[sl], s, t, l);
    }
#pragma unroll
    for (int sl = 0; sl < NSLOT; ++sl) {
      const int i = tid + 256 * sl;
      if (i < NPOS) { gs[i] = gv[sl]; xs[i] = xv[sl]; }
    }
  }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < DW_TH / 8; ++p) {
    const int ly = ty * (DW_TH / 8) + p, gy = y0 + ly, gx = x0 + tx;
    if (gy < H && gx < W) {
      float v = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) v = fmaf(w[ky * 3 + kx], gs[(ly + 2 - ky) * 34 + tx + 2 - kx], v);
      const size_t off = (size_t)gy * W + gx;
      dxp[off] = accumulate ? dxp[off] + v : v;
      if (need_x) {
        const float gc = gs[(ly + 1) * 34 + tx + 1];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) dw[ky * 3 + kx] = fmaf(gc, xs[(ly + ky) * 34 + tx + kx], dw[ky * 3 + kx]);
      }
    }
  }
  }  // tile loop
  if (!need_x) return;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float v = warp_sum(dw[k]);
    if (tx == 0) red[ty][k] = v;
  }
  __syncthreads();
  if (tid < 9) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += red[q][tid];
    const size_t blk = (size_t)n * gridDim.x + blockIdx.x;
    partials[(blk * C + c) * 9 + tid] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d(2) backward: thread per full-resolution pixel; gradient goes to the first maximum
// of its 2x2 window (aten's tie rule), zero elsewhere and in a trailing odd row/column.
__global__ void pool2_bwd_kernel(const float* __restrict__ x, long long x_ss, int C, int H, int W,
                                 const float* __restrict__ sc, const float* __restrict__ sh,
                                 const float* __restrict__ lo, const float* __restrict__ dout,
                                 long long dout_ss, float* __restrict__ din, long long din_ss) {
  // thread = one 2x2 cell (cy, cx) of the full-resolution plane, including partial cells of an odd edge
  const int cx = blockIdx.x * blockDim.x + threadIdx.x;
  const int cy = blockIdx.y * blockDim.y + threadIdx.y;
  const int c = blockIdx.z % C, n = blockIdx.z / C;
  const int iy = 2 * cy, ix = 2 * cx;
  if (ix >= W || iy >= H) return;
  const int Ho = H / 2, Wo = W / 2;
  float* dp = din + (size_t)n * din_ss + (size_t)c * H * W + (size_t)iy * W + ix;
  float r[4] = {0.f, 0.f, 0.f, 0.f};
  if (cy < Ho && cx < Wo) {
    const float* p = x + (size_t)n * x_ss + (size_t)c * H * W + (size_t)iy * W + ix;
    float v[4];
    if ((W & 1) == 0 && (((uintptr_t)p) & 7) == 0) {
      const float2 a = *reinterpret_cast<const float2*>(p), b = *reinterpret_cast<const float2*>(p + W);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
      v[0] = p[0]; v[1] = p[1]; v[2] = p[W]; v[3] = p[W + 1];
    }
    if (sc) {
      const float s = sc[c], t = sh[c], l = lo[c];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = xform_apply(v[q], s, t, l);
    }
    int am = 0;
    float m = v[0];
#pragma unroll
    for (int q = 1; q < 4; ++q)
      if (v[q] > m || isnan(v[q])) { m = v[q]; am = q; }
    const float gv = dout[(size_t)n * dout_ss + (size_t)c * Ho * Wo + (size_t)cy * Wo + cx];
#pragma unroll
    for (int q = 0; q < 4; ++q) r[q] = (q == am) ? gv : 0.f;
  }
  const bool x1 = ix + 1 < W, y1 = iy + 1 < H;
  if (x1 && (W & 1) == 0 && (((uintptr_t)dp) & 7) == 0) {
    *reinterpret_cast<float2*>(dp) = make_float2(r[0], r[1]);
    if (y1) *reinterpret_cast<float2*>(dp + W) = make_float2(r[2], r[3]);
  } else {
    dp[0] = r[0];
    if (x1) dp[1] = r[1];
    if (y1) { dp[W] = r[2]; if (x1) dp[W + 1] = r[3]; }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused producers of the BatchNorm-backward sums (sum dz, sum dz * yhat of the block whose activated output these
// kernels differentiate): the separate bnrelu_bwd_reduce pass over d_a and y is not needed after them.

// MaxPool2d(2) backward + the sums for the block that produced x. partials: [N * gridDim.y * gridDim.x][2][C].
__global__ void pool2_bwd_bn_kernel(const float* __restrict__ x, long long x_ss, int C, int H, int W,
                                    const float* __restrict__ sc, const float* __restrict__ sh,
                                    const float* __restrict__ lo, const float* __restrict__ dout, long long dout_ss,
                                    float* __restrict__ din, long long din_ss, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, float* __restrict__ partials) {
  __shared__ float red[32];
  const int cx = blockIdx.x * blockDim.x + threadIdx.x;
  const int cy = blockIdx.y * blockDim.y + threadIdx.y;
  const int c = blockIdx.z % C, n = blockIdx.z / C;
  const int iy = 2 * cy, ix = 2 * cx;
  const int Ho = H / 2, Wo = W / 2;
  float s1 = 0.f, s2 = 0.f;
  if (ix < W && iy < H) {
    float* dp = din + (size_t)n * din_ss + (size_t)c * H * W + (size_t)iy * W + ix;
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (cy < Ho && cx < Wo) {
      const float* p = x + (size_t)n * x_ss + (size_t)c * H * W + (size_t)iy * W + ix;
      float raw[4], v[4];
      if ((W & 1) == 0 && (((uintptr_t)p) & 7) == 0) {
        const float2 a = *reinterpret_cast<const float2*>(p), b = *reinterpret_cast<const float2*>(p + W);
        raw[0] = a.x; raw[1] = a.y; raw[2] = b.x; raw[3] = b.y;
      } else {
        raw[0] = p[0]; raw[1] = p[1]; raw[2] = p[W]; raw[3] = p[W + 1];
      }
      const float s = sc[c], t = sh[c], l = lo[c];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = xform_apply(raw[q], s, t, l);
      int am = 0;
      float m = v[0], rm = raw[0];
#pragma unroll
      for (int q = 1; q < 4; ++q)
        if (v[q] > m || isnan(v[q])) { m = v[q]; am = q; rm = raw[q]; }
      const float gv = dout[(size_t)n * dout_ss + (size_t)c * Ho * Wo + (size_t)cy * Wo + cx];
#pragma unroll
      for (int q = 0; q < 4; ++q) r[q] = (q == am) ? gv : 0.f;
      const float dz = (fmaf(rm, s, t) > l) ? gv : 0.f;  // the only non-zero d_a of the window sits at the arg-max
      s1 = dz;
      s2 = dz * (rm - mean[c]) * invstd[c];
    }
    const bool x1 = ix + 1 < W, y1 = iy + 1 < H;
    if (x1 && (W & 1) == 0 && (((uintptr_t)dp) & 7) == 0) {
      *reinterpret_cast<float2*>(dp) = make_float2(r[0], r[1]);
      if (y1) *reinterpret_cast<float2*>(dp + W) = make_float2(r[2], r[3]);
    } else {
      dp[0] = r[0];
      if (x1) dp[1] = r[1];
      if (y1) { dp[W] = r[2]; if (x1) dp[W + 1] = r[3]; }
    }
  }
  // block (32 x 8) reduction; block_sum indexes by threadIdx.x only, so flatten first
  const int tid = threadIdx.y * 32 + threadIdx.x, lane = tid & 31, wid = tid >> 5;
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) { red[wid] = s1; red[8 + wid] = s2; }
  __syncthreads();
  if (tid < 2) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[8 * tid + w];
    const size_t row = ((size_t)n * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    partials[(row * 2 + tid) * C + c] = t;
  }
}

// out_conv (C = 8 -> 1) + Sigmoid backward, persistent and vectorised, with the weight/bias partial sums and the
// BatchNorm-backward sums of the block that produced x in registers for the whole kernel:
//   wb_partials [gridDim.x][9]: sum_p dz * a[c] (c < 8), sum_p dz;  bn_partials [gridDim.x][2][8].
__global__ void __launch_bounds__(256)
outconv8_bwd_bn_kernel(const float* __restrict__ dp, const float* __restrict__ prob, const float* __restrict__ x,
                       long long x_ss, long long HW, int N, const float* __restrict__ sc,
                       const float* __restrict__ sh, const float* __restrict__ lo, const float* __restrict__ w,
                       float* __restrict__ d_a, long long da_ss, const float* __restrict__ mean,
                       const float* __restrict__ invstd, float* __restrict__ wb_partials,
                       float* __restrict__ bn_partials) {
  constexpr int C = 8;
  __shared__ float sred[8][32];
  float acc[32];  // [0..8): dW, 8: dbias, [9..17): sum dz_c, [17..25): sum dz_c * yhat, rest unused
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  float s[C], t[C], l[C], wv[C], mu[C], is[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    s[c] = sc ? sc[c] : 1.f; t[c] = sc ? sh[c] : 0.f; l[c] = sc ? lo[c] : -INFINITY; wv[c] = w[c];
    mu[c] = mean ? mean[c] : 0.f; is[c] = mean ? invstd[c] : 0.f;
  }
  const long long q_per_n = HW / 4, total = q_per_n * N;
  for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < total; q += (long long)gridDim.x * 256) {
    const int n = (int)(q / q_per_n);
    const long long i = (q - (long long)n * q_per_n) * 4;
    const float4 p4 = *reinterpret_cast<const float4*>(prob + (size_t)n * HW + i);
    const float4 g4 = *reinterpret_cast<const float4*>(dp + (size_t)n * HW + i);
    const float dz[4] = {g4.x * p4.x * (1.f - p4.x), g4.y * p4.y * (1.f - p4.y), g4.z * p4.z * (1.f - p4.z), g4.w * p4.w * (1.f - p4.w)};
    acc[8] += (dz[0] + dz[1]) + (dz[2] + dz[3]);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 r4 = *reinterpret_cast<const float4*>(x + (size_t)n * x_ss + (size_t)c * HW + i);
      const float raw[4] = {r4.x, r4.y, r4.z, r4.w};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pre = fmaf(raw[e], s[c], t[c]);
        const float a = fmaxf(pre, l[c]);
        o[e] = wv[c] * dz[e];
        acc[c] = fmaf(a, dz[e], acc[c]);
        const float dzc = pre > l[c] ? o[e] : 0.f;
        acc[9 + c] += dzc;
        acc[17 + c] = fmaf(dzc, (raw[e] - mu[c]) * is[c], acc[17 + c]);
      }
      *reinterpret_cast<float4*>(d_a + (size_t)n * da_ss + (size_t)c * HW + i) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // transposing butterfly: lane j ends with the warp total of acc[j]
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < o; ++j) {
      const float send = up ? acc[j] : acc[j + o];
      const float keep = up ? acc[j + o] : acc[j];
      acc[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  sred[wid][lane] = acc[0];
  __syncthreads();
  if (threadIdx.x < 25) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += sred[k][threadIdx.x];
    const int j = threadIdx.x;
    if (j < 9) wb_partials[(size_t)blockIdx.x * 9 + j] = v;
    else if (bn_partials) bn_partials[((size_t)blockIdx.x * 2 + (j >= 17)) * C + (j - 9) % 8] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// ConvTranspose2d backward (data): d_x[ci][iy][ix] = sum_co,k W[ci][co][k] * d_out[co][2iy+ky][2ix+kx].
template <int CI_T>
__global__ void __launch_bounds__(256)
convt_bwd_data_kernel(const float* __restrict__ dout, long long dout_ss, int Cout, int Hs, int Ws,
                      const float* __restrict__ w, int Cin, int Hin, int Win,
                      float* __restrict__ dx, long long dx_ss) {
  __shared__ __align__(16) float sw[8 * 9 * CI_T];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int ix = blockIdx.x * 32 + threadIdx.x, iy = blockIdx.y * 8 + threadIdx.y;
  const int cig = (Cin + CI_T - 1) / CI_T;
  const int ci0 = (blockIdx.z % cig) * CI_T, n = blockIdx.z / cig;
  const bool ok = ix < Win && iy < Hin;
  const size_t HWs = (size_t)Hs * Ws;
  float acc[CI_T];
#pragma unroll
  for (int c = 0; c < CI_T; ++c) acc[c] = 0.f;
  for (int co0 = 0; co0 < Cout; co0 += 8) {
    const int nco = min(8, Cout - co0);
    __syncthreads();
    for (int i = tid; i < nco * 9 * CI_T; i += 256) {
      const int o = i / (9 * CI_T), r = i - o * 9 * CI_T, kk = r / CI_T, c = r - kk * CI_T;
      sw[i] = (ci0 + c < Cin) ? w[((size_t)(ci0 + c) * Cout + co0 + o) * 9 + kk] : 0.f;
    }
    __syncthreads();
    if (!ok) continue;
    for (int o = 0; o < nco; ++o) {
      const float* dp = dout + (size_t)n * dout_ss + (size_t)(co0 + o) * HWs;
      float d[9];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int oy = 2 * iy + ky, ox = 2 * ix + kx;
          d[ky * 3 + kx] = (oy < Hs && ox < Ws) ? dp[(size_t)oy * Ws + ox] : 0.f;
        }
      const float* wo = sw + o * 9 * CI_T;
#pragma unroll
      for (int kk = 0; kk < 9; ++kk)
#pragma unroll
        for (int c = 0; c < CI_T; ++c) acc[c] = fmaf(d[kk], wo[kk * CI_T + c], acc[c]);
    }
  }
  if (!ok) return;
#pragma unroll
  for (int c = 0; c < CI_T; ++c)
    if (ci0 + c < Cin)
      dx[(size_t)n * dx_ss + (size_t)(ci0 + c) * Hin * Win + (size_t)iy * Win + ix] = acc[c];
}

// ConvTranspose2d weight gradient: dW[ci][co][k] = sum_p xact[ci][p] * d_out[co][2iy+ky][2ix+kx].
// Tensor-core version (same fragment trick as pw_wgrad_mma_kernel): rows a = 16 input channels,
// columns b = 16 of the Cout*9 (co, tap) pairs, K = input pixels; operands gathered from global
// memory directly in mma.sync fragment layout, 3xTF32, accumulators flushed every 4 k-steps.
__global__ void __launch_bounds__(256)
convt_wgrad_mma_kernel(const float* __restrict__ x, long long x_ss, int Cin, int Hin, int Win,
                       const float* __restrict__ isc, const float* __restrict__ ish,
                       const float* __restrict__ ilo, const float* __restrict__ dout, long long dout_ss,
                       int Cout, int Hs, int Ws, int N, float* __restrict__ partials) {
  __shared__ float sred[8][256];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  const int CK = Cout * 9;
  const int bt = (CK + 15) / 16;
  const int ci0 = (blockIdx.y / bt) * 16, b0 = (blockIdx.y % bt) * 16;
  const int nci = min(16, Cin - ci0), nb = min(16, CK - b0);
  const size_t HWi = (size_t)Hin * Win, HWs = (size_t)Hs * Ws;
  const long long total = (long long)N * HWi;
  float sc[2], sh[2], lo[2];
  bool civ[2];
  int bco[2], bky[2], bkx[2];
  bool bv[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = g + 8 * h;
    civ[h] = c < nci;
    sc[h] = (civ[h] && isc) ? isc[ci0 + c] : 1.f;
    sh[h] = (civ[h] && isc) ? ish[ci0 + c] : 0.f;
    lo[h] = (civ[h] && isc) ? ilo[ci0 + c] : -INFINITY;
    const int j = b0 + 8 * h + g;  // this lane's B row for n-tile h
    bv[h] = 8 * h + g < nb;
    bco[h] = bv[h] ? j / 9 : 0;
    const int kk = bv[h] ? j - bco[h] * 9 : 0;
    bky[h] = kk / 3;
    bkx[h] = kk - bky[h] * 3;
  }
  float ctot[2][4], c[2][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) { ctot[j][q] = 0.f; c[j][q] = 0.f; }
  int since_flush = 0;
  const long long stride = (long long)gridDim.x * 8 * 8;
  for (long long base = ((long long)blockIdx.x * 8 + wid) * 8; base < total; base += stride) {
    // this lane's two pixels of the 8-pixel k-step
    float av[4];
    float bvv[2][2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const long long pidx = base + t + 4 * e;
      const bool ok = pidx < total;
      const int n = ok ? (int)(pidx / (long long)HWi) : 0;
      const int rem = ok ? (int)(pidx - (long long)n * (long long)HWi) : 0;
      const int iy = rem / Win, ix = rem - iy * Win;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v = 0.f;
        if (ok && civ[h]) v = xform_apply(x[(size_t)n * x_ss + (size_t)(ci0 + g + 8 * h) * HWi + rem], sc[h], sh[h], lo[h]);
        av[h + 2 * e] = v;  // 0 = (g, t), 1 = (g+8, t), 2 = (g, t+4), 3 = (g+8, t+4)
        float w = 0.f;
        const int oy = 2 * iy + bky[h], ox = 2 * ix + bkx[h];
        if (ok && bv[h] && oy < Hs && ox < Ws) w = dout[(size_t)n * dout_ss + (size_t)bco[h] * HWs + (size_t)oy * Ws + ox];
        bvv[h][e] = w;
      }
    }
    uint32_t ah[4], al[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) tf32_split(av[q], ah[q], al[q]);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t bh0, bl0, bh1, bl1;
      tf32_split(bvv[j][0], bh0, bl0);
      tf32_split(bvv[j][1], bh1, bl1);
      mma_tf32(c[j], al, bh0, bh1);
      mma_tf32(c[j], ah, bl0, bl1);
      mma_tf32(c[j], ah, bh0, bh1);
    }
    if (++since_flush == 4) {
      since_flush = 0;
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) { ctot[j][q] += c[j][q]; c[j][q] = 0.f; }
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) ctot[j][q] += c[j][q];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    sred[wid][(g) * 16 + 8 * j + 2 * t] = ctot[j][0];
    sred[wid][(g) * 16 + 8 * j + 2 * t + 1] = ctot[j][1];
    sred[wid][(g + 8) * 16 + 8 * j + 2 * t] = ctot[j][2];
    sred[wid][(g + 8) * 16 + 8 * j + 2 * t + 1] = ctot[j][3];
  }
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += sred[w][tid];
  const int ci = tid >> 4, j = tid & 15;
  if (ci < nci && j < nb) partials[((size_t)blockIdx.x * Cin + ci0 + ci) * CK + b0 + j] = sum;
}

// Per-channel sum of a view over (N, H*W): ConvTranspose2d bias gradient. partials [N*chunks][C].
__global__ void __launch_bounds__(256)
plane_sum_kernel(const float* __restrict__ v, long long v_ss, int C, long long HW,
                 float* __restrict__ partials) {
  __shared__ float red[32];
  const int c = blockIdx.y, n = blockIdx.z;
  const float* p = v + (size_t)n * v_ss + (size_t)c * HW;
  const long long i0 = (long long)blockIdx.x * RED_CHUNK, i1 = min(i0 + RED_CHUNK, HW);
  float a = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) a += p[i];
  a = block_sum(a, red);
  if (threadIdx.x == 0) partials[((size_t)n * gridDim.x + blockIdx.x) * C + c] = a;
}

// Dcol rows (co, ky, kx) of one (sample, output channel): thread = input pixel (qy, qx), nine strided reads of dout.
__global__ void __launch_bounds__(256)
convt_im2col_kernel(const float* __restrict__ dout, long long dout_ss, int Cout, int Hs, int Ws, int Hin, int Win,
                    float* __restrict__ dcol) {
  const int qx = blockIdx.x * 32 + threadIdx.x, qy = blockIdx.y * 8 + threadIdx.y;
  const int co = blockIdx.z % Cout, n = blockIdx.z / Cout;
  if (qx >= Win || qy >= Hin) return;
  const float* dp = dout + (size_t)n * dout_ss + (size_t)co * Hs * Ws;
  const size_t HWi = (size_t)Hin * Win;
  float* cp = dcol + ((size_t)n * Cout + co) * 9 * HWi + (size_t)qy * Win + qx;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int oy = 2 * qy + ky, ox = 2 * qx + kx;
      cp[(size_t)(ky * 3 + kx) * HWi] = (oy < Hs && ox < Ws) ? dp[(size_t)oy * Ws + ox] : 0.f;
    }
}

}  // namespace

bool ocrs_convt_mma_bwd(const float* dout, long long dout_ss, int N, int Cout, int Hs, int Ws, const float* w, int Cin, int Hin,
                        int Win, float* dx, long long dx_ss, cudaStream_t st);  // csrc/det_convt.cu

extern "C" {

int ocrs_finalize_partials(const float* partials, int nblk, int K, float* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const bool al = ((uintptr_t)partials % 16 == 0) && ((uintptr_t)out % 16 == 0);
  if (K % 4 == 0 && al && K >= 16384 && nblk <= 256)
    finalize_partials4_kernel<<<ocrs_cdiv(K / 4, 128), 128, 0, st>>>(partials, nblk, K, out);
  else if (nblk >= 256)
    finalize_partials_kernel<32><<<ocrs_cdiv(K, 32), 1024, 0, st>>>(partials, nblk, K, out);
  else
    finalize_partials_kernel<8><<<ocrs_cdiv(K, 32), 256, 0, st>>>(partials, nblk, K, out);
  OCRS_CHECK_LAUNCH("finalize_partials_kernel");
  return 0;
}

int ocrs_det_outconv_bwd_rows(int N, int H, int W) { return N * ocrs_cdiv((long long)H * W, 256); }
// partials: [rows][C+1] (dW[0..C), db)
int ocrs_det_outconv_bwd(const float* dp, const float* prob, const float* x, long long x_ss, int N,
                         int C, int H, int W, const float* sc, const float* sh, const float* lo,
                         const float* w, float* d_a, long long da_ss, float* partials, void* stream) {
  const long long HW = (long long)H * W;
  dim3 grid(ocrs_cdiv(HW, 256), N);
  outconv_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dp, prob, x, x_ss, C, HW, sc, sh, lo, w,
                                                             d_a, da_ss, partials);
  OCRS_CHECK_LAUNCH("outconv_bwd_kernel");
  return 0;
}

int ocrs_reduce_rows(int N, long long HW) { return N * ocrs_cdiv(HW, RED_CHUNK); }

// partials: [ocrs_reduce_rows][2][C]
int ocrs_bnrelu_bwd_reduce(const float* d_a, long long da_ss, const float* y, long long y_ss, int N,
                           int C, long long HW, const float* sc, const float* sh, const float* lo,
                           const float* mean, const float* invstd, float* partials, void* stream) {
  dim3 grid(ocrs_cdiv(HW, RED_CHUNK), C, N);
  bnrelu_bwd_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_a, da_ss, y, y_ss, C, HW, sc, sh,
                                                                   lo, mean, invstd, partials);
  OCRS_CHECK_LAUNCH("bnrelu_bwd_reduce_kernel");
  return 0;
}

int ocrs_bn_bwd_finalize(const float* partials, int nblk, int C, double count, const float* gamma,
                         const float* mean, const float* invstd, float* dgamma, float* dbeta,
                         float* k1, float* k2, float* k3, int training, void* stream) {
  bn_bwd_finalize_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(partials, nblk, C, count, gamma, mean,
                                                              invstd, dgamma, dbeta, k1, k2, k3, training);
  OCRS_CHECK_LAUNCH("bn_bwd_finalize_kernel");
  return 0;
}

// Gradient side of ConvTranspose2d as a GEMM: Dcol[n][(co, ky, kx)][qy*Win + qx] = dout[n][co][2qy + ky][2qx + kx]
// (0 outside the Hs x Ws crop), the operand shared by the data gradient dx[n] = W Dcol[n] and the weight gradient
// dW = sum_n x[n] Dcol[n]^T (ocrs_gemm_tc_batched) of the levels with >= 64 input channels.
int ocrs_det_convt_im2col(const float* dout, long long dout_ss, int N, int Cout, int Hs, int Ws, int Hin, int Win,
                          float* dcol, void* stream) {
  OCRS_CHECK_ARG(N > 0 && Cout > 0 && Hin > 0 && Win > 0, "convt_im2col: bad dims");
  dim3 block(32, 8), grid(ocrs_cdiv(Win, 32), ocrs_cdiv(Hin, 8), N * Cout);
  convt_im2col_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(dout, dout_ss, Cout, Hs, Ws, Hin, Win, dcol);
  OCRS_CHECK_LAUNCH("convt_im2col_kernel");
  return 0;
}

// dy = k1 * dz + k2 * y + k3 (BatchNorm + ReLU backward of d_a) written contiguously as [N][Cout][HW]: the operand
// of the batched tcgen05 GEMMs for the 1x1 data and weight gradients of the levels with >= 64 channels.
int ocrs_det_dy(const float* d_a, long long da_ss, const float* y, long long y_ss, int N, int Cout, long long HW,
                const float* sc, const float* sh, const float* lo, const float* k1, const float* k2, const float* k3,
                float* dy, void* stream) {
  DyCoef k{sc, sh, lo, k1, k2, k3};
  dim3 grid(ocrs_cdiv(HW, 1024), Cout, N);
  dy_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_a, da_ss, y, y_ss, Cout, HW, k, dy);
  OCRS_CHECK_LAUNCH("dy_kernel");
  return 0;
}

int ocrs_det_pwT_bwd(const float* d_a, long long da_ss, const float* y, long long y_ss, int N,
                     int Cout, long long HW, const float* sc, const float* sh, const float* lo,
                     const float* k1, const float* k2, const float* k3, const float* wpw, int Cin,
                     float* g, long long g_ss, void* stream) {
  DyCoef k{sc, sh, lo, k1, k2, k3};
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (HW % 4 == 0) && (da_ss % 4 == 0) && (y_ss % 4 == 0) && (g_ss % 4 == 0) &&
                   ((uintptr_t)d_a % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)g % 16 == 0);
  const int cit = Cin <= 8 ? 8 : 16;
  const int cig = ocrs_cdiv(Cin, cit);
  if (vec) {
    dim3 grid(ocrs_cdiv(HW, 1024), N * cig);
    if (cit == 8)
      pwT_bwd_kernel<8, 4><<<grid, 256, 0, st>>>(d_a, da_ss, y, y_ss, Cout, HW, k, wpw, Cin, g, g_ss);
    else
      pwT_bwd_kernel<16, 4><<<grid, 256, 0, st>>>(d_a, da_ss, y, y_ss, Cout, HW, k, wpw, Cin, g, g_ss);
  } else {
    dim3 grid(ocrs_cdiv(HW, 256), N * cig);
    if (cit == 8)
      pwT_bwd_kernel<8, 1><<<grid, 256, 0, st>>>(d_a, da_ss, y, y_ss, Cout, HW, k, wpw, Cin, g, g_ss);
    else
      pwT_bwd_kernel<16, 1><<<grid, 256, 0, st>>>(d_a, da_ss, y, y_ss, Cout, HW, k, wpw, Cin, g, g_ss);
  }
  OCRS_CHECK_LAUNCH("pwT_bwd_kernel");
  return 0;
}

int ocrs_det_pw_wgrad_workers(int N, int H, int W) {
  const long long tiles = (long long)N * ocrs_cdiv(W, WG_TW) * ocrs_cdiv(H, WG_TH);
  return (int)(tiles < 3 * OCRS_NUM_SMS ? tiles : 3 * OCRS_NUM_SMS);
}
// partials: [workers][Cout][Cin], fully written (every (co, ci) belongs to exactly one block column).
int ocrs_det_pw_wgrad(const float* d_a, long long da_ss, const float* y, long long y_ss, int N,
                      int Cout, int H, int W, const float* sc, const float* sh, const float* lo,
                      const float* k1, const float* k2, const float* k3, const float* x,
                      long long x_ss, int Cin, const float* isc, const float* ish, const float* ilo,
                      const float* wdw, float* partials, void* stream) {
  DyCoef k{sc, sh, lo, k1, k2, k3};
  const int tiles_x = ocrs_cdiv(W, WG_TW), tiles_y = ocrs_cdiv(H, WG_TH);
  dim3 grid(ocrs_det_pw_wgrad_workers(N, H, W), ocrs_cdiv(Cout, 16) * ocrs_cdiv(Cin, 16));
  pw_wgrad_mma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_a, da_ss, y, y_ss, Cout, k, x, x_ss, Cin, H, W, isc,
                                                               ish, ilo, wdw, N, tiles_x, tiles_y, partials);
  OCRS_CHECK_LAUNCH("pw_wgrad_kernel");
  return 0;
}

constexpr int DW_MAX_BLOCKS_X = 16;
int ocrs_det_dw_bwd_rows(int N, int H, int W) {
  const int tiles = ocrs_cdiv(W, 32) * ocrs_cdiv(H, DW_TH);
  return N * (tiles < DW_MAX_BLOCKS_X ? tiles : DW_MAX_BLOCKS_X);
}
// partials: [rows][C][9] or NULL to skip the weight gradient (and the read of x).
int ocrs_det_dw_bwd(const float* g, long long g_ss, const float* x, long long x_ss, int N, int C, int H,
                    int W, const float* isc, const float* ish, const float* ilo, const float* wdw,
                    float* dx, long long dx_ss, int accumulate, float* partials, void* stream) {
  const int tiles_x = ocrs_cdiv(W, 32), tiles_y = ocrs_cdiv(H, DW_TH);
  const int tiles = tiles_x * tiles_y;
  dim3 grid(tiles < DW_MAX_BLOCKS_X ? tiles : DW_MAX_BLOCKS_X, C, N);
  dw_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, g_ss, x, x_ss, C, H, W, isc, ish, ilo, wdw, dx,
                                                        dx_ss, accumulate, partials, tiles_x, tiles);
  OCRS_CHECK_LAUNCH("dw_bwd_kernel");
  return 0;
}

int ocrs_det_pool2_bwd(const float* x, long long x_ss, int N, int C, int H, int W, const float* sc,
                       const float* sh, const float* lo, const float* dout, long long dout_ss,
                       float* din, long long din_ss, void* stream) {
  dim3 block(32, 8), grid(ocrs_cdiv((W + 1) / 2, 32), ocrs_cdiv((H + 1) / 2, 8), N * C);
  pool2_bwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, x_ss, C, H, W, sc, sh, lo, dout, dout_ss,
                                                             din, din_ss);
  OCRS_CHECK_LAUNCH("pool2_bwd_kernel");
  return 0;
}

// Rows of the [rows][2][C] BatchNorm partials ocrs_det_pool2_bwd_bn writes.
int ocrs_det_pool2_bwd_bn_rows(int N, int H, int W) { return N * ocrs_cdiv((W + 1) / 2, 32) * ocrs_cdiv((H + 1) / 2, 8); }

// ocrs_det_pool2_bwd + the BatchNorm-backward sums (sum dz, sum dz*yhat) of the block that produced x
// (mean / invstd: that block's batch statistics): replaces ocrs_bnrelu_bwd_reduce for it.
int ocrs_det_pool2_bwd_bn(const float* x, long long x_ss, int N, int C, int H, int W, const float* sc,
                          const float* sh, const float* lo, const float* dout, long long dout_ss, float* din,
                          long long din_ss, const float* mean, const float* invstd, float* bn_partials,
                          void* stream) {
  OCRS_CHECK_ARG(sc && mean && invstd && bn_partials, "pool2_bwd_bn: needs the producer's transform and statistics");
  dim3 block(32, 8), grid(ocrs_cdiv((W + 1) / 2, 32), ocrs_cdiv((H + 1) / 2, 8), N * C);
  pool2_bwd_bn_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, x_ss, C, H, W, sc, sh, lo, dout, dout_ss, din,
                                                                din_ss, mean, invstd, bn_partials);
  OCRS_CHECK_LAUNCH("pool2_bwd_bn_kernel");
  return 0;
}

// Blocks (= partial rows) of ocrs_det_outconv8_bwd_bn.
int ocrs_det_outconv8_bwd_blocks(void) { return 4 * OCRS_NUM_SMS; }

// out_conv backward for the 8 -> 1 head (reference models.py:126-129), persistent + vectorised, HW % 4 == 0:
// d_a, weight/bias partials [blocks][9] and, when mean/invstd/bn_partials are given, the BatchNorm-backward sums
// [blocks][2][8] of the block that produced x.
int ocrs_det_outconv8_bwd_bn(const float* dp, const float* prob, const float* x, long long x_ss, int N, long long HW,
                             const float* sc, const float* sh, const float* lo, const float* w, float* d_a,
                             long long da_ss, const float* mean, const float* invstd, float* wb_partials,
                             float* bn_partials, void* stream) {
  OCRS_CHECK_ARG(HW % 4 == 0 && x_ss % 4 == 0 && da_ss % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)d_a % 16 == 0 &&
                     (uintptr_t)dp % 16 == 0 && (uintptr_t)prob % 16 == 0, "outconv8_bwd_bn: needs 16-byte aligned planes");
  outconv8_bwd_bn_kernel<<<ocrs_det_outconv8_bwd_blocks(), 256, 0, (cudaStream_t)stream>>>(
      dp, prob, x, x_ss, HW, N, sc, sh, lo, w, d_a, da_ss, mean, invstd, wb_partials, bn_partials);
  OCRS_CHECK_LAUNCH("outconv8_bwd_bn_kernel");
  return 0;
}

int ocrs_det_convt_bwd_data(const float* dout, long long dout_ss, int N, int Cout, int Hs, int Ws,
                            const float* w, int Cin, int Hin, int Win, float* dx, long long dx_ss,
                            void* stream) {
  dim3 block(32, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (ocrs_convt_mma_bwd(dout, dout_ss, N, Cout, Hs, Ws, w, Cin, Hin, Win, dx, dx_ss, st)) {  // csrc/det_convt.cu
    OCRS_CHECK_LAUNCH("convt_bwd_mma_kernel");
    return 0;
  }
  if (Cin <= 8) {
    dim3 grid(ocrs_cdiv(Win, 32), ocrs_cdiv(Hin, 8), N * ocrs_cdiv(Cin, 8));
    convt_bwd_data_kernel<8><<<grid, block, 0, st>>>(dout, dout_ss, Cout, Hs, Ws, w, Cin, Hin, Win, dx, dx_ss);
  } else {
    dim3 grid(ocrs_cdiv(Win, 32), ocrs_cdiv(Hin, 8), N * ocrs_cdiv(Cin, 16));
    convt_bwd_data_kernel<16><<<grid, block, 0, st>>>(dout, dout_ss, Cout, Hs, Ws, w, Cin, Hin, Win, dx, dx_ss);
  }
  OCRS_CHECK_LAUNCH("convt_bwd_data_kernel");
  return 0;
}

int ocrs_det_convt_wgrad_workers(int N, int Hin, int Win) {
  const long long blocks = ((long long)N * Hin * Win + 255) / 256;
  return (int)(blocks < 2 * OCRS_NUM_SMS ? blocks : 2 * OCRS_NUM_SMS);
}
// partials: [workers][Cin][Cout][9]
int ocrs_det_convt_wgrad(const float* x, long long x_ss, int N, int Cin, int Hin, int Win,
                         const float* isc, const float* ish, const float* ilo, const float* dout,
                         long long dout_ss, int Cout, int Hs, int Ws, float* partials, void* stream) {
  dim3 grid(ocrs_det_convt_wgrad_workers(N, Hin, Win), ocrs_cdiv(Cin, 16) * ocrs_cdiv(Cout * 9, 16));
  convt_wgrad_mma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_ss, Cin, Hin, Win, isc, ish, ilo, dout,
                                                                   dout_ss, Cout, Hs, Ws, N, partials);
  OCRS_CHECK_LAUNCH("convt_wgrad_kernel");
  return 0;
}

// partials: [ocrs_reduce_rows(N, HW)][C]
int ocrs_plane_sum(const float* v, long long v_ss, int N, int C, long long HW, float* partials,
                   void* stream) {
  dim3 grid(ocrs_cdiv(HW, RED_CHUNK), C, N);
  plane_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(v, v_ss, C, HW, partials);
  OCRS_CHECK_LAUNCH("plane_sum_kernel");
  return 0;
}

}  // extern "C"

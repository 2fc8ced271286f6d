// Forward kernels of the text-detection U-Net (reference ocrs_models/models.py:7-143).
//
// Data layout in HBM: planar NCHW fp32 "views" = (base pointer, per-sample stride in elements);
// the channel stride is always H*W, so a view can address a channel range of a wider concat
// buffer (torch.cat at models.py:89 is never materialised as a copy). A block's BatchNorm+ReLU
// is not applied by its producer: the producer writes the raw 1x1-conv output plus per-block
// (sum, sum^2) partials, bn_finalize turns them into per-channel (scale, shift, lo), and every
// consumer applies  max(v*scale+shift, lo)  while loading (ChanXform).
#include "common.cuh"
#include <math.h>

namespace {

constexpr int TW = 32, PPT = 4, TH = 8 * PPT;  // 32x32 pixel tile per 256-thread block
constexpr int CI_CHUNK = 8;
constexpr int SROW = TW + 2;
constexpr int SPLANE = (TH + 2) * SROW;

// ---------------------------------------------------------------------------------------------
// DepthwiseConv block body: y = pw1x1(dw3x3(xform(x))), + BN partial statistics of y.
// models.py:11-22 (both convolutions bias-free, dw padding 1).
template <int CO_T>
__global__ void __launch_bounds__(256, 2)
dwpw_fwd_kernel(const float* __restrict__ x, long long x_ss, int Cin, int H, int W,
                const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                const float* __restrict__ in_lo, const float* __restrict__ wdw,
                const float* __restrict__ wpw, int Cout, float* __restrict__ y, long long y_ss,
                float* __restrict__ partials, int tiles_x) {
  __shared__ float xs[CI_CHUNK * SPLANE];
  __shared__ float sdw[CI_CHUNK * 9];
  __shared__ __align__(16) float spw[CI_CHUNK * CO_T];
  __shared__ float sred[8][2 * CO_T];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int tile = blockIdx.x;
  const int x0 = (tile % tiles_x) * TW, y0 = (tile / tiles_x) * TH;
  const int co0 = blockIdx.y * CO_T, n = blockIdx.z;
  const size_t HW = (size_t)H * W;
  const float* xn = x + (size_t)n * x_ss;

  float acc[PPT][CO_T];
#pragma unroll
  for (int p = 0; p < PPT; ++p)
#pragma unroll
    for (int o = 0; o < CO_T; ++o) acc[p][o] = 0.f;

  // Per-thread tile slots: position (ry, rx) of the haloed tile, fixed for every channel, so the
  // per-element address arithmetic is done once per tile and the loads of a chunk are issued as
  // independent batches (memory-level parallelism instead of a dependent load->store chain).
  constexpr int NSLOT = (SPLANE + 255) / 256;
  int goff[NSLOT];
#pragma unroll
  for (int sl = 0; sl < NSLOT; ++sl) {
    const int pos = tid + 256 * sl;
    const int ry = pos / SROW, rx = pos - ry * SROW;
    const int gy = y0 + ry - 1, gx = x0 + rx - 1;
    goff[sl] = (pos < SPLANE && gy >= 0 && gy < H && gx >= 0 && gx < W) ? gy * W + gx : -1;
  }
  for (int ci0 = 0; ci0 < Cin; ci0 += CI_CHUNK) {
    const int nci = min(CI_CHUNK, Cin - ci0);
    __syncthreads();
#pragma unroll
    for (int sl = 0; sl < NSLOT; ++sl) {
      const int pos = tid + 256 * sl;
      if (pos < SPLANE) {
        float v[CI_CHUNK];
#pragma unroll
        for (int c = 0; c < CI_CHUNK; ++c)
          v[c] = (c < nci && goff[sl] >= 0) ? xn[(size_t)(ci0 + c) * HW + goff[sl]] : 0.f;
        if (in_scale && goff[sl] >= 0) {
#pragma unroll
          for (int c = 0; c < CI_CHUNK; ++c)
            if (c < nci) v[c] = xform_apply(v[c], in_scale[ci0 + c], in_shift[ci0 + c], in_lo[ci0 + c]);
        }
#pragma unroll
        for (int c = 0; c < CI_CHUNK; ++c) xs[c * SPLANE + pos] = v[c];
      }
    }
    for (int i = tid; i < nci * 9; i += 256) sdw[i] = wdw[(size_t)ci0 * 9 + i];
    for (int i = tid; i < nci * CO_T; i += 256) {
      const int c = i / CO_T, o = i - c * CO_T;
      spw[i] = (co0 + o < Cout) ? wpw[(size_t)(co0 + o) * Cin + ci0 + c] : 0.f;
    }
    __syncthreads();
    for (int c = 0; c < nci; ++c) {
      const float* t = xs + c * SPLANE + (ty * PPT) * SROW + tx;
      float w[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) w[k] = sdw[c * 9 + k];
      float r[PPT + 2][3];
#pragma unroll
      for (int rr = 0; rr < PPT + 2; ++rr)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) r[rr][cc] = t[rr * SROW + cc];
      float d[PPT];
#pragma unroll
      for (int p = 0; p < PPT; ++p) {
        float s = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) s = fmaf(r[p + ky][kx], w[ky * 3 + kx], s);
        d[p] = s;
      }
      const float4* w4 = reinterpret_cast<const float4*>(spw + c * CO_T);
#pragma unroll
      for (int o4 = 0; o4 < CO_T / 4; ++o4) {
        const float4 wv = w4[o4];
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
          acc[p][o4 * 4 + 0] = fmaf(d[p], wv.x, acc[p][o4 * 4 + 0]);
          acc[p][o4 * 4 + 1] = fmaf(d[p], wv.y, acc[p][o4 * 4 + 1]);
          acc[p][o4 * 4 + 2] = fmaf(d[p], wv.z, acc[p][o4 * 4 + 2]);
          acc[p][o4 * 4 + 3] = fmaf(d[p], wv.w, acc[p][o4 * 4 + 3]);
        }
      }
    }
  }

  float* yn = y + (size_t)n * y_ss;
  const int gx = x0 + tx;
  float s1[CO_T], s2[CO_T];
#pragma unroll
  for (int o = 0; o < CO_T; ++o) { s1[o] = 0.f; s2[o] = 0.f; }
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int gy = y0 + ty * PPT + p;
    if (gy < H && gx < W) {
#pragma unroll
      for (int o = 0; o < CO_T; ++o) {
        if (co0 + o < Cout) {
          const float v = acc[p][o];
          yn[(size_t)(co0 + o) * HW + (size_t)gy * W + gx] = v;
          s1[o] += v;
          s2[o] = fmaf(v, v, s2[o]);
        }
      }
    }
  }
  if (partials) {
#pragma unroll
    for (int o = 0; o < CO_T; ++o) {
      const float a = warp_sum(s1[o]), b = warp_sum(s2[o]);
      if (tx == 0) { sred[ty][o] = a; sred[ty][CO_T + o] = b; }
    }
    __syncthreads();
    if (tid < 2 * CO_T) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += sred[w][tid];
      const int o = tid % CO_T, which = tid / CO_T;
      if (co0 + o < Cout) {
        const size_t blk = (size_t)n * gridDim.x + tile;
        partials[blk * 2 * Cout + (size_t)which * Cout + co0 + o] = s;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm2d statistics -> consumer transform. nn.BatchNorm2d defaults (models.py:23): eps 1e-5,
// momentum 0.1, biased batch variance for normalisation, unbiased for running_var.
__global__ void bn_finalize_kernel(const float* __restrict__ partials, int nblk, int C,
                                   double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean,
                                   float* running_var, float momentum, float eps, int training,
                                   int relu, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ lo, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, long long* num_batches_tracked) {
  __shared__ double red[2][32];
  const int c = blockIdx.x;
  if (training && num_batches_tracked && c == 0 && threadIdx.x == 0) num_batches_tracked[0] += 1;
  double mean, var;
  if (training) {
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
    mean = a / count;
    var = fmax(b / count - mean * mean, 0.0);
    const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
  } else {
    if (threadIdx.x != 0) return;
    mean = running_mean[c];
    var = running_var[c];
  }
  const double invstd = 1.0 / sqrt(var + (double)eps);
  const double sc = (double)gamma[c] * invstd;
  scale[c] = (float)sc;
  shift[c] = (float)((double)beta[c] - mean * sc);
  lo[c] = relu ? 0.f : -INFINITY;
  mean_out[c] = (float)mean;
  invstd_out[c] = (float)invstd;
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d(2) over the activated tensor (models.py:54), floor semantics.
__global__ void pool2_fwd_kernel(const float* __restrict__ x, long long x_ss, int C, int H, int W,
                                 const float* __restrict__ sc, const float* __restrict__ sh,
                                 const float* __restrict__ lo, float* __restrict__ out,
                                 long long out_ss, int Ho, int Wo) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y * blockDim.y + threadIdx.y;
  const int c = blockIdx.z % C, n = blockIdx.z / C;
  if (ox >= Wo || oy >= Ho) return;
  const float* p = x + (size_t)n * x_ss + (size_t)c * H * W + (size_t)(2 * oy) * W + 2 * ox;
  float a = p[0], b = p[1], d = p[W], e = p[W + 1];
  if (sc) {
    const float s = sc[c], t = sh[c], l = lo[c];
    a = xform_apply(a, s, t, l); b = xform_apply(b, s, t, l);
    d = xform_apply(d, s, t, l); e = xform_apply(e, s, t, l);
  }
  float m = a;
  if (b > m || isnan(b)) m = b;
  if (d > m || isnan(d)) m = d;
  if (e > m || isnan(e)) m = e;
  out[(size_t)n * out_ss + (size_t)c * Ho * Wo + (size_t)oy * Wo + ox] = m;
}

// ---------------------------------------------------------------------------------------------
// ConvTranspose2d(k=3, stride=2) + bias, cropped to (Hs, Ws), written at a channel offset of the
// concat buffer (models.py:76-89). Thread = one 2x2 output quad = input pixel (qy, qx).
constexpr int CT_CHUNK = 8;
template <int CO_T>
__global__ void __launch_bounds__(256)
convt_fwd_kernel(const float* __restrict__ x, long long x_ss, int Cin, int Hin, int Win,
                 const float* __restrict__ sc, const float* __restrict__ sh,
                 const float* __restrict__ lo, const float* __restrict__ w /*[Cin][Cout][3][3]*/,
                 const float* __restrict__ bias, int Cout, float* __restrict__ out,
                 long long out_ss, int Hs, int Ws) {
  __shared__ __align__(16) float sw[CT_CHUNK * 9 * CO_T];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int qx = blockIdx.x * 32 + threadIdx.x, qy = blockIdx.y * 8 + threadIdx.y;
  const int cog = (Cout + CO_T - 1) / CO_T;
  const int co0 = (blockIdx.z % cog) * CO_T, n = blockIdx.z / cog;
  const float* xn = x + (size_t)n * x_ss;
  const size_t HWi = (size_t)Hin * Win;
  float o00[CO_T], o01[CO_T], o10[CO_T], o11[CO_T];
#pragma unroll
  for (int o = 0; o < CO_T; ++o) { o00[o] = 0.f; o01[o] = 0.f; o10[o] = 0.f; o11[o] = 0.f; }
  const bool yin = qy < Hin, xin = qx < Win, ym = qy >= 1 && qy - 1 < Hin, xm = qx >= 1 && qx - 1 < Win;
  for (int ci0 = 0; ci0 < Cin; ci0 += CT_CHUNK) {
    const int nci = min(CT_CHUNK, Cin - ci0);
    __syncthreads();
    for (int i = tid; i < nci * 9 * CO_T; i += 256) {
      const int c = i / (9 * CO_T), r = i - c * 9 * CO_T, k = r / CO_T, o = r - k * CO_T;
      sw[i] = (co0 + o < Cout) ? w[((size_t)(ci0 + c) * Cout + co0 + o) * 9 + k] : 0.f;
    }
    __syncthreads();
    for (int c = 0; c < nci; ++c) {
      const float* xp = xn + (size_t)(ci0 + c) * HWi;
      float x00 = (yin && xin) ? xp[(size_t)qy * Win + qx] : 0.f;
      float x10 = (ym && xin) ? xp[(size_t)(qy - 1) * Win + qx] : 0.f;
      float x01 = (yin && xm) ? xp[(size_t)qy * Win + qx - 1] : 0.f;
      float x11 = (ym && xm) ? xp[(size_t)(qy - 1) * Win + qx - 1] : 0.f;
      if (sc) {
        const float s = sc[ci0 + c], t = sh[ci0 + c], l = lo[ci0 + c];
        x00 = (yin && xin) ? xform_apply(x00, s, t, l) : 0.f;
        x10 = (ym && xin) ? xform_apply(x10, s, t, l) : 0.f;
        x01 = (yin && xm) ? xform_apply(x01, s, t, l) : 0.f;
        x11 = (ym && xm) ? xform_apply(x11, s, t, l) : 0.f;
      }
      const float* wc = sw + c * 9 * CO_T;
#pragma unroll
      for (int o = 0; o < CO_T; ++o) {
        // taps (ky,kx): index ky*3+kx
        o00[o] = fmaf(x00, wc[0 * CO_T + o], o00[o]);
        o00[o] = fmaf(x10, wc[6 * CO_T + o], o00[o]);
        o00[o] = fmaf(x01, wc[2 * CO_T + o], o00[o]);
        o00[o] = fmaf(x11, wc[8 * CO_T + o], o00[o]);
        o01[o] = fmaf(x00, wc[1 * CO_T + o], o01[o]);
        o01[o] = fmaf(x10, wc[7 * CO_T + o], o01[o]);
        o10[o] = fmaf(x00, wc[3 * CO_T + o], o10[o]);
        o10[o] = fmaf(x01, wc[5 * CO_T + o], o10[o]);
        o11[o] = fmaf(x00, wc[4 * CO_T + o], o11[o]);
      }
    }
  }
  const int oy = 2 * qy, ox = 2 * qx;
  float* on = out + (size_t)n * out_ss;
  const size_t HWs = (size_t)Hs * Ws;
#pragma unroll
  for (int o = 0; o < CO_T; ++o) {
    if (co0 + o >= Cout) continue;
    const float b = bias ? bias[co0 + o] : 0.f;
    float* oc = on + (size_t)(co0 + o) * HWs;
    if (oy < Hs && ox < Ws) oc[(size_t)oy * Ws + ox] = o00[o] + b;
    if (oy < Hs && ox + 1 < Ws) oc[(size_t)oy * Ws + ox + 1] = o01[o] + b;
    if (oy + 1 < Hs && ox < Ws) oc[(size_t)(oy + 1) * Ws + ox] = o10[o] + b;
    if (oy + 1 < Hs && ox + 1 < Ws) oc[(size_t)(oy + 1) * Ws + ox + 1] = o11[o] + b;
  }
}

// ---------------------------------------------------------------------------------------------
// out_conv: Conv2d(C -> 1, k=1) + bias + Sigmoid (models.py:126-129).
__global__ void outconv_fwd_kernel(const float* __restrict__ x, long long x_ss, int C, long long HW,
                                   const float* __restrict__ sc, const float* __restrict__ sh,
                                   const float* __restrict__ lo, const float* __restrict__ w,
                                   const float* __restrict__ bias, float* __restrict__ prob) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (i >= HW) return;
  const float* xn = x + (size_t)n * x_ss + i;
  float z = bias[0];
  for (int c = 0; c < C; ++c) {
    float v = xn[(size_t)c * HW];
    if (sc) v = xform_apply(v, sc[c], sh[c], lo[c]);
    z = fmaf(v, w[c], z);
  }
  prob[(size_t)n * HW + i] = 1.f / (1.f + expf(-z));
}

// ---------------------------------------------------------------------------------------------
// Depthwise 3x3 alone (levels with >= 64 channels, where the 1x1 contraction runs as a batched tcgen05 GEMM):
// out[n][c][p] = dw3x3(xform(x))[n][c][p], contiguous [N][C][H][W].
__global__ void __launch_bounds__(256)
dw3x3_fwd_kernel(const float* __restrict__ x, long long x_ss, int C, int H, int W, const float* __restrict__ sc,
                 const float* __restrict__ sh, const float* __restrict__ lo, const float* __restrict__ wdw,
                 float* __restrict__ out) {
  const int px = blockIdx.x * 32 + threadIdx.x, py = blockIdx.y * 8 + threadIdx.y;
  const int c = blockIdx.z % C, n = blockIdx.z / C;
  if (px >= W || py >= H) return;
  const float* xp = x + (size_t)n * x_ss + (size_t)c * H * W;
  float s = 1.f, t = 0.f, l = -INFINITY;
  if (sc) { s = sc[c]; t = sh[c]; l = lo[c]; }
  float acc = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = py + ky - 1, xx = px + kx - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W)
        acc = fmaf(xform_apply(xp[(size_t)yy * W + xx], s, t, l), wdw[(size_t)c * 9 + ky * 3 + kx], acc);
    }
  out[((size_t)n * C + c) * H * W + (size_t)py * W + px] = acc;
}

// Same, four consecutive pixels of a row per thread (W % 4 == 0, 16-byte aligned planes): one float4 + two scalars per
// input row instead of twelve scalar loads.
__global__ void __launch_bounds__(256)
dw3x3_fwd_vec4_kernel(const float* __restrict__ x, long long x_ss, int C, int H, int W, const float* __restrict__ sc,
                      const float* __restrict__ sh, const float* __restrict__ lo, const float* __restrict__ wdw,
                      float* __restrict__ out) {
  const int wq = W >> 2, idx = blockIdx.x * 256 + threadIdx.x;  // rows are short at these levels: flat (row, quad) index
  const int py = idx / wq, px = (idx - py * wq) * 4;
  const int c = blockIdx.y, n = blockIdx.z;
  if (py >= H) return;
  const float* xp = x + (size_t)n * x_ss + (size_t)c * H * W;
  float s = 1.f, t = 0.f, l = -INFINITY;
  if (sc) { s = sc[c]; t = sh[c]; l = lo[c]; }
  float wk[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) wk[k] = wdw[(size_t)c * 9 + k];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = py + ky - 1;
    if (yy < 0 || yy >= H) continue;
    const float* r = xp + (size_t)yy * W + px;
    const float4 m = *reinterpret_cast<const float4*>(r);
    float v[6];
    v[0] = px > 0 ? xform_apply(r[-1], s, t, l) : 0.f;
    v[1] = xform_apply(m.x, s, t, l); v[2] = xform_apply(m.y, s, t, l);
    v[3] = xform_apply(m.z, s, t, l); v[4] = xform_apply(m.w, s, t, l);
    v[5] = px + 4 < W ? xform_apply(r[4], s, t, l) : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) acc[q] = fmaf(v[q + kx], wk[ky * 3 + kx], acc[q]);
  }
  *reinterpret_cast<float4*>(out + ((size_t)n * C + c) * H * W + (size_t)py * W + px) =
      make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// ---------------------------------------------------------------------------------------------
// ConvTranspose2d as a GEMM (levels with >= 64 input channels): the activated input, contiguous [N][C][HW], is the
// B operand of Z[n] = W^T x[n] (Z rows = (co, ky, kx), ocrs_gemm_tc_batched); col2im then gathers the 1, 2 or 4 taps
// that land on every output pixel (stride 2, 3x3: out[2qy + ky][2qx + kx] += Z[(co, ky, kx)][qy][qx]) and adds the bias.
__global__ void __launch_bounds__(256)
activate_kernel(const float* __restrict__ x, long long x_ss, int C, long long HW, const float* __restrict__ sc,
                const float* __restrict__ sh, const float* __restrict__ lo, float* __restrict__ out) {
  const long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  const int c = blockIdx.y, n = blockIdx.z;
  if (i >= HW) return;
  const float* xp = x + (size_t)n * x_ss + (size_t)c * HW + i;
  float* op = out + ((size_t)n * C + c) * HW + i;
  float s = 1.f, t = 0.f, l = -INFINITY;
  if (sc) { s = sc[c]; t = sh[c]; l = lo[c]; }
  if (i + 4 <= HW && (((uintptr_t)xp | (uintptr_t)op) & 15) == 0) {
    float4 v = *reinterpret_cast<const float4*>(xp);
    v.x = xform_apply(v.x, s, t, l); v.y = xform_apply(v.y, s, t, l);
    v.z = xform_apply(v.z, s, t, l); v.w = xform_apply(v.w, s, t, l);
    *reinterpret_cast<float4*>(op) = v;
  } else {
    for (int q = 0; q < 4 && i + q < HW; ++q) op[q] = xform_apply(xp[q], s, t, l);
  }
}

__global__ void __launch_bounds__(256)
convt_col2im_kernel(const float* __restrict__ Z /*[N][Cout*9][Hin*Win]*/, int Cout, int Hin, int Win,
                    const float* __restrict__ bias, float* __restrict__ out, long long out_ss, int Hs, int Ws) {
  // thread = input position (qy, qx) = the 2x2 output quad (2qy.., 2qx..): nine row-contiguous reads of Z, two 8-byte stores.
  // qy == Hin / qx == Win (odd crops) only receive the ky == 2 / kx == 2 taps of the last input row / column.
  const int qx = blockIdx.x * 32 + threadIdx.x, qy = blockIdx.y * 8 + threadIdx.y;
  const int co = blockIdx.z % Cout, n = blockIdx.z / Cout;
  const int oy = 2 * qy, ox = 2 * qx;
  if (ox >= Ws || oy >= Hs) return;
  const size_t HWi = (size_t)Hin * Win;
  const float* zc = Z + ((size_t)n * Cout + co) * 9 * HWi;
  const bool y0 = qy < Hin, y1 = qy >= 1, x0 = qx < Win, x1 = qx >= 1;  // (qy, qx), (qy-1, .), (., qx-1) inside the input
  const size_t p = (size_t)qy * Win + qx;
  auto z = [&](int k, bool ok, size_t at) { return ok ? zc[(size_t)k * HWi + at] : 0.f; };
  const float b = bias ? bias[co] : 0.f;
  const float o00 = b + z(0, y0 && x0, p) + z(6, y1 && x0, p - Win) + z(2, y0 && x1, p - 1) + z(8, y1 && x1, p - Win - 1);
  const float o01 = b + z(1, y0 && x0, p) + z(7, y1 && x0, p - Win);
  const float o10 = b + z(3, y0 && x0, p) + z(5, y0 && x1, p - 1);
  const float o11 = b + z(4, y0 && x0, p);
  float* o = out + (size_t)n * out_ss + (size_t)co * Hs * Ws + (size_t)oy * Ws + ox;
  const bool pair = ox + 1 < Ws, vec = pair && ((reinterpret_cast<uintptr_t>(o) & 7) == 0) && (Ws % 2 == 0);
  if (vec) *reinterpret_cast<float2*>(o) = make_float2(o00, o01);
  else { o[0] = o00; if (pair) o[1] = o01; }
  if (oy + 1 < Hs) {
    if (vec) *reinterpret_cast<float2*>(o + Ws) = make_float2(o10, o11);
    else { o[Ws] = o10; if (pair) o[Ws + 1] = o11; }
  }
}

}  // namespace

bool ocrs_convt_mma_fwd(const float* x, long long x_ss, int N, int Cin, int Hin, int Win, const float* sc, const float* sh,
                        const float* lo, const float* w, const float* bias, int Cout, float* out, long long out_ss, int Hs,
                        int Ws, cudaStream_t st);  // csrc/det_convt.cu

extern "C" {

// a[n][c][p] = max(x * scale + shift, lo) (the producer's folded BatchNorm + ReLU) written contiguously as [N][C][HW]:
// the activated input of ConvTranspose2d (reference models.py:76-78) as a GEMM operand.
int ocrs_det_activate(const float* x, long long x_ss, int N, int C, long long HW, const float* sc, const float* sh,
                      const float* lo, float* out, void* stream) {
  OCRS_CHECK_ARG(N > 0 && C > 0 && HW > 0, "activate: bad dims");
  dim3 grid(ocrs_cdiv(HW, 1024), C, N);
  activate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_ss, C, HW, sc, sh, lo, out);
  OCRS_CHECK_LAUNCH("activate_kernel");
  return 0;
}

// Second half of ConvTranspose2d(k=3, stride=2) + bias as a GEMM: Z [N][Cout*9][Hin*Win] (row = (co, ky, kx), the layout
// of the [Cin][Cout][3][3] weight read as a [Cin][9*Cout] matrix) -> out view [Cout][Hs][Ws] (crop of the 2*Hin+1 output).
int ocrs_det_convt_col2im(const float* Z, int N, int Cout, int Hin, int Win, const float* bias, float* out,
                          long long out_ss, int Hs, int Ws, void* stream) {
  OCRS_CHECK_ARG(N > 0 && Cout > 0 && Hin > 0 && Win > 0, "convt_col2im: bad dims");
  OCRS_CHECK_ARG(Hs <= 2 * Hin + 1 && Ws <= 2 * Win + 1, "convt_col2im: crop %dx%d exceeds %dx%d", Hs, Ws, 2 * Hin + 1,
                 2 * Win + 1);
  dim3 block(32, 8), grid(ocrs_cdiv((Ws + 1) / 2, 32), ocrs_cdiv((Hs + 1) / 2, 8), N * Cout);
  convt_col2im_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(Z, Cout, Hin, Win, bias, out, out_ss, Hs, Ws);
  OCRS_CHECK_LAUNCH("convt_col2im_kernel");
  return 0;
}

// Depthwise 3x3 (pad 1, bias-free, reference models.py:12-17) of the activated input, written contiguously as
// [N][C][H][W]: the B operand of ocrs_gemm_tc_batched for the 1x1 convolution of the levels with >= 64 channels.
int ocrs_det_dw3x3_fwd(const float* x, long long x_ss, int N, int C, int H, int W, const float* sc, const float* sh,
                       const float* lo, const float* wdw, float* out, void* stream) {
  OCRS_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0, "dw3x3_fwd: bad dims");
  dim3 block(32, 8);
  if (W % 4 == 0 && x_ss % 4 == 0 && (((uintptr_t)x | (uintptr_t)out) & 15) == 0) {
    dim3 grid(ocrs_cdiv((long long)H * (W / 4), 256), C, N);
    dw3x3_fwd_vec4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_ss, C, H, W, sc, sh, lo, wdw, out);
  } else {
    dim3 grid(ocrs_cdiv(W, 32), ocrs_cdiv(H, 8), N * C);
    dw3x3_fwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, x_ss, C, H, W, sc, sh, lo, wdw, out);
  }
  OCRS_CHECK_LAUNCH("dw3x3_fwd_kernel");
  return 0;
}

// Number of per-block statistic partial rows ocrs_det_dwpw_fwd writes ([rows][2][Cout] floats).
int ocrs_det_dwpw_partial_rows(int N, int H, int W) {
  return N * ocrs_cdiv(W, TW) * ocrs_cdiv(H, TH);
}

int ocrs_det_dwpw_fwd(const float* x, long long x_ss, int N, int Cin, int H, int W,
                      const float* in_scale, const float* in_shift, const float* in_lo,
                      const float* wdw, const float* wpw, int Cout, float* y, long long y_ss,
                      float* partials, void* stream) {
  OCRS_CHECK_ARG(N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "dwpw_fwd: bad dims");
  const int tiles_x = ocrs_cdiv(W, TW), tiles_y = ocrs_cdiv(H, TH);
  cudaStream_t st = (cudaStream_t)stream;
  if (Cout <= 8) {
    dim3 grid(tiles_x * tiles_y, ocrs_cdiv(Cout, 8), N);
    dwpw_fwd_kernel<8><<<grid, 256, 0, st>>>(x, x_ss, Cin, H, W, in_scale, in_shift, in_lo, wdw,
                                             wpw, Cout, y, y_ss, partials, tiles_x);
  } else {
    dim3 grid(tiles_x * tiles_y, ocrs_cdiv(Cout, 16), N);
    dwpw_fwd_kernel<16><<<grid, 256, 0, st>>>(x, x_ss, Cin, H, W, in_scale, in_shift, in_lo, wdw,
                                              wpw, Cout, y, y_ss, partials, tiles_x);
  }
  OCRS_CHECK_LAUNCH("dwpw_fwd_kernel");
  return 0;
}

// BatchNorm2d statistics from (sum, sum of squares) partial rows -> the consumers' folded (scale, shift, lo) transform,
// mean / invstd for the backward pass, running statistics and (training, optional) num_batches_tracked += 1
// (reference models.py:18-20: nn.BatchNorm2d in train or eval mode).
int ocrs_bn_finalize(const float* partials, int nblk, int C, double count, const float* gamma,
                     const float* beta, float* running_mean, float* running_var, float momentum,
                     float eps, int training, int relu, float* scale, float* shift, float* lo,
                     float* mean_out, float* invstd_out, long long* num_batches_tracked, void* stream) {
  OCRS_CHECK_ARG(C > 0, "bn_finalize: bad channel count");
  bn_finalize_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(
      partials, nblk, C, count, gamma, beta, running_mean, running_var, momentum, eps, training,
      relu, scale, shift, lo, mean_out, invstd_out, num_batches_tracked);
  OCRS_CHECK_LAUNCH("bn_finalize_kernel");
  return 0;
}

int ocrs_det_pool2_fwd(const float* x, long long x_ss, int N, int C, int H, int W, const float* sc,
                       const float* sh, const float* lo, float* out, long long out_ss,
                       void* stream) {
  const int Ho = H / 2, Wo = W / 2;
  OCRS_CHECK_ARG(Ho > 0 && Wo > 0, "pool2_fwd: input %dx%d too small", H, W);
  dim3 block(32, 8), grid(ocrs_cdiv(Wo, 32), ocrs_cdiv(Ho, 8), N * C);
  pool2_fwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, x_ss, C, H, W, sc, sh, lo, out,
                                                             out_ss, Ho, Wo);
  OCRS_CHECK_LAUNCH("pool2_fwd_kernel");
  return 0;
}

int ocrs_det_convt_fwd(const float* x, long long x_ss, int N, int Cin, int Hin, int Win,
                       const float* sc, const float* sh, const float* lo, const float* w,
                       const float* bias, int Cout, float* out, long long out_ss, int Hs, int Ws,
                       void* stream) {
  OCRS_CHECK_ARG(Hs <= 2 * Hin + 1 && Ws <= 2 * Win + 1, "convt_fwd: crop %dx%d exceeds %dx%d", Hs,
                 Ws, 2 * Hin + 1, 2 * Win + 1);
  const int QH = (Hs + 1) / 2, QW = (Ws + 1) / 2;
  dim3 block(32, 8);
  cudaStream_t st = (cudaStream_t)stream;
  // 16-32 channel levels on TMA-friendly shapes: tensor-core tile kernel (csrc/det_convt.cu)
  if (ocrs_convt_mma_fwd(x, x_ss, N, Cin, Hin, Win, sc, sh, lo, w, bias, Cout, out, out_ss, Hs, Ws, st)) {
    OCRS_CHECK_LAUNCH("convt_fwd_mma_kernel");
    return 0;
  }
  if (Cout <= 8) {
    dim3 grid(ocrs_cdiv(QW, 32), ocrs_cdiv(QH, 8), N * ocrs_cdiv(Cout, 8));
    convt_fwd_kernel<8><<<grid, block, 0, st>>>(x, x_ss, Cin, Hin, Win, sc, sh, lo, w, bias, Cout,
                                                out, out_ss, Hs, Ws);
  } else {
    dim3 grid(ocrs_cdiv(QW, 32), ocrs_cdiv(QH, 8), N * ocrs_cdiv(Cout, 16));
    convt_fwd_kernel<16><<<grid, block, 0, st>>>(x, x_ss, Cin, Hin, Win, sc, sh, lo, w, bias, Cout,
                                                 out, out_ss, Hs, Ws);
  }
  OCRS_CHECK_LAUNCH("convt_fwd_kernel");
  return 0;
}

int ocrs_det_outconv_fwd(const float* x, long long x_ss, int N, int C, int H, int W,
                         const float* sc, const float* sh, const float* lo, const float* w,
                         const float* bias, float* prob, void* stream) {
  const long long HW = (long long)H * W;
  dim3 grid(ocrs_cdiv(HW, 256), N);
  outconv_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_ss, C, HW, sc, sh, lo, w, bias,
                                                             prob);
  OCRS_CHECK_LAUNCH("outconv_fwd_kernel");
  return 0;
}

}  // extern "C"

// Recognition (CRNN) kernels that are not plain GEMMs (reference ocrs_models/models.py:179-268).
//
// Activations of the conv stack are NHWC fp32. conv.0 (Cin = 1) is a direct convolution fused
// with ReLU + MaxPool2 (models.py:180-187); every other convolution is an implicit GEMM on tcgen05 (gemm_tc.cu; conv.19, 2x2: im2col + GEMM)
// followed by one fused BatchNorm-affine / ReLU / pool kernel here. The GRU recurrence
// (models.py:245) lives in gru_persist.cu (one cluster-persistent launch per layer); LogSoftmax
// (models.py:250) is one warp per row.
#include "common.cuh"
#include <math.h>

namespace {

// ---------------------------------------------------------------------------------------------
// conv.0: Conv2d(1, 32, 3, pad 1) + bias -> ReLU -> MaxPool2d(2). Thread = pooled pixel x 8 channels.
__device__ __forceinline__ void conv0_patch(const float* __restrict__ x, int H, int W, int n, int py,
                                            int px, float (&p)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int iy = 2 * py - 1 + r, ix = 2 * px - 1 + c;
      p[r][c] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? x[((size_t)n * H + iy) * W + ix] : 0.f;
    }
}

__global__ void __launch_bounds__(256)
conv0_fwd_kernel(const float* __restrict__ x, int N, int H, int W, const float* __restrict__ w,
                 const float* __restrict__ bias, float* __restrict__ out, unsigned* __restrict__ code) {
  __shared__ __align__(16) float sw[9 * 32];
  __shared__ float sb[32];
  for (int i = threadIdx.x; i < 288; i += 256) sw[(i % 9) * 32 + i / 9] = w[i];  // [tap][c]
  if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int Hp = H / 2, Wp = W / 2;
  const long long total = (long long)N * Hp * Wp * 4;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int cg = (int)(i & 3);
  long long r = i >> 2;
  const int px = (int)(r % Wp);
  r /= Wp;
  const int py = (int)(r % Hp), n = (int)(r / Hp);
  float p[4][4];
  conv0_patch(x, H, W, n, py, px, p);
  float best[8];
  unsigned arg = 0;  // 2 bits per channel: window position of the FIRST maximum (aten's tie rule), for the backward pass
#pragma unroll
  for (int c = 0; c < 8; ++c) best[c] = -INFINITY;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int dy = q >> 1, dx = q & 1;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = sb[cg * 8 + c];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float xv = p[dy + k / 3][dx + k % 3];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = fmaf(xv, sw[k * 32 + cg * 8 + c], v[c]);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (q == 0) best[c] = v[c];
      else if (v[c] > best[c]) { best[c] = v[c]; arg = (arg & ~(3u << (2 * c))) | ((unsigned)q << (2 * c)); }
    }
  }
  if (code) {  // bits 0-15: arg-max positions, bits 16-23: pre-activation maximum > 0 (the ReLU gate)
    unsigned pos = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) pos |= (best[c] > 0.f ? 1u : 0u) << c;
    code[i] = arg | (pos << 16);
  }
  float4* o = reinterpret_cast<float4*>(out + ((((size_t)n * Hp + py) * Wp + px) * 32 + cg * 8));
  o[0] = make_float4(fmaxf(best[0], 0.f), fmaxf(best[1], 0.f), fmaxf(best[2], 0.f), fmaxf(best[3], 0.f));
  o[1] = make_float4(fmaxf(best[4], 0.f), fmaxf(best[5], 0.f), fmaxf(best[6], 0.f), fmaxf(best[7], 0.f));
}

// Weight/bias gradient of conv.0 (its input is the image: no data gradient). Recomputes the four
// pre-pool values to route d_out to the first maximum (aten tie rule) when it is positive.
// partials: [gridDim.x][32][10] (9 taps + bias).
__global__ void __launch_bounds__(256)
conv0_bwd_kernel(const float* __restrict__ x, int N, int H, int W, const float* __restrict__ w,
                 const float* __restrict__ bias, const float* __restrict__ dout, const unsigned* __restrict__ code,
                 float* __restrict__ partials) {
  __shared__ __align__(16) float sw[9 * 32];
  __shared__ float sb[32];
  __shared__ float red[8][4][80];
  for (int i = threadIdx.x; i < 288; i += 256) sw[(i % 9) * 32 + i / 9] = w[i];
  if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int Hp = H / 2, Wp = W / 2;
  const long long total = (long long)N * Hp * Wp * 4;
  const int cg = threadIdx.x & 3;
  float acc[8][10];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[c][k] = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    long long r = i >> 2;
    const int px = (int)(r % Wp);
    r /= Wp;
    const int py = (int)(r % Hp), n = (int)(r / Hp);
    float p[4][4];
    conv0_patch(x, H, W, n, py, px, p);
    const float4* gp = reinterpret_cast<const float4*>(dout + ((((size_t)n * Hp + py) * Wp + px) * 32 + cg * 8));
    const float4 g0 = gp[0], g1 = gp[1];
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    unsigned cw = 0;
    if (code) {
      cw = code[i];  // arg-max positions and ReLU gates saved by the forward pass: no recompute of the 4 x 9 x 8 products
    } else {
      float v[4][8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int dy = q >> 1, dx = q & 1;
#pragma unroll
        for (int c = 0; c < 8; ++c) v[q][c] = sb[cg * 8 + c];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float xv = p[dy + k / 3][dx + k % 3];
#pragma unroll
          for (int c = 0; c < 8; ++c) v[q][c] = fmaf(xv, sw[k * 32 + cg * 8 + c], v[q][c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        unsigned am = 0;
        float m = v[0][c];
#pragma unroll
        for (int q = 1; q < 4; ++q)
          if (v[q][c] > m) { m = v[q][c]; am = q; }
        cw |= (am << (2 * c)) | ((m > 0.f ? 1u : 0u) << (16 + c));
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int am = (cw >> (2 * c)) & 3;
      const float gv = ((cw >> (16 + c)) & 1u) ? g[c] : 0.f;
      acc[c][9] += gv;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float gq = (q == am) ? gv : 0.f;
        const int dy = q >> 1, dx = q & 1;
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[c][k] = fmaf(gq, p[dy + k / 3][dx + k % 3], acc[c][k]);
      }
    }
  }
  // reduce over lanes with the same channel group (lane & 3), then over warps
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      float s = acc[c][k];
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      if (lane < 4) red[wid][lane][c * 10 + k] = s;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 320; i += 256) {
    const int g4 = i / 80, e = i - g4 * 80;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += red[q][g4][e];
    partials[(size_t)blockIdx.x * 320 + (g4 * 8 + e / 10) * 10 + e % 10] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm affine (+ReLU) + pooling over NHWC. mode 0 = max, 1 = avg. Pool window (ph, pw),
// floor semantics; out element (n, oy, ox, c) at out[n*son + oy*soh + ox*sow + c].
struct PoolGeom {
  int N, H, W, C, ph, pw, Hp, Wp, mode, relu;
  long long son, soh, sow;
};

__device__ __forceinline__ float bn_act(float v, float s, float t, int relu) {
  v = fmaf(v, s, t);
  return relu ? fmaxf(v, 0.f) : v;
}

__global__ void __launch_bounds__(256)
bn_act_pool_fwd_kernel(const float* __restrict__ y, PoolGeom g, const float* __restrict__ sc,
                       const float* __restrict__ sh, float* __restrict__ out) {
  const int C4 = g.C >> 2;
  const long long total = (long long)g.N * g.Hp * g.Wp * C4;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int c4 = (int)(i % C4);
  long long r = i / C4;
  const int ox = (int)(r % g.Wp);
  r /= g.Wp;
  const int oy = (int)(r % g.Hp), n = (int)(r / g.Hp);
  const float4 s = reinterpret_cast<const float4*>(sc)[c4], t = reinterpret_cast<const float4*>(sh)[c4];
  float4 res = g.mode == 0 ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int dy = 0; dy < g.ph; ++dy)
    for (int dx = 0; dx < g.pw; ++dx) {
      const float4 v = *reinterpret_cast<const float4*>(
          y + (((size_t)n * g.H + oy * g.ph + dy) * g.W + ox * g.pw + dx) * g.C + c4 * 4);
      const float a0 = bn_act(v.x, s.x, t.x, g.relu), a1 = bn_act(v.y, s.y, t.y, g.relu);
      const float a2 = bn_act(v.z, s.z, t.z, g.relu), a3 = bn_act(v.w, s.w, t.w, g.relu);
      if (g.mode == 0) {
        res.x = fmaxf(res.x, a0); res.y = fmaxf(res.y, a1); res.z = fmaxf(res.z, a2); res.w = fmaxf(res.w, a3);
      } else {
        res.x += a0; res.y += a1; res.z += a2; res.w += a3;
      }
    }
  if (g.mode == 1) {
    const float inv = 1.f / (float)(g.ph * g.pw);
    res.x *= inv; res.y *= inv; res.z *= inv; res.w *= inv;
  }
  *reinterpret_cast<float4*>(out + (size_t)n * g.son + (size_t)oy * g.soh + (size_t)ox * g.sow + c4 * 4) = res;
}

// Routed gradient of one pooling window for one channel: returns dz at window position q.
// For max: the first maximum takes g (if it passed the ReLU); avg: every position takes g / count.
__device__ __forceinline__ void window_route(const float* v, int cnt, float s, float t, int relu, int mode,
                                             float g, float* dz) {
  if (mode == 1) {
    const float q = g / (float)cnt;
    for (int i = 0; i < cnt; ++i) dz[i] = q;
    return;
  }
  int am = 0;
  float m = bn_act(v[0], s, t, relu);
  for (int i = 1; i < cnt; ++i) {
    const float a = bn_act(v[i], s, t, relu);
    if (a > m) { m = a; am = i; }
  }
  const bool pass = relu ? (fmaf(v[am], s, t) > 0.f) : true;
  for (int i = 0; i < cnt; ++i) dz[i] = (i == am && pass) ? g : 0.f;
}

constexpr int MAXWIN = 8;
// partials [gridDim.x][2][C]: sum dz, sum dz * yhat.
__global__ void __launch_bounds__(256)
bn_act_pool_bwd_reduce_kernel(const float* __restrict__ y, PoolGeom g, const float* __restrict__ sc,
                              const float* __restrict__ sh, const float* __restrict__ mean,
                              const float* __restrict__ invstd, const float* __restrict__ dout,
                              float* __restrict__ partials) {
  extern __shared__ float red[];  // [256/C4][2][C]
  const int C4 = g.C >> 2, lanes = 256 / C4;
  const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4;
  const long long npix = (long long)g.N * g.Hp * g.Wp;
  const int cnt = g.ph * g.pw;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  float s[4], t[4], mu[4], is[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { s[j] = sc[c4 * 4 + j]; t[j] = sh[c4 * 4 + j]; mu[j] = mean[c4 * 4 + j]; is[j] = invstd[c4 * 4 + j]; }
  for (long long p = (long long)blockIdx.x * lanes + pl; p < npix; p += (long long)gridDim.x * lanes) {
    long long r = p;
    const int ox = (int)(r % g.Wp);
    r /= g.Wp;
    const int oy = (int)(r % g.Hp), n = (int)(r / g.Hp);
    const float4 gv = *reinterpret_cast<const float4*>(dout + (size_t)n * g.son + (size_t)oy * g.soh + (size_t)ox * g.sow + c4 * 4);
    const float gj[4] = {gv.x, gv.y, gv.z, gv.w};
    float v[4][MAXWIN];
    for (int dy = 0; dy < g.ph; ++dy)
      for (int dx = 0; dx < g.pw; ++dx) {
        const float4 q = *reinterpret_cast<const float4*>(
            y + (((size_t)n * g.H + oy * g.ph + dy) * g.W + ox * g.pw + dx) * g.C + c4 * 4);
        const int wi = dy * g.pw + dx;
        v[0][wi] = q.x; v[1][wi] = q.y; v[2][wi] = q.z; v[3][wi] = q.w;
      }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float dz[MAXWIN];
      window_route(v[j], cnt, s[j], t[j], g.relu, g.mode, gj[j], dz);
      for (int wi = 0; wi < cnt; ++wi) {
        a[j] += dz[wi];
        b[j] = fmaf(dz[wi], (v[j][wi] - mu[j]) * is[j], b[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[(pl * 2) * g.C + c4 * 4 + j] = a[j];
    red[(pl * 2 + 1) * g.C + c4 * 4 + j] = b[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * g.C; i += 256) {
    float sum = 0.f;
    for (int q = 0; q < lanes; ++q) sum += red[q * 2 * g.C + i];
    partials[(size_t)blockIdx.x * 2 * g.C + i] = sum;
  }
}

// dy[n,h,w,c] = k1*dz + k2*y + k3 for every input position (positions outside any pooling window
// have dz = 0 but still receive the batch-statistics terms).
__global__ void __launch_bounds__(256)
bn_act_pool_bwd_apply_kernel(const float* __restrict__ y, PoolGeom g, const float* __restrict__ sc,
                             const float* __restrict__ sh, const float* __restrict__ k1,
                             const float* __restrict__ k2, const float* __restrict__ k3,
                             const float* __restrict__ dout, float* __restrict__ dy) {
  const int C4 = g.C >> 2;
  const long long total = (long long)g.N * g.H * g.W * C4;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int c4 = (int)(i % C4);
  long long r = i / C4;
  const int ix = (int)(r % g.W);
  r /= g.W;
  const int iy = (int)(r % g.H), n = (int)(r / g.H);
  const int oy = iy / g.ph, ox = ix / g.pw;
  const int me = (iy - oy * g.ph) * g.pw + (ix - ox * g.pw);
  const bool inside = oy < g.Hp && ox < g.Wp;
  const int cnt = g.ph * g.pw;
  const float4 yv = *reinterpret_cast<const float4*>(y + (((size_t)n * g.H + iy) * g.W + ix) * g.C + c4 * 4);
  const float ymine[4] = {yv.x, yv.y, yv.z, yv.w};
  float dzv[4] = {0.f, 0.f, 0.f, 0.f};
  if (inside) {
    const float4 gv = *reinterpret_cast<const float4*>(dout + (size_t)n * g.son + (size_t)oy * g.soh + (size_t)ox * g.sow + c4 * 4);
    const float gj[4] = {gv.x, gv.y, gv.z, gv.w};
    float v[4][MAXWIN];
    for (int dy_ = 0; dy_ < g.ph; ++dy_)
      for (int dx = 0; dx < g.pw; ++dx) {
        const float4 q = *reinterpret_cast<const float4*>(
            y + (((size_t)n * g.H + oy * g.ph + dy_) * g.W + ox * g.pw + dx) * g.C + c4 * 4);
        const int wi = dy_ * g.pw + dx;
        v[0][wi] = q.x; v[1][wi] = q.y; v[2][wi] = q.z; v[3][wi] = q.w;
      }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float dz[MAXWIN];
      window_route(v[j], cnt, sc[c4 * 4 + j], sh[c4 * 4 + j], g.relu, g.mode, gj[j], dz);
      float pick = 0.f;
      for (int wi = 0; wi < cnt; ++wi) pick = (wi == me) ? dz[wi] : pick;
      dzv[j] = pick;
    }
  }
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = fmaf(k1[c4 * 4 + j], dzv[j], fmaf(k2[c4 * 4 + j], ymine[j], k3[c4 * 4 + j]));
  *reinterpret_cast<float4*>(dy + (((size_t)n * g.H + iy) * g.W + ix) * g.C + c4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
}

// ---- compile-time-window versions of the two kernels above (the windows the CRNN uses) ----
// Everything lives in registers (the generic kernels index their window arrays dynamically and spill to
// local memory), and the apply kernel walks pooling CELLS: one thread reads a window once and writes the
// gradient of all its positions, instead of every position re-reading its whole window.
struct F4 { float v[4]; };
__device__ __forceinline__ F4 ld4(const float* p) {
  const float4 q = *reinterpret_cast<const float4*>(p);
  return F4{{q.x, q.y, q.z, q.w}};
}
__device__ __forceinline__ F4 ldg4(const float* p) {
  const float4 q = __ldg(reinterpret_cast<const float4*>(p));
  return F4{{q.x, q.y, q.z, q.w}};
}

// Routed gradient of a full window for one channel: `am` = position that receives `gpass` (max), or -1 = every
// position receives gpass (avg).
template <int CNT, int MODE>
__device__ __forceinline__ void route_static(const float (&v)[CNT], float s, float t, int relu, float g, int& am,
                                             float& gpass) {
  if (MODE == 1) { am = -1; gpass = g / (float)CNT; return; }
  am = 0;
  float pre = fmaf(v[0], s, t);
  float m = relu ? fmaxf(pre, 0.f) : pre;
#pragma unroll
  for (int i = 1; i < CNT; ++i) {
    const float pi = fmaf(v[i], s, t);
    const float a = relu ? fmaxf(pi, 0.f) : pi;
    if (a > m) { m = a; am = i; pre = pi; }
  }
  gpass = (!relu || pre > 0.f) ? g : 0.f;
}

template <int PH, int PW, int MODE>
__global__ void __launch_bounds__(256)
bn_act_pool_bwd_reduce_t(const float* __restrict__ y, PoolGeom g, const float* __restrict__ sc,
                         const float* __restrict__ sh, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ dout,
                         float* __restrict__ partials) {
  extern __shared__ float red[];  // [256/C4][2][C]
  constexpr int CNT = PH * PW;
  const int C4 = g.C >> 2, lanes = 256 / C4;
  const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4;
  const long long npix = (long long)g.N * g.Hp * g.Wp;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  const F4 s = ldg4(sc + c4 * 4), t = ldg4(sh + c4 * 4), mu = ldg4(mean + c4 * 4), is = ldg4(invstd + c4 * 4);
  const size_t rowC = (size_t)g.W * g.C;
  for (long long p = (long long)blockIdx.x * lanes + pl; p < npix; p += (long long)gridDim.x * lanes) {
    long long r = p;
    const int ox = (int)(r % g.Wp);
    r /= g.Wp;
    const int oy = (int)(r % g.Hp), n = (int)(r / g.Hp);
    const F4 gv = ld4(dout + (size_t)n * g.son + (size_t)oy * g.soh + (size_t)ox * g.sow + c4 * 4);
    const float* y0 = y + (((size_t)n * g.H + oy * PH) * g.W + ox * PW) * g.C + c4 * 4;
    F4 q[CNT];
#pragma unroll
    for (int dy = 0; dy < PH; ++dy)
#pragma unroll
      for (int dx = 0; dx < PW; ++dx) q[dy * PW + dx] = ld4(y0 + dy * rowC + (size_t)dx * g.C);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[CNT];
#pragma unroll
      for (int wi = 0; wi < CNT; ++wi) v[wi] = q[wi].v[j];
      int am;
      float gp;
      route_static<CNT, MODE>(v, s.v[j], t.v[j], g.relu, gv.v[j], am, gp);
      if (MODE == 1) {
        float sv = 0.f;
#pragma unroll
        for (int wi = 0; wi < CNT; ++wi) sv += v[wi];
        a[j] += gp * (float)CNT;
        b[j] = fmaf(gp, (sv - (float)CNT * mu.v[j]) * is.v[j], b[j]);
      } else {
        float vm = v[0];
#pragma unroll
        for (int wi = 1; wi < CNT; ++wi) vm = (wi == am) ? v[wi] : vm;
        a[j] += gp;
        b[j] = fmaf(gp, (vm - mu.v[j]) * is.v[j], b[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[(pl * 2) * g.C + c4 * 4 + j] = a[j];
    red[(pl * 2 + 1) * g.C + c4 * 4 + j] = b[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * g.C; i += 256) {
    float sum = 0.f;
    for (int q2 = 0; q2 < lanes; ++q2) sum += red[q2 * 2 * g.C + i];
    partials[(size_t)blockIdx.x * 2 * g.C + i] = sum;
  }
}

// One thread per (sample, pooling cell, 4 channels); cells cover the whole input, the ragged border cells
// (outside every window) get dz = 0.
template <int PH, int PW, int MODE>
__global__ void __launch_bounds__(256)
bn_act_pool_bwd_apply_t(const float* __restrict__ y, PoolGeom g, const float* __restrict__ sc,
                        const float* __restrict__ sh, const float* __restrict__ k1, const float* __restrict__ k2,
                        const float* __restrict__ k3, const float* __restrict__ dout, float* __restrict__ dy) {
  constexpr int CNT = PH * PW;
  const int C4 = g.C >> 2;
  const int Hc = (g.H + PH - 1) / PH, Wc = (g.W + PW - 1) / PW;
  const long long total = (long long)g.N * Hc * Wc * C4;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int c4 = (int)(i % C4);
  long long r = i / C4;
  const int cx = (int)(r % Wc);
  r /= Wc;
  const int cy = (int)(r % Hc), n = (int)(r / Hc);
  const bool full = cy < g.Hp && cx < g.Wp;
  const size_t rowC = (size_t)g.W * g.C;
  const size_t off0 = (((size_t)n * g.H + cy * PH) * g.W + cx * PW) * g.C + c4 * 4;
  F4 q[CNT];
  bool in[CNT];
#pragma unroll
  for (int dy_ = 0; dy_ < PH; ++dy_)
#pragma unroll
    for (int dx = 0; dx < PW; ++dx) {
      const int wi = dy_ * PW + dx;
      in[wi] = full || (cy * PH + dy_ < g.H && cx * PW + dx < g.W);
      q[wi] = in[wi] ? ld4(y + off0 + dy_ * rowC + (size_t)dx * g.C) : F4{{0.f, 0.f, 0.f, 0.f}};
    }
  F4 gv{{0.f, 0.f, 0.f, 0.f}};
  if (full) gv = ld4(dout + (size_t)n * g.son + (size_t)cy * g.soh + (size_t)cx * g.sow + c4 * 4);
  const F4 s = ldg4(sc + c4 * 4), t = ldg4(sh + c4 * 4);
  const F4 a1 = ldg4(k1 + c4 * 4), a2 = ldg4(k2 + c4 * 4), a3 = ldg4(k3 + c4 * 4);
  F4 o[CNT];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v[CNT];
#pragma unroll
    for (int wi = 0; wi < CNT; ++wi) v[wi] = q[wi].v[j];
    int am = -2;
    float gp = 0.f;
    if (full) route_static<CNT, MODE>(v, s.v[j], t.v[j], g.relu, gv.v[j], am, gp);
#pragma unroll
    for (int wi = 0; wi < CNT; ++wi) {
      const float dz = (am == -1 || am == wi) ? gp : 0.f;
      o[wi].v[j] = fmaf(a1.v[j], dz, fmaf(a2.v[j], v[wi], a3.v[j]));
    }
  }
#pragma unroll
  for (int dy_ = 0; dy_ < PH; ++dy_)
#pragma unroll
    for (int dx = 0; dx < PW; ++dx) {
      const int wi = dy_ * PW + dx;
      if (in[wi])
        *reinterpret_cast<float4*>(dy + off0 + dy_ * rowC + (size_t)dx * g.C) =
            make_float4(o[wi].v[0], o[wi].v[1], o[wi].v[2], o[wi].v[3]);
    }
}

// ReLU backward in place on a GEMM output that was stored post-ReLU: d *= (a > 0).
__global__ void relu_bwd_kernel(const float* __restrict__ a, float* __restrict__ d, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 av = reinterpret_cast<const float4*>(a)[i];
  float4 dv = reinterpret_cast<float4*>(d)[i];
  dv.x = av.x > 0.f ? dv.x : 0.f; dv.y = av.y > 0.f ? dv.y : 0.f;
  dv.z = av.z > 0.f ? dv.z : 0.f; dv.w = av.w > 0.f ? dv.w : 0.f;
  reinterpret_cast<float4*>(d)[i] = dv;
}

// ---------------------------------------------------------------------------------------------
// LogSoftmax over the last dim, one warp per row (models.py:250), and its backward.
__global__ void log_softmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int R, int C) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* xr = x + (size_t)row * C;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, xr[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(xr[c] - m);
  s = warp_sum(s);
  const float lse = m + logf(s);
  for (int c = lane; c < C; c += 32) y[(size_t)row * C + c] = xr[c] - lse;
}
__global__ void log_softmax_bwd_kernel(const float* __restrict__ y, const float* __restrict__ g,
                                       float* __restrict__ dx, int R, int C, int ld_dx) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += g[(size_t)row * C + c];
  s = warp_sum(s);
  for (int c = lane; c < ld_dx; c += 32)  // columns C .. ld_dx-1 are zero padding (16-byte row pitch for the TMA GEMMs)
    dx[(size_t)row * ld_dx + c] = c < C ? g[(size_t)row * C + c] - expf(y[(size_t)row * C + c]) * s : 0.f;
}

// dst[c][r] = src[r][c] (small weight transposes)
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int C) {
  __shared__ float tile[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8)
    if (r0 + i < R && c < C) tile[i][threadIdx.x] = src[(size_t)(r0 + i) * C + c];
  __syncthreads();
  const int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8)
    if (c0 + i < C && r < R) dst[(size_t)(c0 + i) * R + r] = tile[threadIdx.x][i];
}

}  // namespace

extern "C" {

// conv.0 + ReLU + MaxPool2d(2) -> NHWC [N][H/2][W/2][32]. code (optional, [N * H/2 * W/2 * 4] int32 words): per pooled pixel and
// 8-channel group the arg-max window positions and ReLU gates, which ocrs_rec_conv0_bwd then uses instead of recomputing.
int ocrs_rec_conv0_fwd(const float* x, int N, int H, int W, const float* w, const float* bias, float* out,
                       int* code, void* stream) {
  OCRS_CHECK_ARG(H >= 2 && W >= 2, "conv0_fwd: input too small");
  const long long total = (long long)N * (H / 2) * (W / 2) * 4;
  conv0_fwd_kernel<<<ocrs_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, w, bias, out, reinterpret_cast<unsigned*>(code));
  OCRS_CHECK_LAUNCH("conv0_fwd_kernel");
  return 0;
}

int ocrs_rec_conv0_bwd_blocks(void) { return 4 * OCRS_NUM_SMS; }
// Weight / bias gradient of conv.0 (its input is the image: no data gradient). partials: [blocks][32][10] (9 taps + bias).
// code: what ocrs_rec_conv0_fwd saved, or null to recompute the pre-pool values here.
int ocrs_rec_conv0_bwd(const float* x, int N, int H, int W, const float* w, const float* bias,
                       const float* dout, const int* code, float* partials, void* stream) {
  conv0_bwd_kernel<<<ocrs_rec_conv0_bwd_blocks(), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, w, bias, dout, reinterpret_cast<const unsigned*>(code), partials);
  OCRS_CHECK_LAUNCH("conv0_bwd_kernel");
  return 0;
}

static int make_geom(PoolGeom& g, int N, int H, int W, int C, int ph, int pw, int mode, int relu,
                     long long son, long long soh, long long sow) {
  OCRS_CHECK_ARG(C % 4 == 0 && C <= 1024 && 256 % (C / 4) == 0, "bn_act_pool: unsupported channel count %d", C);
  OCRS_CHECK_ARG(ph * pw <= MAXWIN && ph >= 1 && pw >= 1, "bn_act_pool: window %dx%d too large", ph, pw);
  g = PoolGeom{N, H, W, C, ph, pw, H / ph, W / pw, mode, relu, son, soh, sow};
  OCRS_CHECK_ARG(g.Hp > 0 && g.Wp > 0, "bn_act_pool: input smaller than the window");
  return 0;
}

int ocrs_rec_bn_act_pool_fwd(const float* y, int N, int H, int W, int C, int ph, int pw, int mode, int relu,
                             const float* sc, const float* sh, float* out, long long son, long long soh,
                             long long sow, void* stream) {
  PoolGeom g;
  if (make_geom(g, N, H, W, C, ph, pw, mode, relu, son, soh, sow)) return -1;
  const long long total = (long long)N * g.Hp * g.Wp * (C / 4);
  bn_act_pool_fwd_kernel<<<ocrs_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(y, g, sc, sh, out);
  OCRS_CHECK_LAUNCH("bn_act_pool_fwd_kernel");
  return 0;
}

int ocrs_rec_pool_bwd_blocks(void) { return 4 * OCRS_NUM_SMS; }
// partials: [blocks][2][C]
int ocrs_rec_bn_act_pool_bwd_reduce(const float* y, int N, int H, int W, int C, int ph, int pw, int mode,
                                    int relu, const float* sc, const float* sh, const float* mean,
                                    const float* invstd, const float* dout, long long son, long long soh,
                                    long long sow, float* partials, void* stream) {
  PoolGeom g;
  if (make_geom(g, N, H, W, C, ph, pw, mode, relu, son, soh, sow)) return -1;
  const int lanes = 256 / (C / 4);
  const size_t smem = (size_t)lanes * 2 * C * sizeof(float);
  OCRS_CHECK_ARG(smem <= 48 * 1024, "bn_act_pool_bwd_reduce: smem");
  const int blocks = ocrs_rec_pool_bwd_blocks();
  cudaStream_t st = (cudaStream_t)stream;
#define OCRS_POOL_REDUCE(PH_, PW_, MODE_) \
  bn_act_pool_bwd_reduce_t<PH_, PW_, MODE_><<<blocks, 256, smem, st>>>(y, g, sc, sh, mean, invstd, dout, partials)
  if (ph == 2 && pw == 2 && mode == 0) OCRS_POOL_REDUCE(2, 2, 0);
  else if (ph == 2 && pw == 1 && mode == 0) OCRS_POOL_REDUCE(2, 1, 0);
  else if (ph == 4 && pw == 1 && mode == 1) OCRS_POOL_REDUCE(4, 1, 1);
  else bn_act_pool_bwd_reduce_kernel<<<blocks, 256, smem, st>>>(y, g, sc, sh, mean, invstd, dout, partials);
#undef OCRS_POOL_REDUCE
  OCRS_CHECK_LAUNCH("bn_act_pool_bwd_reduce_kernel");
  return 0;
}

int ocrs_rec_bn_act_pool_bwd_apply(const float* y, int N, int H, int W, int C, int ph, int pw, int mode,
                                   int relu, const float* sc, const float* sh, const float* k1,
                                   const float* k2, const float* k3, const float* dout, long long son,
                                   long long soh, long long sow, float* dy, void* stream) {
  PoolGeom g;
  if (make_geom(g, N, H, W, C, ph, pw, mode, relu, son, soh, sow)) return -1;
  const long long total = (long long)N * H * W * (C / 4);
  const long long cells = (long long)N * ocrs_cdiv(H, ph) * ocrs_cdiv(W, pw) * (C / 4);
  cudaStream_t st = (cudaStream_t)stream;
#define OCRS_POOL_APPLY(PH_, PW_, MODE_) \
  bn_act_pool_bwd_apply_t<PH_, PW_, MODE_><<<ocrs_cdiv(cells, 256), 256, 0, st>>>(y, g, sc, sh, k1, k2, k3, dout, dy)
  if (ph == 2 && pw == 2 && mode == 0) OCRS_POOL_APPLY(2, 2, 0);
  else if (ph == 2 && pw == 1 && mode == 0) OCRS_POOL_APPLY(2, 1, 0);
  else if (ph == 4 && pw == 1 && mode == 1) OCRS_POOL_APPLY(4, 1, 1);
  else bn_act_pool_bwd_apply_kernel<<<ocrs_cdiv(total, 256), 256, 0, st>>>(y, g, sc, sh, k1, k2, k3, dout, dy);
#undef OCRS_POOL_APPLY
  OCRS_CHECK_LAUNCH("bn_act_pool_bwd_apply_kernel");
  return 0;
}

int ocrs_relu_bwd(const float* act, float* grad, long long n, void* stream) {
  OCRS_CHECK_ARG(n % 4 == 0, "relu_bwd: length must be a multiple of 4");
  relu_bwd_kernel<<<ocrs_cdiv(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(act, grad, n / 4);
  OCRS_CHECK_LAUNCH("relu_bwd_kernel");
  return 0;
}

int ocrs_log_softmax_fwd(const float* x, float* y, int R, int C, void* stream) {
  log_softmax_fwd_kernel<<<ocrs_cdiv(R, 8), 256, 0, (cudaStream_t)stream>>>(x, y, R, C);
  OCRS_CHECK_LAUNCH("log_softmax_fwd_kernel");
  return 0;
}
// dx [R][ld_dx] (ld_dx >= C; the padding columns are written as zeros) from y = log_softmax(x) [R][C] and g = dL/dy [R][C].
int ocrs_log_softmax_bwd(const float* y, const float* g, float* dx, int R, int C, int ld_dx, void* stream) {
  OCRS_CHECK_ARG(ld_dx >= C, "log_softmax_bwd: row pitch %d < %d columns", ld_dx, C);
  log_softmax_bwd_kernel<<<ocrs_cdiv(R, 8), 256, 0, (cudaStream_t)stream>>>(y, g, dx, R, C, ld_dx);
  OCRS_CHECK_LAUNCH("log_softmax_bwd_kernel");
  return 0;
}

int ocrs_transpose(const float* src, float* dst, int R, int C, void* stream) {
  dim3 grid(ocrs_cdiv(C, 32), ocrs_cdiv(R, 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, dst, R, C);
  OCRS_CHECK_LAUNCH("transpose_kernel");
  return 0;
}

}  // extern "C"

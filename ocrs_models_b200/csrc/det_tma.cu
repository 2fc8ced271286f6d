// TMA-pipelined kernels of the detection U-Net's DepthwiseConv blocks (reference ocrs_models/models.py:7-28 and
// its autograd). Same math and HBM layout as det_fwd.cu / det_bwd.cu (planar NCHW fp32 views, BatchNorm+ReLU
// folded into the consumer's load), different data movement:
//
//   * persistent CTAs (<= 2 per SM) walk 32x32-pixel tiles; the haloed input tile of a channel chunk is ONE
//     cp.async.bulk.tensor (TMA) box {40, 34, chunk} over a 4-D (W, H, C, N) tensor map - the hardware zero-fills
//     the halo outside the image - landing in a 4-stage shared-memory ring guarded by mbarriers, so the loads of
//     the next tiles are in flight while the current one is computed (the old kernels did load -> sync -> compute).
//     The box starts at x0-4, not x0-1: the innermost TMA coordinate must be a multiple of 16 bytes (measured:
//     UTMALDG raises "illegal instruction" otherwise, scripts/probe/tma_probe.cu);
//   * a warp owns 4 rows x 32 consecutive pixels, a thread one column of 4 pixels: conflict-free 4-byte
//     shared-memory reads, 128-byte coalesced global stores per warp instruction;
//   * per-channel statistics (BatchNorm partial sums, weight-gradient partial sums) stay in registers across all
//     the tiles of a CTA and are written once per CTA: partial rows = CTAs, reduced in double by finalize_partials
//     (deterministic, no atomics);
//   * the BatchNorm-backward reduction of the UPSTREAM block (sum dz, sum dz*yhat) is produced by the kernel that
//     writes that block's d_a (the depthwise backward kernel here), instead of a separate pass over d_a and y.
//
// TMA needs 16-byte aligned bases and row strides: W % 4 == 0 (ocrs_det_tma_supported); other shapes (e.g. the
// 75x37 level of the 800x600 training size) keep using det_fwd.cu / det_bwd.cu.
#include "common.cuh"
#include "tma_util.cuh"
#include <math.h>
#include <type_traits>

namespace {

constexpr int TW = 32, TH = 32;       // interior tile
constexpr int BW = 40, BH = 34;       // haloed box: x0-4 .. x0+35, y0-1 .. y0+32 (tile column j = image column x0-4+j)
constexpr int PLANE = BW * BH;        // 1360 floats per channel plane of a box
constexpr int PPT = 4;                // rows per thread
constexpr int NSTAGE = 4;
constexpr int NTHREADS = 256;
// shared-memory floats of a box of `ch` planes, padded so that the next TMA destination stays 128-byte aligned
constexpr int box_floats(int ch) { return (ch * PLANE + 31) / 32 * 32; }
constexpr int box_floats2(int floats) { return (floats + 31) / 32 * 32; }

// Sum over the 32 lanes of v[j] for all j at once; lane l returns the total of column (l % NV).
template <int NV>
__device__ __forceinline__ float warp_colsum(float (&v)[NV], int lane) {
#pragma unroll
  for (int o = NV / 2; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < o; ++j) {
      const float send = up ? v[j] : v[j + o];
      const float keep = up ? v[j + o] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  float r = v[0];
#pragma unroll
  for (int o = NV; o < 32; o <<= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

struct TileWalk {  // spatial tiles (n, ty, tx) assigned round-robin to the CTAs of one channel group
  int tiles_x, tiles_y, sp_total, sp0, sp_stride;
  __device__ __forceinline__ int count() const { return sp0 < sp_total ? (sp_total - sp0 + sp_stride - 1) / sp_stride : 0; }
  __device__ __forceinline__ void decode(int k, int& n, int& x0, int& y0) const {
    const unsigned sp = (unsigned)(sp0 + k * sp_stride), per = (unsigned)(tiles_x * tiles_y);
    const unsigned nn = sp / per, r = sp - nn * per, ry = r / (unsigned)tiles_x;  // unsigned: no sign fix-up code
    n = (int)nn;
    y0 = (int)ry * TH;
    x0 = (int)(r - ry * (unsigned)tiles_x) * TW;
  }
};

// Load the 6x3 window (tile rows 4w..4w+5, tile columns lane+3..lane+5) of one channel plane, apply the producer's
// BatchNorm+ReLU transform and zero what lies outside the image (padding applies AFTER activation).
template <bool XF, bool BORDER>
__device__ __forceinline__ void load_window(const float* t, float sc, float sh, float lo, const bool (&rowok)[PPT + 2],
                                            const bool (&colok)[3], float (&v)[PPT + 2][3]) {
#pragma unroll
  for (int rr = 0; rr < PPT + 2; ++rr)
#pragma unroll
    for (int k = 0; k < 3; ++k) v[rr][k] = t[rr * BW + k];
  if (XF) {
#pragma unroll
    for (int rr = 0; rr < PPT + 2; ++rr)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        v[rr][k] = xform_apply(v[rr][k], sc, sh, lo);
        if (BORDER) v[rr][k] = (rowok[rr] && colok[k]) ? v[rr][k] : 0.f;
      }
  }
}

// Per-thread validity of its window rows / columns for a tile at (x0, y0); returns whether the tile touches the border.
__device__ __forceinline__ bool tile_masks(int x0, int y0, int H, int W, int lane, int warp, bool (&rowok)[PPT + 2],
                                           bool (&colok)[3]) {
  const bool border = x0 == 0 || y0 == 0 || x0 + TW + 1 > W || y0 + TH + 1 > H;  // uniform per CTA
#pragma unroll
  for (int rr = 0; rr < PPT + 2; ++rr) { const int gy = y0 - 1 + PPT * warp + rr; rowok[rr] = gy >= 0 && gy < H; }
#pragma unroll
  for (int q = 0; q < 3; ++q) { const int gx = x0 - 1 + lane + q; colok[q] = gx >= 0 && gx < W; }
  return border;
}

// Packed fp32 FMA (Blackwell FFMA2): two independent multiply-adds per issue slot.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
  return v;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// One stage (FCH channels) of the forward block: depthwise 3x3 on the activated window, then the 1x1 contraction.
template <bool XF, bool BORDER, int CO_T, int FCH>
__device__ __forceinline__ void fwd_chunk(const float* st, const float* sdw, const float* spw, const float* sxf, int Cin,
                                          int c0, const bool (&rowok)[PPT + 2], const bool (&colok)[3],
                                          u64 (&acc)[PPT][CO_T / 2], float* dwo, size_t plane, int W, unsigned okmask) {
#pragma unroll
  for (int c = 0; c < FCH; ++c) {
    float v[PPT + 2][3];
    const int ci = c0 + c;
    float sc = 1.f, sh = 0.f, lo = 0.f;
    if (XF) { sc = sxf[ci]; sh = sxf[Cin + ci]; lo = sxf[2 * Cin + ci]; }
    load_window<XF, BORDER>(st + c * PLANE, sc, sh, lo, rowok, colok, v);
    const float4 w0 = *reinterpret_cast<const float4*>(sdw + ci * 12);
    const float4 w1 = *reinterpret_cast<const float4*>(sdw + ci * 12 + 4);
    const float w8 = sdw[ci * 12 + 8];
    float d[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      float t = v[i][0] * w0.x;
      t = fmaf(v[i][1], w0.y, t); t = fmaf(v[i][2], w0.z, t);
      t = fmaf(v[i + 1][0], w0.w, t); t = fmaf(v[i + 1][1], w1.x, t); t = fmaf(v[i + 1][2], w1.y, t);
      t = fmaf(v[i + 2][0], w1.z, t); t = fmaf(v[i + 2][1], w1.w, t); t = fmaf(v[i + 2][2], w8, t);
      d[i] = t;
    }
    if (dwo) {  // uniform: training-mode forward of co-tile 0 keeps the depthwise output for the weight gradient
#pragma unroll
      for (int i = 0; i < PPT; ++i)
        if (okmask & (1u << i)) dwo[(size_t)ci * plane + (size_t)i * W] = d[i];
    }
    // 1x1 contraction on packed FFMA2: (acc[o], acc[o+1]) += (d, d) * (w[o], w[o+1]) - half the issue slots of 64 FFMA
    u64 dd[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) dd[i] = pack2(d[i], d[i]);
    const float4* w4 = reinterpret_cast<const float4*>(spw + ci * CO_T);
#pragma unroll
    for (int o4 = 0; o4 < CO_T / 4; ++o4) {
      const float4 wv = w4[o4];
      const u64 w01 = pack2(wv.x, wv.y), w23 = pack2(wv.z, wv.w);
#pragma unroll
      for (int i = 0; i < PPT; ++i) {
        acc[i][o4 * 2 + 0] = fma2(dd[i], w01, acc[i][o4 * 2 + 0]);
        acc[i][o4 * 2 + 1] = fma2(dd[i], w23, acc[i][o4 * 2 + 1]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Forward: y = pw1x1(dw3x3(xform(x))) + per-CTA BatchNorm partial sums (models.py:11-22).
struct FwdArgs {
  int Cin, Cout, H, W, N, n_cot;
  const float *in_scale, *in_shift, *in_lo, *wdw, *wpw;
  float* y; long long y_ss;
  float* dwo;       // [N][Cin][H][W] depthwise output saved for the 1x1 weight gradient, or null
  float* partials;  // [gridDim.x / n_cot][2][Cout] or null
  TileWalk walk;    // sp0 / sp_stride filled per CTA
};

template <int CO_T, int FCH>
__global__ void __launch_bounds__(NTHREADS, 2)
sep_fwd_tma_kernel(const __grid_constant__ CUtensorMap xmap, FwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int STAGE_FLOATS = box_floats(FCH);
  float* stages = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + NSTAGE * STAGE_FLOATS);  // 64 bytes reserved
  float* sdw = stages + NSTAGE * STAGE_FLOATS + 16;    // [Cin][12] (9 taps padded to 12 for 16-byte reads)
  float* spw = sdw + a.Cin * 12;                       // [Cin][CO_T]
  float* sred = spw + a.Cin * CO_T;                    // [8][2*CO_T]
  float* sxf = sred + 8 * 2 * CO_T;                    // [3][Cin]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cot = blockIdx.x % a.n_cot, co0 = cot * CO_T;
  TileWalk walk = a.walk;
  walk.sp0 = blockIdx.x / a.n_cot;
  walk.sp_stride = gridDim.x / a.n_cot;
  const int Cin = a.Cin;
  const bool has_xf = a.in_scale != nullptr;

  for (int i = tid; i < Cin * 12; i += NTHREADS) {
    const int c = i / 12, k = i - c * 12;
    sdw[i] = k < 9 ? a.wdw[(size_t)c * 9 + k] : 0.f;
  }
  for (int i = tid; i < Cin * CO_T; i += NTHREADS) {
    const int c = i / CO_T, o = i - c * CO_T;
    spw[i] = (co0 + o < a.Cout) ? a.wpw[(size_t)(co0 + o) * Cin + c] : 0.f;
  }
  if (has_xf)
    for (int i = tid; i < Cin; i += NTHREADS) { sxf[i] = a.in_scale[i]; sxf[Cin + i] = a.in_shift[i]; sxf[2 * Cin + i] = a.in_lo[i]; }
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) tma::mbar_init(tma::smem_u32(&bars[s]), 1);
    tma::fence_barrier_init();
    tma::prefetch_map(&xmap);
  }
  __syncthreads();

  const int nchunks = (Cin + FCH - 1) / FCH;
  const int total = walk.count() * nchunks;
  auto issue = [&](int j) {
    const int k = j / nchunks, ch = j - k * nchunks;
    int n, x0, y0;
    walk.decode(k, n, x0, y0);
    const uint32_t bar = tma::smem_u32(&bars[j % NSTAGE]);
    tma::mbar_expect_tx(bar, FCH * PLANE * 4);  // the box's bytes (zero-filled halo included), not the padded stage
    tma::load_4d(tma::smem_u32(stages + (j % NSTAGE) * STAGE_FLOATS), &xmap, x0 - 4, y0 - 1, ch * FCH, n, bar);
  };
  if (tid == 0)
    for (int j = 0; j < NSTAGE && j < total; ++j) issue(j);

  u64 acc2[PPT][CO_T / 2];
  float stat = 0.f;  // lane l accumulates column l of (sum[0..CO_T), sumsq[0..CO_T)) over all tiles of this CTA
  bool rowok[PPT + 2], colok[3];
  bool border = false;
  int n = 0, x0 = 0, y0 = 0;
  int k = 0, ch = 0;
  const size_t plane = (size_t)a.H * a.W;
  float* dwo_t = nullptr;   // this thread's first pixel of the saved depthwise output (channel 0), current tile
  unsigned okmask = 0;      // which of its PPT pixels are inside the image
  for (int j = 0; j < total; ++j) {
    if (ch == 0) {
      walk.decode(k, n, x0, y0);
#pragma unroll
      for (int i = 0; i < PPT; ++i)
#pragma unroll
        for (int o = 0; o < CO_T / 2; ++o) acc2[i][o] = 0ull;
      border = tile_masks(x0, y0, a.H, a.W, lane, warp, rowok, colok);
      if (a.dwo != nullptr && cot == 0) {
        const int gx = x0 + lane, gy0 = y0 + PPT * warp;
        dwo_t = a.dwo + ((size_t)n * Cin * a.H + gy0) * a.W + gx;
        okmask = 0;
#pragma unroll
        for (int i = 0; i < PPT; ++i) okmask |= (gx < a.W && gy0 + i < a.H) ? (1u << i) : 0u;
      }
    }
    const int s = j % NSTAGE;
    tma::mbar_wait(tma::smem_u32(&bars[s]), (j / NSTAGE) & 1);
    const float* st = stages + s * STAGE_FLOATS + (PPT * warp) * BW + lane + 3;
    const int c0 = ch * FCH;  // Cin is 1 (FCH = 1) or a multiple of FCH (checked on the host)
    if (!has_xf) fwd_chunk<false, false, CO_T, FCH>(st, sdw, spw, sxf, Cin, c0, rowok, colok, acc2, dwo_t, plane, a.W, okmask);
    else if (border) fwd_chunk<true, true, CO_T, FCH>(st, sdw, spw, sxf, Cin, c0, rowok, colok, acc2, dwo_t, plane, a.W, okmask);
    else fwd_chunk<true, false, CO_T, FCH>(st, sdw, spw, sxf, Cin, c0, rowok, colok, acc2, dwo_t, plane, a.W, okmask);
    __syncthreads();  // every thread is done with stage s
    if (tid == 0 && j + NSTAGE < total) issue(j + NSTAGE);
    if (++ch == nchunks) {
      ch = 0;
      ++k;
      const int gx = x0 + lane, gy0 = y0 + PPT * warp;
      float acc[PPT][CO_T];
#pragma unroll
      for (int i = 0; i < PPT; ++i)
#pragma unroll
        for (int o = 0; o < CO_T / 2; ++o) unpack2(acc2[i][o], acc[i][2 * o], acc[i][2 * o + 1]);
      float sv[2 * CO_T];
      float* yp = a.y + (size_t)n * a.y_ss + ((size_t)co0 * a.H + gy0) * a.W + gx;  // Cout % CO_T == 0 (host check)
      const size_t HW = (size_t)a.H * a.W;
      if (x0 + TW <= a.W && y0 + TH <= a.H) {  // full tile (uniform per CTA): no bounds checks
#pragma unroll
        for (int o = 0; o < CO_T; ++o) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < PPT; ++i) {
            s1 += acc[i][o];
            s2 = fmaf(acc[i][o], acc[i][o], s2);
            yp[o * HW + (size_t)i * a.W] = acc[i][o];
          }
          sv[o] = s1;
          sv[CO_T + o] = s2;
        }
      } else {
#pragma unroll
        for (int o = 0; o < CO_T; ++o) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < PPT; ++i) {
            if (gx < a.W && gy0 + i < a.H) {
              s1 += acc[i][o];
              s2 = fmaf(acc[i][o], acc[i][o], s2);
              yp[o * HW + (size_t)i * a.W] = acc[i][o];
            }
          }
          sv[o] = s1;
          sv[CO_T + o] = s2;
        }
      }
      if (a.partials) stat += warp_colsum<2 * CO_T>(sv, lane);
    }
  }
  if (a.partials) {
    if (lane < 2 * CO_T) sred[warp * 2 * CO_T + lane] = stat;
    __syncthreads();
    if (tid < 2 * CO_T) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += sred[w * 2 * CO_T + tid];
      const int which = tid / CO_T, o = tid - which * CO_T;
      if (co0 + o < a.Cout)
        a.partials[((size_t)(blockIdx.x / a.n_cot) * 2 + which) * a.Cout + co0 + o] = s;
    }
  }
}

constexpr int FWD_MAX_CIN = 256;  // the attribute is set once per device for the largest block of the network
template <int CO_T, int FCH>
size_t fwd_smem(int Cin) {
  return (size_t)NSTAGE * box_floats(FCH) * 4 + 64 + (size_t)Cin * (12 + CO_T + 3) * 4 + 8 * 2 * CO_T * 4 + 128;
}

int fwd_ctas(int N, int H, int W, int n_cot) {
  const long long sp = (long long)N * ocrs_cdiv(W, TW) * ocrs_cdiv(H, TH);
  long long per = (2 * OCRS_NUM_SMS) / n_cot;
  if (per < 1) per = 1;
  return (int)(sp < per ? sp : per);
}

// ---------------------------------------------------------------------------------------------------------------
// Depthwise 3x3 backward (per channel): dx = corr(g, flip(w)) [+= old dx], dWdw[k] = sum g * xact(shifted),
// and - for the block that PRODUCED x - the BatchNorm-backward sums over this kernel's output d_a = dx:
//   sum dz, sum dz * (x - mean) * invstd   with   dz = dx * [xform(x) > lo].
// CTA = (channel pair, subset of the spatial tiles): all per-channel sums live in registers for the whole kernel.
struct DwBwdArgs {
  int C, H, W, N, accumulate, ctas_per_chunk;
  const float *isc, *ish, *ilo, *wdw;
  const float *up_mean, *up_invstd;  // upstream BatchNorm statistics (null: no fused reduction)
  float* dx; long long dx_ss;
  float* wpart;    // [ctas_per_chunk][C][9]
  float* bnpart;   // [ctas_per_chunk][2][C] or null
  TileWalk walk;
};
constexpr int DCH = 2;  // channels per stage (g and x boxes of both)

__global__ void __launch_bounds__(NTHREADS, 2)
sep_dw_bwd_tma_kernel(const __grid_constant__ CUtensorMap gmap, const __grid_constant__ CUtensorMap xmap, DwBwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int XOFF = box_floats(DCH);
  constexpr int STAGE_FLOATS = 2 * XOFF;  // [g: DCH planes][x: DCH planes]
  float* stages = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + NSTAGE * STAGE_FLOATS);  // 64 bytes reserved
  float* sred = stages + NSTAGE * STAGE_FLOATS + 16;  // [8][DCH * 11]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = blockIdx.x / a.ctas_per_chunk, c0 = chunk * DCH;
  TileWalk walk = a.walk;
  walk.sp0 = blockIdx.x % a.ctas_per_chunk;
  walk.sp_stride = a.ctas_per_chunk;
  const bool has_xf = a.isc != nullptr;
  const bool need_bn = a.bnpart != nullptr;

  float w[DCH][9], sc[DCH], sh[DCH], lo[DCH], mu[DCH], is[DCH];
#pragma unroll
  for (int c = 0; c < DCH; ++c) {
    const bool v = c0 + c < a.C;
#pragma unroll
    for (int q = 0; q < 9; ++q) w[c][q] = v ? a.wdw[(size_t)(c0 + c) * 9 + q] : 0.f;
    sc[c] = (v && has_xf) ? a.isc[c0 + c] : 1.f;
    sh[c] = (v && has_xf) ? a.ish[c0 + c] : 0.f;
    lo[c] = (v && has_xf) ? a.ilo[c0 + c] : -INFINITY;
    mu[c] = (v && need_bn) ? a.up_mean[c0 + c] : 0.f;
    is[c] = (v && need_bn) ? a.up_invstd[c0 + c] : 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) tma::mbar_init(tma::smem_u32(&bars[s]), 1);
    tma::fence_barrier_init();
    tma::prefetch_map(&gmap);
    tma::prefetch_map(&xmap);
  }
  __syncthreads();
  const int total = walk.count();
  auto issue = [&](int j) {
    int n, x0, y0;
    walk.decode(j, n, x0, y0);
    const uint32_t bar = tma::smem_u32(&bars[j % NSTAGE]);
    float* dst = stages + (j % NSTAGE) * STAGE_FLOATS;
    tma::mbar_expect_tx(bar, 2 * DCH * PLANE * 4);
    tma::load_4d(tma::smem_u32(dst), &gmap, x0 - 4, y0 - 1, c0, n, bar);
    tma::load_4d(tma::smem_u32(dst + XOFF), &xmap, x0 - 4, y0 - 1, c0, n, bar);
  };
  if (tid == 0)
    for (int j = 0; j < NSTAGE && j < total; ++j) issue(j);

  float dwacc[DCH][9], bn1[DCH], bn2[DCH];
#pragma unroll
  for (int c = 0; c < DCH; ++c) {
    bn1[c] = 0.f; bn2[c] = 0.f;
#pragma unroll
    for (int q = 0; q < 9; ++q) dwacc[c][q] = 0.f;
  }
  // One tile. FULL: the tile and its halo lie inside the image (88 % of the tiles at 1024^2): no masks, no bounds checks.
  auto tile_body = [&](auto full_tag, int j, int n, int x0, int y0) {
    constexpr bool FULL = decltype(full_tag)::value;
    const int gx = x0 + lane, gy0 = y0 + PPT * warp;
    const size_t plane = (size_t)a.H * a.W;
    float* dx0 = a.dx + (size_t)n * a.dx_ss + ((size_t)c0 * a.H + gy0) * a.W + gx;
    unsigned ok = (1u << PPT) - 1;  // which of this thread's PPT pixels are inside the image
    if (!FULL) {
      ok = 0;
#pragma unroll
      for (int i = 0; i < PPT; ++i) ok |= (gx < a.W && gy0 + i < a.H) ? (1u << i) : 0u;
    }
    // previous contents of dx (skip connections: two consumers accumulate): issue the loads before waiting
    float old[DCH][PPT];
#pragma unroll
    for (int c = 0; c < DCH; ++c)
#pragma unroll
      for (int i = 0; i < PPT; ++i) {
        old[c][i] = 0.f;
        if (a.accumulate && (FULL || ((ok >> i) & 1u)) && c0 + c < a.C) old[c][i] = dx0[c * plane + (size_t)i * a.W];
      }
    bool rowok[PPT + 2], colok[3];
    if (!FULL) tile_masks(x0, y0, a.H, a.W, lane, warp, rowok, colok);
    const int s = j % NSTAGE;
    tma::mbar_wait(tma::smem_u32(&bars[s]), (j / NSTAGE) & 1);
    const float* st = stages + s * STAGE_FLOATS + (PPT * warp) * BW + lane + 3;
#pragma unroll
    for (int c = 0; c < DCH; ++c) {
      if (c0 + c >= a.C) continue;
      float gv[PPT + 2][3], xv[PPT + 2][3];
      load_window<false, false>(st + c * PLANE, 1.f, 0.f, 0.f, rowok, colok, gv);  // g is zero outside the image (TMA fill)
      load_window<true, !FULL>(st + XOFF + c * PLANE, sc[c], sh[c], lo[c], rowok, colok, xv);
#pragma unroll
      for (int i = 0; i < PPT; ++i) {
        // dx[p] = sum_k w[k] * g[p - (k - 1)]: window element (i + 2 - ky, 2 - kx)
        float t = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) t = fmaf(w[c][ky * 3 + kx], gv[i + 2 - ky][2 - kx], t);
        if (FULL || ((ok >> i) & 1u)) {
          const float gc = gv[i + 1][1];
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) dwacc[c][ky * 3 + kx] = fmaf(gc, xv[i + ky][kx], dwacc[c][ky * 3 + kx]);
          const float o = t + old[c][i];
          dx0[c * plane + (size_t)i * a.W] = o;
          if (need_bn) {
            const float raw = st[XOFF + c * PLANE + (i + 1) * BW + 1];  // xv holds the activated value
            const float dz = (fmaf(raw, sc[c], sh[c]) > lo[c]) ? o : 0.f;
            bn1[c] += dz;
            bn2[c] = fmaf(dz, (raw - mu[c]) * is[c], bn2[c]);
          }
        }
      }
    }
  };
  for (int j = 0; j < total; ++j) {
    int n, x0, y0;
    walk.decode(j, n, x0, y0);
    const bool full = x0 > 0 && y0 > 0 && x0 + TW + 1 <= a.W && y0 + TH + 1 <= a.H;  // uniform per CTA
    if (full) tile_body(std::true_type{}, j, n, x0, y0);
    else tile_body(std::false_type{}, j, n, x0, y0);
    __syncthreads();
    if (tid == 0 && j + NSTAGE < total) issue(j + NSTAGE);
  }
  // CTA reduction of the per-thread sums: DCH * (9 + 2) values
#pragma unroll
  for (int c = 0; c < DCH; ++c) {
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      const float v = warp_sum(dwacc[c][q]);
      if (lane == 0) sred[warp * DCH * 11 + c * 11 + q] = v;
    }
    const float v1 = warp_sum(bn1[c]), v2 = warp_sum(bn2[c]);
    if (lane == 0) { sred[warp * DCH * 11 + c * 11 + 9] = v1; sred[warp * DCH * 11 + c * 11 + 10] = v2; }
  }
  __syncthreads();
  if (tid < DCH * 11) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += sred[wv * DCH * 11 + tid];
    const int c = tid / 11, q = tid - c * 11;
    const int row = blockIdx.x % a.ctas_per_chunk;
    if (c0 + c < a.C) {
      if (q < 9) a.wpart[((size_t)row * a.C + c0 + c) * 9 + q] = s;
      else if (need_bn) a.bnpart[((size_t)row * 2 + (q - 9)) * a.C + c0 + c] = s;
    }
  }
}

int dw_ctas_per_chunk(int N, int H, int W, int C) {
  const long long sp = (long long)N * ocrs_cdiv(W, TW) * ocrs_cdiv(H, TH);
  const int chunks = ocrs_cdiv(C, DCH);
  long long per = (2 * OCRS_NUM_SMS + chunks - 1) / chunks;
  if (per < 1) per = 1;
  return (int)(sp < per ? sp : per);
}


// ---------------------------------------------------------------------------------------------------------------
// Pointwise (1x1) weight gradient dWpw[co][ci] = sum_p dy[co][p] * dwout[ci][p] (dy = BatchNorm/ReLU backward of d_a
// rebuilt on the fly, dwout = dw3x3(xform(x)) recomputed). Same warp-level mma.sync m16n8k8 3xTF32 contraction
// as pw_wgrad_mma_kernel (det_bwd.cu): M = 16 output channels, N = 16 input channels, K = pixels, operands built in
// fragment layout in registers. What changed is everything around it:
//   * the haloed x tile (16 channels x 10 rows x 40 columns) arrives by TMA into a 2-stage ring while the previous
//     tile is computed, and is activated ONCE per element by a vectorised pass into a work buffer whose plane
//     stride (404 floats) makes the fragment reads bank-conflict free (the old kernel activated every element up
//     to nine times while gathering the stencil);
//   * the d_a / y operands of a tile are loaded into registers before waiting for the tile (independent loads);
//   * persistent CTAs, two per SM, no spills.
constexpr int WTW = 32, WTH = 8;                  // tile: one row of 32 pixels per warp (4 k-steps of 8 pixels)
constexpr int WBH = WTH + 2;                      // box rows
constexpr int WPLANE = BW * WBH;                  // 400 floats per channel plane as landed by TMA
constexpr int XS_PLANE = WPLANE + 4;              // activated copy: 404 = 20 (mod 32) -> fragment reads conflict free
constexpr int WCH = 16;

struct PwWgArgs {
  const float *d_a, *y;
  long long da_ss, y_ss;
  const float *sc, *sh, *lo, *k1, *k2, *k3;       // dy = k1 * dz + k2 * y + k3, dz = d_a * [y * sc + sh > lo]
  const float *isc, *ish, *ilo, *wdw;
  float* partials;                                // [gridDim.x][Cout][Cin]
  int Cout, Cin, H, W, N, tiles_x, tiles_y;
};

__device__ __forceinline__ void tf32_split2(float v, uint32_t& hi, uint32_t& lo) {
  // integer add + mask (2 ALU instructions): measured faster than cvt.rna.tf32.f32, which issues on the quarter-rate conversion pipe
  hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32_16n8k8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(NTHREADS, 2)
sep_pw_wgrad_tma_kernel(const __grid_constant__ CUtensorMap xmap, PwWgArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int STAGE_FLOATS = box_floats2(WCH * WPLANE);
  float* stages = reinterpret_cast<float*>(smem_raw);                   // [2][WCH * WPLANE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + 2 * STAGE_FLOATS);
  float* xs = stages + 2 * STAGE_FLOATS + 16;                           // [WCH][XS_PLANE]
  float* sred = xs + WCH * XS_PLANE;                                    // [8][256]
  float* sxf = sred + 8 * 256;                                          // [3][16]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int cit = (a.Cin + 15) / 16;
  const int co0 = (blockIdx.y / cit) * 16, ci0 = (blockIdx.y % cit) * 16;
  const int nco = min(16, a.Cout - co0), nci = min(16, a.Cin - ci0);
  const size_t HW = (size_t)a.H * a.W;
  if (tid < 16) {
    const bool v = tid < nci && a.isc != nullptr;
    sxf[tid] = v ? a.isc[ci0 + tid] : 1.f;
    sxf[16 + tid] = v ? a.ish[ci0 + tid] : 0.f;
    sxf[32 + tid] = v ? a.ilo[ci0 + tid] : -INFINITY;
  }
  // per-lane constants: its two output channels (g, g+8) and its two input channels (g, g+8)
  float ksc[2], ksh[2], klo[2], kk1[2], kk2[2], kk3[2], wd[2][9];
  bool cov[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int o = g + 8 * h;
    cov[h] = o < nco;
    const bool civ = o < nci;
    ksc[h] = cov[h] ? a.sc[co0 + o] : 0.f; ksh[h] = cov[h] ? a.sh[co0 + o] : 0.f; klo[h] = cov[h] ? a.lo[co0 + o] : 0.f;
    kk1[h] = cov[h] ? a.k1[co0 + o] : 0.f; kk2[h] = cov[h] ? a.k2[co0 + o] : 0.f; kk3[h] = cov[h] ? a.k3[co0 + o] : 0.f;
#pragma unroll
    for (int q = 0; q < 9; ++q) wd[h][q] = civ ? a.wdw[(size_t)(ci0 + o) * 9 + q] : 0.f;
  }
  if (tid == 0) {
    tma::mbar_init(tma::smem_u32(&bars[0]), 1);
    tma::mbar_init(tma::smem_u32(&bars[1]), 1);
    tma::fence_barrier_init();
    tma::prefetch_map(&xmap);
  }
  __syncthreads();
  const int tiles = a.tiles_x * a.tiles_y;
  const int total_tiles = a.N * tiles;
  const int mine = blockIdx.x < total_tiles ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto decode = [&](int j, int& n, int& x0, int& y0) {
    const int work = blockIdx.x + j * gridDim.x;
    n = work / tiles;
    const int r = work - n * tiles;
    y0 = (r / a.tiles_x) * WTH;
    x0 = (r % a.tiles_x) * WTW;
  };
  auto issue = [&](int j) {
    int n, x0, y0;
    decode(j, n, x0, y0);
    const uint32_t bar = tma::smem_u32(&bars[j & 1]);
    tma::mbar_expect_tx(bar, WCH * WPLANE * 4);
    tma::load_4d(tma::smem_u32(stages + (j & 1) * STAGE_FLOATS), &xmap, x0 - 4, y0 - 1, ci0, n, bar);
  };
  if (tid == 0) {
    if (mine > 0) issue(0);
    if (mine > 1) issue(1);
  }
  float ctot[2][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) ctot[j][q] = 0.f;

  for (int j = 0; j < mine; ++j) {
    int n, x0, y0;
    decode(j, n, x0, y0);
    // A operands of the whole tile row (4 k-steps x {t, t+4} x {g, g+8}): independent loads, issued before the wait
    const int gy = y0 + warp;
    float dav[4][4], yvv[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int h = q & 1, px = x0 + ks * 8 + t + 4 * (q >> 1);
        dav[ks][q] = 0.f;
        yvv[ks][q] = 0.f;
        if (cov[h] && gy < a.H && px < a.W) {
          const size_t off = (size_t)(co0 + g + 8 * h) * HW + (size_t)gy * a.W + px;
          dav[ks][q] = a.d_a[(size_t)n * a.da_ss + off];
          yvv[ks][q] = a.y[(size_t)n * a.y_ss + off];
        }
      }
    tma::mbar_wait(tma::smem_u32(&bars[j & 1]), (j >> 1) & 1);
    // activate the landed tile once, 4 elements at a time, into the conflict-free work buffer
    {
      const float* st = stages + (j & 1) * STAGE_FLOATS;
      constexpr int V4_PER_PLANE = WPLANE / 4, V4_PER_ROW = BW / 4;
      const int nplanes = (nci + 7) & ~7;  // planes the fragment reads touch; those past nci are zeroed
      for (int e = tid; e < nplanes * V4_PER_PLANE; e += NTHREADS) {
        const int c = e / V4_PER_PLANE, p4 = e - c * V4_PER_PLANE;
        const int r = p4 / V4_PER_ROW, col = (p4 - r * V4_PER_ROW) * 4;
        const int yy = y0 - 1 + r, xx = x0 - 4 + col;
        float4 v = *reinterpret_cast<const float4*>(st + c * WPLANE + p4 * 4);
        const float s = sxf[c], sh_ = sxf[16 + c], l = sxf[32 + c];
        const bool rok = yy >= 0 && yy < a.H && c < nci;
        v.x = (rok && xx >= 0 && xx < a.W) ? xform_apply(v.x, s, sh_, l) : 0.f;
        v.y = (rok && xx + 1 >= 0 && xx + 1 < a.W) ? xform_apply(v.y, s, sh_, l) : 0.f;
        v.z = (rok && xx + 2 >= 0 && xx + 2 < a.W) ? xform_apply(v.z, s, sh_, l) : 0.f;
        v.w = (rok && xx + 3 >= 0 && xx + 3 < a.W) ? xform_apply(v.w, s, sh_, l) : 0.f;
        *reinterpret_cast<float4*>(xs + c * XS_PLANE + p4 * 4) = v;
      }
    }
    __syncthreads();  // xs complete, stage (j & 1) free
    if (tid == 0 && j + 2 < mine) issue(j + 2);
    float c[2][4];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
      for (int q = 0; q < 4; ++q) c[jj][q] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[4], al[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int h = q & 1;
        const float yv = yvv[ks][q];
        const float dz = (fmaf(yv, ksc[h], ksh[h]) > klo[h]) ? dav[ks][q] : 0.f;
        const float dy = fmaf(kk1[h], dz, fmaf(kk2[h], yv, kk3[h]));
        tf32_split2(cov[h] && gy < a.H && x0 + ks * 8 + t + 4 * (q >> 1) < a.W ? dy : 0.f, ah[q], al[q]);
      }
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        if (8 * jj >= nci) continue;  // uniform: no input channels in this n-tile
        // dwout[ci = 8jj + g][pixel t / t+4]: depthwise stencil on the activated tile
        float b[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float* tp = xs + (8 * jj + g) * XS_PLANE + warp * BW + ks * 8 + t + 4 * e + 3;
          float sacc = 0.f;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) sacc = fmaf(tp[ky * BW + kx], wd[jj][ky * 3 + kx], sacc);
          b[e] = sacc;
        }
        uint32_t bh0, bl0, bh1, bl1;
        tf32_split2(b[0], bh0, bl0);
        tf32_split2(b[1], bh1, bl1);
        mma_tf32_16n8k8(c[jj], al, bh0, bh1);
        mma_tf32_16n8k8(c[jj], ah, bl0, bl1);
        mma_tf32_16n8k8(c[jj], ah, bh0, bh1);
      }
    }
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
      for (int q = 0; q < 4; ++q) ctot[jj][q] += c[jj][q];  // flush: the tensor core's own accumulation stays short
    __syncthreads();  // everyone is done with xs before the next tile overwrites it
  }
  // C fragment -> (co, ci): c0 (g, 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1); n-tile jj adds 8 to ci
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    sred[warp * 256 + g * 16 + 8 * jj + 2 * t] = ctot[jj][0];
    sred[warp * 256 + g * 16 + 8 * jj + 2 * t + 1] = ctot[jj][1];
    sred[warp * 256 + (g + 8) * 16 + 8 * jj + 2 * t] = ctot[jj][2];
    sred[warp * 256 + (g + 8) * 16 + 8 * jj + 2 * t + 1] = ctot[jj][3];
  }
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += sred[w * 256 + tid];
  const int o = tid >> 4, ci = tid & 15;
  if (o < nco && ci < nci) a.partials[((size_t)blockIdx.x * a.Cout + co0 + o) * a.Cin + ci0 + ci] = sum;
}

int pw_wgrad_workers(int N, int H, int W, int pairs) {
  const long long tiles = (long long)N * ocrs_cdiv(W, WTW) * ocrs_cdiv(H, WTH);
  long long per = (2 * OCRS_NUM_SMS + pairs - 1) / pairs;
  if (per < 1) per = 1;
  return (int)(tiles < per ? tiles : per);
}

// ---------------------------------------------------------------------------------------------------------------
// 1x1 weight gradient from the SAVED depthwise output: dWpw[co][ci] = sum_p dy[co][p] * dwo[ci][p]. The detection
// step on B200 is bound by instruction issue, not by HBM (ncu: 50-68 % issue-active at 25-40 % DRAM throughput), so
// the forward stores the depthwise output (4*Cin bytes / pixel) and this kernel streams both operands straight from
// global memory in mma.sync fragment layout: no halo, no stencil recompute, no shared memory until the final
// reduction. ~10 warp instructions per pixel instead of ~36 for the recomputing kernel above.
struct PwWg2Args {
  const float *d_a, *y, *dwo;
  long long da_ss, y_ss;
  const float *sc, *sh, *lo, *k1, *k2, *k3;
  float* partials;  // [gridDim.x][Cout][Cin]
  int Cout, Cin, N;
  long long HW;
  // optional fused 1x1 data gradient g[ci][p] = sum_co Wpw[co][ci] * dy[co][p] (Cout <= 16 only): the staged d_a / y
  // tile is already in shared memory, so ocrs_det_pwT_bwd's second pass over d_a and y disappears
  const float* wpw;  // [Cout][Cin] or null
  float* g; long long g_ss;
};
// Operands are staged by cp.async.cg (16 bytes = 4 pixels of one channel plane per request: every warp request
// covers 512 contiguous bytes of ONE plane, where gathering straight in fragment layout touched 8 planes per
// request and ran into the L1 tag rate) into a 4-stage ring of [48 planes][128 pixels] whose plane stride (132)
// makes the fragment reads (8 planes x 4 pixels per warp request) bank-conflict free.
constexpr int GSTAGES = 4;
// NPL = channel planes per operand group (d_a | y | dwo): 16, or 8 for the 8-channel blocks, which then take 256 pixels
// per stage in the same shared memory (their stages were too short to hide the per-stage barrier latency).
template <int NPL> struct GCfg {
  static constexpr int GPX = NPL == 8 ? 256 : 128;   // pixels per stage
  static constexpr int GPS = GPX + 4;                // plane stride in shared memory (= 4 mod 32)
  static constexpr int GPLANES = 3 * NPL;
  static constexpr int KSTEPS = GPX / 64;            // k-steps (8 pixels) per warp and stage
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NPL>
__global__ void __launch_bounds__(NTHREADS, 2)
pw_wgrad_saved_kernel(PwWg2Args a) {
  constexpr int GPX = GCfg<NPL>::GPX, GPS = GCfg<NPL>::GPS, GPLANES = GCfg<NPL>::GPLANES, KSTEPS = GCfg<NPL>::KSTEPS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stages = reinterpret_cast<float*>(smem_raw);  // [GSTAGES][GPLANES][GPS]; reused for the final reduction
  float* gs = stages + GSTAGES * GPLANES * GPS;        // [NPL][GPS] staging of the fused data gradient
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int cit = (a.Cin + 15) / 16;
  const int co0 = (blockIdx.y / cit) * 16, ci0 = (blockIdx.y % cit) * 16;
  const int nco = min(16, a.Cout - co0), nci = min(16, a.Cin - ci0);
  const bool do_g = a.wpw != nullptr;  // host guarantees Cout <= 16, i.e. co0 == 0
  // A fragments of the g-mma (M = ci, K = co with the slot permutation k = t -> co 2t, k = t + 4 -> co 2t + 1, which
  // makes the B reads below bank-conflict free): a0 (ci g, co 2t), a1 (ci g+8, co 2t), a2 (ci g, co 2t+1), a3 (ci g+8, co 2t+1)
  uint32_t wh[2][4], wl[2][4];
  if (do_g) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int ci = g + 8 * (q & 1), co = 8 * ks + 2 * t + (q >> 1);
        const float v = (ci < nci && co < nco) ? a.wpw[(size_t)co * a.Cin + ci0 + ci] : 0.f;
        tf32_split2(v, wh[ks][q], wl[ks][q]);
      }
  }
  float ksc[2], ksh[2], klo[2], kk1[2], kk2[2], kk3[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int o = g + 8 * h;
    const bool cov = o < nco;
    ksc[h] = cov ? a.sc[co0 + o] : 0.f; ksh[h] = cov ? a.sh[co0 + o] : 0.f; klo[h] = cov ? a.lo[co0 + o] : 0.f;
    kk1[h] = cov ? a.k1[co0 + o] : 0.f; kk2[h] = cov ? a.k2[co0 + o] : 0.f; kk3[h] = cov ? a.k3[co0 + o] : 0.f;
  }
  float ctot[2][4], c[2][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) { ctot[j][q] = 0.f; c[j][q] = 0.f; }
  const long long chunks_per_n = (a.HW + GPX - 1) / GPX;
  const long long total = chunks_per_n * a.N;
  const long long mine = blockIdx.x < total ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  // this thread's cp.async requests of a stage: planes warp + 8 r (r < 6), 16-byte column `lane`. The per-plane parts of
  // the addresses are computed once; per stage only the (sample, pixel) offset of each tensor changes.
  const float* pbase[GPLANES / 8];
  bool pok[GPLANES / 8];
#pragma unroll
  for (int r = 0; r < GPLANES / 8; ++r) {
    const int pl = warp + 8 * r, which = pl / NPL, ch = pl % NPL;  // which: 0 = d_a, 1 = y, 2 = dwo
    pok[r] = ch < (which == 2 ? nci : nco);
    pbase[r] = which == 0 ? a.d_a + (size_t)(co0 + ch) * a.HW : which == 1 ? a.y + (size_t)(co0 + ch) * a.HW
                                                                            : a.dwo + (size_t)(ci0 + ch) * a.HW;
  }
  const long long dwo_ss = (long long)a.Cin * a.HW;
  auto issue = [&](long long j) {
    const long long w = blockIdx.x + j * gridDim.x;
    const int n = (int)(w / chunks_per_n);
    const long long p0 = (w - (long long)n * chunks_per_n) * GPX + 4 * lane;
    const long long pc = p0 < a.HW ? p0 : 0;  // keep the (unread) source address inside the tensor
    const long long off[3] = {(long long)n * a.da_ss + pc, (long long)n * a.y_ss + pc, (long long)n * dwo_ss + pc};
    const uint32_t dst = tma::smem_u32(stages + (j % GSTAGES) * (GPLANES * GPS) + 4 * lane + warp * GPS);
#pragma unroll
    for (int r = 0; r < GPLANES / 8; ++r)
      if (pok[r]) {
        const int which = (warp + 8 * r) / NPL;
#pragma unroll
        for (int q = 0; q < GPX / 128; ++q) {  // 128 pixels = one 16-byte column per lane
          const long long rem = a.HW - (p0 + 128 * q);
          const int bytes = rem >= 4 ? 16 : (rem > 0 ? (int)rem * 4 : 0);
          cp_async16(dst + (r * 8 * GPS + 128 * q) * 4, pbase[r] + off[which] + (rem > 0 ? 128 * q : 0), bytes);  // bytes < 16 zero-fills
        }
      }
  };
  // planes of channels this CTA does not have (8-channel blocks, Cin = 1) are zeroed once and never requested
  for (int i = tid; i < GSTAGES * GPLANES * GPS; i += NTHREADS) {
    const int pl = (i / GPS) % GPLANES, ch = pl % NPL;
    if (ch >= (pl / NPL == 2 ? nci : nco)) stages[i] = 0.f;
  }
  __syncthreads();
  for (int j = 0; j < GSTAGES - 1; ++j) {
    if (j < mine) issue(j);
    cp_async_commit();
  }
  for (long long j = 0; j < mine; ++j) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();  // stage j landed for every thread; stage (j - 1) is free again
    if (j + GSTAGES - 1 < mine) issue(j + GSTAGES - 1);
    cp_async_commit();
    float* st = stages + (j % GSTAGES) * (GPLANES * GPS);
#pragma unroll
    for (int u = 0; u < KSTEPS; ++u) {  // warp w owns pixels 8 KSTEPS w .. 8 KSTEPS (w + 1) - 1 of the stage
      const int px = (warp * KSTEPS + u) * 8 + t;
      uint32_t ah[4], al[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // q: 0 = (g, t), 1 = (g+8, t), 2 = (g, t+4), 3 = (g+8, t+4)
        const int h = q & 1, e = q >> 1;
        if (8 * h >= NPL || 8 * h >= nco) { ah[q] = 0u; al[q] = 0u; continue; }  // 8-channel blocks use half of the M tile
        const float da = st[(g + 8 * h) * GPS + px + 4 * e];
        const float yv = st[(NPL + g + 8 * h) * GPS + px + 4 * e];
        const float dz = (fmaf(yv, ksc[h], ksh[h]) > klo[h]) ? da : 0.f;
        // channels past nco have kk* = 0; pixels past HW give dy = k3 but meet a zero-filled dwo
        const float dy = fmaf(kk1[h], dz, fmaf(kk2[h], yv, kk3[h]));
        tf32_split2(dy, ah[q], al[q]);
        if (do_g) st[(g + 8 * h) * GPS + px + 4 * e] = dy;  // in place over d_a: only this warp reads these pixels
      }
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        if (8 * jj >= NPL || 8 * jj >= nci) continue;
        uint32_t bh0, bl0, bh1, bl1;
        tf32_split2(st[(2 * NPL + 8 * jj + g) * GPS + px], bh0, bl0);      // (k = t,     n = ci 8jj + g)
        tf32_split2(st[(2 * NPL + 8 * jj + g) * GPS + px + 4], bh1, bl1);  // (k = t + 4, n = ci 8jj + g)
        mma_tf32_16n8k8(c[jj], al, bh0, bh1);
        mma_tf32_16n8k8(c[jj], ah, bl0, bl1);
        mma_tf32_16n8k8(c[jj], ah, bh0, bh1);
      }
    }
    if (do_g) {
      // g tile of this stage: warp w owns pixels 16w .. 16w+15 (two n-tiles of 8), all 16 input channels
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < KSTEPS; ++nt) {
        const int px = (warp * KSTEPS + nt) * 8;
        float cg[4] = {0.f, 0.f, 0.f, 0.f}, cx[4] = {0.f, 0.f, 0.f, 0.f};  // hi.hi and cross terms apart
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          if (8 * ks >= NPL || 8 * ks >= nco) continue;
          uint32_t bh[2], bl[2];
#pragma unroll
          for (int e = 0; e < 2; ++e)  // B: (k = t -> co 2t, k = t+4 -> co 2t+1; n = pixel g): dy written above
            tf32_split2(st[(8 * ks + 2 * t + e) * GPS + px + g], bh[e], bl[e]);
          mma_tf32_16n8k8(cx, wl[ks], bh[0], bh[1]);
          mma_tf32_16n8k8(cx, wh[ks], bl[0], bl[1]);
          mma_tf32_16n8k8(cg, wh[ks], bh[0], bh[1]);
        }
        // C fragment: c0 (ci g, px 2t), c1 (ci g, px 2t+1), c2 (ci g+8, px 2t), c3 (ci g+8, px 2t+1)
        *reinterpret_cast<float2*>(gs + g * GPS + px + 2 * t) = make_float2(cg[0] + cx[0], cg[1] + cx[1]);
        if (NPL > 8) *reinterpret_cast<float2*>(gs + (g + 8) * GPS + px + 2 * t) = make_float2(cg[2] + cx[2], cg[3] + cx[3]);
      }
      __syncthreads();  // gs complete (the loop-top barrier of the next stage protects its reuse)
      {
        const long long w = blockIdx.x + j * gridDim.x;
        const int n = (int)(w / chunks_per_n);
        const long long p0 = (w - (long long)n * chunks_per_n) * GPX + 4 * lane;
#pragma unroll
        for (int r = 0; r < NPL / 8; ++r) {
          const int ci = warp + 8 * r;
#pragma unroll
          for (int q = 0; q < GPX / 128; ++q)
            if (ci < nci && p0 + 128 * q < a.HW)  // HW % 4 == 0: the four pixels are in or out together
              *reinterpret_cast<float4*>(a.g + (size_t)n * a.g_ss + (size_t)(ci0 + ci) * a.HW + p0 + 128 * q) =
                  *reinterpret_cast<const float4*>(gs + ci * GPS + 4 * lane + 128 * q);
        }
      }
    }
    if ((j & (NPL == 8 ? 1 : 3)) == (NPL == 8 ? 1 : 3)) {  // keep the tensor core's truncating accumulation chains short (8 k-steps)
#pragma unroll
      for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int q = 0; q < 4; ++q) { ctot[jj][q] += c[jj][q]; c[jj][q] = 0.f; }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  float* sred = stages;  // [8][256]
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    sred[warp * 256 + g * 16 + 8 * jj + 2 * t] = ctot[jj][0] + c[jj][0];
    sred[warp * 256 + g * 16 + 8 * jj + 2 * t + 1] = ctot[jj][1] + c[jj][1];
    sred[warp * 256 + (g + 8) * 16 + 8 * jj + 2 * t] = ctot[jj][2] + c[jj][2];
    sred[warp * 256 + (g + 8) * 16 + 8 * jj + 2 * t + 1] = ctot[jj][3] + c[jj][3];
  }
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += sred[w * 256 + tid];
  const int o = tid >> 4, ci = tid & 15;
  if (o < nco && ci < nci) a.partials[((size_t)blockIdx.x * a.Cout + co0 + o) * a.Cin + ci0 + ci] = sum;
}


// ---------------------------------------------------------------------------------------------------------------
// The same for the blocks with 32 output channels (16->32, 32->32, 64->32: 8 of the 26 blocks). One CTA holds ALL 32
// output channels of a 64-pixel stage (d_a 32 | y 32 | dwo 16 planes) and one 16-channel slice of the input channels,
// so the 1x1 data gradient of its slice, g[ci][p] = sum_co Wpw[co][ci] * dy[co][p], is complete inside the CTA and
// ocrs_det_pwT_bwd's separate pass over d_a and y (3.8 ms per detection step) disappears. The CTAs of the other input
// slices re-read the same d_a / y chunk at the same time (L2 hits).
constexpr int G32_PX = 64, G32_PS = G32_PX + 4, G32_PLANES = 80, G32_STAGE = G32_PLANES * G32_PS;

__global__ void __launch_bounds__(NTHREADS, 2)
pw_wgrad_saved32_kernel(PwWg2Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stages = reinterpret_cast<float*>(smem_raw);  // [GSTAGES][80][G32_PS]; reused for the final reduction
  float* gs = stages + GSTAGES * G32_STAGE;            // [16][G32_PS] staging of the data gradient
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int ci0 = blockIdx.y * 16;
  // A fragments of the g-mma (M = ci, K = co in four k-steps, slot permutation k = t -> co 2t, k = t + 4 -> co 2t + 1)
  uint32_t wh[4][4], wl[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ci = g + 8 * (q & 1), co = 8 * ks + 2 * t + (q >> 1);
      tf32_split2(a.wpw[(size_t)co * a.Cin + ci0 + ci], wh[ks][q], wl[ks][q]);
    }
  float ksc[4], ksh[4], klo[4], kk1[4], kk2[4], kk3[4];  // rows g + 8 h of the 32 output channels
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const int o = g + 8 * h;
    ksc[h] = a.sc[o]; ksh[h] = a.sh[o]; klo[h] = a.lo[o]; kk1[h] = a.k1[o]; kk2[h] = a.k2[o]; kk3[h] = a.k3[o];
  }
  float ctot[2][2][4], c[2][2][4];  // [co tile][ci tile]
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) { ctot[m][j][q] = 0.f; c[m][j][q] = 0.f; }
  const long long chunks_per_n = (a.HW + G32_PX - 1) / G32_PX;
  const long long total = chunks_per_n * a.N;
  const long long mine = blockIdx.x < total ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  // cp.async requests of a stage: 80 planes x 16 columns of 16 bytes = 1280 = 5 per thread; request id = tid + 256 r
  const float* pbase[5];
  int pwhich[5];
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const int pl = (tid + 256 * r) >> 4;
    pwhich[r] = pl < 32 ? 0 : (pl < 64 ? 1 : 2);
    pbase[r] = pl < 32 ? a.d_a + (size_t)pl * a.HW : pl < 64 ? a.y + (size_t)(pl - 32) * a.HW : a.dwo + (size_t)(ci0 + pl - 64) * a.HW;
  }
  const long long dwo_ss = (long long)a.Cin * a.HW;
  auto issue = [&](long long j) {
    const long long w = blockIdx.x + j * gridDim.x;
    const int n = (int)(w / chunks_per_n);
    const long long p0 = (w - (long long)n * chunks_per_n) * G32_PX + 4 * (tid & 15);
    const long long pc = p0 < a.HW ? p0 : 0;  // keep the (unread) source address inside the tensor
    const long long off[3] = {(long long)n * a.da_ss + pc, (long long)n * a.y_ss + pc, (long long)n * dwo_ss + pc};
    const long long rem = a.HW - p0;
    const int bytes = rem >= 4 ? 16 : (rem > 0 ? (int)rem * 4 : 0);  // < 16 zero-fills
    const uint32_t dst = tma::smem_u32(stages + (j % GSTAGES) * G32_STAGE + (tid >> 4) * G32_PS + 4 * (tid & 15));
#pragma unroll
    for (int r = 0; r < 5; ++r) cp_async16(dst + r * 16 * G32_PS * 4, pbase[r] + off[pwhich[r]], bytes);
  };
  for (int j = 0; j < GSTAGES - 1; ++j) {
    if (j < mine) issue(j);
    cp_async_commit();
  }
  for (long long j = 0; j < mine; ++j) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();  // stage j landed for every thread; stage (j - 1) and gs are free again
    if (j + GSTAGES - 1 < mine) issue(j + GSTAGES - 1);
    cp_async_commit();
    float* st = stages + (j % GSTAGES) * G32_STAGE;
    const int px = warp * 8 + t;  // warp w owns pixels 8w .. 8w+7 of the stage (one k-step of the weight gradient)
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // q: 0 = (g, t), 1 = (g+8, t), 2 = (g, t+4), 3 = (g+8, t+4)
        const int h = 2 * m + (q & 1), e = q >> 1, o = g + 8 * h;
        const float da = st[o * G32_PS + px + 4 * e];
        const float yv = st[(32 + o) * G32_PS + px + 4 * e];
        const float dz = (fmaf(yv, ksc[h], ksh[h]) > klo[h]) ? da : 0.f;
        const float dy = fmaf(kk1[h], dz, fmaf(kk2[h], yv, kk3[h]));  // pixels past HW: dy = k3, but dwo is zero-filled and g is not stored
        tf32_split2(dy, ah[m][q], al[m][q]);
        st[o * G32_PS + px + 4 * e] = dy;  // in place over d_a: only this warp reads these pixels
      }
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      uint32_t bh0, bl0, bh1, bl1;
      tf32_split2(st[(64 + 8 * jj + g) * G32_PS + px], bh0, bl0);      // (k = t,     n = ci 8jj + g)
      tf32_split2(st[(64 + 8 * jj + g) * G32_PS + px + 4], bh1, bl1);  // (k = t + 4, n = ci 8jj + g)
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        mma_tf32_16n8k8(c[m][jj], al[m], bh0, bh1);
        mma_tf32_16n8k8(c[m][jj], ah[m], bl0, bl1);
        mma_tf32_16n8k8(c[m][jj], ah[m], bh0, bh1);
      }
    }
    // data gradient of this warp's 8 pixels: M = 16 input channels, N = 8 pixels, K = 32 output channels
    __syncwarp();
    {
      float cg[4] = {0.f, 0.f, 0.f, 0.f}, cx[4] = {0.f, 0.f, 0.f, 0.f};  // hi.hi and cross terms apart
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bh[2], bl[2];
#pragma unroll
        for (int e = 0; e < 2; ++e)  // B: (k = t -> co 8ks + 2t, k = t+4 -> co 8ks + 2t + 1; n = pixel g)
          tf32_split2(st[(8 * ks + 2 * t + e) * G32_PS + warp * 8 + g], bh[e], bl[e]);
        mma_tf32_16n8k8(cx, wl[ks], bh[0], bh[1]);
        mma_tf32_16n8k8(cx, wh[ks], bl[0], bl[1]);
        mma_tf32_16n8k8(cg, wh[ks], bh[0], bh[1]);
      }
      // C fragment: c0 (ci g, px 2t), c1 (ci g, px 2t+1), c2 (ci g+8, px 2t), c3 (ci g+8, px 2t+1)
      *reinterpret_cast<float2*>(gs + g * G32_PS + warp * 8 + 2 * t) = make_float2(cg[0] + cx[0], cg[1] + cx[1]);
      *reinterpret_cast<float2*>(gs + (g + 8) * G32_PS + warp * 8 + 2 * t) = make_float2(cg[2] + cx[2], cg[3] + cx[3]);
    }
    __syncthreads();  // gs complete
    {
      const long long w = blockIdx.x + j * gridDim.x;
      const int n = (int)(w / chunks_per_n);
      const long long p0 = (w - (long long)n * chunks_per_n) * G32_PX + 4 * (tid & 15);
      const int ci = tid >> 4;
      if (p0 < a.HW)  // HW % 4 == 0: the four pixels are in or out together
        *reinterpret_cast<float4*>(a.g + (size_t)n * a.g_ss + (size_t)(ci0 + ci) * a.HW + p0) =
            *reinterpret_cast<const float4*>(gs + ci * G32_PS + 4 * (tid & 15));
    }
    if ((j & 7) == 7) {  // keep the tensor core's truncating accumulation chains short (8 k-steps)
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
          for (int q = 0; q < 4; ++q) { ctot[m][jj][q] += c[m][jj][q]; c[m][jj][q] = 0.f; }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  float* sred = stages;  // [8][32 * 16]
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      float* r0 = sred + warp * 512 + (16 * m + g) * 16 + 8 * jj + 2 * t;
      r0[0] = ctot[m][jj][0] + c[m][jj][0];
      r0[1] = ctot[m][jj][1] + c[m][jj][1];
      r0[8 * 16] = ctot[m][jj][2] + c[m][jj][2];
      r0[8 * 16 + 1] = ctot[m][jj][3] + c[m][jj][3];
    }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = tid + 256 * r;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += sred[w * 512 + i];
    a.partials[((size_t)blockIdx.x * 32 + (i >> 4)) * a.Cin + ci0 + (i & 15)] = sum;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// ConvTranspose2d(k3, s2) weight gradient (reference models.py:76-78):
//   dW[ci][co][ky][kx] = sum_{n,iy,ix} xact[n][ci][iy][ix] * dout[n][co][2iy+ky][2ix+kx]
// = nine skinny GEMMs that share the A operand (xact, M = 16 input channels, K = input pixels) and take their B
// operands from nine stride-2 views of the same dout tile. The tile (4 input rows x 64 input columns of 16 input
// channels; 9 x 132 output pixels of 8 output channels) is staged by cp.async in plane-contiguous 16-byte
// requests (2-stage ring); fragments come from shared memory; mma.sync m16n8k8 3xTF32 like the 1x1 kernels.
// The old kernel gathered both operands from global memory in fragment layout (8 planes and stride-2 columns per
// warp request) and re-read x once per 16 (co, tap) pairs.
constexpr int CTR = 4, CTC = 64;             // input rows / columns per tile
constexpr int CT_XPS = CTR * CTC + 4;        // x plane stride (260 = 4 mod 32: conflict-free A fragments)
constexpr int CT_DROWS = 2 * CTR + 1;        // dout rows per tile
constexpr int CT_DCOLS = 2 * CTC + 4;        // dout columns per tile (129 used, 16-byte multiple)
constexpr int CT_DPS = CT_DROWS * CT_DCOLS + 8;   // dout plane stride
constexpr int CT_STAGE = 16 * CT_XPS + 8 * CT_DPS;

struct ConvtWgArgs {
  const float *x, *dout;
  long long x_ss, dout_ss;
  const float *isc, *ish, *ilo;
  float* partials;  // [gridDim.x][Cin][Cout][9]
  int Cin, Cout, Hin, Win, Hs, Ws, N;
};

__global__ void __launch_bounds__(NTHREADS, 2)
convt_wgrad_staged_kernel(ConvtWgArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stages = reinterpret_cast<float*>(smem_raw);  // [2][CT_STAGE]; reused for the final reduction
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int cot = (a.Cout + 7) / 8;
  const int ci0 = (blockIdx.y / cot) * 16, co0 = (blockIdx.y % cot) * 8;
  const int nci = min(16, a.Cin - ci0), nco = min(8, a.Cout - co0);
  float sc[2], sh[2], lo[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const bool v = g + 8 * h < nci && a.isc != nullptr;
    sc[h] = v ? a.isc[ci0 + g + 8 * h] : 1.f;
    sh[h] = v ? a.ish[ci0 + g + 8 * h] : 0.f;
    lo[h] = v ? a.ilo[ci0 + g + 8 * h] : -INFINITY;
  }
  const int tiles_x = (a.Win + CTC - 1) / CTC, tiles_y = (a.Hin + CTR - 1) / CTR;
  const long long total = (long long)a.N * tiles_x * tiles_y;
  const long long mine = blockIdx.x < total ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto decode = [&](long long j, int& n, int& ix0, int& iy0) {
    const long long w = blockIdx.x + j * gridDim.x;
    n = (int)(w / (tiles_x * tiles_y));
    const int r = (int)(w - (long long)n * tiles_x * tiles_y);
    iy0 = (r / tiles_x) * CTR;
    ix0 = (r % tiles_x) * CTC;
  };
  auto issue = [&](long long j) {
    int n, ix0, iy0;
    decode(j, n, ix0, iy0);
    float* st = stages + (j & 1) * CT_STAGE;
    // x: nci planes x CTR rows x 16 requests; requests past the image are zero-filled (src_bytes 0)
    for (int e = tid; e < nci * CTR * (CTC / 4); e += NTHREADS) {
      const int c = e / (CTR * (CTC / 4)), r = (e / (CTC / 4)) % CTR, q = e % (CTC / 4);
      const int iy = iy0 + r, ix = ix0 + 4 * q;
      const bool ok = iy < a.Hin && ix < a.Win;  // Win % 4 == 0: a request is inside or outside as a whole
      const float* src = a.x + (size_t)n * a.x_ss + ((size_t)(ci0 + c) * a.Hin + (ok ? iy : 0)) * a.Win + (ok ? ix : 0);
      cp_async16(tma::smem_u32(st + c * CT_XPS + r * CTC + 4 * q), src, ok ? 16 : 0);
    }
    // dout: nco planes x 9 rows x 33 requests starting at (2 iy0, 2 ix0)
    float* sd = st + 16 * CT_XPS;
    for (int e = tid; e < nco * CT_DROWS * (CT_DCOLS / 4); e += NTHREADS) {
      const int c = e / (CT_DROWS * (CT_DCOLS / 4)), r = (e / (CT_DCOLS / 4)) % CT_DROWS, q = e % (CT_DCOLS / 4);
      const int oy = 2 * iy0 + r, ox = 2 * ix0 + 4 * q;
      const bool ok = oy < a.Hs && ox < a.Ws;
      const int rem = ok ? a.Ws - ox : 0;  // Ws % 4 == 0 keeps this a multiple of 4; the guard is for safety
      const float* src = a.dout + (size_t)n * a.dout_ss + ((size_t)(co0 + c) * a.Hs + (ok ? oy : 0)) * a.Ws + (ok ? ox : 0);
      cp_async16(tma::smem_u32(sd + c * CT_DPS + r * CT_DCOLS + 4 * q), src, rem >= 4 ? 16 : rem * 4);
    }
  };
  // planes of absent channels are zeroed once (their requests are never issued)
  for (int i = tid; i < 2 * CT_STAGE; i += NTHREADS) stages[i] = 0.f;
  __syncthreads();
  if (mine > 0) issue(0);
  cp_async_commit();
  float ctot[9][4], c[9][4];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int q = 0; q < 4; ++q) { ctot[k][q] = 0.f; c[k][q] = 0.f; }
  for (long long j = 0; j < mine; ++j) {
    cp_async_wait<0>();
    __syncthreads();  // stage j landed; stage j^1 is free (everyone finished tile j-1)
    if (j + 1 < mine) issue(j + 1);
    cp_async_commit();
    int n, ix0, iy0;
    decode(j, n, ix0, iy0);
    const float* st = stages + (j & 1) * CT_STAGE;
    const float* sd = st + 16 * CT_XPS;
    // 32 k-steps of 8 input pixels (row r = ks / 8, columns 8 (ks % 8) ..), 4 per warp
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int ks = warp * 4 + u, r = ks >> 3, cx = (ks & 7) * 8;
      const bool rowok = iy0 + r < a.Hin;
      uint32_t ah[4], al[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // q: 0 = (ci g, px t), 1 = (g+8, t), 2 = (g, t+4), 3 = (g+8, t+4)
        const int h = q & 1, e = q >> 1;
        const int px = cx + t + 4 * e;
        float v = 0.f;
        if (rowok && ix0 + px < a.Win && g + 8 * h < nci) v = xform_apply(st[(g + 8 * h) * CT_XPS + r * CTC + px], sc[h], sh[h], lo[h]);
        tf32_split2(v, ah[q], al[q]);
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const int ky = k / 3, kx = k - 3 * ky;
        // B (k = px t / t+4, n = co g): dout[co g][2r + ky][2 px + kx]
        const float* bp = sd + g * CT_DPS + (2 * r + ky) * CT_DCOLS + 2 * (cx + t) + kx;
        uint32_t bh0, bl0, bh1, bl1;
        tf32_split2(bp[0], bh0, bl0);
        tf32_split2(bp[8], bh1, bl1);
        mma_tf32_16n8k8(c[k], al, bh0, bh1);
        mma_tf32_16n8k8(c[k], ah, bl0, bl1);
        mma_tf32_16n8k8(c[k], ah, bh0, bh1);
      }
    }
    if ((j & 1) == 1) {  // flush: the tensor core's truncating accumulation sees chains of 8 k-steps
#pragma unroll
      for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) { ctot[k][q] += c[k][q]; c[k][q] = 0.f; }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // CTA reduction: sred[warp][ci 16][co 8][tap 9]
  float* sred = stages;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float v0 = ctot[k][0] + c[k][0], v1 = ctot[k][1] + c[k][1], v2 = ctot[k][2] + c[k][2], v3 = ctot[k][3] + c[k][3];
    sred[warp * 1152 + (g * 8 + 2 * t) * 9 + k] = v0;          // (ci g, co 2t)
    sred[warp * 1152 + (g * 8 + 2 * t + 1) * 9 + k] = v1;      // (ci g, co 2t+1)
    sred[warp * 1152 + ((g + 8) * 8 + 2 * t) * 9 + k] = v2;    // (ci g+8, co 2t)
    sred[warp * 1152 + ((g + 8) * 8 + 2 * t + 1) * 9 + k] = v3;
  }
  __syncthreads();
  for (int i = tid; i < 1152; i += NTHREADS) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += sred[w * 1152 + i];
    const int ci = i / 72, co = (i / 9) % 8, k = i % 9;
    if (ci < nci && co < nco)
      a.partials[(((size_t)blockIdx.x * a.Cin + ci0 + ci) * a.Cout + co0 + co) * 9 + k] = sum;
  }
}
}  // namespace

extern "C" {

// 1 when the TMA-pipelined DepthwiseConv kernels can run on these views (16-byte aligned planes, W % 4 == 0).
int ocrs_det_tma_supported(const float* x, long long x_ss, const float* y, long long y_ss, int H, int W) {
  return ocrs_plane_tma_ok(x, x_ss, H, W) && ocrs_plane_tma_ok(y, y_ss, H, W) && H >= 8 && W >= 8;
}

// 1 when ocrs_det_sep_fwd handles these channel counts (Cout a multiple of its 8/16-channel tile, Cin < 4 or a multiple of 4).
int ocrs_det_sep_channels_ok(int Cin, int Cout) {
  const int cot = Cout <= 8 ? 8 : 16;
  return Cout % cot == 0 && (Cin < 4 || Cin % 4 == 0) && Cin <= FWD_MAX_CIN;
}

// Rows of the [rows][2][Cout] statistics partials ocrs_det_sep_fwd writes.
int ocrs_det_sep_fwd_rows(int N, int H, int W, int Cout) {
  return fwd_ctas(N, H, W, ocrs_cdiv(Cout, Cout <= 8 ? 8 : 16));
}

// DepthwiseConv block body (reference models.py:11-22), TMA-pipelined: y = pw1x1(dw3x3(xform(x))) and the
// BatchNorm partial sums of y. Same contract as ocrs_det_dwpw_fwd except for the partial-row count; when dw_out
// ([N][Cin][H][W], contiguous) is given, the depthwise output is stored too, for ocrs_det_pw_wgrad_saved.
int ocrs_det_sep_fwd(const float* x, long long x_ss, int N, int Cin, int H, int W, const float* in_scale,
                     const float* in_shift, const float* in_lo, const float* wdw, const float* wpw, int Cout,
                     float* y, long long y_ss, float* partials, float* dw_out, void* stream) {
  OCRS_CHECK_ARG(N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "sep_fwd: bad dims");
  OCRS_CHECK_ARG(ocrs_det_tma_supported(x, x_ss, y, y_ss, H, W), "sep_fwd: views are not TMA-addressable");
  const int cot = Cout <= 8 ? 8 : 16;
  OCRS_CHECK_ARG(Cout % cot == 0 && (Cin < 4 || Cin % 4 == 0) && Cin <= FWD_MAX_CIN,
                 "sep_fwd: channel counts %d -> %d unsupported (ocrs_det_sep_channels_ok)", Cin, Cout);
  FwdArgs a;
  a.Cin = Cin; a.Cout = Cout; a.H = H; a.W = W; a.N = N; a.n_cot = ocrs_cdiv(Cout, cot);
  a.in_scale = in_scale; a.in_shift = in_shift; a.in_lo = in_lo; a.wdw = wdw; a.wpw = wpw;
  a.y = y; a.y_ss = y_ss; a.partials = partials; a.dwo = dw_out;
  a.walk.tiles_x = ocrs_cdiv(W, TW); a.walk.tiles_y = ocrs_cdiv(H, TH);
  a.walk.sp_total = N * a.walk.tiles_x * a.walk.tiles_y; a.walk.sp0 = 0; a.walk.sp_stride = 1;
  const int ctas = fwd_ctas(N, H, W, a.n_cot) * a.n_cot;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap xmap;
  if (Cin < 4) {
    if (ocrs_plane_map(&xmap, x, x_ss, N, Cin, H, W, BW, BH, 1)) return -1;
    const size_t smem = fwd_smem<8, 1>(Cin);
    if (cot == 8) {
      OCRS_SET_SMEM_ONCE((sep_fwd_tma_kernel<8, 1>), (fwd_smem<8, 1>(FWD_MAX_CIN)));
      sep_fwd_tma_kernel<8, 1><<<ctas, NTHREADS, smem, st>>>(xmap, a);
    } else {
      const size_t smem16 = fwd_smem<16, 1>(Cin);
      OCRS_SET_SMEM_ONCE((sep_fwd_tma_kernel<16, 1>), (fwd_smem<16, 1>(FWD_MAX_CIN)));
      sep_fwd_tma_kernel<16, 1><<<ctas, NTHREADS, smem16, st>>>(xmap, a);
    }
  } else {
    if (ocrs_plane_map(&xmap, x, x_ss, N, Cin, H, W, BW, BH, 4)) return -1;
    if (cot == 8) {
      const size_t smem = fwd_smem<8, 4>(Cin);
      OCRS_SET_SMEM_ONCE((sep_fwd_tma_kernel<8, 4>), (fwd_smem<8, 4>(FWD_MAX_CIN)));
      sep_fwd_tma_kernel<8, 4><<<ctas, NTHREADS, smem, st>>>(xmap, a);
    } else {
      const size_t smem = fwd_smem<16, 4>(Cin);
      OCRS_SET_SMEM_ONCE((sep_fwd_tma_kernel<16, 4>), (fwd_smem<16, 4>(FWD_MAX_CIN)));
      sep_fwd_tma_kernel<16, 4><<<ctas, NTHREADS, smem, st>>>(xmap, a);
    }
  }
  OCRS_CHECK_LAUNCH("sep_fwd_tma_kernel");
  return 0;
}

// Rows of the [rows][C][9] weight-gradient partials (and [rows][2][C] BatchNorm partials) of ocrs_det_sep_dw_bwd.
int ocrs_det_sep_dw_bwd_rows(int N, int H, int W, int C) { return dw_ctas_per_chunk(N, H, W, C); }

// Depthwise 3x3 backward, TMA-pipelined: dx (+= when accumulate) and the dw weight-gradient partials, like
// ocrs_det_dw_bwd; when up_mean/up_invstd/bn_partials are given it also emits the BatchNorm-backward sums
// (sum dz, sum dz*yhat) of the block that produced x, computed from the final dx (replaces a separate
// ocrs_bnrelu_bwd_reduce pass over d_a and y of that block).
int ocrs_det_sep_dw_bwd(const float* g, long long g_ss, const float* x, long long x_ss, int N, int C, int H, int W,
                        const float* isc, const float* ish, const float* ilo, const float* wdw, float* dx,
                        long long dx_ss, int accumulate, float* w_partials, const float* up_mean,
                        const float* up_invstd, float* bn_partials, void* stream) {
  OCRS_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0, "sep_dw_bwd: bad dims");
  OCRS_CHECK_ARG(ocrs_det_tma_supported(g, g_ss, x, x_ss, H, W) && ocrs_plane_tma_ok(dx, dx_ss, H, W),
                 "sep_dw_bwd: views are not TMA-addressable");
  OCRS_CHECK_ARG(bn_partials == nullptr || (up_mean && up_invstd && isc), "sep_dw_bwd: fused BN reduction needs the upstream statistics");
  DwBwdArgs a;
  a.C = C; a.H = H; a.W = W; a.N = N; a.accumulate = accumulate;
  a.ctas_per_chunk = dw_ctas_per_chunk(N, H, W, C);
  a.isc = isc; a.ish = ish; a.ilo = ilo; a.wdw = wdw; a.up_mean = up_mean; a.up_invstd = up_invstd;
  a.dx = dx; a.dx_ss = dx_ss; a.wpart = w_partials; a.bnpart = bn_partials;
  a.walk.tiles_x = ocrs_cdiv(W, TW); a.walk.tiles_y = ocrs_cdiv(H, TH);
  a.walk.sp_total = N * a.walk.tiles_x * a.walk.tiles_y; a.walk.sp0 = 0; a.walk.sp_stride = 1;
  CUtensorMap gmap, xmap;
  if (ocrs_plane_map(&gmap, g, g_ss, N, C, H, W, BW, BH, DCH)) return -1;
  if (ocrs_plane_map(&xmap, x, x_ss, N, C, H, W, BW, BH, DCH)) return -1;
  const size_t smem = (size_t)NSTAGE * 2 * box_floats(DCH) * 4 + 64 + 8 * DCH * 11 * 4 + 128;
  OCRS_SET_SMEM_ONCE(sep_dw_bwd_tma_kernel, smem);
  const int ctas = a.ctas_per_chunk * ocrs_cdiv(C, DCH);
  sep_dw_bwd_tma_kernel<<<ctas, NTHREADS, smem, (cudaStream_t)stream>>>(gmap, xmap, a);
  OCRS_CHECK_LAUNCH("sep_dw_bwd_tma_kernel");
  return 0;
}

// Rows ("workers") of the [workers][Cout][Cin] partials of ocrs_det_sep_pw_wgrad.
int ocrs_det_sep_pw_wgrad_workers(int N, int H, int W, int Cout, int Cin) {
  return pw_wgrad_workers(N, H, W, ocrs_cdiv(Cout, 16) * ocrs_cdiv(Cin, 16));
}

// 1x1-convolution weight gradient of a DepthwiseConv block, TMA-staged; same contract as ocrs_det_pw_wgrad
// (partials fully written; reduce with ocrs_finalize_partials).
int ocrs_det_sep_pw_wgrad(const float* d_a, long long da_ss, const float* y, long long y_ss, int N, int Cout, int H,
                          int W, const float* sc, const float* sh, const float* lo, const float* k1, const float* k2,
                          const float* k3, const float* x, long long x_ss, int Cin, const float* isc, const float* ish,
                          const float* ilo, const float* wdw, float* partials, void* stream) {
  OCRS_CHECK_ARG(N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "sep_pw_wgrad: bad dims");
  OCRS_CHECK_ARG(ocrs_plane_tma_ok(x, x_ss, H, W), "sep_pw_wgrad: x is not TMA-addressable");
  PwWgArgs a;
  a.d_a = d_a; a.y = y; a.da_ss = da_ss; a.y_ss = y_ss;
  a.sc = sc; a.sh = sh; a.lo = lo; a.k1 = k1; a.k2 = k2; a.k3 = k3;
  a.isc = isc; a.ish = ish; a.ilo = ilo; a.wdw = wdw; a.partials = partials;
  a.Cout = Cout; a.Cin = Cin; a.H = H; a.W = W; a.N = N;
  a.tiles_x = ocrs_cdiv(W, WTW); a.tiles_y = ocrs_cdiv(H, WTH);
  CUtensorMap xmap;
  if (ocrs_plane_map(&xmap, x, x_ss, N, Cin, H, W, BW, WBH, WCH)) return -1;
  const int pairs = ocrs_cdiv(Cout, 16) * ocrs_cdiv(Cin, 16);
  const size_t smem = (size_t)2 * box_floats2(WCH * WPLANE) * 4 + 64 + (size_t)WCH * XS_PLANE * 4 + 8 * 256 * 4 + 48 * 4 + 128;
  OCRS_SET_SMEM_ONCE(sep_pw_wgrad_tma_kernel, smem);
  dim3 grid(pw_wgrad_workers(N, H, W, pairs), pairs);
  sep_pw_wgrad_tma_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(xmap, a);
  OCRS_CHECK_LAUNCH("sep_pw_wgrad_tma_kernel");
  return 0;
}

// Rows of the [workers][Cout][Cin] partials of ocrs_det_pw_wgrad_saved.
int ocrs_det_pw_wgrad_saved_workers(int N, long long HW, int Cout, int Cin) {
  const int pairs = ocrs_cdiv(Cout, 16) * ocrs_cdiv(Cin, 16);
  const int gpx = (Cout <= 8 && Cin <= 8) ? GCfg<8>::GPX : GCfg<16>::GPX;
  const long long items = (long long)N * ((HW + gpx - 1) / gpx);
  long long per = (2 * OCRS_NUM_SMS + pairs - 1) / pairs;
  if (per > items) per = items;
  return (int)(per < 1 ? 1 : per);
}

// 1x1-convolution weight gradient from the depthwise output saved by ocrs_det_sep_fwd (dw_out, [N][Cin][HW]
// contiguous): partials [workers][Cout][Cin] fully written. Needs 16-byte aligned planes (HW % 4 == 0).
// With wpw ([Cout][Cin], Cout <= 16) it also writes the 1x1 data gradient g[ci][p] = sum_co wpw[co][ci] * dy[co][p]
// (what ocrs_det_pwT_bwd computes) from the same staged tile.
int ocrs_det_pw_wgrad_saved(const float* d_a, long long da_ss, const float* y, long long y_ss, int N, int Cout,
                            long long HW, const float* sc, const float* sh, const float* lo, const float* k1,
                            const float* k2, const float* k3, const float* dw_out, int Cin, float* partials,
                            const float* wpw, float* g, long long g_ss, void* stream) {
  OCRS_CHECK_ARG(N > 0 && Cin > 0 && Cout > 0 && HW > 0, "pw_wgrad_saved: bad dims");
  OCRS_CHECK_ARG(wpw == nullptr || (Cout <= 16 && g != nullptr && g_ss % 4 == 0 && (uintptr_t)g % 16 == 0),
                 "pw_wgrad_saved: the fused data gradient needs Cout <= 16 and a 16-byte aligned g");
  OCRS_CHECK_ARG(HW % 4 == 0 && da_ss % 4 == 0 && y_ss % 4 == 0 && (uintptr_t)d_a % 16 == 0 && (uintptr_t)y % 16 == 0 &&
                     (uintptr_t)dw_out % 16 == 0, "pw_wgrad_saved: planes must be 16-byte aligned (HW %% 4 == 0)");
  PwWg2Args a;
  a.d_a = d_a; a.y = y; a.dwo = dw_out; a.da_ss = da_ss; a.y_ss = y_ss;
  a.sc = sc; a.sh = sh; a.lo = lo; a.k1 = k1; a.k2 = k2; a.k3 = k3; a.partials = partials;
  a.Cout = Cout; a.Cin = Cin; a.N = N; a.HW = HW;
  a.wpw = wpw; a.g = g; a.g_ss = g_ss;
  dim3 grid(ocrs_det_pw_wgrad_saved_workers(N, HW, Cout, Cin), ocrs_cdiv(Cout, 16) * ocrs_cdiv(Cin, 16));
  if (Cout <= 8 && Cin <= 8) {
    const size_t smem = (size_t)(GSTAGES * GCfg<8>::GPLANES + 8) * GCfg<8>::GPS * 4;
    OCRS_SET_SMEM_ONCE(pw_wgrad_saved_kernel<8>, smem);
    pw_wgrad_saved_kernel<8><<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(a);
  } else {
    const size_t smem = (size_t)(GSTAGES * GCfg<16>::GPLANES + 16) * GCfg<16>::GPS * 4;
    OCRS_SET_SMEM_ONCE(pw_wgrad_saved_kernel<16>, smem);
    pw_wgrad_saved_kernel<16><<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(a);
  }
  OCRS_CHECK_LAUNCH("pw_wgrad_saved_kernel");
  return 0;
}

// Rows of the [workers][32][Cin] partials of ocrs_det_pw_wgrad_saved32.
int ocrs_det_pw_wgrad_saved32_workers(int N, long long HW, int Cin) {
  const int slices = ocrs_cdiv(Cin, 16);
  const long long items = (long long)N * ((HW + G32_PX - 1) / G32_PX);
  long long per = (2 * OCRS_NUM_SMS + slices - 1) / slices;
  if (per > items) per = items;
  return (int)(per < 1 ? 1 : per);
}

// ocrs_det_pw_wgrad_saved for the blocks with exactly 32 output channels (Cin a multiple of 16), always with the fused
// 1x1 data gradient: partials [workers][32][Cin] fully written, g [N][Cin][HW] view written.
int ocrs_det_pw_wgrad_saved32(const float* d_a, long long da_ss, const float* y, long long y_ss, int N, long long HW,
                              const float* sc, const float* sh, const float* lo, const float* k1, const float* k2,
                              const float* k3, const float* dw_out, int Cin, float* partials, const float* wpw, float* g,
                              long long g_ss, void* stream) {
  OCRS_CHECK_ARG(N > 0 && Cin > 0 && Cin % 16 == 0 && HW > 0, "pw_wgrad_saved32: Cin %d must be a multiple of 16", Cin);
  OCRS_CHECK_ARG(wpw != nullptr && g != nullptr && g_ss % 4 == 0 && (uintptr_t)g % 16 == 0, "pw_wgrad_saved32: needs wpw and a 16-byte aligned g");
  OCRS_CHECK_ARG(HW % 4 == 0 && da_ss % 4 == 0 && y_ss % 4 == 0 && (uintptr_t)d_a % 16 == 0 && (uintptr_t)y % 16 == 0 &&
                     (uintptr_t)dw_out % 16 == 0, "pw_wgrad_saved32: planes must be 16-byte aligned (HW %% 4 == 0)");
  PwWg2Args a;
  a.d_a = d_a; a.y = y; a.dwo = dw_out; a.da_ss = da_ss; a.y_ss = y_ss;
  a.sc = sc; a.sh = sh; a.lo = lo; a.k1 = k1; a.k2 = k2; a.k3 = k3; a.partials = partials;
  a.Cout = 32; a.Cin = Cin; a.N = N; a.HW = HW;
  a.wpw = wpw; a.g = g; a.g_ss = g_ss;
  dim3 grid(ocrs_det_pw_wgrad_saved32_workers(N, HW, Cin), Cin / 16);
  const size_t smem = (size_t)(GSTAGES * G32_STAGE + 16 * G32_PS) * 4;
  OCRS_SET_SMEM_ONCE(pw_wgrad_saved32_kernel, smem);
  pw_wgrad_saved32_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(a);
  OCRS_CHECK_LAUNCH("pw_wgrad_saved32_kernel");
  return 0;
}

// 1 when ocrs_det_convt_wgrad_staged can run on these views (16-byte aligned rows: Win % 4 == 0 and Ws % 4 == 0).
int ocrs_det_convt_wgrad_staged_ok(const float* x, long long x_ss, int Hin, int Win, const float* dout, long long dout_ss,
                                   int Hs, int Ws) {
  return ocrs_plane_tma_ok(x, x_ss, Hin, Win) && ocrs_plane_tma_ok(dout, dout_ss, Hs, Ws) && Win >= 16;
}

// Rows of the [workers][Cin][Cout][9] partials of ocrs_det_convt_wgrad_staged.
int ocrs_det_convt_wgrad_staged_workers(int N, int Hin, int Win, int Cin, int Cout) {
  const int pairs = ocrs_cdiv(Cin, 16) * ocrs_cdiv(Cout, 8);
  const long long tiles = (long long)N * ocrs_cdiv(Win, CTC) * ocrs_cdiv(Hin, CTR);
  long long per = (2 * OCRS_NUM_SMS + pairs - 1) / pairs;
  if (per > tiles) per = tiles;
  return (int)(per < 1 ? 1 : per);
}

// ConvTranspose2d(k3, s2) weight gradient, cp.async-staged tiles + mma.sync 3xTF32; same contract as
// ocrs_det_convt_wgrad (partials in weight layout [Cin][Cout][3][3] per worker, fully written).
int ocrs_det_convt_wgrad_staged(const float* x, long long x_ss, int N, int Cin, int Hin, int Win, const float* isc,
                                const float* ish, const float* ilo, const float* dout, long long dout_ss, int Cout,
                                int Hs, int Ws, float* partials, void* stream) {
  OCRS_CHECK_ARG(ocrs_det_convt_wgrad_staged_ok(x, x_ss, Hin, Win, dout, dout_ss, Hs, Ws), "convt_wgrad_staged: unaligned views");
  OCRS_CHECK_ARG(Hs <= 2 * Hin + 1 && Ws <= 2 * Win + 1, "convt_wgrad_staged: crop exceeds the transposed-conv output");
  ConvtWgArgs a;
  a.x = x; a.dout = dout; a.x_ss = x_ss; a.dout_ss = dout_ss; a.isc = isc; a.ish = ish; a.ilo = ilo; a.partials = partials;
  a.Cin = Cin; a.Cout = Cout; a.Hin = Hin; a.Win = Win; a.Hs = Hs; a.Ws = Ws; a.N = N;
  size_t smem = (size_t)2 * CT_STAGE * 4;
  if (smem < 8 * 1152 * 4) smem = 8 * 1152 * 4;
  OCRS_SET_SMEM_ONCE(convt_wgrad_staged_kernel, smem);
  dim3 grid(ocrs_det_convt_wgrad_staged_workers(N, Hin, Win, Cin, Cout), ocrs_cdiv(Cin, 16) * ocrs_cdiv(Cout, 8));
  convt_wgrad_staged_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(a);
  OCRS_CHECK_LAUNCH("convt_wgrad_staged_kernel");
  return 0;
}

}  // extern "C"

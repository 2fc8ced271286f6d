// ConvTranspose2d(k=3, stride=2) forward and data gradient of the detection up path (reference ocrs_models/models.py:76-90)
// for the levels with 16-32 channels, which hold the large images: warp-level tensor-core contractions on tiles staged
// once in shared memory (the levels with >= 64 input channels run as batched tcgen05 GEMMs, det_engine._convt_forward_tc).
//
// A stride-2 3x3 transposed convolution splits into four output-parity classes, each a small dense contraction with no
// overlap-add:   out[co][2qy+py][2qx+px] = bias[co] + sum over the taps (ky, kx) with ky == 1 iff py == 1, kx == 1 iff px == 1
//                                           of sum_ci x[ci][qy - (ky == 2)][qx - (kx == 2)] * w[ci][co][ky][kx]
// (1, 2, 2 and 4 taps). Per warp: M = 16 consecutive input pixels of a row (mma.sync m16n8k8 rows), N = output channels,
// K = input channels; the four shifted A fragments of a k-step are loaded and split into TF32 hi/lo once and feed all
// nine taps; 3xTF32 (hi*hi + hi*lo + lo*hi) keeps fp32-class accuracy. The two column classes of a row end up in the
// same thread, so the stores are 8-byte (ox, ox+1) pairs: 64 contiguous bytes per 8 lanes.
// The data gradient is the mirror image: dx[ci][qy][qx] = sum_{co,ky,kx} dout[co][2qy+ky][2qx+kx] * w[ci][co][ky][kx],
// M = input pixels, N = input channels, K = output channels x 9 taps, with the dout tile de-interleaved by column parity
// while it is staged so that the stride-2 fragment reads are unit-stride (bank-conflict free).
#include "common.cuh"

namespace {

__device__ __forceinline__ void tf32_split2(float v, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32_16n8k8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// row stride of a [k][n] weight matrix in shared memory, = 8 or 24 mod 32: the B-fragment reads (k = lane % 4 (+4),
// n = lane / 4) of a warp then hit 32 different banks
__host__ __device__ constexpr int wstride(int n) { return n % 32 == 8 || n % 32 == 24 ? n : (n / 32) * 32 + (n % 32 < 8 ? 8 : (n % 32 < 24 ? 24 : 40)); }

struct ConvtArgs {
  const float* x; long long x_ss;      // forward: activated-on-load input; backward: dx (output)
  const float *sc, *sh, *lo;
  const float* w;                      // [Cin][Cout][3][3]
  const float* bias;
  float* out; long long out_ss;        // forward: out; backward: dout (input)
  int N, Hin, Win, Hs, Ws, QH, QW, tiles_x, tiles_y;
};

// ---------------------------------------------------------------------------------------------------------------
// forward: tile = FQH x 32 input pixels (one row per warp, two 16-pixel m-tiles), all channels.
constexpr int FQH = 8, FQW = 32, FXC = 36;               // staged columns: qx0 - 4 .. qx0 + 31 (16-byte aligned loads)
constexpr int FPL = ((FQH + 1) * FXC / 32) * 32 + 8;     // plane stride = 8 mod 32 (A fragments: k = lane % 4, m = lane / 4)
static_assert(FPL >= (FQH + 1) * FXC, "plane stride");

template <int CI, int CO>
__global__ void __launch_bounds__(256, (CI == 32 && CO == 32) ? 1 : 2)
convt_fwd_mma_kernel(ConvtArgs a) {
  constexpr int WS = wstride(CO), NT = CO / 8, KS = CI / 8;
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                                          // [CI][FPL]
  uint32_t* wh = reinterpret_cast<uint32_t*>(xs + CI * FPL); // [9][CI][WS] hi
  uint32_t* wl = wh + 9 * CI * WS;                           // lo
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < 9 * CI * CO; i += 256) {
    const int co = i % CO, r = i / CO, ci = r % CI, k = r / CI;
    uint32_t h, l;
    tf32_split2(a.w[((size_t)ci * CO + co) * 9 + k], h, l);
    wh[(k * CI + ci) * WS + co] = h;
    wl[(k * CI + ci) * WS + co] = l;
  }
  float bias[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { bias[nt][0] = a.bias ? a.bias[8 * nt + 2 * t] : 0.f; bias[nt][1] = a.bias ? a.bias[8 * nt + 2 * t + 1] : 0.f; }
  const long long tiles = (long long)a.N * a.tiles_y * a.tiles_x;
  const size_t HWi = (size_t)a.Hin * a.Win, HWs = (size_t)a.Hs * a.Ws;
  // The tile is staged through registers one tile ahead (three CTAs per SM with 80 registers were slower than two with 124): the global loads of tile i+1 are in flight while tile i is
  // contracted (the kernel has too few resident warps to hide that latency otherwise: ncu long-scoreboard stalls).
  constexpr int TOTAL = CI * (FQH + 1) * (FXC / 4), NLD = (TOTAL + 255) / 256;
  float4 pre[NLD];
  unsigned inside = 0;  // bit u: pre[u] is inside the image (gets the BatchNorm+ReLU transform; outside stays 0)
  auto fetch = [&](long long tile) {
    const int tx = (int)(tile % a.tiles_x), ty = (int)((tile / a.tiles_x) % a.tiles_y), n = (int)(tile / ((long long)a.tiles_x * a.tiles_y));
    const float* xn = a.x + (size_t)n * a.x_ss;
    inside = 0;
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int i = tid + 256 * u;
      const int c4 = i % (FXC / 4), r = (i / (FXC / 4)) % (FQH + 1), ci = i / ((FXC / 4) * (FQH + 1));
      const int iy = ty * FQH - 1 + r, ix = tx * FQW - 4 + 4 * c4;
      pre[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < TOTAL && iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win) {  // Win % 4 == 0: the four columns are in or out together
        pre[u] = *reinterpret_cast<const float4*>(xn + (size_t)ci * HWi + (size_t)iy * a.Win + ix);
        inside |= 1u << u;
      }
    }
  };
  if ((long long)blockIdx.x < tiles) fetch(blockIdx.x);
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int tx = (int)(tile % a.tiles_x), ty = (int)((tile / a.tiles_x) % a.tiles_y), n = (int)(tile / ((long long)a.tiles_x * a.tiles_y));
    const int qy0 = ty * FQH, qx0 = tx * FQW;
    __syncthreads();  // the previous tile's fragments have been read (and the weights are in place)
    // rows qy0-1 .. qy0+FQH-1, columns qx0-4 .. qx0+31 of every channel, BatchNorm+ReLU applied, 0 outside the image
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int i = tid + 256 * u;
      if (i >= TOTAL) break;
      const int c4 = i % (FXC / 4), r = (i / (FXC / 4)) % (FQH + 1), ci = i / ((FXC / 4) * (FQH + 1));
      float4 v = pre[u];
      if (a.sc && ((inside >> u) & 1u)) {
        const float s = a.sc[ci], h = a.sh[ci], l = a.lo[ci];
        v.x = xform_apply(v.x, s, h, l); v.y = xform_apply(v.y, s, h, l);
        v.z = xform_apply(v.z, s, h, l); v.w = xform_apply(v.w, s, h, l);
      }
      *reinterpret_cast<float4*>(xs + ci * FPL + r * FXC + 4 * c4) = v;
    }
    __syncthreads();
    if (tile + gridDim.x < tiles) fetch(tile + gridDim.x);
    const int qy = qy0 + warp;
    if (qy >= a.QH) continue;  // (whole warp; the barriers above are reached through the loop head)
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int qxb = qx0 + 16 * half;
      if (qxb >= a.QW) break;
      float acc[4][NT][4];  // class = 2 * py + px
#pragma unroll
      for (int cl = 0; cl < 4; ++cl)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[cl][nt][0] = bias[nt][0]; acc[cl][nt][1] = bias[nt][1]; acc[cl][nt][2] = bias[nt][0]; acc[cl][nt][3] = bias[nt][1]; }
#pragma unroll 1
      for (int ks = 0; ks < KS; ++ks) {
        // A fragments of the four shifted views: v = (ky == 2) + 2 * (kx == 2) -> (row - 1 if ky == 2, col - 1 if kx == 2)
        uint32_t ah[4][4], al[4][4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float* base = xs + (8 * ks + t) * FPL + (warp + 1 - (v & 1)) * FXC + 4 + 16 * half + g - (v >> 1);
#pragma unroll
          for (int q = 0; q < 4; ++q)  // q: (m g, k t), (m g+8, k t), (m g, k t+4), (m g+8, k t+4)
            tf32_split2(base[(q >> 1) * 4 * FPL + (q & 1) * 8], ah[v][q], al[v][q]);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const int ky = k / 3, kx = k % 3;
          const int v = (ky == 2 ? 1 : 0) + (kx == 2 ? 2 : 0), cl = (ky == 1 ? 2 : 0) + (kx == 1 ? 1 : 0);
          const uint32_t* bh = wh + (k * CI + 8 * ks + t) * WS + g;
          const uint32_t* bl = wl + (k * CI + 8 * ks + t) * WS + g;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const uint32_t bh0 = bh[8 * nt], bh1 = bh[4 * WS + 8 * nt], bl0 = bl[8 * nt], bl1 = bl[4 * WS + 8 * nt];
            mma_tf32_16n8k8(acc[cl][nt], al[v], bh0, bh1);
            mma_tf32_16n8k8(acc[cl][nt], ah[v], bl0, bl1);
            mma_tf32_16n8k8(acc[cl][nt], ah[v], bh0, bh1);
          }
        }
      }
      // C fragment: c0 (px g, co 2t), c1 (px g, co 2t+1), c2 (px g+8, co 2t), c3 (px g+8, co 2t+1)
      float* on = a.out + (size_t)n * a.out_ss;
#pragma unroll
      for (int py = 0; py < 2; ++py) {
        const int oy = 2 * qy + py;
        if (oy >= a.Hs) continue;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int co = 8 * nt + 2 * t + (e & 1), ox = 2 * (qxb + g + 8 * (e >> 1));
            float* p = on + (size_t)co * HWs + (size_t)oy * a.Ws + ox;
            if (ox + 1 < a.Ws) *reinterpret_cast<float2*>(p) = make_float2(acc[2 * py][nt][e], acc[2 * py + 1][nt][e]);
            else if (ox < a.Ws) *p = acc[2 * py][nt][e];
          }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// data gradient: tile = BQH x 32 input pixels (warp = (row, 16-pixel half)); the dout tile (2 BQH + 1 rows x 65 columns of
// every output channel) is staged split by column parity: ds[co][row][parity][36].
constexpr int BQH = 4, BQW = 32, BDC = 36;
constexpr int BROWS = 2 * BQH + 1;
constexpr int BPL = ((BROWS * 2 * BDC) / 32) * 32 + 8;   // 648 for BQH = 4 (= 8 mod 32)
static_assert(BPL >= BROWS * 2 * BDC, "plane stride");

template <int CI, int CO>
__global__ void __launch_bounds__(256, CO == 32 ? 1 : 2)
convt_bwd_mma_kernel(ConvtArgs a) {
  constexpr int WS = wstride(CI), NT = CI / 8, KS = CO / 8;
  extern __shared__ __align__(16) float smem[];
  float* ds = smem;                                           // [CO][BPL]
  uint32_t* wh = reinterpret_cast<uint32_t*>(ds + CO * BPL);  // [9][CO][WS] hi   (k = co, n = ci)
  uint32_t* wl = wh + 9 * CO * WS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < 9 * CI * CO; i += 256) {
    const int ci = i % CI, r = i / CI, co = r % CO, k = r / CO;
    uint32_t h, l;
    tf32_split2(a.w[((size_t)ci * CO + co) * 9 + k], h, l);
    wh[(k * CO + co) * WS + ci] = h;
    wl[(k * CO + co) * WS + ci] = l;
  }
  const long long tiles = (long long)a.N * a.tiles_y * a.tiles_x;
  const size_t HWi = (size_t)a.Hin * a.Win, HWs = (size_t)a.Hs * a.Ws;
  float* dx = const_cast<float*>(a.x);
  constexpr int TOTAL = CO * BROWS * 18, NLD = (TOTAL + 255) / 256;
  float4 pre[NLD];  // the next tile, in flight while this one is contracted
  auto fetch = [&](long long tile) {
    const int tx = (int)(tile % a.tiles_x), ty = (int)((tile / a.tiles_x) % a.tiles_y), n = (int)(tile / ((long long)a.tiles_x * a.tiles_y));
    const float* dn = a.out + (size_t)n * a.out_ss;
    // rows 2 qy0 .. 2 qy0 + 2 BQH, columns 2 qx0 .. 2 qx0 + 71 (18 float4; Ws % 4 == 0), 0 outside the crop
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int i = tid + 256 * u;
      const int c4 = i % 18, r = (i / 18) % BROWS, co = i / (18 * BROWS);
      const int oy = 2 * ty * BQH + r, ox = 2 * tx * BQW + 4 * c4;
      pre[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < TOTAL && oy < a.Hs && ox < a.Ws) pre[u] = *reinterpret_cast<const float4*>(dn + (size_t)co * HWs + (size_t)oy * a.Ws + ox);
    }
  };
  if ((long long)blockIdx.x < tiles) fetch(blockIdx.x);
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int tx = (int)(tile % a.tiles_x), ty = (int)((tile / a.tiles_x) % a.tiles_y), n = (int)(tile / ((long long)a.tiles_x * a.tiles_y));
    const int qy0 = ty * BQH, qx0 = tx * BQW;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int i = tid + 256 * u;
      if (i >= TOTAL) break;
      const int c4 = i % 18, r = (i / 18) % BROWS, co = i / (18 * BROWS);
      const float4 v = pre[u];
      float* row = ds + co * BPL + r * 2 * BDC;
      *reinterpret_cast<float2*>(row + 2 * c4) = make_float2(v.x, v.z);        // even columns ox, ox + 2
      *reinterpret_cast<float2*>(row + BDC + 2 * c4) = make_float2(v.y, v.w);  // odd columns
    }
    __syncthreads();
    if (tile + gridDim.x < tiles) fetch(tile + gridDim.x);
    const int qy = qy0 + (warp >> 1), qxb = qx0 + 16 * (warp & 1);
    if (qy >= a.Hin || qxb >= a.Win) continue;
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = 0.f; acc[nt][1] = 0.f; acc[nt][2] = 0.f; acc[nt][3] = 0.f; }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int ky = k / 3, kx = k % 3;
      // dout[co][2 qy + ky][2 (qxb + m) + kx]: parity kx & 1, index 16 (warp & 1) + m + (kx == 2)
      const float* abase = ds + (2 * (warp >> 1) + ky) * 2 * BDC + (kx & 1) * BDC + 16 * (warp & 1) + g + (kx >> 1);
#pragma unroll 1
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)  // (m g, k t), (m g+8, k t), (m g, k t+4), (m g+8, k t+4); k = output channel
          tf32_split2(abase[(8 * ks + t + 4 * (q >> 1)) * BPL + (q & 1) * 8], ah[q], al[q]);
        const uint32_t* bh = wh + (k * CO + 8 * ks + t) * WS + g;
        const uint32_t* bl = wl + (k * CO + 8 * ks + t) * WS + g;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const uint32_t bh0 = bh[8 * nt], bh1 = bh[4 * WS + 8 * nt], bl0 = bl[8 * nt], bl1 = bl[4 * WS + 8 * nt];
          mma_tf32_16n8k8(acc[nt], al, bh0, bh1);
          mma_tf32_16n8k8(acc[nt], ah, bl0, bl1);
          mma_tf32_16n8k8(acc[nt], ah, bh0, bh1);
        }
      }
    }
    float* xo = dx + (size_t)n * a.x_ss + (size_t)qy * a.Win;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ci = 8 * nt + 2 * t + (e & 1), qx = qxb + g + 8 * (e >> 1);
        if (qx < a.Win) xo[(size_t)ci * HWi + qx] = acc[nt][e];
      }
  }
}

template <int CI, int CO>
int launch_fwd(const ConvtArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)(CI * FPL + 2 * 9 * CI * wstride(CO)) * 4;
  OCRS_SET_SMEM_ONCE((convt_fwd_mma_kernel<CI, CO>), smem);
  const long long tiles = (long long)a.N * a.tiles_x * a.tiles_y;
  // the tile is staged synchronously: co-resident CTAs hide each other's load latency, so take every slot the SM offers
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, convt_fwd_mma_kernel<CI, CO>, 256, smem);
  if (per_sm < 1) per_sm = 1;
  const int grid = (int)(tiles < (long long)per_sm * OCRS_NUM_SMS ? tiles : (long long)per_sm * OCRS_NUM_SMS);
  convt_fwd_mma_kernel<CI, CO><<<grid, 256, smem, st>>>(a);
  return 0;
}
template <int CI, int CO>
int launch_bwd(const ConvtArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)(CO * BPL + 2 * 9 * CO * wstride(CI)) * 4;
  OCRS_SET_SMEM_ONCE((convt_bwd_mma_kernel<CI, CO>), smem);
  const long long tiles = (long long)a.N * a.tiles_x * a.tiles_y;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, convt_bwd_mma_kernel<CI, CO>, 256, smem);
  if (per_sm < 1) per_sm = 1;
  const int grid = (int)(tiles < (long long)per_sm * OCRS_NUM_SMS ? tiles : (long long)per_sm * OCRS_NUM_SMS);
  convt_bwd_mma_kernel<CI, CO><<<grid, 256, smem, st>>>(a);
  return 0;
}

bool chan_ok(int Cin, int Cout) { return (Cin == 16 && Cout == 8) || (Cin == 32 && (Cout == 16 || Cout == 32)); }

}  // namespace

// Internal (called by ocrs_det_convt_fwd / ocrs_det_convt_bwd_data): 1 if the tensor-core tile kernels take this case.
bool ocrs_convt_mma_fwd(const float* x, long long x_ss, int N, int Cin, int Hin, int Win, const float* sc, const float* sh,
                        const float* lo, const float* w, const float* bias, int Cout, float* out, long long out_ss, int Hs,
                        int Ws, cudaStream_t st) {
  if (!chan_ok(Cin, Cout) || Win % 4 || x_ss % 4 || ((uintptr_t)x & 15) || Ws % 2 || out_ss % 2 || ((uintptr_t)out & 7)) return false;
  ConvtArgs a{x, x_ss, sc, sh, lo, w, bias, out, out_ss, N, Hin, Win, Hs, Ws, (Hs + 1) / 2, (Ws + 1) / 2, 0, 0};
  a.tiles_x = ocrs_cdiv(a.QW, FQW);
  a.tiles_y = ocrs_cdiv(a.QH, FQH);
  if (Cin == 16) launch_fwd<16, 8>(a, st);
  else if (Cout == 16) launch_fwd<32, 16>(a, st);
  else launch_fwd<32, 32>(a, st);
  return true;
}

bool ocrs_convt_mma_bwd(const float* dout, long long dout_ss, int N, int Cout, int Hs, int Ws, const float* w, int Cin, int Hin,
                        int Win, float* dx, long long dx_ss, cudaStream_t st) {
  if (!chan_ok(Cin, Cout) || Ws % 4 || dout_ss % 4 || ((uintptr_t)dout & 15)) return false;
  ConvtArgs a{dx, dx_ss, nullptr, nullptr, nullptr, w, nullptr, const_cast<float*>(dout), dout_ss, N, Hin, Win, Hs, Ws, Hin, Win, 0, 0};
  a.tiles_x = ocrs_cdiv(Win, BQW);
  a.tiles_y = ocrs_cdiv(Hin, BQH);
  if (Cin == 16) launch_bwd<16, 8>(a, st);
  else if (Cout == 16) launch_bwd<32, 16>(a, st);
  else launch_bwd<32, 32>(a, st);
  return true;
}

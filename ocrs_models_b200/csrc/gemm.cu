// fp32 GEMM used by the recognition path: convolutions (over an im2col matrix), GRU input
// projections, the linear head, and all of their data / weight gradients.
//
//   C[M,N] (+)= op(A)[M,K] * op(B)[K,N] (+ bias[N]) (ReLU)      fp32 in, fp32 accumulate
//
// Operand layouts: A is either A[m][k] ("K-major", lda = row stride) or A[k][m] ("M-major");
// B is either B[n][k] ("K-major") or B[k][n] ("N-major"). 128x128x8 block tile, 256 threads,
// 8x8 register tile per thread, register-prefetch double buffering. Optional per-column
// (sum, sum^2) partials of the stored values for a following BatchNorm, and split-K over
// gridDim.z writing [z][M][N] partial products.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 8, LDS_ = BM + 4;

struct GemmArgs {
  const float* A; long long lda;
  const float* B; long long ldb;
  float* C; long long ldc;
  int M, N, K;
  const float* bias;
  int relu, accumulate;
  float* stats;     // [gridDim.y][2][N] or null
  int k_per_split;  // K range per blockIdx.z (multiple of BK)
};

// Load a 128 x 8 operand tile (rows r0.., k range k0..) into smem as s[k][r].
template <bool KMAJOR, bool VEC>
__device__ __forceinline__ void load_tile(const float* __restrict__ P, long long ld, int R, int kend,
                                          int r0, int k0, float (&reg)[4], int tid) {
  if (KMAJOR) {
    const int r = r0 + (tid >> 1), k = k0 + (tid & 1) * 4;
    if (VEC) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < R && k + 3 < kend) v = *reinterpret_cast<const float4*>(P + (size_t)r * ld + k);
      else if (r < R) {
        const float* p = P + (size_t)r * ld;
        v.x = k < kend ? p[k] : 0.f; v.y = k + 1 < kend ? p[k + 1] : 0.f;
        v.z = k + 2 < kend ? p[k + 2] : 0.f; v.w = k + 3 < kend ? p[k + 3] : 0.f;
      }
      reg[0] = v.x; reg[1] = v.y; reg[2] = v.z; reg[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) reg[i] = (r < R && k + i < kend) ? P[(size_t)r * ld + k + i] : 0.f;
    }
  } else {
    const int k = k0 + (tid >> 5), r = r0 + (tid & 31) * 4;
    if (VEC) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < kend && r + 3 < R) v = *reinterpret_cast<const float4*>(P + (size_t)k * ld + r);
      else if (k < kend) {
        const float* p = P + (size_t)k * ld;
        v.x = r < R ? p[r] : 0.f; v.y = r + 1 < R ? p[r + 1] : 0.f;
        v.z = r + 2 < R ? p[r + 2] : 0.f; v.w = r + 3 < R ? p[r + 3] : 0.f;
      }
      reg[0] = v.x; reg[1] = v.y; reg[2] = v.z; reg[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) reg[i] = (k < kend && r + i < R) ? P[(size_t)k * ld + r + i] : 0.f;
    }
  }
}
template <bool KMAJOR>
__device__ __forceinline__ void store_tile(float* s, const float (&reg)[4], int tid) {
  if (KMAJOR) {
    const int r = tid >> 1, k = (tid & 1) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) s[(k + i) * LDS_ + r] = reg[i];
  } else {
    const int k = tid >> 5, r = (tid & 31) * 4;
    *reinterpret_cast<float4*>(s + k * LDS_ + r) = make_float4(reg[0], reg[1], reg[2], reg[3]);
  }
}

template <bool AK, bool BKM, bool VEC>
__global__ void __launch_bounds__(256, 2) gemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BK * LDS_];
  __shared__ __align__(16) float Bs[2][BK * LDS_];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * g.k_per_split;
  const int kend = min(g.K, kbeg + g.k_per_split);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float ra[4], rb[4];
  load_tile<AK, VEC>(g.A, g.lda, g.M, kend, m0, kbeg, ra, tid);
  load_tile<BKM, VEC>(g.B, g.ldb, g.N, kend, n0, kbeg, rb, tid);
  store_tile<AK>(As[0], ra, tid);
  store_tile<BKM>(Bs[0], rb, tid);
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) {
      load_tile<AK, VEC>(g.A, g.lda, g.M, kend, m0, k0 + BK, ra, tid);
      load_tile<BKM, VEC>(g.B, g.ldb, g.N, kend, n0, k0 + BK, rb, tid);
    }
    const float* a = As[buf];
    const float* b = Bs[buf];
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(a + k * LDS_ + ty * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(a + k * LDS_ + 64 + ty * 4);
      const float4 b0 = *reinterpret_cast<const float4*>(b + k * LDS_ + tx * 4);
      const float4 b1 = *reinterpret_cast<const float4*>(b + k * LDS_ + 64 + tx * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      store_tile<AK>(As[buf ^ 1], ra, tid);
      store_tile<BKM>(Bs[buf ^ 1], rb, tid);
      __syncthreads();
      buf ^= 1;
    }
  }
  // epilogue
  float* C = g.C + (size_t)blockIdx.z * g.M * g.ldc;
  float cs[8], cq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { cs[j] = 0.f; cq[j] = 0.f; }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[n];
      if (g.relu) v = fmaxf(v, 0.f);
      float* dst = C + (size_t)m * g.ldc + n;
      if (g.accumulate) v += *dst;
      *dst = v;
      cs[j] += v;
      cq[j] = fmaf(v, v, cq[j]);
    }
  }
  if (g.stats) {
    __syncthreads();
    float* red = As[0];  // [16][128] sums, then squares in Bs
    float* red2 = Bs[0];
    // As/Bs hold 2*8*132 = 2112 floats each >= 16*128 = 2048
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4;
      red[ty * 128 + c] = cs[j];
      red2[ty * 128 + c] = cq[j];
    }
    __syncthreads();
    if (tid < 128) {
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) { s += red[r * 128 + tid]; q += red2[r * 128 + tid]; }
      const int n = n0 + tid;
      if (n < g.N) {
        g.stats[((size_t)blockIdx.y * 2) * g.N + n] = s;
        g.stats[((size_t)blockIdx.y * 2 + 1) * g.N + n] = q;
      }
    }
  }
}

// col[m][ (ky*kw + kx)*C + c ] = x[n][oy+ky-ph][ox+kx-pw][c]  (NHWC, zero padding).
// grid.x = output pixel (n, oy, ox) groups, one warp-row of threads per (tap, 4 channels): no per-element
// 64-bit index arithmetic, 16-byte loads and stores, writes of a block are one contiguous span of col.
constexpr int I2C_PIX = 4;  // output pixels per block
__global__ void __launch_bounds__(256)
im2col_nhwc_kernel(const float* __restrict__ x, int N, int H, int W, int C, int kh, int kw, int ph,
                   int pw, int Ho, int Wo, float* __restrict__ col) {
  const int C4 = C >> 2, row4 = kh * kw * C4;  // float4 per col row
  const long long npix = (long long)N * Ho * Wo;
  const long long p0 = (long long)blockIdx.x * I2C_PIX;
  __shared__ int spix[I2C_PIX][3];
  if (threadIdx.x < I2C_PIX) {
    const long long p = p0 + threadIdx.x;
    const long long q = p / Wo;
    spix[threadIdx.x][0] = (int)(q / Ho);
    spix[threadIdx.x][1] = (int)(q % Ho);
    spix[threadIdx.x][2] = (int)(p % Wo);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < I2C_PIX * row4; e += 256) {
    const int pl = e / row4, r = e - pl * row4;
    const long long p = p0 + pl;
    if (p >= npix) break;
    const int tap = r / C4, c4 = r - tap * C4;
    const int n = spix[pl][0], oy = spix[pl][1], ox = spix[pl][2];
    const int ky = tap / kw, kx = tap - ky * kw;
    const int iy = oy + ky - ph, ix = ox + kx - pw;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = *reinterpret_cast<const float4*>(x + (((size_t)n * H + iy) * W + ix) * C + c4 * 4);
    reinterpret_cast<float4*>(col)[(size_t)p * row4 + r] = v;
  }
}

// Column sums of A[M][N] (row stride lda): partials [chunks][N], chunk = CS_ROWS rows.
constexpr int CS_ROWS = 64;
__global__ void colsum_kernel(const float* __restrict__ A, long long lda, int M, int N,
                              float* __restrict__ partials) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int m0 = blockIdx.y * CS_ROWS, m1 = min(M, m0 + CS_ROWS);
  float s = 0.f;
#pragma unroll 8
  for (int m = m0; m < m1; ++m) s += A[(size_t)m * lda + n];
  partials[(size_t)blockIdx.y * N + n] = s;
}

// Vectorised variant (lda % 4 == 0, 16-byte aligned): 256 threads = 32 column quads x 8 row lanes; every thread
// has 8 independent 16-byte loads in flight.
__global__ void __launch_bounds__(256)
colsum4_kernel(const float* __restrict__ A, long long lda, int M, int N, float* __restrict__ partials) {
  __shared__ float4 red[8][32];
  const int cq = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int n = (blockIdx.x * 32 + cq) * 4;
  const int m0 = blockIdx.y * CS_ROWS;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < N) {
    float4 v[CS_ROWS / 8];
#pragma unroll
    for (int i = 0; i < CS_ROWS / 8; ++i) {
      const int m = m0 + rl + 8 * i;
      v[i] = m < M ? *reinterpret_cast<const float4*>(A + (size_t)m * lda + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < CS_ROWS / 8; ++i) { acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w; }
  }
  red[rl][cq] = acc;
  __syncthreads();
  if (rl == 0 && n < N) {
    float4 t = red[0][cq];
#pragma unroll
    for (int q = 1; q < 8; ++q) { t.x += red[q][cq].x; t.y += red[q][cq].y; t.z += red[q][cq].z; t.w += red[q][cq].w; }
    float* dst = partials + (size_t)blockIdx.y * N + n;
    dst[0] = t.x;
    if (n + 1 < N) dst[1] = t.y;
    if (n + 2 < N) dst[2] = t.z;
    if (n + 3 < N) dst[3] = t.w;
  }
}

}  // namespace

extern "C" {

int ocrs_gemm_stat_rows(int M) { return ocrs_cdiv(M, BM); }

// a_kmajor: A is [M][K] (else [K][M]); b_kmajor: B is [N][K] (else [K][N]).
// splits > 1: C must hold splits*M*ldc floats ([z][M][ldc]); bias/relu/accumulate/stats unsupported.
int ocrs_gemm(const float* A, long long lda, int a_kmajor, const float* B, long long ldb, int b_kmajor,
              float* C, long long ldc, int M, int N, int K, const float* bias, int relu, int accumulate,
              float* stats, int splits, void* stream) {
  OCRS_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad dims %d %d %d", M, N, K);
  OCRS_CHECK_ARG(splits >= 1, "gemm: bad split count");
  OCRS_CHECK_ARG(splits == 1 || (!bias && !relu && !accumulate && !stats), "gemm: split-K takes no epilogue");
  GemmArgs g{A, lda, B, ldb, C, ldc, M, N, K, bias, relu, accumulate, stats, 0};
  int kps = ocrs_cdiv(K, splits);
  kps = ocrs_cdiv(kps, BK) * BK;
  g.k_per_split = kps;
  const int zs = ocrs_cdiv(K, kps);
  dim3 grid(ocrs_cdiv(N, BN), ocrs_cdiv(M, BM), zs);
  const bool vec = (lda % 4 == 0) && (ldb % 4 == 0) && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(AK, BKM)                                                  \
  do {                                                                   \
    if (vec) gemm_kernel<AK, BKM, true><<<grid, 256, 0, st>>>(g);        \
    else gemm_kernel<AK, BKM, false><<<grid, 256, 0, st>>>(g);           \
  } while (0)
  if (a_kmajor && b_kmajor) LAUNCH(true, true);
  else if (a_kmajor && !b_kmajor) LAUNCH(true, false);
  else if (!a_kmajor && b_kmajor) LAUNCH(false, true);
  else LAUNCH(false, false);
#undef LAUNCH
  OCRS_CHECK_LAUNCH("gemm_kernel");
  return 0;
}

// Number of [M][ldc] partial products ocrs_gemm writes for a requested split count.
int ocrs_gemm_splits(int K, int splits) {
  int kps = ocrs_cdiv(K, splits < 1 ? 1 : splits);
  kps = ocrs_cdiv(kps, BK) * BK;
  return ocrs_cdiv(K, kps);
}

int ocrs_im2col_nhwc(const float* x, int N, int H, int W, int C, int kh, int kw, int ph, int pw, int Ho,
                     int Wo, float* col, void* stream) {
  OCRS_CHECK_ARG(C % 4 == 0, "im2col: channel count %d must be a multiple of 4", C);
  const long long npix = (long long)N * Ho * Wo;
  const long long blocks = (npix + I2C_PIX - 1) / I2C_PIX;
  OCRS_CHECK_ARG(blocks < 2147483647LL, "im2col: too many pixels");
  im2col_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, N, H, W, C, kh, kw, ph, pw, Ho, Wo, col);
  OCRS_CHECK_LAUNCH("im2col_nhwc_kernel");
  return 0;
}

int ocrs_colsum_rows(int M) { return ocrs_cdiv(M, CS_ROWS); }
int ocrs_colsum(const float* A, long long lda, int M, int N, float* partials, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (lda % 4 == 0 && N % 4 == 0 && ((uintptr_t)A % 16 == 0)) {
    dim3 grid(ocrs_cdiv(N, 128), ocrs_cdiv(M, CS_ROWS));
    colsum4_kernel<<<grid, 256, 0, st>>>(A, lda, M, N, partials);
  } else {
    dim3 grid(ocrs_cdiv(N, 128), ocrs_cdiv(M, CS_ROWS));
    colsum_kernel<<<grid, 128, 0, st>>>(A, lda, M, N, partials);
  }
  OCRS_CHECK_LAUNCH("colsum_kernel");
  return 0;
}

}  // extern "C"

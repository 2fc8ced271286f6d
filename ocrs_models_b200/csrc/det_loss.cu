// Balanced binary cross-entropy (reference ocrs_models/train_detection.py:225-263), on device.
//
//   pos = t > 0.5, neg = t < 0.5, L = BCE(p, clamp(t,0,1)) with log clamped at -100,
//   k = min(#pos, #neg), loss = mean(topk(pos*L, k) ++ topk(neg*L, k)).
//
// The reference does two .item() host syncs and two torch.topk calls; here k stays on the
// device and each top-k sum is a 3-pass radix select (11+11+9 bits of the non-negative float
// key) for the k-th largest value T, followed by  sum(L > T) + (k - #(L > T)) * T.
// Per-pixel losses are stored once as a signed map: +L for pos, -L for neg, NaN for neither.
// Ties at the threshold share the remaining gradient weight equally (torch.topk picks an
// unspecified subset of equal values; the loss value is identical).
#include "common.cuh"
#include <math.h>

namespace {

// state words (int32)
enum { ST_NPOS = 0, ST_NNEG = 1, ST_K = 2, ST_NAN = 3, ST_PREFIX = 4, ST_KREM = 6, ST_NEQ = 8, ST_WORDS = 16 };
constexpr int NBINS = 2048;

__device__ __forceinline__ int key_bin(unsigned key, int pass) {
  return pass == 0 ? (int)(key >> 20) : (pass == 1 ? (int)((key >> 9) & 0x7ffu) : (int)(key & 0x1ffu));
}
__device__ __forceinline__ unsigned key_prefix(unsigned key, int pass) {
  return pass == 1 ? (key >> 20) : (key >> 9);
}

__global__ void __launch_bounds__(256)
bce_map_kernel(const float* __restrict__ p, const float* __restrict__ t, long long n,
               float* __restrict__ ls, int* __restrict__ st) {
  __shared__ int cnt[2];
  if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
  __syncthreads();
  int np = 0, nn = 0;
  bool bad = false;  // a NaN prediction or target must surface as a NaN loss (the reference raises / propagates)
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float tv = t[i], pv = p[i];
    bad |= (pv != pv) || (tv != tv);
    const float tc = fminf(fmaxf(tv, 0.f), 1.f);
    const float lp = fmaxf(logf(pv), -100.f), l1 = fmaxf(logf(1.f - pv), -100.f);
    const float L = fabsf(-(tc * lp + (1.f - tc) * l1));
    float v = __int_as_float(0x7fc00000);
    if (tv > 0.5f) { v = L; ++np; }
    else if (tv < 0.5f) { v = -L; ++nn; }
    ls[i] = v;
  }
  np = __reduce_add_sync(0xffffffffu, np);
  nn = __reduce_add_sync(0xffffffffu, nn);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&cnt[0], np); atomicAdd(&cnt[1], nn); }
  __syncthreads();
  if (threadIdx.x == 0) { atomicAdd(&st[ST_NPOS], cnt[0]); atomicAdd(&st[ST_NNEG], cnt[1]); }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(&st[ST_NAN], 1);
}

__global__ void select_init_kernel(int* st, unsigned* hist) {
  for (int i = threadIdx.x; i < 2 * NBINS; i += blockDim.x) hist[i] = 0u;
  if (threadIdx.x == 0) {
    const int k = min(st[ST_NPOS], st[ST_NNEG]);
    st[ST_K] = k;
    st[ST_PREFIX] = 0; st[ST_PREFIX + 1] = 0;
    st[ST_KREM] = k; st[ST_KREM + 1] = k;
    st[ST_NEQ] = 0; st[ST_NEQ + 1] = 0;
  }
}

__global__ void __launch_bounds__(256)
hist_kernel(const float* __restrict__ ls, long long n, int pass, const int* __restrict__ st,
            unsigned* __restrict__ hist) {
  __shared__ unsigned sh[2 * NBINS];
  for (int i = threadIdx.x; i < 2 * NBINS; i += 256) sh[i] = 0u;
  __syncthreads();
  const unsigned pre0 = (unsigned)st[ST_PREFIX], pre1 = (unsigned)st[ST_PREFIX + 1];
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float v = ls[i];
    if (v != v) continue;
    const unsigned bits = __float_as_uint(v), cls = bits >> 31, key = bits & 0x7fffffffu;
    if (pass > 0 && key_prefix(key, pass) != (cls ? pre1 : pre0)) continue;
    atomicAdd(&sh[cls * NBINS + key_bin(key, pass)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * NBINS; i += 256)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// One block: pick, per class, the bin holding the k-th largest key; extend the prefix.
__global__ void select_scan_kernel(int pass, int* st, unsigned* hist) {
  __shared__ unsigned sh[2 * NBINS];
  for (int i = threadIdx.x; i < 2 * NBINS; i += blockDim.x) { sh[i] = hist[i]; hist[i] = 0u; }
  __syncthreads();
  if (threadIdx.x < 2) {
    const int c = threadIdx.x;
    const int nb = pass == 2 ? 512 : NBINS;
    int krem = st[ST_KREM + c];
    if (krem > 0) {
      int b = nb - 1;
      long long above = 0;
      for (; b > 0; --b) {
        if (above + (long long)sh[c * NBINS + b] >= krem) break;
        above += sh[c * NBINS + b];
      }
      st[ST_PREFIX + c] = (int)(((unsigned)st[ST_PREFIX + c] << (pass == 2 ? 9 : 11)) | (unsigned)b);
      st[ST_KREM + c] = krem - (int)above;
      st[ST_NEQ + c] = (int)sh[c * NBINS + b];
    }
  }
}

__global__ void __launch_bounds__(256)
bce_sum_kernel(const float* __restrict__ ls, long long n, const int* __restrict__ st,
               float* __restrict__ partials) {
  __shared__ float red[32];
  const unsigned t0 = (unsigned)st[ST_PREFIX], t1 = (unsigned)st[ST_PREFIX + 1];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float v = ls[i];
    if (v != v) continue;
    const unsigned bits = __float_as_uint(v), cls = bits >> 31, key = bits & 0x7fffffffu;
    if (key > (cls ? t1 : t0)) acc += __uint_as_float(key);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

__global__ void bce_loss_finalize_kernel(const float* __restrict__ partials, int nblk,
                                         const int* __restrict__ st, float* __restrict__ loss) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) s += (double)partials[i];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x != 0) return;
  s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  const int k = st[ST_K];
  for (int c = 0; c < 2; ++c)
    s += (double)st[ST_KREM + c] * (double)__uint_as_float((unsigned)st[ST_PREFIX + c]);
  loss[0] = (k > 0 && st[ST_NAN] == 0) ? (float)(s / (2.0 * k)) : __int_as_float(0x7fc00000);
}

// d loss / d p (aten binary_cross_entropy_backward: (p - t) / max(p (1 - p), 1e-12)).
__global__ void __launch_bounds__(256)
bce_grad_kernel(const float* __restrict__ p, const float* __restrict__ t,
                const float* __restrict__ ls, long long n, const int* __restrict__ st,
                const float* __restrict__ gout, float* __restrict__ dp) {
  const int k = st[ST_K];
  const float g = k > 0 ? gout[0] / (2.f * (float)k) : 0.f;
  const unsigned t0 = (unsigned)st[ST_PREFIX], t1 = (unsigned)st[ST_PREFIX + 1];
  const float tie0 = st[ST_NEQ] > 0 ? (float)st[ST_KREM] / (float)st[ST_NEQ] : 0.f;
  const float tie1 = st[ST_NEQ + 1] > 0 ? (float)st[ST_KREM + 1] / (float)st[ST_NEQ + 1] : 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float v = ls[i];
    float r = 0.f;
    if (v == v && k > 0) {
      const unsigned bits = __float_as_uint(v), cls = bits >> 31, key = bits & 0x7fffffffu;
      const unsigned thr = cls ? t1 : t0;
      const float w = key > thr ? 1.f : (key == thr ? (cls ? tie1 : tie0) : 0.f);
      if (w > 0.f) {
        const float pv = p[i], tc = fminf(fmaxf(t[i], 0.f), 1.f);
        r = g * w * (pv - tc) / fmaxf(pv * (1.f - pv), 1e-12f);
      }
    }
    dp[i] = r;
  }
}

}  // namespace

extern "C" {

int ocrs_bce_state_words(void) { return ST_WORDS + 2 * NBINS; }
int ocrs_bce_blocks(void) { return 8 * OCRS_NUM_SMS; }

// state: int32[ocrs_bce_state_words()] zero-filled by the caller; loss_map: float[n];
// partials: float[ocrs_bce_blocks()]; loss: float[1].
int ocrs_balanced_bce_fwd(const float* pred, const float* target, long long n, float* loss_map,
                          int* state, float* partials, float* loss, void* stream) {
  OCRS_CHECK_ARG(n > 0, "balanced_bce_fwd: empty input");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned* hist = reinterpret_cast<unsigned*>(state + ST_WORDS);
  const int nb = ocrs_bce_blocks();
  bce_map_kernel<<<nb, 256, 0, s>>>(pred, target, n, loss_map, state);
  select_init_kernel<<<1, 256, 0, s>>>(state, hist);
  for (int pass = 0; pass < 3; ++pass) {
    hist_kernel<<<nb, 256, 0, s>>>(loss_map, n, pass, state, hist);
    select_scan_kernel<<<1, 1024, 0, s>>>(pass, state, hist);
  }
  bce_sum_kernel<<<nb, 256, 0, s>>>(loss_map, n, state, partials);
  bce_loss_finalize_kernel<<<1, 256, 0, s>>>(partials, nb, state, loss);
  OCRS_CHECK_LAUNCH_N("balanced_bce_fwd", 10);
  return 0;
}

int ocrs_balanced_bce_bwd(const float* pred, const float* target, const float* loss_map, long long n,
                          const int* state, const float* grad_out, float* grad_pred, void* stream) {
  bce_grad_kernel<<<ocrs_bce_blocks(), 256, 0, (cudaStream_t)stream>>>(pred, target, loss_map, n, state,
                                                                       grad_out, grad_pred);
  OCRS_CHECK_LAUNCH("bce_grad_kernel");
  return 0;
}

}  // extern "C"

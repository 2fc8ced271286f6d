// CTC loss forward (alpha lattice) and backward (beta lattice + softmax-folded gradient).
//
// Replaces torch.nn.CTCLoss() as called at reference ocrs_models/train_rec.py:104,121
// (blank = 0, reduction = "mean", zero_infinity = False; arithmetic = aten ctc_loss).
//
// One warp per sample. The 2S+1 lattice states live in registers, K consecutive states per
// lane, so the s-1 / s-2 neighbours of a step need two warp shuffles; the three-way
// log-sum-exp per state is evaluated in fp32 exactly as aten does (max-shifted exp/log).
// log-prob rows for step t+1 are gathered while step t is being reduced.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int CTC_WARPS = 4;
constexpr int CTC_DEPTH_F = 8;  // frames of log-probs in flight per warp (forward)
constexpr int CTC_DEPTH_B = 6;  // frames of (log-probs, alpha) in flight per warp (backward)

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// The lattice runs in the log2 domain (log-probs are scaled by log2(e) on load) so that every
// log-sum-exp is MUFU.EX2 x3 + MUFU.LG2 with no range reduction, and unreachable states carry the
// finite sentinel NEG instead of -inf so the recurrence needs no branches: the serial chain of a
// step is ~15 dependent instructions per state. ex2/lg2.approx are accurate to ~2^-22, far below
// the fp32 drift of a 200-step lattice.
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f, NEG = -1e30f;
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  return m + lg2(ex2(a - m) + ex2(b - m) + ex2(c - m));
}

template <int K>
__global__ void __launch_bounds__(CTC_WARPS * 32, K <= 3 ? 12 : (K <= 6 ? 8 : 4))
ctc_alpha_kernel(const float* __restrict__ lp, const int* __restrict__ targets, int tgt_stride,
                 const int* __restrict__ in_len, const int* __restrict__ tgt_len,
                 float* __restrict__ alpha, float* __restrict__ nll, int T, int N, int C,
                 int blank) {
  extern __shared__ float ring_all[];
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * CTC_WARPS + (threadIdx.x >> 5);
  if (n >= N) return;
  constexpr int LROW = 32 * K;
  const int S = max(tgt_len[n], 0);
  const int L = 2 * S + 1;
  const int Tn = min(in_len[n], T);
  const int* tg = targets + (size_t)n * tgt_stride;
  if (Tn <= 0) {
    if (lane == 0) nll[n] = (S == 0) ? 0.f : INFINITY;
    return;
  }
  int lab[K];
  bool skip[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int s = lane * K + k;
    int l = blank;
    bool sk = false;
    if (s < L && (s & 1)) {
      l = tg[s >> 1];
      sk = (s >= 3) && (tg[(s >> 1) - 1] != l);
    }
    lab[k] = l;
    skip[k] = sk;
  }
  const size_t tstride = (size_t)N * C;
  // log-prob rows stream through a per-warp shared-memory ring, CTC_DEPTH_F frames ahead (cp.async)
  float* ring = ring_all + (size_t)(threadIdx.x >> 5) * CTC_DEPTH_F * C;
  const float* isrc = lp + (size_t)n * C + lane;  // frame to stage next (this lane's first column)
  auto issue = [&](int t) {
    if (t < Tn) {
      float* dst = ring + (t % CTC_DEPTH_F) * C + lane;
      for (int c = 0; c + lane < C; c += 32) cp_async4(dst + c, isrc + c);
    }
    isrc += tstride;
    cp_async_commit();
  };
#pragma unroll 1
  for (int d = 0; d < CTC_DEPTH_F; ++d) issue(d);
  float a[K];
  float* arow = alpha + (size_t)n * T * LROW + lane * K;
  for (int t = 0; t < Tn; ++t) {
    cp_async_wait<CTC_DEPTH_F - 1>();
    __syncwarp();
    const float* slot = ring + (t % CTC_DEPTH_F) * C;
    float e[K];
#pragma unroll
    for (int k = 0; k < K; ++k) e[k] = slot[lab[k]] * LOG2E;
    __syncwarp();
    issue(t + CTC_DEPTH_F);
    if (t == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int s = lane * K + k;
        a[k] = (s < 2 && s < L) ? e[k] : NEG;
      }
    } else {
      float p1 = __shfl_up_sync(0xffffffffu, a[K - 1], 1);
      float p2 = __shfl_up_sync(0xffffffffu, a[K >= 2 ? K - 2 : 0], K >= 2 ? 1 : 2);
      if (lane == 0) { p1 = NEG; p2 = NEG; }
      if (K == 1 && lane < 2) p2 = NEG;
      float an[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float m1 = (k >= 1) ? a[k >= 1 ? k - 1 : 0] : p1;
        const float m2 = (k >= 2) ? a[k >= 2 ? k - 2 : 0] : ((k == 1 && K >= 2) ? p1 : p2);
        an[k] = (lane * K + k < L) ? lse3(a[k], m1, skip[k] ? m2 : NEG) + e[k] : NEG;
      }
#pragma unroll
      for (int k = 0; k < K; ++k) a[k] = fmaxf(an[k], NEG);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) arow[k] = a[k];
    arow += LROW;
  }
  // nll = -logsumexp(alpha[Tn-1][L-1], alpha[Tn-1][L-2])
  float last = NEG, last2 = NEG;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int s = lane * K + k;
    if (s == L - 1) last = a[k];
    if (s == L - 2) last2 = a[k];
  }
  last = warp_max(last);
  last2 = warp_max(last2);
  if (lane == 0) {
    const float r = lse3(last, last2, NEG);
    nll[n] = (r < -1e29f) ? INFINITY : -r * LN2;
  }
}

// mode: 0 = none (gout[n]), 1 = mean (gout[0] / (N * max(S,1))), 2 = sum (gout[0])
template <int K>
__global__ void __launch_bounds__(CTC_WARPS * 32, K <= 3 ? 8 : (K <= 6 ? 5 : 3))
ctc_beta_grad_kernel(const float* __restrict__ lp, const int* __restrict__ targets,
                     int tgt_stride, const int* __restrict__ in_len,
                     const int* __restrict__ tgt_len, const float* __restrict__ alpha,
                     const float* __restrict__ nll, const float* __restrict__ gout, int mode,
                     int zero_infinity, float* __restrict__ grad, int T, int N, int C,
                     int blank) {
  extern __shared__ float occ_all[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int n = blockIdx.x * CTC_WARPS + wid;
  if (n >= N) return;
  float* occ = occ_all + wid * C;
  constexpr int LROW = 32 * K;
  const int S = max(tgt_len[n], 0);
  const int L = 2 * S + 1;
  const int Tn = min(in_len[n], T);
  const int* tg = targets + (size_t)n * tgt_stride;
  const size_t tstride = (size_t)N * C;

  float gs = (mode == 0) ? gout[n] : gout[0];
  if (mode == 1) gs /= ((float)N * (float)max(S, 1));
  const float nl = nll[n];
  const bool dead = zero_infinity && (nl == INFINITY);
  const float nl2 = nl * LOG2E;

  // frames at or beyond the sample's input length get zero gradient
  for (int t = max(Tn, 0); t < T; ++t) {
    float* g = grad + (size_t)t * tstride + (size_t)n * C;
    for (int c = lane; c < C; c += 32) g[c] = 0.f;
  }
  if (Tn <= 0) return;
  if (dead) {
    for (int t = 0; t < Tn; ++t) {
      float* g = grad + (size_t)t * tstride + (size_t)n * C;
      for (int c = lane; c < C; c += 32) g[c] = 0.f;
    }
    return;
  }

  int lab[K];
  bool skip[K];  // transition s -> s+2 allowed
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int s = lane * K + k;
    int l = blank;
    bool sk = false;
    if (s < L && (s & 1)) {
      l = tg[s >> 1];
      sk = (s + 2 < L) && (tg[(s >> 1) + 1] != l);
    }
    lab[k] = l;
    skip[k] = sk;
  }

  // Per-warp ring of CTC_DEPTH_B frames, each [C log-probs | 32*K alphas], filled by cp.async.
  const int slot_floats = C + LROW;
  float* ring = occ_all + CTC_WARPS * C + (size_t)wid * CTC_DEPTH_B * slot_floats;
  const float* isrc = lp + (size_t)(Tn - 1) * tstride + (size_t)n * C + lane;
  const float* asrc = alpha + ((size_t)n * T + (Tn - 1)) * LROW + lane * K;
  auto issue = [&](int t) {
    if (t >= 0) {
      float* dst = ring + (t % CTC_DEPTH_B) * slot_floats;
      for (int c = 0; c + lane < C; c += 32) cp_async4(dst + lane + c, isrc + c);
#pragma unroll
      for (int k = 0; k < K; ++k) cp_async4(dst + C + lane * K + k, asrc + k);
    }
    isrc -= tstride;
    asrc -= LROW;
    cp_async_commit();
  };
#pragma unroll 1
  for (int d = 0; d < CTC_DEPTH_B; ++d) issue(Tn - 1 - d);
  float b[K];
  float* gptr = grad + (size_t)(Tn - 1) * tstride + (size_t)n * C;
  for (int t = Tn - 1; t >= 0; --t) {
    cp_async_wait<CTC_DEPTH_B - 1>();
    __syncwarp();
    const float* slot = ring + (t % CTC_DEPTH_B) * slot_floats;
    float cur[K], av[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      cur[k] = slot[lab[k]] * LOG2E;
      av[k] = slot[C + lane * K + k];
    }
    // beta(t)[s] = lse(beta(t+1)[s], beta(t+1)[s+1], skip ? beta(t+1)[s+2]) + lp[t][l_s]
    if (t == Tn - 1) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int s = lane * K + k;
        b[k] = (s < L && s >= L - 2) ? cur[k] : NEG;
      }
    } else {
      float n1 = __shfl_down_sync(0xffffffffu, b[0], 1);
      float n2 = __shfl_down_sync(0xffffffffu, b[K >= 2 ? 1 : 0], K >= 2 ? 1 : 2);
      if (lane == 31) { n1 = NEG; n2 = NEG; }
      if (K == 1 && lane >= 30) n2 = NEG;
      float bn[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float u1 = (k + 1 < K) ? b[k + 1 < K ? k + 1 : 0] : n1;
        const float u2 = (k + 2 < K) ? b[k + 2 < K ? k + 2 : 0] : ((k + 1 < K) ? n1 : n2);
        bn[k] = (lane * K + k < L) ? lse3(b[k], u1, skip[k] ? u2 : NEG) + cur[k] : NEG;
      }
#pragma unroll
      for (int k = 0; k < K; ++k) b[k] = fmaxf(bn[k], NEG);
    }
    // this frame's gradient row from alpha(t) + beta(t)
    for (int c = lane; c < C; c += 32) occ[c] = 0.f;
    __syncwarp();
    // Half of the lattice states are blanks: their posterior mass is reduced across the warp with
    // shuffles instead of contending on one shared-memory word; labels use (rarely colliding) atomics.
    float blank_mass = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int s = lane * K + k;
      if (s < L) {
        const float v = av[k] + b[k] + nl2 - cur[k];
        if (v > -1e29f) {
          const float e = ex2(v);
          if (s & 1) atomicAdd(&occ[lab[k]], e);
          else blank_mass += e;
        }
      }
    }
    blank_mass = warp_sum(blank_mass);
    __syncwarp();
    for (int c = lane; c < C; c += 32)
      gptr[c] = (ex2(slot[c] * LOG2E) - occ[c] - (c == blank ? blank_mass : 0.f)) * gs;
    gptr -= tstride;
    __syncwarp();
    issue(t - CTC_DEPTH_B);  // frame t is consumed: refill its slot
  }
}

// loss = reduce(nll) with aten's conventions (mean: mean_n(nll_n / max(S_n, 1))).
__global__ void ctc_reduce_kernel(const float* __restrict__ nll, const int* __restrict__ tgt_len,
                                  float* __restrict__ loss, int N, int mode, int zero_infinity) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float v = nll[n];
    if (zero_infinity && v == INFINITY) v = 0.f;
    if (mode == 1) v /= (float)max(tgt_len[n], 1);
    acc += v;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss[0] = (mode == 1) ? acc / (float)N : acc;
}

int pick_k(int max_S) {
  const int L = 2 * max_S + 1;
  const int ks[] = {1, 2, 3, 4, 5, 6, 8, 12, 16};
  for (int k : ks)
    if (32 * k >= L) return k;
  return 0;
}

}  // namespace

#define CTC_DISPATCH(K_, ...)                \
  switch (K_) {                              \
    case 1: { constexpr int K = 1; __VA_ARGS__; break; }   \
    case 2: { constexpr int K = 2; __VA_ARGS__; break; }   \
    case 3: { constexpr int K = 3; __VA_ARGS__; break; }   \
    case 4: { constexpr int K = 4; __VA_ARGS__; break; }   \
    case 5: { constexpr int K = 5; __VA_ARGS__; break; }   \
    case 6: { constexpr int K = 6; __VA_ARGS__; break; }   \
    case 8: { constexpr int K = 8; __VA_ARGS__; break; }   \
    case 12: { constexpr int K = 12; __VA_ARGS__; break; } \
    case 16: { constexpr int K = 16; __VA_ARGS__; break; } \
    default: break;                          \
  }

extern "C" {

// Row length (floats) of the alpha workspace for targets of at most max_S labels; 0 if unsupported.
int ocrs_ctc_alpha_row(int max_S) { return 32 * pick_k(max_S); }

int ocrs_ctc_fwd(const float* log_probs, const int* targets, int tgt_stride,
                 const int* input_lengths, const int* target_lengths, int T, int N, int C,
                 int max_S, int blank, int reduction, int zero_infinity, float* alpha,
                 float* nll, float* loss, void* stream) {
  OCRS_CHECK_ARG(T > 0 && N > 0 && C > 0, "ctc_fwd: bad dims T=%d N=%d C=%d", T, N, C);
  OCRS_CHECK_ARG(blank >= 0 && blank < C, "ctc_fwd: blank %d out of range", blank);
  const int K_ = pick_k(max_S);
  OCRS_CHECK_ARG(K_ > 0, "ctc_fwd: target length %d exceeds supported maximum 255", max_S);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ocrs_cdiv(N, CTC_WARPS);
  const size_t fsmem = (size_t)CTC_WARPS * CTC_DEPTH_F * C * sizeof(float);
  OCRS_CHECK_ARG(fsmem <= 200 * 1024, "ctc_fwd: class count %d too large", C);
  if (fsmem > 48 * 1024)
    CTC_DISPATCH(K_, (cudaFuncSetAttribute(ctc_alpha_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem)));
  CTC_DISPATCH(K_, (ctc_alpha_kernel<K><<<grid, CTC_WARPS * 32, fsmem, st>>>(
                       log_probs, targets, tgt_stride, input_lengths, target_lengths, alpha, nll,
                       T, N, C, blank)));
  OCRS_CHECK_LAUNCH("ctc_alpha_kernel");
  if (loss && reduction != 0) {
    ctc_reduce_kernel<<<1, 256, 0, st>>>(nll, target_lengths, loss, N, reduction, zero_infinity);
    OCRS_CHECK_LAUNCH("ctc_reduce_kernel");
  }
  return 0;
}

int ocrs_ctc_bwd(const float* log_probs, const int* targets, int tgt_stride,
                 const int* input_lengths, const int* target_lengths, int T, int N, int C,
                 int max_S, int blank, int reduction, int zero_infinity, const float* alpha,
                 const float* nll, const float* grad_out, float* grad_log_probs, void* stream) {
  OCRS_CHECK_ARG(T > 0 && N > 0 && C > 0, "ctc_bwd: bad dims T=%d N=%d C=%d", T, N, C);
  const int K_ = pick_k(max_S);
  OCRS_CHECK_ARG(K_ > 0, "ctc_bwd: target length %d exceeds supported maximum 255", max_S);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ocrs_cdiv(N, CTC_WARPS);
  const size_t smem = (size_t)CTC_WARPS * (C + (size_t)CTC_DEPTH_B * (C + 32 * K_)) * sizeof(float);
  OCRS_CHECK_ARG(smem <= 200 * 1024, "ctc_bwd: class count %d / target length too large for the staging ring", C);
  if (smem > 48 * 1024)
    CTC_DISPATCH(K_, (cudaFuncSetAttribute(ctc_beta_grad_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
  CTC_DISPATCH(K_, (ctc_beta_grad_kernel<K><<<grid, CTC_WARPS * 32, smem, st>>>(
                       log_probs, targets, tgt_stride, input_lengths, target_lengths, alpha, nll,
                       grad_out, reduction, zero_infinity, grad_log_probs, T, N, C, blank)));
  OCRS_CHECK_LAUNCH("ctc_beta_grad_kernel");
  return 0;
}

}  // extern "C"

// CTC loss forward (alpha lattice) and backward (beta lattice + softmax-folded gradient).
//
// Replaces torch.nn.CTCLoss() as called at reference ocrs_models/train_rec.py:104,121
// (blank = 0, reduction = "mean", zero_infinity = False; arithmetic = aten ctc_loss).
//
// One warp per sample. The 2S+1 lattice states are held as (blank, label) PAIRS: lane j owns the NP
// consecutive pairs j*NP .. j*NP+NP-1, i.e. states 2i (blank before label i) and 2i+1 (label i). With this
// layout a blank state needs a two-way and a label state a three-way log-sum-exp (6 MUFU per pair), the
// s-1 / s-2 neighbours of a pair live in the same lane except for ONE warp shuffle per step, and a lane's
// alphas are one contiguous 8*NP-byte vector in HBM. The lattice runs in the log2 domain (log-probs are
// scaled by log2(e) when gathered) with the finite sentinel NEG for unreachable states, so the recurrence
// is branch free; padding states are kept at NEG by folding their validity into the scale/offset of the
// gathered log-prob (one FFMA). The forward kernel gathers only the S+1 log-probs it needs straight from
// global memory, CTC_DF frames ahead in a register ring; the backward kernel streams whole frames and the
// saved alphas CTC_DB frames ahead in registers, stages the frame in shared memory for the label gathers
// and accumulates the per-class posteriors there (labels) / with a warp reduction (blanks).
// Measured on B200 (profiles/): the previous one-state-per-register version issued 193 / 400 instructions
// per warp-step (forward / backward) and was issue bound at 37% of HBM peak for N = 8192.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int CTC_WARPS = 4;
constexpr int CTC_DF = 8;   // frames of gathered log-probs in flight per warp (forward); power of two

// Deep prefetch goes through cp.async groups: a register ring of plain loads does not work, the handful of
// scoreboard slots per warp makes every consumer wait for the most recently issued load as well (measured:
// one full DRAM round trip per step).
template <int BYTES>
__device__ __forceinline__ void cp_async(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f, NEG = -1e30f;
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  return m + lg2(ex2(a - m) + ex2(b - m) + ex2(c - m));
}
__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  return m + lg2(1.f + ex2(-fabsf(a - b)));
}

// Per-pair constants of one lane.
template <int NP>
struct Pairs {
  int lab[NP];                  // class of label i (blank when the pair has no label)
  float sc_b[NP], of_b[NP];     // gathered blank log-prob -> log2 domain, or NEG for padding pairs
  float sc_l[NP], of_l[NP];
  bool vl[NP];                  // pair has a label state
};

template <int NP>
__device__ __forceinline__ void store_pairs(float* dst, const float (&b)[NP], const float (&l)[NP]) {
  if (NP == 1) {
    *reinterpret_cast<float2*>(dst) = make_float2(b[0], l[0]);
  } else {
#pragma unroll
    for (int p = 0; p < NP; p += 2)
      *reinterpret_cast<float4*>(dst + 2 * p) = make_float4(b[p], l[p], b[p + 1 < NP ? p + 1 : p], l[p + 1 < NP ? p + 1 : p]);
  }
}
template <int NP>
__device__ __forceinline__ void load_pairs(const float* src, float (&v)[2 * NP]) {
  if (NP == 1) {
    const float2 q = *reinterpret_cast<const float2*>(src);
    v[0] = q.x; v[1] = q.y;
  } else {
#pragma unroll
    for (int p = 0; p < NP; p += 2) {
      const float4 q = *reinterpret_cast<const float4*>(src + 2 * p);
      v[2 * p] = q.x; v[2 * p + 1] = q.y; v[2 * p + 2 < 2 * NP ? 2 * p + 2 : 0] = q.z; v[2 * p + 3 < 2 * NP ? 2 * p + 3 : 1] = q.w;
    }
  }
}

template <int NP, int NCH>
__global__ void __launch_bounds__(CTC_WARPS * 32, (NP <= 2 && NCH <= 4) ? 8 : (NP == 4 ? 3 : 2))
ctc_alpha_kernel(const float* __restrict__ lp, const int* __restrict__ targets, int tgt_stride,
                 const int* __restrict__ in_len, const int* __restrict__ tgt_len,
                 float* __restrict__ alpha, float* __restrict__ nll, int T, int N, int C,
                 int blank, int row) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * CTC_WARPS + (threadIdx.x >> 5);
  if (n >= N) return;
  // lengths and labels come from the caller's tensors: clamp them into what the shared-memory frame and the lane
  // layout can address (torch raises on such inputs; here they must at least never index out of bounds)
  const int S = min(max(tgt_len[n], 0), min(tgt_stride, 32 * NP - 1));
  const int Tn = min(in_len[n], T);
  const int* tg = targets + (size_t)n * tgt_stride;
  if (Tn <= 0) {
    if (lane == 0) nll[n] = (S == 0) ? 0.f : INFINITY;
    return;
  }
  Pairs<NP> P;
  bool skip[NP];  // transition label i-1 -> label i allowed
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int i = lane * NP + p;
    const bool vl = i < S, vb = i <= S;
    const int l = vl ? min(max(tg[i], 0), C - 1) : blank;
    P.lab[p] = l;
    P.vl[p] = vl;
    skip[p] = vl && i >= 1 && min(max(tg[i - 1], 0), C - 1) != l;
    P.sc_b[p] = vb ? LOG2E : 0.f; P.of_b[p] = vb ? 0.f : NEG;
    P.sc_l[p] = vl ? LOG2E : 0.f; P.of_l[p] = vl ? 0.f : NEG;
  }
  const bool stores = lane * NP <= S;
  const size_t tstride = (size_t)N * C;
  // per-warp shared-memory ring of whole frames (C log-probs), CTC_DF frames ahead: the copy is coalesced
  // (scattered 4-byte cp.async gathers cost ~20 shared-memory wavefronts each, measured), the S+1 values a
  // lane needs are gathered from shared memory
  extern __shared__ __align__(16) float ctc_smem[];
  constexpr int fstride = NCH * 32;
  float* ring = ctc_smem + (size_t)(threadIdx.x >> 5) * CTC_DF * fstride + lane;
  const float* fp = lp + (size_t)n * C + lane;  // frame to copy next (this lane's first column)
  auto issue = [&](int t) {
    if (t < Tn) {
      float* dst = ring + (t & (CTC_DF - 1)) * fstride;
#pragma unroll
      for (int j = 0; j < NCH; ++j)
        if (lane + 32 * j < C) cp_async<4>(dst + 32 * j, fp + 32 * j);
      fp += tstride;
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int d = 0; d < CTC_DF; ++d) issue(d);
  // virtual alpha(-1): all mass on state 0, so that the first recurrence step yields aten's initialisation
  float ab[NP], al[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) { ab[p] = (lane == 0 && p == 0) ? 0.f : NEG; al[p] = NEG; }
  float* arow = alpha + (size_t)n * T * row + lane * 2 * NP;
  for (int t = 0; t < Tn; ++t) {
    cp_async_wait<CTC_DF - 2>();  // the refill below runs one step late: at most DF-2 younger groups may be pending
    __syncwarp();
    // every lane is past its gathers of frame t-1: refill that slot (one step late saves a second warp barrier)
    issue(t > 0 ? t - 1 + CTC_DF : Tn);
    const float* slot = ring - lane + (t & (CTC_DF - 1)) * fstride;
    const float vb = slot[blank];
    float eb[NP], el[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      eb[p] = fmaf(vb, P.sc_b[p], P.of_b[p]);
      el[p] = fmaf(slot[P.lab[p]], P.sc_l[p], P.of_l[p]);
    }
    float prev = __shfl_up_sync(0xffffffffu, al[NP - 1], 1);  // label of the pair before this lane's first
    if (lane == 0) prev = NEG;
    float nb[NP], nl_[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const float pl = p == 0 ? prev : al[p > 0 ? p - 1 : 0];
      nb[p] = fmaxf(lse2(ab[p], pl) + eb[p], NEG);
      nl_[p] = fmaxf(lse3(al[p], ab[p], skip[p] ? pl : NEG) + el[p], NEG);
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) { ab[p] = nb[p]; al[p] = nl_[p]; }
    if (stores) store_pairs<NP>(arow, ab, al);
    arow += row;
  }
  // nll = -logsumexp(alpha[Tn-1][2S], alpha[Tn-1][2S-1])
  float last = NEG, last2 = NEG;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int i = lane * NP + p;
    if (i == S) last = ab[p];
    if (i == S - 1) last2 = al[p];
  }
  last = warp_max(last);
  last2 = warp_max(last2);
  if (lane == 0) {
    const float r = lse2(last, last2);
    nll[n] = (r < -1e29f) ? INFINITY : -r * LN2;
  }
}

// mode: 0 = none (gout[n]), 1 = mean (gout[0] / (N * max(S,1))), 2 = sum (gout[0])
// NCH = 32-class chunks per frame (C <= 32 * NCH); DB = frames in flight.
template <int NP, int NCH, int DB>
__global__ void __launch_bounds__(CTC_WARPS * 32, (NP <= 2 && NCH <= 4) ? 8 : 2)
ctc_beta_grad_kernel(const float* __restrict__ lp, const int* __restrict__ targets,
                     int tgt_stride, const int* __restrict__ in_len,
                     const int* __restrict__ tgt_len, const float* __restrict__ alpha,
                     const float* __restrict__ nll, const float* __restrict__ gout, int mode,
                     int zero_infinity, float* __restrict__ grad, int T, int N, int C,
                     int blank, int row) {
  extern __shared__ __align__(16) float ctc_smem[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int n = blockIdx.x * CTC_WARPS + wid;
  if (n >= N) return;
  constexpr int SLOT = (NCH + 2 * NP) * 32;                 // one frame: [NCH*32 log-probs | 32 lanes x 2NP alphas]
  // per-class posterior mass of the label states, as 2^-30 fixed point: shared-memory float atomicAdd is a CAS
  // loop on sm_100 (measured: ~80 instructions per step), integer ATOMS.ADD is native - and order independent,
  // so the gradient is bitwise deterministic
  int* occ = reinterpret_cast<int*>(ctc_smem + (size_t)wid * (NCH * 32 + DB * SLOT));
  float* ring = reinterpret_cast<float*>(occ + NCH * 32);
  const int S = min(max(tgt_len[n], 0), min(tgt_stride, 32 * NP - 1));
  const int Tn = min(in_len[n], T);
  const int* tg = targets + (size_t)n * tgt_stride;
  const size_t tstride = (size_t)N * C;

  float gs = (mode == 0) ? gout[n] : gout[0];
  if (mode == 1) gs /= ((float)N * (float)max(S, 1));
  const float nl = nll[n];
  const bool dead = zero_infinity && (nl == INFINITY);
  const float nl2 = nl * LOG2E;

  // frames at or beyond the sample's input length get zero gradient (all of them for a dropped sample)
  for (int t = dead ? 0 : max(Tn, 0); t < T; ++t) {
    float* g = grad + (size_t)t * tstride + (size_t)n * C;
    for (int c = lane; c < C; c += 32) g[c] = 0.f;
  }
  if (Tn <= 0 || dead) return;
  if (nl == INFINITY) {  // infeasible alignment without zero_infinity: the gradient is undefined (aten: inf/nan)
    for (int t = 0; t < Tn; ++t) {
      float* g = grad + (size_t)t * tstride + (size_t)n * C;
      for (int c = lane; c < C; c += 32) g[c] = __int_as_float(0x7fc00000);
    }
    return;
  }

  Pairs<NP> P;
  bool skip[NP];  // transition label i -> label i+1 allowed
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int i = lane * NP + p;
    const bool vl = i < S, vb = i <= S;
    const int l = vl ? min(max(tg[i], 0), C - 1) : blank;
    P.lab[p] = l;
    P.vl[p] = vl;
    skip[p] = vl && (i + 1 < S) && min(max(tg[i + 1], 0), C - 1) != l;
    P.sc_b[p] = vb ? LOG2E : 0.f; P.of_b[p] = vb ? 0.f : NEG;
    P.sc_l[p] = vl ? LOG2E : 0.f; P.of_l[p] = vl ? 0.f : NEG;
  }
  const bool has_alpha = lane * NP <= S;
#pragma unroll
  for (int j = 0; j < NCH; ++j) occ[lane + 32 * j] = 0;

  // per-warp shared-memory ring of whole frames of log-probs and saved alphas, DB frames ahead (walking
  // backwards in time), filled by cp.async
  const float* fp = lp + (size_t)(Tn - 1) * tstride + (size_t)n * C + lane;
  const float* ap = alpha + ((size_t)n * T + (Tn - 1)) * row + lane * 2 * NP;
  auto issue = [&](int t) {
    if (t >= 0) {
      float* dst = ring + (t & (DB - 1)) * SLOT;
#pragma unroll
      for (int j = 0; j < NCH; ++j)
        if (lane + 32 * j < C) cp_async<4>(dst + lane + 32 * j, fp + 32 * j);
      if (has_alpha) {
        float* ad = dst + NCH * 32 + lane * 2 * NP;
        if (NP == 1) cp_async<8>(ad, ap);
        else {
#pragma unroll
          for (int k = 0; k < 2 * NP; k += 4) cp_async<16>(ad + k, ap + k);
        }
      }
      fp -= tstride;
      ap -= row;
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int d = 0; d < DB; ++d) issue(Tn - 1 - d);
  // virtual beta(Tn): all mass on the final blank, so that the first step yields aten's initialisation
  float bb[NP], bl[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) { bb[p] = (lane * NP + p == S) ? 0.f : NEG; bl[p] = NEG; }
  float* gptr = grad + (size_t)(Tn - 1) * tstride + (size_t)n * C;
  __syncwarp();
  {
    for (int t = Tn - 1; t >= 0; --t) {
      {
        static_assert(DB >= 2, "deferred refill needs two slots");
        cp_async_wait<DB - 2>();
        __syncwarp();
        issue(t < Tn - 1 ? t + 1 - DB : -1);  // frame t+1 is consumed by every lane: refill its slot
        const float* slot = ring + (t & (DB - 1)) * SLOT;
        float rv[NCH], av[2 * NP];
#pragma unroll
        for (int j = 0; j < NCH; ++j) rv[j] = slot[lane + 32 * j];
        if (has_alpha) load_pairs<NP>(slot + NCH * 32 + lane * 2 * NP, av);
        else {
#pragma unroll
          for (int k = 0; k < 2 * NP; ++k) av[k] = NEG;
        }
        const float vb = slot[blank];
        float cb[NP], cl[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          cb[p] = fmaf(vb, P.sc_b[p], P.of_b[p]);
          cl[p] = fmaf(slot[P.lab[p]], P.sc_l[p], P.of_l[p]);
        }
        // beta(t)[s] = lse(beta(t+1)[s], beta(t+1)[s+1], skip ? beta(t+1)[s+2]) + lp[t][l_s]
        float nxb = __shfl_down_sync(0xffffffffu, bb[0], 1);  // blank / label of the pair after this lane's last
        float nxl = __shfl_down_sync(0xffffffffu, bl[0], 1);
        if (lane == 31) { nxb = NEG; nxl = NEG; }
        float nb[NP], nl_[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const float ub = p + 1 < NP ? bb[p + 1 < NP ? p + 1 : 0] : nxb;
          const float ul = p + 1 < NP ? bl[p + 1 < NP ? p + 1 : 0] : nxl;
          nb[p] = fmaxf(lse2(bb[p], bl[p]) + cb[p], NEG);
          nl_[p] = fmaxf(lse3(bl[p], ub, skip[p] ? ul : NEG) + cl[p], NEG);
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) { bb[p] = nb[p]; bl[p] = nl_[p]; }
        // posterior mass of every state: exp(alpha + beta + nll - lp)
        float blank_mass = 0.f;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          blank_mass += ex2(av[2 * p] + bb[p] + nl2 - cb[p]);
          const float e = ex2(av[2 * p + 1] + bl[p] + nl2 - cl[p]);
          if (P.vl[p]) atomicAdd(&occ[P.lab[p]], __float2int_rn(e * 1073741824.f));
        }
        blank_mass = warp_sum(blank_mass);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const int c = lane + 32 * j;
          if (c < C) {
            const float o = (float)occ[c] * (1.f / 1073741824.f);
            occ[c] = 0;
            gptr[c] = (ex2(rv[j] * LOG2E) - o - (c == blank ? blank_mass : 0.f)) * gs;
          }
        }
        gptr -= tstride;
      }
    }
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

int pick_np(int max_S) {
  const int nps[] = {1, 2, 4, 8};
  for (int np : nps)
    if (32 * np >= max_S + 1) return np;
  return 0;
}
int alpha_row(int max_S) {
  const int np = pick_np(max_S);
  return np ? 2 * np * ocrs_cdiv(max_S + 1, np) : 0;
}

}  // namespace

#define CTC_DISPATCH_NP(NP_, ...)                          \
  switch (NP_) {                                           \
    case 1: { constexpr int NP = 1; __VA_ARGS__; break; }  \
    case 2: { constexpr int NP = 2; __VA_ARGS__; break; }  \
    case 4: { constexpr int NP = 4; __VA_ARGS__; break; }  \
    case 8: { constexpr int NP = 8; __VA_ARGS__; break; }  \
    default: break;                                        \
  }

extern "C" {

// Row length (floats) of the alpha workspace for targets of at most max_S labels; 0 if unsupported.
int ocrs_ctc_alpha_row(int max_S) { return alpha_row(max_S < 0 ? 0 : max_S); }

int ocrs_ctc_fwd(const float* log_probs, const int* targets, int tgt_stride,
                 const int* input_lengths, const int* target_lengths, int T, int N, int C,
                 int max_S, int blank, int reduction, int zero_infinity, float* alpha,
                 float* nll, float* loss, void* stream) {
  OCRS_CHECK_ARG(T > 0 && N > 0 && C > 0, "ctc_fwd: bad dims T=%d N=%d C=%d", T, N, C);
  OCRS_CHECK_ARG(blank >= 0 && blank < C, "ctc_fwd: blank %d out of range", blank);
  if (max_S < 0) max_S = 0;
  const int NP_ = pick_np(max_S);
  OCRS_CHECK_ARG(NP_ > 0, "ctc_fwd: target length %d exceeds supported maximum 255", max_S);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ocrs_cdiv(N, CTC_WARPS);
  const int row = alpha_row(max_S);
  OCRS_CHECK_ARG(C <= 32 * 48, "ctc_fwd: class count %d too large (max 1536)", C);
#define CTC_FWD_LAUNCH(NCH_)                                                                                        \
  do {                                                                                                              \
    const size_t fsmem = (size_t)CTC_WARPS * CTC_DF * NCH_ * 32 * sizeof(float);                                    \
    CTC_DISPATCH_NP(NP_, {                                                                                          \
      if (fsmem > 48 * 1024)                                                                                        \
        cudaFuncSetAttribute(ctc_alpha_kernel<NP, NCH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);  \
      ctc_alpha_kernel<NP, NCH_><<<grid, CTC_WARPS * 32, fsmem, st>>>(                                              \
          log_probs, targets, tgt_stride, input_lengths, target_lengths, alpha, nll, T, N, C, blank, row);          \
    });                                                                                                             \
  } while (0)
  if (C <= 128) CTC_FWD_LAUNCH(4);
  else if (C <= 512) CTC_FWD_LAUNCH(16);
  else CTC_FWD_LAUNCH(48);
#undef CTC_FWD_LAUNCH
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
  if (max_S < 0) max_S = 0;
  const int NP_ = pick_np(max_S);
  OCRS_CHECK_ARG(NP_ > 0, "ctc_bwd: target length %d exceeds supported maximum 255", max_S);
  OCRS_CHECK_ARG(C <= 32 * 48, "ctc_bwd: class count %d too large (max 1536)", C);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ocrs_cdiv(N, CTC_WARPS);
  const int row = alpha_row(max_S);
#define CTC_BWD_LAUNCH(NCH_, DB_)                                                                                  \
  do {                                                                                                             \
    CTC_DISPATCH_NP(NP_, {                                                                                         \
      constexpr int DBV = (NCH_ <= 4) ? ((NP <= 2) ? DB_ : 2) : 2;                                                 \
      const size_t smem = (size_t)CTC_WARPS * (NCH_ * 32 + DBV * (NCH_ + 2 * NP) * 32) * sizeof(float);            \
      if (smem > 48 * 1024)                                                                                        \
        cudaFuncSetAttribute(ctc_beta_grad_kernel<NP, NCH_, DBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      ctc_beta_grad_kernel<NP, NCH_, DBV><<<grid, CTC_WARPS * 32, smem, st>>>(                                      \
          log_probs, targets, tgt_stride, input_lengths, target_lengths, alpha, nll, grad_out, reduction,          \
          zero_infinity, grad_log_probs, T, N, C, blank, row);                                                     \
    });                                                                                                            \
  } while (0)
  if (C <= 128) CTC_BWD_LAUNCH(4, 4);
  else if (C <= 512) CTC_BWD_LAUNCH(16, 2);
  else CTC_BWD_LAUNCH(48, 2);
#undef CTC_BWD_LAUNCH
  OCRS_CHECK_LAUNCH("ctc_beta_grad_kernel");
  return 0;
}

}  // extern "C"

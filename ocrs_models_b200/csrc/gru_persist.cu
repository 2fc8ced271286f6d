// Persistent bidirectional-GRU recurrence (forward and BPTT) on thread-block clusters.
// Replaces the 2 x T dependent steps inside nn.GRU(128, 256, bidirectional, num_layers=2) of
// reference ocrs_models/models.py:245,264-266 (gate order r, z, n; fp32 like the reference).
//
// The recurrence is a chain of [N,256] x [256,768] products, each depending on the previous one.
// Work decomposition: one CLUSTER OF 4 CTAs per (direction, group of 4 batch rows); each CTA keeps its
// quarter of W_hh (64 hidden units x 3 gates x 256, fp32, 192 KB) resident in shared memory for the whole
// sequence. Measured on B200 (profiles/r01), the first version of this kernel was bound by shared-memory
// wavefronts (every thread re-read its k-slice of h for one unit: 4100 wavefronts per step, 57% of the
// step) and by the release fence + cluster barrier of every step. This version:
//   * streams each W row from shared memory exactly once per step, 512 contiguous bytes per warp
//     instruction (lane l owns the float4 chunks l and l + 32 of every 256-long row), with the lane's
//     k-slice of h (forward) / d_gh (backward) held in registers -> 2048 / 2304 wavefronts per step;
//   * does the multiply-adds as packed FFMA2 (fma.rn.f32x2), halving the issue slots of the dot products;
//   * reduces the per-lane partial sums with a transposing butterfly (about one shuffle per value);
//   * exchanges the new state through distributed shared memory with st.async + mbarrier complete_tx
//     (each CTA waits on a LOCAL mbarrier for the 4 KB / 12 KB of the next step; double-buffered, no fence
//     and no cluster barrier inside the loop). OCRS_GRU_SYNC=barrier selects the arrive/wait
//     cluster-barrier exchange instead (A/B and fallback).
// h never round-trips through HBM/L2; only the per-step outputs are written (fire and forget) and the
// per-step inputs are prefetched one step ahead. N = 64 -> 2 x 16 clusters = 128 CTAs, one wave on 148 SMs.
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace cg = cooperative_groups;

namespace {

constexpr int H = 256, G3 = 768;
constexpr int RB = 4;                 // batch rows per cluster
constexpr int UQ = 64;                // hidden units per CTA (cluster of 4 covers 256)
constexpr int FWD_THREADS = 512;      // 16 warps x 4 units
constexpr int BWD_WARPS = 8;          // 8 warps x 8 units (the d_gh slice of a warp is shared by 8 units)
constexpr int BWD_THREADS = BWD_WARPS * 32;
constexpr int BWD_UPW = UQ / BWD_WARPS;
constexpr int FWD_SMEM = (3 * UQ * H + 2 * RB * H) * 4 + 16;
constexpr int BWD_SMEM = (UQ * G3 + 2 * RB * G3) * 4 + 16;

typedef unsigned long long u64;

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }
// d = a * b + c on two packed fp32 lanes (Blackwell FFMA2)
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float sum2(u64 v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}

// Transposing butterfly: every lane holds V partial sums; afterwards lane L holds the 32-lane total of value
// (L >> S) & (V - 1), S = 5 - log2(V) (replicated over the 2^S lanes that share those bits).
template <int V>
__device__ __forceinline__ float warp_reduce_vals(float (&v)[V], int lane) {
  int o = 16;
#pragma unroll
  for (int half = V / 2; half >= 1; half >>= 1, o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send = up ? v[j] : v[j + half];
      const float keep = up ? v[j + half] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
#pragma unroll
  for (; o >= 1; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
  return v[0];
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async(uint32_t raddr, float v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(raddr), "r"(__float_as_uint(v)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arm(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Wait for the st.async payload of the peers. The data arrives through the async proxy and is made visible by the
// mbarrier's transaction completion itself, so the default CTA-scope acquire is enough; `.acquire.cluster` made the
// compiler emit CCTL.IVALL (an L1 invalidate, 10 % of the backward kernel's stall samples) on every step and threw
// away the L1 lines the prefetches below bring in.
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

struct FwdArgs {
  const float* gi[2];    // [T*N][768]
  const float* whh[2];   // [768][256]
  const float* bhh[2];   // [768]
  float* out;            // [T][N][512]
  float* gates;          // [T][N][2][4][256]  (r, z, n, gh_n)
  int T, N, groups;      // groups = ceil(N / RB)
};

template <bool ASYNC>
__global__ void __launch_bounds__(FWD_THREADS, 1) gru_fwd_persist_kernel(FwdArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                      // [3][UQ][H]
  float* hs = smem + 3 * UQ * H;         // [2][RB][H]
  const uint32_t mbar = smem_u32(hs + 2 * RB * H);  // two 8-byte mbarriers
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / 4;        // cluster id
  const int d = cid / p.groups, grp = cid % p.groups;
  const int n0 = grp * RB, u0 = rank * UQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = p.T, N = p.N;

  // resident weight slice: smem row g*UQ + j  <-  W_hh row g*256 + u0 + j
  const float* w = p.whh[d];
  for (int i = tid; i < 3 * UQ * (H / 4); i += FWD_THREADS) {
    const int row = i / (H / 4), c4 = i % (H / 4);
    const int g = row / UQ, j = row % UQ;
    const float4 v = *reinterpret_cast<const float4*>(w + (size_t)(g * H + u0 + j) * H + c4 * 4);
    *reinterpret_cast<float4*>(Ws + row * H + c4 * 4) = v;
  }
  for (int i = tid; i < 2 * RB * H; i += FWD_THREADS) hs[i] = 0.f;
  if (ASYNC && tid == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_arm(mbar + 8, RB * H * 4);  // h(1) lands in buffer 1
    mbar_arm(mbar, RB * H * 4);      // h(2) lands in buffer 0
  }
  // After the reduction lane L holds the pre-activations of unit 4*warp + 2*grp2 + (L >> 4), batch row (L >> 2) & 3
  // for both unit pairs grp2; lanes with (L & 3) == grp2 finish pair grp2.
  const int q = lane & 3;
  const bool act = q < 2;
  const int ul = 4 * warp + 2 * (q & 1) + (lane >> 4);
  const int u = u0 + ul;
  const int bme = (lane >> 2) & 3;
  const int n = n0 + bme;
  const bool live = act && n < N;
  const float br = p.bhh[d][u], bz = p.bhh[d][H + u], bn = p.bhh[d][2 * H + u];
  uint32_t rhs[4], rbar[4];
  float* remote[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    rhs[r] = mapa(smem_u32(hs), r);
    rbar[r] = mapa(mbar, r);
    remote[r] = cluster.map_shared_rank(hs, r);
  }
  cluster.sync();

  const float* gi_base = p.gi[d];
  float gr = 0.f, gz = 0.f, gn = 0.f;
  if (live) {
    const int t0 = d == 0 ? 0 : T - 1;
    const float* gp = gi_base + ((size_t)t0 * N + n) * G3;
    gr = gp[u]; gz = gp[H + u]; gn = gp[2 * H + u];
  }
  const ulonglong2* W2 = reinterpret_cast<const ulonglong2*>(Ws);
  for (int s = 0; s < T; ++s) {
    const int t = d == 0 ? s : T - 1 - s;
    const int buf = s & 1;
    // prefetch next step's input projection
    float ngr = 0.f, ngz = 0.f, ngn = 0.f;
    if (live && s + 1 < T) {
      const int tn = d == 0 ? t + 1 : t - 1;
      const float* gp = gi_base + ((size_t)tn * N + n) * G3;
      ngr = gp[u]; ngz = gp[H + u]; ngn = gp[2 * H + u];
    }
    if (ASYNC && s > 0) {
      mbar_wait_cluster(mbar + 8 * buf, (uint32_t)((s - 1) >> 1) & 1u);
      if (tid == 0) mbar_arm(mbar + 8 * buf, RB * H * 4);  // next use of this buffer: h(s + 2)
    }
    // this lane's k-slice of h(s): float4 chunks lane and lane + 32 of each batch row
    const ulonglong2* h2 = reinterpret_cast<const ulonglong2*>(hs + buf * RB * H);
    ulonglong2 hv[RB][2];
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      hv[b][0] = h2[b * (H / 4) + lane];
      hv[b][1] = h2[b * (H / 4) + 32 + lane];
    }
    float tot[2][3];
#pragma unroll
    for (int g2 = 0; g2 < 2; ++g2) {      // two pairs of units
      u64 acc[3][2][RB];
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int ju = 0; ju < 2; ++ju)
#pragma unroll
          for (int b = 0; b < RB; ++b) acc[g][ju][b] = 0ull;
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int ju = 0; ju < 2; ++ju) {
          const int row = g * UQ + 4 * warp + 2 * g2 + ju;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const ulonglong2 wv = W2[row * (H / 4) + 32 * i + lane];
#pragma unroll
            for (int b = 0; b < RB; ++b) {
              acc[g][ju][b] = fma2(wv.x, hv[b][i].x, acc[g][ju][b]);
              acc[g][ju][b] = fma2(wv.y, hv[b][i].y, acc[g][ju][b]);
            }
          }
        }
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        float v[8];
#pragma unroll
        for (int ju = 0; ju < 2; ++ju)
#pragma unroll
          for (int b = 0; b < RB; ++b) v[ju * 4 + b] = sum2(acc[g][ju][b]);
        tot[g2][g] = warp_reduce_vals<8>(v, lane);
      }
    }
    const float ar = q == 0 ? tot[0][0] : tot[1][0];
    const float az = q == 0 ? tot[0][1] : tot[1][1];
    const float an = q == 0 ? tot[0][2] : tot[1][2];
    const float hp = hs[buf * RB * H + bme * H + u];
    const float r = sigm(gr + ar + br);
    const float z = sigm(gz + az + bz);
    const float ghn = an + bn;
    const float nn = tanhf(gn + r * ghn);
    const float hnew = (1.f - z) * nn + z * hp;
    const int off = (buf ^ 1) * RB * H + bme * H + u;
    if (ASYNC) {
      if (act && s + 1 < T) {
#pragma unroll
        for (int rk = 0; rk < 4; ++rk) st_async(rhs[rk] + 4u * off, hnew, rbar[rk] + 8u * (buf ^ 1));
      }
    } else {
      if (act) {
#pragma unroll
        for (int rk = 0; rk < 4; ++rk) remote[rk][off] = hnew;
      }
      cluster_arrive();
    }
    if (live) {
      p.out[((size_t)t * N + n) * 512 + d * H + u] = hnew;
      float* gs = p.gates + (((size_t)t * N + n) * 2 + d) * 4 * H;
      gs[u] = r; gs[H + u] = z; gs[2 * H + u] = nn; gs[3 * H + u] = ghn;
    }
    gr = ngr; gz = ngz; gn = ngn;
    if (!ASYNC) cluster_wait();
  }
  cluster.sync();  // no CTA leaves while a peer could still address its shared memory
}

struct BwdArgs {
  const float* whhT[2];  // [256][768]  (W_hh transposed)
  const float* dout;     // [T][N][512]
  const float* out;      // [T][N][512]
  const float* gates;    // [T][N][2][4][256]
  float* dgi[2];         // [T*N][768]
  float* dgh[2];         // [T*N][768]
  int T, N, groups;
};

template <bool ASYNC>
__global__ void __launch_bounds__(BWD_THREADS, 1) gru_bwd_persist_kernel(BwdArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* Wt = smem;                  // [UQ][G3]
  float* dg = smem + UQ * G3;        // [2][RB][G3]
  const uint32_t mbar = smem_u32(dg + 2 * RB * G3);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / 4;
  const int d = cid / p.groups, grp = cid % p.groups;
  const int n0 = grp * RB, u0 = rank * UQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = p.T, N = p.N;

  const float* wt = p.whhT[d];
  for (int i = tid; i < UQ * (G3 / 4); i += BWD_THREADS) {
    const int row = i / (G3 / 4), c4 = i % (G3 / 4);
    const float4 v = *reinterpret_cast<const float4*>(wt + (size_t)(u0 + row) * G3 + c4 * 4);
    *reinterpret_cast<float4*>(Wt + row * G3 + c4 * 4) = v;
  }
  for (int i = tid; i < 2 * RB * G3; i += BWD_THREADS) dg[i] = 0.f;
  if (ASYNC && tid == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_arm(mbar + 8, RB * G3 * 4);
    mbar_arm(mbar, RB * G3 * 4);
  }
  // After the reduction of unit group g4 (4 units), lane L holds the total of unit BWD_UPW*warp + 4*g4 + (L >> 3),
  // batch row (L >> 1) & 3; lanes with (L & 1) == g4 finish group g4 (BWD_UPW == 8: every lane is busy).
  static_assert(BWD_UPW == 8 || BWD_UPW == 4, "unit groups of 4");
  constexpr int NG4 = BWD_UPW / 4;
  const int g4me = lane & 1;
  const bool act = g4me < NG4;
  const int ul = BWD_UPW * warp + 4 * (act ? g4me : 0) + (lane >> 3);
  const int u = u0 + ul;
  const int bme = (lane >> 1) & 3;
  const int n = n0 + bme;
  const bool live = act && n < N;
  uint32_t rdg[4], rbar[4];
  float* remote[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    rdg[r] = mapa(smem_u32(dg), r);
    rbar[r] = mapa(mbar, r);
    remote[r] = cluster.map_shared_rank(dg, r);
  }
  cluster.sync();

  float carry = 0.f;  // dh(t_next) * z(t_next) for this lane's (row, unit)
  const ulonglong2* W2 = reinterpret_cast<const ulonglong2*>(Wt);
  for (int s = 0; s < T; ++s) {
    // BPTT visits time in the reverse of this direction's forward order
    const int t = d == 0 ? T - 1 - s : s;
    const int tprev = d == 0 ? t - 1 : t + 1;
    const int buf = s & 1;
    float go = 0.f, r = 0.f, z = 0.f, nn = 0.f, ghn = 0.f, hp = 0.f;
    if (live) {
      go = p.dout[((size_t)t * N + n) * 512 + d * H + u];
      const float* gs = p.gates + (((size_t)t * N + n) * 2 + d) * 4 * H;
      r = gs[u]; z = gs[H + u]; nn = gs[2 * H + u]; ghn = gs[3 * H + u];
      if (tprev >= 0 && tprev < T) hp = p.out[((size_t)tprev * N + n) * 512 + d * H + u];
      // next step's operands: the register file is full (the loads above get sunk below the product), so bring the
      // lines into L1 now and let the loads of the next step hit
      const int t2 = d == 0 ? t - 1 : t + 1;
      if (t2 >= 0 && t2 < T) {
        prefetch_l1(p.dout + ((size_t)t2 * N + n) * 512 + d * H + u);
        const float* gs2 = p.gates + (((size_t)t2 * N + n) * 2 + d) * 4 * H;
        prefetch_l1(gs2 + u); prefetch_l1(gs2 + H + u); prefetch_l1(gs2 + 2 * H + u); prefetch_l1(gs2 + 3 * H + u);
        const int t3 = d == 0 ? t2 - 1 : t2 + 1;
        if (t3 >= 0 && t3 < T) prefetch_l1(p.out + ((size_t)t3 * N + n) * 512 + d * H + u);
      }
    }
    if (ASYNC && s > 0) {
      mbar_wait_cluster(mbar + 8 * buf, (uint32_t)((s - 1) >> 1) & 1u);
      if (tid == 0) mbar_arm(mbar + 8 * buf, RB * G3 * 4);
    }
    u64 acc[BWD_UPW][RB];
#pragma unroll
    for (int ju = 0; ju < BWD_UPW; ++ju)
#pragma unroll
      for (int b = 0; b < RB; ++b) acc[ju][b] = 0ull;
    const ulonglong2* g2 = reinterpret_cast<const ulonglong2*>(dg + buf * RB * G3);
#pragma unroll
    for (int gb = 0; gb < 3; ++gb) {     // gate blocks of 256 columns of W_hh^T
      ulonglong2 gv[RB][2];
#pragma unroll
      for (int b = 0; b < RB; ++b) {
        gv[b][0] = g2[b * (G3 / 4) + gb * 64 + lane];
        gv[b][1] = g2[b * (G3 / 4) + gb * 64 + 32 + lane];
      }
#pragma unroll
      for (int ju = 0; ju < BWD_UPW; ++ju) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const ulonglong2 wv = W2[(BWD_UPW * warp + ju) * (G3 / 4) + gb * 64 + 32 * i + lane];
#pragma unroll
          for (int b = 0; b < RB; ++b) {
            acc[ju][b] = fma2(wv.x, gv[b][i].x, acc[ju][b]);
            acc[ju][b] = fma2(wv.y, gv[b][i].y, acc[ju][b]);
          }
        }
      }
    }
    float tot[NG4];
#pragma unroll
    for (int g4 = 0; g4 < NG4; ++g4) {
      float v[16];
#pragma unroll
      for (int ju = 0; ju < 4; ++ju)
#pragma unroll
        for (int b = 0; b < RB; ++b) v[ju * 4 + b] = sum2(acc[4 * g4 + ju][b]);
      tot[g4] = warp_reduce_vals<16>(v, lane);
    }
    const float a = (NG4 == 2 && g4me == 1) ? tot[NG4 - 1] : tot[0];
    const float dh = go + a + carry;
    const float dn = dh * (1.f - z) * (1.f - nn * nn);
    const float dz = dh * (hp - nn) * z * (1.f - z);
    const float dr = dn * ghn * r * (1.f - r);
    carry = dh * z;
    const int off = (buf ^ 1) * RB * G3 + bme * G3;
    if (ASYNC) {
      if (act && s + 1 < T) {
#pragma unroll
        for (int rk = 0; rk < 4; ++rk) {
          const uint32_t bar = rbar[rk] + 8u * (buf ^ 1), base = rdg[rk] + 4u * (off + u);
          st_async(base, dr, bar);
          st_async(base + 4u * H, dz, bar);
          st_async(base + 8u * H, dn * r, bar);
        }
      }
    } else {
      if (act) {
#pragma unroll
        for (int rk = 0; rk < 4; ++rk) {
          remote[rk][off + u] = dr;
          remote[rk][off + H + u] = dz;
          remote[rk][off + 2 * H + u] = dn * r;
        }
      }
      cluster_arrive();
    }
    if (live) {
      float* gi = p.dgi[d] + ((size_t)t * N + n) * G3;
      float* gh = p.dgh[d] + ((size_t)t * N + n) * G3;
      gi[u] = dr; gi[H + u] = dz; gi[2 * H + u] = dn;
      gh[u] = dr; gh[H + u] = dz; gh[2 * H + u] = dn * r;
    }
    if (!ASYNC) cluster_wait();
  }
  cluster.sync();
}

template <typename Args>
int launch_cluster(void (*kern)(Args), Args a, int clusters, int threads, int smem, cudaStream_t st, const char* name) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * 4);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
  if (e != cudaSuccess) {
    ocrs_set_error("%s: cluster launch failed: %s", name, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// OCRS_GRU_SYNC=barrier: exchange through plain DSMEM stores + one cluster barrier per step.
bool use_async() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OCRS_GRU_SYNC");
    v = (e && strcmp(e, "barrier") == 0) ? 0 : 1;
  }
  return v == 1;
}

}  // namespace

extern "C" {

// All T steps of one bidirectional GRU layer in ONE launch. Same contract as ocrs_gru_layer_fwd.
int ocrs_gru_layer_fwd_persist(const float* gi_f, const float* gi_r, const float* whh_f, const float* whh_r,
                               const float* bhh_f, const float* bhh_r, float* out, float* gates, int T, int N,
                               void* stream) {
  OCRS_CHECK_ARG(T > 0 && N > 0, "gru_layer_fwd_persist: bad dims");
  OCRS_SET_SMEM_ONCE(gru_fwd_persist_kernel<true>, FWD_SMEM);
  OCRS_SET_SMEM_ONCE(gru_fwd_persist_kernel<false>, FWD_SMEM);
  const int groups = ocrs_cdiv(N, RB);
  FwdArgs a{{gi_f, gi_r}, {whh_f, whh_r}, {bhh_f, bhh_r}, out, gates, T, N, groups};
  int rc = launch_cluster(use_async() ? gru_fwd_persist_kernel<true> : gru_fwd_persist_kernel<false>, a, 2 * groups,
                          FWD_THREADS, FWD_SMEM, (cudaStream_t)stream, "gru_fwd_persist");
  if (rc) return rc;
  OCRS_CHECK_LAUNCH("gru_fwd_persist_kernel");
  return 0;
}

// BPTT of one layer in ONE launch. whhT_*: [256][768]. Same outputs as ocrs_gru_layer_bwd (no carry buffer).
int ocrs_gru_layer_bwd_persist(const float* whhT_f, const float* whhT_r, const float* dout, const float* out,
                               const float* gates, float* dgi_f, float* dgi_r, float* dgh_f, float* dgh_r, int T,
                               int N, void* stream) {
  OCRS_CHECK_ARG(T > 0 && N > 0, "gru_layer_bwd_persist: bad dims");
  OCRS_SET_SMEM_ONCE(gru_bwd_persist_kernel<true>, BWD_SMEM);
  OCRS_SET_SMEM_ONCE(gru_bwd_persist_kernel<false>, BWD_SMEM);
  const int groups = ocrs_cdiv(N, RB);
  BwdArgs a{{whhT_f, whhT_r}, dout, out, gates, {dgi_f, dgi_r}, {dgh_f, dgh_r}, T, N, groups};
  int rc = launch_cluster(use_async() ? gru_bwd_persist_kernel<true> : gru_bwd_persist_kernel<false>, a, 2 * groups,
                          BWD_THREADS, BWD_SMEM, (cudaStream_t)stream, "gru_bwd_persist");
  if (rc) return rc;
  OCRS_CHECK_LAUNCH("gru_bwd_persist_kernel");
  return 0;
}

}  // extern "C"

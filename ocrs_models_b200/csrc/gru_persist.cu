// Persistent bidirectional-GRU recurrence (forward and BPTT) on thread-block clusters.
// Replaces the 2 x T dependent steps inside nn.GRU(128, 256, bidirectional, num_layers=2) of
// reference ocrs_models/models.py:245,264-266 (gate order r, z, n; fp32 like the reference).
//
// The recurrence is latency-bound: each step is a [N,256] x [256,768] product that depends on the
// previous one. Work decomposition: one CLUSTER OF 4 CTAs per (direction, group of 4 batch rows).
// Each CTA (512 threads = 16 warps, so shared-memory latency overlaps) keeps its quarter of W_hh (64
// hidden units x 3 gates x 256, fp32, 204 KB) resident in shared memory for the whole sequence; per step it computes its 64 units for the 4 rows, applies
// the gate non-linearities, and broadcasts the 256 new h values to the other three CTAs through
// distributed shared memory; one cluster barrier per step, split into
// arrive / wait around the global stores. (A register-resident-weights variant with clusters of 8 was
// measured slower: the DSMEM broadcast volume doubles.) h never round-trips through HBM/L2;
// only the per-step outputs are written (fire and forget) and gi (the precomputed input
// projection) is prefetched one step ahead. Batch groups are independent, so N = 64 is
// 2 x 16 clusters = 128 CTAs, one wave on 148 SMs.
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace {

constexpr int H = 256, G3 = 768;
constexpr int RB = 4;                 // batch rows per cluster
constexpr int UQ = 64;                // hidden units per CTA (cluster of 4 covers 256)
constexpr int LDW = H + 16;           // fwd W row stride (floats): 68 x 16B == 4 mod 8 -> conflict-free LDS.128
constexpr int LDT = G3 + 16;          // bwd W^T row stride
constexpr int FWD_SMEM = (3 * UQ * LDW + 2 * RB * H) * 4;
constexpr int BWD_SMEM = (UQ * LDT + 2 * RB * G3) * 4;

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// Sum v[0..3] over the 4 lanes {kq & 3}; lane kq returns the total of v[kq & 3] (2 + 1 shuffles).
__device__ __forceinline__ float reduce4_transpose(const float (&v)[4], int kq) {
  float b[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = (kq & 2) ? v[j] : v[j + 2];
    const float keep = (kq & 2) ? v[j + 2] : v[j];
    b[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  const float send = (kq & 1) ? b[0] : b[1];
  const float keep = (kq & 1) ? b[1] : b[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

struct FwdArgs {
  const float* gi[2];    // [T*N][768]
  const float* whh[2];   // [768][256]
  const float* bhh[2];   // [768]
  float* out;            // [T][N][512]
  float* gates;          // [T][N][2][4][256]  (r, z, n, gh_n)
  int T, N, groups;      // groups = ceil(N / RB)
};

__global__ void __launch_bounds__(512, 1) gru_fwd_persist_kernel(FwdArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                      // [3*UQ][LDW]
  float* hs = smem + 3 * UQ * LDW;       // [2][RB][H]
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / 4;        // cluster id
  const int d = cid / p.groups, grp = cid % p.groups;
  const int n0 = grp * RB, u0 = rank * UQ;
  const int tid = threadIdx.x, jl = tid >> 3, kq = tid & 7;
  const int T = p.T, N = p.N;

  // resident weight slice: smem row g*UQ + j  <-  W_hh row g*256 + u0 + j
  const float* w = p.whh[d];
  for (int i = tid; i < 3 * UQ * (H / 4); i += 512) {
    const int row = i / (H / 4), c4 = i % (H / 4);
    const int g = row / UQ, j = row % UQ;
    const float4 v = *reinterpret_cast<const float4*>(w + (size_t)(g * H + u0 + j) * H + c4 * 4);
    *reinterpret_cast<float4*>(Ws + row * LDW + c4 * 4) = v;
  }
  for (int i = tid; i < 2 * RB * H; i += 512) hs[i] = 0.f;
  const int u = u0 + jl;
  const float br = p.bhh[d][u], bz = p.bhh[d][H + u], bn = p.bhh[d][2 * H + u];
  const int n = n0 + (kq & 3);           // the batch row this lane finalises (lanes kq >= 4 mirror kq - 4)
  const bool live = n < N && kq < 4;
  float* remote[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) remote[r] = cluster.map_shared_rank(hs, r);
  cluster.sync();

  const float* gi_base = p.gi[d];
  float gr = 0.f, gz = 0.f, gn = 0.f;
  {
    const int t0 = d == 0 ? 0 : T - 1;
    if (live) {
      const float* gp = gi_base + ((size_t)t0 * N + n) * G3;
      gr = gp[u]; gz = gp[H + u]; gn = gp[2 * H + u];
    }
  }
  const float4* W4 = reinterpret_cast<const float4*>(Ws);
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
    float acc[RB][3];
#pragma unroll
    for (int b = 0; b < RB; ++b) { acc[b][0] = 0.f; acc[b][1] = 0.f; acc[b][2] = 0.f; }
    const float4* h4 = reinterpret_cast<const float4*>(hs + buf * RB * H);
#pragma unroll
    for (int i = 0; i < H / 32; ++i) {
      const int k4 = i * 8 + kq;
      const float4 wr = W4[(0 * UQ + jl) * (LDW / 4) + k4];
      const float4 wz = W4[(1 * UQ + jl) * (LDW / 4) + k4];
      const float4 wn = W4[(2 * UQ + jl) * (LDW / 4) + k4];
#pragma unroll
      for (int b = 0; b < RB; ++b) {
        const float4 hv = h4[b * (H / 4) + k4];
        acc[b][0] = dot4(hv, wr, acc[b][0]);
        acc[b][1] = dot4(hv, wz, acc[b][1]);
        acc[b][2] = dot4(hv, wn, acc[b][2]);
      }
    }
    float tot[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      float v[RB];
#pragma unroll
      for (int b = 0; b < RB; ++b) v[b] = acc[b][g] + __shfl_xor_sync(0xffffffffu, acc[b][g], 4);
      tot[g] = reduce4_transpose(v, kq);
    }
    const float ar = tot[0], az = tot[1], an = tot[2];
    const int row = kq & 3;
    const float hp = hs[buf * RB * H + row * H + u];
    const float r = sigm(gr + ar + br);
    const float z = sigm(gz + az + bz);
    const float ghn = an + bn;
    const float nn = tanhf(gn + r * ghn);
    const float hnew = (1.f - z) * nn + z * hp;
    const int off = (buf ^ 1) * RB * H + row * H + u;
    if (kq < 4) {
#pragma unroll
      for (int rk = 0; rk < 4; ++rk) remote[rk][off] = hnew;
    }
    cluster_arrive();
    if (live) {
      p.out[((size_t)t * N + n) * 512 + d * H + u] = hnew;
      float* gs = p.gates + (((size_t)t * N + n) * 2 + d) * 4 * H;
      gs[u] = r; gs[H + u] = z; gs[2 * H + u] = nn; gs[3 * H + u] = ghn;
    }
    gr = ngr; gz = ngz; gn = ngn;
    cluster_wait();
  }
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

__global__ void __launch_bounds__(512, 1) gru_bwd_persist_kernel(BwdArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* Wt = smem;                  // [UQ][LDT]
  float* dg = smem + UQ * LDT;       // [2][RB][G3]
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / 4;
  const int d = cid / p.groups, grp = cid % p.groups;
  const int n0 = grp * RB, u0 = rank * UQ;
  const int tid = threadIdx.x, kl = tid >> 3, jq = tid & 7;
  const int T = p.T, N = p.N;

  const float* wt = p.whhT[d];
  for (int i = tid; i < UQ * (G3 / 4); i += 512) {
    const int row = i / (G3 / 4), c4 = i % (G3 / 4);
    const float4 v = *reinterpret_cast<const float4*>(wt + (size_t)(u0 + row) * G3 + c4 * 4);
    *reinterpret_cast<float4*>(Wt + row * LDT + c4 * 4) = v;
  }
  for (int i = tid; i < 2 * RB * G3; i += 512) dg[i] = 0.f;
  const int u = u0 + kl;
  const int n = n0 + (jq & 3);
  const bool live = n < N && jq < 4;
  float* remote[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) remote[r] = cluster.map_shared_rank(dg, r);
  cluster.sync();

  float carry = 0.f;  // dh(t_next) * z(t_next) for this thread's (row, unit)
  const float4* W4 = reinterpret_cast<const float4*>(Wt);
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
    }
    float acc[RB];
#pragma unroll
    for (int b = 0; b < RB; ++b) acc[b] = 0.f;
    const float4* g4 = reinterpret_cast<const float4*>(dg + buf * RB * G3);
#pragma unroll 8
    for (int i = 0; i < G3 / 32; ++i) {
      const int j4 = i * 8 + jq;
      const float4 wv = W4[kl * (LDT / 4) + j4];
#pragma unroll
      for (int b = 0; b < RB; ++b) acc[b] = dot4(g4[b * (G3 / 4) + j4], wv, acc[b]);
    }
    float v4[RB];
#pragma unroll
    for (int b = 0; b < RB; ++b) v4[b] = acc[b] + __shfl_xor_sync(0xffffffffu, acc[b], 4);
    const float a = reduce4_transpose(v4, jq);
    const int row = jq & 3;
    const float dh = go + a + carry;
    const float dn = dh * (1.f - z) * (1.f - nn * nn);
    const float dz = dh * (hp - nn) * z * (1.f - z);
    const float dr = dn * ghn * r * (1.f - r);
    carry = dh * z;
    const int off = (buf ^ 1) * RB * G3 + row * G3;
    if (jq < 4) {
#pragma unroll
      for (int rk = 0; rk < 4; ++rk) {
        remote[rk][off + u] = dr;
        remote[rk][off + H + u] = dz;
        remote[rk][off + 2 * H + u] = dn * r;
      }
    }
    cluster_arrive();
    if (live) {
      float* gi = p.dgi[d] + ((size_t)t * N + n) * G3;
      float* gh = p.dgh[d] + ((size_t)t * N + n) * G3;
      gi[u] = dr; gi[H + u] = dz; gi[2 * H + u] = dn;
      gh[u] = dr; gh[H + u] = dz; gh[2 * H + u] = dn * r;
    }
    cluster_wait();
  }
}

template <typename Args>
int launch_cluster(void (*kern)(Args), Args a, int clusters, int smem, cudaStream_t st, const char* name) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * 4);
  cfg.blockDim = dim3(512);
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

}  // namespace

extern "C" {

// All T steps of one bidirectional GRU layer in ONE launch. Same contract as ocrs_gru_layer_fwd.
int ocrs_gru_layer_fwd_persist(const float* gi_f, const float* gi_r, const float* whh_f, const float* whh_r,
                               const float* bhh_f, const float* bhh_r, float* out, float* gates, int T, int N,
                               void* stream) {
  OCRS_CHECK_ARG(T > 0 && N > 0, "gru_layer_fwd_persist: bad dims");
  static bool attr_set = false;
  if (!attr_set) {
    OCRS_CUDA(cudaFuncSetAttribute(gru_fwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    attr_set = true;
  }
  const int groups = ocrs_cdiv(N, RB);
  FwdArgs a{{gi_f, gi_r}, {whh_f, whh_r}, {bhh_f, bhh_r}, out, gates, T, N, groups};
  int rc = launch_cluster(gru_fwd_persist_kernel, a, 2 * groups, FWD_SMEM, (cudaStream_t)stream, "gru_fwd_persist");
  if (rc) return rc;
  OCRS_CHECK_LAUNCH("gru_fwd_persist_kernel");
  return 0;
}

// BPTT of one layer in ONE launch. whhT_*: [256][768]. Same outputs as ocrs_gru_layer_bwd (no carry buffer).
int ocrs_gru_layer_bwd_persist(const float* whhT_f, const float* whhT_r, const float* dout, const float* out,
                               const float* gates, float* dgi_f, float* dgi_r, float* dgh_f, float* dgh_r, int T,
                               int N, void* stream) {
  OCRS_CHECK_ARG(T > 0 && N > 0, "gru_layer_bwd_persist: bad dims");
  static bool attr_set = false;
  if (!attr_set) {
    OCRS_CUDA(cudaFuncSetAttribute(gru_bwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    attr_set = true;
  }
  const int groups = ocrs_cdiv(N, RB);
  BwdArgs a{{whhT_f, whhT_r}, dout, out, gates, {dgi_f, dgi_r}, {dgh_f, dgh_r}, T, N, groups};
  int rc = launch_cluster(gru_bwd_persist_kernel, a, 2 * groups, BWD_SMEM, (cudaStream_t)stream, "gru_bwd_persist");
  if (rc) return rc;
  OCRS_CHECK_LAUNCH("gru_bwd_persist_kernel");
  return 0;
}

}  // extern "C"

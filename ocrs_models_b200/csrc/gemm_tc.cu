// tcgen05 (5th-gen tensor core) GEMM for the recognition path, fp32 in / fp32 out with fp32-class
// accuracy through a 3xTF32 split:   a = a_hi + a_lo  (a_hi = a rounded to TF32),
//   A.B  ~=  A_hi.B_hi + A_hi.B_lo + A_lo.B_hi        (relative error ~2^-21, vs 2^-11 for plain TF32)
// which is what the 1e-3 gradient parity target of this port needs (SURVEY finding 6).
// The tensor core's fp32 accumulation TRUNCATES (measured: relative bias -7e-9 per accumulated
// k-element, scripts/tc_precision.py), so one long accumulation chain would cost ~1e-5 at K ~ 1e3.
// The kernel therefore keeps FOUR accumulators in TMEM: the dominant hi.hi products are spread
// round-robin over three of them (k-block i -> accumulator i % 3) and the two small cross terms go
// to the fourth; the epilogue adds them in round-to-nearest fp32. Error ~5e-7 at K = 1152, on par
// with an fp32 FMA loop.
//
// Persistent CTAs (min(#tiles, 148)) of 448 threads walk 128 x BN output tiles (BN in {32, 64, 128}):
//   warp 0        TMA producer: fp32 operand tiles -> 128B-swizzled shared memory (3 stages); weight operands may
//                 arrive already split (hi and lo tiles both by TMA, ocrs_split_tf32)
//   warp 1        TMEM allocation + single-thread tcgen05.mma issue (kind::tf32, accumulators in TMEM)
//   warps 2-9     converters: split each landed stage into hi / lo tiles in place (fence.proxy.async)
//   warps 10-13   epilogue: tcgen05.ld TMEM -> registers -> (+bias, ReLU, C +=) -> 16-byte global stores straight
//                 from registers (+ per-column sum / sum^2 partials for a following BatchNorm); the accumulators go
//                 back to the MMA warp as soon as they are in registers, so the next tile's main loop overlaps
//                 the stores (and the whole drain when the accumulators are double-buffered, 4*BN <= 256 columns)
// Measured (profiles/): the main loop is bound by the shared-memory port - converter LDS/STS, TMA writes and the
// UMMA operand reads of the 12 MMAs per k-block share 128 B/clk - not by the tensor pipe (56% active on conv.9).
// Operands may be K-major (X[rows][K]) or MN-major (X[K][rows]); the latter is what the weight
// gradients dW = dY^T . X need, and is expressed purely through the TMA boxes and the UMMA
// shared-memory descriptors (no transposes in HBM). Split-K slices are extra tiles of the same walk.
#include "common.cuh"
#include <cuda.h>

namespace {

constexpr int TBM = 128;          // tile rows (UMMA M)
constexpr int TBK = 32;           // fp32 elements per k-block = one 128-byte swizzle row
constexpr int STAGES = 3;
constexpr int A_TILE_BYTES = TBM * TBK * 4;          // 16 KB
constexpr int STAGE_BYTES = 4 * A_TILE_BYTES;        // A_hi, A_lo, B_hi, B_lo (B sized for BN = 128)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 128 /*barriers*/ + 4096 /*BN partials*/;

struct TcArgs {
  float* C; long long ldc;
  int M, N, K;
  const float* bias;
  int relu, accumulate;
  float* stats;          // [gridDim.y][2][N] or null
  int kb_per_split;      // k-blocks per blockIdx.z
  int a_mn, b_mn;        // operand is MN-major (stored [K][rows])
  int bn;                // tile columns (UMMA N)
  uint32_t idesc;
  // Implicit-GEMM 3x3 / pad-1 convolution (forward and dgrad): A is the NHWC activation viewed as
  // [M = N*H*W rows][cC channels]; the k-block (tap, channel chunk) is the SAME 2-D TMA box shifted by
  // (ky-1)*W + (kx-1) rows (im2col folded into the TMA coordinates, never materialised), and the rows
  // whose tap falls outside the image are zeroed by the converter warps while they split hi/lo.
  // conv == 2: weight gradient dW[co][(tap, ci)] = sum_pixels dY[p][co] * x[p + shift(tap)][ci]: A = dY
  // (MN-major, plain), B = the activation (MN-major) where every 32-channel chunk of the N tile belongs
  // to one tap and is loaded with that tap's row shift; out-of-image pixels are zeroed by the converter.
  int conv, cH, cW, cC;
  int tiles_m, tiles_n, zs;  // tile grid walked by the persistent CTAs
  int nbuf;                  // TMEM accumulator sets (2 when 4*bn <= 256 columns)
  int b_split;               // B arrives pre-split (map_b = TF32-exact hi, map_b2 = lo): weights, split once per step
  // Batched GEMM (detection deep levels): tile index z = batch item. Operands are 2-D matrices in which the items are
  // stacked along the ROW axis of the tensor map (planar [N][C][HW] tensors: row = n*C + c), so an item is a row offset;
  // every item runs the full K range, or bsplit slices of it (z = item * bsplit + slice: weight gradients, whose K is
  // the pixel axis); C advances by c_zstride elements per z.
  int batched, a_brows, b_brows, bsplit;
  long long c_zstride;       // elements between consecutive z slices of C (split-K partials: M * ldc)
  float* row_stats;          // [zs * tiles_n][2][M] per-row sum / sum of squares of the tile's columns (BatchNorm over N), or null
  int fast;                  // labelled throughput mode: ONE plain TF32 product per k-step (the tensor core truncates the
                             // fp32 operands to 10 mantissa bits), no hi/lo split, 1/3 of the MMAs; NOT parity numerics
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// UMMA shared-memory descriptor (sm_100): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version 1 at [46,48), SWIZZLE_128B = 2 at [61,64).
// layout_type: 2 = SWIZZLE_128B (16-byte chunks), 1 = SWIZZLE_128B_BASE32B (32-byte chunks; the only
// layout the hardware accepts for MN-major 32-bit operands).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

// Sum over the 32 lanes of v[j], for every j at once: lane j returns column j's total (31 shuffles).
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < o; ++j) {
      const float send = up ? v[j] : v[j + o];
      const float keep = up ? v[j + o] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

constexpr int NUM_THREADS = 448;   // warp 0 TMA, warp 1 MMA, warps 2-9 converters, warps 10-13 epilogue
constexpr int EPI_THREAD0 = 320;

// Persistent kernel: CTA b walks tiles b, b + gridDim.x, ... (n fastest, then m, then split-K slice).
// The stage ring and its mbarrier phases run continuously across tiles, so the TMA producer and the
// converter warps prefetch the next tile while the epilogue warps drain the finished accumulators; with
// 4*bn <= 256 TMEM columns the accumulators are double-buffered and the drain is hidden completely.
template <int CONV, bool FAST>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_b2, TcArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + STAGES * STAGE_BYTES;
  // barriers: full[s] = bar0 + 8 s ; conv[s] = +24 ; empty[s] = +48 ; tmem_full[b] = +72 ; tmem_empty[b] = +88
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto conv_bar = [&](int s) { return bar0 + 24u + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 48u + 8u * s; };
  auto tmem_full_bar = [&](int b) { return bar0 + 72u + 8u * b; };
  auto tmem_empty_bar = [&](int b) { return bar0 + 88u + 8u * b; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * STAGE_BYTES + 112);
  float* sm_stats = reinterpret_cast<float*>(base_ptr + STAGES * STAGE_BYTES + 128);  // [4 chunks][4 warps][2][32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_kb = (g.K + TBK - 1) / TBK;
  const uint32_t b_bytes = (uint32_t)g.bn * TBK * 4;
  const int total_tiles = g.tiles_m * g.tiles_n * g.zs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), 256);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full_bar(b), 1);
      mbar_init(tmem_empty_bar(b), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

#define TILE_DECODE(tile_)                                         \
  const int tn_ = (tile_) % g.tiles_n, tr_ = (tile_) / g.tiles_n;  \
  const int tm_ = tr_ % g.tiles_m, tzz_ = tr_ / g.tiles_m;         \
  /* batched: z = (item, K split); tz_ = item (row offsets), tzz_ = output slice */ \
  const int tz_ = g.batched ? tzz_ / g.bsplit : tzz_;              \
  const int m0 = tm_ * TBM, n0 = tn_ * g.bn;                       \
  const int kb0 = (g.batched ? tzz_ % g.bsplit : tzz_) * g.kb_per_split; \
  const int nkb = min(total_kb, kb0 + g.kb_per_split) - kb0;       \
  (void)m0; (void)n0; (void)kb0; (void)tz_; (void)tm_; (void)tzz_;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        TILE_DECODE(tile)
        uint32_t tx_bytes = A_TILE_BYTES + (g.b_split ? 2u : 1u) * b_bytes;
        if (CONV == 2) {
          int nvalid = 0;
          for (int c = 0; c < g.bn / 32; ++c) nvalid += (n0 + 32 * c < 9 * g.cC) ? 1 : 0;
          tx_bytes = A_TILE_BYTES + (uint32_t)nvalid * 4096u;
        }
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t sa = base + s * STAGE_BYTES, sb = sa + 2 * A_TILE_BYTES;
          mbar_expect_tx(full_bar(s), tx_bytes);
          const int k0 = (kb0 + i) * TBK;
          if (CONV == 1) {
            const int cpb = g.cC / TBK, kb = kb0 + i;
            const int tap = kb / cpb, c0 = (kb - tap * cpb) * TBK;
            tma_load_2d(sa, &map_a, c0, m0 + (tap / 3 - 1) * g.cW + (tap % 3 - 1), full_bar(s));
          } else if (!g.a_mn) tma_load_2d(sa, &map_a, k0, m0 + tz_ * g.a_brows, full_bar(s));
          else
            for (int c = 0; c < TBM / 32; ++c) tma_load_2d(sa + c * 4096, &map_a, m0 + 32 * c, k0 + tz_ * g.a_brows, full_bar(s));
          if (CONV == 2) {
            for (int c = 0; c < g.bn / 32; ++c) {
              const int col = n0 + 32 * c;
              if (col < 9 * g.cC) {
                const int tap = col / g.cC, ci0 = col - tap * g.cC;
                tma_load_2d(sb + c * 4096, &map_b, ci0, k0 + (tap / 3 - 1) * g.cW + (tap % 3 - 1), full_bar(s));
              }
            }
          } else if (!g.b_mn) {
            tma_load_2d(sb, &map_b, k0, n0 + tz_ * g.b_brows, full_bar(s));
            if (g.b_split) tma_load_2d(sb + A_TILE_BYTES, &map_b2, k0, n0, full_bar(s));
          } else {
            for (int c = 0; c < g.bn / 32; ++c) tma_load_2d(sb + c * 4096, &map_b, n0 + 32 * c, k0 + tz_ * g.b_brows, full_bar(s));
            if (g.b_split)
              for (int c = 0; c < g.bn / 32; ++c)
                tma_load_2d(sb + A_TILE_BYTES + c * 4096, &map_b2, n0 + 32 * c, k0, full_bar(s));
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // K-major (SWIZZLE_128B): 8 rows x 128 B atoms 1024 B apart (SBO), k-step = +32 B inside the row.
      // MN-major (SWIZZLE_128B_BASE32B): 32-wide MN chunks 4096 B apart (LBO), atoms of 4 k-rows
      // 512 B apart (SBO), k-step (8 k-rows) = +1024 B.
      const uint32_t a_lbo = g.a_mn ? 4096u : 16u, b_lbo = g.b_mn ? 4096u : 16u;
      const uint32_t a_sbo = g.a_mn ? 512u : 1024u, b_sbo = g.b_mn ? 512u : 1024u;
      const uint32_t a_lt = g.a_mn ? 1u : 2u, b_lt = g.b_mn ? 1u : 2u;
      const uint32_t a_step = g.a_mn ? 1024u : 32u, b_step = g.b_mn ? 1024u : 32u;
      uint32_t it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        TILE_DECODE(tile)
        const uint32_t buf = lt % (uint32_t)g.nbuf, use = lt / (uint32_t)g.nbuf;
        mbar_wait(tmem_empty_bar(buf), (use & 1u) ^ 1u);  // epilogue has drained this accumulator set
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tb = tmem_base + buf * 256u;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(conv_bar(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = base + s * STAGE_BYTES, sb = sa + 2 * A_TILE_BYTES;
#pragma unroll
          for (int ks = 0; ks < TBK / 8; ++ks) {
            const uint64_t ah = make_desc(sa + ks * a_step, a_lbo, a_sbo, a_lt);
            const uint64_t al = make_desc(sa + A_TILE_BYTES + ks * a_step, a_lbo, a_sbo, a_lt);
            const uint64_t bh = make_desc(sb + ks * b_step, b_lbo, b_sbo, b_lt);
            const uint64_t bl = make_desc(sb + A_TILE_BYTES + ks * b_step, b_lbo, b_sbo, b_lt);
            // accumulator columns: [0,bn) [bn,2bn) [2bn,3bn) = hi.hi round-robin, [3bn,4bn) = cross terms
            umma_tf32(tb + (uint32_t)((i % 3) * g.bn), ah, bh, g.idesc, (i >= 3 || ks > 0) ? 1u : 0u);
            if (!FAST) {
              umma_tf32(tb + (uint32_t)(3 * g.bn), ah, bl, g.idesc, (i > 0 || ks > 0) ? 1u : 0u);
              umma_tf32(tb + (uint32_t)(3 * g.bn), al, bh, g.idesc, 1u);
            }
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(tmem_full_bar(buf));
      }
    }
  } else if (threadIdx.x < EPI_THREAD0) {
    // ---- converters: split landed fp32 tiles into TF32-exact hi (in place) and lo ----
    // 8 warps (2 per scheduler). Each thread owns 4 float4 of A and 4 of B per stage; all eight are
    // loaded before any is stored so the shared-memory latency overlaps.
    const int ct = threadIdx.x - 64;  // 0..255
    const int b_vec = g.b_split ? 0 : (int)b_bytes / 16;  // pre-split B tiles need no conversion
    auto split4 = [](const float4& v, float4& h, float4& l) {
      h.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xffffe000u); l.x = v.x - h.x;
      h.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xffffe000u); l.y = v.y - h.y;
      h.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xffffe000u); l.z = v.z - h.z;
      h.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xffffe000u); l.w = v.w - h.w;
    };
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      TILE_DECODE(tile)
      // CONV == 1: this thread converts A tile rows (ct >> 3) + 32 q; 9-bit tap validity per row
      uint32_t rmask[4] = {0x1ffu, 0x1ffu, 0x1ffu, 0x1ffu};
      if (CONV == 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int m = m0 + (ct >> 3) + 32 * q;
          const int ox = m % g.cW, oy = (m / g.cW) % g.cH;
          uint32_t bits = 0;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int iy = oy + tap / 3 - 1, ix = ox + tap % 3 - 1;
            if (iy >= 0 && iy < g.cH && ix >= 0 && ix < g.cW) bits |= 1u << tap;
          }
          rmask[q] = bits;
        }
      }
      // CONV == 2: this thread converts B k-row (ct >> 3) of each of the four 32-channel chunks
      int wdy[4] = {0, 0, 0, 0}, wdx[4] = {0, 0, 0, 0};
      bool wchunk[4] = {false, false, false, false};
      int woy = 0, wox = 0;
      if (CONV == 2) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col = n0 + 32 * c;
          wchunk[c] = c < g.bn / 32 && col < 9 * g.cC;
          const int tap = wchunk[c] ? col / g.cC : 0;
          wdy[c] = tap / 3 - 1;
          wdx[c] = tap % 3 - 1;
        }
        const long long m = (long long)kb0 * TBK + (ct >> 3);
        wox = (int)(m % g.cW);
        woy = (int)((m / g.cW) % g.cH);
      }
      int tap_a = 0, chunk_a = 0;  // CONV == 1: (tap, channel chunk) of the current k-block
      const int cpb = CONV == 1 ? g.cC / TBK : 1;
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        float4* ah = reinterpret_cast<float4*>(base_ptr + s * STAGE_BYTES);
        float4* al = ah + A_TILE_BYTES / 16;
        float4* bh = al + A_TILE_BYTES / 16;
        float4* bl = bh + A_TILE_BYTES / 16;
        float4 va[4], vb[4];
        bool ka[4], kb_[4];  // keep (true) or zero (false)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = ct + 256 * q;
          ka[q] = CONV == 1 ? ((rmask[q] >> tap_a) & 1u) != 0 : true;
          kb_[q] = j < b_vec;
          if (CONV == 2) {
            const int iy = woy + wdy[q], ix = wox + wdx[q];
            kb_[q] = kb_[q] && wchunk[q] && iy >= 0 && iy < g.cH && ix >= 0 && ix < g.cW;
          }
          if (FAST) continue;
          va[q] = ah[j];
          vb[q] = (j < b_vec) ? bh[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (FAST) {
          // fast mode: the operands stay as they landed; only out-of-image rows of the implicit convolutions are zeroed
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = ct + 256 * q;
            if (!ka[q]) ah[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (CONV == 2 && j < b_vec && !kb_[q]) bh[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        } else
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = ct + 256 * q;
          float4 h, l;
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          split4(ka[q] ? va[q] : z, h, l);
          ah[j] = h;
          al[j] = l;
          if (j < b_vec) {
            split4(kb_[q] ? vb[q] : z, h, l);
            bh[j] = h;
            bl[j] = l;
          }
        }
        if (CONV == 1) {
          if (++chunk_a == cpb) { chunk_a = 0; ++tap_a; }
        }
        if (CONV == 2) {  // advance this thread's pixel by one k-block (32 pixels)
          wox += TBK;
          while (wox >= g.cW) { wox -= g.cW; if (++woy == g.cH) woy = 0; }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(conv_bar(s));
      }
    }
  } else {
    // ---- epilogue warps: TMEM -> registers -> (+bias, ReLU, accumulate) -> global, straight from registers;
    // the accumulators are released to the MMA warp as soon as the last column chunk has been read ----
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int et = threadIdx.x - EPI_THREAD0;
    const bool vec_ok = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15u) == 0);
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      TILE_DECODE(tile)
      const uint32_t buf = lt % (uint32_t)g.nbuf, use = lt / (uint32_t)g.nbuf;
      mbar_wait(tmem_full_bar(buf), use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int m = m0 + q * 32 + lane;   // output row held by this thread
      const bool row_ok = m < g.M;
      float* crow = g.C + (size_t)tzz_ * g.c_zstride + (size_t)(row_ok ? m : 0) * g.ldc;
      float rs1 = 0.f, rs2 = 0.f;  // row statistics of this tile (g.row_stats)
      const int n_main = min(3, nkb);
      const uint32_t lane_addr = tmem_base + buf * 256u + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < g.bn; c0 += 32) {
        uint32_t r0[32], r1[32];
        float v[32];
        if (!FAST) tmem_ld32(lane_addr + (uint32_t)(3 * g.bn + c0), r0);  // cross terms (smallest magnitude first)
        tmem_ld32(lane_addr + (uint32_t)c0, r1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (FAST ? 0.f : __uint_as_float(r0[j])) + __uint_as_float(r1[j]);
        if (n_main > 1) {
          tmem_ld32(lane_addr + (uint32_t)(g.bn + c0), r0);
          if (n_main > 2) tmem_ld32(lane_addr + (uint32_t)(2 * g.bn + c0), r1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r0[j]);
          if (n_main > 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r1[j]);
          }
        }
        if (c0 + 32 >= g.bn) {  // every column of this accumulator set is in registers: hand it back
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar(buf));
        }
        const int nb = n0 + c0;
        if (g.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < g.N) v[j] += __ldg(g.bias + nb + j);
        }
        if (g.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (g.row_stats) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < g.N) { rs1 += v[j]; rs2 = fmaf(v[j], v[j], rs2); }
        }
        if (row_ok) {
          if (vec_ok) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const int n = nb + 4 * j4;
              if (n + 3 < g.N) {
                float4* dst = reinterpret_cast<float4*>(crow + n);
                if (g.accumulate) {
                  const float4 o = *dst;
                  v[4 * j4] += o.x; v[4 * j4 + 1] += o.y; v[4 * j4 + 2] += o.z; v[4 * j4 + 3] += o.w;
                }
                *dst = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (n + e < g.N) {
                    if (g.accumulate) v[4 * j4 + e] += crow[n + e];
                    crow[n + e] = v[4 * j4 + e];
                  }
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < g.N) {
                if (g.accumulate) v[j] += crow[nb + j];
                crow[nb + j] = v[j];
              }
          }
        }
        if (g.stats) {  // per-column sum / sum of squares over this warp's 32 rows
          float sq[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = row_ok ? v[j] : 0.f;
            sq[j] = v[j] * v[j];
          }
          const float s1 = warp_transpose_sum(v, lane);
          const float s2 = warp_transpose_sum(sq, lane);
          float* dst = sm_stats + ((c0 >> 5) * 4 + q) * 64;
          dst[lane] = s1;
          dst[32 + lane] = s2;
        }
      }
      if (g.row_stats && row_ok) {
        const size_t blk = (size_t)tz_ * g.tiles_n + tn_;
        g.row_stats[(blk * 2) * g.M + m] = rs1;
        g.row_stats[(blk * 2 + 1) * g.M + m] = rs2;
      }
      if (g.stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et < g.bn && n0 + et < g.N) {
          const float* src = sm_stats + ((et >> 5) * 4) * 64 + (et & 31);
          const float s1 = (src[0] + src[64]) + (src[128] + src[192]);
          const float s2 = (src[32] + src[96]) + (src[160] + src[224]);
          g.stats[((size_t)tm_ * 2) * g.N + n0 + et] = s1;
          g.stats[((size_t)tm_ * 2 + 1) * g.N + n0 + et] = s2;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
  }
#undef TILE_DECODE
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 tensor map: `inner` contiguous elements, `outer` rows of stride ld elements.
int make_map(CUtensorMap* map, const float* ptr, long long inner, long long outer, long long ld, int box_inner,
             int box_outer, bool mn_major) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { ocrs_set_error("gemm_tc: cuTensorMapEncodeTiled unavailable"); return -1; }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ocrs_set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
  return 0;
}

// hi = v rounded to TF32 (the converter warps' rounding), lo = v - hi (exact).
__global__ void split_tf32_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = src[i];
  const float h = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
  hi[i] = h;
  lo[i] = v - h;
}

int g_fast_mode = 0;  // process-wide numerics mode of the tensor-core GEMMs (ocrs_gemm_tc_set_fast)

template <int CONV>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap* mb2, TcArgs g, int tiles_n, int tiles_m, int zs,
              cudaStream_t st, const char* what) {
  OCRS_SET_SMEM_ONCE((gemm_tc_kernel<CONV, false>), SMEM_BYTES);
  OCRS_SET_SMEM_ONCE((gemm_tc_kernel<CONV, true>), SMEM_BYTES);
  g.tiles_m = tiles_m;
  g.tiles_n = tiles_n;
  g.zs = zs;
  g.nbuf = (4 * g.bn <= 256) ? 2 : 1;
  const long long total = (long long)tiles_m * tiles_n * zs;
  const int ctas = (int)(total < OCRS_NUM_SMS ? total : OCRS_NUM_SMS);
  g.b_split = mb2 != nullptr;
  g.fast = g_fast_mode;
  if (!g.batched) g.c_zstride = (long long)g.M * g.ldc;
  if (g.fast) gemm_tc_kernel<CONV, true><<<ctas, NUM_THREADS, SMEM_BYTES, st>>>(ma, mb, mb2 ? *mb2 : mb, g);
  else gemm_tc_kernel<CONV, false><<<ctas, NUM_THREADS, SMEM_BYTES, st>>>(ma, mb, mb2 ? *mb2 : mb, g);
  OCRS_CHECK_LAUNCH(what);
  return 0;
}

}  // namespace

extern "C" {

// Numerics mode of every tensor-core GEMM / implicit convolution launched afterwards by this process:
// 0 (default) = parity mode, 3xTF32 with four TMEM accumulators (fp32-class accuracy, what every parity test runs);
// 1 = labelled fast mode, one plain TF32 product (operands truncated to 10 mantissa bits by the tensor core, relative
// error ~1e-3 per product). Pre-split weight operands must not be used in fast mode. Returns the previous mode.
int ocrs_gemm_tc_set_fast(int fast) {
  const int prev = g_fast_mode;
  g_fast_mode = fast ? 1 : 0;
  return prev;
}

// 1 when (lda, ldb, pointers) satisfy the TMA constraints of ocrs_gemm_tc.
int ocrs_gemm_tc_supported(const float* A, long long lda, const float* B, long long ldb) {
  return (lda % 4 == 0) && (ldb % 4 == 0) && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0);
}

static int gemm_tc_impl(const float* A, long long lda, int a_kmajor, const float* B, const float* B_lo, long long ldb,
                        int b_kmajor, float* C, long long ldc, int M, int N, int K, const float* bias, int relu,
                        int accumulate, float* stats, int splits, void* stream) {
  OCRS_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_tc: bad dims %d %d %d", M, N, K);
  OCRS_CHECK_ARG(ocrs_gemm_tc_supported(A, lda, B, ldb), "gemm_tc: operands must be 16-byte aligned with ld %% 4 == 0");
  OCRS_CHECK_ARG(!B_lo || ((uintptr_t)B_lo % 16 == 0), "gemm_tc: B_lo must be 16-byte aligned");
  OCRS_CHECK_ARG(splits >= 1, "gemm_tc: bad split count");
  OCRS_CHECK_ARG(splits == 1 || (!bias && !relu && !accumulate && !stats), "gemm_tc: split-K takes no epilogue");
  const int bn = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
  CUtensorMap ma, mb, mb2;
  if (a_kmajor) { if (make_map(&ma, A, K, M, lda, TBK, TBM, false)) return -1; }
  else          { if (make_map(&ma, A, M, K, lda, 32, TBK, true)) return -1; }
  if (b_kmajor) { if (make_map(&mb, B, K, N, ldb, TBK, bn, false)) return -1; }
  else          { if (make_map(&mb, B, N, K, ldb, 32, TBK, true)) return -1; }
  if (B_lo) {
    if (b_kmajor) { if (make_map(&mb2, B_lo, K, N, ldb, TBK, bn, false)) return -1; }
    else          { if (make_map(&mb2, B_lo, N, K, ldb, 32, TBK, true)) return -1; }
  }
  TcArgs g{C, ldc, M, N, K, bias, relu, accumulate, stats, 0, !a_kmajor, !b_kmajor, bn, 0, 0, 0, 0, 0};
  const int total_kb = ocrs_cdiv(K, TBK);
  g.kb_per_split = ocrs_cdiv(total_kb, splits);
  const int zs = ocrs_cdiv(total_kb, g.kb_per_split);
  // instruction descriptor: D = F32 (bit 4), A/B = TF32 (2 << 7, 2 << 10), majors, N >> 3 at 17, M >> 4 at 24
  g.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)g.a_mn << 15) | ((uint32_t)g.b_mn << 16) |
            ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
  return launch_tc<0>(ma, mb, B_lo ? &mb2 : nullptr, g, ocrs_cdiv(N, bn), ocrs_cdiv(M, TBM), zs, (cudaStream_t)stream, "gemm_tc_kernel");
}

// Same contract as ocrs_gemm (gemm.cu): a_kmajor: A is [M][K] else [K][M]; b_kmajor: B is [N][K] else [K][N].
int ocrs_gemm_tc(const float* A, long long lda, int a_kmajor, const float* B, long long ldb, int b_kmajor,
                 float* C, long long ldc, int M, int N, int K, const float* bias, int relu, int accumulate,
                 float* stats, int splits, void* stream) {
  return gemm_tc_impl(A, lda, a_kmajor, B, nullptr, ldb, b_kmajor, C, ldc, M, N, K, bias, relu, accumulate, stats, splits, stream);
}

// ocrs_gemm_tc with the B operand (a weight matrix) already split by ocrs_split_tf32 into B_hi / B_lo of the same
// layout: both halves arrive by TMA and the converter warps only touch A (14% less shared-memory traffic per k-block).
int ocrs_gemm_tc_presplit(const float* A, long long lda, int a_kmajor, const float* B_hi, const float* B_lo, long long ldb,
                          int b_kmajor, float* C, long long ldc, int M, int N, int K, const float* bias, int relu,
                          int accumulate, float* stats, int splits, void* stream) {
  OCRS_CHECK_ARG(B_lo != nullptr, "gemm_tc_presplit: B_lo is null");
  return gemm_tc_impl(A, lda, a_kmajor, B_hi, B_lo, ldb, b_kmajor, C, ldc, M, N, K, bias, relu, accumulate, stats, splits, stream);
}

static int gemm_tc_batched_impl(const float* A, long long lda, int a_kmajor, int a_rows, int a_brows, const float* B,
                                long long ldb, int b_kmajor, int b_rows, int b_brows, float* C, long long ldc,
                                long long c_batch_stride, int M, int N, int K, int batch, float* row_stats, int k_splits,
                                void* stream) {
  OCRS_CHECK_ARG(M > 0 && N > 0 && K > 0 && batch > 0 && k_splits >= 1, "gemm_tc_batched: bad dims");
  OCRS_CHECK_ARG(k_splits == 1 || row_stats == nullptr, "gemm_tc_batched: row statistics need the whole K range per item");
  OCRS_CHECK_ARG(ocrs_gemm_tc_supported(A, lda, B, ldb), "gemm_tc_batched: operands must be 16-byte aligned with ld %% 4 == 0");
  OCRS_CHECK_ARG(K % TBK == 0 || (a_brows == 0 && b_brows == 0) || (a_kmajor && b_kmajor),
                 "gemm_tc_batched: K must be a multiple of 32 when items are stacked along K rows");
  const int bn = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
  CUtensorMap ma, mb;
  if (a_kmajor) { if (make_map(&ma, A, K, a_rows, lda, TBK, TBM, false)) return -1; }
  else          { if (make_map(&ma, A, M, a_rows, lda, 32, TBK, true)) return -1; }
  if (b_kmajor) { if (make_map(&mb, B, K, b_rows, ldb, TBK, bn, false)) return -1; }
  else          { if (make_map(&mb, B, N, b_rows, ldb, 32, TBK, true)) return -1; }
  TcArgs g{C, ldc, M, N, K, nullptr, 0, 0, nullptr, 0, !a_kmajor, !b_kmajor, bn, 0, 0, 0, 0, 0};
  g.kb_per_split = ocrs_cdiv(ocrs_cdiv(K, TBK), k_splits);
  g.bsplit = ocrs_cdiv(ocrs_cdiv(K, TBK), g.kb_per_split);
  g.batched = 1;
  g.a_brows = a_brows;
  g.b_brows = b_brows;
  g.c_zstride = c_batch_stride;
  g.row_stats = row_stats;
  g.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)g.a_mn << 15) | ((uint32_t)g.b_mn << 16) |
            ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
  return launch_tc<0>(ma, mb, nullptr, g, ocrs_cdiv(N, bn), ocrs_cdiv(M, TBM), batch * g.bsplit, (cudaStream_t)stream, "gemm_tc_kernel(batched)");
}

// Batched C[z] = op(A[z]) op(B[z]) on the same tcgen05 kernel, z < batch, for operands whose items are stacked along
// the row axis of a 2-D matrix (planar NCHW activations: row = n*C + c, row pitch H*W). a_brows / b_brows = rows per
// item (0: the operand is shared by all items, e.g. a weight matrix); a_rows / b_rows = total rows of the 2-D matrix.
// a_kmajor: A rows are M (each row K long), else rows are K (each row M long); same for B with N. C[z] starts at
// C + z * c_batch_stride, row pitch ldc. row_stats (optional): [batch * ceil(N / bn)][2][M] per-row sums and sums of
// squares (BatchNorm statistics over N when the rows are channels). Used by the detection levels with >= 64 channels.
int ocrs_gemm_tc_batched(const float* A, long long lda, int a_kmajor, int a_rows, int a_brows, const float* B,
                         long long ldb, int b_kmajor, int b_rows, int b_brows, float* C, long long ldc,
                         long long c_batch_stride, int M, int N, int K, int batch, float* row_stats, void* stream) {
  return gemm_tc_batched_impl(A, lda, a_kmajor, a_rows, a_brows, B, ldb, b_kmajor, b_rows, b_brows, C, ldc, c_batch_stride, M, N,
                              K, batch, row_stats, 1, stream);
}

// K slices ocrs_gemm_tc_batched_splitk really uses when asked for `want` (whole 32-wide k-blocks per slice).
int ocrs_gemm_tc_batched_splits(int K, int want) {
  const int kb = ocrs_cdiv(K, TBK), per = ocrs_cdiv(kb, want < 1 ? 1 : want);
  return ocrs_cdiv(kb, per);
}

// ocrs_gemm_tc_batched with every item's K range cut into ocrs_gemm_tc_batched_splits(K, k_splits) slices that run as
// separate tiles: C holds batch * splits partial results, slice (item z, split s) at C + (z * splits + s) * c_batch_stride.
// For the weight gradients of the detection levels, where K is the pixel axis of one sample and M x N is one tile.
int ocrs_gemm_tc_batched_splitk(const float* A, long long lda, int a_kmajor, int a_rows, int a_brows, const float* B,
                                long long ldb, int b_kmajor, int b_rows, int b_brows, float* C, long long ldc,
                                long long c_batch_stride, int M, int N, int K, int batch, int k_splits, void* stream) {
  OCRS_CHECK_ARG(a_kmajor && b_kmajor, "gemm_tc_batched_splitk: both operands must be K-major");
  return gemm_tc_batched_impl(A, lda, a_kmajor, a_rows, a_brows, B, ldb, b_kmajor, b_rows, b_brows, C, ldc, c_batch_stride, M, N,
                              K, batch, nullptr, k_splits, stream);
}

// Rows of the row_stats partials of ocrs_gemm_tc_batched.
int ocrs_gemm_tc_batched_stat_rows(int N, int batch) { return batch * ocrs_cdiv(N, N <= 32 ? 32 : (N <= 64 ? 64 : 128)); }

// hi[i] = src[i] rounded to TF32, lo[i] = src[i] - hi[i]: the operand split of the 3xTF32 GEMM, done once per
// step for weight matrices.
int ocrs_split_tf32(const float* src, float* hi, float* lo, long long n, void* stream) {
  OCRS_CHECK_ARG(n >= 0, "split_tf32: bad length");
  if (n == 0) return 0;
  split_tf32_kernel<<<ocrs_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(src, hi, lo, n);
  OCRS_CHECK_LAUNCH("split_tf32_kernel");
  return 0;
}

static int conv3x3_tc_impl(const float* x, int N, int H, int W, int Cin, const float* wp, const float* wp_lo, int Cout,
                           float* out, long long ldc, const float* bias, int relu, float* stats, void* stream) {
  OCRS_CHECK_ARG(Cin % TBK == 0 && Cin > 0, "conv3x3_tc: Cin %d must be a multiple of 32", Cin);
  OCRS_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)wp % 16 == 0) && ((uintptr_t)wp_lo % 16 == 0),
                 "conv3x3_tc: operands must be 16-byte aligned");
  const long long Ml = (long long)N * H * W;
  OCRS_CHECK_ARG(Ml < 2147483647LL - 128, "conv3x3_tc: too many pixels");
  const int M = (int)Ml, K = 9 * Cin;
  const int bn = Cout <= 32 ? 32 : (Cout <= 64 ? 64 : 128);
  CUtensorMap ma, mb, mb2;
  if (make_map(&ma, x, Cin, M, Cin, TBK, TBM, false)) return -1;
  if (make_map(&mb, wp, K, Cout, K, TBK, bn, false)) return -1;
  if (wp_lo && make_map(&mb2, wp_lo, K, Cout, K, TBK, bn, false)) return -1;
  TcArgs g{out, ldc, M, Cout, K, bias, relu, 0, stats, K / TBK, 0, 0, bn, 0, 1, H, W, Cin};
  g.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
  return launch_tc<1>(ma, mb, wp_lo ? &mb2 : nullptr, g, ocrs_cdiv(Cout, bn), ocrs_cdiv(M, TBM), 1, (cudaStream_t)stream,
                      "gemm_tc_kernel(conv3x3)");
}

// 3x3 / pad-1 / stride-1 convolution as an implicit GEMM on the tensor cores (no im2col buffer):
// x: NHWC [N][H][W][Cin] (Cin % 32 == 0), wp: [Cout][(ky, kx, ci)], out: [N*H*W][ldc].
// Forward convolution and, with the flipped/transposed weights, the data gradient.
int ocrs_conv3x3_tc(const float* x, int N, int H, int W, int Cin, const float* wp, int Cout, float* out,
                    long long ldc, const float* bias, int relu, float* stats, void* stream) {
  return conv3x3_tc_impl(x, N, H, W, Cin, wp, nullptr, Cout, out, ldc, bias, relu, stats, stream);
}

// ocrs_conv3x3_tc with the packed weights already split by ocrs_split_tf32 (wp_hi, wp_lo).
int ocrs_conv3x3_tc_presplit(const float* x, int N, int H, int W, int Cin, const float* wp_hi, const float* wp_lo, int Cout,
                             float* out, long long ldc, const float* bias, int relu, float* stats, void* stream) {
  OCRS_CHECK_ARG(wp_lo != nullptr, "conv3x3_tc_presplit: wp_lo is null");
  return conv3x3_tc_impl(x, N, H, W, Cin, wp_hi, wp_lo, Cout, out, ldc, bias, relu, stats, stream);
}

// Weight gradient of the same convolution, also without an im2col buffer:
// dwp[z][Cout][(ky, kx, ci)] (split-K partials, z < ocrs_gemm_tc_splits(N*H*W, splits)) from
// dy: [N*H*W][Cout] and x: NHWC [N][H][W][Cin]. Cin % 32 == 0, Cout % 4 == 0.
int ocrs_conv3x3_wgrad_tc(const float* dy, const float* x, int N, int H, int W, int Cin, int Cout,
                          float* dwp, int splits, void* stream) {
  OCRS_CHECK_ARG(Cin % TBK == 0 && Cout % 4 == 0, "conv3x3_wgrad_tc: unsupported channel counts %d -> %d", Cin, Cout);
  OCRS_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)dy % 16 == 0), "conv3x3_wgrad_tc: operands must be 16-byte aligned");
  OCRS_CHECK_ARG(splits >= 1, "conv3x3_wgrad_tc: bad split count");
  const long long Kl = (long long)N * H * W;
  OCRS_CHECK_ARG(Kl < 2147483647LL - 128, "conv3x3_wgrad_tc: too many pixels");
  // N tile: 96 columns when that needs no more tiles than 128 (Cin = 32: 288 = 3 x 96 instead of 2.25 x 128: no MMA columns or
  // B traffic wasted on padding; measured 510 -> 453 us). With more tiles (Cin = 64: 6 x 96 vs 4.5 x 128) the extra pass over
  // the dY operand costs more than the padding (224 -> 289 us).
  const int K = (int)Kl, Nn = 9 * Cin, bn = (ocrs_cdiv(Nn, 96) == ocrs_cdiv(Nn, 128)) ? 96 : 128;
  CUtensorMap ma, mb;
  if (make_map(&ma, dy, Cout, K, Cout, 32, TBK, true)) return -1;
  if (make_map(&mb, x, Cin, K, Cin, 32, TBK, true)) return -1;
  TcArgs g{dwp, Nn, Cout, Nn, K, nullptr, 0, 0, nullptr, 0, 1, 1, bn, 0, 2, H, W, Cin};
  const int total_kb = ocrs_cdiv(K, TBK);
  g.kb_per_split = ocrs_cdiv(total_kb, splits);
  const int zs = ocrs_cdiv(total_kb, g.kb_per_split);
  g.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(bn >> 3) << 17) |
            ((uint32_t)(TBM >> 4) << 24);
  return launch_tc<2>(ma, mb, nullptr, g, ocrs_cdiv(Nn, bn), ocrs_cdiv(Cout, TBM), zs, (cudaStream_t)stream,
                      "gemm_tc_kernel(conv3x3 wgrad)");
}

// Number of [M][ldc] partial products ocrs_gemm_tc writes for a requested split count.
int ocrs_gemm_tc_splits(int K, int splits) {
  const int total_kb = ocrs_cdiv(K, TBK);
  const int per = ocrs_cdiv(total_kb, splits < 1 ? 1 : splits);
  return ocrs_cdiv(total_kb, per);
}

}  // extern "C"

// Device-side recognition accuracy bookkeeping: greedy CTC decode + character edit distance, replacing the host loop
// of reference ocrs_models/train_rec.py:29-68 (RecognitionAccuracyStats.update: argmax -> .tolist() -> per-sample
// ctc_greedy_decode_text / decode_text (ocrs_models/datasets/util.py:132-177) -> pylev.levenshtein), which at a
// ~9 ms training step costs several steps of host time per batch.
//
// One warp per sample:
//   1. frames are taken 32 at a time, one frame per lane: arg-max over the C classes (first maximum, like torch.argmax);
//   2. greedy CTC: a frame survives if its label differs from the previous frame's label and is not the blank
//      (util.py:163-177); survivors are visited in order through the ballot mask;
//   3. every surviving label advances one row of the Levenshtein table against the target (all non-blank labels of
//      the padded row, util.py:147: decode_text skips zeros wherever they are). A row update
//      D[i][j] = min(D[i-1][j] + 1, D[i-1][j-1] + [t_j != p_i], D[i][j-1] + 1) is done in parallel over j through
//      D[i][j] = j + min_{k <= j}(A[k] - k),  A[j] = min(D[i-1][j] + 1, D[i-1][j-1] + [t_j != p_i]),  A[0] = i
//      (a prefix-min scan: JPL consecutive target positions per lane + 5 shuffle steps).
// Labels map one-to-one to the alphabet's characters, so the distance over labels equals the distance over text.
#include "common.cuh"

namespace {

constexpr int JPL = 8;  // target positions per lane: targets up to 256 labels

__global__ void __launch_bounds__(128)
ctc_greedy_cer_kernel(const float* __restrict__ lp, int T, int N, int C, const int* __restrict__ pred_len,
                      const int* __restrict__ targets, long long tgt_stride, int S_pad, int blank,
                      int* __restrict__ edit_dist, int* __restrict__ decoded, int* __restrict__ decoded_len,
                      long long* __restrict__ total_errors) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  // target: compact the non-blank labels of the padded row into lane-owned registers (positions lane*JPL + q)
  int tgt[JPL];
#pragma unroll
  for (int q = 0; q < JPL; ++q) tgt[q] = -1;
  int S = 0;
  {
    // the row is short (<= 256): lane-strided read, ballot compaction 32 labels at a time
    for (int base = 0; base < S_pad; base += 32) {
      const int j = base + lane;
      const int v = j < S_pad ? targets[(size_t)n * tgt_stride + j] : 0;
      const unsigned keep = __ballot_sync(0xffffffffu, v != 0);
      const int pos = S + __popc(keep & ((1u << lane) - 1));  // compacted position of this lane's label
      // deliver label to the owning lane/slot: position p lives in lane p / JPL, slot p % JPL
#pragma unroll
      for (int src = 0; src < 32; ++src) {
        const int sv = __shfl_sync(0xffffffffu, v, src);
        const int sp = __shfl_sync(0xffffffffu, pos, src);
        if (((keep >> src) & 1u) && sp / JPL == lane) {
#pragma unroll
          for (int q = 0; q < JPL; ++q)
            if (sp % JPL == q) tgt[q] = sv;
        }
      }
      S += __popc(keep);
    }
  }
  // Levenshtein row 0: D[0][j] = j
  int D[JPL];
#pragma unroll
  for (int q = 0; q < JPL; ++q) D[q] = lane * JPL + q + 1;
  int P = 0;  // decoded length so far = current row index
  const int len = min(max(pred_len[n], 0), T);
  int prev = -1;  // label of the previous frame (none yet: util.py starts with last_cls = None)
  for (int t0 = 0; t0 < len; t0 += 32) {
    const int t = t0 + lane;
    int lab = blank;
    if (t < len) {
      const float* row = lp + ((size_t)t * N + n) * C;
      float best = row[0];
      int bi = 0;
      for (int c = 1; c < C; ++c) {
        const float v = row[c];
        if (v > best || (v != v && best == best)) { best = v; bi = c; }  // first maximum; NaN wins like torch.argmax
      }
      lab = bi;
    }
    int before = __shfl_up_sync(0xffffffffu, lab, 1);
    if (lane == 0) before = prev;
    const bool keep = t < len && lab != before && lab != blank;
    prev = __shfl_sync(0xffffffffu, lab, 31);  // a full chunk ends at lane 31; a partial one ends the loop
    unsigned mask = __ballot_sync(0xffffffffu, keep);
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const int p = __shfl_sync(0xffffffffu, lab, src);
      if (decoded && lane == 0) decoded[(size_t)n * T + P] = p;
      ++P;
      // one table row: A[j] - j, local prefix-min, warp exclusive prefix-min, D[j] = j + min
      int left_old = __shfl_up_sync(0xffffffffu, D[JPL - 1], 1);  // D[i-1][j-1] of this lane's first position
      if (lane == 0) left_old = P - 1;                            // D[i-1][0] = i - 1
      int m[JPL];
      int run = 0x3fffffff;
#pragma unroll
      for (int q = 0; q < JPL; ++q) {
        const int j = lane * JPL + q + 1;
        const int diag = (q == 0 ? left_old : D[q - 1]) + (tgt[q] != p ? 1 : 0);
        const int a = min(D[q] + 1, diag);
        run = min(run, a - j);
        m[q] = run;
      }
      // exclusive prefix-min over lanes of `run`, seeded with A[0] - 0 = P
      int incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = min(incl, v);
      }
      int excl = __shfl_up_sync(0xffffffffu, incl, 1);
      excl = lane == 0 ? P : min(excl, P);
      // D[q-1] (old) was consumed above before being overwritten: update back to front is not needed because m[] holds all
#pragma unroll
      for (int q = 0; q < JPL; ++q) D[q] = lane * JPL + q + 1 + min(excl, m[q]);
    }
  }
  // result D[P][S]; S = 0 -> P
  int res = P;
  if (S > 0) {
    const int owner = (S - 1) / JPL, slot = (S - 1) % JPL;
    int v = 0;
#pragma unroll
    for (int q = 0; q < JPL; ++q)
      if (slot == q) v = D[q];
    res = __shfl_sync(0xffffffffu, v, owner);
  }
  if (lane == 0) {
    edit_dist[n] = res;
    if (decoded_len) decoded_len[n] = P;
    if (total_errors) atomicAdd(reinterpret_cast<unsigned long long*>(total_errors), (unsigned long long)res);
  }
}

}  // namespace

extern "C" {

// Longest target row (S_pad) ocrs_ctc_greedy_cer accepts.
int ocrs_ctc_greedy_cer_max_targets(void) { return 32 * JPL; }

// Greedy CTC decode + character edit distance per sample, fully on the device (reference
// ocrs_models/train_rec.py:29-68 with datasets/util.py:132-177 and pylev.levenshtein):
//   lp [T][N][C] log-probs (or any scores: only the arg-max matters), pred_len [N] int32 frames to decode,
//   targets [N][tgt_stride] int32 padded with the blank; edit_dist [N] int32 out;
//   decoded [N][T] int32 / decoded_len [N] optional outputs of the decoded label sequences;
//   total_errors: optional int64 accumulator (+= sum of edit_dist, integer atomics: order independent).
int ocrs_ctc_greedy_cer(const float* lp, int T, int N, int C, const int* pred_len, const int* targets,
                        long long tgt_stride, int S_pad, int blank, int* edit_dist, int* decoded, int* decoded_len,
                        long long* total_errors, void* stream) {
  OCRS_CHECK_ARG(T > 0 && N > 0 && C > 0 && S_pad >= 0, "ctc_greedy_cer: bad dims");
  OCRS_CHECK_ARG(S_pad <= 32 * JPL, "ctc_greedy_cer: target rows longer than %d labels are not supported", 32 * JPL);
  ctc_greedy_cer_kernel<<<ocrs_cdiv(N, 4), 128, 0, (cudaStream_t)stream>>>(lp, T, N, C, pred_len, targets, tgt_stride,
                                                                          S_pad, blank, edit_dist, decoded,
                                                                          decoded_len, total_errors);
  OCRS_CHECK_LAUNCH("ctc_greedy_cer_kernel");
  return 0;
}

}  // extern "C"

// Post-loop decode on the GPU: piano roll -> note list.
// Replaces the per-note Python `while` loop of extract_notes_wo_velocity (task/utils.py:4-54; called on every finished
// roll with onsets = frames = roll at task/diffusion.py:599-602).  Integer / byte work, HBM-bound: one coalesced read of
// the roll(s), bit-exact output including the note order (frame-major, then pitch = np.nonzero order).
//
//   pass 1  one thread per (roll, pitch) walks time BACKWARDS keeping `end` = first frame >= t where neither the onset
//           nor the frame activation is on (the reference's while loop), and marks every rising onset edge whose frame
//           activation is on with that end (int16 scratch, -1 elsewhere).  Warps read 32 neighbouring pitches: coalesced.
//   pass 2  one block per roll walks time forwards; a ballot prefix over the 88 pitches of a frame gives each note its
//           slot, so notes come out in the reference's order without a sort.
#include "common.cuh"
#include "kernels.h"

namespace drb {

__global__ void notes_mark_kernel(const float* __restrict__ onsets, const float* __restrict__ frames, int16_t* __restrict__ endmark,
                                  int B, int T, int P, float on_thr, float fr_thr, int rule) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * P) return;
  const int b = idx / P, p = idx - b * P;
  const float* on = onsets + (size_t)b * T * P + p;
  const float* fr = frames + (size_t)b * T * P + p;
  int16_t* em = endmark + (size_t)b * T * P + p;
  int end = T;
  bool o_cur = on[(size_t)(T - 1) * P] > on_thr;
  for (int t = T - 1; t >= 0; --t) {
    const bool f_cur = fr[(size_t)t * P] > fr_thr;
    const bool o_prev = t > 0 ? (on[(size_t)(t - 1) * P] > on_thr) : false;
    if (!(o_cur || f_cur)) end = t;
    // onset_diff: onsets[t] - onsets[t-1] == 1 (first row: onsets[0] == 1); rule1: and frames[t] == 1; rule2: nothing more
    const bool edge = o_cur && !o_prev && (f_cur || rule == 2);
    em[(size_t)t * P] = edge ? (int16_t)end : (int16_t)-1;   // edge implies o_cur, so end > t
    o_cur = o_prev;
  }
}

__global__ void __launch_bounds__(128) notes_compact_kernel(const int16_t* __restrict__ endmark, int32_t* __restrict__ pitches,
                                                            int32_t* __restrict__ intervals, int32_t* __restrict__ counts,
                                                            int T, int P, int max_notes) {
  __shared__ int warp_cnt[4];
  const int b = blockIdx.x, p = threadIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int16_t* em = endmark + (size_t)b * T * P;
  int32_t* out_p = pitches + (size_t)b * max_notes;
  int32_t* out_i = intervals + (size_t)b * max_notes * 2;
  int base = 0;
  for (int t = 0; t < T; ++t) {
    const int e = p < P ? (int)em[(size_t)t * P + p] : -1;
    const bool has = e >= 0;
    const unsigned m = __ballot_sync(0xffffffffu, has);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) { if (w < warp) off += warp_cnt[w]; tot += warp_cnt[w]; }
    if (has) {
      const int slot = base + off + __popc(m & ((1u << lane) - 1u));
      if (slot < max_notes) { out_p[slot] = p; out_i[2 * slot] = t; out_i[2 * slot + 1] = e; }
    }
    base += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[b] = base;
}

// Frame-level confusion counts of test_step's precision_recall_fscore_support(label.flatten(), pred.flatten() > thr,
// average='binary') (task/diffusion.py:378-380): TP, FP, FN over all elements.  One coalesced pass, integer counts.
__global__ void __launch_bounds__(256) frame_counts_kernel(const float* __restrict__ pred, const float* __restrict__ label, size_t n,
                                                           float thr, unsigned long long* __restrict__ counts) {
  unsigned tp = 0, fp = 0, fn = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const bool p = pred[i] > thr, l = label[i] == 1.0f;      // pos_label = 1
    tp += p && l; fp += p && !l; fn += !p && l;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tp += __shfl_down_sync(0xffffffffu, tp, o); fp += __shfl_down_sync(0xffffffffu, fp, o); fn += __shfl_down_sync(0xffffffffu, fn, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (tp) atomicAdd(counts + 0, (unsigned long long)tp);
    if (fp) atomicAdd(counts + 1, (unsigned long long)fp);
    if (fn) atomicAdd(counts + 2, (unsigned long long)fn);
  }
}

}  // namespace drb

using namespace drb;

extern "C" int drb_frame_counts(const float* pred, const float* label, int64_t n, float threshold, uint64_t* counts3, void* stream) {
  if (!pred || !label || !counts3 || n <= 0) { set_error("frame_counts: bad argument"); return DRB_E_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  DRB_CUDA(cudaMemsetAsync(counts3, 0, 3 * sizeof(uint64_t), s));
  size_t blocks = ((size_t)n + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  frame_counts_kernel<<<(unsigned)blocks, 256, 0, s>>>(pred, label, (size_t)n, threshold, reinterpret_cast<unsigned long long*>(counts3));
  DRB_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t drb_extract_notes_scratch_bytes(int32_t B, int32_t T, int32_t P) {
  if (B <= 0 || T <= 0 || P <= 0) return 0;
  return (size_t)B * T * P * sizeof(int16_t);
}

extern "C" int drb_extract_notes(const float* onsets, const float* frames, int32_t B, int32_t T, int32_t P, float onset_threshold,
                                 float frame_threshold, int32_t rule, void* scratch, int32_t* pitches, int32_t* intervals,
                                 int32_t* counts, int32_t max_notes, void* stream) {
  if (!onsets || !frames || !scratch || !pitches || !intervals || !counts || B <= 0 || T <= 0 || P <= 0 || P > 128 || T > 32767 ||
      max_notes <= 0 || (rule != 1 && rule != 2)) {
    set_error("extract_notes: bad argument (need 0 < P <= 128, 0 < T <= 32767, rule 1 or 2)");
    return DRB_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int n = B * P;
  notes_mark_kernel<<<(n + 127) / 128, 128, 0, s>>>(onsets, frames, (int16_t*)scratch, B, T, P, onset_threshold, frame_threshold, rule);
  DRB_LAUNCH_CHECK();
  notes_compact_kernel<<<B, 128, 0, s>>>((const int16_t*)scratch, pitches, intervals, counts, T, P, max_notes);
  DRB_LAUNCH_CHECK();
  return 0;
}

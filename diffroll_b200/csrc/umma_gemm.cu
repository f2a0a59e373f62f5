// tcgen05 (5th-gen tensor core) kernels for the contractions of the ResidualBlock stack
// (model/diffwave.py:134-151, 680-684), sm_100a only.
//
//   gate kernel  : y[t, n] = sum_{tap, c} xin[t + (tap-k/2)*dil, c] * Wd[n, tap, c]  (+ spec[t, :] . Wc[n, :])  + bias1[n]
//                  z_l[t, c] = sigmoid(y[t, c]) * tanh(y[t, C + c])                     -> operand pair, kept for every layer l
//   zgemm RES    : o[t, c] = sum_k z_l[t, k] * Wo_l[c, k] + bo_l[c]            (residual half of output_projection)
//                  x[t, c] = (x[t, c] + o[t, c]) / sqrt(2)    -> fp32, and the operand pair of (x + d_{l+1})
//   zgemm HEAD   : h[t, n] = relu( sum_l sum_k z_l[t, k] * Wcomp[n, l*C + k] + bcomp[n] )
//                  with Wcomp_l = skip_projection . Wo_l[skip half] / sqrt(L): the skip sum (diffwave.py:680), the
//                  1/sqrt(L) scale and skip_projection (:682-684) are one long-K GEMM over the stored z_l instead of
//                  15 read-modify-write passes over a skip buffer.
//
// All are implicit GEMMs with M = time (128-frame tiles, one roll per tile), N = 256 output channels, K-slabs of 64
// input channels.  The dilated taps are NOT materialised: each (tap, 64 channel) K-slab is a TMA box load of the
// activation tensor [roll][frame][channel] at frame offset (tap - k/2)*dil; frames outside [0, T) are zero-filled by
// the TMA unit, which is exactly the conv's zero padding and cannot bleed into the neighbouring roll.
//
// fp32 parity (|delta| < 1e-3 after 200 chained steps) needs more than one 16-bit product, so
// activations and weights are kept as operand PAIRS (2-byte main + 2-byte aux per element, formats in `Cfg` below and
// DESIGN.md section 2); the default f16e5 issues one fp16 MMA plus one e5m2 correction MMA per K-step into the same
// TMEM accumulator (bf16x3: three bf16 MMAs; f16f8: an e4m3 correction into a second accumulator).
//
// CTA pairs (PAIR = true, the default when the number of M tiles is even): two CTAs of a cluster issue ONE
// tcgen05.mma.cta_group::2 with M = 256.  Each CTA stages its own 128 activation rows and only HALF of the weight
// tile (128 of the 256 output rows), so the shared-memory traffic per MMA (operand reads + TMA fills), which bounds
// the single-CTA kernel, drops by a third and the ring holds 3 stages instead of 2.
//
// Kernels, from the general fallback to the ones the benchmark runs:
//   umma_gate_kernel<P, PAIR>        one tile per CTA (pair), one TMA box per tap          (odd tile counts, P = 0)
//   umma_gate_win_kernel<P>          + tap window fetched once per 64-channel chunk        (f16f8: needs 512 TMEM columns)
//   umma_gate_pers_kernel<P, DUAL>   + persistent pairs, two accumulator stages, conditioner term from the per-clip
//                                      fp32 table in the epilogue; DUAL = layer-0 conv shared by both guidance branches
//   umma_gate_n4_kernel<DUAL>        the default: f16n4 operands (fp16 product + block-scaled fp4 correction), scale factors in TMEM
//   umma_conv_lin_pers_kernel<P>     linear fp32 epilogue, persistent pairs over (tile, tap pass) units: training forward / dgrad and
//                                      the per-clip conditioner tables; lin_fixup_kernel adds the parked partial tiles
//   umma_zgemm_kernel<P, PAIR>       RES / HEAD, one tile per CTA (pair); mode 3 = output projection + guidance + posterior,
//                                      mode 4 = plain GEMM with optional split-K (training weight gradients)
//   umma_res_pers_kernel<P, XF>      RES, persistent pairs, register-prefetched x tile, TMA-staged stores (XF = 4: f16n4 emission)
//   umma_head_pers_kernel<P>         HEAD (K = L * C), persistent pairs, operand-pair output
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warp 3 = bias stager /
// second producer, warps 4-11 = epilogue (TMEM -> registers -> swizzled smem staging -> TMA store).  smem ring:
// full/empty mbarriers per stage.
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"

namespace drb {

constexpr int TILE_M = 128;
constexpr int TILE_N = 256;
constexpr int TILE_K = 64;
constexpr int A_TILE_BYTES = TILE_M * TILE_K * 2;  // 16 KB
constexpr int B_TILE_BYTES = TILE_N * TILE_K * 2;  // 32 KB
constexpr int UMMA_K = 16;
constexpr int CHUNK_BYTES = TILE_M * 128;          // one 128-row x 128-byte swizzled staging box (16 KB)
constexpr int EPI_BAR = 1;                         // named barrier of the epilogue warps
constexpr int NUM_THREADS = 384;                   // warps 0-3: producer / MMA / TMEM / bias staging; warps 4-11: epilogue
constexpr int EPI_THREADS = 256;

// Precision modes (template parameter P):
//   0  bf16     one product                                     main operands only
//   1  bf16x3   bf16 hi/lo, 3 products into one accumulator      aux = bf16 lo tiles
//   2  f16f8    fp16 product + e4m3 correction product          aux = e4m3 tiles [lo*SA | hi] / [hi*SW | lo*SA*SW],
//                                                                second accumulator (TMEM columns 256..511)
//   3  f16e5    fp16 product + e5m2 correction product          aux = e5m2 tiles, unit total scale: same accumulator
//   4  f16x3    fp16 hi/lo (lo = fp16(v - hi): 22 mantissa bits), 3 products into one accumulator: fp32-grade, used by the
//               TRAINING forward (csrc/train.cu), whose rounding would otherwise move ReLU / gate / loss derivatives
template <int P, bool PAIR>
struct Cfg {
  static constexpr bool kAux = P != 0;
  static constexpr int kBBytes = PAIR ? B_TILE_BYTES / 2 : B_TILE_BYTES;  // weight rows staged by this CTA: 128 or 256
  static constexpr int kStageBytes = (kAux ? 2 : 1) * (A_TILE_BYTES + kBBytes);
  static constexpr int kRingBytes = 196608;                                // 192 KB: also the epilogue staging area
  static constexpr int kStages = kRingBytes / kStageBytes;                 // 2 (single) / 3 (pair) with aux operands
  static constexpr int kSmemBytes = kRingBytes + 1024 /*align slack*/ + 256 /*barriers*/ + 2048 /*bias, d_next*/;
  static constexpr uint32_t kTmemCols = P == 2 ? 512 : 256;
  static constexpr int kAuxMul = (P == 2 || P == 3) ? 2 : 1;  // aux element coordinate = kAuxMul * main element coordinate
  static constexpr int kAOff = 0, kAAuxOff = A_TILE_BYTES;
  static constexpr int kBOff = (kAux ? 2 : 1) * A_TILE_BYTES, kBAuxOff = kBOff + kBBytes;
  static constexpr int kMaxStages = 6;
};

struct alignas(64) GateParams {
  CUtensorMap xh, xl, sh, sl, wd_h, wd_l, wc_h, wc_l, zh, zl;
  CUtensorMap xwh, xwl;  // activation maps whose box spans the whole tap window: 128 + (taps-1)*dil frames
  int win_rows, n_items;
  int NB, n_cond, T, C, taps, dil, cond_slabs, tiles_t, n_blocks, z_group0;
  const float* bias_cond;
  const float* bias_unc;
  const float* inv_scale;  // f16f8: 1 / (SA * SW), f16e5: 1 / SW of this layer's weights (device scalar)
  // Persistent kernel only.  The conditioner projection is step-invariant, so it is computed once per clip in fp32
  // (cond[roll][t][2C], natural channel order: gate 0..C-1, filter C..2C-1) and ADDED IN THE EPILOGUE of conditional
  // rolls (nb < n_cond) instead of being re-contracted every step as extra K-slabs (those run for nb < n_cond_mma).
  // Layer-0 branch sharing (DUAL): both guidance branches read the SAME x at layer 0, so the dilated conv is computed
  // once per roll and the epilogue emits two gated outputs: the unconditional one (bias_unc) into roll nb + dual_off and
  // the conditional one (bias_cond + cond) into roll nb.
  const float* cond;
  int n_cond_mma;
  int dual_off;
  // f16n4 kernel only: aux activation window map (64-byte rows of e2m1 codes, SWIZZLE_64B), aux weight map, weight scale-factor
  // atoms (2 KB per N block and K-slab, un-swizzled), activation scale factors [roll][chunk][frame][8]
  CUtensorMap xw4, wd4, wsf;
  const uint8_t* xs;
  // Linear epilogue (training, csrc/train.cu): instead of the gate, every accumulator column leaves as fp32
  // out[(roll * T + t) * ldo + nblk * 256 + column] = acc / scale + bias[column]  (weights in natural row order).
  float* lin_out;
  int ldo;
  // A launch may cover only the taps [tap_lo, tap_lo + tap_n) (tap_n == 0: all) and ADD its result to lin_out (lin_acc): the tensor
  // core's fp32 accumulation loses ~1e-8 of the running sum per K-step, which a 4608-long chain turns into 4e-5 -- short chains summed
  // in exact fp32 in global memory keep the training forward at the fp32 floor (csrc/train.cu)
  int tap_lo, tap_n, lin_acc;
  float* lin_scratch;   // persistent linear kernel: [pairs][2][128][256] fp32 partial tiles, or nullptr (ranges cut at whole items)
  int tap_span;   // persistent linear kernel: taps [tap_lo, tap_lo + tap_span) in passes of tap_n, summed in out[] (0: one pass)
};

struct alignas(64) ZGemmParams {
  CUtensorMap zh, zl, w_h, w_l, out32, xh, xl;
  int NB, T, C, tiles_t, n_blocks, nslabs, spg, z_group0, group_stride, mode;  // mode 0 = RES, 1 = HEAD
  const float* bias;
  const float* dnext;      // RES: next layer's diffusion_projection table [timesteps][C]
  const float* inv_scale;
  const int* steps;        // per-sample diffusion steps (device, [bsamp]) or nullptr: every roll uses row t_uniform
  int t_uniform, bsamp;
  unsigned int* range_max; // RES: running max |x + d_next| over the emitted operands (fp32 bits), or nullptr
  uint8_t* xs;             // f16n4: activation scale factors [roll][chunk][frame][8] (xl then maps the 64-byte e2m1 rows)
  // mode 1 with h_pair: relu output as an operand pair through xh / xl.  mode 3: head output projection + guidance + posterior
  int h_pair, dual_B, F;
  int ksplit;              // mode 4: NB counts K ranges of nslabs slabs each (see the producer)
  drb_update upd;
  const float* x_t; const float* noise; float* x_prev; float* net_out;
};

struct SmemView {
  uint8_t* stage0;
  uint64_t* full;
  uint64_t* empty;
  uint64_t* tmem_full;
  uint64_t* xin_full;   // [2]: early / late x-tile boxes (zgemm RES)
  uint64_t* afull;      // [2] activation-window buffers (gate window kernel)
  uint64_t* aempty;     // [2]
  uint32_t* tmem_ptr;
  float* sbias;         // [256] bias of this tile's columns
  float* sdn;           // [256] d_next of this tile's channels (zgemm RES)
};

template <int P, bool PAIR>
__device__ __forceinline__ SmemView carve(uint8_t* raw) {
  uint32_t a = smem_u32(raw);
  uint32_t pad = ((a + 1023u) & ~1023u) - a;
  SmemView v;
  v.stage0 = raw + pad;
  uint8_t* bars = v.stage0 + Cfg<P, PAIR>::kRingBytes;
  v.full = reinterpret_cast<uint64_t*>(bars);
  v.empty = v.full + Cfg<P, PAIR>::kMaxStages;
  v.tmem_full = v.empty + Cfg<P, PAIR>::kMaxStages;
  v.xin_full = v.tmem_full + 1;
  v.afull = v.xin_full + 2;
  v.aempty = v.afull + 2;
  v.tmem_ptr = reinterpret_cast<uint32_t*>(v.aempty + 2);
  v.sbias = reinterpret_cast<float*>(bars + 256);
  v.sdn = v.sbias + 256;
  return v;
}

template <int P, bool PAIR>
__device__ __forceinline__ void prologue(const SmemView& sv, int warp) {
  if (warp == 1 && elect_one()) {
    // pair: the leader's full barrier collects one arrive.expect_tx from each CTA's producer
    for (int i = 0; i < Cfg<P, PAIR>::kStages; ++i) { mbar_init(&sv.full[i], PAIR ? 2 : 1); mbar_init(&sv.empty[i], 1); }
    mbar_init(sv.tmem_full, 1);
    mbar_init(&sv.xin_full[0], 1); mbar_init(&sv.xin_full[1], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&sv.afull[i], PAIR ? 2 : 1); mbar_init(&sv.aempty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) { tmem_alloc_pair(sv.tmem_ptr, Cfg<P, PAIR>::kTmemCols); tmem_relinquish_pair(); }
    else { tmem_alloc(sv.tmem_ptr, Cfg<P, PAIR>::kTmemCols); tmem_relinquish(); }
  }
  pdl_trigger();                 // the next kernel may start its own prologue on SMs this grid has left
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // both CTAs' barriers and TMEM exist before any remote arrive / paired MMA
  tc_fence_after();
  pdl_wait();                    // everything above touched only constants; from here on we read the previous kernel's output
}

template <int P, bool PAIR>
__device__ __forceinline__ void teardown(uint32_t tmem_base, int warp) {
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer may still touch this CTA's barriers / TMEM until its work has retired
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, Cfg<P, PAIR>::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg<P, PAIR>::kTmemCols);
  }
}

// Producer-side helpers.  `fb` is the address the TMA completion goes to: this CTA's full barrier, or (pair) the
// leader CTA's full barrier as a shared::cluster address.
template <bool PAIR>
__device__ __forceinline__ void prod_expect(uint64_t* full_local, uint32_t fb, uint32_t bytes) {
  if (PAIR) mbar_expect_tx_cluster(fb, bytes); else mbar_expect_tx(full_local, bytes);
}
template <bool PAIR>
__device__ __forceinline__ void load_a(uint8_t* dst, const void* tmap, uint64_t* full_local, uint32_t fb, int c0, int c1, int c2) {
  if (PAIR) tma_load_3d_pair(dst, tmap, fb, c0, c1, c2); else tma_load_3d(dst, tmap, full_local, c0, c1, c2);
}
// weight tile: 256 rows = two 128-row boxes (single CTA), or this CTA's 128-row half (pair)
template <bool PAIR>
__device__ __forceinline__ void load_b(uint8_t* dst, const void* tmap, uint64_t* full_local, uint32_t fb, int col, int row0, uint32_t rank) {
  if (PAIR) {
    tma_load_2d_pair(dst, tmap, fb, col, row0 + (int)rank * (TILE_N / 2));
  } else {
    tma_load_2d(dst, tmap, full_local, col, row0);
    tma_load_2d(dst + B_TILE_BYTES / 2, tmap, full_local, col, row0 + TILE_N / 2);
  }
}

// MMA issue for one K-slab (64 channels = 4 UMMA K-steps) resident in stage memory.
template <int P, bool PAIR>
__device__ __forceinline__ void issue_slab(uint8_t* st, uint32_t tmem_d, bool first_slab) {
  using C = Cfg<P, PAIR>;
  const uint32_t a_hi = smem_u32(st) + C::kAOff, a_lo = smem_u32(st) + C::kAAuxOff;
  const uint32_t b_hi = smem_u32(st) + C::kBOff, b_lo = smem_u32(st) + C::kBAuxOff;
  constexpr int M = PAIR ? 2 * TILE_M : TILE_M;
  constexpr uint32_t idesc = P >= 2 ? make_idesc_fmt0(M, TILE_N) : make_idesc_bf16(M, TILE_N);
  constexpr uint32_t idesc_e5 = make_idesc_bf16(M, TILE_N);   // kind::f8f6f4 format code 1 = e5m2 (same bits as bf16 for kind::f16)
  const uint32_t acc0 = first_slab ? 0u : 1u;
#pragma unroll
  for (int k = 0; k < TILE_K / UMMA_K; ++k) {
    const uint32_t ko = k * UMMA_K * 2;  // byte advance inside the 128-byte swizzled row (16 x 2 B, or 32 x 1 B for e4m3)
    const uint64_t da_hi = make_sw128_desc(a_hi + ko), db_hi = make_sw128_desc(b_hi + ko);
    const uint32_t acc = k == 0 ? acc0 : 1u;
    if (PAIR) {
      umma_bf16_pair(tmem_d, da_hi, db_hi, idesc, acc);   // kind::f16: bf16 or fp16 per idesc
      if (P == 1 || P == 4) {
        umma_bf16_pair(tmem_d, make_sw128_desc(a_lo + ko), db_hi, idesc, 1u);
        umma_bf16_pair(tmem_d, da_hi, make_sw128_desc(b_lo + ko), idesc, 1u);
      }
      if (P == 2) umma_f8_pair(tmem_d + 256, make_sw128_desc(a_lo + ko), make_sw128_desc(b_lo + ko), idesc, acc);
      if (P == 3) umma_f8_pair(tmem_d, make_sw128_desc(a_lo + ko), make_sw128_desc(b_lo + ko), idesc_e5, 1u);
    } else {
      umma_bf16(tmem_d, da_hi, db_hi, idesc, acc);
      if (P == 1 || P == 4) {
        umma_bf16(tmem_d, make_sw128_desc(a_lo + ko), db_hi, idesc, 1u);
        umma_bf16(tmem_d, da_hi, make_sw128_desc(b_lo + ko), idesc, 1u);
      }
      if (P == 2) umma_f8(tmem_d + 256, make_sw128_desc(a_lo + ko), make_sw128_desc(b_lo + ko), idesc, acc);
      if (P == 3) umma_f8(tmem_d, make_sw128_desc(a_lo + ko), make_sw128_desc(b_lo + ko), idesc_e5, 1u);
    }
  }
}

// Consumer loop, one elected thread (pair: of the leader CTA only).
template <int P, bool PAIR>
__device__ __forceinline__ void mma_loop(const SmemView& sv, uint32_t tmem_base, int nslabs, int nst) {
  int stage = 0; uint32_t phase = 0;
  for (int s = 0; s < nslabs; ++s) {
    mbar_wait(&sv.full[stage], phase);
    tc_fence_after();
    issue_slab<P, PAIR>(sv.stage0 + stage * Cfg<P, PAIR>::kStageBytes, tmem_base, s == 0);
    if (PAIR) umma_commit_pair(&sv.empty[stage]);  // frees the slot in both CTAs when these MMAs retire
    else umma_commit(&sv.empty[stage]);
    if (++stage == nst) { stage = 0; phase ^= 1; }
  }
  if (PAIR) umma_commit_pair(sv.tmem_full); else umma_commit(sv.tmem_full);
}

// sigmoid(g) * tanh(f) with the SFU exp2/rcp approximations (abs error ~2e-7, far below the bf16 hi/lo split error)
__device__ __forceinline__ float gate_act(float g, float f) {
  const float sg = __fdividef(1.f, 1.f + __expf(-g));
  const float th = 1.f - __fdividef(2.f, __expf(2.f * f) + 1.f);
  return sg * th;
}

// 32 consecutive accumulator columns of this thread's row; f16f8 adds the scaled correction accumulator.
// The whole warp must be converged at the call (tcgen05.ld is .sync.aligned).
template <int P>
__device__ __forceinline__ void load_acc32(uint32_t taddr, float inv_scale, float (&v)[32]) {
  uint32_t r[32];
  __syncwarp();
  tmem_ld32(taddr, r);
  if (P == 2) {
    uint32_t c[32];
    tmem_ld32(taddr + 256, c);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(c[i]), inv_scale, __uint_as_float(r[i]));
  } else if (P == 3 || P == 4) {   // f16e5 / f16x3: weights were split as W * SW (a power of two, keeps small weights out of the fp16 subnormals)
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * inv_scale;
  } else {
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  }
}

// Largest |value| (as fp32 bits; NaN > inf > finite in this order) of the activation operands a kernel emitted: the fp16
// main part of the f16 formats overflows at 65504, so the host can tell when a run left the format's range.
__device__ __forceinline__ void range_note(unsigned int& umax, const float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) umax = max(umax, __float_as_uint(v[i]) & 0x7fffffffu);
}
__device__ __forceinline__ void range_publish(unsigned int* dst, unsigned int umax) {
  if (!dst) return;
  umax = __reduce_max_sync(0xffffffffu, umax);
  if ((threadIdx.x & 31) == 0 && umax > *reinterpret_cast<volatile unsigned int*>(dst)) atomicMax(dst, umax);
}

// Stage 16 consecutive channel values (channel offset ch inside a 64-channel box, multiple of 16) of one row into the
// swizzled operand boxes of the next GEMM.  main box: [128 rows][64 channels] of 2-byte elements; aux box: bf16 lo
// parts (same shape) or e4m3 bytes [lo*SA (64) | hi (64)] per row.
template <int P>
__device__ __forceinline__ void stage16(uint32_t main_box, uint32_t aux_box, int row, int ch, const float (&v)[16]) {
  if (P >= 2) {
    uint32_t h[8], lo[4], hi[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float t[4] = {v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]};
      if (P == 2) split_f16f8_x4(t, h[2 * q], h[2 * q + 1], lo[q], hi[q]);
      else split_f16e5_x4(t, h[2 * q], h[2 * q + 1], lo[q], hi[q]);
    }
    sts128u(main_box + sw128_off(row, ch / 8), h[0], h[1], h[2], h[3]);
    sts128u(main_box + sw128_off(row, ch / 8 + 1), h[4], h[5], h[6], h[7]);
    sts128u(aux_box + sw128_off(row, ch / 16), lo[0], lo[1], lo[2], lo[3]);
    sts128u(aux_box + sw128_off(row, 4 + ch / 16), hi[0], hi[1], hi[2], hi[3]);
  } else {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_pack2(v[8 * hf + 2 * e], v[8 * hf + 2 * e + 1], hi[e], lo[e]);
      sts128u(main_box + sw128_off(row, ch / 8 + hf), hi[0], hi[1], hi[2], hi[3]);
      if (P == 1) sts128u(aux_box + sw128_off(row, ch / 8 + hf), lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// Gate epilogue (8 warps): accumulator -> bias -> sigmoid*tanh -> operand split -> swizzled staging -> TMA store of z.
template <int P>
__device__ __forceinline__ void gate_epilogue(const SmemView& sv, const GateParams& p, uint32_t tmem_base, int warp, int lane,
                                              int nb, int nblk, int t0) {
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int grp = (warp - 4) >> 2;       // two groups of 4 warps split the tile's channels
    const int row = q * 32 + lane;
    const float* bias = sv.sbias;
    mbar_wait(sv.tmem_full, 0);  // every MMA has retired: accumulator complete, ring memory free
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (p.lin_out) {                       // training: plain fp32 result, natural column order, straight from registers
      const float inv_l = (P >= 2) ? __ldg(p.inv_scale) : 0.f;
      const int t = t0 + row;
      float* dst = p.lin_out + ((size_t)nb * p.T + (t < p.T ? t : 0)) * (size_t)p.ldo + nblk * TILE_N + grp * 128;
#pragma unroll 1
      for (int c4 = 0; c4 < 4; ++c4) {
        float v[32];
        load_acc32<P>(taddr + grp * 128 + c4 * 32, inv_l, v);
        if (t < p.T) {
          // the bias is read from global memory HERE, after griddepcontrol.wait: in training it is produced by the kernel launched
          // just before this one, and the shared-memory copy staged in the prologue may predate it (programmatic dependent launch).
          // Accumulating launches fetch their 8 float4 of old values BEFORE the first store of the chunk: interleaved load/store
          // pairs on the same row ran 9x slower (570 vs 66 us per launch), every load queued behind the previous store.
          float4 o4[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            o4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias_cond) o4[i] = *reinterpret_cast<const float4*>(p.bias_cond + nblk * TILE_N + grp * 128 + c4 * 32 + i * 4);
          }
          if (p.lin_acc) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 q = __ldcg(reinterpret_cast<const float4*>(dst + c4 * 32 + i * 4));
              o4[i].x += q.x; o4[i].y += q.y; o4[i].z += q.z; o4[i].w += q.w;
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(dst + c4 * 32 + i * 4) = make_float4(v[4 * i] + o4[i].x, v[4 * i + 1] + o4[i].y, v[4 * i + 2] + o4[i].z, v[4 * i + 3] + o4[i].w);
        }
      }
      tc_fence_before();
      return;
    }
    // staging: z main boxes 0,1 then z aux boxes 0,1; each [128 frames][128 bytes], 128-byte swizzled
    const uint32_t stg = smem_u32(sv.stage0);
    const float inv = (P >= 2) ? __ldg(p.inv_scale) : 0.f;
#pragma unroll 1
    for (int c2 = 0; c2 < 2; ++c2) {
      const int ch = grp * 2 + c2;         // 32 gate + 32 filter columns; group g fills box g
      float g[32], f[32];
      load_acc32<P>(taddr + ch * 32, inv, g);
      load_acc32<P>(taddr + 128 + ch * 32, inv, f);
      const uint32_t box_m = stg + (ch >> 1) * CHUNK_BYTES, box_a = box_m + 2 * CHUNK_BYTES;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j = hf * 16 + i;
          z[i] = gate_act(g[j] + bias[ch * 32 + j], f[j] + bias[128 + ch * 32 + j]);
        }
        stage16<P>(box_m, box_a, row, (ch & 1) * 32 + hf * 16, z);
      }
    }
    tc_fence_before();
    fence_proxy_async();  // make the generic-proxy smem writes visible to the TMA (async proxy)
    named_bar_sync(EPI_BAR, EPI_THREADS);
    if (warp == 4 && elect_one()) {
      const int c0 = nblk * (TILE_N / 2);
      tma_store_3d(&p.zh, sv.stage0, c0, t0, p.z_group0 + nb);
      tma_store_3d(&p.zh, sv.stage0 + CHUNK_BYTES, c0 + 64, t0, p.z_group0 + nb);
      if ((P != 0)) {
        tma_store_3d(&p.zl, sv.stage0 + 2 * CHUNK_BYTES, (P >= 2 ? 2 : 1) * c0, t0, p.z_group0 + nb);
        tma_store_3d(&p.zl, sv.stage0 + 3 * CHUNK_BYTES, (P >= 2 ? 2 : 1) * (c0 + 64), t0, p.z_group0 + nb);
      }
      tma_store_commit();
      tma_store_wait_read<0>();  // smem must stay valid until the bulk stores have read it
    }
}

// ---------------------------------------------------------------------------------------------
// gate kernel
// ---------------------------------------------------------------------------------------------
template <int P, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1) umma_gate_kernel(const __grid_constant__ GateParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmemView sv = carve<P, PAIR>(smem_raw);
  using CF = Cfg<P, PAIR>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile mapping: a CTA pair = two consecutive M tiles (frames x roll) of the same N block
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  int cid = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nblk = cid % p.n_blocks; cid /= p.n_blocks;
  const int mt = PAIR ? cid * 2 + (int)rank : cid;
  const int tt = mt % p.tiles_t;
  const int nb = mt / p.tiles_t;
  const int t0 = tt * TILE_M;
  const int cpt = p.C / TILE_K;  // K-slabs per tap
  const int conv_slabs = (p.tap_n > 0 ? p.tap_n : p.taps) * cpt;
  // Both CTAs of a cluster run the same slab list.  If only the first tile of the pair is conditional, the second
  // one runs the conditioner slabs too: its spectrogram coordinate (roll >= n_cond) is out of bounds, so TMA feeds zeros.
  const int nb_first = PAIR ? (cid * 2) / p.tiles_t : nb;
  const int nslabs = conv_slabs + (nb_first < p.n_cond ? p.cond_slabs : 0);
  const int half = p.taps / 2;
  constexpr int AM = CF::kAuxMul;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.xh); tma_prefetch_desc(&p.wd_h); tma_prefetch_desc(&p.zh);
    if (CF::kAux) { tma_prefetch_desc(&p.xl); tma_prefetch_desc(&p.wd_l); tma_prefetch_desc(&p.zl); }
  }
  if (warp == 3) {  // this tile's 256 bias values -> smem (the epilogue reads them as warp-wide broadcasts)
    const float* bsrc = (nb < p.n_cond ? p.bias_cond : p.bias_unc);
    for (int i = lane; i < TILE_N; i += 32) sv.sbias[i] = bsrc ? __ldg(bsrc + nblk * TILE_N + i) : 0.f;
  }
  prologue<P, PAIR>(sv, warp);
  const uint32_t tmem_base = *sv.tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int s = 0; s < nslabs; ++s) {
        mbar_wait(&sv.empty[stage], phase ^ 1);
        uint8_t* st = sv.stage0 + stage * CF::kStageBytes;
        uint64_t* fl = &sv.full[stage];
        const uint32_t fb = PAIR ? mapa_cluster(smem_u32(fl), 0) : 0u;
        prod_expect<PAIR>(fl, fb, CF::kStageBytes);
        if (s < conv_slabs) {
          const int tap = p.tap_lo + s / cpt, cc = s % cpt;
          const int trow = t0 + (tap - half) * p.dil;
          load_a<PAIR>(st + CF::kAOff, &p.xh, fl, fb, cc * TILE_K, trow, nb);
          load_b<PAIR>(st + CF::kBOff, &p.wd_h, fl, fb, tap * p.C + cc * TILE_K, nblk * TILE_N, rank);
          if (CF::kAux) {
            load_a<PAIR>(st + CF::kAAuxOff, &p.xl, fl, fb, AM * cc * TILE_K, trow, nb);
            load_b<PAIR>(st + CF::kBAuxOff, &p.wd_l, fl, fb, AM * (tap * p.C + cc * TILE_K), nblk * TILE_N, rank);
          }
        } else {
          const int cc = s - conv_slabs;
          load_a<PAIR>(st + CF::kAOff, &p.sh, fl, fb, cc * TILE_K, t0, nb);
          load_b<PAIR>(st + CF::kBOff, &p.wc_h, fl, fb, cc * TILE_K, nblk * TILE_N, rank);
          if (CF::kAux) {
            load_a<PAIR>(st + CF::kAAuxOff, &p.sl, fl, fb, AM * cc * TILE_K, t0, nb);
            load_b<PAIR>(st + CF::kBAuxOff, &p.wc_l, fl, fb, AM * cc * TILE_K, nblk * TILE_N, rank);
          }
        }
        if (++stage == CF::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) mma_loop<P, PAIR>(sv, tmem_base, nslabs, CF::kStages);
  } else if (warp >= 4) {
    gate_epilogue<P>(sv, p, tmem_base, warp, lane, nb, nblk, t0);
  }
  teardown<P, PAIR>(tmem_base, warp);
}


// ---------------------------------------------------------------------------------------------
// Linear conv, persistent form (training forward, conditioner tables).  The one-tile-per-CTA kernel above with the linear epilogue
// spent two thirds of a launch outside the MMAs (prologue + a serialised epilogue per tile, 66 us for 7 us of tensor work at one tap
// per launch).  Here one CTA pair per SM pair walks its (tile, tap pass) units: two TMEM accumulator stages, so the epilogue of unit
// i overlaps the MMAs of unit i + 1; the tap passes of a tile run back to back in the same CTA (the exact-fp32 sum across passes
// goes through the tile's 128 KB of out[] while it is still in L2: a thread re-reads only addresses it wrote itself); and the
// epilogue transposes each 32x32 accumulator block through a private 4 KB of shared memory so that every global load / store
// instruction of a warp covers four whole 128-byte lines (thread = row gave 32 lines x 16 bytes per instruction).
//   passes: taps [tap_lo, tap_lo + span) in groups of tap_n; pass 0 carries the bias (and lin_acc), the last one the 1x1 slabs.
//   smem: 3 stages x 64 KB | barriers | 8 x 4 KB transposition buffers
// ---------------------------------------------------------------------------------------------
constexpr int CLP_SMEM = 196608 + 1024 /*align*/ + 256 /*barriers*/ + 8 * 4096;
static_assert(CLP_SMEM <= 232448, "linear conv kernel: shared memory over the 227 KB limit");

template <int P>
__global__ void __launch_bounds__(NUM_THREADS, 1) umma_conv_lin_pers_kernel(const __grid_constant__ GateParams p) {
  static_assert(P == 3 || P == 4, "single 256-column accumulator formats");
  constexpr bool PAIR = true;
  using CF = Cfg<P, PAIR>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmemView sv = carve<P, PAIR>(smem_raw);
  uint64_t* const tfull = sv.afull;      // [2] accumulator stage complete (one per CTA, multicast commit)
  uint64_t* const tempty = sv.aempty;    // [2] leader: 8 epilogue warps x 2 CTAs have read the stage
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);
  constexpr int AM = CF::kAuxMul;
  const int cpt = p.C / TILE_K;
  const int half = p.taps / 2;
  const int per_pass = p.tap_n > 0 ? p.tap_n : p.taps;
  const int span = p.tap_span > 0 ? p.tap_span : per_pass;
  const int n_pass = (span + per_pass - 1) / per_pass;
  const int cslabs = p.n_cond > 0 ? p.cond_slabs : 0;
  // This pair's units: the contiguous range [u_lo, u_hi) of u = item * n_pass + pass.  With a scratch buffer the cut may fall
  // INSIDE an item (n_items * n_pass units over the pairs instead of whole items: 160 items on 74 pairs are 3 rounds of whole
  // items but 19.5 -> 20 of 27 pass-units): the pair that starts mid-item sums its passes of that item into its own 256 x 256
  // scratch tile and lin_fixup_kernel adds that tile to out[] afterwards -- the launcher guarantees that a range is at least one
  // item long, so an item has at most one such partial and the result does not depend on timing.
  const int gran = p.lin_scratch ? 1 : n_pass;
  const int n_units = p.n_items * n_pass;
  const int u_lo = (int)((long long)pair_id * (n_units / gran) / n_pairs) * gran;
  const int u_hi = (int)((long long)(pair_id + 1) * (n_units / gran) / n_pairs) * gran;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.xh); tma_prefetch_desc(&p.wd_h); tma_prefetch_desc(&p.xl); tma_prefetch_desc(&p.wd_l);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < CF::kStages; ++i) { mbar_init(&sv.full[i], 2); mbar_init(&sv.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc_pair(sv.tmem_ptr, 512); tmem_relinquish_pair(); }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *sv.tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- producer (each CTA: its 128 frames and its half of the weight tile) ----------------
      int stage = 0; uint32_t phase = 0;
      for (int u = u_lo; u < u_hi; ++u) {
        const int item = u / n_pass, ps = u - item * n_pass;
        const int nblk = item % p.n_blocks;
        const int mt = (item / p.n_blocks) * 2 + (int)rank;
        const int nb = mt / p.tiles_t, t0 = (mt % p.tiles_t) * TILE_M;
        const int tap0 = p.tap_lo + ps * per_pass;
        const int ntap = (ps + 1) * per_pass <= span ? per_pass : span - ps * per_pass;
        const int conv_slabs = ntap * cpt;
        const int nslabs = conv_slabs + (ps == n_pass - 1 ? cslabs : 0);
        for (int sl = 0; sl < nslabs; ++sl) {
          mbar_wait(&sv.empty[stage], phase ^ 1);
          uint8_t* st = sv.stage0 + stage * CF::kStageBytes;
          uint64_t* fl = &sv.full[stage];
          const uint32_t fb = mapa_cluster(smem_u32(fl), 0);
          prod_expect<PAIR>(fl, fb, CF::kStageBytes);
          if (sl < conv_slabs) {
            const int tap = tap0 + sl / cpt, cc = sl % cpt;
            const int trow = t0 + (tap - half) * p.dil;
            load_a<PAIR>(st + CF::kAOff, &p.xh, fl, fb, cc * TILE_K, trow, nb);
            load_b<PAIR>(st + CF::kBOff, &p.wd_h, fl, fb, tap * p.C + cc * TILE_K, nblk * TILE_N, rank);
            load_a<PAIR>(st + CF::kAAuxOff, &p.xl, fl, fb, AM * cc * TILE_K, trow, nb);
            load_b<PAIR>(st + CF::kBAuxOff, &p.wd_l, fl, fb, AM * (tap * p.C + cc * TILE_K), nblk * TILE_N, rank);
          } else {
            const int cc = sl - conv_slabs;
            load_a<PAIR>(st + CF::kAOff, &p.sh, fl, fb, cc * TILE_K, t0, nb);
            load_b<PAIR>(st + CF::kBOff, &p.wc_h, fl, fb, cc * TILE_K, nblk * TILE_N, rank);
            load_a<PAIR>(st + CF::kAAuxOff, &p.sl, fl, fb, AM * cc * TILE_K, t0, nb);
            load_b<PAIR>(st + CF::kBAuxOff, &p.wc_l, fl, fb, AM * cc * TILE_K, nblk * TILE_N, rank);
          }
          if (++stage == CF::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      // ---------------- MMA issuer (leader CTA) ----------------
      int stage = 0; uint32_t phase = 0;
      int ucnt = 0;
      for (int u = u_lo; u < u_hi; ++u, ++ucnt) {
        const int ps = u % n_pass;
        const int ntap = (ps + 1) * per_pass <= span ? per_pass : span - ps * per_pass;
        const int nslabs = ntap * cpt + (ps == n_pass - 1 ? cslabs : 0);
        const int as = ucnt & 1;
        mbar_wait(&tempty[as], ((ucnt >> 1) & 1) ^ 1);   // both CTAs' epilogues have read this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)as * 256u;
        for (int sl = 0; sl < nslabs; ++sl) {
          mbar_wait(&sv.full[stage], phase);
          tc_fence_after();
          issue_slab<P, PAIR>(sv.stage0 + stage * CF::kStageBytes, tmem_d, sl == 0);
          umma_commit_pair(&sv.empty[stage]);
          if (++stage == CF::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull[as]);
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue (8 warps), overlapped with the next unit's MMAs ----------------
    const int q = warp & 3, grp = (warp - 4) >> 2;
    const uint32_t sbuf = smem_u32(sv.sbias) + (uint32_t)(warp - 4) * 4096u;   // this warp's 32 rows x 128 bytes
    const float inv_l = __ldg(p.inv_scale);
    const int l8 = lane & 7, lr = lane >> 3;
    int ucnt = 0;
    for (int u = u_lo; u < u_hi; ++u, ++ucnt) {
      const int item = u / n_pass, ps = u - item * n_pass;
      const int nblk = item % p.n_blocks;
      const int mt = (item / p.n_blocks) * 2 + (int)rank;
      const int nb = mt / p.tiles_t, t0 = (mt % p.tiles_t) * TILE_M;
      const int tq = t0 + q * 32;                                    // first frame of this warp's 32 rows
      const int colb = nblk * TILE_N + grp * 128 + l8 * 4;           // this lane's 4 columns of every 32-column block
      // an item whose first pass belongs to the previous pair: this pair's passes of it are summed in its scratch tile
      const bool partial = item * n_pass < u_lo;
      float* const dst = partial ? p.lin_scratch + ((size_t)(2 * pair_id + (int)rank) * TILE_M + q * 32) * TILE_N + grp * 128 + l8 * 4
                                 : p.lin_out + ((size_t)nb * p.T + tq) * (size_t)p.ldo + colb;
      const size_t ldd = partial ? (size_t)TILE_N : (size_t)p.ldo;
      const int as = ucnt & 1;
      const bool acc = partial ? u > u_lo : (ps > 0 || p.lin_acc != 0);
      const float* bias = (ps == 0 && !partial) ? p.bias_cond : nullptr;
      mbar_wait(&tfull[as], (ucnt >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * 256u + (uint32_t)grp * 128u;
#pragma unroll 1
      for (int c4 = 0; c4 < 4; ++c4) {
        float v[32];
        load_acc32<P>(taddr + c4 * 32, inv_l, v);
        if (c4 == 3) {   // this warp's last TMEM read of the stage: hand it back to the MMA issuer (leader CTA)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&tempty[as]), 0));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)   // thread = row: 8 x 16 bytes into the row, 16-byte units XOR-swizzled by the row
          sts128(sbuf + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        __syncwarp();
        // 8 lanes per row: instruction k moves rows 4k .. 4k+3 as four whole 128-byte lines.  All loads of the block before its
        // first store (a load queued behind a store to the same row was 9x slower in the one-tile kernel).
        float4 o4[8];
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) b4 = *reinterpret_cast<const float4*>(bias + colb + c4 * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int r = 4 * k + lr;
          o4[k] = b4;
          if (acc && tq + r < p.T) {
            const float4 qv = __ldcg(reinterpret_cast<const float4*>(dst + (size_t)r * ldd + c4 * 32));
            o4[k].x += qv.x; o4[k].y += qv.y; o4[k].z += qv.z; o4[k].w += qv.w;
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int r = 4 * k + lr;
          const float4 a = lds128(sbuf + (uint32_t)r * 128u + (uint32_t)((l8 ^ (r & 7)) << 4));
          if (tq + r < p.T)
            *reinterpret_cast<float4*>(dst + (size_t)r * ldd + c4 * 32) = make_float4(a.x + o4[k].x, a.y + o4[k].y, a.z + o4[k].z, a.w + o4[k].w);
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}


// out[] += the partial tile of every pair whose unit range starts inside an item (see umma_conv_lin_pers_kernel).  One block per CTA
// of the persistent grid; u_lo is recomputed exactly as there.
__global__ void __launch_bounds__(256) lin_fixup_kernel(const float* __restrict__ scratch, float* __restrict__ out, int ldo, int T, int tiles_t,
                                                        int n_blocks, int n_items, int n_pass, int n_pairs) {
  const int pair_id = (int)(blockIdx.x >> 1), rank = (int)(blockIdx.x & 1);
  const int n_units = n_items * n_pass;
  const int u_lo = (int)((long long)pair_id * n_units / n_pairs);
  const int item = u_lo / n_pass;
  if (item * n_pass == u_lo) return;                       // the range starts at an item boundary: nothing parked
  const int nblk = item % n_blocks, mt = (item / n_blocks) * 2 + rank;
  const int nb = mt / tiles_t, t0 = (mt % tiles_t) * TILE_M;
  const float4* src = reinterpret_cast<const float4*>(scratch + (size_t)(2 * pair_id + rank) * TILE_M * TILE_N);
  for (int i = threadIdx.x; i < TILE_M * TILE_N / 4; i += 256) {
    const int r = i / (TILE_N / 4), c = (i - r * (TILE_N / 4)) * 4;
    if (t0 + r >= T) continue;
    float4* d = reinterpret_cast<float4*>(out + ((size_t)nb * T + t0 + r) * (size_t)ldo + nblk * TILE_N + c);
    const float4 a = src[i];
    float4 o = *d;
    o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    *d = o;
  }
}

// ---------------------------------------------------------------------------------------------
// gate kernel, window variant (CTA pairs, aux operands).  The 9 dilated taps of one 64-channel chunk read the SAME
// activation rows shifted by tap*dil frames, so the window [t0 - 4*dil, t0 + 128 + 4*dil) is fetched ONCE per chunk and
// every tap's A descriptor just starts tap*dil rows (x 128 B) further into it.  Shared-memory fill traffic per K-slab
// drops from 64 KB to ~37 KB (the weight half-tile dominates), which matters because MMA operand reads + TMA fills
// run right at the 128 B/clk shared-memory limit otherwise.
//   ring: 2 window buffers x (24 KB main + 24 KB aux)  |  3 weight stages x (16 KB main + 16 KB aux)
// ---------------------------------------------------------------------------------------------
constexpr int WIN_BUF_BYTES = 49152;   // 192 rows x 128 B, main + aux
constexpr int WIN_HALF = 24576;
constexpr int WIN_BSTAGE = 32768;
constexpr int WIN_BSTAGES = 3;

template <int P>
__global__ void __launch_bounds__(NUM_THREADS, 1) umma_gate_win_kernel(const __grid_constant__ GateParams p) {
  static_assert(P != 0, "window kernel needs aux operands");
  constexpr bool PAIR = true;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmemView sv = carve<P, PAIR>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  int cid = (int)(blockIdx.x >> 1);
  const int nblk = cid % p.n_blocks; cid /= p.n_blocks;
  const int mt = cid * 2 + (int)rank;
  const int tt = mt % p.tiles_t;
  const int nb = mt / p.tiles_t;
  const int t0 = tt * TILE_M;
  const int cpt = p.C / TILE_K;                       // conv chunks (64 channels each)
  const int nb_first = (cid * 2) / p.tiles_t;
  const int nchunks = cpt + (nb_first < p.n_cond ? p.cond_slabs : 0);   // conditioner chunks have a single slab
  const int half = p.taps / 2;
  constexpr int AM = P >= 2 ? 2 : 1;
  uint8_t* const wbuf = sv.stage0;                                   // 2 window buffers
  uint8_t* const bring = sv.stage0 + 2 * WIN_BUF_BYTES;              // 3 weight stages

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.xwh); tma_prefetch_desc(&p.xwl); tma_prefetch_desc(&p.wd_h); tma_prefetch_desc(&p.wd_l);
    tma_prefetch_desc(&p.zh); tma_prefetch_desc(&p.zl);
  }
  if (warp == 3) {
    const float* bsrc = (nb < p.n_cond ? p.bias_cond : p.bias_unc) + nblk * TILE_N;
    for (int i = lane; i < TILE_N; i += 32) sv.sbias[i] = __ldg(bsrc + i);
  }
  prologue<P, PAIR>(sv, warp);
  const uint32_t tmem_base = *sv.tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int a_fill = 0;                 // chunks whose window load has been issued
      auto issue_window = [&](int c) {
        const int ai = c & 1;
        mbar_wait(&sv.aempty[ai], ((c >> 1) & 1) ^ 1);
        const uint32_t fb = mapa_cluster(smem_u32(&sv.afull[ai]), 0);
        uint8_t* dst = wbuf + ai * WIN_BUF_BYTES;
        if (c < cpt) {
          mbar_expect_tx_cluster(fb, 2u * (uint32_t)p.win_rows * 128u);
          tma_load_3d_pair(dst, &p.xwh, fb, c * TILE_K, t0 - half * p.dil, nb);
          tma_load_3d_pair(dst + WIN_HALF, &p.xwl, fb, AM * c * TILE_K, t0 - half * p.dil, nb);
        } else {
          mbar_expect_tx_cluster(fb, 2u * A_TILE_BYTES);
          tma_load_3d_pair(dst, &p.sh, fb, (c - cpt) * TILE_K, t0, nb);
          tma_load_3d_pair(dst + WIN_HALF, &p.sl, fb, AM * (c - cpt) * TILE_K, t0, nb);
        }
        a_fill = c + 1;
      };
      issue_window(0);
      int bs = 0; uint32_t bphase = 0;
      for (int c = 0; c < nchunks; ++c) {
        const int nsl = c < cpt ? p.taps : 1;
        for (int j = 0; j < nsl; ++j) {
          // prefetch the next chunk's window once the MMAs are safely past the previous user of that buffer
          if (a_fill == c + 1 && c + 1 < nchunks && j >= (nsl > 4 ? 4 : nsl - 1)) issue_window(c + 1);
          mbar_wait(&sv.empty[bs], bphase ^ 1);
          const uint32_t fb = mapa_cluster(smem_u32(&sv.full[bs]), 0);
          uint8_t* dst = bring + bs * WIN_BSTAGE;
          mbar_expect_tx_cluster(fb, WIN_BSTAGE);
          const int row0 = nblk * TILE_N + (int)rank * (TILE_N / 2);
          if (c < cpt) {
            const int col = j * p.C + c * TILE_K;
            tma_load_2d_pair(dst, &p.wd_h, fb, col, row0);
            tma_load_2d_pair(dst + WIN_BSTAGE / 2, &p.wd_l, fb, AM * col, row0);
          } else {
            const int col = (c - cpt) * TILE_K;
            tma_load_2d_pair(dst, &p.wc_h, fb, col, row0);
            tma_load_2d_pair(dst + WIN_BSTAGE / 2, &p.wc_l, fb, AM * col, row0);
          }
          if (++bs == WIN_BSTAGES) { bs = 0; bphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = P >= 2 ? make_idesc_fmt0(2 * TILE_M, TILE_N) : make_idesc_bf16(2 * TILE_M, TILE_N);
      constexpr uint32_t idesc_e5 = make_idesc_bf16(2 * TILE_M, TILE_N);
      int bs = 0; uint32_t bphase = 0;
      bool first = true;
      for (int c = 0; c < nchunks; ++c) {
        const int ai = c & 1;
        mbar_wait(&sv.afull[ai], (c >> 1) & 1);
        const int nsl = c < cpt ? p.taps : 1;
        const uint32_t a_main = smem_u32(wbuf + ai * WIN_BUF_BYTES), a_aux = a_main + WIN_HALF;
        for (int j = 0; j < nsl; ++j) {
          mbar_wait(&sv.full[bs], bphase);
          tc_fence_after();
          const uint32_t shift = c < cpt ? (uint32_t)(j * p.dil) * 128u : 0u;   // tap j starts j*dil frames into the window
          const uint32_t b_main = smem_u32(bring + bs * WIN_BSTAGE), b_aux = b_main + WIN_BSTAGE / 2;
#pragma unroll
          for (int k = 0; k < TILE_K / UMMA_K; ++k) {
            const uint32_t ko = k * UMMA_K * 2;
            const uint64_t da = make_sw128_desc(a_main + shift + ko), db = make_sw128_desc(b_main + ko);
            const uint32_t acc = (first && k == 0) ? 0u : 1u;
            umma_bf16_pair(tmem_base, da, db, idesc, acc);
            if (P == 1) {
              umma_bf16_pair(tmem_base, make_sw128_desc(a_aux + shift + ko), db, idesc, 1u);
              umma_bf16_pair(tmem_base, da, make_sw128_desc(b_aux + ko), idesc, 1u);
            }
            if (P == 2) umma_f8_pair(tmem_base + 256, make_sw128_desc(a_aux + shift + ko), make_sw128_desc(b_aux + ko), idesc, acc);
            if (P == 3) umma_f8_pair(tmem_base, make_sw128_desc(a_aux + shift + ko), make_sw128_desc(b_aux + ko), idesc_e5, 1u);
          }
          first = false;
          umma_commit_pair(&sv.empty[bs]);
          if (++bs == WIN_BSTAGES) { bs = 0; bphase ^= 1; }
        }
        umma_commit_pair(&sv.aempty[ai]);   // window buffer free (in both CTAs) once this chunk's MMAs retire
      }
      umma_commit_pair(sv.tmem_full);
    }
  } else if (warp >= 4) {
    gate_epilogue<P>(sv, p, tmem_base, warp, lane, nb, nblk, t0);
  }
  teardown<P, PAIR>(tmem_base, warp);
}


// ---------------------------------------------------------------------------------------------
// gate kernel, persistent window variant (CTA pairs; precisions with ONE 256-column accumulator: bf16x3, f16e5).
// One CTA pair per SM pair loops over its tiles; TMEM holds two accumulator stages, so the epilogue of tile i (TMEM
// reads, gate math, operand split, TMA store) overlaps the MMAs of tile i+1, and barrier setup / TMEM allocation / the
// first TMA round trip are paid once per launch instead of once per tile.
//   smem: 2 window buffers (96 KB) | 3 weight stages (96 KB) | 32 KB epilogue staging (main boxes, then aux boxes)
// ---------------------------------------------------------------------------------------------
constexpr int PW_RING = 2 * WIN_BUF_BYTES + WIN_BSTAGES * WIN_BSTAGE;   // 192 KB
constexpr int PW_STAGING = 2 * CHUNK_BYTES;                             // 32 KB
constexpr int PW_SMEM = PW_RING + PW_STAGING + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*bias*/;

template <int P>
__device__ __forceinline__ void pack16(const float (&v)[16], uint32_t (&m)[8], uint32_t (&a)[8]) {
  if (P >= 2) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float t[4] = {v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]};
      if (P == 2) split_f16f8_x4(t, m[2 * q], m[2 * q + 1], a[q], a[4 + q]);
      else split_f16e5_x4(t, m[2 * q], m[2 * q + 1], a[q], a[4 + q]);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) split_pack2(v[2 * e], v[2 * e + 1], m[e], a[e]);
  }
}

struct GateTile { int nblk, nb, t0, nchunks; };

template <int P, bool DUAL>
__global__ void __launch_bounds__(NUM_THREADS, 1) umma_gate_pers_kernel(const __grid_constant__ GateParams p) {
  static_assert(P == 1 || P == 3, "needs a single 256-column accumulator");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);
  constexpr int AM = P >= 2 ? 2 : 1;

  uint8_t* ring;
  { uint32_t a = smem_u32(smem_raw); ring = smem_raw + (((a + 1023u) & ~1023u) - a); }
  uint8_t* const wbuf = ring;
  uint8_t* const bring = ring + 2 * WIN_BUF_BYTES;
  uint8_t* const staging = ring + PW_RING;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(ring + PW_RING + PW_STAGING);
  uint64_t* const bfull = bars;            // [3]
  uint64_t* const bempty = bars + 3;       // [3]
  uint64_t* const afull = bars + 6;        // [2]
  uint64_t* const aempty = bars + 8;       // [2]
  uint64_t* const tfull = bars + 10;       // [2]
  uint64_t* const tempty = bars + 12;      // [2]
  uint32_t* const tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);
  float* const sbias = reinterpret_cast<float*>(ring + PW_RING + PW_STAGING + 256);

  const int cpt = p.C / TILE_K;
  const int half = p.taps / 2;
  auto tile_of = [&](int item) -> GateTile {
    GateTile t;
    t.nblk = item % p.n_blocks;
    const int pm = item / p.n_blocks;
    const int mt = pm * 2 + (int)rank;
    t.nb = mt / p.tiles_t;
    t.t0 = (mt % p.tiles_t) * TILE_M;
    const int nb_first = (pm * 2) / p.tiles_t;
    t.nchunks = cpt + (nb_first < p.n_cond_mma ? p.cond_slabs : 0);
    return t;
  };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.xwh); tma_prefetch_desc(&p.xwl); tma_prefetch_desc(&p.wd_h); tma_prefetch_desc(&p.wd_l);
    tma_prefetch_desc(&p.zh); tma_prefetch_desc(&p.zl);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < 3; ++i) { mbar_init(&bfull[i], 2); mbar_init(&bempty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&afull[i], 2); mbar_init(&aempty[i], 1);
      mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16);   // 8 epilogue warps x 2 CTAs drain an accumulator stage
    }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc_pair(tmem_ptr, 512); tmem_relinquish_pair(); }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- producer: windows and weight stages for every tile of this pair, in order ----------------
      int a_issued = 0;            // windows issued so far (global over tiles)
      int w_item = pair_id, w_c = 0;                // next window to issue
      GateTile wt = tile_of(w_item < p.n_items ? w_item : 0);
      auto issue_window = [&]() {
        const int ai = a_issued & 1;
        mbar_wait(&aempty[ai], ((a_issued >> 1) & 1) ^ 1);
        const uint32_t fb = mapa_cluster(smem_u32(&afull[ai]), 0);
        uint8_t* dst = wbuf + ai * WIN_BUF_BYTES;
        if (w_c < cpt) {
          mbar_expect_tx_cluster(fb, 2u * (uint32_t)p.win_rows * 128u);
          tma_load_3d_pair(dst, &p.xwh, fb, w_c * TILE_K, wt.t0 - half * p.dil, wt.nb);
          tma_load_3d_pair(dst + WIN_HALF, &p.xwl, fb, AM * w_c * TILE_K, wt.t0 - half * p.dil, wt.nb);
        } else {
          mbar_expect_tx_cluster(fb, 2u * A_TILE_BYTES);
          tma_load_3d_pair(dst, &p.sh, fb, (w_c - cpt) * TILE_K, wt.t0, wt.nb);
          tma_load_3d_pair(dst + WIN_HALF, &p.sl, fb, AM * (w_c - cpt) * TILE_K, wt.t0, wt.nb);
        }
        ++a_issued;
        if (++w_c == wt.nchunks) { w_c = 0; w_item += n_pairs; if (w_item < p.n_items) wt = tile_of(w_item); }
      };
      int gidx = 0;                // global index of the chunk whose weight slabs are being issued
      int bcnt = 0;
      if (w_item < p.n_items) issue_window();
      for (int item = pair_id; item < p.n_items; item += n_pairs) {
        const GateTile ti = tile_of(item);
        const int row0 = ti.nblk * TILE_N + (int)rank * (TILE_N / 2);
        for (int c = 0; c < ti.nchunks; ++c, ++gidx) {
          const int nsl = c < cpt ? p.taps : 1;
          for (int j = 0; j < nsl; ++j) {
            if (a_issued == gidx + 1 && w_item < p.n_items && j >= (nsl > 4 ? 4 : nsl - 1)) issue_window();
            const int bs = bcnt % WIN_BSTAGES;
            mbar_wait(&bempty[bs], ((bcnt / WIN_BSTAGES) & 1) ^ 1);
            const uint32_t fb = mapa_cluster(smem_u32(&bfull[bs]), 0);
            uint8_t* dst = bring + bs * WIN_BSTAGE;
            mbar_expect_tx_cluster(fb, WIN_BSTAGE);
            if (c < cpt) {
              const int col = j * p.C + c * TILE_K;
              tma_load_2d_pair(dst, &p.wd_h, fb, col, row0);
              tma_load_2d_pair(dst + WIN_BSTAGE / 2, &p.wd_l, fb, AM * col, row0);
            } else {
              const int col = (c - cpt) * TILE_K;
              tma_load_2d_pair(dst, &p.wc_h, fb, col, row0);
              tma_load_2d_pair(dst + WIN_BSTAGE / 2, &p.wc_l, fb, AM * col, row0);
            }
            ++bcnt;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      // ---------------- MMA issuer (leader CTA) ----------------
      constexpr uint32_t idesc = P >= 2 ? make_idesc_fmt0(2 * TILE_M, TILE_N) : make_idesc_bf16(2 * TILE_M, TILE_N);
      constexpr uint32_t idesc_e5 = make_idesc_bf16(2 * TILE_M, TILE_N);
      int gidx = 0, bcnt = 0, tcnt = 0;
      for (int item = pair_id; item < p.n_items; item += n_pairs, ++tcnt) {
        const GateTile ti = tile_of(item);
        const int as = tcnt & 1;
        mbar_wait(&tempty[as], ((tcnt >> 1) & 1) ^ 1);   // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)as * 256u;
        bool first = true;
        for (int c = 0; c < ti.nchunks; ++c, ++gidx) {
          const int ai = gidx & 1;
          mbar_wait(&afull[ai], (gidx >> 1) & 1);
          const int nsl = c < cpt ? p.taps : 1;
          const uint32_t a_main = smem_u32(wbuf + ai * WIN_BUF_BYTES), a_aux = a_main + WIN_HALF;
          for (int j = 0; j < nsl; ++j, ++bcnt) {
            const int bs = bcnt % WIN_BSTAGES;
            mbar_wait(&bfull[bs], (bcnt / WIN_BSTAGES) & 1);
            tc_fence_after();
            const uint32_t shift = c < cpt ? (uint32_t)(j * p.dil) * 128u : 0u;
            const uint32_t b_main = smem_u32(bring + bs * WIN_BSTAGE), b_aux = b_main + WIN_BSTAGE / 2;
#pragma unroll
            for (int k = 0; k < TILE_K / UMMA_K; ++k) {
              const uint32_t ko = k * UMMA_K * 2;
              const uint64_t da = make_sw128_desc(a_main + shift + ko), db = make_sw128_desc(b_main + ko);
              umma_bf16_pair(tmem_d, da, db, idesc, (first && k == 0) ? 0u : 1u);
              if (P == 1) {
                umma_bf16_pair(tmem_d, make_sw128_desc(a_aux + shift + ko), db, idesc, 1u);
                umma_bf16_pair(tmem_d, da, make_sw128_desc(b_aux + ko), idesc, 1u);
              }
              if (P == 3) umma_f8_pair(tmem_d, make_sw128_desc(a_aux + shift + ko), make_sw128_desc(b_aux + ko), idesc_e5, 1u);
            }
            first = false;
            umma_commit_pair(&bempty[bs]);
          }
          umma_commit_pair(&aempty[ai]);
        }
        umma_commit_pair(&tfull[as]);
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue (8 warps), overlapped with the next tile's MMAs ----------------
    const int q = warp & 3, grp = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int etid = (int)threadIdx.x - 128;
    const bool issuer = (warp == 4) && (lane == 0);
    const uint32_t stg = smem_u32(staging);
    const float inv = (P >= 2) ? __ldg(p.inv_scale) : 0.f;
    int tcnt = 0;
    for (int item = pair_id; item < p.n_items; item += n_pairs, ++tcnt) {
      const GateTile ti = tile_of(item);
      const int as = tcnt & 1;
#pragma unroll 1
      for (int pass = 0; pass < (DUAL ? 2 : 1); ++pass) {
      // DUAL: pass 0 = unconditional output (roll nb + dual_off), pass 1 = conditional output (roll nb)
      const bool is_cond = DUAL ? (pass == 1) : (ti.nb < p.n_cond);
      const int zroll = p.z_group0 + ti.nb + ((DUAL && pass == 0) ? p.dual_off : 0);
      // previous tile: every thread is done with sbias, and the aux stores have finished reading the staging area
      if (issuer) tma_store_wait_read<0>();
      named_bar_sync(EPI_BAR, EPI_THREADS);
      sbias[etid] = __ldg((is_cond ? p.bias_cond : p.bias_unc) + ti.nblk * TILE_N + etid);
      named_bar_sync(EPI_BAR, EPI_THREADS);
      if (pass == 0) mbar_wait(&tfull[as], (tcnt >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * 256u;
      uint32_t zm[2][2][8], za[2][2][8];   // [c2][hf]: packed main / aux words of 16 channels
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int ch = grp * 2 + c2;
        float g[32], f[32];
        load_acc32<P>(taddr + ch * 32, inv, g);
        load_acc32<P>(taddr + 128 + ch * 32, inv, f);
        if (is_cond && p.cond != nullptr) {   // + conditioner_projection(spec) of this frame, fp32, computed once per clip
          const int tf = ti.t0 + row;
          if (tf < p.T) {
            const float* cg = p.cond + ((size_t)ti.nb * p.T + tf) * (size_t)(2 * p.C) + ti.nblk * (TILE_N / 2) + ch * 32;
            const float* cf = cg + p.C;
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(cg) + v);
              const float4 b = __ldg(reinterpret_cast<const float4*>(cf) + v);
              g[4 * v] += a.x; g[4 * v + 1] += a.y; g[4 * v + 2] += a.z; g[4 * v + 3] += a.w;
              f[4 * v] += b.x; f[4 * v + 1] += b.y; f[4 * v + 2] += b.z; f[4 * v + 3] += b.w;
            }
          }
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float z[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int j = hf * 16 + i;
            z[i] = gate_act(g[j] + sbias[ch * 32 + j], f[j] + sbias[128 + ch * 32 + j]);
          }
          pack16<P>(z, zm[c2][hf], za[c2][hf]);
        }
      }
      // this warp's TMEM reads are complete: hand the accumulator stage back to the MMA issuer (leader CTA)
      tc_fence_before();
      __syncwarp();
      if ((!DUAL || pass == 1) && lane == 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&tempty[as]), 0));
      // pass A: main boxes (group g owns box g)
      const uint32_t box = stg + grp * CHUNK_BYTES;
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int chn = c2 * 32 + hf * 16;
          sts128u(box + sw128_off(row, chn / 8), zm[c2][hf][0], zm[c2][hf][1], zm[c2][hf][2], zm[c2][hf][3]);
          sts128u(box + sw128_off(row, chn / 8 + 1), zm[c2][hf][4], zm[c2][hf][5], zm[c2][hf][6], zm[c2][hf][7]);
        }
      fence_proxy_async();
      named_bar_sync(EPI_BAR, EPI_THREADS);
      const int c0 = ti.nblk * (TILE_N / 2);
      if (issuer) {
        tma_store_3d(&p.zh, staging, c0, ti.t0, zroll);
        tma_store_3d(&p.zh, staging + CHUNK_BYTES, c0 + 64, ti.t0, zroll);
        tma_store_commit();
        tma_store_wait_read<0>();
      }
      named_bar_sync(EPI_BAR, EPI_THREADS);
      // pass B: aux boxes
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int chn = c2 * 32 + hf * 16;
          if (P >= 2) {
            sts128u(box + sw128_off(row, chn / 16), za[c2][hf][0], za[c2][hf][1], za[c2][hf][2], za[c2][hf][3]);
            sts128u(box + sw128_off(row, 4 + chn / 16), za[c2][hf][4], za[c2][hf][5], za[c2][hf][6], za[c2][hf][7]);
          } else {
            sts128u(box + sw128_off(row, chn / 8), za[c2][hf][0], za[c2][hf][1], za[c2][hf][2], za[c2][hf][3]);
            sts128u(box + sw128_off(row, chn / 8 + 1), za[c2][hf][4], za[c2][hf][5], za[c2][hf][6], za[c2][hf][7]);
          }
        }
      fence_proxy_async();
      named_bar_sync(EPI_BAR, EPI_THREADS);
      if (issuer) {
        tma_store_3d(&p.zl, staging, AM * c0, ti.t0, zroll);
        tma_store_3d(&p.zl, staging + CHUNK_BYTES, AM * (c0 + 64), ti.t0, zroll);
        tma_store_commit();
      }
      }   // pass
    }
    if (issuer) tma_store_wait_read<0>();
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------
// gate kernel, f16n4 variant: fp16 main product + ONE block-scaled fp4 correction product (1.5 MMA units per K-step
// instead of 2).  Same persistent CTA pairs, tap window and epilogue as umma_gate_pers_kernel, plus what kind::mxf4nvf4
// needs (layouts verified by profiles/experiments/nv4_probe.cu on a B200):
//   * aux operands are 64-byte rows of e2m1 codes [part 0 | part 1] per 64-channel chunk, 64-byte swizzle; a tap is still a
//     row-shifted descriptor into the window;
//   * scale factors live in TMEM.  The weight side's 2 KB of atoms per (N block, K-slab) are precomputed and arrive with the
//     weight stage.  The activation side's depend on the tap's row shift: an atom holds, at row l, the scales of window rows
//     s + l, s + 32 + l, s + 64 + l, s + 96 + l (s = tap * dilation).  Warp 3 of EACH CTA therefore writes, once per WINDOW,
//     an image whose row R is {sf[R], sf[R + 32], sf[R + 64], sf[R + 96]} (16 bytes, all four from the lane's own prefetched
//     registers): the atom of any tap is then the 512 contiguous bytes starting at image row s, and the MMA thread copies
//     it (and the weight atoms) to TMEM with tcgen05.cp right before the slab's MMAs (two alternating 24-column sets).
//     (First version: the warp gathered two atoms PER K-SLAB into a 4-slot ring; ncu showed it spending 71 % of its time
//     in the membar of the release.cluster arrive, one per slab, which held the tensor pipe at 48 %.)
//   * the scale factors take TMEM columns, so there is ONE accumulator stage: the epilogue warps (216 registers each, taken
//     from the producer warps with setmaxnreg) drain the whole accumulator into registers and hand it back at once; the
//     next tile's MMAs wait ~1 us per tile instead of overlapping the whole epilogue.
//   * FOUR weight stages: a K-slab is only 768 MMA cycles, so the 3-stage ring of the f16e5 kernel (3072 cycles of cover
//     there) would cover 2304 cycles of TMA round trip; the staging area shrinks to two boxes (main boxes, then aux
//     boxes from registers) to make room.
//   smem: 2 windows x (24 + 12 KB) | 4 weight stages x (16 + 8 + 2 KB) | 2 SF images x 3 KB | 32 KB staging
// ---------------------------------------------------------------------------------------------
constexpr int N4_WIN_MAIN = 24576, N4_WIN_AUX = 12288, N4_WIN_BUF = N4_WIN_MAIN + N4_WIN_AUX;
constexpr int N4_B_MAIN = 16384, N4_B_AUX = 8192, N4_B_SF = 2048, N4_BSTAGE = N4_B_MAIN + N4_B_AUX + N4_B_SF;
constexpr int N4_BSTAGES = 4;
constexpr int N4_SFIMG_PART = 96 * 16, N4_SFIMG_BUF = 2 * N4_SFIMG_PART;   // image rows 0 .. win_rows - 97 (<= 96), lo part then hi part
constexpr int N4_OFF_BRING = 2 * N4_WIN_BUF;
constexpr int N4_OFF_SFIMG = N4_OFF_BRING + N4_BSTAGES * N4_BSTAGE;
constexpr int N4_OFF_STAGING = (N4_OFF_SFIMG + 2 * N4_SFIMG_BUF + 1023) / 1024 * 1024;
constexpr int N4_OFF_BARS = N4_OFF_STAGING + 2 * CHUNK_BYTES;
constexpr int N4_SMEM = N4_OFF_BARS + 256 /*barriers*/ + 1024 /*bias*/ + 1024 /*align*/;
static_assert(N4_SMEM <= 232448, "f16n4 gate kernel: shared memory over the 227 KB limit");

template <bool DUAL>
__global__ void __launch_bounds__(NUM_THREADS, 1) umma_gate_n4_kernel(const __grid_constant__ GateParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);

  uint8_t* ring;
  { uint32_t a = smem_u32(smem_raw); ring = smem_raw + (((a + 1023u) & ~1023u) - a); }
  uint8_t* const wbuf = ring;
  uint8_t* const bring = ring + N4_OFF_BRING;
  uint8_t* const sfimg = ring + N4_OFF_SFIMG;
  uint8_t* const staging = ring + N4_OFF_STAGING;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(ring + N4_OFF_BARS);
  uint64_t* const bfull = bars;            // [4]
  uint64_t* const bempty = bars + 4;       // [4]
  uint64_t* const afull = bars + 8;        // [2]
  uint64_t* const aempty = bars + 10;      // [2]
  uint64_t* const tfull = bars + 12;
  uint64_t* const tempty = bars + 13;
  uint32_t* const tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);
  float* const sbias = reinterpret_cast<float*>(ring + N4_OFF_BARS + 256);

  const int cpt = p.C / TILE_K;
  const int half = p.taps / 2;
  auto tile_of = [&](int item) -> GateTile {
    GateTile t;
    t.nblk = item % p.n_blocks;
    const int pm = item / p.n_blocks;
    const int mt = pm * 2 + (int)rank;
    t.nb = mt / p.tiles_t;
    t.t0 = (mt % p.tiles_t) * TILE_M;
    t.nchunks = cpt;
    return t;
  };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.xwh); tma_prefetch_desc(&p.xw4); tma_prefetch_desc(&p.wd_h); tma_prefetch_desc(&p.wd4);
    tma_prefetch_desc(&p.wsf); tma_prefetch_desc(&p.zh); tma_prefetch_desc(&p.zl);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < N4_BSTAGES; ++i) { mbar_init(&bfull[i], 2); mbar_init(&bempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&afull[i], 4); mbar_init(&aempty[i], 1); }   // full: 2 producers (TMA bytes) + 2 SF warps
    mbar_init(tfull, 1); mbar_init(tempty, 16);    // 8 epilogue warps x 2 CTAs drain the accumulator
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc_pair(tmem_ptr, 512); tmem_relinquish_pair(); }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr;
  // register split inside the CTA's own pool (384 x 168 registers at launch): 128 x 64 + 256 x 216 = 63488 <= 64512.
  // (setmaxnreg.inc waits for registers freed by THIS CTA's dec: a split summing above the launch allocation hangs.)
  if (warp < 4) {
  setmaxnreg_dec<64>();
  if (warp == 0) {
    if (elect_one()) {
      // ---------------- producer: windows (main + aux) and weight stages (main + aux + scale atoms) ----------------
      int a_issued = 0;
      int w_item = pair_id, w_c = 0;
      GateTile wt = tile_of(w_item < p.n_items ? w_item : 0);
      auto issue_window = [&]() {
        const int ai = a_issued & 1;
        mbar_wait(&aempty[ai], ((a_issued >> 1) & 1) ^ 1);
        const uint32_t fb = mapa_cluster(smem_u32(&afull[ai]), 0);
        uint8_t* dst = wbuf + ai * N4_WIN_BUF;
        mbar_expect_tx_cluster(fb, (uint32_t)p.win_rows * 192u);
        tma_load_3d_pair(dst, &p.xwh, fb, w_c * TILE_K, wt.t0 - half * p.dil, wt.nb);
        tma_load_3d_pair(dst + N4_WIN_MAIN, &p.xw4, fb, w_c * TILE_K, wt.t0 - half * p.dil, wt.nb);
        ++a_issued;
        if (++w_c == cpt) { w_c = 0; w_item += n_pairs; if (w_item < p.n_items) wt = tile_of(w_item); }
      };
      int gidx = 0, bcnt = 0;
      if (w_item < p.n_items) issue_window();
      for (int item = pair_id; item < p.n_items; item += n_pairs) {
        const GateTile ti = tile_of(item);
        const int row0 = ti.nblk * TILE_N + (int)rank * (TILE_N / 2);
        for (int c = 0; c < cpt; ++c, ++gidx) {
          for (int j = 0; j < p.taps; ++j) {
            if (a_issued == gidx + 1 && w_item < p.n_items && j >= (p.taps > 4 ? 4 : p.taps - 1)) issue_window();
            const int bs = bcnt % N4_BSTAGES;
            mbar_wait(&bempty[bs], ((bcnt / N4_BSTAGES) & 1) ^ 1);
            const uint32_t fb = mapa_cluster(smem_u32(&bfull[bs]), 0);
            uint8_t* dst = bring + bs * N4_BSTAGE;
            mbar_expect_tx_cluster(fb, N4_BSTAGE);
            const int slab = j * cpt + c;
            tma_load_2d_pair(dst, &p.wd_h, fb, slab * TILE_K, row0);
            tma_load_2d_pair(dst + N4_B_MAIN, &p.wd4, fb, slab * TILE_K, row0);
            tma_load_2d_pair(dst + N4_B_MAIN + N4_B_AUX, &p.wsf, fb, 0, (ti.nblk * p.taps * cpt + slab) * 16);
            ++bcnt;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      // ---------------- MMA issuer (leader CTA) ----------------
      constexpr uint32_t idesc = make_idesc_fmt0(2 * TILE_M, TILE_N);
      constexpr uint32_t idesc4 = make_idesc_nv4(2 * TILE_M, TILE_N);
      int gidx = 0, bcnt = 0, tcnt = 0;
      for (int item = pair_id; item < p.n_items; item += n_pairs, ++tcnt) {
        mbar_wait(tempty, (tcnt & 1) ^ 1);     // both CTAs' epilogues hold the previous tile's accumulator in registers
        tc_fence_after();
        bool first = true;
        for (int c = 0; c < cpt; ++c, ++gidx) {
          const int ai = gidx & 1;
          mbar_wait(&afull[ai], (gidx >> 1) & 1);
          const uint32_t a_main = smem_u32(wbuf + ai * N4_WIN_BUF), a_aux = a_main + N4_WIN_MAIN;
          for (int j = 0; j < p.taps; ++j, ++bcnt) {
            const int bs = bcnt % N4_BSTAGES;
            mbar_wait(&bfull[bs], (bcnt / N4_BSTAGES) & 1);
            tc_fence_after();
            const uint32_t b_main = smem_u32(bring + bs * N4_BSTAGE), b_aux = b_main + N4_B_MAIN, b_sf = b_aux + N4_B_AUX;
            const uint32_t sfc = tmem_base + 256u + (uint32_t)(bcnt & 1) * 24u;   // SFA 2 x 4 columns, SFB 2 x 8 columns
            const uint32_t rshift = (uint32_t)(j * p.dil);
            const uint32_t a_sf = smem_u32(sfimg + ai * N4_SFIMG_BUF) + rshift * 16u;   // this tap's atoms start at image row j * dil
            utccp_sf_pair(sfc, make_sfatom_desc(a_sf));
            utccp_sf_pair(sfc + 4, make_sfatom_desc(a_sf + N4_SFIMG_PART));
#pragma unroll
            for (int i = 0; i < 4; ++i) utccp_sf_pair(sfc + 8 + 4 * i, make_sfatom_desc(b_sf + 512 * i));
#pragma unroll
            for (int k = 0; k < TILE_K / UMMA_K; ++k) {
              const uint32_t ko = k * UMMA_K * 2;
              umma_bf16_pair(tmem_base, make_sw128_desc(a_main + rshift * 128u + ko), make_sw128_desc(b_main + ko), idesc,
                             (first && k == 0) ? 0u : 1u);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_nv4_pair(tmem_base, make_sw64_desc(a_aux + rshift * 64u + 32u * k), make_sw64_desc(b_aux + 32u * k), idesc4, 1u,
                            sfc + 4 * k, sfc + 8 + 8 * k);
            first = false;
            umma_commit_pair(&bempty[bs]);
          }
          umma_commit_pair(&aempty[ai]);
        }
        umma_commit_pair(tfull);
      }
    }
  } else if (warp == 3) {
    // ---------------- scale-factor warp (both CTAs): per-row window scales -> one shift-addressable image per window ----------------
    const int nch = p.C / TILE_K;
    auto load_win = [&](const GateTile& t, int c, uint2 (&v)[6]) {
      const int tw0 = t.t0 - half * p.dil;
      const uint2* src = reinterpret_cast<const uint2*>(p.xs) + ((size_t)t.nb * nch + c) * p.T;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int wr = lane + 32 * i, tt = tw0 + wr;
        v[i] = (wr < p.win_rows && tt >= 0 && tt < p.T) ? __ldg(src + tt) : make_uint2(0u, 0u);
      }
    };
    uint2 cur[6];
    int wcnt = 0;
    if (pair_id < p.n_items) { const GateTile t0 = tile_of(pair_id); load_win(t0, 0, cur); }
    for (int item = pair_id; item < p.n_items; item += n_pairs) {
      const GateTile ti = tile_of(item);
      for (int c = 0; c < cpt; ++c, ++wcnt) {
        const int ai = wcnt & 1;
        mbar_wait(&aempty[ai], ((wcnt >> 1) & 1) ^ 1);          // the MMAs (and scale copies) of the window two back have retired
        const uint32_t img = smem_u32(sfimg + ai * N4_SFIMG_BUF) + lane * 16;
#pragma unroll
        for (int m = 0; m < 3; ++m) {                            // image rows lane, lane + 32, lane + 64
          sts128u(img + m * 512, cur[m].x, cur[m + 1].x, cur[m + 2].x, cur[m + 3].x);                   // lo part (instruction 0)
          sts128u(img + N4_SFIMG_PART + m * 512, cur[m].y, cur[m + 1].y, cur[m + 2].y, cur[m + 3].y);   // hi part (instruction 1)
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&afull[ai]), 0));
        // next window's scales: issued after the arrive (its release fence would wait for them), needed one chunk from now
        if (c + 1 < cpt) load_win(ti, c + 1, cur);
        else if (item + n_pairs < p.n_items) { const GateTile tn = tile_of(item + n_pairs); load_win(tn, 0, cur); }
      }
    }
  }
  } else {
    setmaxnreg_inc<216>();
    // ---------------- epilogue (8 warps): drain the accumulator into registers, release it, then gate / split / store ----------------
    const int q = warp & 3, grp = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int etid = (int)threadIdx.x - 128;
    const bool issuer = (warp == 4) && (lane == 0);
    const uint32_t stg = smem_u32(staging);
    const float inv = __ldg(p.inv_scale);
    int tcnt = 0;
    for (int item = pair_id; item < p.n_items; item += n_pairs, ++tcnt) {
      const GateTile ti = tile_of(item);
      mbar_wait(tfull, tcnt & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      uint32_t acc[2][2][32];                        // [c2][gate | filter][column]
      __syncwarp();
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        tmem_ld32(taddr + (grp * 2 + c2) * 32, acc[c2][0]);
        tmem_ld32(taddr + 128 + (grp * 2 + c2) * 32, acc[c2][1]);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_cluster(smem_u32(tempty), 0));
#pragma unroll 1
      for (int pass = 0; pass < (DUAL ? 2 : 1); ++pass) {
        const bool is_cond = DUAL ? (pass == 1) : (ti.nb < p.n_cond);
        const int zroll = p.z_group0 + ti.nb + ((DUAL && pass == 0) ? p.dual_off : 0);
        if (issuer) tma_store_wait_read<0>();        // the previous stores no longer read the staging boxes
        named_bar_sync(EPI_BAR, EPI_THREADS);        // ... and every thread is done with sbias
        sbias[etid] = __ldg((is_cond ? p.bias_cond : p.bias_unc) + ti.nblk * TILE_N + etid);
        named_bar_sync(EPI_BAR, EPI_THREADS);
        const int tf = ti.t0 + row;
        const bool add_cond = is_cond && p.cond != nullptr && tf < p.T;
        const uint32_t box = stg + grp * CHUNK_BYTES;
        uint32_t zak[2][2][8];                       // aux words, staged after the main boxes have left
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int ch = grp * 2 + c2;
          const float* cg = p.cond + ((size_t)ti.nb * p.T + (add_cond ? tf : 0)) * (size_t)(2 * p.C) + ti.nblk * (TILE_N / 2) + ch * 32;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float z[16];
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
              float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
              if (add_cond) {
                a = __ldg(reinterpret_cast<const float4*>(cg) + hf * 4 + v4);
                b = __ldg(reinterpret_cast<const float4*>(cg + p.C) + hf * 4 + v4);
              }
              const float ca[4] = {a.x, a.y, a.z, a.w}, cb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = hf * 16 + v4 * 4 + e;
                const float g = __uint_as_float(acc[c2][0][j]) * inv + ca[e];
                const float f = __uint_as_float(acc[c2][1][j]) * inv + cb[e];
                z[v4 * 4 + e] = gate_act(g + sbias[ch * 32 + j], f + sbias[128 + ch * 32 + j]);
              }
            }
            uint32_t zm[8];
            pack16<3>(z, zm, zak[c2][hf]);
            const int chn = c2 * 32 + hf * 16;
            sts128u(box + sw128_off(row, chn / 8), zm[0], zm[1], zm[2], zm[3]);
            sts128u(box + sw128_off(row, chn / 8 + 1), zm[4], zm[5], zm[6], zm[7]);
          }
        }
        fence_proxy_async();
        named_bar_sync(EPI_BAR, EPI_THREADS);
        const int c0 = ti.nblk * (TILE_N / 2);
        if (issuer) {
          tma_store_3d(&p.zh, staging, c0, ti.t0, zroll);
          tma_store_3d(&p.zh, staging + CHUNK_BYTES, c0 + 64, ti.t0, zroll);
          tma_store_commit();
          tma_store_wait_read<0>();
        }
        named_bar_sync(EPI_BAR, EPI_THREADS);
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int chn = c2 * 32 + hf * 16;
            sts128u(box + sw128_off(row, chn / 16), zak[c2][hf][0], zak[c2][hf][1], zak[c2][hf][2], zak[c2][hf][3]);
            sts128u(box + sw128_off(row, 4 + chn / 16), zak[c2][hf][4], zak[c2][hf][5], zak[c2][hf][6], zak[c2][hf][7]);
          }
        fence_proxy_async();
        named_bar_sync(EPI_BAR, EPI_THREADS);
        if (issuer) {
          tma_store_3d(&p.zl, staging, 2 * c0, ti.t0, zroll);
          tma_store_3d(&p.zl, staging + CHUNK_BYTES, 2 * (c0 + 64), ti.t0, zroll);
          tma_store_commit();
        }
      }   // pass
    }
    if (issuer) tma_store_wait_read<0>();
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------
// zgemm kernel: A = stored z of one layer (RES) or of all layers (HEAD)
// ---------------------------------------------------------------------------------------------
template <int P, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1) umma_zgemm_kernel(const __grid_constant__ ZGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmemView sv = carve<P, PAIR>(smem_raw);
  using CF = Cfg<P, PAIR>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  int cid = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nblk = cid % p.n_blocks; cid /= p.n_blocks;
  // mode 3 with a guidance pair: the two CTAs of a pair take the SAME frames of roll b (conditional) and roll b + dual_B
  const bool dual3 = PAIR && p.mode == 3 && p.dual_B > 0;
  const int mt = dual3 ? cid : (PAIR ? cid * 2 + (int)rank : cid);
  const int tt = mt % p.tiles_t;
  const int nb = mt / p.tiles_t + (dual3 ? (int)rank * p.dual_B : 0);
  const int t0 = tt * TILE_M;
  const int n_base = nblk * TILE_N;
  const bool res = p.mode == 0;
  constexpr int AM = CF::kAuxMul;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.zh); tma_prefetch_desc(&p.w_h); tma_prefetch_desc(&p.out32);
    if (CF::kAux) { tma_prefetch_desc(&p.zl); tma_prefetch_desc(&p.w_l); }
  }
  if (warp == 3) {
    for (int i = lane; i < TILE_N; i += 32) {
      sv.sbias[i] = (p.bias && (p.mode != 3 || i < p.F)) ? __ldg(p.bias + n_base + i) : 0.f;
      if (res) sv.sdn[i] = __ldg(p.dnext + (size_t)(p.steps ? __ldg(p.steps + nb % p.bsamp) : p.t_uniform) * p.C + n_base + i);
    }
  }
  // RES on CTA pairs: the K loop is only C/64 slabs, so it runs on a 2-stage ring and the third stage's 64 KB hold the
  // first four fp32 x boxes, prefetched while the MMAs run; the other four follow once the ring is free.
  const bool early = res && PAIR && CF::kAux;
  const int nst = early ? 2 : CF::kStages;
  auto xbox_off = [&](int c) -> uint32_t {
    return early ? (c < 4 ? 2u * CF::kStageBytes + (uint32_t)c * CHUNK_BYTES : (uint32_t)(c - 4) * CHUNK_BYTES) : (uint32_t)c * CHUNK_BYTES;
  };
  const uint32_t set_off = early ? 4u * CHUNK_BYTES : 8u * CHUNK_BYTES;
  prologue<P, PAIR>(sv, warp);
  const uint32_t tmem_base = *sv.tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      if (early) {
        mbar_expect_tx(&sv.xin_full[0], 4 * CHUNK_BYTES);
        for (int c = 0; c < 4; ++c) tma_load_3d(sv.stage0 + xbox_off(c), &p.out32, &sv.xin_full[0], n_base + c * 32, t0, nb);
      }
      int stage = 0; uint32_t phase = 0;
      for (int s = 0; s < p.nslabs; ++s) {
        mbar_wait(&sv.empty[stage], phase ^ 1);
        uint8_t* st = sv.stage0 + stage * CF::kStageBytes;
        uint64_t* fl = &sv.full[stage];
        const uint32_t fb = PAIR ? mapa_cluster(smem_u32(fl), 0) : 0u;
        const int grp = s / p.spg;
        // split-K (mode 4, training weight gradients): the "roll" index of a tile selects its K range, every split writes its own
        // partial result [nb][rows][columns] and a reduction kernel adds them up in fp32
        const int ks = p.ksplit ? nb * p.nslabs : 0;
        const int cc = s - grp * p.spg + ks;
        const int zrow = p.ksplit ? 0 : p.z_group0 + grp * p.group_stride + nb;
        prod_expect<PAIR>(fl, fb, CF::kStageBytes);
        load_a<PAIR>(st + CF::kAOff, &p.zh, fl, fb, cc * TILE_K, t0, zrow);
        load_b<PAIR>(st + CF::kBOff, &p.w_h, fl, fb, (ks + s) * TILE_K, n_base, rank);
        if (CF::kAux) {
          load_a<PAIR>(st + CF::kAAuxOff, &p.zl, fl, fb, AM * cc * TILE_K, t0, zrow);
          load_b<PAIR>(st + CF::kBAuxOff, &p.w_l, fl, fb, AM * (ks + s) * TILE_K, n_base, rank);
        }
        if (++stage == nst) { stage = 0; phase ^= 1; }
      }
      if (res) {
        // Residual update needs the fp32 x tile: fetch the remaining boxes into the (now free) ring.
        mbar_wait(sv.tmem_full, 0);
        const int c_lo = early ? 4 : 0;
        mbar_expect_tx(&sv.xin_full[1], (8 - c_lo) * CHUNK_BYTES);
        for (int c = c_lo; c < 8; ++c) tma_load_3d(sv.stage0 + xbox_off(c), &p.out32, &sv.xin_full[1], n_base + c * 32, t0, nb);
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) mma_loop<P, PAIR>(sv, tmem_base, p.nslabs, nst);
  } else if (warp >= 4 && p.mode != 3) {
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int hc = (warp - 4) >> 2;        // two groups of 4 warps: each takes one 32-channel fp32 box per iteration
    const int row = q * 32 + lane;
    const bool issuer = (warp == 4) && (lane == 0);
    const bool hpair = p.mode == 1 && p.h_pair;   // HEAD: the ReLU output leaves as an operand pair (input of the tensor-core projection)
    mbar_wait(sv.tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t stg = smem_u32(sv.stage0);
    const float rsqrt2 = 0.70710678118654752f;
    const float inv = (P >= 2) ? __ldg(p.inv_scale) : 0.f;
    const uint32_t pset_off = hpair ? 0u : set_off;
    unsigned int umax = 0;
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {       // 64 output channels per iteration
      // operand staging of the next GEMM (x + d_next split, or h): two alternating sets of {main box, aux box}
      const uint32_t set = stg + pset_off + (it & 1) * 2 * CHUNK_BYTES;
      if ((res || hpair) && it >= 2) {     // the TMA stores of iteration it-2 must have finished reading this set
        if (issuer) tma_store_wait_read<1>();
        named_bar_sync(EPI_BAR, EPI_THREADS);
      }
      const int cbox = it * 2 + hc;
      float o[32];
      load_acc32<P>(taddr + cbox * 32, inv, o);
      const uint32_t box = stg + xbox_off(cbox);
      const float* bs = sv.sbias + cbox * 32;
      if (res) {
        if (it == 0 || (early && it == 2) ) mbar_wait(&sv.xin_full[(early && it == 0) ? 0 : 1], 0);
        const float* dn = sv.sdn + cbox * 32;
#pragma unroll
        for (int g16 = 0; g16 < 2; ++g16) {
          float xin[16];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int v = g16 * 4 + u, i = v * 4;
            const uint32_t xa = box + sw128_off(row, v);
            float4 x = lds128(xa);
            x.x = (x.x + (o[i + 0] + bs[i + 0])) * rsqrt2;   // (x + residual) / sqrt(2.0)   diffwave.py:151
            x.y = (x.y + (o[i + 1] + bs[i + 1])) * rsqrt2;
            x.z = (x.z + (o[i + 2] + bs[i + 2])) * rsqrt2;
            x.w = (x.w + (o[i + 3] + bs[i + 3])) * rsqrt2;
            sts128(xa, x);  // in place: same thread, same address
            xin[u * 4 + 0] = x.x + dn[i + 0]; xin[u * 4 + 1] = x.y + dn[i + 1];
            xin[u * 4 + 2] = x.z + dn[i + 2]; xin[u * 4 + 3] = x.w + dn[i + 3];
          }
          range_note(umax, xin);
          stage16<P>(set, set + CHUNK_BYTES, row, hc * 32 + g16 * 16, xin);
        }
      } else if (hpair) {
#pragma unroll
        for (int g16 = 0; g16 < 2; ++g16) {
          float hv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) hv[i] = fmaxf(o[g16 * 16 + i] + bs[g16 * 16 + i], 0.f);   // F.relu(skip_projection(.))   diffwave.py:683-684
          stage16<P>(set, set + CHUNK_BYTES, row, hc * 32 + g16 * 16, hv);
        }
      } else {
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const int i = v * 4;
          float4 h;
          const float floor_ = (p.mode == 2 || p.mode == 4) ? -INFINITY : 0.f;   // HEAD: ReLU; conditioner tables / plain GEMM: linear
          h.x = fmaxf(o[i + 0] + bs[i + 0], floor_);
          h.y = fmaxf(o[i + 1] + bs[i + 1], floor_);
          h.z = fmaxf(o[i + 2] + bs[i + 2], floor_);
          h.w = fmaxf(o[i + 3] + bs[i + 3], floor_);
          sts128(box + sw128_off(row, v), h);
        }
      }
      fence_proxy_async();
      named_bar_sync(EPI_BAR, EPI_THREADS);
      if (issuer) {
        // mode 2: the weight rows of an N block are [128 gate | 128 filter] channels; the table keeps the natural order
        const int c0 = p.mode == 2 ? (it < 2 ? nblk * (TILE_N / 2) + it * 64 : p.C / 2 + nblk * (TILE_N / 2) + (it - 2) * 64)
                                   : n_base + it * 64;
        if (!hpair) {
          tma_store_3d(&p.out32, sv.stage0 + xbox_off(it * 2), c0, t0, nb);
          tma_store_3d(&p.out32, sv.stage0 + xbox_off(it * 2 + 1), c0 + 32, t0, nb);
        }
        if (res || hpair) {
          tma_store_3d(&p.xh, sv.stage0 + pset_off + (it & 1) * 2 * CHUNK_BYTES, c0, t0, nb);
          if (CF::kAux) tma_store_3d(&p.xl, sv.stage0 + pset_off + (it & 1) * 2 * CHUNK_BYTES + CHUNK_BYTES, AM * c0, t0, nb);
        }
        tma_store_commit();
      }
    }
    tc_fence_before();
    if (issuer) tma_store_wait_read<0>();
    if (res) range_publish(p.range_max, umax);
  }
  // ---- mode 3: output_projection + guidance combine + posterior update (model/diffwave.py:685, task/diffusion.py:1009-1023) ----
  // Both CTAs park their projection tile [row][pitch] (fp32, row stride 92 floats) in the first CTA's idle operand ring -- the
  // second one through distributed shared memory --; after a cluster barrier the first CTA's epilogue threads walk the tile in
  // float4 steps along the pitches, so x_t / noise / x_prev rows are read and written fully coalesced.  The walk is a ROLLED
  // loop on purpose: the first version kept the 64 columns of a row in registers and unrolled the update per column, ~150 KB
  // of straight-line code executed once per CTA, and ran at instruction-fetch speed (50 us per wave, 'no_inst' stalls).
  if (p.mode == 3) {
    constexpr int RS = 92;                 // row stride in floats: 16-byte aligned rows, at most 4-way conflicts on the column-wise writes
    const bool epi = warp >= 4;
    float* const tile_c = reinterpret_cast<float*>(sv.stage0);                        // this CTA's own tile (conditional branch when dual)
    float* const tile_u = reinterpret_cast<float*>(sv.stage0) + TILE_M * RS;          // the pair's second CTA's tile (unconditional branch)
    if (epi) {
      const int q = warp & 3, hc = (warp - 4) >> 2;
      const int row = q * 32 + lane;
      mbar_wait(sv.tmem_full, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)hc * 64u;
      const float inv = (P >= 2) ? __ldg(p.inv_scale) : 0.f;
      const bool remote = dual3 && rank == 1;
      const uint32_t dst = remote ? mapa_cluster(smem_u32(tile_u), 0) : smem_u32(tile_c);
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        float a[32];
        load_acc32<P>(taddr + hlf * 32, inv, a);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int c = hc * 64 + hlf * 32 + i;
          if (c < p.F) {
            const uint32_t ad = dst + (uint32_t)(row * RS + c) * 4u;
            if (remote) asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ad), "f"(a[i]) : "memory");
            else asm volatile("st.shared.f32 [%0], %1;" ::"r"(ad), "f"(a[i]) : "memory");
          }
        }
      }
      tc_fence_before();
    }
    if (dual3) cluster_sync_all();         // every thread of both CTAs: both tiles are complete and visible in the first CTA
    else __syncthreads();
    if (epi && (!dual3 || rank == 0)) {
      const int etid = (int)threadIdx.x - 128;
      const int g_per_row = p.F >> 2;
      const float w = p.upd.w;
      const bool needs_x = !(p.upd.mode == DRB_UPD_X0_FINAL || p.upd.mode == DRB_UPD_NONE);
#pragma unroll 1
      for (int i = etid; i < TILE_M * g_per_row; i += EPI_THREADS) {
        const int row = i / g_per_row, c = (i - row * g_per_row) * 4;
        const int t = t0 + row;
        if (t >= p.T) continue;
        const size_t base = ((size_t)nb * p.T + t) * (size_t)p.F + c;
        const float4 cv = *reinterpret_cast<const float4*>(tile_c + row * RS + c);
        const float4 bv = *reinterpret_cast<const float4*>(sv.sbias + c);
        float net[4] = {cv.x + bv.x, cv.y + bv.y, cv.z + bv.z, cv.w + bv.w};
        if (dual3) {                                     // (1 + w) * x0_c - w * x0_0     task/diffusion.py:1009
          const float4 uv = *reinterpret_cast<const float4*>(tile_u + row * RS + c);
          net[0] = (1.f + w) * net[0] - w * (uv.x + bv.x); net[1] = (1.f + w) * net[1] - w * (uv.y + bv.y);
          net[2] = (1.f + w) * net[2] - w * (uv.z + bv.z); net[3] = (1.f + w) * net[3] - w * (uv.w + bv.w);
        }
        if (p.net_out) *reinterpret_cast<float4*>(p.net_out + base) = make_float4(net[0], net[1], net[2], net[3]);
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), nz = x;
        if (needs_x) x = *reinterpret_cast<const float4*>(p.x_t + base);
        if (p.upd.has_noise) nz = *reinterpret_cast<const float4*>(p.noise + base);
        float4 o;
        o.x = posterior_update(p.upd, net[0], x.x, nz.x); o.y = posterior_update(p.upd, net[1], x.y, nz.y);
        o.z = posterior_update(p.upd, net[2], x.z, nz.z); o.w = posterior_update(p.upd, net[3], x.w, nz.w);
        *reinterpret_cast<float4*>(p.x_prev + base) = o;
      }
    }
  }
  teardown<P, PAIR>(tmem_base, warp);
}


// ---------------------------------------------------------------------------------------------
// zgemm RES, persistent variant (CTA pairs; bf16x3 / f16e5).  The K loop of the residual GEMM is only C/64 slabs, so the
// one-tile-per-CTA kernel is dominated by its epilogue and per-CTA setup.  Here one CTA pair per SM pair loops over its
// tiles with two TMEM accumulator stages; the fp32 x tile streams through a 4-box ring filled by a second producer
// thread, is updated in place, and leaves (with the next layer's operand pair) through TMA stores, all overlapped with
// the MMAs of the following tile.
//   smem: 2 operand stages (128 KB) | x ring 4 x 16 KB | operand staging 32 KB
// ---------------------------------------------------------------------------------------------
constexpr int RP_STAGE = 65536;
constexpr int RP_STAGES = 2;
constexpr int RP_XBOXES = 4;    // 16 KB boxes after the operand stages: 2 landing boxes (x loads) + 2 store-staging boxes (x')
constexpr int RP_LANDING = 2;
constexpr int RP_RING = RP_STAGES * RP_STAGE + RP_XBOXES * CHUNK_BYTES;   // 192 KB
constexpr int RP_SMEM = RP_RING + 2 * CHUNK_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*bias*/;   // 231,680 <= 227 KB

template <int P, int XF>
__global__ void __launch_bounds__(NUM_THREADS, 1) umma_res_pers_kernel(const __grid_constant__ ZGemmParams p) {
  static_assert(P == 1 || P == 3, "needs a single 256-column accumulator");
  static_assert(XF == 0 || (XF == 4 && P == 3), "f16n4 activations pair with f16e5 z operands");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);
  constexpr int AM = P >= 2 ? 2 : 1;

  uint8_t* ring;
  { uint32_t a = smem_u32(smem_raw); ring = smem_raw + (((a + 1023u) & ~1023u) - a); }
  uint8_t* const xring = ring + RP_STAGES * RP_STAGE;
  uint8_t* const staging = ring + RP_RING;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(ring + RP_RING + 2 * CHUNK_BYTES);
  uint64_t* const full = bars;             // [2] operand stages (leader's is used)
  uint64_t* const empty = bars + 2;        // [2]
  uint64_t* const tfull = bars + 4;        // [2]
  uint64_t* const tempty = bars + 6;       // [2]
  uint64_t* const xfull = bars + 8;        // [4] x boxes (per CTA)
  uint64_t* const xempty = bars + 12;      // [4]
  uint32_t* const tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);
  float* const sbias = reinterpret_cast<float*>(ring + RP_RING + 2 * CHUNK_BYTES + 256);

  struct Tile { int n_base, nb, t0; };
  auto tile_of = [&](int item) -> Tile {
    Tile t;
    t.n_base = (item % p.n_blocks) * TILE_N;
    const int mt = (item / p.n_blocks) * 2 + (int)rank;
    t.nb = mt / p.tiles_t;
    t.t0 = (mt % p.tiles_t) * TILE_M;
    return t;
  };
  const int n_items = (p.NB * p.tiles_t / 2) * p.n_blocks;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.zh); tma_prefetch_desc(&p.zl); tma_prefetch_desc(&p.w_h); tma_prefetch_desc(&p.w_l);
    tma_prefetch_desc(&p.out32); tma_prefetch_desc(&p.xh); tma_prefetch_desc(&p.xl);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 2); mbar_init(&empty[i], 1);
      mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16);
    }
    for (int i = 0; i < RP_XBOXES; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 4); }   // 4 warps read a landing box
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc_pair(tmem_ptr, 512); tmem_relinquish_pair(); }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {          // operand producer
      int cnt = 0;
      for (int item = pair_id; item < n_items; item += n_pairs) {
        const Tile ti = tile_of(item);
        for (int s = 0; s < p.nslabs; ++s, ++cnt) {
          const int st_i = cnt % RP_STAGES;
          mbar_wait(&empty[st_i], ((cnt / RP_STAGES) & 1) ^ 1);
          uint8_t* st = ring + st_i * RP_STAGE;
          const uint32_t fb = mapa_cluster(smem_u32(&full[st_i]), 0);
          mbar_expect_tx_cluster(fb, RP_STAGE);
          const int zrow = p.z_group0 + ti.nb;
          const int row0 = ti.n_base + (int)rank * (TILE_N / 2);
          tma_load_3d_pair(st, &p.zh, fb, s * TILE_K, ti.t0, zrow);
          tma_load_3d_pair(st + A_TILE_BYTES, &p.zl, fb, AM * s * TILE_K, ti.t0, zrow);
          tma_load_2d_pair(st + 2 * A_TILE_BYTES, &p.w_h, fb, s * TILE_K, row0);
          tma_load_2d_pair(st + 2 * A_TILE_BYTES + B_TILE_BYTES / 2, &p.w_l, fb, AM * s * TILE_K, row0);
        }
      }
    }
  } else if (warp == 3) {
    if (elect_one()) {          // x-box producer (this CTA's fp32 residual tiles)
      int gb = 0;
      for (int item = pair_id; item < n_items; item += n_pairs) {
        const Tile ti = tile_of(item);
        for (int c = 0; c < 8; ++c, ++gb) {
          const int sl = gb % RP_LANDING;
          mbar_wait(&xempty[sl], ((gb / RP_LANDING) & 1) ^ 1);
          mbar_expect_tx(&xfull[sl], CHUNK_BYTES);
          tma_load_3d(xring + sl * CHUNK_BYTES, &p.out32, &xfull[sl], ti.n_base + c * 32, ti.t0, ti.nb);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {   // MMA issuer
      constexpr uint32_t idesc = P >= 2 ? make_idesc_fmt0(2 * TILE_M, TILE_N) : make_idesc_bf16(2 * TILE_M, TILE_N);
      constexpr uint32_t idesc_e5 = make_idesc_bf16(2 * TILE_M, TILE_N);
      int cnt = 0, tcnt = 0;
      for (int item = pair_id; item < n_items; item += n_pairs, ++tcnt) {
        const int as = tcnt & 1;
        mbar_wait(&tempty[as], ((tcnt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)as * 256u;
        for (int s = 0; s < p.nslabs; ++s, ++cnt) {
          const int st_i = cnt % RP_STAGES;
          mbar_wait(&full[st_i], (cnt / RP_STAGES) & 1);
          tc_fence_after();
          const uint32_t a_main = smem_u32(ring + st_i * RP_STAGE), a_aux = a_main + A_TILE_BYTES;
          const uint32_t b_main = a_main + 2 * A_TILE_BYTES, b_aux = b_main + B_TILE_BYTES / 2;
#pragma unroll
          for (int k = 0; k < TILE_K / UMMA_K; ++k) {
            const uint32_t ko = k * UMMA_K * 2;
            const uint64_t da = make_sw128_desc(a_main + ko), db = make_sw128_desc(b_main + ko);
            umma_bf16_pair(tmem_d, da, db, idesc, (s == 0 && k == 0) ? 0u : 1u);
            if (P == 1) {
              umma_bf16_pair(tmem_d, make_sw128_desc(a_aux + ko), db, idesc, 1u);
              umma_bf16_pair(tmem_d, da, make_sw128_desc(b_aux + ko), idesc, 1u);
            }
            if (P == 3) umma_f8_pair(tmem_d, make_sw128_desc(a_aux + ko), make_sw128_desc(b_aux + ko), idesc_e5, 1u);
          }
          umma_commit_pair(&empty[st_i]);
        }
        umma_commit_pair(&tfull[as]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3, hc = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int etid = (int)threadIdx.x - 128;
    const bool issuer = (warp == 4) && (lane == 0);
    const uint32_t stg = smem_u32(staging);
    const float rsqrt2 = 0.70710678118654752f;
    const float inv = (P >= 2) ? __ldg(p.inv_scale) : 0.f;
    unsigned int umax = 0;
    // x pipeline: the fp32 x tile arrives by TMA in a landing box (one per 4-warp group), is pulled into registers one
    // iteration AHEAD and the box handed straight back to the x producer, so the next load has a whole iteration to
    // arrive (ncu: 22 % of this kernel's stall samples sat on the x barrier when a box stayed occupied until its TMA
    // store had drained).  x' and the next layer's operand pair leave through separate store-staging boxes.
    uint8_t* const xstage = xring + RP_LANDING * CHUNK_BYTES;
    const uint32_t land = smem_u32(xring + hc * CHUNK_BYTES);
    const uint32_t xst = smem_u32(xstage + hc * CHUNK_BYTES);
    const int my_tiles = pair_id < n_items ? (n_items - pair_id + n_pairs - 1) / n_pairs : 0;
    const int total_it = 4 * my_tiles;
    float xc[32];
    auto fetch_x = [&](int g) {                     // x values of global iteration g -> registers, landing box released
      mbar_wait(&xfull[hc], g & 1);
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const float4 t = lds128(land + sw128_off(row, v));
        xc[4 * v] = t.x; xc[4 * v + 1] = t.y; xc[4 * v + 2] = t.z; xc[4 * v + 3] = t.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&xempty[hc]);
    };
    if (total_it > 0) fetch_x(0);
    int tcnt = 0, git = 0;
    for (int item = pair_id; item < n_items; item += n_pairs, ++tcnt) {
      const Tile ti = tile_of(item);
      const int as = tcnt & 1;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * 256u;
#pragma unroll 1
      for (int it = 0; it < 4; ++it, ++git) {
        // Everything up to the staging writes (TMEM load, residual update, operand split, all in registers) overlaps the
        // TMA stores of the previous iteration, which are still reading the staging boxes; only then does the issuer
        // wait for those reads.  (Measured alternative: writing x' and the operand pair straight from registers to
        // global memory, no staging and no barriers, is 50 % SLOWER -- 1.84 vs 1.23 ms per step: per-row 16-byte stores
        // cost more than the synchronisation they remove.)
        if (it == 0) {
          // every thread has passed the last barrier of the previous tile, i.e. finished reading its sbias
          sbias[etid] = __ldg(p.bias + ti.n_base + etid);
          named_bar_sync(EPI_BAR, EPI_THREADS);
          mbar_wait(&tfull[as], (tcnt >> 1) & 1);
          tc_fence_after();
        }
        const int cbox = it * 2 + hc;
        float o[32];
        load_acc32<P>(taddr + cbox * 32, inv, o);
        if (it == 3) {                              // last TMEM read of this tile: hand the accumulator stage back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&tempty[as]), 0));
        }
        const float* bs = sbias + cbox * 32;
        // warp-uniform addresses: L1 broadcast, off the critical path
        const float* dn = p.dnext + (size_t)(p.steps ? __ldg(p.steps + ti.nb % p.bsamp) : p.t_uniform) * p.C + ti.n_base + cbox * 32;
        uint32_t pm[2][8], pa[2][8];                // packed operand pair of this thread's 32 channels
        uint32_t sfl[2], sfh[2];                    // f16n4: scale bytes of the two 16-channel blocks (lo part, hi part)
#pragma unroll
        for (int g16 = 0; g16 < 2; ++g16) {
          float xin[16];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = (g16 * 4 + u) * 4;
            const float4 d4 = __ldg(reinterpret_cast<const float4*>(dn + i));
#pragma unroll
            for (int e = 0; e < 4; ++e) xc[i + e] = (xc[i + e] + (o[i + e] + bs[i + e])) * rsqrt2;   // (x + residual) / sqrt(2.0)   diffwave.py:151
            xin[u * 4 + 0] = xc[i + 0] + d4.x; xin[u * 4 + 1] = xc[i + 1] + d4.y;
            xin[u * 4 + 2] = xc[i + 2] + d4.z; xin[u * 4 + 3] = xc[i + 3] + d4.w;
          }
          range_note(umax, xin);
          if (XF == 4) {   // pa: [lo codes (2 words) | hi codes (2 words)]
            uint32_t lo[2], hi[2];
            split_f16n4_x16(xin, pm[g16], lo, hi, sfl[g16], sfh[g16]);
            pa[g16][0] = lo[0]; pa[g16][1] = lo[1]; pa[g16][2] = hi[0]; pa[g16][3] = hi[1];
          } else {
            pack16<P>(xin, pm[g16], pa[g16]);
          }
        }
        if (issuer) tma_store_wait_read<0>();       // the previous iteration's stores no longer read the staging boxes
        named_bar_sync(EPI_BAR, EPI_THREADS);
#pragma unroll
        for (int v = 0; v < 8; ++v) sts128(xst + sw128_off(row, v), make_float4(xc[4 * v], xc[4 * v + 1], xc[4 * v + 2], xc[4 * v + 3]));
#pragma unroll
        for (int g16 = 0; g16 < 2; ++g16) {
          const int chn = hc * 32 + g16 * 16;
          sts128u(stg + sw128_off(row, chn / 8), pm[g16][0], pm[g16][1], pm[g16][2], pm[g16][3]);
          sts128u(stg + sw128_off(row, chn / 8 + 1), pm[g16][4], pm[g16][5], pm[g16][6], pm[g16][7]);
          if (XF == 4) {
            // aux box: 64-byte rows [lo codes 32 B | hi codes 32 B], 64-byte swizzle; this thread owns bytes [16 hc, 16 hc + 16) of each half
            if (g16 == 1) {
              sts128u(stg + CHUNK_BYTES + sw64_off(row, hc), pa[0][0], pa[0][1], pa[1][0], pa[1][1]);
              sts128u(stg + CHUNK_BYTES + sw64_off(row, 2 + hc), pa[0][2], pa[0][3], pa[1][2], pa[1][3]);
            }
          } else if (P >= 2) {
            sts128u(stg + CHUNK_BYTES + sw128_off(row, chn / 16), pa[g16][0], pa[g16][1], pa[g16][2], pa[g16][3]);
            sts128u(stg + CHUNK_BYTES + sw128_off(row, 4 + chn / 16), pa[g16][4], pa[g16][5], pa[g16][6], pa[g16][7]);
          } else {
            sts128u(stg + CHUNK_BYTES + sw128_off(row, chn / 8), pa[g16][0], pa[g16][1], pa[g16][2], pa[g16][3]);
            sts128u(stg + CHUNK_BYTES + sw128_off(row, chn / 8 + 1), pa[g16][4], pa[g16][5], pa[g16][6], pa[g16][7]);
          }
        }
        fence_proxy_async();
        named_bar_sync(EPI_BAR, EPI_THREADS);
        if (issuer) {
          const int c0 = ti.n_base + it * 64;
          tma_store_3d(&p.out32, xstage, c0, ti.t0, ti.nb);
          tma_store_3d(&p.out32, xstage + CHUNK_BYTES, c0 + 32, ti.t0, ti.nb);
          tma_store_3d(&p.xh, staging, c0, ti.t0, ti.nb);
          tma_store_3d(&p.xl, staging + CHUNK_BYTES, XF == 4 ? c0 : AM * c0, ti.t0, ti.nb);   // f16n4: 64 aux bytes per 64 channels
          tma_store_commit();
        }
        if (XF == 4 && ti.t0 + row < p.T) {   // scale factors [roll][chunk][frame][lo x4 | hi x4]: this thread's blocks 2 hc, 2 hc + 1
          uint8_t* sp = p.xs + (((size_t)ti.nb * (p.C / TILE_K) + (ti.n_base / TILE_K + it)) * p.T + (ti.t0 + row)) * 8 + 2 * hc;
          *reinterpret_cast<uint16_t*>(sp) = (uint16_t)(sfl[0] | (sfl[1] << 8));
          *reinterpret_cast<uint16_t*>(sp + 4) = (uint16_t)(sfh[0] | (sfh[1] << 8));
        }
        if (git + 1 < total_it) fetch_x(git + 1);   // next iteration's x: usually landed long ago
      }
    }
    if (issuer) tma_store_wait_read<0>();
    range_publish(p.range_max, umax);
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------
// zgemm HEAD, persistent variant (CTA pairs; bf16x3 / f16e5; output = operand pair of relu(.) for the tensor-core projection).
// The one-tile-per-CTA kernel pays ~40 us per wave around an 80 us K loop (prologue, first TMA round trip, four store
// iterations, teardown): here one CTA pair per SM pair loops over its tiles with two TMEM accumulator stages, so the
// epilogue of tile i runs under the 120 K-slabs of tile i+1.
//   smem: 3 operand stages x 64 KB | one staging set (main + aux box, 32 KB)
// ---------------------------------------------------------------------------------------------
constexpr int HP_STAGE = 65536, HP_STAGES = 3;
constexpr int HP_RING = HP_STAGES * HP_STAGE;
constexpr int HP_SMEM = HP_RING + 2 * CHUNK_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*bias*/;

template <int P>
__global__ void __launch_bounds__(NUM_THREADS, 1) umma_head_pers_kernel(const __grid_constant__ ZGemmParams p) {
  static_assert(P == 1 || P == 3, "needs a single 256-column accumulator");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);
  constexpr int AM = P >= 2 ? 2 : 1;

  uint8_t* ring;
  { uint32_t a = smem_u32(smem_raw); ring = smem_raw + (((a + 1023u) & ~1023u) - a); }
  uint8_t* const staging = ring + HP_RING;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(ring + HP_RING + 2 * CHUNK_BYTES);
  uint64_t* const full = bars;             // [3]
  uint64_t* const empty = bars + 3;        // [3]
  uint64_t* const tfull = bars + 6;        // [2]
  uint64_t* const tempty = bars + 8;       // [2]
  uint32_t* const tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);
  float* const sbias = reinterpret_cast<float*>(ring + HP_RING + 2 * CHUNK_BYTES + 256);

  struct Tile { int n_base, nb, t0; };
  auto tile_of = [&](int item) -> Tile {
    Tile t;
    t.n_base = (item % p.n_blocks) * TILE_N;
    const int mt = (item / p.n_blocks) * 2 + (int)rank;
    t.nb = mt / p.tiles_t;
    t.t0 = (mt % p.tiles_t) * TILE_M;
    return t;
  };
  const int n_items = (p.NB * p.tiles_t / 2) * p.n_blocks;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.zh); tma_prefetch_desc(&p.zl); tma_prefetch_desc(&p.w_h); tma_prefetch_desc(&p.w_l);
    tma_prefetch_desc(&p.xh); tma_prefetch_desc(&p.xl);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < HP_STAGES; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc_pair(tmem_ptr, 512); tmem_relinquish_pair(); }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {          // operand producer
      int cnt = 0;
      for (int item = pair_id; item < n_items; item += n_pairs) {
        const Tile ti = tile_of(item);
        const int row0 = ti.n_base + (int)rank * (TILE_N / 2);
        for (int s = 0; s < p.nslabs; ++s, ++cnt) {
          const int st_i = cnt % HP_STAGES;
          mbar_wait(&empty[st_i], ((cnt / HP_STAGES) & 1) ^ 1);
          uint8_t* st = ring + st_i * HP_STAGE;
          const uint32_t fb = mapa_cluster(smem_u32(&full[st_i]), 0);
          mbar_expect_tx_cluster(fb, HP_STAGE);
          const int grp = s / p.spg, cc = s - grp * p.spg;
          const int zrow = p.z_group0 + grp * p.group_stride + ti.nb;
          tma_load_3d_pair(st, &p.zh, fb, cc * TILE_K, ti.t0, zrow);
          tma_load_3d_pair(st + A_TILE_BYTES, &p.zl, fb, AM * cc * TILE_K, ti.t0, zrow);
          tma_load_2d_pair(st + 2 * A_TILE_BYTES, &p.w_h, fb, s * TILE_K, row0);
          tma_load_2d_pair(st + 2 * A_TILE_BYTES + B_TILE_BYTES / 2, &p.w_l, fb, AM * s * TILE_K, row0);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {   // MMA issuer
      constexpr uint32_t idesc = P >= 2 ? make_idesc_fmt0(2 * TILE_M, TILE_N) : make_idesc_bf16(2 * TILE_M, TILE_N);
      constexpr uint32_t idesc_e5 = make_idesc_bf16(2 * TILE_M, TILE_N);
      int cnt = 0, tcnt = 0;
      for (int item = pair_id; item < n_items; item += n_pairs, ++tcnt) {
        const int as = tcnt & 1;
        mbar_wait(&tempty[as], ((tcnt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)as * 256u;
        for (int s = 0; s < p.nslabs; ++s, ++cnt) {
          const int st_i = cnt % HP_STAGES;
          mbar_wait(&full[st_i], (cnt / HP_STAGES) & 1);
          tc_fence_after();
          const uint32_t a_main = smem_u32(ring + st_i * HP_STAGE), a_aux = a_main + A_TILE_BYTES;
          const uint32_t b_main = a_main + 2 * A_TILE_BYTES, b_aux = b_main + B_TILE_BYTES / 2;
#pragma unroll
          for (int k = 0; k < TILE_K / UMMA_K; ++k) {
            const uint32_t ko = k * UMMA_K * 2;
            const uint64_t da = make_sw128_desc(a_main + ko), db = make_sw128_desc(b_main + ko);
            umma_bf16_pair(tmem_d, da, db, idesc, (s == 0 && k == 0) ? 0u : 1u);
            if (P == 1) {
              umma_bf16_pair(tmem_d, make_sw128_desc(a_aux + ko), db, idesc, 1u);
              umma_bf16_pair(tmem_d, da, make_sw128_desc(b_aux + ko), idesc, 1u);
            }
            if (P == 3) umma_f8_pair(tmem_d, make_sw128_desc(a_aux + ko), make_sw128_desc(b_aux + ko), idesc_e5, 1u);
          }
          umma_commit_pair(&empty[st_i]);
        }
        umma_commit_pair(&tfull[as]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3, hc = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int etid = (int)threadIdx.x - 128;
    const bool issuer = (warp == 4) && (lane == 0);
    const uint32_t stg = smem_u32(staging);
    const float inv = (P >= 2) ? __ldg(p.inv_scale) : 0.f;
    int tcnt = 0;
    for (int item = pair_id; item < n_items; item += n_pairs, ++tcnt) {
      const Tile ti = tile_of(item);
      const int as = tcnt & 1;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * 256u;
#pragma unroll 1
      for (int it = 0; it < 4; ++it) {
        if (it == 0) {
          named_bar_sync(EPI_BAR, EPI_THREADS);       // every thread is done with the previous tile's sbias
          sbias[etid] = p.bias ? __ldg(p.bias + ti.n_base + etid) : 0.f;
          named_bar_sync(EPI_BAR, EPI_THREADS);
          mbar_wait(&tfull[as], (tcnt >> 1) & 1);
          tc_fence_after();
        }
        const int cbox = it * 2 + hc;
        float o[32];
        load_acc32<P>(taddr + cbox * 32, inv, o);
        if (it == 3) {                                // last TMEM read of this tile: hand the accumulator stage back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&tempty[as]), 0));
        }
        const float* bs = sbias + cbox * 32;
        uint32_t pm[2][8], pa[2][8];
#pragma unroll
        for (int g16 = 0; g16 < 2; ++g16) {
          float hv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) hv[i] = fmaxf(o[g16 * 16 + i] + bs[g16 * 16 + i], 0.f);   // F.relu(skip_projection(.))   diffwave.py:683-684
          pack16<P>(hv, pm[g16], pa[g16]);
        }
        if (issuer) tma_store_wait_read<0>();         // the previous iteration's stores no longer read the staging boxes
        named_bar_sync(EPI_BAR, EPI_THREADS);
#pragma unroll
        for (int g16 = 0; g16 < 2; ++g16) {
          const int chn = hc * 32 + g16 * 16;
          sts128u(stg + sw128_off(row, chn / 8), pm[g16][0], pm[g16][1], pm[g16][2], pm[g16][3]);
          sts128u(stg + sw128_off(row, chn / 8 + 1), pm[g16][4], pm[g16][5], pm[g16][6], pm[g16][7]);
          if (P >= 2) {
            sts128u(stg + CHUNK_BYTES + sw128_off(row, chn / 16), pa[g16][0], pa[g16][1], pa[g16][2], pa[g16][3]);
            sts128u(stg + CHUNK_BYTES + sw128_off(row, 4 + chn / 16), pa[g16][4], pa[g16][5], pa[g16][6], pa[g16][7]);
          } else {
            sts128u(stg + CHUNK_BYTES + sw128_off(row, chn / 8), pa[g16][0], pa[g16][1], pa[g16][2], pa[g16][3]);
            sts128u(stg + CHUNK_BYTES + sw128_off(row, chn / 8 + 1), pa[g16][4], pa[g16][5], pa[g16][6], pa[g16][7]);
          }
        }
        fence_proxy_async();
        named_bar_sync(EPI_BAR, EPI_THREADS);
        if (issuer) {
          const int c0 = ti.n_base + it * 64;
          tma_store_3d(&p.xh, staging, c0, ti.t0, ti.nb);
          tma_store_3d(&p.xl, staging + CHUNK_BYTES, AM * c0, ti.t0, ti.nb);
          tma_store_commit();
        }
      }
    }
    if (issuer) tma_store_wait_read<0>();
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

int umma_init() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    set_error("cuTensorMapEncodeTiled not available (%s)", cudaGetErrorString(e));
    return DRB_E_DRIVER;
  }
  g_encode = (EncodeTiledFn)fn;
  cudaError_t ee = cudaSuccess;
  auto set = [&](const void* fn, int bytes) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess && ee == cudaSuccess) ee = e;
  };
  set((const void*)umma_gate_kernel<0, false>, Cfg<0, false>::kSmemBytes); set((const void*)umma_gate_kernel<0, true>, Cfg<0, false>::kSmemBytes);
  set((const void*)umma_gate_kernel<1, false>, Cfg<1, false>::kSmemBytes); set((const void*)umma_gate_kernel<1, true>, Cfg<1, false>::kSmemBytes);
  set((const void*)umma_gate_kernel<2, false>, Cfg<2, false>::kSmemBytes); set((const void*)umma_gate_kernel<2, true>, Cfg<2, false>::kSmemBytes);
  set((const void*)umma_gate_kernel<3, false>, Cfg<3, false>::kSmemBytes); set((const void*)umma_gate_kernel<3, true>, Cfg<3, false>::kSmemBytes);
  set((const void*)umma_gate_kernel<4, false>, Cfg<4, false>::kSmemBytes); set((const void*)umma_gate_kernel<4, true>, Cfg<4, false>::kSmemBytes);
  set((const void*)umma_gate_win_kernel<1>, Cfg<1, true>::kSmemBytes); set((const void*)umma_gate_win_kernel<2>, Cfg<2, true>::kSmemBytes);
  set((const void*)umma_gate_win_kernel<3>, Cfg<3, true>::kSmemBytes);
  set((const void*)umma_gate_pers_kernel<1, false>, PW_SMEM); set((const void*)umma_gate_pers_kernel<3, false>, PW_SMEM);
  set((const void*)umma_gate_pers_kernel<1, true>, PW_SMEM); set((const void*)umma_gate_pers_kernel<3, true>, PW_SMEM);
  set((const void*)umma_res_pers_kernel<1, 0>, RP_SMEM); set((const void*)umma_res_pers_kernel<3, 0>, RP_SMEM);
  set((const void*)umma_res_pers_kernel<3, 4>, RP_SMEM);
  set((const void*)umma_head_pers_kernel<1>, HP_SMEM); set((const void*)umma_head_pers_kernel<3>, HP_SMEM);
  set((const void*)umma_conv_lin_pers_kernel<3>, CLP_SMEM); set((const void*)umma_conv_lin_pers_kernel<4>, CLP_SMEM);
  set((const void*)umma_gate_n4_kernel<false>, N4_SMEM); set((const void*)umma_gate_n4_kernel<true>, N4_SMEM);
  set((const void*)umma_zgemm_kernel<0, false>, Cfg<0, false>::kSmemBytes); set((const void*)umma_zgemm_kernel<0, true>, Cfg<0, false>::kSmemBytes);
  set((const void*)umma_zgemm_kernel<1, false>, Cfg<1, false>::kSmemBytes); set((const void*)umma_zgemm_kernel<1, true>, Cfg<1, false>::kSmemBytes);
  set((const void*)umma_zgemm_kernel<2, false>, Cfg<2, false>::kSmemBytes); set((const void*)umma_zgemm_kernel<2, true>, Cfg<2, false>::kSmemBytes);
  set((const void*)umma_zgemm_kernel<3, false>, Cfg<3, false>::kSmemBytes); set((const void*)umma_zgemm_kernel<3, true>, Cfg<3, false>::kSmemBytes);
  if (ee) {
    set_error("cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(ee));
    g_encode = nullptr;
    return (int)ee;
  }
  return 0;
}

static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, int dtype, int swizzle_bytes = 128) {  // dtype: 0 bf16, 1 fp32, 2 fp16, 3 uint8
  if (!g_encode) { int r = umma_init(); if (r) return r; }
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                 : dtype == 3 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                                                       : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(m, dt, (cuuint32_t)rank,
                        const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return DRB_E_DRIVER; }
  return 0;
}

static int esize(int dtype) { return dtype == 1 ? 4 : dtype == 3 ? 1 : 2; }

// [rows][cols] row-major, box = box_rows x (128 bytes of columns)
int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, int dtype) {
  const int es = esize(dtype);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * es};
  cuuint32_t box[2] = {(cuuint32_t)(128 / es), box_rows};
  return encode(m, base, 2, dims, strides, box, dtype);
}

// [d2][d1][d0], box = 1 x box1 x (128 bytes of d0)
int make_tmap_3d(CUtensorMap* m, const void* base, uint64_t d2, uint64_t d1, uint64_t d0, uint32_t box1, int dtype) {
  const int es = esize(dtype);
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * es, d1 * d0 * es};
  cuuint32_t box[3] = {(cuuint32_t)(128 / es), box1, 1};
  return encode(m, base, 3, dims, strides, box, dtype);
}

// byte tensors with an explicit inner box width and swizzle (f16n4: 64-byte e2m1 rows under SWIZZLE_64B, un-swizzled scale atoms)
int make_tmap_2d_bytes(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols, int swizzle_bytes) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols};
  cuuint32_t box[2] = {box_cols, box_rows};
  return encode(m, base, 2, dims, strides, box, 3, swizzle_bytes);
}
int make_tmap_3d_bytes(CUtensorMap* m, const void* base, uint64_t d2, uint64_t d1, uint64_t d0, uint32_t box1, uint32_t box0, int swizzle_bytes) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0, d1 * d0};
  cuuint32_t box[3] = {box0, box1, 1};
  return encode(m, base, 3, dims, strides, box, 3, swizzle_bytes);
}

static int g_pdl = -1;

// Launch `kernel` on `grid` CTAs of 256 threads; cluster = 2 consecutive CTAs (a CTA pair) when mc.
template <class Params>
static int launch_k(void (*kernel)(Params), const Params& p, int grid, int smem, bool mc, cudaStream_t s) {
  if (g_pdl < 0) { const char* e = getenv("DRB_NO_PDL"); g_pdl = (e && e[0] == '1') ? 0 : 1; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = mc ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: overlap this kernel's prologue with its predecessor
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_pdl ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p);
  count_launch();
  if (e != cudaSuccess) { set_error("cudaLaunchKernelEx: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

int launch_umma_gate(const UmmaMaps& maps, const UmmaLayer& L, const UmmaGate& g, cudaStream_t s) {
  if (g.C % TILE_N || g.Mp % TILE_K || (g.taps & 1) == 0) {
    set_error("umma_gate: unsupported C=%d Mp=%d taps=%d", g.C, g.Mp, g.taps);
    return DRB_E_INVALID;
  }
  GateParams p;
  p.xh = maps.xh; p.xl = maps.xl; p.sh = maps.sh; p.sl = maps.sl; p.zh = maps.zh; p.zl = maps.zl;
  p.wd_h = L.wd_h; p.wd_l = L.wd_l; p.wc_h = L.wc_h; p.wc_l = L.wc_l;
  p.NB = g.NB; p.n_cond = g.n_cond; p.T = g.T; p.C = g.C; p.taps = g.taps; p.dil = g.dil;
  p.cond_slabs = g.Mp / TILE_K; p.tiles_t = (g.T + TILE_M - 1) / TILE_M; p.n_blocks = 2 * g.C / TILE_N;
  p.z_group0 = g.z_group0;
  p.bias_cond = g.bias_cond; p.bias_unc = g.bias_unc;
  p.cond = nullptr; p.n_cond_mma = p.n_cond; p.dual_off = 0; p.xs = nullptr; p.lin_out = nullptr; p.ldo = 0;
  p.tap_lo = 0; p.tap_n = 0; p.lin_acc = 0; p.tap_span = 0; p.lin_scratch = nullptr;
  const int grid = p.NB * p.tiles_t * p.n_blocks;
  p.inv_scale = g.inv_scale;
  const bool mc = g.pair && ((p.NB * p.tiles_t) % 2 == 0);
  // window variant: the tap window 128 + (taps-1)*dil frames must fit the 192-row buffers
  const int win_rows = TILE_M + (g.taps - 1) * g.dil;
  if (mc && g.window && g.prec != 0 && win_rows <= 192 && g.xwh && g.xwl) {
    p.xwh = *g.xwh; p.xwl = *g.xwl; p.win_rows = win_rows;
    p.n_items = (p.NB * p.tiles_t / 2) * p.n_blocks;
    if (g.n4) {   // f16n4: persistent pairs only, conditioner term always from the per-clip table
      if (!g.persistent || !g.xw4 || !g.wd4 || !g.wsf || !g.xs || (g.n_cond > 0 && !g.cond) || win_rows > 192) {
        set_error("umma_gate: f16n4 needs the persistent window kernel and the conditioner tables"); return DRB_E_INVALID;
      }
      int n_sm = 148;
      { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
      p.xw4 = *g.xw4; p.wd4 = *g.wd4; p.wsf = *g.wsf; p.xs = g.xs;
      p.cond = g.cond; p.n_cond_mma = 0;
      const bool dual = g.dual_B > 0 && g.cond && g.NB == 2 * g.dual_B && ((g.dual_B * p.tiles_t) % 2 == 0);
      if (dual) { p.NB = g.dual_B; p.dual_off = g.dual_B; p.n_items = (p.NB * p.tiles_t / 2) * p.n_blocks; }
      const int pairs = p.n_items < n_sm / 2 ? p.n_items : n_sm / 2;
      return dual ? launch_k(umma_gate_n4_kernel<true>, p, 2 * pairs, N4_SMEM, true, s)
                  : launch_k(umma_gate_n4_kernel<false>, p, 2 * pairs, N4_SMEM, true, s);
    }
    if (g.persistent && (g.prec == 1 || g.prec == 3)) {   // one CTA pair per SM pair, looping over its tiles
      int n_sm = 148;
      { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
      if (g.cond) { p.cond = g.cond; p.n_cond_mma = 0; }   // conditioner term added in the epilogue, no cond K-slabs
      else if (g.need_tables) { set_error("umma_gate: conditioner tables are required but were not given"); return DRB_E_INVALID; }
      // layer-0 branch sharing: conv tiles over the conditional rolls only, two gated outputs per tile
      const bool dual = g.dual_B > 0 && g.cond && g.NB == 2 * g.dual_B && ((g.dual_B * p.tiles_t) % 2 == 0);
      if (dual) {
        p.NB = g.dual_B; p.dual_off = g.dual_B;
        p.n_items = (p.NB * p.tiles_t / 2) * p.n_blocks;
      }
      const int pairs = p.n_items < n_sm / 2 ? p.n_items : n_sm / 2;
      if (dual) return g.prec == 1 ? launch_k(umma_gate_pers_kernel<1, true>, p, 2 * pairs, PW_SMEM, true, s)
                                   : launch_k(umma_gate_pers_kernel<3, true>, p, 2 * pairs, PW_SMEM, true, s);
      return g.prec == 1 ? launch_k(umma_gate_pers_kernel<1, false>, p, 2 * pairs, PW_SMEM, true, s)
                         : launch_k(umma_gate_pers_kernel<3, false>, p, 2 * pairs, PW_SMEM, true, s);
    }
    if (g.need_tables) { set_error("umma_gate: conditioner tables are required but this kernel variant contracts the spectrogram"); return DRB_E_INVALID; }
    return g.prec == 1 ? launch_k(umma_gate_win_kernel<1>, p, grid, Cfg<1, true>::kSmemBytes, true, s)
         : g.prec == 2 ? launch_k(umma_gate_win_kernel<2>, p, grid, Cfg<2, true>::kSmemBytes, true, s)
                       : launch_k(umma_gate_win_kernel<3>, p, grid, Cfg<3, true>::kSmemBytes, true, s);
  }
  if (g.n4) { set_error("umma_gate: f16n4 needs CTA pairs (even tile count) and the tap window"); return DRB_E_INVALID; }
  if (g.need_tables) { set_error("umma_gate: conditioner tables are required but this kernel variant contracts the spectrogram"); return DRB_E_INVALID; }
  p.xwh = maps.xh; p.xwl = maps.xl; p.win_rows = TILE_M; p.n_items = 0;
  if (g.prec == 1) return mc ? launch_k(umma_gate_kernel<1, true>, p, grid, Cfg<1, false>::kSmemBytes, true, s)
                             : launch_k(umma_gate_kernel<1, false>, p, grid, Cfg<1, false>::kSmemBytes, false, s);
  if (g.prec == 2) return mc ? launch_k(umma_gate_kernel<2, true>, p, grid, Cfg<2, false>::kSmemBytes, true, s)
                             : launch_k(umma_gate_kernel<2, false>, p, grid, Cfg<2, false>::kSmemBytes, false, s);
  if (g.prec == 3) return mc ? launch_k(umma_gate_kernel<3, true>, p, grid, Cfg<3, false>::kSmemBytes, true, s)
                             : launch_k(umma_gate_kernel<3, false>, p, grid, Cfg<3, false>::kSmemBytes, false, s);
  return mc ? launch_k(umma_gate_kernel<0, true>, p, grid, Cfg<0, false>::kSmemBytes, true, s)
            : launch_k(umma_gate_kernel<0, false>, p, grid, Cfg<0, false>::kSmemBytes, false, s);
}

// Dilated conv (+ optional 1x1 conditioner term as extra K-slabs) with a linear fp32 epilogue: the training forward and its
// transposed form (dgrad), csrc/train.cu.  f16e5 operand pairs, one tile per CTA (pair).
int launch_umma_conv_lin(const UmmaConvLin& c, cudaStream_t s) {
  if ((c.Nout % TILE_N) || (c.Cin % TILE_K) || (c.taps & 1) == 0 || (c.Mp % TILE_K) || !c.out || (c.ldo & 3)) {
    set_error("umma_conv_lin: unsupported Cin=%d Nout=%d taps=%d Mp=%d", c.Cin, c.Nout, c.taps, c.Mp);
    return DRB_E_INVALID;
  }
  GateParams p;
  memset(&p, 0, sizeof(p));
  p.xh = *c.ah; p.xl = *c.al; p.wd_h = *c.wh; p.wd_l = *c.wl;
  p.zh = *c.ah; p.zl = *c.al; p.xwh = *c.ah; p.xwl = *c.al;      // never used by this epilogue; valid descriptors for the prefetch
  if (c.Mp > 0) { p.sh = *c.sh; p.sl = *c.sl; p.wc_h = *c.wch; p.wc_l = *c.wcl; }
  else { p.sh = *c.ah; p.sl = *c.al; p.wc_h = *c.wh; p.wc_l = *c.wl; }
  p.NB = c.NB; p.n_cond = c.Mp > 0 ? c.NB : 0; p.T = c.T; p.C = c.Cin; p.taps = c.taps; p.dil = c.dil;
  p.cond_slabs = c.Mp / TILE_K; p.tiles_t = (c.T + TILE_M - 1) / TILE_M; p.n_blocks = c.Nout / TILE_N; p.z_group0 = 0;
  p.bias_cond = c.bias; p.bias_unc = c.bias; p.inv_scale = c.inv_scale;
  p.cond = nullptr; p.n_cond_mma = p.n_cond; p.dual_off = 0; p.xs = nullptr; p.win_rows = TILE_M; p.n_items = 0;
  p.lin_out = c.out; p.ldo = c.ldo;
  p.tap_lo = c.tap_lo; p.tap_n = c.tap_n; p.lin_acc = c.accumulate; p.tap_span = c.tap_span;
  if (c.tap_n < 0 || c.tap_lo < 0 || c.tap_span < 0 || c.tap_lo + (c.tap_span > 0 ? c.tap_span : c.tap_n) > c.taps) {
    set_error("umma_conv_lin: bad tap range"); return DRB_E_INVALID;
  }
  const int grid = p.NB * p.tiles_t * p.n_blocks;
  const bool mc = c.pair && ((p.NB * p.tiles_t) % 2 == 0);
  static int pers = -1;   // DRB_LIN_PERS=0: the one-tile-per-CTA kernel (A/B runs)
  if (pers < 0) { const char* e = getenv("DRB_LIN_PERS"); pers = (e && e[0] == '0') ? 0 : 1; }
  if (mc && pers && (c.prec == 4 || c.prec == 3)) {   // persistent CTA pairs, every tap pass of a tile inside ONE launch
    int n_sm = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    p.n_items = (p.NB * p.tiles_t / 2) * p.n_blocks;
    const int pairs = p.n_items < n_sm / 2 ? p.n_items : n_sm / 2;
    // pass-granular ranges (a pair may start inside an item and park that part in the caller's scratch) when there is more than
    // one pass, the ranges are at least one item long (at most one partial per item) and the item count does not divide evenly
    const int per_pass = c.tap_n > 0 ? c.tap_n : c.taps, span = c.tap_span > 0 ? c.tap_span : per_pass;
    const int n_pass = (span + per_pass - 1) / per_pass;
    static int bal = -1;   // DRB_LIN_BALANCE=0: whole items per pair (A/B runs)
    if (bal < 0) { const char* e = getenv("DRB_LIN_BALANCE"); bal = (e && e[0] == '0') ? 0 : 1; }
    const bool split = bal && c.scratch && n_pass > 1 && (p.n_items % pairs) != 0 && (p.n_items * n_pass) / pairs >= n_pass &&
                       c.scratch_bytes >= (size_t)pairs * 2 * TILE_M * TILE_N * sizeof(float);
    p.lin_scratch = split ? c.scratch : nullptr;
    int r = c.prec == 4 ? launch_k(umma_conv_lin_pers_kernel<4>, p, 2 * pairs, CLP_SMEM, true, s)
                        : launch_k(umma_conv_lin_pers_kernel<3>, p, 2 * pairs, CLP_SMEM, true, s);
    if (r || !split) return r;
    lin_fixup_kernel<<<2 * pairs, 256, 0, s>>>(c.scratch, c.out, c.ldo, c.T, p.tiles_t, p.n_blocks, p.n_items, n_pass, pairs);
    DRB_LAUNCH_CHECK();
    return 0;
  }
  if (c.tap_span > 0 && c.tap_n > 0 && c.tap_span > c.tap_n) {   // one launch per pass, summed in out[] across launches
    UmmaConvLin one = c;
    one.tap_span = 0;
    for (int t = 0; t < c.tap_span; t += c.tap_n) {
      one.tap_lo = c.tap_lo + t; one.tap_n = t + c.tap_n <= c.tap_span ? c.tap_n : c.tap_span - t;
      one.accumulate = (t > 0 || c.accumulate) ? 1 : 0;
      one.bias = t == 0 ? c.bias : nullptr;
      one.Mp = t + c.tap_n >= c.tap_span ? c.Mp : 0;
      const int r = launch_umma_conv_lin(one, s);
      if (r) return r;
    }
    return 0;
  }
  if (c.tap_span > 0) p.tap_n = c.tap_span;
  if (c.prec == 4) return mc ? launch_k(umma_gate_kernel<4, true>, p, grid, Cfg<4, false>::kSmemBytes, true, s)
                             : launch_k(umma_gate_kernel<4, false>, p, grid, Cfg<4, false>::kSmemBytes, false, s);
  return mc ? launch_k(umma_gate_kernel<3, true>, p, grid, Cfg<3, false>::kSmemBytes, true, s)
            : launch_k(umma_gate_kernel<3, false>, p, grid, Cfg<3, false>::kSmemBytes, false, s);
}

int launch_umma_zgemm(const UmmaMaps& maps, const UmmaZGemm& z, cudaStream_t s) {
  if (z.C % TILE_N) { set_error("umma_zgemm: unsupported C=%d", z.C); return DRB_E_INVALID; }
  ZGemmParams p;
  p.zh = maps.zh; p.zl = maps.zl; p.w_h = *z.w_h; p.w_l = *z.w_l; p.out32 = *z.out32; p.xh = maps.xh; p.xl = maps.xl;
  p.NB = z.NB; p.T = z.T; p.C = z.C; p.tiles_t = (z.T + TILE_M - 1) / TILE_M; p.n_blocks = z.C / TILE_N;
  p.spg = z.C / TILE_K; p.nslabs = z.groups * p.spg;
  if (z.mode == 2 || z.mode == 4) {   // A = caller's operand pair (2: spectrogram; 4: any), C counts the output columns, K = nslabs64 slabs
    if (!z.a_h || !z.a_l || z.nslabs64 <= 0) { set_error("umma_zgemm: mode %d needs the A operand maps", z.mode); return DRB_E_INVALID; }
    p.zh = *z.a_h; p.zl = *z.a_l; p.spg = z.nslabs64; p.nslabs = z.nslabs64;
  } p.z_group0 = z.z_group0; p.group_stride = z.group_stride;
  p.mode = z.mode; p.bias = z.bias; p.dnext = z.dnext; p.steps = z.steps; p.t_uniform = z.t_uniform; p.bsamp = z.bsamp > 0 ? z.bsamp : 1;
  p.range_max = z.range_max; p.xs = nullptr;
  p.ksplit = 0;
  if (z.ksplit) {
    if (z.mode != 4 || !z.pair || (p.tiles_t & 1)) { set_error("umma_zgemm: split-K needs mode 4 on CTA pairs with an even row-tile count"); return DRB_E_INVALID; }
    p.ksplit = 1;
  }
  p.h_pair = 0; p.dual_B = 0; p.F = 0; p.x_t = nullptr; p.noise = nullptr; p.x_prev = nullptr; p.net_out = nullptr;
  p.upd.mode = DRB_UPD_NONE; p.upd.has_noise = 0; p.upd.w = 0.f;
  if (z.mode == 1 && z.hp_h && z.hp_l) { p.h_pair = 1; p.xh = *z.hp_h; p.xl = *z.hp_l; }
  int grid = p.NB * p.tiles_t * p.n_blocks;
  p.inv_scale = z.inv_scale;
  bool mc = z.pair && ((p.NB * p.tiles_t) % 2 == 0);
  if (z.mode == 3) {   // head output projection: A = h pair (K = C), one N block of padded output rows
    if (!z.a_h || !z.a_l || !z.upd || !z.x_prev || z.F <= 0 || z.F > TILE_N || (z.F & 3)) { set_error("umma_zgemm: mode 3 arguments"); return DRB_E_INVALID; }
    p.zh = *z.a_h; p.zl = *z.a_l; p.n_blocks = 1; p.z_group0 = 0; p.group_stride = 0;
    p.F = z.F; p.upd = *z.upd; p.x_t = z.x_t; p.noise = z.noise; p.x_prev = z.x_prev; p.net_out = z.net_out;
    if (z.dual_B > 0) {   // pair = (conditional roll, unconditional roll) of the same frames
      if (!z.pair || z.NB != 2 * z.dual_B) { set_error("umma_zgemm: mode 3 guidance pair needs CTA pairs"); return DRB_E_INVALID; }
      p.dual_B = z.dual_B; mc = true; grid = 2 * z.dual_B * p.tiles_t;
    } else {
      grid = p.NB * p.tiles_t;
    }
  }
  if (mc && z.persistent && z.mode == 0 && (z.prec == 1 || z.prec == 3)) {
    int n_sm = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    const int n_items = (p.NB * p.tiles_t / 2) * p.n_blocks;
    const int pairs = n_items < n_sm / 2 ? n_items : n_sm / 2;
    if (z.x_n4) {
      if (z.prec != 3 || !z.xl4 || !z.xs) { set_error("umma_zgemm: f16n4 activations need f16e5 z operands"); return DRB_E_INVALID; }
      p.xl = *z.xl4; p.xs = z.xs;
      return launch_k(umma_res_pers_kernel<3, 4>, p, 2 * pairs, RP_SMEM, true, s);
    }
    return z.prec == 1 ? launch_k(umma_res_pers_kernel<1, 0>, p, 2 * pairs, RP_SMEM, true, s)
                       : launch_k(umma_res_pers_kernel<3, 0>, p, 2 * pairs, RP_SMEM, true, s);
  }
  if (mc && z.persistent && z.mode == 1 && p.h_pair && (z.prec == 1 || z.prec == 3)) {   // persistent HEAD (operand-pair output)
    int n_sm = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    const int n_items = (p.NB * p.tiles_t / 2) * p.n_blocks;
    const int pairs = n_items < n_sm / 2 ? n_items : n_sm / 2;
    return z.prec == 1 ? launch_k(umma_head_pers_kernel<1>, p, 2 * pairs, HP_SMEM, true, s)
                       : launch_k(umma_head_pers_kernel<3>, p, 2 * pairs, HP_SMEM, true, s);
  }
  if (z.x_n4 && z.mode == 0) { set_error("umma_zgemm: f16n4 needs the persistent CTA-pair kernels (even tile count)"); return DRB_E_INVALID; }
  if (z.prec == 1) return mc ? launch_k(umma_zgemm_kernel<1, true>, p, grid, Cfg<1, false>::kSmemBytes, true, s)
                             : launch_k(umma_zgemm_kernel<1, false>, p, grid, Cfg<1, false>::kSmemBytes, false, s);
  if (z.prec == 2) return mc ? launch_k(umma_zgemm_kernel<2, true>, p, grid, Cfg<2, false>::kSmemBytes, true, s)
                             : launch_k(umma_zgemm_kernel<2, false>, p, grid, Cfg<2, false>::kSmemBytes, false, s);
  if (z.prec == 3) return mc ? launch_k(umma_zgemm_kernel<3, true>, p, grid, Cfg<3, false>::kSmemBytes, true, s)
                             : launch_k(umma_zgemm_kernel<3, false>, p, grid, Cfg<3, false>::kSmemBytes, false, s);
  return mc ? launch_k(umma_zgemm_kernel<0, true>, p, grid, Cfg<0, false>::kSmemBytes, true, s)
            : launch_k(umma_zgemm_kernel<0, false>, p, grid, Cfg<0, false>::kSmemBytes, false, s);
}

}  // namespace drb

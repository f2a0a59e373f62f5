// Host-side launchers shared between the translation units of libdiffroll_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include "../../include/diffroll_b200.h"

namespace drb {

// ------------------------------- fp32 CUDA-core GEMM (simt_kernels.cu) -------------------------
// C[m][n] = epi( sum_k Aeff[m][k] * W[n][k] + bias[n] ),  m in [0,M), n in [0,N)
//   Aeff[m][k]: k = tap*Ck + c;  row m = seg*T + t;  source row t' = t + (tap - taps/2)*dil
//               Aeff = (alpha*A[(seg*T+t')*lda + c] + beta*A2[...]) / a_div + addvec[c]   if 0<=t'<T else 0
struct SimtGemm {
  const float* A = nullptr;
  const float* A2 = nullptr;  // optional second source (guidance combine), same indexing
  float alpha = 1.f, beta = 0.f, a_div = 1.f;  // Aeff = (alpha*A + beta*A2) / a_div + addvec
  const float* addvec = nullptr;  // [Ck], optional
  // per-segment addvec row (per-sample diffusion steps): row = addvec + addvec_steps[seg % addvec_mod] * addvec_stride
  const int* addvec_steps = nullptr;
  int addvec_mod = 1, addvec_stride = 0;
  int lda = 0;
  int T = 1;  // rows per segment (conv boundary)
  int taps = 1, dil = 1, Ck = 0;
  const float* W = nullptr;
  int ldw = 0;
  const float* bias = nullptr;
  int act = 0;         // 0 none, 1 relu, 2 silu
  int accumulate = 0;  // C += ...
  int skinny = 0;      // M <= 32 rows: one warp per output column (training: per-roll projections); summation order differs from the tiled kernel
  float* C = nullptr;
  int ldc = 0;
  int M = 0, N = 0;
  // optional fused posterior update (out = upd(net, x_t, noise)); net written to net_out if non-null
  const drb_update* upd = nullptr;  // host pointer, copied by value
  const float* x_t = nullptr;
  const float* noise = nullptr;
  float* net_out = nullptr;
};
int launch_simt_gemm(const SimtGemm& g, cudaStream_t s);

// elementwise (simt_kernels.cu)
int launch_gate(const float* y, float* z, int M, int C, cudaStream_t s);  // z = sigmoid(y[:, :C]) * tanh(y[:, C:])
int launch_res_skip(const float* o, float* x, float* skip, int M, int C, int first, int do_res, cudaStream_t s);
// x32 rows [0,Mb) hold relu(in_proj); duplicate them `copies` times (branches) and emit the tensor-path operand of x + d
// fmt: 0 none, 1 bf16 hi/lo, 2 fp16 + e4m3 correction bytes
int launch_prep_xin(float* x32, void* xmain, void* xaux, const float* dvec, int Mb, int C, int copies, int fmt, cudaStream_t s);
// tensor path: relu(input_projection(x_t)) -> fp32 x32 for every branch copy + operand pair of x + dtab0[t] in ONE kernel;
// t = steps[row / T] when steps != nullptr (per-sample diffusion steps), else the uniform t
// fmt 4 (f16n4): xaux = e2m1 bytes [rows][C], xsf = scale factors [roll][C/64][T][8]
int launch_in_proj_fused(const float* x_t, const float* W, const float* bias, const float* dtab0, const int* steps, int t,
                         int M, int T, int F, int C, int copies, int fmt, float* x32, void* xmain, void* xaux, uint8_t* xsf,
                         unsigned int* range_max, cudaStream_t s);
// f16n4 weights: w [OC][K] fp32 (tap-major K, K % 64 == 0), rows permuted into 256-wide gate/filter blocks (C > 0) ->
// main fp16(W * SW) [OC][K], aux e2m1 bytes [OC][K] ([hi part 32 B | lo part 32 B] per 64-wide K-slab) and the scale atoms
// [OC / 256][K / 64][2 instructions][2 atoms][512 B]; SW = scale[0]
int launch_repack_n4(const float* w, void* mainp, void* auxp, uint8_t* sf_atoms, int OC, int K, int interleave_C,
                     const float* scale, cudaStream_t s);
// weight repacks
int launch_repack_conv_fp32(const float* w, float* out, int OC, int C, int k, cudaStream_t s);  // [OC][C][k] -> [OC][k][C]
// [OC][Kin] fp32 (k index already tap-major) -> operand pair (fmt 1 or 2), rows optionally permuted into 256-wide
// gate/filter blocks (interleave_C > 0), K padded to Kp; fmt 2 and 3 read SW from scale[0]
int launch_repack_split(const float* w, void* mainp, void* auxp, int OC, int Kin, int Kp, int interleave_C, int fmt,
                        const float* scale, cudaStream_t s);
// scale2[0] = SW = 2^floor(log2(224 / max|w|)), scale2[1] = 1/(sa*SW) from max|w| over one or two tensors (scale2 must
// have 3 floats of space).  sa: F8_SA for f16f8 (scale of the activation residual), 1 for f16e5.
int launch_weight_scale(const float* w0, size_t n0, const float* w1, size_t n1, float* scale2, float sa, cudaStream_t s,
                        float target = 224.f);
int launch_pad_rows(const float* src, float* dst, int rows, int Kin, int Kp, cudaStream_t s);
// bias1[n'] (interleaved) = bd[n] + bc[n] (- sum_k Wc[n][k] if uncond)
int launch_bias1(const float* bd, const float* bc, const float* wc, float* out_cond, float* out_unc,
                 float* out_cond_nat, float* out_unc_nat, int C, int n_mels, cudaStream_t s);
int launch_fill(float* p, float v, size_t n, cudaStream_t s);
// Wcomp[n][l*C + k] = sum_j Ws[n][j] * Wo_l[C + j][k] / sqrt(L)   (one layer l per launch)
int launch_compose_skip(const float* ws, const float* wo, float* wcomp, int C, int L, int layer, cudaStream_t s);
// bcomp[n] = bs[n] + sum_j Ws[n][j] * bsum[j] / sqrt(L), bsum = sum_l bo_l[C + j]
int launch_compose_bias(const float* ws, const float* bs, const float* const* bo_dev_ptrs_host, int C, int L, float* bsum_tmp,
                        float* bcomp, cudaStream_t s);

// ------------------------------- mel front-end (mel.cu) ----------------------------------------
struct MelPlan;
size_t mel_workspace_bytes(const drb_config& c);
int mel_create(MelPlan** mp, const drb_config& c, const float* window, const float* fb, void* ws, size_t ws_bytes,
               cudaStream_t s);
void mel_destroy(MelPlan* mp);
// logmel -> normalised -> masked.  spec_out [B][n_mels][T] fp32 (nullable), spec32 [B][T][Mp] fp32 zero-padded,
// spec_main/spec_aux: tensor-path operand pair of the spectrogram (fmt 1: bf16 hi/lo [B][T][Mp]; fmt 2: fp16 + e4m3 bytes)
int mel_forward(MelPlan* mp, const float* waveform, float* spec_out, float* spec32, void* spec_main, void* spec_aux, int fmt,
                int Mp, int T, int it0, int it1, int if0, int if1, cudaStream_t s);
float* mel_logmel_ptr(MelPlan* mp, size_t* bytes);

// ------------------------------- tcgen05 GEMMs (umma_gemm.cu) ----------------------------------
struct UmmaLayer {
  CUtensorMap wd_h, wd_l;  // [2C rows (interleaved)][k*C] bf16
  CUtensorMap wc_h, wc_l;  // [2C][Mp]
  CUtensorMap wo_h, wo_l;  // [2C][C]
};
// Posterior update of one element, same operation order as the reference's fp32 tensor expressions (DRB_UPD_* in diffroll_b200.h).
#ifdef __CUDACC__
__device__ __forceinline__ float posterior_update(const drb_update& u, float net, float x, float n) {
  switch (u.mode) {
    case DRB_UPD_X0: {
      float r = u.s[0] * net + u.s[1] * (x - u.s[2] * net) / u.s[3];
      return u.has_noise ? r + u.s[4] * n : r;
    }
    case DRB_UPD_X0_FINAL: return net / u.s[0];
    case DRB_UPD_EPS_DDPM: {
      float r = u.s[0] * (x - u.s[1] * net / u.s[2]);
      return u.has_noise ? r + u.s[3] * n : r;
    }
    case DRB_UPD_EPS_DDIM: {
      float r = u.s[0] * ((x - u.s[1] * net) / u.s[2]) + u.s[3] * net;
      return u.has_noise ? r + u.s[4] * n : r;
    }
    case DRB_UPD_EPS_FINAL: return (x - u.s[0] * net) / u.s[1];
    default: return net;
  }
}
#endif

struct UmmaMaps {
  CUtensorMap xh, xl;      // [NB][T][C]
  CUtensorMap zh, zl;      // [L*NB][T][C]: gated activations of every layer
  CUtensorMap x32, h32;    // fp32 [NB][T][C], box 32 channels x 128 frames
  CUtensorMap wcomp_h, wcomp_l;  // [C][L*C] composed skip/head weights
  CUtensorMap sh, sl;      // [B][T][Mp]
  CUtensorMap hh, hl;      // [NB][T][C]: operand pair of the head's hidden activations relu(skip_projection(...)) (tensor-core head)
  CUtensorMap wout_h, wout_l;  // [256][C]: output_projection rows padded to one N block
};
struct UmmaGate {  // y = conv(xin) + cond + bias1 ; z = sigmoid(gate)*tanh(filter) -> zh/zl[z_group0 + roll]
  int NB, n_cond, T, C, taps, dil, Mp, prec, z_group0;  // prec: 0 bf16, 1 bf16x3, 2 f16f8
  int pair = 1;  // CTA pairs (cta_group::2, M = 256) when the M-tile count is even
  int persistent = 1;  // persistent CTA pairs with two TMEM accumulator stages (bf16x3 / f16e5 only)
  int window = 1;  // fetch the tap window once per channel chunk (needs pair, aux operands and the window maps)
  const CUtensorMap *xwh = nullptr, *xwl = nullptr;  // activation maps with box rows = 128 + (taps-1)*dil
  const float* inv_scale;
  const float* bias_cond;  // interleaved [2C]
  const float* bias_unc;
  // Persistent kernel only.  cond: this layer's conditioner projection of the spectrogram, fp32 [n_cond][T][2C], computed
  // once per clip and added in the epilogue of conditional rolls (nullptr = contract it as K-slabs every step).
  // dual_B > 0 (layer 0, both branches read the same x): conv once over dual_B rolls, two gated outputs per tile.
  const float* cond = nullptr;
  int dual_B = 0;
  int need_tables = 0;   // the conditioner term exists ONLY as table rows (learned clips): fail instead of contracting K-slabs
  // f16n4 (prec must be 3 for the z output): fp16 main + block-scaled e2m1 correction operands
  int n4 = 0;
  const CUtensorMap *xw4 = nullptr, *wd4 = nullptr, *wsf = nullptr;   // aux window map, aux weight map, weight scale atoms
  const uint8_t* xs = nullptr;                                        // activation scale factors [roll][chunk][frame][8]
};
struct UmmaZGemm {  // A = stored z (groups x C channels of K), B = w maps, fp32 output tile through `out32`
  int pair = 1, persistent = 1;
  int NB, T, C, prec, mode;         // mode 0: residual update (x32 in place, xh/xl of x + dnext); 1: relu -> h;
                                    // 2: conditioner tables: A = spectrogram pair (K = nslabs64 * 64), N = 2C gate-interleaved weight
                                    //    rows, plain fp32 result stored in NATURAL channel order [gate 0..C-1 | filter C..2C-1]
  int nslabs64 = 0;                 // mode 2: K-slabs (Mp / 64)
  int ksplit = 0;                   // mode 4: split-K -- NB = number of K ranges (nslabs64 slabs each); split i writes out32[i][row][n]
                                    // 3: head output projection, see below
                                    // 4: plain GEMM out32[row][n] = sum_k A[row][k] W[n][k] (A = a_h/a_l maps [NB][T rows][K], C = output
                                    //    columns, K = nslabs64 * 64), natural column order, no activation (training: weight gradients)
  const float* inv_scale;
  int groups, z_group0, group_stride;
  const CUtensorMap *w_h, *w_l, *out32;
  const float* bias;
  const float* dnext;               // RES: diffusion_projection table of the NEXT layer, [timesteps][C]
  int t_uniform = 0;                // row of dnext when steps == nullptr
  const int* steps = nullptr;       // per-sample diffusion steps [bsamp] (device); roll nb uses steps[nb % bsamp]
  int bsamp = 1;
  unsigned int* range_max = nullptr;  // RES: device word, atomicMax of |x + d_next| (fp32 bits) over the emitted operands
  const CUtensorMap *a_h = nullptr, *a_l = nullptr;   // mode 2 / 3: A operand maps (spectrogram pair / head hidden pair) instead of the stored z
  // mode 1 with h_pair: the ReLU output leaves as an operand pair through xh/xl-style maps instead of fp32 through out32
  const CUtensorMap *hp_h = nullptr, *hp_l = nullptr;
  // mode 3: output_projection (N = pitches <= 256 padded) + guidance combine + posterior update     model/diffwave.py:685, task/diffusion.py:1009-1023
  //   dual_B > 0: CTA pair = (conditional roll b, unconditional roll b + dual_B) of the same frames; the pair's second CTA hands
  //   its accumulator to the first one through distributed shared memory and the first one writes x_prev.
  int dual_B = 0, F = 0;
  const drb_update* upd = nullptr;    // host pointer, copied by value
  const float* x_t = nullptr; const float* noise = nullptr; float* x_prev = nullptr; float* net_out = nullptr;
  int x_n4 = 0;                       // RES emits the next layer's operand in the f16n4 format (xl4 = 64-byte-row aux map, xs = scales)
  const CUtensorMap* xl4 = nullptr;
  uint8_t* xs = nullptr;
};
struct UmmaConvLin {  // out[roll][t][n] = sum_{tap,c} A[roll][t + (tap - taps/2) dil][c] W[n][tap*Cin + c] (+ sum_m S[roll][t][m] Wc[n][m]) + bias[n]
  const CUtensorMap *ah = nullptr, *al = nullptr;    // activation pair [NB][T][Cin] (f16e5)
  const CUtensorMap *wh = nullptr, *wl = nullptr;    // weight pair [Nout][taps*Cin], natural row order
  const CUtensorMap *sh = nullptr, *sl = nullptr, *wch = nullptr, *wcl = nullptr;   // optional 1x1 term: [NB][T][Mp] and [Nout][Mp]
  int NB = 0, T = 0, Cin = 0, Nout = 0, taps = 1, dil = 1, Mp = 0, pair = 1;
  int tap_lo = 0, tap_n = 0, accumulate = 0;          // taps [tap_lo, tap_lo + tap_n) only (0: all); out += result
  float* scratch = nullptr; size_t scratch_bytes = 0; // optional: lets the persistent kernel cut its work between passes of an item (pairs x 256 KB)
  int tap_span = 0;                                  // > 0: taps [tap_lo, tap_lo + tap_span) in passes of tap_n taps, passes summed in out[] in fp32
  int prec = 3;                                      // 3 f16e5 pairs (aux = bytes [.][2C]); 4 f16x3 (aux = fp16 lo [.][C]): fp32-grade
  const float* inv_scale = nullptr;                  // device scalar: 1 / (product of the operand scales)
  const float* bias = nullptr;                       // [Nout] or nullptr
  float* out = nullptr; int ldo = 0;
};
int launch_umma_conv_lin(const UmmaConvLin& c, cudaStream_t s);
// fp32 rows (+ optional per-segment addvec, * optional device scale) -> f16e5 operand pair [M][C] (main fp16, aux bytes [M][2C])
// fmt 3: f16e5; fmt 5: f16x3 (fp16 hi + fp16 lo, aux [M][C] halves)
int launch_split_pair(const float* src, int ld, const float* addvec, int av_stride, int T, const float* scale, void* mainp, void* auxp,
                      int M, int C, cudaStream_t s, int fmt = 3);
// fp32 rows -> TRANSPOSED f16e5 operand pair (training weight gradients contract over rolls x frames): out row tap*C + c holds
// src[m + (tap - taps/2) * dil][c] (+ addvec[roll][c]) * scale for m = 0..M-1, zero outside the roll; main [taps*C][M] fp16,
// aux [taps*C][2M] bytes ([lo x 64 | hi x 64] per 64 rows m).  M % 64 == 0, C % 64 == 0, T % 64 == 0.
// bside: the pair is the GEMM's B ("weight") operand ([hi * 2^-4 | lo * 2^8]) instead of the A form ([lo * 2^4 | hi * 2^-8]).
int launch_split_pair_T(const float* src, int ld, const float* addvec, int av_stride, int T, int taps, int dil, const float* scale,
                        void* mainp, void* auxp, int M, int C, int bside, cudaStream_t s);
int umma_init();  // resolves cuTensorMapEncodeTiled
// dtype: 0 bf16, 1 fp32, 2 fp16, 3 uint8; the box always spans 128 bytes of the innermost dimension
int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, int dtype);
int make_tmap_3d(CUtensorMap* m, const void* base, uint64_t d2, uint64_t d1, uint64_t d0, uint32_t box1, int dtype);
// byte tensors, explicit inner box width (bytes) and swizzle (128 / 64 / 0)
int make_tmap_2d_bytes(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols, int swizzle_bytes);
int make_tmap_3d_bytes(CUtensorMap* m, const void* base, uint64_t d2, uint64_t d1, uint64_t d0, uint32_t box1, uint32_t box0, int swizzle_bytes);
int launch_umma_gate(const UmmaMaps& maps, const UmmaLayer& L, const UmmaGate& g, cudaStream_t s);
int launch_umma_zgemm(const UmmaMaps& maps, const UmmaZGemm& z, cudaStream_t s);

}  // namespace drb

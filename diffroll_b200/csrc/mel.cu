// STFT -> power -> HTK mel -> log -> per-clip min-max -> mask.  Replaces, for one clip batch,
//   torchaudio MelSpectrogram(...)            model/diffwave.py:635,643   (third-party torchaudio 0.11.0)
//   torch.log(spec + 1e-6)                    model/diffwave.py:644
//   Normalization(0,1,'imagewise')            model/diffwave.py:645, model/utils.py:21-32
//   inpainting masks                          model/diffwave.py:649-654
// n_fft = 2048 (the reference's configuration): ONE kernel per clip batch, a CTA per frame -- reflect-padded, windowed frame
// straight from the waveform into shared memory, a 1024-point complex Stockham radix-4 FFT there (the 2048 real samples packed
// as 1024 complex ones, untangled afterwards), power, banded mel filter, log, min-max: neither the frames (168 MB at 32 clips)
// nor the complex spectrum (168 MB) exist in HBM.  Other n_fft: cuFFT (R2C, batched) between a framing kernel and the
// power/mel consumer kernel.  The mel "matmul" uses the band structure of the filterbank (<= ~24 non-zeros per mel bin)
// instead of a dense [1025 x 229] product.
#include <cufft.h>
#include <limits.h>
#include "common.cuh"
#include "kernels.h"

namespace drb {

struct MelPlan {
  int B, L, n_fft, hop, n_mels, nF, nbins;
  cufftHandle fft;
  bool fft_ok;
  const float* window;  // caller-owned (state_dict buffer)
  const float* fb;      // caller-owned [nbins][n_mels]
  float* frames;        // [B*nF][n_fft]
  float2* spectrum;     // [B*nF][nbins]
  float* logmel;        // [B][nF][n_mels]
  int* band_start;      // [n_mels]
  int* band_len;        // [n_mels]
  float* wnorm;         // [1] sqrt(sum w^2)
  int* minmax;          // [B][2] order-preserving int encodings
  void* fft_work;
  size_t fft_work_bytes;
  bool fused;           // n_fft == 2048: stft_mel_kernel, no frames / spectrum / cuFFT plan
  float2* twiddle;      // [2048] exp(-2 pi i k / 2048)
  float* fbT;           // [MEL_BAND_MAX][MEL_PITCH] compact filterbank (fused kernel)
};

static bool mel_fused(int n_fft, int n_mels = 1) {
  static int off = -1;   // DRB_MEL_CUFFT=1: the cuFFT pipeline also for n_fft = 2048 (A/B runs)
  if (off < 0) { const char* e = getenv("DRB_MEL_CUFFT"); off = (e && e[0] == '1') ? 1 : 0; }
  return n_fft == 2048 && n_mels <= 256 && !off;
}

static size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

static size_t fft_work_estimate(int n_fft, int batch) {
  // Deterministic (no CUDA context needed): one spectrum-sized scratch.  mel_create falls back to cuFFT-owned
  // scratch if the real requirement turns out larger.
  return align_up((size_t)batch * (n_fft / 2 + 1) * sizeof(float2) + (1 << 20));
}

size_t mel_workspace_bytes(const drb_config& c) {
  const int nF = c.wave_len / c.hop_length + 1, nbins = c.n_fft / 2 + 1;
  const size_t rows = (size_t)c.batch * nF;
  const bool fused = mel_fused(c.n_fft, c.n_mels);
  size_t b = 0;
  if (!fused) {
    b += align_up(rows * c.n_fft * sizeof(float));
    b += align_up(rows * nbins * sizeof(float2));
  }
  b += align_up(rows * c.n_mels * sizeof(float));
  b += align_up(c.n_mels * sizeof(int)) * 2;
  b += align_up(sizeof(float));
  b += align_up((size_t)c.batch * 2 * sizeof(int));
  b += fused ? align_up(2048 * sizeof(float2)) + align_up((size_t)64 * 256 * sizeof(float)) : fft_work_estimate(c.n_fft, (int)rows);
  return b;
}

__global__ void twiddle_kernel(float2* __restrict__ W) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= 2048) return;
  double sn, cs;
  sincospi(-2.0 * (double)k / 2048.0, &sn, &cs);
  W[k] = make_float2((float)cs, (float)sn);
}

__global__ void band_kernel(const float* __restrict__ fb, int nbins, int n_mels, int* __restrict__ start,
                            int* __restrict__ len) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_mels) return;
  int first = -1, last = -1;
  for (int f = 0; f < nbins; ++f)
    if (fb[(size_t)f * n_mels + m] != 0.f) { if (first < 0) first = f; last = f; }
  start[m] = first < 0 ? 0 : first;
  len[m] = first < 0 ? 0 : last - first + 1;
}

__global__ void wnorm_kernel(const float* __restrict__ w, int n, float* __restrict__ out) {
  // window.pow(2.).sum().sqrt()   (torchaudio functional.spectrogram, normalized="window")
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)w[i] * (double)w[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) out[0] = (float)sqrt(red[0]);
}

__global__ void frame_kernel(const float* __restrict__ wave, const float* __restrict__ window, float* __restrict__ frames,
                             int L, int n_fft, int hop, int nF) {
  // center=True, pad_mode='reflect': padded[p] = wave[reflect(p - n_fft/2)]; frame t covers p in [t*hop, t*hop+n_fft)
  const int t = blockIdx.x, b = blockIdx.y;
  const float* w = wave + (size_t)b * L;
  float* out = frames + ((size_t)b * nF + t) * n_fft;
  const int base = t * hop - n_fft / 2;
  for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
    int i = base + n;
    if (i < 0) i = -i;
    if (i >= L) i = 2 * (L - 1) - i;
    out[n] = window[n] * w[i];
  }
}

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void minmax_init_kernel(int* mm, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) { mm[2 * i] = INT_MAX; mm[2 * i + 1] = INT_MIN; }
}

__global__ void __launch_bounds__(256) power_mel_kernel(const float2* __restrict__ spectrum, const float* __restrict__ fb,
                                                        const int* __restrict__ band_start, const int* __restrict__ band_len,
                                                        const float* __restrict__ wnorm, float* __restrict__ logmel,
                                                        int* __restrict__ minmax, int nbins, int n_mels, int nF) {
  extern __shared__ float pw[];  // [nbins]
  __shared__ int smin[8], smax[8];
  const int t = blockIdx.x, b = blockIdx.y;
  const float2* sp = spectrum + ((size_t)b * nF + t) * nbins;
  const float s = wnorm[0];
  for (int f = threadIdx.x; f < nbins; f += blockDim.x) {
    float2 v = sp[f];
    float re = v.x / s, im = v.y / s;  // spec_f /= window.pow(2).sum().sqrt()
    pw[f] = re * re + im * im;         // .abs().pow(2)
  }
  __syncthreads();
  int lmin = INT_MAX, lmax = INT_MIN;
  for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
    const int f0 = band_start[m], n = band_len[m];
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc = fmaf(pw[f0 + i], fb[(size_t)(f0 + i) * n_mels + m], acc);
    float v = logf(acc + 1e-6f);  // torch.log(spec + 1e-6)
    logmel[((size_t)b * nF + t) * n_mels + m] = v;
    if (v == v) { int o = f2ord(v); lmin = min(lmin, o); lmax = max(lmax, o); }
  }
  for (int o = 16; o > 0; o >>= 1) {
    lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
    lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lmin; smax[threadIdx.x >> 5] = lmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) { lmin = min(lmin, smin[i]); lmax = max(lmax, smax[i]); }
    atomicMin(&minmax[2 * b], lmin);
    atomicMax(&minmax[2 * b + 1], lmax);
  }
}

// Compact copy of the filterbank for the fused kernel: fbT[i][m] = i-th non-zero weight of mel bin m (n_mels <= MEL_PITCH), so the
// threads of a warp (one mel bin each) read one coalesced row per step of their band walk.
constexpr int MEL_BAND_MAX = 64, MEL_PITCH = 256;
__global__ void band_pack_kernel(const float* __restrict__ fb, const int* __restrict__ start, const int* __restrict__ len, int n_mels,
                                 float* __restrict__ fbT) {
  const int m = blockIdx.x, i = threadIdx.x;   // MEL_BAND_MAX threads
  fbT[(size_t)i * MEL_PITCH + m] = i < len[m] ? fb[(size_t)(start[m] + i) * n_mels + m] : 0.f;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// One frame per CTA (256 threads): frame t of clip b -> log-mel row [n_mels] + the clip's running min / max.
//   z[c] = w[2c] x[2c] + i w[2c+1] x[2c+1], c < 1024;  Z = FFT_1024(z): Stockham autosort, radix 4, five stages, one butterfly per
//   thread and stage, ping-pong between two shared buffers (natural order out, no bit reversal);
//   X[k] = E[k] + W^k O[k], E = (Z[k] + conj Z[1024-k]) / 2, O = (Z[k] - conj Z[1024-k]) / 2i, k = 0 .. 1024  (W = exp(-2 pi i / 2048)).
__global__ void __launch_bounds__(256) stft_mel_kernel(const float* __restrict__ wave, const float* __restrict__ window,
                                                       const float2* __restrict__ W, const float* __restrict__ fbT, const float* __restrict__ fb,
                                                       const int* __restrict__ band_start, const int* __restrict__ band_len,
                                                       const float* __restrict__ wnorm, float* __restrict__ logmel, int* __restrict__ minmax,
                                                       int L, int hop, int nF, int n_mels, int vec_ok) {
  __shared__ float2 bufA[1024], bufB[1024];
  __shared__ float pw[1025];
  __shared__ int smin[8], smax[8];
  const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* w = wave + (size_t)b * L;
  const int base = t * hop - 1024;   // center=True, pad_mode='reflect': padded[p] = wave[reflect(p - n_fft/2)]
  const bool interior = vec_ok && base >= 0 && base + 2048 <= L;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = tid + 256 * j;
    float x0, x1;
    if (interior) {
      const float2 v = *reinterpret_cast<const float2*>(w + base + 2 * c);
      x0 = v.x; x1 = v.y;
    } else {
      int i0 = base + 2 * c, i1 = i0 + 1;
      if (i0 < 0) i0 = -i0;
      if (i0 >= L) i0 = 2 * (L - 1) - i0;
      if (i1 < 0) i1 = -i1;
      if (i1 >= L) i1 = 2 * (L - 1) - i1;
      x0 = w[i0]; x1 = w[i1];
    }
    const float2 wn = *reinterpret_cast<const float2*>(window + 2 * c);
    bufA[c] = make_float2(wn.x * x0, wn.y * x1);
  }
  __syncthreads();
  float2* x = bufA;
  float2* y = bufB;
#pragma unroll
  for (int st = 0; st < 5; ++st) {
    const int s = 1 << (2 * st), m = 256 >> (2 * st);           // stride, quarter length of this stage's sub-transforms
    const int p = tid >> (2 * st), q = tid & (s - 1);
    const float2 a = x[q + s * p], bq = x[q + s * (p + m)], c = x[q + s * (p + 2 * m)], d = x[q + s * (p + 3 * m)];
    const int idx = p << (2 * st + 1);                          // exp(-2 pi i p / n) = W[p * 2048 / n]
    const float2 apc = make_float2(a.x + c.x, a.y + c.y), amc = make_float2(a.x - c.x, a.y - c.y);
    const float2 bpd = make_float2(bq.x + d.x, bq.y + d.y), jbmd = make_float2(-(bq.y - d.y), bq.x - d.x);   // j (b - d)
    float2 y0 = make_float2(apc.x + bpd.x, apc.y + bpd.y);
    float2 y1 = make_float2(amc.x - jbmd.x, amc.y - jbmd.y);
    float2 y2 = make_float2(apc.x - bpd.x, apc.y - bpd.y);
    float2 y3 = make_float2(amc.x + jbmd.x, amc.y + jbmd.y);
    if (st < 4) {                                               // last stage: n = 4, p = 0, all twiddles 1
      y1 = cmul(__ldg(W + idx), y1); y2 = cmul(__ldg(W + 2 * idx), y2); y3 = cmul(__ldg(W + 3 * idx), y3);
    }
    const int o = q + s * 4 * p;
    y[o] = y0; y[o + s] = y1; y[o + 2 * s] = y2; y[o + 3 * s] = y3;
    __syncthreads();
    float2* tmp = x; x = y; y = tmp;
  }
  const float inv_sN = 1.f / wnorm[0];
  for (int k = tid; k <= 1024; k += 256) {
    const float2 zk = x[k & 1023], zr = x[(1024 - k) & 1023];
    const float2 e = make_float2(0.5f * (zk.x + zr.x), 0.5f * (zk.y - zr.y));        // (Z[k] + conj Z[N-k]) / 2
    const float2 o = make_float2(0.5f * (zk.y + zr.y), -0.5f * (zk.x - zr.x));       // (Z[k] - conj Z[N-k]) / 2i
    const float2 wo = cmul(__ldg(W + k), o);
    const float re = (e.x + wo.x) * inv_sN, im = (e.y + wo.y) * inv_sN;              // spec_f /= window.pow(2).sum().sqrt()
    pw[k] = re * re + im * im;                                                       // .abs().pow(2)
  }
  __syncthreads();
  // one thread per mel bin walking its band: weight i of every bin sits in row i of the compact table, so a warp's loads are one
  // coalesced row per step (the full matrix gave every thread its own cache line per frequency: 165 us of L2 latency per clip batch)
  int lmin = INT_MAX, lmax = INT_MIN;
  const int warp = tid >> 5, lane = tid & 31;
  for (int mI = tid; mI < n_mels; mI += 256) {
    const int f0 = band_start[mI], n = band_len[mI];
    float acc = 0.f;
#pragma unroll 4
    for (int i = 0; i < n; ++i)   // weights beyond the compact rows (a filter wider than MEL_BAND_MAX bins) come from the full matrix
      acc = fmaf(pw[f0 + i], i < MEL_BAND_MAX ? __ldg(fbT + (size_t)i * MEL_PITCH + mI) : __ldg(fb + (size_t)(f0 + i) * n_mels + mI), acc);
    const float v = logf(acc + 1e-6f);  // torch.log(spec + 1e-6)
    logmel[((size_t)b * nF + t) * n_mels + mI] = v;
    if (v == v) { const int o = f2ord(v); lmin = min(lmin, o); lmax = max(lmax, o); }
  }
  for (int o = 16; o > 0; o >>= 1) {
    lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
    lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  }
  if (lane == 0) { smin[warp] = lmin; smax[warp] = lmax; }
  __syncthreads();
  if (tid == 0) {
    lmin = smin[0]; lmax = smax[0];
    for (int i = 1; i < 8; ++i) { lmin = min(lmin, smin[i]); lmax = max(lmax, smax[i]); }
    atomicMin(&minmax[2 * b], lmin);
    atomicMax(&minmax[2 * b + 1], lmax);
  }
}

__global__ void __launch_bounds__(256) spec_finalize_kernel(const float* __restrict__ logmel, const int* __restrict__ minmax,
                                                            float* __restrict__ spec_out, float* __restrict__ spec32,
                                                            void* __restrict__ spec_main, void* __restrict__ spec_aux, int fmt,
                                                            int n_mels, int nF, int T, int Mp, int it0, int it1, int if0, int if1) {
  extern __shared__ float tile[];  // [32][n_mels+1]
  const int t0 = blockIdx.x * 32, b = blockIdx.y;
  const float vmin = ord2f(minmax[2 * b]), vmax = ord2f(minmax[2 * b + 1]);
  const float range = vmax - vmin;
  const bool has_t = it1 > it0, has_f = if1 > if0;
  const int ld = n_mels + 1;
  for (int idx = threadIdx.x; idx < 32 * Mp; idx += blockDim.x) {
    const int r = idx / Mp, m = idx - r * Mp, t = t0 + r;
    if (t >= T) continue;
    float v = 0.f;
    if (m < n_mels) {
      v = (logmel[((size_t)b * nF + t) * n_mels + m] - vmin) / range;  // (x - min)/(max - min) * (1-0) + 0
      if (v != v) v = 0.f;                                             // x_scaled[isnan] = min
      const bool in_t = t >= it0 && t < it1, in_f = m >= if0 && m < if1;
      const bool masked = (has_t && has_f) ? (in_t && in_f) : (has_t ? in_t : (has_f ? in_f : false));
      if (masked) v = -1.f;                                            // model/diffwave.py:649-654
      tile[r * ld + m] = v;
    }
    const size_t o = ((size_t)b * T + t) * Mp + m;
    if (spec32) spec32[o] = v;
    if (fmt == 1) {
      __nv_bfloat16 h, l; split_bf16(v, h, l);
      reinterpret_cast<__nv_bfloat16*>(spec_main)[o] = h; reinterpret_cast<__nv_bfloat16*>(spec_aux)[o] = l;
    } else if (fmt >= 2) {   // fp16 + fp8 bytes [lo (64) | hi (64)] per 64-mel chunk (scales: see common.cuh)
      const __half h = __float2half_rn(v);
      const float hf = __half2float(h);
      reinterpret_cast<__half*>(spec_main)[o] = h;
      uint8_t* a8 = reinterpret_cast<uint8_t*>(spec_aux) + ((size_t)b * T + t) * 2 * Mp + (size_t)(m >> 6) * 128 + (m & 63);
      if (fmt == 2) {
        a8[0] = (uint8_t)__nv_cvt_float_to_fp8((v - hf) * F8_SA, __NV_SATFINITE, __NV_E4M3);
        a8[64] = (uint8_t)__nv_cvt_float_to_fp8(hf, __NV_SATFINITE, __NV_E4M3);
      } else {
        a8[0] = (uint8_t)__nv_cvt_float_to_fp8((v - hf) * 16.f, __NV_SATFINITE, __NV_E5M2);
        a8[64] = (uint8_t)__nv_cvt_float_to_fp8(hf * 0.00390625f, __NV_SATFINITE, __NV_E5M2);
      }
    }
  }
  __syncthreads();
  if (spec_out) {
    for (int idx = threadIdx.x; idx < 32 * n_mels; idx += blockDim.x) {
      const int m = idx >> 5, r = idx & 31, t = t0 + r;
      if (t < T) spec_out[((size_t)b * n_mels + m) * T + t] = tile[r * ld + m];
    }
  }
}

int mel_create(MelPlan** out, const drb_config& c, const float* window, const float* fb, void* ws, size_t ws_bytes,
               cudaStream_t s) {
  MelPlan* p = new MelPlan();
  p->B = c.batch; p->L = c.wave_len; p->n_fft = c.n_fft; p->hop = c.hop_length; p->n_mels = c.n_mels;
  p->nF = c.wave_len / c.hop_length + 1; p->nbins = c.n_fft / 2 + 1;
  p->window = window; p->fb = fb; p->fft_ok = false;
  const size_t rows = (size_t)p->B * p->nF;
  char* base = (char*)ws; size_t off = 0;
  auto take = [&](size_t bytes) { void* r = base + off; off += align_up(bytes); return r; };
  p->fused = mel_fused(p->n_fft, p->n_mels);
  p->frames = nullptr; p->spectrum = nullptr; p->twiddle = nullptr;
  if (!p->fused) {
    p->frames = (float*)take(rows * p->n_fft * sizeof(float));
    p->spectrum = (float2*)take(rows * p->nbins * sizeof(float2));
  }
  p->logmel = (float*)take(rows * p->n_mels * sizeof(float));
  p->band_start = (int*)take(p->n_mels * sizeof(int));
  p->band_len = (int*)take(p->n_mels * sizeof(int));
  p->wnorm = (float*)take(sizeof(float));
  p->minmax = (int*)take((size_t)p->B * 2 * sizeof(int));
  if (p->fused) {
    p->twiddle = (float2*)take(2048 * sizeof(float2));
    if (off > ws_bytes) { set_error("mel workspace too small"); delete p; return DRB_E_WORKSPACE; }
    p->fbT = (float*)take((size_t)MEL_BAND_MAX * MEL_PITCH * sizeof(float));
    if (off > ws_bytes) { set_error("mel workspace too small"); delete p; return DRB_E_WORKSPACE; }
    twiddle_kernel<<<8, 256, 0, s>>>(p->twiddle);
    DRB_LAUNCH_CHECK();
    band_kernel<<<(p->n_mels + 127) / 128, 128, 0, s>>>(fb, p->nbins, p->n_mels, p->band_start, p->band_len);
    DRB_LAUNCH_CHECK();
    band_pack_kernel<<<p->n_mels, MEL_BAND_MAX, 0, s>>>(fb, p->band_start, p->band_len, p->n_mels, p->fbT);
    DRB_LAUNCH_CHECK();
    wnorm_kernel<<<1, 256, 0, s>>>(window, p->n_fft, p->wnorm);
    DRB_LAUNCH_CHECK();
    *out = p;
    return 0;
  }
  p->fft_work = base + off;
  p->fft_work_bytes = ws_bytes > off ? ws_bytes - off : 0;

  if (cufftCreate(&p->fft) != CUFFT_SUCCESS) { set_error("cufftCreate failed"); delete p; return DRB_E_CUFFT; }
  p->fft_ok = true;
  cufftSetAutoAllocation(p->fft, 0);
  int n[1] = {p->n_fft};
  size_t need = 0;
  cufftResult r = cufftMakePlanMany(p->fft, 1, n, nullptr, 1, p->n_fft, nullptr, 1, p->nbins, CUFFT_R2C, (int)rows, &need);
  if (r != CUFFT_SUCCESS) { set_error("cufftMakePlanMany failed (%d)", (int)r); mel_destroy(p); return DRB_E_CUFFT; }
  if (need > p->fft_work_bytes) {
    // The estimate used to size the workspace was too small for this cuFFT version: let cuFFT own its scratch.
    cufftDestroy(p->fft);
    if (cufftCreate(&p->fft) != CUFFT_SUCCESS) { p->fft_ok = false; set_error("cufftCreate failed"); mel_destroy(p); return DRB_E_CUFFT; }
    r = cufftMakePlanMany(p->fft, 1, n, nullptr, 1, p->n_fft, nullptr, 1, p->nbins, CUFFT_R2C, (int)rows, &need);
    if (r != CUFFT_SUCCESS) { set_error("cufftMakePlanMany failed (%d)", (int)r); mel_destroy(p); return DRB_E_CUFFT; }
  } else if (cufftSetWorkArea(p->fft, p->fft_work) != CUFFT_SUCCESS) {
    set_error("cufftSetWorkArea failed"); mel_destroy(p); return DRB_E_CUFFT;
  }

  band_kernel<<<(p->n_mels + 127) / 128, 128, 0, s>>>(fb, p->nbins, p->n_mels, p->band_start, p->band_len);
  DRB_LAUNCH_CHECK();
  wnorm_kernel<<<1, 256, 0, s>>>(window, p->n_fft, p->wnorm);
  DRB_LAUNCH_CHECK();
  *out = p;
  return 0;
}

void mel_destroy(MelPlan* p) {
  if (!p) return;
  if (p->fft_ok) cufftDestroy(p->fft);
  delete p;
}

float* mel_logmel_ptr(MelPlan* p, size_t* bytes) {
  if (bytes) *bytes = (size_t)p->B * p->nF * p->n_mels * sizeof(float);
  return p->logmel;
}

int mel_forward(MelPlan* p, const float* waveform, float* spec_out, float* spec32, void* spec_main, void* spec_aux, int fmt,
                int Mp, int T, int it0, int it1, int if0, int if1, cudaStream_t s) {
  if (T > p->nF) { set_error("mel_forward: T=%d > frames %d", T, p->nF); return DRB_E_INVALID; }
  dim3 grid(p->nF, p->B);
  if (p->fused) {
    minmax_init_kernel<<<(p->B + 127) / 128, 128, 0, s>>>(p->minmax, p->B);
    DRB_LAUNCH_CHECK();
    const int vec_ok = ((p->L | p->hop) & 1) == 0 && ((uintptr_t)waveform & 7) == 0 && ((uintptr_t)p->window & 7) == 0;
    if (((uintptr_t)p->window & 7) != 0) { set_error("mel_forward: window must be 8-byte aligned"); return DRB_E_INVALID; }
    stft_mel_kernel<<<grid, 256, 0, s>>>(waveform, p->window, p->twiddle, p->fbT, p->fb, p->band_start, p->band_len, p->wnorm, p->logmel,
                                         p->minmax, p->L, p->hop, p->nF, p->n_mels, vec_ok);
    DRB_LAUNCH_CHECK();
    dim3 g2f((T + 31) / 32, p->B);
    const size_t smf = (size_t)32 * (p->n_mels + 1) * sizeof(float);
    spec_finalize_kernel<<<g2f, 256, smf, s>>>(p->logmel, p->minmax, spec_out, spec32, spec_main, spec_aux, fmt, p->n_mels, p->nF,
                                              T, Mp, it0, it1, if0, if1);
    DRB_LAUNCH_CHECK();
    return 0;
  }
  frame_kernel<<<grid, 256, 0, s>>>(waveform, p->window, p->frames, p->L, p->n_fft, p->hop, p->nF);
  DRB_LAUNCH_CHECK();
  if (cufftSetStream(p->fft, s) != CUFFT_SUCCESS) { set_error("cufftSetStream failed"); return DRB_E_CUFFT; }
  cufftResult r = cufftExecR2C(p->fft, p->frames, (cufftComplex*)p->spectrum);
  if (r != CUFFT_SUCCESS) { set_error("cufftExecR2C failed (%d)", (int)r); return DRB_E_CUFFT; }
  minmax_init_kernel<<<(p->B + 127) / 128, 128, 0, s>>>(p->minmax, p->B);
  DRB_LAUNCH_CHECK();
  power_mel_kernel<<<grid, 256, p->nbins * sizeof(float), s>>>(p->spectrum, p->fb, p->band_start, p->band_len, p->wnorm,
                                                              p->logmel, p->minmax, p->nbins, p->n_mels, p->nF);
  DRB_LAUNCH_CHECK();
  dim3 g2((T + 31) / 32, p->B);
  size_t sm = (size_t)32 * (p->n_mels + 1) * sizeof(float);
  spec_finalize_kernel<<<g2, 256, sm, s>>>(p->logmel, p->minmax, spec_out, spec32, spec_main, spec_aux, fmt, p->n_mels, p->nF,
                                           T, Mp, it0, it1, if0, if1);
  DRB_LAUNCH_CHECK();
  return 0;
}

}  // namespace drb

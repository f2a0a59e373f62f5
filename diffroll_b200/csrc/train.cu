// Row f3 of SURVEY.md section 8: the TRAINING step of ClassifierFreeDiffRoll, forward with saved activations and the
// full backward pass to every one of the 132 state_dict tensors, plus Adam.
//   reference: SpecRollDiffusion.step / training_step   task/diffusion.py:258-270, 651-763
//              ClassifierFreeDiffRoll.forward            model/diffwave.py:637-686   (training mode: spec dropout :646-647, 689-693)
//              ResidualBlock.forward                     model/diffwave.py:134-151
//              DiffusionEmbedding.forward                model/diffwave.py:66-75
//              configure_optimizers (torch.optim.Adam)   task/diffusion.py:1057-1067
//
// Arithmetic.  DRB_TRAIN_TC=0: fp32 on the CUDA cores throughout, like the reference trains -- the contractions reuse the
// generic fp32 GEMM of the validation path (simt_kernels.cu: NT form with the dilated-tap gather) for the forward and for
// every "gradient with respect to the input" product (dgrad: the same gather with mirrored taps over a transposed weight
// copy), and simt_wgrad_kernel for every "gradient with respect to a weight" product (contraction over rolls x frames, TN
// form, the same tap gather on the activation side, fp32 atomics).  Default: the five products that carry 97 % of the
// FLOPs run on the tensor cores (umma_gemm.cu): the forward conv and output_projection in f16x3 (fp16 hi + lo, fp32-grade)
// with one launch per tap and exact fp32 accumulation across launches, conv dgrad and the 1x1 dgrad through the same
// kernel over scaled f16e5 pairs, conv wgrad as one plain GEMM over transposed im2col pairs.  DESIGN.md section 8.
//
// Memory (one drb_train plan, caller-provided workspace): the inputs x_l, pre-activations y_l and gated activations z_l
// of every layer are kept for the backward pass (L x M x 4C floats, 2.5 GB at 32 rolls x 640 frames).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "kernels.h"
#include "../../include/diffroll_b200.h"

namespace drb {

// ---------------------------------------------------------------------------------------------
// wgrad:  dW[n][c][tap] (+)= sum_m G[m][n] * Aeff(m, tap, c)
//   Aeff(m, tap, c) = X[(seg*T + t')*ldx + c] + addvec[seg][c]   with m = seg*T + t, t' = t + (tap - taps/2)*dil, 0 <= t' < T
//   (else 0): exactly the operand the forward GEMM contracted with W[n][tap*Ck + c].
// 128 x 128 output tile per block, 16 rows of m per stage, 8 x 8 outputs per thread; blockIdx.z splits the m range and
// the partial tiles are accumulated with atomicAdd (dW is zeroed, or holds the gradient being accumulated, beforehand).
// Output address: dW + n*sn + c*sc + tap*st  (PyTorch conv weights are [out][in][k]: sn = Ck*k, sc = k, st = 1).
// ---------------------------------------------------------------------------------------------
struct WgradDev {
  const float* G; int ldg; const float* X; int ldx; const float* addvec; int av_stride;
  int M, T, N, Ck, taps, dil, rows_per_split;
  float* dW; long long sn, sc, st; float out_scale;
};

__global__ void __launch_bounds__(256, 2) simt_wgrad_kernel(const WgradDev g) {
  constexpr int WK = 16;
  __shared__ __align__(16) float Gs[2][WK][128 + 4];
  __shared__ __align__(16) float Xs[2][WK][128 + 4];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * 128;
  const int cblocks = (g.Ck + 127) / 128;
  const int tap = blockIdx.y / cblocks, c0 = (blockIdx.y - tap * cblocks) * 128;
  const int m_lo = blockIdx.z * g.rows_per_split, m_hi = min(g.M, m_lo + g.rows_per_split);
  const int shift = (tap - g.taps / 2) * g.dil;
  // loader: thread -> (row r = tid / 16 of the stage, 8 consecutive columns starting at (tid % 16) * 8)
  const int lr = tid >> 4, lc = (tid & 15) * 8;
  float4 rg[2], rx[2];
  auto load = [&](int m0) {
    const int m = m0 + lr;
    const bool m_ok = m < m_hi;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int n = n0 + lc + q * 4;
      if (m_ok && n < g.N) {
        const float* p = g.G + (size_t)m * g.ldg + n;
        if (n + 3 < g.N) v = *reinterpret_cast<const float4*>(p);
        else { v.x = p[0]; if (n + 1 < g.N) v.y = p[1]; if (n + 2 < g.N) v.z = p[2]; }
      }
      rg[q] = v;
    }
    int seg = 0, t2 = -1;
    if (m_ok) { seg = m / g.T; t2 = m - seg * g.T + shift; }
    const bool x_ok = m_ok && t2 >= 0 && t2 < g.T;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int c = c0 + lc + q * 4;
      if (x_ok && c < g.Ck) {
        const float* p = g.X + (size_t)(seg * g.T + t2) * g.ldx + c;
        if (c + 3 < g.Ck) v = *reinterpret_cast<const float4*>(p);
        else { v.x = p[0]; if (c + 1 < g.Ck) v.y = p[1]; if (c + 2 < g.Ck) v.z = p[2]; }
        if (g.addvec) {
          const float* a = g.addvec + (size_t)seg * g.av_stride + c;
          v.x += a[0]; if (c + 1 < g.Ck) v.y += a[1]; if (c + 2 < g.Ck) v.z += a[2]; if (c + 3 < g.Ck) v.w += a[3];
        }
      }
      rx[q] = v;
    }
  };
  auto store = [&](int buf) {
    *reinterpret_cast<float4*>(&Gs[buf][lr][lc]) = rg[0]; *reinterpret_cast<float4*>(&Gs[buf][lr][lc + 4]) = rg[1];
    *reinterpret_cast<float4*>(&Xs[buf][lr][lc]) = rx[0]; *reinterpret_cast<float4*>(&Xs[buf][lr][lc + 4]) = rx[1];
  };
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int nk = (m_hi - m_lo + WK - 1) / WK;
  if (nk <= 0) return;
  load(m_lo); store(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) load(m_lo + (kb + 1) * WK);
#pragma unroll
    for (int k = 0; k < WK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&Gs[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&Gs[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Xs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Xs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kb + 1 < nk) { store(buf ^ 1); __syncthreads(); }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (n >= g.N) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (c >= g.Ck) continue;
      atomicAdd(g.dW + (long long)n * g.sn + (long long)c * g.sc + (long long)tap * g.st, acc[i][j] * g.out_scale);
    }
  }
}

struct Wgrad {
  const float* G = nullptr; int ldg = 0; const float* X = nullptr; int ldx = 0; const float* addvec = nullptr; int av_stride = 0;
  int M = 0, T = 1, N = 0, Ck = 0, taps = 1, dil = 1;
  float* dW = nullptr; long long sn = 0, sc = 1, st = 0;
  float out_scale = 1.f;   // dW += out_scale * (G^T Aeff): the forward's division of the GEMM input (skip / sqrt(L))
};
static int launch_wgrad(const Wgrad& w, cudaStream_t s) {
  WgradDev g;
  g.G = w.G; g.ldg = w.ldg; g.X = w.X; g.ldx = w.ldx; g.addvec = w.addvec; g.av_stride = w.av_stride;
  g.M = w.M; g.T = w.T; g.N = w.N; g.Ck = w.Ck; g.taps = w.taps; g.dil = w.dil; g.dW = w.dW; g.sn = w.sn; g.sc = w.sc; g.st = w.st; g.out_scale = w.out_scale;
  if (w.M <= 0 || w.N <= 0 || w.Ck <= 0 || (w.ldg & 3) || (w.ldx & 3) || ((uintptr_t)w.G & 15) || ((uintptr_t)w.X & 15)) {
    set_error("wgrad: unsupported shape M=%d N=%d Ck=%d ldg=%d ldx=%d", w.M, w.N, w.Ck, w.ldg, w.ldx);
    return DRB_E_INVALID;
  }
  const int nb = (w.N + 127) / 128, cb = (w.Ck + 127) / 128 * w.taps;
  // enough blocks to fill the machine a few times over; every split covers a multiple of 16 rows
  int splits = (4 * 148 + nb * cb - 1) / (nb * cb);
  const int max_splits = (w.M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  g.rows_per_split = ((w.M + splits - 1) / splits + 15) / 16 * 16;
  splits = (w.M + g.rows_per_split - 1) / g.rows_per_split;
  dim3 grid(nb, cb, splits);
  simt_wgrad_kernel<<<grid, 256, 0, s>>>(g);
  DRB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// elementwise / reduction kernels of the training step
// ---------------------------------------------------------------------------------------------
// out[seg][n] (+)= sum over the rows of segment seg of G[row][n]   (rows_per_seg = M: one segment = a plain column sum)
__global__ void colsum_kernel(const float* __restrict__ G, int ldg, int N, int rows_per_seg, int chunk, float* __restrict__ out, int ldo) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int seg = blockIdx.z;
  const int r0 = blockIdx.y * chunk, r1 = min(rows_per_seg, r0 + chunk);
  if (n >= N) return;
  const float* p = G + ((size_t)seg * rows_per_seg + r0) * ldg + n;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int r = r0;
  for (; r + 3 < r1; r += 4, p += 4 * (size_t)ldg) { s0 += p[0]; s1 += p[ldg]; s2 += p[2 * (size_t)ldg]; s3 += p[3 * (size_t)ldg]; }
  for (; r < r1; ++r, p += ldg) s0 += p[0];
  atomicAdd(out + (size_t)seg * ldo + n, (s0 + s1) + (s2 + s3));
}
static int launch_colsum(const float* G, int ldg, int N, int segs, int rows_per_seg, float* out, int ldo, cudaStream_t s) {
  const int chunk = 64;
  dim3 grid((N + 127) / 128, (rows_per_seg + chunk - 1) / chunk, segs);
  colsum_kernel<<<grid, 128, 0, s>>>(G, ldg, N, rows_per_seg, chunk, out, ldo);
  DRB_LAUNCH_CHECK();
  return 0;
}

// x_out = (x_in + o[:, :C]) / sqrt(2)  (skipped for the last layer) ; skip (+)= o[:, C:]      model/diffwave.py:150-151, 680
__global__ void res_skip_fwd_kernel(const float* __restrict__ o, const float* __restrict__ x_in, float* __restrict__ x_out,
                                    float* __restrict__ skip, size_t M, int C, int first, int do_res) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C / 4) return;
  const size_t m = (i * 4) / C; const int c = (int)((i * 4) % C);
  const float rs2 = 1.41421356237309515f;
  if (do_res) {
    const float4 r = *reinterpret_cast<const float4*>(o + m * 2 * C + c);
    float4 xv = *reinterpret_cast<const float4*>(x_in + m * C + c);
    xv.x = (xv.x + r.x) / rs2; xv.y = (xv.y + r.y) / rs2; xv.z = (xv.z + r.z) / rs2; xv.w = (xv.w + r.w) / rs2;
    *reinterpret_cast<float4*>(x_out + m * C + c) = xv;
  }
  float4 sk = *reinterpret_cast<const float4*>(o + m * 2 * C + C + c);
  if (!first) {
    const float4 p = *reinterpret_cast<const float4*>(skip + m * C + c);
    sk.x += p.x; sk.y += p.y; sk.z += p.z; sk.w += p.w;
  }
  *reinterpret_cast<float4*>(skip + m * C + c) = sk;
}

// g_o[m] = [ g_x[m] / sqrt(2)  (zeros for the last layer) | g_skip[m] ]      gradient of the layer's output_projection output
__global__ void make_go_kernel(const float* __restrict__ gx, const float* __restrict__ gskip, float* __restrict__ go, size_t M, int C,
                               int has_res) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C / 4) return;
  const size_t m = (i * 4) / C; const int c = (int)((i * 4) % C);
  const float rs2 = 1.41421356237309515f;
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (has_res) { r = *reinterpret_cast<const float4*>(gx + m * C + c); r.x /= rs2; r.y /= rs2; r.z /= rs2; r.w /= rs2; }
  *reinterpret_cast<float4*>(go + m * 2 * C + c) = r;
  *reinterpret_cast<float4*>(go + m * 2 * C + C + c) = *reinterpret_cast<const float4*>(gskip + m * C + c);
}

// z = sigmoid(y_g) * tanh(y_f):  g_yg = g_z * tanh(y_f) * s (1 - s),  g_yf = g_z * s * (1 - tanh(y_f)^2)      model/diffwave.py:146-147
__global__ void gate_bwd_kernel(const float* __restrict__ y, const float* __restrict__ gz, float* __restrict__ gy, size_t M, int C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C / 4) return;
  const size_t m = (i * 4) / C; const int c = (int)((i * 4) % C);
  const float4 g4 = *reinterpret_cast<const float4*>(y + m * 2 * C + c);
  const float4 f4 = *reinterpret_cast<const float4*>(y + m * 2 * C + C + c);
  const float4 z4 = *reinterpret_cast<const float4*>(gz + m * C + c);
  const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, ff[4] = {f4.x, f4.y, f4.z, f4.w}, zz[4] = {z4.x, z4.y, z4.z, z4.w};
  float og[4], of[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float sg = 1.f / (1.f + expf(-gg[e])), th = tanhf(ff[e]);
    og[e] = zz[e] * th * sg * (1.f - sg);
    of[e] = zz[e] * sg * (1.f - th * th);
  }
  *reinterpret_cast<float4*>(gy + m * 2 * C + c) = make_float4(og[0], og[1], og[2], og[3]);
  *reinterpret_cast<float4*>(gy + m * 2 * C + C + c) = make_float4(of[0], of[1], of[2], of[3]);
}

// g_x_l = g_u (+ g_x_{l+1} / sqrt(2)) in place on gu ; the dead gradient of the last layer's residual output is skipped
__global__ void add_res_grad_kernel(float* __restrict__ gu, const float* __restrict__ gx_next, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float rs2 = 1.41421356237309515f;
  float4 a = reinterpret_cast<float4*>(gu)[i];
  const float4 b = reinterpret_cast<const float4*>(gx_next)[i];
  a.x += b.x / rs2; a.y += b.y / rs2; a.z += b.z / rs2; a.w += b.w / rs2;
  reinterpret_cast<float4*>(gu)[i] = a;
}

// g *= (act > 0)  (ReLU backward, from the saved post-activation), optionally scaled
__global__ void relu_bwd_kernel(float* __restrict__ g, const float* __restrict__ act, size_t n4, float scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = reinterpret_cast<float4*>(g)[i];
  const float4 h = reinterpret_cast<const float4*>(act)[i];
  a.x = h.x > 0.f ? a.x * scale : 0.f; a.y = h.y > 0.f ? a.y * scale : 0.f;
  a.z = h.z > 0.f ? a.z * scale : 0.f; a.w = h.w > 0.f ? a.w * scale : 0.f;
  reinterpret_cast<float4*>(g)[i] = a;
}
__global__ void scale_kernel(float* __restrict__ g, size_t n, float scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g[i] *= scale;
}
// silu(p) = p * sigmoid(p);  g_p = g * (s + p s (1 - s))      model/diffwave.py:53-55
__global__ void silu_bwd_kernel(float* __restrict__ g, const float* __restrict__ p, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = p[i], s = 1.f / (1.f + expf(-v));
  g[i] *= s + v * s * (1.f - s);
}
// out[b][:] = table[steps[b]][:]   (DiffusionEmbedding lookup, model/diffwave.py:67-68)
__global__ void gather_rows_kernel(const float* __restrict__ table, const int* __restrict__ steps, float* __restrict__ out, int B, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * W) return;
  const int b = i / W, j = i - b * W;
  out[i] = table[(size_t)steps[b] * W + j];
}
__global__ void iota_kernel(int* p, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; }

// dst[c][r] = src[r][c]  (weights: [rows][cols] -> [cols][rows]); 32 x 32 tiles through shared memory
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dst[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}
static int launch_transpose(const float* src, float* dst, int rows, int cols, cudaStream_t s) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, s>>>(src, dst, rows, cols);
  DRB_LAUNCH_CHECK();
  return 0;
}
// Dilated-conv weight for the dgrad GEMM: out[c][tap'*OC + n] = w[n][c][k-1-tap']   (w: PyTorch [OC][C][k])
// The dgrad is the same tap gather as the forward over g_y with the taps mirrored: g_u[t] = sum_tap g_y[t - (tap-k/2) dil] . w[:, :, tap]
__global__ void repack_dgrad_kernel(const float* __restrict__ w, float* __restrict__ out, int OC, int C, int k) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)OC * C * k) return;
  const int n = (int)(i % OC); const size_t r = i / OC; const int tp = (int)(r % k); const int c = (int)(r / k);
  out[i] = w[((size_t)n * C + c) * k + (k - 1 - tp)];
}
// spec [B][n_mels][T] (the module's layout) -> [B*T][Mp] time-major, zero padded to Mp columns
__global__ void spec_to_rows_kernel(const float* __restrict__ spec, float* __restrict__ out, int B, int nm, int T, int Mp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int m = m0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (m < nm && t < T) ? spec[((size_t)b * nm + m) * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int t = t0 + i, m = m0 + threadIdx.x;
    if (t < T && m < Mp) out[((size_t)b * T + t) * Mp + m] = tile[threadIdx.x][i];
  }
}

// rows [B*T][Mp] time-major -> spec layout [B][n_mels][T] (the inverse of spec_to_rows_kernel), optionally added to the destination
__global__ void rows_to_spec_kernel(const float* __restrict__ rows, float* __restrict__ spec, int B, int nm, int T, int Mp, int accumulate) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int t = t0 + i, m = m0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < T && m < Mp) ? rows[((size_t)b * T + t) * Mp + m] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int m = m0 + i, t = t0 + threadIdx.x;
    if (m < nm && t < T) {
      const size_t j = ((size_t)b * nm + m) * T + t;
      spec[j] = tile[threadIdx.x][i] + (accumulate ? spec[j] : 0.f);
    }
  }
}

// grad[n][c][tap] += dwt[n][tap * C + c]   (tap-major GEMM result -> PyTorch conv weight layout [out][in][k])
__global__ void untap_add_kernel(const float* __restrict__ dwt, float* __restrict__ grad, int OC, int C, int k) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)OC * C * k) return;
  const int tap = (int)(i % k); const size_t r = i / k; const int c = (int)(r % C); const size_t n = r / C;
  grad[i] += dwt[n * ((size_t)k * C) + (size_t)tap * C + c];
}
__global__ void add2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}
// inverse scale of a product of two scaled operands: out[0] = a[1] * b[1]   (slots as written by launch_weight_scale: {S, 1/S})
__global__ void mul_inv_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out) { out[0] = a[1] * b[1]; }

// Adam, torch.optim.Adam semantics (amsgrad=False, maximize=False; weight_decay adds wd * p to the gradient):
//   m = b1 m + (1 - b1) g ; v = b2 v + (1 - b2) g^2 ; p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                            float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i];
  const float pi = p[i];
  if (wd != 0.f) gi = fmaf(wd, pi, gi);
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = pi - (lr / bc1) * (mi / denom);
}

// The same update for a whole parameter list in ONE launch: table[t] = {param, grad, exp_avg, exp_avg_sq, numel} (five 64-bit words, device
// memory), block_map[b] = {tensor, first element / 4096}: every block owns 4096 consecutive elements of one tensor.
constexpr int ADAM_CHUNK = 4096;
__global__ void __launch_bounds__(256) adam_multi_kernel(const unsigned long long* __restrict__ table, const int2* __restrict__ block_map,
                                                         float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
  const int2 bm = block_map[blockIdx.x];
  const unsigned long long* e = table + (size_t)bm.x * 5;
  float* __restrict__ p = reinterpret_cast<float*>(e[0]);
  const float* __restrict__ g = reinterpret_cast<const float*>(e[1]);
  float* __restrict__ m = reinterpret_cast<float*>(e[2]);
  float* __restrict__ v = reinterpret_cast<float*>(e[3]);
  const size_t n = (size_t)e[4], lo = (size_t)bm.y * ADAM_CHUNK, hi = lo + ADAM_CHUNK < n ? lo + ADAM_CHUNK : n;
  for (size_t i = lo + threadIdx.x; i < hi; i += 256) {
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// d loss / d prediction for the mean l1 / l2 / smooth-l1 losses of p_losses (task/diffusion.py:792-802), times a per-roll factor
// (training mode 'ex_0': pred_roll = (x_t - s1[t] eps) / sa[t]  ->  d pred_roll / d eps = -s1[t] / sa[t])
__global__ void loss_grad_kernel(const float* __restrict__ label, const float* __restrict__ pred, float* __restrict__ g, size_t n,
                                 size_t per_roll, int loss_type, const float* __restrict__ roll_scale, float inv_n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = pred[i] - label[i];
  float r;
  if (loss_type == 0) r = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);      // l1
  else if (loss_type == 1) r = 2.f * d;                                 // l2
  else r = fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f);                 // huber = smooth_l1_loss, beta = 1
  r *= inv_n;
  if (roll_scale) r *= roll_scale[i / per_roll];
  g[i] = r;
}

}  // namespace drb

using namespace drb;

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
struct drb_train {
  drb_train_config cfg;
  char* ws;
  int Mp;
  std::vector<int> dil;
  // workspace offsets (bytes)
  size_t xs, ys, zs, skip, hbuf, obuf, spec, e0, p1, s1, p2, emb, dl, iota, wtmp, gskip, gx, gu, gz, gemb, gd, gsmall, total;
  bool fwd_done = false;
  // drb_train_set_spec_grad: where the next backward leaves d loss / d spec ([B][n_mels][T]); nullptr = not wanted (the default:
  // the spectrogram is data).  condition='trainable_spec' needs it for the rolls conditioned on the learned table.
  float* g_spec_out = nullptr;
  size_t gspec = 0;          // [B*T][Mp] fp32 accumulator over the layers
  // tensor-core path of the two dilated-conv products that carry 2/3 of the step's FLOPs (forward conv, its transposed form):
  // f16e5 operand pairs staged per layer, umma_gate_kernel with the linear epilogue (DRB_TRAIN_TC=0: everything on the CUDA cores)
  bool tc = false;
  // DRB_TRAIN_TC bits: 1 forward conv, 2 dgrad conv, 4 wgrad conv, 8 the 1x1 output_projection (forward and dgrad), 16 the 1x1 weight gradients on the
  // tensor cores (default all).  Products that are linear in a gradient use f16e5 pairs (2 MMA units; their 2^-15 operand rounding stays a ~5e-5
  // relative error of that gradient).  FORWARD products use f16x3 (fp16 hi + fp16 lo, 3 units, 22 mantissa bits): an f16e5 forward
  // moved the activations by 1e-4 and with them ReLU / gate / L1-loss derivatives (measured worst gradient error 9e-4 .. 1.5e-2
  // instead of 6e-5), so the forward gets fp32-grade operands.
  int tc_mask = 31;
  int dgrad_taps = 1;        // DRB_TRAIN_DGRAD_TAPS: taps per pass of the transposed conv (9: one pass, the round-2 form)
  size_t wtmp_bytes = 0;
  int taps_per_launch = 1;   // DRB_TRAIN_TAPS: taps of the forward conv per tensor-core launch (accumulation chain = taps * C terms)
  size_t uh, ul, gh, gl, sph, spl, wh, wl, wch, wcl, bnat, scal, gth, gtl, uth, utl;
  CUtensorMap m_uh, m_ul, m_gh, m_gl, m_sh, m_sl, m_wfh, m_wfl, m_wgh, m_wgl, m_wch, m_wcl;
  CUtensorMap m_ul5, m_sl5, m_wfl5, m_wcl5, m_woh, m_wol5, m_woTh, m_woTl;   // f16x3 aux views (fp16 lo) and the output_projection weights
  CUtensorMap m_gth, m_gtl, m_uth, m_utl, m_dwt;   // transposed pairs g_y^T [2C][M], im2col(x + d)^T [k*C][M]; tap-major weight gradient [2C][k*C]
  bool tc_wgrad = false;
  // 1x1 weight gradients (output_projection, conditioner_projection, skip_projection) as split-K tcgen05 GEMMs: few output tiles
  // (16 / 8 / 8) against K = rolls x frames, so the K range is cut into `splits` launches-worth of tiles that fill the SMs; every
  // split leaves an fp32 partial [splits][rows][cols] in wtmp and splitk_reduce_kernel adds them into the gradient
  struct SplitK { int splits = 0, slabs = 0; CUtensorMap out; };
  SplitK sk_wo, sk_wc, sk_sk;
  template <class Tp> Tp* at(size_t off) const { return reinterpret_cast<Tp*>(ws + off); }
};

static size_t train_layout(drb_train& p) {
  const drb_train_config& c = p.cfg;
  const size_t B = c.batch, T = c.frames, C = c.residual_channels, L = c.residual_layers, k = c.kernel_size, M = B * T;
  p.Mp = (c.n_mels + 63) / 64 * 64;   // one K padding for the CUDA-core and the tensor-core conditioner products
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t r = off; off += (bytes + 255) / 256 * 256; return r; };
  p.xs = take(L * M * C * 4); p.ys = take(L * M * 2 * C * 4); p.zs = take(L * M * C * 4);
  p.skip = take(M * C * 4); p.hbuf = take(M * C * 4); p.obuf = take(M * 2 * C * 4);
  p.spec = take(M * (size_t)p.Mp * 4);
  p.e0 = take(B * 128 * 4); p.p1 = take(B * 512 * 4); p.s1 = take(B * 512 * 4); p.p2 = take(B * 512 * 4); p.emb = take(B * 512 * 4);
  p.dl = take(L * B * C * 4); p.iota = take(B * 4);
  size_t wmax = 2 * C * k * C;                       // repacked / transposed weight scratch (largest: the dilated conv)
  if (wmax < 4 * C * (size_t)p.Mp) wmax = 4 * C * (size_t)p.Mp;   // padded conditioner weights + their transpose (spec gradient)
  p.wtmp = take(wmax * 4); p.wtmp_bytes = wmax * 4;
  p.gskip = take(M * C * 4); p.gx = take(M * C * 4); p.gu = take(M * C * 4); p.gz = take(M * C * 4);
  p.gemb = take(B * 512 * 4); p.gd = take(B * C * 4); p.gsmall = take(B * 512 * 4);
  // operand pairs (2 + 2 bytes per element): x + d [M][C], g_y [M][2C], spectrogram [M][Mp], one layer's conv weights, conditioner weights
  p.uh = take(M * C * 2); p.ul = take(M * C * 2); p.gh = take(M * 2 * C * 2); p.gl = take(M * 2 * C * 2);
  p.sph = take(M * (size_t)p.Mp * 2); p.spl = take(M * (size_t)p.Mp * 2);
  p.wh = take(2 * C * k * C * 2); p.wl = take(2 * C * k * C * 2); p.wch = take(2 * C * (size_t)p.Mp * 2); p.wcl = take(2 * C * (size_t)p.Mp * 2);
  p.bnat = take(2 * C * 4); p.scal = take(16 * 4 * 4);
  p.gth = take(2 * C * M * 2); p.gtl = take(2 * C * M * 2); p.uth = take(k * C * M * 2); p.utl = take(k * C * M * 2);
  p.gspec = take(M * (size_t)p.Mp * 4);
  p.total = off;
  return off;
}

static bool train_cfg_ok(const drb_train_config& c) {
  return c.batch > 0 && c.frames > 0 && c.pitches > 0 && (c.pitches % 4) == 0 && c.residual_channels > 0 && (c.residual_channels % 16) == 0 &&
         c.residual_layers > 0 && c.kernel_size > 0 && (c.kernel_size & 1) && c.dilation_base > 0 && c.dilation_bound > 0 && c.n_mels > 0 &&
         c.timesteps > 0;
}

static bool params_ok(const drb_train_params* q, int L) {
  if (!q || !q->in_w || !q->in_b || !q->e1w || !q->e1b || !q->e2w || !q->e2b || !q->skw || !q->skb || !q->hdw || !q->hdb) return false;
  if (!q->wd || !q->bd || !q->wdp || !q->bdp || !q->wc || !q->bc || !q->wo || !q->bo) return false;
  for (int l = 0; l < L; ++l)
    if (!q->wd[l] || !q->bd[l] || !q->wdp[l] || !q->bdp[l] || !q->wc[l] || !q->bc[l] || !q->wo[l] || !q->bo[l]) return false;
  return true;
}

#define TR(expr) do { int _r = (expr); if (_r) return _r; } while (0)

// grad[r][c] += scale * sum_s part[s][r][c]   (c < cols_used; part rows have `cols` floats)
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, size_t stride, int rows, int cols, int cols_used,
                                     float* __restrict__ grad, int ld_grad, float scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols_used) return;
  const int r = (int)(i / cols_used), c = (int)(i - (size_t)r * cols_used);
  const float* q = part + (size_t)r * cols + c;
  float acc = 0.f;
  for (int sN = 0; sN < splits; ++sN) acc += q[(size_t)sN * stride];
  grad[(size_t)r * ld_grad + c] += scale * acc;
}

// dW[rows][cols_used] += out_scale * G^T X over the rolls x frames (G [M][rows] scaled by gscale = {S, 1/S}; X [M][cols]) on the tensor
// cores: transposed f16e5 operand pairs in the conv weight gradient's buffers, split-K zgemm, fp32 reduction.  g_ready: G^T is in place.
static int wgrad_splitk(drb_train* p, const drb_train::SplitK& q, const float* G, int ldg, int rows, const float* gscale, bool g_ready,
                        const float* X, int ldx, int cols, float* dW, int ld_dw, int cols_used, float out_scale, cudaStream_t s) {
  const int T = p->cfg.frames, M = p->cfg.batch * T;
  if (!g_ready) TR(launch_split_pair_T(G, ldg, nullptr, 0, T, 1, 1, gscale, p->ws + p->gth, p->ws + p->gtl, M, rows, 0, s));
  TR(launch_split_pair_T(X, ldx, nullptr, 0, T, 1, 1, nullptr, p->ws + p->uth, p->ws + p->utl, M, cols, 1, s));
  UmmaZGemm zg;
  zg.pair = 1; zg.persistent = 0; zg.NB = q.splits; zg.T = rows; zg.C = cols; zg.prec = 3; zg.mode = 4; zg.groups = 1; zg.z_group0 = 0; zg.group_stride = 0;
  zg.inv_scale = gscale + 1; zg.w_h = &p->m_uth; zg.w_l = &p->m_utl; zg.out32 = &q.out; zg.bias = nullptr; zg.dnext = nullptr;
  zg.a_h = &p->m_gth; zg.a_l = &p->m_gtl; zg.nslabs64 = q.slabs; zg.ksplit = 1;
  UmmaMaps dummy;
  dummy.xh = p->m_gth; dummy.xl = p->m_gtl; dummy.zh = p->m_gth; dummy.zl = p->m_gtl;
  TR(launch_umma_zgemm(dummy, zg, s));
  const size_t n = (size_t)rows * cols_used;
  splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p->at<float>(p->wtmp), q.splits, (size_t)rows * cols, rows, cols, cols_used, dW,
                                                                    ld_dw, out_scale);
  DRB_LAUNCH_CHECK();
  return 0;
}

static inline unsigned nblk(size_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }

extern "C" {

size_t drb_train_workspace_bytes(const drb_train_config* cfg) {
  if (!cfg || !train_cfg_ok(*cfg)) { set_error("invalid drb_train_config"); return 0; }
  drb_train tmp; tmp.cfg = *cfg;
  return train_layout(tmp);
}

int drb_train_create(drb_train** out, const drb_train_config* cfg, void* workspace, size_t ws_bytes, void* stream) {
  if (!out || !cfg || !workspace) { set_error("null argument"); return DRB_E_INVALID; }
  if (!train_cfg_ok(*cfg)) { set_error("invalid drb_train_config"); return DRB_E_INVALID; }
  drb_train* p = new drb_train();
  p->cfg = *cfg; p->ws = (char*)workspace;
  if (ws_bytes < train_layout(*p) || ((uintptr_t)workspace & 255)) {
    set_error("train workspace: need %zu bytes 256-aligned, got %zu", p->total, ws_bytes);
    delete p; return DRB_E_WORKSPACE;
  }
  for (int i = 0; i < cfg->residual_layers; ++i) {   // dilation_base ** (i % dilation_bound)   model/diffwave.py:624
    int d = 1;
    for (int j = 0; j < i % cfg->dilation_bound; ++j) d *= cfg->dilation_base;
    p->dil.push_back(d);
  }
  iota_kernel<<<nblk(cfg->batch), 256, 0, (cudaStream_t)stream>>>(p->at<int>(p->iota), cfg->batch);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("train_create: %s", cudaGetErrorString(e)); delete p; return (int)e; }
  {
    const char* env = getenv("DRB_TRAIN_TC");
    const int C = cfg->residual_channels, B = cfg->batch, T = cfg->frames, k = cfg->kernel_size, Mp = p->Mp;
    p->tc = (C % 256) == 0 && !(env && env[0] == '0');
    if (env && env[0] >= '1' && env[0] <= '9') p->tc_mask = atoi(env);
    const char* et = getenv("DRB_TRAIN_TAPS");
    if (et && atoi(et) >= 1) p->taps_per_launch = atoi(et);
    const char* ed = getenv("DRB_TRAIN_DGRAD_TAPS");
    if (ed && atoi(ed) >= 1) p->dgrad_taps = atoi(ed);
    if (p->tc) {
      int r = 0;
      auto mk3 = [&](CUtensorMap* m, size_t off, int d0, int dtype) { if (!r) r = make_tmap_3d(m, p->ws + off, B, T, d0, 128, dtype); };
      auto mk2 = [&](CUtensorMap* m, size_t off, int rows, uint64_t cols, int dtype) { if (!r) r = make_tmap_2d(m, p->ws + off, rows, cols, 128, dtype); };
      mk3(&p->m_uh, p->uh, C, 2); mk3(&p->m_ul, p->ul, 2 * C, 3);
      mk3(&p->m_gh, p->gh, 2 * C, 2); mk3(&p->m_gl, p->gl, 4 * C, 3);
      mk3(&p->m_sh, p->sph, Mp, 2); mk3(&p->m_sl, p->spl, 2 * Mp, 3);
      mk2(&p->m_wfh, p->wh, 2 * C, (uint64_t)k * C, 2); mk2(&p->m_wfl, p->wl, 2 * C, (uint64_t)2 * k * C, 3);
      mk2(&p->m_wgh, p->wh, C, (uint64_t)k * 2 * C, 2); mk2(&p->m_wgl, p->wl, C, (uint64_t)2 * k * 2 * C, 3);
      mk2(&p->m_wch, p->wch, 2 * C, Mp, 2); mk2(&p->m_wcl, p->wcl, 2 * C, (uint64_t)2 * Mp, 3);
      mk3(&p->m_ul5, p->ul, C, 2); mk3(&p->m_sl5, p->spl, Mp, 2);
      mk2(&p->m_wfl5, p->wl, 2 * C, (uint64_t)k * C, 2); mk2(&p->m_wcl5, p->wcl, 2 * C, Mp, 2);
      mk2(&p->m_woh, p->wh, 2 * C, C, 2); mk2(&p->m_wol5, p->wl, 2 * C, C, 2);
      mk2(&p->m_woTh, p->wh, C, 2 * C, 2); mk2(&p->m_woTl, p->wl, C, (uint64_t)4 * C, 3);
      const uint64_t Mr = (uint64_t)B * T;
      p->tc_wgrad = (T % 64) == 0 && ((uint64_t)k * C) % 256 == 0;
      if (p->tc_wgrad) {
        if (!r) r = make_tmap_3d(&p->m_gth, p->ws + p->gth, 1, 2 * C, Mr, 128, 2);
        if (!r) r = make_tmap_3d(&p->m_gtl, p->ws + p->gtl, 1, 2 * C, 2 * Mr, 128, 3);
        if (!r) r = make_tmap_2d(&p->m_uth, p->ws + p->uth, (uint64_t)k * C, Mr, 128, 2);
        if (!r) r = make_tmap_2d(&p->m_utl, p->ws + p->utl, (uint64_t)k * C, 2 * Mr, 128, 3);
        if (!r) r = make_tmap_3d(&p->m_dwt, p->ws + p->wtmp, 1, 2 * C, (uint64_t)k * C, 128, 1);
      }
      if (p->tc_wgrad) {
        int n_sm = 148;
        { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
        size_t cap = (size_t)2 * C * k * C;
        if (cap < (size_t)2 * C * Mp) cap = (size_t)2 * C * Mp;
        auto mk_split = [&](drb_train::SplitK& q, int rows, int cols) {
          const int total = (int)(Mr / 64), tiles = (rows / 128) * (cols / 256);
          int best = 0;
          if (rows % 256 == 0 && cols % 256 == 0 && rows <= 2 * C && cols <= k * C)   // even row-tile count (CTA pairs), whole N blocks
            for (int sn = 1; sn <= total; ++sn)
              if (total % sn == 0 && sn * tiles <= n_sm && (size_t)sn * rows * cols <= cap) best = sn;
          q.splits = best; q.slabs = best ? total / best : 0;
          if (best && !r) r = make_tmap_3d(&q.out, p->ws + p->wtmp, best, rows, cols, 128, 1);
        };
        mk_split(p->sk_wo, 2 * C, C); mk_split(p->sk_wc, 2 * C, Mp); mk_split(p->sk_sk, C, C);
      }
      if (r) { delete p; return r; }
    }
  }
  *out = p;
  return 0;
}

void drb_train_destroy(drb_train* p) { delete p; }

// Ask the following drb_train_backward calls for d loss / d spec as well: g_spec [B][n_mels][T] fp32 device, OVERWRITTEN by every
// backward (the sum over the layers of g_y . W_c, model/diffwave.py:143); NULL switches it off again.  The spectrogram is data
// for condition='fixed'; condition='trainable_spec' conditions the dropped rolls on a parameter (:695-699) whose gradient is
// the sum of this tensor over those rolls.
int drb_train_set_spec_grad(drb_train* p, float* g_spec) {
  if (!p) return DRB_E_INVALID;
  p->g_spec_out = g_spec;
  return 0;
}

// Forward pass in training form.  x_t [B][T][F]; spec [B][n_mels][T] exactly as the network sees it (normalised log-mel with the
// spec-dropout rows and masks already set to -1); steps [B] int32; emb_table [timesteps][128].  pred [B][T][F].
int drb_train_forward(drb_train* p, const drb_train_params* w, const float* x_t, const float* spec, const int32_t* steps,
                      const float* emb_table, float* pred, void* stream) {
  if (!p || !x_t || !spec || !steps || !emb_table || !pred || !params_ok(w, p->cfg.residual_layers)) {
    set_error("train_forward: null argument"); return DRB_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const drb_train_config& c = p->cfg;
  const int B = c.batch, T = c.frames, C = c.residual_channels, L = c.residual_layers, k = c.kernel_size, F = c.pitches, Mp = p->Mp;
  const size_t M = (size_t)B * T;
  float* wtmp = p->at<float>(p->wtmp);
  {  // spectrogram to time-major rows
    dim3 grid((T + 31) / 32, (Mp + 31) / 32, B), block(32, 8);
    spec_to_rows_kernel<<<grid, block, 0, s>>>(spec, p->at<float>(p->spec), B, c.n_mels, T, Mp);
    DRB_LAUNCH_CHECK();
    if (p->tc && (p->tc_mask & 1))
      TR(launch_split_pair(p->at<float>(p->spec), Mp, nullptr, 0, T, nullptr, p->ws + p->sph, p->ws + p->spl, (int)M, Mp, s, 5));
  }
  {  // diffusion embedding: table lookup, silu(projection1), silu(projection2)      model/diffwave.py:66-75
    gather_rows_kernel<<<nblk((size_t)B * 128), 256, 0, s>>>(emb_table, steps, p->at<float>(p->e0), B, 128);
    DRB_LAUNCH_CHECK();
    SimtGemm g;
    g.A = p->at<float>(p->e0); g.lda = 128; g.T = 1; g.Ck = 128; g.W = w->e1w; g.ldw = 128; g.bias = w->e1b; g.C = p->at<float>(p->p1); g.ldc = 512; g.M = B; g.N = 512; g.skinny = 1;
    TR(launch_simt_gemm(g, s));
    g.act = 2; g.C = p->at<float>(p->s1);
    TR(launch_simt_gemm(g, s));
    SimtGemm h;
    h.A = p->at<float>(p->s1); h.lda = 512; h.T = 1; h.Ck = 512; h.W = w->e2w; h.ldw = 512; h.bias = w->e2b; h.C = p->at<float>(p->p2); h.ldc = 512; h.M = B; h.N = 512; h.skinny = 1;
    TR(launch_simt_gemm(h, s));
    h.act = 2; h.C = p->at<float>(p->emb);
    TR(launch_simt_gemm(h, s));
  }
  {  // x_0 = relu(input_projection(x_t))      model/diffwave.py:667-668
    SimtGemm g;
    g.A = x_t; g.lda = F; g.T = T; g.Ck = F; g.W = w->in_w; g.ldw = F; g.bias = w->in_b; g.act = 1;
    g.C = p->at<float>(p->xs); g.ldc = C; g.M = (int)M; g.N = C;
    TR(launch_simt_gemm(g, s));
  }
  for (int l = 0; l < L; ++l) {
    float* x_l = p->at<float>(p->xs) + (size_t)l * M * C;
    float* y_l = p->at<float>(p->ys) + (size_t)l * M * 2 * C;
    float* z_l = p->at<float>(p->zs) + (size_t)l * M * C;
    float* d_l = p->at<float>(p->dl) + (size_t)l * B * C;
    SimtGemm d;  // diffusion_projection(emb)   model/diffwave.py:137
    d.A = p->at<float>(p->emb); d.lda = 512; d.T = 1; d.Ck = 512; d.W = w->wdp[l]; d.ldw = 512; d.bias = w->bdp[l]; d.C = d_l; d.ldc = C; d.M = B; d.N = C; d.skinny = 1;
    TR(launch_simt_gemm(d, s));
    TR(launch_repack_conv_fp32(w->wd[l], wtmp, 2 * C, C, k, s));   // [2C][C][k] -> tap-major [2C][k*C]
    if (p->tc && (p->tc_mask & 1)) {
      // dilated_conv(x + d) + conditioner_projection(spec) (:138-144) as ONE tcgen05 implicit GEMM: 9 x C/64 tap slabs + Mp/64
      // conditioner slabs into the same accumulator, fp32 pre-activations out
      float* sw = p->at<float>(p->scal);
      TR(launch_split_pair(x_l, C, d_l, C, T, nullptr, p->ws + p->uh, p->ws + p->ul, (int)M, C, s, 5));
      TR(launch_weight_scale(wtmp, (size_t)2 * C * k * C, w->wc[l], (size_t)2 * C * c.n_mels, sw, 1.f, s));
      TR(launch_repack_split(wtmp, p->ws + p->wh, p->ws + p->wl, 2 * C, k * C, k * C, 0, 5, sw, s));
      TR(launch_repack_split(w->wc[l], p->ws + p->wch, p->ws + p->wcl, 2 * C, c.n_mels, Mp, 0, 5, sw, s));
      add2_kernel<<<nblk(2 * C), 256, 0, s>>>(w->bd[l], w->bc[l], p->at<float>(p->bnat), 2 * C);
      DRB_LAUNCH_CHECK();
      UmmaConvLin cv;
      cv.ah = &p->m_uh; cv.al = &p->m_ul5; cv.wh = &p->m_wfh; cv.wl = &p->m_wfl5; cv.sh = &p->m_sh; cv.sl = &p->m_sl5; cv.wch = &p->m_wch; cv.wcl = &p->m_wcl5;
      cv.prec = 4;
      cv.NB = B; cv.T = T; cv.Cin = C; cv.Nout = 2 * C; cv.taps = k; cv.dil = p->dil[l]; cv.Mp = Mp; cv.inv_scale = sw + 1;
      cv.bias = p->at<float>(p->bnat); cv.out = y_l; cv.ldo = 2 * C;
      // one launch per tap (the first carries the bias, the last the conditioner slabs): accumulation chains of C (+ Mp) instead
      // of k*C + Mp terms, summed across launches in exact fp32
      // (one kernel: the persistent linear conv runs the passes of a tile back to back; DRB_LIN_PERS=0 launches once per pass)
      cv.tap_lo = 0; cv.tap_n = p->taps_per_launch < k ? p->taps_per_launch : k; cv.tap_span = k; cv.accumulate = 0;
      cv.scratch = wtmp; cv.scratch_bytes = p->wtmp_bytes;   // wtmp is free again: the weights were split into wh / wl above
      TR(launch_umma_conv_lin(cv, s));
    } else {
    SimtGemm g;  // dilated_conv(x + d)   :138-139
    g.A = x_l; g.lda = C; g.T = T; g.taps = k; g.dil = p->dil[l]; g.Ck = C; g.addvec = d_l; g.addvec_steps = p->at<int>(p->iota);
    g.addvec_mod = B; g.addvec_stride = C; g.W = wtmp; g.ldw = k * C; g.bias = w->bd[l]; g.C = y_l; g.ldc = 2 * C; g.M = (int)M; g.N = 2 * C;
    TR(launch_simt_gemm(g, s));
    TR(launch_pad_rows(w->wc[l], wtmp, 2 * C, c.n_mels, Mp, s));
    SimtGemm q;  // + conditioner_projection(spec)   :143-144
    q.A = p->at<float>(p->spec); q.lda = Mp; q.T = T; q.Ck = Mp; q.W = wtmp; q.ldw = Mp; q.bias = w->bc[l]; q.accumulate = 1;
    q.C = y_l; q.ldc = 2 * C; q.M = (int)M; q.N = 2 * C;
    TR(launch_simt_gemm(q, s));
    }
    TR(launch_gate(y_l, z_l, (int)M, C, s));
    if (p->tc && (p->tc_mask & 8)) {   // output_projection(z) (:149) = the conv kernel with one tap, f16x3 operands
      float* sw = p->at<float>(p->scal);
      TR(launch_split_pair(z_l, C, nullptr, 0, T, nullptr, p->ws + p->uh, p->ws + p->ul, (int)M, C, s, 5));
      TR(launch_weight_scale(w->wo[l], (size_t)2 * C * C, nullptr, 0, sw, 1.f, s));
      TR(launch_repack_split(w->wo[l], p->ws + p->wh, p->ws + p->wl, 2 * C, C, C, 0, 5, sw, s));
      UmmaConvLin cv;
      cv.ah = &p->m_uh; cv.al = &p->m_ul5; cv.wh = &p->m_woh; cv.wl = &p->m_wol5; cv.prec = 4;
      cv.NB = B; cv.T = T; cv.Cin = C; cv.Nout = 2 * C; cv.taps = 1; cv.dil = 1; cv.Mp = 0; cv.inv_scale = sw + 1;
      cv.bias = w->bo[l]; cv.out = p->at<float>(p->obuf); cv.ldo = 2 * C;
      TR(launch_umma_conv_lin(cv, s));
    } else {
    SimtGemm o;  // output_projection(z)   :149
    o.A = z_l; o.lda = C; o.T = T; o.Ck = C; o.W = w->wo[l]; o.ldw = C; o.bias = w->bo[l]; o.C = p->at<float>(p->obuf); o.ldc = 2 * C; o.M = (int)M; o.N = 2 * C;
    TR(launch_simt_gemm(o, s));
    }
    const int do_res = l < L - 1;
    res_skip_fwd_kernel<<<nblk(M * C / 4), 256, 0, s>>>(p->at<float>(p->obuf), x_l, do_res ? x_l + M * C : nullptr, p->at<float>(p->skip), M, C,
                                                        l == 0, do_res);
    DRB_LAUNCH_CHECK();
  }
  {  // h = relu(skip_projection(skip / sqrt(L))) ; pred = output_projection(h)      :680-685
    SimtGemm g;
    g.A = p->at<float>(p->skip); g.lda = C; g.T = T; g.Ck = C; g.a_div = sqrtf((float)L); g.W = w->skw; g.ldw = C; g.bias = w->skb; g.act = 1;
    g.C = p->at<float>(p->hbuf); g.ldc = C; g.M = (int)M; g.N = C;
    TR(launch_simt_gemm(g, s));
    SimtGemm o;
    o.A = p->at<float>(p->hbuf); o.lda = C; o.T = T; o.Ck = C; o.W = w->hdw; o.ldw = C; o.bias = w->hdb; o.C = pred; o.ldc = F; o.M = (int)M; o.N = F;
    TR(launch_simt_gemm(o, s));
  }
  p->fwd_done = true;
  return 0;
}

// Backward pass of the last drb_train_forward.  g_pred [B][T][F] = d loss / d pred.  Every tensor of `grads` (same shapes as the
// parameters) receives its gradient; accumulate != 0 adds to what is there (second forward of a two-dataset batch).
// g_x_t (optional) [B][T][F] = d loss / d x_t.  The saved pre-activations y_l are overwritten (one backward per forward).
int drb_train_backward(drb_train* p, const drb_train_params* w, const drb_train_params* gr, const float* x_t, const float* g_pred,
                       int32_t accumulate, float* g_x_t, void* stream) {
  if (!p || !x_t || !g_pred || !params_ok(w, p->cfg.residual_layers) || !params_ok(gr, p->cfg.residual_layers)) {
    set_error("train_backward: null argument"); return DRB_E_INVALID;
  }
  if (!p->fwd_done) { set_error("train_backward: no forward pass to differentiate"); return DRB_E_STATE; }
  p->fwd_done = false;
  cudaStream_t s = (cudaStream_t)stream;
  const drb_train_config& c = p->cfg;
  const int B = c.batch, T = c.frames, C = c.residual_channels, L = c.residual_layers, k = c.kernel_size, F = c.pitches, Mp = p->Mp;
  const size_t M = (size_t)B * T;
  float* wtmp = p->at<float>(p->wtmp);
  if (!accumulate) {
    auto zero = [&](float* q, size_t n) { return cudaMemsetAsync(q, 0, n * 4, s); };
    DRB_CUDA(zero(gr->in_w, (size_t)C * F)); DRB_CUDA(zero(gr->in_b, C));
    DRB_CUDA(zero(gr->e1w, 512 * 128)); DRB_CUDA(zero(gr->e1b, 512)); DRB_CUDA(zero(gr->e2w, 512 * 512)); DRB_CUDA(zero(gr->e2b, 512));
    DRB_CUDA(zero(gr->skw, (size_t)C * C)); DRB_CUDA(zero(gr->skb, C)); DRB_CUDA(zero(gr->hdw, (size_t)F * C)); DRB_CUDA(zero(gr->hdb, F));
    for (int l = 0; l < L; ++l) {
      DRB_CUDA(zero(gr->wd[l], (size_t)2 * C * C * k)); DRB_CUDA(zero(gr->bd[l], 2 * C));
      DRB_CUDA(zero(gr->wdp[l], (size_t)C * 512)); DRB_CUDA(zero(gr->bdp[l], C));
      DRB_CUDA(zero(gr->wc[l], (size_t)2 * C * c.n_mels)); DRB_CUDA(zero(gr->bc[l], 2 * C));
      DRB_CUDA(zero(gr->wo[l], (size_t)2 * C * C)); DRB_CUDA(zero(gr->bo[l], 2 * C));
    }
  }
  float* gspec = p->g_spec_out ? p->at<float>(p->gspec) : nullptr;
  if (gspec) DRB_CUDA(cudaMemsetAsync(gspec, 0, M * (size_t)Mp * 4, s));
  float* h = p->at<float>(p->hbuf);
  float* gx = p->at<float>(p->gx);        // g_x_{l+1}
  float* gu = p->at<float>(p->gu);        // g_h, then g_u of the current layer
  float* gz = p->at<float>(p->gz);
  float* gskip = p->at<float>(p->gskip);
  float* go = p->at<float>(p->obuf);      // [M][2C]
  float* emb = p->at<float>(p->emb);
  float* gemb = p->at<float>(p->gemb);
  float* gd = p->at<float>(p->gd);
  const float sqrtL = sqrtf((float)L);
  // ---- head: output_projection, ReLU, skip_projection(skip / sqrt(L))      model/diffwave.py:680-685 ----
  {
    Wgrad a;
    a.G = g_pred; a.ldg = F; a.X = h; a.ldx = C; a.M = (int)M; a.T = T; a.N = F; a.Ck = C; a.dW = gr->hdw; a.sn = C;
    TR(launch_wgrad(a, s));
    TR(launch_colsum(g_pred, F, F, 1, (int)M, gr->hdb, F, s));
    TR(launch_transpose(w->hdw, wtmp, F, C, s));                    // [F][C] -> [C][F]
    SimtGemm g;                                                     // g_h = g_pred . W_out
    g.A = g_pred; g.lda = F; g.T = T; g.Ck = F; g.W = wtmp; g.ldw = F; g.C = gu; g.ldc = C; g.M = (int)M; g.N = C;
    TR(launch_simt_gemm(g, s));
    relu_bwd_kernel<<<nblk(M * C / 4), 256, 0, s>>>(gu, h, M * C / 4, 1.f);
    DRB_LAUNCH_CHECK();
    if (p->tc && (p->tc_mask & 16) && p->sk_sk.splits) {
      float* sg = p->at<float>(p->scal) + 4;
      TR(launch_weight_scale(gu, M * C, nullptr, 0, sg, 1.f, s));
      TR(wgrad_splitk(p, p->sk_sk, gu, C, C, sg, false, p->at<float>(p->skip), C, C, gr->skw, C, C, 1.f / sqrtL, s));
    } else {
    Wgrad b;
    b.G = gu; b.ldg = C; b.X = p->at<float>(p->skip); b.ldx = C; b.M = (int)M; b.T = T; b.N = C; b.Ck = C; b.dW = gr->skw; b.sn = C;
    b.out_scale = 1.f / sqrtL;
    TR(launch_wgrad(b, s));
    }
    TR(launch_colsum(gu, C, C, 1, (int)M, gr->skb, C, s));
    TR(launch_transpose(w->skw, wtmp, C, C, s));
    SimtGemm q;                                                     // g_skip = (g_hp . W_sp) / sqrt(L), the same for every layer
    q.A = gu; q.lda = C; q.T = T; q.Ck = C; q.a_div = sqrtL; q.W = wtmp; q.ldw = C; q.C = gskip; q.ldc = C; q.M = (int)M; q.N = C;
    TR(launch_simt_gemm(q, s));
  }
  // ---- residual layers, last to first      model/diffwave.py:134-151 ----
  for (int l = L - 1; l >= 0; --l) {
    float* x_l = p->at<float>(p->xs) + (size_t)l * M * C;
    float* y_l = p->at<float>(p->ys) + (size_t)l * M * 2 * C;       // becomes g_y in place
    float* z_l = p->at<float>(p->zs) + (size_t)l * M * C;
    float* d_l = p->at<float>(p->dl) + (size_t)l * B * C;
    const int has_res = l < L - 1;                                   // the last layer's residual output is never used (:676-680)
    make_go_kernel<<<nblk(M * C / 4), 256, 0, s>>>(gx, gskip, go, M, C, has_res);
    DRB_LAUNCH_CHECK();
    const bool tc_o = p->tc && (p->tc_mask & 8), tc_w1 = p->tc && (p->tc_mask & 16) && p->sk_wo.splits && p->sk_wc.splits;
    float* so = p->at<float>(p->scal) + 4;                           // {S_o, 1 / S_o}: power-of-two scale that lifts g_o into fp16's range
    if (tc_o || tc_w1) TR(launch_weight_scale(go, M * 2 * C, nullptr, 0, so, 1.f, s));
    if (tc_w1) {                                                     // output_projection.weight [2C][C][1]
      TR(wgrad_splitk(p, p->sk_wo, go, 2 * C, 2 * C, so, false, z_l, C, C, gr->wo[l], C, C, 1.f, s));
    } else {
    Wgrad a;                                                         // output_projection: weight [2C][C][1], bias
    a.G = go; a.ldg = 2 * C; a.X = z_l; a.ldx = C; a.M = (int)M; a.T = T; a.N = 2 * C; a.Ck = C; a.dW = gr->wo[l]; a.sn = C;
    TR(launch_wgrad(a, s));
    }
    TR(launch_colsum(go, 2 * C, 2 * C, 1, (int)M, gr->bo[l], 2 * C, s));
    TR(launch_transpose(w->wo[l], wtmp, 2 * C, C, s));               // [2C][C] -> [C][2C]
    if (tc_o) {                                                      // g_z = g_o . W_o: one-tap conv kernel over scaled f16e5 pairs
      float* sw = p->at<float>(p->scal) + 8; float* sc = p->at<float>(p->scal) + 12;
      TR(launch_split_pair(go, 2 * C, nullptr, 0, T, so, p->ws + p->gh, p->ws + p->gl, (int)M, 2 * C, s));
      TR(launch_weight_scale(wtmp, (size_t)2 * C * C, nullptr, 0, sw, 1.f, s));
      TR(launch_repack_split(wtmp, p->ws + p->wh, p->ws + p->wl, C, 2 * C, 2 * C, 0, 3, sw, s));
      mul_inv_kernel<<<1, 1, 0, s>>>(so, sw, sc);
      DRB_LAUNCH_CHECK();
      UmmaConvLin cv;
      cv.ah = &p->m_gh; cv.al = &p->m_gl; cv.wh = &p->m_woTh; cv.wl = &p->m_woTl;
      cv.NB = B; cv.T = T; cv.Cin = 2 * C; cv.Nout = C; cv.taps = 1; cv.dil = 1; cv.Mp = 0; cv.inv_scale = sc;
      cv.bias = nullptr; cv.out = gz; cv.ldo = C;
      TR(launch_umma_conv_lin(cv, s));
    } else {
    SimtGemm g;                                                      // g_z = g_o . W_o
    g.A = go; g.lda = 2 * C; g.T = T; g.Ck = 2 * C; g.W = wtmp; g.ldw = 2 * C; g.C = gz; g.ldc = C; g.M = (int)M; g.N = C;
    TR(launch_simt_gemm(g, s));
    }
    gate_bwd_kernel<<<nblk(M * C / 4), 256, 0, s>>>(y_l, gz, y_l, M, C);
    DRB_LAUNCH_CHECK();
    const float* gy = y_l;
    TR(launch_colsum(gy, 2 * C, 2 * C, 1, (int)M, gr->bd[l], 2 * C, s));     // dilated_conv.bias
    TR(launch_colsum(gy, 2 * C, 2 * C, 1, (int)M, gr->bc[l], 2 * C, s));     // conditioner_projection.bias (added to the same y)
    if (gspec) {   // d loss / d spec += g_y . W_c (the conditioner is a 1x1 conv, :143); fp32 CUDA cores, wtmp is free between the two products around it
      float* wcp = wtmp; float* wct = wtmp + (size_t)2 * C * Mp;
      TR(launch_pad_rows(w->wc[l], wcp, 2 * C, c.n_mels, Mp, s));             // [2C][n_mels] -> [2C][Mp]
      TR(launch_transpose(wcp, wct, 2 * C, Mp, s));                           // -> [Mp][2C]
      SimtGemm q;
      q.A = gy; q.lda = 2 * C; q.T = T; q.Ck = 2 * C; q.W = wct; q.ldw = 2 * C; q.accumulate = 1; q.C = gspec; q.ldc = Mp; q.M = (int)M; q.N = Mp;
      TR(launch_simt_gemm(q, s));
    }
    float* sg = p->at<float>(p->scal) + 4;                           // {S_g, 1 / S_g}: power-of-two scale that lifts g_y into fp16's range
    const bool tc_w = p->tc && p->tc_wgrad && (p->tc_mask & 4), tc_d = p->tc && (p->tc_mask & 2);
    if (tc_w || tc_d || tc_w1) TR(launch_weight_scale(gy, M * 2 * C, nullptr, 0, sg, 1.f, s));
    if (tc_w || tc_w1) TR(launch_split_pair_T(gy, 2 * C, nullptr, 0, T, 1, 1, sg, p->ws + p->gth, p->ws + p->gtl, (int)M, 2 * C, 0, s));
    if (tc_w1) {                                                     // conditioner_projection.weight [2C][n_mels][1], g_y^T shared with the conv
      TR(wgrad_splitk(p, p->sk_wc, gy, 2 * C, 2 * C, sg, true, p->at<float>(p->spec), Mp, Mp, gr->wc[l], c.n_mels, c.n_mels, 1.f, s));
    } else {
    Wgrad cw;                                                        // conditioner_projection.weight [2C][n_mels][1]
    cw.G = gy; cw.ldg = 2 * C; cw.X = p->at<float>(p->spec); cw.ldx = Mp; cw.M = (int)M; cw.T = T; cw.N = 2 * C; cw.Ck = c.n_mels;
    cw.dW = gr->wc[l]; cw.sn = c.n_mels;
    TR(launch_wgrad(cw, s));
    }
    if (tc_w) {
      // dilated_conv.weight on the tensor cores: one plain GEMM  dW[n][tap*C + c] = sum_m g_y^T[n][m] * im2col(x + d)^T[tap*C + c][m]
      // over transposed f16e5 operand pairs (K = rolls x frames), fp32 result in tap-major order, then added into [2C][C][k]
      TR(launch_split_pair_T(x_l, C, d_l, C, T, k, p->dil[l], nullptr, p->ws + p->uth, p->ws + p->utl, (int)M, C, 1, s));
      UmmaZGemm zg;
      zg.pair = 1; zg.persistent = 0; zg.NB = 1; zg.T = 2 * C; zg.C = k * C; zg.prec = 3; zg.mode = 4; zg.groups = 1; zg.z_group0 = 0; zg.group_stride = 0;
      zg.inv_scale = sg + 1; zg.w_h = &p->m_uth; zg.w_l = &p->m_utl; zg.out32 = &p->m_dwt; zg.bias = nullptr; zg.dnext = nullptr;
      zg.a_h = &p->m_gth; zg.a_l = &p->m_gtl; zg.nslabs64 = (int)(M / 64);
      UmmaMaps dummy;
      dummy.xh = p->m_gth; dummy.xl = p->m_gtl; dummy.zh = p->m_gth; dummy.zl = p->m_gtl;
      TR(launch_umma_zgemm(dummy, zg, s));
      const size_t n = (size_t)2 * C * C * k;
      untap_add_kernel<<<nblk(n), 256, 0, s>>>(wtmp, gr->wd[l], 2 * C, C, k);
      DRB_LAUNCH_CHECK();
    } else {
    Wgrad dw;                                                        // dilated_conv.weight [2C][C][k]: the forward's tap gather of x + d
    dw.G = gy; dw.ldg = 2 * C; dw.X = x_l; dw.ldx = C; dw.addvec = d_l; dw.av_stride = C; dw.M = (int)M; dw.T = T; dw.N = 2 * C; dw.Ck = C;
    dw.taps = k; dw.dil = p->dil[l]; dw.dW = gr->wd[l]; dw.sn = (long long)C * k; dw.sc = k; dw.st = 1;
    TR(launch_wgrad(dw, s));
    }
    {                                                                // g_u = transposed dilated conv of g_y
      const size_t n = (size_t)2 * C * C * k;
      repack_dgrad_kernel<<<nblk(n), 256, 0, s>>>(w->wd[l], wtmp, 2 * C, C, k);
      DRB_LAUNCH_CHECK();
      if (tc_d) {
        // the same tcgen05 kernel with 2C input channels: g_y is scaled into fp16's range by a per-tensor power of two first
        // (gradients of a mean loss sit around 1e-6), the weights by theirs; the epilogue divides by the product
        float* sw = p->at<float>(p->scal) + 8; float* sc = p->at<float>(p->scal) + 12;
        TR(launch_split_pair(gy, 2 * C, nullptr, 0, T, sg, p->ws + p->gh, p->ws + p->gl, (int)M, 2 * C, s));
        TR(launch_weight_scale(wtmp, n, nullptr, 0, sw, 1.f, s));
        TR(launch_repack_split(wtmp, p->ws + p->wh, p->ws + p->wl, C, k * 2 * C, k * 2 * C, 0, 3, sw, s));
        mul_inv_kernel<<<1, 1, 0, s>>>(sg, sw, sc);
        DRB_LAUNCH_CHECK();
        UmmaConvLin cv;
        cv.ah = &p->m_gh; cv.al = &p->m_gl; cv.wh = &p->m_wgh; cv.wl = &p->m_wgl;
        cv.NB = B; cv.T = T; cv.Cin = 2 * C; cv.Nout = C; cv.taps = k; cv.dil = p->dil[l]; cv.Mp = 0; cv.inv_scale = sc;
        cv.bias = nullptr; cv.out = gu; cv.ldo = C;
        // one pass per tap, summed in fp32 in g_u (chains of 2C instead of 9 * 2C terms: worst gradient error 6e-5 -> 3.5e-5 at the same
        // speed).  No scratch, i.e. whole items per pair: cutting the 80 items into balanced pass ranges was measured 2 ms SLOWER per
        // step -- this product streams ~1.1 GB of operands from L2 per launch, and pairs that no longer walk the same tap in lockstep
        // lose the sharing of those reads (the forward conv, 160 items, gains 0.5 ms from the same cut)
        cv.tap_lo = 0; cv.tap_n = p->dgrad_taps < k ? p->dgrad_taps : k; cv.tap_span = k;
        TR(launch_umma_conv_lin(cv, s));
      } else {
      SimtGemm u;
      u.A = gy; u.lda = 2 * C; u.T = T; u.taps = k; u.dil = p->dil[l]; u.Ck = 2 * C; u.W = wtmp; u.ldw = k * 2 * C; u.C = gu; u.ldc = C;
      u.M = (int)M; u.N = C;
      TR(launch_simt_gemm(u, s));
      }
    }
    // diffusion_projection: d enters only through the conv input x + d (broadcast over the frames of a roll)
    DRB_CUDA(cudaMemsetAsync(gd, 0, (size_t)B * C * 4, s));
    TR(launch_colsum(gu, C, C, B, T, gd, C, s));
    Wgrad pw;
    pw.G = gd; pw.ldg = C; pw.X = emb; pw.ldx = 512; pw.M = B; pw.T = 1; pw.N = C; pw.Ck = 512; pw.dW = gr->wdp[l]; pw.sn = 512;
    TR(launch_wgrad(pw, s));
    TR(launch_colsum(gd, C, C, 1, B, gr->bdp[l], C, s));
    TR(launch_transpose(w->wdp[l], wtmp, C, 512, s));                // [C][512] -> [512][C]
    SimtGemm e;
    e.A = gd; e.lda = C; e.T = 1; e.Ck = C; e.W = wtmp; e.ldw = C; e.C = gemb; e.ldc = 512; e.M = B; e.N = 512; e.accumulate = l < L - 1; e.skinny = 1;
    TR(launch_simt_gemm(e, s));
    if (has_res) {
      add_res_grad_kernel<<<nblk(M * C / 4), 256, 0, s>>>(gu, gx, M * C / 4);
      DRB_LAUNCH_CHECK();
    }
    float* t = gx; gx = gu; gu = t;                                  // gx now holds g_x_l
  }
  // ---- input projection + ReLU      model/diffwave.py:667-668 ----
  {
    relu_bwd_kernel<<<nblk(M * C / 4), 256, 0, s>>>(gx, p->at<float>(p->xs), M * C / 4, 1.f);
    DRB_LAUNCH_CHECK();
    Wgrad a;
    a.G = gx; a.ldg = C; a.X = x_t; a.ldx = F; a.M = (int)M; a.T = T; a.N = C; a.Ck = F; a.dW = gr->in_w; a.sn = F;
    TR(launch_wgrad(a, s));
    TR(launch_colsum(gx, C, C, 1, (int)M, gr->in_b, C, s));
    if (g_x_t) {
      TR(launch_transpose(w->in_w, wtmp, C, F, s));                  // [C][F] -> [F][C]
      SimtGemm g;
      g.A = gx; g.lda = C; g.T = T; g.Ck = C; g.W = wtmp; g.ldw = C; g.C = g_x_t; g.ldc = F; g.M = (int)M; g.N = F;
      TR(launch_simt_gemm(g, s));
    }
  }
  if (gspec) {   // the layers' sum, back in the module's layout
    dim3 grid((T + 31) / 32, (Mp + 31) / 32, B), block(32, 8);
    rows_to_spec_kernel<<<grid, block, 0, s>>>(gspec, p->g_spec_out, B, c.n_mels, T, Mp, 0);
    DRB_LAUNCH_CHECK();
  }
  // ---- diffusion embedding MLP      model/diffwave.py:66-75 ----
  {
    float* gs = p->at<float>(p->gsmall);
    silu_bwd_kernel<<<nblk((size_t)B * 512), 256, 0, s>>>(gemb, p->at<float>(p->p2), (size_t)B * 512);
    DRB_LAUNCH_CHECK();
    Wgrad a;
    a.G = gemb; a.ldg = 512; a.X = p->at<float>(p->s1); a.ldx = 512; a.M = B; a.T = 1; a.N = 512; a.Ck = 512; a.dW = gr->e2w; a.sn = 512;
    TR(launch_wgrad(a, s));
    TR(launch_colsum(gemb, 512, 512, 1, B, gr->e2b, 512, s));
    TR(launch_transpose(w->e2w, wtmp, 512, 512, s));
    SimtGemm g;
    g.A = gemb; g.lda = 512; g.T = 1; g.Ck = 512; g.W = wtmp; g.ldw = 512; g.C = gs; g.ldc = 512; g.M = B; g.N = 512; g.skinny = 1;
    TR(launch_simt_gemm(g, s));
    silu_bwd_kernel<<<nblk((size_t)B * 512), 256, 0, s>>>(gs, p->at<float>(p->p1), (size_t)B * 512);
    DRB_LAUNCH_CHECK();
    Wgrad b;
    b.G = gs; b.ldg = 512; b.X = p->at<float>(p->e0); b.ldx = 128; b.M = B; b.T = 1; b.N = 512; b.Ck = 128; b.dW = gr->e1w; b.sn = 128;
    TR(launch_wgrad(b, s));
    TR(launch_colsum(gs, 512, 512, 1, B, gr->e1b, 512, s));
  }
  return 0;
}

// d loss / d prediction of p_losses (task/diffusion.py:792-802; loss_type 0 l1, 1 l2, 2 huber), mean over all n elements, optionally
// times a per-roll factor (training mode 'ex_0').  label, pred, g_pred: n floats; roll_scale: n / per_roll floats or NULL.
int drb_loss_grad(const float* label, const float* pred, float* g_pred, size_t n, size_t per_roll, int32_t loss_type,
                  const float* roll_scale, void* stream) {
  if (!label || !pred || !g_pred || n == 0 || per_roll == 0 || loss_type < 0 || loss_type > 2) { set_error("loss_grad: bad argument"); return DRB_E_INVALID; }
  loss_grad_kernel<<<nblk(n), 256, 0, (cudaStream_t)stream>>>(label, pred, g_pred, n, per_roll, loss_type, roll_scale, 1.f / (float)n);
  DRB_LAUNCH_CHECK();
  return 0;
}

// One Adam update of a parameter tensor (torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8, weight_decay 0, amsgrad off).
// step = 1 for the first update.  task/diffusion.py:1057-1059
int drb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int32_t step, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || step < 1) { set_error("adam: bad argument"); return DRB_E_INVALID; }
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<nblk(n), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, (float)bc1,
                                                         (float)sqrt(bc2));
  DRB_LAUNCH_CHECK();
  return 0;
}

int drb_adam_step_multi(const void* table, const void* block_map, int32_t n_blocks, float lr, float beta1, float beta2, float eps,
                        float weight_decay, int32_t step, void* stream) {
  if (!table || !block_map || n_blocks < 0 || step < 1) { set_error("adam_multi: bad argument"); return DRB_E_INVALID; }
  if (n_blocks == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_multi_kernel<<<(unsigned)n_blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(table),
                                                                           reinterpret_cast<const int2*>(block_map), lr, beta1, beta2, eps,
                                                                           weight_decay, (float)bc1, (float)sqrt(bc2));
  DRB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

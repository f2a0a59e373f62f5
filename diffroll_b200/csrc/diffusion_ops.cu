// Elementwise / reduction kernels of the diffusion (validation) step around the network forward
// (forward-only part: SpecRollDiffusion.step, task/diffusion.py:651-763):
//   q_sample            task/diffusion.py:31-46     x_t = sqrt(abar_t) * x_0 + sqrt(1 - abar_t) * noise, t per roll
//   extract_x0          task/diffusion.py:49-65     x_0 = (x_t - sqrt(1 - abar_t) * eps) / sqrt(abar_t)
//   p_losses            task/diffusion.py:792-802   mean |a-b|, mean (a-b)^2, mean smooth_l1(a-b) (beta = 1)
//   Normalization       model/utils.py:21-32        per-roll min-max of the label roll ('imagewise'), NaN -> min value
// All HBM-bound: one coalesced float4 pass each (the loss: one pass + a 1-block finish).
#include "common.cuh"
#include "kernels.h"

namespace drb {

// mode 0: out = sa[t]*a + s1[t]*b            (q_sample: a = x_start, b = noise)
// mode 1: out = (a - s1[t]*b) / sa[t]        (extract_x0: a = x_t, b = epsilon)
__global__ void diffuse_kernel(const float* __restrict__ a, const float* __restrict__ b, const int* __restrict__ steps,
                               const float* __restrict__ sa, const float* __restrict__ s1, float* __restrict__ out,
                               int B, size_t n_per, int mode) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= (size_t)B * n_per) return;
  const int roll = (int)(i / n_per);           // n_per % 4 == 0: a float4 never straddles two rolls
  const int t = __ldg(steps + roll);
  const float ca = __ldg(sa + t), cb = __ldg(s1 + t);
  const float4 x = *reinterpret_cast<const float4*>(a + i), y = *reinterpret_cast<const float4*>(b + i);
  float4 o;
  if (mode == 0) {
    o.x = ca * x.x + cb * y.x; o.y = ca * x.y + cb * y.y; o.z = ca * x.z + cb * y.z; o.w = ca * x.w + cb * y.w;
  } else {
    o.x = (x.x - cb * y.x) / ca; o.y = (x.y - cb * y.y) / ca; o.z = (x.z - cb * y.z) / ca; o.w = (x.w - cb * y.w) / ca;
  }
  *reinterpret_cast<float4*>(out + i) = o;
}

constexpr int LOSS_BLOCKS = 592;   // 4 x 148 SMs

__device__ __forceinline__ float loss_term(float a, float b, int type) {
  const float d = a - b, ad = fabsf(d);
  if (type == 0) return ad;                                  // F.l1_loss
  if (type == 1) return d * d;                               // F.mse_loss
  return ad < 1.f ? 0.5f * d * d : ad - 0.5f;                // F.smooth_l1_loss, beta = 1
}

__global__ void __launch_bounds__(256) loss_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n,
                                                           int type, double* __restrict__ partial) {
  double acc = 0.0;
  const size_t n4 = n / 4, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 x = *reinterpret_cast<const float4*>(a + 4 * i), y = *reinterpret_cast<const float4*>(b + 4 * i);
    acc += (double)((loss_term(x.x, y.x, type) + loss_term(x.y, y.y, type)) + (loss_term(x.z, y.z, type) + loss_term(x.w, y.w, type)));
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - 4 * n4)) acc += (double)loss_term(a[4 * n4 + threadIdx.x], b[4 * n4 + threadIdx.x], type);
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) loss_finish_kernel(const double* __restrict__ partial, int nblocks, size_t n, float* __restrict__ out) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) acc += partial[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(red[0] / (double)n);
}

// one block per roll: min / max, then (x - min) / (max - min) * (hi - lo) + lo, NaN -> lo   (model/utils.py:25-32)
__global__ void __launch_bounds__(256) normalize_imagewise_kernel(const float* __restrict__ x, float* __restrict__ out, size_t n_per,
                                                                  float lo, float hi) {
  const float* src = x + (size_t)blockIdx.x * n_per;
  float* dst = out + (size_t)blockIdx.x * n_per;
  float mn = INFINITY, mx = -INFINITY;
  for (size_t i = threadIdx.x; i < n_per; i += 256) { const float v = src[i]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
  __shared__ float smn[256], smx[256];
  smn[threadIdx.x] = mn; smx[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { smn[threadIdx.x] = fminf(smn[threadIdx.x], smn[threadIdx.x + s]); smx[threadIdx.x] = fmaxf(smx[threadIdx.x], smx[threadIdx.x + s]); }
    __syncthreads();
  }
  mn = smn[0]; mx = smx[0];
  const float range = mx - mn;
  for (size_t i = threadIdx.x; i < n_per; i += 256) {
    float v = (src[i] - mn) / range;          // 0/0 -> NaN for a constant roll, like the reference
    v = v * (hi - lo) + lo;
    dst[i] = isnan(v) ? lo : v;
  }
}

}  // namespace drb

using namespace drb;

extern "C" {

static int diffuse(const float* a, const float* b, const int32_t* steps, const float* sa, const float* s1, float* out, int32_t B,
                   int64_t n_per, int mode, void* stream) {
  if (!a || !b || !steps || !sa || !s1 || !out || B <= 0 || n_per <= 0 || (n_per % 4)) {
    set_error("q_sample/extract_x0: bad argument (n_per must be a multiple of 4)"); return DRB_E_INVALID;
  }
  const size_t n4 = (size_t)B * (size_t)n_per / 4;
  diffuse_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b, steps, sa, s1, out, B, (size_t)n_per, mode);
  DRB_LAUNCH_CHECK();
  return 0;
}

int drb_q_sample(const float* x_start, const float* noise, const int32_t* steps, const float* sqrt_alphas_cumprod,
                 const float* sqrt_one_minus_alphas_cumprod, float* x_t, int32_t B, int64_t n_per, void* stream) {
  return diffuse(x_start, noise, steps, sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod, x_t, B, n_per, 0, stream);
}

int drb_extract_x0(const float* x_t, const float* epsilon, const int32_t* steps, const float* sqrt_alphas_cumprod,
                   const float* sqrt_one_minus_alphas_cumprod, float* x0, int32_t B, int64_t n_per, void* stream) {
  return diffuse(x_t, epsilon, steps, sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod, x0, B, n_per, 1, stream);
}

size_t drb_p_losses_scratch_bytes(void) { return LOSS_BLOCKS * sizeof(double); }

int drb_p_losses(const float* label, const float* prediction, int64_t n, int32_t loss_type, void* scratch, float* loss_out,
                 void* stream) {
  if (!label || !prediction || !scratch || !loss_out || n <= 0 || loss_type < 0 || loss_type > 2) {
    set_error("p_losses: bad argument"); return DRB_E_INVALID;
  }
  if (((uintptr_t)label | (uintptr_t)prediction) & 15) { set_error("p_losses: inputs must be 16-byte aligned"); return DRB_E_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  loss_partial_kernel<<<LOSS_BLOCKS, 256, 0, s>>>(label, prediction, (size_t)n, loss_type, (double*)scratch);
  DRB_LAUNCH_CHECK();
  loss_finish_kernel<<<1, 256, 0, s>>>((const double*)scratch, LOSS_BLOCKS, (size_t)n, loss_out);
  DRB_LAUNCH_CHECK();
  return 0;
}

int drb_normalize_imagewise(const float* x, float* out, int32_t B, int64_t n_per, float lo, float hi, void* stream) {
  if (!x || !out || B <= 0 || n_per <= 0) { set_error("normalize_imagewise: bad argument"); return DRB_E_INVALID; }
  normalize_imagewise_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, out, (size_t)n_per, lo, hi);
  DRB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

// Micro-benchmark (round-2 groundwork, NOT part of the product): sustained issue rate of the tcgen05.mma kinds on every
// SM of a B200, to decide whether a block-scaled fp4 correction product (kind::mxf4nvf4, K = 64 per instruction) would
// really cost half of the e5m2 correction product (kind::f8f6f4, K = 32) the gate kernel issues today.
//   M = 128, N = 256 per CTA (cta_group::1), one CTA per SM, operands = one 128-byte-swizzled K-slab resident in shared
//   memory (zeros; no loads inside the loop), accumulator = 256 TMEM columns, scale factors = 1.0 in TMEM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../diffroll_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"

using namespace drb;

namespace drb { void set_error(const char*, ...) {} void count_launch(int) {} }

__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// block-scaled instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptorBlockScaled): a_format [7,10),
// b_format [10,13), n_dim [17,23), scale_format [23] (0 = ue4m3, 1 = ue8m0), m_dim [24,29)
__host__ __device__ constexpr uint32_t idesc_bs(int M, int N, int afmt, int bfmt, int sf_ue8m0) {
  return ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)sf_ue8m0 << 23) | ((uint32_t)(M >> 4) << 24);
}

template <int KIND>
__device__ __forceinline__ void mma_one(uint32_t d, uint64_t a, uint64_t b, uint32_t sfa, uint32_t sfb) {
  if (KIND == 0) umma_bf16(d, a, b, make_idesc_fmt0(128, 256), 1u);            // kind::f16, fp16 inputs, K = 16
  if (KIND == 1) umma_f8(d, a, b, make_idesc_bf16(128, 256), 1u);              // kind::f8f6f4, e5m2 inputs, K = 32
  if (KIND == 2) {                                                             // kind::mxf8f6f4, e5m2, ue8m0 per 32, K = 32
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale.scale_vec::1X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc_bs(128, 256, 1, 1, 1)), "r"(1u), "r"(sfa), "r"(sfb) : "memory");
  }
  if (KIND == 3) {                                                             // kind::mxf4nvf4, e2m1, ue4m3 per 16, K = 64
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::mxf4nvf4.block_scale.scale_vec::4X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc_bs(128, 256, 1, 1, 0)), "r"(1u), "r"(sfa), "r"(sfb) : "memory");
  }
  if (KIND == 4) {                                                             // kind::mxf4, e2m1, ue8m0 per 32, K = 64
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc_bs(128, 256, 1, 1, 1)), "r"(1u), "r"(sfa), "r"(sfb) : "memory");
  }
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* cycles, unsigned long long* ns) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint32_t a0 = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((a0 + 1023u) & ~1023u) - a0);
  uint8_t* A = base;                    // 128 rows x 128 B
  uint8_t* B = base + 16384;            // 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 16384 + 32768);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  // scale factors = 1.0 (ue4m3 0x38 / ue8m0 0x7f) in columns 256..287, every lane quadrant
  const uint32_t one = (KIND == 3) ? 0x38383838u : 0x7f7f7f7fu;
  for (int c = 0; c < 32; ++c) tmem_st1(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c, one);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && elect_one()) {
    const uint32_t sfa = tmem + 256, sfb = tmem + 272;
    const uint32_t sa = smem_u32(A), sb = smem_u32(B);
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_one<KIND>(tmem, make_sw128_desc(sa + k * 32), make_sw128_desc(sb + k * 32), sfa, sfb);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    cycles[blockIdx.x] = c1 - c0;
    ns[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int KIND>
static int run(const char* name, int kdim, int iters, int n_sm) {
  long long* cyc; unsigned long long* ns;
  cudaMalloc(&cyc, n_sm * sizeof(long long)); cudaMalloc(&ns, n_sm * sizeof(unsigned long long));
  const int smem = 16384 + 32768 + 1024 + 64;
  cudaFuncSetAttribute((const void*)rate_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<KIND><<<n_sm, 128, smem>>>(1000, cyc, ns);           // warm-up
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("{\"kind\": \"%s\", \"error\": \"%s\"}\n", name, cudaGetErrorString(e)); return 1; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate_kernel<KIND><<<n_sm, 128, smem>>>(iters, cyc, ns);
  cudaEventRecord(e1);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("{\"kind\": \"%s\", \"error\": \"%s\"}\n", name, cudaGetErrorString(e)); return 1; }
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> hc(n_sm); std::vector<unsigned long long> hn(n_sm);
  cudaMemcpy(hc.data(), cyc, n_sm * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaMemcpy(hn.data(), ns, n_sm * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  double csum = 0, nmax = 0; for (int i = 0; i < n_sm; ++i) { csum += (double)hc[i]; if ((double)hn[i] > nmax) nmax = (double)hn[i]; }
  const double mmas = 4.0 * iters;
  const double flops = (double)n_sm * mmas * 2.0 * 128 * 256 * kdim;
  printf("{\"kind\": \"%s\", \"K\": %d, \"mma_per_cta\": %.0f, \"event_ms\": %.2f, \"cycles_per_mma\": %.1f, \"tflops\": %.1f, \"sm_mhz_effective\": %.0f}\n",
         name, kdim, mmas, ms, csum / n_sm / mmas, flops / (ms * 1e-3) / 1e12, csum / n_sm / (nmax * 1e-9) / 1e6);
  cudaFree(cyc); cudaFree(ns);
  return 0;
}

int main(int argc, char** argv) {
  int n_sm = 148; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  const int iters = argc > 1 ? atoi(argv[1]) : 2000000;
  int bad = 0;
  bad += run<0>("f16 (kind::f16, fp16 inputs)", 16, iters, n_sm);
  bad += run<1>("f8f6f4 (e5m2)", 32, iters, n_sm);
  bad += run<2>("mxf8f6f4.block_scale 1X (e5m2, ue8m0)", 32, iters, n_sm);
  bad += run<3>("mxf4nvf4.block_scale 4X (e2m1, ue4m3 per 16)", 64, iters, n_sm);
  bad += run<4>("mxf4.block_scale 2X (e2m1, ue8m0 per 32)", 64, iters, n_sm);
  return bad ? 1 : 0;
}

// Known-answer probe (round-2 groundwork for the nvfp4 correction product of the gate kernel; NOT part of the product).
//
// One CTA pair computes  D[256 x 256] = A[256 x 128] . B[256 x 128]^T  with  tcgen05.mma.cta_group::2.kind::mxf4nvf4
// .block_scale.scale_vec::4X  (e2m1 codes, one ue4m3 scale per 16 elements), two K = 64 instructions, exactly the way the
// gate kernel would issue its correction product:
//   * A rows come from a TMA-written, 64-byte-swizzled WINDOW of 192 rows per CTA, and the MMA reads the 128 rows that
//     start `shift` rows into it (descriptor start address + shift * 64 B): does the row-shifted view work under
//     SWIZZLE_64B the way it does under SWIZZLE_128B?
//   * the A-side scale factors of the shifted rows are gathered by a warp into the 512-byte "SF atom"
//     (byte offset 16*(r%32) + 4*(r/32) + kblock, cutlass/detail/sm100_blockscaled_layout.hpp) and copied to TMEM with
//     tcgen05.cp.cta_group::2.32x128b.warpx4; the B-side atoms (N = 256: two atoms per instruction) are precomputed.
//   * both CTAs hold their own A rows / SFA and ALL of SFB.
// Data patterns isolate what breaks: P0 everything 1.0; P1 random SFA; P2 random SFB; P3 random codes; P4 all random.
// The host decodes the codes under both nibble orders and reports which one matches.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../diffroll_b200/csrc nv4_probe.cu -o nv4_probe -lcuda
#include <cuda.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "common.cuh"

using namespace drb;
namespace drb { void set_error(const char*, ...) {} void count_launch(int) {} }

constexpr int WIN_ROWS = 192, ROW_B = 64, NROWS_B = 128;   // per CTA: A window 192 x 64 B, B half 128 x 64 B

__host__ __device__ constexpr uint32_t idesc_nv4(int M, int N) {   // e2m1 x e2m1, ue4m3 scales, K-major both
  return (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (0u << 23) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t make_sw64_desc(uint32_t saddr) {    // K-major, 64-byte swizzle: 8-row groups 512 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t saddr) {    // 32 rows x 16 B, rows 16 B apart, 8-row groups 128 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(16 >> 4) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void utccp_pair(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_nv4_pair(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, uint32_t sfa, uint32_t sfb) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::mxf4nvf4.block_scale.scale_vec::4X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb) : "memory");
}

struct alignas(64) Params {
  CUtensorMap amap, bmap;      // A windows [2*192 rows][64 B], B [256 rows][64 B], both SWIZZLE_64B, box 64 B x rows
  const uint8_t* sfw;          // [2][192][8]: window-row scale factors (4 for instruction 0, 4 for instruction 1)
  const uint8_t* sfb_atoms;    // [2 instr][2 atoms][512]
  float* out;                  // [256][256]
  int shift;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1) probe_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint32_t a0 = smem_u32(smem_raw);
  uint8_t* base = smem_raw + (((a0 + 1023u) & ~1023u) - a0);
  uint8_t* sA = base;                              // 192 x 64 = 12288
  uint8_t* sB = base + 12288;                      // 128 x 64 = 8192
  uint8_t* sSFW = base + 20480;                    // 192 x 8 = 1536
  uint8_t* sSFA = base + 22528;                    // 2 atoms x 512 (1024-aligned)
  uint8_t* sSFB = base + 23552;                    // 4 atoms x 512
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 25600);
  uint64_t* full = bars; uint64_t* done = bars + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) { mbar_init(full, 2); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc_pair(tptr, 512); tmem_relinquish_pair(); }
  // scale factors: window rows -> smem, B atoms -> smem (generic proxy), then the A atoms of the shifted rows
  for (int i = threadIdx.x; i < WIN_ROWS * 8 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(sSFW)[i] = reinterpret_cast<const uint32_t*>(p.sfw + (size_t)rank * WIN_ROWS * 8)[i];
  for (int i = threadIdx.x; i < 4 * 512 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(sSFB)[i] = reinterpret_cast<const uint32_t*>(p.sfb_atoms)[i];
  __syncthreads();
  if (warp == 2) {   // the "SF warp": atom_k[lane][q] = SF[row 32q + lane + shift][instruction k]
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint2 v = *reinterpret_cast<const uint2*>(sSFW + (size_t)(32 * q + lane + p.shift) * 8);
      *reinterpret_cast<uint32_t*>(sSFA + lane * 16 + q * 4) = v.x;
      *reinterpret_cast<uint32_t*>(sSFA + 512 + lane * 16 + q * 4) = v.y;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tptr;

  if (warp == 0 && elect_one()) {
    const uint32_t fb = mapa_cluster(smem_u32(full), 0);
    mbar_expect_tx_cluster(fb, WIN_ROWS * ROW_B + NROWS_B * ROW_B);
    tma_load_2d_pair(sA, &p.amap, fb, 0, (int)rank * WIN_ROWS);
    tma_load_2d_pair(sB, &p.bmap, fb, 0, (int)rank * NROWS_B);
  }
  if (warp == 1 && rank == 0 && elect_one()) {
    mbar_wait(full, 0);
    tc_fence_after();
    const uint32_t sfa_t = tmem + 256, sfb_t = tmem + 272;     // SFA: 4 columns per instruction; SFB: 8 per instruction
    utccp_pair(sfa_t, make_nosw_desc(smem_u32(sSFA)));
    utccp_pair(sfa_t + 4, make_nosw_desc(smem_u32(sSFA + 512)));
    for (int k = 0; k < 2; ++k)
      for (int h = 0; h < 2; ++h) utccp_pair(sfb_t + k * 8 + h * 4, make_nosw_desc(smem_u32(sSFB + (k * 2 + h) * 512)));
    const uint32_t aaddr = smem_u32(sA) + (uint32_t)p.shift * ROW_B;
    for (int k = 0; k < 2; ++k)
      umma_nv4_pair(tmem, make_sw64_desc(aaddr + k * 32), make_sw64_desc(smem_u32(sB) + k * 32), idesc_nv4(256, 256),
                    k ? 1u : 0u, sfa_t + k * 4, sfb_t + k * 8);
    umma_commit_pair(done);
  }
  if (warp >= 2) {   // 4 warps: one TMEM lane quarter each
    mbar_wait(done, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int row = (int)rank * 128 + q * 32 + lane;
    for (int c = 0; c < 8; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) p.out[(size_t)row * 256 + c * 32 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_pair(tmem, 512); }
}

// ---------------------------------------------------------------------------------------------------------------
static float e2m1_val(int c) { static const float t[8] = {0.f, 0.5f, 1.f, 1.5f, 2.f, 3.f, 4.f, 6.f}; float v = t[c & 7]; return (c & 8) ? -v : v; }
static float ue4m3_val(uint8_t b) {
  const int e = (b >> 3) & 15, m = b & 7;
  return e == 0 ? std::ldexp((float)m / 8.f, -6) : std::ldexp(1.f + (float)m / 8.f, e - 7);
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) { printf("no encode fn\n"); return 2; }
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  const int smem = 25600 + 64 + 1024;
  cudaFuncSetAttribute((const void*)probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  uint8_t *dA, *dB, *dSFW, *dSFB; float* dOut;
  cudaMalloc(&dA, 2 * WIN_ROWS * ROW_B); cudaMalloc(&dB, 256 * ROW_B); cudaMalloc(&dSFW, 2 * WIN_ROWS * 8); cudaMalloc(&dSFB, 4 * 512);
  cudaMalloc(&dOut, 256 * 256 * 4);
  Params prm;
  {
    cuuint64_t dims[2] = {ROW_B, 2 * WIN_ROWS}; cuuint64_t strides[1] = {ROW_B}; cuuint32_t box[2] = {ROW_B, WIN_ROWS}; cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&prm.amap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t dimsb[2] = {ROW_B, 256}; cuuint32_t boxb[2] = {ROW_B, NROWS_B};
    CUresult r2 = encode(&prm.bmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dB, dimsb, strides, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed %d %d\n", (int)r, (int)r2); return 2; }
  }
  prm.sfw = dSFW; prm.sfb_atoms = dSFB; prm.out = dOut;
  std::vector<uint8_t> hA(2 * WIN_ROWS * ROW_B), hB(256 * ROW_B), hSFW(2 * WIN_ROWS * 8), hSFBrow(256 * 8), hSFBat(4 * 512);
  std::vector<float> hOut(256 * 256);
  srand(1234);
  int fails = 0;
  const int shifts[5] = {0, 1, 8, 37, 64};
  for (int pat = 0; pat < 5; ++pat)
    for (int si = 0; si < 5; ++si) {
      const int shift = shifts[si];
      const bool rnd_sfa = pat == 1 || pat == 4, rnd_sfb = pat == 2 || pat == 4, rnd_code = pat >= 3;
      for (auto& b : hA) b = rnd_code ? (uint8_t)(rand() & 255) : 0x22;      // 0x2 = +1.0 in both nibbles
      for (auto& b : hB) b = rnd_code ? (uint8_t)(rand() & 255) : 0x22;
      auto rsf = [&](bool rnd) -> uint8_t { return rnd ? (uint8_t)(((5 + rand() % 5) << 3) | (rand() & 7)) : 0x38; };   // 2^-2..2^2 x mantissa | 1.0
      for (auto& b : hSFW) b = rsf(rnd_sfa);
      for (auto& b : hSFBrow) b = rsf(rnd_sfb);
      for (int k = 0; k < 2; ++k) for (int n = 0; n < 256; ++n) for (int kb = 0; kb < 4; ++kb)
        hSFBat[(size_t)(k * 2 + n / 128) * 512 + 16 * (n % 32) + 4 * ((n % 128) / 32) + kb] = hSFBrow[n * 8 + k * 4 + kb];
      cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
      cudaMemcpy(dSFW, hSFW.data(), hSFW.size(), cudaMemcpyHostToDevice); cudaMemcpy(dSFB, hSFBat.data(), hSFBat.size(), cudaMemcpyHostToDevice);
      cudaMemset(dOut, 0xff, 256 * 256 * 4);
      prm.shift = shift;
      probe_kernel<<<2, 192, smem>>>(prm);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("{\"pattern\": %d, \"shift\": %d, \"error\": \"%s\"}\n", pat, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hOut.data(), dOut, 256 * 256 * 4, cudaMemcpyDeviceToHost);
      double worst[2] = {0, 0}; int bad_m[2] = {-1, -1}, bad_n[2] = {-1, -1}; double bad_got[2] = {0, 0}, bad_exp[2] = {0, 0};
      for (int order = 0; order < 2; ++order)       // 0: element 2i in the low nibble; 1: in the high nibble
        for (int m = 0; m < 256; ++m) {
          const int cta = m / 128, wr = m % 128 + shift;
          const uint8_t* arow = &hA[(size_t)(cta * WIN_ROWS + wr) * ROW_B];
          const uint8_t* asf = &hSFW[(size_t)(cta * WIN_ROWS + wr) * 8];
          for (int n = 0; n < 256; ++n) {
            const uint8_t* brow = &hB[(size_t)n * ROW_B];
            double acc = 0;
            for (int kk = 0; kk < 128; ++kk) {
              const int sh = ((kk & 1) ^ order) ? 4 : 0;
              const float a = e2m1_val((arow[kk >> 1] >> sh) & 15) * ue4m3_val(asf[kk / 16]);
              const float b = e2m1_val((brow[kk >> 1] >> sh) & 15) * ue4m3_val(hSFBrow[n * 8 + kk / 16]);
              acc += (double)a * b;
            }
            const double got = hOut[(size_t)m * 256 + n], err = std::fabs(got - acc) / (1.0 + std::fabs(acc));
            if (!(err <= worst[order])) { worst[order] = std::isfinite(err) ? err : 1e30; bad_m[order] = m; bad_n[order] = n; bad_got[order] = got; bad_exp[order] = acc; }
          }
        }
      const int best = worst[0] <= worst[1] ? 0 : 1;
      const bool ok = worst[best] < 1e-5;
      if (!ok) ++fails;
      printf("{\"pattern\": %d, \"shift\": %d, \"ok\": %s, \"nibble_order\": \"%s\", \"rel_err\": %.3e, \"other_order_err\": %.3e, "
             "\"worst\": {\"m\": %d, \"n\": %d, \"got\": %.6g, \"expected\": %.6g}}\n",
             pat, shift, ok ? "true" : "false", best == 0 ? "even element in low nibble" : "even element in high nibble", worst[best],
             worst[1 - best], bad_m[best], bad_n[best], bad_got[best], bad_exp[best]);
    }
  printf("{\"summary\": \"%d of 25 cases failed\"}\n", fails);
  return fails ? 1 : 0;
}

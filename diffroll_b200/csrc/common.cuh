// Shared device helpers: error plumbing, bf16 hi/lo split, mbarrier / TMA / tcgen05 PTX wrappers.
// sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_fp4.h>
#include <stdint.h>
#include <stdio.h>

namespace drb {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define DRB_CUDA(expr)                                                               \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      drb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return (int)_e;                                                                \
    }                                                                                \
  } while (0)

#define DRB_LAUNCH_CHECK()                                                           \
  do {                                                                               \
    drb::count_launch();                                                             \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      drb::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return (int)_e;                                                                \
    }                                                                                \
  } while (0)

// ---------------------------------------------------------------------------------------------
// bf16 hi/lo split:  v ~= hi + lo with hi = bf16(v), lo = bf16(v - hi)   (|err| <= 2^-17 |v|)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// two floats -> packed hi pair and packed lo pair
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat16 ah, al, bh, bl;
  split_bf16(a, ah, al);
  split_bf16(b, bh, bl);
  hi = pack_bf16x2(ah, bh);
  lo = pack_bf16x2(al, bl);
}

// ---------------------------------------------------------------------------------------------
// fp16 + fp8 split ("f16f8"):  v ~= h + l,  h = fp16(v),  l = v - h  (|l| <= 2^-12 |v|).
//   main product   : h_a * h_b                                 (kind::f16, fp16 inputs)
//   correction     : [l_a*SA | e4m3(h_a)] . [e4m3(h_b*SW) | e4m3(l_b*SA*SW)]   (kind::f8f6f4, K concatenated)
// Both correction terms carry the common scale SA*SW, so they share one accumulator; the e4m3 rounding (2^-4) of a
// 2^-12-sized term leaves ~2^-16 relative error, like the bf16 hi/lo 3-product scheme, for 2 MMA-units instead of 3.
// ---------------------------------------------------------------------------------------------
constexpr float F8_SA = 4096.f;  // 2^12: scale of the activation residual before e4m3 rounding

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint16_t pack_e4m3x2(float a, float b) {
  return (uint16_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
}
// 4 consecutive values -> 2 packed fp16 pairs, 4 residual bytes (scaled by F8_SA), 4 e4m3 bytes of the fp16 value
__device__ __forceinline__ void split_f16f8_x4(const float (&v)[4], uint32_t& h01, uint32_t& h23, uint32_t& lo4, uint32_t& hi4) {
  const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  h01 = *reinterpret_cast<const uint32_t*>(&a);
  h23 = *reinterpret_cast<const uint32_t*>(&b);
  const float2 fa = __half22float2(a), fb = __half22float2(b);
  lo4 = (uint32_t)pack_e4m3x2((v[0] - fa.x) * F8_SA, (v[1] - fa.y) * F8_SA) |
        ((uint32_t)pack_e4m3x2((v[2] - fb.x) * F8_SA, (v[3] - fb.y) * F8_SA) << 16);
  hi4 = (uint32_t)pack_e4m3x2(fa.x, fa.y) | ((uint32_t)pack_e4m3x2(fb.x, fb.y) << 16);
}

// "f16e5": same idea with e5m2 corrections whose scales multiply to ONE, so the correction MMA accumulates straight into the
// main accumulator (no second accumulator -> TMEM has room for two accumulator stages):
//   [l_a*2^4 | e5m2(h_a*2^-8)] . [e5m2(h_b*2^-4) | e5m2(l_b*2^8)]     (e5m2's 5 exponent bits give the range, its 2 mantissa
//   bits leave ~2^-15 relative error: 200-step chain error ~2e-4 instead of ~1e-4, still 5x inside the 1e-3 bar)
__device__ __forceinline__ uint16_t pack_e5m2x2(float a, float b) {
  return (uint16_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E5M2);
}
__device__ __forceinline__ void split_f16e5_x4(const float (&v)[4], uint32_t& h01, uint32_t& h23, uint32_t& lo4, uint32_t& hi4) {
  const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  h01 = *reinterpret_cast<const uint32_t*>(&a);
  h23 = *reinterpret_cast<const uint32_t*>(&b);
  const float2 fa = __half22float2(a), fb = __half22float2(b);
  lo4 = (uint32_t)pack_e5m2x2((v[0] - fa.x) * 16.f, (v[1] - fa.y) * 16.f) |
        ((uint32_t)pack_e5m2x2((v[2] - fb.x) * 16.f, (v[3] - fb.y) * 16.f) << 16);
  hi4 = (uint32_t)pack_e5m2x2(fa.x * 0.00390625f, fa.y * 0.00390625f) |
        ((uint32_t)pack_e5m2x2(fb.x * 0.00390625f, fb.y * 0.00390625f) << 16);
}

// the B-side ("weight") form of the same split: [e5m2(h * 2^-4) | e5m2(l * 2^8)], pairing with the activation form above
__device__ __forceinline__ void split_f16e5w_x4(const float (&v)[4], uint32_t& h01, uint32_t& h23, uint32_t& first4, uint32_t& second4) {
  const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  h01 = *reinterpret_cast<const uint32_t*>(&a);
  h23 = *reinterpret_cast<const uint32_t*>(&b);
  const float2 fa = __half22float2(a), fb = __half22float2(b);
  first4 = (uint32_t)pack_e5m2x2(fa.x * 0.0625f, fa.y * 0.0625f) | ((uint32_t)pack_e5m2x2(fb.x * 0.0625f, fb.y * 0.0625f) << 16);
  second4 = (uint32_t)pack_e5m2x2((v[0] - fa.x) * 256.f, (v[1] - fa.y) * 256.f) |
            ((uint32_t)pack_e5m2x2((v[2] - fb.x) * 256.f, (v[3] - fb.y) * 256.f) << 16);
}

// ---------------------------------------------------------------------------------------------
// "f16n4": fp16 main product + ONE block-scaled fp4 correction product (kind::mxf4nvf4, K = 64 per instruction, half the
// issue cycles of the e5m2 correction).  Per 64-channel chunk an activation row carries 64 B of e2m1 codes
// [lo part (64 codes) | hi part (64 codes)] and 8 ue4m3 scales (one per 16 codes); weights carry [hi | lo], so
//   instruction 0:  (l_a * 2^10) . (h_w * 2^-10)      instruction 1:  (h_a * 2^4) . (l_w * 2^-4)
// with h = fp16(v), l = v - h and weights pre-multiplied by the per-tensor power of two SW (max |W * SW| in (2^14, 2^15]).
// The powers of two put every block scale (block max / 6) inside ue4m3's range [2^-9, 448] for |activation| in
// [2^-6, 2^8]; the scale is rounded UP to the next ue4m3 so the block maximum never saturates.  Emulated on real layer
// inputs (profiles/experiments/n4_emulation.py) the per-layer rms error is 4.3e-5 (f16e5: 2.1e-5, one fp16 product: 2.9e-4).
// ---------------------------------------------------------------------------------------------
constexpr float N4_ALO = 1024.f, N4_AHI = 16.f;              // activation lo / hi part multipliers
constexpr float N4_WHI = 1.f / 1024.f, N4_WLO = 1.f / 16.f;  // weight hi / lo part multipliers
constexpr float N4_WTARGET = 32768.f;                         // SW = 2^floor(log2(N4_WTARGET / max|W|))

__device__ __forceinline__ float e4m3_to_float(uint32_t b) {
  return __half2float(__half(__nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)b, __NV_E4M3)));
}
// 16 values -> 8 bytes of e2m1 codes (value 2i in the low nibble of byte i) and one ue4m3 scale byte >= max|v| / 6
// The block is v[i] * mul with mul a power of two (the part multipliers N4_ALO / N4_AHI): folded into the block maximum and into the
// reciprocal scale instead of 16 multiplies -- both scalings are exact, so the codes are bit-identical to scaling the values first.
__device__ __forceinline__ void n4_block16(const float (&v)[16], uint32_t& c0, uint32_t& c1, uint32_t& sf, float mul = 1.f) {
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) amax = fmaxf(amax, fabsf(v[i]));
  const float want = (amax * mul) * (1.f / 6.f);
  uint32_t b = (uint32_t)__nv_cvt_float_to_fp8(want, __NV_SATFINITE, __NV_E4M3);
  float q = e4m3_to_float(b);
  if (q < want && b < 0x7eu) { ++b; q = e4m3_to_float(b); }
  const float inv = (q > 0.f ? __frcp_rn(q) : 0.f) * mul;
  uint32_t w[2] = {0u, 0u};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t byte = (uint32_t)__nv_cvt_float2_to_fp4x2(make_float2(v[2 * i] * inv, v[2 * i + 1] * inv), __NV_E2M1, cudaRoundNearest);
    w[i >> 2] |= (byte & 0xffu) << (8 * (i & 3));
  }
  c0 = w[0]; c1 = w[1]; sf = b;
}
// 16 fp32 values -> 8 packed fp16 pairs + the two f16n4 blocks (lo part, hi part) of an ACTIVATION row
__device__ __forceinline__ void split_f16n4_x16(const float (&v)[16], uint32_t (&h)[8], uint32_t (&lo)[2], uint32_t (&hi)[2],
                                                uint32_t& sf_lo, uint32_t& sf_hi) {
  float l[16], g[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __half2 a = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&a);
    const float2 f = __half22float2(a);
    l[2 * i] = v[2 * i] - f.x; l[2 * i + 1] = v[2 * i + 1] - f.y;
    g[2 * i] = f.x; g[2 * i + 1] = f.y;
  }
  n4_block16(l, lo[0], lo[1], sf_lo, N4_ALO);
  n4_block16(g, hi[0], hi[1], sf_hi, N4_AHI);
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must trap (visible CUDA error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("drb: mbarrier timeout block %d thread %d\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

// TMA tile loads (global -> shared, completes on an mbarrier).  Coordinates are signed; elements
// outside the tensor (including negative coordinates) are zero-filled.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster cooperate on one M=256 MMA --------------------------
// Map a shared::cta address of THIS CTA to the shared::cluster address of the same offset in CTA `rank`.
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// arrive + expect_tx on a (possibly remote) barrier given by its shared::cluster address
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster_addr), "r"(bytes) : "memory");
}
// plain arrive on a (possibly remote) barrier given by its shared::cluster address
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// TMA loads issued by either CTA of a pair: data lands in the issuing CTA's smem, completion bytes go to the barrier
// at `bar_cluster_addr` (the leader CTA's full barrier).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// Programmatic dependent launch: `pdl_trigger` lets the next kernel of the stream start its prologue while this grid is
// still running; `pdl_wait` blocks until the previous grid has completed and its memory is visible.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA tile store (shared -> global, bulk async-group completion).  Rows/columns outside the tensor are clipped.
__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"((uint64_t)tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {  // at most N groups still reading their smem source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  __syncwarp();  // bar.sync is warp-aligned: reconverge first
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Byte offset of (row, 16-byte chunk) inside a 128-byte-swizzled tile whose rows are 128 bytes (base 1024-aligned).
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk16) {
  return (uint32_t)row * 128u + (uint32_t)((chunk16 ^ (row & 7)) << 4);
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tmap) : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; kind::f16 (bf16 inputs, fp32 accumulate); one thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f8f6f4 (e4m3 inputs here, fp32 accumulate), K = 32 per instruction.
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// CTA-pair forms (issued by the leader CTA only).  M = 256: each CTA supplies 128 rows of A and half the rows of B from
// the same smem offsets, and receives its 128 accumulator rows in its own TMEM.
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (when all prior MMAs of the pair retire) on the barrier at this smem offset in both CTAs.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)0x3)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {  // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive columns (one row per thread).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory matrix descriptor (rows of 64 bf16 = 128 B, 8-row
// groups 1024 B apart).  Fields: start>>4 [0,14), LBO>>4 [16,30) (ignored for swizzled K-major),
// SBO>>4 [32,46), version=1 [46,48), base_offset [49,52) = 0, layout SWIZZLE_128B=2 [61,64).
// The start may be ANY 16-byte-aligned address inside a swizzled tile (e.g. a view shifted by r rows = r*128 B into a
// TMA-written window): measured on B200, the swizzle XOR is taken from the absolute shared-memory address bits, so a
// row-shifted view needs base_offset = 0; setting base_offset = (addr >> 7) & 7 produces garbage.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// K-major, 64-byte-swizzled operand (rows of 64 B = 128 e2m1 codes, 8-row groups 512 B apart).  Row-shifted views work like
// under SWIZZLE_128B (verified on a B200: profiles/experiments/nv4_probe.cu).
__device__ __forceinline__ uint64_t make_sw64_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// Un-swizzled 32 rows x 16 B (one 512-byte scale-factor atom: rows 16 B apart, 8-row groups 128 B apart), source of tcgen05.cp
__device__ __forceinline__ uint64_t make_sfatom_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(16 >> 4) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// smem -> TMEM copy of one scale-factor atom in BOTH CTAs of a pair (each from its own smem at this offset): 32 lanes x 4
// columns, replicated over the four lane quarters
__device__ __forceinline__ void utccp_sf_pair(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// block-scaled fp4 MMA of a CTA pair: e2m1 codes, one ue4m3 scale per 16 codes (scale_vec::4X), K = 64
__device__ __forceinline__ void umma_nv4_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                                              uint32_t tmem_sfa, uint32_t tmem_sfb) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::mxf4nvf4.block_scale.scale_vec::4X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}
// instruction descriptor of the block-scaled fp4 MMA: a/b format 1 = e2m1 [7,10) / [10,13), scale format 0 = ue4m3 [23],
// N>>3 [17,23), M>>4 [24,29), both K-major
__host__ __device__ constexpr uint32_t make_idesc_nv4(int M, int N) {
  return (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
// byte offset of (row, 16-byte chunk) inside a 64-byte-swizzled tile whose rows are 64 bytes (base 512-aligned)
__device__ __forceinline__ uint32_t sw64_off(int row, int chunk16) {
  return (uint32_t)row * 64u + (uint32_t)((chunk16 ^ ((row >> 1) & 3)) << 4);
}

// Instruction descriptor, kind::f16: c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1, both K-major,
// N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// fp16 inputs (kind::f16, a/b format 0) and e4m3 inputs (kind::f8f6f4, a/b format 0): same bit pattern, fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_fmt0(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace drb

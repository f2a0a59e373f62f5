// tcgen05 (5th-gen tensor core) kernels for the contractions of the ResidualBlock stack
// (model/diffwave.py:134-151, 680-684), sm_100a only.
//
//   gate kernel  : y[t, n] = sum_{tap, c} xin[t + (tap-k/2)*dil, c] * Wd[n, tap, c]  (+ spec[t, :] . Wc[n, :])  + bias1[n]
//                  z_l[t, c] = sigmoid(y[t, c]) * tanh(y[t, C + c])                     -> bf16 hi/lo, kept for every layer l
//   zgemm RES    : o[t, c] = sum_k z_l[t, k] * Wo_l[c, k] + bo_l[c]            (residual half of output_projection)
//                  x[t, c] = (x[t, c] + o[t, c]) / sqrt(2)    -> fp32, and bf16 hi/lo of (x + d_{l+1})
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
// fp32 parity (|delta| < 1e-3 after 200 chained steps) needs more than one bf16 product, so
// activations and weights are kept as bf16 hi + lo pairs and every K-step issues three MMAs into the same TMEM
// accumulator: hi*hi + lo*hi + hi*lo.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue
// (TMEM -> registers -> swizzled smem staging -> TMA store).  smem ring: full/empty mbarriers per stage; after the last
// MMA retires the ring memory is reused as epilogue staging.
#include "common.cuh"
#include "kernels.h"

namespace drb {

constexpr int TILE_M = 128;
constexpr int TILE_N = 256;
constexpr int TILE_K = 64;
constexpr int A_TILE_BYTES = TILE_M * TILE_K * 2;  // 16 KB
constexpr int B_TILE_BYTES = TILE_N * TILE_K * 2;  // 32 KB
constexpr int UMMA_K = 16;
constexpr uint32_t TMEM_COLS = 256;
constexpr int CHUNK_BYTES = TILE_M * 128;          // one 128-row x 128-byte swizzled staging box (16 KB)
constexpr int EPI_BAR = 1;                         // named barrier of the 4 epilogue warps

template <bool THREE>
struct Cfg {
  static constexpr int kStageBytes = THREE ? 2 * (A_TILE_BYTES + B_TILE_BYTES) : (A_TILE_BYTES + B_TILE_BYTES);
  static constexpr int kStages = THREE ? 2 : 4;
  static constexpr int kRingBytes = kStages * kStageBytes;  // 192 KB
  static constexpr int kSmemBytes = kRingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct alignas(64) GateParams {
  CUtensorMap xh, xl, sh, sl, wd_h, wd_l, wc_h, wc_l, zh, zl;
  int NB, n_cond, T, C, taps, dil, cond_slabs, tiles_t, n_blocks, z_group0;
  const float* bias_cond;
  const float* bias_unc;
};

struct alignas(64) ZGemmParams {
  CUtensorMap zh, zl, w_h, w_l, out32, xh, xl;
  int NB, T, C, tiles_t, n_blocks, nslabs, spg, z_group0, group_stride, mode;  // mode 0 = RES, 1 = HEAD
  const float* bias;
  const float* dnext;
};

struct SmemView {
  uint8_t* stage0;
  uint64_t* full;
  uint64_t* empty;
  uint64_t* tmem_full;
  uint64_t* xin_full;
  uint32_t* tmem_ptr;
};

template <bool THREE>
__device__ __forceinline__ SmemView carve(uint8_t* raw) {
  uint32_t a = smem_u32(raw);
  uint32_t pad = ((a + 1023u) & ~1023u) - a;
  SmemView v;
  v.stage0 = raw + pad;
  uint8_t* bars = v.stage0 + Cfg<THREE>::kRingBytes;
  v.full = reinterpret_cast<uint64_t*>(bars);
  v.empty = v.full + Cfg<THREE>::kStages;
  v.tmem_full = v.empty + Cfg<THREE>::kStages;
  v.xin_full = v.tmem_full + 1;
  v.tmem_ptr = reinterpret_cast<uint32_t*>(v.xin_full + 1);
  return v;
}

template <bool THREE>
__device__ __forceinline__ void prologue(const SmemView& sv, int warp) {
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < Cfg<THREE>::kStages; ++i) { mbar_init(&sv.full[i], 1); mbar_init(&sv.empty[i], 1); }
    mbar_init(sv.tmem_full, 1);
    mbar_init(sv.xin_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(sv.tmem_ptr, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

// MMA issue for one K-slab (64 channels = 4 UMMA K-steps) resident in stage memory.
template <bool THREE>
__device__ __forceinline__ void issue_slab(uint8_t* st, uint32_t tmem_d, bool first_slab) {
  const uint32_t a_hi = smem_u32(st);
  const uint32_t a_lo = a_hi + A_TILE_BYTES;
  const uint32_t b_hi = a_hi + (THREE ? 2 * A_TILE_BYTES : A_TILE_BYTES);
  const uint32_t b_lo = b_hi + B_TILE_BYTES;
  constexpr uint32_t idesc = make_idesc_bf16(TILE_M, TILE_N);
#pragma unroll
  for (int k = 0; k < TILE_K / UMMA_K; ++k) {
    const uint32_t ko = k * UMMA_K * 2;  // byte advance inside the 128-byte swizzled row
    const uint64_t da_hi = make_sw128_desc(a_hi + ko), db_hi = make_sw128_desc(b_hi + ko);
    umma_bf16(tmem_d, da_hi, db_hi, idesc, (first_slab && k == 0) ? 0u : 1u);
    if (THREE) {
      umma_bf16(tmem_d, make_sw128_desc(a_lo + ko), db_hi, idesc, 1u);
      umma_bf16(tmem_d, da_hi, make_sw128_desc(b_lo + ko), idesc, 1u);
    }
  }
}

template <bool THREE>
__device__ __forceinline__ void mma_loop(const SmemView& sv, uint32_t tmem_base, int nslabs) {
  int stage = 0; uint32_t phase = 0;
  for (int s = 0; s < nslabs; ++s) {
    mbar_wait(&sv.full[stage], phase);
    tc_fence_after();
    issue_slab<THREE>(sv.stage0 + stage * Cfg<THREE>::kStageBytes, tmem_base, s == 0);
    umma_commit(&sv.empty[stage]);  // frees the smem slot when these MMAs retire
    if (++stage == Cfg<THREE>::kStages) { stage = 0; phase ^= 1; }
  }
  umma_commit(sv.tmem_full);
}

// sigmoid(g) * tanh(f) with the SFU exp2/rcp approximations (abs error ~2e-7, far below the bf16 hi/lo split error)
__device__ __forceinline__ float gate_act(float g, float f) {
  const float sg = __fdividef(1.f, 1.f + __expf(-g));
  const float th = 1.f - __fdividef(2.f, __expf(2.f * f) + 1.f);
  return sg * th;
}

// ---------------------------------------------------------------------------------------------
// gate kernel
// ---------------------------------------------------------------------------------------------
template <bool THREE>
__global__ void __launch_bounds__(256, 1) umma_gate_kernel(const __grid_constant__ GateParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmemView sv = carve<THREE>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int bid = blockIdx.x;
  const int nblk = bid % p.n_blocks; bid /= p.n_blocks;
  const int tt = bid % p.tiles_t;
  const int nb = bid / p.tiles_t;
  const int t0 = tt * TILE_M;
  const int cpt = p.C / TILE_K;  // K-slabs per tap
  const int conv_slabs = p.taps * cpt;
  const int nslabs = conv_slabs + (nb < p.n_cond ? p.cond_slabs : 0);
  const int half = p.taps / 2;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.xh); tma_prefetch_desc(&p.wd_h); tma_prefetch_desc(&p.zh);
    if (THREE) { tma_prefetch_desc(&p.xl); tma_prefetch_desc(&p.wd_l); tma_prefetch_desc(&p.zl); }
  }
  prologue<THREE>(sv, warp);
  const uint32_t tmem_base = *sv.tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int s = 0; s < nslabs; ++s) {
        mbar_wait(&sv.empty[stage], phase ^ 1);
        uint8_t* st = sv.stage0 + stage * Cfg<THREE>::kStageBytes;
        uint8_t* a_hi = st; uint8_t* a_lo = st + A_TILE_BYTES;
        uint8_t* b_hi = st + (THREE ? 2 * A_TILE_BYTES : A_TILE_BYTES); uint8_t* b_lo = b_hi + B_TILE_BYTES;
        mbar_expect_tx(&sv.full[stage], Cfg<THREE>::kStageBytes);
        if (s < conv_slabs) {
          const int tap = s / cpt, cc = s - tap * cpt;
          const int trow = t0 + (tap - half) * p.dil;
          tma_load_3d(a_hi, &p.xh, &sv.full[stage], cc * TILE_K, trow, nb);
          tma_load_2d(b_hi, &p.wd_h, &sv.full[stage], tap * p.C + cc * TILE_K, nblk * TILE_N);
          if (THREE) {
            tma_load_3d(a_lo, &p.xl, &sv.full[stage], cc * TILE_K, trow, nb);
            tma_load_2d(b_lo, &p.wd_l, &sv.full[stage], tap * p.C + cc * TILE_K, nblk * TILE_N);
          }
        } else {
          const int cc = s - conv_slabs;
          tma_load_3d(a_hi, &p.sh, &sv.full[stage], cc * TILE_K, t0, nb);
          tma_load_2d(b_hi, &p.wc_h, &sv.full[stage], cc * TILE_K, nblk * TILE_N);
          if (THREE) {
            tma_load_3d(a_lo, &p.sl, &sv.full[stage], cc * TILE_K, t0, nb);
            tma_load_2d(b_lo, &p.wc_l, &sv.full[stage], cc * TILE_K, nblk * TILE_N);
          }
        }
        if (++stage == Cfg<THREE>::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) mma_loop<THREE>(sv, tmem_base, nslabs);
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const float* bias = (nb < p.n_cond ? p.bias_cond : p.bias_unc) + nblk * TILE_N;
    mbar_wait(sv.tmem_full, 0);  // every MMA has retired: accumulator complete, ring memory free
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    // staging: z hi boxes 0,1 then z lo boxes 0,1; each [128 frames][64 channels] bf16, 128-byte swizzled
    const uint32_t stg = smem_u32(sv.stage0);
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t g[32], f[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: the whole warp must arrive together
      tmem_ld32(taddr + ch * 32, g);
      tmem_ld32(taddr + 128 + ch * 32, f);
      tmem_ld_wait();
      const uint32_t box_h = stg + (ch >> 1) * CHUNK_BYTES, box_l = box_h + 2 * CHUNK_BYTES;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = v * 8 + e * 2;
          const float z0 = gate_act(__uint_as_float(g[i]) + __ldg(bias + ch * 32 + i),
                                    __uint_as_float(f[i]) + __ldg(bias + 128 + ch * 32 + i));
          const float z1 = gate_act(__uint_as_float(g[i + 1]) + __ldg(bias + ch * 32 + i + 1),
                                    __uint_as_float(f[i + 1]) + __ldg(bias + 128 + ch * 32 + i + 1));
          split_pack2(z0, z1, hi[e], lo[e]);
        }
        const uint32_t off = sw128_off(row, (ch & 1) * 4 + v);
        sts128u(box_h + off, hi[0], hi[1], hi[2], hi[3]);
        sts128u(box_l + off, lo[0], lo[1], lo[2], lo[3]);
      }
    }
    tc_fence_before();
    fence_proxy_async();  // make the generic-proxy smem writes visible to the TMA (async proxy)
    named_bar_sync(EPI_BAR, 128);
    if (warp == 4 && elect_one()) {
      const int c0 = nblk * (TILE_N / 2);
      tma_store_3d(&p.zh, sv.stage0, c0, t0, p.z_group0 + nb);
      tma_store_3d(&p.zh, sv.stage0 + CHUNK_BYTES, c0 + 64, t0, p.z_group0 + nb);
      if (THREE) {
        tma_store_3d(&p.zl, sv.stage0 + 2 * CHUNK_BYTES, c0, t0, p.z_group0 + nb);
        tma_store_3d(&p.zl, sv.stage0 + 3 * CHUNK_BYTES, c0 + 64, t0, p.z_group0 + nb);
      }
      tma_store_commit();
      tma_store_wait_read<0>();  // smem must stay valid until the bulk stores have read it
    }
  }
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------
// zgemm kernel: A = stored z of one layer (RES) or of all layers (HEAD)
// ---------------------------------------------------------------------------------------------
template <bool THREE>
__global__ void __launch_bounds__(256, 1) umma_zgemm_kernel(const __grid_constant__ ZGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmemView sv = carve<THREE>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int bid = blockIdx.x;
  const int nblk = bid % p.n_blocks; bid /= p.n_blocks;
  const int tt = bid % p.tiles_t;
  const int nb = bid / p.tiles_t;
  const int t0 = tt * TILE_M;
  const int n_base = nblk * TILE_N;
  const bool res = p.mode == 0;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.zh); tma_prefetch_desc(&p.w_h); tma_prefetch_desc(&p.out32);
    if (THREE) { tma_prefetch_desc(&p.zl); tma_prefetch_desc(&p.w_l); }
  }
  prologue<THREE>(sv, warp);
  const uint32_t tmem_base = *sv.tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int s = 0; s < p.nslabs; ++s) {
        mbar_wait(&sv.empty[stage], phase ^ 1);
        uint8_t* st = sv.stage0 + stage * Cfg<THREE>::kStageBytes;
        uint8_t* a_hi = st; uint8_t* a_lo = st + A_TILE_BYTES;
        uint8_t* b_hi = st + (THREE ? 2 * A_TILE_BYTES : A_TILE_BYTES); uint8_t* b_lo = b_hi + B_TILE_BYTES;
        const int grp = s / p.spg, cc = s - grp * p.spg;
        const int zrow = p.z_group0 + grp * p.group_stride + nb;
        mbar_expect_tx(&sv.full[stage], Cfg<THREE>::kStageBytes);
        tma_load_3d(a_hi, &p.zh, &sv.full[stage], cc * TILE_K, t0, zrow);
        tma_load_2d(b_hi, &p.w_h, &sv.full[stage], s * TILE_K, n_base);
        if (THREE) {
          tma_load_3d(a_lo, &p.zl, &sv.full[stage], cc * TILE_K, t0, zrow);
          tma_load_2d(b_lo, &p.w_l, &sv.full[stage], s * TILE_K, n_base);
        }
        if (++stage == Cfg<THREE>::kStages) { stage = 0; phase ^= 1; }
      }
      if (res) {
        // Residual update needs the fp32 x tile: fetch it into the (now free) ring as 8 swizzled [128][32] fp32 boxes.
        mbar_wait(sv.tmem_full, 0);
        mbar_expect_tx(sv.xin_full, 8 * CHUNK_BYTES);
        for (int c = 0; c < 8; ++c) tma_load_3d(sv.stage0 + c * CHUNK_BYTES, &p.out32, sv.xin_full, n_base + c * 32, t0, nb);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) mma_loop<THREE>(sv, tmem_base, p.nslabs);
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const float* bias = p.bias + n_base;
    const bool issuer = (warp == 4) && (lane == 0);
    mbar_wait(sv.tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t stg = smem_u32(sv.stage0);
    const float rs2 = 1.41421356237309515f;
    if (res) mbar_wait(sv.xin_full, 0);
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {       // 64 output channels per iteration
      uint32_t o[64];
      __syncwarp();
      tmem_ld32(taddr + it * 64, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
      tmem_ld32(taddr + it * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
      tmem_ld_wait();
      if (res) {
        // bf16 staging (x + d_next split): two alternating sets of {hi box, lo box} behind the 8 fp32 boxes
        const uint32_t set = stg + 8 * CHUNK_BYTES + (it & 1) * 2 * CHUNK_BYTES;
        if (it >= 2) {                     // the TMA stores of iteration it-2 must have finished reading this set
          if (issuer) tma_store_wait_read<1>();
          named_bar_sync(EPI_BAR, 128);
        }
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {   // fp32 box (32 channels) inside this iteration
          const uint32_t xbox = stg + (it * 2 + hc) * CHUNK_BYTES;
          const float* dn = p.dnext + n_base + it * 64 + hc * 32;
#pragma unroll
          for (int v = 0; v < 8; v += 2) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int i = hc * 32 + (v + u) * 4;
              const uint32_t xa = xbox + sw128_off(row, v + u);
              float4 x = lds128(xa);
              const float4 d = __ldg(reinterpret_cast<const float4*>(dn + (v + u) * 4));
              x.x = (x.x + (__uint_as_float(o[i + 0]) + __ldg(bias + it * 64 + i + 0))) / rs2;
              x.y = (x.y + (__uint_as_float(o[i + 1]) + __ldg(bias + it * 64 + i + 1))) / rs2;
              x.z = (x.z + (__uint_as_float(o[i + 2]) + __ldg(bias + it * 64 + i + 2))) / rs2;
              x.w = (x.w + (__uint_as_float(o[i + 3]) + __ldg(bias + it * 64 + i + 3))) / rs2;
              sts128(xa, x);  // in place: same thread, same address
              split_pack2(x.x + d.x, x.y + d.y, hi[u * 2 + 0], lo[u * 2 + 0]);
              split_pack2(x.z + d.z, x.w + d.w, hi[u * 2 + 1], lo[u * 2 + 1]);
            }
            const uint32_t off = sw128_off(row, hc * 4 + (v >> 1));
            sts128u(set + off, hi[0], hi[1], hi[2], hi[3]);
            sts128u(set + CHUNK_BYTES + off, lo[0], lo[1], lo[2], lo[3]);
          }
        }
        fence_proxy_async();
        named_bar_sync(EPI_BAR, 128);
        if (issuer) {
          const int c0 = n_base + it * 64;
          tma_store_3d(&p.out32, sv.stage0 + (it * 2) * CHUNK_BYTES, c0, t0, nb);
          tma_store_3d(&p.out32, sv.stage0 + (it * 2 + 1) * CHUNK_BYTES, c0 + 32, t0, nb);
          tma_store_3d(&p.xh, sv.stage0 + 8 * CHUNK_BYTES + (it & 1) * 2 * CHUNK_BYTES, c0, t0, nb);
          if (THREE) tma_store_3d(&p.xl, sv.stage0 + 8 * CHUNK_BYTES + (it & 1) * 2 * CHUNK_BYTES + CHUNK_BYTES, c0, t0, nb);
          tma_store_commit();
        }
      } else {
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
          const uint32_t hbox = stg + (it * 2 + hc) * CHUNK_BYTES;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const int i = hc * 32 + v * 4;
            float4 h;
            h.x = fmaxf(__uint_as_float(o[i + 0]) + __ldg(bias + it * 64 + i + 0), 0.f);
            h.y = fmaxf(__uint_as_float(o[i + 1]) + __ldg(bias + it * 64 + i + 1), 0.f);
            h.z = fmaxf(__uint_as_float(o[i + 2]) + __ldg(bias + it * 64 + i + 2), 0.f);
            h.w = fmaxf(__uint_as_float(o[i + 3]) + __ldg(bias + it * 64 + i + 3), 0.f);
            sts128(hbox + sw128_off(row, v), h);
          }
        }
        fence_proxy_async();
        named_bar_sync(EPI_BAR, 128);
        if (issuer) {
          const int c0 = n_base + it * 64;
          tma_store_3d(&p.out32, sv.stage0 + (it * 2) * CHUNK_BYTES, c0, t0, nb);
          tma_store_3d(&p.out32, sv.stage0 + (it * 2 + 1) * CHUNK_BYTES, c0 + 32, t0, nb);
          tma_store_commit();
        }
      }
    }
    tc_fence_before();
    if (issuer) tma_store_wait_read<0>();
  }
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
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
  cudaError_t e1, e2, e3, e4;
  e1 = cudaFuncSetAttribute(umma_gate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<true>::kSmemBytes);
  e2 = cudaFuncSetAttribute(umma_gate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<false>::kSmemBytes);
  e3 = cudaFuncSetAttribute(umma_zgemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<true>::kSmemBytes);
  e4 = cudaFuncSetAttribute(umma_zgemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<false>::kSmemBytes);
  if (e1 || e2 || e3 || e4) {
    set_error("cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(e1 ? e1 : e2 ? e2 : e3 ? e3 : e4));
    g_encode = nullptr;
    return (int)(e1 ? e1 : e2 ? e2 : e3 ? e3 : e4);
  }
  return 0;
}

static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, bool f32) {
  if (!g_encode) { int r = umma_init(); if (r) return r; }
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                        const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return DRB_E_DRIVER; }
  return 0;
}

int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  return encode(m, base, 2, dims, strides, box, false);
}

int make_tmap_3d(CUtensorMap* m, const void* base, uint64_t d2, uint64_t d1, uint64_t d0, uint32_t box1, uint32_t box0) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 2, d1 * d0 * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  return encode(m, base, 3, dims, strides, box, false);
}

int make_tmap_3d_f32(CUtensorMap* m, const void* base, uint64_t d2, uint64_t d1, uint64_t d0, uint32_t box1, uint32_t box0) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 4, d1 * d0 * 4};
  cuuint32_t box[3] = {box0, box1, 1};
  return encode(m, base, 3, dims, strides, box, true);
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
  const int grid = p.NB * p.tiles_t * p.n_blocks;
  if (g.three) umma_gate_kernel<true><<<grid, 256, Cfg<true>::kSmemBytes, s>>>(p);
  else umma_gate_kernel<false><<<grid, 256, Cfg<false>::kSmemBytes, s>>>(p);
  DRB_LAUNCH_CHECK();
  return 0;
}

int launch_umma_zgemm(const UmmaMaps& maps, const UmmaZGemm& z, cudaStream_t s) {
  if (z.C % TILE_N) { set_error("umma_zgemm: unsupported C=%d", z.C); return DRB_E_INVALID; }
  ZGemmParams p;
  p.zh = maps.zh; p.zl = maps.zl; p.w_h = *z.w_h; p.w_l = *z.w_l; p.out32 = *z.out32; p.xh = maps.xh; p.xl = maps.xl;
  p.NB = z.NB; p.T = z.T; p.C = z.C; p.tiles_t = (z.T + TILE_M - 1) / TILE_M; p.n_blocks = z.C / TILE_N;
  p.spg = z.C / TILE_K; p.nslabs = z.groups * p.spg; p.z_group0 = z.z_group0; p.group_stride = z.group_stride;
  p.mode = z.mode; p.bias = z.bias; p.dnext = z.dnext;
  const int grid = p.NB * p.tiles_t * p.n_blocks;
  if (z.three) umma_zgemm_kernel<true><<<grid, 256, Cfg<true>::kSmemBytes, s>>>(p);
  else umma_zgemm_kernel<false><<<grid, 256, Cfg<false>::kSmemBytes, s>>>(p);
  DRB_LAUNCH_CHECK();
  return 0;
}

}  // namespace drb

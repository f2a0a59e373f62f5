// C ABI of libdiffroll_b200.so: plan construction (weight repack, TMA descriptors, cuFFT plan) and the
// per-step entry points.  See include/diffroll_b200.h for the contract and the reference lines replaced.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <nvtx3/nvToolsExt.h>
#include "common.cuh"
#include "kernels.h"

namespace drb {

// Opt-in NVTX ranges (DRB_NVTX=1): one range per sampler step and per residual layer, so a timeline tool shows the
// loop's structure.  The header-only NVTX v3 costs nothing when no tool is attached; off by default anyway.
static int g_nvtx = -1;
struct NvtxRange {
  bool on;
  NvtxRange(const char* fmt, int a, int b = 0) {
    if (g_nvtx < 0) { const char* e = getenv("DRB_NVTX"); g_nvtx = (e && e[0] == '1') ? 1 : 0; }
    on = g_nvtx == 1;
    if (on) { char buf[64]; snprintf(buf, sizeof(buf), fmt, a, b); nvtxRangePushA(buf); }
  }
  ~NvtxRange() { if (on) nvtxRangePop(); }
};

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }

static size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct Layout {  // byte offsets into the workspace
  size_t dtab, emb1, emb2, x32, skip, hbuf, spec32, bias, wtmp, mel, xpp;
  size_t wd32, wc32, ybuf, z32;                        // fp32 path
  size_t xh, xl, zh, zl, sh, sl, wdh, wdl, wch, wcl, woh, wol, wcomp32, wcomph, wcompl, bcomp, bsum, wscale;  // tensor path
  size_t wouth, woutl;                                  // output_projection rows padded to 256, operand pair (tensor-core head)
  size_t dtab2, iota;                                   // per-roll time tables of fractional diffusion steps (drb_plan_set_step_embeddings)
  size_t sp5h, sp5l, wc5h, wc5l;                        // f16x3 (fp16 hi + fp16 lo) pairs of the spectrogram and the conditioner weights (tables)
  size_t xs, wsf4;                                      // f16n4: activation scale factors [NB][C/64][T][8]; weight scale atoms [L][2C/256][k*C/64][2048]
  size_t range;                                         // one word: max |activation operand| (fp32 bits) since the last reset
  size_t wcpad, cond;                                   // fp32 Wc of every layer padded to Mp; conditioner projections of the spectrogram [L][B][T][2C]
  size_t total;
  int NBcap, Mp, KC;
  int Bs;   // clips held by spec32 / the conditioner tables: batch, or 2 x batch with the learned unconditional spectrogram
};

static bool cfg_ok(const drb_config& c) {
  if (c.batch <= 0 || c.frames <= 0 || c.pitches <= 0 || c.pitches % 4 || c.wave_len <= 0) return false;
  if (c.residual_channels <= 0 || c.residual_channels % 256) return false;
  if (c.residual_layers <= 0 || c.kernel_size <= 0 || !(c.kernel_size & 1)) return false;
  if (c.dilation_base <= 0 || c.dilation_bound <= 0 || c.n_mels <= 0 || c.n_fft <= 0 || c.hop_length <= 0) return false;
  if (c.timesteps <= 0 || c.precision < 0 || c.precision > 5 || c.branches < 0 || c.branches > 5) return false;
  if (c.wave_len / c.hop_length + 1 < c.frames) return false;
  if (c.wave_len <= c.n_fft / 2) return false;  // reflect padding needs pad < length
  return true;
}

// f16n4 exists only as persistent CTA-pair kernels: a plan whose M-tile count can be odd (batch * ceil(frames / 128) odd) or
// that was told not to use pairs / windows / persistence runs as f16e5 (same parity grade, 2 MMA units).
static drb_config effective_config(const drb_config& in) {
  drb_config c = in;
  if (c.precision == DRB_PREC_F16N4) {
    auto off = [](const char* name) { const char* e = getenv(name); return e && e[0] == '1'; };
    const int tiles_t = (c.frames + 127) / 128;
    const bool ok = ((c.batch * tiles_t) % 2 == 0) && !off("DRB_NO_PAIR") && !off("DRB_NO_WINDOW") && !off("DRB_NO_PERSIST") &&
                    !off("DRB_NO_CONDPRE") && !off("DRB_NO_N4");
    int win = 0, d = 1;
    for (int i = 0; i < c.dilation_bound && i < c.residual_layers; ++i) { win = 128 + (c.kernel_size - 1) * d; d *= c.dilation_base; }
    if (!ok || win > 192) c.precision = DRB_PREC_F16E5;
  }
  return c;
}

static Layout make_layout(const drb_config& c) {
  Layout l; memset(&l, 0, sizeof(l));
  const size_t B = c.batch, T = c.frames, C = c.residual_channels, L = c.residual_layers, k = c.kernel_size;
  l.NBcap = (c.branches == DRB_BRANCH_COND_UNCOND || c.branches == DRB_BRANCH_COND_ZEROSPEC || c.branches == DRB_BRANCH_COND_LEARNED) ? 2 * c.batch : c.batch;
  // DRB_BRANCH_COND_LEARNED: the second branch is conditioned on a learned spectrogram (condition='trainable_spec',
  // model/diffwave.py:657-658).  It is kept as `batch` more clips behind the real ones, so every kernel sees 2 x batch
  // conditional rolls and nothing in the step changes shape.
  l.Bs = (c.branches == DRB_BRANCH_COND_LEARNED || c.branches == DRB_BRANCH_LEARNED) ? 2 * c.batch : c.batch;
  l.Mp = (c.n_mels + 63) / 64 * 64;
  l.KC = (int)(k * C);
  const size_t NB = l.NBcap, Mp = l.Mp, rows = NB * T;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t r = off; off += align_up(bytes); return r; };
  l.dtab = take(L * c.timesteps * C * 4);
  l.dtab2 = take(L * c.timesteps * C * 4);
  l.iota = take(B * 4);
  l.emb1 = take((size_t)c.timesteps * 512 * 4);
  l.emb2 = take((size_t)c.timesteps * 512 * 4);
  l.x32 = take(rows * C * 4);
  l.skip = take(rows * C * 4);
  l.hbuf = take(rows * C * 4);
  l.spec32 = take((size_t)l.Bs * T * Mp * 4);
  l.bias = take(L * 4 * 2 * C * 4);
  l.wtmp = take(2 * C * k * C * 4);
  l.mel = take(mel_workspace_bytes(c));
  l.xpp = take(B * T * c.pitches * 4);
  l.range = take(256);
  if (c.precision == DRB_PREC_FP32) {
    l.wd32 = take(L * 2 * C * k * C * 4);
    l.wc32 = take(L * 2 * C * Mp * 4);
    l.ybuf = take(rows * 2 * C * 4);
    l.z32 = take(rows * C * 4);
  } else {
    l.xh = take(rows * C * 2); l.xl = take(rows * C * 2);
    l.zh = take(L * rows * C * 2); l.zl = take(L * rows * C * 2);  // gated activations of every layer (head GEMM)
    l.sh = take(B * T * Mp * 2); l.sl = take(B * T * Mp * 2);
    l.wdh = take(L * 2 * C * k * C * 2); l.wdl = take(L * 2 * C * k * C * 2);
    l.wch = take(L * 2 * C * Mp * 2); l.wcl = take(L * 2 * C * Mp * 2);
    l.woh = take(L * 2 * C * C * 2); l.wol = take(L * 2 * C * C * 2);
    l.wcomp32 = take(C * L * C * 4); l.wcomph = take(C * L * C * 2); l.wcompl = take(C * L * C * 2);
    l.bcomp = take(C * 4); l.bsum = take(C * 4);
    l.wscale = take((4 * L + 2) * 4 * 4);
    l.wouth = take(256 * C * 2); l.woutl = take(256 * C * 2);
    if (c.precision == DRB_PREC_F16N4) {
      l.xs = take(NB * (C / 64) * T * 8);
      l.wsf4 = take(L * (2 * C / 256) * (k * C / 64) * 2048);
    }  // f16f8: {SW, 1/(SA*SW), scratch, -} per gate / out weight set and the head
    if (c.branches != DRB_BRANCH_UNCOND && c.precision != DRB_PREC_F16F8 && c.precision != DRB_PREC_BF16) {
      l.wcpad = take(L * 2 * C * Mp * 4); l.cond = take(L * (size_t)l.Bs * T * 2 * C * 4);
      l.sp5h = take(B * T * Mp * 2); l.sp5l = take(B * T * Mp * 2); l.wc5h = take(L * 2 * C * Mp * 2); l.wc5l = take(L * 2 * C * Mp * 2);
    }
  }
  l.total = off;
  return l;
}

}  // namespace drb

using namespace drb;

struct drb_plan {
  drb_config cfg;
  Layout lay;
  char* ws;
  int NB, n_cond;      // active branches
  bool zero_spec;      // second branch = conditional forward on an all-zero spectrogram (cfdg_ddim_x0)
  bool tables_ready, spec_ready;
  bool cond_ready = false;   // cond tables hold the conditioner projections of the CURRENT spectrogram
  bool cond_use = true;      // drb_plan_use_cond_tables: steps read the tables when they are ready
  bool learned = false;      // DRB_BRANCH_COND_LEARNED / DRB_BRANCH_LEARNED is in force: rolls read the learned clips
  int clip0 = 0;             // first clip roll 0 reads: 0, or batch when EVERY roll reads a learned clip (DRB_BRANCH_LEARNED)
  bool uspec_ready = false;  // drb_plan_set_uncond_spec has filled the learned clips of spec32
  bool ucond_ready = false;  // ... and their half of the conditioner tables is built
  int pair = 1;        // CTA pairs (cta_group::2); DRB_NO_PAIR=1 selects the single-CTA kernels, for A/B runs
  std::vector<int> dil;
  // weight pointers used in place (caller keeps them alive)
  const float *in_w, *in_b, *e1w, *e1b, *e2w, *e2b, *skw, *skb, *hdw, *hdb;
  std::vector<const float*> dpw, dpb, wo32, bo;
  MelPlan* mel;
  UmmaMaps maps;
  std::vector<UmmaLayer> layers;
  std::vector<CUtensorMap> win_h, win_l;  // per layer: x operand maps whose box covers the layer's whole tap window
  std::vector<CUtensorMap> win4, wd4, wsf4;   // f16n4 per layer: aux window map, aux weight map, weight scale atoms
  CUtensorMap xl4;                            // f16n4: RES store map of the 64-byte e2m1 rows
  int window = 1, persistent = 1;
  const int32_t* steps = nullptr;   // per-sample diffusion steps (device, [batch]); nullptr = the t_index arguments
  int condpre = 1;     // conditioner projections precomputed per clip, added in the gate epilogue (DRB_NO_CONDPRE=1: off)
  int share0 = 1;      // layer-0 branch sharing (DRB_NO_SHARE0=1 turns it off, for A/B runs); needs condpre
  // optional per-kernel timing (drb_plan_profile): events recorded on the launching stream around every kernel class
  cudaStream_t copy_stream = nullptr;  // trajectory copies (drb_sample_loop)
  cudaEvent_t ev_step[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
  bool prof = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<int, int>> ev_spans[5];  // 0 gate kernel, 1 out kernel, 2 in_proj+prep, 3 head (GEMM + projection), 4 head projection + posterior alone
  size_t ev_used = 0;
  int ev_mark(cudaStream_t s) {
    if (ev_used == ev_pool.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return -1; ev_pool.push_back(e); }
    cudaEventRecord(ev_pool[ev_used], s);
    return (int)ev_used++;
  }
  template <class Tp> Tp* at(size_t off) const { return reinterpret_cast<Tp*>(ws + off); }
  float* bias_ptr(int layer, int which) const {  // which: 0 cond-interleaved, 1 unc-interleaved, 2 cond-natural, 3 unc-natural
    return at<float>(lay.bias) + ((size_t)layer * 4 + which) * 2 * cfg.residual_channels;
  }
  // tensor-path arithmetic: kernel template mode (0 bf16, 1 bf16x3, 2 f16f8) and operand format (1 bf16 hi/lo, 2 fp16+e4m3)
  // f16n4: only the gate kernel's activation / weight operands are block-scaled fp4; z operands, RES and HEAD stay f16e5
  bool n4() const { return cfg.precision == DRB_PREC_F16N4; }
  int prec() const { return cfg.precision == DRB_PREC_BF16X3 ? 1 : cfg.precision == DRB_PREC_F16F8 ? 2 : (cfg.precision == DRB_PREC_F16E5 || n4()) ? 3 : 0; }
  int fmt() const { return cfg.precision == DRB_PREC_FP32 ? 0 : cfg.precision == DRB_PREC_F16F8 ? 2 : (cfg.precision == DRB_PREC_F16E5 || n4()) ? 3 : 1; }
  int xfmt() const { return n4() ? 4 : fmt(); }   // format of the x operand pair (in_proj / RES -> gate kernel)
  float* wscale(int slot) const { return at<float>(lay.wscale) + 4 * slot; }  // slot 2l: gate weights, 2l+1: Wo, 2L: head, 2L+1+l: Wc (f16n4)
  int wc_slot(int layer) const { return n4() ? 2 * cfg.residual_layers + 1 + layer : 2 * layer; }   // f16n4 scales the conv weights alone
  int wout_slot() const { return 3 * cfg.residual_layers + 1; }
  int wc5_slot(int layer) const { return 3 * cfg.residual_layers + 2 + layer; }
  // conditioner tables on the tensor cores at fp32 grade: f16x3 operands (22 mantissa bits) and K = Mp = 256 per accumulation
  // chain, i.e. at the floor of the tensor core's fp32 accumulation (8e-6 on the network output, csrc/train.cu)
  CUtensorMap sp5h, sp5l;
  std::vector<CUtensorMap> wc5h, wc5l;
  int cond_tc = 0;
  // tensor-core head: HEAD leaves relu(skip_projection) as an operand pair, output_projection + guidance + posterior run as
  // one tcgen05 kernel (DRB_NO_HEAD_TC=1: fp32 h + the CUDA-core projection kernel, for A/B runs)
  int head_tc = 0;
  std::vector<CUtensorMap> cond32;   // per layer: fp32 [B][T][2C] map of the conditioner table (tensor-core build)
  bool dtab_alt = false;       // the per-roll tables of drb_plan_set_step_embeddings are in force (row = roll index)
  const float* dvec(int layer, int t) const {
    return at<float>(dtab_alt ? lay.dtab2 : lay.dtab) + ((size_t)layer * cfg.timesteps + t) * cfg.residual_channels;
  }
};

extern "C" {

int drb_version(void) { return DRB_VERSION; }
const char* drb_last_error(void) { return g_err; }
int64_t drb_launch_count(int32_t reset) { int64_t v = g_launches; if (reset) g_launches = 0; return v; }

size_t drb_plan_workspace_bytes(const drb_config* cfg) {
  if (!cfg || !cfg_ok(*cfg)) { set_error("invalid drb_config"); return 0; }
  return make_layout(effective_config(*cfg)).total;
}

int drb_plan_set_branches(drb_plan* p, int32_t branches) {
  if (!p) return DRB_E_INVALID;
  const int B = p->cfg.batch;
  int NB, nc;
  if (branches == DRB_BRANCH_COND_UNCOND) { NB = 2 * B; nc = B; }
  else if (branches == DRB_BRANCH_COND) { NB = B; nc = B; }
  else if (branches == DRB_BRANCH_UNCOND) { NB = B; nc = 0; }
  else if (branches == DRB_BRANCH_COND_ZEROSPEC) { NB = 2 * B; nc = B; }
  else if (branches == DRB_BRANCH_COND_LEARNED) {   // every roll reads a conditioner table: its clip's, or the learned one
    if (p->lay.Bs != 2 * B) { set_error("plan was not created with DRB_BRANCH_COND_LEARNED"); return DRB_E_INVALID; }
    NB = 2 * B; nc = 2 * B;
  }
  else if (branches == DRB_BRANCH_LEARNED) {        // one forward per roll, each conditioned on the learned table
    if (p->lay.Bs != 2 * B) { set_error("plan was not created with DRB_BRANCH_COND_LEARNED / DRB_BRANCH_LEARNED"); return DRB_E_INVALID; }
    NB = B; nc = B;
  }
  else { set_error("bad branches %d", branches); return DRB_E_INVALID; }
  if (NB > p->lay.NBcap) { set_error("plan was created for a single branch"); return DRB_E_INVALID; }
  p->NB = NB; p->n_cond = nc; p->zero_spec = branches == DRB_BRANCH_COND_ZEROSPEC;
  p->learned = branches == DRB_BRANCH_COND_LEARNED || branches == DRB_BRANCH_LEARNED;
  p->clip0 = branches == DRB_BRANCH_LEARNED ? B : 0;
  return 0;
}

int drb_plan_set_steps(drb_plan* p, const int32_t* steps_dev) {
  if (!p) return DRB_E_INVALID;
  p->steps = steps_dev;
  return 0;
}

int drb_plan_precision(const drb_plan* p) { return p ? p->cfg.precision : DRB_E_INVALID; }

int drb_plan_create(drb_plan** out, const drb_config* cfg_in, const drb_weights* w, void* workspace, size_t ws_bytes,
                    void* stream) {
  if (!out || !cfg_in || !w || !workspace) { set_error("null argument"); return DRB_E_INVALID; }
  if (!cfg_ok(*cfg_in)) { set_error("invalid drb_config"); return DRB_E_INVALID; }
  const drb_config cfg_eff = effective_config(*cfg_in);
  const drb_config* cfg = &cfg_eff;
  cudaStream_t s = (cudaStream_t)stream;
  Layout lay = make_layout(*cfg);
  if (ws_bytes < lay.total || ((uintptr_t)workspace & 255)) {
    set_error("workspace: need %zu bytes 256-aligned, got %zu", lay.total, ws_bytes);
    return DRB_E_WORKSPACE;
  }
  drb_plan* p = new drb_plan();
  p->cfg = *cfg; p->lay = lay; p->ws = (char*)workspace; p->mel = nullptr;
  p->tables_ready = false; p->spec_ready = false;
  { const char* e = getenv("DRB_NO_PAIR"); p->pair = (e && e[0] == '1') ? 0 : 1; }
  { const char* e = getenv("DRB_NO_WINDOW"); p->window = (e && e[0] == '1') ? 0 : 1; }
  { const char* e = getenv("DRB_NO_PERSIST"); p->persistent = (e && e[0] == '1') ? 0 : 1; }
  { const char* e = getenv("DRB_NO_SHARE0"); p->share0 = (e && e[0] == '1') ? 0 : 1; }
  { const char* e = getenv("DRB_NO_CONDPRE"); p->condpre = (e && e[0] == '1') ? 0 : 1; }
  if (lay.cond == 0 || !p->persistent) p->condpre = 0;   // only the persistent gate kernel (bf16x3 / f16e5) implements it
  if (!p->condpre || lay.NBcap != 2 * cfg->batch) p->share0 = 0;
  if (lay.Bs != cfg->batch && cfg->precision != DRB_PREC_FP32 && !p->condpre) {
    set_error("DRB_BRANCH_COND_LEARNED needs the per-clip conditioner tables (precision fp32, bf16x3, f16e5 or f16n4; persistent kernels)");
    delete p; return DRB_E_INVALID;
  }
  const int C = cfg->residual_channels, L = cfg->residual_layers, k = cfg->kernel_size, Mp = lay.Mp, T = cfg->frames;
  p->in_w = w->input_projection_w; p->in_b = w->input_projection_b;
  p->e1w = w->emb_projection1_w; p->e1b = w->emb_projection1_b; p->e2w = w->emb_projection2_w; p->e2b = w->emb_projection2_b;
  p->skw = w->skip_projection_w; p->skb = w->skip_projection_b; p->hdw = w->head_projection_w; p->hdb = w->head_projection_b;
  int rc = drb_plan_set_branches(p, cfg->branches);
  if (rc) { delete p; return rc; }
#define PLAN_TRY(expr) do { int _r = (expr); if (_r) { drb_plan_destroy(p); return _r; } } while (0)
#define PLAN_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("%s -> %s", #expr, cudaGetErrorString(_e)); drb_plan_destroy(p); return (int)_e; } } while (0)
  const bool tensor = cfg->precision != DRB_PREC_FP32;
  if (tensor) PLAN_TRY(umma_init());
  PLAN_TRY(mel_create(&p->mel, *cfg, w->stft_window, w->mel_fb, p->ws + lay.mel, mel_workspace_bytes(*cfg), s));
  for (int i = 0; i < L; ++i) {
    int e = i % cfg->dilation_bound, d = 1;
    for (int j = 0; j < e; ++j) d *= cfg->dilation_base;  // dilation_base**(i % dilation_bound)  model/diffwave.py:624
    p->dil.push_back(d);
    p->dpw.push_back(w->diffusion_projection_w[i]); p->dpb.push_back(w->diffusion_projection_b[i]);
    p->wo32.push_back(w->output_projection_w[i]); p->bo.push_back(w->output_projection_b[i]);
    PLAN_TRY(launch_bias1(w->dilated_conv_b[i], w->conditioner_projection_b[i], w->conditioner_projection_w[i],
                          p->bias_ptr(i, 0), p->bias_ptr(i, 1), p->bias_ptr(i, 2), p->bias_ptr(i, 3), C, cfg->n_mels, s));
    float* tmp = p->at<float>(lay.wtmp);
    if (!tensor) {
      PLAN_TRY(launch_repack_conv_fp32(w->dilated_conv_w[i], p->at<float>(lay.wd32) + (size_t)i * 2 * C * k * C, 2 * C, C, k, s));
      PLAN_TRY(launch_pad_rows(w->conditioner_projection_w[i], p->at<float>(lay.wc32) + (size_t)i * 2 * C * Mp, 2 * C,
                               cfg->n_mels, Mp, s));
    } else {
      PLAN_TRY(launch_repack_conv_fp32(w->dilated_conv_w[i], tmp, 2 * C, C, k, s));
      const int fmt = p->fmt();
      const int dm = fmt >= 2 ? 2 : 0, da = fmt >= 2 ? 3 : 0, am = fmt >= 2 ? 2 : 1;  // tensor-map dtypes, aux width factor
      if (fmt >= 2) {   // per-tensor power-of-two weight scales (f16f8: e4m3 range; f16e5: keeps small weights normal in fp16)
        const float sa = fmt == 2 ? F8_SA : 1.f;
        if (!p->n4())
          PLAN_TRY(launch_weight_scale(tmp, (size_t)2 * C * k * C, w->conditioner_projection_w[i], (size_t)2 * C * cfg->n_mels,
                                       p->wscale(2 * i), sa, s));
        PLAN_TRY(launch_weight_scale(w->output_projection_w[i], (size_t)2 * C * C, nullptr, 0, p->wscale(2 * i + 1), sa, s));
      }
      char* wdh = p->ws + lay.wdh + (size_t)i * 2 * C * k * C * 2;
      char* wdl = p->ws + lay.wdl + (size_t)i * 2 * C * k * C * 2;
      if (p->n4()) {   // own scale (only the conv weights, up to 2^15), fp16 main + e2m1 aux + scale atoms
        uint8_t* sfa = p->at<uint8_t>(lay.wsf4) + (size_t)i * (2 * C / 256) * (k * C / 64) * 2048;
        PLAN_TRY(launch_weight_scale(tmp, (size_t)2 * C * k * C, nullptr, 0, p->wscale(2 * i), 1.f, s, N4_WTARGET));
        PLAN_TRY(launch_repack_n4(tmp, wdh, wdl, sfa, 2 * C, k * C, C, p->wscale(2 * i), s));
      } else {
        PLAN_TRY(launch_repack_split(tmp, wdh, wdl, 2 * C, k * C, k * C, C, fmt, p->wscale(2 * i), s));
      }
      char* wch = p->ws + lay.wch + (size_t)i * 2 * C * Mp * 2;
      char* wcl = p->ws + lay.wcl + (size_t)i * 2 * C * Mp * 2;
      if (p->n4() && fmt >= 2)   // the f16n4 conv-weight scale targets 2^15 for the e2m1 split: the conditioner weights get their own
        PLAN_TRY(launch_weight_scale(w->conditioner_projection_w[i], (size_t)2 * C * cfg->n_mels, nullptr, 0, p->wscale(p->wc_slot(i)), 1.f, s));
      PLAN_TRY(launch_repack_split(w->conditioner_projection_w[i], wch, wcl, 2 * C, cfg->n_mels, Mp, C, fmt, p->wscale(p->wc_slot(i)), s));
      char* woh = p->ws + lay.woh + (size_t)i * 2 * C * C * 2;
      char* wol = p->ws + lay.wol + (size_t)i * 2 * C * C * 2;
      PLAN_TRY(launch_repack_split(w->output_projection_w[i], woh, wol, 2 * C, C, C, 0, fmt, p->wscale(2 * i + 1), s));
      UmmaLayer ul;
      PLAN_TRY(make_tmap_2d(&ul.wd_h, wdh, 2 * C, (uint64_t)k * C, 128, dm));
      PLAN_TRY(make_tmap_2d(&ul.wd_l, wdl, 2 * C, (uint64_t)am * k * C, 128, da));
      PLAN_TRY(make_tmap_2d(&ul.wc_h, wch, 2 * C, Mp, 128, dm));
      PLAN_TRY(make_tmap_2d(&ul.wc_l, wcl, 2 * C, (uint64_t)am * Mp, 128, da));
      PLAN_TRY(make_tmap_2d(&ul.wo_h, woh, 2 * C, C, 128, dm));
      PLAN_TRY(make_tmap_2d(&ul.wo_l, wol, 2 * C, (uint64_t)am * C, 128, da));
      p->layers.push_back(ul);
      if (p->n4()) {
        CUtensorMap m4, ms;
        PLAN_TRY(make_tmap_2d_bytes(&m4, wdl, 2 * C, (uint64_t)k * C, 128, 64, 64));
        PLAN_TRY(make_tmap_2d_bytes(&ms, p->at<uint8_t>(lay.wsf4) + (size_t)i * (2 * C / 256) * (k * C / 64) * 2048,
                                    (uint64_t)(2 * C / 256) * (k * C / 64) * 16, 128, 16, 128, 0));
        p->wd4.push_back(m4); p->wsf4.push_back(ms);
      }
    }
  }
  if (p->condpre)   // fp32 conditioner weights, K padded to Mp (the spectrogram rows are zero-padded alike)
    for (int i = 0; i < L; ++i)
      PLAN_TRY(launch_pad_rows(w->conditioner_projection_w[i], p->at<float>(lay.wcpad) + (size_t)i * 2 * C * Mp, 2 * C,
                               cfg->n_mels, Mp, s));
  if (tensor) {
    const uint64_t NBc = lay.NBcap;
    const int fmt = p->fmt();
    const int dm = fmt >= 2 ? 2 : 0, da = fmt >= 2 ? 3 : 0;
    const uint64_t am = fmt >= 2 ? 2 : 1;
    PLAN_TRY(make_tmap_3d(&p->maps.xh, p->ws + lay.xh, NBc, T, C, 128, dm));
    PLAN_TRY(make_tmap_3d(&p->maps.xl, p->ws + lay.xl, NBc, T, am * C, 128, da));
    for (int i = 0; i < L; ++i) {
      const int rows = 128 + (k - 1) * p->dil[i];
      CUtensorMap mh, ml;
      if (rows <= 192) {
        PLAN_TRY(make_tmap_3d(&mh, p->ws + lay.xh, NBc, T, C, rows, dm));
        PLAN_TRY(make_tmap_3d(&ml, p->ws + lay.xl, NBc, T, am * C, rows, da));
      } else { mh = p->maps.xh; ml = p->maps.xl; }
      p->win_h.push_back(mh); p->win_l.push_back(ml);
      if (p->n4()) {
        if (rows > 192) { set_error("f16n4: tap window of layer %d (%d frames) exceeds 192", i, rows); drb_plan_destroy(p); return DRB_E_INVALID; }
        CUtensorMap m4;
        PLAN_TRY(make_tmap_3d_bytes(&m4, p->ws + lay.xl, NBc, T, C, rows, 64, 64));
        p->win4.push_back(m4);
      }
    }
    if (p->n4()) PLAN_TRY(make_tmap_3d_bytes(&p->xl4, p->ws + lay.xl, NBc, T, C, 128, 64, 64));
    PLAN_TRY(make_tmap_3d(&p->maps.zh, p->ws + lay.zh, (uint64_t)L * NBc, T, C, 128, dm));
    PLAN_TRY(make_tmap_3d(&p->maps.zl, p->ws + lay.zl, (uint64_t)L * NBc, T, am * C, 128, da));
    PLAN_TRY(make_tmap_3d(&p->maps.sh, p->ws + lay.sh, cfg->batch, T, Mp, 128, dm));
    PLAN_TRY(make_tmap_3d(&p->maps.sl, p->ws + lay.sl, cfg->batch, T, am * Mp, 128, da));
    if (lay.cond) {   // f16x3 operands of the per-clip conditioner tables (DRB_COND_SIMT=1: fp32 CUDA-core build, for A/B runs)
      const char* e = getenv("DRB_COND_SIMT");
      p->cond_tc = (2 * C) % 256 == 0 && Mp % 64 == 0 && !(e && e[0] == '1');
      if (p->cond_tc) {
        PLAN_TRY(make_tmap_3d(&p->sp5h, p->ws + lay.sp5h, cfg->batch, T, Mp, 128, 2));
        PLAN_TRY(make_tmap_3d(&p->sp5l, p->ws + lay.sp5l, cfg->batch, T, Mp, 128, 2));
        for (int i = 0; i < L; ++i) {
          char* h5 = p->ws + lay.wc5h + (size_t)i * 2 * C * Mp * 2;
          char* l5 = p->ws + lay.wc5l + (size_t)i * 2 * C * Mp * 2;
          PLAN_TRY(launch_weight_scale(w->conditioner_projection_w[i], (size_t)2 * C * cfg->n_mels, nullptr, 0, p->wscale(p->wc5_slot(i)), 1.f, s));
          PLAN_TRY(launch_repack_split(w->conditioner_projection_w[i], h5, l5, 2 * C, cfg->n_mels, Mp, 0, 5, p->wscale(p->wc5_slot(i)), s));
          CUtensorMap mh, ml;
          PLAN_TRY(make_tmap_2d(&mh, h5, 2 * C, Mp, 128, 2));
          PLAN_TRY(make_tmap_2d(&ml, l5, 2 * C, Mp, 128, 2));
          p->wc5h.push_back(mh); p->wc5l.push_back(ml);
        }
      }
    }
    if (lay.cond)   // conditioner tables built on the tensor cores (drb_cond_tables): one fp32 map per layer
      for (int i = 0; i < L; ++i) {
        CUtensorMap mc;
        PLAN_TRY(make_tmap_3d(&mc, p->at<float>(lay.cond) + (size_t)i * lay.Bs * T * 2 * C, cfg->batch, T, 2 * (uint64_t)C, 128, 1));
        p->cond32.push_back(mc);
      }
    PLAN_TRY(make_tmap_3d(&p->maps.x32, p->ws + lay.x32, NBc, T, C, 128, 1));
    PLAN_TRY(make_tmap_3d(&p->maps.h32, p->ws + lay.hbuf, NBc, T, C, 128, 1));
    {   // the same storage seen as an operand pair [NB][T][C] (2 + 2 bytes per element) and the padded output_projection
      const size_t rows_all = (size_t)NBc * T;
      PLAN_TRY(make_tmap_3d(&p->maps.hh, p->ws + lay.hbuf, NBc, T, C, 128, dm));
      PLAN_TRY(make_tmap_3d(&p->maps.hl, p->ws + lay.hbuf + rows_all * C * 2, NBc, T, am * C, 128, da));
      const char* e = getenv("DRB_NO_HEAD_TC");
      p->head_tc = (p->prec() == 1 || p->prec() == 3) && cfg->pitches <= 256 && (cfg->pitches % 4) == 0 && !(e && e[0] == '1');
      if (p->head_tc) {
        float* tmp = p->at<float>(lay.wtmp);
        PLAN_CUDA(cudaMemsetAsync(tmp, 0, (size_t)256 * C * 4, s));
        PLAN_CUDA(cudaMemcpyAsync(tmp, p->hdw, (size_t)cfg->pitches * C * 4, cudaMemcpyDeviceToDevice, s));
        if (fmt >= 2) PLAN_TRY(launch_weight_scale(tmp, (size_t)256 * C, nullptr, 0, p->wscale(p->wout_slot()), fmt == 2 ? F8_SA : 1.f, s));
        PLAN_TRY(launch_repack_split(tmp, p->ws + lay.wouth, p->ws + lay.woutl, 256, C, C, 0, fmt, p->wscale(p->wout_slot()), s));
        PLAN_TRY(make_tmap_2d(&p->maps.wout_h, p->ws + lay.wouth, 256, C, 128, dm));
        PLAN_TRY(make_tmap_2d(&p->maps.wout_l, p->ws + lay.woutl, 256, am * C, 128, da));
      }
    }
    // skip sum + 1/sqrt(L) + skip_projection composed into one [C][L*C] weight over the stored z of all layers
    for (int i = 0; i < L; ++i)
      PLAN_TRY(launch_compose_skip(p->skw, p->wo32[i], p->at<float>(lay.wcomp32), C, L, i, s));
    PLAN_TRY(launch_compose_bias(p->skw, p->skb, p->bo.data(), C, L, p->at<float>(lay.bsum), p->at<float>(lay.bcomp), s));
    if (fmt >= 2) PLAN_TRY(launch_weight_scale(p->at<float>(lay.wcomp32), (size_t)C * L * C, nullptr, 0, p->wscale(2 * L),
                                               fmt == 2 ? F8_SA : 1.f, s));
    PLAN_TRY(launch_repack_split(p->at<float>(lay.wcomp32), p->ws + lay.wcomph, p->ws + lay.wcompl, C, L * C, L * C, 0, fmt,
                                 p->wscale(2 * L), s));
    PLAN_TRY(make_tmap_2d(&p->maps.wcomp_h, p->ws + lay.wcomph, C, (uint64_t)L * C, 128, dm));
    PLAN_TRY(make_tmap_2d(&p->maps.wcomp_l, p->ws + lay.wcompl, C, am * L * C, 128, da));
  }
  if (cudaMemsetAsync(p->ws + lay.range, 0, 256, s) != cudaSuccess) { set_error("plan_create: memset failed"); drb_plan_destroy(p); return DRB_E_INVALID; }
#undef PLAN_TRY
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { set_error("plan_create sync: %s", cudaGetErrorString(e)); drb_plan_destroy(p); return (int)e; }
  *out = p;
  return 0;
}

int drb_plan_destroy(drb_plan* p) {
  if (!p) return 0;
  mel_destroy(p->mel);
  for (auto e : p->ev_pool) cudaEventDestroy(e);
  if (p->copy_stream) {
    cudaStreamSynchronize(p->copy_stream);
    for (int k = 0; k < 2; ++k) { cudaEventDestroy(p->ev_step[k]); cudaEventDestroy(p->ev_copy[k]); }
    cudaStreamDestroy(p->copy_stream);
  }
  delete p;
  return 0;
}

int drb_time_tables(drb_plan* p, const float* emb_table, void* stream) {
  if (!p || !emb_table) return DRB_E_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  const int TS = p->cfg.timesteps, C = p->cfg.residual_channels;
  SimtGemm g;
  g.A = emb_table; g.lda = 128; g.T = TS; g.Ck = 128; g.W = p->e1w; g.ldw = 128; g.bias = p->e1b; g.act = 2;
  g.C = p->at<float>(p->lay.emb1); g.ldc = 512; g.M = TS; g.N = 512;
  int r = launch_simt_gemm(g, s); if (r) return r;
  g.A = p->at<float>(p->lay.emb1); g.lda = 512; g.Ck = 512; g.W = p->e2w; g.ldw = 512; g.bias = p->e2b;
  g.C = p->at<float>(p->lay.emb2);
  r = launch_simt_gemm(g, s); if (r) return r;
  for (int i = 0; i < p->cfg.residual_layers; ++i) {
    g.A = p->at<float>(p->lay.emb2); g.W = p->dpw[i]; g.bias = p->dpb[i]; g.act = 0;
    g.C = p->at<float>(p->lay.dtab) + (size_t)i * TS * C; g.ldc = C; g.N = C;
    r = launch_simt_gemm(g, s); if (r) return r;
  }
  p->tables_ready = true;
  return 0;
}

// Fractional diffusion steps (DiffusionEmbedding._lerp_embedding, model/diffwave.py:76-81): the caller interpolates the sinusoid
// table per roll; this runs projection1 / projection2 / every diffusion_projection on those `batch` rows into a second table
// whose row index is the ROLL, and makes the plan read it (with roll-indexed steps) until it is cleared with emb_rows == NULL.
namespace { __global__ void iota_i32_kernel(int* p, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; } }
int drb_plan_set_step_embeddings(drb_plan* p, const float* emb_rows, void* stream) {
  if (!p) return DRB_E_INVALID;
  if (!emb_rows) { p->dtab_alt = false; p->steps = nullptr; return 0; }
  cudaStream_t s = (cudaStream_t)stream;
  const int B = p->cfg.batch, TS = p->cfg.timesteps, C = p->cfg.residual_channels;
  if (B > TS) { set_error("set_step_embeddings: batch %d exceeds the table rows (timesteps %d)", B, TS); return DRB_E_INVALID; }
  SimtGemm g;
  g.A = emb_rows; g.lda = 128; g.T = B; g.Ck = 128; g.W = p->e1w; g.ldw = 128; g.bias = p->e1b; g.act = 2;
  g.C = p->at<float>(p->lay.emb1); g.ldc = 512; g.M = B; g.N = 512;
  int r = launch_simt_gemm(g, s); if (r) return r;
  g.A = p->at<float>(p->lay.emb1); g.lda = 512; g.Ck = 512; g.W = p->e2w; g.ldw = 512; g.bias = p->e2b;
  g.C = p->at<float>(p->lay.emb2);
  r = launch_simt_gemm(g, s); if (r) return r;
  for (int i = 0; i < p->cfg.residual_layers; ++i) {
    g.A = p->at<float>(p->lay.emb2); g.W = p->dpw[i]; g.bias = p->dpb[i]; g.act = 0;
    g.C = p->at<float>(p->lay.dtab2) + (size_t)i * TS * C; g.ldc = C; g.N = C;
    r = launch_simt_gemm(g, s); if (r) return r;
  }
  iota_i32_kernel<<<(B + 255) / 256, 256, 0, s>>>(p->at<int>(p->lay.iota), B);
  DRB_LAUNCH_CHECK();
  p->dtab_alt = true; p->steps = p->at<int>(p->lay.iota);
  // NB: emb1 / emb2 are scratch shared with drb_time_tables, whose results (dtab) stay valid
  return 0;
}

int drb_mel_forward(drb_plan* p, const float* waveform, float* spec_out, int32_t it0, int32_t it1, int32_t if0,
                    int32_t if1, void* stream) {
  if (!p || !waveform) return DRB_E_INVALID;
  const bool tensor = p->cfg.precision != DRB_PREC_FP32;
  int r = mel_forward(p->mel, waveform, spec_out, p->at<float>(p->lay.spec32), tensor ? p->ws + p->lay.sh : nullptr,
                      tensor ? p->ws + p->lay.sl : nullptr, p->fmt(), p->lay.Mp, p->cfg.frames, it0, it1, if0, if1,
                      (cudaStream_t)stream);
  p->cond_ready = false;   // built lazily: drb_cond_tables / drb_sample_loop (a single forward is cheaper without them)
  if (r == 0) p->spec_ready = true;
  return r;
}

// conditioner_projection_l(spec) (model/diffwave.py:143) does not depend on the timestep: computed once per clip, at fp32
// grade, for every layer; the gate kernel then adds it in its epilogue instead of contracting it every step.  One call covers
// `batch` clips starting at clip `clip0` of spec32 (0: the clips of drb_mel_forward; batch: the learned clips of
// drb_plan_set_uncond_spec) and writes the same clip range of every layer's table -- the launches are the same either way.
static int build_cond_tables(drb_plan* p, int clip0, void* stream) {
  const drb_config& c = p->cfg;
  const size_t C2 = 2 * (size_t)c.residual_channels, per = (size_t)p->lay.Bs * c.frames * C2;
  const float* spec = p->at<float>(p->lay.spec32) + (size_t)clip0 * c.frames * p->lay.Mp;
  float* cond = p->at<float>(p->lay.cond) + (size_t)clip0 * c.frames * C2;
  if (p->cond_tc && p->cfg.precision != DRB_PREC_FP32) {
    // default: tcgen05 with f16x3 operands (fp16 hi + fp16 lo) and a single K = Mp chain per output -- fp32-grade, 15 short
    // launches instead of 15 x 0.37 ms of fp32 FMA per clip
    int r = launch_split_pair(spec, p->lay.Mp, nullptr, 0, c.frames, nullptr, p->ws + p->lay.sp5h, p->ws + p->lay.sp5l,
                              c.batch * c.frames, p->lay.Mp, (cudaStream_t)stream, 5);
    if (r) return r;
    for (int l = 0; l < c.residual_layers; ++l) {
      UmmaConvLin cv;
      cv.ah = &p->sp5h; cv.al = &p->sp5l; cv.wh = &p->wc5h[l]; cv.wl = &p->wc5l[l]; cv.prec = 4; cv.pair = p->pair;
      cv.NB = c.batch; cv.T = c.frames; cv.Cin = p->lay.Mp; cv.Nout = (int)C2; cv.taps = 1; cv.dil = 1; cv.Mp = 0;
      cv.inv_scale = p->wscale(p->wc5_slot(l)) + 1; cv.bias = nullptr; cv.out = cond + (size_t)l * per; cv.ldo = (int)C2;
      r = launch_umma_conv_lin(cv, (cudaStream_t)stream); if (r) return r;
    }
    return 0;
  }
  static int tc_cond = -1;
  if (tc_cond < 0) { const char* e = getenv("DRB_COND_TC"); tc_cond = (e && e[0] == '1') ? 1 : 0; }
  if (tc_cond && clip0 == 0 && p->lay.Bs == c.batch && (p->prec() == 1 || p->prec() == 3) && (int)p->cond32.size() == c.residual_layers) {
    // OPT-IN tensor-core build (DRB_COND_TC=1): the spectrogram operand pair [B][T][Mp] against the conditioner weights in the
    // plan's pair format, K = Mp, plain fp32 result in natural channel order: 15 short tcgen05 launches instead of 15 x 0.37 ms
    // of fp32 FMA per clip.  Measured (B200, configs[1]): per-clip work 5.5 -> 0.9 ms, i.e. +3.6 % on a 20-step e2e run and
    // +0.3 % on a real 200-step chain, but the tables then carry the pair format's rounding into EVERY step: 200-step B=32
    // chain error 2.9e-4 -> 3.6e-4 (f16n4), 2.0e-4 -> 2.5e-4 (f16e5).  The default keeps the exact fp32 tables: parity first.
    for (int l = 0; l < c.residual_layers; ++l) {
      UmmaZGemm uz;
      uz.pair = p->pair; uz.persistent = 0; uz.NB = c.batch; uz.T = c.frames; uz.C = (int)C2; uz.prec = p->prec(); uz.mode = 2;
      uz.inv_scale = p->wscale(p->wc_slot(l)) + 1; uz.groups = 1; uz.z_group0 = 0; uz.group_stride = 0;
      uz.w_h = &p->layers[l].wc_h; uz.w_l = &p->layers[l].wc_l; uz.out32 = &p->cond32[l]; uz.bias = nullptr; uz.dnext = nullptr;
      uz.a_h = &p->maps.sh; uz.a_l = &p->maps.sl; uz.nslabs64 = p->lay.Mp / 64;
      int r = launch_umma_zgemm(p->maps, uz, (cudaStream_t)stream); if (r) return r;
    }
    return 0;
  }
  for (int l = 0; l < c.residual_layers; ++l) {
    SimtGemm g;
    g.A = spec; g.lda = p->lay.Mp; g.T = c.frames; g.Ck = p->lay.Mp;
    g.W = p->at<float>(p->lay.wcpad) + (size_t)l * C2 * p->lay.Mp; g.ldw = p->lay.Mp;
    g.C = cond + (size_t)l * per; g.ldc = (int)C2; g.M = c.batch * c.frames; g.N = (int)C2;
    int r = launch_simt_gemm(g, (cudaStream_t)stream); if (r) return r;
  }
  return 0;
}

int drb_cond_tables(drb_plan* p, void* stream) {
  if (!p) return DRB_E_INVALID;
  if (!p->condpre) return 0;
  if (!p->cond_ready && p->clip0 == 0) {   // (DRB_BRANCH_LEARNED never reads the audio clips' rows)
    if (!p->spec_ready) { set_error("drb_mel_forward has not been called"); return DRB_E_STATE; }
    int r = build_cond_tables(p, 0, stream); if (r) return r;
    p->cond_ready = true;
  }
  if (p->learned && !p->ucond_ready) {   // the learned clips change with the parameter, not with the audio clip
    if (!p->uspec_ready) { set_error("drb_plan_set_uncond_spec has not been called"); return DRB_E_STATE; }
    int r = build_cond_tables(p, p->cfg.batch, stream); if (r) return r;
    p->ucond_ready = true;
  }
  return 0;
}

namespace {
// spec32 clip layout [clip][T][Mp] (mel bins fastest, zero beyond n_mels) from the learned table [n_mels][ld], one copy per clip
__global__ void uncond_spec_kernel(const float* __restrict__ src, int ld, int n_mels, float* __restrict__ dst, int clips, int T, int Mp) {
  const size_t n = (size_t)clips * T * Mp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i % Mp), t = (int)((i / Mp) % T);
    dst[i] = m < n_mels ? __ldg(src + (size_t)m * ld + t) : 0.f;
  }
}
}  // namespace

int drb_plan_set_uncond_spec(drb_plan* p, const float* spec, int32_t ld, void* stream) {
  if (!p || !spec) { set_error("set_uncond_spec: null argument"); return DRB_E_INVALID; }
  const drb_config& c = p->cfg;
  if (p->lay.Bs != 2 * c.batch) { set_error("plan was not created with DRB_BRANCH_COND_LEARNED"); return DRB_E_INVALID; }
  if (ld < c.frames) { set_error("set_uncond_spec: the table holds %d frames, the plan needs %d", ld, c.frames); return DRB_E_INVALID; }
  float* dst = p->at<float>(p->lay.spec32) + (size_t)c.batch * c.frames * p->lay.Mp;
  const size_t n = (size_t)c.batch * c.frames * p->lay.Mp;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  uncond_spec_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(spec, ld, c.n_mels, dst, c.batch, c.frames, p->lay.Mp);
  DRB_CUDA(cudaGetLastError());
  ++g_launches;
  p->uspec_ready = true; p->ucond_ready = false;
  return 0;
}

int drb_plan_use_cond_tables(drb_plan* p, int32_t enable) {
  if (!p) return DRB_E_INVALID;
  p->cond_use = enable != 0;
  return 0;
}

int drb_in_proj(drb_plan* p, const float* x_t, int32_t t_index, void* stream) {
  if (!p || !x_t || t_index < 0 || t_index >= p->cfg.timesteps) { set_error("in_proj: bad argument"); return DRB_E_INVALID; }
  if (!p->tables_ready) { set_error("drb_time_tables has not been called"); return DRB_E_STATE; }
  cudaStream_t s = (cudaStream_t)stream;
  const int B = p->cfg.batch, T = p->cfg.frames, C = p->cfg.residual_channels, F = p->cfg.pitches;
  const bool tensor = p->cfg.precision != DRB_PREC_FP32;
  const int e0 = p->prof ? p->ev_mark(s) : -1;
  int r;
  if (tensor) {  // relu(input_projection(x_t)) for every branch copy + operand pair of x + d_0(t), one kernel
    r = launch_in_proj_fused(x_t, p->in_w, p->in_b, p->dvec(0, 0), p->steps, t_index, B * T, T, F, C, p->NB / B, p->xfmt(),
                             p->at<float>(p->lay.x32), p->ws + p->lay.xh, p->ws + p->lay.xl,
                             p->n4() ? p->at<uint8_t>(p->lay.xs) : nullptr, p->at<unsigned int>(p->lay.range), s);
  } else {
    SimtGemm g;  // relu(input_projection(x_t))   model/diffwave.py:667-668 ; x_t [B,1,T,88] is already [B*T][88]
    g.A = x_t; g.lda = F; g.T = T; g.Ck = F; g.W = p->in_w; g.ldw = F; g.bias = p->in_b; g.act = 1;
    g.C = p->at<float>(p->lay.x32); g.ldc = C; g.M = B * T; g.N = C;
    r = launch_simt_gemm(g, s); if (r) return r;
    r = launch_prep_xin(p->at<float>(p->lay.x32), nullptr, nullptr, nullptr, B * T, C, p->NB / B, 0, s);
  }
  if (p->prof && r == 0) p->ev_spans[2].push_back({e0, p->ev_mark(s)});
  return r;
}

int drb_resblock_forward(drb_plan* p, int32_t layer, int32_t t_index, void* stream) {
  if (!p || layer < 0 || layer >= p->cfg.residual_layers || t_index < 0 || t_index >= p->cfg.timesteps) {
    set_error("resblock: bad argument"); return DRB_E_INVALID;
  }
  if (!p->tables_ready) { set_error("drb_time_tables has not been called"); return DRB_E_STATE; }
  if (p->n_cond > 0 && p->clip0 == 0 && !p->spec_ready) { set_error("drb_mel_forward has not been called"); return DRB_E_STATE; }
  if (p->learned && !p->uspec_ready) { set_error("drb_plan_set_uncond_spec has not been called"); return DRB_E_STATE; }
  NvtxRange nv("drb.resblock l=%d t=%d", layer, t_index);
  cudaStream_t s = (cudaStream_t)stream;
  const drb_config& c = p->cfg;
  const int B = c.batch, T = c.frames, C = c.residual_channels, L = c.residual_layers, k = c.kernel_size, Mp = p->lay.Mp;
  const int NB = p->NB, nc = p->n_cond;
  const int first = layer == 0, do_res = layer < L - 1;
  int r;
  if (c.precision == DRB_PREC_FP32) {
    float* x32 = p->at<float>(p->lay.x32); float* y = p->at<float>(p->lay.ybuf); float* z = p->at<float>(p->lay.z32);
    SimtGemm g;  // dilated_conv(x + d)   model/diffwave.py:138-139
    g.lda = C; g.T = T; g.taps = k; g.dil = p->dil[layer]; g.Ck = C; g.addvec = p->dvec(layer, p->steps ? 0 : t_index);
    g.addvec_steps = p->steps; g.addvec_mod = B; g.addvec_stride = C;
    g.W = p->at<float>(p->lay.wd32) + (size_t)layer * 2 * C * k * C; g.ldw = k * C; g.ldc = 2 * C; g.N = 2 * C;
    if (nc > 0) {
      g.A = x32; g.C = y; g.M = nc * T; g.bias = p->bias_ptr(layer, 2);
      r = launch_simt_gemm(g, s); if (r) return r;
      SimtGemm q;  // + conditioner_projection(spec)   model/diffwave.py:143-144
      q.A = p->at<float>(p->lay.spec32) + (size_t)p->clip0 * T * Mp; q.lda = Mp; q.T = T; q.Ck = Mp;
      q.W = p->at<float>(p->lay.wc32) + (size_t)layer * 2 * C * Mp; q.ldw = Mp; q.accumulate = 1;
      q.C = y; q.ldc = 2 * C; q.M = nc * T; q.N = 2 * C;
      r = launch_simt_gemm(q, s); if (r) return r;
    }
    if (NB > nc) {
      g.A = x32 + (size_t)nc * T * C; g.C = y + (size_t)nc * T * 2 * C; g.M = (NB - nc) * T; g.bias = p->bias_ptr(layer, p->zero_spec ? 2 : 3);
      r = launch_simt_gemm(g, s); if (r) return r;
    }
    r = launch_gate(y, z, NB * T, C, s); if (r) return r;
    SimtGemm o;  // output_projection(z)   model/diffwave.py:149
    o.A = z; o.lda = C; o.T = T; o.Ck = C; o.W = p->wo32[layer]; o.ldw = C; o.bias = p->bo[layer];
    o.C = y; o.ldc = 2 * C; o.M = NB * T; o.N = 2 * C;
    r = launch_simt_gemm(o, s); if (r) return r;
    return launch_res_skip(y, x32, p->at<float>(p->lay.skip), NB * T, C, first, do_res, s);
  }
  UmmaGate ug;
  ug.NB = NB; ug.n_cond = nc; ug.T = T; ug.C = C; ug.taps = k; ug.dil = p->dil[layer]; ug.Mp = Mp;
  ug.pair = p->pair; ug.window = p->window; ug.persistent = p->persistent; ug.xwh = &p->win_h[layer]; ug.xwl = &p->win_l[layer]; ug.prec = p->prec(); ug.z_group0 = layer * p->lay.NBcap; ug.inv_scale = p->wscale(2 * layer) + 1;
  ug.bias_cond = p->bias_ptr(layer, 0); ug.bias_unc = p->bias_ptr(layer, p->zero_spec ? 0 : 1);
  // the f16n4 kernel has no conditioner K-slabs, and the learned clips exist only as table rows: there the per-clip tables are
  // always used (built on demand)
  const bool need_tables = p->n4() || p->learned;
  auto tables_ready = [&]() { return (p->clip0 > 0 || p->cond_ready) && (!p->learned || p->ucond_ready); };
  if (need_tables && nc > 0 && !tables_ready()) { r = drb_cond_tables(p, stream); if (r) return r; }
  ug.need_tables = p->learned ? 1 : 0;
  if (p->n4()) {
    ug.n4 = 1; ug.xw4 = &p->win4[layer]; ug.wd4 = &p->wd4[layer]; ug.wsf = &p->wsf4[layer]; ug.xs = p->at<uint8_t>(p->lay.xs);
  }
  if (p->condpre && tables_ready() && (p->cond_use || need_tables) && nc > 0) {
    ug.cond = p->at<float>(p->lay.cond) + ((size_t)layer * p->lay.Bs + p->clip0) * T * 2 * C;
    if (first && p->share0 && NB == 2 * B && nc == B) ug.dual_B = B;
  }
  const int e0 = p->prof ? p->ev_mark(s) : -1;
  r = launch_umma_gate(p->maps, p->layers[layer], ug, s); if (r) return r;
  const int e1 = p->prof ? p->ev_mark(s) : -1;
  if (do_res) {  // the last layer's residual half is dead; every layer's skip half is deferred to the head GEMM
    UmmaZGemm uz;
    uz.pair = p->pair; uz.persistent = p->persistent; uz.NB = NB; uz.T = T; uz.C = C; uz.prec = ug.prec; uz.mode = 0; uz.groups = 1; uz.z_group0 = ug.z_group0;
    uz.inv_scale = p->wscale(2 * layer + 1) + 1;
    uz.group_stride = p->lay.NBcap; uz.w_h = &p->layers[layer].wo_h; uz.w_l = &p->layers[layer].wo_l; uz.out32 = &p->maps.x32;
    uz.bias = p->bo[layer]; uz.dnext = p->dvec(layer + 1, 0); uz.t_uniform = t_index; uz.steps = p->steps; uz.bsamp = B;
    uz.range_max = p->at<unsigned int>(p->lay.range);
    if (p->n4()) { uz.x_n4 = 1; uz.xl4 = &p->xl4; uz.xs = p->at<uint8_t>(p->lay.xs); }
    r = launch_umma_zgemm(p->maps, uz, s); if (r) return r;
  }
  if (p->prof) {
    const int e2 = p->ev_mark(s);
    p->ev_spans[0].push_back({e0, e1}); p->ev_spans[1].push_back({e1, e2});
  }
  return 0;
}

int drb_head_posterior_step(drb_plan* p, const float* x_t, const float* noise, float* x_prev, float* net_out,
                            const drb_update* upd, void* stream) {
  if (!p || !x_prev || !upd) { set_error("head: null argument"); return DRB_E_INVALID; }
  if (upd->has_noise && !noise) { set_error("head: has_noise without a noise pointer"); return DRB_E_INVALID; }
  const bool needs_x = !(upd->mode == DRB_UPD_X0_FINAL || upd->mode == DRB_UPD_NONE);
  if (needs_x && !x_t) { set_error("head: update mode %d needs x_t", upd->mode); return DRB_E_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  const drb_config& c = p->cfg;
  const int B = c.batch, T = c.frames, C = c.residual_channels, F = c.pitches;
  float* h = p->at<float>(p->lay.hbuf);
  SimtGemm g;  // relu(skip_projection(skip / sqrt(L)))   model/diffwave.py:682-684
  g.A = p->at<float>(p->lay.skip); g.lda = C; g.T = T; g.Ck = C; g.a_div = sqrtf((float)c.residual_layers);
  g.W = p->skw; g.ldw = C; g.bias = p->skb; g.act = 1; g.C = h; g.ldc = C; g.M = p->NB * T; g.N = C;
  const int e0 = p->prof ? p->ev_mark(s) : -1;
  int r;
  if (c.precision == DRB_PREC_FP32) {
    r = launch_simt_gemm(g, s); if (r) return r;
  } else {  // one long-K tensor-core GEMM over the stored z of all layers (skip sum, 1/sqrt(L), skip_projection, ReLU)
    UmmaZGemm uz;
    uz.pair = p->pair; uz.persistent = p->persistent; uz.NB = p->NB; uz.T = T; uz.C = C; uz.prec = p->prec(); uz.mode = 1; uz.groups = c.residual_layers;
    uz.inv_scale = p->wscale(2 * c.residual_layers) + 1;
    uz.z_group0 = 0; uz.group_stride = p->lay.NBcap; uz.w_h = &p->maps.wcomp_h; uz.w_l = &p->maps.wcomp_l; uz.out32 = &p->maps.h32;
    uz.bias = p->at<float>(p->lay.bcomp); uz.dnext = nullptr;
    const bool tc = p->head_tc && (p->NB == B || (p->NB == 2 * B && p->pair));
    if (tc) { uz.hp_h = &p->maps.hh; uz.hp_l = &p->maps.hl; }
    r = launch_umma_zgemm(p->maps, uz, s); if (r) return r;
    if (tc) {   // output_projection + guidance combine + posterior update on the tensor cores (one N block of padded rows)
      UmmaZGemm uo;
      uo.pair = p->pair; uo.persistent = 0; uo.NB = p->NB; uo.T = T; uo.C = C; uo.prec = p->prec(); uo.mode = 3; uo.groups = 1;
      uo.inv_scale = p->wscale(p->wout_slot()) + 1; uo.z_group0 = 0; uo.group_stride = 0;
      uo.w_h = &p->maps.wout_h; uo.w_l = &p->maps.wout_l; uo.out32 = &p->maps.h32; uo.bias = p->hdb; uo.dnext = nullptr;
      uo.a_h = &p->maps.hh; uo.a_l = &p->maps.hl; uo.dual_B = p->NB == 2 * B ? B : 0; uo.F = F; uo.upd = upd;
      uo.x_t = x_t; uo.noise = noise; uo.x_prev = x_prev; uo.net_out = net_out;
      const int e1 = p->prof ? p->ev_mark(s) : -1;
      r = launch_umma_zgemm(p->maps, uo, s);
      if (p->prof && r == 0) { const int e2 = p->ev_mark(s); p->ev_spans[3].push_back({e0, e2}); p->ev_spans[4].push_back({e1, e2}); }
      return r;
    }
  }
  SimtGemm o;  // output_projection + guidance combine + posterior update
  o.A = h; o.lda = C; o.T = T; o.Ck = C; o.W = p->hdw; o.ldw = C; o.bias = p->hdb;
  if (p->NB == 2 * B) {  // (1+w)*x0_c - w*x0_u, applied to the (linear) head's input   task/diffusion.py:1009
    o.A2 = h + (size_t)B * T * C; o.alpha = 1.f + upd->w; o.beta = -upd->w;
  }
  o.C = x_prev; o.ldc = F; o.M = B * T; o.N = F; o.upd = upd; o.x_t = x_t; o.noise = noise; o.net_out = net_out;
  const int e1 = p->prof ? p->ev_mark(s) : -1;
  r = launch_simt_gemm(o, s);
  if (p->prof && r == 0) { const int e2 = p->ev_mark(s); p->ev_spans[3].push_back({e0, e2}); p->ev_spans[4].push_back({e1, e2}); }
  return r;
}

int drb_sample_step(drb_plan* p, const float* x_t, const float* noise, float* x_prev, int32_t t_index,
                    const drb_update* upd, void* stream) {
  NvtxRange nv("drb.sample_step t=%d", t_index);
  int r = drb_in_proj(p, x_t, t_index, stream); if (r) return r;
  for (int l = 0; l < p->cfg.residual_layers; ++l) { r = drb_resblock_forward(p, l, t_index, stream); if (r) return r; }
  return drb_head_posterior_step(p, x_t, noise, x_prev, nullptr, upd, stream);
}

int drb_sample_loop(drb_plan* p, float* x, const float* noise, const drb_update* updates_host, int32_t t_start,
                    int32_t t_stop, float* trajectory, void* stream) {
  if (!p || !x || !updates_host || t_start <= t_stop || t_stop < 0 || t_start > p->cfg.timesteps) {
    set_error("sample_loop: bad argument"); return DRB_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)p->cfg.batch * p->cfg.frames * p->cfg.pitches;
  size_t j = 0;
  if (p->n_cond > 0) { p->cond_use = true; int r = drb_cond_tables(p, stream); if (r) return r; }
  if (!trajectory) {
    for (int t = t_start - 1, i = 0; t >= t_stop; --t, ++i) {
      const drb_update* u = &updates_host[i];
      const float* nz = nullptr;
      if (u->has_noise) { if (!noise) { set_error("sample_loop: noise missing"); return DRB_E_INVALID; } nz = noise + (j++) * n; }
      int r = drb_sample_step(p, x, nz, x, t, u, stream); if (r) return r;
    }
    return 0;
  }
  // With a trajectory (the reference's per-step host copy, task/diffusion.py:530) the roll ping-pongs between the
  // caller's buffer and a plan-owned one, so the copy of step i (on a side stream) overlaps the compute of step i+1.
  if (!p->copy_stream) {
    DRB_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      DRB_CUDA(cudaEventCreateWithFlags(&p->ev_step[k], cudaEventDisableTiming));
      DRB_CUDA(cudaEventCreateWithFlags(&p->ev_copy[k], cudaEventDisableTiming));
    }
  }
  float* buf[2] = {x, p->at<float>(p->lay.xpp)};
  int i = 0;
  for (int t = t_start - 1; t >= t_stop; --t, ++i) {
    const drb_update* u = &updates_host[i];
    const float* nz = nullptr;
    if (u->has_noise) { if (!noise) { set_error("sample_loop: noise missing"); return DRB_E_INVALID; } nz = noise + (j++) * n; }
    float* src = buf[i & 1];
    float* dst = buf[(i + 1) & 1];
    if (i >= 2) DRB_CUDA(cudaStreamWaitEvent(s, p->ev_copy[i & 1], 0));  // dst was the source of the copy of step i-2
    int r = drb_sample_step(p, src, nz, dst, t, u, stream); if (r) return r;
    DRB_CUDA(cudaEventRecord(p->ev_step[i & 1], s));
    DRB_CUDA(cudaStreamWaitEvent(p->copy_stream, p->ev_step[i & 1], 0));
    DRB_CUDA(cudaMemcpyAsync(trajectory + (size_t)i * n, dst, n * sizeof(float), cudaMemcpyDefault, p->copy_stream));
    DRB_CUDA(cudaEventRecord(p->ev_copy[i & 1], p->copy_stream));
  }
  if (i & 1) DRB_CUDA(cudaMemcpyAsync(x, buf[1], n * sizeof(float), cudaMemcpyDeviceToDevice, s));  // result back in place
  DRB_CUDA(cudaStreamWaitEvent(s, p->ev_copy[0], 0));   // the caller's stream now also covers the trajectory copies
  if (i >= 2) DRB_CUDA(cudaStreamWaitEvent(s, p->ev_copy[1], 0));
  return 0;
}

int drb_plan_profile(drb_plan* p, int32_t enable) {
  if (!p) return DRB_E_INVALID;
  p->prof = enable != 0;
  p->ev_used = 0;
  for (auto& v : p->ev_spans) v.clear();
  return 0;
}

int drb_plan_profile_read(drb_plan* p, double* ms_total, int64_t* launches) {
  if (!p || !ms_total || !launches) return DRB_E_INVALID;
  DRB_CUDA(cudaDeviceSynchronize());
  for (int k = 0; k < 4; ++k) {
    double tot = 0.0;
    for (auto& sp : p->ev_spans[k]) {
      float ms = 0.f;
      DRB_CUDA(cudaEventElapsedTime(&ms, p->ev_pool[sp.first], p->ev_pool[sp.second]));
      tot += ms;
    }
    ms_total[k] = tot; launches[k] = (int64_t)p->ev_spans[k].size();
  }
  return 0;
}

int drb_plan_range_stats(drb_plan* p, float* max_abs, int32_t reset, void* stream) {
  if (!p || !max_abs) return DRB_E_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned int bits = 0;
  DRB_CUDA(cudaMemcpyAsync(&bits, p->ws + p->lay.range, sizeof(bits), cudaMemcpyDeviceToHost, s));
  if (reset) DRB_CUDA(cudaMemsetAsync(p->ws + p->lay.range, 0, sizeof(bits), s));
  DRB_CUDA(cudaStreamSynchronize(s));
  memcpy(max_abs, &bits, sizeof(bits));   // NaN operands read back as NaN, overflowed ones as inf or > 65504
  return 0;
}

int drb_plan_profile_read2(drb_plan* p, double* ms_total, int64_t* launches, int32_t n_classes, double* gate_ms_per_layer,
                           int32_t n_layers) {
  if (!p || !ms_total || !launches || n_classes < 1 || n_classes > 5) return DRB_E_INVALID;
  DRB_CUDA(cudaDeviceSynchronize());
  for (int k = 0; k < n_classes; ++k) {
    double tot = 0.0;
    for (auto& sp : p->ev_spans[k]) {
      float ms = 0.f;
      DRB_CUDA(cudaEventElapsedTime(&ms, p->ev_pool[sp.first], p->ev_pool[sp.second]));
      tot += ms;
    }
    ms_total[k] = tot; launches[k] = (int64_t)p->ev_spans[k].size();
  }
  if (gate_ms_per_layer && n_layers == p->cfg.residual_layers) {   // spans are recorded layer 0..L-1 every step
    for (int l = 0; l < n_layers; ++l) gate_ms_per_layer[l] = 0.0;
    for (size_t i = 0; i < p->ev_spans[0].size(); ++i) {
      float ms = 0.f;
      DRB_CUDA(cudaEventElapsedTime(&ms, p->ev_pool[p->ev_spans[0][i].first], p->ev_pool[p->ev_spans[0][i].second]));
      gate_ms_per_layer[i % n_layers] += ms;
    }
  }
  return 0;
}

int drb_plan_buffer(drb_plan* p, const char* name, void** ptr, size_t* bytes) {
  if (!p || !name || !ptr) return DRB_E_INVALID;
  const drb_config& c = p->cfg;
  const size_t rows = (size_t)p->lay.NBcap * c.frames, C = c.residual_channels;
  const bool tensor = c.precision != DRB_PREC_FP32;
  std::string n(name);
  size_t off = 0, sz = 0; bool ok = true;
  if (n == "x32") { off = p->lay.x32; sz = rows * C * 4; }
  else if (n == "skip") { off = p->lay.skip; sz = rows * C * 4; }
  else if (n == "h") {
    if (tensor && p->head_tc) { set_error("'h' is kept as an operand pair by the tensor-core head (DRB_NO_HEAD_TC=1 keeps fp32)"); return DRB_E_STATE; }
    off = p->lay.hbuf; sz = rows * C * 4;
  }
  else if (n == "dtab") { off = p->lay.dtab; sz = (size_t)c.residual_layers * c.timesteps * C * 4; }
  else if (n == "spec32") { off = p->lay.spec32; sz = (size_t)p->lay.Bs * c.frames * p->lay.Mp * 4; }
  else if (n == "y" && !tensor) { off = p->lay.ybuf; sz = rows * 2 * C * 4; }
  else if (n == "z32" && !tensor) { off = p->lay.z32; sz = rows * C * 4; }
  else if (n == "xh" && tensor) { off = p->lay.xh; sz = rows * C * 2; }
  else if (n == "xl" && tensor) { off = p->lay.xl; sz = rows * C * 2; }
  else if (n == "zh" && tensor) { off = p->lay.zh; sz = rows * C * 2; }
  else if (n == "zl" && tensor) { off = p->lay.zl; sz = rows * C * 2; }
  else if (n == "logmel") { *ptr = mel_logmel_ptr(p->mel, bytes); return 0; }
  else ok = false;
  if (!ok) { set_error("unknown buffer '%s'", name); return DRB_E_INVALID; }
  *ptr = p->ws + off; if (bytes) *bytes = sz;
  return 0;
}

}  // extern "C"

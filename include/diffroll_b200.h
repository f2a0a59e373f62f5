/* diffroll_b200 — C ABI of the B200-native DiffRoll sampling hot path.
 *
 * The reference (sony/DiffRoll) is pure Python and has no FFI of its own; the
 * boundary it exposes for this path is the PyTorch module surface
 *   ClassifierFreeDiffRoll.forward            model/diffwave.py:637-686
 *   SpecRollDiffusion.<sampler>(x, wav, t)    task/diffusion.py:804-1055
 *   SpecRollDiffusion.predict_step / sampling task/diffusion.py:513-534, 765-790
 * Each entry point below replaces the arithmetic of the cited reference lines.
 * The Python host side (diffroll_b200/model.py, task.py) keeps the reference's
 * names and argument meaning and calls these through ctypes.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *    its name ends in _host;
 *  - every call is asynchronous on the given stream (a cudaStream_t passed as
 *    void*), never allocates or frees caller memory, never throws;
 *  - return value 0 = ok, otherwise a negative DRB_E_* code or a positive
 *    cudaError_t; drb_last_error() gives the message for the calling thread;
 *  - a drb_plan owns repacked weights, TMA descriptors and the cuFFT plan; all
 *    of its device storage lives in the caller-provided workspace.
 */
#ifndef DIFFROLL_B200_H
#define DIFFROLL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRB_VERSION 100

#define DRB_E_INVALID   (-1)  /* bad argument / unsupported shape            */
#define DRB_E_WORKSPACE (-2)  /* workspace too small or misaligned           */
#define DRB_E_CUFFT     (-3)
#define DRB_E_DRIVER    (-4)  /* cuTensorMapEncodeTiled unavailable / failed */
#define DRB_E_STATE     (-5)  /* call order (e.g. step before time tables)   */

/* arithmetic of the contractions */
#define DRB_PREC_FP32    0  /* fp32 CUDA-core kernels (validation path)                              */
#define DRB_PREC_BF16X3  1  /* tcgen05, bf16 hi/lo split, 3 products, fp32 accumulate (parity mode) */
#define DRB_PREC_BF16    2  /* tcgen05, single bf16 product (fast mode, NOT within 1e-3)            */
#define DRB_PREC_F16F8   3  /* tcgen05, fp16 product + e4m3 correction product (2 MMA units, parity-grade) */
#define DRB_PREC_F16E5   4  /* tcgen05, fp16 product + e5m2 correction product into the same accumulator   */
#define DRB_PREC_F16N4   5  /* as F16E5, but the dilated conv's correction product is block-scaled fp4 (kind::mxf4nvf4,
                               e2m1 codes + one ue4m3 scale per 16 channels): 1.5 instead of 2 MMA units per K-step.
                               Needs an even number of 128-frame tiles (CTA pairs) and tap windows <= 192 frames. */

/* which network branches one step evaluates */
#define DRB_BRANCH_COND_UNCOND 0  /* classifier-free pair: (1+w)*cond - w*uncond   task/diffusion.py:1007-1009 */
#define DRB_BRANCH_COND        1  /* one conditional forward                       task/diffusion.py:839        */
#define DRB_BRANCH_UNCOND      2  /* spec == -1 (sampling=True)                    task/diffusion.py:979        */
#define DRB_BRANCH_COND_ZEROSPEC 3 /* pair: cond + cond on a zero waveform (its normalised spectrogram is all 0)
                                      task/diffusion.py:1038-1040 (cfdg_ddim_x0 omits sampling=True) */
#define DRB_BRANCH_COND_LEARNED 4  /* pair: cond + cond on the LEARNED unconditional spectrogram of condition='trainable_spec'
                                      (model/diffwave.py:600-605, 657-658: with sampling=True the spectrogram is the
                                      [n_mels, 641] parameter instead of -1).  The plan must be CREATED with this value
                                      (it holds `batch` extra clips and twice the conditioner tables) and given the table
                                      with drb_plan_set_uncond_spec; it may be switched to the other modes and back. */
#define DRB_BRANCH_LEARNED      5  /* one forward per roll, every roll conditioned on that learned spectrogram: a sampling=True
                                      forward of condition='trainable_spec' on its own (generation_ddpm_x0,
                                      task/diffusion.py:979).  Needs a plan created with 4 or 5; the tensor-core precisions
                                      need an even number of 128-frame tiles (CTA pairs), else the step fails and the caller
                                      runs mode 4 at guidance weight -1, which gives the same result. */

/* posterior-update formulas, evaluated in the reference's operation order with fp32 scalars s[0..4] */
#define DRB_UPD_X0        0  /* s0*net + s1*(x - s2*net)/s3 + s4*noise      task/diffusion.py:1018-1023 */
#define DRB_UPD_X0_FINAL  1  /* net / s0                                    task/diffusion.py:1016      */
#define DRB_UPD_EPS_DDPM  2  /* s0*(x - s1*net/s2) + s3*noise               task/diffusion.py:819-829   */
#define DRB_UPD_EPS_DDIM  3  /* s0*((x - s1*net)/s2) + s3*net + s4*noise    task/diffusion.py:886-889, 904-909 */
#define DRB_UPD_EPS_FINAL 4  /* (x - s0*net)/s1                             task/diffusion.py:885, 903  */
#define DRB_UPD_NONE      5  /* out = net (plain forward)                   model/diffwave.py:686       */

typedef struct drb_config {
  int32_t batch;             /* B: rolls in this plan                                        */
  int32_t frames;            /* T: roll frames (640)                                          */
  int32_t pitches;           /* 88                                                            */
  int32_t wave_len;          /* L: samples per clip (327680)                                  */
  int32_t residual_channels; /* C (512)                                                       */
  int32_t residual_layers;   /* (15)                                                          */
  int32_t kernel_size;       /* k (9), odd                                                    */
  int32_t dilation_base;     /* 2                                                             */
  int32_t dilation_bound;    /* 4 : dilation of layer i = base^(i % bound)  model/diffwave.py:624 */
  int32_t n_mels;            /* 229                                                           */
  int32_t n_fft;             /* 2048                                                          */
  int32_t hop_length;        /* 512                                                           */
  int32_t timesteps;         /* rows of the diffusion-embedding table                         */
  int32_t precision;         /* DRB_PREC_*                                                    */
  int32_t branches;          /* DRB_BRANCH_*                                                  */
  int32_t reserved;
} drb_config;

/* fp32 device pointers in the reference's state_dict layout. */
typedef struct drb_weights {
  const float* input_projection_w;      /* [C,88,1]   */
  const float* input_projection_b;      /* [C]        */
  const float* emb_projection1_w;       /* [512,128]  */
  const float* emb_projection1_b;       /* [512]      */
  const float* emb_projection2_w;       /* [512,512]  */
  const float* emb_projection2_b;       /* [512]      */
  const float* const* dilated_conv_w;   /* host array of L device ptrs, each [2C,C,k] */
  const float* const* dilated_conv_b;   /* [2C]       */
  const float* const* diffusion_projection_w; /* [C,512] */
  const float* const* diffusion_projection_b; /* [C]     */
  const float* const* conditioner_projection_w; /* [2C,n_mels,1] */
  const float* const* conditioner_projection_b; /* [2C]          */
  const float* const* output_projection_w;      /* [2C,C,1]      */
  const float* const* output_projection_b;      /* [2C]          */
  const float* skip_projection_w;       /* [C,C,1]    */
  const float* skip_projection_b;       /* [C]        */
  const float* head_projection_w;       /* [88,C,1]  (state_dict key output_projection.weight) */
  const float* head_projection_b;       /* [88]       */
  const float* stft_window;             /* [n_fft]            mel_layer.spectrogram.window */
  const float* mel_fb;                  /* [n_fft/2+1,n_mels] mel_layer.mel_scale.fb       */
} drb_weights;

typedef struct drb_update {
  int32_t mode;        /* DRB_UPD_*                           */
  int32_t has_noise;   /* 0: the noise pointer is not read    */
  float   s[5];
  float   w;           /* classifier-free guidance weight     */
} drb_update;

typedef struct drb_plan drb_plan;

int         drb_version(void);
const char* drb_last_error(void);

/* Bytes of device workspace a plan for cfg needs (0 on invalid cfg). */
size_t drb_plan_workspace_bytes(const drb_config* cfg);

/* Repack weights (tap-major, gate/filter interleave, bf16 hi/lo split), build TMA
 * descriptors and the cuFFT plan.  Replaces module construction + .to(device):
 * model/diffwave.py:580-635.  Synchronous on `stream` when it returns. */
int drb_plan_create(drb_plan** plan, const drb_config* cfg, const drb_weights* w,
                    void* workspace, size_t workspace_bytes, void* stream);
int drb_plan_destroy(drb_plan* plan);

/* Select which branches the following step calls evaluate (DRB_BRANCH_*).  A plan created with
 * DRB_BRANCH_COND_UNCOND has room for both and may be switched to either single branch. */
int drb_plan_set_branches(drb_plan* plan, int32_t branches);

/* Per-sample diffusion steps (ClassifierFreeDiffRoll.forward takes diffusion_step int64[B], model/diffwave.py:637,670;
 * the samplers pass one value repeated, the training / validation step a different one per roll,
 * task/diffusion.py:667).  steps_dev: device array of `batch` int32 values in [0, timesteps), read by the following
 * drb_in_proj / drb_resblock_forward / drb_sample_step calls INSTEAD of their t_index argument (it must stay valid
 * until they have run); NULL returns to the uniform t_index. */
int drb_plan_set_steps(drb_plan* plan, const int32_t* steps_dev);

/* DiffusionEmbedding MLP + every layer's diffusion_projection for all timesteps:
 * model/diffwave.py:65-74,138.  emb_table is the [timesteps,128] sinusoid table
 * built by the host with the reference's own expression (model/diffwave.py:83-88). */
int drb_time_tables(drb_plan* plan, const float* emb_table, void* stream);

/* STFT -> power -> HTK mel -> log -> per-clip min-max -> inpainting mask:
 * model/diffwave.py:643-654 + model/utils.py:21-32 + torchaudio MelSpectrogram.
 * waveform [B,L]; spec_out [B,n_mels,T] (may be NULL).  it0<it1 / if0<if1 select
 * the masked frame / mel ranges (set both ends to 0 for "no mask").  Also
 * precomputes nothing step-dependent; call once per clip. */
int drb_mel_forward(drb_plan* plan, const float* waveform, float* spec_out,
                    int32_t it0, int32_t it1, int32_t if0, int32_t if1, void* stream);

/* Optional, once per clip after drb_mel_forward: conditioner_projection_l(spec) of every layer (model/diffwave.py:143,
 * step-invariant) in fp32, added by the gate kernel's epilogue from then on instead of being contracted as extra
 * K-slabs every step.  Costs about as much as one network forward, so it pays off from the second step on the same
 * clip: drb_sample_loop always builds and uses the tables; a lone forward is cheaper without them.  Results with and
 * without differ by fp32-vs-operand-pair rounding of that term only (~1e-6); drb_plan_use_cond_tables(plan, 0) makes
 * the following steps ignore tables that happen to be ready, so a result never depends on call history. */
int drb_cond_tables(drb_plan* plan, void* stream);

/* DRB_BRANCH_COND_LEARNED only: the learned unconditional spectrogram `trainable_parameters` (model/diffwave.py:601-604),
 * fp32 device [n_mels][ld] with ld >= frames (the reference's table has 641 frames and is trimmed to the roll,
 * model/diffwave.py:30-39,662).  Copied into the plan; call again after the parameter changes.  Its conditioner
 * projections are (re)built by the next drb_cond_tables / step. */
int drb_plan_set_uncond_spec(drb_plan* plan, const float* spec, int32_t ld, void* stream);
int drb_plan_use_cond_tables(drb_plan* plan, int32_t enable);

/* input_projection + ReLU (model/diffwave.py:640,667-668) for timestep t_index. x_t [B,1,T,88]. */
int drb_in_proj(drb_plan* plan, const float* x_t, int32_t t_index, void* stream);

/* One ResidualBlock (model/diffwave.py:134-151) on every branch, plus the skip sum (:680). */
int drb_resblock_forward(drb_plan* plan, int32_t layer, int32_t t_index, void* stream);

/* skip/sqrt(L) -> skip_projection -> ReLU -> output_projection (model/diffwave.py:682-686),
 * guidance combine and posterior update (task/diffusion.py:1009-1023).
 * net_out (optional) receives the combined network output [B,1,T,88]. */
int drb_head_posterior_step(drb_plan* plan, const float* x_t, const float* noise, float* x_prev,
                            float* net_out, const drb_update* upd, void* stream);

/* drb_in_proj + L x drb_resblock_forward + drb_head_posterior_step. */
int drb_sample_step(drb_plan* plan, const float* x_t, const float* noise, float* x_prev,
                    int32_t t_index, const drb_update* upd, void* stream);

/* The loop of predict_step (task/diffusion.py:528-534) for t = t_start-1 .. t_stop:
 * x [B,1,T,88] is updated in place; noise [n,B,1,T,88] is consumed one slice per step
 * with has_noise; updates_host is a host array of (t_start-t_stop) drb_update, first
 * entry = highest t.  trajectory (optional, device or pinned-host-mapped) receives
 * every step's x, [t_start-t_stop,B,1,T,88]. */
int drb_sample_loop(drb_plan* plan, float* x, const float* noise, const drb_update* updates_host,
                    int32_t t_start, int32_t t_stop, float* trajectory, void* stream);

/* Number of kernels launched by this library on the calling thread since the last reset. */
int64_t drb_launch_count(int32_t reset);

/* Post-loop decode: extract_notes_wo_velocity(onsets, frames, ..., rule) with rule = 1
 * ('rule1': an onset also needs its frame activation, the sampling path's default) or 2 ('rule2': rising onset edges only),
 * task/utils.py:4-54, as called on every finished roll at task/diffusion.py:599-602.  onsets/frames [B,T,P] fp32 (may be
 * the same pointer); outputs per roll b: counts[b] notes in the reference's order (frame-major, then pitch),
 * pitches[b*max_notes + i], intervals[(b*max_notes + i)*2 + {0,1}] = (onset frame, offset frame).  Notes beyond
 * max_notes are counted but not stored (T*P/2 + P is always enough).  scratch: drb_extract_notes_scratch_bytes. */
size_t drb_extract_notes_scratch_bytes(int32_t B, int32_t T, int32_t P);
int drb_extract_notes(const float* onsets, const float* frames, int32_t B, int32_t T, int32_t P, float onset_threshold,
                      float frame_threshold, int32_t rule, void* scratch, int32_t* pitches, int32_t* intervals,
                      int32_t* counts, int32_t max_notes, void* stream);

/* Frame-level confusion counts behind test_step's precision_recall_fscore_support(label.flatten(),
 * pred.flatten() > threshold, average='binary'), task/diffusion.py:378-380: counts3 (device) = {TP, FP, FN} with
 * positive label == 1.0f.  precision = TP/(TP+FP), recall = TP/(TP+FN), f1 = 2PR/(P+R) are left to the host. */
int drb_frame_counts(const float* pred, const float* label, int64_t n, float threshold, uint64_t* counts3, void* stream);

/* Forward-only part of the reference's (validation) step around the network forward, SURVEY section 8 row f3
 * (task/diffusion.py:651-763).  steps: device int32[B], one diffusion step per roll; the two tables are the
 * [timesteps] fp32 schedule tensors of task/diffusion.py:250-251 on the device; n_per = elements per roll, % 4 == 0.
 *   drb_q_sample    x_t = sqrt_alphas_cumprod[t]*x_start + sqrt_one_minus_alphas_cumprod[t]*noise   task/diffusion.py:31-46
 *   drb_extract_x0  x0  = (x_t - sqrt_one_minus_alphas_cumprod[t]*epsilon)/sqrt_alphas_cumprod[t]   task/diffusion.py:49-65
 *   drb_p_losses    loss_type 0: mean|a-b| (F.l1_loss), 1: mean (a-b)^2, 2: smooth-l1 (beta 1)       task/diffusion.py:792-802
 *                   -> loss_out[0] (device); scratch: drb_p_losses_scratch_bytes() bytes
 *   drb_normalize_imagewise  per-roll min-max to [lo,hi], NaN (constant roll) -> lo                  model/utils.py:21-32 */
int drb_q_sample(const float* x_start, const float* noise, const int32_t* steps, const float* sqrt_alphas_cumprod,
                 const float* sqrt_one_minus_alphas_cumprod, float* x_t, int32_t B, int64_t n_per, void* stream);
int drb_extract_x0(const float* x_t, const float* epsilon, const int32_t* steps, const float* sqrt_alphas_cumprod,
                   const float* sqrt_one_minus_alphas_cumprod, float* x0, int32_t B, int64_t n_per, void* stream);
size_t drb_p_losses_scratch_bytes(void);
int drb_p_losses(const float* label, const float* prediction, int64_t n, int32_t loss_type, void* scratch,
                 float* loss_out, void* stream);
int drb_normalize_imagewise(const float* x, float* out, int32_t B, int64_t n_per, float lo, float hi, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training step, SURVEY section 8 row f3: forward with saved activations, backward to all 132 state_dict tensors, Adam.
 *   reference: SpecRollDiffusion.step / training_step   task/diffusion.py:258-270, 651-763   (loss: p_losses :792-802)
 *              ClassifierFreeDiffRoll.forward, training mode (spec dropout)   model/diffwave.py:637-699
 *              configure_optimizers = torch.optim.Adam(lr)                    task/diffusion.py:1057-1059
 * fp32 CUDA-core arithmetic (the reference trains in fp32).  The parameter / gradient structs hold DEVICE pointers to
 * tensors in the reference's own state_dict layout (conv weights [out][in][k], linear weights [out][in]); the arrays
 * wd ... bo are HOST arrays of residual_layers device pointers.  Gradients are accumulated with fp32 atomics (the
 * summation order, hence the last bits, can differ from run to run). */
typedef struct drb_train_config {
  int32_t batch, frames, pitches, residual_channels, residual_layers, kernel_size, dilation_base, dilation_bound, n_mels,
          timesteps;
} drb_train_config;

typedef struct drb_train_params {
  float *in_w, *in_b;                 /* input_projection            [C][88][1], [C]            model/diffwave.py:597  */
  float *e1w, *e1b, *e2w, *e2b;       /* diffusion_embedding.projection1 [512][128], projection2 [512][512]  :62-63 */
  float *skw, *skb;                   /* skip_projection             [C][C][1], [C]                          :628   */
  float *hdw, *hdb;                   /* output_projection           [88][C][1], [88]                        :629   */
  float *const *wd, *const *bd;       /* residual_layers.l.dilated_conv          [2C][C][k], [2C]            :112   */
  float *const *wdp, *const *bdp;     /* residual_layers.l.diffusion_projection  [C][512], [C]               :118   */
  float *const *wc, *const *bc;       /* residual_layers.l.conditioner_projection [2C][n_mels][1], [2C]      :120   */
  float *const *wo, *const *bo;       /* residual_layers.l.output_projection     [2C][C][1], [2C]            :124   */
} drb_train_params;

typedef struct drb_train drb_train;

size_t drb_train_workspace_bytes(const drb_train_config* cfg);
int drb_train_create(drb_train** out, const drb_train_config* cfg, void* workspace, size_t workspace_bytes, void* stream);
void drb_train_destroy(drb_train* plan);
/* x_t [B][T][88]; spec [B][n_mels][T] exactly as the network sees it (normalised log-mel, spec-dropout rolls and masks
 * already -1); steps int32[B]; emb_table [timesteps][128] (DiffusionEmbedding._build_embedding); pred [B][T][88]. */
int drb_train_forward(drb_train* plan, const drb_train_params* params, const float* x_t, const float* spec,
                      const int32_t* steps, const float* emb_table, float* pred, void* stream);
/* Optional: make the following drb_train_backward calls also leave d loss / d spec in g_spec ([B][n_mels][T] fp32 device,
 * overwritten by every backward; NULL = off, the default).  condition='trainable_spec' conditions the dropped rolls on the
 * parameter `trainable_parameters` (model/diffwave.py:695-699): its gradient is this tensor summed over those rolls. */
int drb_train_set_spec_grad(drb_train* plan, float* g_spec);
/* Backward of the last drb_train_forward: g_pred = d loss / d pred [B][T][88]; every tensor of grads receives its gradient
 * (accumulate != 0: added to its content); g_x_t (optional) = d loss / d x_t. */
int drb_train_backward(drb_train* plan, const drb_train_params* params, const drb_train_params* grads, const float* x_t,
                       const float* g_pred, int32_t accumulate, float* g_x_t, void* stream);
/* d p_losses / d prediction (loss_type 0 l1, 1 l2, 2 huber; mean over n), times roll_scale[i / per_roll] when given
 * (training mode 'ex_0': d extract_x0 / d epsilon = -sqrt_one_minus_alphas_cumprod[t] / sqrt_alphas_cumprod[t]). */
int drb_loss_grad(const float* label, const float* pred, float* g_pred, size_t n, size_t per_roll, int32_t loss_type,
                  const float* roll_scale, void* stream);
/* One torch.optim.Adam update of one tensor (amsgrad off); step = 1 for the first update. */
int drb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int32_t step, void* stream);
/* The same update for a list of tensors in one launch (torch's foreach / fused Adam, task/diffusion.py:1057-1059 steps all 130
 * tensors together).  table: device array of n_tensors x 5 uint64 {param, grad, exp_avg, exp_avg_sq, numel}; block_map: device array
 * of n_blocks x 2 int32 {tensor index, chunk index}, one entry per 4096-element chunk of every tensor. */
int drb_adam_step_multi(const void* table, const void* block_map, int32_t n_blocks, float lr, float beta1, float beta2,
                        float eps, float weight_decay, int32_t step, void* stream);

/* Measurement hooks (bench.py): with profiling enabled every step records CUDA events on the launching stream
 * around each kernel class; drb_plan_profile_read synchronises and returns, for class k in
 * {0: gate kernel, 1: out kernel, 2: in_proj, 3: head}, the summed milliseconds and the number of spans. */
int drb_plan_profile(drb_plan* plan, int32_t enable);
int drb_plan_profile_read(drb_plan* plan, double* ms_total4, int64_t* launches4);
/* Same with n_classes <= 5 classes (4: the head's output projection + guidance + posterior kernel alone, also part of
 * class 3) and, when gate_ms_per_layer != NULL and n_layers == residual_layers, the gate-kernel milliseconds summed per
 * layer (layer 0 is the launch shared by both guidance branches). */
int drb_plan_profile_read2(drb_plan* plan, double* ms_total, int64_t* launches, int32_t n_classes,
                           double* gate_ms_per_layer, int32_t n_layers);

/* Fractional diffusion steps, DiffusionEmbedding._lerp_embedding (model/diffwave.py:76-81) as reached from forward() with a
 * floating-point diffusion_step: emb_rows = device [batch][128], the caller's interpolated sinusoid rows, one per roll.  The plan
 * runs the embedding MLP and every diffusion_projection on them and uses the result (per roll) for all following calls until
 * this is called again with emb_rows == NULL.  Needs batch <= timesteps. */
int drb_plan_set_step_embeddings(drb_plan* plan, const float* emb_rows, void* stream);

/* The DRB_PREC_* a plan actually computes in: DRB_PREC_F16N4 is granted only when every launch can run as CTA pairs
 * (batch * ceil(frames / 128) even) with tap windows <= 192 frames; otherwise the plan runs as DRB_PREC_F16E5. */
int drb_plan_precision(const drb_plan* plan);

/* Range guard of the fp16-based operand formats (f16e5, f16f8).  Their main product rounds activations to fp16, which
 * overflows at 65504; weights are pre-scaled per tensor by a power of two, activations are not.  Every kernel that
 * emits an activation operand (in_proj, the residual update) folds max |value| into one device word; this call copies
 * it to *max_abs (synchronises the stream; NaN operands read back as NaN) and optionally resets it.  The Python module
 * re-runs a call in bf16x3 (fp32 exponent range) when the value is not finite or above 3e4.  No reference counterpart:
 * the reference computes in fp32 (model/diffwave.py:134-151). */
int drb_plan_range_stats(drb_plan* plan, float* max_abs, int32_t reset, void* stream);

/* Debug / test access to plan-owned device buffers ("x32","skip","xh","xl","zh","zl","spec_h","dtab","logmel"). */
int drb_plan_buffer(drb_plan* plan, const char* name, void** ptr, size_t* bytes);

#ifdef __cplusplus
}
#endif
#endif /* DIFFROLL_B200_H */

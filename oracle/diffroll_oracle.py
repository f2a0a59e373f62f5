"""CPU oracle: a plain-torch restatement of DiffRoll's sampling hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
file, and only as the checker or the timed CPU baseline.  The product path
(``diffroll_b200``) never imports it and has no CPU fallback.

Parity status: PINNED.  The reference has no tests or golden vectors of its own
(SURVEY.md §4), so this restatement is pinned against the reference itself:
``oracle/make_golden.py`` imports the unmodified reference in the build
container (via ``oracle/ref_shim.py``), runs it on the seeded inputs of
``diffroll_b200/synthetic.py`` and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those vectors, and
``tests/test_oracle_vs_reference.py`` (container only) against the live reference.

The restatement follows the reference op for op, *including* its redundancies
(mel recomputed twice per step, conditioner projection recomputed every step),
so that timing it is an honest stand-in for the reference's CPU path:

  linear_beta_schedule / schedule tables   task/diffusion.py:28-29, 239-256
  mel front-end (torchaudio, third-party,
     pinned 0.11.0, requirements.txt:12)   model/diffwave.py:635,643-645
  Normalization('imagewise')               model/utils.py:21-32
  DiffusionEmbedding                       model/diffwave.py:58-88
  ResidualBlock.forward                    model/diffwave.py:134-151
  ClassifierFreeDiffRoll.forward           model/diffwave.py:637-686
  samplers                                 task/diffusion.py:804-1055
  predict_step loop                        task/diffusion.py:513-534
"""
from __future__ import annotations

from math import sqrt

import torch
import torch.nn.functional as F


class Schedule:
    """task/diffusion.py:239-256 — all fp32 CPU tensors, built with the same torch expressions."""

    def __init__(self, beta_start, beta_end, timesteps):
        self.betas = torch.linspace(beta_start, beta_end, timesteps)
        alphas = 1.0 - self.betas
        alphas_cumprod = torch.cumprod(alphas, axis=0)
        alphas_cumprod_prev = F.pad(alphas_cumprod[:-1], (1, 0), value=1.0)
        self.sqrt_recip_alphas = torch.sqrt(1.0 / alphas)
        self.sqrt_alphas_cumprod = torch.sqrt(alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - alphas_cumprod)
        self.posterior_variance = self.betas * (1.0 - alphas_cumprod_prev) / (1 - alphas_cumprod)
        self.alphas = alphas
        self.timesteps = timesteps


def build_embedding(max_steps):
    """model/diffwave.py:83-88."""
    steps = torch.arange(max_steps).unsqueeze(1)
    dims = torch.arange(64).unsqueeze(0)
    table = steps * 10.0 ** (dims * 4.0 / 63.0)
    return torch.cat([torch.sin(table), torch.cos(table)], dim=1)


def silu(x):
    return x * torch.sigmoid(x)  # model/diffwave.py:53-55


def mel_spectrogram(waveform, window, fb, n_fft=2048, hop_length=512):
    """torchaudio MelSpectrogram(power=2, normalized=True, center=True, reflect) restated:
    torchaudio/functional/functional.py `spectrogram` + transforms `MelScale.forward`."""
    spec_f = torch.stft(waveform, n_fft=n_fft, hop_length=hop_length, win_length=n_fft,
                        window=window, center=True, pad_mode="reflect", normalized=False,
                        onesided=True, return_complex=True)
    spec_f = spec_f / window.pow(2.0).sum().sqrt()
    power = spec_f.abs().pow(2.0)                                  # [B, 1025, frames]
    return torch.matmul(power.transpose(-1, -2), fb).transpose(-1, -2)  # [B, n_mels, frames]


def normalize_imagewise(x, lo=0.0, hi=1.0):
    """model/utils.py:21-32."""
    x_max = x.flatten(1).max(1, keepdim=True)[0].unsqueeze(1)
    x_min = x.flatten(1).min(1, keepdim=True)[0].unsqueeze(1)
    x_std = (x - x_min) / (x_max - x_min)
    x_scaled = x_std * (hi - lo) + lo
    x_scaled[torch.isnan(x_scaled)] = lo
    return x_scaled


class OracleDiffRoll:
    """Functional restatement of ClassifierFreeDiffRoll + SpecRollDiffusion, condition='fixed' or 'trainable_spec' (the learned
    unconditional spectrogram of model/diffwave.py:600-605, 657-658, 695-699)."""

    def __init__(self, hp, state_dict, dtype=torch.float32, device=None):
        # device: None/CPU for the checker and the CPU baseline; "cuda" runs the same torch ops eagerly on the GPU
        # (cuDNN/cuBLAS), which is what the reference itself does on a GPU box (tests/test_gpu_eager_baseline.py)
        self.hp = hp
        self.dtype = dtype
        self.sd = {k: v.to(device=device, dtype=dtype) for k, v in state_dict.items()}
        self.sched = Schedule(hp["beta_start"], hp["beta_end"], hp["timesteps"])
        self.embedding = build_embedding(hp["timesteps"]).to(device=device, dtype=dtype)
        self.L = hp["residual_layers"]
        self.k = hp["kernel_size"]
        self.dilations = [hp["dilation_base"] ** (i % hp["dilation_bound"]) for i in range(self.L)]
        self.w = hp["sampling"].get("w", 0.0) if isinstance(hp["sampling"], dict) else hp["sampling"].w
        self.reverse_diffusion = getattr(self, hp["sampling"]["type"])

    # ---- network -------------------------------------------------------------------
    def spec_frontend(self, waveform):
        sd = self.sd
        sa = self.hp["spec_args"]
        spec = mel_spectrogram(waveform.to(self.dtype), sd["mel_layer.spectrogram.window"],
                               sd["mel_layer.mel_scale.fb"], sa["n_fft"], sa["hop_length"])
        spec = torch.log(spec + 1e-6)
        return normalize_imagewise(spec, 0.0, 1.0)

    def forward(self, x_t, waveform, diffusion_step, sampling=False, inpainting_t=None, inpainting_f=None, dropout_mask=None):
        """model/diffwave.py:637-686.  dropout_mask ([B], nonzero = drop): the training-mode spec dropout of :646-647 /
        fixed_dropout :689-693 with the Bernoulli draw injected."""
        sd = self.sd
        x_t = x_t.to(self.dtype).squeeze(1).transpose(1, 2)
        spec = self.spec_frontend(waveform)
        learned = self.hp.get("condition", "fixed") == "trainable_spec"
        if dropout_mask is not None:
            if learned:                                                # trainable_dropout, model/diffwave.py:695-699
                spec[dropout_mask.to(spec.device).bool()] = sd["trainable_parameters"]
            else:
                spec[dropout_mask.to(spec.device).bool()] = -1
        if inpainting_t and inpainting_f is None:
            spec[:, :, int(inpainting_t[0]):int(inpainting_t[1])] = -1
        elif inpainting_t is None and inpainting_f:
            spec[:, int(inpainting_f[0]):int(inpainting_f[1]), :] = -1
        elif inpainting_t and inpainting_f:
            spec[:, int(inpainting_f[0]):int(inpainting_f[1]), int(inpainting_t[0]):int(inpainting_t[1])] = -1
        if sampling is True:
            # :657-660; the learned table is 2-D [n_mels, 641]: conv1d treats it as one unbatched clip and the block's sum broadcasts it
            spec = sd["trainable_parameters"] if learned else torch.full_like(spec, -1)
        T_min = min(x_t.shape[-1], spec.shape[-1])
        x_t = x_t[..., :T_min]
        spectrogram = spec[..., :T_min]

        x = F.relu(F.conv1d(x_t, sd["input_projection.weight"], sd["input_projection.bias"]))
        if diffusion_step.dtype in (torch.int32, torch.int64):
            e = self.embedding[diffusion_step]
        else:                                                          # _lerp_embedding, model/diffwave.py:76-81
            tt = diffusion_step.to(self.embedding.device)
            lo_i, hi_i = torch.floor(tt).long(), torch.ceil(tt).long()
            lo_e, hi_e = self.embedding[lo_i], self.embedding[hi_i]
            e = lo_e + (hi_e - lo_e) * (tt - lo_i).to(self.dtype).unsqueeze(-1)
        e = silu(F.linear(e, sd["diffusion_embedding.projection1.weight"], sd["diffusion_embedding.projection1.bias"]))
        e = silu(F.linear(e, sd["diffusion_embedding.projection2.weight"], sd["diffusion_embedding.projection2.bias"]))

        skip = None
        for i in range(self.L):
            p = f"residual_layers.{i}."
            d = self.dilations[i]
            dstep = F.linear(e, sd[p + "diffusion_projection.weight"], sd[p + "diffusion_projection.bias"]).unsqueeze(-1)
            y = x + dstep
            pad = ((self.k - 1) * (d - 1) + self.k - 1) // 2
            y = F.conv1d(y, sd[p + "dilated_conv.weight"], sd[p + "dilated_conv.bias"], padding=pad, dilation=d)
            y = y + F.conv1d(spectrogram, sd[p + "conditioner_projection.weight"], sd[p + "conditioner_projection.bias"])
            gate, filt = torch.chunk(y, 2, dim=1)
            y = torch.sigmoid(gate) * torch.tanh(filt)
            y = F.conv1d(y, sd[p + "output_projection.weight"], sd[p + "output_projection.bias"])
            residual, s = torch.chunk(y, 2, dim=1)
            x = (x + residual) / sqrt(2.0)
            skip = s if skip is None else s + skip
        x = skip / sqrt(self.L)
        x = F.relu(F.conv1d(x, sd["skip_projection.weight"], sd["skip_projection.bias"]))
        x = F.conv1d(x, sd["output_projection.weight"], sd["output_projection.bias"])
        return x.transpose(1, 2).unsqueeze(1), spectrogram

    __call__ = forward

    # ---- samplers (noise injected instead of randn_like; SURVEY.md §7 hard part 5) -----
    def _t(self, x, t_index):
        return torch.tensor(t_index).repeat(x.shape[0])

    def _x0_update(self, x, x0_pred, t_index, noise, ddim=False):
        """task/diffusion.py:1013-1023 (shared verbatim by :836-851, :957-967, :985-995, :860-871, :1043-1053)."""
        s = self.sched
        if t_index == 0:
            return x0_pred / s.sqrt_alphas_cumprod[t_index]
        if ddim:
            sigma = 0
            return (s.sqrt_alphas_cumprod[t_index - 1]) * x0_pred + (
                torch.sqrt(1 - s.sqrt_alphas_cumprod[t_index - 1] ** 2 - sigma ** 2) * (
                    x - s.sqrt_alphas_cumprod[t_index] * x0_pred) / s.sqrt_one_minus_alphas_cumprod[t_index]) + (
                sigma * noise)
        sigma = (s.sqrt_one_minus_alphas_cumprod[t_index - 1] / s.sqrt_one_minus_alphas_cumprod[t_index]) * (
            torch.sqrt(1 - s.alphas[t_index]))
        return (s.sqrt_alphas_cumprod[t_index - 1]) * x0_pred + (
            torch.sqrt(1 - s.sqrt_alphas_cumprod[t_index - 1] ** 2 - sigma ** 2) * (
                x - s.sqrt_alphas_cumprod[t_index] * x0_pred) / s.sqrt_one_minus_alphas_cumprod[t_index]) + (
            sigma * noise)

    def inpainting_ddpm_x0(self, x, waveform, t_index, noise=None):      # task/diffusion.py:999-1025
        t = self._t(x, t_index)
        x0_c, spec = self(x, waveform, t, inpainting_t=self.hp["inpainting_t"], inpainting_f=self.hp["inpainting_f"])
        x0_0, _ = self(x, torch.zeros_like(waveform), t, sampling=True)
        x0 = (1 + self.w) * x0_c - self.w * x0_0
        return self._x0_update(x.to(self.dtype), x0, t_index, noise), spec

    def cfdg_ddpm_x0(self, x, waveform, t_index, noise=None):            # task/diffusion.py:943-969
        t = self._t(x, t_index)
        x0_c, spec = self(x, waveform, t)
        x0_0, _ = self(x, torch.zeros_like(waveform), t, sampling=True)
        x0 = (1 + self.w) * x0_c - self.w * x0_0
        return self._x0_update(x.to(self.dtype), x0, t_index, noise), spec

    def generation_ddpm_x0(self, x, waveform, t_index, noise=None):      # task/diffusion.py:971-997
        t = self._t(x, t_index)
        x0, spec = self(x, torch.zeros_like(waveform), t, sampling=True)
        return self._x0_update(x.to(self.dtype), x0, t_index, noise), spec

    def ddpm_x0(self, x, waveform, t_index, noise=None):                 # task/diffusion.py:831-853
        x0, spec = self(x, waveform, self._t(x, t_index))
        return self._x0_update(x.to(self.dtype), x0, t_index, noise), spec

    def ddim_x0(self, x, waveform, t_index, noise=None):                 # task/diffusion.py:855-875
        x0, spec = self(x, waveform, self._t(x, t_index))
        return self._x0_update(x.to(self.dtype), x0, t_index, noise, ddim=True), spec

    def cfdg_ddim_x0(self, x, waveform, t_index, noise=None):            # task/diffusion.py:1027-1055
        t = self._t(x, t_index)
        x0_c, spec = self(x, waveform, t)
        x0_0, _ = self(x, torch.zeros_like(waveform), t)                 # NB: no sampling=True here (:1039)
        x0 = (1 + self.w) * x0_c - self.w * x0_0
        return self._x0_update(x.to(self.dtype), x0, t_index, noise, ddim=True), spec

    def ddpm(self, x, waveform, t_index, noise=None):                    # task/diffusion.py:804-829
        s = self.sched
        eps, spec = self(x, waveform, self._t(x, t_index))
        x = x.to(self.dtype)
        mean = s.sqrt_recip_alphas[t_index] * (x - s.betas[t_index] * eps / s.sqrt_one_minus_alphas_cumprod[t_index])
        if t_index == 0:
            return mean, spec
        return mean + torch.sqrt(s.posterior_variance[t_index]) * noise, spec

    def ddim(self, x, waveform, t_index, noise=None):                    # task/diffusion.py:877-892
        s = self.sched
        eps, spec = self(x, waveform, self._t(x, t_index))
        x = x.to(self.dtype)
        x0 = (x - s.sqrt_one_minus_alphas_cumprod[t_index] * eps) / s.sqrt_alphas_cumprod[t_index]
        if t_index == 0:
            return x0, spec
        return s.sqrt_alphas_cumprod[t_index - 1] * x0 + s.sqrt_one_minus_alphas_cumprod[t_index - 1] * eps, spec

    def ddim2ddpm(self, x, waveform, t_index, noise=None):               # task/diffusion.py:894-911
        s = self.sched
        eps, spec = self(x, waveform, self._t(x, t_index))
        x = x.to(self.dtype)
        x0 = (x - s.sqrt_one_minus_alphas_cumprod[t_index] * eps) / s.sqrt_alphas_cumprod[t_index]
        if t_index == 0:
            return x0, spec
        sigma = (s.sqrt_one_minus_alphas_cumprod[t_index - 1] / s.sqrt_one_minus_alphas_cumprod[t_index]) * (
            torch.sqrt(1 - s.alphas[t_index]))
        return (s.sqrt_alphas_cumprod[t_index - 1] * x0
                + torch.sqrt(1 - s.sqrt_alphas_cumprod[t_index - 1] ** 2 - sigma ** 2) * eps + sigma * noise), spec

    # ---- forward-only (validation) step: task/diffusion.py:651-763 as run by validation_step :271-276 -------------
    @torch.no_grad()
    def step(self, batch, t, noise):
        """t [B] and noise [B,1,T,88] are injected instead of torch.randint (:667) / torch.randn_like (:670)."""
        hp, s = self.hp, self.sched
        two = isinstance(batch, list)
        first = batch[0] if two else batch
        lo, hi = hp["norm_args"][0], hp["norm_args"][1]
        roll = normalize_imagewise(first["frame"].to(self.dtype), lo, hi).unsqueeze(1)
        waveform = first["audio"]

        def q_sample(x_start):                                        # task/diffusion.py:31-46
            a = s.sqrt_alphas_cumprod[t][:, None, None, None].to(x_start.device)
            b = s.sqrt_one_minus_alphas_cumprod[t][:, None, None, None].to(x_start.device)
            return a * x_start + b * noise

        def extract_x0(x_t, eps):                                     # task/diffusion.py:49-65
            a = s.sqrt_alphas_cumprod[t][:, None, None, None].to(x_t.device)
            b = s.sqrt_one_minus_alphas_cumprod[t][:, None, None, None].to(x_t.device)
            return (x_t - b * eps) / a

        def p_losses(label, prediction):                              # task/diffusion.py:792-802
            fn = {"l1": F.l1_loss, "l2": F.mse_loss, "huber": F.smooth_l1_loss}[hp["loss_type"]]
            return fn(label, prediction)

        x_t = q_sample(roll)
        mode = hp["training"]["mode"]
        losses, tensors = {}, {}
        if mode == "epsilon":
            eps, spec = self(x_t, waveform, t)
            losses["diffusion_loss"] = p_losses(noise, eps)
            pred_roll = extract_x0(x_t, eps)
        elif mode == "x_0":
            pred_roll, spec = self(x_t, waveform, t)
            losses["diffusion_loss"] = p_losses(roll, pred_roll)
            if two:
                roll2 = normalize_imagewise(batch[1]["frame"].to(self.dtype), lo, hi).unsqueeze(1)
                pred_roll2, spec2 = self(q_sample(roll2), batch[1]["audio"], t, sampling=True)
                losses["unconditional_diffusion_loss"] = p_losses(roll2, pred_roll2)
                tensors.update(spec2=spec2, label_roll2=roll2, pred_roll2=pred_roll2)
        elif mode == "ex_0":
            eps, spec = self(x_t, waveform, t)
            pred_roll = extract_x0(x_t, eps)
            losses["diffusion_loss"] = p_losses(roll, pred_roll)
        else:
            raise ValueError(mode)
        tensors.update(pred_roll=pred_roll, label_roll=roll, spec=spec)
        return losses, tensors

    # ---- training step with autograd: task/diffusion.py:258-270 over step :651-763 in training mode ---------------
    def train_step(self, batch, t, noise, dropout_mask=None, want_input_grad=False):
        """Total loss over hp['loss_keys'] and its gradient with respect to every parameter tensor (torch autograd over the
        restated forward), with t / noise / the spec-dropout mask injected.  Returns (losses, grads: name -> tensor, g_x_t)."""
        hp, s = self.hp, self.sched
        two = isinstance(batch, list)
        first = batch[0] if two else batch
        lo, hi = hp["norm_args"][0], hp["norm_args"][1]
        with torch.no_grad():
            roll = normalize_imagewise(first["frame"].to(self.dtype), lo, hi).unsqueeze(1)
            a = s.sqrt_alphas_cumprod[t][:, None, None, None].to(device=roll.device, dtype=self.dtype)
            b = s.sqrt_one_minus_alphas_cumprod[t][:, None, None, None].to(device=roll.device, dtype=self.dtype)
            nz = noise.to(self.dtype)
            x_t = a * roll + b * nz
        names = [k for k in self.sd if not k.startswith("mel_layer.")]
        saved = {k: self.sd[k] for k in names}
        leaves = {k: saved[k].detach().clone().requires_grad_(True) for k in names}
        self.sd.update(leaves)
        x_in = x_t.clone().requires_grad_(want_input_grad)
        fn = {"l1": F.l1_loss, "l2": F.mse_loss, "huber": F.smooth_l1_loss}[hp["loss_type"]]
        try:
            with torch.enable_grad():
                mode = hp["training"]["mode"]
                losses = {}
                net, _ = self(x_in, first["audio"], t, dropout_mask=dropout_mask)
                if mode == "epsilon":
                    losses["diffusion_loss"] = fn(nz, net)
                elif mode == "x_0":
                    losses["diffusion_loss"] = fn(roll, net)
                    if two:
                        roll2 = normalize_imagewise(batch[1]["frame"].to(self.dtype), lo, hi).unsqueeze(1)
                        net2, _ = self(a * roll2 + b * nz, batch[1]["audio"], t, sampling=True)
                        losses["unconditional_diffusion_loss"] = fn(roll2, net2)
                elif mode == "ex_0":
                    losses["diffusion_loss"] = fn(roll, (x_in - b * net) / a)
                else:
                    raise ValueError(mode)
                total = sum(losses[k] for k in hp["loss_keys"])
                total.backward()
        finally:
            self.sd.update(saved)
        grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
        return {k: v.detach() for k, v in losses.items()}, grads, (x_in.grad if want_input_grad else None)

    # ---- the loop (task/diffusion.py:513-534), noise[i] used at the i-th step with t>0 -------
    @torch.no_grad()
    def sample_loop(self, x_T, waveform, noise, t_start=None, t_stop=0, keep_numpy=True):
        x = x_T
        spec = None
        T = self.hp["timesteps"] if t_start is None else t_start
        i = 0
        trajectory = []
        for t_index in reversed(range(t_stop, T)):
            n = None
            if t_index > 0:
                n = noise[i].to(self.dtype); i += 1
            x, spec = self.reverse_diffusion(x, waveform, t_index, noise=n)
            if keep_numpy:
                trajectory.append(x.detach().cpu().numpy())  # the per-step host copy of :530
        return x, spec, trajectory

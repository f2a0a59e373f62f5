"""Host-side mirror of the reference's task layer for the sampling path.

``SpecRollDiffusion`` keeps the names, argument meaning and error behaviour of
task/diffusion.py:219-256 (schedule), :513-534 (``predict_step``), :765-790
(``sampling``) and :804-1055 (the nine ``reverse_diffusion`` samplers).  The
arithmetic is not here: every sampler reduces to "which branches does the
network evaluate" plus five fp32 scalars for the fused posterior-update
epilogue, and hands both to the CUDA engine (diffroll_b200/engine.py).

Lightning is not required: the class is an ``nn.Module`` that offers the bits of
``LightningModule`` the reference scripts touch (``hparams``,
``load_from_checkpoint``, ``predict_step``/``test``-style entry points, ``log``).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import DrbUpdate


class AttributeDict(dict):
    """Attribute-access dict (stand-in for Lightning's AttributeDict / OmegaConf DictConfig)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(obj):
    if isinstance(obj, AttributeDict):
        return obj
    if hasattr(obj, "items") and not isinstance(obj, torch.Tensor):
        return AttributeDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)) or type(obj).__name__ == "ListConfig":
        return [to_attr(v) for v in obj]
    return obj


def linear_beta_schedule(beta_start, beta_end, timesteps):
    return torch.linspace(beta_start, beta_end, timesteps)  # task/diffusion.py:28-29


def _upd(mode, s=(), has_noise=False, w=0.0):
    u = DrbUpdate()
    u.mode = mode
    u.has_noise = 1 if has_noise else 0
    for i, v in enumerate(s):
        u.s[i] = float(v)
    u.w = float(w)
    return u


class SpecRollDiffusion(nn.Module):
    def __init__(self, lr, timesteps, loss_type, loss_keys, beta_start, beta_end, frame_threshold,
                 training, sampling, debug=False, generation_filter=0.0):
        super().__init__()
        hp = getattr(self, "_pending_hparams", AttributeDict())
        hp.update(dict(lr=lr, timesteps=timesteps, loss_type=loss_type, loss_keys=loss_keys,
                       beta_start=beta_start, beta_end=beta_end, frame_threshold=frame_threshold,
                       training=to_attr(training), sampling=to_attr(sampling), debug=debug,
                       generation_filter=generation_filter))
        self._hparams = hp
        # Schedule: the same torch expressions as task/diffusion.py:239-256, plain fp32 CPU tensors.
        self.betas = linear_beta_schedule(beta_start, beta_end, timesteps=timesteps)
        alphas = 1. - self.betas
        alphas_cumprod = torch.cumprod(alphas, axis=0)
        alphas_cumprod_prev = F.pad(alphas_cumprod[:-1], (1, 0), value=1.0)
        self.sqrt_recip_alphas = torch.sqrt(1.0 / alphas)
        self.sqrt_alphas_cumprod = torch.sqrt(alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1. - alphas_cumprod)
        self.posterior_variance = self.betas * (1. - alphas_cumprod_prev) / (1 - alphas_cumprod)
        self.alphas = alphas
        # AttributeError for an unknown sampler name, like the reference's getattr (task/diffusion.py:255)
        self.reverse_diffusion = getattr(self, self._hparams.sampling.type)

    # ---- Lightning-ish surface ---------------------------------------------------------
    @property
    def hparams(self):
        return self._hparams

    def log(self, *a, **k):
        pass

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location="cpu", strict=True, **overrides):
        """Lightning-format checkpoint: {'state_dict', 'hyper_parameters'}; keyword overrides win
        (sampling.py:54-65, test.py:30)."""
        ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(overrides)
        model = cls(**hp)
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        return model

    # ---- update coefficients: fp32 scalar arithmetic exactly as the reference writes it ----
    def _x0_update(self, t_index, ddim=False):
        """task/diffusion.py:1013-1023 (same text at :836-851, :957-967, :985-995; sigma=0 at :860-871, :1043-1053)."""
        if t_index == 0:
            return _upd(_lib.UPD_X0_FINAL, [self.sqrt_alphas_cumprod[t_index]])
        if ddim:
            sigma = torch.tensor(0.0)
        else:
            sigma = (self.sqrt_one_minus_alphas_cumprod[t_index - 1] / self.sqrt_one_minus_alphas_cumprod[t_index]) * (
                torch.sqrt(1 - self.alphas[t_index]))
        coef = torch.sqrt(1 - self.sqrt_alphas_cumprod[t_index - 1] ** 2 - sigma ** 2)
        return _upd(_lib.UPD_X0, [self.sqrt_alphas_cumprod[t_index - 1], coef, self.sqrt_alphas_cumprod[t_index],
                                  self.sqrt_one_minus_alphas_cumprod[t_index], sigma], has_noise=not ddim)

    def _w(self):
        return float(self.hparams.sampling.w)   # AttributeError when absent, like the reference's hparams access (:1009)

    # ---- samplers: (x, waveform, t_index) -> (x_prev, spec) ----------------------------------
    def inpainting_ddpm_x0(self, x, waveform, t_index, noise=None):
        upd = self._x0_update(t_index)
        upd.w = self._w()
        return self._step(x, waveform, t_index, upd, _lib.BRANCH_COND_UNCOND, noise,
                          self.hparams.inpainting_t, self.hparams.inpainting_f)

    def cfdg_ddpm_x0(self, x, waveform, t_index, noise=None):
        upd = self._x0_update(t_index)
        upd.w = self._w()
        return self._step(x, waveform, t_index, upd, _lib.BRANCH_COND_UNCOND, noise)

    def generation_ddpm_x0(self, x, waveform, t_index, noise=None):
        return self._step(x, waveform, t_index, self._x0_update(t_index), _lib.BRANCH_UNCOND, noise)

    def ddpm_x0(self, x, waveform, t_index, noise=None):
        return self._step(x, waveform, t_index, self._x0_update(t_index), _lib.BRANCH_COND, noise)

    def ddim_x0(self, x, waveform, t_index, noise=None):
        return self._step(x, waveform, t_index, self._x0_update(t_index, ddim=True), _lib.BRANCH_COND, noise)

    def cfdg_ddim_x0(self, x, waveform, t_index, noise=None):
        # task/diffusion.py:1027-1055.  The reference's second forward runs on a zero waveform WITHOUT
        # sampling=True (:1039), i.e. it is conditioned on an all-zero normalised spectrogram, not on -1.
        upd = self._x0_update(t_index, ddim=True)
        upd.w = self._w()
        return self._step(x, waveform, t_index, upd, _lib.BRANCH_COND_ZEROSPEC, noise)

    def ddpm(self, x, waveform, t_index, noise=None):                       # task/diffusion.py:804-829
        s = [self.sqrt_recip_alphas[t_index], self.betas[t_index], self.sqrt_one_minus_alphas_cumprod[t_index],
             torch.sqrt(self.posterior_variance[t_index]) if t_index > 0 else 0.0]
        return self._step(x, waveform, t_index, _upd(_lib.UPD_EPS_DDPM, s, has_noise=t_index > 0), _lib.BRANCH_COND, noise)

    def ddim(self, x, waveform, t_index, noise=None):                       # task/diffusion.py:877-892
        if t_index == 0:
            upd = _upd(_lib.UPD_EPS_FINAL, [self.sqrt_one_minus_alphas_cumprod[0], self.sqrt_alphas_cumprod[0]])
        else:
            upd = _upd(_lib.UPD_EPS_DDIM, [self.sqrt_alphas_cumprod[t_index - 1], self.sqrt_one_minus_alphas_cumprod[t_index],
                                           self.sqrt_alphas_cumprod[t_index], self.sqrt_one_minus_alphas_cumprod[t_index - 1], 0.0])
        return self._step(x, waveform, t_index, upd, _lib.BRANCH_COND, noise)

    def ddim2ddpm(self, x, waveform, t_index, noise=None):                  # task/diffusion.py:894-911
        if t_index == 0:
            upd = _upd(_lib.UPD_EPS_FINAL, [self.sqrt_one_minus_alphas_cumprod[0], self.sqrt_alphas_cumprod[0]])
        else:
            sigma = (self.sqrt_one_minus_alphas_cumprod[t_index - 1] / self.sqrt_one_minus_alphas_cumprod[t_index]) * (
                torch.sqrt(1 - self.alphas[t_index]))
            upd = _upd(_lib.UPD_EPS_DDIM, [self.sqrt_alphas_cumprod[t_index - 1], self.sqrt_one_minus_alphas_cumprod[t_index],
                                           self.sqrt_alphas_cumprod[t_index],
                                           torch.sqrt(1 - self.sqrt_alphas_cumprod[t_index - 1] ** 2 - sigma ** 2), sigma],
                       has_noise=True)
        return self._step(x, waveform, t_index, upd, _lib.BRANCH_COND, noise)

    # ---- loops ----------------------------------------------------------------------------------
    def _all_updates(self):
        """The drb_update of every step t = T-1 .. 0 for the configured sampler, without running it."""
        # 10 ms of host work for a 200-step chain (torch scalar arithmetic per step) during which the GPU would idle at the
        # head of every sample_loop call: memoised on everything the sampler methods read.
        hp = self.hparams
        sched = (self.betas, self.alphas, self.sqrt_recip_alphas, self.sqrt_alphas_cumprod,
                 self.sqrt_one_minus_alphas_cumprod, self.posterior_variance)
        key = (hp.sampling.type, repr(hp.sampling.get("w", None)), hp.timesteps, repr(hp.get("inpainting_t", None)),
               repr(hp.get("inpainting_f", None)), getattr(self.reverse_diffusion, "__func__", self.reverse_diffusion),
               tuple((id(t), t._version) for t in sched))
        hit = self.__dict__.get("_updates_memo")
        if hit is not None and hit[0] == key:
            return hit[1]
        probe = _UpdateProbe(self)
        ups = []
        for t_index in reversed(range(hp.timesteps)):
            ups.append(probe.capture(t_index))
        self.__dict__["_updates_memo"] = (key, (ups, probe.branches, probe.masks), sched)   # sched kept alive: ids stay unique
        return ups, probe.branches, probe.masks

    # device bytes of pre-drawn noise held at any time by sample_loop (the reference draws one step at a time, :1023)
    NOISE_CHUNK_BYTES = 1 << 30
    # trajectories up to this size live in pinned host memory (async copies overlap compute); larger ones are pageable
    PINNED_TRAJ_BYTES = 8 << 30

    @torch.no_grad()
    def sample_loop(self, x_T, waveform, noise=None, keep_trajectory=False, n_steps=None, generator=None, shard=None):
        """The loop body of predict_step / sampling (task/diffusion.py:528-534, 779-788) as a few library calls.

        noise: optional pre-drawn [n_noisy_steps, B, 1, T, 88] (device or host tensor).  By default the noise is drawn
        step by step with ``torch.randn_like`` on the roll's device (``generator``: a CUDA ``torch.Generator`` to draw
        from instead of the default one), in the reference's order (descending t, only steps with noise), in chunks
        of at most NOISE_CHUNK_BYTES so a 1000-step chain does not hold 1000 noise tensors.
        shard: (global_batch, lo, hi) when x_T holds rows lo:hi of a global batch split over ranks: every step's noise
        is then drawn for the GLOBAL batch and sliced, so an N-rank run consumes the generator exactly like a 1-rank
        run and returns the same rolls.
        n_steps: run only the first n_steps of the chain (t = T-1 .. T-n_steps); default the whole chain.
        Returns (x_0, spec, trajectory) -- trajectory is a host tensor [steps, B, 1, T, 88] or None; it is reused by
        the next call with the same shape (copy it if you need to keep it).
        """
        ups, branches, masks = self._all_updates()
        T_all = self.hparams.timesteps
        if n_steps is not None:
            if not 0 < n_steps <= T_all:
                raise ValueError("n_steps must be in [1, timesteps]")
            ups = ups[:n_steps]
        eng, x, spec = self._prepare(x_T, waveform, branches, *masks)
        if getattr(self, "_learned", False):      # condition='trainable_spec': see ClassifierFreeDiffRoll._learned_upd
            ups = [self._learned_upd(u, branches) for u in ups]
        x = x.clone()
        n_total = sum(1 for u in ups if u.has_noise)
        rng_state = None
        if noise is None and self._range_guarded():     # a re-run in the range-safe format must see the same noise
            rng_state = generator.get_state() if generator is not None else torch.cuda.get_rng_state(x.device)
        if noise is not None:
            if noise.shape[0] < n_total:
                raise ValueError(f"need {n_total} noise slices, got {noise.shape[0]}")
            resident = noise.is_cuda and noise.dtype == torch.float32 and noise.shape[3] == x.shape[2] and noise.is_contiguous()
        if shard is not None:
            gB, lo, hi = (int(v) for v in shard)
            if hi - lo != x.shape[0] or not 0 <= lo <= hi <= gB:
                raise ValueError(f"shard {shard} does not describe a batch of {x.shape[0]} rolls")
        traj = None
        if keep_trajectory:
            # pinned host memory is expensive to allocate (~0.5 s per GB): keep one buffer per shape
            shape = (len(ups),) + tuple(x.shape)
            traj = self.__dict__.get("_traj_buf")
            if traj is None or tuple(traj.shape) != shape:
                self.__dict__["_traj_buf"] = None
                nbytes = 4
                for d in shape:
                    nbytes *= d
                traj = torch.empty(shape, dtype=torch.float32, pin_memory=nbytes <= self.PINNED_TRAJ_BYTES)
                self.__dict__["_traj_buf"] = traj
        per_step = x.numel() * 4
        chunk = max(1, min(len(ups), self.NOISE_CHUNK_BYTES // per_step))
        k = 0                                            # noisy steps consumed so far
        for i0 in range(0, len(ups), chunk):
            cu = ups[i0:i0 + chunk]
            nn = sum(1 for u in cu if u.has_noise)
            nz = None
            if nn and noise is not None:
                nz = noise[k:k + nn] if resident else noise[k:k + nn].to(device=x.device, dtype=torch.float32)[..., :x.shape[2], :].contiguous()
            elif nn:
                nz = torch.empty((nn,) + tuple(x.shape), device=x.device)
                for j in range(nn):                      # same generator consumption order as task/diffusion.py:1023
                    if shard is not None:
                        nz[j] = torch.randn((gB,) + tuple(x.shape[1:]), device=x.device, generator=generator)[lo:hi]
                    else:                                # drawn in place: the same Philox consumption as randn_like(x)
                        torch.randn(tuple(x.shape), generator=generator, out=nz[j])
            k += nn
            eng.loop(x, nz, cu, T_all - i0, T_all - i0 - len(cu), None if traj is None else traj[i0:i0 + len(cu)])
        if not self._range_ok(eng):
            if rng_state is not None:
                if generator is not None:
                    generator.set_state(rng_state)
                else:
                    torch.cuda.set_rng_state(rng_state, x.device)
            self._range_fallback()
            return self.sample_loop(x_T, waveform, noise=noise, keep_trajectory=keep_trajectory, n_steps=n_steps,
                                    generator=generator, shard=shard)
        return x, spec, traj

    def release_buffers(self):
        """Drop the cached host trajectory buffer and every engine (device workspaces) of this model."""
        self.__dict__.pop("_traj_buf", None)
        for eng, _ in getattr(self, "_engines", {}).values():
            eng.close()
        if hasattr(self, "_engines"):
            self._engines.clear()
        for eng in self.__dict__.pop("_train_engines", {}).values():
            eng.close()

    @torch.no_grad()
    def predict_step(self, batch, batch_idx=0, generator=None, shard=None):
        """task/diffusion.py:513-534.  batch = (x_T [B,1,T,88], waveform [B,L][, roll_label]).

        The reference copies every intermediate roll to the host (:530) and then writes figures/MIDI
        (:540-618, out of scope).  Here the per-step copies are asynchronous into pinned memory and the
        method returns ``(roll_pred, noise_list, spec)`` instead of None; ``noise_list`` has the
        reference's structure: [(x_T, timesteps), (x_{T-1} as numpy, T-1), ..., (x_0 as numpy, 0)].  The numpy entries
        are views of one pinned host buffer that the next call with the same shape overwrites; ``roll_pred`` is a copy.
        """
        noise, waveform = batch[0], batch[1]
        x0, spec, traj = self.sample_loop(noise, waveform, keep_trajectory=True, generator=generator, shard=shard)
        torch.cuda.current_stream(x0.device).synchronize()
        noise_list = [(noise, self.hparams.timesteps)]
        tnp = traj.numpy()
        for i, t_index in enumerate(reversed(range(self.hparams.timesteps))):
            noise_list.append((tnp[i], t_index))
        roll_pred = noise_list[-1][0].copy()   # independent of the reused pinned trajectory buffer
        return roll_pred, noise_list, spec

    @torch.no_grad()
    def sampling(self, batch, batch_idx=0, zero_copy=False):
        """task/diffusion.py:765-790: batch = {'frame': [B,T,88], 'audio': [B,L]} -> (noise_list, spec).

        The reference returns independent ``.cpu().numpy()`` arrays (:784); so does this method.  ``zero_copy=True``
        returns views of the single reused host trajectory buffer instead (the next sampling / predict_step / test_step
        call with the same shape overwrites them)."""
        roll = self.normalize(batch["frame"]).unsqueeze(1)
        waveform = batch["audio"]
        noise = torch.randn_like(roll)
        if self.hparams.debug:
            raise NotImplementedError("debug=True conditions on the label roll (DiffRollDebug); not on this path")
        _, noise_list, spec = self.predict_step((noise, waveform), batch_idx)
        if not zero_copy:
            noise_list = [noise_list[0]] + [(a.copy(), t) for a, t in noise_list[1:]]
        return noise_list, spec

    @torch.no_grad()
    def notes_from_rolls(self, rolls):
        """The first half of predict_step's MIDI tail (task/diffusion.py:599-602) on the GPU: for every finished roll
        ``extract_notes_wo_velocity(np_frame, np_frame)`` with the default 0.5 thresholds.  rolls: CUDA tensor
        [B,1,T,88] -> list of (pitches, intervals) numpy arrays in the reference's order.  Scaling to seconds / Hz
        and writing .mid files stay on the host (I/O, out of scope)."""
        from .notes import extract_notes_batch
        r = rolls[:, 0] if rolls.ndim == 4 else rolls
        return extract_notes_batch(r, r)

    @torch.no_grad()
    def test_step(self, batch, batch_idx=0):
        """task/diffusion.py:312-418 up to the host-library boundary: ``sampling`` (the whole reverse chain from fresh
        noise), the frame-level precision / recall / F1 of :378-380 and the note lists of :382-396 for the estimate
        and the label, all on the GPU.  The mir_eval note metrics, figures, GIF, audio and MIDI files of the reference
        are host I/O and not built.  batch = {'frame': [B,T,88], 'audio': [B,L]} on the GPU.
        Returns a dict: frame_p, frame_r, frame_f1, counts (tp, fp, fn), notes_est / notes_ref (lists of
        (pitches, intervals) in frames, the reference's order), roll_pred (numpy [B,1,T,88]), spec."""
        from .notes import extract_notes_batch, frame_precision_recall_f1
        noise_list, spec = self.sampling(batch, batch_idx, zero_copy=True)
        roll_pred = noise_list[-1][0].copy()   # independent of the reused host trajectory buffer
        thr = self.hparams.frame_threshold
        label = batch["frame"]
        pred_dev = torch.from_numpy(roll_pred).to(label.device)
        T = pred_dev.shape[2]
        lab = label[:, :T, :].to(torch.float32).contiguous()          # trim_spec_roll may have shortened the roll
        p, r, f1, counts = frame_precision_recall_f1(lab, pred_dev[:, 0], thr)
        self.log("Test/Frame_F1", f1)
        est = extract_notes_batch(pred_dev[:, 0], pred_dev[:, 0], thr, thr)
        ref = extract_notes_batch(lab, lab, thr, thr)
        return dict(frame_p=p, frame_r=r, frame_f1=f1, counts=counts, notes_est=est, notes_ref=ref,
                    roll_pred=roll_pred, spec=spec)

    def p_losses(self, label, prediction, loss_type="l1"):
        """task/diffusion.py:792-802 (mean l1 / l2 / smooth-l1) as one reduction kernel; 0-dim CUDA tensor."""
        from .diffusion_ops import p_losses
        return p_losses(label, prediction, loss_type)

    @torch.no_grad()
    def step(self, batch, t=None, noise=None):
        """task/diffusion.py:651-763, forward only (what ``validation_step`` :271-276 runs): draw a diffusion step per
        roll, ``q_sample`` the normalised label roll, one network forward at per-roll steps, and the diffusion loss.

        batch: {'frame': [B,T,88], 'audio': [B,L]} on the GPU, or a list of two such dicts (the second one is run
        unconditionally, :707-719).  ``t`` / ``noise`` default to the reference's draws (``torch.randint`` :667,
        ``torch.randn_like`` :670, in that order on the roll's device) and can be injected for parity tests.
        Returns (losses, tensors) with the reference's keys.  Gradients are not built: the backward pass is SURVEY.md
        row f3, not part of this path."""
        from .diffusion_ops import extract_x0, q_sample
        two = isinstance(batch, list)
        first = batch[0] if two else batch
        roll = self.normalize(first["frame"]).unsqueeze(1)
        waveform = first["audio"]
        batch_size = roll.shape[0]
        device = roll.device
        if t is None:
            t = torch.randint(0, self.hparams.timesteps, (batch_size,), device=device).long()
        if noise is None:
            noise = torch.randn_like(roll)
        sa, s1 = self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod
        x_t = q_sample(roll, t, sa, s1, noise)
        mode = self.hparams.training.mode
        if self.hparams.debug:
            raise NotImplementedError("debug=True conditions on the label roll (DiffRollDebug); not on this path")
        losses, tensors = {}, {}
        if mode == 'epsilon':
            epsilon_pred, spec = self(x_t, waveform, t)
            losses["diffusion_loss"] = self.p_losses(noise, epsilon_pred, loss_type=self.hparams.loss_type)
            pred_roll = extract_x0(x_t, epsilon_pred, t, sa, s1)
        elif mode == 'x_0':
            pred_roll, spec = self(x_t, waveform, t)
            losses["diffusion_loss"] = self.p_losses(roll, pred_roll, loss_type=self.hparams.loss_type)
            if two:
                roll2 = self.normalize(batch[1]["frame"]).unsqueeze(1)
                x_t2 = q_sample(roll2, t, sa, s1, noise)
                pred_roll2, spec2 = self(x_t2, batch[1]["audio"], t, sampling=True)
                losses["unconditional_diffusion_loss"] = self.p_losses(roll2, pred_roll2, loss_type=self.hparams.loss_type)
                tensors.update(spec2=spec2, label_roll2=roll2, pred_roll2=pred_roll2)
        elif mode == 'ex_0':
            epsilon_pred, spec = self(x_t, waveform, t)
            pred_roll = extract_x0(x_t, epsilon_pred, t, sa, s1)
            losses["diffusion_loss"] = self.p_losses(roll, pred_roll, loss_type=self.hparams.loss_type)
        else:
            raise ValueError(f"training mode {mode} is not supported. Please either use 'x_0' or 'epsilon'.")
        tensors.update(pred_roll=pred_roll, label_roll=roll, spec=spec)
        return losses, tensors

    # ---- training step ------------------------------------------------------------------
    def fixed_dropout(self, x, p, masked_value=-1, mask=None):
        """model/diffwave.py:689-693: every roll's conditioning is replaced by ``masked_value`` with probability ``p``.
        The reference samples ``torch.distributions.Bernoulli(probs=p).sample((B,))``, i.e. B draws from the CPU generator;
        ``torch.bernoulli`` over a CPU vector of p consumes the generator identically.  ``mask`` ([B], nonzero = drop)
        overrides the draw (parity tests)."""
        if mask is None:
            mask = torch.bernoulli(torch.full((x.shape[0],), float(p)))
        drop = mask.to(device=x.device).bool()
        x[drop] = masked_value
        return x

    uncon_dropout = fixed_dropout        # condition == 'fixed' (model/diffwave.py:617)

    def _train_engine(self, batch, frames):
        from .train import TrainEngine
        key = (int(batch), int(frames), str(next(self.parameters()).device))
        cache = self.__dict__.setdefault("_train_engines", {})
        eng = cache.get(key)
        if eng is None:
            for old in list(cache.values()):      # one training shape at a time: a workspace is gigabytes
                old.close()
            cache.clear()
            eng = cache[key] = TrainEngine(self, batch, frames)
        return eng

    @torch.no_grad()
    def training_step(self, batch, batch_idx=0, t=None, noise=None, dropout_mask=None, want_input_grad=False):
        """task/diffusion.py:258-270 + ``step`` :651-763 in training mode: draw a diffusion step per roll, ``q_sample`` the
        normalised label roll, the network forward WITH spec dropout (model/diffwave.py:646-647), the losses of
        ``hparams.loss_keys`` -- and, because there is no autograd graph here, the whole backward pass: on return every
        parameter's ``.grad`` holds d(total loss)/d(parameter) (added to an existing gradient, like ``loss.backward()``).
        Returns the total loss (0-dim CUDA tensor).  ``t`` / ``noise`` / ``dropout_mask`` default to the reference's draws
        and can be injected for parity tests; the last step's (losses, tensors) are kept in ``self.last_step``."""
        from .diffusion_ops import extract_x0, q_sample
        from .train import loss_grad
        two = isinstance(batch, list)
        first = batch[0] if two else batch
        roll = self.normalize(first["frame"]).unsqueeze(1)
        waveform = first["audio"]
        B, _, T, _ = roll.shape
        device = roll.device
        if t is None:
            t = torch.randint(0, self.hparams.timesteps, (B,), device=device).long()
        if noise is None:
            noise = torch.randn_like(roll)
        sa, s1 = self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod
        x_t = q_sample(roll, t, sa, s1, noise)
        mode = self.hparams.training.mode
        if self.hparams.debug:
            raise NotImplementedError("debug=True conditions on the label roll (DiffRollDebug); not on this path")
        if mode not in ("epsilon", "x_0", "ex_0"):
            raise ValueError(f"training mode {mode} is not supported. Please either use 'x_0' or 'epsilon'.")
        learned = getattr(self, "_learned", False)      # condition='trainable_spec': the dropped rolls read a parameter
        from . import _lib
        _, _, spec = self._prepare(x_t, waveform, _lib.BRANCH_COND, self.hparams.inpainting_t, self.hparams.inpainting_f, mel_only=True)
        if spec.shape[-1] != T:
            raise NotImplementedError("training step: the clip must cover the whole roll (trim_spec_roll would shorten it)")
        spec = spec.clone()
        drop = None
        if self.training:                                  # model/diffwave.py:646-647
            if learned and dropout_mask is None:           # the same draw uncon_dropout would make; kept to route the gradient
                dropout_mask = torch.bernoulli(torch.full((B,), float(self.hparams.spec_dropout)))
            spec = self.uncon_dropout(spec, self.hparams.spec_dropout, mask=dropout_mask)
            if learned:
                drop = dropout_mask.to(device=device).bool()
        eng = self._train_engine(B, T)
        keys = list(self.hparams.loss_keys)
        losses, tensors = {}, {}
        net = eng.forward(x_t, spec, t)
        acc = any(q.grad is not None for q in self.parameters())
        if mode == "epsilon":
            losses["diffusion_loss"] = self.p_losses(noise, net, loss_type=self.hparams.loss_type)
            g = loss_grad(noise, net, self.hparams.loss_type)
            pred_roll = extract_x0(x_t, net, t, sa, s1)
        elif mode == "x_0":
            pred_roll = net
            losses["diffusion_loss"] = self.p_losses(roll, pred_roll, loss_type=self.hparams.loss_type)
            g = loss_grad(roll, pred_roll, self.hparams.loss_type)
        else:                                              # 'ex_0': the loss is taken on x0 extracted from the predicted noise
            pred_roll = extract_x0(x_t, net, t, sa, s1)
            losses["diffusion_loss"] = self.p_losses(roll, pred_roll, loss_type=self.hparams.loss_type)
            scale = -(s1.to(device)[t] / sa.to(device)[t])
            g = loss_grad(roll, pred_roll, self.hparams.loss_type, roll_scale=scale)
        gx = None
        if "diffusion_loss" in keys:
            gx = eng.backward(g, accumulate=acc, want_input_grad=want_input_grad,
                              want_spec_grad=drop is not None and bool(dropout_mask.bool().any()))
            acc = True
            if eng.spec_grad is not None:                  # trainable_dropout (:695-699): d loss / d table = sum over the dropped rolls
                self.trainable_parameters.grad[:, :T].add_(eng.spec_grad[drop].sum(0))
        tensors.update(pred_roll=pred_roll, label_roll=roll, spec=spec)
        if two and mode == "x_0":                          # second dataset: one more, unconditional, forward (:707-719)
            roll2 = self.normalize(batch[1]["frame"]).unsqueeze(1)
            x_t2 = q_sample(roll2, t, sa, s1, noise)
            spec2 = self._uncond_spec_like(spec)           # sampling=True, model/diffwave.py:656-660
            pred_roll2 = eng.forward(x_t2, spec2, t)
            losses["unconditional_diffusion_loss"] = self.p_losses(roll2, pred_roll2, loss_type=self.hparams.loss_type)
            if "unconditional_diffusion_loss" in keys:
                eng.backward(loss_grad(roll2, pred_roll2, self.hparams.loss_type), accumulate=acc, want_spec_grad=learned)
                if learned:                                # every roll of the second batch is conditioned on the table (:657-658)
                    self.trainable_parameters.grad[:, :T].add_(eng.spec_grad.sum(0))
            tensors.update(spec2=spec2, label_roll2=roll2, pred_roll2=pred_roll2)
        total_loss = 0
        for k in keys:
            total_loss = total_loss + losses[k]
            self.log(f"Train/{k}", losses[k])
        self.last_step = (losses, tensors, gx)
        return total_loss

    def configure_optimizers(self):
        """task/diffusion.py:1057-1067: ``torch.optim.Adam(self.parameters(), lr=self.hparams.lr)`` as fused CUDA updates."""
        from .train import Adam
        return [Adam(self.parameters(), lr=self.hparams.lr)]

    @torch.no_grad()
    def validation_step(self, batch, batch_idx=0):
        """task/diffusion.py:271-276 without the figure logging: returns the summed loss over ``hparams.loss_keys``."""
        losses, _ = self.step(batch)
        total_loss = 0
        for k in self.hparams.loss_keys:
            total_loss = total_loss + losses[k]
            self.log(f"Val/{k}", losses[k])
        return total_loss

    # range guard of the fp16-based operand formats: no-ops here, implemented by the model subclass
    def _range_guarded(self):
        return False

    def _range_ok(self, eng):
        return True

    def _range_fallback(self):
        raise NotImplementedError

    # provided by the model subclass
    def _step(self, x, waveform, t_index, upd, branches, noise=None, inpainting_t=None, inpainting_f=None):
        raise NotImplementedError

    def _prepare(self, x, waveform, branches, inpainting_t=None, inpainting_f=None, mel_only=False):
        raise NotImplementedError


class _UpdateProbe:
    """Runs a sampler method with ``_step`` intercepted, to read off its update struct, branches and masks."""

    def __init__(self, model):
        self.model = model
        self.branches = None
        self.masks = (None, None)

    def capture(self, t_index):
        got = {}

        def fake_step(x, waveform, t, upd, branches, noise=None, inpainting_t=None, inpainting_f=None):
            got["upd"] = upd
            self.branches = branches
            self.masks = (inpainting_t, inpainting_f)
            return None, None

        m = self.model
        orig = m.__dict__.get("_step")
        m.__dict__["_step"] = fake_step
        try:
            m.reverse_diffusion(None, None, t_index)
        finally:
            if orig is None:
                del m.__dict__["_step"]
            else:
                m.__dict__["_step"] = orig
        return got["upd"]

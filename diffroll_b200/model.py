"""``ClassifierFreeDiffRoll`` with the reference's module surface (model/diffwave.py:579-699).

Same constructor keywords, same ``state_dict`` keys and shapes, same
``forward(x_t, waveform, diffusion_step, sampling=False, inpainting_t=None, inpainting_f=None)
-> (pred [B,1,T,88], spec [B,n_mels,T])``.  The parameters live in ordinary ``nn.Conv1d`` /
``nn.Linear`` containers so reference checkpoints load unchanged, but those containers are
never *called*: every forward goes through the CUDA engine.  There is no CPU path.
"""
from __future__ import annotations

from math import sqrt  # noqa: F401  (kept for parity with the reference's namespace)

import weakref

import torch
import torch.nn as nn

from . import _lib
from ._lib import DrbUpdate
from .engine import Engine
from .synthetic import hann_window, melscale_fbanks
from .task import AttributeDict, SpecRollDiffusion, _upd, to_attr


def _minmax_rescale(x, dims, lo, hi, nan_to):
    """Scale ``x`` so that its minimum / maximum over ``dims`` land on ``lo`` / ``hi``; a constant slice (0/0) becomes ``nan_to``."""
    top = x.amax(dim=dims, keepdim=True)
    bottom = x.amin(dim=dims, keepdim=True)
    unit = (x - bottom) / (top - bottom)
    if nan_to[0] == "unit":                       # framewise: the NaNs are replaced BEFORE the affine map (model/utils.py:13-15)
        unit = torch.where(torch.isnan(unit), torch.zeros_like(unit), unit)
        return unit * (hi - lo) + lo
    out = unit * (hi - lo) + lo                   # imagewise: replaced after it, by the lower bound (model/utils.py:28-30)
    return torch.where(torch.isnan(out), torch.full_like(out, lo), out)


class Normalization:
    """Min-max scaling of label rolls / spectrograms with the reference's call surface (model/utils.py:2-38):
    ``Normalization(min, max, mode)(x)``.  'imagewise' scales each batch element over all of its entries, 'framewise' over
    dim 1 only.  CUDA inputs of the imagewise mode go through one kernel (``drb_normalize_imagewise``); the torch
    expressions below serve host tensors (labels before they are moved to the GPU) and are not on the hot path."""

    MODES = ("framewise", "imagewise")

    def __init__(self, min, max, mode='imagewise'):
        self.lo, self.hi, self.mode = min, max, mode
        if mode not in self.MODES:                # the reference prints a hint and fails later at call time; fail at the source
            raise ValueError(f"Normalization mode must be one of {self.MODES}, got {mode!r}")

    def normalize(self, x):
        if self.mode == "framewise":
            return _minmax_rescale(x, (1,), self.lo, self.hi, ("unit",))
        if x.is_cuda:
            from .diffusion_ops import normalize_imagewise
            return normalize_imagewise(x, self.lo, self.hi)
        return _minmax_rescale(x, tuple(range(1, x.dim())), self.lo, self.hi, ("out",))

    __call__ = normalize


def Conv1d(*args, **kwargs):
    layer = nn.Conv1d(*args, **kwargs)
    nn.init.kaiming_normal_(layer.weight)  # model/diffwave.py:41-44
    return layer


class DiffusionEmbedding(nn.Module):
    """Parameter container + the sinusoid table of model/diffwave.py:58-88 (built with the same expression)."""

    def __init__(self, max_steps):
        super().__init__()
        self.register_buffer('embedding', self._build_embedding(max_steps), persistent=False)
        self.projection1 = nn.Linear(128, 512)
        self.projection2 = nn.Linear(512, 512)

    def _build_embedding(self, max_steps):
        steps = torch.arange(max_steps).unsqueeze(1)
        dims = torch.arange(64).unsqueeze(0)
        table = steps * 10.0 ** (dims * 4.0 / 63.0)
        return torch.cat([torch.sin(table), torch.cos(table)], dim=1)


class ResidualBlock(nn.Module):
    """Parameter container with the reference's names and shapes (model/diffwave.py:107-132)."""

    def __init__(self, n_mels, residual_channels, dilation, kernel_size=3, uncond=False):
        super().__init__()
        self.dilated_conv = Conv1d(residual_channels, 2 * residual_channels, kernel_size,
                                   padding=((kernel_size - 1) * (dilation - 1) + kernel_size - 1) // 2,
                                   dilation=dilation)
        self.diffusion_projection = nn.Linear(512, residual_channels)
        self.conditioner_projection = None if uncond else Conv1d(n_mels, 2 * residual_channels, 1)
        self.output_projection = Conv1d(residual_channels, 2 * residual_channels, 1)


class _MelBuffers(nn.Module):
    """Holds ``mel_layer.spectrogram.window`` and ``mel_layer.mel_scale.fb`` under the reference's keys."""

    def __init__(self, sample_rate, n_fft, n_mels, f_min, f_max):
        super().__init__()
        self.spectrogram = nn.Module()
        self.spectrogram.register_buffer("window", hann_window(n_fft))
        self.mel_scale = nn.Module()
        self.mel_scale.register_buffer("fb", melscale_fbanks(n_fft // 2 + 1, float(f_min), float(f_max), n_mels, sample_rate))


class ClassifierFreeDiffRoll(SpecRollDiffusion):
    MAX_ENGINES = 4   # engines (one per batch/frames/wave_len/device) kept alive at once
    # Range guard.  The f16e5 / f16f8 formats round activations to fp16 (max 65504); weights are
    # pre-scaled per tensor by a power of two, activations are not.  Every call reads back the largest |operand| its
    # kernels emitted; a value that is not finite or above F16_RANGE_LIMIT switches this model to bf16x3 (fp32 exponent
    # range, same parity grade, ~20 % slower) and re-runs the call.  ``range_check = False`` skips the read-back.
    F16_RANGE_LIMIT = 3.0e4
    # f16n4 adds block scales stored as ue4m3 (max 448): its activation range ends near 2^7 (csrc/common.cuh); beyond
    # N4_RANGE_LIMIT the model steps down to f16e5 first.
    N4_RANGE_LIMIT = 100.0
    range_check = True

    def __init__(self, residual_channels, unconditional, condition, n_mels, norm_args,
                 residual_layers=30, kernel_size=3, dilation_base=1, dilation_bound=4, spec_args={},
                 spec_dropout=0.5, inpainting_t=None, inpainting_f=None, precision="f16n4", **kwargs):
        spec_args = to_attr(dict(spec_args))
        self._pending_hparams = AttributeDict(
            residual_channels=residual_channels, unconditional=unconditional, condition=condition, n_mels=n_mels,
            norm_args=norm_args, residual_layers=residual_layers, kernel_size=kernel_size,
            dilation_base=dilation_base, dilation_bound=dilation_bound, spec_args=spec_args,
            spec_dropout=spec_dropout, inpainting_t=inpainting_t, inpainting_f=inpainting_f)
        self.spec_dropout = spec_dropout
        super().__init__(**kwargs)
        del self._pending_hparams
        if condition == 'trainable_z':
            # model/diffwave.py:616 calls ResidualBlockz(n_mels, residual_channels, dilation, kernel_size, uncond=...) but its
            # __init__ (:154) takes (n_mels, residual_channels, dilation, uncond): the reference cannot construct this variant
            print(f"================trainable_z layers=================")
            raise TypeError("ResidualBlockz.__init__() got multiple values for argument 'uncond'")
        elif condition not in ('fixed', 'trainable_spec'):
            raise ValueError("unrecognized condition '{condition}'")  # sic: model/diffwave.py:610
        # condition == 'trainable_spec' (model/diffwave.py:600-605): the unconditional branch is conditioned on a learned
        # spectrogram instead of -1 -- in sampling the whole [n_mels, 641] table (:657-658), in training the dropped rolls (:695-699)
        self._learned = condition == 'trainable_spec'
        # unconditional=True builds residual blocks without conditioner_projection (model/diffwave.py:125-128), but
        # ClassifierFreeDiffRoll.forward always hands them a spectrogram, which ResidualBlock.forward rejects (:135-136):
        # the reference constructs such a model and fails its first forward with an AssertionError.  Same here.
        self._uncond_blocks = bool(unconditional)
        self.precision = precision
        self.input_projection = Conv1d(88, residual_channels, 1)
        self.diffusion_embedding = DiffusionEmbedding(len(self.betas))
        if self._learned:
            self.register_parameter("trainable_parameters", nn.Parameter(torch.full((spec_args["n_mels"], 641), -1.0)))  # :601-604
            self.uncon_dropout = self.trainable_dropout
        self.residual_layers = nn.ModuleList([
            ResidualBlock(n_mels, residual_channels, dilation_base ** (i % dilation_bound), kernel_size, uncond=unconditional)
            for i in range(residual_layers)])
        self.skip_projection = Conv1d(residual_channels, residual_channels, 1)
        self.output_projection = Conv1d(residual_channels, 88, 1)
        nn.init.zeros_(self.output_projection.weight)
        self.normalize_spec = Normalization(0, 1, norm_args[2])
        self.normalize = Normalization(norm_args[0], norm_args[1], norm_args[2])
        sa = spec_args
        for key, want in (("center", True), ("normalized", True), ("pad_mode", "reflect")):
            if sa.get(key, want) != want:
                raise NotImplementedError(f"mel front-end built for {key}={want!r} (config/spec/mel.yaml)")
        self.mel_layer = _MelBuffers(sa["sample_rate"], sa["n_fft"], sa["n_mels"], sa.get("f_min", 0), sa.get("f_max", sa["sample_rate"] // 2))
        self._engines = {}
        self._mel_key = None
        self._mel_ref = None
        self._spec = None
        self._learned_pair = False

    # ---- engine management ------------------------------------------------------------------------
    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def _engine(self, batch, frames, wave_len, device, any_version=False):
        key = (batch, frames, wave_len, self.precision, str(device))   # one model is either 'fixed' or 'trainable_spec' for life
        ent = self._engines.get(key)
        if ent is not None and any_version:      # mel front-end only (training step): it has no trainable tensors
            return ent[0]
        ver = self._weights_version()
        if ent is not None and ent[1] == ver:
            return ent[0]
        if ent is not None:
            ent[0].close()
        for k in [k for k, e in self._engines.items() if e[1] != ver]:
            self._engines.pop(k)[0].close()
        while len(self._engines) >= self.MAX_ENGINES:       # bounded: a workspace is gigabytes; least recently built goes first
            self._engines.pop(next(iter(self._engines)))[0].close()
        eng = Engine(self.state_dict(), self.hparams, batch, frames, wave_len, self.diffusion_embedding.embedding,
                     precision=self.precision, device=device,
                     branches=_lib.BRANCH_COND_LEARNED if self._learned else _lib.BRANCH_COND_UNCOND)
        if self._learned:     # a later change of the parameter bumps its version and rebuilds the engine like any weight
            eng.set_uncond_spec(self.trainable_parameters)
        self._engines[key] = (eng, ver)
        self._mel_key = None
        return eng

    def _prepare(self, x, waveform, branches, inpainting_t=None, inpainting_f=None, mel_only=False):
        """Pick the engine for these shapes, run the (cached) mel front-end, select branches.  mel_only: the caller wants
        the spectrogram alone (training step), so an engine built from older parameter values is good enough."""
        if not x.is_cuda:
            raise _lib.DrbError("ClassifierFreeDiffRoll (diffroll_b200) runs on CUDA tensors only; there is no CPU path")
        B, _, T, Fp = x.shape
        if Fp != 88:
            raise ValueError("piano roll must have 88 pitches")
        sa = self.hparams.spec_args
        wave_len = waveform.shape[-1]
        n_frames = wave_len // sa["hop_length"] + 1
        T_min = min(T, n_frames)                                  # trim_spec_roll, model/diffwave.py:30-39,662
        eng = self._engine(B, T_min, wave_len, x.device, any_version=mel_only)
        xx = x.to(torch.float32)
        if T_min != T:
            xx = xx[:, :, :T_min, :]
        xx = xx.contiguous()
        if self._learned and branches in (_lib.BRANCH_UNCOND, _lib.BRANCH_COND_UNCOND):
            # the learned table has 641 frames (model/diffwave.py:601) and is trimmed like a spectrogram (:662); the reference's
            # guidance pair needs both branches to come out equally long, and so does the engine
            if min(T, self.trainable_parameters.shape[-1]) != T_min:
                raise RuntimeError(f"trainable_spec: the learned spectrogram trims the roll to "
                                   f"{min(T, self.trainable_parameters.shape[-1])} frames, the clip to {T_min}")
        # A sampling=True forward on its own under 'trainable_spec' (the generation sampler, a plain forward): every roll is
        # conditioned on the table.  DRB_BRANCH_LEARNED does exactly that, but the table exists only as rows of the conditioner
        # tables, which the CTA-pair kernels read: the tensor-core formats need an even number of 128-frame tiles.  Otherwise the
        # step runs as the (clip, table) pair at guidance weight -1 -- (1 + w) * cond - w * learned = learned exactly, see
        # _learned_upd -- whose clip half only has to be finite (the mel front-end guarantees that for any waveform).
        self._learned_pair = (self._learned and branches == _lib.BRANCH_UNCOND and self.precision != "fp32"
                              and (B * ((T_min + 127) // 128)) % 2 != 0)
        if branches == _lib.BRANCH_UNCOND and not self._learned_pair:
            if self._learned:
                spec = self.trainable_parameters.detach()[..., :T_min]          # 2-D, like the reference's return value (:658,662)
                branches = _lib.BRANCH_LEARNED
            else:
                spec = torch.full((B, sa["n_mels"], T_min), -1.0, device=x.device)   # model/diffwave.py:660
        else:
            wav = waveform.to(device=x.device, dtype=torch.float32).contiguous()
            it = list(inpainting_t) if inpainting_t else None
            itf = list(inpainting_f) if inpainting_f else None
            # The spectrogram is step-invariant; the reference recomputes it twice per step.  Reuse it only while the
            # caller passes the very same (still alive, unmodified) tensor object: an address alone could be recycled.
            key = (id(eng), waveform._version, tuple(waveform.shape), tuple(it or ()), tuple(itf or ()))
            same = self._mel_ref is not None and self._mel_ref() is waveform
            if not same or key != self._mel_key or self._spec is None:
                self._spec = eng.mel(wav, it, itf)
                self._mel_key = key
                self._mel_ref = weakref.ref(waveform)
            spec = self._spec
        if self._learned and branches in (_lib.BRANCH_UNCOND, _lib.BRANCH_COND_UNCOND):   # the (clip, learned table) pair
            if branches == _lib.BRANCH_UNCOND:
                spec = self.trainable_parameters.detach()[..., :T_min]
            branches = _lib.BRANCH_COND_LEARNED
        eng.set_branches(branches)
        return eng, xx, spec

    def _learned_upd(self, upd, branches):
        """The update struct a step really runs with, after ``_prepare``: under condition='trainable_spec' a sampling=True forward
        that had to run as the learned pair takes guidance weight -1 (see _prepare).  Returns ``upd`` itself otherwise."""
        if not (self._learned and branches == _lib.BRANCH_UNCOND and self._learned_pair):
            return upd
        u = DrbUpdate.from_buffer_copy(upd)
        u.w = -1.0
        return u

    def _range_guarded(self):
        return self.range_check and self.precision in ("f16n4", "f16e5", "f16f8")

    def _range_limit(self):
        return self.N4_RANGE_LIMIT if self.precision == "f16n4" else self.F16_RANGE_LIMIT

    def _range_ok(self, eng):
        if not self._range_guarded():
            return True
        m = eng.range_max(reset=True)
        self._range_seen = m
        return m == m and m <= self._range_limit()

    def _range_fallback(self):
        import warnings
        m = getattr(self, "_range_seen", float("nan"))
        nxt = "f16e5" if (self.precision == "f16n4" and m == m and m <= self.F16_RANGE_LIMIT) else "bf16x3"
        warnings.warn(f"diffroll_b200: activations left the range of precision='{self.precision}' "
                      f"(max |x| = {m:g}, limit {self._range_limit():g}); re-running in '{nxt}'", RuntimeWarning, stacklevel=3)
        self.precision = nxt
        self._mel_key = None

    def _step(self, x, waveform, t_index, upd, branches, noise=None, inpainting_t=None, inpainting_f=None):
        eng, xx, spec = self._prepare(x, waveform, branches, inpainting_t, inpainting_f)
        if upd.has_noise:
            if noise is None:
                noise = torch.randn_like(xx)       # task/diffusion.py:1023
            else:
                noise = noise.to(device=xx.device, dtype=torch.float32)[:, :, :xx.shape[2], :].contiguous()
        else:
            noise = None
        out = eng.step(xx, noise, t_index, self._learned_upd(upd, branches))
        if not self._range_ok(eng):
            self._range_fallback()
            return self._step(x, waveform, t_index, upd, branches, noise, inpainting_t, inpainting_f)
        return out, spec

    # ---- reference forward ---------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x_t, waveform, diffusion_step, sampling=False, inpainting_t=None, inpainting_f=None):
        """model/diffwave.py:637-686.  ``diffusion_step``: int64[B]; the samplers pass one value repeated
        (``tensor(t).repeat(B)``), the training / validation step a different one per roll (task/diffusion.py:667)."""
        assert not self._uncond_blocks, \
            "ResidualBlock built with uncond=True received a conditioner (model/diffwave.py:135-136)"
        if self.training:
            return self._forward_training(x_t, waveform, diffusion_step, sampling, inpainting_t, inpainting_f)
        if torch.is_tensor(diffusion_step) and diffusion_step.dtype not in (torch.int32, torch.int64):
            return self._forward_fractional(x_t, waveform, diffusion_step, sampling, inpainting_t, inpainting_f)
        if torch.is_tensor(diffusion_step):
            steps = diffusion_step.flatten()
            if steps.numel() != x_t.shape[0]:
                raise ValueError("diffusion_step must hold one step per roll")
            lo, hi = int(steps.min()), int(steps.max())
            if lo < 0 or hi >= len(self.betas):
                raise IndexError(f"diffusion_step out of range [0, {len(self.betas)})")   # the embedding lookup of :670 would raise
            t0, per_roll = lo, lo != hi
        else:
            t0, per_roll = int(diffusion_step), False
        branches = _lib.BRANCH_UNCOND if sampling is True else _lib.BRANCH_COND
        eng, xx, spec = self._prepare(x_t, waveform, branches, inpainting_t, inpainting_f)
        none = self._learned_upd(_upd(_lib.UPD_NONE), branches)
        if not per_roll:
            pred = eng.step(xx, None, t0, none)
        else:
            eng.set_steps(steps)
            try:
                pred = eng.step(xx, None, t0, none)
            finally:
                eng.set_steps(None)
        if not self._range_ok(eng):
            self._range_fallback()
            return self.forward(x_t, waveform, diffusion_step, sampling, inpainting_t, inpainting_f)
        return pred, spec

    def _forward_fractional(self, x_t, waveform, diffusion_step, sampling, inpainting_t, inpainting_f):
        """Floating-point ``diffusion_step`` (model/diffwave.py:66-81): the sinusoid table is interpolated linearly between the two
        neighbouring integer steps, per roll, and the embedding MLP + the 15 diffusion projections run on those rows."""
        t = diffusion_step.flatten().to(device=x_t.device, dtype=torch.float32)
        if t.numel() != x_t.shape[0]:
            raise ValueError("diffusion_step must hold one step per roll")
        table = self.diffusion_embedding.embedding.to(x_t.device)
        low_idx, high_idx = torch.floor(t).long(), torch.ceil(t).long()       # an out-of-range step fails the lookup, as in the reference
        low, high = table[low_idx], table[high_idx]
        rows = low + (high - low) * (t - low_idx).unsqueeze(-1)               # _lerp_embedding (:76-81; t [B] against rows [B,128])
        branches = _lib.BRANCH_UNCOND if sampling is True else _lib.BRANCH_COND
        eng, xx, spec = self._prepare(x_t, waveform, branches, inpainting_t, inpainting_f)
        if x_t.shape[0] > len(self.betas):
            raise NotImplementedError("fractional diffusion steps need batch <= timesteps (per-roll table rows)")
        eng.set_step_embeddings(rows)
        try:
            pred = eng.step(xx, None, 0, self._learned_upd(_upd(_lib.UPD_NONE), branches))
        finally:
            eng.set_step_embeddings(None)
        if not self._range_ok(eng):
            self._range_fallback()
            return self.forward(x_t, waveform, diffusion_step, sampling, inpainting_t, inpainting_f)
        return pred, spec

    def _forward_training(self, x_t, waveform, diffusion_step, sampling, inpainting_t, inpainting_f, dropout_mask=None):
        """model/diffwave.py:637-686 with ``self.training`` set: spec dropout (:646-647), then the fp32 training forward that
        keeps its activations for ``TrainEngine.backward`` (csrc/train.cu).  No autograd graph is attached to the result."""
        if not torch.is_tensor(diffusion_step) or diffusion_step.dtype not in (torch.int32, torch.int64):
            raise NotImplementedError("fractional diffusion steps (_lerp_embedding) are not built")
        _, _, spec = self._prepare(x_t, waveform, _lib.BRANCH_COND, inpainting_t, inpainting_f, mel_only=True)
        if spec.shape[-1] != x_t.shape[2]:
            raise NotImplementedError("training forward: the clip must cover the whole roll (trim_spec_roll would shorten it)")
        spec = self.uncon_dropout(spec.clone(), self.hparams.spec_dropout, mask=dropout_mask)
        if inpainting_t or inpainting_f:     # the masks were applied before the dropout in the cached spectrogram; re-apply (:649-654)
            t0, t1 = (int(inpainting_t[0]), int(inpainting_t[1])) if inpainting_t else (0, spec.shape[2])
            f0, f1 = (int(inpainting_f[0]), int(inpainting_f[1])) if inpainting_f else (0, spec.shape[1])
            spec[:, f0:f1, t0:t1] = -1
        if sampling is True:                 # model/diffwave.py:656-660 (the reference hands the 2-D table itself to the blocks)
            spec = self._uncond_spec_like(spec)
        eng = self._train_engine(x_t.shape[0], x_t.shape[2])
        return eng.forward(x_t, spec, diffusion_step.flatten()), spec

    def _uncond_spec_like(self, spec):
        """What sampling=True conditions on, in the shape of ``spec`` [B, n_mels, T]: -1, or the learned table."""
        if self._learned:
            return self.trainable_parameters.detach()[..., :spec.shape[-1]].expand_as(spec).contiguous()
        return torch.full_like(spec, -1.0)

    def trainable_dropout(self, x, p, mask=None):
        """model/diffwave.py:695-699: the dropped rolls are conditioned on the learned spectrogram instead of -1.  Same Bernoulli
        draw as ``fixed_dropout``; ``x`` is already trimmed to the roll here, so the table is trimmed alike."""
        if mask is None:
            mask = torch.bernoulli(torch.full((x.shape[0],), float(p)))
        drop = mask.to(device=x.device).bool()
        x[drop] = self.trainable_parameters.detach()[..., :x.shape[-1]]
        return x

    def train(self, mode=True):
        return super().train(mode)

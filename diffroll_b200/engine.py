"""Thin Python owner of a ``drb_plan``: device workspace, weight pointers, stream plumbing.

PyTorch is used here only for device memory and streams; all arithmetic of the
sampling path happens inside libdiffroll_b200.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import DrbConfig, DrbUpdate, DrbWeights


def _stream(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    """One plan = one (batch, frames, wave_len, precision) configuration on one GPU."""

    def __init__(self, state, hp, batch, frames, wave_len, emb_table, precision="f16n4",
                 branches=_lib.BRANCH_COND_UNCOND, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.DrbError("diffroll_b200 needs a CUDA device: there is no CPU path")
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {list(_lib.PRECISIONS)}")
        sa = hp["spec_args"]
        L = hp["residual_layers"]
        cfg = DrbConfig(batch=batch, frames=frames, pitches=88, wave_len=wave_len,
                        residual_channels=hp["residual_channels"], residual_layers=L,
                        kernel_size=hp["kernel_size"], dilation_base=hp["dilation_base"],
                        dilation_bound=hp["dilation_bound"], n_mels=sa["n_mels"], n_fft=sa["n_fft"],
                        hop_length=sa["hop_length"], timesteps=emb_table.shape[0],
                        precision=_lib.PRECISIONS[precision], branches=branches, reserved=0)
        self.cfg = cfg
        self.precision = precision
        self.batch, self.frames, self.wave_len = batch, frames, wave_len
        self.n_mels = sa["n_mels"]
        need = self.lib.drb_plan_workspace_bytes(C.byref(cfg))
        if need == 0:
            raise _lib.DrbError("unsupported configuration: " + self.lib.drb_last_error().decode())
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
            base = self.workspace.data_ptr()
            self._ws_off = (-base) % 256
            # keep every weight tensor alive and contiguous fp32 on this device
            self._w = {k: v.detach().to(device=self.device, dtype=torch.float32).contiguous() for k, v in state.items()}
            w = self._w

            def arr(fmt):
                a = (C.c_void_p * L)(*[w[fmt.format(i)].data_ptr() for i in range(L)])
                self._keep.append(a)
                return C.cast(a, C.POINTER(C.c_void_p))

            self._keep = []
            ws = DrbWeights(
                input_projection_w=w["input_projection.weight"].data_ptr(),
                input_projection_b=w["input_projection.bias"].data_ptr(),
                emb_projection1_w=w["diffusion_embedding.projection1.weight"].data_ptr(),
                emb_projection1_b=w["diffusion_embedding.projection1.bias"].data_ptr(),
                emb_projection2_w=w["diffusion_embedding.projection2.weight"].data_ptr(),
                emb_projection2_b=w["diffusion_embedding.projection2.bias"].data_ptr(),
                dilated_conv_w=arr("residual_layers.{}.dilated_conv.weight"),
                dilated_conv_b=arr("residual_layers.{}.dilated_conv.bias"),
                diffusion_projection_w=arr("residual_layers.{}.diffusion_projection.weight"),
                diffusion_projection_b=arr("residual_layers.{}.diffusion_projection.bias"),
                conditioner_projection_w=arr("residual_layers.{}.conditioner_projection.weight"),
                conditioner_projection_b=arr("residual_layers.{}.conditioner_projection.bias"),
                output_projection_w=arr("residual_layers.{}.output_projection.weight"),
                output_projection_b=arr("residual_layers.{}.output_projection.bias"),
                skip_projection_w=w["skip_projection.weight"].data_ptr(),
                skip_projection_b=w["skip_projection.bias"].data_ptr(),
                head_projection_w=w["output_projection.weight"].data_ptr(),
                head_projection_b=w["output_projection.bias"].data_ptr(),
                stft_window=w["mel_layer.spectrogram.window"].data_ptr(),
                mel_fb=w["mel_layer.mel_scale.fb"].data_ptr(),
            )
            plan = C.c_void_p()
            _lib.check(self.lib.drb_plan_create(C.byref(plan), C.byref(cfg), C.byref(ws),
                                                C.c_void_p(base + self._ws_off), C.c_size_t(need), _stream(self.device)),
                       "drb_plan_create")
            self.plan = plan
            self._emb = emb_table.detach().to(device=self.device, dtype=torch.float32).contiguous()
            _lib.check(self.lib.drb_time_tables(self.plan, _ptr(self._emb), _stream(self.device)), "drb_time_tables")
        self.workspace_bytes = int(need)
        # f16n4 is granted only to shapes that run as CTA pairs; otherwise the plan computes in f16e5 (drb_plan_precision)
        code = int(self.lib.drb_plan_precision(self.plan))
        self.effective_precision = {v: k for k, v in _lib.PRECISIONS.items()}.get(code, precision)
        self.branches = branches

    def close(self):
        if getattr(self, "plan", None):
            self.lib.drb_plan_destroy(self.plan)
            self.plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------
    def set_branches(self, branches):
        if branches != self.branches:
            _lib.check(self.lib.drb_plan_set_branches(self.plan, branches), "drb_plan_set_branches")
            self.branches = branches

    def set_uncond_spec(self, table):
        """condition='trainable_spec' (model/diffwave.py:601-604,657-658): ``table`` [n_mels, >= frames] is the learned
        unconditional spectrogram; the plan (created with BRANCH_COND_LEARNED) keeps its own copy."""
        t = table.detach().to(device=self.device, dtype=torch.float32).contiguous()
        if t.ndim != 2 or t.shape[0] != self.n_mels or t.shape[1] < self.frames:
            raise ValueError(f"learned spectrogram: expected [{self.n_mels}, >= {self.frames}], got {tuple(t.shape)}")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.drb_plan_set_uncond_spec(self.plan, _ptr(t), int(t.shape[1]), _stream(self.device)),
                       "drb_plan_set_uncond_spec")

    def set_step_embeddings(self, emb_rows):
        """Fractional diffusion steps (model/diffwave.py:76-81): ``emb_rows`` [B,128] = the caller's interpolated sinusoid rows, one
        per roll; ``None`` returns to the integer tables."""
        if emb_rows is None:
            self._emb_rows = None
            _lib.check(self.lib.drb_plan_set_step_embeddings(self.plan, C.c_void_p(0), _stream(self.device)), "drb_plan_set_step_embeddings")
            return
        e = emb_rows.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(e.shape) != (self.batch, 128):
            raise ValueError(f"step embeddings: expected shape ({self.batch}, 128), got {tuple(e.shape)}")
        self._emb_rows = e
        with torch.cuda.device(self.device):
            _lib.check(self.lib.drb_plan_set_step_embeddings(self.plan, _ptr(e), _stream(self.device)), "drb_plan_set_step_embeddings")

    def set_steps(self, steps):
        """Per-sample diffusion steps for the following step()/forward calls (model/diffwave.py:637,670: ``diffusion_step``
        is int64[B]); ``None`` returns to the uniform ``t_index`` argument.  The int32 device copy is kept alive here."""
        if steps is None:
            self._steps = None
            _lib.check(self.lib.drb_plan_set_steps(self.plan, C.c_void_p(0)), "drb_plan_set_steps")
            return
        st = steps.to(device=self.device, dtype=torch.int32).contiguous()
        if tuple(st.shape) != (self.batch,):
            raise ValueError(f"diffusion_step: expected shape ({self.batch},), got {tuple(st.shape)}")
        self._steps = st
        _lib.check(self.lib.drb_plan_set_steps(self.plan, _ptr(st)), "drb_plan_set_steps")

    def _check(self, t, shape, name):
        if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError(f"{name}: expected a contiguous fp32 tensor on {self.device}")
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")

    def mel(self, waveform, inpainting_t=None, inpainting_f=None, want_spec=True):
        """model/diffwave.py:643-654: returns the normalised, masked spectrogram [B, n_mels, T]."""
        self._check(waveform, (self.batch, self.wave_len), "waveform")
        nF = self.wave_len // self.cfg.hop_length + 1
        it0 = it1 = if0 = if1 = 0
        if inpainting_t:
            it0, it1, _ = slice(int(inpainting_t[0]), int(inpainting_t[1])).indices(nF)
            it1 = max(it1, it0)
        if inpainting_f:
            if0, if1, _ = slice(int(inpainting_f[0]), int(inpainting_f[1])).indices(self.n_mels)
            if1 = max(if1, if0)
        # The library reads it1 > it0 / if1 > if0 as "this mask was requested".  A requested-but-empty range masks
        # nothing, like the reference's empty slice assignment (model/diffwave.py:649-654) -- and when BOTH ranges were
        # requested the masked region is their product (:653), so one empty range empties the whole mask.
        if (inpainting_t and it1 == it0) or (inpainting_f and if1 == if0):
            it0 = it1 = if0 = if1 = 0
        with torch.cuda.device(self.device):
            spec = torch.empty(self.batch, self.n_mels, self.frames, device=self.device) if want_spec else None
            _lib.check(self.lib.drb_mel_forward(self.plan, _ptr(waveform), _ptr(spec), it0, it1, if0, if1,
                                                _stream(self.device)), "drb_mel_forward")
        return spec

    def step(self, x_t, noise, t_index, upd: DrbUpdate, out=None, net_out=None):
        """One reverse-diffusion step: in_proj, residual blocks, head, guidance, posterior update."""
        shape = (self.batch, 1, self.frames, 88)
        self._check(x_t, shape, "x_t")
        if noise is not None:
            self._check(noise, shape, "noise")
        if out is None:
            out = torch.empty_like(x_t)
        with torch.cuda.device(self.device):
            return self._step_on_device(x_t, noise, t_index, upd, out, net_out)

    def _step_on_device(self, x_t, noise, t_index, upd, out, net_out):
        s = _stream(self.device)
        # The per-clip conditioner tables cost about one forward.  A sampler step is one of a chain over the same clip
        # (task/diffusion.py:528-529): build them (no-op when ready) and use them.  A plain forward (UPD_NONE: forward(),
        # the validation step) runs without them, whether or not they happen to be ready: results never depend on history.
        sampler_step = upd.mode != _lib.UPD_NONE and self.branches != _lib.BRANCH_UNCOND
        if sampler_step:
            _lib.check(self.lib.drb_cond_tables(self.plan, s), "drb_cond_tables")
        _lib.check(self.lib.drb_plan_use_cond_tables(self.plan, 1 if sampler_step else 0), "drb_plan_use_cond_tables")
        if net_out is None:
            _lib.check(self.lib.drb_sample_step(self.plan, _ptr(x_t), _ptr(noise), _ptr(out), int(t_index),
                                                C.byref(upd), s), "drb_sample_step")
        else:
            _lib.check(self.lib.drb_in_proj(self.plan, _ptr(x_t), int(t_index), s), "drb_in_proj")
            for layer in range(self.cfg.residual_layers):
                _lib.check(self.lib.drb_resblock_forward(self.plan, layer, int(t_index), s), "drb_resblock_forward")
            _lib.check(self.lib.drb_head_posterior_step(self.plan, _ptr(x_t), _ptr(noise), _ptr(out), _ptr(net_out),
                                                        C.byref(upd), s), "drb_head_posterior_step")
        return out

    def loop(self, x, noise, updates, t_start, t_stop=0, trajectory=None):
        """task/diffusion.py:528-534 for t = t_start-1 .. t_stop, x updated in place."""
        self._check(x, (self.batch, 1, self.frames, 88), "x")
        n = t_start - t_stop
        arr = (DrbUpdate * n)(*updates)
        n_noise = sum(1 for u in updates if u.has_noise)
        if n_noise:
            self._check(noise, (noise.shape[0], self.batch, 1, self.frames, 88), "noise")
            if noise.shape[0] < n_noise:
                raise ValueError(f"need {n_noise} noise slices, got {noise.shape[0]}")
        tp = C.c_void_p(trajectory.data_ptr()) if trajectory is not None else C.c_void_p(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.drb_sample_loop(self.plan, _ptr(x), _ptr(noise) if n_noise else C.c_void_p(0), arr,
                                                int(t_start), int(t_stop), tp, _stream(self.device)), "drb_sample_loop")
        return x

    def range_max(self, reset=True):
        """Largest |activation operand| the fp16-based kernels emitted since the last reset (synchronises)."""
        v = C.c_float()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.drb_plan_range_stats(self.plan, C.byref(v), 1 if reset else 0, _stream(self.device)),
                       "drb_plan_range_stats")
        return float(v.value)

    def buffer(self, name, dtype=torch.float32):
        """Debug view of a plan-owned device buffer."""
        p = C.c_void_p()
        nbytes = C.c_size_t()
        _lib.check(self.lib.drb_plan_buffer(self.plan, name.encode(), C.byref(p), C.byref(nbytes)), "drb_plan_buffer")
        off = p.value - self.workspace.data_ptr()
        return self.workspace[off:off + nbytes.value].view(dtype)

    def profile(self, enable=True):
        _lib.check(self.lib.drb_plan_profile(self.plan, 1 if enable else 0), "drb_plan_profile")

    def profile_read(self):
        """{class: (total_ms, spans)} for gate / out / in_proj / head kernels since profile(True)."""
        ms = (C.c_double * 4)()
        n = (C.c_int64 * 4)()
        _lib.check(self.lib.drb_plan_profile_read(self.plan, ms, n), "drb_plan_profile_read")
        return {k: (ms[i], int(n[i])) for i, k in enumerate(("gate", "out", "in_proj", "head"))}

    def profile_read_detail(self):
        """({class: (total_ms, spans)} incl. 'head_out' = the head's projection + posterior kernel alone, gate ms per layer)."""
        L = self.cfg.residual_layers
        ms = (C.c_double * 5)()
        n = (C.c_int64 * 5)()
        per = (C.c_double * L)()
        _lib.check(self.lib.drb_plan_profile_read2(self.plan, ms, n, 5, per, L), "drb_plan_profile_read2")
        names = ("gate", "out", "in_proj", "head", "head_out")
        return {k: (ms[i], int(n[i])) for i, k in enumerate(names)}, [per[i] for i in range(L)]

    def launch_count(self, reset=False):
        return int(self.lib.drb_launch_count(1 if reset else 0))

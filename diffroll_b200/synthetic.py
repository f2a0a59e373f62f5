"""Deterministic synthetic weights and inputs for tests and benchmarks.

There is no network for checkpoints or datasets, so every parity test and the
benchmark use tensors generated here.  The generator is independent of module
construction order: each tensor of the reference ``state_dict`` (key set from
SURVEY.md §8b, probed on model/diffwave.py:579-635) is drawn from one CPU
``torch.Generator`` in sorted-key order, so the reference model, the oracle and
the CUDA path can all be loaded with bit-identical weights.

The reference zero-initialises ``output_projection.weight``
(model/diffwave.py:630), which makes the network output independent of its
input; a parity test on such weights is vacuous, so the head is drawn from
N(0, 1/C) here (SURVEY.md §8c "mandatory harness fix").
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

MEL_ARGS = dict(  # config/spec/mel.yaml:2-10 with sampling_rate/hop_length of config/sampling.yaml:2-4
    sample_rate=16000, n_fft=2048, hop_length=512, n_mels=229,
    f_min=0, f_max=8000, center=True, normalized=True, pad_mode="reflect",
)


def default_hparams(timesteps=200, sampling_type="inpainting_ddpm_x0", w=0.5,
                    inpainting_t=None, inpainting_f=None, residual_channels=512,
                    residual_layers=15, kernel_size=9, condition="fixed"):
    """Hyper-parameters of configs[1] (config/model/ClassifierFreeDiffRoll.yaml + config/task/transcription.yaml, k=9)."""
    return dict(
        residual_channels=residual_channels, unconditional=False, condition=condition,
        n_mels=229, norm_args=[0, 1, "imagewise"], residual_layers=residual_layers,
        kernel_size=kernel_size, dilation_base=2, dilation_bound=4,
        spec_args=dict(MEL_ARGS), spec_dropout=0.1,
        inpainting_t=inpainting_t, inpainting_f=inpainting_f,
        lr=1e-4, timesteps=timesteps, loss_type="l2", loss_keys=["diffusion_loss"],
        beta_start=1e-4, beta_end=0.02, frame_threshold=0.5,
        training=dict(mode="x_0"), sampling=dict(type=sampling_type, w=w),
        debug=False, generation_filter=0.02,
    )


def hann_window(n_fft=2048):
    return torch.hann_window(n_fft, periodic=True, dtype=torch.float32)


def melscale_fbanks(n_freqs=1025, f_min=0.0, f_max=8000.0, n_mels=229, sample_rate=16000):
    """HTK mel filterbank, norm=None, as torchaudio builds it for MelScale
    (torchaudio/functional/functional.py `melscale_fbanks`; third-party, pinned
    torchaudio==0.11.0 in the reference's requirements.txt:12).  Returns [n_freqs, n_mels] fp32."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    zero = torch.zeros(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    return torch.max(zero, torch.min(down_slopes, up_slopes))


def state_dict_shapes(hp):
    C = hp["residual_channels"]; L = hp["residual_layers"]; k = hp["kernel_size"]; M = hp["n_mels"]
    n_fft = hp["spec_args"]["n_fft"]
    s = OrderedDict()
    s["input_projection.weight"] = (C, 88, 1); s["input_projection.bias"] = (C,)
    s["diffusion_embedding.projection1.weight"] = (512, 128); s["diffusion_embedding.projection1.bias"] = (512,)
    s["diffusion_embedding.projection2.weight"] = (512, 512); s["diffusion_embedding.projection2.bias"] = (512,)
    for i in range(L):
        p = f"residual_layers.{i}."
        s[p + "dilated_conv.weight"] = (2 * C, C, k); s[p + "dilated_conv.bias"] = (2 * C,)
        s[p + "diffusion_projection.weight"] = (C, 512); s[p + "diffusion_projection.bias"] = (C,)
        if not hp["unconditional"]:
            s[p + "conditioner_projection.weight"] = (2 * C, M, 1); s[p + "conditioner_projection.bias"] = (2 * C,)
        s[p + "output_projection.weight"] = (2 * C, C, 1); s[p + "output_projection.bias"] = (2 * C,)
    s["skip_projection.weight"] = (C, C, 1); s["skip_projection.bias"] = (C,)
    s["output_projection.weight"] = (88, C, 1); s["output_projection.bias"] = (88,)
    s["mel_layer.spectrogram.window"] = (n_fft,)
    s["mel_layer.mel_scale.fb"] = (n_fft // 2 + 1, M)
    if hp.get("condition", "fixed") == "trainable_spec":
        s["trainable_parameters"] = (hp["spec_args"]["n_mels"], 641)   # model/diffwave.py:601 (641 is hard-coded there)
    return s


def make_state_dict(hp, seed=0):
    """Reference-shaped ``state_dict`` with seeded values (kaiming-normal-like
    weights, small uniform biases, non-zero output head)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    shapes = state_dict_shapes(hp)
    for key in sorted(shapes):
        shp = shapes[key]
        if key == "mel_layer.spectrogram.window":
            sd[key] = hann_window(shp[0]); continue
        if key == "mel_layer.mel_scale.fb":
            sa = hp["spec_args"]
            sd[key] = melscale_fbanks(shp[0], float(sa["f_min"]), float(sa["f_max"]), sa["n_mels"], sa["sample_rate"]); continue
        if key == "trainable_parameters":   # the reference starts it at -1 (= the fixed variant); a trained one is anything in [-1, 1]
            sd[key] = torch.rand(shp, generator=g) * 1.6 - 1.0; continue
        fan_in = 1
        for d in shp[1:]:
            fan_in *= d
        if key.endswith(".weight"):
            if key == "output_projection.weight":
                std = fan_in ** -0.5
            elif "projection1" in key or "projection2" in key or "diffusion_projection" in key:
                std = (1.0 / fan_in) ** 0.5
            else:
                std = (2.0 / fan_in) ** 0.5  # nn.init.kaiming_normal_ (model/diffwave.py:41-44)
            sd[key] = torch.randn(shp, generator=g) * std
        else:
            sd[key] = (torch.rand(shp, generator=g) * 2 - 1) * 0.05
    return OrderedDict((k, sd[k]) for k in shapes)


def make_inputs(batch, timesteps, seed=123, n_noise=None, T=640, wav_len=327680):
    """x_T, waveform and the pre-drawn per-step posterior noise, one generator,
    drawn in the order the reference consumes them (sampling.py:27,45; task/diffusion.py:1023)."""
    g = torch.Generator().manual_seed(seed)
    x_T = torch.randn(batch, 1, T, 88, generator=g)
    waveform = torch.randn(batch, wav_len, generator=g)
    n = timesteps - 1 if n_noise is None else n_noise
    noise = torch.randn(n, batch, 1, T, 88, generator=g) if n > 0 else torch.empty(0, batch, 1, T, 88)
    return x_T, waveform, noise


def make_labelled_batch(B=4, T=128, wav_len=65536, seed=77):
    """Synthetic labelled batch of the (validation) step, task/diffusion.py:651-670: a binary piano roll [B,T,88] (the
    last roll empty: the Normalization NaN -> min case of model/utils.py:31), a waveform, one diffusion step per roll
    and the label noise [B,1,T,88].  Returns (frame, audio, t, noise)."""
    g = torch.Generator().manual_seed(seed)
    frame = (torch.rand(B, T, 88, generator=g) < 0.06).float()
    frame[B - 1] = 0.0
    audio = torch.randn(B, wav_len, generator=g)
    t = torch.tensor([3, 150, 0, 199, 77, 12, 181, 64][:B])
    noise = torch.randn(B, 1, T, 88, generator=g)
    return frame, audio, t, noise

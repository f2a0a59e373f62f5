"""Generate tests/golden/*.npz by running the UNMODIFIED reference (container only).

    python oracle/make_golden.py [--only forward,steps,chain,chain_inpaint,chain_gen,valstep,learned]

The reference is imported from /root/reference through oracle/ref_shim.py, loaded
with the seeded weights of diffroll_b200/synthetic.py (strict=True, so the
132-key state_dict contract of SURVEY.md §8b is checked too) and driven through
its own ``forward`` / sampler methods / ``predict_step``-style loop
(task/diffusion.py:528-534).  ``torch.randn_like`` is replaced by a queue of
pre-drawn tensors for the duration of each sampler call so the CUDA path can be
fed the same noise (a CPU and a CUDA generator can never share a stream).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from diffroll_b200.synthetic import default_hparams, make_inputs, make_labelled_batch, make_state_dict  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class NoiseQueue:
    def __init__(self):
        self.q = []
        self._orig = torch.randn_like

    def __enter__(self):
        torch.randn_like = lambda x, *a, **k: self.q.pop(0).to(x.dtype)
        return self

    def __exit__(self, *a):
        torch.randn_like = self._orig


def ref_model(hp, dtype=torch.float32):
    m = ref_shim.build_reference_model(hp)
    missing = m.load_state_dict(make_state_dict(hp), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    if dtype == torch.float64:
        m = m.double()
        m.diffusion_embedding.embedding = m.diffusion_embedding.embedding.double()
    return m.eval()


@torch.no_grad()
def run_chain(m, x_T, waveform, noise, keep=()):
    """The loop body of predict_step (task/diffusion.py:528-534) with injected noise."""
    x = x_T
    kept = {}
    with NoiseQueue() as nq:
        nq.q = [n for n in noise]
        for t_index in reversed(range(0, m.hparams.timesteps)):
            x, spec = m.reverse_diffusion(x, waveform, t_index)
            _ = x.detach().cpu().numpy()
            if t_index in keep:
                kept[t_index] = x.clone()
    return x, spec, kept


def gen_forward():
    hp = default_hparams()
    m = ref_model(hp)
    x_T, wav, _ = make_inputs(2, 200, seed=123, n_noise=0)
    t = torch.tensor(37).repeat(2)
    with torch.no_grad():
        pred_c, spec_c = m(x_T, wav, t)
        pred_u, spec_u = m(x_T, torch.zeros_like(wav), t, sampling=True)
        pred_m, spec_m = m(x_T, wav, t, inpainting_t=[100, 420], inpainting_f=None)
        pred_f, spec_f = m(x_T, wav, t, inpainting_t=[100, 420], inpainting_f=[30, 99])
    np.savez(os.path.join(GOLD, "forward_b2_t37.npz"),
             pred_c=pred_c.numpy(), spec_c=spec_c.numpy(), pred_u=pred_u.numpy(),
             pred_m=pred_m.numpy(), spec_m=spec_m[:, :, ::8].numpy(), pred_f=pred_f.numpy())


def gen_steps():
    """One step of every sampler, short clip (T=128, L=65536) to keep fixtures small."""
    out = {}
    for name in ["inpainting_ddpm_x0", "cfdg_ddpm_x0", "generation_ddpm_x0", "ddpm_x0", "ddim_x0",
                 "cfdg_ddim_x0", "ddpm", "ddim", "ddim2ddpm"]:
        hp = default_hparams(sampling_type=name, inpainting_t=[32, 96] if name == "inpainting_ddpm_x0" else None)
        m = ref_model(hp)
        x_T, wav, noise = make_inputs(2, 200, seed=7, n_noise=1, T=128, wav_len=65536)
        for t_index in (199, 1, 0):
            with NoiseQueue() as nq, torch.no_grad():
                nq.q = [noise[0]]
                x_prev, _ = m.reverse_diffusion(x_T, wav, t_index)
            out[f"{name}_t{t_index}"] = x_prev.numpy()
    np.savez(os.path.join(GOLD, "steps_T128.npz"), **out)


def gen_chain(tag, hp, B, seed, T=640, wav_len=327680, keep=(), fp64=False):
    x_T, wav, noise = make_inputs(B, hp["timesteps"], seed=seed, T=T, wav_len=wav_len)
    m = ref_model(hp)
    t0 = time.time()
    x0, spec, kept = run_chain(m, x_T, wav, noise, keep)
    dt = time.time() - t0
    out = dict(final=x0.numpy(), seconds=np.float64(dt))
    for t, v in kept.items():
        out[f"t{t}"] = v.numpy()
    if fp64:
        m64 = ref_model(hp, torch.float64)
        x64, _, _ = run_chain(m64, x_T.double(), wav.double(), noise.double())
        out["final_fp64"] = x64.numpy().astype(np.float32)
        out["fp32_vs_fp64_maxabs"] = np.float64((x64 - x0.double()).abs().max().item())
    np.savez(os.path.join(GOLD, f"chain_{tag}.npz"), **out)
    print(tag, "done in", dt, "s", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


def gen_valstep():
    """SpecRollDiffusion.step (task/diffusion.py:651-763) in eval mode, i.e. what validation_step (:271-276) runs:
    normalise, q_sample at per-roll steps, forward, loss -- for the three training modes and the three loss types,
    plus the two-dataset variant (:707-719).  torch.randint / randn_like are pinned to the injected t / noise."""
    frame, audio, t, noise = make_labelled_batch()
    frame2, audio2, _, _ = make_labelled_batch(seed=78)
    out = {}
    orig_randint = torch.randint
    for mode, loss_type in (("x_0", "l2"), ("x_0", "l1"), ("epsilon", "huber"), ("ex_0", "l2")):
        hp = default_hparams()
        hp["training"] = dict(mode=mode)
        hp["loss_type"] = loss_type
        m = ref_model(hp)
        torch.randint = lambda *a, **k: t.clone()
        try:
            with NoiseQueue() as nq, torch.no_grad():
                nq.q = [noise.clone()]
                losses, tensors = m.step({"frame": frame.clone(), "audio": audio.clone()})
        finally:
            torch.randint = orig_randint
        tag = f"{mode}_{loss_type}"
        out[f"{tag}_loss"] = np.float64(losses["diffusion_loss"].item())
        out[f"{tag}_pred_roll"] = tensors["pred_roll"].numpy()
        out["label_roll"] = tensors["label_roll"].numpy()      # the same normalised label for every mode
    hp = default_hparams()
    m = ref_model(hp)
    torch.randint = lambda *a, **k: t.clone()
    try:
        with NoiseQueue() as nq, torch.no_grad():
            nq.q = [noise.clone()]
            losses, tensors = m.step([{"frame": frame.clone(), "audio": audio.clone()},
                                      {"frame": frame2.clone(), "audio": audio2.clone()}])
    finally:
        torch.randint = orig_randint
    out["two_loss"] = np.float64(losses["diffusion_loss"].item())
    out["two_uncond_loss"] = np.float64(losses["unconditional_diffusion_loss"].item())
    out["two_pred_roll2"] = tensors["pred_roll2"].numpy()
    np.savez_compressed(os.path.join(GOLD, "valstep_b4_T128.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items()})


def gen_learned():
    """condition='trainable_spec' (model/diffwave.py:600-605, 657-658, 695-699): the 133-key state_dict loads strict=True; one
    sampling=True forward, single steps of the two samplers that read the learned table, the validation step's second
    (unconditional) dataset and a whole 200-step guided chain, short clip."""
    out = {}
    x_T, wav, noise = make_inputs(2, 200, seed=7, n_noise=1, T=128, wav_len=65536)
    for name in ("cfdg_ddpm_x0", "generation_ddpm_x0"):
        hp = default_hparams(sampling_type=name, condition="trainable_spec")
        m = ref_model(hp)
        for t_index in (199, 1, 0):
            with NoiseQueue() as nq, torch.no_grad():
                nq.q = [noise[0]]
                x_prev, _ = m.reverse_diffusion(x_T, wav, t_index)
            out[f"{name}_t{t_index}"] = x_prev.numpy()
    with torch.no_grad():
        pred_u, spec_u = m(x_T, wav, torch.tensor([37, 150]), sampling=True)
    assert spec_u.shape == (229, 128)
    out["pred_u"] = pred_u.numpy()
    frame, audio, t, nz = make_labelled_batch(B=2)
    frame2, audio2, _, _ = make_labelled_batch(B=2, seed=78)
    orig_randint = torch.randint
    torch.randint = lambda *a, **k: t.clone()
    try:
        with NoiseQueue() as nq, torch.no_grad():
            nq.q = [nz.clone()]
            losses, tensors = m.step([{"frame": frame.clone(), "audio": audio.clone()}, {"frame": frame2.clone(), "audio": audio2.clone()}])
    finally:
        torch.randint = orig_randint
    out["two_loss"] = np.float64(losses["diffusion_loss"].item())
    out["two_uncond_loss"] = np.float64(losses["unconditional_diffusion_loss"].item())
    out["two_pred_roll2"] = tensors["pred_roll2"].numpy()
    hp = default_hparams(sampling_type="cfdg_ddpm_x0", condition="trainable_spec")
    xc, wc, nc = make_inputs(2, 200, seed=13, T=128, wav_len=65536)
    x0, _, _ = run_chain(ref_model(hp), xc, wc, nc)
    out["chain_final"] = x0.numpy()
    np.savez_compressed(os.path.join(GOLD, "learned_T128.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items()})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="forward,steps,chain,chain_inpaint,chain_gen,valstep")
    a = ap.parse_args()
    only = set(a.only.split(","))
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    if "forward" in only:
        gen_forward(); print("forward done")
    if "steps" in only:
        gen_steps(); print("steps done")
    if "valstep" in only:
        gen_valstep(); print("valstep done")
    if "learned" in only:
        gen_learned(); print("learned done")
    if "chain" in only:       # configs[0]/[1] shape: transcription, 200 steps, full 640-frame clip
        gen_chain("transcription_b1_200", default_hparams(), 1, seed=123, keep=(150, 100, 50), fp64=True)
    if "chain_inpaint" in only:  # configs[3] shape: 50 % frame mask
        gen_chain("inpaint_b2_200_T128", default_hparams(inpainting_t=[0, 64]), 2, seed=11, T=128, wav_len=65536, keep=(180,))
    if "chain_gen" in only:   # configs[2] shape: unconditional, 1000 steps (short clip)
        gen_chain("generation_b1_1000_T128", default_hparams(timesteps=1000, sampling_type="generation_ddpm_x0"),
                  1, seed=5, T=128, wav_len=65536, keep=(960,))


if __name__ == "__main__":
    main()

"""The CPU oracle (oracle/diffroll_oracle.py) against the golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  This is what pins the oracle; it runs without a GPU."""
import numpy as np
import pytest
import torch

from conftest import golden
from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict
from oracle.diffroll_oracle import OracleDiffRoll

# The oracle restates the reference op for op on the same torch CPU kernels, so it should agree to
# rounding noise; 2e-5 leaves room for thread-count dependent reduction order on another host.
TOL = 2e-5


def _oracle(**kw):
    hp = default_hparams(**kw)
    return OracleDiffRoll(hp, make_state_dict(hp)), hp


def _err(a, b):
    return float(np.abs(a.numpy().astype(np.float64) - b.astype(np.float64)).max())


def test_forward_cond_uncond_masks():
    g = golden("forward_b2_t37.npz")
    o, _ = _oracle()
    x_T, wav, _ = make_inputs(2, 200, seed=123, n_noise=0)
    t = torch.tensor(37).repeat(2)
    with torch.no_grad():
        pc, sc = o(x_T, wav, t)
        pu, su = o(x_T, torch.zeros_like(wav), t, sampling=True)
        pm, sm = o(x_T, wav, t, inpainting_t=[100, 420])
        pf, _ = o(x_T, wav, t, inpainting_t=[100, 420], inpainting_f=[30, 99])
    assert pc.shape == (2, 1, 640, 88) and sc.shape == (2, 229, 640)
    assert _err(pc, g["pred_c"]) < TOL and _err(sc, g["spec_c"]) < TOL
    assert _err(pu, g["pred_u"]) < TOL and float(su.max()) == -1.0
    assert _err(pm, g["pred_m"]) < TOL and _err(sm[:, :, ::8], g["spec_m"]) < TOL
    assert _err(pf, g["pred_f"]) < TOL
    # a vacuous comparison would pass on a constant output: the golden must vary along time
    assert g["pred_c"].std(axis=2).min() > 1e-3


@pytest.mark.parametrize("name", ["inpainting_ddpm_x0", "cfdg_ddpm_x0", "generation_ddpm_x0", "ddpm_x0", "ddim_x0",
                                  "cfdg_ddim_x0", "ddpm", "ddim", "ddim2ddpm"])
def test_every_sampler_single_step(name):
    g = golden("steps_T128.npz")
    o, _ = _oracle(sampling_type=name, inpainting_t=[32, 96] if name == "inpainting_ddpm_x0" else None)
    x_T, wav, noise = make_inputs(2, 200, seed=7, n_noise=1, T=128, wav_len=65536)
    for t_index in (199, 1, 0):
        with torch.no_grad():
            x_prev, _ = o.reverse_diffusion(x_T, wav, t_index, noise=noise[0])
        ref = g[f"{name}_t{t_index}"]
        assert _err(x_prev, ref) < TOL * max(1.0, float(np.abs(ref).max())), (name, t_index)


def test_chain_inpainting_first_20_steps():
    g = golden("chain_inpaint_b2_200_T128.npz")
    if "t180" not in g.files:
        pytest.skip("fixture has no t180 checkpoint")
    o, hp = _oracle(inpainting_t=[0, 64])
    x_T, wav, noise = make_inputs(2, 200, seed=11, T=128, wav_len=65536)
    x, spec, _ = o.sample_loop(x_T, wav, noise, t_stop=180)
    assert _err(x, g["t180"]) < 5 * TOL
    assert float(spec[:, :, :64].max()) == -1.0


def test_chain_generation_1000_first_40_steps():
    g = golden("chain_generation_b1_1000_T128.npz")
    if "t960" not in g.files:
        pytest.skip("fixture has no t960 checkpoint")
    o, hp = _oracle(timesteps=1000, sampling_type="generation_ddpm_x0")
    x_T, wav, noise = make_inputs(1, 1000, seed=5, T=128, wav_len=65536)
    x, _, _ = o.sample_loop(x_T, wav, noise, t_stop=960)
    assert _err(x, g["t960"]) < 5 * TOL


@pytest.mark.slow
def test_chain_transcription_first_50_steps():
    g = golden("chain_transcription_b1_200.npz")
    o, hp = _oracle()
    x_T, wav, noise = make_inputs(1, 200, seed=123)
    x, _, _ = o.sample_loop(x_T, wav, noise, t_stop=150)
    assert _err(x, g["t150"]) < 5 * TOL


def test_golden_chain_records_reference_noise_floor():
    g = golden("chain_transcription_b1_200.npz")
    assert g["final"].shape == (1, 1, 640, 88)
    assert float(g["fp32_vs_fp64_maxabs"]) < 1e-4      # reference fp32 vs its own fp64 run: ~4e-6
    assert float(np.abs(g["final"] - g["final_fp64"]).max()) < 1e-4


def test_mel_restatement_matches_torchaudio():
    """The oracle's mel front-end against torchaudio.transforms.MelSpectrogram (the third-party code the
    reference calls, model/diffwave.py:635,643) when torchaudio is importable."""
    torchaudio = pytest.importorskip("torchaudio")
    from diffroll_b200.synthetic import MEL_ARGS, hann_window, melscale_fbanks
    from oracle.diffroll_oracle import mel_spectrogram
    layer = torchaudio.transforms.MelSpectrogram(**MEL_ARGS)
    assert torch.equal(layer.spectrogram.window, hann_window(2048))
    assert torch.allclose(layer.mel_scale.fb, melscale_fbanks(), atol=0, rtol=0)
    wav = torch.randn(2, 32768, generator=torch.Generator().manual_seed(1))
    a = layer(wav)
    b = mel_spectrogram(wav, layer.spectrogram.window, layer.mel_scale.fb)
    assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)
    fb = layer.mel_scale.fb
    assert int((fb != 0).sum()) == 2034 and int((fb != 0).sum(0).max()) <= 24


VALSTEP_CASES = [("x_0", "l2"), ("x_0", "l1"), ("epsilon", "huber"), ("ex_0", "l2")]


@pytest.mark.parametrize("mode,loss_type", VALSTEP_CASES)
def test_validation_step_vs_reference(mode, loss_type):
    """SpecRollDiffusion.step in eval mode (task/diffusion.py:651-763): per-roll diffusion steps, q_sample, forward, loss."""
    from diffroll_b200.synthetic import make_labelled_batch
    g = golden("valstep_b4_T128.npz")
    hp = default_hparams()
    hp["training"] = dict(mode=mode)
    hp["loss_type"] = loss_type
    o = OracleDiffRoll(hp, make_state_dict(hp))
    frame, audio, t, noise = make_labelled_batch()
    losses, tensors = o.step({"frame": frame, "audio": audio}, t, noise)
    tag = f"{mode}_{loss_type}"
    ref = g[f"{tag}_pred_roll"]
    assert _err(tensors["pred_roll"], ref) < TOL * max(1.0, float(np.abs(ref).max()))
    assert _err(tensors["label_roll"], g["label_roll"]) == 0.0
    assert float(tensors["label_roll"][-1].abs().max()) == 0.0          # empty roll: NaN -> min (model/utils.py:31)
    assert abs(float(losses["diffusion_loss"]) - float(g[f"{tag}_loss"])) < 1e-5 * max(1.0, float(g[f"{tag}_loss"]))


def test_validation_step_two_datasets_vs_reference():
    from diffroll_b200.synthetic import make_labelled_batch
    g = golden("valstep_b4_T128.npz")
    hp = default_hparams()
    o = OracleDiffRoll(hp, make_state_dict(hp))
    frame, audio, t, noise = make_labelled_batch()
    frame2, audio2, _, _ = make_labelled_batch(seed=78)
    losses, tensors = o.step([{"frame": frame, "audio": audio}, {"frame": frame2, "audio": audio2}], t, noise)
    assert abs(float(losses["diffusion_loss"]) - float(g["two_loss"])) < 1e-5 * max(1.0, float(g["two_loss"]))
    assert abs(float(losses["unconditional_diffusion_loss"]) - float(g["two_uncond_loss"])) < 1e-5 * max(1.0, float(g["two_uncond_loss"]))
    assert _err(tensors["pred_roll2"], g["two_pred_roll2"]) < TOL * max(1.0, float(np.abs(g["two_pred_roll2"]).max()))


def test_trainable_spec_vs_golden():
    """condition='trainable_spec' (model/diffwave.py:600-605, 657-658): the oracle's learned-table branch against the unmodified
    reference -- single steps of the two samplers that read the table, a sampling=True forward at per-roll steps, the validation
    step's unconditional second dataset."""
    g = golden("learned_T128.npz")
    x_T, wav, noise = make_inputs(2, 200, seed=7, n_noise=1, T=128, wav_len=65536)
    for name in ("cfdg_ddpm_x0", "generation_ddpm_x0"):
        o, _ = _oracle(sampling_type=name, condition="trainable_spec")
        for t_index in (199, 1, 0):
            with torch.no_grad():
                x_prev, _ = o.reverse_diffusion(x_T, wav, t_index, noise=noise[0])
            ref = g[f"{name}_t{t_index}"]
            assert _err(x_prev, ref) < TOL * max(1.0, float(np.abs(ref).max())), (name, t_index)
    with torch.no_grad():
        pu, su = o(x_T, wav, torch.tensor([37, 150]), sampling=True)
    assert su.shape == (229, 128) and _err(pu, g["pred_u"]) < TOL
    from diffroll_b200.synthetic import make_labelled_batch
    frame, audio, t, nz = make_labelled_batch(B=2)
    frame2, audio2, _, _ = make_labelled_batch(B=2, seed=78)
    losses, tensors = o.step([{"frame": frame, "audio": audio}, {"frame": frame2, "audio": audio2}], t, nz)
    assert abs(float(losses["unconditional_diffusion_loss"]) - float(g["two_uncond_loss"])) < 1e-5
    assert _err(tensors["pred_roll2"], g["two_pred_roll2"]) < TOL
    # the table matters: the fixed variant's unconditional forward is somewhere else
    f, _ = _oracle(sampling_type="generation_ddpm_x0")
    with torch.no_grad():
        pf, _ = f(x_T, wav, torch.tensor([37, 150]), sampling=True)
    assert _err(pf, g["pred_u"]) > 1e-3

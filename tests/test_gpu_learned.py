"""GPU parity of condition='trainable_spec' (model/diffwave.py:600-605, 657-658, 695-699): the unconditional branch is
conditioned on a learned [n_mels, 641] spectrogram instead of -1.  The CUDA path (DRB_BRANCH_COND_LEARNED: the learned clips
sit behind the real ones, every roll reads a conditioner table) against golden vectors of the unmodified reference
(oracle/make_golden.py gen_learned) and against the oracle on the same inputs.

Tolerances as in tests/test_gpu_parity.py: one forward / sampler step 2e-4 (fp32 path) or 5e-4 (tensor path), a whole chain 1e-3.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import golden
from diffroll_b200.synthetic import default_hparams, make_inputs, make_labelled_batch, make_state_dict

pytestmark = pytest.mark.gpu
TOL_STEP = {"fp32": 2e-4, "bf16x3": 5e-4, "f16e5": 5e-4, "f16n4": 5e-4}
TOL_FINAL = 1e-3


def _model(precision, name):
    import diffroll_b200 as M
    hp = default_hparams(sampling_type=name, condition="trainable_spec")
    m = M.ClassifierFreeDiffRoll(**hp, precision=precision)
    m.load_state_dict(make_state_dict(hp), strict=True)
    return m.cuda().eval(), hp


def _close(m):
    m.release_buffers()


def _err(a, ref):
    return float(np.abs(a.detach().cpu().numpy().astype(np.float64) - np.asarray(ref, dtype=np.float64)).max())


def _record(msg):
    from test_gpu_parity import record
    record(msg)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16e5", "f16n4"])
def test_learned_single_steps_and_forward_vs_golden(precision):
    g = golden("learned_T128.npz")
    x_T, wav, noise = make_inputs(2, 200, seed=7, n_noise=1, T=128, wav_len=65536)
    worst = 0.0
    for name in ("cfdg_ddpm_x0", "generation_ddpm_x0"):
        m, _ = _model(precision, name)
        for t_index in (199, 1, 0):
            x_prev, _ = m.reverse_diffusion(x_T.cuda(), wav.cuda(), t_index, noise=noise[0].cuda())
            ref = g[f"{name}_t{t_index}"]
            e = _err(x_prev, ref) / max(1.0, float(np.abs(ref).max()))
            worst = max(worst, e)
            assert e < TOL_STEP[precision], (name, t_index, e)
        if name == "generation_ddpm_x0":
            # a plain forward with sampling=True, per-roll steps: every roll conditioned on the table; the spectrogram comes back
            # 2-D and trimmed like the reference's (model/diffwave.py:658,662)
            pred, spec = m(x_T.cuda(), wav.cuda(), torch.tensor([37, 150]).cuda(), sampling=True)
            assert spec.shape == (229, 128) and torch.equal(spec, m.trainable_parameters.detach()[:, :128])
            e = _err(pred, g["pred_u"])
            worst = max(worst, e)
            assert e < TOL_STEP[precision], e
            assert m._engines and all(e_.effective_precision in (precision, "f16e5") for e_, _ in m._engines.values())
        _close(m)
    _record(f"trainable_spec[{precision}] sampler steps + sampling=True forward vs reference goldens: worst max|delta| = {worst:.3e}")


def test_learned_validation_step_chain_and_parameter_update():
    """The two-dataset validation step (its second forward is sampling=True, task/diffusion.py:707-719), a whole guided 200-step
    chain through sample_loop, a conditional forward on the same engine, and a change of the parameter (the engine follows it)."""
    from oracle.diffroll_oracle import OracleDiffRoll
    g = golden("learned_T128.npz")
    m, hp = _model("f16n4", "cfdg_ddpm_x0")
    frame, audio, t, nz = make_labelled_batch(B=2)
    frame2, audio2, _, _ = make_labelled_batch(B=2, seed=78)
    losses, tensors = m.step([{"frame": frame.cuda(), "audio": audio.cuda()}, {"frame": frame2.cuda(), "audio": audio2.cuda()}],
                             t=t.cuda(), noise=nz.cuda())
    for k, ref in (("diffusion_loss", g["two_loss"]), ("unconditional_diffusion_loss", g["two_uncond_loss"])):
        assert abs(float(losses[k]) - float(ref)) < 1e-3 * max(1.0, float(ref)), k
    assert _err(tensors["pred_roll2"], g["two_pred_roll2"]) < TOL_STEP["f16n4"]
    xc, wc, nc = make_inputs(2, 200, seed=13, T=128, wav_len=65536)
    x0, spec, _ = m.sample_loop(xc.cuda(), wc.cuda(), noise=nc.cuda())
    e = _err(x0, g["chain_final"])
    _record(f"trainable_spec[f16n4] 200-step cfdg chain (2 rolls x 128 frames) vs reference golden: max|delta| = {e:.3e}")
    assert e < TOL_FINAL and spec.shape == (2, 229, 128)
    # conditional forward (BRANCH_COND on the learned-capable plan), then the parameter moves: results follow the oracle
    sd = make_state_dict(hp)
    orc = OracleDiffRoll(hp, sd)                     # on the CPU: exact fp32 (cuDNN would run the convolutions in TF32)
    steps = torch.tensor([57, 3])
    with torch.no_grad():
        ref_c, _ = orc(xc, wc, steps)
    pred_c, _ = m(xc.cuda(), wc.cuda(), steps.cuda())
    assert _err(pred_c, ref_c.numpy()) < TOL_STEP["f16n4"]
    with torch.no_grad():
        m.trainable_parameters.mul_(0.25)
    orc.sd["trainable_parameters"] = orc.sd["trainable_parameters"] * 0.25
    with torch.no_grad():
        ref_u, _ = orc(xc, wc, steps, sampling=True)
        ref_s, _ = orc.reverse_diffusion(xc, wc, 120, noise=nc[0])
    pred_u, _ = m(xc.cuda(), wc.cuda(), steps.cuda(), sampling=True)
    x_prev, _ = m.reverse_diffusion(xc.cuda(), wc.cuda(), 120, noise=nc[0].cuda())
    eu, es = _err(pred_u, ref_u.numpy()), _err(x_prev, ref_s.numpy())
    _record(f"trainable_spec[f16n4] after an in-place change of the table, vs the CPU oracle: forward {eu:.3e}, cfdg step {es:.3e}")
    assert eu < TOL_STEP["f16n4"] and es < TOL_STEP["f16n4"]
    assert _err(pred_u, g["pred_u"]) > 1e-3          # and it did move
    _close(m)


@pytest.mark.parametrize("case", ["one", "two"])
def test_learned_training_step_gradients_vs_reference_golden(case):
    """Row f3 under condition='trainable_spec': trainable_dropout (model/diffwave.py:695-699) conditions the dropped roll on the
    table, so the backward pass also owes d loss / d trainable_parameters (drb_train_set_spec_grad: g_y . W_c summed over the
    layers, then over the rolls that read the table).  Goldens: the live reference's training_step + backward() on the full
    640-frame clip (oracle/make_golden_train.py --learned); 'two' adds the unconditional second dataset, every roll on the table.
    Tolerance as in tests/test_gpu_train.py: every gradient within 1e-3 of its own max |value|."""
    import diffroll_b200 as M
    gold = golden("trainstep_learned_b2_T640.npz")
    frame, audio, t, noise = make_labelled_batch(B=2, T=640, wav_len=327680)
    hp = default_hparams(condition="trainable_spec")
    batch = {"frame": frame.cuda(), "audio": audio.cuda()}
    if case == "two":
        frame2, audio2, _, _ = make_labelled_batch(B=2, T=640, wav_len=327680, seed=78)
        hp["loss_keys"] = ["diffusion_loss", "unconditional_diffusion_loss"]
        batch = [batch, {"frame": frame2.cuda(), "audio": audio2.cuda()}]
    m = M.ClassifierFreeDiffRoll(**hp)
    m.load_state_dict(make_state_dict(hp), strict=True)
    m = m.cuda().train()
    total = m.training_step(batch, 0, t=t.cuda(), noise=noise.cuda(), dropout_mask=torch.from_numpy(gold["mask"]))
    torch.cuda.synchronize()
    assert abs(float(total) - float(gold[f"{case}/total_loss"])) < 2e-5

    def sample_of(g):
        flat = g.detach().reshape(-1)
        return flat[::max(1, -(-flat.numel() // 1024))]

    worst, worst_name = 0.0, ""
    for name, p in m.named_parameters():
        ref = gold[f"{case}/grad/{name}"]
        err = float(np.abs(sample_of(p.grad).cpu().numpy() - ref).max()) / max(float(np.abs(ref).max()), 1e-12)
        if err > worst:
            worst, worst_name = err, name
    ref = gold[f"{case}/grad/trainable_parameters"]
    e_tab = float(np.abs(sample_of(m.trainable_parameters.grad).cpu().numpy() - ref).max()) / float(np.abs(ref).max())
    n_tab = abs(float(m.trainable_parameters.grad.double().norm()) - float(gold[f"{case}/norm/trainable_parameters"])) / float(gold[f"{case}/norm/trainable_parameters"])
    _record(f"train[trainable_spec, {case}] B=2 T=640 vs live-reference golden: loss {float(total):.6f}, worst gradient rel. max|delta| = "
            f"{worst:.3e} ({worst_name}); table gradient {e_tab:.3e}, its norm {n_tab:.3e}")
    assert worst < 1e-3 and e_tab < 1e-3 and n_tab < 1e-3, (worst, worst_name, e_tab, n_tab)
    assert float(m.trainable_parameters.grad[:, 640].abs().max()) == 0.0      # the frame trim_spec_roll cuts off
    before = m.trainable_parameters.detach().clone()
    m.configure_optimizers()[0].step()
    assert float((m.trainable_parameters.detach() - before).abs().max()) > 0.0
    _close(m)


def test_learned_plan_contract_through_the_c_abi():
    """DRB_BRANCH_COND_LEARNED must be requested at plan creation; a step before drb_plan_set_uncond_spec is a state error."""
    from diffroll_b200 import _lib
    from diffroll_b200.engine import Engine
    from diffroll_b200.model import DiffusionEmbedding
    hp = default_hparams(condition="trainable_spec")
    sd = {k: v.cuda() for k, v in make_state_dict(hp).items()}
    emb = DiffusionEmbedding(200).embedding
    plain = Engine(sd, hp, 2, 128, 65536, emb, precision="f16e5")
    with pytest.raises(_lib.DrbError, match="DRB_BRANCH_COND_LEARNED"):
        plain.set_branches(_lib.BRANCH_COND_LEARNED)
    with pytest.raises(_lib.DrbError, match="DRB_BRANCH_COND_LEARNED"):
        plain.set_uncond_spec(sd["trainable_parameters"])
    eng = Engine(sd, hp, 2, 128, 65536, emb, precision="f16e5", branches=_lib.BRANCH_COND_LEARNED)
    assert eng.workspace_bytes > plain.workspace_bytes
    x_T, wav, _ = make_inputs(2, 200, seed=7, n_noise=0, T=128, wav_len=65536)
    eng.mel(wav.cuda())
    from diffroll_b200.task import _upd
    with pytest.raises(_lib.DrbError, match="drb_plan_set_uncond_spec"):
        eng.step(x_T.cuda(), None, 5, _upd(_lib.UPD_NONE))
    with pytest.raises(ValueError):
        eng.set_uncond_spec(sd["trainable_parameters"][:, :100])
    eng.set_uncond_spec(sd["trainable_parameters"])
    out = eng.step(x_T.cuda(), None, 5, _upd(_lib.UPD_NONE, w=-1.0))
    assert torch.isfinite(out).all()
    # DRB_BRANCH_LEARNED (every roll on the table, one forward) == the pair at guidance weight -1, bit for bit or to the last ulp
    eng.set_branches(_lib.BRANCH_LEARNED)
    single = eng.step(x_T.cuda(), None, 5, _upd(_lib.UPD_NONE))
    assert float((single - out).abs().max()) < 1e-6
    plain.close(); eng.close()


@pytest.mark.parametrize("precision", ["fp32", "f16n4"])
def test_learned_generation_with_an_odd_tile_count(precision):
    """One roll of 128 frames is a single tile: no CTA pair, so the tensor-core formats run the sampling=True forward as the
    (clip, table) pair at guidance weight -1 instead of DRB_BRANCH_LEARNED; fp32 takes the single branch.  Same numbers either way."""
    from oracle.diffroll_oracle import OracleDiffRoll
    m, hp = _model(precision, "generation_ddpm_x0")
    x_T, wav, noise = make_inputs(1, 200, seed=21, n_noise=1, T=128, wav_len=65536)
    orc = OracleDiffRoll(hp, make_state_dict(hp))
    with torch.no_grad():
        ref, _ = orc.reverse_diffusion(x_T, wav, 150, noise=noise[0])
    x_prev, spec = m.reverse_diffusion(x_T.cuda(), wav.cuda(), 150, noise=noise[0].cuda())
    assert m._learned_pair == (precision != "fp32") and spec.shape == (229, 128)
    e = _err(x_prev, ref.numpy())
    _record(f"trainable_spec[{precision}] generation step, 1 roll x 128 frames (odd tile count): max|delta| vs CPU oracle = {e:.3e}")
    assert e < TOL_STEP[precision]
    _close(m)

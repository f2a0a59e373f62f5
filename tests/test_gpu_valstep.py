"""GPU parity of the forward-only (validation) step, SURVEY.md section 8 row f3 (forward part): q_sample at per-roll
steps, network forward, extract_x0, loss (task/diffusion.py:651-763, run by validation_step :271-276) against the
golden vectors of the unmodified reference (oracle/make_golden.py gen_valstep), and the four small kernels against
their torch expressions bit for bit / to reduction order.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden
from diffroll_b200.synthetic import default_hparams, make_labelled_batch, make_state_dict

pytestmark = pytest.mark.gpu
TOL_STEP = {"fp32": 2e-4, "f16e5": 5e-4}     # one forward, as in tests/test_gpu_parity.py


def _model(precision, mode="x_0", loss_type="l2"):
    import diffroll_b200 as M
    hp = default_hparams()
    hp["training"] = dict(mode=mode)
    hp["loss_type"] = loss_type
    m = M.ClassifierFreeDiffRoll(**hp, precision=precision)
    m.load_state_dict(make_state_dict(hp))
    return m.cuda().eval()


def _close(m):
    for e, _ in m._engines.values():
        e.close()


@pytest.mark.parametrize("precision", ["fp32", "f16e5"])
@pytest.mark.parametrize("mode,loss_type", [("x_0", "l2"), ("x_0", "l1"), ("epsilon", "huber"), ("ex_0", "l2")])
def test_validation_step_vs_golden(mode, loss_type, precision):
    g = golden("valstep_b4_T128.npz")
    m = _model(precision, mode, loss_type)
    frame, audio, t, noise = make_labelled_batch()
    losses, tensors = m.step({"frame": frame.cuda(), "audio": audio.cuda()}, t=t.cuda(), noise=noise.cuda())
    tag = f"{mode}_{loss_type}"
    ref = g[f"{tag}_pred_roll"]
    scale = max(1.0, float(np.abs(ref).max()))
    err = float(np.abs(tensors["pred_roll"].cpu().numpy().astype(np.float64) - ref).max())
    assert err < TOL_STEP[precision] * scale, (tag, err)
    assert np.array_equal(tensors["label_roll"].cpu().numpy(), g["label_roll"])
    ref_loss = float(g[f"{tag}_loss"])
    assert abs(float(losses["diffusion_loss"]) - ref_loss) < 1e-3 * max(1.0, ref_loss), (tag, float(losses["diffusion_loss"]), ref_loss)
    assert float(m.validation_step({"frame": frame.cuda(), "audio": audio.cuda()})) > 0.0   # own draws of t / noise
    _close(m)


def test_validation_step_two_datasets_vs_golden():
    g = golden("valstep_b4_T128.npz")
    m = _model("f16e5")
    frame, audio, t, noise = make_labelled_batch()
    frame2, audio2, _, _ = make_labelled_batch(seed=78)
    losses, tensors = m.step([{"frame": frame.cuda(), "audio": audio.cuda()}, {"frame": frame2.cuda(), "audio": audio2.cuda()}],
                             t=t.cuda(), noise=noise.cuda())
    for k, ref in (("diffusion_loss", g["two_loss"]), ("unconditional_diffusion_loss", g["two_uncond_loss"])):
        assert abs(float(losses[k]) - float(ref)) < 1e-3 * max(1.0, float(ref)), k
    ref = g["two_pred_roll2"]
    err = float(np.abs(tensors["pred_roll2"].cpu().numpy().astype(np.float64) - ref).max())
    assert err < TOL_STEP["f16e5"] * max(1.0, float(np.abs(ref).max()))
    _close(m)


def test_diffusion_ops_vs_torch_expressions():
    """q_sample / extract_x0 / p_losses / Normalization kernels against the reference's torch expressions on the GPU."""
    from diffroll_b200 import diffusion_ops as ops
    m = _model("f16e5")
    gen = torch.Generator(device="cuda").manual_seed(5)
    B = 32
    x0 = torch.randn(B, 1, 640, 88, device="cuda", generator=gen)
    nz = torch.randn(B, 1, 640, 88, device="cuda", generator=gen)
    t = torch.randint(0, 200, (B,), device="cuda", generator=gen)
    sa, s1 = m.sqrt_alphas_cumprod.cuda(), m.sqrt_one_minus_alphas_cumprod.cuda()
    a, b = sa[t][:, None, None, None], s1[t][:, None, None, None]
    x_t = ops.q_sample(x0, t, m.sqrt_alphas_cumprod, m.sqrt_one_minus_alphas_cumprod, nz)
    assert float((x_t - (a * x0 + b * nz)).abs().max()) < 1e-6       # fma contraction vs two roundings
    back = ops.extract_x0(x_t, nz, t, m.sqrt_alphas_cumprod, m.sqrt_one_minus_alphas_cumprod)
    assert float((back - (x_t - b * nz) / a).abs().max()) < 1e-5 * float(1.0 / sa.min())
    for lt, fn in (("l1", F.l1_loss), ("l2", F.mse_loss), ("huber", F.smooth_l1_loss)):
        got, ref = float(ops.p_losses(x0, x_t, lt)), float(fn(x0.double(), x_t.double()))
        assert abs(got - ref) < 1e-6 * max(1.0, abs(ref)), (lt, got, ref)
    got = float(ops.p_losses(x0.flatten()[:1003 * 4 + 3], x_t.flatten()[:1003 * 4 + 3], "l1"))   # ragged tail
    assert abs(got - float(F.l1_loss(x0.flatten()[:4015].double(), x_t.flatten()[:4015].double()))) < 1e-6
    with pytest.raises(NotImplementedError):
        ops.p_losses(x0, x_t, "l3")
    roll = (torch.rand(8, 640, 88, device="cuda", generator=gen) < 0.05).float() * 3.0 + 1.0
    roll[3] = 2.0                                                         # constant roll: NaN -> min
    n = ops.normalize_imagewise(roll, 0.0, 1.0)
    mx = roll.flatten(1).max(1)[0][:, None, None]; mn = roll.flatten(1).min(1)[0][:, None, None]
    ref = (roll - mn) / (mx - mn)
    ref[torch.isnan(ref)] = 0.0
    assert torch.equal(n, ref) and float(n[3].abs().max()) == 0.0
    with pytest.raises(Exception):
        ops.q_sample(x0.cpu(), t.cpu(), m.sqrt_alphas_cumprod, m.sqrt_one_minus_alphas_cumprod, nz.cpu())
    _close(m)

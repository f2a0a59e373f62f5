#!/usr/bin/env python
"""Sampling entry point with the reference's command line (sampling.py there: Hydra `main(cfg)`):

    python sampling.py task=transcription dataloader.batch_size=32 dataset.num_samples=32 model.args.kernel_size=9
    python sampling.py task=generation task.timesteps=1000 checkpoint_path=weights/xyz.ckpt
    torchrun --nproc-per-node 8 sampling.py task=transcription dataset.num_samples=256 dataloader.batch_size=32

It composes config/sampling.yaml, builds `getattr(Model, cfg.model.name)` exactly like the reference
(sampling.py:54 there) — from a Lightning checkpoint when `checkpoint_path` is given, otherwise with seeded synthetic
weights — and drives `predict_step` over the batches.  Each rank of a torchrun launch takes a contiguous shard of
every batch; finished rolls are all-gathered and rank 0 writes them to `output_path`.  Figures, GIFs and MIDI export of
the reference's predict_step tail are out of scope.
"""
from __future__ import annotations

import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import diffroll_b200 as Model  # noqa: E402
from diffroll_b200.config import compose, default_config_path  # noqa: E402
from diffroll_b200.dist import all_gather_rolls, init_from_env, shard_bounds  # noqa: E402
from diffroll_b200.synthetic import make_state_dict  # noqa: E402


def build_model(cfg):
    task = dict(cfg.task)
    task.pop("name", None)
    cls = getattr(Model, cfg.model.name)
    if cfg.checkpoint_path:
        model = cls.load_from_checkpoint(cfg.checkpoint_path, sampling=cfg.task.sampling,
                                         frame_threshold=cfg.task.frame_threshold,
                                         generation_filter=cfg.task.generation_filter,
                                         inpainting_t=cfg.task.inpainting_t, inpainting_f=cfg.task.inpainting_f,
                                         precision=cfg.precision)
    else:
        kw = dict(cfg.model.args)
        kw.update(task)
        kw["spec_args"] = dict(cfg.spec.args)
        model = cls(**kw, precision=cfg.precision)
        hp = dict(model.hparams)
        hp["spec_args"] = dict(hp["spec_args"])
        model.load_state_dict(make_state_dict(hp, seed=cfg.seed))
    return model.cuda().eval()


def build_waveforms(cfg, n):
    if cfg.dataset.name == "Sampling":
        return torch.randn(n, cfg.sequence_length)                      # reference sampling.py:45
    if cfg.dataset.name == "Tensor":
        wav = torch.load(cfg.dataset.args.path).float()
        if wav.ndim != 2 or wav.shape[1] != cfg.sequence_length:
            raise ValueError(f"expected waveforms [N, {cfg.sequence_length}], got {tuple(wav.shape)}")
        return wav[:n]
    raise NotImplementedError(f"dataset '{cfg.dataset.name}': only in-memory tensors are built here (file decoding is I/O, out of scope)")


def main(argv=None):
    cfg = compose(default_config_path(), list(sys.argv[1:] if argv is None else argv))
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    torch.manual_seed(cfg.seed)
    S = cfg.dataset.num_samples
    x = torch.randn(S, 1, 640, 88)                                      # reference sampling.py:27
    waveform = build_waveforms(cfg, S)
    S = min(S, waveform.shape[0])
    model = build_model(cfg)
    bs = cfg.dataloader.batch_size
    rolls = []
    t0 = time.time()
    for lo in range(0, S, bs):
        hi = min(S, lo + bs)
        a, b = shard_bounds(hi - lo, rank, world)
        if b > a:
            batch = (x[lo + a:lo + b].cuda(non_blocking=True), waveform[lo + a:lo + b].cuda(non_blocking=True))
            # Step noise (task/diffusion.py:1023) is drawn for the WHOLE batch from a generator seeded per batch and
            # sliced to this rank's rolls: shards are statistically independent and an N-rank run returns exactly the
            # rolls of a 1-rank run (every rank seeding its own generator alike would repeat one noise sequence per shard).
            gen = torch.Generator(device="cuda").manual_seed(int(cfg.seed) + 1 + lo // bs)
            roll_pred, _, _ = model.predict_step(batch, lo // bs, generator=gen, shard=(hi - lo, a, b))
            part = torch.from_numpy(roll_pred).cuda()
        else:
            part = torch.empty(0, 1, 640, 88, device="cuda")
        rolls.append(all_gather_rolls(part, hi - lo).cpu())
    if rank == 0:
        out = torch.cat(rolls, 0)
        os.makedirs(os.path.dirname(cfg.output_path) or ".", exist_ok=True)
        torch.save({"rolls": out, "frame_threshold": cfg.task.frame_threshold, "sampler": cfg.task.sampling.type}, cfg.output_path)
        print(f"{out.shape[0]} rolls x {cfg.task.timesteps} steps ({cfg.task.sampling.type}) in {time.time() - t0:.1f} s -> {cfg.output_path}")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

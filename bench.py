#!/usr/bin/env python
"""Benchmark of the DiffRoll sampling hot path (BASELINE.json: diffusion sampling steps/sec, B=32,
640x88 roll, 200 steps) on N B200s of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config {1,2,3,4}] [--impl reference] [--lean]

A "step" is one reverse-diffusion timestep (task/diffusion.py:999-1025: conditional + unconditional
network forward, guidance combine, posterior update) applied to one batch of synthetic rolls.

--config selects the BASELINE.json configuration the headline `value` is measured on (default 1, the one the
metric is quoted on):
  1  configs[1]  32 rolls PER GPU, transcription (inpainting_ddpm_x0 w=0.5, no masks), 200 timesteps   weak scaling
                 (N=8 is configs[4]: 256 rolls over 8 GPUs)
  2  configs[2]  64 rolls per GPU, unconditional generation (generation_ddpm_x0, spec == -1), 1000 timesteps   weak
  3  configs[3]  32 rolls IN TOTAL split over the N ranks, inpainting with frames [0,320) masked             strong
  4  configs[4]  256 rolls in total split over the N ranks, transcription                                    strong
With the default --config 1 the line also carries, as extra keys, short measurements of the other three
(`configs2`, `strong`) and of the incumbent GPU path (`gpu_eager_baseline`: the reference's torch ops run eagerly
through cuDNN / cuBLAS on the same GPU, TF32 off and on); --lean skips those extras (ncu runs).

Own arm, one JSON line on rank 0:
  value     steps/s, whole job, inputs resident in HBM (x_t, pre-drawn noise, spectrogram, weights)
  e2e       the same metric through the public API (ClassifierFreeDiffRoll.sample_loop = predict_step's
            loop) with pinned HOST inputs: H2D of x_T and the waveform, mel front-end, per-step noise
            draw, K steps, and a D2H copy of every step's roll (the reference's per-step .cpu(), :530)
  roofline  tcgen05 gate kernel (dilated conv + gate): algorithmic FLOPs / CUDA-event time, vs measured bf16 peak
  roofline_hbm  the HBM-bound kernels (posterior epilogue, in_proj, mel front-end): algorithmic bytes / CUDA-event time
  cpu_baseline  oracle port (oracle/diffroll_oracle.py) on the host cores, bounded sample at the same batch

Reference arm (--impl reference): the reference is pure Python and cannot travel to the GPU box, so
its CPU path is timed through the oracle port (op-for-op restatement, same torch CPU kernels) on all
host threads, K steps of the SAME workload (batch 32, no rescaling).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "diffusion sampling steps/sec (B=32, 640x88 roll, 200 steps)"
UNIT = "steps/s"
BATCH = 32
E2E_REPEATS = 3      # end-to-end calls timed per configuration; the median is reported, all are listed
FRAMES = 640
WAVE_LEN = 327680
C, L, KSIZE, PITCHES, NMELS = 512, 15, 9, 88, 229
# algorithmic FLOPs (2 x MAC), SURVEY.md section 8(d)
FLOP_DILATED = 2.0 * FRAMES * (2 * C) * (KSIZE * C)      # 6.0398e9 per roll-branch-layer
FLOP_BRANCH = 100.79e9                                   # one network forward of one roll, algorithmic minimum

CONFIGS = {
    1: dict(tag="configs[1]", batch=BATCH, scaling="weak", hp={}, branches=2,
            what="transcription: inpainting_ddpm_x0 w=0.5 (2 network forwards/step), timesteps=200"),
    2: dict(tag="configs[2]", batch=64, scaling="weak", hp=dict(timesteps=1000, sampling_type="generation_ddpm_x0"), branches=1,
            what="unconditional generation: generation_ddpm_x0, spec == -1 (1 network forward/step), timesteps=1000"),
    3: dict(tag="configs[3]", global_batch=32, scaling="strong", hp=dict(inpainting_t=[0, 320]), branches=2,
            what="inpainting: inpainting_ddpm_x0 w=0.5, frames [0,320) of the spectrogram masked, timesteps=200"),
    4: dict(tag="configs[4]", global_batch=256, scaling="strong", hp={}, branches=2,
            what="transcription: inpainting_ddpm_x0 w=0.5, timesteps=200"),
}


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


_REAL_STDOUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write there too (NCCL prints its version banner on stdout
    at NCCL_DEBUG=VERSION/WARN/INFO even with NCCL_DEBUG_FILE set, as measured on the GPU box).  Keep a private
    duplicate of the real stdout for the result line and point fd 1 at stderr for everything else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# CPU side (oracle port) — used by cpu_baseline and by --impl reference
# --------------------------------------------------------------------------------------------------
def cpu_steps_per_s(n_steps, warmup, max_seconds=300.0):
    """Times the reference's sampler step (inpainting_ddpm_x0: 2 forwards incl. the mel front-end it recomputes,
    posterior update, per-step .cpu().numpy()) via the oracle port on all host threads, at the benchmark's own batch
    (32 rolls, no rescaling).  If the first step shows that warmup + n_steps would take longer than max_seconds, fewer
    steps are timed (the per-step cost is constant).  Returns (steps/s, description, threads)."""
    import torch
    from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict
    from oracle.diffroll_oracle import OracleDiffRoll
    # torchrun exports OMP_NUM_THREADS=1 to every worker; the CPU arm must still use all the host cores it can
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    if torch.get_num_threads() < avail:
        torch.set_num_threads(avail)
    threads = torch.get_num_threads()
    hp = default_hparams()
    T = hp["timesteps"]
    orc = OracleDiffRoll(hp, make_state_dict(hp))
    x, w, nz = make_inputs(BATCH, T, seed=123, n_noise=n_steps + warmup + 1)
    with torch.no_grad():
        t_index, i = T - 1, 0
        t0 = time.perf_counter()
        x, _ = orc.reverse_diffusion(x, w, t_index, noise=nz[i]); i += 1; t_index -= 1      # first step: also the probe
        probe = time.perf_counter() - t0
        n_warm = max(warmup - 1, 0)
        n_timed = n_steps
        if (n_warm + n_timed) * probe > max_seconds:
            n_warm = min(n_warm, 1)
            n_timed = max(1, int(max_seconds / probe) - n_warm)
        for _ in range(n_warm):
            x, _ = orc.reverse_diffusion(x, w, t_index, noise=nz[i]); i += 1; t_index -= 1
        t0 = time.perf_counter()
        for _ in range(n_timed):
            x, _ = orc.reverse_diffusion(x, w, t_index, noise=nz[i]); i += 1; t_index -= 1
            _ = x.detach().cpu().numpy()
        dt = time.perf_counter() - t0
    val = n_timed / dt
    desc = (f"{n_timed} timed steps (after {n_warm + 1} warm-up) of inpainting_ddpm_x0 at batch {BATCH} on {threads} threads, "
            f"{dt:.1f} s; no rescaling (same batch as the GPU arm)")
    return val, desc, threads


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    val, desc, threads = cpu_steps_per_s(steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1000.0 / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: batch=32 synthetic 640x88 rolls + 229-bin mel, inpainting_ddpm_x0 w=0.5, "
                               "timesteps=200, ClassifierFreeDiffRoll k=9 (CPU, host cores)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, rank, world, local):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local = rank, world, local
        self.dev = torch.device("cuda", local)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure(ctx, cfg_id, K, W, precision, want_e2e=True, want_profile=True, sampler=None):
    """Times one BASELINE configuration on this job's ranks.  Returns a dict of raw measurements (rank-max times)."""
    torch = ctx.torch
    import diffroll_b200 as M
    from diffroll_b200 import _lib
    from diffroll_b200.dist import all_gather_rolls, shard_bounds
    from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict
    cfg = CONFIGS[cfg_id]
    lib = _lib.load()
    hp = default_hparams(**cfg["hp"])
    TS = hp["timesteps"]
    if cfg["scaling"] == "weak":
        batch, global_batch = cfg["batch"], cfg["batch"] * ctx.world
    else:
        lo, hi = shard_bounds(cfg["global_batch"], ctx.rank, ctx.world)
        batch, global_batch = hi - lo, cfg["global_batch"]
    model = M.ClassifierFreeDiffRoll(**hp, precision=precision)
    model.load_state_dict(make_state_dict(hp))
    model = model.cuda().eval()
    x_T, wav, _ = make_inputs(batch, TS, seed=123 + ctx.rank, n_noise=0)
    x_host, w_host = x_T.pin_memory(), wav.pin_memory()
    ups, branches, masks = model._all_updates()

    # ---- resident run -----------------------------------------------------------------------------
    x_dev, w_dev = x_host.to(ctx.dev), w_host.to(ctx.dev)
    eng, xx, spec = model._prepare(x_dev, w_dev, branches, *masks)
    n_noise = min(max(K, W, 3, 20), TS - 1)
    gen = torch.Generator(device=ctx.dev).manual_seed(1000 + ctx.rank)
    noise = torch.randn((n_noise,) + tuple(xx.shape), device=ctx.dev, generator=gen)

    def run_steps(n):
        """n timesteps as chains of <= n_noise steps starting at t = T-1 (x restarted per chain)."""
        done = 0
        while done < n:
            m = min(n_noise, n - done)
            x = xx.clone()
            eng.loop(x, noise, ups[:m], TS, TS - m)
            done += m

    run_steps(max(W, 3))
    torch.cuda.synchronize()
    ctx.barrier()   # NCCL's lazy communicator setup takes seconds: keep it out of the clock samples and the timed region
    sampler_started = False
    if sampler is not None and ctx.rank == 0:
        sampler.start(); sampler_started = True
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier(); torch.cuda.synchronize()
    lib.drb_launch_count(1)
    ev0.record()
    run_steps(K)
    ev1.record()
    torch.cuda.synchronize(); ctx.barrier()
    launches = int(lib.drb_launch_count(0))
    ms = ctx.max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if sampler_started else None
    out = dict(cfg=cfg, hp=hp, batch=batch, global_batch=global_batch, ms=ms, launches=launches, clocks=clocks,
               workspace_bytes=eng.workspace_bytes, range_max=eng.range_max(reset=True), precision=eng.effective_precision)

    # ---- per-kernel CUDA-event timing inside a running loop (roofline) ------------------------------
    if want_profile:
        n_prof = min(K, n_noise, 20)
        eng.profile(True)
        x = xx.clone()
        eng.loop(x, noise, ups[:n_prof], TS, TS - n_prof)
        prof, gate_layers = eng.profile_read_detail()
        eng.profile(False)
        out.update(prof=prof, gate_layers=gate_layers, n_prof=n_prof)
        if branches != _lib.BRANCH_UNCOND:   # mel front-end alone (once per clip): CUDA events around drb_mel_forward
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.mel(w_dev, *masks, want_spec=False)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                eng.mel(w_dev, *masks, want_spec=False)
            e1.record()
            torch.cuda.synchronize()
            out["mel_ms"] = e0.elapsed_time(e1) / 5
            model._mel_key = None

    # ---- end to end through the public API, host buffers ---------------------------------------------
    if want_e2e:
        Ke = min(K, TS)
        # warm-up with the same shape: allocates the pinned trajectory buffer the public API keeps per shape
        x0w, _, _ = model.sample_loop(x_host.to(ctx.dev, non_blocking=True), w_host.to(ctx.dev, non_blocking=True), keep_trajectory=True, n_steps=Ke)
        if ctx.world > 1:   # NCCL sets a collective of a new size up lazily (60 ms on the first call): part of the warm-up
            all_gather_rolls(x0w, global_batch)
        del x0w
        # one chain is one host call: E2E_REPEATS independent calls (new device copies of the host buffers each time, so the mel
        # front-end and the conditioner tables run again), each timed on the device as the max over ranks; the MEDIAN is reported
        # and all of them are listed (a single call is exposed to one-off host hiccups)
        e2e_all = []
        for _ in range(E2E_REPEATS):
            torch.cuda.synchronize(); ctx.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            xd = x_host.to(ctx.dev, non_blocking=True)
            wd = w_host.to(ctx.dev, non_blocking=True)
            x0, _, traj = model.sample_loop(xd, wd, keep_trajectory=True, n_steps=Ke)
            if ctx.world > 1:   # the path's only collective: one all-gather of the finished rolls over NCCL / NVLink
                rolls = all_gather_rolls(x0, global_batch)
                assert rolls.shape[0] == global_batch
            e1.record()
            torch.cuda.synchronize(); ctx.barrier()
            e2e_all.append(ctx.max_over_ranks(e0.elapsed_time(e1)))
            del xd, wd
        out.update(ms_e2e=sorted(e2e_all)[len(e2e_all) // 2], ms_e2e_all=e2e_all, Ke=Ke,
                   h2d=(x_host.numel() + (w_host.numel() if branches != _lib.BRANCH_UNCOND else 0)) * 4 / Ke,
                   d2h=x_host.numel() * 4)
    model.release_buffers()
    del model, eng, noise
    torch.cuda.empty_cache()
    return out


def steps_per_s(m, world, K, key="ms"):
    mult = world if m["cfg"]["scaling"] == "weak" else 1
    return mult * K / (m[key] / 1000.0)


def gate_roofline(m, peaks, precision, world, K, traffic):
    prof, n_prof = m["prof"], m["n_prof"]
    gate_ms, gate_n = prof["gate"]
    nb = m["batch"] * m["cfg"]["branches"]
    gate_flops = nb * FLOP_DILATED                    # per launch, algorithmic (conditioner GEMM and the operand split excluded)
    gate_avg_s = gate_ms / max(gate_n, 1) / 1000.0
    achieved = gate_flops / gate_avg_s / 1e12 if gate_n else None
    step_ms_prof = sum(prof[k][0] for k in ("gate", "out", "in_proj", "head")) / n_prof
    full = m["gate_layers"][1:]                       # layers 1..L-1: launches that issue every MMA (layer 0 may be shared)
    full_avg_s = sum(full) / max(len(full), 1) / n_prof / 1000.0
    r = {
        "kernel": {"f16n4": "umma_gate_n4_kernel (f16n4: fp16 + block-scaled fp4 correction, persistent CTA pairs, tap window)",
                   "f16e5": "umma_gate_pers_kernel<3> (f16e5, persistent CTA pairs, tap window)",
                   "f16f8": "umma_gate_win_kernel<2> (f16f8, CTA pairs, tap window)",
                   "bf16x3": "umma_gate_pers_kernel<1> (bf16x3, persistent CTA pairs, tap window)"}.get(precision, f"umma_gate_kernel<{precision}>"),
        "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
        "frac": (achieved / peaks["bf16_sustained"]) if achieved else None, "traffic": traffic,
        "traffic_source": "static: one ncu --set full capture of this kernel at this shape (profiles/roofline_traffic.json), "
                          "not re-measured by this run",
        "peak_source": f"{peaks['source']} bf16 dense, sustained (kernel timed inside a long step); burst {peaks['bf16']}",
        "mma_multiplicity": {"bf16x3": 3, "f16f8": 2, "f16e5": 2, "f16n4": 1.5}.get(precision, 1),
        "algorithmic_flops_per_launch": gate_flops,
        "avg_launch_ms": gate_avg_s * 1e3, "launches_timed": gate_n,
        "full_launch_avg_ms": full_avg_s * 1e3,
        "frac_full_launches": gate_flops / full_avg_s / 1e12 / peaks["bf16_sustained"] if full_avg_s > 0 else None,
        "share_of_step": (gate_ms / n_prof) / step_ms_prof if step_ms_prof else None,
        "per_step_ms": {k: v[0] / n_prof for k, v in prof.items()},
        "whole_step_algorithmic_tflops": FLOP_BRANCH * nb * (K / (m["ms"] / 1000.0)) / 1e12,
    }
    return r


def hbm_rooflines(m, peaks):
    """The HBM-bound kernels of the path: algorithmic bytes / CUDA-event time vs measured copy bandwidth."""
    prof, n_prof, B = m["prof"], m["n_prof"], m["batch"]
    nb = B * m["cfg"]["branches"]
    roll = B * FRAMES * PITCHES * 4
    out = {}

    def entry(bytes_, us, what):
        gbs = bytes_ / (us * 1e-6) / 1e9 if us else None
        return {"bytes": bytes_, "us": us, "achieved_gbs": gbs, "frac": gbs / peaks["hbm_gbs"] if gbs else None, "what": what}

    ho_ms, ho_n = prof["head_out"]
    if ho_n:
        out["posterior"] = entry(nb * FRAMES * C * 4 + 4 * roll, ho_ms / ho_n * 1e3,
                                 "output_projection + guidance + posterior update: h fp32 [NB,T,512] in, x_t, noise in, x_{t-1} (+ x0) out")
    ip_ms, ip_n = prof["in_proj"]
    if ip_n:
        out["in_proj"] = entry(roll + nb * FRAMES * C * (4 + 4), ip_ms / ip_n * 1e3,
                               "input_projection + ReLU: x_t in, fp32 residual stream + operand pair [NB,T,512] out")
    if "mel_ms" in m:
        out["mel"] = entry(B * WAVE_LEN * 4 + B * NMELS * FRAMES * 4, m["mel_ms"] * 1e3,
                           "STFT -> mel -> log -> min-max -> mask, once per clip: waveform in, spectrogram out")
    out["peak_gbs"] = peaks["hbm_gbs"]
    out["peak_source"] = f"{peaks['source']} copy bandwidth"
    return out


def gpu_eager_baseline(ctx):
    """SURVEY.md section 8(d) 'kernel to beat': the reference's own torch ops run eagerly on this GPU (cuDNN / cuBLAS),
    fp32 with TF32 off, then with TF32 on (torch 1.11's default for convs; fails the 1e-3 parity bar)."""
    torch = ctx.torch
    from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict
    from oracle.diffroll_oracle import OracleDiffRoll
    hp = default_hparams()
    x_T, wav, noise = make_inputs(BATCH, 200, seed=123, n_noise=1)
    x, w, nz = x_T.to(ctx.dev), wav.to(ctx.dev), noise[0].to(ctx.dev)
    orc = OracleDiffRoll(hp, make_state_dict(hp), device=ctx.dev)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    res = {}
    try:
        for name, flag in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = flag
            torch.backends.cuda.matmul.allow_tf32 = flag
            with torch.no_grad():
                orc.reverse_diffusion(x, w, 199, noise=nz)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    orc.reverse_diffusion(x, w, 199, noise=nz)
                e1.record()
                torch.cuda.synchronize()
            res[f"{name}_steps_per_s"] = 3000.0 / e0.elapsed_time(e1)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    res["what"] = "reference torch ops (oracle port) eager on this GPU, configs[1] step at batch 32, 3 timed steps per mode"
    del orc
    torch.cuda.empty_cache()
    return res


def train_step_extra(ctx, B=16, iters=3):
    """SURVEY.md section 8 row f3, measured beside the headline: one training step (training_step + fused Adam; fp32 CUDA-core
    forward + hand-written backward, csrc/train.cu) on B rolls x 640 frames, against torch autograd over the oracle +
    torch.optim.Adam run eagerly on this GPU (cuDNN / cuBLAS) in fp32 with TF32 off (same arithmetic grade) and on."""
    torch = ctx.torch
    import diffroll_b200 as M
    from diffroll_b200.synthetic import default_hparams, make_labelled_batch, make_state_dict
    from oracle.diffroll_oracle import OracleDiffRoll
    hp = default_hparams(); hp["lr"] = 1e-4
    frame, audio, _, noise = make_labelled_batch(B=B, T=FRAMES, wav_len=WAVE_LEN, seed=5)
    t = ((torch.arange(B) * 37) % hp["timesteps"]).to(ctx.dev)
    batch = {"frame": frame.to(ctx.dev), "audio": audio.to(ctx.dev)}
    nz = noise.to(ctx.dev)
    mask = (torch.arange(B) % 4 == 1).long()
    res = {"what": f"one training step (forward with spec dropout, backward to 130 tensors, Adam) on {B} rolls x {FRAMES} frames; "
                   "ours = csrc/train.cu (tcgen05 products at fp32-grade gradient parity; fp32 CUDA-core variant beside it); "
                   "eager = torch autograd over the oracle + torch.optim.Adam on this GPU",
           "batch": B, "algorithmic_tflop_per_step": 3 * FLOP_BRANCH * B / 1e12}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn):
        fn(); torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    # default: tensor-core products (f16x3 forward with per-tap fp32 accumulation, f16e5 backward); then everything in fp32 on the CUDA cores
    for key, env in (("ours_ms", None), ("ours_fp32_cuda_cores_ms", "0")):
        old_env = os.environ.get("DRB_TRAIN_TC")
        if env is not None:
            os.environ["DRB_TRAIN_TC"] = env
        try:
            m = M.ClassifierFreeDiffRoll(**hp); m.load_state_dict(make_state_dict(hp)); m = m.to(ctx.dev).train()
            opt = m.configure_optimizers()[0]

            def ours():
                opt.zero_grad(); m.training_step(batch, 0, t=t, noise=nz, dropout_mask=mask); opt.step()
            res[key] = timed(ours)
            if env is None:
                res["ours_steps_per_s"] = 1000.0 / res[key]
                res["ours_algorithmic_tflops"] = res["algorithmic_tflop_per_step"] / (res[key] / 1000.0)
                res["workspace_bytes"] = list(m._train_engines.values())[0].workspace_bytes
            m.release_buffers(); del m, opt
            torch.cuda.empty_cache()
        finally:
            if env is not None:
                if old_env is None:
                    os.environ.pop("DRB_TRAIN_TC", None)
                else:
                    os.environ["DRB_TRAIN_TC"] = old_env
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, flag in (("eager_fp32_ms", False), ("eager_tf32_ms", True)):
            torch.backends.cudnn.allow_tf32 = flag; torch.backends.cuda.matmul.allow_tf32 = flag
            orc = OracleDiffRoll(hp, make_state_dict(hp), device=ctx.dev)
            params = {k: torch.nn.Parameter(v.clone()) for k, v in orc.sd.items() if not k.startswith("mel_layer.")}
            ropt = torch.optim.Adam(params.values(), lr=hp["lr"])

            def ref():
                orc.sd.update({k: q.data for k, q in params.items()})
                _, grads, _ = orc.train_step(batch, t.cpu(), nz, dropout_mask=mask)
                for k, q in params.items():
                    q.grad = grads[k]
                ropt.step()
            res[name] = timed(ref)
            del orc, params, ropt
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return res


def run_b200(args, rank, world, local):
    import torch
    from diffroll_b200 import build as _build

    if not os.path.exists(_build.LIB):       # a checkout without the (git-ignored) built library: compile it once
        if local == 0:
            _build.build()
        if world > 1:
            import torch.distributed as _d
            _d.barrier()
    torch.cuda.set_device(local)
    ctx = Ctx(rank, world, local)
    K, W = args.steps, max(args.warmup, 0)
    peaks = read_peaks()
    cfg_id = args.config
    cfg = CONFIGS[cfg_id]
    if args.batch:
        cfg = dict(cfg); cfg["batch" if cfg["scaling"] == "weak" else "global_batch"] = args.batch
        CONFIGS[cfg_id] = cfg

    main = measure(ctx, cfg_id, K, W, args.precision, sampler=ClockSampler(local))
    value = steps_per_s(main, world, K)

    extras = {}
    if cfg_id == 1 and not args.lean:
        # the other BASELINE.json configurations, short runs (never the headline value)
        Kx = min(K, 20)
        try:
            m2 = measure(ctx, 2, Kx, 3, args.precision)
            r2 = gate_roofline(m2, peaks, m2["precision"], world, Kx, None)
            extras["configs2"] = {
                "workload": f"configs[2]: batch={m2['batch']} per GPU, {CONFIGS[2]['what']}",
                "value": steps_per_s(m2, world, Kx), "unit": UNIT, "steps": Kx, "ms_per_step": m2["ms"] / Kx,
                "e2e": {"value": steps_per_s(m2, world, m2["Ke"], "ms_e2e"), "amortised_over": m2["Ke"]},
                "gate_frac": r2["frac"], "gate_avg_launch_ms": r2["avg_launch_ms"], "per_step_ms": r2["per_step_ms"],
                "workspace_bytes": m2["workspace_bytes"], "scaling": "weak", "precision": m2["precision"]}
        except Exception as e:
            extras["configs2"] = {"error": repr(e)}
        strong = {}
        for cid, key in ((3, "configs3_global32_inpainting"), (4, "global256_transcription")):
            try:
                ms_ = measure(ctx, cid, Kx, 3, args.precision, want_profile=False)
                strong[key] = {"workload": f"{CONFIGS[cid]['tag']}: {ms_['global_batch']} rolls in total over {world} GPU(s) "
                                           f"({ms_['batch']} on rank 0), {CONFIGS[cid]['what']}",
                               "value": steps_per_s(ms_, world, Kx), "unit": UNIT, "steps": Kx, "ms_per_step": ms_["ms"] / Kx,
                               "sample_steps_per_s": steps_per_s(ms_, world, Kx) * ms_["global_batch"],
                               "e2e": {"value": steps_per_s(ms_, world, ms_["Ke"], "ms_e2e"), "amortised_over": ms_["Ke"]},
                               "workspace_bytes": ms_["workspace_bytes"], "scaling": "strong"}
            except Exception as e:
                strong[key] = {"error": repr(e)}
        extras["strong"] = strong
        if rank == 0:
            try:
                extras["gpu_eager_baseline"] = gpu_eager_baseline(ctx)
            except Exception as e:
                extras["gpu_eager_baseline"] = {"error": repr(e)}
            try:
                extras["train_step"] = train_step_extra(ctx)
            except Exception as e:
                extras["train_step"] = {"error": repr(e)}
        ctx.barrier()

    if rank != 0:
        return
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if cfg_id == 1 and os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
                traffic = tj.get("umma_gate_n4_kernel_dram_bytes_per_launch", tj.get("umma_gate_kernel_dram_bytes_per_launch"))
        except Exception:
            traffic = None
    prec = main["precision"]       # what the plan actually computes in (f16n4 steps down to f16e5 on odd tile counts)
    roofline = gate_roofline(main, peaks, prec, world, K, traffic)
    cpu = None
    if not args.no_cpu_baseline:
        try:
            v, desc, threads = cpu_steps_per_s(3, 1, max_seconds=30.0)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc}
        except Exception as e:  # the GPU numbers must still be printed
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    per_gpu = f"batch={main['batch']} per GPU" if cfg["scaling"] == "weak" else \
        f"{main['global_batch']} rolls in total split over {world} GPU(s) ({main['batch']} on rank 0)"
    e2e_value = steps_per_s(main, world, main["Ke"], "ms_e2e")
    line = {
        "metric": METRIC if cfg_id == 1 else f"diffusion sampling steps/sec ({cfg['tag']})",
        "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": main["ms"] / K, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
        "dtype": {"f16n4": "f16n4 (fp16 tcgen05 product + block-scaled e2m1 correction product, one fp32 accumulator; "
                           "RES / HEAD GEMMs: f16e5)",
                  "bf16x3": "bf16x3 (bf16 hi/lo split, 3 tcgen05 products, fp32 accumulate)",
                  "f16f8": "f16f8 (fp16 tcgen05 product + e4m3 correction product, fp32 accumulate)",
                  "f16e5": "f16e5 (fp16 tcgen05 product + e5m2 correction product, one fp32 accumulator)"}.get(prec, prec),
        "data": "synthetic",
        "config": {"workload": f"{cfg['tag']}: {per_gpu}, synthetic 640x88 rolls + 229-bin mel, {cfg['what']}, "
                               "ClassifierFreeDiffRoll k=9, random weights",
                   "l2": "per-step working set (weights 0.6 GB + activations 0.6 GB) exceeds the 126 MB L2; no flush needed",
                   "parallelism": f"dp{world} (independent batch shards, no data-path collective; one all-gather of the "
                                  "finished rolls inside the e2e region when N>1)"},
        "sample_steps_per_s": value * (main["batch"] if cfg["scaling"] == "weak" else main["global_batch"]),
        "chain_wall_s": main["hp"]["timesteps"] / (K / (main["ms"] / 1000.0)),   # one whole chain of a shard
        "clocks": main["clocks"], "gpu_launches": main["launches"],
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": main["h2d"], "d2h_bytes_per_step": main["d2h"],
                "steps": main["Ke"], "ms_per_step": main["ms_e2e"] / main["Ke"], "amortised_over": main["Ke"],
                "repeats": len(main["ms_e2e_all"]), "ms_per_call_all": [round(v, 3) for v in main["ms_e2e_all"]], "statistic": "median",
                "note": "per-clip work (H2D of the clip, mel front-end, conditioner tables) is spread over `amortised_over` "
                        "steps; a full chain spreads it over all of its timesteps"},
        "workspace_bytes": main["workspace_bytes"], "activation_operand_max_abs": main["range_max"],
        "roofline": roofline, "roofline_hbm": hbm_rooflines(main, peaks), "cpu_baseline": cpu,
    }
    line.update(extras)
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4])
    ap.add_argument("--batch", type=int, default=0, help="override the configuration's (per-GPU or global) batch")
    ap.add_argument("--precision", default="f16n4", choices=["f16n4", "f16e5", "f16f8", "bf16x3", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lean", action="store_true", help="headline configuration only (no configs2 / strong / eager extras)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        # NCCL prints its version banner on stdout at the VERSION/WARN/INFO levels: send its log to stderr instead so
        # that stdout carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    try:
        run_b200(args, rank, world, local)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

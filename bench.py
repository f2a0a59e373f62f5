#!/usr/bin/env python
"""Benchmark of the DiffRoll sampling hot path (BASELINE.json: diffusion sampling steps/sec, B=32,
640x88 roll, 200 steps) on N B200s of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one reverse-diffusion timestep (task/diffusion.py:999-1025: conditional + unconditional
network forward, guidance combine, posterior update) applied to one batch of 32 synthetic rolls.
Weak scaling: every rank runs its own batch of 32 (N=8 is BASELINE.json configs[4], B=256).

Own arm, one JSON line on rank 0:
  value     steps/s, whole job, inputs resident in HBM (x_t, pre-drawn noise, spectrogram, weights)
  e2e       the same metric through the public API (ClassifierFreeDiffRoll.sample_loop = predict_step's
            loop) with pinned HOST inputs: H2D of x_T and the waveform, mel front-end, per-step noise
            draw, K steps, and a D2H copy of every step's roll (the reference's per-step .cpu(), :530)
  roofline  tcgen05 gate kernel (dilated conv + gate): algorithmic FLOPs / CUDA-event time, vs measured bf16 peak
  cpu_baseline  oracle port (oracle/diffroll_oracle.py) on the host cores, bounded sample

Reference arm (--impl reference): the reference is pure Python and cannot travel to the GPU box, so
its CPU path is timed through the oracle port (op-for-op restatement, same torch CPU kernels) on all
host threads, on a bounded sample of the same workload.
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
FRAMES = 640
WAVE_LEN = 327680
TIMESTEPS = 200
C, L, KSIZE = 512, 15, 9
# algorithmic FLOPs (2 x MAC), SURVEY.md section 8(d)
FLOP_DILATED = 2.0 * FRAMES * (2 * C) * (KSIZE * C)      # 6.0398e9 per roll-branch-layer
FLOP_OUTPROJ = 2.0 * FRAMES * (2 * C) * C                # 0.6711e9
FLOP_STEP_PER_ROLL = 201.6e9                             # both branches, algorithmic minimum


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
def cpu_steps_per_s(n_steps, warmup, budget_s, fixed_batch=None):
    """Times the reference's sampler step (inpainting_ddpm_x0: 2 forwards incl. the mel front-end it recomputes,
    posterior update, per-step .cpu().numpy()) via the oracle port on all host threads.  Returns
    (steps/s scaled to B=32, description, threads)."""
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
    orc = OracleDiffRoll(hp, make_state_dict(hp))
    x1, w1, n1 = make_inputs(1, TIMESTEPS, seed=123, n_noise=1)
    with torch.no_grad():
        t0 = time.perf_counter()
        orc.reverse_diffusion(x1, w1, TIMESTEPS - 1, noise=n1[0])
        t1 = time.perf_counter() - t0
    b = fixed_batch
    if b is None:
        b = 1
        while b < BATCH and (n_steps + warmup) * (2 * b) * t1 <= budget_s:
            b *= 2
    x, w, nz = make_inputs(b, TIMESTEPS, seed=123, n_noise=n_steps + warmup)
    with torch.no_grad():
        i = 0
        t_index = TIMESTEPS - 1
        for _ in range(warmup):
            x, _ = orc.reverse_diffusion(x, w, t_index, noise=nz[i]); i += 1; t_index -= 1
        t0 = time.perf_counter()
        for _ in range(n_steps):
            x, _ = orc.reverse_diffusion(x, w, t_index, noise=nz[i]); i += 1; t_index -= 1
            _ = x.detach().cpu().numpy()
        dt = time.perf_counter() - t0
    val = n_steps / dt * (b / BATCH)
    desc = (f"{n_steps} timed steps (after {warmup} warm-up) of inpainting_ddpm_x0 at batch {b} on {threads} threads, "
            f"{dt:.1f} s; steps/s scaled by {b}/{BATCH} to the B={BATCH} workload (cost is linear in batch)")
    return val, desc, threads


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    val, desc, threads = cpu_steps_per_s(steps, warm, budget_s=150.0)
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
def run_b200(args, rank, world, local):
    import torch
    import torch.distributed as dist
    import diffroll_b200 as M
    from diffroll_b200 import _lib, build as _build
    from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict

    if not os.path.exists(_build.LIB):       # a checkout without the (git-ignored) built library: compile it once
        if local == 0:
            _build.build()
        if world > 1:
            import torch.distributed as _d
            _d.barrier()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    K, W = args.steps, max(args.warmup, 0)
    hp = default_hparams()
    model = M.ClassifierFreeDiffRoll(**hp, precision=args.precision)
    model.load_state_dict(make_state_dict(hp))
    model = model.cuda().eval()
    lib = _lib.load()

    # every rank: its own batch of 32 rolls (weak scaling), seeded per rank
    x_T, wav, _ = make_inputs(args.batch, TIMESTEPS, seed=123 + rank, n_noise=0)
    x_host, w_host = x_T.pin_memory(), wav.pin_memory()
    ups, branches, masks = model._all_updates()

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident run -----------------------------------------------------------------------------
    x_dev, w_dev = x_host.to(dev), w_host.to(dev)
    eng, xx, spec = model._prepare(x_dev, w_dev, branches, *masks)
    n_noise = min(max(K, W, 3, 20), TIMESTEPS)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    noise = torch.randn((n_noise,) + tuple(xx.shape), device=dev, generator=gen)

    def run_steps(n):
        """n timesteps as chains of <= TIMESTEPS steps starting at t = T-1 (x restarted per chain)."""
        done = 0
        while done < n:
            m = min(TIMESTEPS, n - done)
            x = xx.clone()
            eng.loop(x, noise, ups[:m], TIMESTEPS, TIMESTEPS - m)
            done += m

    run_steps(max(W, 3))
    torch.cuda.synchronize()
    barrier()   # NCCL's lazy communicator setup takes seconds: keep it out of the clock samples and the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    lib.drb_launch_count(1)
    ev0.record()
    run_steps(K)
    ev1.record()
    torch.cuda.synchronize(); barrier()
    launches = int(lib.drb_launch_count(0))
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if rank == 0 else None
    value = world * K / (ms / 1000.0)

    # ---- per-kernel CUDA-event timing inside a running loop (roofline) ------------------------------
    n_prof = min(K, 20)
    eng.profile(True)
    x = xx.clone()
    eng.loop(x, noise, ups[:n_prof], TIMESTEPS, TIMESTEPS - n_prof)
    prof = eng.profile_read()
    eng.profile(False)

    # ---- end to end through the public API, host buffers ---------------------------------------------
    Ke = min(K, TIMESTEPS)
    # warm-up with the same shape: allocates the pinned trajectory buffer the public API keeps per shape
    model.sample_loop(x_host.to(dev, non_blocking=True), w_host.to(dev, non_blocking=True), keep_trajectory=True, n_steps=Ke)
    torch.cuda.synchronize(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    xd = x_host.to(dev, non_blocking=True)
    wd = w_host.to(dev, non_blocking=True)
    x0, _, traj = model.sample_loop(xd, wd, keep_trajectory=True, n_steps=Ke)
    if world > 1:   # the path's only collective: one all-gather of the finished rolls over NCCL / NVLink
        from diffroll_b200.dist import all_gather_rolls
        rolls = all_gather_rolls(x0, world * args.batch)
        assert rolls.shape[0] == world * args.batch
    e1.record()
    torch.cuda.synchronize(); barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * Ke / (ms_e2e / 1000.0)
    h2d = (x_host.numel() + w_host.numel()) * 4 / Ke
    d2h = x_host.numel() * 4

    if rank != 0:
        return
    peaks = read_peaks()
    gate_ms, gate_n = prof["gate"]
    nb = 2 * args.batch
    gate_flops = nb * FLOP_DILATED                    # per launch, algorithmic (conditioner GEMM and x3 split excluded)
    gate_avg_s = gate_ms / max(gate_n, 1) / 1000.0
    achieved = gate_flops / gate_avg_s / 1e12 if gate_n else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get("umma_gate_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    step_ms_prof = sum(v[0] for v in prof.values()) / n_prof
    roofline = {
        "kernel": {"f16e5": "umma_gate_pers_kernel<3> (f16e5, persistent CTA pairs, tap window)",
                   "f16f8": "umma_gate_win_kernel<2> (f16f8, CTA pairs, tap window)",
                   "bf16x3": "umma_gate_pers_kernel<1> (bf16x3, persistent CTA pairs, tap window)"}.get(args.precision, f"umma_gate_kernel<{args.precision}>"),
        "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
        "frac": (achieved / peaks["bf16_sustained"]) if achieved else None, "traffic": traffic,
        "peak_source": f"{peaks['source']} bf16 dense, sustained (kernel timed inside a long step); burst {peaks['bf16']}",
        "mma_multiplicity": {"bf16x3": 3, "f16f8": 2, "f16e5": 2}.get(args.precision, 1),
        "algorithmic_flops_per_launch": gate_flops,
        "avg_launch_ms": gate_avg_s * 1e3, "launches_timed": gate_n,
        "share_of_step": (gate_ms / n_prof) / step_ms_prof if step_ms_prof else None,
        "per_step_ms": {k: v[0] / n_prof for k, v in prof.items()},
        "whole_step_algorithmic_tflops": FLOP_STEP_PER_ROLL * args.batch * (value / world) / 1e12,
    }
    cpu = None
    if not args.no_cpu_baseline:
        try:
            v, desc, threads = cpu_steps_per_s(2, 1, budget_s=25.0)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc}
        except Exception as e:  # the GPU numbers must still be printed
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16x3": "bf16x3 (bf16 hi/lo split, 3 tcgen05 products, fp32 accumulate)",
                  "f16f8": "f16f8 (fp16 tcgen05 product + e4m3 correction product, fp32 accumulate)",
                  "f16e5": "f16e5 (fp16 tcgen05 product + e5m2 correction product, one fp32 accumulator)"}.get(args.precision, args.precision),
        "data": "synthetic",
        "config": {"workload": f"configs[1]: batch={args.batch} per GPU, synthetic 640x88 rolls + 229-bin mel, "
                               "inpainting_ddpm_x0 w=0.5 (2 network forwards/step), timesteps=200, ClassifierFreeDiffRoll k=9, random weights",
                   "l2": "per-step working set (weights 0.6 GB + activations 0.6 GB) exceeds the 126 MB L2; no flush needed",
                   "parallelism": f"dp{world} (independent batch shards, no data-path collective; one all-gather of the "
                                  "finished rolls inside the e2e region when N>1)"},
        "sample_steps_per_s": value * args.batch,                 # B x steps/s, whole job
        "chain_wall_s": TIMESTEPS / (value / world),              # one 200-step chain of a 32-roll shard
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                "ms_per_step": ms_e2e / Ke},
        "roofline": roofline, "cpu_baseline": cpu,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--precision", default="f16e5", choices=["f16e5", "f16f8", "bf16x3", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

"""Multi-GPU tests (need >= 2 CUDA devices; `gpurun --gpus N`): an N-rank sharded run over NCCL returns exactly the
1-rank result per roll -- through dist.sample_sharded (pre-drawn and generator-drawn noise, ragged shards) and through
the sampling.py entry point under torchrun (every rank used to repeat one noise sequence)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DRB_ROOT"])
import diffroll_b200 as M
from diffroll_b200.dist import init_from_env, sample_sharded
from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict
rank, world, local = init_from_env("nccl")
torch.cuda.set_device(local)
B = int(os.environ["DRB_B"])
hp = default_hparams(timesteps=6)
m = M.ClassifierFreeDiffRoll(**hp); m.load_state_dict(make_state_dict(hp)); m = m.cuda().eval()
x_T, wav, noise = make_inputs(B, 6, seed=77, T=128, wav_len=65536)
x0, _ = sample_sharded(m, x_T, wav, noise)                       # pre-drawn global noise, sliced per rank
g = torch.Generator(device="cuda").manual_seed(4242)
x0g, _ = sample_sharded(m, x_T, wav, None, generator=g)          # noise drawn per step for the global batch, sliced
if rank == 0:
    torch.save({"pre": x0.cpu(), "gen": x0g.cpu()}, os.environ["DRB_OUT"])
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _torchrun(n, script_args, env):
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                           "--master-addr", "127.0.0.1", "--master-port", str(_free_port())] + script_args, env=env, timeout=900)


def _record(msg):
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "parity_numbers.log"), "a") as f:
        f.write(msg + "\n")
    print(msg)


# (ranks, global batch): even shards on 2 ranks; ragged shards (sizes differ by one, incl. 2,2,2,2,1,1,1,1) on 4 and 8
CASES = [(2, 4), (2, 3), (4, 6), (8, 12), (8, 32)]


@pytest.mark.parametrize("world,B", CASES)
def test_n_rank_sharded_run_equals_single_rank(tmp_path, world, B):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import diffroll_b200 as M
    from diffroll_b200.dist import shard_bounds
    from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict
    out = tmp_path / "x0.pt"
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DRB_ROOT=ROOT, DRB_OUT=str(out), DRB_B=str(B))
    _torchrun(world, [str(script)], env)
    got = torch.load(out)
    hp = default_hparams(timesteps=6)
    m = M.ClassifierFreeDiffRoll(**hp); m.load_state_dict(make_state_dict(hp)); m = m.cuda().eval()
    x_T, wav, noise = make_inputs(B, 6, seed=77, T=128, wav_len=65536)
    # the 1-rank reference runs every shard with the same per-rank batch size, so the same kernels see the same tiles:
    # the comparison is bit-exact
    parts, parts_g = [], []
    for r in range(world):
        lo, hi = shard_bounds(B, r, world)
        if hi == lo:
            continue
        parts.append(m.sample_loop(x_T[lo:hi].cuda(), wav[lo:hi].cuda(), noise=noise[:, lo:hi].cuda())[0].cpu())
        g = torch.Generator(device="cuda").manual_seed(4242)
        parts_g.append(m.sample_loop(x_T[lo:hi].cuda(), wav[lo:hi].cuda(), generator=g, shard=(B, lo, hi))[0].cpu())
    assert torch.equal(got["pre"], torch.cat(parts, 0))
    assert torch.equal(got["gen"], torch.cat(parts_g, 0))
    # and against ONE rank holding the whole batch.  Every roll's arithmetic is independent of its neighbours, but the
    # operand FORMAT is chosen per plan: f16n4 needs an even number of 128-frame tiles (CTA pairs), so a 1-roll shard runs
    # in f16e5 while the whole batch runs in f16n4 (or the other way round for an odd batch).  Equal formats -> equal bits;
    # different formats -> both within the parity bar of the fp32 reference, i.e. within 1e-3 of each other.
    precs = set()
    for eng, _ in m._engines.values():
        precs.add(eng.effective_precision)
    whole = m.sample_loop(x_T.cuda(), wav.cuda(), noise=noise.cuda())[0].cpu()
    g = torch.Generator(device="cuda").manual_seed(4242)
    whole_g = m.sample_loop(x_T.cuda(), wav.cuda(), generator=g)[0].cpu()
    for eng, _ in m._engines.values():
        precs.add(eng.effective_precision)
    e1, e2 = float((got["pre"] - whole).abs().max()), float((got["gen"] - whole_g).abs().max())
    _record(f"dist: {world} ranks x global batch {B} (NCCL all-gather): bit-identical to per-shard 1-rank runs; "
            f"vs one rank holding the whole batch max|delta| = {e1:.2e} (pre-drawn noise), {e2:.2e} (generator-drawn); "
            f"operand formats in play: {sorted(precs)}")
    tol = 0.0 if len(precs) == 1 else 1e-3 * max(1.0, float(whole.abs().max()))
    assert e1 <= tol and e2 <= tol, (e1, e2, sorted(precs))


def _run_sampling(tmp_path, world, tag):
    out = tmp_path / f"rolls_{tag}.pt"
    args = [os.path.join(ROOT, "sampling.py"), "task=transcription", "task.timesteps=5", "dataset.num_samples=6",
            "dataloader.batch_size=4", f"output_path={out}", "seed=11"]
    env = dict(os.environ)
    if world == 1:
        subprocess.check_call([sys.executable] + args, env=env, timeout=900)
    else:
        _torchrun(world, args, env)
    return torch.load(out)["rolls"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sampling_entry_point_torchrun_equals_single_process(tmp_path):
    """sampling.py (reference: sampling.py:22-73) with 2 ranks vs 1 process: same seed -> same rolls.  Batches of 4 + 2
    rolls: the second batch is split 1 + 1."""
    one = _run_sampling(tmp_path, 1, "one")
    two = _run_sampling(tmp_path, 2, "two")
    assert one.shape == two.shape == (6, 1, 640, 88)
    err = float((one - two).abs().max())
    _record(f"sampling.py: torchrun 2 ranks vs 1 process, 6 rolls x 5 steps: max|delta| = {err:.2e}")
    # the second batch (2 rolls) is split 1 + 1: one-roll shards compute in f16e5, the single process in f16n4 (see above)
    assert err < 1e-3 * max(1.0, float(one.abs().max()))
    # shards must not repeat one noise sequence: rolls of different shards differ
    assert float((two[0] - two[2]).abs().max()) > 1e-3

"""Multi-GPU tests (need >= 2 CUDA devices): an N-rank sharded run returns exactly the 1-rank result per sample."""
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
hp = default_hparams(timesteps=6)
m = M.ClassifierFreeDiffRoll(**hp); m.load_state_dict(make_state_dict(hp)); m = m.cuda().eval()
x_T, wav, noise = make_inputs(4, 6, seed=77, T=128, wav_len=65536)
x0, _ = sample_sharded(m, x_T, wav, noise)
if rank == 0:
    torch.save(x0.cpu(), os.environ["DRB_OUT"])
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_sharded_run_equals_single_rank(tmp_path):
    import diffroll_b200 as M
    from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict
    out = tmp_path / "x0.pt"
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DRB_ROOT=ROOT, DRB_OUT=str(out))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)], env=env, timeout=600)
    got = torch.load(out)
    hp = default_hparams(timesteps=6)
    m = M.ClassifierFreeDiffRoll(**hp); m.load_state_dict(make_state_dict(hp)); m = m.cuda().eval()
    x_T, wav, noise = make_inputs(4, 6, seed=77, T=128, wav_len=65536)
    # the 1-rank reference runs the same per-rank batch size (2), so the same kernels see the same tiles
    parts = [m.sample_loop(x_T[i:i + 2].cuda(), wav[i:i + 2].cuda(), noise=noise[:, i:i + 2].cuda())[0].cpu() for i in (0, 2)]
    assert torch.equal(got, torch.cat(parts, 0))
    whole = m.sample_loop(x_T.cuda(), wav.cuda(), noise=noise.cuda())[0].cpu()
    assert float((got - whole).abs().max()) < 1e-5

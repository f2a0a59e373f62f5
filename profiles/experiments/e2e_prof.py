"""One e2e sample_loop call (B=32, 3 steps) for an ncu launch list of the per-clip work (mel, conditioner tables)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import diffroll_b200 as M
from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict
hp = default_hparams()
m = M.ClassifierFreeDiffRoll(**hp); m.load_state_dict(make_state_dict(hp)); m = m.cuda().eval()
x, w, _ = make_inputs(32, hp["timesteps"], seed=3, n_noise=0)
x, w = x.cuda(), w.cuda()
for it in range(2):
    m.sample_loop(x, w.clone(), n_steps=3); torch.cuda.synchronize()   # a new waveform object: mel + conditioner tables run again
    if it == 0: torch.cuda.profiler.start()
torch.cuda.profiler.stop()

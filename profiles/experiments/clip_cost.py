"""Per-clip cost of the public sampling call: sample_loop(n_steps=1) on a new waveform object (mel + conditioner tables + 1 step),
and n_steps=20, CUDA events, 6 repetitions each."""
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
for n in (1, 20):
    ts = []
    for it in range(7):
        wc = w.clone(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.sample_loop(x, wc, n_steps=n, keep_trajectory=(n == 20)); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 2))
    print(f"DRB_LIN_PERS={os.environ.get('DRB_LIN_PERS', 'default')} n_steps={n}: ms per call {ts}")

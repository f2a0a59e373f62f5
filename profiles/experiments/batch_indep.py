"""Why does a 2-rank run differ by 2e-4 from one rank holding the whole batch (tests/test_gpu_dist.py)?  Reproduce on one GPU."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import diffroll_b200 as M
from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict
for prec in ("f16n4", "f16e5"):
    hp = default_hparams(timesteps=6)
    m = M.ClassifierFreeDiffRoll(**hp, precision=prec); m.load_state_dict(make_state_dict(hp)); m = m.cuda().eval()
    x_T, wav, noise = make_inputs(4, 6, seed=77, T=128, wav_len=65536)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        whole = m.sample_loop(x_T.cuda(), wav.cuda(), noise=noise.cuda())[0].cpu()
        print(prec, "whole: precision now", m.precision, "range seen", getattr(m, "_range_seen", None), [str(c.message)[:90] for c in caught])
    m2 = M.ClassifierFreeDiffRoll(**hp, precision=prec); m2.load_state_dict(make_state_dict(hp)); m2 = m2.cuda().eval()
    parts = []
    for lo, hi in ((0, 2), (2, 4)):
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            parts.append(m2.sample_loop(x_T[lo:hi].cuda(), wav[lo:hi].cuda(), noise=noise[:, lo:hi].cuda())[0].cpu())
            print(prec, "shard", lo, hi, "precision now", m2.precision, "range seen", getattr(m2, "_range_seen", None), [str(c.message)[:90] for c in caught])
    d = (torch.cat(parts, 0) - whole).abs()
    print(prec, "max|delta| whole vs shards per roll:", [float(d[i].max()) for i in range(4)], "|x|max", float(whole.abs().max()))

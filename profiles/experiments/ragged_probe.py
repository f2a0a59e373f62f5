"""Why does the 3 x 100 ragged training case sit at 8e-3?  Seeds x arithmetic modes, against GPU autograd over the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import diffroll_b200 as M
from diffroll_b200.synthetic import default_hparams, make_labelled_batch, make_state_dict
from oracle.diffroll_oracle import OracleDiffRoll
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
hp = default_hparams(); hp["training"] = dict(mode="x_0"); hp["loss_type"] = "huber"
for seed in (21, 22, 23):
    frame, audio, t, noise = make_labelled_batch(B=3, T=100, wav_len=65536, seed=seed)
    mask = torch.tensor([1, 0, 0])
    batch = {"frame": frame.cuda(), "audio": audio.cuda()}
    orc = OracleDiffRoll(hp, make_state_dict(hp), device="cuda")
    losses, grads, _ = orc.train_step(batch, t, noise.cuda(), dropout_mask=mask)
    got = {}
    for tc in ("0", "15"):
        os.environ["DRB_TRAIN_TC"] = tc
        m = M.ClassifierFreeDiffRoll(**hp); m.load_state_dict(make_state_dict(hp)); m = m.cuda().train()
        total = m.training_step(batch, 0, t=t.cuda(), noise=noise.cuda(), dropout_mask=mask)
        worst, wn, wl2, wl2n = 0.0, "", 0.0, ""
        for name, p in m.named_parameters():
            ref = grads[name]
            err = float((p.grad - ref).abs().max()) / max(float(ref.abs().max()), 1e-12)
            l2 = float((p.grad - ref).norm()) / max(float(ref.norm()), 1e-12)
            if err > worst: worst, wn = err, name
            if l2 > wl2: wl2, wl2n = l2, name
        got[tc] = {n: p.grad.clone() for n, p in m.named_parameters()}
        print(f"seed {seed} TC={tc}: loss diff {abs(float(total) - float(losses['diffusion_loss'])):.2e} worst max {worst:.3e} ({wn}) worst L2 {wl2:.3e} ({wl2n})", flush=True)
        m.release_buffers()
    d = max(float((got["0"][n] - got["15"][n]).norm()) / max(float(got["0"][n].norm()), 1e-12) for n in got["0"])
    print(f"seed {seed}: TC=15 vs TC=0 worst rel L2 {d:.3e}")

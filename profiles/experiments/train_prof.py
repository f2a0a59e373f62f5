"""Two training steps (B=16 x 640 frames) for an ncu launch list of the second one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import diffroll_b200 as M
from diffroll_b200.synthetic import default_hparams, make_labelled_batch, make_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
hp = default_hparams(); hp["lr"] = 1e-4
frame, audio, t, noise = make_labelled_batch(B=B, T=640, wav_len=327680, seed=5)
t = (torch.arange(B) * 37) % 200
batch = {"frame": frame.cuda(), "audio": audio.cuda()}
mask = (torch.arange(B) % 4 == 1).long()
m = M.ClassifierFreeDiffRoll(**hp); m.load_state_dict(make_state_dict(hp)); m = m.cuda().train()
opt = m.configure_optimizers()[0]
for it in range(2):
    opt.zero_grad(); m.training_step(batch, 0, t=t.cuda(), noise=noise.cuda(), dropout_mask=mask); opt.step()
    torch.cuda.synchronize()
    if it == 0: torch.cuda.profiler.start()
torch.cuda.profiler.stop()

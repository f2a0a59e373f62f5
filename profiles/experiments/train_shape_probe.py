"""Gradient parity of the training step vs GPU autograd for a list of (B, T) shapes and DRB_TRAIN_TC masks (set in the environment)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import diffroll_b200 as M
from diffroll_b200.synthetic import default_hparams, make_labelled_batch, make_state_dict
from oracle.diffroll_oracle import OracleDiffRoll
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
hp = default_hparams(); hp["training"] = dict(mode="x_0"); hp["loss_type"] = "huber"
for B, T in ((3, 128), (2, 100), (3, 100), (4, 192)):
    frame, audio, t, noise = make_labelled_batch(B=B, T=T, wav_len=131072, seed=21)
    mask = torch.zeros(B, dtype=torch.long); mask[0] = 1
    batch = {"frame": frame.cuda(), "audio": audio.cuda()}
    m = M.ClassifierFreeDiffRoll(**hp); m.load_state_dict(make_state_dict(hp)); m = m.cuda().train()
    total = m.training_step(batch, 0, t=t.cuda(), noise=noise.cuda(), dropout_mask=mask)
    orc = OracleDiffRoll(hp, make_state_dict(hp), device="cuda")
    losses, grads, _ = orc.train_step(batch, t, noise.cuda(), dropout_mask=mask)
    worst, wn = 0.0, ""
    for name, p in m.named_parameters():
        ref = grads[name]; err = float((p.grad - ref).abs().max()) / max(float(ref.abs().max()), 1e-12)
        if err > worst: worst, wn = err, name
    print(f"TC={os.environ.get('DRB_TRAIN_TC','default')} B={B} T={T}: loss diff {abs(float(total)-float(losses['diffusion_loss'])):.2e} worst grad {worst:.3e} ({wn})")
    m.release_buffers()

"""Training-step timing (row f3): ours (training_step + fused Adam, fp32 CUDA cores) vs torch autograd over the oracle + torch.optim.Adam
eagerly on the same GPU (cuDNN / cuBLAS; fp32 with TF32 off, then TF32 on).  usage: python profiles/experiments/train_bench.py [B] [iters]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import diffroll_b200 as M
from diffroll_b200.synthetic import default_hparams, make_labelled_batch, make_state_dict
from oracle.diffroll_oracle import OracleDiffRoll

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
hp = default_hparams(); hp["lr"] = 1e-4
frame, audio, t, noise = make_labelled_batch(B=B, T=640, wav_len=327680, seed=5)
t = (torch.arange(B) * 37) % 200      # make_labelled_batch carries 8 steps: one per roll for any B
batch = {"frame": frame.cuda(), "audio": audio.cuda()}
mask = (torch.arange(B) % 4 == 1).long()
res = {"batch": B, "frames": 640}

m = M.ClassifierFreeDiffRoll(**hp); m.load_state_dict(make_state_dict(hp)); m = m.cuda().train()
opt = m.configure_optimizers()[0]
def ours():
    opt.zero_grad(); m.training_step(batch, 0, t=t.cuda(), noise=noise.cuda(), dropout_mask=mask); opt.step()
ours(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters): ours()
e1.record(); torch.cuda.synchronize()
res["ours_ms"] = e0.elapsed_time(e1) / iters
res["train_workspace_bytes"] = list(m._train_engines.values())[0].workspace_bytes
m.release_buffers(); del m, opt; torch.cuda.empty_cache()

for name, flag in (() if "noeager" in sys.argv else (("eager_fp32_ms", False), ("eager_tf32_ms", True))):
    torch.backends.cudnn.allow_tf32 = flag; torch.backends.cuda.matmul.allow_tf32 = flag
    orc = OracleDiffRoll(hp, make_state_dict(hp), device="cuda")
    params = {k: torch.nn.Parameter(v.clone()) for k, v in orc.sd.items() if not k.startswith("mel_layer.")}
    ropt = torch.optim.Adam(params.values(), lr=hp["lr"])
    def ref():
        orc.sd.update({k: q.data for k, q in params.items()})
        losses, grads, _ = orc.train_step(batch, t, noise.cuda(), dropout_mask=mask)
        for k, q in params.items(): q.grad = grads[k]
        ropt.step()
    ref(); torch.cuda.synchronize()
    e0.record()
    for _ in range(iters): ref()
    e1.record(); torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1) / iters
    del orc, params, ropt; torch.cuda.empty_cache()
print(json.dumps(res))

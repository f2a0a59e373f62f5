"""Accuracy of the tensor-core products of the TRAINING forward: pred with DRB_TRAIN_TC masks vs the all-fp32 CUDA-core forward."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import torch
    import diffroll_b200 as M
    from diffroll_b200.synthetic import default_hparams, make_labelled_batch, make_state_dict
    hp = default_hparams()
    frame, audio, t, noise = make_labelled_batch(B=2)
    m = M.ClassifierFreeDiffRoll(**hp); m.load_state_dict(make_state_dict(hp)); m = m.cuda().train()
    x = noise.cuda()
    pred, spec = m._forward_training(x, audio.cuda(), t.cuda(), False, None, None, dropout_mask=torch.tensor([0, 1]))
    eng = list(m._train_engines.values())[0]
    torch.save({"pred": pred.cpu()}, sys.argv[1])
else:
    import torch
    outs = {}
    for mask in ("0", "1", "8", "9"):
        path = f"/tmp/pred_{mask}.pt"
        subprocess.check_call([sys.executable, __file__, path], env=dict(os.environ, DRB_TRAIN_TC=mask))
        outs[mask] = torch.load(path)["pred"]
    ref = outs["0"]
    for mask in ("1", "8", "9"):
        d = (outs[mask] - ref).abs()
        print(f"mask {mask}: max|delta pred| = {float(d.max()):.3e}  (|pred|max {float(ref.abs().max()):.3f})")

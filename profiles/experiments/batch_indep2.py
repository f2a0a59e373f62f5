"""f16n4: determinism and batch independence of ONE forward (cond branch only, then both branches), per layer."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import diffroll_b200 as M
from diffroll_b200 import _lib
from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict
hp = default_hparams(timesteps=6)
sd = make_state_dict(hp)
x_T, wav, noise = make_inputs(4, 6, seed=77, T=128, wav_len=65536)
def run(lo, hi, branches, layers=15):
    m = M.ClassifierFreeDiffRoll(**hp, precision="f16n4"); m.load_state_dict(sd); m = m.cuda().eval()
    x = x_T[lo:hi].cuda(); w = wav[lo:hi].cuda()
    eng, xx, _ = m._prepare(x, w, branches)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(eng.lib.drb_in_proj(eng.plan, C.c_void_p(xx.data_ptr()), 3, s), "in_proj")
    outs = [eng.buffer("x32").clone()]
    for l in range(layers):
        _lib.check(eng.lib.drb_resblock_forward(eng.plan, l, 3, s), "res")
        torch.cuda.synchronize()
        outs.append(eng.buffer("x32").clone())
    eff = eng.effective_precision
    m.release_buffers()
    return outs, eff
for br, name in ((_lib.BRANCH_COND, "cond"), (_lib.BRANCH_COND_UNCOND, "cond+uncond")):
    a, ea = run(0, 4, br); b, eb = run(0, 4, br); c, ec = run(0, 2, br)
    per = 128 * 512
    print(name, "effective precisions", ea, eb, ec)
    for l in range(0, 15):
        same = float((a[l] - b[l]).abs().max())
        n = c[l].numel() // (2 if br == _lib.BRANCH_COND_UNCOND else 1)
        # rolls 0,1 of the whole batch vs the 2-roll shard (conditional branch rows come first)
        d = float((a[l].flatten()[:2 * per] - c[l].flatten()[:2 * per]).abs().max())
        print(f"  after layer {l:2d}: rerun max|delta| {same:.3e}   B=4 vs B=2 shard (cond rolls 0,1) {d:.3e}")
        if l > 3 and d > 0: break

"""Experiment: does running the batch as TWO half-batches on two CUDA streams (two host threads, each in its own
drb_sample_loop) hide the tile-quantisation tails of the persistent kernels (640 pair-tiles on 74 SM pairs = 8.65 -> 9
rounds) by letting one half's kernels fill the SMs the other half's last round leaves idle?   python two_stream.py"""
import os, sys, threading, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import diffroll_b200 as M
from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict

STEPS = int(os.environ.get("STEPS", "100"))
hp = default_hparams()
sd = make_state_dict(hp)


def build(batch, seed):
    m = M.ClassifierFreeDiffRoll(**hp)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x_T, wav, _ = make_inputs(batch, 200, seed=seed, n_noise=0)
    ups, branches, masks = m._all_updates()
    eng, xx, _ = m._prepare(x_T.cuda(), wav.cuda(), branches, *masks)
    noise = torch.randn((STEPS,) + tuple(xx.shape), device="cuda")
    return m, eng, xx, noise, ups


def run(eng, xx, noise, ups, stream, n):
    with torch.cuda.stream(stream):
        x = xx.clone()
        eng.loop(x, noise, ups[:n], 200, 200 - n)


def timed(jobs):
    for j in jobs:                       # warm-up, sequential (lazy one-time initialisation inside the library)
        run(*j, 5); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    ths = [threading.Thread(target=run, args=(*j, STEPS)) for j in jobs]
    for t in ths: t.start()
    for t in ths: t.join()
    for j in jobs: torch.cuda.current_stream().wait_stream(j[4])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


res = {}
m, eng, xx, noise, ups = build(32, 1)
ms = timed([(eng, xx, noise, ups, torch.cuda.Stream())])
res["one_stream_b32_steps_per_s"] = STEPS / ms * 1e3
del m, eng, xx, noise
torch.cuda.empty_cache()
a = build(16, 2); b = build(16, 3)
ms = timed([(a[1], a[2], a[3], a[4], torch.cuda.Stream()), (b[1], b[2], b[3], b[4], torch.cuda.Stream())])
res["two_streams_2xb16_steps_per_s"] = STEPS / ms * 1e3
ms = timed([(a[1], a[2], a[3], a[4], torch.cuda.Stream())])
res["one_stream_b16_steps_per_s_(16 rolls only)"] = STEPS / ms * 1e3
del a, b
torch.cuda.empty_cache()
q = [build(8, 10 + i) for i in range(4)]
ms = timed([(j[1], j[2], j[3], j[4], torch.cuda.Stream()) for j in q])
res["four_streams_4xb8_steps_per_s"] = STEPS / ms * 1e3
res["pdl"] = os.environ.get("DRB_NO_PDL", "0") != "1"
print(json.dumps(res))

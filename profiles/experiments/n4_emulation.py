"""Emulation (CPU, fp64 accumulate) of the operand formats of the gate GEMM on real layer inputs: relative rms error of
y = dilated_conv(x + d) for  fp16 only | f16e5 (fp16 + e5m2 correction) | f16n4 (fp16 + block-scaled e2m1 correction with
ue4m3 scales per 16 channels, the scales and global powers of two the kernel uses).  Round-2 groundwork for the nvfp4
correction product; run in the build container:  python profiles/experiments/n4_emulation.py"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict
import torch.nn.functional as F

E2M1 = torch.tensor([0., .5, 1., 1.5, 2., 3., 4., 6.], dtype=torch.float64)

def q_e2m1(v):            # round to nearest e2m1 magnitude grid, saturating at 6
    a = v.abs().clamp(max=6.0)
    idx = (a.unsqueeze(-1) - E2M1).abs().argmin(-1)
    return torch.sign(v) * E2M1[idx]

def q_fp8(v, fmt):
    return v.to(torch.float32).to(fmt).to(torch.float64)

def q_ue4m3_up(s):        # smallest e4m3 >= s (scale must cover the block max)
    q = s.to(torch.float32).to(torch.float8_e4m3fn).to(torch.float64)
    bump = q < s
    bits = s.to(torch.float32).to(torch.float8_e4m3fn).view(torch.uint8).to(torch.int16)
    up = (bits + 1).clamp(max=126).to(torch.uint8).view(torch.float8_e4m3fn).to(torch.float64)
    return torch.where(bump, up, q)

def block_n4(v, block=16):     # v [..., K] with K % 16 == 0 -> dequantised block-scaled e2m1
    shp = v.shape
    b = v.reshape(*shp[:-1], shp[-1] // block, block)
    amax = b.abs().amax(-1, keepdim=True)
    sf = q_ue4m3_up(amax / 6.0)
    sf_safe = torch.where(sf > 0, sf, torch.ones_like(sf))
    return (q_e2m1(b / sf_safe) * sf).reshape(shp)

def main():
    torch.manual_seed(0)
    hp = default_hparams()
    sd = make_state_dict(hp)
    # a real layer input: run the first layers of the network in fp32 on a short clip
    from oracle.diffroll_oracle import OracleDiffRoll
    orc = OracleDiffRoll(hp, sd)
    x_T, wav, _ = make_inputs(1, 200, seed=5, n_noise=0, T=128, wav_len=65536)
    res = {}
    with torch.no_grad():
        sdd = orc.sd
        x = F.relu(F.conv1d(x_T.squeeze(1).transpose(1, 2), sdd["input_projection.weight"], sdd["input_projection.bias"]))
        e = orc.embedding[torch.tensor([150])]
        e = F.linear(e, sdd["diffusion_embedding.projection1.weight"], sdd["diffusion_embedding.projection1.bias"]); e = e * torch.sigmoid(e)
        e = F.linear(e, sdd["diffusion_embedding.projection2.weight"], sdd["diffusion_embedding.projection2.bias"]); e = e * torch.sigmoid(e)
        spec = orc.spec_frontend(wav)[..., :128]
        for layer in range(6):
            p = f"residual_layers.{layer}."
            d = orc.dilations[layer]
            dstep = F.linear(e, sdd[p + "diffusion_projection.weight"], sdd[p + "diffusion_projection.bias"]).unsqueeze(-1)
            xin = (x + dstep)[0].T.double()                       # [T, C]
            W = sdd[p + "dilated_conv.weight"].double()           # [2C, C, k]
            T, C = xin.shape
            pad = 4 * d
            xp = F.pad(xin.T, (pad, pad)).T                        # [T + 8d, C]
            cols = torch.stack([xp[j * d: j * d + T] for j in range(9)], 1).reshape(T, 9 * C)      # tap-major K
            Wk = W.permute(0, 2, 1).reshape(W.shape[0], 9 * C)
            y = cols @ Wk.T
            # operand splits
            a_hi = cols.to(torch.float16).double(); a_lo = cols - a_hi
            sw = 2.0 ** torch.floor(torch.log2(32768.0 / Wk.abs().max()))
            Ws = Wk * sw
            w_hi = Ws.to(torch.float16).double(); w_lo = Ws - w_hi
            main = a_hi @ w_hi.T
            def rel(yq):
                return float(((yq / sw - y).pow(2).mean() / y.pow(2).mean()).sqrt())
            out = {"fp16_only": rel(main)}
            # f16e5: [lo*2^4 | hi*2^-8] . [hi*2^-4 | lo*2^8] in e5m2 (weights unscaled there; scale-free comparison)
            c5 = q_fp8(a_lo * 16, torch.float8_e5m2) @ q_fp8(w_hi / 16, torch.float8_e5m2).T + \
                 q_fp8(a_hi / 256, torch.float8_e5m2) @ q_fp8(w_lo * 256, torch.float8_e5m2).T
            out["f16e5"] = rel(main + c5)
            # f16n4: block-scaled e2m1, A_LO 2^10 / W_HI 2^-10, A_HI 2^4 / W_LO 2^-4
            c4 = block_n4(a_lo * 1024) @ block_n4(w_hi / 1024).T + block_n4(a_hi * 16) @ block_n4(w_lo / 16).T
            out["f16n4"] = rel(main + c4)
            out["exact_split"] = rel(main + a_lo @ w_hi.T + a_hi @ w_lo.T)
            out["act_absmax"] = float(cols.abs().max())
            res[f"layer{layer}"] = out
            # advance the network (fp32, reference ops)
            yy = F.conv1d(x + dstep, sdd[p + "dilated_conv.weight"], sdd[p + "dilated_conv.bias"], padding=pad, dilation=d)
            yy = yy + F.conv1d(spec, sdd[p + "conditioner_projection.weight"], sdd[p + "conditioner_projection.bias"])
            g, f = torch.chunk(yy, 2, 1)
            z = torch.sigmoid(g) * torch.tanh(f)
            o = F.conv1d(z, sdd[p + "output_projection.weight"], sdd[p + "output_projection.bias"])
            r, _ = torch.chunk(o, 2, 1)
            x = (x + r) / 2 ** 0.5
    print(json.dumps(res, indent=1))

if __name__ == "__main__":
    main()

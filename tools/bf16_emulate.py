"""CPU emulation of where the engine rounds, to rank the rounding sources of the SegModel forward against north_star's
bf16 bound (rel-L2 <= 1e-2) WITHOUT a GPU.  Runs the oracle network in fp32 and injects roundings at the places the
kernels round: GEMM-operand weights, the pre-normalisation conv output y, the normalised activation a.

    python tools/bf16_emulate.py [plan] [patch] --y bf16|fp16|fp32 --a bf16 --w bf16

Test infrastructure / development tool (imports oracle/); not part of the product path.
"""
from __future__ import annotations

import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import seg_model as ref_seg  # noqa: E402
from oracle.parity import rel_l2  # noqa: E402

DT = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": None}


def rnd(t, kind):
    d = DT[kind]
    return t if d is None else t.to(d).float()


def block(m, x, cfg, first=False, wkind=None):
    """ConvDropoutNormReLU with the engine's roundings (bias dropped: exact under InstanceNorm)."""
    w = m.conv.weight if first else rnd(m.conv.weight, wkind or cfg.w)  # the 1-channel stem keeps fp32 accuracy (hi/lo split)
    y = F.conv3d(x, w, None, m.conv.stride, m.conv.padding)
    mean = y.mean((2, 3, 4), keepdim=True)                     # statistics from the fp32 accumulators
    var = y.var((2, 3, 4), keepdim=True, unbiased=False)
    y = rnd(y, cfg.y)
    a = (y - mean) * torch.rsqrt(var + m.norm.eps) * m.norm.weight.view(1, -1, 1, 1, 1) + m.norm.bias.view(1, -1, 1, 1, 1)
    a = F.leaky_relu(a, 0.01)
    return rnd(a, cfg.a)


def forward(net, x, cfg):
    skips = []
    first = True
    for st in net.encoder.stages:
        for m in st[0].convs:
            x = block(m, x, cfg, first)
            first = False
        skips.append(x)
    low = skips[-1]
    dec = net.decoder
    for s, (up, stage) in enumerate(zip(dec.transpconvs, dec.stages)):
        u = rnd(F.conv_transpose3d(low, rnd(up.weight, cfg.w), up.bias, up.stride), cfg.cat)
        low = torch.cat((u, rnd(skips[-(s + 2)], cfg.cat)), 1)   # the [up | skip] concat buffer and its consumer's operands
        for i, m in enumerate(stage.convs):
            low = block(m, low, cfg, wkind=cfg.cat if i == 0 else None)
    seg = dec.seg_layers[-1]
    out = F.conv3d(low, seg.weight, seg.bias)                  # fp32 weights, bf16 activations, fp32 accumulate
    f = F.interpolate(low, scale_factor=(net.upscale, 1, 1), mode="trilinear", align_corners=True)
    f = rnd(f, cfg.up)
    c0, c2 = net.sr_head[0], net.sr_head[2]
    h = rnd(F.relu(F.conv3d(f, rnd(c0.weight, cfg.sr), c0.bias, 1, 1)), cfg.sr)
    hr = F.conv3d(h, rnd(c2.weight, cfg.sr), c2.bias, 1, 2)
    return out, hr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("plan", nargs="?", default="3d_fullres")
    ap.add_argument("patch", nargs="?", default="64,64,64")
    ap.add_argument("--y", default="bf16")
    ap.add_argument("--a", default="bf16")
    ap.add_argument("--w", default="bf16")
    ap.add_argument("--up", default="bf16")
    ap.add_argument("--sr", default="bf16", help="sr_head operands (weights and the hidden activation)")
    ap.add_argument("--cat", default="bf16", help="dtype of the decoder concat buffer and of the conv that reads it")
    ap.add_argument("--seed", type=int, default=0)
    cfg = ap.parse_args()
    patch = tuple(int(v) for v in cfg.patch.split(","))
    net = ref_seg.build(cfg.plan)
    x = torch.randn((1, 1, *patch), generator=torch.Generator().manual_seed(cfg.seed))
    with torch.no_grad():
        out_r, hr_r = net(x)
        out_e, hr_e = forward(net, x, cfg)
    agree = float((out_e.argmax(1) == out_r.argmax(1)).double().mean())
    print(f"{cfg.plan} {patch} y={cfg.y} a={cfg.a} w={cfg.w} up={cfg.up}: logits {rel_l2(out_e, out_r):.3e}  "
          f"hr_logits {rel_l2(hr_e, hr_r):.3e}  argmax {agree:.5f}")


if __name__ == "__main__":
    main()

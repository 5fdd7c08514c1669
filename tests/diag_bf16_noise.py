"""Diagnostic: how much of the engine-vs-fp32 difference is inherent to bf16 arithmetic?  Compares, against the fp32
oracle on the same GPU (TF32 off): (a) this engine, (b) the oracle under torch.autocast(bf16) (cuDNN)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import seg_model as ref_seg
from rehrseg_b200 import seg_model as sm

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def run(model, x, g1, g2, autocast=False):
    for p in model.parameters():
        p.grad = None
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out, up = model(x)
    else:
        out, up = model(x)
    loss = (out.float() * g1).sum() / out.numel() + (up.float() * g2).sum() / up.numel()
    loss.backward()
    return out.float().detach(), up.float().detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


for plan, patch, batch in (("tiny", (16, 32, 32), 2), ("3d_fullres", (64, 64, 64), 1)):
    ref = ref_seg.build(plan).cuda()
    mine = sm.SegModel(**ref_seg.plan_kwargs(plan)).cuda()
    mine.load_state_dict(ref.state_dict())
    x = torch.randn((batch, 1, *patch), generator=torch.Generator().manual_seed(0)).cuda()
    o32, u32, g32 = run(ref, x, 1, 1) if False else (None, None, None)
    out_shape = ref(x)[0].shape
    up_shape = ref(x)[1].shape
    g1 = torch.randn(out_shape, generator=torch.Generator().manual_seed(1)).cuda()
    g2 = torch.randn(up_shape, generator=torch.Generator().manual_seed(2)).cuda()
    o32, u32, g32 = run(ref, x, g1, g2)
    for name, (o, u, g) in (("engine", run(mine, x, g1, g2)), ("autocast_bf16", run(ref, x, g1, g2, autocast=True))):
        num = den = 0.0
        worst = (0.0, "")
        for k in g32:
            if k.endswith("conv.bias") and ".convs." in k:
                continue
            d = (g[k].double() - g32[k].double())
            num += float(d.pow(2).sum()); den += float(g32[k].double().pow(2).sum())
            r = rel(g[k], g32[k])
            if r > worst[0]:
                worst = (r, k)
        print(f"{plan:10s} {name:14s} logits {rel(o, o32):.4e} hr {rel(u, u32):.4e} argmax {float((o.argmax(1) == o32.argmax(1)).double().mean()):.5f} "
              f"grads_global {(num / den) ** 0.5:.4e} worst {worst[0]:.3e} {worst[1]}", flush=True)
    # per-parameter table for the engine
    o, u, g = run(mine, x, g1, g2)
    for k in g32:
        if "all_modules" in k or "decoder.encoder" in k:
            continue
        print(f"    {k:60s} {rel(g[k], g32[k]):.3e}")

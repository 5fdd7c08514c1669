"""Development tool: device time of the loss pieces of the C4 joint step (train_step.py) at full size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import train_step as ts

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
B, D, HW = 2, 16, 256


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


lo = torch.randn((B, 2, D, HW, HW), device=dev, generator=g, requires_grad=True)
hi = torch.randn((B, 2, 4 * D, HW, HW), device=dev, generator=g, requires_grad=True)
lab_lo = (torch.rand((B, 1, D, HW, HW), device=dev, generator=g) > 0.8).float()
lab_hi = (torch.rand((B, 1, 4 * D, HW, HW), device=dev, generator=g) > 0.8).float()
unc = torch.rand((B, 1, D, HW, HW), device=dev, generator=g) * 0.99 + 0.01
fs = torch.randn((B, 64, D, HW // 2, HW // 2), device=dev, generator=g, requires_grad=True)
ft = torch.randn((B, 64, D, HW // 2, HW // 2), device=dev, generator=g)
lr_obj, hr_obj = ts.build_loss(False, 0), ts.build_loss(False, 1)
dist = ts.Distiller(64, 64, 0.0, 1.0, 1.0).to(dev)


def fb(f):
    def run():
        for t in (lo, hi, fs):
            t.grad = None
        f().backward()
    return run


print("lr loss (CE x unc)      fwd+bwd %.3f ms" % timed(fb(lambda: lr_obj(lo, lab_lo, unc))))
print("hr loss (CE + Dice)     fwd+bwd %.3f ms" % timed(fb(lambda: hr_obj(hi, lab_hi, None))))
print("distiller               fwd+bwd %.3f ms" % timed(fb(lambda: dist(fs, ft))))
print("  structure_loss        fwd+bwd %.3f ms" % timed(fb(lambda: ts.structure_loss(fs, ft, 0.5))))
print("  1x1x1 conv + cosine   fwd+bwd %.3f ms" % timed(fb(lambda: ts.cosine_distance_loss(dist.distill(fs), ft))))
print("  cosine only           fwd+bwd %.3f ms" % timed(fb(lambda: ts.cosine_distance_loss(fs, ft))))

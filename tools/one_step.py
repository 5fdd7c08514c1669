"""One C1 step (PlainConvUNet 2x1x128^3 fwd+bwd, bf16 engine) after `--warm` untimed steps: the command ncu wraps for the
per-launch list (profiles/).  `--segmodel` runs the full SegModel (with the x4 SR head) instead."""
import sys, os, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import seg_model as sm, functional as Fn
ap = argparse.ArgumentParser()
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--segmodel", action="store_true")
a = ap.parse_args()
torch.manual_seed(0)
m = (sm.plainconv_3d_fullres() if a.segmodel else sm.plainconv_unet_3d_fullres()).cuda()
x = torch.randn(2, 1, 128, 128, 128, device='cuda')
g = torch.randn(2, 2, 128, 128, 128, device='cuda')
def step():
    for p in m.parameters():
        p.grad = None
    Fn.clear_weight_cache()
    out = m(x)
    if a.segmodel:
        loss = (out[0].float() * g).mean() + out[1].float().mean()
    else:
        loss = (out.float() * g).mean()
    loss.backward()
for i in range(a.warm):
    step()
torch.cuda.synchronize()
l0 = Fn.launches()
step()
torch.cuda.synchronize()
print("launches in the measured step:", Fn.launches() - l0)

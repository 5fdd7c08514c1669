"""Run every HBM-bound helper of configs 3 and 5 a few times at full size: the command `ncu` wraps for their per-kernel DRAM bytes
(sw_accumulate / sw_finalize, blur1d, rot90, fba_combine, mean_stack, upsample_d, pointwise head, layout adapters)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import functional as Fn, sliding_window as sw, volume_ops as vo

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(5)
for _ in range(2):
    logits = torch.zeros((2, 256, 256, 256), dtype=torch.half, device=dev)
    npred = torch.zeros((256, 256, 256), dtype=torch.half, device=dev)
    pred = torch.randn((2, 128, 128, 128), device=dev, generator=g).half()
    gauss = sw.importance_map([128, 128, 128], 1. / 8, 10, device=torch.device(dev))
    sw.sw_accumulate(logits, npred, pred, gauss, (64, 64, 64))
    npred += 1
    sw.sw_finalize(logits, npred)
    hr = torch.rand((160, 1, 512, 512), device=dev, generator=g)
    taps = torch.exp(-0.5 * ((torch.arange(9.) - 4) / (3.873 / 2.355)) ** 2)
    vo.blur_along_x(hr, (taps / taps.sum()).reshape(1, 1, 9, 1).to(dev))
    vols = [torch.rand((512, 512, 160), device=dev, generator=g) for _ in range(4)]
    vo.rotate_vol_2d(vols[0], 90)
    vo.mean_fuse(vols)
    vo.fba(vols, "infinity")
    vo.fba(vols, 2.0)
    feats = torch.randn((2, 128, 128, 128, 32), device=dev, generator=g).to(torch.bfloat16)
    Fn.upsample_linear_d(feats[:1], 512)
    w = torch.randn((2, 32, 1, 1, 1), device=dev)
    b = torch.zeros((2,), device=dev)
    Fn.seg_head(feats, w, b)
    Fn.from_channels_last(feats)
torch.cuda.synchronize()
print("done")

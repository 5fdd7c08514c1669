"""Development check: stride-2 marching weight gradient vs torch autograd over a few shapes (prints relative errors)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from rehrseg_b200 import functional as Fn
Fn.S2_WGRAD_MARCH_MIN_VOXELS = 0
torch.backends.cudnn.allow_tf32 = False
for n, cin, cout, dhw in [(1, 32, 64, (16, 128, 128)), (2, 32, 64, (32, 64, 64)), (1, 32, 64, (33, 64, 64)), (1, 32, 64, (32, 66, 64)),
                          (1, 32, 64, (32, 64, 70)), (2, 32, 64, (33, 66, 70)), (2, 32, 32, (33, 66, 70)), (2, 64, 32, (33, 66, 70))]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn((n, cin, *dhw), generator=g).cuda().to(torch.bfloat16).float()
    w = torch.zeros((cout, cin, 3, 3, 3), device="cuda", requires_grad=True)
    y = F.conv3d(x, w, None, stride=2, padding=1)
    dy = torch.randn(y.shape, generator=g).cuda().to(torch.bfloat16).float()
    (ref,) = torch.autograd.grad(y, w, dy)
    xcl = x.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16)
    gcl = dy.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16)
    route = Fn.wgrad_route(xcl.shape, gcl.shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    dw = Fn.conv3d_wgrad_raw(xcl, gcl, w.shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    err = float((dw - ref).norm() / ref.norm())
    per_tap = ((dw - ref).pow(2).sum((0, 1)) / ref.pow(2).sum((0, 1))).sqrt().flatten().tolist()
    print(n, cin, cout, dhw, route, f"rel {err:.2e}", "bad taps:", [i for i, e in enumerate(per_tap) if e > 1e-2], flush=True)

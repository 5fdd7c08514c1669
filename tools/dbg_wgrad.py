import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import functional as Fn
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
for (n, ci, co, dhw) in [(2, 32, 32, (16, 32, 32)), (2, 64, 32, (16, 32, 32)), (1, 32, 32, (32, 64, 64)), (2, 32, 32, (16, 32, 16)),
                         (2, 32, 32, (16, 16, 32)), (1, 32, 32, (16, 32, 32)), (2, 32, 32, (20, 32, 16)), (2, 32, 32, (128, 128, 128))]:
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((n, *dhw, ci), device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn((n, *dhw, co), device="cuda", generator=g).to(torch.bfloat16)
    ws = (co, ci, 3, 3, 3)
    Fn.USE_MARCH = True; a = Fn.conv3d_wgrad_raw(x, dy, ws, (3, 3, 3), (1, 1, 1), (1, 1, 1))
    Fn.USE_MARCH = False; b = Fn.conv3d_wgrad_raw(x, dy, ws, (3, 3, 3), (1, 1, 1), (1, 1, 1))
    torch.cuda.synchronize()
    e = (a - b).abs()
    print(n, ci, co, dhw, "rel", rel(a, b), "per-kd", [round(rel(a[:, :, k], b[:, :, k]), 4) for k in range(3)],
          "per-kh", [round(rel(a[:, :, :, k], b[:, :, :, k]), 4) for k in range(3)], "per-kw", [round(rel(a[..., k], b[..., k]), 4) for k in range(3)])

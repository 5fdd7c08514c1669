import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from rehrseg_b200 import functional as Fn
torch.backends.cudnn.allow_tf32 = False
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
n, ci, co, dhw = 2, 32, 32, (16, 32, 32)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((n, *dhw, ci), device="cuda", generator=g).to(torch.bfloat16)
dy = torch.randn((n, *dhw, co), device="cuda", generator=g).to(torch.bfloat16)
Fn.USE_MARCH = True
a = Fn.conv3d_wgrad_raw(x, dy, (co, ci, 3, 3, 3), (3, 3, 3), (1, 1, 1), (1, 1, 1))
w5 = torch.zeros((co, ci, 7, 7, 7), device="cuda", requires_grad=True)
y = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w5, padding=3)
(d5,) = torch.autograd.grad(y, w5, dy.float().permute(0, 4, 1, 2, 3))
for kd in range(3):
    for kh in range(3):
        for kw in range(3):
            best = (9, None)
            for a_ in range(7):
                for b_ in range(7):
                    for c_ in range(7):
                        r = rel(a[:, :, kd, kh, kw], d5[:, :, a_, b_, c_])
                        if r < best[0]: best = (r, (a_ - 2, b_ - 2, c_ - 2))
            print((kd, kh, kw), "best match", best[1], round(best[0], 4))

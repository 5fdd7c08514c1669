"""Development tool: kernel-level breakdown (torch.profiler, CUDA activities) of one fwd+bwd of the config-4 student:
anisotropic SegModel (x4 SR head) on [2,1,16,256,256]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from rehrseg_b200 import seg_model as sm
torch.manual_seed(0)
kw = sm.fullres_kwargs()
kw.update(kernel_sizes=[[1, 3, 3], [1, 3, 3]] + [[3, 3, 3]] * 4, strides=[[1, 1, 1], [1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2], [1, 2, 2]])
m = sm.SegModel(upscale=4, **kw).cuda()
x = torch.randn(2, 1, 16, 256, 256, device="cuda")
def step():
    for p in m.parameters():
        p.grad = None
    out, up, skips = m(x, return_inetermediate_feature=True)
    (out.float().mean() + up.float().mean() + skips[1].float().mean()).backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted([(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0], key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device ms {tot:.3f}")
for k, ms, n in rows[:40]:
    print(f"{ms:9.3f} ms {100 * ms / tot:5.1f}% x{n:4d}  {k[:110]}")

"""Development tool: kernel-level breakdown of one C1 step (PlainConvUNet 2x1x128^3 fwd+bwd) with torch.profiler."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from rehrseg_b200 import seg_model as sm, functional as Fn

full = len(sys.argv) > 1 and sys.argv[1] == "segmodel"
torch.manual_seed(0)
m = (sm.plainconv_3d_fullres() if full else sm.plainconv_unet_3d_fullres()).cuda()
x = torch.randn(2, 1, 128, 128, 128, device='cuda')


def step():
    for p in m.parameters():
        p.grad = None
    Fn.clear_weight_cache()
    out = m(x)
    loss = (out[0].float().mean() + out[1].float().mean()) if full else out.float().mean()
    loss.backward()


for i in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device ms {tot:.3f}")
for k, ms, n in rows[:45]:
    print(f"{ms:9.3f} ms {100 * ms / tot:5.1f}% x{n:4d}  {k[:110]}")

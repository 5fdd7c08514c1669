"""Development tool: kernel-level breakdown of one FLAVR fwd+bwd step (C2) with torch.profiler."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from rehrseg_b200 import flavr
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
m = flavr.UNet_3D_3D(2, "unet_18", 4, 4, False, "concat", "transpose", use_uncertainty=False).cuda()
x = torch.rand((B, 2, 4, 256, 256), device="cuda")
def step():
    for p in m.parameters():
        p.grad = None
    m(x.clone()).float().mean().backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted([(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0], key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"B={B} total device ms {tot:.3f}")
for k, ms, n in rows[:30]:
    print(f"{ms:9.3f} ms {100 * ms / tot:5.1f}% x{n:4d}  {k[:120]}")

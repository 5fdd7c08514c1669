# scratch probe: C1 fwd+bwd timing per phase (not a bench line)
import time, torch, sys
sys.path.insert(0, '/root/repo')
from rehrseg_b200 import seg_model as sm, functional as Fn
torch.manual_seed(0)
m = sm.plainconv_3d_fullres().cuda()
x = torch.randn(2, 1, 128, 128, 128, device='cuda')
def step():
    out, up = m(x)
    loss = out.float().mean() + up.float().mean()
    loss.backward()
for i in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 5
for i in range(n): step()
e1.record(); torch.cuda.synchronize()
print("C1 full SegModel fwd+bwd ms/step:", e0.elapsed_time(e1) / n)
with torch.no_grad():
    for i in range(2): m(x)
    torch.cuda.synchronize(); e0.record()
    for i in range(n): m(x)
    e1.record(); torch.cuda.synchronize()
print("C1 fwd only ms:", e0.elapsed_time(e1) / n)
print("mem GB", torch.cuda.max_memory_allocated() / 1e9)

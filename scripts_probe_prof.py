# scratch: kernel-level breakdown of one C1 step with torch.profiler
import torch, sys
sys.path.insert(0, '/root/repo')
from torch.profiler import profile, ProfilerActivity
from rehrseg_b200 import seg_model as sm
torch.manual_seed(0)
m = sm.plainconv_3d_fullres().cuda()
x = torch.randn(2, 1, 128, 128, 128, device='cuda')
def step():
    out, up = m(x)
    loss = out.float().mean() + up.float().mean()
    loss.backward()
for i in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))

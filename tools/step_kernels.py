"""Development tool: per-launch device time of the conv-engine calls of one C1 step (CUDA events), with shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import seg_model as sm, functional as Fn
torch.manual_seed(0)
ANISO = "--aniso" in sys.argv   # config 4 student: anisotropic SegModel (x4 SR head) on [2,1,16,256,256]
if ANISO:
    kw = sm.fullres_kwargs()
    kw.update(kernel_sizes=[[1, 3, 3], [1, 3, 3]] + [[3, 3, 3]] * 4, strides=[[1, 1, 1], [1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2], [1, 2, 2]])
    m = sm.SegModel(upscale=4, **kw).cuda()
    x = torch.randn(2, 1, 16, 256, 256, device='cuda')
else:
    m = sm.plainconv_unet_3d_fullres().cuda()
    x = torch.randn(2, 1, 128, 128, 128, device='cuda')
def step():
    for p in m.parameters():
        p.grad = None
    out = m(x)
    if ANISO:
        (out[0].float().mean() + out[1].float().mean()).backward()
    else:
        out.float().mean().backward()
for i in range(3):
    step()
torch.cuda.synchronize()
with Fn.kernel_timer() as kt:
    torch.cuda._sleep(300_000_000)
    step()
rows = kt.rows()
tot = sum(r[2] for r in rows)
print(f"conv-engine calls: {len(rows)}, total {tot:.3f} ms")
for name, tag, ms, fl in sorted(rows, key=lambda r: -r[2]):
    print(f"{ms:8.3f} ms {fl / ms / 1e9:8.1f} TF/s  {name:26s} {tag}")

"""Development tool: one C1 step (fwd+bwd, weights re-packed) eager vs captured in ONE CUDA graph and replayed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import seg_model as sm, functional as Fn
torch.manual_seed(0)
m = sm.plainconv_unet_3d_fullres().cuda()
x = torch.randn(2, 1, 128, 128, 128, device='cuda')
g = torch.randn(2, 2, 128, 128, 128, device='cuda')
params = list(m.parameters())


def step():
    for p in params:
        p.grad = None
    Fn.clear_weight_cache()
    out = m(x)
    loss = torch.dot(out.float().reshape(-1), g.reshape(-1)) / out.numel()
    loss.backward()
    return loss


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


print("eager  %.3f ms/step" % timed(step))
l_eager = float(step())
grads_eager = [p.grad.clone() for p in params if p.grad is not None]
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2):
        step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
for p in params:
    p.grad = None
with torch.cuda.graph(graph):
    loss = step()
torch.cuda.synchronize()
print("graph  %.3f ms/step" % timed(graph.replay))
graph.replay()
torch.cuda.synchronize()
grads_graph = [p.grad for p in params if p.grad is not None]
print("loss eager %.6f graph %.6f" % (l_eager, float(loss)))
print("max grad diff", max(float((a - b).abs().max()) for a, b in zip(grads_eager, grads_graph)))

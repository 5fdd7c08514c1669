"""Development tool: the C1 step exactly as bench.py replays it (GraphedTrainStep, 2x1x128^3), ms/step over `--steps` replays
after a warm-up; prints one line.  For quick A/B runs of environment knobs (REHR_*) without bench.py's other legs.
`--check`: also compare loss and gradients of the replayed step with a plain python-launched step (same process)."""
import sys, os, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import seg_model as sm, functional as Fn
from rehrseg_b200.graphs import GraphedTrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--label", default="")
ap.add_argument("--check", action="store_true")
ap.add_argument("--segmodel", action="store_true")
a = ap.parse_args()
torch.manual_seed(0)
m = (sm.plainconv_3d_fullres() if a.segmodel else sm.plainconv_unet_3d_fullres()).cuda()
x = torch.randn(2, 1, 128, 128, 128, device="cuda")
g = torch.randn(2, 2, 128, 128, 128, device="cuda")


def loss_fn(out, g):
    if a.segmodel:
        return torch.dot(out[0].float().reshape(-1), g.reshape(-1)) / out[0].numel() + out[1].float().mean()
    return torch.dot(out.float().reshape(-1), g.reshape(-1)) / out.numel()


gs = GraphedTrainStep(m, loss_fn, (x, g))
for _ in range(3):
    gs.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    gs.replay()
e1.record()
torch.cuda.synchronize()
msg = f"{a.label or 'step'}: {e0.elapsed_time(e1) / a.steps:.3f} ms/step"
if a.check:
    loss_g = float(gs.replay())
    grads_g = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    defer, Fn.WGRAD_DEFER_JOIN = Fn.WGRAD_DEFER_JOIN, False
    for p in m.parameters():
        p.grad = None
    Fn.clear_weight_cache()
    loss_e = loss_fn(m(x), g)
    loss_e.backward()
    torch.cuda.synchronize()
    worst = max(float((p.grad - grads_g[n]).abs().max() / (p.grad.abs().max() + 1e-30)) for n, p in m.named_parameters() if p.grad is not None)
    msg += f" | loss graph {loss_g:.6f} eager {float(loss_e):.6f} | worst grad diff (rel. to max) {worst:.2e}"
print(msg)

"""Run ONE pass (fwd | dgrad | wgrad) of one C1 layer shape a few times: the command `ncu --set full -k regex:<kernel>` wraps.
    python tools/one_layer.py <cin> <cout> <spatial> <stride> <fwd|dgrad|wgrad> [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import functional as Fn
cin, cout, sp, st = (int(v) for v in sys.argv[1:5])
what = sys.argv[5]
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 2
n = 2
x = torch.randn((n, sp, sp, sp, cin), device="cuda").to(torch.bfloat16)
w = torch.randn((cout, cin, 3, 3, 3), device="cuda") / (27 * cin) ** 0.5
k, s, p = (3, 3, 3), (st,) * 3, (1, 1, 1)
y, _, _ = Fn.conv3d_raw(x, w, None, k, s, p, want_stats=True)
dy = torch.randn_like(y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    if what == "fwd":
        Fn.conv3d_raw(x, w, None, k, s, p, want_stats=True)
    elif what == "dgrad":
        Fn.conv3d_dgrad_raw(dy, w, x.shape, k, s, p)
    else:
        Fn.conv3d_wgrad_raw(x, dy, w.shape, k, s, p)
e1.record()
torch.cuda.synchronize()
print("done", what, cin, cout, sp, st, f"{e0.elapsed_time(e1) / iters * 1e3:.1f} us per launch")

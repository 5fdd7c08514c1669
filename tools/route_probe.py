"""Development tool: marching vs tapped-GEMM kernel on one stride-1 k3 layer shape (forward and input gradient), L2 flushed
between launches.  usage: route_probe.py cin cout spatial|DxHxW [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import functional as Fn
cin, cout = int(sys.argv[1]), int(sys.argv[2])
dims = [int(v) for v in sys.argv[3].split("x")]
dd, hh, ww = dims if len(dims) == 3 else dims * 3
sp = sys.argv[3]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
k, s, p = (3, 3, 3), (1, 1, 1), (1, 1, 1)
x = torch.randn((n, dd, hh, ww, cin), device="cuda").to(torch.bfloat16)
w = torch.randn((cout, cin, 3, 3, 3), device="cuda") / (27 * cin) ** 0.5
dy = torch.randn((n, dd, hh, ww, cout), device="cuda").to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
fl = 2.0 * n * dd * hh * ww * cin * cout * 27


def timed(fn, it=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(it):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / it * 1e3


Fn.MARCH_MAX_WIDE = 1 << 30          # compare the kernels themselves, not the routing rule
for use in (True, False):
    Fn.USE_MARCH = use
    Fn.clear_weight_cache()
    tf = timed(lambda: Fn.conv3d_raw(x, w, None, k, s, p, want_stats=True))
    td = timed(lambda: Fn.conv3d_dgrad_raw(dy, w, x.shape, k, s, p))
    tw = timed(lambda: Fn.conv3d_wgrad_raw(x, dy, w.shape, k, s, p))
    print(f"{cin}->{cout} @{sp} n={n} {'march ' if use else 'tapped'}: fwd {tf:7.1f} us ({fl / tf / 1e6:6.0f} TF/s)  dgrad {td:7.1f} us ({fl / td / 1e6:6.0f} TF/s)"
          f"  wgrad {tw:7.1f} us ({fl / tw / 1e6:6.0f} TF/s)")

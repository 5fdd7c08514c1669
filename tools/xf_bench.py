"""Micro-benchmark of the normalise-on-load variants of the marching kernels against their plain versions (same shapes, L2
flushed between iterations).  REHR_MARCH_DEBUG=8 / 16 isolate the pipeline-latency and the shared-memory-traffic share."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import functional as Fn

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=6):
    fn(); fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for cin, cout, d in ((64, 32, 128), (32, 32, 128), (64, 64, 64), (128, 64, 64)):
    n = 2
    g = torch.Generator().manual_seed(0)
    y = torch.randn((n, d, d, d, cin), generator=g, dtype=torch.float16).cuda().view(torch.bfloat16)
    norm = torch.ones((n, 3, cin), device="cuda"); norm[:, 1] = 0.1; norm[:, 2] = 0.01
    w = (torch.randn((cout, cin, 3, 3, 3), generator=g) / (27 * cin) ** 0.5).cuda()
    dy = torch.randn((n, d, d, d, cout), generator=g, dtype=torch.bfloat16).cuda()
    k, s, p = (3, 3, 3), (1, 1, 1), (1, 1, 1)
    fl = 2.0 * n * d ** 3 * cout * cin * 27
    abf = y.view(torch.float16).to(torch.bfloat16)
    t0 = timeit(lambda: Fn.conv3d_raw(y, w, None, k, s, p, want_stats=True, x_h=True, y_h=True))
    t1 = timeit(lambda: Fn.conv3d_raw(y, w, None, k, s, p, want_stats=True, x_h=True, y_h=True, norm=norm, op_h=True))
    t2 = timeit(lambda: Fn.conv3d_wgrad_raw(abf, dy, w.shape, k, s, p))
    t3 = timeit(lambda: Fn.conv3d_wgrad_raw(y, dy, w.shape, k, s, p, norm=norm, x_h=True))
    print(f"{cin}->{cout} @{d}^3: fwd {t0:.3f} ms ({fl / t0 / 1e9:.0f} TF/s) | fwd on-load {t1:.3f} ms ({fl / t1 / 1e9:.0f}) | "
          f"wgrad {t2:.3f} ms ({fl / t2 / 1e9:.0f}) | wgrad on-load {t3:.3f} ms ({fl / t3 / 1e9:.0f})", flush=True)

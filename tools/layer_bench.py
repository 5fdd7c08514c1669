"""Per-layer microbenchmark of the conv engine on the C1 layer shapes: fwd / dgrad / wgrad
device time (CUDA events, L2 flushed between iterations) and TFLOP/s.  Development tool, not the bench line."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import functional as Fn

LAYERS = [  # name, cin, cout, spatial (of the INPUT), stride
    ("enc0.1", 32, 32, 128, 1), ("enc1.0", 32, 64, 128, 2), ("enc1.1", 64, 64, 64, 1), ("enc2.0", 64, 128, 64, 2),
    ("enc2.1", 128, 128, 32, 1), ("enc3.0", 128, 256, 32, 2), ("enc3.1", 256, 256, 16, 1), ("enc4.0", 256, 320, 16, 2),
    ("enc4.1", 320, 320, 8, 1), ("enc5.0", 320, 320, 8, 2), ("enc5.1", 320, 320, 4, 1),
    ("dec0.0", 640, 320, 8, 1), ("dec1.0", 512, 256, 16, 1), ("dec2.0", 256, 128, 32, 1), ("dec3.0", 128, 64, 64, 1),
    ("dec4.0", 64, 32, 128, 1),
]
if os.environ.get("S2_WGRAD_MIN") is not None:
    Fn.S2_WGRAD_MARCH_MIN_VOXELS = int(os.environ["S2_WGRAD_MIN"])
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda._sleep(1_500_000)  # let the CPU run ahead so launch latency is not inside the timed region
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


only = sys.argv[1:]
rows = []
for name, cin, cout, sp, st in LAYERS:
    if only and name not in only:
        continue
    n = 2
    x = torch.randn((n, sp, sp, sp, cin), device="cuda").to(torch.bfloat16)
    w = (torch.randn((cout, cin, 3, 3, 3), device="cuda") / (27 * cin) ** 0.5)
    b = torch.randn((cout,), device="cuda")
    k, s, p = (3, 3, 3), (st,) * 3, (1, 1, 1)
    y, _, _ = Fn.conv3d_raw(x, w, b, k, s, p, want_stats=True)
    dy = torch.randn_like(y)
    osp = y.shape[1]
    flops = 2.0 * n * osp ** 3 * cout * cin * 27
    tf = timeit(lambda: Fn.conv3d_raw(x, w, b, k, s, p, want_stats=True))
    td = timeit(lambda: Fn.conv3d_dgrad_raw(dy, w, x.shape, k, s, p))
    tw = timeit(lambda: Fn.conv3d_wgrad_raw(x, dy, w.shape, k, s, p))
    rows.append((name, cin, cout, sp, st, flops / 1e9, tf, td, tw))
    print(f"{name:7s} {cin:4d}->{cout:4d} in{sp:4d}^3 s{st}  {flops/1e9:7.1f} GF | fwd {tf:8.3f} ms {flops/tf/1e9:7.1f} TF/s | "
          f"dgrad {td:8.3f} ms {flops/td/1e9:7.1f} | wgrad {tw:8.3f} ms {flops/tw/1e9:7.1f}", flush=True)
tot = [sum(r[i] for r in rows) for i in (5, 6, 7, 8)]
print(f"total {tot[0]:.1f} GF: fwd {tot[1]:.2f} ms, dgrad {tot[2]:.2f} ms, wgrad {tot[3]:.2f} ms")

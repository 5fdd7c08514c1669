"""Development tool: device time of the k2s2 transposed convs of the C1 decoder (fwd into a dense buffer / into the concat
buffer, dgrad, wgrad, bias grad)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import functional as Fn

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda._sleep(1_500_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


only = sys.argv[1:]
for name, cin, cout, sp in [("dec4", 64, 32, 64), ("dec3", 128, 64, 32), ("dec2", 256, 128, 16), ("dec1", 320, 256, 8), ("dec0", 320, 320, 4)]:
    if only and name not in only:
        continue
    n = 2
    x = torch.randn((n, sp, sp, sp, cin), device="cuda").to(torch.bfloat16).requires_grad_(True)
    w = (torch.randn((cin, cout, 2, 2, 2), device="cuda") / cin ** 0.5).requires_grad_(True)
    b = torch.randn((cout,), device="cuda").requires_grad_(True)
    k = s = (2, 2, 2)
    with torch.no_grad():
        t_dense = timeit(lambda: Fn.conv_transpose(x, w, b, k, s))
    y = Fn.conv_transpose(x, w, b, k, s)
    dy = torch.randn_like(y)
    t_bwd = timeit(lambda: torch.autograd.grad(y, (x, w, b), dy, retain_graph=True))
    out_mb = y.numel() * 2 / 1e6
    in_mb = x.numel() * 2 / 1e6
    print(f"{name} {cin}->{cout} in{sp}^3: fwd dense {t_dense*1e3:7.1f} us ({(in_mb+out_mb)/t_dense/1e3:5.2f} TB/s of {in_mb+out_mb:.0f} MB)  "
          f"bwd (dgrad+wgrad+bias) {t_bwd*1e3:7.1f} us", flush=True)

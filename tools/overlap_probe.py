"""Can an HBM-bound InstanceNorm pass run NEXT to a tensor-core conv CTA?  Times a marching conv (fwd 64->32 @128^3, one sample)
and the InstanceNorm passes of another sample alone and concurrently on two streams.  Run once per REHR_SMEM_RESERVE_KB value
(the library reads it once per process): with 0 the conv CTAs take all 227 KB and nothing else becomes resident."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import functional as F, _lib as L
from rehrseg_b200._lib import lib, ptr, rt, check, stream_ptr

dev = torch.device("cuda")
torch.manual_seed(0)
S = int(os.environ.get("PROBE_SIZE", "128"))
cin, cout = 64, 32
k3, s1, p1 = (3, 3, 3), (1, 1, 1), (1, 1, 1)
x = (torch.randn(1, S, S, S, cin, device=dev) * 0.5).to(torch.float16).view(torch.bfloat16)
xb = torch.randn(1, S, S, S, cin, device=dev).to(torch.bfloat16)
w = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.05
y2 = torch.randn(1, S, S, S, cout, device=dev).to(torch.float16).view(torch.bfloat16)
da = torch.randn(1, S, S, S, cout, device=dev).to(torch.bfloat16)
mean = torch.zeros(1, cout, device=dev); rstd = torch.ones(1, cout, device=dev)
g = torch.ones(cout, device=dev); b = torch.zeros(cout, device=dev)
a = torch.empty_like(y2); a2 = torch.empty_like(y2); dy = torch.empty_like(y2)
sums = torch.zeros(1, cout, 2, device=dev)
yt = rt(y2, True)
tiles = lib().rehr_instnorm_stats_tiles(C.byref(yt))
partial = torch.empty(1, tiles, cout, 2, device=dev)
yout = torch.empty(1, S, S, S, cout, dtype=torch.bfloat16, device=dev)
dxout = None


def conv():
    F.conv3d_raw(x, w, None, k3, s1, p1, want_stats=True, x_h=True, y_h=True, out=yout)


def wgrad():
    F.conv3d_wgrad_raw(xb, da, w.shape, k3, s1, p1)


def in_fwd():
    at, a2t = rt(a, True), rt(a2, False)
    check(lib().rehr_instnorm_lrelu_apply(C.byref(yt), ptr(mean), ptr(rstd), ptr(g), ptr(b), 0.01, C.byref(at), C.byref(a2t), stream_ptr()))


def in_bwd():
    dat, dyt = rt(da), rt(dy)
    check(lib().rehr_instnorm_lrelu_bwd_reduce(C.byref(yt), C.byref(dat), None, ptr(mean), ptr(rstd), ptr(g), ptr(b), 0.01, ptr(partial), stream_ptr()))
    check(lib().rehr_instnorm_lrelu_bwd_apply(C.byref(yt), C.byref(dat), None, ptr(mean), ptr(rstd), ptr(g), ptr(b), 0.01, ptr(sums), C.byref(dyt), stream_ptr()))


sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
REP = 8


def run(fa, fb):
    """REP x fa on stream A next to REP x fb on stream B (either may be None); returns ms per repetition."""
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    sA.wait_event(e0); sB.wait_event(e0)
    for _ in range(REP):
        if fa:
            with torch.cuda.stream(sA):
                fa()
        if fb:
            with torch.cuda.stream(sB):
                fb()
    main.wait_stream(sA); main.wait_stream(sB)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REP


for name, fa in (("conv_fwd 64->32", conv), ("wgrad 64->32", wgrad)):
    for nb, fb in (("in_fwd(apply+twin)", in_fwd), ("in_bwd(reduce+apply)", in_bwd)):
        for _ in range(2):
            ta, tb, tab = run(fa, None), run(None, fb), run(fa, fb)
        print(f"reserve={os.environ.get('REHR_SMEM_RESERVE_KB', 'default')} S={S}  {name}: {ta*1e3:7.1f} us | {nb}: {tb*1e3:7.1f} us | "
              f"concurrent: {tab*1e3:7.1f} us  (sum {1e3*(ta+tb):7.1f}, hidden {100*(ta+tb-tab)/tb:5.1f} % of the IN pass)")

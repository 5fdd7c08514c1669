"""CUDA-event timings of the HBM-bound helper kernels of configs 3 and 5 at full size (sw_accumulate / sw_finalize on a 256^3 volume,
blur1d / rot90 / mean_stack / fba_combine on 512x512x160): algorithmic bytes / time against the measured HBM peak.  Every call works on
buffers that together exceed the 126 MB L2 (or a 256 MB scratch write evicts it first), 3 warm-up + 10 timed calls, median.
Usage: python tools/hbm_kernels_bench.py [out.json]"""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import sliding_window as sw, volume_ops as vo
from rehrseg_b200._lib import lib, ptr, stream_ptr, check

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10, warm=3):
    ts = []
    for i in range(warm + reps):
        scratch.fill_(i & 1)                      # evict L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    return statistics.median(ts)


peaks = {}
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        peaks = json.load(f)
except OSError:
    pass
hbm_peak = float(peaks.get("hbm_gbs", peaks.get("hbm_gb_s", 6460.0)))
out = {"hbm_peak_gbs": hbm_peak, "kernels": {}}


def row(name, ms, nbytes, note):
    gbs = nbytes / ms / 1e6
    out["kernels"][name] = {"ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(gbs, 1),
                            "frac_hbm_peak": round(gbs / hbm_peak, 3), "note": note}
    print(f"{name:22s} {ms * 1e3:8.1f} us {nbytes / 1e6:8.1f} MB {gbs:8.1f} GB/s {gbs / hbm_peak:6.1%}  {note}")


# --- sliding-window blend: 2 classes, 128^3 tile into a 256^3 volume (fp16 accumulators, fp16 prediction, Gaussian)
logits = torch.zeros((2, 256, 256, 256), dtype=torch.half, device=dev)
npred = torch.zeros((256, 256, 256), dtype=torch.half, device=dev)
pred = torch.randn((2, 128, 128, 128), device=dev, generator=g).half()
gauss = sw.importance_map([128, 128, 128], 1. / 8, 10, device=dev)
tv = 128 ** 3
ms = timed(lambda: sw.sw_accumulate(logits, npred, pred, gauss, (64, 64, 64)))
row("sw_accumulate", ms, tv * 2 * (2 + 1 + 2 * 2 + 2), "read pred(2) + gauss, read+write logits(2) and npred, 2 B each")
npred.fill_(3.0)
vv = 256 ** 3
flag = torch.zeros((1,), dtype=torch.int32, device=dev)
ms = timed(lambda: check(lib().rehr_sw_finalize(ptr(logits), ptr(npred), 2, vv, ptr(flag), stream_ptr()), "sw_finalize"))
row("sw_finalize", ms, vv * 2 * (1 + 2 * 2), "read npred, read+write logits(2)")
del logits, npred, pred

# --- C5 volume helpers on 512x512x160 fp32
hr = torch.rand((160, 1, 512, 512), device=dev, generator=g)
taps = torch.exp(-0.5 * ((torch.arange(9.) - 4) / (3.873 / 2.355)) ** 2)
k9 = (taps / taps.sum()).reshape(1, 1, 9, 1).to(dev)
ms = timed(lambda: vo.blur_along_x(hr, k9))
row("blur1d", ms, hr.numel() * 8, "read + write fp32 (includes the output allocation)")
vols = [torch.rand((512, 512, 160), device=dev, generator=g) for _ in range(4)]
ms = timed(lambda: vo.rotate_vol_2d(vols[0], 90))
row("rot90", ms, vols[0].numel() * 8, "read + write fp32")
ms = timed(lambda: vo.mean_fuse(vols))
row("mean_stack", ms, vols[0].numel() * 4 * 5, "4 reads + 1 write fp32")
# the spectral combine alone, on 4 half-spectra of the same volume (cuFFT is a library call, timed separately in bench_extras)
spec = [torch.fft.rfftn(v) for v in vols]
outc = torch.empty_like(spec[0])
views = [torch.view_as_real(s) for s in spec]
arr = vo._ptr_array(views)
n = spec[0].numel()
for p, tag in ((-1.0, "fba_combine_inf"), (2.0, "fba_combine_p2")):
    ms = timed(lambda: check(lib().rehr_fba_combine(arr, 4, p, ptr(torch.view_as_real(outc)), n, stream_ptr()), "fba"))
    row(tag, ms, n * 8 * 5, "4 complex64 reads + 1 write")

# --- depth-only linear up-sampling of the SR head at the C4 shape: [2, 16, 256, 256, 32] bf16 -> depth 64
del vols, spec, outc, views, hr
from rehrseg_b200 import functional as Fn
feats = torch.randn((2, 16, 256, 256, 32), device=dev, generator=g).to(torch.bfloat16)
with torch.no_grad():
    ms = timed(lambda: Fn.upsample_linear_d(feats, 64))
row("upsample_d", ms, feats.numel() * 2 * 5, "1 read + 4x write bf16 (includes the output allocation)")

if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        json.dump(out, f, indent=1)

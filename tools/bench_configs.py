"""Development tool: device timings of the other BASELINE configs (C2 FLAVR fwd+bwd, C3 sliding window, C5 pipeline pieces).
Not the bench line (bench.py measures C1); results are quoted in DESIGN.md."""
import sys, os, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehrseg_b200 import seg_model as sm, flavr, sliding_window as sw, volume_ops as vo, functional as Fn

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="c2,c3,c5")
a = ap.parse_args()
what = a.what.split(",")


def timeit(fn, iters=5, warm=4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if "c2" in what:
    for unc, tf in ((False, 1.650), (True, 1.834)):
        torch.manual_seed(0)
        m = flavr.UNet_3D_3D(2, "unet_18", 4, 4, False, "concat", "transpose", use_uncertainty=unc).cuda()
        for B in (1, 8):
            x = torch.rand((B, 2, 4, 256, 256), device="cuda")

            def step():
                for p in m.parameters():
                    p.grad = None
                out = m(x.clone())
                out = out[0] if unc else out
                out.float().mean().backward()
            ms = timeit(step)
            print(f"C2 FLAVR {'UASR' if unc else 'plain'} fwd+bwd B={B} 256x256: {ms:.2f} ms/step, {B / ms * 1e3:.1f} samples/s, "
                  f"{B * tf / ms * 1e3:.0f} TFLOP/s", flush=True)
        with torch.no_grad():
            x = torch.rand((8, 2, 4, 256, 256), device="cuda")
            ms = timeit(lambda: m(x.clone()))
            print(f"   fwd only B=8: {ms:.2f} ms, {8 * (0.550 if not unc else 0.611) / ms * 1e3:.0f} TFLOP/s", flush=True)
        del m

if "c3" in what:
    torch.manual_seed(0)
    m = sm.plainconv_3d_fullres().cuda().eval()
    x = torch.randn((1, 1, 128, 128, 128), device="cuda")
    with torch.no_grad():
        ms = timeit(lambda: m(x))
    print(f"C3 one SegModel tile forward (U-Net + x4 SR head): {ms:.2f} ms", flush=True)
    unet = sm.plainconv_unet_3d_fullres().cuda().eval()
    with torch.no_grad():
        ms_u = timeit(lambda: unet(x))
    print(f"C3 one PlainConvUNet tile forward (no SR head): {ms_u:.2f} ms", flush=True)
    vol = torch.randn((1, 256, 256, 256), device="cuda")
    slicers = sw._internal_get_sliding_window_slicers(vol.shape[1:], patch_size=[128, 128, 128])
    with torch.no_grad():
        t0 = time.perf_counter()
        out = sw._internal_predict_sliding_window_return_logits(vol, slicers, m, True, 0, 1, [128, 128, 128], use_gaussian=True,
                                                                deep_supervision=False)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
    print(f"C3 256^3 volume, {len(slicers)} tiles x 8 mirror passes = {len(slicers) * 8} forwards: {t:.2f} s/volume "
          f"({206.2 / t:.0f} TFLOP/s; the SR head is skipped because only output 0 is read)", flush=True)

if "c5" in what:
    g = torch.Generator(device="cuda").manual_seed(5)
    hr = torch.rand((160, 1, 512, 512), device="cuda", generator=g)
    taps = torch.exp(-0.5 * ((torch.arange(9.) - 4) / (3.873 / 2.355)) ** 2)
    k = (taps / taps.sum()).reshape(1, 1, 9, 1).cuda()
    ms = timeit(lambda: vo.blur_along_x(hr, k))
    print(f"C5 blur 160x512x512: {ms:.3f} ms, {hr.numel() * 8 / ms / 1e6:.0f} GB/s", flush=True)
    vols = [torch.rand((512, 512, 160), device="cuda", generator=g) for _ in range(4)]
    for p in ("infinity", 2.0):
        ms = timeit(lambda: vo.fba(vols, p))
        print(f"C5 fba(p={p}) 4 x 512x512x160 incl. cuFFT: {ms:.2f} ms", flush=True)
    ms = timeit(lambda: vo.mean_fuse(vols))
    print(f"C5 mean fusion: {ms:.3f} ms, {5 * vols[0].numel() * 4 / ms / 1e6:.0f} GB/s", flush=True)
    ms = timeit(lambda: vo.rotate_vol_2d(vols[0], 90))
    print(f"C5 rot90 512x512x160: {ms:.3f} ms, {2 * vols[0].numel() * 4 / ms / 1e6:.0f} GB/s", flush=True)
    torch.manual_seed(0)
    m = flavr.UNet_3D_3D(2, "unet_18", 4, 4, False, "concat", "transpose", use_uncertainty=False).cuda().eval()
    lr = torch.rand((41, 2, 512, 512), device="cuda", generator=g)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = flavr.apply_to_vol_flavr(m, lr, max_batch=4)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    print(f"C5 one orientation: 40 windows of [2,4,512,512] -> {tuple(out.shape)}: {t:.2f} s ({40 * 2.2 / t:.0f} TFLOP/s)", flush=True)

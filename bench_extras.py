"""The other BASELINE.json configs, timed in the same run as bench.py's C1 line so that every config carries a driver-visible
number (BASELINE.json `metric` also names "volumes/sec SW infer"):

  C2  FLAVR UNet3D self-SR fwd+bwd on [8,2,4,256,256] (plain head)                      -> flavr_samples_per_s
  C3  sliding-window Gaussian-blended inference over a 256^3 volume, 27 tiles x 8 mirror
      passes, tiles x mirrors sharded over the ranks                                     -> sw_volumes_per_s
  C4  joint SR+seg step (anisotropic SegModel student, UASR FLAVR teacher sweep,
      Distiller, SGD), batch 2 per GPU, h2d copies inside                                -> joint_ms_per_step
  C5  blur degradation + 4-orientation SR sweep + FBA / mean fusion on 512x512x160
      (one orientation sweep is timed and multiplied by the exact orientation count)    -> c5_pipeline_s
  cuDNN  torch eager + bf16 autocast, channels_last_3d, C1 shape on the same GPU        -> eager_gpu_ms_per_step

Every entry: device time from CUDA events (wall clock where host work is part of the path, said so), algorithmic work from
SURVEY.md section 8(d), fraction of the MEASURED bf16 / HBM peak.  Failures never break the main line: the entry carries "error".
"""
from __future__ import annotations

import os
import time

import torch


def _events(fn, iters, warm):
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


def c2_flavr(dev, peaks, steps=8, warm=3, uasr=False) -> dict:
    from rehrseg_b200 import flavr
    torch.manual_seed(0)
    m = flavr.UNet_3D_3D(2, "unet_18", 4, 4, False, "concat", "transpose", use_uncertainty=uasr).to(dev)
    B = 8
    x = torch.rand((B, 2, 4, 256, 256), device=dev)

    # the step is ~700 launches of 5-300 us: issued from Python it is bound by the host (26.5 ms in a quiet process, 35 ms here next
    # to bench.py's sampler / CPU-baseline threads), so it is replayed from one CUDA graph like the C1 step
    from rehrseg_b200.graphs import GraphedTrainStep

    class _Fresh(torch.nn.Module):          # UNet_3D_3D.forward subtracts the mean from images[:, 0:1] IN PLACE (FLAVR_arch.py:172)
        def __init__(self, net):
            super().__init__()
            self.net = net

        def forward(self, images):
            return self.net(images.clone())

    loss_fn = (lambda out: out[0].float().mean() + out[1].float().mean()) if uasr else (lambda out: out.float().mean())
    gs = GraphedTrainStep(_Fresh(m), loss_fn, (x,))
    ms = _events(gs.replay, steps, warm)
    gs.close()
    tf = B * (1.834 if uasr else 1.650) / ms * 1e3     # 1.650 / 1.834 TFLOP fwd+bwd per 256^2 sample (plain / UASR head)
    return {"config": f"C2 FLAVR UNet3D self-SR fwd+bwd, [8,2,4,256,256], {'UASR' if uasr else 'plain'} head, bf16, one CUDA graph per step", "flavr_samples_per_s": round(B / ms * 1e3, 2),
            "ms_per_step": round(ms, 3), "tflops": round(tf, 1), "frac_bf16_peak_burst": round(tf / float(peaks["bf16_tflops"]), 4)}


def c3_sliding_window(dev, peaks, world, iters=2) -> dict:
    import torch.distributed as dist
    from rehrseg_b200 import seg_model as sm, sliding_window as sw
    torch.manual_seed(0)
    model = sm.plainconv_3d_fullres().to(dev).eval()
    vol = torch.randn((1, 256, 256, 256), generator=torch.Generator().manual_seed(3)).to(dev)
    patch = [128, 128, 128]
    slicers = sw._internal_get_sliding_window_slicers(vol.shape[1:], patch_size=patch)

    def run():
        with torch.no_grad():
            return sw.predict_sliding_window_sharded(vol, slicers, model, out_idx=0, patch_size=patch, use_gaussian=True,
                                                     deep_supervision=False)

    run()                       # graph capture of the tile forward, importance map
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = run()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters / 1e3], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s = float(t)
    tf = 206.2 / s              # 216 U-Net forwards = 206.2 TFLOP (the SR head is not computed: only output 0 is read)
    return {"config": f"C3 sliding window, 256^3 volume, {len(slicers)} tiles x 8 mirror passes, Gaussian fp16 blend, (tile, mirror) units "
                      f"sharded over {world} GPU(s)", "sw_volumes_per_s": round(1.0 / s, 3), "s_per_volume": round(s, 4),
            "tflops_all_gpus": round(tf, 1), "frac_bf16_peak_burst": round(tf / world / float(peaks["bf16_tflops"]), 4),
            "finite": bool(torch.isfinite(out.float()).all())}


def c4_joint(dev, peaks, steps=6, warm=3, teacher_keys=None) -> dict:
    import itertools
    from rehrseg_b200 import flavr, loss_ops, seg_model as sm, train_step as ts
    aniso = dict(input_channels=1, n_stages=6, features_per_stage=[32, 64, 128, 256, 320, 320], conv_op=torch.nn.Conv3d,
                 kernel_sizes=[[1, 3, 3], [1, 3, 3]] + [[3, 3, 3]] * 4,
                 strides=[[1, 1, 1], [1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2], [1, 2, 2]], n_conv_per_stage=[2] * 6, num_classes=2,
                 upscale=4, n_conv_per_stage_decoder=[2] * 5, conv_bias=True, norm_op=torch.nn.InstanceNorm3d,
                 norm_op_kwargs={"eps": 1e-5, "affine": True}, dropout_op=None, dropout_op_kwargs=None, nonlin=torch.nn.LeakyReLU,
                 nonlin_kwargs={"inplace": True}, deep_supervision=False)
    torch.manual_seed(1234)
    student = sm.SegModel(**aniso).to(dev)
    teacher = flavr.UNet_3D_3D(2, "unet_18", 4, 4, False, "concat", "transpose", use_uncertainty=True).to(dev).eval()
    distiller = loss_ops.FusedDistiller(64, 64, 0.0, 1.0, 1.0).to(dev)
    opt = torch.optim.SGD(itertools.chain(student.parameters(), distiller.parameters()), lr=1e-3, momentum=0.99, nesterov=True,
                          weight_decay=3e-5)
    lr_obj, hr_obj = loss_ops.build_fused_loss(False, 0), loss_ops.build_fused_loss(False, 1)
    g = torch.Generator().manual_seed(4)
    B, D, HW = 2, 16, 256
    host = (torch.randn((B, 1, D, HW, HW), generator=g).pin_memory(),
            (torch.rand((B, 1, D, HW, HW), generator=g) > 0.8).float().pin_memory(),
            (torch.rand((B, 1, 4 * D, HW, HW), generator=g) > 0.8).float().pin_memory(),
            (torch.rand((B, 1, D, HW, HW), generator=g) * 0.99 + 0.01).pin_memory())

    # the whole iteration (teacher sweep, student forward, losses, backward) replayed from one CUDA graph; per step: batch copied
    # from pinned host memory into the static inputs, replay, SGD step, loss read back like the training script printing it
    gs = ts.GraphedJointStep(student, tuple(t.to(dev) for t in host), lr_obj, hr_obj, teacher, distiller, teacher_keys=teacher_keys)

    def step():
        out = gs(host)
        opt.step()
        float(out["loss"])

    ms = _events(step, steps, warm)
    gs.close()
    if teacher_keys is not None:
        # Opt-in variant, NOT the reference's workload: the loop reads only features_sr[1] (train_all.py:550), so the teacher stops
        # after layer1 (stem 2.47 + 4 x 14.50 GFLOP per sample x 30 samples = 1.81 TFLOP instead of 10.99).  Same losses and gradients.
        tf = (4.50 + 1.81) / ms * 1e3
        return {"config": "C4 variant (opt-in `teacher_keys=(1,)`, not the reference's workload): the teacher sweep stops after the one "
                          "feature map the stage-2 loop reads; identical losses and gradients; everything else as c4_joint",
                "joint_ms_per_step_needed_features_only": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 2),
                "tflops_of_the_work_done": round(tf, 1), "frac_bf16_peak_burst": round(tf / float(peaks["bf16_tflops"]), 4)}
    tf = 15.5 / ms * 1e3        # student fwd+bwd 4.50 + teacher sweep 10.99 TFLOP per GPU-step
    return {"config": "C4 joint SR+seg step: anisotropic SegModel student [2,1,16,256,256] x4 SR head, UASR FLAVR teacher sweep (15 windows), "
                      "uncertainty-weighted CE + CE/Dice + Distiller(64,64,0,1,1), SGD, batch from pinned host memory, one CUDA graph per step",
            "joint_ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 2), "tflops": round(tf, 1),
            "frac_bf16_peak_burst": round(tf / float(peaks["bf16_tflops"]), 4)}


def c5_pipeline(dev, peaks) -> dict:
    from rehrseg_b200 import flavr, volume_ops as vo
    g = torch.Generator(device=dev).manual_seed(5)
    hr = torch.rand((160, 1, 512, 512), device=dev, generator=g)
    taps = torch.exp(-0.5 * ((torch.arange(9.) - 4) / (3.873 / 2.355)) ** 2)
    k = (taps / taps.sum()).reshape(1, 1, 9, 1).to(dev)
    t_blur = _events(lambda: vo.blur_along_x(hr, k), 5, 2)
    vols = [torch.rand((512, 512, 160), device=dev, generator=g) for _ in range(4)]
    t_fba = _events(lambda: vo.fba(vols, "infinity"), 3, 2)
    t_mean = _events(lambda: vo.mean_fuse(vols), 5, 2)
    t_rot = _events(lambda: vo.rotate_vol_2d(vols[0], 90), 5, 2)
    torch.manual_seed(0)
    m = flavr.UNet_3D_3D(2, "unet_18", 4, 4, False, "concat", "transpose", use_uncertainty=False).to(dev).eval()
    lr = torch.rand((41, 2, 512, 512), device=dev, generator=g)
    flavr.apply_to_vol_flavr(m, lr[:9], max_batch=4)   # warm-up (weight packs, allocator, capture of the 4-window forward)
    # best of 3 single sweeps: the sweep interleaves host work (window stacking, 10 graph launches) with the GPU's, and one stalled
    # host thread next to bench.py's sampler / CPU-baseline threads tripled a single measurement (0.16 -> 0.40 s) in one run
    t_sweep = min(_events(lambda: flavr.apply_to_vol_flavr(m, lr, max_batch=4), 1, 0) for _ in range(3))
    orient = 4
    total = (t_blur + orient * (t_sweep + 2 * t_rot) + t_fba + t_mean) / 1e3
    nbytes = hr.numel() * 8
    return {"config": "C5 self-SR pipeline on 512x512x160: blur (9 taps) + 4 orientations x (rot90, 40-window FLAVR sweep, rot-90) + "
                      "fba(p=inf) and mean fusion; best of 3 timed sweeps, multiplied by the orientation count", "c5_pipeline_s": round(total, 4),
            "blur_ms": round(t_blur, 4), "blur_gbs": round(nbytes / t_blur / 1e6, 1),
            "blur_frac_hbm_peak": round(nbytes / t_blur / 1e6 / float(peaks["hbm_gbs"]), 4),
            "fba_ms": round(t_fba, 3), "mean_fuse_ms": round(t_mean, 4), "mean_fuse_gbs": round(5 * vols[0].numel() * 4 / t_mean / 1e6, 1),
            "rot90_ms": round(t_rot, 4), "rot90_gbs": round(2 * vols[0].numel() * 4 / t_rot / 1e6, 1),
            "sweep_s_per_orientation": round(t_sweep / 1e3, 4), "sweep_tflops": round(40 * 2.2 / t_sweep * 1e3, 1),
            "frac_bf16_peak_burst": round(orient * 40 * 2.2 / total / float(peaks["bf16_tflops"]), 4)}


def loaders(dev) -> dict:
    """The two data pipelines in front of the networks, on the GPU: SR-stage sample synthesis
    (degrade.SRTrainSampler, 96x96 pairs from a resident 512x512x160 volume) and the stage-2 spatial augmentation
    (augment.spatial_transform_dummy_2d, one batch of 2 patches = 112 slices of 256^2, rotation + scaling forced on).  Wall clock
    (their cost is host launches), next to the oracle's CPU statements in one process on a bounded sample."""
    import random
    import numpy as np
    from rehrseg_b200 import augment, degrade
    from oracle import augment as oa, degrade as od
    rng = np.random.RandomState(0)
    img = rng.rand(512, 512, 160, 1).astype(np.float32)
    lab = (rng.rand(512, 512, 160, 1) > 0.7).astype(np.uint8)
    taps = np.exp(-0.5 * ((np.arange(9.0) - 4) / (3.873 / 2.355)) ** 2)
    kernel = torch.tensor(taps / taps.sum(), dtype=torch.float32).reshape(1, 1, 9, 1)
    ds = degrade.SRTrainSampler([96, 96, 1], 4.0, blur=True, random_flip=True, blur_kernel=kernel.to(dev))
    t0 = time.perf_counter()
    ds.add_subject(img, lab)
    torch.cuda.synchronize()
    t_pre = time.perf_counter() - t0
    random.seed(0)
    B = 32
    for _ in range(3):
        ds.batch([0] * B)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        ds.batch([0] * B)
    torch.cuda.synchronize()
    sr_rate = 10 * B / (time.perf_counter() - t0)
    # CPU statements of the reference's __getitem__ (crop of an already filtered volume: a slab keeps the sample bounded)
    fx = np.ascontiguousarray(img.transpose(2, 3, 0, 1))
    fy = np.ascontiguousarray(img.transpose(2, 3, 1, 0))
    random.seed(0)
    t0 = time.perf_counter()
    for _ in range(32):
        od.train_sample(img, lab, fx, fy, [96, 96, 1], 4.0, True, True)
    sr_cpu = 32 / (time.perf_counter() - t0)

    class Forced:          # every sample rotates and scales (the reference: 20 % each)
        def __init__(self, seed):
            self.r = np.random.RandomState(seed)

        def uniform(self, *a):
            return self.r.uniform(*a) if a else 0.0

        def random(self):
            return self.r.random_sample()

    b, z, X, Y = 2, 16, 256, 256
    dd = {"data": rng.randn(b, 1, z, X, Y).astype(np.float32), "seg": (rng.rand(b, 1, z, X, Y) > 0.5).astype(np.float32),
          "seg_sr": (rng.rand(b, 1, 4 * z, X, Y) > 0.5).astype(np.float32), "uncertainty": rng.rand(b, 1, z, X, Y).astype(np.float32)}
    ddev = {k: torch.from_numpy(v).to(dev) for k, v in dd.items()}
    for _ in range(3):
        augment.spatial_transform_dummy_2d(ddev, (z, X, Y), rng=Forced(1), seg_labels=(0.0, 1.0))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        augment.spatial_transform_dummy_2d(ddev, (z, X, Y), rng=Forced(1), seg_labels=(0.0, 1.0))
    torch.cuda.synchronize()
    aug_ms = (time.perf_counter() - t0) / 10 * 1e3
    small = {k: v[:1, :, :(4 if k != "seg_sr" else 16)] for k, v in dd.items()}      # bounded CPU sample: 1 patch, 4 LR slices
    t0 = time.perf_counter()
    oa.spatial_transform_dummy_2d({k: v.copy() for k, v in small.items()}, (4, X, Y), rng=Forced(1))
    aug_cpu_s = (time.perf_counter() - t0) * (b * z / 4)                                # scaled to the full batch by the slice count
    return {"config": "GPU data pipelines: SR-stage sample synthesis (96x96 pairs, resident 512x512x160 volume) and stage-2 spatial "
                      "augmentation (2 patches [1,16,256,256] + LR/HR labels + uncertainty, rotation and scaling forced)",
            "sr_sampler_samples_per_s": round(sr_rate, 0), "sr_sampler_cpu_port_samples_per_s": round(sr_cpu, 0),
            "sr_volume_prefilter_s": round(t_pre, 3), "stage2_augment_ms_per_batch": round(aug_ms, 3),
            "stage2_augment_patches_per_s": round(b / aug_ms * 1e3, 0),
            "stage2_augment_cpu_port_s_per_batch": round(aug_cpu_s, 3), "cpu_port": "oracle statements, 1 process; augmentation: 1/8 of "
            "the batch's slices timed and scaled"}


def eager_gpu(dev, make_oracle_unet, batch, patch, steps=3, warm=2) -> dict:
    """The reference's own GPU path on the same box: torch eager + cuDNN, bf16 autocast, channels_last_3d (informational)."""
    model = make_oracle_unet().to(dev).to(memory_format=torch.channels_last_3d)
    x = torch.randn((batch, 1, *patch), device=dev)
    gt = torch.randn((batch, 2, *patch), device=dev)

    def step():
        for p in model.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(x)
        (torch.dot(out.float().reshape(-1), gt.reshape(-1)) / out.numel()).backward()

    ms = _events(step, steps, warm)
    return {"config": "C1 shape through torch eager + cuDNN, bf16 autocast, channels_last_3d, same GPU", "eager_gpu_ms_per_step": round(ms, 3),
            "patches_per_s": round(batch / ms * 1e3, 2)}


def run_all(dev, peaks, rank, world, make_oracle_unet, batch, patch) -> dict:
    """rank 0 returns {name: entry}; C3 runs on every rank (sharded), the single-GPU configs only when world == 1."""
    out = {}

    def guarded(name, fn):
        t0 = time.perf_counter()
        try:
            out[name] = fn()
        except Exception as e:  # noqa: BLE001 -- an extra must never take the main bench line down
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        out[name]["wall_s"] = round(time.perf_counter() - t0, 1)
        torch.cuda.empty_cache()

    guarded("c3_sliding_window", lambda: c3_sliding_window(dev, peaks, world))
    if world == 1:
        guarded("c2_flavr", lambda: c2_flavr(dev, peaks))
        guarded("c2_flavr_uasr", lambda: c2_flavr(dev, peaks, uasr=True))
        guarded("c4_joint", lambda: c4_joint(dev, peaks))
        guarded("c5_pipeline", lambda: c5_pipeline(dev, peaks))
        guarded("loaders", lambda: loaders(dev))
        if os.environ.get("REHR_BENCH_EAGER", "1") != "0":
            guarded("eager_gpu", lambda: eager_gpu(dev, make_oracle_unet, batch, patch))
        guarded("c4_joint_needed_features", lambda: c4_joint(dev, peaks, teacher_keys=(1,)))   # opt-in variant, last: see c4_joint
    return out if rank == 0 else {}

"""GPU-vs-oracle parity helpers -- TEST INFRASTRUCTURE (used by tests/, smoke() and bench.py's checker only)."""
from __future__ import annotations

import torch

from . import seg_model as ref_seg


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def segmodel_parity(patch=(64, 64, 64), batch=1, plan="3d_fullres", device="cuda:0", backward=True, seed=0,
                    threads: int | None = None, autocast_baseline: bool = False) -> dict:
    """Run the oracle SegModel on CPU (fp32) and the B200 engine on `device` with identical weights and input
    (SURVEY.md section 8(d), config 1 protocol: x ~ N(0,1) seed 0; cotangent g ~ N(0,1) seed 1,
    loss = <logits, g>/numel + <hr_logits, g2>/numel)."""
    from rehrseg_b200 import seg_model as sm

    if threads:
        torch.set_num_threads(threads)
    ref = ref_seg.build(plan)
    mine = sm.SegModel(**ref_seg.plan_kwargs(plan))
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(device)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((batch, 1, *patch), generator=g)
    out_r, up_r = ref(x)
    out_m, up_m = mine(x.to(device))
    res = {
        "rel_l2_logits": rel_l2(out_m, out_r),
        "rel_l2_hr_logits": rel_l2(up_m, up_r),
        "argmax_agreement": float((out_m.argmax(1).cpu() == out_r.argmax(1)).double().mean()),
    }
    # Argmax agreement restricted to voxels whose fp32 decision margin exceeds 4x the RMS logit error: with random-init
    # weights the two class logits are nearly tied almost everywhere, so raw agreement measures the tie density, not
    # the kernels (torch's own bf16 autocast path scores the same 99.6-99.7 % -- see `autocast_*` below and DESIGN.md).
    om, orr = out_m.float().cpu(), out_r.float()
    margin = (orr[:, 0] - orr[:, 1]).detach().abs()
    tau = 4.0 * float((om - orr).detach().pow(2).mean().sqrt())
    sel = margin > tau
    res["argmax_agreement_clear_margin"] = float((om.argmax(1) == orr.argmax(1))[sel].double().mean()) if bool(sel.any()) else 1.0
    res["clear_margin_fraction"] = float(sel.double().mean())
    if autocast_baseline:
        # the reference's own GPU bf16 path (torch.autocast -> cuDNN) on the same weights / input, same fp32 yardstick
        ref_gpu = ref_seg.build(plan)
        ref_gpu.load_state_dict(ref.state_dict())
        ref_gpu = ref_gpu.to(device)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            out_a, up_a = ref_gpu(x.to(device))
        res["autocast_rel_l2_logits"] = rel_l2(out_a.float(), out_r)
        res["autocast_argmax_agreement"] = float((out_a.float().argmax(1).cpu() == out_r.argmax(1)).double().mean())
        del ref_gpu
    if backward:
        g1 = torch.randn(out_r.shape, generator=torch.Generator().manual_seed(seed + 1))
        g2 = torch.randn(up_r.shape, generator=torch.Generator().manual_seed(seed + 2))
        loss_r = (out_r * g1).sum() / out_r.numel() + (up_r * g2).sum() / up_r.numel()
        loss_r.backward()
        loss_m = (out_m * g1.to(device)).sum() / out_m.numel() + (up_m * g2.to(device)).sum() / up_m.numel()
        loss_m.backward()
        worst, worst_name = 0.0, ""
        num = den = 0.0
        pr = dict(ref.named_parameters())
        for name, p in mine.named_parameters():
            if p.grad is None or pr[name].grad is None:
                continue
            a, b = p.grad.detach().double().cpu(), pr[name].grad.detach().double()
            if name.endswith("conv.bias") and ".convs." in name:
                continue  # bias before InstanceNorm: exact gradient is 0, the reference's value is rounding noise
            num += float((a - b).pow(2).sum())
            den += float(b.pow(2).sum())
            r = float((a - b).norm() / (b.norm() + 1e-30))
            if r > worst:
                worst, worst_name = r, name
        res["rel_l2_grads_global"] = (num / max(den, 1e-300)) ** 0.5
        res["rel_l2_grads_worst"] = worst
        res["worst_grad"] = worst_name
    return res

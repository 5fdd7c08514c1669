"""Batched weight re-pack (csrc/pack_batch.cu, rehr_pack_batch_begin / _launch): every cached 16-bit operand copy of a model,
re-packed in ONE launch from updated parameters, must be bit-identical to the copies the individual pack kernels produce."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _one_step(model, x):
    for p in model.parameters():
        p.grad = None
    out = model(x)
    out = out[0] if isinstance(out, (tuple, list)) else out
    out.float().mean().backward()


@pytest.mark.parametrize("which", ["segmodel_64", "flavr"])
def test_batched_repack_is_bit_identical(which, monkeypatch):
    from rehrseg_b200 import functional as Fn, seg_model as sm, flavr
    torch.manual_seed(0)
    if which == "segmodel_64":
        model = sm.plainconv_3d_fullres().cuda()          # all layouts: marching fwd / dgrad, stride-2 dgrad classes, generic, tconv, k5
        x = torch.randn(1, 1, 64, 64, 64, device="cuda")
        run = lambda: _one_step(model, x)
    else:
        model = flavr.UNet_3D_3D(2, "unet_18", 4, 4, False, "concat", "transpose", use_uncertainty=False).cuda()
        x = torch.rand(1, 2, 4, 64, 64, device="cuda")
        run = lambda: _one_step(model, x.clone())
    Fn.clear_weight_cache()
    run()                                                  # fills the cache lazily (individual pack kernels)
    assert len(Fn._wcache) > 10
    with torch.no_grad():
        for p in model.parameters():
            p.add_(torch.randn_like(p) * 0.05)             # "optimizer step": new values, bumped version counters
    monkeypatch.setattr(Fn, "PACK_BATCHED", True)
    l0 = Fn.launches()
    n = Fn.refresh_weight_cache()                          # ONE batched launch on the side stream, in place
    assert n == len(Fn._wcache) and Fn.launches() - l0 <= len(Fn.PACK_BATCH_SPLITS) + 1
    Fn._join_prepack()
    torch.cuda.synchronize()
    batched = {k: v[2].clone() for k, v in Fn._wcache.items()}
    kinds = {k[1] for k in batched}
    n_batched = len(batched)
    Fn.clear_weight_cache()
    run()                                                  # the same copies through the individual kernels
    torch.cuda.synchronize()
    single = {k: v[2] for k, v in Fn._wcache.items()}
    # (FLAVR sees its 2-D convs as 5-D views made on the fly: those few copies are keyed by a temporary and are re-packed per call)
    common = set(single) & set(batched)
    assert len(common) >= 0.8 * len(batched)
    for k in common:
        assert torch.equal(single[k].view(torch.int16), batched[k].view(torch.int16)), k[1:]
    if which == "segmodel_64":
        assert {"march_fwd", "march_dgrad", "s2dgrad", "fwd", "dgrad", "tconv_fused"} <= kinds, kinds


def test_batch_recording_is_per_thread_and_abortable():
    from rehrseg_b200._lib import lib
    assert lib().rehr_pack_batch_launch(None) != 0          # nothing recorded: refused, no launch
    assert lib().rehr_pack_batch_begin() == 0
    assert lib().rehr_pack_batch_abort() == 0
    assert lib().rehr_pack_batch_launch(None) != 0

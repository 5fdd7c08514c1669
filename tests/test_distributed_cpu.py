"""CPU / gloo, world_size 2: the N > 1 host logic -- tile round-robin + one all-reduce of the sharded sliding window
(rehrseg_b200.sliding_window.predict_sliding_window_sharded) and the flat-bucket gradient mean bench.py uses -- checked
against the single-process oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _net():
    torch.manual_seed(7)
    conv = torch.nn.Conv3d(1, 2, 3, padding=1)
    conv.requires_grad_(False)
    return conv


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import third_party as tp, volume as ov
    from rehrseg_b200 import sliding_window as sw
    data = torch.randn((1, 24, 20, 28), generator=torch.Generator().manual_seed(11))
    patch = [16, 16, 16]
    slicers = sw._internal_get_sliding_window_slicers(data.shape[1:], patch_size=patch)
    net = _net()
    g = tp.compute_gaussian(tuple(patch), sigma_scale=1. / 8, value_scaling_factor=10, device=torch.device("cpu"))
    seen = []

    def accumulate(d, mine):  # CPU stand-in for the rehr_sw_accumulate kernel (same fp16 arithmetic)
        logits = torch.zeros((2, *d.shape[1:]), dtype=torch.half)
        npred = torch.zeros(d.shape[1:], dtype=torch.half)
        for i, sl in mine:
            seen.append(i)
            pred = ov.mirror_and_predict(net, d[sl][None], None, False)[0]
            logits[sl] += pred * g
            npred[sl[1:]] += g
        return logits, npred

    out = sw.predict_sliding_window_sharded(data, slicers, net, accumulate_fn=accumulate)
    # flat-bucket gradient mean as in bench.py
    grads = [torch.full((3, 2), float(rank + 1)), torch.full((5,), float(10 * (rank + 1)))]
    flat = torch.cat([t.reshape(-1) for t in grads])
    dist.all_reduce(flat)
    flat /= world
    q.put((rank, seen, out.numpy(), flat.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sliding_window_and_grad_bucket_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import volume as ov
    data = torch.randn((1, 24, 20, 28), generator=torch.Generator().manual_seed(11))
    slicers = ov.sliding_window_slicers(data.shape[1:], [16, 16, 16])
    want = ov.sliding_window_logits(data, slicers, _net(), None, 1, [16, 16, 16], True, False).float().numpy()
    assert sorted(res[0][1] + res[1][1]) == list(range(len(slicers)))           # every tile exactly once
    assert res[0][1] == list(range(0, len(slicers), 2)) and res[1][1] == list(range(1, len(slicers), 2))
    for _, _, out, flat in res:
        assert np.array_equal(out, res[0][2])                                     # all ranks agree
        err = np.linalg.norm(out.astype(np.float32) - want) / np.linalg.norm(want)
        assert err <= 2e-3, err                                                   # fp16 sums, different order
        assert np.allclose(flat[:6], 1.5) and np.allclose(flat[6:], 15.0)

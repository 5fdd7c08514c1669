"""Development tool (run under torchrun): config 3 -- sliding-window Gaussian-blended inference of the SegModel over a synthetic
256^3 volume with the 27 tiles dealt round-robin over the ranks (rehrseg_b200.sliding_window.predict_sliding_window_sharded),
and a 1-vs-N consistency check of the blended logits.  Prints volumes/s on rank 0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from rehrseg_b200 import seg_model as sm, sliding_window as sw

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = sm.plainconv_3d_fullres().to(dev).eval()
vol = torch.randn((1, 256, 256, 256), generator=torch.Generator().manual_seed(3)).to(dev)
patch = [128, 128, 128]
slicers = sw._internal_get_sliding_window_slicers(vol.shape[1:], patch_size=patch)


def run():
    with torch.no_grad():
        return sw.predict_sliding_window_sharded(vol, slicers, model, out_idx=0, patch_size=patch, use_gaussian=True, deep_supervision=False)


out = run()  # warm-up (graph capture, gaussian)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 3
for _ in range(n):
    out = run()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t = (time.perf_counter() - t0) / n
tt = torch.tensor([t], device=dev)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    with torch.no_grad():
        ref = sw._internal_predict_sliding_window_return_logits(vol, slicers, model, True, 0, 1, patch, use_gaussian=True, deep_supervision=False)
    err = float((out.float() - ref.float()).norm() / ref.float().norm())
    agree = float((out.argmax(0) == ref.argmax(0)).double().mean())
    print(f"C3 sharded sliding window, {world} GPU(s): {float(tt):.3f} s/volume = {1 / float(tt):.2f} volumes/s; "
          f"vs single-GPU tile order: rel L2 {err:.2e}, argmax agreement {agree:.5f}", flush=True)
if world > 1:
    dist.destroy_process_group()

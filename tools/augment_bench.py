"""Development tool: the stage-2 spatial augmentation of one batch (B = 2 patches [1,16,256,256] + LR / HR segmentations + uncertainty,
rotation and scaling forced on) on the GPU (rehrseg_b200.augment) next to the oracle's CPU path (scipy map_coordinates, one process)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rehrseg_b200 import augment
from oracle import augment as oa

rng = np.random.RandomState(0)
b, z, X, Y = 2, 16, 256, 256
dd = {"data": rng.randn(b, 1, z, X, Y).astype(np.float32), "seg": (rng.rand(b, 1, z, X, Y) > 0.5).astype(np.float32),
      "seg_sr": (rng.rand(b, 1, 4 * z, X, Y) > 0.5).astype(np.float32), "uncertainty": rng.rand(b, 1, z, X, Y).astype(np.float32)}
dev = {k: torch.from_numpy(v).cuda() for k, v in dd.items()}


class Forced:          # every sample rotates and scales (the reference: 20 % each)
    def __init__(self, seed):
        self.r = np.random.RandomState(seed)
    def uniform(self, *a):
        return self.r.uniform(*a) if a else 0.0
    def random(self):
        return self.r.random_sample()


for _ in range(3):
    augment.spatial_transform_dummy_2d(dev, (z, X, Y), rng=Forced(1), seg_labels=(0.0, 1.0))
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 20
for _ in range(n):
    out = augment.spatial_transform_dummy_2d(dev, (z, X, Y), rng=Forced(1), seg_labels=(0.0, 1.0))
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"GPU: batch of {b} patches, {z + z + 4 * z + z} slices of {X}x{Y} each: {dt * 1e3:.2f} ms = {b / dt:.0f} patches/s")
t0 = time.perf_counter()
oa.spatial_transform_dummy_2d({k: v.copy() for k, v in dd.items()}, (z, X, Y), rng=Forced(1))
dt = time.perf_counter() - t0
print(f"CPU (scipy map_coordinates, one process): {dt:.2f} s per batch = {b / dt:.2f} patches/s")

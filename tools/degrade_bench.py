"""Development tool: SR-stage sample synthesis rate (rehrseg_b200.degrade.SRTrainSampler.batch on the GPU) next to the oracle's CPU
restatement of the reference's `__getitem__` on the same volume (512 x 512 x 160, patch 96 x 96 x 1 (the 2-D pairs the SR stage trains on), slice separation 4)."""
import sys, os, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rehrseg_b200 import degrade
from oracle import degrade as od

rng = np.random.RandomState(0)
X, Y, Z = 512, 512, 160
img = rng.rand(X, Y, Z, 1).astype(np.float32)
lab = (rng.rand(X, Y, Z, 1) > 0.7).astype(np.uint8)
taps = np.exp(-0.5 * ((np.arange(9.0) - 4) / (3.873 / 2.355)) ** 2)
kernel = torch.tensor(taps / taps.sum(), dtype=torch.float32).reshape(1, 1, 9, 1)
ps, sep, B = [96, 96, 1], 4.0, 32
ds = degrade.SRTrainSampler(ps, sep, blur=True, random_flip=True, blur_kernel=kernel.cuda())
t0 = time.perf_counter()
ds.add_subject(img, lab)
torch.cuda.synchronize()
print(f"preload + two blur passes of a 512x512x160 volume on the GPU: {time.perf_counter() - t0:.3f} s")
random.seed(0)
for _ in range(3):
    ds.batch([0] * B)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 20
for _ in range(n):
    lr, hr = ds.batch([0] * B)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"GPU batch of {B}: lr {tuple(lr.shape)} hr {tuple(hr.shape)}  {dt * 1e3:.2f} ms = {B / dt:.0f} samples/s (wall clock, host launches included)")
image = np.concatenate([img, lab.astype(np.float32)], axis=-1)
t0 = time.perf_counter()
fx, fy = od.blur_prefilter(image, kernel)
print(f"CPU pre-filter (reference path): {time.perf_counter() - t0:.3f} s")
random.seed(0)
t0 = time.perf_counter()
m = 64
for _ in range(m):
    od.train_sample(img, lab, fx, fy, ps, sep, True, True)
dt = (time.perf_counter() - t0) / m
print(f"CPU restatement of the reference __getitem__ (one process, as num_workers=0): {dt * 1e3:.2f} ms per sample = {1 / dt:.0f} samples/s")

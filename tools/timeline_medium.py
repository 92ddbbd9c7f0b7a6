import os, sys
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
from __graft_entry__ import load_package
sa = load_package()
al = sa.Aligner(0)
rng = np.random.default_rng(0)
nb, L = 8, 8191
T = rng.integers(0, 22, nb * L, dtype=np.uint8); P = rng.integers(0, 22, nb * L, dtype=np.uint8)
toff = np.arange(nb + 1, dtype=np.int64) * L; poff = toff.copy()
mat = np.full((23, 23), -2, np.int32); np.fill_diagonal(mat, 6)
for _ in range(2): al.align_batch(0, 23, mat, 5, T, toff, P, poff)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    al.align_batch(0, 23, mat, 5, T, toff, P, poff)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev: print(f"{(e.time_range.start - t0) / 1e3:8.3f} ms +{(e.time_range.end - e.time_range.start) / 1e3:7.3f} ms {e.name[:40]}")

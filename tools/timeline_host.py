#!/usr/bin/env python
"""Dev tool: kernel + copy timeline (CUPTI through torch.profiler) of one sa_align_batch call on pinned host buffers.
usage: python tools/timeline_host.py [pairs]   -> start (ms), duration (ms), name; short events are merged per kind"""
import os
import sys

import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402
import synth, helpers  # noqa: E402

sa = load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
T, toff, P, poff = synth.synthetic_batch(N, seed=2024)
mat = helpers.matrices()["protein/blosum62.txt"]
torch.zeros(1, device="cuda")
arena = int(toff[-1] + poff[-1])
T, toff, P, poff = (sa.pinned_copy(x) for x in (T, toff, P, poff))
out = dict(results=sa.pinned_empty(N, sa.RESULT_DTYPE), aln_off=sa.pinned_empty(N, np.uint64),
           aligned_text=sa.pinned_empty(arena, np.uint8), aligned_pattern=sa.pinned_empty(arena, np.uint8))
al = sa.Aligner(0)


def step():
    al.align_batch(1, 23, mat, 5, T, toff, P, poff, out=out)


step(); step()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    d = (e.time_range.end - e.time_range.start) / 1e3
    name = e.name
    if d < 0.05 and "Memcpy" not in name:
        continue
    if "Memcpy" in name and d < 0.1:
        continue
    print(f"{(e.time_range.start - t0) / 1e3:9.3f} ms  +{d:8.3f} ms  {name[:60]}")

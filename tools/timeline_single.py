#!/usr/bin/env python
"""Dev tool: kernel timeline (CUPTI through torch.profiler) of one single-pair call.  usage: timeline_single.py n m [mode]"""
import os, sys, json
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
from __graft_entry__ import load_package
import synth
sa = load_package()
n = int(sys.argv[1]); mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t, p = synth.synthetic_pair(n, 12345, 54321)
blast = np.full((4, 4), -4, np.int32); np.fill_diagonal(blast, 5)
al = sa.Aligner(0)
os.environ["SA_FORCE_PATH"] = "long"
for _ in range(2):
    al.align(mode, 4, blast, 5, t, p)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    al.align(mode, 4, blast, 5, t, p)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    print(f"{(e.time_range.start - t0) / 1e3:9.3f} ms  +{(e.time_range.end - e.time_range.start) / 1e3:8.3f} ms  {e.name[:60]}")

#!/usr/bin/env python
"""Dev tool: kernel timeline (CUPTI through torch.profiler) of one pipelined device-batch step.
usage: python tools/timeline.py [pairs]   -> prints stream, start (ms), duration (ms), kernel name"""
import os
import sys
import json

import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
from __graft_entry__ import load_package  # noqa: E402
import synth  # noqa: E402

sa = load_package()
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
T, toff, P, poff = synth.synthetic_batch(pairs, seed=2024)
mat = np.asarray(json.load(open(os.path.join(ROOT, "tests/golden/matrices.json")))["protein/blosum62.txt"], np.int32)
dev = torch.device("cuda:0")
dT, dP = torch.from_numpy(T).to(dev), torch.from_numpy(P).to(dev)
dto, dpo = torch.from_numpy(toff).to(dev), torch.from_numpy(poff).to(dev)
arena = int(toff[-1] + poff[-1])
res = torch.zeros(pairs * 4, dtype=torch.int64, device=dev)
aoff = torch.zeros(pairs, dtype=torch.int64, device=dev)
oT = torch.zeros(arena, dtype=torch.uint8, device=dev)
oP = torch.zeros(arena, dtype=torch.uint8, device=dev)
max_n = int((toff[1:] - toff[:-1]).max()); max_m = int((poff[1:] - poff[:-1]).max())
al = sa.Aligner(0)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)


def step():
    al.align_batch_device(1, 23, mat, 5, pairs, dT.data_ptr(), dto.data_ptr(), dP.data_ptr(), dpo.data_ptr(), res.data_ptr(),
                          aoff.data_ptr(), oT.data_ptr(), oP.data_ptr(), arena, max_n, max_m, stream=ts.cuda_stream)
    torch.cuda.synchronize()


step(); step()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    d = (e.time_range.end - e.time_range.start) / 1e3
    if d < 0.02:
        continue
    print(f"{(e.time_range.start - t0) / 1e3:9.3f} ms  +{d:8.3f} ms  {e.name[:70]}")

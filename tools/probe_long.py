#!/usr/bin/env python
"""Dev tool: per-step cost and per-strip lag of the long-pair kernel (fill-only timings)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
from __graft_entry__ import load_package
sa = load_package()
al = sa.Aligner(0)
os.environ["SA_FORCE_PATH"] = "long"
rng = np.random.default_rng(0)
blast = np.full((4, 4), -4, np.int32); np.fill_diagonal(blast, 5)
mode = int(os.environ.get("PROBE_MODE", "0"))
for R in [int(x) for x in os.environ.get("PROBE_R", "4,8,16").split(",")]:
    os.environ["SA_LONG_R"] = str(R)
    for strips, n in [(1, 50000), (2, 50000), (4, 50000), (16, 50000), (64, 50000), (64, 5000), (256, 20000)]:
        m = 32 * R * strips
        t = rng.integers(0, 4, n, dtype=np.uint8); p = rng.integers(0, 4, m, dtype=np.uint8)
        best = 1e9
        for _ in range(3):
            al.fill_only(mode, 4, blast, 5, t, p)
            best = min(best, al.timing()["fill_us"])
        steps = n + 31
        print(f"R={R:2d} strips={strips:4d} n={n:6d}: fill {best/1e3:8.3f} ms  {best*1e3/steps:7.1f} ns/step(n+31)  "
              f"{(m+1)*(n+1)/best/1e3:8.1f} GCUPS", flush=True)

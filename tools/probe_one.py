#!/usr/bin/env python
"""Dev tool: one fill-only call of the long kernel (for ncu): PROBE_R, PROBE_STRIPS, PROBE_N, PROBE_MODE."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
from __graft_entry__ import load_package
sa = load_package()
al = sa.Aligner(0)
os.environ["SA_FORCE_PATH"] = "long"
R = int(os.environ.get("PROBE_R", "4")); strips = int(os.environ.get("PROBE_STRIPS", "1")); n = int(os.environ.get("PROBE_N", "20000"))
os.environ["SA_LONG_R"] = str(R)
rng = np.random.default_rng(0)
blast = np.full((4, 4), -4, np.int32); np.fill_diagonal(blast, 5)
t = rng.integers(0, 4, n, dtype=np.uint8); p = rng.integers(0, 4, 32 * R * strips, dtype=np.uint8)
for _ in range(2):
    print(al.fill_only(int(os.environ.get("PROBE_MODE", "0")), 4, blast, 5, t, p), al.timing()["fill_us"])

#!/usr/bin/env python
"""Dev tool: one fill-only call of the 100 000 x 95 217 pair (for ncu)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
from __graft_entry__ import load_package
import synth
sa = load_package()
al = sa.Aligner(0)
t, p = synth.synthetic_pair(100000, 12345, 54321)
blast = np.full((4, 4), -4, np.int32); np.fill_diagonal(blast, 5)
for _ in range(2):
    print(al.fill_only(0, 4, blast, 5, t, p), al.timing()["fill_us"])

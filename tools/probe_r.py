#!/usr/bin/env python
"""Dev tool: fill-only time of square-ish single pairs for every strip height R."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
sa = load_package()
al = sa.Aligner(0)
os.environ["SA_FORCE_PATH"] = "long"
rng = np.random.default_rng(0)
mat = np.full((23, 23), -2, np.int32); np.fill_diagonal(mat, 6)
for (m, n) in [(3700, 3900), (8192, 8192), (32768, 32768), (95000, 100000), (250000, 250000)]:
    t = rng.integers(0, 22, n, dtype=np.uint8); p = rng.integers(0, 22, m, dtype=np.uint8)
    row = []
    for R in (2, 4, 6, 8, 12, 16):
        os.environ["SA_LONG_R"] = str(R)
        best = 1e9
        for _ in range(2):
            al.fill_only(int(os.environ.get("PROBE_MODE", "0")), 23, mat, 5, t, p)
            best = min(best, al.timing()["fill_us"])
        row.append(f"R={R}:{best/1e3:8.3f}ms")
    os.environ.pop("SA_LONG_R")
    al.fill_only(0, 23, mat, 5, t, p)
    print(f"{m:7d} x {n:7d}: " + "  ".join(row) + f"   auto:{al.timing()['fill_us']/1e3:8.3f}ms", flush=True)

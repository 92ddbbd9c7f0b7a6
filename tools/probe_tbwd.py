import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
from __graft_entry__ import load_package
import synth
sa = load_package()
al = sa.Aligner(0)
rng = np.random.default_rng(0)
mat = np.full((23, 23), -2, np.int32); np.fill_diagonal(mat, 6)
blast = np.full((4, 4), -4, np.int32); np.fill_diagonal(blast, 5)
cases = []
for L in (8191, 32767):
    cases.append((f"random protein {L}", 23, mat, rng.integers(0, 22, L, dtype=np.uint8), rng.integers(0, 22, L, dtype=np.uint8)))
t, p = synth.synthetic_pair(100000, 12345, 54321)
cases.append(("c3 similar dna", 4, blast, t, p))
for wd in ("", "8", "16", "32"):
    if wd: os.environ["SA_TB_WD"] = wd
    else: os.environ.pop("SA_TB_WD", None)
    row = []
    for name, alpha, m, t, p in cases:
        for mode in (0, 1):
            best = 1e9
            for _ in range(3):
                al.align(mode, alpha, m, 5, t, p); best = min(best, al.timing()["traceback_us"])
            row.append(f"{name}/{'NW' if mode == 0 else 'SW'}: tb {best / 1e3:.2f} ms")
    print("WD", wd or "auto", " | ".join(row), flush=True)

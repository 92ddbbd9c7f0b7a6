"""Fill / traceback time of the long-pair kernels on one pair for a list of tile shapes (dev tool, needs a GPU).
    python tools/probe_tile.py [length] [shape ...]      shape = R,C  or  strip:R (the one-column kernel)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
import synth  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

sa = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
shapes = sys.argv[2:] or ["4,4", "8,4", "8,2", "4,8", "2,8", "16,4", "8,8", "strip:8"]
mode = int(os.environ.get("MODE", "0"))
protein = os.environ.get("PROTEIN", "0") == "1"
t, p = synth.synthetic_pair(n, 12345, 54321, protein=protein)
if os.environ.get("SLICE"):          # a column slice of the pair (config 5 shape: few columns, all rows)
    t = t[:int(os.environ["SLICE"])]
if os.environ.get("M"):
    p = p[:int(os.environ["M"])]
alpha = 23 if protein else 4
mat = np.full((alpha, alpha), -4, np.int32)
np.fill_diagonal(mat, 5)
al = sa.Aligner(0)
cells = (len(t) + 1) * (len(p) + 1)
for sh in shapes:
    for k in ("SA_TILE", "SA_LONG_R", "SA_LONG_KERNEL"):
        os.environ.pop(k, None)
    if sh.startswith("strip:"):
        os.environ["SA_LONG_R"] = sh.split(":")[1]
    else:
        os.environ["SA_TILE"] = sh
    os.environ["SA_FORCE_PATH"] = "long"
    best = None
    for rep in range(int(os.environ.get("REPS", "4"))):
        a = al.align(mode, alpha, mat, 5, t, p)
        tm = al.timing()
        if best is None or tm["fill_us"] < best["fill_us"]:
            best = tm
    print(f"{sh:8s} n={len(t)} m={len(p)} mode={mode} fill {best['fill_us'] / 1e3:8.3f} ms ({cells / best['fill_us'] / 1e3:7.1f} GCUPS)  "
          f"traceback {best['traceback_us'] / 1e3:7.3f} ms  score {a.score} len {a.aln_len}", flush=True)
al.close()

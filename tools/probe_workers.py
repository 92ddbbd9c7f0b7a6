import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
from __graft_entry__ import load_package
sa = load_package()
rng = np.random.default_rng(0)
nb, L = 32, 8191
T = rng.integers(0, 22, nb * L, dtype=np.uint8); P = rng.integers(0, 22, nb * L, dtype=np.uint8)
toff = np.arange(nb + 1, dtype=np.int64) * L; poff = toff.copy()
mat = np.full((23, 23), -2, np.int32); np.fill_diagonal(mat, 6)
for w in (1, 2, 4, 8, 16):
    os.environ["SA_LONG_WORKERS"] = str(w)
    al = sa.Aligner(0)
    al.align_batch(0, 23, mat, 5, T, toff, P, poff)
    t0 = time.perf_counter(); al.align_batch(0, 23, mat, 5, T, toff, P, poff); dt = time.perf_counter() - t0
    print("workers", w, f"{dt * 1e3:.1f} ms", al.timing()["fill_us"] / 1e3, al.timing()["traceback_us"] / 1e3)
    al.close()

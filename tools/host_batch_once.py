"""One warm-up and one timed sa_align_batch call on host buffers (for an ncu launch list of the staged host
pipeline): python tools/host_batch_once.py [pairs]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
import synth, helpers
sa = load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
T, toff, P, poff = synth.synthetic_batch(N, seed=2024)
mat = helpers.matrices()["protein/blosum62.txt"]
al = sa.Aligner(0)
arena = int(toff[-1] + poff[-1])
for kind in (["pageable"] if os.environ.get("ONCE_PAGEABLE_ONLY") else ["pageable", "pinned"]):
    if kind == "pinned":
        T, toff, P, poff = (sa.pinned_copy(x) for x in (T, toff, P, poff))
        outb = dict(results=sa.pinned_empty(N, sa.RESULT_DTYPE), aln_off=sa.pinned_empty(N, np.uint64),
                    aligned_text=sa.pinned_empty(arena, np.uint8), aligned_pattern=sa.pinned_empty(arena, np.uint8))
    else:
        outb = None
    for _ in range(3):
        t0 = time.perf_counter()
        out = al.align_batch(1, 23, mat, 5, T, toff, P, poff, out=outb)
        dt = time.perf_counter() - t0
    t = al.timing()
    print(f"{N} pairs, {kind} host buffers: {dt * 1e3:.2f} ms, launches {t['kernel_launches']}, d2h {t['d2h_bytes']} B, "
          f"packed bytes {int(out['results']['aln_len'].sum())}", flush=True)

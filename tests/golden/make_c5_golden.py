"""Pins BASELINE config 5: the score of the 1 000 000 x ~951 000 global alignment, computed on the CPU by
oracle/sa_oracle.c's two-row restatement of fillMatrixNW (alignSequenceCPU.cpp:232-277; the reference itself would need
a 1 TB direction matrix, :305).  ~10^12 cells: tens of minutes on one core.  Writes tests/golden/c5_golden.json, which
bench.py / bench_c5.py and tests/test_gpu_parity.py compare the GPU result with.

    python tests/golden/make_c5_golden.py [length]        (default 1000000; 100000 reproduces the C3-sized check quickly)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
import synth  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402


def main():
    length = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    t, p = synth.synthetic_pair(length, 777, 778)
    blast = np.array([[5 if i == j else -4 for j in range(4)] for i in range(4)], np.int32)
    t0 = time.time()
    score, _ = Oracle().score_only(0, 4, blast, 5, t, p)
    dt = time.time() - t0
    path = os.path.join(ROOT, "tests", "golden", "c5_golden.json")
    rec = json.load(open(path)) if os.path.exists(path) else {}
    rec[str(length)] = dict(n=int(len(t)), m=int(len(p)), score=int(score), seeds=[777, 778], matrix="dna/blast.txt", gap=5, mode="global",
                            how="oracle/sa_oracle.c sa_oracle_score_only (two rolling rows), 1 thread", seconds=round(dt, 1),
                            gcups=round((len(t) + 1) * (len(p) + 1) / dt / 1e9, 3))
    json.dump(rec, open(path, "w"), indent=1, sort_keys=True)
    print(rec[str(length)])


if __name__ == "__main__":
    main()

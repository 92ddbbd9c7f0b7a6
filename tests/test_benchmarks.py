"""The reference's benchmark harness on the B200 path (benchmarks.py): the tables keep the format of
tests/benchmarks.cu (`-----  rows x cols  -----`, `GPU = ... ms`, `MCUPS: ...`) and the scores behind the timed fills
equal the CPU oracle's."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import helpers
from gpu_common import ROOT, load_package

pytestmark = pytest.mark.gpu


def test_throughput_and_latency_tables_keep_the_reference_format():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks.py"), "--mode", "throughput", "--max-size", "1024"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    txt = out.stdout
    assert "Global alignment benchmark:" in txt and "Local alignment benchmark:" in txt
    sizes = re.findall(r"-----  (\d+) x (\d+)  -----", txt)
    assert ("256", "256") in sizes and ("1024", "1024") in sizes and ("1024", "32768") in sizes
    mcups = [int(x) for x in re.findall(r"MCUPS: (\d+)", txt)]
    assert len(mcups) == len(sizes) and all(m > 0 for m in mcups)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks.py"), "--mode", "latency", "--max-size", "512"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert len(re.findall(r"GPU = [0-9.]+ ms", out.stdout)) >= 4


def test_fill_only_scores_of_the_harness_recipe_vs_oracle():
    """fillDummyRequest's recipe (random residues in 0..21, blosum50, gap 5; benchmarks.cu:21-42): the fill-only entry the
    throughput table times returns the oracle's score, global and local, on the harness' short-and-wide shapes."""
    sys.path.insert(0, ROOT)
    from oracle.oracle_py import Oracle
    sa = load_package()
    oracle = Oracle()
    al = sa.Aligner(0)
    mat = helpers.matrices()["protein/blosum50.txt"]
    rng = np.random.default_rng(0)
    try:
        for rows, cols in ((256, 256), (257, 4097), (1024, 2048)):
            t, p = rng.integers(0, 22, cols - 1, dtype=np.uint8), rng.integers(0, 22, rows - 1, dtype=np.uint8)
            for mode in (0, 1):
                assert al.fill_only(mode, 23, mat, 5, t, p)[0] == oracle.align(mode, 23, mat, 5, t, p).score, (rows, cols, mode)
    finally:
        al.close()

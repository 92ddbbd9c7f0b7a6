#!/usr/bin/env python
"""Dev tool: a host batch with the whole length spectrum the batch kernels take (text 1..4096, pattern 1..1536, tiny,
ragged, similar and unrelated pairs) against the oracle, both modes and alphabets.  Test infrastructure (uses oracle/)."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
from __graft_entry__ import load_package
import helpers
from oracle.oracle_py import Oracle
sa = load_package(); al = sa.Aligner(0); orc = Oracle()
rng = np.random.default_rng(2025)
mats = helpers.matrices()
bad = 0
for alpha, mat in ((4, mats["dna/blast.txt"]), (23, mats["protein/blosum50.txt"])):
    for mode in (0, 1):
        N = 700
        n = np.concatenate((rng.integers(1, 40, 150), rng.integers(1, 700, 400), rng.integers(700, 4097, 150)))
        m = np.concatenate((rng.integers(1, 40, 150), rng.integers(1, 400, 400), rng.integers(300, 1537, 150)))
        rng.shuffle(n); rng.shuffle(m)
        toff = np.concatenate(([0], np.cumsum(n))).astype(np.int64); poff = np.concatenate(([0], np.cumsum(m))).astype(np.int64)
        T = rng.integers(0, alpha, toff[-1], dtype=np.uint8); P = rng.integers(0, alpha, poff[-1], dtype=np.uint8)
        # make a third of them similar pairs
        for i in range(0, N, 3):
            k = min(n[i], m[i]); P[poff[i]:poff[i] + k] = T[toff[i]:toff[i] + k]
        gap = int(rng.integers(1, 12))
        out = al.align_batch(mode, alpha, mat, gap, T, toff, P, poff)
        for i in range(N):
            w = orc.align(mode, alpha, mat, gap, T[toff[i]:toff[i + 1]], P[poff[i]:poff[i + 1]])
            g = sa.unpack_batch(out, i)
            if g.key() != w.key():
                bad += 1
                if bad < 5: print("MISMATCH", alpha, mode, i, n[i], m[i], g.score, w.score, g.aln_len, w.aln_len)
        print("alpha", alpha, "mode", mode, "gap", gap, "done, mismatches so far", bad, flush=True)
print("TOTAL MISMATCHES", bad)

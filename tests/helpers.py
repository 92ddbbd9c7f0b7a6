"""Shared helpers for the parity tests: golden fixtures and input resolution."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
import synth  # noqa: E402

_seqs = None
_mats = None
_goldens = None


def sequences():
    global _seqs
    if _seqs is None:
        z = np.load(os.path.join(GOLD, "sequences.npz"))
        _seqs = {k: z[k] for k in z.files}
    return _seqs


def matrices():
    global _mats
    if _mats is None:
        _mats = {k: np.asarray(v, np.int32) for k, v in json.load(open(os.path.join(GOLD, "matrices.json"))).items()}
    return _mats


def goldens():
    global _goldens
    if _goldens is None:
        _goldens = json.load(open(os.path.join(GOLD, "reference_goldens.json")))
    return _goldens


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def golden_inputs(g):
    """-> (text, pattern, matrix) for a golden record (text is the longer)."""
    if "text" in g:
        t, p = np.asarray(g["text"], np.uint8), np.asarray(g["pattern"], np.uint8)
    elif "files" in g:
        a, b = (sequences()[f] for f in g["files"])
        t, p = (a, b) if len(a) >= len(b) else (b, a)
    elif "synth" in g:
        s = g["synth"]
        t, p = synth.synthetic_pair(s["n"], s["seed_base"], s["seed_mut"], protein=s["protein"])
    else:
        raise KeyError(g["name"])
    mat = np.asarray(g["matrix_values"], np.int32) if "matrix_values" in g else matrices()[g["matrix"]]
    assert len(t) == g["n"] and len(p) == g["m"], g["name"]
    return t, p, mat


def has_inputs(g):
    if "files" in g:
        return all(f in sequences() for f in g["files"])
    return "text" in g or "synth" in g


def check_against_golden(aln, g):
    """aln: object with score/aln_len/start_text/start_pattern/aligned_text/aligned_pattern."""
    r = g["ref"]
    assert aln.score == r["score"], (g["name"], "score", aln.score, r["score"])
    assert aln.aln_len == r["aln_len"], (g["name"], "aln_len", aln.aln_len, r["aln_len"])
    assert aln.start_text == r["start_text"], (g["name"], "start_text", aln.start_text, r["start_text"])
    assert aln.start_pattern == r["start_pattern"], (g["name"], "start_pattern")
    assert sha(aln.aligned_text) == r["sha_text"], (g["name"], "aligned_text")
    assert sha(aln.aligned_pattern) == r["sha_pattern"], (g["name"], "aligned_pattern")
    for k, v in g.get("expect", {}).items():   # values copied from the reference's tests
        got = getattr(aln, k)
        got = got.decode() if isinstance(got, bytes) else got
        assert got == v, (g["name"], k)


def random_case(rng, alpha, n_max=200, m_max=None, similar=True):
    """Random pair; similar=True makes the pattern a mutated copy (long paths, many ties)."""
    n = int(rng.integers(1, n_max + 1))
    t = rng.integers(0, alpha, n, dtype=np.uint8)
    if similar:
        p = synth.mutate_indices_numpy(t, rng, alpha)
        if len(p) == 0:
            p = t[:1].copy()
    else:
        m = int(rng.integers(1, (m_max or n) + 1))
        p = rng.integers(0, alpha, m, dtype=np.uint8)
    if len(p) > len(t):
        t, p = p, t
    return t, p


class NumpyStripEngine:
    """CPU stand-in for GpuStripEngine (test infrastructure): one column slice of a global alignment,
    restating alignSequenceCPU.cpp:232-277 (fill: DIAG only on a strict win, gap tie -> LEFT) and :64-114
    (traceback) on the slice.  Columns travel as 4*H like on the device."""

    def __init__(self, matrix, gap, text_slice, col0, pattern, alphabet):
        import torch
        self.torch = torch
        self.S, self.g = np.asarray(matrix, np.int64), int(gap)
        self.t, self.p = np.asarray(text_slice, np.int64), np.asarray(pattern, np.int64)
        self.n, self.m, self.col0, self.alphabet = len(self.t), len(self.p), int(col0), alphabet
        self.n_total = None            # set by the caller when the pipelined protocol is used

    def fill(self, left_col):
        n, m, g = self.n, self.m, self.g
        H = np.zeros((m + 1, n + 1), np.int64)
        D = np.zeros((m + 1, n + 1), np.int8)      # 0 diag, 1 top, 2 left
        H[:, 0] = -g * np.arange(m + 1) if left_col is None else np.asarray(left_col.numpy(), np.int64) // 4
        H[0, :] = -g * (self.col0 + np.arange(n + 1))
        for i in range(1, m + 1):
            for j in range(1, n + 1):
                left, top, diag = H[i, j - 1] - g, H[i - 1, j] - g, H[i - 1, j - 1] + self.S[self.p[i - 1], self.t[j - 1]]
                if diag > max(left, top): H[i, j], D[i, j] = diag, 0
                elif left >= top: H[i, j], D[i, j] = left, 2
                else: H[i, j], D[i, j] = top, 1
        self.H, self.D = H, D
        return self.torch.from_numpy((4 * H[:, n]).astype(np.int32))

    # row-chunked form (strips.fill_slice_pipelined)
    def begin(self, chunk_rows_hint):
        n, m, g = self.n, self.m, self.g
        self.H = np.zeros((m + 1, n + 1), np.int64)
        self.D = np.zeros((m + 1, n + 1), np.int8)
        self.H[0, :] = -g * (self.col0 + np.arange(n + 1))
        return (int(chunk_rows_hint) + 7) // 8 * 8          # pretend strips are 8 rows high

    def column_buffer(self):
        return self.torch.zeros(self.m + 1, dtype=self.torch.int32)

    def row_buffer(self):
        return self.torch.zeros(max(self.n, 1), dtype=self.torch.int32)

    def fill_rows(self, row0, rows, left, right, top, bottom):
        n, g, H, D = self.n, self.g, self.H, self.D
        if top is not None:
            assert np.array_equal(np.asarray(top.numpy()[:n], np.int64), 4 * H[row0, 1:]), "top row hand-over"
        lo = row0 + 1 if row0 else 0
        H[lo:row0 + rows + 1, 0] = (-g * np.arange(lo, row0 + rows + 1) if left is None
                                    else np.asarray(left.numpy()[lo:row0 + rows + 1], np.int64) // 4)
        for i in range(row0 + 1, row0 + rows + 1):
            for j in range(1, n + 1):
                l_, t_, d_ = H[i, j - 1] - g, H[i - 1, j] - g, H[i - 1, j - 1] + self.S[self.p[i - 1], self.t[j - 1]]
                if d_ > max(l_, t_): H[i, j], D[i, j] = d_, 0
                elif l_ >= t_: H[i, j], D[i, j] = l_, 2
                else: H[i, j], D[i, j] = t_, 1
        right[lo:row0 + rows + 1] = self.torch.from_numpy((4 * H[lo:row0 + rows + 1, n]).astype(np.int32))
        if bottom is not None:
            bottom[:n] = self.torch.from_numpy((4 * H[row0 + rows, 1:]).astype(np.int32))

    def score(self):
        return int(self.H[self.m, self.n])

    def traceback(self, start_row):
        i, j = int(start_row), self.n
        gapc = self.alphabet[len(self.alphabet) - 1:]
        oT, oP = [], []
        ti, pi = j - 1, i - 1
        while j > 0 or (self.col0 == 0 and i > 0):
            d = 1 if j == 0 else 2 if i == 0 else int(self.D[i, j])
            takeT, takeP = d != 1, d != 2
            oT.append(self.alphabet[self.t[ti]:self.t[ti] + 1] if takeT else gapc)
            oP.append(self.alphabet[self.p[pi]:self.p[pi] + 1] if takeP else gapc)
            ti -= takeT; pi -= takeP; i -= takeP; j -= takeT
        return b"".join(reversed(oT)), b"".join(reversed(oP)), i, max(ti, 0), max(pi, 0)

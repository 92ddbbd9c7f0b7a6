"""ctypes bindings for the TEST-ONLY checkers under oracle/.

* ``Oracle``    -> oracle/libsa_oracle.so   (our C restatement, oracle/sa_oracle.c)
* ``Reference`` -> oracle/_ref/libsa_ref_O{0,3}.so (the unmodified reference,
  compiled by oracle/Makefile from /root/reference; travels to the GPU box as a
  prebuilt, git-ignored .so)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

DNA_ALPHABET = b"ATCG-"                       # SequenceAlignment.hpp:56
PROTEIN_ALPHABET = b"ARNDCQEGHILKMFPSTWYVBZX-"  # SequenceAlignment.hpp:57-58


def alphabet_for(alpha: int) -> bytes:
    return PROTEIN_ALPHABET if alpha == 23 else DNA_ALPHABET


@dataclass
class Alignment:
    score: int
    aln_len: int
    start_text: int
    start_pattern: int
    aligned_text: bytes
    aligned_pattern: bytes

    def key(self):
        return (self.score, self.aln_len, self.start_text, self.start_pattern,
                self.aligned_text, self.aligned_pattern)


class _Res(C.Structure):
    _fields_ = [("score", C.c_int32), ("aln_len", C.c_uint64),
                ("start_text", C.c_uint64), ("start_pattern", C.c_uint64)]


def build(ref: bool = True) -> None:
    """Compile the checkers (gcc for the restatement; nvcc on the reference's own
    sources for _ref when /root/reference exists)."""
    targets = ["libsa_oracle.so"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint8))


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


class Oracle:
    """oracle/sa_oracle.c"""

    def __init__(self):
        path = os.path.join(HERE, "libsa_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)
        self.lib.sa_oracle_align.restype = C.c_int
        self.lib.sa_oracle_align.argtypes = [
            C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_char_p,
            C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
            C.c_void_p, C.c_void_p, C.POINTER(_Res), C.c_void_p]
        self.lib.sa_oracle_score_only.restype = C.c_int
        self.lib.sa_oracle_score_only.argtypes = [
            C.c_int, C.c_int, C.c_void_p, C.c_int,
            C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
            C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
        self.lib.sa_oracle_rescore.restype = C.c_int64
        self.lib.sa_oracle_rescore.argtypes = [
            C.c_char_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_int, C.c_void_p, C.c_int]

    def align(self, mode, alpha, matrix, gap, text, pattern, want_dirs=False):
        text, pattern, matrix = _u8(text), _u8(pattern), _i32(matrix)
        n, m = len(text), len(pattern)
        outT = np.empty(max(1, n + m), np.uint8)
        outP = np.empty(max(1, n + m), np.uint8)
        dirs = np.empty((m + 1, n + 1), np.int8) if want_dirs else None
        res = _Res()
        rc = self.lib.sa_oracle_align(
            mode, alpha, matrix.ctypes.data, gap, alphabet_for(alpha),
            text.ctypes.data, n, pattern.ctypes.data, m,
            outT.ctypes.data, outP.ctypes.data, C.byref(res),
            dirs.ctypes.data if want_dirs else None)
        if rc:
            raise MemoryError("oracle allocation failed")
        aln = Alignment(res.score, res.aln_len, res.start_text, res.start_pattern,
                        outT[:res.aln_len].tobytes(), outP[:res.aln_len].tobytes())
        return (aln, dirs) if want_dirs else aln

    def score_only(self, mode, alpha, matrix, gap, text, pattern):
        text, pattern, matrix = _u8(text), _u8(pattern), _i32(matrix)
        score, arg = C.c_int32(), C.c_uint64()
        rc = self.lib.sa_oracle_score_only(mode, alpha, matrix.ctypes.data, gap,
                                           text.ctypes.data, len(text),
                                           pattern.ctypes.data, len(pattern),
                                           C.byref(score), C.byref(arg))
        if rc:
            raise MemoryError
        return score.value, arg.value

    def check_batch(self, mode, alpha, matrix, gap, text, text_off, pattern, pattern_off, out, idx=None, nthreads=0):
        """Every Response field and both strings of the pairs ``idx`` (default: all) of a sa_align_batch-style result
        ``out`` (results / aln_off / aligned_text / aligned_pattern arrays) against sa_oracle_align, on ``nthreads``
        host threads (0 = all).  Returns (mismatching pairs, first mismatching pair or -1)."""
        return _check_batch(self.lib.sa_oracle_check_batch, True, mode, alpha, matrix, gap, text, text_off, pattern,
                            pattern_off, out, idx, nthreads)

    def rescore(self, aligned_text: bytes, aligned_pattern: bytes, alpha, matrix, gap) -> int:
        matrix = _i32(matrix)
        return self.lib.sa_oracle_rescore(aligned_text, aligned_pattern, len(aligned_text),
                                          alphabet_for(alpha), alpha, matrix.ctypes.data, gap)


def _check_batch(fn, with_alphabet, mode, alpha, matrix, gap, text, text_off, pattern, pattern_off, out, idx, nthreads):
    text, pattern = _u8(text), _u8(pattern)
    toff = np.ascontiguousarray(text_off, dtype=np.int64)
    poff = np.ascontiguousarray(pattern_off, dtype=np.int64)
    full = np.zeros(max(23 * 23, alpha * alpha), np.int32)
    full[:alpha * alpha] = _i32(matrix).ravel()[:alpha * alpha]
    N = len(toff) - 1
    if idx is None:
        pidx, nidx = None, N
    else:
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        pidx, nidx = idx.ctypes.data, len(idx)
    res = np.ascontiguousarray(out["results"])
    assert res.dtype.itemsize == 32 and len(res) >= N
    aoff = np.ascontiguousarray(out["aln_off"], dtype=np.uint64)
    aT, aP = np.ascontiguousarray(out["aligned_text"]), np.ascontiguousarray(out["aligned_pattern"])
    first = C.c_int64(-1)
    fn.restype = C.c_uint64
    args = [C.c_int(mode), C.c_int(alpha), C.c_void_p(full.ctypes.data), C.c_int(gap)]
    if with_alphabet:
        args.append(C.c_char_p(alphabet_for(alpha)))
    args += [C.c_void_p(text.ctypes.data), C.c_void_p(toff.ctypes.data), C.c_void_p(pattern.ctypes.data),
             C.c_void_p(poff.ctypes.data), C.c_void_p(pidx), C.c_uint64(nidx), C.c_void_p(res.ctypes.data),
             C.c_void_p(aoff.ctypes.data), C.c_void_p(aT.ctypes.data), C.c_void_p(aP.ctypes.data),
             C.c_int(nthreads or (os.cpu_count() or 1)), C.byref(first)]
    bad = fn(*args)
    return int(bad), int(first.value)


class Reference:
    """The unmodified reference's CPU path (oracle/ref_shim.cu)."""

    def __init__(self, opt: str = "O3"):
        path = os.path.join(HERE, "_ref", f"libsa_ref_{opt}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.opt = opt
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_align_cpu.restype = C.c_int
        L.ref_align_cpu.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                    C.POINTER(C.c_int), C.POINTER(C.c_uint64),
                                    C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                    C.c_void_p, C.c_void_p]
        L.ref_align_cpu_batch.restype = C.c_int
        L.ref_align_cpu_batch.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_fill_cpu.restype = C.c_int
        L.ref_fill_cpu.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                   C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
        L.ref_read_sequence.restype = C.c_int64
        L.ref_read_sequence.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_uint64]
        L.ref_validate_and_transform.restype = C.c_int
        L.ref_validate_and_transform.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.ref_parse_score_matrix.restype = C.c_int
        L.ref_parse_score_matrix.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
        L.ref_pretty_print.restype = C.c_uint64
        L.ref_pretty_print.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_uint64,
                                       C.c_uint64, C.c_int, C.c_void_p, C.c_uint64]

    @staticmethod
    def available(opt: str = "O3") -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", f"libsa_ref_{opt}.so"))

    def align(self, mode, alpha, matrix, gap, text, pattern) -> Alignment:
        text, pattern = _u8(text), _u8(pattern)
        n, m = len(text), len(pattern)
        if n < m:
            raise ValueError("reference needs text >= pattern (2*text output capacity)")
        full = np.zeros(23 * 23, np.int32)
        full[:alpha * alpha] = _i32(matrix).ravel()[:alpha * alpha]
        outT = np.empty(max(1, 2 * n), np.uint8)
        outP = np.empty(max(1, 2 * n), np.uint8)
        score, ln, st, sp = C.c_int(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        rc = self.lib.ref_align_cpu(mode, alpha, full.ctypes.data, gap,
                                    text.ctypes.data, n, pattern.ctypes.data, m,
                                    C.byref(score), C.byref(ln), C.byref(st), C.byref(sp),
                                    outT.ctypes.data, outP.ctypes.data)
        if rc:
            raise MemoryError("reference returned error")
        return Alignment(score.value, ln.value, st.value, sp.value,
                         outT[:ln.value].tobytes(), outP[:ln.value].tobytes())

    def align_batch(self, mode, alpha, matrix, gap, text, text_off, pattern, pattern_off, nthreads=1):
        """Scores and alignment lengths of a CSR batch on ``nthreads`` host threads."""
        text, pattern = _u8(text), _u8(pattern)
        toff = np.ascontiguousarray(text_off, dtype=np.int64)
        poff = np.ascontiguousarray(pattern_off, dtype=np.int64)
        full = np.zeros(23 * 23, np.int32)
        full[:alpha * alpha] = _i32(matrix).ravel()[:alpha * alpha]
        N = len(toff) - 1
        scores = np.zeros(N, np.int32)
        lens = np.zeros(N, np.uint64)
        bad = self.lib.ref_align_cpu_batch(mode, alpha, full.ctypes.data, gap, text.ctypes.data, toff.ctypes.data,
                                           pattern.ctypes.data, poff.ctypes.data, N, nthreads,
                                           scores.ctypes.data, lens.ctypes.data)
        if bad:
            raise MemoryError(f"{bad} pairs failed in the reference")
        return scores, lens

    def check_batch(self, mode, alpha, matrix, gap, text, text_off, pattern, pattern_off, out, idx=None, nthreads=0):
        """Like Oracle.check_batch, but the truth is the UNMODIFIED alignSequenceCPU (ref_check_batch)."""
        return _check_batch(self.lib.ref_check_batch, False, mode, alpha, matrix, gap, text, text_off, pattern,
                            pattern_off, out, idx, nthreads)

    def fill(self, mode, alpha, matrix, gap, text, pattern, M=None):
        text, pattern = _u8(text), _u8(pattern)
        n, m = len(text), len(pattern)
        full = np.zeros(23 * 23, np.int32)
        full[:alpha * alpha] = _i32(matrix).ravel()[:alpha * alpha]
        if M is None:
            M = np.empty((m + 1) * (n + 1), np.int8)
        score, arg = C.c_int(), C.c_uint64()
        self.lib.ref_fill_cpu(mode, alpha, full.ctypes.data, gap, text.ctypes.data, n,
                              pattern.ctypes.data, m, M.ctypes.data, C.byref(score), C.byref(arg))
        return score.value, arg.value, M

    def read_sequence(self, fname: str, alpha: int) -> np.ndarray:
        cap = os.path.getsize(fname) + 16
        buf = np.empty(cap, np.uint8)
        k = self.lib.ref_read_sequence(fname.encode(), alpha, buf.ctypes.data, cap)
        if k < 0:
            raise ValueError(f"reference could not read {fname}")
        return buf[:k].copy()

    def validate_and_transform(self, s: bytes, alpha: int) -> np.ndarray:
        buf = np.frombuffer(s, np.uint8).copy()
        k = self.lib.ref_validate_and_transform(buf.ctypes.data, len(buf), alpha)
        return buf[:max(k, 0)].copy()

    def parse_score_matrix(self, fname: str, alpha: int) -> np.ndarray:
        out = np.zeros(alpha * alpha, np.int32)
        rc = self.lib.ref_parse_score_matrix(fname.encode(), alpha, out.ctypes.data)
        if rc != 0:
            raise ValueError(f"reference could not parse {fname}")
        return out

    def pretty_print(self, aln: Alignment) -> bytes:
        need = self.lib.ref_pretty_print(aln.aligned_text, aln.aligned_pattern, aln.aln_len,
                                         aln.start_text, aln.start_pattern, aln.score, None, 0)
        buf = C.create_string_buffer(int(need) + 1)
        self.lib.ref_pretty_print(aln.aligned_text, aln.aligned_pattern, aln.aln_len,
                                  aln.start_text, aln.start_pattern, aln.score, buf, need)
        return buf.raw[:need]


class ReferenceGpu:
    """The unmodified reference's GPU path (alignSequenceGPU.cu) as a secondary baseline on the same device.
    bench=True loads the -DBENCHMARK build: fill_micros() returns what tests/benchmarks.cu:171-175 measures
    (kernels + device-to-host copy of the 1 byte/cell direction matrix, no allocation / H2D / traceback)."""

    def __init__(self, bench: bool = True):
        path = os.path.join(HERE, "_ref", "libsa_refgpu_bench.so" if bench else "libsa_ref_O3.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.bench = bench
        self.lib = C.CDLL(path)
        self.lib.ref_align_gpu.restype = C.c_uint64
        self.lib.ref_align_gpu.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p,
                                           C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                           C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p]

    def _call(self, mode, alpha, matrix, gap, text, pattern, outT=None, outP=None):
        text, pattern = _u8(text), _u8(pattern)
        full = np.zeros(23 * 23, np.int32)
        full[:alpha * alpha] = _i32(matrix).ravel()[:alpha * alpha]
        score, ln, st, sp = C.c_int(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        r = self.lib.ref_align_gpu(mode, alpha, full.ctypes.data, gap, text.ctypes.data, len(text), pattern.ctypes.data,
                                   len(pattern), C.byref(score), C.byref(ln), C.byref(st), C.byref(sp),
                                   outT.ctypes.data if outT is not None else None, outP.ctypes.data if outP is not None else None)
        return r, score.value, ln.value, st.value, sp.value

    def fill_micros(self, mode, alpha, matrix, gap, text, pattern) -> int:
        assert self.bench
        return int(self._call(mode, alpha, matrix, gap, text, pattern)[0])

    def align(self, mode, alpha, matrix, gap, text, pattern) -> Alignment:
        assert not self.bench
        n = len(text)
        outT = np.empty(max(1, 2 * n), np.uint8); outP = np.empty(max(1, 2 * n), np.uint8)
        r, score, ln, st, sp = self._call(mode, alpha, matrix, gap, text, pattern, outT, outP)
        if r:
            raise RuntimeError(f"reference alignSequenceGPU failed ({r})")
        return Alignment(score, ln, st, sp, outT[:ln].tobytes(), outP[:ln].tobytes())

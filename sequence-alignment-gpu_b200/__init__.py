"""sequence-alignment-gpu_b200 -- host side of the B200-native pairwise alignment hot path.

The product is ``libsa_b200.so`` (hand-written sm_100a kernels behind the C ABI of
``include/sa_b200.h``).  This module is the Python mirror of the reference's
operator interface for that path:

* ``Request`` / ``Response`` / ``alignSequenceGPU`` follow
  ``/root/reference/SequenceAlignment.hpp:71-131`` (same field names, same
  ``0 = ok, 1 = error + MEM_ERROR on stdout`` convention, ``alignSequenceGPU.cu:541-546``);
* ``Aligner`` is the thin object over the C ABI (single pair, host batch,
  device-resident batch for torch tensors).

There is NO CPU fallback and nothing here imports ``oracle/``: if the CUDA
library is missing or no GPU is present the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from dataclasses import dataclass, field
from enum import IntEnum

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SA_B200_LIB") or os.path.join(_HERE, "libsa_b200.so")      # (SA_B200_LIB: dev aid, A/B builds)
CSRC = os.path.join(_HERE, "csrc")


# --- SequenceAlignment.hpp:10-22 -------------------------------------------------------------
class programArgs(IntEnum):
    CPU = 0
    GPU = 1
    DNA = 2
    PROTEIN = 3
    GLOBAL = 4
    LOCAL = 5
    SEMI_GLOBAL = 6
    SCORE_MATRIX = 7
    GAP_PENALTY = 8


NUM_DNA_CHARS = 4                                   # SequenceAlignment.hpp:52
NUM_PROTEIN_CHARS = 23                              # :53
DNA_ALPHABET = b"ATCG-"                             # :56
PROTEIN_ALPHABET = b"ARNDCQEGHILKMFPSTWYVBZX-"      # :57-58
DEFAULT_GAP_PENALTY = 5                             # :66
MEM_ERROR = "error: sequence is too long, not enough memory\n"   # :46

SA_GLOBAL, SA_LOCAL = 0, 1


class SaError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"sa_b200 error {status}: {msg}")
        self.status = status


class _Scoring(C.Structure):
    _fields_ = [("mode", C.c_int32), ("alphabet_size", C.c_int32), ("score_matrix", C.c_void_p),
                ("gap", C.c_int32), ("alphabet", C.c_char_p)]


class _Result(C.Structure):
    _fields_ = [("score", C.c_int32), ("aln_len", C.c_uint64), ("start_text", C.c_uint64),
                ("start_pattern", C.c_uint64)]


class _Timing(C.Structure):
    _fields_ = [("h2d_us", C.c_double), ("fill_us", C.c_double), ("traceback_us", C.c_double),
                ("d2h_us", C.c_double), ("total_us", C.c_double), ("cells", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


class _Batch(C.Structure):
    _fields_ = [("n_pairs", C.c_uint64), ("text", C.c_void_p), ("text_off", C.c_void_p),
                ("pattern", C.c_void_p), ("pattern_off", C.c_void_p)]


class _BatchOut(C.Structure):
    _fields_ = [("results", C.c_void_p), ("aln_off", C.c_void_p), ("aligned_text", C.c_void_p),
                ("aligned_pattern", C.c_void_p), ("arena_capacity", C.c_uint64), ("stats", C.c_void_p)]


class _Stats(C.Structure):
    _fields_ = [("identity", C.c_uint64), ("gaps", C.c_uint64)]


RESULT_DTYPE = np.dtype([("score", np.int32), ("_pad", np.int32), ("aln_len", np.uint64),
                         ("start_text", np.uint64), ("start_pattern", np.uint64)])
assert RESULT_DTYPE.itemsize == C.sizeof(_Result) == 32

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libsa_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "../libsa_b200.so"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libsa_b200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


def lib() -> C.CDLL:
    """Loads the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: run __graft_entry__.build() / make -C {CSRC}")
        L = C.CDLL(LIB_PATH)
        L.sa_version.restype = C.c_char_p
        L.sa_host_alloc.restype = C.c_void_p
        L.sa_host_alloc.argtypes = [C.c_uint64]
        L.sa_host_free.restype = None
        L.sa_host_free.argtypes = [C.c_void_p]
        L.sa_host_register.argtypes = [C.c_void_p, C.c_uint64]
        L.sa_host_unregister.argtypes = [C.c_void_p]
        L.sa_status_string.restype = C.c_char_p
        L.sa_status_string.argtypes = [C.c_int]
        L.sa_device_count.restype = C.c_int
        L.sa_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.sa_destroy.argtypes = [C.c_void_p]
        L.sa_last_timing.argtypes = [C.c_void_p, C.POINTER(_Timing)]
        L.sa_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_longlong]
        L.sa_get_option.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_longlong)]
        L.sa_last_cuda_error.argtypes = [C.c_void_p]
        L.sa_context_stream.restype = C.c_void_p
        L.sa_context_stream.argtypes = [C.c_void_p]
        L.sa_align.argtypes = [C.c_void_p, C.POINTER(_Scoring), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                               C.POINTER(_Result), C.c_void_p, C.c_void_p, C.c_uint64]
        L.sa_fill_only.argtypes = [C.c_void_p, C.POINTER(_Scoring), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
        L.sa_align_batch.argtypes = [C.c_void_p, C.POINTER(_Scoring), C.POINTER(_Batch), C.POINTER(_BatchOut)]
        L.sa_align_batch_device.argtypes = [C.c_void_p, C.POINTER(_Scoring), C.POINTER(_Batch), C.POINTER(_BatchOut),
                                            C.c_uint32, C.c_uint32, C.c_void_p]
        L.sa_align_device.argtypes = [C.c_void_p, C.POINTER(_Scoring), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sa_strip_fill.argtypes = [C.c_void_p, C.POINTER(_Scoring), C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
                                    C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sa_strip_begin.argtypes = [C.c_void_p, C.POINTER(_Scoring), C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
                                     C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p]
        L.sa_strip_fill_rows.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
        L.sa_peer_alloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.c_char_p]
        L.sa_peer_open.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.sa_peer_close.argtypes = [C.c_void_p, C.c_void_p]
        L.sa_peer_free.argtypes = [C.c_void_p, C.c_void_p]
        L.sa_strip_fill_linked.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.sa_strip_linked_status.argtypes = [C.c_void_p, C.c_void_p]
        L.sa_strip_traceback.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.sa_validate_and_transform.restype = C.c_int64
        L.sa_validate_and_transform.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p, C.c_int, C.c_char_p]
        L.sa_read_sequence_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_char_p]
        L.sa_parse_score_matrix_file.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
        L.sa_read_fasta_batch.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_uint64), C.c_char_p]
        L.sa_free.argtypes = [C.c_void_p]
        L.sa_pretty_print.restype = C.c_uint64
        L.sa_pretty_print.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.c_void_p,
                                      C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.sa_partition_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        _lib = L
    return _lib


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


def _alphabet_for(alpha: int) -> bytes:
    if alpha == NUM_PROTEIN_CHARS:
        return PROTEIN_ALPHABET
    if alpha == NUM_DNA_CHARS:
        return DNA_ALPHABET
    return bytes(range(65, 65 + alpha)) + b"-"


class Aligner:
    """One context = one GPU (one process per GPU in the multi-GPU runs)."""

    def __init__(self, device: int = 0):
        self._L = lib()
        self._ctx = C.c_void_p()
        rc = self._L.sa_create(device, C.byref(self._ctx))
        if rc != 0:
            raise SaError(rc, self._L.sa_status_string(rc).decode())
        self.device = device

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx:
            self._L.sa_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers
    def _scoring(self, mode, alpha, matrix, gap, alphabet=None):
        mat = np.ascontiguousarray(np.asarray(matrix, dtype=np.int32).ravel()[:alpha * alpha])
        alph = alphabet or _alphabet_for(alpha)
        sc = _Scoring(int(mode), int(alpha), mat.ctypes.data, int(gap), alph)
        sc._keep = (mat, alph)
        return sc

    def _check(self, rc):
        if rc != 0:
            raise SaError(rc, self._L.sa_status_string(rc).decode() +
                          f" (cuda error {self._L.sa_last_cuda_error(self._ctx)})")

    def set_option(self, name: str, value: int):
        """Tuning knob of this context (sa_set_option): e.g. dev_dirs_budget_mb, batch_min_chunks, ckpt_rows."""
        self._check(self._L.sa_set_option(self._ctx, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        v = C.c_longlong()
        self._check(self._L.sa_get_option(self._ctx, name.encode(), C.byref(v)))
        return int(v.value)

    def timing(self) -> dict:
        t = _Timing()
        self._L.sa_last_timing(self._ctx, C.byref(t))
        return {k: getattr(t, k) for k, _ in _Timing._fields_}

    # -- single pair (sa_align)
    def align(self, mode, alpha, matrix, gap, text, pattern, alphabet=None) -> Alignment:
        text = np.ascontiguousarray(np.asarray(text, dtype=np.uint8))
        pattern = np.ascontiguousarray(np.asarray(pattern, dtype=np.uint8))
        n, m = len(text), len(pattern)
        sc = self._scoring(mode, alpha, matrix, gap, alphabet)
        outT = np.empty(max(1, n + m), np.uint8)
        outP = np.empty(max(1, n + m), np.uint8)
        res = _Result()
        self._check(self._L.sa_align(self._ctx, C.byref(sc), text.ctypes.data, n, pattern.ctypes.data, m,
                                     C.byref(res), outT.ctypes.data, outP.ctypes.data, n + m))
        return Alignment(res.score, res.aln_len, res.start_text, res.start_pattern,
                         outT[:res.aln_len].tobytes(), outP[:res.aln_len].tobytes())

    def fill_only(self, mode, alpha, matrix, gap, text, pattern):
        text = np.ascontiguousarray(np.asarray(text, dtype=np.uint8))
        pattern = np.ascontiguousarray(np.asarray(pattern, dtype=np.uint8))
        sc = self._scoring(mode, alpha, matrix, gap)
        score, arg = C.c_int32(), C.c_uint64()
        self._check(self._L.sa_fill_only(self._ctx, C.byref(sc), text.ctypes.data, len(text), pattern.ctypes.data,
                                         len(pattern), C.byref(score), C.byref(arg)))
        return score.value, arg.value

    # -- device-resident single pair (sa_align_device)
    def align_device(self, mode, alpha, matrix, gap, d_text, n, d_pattern, m, d_out_text, d_out_pattern, d_result4,
                     stream=0, alphabet=None):
        sc = self._scoring(mode, alpha, matrix, gap, alphabet)
        self._check(self._L.sa_align_device(self._ctx, C.byref(sc), d_text, n, d_pattern, m, d_out_text,
                                            d_out_pattern, d_result4, C.c_void_p(stream)))

    # -- column slice of one global alignment (sa_strip_fill / sa_strip_traceback); raw device pointers
    def strip_fill(self, alpha, matrix, gap, d_text, n, col0, n_total, d_pattern, m, d_left_col, d_right_col, d_score=0,
                   stream=0, alphabet=None):
        sc = self._scoring(0, alpha, matrix, gap, alphabet)
        self._check(self._L.sa_strip_fill(self._ctx, C.byref(sc), d_text, n, col0, n_total, d_pattern, m,
                                          C.c_void_p(d_left_col or None), C.c_void_p(d_right_col or None),
                                          C.c_void_p(d_score or None), C.c_void_p(stream)))

    def strip_begin(self, alpha, matrix, gap, d_text, n, col0, n_total, d_pattern, m, chunk_rows_hint=0, stream=0, alphabet=None):
        """Plan a slice for row-chunked filling; returns the chunk height to use (multiple of the strip height)."""
        sc = self._scoring(0, alpha, matrix, gap, alphabet)
        out = C.c_uint64(0)
        self._check(self._L.sa_strip_begin(self._ctx, C.byref(sc), d_text, n, col0, n_total, d_pattern, m,
                                           int(chunk_rows_hint), C.byref(out), C.c_void_p(stream)))
        return int(out.value)

    def strip_fill_rows(self, row0, rows, d_left_col, d_right_col, d_top_row, d_bottom_row, d_score=0, stream=0):
        self._check(self._L.sa_strip_fill_rows(self._ctx, int(row0), int(rows), C.c_void_p(d_left_col or None),
                                               C.c_void_p(d_right_col or None), C.c_void_p(d_top_row or None),
                                               C.c_void_p(d_bottom_row or None), C.c_void_p(d_score or None),
                                               C.c_void_p(stream)))

    # -- slices linked inside the launch (peer border buffers through CUDA IPC)
    def peer_alloc(self, nbytes):
        """-> (device pointer, 64-byte IPC handle) of a zeroed border buffer in this GPU's memory."""
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        self._check(self._L.sa_peer_alloc(self._ctx, int(nbytes), C.byref(ptr), handle))
        return ptr.value, handle.raw

    def peer_open(self, handle: bytes):
        ptr = C.c_void_p()
        self._check(self._L.sa_peer_open(self._ctx, handle, C.byref(ptr)))
        return ptr.value

    def peer_close(self, ptr):
        self._check(self._L.sa_peer_close(self._ctx, C.c_void_p(ptr)))

    def peer_free(self, ptr):
        self._check(self._L.sa_peer_free(self._ctx, C.c_void_p(ptr)))

    def strip_fill_linked(self, d_left_col64, d_right_col64, tag, d_score=0, stream=0):
        self._check(self._L.sa_strip_fill_linked(self._ctx, C.c_void_p(d_left_col64 or None), C.c_void_p(d_right_col64 or None),
                                                 int(tag), C.c_void_p(d_score or None), C.c_void_p(stream)))

    def strip_linked_status(self, stream=0):
        self._check(self._L.sa_strip_linked_status(self._ctx, C.c_void_p(stream)))

    def strip_traceback(self, start_row, d_out_text, d_out_pattern, cap, d_result4, stream=0):
        self._check(self._L.sa_strip_traceback(self._ctx, int(start_row), d_out_text, d_out_pattern, int(cap), d_result4,
                                               C.c_void_p(stream)))

    # -- host batch (sa_align_batch); CSR numpy arrays in, numpy arrays out
    def align_batch(self, mode, alpha, matrix, gap, text, text_off, pattern, pattern_off, out=None, alphabet=None):
        text = np.ascontiguousarray(text, dtype=np.uint8)
        pattern = np.ascontiguousarray(pattern, dtype=np.uint8)
        text_off = np.ascontiguousarray(text_off, dtype=np.int64)
        pattern_off = np.ascontiguousarray(pattern_off, dtype=np.int64)
        N = len(text_off) - 1
        arena = int(text_off[N] - text_off[0] + pattern_off[N] - pattern_off[0])
        if out is None:
            out = dict(results=np.zeros(N, RESULT_DTYPE), aln_off=np.zeros(N, np.uint64),
                       aligned_text=np.empty(max(arena, 1), np.uint8), aligned_pattern=np.empty(max(arena, 1), np.uint8))
        sc = self._scoring(mode, alpha, matrix, gap, alphabet)
        b = _Batch(N, text.ctypes.data, text_off.ctypes.data, pattern.ctypes.data, pattern_off.ctypes.data)
        o = _BatchOut(out["results"].ctypes.data, out["aln_off"].ctypes.data, out["aligned_text"].ctypes.data,
                      out["aligned_pattern"].ctypes.data, arena, out["stats"].ctypes.data if out.get("stats") is not None else None)
        self._check(self._L.sa_align_batch(self._ctx, C.byref(sc), C.byref(b), C.byref(o)))
        return out

    def last_stats(self):
        """(identity, gaps) of the last align() call, counted on the device during string emission (sa_last_stats)."""
        st = _Stats()
        self._check(self._L.sa_last_stats(self._ctx, C.byref(st)))
        return int(st.identity), int(st.gaps)

    # -- device-resident batch (sa_align_batch_device): raw device pointers, e.g. torch data_ptr()
    def align_batch_device(self, mode, alpha, matrix, gap, n_pairs, d_text, d_text_off, d_pattern, d_pattern_off,
                           d_results, d_aln_off, d_out_text, d_out_pattern, arena_capacity, max_text_len,
                           max_pattern_len, stream=0, alphabet=None, d_stats=0):
        sc = self._scoring(mode, alpha, matrix, gap, alphabet)
        b = _Batch(n_pairs, d_text, d_text_off, d_pattern, d_pattern_off)
        o = _BatchOut(d_results, d_aln_off, d_out_text, d_out_pattern, arena_capacity, d_stats or None)
        self._check(self._L.sa_align_batch_device(self._ctx, C.byref(sc), C.byref(b), C.byref(o),
                                                  int(max_text_len), int(max_pattern_len), C.c_void_p(stream)))


class _PinnedOwner:
    """Frees a sa_host_alloc block when the last numpy view of it goes away."""

    def __init__(self, nbytes):
        self.ptr = lib().sa_host_alloc(int(nbytes))
        if not self.ptr:
            raise SaError(-2, "pinned host allocation failed")

    def __del__(self):
        try:
            lib().sa_host_free(C.c_void_p(self.ptr))
        except Exception:
            pass


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """numpy array in page-locked host memory (sa_host_alloc): the buffers sa_align_batch copies at PCIe speed."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    owner = _PinnedOwner(max(1, n * dtype.itemsize))
    buf = (C.c_ubyte * max(1, n * dtype.itemsize)).from_address(owner.ptr)
    buf._owner = owner                       # array -> ctypes buffer -> owner: freed with the last view
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def pinned_copy(a) -> np.ndarray:
    a = np.asarray(a)
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


def unpack_batch(out, i) -> Alignment:
    r = out["results"][i]
    off, ln = int(out["aln_off"][i]), int(r["aln_len"])
    return Alignment(int(r["score"]), ln, int(r["start_text"]), int(r["start_pattern"]),
                     out["aligned_text"][off:off + ln].tobytes(), out["aligned_pattern"][off:off + ln].tobytes())


class _Options(C.Structure):
    _fields_ = [("n_devices", C.c_int32), ("devices", C.c_int32 * 8)]


def align_batch_multi(devices, mode, alpha, matrix, gap, text, text_off, pattern, pattern_off, out=None, alphabet=None):
    """sa_align_batch_multi: one host batch sharded over ``devices`` (CUDA ordinals) inside this process, one host thread
    and one cached context per device, no collective."""
    text = np.ascontiguousarray(text, dtype=np.uint8)
    pattern = np.ascontiguousarray(pattern, dtype=np.uint8)
    text_off = np.ascontiguousarray(text_off, dtype=np.int64)
    pattern_off = np.ascontiguousarray(pattern_off, dtype=np.int64)
    N = len(text_off) - 1
    arena = int(text_off[N] - text_off[0] + pattern_off[N] - pattern_off[0])
    if out is None:
        out = dict(results=np.zeros(N, RESULT_DTYPE), aln_off=np.zeros(N, np.uint64),
                   aligned_text=np.empty(max(arena, 1), np.uint8), aligned_pattern=np.empty(max(arena, 1), np.uint8))
    m = np.ascontiguousarray(np.asarray(matrix, np.int32).ravel())
    sc = _Scoring(int(mode), int(alpha), m.ctypes.data, int(gap), alphabet or _alphabet_for(alpha))
    opt = _Options(len(devices), (C.c_int32 * 8)(*list(devices)))
    b = _Batch(N, text.ctypes.data, text_off.ctypes.data, pattern.ctypes.data, pattern_off.ctypes.data)
    o = _BatchOut(out["results"].ctypes.data, out["aln_off"].ctypes.data, out["aligned_text"].ctypes.data,
                  out["aligned_pattern"].ctypes.data, arena, out["stats"].ctypes.data if out.get("stats") is not None else None)
    L = lib()
    rc = L.sa_align_batch_multi(C.byref(opt), C.byref(sc), C.byref(b), C.byref(o))
    if rc != 0:
        raise SaError(rc, L.sa_status_string(rc).decode())
    return out


def multi_timing(k: int) -> dict:
    """Timing of device k of the last align_batch_multi call."""
    t = _Timing()
    lib().sa_multi_last_timing(int(k), C.byref(t))
    return {f: getattr(t, f) for f, _ in _Timing._fields_}


def partition_batch(text_off, pattern_off, world: int) -> np.ndarray:
    """Cell-balanced contiguous split of a batch over ``world`` ranks (sa_partition_batch)."""
    text_off = np.ascontiguousarray(text_off, dtype=np.int64)
    pattern_off = np.ascontiguousarray(pattern_off, dtype=np.int64)
    first = np.zeros(world + 1, np.uint64)
    rc = lib().sa_partition_batch(text_off.ctypes.data, pattern_off.ctypes.data, len(text_off) - 1, world,
                                  first.ctypes.data)
    if rc != 0:
        raise SaError(rc, "partition failed")
    return first


# --- the reference's operator interface (SequenceAlignment.hpp:71-131) -------------------------
@dataclass
class Request:
    deviceType: programArgs = programArgs.GPU
    sequenceType: programArgs = programArgs.DNA
    alignmentType: programArgs = programArgs.GLOBAL
    textBytes: np.ndarray = None          # alphabet indices, uint8
    textNumBytes: int = 0
    patternBytes: np.ndarray = None
    patternNumBytes: int = 0
    alphabet: bytes = DNA_ALPHABET
    alphabetSize: int = NUM_DNA_CHARS
    scoreMatrix: np.ndarray = field(default_factory=lambda: np.zeros(NUM_PROTEIN_CHARS ** 2, np.int32))
    gapPenalty: int = DEFAULT_GAP_PENALTY


@dataclass
class Response:
    alignedTextBytes: bytes = b""
    alignedPatternBytes: bytes = b""
    numAlignmentBytes: int = 0
    startInAlignedText: int = 0
    startInAlignedPattern: int = 0
    score: int = 0


_default_aligner = None


def _aligner() -> Aligner:
    global _default_aligner
    if _default_aligner is None:
        _default_aligner = Aligner(0)      # the reference uses device 0, alignSequenceGPU.cu:476
    return _default_aligner


def alignSequenceGPU(request: Request, response: Response) -> int:
    """Drop-in for SequenceAlignment::alignSequenceGPU (alignSequenceGPU.cu:463): returns 0 on
    success, 1 after printing MEM_ERROR to stdout on failure; anything but GLOBAL/LOCAL is a
    silent no-op returning 0 (alignSequenceCPU.cpp:318-328)."""
    if request.alignmentType not in (programArgs.GLOBAL, programArgs.LOCAL):
        return 0
    mode = SA_GLOBAL if request.alignmentType == programArgs.GLOBAL else SA_LOCAL
    try:
        a = _aligner().align(mode, request.alphabetSize, request.scoreMatrix, request.gapPenalty,
                             request.textBytes[:request.textNumBytes], request.patternBytes[:request.patternNumBytes],
                             alphabet=request.alphabet)
    except SaError as e:
        if e.status in (-2, -7):
            sys.stdout.write(MEM_ERROR)
        else:
            sys.stdout.write("error: could not copy from device memory\n")
        return 1
    response.alignedTextBytes = a.aligned_text
    response.alignedPatternBytes = a.aligned_pattern
    response.numAlignmentBytes = a.aln_len
    response.startInAlignedText = a.start_text
    response.startInAlignedPattern = a.start_pattern
    response.score = a.score
    return 0


# ---- front end (utilities.cpp:31-129), through the C ABI -------------------------------------------------------
def validateAndTransform(data: bytes, alphabet: bytes, alphabetSize: int) -> np.ndarray:
    """validateAndTransform (utilities.cpp:31-63): FASTA headers skipped, case folded, non-letters dropped, residues
    -> alphabet indices.  Like the reference, a letter outside the alphabet yields an EMPTY result (it returns 0)
    after naming the letter on stderr."""
    buf = C.create_string_buffer(bytes(data), len(data))
    bad = C.create_string_buffer(1)
    k = lib().sa_validate_and_transform(buf, len(data), alphabet, alphabetSize, bad)
    if k == 0 and bad.raw != b"\x00":
        sys.stderr.write(f"'{bad.raw.decode('latin1')}' letter not in alphabet.\n")
    return np.frombuffer(buf.raw[:k], np.uint8).copy()


def readSequenceFile(fname: str, request: Request) -> int:
    """readSequenceFile (utilities.cpp:65-104): the first call fills the text, the second the pattern; returns 0,
    or -1 when the file does not exist."""
    if not os.path.isfile(fname):
        sys.stderr.write(f"{fname} file does not exist\n")
        return -1
    seq = validateAndTransform(open(fname, "rb").read(), request.alphabet, request.alphabetSize)
    if request.textNumBytes == 0 and len(seq) > 0:
        request.textBytes, request.textNumBytes = seq, len(seq)
    elif request.patternNumBytes == 0 and len(seq) > 0:
        request.patternBytes, request.patternNumBytes = seq, len(seq)
    return 0


def parseScoreMatrixFile(fname: str, alphabetSize: int, buffer: np.ndarray) -> int:
    """parseScoreMatrixFile (utilities.cpp:106-129) including its return values: -1 for a malformed file, and 0 --
    with a message on stderr and the buffer untouched -- for a MISSING one."""
    if not os.path.isfile(fname):
        sys.stderr.write(f"{fname} file does not exist\n")
        return 0
    tmp = np.zeros(alphabetSize * alphabetSize, np.int32)
    if lib().sa_parse_score_matrix_file(fname.encode(), alphabetSize, tmp.ctypes.data) != 0:
        return -1
    buffer[:alphabetSize * alphabetSize] = tmp
    return 0


def read_fasta_batch(fname: str, alphabet: bytes, alphabetSize: int):
    """Multi-record FASTA -> (residues uint8, offsets int64) in the CSR form align_batch takes (new surface)."""
    res, off, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
    bad = C.create_string_buffer(1)
    rc = lib().sa_read_fasta_batch(fname.encode(), alphabet, alphabetSize, C.byref(res), C.byref(off), C.byref(n), bad)
    if rc != 0:
        raise SaError(rc, f"cannot read {fname}" + (f": letter '{bad.raw.decode('latin1')}' not in alphabet" if bad.raw != b"\x00" else ""))
    try:
        offsets = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_int64)), shape=(n.value + 1,)).copy()
        total = int(offsets[-1])
        residues = (np.ctypeslib.as_array(C.cast(res, C.POINTER(C.c_uint8)), shape=(max(total, 1),))[:total]).copy()
    finally:
        lib().sa_free(res); lib().sa_free(off)
    return residues, offsets


def prettyAlignmentPrint(response: Response) -> bytes:
    """prettyAlignmentPrint (utilities.cpp:253-315): the report the reference's driver prints, byte for byte."""
    L = lib()
    args = (response.alignedTextBytes, response.alignedPatternBytes, response.numAlignmentBytes,
            response.startInAlignedText, response.startInAlignedPattern, response.score)
    need = L.sa_pretty_print(*args, None, 0, None, None)
    buf = C.create_string_buffer(int(need) + 1)
    L.sa_pretty_print(*args, buf, need, None, None)
    return buf.raw[:need]


def alignment_stats(aligned_text: bytes, aligned_pattern: bytes):
    """(identical columns, gap columns) as prettyAlignmentPrint counts them."""
    ident, gaps = C.c_uint64(), C.c_uint64()
    lib().sa_pretty_print(aligned_text, aligned_pattern, len(aligned_text), 0, 0, 0, None, 0, C.byref(ident), C.byref(gaps))
    return ident.value, gaps.value

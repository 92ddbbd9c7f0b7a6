"""One GLOBAL alignment split into column slices over several GPUs (BASELINE config 5).

Rank k owns the text columns [k*W, (k+1)*W), W = ceil(n / world), all pattern rows, and the packed
direction words of that slice (a 1 M x 1 M pair needs 250 GB of directions: 125 GB per GPU at N = 2).

  fill       rank k receives 4*H(., col0) -- the right-most column of rank k-1, m+1 int32 -- fills its
             slice (sa_strip_fill) and sends its own right-most column to rank k+1.
  traceback  walks right to left: the last rank starts at row m on its right edge, follows the path to
             its left edge (sa_strip_traceback) and hands the row to the rank on its left.
  result     the aligned strings are the concatenation of the pieces in rank order.

Two ways of handing the border column over:

  linked (align_pair_strips_linked, the default of bench_c5.py): every rank launches its whole slice at once; a
      strip of slice k+1 starts when the same strip of slice k has written its right-most column into rank k+1's
      border buffer -- plain 8-byte {4H, tag} stores over NVLink into memory mapped with CUDA IPC, polled locally.
      Neighbouring GPUs overlap: 1 M x 951 k takes 0.325 s on 2 GPUs, 0.278 s on 4.
  sequential (align_pair_strips): torch.distributed send/recv of the whole column (NCCL between GPUs, gloo in the
      CPU tests); a rank starts when its left neighbour has finished (0.44 s).  The row-chunk variant of it
      (chunks > 1, sa_strip_fill_rows) is exact and tested but slower with the present kernel: every launch sweeps
      the whole slice width serially, so K chunks cost K sweeps.

`engine` is anything with fill(left_col) -> right_col, score(), traceback(start_row): GpuStripEngine
below for the product path; the CPU tests plug in a numpy restatement.
"""
from __future__ import annotations

import numpy as np


def slice_columns(n: int, world: int):
    """[(col0, width)] of every rank; trailing ranks may get an empty slice when n < world."""
    W = (n + world - 1) // world
    out = []
    for k in range(world):
        c0 = min(k * W, n)
        out.append((c0, min(W, n - c0)))
    return out


class GpuStripEngine:
    """One column slice on one GPU through the C ABI (torch only provides device memory and the stream)."""

    def __init__(self, aligner, alpha, matrix, gap, text_slice, col0, n_total, pattern, device="cuda:0", alphabet=None):
        import torch
        self.torch = torch
        self.al, self.alpha, self.matrix, self.gap, self.alphabet = aligner, alpha, matrix, gap, alphabet
        self.dev = torch.device(device)
        self.n, self.m, self.col0, self.n_total = len(text_slice), len(pattern), int(col0), int(n_total)
        self.d_text = torch.from_numpy(np.ascontiguousarray(text_slice, dtype=np.uint8)).to(self.dev)
        self.d_pat = torch.from_numpy(np.ascontiguousarray(pattern, dtype=np.uint8)).to(self.dev)
        self.d_score = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.fill_ms = None

    def column_buffer(self):
        return self.torch.empty(self.m + 1, dtype=self.torch.int32, device=self.dev)

    def fill(self, left_col):
        torch = self.torch
        right = self.column_buffer()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.stream):
            e0.record()
            self.al.strip_fill(self.alpha, self.matrix, self.gap, self.d_text.data_ptr(), self.n, self.col0,
                               self.n_total, self.d_pat.data_ptr(), self.m, left_col.data_ptr() if left_col is not None else 0,
                               right.data_ptr(), self.d_score.data_ptr(), stream=self.stream.cuda_stream,
                               alphabet=self.alphabet)
            e1.record()
        self.stream.synchronize()
        self.fill_ms = e0.elapsed_time(e1)
        self._keep = left_col
        return right

    # -- row-chunked filling (the multi-GPU pipeline): no host synchronisation, everything is ordered on self.stream
    def begin(self, chunk_rows_hint):
        self.chunk = self.al.strip_begin(self.alpha, self.matrix, self.gap, self.d_text.data_ptr(), self.n, self.col0,
                                         self.n_total, self.d_pat.data_ptr(), self.m, chunk_rows_hint,
                                         stream=self.stream.cuda_stream, alphabet=self.alphabet)
        return self.chunk

    def row_buffer(self):
        return self.torch.empty(max(self.n, 1), dtype=self.torch.int32, device=self.dev)

    def fill_rows(self, row0, rows, left_col, right_col, top_row, bottom_row):
        self.al.strip_fill_rows(row0, rows, left_col.data_ptr() if left_col is not None else 0, right_col.data_ptr(),
                                top_row.data_ptr() if top_row is not None else 0,
                                bottom_row.data_ptr() if bottom_row is not None else 0, self.d_score.data_ptr(),
                                stream=self.stream.cuda_stream)

    # -- slices linked inside the launch: border buffers in peer memory (CUDA IPC), no host hand-off at all
    def linked_setup(self):
        """Allocates this slice's border buffer ({4H, tag} words); returns its IPC handle for the LEFT neighbour."""
        if not hasattr(self, "border_ptr"):
            self.border_ptr, self.border_handle = self.al.peer_alloc(8 * (self.m + 1))
            self.right_ptr = 0
        return self.border_handle

    def linked_connect(self, right_handle=None, right_ptr=None):
        """right_handle: IPC handle of the right neighbour's border buffer (another process); right_ptr: its device
        pointer when the neighbour lives in this process (tests on one GPU)."""
        self.right_ptr = self.al.peer_open(right_handle) if right_handle is not None else (right_ptr or 0)
        self._right_mapped = right_handle is not None

    def fill_linked(self, tag):
        torch = self.torch
        self.begin(0)
        self._e0, self._e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            self._e0.record()
            self.al.strip_fill_linked(self.border_ptr if self.col0 > 0 else 0, self.right_ptr, tag, self.d_score.data_ptr(),
                                      stream=self.stream.cuda_stream)
            self._e1.record()

    def linked_finish(self):
        self.al.strip_linked_status(stream=self.stream.cuda_stream)       # synchronises; raises if a neighbour never delivered
        self.fill_ms = self._e0.elapsed_time(self._e1)

    def linked_release(self, barrier=None):
        """Unmaps the right neighbour's border buffer, then (after `barrier()`, so that nobody still has it mapped)
        frees this slice's own."""
        if getattr(self, "_right_mapped", False):
            self.al.peer_close(self.right_ptr)
            self._right_mapped = False
        self.right_ptr = 0
        if barrier is not None:
            barrier()
        if hasattr(self, "border_ptr"):
            self.al.peer_free(self.border_ptr)
            del self.border_ptr

    def score(self):
        return int(self.d_score.item())

    def traceback(self, start_row):
        torch = self.torch
        cap = self.n + self.m
        oT = torch.empty(cap, dtype=torch.uint8, device=self.dev)
        oP = torch.empty(cap, dtype=torch.uint8, device=self.dev)
        res = torch.zeros(4, dtype=torch.int64, device=self.dev)
        with torch.cuda.stream(self.stream):
            self.al.strip_traceback(start_row, oT.data_ptr(), oP.data_ptr(), cap, res.data_ptr(),
                                    stream=self.stream.cuda_stream)
        self.stream.synchronize()
        ln, exit_row, ti, pi = (int(x) for x in res.tolist())
        return (oT[cap - ln:].cpu().numpy().tobytes(), oP[cap - ln:].cpu().numpy().tobytes(), exit_row, ti, pi)


def align_pair_strips_local(engines, m, chunks: int = 1):
    """All slices in ONE process (any number of slices on one GPU): the same hand-offs without a network.
    chunks > 1 fills every slice in row chunks (the kernel path of the multi-GPU pipeline)."""
    col = None
    for e in engines:
        if e.n == 0:
            continue
        if chunks <= 1:
            col = e.fill(col)
            continue
        chunk = e.begin((m + chunks - 1) // chunks)
        right, bufs, top, row0, c = e.column_buffer(), [e.row_buffer(), e.row_buffer()], None, 0, 0
        while row0 < m:
            rows = min(chunk, m - row0)
            bottom = None if row0 + rows == m else bufs[c & 1]
            e.fill_rows(row0, rows, col, right, top, bottom)
            top, row0, c = bottom, row0 + rows, c + 1
        if hasattr(e, "stream"):
            e.stream.synchronize()
        e._keep = (col, bufs)
        col = right
    live = [e for e in engines if e.n > 0]
    score = live[-1].score()
    row, pieces, ti, pi = m, [], 0, 0
    for e in reversed(live):
        t, p, row, ti, pi = e.traceback(row)
        pieces.append((t, p))
    pieces.reverse()
    return score, b"".join(t for t, _ in pieces), b"".join(p for _, p in pieces), ti, pi


def fill_slice_pipelined(engine, m, rank: int, world: int, make_column, chunks: int, group=None):
    """Fill this rank's slice in `chunks` row chunks: a chunk starts when that part of the left neighbour's
    right-most column has arrived, and its own part of the right-most column is sent on at once, so that rank
    k works on chunk c while rank k+1 works on chunk c-1.  Every operation is stream-ordered (engine.stream is
    made current for the GPU engine); the host never waits inside the loop.  Returns the right-most column."""
    import contextlib
    import torch.distributed as dist
    ctx = engine.torch.cuda.stream(engine.stream) if hasattr(engine, "stream") else contextlib.nullcontext()
    with ctx:
        chunk = engine.begin((m + chunks - 1) // chunks)
        left = make_column() if rank > 0 else None
        right = make_column()
        rows_bufs = [engine.row_buffer(), engine.row_buffer()]
        top, row0, c = None, 0, 0
        while row0 < m:
            rows = min(chunk, m - row0)
            lo, hi = (row0 + 1 if c else 0), row0 + rows + 1           # column entries new in this chunk
            if rank > 0:
                dist.recv(left[lo:hi], src=rank - 1, group=group)
            last = row0 + rows == m
            bottom = None if last else rows_bufs[c & 1]
            engine.fill_rows(row0, rows, left, right, top, bottom)
            if rank + 1 < world:
                dist.send(right[lo:hi], dst=rank + 1, group=group)
            top, row0, c = bottom, row0 + rows, c + 1
    engine._keep = (left, right, rows_bufs)
    return right


_linked_calls = [0]


def align_pair_strips_linked(engine, m, rank: int, world: int, group=None):
    """One slice per rank, linked inside the launches (GpuStripEngine only): every rank launches its whole slice at
    once; strip s of rank k+1 starts when strip s of rank k has written its right-most column into rank k+1's border
    buffer over NVLink.  Returns like align_pair_strips."""
    import torch
    import torch.distributed as dist
    if not hasattr(engine, "right_ptr"):
        handles = [None] * world
        dist.all_gather_object(handles, engine.linked_setup(), group=group)
        engine.linked_connect(right_handle=handles[rank + 1] if rank + 1 < world else None)
    _linked_calls[0] += 1
    dist.barrier(group=group)                 # every border buffer exists and is mapped before anyone writes
    engine.fill_linked(_linked_calls[0])
    engine.linked_finish()
    row_t = torch.zeros(1, dtype=torch.int64, device=engine.dev)
    if rank + 1 < world:
        dist.recv(row_t, src=rank + 1, group=group)
        row = int(row_t.item())
    else:
        row = m
    t, p, row, ti, pi = engine.traceback(row)
    score = engine.score()
    if rank > 0:
        row_t.fill_(row)
        dist.send(row_t, dst=rank - 1, group=group)
    parts = [None] * world
    dist.all_gather_object(parts, (t, p, ti, pi, score), group=group)
    return (parts[-1][4], b"".join(x[0] for x in parts), b"".join(x[1] for x in parts), parts[0][2], parts[0][3])


def align_pair_strips_linked_local(engines, m, tag=1):
    """The linked protocol with all slices in ONE process on one GPU (tests): the kernels cannot run side by side
    there, so they are launched left to right, each finding its border already delivered."""
    live = [e for e in engines if e.n > 0]
    for e in live:
        e.linked_setup()
    for e, nxt in zip(live, live[1:] + [None]):
        e.linked_connect(right_ptr=nxt.border_ptr if nxt is not None else 0)
    for e in live:
        e.fill_linked(tag)
        e.linked_finish()
    score = live[-1].score()
    row, pieces, ti, pi = m, [], 0, 0
    for e in reversed(live):
        t, p, row, ti, pi = e.traceback(row)
        pieces.append((t, p))
    pieces.reverse()
    return score, b"".join(t for t, _ in pieces), b"".join(p for _, p in pieces), ti, pi


def align_pair_strips(engine, m, rank: int, world: int, make_column, group=None, chunks: int = 1):
    """One slice per rank.  `make_column()` returns an empty (m+1) int32 tensor on the device the
    process group communicates on.  chunks > 1 overlaps the fills of neighbouring ranks (row chunks).
    Returns (score, aligned_text, aligned_pattern, text_idx, pattern_idx) on every rank."""
    import torch
    import torch.distributed as dist
    if chunks > 1 and min(w for _, w in slice_columns(engine.n_total, world)) > 0:
        right = fill_slice_pipelined(engine, m, rank, world, make_column, chunks, group)
    else:
        left = None
        if rank > 0:
            left = make_column()
            dist.recv(left, src=rank - 1, group=group)
        right = engine.fill(left) if engine.n > 0 else left
        if rank + 1 < world:
            dist.send(right, dst=rank + 1, group=group)
    # score of the whole pair: known on the last rank that owns columns
    row_t = torch.zeros(1, dtype=torch.int64, device=right.device)
    if rank + 1 < world:
        dist.recv(row_t, src=rank + 1, group=group)
        row = int(row_t.item())
    else:
        row = m
    if engine.n > 0:
        t, p, row, ti, pi = engine.traceback(row)
        score = engine.score()
    else:
        t, p, ti, pi, score = b"", b"", 0, 0, None
    if rank > 0:
        row_t.fill_(row)
        dist.send(row_t, dst=rank - 1, group=group)
    parts = [None] * world
    dist.all_gather_object(parts, (t, p, ti, pi, score), group=group)
    score = [x[4] for x in parts if x[4] is not None][-1]
    return (score, b"".join(x[0] for x in parts), b"".join(x[1] for x in parts), parts[0][2], parts[0][3])

#!/usr/bin/env python
"""benchmarks.py -- the reference's benchmark harness (tests/benchmarks.cu) on the B200 path.

Regenerates the tables the reference publishes under tests/benchmarkResults/*.out with the same
workload recipe and the same output format, so the numbers can be laid side by side:

  * data: random protein sequences `rand() % 22`, blosum50.txt, gap 5   (benchmarks.cu:21-42)
  * sizes are `numRows x numCols` = (pattern+1) x (text+1); MCUPS = rows*cols / microseconds (:85,165,179)
  * best of NUM_REPEATS = 5                                               (:5)

modes (benchmarks.cu:366-404):
  throughput  fill only, no allocation / H2D / traceback in the timed region  (benchmarkFillMatrixThroughput,
              the `#define BENCHMARK` return value of alignSequenceGPU.cu:613-626) -> sa_fill_only + sa_last_timing
  latency     one full call incl. copies and traceback                         (benchmarkEndToEndLatency)
  batch       N x 8192^2 alignments back to back                               (benchmarkEndToEndBatch)
  maxlength   the longest runs the reference reports (120000^2, 500000^2)      (benchmarkMaxLength), fill only

`--cpu` also times the unmodified reference CPU path (oracle/_ref, test infrastructure) for the sizes
it can hold in memory, like the reference harness does.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

NUM_REPEATS = 5
NW_SIZES = [(256, 256), (512, 512), (1024, 1024), (2048, 2048), (4096, 4096), (8192, 8192), (16384, 16384),
            (32768, 32768), (65536, 65536)]                       # benchmarks.cu:104-115
SW_SIZES = [(256, 32768), (512, 32768), (1024, 32768), (2048, 32768), (4096, 32768), (8192, 32768),
            (16384, 32768), (32768, 32768)]                       # :116-126
LAT_NW_SIZES = [(256, 256), (512, 512), (1024, 1024), (4096, 4096), (8192, 8192), (16384, 16384), (32768, 32768),
                (65536, 65536)]                                   # :193-203


def blosum50():
    import json
    return np.asarray(json.load(open(os.path.join(ROOT, "tests", "golden", "matrices.json")))["protein/blosum50.txt"], np.int32)


def dummy_request(rows, cols, rng):
    """fillDummyRequest (benchmarks.cu:21-42): random residues in 0..21."""
    return rng.integers(0, 22, cols - 1, dtype=np.uint8), rng.integers(0, 22, rows - 1, dtype=np.uint8)


def throughput(al, mode, sizes, cpu, rng, refgpu=None):
    name = "Global" if mode == 0 else "Local"
    print(f"\n{name} alignment benchmark:")
    mat = blosum50()
    ref = None
    if cpu:
        from oracle.oracle_py import Reference
        ref = Reference("O3")
    for rows, cols in sizes:
        print(f"-----  {rows} x {cols}  -----")
        t, p = dummy_request(rows, cols, rng)
        cpu_us = None
        if ref is not None and rows * cols <= (1 << 32):
            best = 1e30
            M = np.empty(rows * cols, np.int8)
            for _ in range(2):
                t0 = time.perf_counter()
                ref.fill(mode, 23, mat, 5, t, p, M)
                best = min(best, (time.perf_counter() - t0) * 1e6)
            cpu_us = max(1.0, best)
            print(f"CPU = {int(cpu_us / 1000)} ms\nMCUPS: {int(rows * cols / cpu_us)}\n")
        best = 1e30
        for _ in range(NUM_REPEATS):
            al.fill_only(mode, 23, mat, 5, t, p)
            best = min(best, al.timing()["fill_us"])
        gpu_us = max(1.0, best)
        print(f"GPU = {gpu_us / 1000:.3f} ms\nMCUPS: {int(rows * cols / gpu_us)}\n")
        if cpu_us:
            print(f"GPU Speedup = {cpu_us / gpu_us:.1f}")
        if refgpu is not None and rows * cols <= (1 << 32):
            # the reference's own kernels on this GPU, timed by its own BENCHMARK switch (fill + D2H of 1 B/cell)
            rbest = min(refgpu.fill_micros(mode, 23, mat, 5, t, p) for _ in range(3))
            rbest = max(1, rbest)
            print(f"reference GPU (alignSequenceGPU.cu, same device) = {rbest / 1000:.3f} ms\nMCUPS: {int(rows * cols / rbest)}")
            print(f"speed-up over the reference GPU path = {rbest / gpu_us:.1f}\n")


def latency(al, mode, sizes, rng):
    name = "Global" if mode == 0 else "Local"
    print(f"\n{name} alignment latency (end-to-end) benchmark:")
    mat = blosum50()
    for rows, cols in sizes:
        print(f"-----  {rows} x {cols}  -----")
        t, p = dummy_request(rows, cols, rng)
        best = 1e30
        for _ in range(NUM_REPEATS):
            t0 = time.perf_counter()
            al.align(mode, 23, mat, 5, t, p)
            best = min(best, time.perf_counter() - t0)
        print(f"GPU = {best * 1e3:.3f} ms")


def batch(al, sa, nb, rng):
    print(f"\nGlobal alignment batch ({nb}x) benchmark:")
    rows = cols = 8192
    print(f"-----  {rows} x {cols}  -----")
    mat = blosum50()
    reqs = [dummy_request(rows, cols, rng) for _ in range(nb)]
    al.align(0, 23, mat, 5, *reqs[0])        # warm-up, like the reference (:314-315)
    t0 = time.perf_counter()
    for t, p in reqs:
        al.align(0, 23, mat, 5, t, p)
    print(f"GPU = {(time.perf_counter() - t0) * 1e3:.1f} ms")
    # the same requests as ONE sa_align_batch call: the dispatcher runs several such pairs side by side
    T = np.concatenate([t for t, _ in reqs]); P = np.concatenate([p for _, p in reqs])
    toff = np.arange(nb + 1, dtype=np.int64) * (cols - 1); poff = np.arange(nb + 1, dtype=np.int64) * (rows - 1)
    al.align_batch(0, 23, mat, 5, T, toff, P, poff)
    t0 = time.perf_counter()
    al.align_batch(0, 23, mat, 5, T, toff, P, poff)
    print(f"GPU, one sa_align_batch call = {(time.perf_counter() - t0) * 1e3:.1f} ms")


def maxlength(al, mode, rng, sizes):
    name = "Global" if mode == 0 else "Local"
    print(f"\n{name} alignment benchmark:")
    mat = blosum50()
    for rows, cols in sizes:
        print(f"-----  {rows} x {cols}  -----")
        t, p = dummy_request(rows, cols, rng)
        al.fill_only(mode, 23, mat, 5, t, p)
        us = max(1.0, al.timing()["fill_us"])
        print(f"GPU = {us / 1000:.1f} ms\nMCUPS: {int(rows * cols / us)}\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="throughput", choices=["throughput", "latency", "batch", "maxlength", "all"])
    ap.add_argument("--cpu", action="store_true", help="also time the reference CPU fill (oracle/_ref)")
    ap.add_argument("--ref-gpu", action="store_true",
                    help="also time the reference's own GPU path on this device (oracle/_ref/libsa_refgpu_bench.so)")
    ap.add_argument("--max-size", type=int, default=65536)
    args = ap.parse_args()
    sa = load_package()
    import ctypes
    al = sa.Aligner(0)
    rng = np.random.default_rng(0)
    print("Benchmark on GPU: B200 (libsa_b200, %s)" % sa.lib().sa_version().decode())
    refgpu = None
    if args.ref_gpu:
        from oracle.oracle_py import ReferenceGpu
        refgpu = ReferenceGpu(bench=True)
    if args.mode in ("throughput", "all"):
        throughput(al, 0, [s for s in NW_SIZES if s[0] <= args.max_size], args.cpu, rng, refgpu)
        throughput(al, 1, [s for s in SW_SIZES if s[0] <= args.max_size], args.cpu, rng, refgpu)
    if args.mode in ("latency", "all"):
        latency(al, 0, [s for s in LAT_NW_SIZES if s[0] <= args.max_size], rng)
        latency(al, 1, [s for s in SW_SIZES if s[0] <= args.max_size], rng)
    if args.mode in ("batch", "all"):
        for nb in (1, 2, 4, 8, 16, 32):
            batch(al, sa, nb, rng)
    if args.mode in ("maxlength", "all"):
        maxlength(al, 1, rng, [(120000, 120000), (500000, 500000)])
    al.close()


if __name__ == "__main__":
    main()

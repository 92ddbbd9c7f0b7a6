"""The real drop-in: a C++ caller written against include/SequenceAlignment.hpp like the reference's mainDriver.cu,
linked against libsa_b200.so only (tests/cpp/dropin_main.cpp), compared byte for byte with what the UNMODIFIED reference
CLI printed for the same command lines (tests/golden/cli_goldens.json, made by tests/golden/make_cli_goldens.py from
/root/reference/mainDriver.cu with `-c`).  Also: alignSequenceGPUBatch from C++, the -DBENCHMARK build of the library,
and the multi-GPU dispatcher behind the batch entry."""
import json
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

import cli_cases
import helpers
from gpu_common import load_package

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sequence-alignment-gpu_b200")


def _build(tmp, lib="sa_b200"):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    load_package().lib()          # the library must have been built
    exe = os.path.join(tmp, "dropin_" + lib)
    r = subprocess.run(["g++", "-std=c++14", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp"),
                        "-o", exe, "-L" + PKG, "-l" + lib, "-Wl,-rpath," + PKG], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("dropin"))
    cli_cases.write_inputs(d)
    return d


@pytest.fixture(scope="module")
def exe(workdir):
    return _build(workdir)


def _goldens():
    return json.load(open(os.path.join(helpers.GOLD, "cli_goldens.json")))


def _run(exe, args, cwd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([exe] + args, cwd=cwd, capture_output=True, env=e)


def test_cli_error_paths_match_the_reference(exe, workdir):
    """parseArguments failures never reach the device: usage, missing / unreadable inputs, bad numbers -- stdout,
    stderr and exit code of the reference CLI, on any machine."""
    n = 0
    for g in _goldens():
        if g["returncode"] == 0:
            continue
        r = _run(exe, [a.replace("{dev}", "-g") for a in g["args"]], workdir)
        assert r.returncode == g["returncode"], g["name"]
        assert r.stdout.decode("latin1") == g["stdout"], g["name"]
        assert r.stderr.decode("latin1") == g["stderr"], g["name"]
        n += 1
    assert n >= 6


def test_library_exports_the_front_end_of_the_reference(exe):
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(PKG, "libsa_b200.so")], capture_output=True, text=True).stdout
    for sym in ("parseArguments", "readSequenceFile", "validateAndTransform", "parseScoreMatrixFile", "prettyAlignmentPrint",
                "indexOfLetter", "getScore", "alignSequenceGPUBatch", "sa_align_batch_multi"):
        assert sym in out, sym
    assert os.path.exists(os.path.join(PKG, "libsa_b200_bench.so"))


@pytest.mark.gpu
def test_cli_output_equals_the_reference_cli(exe, workdir):
    """./alignSequence -g ... through the drop-in == ./alignSequence -c ... of the reference, byte for byte."""
    for g in _goldens():
        r = _run(exe, [a.replace("{dev}", "-g") for a in g["args"]], workdir)
        assert r.returncode == g["returncode"], (g["name"], r.stderr[-500:])
        assert r.stdout.decode("latin1") == g["stdout"], g["name"]
        assert r.stderr.decode("latin1") == g["stderr"], g["name"]
    g = next(x for x in _goldens() if x["name"].startswith("C2"))
    r = _run(exe, ["--twice"] + [a.replace("{dev}", "-g") for a in g["args"]], workdir)       # a reused Response
    assert r.returncode == 0 and r.stdout.decode("latin1") == g["stdout"]


def _write_batch_case(path, mode, alpha, mat, gap, pairs):
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", mode, alpha, gap, len(pairs)))
        f.write(np.asarray(mat, np.int32).tobytes())
        f.write(np.asarray([[len(t), len(p)] for t, p in pairs], np.uint32).tobytes())
        for t, p in pairs:
            f.write(np.asarray(t, np.uint8).tobytes())
            f.write(np.asarray(p, np.uint8).tobytes())


def _check_batch_output(stdout, oracle, mode, alpha, mat, gap, pairs):
    lines = stdout.decode("latin1").split("\n")
    assert len(lines) == len(pairs) + 1 and lines[-1] == ""
    for (t, p), line in zip(pairs, lines):
        want = oracle.align(mode, alpha, mat, gap, t, p)
        f = line.split(" ")
        got = (int(f[0]), int(f[1]), int(f[2]), int(f[3]), f[4].encode("latin1"), f[5].encode("latin1"))
        assert got == (want.score, want.aln_len, want.start_text, want.start_pattern, want.aligned_text, want.aligned_pattern)


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [None, "0", "1"])
def test_cpp_batch_entry_vs_oracle(exe, workdir, oracle, devices):
    """SequenceAlignment::alignSequenceGPUBatch called from C++ (pinned staging, the multi-GPU dispatcher with one
    device): every field and both strings of every pair against the oracle, both modes, short and long members."""
    rng = np.random.default_rng(17)
    b62 = helpers.matrices()["protein/blosum62.txt"]
    blast = helpers.matrices()["dna/blast.txt"]
    env = {} if devices is None else {"SA_DEVICES": devices}
    for mode, alpha, mat, nmax, count in ((1, 23, b62, 330, 300), (0, 4, blast, 120, 60), (0, 23, b62, 2500, 5)):
        pairs = [helpers.random_case(rng, alpha, n_max=nmax, similar=bool(i % 3)) for i in range(count)]
        pairs = [(t, p) for t, p in pairs if len(t) and len(p)]
        case = os.path.join(workdir, "batch.bin")
        _write_batch_case(case, mode, alpha, mat, 5, pairs)
        r = _run(exe, ["--batch", case], workdir, env)
        assert r.returncode == 0, r.stderr[-500:]
        _check_batch_output(r.stdout, oracle, mode, alpha, mat, 5, pairs)


@pytest.mark.gpu
def test_cpp_batch_entry_two_gpus(exe, workdir, oracle):
    """The dispatcher behind alignSequenceGPUBatch sharding one call over two GPUs (SA_DEVICES=0,1)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(18)
    b62 = helpers.matrices()["protein/blosum62.txt"]
    pairs = [helpers.random_case(rng, 23, n_max=330, similar=True) for _ in range(1500)]
    case = os.path.join(workdir, "batch2.bin")
    _write_batch_case(case, 1, 23, b62, 5, pairs)
    r = _run(exe, ["--batch", case], workdir, {"SA_DEVICES": "0,1"})
    assert r.returncode == 0, r.stderr[-500:]
    _check_batch_output(r.stdout, oracle, 1, 23, b62, 5, pairs)


@pytest.mark.gpu
def test_benchmark_build_returns_microseconds(workdir, exe):
    """libsa_b200_bench.so (-DBENCHMARK): alignSequenceGPU returns the microseconds of fill + D2H instead of 0
    (alignSequenceGPU.cu:613-626), which is what tests/benchmarks.cu:171-175 divides the cell count by."""
    g = next(x for x in _goldens() if x["name"].startswith("C1"))
    args = ["--return"] + [a.replace("{dev}", "-g") for a in g["args"]]
    plain = _run(exe, args, workdir)
    assert plain.returncode == 0 and plain.stdout.split()[0] == b"0"
    bench = _run(_build(workdir, "sa_b200_bench"), args, workdir)
    assert bench.returncode == 0
    us, score = (int(x) for x in bench.stdout.split())
    assert 50 <= us < 1_000_000 and score == int(plain.stdout.split()[1]) == 15344


@pytest.mark.gpu
def test_batch_multi_dispatcher_python(oracle):
    """sa_align_batch_multi through ctypes: one device, and every device of the box; packed strings, absolute aln_off."""
    import torch
    import synth
    sa = load_package()
    T, toff, P, poff = synth.synthetic_batch(12000, seed=61)
    mat = helpers.matrices()["protein/blosum62.txt"]
    for devs in ([0], list(range(min(torch.cuda.device_count(), 8)))):
        out = sa.align_batch_multi(devs, 1, 23, mat, 5, T, toff, P, poff)
        assert oracle.check_batch(1, 23, mat, 5, T, toff, P, poff, out) == (0, -1), devs
        assert sa.multi_timing(0)["cells"] > 0

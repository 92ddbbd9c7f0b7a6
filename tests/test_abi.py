"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol
include/sa_b200.h declares (no compute calls without a GPU), and the host-side mirror of
the reference interface behaves."""
import ctypes
import os
import re

import numpy as np
import pytest

from gpu_common import ROOT, load_package


@pytest.fixture(scope="module")
def sa():
    mod = load_package()
    if not os.path.exists(mod.LIB_PATH):
        mod.build()
    return mod


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "sa_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sa_[a-z_0-9]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(sa):
    lib = ctypes.CDLL(sa.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sa_b200.h but not exported"


def test_cpp_shim_exports_reference_entry_point(sa):
    import subprocess
    out = subprocess.run(["nm", "-DC", sa.LIB_PATH], capture_output=True, text=True).stdout
    assert "SequenceAlignment::alignSequenceGPU(SequenceAlignment::Request const&, SequenceAlignment::Response*)" in out
    assert "SequenceAlignment::alignSequenceGPUBatch" in out


def test_no_gpu_fails_loudly(sa):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert sa.lib().sa_device_count() == 0
    with pytest.raises(sa.SaError) as e:
        sa.Aligner(0)
    assert e.value.status == -1          # SA_ERR_NO_DEVICE: no CPU fallback


def test_result_struct_layout(sa):
    assert sa.RESULT_DTYPE.itemsize == 32
    assert sa.RESULT_DTYPE.fields["aln_len"][1] == 8
    assert sa.RESULT_DTYPE.fields["start_pattern"][1] == 24


def test_partition_batch_balances_cells(sa):
    rng = np.random.default_rng(0)
    n = rng.integers(50, 400, 5000)
    m = rng.integers(50, 400, 5000)
    toff = np.concatenate(([0], np.cumsum(n)))
    poff = np.concatenate(([0], np.cumsum(m)))
    cells = (n + 1) * (m + 1)
    for world in (1, 2, 4, 8):
        first = sa.partition_batch(toff, poff, world)
        assert first[0] == 0 and first[-1] == 5000 and np.all(np.diff(first.astype(np.int64)) > 0)
        share = [cells[int(first[r]):int(first[r + 1])].sum() for r in range(world)]
        assert max(share) / (cells.sum() / world) < 1.01


def test_product_does_not_touch_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "sequence-alignment-gpu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "sa_oracle" not in src and "oracle_py" not in src and "libsa_ref" not in src, f


def test_reference_driver_links_against_the_library(sa, tmp_path):
    """INTEGRATION.md recipe: the reference's own mainDriver.cu, built with the one-line header change
    (unity include of alignSequenceGPU.cu removed), links against libsa_b200.so and its CPU path still
    answers the reference's golden case.  Needs the reference tree (build container only)."""
    import shutil
    import subprocess
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "mainDriver.cu")) or shutil.which("nvcc") is None:
        pytest.skip("reference tree / nvcc not available")
    for f in ("mainDriver.cu", "SequenceAlignment.hpp", "utilities.cpp", "alignSequenceCPU.cpp"):
        shutil.copy(os.path.join(ref, f), tmp_path / f)          # scratch copy, never committed
    hdr = (tmp_path / "SequenceAlignment.hpp").read_text()
    assert '#include "alignSequenceGPU.cu"' in hdr
    (tmp_path / "SequenceAlignment.hpp").write_text(hdr.replace('#include "alignSequenceGPU.cu"', ""))
    libdir = os.path.dirname(sa.LIB_PATH)
    exe = tmp_path / "alignSequence"
    r = subprocess.run(["nvcc", "-std=c++14", "-m64", "--expt-relaxed-constexpr", "-include", "cstdint", "-w",
                        str(tmp_path / "mainDriver.cu"), "-o", str(exe), f"-L{libdir}", "-lsa_b200",
                        f"-Xlinker=-rpath={libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([str(exe), "-c", "--global", "data/dna/dna_01.txt", "data/dna/dna_02.txt"], cwd=ref,
                         capture_output=True, text=True)
    assert out.returncode == 0 and "# Score: \t-4" in out.stdout, out.stdout + out.stderr

"""Helpers shared by the -m gpu tests: package loading and comparison against the oracle."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_package():
    """Imports sequence-alignment-gpu_b200/ (not a valid identifier) as module ``sa_b200``."""
    if "sa_b200" in sys.modules:
        return sys.modules["sa_b200"]
    pkg_dir = os.path.join(ROOT, "sequence-alignment-gpu_b200")
    spec = importlib.util.spec_from_file_location("sa_b200", os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["sa_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def assert_same(got, want, ctx=""):
    assert got.score == want.score, (ctx, "score", got.score, want.score)
    assert got.aln_len == want.aln_len, (ctx, "aln_len", got.aln_len, want.aln_len)
    assert got.start_text == want.start_text, (ctx, "start_text", got.start_text, want.start_text)
    assert got.start_pattern == want.start_pattern, (ctx, "start_pattern", got.start_pattern, want.start_pattern)
    assert got.aligned_text == want.aligned_text, (ctx, "aligned_text")
    assert got.aligned_pattern == want.aligned_pattern, (ctx, "aligned_pattern")

"""CPU tests: pin oracle/sa_oracle.c against the reference's golden vectors and,
when oracle/_ref is built, against the unmodified reference itself."""
import numpy as np
import pytest

import helpers


def test_golden_file_has_reference_kat():
    names = [g["name"] for g in helpers.goldens()]
    for k in ("tests.cu:DNA_01", "tests.cu:DNA_04", "tests.cu:PROTEIN_01", "tests.cu:LOCAL_DNA_01",
              "tests.cu:LOCAL_PROTEIN_01"):
        assert k in names
    assert len(names) > 300


@pytest.mark.parametrize("chunk", range(8))
def test_oracle_matches_goldens(oracle, chunk):
    gs = [g for g in helpers.goldens() if helpers.has_inputs(g)]
    # the two 27k x 26k cases cost ~10 s each on one core; keep one per mode in chunk 0
    for g in gs[chunk::8]:
        if g["n"] * g["m"] > 2e8 and chunk != 0 and not g["name"].startswith("long:"):
            continue
        t, p, mat = helpers.golden_inputs(g)
        aln = oracle.align(g["mode"], g["alpha"], mat, g["gap"], t, p)
        helpers.check_against_golden(aln, g)


@pytest.mark.parametrize("alpha,mname", [(4, "dna/blast.txt"), (23, "protein/blosum62.txt")])
@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_matches_reference_random(oracle, reference, alpha, mname, mode):
    rng = np.random.default_rng(100 * alpha + mode)
    mat = helpers.matrices()[mname]
    for it in range(150):
        t, p = helpers.random_case(rng, alpha, n_max=160, similar=bool(it % 2))
        gap = int(rng.integers(0, 12))
        a = oracle.align(mode, alpha, mat, gap, t, p)
        b = reference.align(mode, alpha, mat, gap, t, p)
        assert a.key() == b.key(), (alpha, mode, it, gap, t.tolist(), p.tolist())


def test_oracle_random_matrix_vs_reference(oracle, reference):
    rng = np.random.default_rng(7)
    for it in range(100):
        alpha = 4 if it % 2 else 23
        mat = rng.integers(-9, 12, (alpha, alpha)).astype(np.int32)   # asymmetric on purpose
        t, p = helpers.random_case(rng, alpha, n_max=120, similar=bool(it % 3))
        gap = int(rng.integers(1, 9))
        for mode in (0, 1):
            assert oracle.align(mode, alpha, mat, gap, t, p).key() == \
                reference.align(mode, alpha, mat, gap, t, p).key()


def test_oracle_direction_matrix_matches_reference_fill(oracle, reference):
    rng = np.random.default_rng(3)
    mat = helpers.matrices()["dna/blast.txt"]
    for mode in (0, 1):
        t, p = helpers.random_case(rng, 4, n_max=300)
        _, dirs = oracle.align(mode, 4, mat, 5, t, p, want_dirs=True)
        score, arg, M = reference.fill(mode, 4, mat, 5, t, p)
        assert np.array_equal(dirs.ravel(), M)


def test_score_only_and_rescore(oracle):
    rng = np.random.default_rng(11)
    mat = helpers.matrices()["protein/blosum62.txt"]
    for mode in (0, 1):
        t, p = helpers.random_case(rng, 23, n_max=400)
        a = oracle.align(mode, 23, mat, 5, t, p)
        s, _ = oracle.score_only(mode, 23, mat, 5, t, p)
        assert s == a.score
        assert oracle.rescore(a.aligned_text, a.aligned_pattern, 23, mat, 5) == a.score


def test_sw_zero_score_wraps_start_indices(oracle):
    mat = helpers.matrices()["dna/blast.txt"]
    a = oracle.align(1, 4, mat, 5, np.zeros(4, np.uint8), np.ones(3, np.uint8))
    assert (a.score, a.aln_len) == (0, 0)
    assert a.start_text == a.start_pattern == 2**64 - 1   # alignSequenceCPU.cpp:13-14,56-57


def _oracle_batch_out(oracle, mode, alpha, mat, gap, T, toff, P, poff):
    """A result set in sa_align_batch's layout, produced by the oracle pair by pair (packed strings)."""
    N = len(toff) - 1
    dt = np.dtype([("score", "<i4"), ("_pad", "<i4"), ("aln_len", "<u8"), ("start_text", "<u8"), ("start_pattern", "<u8")])
    res, aoff = np.zeros(N, dt), np.zeros(N, np.uint64)
    aT, aP, pos = bytearray(), bytearray(), 0
    for i in range(N):
        a = oracle.align(mode, alpha, mat, gap, T[toff[i]:toff[i + 1]], P[poff[i]:poff[i + 1]])
        res[i] = (a.score, 0, a.aln_len, a.start_text, a.start_pattern)
        aoff[i] = pos
        aT += a.aligned_text
        aP += a.aligned_pattern
        pos += a.aln_len
    return dict(results=res, aln_off=aoff, aligned_text=np.frombuffer(bytes(aT) + b"\0", np.uint8).copy(),
                aligned_pattern=np.frombuffer(bytes(aP) + b"\0", np.uint8).copy())


@pytest.mark.parametrize("mode", [0, 1])
def test_batch_checkers_accept_truth_and_catch_corruption(oracle, reference, mode):
    """oracle.check_batch / reference.check_batch (the multi-threaded checkers behind the GPU batch tests and bench.py's
    `verified`) agree with the pair-by-pair oracle and flag a single changed field or string byte."""
    rng = np.random.default_rng(31 + mode)
    mat = helpers.matrices()["protein/blosum62.txt"]
    texts, pats = [], []
    for i in range(60):
        t, p = helpers.random_case(rng, 23, n_max=120, similar=bool(i % 3))
        if len(p) > len(t):
            t, p = p, t
        texts.append(t)
        pats.append(p)
    toff = np.concatenate(([0], np.cumsum([len(t) for t in texts]))).astype(np.int64)
    poff = np.concatenate(([0], np.cumsum([len(p) for p in pats]))).astype(np.int64)
    T, P = np.concatenate(texts), np.concatenate(pats)
    out = _oracle_batch_out(oracle, mode, 23, mat, 5, T, toff, P, poff)
    for chk in (oracle, reference):
        assert chk.check_batch(mode, 23, mat, 5, T, toff, P, poff, out, nthreads=3) == (0, -1)
        assert chk.check_batch(mode, 23, mat, 5, T, toff, P, poff, out, idx=[5, 17, 59], nthreads=2) == (0, -1)
    bad = dict(out, results=out["results"].copy())
    bad["results"]["start_pattern"][17] += 1
    k = int(np.flatnonzero(out["results"]["aln_len"] > 0)[-1])
    txt = out["aligned_text"].copy()
    txt[int(out["aln_off"][k])] ^= 1
    for chk in (oracle, reference):
        assert chk.check_batch(mode, 23, mat, 5, T, toff, P, poff, bad, nthreads=3) == (1, 17)
        assert chk.check_batch(mode, 23, mat, 5, T, toff, P, poff, dict(out, aligned_text=txt), nthreads=3) == (1, k)

"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI
(libsa_b200.so via ctypes), against the oracle and the committed reference goldens.
Bit-exact: score, numAlignmentBytes, both start indices, both aligned strings."""
import os

import numpy as np
import pytest

import helpers
from gpu_common import assert_same, load_package

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sa():
    return load_package()


@pytest.fixture(scope="module")
def aligner(sa):
    a = sa.Aligner(0)
    yield a
    a.close()


@pytest.fixture()
def force_path(monkeypatch):
    def _set(path=None, R=None):
        for k, v in (("SA_FORCE_PATH", path), ("SA_LONG_R", R)):
            if v is None:
                monkeypatch.delenv(k, raising=False)
            else:
                monkeypatch.setenv(k, str(v))
    return _set


def test_library_reports_device(sa):
    assert sa.lib().sa_device_count() >= 1
    assert b"sm_100a" in sa.lib().sa_version()


def test_reference_known_answer_vectors(aligner):
    """The reference's own golden vectors (tests/tests.cu:116-368)."""
    n = 0
    for g in helpers.goldens():
        if not g["name"].startswith("tests.cu:"):
            continue
        t, p, mat = helpers.golden_inputs(g)
        a = aligner.align(g["mode"], g["alpha"], mat, g["gap"], t, p)
        helpers.check_against_golden(a, g)
        n += 1
    assert n >= 12


@pytest.mark.parametrize("chunk", range(4))
def test_all_goldens_default_routing(aligner, chunk):
    """Every committed reference output: data/ all-pairs sweep, C1/C2, matrices, synthetic, edges."""
    gs = [g for g in helpers.goldens() if helpers.has_inputs(g)]
    for g in gs[chunk::4]:
        t, p, mat = helpers.golden_inputs(g)
        a = aligner.align(g["mode"], g["alpha"], mat, g["gap"], t, p)
        helpers.check_against_golden(a, g)


@pytest.mark.parametrize("R", [2, 4, 6, 8, 12, 16])
def test_goldens_through_long_kernel(aligner, force_path, R):
    """Short and medium goldens forced through the persistent strip kernel at every R."""
    force_path("long", R)
    gs = [g for g in helpers.goldens() if helpers.has_inputs(g) and g["n"] * g["m"] < 3e6]
    for g in gs[R % 3::3]:
        t, p, mat = helpers.golden_inputs(g)
        a = aligner.align(g["mode"], g["alpha"], mat, g["gap"], t, p)
        helpers.check_against_golden(a, g)


@pytest.mark.parametrize("path", ["batch", "long"])
@pytest.mark.parametrize("alpha", [4, 23])
@pytest.mark.parametrize("mode", [0, 1])
def test_random_pairs_vs_oracle(aligner, oracle, force_path, path, alpha, mode):
    force_path(path)
    rng = np.random.default_rng(1000 + 10 * alpha + mode)
    for it in range(120):
        mat = rng.integers(-9, 12, (alpha, alpha)).astype(np.int32) if it % 3 == 0 else \
            helpers.matrices()["dna/blast.txt" if alpha == 4 else "protein/blosum62.txt"]
        nmax = [12, 70, 300, 1100][it % 4]
        t, p = helpers.random_case(rng, alpha, n_max=nmax, similar=bool(it % 2))
        gap = int(rng.integers(0, 12))
        got = aligner.align(mode, alpha, mat, gap, t, p)
        want = oracle.align(mode, alpha, mat, gap, t, p)
        assert_same(got, want, (path, alpha, mode, it, gap, len(t), len(p)))


def test_low_complexity_ties(aligner, oracle, force_path):
    """Homopolymers / short repeats: every cell is a tie, arg-max has many candidates."""
    mat = helpers.matrices()["dna/blast.txt"]
    for path in ("batch", "long"):
        force_path(path)
        for mode in (0, 1):
            for t, p in ((np.zeros(333, np.uint8), np.zeros(97, np.uint8)),
                         (np.tile(np.arange(4, dtype=np.uint8), 90), np.tile(np.arange(4, dtype=np.uint8), 33)),
                         (np.tile(np.array([0, 0, 1], np.uint8), 100), np.tile(np.array([0, 1], np.uint8), 60))):
                for gap in (0, 1, 5):
                    assert_same(aligner.align(mode, 4, mat, gap, t, p), oracle.align(mode, 4, mat, gap, t, p),
                                (path, mode, gap, len(t)))


def test_pattern_longer_than_text(aligner, oracle):
    """Direct API callers may pass pattern > text (the reference would overflow its 2*text buffers)."""
    rng = np.random.default_rng(5)
    mat = helpers.matrices()["dna/blast.txt"]
    t = rng.integers(0, 4, 50, dtype=np.uint8)
    p = rng.integers(0, 4, 400, dtype=np.uint8)
    for mode in (0, 1):
        assert_same(aligner.align(mode, 4, mat, 5, t, p), oracle.align(mode, 4, mat, 5, t, p))


def test_argument_errors(sa, aligner):
    mat = helpers.matrices()["dna/blast.txt"]
    with pytest.raises(sa.SaError):
        aligner.align(0, 4, mat, 5, np.zeros(0, np.uint8), np.zeros(3, np.uint8))
    with pytest.raises(sa.SaError):      # beyond the two-plane profile (|S| <= 4064); 1000 is fine, see test_wide_score_matrices
        aligner.align(0, 4, np.full(16, 5000, np.int32), 5, np.zeros(3, np.uint8), np.zeros(3, np.uint8))


def test_host_batch_vs_oracle(sa, aligner, oracle):
    rng = np.random.default_rng(77)
    mat = helpers.matrices()["protein/blosum62.txt"]
    for mode in (0, 1):
        texts, pats = [], []
        for i in range(700):
            t, p = helpers.random_case(rng, 23, n_max=[40, 180, 350][i % 3], similar=bool(i % 2))
            texts.append(t)
            pats.append(p)
        toff = np.concatenate(([0], np.cumsum([len(t) for t in texts]))).astype(np.int64)
        poff = np.concatenate(([0], np.cumsum([len(p) for p in pats]))).astype(np.int64)
        out = aligner.align_batch(mode, 23, mat, 5, np.concatenate(texts), toff, np.concatenate(pats), poff)
        for i in range(len(texts)):
            assert_same(sa.unpack_batch(out, i), oracle.align(mode, 23, mat, 5, texts[i], pats[i]), (mode, i))


def test_host_batch_with_long_members(sa, aligner, oracle):
    """A batch that mixes short pairs with pairs too long for the batch kernel."""
    rng = np.random.default_rng(78)
    mat = helpers.matrices()["dna/blast.txt"]
    texts, pats = [], []
    for nmax in (50, 3000, 200, 700, 9):
        t, p = helpers.random_case(rng, 4, n_max=nmax)
        texts.append(t)
        pats.append(p)
    toff = np.concatenate(([0], np.cumsum([len(t) for t in texts]))).astype(np.int64)
    poff = np.concatenate(([0], np.cumsum([len(p) for p in pats]))).astype(np.int64)
    for mode in (0, 1):
        out = aligner.align_batch(mode, 4, mat, 5, np.concatenate(texts), toff, np.concatenate(pats), poff)
        for i in range(len(texts)):
            assert_same(sa.unpack_batch(out, i), oracle.align(mode, 4, mat, 5, texts[i], pats[i]), (mode, i))


def test_device_batch_torch_tensors(sa, aligner, oracle):
    """sa_align_batch_device on torch tensors / torch's current stream (the bench's `value` path)."""
    import torch
    import synth
    T, toff, P, poff = synth.synthetic_batch(3000, seed=9, lo=250, hi=350)
    mat = helpers.matrices()["protein/blosum62.txt"]
    dev = torch.device("cuda:0")
    dT, dP = torch.from_numpy(T).to(dev), torch.from_numpy(P).to(dev)
    dto, dpo = torch.from_numpy(toff).to(dev), torch.from_numpy(poff).to(dev)
    N = len(toff) - 1
    arena = int(toff[-1] + poff[-1])
    res = torch.zeros(N * 4, dtype=torch.int64, device=dev)
    aoff = torch.zeros(N, dtype=torch.int64, device=dev)
    oT = torch.zeros(arena, dtype=torch.uint8, device=dev)
    oP = torch.zeros(arena, dtype=torch.uint8, device=dev)
    max_n = int((toff[1:] - toff[:-1]).max())
    max_m = int((poff[1:] - poff[:-1]).max())
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    for mode in (1, 0):
        aligner.align_batch_device(mode, 23, mat, 5, N, dT.data_ptr(), dto.data_ptr(), dP.data_ptr(), dpo.data_ptr(),
                                   res.data_ptr(), aoff.data_ptr(), oT.data_ptr(), oP.data_ptr(), arena, max_n, max_m,
                                   stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        out = dict(results=res.cpu().numpy().view(sa.RESULT_DTYPE), aln_off=aoff.cpu().numpy().astype(np.uint64),
                   aligned_text=oT.cpu().numpy(), aligned_pattern=oP.cpu().numpy())
        for i in range(0, N, 7):
            t, p = T[toff[i]:toff[i + 1]], P[poff[i]:poff[i + 1]]
            assert_same(sa.unpack_batch(out, i), oracle.align(mode, 23, mat, 5, t, p), (mode, i))


def test_reference_operator_interface(sa):
    """Request / Response / alignSequenceGPU mirror SequenceAlignment.hpp:71-131."""
    g = next(x for x in helpers.goldens() if x["name"] == "tests.cu:LOCAL_DNA_01")
    t, p, mat = helpers.golden_inputs(g)
    rq = sa.Request(deviceType=sa.programArgs.GPU, sequenceType=sa.programArgs.DNA, alignmentType=sa.programArgs.LOCAL,
                    textBytes=t, textNumBytes=len(t), patternBytes=p, patternNumBytes=len(p), alphabet=sa.DNA_ALPHABET,
                    alphabetSize=4, gapPenalty=5)
    rq.scoreMatrix[:16] = mat
    rs = sa.Response()
    assert sa.alignSequenceGPU(rq, rs) == 0
    assert (rs.score, rs.alignedTextBytes, rs.alignedPatternBytes) == (20, b"ACAC", b"ACAC")
    assert (rs.startInAlignedText, rs.startInAlignedPattern, rs.numAlignmentBytes) == (248, 0, 4)
    rq.alignmentType = sa.programArgs.SEMI_GLOBAL       # declared, never implemented: silent no-op
    assert sa.alignSequenceGPU(rq, sa.Response()) == 0


@pytest.mark.parametrize("variant", [dict(SA_TB="serial"), dict(SA_TB_WD="16"), dict(SA_TB_WD="200"), dict(SA_LONG_R="4", SA_TB_WD="32"), dict(SA_TB_BAND="1"), dict(SA_LONG_R="2")])
def test_long_traceback_variants(aligner, oracle, force_path, monkeypatch, variant):
    """Parallel traceback (walkers / resolve / segments) vs the serial device walk vs the oracle,
    with candidate spacings that force both the merged-bracket and the fallback paths."""
    force_path("long")
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(4242)
    blast = helpers.matrices()["dna/blast.txt"]
    b62 = helpers.matrices()["protein/blosum62.txt"]
    cases = []
    for it in range(10):
        alpha, mat = (4, blast) if it % 2 == 0 else (23, b62)
        t, p = helpers.random_case(rng, alpha, n_max=[400, 2500, 6000][it % 3], similar=it % 4 != 3)
        cases.append((alpha, mat, t, p))
    # dissimilar sequences: paths meander, brackets often disagree -> serial fallback segments
    cases.append((4, blast, rng.integers(0, 4, 3000, dtype=np.uint8), rng.integers(0, 4, 2500, dtype=np.uint8)))
    for alpha, mat, t, p in cases:
        for mode in (0, 1):
            for gap in (2, 7):
                assert_same(aligner.align(mode, alpha, mat, gap, t, p), oracle.align(mode, alpha, mat, gap, t, p),
                            (variant, alpha, mode, gap, len(t), len(p)))


def test_full_size_c3_known_answer_and_rescore(aligner, oracle):
    """BASELINE config 3 at full size (100 000 x 95 217 synthetic DNA, NW): the reference CPU's answer for
    this seeded pair (score 399463, 100254 columns, starts 0/0) plus size-independent
    properties: re-scoring the emitted alignment reproduces the score, the strings spell the inputs."""
    import synth
    t, p = synth.synthetic_pair(100000, 12345, 54321)
    assert (len(t), len(p)) == (100000, 95217)
    blast = helpers.matrices()["dna/blast.txt"]
    a = aligner.align(0, 4, blast, 5, t, p)
    assert (a.score, a.aln_len, a.start_text, a.start_pattern) == (399463, 100254, 0, 0)
    assert oracle.rescore(a.aligned_text, a.aligned_pattern, 4, blast, 5) == a.score
    letters = np.frombuffer(b"ATCG", np.uint8)
    assert a.aligned_text.replace(b"-", b"") == letters[t].tobytes()
    assert a.aligned_pattern.replace(b"-", b"") == letters[p].tobytes()
    # the same pair as a local alignment: properties only (the CPU reference needs 9.5 GB and ~100 s)
    b = aligner.align(1, 4, blast, 5, t, p)
    assert b.score >= a.score and oracle.rescore(b.aligned_text, b.aligned_pattern, 4, blast, 5) == b.score
    s0, s1 = b.start_text, b.start_pattern
    assert b.aligned_text.replace(b"-", b"") in letters[t].tobytes() and b.aligned_pattern.replace(b"-", b"") in letters[p].tobytes()
    assert s0 < len(t) and s1 < len(p)


def _strip_align(sa, alpha, mat, gap, t, p, world, chunks=1):
    """Column slices of one global alignment, all on cuda:0, one context per slice (strips.py)."""
    from sa_b200 import strips
    als = [sa.Aligner(0) for _ in range(world)]
    try:
        eng = [strips.GpuStripEngine(al, alpha, mat, gap, t[c0:c0 + w], c0, len(t), p)
               for al, (c0, w) in zip(als, strips.slice_columns(len(t), world))]
        score, at, ap, ti, pi = strips.align_pair_strips_local(eng, len(p), chunks=chunks)
    finally:
        for al in als:
            al.close()
    return sa.Alignment(score, len(at), ti, pi, at, ap)


@pytest.mark.parametrize("tb", ["parallel", "serial"])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_column_slices_vs_oracle(sa, oracle, monkeypatch, world, tb):
    """BASELINE config 5 code path (sa_strip_fill / sa_strip_traceback) at sizes the oracle can check:
    the concatenated pieces must equal the single-matrix reference result bit for bit."""
    if tb == "serial":
        monkeypatch.setenv("SA_TB", "serial")
    rng = np.random.default_rng(100 + world)
    blast = helpers.matrices()["dna/blast.txt"]
    b62 = helpers.matrices()["protein/blosum62.txt"]
    for alpha, mat, gap, n, m in ((4, blast, 5, 2500, 2300), (4, blast, 2, 700, 1900), (23, b62, 7, 1500, 1500),
                                  (4, blast, 5, 9, 4), (4, blast, 5, 3000, 37)):
        import synth
        t = rng.integers(0, alpha, n, dtype=np.uint8)
        base = synth.mutate_indices_numpy(t, rng, alpha)
        p = base[:m] if m <= len(base) else np.concatenate((base, rng.integers(0, alpha, m - len(base), dtype=np.uint8)))
        assert_same(_strip_align(sa, alpha, mat, gap, t, p, world), oracle.align(0, alpha, mat, gap, t, p), (world, alpha, n, m))
        if tb == "parallel":      # the same slices filled in row chunks (the kernel path of the multi-GPU pipeline)
            assert_same(_strip_align(sa, alpha, mat, gap, t, p, world, chunks=3), oracle.align(0, alpha, mat, gap, t, p),
                        (world, alpha, n, m, "row chunks"))
    # low-complexity input: every tie-break rule on the slice borders
    t = np.zeros(1200, np.uint8); p = np.zeros(1100, np.uint8); p[::9] = 1
    assert_same(_strip_align(sa, 4, blast, 5, t, p, world), oracle.align(0, 4, blast, 5, t, p), (world, "ties"))


@pytest.mark.parametrize("world", [2, 5])
def test_linked_slices_vs_oracle(sa, oracle, world):
    """The in-launch hand-off of the border column ({4H, tag} words in the neighbour's buffer, sa_strip_fill_linked):
    same result as the oracle.  On one GPU the slices are launched left to right (the protocol, not the overlap)."""
    from sa_b200 import strips
    rng = np.random.default_rng(300 + world)
    blast = helpers.matrices()["dna/blast.txt"]
    for n, m in ((2500, 2300), (900, 1700), (12, 7)):
        import synth
        t = rng.integers(0, 4, n, dtype=np.uint8)
        base = synth.mutate_indices_numpy(t, rng, 4)
        p = base[:m] if m <= len(base) else np.concatenate((base, rng.integers(0, 4, m - len(base), dtype=np.uint8)))
        als = [sa.Aligner(0) for _ in range(world)]
        try:
            eng = [strips.GpuStripEngine(al, 4, blast, 5, t[c0:c0 + w], c0, len(t), p)
                   for al, (c0, w) in zip(als, strips.slice_columns(len(t), world))]
            for tag in (1, 2):          # twice: the tags of the first call must not satisfy the second
                score, at, ap, ti, pi = strips.align_pair_strips_linked_local(eng, len(p), tag=tag)
                assert_same(sa.Alignment(score, len(at), ti, pi, at, ap), oracle.align(0, 4, blast, 5, t, p), (world, n, m, tag))
            for e in eng:
                if hasattr(e, "border_ptr"):
                    e.al.peer_free(e.border_ptr)
        finally:
            for al in als:
                al.close()


def test_column_slices_full_size_c3(sa, aligner):
    """The slice path on the 100 000 x 95 217 pair (4 slices): same answer as the single-matrix path and
    the reference's known answer for this pair."""
    import synth
    t, p = synth.synthetic_pair(100000, 12345, 54321)
    blast = helpers.matrices()["dna/blast.txt"]
    a = _strip_align(sa, 4, blast, 5, t, p, 4)
    assert (a.score, a.aln_len, a.start_text, a.start_pattern) == (399463, 100254, 0, 0)
    assert_same(_strip_align(sa, 4, blast, 5, t, p, 2, chunks=5), a, "c3 slices in row chunks")
    assert_same(a, aligner.align(0, 4, blast, 5, t, p), "c3 slices vs single matrix")


def test_full_size_c4_batch_properties(sa, aligner, oracle):
    """BASELINE config 4 shape (100 000 pairs of the 1 M recipe): every pair's alignment re-scores to its
    score, a sample is compared field by field with the oracle, s16x2 and s32 kernels agree on all pairs."""
    import synth
    T, toff, P, poff = synth.synthetic_batch(100000, seed=2024)
    mat = helpers.matrices()["protein/blosum62.txt"]
    out = aligner.align_batch(1, 23, mat, 5, T, toff, P, poff)
    res = out["results"]
    assert int(res["score"].min()) > 0
    rng = np.random.default_rng(0)
    for i in rng.integers(0, 100000, 400):
        got = sa.unpack_batch(out, int(i))
        assert_same(got, oracle.align(1, 23, mat, 5, T[toff[i]:toff[i + 1]], P[poff[i]:poff[i + 1]]), int(i))
    for i in range(0, 100000, 37):
        got = sa.unpack_batch(out, i)
        assert oracle.rescore(got.aligned_text, got.aligned_pattern, 23, mat, 5) == got.score, i
    os.environ["SA_BATCH_S16"] = "0"
    try:
        out32 = aligner.align_batch(1, 23, mat, 5, T, toff, P, poff)
    finally:
        os.environ.pop("SA_BATCH_S16", None)
    assert np.array_equal(out32["results"], res)
    h = lambda o: (o["aligned_text"].tobytes(), o["aligned_pattern"].tobytes(), o["aln_off"].tobytes())
    k = 5000   # strings of the first pairs byte for byte (the arenas' unused slack is not compared)
    for i in range(k):
        assert sa.unpack_batch(out, i).key() == sa.unpack_batch(out32, i).key()


def test_reference_gpu_path_agrees(aligner):
    """Cross-check against the reference's OWN GPU path (alignSequenceGPU.cu, unmodified, oracle/_ref): the drop-in
    must return what it returns.  Skipped when the reference build did not travel."""
    from oracle.oracle_py import ReferenceGpu
    try:
        ref = ReferenceGpu(bench=False)
    except (FileNotFoundError, OSError, AttributeError) as e:
        pytest.skip(f"reference GPU build not available: {e}")
    rng = np.random.default_rng(77)
    blast = helpers.matrices()["dna/blast.txt"]
    b50 = helpers.matrices()["protein/blosum50.txt"]
    for mode, alpha, mat, n in ((0, 4, blast, 700), (1, 4, blast, 900), (0, 23, b50, 1200), (1, 23, b50, 2500)):
        t, p = helpers.random_case(rng, alpha, n)
        assert_same(aligner.align(mode, alpha, mat, 5, t, p), ref.align(mode, alpha, mat, 5, t, p), ("reference GPU", mode, alpha, n))


def test_linked_slices_two_gpus(sa):
    """The real thing on two GPUs (skipped on a one-GPU box): two processes, CUDA IPC border buffers, both slices
    launched at once; rank 0 also runs the single-matrix path and compares (bench_c5.py --check)."""
    import json
    import subprocess
    import sys
    if sa.lib().sa_device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29633", os.path.join(root, "bench_c5.py"), "--length", "60000", "--steps", "2", "--check"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and "linked" in line["config"]["pipeline"]
    assert all(v is True for k, v in line["checks"].items() if k != "starts"), line["checks"]
    assert line["checks"]["equals_single_matrix_path"] is True


@pytest.mark.parametrize("gap", [0, 1, 31, 32, 40, 300])
def test_gap_penalty_extremes(sa, aligner, oracle, force_path, gap):
    """Gap penalties at the edges of the fast paths: 0 (ties everywhere, sentinel columns do not decay), 31 / 32 (the NW
    form of the straight-line batch kernel needs 4*gap - 128 < 2), 40 and 300 (s16 guard -> s32 kernels).  Single pairs
    through both kernels and a batch, both modes, against the oracle."""
    rng = np.random.default_rng(900 + gap)
    blast = helpers.matrices()["dna/blast.txt"]
    b62 = helpers.matrices()["protein/blosum62.txt"]
    for path in ("batch", "long"):
        force_path(path)
        for mode in (0, 1):
            for alpha, mat in ((4, blast), (23, b62)):
                t, p = helpers.random_case(rng, alpha, 330)
                assert_same(aligner.align(mode, alpha, mat, gap, t, p), oracle.align(mode, alpha, mat, gap, t, p), (path, mode, alpha, gap))
    force_path(None)
    import synth
    T, toff, P, poff = synth.synthetic_batch(300, seed=gap + 1, lo=200, hi=340)
    for mode in (0, 1):
        out = aligner.align_batch(mode, 23, b62, gap, T, toff, P, poff)
        for i in range(0, 300, 7):
            want = oracle.align(mode, 23, b62, gap, T[toff[i]:toff[i + 1]], P[poff[i]:poff[i + 1]])
            assert_same(sa.unpack_batch(out, i), want, ("batch", mode, gap, i))


def test_wide_score_matrices(sa, aligner, oracle):
    """Score matrices beyond +-31 (the reference takes any int): they run through the long-pair kernel with the
    two-plane profile, whatever the pair size -- single pairs, a host batch; the device-resident batch refuses them."""
    rng = np.random.default_rng(4064)
    big_dna = np.full((4, 4), -77, np.int32); np.fill_diagonal(big_dna, 100)
    b62x25 = helpers.matrices()["protein/blosum62.txt"].astype(np.int32) * 25
    edge = rng.integers(-4064, 4065, (23, 23)).astype(np.int32); edge[0, 0] = 4064; edge[1, 2] = -4064
    for alpha, mat, gap in ((4, big_dna, 60), (23, b62x25, 125), (23, edge, 900), (4, big_dna, 5)):
        for mode in (0, 1):
            for n in (7, 300, 2100):
                t, p = helpers.random_case(rng, alpha, n)
                assert_same(aligner.align(mode, alpha, mat, gap, t, p), oracle.align(mode, alpha, mat, gap, t, p), ("wide", alpha, gap, mode, n))
    import synth
    T, toff, P, poff = synth.synthetic_batch(24, seed=9, lo=60, hi=200)
    out = aligner.align_batch(1, 23, b62x25, 125, T, toff, P, poff)
    for i in range(24):
        assert_same(sa.unpack_batch(out, i), oracle.align(1, 23, b62x25, 125, T[toff[i]:toff[i + 1]], P[poff[i]:poff[i + 1]]), ("wide batch", i))
    too_big = big_dna.copy(); too_big[0, 0] = 5000
    with pytest.raises(sa.SaError) as e:
        aligner.align(0, 4, too_big, 5, T[:50] % 4, P[:40] % 4)
    assert e.value.status == -5
    # 4*H must fit 32 bits: a gap of 2^20 on a 600-residue pair would not -- refused loudly, not wrapped
    with pytest.raises(sa.SaError) as e:
        aligner.align(0, 4, big_dna, 1 << 20, T[:600] % 4, P[:580] % 4)
    assert e.value.status == -5
    # after a wide call the context must be back on the fast paths
    t, p = helpers.random_case(rng, 4, 200)
    blast = helpers.matrices()["dna/blast.txt"]
    assert_same(aligner.align(0, 4, blast, 5, t, p), oracle.align(0, 4, blast, 5, t, p), "narrow after wide")


def test_batch_length_spectrum(sa, aligner, oracle):
    """One host batch spanning everything the batch kernels take -- texts 1..4096, patterns 1..1536, tiny and ragged
    pairs, similar and unrelated ones (every class, lanes without rows, pattern > text) -- against the oracle."""
    rng = np.random.default_rng(2025)
    b50 = helpers.matrices()["protein/blosum50.txt"]
    N = 360
    for mode in (0, 1):
        n = np.concatenate((rng.integers(1, 40, 80), rng.integers(1, 700, 200), rng.integers(700, 4097, 80)))
        m = np.concatenate((rng.integers(1, 40, 80), rng.integers(1, 400, 200), rng.integers(300, 1537, 80)))
        rng.shuffle(n); rng.shuffle(m)
        toff = np.concatenate(([0], np.cumsum(n))).astype(np.int64); poff = np.concatenate(([0], np.cumsum(m))).astype(np.int64)
        T = rng.integers(0, 23, toff[-1], dtype=np.uint8); P = rng.integers(0, 23, poff[-1], dtype=np.uint8)
        for i in range(0, N, 3):
            k = min(n[i], m[i]); P[poff[i]:poff[i] + k] = T[toff[i]:toff[i] + k]
        out = aligner.align_batch(mode, 23, b50, 4, T, toff, P, poff)
        for i in range(N):
            assert_same(sa.unpack_batch(out, i), oracle.align(mode, 23, b50, 4, T[toff[i]:toff[i + 1]], P[poff[i]:poff[i + 1]]),
                        ("spectrum", mode, i, int(n[i]), int(m[i])))


def test_host_batch_of_medium_pairs_runs_concurrently(sa, aligner, oracle):
    """A host batch whose members are all too long for the batch kernels: they go through the long-pair kernels on
    several worker sub-contexts side by side (SA_LONG_WORKERS); every Response must still equal the oracle's."""
    rng = np.random.default_rng(8192)
    blast = helpers.matrices()["dna/blast.txt"]
    n = rng.integers(1700, 5200, 14); m = rng.integers(1600, 4800, 14)
    toff = np.concatenate(([0], np.cumsum(n))).astype(np.int64); poff = np.concatenate(([0], np.cumsum(m))).astype(np.int64)
    T = rng.integers(0, 4, toff[-1], dtype=np.uint8); P = rng.integers(0, 4, poff[-1], dtype=np.uint8)
    for i in range(0, 14, 2):
        k = min(n[i], m[i]); P[poff[i]:poff[i] + k] = T[toff[i]:toff[i] + k]; P[poff[i]:poff[i] + k:11] ^= 1
    for mode in (0, 1):
        out = aligner.align_batch(mode, 4, blast, 5, T, toff, P, poff)
        for i in range(14):
            assert_same(sa.unpack_batch(out, i), oracle.align(mode, 4, blast, 5, T[toff[i]:toff[i + 1]], P[poff[i]:poff[i + 1]]),
                        ("medium batch", mode, i))


def test_host_batch_staged_pipeline_equals_slot_pipeline(sa, aligner, oracle):
    """The host pipelines of sa_align_batch (staged: one fill stream + traceback / copy streams, strings packed on the
    device before the copy or copied as whole slots; slots: one stream per chunk) return the same results and strings
    on a 20 011-pair batch with ragged chunk sizes, and a sample equals the oracle."""
    import synth
    T, toff, P, poff = synth.synthetic_batch(20011, seed=4242)
    mat = helpers.matrices()["protein/blosum62.txt"]
    rng = np.random.default_rng(5)
    keys = ("SA_HOST_PIPELINE", "SA_HOST_CHUNKS", "SA_HOST_PACK")
    for mode in (0, 1):
        outs = {}
        for cfg in (("staged", "5", "1"), ("staged", "5", "0"), ("slots", "5", "1"), ("staged", "3", "1"), ("staged", "1", "1")):
            os.environ.update(dict(zip(keys, cfg)))
            try:
                outs[cfg] = aligner.align_batch(mode, 23, mat, 5, T, toff, P, poff)
            finally:
                for k in keys:
                    os.environ.pop(k, None)
        a = outs["staged", "5", "1"]
        total = int(a["results"]["aln_len"].sum())
        # packed: pair p's strings start where pair p-1's end
        assert np.array_equal(a["aln_off"], np.concatenate(([0], np.cumsum(a["results"]["aln_len"])[:-1])).astype(np.uint64))
        assert aligner.timing()["d2h_bytes"] == 40 * 20011 + 2 * total
        for key, o in outs.items():
            for f in ("score", "aln_len", "start_text", "start_pattern"):      # (not the struct's padding word)
                assert np.array_equal(o["results"][f], a["results"][f]), (mode, key, f)
        for i in list(range(0, 20011, 53)) + [20010]:
            ka = sa.unpack_batch(a, i).key()
            for key, o in outs.items():
                assert sa.unpack_batch(o, i).key() == ka, (mode, key, i)
        for i in rng.integers(0, 20011, 60):
            assert_same(sa.unpack_batch(a, int(i)), oracle.align(mode, 23, mat, 5, T[toff[i]:toff[i + 1]], P[poff[i]:poff[i + 1]]), (mode, int(i)))


def test_pinned_host_buffers(sa, aligner, oracle):
    """sa_host_alloc / sa_host_register: a batch in page-locked buffers gives the same answers as in pageable ones."""
    import synth
    T, toff, P, poff = synth.synthetic_batch(9000, seed=99)
    mat = helpers.matrices()["protein/blosum62.txt"]
    ref = aligner.align_batch(1, 23, mat, 5, T, toff, P, poff)
    arena = int(toff[-1] + poff[-1])
    out = dict(results=sa.pinned_empty(9000, sa.RESULT_DTYPE), aln_off=sa.pinned_empty(9000, np.uint64),
               aligned_text=sa.pinned_empty(arena, np.uint8), aligned_pattern=sa.pinned_empty(arena, np.uint8))
    got = aligner.align_batch(1, 23, mat, 5, sa.pinned_copy(T), sa.pinned_copy(toff), sa.pinned_copy(P), sa.pinned_copy(poff), out=out)
    for i in range(0, 9000, 7):
        assert sa.unpack_batch(got, i).key() == sa.unpack_batch(ref, i).key(), i
    i = 4321
    assert_same(sa.unpack_batch(got, i), oracle.align(1, 23, mat, 5, T[toff[i]:toff[i + 1]], P[poff[i]:poff[i + 1]]), i)
    # registering an existing buffer, and the error paths
    buf = np.zeros(1 << 20, np.uint8)
    assert sa.lib().sa_host_register(buf.ctypes.data, buf.nbytes) == 0
    assert sa.lib().sa_host_unregister(buf.ctypes.data) == 0
    assert sa.lib().sa_host_register(None, 16) != 0
    del out, got


def _device_batch(sa, al, mode, mat, T, toff, P, poff, max_n=None, max_m=None):
    """sa_align_batch_device on torch tensors; returns the result set as host numpy arrays."""
    import torch
    dev = torch.device("cuda:0")
    dT, dP = torch.from_numpy(T).to(dev), torch.from_numpy(P).to(dev)
    dto, dpo = torch.from_numpy(toff).to(dev), torch.from_numpy(poff).to(dev)
    N = len(toff) - 1
    arena = int(toff[-1] + poff[-1])
    res = torch.zeros(N * 4, dtype=torch.int64, device=dev)
    aoff = torch.zeros(N, dtype=torch.int64, device=dev)
    oT = torch.zeros(arena, dtype=torch.uint8, device=dev)
    oP = torch.zeros(arena, dtype=torch.uint8, device=dev)
    max_n = max_n or int((toff[1:] - toff[:-1]).max())
    max_m = max_m or int((poff[1:] - poff[:-1]).max())
    al.align_batch_device(mode, 23, mat, 5, N, dT.data_ptr(), dto.data_ptr(), dP.data_ptr(), dpo.data_ptr(),
                          res.data_ptr(), aoff.data_ptr(), oT.data_ptr(), oP.data_ptr(), arena, max_n, max_m,
                          stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return dict(results=res.cpu().numpy().view(sa.RESULT_DTYPE), aln_off=aoff.cpu().numpy().astype(np.uint64),
                aligned_text=oT.cpu().numpy(), aligned_pattern=oP.cpu().numpy())


@pytest.mark.parametrize("mode", [1, 0])
def test_device_batch_pipelined_chunks_all_pairs(sa, oracle, monkeypatch, mode):
    """The path behind the bench's `value`: sa_align_batch_device with the chunk pipeline engaged (>= 8192 pairs) and a
    direction budget small enough for >= 5 chunks, so that fill / traceback overlap and both buffer sets are reused.
    EVERY pair is compared field by field and string by string with the unmodified reference (all host threads),
    and the whole result set with the host-buffer path."""
    import synth
    from oracle.oracle_py import Reference
    N = 24000
    T, toff, P, poff = synth.synthetic_batch(N, seed=777)
    mat = helpers.matrices()["protein/blosum62.txt"]
    monkeypatch.setenv("SA_DIRS_BUDGET_MB", "160")
    al = sa.Aligner(0)
    try:
        out = _device_batch(sa, al, mode, mat, T, toff, P, poff)
        # the chunks of the call: 4 events (fill start/end, traceback start/end) per chunk in the context's timing pool
        launches = al.timing()["kernel_launches"]
    finally:
        al.close()
    per_chunk = 3 + 1 + 9          # binning kernels, traceback, at most 9 class kernels
    assert launches >= 5 * 5, launches          # >= 5 chunks x (3 binning + >= 1 class + 1 traceback)
    assert launches / per_chunk < 40
    chk = Reference("O3") if Reference.available("O3") else oracle
    assert chk.check_batch(mode, 23, mat, 5, T, toff, P, poff, out) == (0, -1)
    monkeypatch.delenv("SA_DIRS_BUDGET_MB")
    al2 = sa.Aligner(0)
    try:
        host = al2.align_batch(mode, 23, mat, 5, T, toff, P, poff)
    finally:
        al2.close()
    for f in ("score", "aln_len", "start_text", "start_pattern"):
        assert np.array_equal(host["results"][f], out["results"][f]), f
    # strings: the host path returns them packed, the device path in slots -- compare through a checksum of every pair
    def digest(o):
        ln = o["results"]["aln_len"].astype(np.int64)
        off = o["aln_off"].astype(np.int64)
        idx = np.repeat(off - np.concatenate(([0], np.cumsum(ln)[:-1])), ln) + np.arange(int(ln.sum()))
        return helpers.sha(o["aligned_text"][idx].tobytes()), helpers.sha(o["aligned_pattern"][idx].tobytes())
    assert digest(host) == digest(out)


@pytest.mark.parametrize("pack", ["1", "0"])
def test_host_batch_csr_with_nonzero_first_offset(sa, aligner, oracle, monkeypatch, pack):
    """A CSR batch whose first offsets are not 0 (a view into a larger batch) through the staged host pipeline, packed
    and unpacked: aln_off must count from the start of the output arenas either way."""
    import synth
    T, toff, P, poff = synth.synthetic_batch(9500, seed=31)
    k = 700
    mat = helpers.matrices()["protein/blosum62.txt"]
    monkeypatch.setenv("SA_HOST_PACK", pack)
    out = aligner.align_batch(1, 23, mat, 5, T, toff[k:], P, poff[k:])
    N = len(toff) - 1 - k
    arena = int(toff[-1] - toff[k] + poff[-1] - poff[k])
    assert int(out["aln_off"].max()) < arena
    sub = dict(out)
    assert oracle.check_batch(1, 23, mat, 5, T[toff[k]:], toff[k:] - toff[k], P[poff[k]:], poff[k:] - poff[k], sub) == (0, -1)
    if pack == "1":
        assert np.array_equal(out["aln_off"], np.concatenate(([0], np.cumsum(out["results"]["aln_len"])[:-1])).astype(np.uint64))
    assert N == len(out["results"])


def test_context_options_replace_the_environment(sa, oracle):
    """sa_set_option / sa_get_option: the chunking of the device batch set through the C ABI instead of SA_* variables
    (eight or more chunks of a 12 000-pair batch), results equal to the oracle's; unknown names and bad values refused."""
    import synth
    al = sa.Aligner(0)
    try:
        assert al.get_option("dev_dirs_budget_mb") == 3500
        al.set_option("dev_dirs_budget_mb", 64)
        al.set_option("tb_blocks_per_sm", 2)
        assert al.get_option("dev_dirs_budget_mb") == 64 and al.get_option("tb_blocks_per_sm") == 2
        for bad in (("no_such_option", 1), ("dev_dirs_budget_mb", 0), ("tb_blocks_per_sm", 99), ("ckpt_rows", -1)):
            with pytest.raises(sa.SaError):
                al.set_option(*bad)
        with pytest.raises(sa.SaError):
            al.get_option("no_such_option")
        T, toff, P, poff = synth.synthetic_batch(12000, seed=99)
        mat = helpers.matrices()["protein/blosum62.txt"]
        out = _device_batch(sa, al, 1, mat, T, toff, P, poff)
        launches = al.timing()["kernel_launches"]
        assert launches >= 8 * 5, launches
        assert oracle.check_batch(1, 23, mat, 5, T, toff, P, poff, out) == (0, -1)
        # the checkpointed traceback by option: a 3000 x 2800 global alignment in chunks of 700 rows
        rng = np.random.default_rng(5)
        t = rng.integers(0, 4, 3000, dtype=np.uint8)
        p2 = t[:2800].copy(); p2[::17] = (p2[::17] + 1) % 4
        dna = helpers.matrices()["dna/blast.txt"]
        want = oracle.align(0, 4, dna, 5, t, p2)
        al.set_option("ckpt_rows", 700)
        got = al.align(0, 4, dna, 5, t, p2)
        assert got.key() == want.key()
        al.set_option("ckpt_rows", 0)
        assert al.align(0, 4, dna, 5, t, p2).key() == want.key()
    finally:
        al.close()


def test_residues_outside_the_alphabet_are_refused(sa, aligner):
    """include/sa_b200.h: a residue >= alphabet_size is SA_ERR_ARGUMENT from the host-buffer entry points."""
    import synth
    mat = helpers.matrices()["dna/blast.txt"]
    t, p = np.array([0, 1, 2, 3, 4], np.uint8), np.array([0, 1, 2], np.uint8)
    with pytest.raises(sa.SaError):
        aligner.align(0, 4, mat, 5, t, p)
    with pytest.raises(sa.SaError):
        aligner.align(1, 4, mat, 5, p, np.frombuffer(b"ACGT", np.uint8))          # ASCII instead of indices
    T, toff, P, poff = synth.synthetic_batch(9000, seed=5)
    bl = helpers.matrices()["protein/blosum62.txt"]
    for N in (9000, 300):                                                           # staged pipeline and the slot pipeline
        T2 = T.copy()
        T2[int(toff[N - 3]) + 7] = 23
        with pytest.raises(sa.SaError):
            aligner.align_batch(1, 23, bl, 5, T2, toff[:N + 1], P, poff[:N + 1])
        aligner.align_batch(1, 23, bl, 5, T, toff[:N + 1], P, poff[:N + 1])         # the context is still usable


def test_device_batch_marks_pairs_it_cannot_take(sa, aligner, oracle):
    """sa_align_batch_device: a pair longer than the caller's max_text_len / max_pattern_len is not aligned and is
    marked with the sentinel result instead of overrunning the windows sized from those bounds."""
    import synth
    T, toff, P, poff = synth.synthetic_batch(600, seed=12, lo=100, hi=200)
    mat = helpers.matrices()["protein/blosum62.txt"]
    lens_n, lens_m = toff[1:] - toff[:-1], poff[1:] - poff[:-1]
    cut_n = int(lens_n.max()) - 2          # the longest texts are above the bound
    out = _device_batch(sa, aligner, 1, mat, T, toff, P, poff, max_n=cut_n, max_m=int(lens_m.max()))
    over = np.flatnonzero(lens_n > cut_n)
    assert 1 <= len(over) < 100
    for i in over:
        assert int(out["results"]["score"][i]) == -2**31 and int(out["results"]["aln_len"][i]) == 0
    keep = np.flatnonzero(lens_n <= cut_n).astype(np.uint64)
    assert oracle.check_batch(1, 23, mat, 5, T, toff, P, poff, out, idx=keep) == (0, -1)


TILE_SHAPES = ["4,4", "8,4", "8,2", "4,8", "2,8", "16,4", "8,8"]


@pytest.mark.parametrize("tile", TILE_SHAPES)
def test_goldens_through_tile_kernel(aligner, force_path, monkeypatch, tile):
    """Short and medium goldens forced through the register-tiled long-pair kernel (sa_tile.cuh) at every tile shape."""
    force_path("long")
    monkeypatch.setenv("SA_TILE", tile)
    gs = [g for g in helpers.goldens() if helpers.has_inputs(g) and g["n"] * g["m"] < 3e6]
    k = TILE_SHAPES.index(tile)
    for g in gs[k % 2::2]:
        t, p, mat = helpers.golden_inputs(g)
        a = aligner.align(g["mode"], g["alpha"], mat, g["gap"], t, p)
        helpers.check_against_golden(a, g)


@pytest.mark.parametrize("tile", TILE_SHAPES)
def test_tile_kernel_random_and_ties_vs_oracle(aligner, oracle, force_path, monkeypatch, tile):
    """Random pairs of every length class around the tile / group boundaries (n mod C, n around 32*C and TG*C, m around
    32*R), random matrices and gaps, both modes; then low-complexity inputs where every cell is a tie and the SW arg-max
    has many equal candidates in different tiles, lanes and strips."""
    force_path("long")
    monkeypatch.setenv("SA_TILE", tile)
    R, C = (int(x) for x in tile.split(","))
    rng = np.random.default_rng(hash(tile) % 10000)
    blast = helpers.matrices()["dna/blast.txt"]
    b62 = helpers.matrices()["protein/blosum62.txt"]
    lens = [1, 2, C - 1, C, C + 1, 31 * C, 32 * C, 32 * C + 1, 33 * C + C // 2, 8 * C * 5 + 3, 700, 1500]
    rows = [1, R, R + 1, 32 * R - 1, 32 * R, 32 * R + 1, 70 * R, 1300]
    for it in range(60):
        alpha = 4 if it % 2 == 0 else 23
        mat = rng.integers(-9, 12, (alpha, alpha)).astype(np.int32) if it % 3 == 0 else (blast if alpha == 4 else b62)
        n, m = lens[it % len(lens)], rows[(it // 2) % len(rows)]
        t = rng.integers(0, alpha, max(1, n), dtype=np.uint8)
        if it % 4 < 2 and m <= len(t):
            p = t[:m].copy()
            p[rng.random(len(p)) < 0.1] = rng.integers(0, alpha)
        else:
            p = rng.integers(0, alpha, m, dtype=np.uint8)
        gap = int(rng.integers(0, 12))
        for mode in (0, 1):
            assert_same(aligner.align(mode, alpha, mat, gap, t, p), oracle.align(mode, alpha, mat, gap, t, p),
                        (tile, it, alpha, mode, gap, len(t), len(p)))
    for mode in (0, 1):
        for t, p in ((np.zeros(333, np.uint8), np.zeros(297, np.uint8)),
                     (np.tile(np.arange(4, dtype=np.uint8), 190), np.tile(np.arange(4, dtype=np.uint8), 133)),
                     (np.tile(np.array([0, 0, 1], np.uint8), 300), np.tile(np.array([0, 1], np.uint8), 260))):
            for gap in (0, 1, 5):
                assert_same(aligner.align(mode, 4, blast, gap, t, p), oracle.align(mode, 4, blast, gap, t, p), (tile, mode, gap, len(t)))


@pytest.mark.parametrize("tile", ["4,4", "8,4", "8,2"])
def test_tile_kernel_slices_and_row_chunks(sa, oracle, monkeypatch, tile):
    """Column slices (int32 border columns, row chunks, linked {4H, tag} borders) through the tiled kernel."""
    from sa_b200 import strips
    import synth
    monkeypatch.setenv("SA_TILE", tile)
    rng = np.random.default_rng(77)
    blast = helpers.matrices()["dna/blast.txt"]
    t = rng.integers(0, 4, 2700, dtype=np.uint8)
    p = synth.mutate_indices_numpy(t, rng, 4)[:2450]
    want = oracle.align(0, 4, blast, 5, t, p)
    for world in (2, 3):
        assert_same(_strip_align(sa, 4, blast, 5, t, p, world), want, (tile, world))
        assert_same(_strip_align(sa, 4, blast, 5, t, p, world, chunks=3), want, (tile, world, "row chunks"))
        als = [sa.Aligner(0) for _ in range(world)]
        try:
            eng = [strips.GpuStripEngine(al, 4, blast, 5, t[c0:c0 + w], c0, len(t), p)
                   for al, (c0, w) in zip(als, strips.slice_columns(len(t), world))]
            for tag in (1, 2):
                score, at, ap, ti, pi = strips.align_pair_strips_linked_local(eng, len(p), tag=tag)
                assert_same(sa.Alignment(score, len(at), ti, pi, at, ap), want, (tile, world, "linked", tag))
            for e in eng:
                if hasattr(e, "border_ptr"):
                    e.al.peer_free(e.border_ptr)
        finally:
            for al in als:
                al.close()


def test_tile_kernel_many_strips_in_waves(aligner, oracle, force_path, monkeypatch):
    """More strips than resident warps: the persistent loop takes strips in waves and the boundary ring wraps."""
    force_path("long")
    monkeypatch.setenv("SA_TILE", "2,8")
    import synth
    rng = np.random.default_rng(3)
    b62 = helpers.matrices()["protein/blosum62.txt"]
    t = rng.integers(0, 23, 600, dtype=np.uint8)
    p = np.concatenate([synth.mutate_indices_numpy(t, rng, 23) for _ in range(520)])[:300000]       # 4688 strips of 64 rows
    for mode in (0, 1):
        got = aligner.align(mode, 23, b62, 5, t, p)
        s, _ = oracle.score_only(mode, 23, b62, 5, t, p)
        assert got.score == s
        assert oracle.rescore(got.aligned_text, got.aligned_pattern, 23, b62, 5) == got.score


def test_device_side_identity_and_gap_counts(sa, aligner, oracle, force_path, monkeypatch):
    """the two counts of prettyAlignmentPrint (utilities.cpp:262-283) come from the device, taken while the
    strings are emitted: sa_last_stats for single pairs (batch kernels, tiled and one-column long-pair kernels, serial
    traceback) and sa_batch_out.stats for batches -- equal to a host pass over the emitted strings."""
    rng = np.random.default_rng(8)
    blast = helpers.matrices()["dna/blast.txt"]
    b62 = helpers.matrices()["protein/blosum62.txt"]

    def counts(a):
        at, ap = np.frombuffer(a.aligned_text, np.uint8), np.frombuffer(a.aligned_pattern, np.uint8)
        return int((at == ap).sum()), int(((at == 45) | (ap == 45)).sum())

    for path, env in (("batch", {}), ("long", {}), ("long", {"SA_LONG_KERNEL": "strip"}), ("long", {"SA_TB": "serial"}), ("long", {"SA_TILE": "4,4"})):
        force_path(path)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for it in range(6):
            alpha, mat = (4, blast) if it % 2 == 0 else (23, b62)
            t, p = helpers.random_case(rng, alpha, n_max=[60, 300, 1500][it % 3] if path == "long" else 300, similar=it % 3 != 2)
            for mode in (0, 1):
                a = aligner.align(mode, alpha, mat, 3, t, p)
                assert_same(a, oracle.align(mode, alpha, mat, 3, t, p), (path, env, it, mode))
                if a.aln_len:
                    assert aligner.last_stats() == counts(a), (path, env, it, mode)
        for k in env:
            monkeypatch.delenv(k)
    force_path(None)
    import synth
    for N in (500, 9000):                  # the slot pipeline and the staged, packed pipeline
        T, toff, P, poff = synth.synthetic_batch(N, seed=77, lo=40, hi=300)
        arena = int(toff[-1] + poff[-1])
        out = dict(results=np.zeros(N, sa.RESULT_DTYPE), aln_off=np.zeros(N, np.uint64), aligned_text=np.empty(arena, np.uint8),
                   aligned_pattern=np.empty(arena, np.uint8), stats=np.full(2 * N, 0xffffffff, np.uint32))
        for mode in (1, 0):
            aligner.align_batch(mode, 23, b62, 5, T, toff, P, poff, out=out)
            for i in list(range(0, N, 97)) + [N - 1]:
                a = sa.unpack_batch(out, i)
                assert (int(out["stats"][2 * i]), int(out["stats"][2 * i + 1])) == counts(a), (N, mode, i)


def test_config5_recipe_score_pinned_by_the_cpu_oracle(sa, aligner):
    """BASELINE config 5's seeded pair (seeds 777 / 778) at 100 000 letters, as column slices: the score equals
    the CPU oracle's (tests/golden/c5_golden.json, made by tests/golden/make_c5_golden.py; the same file pins the
    1 000 000-letter score 3 972 370 that bench.py / bench_c5.py compare the multi-GPU run with)."""
    import json
    import synth
    gold = json.load(open(os.path.join(helpers.GOLD, "c5_golden.json")))
    assert gold["1000000"]["score"] == 3972370 and gold["1000000"]["m"] == 950793
    t, p = synth.synthetic_pair(100000, 777, 778)
    g = gold["100000"]
    assert (len(t), len(p)) == (g["n"], g["m"])
    blast = helpers.matrices()["dna/blast.txt"]
    a = _strip_align(sa, 4, blast, 5, t, p, 3)
    assert a.score == g["score"]
    assert_same(a, aligner.align(0, 4, blast, 5, t, p), "slices vs single matrix")


@pytest.mark.parametrize("variant", [dict(SA_CKPT_ROWS="300"), dict(SA_CKPT_ROWS="1000", SA_TILE="8,2"), dict(SA_CKPT_ROWS="700", SA_LONG_KERNEL="strip"),
                                     dict(SA_CKPT_ROWS="129", SA_TB_WD="16"), dict(SA_CKPT_ROWS="2000", SA_TB_BAND="1")])
def test_checkpointed_traceback_vs_oracle(aligner, oracle, force_path, monkeypatch, variant):
    """Linear-space traceback: the matrix is filled in row chunks that keep only an H row each, then every
    chunk is filled again from its checkpoint and walked, last chunk first.  Same Response as the single-matrix path and
    the oracle, byte for byte, whatever the chunk height, kernel and candidate spacing; the device-side identity / gap
    counts add up over the chunks.  (Local alignments never take this path.)"""
    force_path("long")
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(777)
    blast = helpers.matrices()["dna/blast.txt"]
    b62 = helpers.matrices()["protein/blosum62.txt"]
    cases = []
    for it in range(8):
        alpha, mat = (4, blast) if it % 2 == 0 else (23, b62)
        t, p = helpers.random_case(rng, alpha, n_max=[700, 2500, 6000][it % 3], similar=it % 4 != 3)
        cases.append((alpha, mat, t, p))
    cases.append((4, blast, rng.integers(0, 4, 3000, dtype=np.uint8), rng.integers(0, 4, 2500, dtype=np.uint8)))      # unrelated: the path meanders
    cases.append((4, blast, np.zeros(2000, np.uint8), np.zeros(1700, np.uint8)))                                    # ties everywhere
    cases.append((4, blast, rng.integers(0, 4, 900, dtype=np.uint8), rng.integers(0, 4, 2600, dtype=np.uint8)))       # more rows than columns
    for alpha, mat, t, p in cases:
        for gap in (2, 7):
            want = oracle.align(0, alpha, mat, gap, t, p)
            got = aligner.align(0, alpha, mat, gap, t, p)
            assert_same(got, want, (variant, alpha, gap, len(t), len(p)))
            st = aligner.last_stats()
            ident = sum(1 for a, b in zip(want.aligned_text, want.aligned_pattern) if a == b)
            gaps = sum(1 for a, b in zip(want.aligned_text, want.aligned_pattern) if a == ord("-") or b == ord("-"))
            assert st == (ident, gaps), (variant, len(t), len(p))
            assert aligner.fill_only(0, alpha, mat, gap, t, p)[0] == want.score          # score only: the first pass alone
            # a local alignment of the same pair is untouched by the switch
            assert_same(aligner.align(1, alpha, mat, gap, t, p), oracle.align(1, alpha, mat, gap, t, p), (variant, "local"))


def test_checkpointed_traceback_full_size_c3(aligner, oracle, monkeypatch):
    """Config 3 (100 000 x 95 217) through the checkpointed path in 5 row chunks: the reference's known answer."""
    import synth
    t, p = synth.synthetic_pair(100000, 12345, 54321)
    blast = helpers.matrices()["dna/blast.txt"]
    plain = aligner.align(0, 4, blast, 5, t, p)
    launches = aligner.timing()["kernel_launches"]
    monkeypatch.setenv("SA_CKPT_ROWS", "20000")
    a = aligner.align(0, 4, blast, 5, t, p)
    assert aligner.timing()["kernel_launches"] >= 5 * launches          # 5 fills + 4 re-fills + 5 tracebacks
    assert (a.score, a.aln_len, a.start_text, a.start_pattern) == (399463, 100254, 0, 0)
    assert_same(a, plain, "checkpointed vs single matrix")

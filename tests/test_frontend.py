"""Front end (csrc/sa_frontend.cpp through the C ABI and its Python mirror) against the reference's own parser
(utilities.cpp, through oracle/_ref -- test infrastructure): validateAndTransform, readSequenceFile,
parseScoreMatrixFile, plus the new multi-FASTA batch reader.  CPU only."""
import json
import os

import numpy as np
import pytest

import helpers
from gpu_common import load_package

DNA, PROT = b"ATCG-", b"ARNDCQEGHILKMFPSTWYVBZX-"

CASES = [
    b"ACGTACGT\n",
    b">seq1 some header with ACGT letters > and more\nacgtNOPE"[:-4] + b"\nAC GT\r\nTTga\n",
    b"  12 acgt 34 tgca *-.\n\n\n",
    b"ACGT>inline header starts mid-line ACGT\nTTTT\n>second\nGG\nCC",
    b">only a header, no newline at the end",
    b"",
    bytes(range(256)).replace(b">", b"") .translate(None, bytes(c for c in range(256) if chr(c).upper() in "BDEFHIJKLMNOPQRSUVWXYZ")),
    b"acgt[]{}`~^_|\x7f\x80\xff\xc3\xa9ACGT",
]


def test_validate_and_transform_matches_reference(reference):
    sa = load_package()
    rng = np.random.default_rng(5)
    cases = list(CASES)
    for _ in range(40):          # random soups of residues, blanks, digits, headers and line ends
        toks = rng.choice([b"A", b"c", b"G", b"t", b" ", b"\n", b"7", b">hdr x\n", b"\r\n", b"-", b"*"], rng.integers(1, 200))
        cases.append(b"".join(toks))
    for data in cases:
        want = reference.validate_and_transform(data, 4)
        got = sa.validateAndTransform(data, DNA, 4)
        assert np.array_equal(got, want), data[:60]
    prot = b">sp|P69905|HBA_HUMAN\nMVLSPADKTNVKAAWGKVGAHAGEYGAEALERMFLSFPTTKTYFPHF\nbzx  mvls*\n"
    assert np.array_equal(sa.validateAndTransform(prot, PROT, 23), reference.validate_and_transform(prot, 23))


def test_letter_outside_alphabet_is_the_references_failure(reference, capfd):
    sa = load_package()
    data = b"ACGTNACGT"                       # N is a letter but not a DNA letter: the reference returns 0
    assert len(reference.validate_and_transform(data, 4)) == 0
    assert len(sa.validateAndTransform(data, DNA, 4)) == 0
    assert "'N' letter not in alphabet." in capfd.readouterr().err
    assert len(sa.validateAndTransform(b"ACGTJ", PROT, 23)) == 0      # J, O, U are not protein letters either


def test_read_sequence_file_fills_text_then_pattern(reference, tmp_path):
    sa = load_package()
    f1, f2 = tmp_path / "a.fasta", tmp_path / "b.txt"
    f1.write_bytes(b">x\nACGTTGCA\nAC\n"); f2.write_bytes(b"ttga\n")
    rq = sa.Request()
    assert sa.readSequenceFile(str(f1), rq) == 0 and sa.readSequenceFile(str(f2), rq) == 0
    assert np.array_equal(rq.textBytes[:rq.textNumBytes], reference.read_sequence(str(f1), 4))
    assert np.array_equal(rq.patternBytes[:rq.patternNumBytes], reference.read_sequence(str(f2), 4))
    assert sa.readSequenceFile(str(tmp_path / "missing"), rq) == -1


def test_parse_score_matrix_file(reference, tmp_path):
    sa = load_package()
    mats = helpers.matrices()
    for name, alpha in (("dna/blast.txt", 4), ("protein/blosum62.txt", 23), ("protein/blosum50.txt", 23)):
        f = tmp_path / "m.txt"
        m = np.asarray(mats[name], np.int32).reshape(alpha, alpha)
        f.write_text("\n".join("  ".join(f"{v:3d}" for v in row) for row in m) + "\n")
        buf = np.zeros(23 * 23, np.int32)
        assert sa.parseScoreMatrixFile(str(f), alpha, buf) == 0
        assert np.array_equal(buf[:alpha * alpha], reference.parse_score_matrix(str(f), alpha))
        assert np.array_equal(buf[:alpha * alpha], m.ravel())
    bad = tmp_path / "bad.txt"; bad.write_text("1 2 3 x 5")
    assert sa.parseScoreMatrixFile(str(bad), 4, np.zeros(16, np.int32)) == -1
    untouched = np.full(16, 7, np.int32)
    assert sa.parseScoreMatrixFile(str(tmp_path / "missing.txt"), 4, untouched) == 0 and (untouched == 7).all()   # the reference's quirk


def test_read_fasta_batch(reference, tmp_path):
    sa = load_package()
    f = tmp_path / "multi.fasta"
    recs = [b"ACGTAC", b"ttgacc\nGGA", b"A", b"CCCC\r\nGG"]
    f.write_bytes(b"".join(b">r%d header\n" % i + r + b"\n" for i, r in enumerate(recs)))
    res, off = sa.read_fasta_batch(str(f), DNA, 4)
    assert len(off) == len(recs) + 1 and off[0] == 0
    for i, r in enumerate(recs):
        assert np.array_equal(res[off[i]:off[i + 1]], reference.validate_and_transform(r, 4)), i
    # a plain single-sequence file is one record, and the concatenation equals what the reference reads from it
    g = tmp_path / "plain.txt"; g.write_bytes(b"ACGT\nACGT\n")
    res, off = sa.read_fasta_batch(str(g), DNA, 4)
    assert list(off) == [0, 8] and np.array_equal(res, reference.read_sequence(str(g), 4))
    with pytest.raises(sa.SaError):
        h = tmp_path / "bad.fasta"; h.write_bytes(b">x\nACGTN\n")
        sa.read_fasta_batch(str(h), DNA, 4)


def test_pretty_print_matches_reference(reference, oracle):
    """prettyAlignmentPrint byte for byte, on golden alignments of all kinds (short, multi-line, wide position columns,
    local alignments with non-zero starts) and the empty alignment."""
    sa = load_package()
    from oracle.oracle_py import Alignment
    gs = helpers.goldens()
    small = [g for g in gs if g["n"] * g["m"] <= 4_000_000]
    picked = [g for g in small if g["ref"]["aln_len"] > 0][:: max(1, len(small) // 40)] + [g for g in small if g["ref"]["aln_len"] == 0][:2]
    assert len(picked) > 20
    for g in picked:
        t, p, mat = helpers.golden_inputs(g)
        a = oracle.align(g["mode"], g["alpha"], mat, g["gap"], t, p)
        rs = sa.Response(alignedTextBytes=a.aligned_text, alignedPatternBytes=a.aligned_pattern, numAlignmentBytes=a.aln_len,
                         startInAlignedText=a.start_text, startInAlignedPattern=a.start_pattern, score=a.score)
        want = reference.pretty_print(Alignment(a.score, a.aln_len, a.start_text, a.start_pattern, a.aligned_text, a.aligned_pattern))
        assert sa.prettyAlignmentPrint(rs) == want, g.get("name")
    ident, gaps = sa.alignment_stats(b"AC-GT", b"ACTG-")
    assert (ident, gaps) == (3, 2)

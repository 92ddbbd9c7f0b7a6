#!/usr/bin/env python
"""Generate tests/golden/* from the UNMODIFIED reference (run in the build container).

Needs /root/reference (data/, scoreMatrices/) and oracle/_ref/libsa_ref_O3.so
(``make -C oracle``).  Nothing here runs on the GPU box: the outputs are small,
committed fixtures.

Outputs
  sequences.npz            index-encoded sequences of data/dna + data/protein
                           (as read by the reference's own readSequenceFile /
                           validateAndTransform, utilities.cpp:65,31); files longer
                           than MAX_LEN residues are left out to keep the repo small
  matrices.json            every parseable scoreMatrices/*/*.txt as parsed by the
                           reference's parseScoreMatrixFile (utilities.cpp:106)
  reference_goldens.json   (a) the known-answer vectors of the reference's own tests
                           (tests/tests.cu:116-368) with the expected values copied
                           from that file, and (b) reference alignSequenceCPU outputs
                           for the all-pairs sweep of tests/tests.cu:463-551, the
                           BASELINE configs C1/C2, synthetic mutate.py-style pairs and
                           edge cases -- every Response field, strings as sha256
                           (plus the literal strings when short).
"""
import hashlib
import itertools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
from oracle.oracle_py import Reference  # noqa: E402
import synth  # noqa: E402

REF = os.environ.get("SA_REF_DIR", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
MAX_LEN = 30000          # residues; longer data files are not committed
SWEEP_MAX_TEXT = 8000    # all-pairs sweep bound (the reference's tests use 20000)


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def record(aln, keep_strings=False):
    d = dict(score=aln.score, aln_len=aln.aln_len, start_text=aln.start_text,
             start_pattern=aln.start_pattern, sha_text=sha(aln.aligned_text),
             sha_pattern=sha(aln.aligned_pattern))
    if keep_strings or aln.aln_len <= 200:
        d["aligned_text"] = aln.aligned_text.decode()
        d["aligned_pattern"] = aln.aligned_pattern.decode()
    return d


def main():
    ref = Reference("O3")
    seqs = {}
    for kind, alpha in (("dna", 4), ("protein", 23)):
        d = os.path.join(REF, "data", kind)
        for f in sorted(os.listdir(d)):
            arr = ref.read_sequence(os.path.join(d, f), alpha)
            if len(arr) <= MAX_LEN:
                seqs[f"{kind}/{f}"] = arr
    np.savez_compressed(os.path.join(OUT, "sequences.npz"), **seqs)

    mats = {}
    for kind, alpha in (("dna", 4), ("protein", 23)):
        d = os.path.join(REF, "scoreMatrices", kind)
        for f in sorted(os.listdir(d)):
            try:
                mats[f"{kind}/{f}"] = ref.parse_score_matrix(os.path.join(d, f), alpha).tolist()
            except ValueError:
                pass  # blosum100.txt: trailing '-' tokens, the reference's parser fails too
    json.dump(mats, open(os.path.join(OUT, "matrices.json"), "w"))

    blast, b50, b62 = mats["dna/blast.txt"], mats["protein/blosum50.txt"], mats["protein/blosum62.txt"]
    enc = lambda s, a: ref.validate_and_transform(s.encode(), a)

    def run(name, mode, alpha, matrix, gap, text, pattern, cite=None, expect=None, keep=False,
            inline=False, matrix_name=None):
        if len(text) < len(pattern):          # utilities.cpp:225-230
            text, pattern = pattern, text
        t0 = time.time()
        aln = ref.align(mode, alpha, matrix, gap, text, pattern)
        g = dict(name=name, mode=mode, alpha=alpha, gap=gap, n=int(len(text)), m=int(len(pattern)),
                 matrix=matrix_name, ref=record(aln, keep))
        if cite:
            g["cite"] = cite
        if expect:
            g["expect"] = expect            # values copied from the reference's tests
            for k, v in expect.items():
                got = getattr(aln, k) if not k.startswith("aligned") else getattr(aln, k).decode()
                assert got == v, (name, k, got, v)
        if inline:
            g["text"] = np.asarray(text).tolist()
            g["pattern"] = np.asarray(pattern).tolist()
        print(f"{name:60s} score={aln.score:8d} len={aln.aln_len:7d} {time.time()-t0:6.2f}s")
        return g

    goldens = []
    S = seqs
    # ---- (a) the reference's own known-answer tests, tests/tests.cu ----
    kat = [
        ("tests.cu:DNA_01", 0, 4, blast, 5, S["dna/dna_01.txt"], S["dna/dna_02.txt"], "tests/tests.cu:119-133", dict(score=-4)),
        ("tests.cu:DNA_02", 0, 4, blast, 5, enc("GCCT", 4), enc("GGTC", 4), "tests/tests.cu:135-161", dict(score=-4)),
        ("tests.cu:DNA_03", 0, 4, blast, 5, enc("TTCGCCT", 4), enc("CTCGGTC", 4), "tests/tests.cu:163-189", dict(score=2)),
        ("tests.cu:DNA_04", 0, 4, blast, 5,
         enc("CATAAAACTCTCGGTCGGGCTTAGTACCAGGACCGGCGCACCAGAGTGTCAATCACGACCCTTCACACTTTGTGC", 4),
         enc("ATGAAGTTGTTCGCCTTACTTTTAATTCTACTCTCTCCTCGAGATTCGTCCGCTGAAAAATCTCTCAGCG", 4),
         "tests/tests.cu:191-232",
         dict(score=22,
              aligned_text="CATAAAACTCTCGGTCGGGCTTAGTACCAGGAC--CGGCGCACCA-GAG-TGTCAATCACGACCCTTCACACTTTGT--GC-",
              aligned_pattern="-ATGAAG-T-T-GTTCGC-CTTACTTTTAATTCTACT-CTCTCCTCGAGAT-TCG-TC-CG-C--TGAAAAATCTCTCAGCG")),
        ("tests.cu:DNA_05", 0, 4, blast, 5, S["dna/NC_018874.txt"], S["dna/GCA_003231495.txt"], "tests/tests.cu:234-248", dict(score=-5991)),
        ("tests.cu:PROTEIN_01", 0, 23, b50, 5,
         enc("MVLSPADKTNVKAAWGKVGAHAGEYGAEALERMFLSFPTTKTYFPHFDLSHGSAQVKGHGKKVADALTNAVAHVDDMPNALSALSDLHAHKLRVDPVNFKLLSHCLLVTLAAHLPAEFTPAVHASLDKFLASVSTVLTSKYR", 23),
         enc("MVLSGEDKSNIKAAWGKIGGHGAEYGAEALERMFASFPTTKTYFPHFDVSHGSAQVKGHGKKVADALASAAGHLDDLPGALSALSDLHAHKLRVDPVNFKLLSHCLLVTLASHHPADFTPAVHASLDKFLASVSTVLTSKYR", 23),
         "tests/tests.cu:251-292",
         dict(score=821,
              aligned_text="MVLSPADKTNVKAAWGKVGAHAGEYGAEALERMFLSFPTTKTYFPHFDLSHGSAQVKGHGKKVADALTNAVAHVDDMPNALSALSDLHAHKLRVDPVNFKLLSHCLLVTLAAHLPAEFTPAVHASLDKFLASVSTVLTSKYR",
              aligned_pattern="MVLSGEDKSNIKAAWGKIGGHGAEYGAEALERMFASFPTTKTYFPHFDVSHGSAQVKGHGKKVADALASAAGHLDDLPGALSALSDLHAHKLRVDPVNFKLLSHCLLVTLASHHPADFTPAVHASLDKFLASVSTVLTSKYR")),
        ("tests.cu:PROTEIN_02", 0, 23, b50, 5, S["protein/P02232.fasta"], S["protein/P03989.fasta"], "tests/tests.cu:294-308", dict(score=-597)),
        ("tests.cu:PROTEIN_03", 0, 23, b50, 5, S["protein/P05013.fasta"], S["protein/P07327.fasta"], "tests/tests.cu:310-324", dict(score=-423)),
        ("tests.cu:LOCAL_DNA_01", 1, 4, blast, 5, S["dna/GCA_003231495.txt"], S["dna/dna_01.txt"], "tests/tests.cu:330-350",
         dict(score=20, aligned_text="ACAC", aligned_pattern="ACAC", start_text=248, start_pattern=0)),
        ("tests.cu:LOCAL_PROTEIN_01", 1, 23, b50, 10, S["protein/P08519.fasta"], S["protein/P10635.fasta"], "tests/tests.cu:352-366",
         dict(score=57, start_text=4203, start_pattern=94)),
        # differential cases of the reference's GPU tests (inputs only; CPU is the truth)
        ("tests.cu:GPU_GLOBAL_PROTEIN_01", 0, 23, b50, 11, S["protein/P10635.fasta"], S["protein/P02232.fasta"], "tests/tests.cu:372-390", None),
        ("tests.cu:GPU_GLOBAL_PROTEIN_02", 0, 23, b50, 5, S["protein/P27895.fasta"], S["protein/P27895.fasta"], "tests/tests.cu:392-411", None),
    ]
    for name, mode, alpha, mat, gap, t, p, cite, exp in kat:
        files = [next((k for k, v in S.items() if v is x), None) for x in (t, p)]
        inline = None in files
        goldens.append(run(name, mode, alpha, mat, gap, t, p, cite, exp, keep=True, inline=inline,
                           matrix_name="dna/blast.txt" if alpha == 4 else "protein/blosum50.txt"))
        if not inline:
            goldens[-1]["files"] = files

    # ---- (b) BASELINE configs C1 / C2 ----
    goldens.append(run("C1:NC_018874xmutated NW blast g5", 0, 4, blast, 5, S["dna/NC_018874.txt"],
                       S["dna/mutated_NC_018874.txt"], "BASELINE.json configs[0]", matrix_name="dna/blast.txt"))
    goldens[-1]["files"] = ["dna/NC_018874.txt", "dna/mutated_NC_018874.txt"]
    goldens.append(run("C2:P33450xmutated SW blosum62 g5", 1, 23, b62, 5, S["protein/P33450.fasta"],
                       S["protein/mutated_P33450.fasta"], "BASELINE.json configs[1]", matrix_name="protein/blosum62.txt"))
    goldens[-1]["files"] = ["protein/P33450.fasta", "protein/mutated_P33450.fasta"]

    # ---- (c) all-pairs sweep, tests/tests.cu:463-551 (dna gap 11, protein gap 5, both modes) ----
    for kind, alpha, mat, mname, gap in (("dna", 4, blast, "dna/blast.txt", 11),
                                         ("protein", 23, b50, "protein/blosum50.txt", 5)):
        names = [k for k in S if k.startswith(kind + "/") and len(S[k]) <= SWEEP_MAX_TEXT]
        for a, b in itertools.combinations(names, 2):
            for mode in (0, 1):
                g = run(f"sweep:{a}x{b}:{'NW' if mode == 0 else 'SW'}", mode, alpha, mat, gap, S[a], S[b],
                        "tests/tests.cu:463-551", matrix_name=mname)
                g["files"] = [a, b]
                goldens.append(g)

    # ---- (d) one long real pair + other matrices ----
    g = run("long:NC_034972.1xmutated NW blast g5", 0, 4, blast, 5, S["dna/NC_034972.1.txt"],
            S["dna/mutated_NC_034972.1.txt"], "SURVEY.md 8c", matrix_name="dna/blast.txt")
    g["files"] = ["dna/NC_034972.1.txt", "dna/mutated_NC_034972.1.txt"]
    goldens.append(g)
    g = run("long:NC_034972.1xmutated SW dnaMat g3", 1, 4, mats["dna/dnaMat.txt"], 3, S["dna/NC_034972.1.txt"],
            S["dna/mutated_NC_034972.1.txt"], None, matrix_name="dna/dnaMat.txt")
    g["files"] = ["dna/NC_034972.1.txt", "dna/mutated_NC_034972.1.txt"]
    goldens.append(g)
    for mname in sorted(mats):
        if not mname.startswith("protein/"):
            continue
        for mode in (0, 1):
            g = run(f"matrix:{mname}:P04775xP07756:{'NW' if mode == 0 else 'SW'}", mode, 23, mats[mname], 7,
                    S["protein/P04775.fasta"], S["protein/P07756.fasta"], None, matrix_name=mname)
            g["files"] = ["protein/P04775.fasta", "protein/P07756.fasta"]
            goldens.append(g)

    # ---- (e) synthetic mutate.py-style pairs (seeded restatement, synth.py) ----
    for n, s1, s2, prot in ((1000, 1, 2, False), (5000, 3, 4, False), (20000, 5, 6, False),
                            (300, 7, 8, True), (2500, 9, 10, True)):
        t, p = synth.synthetic_pair(n, s1, s2, protein=prot)
        alpha, mat, mname = (23, b62, "protein/blosum62.txt") if prot else (4, blast, "dna/blast.txt")
        for mode in (0, 1):
            g = run(f"synth:n{n}:s{s1}/{s2}:{'prot' if prot else 'dna'}:{'NW' if mode == 0 else 'SW'}",
                    mode, alpha, mat, 5, t, p, "SURVEY.md 9.7", matrix_name=mname)
            g["synth"] = dict(n=n, seed_base=s1, seed_mut=s2, protein=prot)
            goldens.append(g)

    # ---- (f) edge cases ----
    edge = [
        ("edge:SW all-mismatch AAAAxTTT", 1, 4, blast, 5, enc("AAAA", 4), enc("TTT", 4)),
        ("edge:NW 1x1 match", 0, 4, blast, 5, enc("A", 4), enc("A", 4)),
        ("edge:NW 1x1 mismatch", 0, 4, blast, 5, enc("A", 4), enc("T", 4)),
        ("edge:SW 1x1 match", 1, 4, blast, 5, enc("G", 4), enc("G", 4)),
        ("edge:NW 9x1", 0, 4, blast, 5, enc("ACGTACGTA", 4), enc("G", 4)),
        ("edge:SW poly-A ties", 1, 4, blast, 5, enc("A" * 97, 4), enc("A" * 33, 4)),
        ("edge:NW poly-A ties", 0, 4, blast, 2, enc("A" * 97, 4), enc("A" * 33, 4)),
        ("edge:NW gap0", 0, 4, blast, 0, enc("ACGTTGCAAGCT" * 5, 4), enc("TGCATGCCAGT" * 4, 4)),
        ("edge:SW gap0", 1, 4, blast, 0, enc("ACGTTGCAAGCT" * 5, 4), enc("TGCATGCCAGT" * 4, 4)),
        ("edge:SW repeat ties", 1, 4, blast, 5, enc("ACGT" * 40, 4), enc("ACGT" * 9, 4)),
    ]
    for name, mode, alpha, mat, gap, t, p in edge:
        goldens.append(run(name, mode, alpha, mat, gap, t, p, None, keep=True, inline=True,
                           matrix_name="dna/blast.txt"))
    # asymmetric matrix pins the [pattern][text] orientation (alignSequenceCPU.cpp:172)
    asym = (np.arange(16, dtype=np.int32).reshape(4, 4) * 3 - 20)
    asym[np.arange(4), np.arange(4)] = 9
    for mode in (0, 1):
        t, p = synth.synthetic_pair(400, 11, 12)
        g = run(f"edge:asymmetric matrix:{'NW' if mode == 0 else 'SW'}", mode, 4, asym.ravel().tolist(), 4, t, p,
                "alignSequenceCPU.cpp:172", inline=True)
        g["matrix_values"] = asym.ravel().tolist()
        goldens.append(g)

    json.dump(goldens, open(os.path.join(OUT, "reference_goldens.json"), "w"), indent=0)
    print(len(goldens), "goldens written")


if __name__ == "__main__":
    main()

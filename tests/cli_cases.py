"""Command lines of the drop-in CLI test and the input files they read (written from the committed fixtures, so that the
GPU box -- which has no /root/reference -- sees the same bytes the reference CLI saw when the goldens were made)."""
import os

import numpy as np

import helpers

DNA = "ATCG"
PROTEIN = "ARNDCQEGHILKMFPSTWYVBZX"

SEQ_FILES = {          # file in the scratch directory -> (fixture, alphabet)
    "data/dna/dna_01.txt": ("dna/dna_01.txt", DNA), "data/dna/dna_02.txt": ("dna/dna_02.txt", DNA),
    "data/dna/NC_018874.txt": ("dna/NC_018874.txt", DNA), "data/dna/mutated_NC_018874.txt": ("dna/mutated_NC_018874.txt", DNA),
    "data/dna/GCA_003231495.txt": ("dna/GCA_003231495.txt", DNA),
    "data/protein/P33450.fasta": ("protein/P33450.fasta", PROTEIN), "data/protein/mutated_P33450.fasta": ("protein/mutated_P33450.fasta", PROTEIN),
    "data/protein/P08519.fasta": ("protein/P08519.fasta", PROTEIN), "data/protein/P10635.fasta": ("protein/P10635.fasta", PROTEIN),
    "data/protein/P02232.fasta": ("protein/P02232.fasta", PROTEIN), "data/protein/P03989.fasta": ("protein/P03989.fasta", PROTEIN),
}
MATRIX_FILES = {"scoreMatrices/dna/blast.txt": "dna/blast.txt", "scoreMatrices/protein/blosum50.txt": "protein/blosum50.txt",
                "scoreMatrices/protein/blosum62.txt": "protein/blosum62.txt"}

# {dev} is -c for the reference (goldens) and -g for the drop-in
CASES = [
    dict(name="smoke dna_01 x dna_02 NW", args=["{dev}", "--global", "data/dna/dna_01.txt", "data/dna/dna_02.txt"]),
    dict(name="C1 NC_018874 NW defaults", args=["{dev}", "-d", "data/dna/NC_018874.txt", "data/dna/mutated_NC_018874.txt"]),
    dict(name="C2 P33450 SW blosum62", args=["{dev}", "-p", "--local", "-s", "scoreMatrices/protein/blosum62.txt",
                                            "data/protein/P33450.fasta", "data/protein/mutated_P33450.fasta"]),
    dict(name="tests.cu local DNA_01 (pattern first: swapped)", args=["{dev}", "--local", "data/dna/dna_01.txt", "data/dna/GCA_003231495.txt"]),
    dict(name="tests.cu local PROTEIN_01 gap 10", args=["{dev}", "--protein", "--local", "--gap-penalty", "10",
                                                       "data/protein/P08519.fasta", "data/protein/P10635.fasta"]),
    dict(name="tests.cu global PROTEIN_02 default blosum50", args=["-p", "{dev}", "data/protein/P02232.fasta", "data/protein/P03989.fasta"]),
    dict(name="lower case + fasta header + junk", args=["{dev}", "--local", "data/dna/messy_a.txt", "data/dna/messy_b.txt"]),
    dict(name="zero-score local alignment prints nothing", args=["{dev}", "--local", "data/dna/all_a.txt", "data/dna/all_t.txt"]),
    dict(name="no arguments: usage", args=[]),
    dict(name="missing sequence file", args=["{dev}", "data/dna/dna_01.txt", "data/dna/nope.txt"]),
    dict(name="only one sequence", args=["{dev}", "data/dna/dna_01.txt"]),
    dict(name="gap penalty not an integer", args=["{dev}", "--gap-penalty", "five", "data/dna/dna_01.txt", "data/dna/dna_02.txt"]),
    dict(name="corrupt score matrix", args=["{dev}", "-s", "scoreMatrices/dna/corrupt.txt", "data/dna/dna_01.txt", "data/dna/dna_02.txt"]),
    dict(name="letter outside the alphabet", args=["{dev}", "data/dna/dna_01.txt", "data/dna/has_protein_letter.txt"]),
]


def write_inputs(root):
    seqs, mats = helpers.sequences(), helpers.matrices()
    for rel, (fixture, letters) in SEQ_FILES.items():
        path = os.path.join(root, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        s = np.frombuffer(letters.encode(), np.uint8)[seqs[fixture]].tobytes().decode()
        with open(path, "w") as f:
            f.write(">" + fixture + " written by tests/cli_cases.py\n")
            for k in range(0, len(s), 70):
                f.write(s[k:k + 70] + "\n")
    for rel, name in MATRIX_FILES.items():
        path = os.path.join(root, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        m = mats[name]
        a = int(round(len(m.ravel()) ** 0.5))
        with open(path, "w") as f:
            for row in m.reshape(a, a):
                f.write(" ".join(str(int(v)) for v in row) + "\n")
    extra = {
        "data/dna/messy_a.txt": ">hdr with > inside\nacgtACGTTTGACCA 123 **\n>second header acgt\nGGCATTA-CAG\n",
        "data/dna/messy_b.txt": "AcGtAcGtTtGaCcAgG\ncatta cag\n",
        "data/dna/all_a.txt": "AAAAAAAA\n", "data/dna/all_t.txt": "TTTTT\n",
        "data/dna/has_protein_letter.txt": "ACGTHACGT\n",
        "scoreMatrices/dna/corrupt.txt": "5 -4 -4 -4\n-4 5 x -4\n-4 -4 5 -4\n-4 -4 -4 5\n",
    }
    for rel, text in extra.items():
        with open(os.path.join(root, rel), "w") as f:
            f.write(text)

"""Seeded synthetic inputs in the style of the reference's mutate.py.

The reference's generator (mutate.py:42-59) is unseeded; these are seeded
re-statements that keep its per-character decision order:
    r1 < 0.05 -> delete; else r2 < 0.02 -> emit a random letter (the "insertion"
    replaces the character); else r3 < 0.05 -> substitute by a different letter;
    else keep.
``mutate_like_reference`` reproduces the exact Python ``random`` call sequence of
mutate.py and is used for the single-pair configs; the
vectorised numpy variant is used for the 1M-pair batch where a Python loop over
3e8 characters is not practical.
"""
from __future__ import annotations

import random

import numpy as np

DNA_LETTERS = ['A', 'T', 'C', 'G']                       # mutate.py:21
PROTEIN_LETTERS = ['A', 'R', 'N', 'D', 'C', 'Q', 'E', 'G', 'H', 'I', 'L', 'K', 'M', 'F', 'P',
                   'S', 'T', 'W', 'Y', 'V', 'B', 'Z', 'X']  # mutate.py:22-23

DELETION_CHANCE = 0.05       # mutate.py:4
INSERTION_CHANCE = 0.02      # mutate.py:5
SUBSTITUTION_CHANCE = 0.05   # mutate.py:6


def base_sequence_text(n: int, seed: int, letters=DNA_LETTERS) -> str:
    """n i.i.d. uniform letters, written like a sequence file: 70-char lines."""
    rng = random.Random(seed)
    s = ''.join(rng.choice(letters) for _ in range(n))
    return '\n'.join(s[i:i + 70] for i in range(0, n, 70)) + '\n'


def mutate_like_reference(file_text: str, seed: int, letters=DNA_LETTERS) -> str:
    """mutate.py:36-59 with a seeded generator (newlines go through the process too)."""
    rng = random.Random(seed)
    out = []
    for line in file_text.splitlines(keepends=True):
        if line.lstrip()[:1] == '>':
            out.append(line)
            continue
        for c in line:
            c = c.upper()
            if rng.random() < DELETION_CHANCE:
                pass
            elif rng.random() < INSERTION_CHANCE:
                out.append(rng.choice(letters))
            elif rng.random() < SUBSTITUTION_CHANCE:
                out.append(rng.choice([l for l in letters if l != c]))
            else:
                out.append(c)
    return ''.join(out)


def encode_letters(file_text: str, letters) -> np.ndarray:
    """Letters -> alphabet indices, skipping non-letters and FASTA '>' lines
    (what the reference's validateAndTransform does, utilities.cpp:31-63)."""
    lut = np.full(256, 255, np.uint8)
    for i, l in enumerate(letters):
        lut[ord(l)] = i
        lut[ord(l.lower())] = i
    keep = []
    ignore = False
    for ch in file_text:
        if not ignore and ch == '>':
            ignore = True
        elif ignore and ch == '\n':
            ignore = False
        if ignore:
            continue
        if ch.isalpha() and ch.isascii():
            keep.append(ord(ch))
    arr = lut[np.asarray(keep, dtype=np.uint8)] if keep else np.zeros(0, np.uint8)
    if (arr == 255).any():
        raise ValueError("letter not in alphabet")
    return arr


def synthetic_pair(n: int, seed_base: int, seed_mut: int, protein: bool = False):
    """(text, pattern) index arrays; text is the longer one like parseArguments
    arranges (utilities.cpp:225-230)."""
    letters = PROTEIN_LETTERS if protein else DNA_LETTERS
    base_txt = base_sequence_text(n, seed_base, letters if not protein else letters[:22])
    mut_txt = mutate_like_reference(base_txt, seed_mut, letters)
    a, b = encode_letters(base_txt, letters), encode_letters(mut_txt, letters)
    return (a, b) if len(a) >= len(b) else (b, a)


def mutate_indices_numpy(base: np.ndarray, rng: np.random.Generator, alpha: int) -> np.ndarray:
    """Vectorised mutate.py process on an index-encoded sequence."""
    n = len(base)
    r1, r2, r3 = rng.random(n), rng.random(n), rng.random(n)
    delete = r1 < DELETION_CHANCE
    insert = ~delete & (r2 < INSERTION_CHANCE)
    subst = ~delete & ~insert & (r3 < SUBSTITUTION_CHANCE)
    out = base.copy()
    out[insert] = rng.integers(0, alpha, int(insert.sum()), dtype=np.uint8)
    # a different letter: add 1..alpha-1 modulo alpha
    out[subst] = (out[subst] + rng.integers(1, alpha, int(subst.sum()), dtype=np.uint8)) % alpha
    return out[~delete]


def synthetic_batch(n_pairs: int, seed: int = 2024, lo: int = 250, hi: int = 350,
                    alpha: int = 23, base_letters: int = 22, chunk: int = 65536):
    """C4-style batch: base length uniform in [lo, hi], residues
    uniform over the first ``base_letters`` letters (tests/benchmarks.cu:35-38),
    partner = mutate.py-style mutation over the full alphabet.  Returns
    concatenated CSR arrays (text, text_off, pattern, pattern_off) with
    text >= pattern per pair."""
    rng = np.random.default_rng(seed)
    texts, pats = [], []
    toff = np.zeros(n_pairs + 1, np.int64)
    poff = np.zeros(n_pairs + 1, np.int64)
    done = 0
    while done < n_pairs:
        k = min(chunk, n_pairs - done)
        lens = rng.integers(lo, hi + 1, k)
        tot = int(lens.sum())
        base = rng.integers(0, base_letters, tot, dtype=np.uint8)
        r1, r2, r3 = rng.random(tot), rng.random(tot), rng.random(tot)
        delete = r1 < DELETION_CHANCE
        insert = ~delete & (r2 < INSERTION_CHANCE)
        subst = ~delete & ~insert & (r3 < SUBSTITUTION_CHANCE)
        mut = base.copy()
        mut[insert] = rng.integers(0, alpha, int(insert.sum()), dtype=np.uint8)
        mut[subst] = (mut[subst] + rng.integers(1, alpha, int(subst.sum()), dtype=np.uint8)) % alpha
        keep = ~delete
        ends = np.cumsum(lens)
        starts = ends - lens
        kept_cum = np.concatenate(([0], np.cumsum(keep)))
        mlens = kept_cum[ends] - kept_cum[starts]
        mut = mut[keep]
        # guard: a partner must not be empty
        if (mlens == 0).any():
            raise RuntimeError("empty mutated sequence; raise lo")
        texts.append(base)
        pats.append(mut)
        toff[done + 1:done + k + 1] = toff[done] + ends
        poff[done + 1:done + k + 1] = poff[done] + np.cumsum(mlens)
        done += k
    return (np.concatenate(texts), toff, np.concatenate(pats), poff)

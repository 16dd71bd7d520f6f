#!/usr/bin/env python
"""Writes the inputs of the reference pin kit (tools/make_reference_fixtures.sh): a small genome and reads that
exercise every rule SURVEY.md Appendix A marks [upstream, from memory] -- ties in the window minimum (low
complexity), both strands, N / IUPAC / lower case in index and query, lengths around k and k + w - 1, prefix
trimming, k > 32 -- sized so that the reference's outputs fit in the repository (a few hundred KB gzipped).

usage: python tests/gen_fixture_inputs.py OUTDIR        (deterministic: numpy PCG64, fixed seeds)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import py_oracle as P  # noqa: E402  (pure Python; used to FIND inputs that separate hypotheses, not to judge them)

ACGT = np.frombuffer(b"ACGT", np.uint8)
COMP = np.zeros(256, np.uint8)
for a, b in zip(b"ACGTacgtNn", b"TGCAtgcaNn"):
    COMP[a] = b


def rand_seq(rng, n):
    return ACGT[rng.integers(0, 4, n)]


def revcomp(s):
    return COMP[s[::-1]]


def genome(rng):
    c1 = rand_seq(rng, 120_000)
    c2 = rand_seq(rng, 40_000)
    iupac = np.frombuffer(b"NRYSWKMBDHVn", np.uint8)
    pos = rng.integers(0, len(c2), 300)
    c2[pos] = iupac[rng.integers(0, len(iupac), 300)]
    lower = rng.integers(0, len(c2) - 200, 20)
    for p in lower:                                  # soft-masked stretches
        c2[p:p + 150] = np.frombuffer(bytes(c2[p:p + 150]).lower(), np.uint8)
    parts = []
    for unit, reps in ((b"A", 400), (b"AC", 300), (b"ACGT", 200), (b"ACGTTGCAAT", 150), (b"T", 350), (b"GGC", 250)):
        parts.append(np.tile(np.frombuffer(unit, np.uint8), reps))
        parts.append(rand_seq(rng, 700))
    c3 = np.concatenate(parts)
    ties = tie_windows(rng)
    c4 = np.concatenate([np.concatenate([t, rand_seq(rng, 60)]) for t in ties])
    return [("chr_random", c1), ("chr_iupac_softmasked", c2), ("chr_low_complexity", c3), ("chr_tie_windows", c4)], ties


def tie_windows(rng, want=24, k=31, w=15):
    """45-mers (exactly one window) whose two smallest k-mer hashes agree in the upper 16 bits and disagree below, with
    the full 32-bit order opposite to the position order: "compare the top 16 bits only" and "compare all 32 bits" pick
    different k-mers there.  Found by search under the working hypothesis' hash (vectorised closed form of
    py_oracle.nthash); which k-mer the REFERENCE picks is what the fixture records."""
    F = np.array(P.F, np.uint32)
    l = k + w - 1
    out = []

    def rotl(x, r):
        r %= 32
        return x if r == 0 else ((x << np.uint32(r)) | (x >> np.uint32(32 - r)))

    while len(out) < want:
        n = 200_000
        codes_acgt = rng.integers(0, 4, (n, l))
        seqs = ACGT[codes_acgt]
        codes = (seqs >> 1) & 3
        h = np.zeros((n, w), np.uint32)
        for p in range(w):
            fw = np.zeros(n, np.uint32)
            rc = np.zeros(n, np.uint32)
            for i in range(k):
                c = codes[:, p + i]
                fw ^= rotl(F[c], k - 1 - i)
                rc ^= rotl(F[c ^ 2], i)
            h[:, p] = fw + rc
        top = h >> np.uint32(16)
        m = top.min(axis=1, keepdims=True)
        is_min = top == m
        first = is_min.argmax(axis=1)
        full_arg = np.where(is_min, h, np.uint32(0xFFFFFFFF)).argmin(axis=1)
        for j in np.nonzero((is_min.sum(axis=1) >= 2) & (full_arg != first))[0]:
            seq = seqs[j].copy()
            assert P.nthash(P.codes_of(bytes(seq)), int(first[j]), k) == int(h[j, first[j]])
            out.append(seq)
            if len(out) == want:
                break
    return out


def mutate(rng, s, rate):
    s = s.copy()
    m = rng.random(len(s)) < rate
    s[m] = ACGT[rng.integers(0, 4, int(m.sum()))]
    return s


def sample(rng, contigs, n):
    g = contigs[rng.integers(0, len(contigs))][1]
    n = min(n, len(g))
    p = int(rng.integers(0, len(g) - n + 1))
    s = g[p:p + n].copy()
    return revcomp(s) if rng.random() < 0.5 else s


def write_fasta(path, recs, width=0):
    with open(path, "wb") as f:
        for name, s in recs:
            f.write(b">" + name.encode() + b"\n")
            b = bytes(s)
            if width:
                for i in range(0, len(b), width):
                    f.write(b[i:i + width] + b"\n")
            else:
                f.write(b + b"\n")


def write_fastq(path, recs):
    with open(path, "wb") as f:
        for name, s in recs:
            f.write(b"@" + name.encode() + b"\n" + bytes(s) + b"\n+\n" + b"I" * len(s) + b"\n")


def main(out):
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(20261018)
    contigs, ties = genome(rng)
    write_fasta(os.path.join(out, "genome.fa"), contigs, width=80)     # multi-line FASTA, like real references
    single = []
    lens = list(range(25, 64)) + [150] * 500 + [int(x) for x in rng.integers(64, 420, 500)]
    for i, n in enumerate(lens):
        kind = rng.random()
        if kind < 0.6:
            s = mutate(rng, sample(rng, contigs, n), 0.02)
        elif kind < 0.8:
            s = rand_seq(rng, n)
        elif kind < 0.9:
            s = mutate(rng, sample(rng, contigs, n), 0.02)
            s[rng.integers(0, len(s), max(1, len(s) // 60))] = ord("N")
        else:
            s = np.frombuffer(bytes(sample(rng, contigs, n)).lower(), np.uint8)
        single.append((f"s{i}_len{len(s)}", s))
    for i, t in enumerate(ties):                     # one-window reads, both strands
        single.append((f"tie{i}_fw", t))
        single.append((f"tie{i}_rc", revcomp(t)))
    write_fastq(os.path.join(out, "reads_single.fq"), single)
    r1, r2 = [], []
    for i in range(600):
        g = contigs[0][1] if rng.random() < 0.8 else contigs[2][1]
        ins = int(rng.integers(160, 420))
        p = int(rng.integers(0, len(g) - ins))
        frag = g[p:p + ins]
        a, b = frag[:150].copy(), revcomp(frag[-150:])
        if rng.random() < 0.2:
            a, b = rand_seq(rng, 150), rand_seq(rng, 150)
        a, b = mutate(rng, a, 0.01), mutate(rng, b, 0.01)
        if rng.random() < 0.5:
            a, b = b, a
        r1.append((f"p{i}/1", a))
        r2.append((f"p{i}/2", b))
    write_fastq(os.path.join(out, "reads_r1.fq"), r1)
    write_fastq(os.path.join(out, "reads_r2.fq"), r2)
    long_reads = []
    for i in range(30):
        n = int(np.clip(rng.gamma(2.0, 4000.0), 1100, 30000))
        s = mutate(rng, sample(rng, contigs[:1], n), 0.05) if rng.random() < 0.6 else rand_seq(rng, n)
        long_reads.append((f"l{i}_len{len(s)}", s))
    write_fastq(os.path.join(out, "reads_long.fq"), long_reads)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/reference_c1")

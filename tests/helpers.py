"""Shared synthetic-data helpers for the test-suite (seeded; numpy only)."""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGTacgtNn", b"TGCAtgcaNn"):
    _COMP[_a] = _b


def random_genome(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return ACGT[rng.integers(0, 4, n)]


def revcomp(a: np.ndarray) -> np.ndarray:
    return _COMP[a[::-1]]


def sample_reads(genome: np.ndarray, n_reads: int, read_len, seed: int, frac_genome=0.5, sub_rate=0.01,
                 n_rate=0.001, lower_rate=0.0):
    """-> list of uint8 arrays.  read_len: int or (lo, hi) range."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_reads):
        ln = read_len if isinstance(read_len, int) else int(rng.integers(read_len[0], read_len[1] + 1))
        if rng.random() < frac_genome and len(genome) > ln:
            p = int(rng.integers(0, len(genome) - ln + 1))
            r = genome[p:p + ln].copy()
            if rng.random() < 0.5:
                r = revcomp(r)
            m = rng.random(ln) < sub_rate
            r[m] = ACGT[rng.integers(0, 4, int(m.sum()))]
        else:
            r = ACGT[rng.integers(0, 4, ln)]
        if n_rate > 0 and rng.random() < n_rate * 1000 * 0.001 + n_rate and ln > 0:
            r[int(rng.integers(0, ln))] = ord("N")
        if lower_rate > 0:
            m = rng.random(ln) < lower_rate
            r[m] |= 0x20
        out.append(r)
    return out


def concat(records):
    off = np.zeros(len(records) + 1, np.uint64)
    if records:
        off[1:] = np.cumsum([len(r) for r in records], dtype=np.uint64)
    bases = np.concatenate(records) if records else np.zeros(0, np.uint8)
    return np.ascontiguousarray(bases, np.uint8), off

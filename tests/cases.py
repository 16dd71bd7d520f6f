"""Parity cases shared by the emulator tests (CPU) and the GPU tests (through the C ABI).

Each case is a dict: name, records (list of uint8 arrays), paired, prefix, abs, rel, deplete, and
index_records (sequences the index is built from with the ORACLE's index-flavour extraction).
"""
from __future__ import annotations

import numpy as np

import helpers as H


def _b(s):
    return np.frombuffer(s.encode() if isinstance(s, str) else s, np.uint8).copy()


def make_cases():
    cases = []
    g = H.random_genome(120_000, 11)

    def add(name, records, paired=False, prefix=0, abs_=2, rel=0.01, deplete=False, index_records=None, extra_keys=None):
        cases.append(dict(name=name, records=records, paired=paired, prefix=prefix, abs=abs_, rel=rel, deplete=deplete,
                          index_records=index_records if index_records is not None else [g], extra_keys=extra_keys))

    r150 = H.sample_reads(g, 2000, 150, 21)
    add("single_150_search", r150)
    add("single_150_deplete", r150, deplete=True)
    add("paired_150_deplete", r150, paired=True, deplete=True)
    add("paired_150_search_abs1", r150, paired=True, abs_=1, rel=0.0)
    ragged = H.sample_reads(g, 2500, (0, 420), 22, n_rate=0.05, lower_rate=0.1)
    add("ragged_0_420_N_lower", ragged)
    add("ragged_paired", ragged, paired=True, deplete=True)
    add("ragged_prefix_80", ragged, prefix=80)
    add("ragged_prefix_20_below_k", ragged, prefix=20)
    add("rel_threshold_half", r150[:500], rel=0.5)
    add("abs_threshold_9", r150[:500], abs_=9, rel=0.0)
    # lengths around k and l = k + w - 1
    edge = [g[100:100 + n].copy() for n in (0, 1, 30, 31, 32, 44, 45, 46, 59, 60, 61)] * 3
    add("lengths_around_k_and_l", edge)
    add("lengths_around_k_and_l_paired", edge + [g[5:50].copy()], paired=True)
    # trailing newline on records (src/filter_common.rs:229)
    nl = [np.concatenate([r, _b("\n")]) for r in r150[:300]]
    add("trailing_newline", nl)
    add("trailing_newline_prefix", nl, prefix=151)
    # many tiny records: more than MAXR records inside one tile
    tiny = [g[i * 7:i * 7 + (i % 9)].copy() for i in range(3000)] + r150[:50]
    add("tiny_records", tiny)
    add("tiny_records_paired", tiny, paired=True)
    add("empty_records_only", [np.zeros(0, np.uint8)] * 40)
    # low complexity / repeats: duplicate minimizers inside a record and across mates
    rep = [_b("ACGT" * 40), _b("A" * 150), _b("AC" * 75), _b("ACGTTGCA" * 19), _b("A" * 60 + "C" * 60 + "A" * 60)]
    add("low_complexity", rep * 20, index_records=[_b("ACGT" * 100), _b("A" * 200), g[:5000]])
    mates = []
    rng = np.random.default_rng(5)
    for _ in range(400):  # overlapping mates share minimizers
        p = int(rng.integers(0, len(g) - 400))
        ins = int(rng.integers(120, 320))
        mates.append(g[p:p + 150].copy())
        mates.append(H.revcomp(g[p + ins - 150:p + ins]))
    add("overlapping_mates", mates, paired=True, deplete=True)
    add("overlapping_mates_search", mates, paired=True)
    # every window emits a pick (poly-A: non-canonical windows take the rightmost minimum): a tile
    # with more picks than one pass holds must be split
    add("dense_picks_polyA", [_b("A" * 150)] * 200 + r150[:30] + [_b("A" * 1000)] * 8, index_records=[_b("A" * 200), g[:3000]])
    add("dense_picks_polyA_paired", [_b("A" * 150)] * 200 + [_b("T" * 151)] * 100, paired=True, deplete=True,
        index_records=[_b("A" * 200), g[:3000]])
    # all-N and mixed garbage bytes
    junk = [_b("N" * 150), _b("ACGT" * 20 + "N" + "ACGT" * 20), np.arange(256, dtype=np.uint8), _b("acgtn" * 40)]
    add("non_acgt_bytes", junk * 10 + r150[:100])
    # 2 x 250 / 2 x 300 pairs (longer short units)
    r300 = H.sample_reads(g, 600, (240, 301), 23)
    add("paired_250_300", r300, paired=True, deplete=True)
    # units right at the short/long boundary
    r1024 = [g[i * 100:i * 100 + n].copy() for i, n in enumerate((1024, 1023, 1000, 512, 512, 1024, 900))]
    add("units_up_to_1024", r1024 + r150[:64])
    # the EMPTY sentinel as a key of the index, and keys with tiny values
    add("sentinel_key_in_index", r150[:200], extra_keys=np.array([0xFFFFFFFFFFFFFFFF, 0, 1, 250, 251, 65535, 65536], np.uint64))
    return cases


def make_long_cases():
    g = H.random_genome(300_000, 12)
    cases = []
    long_reads = H.sample_reads(g, 120, (1025, 30_000), 31, sub_rate=0.05)
    cases.append(dict(name="long_reads_search", records=long_reads, paired=False, prefix=0, abs=2, rel=0.01, deplete=False,
                      index_records=[g], extra_keys=None))
    mixed = []
    short = H.sample_reads(g, 300, 150, 32)
    for i, r in enumerate(long_reads[:40]):
        mixed += short[i * 5:(i + 1) * 5] + [r]
    cases.append(dict(name="mixed_short_long", records=mixed, paired=False, prefix=0, abs=2, rel=0.01, deplete=True,
                      index_records=[g], extra_keys=None))
    cases.append(dict(name="mixed_short_long_paired", records=mixed[:len(mixed) // 2 * 2], paired=True, prefix=0, abs=2,
                      rel=0.01, deplete=False, index_records=[g], extra_keys=None))
    rep = [np.tile(np.frombuffer(b"ACGTTGCAAT", np.uint8), 3000), np.tile(g[:700], 12)]
    cases.append(dict(name="long_repeats", records=rep, paired=False, prefix=0, abs=1, rel=0.0, deplete=False,
                      index_records=[g], extra_keys=None))
    cases.append(dict(name="long_prefix", records=long_reads[:30], paired=False, prefix=5000, abs=2, rel=0.01,
                      deplete=False, index_records=[g], extra_keys=None))
    return cases

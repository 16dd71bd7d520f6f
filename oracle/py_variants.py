"""The 64 neighbouring hypotheses about the un-vendored upstream arithmetic (simd-minimizers 1.3.0 / packed-seq 3.2.1),
as one parametrised pure-Python restatement.

TEST INFRASTRUCTURE ONLY (tests/test_hypothesis_sweep.py, tests/test_reference_fixtures.py).  SURVEY.md A.2-A.3 rest on
recollection of the upstream source; this module makes the claim "the reference's own tests reject only one family of
16 variants, the other 48 pass" reproducible, and lets a reference-derived fixture (tools/make_reference_fixtures.sh)
say which single variant is the reference's.  The working hypothesis (what oracle/deacon_oracle.c and the CUDA kernels
implement) is Variant() with every field at its default.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

from . import py_oracle as P

M32 = 0xFFFFFFFF
# the four classic ntHash seeds in their usual A, C, G, T order
SEEDS64 = {"A": 0x3C8BFBB395C60474, "C": 0x3193C18562A02B4C, "G": 0x20323ED082572324, "T": 0x295549F54BE24456}


@dataclass(frozen=True)
class Variant:
    table_by_code: bool = True   # True: the A,C,G,T-ordered table is indexed by the packing code (A,C,T,G), so code 2 (T) takes the
                                 # "G" seed and code 3 (G) the "T" seed; False: every base takes its own seed
    low32: bool = True           # low / high 32 bits of the 64-bit seeds
    add: bool = True             # canonical hash = fw + rc (wrapping) / fw ^ rc
    top16: bool = True           # window comparison on the upper 16 bits only / on all 32 bits
    tg_majority: bool = True     # canonical window: more T|G than A|C / more A|C than T|G
    left_on_canonical: bool = True   # canonical window takes the leftmost minimum (else the rightmost), the other strand the opposite

    def name(self):
        return "".join("1" if v else "0" for v in (self.table_by_code, self.low32, self.add, self.top16, self.tg_majority,
                                                   self.left_on_canonical))


ALL = [Variant(*bits) for bits in itertools.product((True, False), repeat=6)]
WORKING = Variant()


def seed_table(v: Variant):
    order = "ACGT"
    t = [SEEDS64[c] for c in order]                      # table as written upstream: A, C, G, T
    if v.table_by_code:
        by_code = t                                      # code 0..3 = A, C, T, G reads entries 0..3 = "A", "C", "G", "T" seeds
    else:
        by_code = [SEEDS64["A"], SEEDS64["C"], SEEDS64["T"], SEEDS64["G"]]
    return [(s & M32) if v.low32 else (s >> 32) for s in by_code]


def nthash(v: Variant, F, codes, p, k):
    fw = rc = 0
    for i in range(k):
        fw ^= P.rotl32(F[codes[p + i]], k - 1 - i)
        rc ^= P.rotl32(F[codes[p + i] ^ 2], i)
    return ((fw + rc) & M32) if v.add else (fw ^ rc)


def minimizer_positions(v: Variant, codes, k, w):
    n, l = len(codes), k + w - 1
    if n < l:
        return []
    F = seed_table(v)
    keys = [nthash(v, F, codes, p, k) for p in range(n - k + 1)]
    if v.top16:
        keys = [x >> 16 for x in keys]
    out, prev = [], None
    for j in range(n - l + 1):
        win = keys[j:j + w]
        m = min(win)
        left = j + win.index(m)
        right = j + (w - 1 - win[::-1].index(m))
        tg = sum((c >> 1) & 1 for c in codes[j:j + l])
        canonical = (2 * tg > l) if v.tg_majority else (2 * tg < l)
        pick = (left if canonical else right) if v.left_on_canonical else (right if canonical else left)
        if prev is None or pick != prev:
            out.append(pick)
        prev = pick
    return out


def extract_filter(v: Variant, seq: bytes, k=31, w=15, prefix_len=0):
    """src/filter_common.rs:211-310 under variant v -> (hashes, positions)."""
    if len(seq) < k:
        return [], []
    eff = seq[:prefix_len] if (prefix_len > 0 and len(seq) > prefix_len) else seq
    if eff.endswith(b"\n"):
        eff = eff[:-1]
    codes = P.codes_of(eff)
    hs, ps = [], []
    for p in minimizer_positions(v, codes, k, w):
        if all(b in P.ACGT for b in eff[p:p + k]):
            ps.append(p)
            hs.append(P.kmer_hash(codes, p, k))
    return hs, ps


def extract_index(v: Variant, seq: bytes, k=31, w=15):
    """src/minimizers.rs:125-191 (entropy threshold 0) under variant v."""
    if len(seq) < k:
        return []
    mapped = bytes(P._IUPAC.get(b, ord("C")) for b in seq)
    codes = P.codes_of(mapped)
    return [P.kmer_hash(codes, p, k) for p in minimizer_positions(v, codes, k, w) if all(b in P.ACGT for b in seq[p:p + k])]


def run_kat(v: Variant, case: dict) -> bool:
    """One behavioural known-answer test of the reference (tests/golden/reference_kats.json) under variant v."""
    k, w = case["k"], case["w"]
    index = set()
    for r in case["ref"]:
        index.update(extract_index(v, r.encode(), k, w))
    if "reads" in case:
        units = [[r] for r in case["reads"]]
    else:
        units = [[a, b] for a, b in zip(case["reads1"], case["reads2"])]
    keeps = []
    for unit in units:
        hashes = []
        for r in unit:
            hashes += extract_filter(v, r.encode(), k, w, case.get("prefix", 0))[0]
        keeps.append(int(P.should_keep(index, hashes, case["abs"], case["rel"], case["deplete"])[0]))
    return keeps == case["expect_keep"]

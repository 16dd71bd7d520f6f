"""CPU tests: the oracle against every golden vector and known-answer test available for this path."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import py_oracle as P

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def test_xxh3_against_xxhash_module_vectors():
    for v in _load("xxh3_vectors.json")["vectors"]:
        val = int(v["value"], 16)
        got = O.xxh3_u64(val) if v["bytes"] == 8 else O.xxh3_u128(val)
        assert got == int(v["hash"], 16)
        got_py = P.xxh3_u64(val) if v["bytes"] == 8 else P.xxh3_u128(val)
        assert got_py == int(v["hash"], 16)


def test_xxh3_live_against_xxhash_module():
    xxhash = pytest.importorskip("xxhash")
    rng = np.random.default_rng(1)
    for v in rng.integers(0, 2**63, 2000, dtype=np.uint64):
        v = int(v) * 2 + 1
        assert O.xxh3_u64(v & (2**64 - 1)) == xxhash.xxh3_64_intdigest((v & (2**64 - 1)).to_bytes(8, "little"))
        big = (v * 0x9E3779B97F4A7C15) & (2**114 - 1)
        assert O.xxh3_u128(big) == xxhash.xxh3_64_intdigest(big.to_bytes(16, "little"))


def test_survey_hypothesis_vectors():
    d = _load("hypothesis_vectors.json")
    for c in d["extract_filter"]:
        h, p = O.extract_filter(c["seq"], c["k"], c["w"])
        assert list(p) == c["positions"]
        assert [hex(int(x)) for x in h] == c["hashes"]
    for c in d["canonical_values"]:
        codes = P.codes_of(c["kmer"].encode())
        assert hex(P.canonical_value(codes, 0, len(codes))) == c["value"]
        assert hex(P.kmer_hash(codes, 0, len(codes))) == c["hash"]
    for c in d["xxh3"]:
        val = int(c["value"], 16)
        assert hex(O.xxh3_u64(val) if c["bytes"] == 8 else O.xxh3_u128(val)) == c["hash"]


def test_extract_vectors_from_independent_restatement():
    for c in _load("extract_vectors.json")["cases"]:
        h, p = O.extract_filter(c["seq"], c["k"], c["w"], c["prefix"])
        assert [hex(int(x)) for x in h] == c["filter_hashes"], c["kind"]
        assert list(map(int, p)) == c["filter_positions"]
        hi = O.extract_index(c["seq"], c["k"], c["w"])
        assert [hex(int(x)) for x in hi] == c["index_hashes"], c["kind"]


def test_c_oracle_vs_python_restatement_random():
    rng = np.random.default_rng(7)
    for _ in range(60):
        n = int(rng.integers(0, 400))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTACGTACGTNacgt", np.uint8), n).tolist())
        k, w = [(31, 15), (21, 11), (5, 3), (41, 15), (31, 1)][int(rng.integers(0, 5))]
        prefix = int(rng.choice([0, 0, 60]))
        h, p = O.extract_filter(seq, k, w, prefix)
        hp, pp = P.extract_filter(seq, k, w, prefix)
        assert list(map(int, h)) == hp and list(map(int, p)) == pp
        assert list(map(int, O.extract_index(seq, k, w))) == P.extract_index(seq, k, w)


def test_rolling_equals_bruteforce_positions():
    rng = np.random.default_rng(8)
    for _ in range(40):
        n = int(rng.integers(0, 600))
        # few distinct symbols -> many ties in the 16-bit keys
        codes = rng.integers(0, int(rng.integers(1, 5)), n).astype(np.uint8)
        for k, w in ((31, 15), (7, 4), (33, 9)):
            a = O.minimizer_positions(codes, k, w)
            b = O.minimizer_positions(codes, k, w, brute=True)
            assert np.array_equal(a, b)
            assert list(a) == P.minimizer_positions(list(map(int, codes)), k, w)


def test_strand_invariance_of_hash_sets():
    rng = np.random.default_rng(9)
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    for _ in range(30):
        s = rng.choice(np.frombuffer(b"ACGT", np.uint8), 300)
        rc = np.array([comp[int(x)] for x in s[::-1]], np.uint8)
        assert sorted(map(int, O.extract_filter(s)[0])) == sorted(map(int, O.extract_filter(rc)[0]))


def _run_kat(case, index_fn, filter_fn):
    idx = index_fn([c.encode() for c in case["ref"]], case["k"], case["w"])
    if "reads" in case:
        recs, paired = [r.encode() for r in case["reads"]], False
    else:
        recs, paired = [], True
        for a, b in zip(case["reads1"], case["reads2"]):
            recs += [a.encode(), b.encode()]
    return filter_fn(idx, recs, paired, case)


def test_reference_behavioural_known_answers():
    """tests/filter_tests.rs scenarios, run through the oracle."""
    for case in _load("reference_kats.json")["cases"]:
        def index_fn(refs, k, w):
            return O.index_build(refs, k, w)

        def filter_fn(idx, recs, paired, c):
            bases, off = O.concat_records(recs)
            return O.filter_batch(idx, bases, off, paired=paired, k=c["k"], w=c["w"], abs_thr=c["abs"],
                                  rel_thr=c["rel"], deplete=c["deplete"])
        keep, hits, total = _run_kat(case, index_fn, filter_fn)
        assert list(map(int, keep)) == case["expect_keep"], case["name"]
        if "expect_hits" in case:
            assert list(map(int, hits)) == case["expect_hits"], case["name"]


def test_required_hits_rule():
    # src/filter_common.rs:84-96 incl. quirk C.11 (rel = 0 still requires 1 when total > 0)
    assert O.required_hits(2, 0.01, 0) == 2
    assert O.required_hits(1, 0.0, 10) == 1
    assert O.required_hits(0, 0.0, 10) == 1
    assert O.required_hits(0, 0.0, 0) == 0
    assert O.required_hits(2, 0.01, 250) == 3      # round(2.5) = 3: half away from zero
    assert O.required_hits(2, 0.01, 249) == 2
    assert O.required_hits(1, 0.5, 3) == 2         # round(1.5) = 2
    assert O.required_hits(1, float("nan"), 3) == 1
    for total in range(0, 400):
        for rel in (0.0, 0.01, 0.015, 0.1, 0.5, 1.0):
            assert O.required_hits(2, rel, total) == P.required_hits(2, rel, total)
    assert O.meets_criteria(0, 0, 2, 0.01, True) and not O.meets_criteria(0, 0, 2, 0.01, False)


def test_entropy_reference_unit_test_ranges():
    # src/minimizers.rs:252-386: ranges asserted by the reference's own unit tests
    assert O.scaled_entropy(b"AAAAAAAAAA") == 0.0
    assert abs(O.scaled_entropy(b"ACGTACGTAC") - 0.985) < 0.02
    assert O.scaled_entropy(b"ACGT") == 1.0                      # k < 10 always passes
    assert 0.45 < O.scaled_entropy(b"AAAAACCCCC") < 0.55
    assert O.scaled_entropy(b"ACGTACGTACGTACGTACGTACGTACGTACG") > 0.95


def test_idx_codec_roundtrip_and_short_encodings():
    keys = np.array([0, 1, 250, 251, 65535, 65536, 2**32 - 1, 2**32, 2**64 - 1, 0x1234567890ABCDEF], np.uint64)
    blob = O.idx_encode(keys, 31, 15)
    assert blob[:3] == bytes([2, 31, 15]) and blob[3] == len(keys)
    assert len(blob) == 4 + 1 + 1 + 1 + 3 + 3 + 5 + 5 + 9 + 9 + 9
    ver, k, w, got = O.idx_decode(blob)
    assert (ver, k, w) == (2, 31, 15) and np.array_equal(got, keys)
    bad = bytes([3]) + blob[1:]
    with pytest.raises(ValueError, match="Unsupported index format version"):   # src/index.rs:34-40
        O.idx_decode(bad)


def test_index_set_and_batch_operators():
    rng = np.random.default_rng(3)
    keys = rng.integers(0, 2**63, 5000, dtype=np.uint64)
    s = O.IndexSet(np.concatenate([keys, keys[:100], np.array([0], np.uint64)]), threads=4)
    assert len(s) == len(np.unique(keys)) + 1 and int(keys[5]) in s and 12345 not in s and 0 in s
    # lookup_batch == per-record python
    hashes = np.concatenate([keys[:30], keys[:10], rng.integers(0, 2**63, 20, dtype=np.uint64)])
    off = np.array([0, 30, 40, 40, 60], np.uint64)
    keep, hits, total = O.lookup_batch(s, hashes, off, abs_thr=2, rel_thr=0.01, deplete=False, threads=2)
    assert list(hits) == [30, 10, 0, 0] and list(total) == [30, 10, 0, 20] and list(keep) == [1, 1, 0, 0]

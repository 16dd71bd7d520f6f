"""CPU tests of the generic (k, w) code (dcn_generic.cuh compiled for the host): B3 extraction in
both flavours and the any-(k, w) filter must agree bit-for-bit with the oracle."""
import json
import os

import numpy as np
import pytest

import emu_harness as E
import helpers as H
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KW = [(31, 15), (31, 1), (5, 5), (5, 3), (41, 15), (56, 2), (21, 11), (32, 2), (33, 1), (1, 1), (16, 254)]


def entropy_bitmap(k, thr, stride=64):
    """Same table the library builds on the host (dcn_api.cu build_entropy_bitmap), evaluated with
    the oracle's restatement of src/minimizers.rs:73-121."""
    bits = np.zeros(stride ** 3 // 32, np.uint32)
    for a in range(k + 1):
        for c in range(k + 1 - a):
            for g in range(k + 1 - a - c):
                kmer = b"A" * a + b"C" * c + b"G" * g + b"T" * (k - a - c - g)
                if O.scaled_entropy(kmer, k) >= np.float32(thr):
                    idx = (a * stride + c) * stride + g
                    bits[idx >> 5] |= np.uint32(1 << (idx & 31))
    return bits


def _records(seed):
    g = H.random_genome(30_000, seed)
    recs = H.sample_reads(g, 150, (0, 700), seed + 1, n_rate=0.05, lower_rate=0.1)
    recs += [g[:6000].copy(), np.frombuffer(b"ACGTNNNNRYKMacgtnryk" * 40, np.uint8).copy(), np.frombuffer(b"A" * 900, np.uint8).copy(),
             np.frombuffer(b"ACGTACGTACGT", np.uint8).copy(), np.zeros(0, np.uint8), np.arange(256, dtype=np.uint8),
             np.concatenate([g[100:400], np.frombuffer(b"\n", np.uint8)])]
    return g, recs


@pytest.mark.parametrize("k,w", KW)
@pytest.mark.parametrize("cstride", [256, 7])
def test_generic_extract_filter_flavour(k, w, cstride):
    if k > 56:
        pytest.skip("filter side asserts k <= 56")
    _, recs = _records(100 + k)
    bases, off = H.concat(recs)
    for prefix in (0, 90):
        h, p, oo = E.generic_extract(bases, off, 0, k, w, prefix, cstride=cstride)
        assert not isinstance(h, int)
        for i, r in enumerate(recs):
            wh, wp = O.extract_filter(r, k, w, prefix)
            a, b = int(oo[i]), int(oo[i + 1])
            assert np.array_equal(h[a:b], wh) and np.array_equal(p[a:b], wp), (k, w, prefix, i)


@pytest.mark.parametrize("k,w", KW + [(57, 1)])
def test_generic_extract_index_flavour(k, w):
    _, recs = _records(200 + k)
    bases, off = H.concat(recs)
    for thr in (0.0, 0.5):
        bm = entropy_bitmap(k, thr) if thr else None
        h, p, oo = E.generic_extract(bases, off, 1, k, w, entropy_bitmap=bm, cstride=64)
        for i, r in enumerate(recs):
            want = O.extract_index(r, k, w, thr)
            assert np.array_equal(h[int(oo[i]):int(oo[i + 1])], want), (k, w, thr, i)


def test_generic_extract_reports_required_capacity():
    _, recs = _records(7)
    bases, off = H.concat(recs)
    h, p, oo = E.generic_extract(bases, off, 0, 31, 15)
    rc, _, oo2 = E.generic_extract(bases, off, 0, 31, 15, cap=10)
    assert rc == -6 and int(oo2[-1]) == len(h) and np.array_equal(oo, oo2)


@pytest.mark.parametrize("k,w", [(31, 1), (5, 5), (41, 15), (21, 11), (56, 2)])
@pytest.mark.parametrize("paired", [False, True])
def test_generic_filter_matches_oracle(k, w, paired):
    g, recs = _records(300 + k)
    idx = O.index_build([g[:20_000]], k, w)
    bases, off = H.concat(recs[: len(recs) // 2 * 2])
    for (a, r, dep, prefix) in ((2, 0.01, False, 0), (1, 0.0, True, 0), (2, 0.2, True, 100)):
        rc, kk, hh, tt = E.generic_filter(idx.keys(), bases, off, k, w, paired, prefix, a, r, dep, cstride=64)
        assert rc == 0
        ok, oh, ot = O.filter_batch(idx, bases, off, paired=paired, prefix_len=prefix, k=k, w=w, abs_thr=a, rel_thr=r, deplete=dep)
        assert np.array_equal(tt, ot) and np.array_equal(hh, oh) and np.array_equal(kk, ok), (k, w, paired, a, r, dep)


def test_reference_known_answers_all_kw():
    """Every behavioural known-answer test of tests/filter_tests.rs, including the ones with
    non-default (k, w): k=31 w=1 (:1133-1187), k=5 w=5 (:1190-1251), k=41 (:1254-1296)."""
    with open(os.path.join(GOLD, "reference_kats.json")) as f:
        kats = json.load(f)["cases"]
    for c in kats:
        k, w = c["k"], c["w"]
        idx = O.index_build([r.encode() for r in c["ref"]], k, w)
        if "reads" in c:
            recs, paired = [r.encode() for r in c["reads"]], False
        else:
            recs, paired = [x.encode() for pair in zip(c["reads1"], c["reads2"]) for x in pair], True
        bases, off = O.concat_records(recs)
        rc, kk, hh, tt = E.generic_filter(idx.keys(), bases, off, k, w, paired, 0, c["abs"], c["rel"], c["deplete"])
        assert rc == 0 and list(map(int, kk)) == c["expect_keep"], c["name"]
        if "expect_hits" in c:
            assert list(map(int, hh)) == c["expect_hits"], c["name"]

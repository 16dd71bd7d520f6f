"""CPU tests: the CUDA tile pipeline, compiled for the host and run phase by phase, must agree
bit-for-bit with the oracle.  This checks the kernel's LOGIC without a GPU; the GPU tests
(test_gpu_parity.py) then check the same cases through the C ABI on a B200."""
import numpy as np
import pytest

import cases as CASES
import emu_harness as E
import helpers as H
from oracle import oracle as O


def _index_for(case):
    idx = O.index_build([np.asarray(r, np.uint8) for r in case["index_records"]], 31, 15)
    if case.get("extra_keys") is not None:
        idx.insert(case["extra_keys"])
    return idx


@pytest.fixture(params=["warp", "cta"])
def impl(request):
    """Both mappings of the short path: warp tiles (+ CTA tail), which is what the library runs, and the CTA-tile
    fused kernel kept behind DCN_FUSED_IMPL=cta."""
    E.set_impl(request.param)
    yield request.param
    E.set_impl("warp")


@pytest.mark.parametrize("packed", [False, True], ids=["ascii", "packed"])
@pytest.mark.parametrize("case", CASES.make_cases(), ids=lambda c: c["name"])
def test_short_path_matches_oracle(case, packed, impl):
    idx = _index_for(case)
    bases, off = H.concat(case["records"])
    rc, k, h, t = E.filter_batch(idx.keys(), bases, off, case["paired"], case["prefix"], case["abs"], case["rel"], case["deplete"],
                                 packed=packed)
    assert rc == 0
    if impl == "warp" and case["name"] == "dense_picks_polyA":
        assert E.last_overflow_units() == 8   # the 1000-base poly-A units emit more picks than a warp pass holds: CTA tail
    ok, oh, ot = O.filter_batch(idx, bases, off, paired=case["paired"], prefix_len=case["prefix"], abs_thr=case["abs"],
                                rel_thr=case["rel"], deplete=case["deplete"])
    assert np.array_equal(t, ot), "total minimizers differ"
    assert np.array_equal(h, oh), "distinct hit counts differ"
    assert np.array_equal(k, ok), "keep decisions differ"


def test_high_load_factor_table_probing():
    g = H.random_genome(60_000, 3)
    idx = O.index_build([g], 31, 15)
    reads = H.sample_reads(g, 500, 150, 4)
    bases, off = H.concat(reads)
    ok, oh, ot = O.filter_batch(idx, bases, off)
    for load in (0.05, 0.5, 0.9):
        rc, k, h, t = E.filter_batch(idx.keys(), bases, off, load=load)
        assert rc == 0 and np.array_equal(h, oh) and np.array_equal(t, ot) and np.array_equal(k, ok)


@pytest.mark.parametrize("packed", [False, True], ids=["ascii", "packed"])
@pytest.mark.parametrize("case", CASES.make_long_cases(), ids=lambda c: c["name"])
def test_long_path_matches_oracle(case, packed, impl):
    idx = _index_for(case)
    bases, off = H.concat(case["records"])
    rc, k, h, t = E.filter_batch(idx.keys(), bases, off, case["paired"], case["prefix"], case["abs"], case["rel"], case["deplete"],
                                 packed=packed)
    assert rc == 0
    ok, oh, ot = O.filter_batch(idx, bases, off, paired=case["paired"], prefix_len=case["prefix"], abs_thr=case["abs"],
                                rel_thr=case["rel"], deplete=case["deplete"])
    assert np.array_equal(t, ot), "total minimizers differ"
    assert np.array_equal(h, oh), "distinct hit counts differ"
    assert np.array_equal(k, ok), "keep decisions differ"


def test_long_path_reports_dedup_overflow():
    case = CASES.make_long_cases()[0]
    idx = _index_for(case)
    bases, off = H.concat(case["records"])
    E.set_dedup_cap(64)
    try:
        rc, *_ = E.filter_batch(idx.keys(), bases, off)
        assert rc == -6      # DCN_ERR_OVERFLOW: the API retries with a larger set, never drops hits
    finally:
        E.set_dedup_cap(0)


def test_long_path_hit_list_overflow_branch():
    """Warp-tile long path: the hits of a chunk are compacted into a list before they go through the distinct-hit set;
    a chunk with more hits than the list holds inserts the rest lane by lane.  A list of 5 entries sends most hits of
    every chunk down that branch; results must not change."""
    for case in CASES.make_long_cases()[:3]:
        idx = _index_for(case)
        bases, off = H.concat(case["records"])
        want = O.filter_batch(idx, bases, off, paired=case["paired"], prefix_len=case["prefix"], abs_thr=case["abs"],
                              rel_thr=case["rel"], deplete=case["deplete"])
        E.set_hit_list_cap(5)
        try:
            rc, k, h, t = E.filter_batch(idx.keys(), bases, off, case["paired"], case["prefix"], case["abs"], case["rel"], case["deplete"])
        finally:
            E.set_hit_list_cap(0)
        assert rc == 0
        assert np.array_equal(t, want[2]) and np.array_equal(h, want[1]) and np.array_equal(k, want[0])


def test_index_extraction_matches_oracle_set():
    g = H.random_genome(50_000, 41)
    recs = [g[:20_000].copy(), g[20_000:20_040].copy(), np.zeros(0, np.uint8), g[20_040:].copy(),
            np.frombuffer(b"ACGTNNNNRYKMacgtnryk" * 300, np.uint8).copy(), np.frombuffer(b"A" * 9000, np.uint8).copy()]
    recs[0][[100, 5000, 5001, 19_999]] = [ord("N"), ord("R"), ord("y"), ord("-")]
    bases, off = H.concat(recs)
    got = np.unique(E.index_extract(bases, off))
    want = O.index_build((bases, off), 31, 15).keys()
    assert np.array_equal(got, want)


def test_host_packer_simd_matches_scalar_and_definition():
    """dcn_host_pack.cpp: the AVX2/BMI2 packer == the scalar loop == the definition
    (code = (byte >> 1) & 3, src/filter_common.rs:238; non-ACGT mask, :245-258), all 256 byte values."""
    rng = np.random.default_rng(3)
    for n in (0, 1, 15, 16, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1000, 4097):
        b = rng.integers(0, 256, n).astype(np.uint8)
        if n >= 64:
            b[:40] = np.frombuffer(b"ACGTacgtNnRYKM\n-*" * 3, np.uint8)[:40]
        c0, i0 = E.pack_ascii(b, simd=0)
        for level in (1, 2):   # 1 = best the CPU has (AVX-512 BW where present), 2 = the AVX2 + BMI2 path
            c1, i1 = E.pack_ascii(b, simd=level)
            assert np.array_equal(c1, c0) and np.array_equal(i1, i0), (n, level)
        pad = np.zeros(len(c0) * 16, np.uint8)
        pad[:n] = b
        code = ((pad >> 1) & 3).astype(np.uint32).reshape(-1, 16)
        want_c = (code << (2 * np.arange(16, dtype=np.uint32))).sum(axis=1).astype(np.uint32)
        bad = ~np.isin(pad & 0xDF, np.frombuffer(b"ACGT", np.uint8))
        want_i = (bad.reshape(-1, 16).astype(np.uint32) << np.arange(16, dtype=np.uint32)).sum(axis=1).astype(np.uint16)
        assert np.array_equal(c0, want_c) and np.array_equal(i0, want_i), n


@pytest.mark.parametrize("prefix", [0, 100])
def test_warp_tile_extraction_matches_oracle(prefix):
    """B3 through the warp tiles (extract_warp_kernel + extract_tail_kernel for records of more than a warp pass's
    picks): per record the hashes AND positions of get_minimizer_hashes_and_positions (src/filter_common.rs:211-310)."""
    g = H.random_genome(80_000, 77)
    recs = H.sample_reads(g, 700, (0, 420), 78, n_rate=0.02, lower_rate=0.1)
    recs += [np.frombuffer(b"A" * 1000, np.uint8).copy()] * 3 + [np.frombuffer(b"ACGT" * 60, np.uint8).copy()] * 5   # dense picks
    recs += [np.concatenate([r, np.frombuffer(b"\n", np.uint8)]) for r in recs[:50]]
    bases, off = H.concat(recs)
    h, p, oo = E.tile_extract(bases, off, prefix)
    assert E.last_overflow_units() == (3 if prefix == 0 else 0)   # the untrimmed poly-A records go through the CTA-tile extraction
    for r, rec in enumerate(recs):
        wh, wp = O.extract_filter(rec, 31, 15, prefix)
        a, b = int(oo[r]), int(oo[r + 1])
        assert np.array_equal(h[a:b], wh) and np.array_equal(p[a:b], wp), r

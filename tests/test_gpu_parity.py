"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the
oracle on the same seeded inputs, the committed golden fixtures and the reference's behavioural
known-answer tests.  Bit-exact: integer work only."""
import json
import os

import numpy as np
import pytest

import cases as CASES
import helpers as H
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _index_for(case):
    idx = O.index_build([np.asarray(r, np.uint8) for r in case["index_records"]], 31, 15)
    if case.get("extra_keys") is not None:
        idx.insert(case["extra_keys"])
    return idx


def _check(gpu, idx, case):
    from deacon_server_b200 import IndexHeader
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    assert gpu.index_info()["n_keys"] == len(idx)
    bases, off = H.concat(case["records"])
    k, h, t = gpu.filter_batch(bases, off, paired=case["paired"], prefix_length=case["prefix"], abs_threshold=case["abs"],
                               rel_threshold=case["rel"], deplete=case["deplete"])
    ok, oh, ot = O.filter_batch(idx, bases, off, paired=case["paired"], prefix_len=case["prefix"], abs_thr=case["abs"],
                                rel_thr=case["rel"], deplete=case["deplete"], threads=8)
    assert np.array_equal(t, ot), "total minimizers differ"
    assert np.array_equal(h, oh), "distinct hit counts differ"
    assert np.array_equal(k, ok), "keep decisions differ"
    # the same call with host packing off (ASCII bytes over PCIe, converted on the GPU)
    gpu.host_pack_threads(0)
    try:
        k2, h2, t2 = gpu.filter_batch(bases, off, paired=case["paired"], prefix_length=case["prefix"], abs_threshold=case["abs"],
                                      rel_threshold=case["rel"], deplete=case["deplete"])
    finally:
        gpu.host_pack_threads(4)
    assert np.array_equal(t2, ot) and np.array_equal(h2, oh) and np.array_equal(k2, ok), "ASCII ingest differs"
    # and with the batch packed by the caller (dcn_pack_ascii + dcn_newline_bits -> dcn_filter_batch_packed)
    from deacon_server_b200 import api
    codes, inv = api.pack_ascii(bases)
    nl = api.newline_bits(bases, off, 31, case["prefix"])
    k3, h3, t3 = gpu.filter_batch_packed(codes, inv, nl, off, paired=case["paired"], prefix_length=case["prefix"],
                                         abs_threshold=case["abs"], rel_threshold=case["rel"], deplete=case["deplete"])
    assert np.array_equal(t3, ot) and np.array_equal(h3, oh) and np.array_equal(k3, ok), "caller-packed ingest differs"


@pytest.mark.parametrize("case", CASES.make_cases(), ids=lambda c: c["name"])
def test_filter_batch_matches_oracle(gpu, case):
    _check(gpu, _index_for(case), case)


def test_load_factors(gpu):
    g = H.random_genome(60_000, 3)
    idx = O.index_build([g], 31, 15)
    reads = H.sample_reads(g, 2000, 150, 4)
    case = dict(records=reads, paired=True, prefix=0, abs=2, rel=0.01, deplete=True)
    for load in (0.05, 0.5, 0.9):
        gpu.set_load_factor(load)
        _check(gpu, idx, case)
    gpu.set_load_factor(0.5)


def test_reference_behavioural_known_answers_k31_w15(gpu):
    """tests/filter_tests.rs scenarios that use the default k=31, w=15 (the CUDA fast path)."""
    from deacon_server_b200 import IndexHeader
    with open(os.path.join(GOLD, "reference_kats.json")) as f:
        kats = json.load(f)["cases"]
    ran = 0
    for c in kats:
        if (c["k"], c["w"]) != (31, 15):
            continue
        idx = O.index_build([r.encode() for r in c["ref"]], 31, 15)
        gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
        if "reads" in c:
            recs, paired = [r.encode() for r in c["reads"]], False
        else:
            recs, paired = [x.encode() for pair in zip(c["reads1"], c["reads2"]) for x in pair], True
        bases, off = O.concat_records(recs)
        k, h, t = gpu.filter_batch(bases, off, paired=paired, abs_threshold=c["abs"], rel_threshold=c["rel"], deplete=c["deplete"])
        assert list(map(int, k)) == c["expect_keep"], c["name"]
        if "expect_hits" in c:
            assert list(map(int, h)) == c["expect_hits"], c["name"]
        ran += 1
    assert ran >= 6


def _config1_reads(g, n, seed):
    """50 % sampled from the genome (random strand, 1 % substitutions), 50 % random,
    0.1 % of the reads carry one N - vectorised so that the full 1 M-read size is generated in a second."""
    rng = np.random.default_rng(seed)
    pos = rng.integers(0, len(g) - 150, n)
    reads = g[pos[:, None] + np.arange(150)[None, :]]
    rc = rng.random(n) < 0.5
    reads[rc] = H._COMP[reads[rc][:, ::-1]]
    rnd = rng.random(n) < 0.5
    reads[rnd] = H.ACGT[rng.integers(0, 4, (int(rnd.sum()), 150))]
    sub = rng.random((n, 150)) < 0.01
    sub[rnd] = False
    reads[sub] = H.ACGT[rng.integers(0, 4, int(sub.sum()))]
    withn = np.flatnonzero(rng.random(n) < 0.001)
    reads[withn, rng.integers(0, 150, len(withn))] = ord("N")
    return np.ascontiguousarray(reads.reshape(-1)), np.arange(n + 1, dtype=np.uint64) * np.uint64(150)


def test_config1_full_size_and_counters(gpu):
    """BASELINE config 1 at its full size: 1 M single-end 150 bp reads vs a 10 Mbp random reference indexed with
    k=31 w=15 (index built on the GPU and checked against the oracle's), search mode, -a 2 -r 0.01; every
    (keep, hits, total) against the oracle, and the six ProcessingStats counters (src/local_filter.rs:179-187, 347-371)."""
    g = H.random_genome(10_000_000, 1)
    gb, go = H.concat([g])
    keys = gpu.index_build(gb, go, 31, 15, 0.0, make_resident=True)
    idx = O.index_build([g], 31, 15, threads=8)
    assert np.array_equal(keys, idx.keys()) and 1_200_000 < len(keys) < 1_300_000
    bases, off = _config1_reads(g, 1_000_000, 2)
    reads = [None] * 1_000_000
    gpu.stats_reset()
    k, h, t = gpu.filter_batch(bases, off)
    ok, oh, ot = O.filter_batch(idx, bases, off, threads=8)
    assert np.array_equal(t, ot) and np.array_equal(h, oh) and np.array_equal(k, ok)
    st = gpu.stats()
    lens = np.diff(off).astype(np.int64)
    assert st["total_seqs"] == len(reads) and st["total_bp"] == int(lens.sum())
    assert 0.45 < ok.mean() < 0.55 and int(oh.max()) >= 14
    assert st["output_seq_counter"] == int(ok.sum()) and st["filtered_seqs"] == len(reads) - int(ok.sum())
    assert st["output_bp"] == int(lens[ok.astype(bool)].sum()) and st["filtered_bp"] == int(lens[~ok.astype(bool)].sum())
    from deacon_server_b200 import parallel as P
    assert P.counters_of(off, k, paired=False) == st                      # host mirror of stats_kernel
    assert P.reduce_counters(st) == st                                     # single process: identity
    gpu.stats_reset()
    k2, _, _ = gpu.filter_batch(bases, off, paired=True, deplete=True)
    assert P.counters_of(off, k2, paired=True) == gpu.stats()


def test_chunked_pipeline_many_chunks(gpu, monkeypatch):
    """Host-pointer path with > 2 chunks in flight (chunk size is read once per process; force small batches
    by calling with several MB of reads)."""
    from deacon_server_b200 import IndexHeader
    g = H.random_genome(400_000, 5)
    idx = O.index_build([g], 31, 15, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    rng = np.random.default_rng(6)
    pos = rng.integers(0, len(g) - 150, 600_000)
    bases = g[(pos[:, None] + np.arange(150)[None, :])].reshape(-1).copy()     # 90 MB -> 3 chunks of 32 MB
    off = (np.arange(len(pos) + 1, dtype=np.uint64) * np.uint64(150))
    k, h, t = gpu.filter_batch(bases, off, paired=True, deplete=True)
    ok, oh, ot = O.filter_batch(idx, bases, off, paired=True, deplete=True, threads=8)
    assert np.array_equal(t, ot) and np.array_equal(h, oh) and np.array_equal(k, ok)


def test_equal_length_chunks_do_not_ship_their_offsets(gpu):
    """Host ingest: a chunk whose records all have one length gets its rec_off written on the device (arithmetic sequence)
    instead of copied; one odd record anywhere in the chunk switches that chunk back to the copy.  Results never change."""
    from deacon_server_b200 import IndexHeader
    g = H.random_genome(300_000, 15)
    idx = O.index_build([g], 31, 15, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    gpu.host_pack_threads(0)   # ASCII route: the bytes moved are the bases plus whatever offsets are shipped
    try:
        rng = np.random.default_rng(16)
        n = 500_000                                                   # 75 MB: three chunks
        pos = rng.integers(0, len(g) - 151, n)
        reads = g[(pos[:, None] + np.arange(150)[None, :])]
        bases = reads.reshape(-1).copy()
        off = np.arange(n + 1, dtype=np.uint64) * np.uint64(150)
        for paired in (False, True):
            k, h, t = gpu.filter_batch(bases, off, paired=paired)
            ok, oh, ot = O.filter_batch(idx, bases, off, paired=paired, threads=8)
            assert np.array_equal(t, ot) and np.array_equal(h, oh) and np.array_equal(k, ok)
            h2d, d2h = gpu.last_transfer_bytes()
            assert len(bases) <= h2d < len(bases) + 64 and d2h == 9 * len(k)
        # one 151-base record in the last chunk: only that chunk ships its offsets
        lens = np.full(n, 150, np.uint64)
        lens[n - 7] = 151
        off2 = np.zeros(n + 1, np.uint64)
        off2[1:] = np.cumsum(lens)
        bases2 = np.insert(bases, int(off2[n - 7]) + 150, ord("A"))
        k, h, t = gpu.filter_batch(bases2, off2)
        ok, oh, ot = O.filter_batch(idx, bases2, off2, threads=8)
        assert np.array_equal(t, ot) and np.array_equal(h, oh) and np.array_equal(k, ok)
        h2d, _ = gpu.last_transfer_bytes()
        assert len(bases2) + 8 * 10_000 < h2d < len(bases2) + 8 * (n + 1)
    finally:
        gpu.host_pack_threads(4)


def test_two_route_ingest_pinned_ragged(gpu):
    """Host ingest with pinned caller buffers: the batch is cut into atoms, the copy engine ships ASCII chunks from the
    front while packer threads pack atoms from the back (dcn_api.cu filter_pipeline).  Ragged records (some shorter
    than k, some ending in a newline, some with N) so every atom ships its offsets and newline flags; whatever way
    the two routes split the batch, the results are the oracle's."""
    import torch
    from deacon_server_b200 import IndexHeader
    g = H.random_genome(500_000, 21)
    idx = O.index_build([g], 31, 15, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    rng = np.random.default_rng(22)
    n = 1_400_000
    lens = rng.integers(100, 251, n).astype(np.uint64)
    lens[rng.integers(0, n, 2000)] = rng.integers(0, 31, 2000).astype(np.uint64)        # shorter than k, some empty
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])                                                                # ~245 MB: 7-8 chunks of 32 MB
    assert total > 7 * (32 << 20)
    # reads = one long random walk over the genome (half of it host), cut at the record boundaries
    start = rng.integers(0, len(g) - 400, total // 256 + 2)
    bases = g[(start[:, None] + np.arange(256)[None, :])].reshape(-1)[:total].copy()
    rnd = rng.integers(0, 2, total // 4096 + 1).astype(bool).repeat(4096)[:total]
    bases[rnd] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(rnd.sum()))]
    bases[rng.integers(0, total, 5000)] = ord("N")
    nl = rng.integers(0, n, 3000)
    nl = nl[lens[nl] >= 31]
    bases[(off[nl + 1] - 1).astype(np.int64)] = ord("\n")                              # records ending in a newline
    hb = torch.from_numpy(bases).pin_memory()
    ho = torch.from_numpy(off.view(np.int64)).pin_memory()
    np_units = n // 2
    hk = torch.zeros(np_units, dtype=torch.uint8).pin_memory()
    hh = torch.zeros(np_units, dtype=torch.int32).pin_memory()
    ht = torch.zeros(np_units, dtype=torch.int32).pin_memory()
    ok, oh, ot = O.filter_batch(idx, bases, off, paired=True, deplete=True, threads=8)
    assert 0.2 < ok.mean() < 0.8
    try:
        for threads, fraction in ((6, -1.0), (3, 0.5), (6, 1.0), (0, -1.0)):
            gpu.host_pack_threads(threads)
            gpu.host_pack_fraction(fraction)
            for _ in range(2):                                                          # the second call reuses the packers' blobs
                hk.zero_(); hh.zero_(); ht.zero_()
                gpu.stats_reset()
                gpu.filter_batch_ptr(hb.data_ptr(), ho.data_ptr(), n, True, 0, 2, 0.01, True, hk.data_ptr(), hh.data_ptr(), ht.data_ptr())
                assert np.array_equal(ht.numpy().view(np.uint32), ot) and np.array_equal(hh.numpy().view(np.uint32), oh)
                assert np.array_equal(hk.numpy(), ok)
                # the six summary counters are kept by the kernels, one share per launch of the pipeline, whatever its form
                from deacon_server_b200 import parallel as P
                assert gpu.stats() == P.counters_of(off, ok, True), (threads, fraction)
                h2d, d2h = gpu.last_transfer_bytes()
                assert d2h == 9 * np_units
                if threads == 0:
                    assert h2d >= total + 8 * n                                         # ASCII only: every base and offset crosses
                elif fraction == 1.0:
                    assert h2d < 0.5 * total                                            # packed only: 0.375 B/bp + offsets + flags
                else:
                    assert h2d < total + 8 * n                                          # some atoms went packed
    finally:
        gpu.host_pack_fraction(-1.0)
        gpu.host_pack_threads(4)


def test_two_route_ingest_long_reads(gpu):
    """The same two-route ingest on ONT-like records (config 3 shape): units longer than an atom, long-path units in
    packed atoms and in ASCII chunks, search mode."""
    import torch
    from deacon_server_b200 import IndexHeader
    g = H.random_genome(600_000, 31)
    idx = O.index_build([g], 31, 15, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    rng = np.random.default_rng(32)
    lens = np.clip(rng.gamma(2.0, 5000.0, 26_000), 200, 200_000).astype(np.uint64)
    lens[7] = 6_000_000                                                                # one record longer than an atom
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    assert total > 7 * (32 << 20)
    start = rng.integers(0, len(g) - 5000, total // 4096 + 2)
    bases = g[(start[:, None] + np.arange(4096)[None, :])].reshape(-1)[:total].copy()   # 4 kb stretches of the genome
    rnd = rng.integers(0, 2, total // 65536 + 1).astype(bool).repeat(65536)[:total]
    bases[rnd] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(rnd.sum()))]
    bases[rng.integers(0, total, 20_000)] = ord("N")
    hb = torch.from_numpy(bases).pin_memory()
    ho = torch.from_numpy(off.view(np.int64)).pin_memory()
    n = len(lens)
    hk = torch.zeros(n, dtype=torch.uint8).pin_memory()
    hh = torch.zeros(n, dtype=torch.int32).pin_memory()
    ht = torch.zeros(n, dtype=torch.int32).pin_memory()
    ok, oh, ot = O.filter_batch(idx, bases, off, paired=False, deplete=False, threads=8)
    assert int(oh.max()) > 1000
    try:
        for threads, fraction in ((6, -1.0), (4, 0.5), (6, 1.0)):
            gpu.host_pack_threads(threads)
            gpu.host_pack_fraction(fraction)
            hk.zero_(); hh.zero_(); ht.zero_()
            gpu.filter_batch_ptr(hb.data_ptr(), ho.data_ptr(), n, False, 0, 2, 0.01, False, hk.data_ptr(), hh.data_ptr(), ht.data_ptr())
            assert np.array_equal(ht.numpy().view(np.uint32), ot) and np.array_equal(hh.numpy().view(np.uint32), oh)
            assert np.array_equal(hk.numpy(), ok)
    finally:
        gpu.host_pack_fraction(-1.0)
        gpu.host_pack_threads(4)


def test_two_route_ingest_few_huge_units(gpu):
    """A batch of a handful of chromosome-sized records (fewer than eight atoms of >= 32 units): the arena form declines
    it and the chunk form, whose atoms hold single units, runs both routes."""
    import torch
    from deacon_server_b200 import IndexHeader
    g = H.random_genome(400_000, 61)
    idx = O.index_build([g], 31, 15, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    rng = np.random.default_rng(62)
    lens = np.full(12, 20_000_000, np.uint64)
    lens[5] = 31_000_000
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    start = rng.integers(0, len(g) - 70_000, total // 65536 + 2)
    bases = g[(start[:, None] + np.arange(65536)[None, :])].reshape(-1)[:total].copy()   # 64 kb stretches of the genome
    rnd = rng.integers(0, 2, total // 1_000_000 + 1).astype(bool).repeat(1_000_000)[:total]
    bases[rnd] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(rnd.sum()))]
    bases[rng.integers(0, total, 5000)] = ord("N")
    hb = torch.from_numpy(bases).pin_memory()
    ho = torch.from_numpy(off.view(np.int64)).pin_memory()
    n = len(lens)
    ok, oh, ot = O.filter_batch(idx, bases, off, paired=False, deplete=False, threads=8)
    assert int(oh.min()) > 1000
    try:
        for threads in (6, 0):
            gpu.host_pack_threads(threads)
            k, h, t = (torch.zeros(n, dtype=torch.uint8).pin_memory(), torch.zeros(n, dtype=torch.int32).pin_memory(),
                       torch.zeros(n, dtype=torch.int32).pin_memory())
            gpu.filter_batch_ptr(hb.data_ptr(), ho.data_ptr(), n, False, 0, 2, 0.01, False, k.data_ptr(), h.data_ptr(), t.data_ptr())
            assert np.array_equal(t.numpy().view(np.uint32), ot) and np.array_equal(h.numpy().view(np.uint32), oh)
            assert np.array_equal(k.numpy(), ok)
    finally:
        gpu.host_pack_threads(4)


_SMALL_ATOMS = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import helpers as H
import deacon_server_b200 as d
from oracle import oracle as O
g = H.random_genome(150_000, 41)
idx = O.index_build([g], 31, 15, threads=8)
gpu = d.DeaconGpu(0)
gpu.index_upload(idx.keys(), d.IndexHeader(2, 31, 15))
rng = np.random.default_rng(42)
n = 120_000
lens = rng.integers(0, 400, n).astype(np.uint64)
lens[::997] = 9_000                                    # a few long-path units (> DCN_MAX_SHORT) among the short ones
off = np.zeros(n + 1, np.uint64); off[1:] = np.cumsum(lens)
total = int(off[-1])
start = rng.integers(0, len(g) - 300, total // 128 + 2)
bases = g[(start[:, None] + np.arange(128)[None, :])].reshape(-1)[:total].copy()
bases[rng.integers(0, total, total // int(sys.argv[2]))] = ord("N")
nl = rng.integers(0, n, n // 10); nl = nl[lens[nl] >= 1]
bases[(off[nl + 1] - 1).astype(np.int64)] = 10         # every tenth record ends in a newline
for paired, prefix in ((False, 0), (True, 0), (False, 120)):
    want = O.filter_batch(idx, bases, off, paired=paired, prefix_len=prefix, deplete=paired, threads=8)
    for threads, fraction in ((0, -1.0), (5, -1.0), (3, 0.5), (5, 1.0)):
        gpu.host_pack_threads(threads); gpu.host_pack_fraction(fraction)
        gpu.stats_reset()
        got = gpu.filter_batch(bases, off, paired=paired, prefix_length=prefix, deplete=paired)
        for a, b in zip(got, want):
            assert np.array_equal(a, b), (paired, prefix, threads, fraction)
        from deacon_server_b200 import parallel as P
        assert gpu.stats() == P.counters_of(off, want[0], paired), (paired, prefix, threads, fraction)
print("small atoms ok")
"""


@pytest.mark.parametrize("n_every,sparse,extra", [(500, "1", {}), (40000, "1", {}), (500, "0", {}),
                                                  (40000, "1", {"DCN_PIPELINE": "chunks"}),
                                                  (500, "1", {"DCN_LAUNCH_ATOMS": "3", "DCN_LAUNCH_CAP_MB": "1", "DCN_PACKER_STAGES": "1"}),
                                                  (40000, "1", {"DCN_DEDUP_SHRINK": "8"})],
                         ids=["N-rich:dense-fallback", "N-rare:sparse-mask", "dense-wire", "chunk-pipeline", "arena:many-small-launches",
                              "distinct-hit-set-overflow"])
def test_two_route_ingest_small_atoms(tmp_path, n_every, sparse, extra):
    """The two-route ingest with 1 MB chunks (128 KB atoms; DCN_CHUNK_MB is read once per process, hence the
    subprocess): hundreds of atoms per call, every kind of record boundary inside them -- empty and sub-k records,
    newline-terminated ones, N runs, long-path units, prefix trimming -- on pageable buffers, all splits.  The packer
    threads ship the non-ACGT bits as a sparse exception list (an N every 40 000 bases: few blocks listed) and fall
    back to the dense mask when more than one 32-base block in 32 is listed (an N every 500 bases);
    DCN_SPARSE_MASK=0 keeps the dense wire form of round 1.  The calls with packer threads run the arena form of the
    pipeline (kernels over whatever contiguous range has arrived); DCN_PIPELINE=chunks runs the one-kernel-per-chunk form
    that still serves the single-route cases, and tiny launch limits make the arena form launch hundreds of ranges.
    DCN_DEDUP_SHRINK starts the distinct-hit sets of the long units at an eighth of their size: every launch with a long
    unit overflows and is repeated with a larger set (by the call itself in the chunk form, at retire time in the arena
    form, where nothing may have been committed to the counters by the overflowed attempt)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "small_atoms.py"
    script.write_text(_SMALL_ATOMS)
    env = dict(os.environ, DCN_CHUNK_MB="1", DCN_SPARSE_MASK=sparse, **extra)
    r = subprocess.run([sys.executable, str(script), root, str(n_every)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "small atoms ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_caller_packed_sparse_form_matches_ascii(gpu):
    """dcn_filter_batch_packed_sparse (2-bit codes + a sparse list of the 32-base blocks with a non-ACGT base, what a
    parser that packs while it parses can emit) == dcn_filter_batch_packed (dense mask) == dcn_filter_batch (ASCII) on
    ragged records with N runs, newline-terminated records, prefixes and pairs; 70 MB so that several pipeline chunks
    split the exception list; an unsorted list is refused."""
    from deacon_server_b200 import IndexHeader, DeaconCudaError, api as A
    g = H.random_genome(300_000, 51)
    idx = O.index_build([g], 31, 15, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    rng = np.random.default_rng(52)
    n = 500_000
    lens = rng.integers(60, 220, n).astype(np.uint64)
    lens[rng.integers(0, n, 3000)] = rng.integers(0, 31, 3000).astype(np.uint64)
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    assert total > (64 << 20)                                                         # three 32 MB chunks
    start = rng.integers(0, len(g) - 300, total // 128 + 2)
    bases = g[(start[:, None] + np.arange(128)[None, :])].reshape(-1)[:total].copy()
    bases[rng.integers(0, total, total // 3000)] = ord("N")
    run = int(rng.integers(0, total - 5000))
    bases[run:run + 4000] = ord("N")                                                  # a long run: many consecutive listed blocks
    nl = rng.integers(0, n, 5000); nl = nl[lens[nl] >= 31]
    bases[(off[nl + 1] - 1).astype(np.int64)] = 10
    gpu.host_pack_threads(0)
    try:
        for paired, prefix in ((False, 0), (True, 0), (False, 100)):
            want = gpu.filter_batch(bases, off, paired=paired, prefix_length=prefix, deplete=paired)
            ok = O.filter_batch(idx, bases, off, paired=paired, prefix_len=prefix, deplete=paired, threads=8)
            for a, b in zip(want, ok):
                assert np.array_equal(a, b)
            codes, exc, nlb = A.pack_records_sparse(bases, off, 31, prefix)
            assert 100 < len(exc) < total // 32 // 20
            got = gpu.filter_batch_packed_sparse(codes, exc, nlb, off, paired=paired, prefix_length=prefix, deplete=paired)
            for a, b in zip(got, want):
                assert np.array_equal(a, b), (paired, prefix)
            h2d, _ = gpu.last_transfer_bytes()
            assert h2d < 0.27 * total + 8 * (n + 1) + 8 * len(exc) + 4096
            codes_d, inv_d, nl_d = A.pack_records(bases, off, 31, prefix)
            got_d = gpu.filter_batch_packed(codes_d, inv_d, nl_d, off, paired=paired, prefix_length=prefix, deplete=paired)
            for a, b in zip(got_d, want):
                assert np.array_equal(a, b), (paired, prefix)
        with pytest.raises(DeaconCudaError):
            gpu.filter_batch_packed_sparse(codes, exc[::-1].copy(), nlb, off)
        # no exception at all: ACGT only, the batch ends on a block edge (no padding either)
        reads = [g[i * 160:i * 160 + 160].copy() for i in range(1000)]
        cb, co = H.concat(reads)
        c2, e2, n2 = A.pack_records_sparse(cb, co, 31, 0)
        assert len(e2) == 0
        got = gpu.filter_batch_packed_sparse(c2, e2, n2, co)
        for a, b in zip(got, gpu.filter_batch(cb, co)):
            assert np.array_equal(a, b)
    finally:
        gpu.host_pack_threads(4)


@pytest.mark.parametrize("layout", ["edge-lengths", "empty-runs", "segment-straddle", "tiny-batch", "all-long", "tile-limits"])
def test_warp_tile_planner_adversarial_layouts(gpu, layout):
    """wplan_kernel plans ~5 tiles per memory round trip: every lane works out the tile that would start at its unit and
    the warp hops through those counts.  Record layouts that sit on its limits -- units of exactly the longest short
    length next to long ones, hundreds of empty records in a row (32 units / 64 records per tile), units straddling the
    64 KB segment edges, a batch smaller than one segment, long units only, units that fill a tile to the byte -- must
    give the oracle's result for every unit, single and paired."""
    from deacon_server_b200 import IndexHeader
    g = H.random_genome(300_000, 71)
    idx = O.index_build([g], 31, 15, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    rng = np.random.default_rng(72)
    if layout == "edge-lengths":
        lens = rng.choice([1023, 1024, 1025, 1026, 31, 45, 44, 2048, 300], 6000)
    elif layout == "empty-runs":
        lens = rng.integers(100, 200, 20_000)
        for s0 in rng.integers(0, 19_000, 30):
            lens[s0:s0 + int(rng.integers(40, 700))] = 0
        lens[rng.integers(0, 20_000, 2000)] = rng.integers(1, 31, 2000)
    elif layout == "segment-straddle":
        lens = np.full(9000, 4096 // 8, np.int64)            # 512-base units: starts on every 64 KB edge ...
        lens[::7] = 65536 // 64 - 1                          # ... then drifting across them one base at a time
        lens[3::11] = 1024
    elif layout == "tiny-batch":
        lens = rng.integers(0, 400, 37)
    elif layout == "all-long":
        lens = rng.integers(1025, 30_000, 400)
    else:
        lens = rng.choice([1536, 1521, 1520, 768, 760, 512, 384, 48, 47], 8000)   # > 1024 are long: tiles of exactly / almost 1536 bases
    lens = np.asarray(lens, np.uint64)
    if len(lens) % 2:
        lens = lens[:-1]
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    start = rng.integers(0, len(g) - 3000, total // 2048 + 2)
    bases = g[(start[:, None] + np.arange(2048)[None, :])].reshape(-1)[:total].copy()
    bases[rng.integers(0, max(total, 1), total // 5000 + 1)] = ord("N")
    gpu.host_pack_threads(0)
    try:
        for paired in (False, True):
            got = gpu.filter_batch(bases, off, paired=paired, deplete=paired)
            want = O.filter_batch(idx, bases, off, paired=paired, deplete=paired, threads=8)
            for a, b, name in zip(got, want, ("keep", "hits", "total")):
                assert np.array_equal(a, b), (layout, paired, name, int(np.flatnonzero(a != b)[0]))
    finally:
        gpu.host_pack_threads(4)


def test_device_pointer_api_matches_host_api(gpu):
    import torch
    from deacon_server_b200 import IndexHeader
    g = H.random_genome(200_000, 7)
    idx = O.index_build([g], 31, 15, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    reads = H.sample_reads(g, 20_000, 150, 8)
    bases, off = H.concat(reads)
    k, h, t = gpu.filter_batch(bases, off, paired=True, deplete=True)
    dev = torch.device("cuda:0")
    d_b = torch.from_numpy(bases).to(dev)
    d_o = torch.from_numpy(off.view(np.int64)).to(dev)
    nu = len(reads) // 2
    d_k = torch.zeros(nu, dtype=torch.uint8, device=dev)
    d_h = torch.zeros(nu, dtype=torch.int32, device=dev)
    d_t = torch.zeros(nu, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    gpu.filter_batch_device(d_b, d_o, len(reads), len(bases), d_k, d_h, d_t, paired=True, deplete=True, stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(d_k.cpu().numpy(), k)
    assert np.array_equal(d_h.cpu().numpy().view(np.uint32), h)
    assert np.array_equal(d_t.cpu().numpy().view(np.uint32), t)


def test_device_pointer_api_with_length_promise(gpu):
    """dcn_filter_batch_device_hint: with the caller's promise that every unit is short the call only enqueues
    (no readback); results and the six counters are the same; a broken promise is reported by the next call."""
    import torch
    from deacon_server_b200 import IndexHeader, DeaconCudaError
    g = H.random_genome(300_000, 17)
    idx = O.index_build([g], 31, 15, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    reads = H.sample_reads(g, 30_000, (60, 151), 18, n_rate=0.01)
    bases, off = H.concat(reads)
    ok, oh, ot = O.filter_batch(idx, bases, off, paired=True, deplete=True, threads=8)
    dev = torch.device("cuda:0")
    d_b = torch.from_numpy(bases).to(dev)
    d_o = torch.from_numpy(off.view(np.int64)).to(dev)
    nu = len(reads) // 2
    st = torch.cuda.current_stream().cuda_stream
    gpu.stats_reset()
    outs = []
    for hint in (0, 302, 1024):
        d_k = torch.zeros(nu, dtype=torch.uint8, device=dev)
        d_h = torch.zeros(nu, dtype=torch.int32, device=dev)
        d_t = torch.zeros(nu, dtype=torch.int32, device=dev)
        gpu.filter_batch_device(d_b, d_o, len(reads), len(bases), d_k, d_h, d_t, paired=True, deplete=True, stream=st,
                                max_unit_len=hint)
        outs.append((d_k, d_h, d_t))
    torch.cuda.synchronize()
    for d_k, d_h, d_t in outs:
        assert np.array_equal(d_k.cpu().numpy(), ok)
        assert np.array_equal(d_h.cpu().numpy().view(np.uint32), oh)
        assert np.array_equal(d_t.cpu().numpy().view(np.uint32), ot)
    c = gpu.stats()
    assert c["total_seqs"] == 3 * len(reads) and c["total_bp"] == 3 * len(bases)
    assert c["output_seq_counter"] == 3 * 2 * int(ok.sum())
    # a batch with one long unit under a "short units only" promise: detected on the device, reported by the next call
    long_reads = reads[:100] + [g[:5000].copy(), g[5000:5150].copy()] + reads[100:200]
    b2, o2 = H.concat(long_reads)
    d_b2 = torch.from_numpy(b2).to(dev)
    d_o2 = torch.from_numpy(o2.view(np.int64)).to(dev)
    n2 = len(long_reads) // 2
    d_k = torch.zeros(n2, dtype=torch.uint8, device=dev)
    d_h = torch.zeros(n2, dtype=torch.int32, device=dev)
    d_t = torch.zeros(n2, dtype=torch.int32, device=dev)
    gpu.filter_batch_device(d_b2, d_o2, len(long_reads), len(b2), d_k, d_h, d_t, paired=True, deplete=True, stream=st,
                            max_unit_len=302)
    torch.cuda.synchronize()
    if os.environ.get("DCN_FUSED_IMPL") != "cta":     # (the CTA-tile A/B path ignores the promise: it always asks the device)
        with pytest.raises(DeaconCudaError):
            gpu.filter_batch_device(d_b2, d_o2, len(long_reads), len(b2), d_k, d_h, d_t, paired=True, deplete=True, stream=st)
    # the error is reported once; without the promise the same batch is classified in full
    gpu.filter_batch_device(d_b2, d_o2, len(long_reads), len(b2), d_k, d_h, d_t, paired=True, deplete=True, stream=st)
    torch.cuda.synchronize()
    k2, h2, t2 = O.filter_batch(idx, b2, o2, paired=True, deplete=True, threads=8)
    assert np.array_equal(d_k.cpu().numpy(), k2) and np.array_equal(d_h.cpu().numpy().view(np.uint32), h2)
    assert np.array_equal(d_t.cpu().numpy().view(np.uint32), t2)
    gpu.stats_reset()


@pytest.mark.parametrize("case", CASES.make_long_cases(), ids=lambda c: c["name"])
def test_long_path_matches_oracle(gpu, case):
    """Units longer than 1024 bases: chunked kernel + global (hash, unit) distinct-hit set."""
    _check(gpu, _index_for(case), case)


def test_ont_like_reads_config3_shape(gpu):
    """BASELINE config 3 shape, scaled: gamma(2) lengths, mean 10 kbp, 5 % substitutions, search mode."""
    from deacon_server_b200 import IndexHeader
    g = H.random_genome(2_000_000, 51)
    idx = O.index_build([g], 31, 15, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    rng = np.random.default_rng(52)
    lens = np.clip(rng.gamma(2.0, 5000.0, 600), 200, 200_000).astype(int)
    reads = []
    for i, ln in enumerate(lens):
        if i % 2 == 0:
            p = int(rng.integers(0, len(g) - ln))
            r = g[p:p + ln].copy()
            m = rng.random(ln) < 0.05
            r[m] = H.ACGT[rng.integers(0, 4, int(m.sum()))]
        else:
            r = H.ACGT[rng.integers(0, 4, ln)]
        reads.append(r)
    bases, off = H.concat(reads)
    k, h, t = gpu.filter_batch(bases, off)
    ok, oh, ot = O.filter_batch(idx, bases, off, threads=8)
    assert np.array_equal(t, ot) and np.array_equal(h, oh) and np.array_equal(k, ok)
    assert int(oh.max()) > 500   # long host-derived reads carry hundreds of distinct hits


def test_index_build_matches_oracle_key_set(gpu):
    """Config 4 shape, scaled: GPU extraction + radix sort + unique == the oracle's key set; the
    resident table then answers like the oracle's set."""
    g = H.random_genome(3_000_000, 61)
    recs = [g[:1_000_000].copy(), g[1_000_000:1_000_030].copy(), np.zeros(0, np.uint8), g[1_000_030:].copy(),
            np.frombuffer(b"ACGTNNNNRYKMacgtnryk" * 500, np.uint8).copy(), np.frombuffer(b"A" * 20_000, np.uint8).copy()]
    recs[0][[100, 5000, 5001, 999_999]] = [ord("N"), ord("R"), ord("y"), ord("-")]
    bases, off = H.concat(recs)
    keys = gpu.index_build(bases, off, 31, 15, 0.0, make_resident=True)
    want = O.index_build((bases, off), 31, 15, threads=8).keys()
    assert np.array_equal(keys, want)
    assert gpu.index_info()["n_keys"] == len(want)
    reads = H.sample_reads(g, 5000, 150, 62)
    rb, ro = H.concat(reads)
    k, h, t = gpu.filter_batch(rb, ro, paired=True, deplete=True)
    ok, oh, ot = O.filter_batch(O.IndexSet(want), rb, ro, paired=True, deplete=True, threads=8)
    assert np.array_equal(t, ot) and np.array_equal(h, oh) and np.array_equal(k, ok)


def test_index_build_entropy_filter(gpu):
    """-e 0.5 (src/minimizers.rs:163-168) on a reference with low-complexity inserts."""
    g = H.random_genome(400_000, 71)
    rng = np.random.default_rng(72)
    for _ in range(80):   # homopolymer / dinucleotide / low-entropy runs
        p = int(rng.integers(0, len(g) - 400))
        unit = [b"A", b"T", b"AC", b"GT", b"AAC", b"AAAAAAAG"][int(rng.integers(0, 6))]
        run = np.frombuffer((unit * 400)[:int(rng.integers(60, 400))], np.uint8)
        g[p:p + len(run)] = run
    bases, off = H.concat([g])
    for thr in (0.5, 0.25, 0.9):
        keys = gpu.index_build(bases, off, 31, 15, thr, make_resident=False)
        want = O.index_build((bases, off), 31, 15, entropy=thr, threads=8).keys()
        assert np.array_equal(keys, want), f"entropy threshold {thr}"
    full = gpu.index_build(bases, off, 31, 15, 0.0, make_resident=False)
    assert len(full) > len(gpu.index_build(bases, off, 31, 15, 0.5, make_resident=False))


def test_idx_file_roundtrip_through_gpu(gpu, tmp_path):
    """write_minimizers -> load_minimizer_hashes -> resident table (src/index.rs:80-164)."""
    from deacon_server_b200 import IndexHeader, load_minimizer_hashes, write_minimizers
    g = H.random_genome(300_000, 81)
    bases, off = H.concat([g])
    keys = gpu.index_build(bases, off, 31, 15, 0.0, make_resident=False)
    p = tmp_path / "ref.idx"
    write_minimizers(keys, IndexHeader(2, 31, 15), p)
    ver, k, w, okeys = O.idx_decode(open(p, "rb").read())       # the oracle's codec reads what we wrote
    assert (ver, k, w) == (2, 31, 15) and np.array_equal(okeys, keys)
    got, hdr = load_minimizer_hashes(p)
    hdr2 = gpu.load_index(p)
    assert hdr2.kmer_length == 31 and gpu.index_info()["n_keys"] == len(keys) and np.array_equal(got, keys)


def test_unsupported_kw_fails_loudly(gpu):
    """Parameters the reference itself rejects: even k + w - 1 (src/index.rs:186-194), k > 56 when
    filtering (src/filter_common.rs:269-272).  Errors, never a silent CPU path."""
    from deacon_server_b200 import DeaconCudaError, IndexHeader
    reads = (np.frombuffer(b"ACGT" * 40, np.uint8), np.array([0, 160], np.uint64))
    gpu.index_upload(np.array([1, 2, 3], np.uint64), IndexHeader(2, 21, 10))
    with pytest.raises(DeaconCudaError, match="odd"):
        gpu.filter_batch(*reads)
    gpu.index_upload(np.array([1, 2, 3], np.uint64), IndexHeader(2, 57, 1))
    with pytest.raises(DeaconCudaError, match="56"):
        gpu.filter_batch(*reads)
    with pytest.raises(DeaconCudaError, match="odd"):
        gpu.index_build(*reads, 31, 16)


# --------------------------------------------------------------------------- any (k, w) + B3 extraction
GENERIC_KW = [(31, 1), (5, 5), (41, 15), (21, 11), (56, 2), (32, 2), (33, 1)]


def _generic_records(seed):
    g = H.random_genome(60_000, seed)
    recs = H.sample_reads(g, 1500, (0, 700), seed + 1, n_rate=0.05, lower_rate=0.1)
    recs += [g[:20_000].copy(), np.frombuffer(b"ACGTNNNNRYKMacgtnryk" * 40, np.uint8).copy(), np.frombuffer(b"A" * 900, np.uint8).copy(),
             np.zeros(0, np.uint8), np.arange(256, dtype=np.uint8), np.concatenate([g[100:400], np.frombuffer(b"\n", np.uint8)])]
    return g, recs


@pytest.mark.parametrize("k,w", GENERIC_KW)
def test_generic_kw_filter_matches_oracle(gpu, k, w):
    """Indexes with non-default parameters (generic kernel; u128 k-mers above k = 32)."""
    from deacon_server_b200 import IndexHeader
    g, recs = _generic_records(400 + k)
    idx = O.index_build([g[:40_000]], k, w, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, k, w))
    bases, off = H.concat(recs[: len(recs) // 2 * 2])
    for paired in (False, True):
        for (a, r, dep, prefix) in ((2, 0.01, False, 0), (1, 0.0, True, 0), (2, 0.2, True, 100)):
            kk, hh, tt = gpu.filter_batch(bases, off, paired=paired, prefix_length=prefix, abs_threshold=a, rel_threshold=r, deplete=dep)
            ok, oh, ot = O.filter_batch(idx, bases, off, paired=paired, prefix_len=prefix, k=k, w=w, abs_thr=a, rel_thr=r,
                                        deplete=dep, threads=8)
            assert np.array_equal(tt, ot) and np.array_equal(hh, oh) and np.array_equal(kk, ok), (k, w, paired, a, r, dep)


def test_reference_behavioural_known_answers_all_kw(gpu):
    """tests/filter_tests.rs scenarios with non-default parameters: k=31 w=1 (:1133-1187), k=5 w=5
    (:1190-1251), k=41 (:1254-1296), through the C ABI."""
    from deacon_server_b200 import IndexHeader
    with open(os.path.join(GOLD, "reference_kats.json")) as f:
        kats = json.load(f)["cases"]
    ran = 0
    for c in kats:
        k, w = c["k"], c["w"]
        if (k, w) == (31, 15):
            continue
        bases, off = O.concat_records([r.encode() for r in c["ref"]])
        keys = gpu.index_build(bases, off, k, w, 0.0, make_resident=True)         # the index is built on the GPU too
        assert np.array_equal(keys, O.index_build((bases, off), k, w).keys()), c["name"]
        recs = [r.encode() for r in c["reads"]]
        rb, ro = O.concat_records(recs)
        kk, hh, tt = gpu.filter_batch(rb, ro, abs_threshold=c["abs"], rel_threshold=c["rel"], deplete=c["deplete"])
        assert list(map(int, kk)) == c["expect_keep"], c["name"]
        ran += 1
    assert ran == 3


@pytest.mark.parametrize("k,w", [(31, 15)] + GENERIC_KW)
def test_extract_matches_oracle(gpu, k, w):
    """dcn_extract (B3): get_minimizer_hashes_and_positions (src/filter_common.rs:211-310) and
    fill_minimizer_hashes (src/minimizers.rs:125-191) per record, hashes AND positions, in order."""
    _, recs = _generic_records(500 + k)
    bases, off = H.concat(recs)
    for prefix in (0, 90):
        h, p, oo = gpu.extract(bases, off, 0, k, w, prefix)
        for i, r in enumerate(recs):
            wh, wp = O.extract_filter(r, k, w, prefix)
            a, b = int(oo[i]), int(oo[i + 1])
            assert np.array_equal(h[a:b], wh) and np.array_equal(p[a:b], wp), (k, w, prefix, i)
    for thr in (0.0, 0.5):
        h, _, oo = gpu.extract(bases, off, 1, k, w, 0, thr)
        for i, r in enumerate(recs):
            assert np.array_equal(h[int(oo[i]):int(oo[i + 1])], O.extract_index(r, k, w, thr)), (k, w, thr, i)


@pytest.mark.parametrize("case", [c for c in CASES.make_cases() if c["name"] in (
    "single_150_search", "ragged_0_420_N_lower", "ragged_prefix_80", "ragged_prefix_20_below_k", "lengths_around_k_and_l",
    "trailing_newline", "trailing_newline_prefix", "tiny_records", "empty_records_only", "low_complexity", "dense_picks_polyA",
    "non_acgt_bytes", "units_up_to_1024")], ids=lambda c: c["name"])
def test_extract_tile_pipeline_matches_oracle(gpu, case):
    """dcn_extract's fast path (short records, k=31 w=15: phases 1-4 of the fused kernel + CSR compaction): hashes and
    positions per record == get_minimizer_hashes_and_positions (src/filter_common.rs:211-310), and == the generic kernels."""
    recs = [r for r in case["records"] if len(r) <= 1024]
    bases, off = H.concat(recs)
    h, p, oo = gpu.extract(bases, off, 0, 31, 15, case["prefix"])
    for i, r in enumerate(recs):
        wh, wp = O.extract_filter(r, 31, 15, case["prefix"])
        a, b = int(oo[i]), int(oo[i + 1])
        assert np.array_equal(h[a:b], wh) and np.array_equal(p[a:b], wp), (case["name"], i)
    os.environ["DCN_EXTRACT_GENERIC"] = "1"
    try:
        h2, p2, oo2 = gpu.extract(bases, off, 0, 31, 15, case["prefix"])
    finally:
        del os.environ["DCN_EXTRACT_GENERIC"]
    assert np.array_equal(h, h2) and np.array_equal(p, p2) and np.array_equal(oo, oo2)


def test_extract_then_lookup_equals_filter(gpu):
    """Client/server split of the batch engine (src/remote_filter.rs:762-790): B3 on the client, B2 on the server,
    must reproduce B1's decisions - all three on the GPU."""
    from deacon_server_b200 import IndexHeader
    g = H.random_genome(500_000, 131)
    idx = O.index_build([g], 31, 15, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    reads = H.sample_reads(g, 60_000, 150, 132)
    bases, off = H.concat(reads)
    h, p, oo = gpu.extract(bases, off, 0, 31, 15, 0)
    pair_off = oo[::2].copy()                                   # pooled hash list per pair (src/filter_common.rs:312-348)
    k2, h2, t2 = gpu.lookup_batch(h, pair_off, 2, 0.01, True)
    k1, h1, t1 = gpu.filter_batch(bases, off, paired=True, deplete=True)
    assert np.array_equal(k1, k2) and np.array_equal(h1, h2) and np.array_equal(t1, t2)


def test_extract_single_record_api_and_overflow(gpu):
    from deacon_server_b200 import DeaconCudaError
    g = H.random_genome(5000, 77)
    h, p = gpu.get_minimizer_hashes_and_positions(g, 0, 31, 15)
    wh, wp = O.extract_filter(g, 31, 15, 0)
    assert np.array_equal(h, wh) and np.array_equal(p, wp)
    assert np.array_equal(gpu.compute_minimizer_hashes(g, 31, 15, 0.0), O.extract_index(g, 31, 15, 0.0))
    bases, off = H.concat([g])
    with pytest.raises(DeaconCudaError, match="out_cap"):
        gpu.extract(bases, off, 0, 31, 15, cap=10)


@pytest.mark.parametrize("k,w,thr", [(21, 11, 0.0), (41, 15, 0.5), (57, 1, 0.0), (15, 5, 0.7)])
def test_index_build_generic_kw(gpu, k, w, thr):
    """`deacon index build -k K -w W [-e E]` for non-default parameters == the oracle's key set."""
    g = H.random_genome(400_000, 600 + k)
    g[1000:1300] = ord("A")
    g[5000:5400] = np.frombuffer(b"AC" * 200, np.uint8)
    recs = [g[:250_000].copy(), g[250_000:250_020].copy(), np.zeros(0, np.uint8), g[250_020:].copy(),
            np.frombuffer(b"ACGTNNNNRYKMacgtnryk" * 500, np.uint8).copy()]
    bases, off = H.concat(recs)
    keys = gpu.index_build(bases, off, k, w, thr, make_resident=True)
    want = O.index_build((bases, off), k, w, entropy=thr, threads=8).keys()
    assert np.array_equal(keys, want)
    info = gpu.index_info()
    assert info["n_keys"] == len(want) and info["kmer_length"] == k and info["window_size"] == w


# --------------------------------------------------------------------------- B2: pre-hashed records
def _hash_lists(idx, g, seed, n_rec, paired):
    """Per-record hash lists as the reference client builds them (src/remote_filter.rs:762-774,
    963-975): oracle extraction of simulated reads, mates pooled for pairs."""
    rng = np.random.default_rng(seed)
    reads = H.sample_reads(g, n_rec * (2 if paired else 1), (0, 400), seed)
    lists = []
    for i in range(n_rec):
        if paired:
            h = np.concatenate([O.extract_filter(reads[2 * i])[0], O.extract_filter(reads[2 * i + 1])[0]])
        else:
            h = O.extract_filter(reads[i])[0]
        if i % 7 == 0 and len(h):          # repeated hashes must count once (src/filter_common.rs:143-145)
            h = np.concatenate([h, h[::-1], h[:3]])
        if i % 11 == 0:                    # a few values that are not in the index
            h = np.concatenate([h, rng.integers(0, 2**63, 5).astype(np.uint64)])
        lists.append(np.ascontiguousarray(h, np.uint64))
    return lists


@pytest.mark.parametrize("paired", [False, True])
def test_lookup_batch_matches_oracle(gpu, paired):
    """dcn_lookup_batch == unpaired_should_keep / paired_should_keep (src/remote_filter.rs:230-301)."""
    from deacon_server_b200 import IndexHeader
    g = H.random_genome(300_000, 91)
    idx = O.index_build([g], 31, 15, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    lists = _hash_lists(idx, g, 92 + int(paired), 3000, paired)
    lists[5] = np.zeros(0, np.uint64)                                   # empty record
    keys = idx.keys()
    lists[6] = np.tile(keys[:40], 3)                                   # 120 hashes, 40 distinct hits
    lists[7] = np.concatenate([keys[:33], keys[:33]])                  # duplicates straddle a warp round
    big = np.concatenate([keys[:3000], keys[1000:2500], np.arange(1, 2000, dtype=np.uint64)])
    lists[8] = big                                                     # > 1024 hashes: global (hash, record) set
    lists[9] = np.tile(keys[100:1500], 2)
    off = np.zeros(len(lists) + 1, np.uint64)
    off[1:] = np.cumsum([len(x) for x in lists], dtype=np.uint64)
    hashes = np.concatenate(lists)
    for (a, r, dep) in ((2, 0.01, False), (2, 0.01, True), (1, 0.0, False), (3, 0.5, True)):
        k, h, t = gpu.lookup_batch(hashes, off, a, r, dep)
        ok, oh, ot = O.lookup_batch(idx, hashes, off, a, r, dep, threads=8)
        assert np.array_equal(t, ot) and np.array_equal(h, oh) and np.array_equal(k, ok), (a, r, dep)
    assert int(h[6]) == 40 and int(h[7]) == 33 and int(h[8]) == 3000


def test_lookup_equals_filter_on_same_reads(gpu):
    """B1 (raw sequences) and B2 (pre-hashed, oracle-extracted) must give the same decisions."""
    from deacon_server_b200 import IndexHeader
    g = H.random_genome(200_000, 95)
    idx = O.index_build([g], 31, 15, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    reads = H.sample_reads(g, 4000, 150, 96)
    bases, off = H.concat(reads)
    k1, h1, t1 = gpu.filter_batch(bases, off, paired=True, deplete=True)
    lists = [np.concatenate([O.extract_filter(reads[2 * i])[0], O.extract_filter(reads[2 * i + 1])[0]]) for i in range(2000)]
    res = gpu.paired_should_keep([(x, [], []) for x in lists], 31, 2, 0.01, True)
    assert [r[0] for r in res] == [bool(x) for x in k1]
    assert [r[1] for r in res] == [int(x) for x in h1] and [r[2] for r in res] == [int(x) for x in t1]


# --------------------------------------------------------------------------- .idx codec + union / diff on the GPU
def _mixed_keys(seed, n):
    rng = np.random.default_rng(seed)
    return np.unique(rng.integers(0, 2**64, n, dtype=np.uint64))


def test_idx_decode_encode_matches_oracle_codec(gpu):
    """dcn_idx_decode / dcn_idx_encode == the oracle's restatement of src/index.rs:57-72,130-164, for the usual
    all-9-byte body (GPU decode) and for bodies with short varints (sequential scan)."""
    for keys in (_mixed_keys(1, 200_000),
                 np.array([0, 5, 250, 251, 70000, 2**32 - 1, 2**32, 2**64 - 1], np.uint64),
                 np.concatenate([np.arange(0, 300, dtype=np.uint64), _mixed_keys(2, 5000)]),
                 np.zeros(0, np.uint64)):
        keys = np.unique(keys)
        data = O.idx_encode(np.random.default_rng(3).permutation(keys), 31, 15)       # a set in arbitrary order, like the reference writes
        hdr, n_file, n_set = gpu.idx_decode(data)
        assert (hdr.format_version, hdr.kmer_length, hdr.window_size) == (2, 31, 15) and n_file == n_set == len(keys)
        assert np.array_equal(gpu.working_keys(), keys)
        out = gpu.idx_encode()
        assert out == O.idx_encode(keys, 31, 15)                                       # byte-identical to the oracle's writer on sorted keys
        ver, k, w, back = O.idx_decode(out)
        assert (ver, k, w) == (2, 31, 15) and np.array_equal(back, keys)
    from deacon_server_b200 import DeaconCudaError
    with pytest.raises(DeaconCudaError, match="version"):
        gpu.idx_decode(bytes([1, 31, 15, 0]))
    with pytest.raises(DeaconCudaError, match="truncated|malformed"):
        gpu.idx_decode(O.idx_encode(_mixed_keys(4, 100), 31, 15)[:-3])


def test_key_sort_prefix_runs_and_fallback(gpu):
    """The working key set is sorted by five radix passes over bits 24..63 and a fix-up of the runs of keys that share
    those bits (sort_prefix_runs_kernel); keys that are not hash-like make runs too long for it and get all eight passes.
    Both must give sorted unique keys: hash-like keys with planted prefix collisions (runs of 2 .. 20, duplicates among
    them), and sequential keys (one run of a million)."""
    rng = np.random.default_rng(81)
    base = rng.integers(0, 2**63, 300_000, dtype=np.uint64) * np.uint64(2)
    planted = []
    for p in base[:3000]:
        m = int(rng.integers(2, 21))
        low = rng.integers(0, 1 << 24, m, dtype=np.uint64)
        planted.append((p & np.uint64(0xFFFFFFFFFF000000)) | low)
        planted.append(planted[-1][: m // 2])                      # duplicates inside the run
    hashy = np.concatenate([base] + planted)
    for keys in (hashy, np.arange(5, 1_000_005, dtype=np.uint64), np.concatenate([hashy, np.arange(0, 100_000, dtype=np.uint64)])):
        data = O.idx_encode(rng.permutation(keys), 31, 15)
        hdr, n_file, n_set = gpu.idx_decode(data)
        want = np.unique(keys)
        assert n_file == len(keys) and n_set == len(want)
        assert np.array_equal(gpu.working_keys(), want)


def test_index_union_and_diff_match_set_algebra(gpu):
    """index::union (src/index.rs:563-664) and index::diff index - index (:421-537)."""
    a, b, c = _mixed_keys(11, 300_000), _mixed_keys(12, 200_000), _mixed_keys(13, 1000)
    b = np.unique(np.concatenate([b, a[::3]]))          # overlap
    gpu.idx_decode(O.idx_encode(a, 31, 15))
    assert gpu.index_union(O.idx_encode(b, 31, 15)) == len(np.union1d(a, b))
    assert gpu.index_union(O.idx_encode(c, 31, 15)) == len(np.union1d(np.union1d(a, b), c))
    assert np.array_equal(gpu.working_keys(), np.union1d(np.union1d(a, b), c))
    gpu.idx_decode(O.idx_encode(a, 31, 15))
    assert gpu.index_diff(O.idx_encode(b, 31, 15)) == len(np.setdiff1d(a, b))
    assert np.array_equal(gpu.working_keys(), np.setdiff1d(a, b))
    assert gpu.idx_encode() == O.idx_encode(np.setdiff1d(a, b), 31, 15)
    from deacon_server_b200 import DeaconCudaError
    with pytest.raises(DeaconCudaError, match="Incompatible headers"):
        gpu.index_union(O.idx_encode(c, 21, 11))
    # the working set becomes the resident index and answers lookups
    gpu.index_make_resident()
    d = np.setdiff1d(a, b)
    q = np.concatenate([d[:50], b[:50]])
    k, h, t = gpu.lookup_batch(q, np.array([0, 50, 100], np.uint64), 1, 0.0, False)
    assert list(map(int, h)) == [50, int(np.isin(b[:50], d).sum())]


@pytest.mark.parametrize("k,w", [(31, 15), (21, 11)])
def test_index_diff_against_sequences(gpu, k, w):
    """index diff IDX reads.fa (stream_diff_fastx, src/index.rs:311-419) == index - index of the same sequences
    (the equality tests/index_tests.rs:168-285 checks)."""
    g = H.random_genome(300_000, 21 + k)
    second = [g[50_000:120_000].copy(), g[200_000:200_040].copy(), np.frombuffer(b"ACGTNNNNRYKMacgtnryk" * 100, np.uint8).copy()]
    bases, off = H.concat([g])
    a = gpu.index_build(bases, off, k, w, 0.0, make_resident=False)
    sb, so = H.concat(second)
    remaining = gpu.index_diff_sequences(sb, so)
    want = np.setdiff1d(a, O.index_build((sb, so), k, w).keys())
    assert remaining == len(want) and np.array_equal(gpu.working_keys(), want)
    assert len(want) < len(a)


# --------------------------------------------------------------------------- ABI contract: errors, contexts
def test_error_codes_and_messages(gpu):
    """Every failure is an error code + dcn_last_error text (the reference returns anyhow::Result or panics;
    the library never aborts and never falls back to the CPU)."""
    import ctypes as C
    import deacon_server_b200 as d
    from deacon_server_b200 import DeaconCudaError, IndexHeader
    lib = d.load()
    fresh = d.DeaconGpu(0)
    try:
        b = np.frombuffer(b"ACGT" * 40, np.uint8)
        off = np.array([0, 160], np.uint64)
        with pytest.raises(DeaconCudaError) as e:
            fresh.filter_batch(b, off)
        assert e.value.code == -3 and "no index resident" in str(e.value)                     # DCN_ERR_NO_INDEX
        with pytest.raises(DeaconCudaError) as e:
            fresh.lookup_batch(np.array([1], np.uint64), np.array([0, 1], np.uint64))
        assert e.value.code == -3
        with pytest.raises(DeaconCudaError) as e:
            fresh.idx_encode()
        assert e.value.code == -3
        fresh.index_upload(np.array([5, 6, 7], np.uint64), IndexHeader(2, 31, 15))
        with pytest.raises(DeaconCudaError) as e:
            fresh.filter_batch(np.tile(b, 3), np.array([0, 160, 320, 480], np.uint64), paired=True)
        assert e.value.code == -2 and "even record count" in str(e.value)                       # DCN_ERR_ARG
        k = np.zeros(1, np.uint8)
        assert lib.dcn_filter_batch(fresh._ctx, b.ctypes.data, off.ctypes.data, 1, 0, 0, 2, 0.01, 0, None, None, None) == -2
        assert lib.dcn_filter_batch(None, b.ctypes.data, off.ctypes.data, 1, 0, 0, 2, 0.01, 0, k.ctypes.data, None, None) == -2
        with pytest.raises(DeaconCudaError, match="load factor"):
            fresh.set_load_factor(0.99)
        with pytest.raises(DeaconCudaError, match="unknown flavour"):
            fresh.extract(b, off, flavour=7)
        # zero records is a no-op, not an error
        kk, hh, tt = fresh.filter_batch(np.zeros(0, np.uint8), np.array([0], np.uint64))
        assert len(kk) == 0
        assert lib.dcn_ctx_create(99) is None and b"out of range" in lib.dcn_last_error(None)
    finally:
        fresh.close()


def test_two_contexts_are_independent(gpu):
    """One ctx per index / per worker (SURVEY 8b threading): different tables, same GPU, interleaved calls."""
    import deacon_server_b200 as d
    from deacon_server_b200 import IndexHeader
    g1, g2 = H.random_genome(80_000, 201), H.random_genome(80_000, 202)
    i1, i2 = O.index_build([g1], 31, 15), O.index_build([g2], 31, 15)
    other = d.DeaconGpu(0)
    try:
        gpu.index_upload(i1.keys(), IndexHeader(2, 31, 15))
        other.index_upload(i2.keys(), IndexHeader(2, 31, 15))
        reads = H.sample_reads(g1, 1500, 150, 203) + H.sample_reads(g2, 1500, 150, 204)
        bases, off = H.concat(reads)
        for _ in range(2):
            a = gpu.filter_batch(bases, off)
            b = other.filter_batch(bases, off)
            assert np.array_equal(a[1], O.filter_batch(i1, bases, off)[1]) and np.array_equal(b[1], O.filter_batch(i2, bases, off)[1])
        assert int(a[1][:1500].sum()) > 10 * int(a[1][1500:].sum()) and int(b[1][1500:].sum()) > 10 * int(b[1][:1500].sum())
    finally:
        other.close()


def test_debug_hit_kmers_match_reference_semantics(gpu):
    """--debug: the k-mers of the counted hits, in hash-list order (sequence_matches, src/filter_common.rs:129-155),
    and the DEBUG line of the default engine (src/local_filter.rs:354-363)."""
    from deacon_server_b200 import IndexHeader
    g = H.random_genome(50_000, 301)
    idx = O.index_build([g], 31, 15)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    read = np.concatenate([g[1000:1150], g[1000:1100]])            # the second part repeats minimizers of the first
    read[60] = ord("N")
    hashes, pos = O.extract_filter(read, 31, 15, 0)
    want, seen = [], set()
    for h, p in zip(hashes, pos):                                   # the reference's loop, restated
        if int(h) in idx and int(h) not in seen:
            seen.add(int(h))
            want.append(bytes(read[int(p):int(p) + 31]).decode())
    (keep, hits, total, kmers), = gpu.unpaired_should_keep([(hashes, pos, read)], 31, 2, 0.01, False, debug=True)
    assert kmers == want and hits == len(want) and total == len(hashes) and len(want) < len(hashes)
    keep2, hits2, total2, kmers2, line = gpu.should_keep_sequence_debug("read1", read)
    assert (keep2, hits2, total2, kmers2) == (keep, hits, total, kmers)
    assert line == f"DEBUG: read1 hits={hits}/{total} keep=true kmers=[{','.join(want)}]"
    # flags of many records at once agree with per-record oracle counts
    reads = H.sample_reads(g, 500, 150, 302)
    lists = [O.extract_filter(r, 31, 15, 0)[0] for r in reads]
    off = np.zeros(len(lists) + 1, np.uint64)
    off[1:] = np.cumsum([len(x) for x in lists], dtype=np.uint64)
    k, h, t, fl = gpu.lookup_batch_flags(np.concatenate(lists), off)
    ok, oh, ot = O.lookup_batch(idx, np.concatenate(lists), off)
    assert np.array_equal(h, oh) and np.array_equal(np.add.reduceat(fl.astype(np.uint32), off[:-1].astype(np.int64))[oh > 0], oh[oh > 0])


def test_extract_size_query_and_working_set_release(gpu):
    """dcn_extract with out_cap = 0 and no output arrays is the documented size query: DCN_ERR_OVERFLOW with the need in
    out_off[n_rec] -- also when exactly one minimizer comes out (the copy to a null array used to be attempted).
    dcn_working_set_release drops the key set and scratch; the resident table keeps answering."""
    import ctypes as C
    from deacon_server_b200 import IndexHeader
    g = H.random_genome(5_000, 91)
    one = g[:45].copy()                      # one window -> exactly one minimizer
    bases, off = H.concat([one])
    out_off = np.zeros(2, np.uint64)
    rc = gpu._lib.dcn_extract(gpu._ctx, 0, bases.ctypes.data, off.ctypes.data, 1, 31, 15, 0, C.c_float(0.0), None, None,
                              out_off.ctypes.data, 0)
    assert rc == -6 and int(out_off[1]) == 1          # DCN_ERR_OVERFLOW, need = 1
    h, p, oo = gpu.extract(bases, off, 0, 31, 15, 0)
    assert len(h) == 1 and np.array_equal(h, O.extract_filter(one, 31, 15, 0)[0])
    idx = O.index_build([g], 31, 15)
    blob = O.idx_encode(idx.keys(), 31, 15)
    gpu.idx_decode(blob, mode=0, make_resident=True)
    assert gpu.working_set_info()["n_keys"] == len(idx)
    gpu.working_set_release()
    assert gpu.working_set_info()["n_keys"] == 0
    reads = H.sample_reads(g, 200, 150, 92)
    b2, o2 = H.concat(reads)
    k, hh, t = gpu.filter_batch(b2, o2)
    ok, oh, ot = O.filter_batch(idx, b2, o2)
    assert np.array_equal(k, ok) and np.array_equal(hh, oh) and np.array_equal(t, ot)

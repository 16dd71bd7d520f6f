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


def test_config1_shape_and_counters(gpu):
    """BASELINE config 1 shape, scaled: single-end 150 bp reads vs a random reference, -a 2 -r 0.01;
    also checks the six ProcessingStats counters (src/local_filter.rs:179-187, 347-371)."""
    from deacon_server_b200 import IndexHeader
    g = H.random_genome(1_000_000, 1)
    idx = O.index_build([g], 31, 15, threads=8)
    gpu.index_upload(idx.keys(), IndexHeader(2, 31, 15))
    reads = H.sample_reads(g, 100_000, 150, 2)
    bases, off = H.concat(reads)
    gpu.stats_reset()
    k, h, t = gpu.filter_batch(bases, off)
    ok, oh, ot = O.filter_batch(idx, bases, off, threads=8)
    assert np.array_equal(t, ot) and np.array_equal(h, oh) and np.array_equal(k, ok)
    st = gpu.stats()
    lens = np.diff(off).astype(np.int64)
    assert st["total_seqs"] == len(reads) and st["total_bp"] == int(lens.sum())
    assert st["output_seq_counter"] == int(ok.sum()) and st["filtered_seqs"] == len(reads) - int(ok.sum())
    assert st["output_bp"] == int(lens[ok.astype(bool)].sum()) and st["filtered_bp"] == int(lens[~ok.astype(bool)].sum())


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

"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/deacon_cuda.h
declares (no compute calls: there is no GPU here), and the host-side mirror logic (.idx codec,
classification rule) agrees with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import deacon_server_b200 as D
from deacon_server_b200 import _lib, api
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "deacon_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = D.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libdeacon_cuda.so does not export {n}"
    assert sorted(_lib.SIGNATURES) == names, "python binding table out of sync with the header"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(D.DeaconCudaError, match="no CUDA device|no CPU fallback|device"):
        D.DeaconGpu(0)


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "deacon_server_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} references the oracle"


def test_idx_codec_matches_oracle_codec(tmp_path):
    rng = np.random.default_rng(5)
    keys = np.unique(rng.integers(0, 2**63, 3000, dtype=np.uint64) * np.uint64(2) + np.uint64(1))
    hdr = api.IndexHeader(2, 31, 15)
    assert api.encode_index(keys, hdr) == O.idx_encode(keys, 31, 15)
    mixed = np.array([0, 5, 250, 251, 70000, 2**32, 2**64 - 1], np.uint64)
    assert api.encode_index(mixed, hdr) == O.idx_encode(mixed, 31, 15)
    got, h2 = api.decode_index(O.idx_encode(mixed, 21, 11))
    assert np.array_equal(got, mixed) and (h2.kmer_length, h2.window_size) == (21, 11)
    p = tmp_path / "x.idx"
    api.write_minimizers(keys[::-1], hdr, p)
    got, h3 = api.load_minimizer_hashes(p)
    assert np.array_equal(got, keys) and h3.format_version == 2
    with pytest.raises(ValueError, match="Unsupported index format version"):
        api.decode_index(bytes([1, 31, 15, 0]))


def test_host_classification_matches_oracle():
    for total in list(range(0, 300)) + [10**6, 2**31]:
        for rel in (0.0, 0.01, 0.015, 0.25, 0.5, 1.0):
            for abs_ in (0, 1, 2, 7):
                assert api.calculate_required_hits(abs_, rel, total) == O.required_hits(abs_, rel, total)
    assert api.meets_filtering_criteria(1, 48, 2, 0.01, True) is True    # tests/filter_tests.rs:943-1015
    assert api.meets_filtering_criteria(0, 0, 2, 0.01, False) is False


def test_pack_records_matches_the_two_separate_passes():
    """dcn_pack_records (what the pipeline's packer threads run per chunk; no GPU needed) == dcn_pack_ascii +
    dcn_newline_bits: the newline flags come from the packer's list of blocks with a non-ACGT byte, so every place a
    flag can hide is exercised -- newline-terminated records of every length around k and the prefix, newlines in the
    middle of records, runs of N, empty records, records ending on 32- and 64-base block edges."""
    rng = np.random.default_rng(12)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    for n_rec, prefix in ((0, 0), (1, 0), (700, 0), (700, 40), (5000, 0), (5000, 100)):
        lens = rng.integers(0, 130, n_rec).astype(np.uint64)
        if n_rec > 100:
            lens[::7] = 64                                              # block-edge endings
            lens[3::11] = 0
            lens[5::13] = rng.integers(28, 36, len(lens[5::13]))        # around k
        off = np.zeros(n_rec + 1, np.uint64)
        off[1:] = np.cumsum(lens)
        nb = int(off[-1])
        b = acgt[rng.integers(0, 4, nb)] if nb else np.zeros(0, np.uint8)
        if nb:
            b[rng.integers(0, nb, nb // 40)] = ord("N")
            b[rng.integers(0, nb, nb // 60)] = 10                       # newlines anywhere
            ends = off[1:][lens > 0] - np.uint64(1)
            b[ends[rng.random(len(ends)) < 0.5].astype(np.int64)] = 10  # half of the records end in one
            pre = (off[:-1] + np.uint64(prefix) - np.uint64(1))[lens > prefix] if prefix else np.zeros(0, np.uint64)
            b[pre[rng.random(len(pre)) < 0.5].astype(np.int64)] = 10    # ... or their prefix does
        codes, inv, nl = api.pack_records(b, off, 31, prefix)
        want_c, want_i = api.pack_ascii(b)
        want_nl = api.newline_bits(b, off, 31, prefix)
        assert np.array_equal(codes, want_c) and np.array_equal(inv, want_i), (n_rec, prefix)
        assert np.array_equal(nl, want_nl), (n_rec, prefix)
        if n_rec >= 700:
            assert int(np.unpackbits(nl.view(np.uint8)).sum()) > n_rec // 8
        # the sparse form (dcn_pack_records_sparse): the same codes and flags, and a list that rebuilds the dense mask
        codes_s, exc, nl_s = api.pack_records_sparse(b, off, 31, prefix)
        assert np.array_equal(codes_s, want_c) and np.array_equal(nl_s, want_nl), (n_rec, prefix)
        dense = np.zeros(len(want_i) // 2, np.uint32)
        assert np.all(np.diff(exc[:, 0].astype(np.int64)) > 0)          # ascending, every block once
        dense[exc[:, 0]] = exc[:, 1]
        assert np.array_equal(dense.view(np.uint16), want_i), (n_rec, prefix)
        assert np.all(exc[:, 1] != 0)                                   # only blocks that hold a non-ACGT base (or padding)


def test_product_packer_matches_definition():
    """dcn_pack_ascii (the ingest stage of dcn_filter_batch; runs on the host, no GPU needed):
    code = (byte >> 1) & 3 (src/filter_common.rs:238), non-ACGT mask (:245-258)."""
    rng = np.random.default_rng(11)
    for n in (0, 1, 31, 32, 33, 100_003):
        b = rng.integers(0, 256, n).astype(np.uint8)
        if n > 1000:
            b[: n // 2] = np.frombuffer(b"ACGTacgtN", np.uint8)[rng.integers(0, 9, n // 2)]
        codes, inv = api.pack_ascii(b)
        pad = np.zeros(len(codes) * 16, np.uint8)
        pad[:n] = b
        code = ((pad >> 1) & 3).astype(np.uint32).reshape(-1, 16)
        want_c = (code << (2 * np.arange(16, dtype=np.uint32))).sum(axis=1).astype(np.uint32)
        bad = ~np.isin(pad & 0xDF, np.frombuffer(b"ACGT", np.uint8))
        want_i = (bad.reshape(-1, 16).astype(np.uint32) << np.arange(16, dtype=np.uint32)).sum(axis=1).astype(np.uint16)
        assert np.array_equal(codes, want_c) and np.array_equal(inv, want_i), n

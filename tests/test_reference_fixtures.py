"""Parity pin: the oracle (CPU, here) and the CUDA path (GPU box) against outputs of the UNMODIFIED reference binary.

The fixtures are written by tools/make_reference_fixtures.sh (needs cargo; this image has none) into
tests/golden/reference_c1/.  While they are absent every test here is skipped and parity for minimizer selection
stays "unpinned" (DESIGN.md 2).  What is compared, per SURVEY.md 8c:
  * .idx files   -> 3-byte header, count, key SET  (src/index.rs:130-164)                      vs oracle index build
  * DEBUG lines  -> per record / pair hits, total, keep and the hit k-mers in order
                    (src/local_filter.rs:354-363, 424-434; src/filter_common.rs:129-198)        vs oracle / GPU filter
  * summary JSON -> seqs_in / seqs_out / bp_in / bp_out (src/filter_common.rs:10-38)            vs the kept flags
DCN_REFERENCE_FIXTURES overrides the directory (the self-test below points it at oracle-made files to check this
file's own parsing, which proves nothing about parity).
"""
import gzip
import json
import os
import re

import numpy as np
import pytest

import helpers as H
from oracle import oracle as O

FIX = os.environ.get("DCN_REFERENCE_FIXTURES", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_c1"))
HAVE = os.path.exists(os.path.join(FIX, "k31w15.idx"))
need_fixtures = pytest.mark.skipif(not HAVE, reason="no reference-derived fixtures: run tools/make_reference_fixtures.sh on a box with cargo")

RUNS = [   # (debug file stem, index, reads, paired, prefix, abs, rel, deplete)
    ("single", "k31w15", ["reads_single"], False, 0, 2, 0.01, False),
    ("single_p80_deplete", "k31w15", ["reads_single"], False, 80, 2, 0.01, True),
    ("paired_deplete", "k31w15", ["reads_r1", "reads_r2"], True, 0, 2, 0.01, True),
    ("long", "k31w15", ["reads_long"], False, 0, 2, 0.01, False),
    ("single_k41", "k41w15", ["reads_single"], False, 0, 1, 0.01, False),
    ("single_k21", "k21w11", ["reads_single"], False, 0, 2, 0.01, False),
]
INDEXES = [("k31w15", 31, 15, 0.0), ("k31w15_e05", 31, 15, 0.5), ("k41w15", 41, 15, 0.0), ("k21w11", 21, 11, 0.0)]


def _open(stem):
    p = os.path.join(FIX, stem)
    return gzip.open(p + ".gz", "rb") if os.path.exists(p + ".gz") else open(p, "rb")


def read_fastx(stem):
    """-> [(id, sequence bytes)]; FASTA may be multi-line (needletail joins the lines, src/index.rs:225-246)."""
    out = []
    with _open(stem) as f:
        lines = f.read().split(b"\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith(b"@"):
            out.append((ln[1:].decode(), lines[i + 1]))
            i += 4
        elif ln.startswith(b">"):
            j = i + 1
            while j < len(lines) and not lines[j].startswith(b">"):
                j += 1
            out.append((ln[1:].decode(), b"".join(lines[i + 1:j])))
            i = j
        else:
            i += 1
    return out


def load_idx(name):
    with open(os.path.join(FIX, name + ".idx"), "rb") as f:
        return O.idx_decode(f.read())


DEBUG_RE = re.compile(r"^DEBUG: (.*) hits=(\d+)/(\d+) keep=(true|false) kmers=\[(.*)\]$")


def read_debug(stem):
    out = {}
    order = []
    with _open(stem + "_debug.txt") as f:
        for ln in f.read().decode().splitlines():
            m = DEBUG_RE.match(ln)
            assert m, f"unparsable DEBUG line: {ln[:120]}"
            kmers = m.group(5).split(",") if m.group(5) else []
            out[m.group(1)] = (int(m.group(2)), int(m.group(3)), m.group(4) == "true", kmers)
            order.append(m.group(1))
    return out, order


def expected_units(run, index_keys, k, w):
    """What the oracle says for every unit of a run: (id, hits, total, keep, hit k-mers in order)."""
    stem, _idx, reads, paired, prefix, abs_thr, rel, deplete = run
    recs = [read_fastx(r + ".fq") for r in reads]
    keyset = set(int(x) for x in index_keys)
    idx = O.IndexSet(np.asarray(index_keys, np.uint64))
    units = []
    n = len(recs[0])
    for i in range(n):
        mates = [r[i] for r in recs]
        seen, kmers, total = set(), [], 0
        for _id, seq in mates:
            hs, ps = O.extract_filter(seq, k, w, prefix)
            eff = seq[:prefix] if (prefix > 0 and len(seq) > prefix) else seq
            total += len(hs)
            for h, p in zip(hs, ps):
                h = int(h)
                if h in keyset and h not in seen:
                    seen.add(h)
                    kmers.append(eff[int(p):int(p) + k].decode())
        req = O.required_hits(abs_thr, rel, total)
        keep = (len(seen) < req) if deplete else (len(seen) >= req)
        units.append(("/".join(m[0] for m in mates), len(seen), total, keep, kmers, sum(len(m[1]) for m in mates)))
    # the batch call must agree with the per-record walk above (it is what the GPU is compared with elsewhere)
    flat = [np.frombuffer(m[1], np.uint8) for i in range(n) for m in [r[i] for r in recs]]
    bases, off = H.concat(flat)
    bk, bh, bt = O.filter_batch(idx, bases, off, paired=paired, prefix_len=prefix, k=k, w=w, abs_thr=abs_thr, rel_thr=rel, deplete=deplete)
    assert [u[1] for u in units] == list(map(int, bh)) and [u[2] for u in units] == list(map(int, bt))
    assert [u[3] for u in units] == [bool(x) for x in bk]
    return units, (bases, off)


def check_run(run, units):
    stem, _idx, _reads, paired, *_ = run
    dbg, _order = read_debug(stem)
    n_lines = 0
    for uid, hits, total, keep, kmers, _bp in units:
        if paired and hits == 0:
            assert uid not in dbg, f"{uid}: the reference prints pairs with hits only"
            continue
        assert uid in dbg, f"{uid}: no DEBUG line"
        n_lines += 1
        rh, rt, rk, rkm = dbg[uid]
        assert (rh, rt) == (hits, total), f"{uid}: reference hits/total {rh}/{rt}, ours {hits}/{total}"
        assert rk == keep, f"{uid}: keep differs"
        assert rkm == kmers, f"{uid}: hit k-mers (i.e. minimizer positions) differ"
    assert n_lines == len(dbg)


def check_summary(run, units):
    stem, *_ , = run
    paired = run[3]
    with open(os.path.join(FIX, stem + "_summary.json")) as f:
        s = json.load(f)
    per = 2 if paired else 1
    assert s["seqs_in"] == per * len(units)
    assert s["seqs_out"] == per * sum(1 for u in units if u[3])
    assert s["bp_in"] == sum(u[5] for u in units)
    assert s["bp_out"] == sum(u[5] for u in units if u[3])


@need_fixtures
@pytest.mark.parametrize("name,k,w,entropy", INDEXES, ids=[x[0] for x in INDEXES])
def test_reference_index_keyset_equals_oracle(name, k, w, entropy):
    ver, rk, rw, keys = load_idx(name)
    assert (ver, rk, rw) == (2, k, w)
    genome = [np.frombuffer(s, np.uint8) for _id, s in read_fastx("genome.fa")]
    want = O.index_build(genome, k, w, entropy).keys()
    assert len(keys) == len(np.unique(keys)), "the reference writes a set"
    assert np.array_equal(np.sort(keys), want), "oracle index != reference index (key set)"


@need_fixtures
@pytest.mark.parametrize("run", RUNS, ids=[r[0] for r in RUNS])
def test_reference_debug_lines_equal_oracle(run):
    _ver, k, w, keys = load_idx(run[1])
    units, _ = expected_units(run, keys, k, w)
    check_run(run, units)
    check_summary(run, units)


@need_fixtures
@pytest.mark.gpu
@pytest.mark.parametrize("run", RUNS, ids=[r[0] for r in RUNS])
def test_reference_debug_lines_equal_gpu(gpu, run):
    """The same comparison with the CUDA path through the C ABI in place of the oracle's numbers."""
    from deacon_server_b200 import IndexHeader
    _ver, k, w, keys = load_idx(run[1])
    units, (bases, off) = expected_units(run, keys, k, w)
    stem, _idx, _reads, paired, prefix, abs_thr, rel, deplete = run
    gpu.index_upload(np.asarray(keys, np.uint64), IndexHeader(2, k, w))
    gk, gh, gt = gpu.filter_batch(bases, off, paired=paired, prefix_length=prefix, abs_threshold=abs_thr, rel_threshold=rel, deplete=deplete)
    dbg, _ = read_debug(stem)
    for (uid, _h, _t, _k, _km, _bp), h, t, kp in zip(units, gh, gt, gk):
        if uid in dbg:
            assert dbg[uid][:3] == (int(h), int(t), bool(kp)), uid
        else:
            assert paired and int(h) == 0, uid


@need_fixtures
@pytest.mark.gpu
@pytest.mark.parametrize("name,k,w,entropy", INDEXES, ids=[x[0] for x in INDEXES])
def test_reference_index_keyset_equals_gpu_build(gpu, name, k, w, entropy):
    _ver, _k, _w, keys = load_idx(name)
    genome = [np.frombuffer(s, np.uint8) for _id, s in read_fastx("genome.fa")]
    bases, off = H.concat(genome)
    got = gpu.index_build(bases, off, k, w, entropy, make_resident=False)
    assert np.array_equal(np.sort(keys), np.sort(got))


@need_fixtures
def test_reference_fixtures_single_out_the_working_hypothesis():
    """With reference output in hand the 64-variant sweep (test_hypothesis_sweep.py) must collapse to ONE hypothesis:
    the working one and its algebraic twin ("A|C majority, rightmost on canonical" is the same function as "T|G
    majority, leftmost on canonical" because k + w - 1 is odd)."""
    from oracle import py_variants as V
    _ver, k, w, keys = load_idx("k31w15")
    keyset = set(int(x) for x in keys)
    dbg, order = read_debug("single")
    recs = dict(read_fastx("reads_single.fq"))
    survivors = []
    for v in V.ALL:
        ok = True
        for uid in order:
            seq = recs[uid]
            hs, ps = V.extract_filter(v, seq, k, w, 0)
            seen, kmers = set(), []
            for h, p in zip(hs, ps):
                if h in keyset and h not in seen:
                    seen.add(h)
                    kmers.append(seq[p:p + k].decode())
            if (len(seen), len(hs), kmers) != (dbg[uid][0], dbg[uid][1], dbg[uid][3]):
                ok = False
                break
        if ok:
            survivors.append(v)
    twin = V.Variant(tg_majority=False, left_on_canonical=False)
    assert set(survivors) == {V.WORKING, twin}, f"variants consistent with the reference: {[v.name() for v in survivors]}"


def test_this_file_on_oracle_made_fixtures(tmp_path):
    """Self-check of the parsers and conventions above: files in the kit's layout, written from the oracle's own output
    (tests/fake_reference.py), must pass every CPU test of this file.  Says nothing about parity."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    if os.environ.get("DCN_REFERENCE_FIXTURES"):
        pytest.skip("already running on substituted fixtures")
    out = str(tmp_path / "fake")
    subprocess.check_call([sys.executable, os.path.join(here, "fake_reference.py"), out])
    env = dict(os.environ, DCN_REFERENCE_FIXTURES=out)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "not gpu", os.path.abspath(__file__)], env=env,
                       capture_output=True, text=True, cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " skipped" in r.stdout and "11 passed" in r.stdout, r.stdout[-500:]

"""The C++ host driver end to end on a B200 (run with -m gpu): files in, files out, like the reference's own
command-line tests (tests/filter_tests.rs, tests/index_tests.rs).  Expected outputs are made here from the oracle's
decisions and the reference's record format (src/local_filter.rs:60-92): output records, summary counters and .idx
key sets must be identical."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "deacon_server_b200", "deacon-b200")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def run(*args, stdin=None, check=True):
    p = subprocess.run([BIN, *map(str, args)], input=stdin, capture_output=True)
    if check:
        assert p.returncode == 0, p.stderr.decode()
    return p


def fasta(records, width=0, names=None):
    out = b""
    for i, r in enumerate(records):
        r = bytes(r)
        out += b">" + (names[i] if names else b"seq%d" % i) + b"\n"
        out += (b"\n".join(r[j:j + width] for j in range(0, len(r), width)) if width and r else r) + b"\n"
    return out


def fastq_records(records, prefix=b"read"):
    return [(prefix + b"%d extra=%d" % (i, i * 7), bytes(r), bytes([33 + (i + j) % 40 for j in range(len(r))])) for i, r in enumerate(records)]


def fastq(recs):
    return b"".join(b"@" + i + b"\n" + s + b"\n+\n" + q + b"\n" for i, s, q in recs)


def expected_output(recs, keep_units, paired, rename=False, split=False):
    """format_record_to_buffer over the kept units, global rename counter."""
    outs = [b"", b""]
    counter = 0
    rpu = 2 if paired else 1
    for u, k in enumerate(keep_units):
        if not k:
            continue
        for m in range(rpu):
            i, s, q = recs[u * rpu + m]
            counter += 1
            name = str(counter).encode() if rename else i
            rec = (b"@" + name + b"\n" + s + b"\n+\n" + q + b"\n") if q is not None else (b">" + name + b"\n" + s + b"\n")
            outs[m if split else 0] += rec
    return outs


def oracle_keep(idx, seqs, paired, k=31, w=15, **kw):
    bases, off = H.concat([np.frombuffer(s, np.uint8) for s in seqs])
    return O.filter_batch(idx, bases, off, paired=paired, k=k, w=w, threads=8, **kw)


def test_reference_known_answers_through_the_cli(tmp_path):
    """tests/filter_tests.rs restated: index build -> filter, files on both sides."""
    cases = json.load(open(os.path.join(GOLD, "reference_kats.json")))["cases"]
    for c in cases:
        d = tmp_path / c["name"]
        d.mkdir()
        (d / "ref.fa").write_bytes(fasta([r.encode() for r in c["ref"]], width=10 if "multiline" in c["name"] else 0))
        run("index", "build", "-k", c["k"], "-w", c["w"], "-q", "-o", d / "ref.idx", d / "ref.fa")
        args = ["filter", d / "ref.idx"]
        paired = "reads1" in c
        if paired:
            r1, r2 = fastq_records([r.encode() for r in c["reads1"]], b"a"), fastq_records([r.encode() for r in c["reads2"]], b"b")
            (d / "r1.fq").write_bytes(fastq(r1))
            (d / "r2.fq").write_bytes(fastq(r2))
            recs = [x for pair in zip(r1, r2) for x in pair]
            args += [d / "r1.fq", d / "r2.fq"]
        else:
            recs = fastq_records([r.encode() for r in c["reads"]])
            (d / "r.fq").write_bytes(fastq(recs))
            args += [d / "r.fq"]
        args += ["-a", c["abs"], "-r", c["rel"], "-o", d / "out.fq", "-s", d / "summary.json", "-q"]
        if c["deplete"]:
            args.append("--deplete")
        run(*args)
        assert (d / "out.fq").read_bytes() == expected_output(recs, c["expect_keep"], paired)[0], c["name"]
        s = json.load(open(d / "summary.json"))
        rpu = 2 if paired else 1
        assert s["seqs_in"] == len(recs) and s["seqs_out"] == rpu * sum(c["expect_keep"]), c["name"]
        assert s["k"] == c["k"] and s["w"] == c["w"] and s["deplete"] == c["deplete"]


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    g = H.random_genome(400_000, 21)
    contigs = [g[:150_000], g[150_000:150_020], g[150_020:]]   # one contig shorter than k
    (d / "ref.fa").write_bytes(fasta(contigs, width=80, names=[b"chr1 test", b"tiny", b"chr2"]))
    p = run("index", "build", "-o", d / "ref.idx", d / "ref.fa")
    idx = O.index_build(contigs, 31, 15, threads=8)
    return {"dir": d, "genome": g, "contigs": contigs, "idx": idx, "build_stderr": p.stderr}


def test_index_build_file_matches_oracle_key_set(world):
    from deacon_server_b200 import api
    keys, hdr = api.decode_index((world["dir"] / "ref.idx").read_bytes())
    assert (hdr.format_version, hdr.kmer_length, hdr.window_size) == (2, 31, 15)
    assert np.array_equal(np.sort(keys), np.sort(world["idx"].keys()))
    err = world["build_stderr"].decode()
    assert "chr1 test (150000bp)" in err and f"Indexed {len(world['idx'])} minimizers from 3 sequence(s) (400000bp)" in err
    info = run("index", "info", world["dir"] / "ref.idx").stderr.decode()
    assert f"Distinct minimizer count: {len(world['idx'])}" in info and "K-mer length (k): 31" in info
    # k + w - 1 must be odd (src/index.rs:186-194)
    p = run("index", "build", "-k", 30, "-w", 15, world["dir"] / "ref.fa", check=False)
    assert p.returncode == 1 and b"must be odd" in p.stderr


@pytest.mark.parametrize("mode", ["search", "deplete_prefix", "rename_gz", "multi_batch"])
def test_single_end_outputs_identical(world, mode):
    d = world["dir"]
    reads = H.sample_reads(world["genome"], 6000, (20, 400), 31, lower_rate=0.02)
    recs = fastq_records(reads)
    (d / "reads.fq").write_bytes(fastq(recs))
    kw, args = {}, []
    if mode == "deplete_prefix":
        kw, args = dict(deplete=True, prefix_len=80, abs_thr=1), ["-d", "-p", 80, "-a", 1]
    out = d / ("out_%s.fq%s" % (mode, ".gz" if mode == "rename_gz" else ""))
    if mode == "rename_gz":
        args += ["-R", "--compression-level", 4]
    if mode == "multi_batch":
        args += ["--batch-mbp", 1, "-t", 3]
    run("filter", d / "ref.idx", d / "reads.fq", "-o", out, "-s", d / "s.json", *args)
    keep, hits, total = oracle_keep(world["idx"], [r[1] for r in recs], False, **kw)
    want = expected_output(recs, keep, False, rename=(mode == "rename_gz"))[0]
    got = out.read_bytes()
    assert (gzip.decompress(got) if mode == "rename_gz" else got) == want
    s = json.load(open(d / "s.json"))
    lens = np.array([len(r[1]) for r in recs])
    assert s["seqs_in"] == len(recs) and s["seqs_out"] == int(keep.sum()) and s["seqs_removed"] == int((keep == 0).sum())
    assert s["bp_in"] == int(lens.sum()) and s["bp_out"] == int(lens[keep == 1].sum()) and s["bp_removed"] == int(lens[keep == 0].sum())
    assert s["seqs_out_proportion"] == int(keep.sum()) / len(recs)
    assert s["index"] == str(d / "ref.idx") and s["input2"] is None and s["rename"] == (mode == "rename_gz")


def test_paired_files_split_outputs_and_interleaved_stdin(world):
    d = world["dir"]
    m1 = H.sample_reads(world["genome"], 3000, 150, 41)
    m2 = H.sample_reads(world["genome"], 3000, (100, 150), 42)
    r1, r2 = fastq_records(m1, b"p"), fastq_records(m2, b"p")
    (d / "r1.fq").write_bytes(fastq(r1))
    (d / "r2.fq.gz").write_bytes(gzip.compress(fastq(r2)))
    recs = [x for pair in zip(r1, r2) for x in pair]
    keep, hits, total = oracle_keep(world["idx"], [r[1] for r in recs], True, deplete=True)
    run("filter", "-d", d / "ref.idx", d / "r1.fq", d / "r2.fq.gz", "-o", d / "o1.fq", "-O", d / "o2.fq", "--batch-mbp", 1, "-s", d / "ps.json")
    w1, w2 = expected_output(recs, keep, True, split=True)
    assert (d / "o1.fq").read_bytes() == w1 and (d / "o2.fq").read_bytes() == w2
    s = json.load(open(d / "ps.json"))
    assert s["seqs_in"] == 6000 and s["seqs_out"] == 2 * int(keep.sum()) and s["output2"] == str(d / "o2.fq")
    # interleaved pairs on stdin ("-" "-"), interleaved on stdout (src/local_filter.rs:593-600, 697-700)
    p = run("filter", "-d", "-q", d / "ref.idx", "-", "-", stdin=fastq(recs))
    assert p.stdout == expected_output(recs, keep, True)[0]
    # a lone record cannot be paired
    p = run("filter", "-d", "-q", d / "ref.idx", "-", "-", stdin=fastq(recs[:5]), check=False)
    assert p.returncode == 1 and b"odd number" in p.stderr
    # files of unequal length
    (d / "short.fq").write_bytes(fastq(r2[:100]))
    p = run("filter", "-q", d / "ref.idx", d / "r1.fq", d / "short.fq", "-o", d / "x.fq", check=False)
    assert p.returncode == 1 and b"different numbers of records" in p.stderr


def test_fasta_reads_multiline_and_long_reads(world):
    """FASTA in -> FASTA out (newline-free sequence line); records longer than 1024 bases take the kernel's long path."""
    d = world["dir"]
    reads = H.sample_reads(world["genome"], 300, (500, 30000), 51, sub_rate=0.05)
    names = [b"ont_%d" % i for i in range(len(reads))]
    (d / "long.fa").write_bytes(fasta(reads, width=70, names=names))
    run("filter", d / "ref.idx", d / "long.fa", "-o", d / "long_out.fa", "-q", "-a", 5, "-r", 0.05)
    recs = [(n, bytes(r), None) for n, r in zip(names, reads)]
    keep, _, _ = oracle_keep(world["idx"], [r[1] for r in recs], False, abs_thr=5, rel_thr=0.05)
    assert 0 < keep.sum() < len(keep)
    assert (d / "long_out.fa").read_bytes() == expected_output(recs, keep, False)[0]


def test_debug_lines(world, gpu):
    """--debug: one line per single-end record with the matching k-mers (src/local_filter.rs:351-363); pairs only with hits and
    never with k-mers."""
    d = world["dir"]
    from deacon_server_b200 import IndexHeader
    gpu.index_upload(world["idx"].keys(), IndexHeader(2, 31, 15))
    reads = H.sample_reads(world["genome"], 60, (40, 300), 61)
    recs = fastq_records(reads)
    (d / "dbg.fq").write_bytes(fastq(recs))
    p = run("filter", "--debug", d / "ref.idx", d / "dbg.fq", "-o", d / "dbg_out.fq")
    want = [gpu.should_keep_sequence_debug(i.decode(), np.frombuffer(s, np.uint8))[4] for i, s, _ in recs]
    assert p.stderr.decode().splitlines() == want
    assert any("kmers=[A" in ln or "kmers=[C" in ln or "kmers=[G" in ln or "kmers=[T" in ln for ln in want)
    pr = [x for pair in zip(recs[0::2], recs[1::2]) for x in pair]
    keep, hits, total = oracle_keep(world["idx"], [r[1] for r in pr], True)
    p = run("filter", "--debug", d / "ref.idx", "-", "-", stdin=fastq(pr))
    want = [f"DEBUG: {pr[2 * u][0].decode()}/{pr[2 * u + 1][0].decode()} hits={hits[u]}/{total[u]} keep={'true' if keep[u] else 'false'} kmers=[]"
            for u in range(len(keep)) if hits[u] > 0]
    assert p.stderr.decode().splitlines() == want and p.stdout == expected_output(pr, keep, True)[0]


def test_index_union_diff_through_the_cli(world):
    """tests/index_tests.rs: union, diff index - index, diff index - FASTX (explicit and auto-detected parameters)."""
    from deacon_server_b200 import api
    d = world["dir"]
    g2 = H.random_genome(100_000, 77)
    other = [g2, world["genome"][50_000:90_000]]
    (d / "other.fa").write_bytes(fasta(other))
    run("index", "build", "-q", "-o", d / "other.idx", d / "other.fa")
    a = set(world["idx"].keys().tolist())
    b = set(O.index_build(other, 31, 15, threads=8).keys().tolist())

    def keys_of(path):
        return set(api.decode_index(path.read_bytes())[0].tolist())

    run("index", "union", "-o", d / "u.idx", d / "ref.idx", d / "other.idx")
    assert keys_of(d / "u.idx") == a | b
    run("index", "diff", "-o", d / "d.idx", d / "ref.idx", d / "other.idx")
    assert keys_of(d / "d.idx") == a - b
    for extra in ([], ["-k", 31, "-w", 15]):
        run("index", "diff", *extra, "-o", d / "dx.idx", d / "ref.idx", d / "other.fa")
        assert keys_of(d / "dx.idx") == a - b
    p = run("index", "diff", d / "ref.idx", d / "other.idx")   # "-": the index goes to stdout
    assert set(api.decode_index(p.stdout)[0].tolist()) == a - b
    p = run("index", "diff", "-k", 21, "-w", 11, d / "ref.idx", d / "other.fa", check=False)
    assert p.returncode == 1 and b"must match first index" in p.stderr
    run("index", "build", "-q", "-k", 21, "-w", 11, "-o", d / "k21.idx", d / "other.fa")
    p = run("index", "diff", d / "ref.idx", d / "k21.idx", check=False)
    assert p.returncode == 1 and b"Incompatible headers: second index has k=21, w=11, but first index has k=31, w=15" in p.stderr
    p = run("index", "union", d / "ref.idx", d / "k21.idx", check=False)
    assert p.returncode == 1 and b"Incompatible headers" in p.stderr


def test_batches_sharded_over_two_gpus_give_the_same_files(world):
    """--devices 0,1: batches dealt to one context per GPU (index replicated), written back in input order."""
    import deacon_server_b200 as d
    if d.load().dcn_device_count() < 2:
        pytest.skip("needs two GPUs")
    dd = world["dir"]
    reads = H.sample_reads(world["genome"], 20000, (50, 250), 71)
    (dd / "many.fq").write_bytes(fastq(fastq_records(reads)))
    run("filter", "-d", "-q", dd / "ref.idx", dd / "many.fq", "-o", dd / "g1.fq", "--batch-mbp", 1, "--devices", 0)
    run("filter", "-d", "-q", dd / "ref.idx", dd / "many.fq", "-o", dd / "g2.fq", "--batch-mbp", 1, "--devices", "0,1")
    assert (dd / "g1.fq").read_bytes() == (dd / "g2.fq").read_bytes()


def test_empty_input_and_a_record_larger_than_a_batch(world):
    d = world["dir"]
    (d / "empty.fq").write_bytes(b"")
    run("filter", "-q", d / "ref.idx", d / "empty.fq", "-o", d / "empty_out.fq", "-s", d / "empty.json")
    s = json.load(open(d / "empty.json"))
    assert (d / "empty_out.fq").read_bytes() == b"" and s["seqs_in"] == 0 and s["bp_in"] == 0 and s["seqs_out_proportion"] == 0.0
    # one 2.5 Mbp record (a stretch of the reference with 3 % substitutions) between short ones, batches of 1 Mbp
    big = np.concatenate([world["genome"]] * 7)[:2_500_000].copy()
    rng = np.random.default_rng(5)
    m = rng.random(len(big)) < 0.03
    big[m] = H.ACGT[rng.integers(0, 4, int(m.sum()))]
    reads = H.sample_reads(world["genome"], 50, 150, 81)
    seqs = reads[:25] + [big] + reads[25:]
    names = [b"r%d" % i for i in range(len(seqs))]
    (d / "mixed.fa").write_bytes(fasta(seqs, width=60, names=names))
    run("filter", "-q", d / "ref.idx", d / "mixed.fa", "-o", d / "mixed_out.fa", "--batch-mbp", 1, "-r", 0.02)
    recs = [(n, bytes(r), None) for n, r in zip(names, seqs)]
    keep, hits, total = oracle_keep(world["idx"], [r[1] for r in recs], False, rel_thr=0.02)
    assert keep[25] == 1 and total[25] > 100_000
    assert (d / "mixed_out.fa").read_bytes() == expected_output(recs, keep, False)[0]


def test_missing_index_and_unreadable_inputs(world):
    d = world["dir"]
    p = run("filter", d / "nope.idx", d / "reads.fq", check=False)
    assert p.returncode == 1 and b"Failed to open index file" in p.stderr
    (d / "junk.idx").write_bytes(b"\x07\x1f\x0fnot an index")
    p = run("filter", d / "junk.idx", d / "reads.fq", check=False)
    assert p.returncode == 1 and b"Unsupported index format version" in p.stderr
